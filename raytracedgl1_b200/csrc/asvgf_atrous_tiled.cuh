// asvgf_atrous_tiled.cuh -- the production a-trous kernel (all iterations), sm_100a.
//
// Same arithmetic as CmSVGFAtrous_Iter0.comp:158-335 / CmSVGFAtrous.comp:65-349 (see k_atrous in asvgf_kernels.cuh for
// the literal form), restructured for the B200 SM, where the pass is instruction-issue bound (ncu: the literal form
// runs 2200 instructions per pixel with DRAM at 6 %; profiles/r01_v1_*, r01_v2_*):
//
//  * DILATED TILE.  For stride s a pixel only ever touches rows y, y+-s.  A CTA therefore owns 64 contiguous columns of
//    8 rows spaced s apart (rows y0+r, y0+r+s, ...; the s residues r go to s different CTAs).  Its halo is one dilated
//    row above/below and s columns left/right, so the read amplification is (10/8)*((64+2s)/64) = 1.29..1.56 for
//    every stride instead of the (1+2s/T)^2 of a square tile, and every global access is still a full-line row segment.
//    (64 x 8 keeps a tile at 41-51 KB of shared memory and 64 registers per thread: four CTAs = 32 warps per SM.)
//  * DECODE ONCE.  Each tile pixel is unpacked once on the way into shared memory -- normal through the exact sin/cos
//    table, E5B9G9R9 and fp16 to fp32, luminance -- into four float4 planes, so a tap costs four conflict-free
//    LDS.128 and no conversions.  The stride is a template parameter: every tap address is an immediate offset.
//  * WEIGHTS THROUGH ONE LOG.  w_n = (n.n')^128 and the specular lobe (n.n')^P share L = lg2(n.n'), and the wavelet
//    kernel weight / inside-the-render-area test is folded into the exponent (lk = log2 k, or -inf outside):
//      wBase = ex2(128 L + lk - wz^2 log2e),  wDiff = ex2(128 L + lk - (wz^2 + wl) log2e),  wSpec = r * ex2(. + P L)
//    i.e. 1 lg2 + 3 ex2 on the MUFU pipe and ~40 FMA-pipe instructions per tap instead of two powf and two expf.
//    Relative weight error <= ~3e-4 (|P|<=1152 times the 2^-22 absolute error of lg2.approx), inside the 1e-3 output
//    tolerance because the weights are normalised; integer/boolean outputs do not depend on this kernel.
#pragma once

#include "asvgf_fast.cuh"

#include <type_traits>

namespace asvgf
{

// Tile height / CTAs per SM.  Measured at 4K (ms per iteration, plain / composing), one tile per CTA:
//   16 rows, 2 CTAs/SM (128 regs, 16 warps/SM): 0.22 / 0.25      8 rows, 3 CTAs/SM (80 regs): 0.21 / 0.23
//    8 rows, 4 CTAs/SM ( 64 regs, 32 warps/SM): 0.19 / 0.23     16 rows, 2 x 512 threads (64 regs): 0.20 / 0.23
// The pass is bound by issue latency (a long dependent chain per tap), so the warps per SM matter more than the 11 %
// of extra halo rows the shorter tile stages.
#ifndef ASVGF_AT_TILE_H
    #define ASVGF_AT_TILE_H 8
#endif
#ifndef ASVGF_AT_THREADS
    #define ASVGF_AT_THREADS 256
#endif
#ifndef ASVGF_AT_MIN_CTAS
    #define ASVGF_AT_MIN_CTAS 4
#endif
#ifndef ASVGF_AT_MULTI_TILE
    #define ASVGF_AT_MULTI_TILE 0 // 1: a CTA walks a.tilesPerCta tiles with the next tile's loads in flight (measured equal to one tile per CTA)
#endif
constexpr bool kAtMultiTile = ASVGF_AT_MULTI_TILE != 0;
constexpr int kAtTileW   = 64;               // columns per CTA
constexpr int kAtTileH   = ASVGF_AT_TILE_H;  // dilated rows per CTA
constexpr int kAtThreads = ASVGF_AT_THREADS; // 4 pixels per thread
constexpr int kAtBytesPerTilePixel = 64;

ASVGF_DEVICE float ex2Approx( float x ) { return fex2( x ); }
// exp2 that keeps denormal results (MUFU.EX2 flushes them): a vanishing weight must stay non-zero exactly as long as
// the shader's exp() does, because an infinite fp16 variance times a zero weight is NaN but times a denormal weight is inf
ASVGF_DEVICE float ex2ApproxDenorm( float x )
{
    const bool  tiny = x < -126.0f;
    const float r    = fex2( tiny ? x * 0.5f : x );
    return tiny ? r * r : r;
}

// B10G11R11_UFLOAT <-> fp32 through the fp16 bit layout (a uf11 is the top 11 bits of a positive half, a uf10 the top 10)
ASVGF_DEVICE float3 unpackUfloat111110Fast( uint32_t p )
{
    return make_float3( h2f( ( uint16_t )( ( p & 0x7ffu ) << 4 ) ), h2f( ( uint16_t )( ( ( p >> 11 ) & 0x7ffu ) << 4 ) ), h2f( ( uint16_t )( ( p >> 22 ) << 5 ) ) );
}
template< uint32_t MBITS > ASVGF_DEVICE uint32_t floatToUfloatFast( float f ) // RTNE; NaN, negatives -> 0; overflow -> max finite
{
    constexpr uint32_t maxCode = ( 30u << MBITS ) | ( ( 1u << MBITS ) - 1u );
    constexpr uint32_t drop    = 23u - MBITS;
    if( !( f > 0.0f ) ) return 0u;
    if( f < 6.103515625e-05f ) return __float2uint_rn( f * ( float )( 1u << ( 14u + MBITS ) ) ); // below 2^-14: m * 2^(-14-MBITS)
    const uint32_t x = __float_as_uint( f );
    const uint32_t r = ( x - 0x38000000u + ( ( 1u << ( drop - 1u ) ) - 1u ) + ( ( x >> drop ) & 1u ) ) >> drop;
    return min( r, maxCode );
}
ASVGF_DEVICE uint32_t packUfloat111110Fast( float3 c ) { return floatToUfloatFast< 6 >( c.x ) | ( floatToUfloatFast< 6 >( c.y ) << 11 ) | ( floatToUfloatFast< 5 >( c.z ) << 22 ); }

// raw (still packed) texels of one tile pixel and of one output pixel's centre-only inputs: what a thread keeps in
// registers while the loads of the NEXT tile are in flight behind the arithmetic of the current one
// (no arithmetic may touch a loaded value before the tile is committed: a use would stall the thread on the load)
struct AtrousRaw { uint32_t n, s, i; uint2 d; uint16_t depth, mr; };
template< bool ITER0, bool LAST > struct AtrousCentre { uint2 histLen, tp; uint32_t albedo; uint16_t gradDepth, sigma, metRough; uint8_t sky; };

template< int STEP, bool ITER0, bool LAST >
__global__ void __launch_bounds__( kAtThreads, ASVGF_AT_MIN_CTAS ) k_atrous_tiled( const __grid_constant__ AtrousArgs a, const __grid_constant__ Uniform u )
{
    extern __shared__ __align__( 16 ) uint8_t smemRaw[];
    constexpr int TW    = kAtTileW + 2 * STEP; // tile width incl. halo columns
    constexpr int TROWS = kAtTileH + 2;
    constexpr int NPX   = TW * TROWS;
    constexpr int NIT   = ( NPX + kAtThreads - 1 ) / kAtThreads;
    constexpr int PIX   = ( kAtTileW * kAtTileH ) / kAtThreads;
    float4* sA = reinterpret_cast< float4* >( smemRaw ); // normal.xyz, depth
    float4* sB = sA + NPX;                               // diffuse rgb, variance
    float4* sC = sB + NPX;                               // specular rgb, luminance of the diffuse rgb
    float4* sD = sC + NPX;                               // indirect rgb, roughness

    // dilated rows: blockIdx.y enumerates (band, residue); row(j) = yBase + j*STEP, j = -1 .. kAtTileH.
    // A CTA walks a.tilesPerCta consecutive 64-column tiles of that row set; while it filters tile t from shared
    // memory, the global loads of tile t+1 are already in flight into registers (software pipeline), so only the first
    // tile's HBM latency is exposed.
    const int band  = blockIdx.y / STEP, res = blockIdx.y - band * STEP;
    const int yBase = u.yBegin + band * ( kAtTileH * STEP ) + res;
    const int nTilesCta = kAtMultiTile ? a.tilesPerCta : 1;
    const int xFirst = blockIdx.x * nTilesCta * kAtTileW;
    const Rows rows = rowsOf( a.depth );

    AtrousRaw raw[ NIT ];
    AtrousCentre< ITER0, LAST > cen[ PIX ], cenNext[ kAtMultiTile ? PIX : 1 ];

    auto issue = [ & ]( int xBase, AtrousCentre< ITER0, LAST >( &cc )[ PIX ] )
    {
#pragma unroll
        for( int k = 0; k < PIX; k++ )
        {
            const int  p  = threadIdx.x + k * kAtThreads;
            const int  j  = p / kAtTileW, c = p - j * kAtTileW;
            const int  px = xBase + c, py = yBase + j * STEP;
            const bool ok = px < u.W && py < u.yEnd;
            cc[ k ].gradDepth = ldIf< uint16_t >( a.depthGrad, px, py, ok, ( uint16_t )0 );
            cc[ k ].histLen   = ldIf< uint2 >( a.histLen, px, py, ok, make_uint2( 0u, 0u ) );
            if( !ITER0 ) cc[ k ].sigma = ldIf< uint16_t >( a.filteredVariance, px, py, ok, ( uint16_t )0 );
            if( LAST )
            {
                const int sepx = u.W / 2, isOdd = px >= sepx ? 1 : 0;
                const int regX = ( px - isOdd * sepx ) * 2 + ( ( isOdd + py ) & 1 ); // ShaderCommonGLSLFunc.h:331-342
                cc[ k ].albedo   = ldIf< uint32_t >( a.albedo, regX, py, ok, 0u );
                cc[ k ].tp       = ldIf< uint2 >( a.throughput, px, py, ok, make_uint2( 0u, 0u ) );
                cc[ k ].sky      = ldIf< uint8_t >( a.isSky, px, py, ok, ( uint8_t )0 );
                cc[ k ].metRough = ldIf< uint16_t >( a.metRough, px, py, ok, ( uint16_t )0 );
            }
        }
#pragma unroll
        for( int it = 0; it < NIT; it++ )
        {
            const int  e = threadIdx.x + it * kAtThreads;
            const int  r = e / TW, c = e - r * TW;
            const int  gx = xBase - STEP + c, gy = yBase + ( r - 1 ) * STEP;
            const bool ok = e < NPX && rows.in( gx, gy );
            raw[ it ].n  = ldIf< uint32_t >( a.normal, gx, gy, ok, 0u );
            raw[ it ].depth = ldIf< uint16_t >( a.depth, gx, gy, ok, ( uint16_t )0 );
            raw[ it ].mr    = ldIf< uint16_t >( a.metRough, gx, gy, ok, ( uint16_t )0 );
            raw[ it ].d  = ldIf< uint2 >( a.srcDiff, gx, gy, ok, make_uint2( 0u, 0u ) );
            raw[ it ].s  = ldIf< uint32_t >( a.srcSpec, gx, gy, ok, 0u );
            raw[ it ].i  = ldIf< uint32_t >( a.srcIndir, gx, gy, ok, 0u );
        }
    };
    // unpack every tile pixel once into the four float4 planes
    auto commit = [ & ]()
    {
#pragma unroll
        for( int it = 0; it < NIT; it++ )
        {
            const int e = threadIdx.x + it * kAtThreads;
            if( e >= NPX ) break;
            const float3 n  = decodeNormal( a.lut, raw[ it ].n );
            const float2 rg = __half22float2( *reinterpret_cast< const __half2* >( &raw[ it ].d.x ) );
            const float2 bv = __half22float2( *reinterpret_cast< const __half2* >( &raw[ it ].d.y ) );
            const float3 sp = decodeE5( raw[ it ].s ), in = decodeE5( raw[ it ].i );
            sA[ e ] = make_float4( n.x, n.y, n.z, h2f( raw[ it ].depth ) );
            sB[ e ] = make_float4( rg.x, rg.y, bv.x, bv.y );
            sC[ e ] = make_float4( sp.x, sp.y, sp.z, luminance( make_float3( rg.x, rg.y, bv.x ) ) );
            sD[ e ] = make_float4( in.x, in.y, in.z, unorm8ToFloatFast( ( uint32_t )raw[ it ].mr >> 8 ) );
        }
    };

    issue( xFirst, cen );
  for( int t = 0; t < nTilesCta; t++ )
  {
    const int xBase = xFirst + t * kAtTileW;
    if( xBase >= u.W ) break;
    if( kAtMultiTile && t > 0 )
    {
        __syncthreads(); // every warp is done reading the previous tile
#pragma unroll
        for( int k = 0; k < PIX; k++ ) cen[ k ] = cenNext[ kAtMultiTile ? k : 0 ];
    }
    commit();
    __syncthreads();
    if constexpr( kAtMultiTile )
    {
        if( t + 1 < nTilesCta && xBase + kAtTileW < u.W ) issue( xBase + kAtTileW, cenNext ); // in flight during the arithmetic below
    }

    const int       sep   = u.W / 2;
    constexpr float LOG2E = 1.4426950408889634f;
    const float     NINF  = __uint_as_float( 0xff800000u );

#pragma unroll
    for( int k = 0; k < PIX; k++ )
    {
        const int p  = threadIdx.x + k * kAtThreads;
        const int j  = p / kAtTileW, c = p - j * kAtTileW;
        const int px = xBase + c, py = yBase + j * STEP;
        if( px >= u.W || py >= u.yEnd ) continue;

        const int    ce = ( j + 1 ) * TW + ( c + STEP );
        const float4 cA = sA[ ce ], cB = sB[ ce ], cC = sC[ ce ], cD = sD[ ce ];
        float3 outDiff = make_float3( cB.x, cB.y, cB.z ), outSpec = make_float3( cC.x, cC.y, cC.z ), outIndir = make_float3( cD.x, cD.y, cD.z );
        float  outVariance = cB.w;

        const float depth     = cA.w;
        const float roughness = cD.w;
        const bool  invalid   = depth < 0.0f || depth > 10000.0f;

        if( invalid )
        {
            if( ITER0 )
            {
                outVariance = 0.0f;
                outDiff = outSpec = outIndir = make_float3( 0.0f, 0.0f, 0.0f );
            }
        }
        else
        {
            const float gradDepth = h2f( cen[ k ].gradDepth );
            const float l         = cC.w;

            float sigmaL;
            if( ITER0 ) // Iter0.comp:158-187: 3x3 Gaussian {1/4,1/8,1/16} of the variance (products by powers of two are exact)
            {
                float r = 0.0f;
#pragma unroll
                for( int yy = -1; yy <= 1; yy++ )
#pragma unroll
                    for( int xx = -1; xx <= 1; xx++ )
                        r = fmaf( sB[ ce + yy * TW + xx ].w, ( xx == 0 ? 0.5f : 0.25f ) * ( yy == 0 ? 0.5f : 0.25f ), r );
                r      = __fsqrt_rn( fmaxf( r, 0.0f ) );
                *atW< uint16_t >( a.filteredVariance, px, py ) = f2h( r );
                sigmaL = r;
            }
            else
            {
                sigmaL = h2f( cen[ k ].sigma );
            }
            const float kLum = LOG2E * frcp( fmaf( 4.0f, sigmaL, 0.00001f ) );

            const float wRoughMultiplier  = saturatef( fmaf( roughness, 30.0f, -( float )a.iteration ) );
            const float wR10              = wRoughMultiplier * 10.0f;
            const float historyLengthSpec = __low2float( *reinterpret_cast< const __half2* >( &cen[ k ].histLen.y ) );
            const float alpha             = roughness * roughness;
            float       P                 = fmaxf( 0.01f, fmaf( 2.0f, frcp( fmaf( alpha, alpha, 1e-4f ) ), -2.0f ) ); // BRDF.h:27-29
            P                             = clampf( P, 8.0f, 1024.0f ) * saturatef( historyLengthSpec * 0.125f );
            // 1/max(g,.01) = min(1/g,100) and 1/max(2g,.01) = min(.5/g,100) from one MUFU (g >= 0; 1/0 = inf), pre-scaled by
            // sqrt(log2 e) so that -wZ*wZ is already the base-2 exponent
            constexpr float SQRT_LOG2E = 1.2011224087864498f;
            const float rgd   = frcp( gradDepth );
            const float invK1 = SQRT_LOG2E * fminf( rgd, 100.0f ), invK2 = SQRT_LOG2E * fminf( 0.5f * rgd, 100.0f );

            // taps outside the checkerboarded render area (ShaderCommonGLSLFunc.h:345-366: they never cross x = W/2 or
            // leave the frame) get exponent -inf; interior pixels take the loop whose wavelet weights are immediates
            const int  ax0 = px >= sep ? sep : 0, ax1 = ax0 + sep;
            const bool okL = px - STEP >= ax0, okR = px + STEP < ax1, okT = py - STEP >= 0, okB = py + STEP < u.H;
            const float PINF = __uint_as_float( 0x7f800000u );

            float weightSum = 1.0f, weightSumSpec = 1.0f, weightSumIndir = 1.0f;

            auto taps = [ & ]( auto interiorTag )
            {
                constexpr bool INTERIOR = decltype( interiorTag )::value;
#pragma unroll
                for( int yy = -1; yy <= 1; yy++ )
                {
#pragma unroll
                    for( int xx = -1; xx <= 1; xx++ )
                    {
                        if( xx == 0 && yy == 0 ) continue;
                        const int    qe = ce + yy * TW + xx * STEP;
                        const float4 qA = sA[ qe ], qB = sB[ qe ], qC = sC[ qe ], qD = sD[ qe ];

                        const float nN = fmaxf( 0.0f, fmaf( cA.z, qA.z, fmaf( cA.y, qA.y, cA.x * qA.x ) ) );
                        const float L  = fmaxf( flg2( nN ), -126.0f );
                        const float wZ = fabsf( depth - qA.w ) * ( ( xx != 0 && yy != 0 ) ? invK2 : invK1 );
                        float       lk = ( xx != 0 && yy != 0 ) ? -2.0f : -1.0f; // log2 of the wavelet weight
                        if( !INTERIOR )
                        {
                            const bool in = ( xx < 0 ? okL : xx > 0 ? okR : true ) && ( yy < 0 ? okT : yy > 0 ? okB : true );
                            lk            = in ? lk : NINF;
                        }
                        const float e0    = fmaf( 128.0f, L, lk );
                        const float eBase = fmaf( -wZ, wZ, e0 );
                        const float eDiff = fmaf( fabsf( l - qC.w ), -kLum, eBase );
                        const float wBase = fex2( eBase );
                        const float wDiff = fex2( eDiff );
                        float       rfac; // max(0, 1 - 10|dr|) * wRough, both factors in [0,1]
                        asm( "fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"( rfac ) : "f"( fabsf( roughness - qD.w ) ), "f"( -wR10 ), "f"( wRoughMultiplier ) );
                        const float wSpec = rfac * fex2( fmaf( P, L, eBase ) );

                        outDiff.x = fmaf( qB.x, wDiff, outDiff.x ); outDiff.y = fmaf( qB.y, wDiff, outDiff.y ); outDiff.z = fmaf( qB.z, wDiff, outDiff.z );
                        outSpec.x = fmaf( qC.x, wSpec, outSpec.x ); outSpec.y = fmaf( qC.y, wSpec, outSpec.y ); outSpec.z = fmaf( qC.z, wSpec, outSpec.z );
                        outIndir.x = fmaf( qD.x, wBase, outIndir.x ); outIndir.y = fmaf( qD.y, wBase, outIndir.y ); outIndir.z = fmaf( qD.z, wBase, outIndir.z );
                        weightSum += wDiff; weightSumSpec += wSpec; weightSumIndir += wBase;
                        // (var*w)*w: MUFU.EX2 flushes a denormal weight to zero where the shader's exp() keeps it, which only
                        // shows when the tap's fp16 variance is +inf (inf*denormal = inf, inf*0 = NaN): only that case pays
                        // for the denormal-preserving exponential
                        if( qB.w == PINF )
                        {
                            const float wd = ex2ApproxDenorm( eDiff );
                            outVariance    = fmaf( __fmul_rn( qB.w, wd ), wd, outVariance );
                        }
                        else
                        {
                            outVariance = fmaf( __fmul_rn( qB.w, wDiff ), wDiff, outVariance );
                        }
                    }
                }
            };
            if( okL && okR && okT && okB ) taps( std::true_type{} );
            else taps( std::false_type{} );
            const float inv = frcp( weightSum ), invSpec = frcp( weightSumSpec ), invIndir = frcp( weightSumIndir );
            outDiff     = make_float3( outDiff.x * inv, outDiff.y * inv, outDiff.z * inv );
            outVariance = outVariance * ( inv * inv );
            outSpec     = make_float3( outSpec.x * invSpec, outSpec.y * invSpec, outSpec.z * invSpec );
            outIndir    = make_float3( outIndir.x * invIndir, outIndir.y * invIndir, outIndir.z * invIndir );
        }

        if( !LAST )
        {
            const __half2 o0 = __floats2half2_rn( outDiff.x, outDiff.y ), o1 = __floats2half2_rn( outDiff.z, outVariance );
            *atW< uint2 >( a.dstDiff, px, py ) = make_uint2( *reinterpret_cast< const uint32_t* >( &o0 ), *reinterpret_cast< const uint32_t* >( &o1 ) );
            *atW< uint32_t >( a.dstSpec, px, py )  = encodeE5Fast( outSpec );
            *atW< uint32_t >( a.dstIndir, px, py ) = encodeE5Fast( outIndir );
            continue;
        }

        // composition, CmSVGFAtrous.comp:278-348
        const int     isOdd = px >= sep ? 1 : 0;
        const int     regX  = ( px - isOdd * sep ) * 2 + ( ( isOdd + py ) & 1 ); // ShaderCommonGLSLFunc.h:331-342
        float3        albedo = unpackUfloat111110Fast( cen[ k ].albedo );
        const uint2   tp     = cen[ k ].tp;
        const float2  tp01   = __half22float2( *reinterpret_cast< const __half2* >( &tp.x ) );
        const float3  throughput = make_float3( tp01.x, tp01.y, __low2float( *reinterpret_cast< const __half2* >( &tp.y ) ) );
        const uint32_t sky   = cen[ k ].sky;
        if( ( u.debugShowFlags & ( 1u << 9 ) ) != 0 ) albedo = make_float3( 1.0f, 1.0f, 1.0f );

        float3 illuminated;
        if( sky == 0u )
        {
            const float metallic = unorm8ToFloatFast( ( uint32_t )cen[ k ].metRough & 0xffu );
            const float t  = saturatef( ( roughness - 0.5f ) * 4.0f );            // smoothstep(0.5, 0.75, roughness)
            const float sm = t * t * ( 3.0f - 2.0f * t );
            const float3 specular = fmix3( outSpec, outIndir, sm );
            const float3 roS = fmix3( make_float3( 0.04f, 0.04f, 0.04f ), albedo, metallic ); // BRDF.h:49-53
            const float  om  = 1.0f - metallic;
            illuminated = make_float3( fmaf( ( outDiff.x + outIndir.x ) * albedo.x, om, specular.x * roS.x ), fmaf( ( outDiff.y + outIndir.y ) * albedo.y, om, specular.y * roS.y ),
                                       fmaf( ( outDiff.z + outIndir.z ) * albedo.z, om, specular.z * roS.z ) );
            if( ( u.debugShowFlags & ( 1u << 5 ) ) != 0 ) illuminated = make_float3( outDiff.x * albedo.x, outDiff.y * albedo.y, outDiff.z * albedo.z );
            else if( ( u.debugShowFlags & ( 1u << 6 ) ) != 0 ) illuminated = make_float3( specular.x * roS.x, specular.y * roS.y, specular.z * roS.z );
            else if( ( u.debugShowFlags & ( 1u << 7 ) ) != 0 ) illuminated = outIndir;
            const float la = flum( albedo ); // BRDF.h:57-64
            float       ambient = 1.0f;
            if( !( la > 0.02f ) ) { const float q = ( la - 0.02f ) * 50.0f; ambient = 1.0f - q * q; }
            illuminated = make_float3( illuminated.x * throughput.x * ambient, illuminated.y * throughput.y * ambient, illuminated.z * throughput.z * ambient );
        }
        else
        {
            illuminated = make_float3( albedo.x * throughput.x, albedo.y * throughput.y, albedo.z * throughput.z );
        }
        *atW< uint16_t >( a.histogramInput, regX, py ) = f2h( fmaxf( 0.0f, flum( illuminated ) ) );
        if( ( float )regX < mulr( u.lightmapScreenCoverage, u.renderWidth ) ) illuminated = make_float3( albedo.x * throughput.x, albedo.y * throughput.y, albedo.z * throughput.z ); // classicShading
        *atW< uint32_t >( a.preFinal, regX, py ) =
            packUfloat111110Fast( make_float3( fmaxf( 0.0f, illuminated.x ), fmaxf( 0.0f, illuminated.y ), fmaxf( 0.0f, illuminated.z ) ) );
    }
  } // tiles of this CTA
}

constexpr size_t atrousTiledSmemBytes( int step ) { return ( size_t )( kAtTileW + 2 * step ) * ( kAtTileH + 2 ) * kAtBytesPerTilePixel; }

} // namespace asvgf
