// asvgf_atrous_fp32.cuh -- EXPERIMENT, not part of libasvgf.so: the a-trous kernel on four fp32 planes (64 B per staged texel), the
// production kernel until the packed tile of asvgf_atrous_half.cuh replaced it (1.3865 -> 1.3238 ms per 4K frame, identical
// results), with its two-pixel PAIR form.  Kept as written when it was retired; it builds against asvgf_atrous_tiled.cuh
// (asvgf.cu would include this file and launch k_atrous_tiled< STEP, ITER0, LAST, PAIR, CENTRE > with atrousSmemBytes( step )).
#pragma once

#include "../raytracedgl1_b200/csrc/asvgf_atrous_tiled.cuh"

namespace asvgf
{

// Staging of one tile.  UNI: the guide plane and the three colour images hold the same number of texels per row (the library's
// own allocation): one texel offset serves the four loads of a texel.  INTERIOR: every staged texel lies inside the image and the
// stored rows (all tiles but the frame / band border): no bounds tests, no zero fill.  Both are uniform per CTA; the general form
// (own pitch per image, every texel tested) is the third copy.  (ncu / SASS of the single general form: 90 instructions per staged
// texel at STEP < 4 and 66 at STEP >= 4, of which 37 decode, 8 move data and the rest was addressing, bounds and zero fill.)
template< int STEP, bool UNI, bool INTERIOR >
ASVGF_DEVICE void atrousStageImpl( const AtrousArgs& a, const Uniform& u, const Rows& rows, int xBase, int yBase, float4* sA, float4* sB, float4* sC, float4* sD )
{
    using G = AtrousGeom< STEP >;
    const uint32_t pitchT = a.guide.pitchT;
    // texel offset of staged (row 0, column 0); wraps for tiles that start outside the image -- only dereferenced under `ok`
    const uint32_t off0 = ( uint32_t )( ( yBase - STEP ) * ( int )pitchT + xBase - G::HALO );
    const float4*   bG = reinterpret_cast< const float4* >( a.guide.p0 );
    const uint2*    bD = reinterpret_cast< const uint2* >( a.srcDiff.p0 );
    const uint32_t* bS = reinterpret_cast< const uint32_t* >( a.srcSpec.p0 );
    const uint32_t* bI = reinterpret_cast< const uint32_t* >( a.srcIndir.p0 );
    if constexpr( G::VEC )
    {
        constexpr int NGRP = G::GROUPS * G::TROWS;
#pragma unroll
        for( int g0 = 0; g0 < NGRP; g0 += kAtThreads )
        {
            const int g = g0 + threadIdx.x;
            if( NGRP % kAtThreads != 0 && g >= NGRP ) break;
            const int  r = g / G::GROUPS, cg = g - r * G::GROUPS;
            const int  gx = xBase - G::HALO + 4 * cg, gy = yBase + ( r - 1 ) * STEP;
            const bool rowOk = INTERIOR || ( gy >= rows.y0 && gy < rows.y1 );
            const bool ok01 = INTERIOR || ( rowOk && gx >= 0 && gx < u.W ), ok23 = INTERIOR || ( rowOk && gx >= 0 && gx + 2 < u.W ); // gx is a multiple of 4, W is even
            float4 ga[ 4 ];
            uint4  d01, d23, sp, in;
            {
                const uint32_t off = off0 + ( uint32_t )r * ( ( uint32_t )STEP * pitchT ) + 4u * ( uint32_t )cg;
                const float4*  pg = UNI ? bG + off : at< float4 >( a.guide, gx, gy );
                const uint4*   pd = UNI ? reinterpret_cast< const uint4* >( bD + off ) : reinterpret_cast< const uint4* >( at< uint2 >( a.srcDiff, gx, gy ) );
                const uint4*   ps = UNI ? reinterpret_cast< const uint4* >( bS + off ) : reinterpret_cast< const uint4* >( at< uint32_t >( a.srcSpec, gx, gy ) );
                const uint4*   pi = UNI ? reinterpret_cast< const uint4* >( bI + off ) : reinterpret_cast< const uint4* >( at< uint32_t >( a.srcIndir, gx, gy ) );
                if constexpr( INTERIOR )
                {
                    ga[ 0 ] = __ldg( pg ); ga[ 1 ] = __ldg( pg + 1 ); ga[ 2 ] = __ldg( pg + 2 ); ga[ 3 ] = __ldg( pg + 3 );
                    d01 = __ldg( pd ); d23 = __ldg( pd + 1 ); sp = __ldg( ps ); in = __ldg( pi );
                }
                else
                {
                    const float4 z4 = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
                    const uint4  z  = make_uint4( 0u, 0u, 0u, 0u );
                    ga[ 0 ] = ok01 ? __ldg( pg ) : z4;
                    ga[ 1 ] = ok01 ? __ldg( pg + 1 ) : z4;
                    ga[ 2 ] = ok23 ? __ldg( pg + 2 ) : z4;
                    ga[ 3 ] = ok23 ? __ldg( pg + 3 ) : z4;
                    d01 = ok01 ? __ldg( pd ) : z;
                    d23 = ok23 ? __ldg( pd + 1 ) : z;
                    sp  = ok01 ? __ldg( ps ) : z;
                    in  = ok01 ? __ldg( pi ) : z;
                }
            }
            const int      e0 = r * G::RP + cg; // texel k of the group -> quarter row k
            const uint32_t dd[ 8 ] = { d01.x, d01.y, d01.z, d01.w, d23.x, d23.y, d23.z, d23.w };
            const uint32_t ss[ 4 ] = { sp.x, sp.y, sp.z, sp.w }, ii[ 4 ] = { in.x, in.y, in.z, in.w };
#pragma unroll
            for( int k = 0; k < 4; k++ )
            {
                const float2 rg = __half22float2( *reinterpret_cast< const __half2* >( &dd[ 2 * k ] ) );
                const float2 bv = __half22float2( *reinterpret_cast< const __half2* >( &dd[ 2 * k + 1 ] ) );
                const float3 s3 = decodeE5( ss[ k ] ), i3 = decodeE5( ii[ k ] );
                sA[ e0 + k * G::GP ] = make_float4( ga[ k ].x, ga[ k ].y, ga[ k ].z, guideDepth( ga[ k ].w ) );
                sB[ e0 + k * G::GP ] = make_float4( rg.x, rg.y, bv.x, bv.y );
                sC[ e0 + k * G::GP ] = make_float4( s3.x, s3.y, s3.z, luminance( make_float3( rg.x, rg.y, bv.x ) ) );
                sD[ e0 + k * G::GP ] = make_float4( i3.x, i3.y, i3.z, guideRoughByte( ga[ k ].w ) );
            }
        }
    }
    else
    {
        constexpr int NIT = ( G::NPX + kAtThreads - 1 ) / kAtThreads;
        float4   ga[ NIT ];
        uint2    dd[ NIT ];
        uint32_t ss[ NIT ], ii[ NIT ];
#pragma unroll
        for( int it = 0; it < NIT; it++ ) // all loads first: one memory round trip for the tile
        {
            const int  e = threadIdx.x + it * kAtThreads;
            const int  r = e / G::TW, c = e - r * G::TW;
            const int  gx = xBase - G::HALO + c, gy = yBase + ( r - 1 ) * STEP;
            const bool ok = ( ( it + 1 ) * kAtThreads <= G::NPX || e < G::NPX ) && ( INTERIOR || rows.in( gx, gy ) ); // (threadIdx.x < kAtThreads)
            const uint32_t off = off0 + ( uint32_t )r * ( ( uint32_t )STEP * pitchT ) + ( uint32_t )c;
            ga[ it ] = ok ? __ldg( UNI ? bG + off : at< float4 >( a.guide, gx, gy ) ) : make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
            dd[ it ] = ok ? __ldg( UNI ? bD + off : at< uint2 >( a.srcDiff, gx, gy ) ) : make_uint2( 0u, 0u );
            ss[ it ] = ok ? __ldg( UNI ? bS + off : at< uint32_t >( a.srcSpec, gx, gy ) ) : 0u;
            ii[ it ] = ok ? __ldg( UNI ? bI + off : at< uint32_t >( a.srcIndir, gx, gy ) ) : 0u;
        }
#pragma unroll
        for( int it = 0; it < NIT; it++ )
        {
            const int e = threadIdx.x + it * kAtThreads;
            if( ( it + 1 ) * kAtThreads > G::NPX && e >= G::NPX ) break;
            const float2 rg = __half22float2( *reinterpret_cast< const __half2* >( &dd[ it ].x ) );
            const float2 bv = __half22float2( *reinterpret_cast< const __half2* >( &dd[ it ].y ) );
            const float3 s3 = decodeE5( ss[ it ] ), i3 = decodeE5( ii[ it ] );
            sA[ e ] = make_float4( ga[ it ].x, ga[ it ].y, ga[ it ].z, guideDepth( ga[ it ].w ) );
            sB[ e ] = make_float4( rg.x, rg.y, bv.x, bv.y );
            sC[ e ] = make_float4( s3.x, s3.y, s3.z, luminance( make_float3( rg.x, rg.y, bv.x ) ) );
            sD[ e ] = make_float4( i3.x, i3.y, i3.z, guideRoughByte( ga[ it ].w ) );
        }
    }
}

template< int STEP >
ASVGF_DEVICE void atrousStage( const AtrousArgs& a, const Uniform& u, const Rows& rows, int xBase, int yBase, float4* sA, float4* sB, float4* sC, float4* sD )
{
    using G = AtrousGeom< STEP >;
#ifdef ASVGF_AT_GENERAL_STAGING
    atrousStageImpl< STEP, false, false >( a, u, rows, xBase, yBase, sA, sB, sC, sD );
#else
    // staged rows yBase - STEP .. yBase + kAtTileH * STEP, staged columns xBase - HALO .. xBase + kAtTileW + HALO - 1
    const bool interior = xBase - G::HALO >= 0 && xBase + kAtTileW + G::HALO <= u.W && yBase - STEP >= rows.y0 && yBase + kAtTileH * STEP < rows.y1;
    if( a.uniformPitch )
    {
        if( interior ) atrousStageImpl< STEP, true, true >( a, u, rows, xBase, yBase, sA, sB, sC, sD );
        else atrousStageImpl< STEP, true, false >( a, u, rows, xBase, yBase, sA, sB, sC, sD );
    }
    else atrousStageImpl< STEP, false, false >( a, u, rows, xBase, yBase, sA, sB, sC, sD );
#endif
}


// the pixel state with the tap of the retired kernel (four fp32 planes)
struct AtrousPixFp32 : AtrousPix
{
    // one tap: lk = log2 of the wavelet weight, or -inf outside the render area
    ASVGF_DEVICE void tap( const float4 qA, const float4 qB, const float4 qC, const float4 qD, const bool diag, const float lk )
    {
        const float TINY = __uint_as_float( 0x00800000u ); // 2^-126: lg2 = -126, so P * L stays finite when P = 0
        const float nN    = fmaxf( TINY, fmaf( n.z, qA.z, fmaf( n.y, qA.y, n.x * qA.x ) ) );
        const float L     = flg2( nN );
        const float wZ    = fabsf( depth - qA.w ) * ( diag ? invK2 : invK1 );
        const float eN    = fmaf( 128.0f, L, lk );
        const float eBase = fmaf( -wZ, wZ, eN );
        const float eDiff = fmaf( fabsf( l - qC.w ), -kLum, eBase );
        const float wBase = fex2( eBase );
        const float wDiff = fex2( eDiff );
        float       rfac; // max(0, 1 - 10|dr|) * wRough, both factors in [0,1]
        asm( "fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"( rfac ) : "f"( fabsf( roughByte - qD.w ) ), "f"( -wR10 ), "f"( wRough ) );
        const float wSpec = rfac * fex2( fmaf( P, L, eBase ) );
        diff.x = fmaf( qB.x, wDiff, diff.x ); diff.y = fmaf( qB.y, wDiff, diff.y ); diff.z = fmaf( qB.z, wDiff, diff.z );
        spec.x = fmaf( qC.x, wSpec, spec.x ); spec.y = fmaf( qC.y, wSpec, spec.y ); spec.z = fmaf( qC.z, wSpec, spec.z );
        indir.x = fmaf( qD.x, wBase, indir.x ); indir.y = fmaf( qD.y, wBase, indir.y ); indir.z = fmaf( qD.z, wBase, indir.z );
        wSum += wDiff; wSumSpec += wSpec; wSumIndir += wBase;
        // (var*w)*w.  A vanishing weight next to an infinite fp16 variance gives inf*0 = NaN here where the shader's exp() may
        // still hold a denormal (inf): non-finite either way, and GLSL leaves the underflow of exp() to the implementation.
        variance = fmaf( __fmul_rn( qB.w, wDiff ), wDiff, variance );
    }
};

// PAIR = false (default): a thread filters its two pixels one after the other (9 tile texels each): 64 registers, 4 CTAs per SM.
//               Bound by shared-memory bandwidth: 36 LDS.128 per pixel = 0.128 ms per iteration at 4K at 128 B/clk/SM, measured
//               0.175-0.198 (ncu: LDS pipe 65 %, short-scoreboard + MIO stalls 41 %, issue 55 %).
// PAIR = true:  a thread filters two vertically adjacent (dilated) pixels TOGETHER: their 3x3 neighbourhoods share two of
//               four rows, so 12 tile texels serve 16 tap evaluations -- 2/3 of the shared-memory traffic -- but the two pixel
//               states need 80 registers (3 CTAs per SM) and the kernel lands 2 % behind (ASVGF_ATROUS_PAIR=1).
#ifndef ASVGF_AT_PAIR_MIN_CTAS
    #define ASVGF_AT_PAIR_MIN_CTAS 3
#endif
// CENTRE = true (iterations >= 1 inside asvgfDenoise): the per-pixel weight constants come from the plane iteration 0 wrote.
template< int STEP, bool ITER0, bool LAST, bool PAIR, bool CENTRE >
__global__ void __launch_bounds__( kAtThreads, PAIR ? ASVGF_AT_PAIR_MIN_CTAS : ASVGF_AT_MIN_CTAS ) k_atrous_tiled( const __grid_constant__ AtrousArgs a, const __grid_constant__ Uniform u )
{
    pdlLaunchDependents();
    extern __shared__ __align__( 128 ) uint8_t smemRaw[];
    using G = AtrousGeom< STEP >;
    constexpr int RP = G::RP, NPX = G::NPX, HALO = G::HALO;
    constexpr int PIX = ( kAtTileW * kAtTileH ) / kAtThreads; // 2
    static_assert( PIX == 2, "two pixels per thread" );
    float4* sA = reinterpret_cast< float4* >( smemRaw ); // normal.xyz, depth
    float4* sB = sA + NPX;                               // diffuse rgb, variance
    float4* sC = sB + NPX;                               // specular rgb, luminance of the diffuse rgb
    float4* sD = sC + NPX;                               // indirect rgb, roughness byte (0..255 as float)

    // dilated rows: blockIdx.y enumerates (band, residue); row(j) = yBase + j*STEP, j = -1 .. kAtTileH
    const int band  = blockIdx.y / STEP, res = blockIdx.y - band * STEP;
    const int yBase = u.yBegin + band * ( kAtTileH * STEP ) + res;
    const int xBase = blockIdx.x * kAtTileW;
    const Rows rows = rowsOf( a.depth );

    // this thread's two pixels: column c, tile rows jOf(0), jOf(1).  PAIR: adjacent rows 2m, 2m+1; else rows w, w+4
    const int c  = threadIdx.x & ( kAtTileW - 1 ), m = threadIdx.x / kAtTileW; // m = 0..3
    auto      jOf = [ & ]( int k ) { return PAIR ? 2 * m + k : m + 4 * k; };
    const int px = xBase + c;

    // ---- tiles without a single valid surface (sky): every pixel passes its centre texel through (CmSVGFAtrous.comp:96-99;
    //      iteration 0 stores zeros, CmSVGFAtrous_Iter0.comp:208-216) and no neighbourhood is read: skip staging and taps ----
#ifndef ASVGF_AT_NO_SKY_TILES
    {
        bool anyValid = false;
#pragma unroll
        for( int k = 0; k < PIX; k++ )
        {
            const int  py = yBase + jOf( k ) * STEP;
            const bool ok = px < u.W && py < u.yEnd;
            const float dz = h2f( ldIf< uint16_t >( a.depth, px, py, ok, ( uint16_t )0x7c00 ) ); // the depth the guide plane carries
            anyValid = anyValid || ( ok && !( dz < 0.0f || dz > 10000.0f ) );
        }
        pdlWait(); // the depth image is the producer's (complete before the frame's first pass started); everything below reads this frame's passes
        if( !__syncthreads_or( anyValid ) )
        {
#pragma unroll
            for( int k = 0; k < PIX; k++ )
            {
                const int py = yBase + jOf( k ) * STEP;
                if( px >= u.W || py >= u.yEnd ) continue;
                float3 oDiff = make_float3( 0.0f, 0.0f, 0.0f ), oSpec = oDiff, oIndir = oDiff;
                float  oVar  = 0.0f;
                if( !ITER0 )
                {
                    const uint2  dd = __ldg( at< uint2 >( a.srcDiff, px, py ) );
                    const float2 rg = __half22float2( *reinterpret_cast< const __half2* >( &dd.x ) );
                    const float2 bv = __half22float2( *reinterpret_cast< const __half2* >( &dd.y ) );
                    oDiff = make_float3( rg.x, rg.y, bv.x ); oVar = bv.y;
                    oSpec  = decodeE5( __ldg( at< uint32_t >( a.srcSpec, px, py ) ) );
                    oIndir = decodeE5( __ldg( at< uint32_t >( a.srcIndir, px, py ) ) );
                }
                atrousStore< LAST >( a, u, px, py, oDiff, oVar, oSpec, oIndir );
            }
            return;
        }
    }
#else
    pdlWait();
#endif

    // ---- centre-only inputs of this thread's pixels: in flight while the tile is staged ----
    static_assert( !( ITER0 && CENTRE ), "iteration 0 writes the centre plane" );
    uint16_t cGrad[ PIX ], cSigma[ PIX ], cSpecLen[ PIX ];
    if constexpr( CENTRE )
    {
        // no registers held across the staging phase: the plane is prefetched now and loaded (from L1) when the pixel starts
#pragma unroll
        for( int k = 0; k < PIX; k++ )
        {
            const int py = yBase + jOf( k ) * STEP;
            if( px < u.W && py < u.yEnd ) asm volatile( "prefetch.global.L1 [%0];" ::"l"( at< float4 >( a.centre, px, py ) ) );
        }
    }
    else
    {
#pragma unroll
        for( int k = 0; k < PIX; k++ )
        {
            const int  py = yBase + jOf( k ) * STEP;
            const bool ok = px < u.W && py < u.yEnd;
            cGrad[ k ]    = ldIf< uint16_t >( a.depthGrad, px, py, ok, ( uint16_t )0 );
            cSpecLen[ k ] = ok ? __ldg( reinterpret_cast< const uint16_t* >( at< uint2 >( a.histLen, px, py ) ) + 2 ) : ( uint16_t )0; // .b of rgba16f
            cSigma[ k ]   = ITER0 ? ( uint16_t )0 : ldIf< uint16_t >( a.filteredVariance, px, py, ok, ( uint16_t )0 );
        }
    }

    if constexpr( LAST )
    {
        // the composition reads four more images at the centre pixel (CmSVGFAtrous.comp:280-301); no registers to spare for them
        // across the tap loop, so their lines are prefetched here and loaded (from L1) in the epilogue
#pragma unroll
        for( int k = 0; k < PIX; k++ )
        {
            const int py = yBase + jOf( k ) * STEP;
            if( px < u.W && py < u.yEnd )
            {
                const int sepx = u.W / 2, isOdd = px >= sepx ? 1 : 0, regX = ( px - isOdd * sepx ) * 2 + ( ( isOdd + py ) & 1 );
                asm volatile( "prefetch.global.L1 [%0];" ::"l"( at< uint32_t >( a.albedo, regX, py ) ) );
                asm volatile( "prefetch.global.L1 [%0];" ::"l"( at< uint2 >( a.throughput, px, py ) ) );
                asm volatile( "prefetch.global.L1 [%0];" ::"l"( at< uint16_t >( a.metRough, px, py ) ) );
                asm volatile( "prefetch.global.L1 [%0];" ::"l"( at< uint8_t >( a.isSky, px, py ) ) );
            }
        }
    }
    atrousStage< STEP >( a, u, rows, xBase, yBase, sA, sB, sC, sD );
    __syncthreads();

    const int   sep  = u.W / 2;
    const float NINF = __uint_as_float( 0xff800000u );
    // every tap of every pixel of this tile inside its render area (ShaderCommonGLSLFunc.h:345-366)?  Uniform per CTA.
    const bool tileInterior = ( xBase >= sep ? xBase - STEP >= sep : xBase + kAtTileW + STEP <= sep ) && xBase - STEP >= 0 && xBase + kAtTileW + STEP <= u.W &&
                              yBase - STEP >= 0 && yBase + kAtTileH * STEP < u.H;
    const int  ax0 = px >= sep ? sep : 0, ax1 = ax0 + sep;
    const bool okL = px - STEP >= ax0, okR = px + STEP < ax1;
    constexpr int dC = G::dCol;

    // Iter0.comp:158-187: 3x3 Gaussian {1/4,1/8,1/16} of the variance (products by powers of two are exact)
    auto sigmaOf = [ & ]( int ce, int k ) {
        if constexpr( CENTRE ) return 0.0f;
        if( !ITER0 ) return h2f( cSigma[ k ] );
        float r = 0.0f;
#pragma unroll
        for( int yy = -1; yy <= 1; yy++ )
#pragma unroll
            for( int xx = -1; xx <= 1; xx++ )
                r = fmaf( sB[ ce + yy * RP + xx * dC ].w, ( xx == 0 ? 0.5f : 0.25f ) * ( yy == 0 ? 0.5f : 0.25f ), r );
        return __fsqrt_rn( fmaxf( r, 0.0f ) );
    };
    // centre of pixel k at slot ce: derive (or fetch) the weight constants; iteration 0 publishes them and the filtered sigma
    auto begin = [ & ]( AtrousPix& s, int ce, int k, int py, bool inFrame ) {
        if constexpr( CENTRE )
        {
            const float4 cK = inFrame ? __ldg( at< float4 >( a.centre, px, py ) ) : make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
            s.initFromCentre( sA[ ce ], sB[ ce ], sC[ ce ], sD[ ce ], cK, a.iteration );
        }
        else
        {
            const float sigmaL = sigmaOf( ce, k );
            s.init( sA[ ce ], sB[ ce ], sC[ ce ], sD[ ce ], h2f( cGrad[ k ] ), sigmaL, h2f( cSpecLen[ k ] ), a.iteration );
            if( ITER0 && inFrame )
            {
                if( !s.invalid ) *atW< uint16_t >( a.filteredVariance, px, py ) = f2h( sigmaL );
                *atW< float4 >( a.centre, px, py ) = make_float4( s.kLum, s.invK1, s.invK2, s.P );
            }
        }
    };
    auto finish = [ & ]( AtrousPix& s, int py, int ce ) {
        if( s.invalid ) // :96-99 pass-through of the centre (iteration 0, Iter0.comp:208-216: zeros); in the pair form the sums may hold taps
        {
            const float4 cB = sB[ ce ], cC = sC[ ce ], cD = sD[ ce ];
            s.diff = make_float3( cB.x, cB.y, cB.z ); s.variance = cB.w; s.spec = make_float3( cC.x, cC.y, cC.z ); s.indir = make_float3( cD.x, cD.y, cD.z );
            if( ITER0 ) { s.variance = 0.0f; s.diff = s.spec = s.indir = make_float3( 0.0f, 0.0f, 0.0f ); }
        }
        else s.normalise();
        atrousStore< LAST >( a, u, px, py, s.diff, s.variance, s.spec, s.indir );
    };

    if constexpr( !PAIR )
    {
#pragma unroll
        for( int k = 0; k < PIX; k++ )
        {
            const int j = jOf( k ), py = yBase + j * STEP;
            if( px >= u.W || py >= u.yEnd ) continue;
            const int ce = G::slot( j + 1, c + HALO );
            AtrousPixFp32 s;
            begin( s, ce, k, py, true );
            if( !s.invalid )
            {
                // interior tiles (all but the frame / checkerboard-separator borders): every tap is inside, lk is an immediate
                auto taps = [ & ]( auto interior ) {
                    constexpr bool INTERIOR = decltype( interior )::value;
                    const bool okT = py - STEP >= 0, okB = py + STEP < u.H;
#pragma unroll
                    for( int yy = -1; yy <= 1; yy++ )
#pragma unroll
                        for( int xx = -1; xx <= 1; xx++ )
                        {
                            if( xx == 0 && yy == 0 ) continue;
                            const bool  diag = xx != 0 && yy != 0;
                            const int   qe   = ce + yy * RP + xx * dC;
                            const bool  in   = INTERIOR || ( ( xx < 0 ? okL : xx > 0 ? okR : true ) && ( yy < 0 ? okT : yy > 0 ? okB : true ) );
                            s.tap( sA[ qe ], sB[ qe ], sC[ qe ], sD[ qe ], diag, in ? ( diag ? -2.0f : -1.0f ) : NINF );
                        }
                };
                if( tileInterior ) taps( std::true_type{} );
                else taps( std::false_type{} );
            }
            finish( s, py, ce );
        }
    }
    else
    {
        const int j0 = jOf( 0 ), py0 = yBase + j0 * STEP, py1 = py0 + STEP;
        if( px >= u.W || py0 >= u.yEnd ) return;
        const bool have1 = py1 < u.yEnd;
        const int  ce0 = G::slot( j0 + 1, c + HALO ), ce1 = ce0 + RP;
        AtrousPixFp32 s0, s1;
        begin( s0, ce0, 0, py0, true );
        begin( s1, ce1, 1, py1, have1 );
        if( !( s0.invalid && s1.invalid ) )
        {
            // staged rows j0 .. j0+3 (= tile rows j0-1 .. j0+2): row t feeds pixel 0 as dy = t-1 and pixel 1 as dy = t-2
            const bool okT0 = py0 - STEP >= 0, okB0 = py0 + STEP < u.H, okT1 = py1 - STEP >= 0, okB1 = py1 + STEP < u.H;
#pragma unroll
            for( int t = 0; t < 4; t++ )
#pragma unroll
                for( int xx = -1; xx <= 1; xx++ )
                {
                    const int    qe = ce0 + ( t - 1 ) * RP + xx * dC;
                    const float4 qA = sA[ qe ], qB = sB[ qe ], qC = sC[ qe ], qD = sD[ qe ];
                    const bool   okX = xx < 0 ? okL : xx > 0 ? okR : true;
                    if( t <= 2 && !( t == 1 && xx == 0 ) ) // pixel 0: dy = t - 1
                    {
                        const int  yy = t - 1;
                        const bool diag = xx != 0 && yy != 0;
                        const bool in = tileInterior || ( okX && ( yy < 0 ? okT0 : yy > 0 ? okB0 : true ) );
                        s0.tap( qA, qB, qC, qD, diag, in ? ( diag ? -2.0f : -1.0f ) : NINF );
                    }
                    if( t >= 1 && !( t == 2 && xx == 0 ) ) // pixel 1: dy = t - 2
                    {
                        const int  yy = t - 2;
                        const bool diag = xx != 0 && yy != 0;
                        const bool in = tileInterior || ( okX && ( yy < 0 ? okT1 : yy > 0 ? okB1 : true ) );
                        s1.tap( qA, qB, qC, qD, diag, in ? ( diag ? -2.0f : -1.0f ) : NINF );
                    }
                }
        }
        finish( s0, py0, ce0 );
        if( have1 ) finish( s1, py1, ce1 );
    }
}


} // namespace asvgf
