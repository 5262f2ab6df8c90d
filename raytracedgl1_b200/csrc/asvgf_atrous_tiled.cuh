// asvgf_atrous_tiled.cuh -- the production a-trous kernel (all iterations), sm_100a.
//
// Same arithmetic as CmSVGFAtrous_Iter0.comp:158-335 / CmSVGFAtrous.comp:65-349 (see k_atrous in asvgf_kernels.cuh for
// the literal form), restructured for the B200 SM, where the literal form is instruction-issue bound (ncu: 2200
// instructions per pixel, DRAM at 6 % -- profiles/r01_v1_*):
//
//  * DILATED TILE.  For stride s a pixel only ever touches rows y, y+-s.  A CTA therefore owns 64 contiguous columns of
//    16 rows spaced s apart (rows y0+r, y0+r+s, ...; the s residues r go to s different CTAs).  Its halo is one dilated
//    row above/below and s columns left/right, so the read amplification is (18/16)*((64+2s)/64) = 1.16..1.41 for
//    every stride instead of the (1+2s/T)^2 of a square tile, and every global access is still a full-line row segment.
//  * DECODE ONCE.  Each tile pixel is unpacked once on the way into shared memory: normal through the exact sin/cos
//    table, E5B9G9R9 -> fp16 triples (every E5 value is exactly representable in fp16), luminance in the oracle's
//    operation order.  The 8 taps then cost shared-memory loads and converts only.
//  * WEIGHTS THROUGH ONE LOG.  w_n = (n.n')^128 and the specular lobe (n.n')^P share L = lg2(n.n'):
//      wBase = k * ex2(128 L - wz^2 log2e),  wDiff = k * ex2(128 L - (wz^2 + wl) log2e),  wSpec = wBase * r * ex2(P L)
//    i.e. 1 lg2 + 3 ex2 on the MUFU pipe instead of two powf and two expf (~150 FMA-pipe instructions) per tap.
//    Relative weight error <= ~3e-4 (|P|<=1152 times the 2^-22 absolute error of lg2.approx), inside the 1e-3 output
//    tolerance because the weights are normalised; integer/boolean outputs do not depend on this kernel.
#pragma once

#include "asvgf_kernels.cuh"

namespace asvgf
{

constexpr int kAtTileW   = 64;  // columns per CTA
constexpr int kAtTileH   = 16;  // dilated rows per CTA
constexpr int kAtThreads = 256; // 4 pixels per thread
constexpr int kAtBytesPerTilePixel = 16 + 4 + 8 + 8 + 8;

ASVGF_DEVICE float ex2Approx( float x ) { float r; asm( "ex2.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ); return r; }
// exp2 that keeps denormal results (MUFU.EX2 flushes them): a vanishing weight must stay non-zero exactly as long as
// the shader's exp() does, because an infinite fp16 variance times a zero weight is NaN but times a denormal weight is inf
ASVGF_DEVICE float ex2ApproxDenorm( float x )
{
    const bool  tiny = x < -126.0f;
    const float r    = ex2Approx( tiny ? x * 0.5f : x );
    return tiny ? r * r : r;
}
ASVGF_DEVICE float lg2Approx( float x ) { float r; asm( "lg2.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ); return r; }
ASVGF_DEVICE float rcpApprox( float x ) { float r; asm( "rcp.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ); return r; }

// E5B9G9R9 -> three exact fp16 values packed as (r | g<<16, b)
ASVGF_DEVICE uint2 e5ToHalf3( uint32_t p )
{
    const float3 c = decodeE5( p );
    const __half2 rg = __floats2half2_rn( c.x, c.y );
    const __half2 b0 = __floats2half2_rn( c.z, 0.0f );
    return make_uint2( *reinterpret_cast< const uint32_t* >( &rg ), *reinterpret_cast< const uint32_t* >( &b0 ) );
}
ASVGF_DEVICE float3 half3ToFloat( uint2 v )
{
    const float2 rg = __half22float2( *reinterpret_cast< const __half2* >( &v.x ) );
    const float  b  = __low2float( *reinterpret_cast< const __half2* >( &v.y ) );
    return make_float3( rg.x, rg.y, b );
}

template< bool ITER0, bool LAST >
__global__ void __launch_bounds__( kAtThreads ) k_atrous_tiled( const __grid_constant__ AtrousArgs a, const __grid_constant__ Uniform u )
{
    extern __shared__ __align__( 16 ) uint8_t smemRaw[];
    const int step = ITER0 ? 1 : ( 1 << a.iteration );
    const int tw   = kAtTileW + 2 * step; // tile width incl. halo columns
    constexpr int TROWS = kAtTileH + 2;
    const int npx  = tw * TROWS;
    float4*   sND    = reinterpret_cast< float4* >( smemRaw );               // normal.xyz, (depth half bits << 16 | roughness unorm8)
    uint2*    sDiff  = reinterpret_cast< uint2* >( sND + npx );              // diffuse rgb + variance, 4 x fp16 as stored
    uint2*    sSpec  = sDiff + npx;                                          // specular as 3 x fp16
    uint2*    sIndir = sSpec + npx;                                          // indirect as 3 x fp16
    float*    sLum   = reinterpret_cast< float* >( sIndir + npx );           // luminance of the diffuse rgb

    // dilated rows: blockIdx.y enumerates (band, residue); row(j) = yBase + j*step, j = -1 .. kAtTileH
    const int band  = blockIdx.y / step, res = blockIdx.y - band * step;
    const int yBase = u.yBegin + band * ( kAtTileH * step ) + res;
    const int xBase = blockIdx.x * kAtTileW;

    // ---- stage the tile: unpack every pixel once ----
    for( int e = threadIdx.x; e < npx; e += kAtThreads )
    {
        const int r = e / tw, c = e - r * tw;
        const int gx = xBase - step + c, gy = yBase + ( r - 1 ) * step;
        const uint32_t encN  = fetchU32( a.normal, gx, gy );
        const float3   n     = decodeNormal( a.lut, encN );
        uint32_t       dbits = 0, rbits = 0;
        uint2          dv    = make_uint2( 0u, 0u );
        if( inImage( a.depth, gx, gy ) )
        {
            dbits = __ldg( texel< uint16_t >( a.depth, gx, gy ) );
            rbits = __ldg( texel< uint16_t >( a.metRough, gx, gy ) ) >> 8;
            dv    = __ldg( texel< uint2 >( a.srcDiff, gx, gy ) );
        }
        sND[ e ]    = make_float4( n.x, n.y, n.z, __uint_as_float( ( dbits << 16 ) | rbits ) );
        sDiff[ e ]  = dv;
        sSpec[ e ]  = e5ToHalf3( fetchU32( a.srcSpec, gx, gy ) );
        sIndir[ e ] = e5ToHalf3( fetchU32( a.srcIndir, gx, gy ) );
        const float2 rg = __half22float2( *reinterpret_cast< const __half2* >( &dv.x ) );
        const float  b  = __low2float( *reinterpret_cast< const __half2* >( &dv.y ) );
        sLum[ e ] = luminance( make_float3( rg.x, rg.y, b ) );
    }
    __syncthreads();

    const int   sep     = u.W / 2;
    constexpr float LOG2E = 1.4426950408889634f;

    for( int k = 0; k < ( kAtTileW * kAtTileH ) / kAtThreads; k++ )
    {
        const int p  = threadIdx.x + k * kAtThreads;
        const int j  = p / kAtTileW, c = p - j * kAtTileW;
        const int px = xBase + c, py = yBase + j * step;
        if( px >= u.W || py >= u.yEnd ) continue;

        // checkerboarded render area of this pixel (ShaderCommonGLSLFunc.h:345-358): taps never cross x = W/2
        const int ax0 = px >= sep ? sep : 0, ax1 = ax0 + sep;

        const int    ce  = ( j + 1 ) * tw + ( c + step );
        const float4 cnd = sND[ ce ];
        const uint2  cdv = sDiff[ ce ];
        const float2 crg = __half22float2( *reinterpret_cast< const __half2* >( &cdv.x ) );
        const float2 cbv = __half22float2( *reinterpret_cast< const __half2* >( &cdv.y ) );
        float3 outDiff     = make_float3( crg.x, crg.y, cbv.x );
        float  outVariance = cbv.y;
        float3 outSpec     = half3ToFloat( sSpec[ ce ] );
        float3 outIndir    = half3ToFloat( sIndir[ ce ] );

        const uint32_t cw      = __float_as_uint( cnd.w );
        const float    depth   = h2f( ( uint16_t )( cw >> 16 ) );
        const bool     invalid = depth < 0.0f || depth > 10000.0f;

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
            const float3 normal    = make_float3( cnd.x, cnd.y, cnd.z );
            const float  gradDepth = fetchR16F( a.depthGrad, px, py );
            const float  l         = sLum[ ce ];

            float sigmaL;
            if( ITER0 ) // Iter0.comp:158-187, exact operation order
            {
                float r = 0.0f;
#pragma unroll
                for( int yy = -1; yy <= 1; yy++ )
#pragma unroll
                    for( int xx = -1; xx <= 1; xx++ )
                    {
                        const uint2 q = sDiff[ ce + yy * tw + xx ];
                        const float v = __high2float( *reinterpret_cast< const __half2* >( &q.y ) );
                        const float w = ( xx == 0 ? 0.5f : 0.25f ) * ( yy == 0 ? 0.5f : 0.25f );
                        r             = addr( r, mulr( v, w ) );
                    }
                r = __fsqrt_rn( fmaxf( r, 0.0f ) );
                storeR16F( a.filteredVariance, px, py, r );
                sigmaL = r;
            }
            else
            {
                sigmaL = fetchR16F( a.filteredVariance, px, py );
            }
            const float kLum = mulr( divr( 1.0f, addr( mulr( 4.0f, sigmaL ), 0.00001f ) ), LOG2E );

            const float roughness        = unorm8ToFloat( cw & 0xffu );
            const float wRoughMultiplier = clampf( subr( mulr( roughness, 30.0f ), ( float )a.iteration ), 0.0f, 1.0f );
            const float historyLengthSpec = fetchRGBA16F( a.histLen, px, py ).z;
            const float normalWeightScale = clampf( divr( historyLengthSpec, 8.0f ), 0.0f, 1.0f );
            const float alpha             = mulr( roughness, roughness );
            float       P                 = fmaxf( 0.01f, subr( divr( 2.0f, addr( mulr( alpha, alpha ), 1e-4f ) ), 2.0f ) );
            P                             = mulr( clampf( P, 8.0f, 1024.0f ), normalWeightScale );
            const bool  useP              = P > 0.0f;
            const float invK1 = rcpApprox( fmaxf( gradDepth, 0.01f ) ), invK2 = rcpApprox( fmaxf( mulr( gradDepth, 2.0f ), 0.01f ) );

            float weightSum = 1.0f, weightSumSpec = 1.0f, weightSumIndir = 1.0f;

#pragma unroll
            for( int yy = -1; yy <= 1; yy++ )
            {
                const int  qy   = py + yy * step;
                const bool yIn  = qy >= 0 && qy < u.H;
#pragma unroll
                for( int xx = -1; xx <= 1; xx++ )
                {
                    if( xx == 0 && yy == 0 ) continue;
                    const int   qx  = px + xx * step;
                    const float kf  = ( yIn && qx >= ax0 && qx < ax1 ) ? ( yy == 0 ? 1.0f : 0.5f ) * ( xx == 0 ? 1.0f : 0.5f ) : 0.0f;
                    const int   qe  = ce + yy * tw + xx * step;
                    const float4 qnd = sND[ qe ];
                    const uint2  qdv = sDiff[ qe ];
                    const float  lQ  = sLum[ qe ];
                    const float3 specQ  = half3ToFloat( sSpec[ qe ] );
                    const float3 indirQ = half3ToFloat( sIndir[ qe ] );

                    const uint32_t qw     = __float_as_uint( qnd.w );
                    const float    depthQ = h2f( ( uint16_t )( qw >> 16 ) );
                    const float    roughQ = ( float )( qw & 0xffu ) * ( 1.0f / 255.0f );
                    const float2   qrg    = __half22float2( *reinterpret_cast< const __half2* >( &qdv.x ) );
                    const float2   qbv    = __half22float2( *reinterpret_cast< const __half2* >( &qdv.y ) );

                    const float nN = fmaxf( 0.0f, normal.x * qnd.x + normal.y * qnd.y + normal.z * qnd.z );
                    const float L  = lg2Approx( nN );
                    const float wZ = fabsf( depth - depthQ ) * ( ( xx != 0 && yy != 0 ) ? invK2 : invK1 );
                    const float eBase = 128.0f * L - ( wZ * wZ ) * LOG2E;
                    const float eDiff = eBase - fabsf( l - lQ ) * kLum;
                    const float wBase = ex2Approx( eBase ) * kf;
                    const float wDiff = ex2ApproxDenorm( eDiff ) * kf;
                    const float rfac  = fmaxf( 0.0f, 1.0f - 10.0f * fabsf( roughness - roughQ ) ) * wRoughMultiplier;
                    const float wSpec = wBase * rfac * ex2Approx( useP ? P * L : 0.0f );

                    outDiff.x += qrg.x * wDiff; outDiff.y += qrg.y * wDiff; outDiff.z += qbv.x * wDiff;
                    outSpec.x += specQ.x * wSpec; outSpec.y += specQ.y * wSpec; outSpec.z += specQ.z * wSpec;
                    outIndir.x += indirQ.x * wBase; outIndir.y += indirQ.y * wBase; outIndir.z += indirQ.z * wBase;
                    outVariance = fmaf( __fmul_rn( qbv.y, wDiff ), wDiff, outVariance ); // (var*w)*w: inf*denormal stays inf, like the shader
                    weightSum += wDiff; weightSumSpec += wSpec; weightSumIndir += wBase;
                }
            }
            const float inv = divr( 1.0f, weightSum ), invSpec = divr( 1.0f, weightSumSpec ), invIndir = divr( 1.0f, weightSumIndir );
            outDiff     = outDiff * inv;
            outVariance = mulr( outVariance, mulr( inv, inv ) );
            outSpec     = outSpec * invSpec;
            outIndir    = outIndir * invIndir;
        }

        if( !LAST )
        {
            storeRGBA16F( a.dstDiff, px, py, make_float4( outDiff.x, outDiff.y, outDiff.z, outVariance ) );
            storeU32( a.dstSpec, px, py, encodeE5( outSpec ) );
            storeU32( a.dstIndir, px, py, encodeE5( outIndir ) );
            continue;
        }

        // composition, CmSVGFAtrous.comp:278-348 (exact operation order, as in k_atrous)
        const int2   reg        = regularPixFromCheckerboardPix( u, px, py );
        float3       albedo     = unpackUfloat111110( fetchU32( a.albedo, reg.x, reg.y ) );
        const float4 tp4        = fetchRGBA16F( a.throughput, px, py );
        const float3 throughput = make_float3( tp4.x, tp4.y, tp4.z );
        if( ( u.debugShowFlags & ( 1u << 9 ) ) != 0 ) albedo = make_float3( 1.0f, 1.0f, 1.0f );

        float3 illuminated;
        if( fetchU8( a.isSky, px, py ) == 0u )
        {
            const float  roughness = unorm8ToFloat( cw & 0xffu );
            const float  metallic  = fetchRG8( a.metRough, px, py ).x;
            const float3 specular  = mix3r( outSpec, outIndir, smoothstepr( 0.5f, 0.5f + 0.25f, roughness ) );
            const float3 roS       = mix3r( make_float3( 0.04f, 0.04f, 0.04f ), albedo, metallic );
            const float3 roD       = albedo * subr( 1.0f, metallic );
            illuminated            = ( outDiff + outIndir ) * roD + specular * roS;
            if( ( u.debugShowFlags & ( 1u << 5 ) ) != 0 ) illuminated = outDiff * albedo;
            else if( ( u.debugShowFlags & ( 1u << 6 ) ) != 0 ) illuminated = specular * roS;
            else if( ( u.debugShowFlags & ( 1u << 7 ) ) != 0 ) illuminated = outIndir;
            illuminated = illuminated * throughput;
            const float la = luminance( albedo );
            float       ambient = 1.0f;
            if( !( la > 0.02f ) ) { const float q = divr( subr( la, 0.02f ), 0.02f ); ambient = subr( 1.0f, mulr( q, q ) ); }
            illuminated = illuminated * ambient;
        }
        else
        {
            illuminated = albedo * throughput;
        }
        storeR16F( a.histogramInput, reg.x, reg.y, fmaxf( 0.0f, luminance( illuminated ) ) );
        if( ( float )reg.x < mulr( u.lightmapScreenCoverage, u.renderWidth ) ) illuminated = albedo * throughput;
        storeU32( a.preFinal, reg.x, reg.y, packUfloat111110( make_float3( fmaxf( 0.0f, illuminated.x ), fmaxf( 0.0f, illuminated.y ), fmaxf( 0.0f, illuminated.z ) ) ) );
    }
}

inline size_t atrousTiledSmemBytes( int step ) { return ( size_t )( kAtTileW + 2 * step ) * ( kAtTileH + 2 ) * kAtBytesPerTilePixel; }

} // namespace asvgf
