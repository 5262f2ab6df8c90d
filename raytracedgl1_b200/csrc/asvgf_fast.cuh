// asvgf_fast.cuh -- production kernels of the gradient, temporal, anti-firefly and variance passes (sm_100a).
//
// Same results as the literal kernels of asvgf_kernels.cuh (which restate the shaders expression by expression), but
// written for the B200 SM, where the literal forms are instruction-issue bound, not HBM bound (ncu, profiles/r01_v2_*:
// k_temporal 3040 instructions per pixel with DRAM at 12 %, k_antifirefly 1860 per pixel with DRAM at 3 %):
//
//   * everything that decides an INTEGER/BOOLEAN output (reprojected pixel positions, bilinear tap weights, the depth /
//     normal / render-area disocclusion tests, the chosen stratum slot) keeps the explicitly rounded operation order
//     of the oracle, so GradientPrevPix and the debug mask stay bit-exact;
//   * colour arithmetic uses fused multiply-adds, MUFU reciprocals and ex2/lg2 instead of IEEE divisions and
//     powf/expf/logf: relative error ~1e-6, far inside the 1e-3 output tolerance;
//   * image addressing is hoisted (one row pointer per image instead of a bounds-checked 64-bit address per fetch).
//
//   k_gradient_samples_fast ....... RtGradients.rgen:44-181 (image-space part)
//   k_gradient_atrous_fast ........ CmASVGFGradientAtrous.comp:42-105
//   k_temporal_fast ............... CmSVGFTemporalAccumulation.comp:58-385
//   k_antifirefly_fast ............ CmAntiFirefly.comp:47-159
//   k_variance_fast ............... CmSVGFEstimateVariance.comp:74-205
#pragma once

#include "asvgf_args.cuh"

namespace asvgf
{

// ------------------------------------------------------------------------------------------------------------
// fast scalar helpers (colour math only)
// ------------------------------------------------------------------------------------------------------------
ASVGF_DEVICE float fex2( float x ) { float r; asm( "ex2.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ); return r; }
ASVGF_DEVICE float flg2( float x ) { float r; asm( "lg2.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ); return r; }
ASVGF_DEVICE float frcp( float x ) { float r; asm( "rcp.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ); return r; }
ASVGF_DEVICE float frsqrt( float x ) { float r; asm( "rsqrt.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ); return r; }
ASVGF_DEVICE float fsqrt( float x ) { float r; asm( "sqrt.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ); return r; }
ASVGF_DEVICE float fdiv( float a, float b ) { return a * frcp( b ); }
// GLSL pow(x,y) for x >= 0, y > 0: exp2(y*log2(x)); pow(0,y) = 0
ASVGF_DEVICE float fpow( float x, float y ) { return fex2( y * flg2( x ) ); }
ASVGF_DEVICE float fmix( float a, float b, float t ) { return fmaf( b, t, a * ( 1.0f - t ) ); }
ASVGF_DEVICE float3 fmix3( float3 a, float3 b, float t )
{
    const float s = 1.0f - t;
    return make_float3( fmaf( b.x, t, a.x * s ), fmaf( b.y, t, a.y * s ), fmaf( b.z, t, a.z * s ) );
}
ASVGF_DEVICE float flum( float3 c ) { return fmaf( 0.0721f, c.z, fmaf( 0.7154f, c.y, 0.2125f * c.x ) ); }

// c/255 correctly rounded without a division: one Newton step on c*(1/255) is exact for every c in 0..255
// (checked exhaustively on the host, tests/test_oracle_codecs.py)
ASVGF_DEVICE float unorm8ToFloatFast( uint32_t c )
{
    const float cf = ( float )c, k = 1.0f / 255.0f;
    const float q  = cf * k;
    return fmaf( fmaf( -255.0f, q, cf ), k, q );
}

// E5B9G9R9 encoder for components already inside [0, 65408] (Utils.h:162-192 without the clamps)
ASVGF_DEVICE uint32_t encodeE5Unclamped( float r, float g, float b )
{
    const float maxC = fmaxf( r, fmaxf( g, b ) );
    if( maxC == 0.0f ) return 0u;
    const int ef        = ( int )( ( __float_as_uint( maxC ) >> 23 ) & 0xffu ) - 127;
    int       expShared = max( -16, ef ) + 16;
    float     s         = __uint_as_float( ( uint32_t )( 151 - expShared ) << 23 );
    if( __float2int_rn( maxC * s ) == 512 )
    {
        expShared += 1;
        s = __uint_as_float( ( uint32_t )( 151 - expShared ) << 23 );
    }
    const uint32_t ri = ( uint32_t )__float2int_rn( r * s ), gi = ( uint32_t )__float2int_rn( g * s ), bi = ( uint32_t )__float2int_rn( b * s );
    return ( ( uint32_t )expShared << 27 ) | ( bi << 18 ) | ( gi << 9 ) | ri;
}
// general encoder with fminf/fmaxf clamps (NaN -> 0 like the oracle's clamp)
ASVGF_DEVICE uint32_t encodeE5Fast( float3 c )
{
    return encodeE5Unclamped( fminf( fmaxf( c.x, 0.0f ), 65408.0f ), fminf( fmaxf( c.y, 0.0f ), 65408.0f ), fminf( fmaxf( c.z, 0.0f ), 65408.0f ) );
}

// row addressing: all full-resolution images share (W, H) and the stored row range
struct Rows
{
    int w, y0, y1; // columns, stored rows [y0,y1) (clipped to the frame)
    ASVGF_DEVICE bool in( int x, int y ) const { return ( unsigned )x < ( unsigned )w && y >= y0 && y < y1; }
};
ASVGF_DEVICE Rows rowsOf( const Img& im )
{
    Rows r;
    r.w  = im.w;
    r.y0 = max( im.row0, 0 );
    r.y1 = min( im.row0 + im.rows, im.h );
    return r;
}
template< typename T > ASVGF_DEVICE T ldIf( const Img& im, int x, int y, bool ok, T zero )
{
    return ok ? __ldg( at< T >( im, x, y ) ) : zero;
}

// exact a/b < 0.1f (the reference's depth test, ShaderCommonGLSLFunc.h:368-371) with an approximate-quotient
// fast path; only quotients within 1e-5 of the threshold (or non-finite) take the IEEE division
ASVGF_DEVICE bool quotientBelowTenth( float a, float b )
{
    const float q = a * frcp( b );
    if( fabsf( q - 0.1f ) > 1.0e-5f ) return q < 0.1f; // NaN/Inf fall through (comparison false / decided below)
    return divr( a, b ) < 0.1f;
}

// ------------------------------------------------------------------------------------------------------------
// G: gradient samples.  Integer outputs (slot, positions) exact; alpha arithmetic fast.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__( kBlockW* kBlockH ) k_gradient_samples_fast( const __grid_constant__ GradientSamplesArgs a, const __grid_constant__ Uniform u )
{
    const int gx = blockIdx.x * kBlockW + threadIdx.x;
    const int gy = a.gyBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( gx >= a.disPing.w || gy >= a.gyEnd ) return;
    constexpr int S = 3;
    if( gx * S >= u.W || gy * S >= u.H ) return;

    uint32_t* outGrad = atW< uint32_t >( a.disPing, gx, gy );
    uint8_t*  outSlot = atW< uint8_t >( a.gradientPrevPix, gx, gy );
    if( u.gradientMult[ 0 ] < 0.001f && u.gradientMult[ 1 ] < 0.001f && u.gradientMult[ 2 ] < 0.001f )
    {
        *outGrad = 0u; *outSlot = 10;
        return;
    }
    const uint32_t prevStrataPos = *outSlot;
    const Rows     full          = rowsOf( a.motion );

    // the 9 candidates: motion fetches first (independent), then the dependent luminance gathers
    uint2 mv[ 9 ];
#pragma unroll
    for( int k = 0; k < 9; k++ )
    {
        const int cx = gx * S + k % 3, cy = gy * S + k / 3;
        mv[ k ] = ldIf< uint2 >( a.motion, cx, cy, full.in( cx, cy ), make_uint2( 0u, 0u ) );
    }
    // reprojected positions of all nine candidates, then all nine luminance gathers in flight together
    int      ppx[ 9 ], ppy[ 9 ];
    uint32_t lum[ 9 ];
#pragma unroll
    for( int k = 0; k < 9; k++ )
    {
        const int    cx = gx * S + k % 3, cy = gy * S + k / 3;
        const float2 m  = __half22float2( *reinterpret_cast< const __half2* >( &mv[ k ].x ) );
        const float  fx = mulr( addr( mulr( addr( ( float )cx, 0.5f ), u.invHalfW ), m.x ), u.halfW );
        const float  fy = mulr( addr( mulr( addr( ( float )cy, 0.5f ), u.invH ), m.y ), u.renderHeight );
        ppx[ k ] = f2i( floorf( fx ) ); ppy[ k ] = f2i( floorf( fy ) );
        lum[ k ] = ldIf< uint32_t >( a.gradientInputsPrev, ppx[ k ], ppy[ k ], ( uint32_t )k != prevStrataPos && full.in( ppx[ k ], ppy[ k ] ), 0u );
    }
    int      prevX = 0, prevY = 0, curX = 0, curY = 0;
    uint32_t curStrataPos = 0;
    float2   prevLum = make_float2( 0.0f, 0.0f );
    float    prevMax = 0.0f;
#pragma unroll
    for( int k = 0; k < 9; k++ ) // :75-94, candidates in slot order; last frame's slot is skipped
    {
        if( ( uint32_t )k == prevStrataPos ) continue;
        const float2 pl = __half22float2( *reinterpret_cast< const __half2* >( &lum[ k ] ) );
        const float  pm = fmaxf( pl.x, pl.y );
        if( pm > prevMax )
        {
            prevX = ppx[ k ]; prevY = ppy[ k ]; curX = gx * S + k % 3; curY = gy * S + k / 3;
            curStrataPos = ( uint32_t )k;
            prevLum = pl; prevMax = pm;
        }
    }

    const int sep = u.W / 2, ax0 = curX >= sep ? sep : 0;
    const bool inArea = prevY >= 0 && prevY < u.H && prevX >= ax0 && prevX < ax0 + sep;
    if( !inArea || fetchU8( a.isSky, curX, curY ) != 0u || ( prevLum.x <= 0.0f && prevLum.y <= 0.0f ) )
    {
        *outGrad = 0u; *outSlot = 10;
        return;
    }
    const float2 fwd = fetchRG16F( a.fwdLum, curX, curY );

    auto antilag = []( float gradSample, float normFactor ) { return normFactor > 0.0001f ? saturatef( fabsf( gradSample ) * frcp( normFactor ) ) : 0.0f; };
    float gD = antilag( fwd.x - prevLum.x, fmaxf( fwd.x, prevLum.x ) );
    float gI;
    {
        float       gradSample = fmaxf( 0.0f, prevLum.x - fwd.x );
        const float brightness = flum( fetchE5( a.indirAccumPrev, prevX, prevY ) );
        const float t          = saturatef( ( brightness - 1.0f ) * ( 1.0f / 9.0f ) );
        gradSample *= 1.0f - t * t * ( 3.0f - 2.0f * t );
        gI = antilag( gradSample, fmaxf( fwd.x, prevLum.x ) );
    }
    float gS = antilag( fwd.y - prevLum.y, fmaxf( fwd.y, prevLum.y ) );
    gD *= u.gradientMult[ 0 ]; gI *= u.gradientMult[ 1 ]; gS *= u.gradientMult[ 2 ];
    {
        const uint32_t h = fetchU32( a.disHistory, prevX / S, prevY / S ); // prevX,prevY >= 0 here
        gD = 0.5f * gD + 0.5f * unorm8ToFloatFast( h & 0xffu );
        gI = 0.5f * gI + 0.5f * unorm8ToFloatFast( ( h >> 8 ) & 0xffu );
        gS = 0.5f * gS + 0.5f * unorm8ToFloatFast( ( h >> 16 ) & 0xffu );
    }
    *outGrad = floatToUnorm8( gD ) | ( floatToUnorm8( gI ) << 8 ) | ( floatToUnorm8( gS ) << 16 );
    *outSlot = ( uint8_t )curStrataPos;
}

// ------------------------------------------------------------------------------------------------------------
// GA: gradient a-trous.  Bit-exact with the oracle: c/255 exact, tap products are exact (weights are powers of two),
// sums are single roundings in tap order, the final division is exact (a multiply when the weight sum is 4).
// (A variant that ran all four iterations of a tile in shared memory behind a 15-pixel halo was measured at 102 us
// against 4 x 19 us for these four launches at 4K -- the 2x redundant halo work costs more than three launches -- and
// was dropped.)
// ------------------------------------------------------------------------------------------------------------
ASVGF_DEVICE uint32_t gradientAtrousTexel( const uint32_t ( &tap )[ 9 ], const float ( &w )[ 9 ], float weightSum )
{
    float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f;
#pragma unroll
    for( int k = 0; k < 9; k++ )
    {
        o0 = fmaf( unorm8ToFloatFast( tap[ k ] & 0xffu ), w[ k ], o0 ); // product exact -> same as mul then add
        o1 = fmaf( unorm8ToFloatFast( ( tap[ k ] >> 8 ) & 0xffu ), w[ k ], o1 );
        o2 = fmaf( unorm8ToFloatFast( ( tap[ k ] >> 16 ) & 0xffu ), w[ k ], o2 );
    }
    if( weightSum == 4.0f ) { o0 *= 0.25f; o1 *= 0.25f; o2 *= 0.25f; }
    else { o0 = divr( o0, weightSum ); o1 = divr( o1, weightSum ); o2 = divr( o2, weightSum ); }
    return floatToUnorm8( o0 ) | ( floatToUnorm8( o1 ) << 8 ) | ( floatToUnorm8( o2 ) << 16 );
}

__global__ void __launch_bounds__( kBlockW* kBlockH ) k_gradient_atrous_fast( const __grid_constant__ GradientAtrousArgs a, const __grid_constant__ Uniform u )
{
    const int gx = blockIdx.x * kBlockW + threadIdx.x;
    const int gy = a.gyBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( gx >= a.dst.w || gy >= a.gyEnd ) return;
    constexpr int S = 3;
    if( gx * S >= u.W || gy * S >= u.H ) return;
    const int sep = u.W / 2, areaH = u.H / S;
    const int ax0 = ( gx * S >= sep ? sep : 0 ) / S, ax1 = ( gx * S >= sep ? 2 * sep : sep ) / S;
    const int step = 1 << a.iteration;
    const int sy0 = max( a.src.row0, 0 ), sy1 = min( a.src.row0 + a.src.rows, a.src.h );

    uint32_t tap[ 9 ];
    float    w[ 9 ];
    float    weightSum = 0.0f;
#pragma unroll
    for( int k = 0; k < 9; k++ )
    {
        const int xx = k % 3 - 1, yy = k / 3 - 1;
        const int qx = gx + xx * step, qy = gy + yy * step;
        const bool inside = qy >= 0 && qy < areaH && qx >= ax0 && qx < ax1;
        tap[ k ] = ldIf< uint32_t >( a.src, qx, qy, inside && qy >= sy0 && qy < sy1 && qx < a.src.w, 0u );
        w[ k ]   = inside ? ( yy == 0 ? 1.0f : 0.5f ) * ( xx == 0 ? 1.0f : 0.5f ) : 0.0f;
        weightSum = addr( weightSum, w[ k ] );
    }
    // A-SVGF gradients are zero wherever the lighting did not change, so most 3x3 neighbourhoods hold nine equal texels.
    // Then the weighted mean is that texel: sum(v w)/sum(w) differs from v = c/255 by < 1e-6 relative (at most ten
    // roundings), i.e. by < 2e-4 of a code after the * 255, and rounds back to c.  Exact, and skips the 27 conversions.
    const uint32_t c = tap[ 4 ] & 0xffffffu; // the centre tap is always inside
    bool           same = true;
#pragma unroll
    for( int k = 0; k < 9; k++ ) same = same && ( w[ k ] == 0.0f || ( tap[ k ] & 0xffffffu ) == c );
    uint32_t o = c;
    if( !same ) o = gradientAtrousTexel( tap, w, weightSum );
    *atW< uint32_t >( a.dst, gx, gy ) = o;
    if( a.iteration == 0 ) *atW< uint32_t >( a.history, gx, gy ) = o;
}

// ------------------------------------------------------------------------------------------------------------
// T: temporal accumulation
// ------------------------------------------------------------------------------------------------------------
ASVGF_DEVICE float3 decodeNormalExact( const NormalLut& lut, uint32_t p ) { return decodeNormal( lut, p ); }

struct TapTest
{
    bool  ok[ 4 ];
};

// the 2x2 consistency gather of CmSVGFTemporalAccumulation.comp:149-162 / :258-270 (exact)
// encN = the centre pixel's packed normal: a tap whose packed normal has the same bits decodes to the same vector, and
// dot(n, n) of a decoded (unit) normal is 1 up to rounding, i.e. > 0.95 -- so when all four taps match (flat surfaces,
// most pixels) the eight sin/cos table gathers and the dot products are skipped; the result is identical by construction.
ASVGF_DEVICE uint32_t consistencyBits( const TemporalArgs& a, const Rows& rows, int bx, int by, int ax0, int ax1, int H, float depth, float motionZ, float3 normal,
                                       uint32_t encN )
{
    uint32_t dz[ 4 ], en[ 4 ];
    bool     in[ 4 ];
#pragma unroll
    for( int t = 0; t < 4; t++ )
    {
        const int x = bx + ( t & 1 ), y = by + ( t >> 1 );
        in[ t ] = y >= 0 && y < H && x >= ax0 && x < ax1 && y >= rows.y0 && y < rows.y1;
        dz[ t ] = ldIf< uint16_t >( a.depthPrev, x, y, in[ t ], ( uint16_t )0 );
        en[ t ] = ldIf< uint32_t >( a.normalPrev, x, y, in[ t ], encN ); // taps outside never count: let them match
    }
    uint32_t    nOk = 15u;
    if( ( en[ 0 ] != encN ) | ( en[ 1 ] != encN ) | ( en[ 2 ] != encN ) | ( en[ 3 ] != encN ) )
    {
        float2 sp[ 4 ], st[ 4 ];
#pragma unroll
        for( int t = 0; t < 4; t++ )
        {
            sp[ t ] = __ldg( a.lut.phi + ( en[ t ] >> 16 ) );
            st[ t ] = __ldg( a.lut.theta + ( en[ t ] & 0xffffu ) );
        }
        nOk = 0u;
#pragma unroll
        for( int t = 0; t < 4; t++ )
        {
            const float3 nPrev = make_float3( mulr( sp[ t ].x, st[ t ].y ), mulr( sp[ t ].x, st[ t ].x ), sp[ t ].y );
            nOk |= dot3r( normal, nPrev ) > 0.95f ? ( 1u << t ) : 0u;
        }
    }
    uint32_t bits = 0;
    const float az = fabsf( depth );
#pragma unroll
    for( int t = 0; t < 4; t++ )
    {
        const float num = fabsf( addr( subr( depth, h2f( ( uint16_t )dz[ t ] ) ), motionZ ) );
        const bool  ok  = in[ t ] && ( ( nOk >> t ) & 1u ) && quotientBelowTenth( num, az );
        bits |= ok ? ( 1u << t ) : 0u;
    }
    return bits;
}

ASVGF_DEVICE float specAccumSpeedFast( float minFrames, float Amax, float roughness, float powA, float parallax ) // :80-92, powA = pow(sat(1-NoV), 0.66)
{
    const float b                   = fmaf( roughness, roughness, 1.1f );
    const float parallaxSensitivity = ( b + powA ) * frcp( b - powA );
    const float powerScale          = fmaf( parallax, parallaxSensitivity, 1.0f );
    float       f                   = 1.0f - fex2( -200.0f * roughness * roughness );
    f *= fpow( saturatef( roughness ), 0.5f * powerScale );
    return clampf( 256.0f * f, minFrames, Amax );
}

#ifndef ASVGF_T_EAGER_HISTORY
    #define ASVGF_T_EAGER_HISTORY 0
#endif
#ifndef ASVGF_T_MIN_CTAS
    #define ASVGF_T_MIN_CTAS 4 // 64 registers (100 B of spills): measured 0.485 ms vs 0.523 (3 CTAs, 80 regs) vs 0.602 (2 CTAs)
#endif
__global__ void __launch_bounds__( kBlockW* kBlockH, ASVGF_T_MIN_CTAS ) k_temporal_fast( const __grid_constant__ TemporalArgs a, const __grid_constant__ Uniform u )
{
    const int px = blockIdx.x * kBlockW + threadIdx.x;
    const int py = u.yBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( px >= u.W || py >= u.yEnd ) return;
    const Rows rows = rowsOf( a.depth );

    // ---- current pixel: all loads up front ----
    const uint32_t eD = __ldg( at< uint32_t >( a.unfDirect, px, py ) );
    const uint32_t eI = __ldg( at< uint32_t >( a.unfIndir, px, py ) );
    const uint32_t eS = __ldg( at< uint32_t >( a.unfSpec, px, py ) );
    const uint2    rv = __ldg( at< uint2 >( a.viewDir, px, py ) );
    const uint2    mv = __ldg( at< uint2 >( a.motion, px, py ) );
    const uint16_t dv = __ldg( at< uint16_t >( a.depth, px, py ) );
    const uint32_t nv32 = __ldg( at< uint32_t >( a.normal, px, py ) );
    const uint16_t mr = __ldg( at< uint16_t >( a.metRough, px, py ) );
    const float4   sp4 = __ldg( at< float4 >( a.surfPos, px, py ) );
    const uint32_t gr = __ldg( at< uint32_t >( a.disPing, px / 3, py / 3 ) );

    const float3 unfilteredDiff = decodeE5( eD ), unfilteredIndir = decodeE5( eI ), unfilteredSpec = decodeE5( eS );
    const float2 ray01 = __half22float2( *reinterpret_cast< const __half2* >( &rv.x ) );
    const float2 ray23 = __half22float2( *reinterpret_cast< const __half2* >( &rv.y ) );
    const float3 rayDir = make_float3( ray01.x, ray01.y, ray23.x );
    const float  rayHit = ray23.y;
    const float2 mot01 = __half22float2( *reinterpret_cast< const __half2* >( &mv.x ) );
    const float  motionZ = __low2float( *reinterpret_cast< const __half2* >( &mv.y ) );
    const float  depth = h2f( dv );
    const float3 normal = decodeNormal( a.lut, nv32 );
    const float  roughness = unorm8ToFloatFast( ( uint32_t )mr >> 8 );

    const int sep = u.W / 2, ax0 = px >= sep ? sep : 0, ax1 = ax0 + sep;

    const float l  = luminance( unfilteredDiff );
    const float l2 = mulr( l, l );

    // ---- reprojection (exact) ----
    int   pixPrevX, pixPrevY;
    float subX, subY;
    {
        const float ppx = mulr( addr( mulr( addr( ( float )px, 0.5f ), u.invHalfW ), mot01.x ), u.halfW );
        const float ppy = mulr( addr( mulr( addr( ( float )py, 0.5f ), u.invH ), mot01.y ), u.renderHeight );
        const float qx = subr( ppx, 0.5f ), qy = subr( ppy, 0.5f );
        const float fx = floorf( qx ), fy = floorf( qy );
        pixPrevX = f2i( fx ); pixPrevY = f2i( fy );
        subX = subr( qx, fx ); subY = subr( qy, fy );
    }
    const float bw[ 4 ] = { mulr( subr( 1.0f, subX ), subr( 1.0f, subY ) ), mulr( subX, subr( 1.0f, subY ) ),
                            mulr( subr( 1.0f, subX ), subY ), mulr( subX, subY ) };

    // ---- history fetches ----
    uint2    hDiff[ 4 ], hLen[ 4 ];
    uint32_t hMom[ 4 ], hSpec[ 4 ], hInd[ 4 ];
#if ASVGF_T_EAGER_HISTORY
    // issued for every tap inside the render area BEFORE the consistency tests: one serial memory round trip less
    // (the tests need depth/normal of the same texels, then the sin/cos table); inconsistent taps are loaded in vain
#pragma unroll
    for( int t = 0; t < 4; t++ )
    {
        const int  x = pixPrevX + ( t & 1 ), y = pixPrevY + ( t >> 1 );
        const bool ok = y >= 0 && y < u.H && x >= ax0 && x < ax1 && y >= rows.y0 && y < rows.y1;
        hDiff[ t ] = ldIf< uint2 >( a.diffHistory, x, y, ok, make_uint2( 0u, 0u ) );
        hMom[ t ]  = ldIf< uint32_t >( a.momentsPrev, x, y, ok, 0u );
        hLen[ t ]  = ldIf< uint2 >( a.histLenPrev, x, y, ok, make_uint2( 0u, 0u ) );
        hSpec[ t ] = ldIf< uint32_t >( a.specPrev, x, y, ok, 0u );
        hInd[ t ]  = ldIf< uint32_t >( a.indirPrev, x, y, ok, 0u );
    }
    uint32_t mask = consistencyBits( a, rows, pixPrevX, pixPrevY, ax0, ax1, u.H, depth, motionZ, normal, nv32 );
#else
    uint32_t mask = consistencyBits( a, rows, pixPrevX, pixPrevY, ax0, ax1, u.H, depth, motionZ, normal, nv32 );
    // of the consistent taps only (all in flight together)
#pragma unroll
    for( int t = 0; t < 4; t++ )
    {
        const int  x = pixPrevX + ( t & 1 ), y = pixPrevY + ( t >> 1 );
        const bool ok = ( mask >> t ) & 1u;
        hDiff[ t ] = ldIf< uint2 >( a.diffHistory, x, y, ok, make_uint2( 0u, 0u ) );
        hMom[ t ]  = ldIf< uint32_t >( a.momentsPrev, x, y, ok, 0u );
        hLen[ t ]  = ldIf< uint2 >( a.histLenPrev, x, y, ok, make_uint2( 0u, 0u ) );
        hSpec[ t ] = ldIf< uint32_t >( a.specPrev, x, y, ok, 0u );
        hInd[ t ]  = ldIf< uint32_t >( a.indirPrev, x, y, ok, 0u );
    }
#endif
    const bool   sppOk = rows.in( pixPrevX, pixPrevY );
    const float4 spp4  = ldIf< float4 >( a.surfPosPrev, pixPrevX, pixPrevY, sppOk, make_float4( 0.0f, 0.0f, 0.0f, 0.0f ) );

    float4 historyLengthDIS = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
    float3 diffPrev = make_float3( 0.0f, 0.0f, 0.0f ), specPrevOnSurf = diffPrev, indirPrev = diffPrev;
    float2 momentsPrev = make_float2( 0.0f, 0.0f );
    float  weightSum   = 0.0f;
#pragma unroll
    for( int t = 0; t < 4; t++ )
    {
        if( !( ( mask >> t ) & 1u ) ) continue;
        const float  w  = bw[ t ];
        const float2 d01 = __half22float2( *reinterpret_cast< const __half2* >( &hDiff[ t ].x ) );
        const float  d2  = __low2float( *reinterpret_cast< const __half2* >( &hDiff[ t ].y ) );
        diffPrev.x = fmaf( d01.x, w, diffPrev.x ); diffPrev.y = fmaf( d01.y, w, diffPrev.y ); diffPrev.z = fmaf( d2, w, diffPrev.z );
        const float2 mp = __half22float2( *reinterpret_cast< const __half2* >( &hMom[ t ] ) );
        momentsPrev.x = fmaf( mp.x, w, momentsPrev.x ); momentsPrev.y = fmaf( mp.y, w, momentsPrev.y );
        const float2 h01 = __half22float2( *reinterpret_cast< const __half2* >( &hLen[ t ].x ) );
        const float2 h23 = __half22float2( *reinterpret_cast< const __half2* >( &hLen[ t ].y ) );
        historyLengthDIS.x = fmaf( h01.x, w, historyLengthDIS.x ); historyLengthDIS.y = fmaf( h01.y, w, historyLengthDIS.y );
        historyLengthDIS.z = fmaf( h23.x, w, historyLengthDIS.z ); historyLengthDIS.w = fmaf( h23.y, w, historyLengthDIS.w );
        const float3 sc = decodeE5( hSpec[ t ] ), ic = decodeE5( hInd[ t ] );
        specPrevOnSurf.x = fmaf( sc.x, w, specPrevOnSurf.x ); specPrevOnSurf.y = fmaf( sc.y, w, specPrevOnSurf.y ); specPrevOnSurf.z = fmaf( sc.z, w, specPrevOnSurf.z );
        indirPrev.x = fmaf( ic.x, w, indirPrev.x ); indirPrev.y = fmaf( ic.y, w, indirPrev.y ); indirPrev.z = fmaf( ic.z, w, indirPrev.z );
        weightSum = addr( weightSum, w ); // exact order: decides the weightSum > 1e-4 bit
    }

    const float antilagAlphaDiff  = unorm8ToFloatFast( gr & 0xffu );
    const float antilagAlphaIndir = unorm8ToFloatFast( ( gr >> 8 ) & 0xffu );
    const float antilagAlphaSpec  = unorm8ToFloatFast( ( gr >> 16 ) & 0xffu );

    float3 specColorAccum, diffColorAccum, indirAccum;
    float2 diffMomentsAccum;
    float  specHistoryLengthAccum, specHitDistanceAccum, diffHistoryLengthAccum, indirHistoryLengthAccum;

    if( weightSum > 0.0001f )
    {
        mask |= 1u << 4;
        const float inv = frcp( weightSum );
        specPrevOnSurf = make_float3( specPrevOnSurf.x * inv, specPrevOnSurf.y * inv, specPrevOnSurf.z * inv );
        const float specHistoryLength   = historyLengthDIS.z * inv;
        const float specHitDistancePrev = historyLengthDIS.w * inv;

        const float3 surfPos = make_float3( sp4.x, sp4.y, sp4.z ), surfPosPrev = make_float3( spp4.x, spp4.y, spp4.z );
        const float  nv = fmaxf( 0.0f, -( normal.x * rayDir.x + normal.y * rayDir.y + normal.z * rayDir.z ) );

        const float minFrames = u.specMinFrames;
        const float oneMinusS = 1.0f - antilagAlphaSpec;
        const float om2 = oneMinusS * oneMinusS;
        const float A = clampf( fmaf( specHistoryLength, om2 * om2 * oneMinusS, 1.0f ), minFrames, 256.0f ); // pow(x,5)

        float parallax; // :58-69
        {
            const float3 q     = make_float3( surfPosPrev.x - u.cameraDelta[ 0 ], surfPosPrev.y - u.cameraDelta[ 1 ], surfPosPrev.z - u.cameraDelta[ 2 ] );
            const float  i1    = frsqrt( surfPos.x * surfPos.x + surfPos.y * surfPos.y + surfPos.z * surfPos.z );
            const float  i2    = frsqrt( q.x * q.x + q.y * q.y + q.z * q.z );
            const float  cosa  = saturatef( ( surfPos.x * q.x + surfPos.y * q.y + surfPos.z * q.z ) * ( i1 * i2 ) );
            parallax           = fsqrt( fmaxf( 0.0f, 1.0f - cosa * cosa ) ) * frcp( fmaxf( cosa, 1e-6f ) ) * 60.0f;
        }
        const float powA    = fpow( saturatef( 1.0f - nv ), 0.66f );
        const float AOnSurf = specAccumSpeedFast( minFrames, A, roughness, powA, parallax );

        const float  aOnSurf       = frcp( 1.0f + AOnSurf );
        const float3 specOnSurf    = fmix3( specPrevOnSurf, unfilteredSpec, aOnSurf );
        const float  hitDistOnSurf = fmix( specHitDistancePrev, rayHit, aOnSurf );

        specColorAccum         = specOnSurf;
        specHistoryLengthAccum = AOnSurf;
        specHitDistanceAccum   = hitDistOnSurf;

        float domFactor; // :93-99
        {
            const float da = 0.298475f * 0.69314718056f * flg2( 39.4115f - 39.0029f * roughness );
            domFactor      = saturatef( fmaf( fpow( saturatef( 1.0f - nv ), 10.8649f ), 1.0f - da, da ) );
        }
        const float  vs         = hitDistOnSurf * domFactor;
        const float3 virtualPos = make_float3( fmaf( rayDir.x, vs, surfPos.x ), fmaf( rayDir.y, vs, surfPos.y ), fmaf( rayDir.z, vs, surfPos.z ) );

        // virtual-point reprojection (exact operation order from here to the tap positions, :232-249)
        int   pixPrevSpecX, pixPrevSpecY;
        float subSX, subSY;
        {
            const float4 vp       = make_float4( virtualPos.x, virtualPos.y, virtualPos.z, 1.0f );
            const float4 viewCur  = mulMat( u.view, vp );
            const float4 viewPrev = mulMat( u.viewPrev, vp );
            const float4 clipCur  = mulMat( u.projection, viewCur );
            const float4 clipPrev = mulMat( u.projectionPrev, viewPrev );
            const float  scx = addr( mulr( divr( clipCur.x, clipCur.w ), 0.5f ), 0.5f ), scy = addr( mulr( divr( clipCur.y, clipCur.w ), 0.5f ), 0.5f );
            const float  spx = addr( mulr( divr( clipPrev.x, clipPrev.w ), 0.5f ), 0.5f ), spy = addr( mulr( divr( clipPrev.y, clipPrev.w ), 0.5f ), 0.5f );
            const float  ppx = mulr( addr( mulr( addr( ( float )px, 0.5f ), u.invHalfW ), subr( spx, scx ) ), u.halfW );
            const float  ppy = mulr( addr( mulr( addr( ( float )py, 0.5f ), u.invH ), subr( spy, scy ) ), u.renderHeight );
            const float  qx = subr( ppx, 0.5f ), qy = subr( ppy, 0.5f );
            const float  fx = floorf( qx ), fy = floorf( qy );
            pixPrevSpecX = f2i( fx ); pixPrevSpecY = f2i( fy );
            subSX = subr( qx, fx ); subSY = subr( qy, fy );
        }
        const float bws[ 4 ] = { mulr( subr( 1.0f, subSX ), subr( 1.0f, subSY ) ), mulr( subSX, subr( 1.0f, subSY ) ),
                                 mulr( subr( 1.0f, subSX ), subSY ), mulr( subSX, subSY ) };
        // the same four texels as the first gather give the same test results (same centre depth/normal/motion)
        const uint32_t vbits = ( pixPrevSpecX == pixPrevX && pixPrevSpecY == pixPrevY )
                                   ? ( mask & 15u )
                                   : consistencyBits( a, rows, pixPrevSpecX, pixPrevSpecY, ax0, ax1, u.H, depth, motionZ, normal, nv32 );
        float weightSumVirtual = 0.0f;
#pragma unroll
        for( int t = 0; t < 4; t++ )
            if( ( vbits >> t ) & 1u ) weightSumVirtual = addr( weightSumVirtual, bws[ t ] );

        if( weightSumVirtual > 0.0001f ) // :282-307
        {
            mask |= 1u << 5;
            float       Avirt    = specAccumSpeedFast( minFrames, A, roughness, powA, 0.0f );
            // mix(1/(1+Amin), 1/(1+Avirt), 1.0) = 0*(1/(1+Amin)) + 1/(1+Avirt): Avirt is unchanged unless 1/(1+Amin) is not finite
            const float Ahitdist = fminf( Avirt, 16.0f );
            const float aVirt    = frcp( 1.0f + Avirt );
            const float3 specVirt = fmix3( specPrevOnSurf, unfilteredSpec, aVirt );
            const float  amount   = domFactor;
            const float  al       = fmix( aOnSurf, aVirt, amount );
            specColorAccum         = fmix3( specOnSurf, specVirt, amount );
            specHistoryLengthAccum = frcp( al ) - 1.0f;
            specHitDistanceAccum   = fmix( specHitDistancePrev, rayHit, frcp( 1.0f + Ahitdist ) );
        }

        // direct diffuse and indirect (:326-367)
        diffPrev    = make_float3( diffPrev.x * inv, diffPrev.y * inv, diffPrev.z * inv );
        momentsPrev = make_float2( momentsPrev.x * inv, momentsPrev.y * inv );
        indirPrev   = make_float3( indirPrev.x * inv, indirPrev.y * inv, indirPrev.z * inv );
        {
            const float om = 1.0f - antilagAlphaDiff, o2 = om * om, o4 = o2 * o2;
            float       len = historyLengthDIS.x * inv * ( o4 * o4 * o2 ); // pow(x,10)
            len             = clampf( len + 1.0f, 1.0f, 256.0f );
            float alpha     = fmaxf( 0.01f, frcp( len ) );
            alpha           = fmix( alpha, 1.0f, antilagAlphaDiff );
            diffColorAccum   = fmix3( diffPrev, unfilteredDiff, alpha );
            diffMomentsAccum = make_float2( fmix( momentsPrev.x, l, alpha ), fmix( momentsPrev.y, l2, alpha ) );
            diffHistoryLengthAccum = len;
        }
        {
            const float om = 1.0f - antilagAlphaIndir, o2 = om * om, o4 = o2 * o2;
            float       len = historyLengthDIS.y * inv * ( o4 * o4 * o2 );
            len             = clampf( len + 1.0f, 1.0f, 256.0f );
            float alpha     = fmaxf( 0.01f, frcp( len ) );
            alpha           = fmix( alpha, 1.0f, antilagAlphaIndir );
            indirAccum      = fmix3( indirPrev, unfilteredIndir, alpha );
            indirHistoryLengthAccum = len;
        }
    }
    else // :310-315, :368-376
    {
        specColorAccum = unfilteredSpec; specHistoryLengthAccum = 1.0f; specHitDistanceAccum = rayHit;
        diffColorAccum = unfilteredDiff; diffMomentsAccum = make_float2( l, l2 ); diffHistoryLengthAccum = 1.0f;
        indirAccum = unfilteredIndir; indirHistoryLengthAccum = 1.0f;
    }

    *atW< uint32_t >( a.outDiffTemporary, px, py ) = encodeE5Fast( diffColorAccum );
    {
        const __half2 m = __floats2half2_rn( diffMomentsAccum.x, diffMomentsAccum.y );
        *atW< uint32_t >( a.outMoments, px, py ) = *reinterpret_cast< const uint32_t* >( &m );
    }
    {
        const __half2 h0 = __floats2half2_rn( diffHistoryLengthAccum, indirHistoryLengthAccum ), h1 = __floats2half2_rn( specHistoryLengthAccum, specHitDistanceAccum );
        *atW< uint2 >( a.outHistLen, px, py ) = make_uint2( *reinterpret_cast< const uint32_t* >( &h0 ), *reinterpret_cast< const uint32_t* >( &h1 ) );
    }
    *atW< uint32_t >( a.outSpecAccum, px, py ) = encodeE5Fast( specColorAccum );
    *atW< uint32_t >( a.outIndirAccum, px, py ) = encodeE5Fast( indirAccum );
    if( a.writeMask ) *atW< uint8_t >( a.outMask, px, py ) = ( uint8_t )mask;
}

// ------------------------------------------------------------------------------------------------------------
// F: anti-firefly.  One CTA = one 64x64 tile of ONE of the three E5 images (blockIdx.z).  Every E5B9G9R9 component
// is exactly representable in fp16, so the tile is decoded once into fp16 planes and the 5x5-minus-centre min/max
// runs as packed half2 min/max: separable (row windows of 5, and of 4 without the centre, for two adjacent output
// columns per lane from three 8-byte shared loads; then a sliding vertical combine in registers).
// ------------------------------------------------------------------------------------------------------------
constexpr int kAfTile = 64, kAfR = 2, kAfT = kAfTile + 2 * kAfR, kAfThreads = 256, kAfStrip = 8;
constexpr int kAfBPitch = kAfT + 4; // halves per row of the B plane (keeps rows 8-byte aligned)

ASVGF_DEVICE uint32_t hmin2u( uint32_t x, uint32_t y )
{
    const __half2 r = __hmin2( *reinterpret_cast< const __half2* >( &x ), *reinterpret_cast< const __half2* >( &y ) );
    return *reinterpret_cast< const uint32_t* >( &r );
}
ASVGF_DEVICE uint32_t hmax2u( uint32_t x, uint32_t y )
{
    const __half2 r = __hmax2( *reinterpret_cast< const __half2* >( &x ), *reinterpret_cast< const __half2* >( &y ) );
    return *reinterpret_cast< const uint32_t* >( &r );
}

struct AfRow // one tile row's contribution for a lane's two output columns: windows of RG (per column) and of B (two columns packed)
{
    uint32_t rg0, rg1, b;
};

__global__ void __launch_bounds__( kAfThreads ) k_antifirefly_fast( const __grid_constant__ AntiFireflyArgs a, const __grid_constant__ Uniform u )
{
    __shared__ __align__( 16 ) uint32_t sRG[ kAfT ][ kAfT ];
    __shared__ __align__( 16 ) uint16_t sB[ kAfT ][ kAfBPitch ];
    const Img& src = a.src[ blockIdx.z ];
    const Img& dst = a.dst[ blockIdx.z ];
    const Rows rows = rowsOf( src );
    const int  bx = blockIdx.x * kAfTile - kAfR, by = u.yBegin + blockIdx.y * kAfTile - kAfR;

    constexpr int AF_UNROLL = 5; // loads in flight per thread while staging
    for( int base = threadIdx.x; base < kAfT * kAfT; base += kAfThreads * AF_UNROLL )
    {
        uint32_t raw[ AF_UNROLL ];
#pragma unroll
        for( int k = 0; k < AF_UNROLL; k++ )
        {
            const int e = base + k * kAfThreads, r = e / kAfT, c = e - r * kAfT, gx = bx + c, gy = by + r;
            raw[ k ]    = ldIf< uint32_t >( src, gx, gy, e < kAfT * kAfT && rows.in( gx, gy ), 0u );
        }
#pragma unroll
        for( int k = 0; k < AF_UNROLL; k++ )
        {
            const int e = base + k * kAfThreads, r = e / kAfT, c = e - r * kAfT;
            if( e >= kAfT * kAfT ) break;
            const float3  v  = decodeE5( raw[ k ] );
            const __half2 rg = __floats2half2_rn( v.x, v.y );
            sRG[ r ][ c ] = *reinterpret_cast< const uint32_t* >( &rg );
            sB[ r ][ c ]  = __half_as_ushort( __float2half_rn( v.z ) );
        }
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = 2 * lane;                 // first tile column of the two 5-wide windows (output tile columns c0, c0+1)
    const int t0 = warp * kAfStrip;          // first tile row this warp needs
    const int gx = blockIdx.x * kAfTile + c0;
    const bool enabled = u.antiFireflyEnabled != 0;

    AfRow h5min[ 2 ], h5max[ 2 ];            // h5 of rows j-1, j (ring by parity)
    AfRow pmin[ 4 ], pmax[ 4 ];              // p[k] = min/max( h5[k], h5[k+1] ), ring of 4
    AfRow h4min[ 3 ], h4max[ 3 ], ctr[ 3 ];  // own-row windows without the centre, and the centre texels, ring of 3

#pragma unroll
    for( int j = 0; j < kAfStrip + 2 * kAfR; j++ )
    {
        const int t = t0 + j;
        const uint2 q0 = *reinterpret_cast< const uint2* >( &sRG[ t ][ c0 ] );
        const uint2 q1 = *reinterpret_cast< const uint2* >( &sRG[ t ][ c0 + 2 ] );
        const uint2 q2 = *reinterpret_cast< const uint2* >( &sRG[ t ][ c0 + 4 ] );
        const uint32_t v0 = q0.x, v1 = q0.y, v2 = q1.x, v3 = q1.y, v4 = q2.x, v5 = q2.y;
        const uint32_t X0 = *reinterpret_cast< const uint32_t* >( &sB[ t ][ c0 ] );
        const uint32_t X1 = *reinterpret_cast< const uint32_t* >( &sB[ t ][ c0 + 2 ] );
        const uint32_t X2 = *reinterpret_cast< const uint32_t* >( &sB[ t ][ c0 + 4 ] );
        const uint32_t Y0 = __byte_perm( X0, X1, 0x5432 ), Y1 = __byte_perm( X1, X2, 0x5432 ); // (b1,b2), (b3,b4)

        AfRow mn, mx, n4, x4;
        {
            const uint32_t m01 = hmin2u( v0, v1 ), m12 = hmin2u( v1, v2 ), m34 = hmin2u( v3, v4 ), m45 = hmin2u( v4, v5 );
            const uint32_t m = hmin2u( m12, m34 );
            mn.rg0 = hmin2u( v0, m ); mn.rg1 = hmin2u( m, v5 );
            n4.rg0 = hmin2u( m01, m34 ); n4.rg1 = hmin2u( m12, m45 );
            const uint32_t a0 = hmin2u( X0, Y0 ), a1 = hmin2u( X1, Y1 ), a2 = hmin2u( Y1, X2 );
            mn.b = hmin2u( hmin2u( a0, a1 ), X2 );
            n4.b = hmin2u( a0, a2 );
        }
        {
            const uint32_t m01 = hmax2u( v0, v1 ), m12 = hmax2u( v1, v2 ), m34 = hmax2u( v3, v4 ), m45 = hmax2u( v4, v5 );
            const uint32_t m = hmax2u( m12, m34 );
            mx.rg0 = hmax2u( v0, m ); mx.rg1 = hmax2u( m, v5 );
            x4.rg0 = hmax2u( m01, m34 ); x4.rg1 = hmax2u( m12, m45 );
            const uint32_t a0 = hmax2u( X0, Y0 ), a1 = hmax2u( X1, Y1 ), a2 = hmax2u( Y1, X2 );
            mx.b = hmax2u( hmax2u( a0, a1 ), X2 );
            x4.b = hmax2u( a0, a2 );
        }
        h5min[ j & 1 ] = mn; h5max[ j & 1 ] = mx;
        h4min[ j % 3 ] = n4; h4max[ j % 3 ] = x4;
        ctr[ j % 3 ]   = AfRow{ v2, v3, X1 };
        if( j >= 1 )
        {
            const AfRow& o = h5min[ ( j - 1 ) & 1 ];
            const AfRow& p = h5max[ ( j - 1 ) & 1 ];
            pmin[ ( j - 1 ) & 3 ] = AfRow{ hmin2u( o.rg0, mn.rg0 ), hmin2u( o.rg1, mn.rg1 ), hmin2u( o.b, mn.b ) };
            pmax[ ( j - 1 ) & 3 ] = AfRow{ hmax2u( p.rg0, mx.rg0 ), hmax2u( p.rg1, mx.rg1 ), hmax2u( p.b, mx.b ) };
        }
        if( j >= 4 )
        {
            // output row with centre at tile row i = j-2: window rows i-2..i+2 = p[i-2] (rows i-2,i-1), h4[i], p[i+1] (rows i+1,i+2)
            const int    i  = j - 2;
            const AfRow& pa = pmin[ ( i - 2 ) & 3 ];
            const AfRow& pb = pmin[ ( i + 1 ) & 3 ];
            const AfRow& qa = pmax[ ( i - 2 ) & 3 ];
            const AfRow& qb = pmax[ ( i + 1 ) & 3 ];
            const AfRow& n  = h4min[ i % 3 ];
            const AfRow& x  = h4max[ i % 3 ];
            AfRow        c  = ctr[ i % 3 ];
            if( enabled )
            {
                c.rg0 = hmin2u( hmax2u( c.rg0, hmin2u( hmin2u( pa.rg0, pb.rg0 ), n.rg0 ) ), hmax2u( hmax2u( qa.rg0, qb.rg0 ), x.rg0 ) );
                c.rg1 = hmin2u( hmax2u( c.rg1, hmin2u( hmin2u( pa.rg1, pb.rg1 ), n.rg1 ) ), hmax2u( hmax2u( qa.rg1, qb.rg1 ), x.rg1 ) );
                c.b   = hmin2u( hmax2u( c.b, hmin2u( hmin2u( pa.b, pb.b ), n.b ) ), hmax2u( hmax2u( qa.b, qb.b ), x.b ) );
            }
            const int gy = by + t0 + i; // tile row -> frame row
            if( gx < u.W && gy < u.yEnd )
            {
                const float2 f0 = __half22float2( *reinterpret_cast< const __half2* >( &c.rg0 ) );
                const float2 f1 = __half22float2( *reinterpret_cast< const __half2* >( &c.rg1 ) );
                const float2 fb = __half22float2( *reinterpret_cast< const __half2* >( &c.b ) );
                const uint2  o  = make_uint2( encodeE5Unclamped( f0.x, f0.y, fb.x ), encodeE5Unclamped( f1.x, f1.y, fb.y ) );
                *reinterpret_cast< uint2* >( atW< uint32_t >( dst, gx, gy ) ) = o;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// V: variance estimation
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__( kBlockW* kBlockH ) k_variance_fast( const __grid_constant__ VarianceArgs a, const __grid_constant__ Uniform u )
{
    const int px = blockIdx.x * kBlockW + threadIdx.x;
    const int py = u.yBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( px >= u.W || py >= u.yEnd ) return;
    const Rows rows = rowsOf( a.depth );

    const uint16_t dv = __ldg( at< uint16_t >( a.depth, px, py ) );
    const uint32_t ce = __ldg( at< uint32_t >( a.color, px, py ) );
    const uint2    hl = __ldg( at< uint2 >( a.histLen, px, py ) );
    const uint32_t mo = __ldg( at< uint32_t >( a.moments, px, py ) );

    const float pixDepth = h2f( dv );
    float3      color    = decodeE5( ce );
    float       variance;

    if( pixDepth < 0.0f || pixDepth > 10000.0f ) variance = 0.0f; // :143-147
    else
    {
        const float  historyLength   = __low2float( *reinterpret_cast< const __half2* >( &hl.x ) );
        const float2 temporalMoments = __half22float2( *reinterpret_cast< const __half2* >( &mo ) );
        if( historyLength > 4.0f ) // :153-161
        {
            variance = fmaxf( 0.0f, subr( temporalMoments.y, mulr( temporalMoments.x, temporalMoments.x ) ) ) * frcp( fmaxf( 1.0f, historyLength - 4.0f ) ); // the cancellation keeps the oracle's rounding
        }
        else
        {
            const float  l  = flum( color );
            float        m1 = l + temporalMoments.x, m2 = fmaf( l, l, temporalMoments.y );
            float        weightSum = 1.0f;
            const float3 n = decodeNormal( a.lut, __ldg( at< uint32_t >( a.normal, px, py ) ) );
            constexpr float LOG2E = 1.4426950408889634f;
#pragma unroll
            for( int k = 0; k < 9; k++ ) // :168-196; out-of-image and cross-separator neighbours participate as zeros
            {
                if( k == 4 ) continue;
                const int      qx = px + k % 3 - 1, qy = py + k / 3 - 1;
                const bool     ok = rows.in( qx, qy );
                const float3   colorQ = decodeE5( ldIf< uint32_t >( a.color, qx, qy, ok, 0u ) );
                const float    depthQ = h2f( ldIf< uint16_t >( a.depth, qx, qy, ok, ( uint16_t )0 ) );
                const float3   nQ     = decodeNormal( a.lut, ldIf< uint32_t >( a.normal, qx, qy, ok, 0u ) );
                const float    lQ     = flum( colorQ );
                const float    zm     = fmaxf( pixDepth, depthQ );
                const float    ez     = fabsf( pixDepth - depthQ ) * ( zm <= 0.0f ? 0.0f : frcp( zm ) );
                const float    nN     = fmaxf( 0.0f, n.x * nQ.x + n.y * nQ.y + n.z * nQ.z );
                const float    w      = fex2( fmaf( 128.0f, fmaxf( flg2( nN ), -1.0e4f ), -ez * LOG2E ) ); // exp(-ez) * nN^128
                if( !isnan( w ) )
                {
                    m1 = fmaf( lQ, w, m1 );
                    m2 = fmaf( lQ * lQ, w, m2 );
                    color.x = fmaf( colorQ.x, w, color.x ); color.y = fmaf( colorQ.y, w, color.y ); color.z = fmaf( colorQ.z, w, color.z );
                    weightSum += w;
                }
            }
            const float inv = weightSum <= 0.0f ? 0.0f : frcp( weightSum );
            m1 *= inv; m2 *= inv;
            color = make_float3( color.x * inv, color.y * inv, color.z * inv );
            variance = fmaxf( 0.0f, subr( m2, mulr( m1, m1 ) ) ) * fmaf( 2.0f, 1.0f - historyLength * 0.25f, 1.0f );
        }
    }
    const __half2 o0 = __floats2half2_rn( color.x, color.y ), o1 = __floats2half2_rn( color.z, variance );
    *atW< uint2 >( a.out, px, py ) = make_uint2( *reinterpret_cast< const uint32_t* >( &o0 ), *reinterpret_cast< const uint32_t* >( &o1 ) );
}

} // namespace asvgf
