// asvgf_fast.cuh -- production kernels of the gradient, temporal, anti-firefly and variance passes (sm_100a).
//
// Written for the B200 SM, where a one-thread-per-pixel translation of the shaders (experiments/asvgf_literal_kernels.cuh, not
// part of the library) is instruction-issue bound, not HBM bound (ncu, profiles/r01_v2_*: k_temporal 3040 instructions per pixel
// with DRAM at 12 %, k_antifirefly 1860 per pixel with DRAM at 3 %):
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
//   k_temporal_fast ............... CmSVGFTemporalAccumulation.comp:58-385   (asvgf_temporal.cuh)
//   k_antifirefly_fast ............ CmAntiFirefly.comp:47-159
//   k_variance_fast ............... CmSVGFEstimateVariance.comp:74-205
#pragma once

#include "asvgf_args.cuh"

namespace asvgf
{

// ------------------------------------------------------------------------------------------------------------
// fast scalar helpers (colour math only)
// ------------------------------------------------------------------------------------------------------------
#ifndef ASVGF_HOSTSIM
ASVGF_DEVICE float fex2( float x ) { float r; asm( "ex2.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ); return r; }
ASVGF_DEVICE float flg2( float x ) { float r; asm( "lg2.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ); return r; }
ASVGF_DEVICE float frcp( float x ) { float r; asm( "rcp.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ); return r; }
ASVGF_DEVICE float frsqrt( float x ) { float r; asm( "rsqrt.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ); return r; }
ASVGF_DEVICE float fsqrt( float x ) { float r; asm( "sqrt.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) ); return r; }
#else // tests/hostsim compiles this file for the host (test infrastructure, never the product): the MUFU approximations are libm calls there
ASVGF_DEVICE float fex2( float x ) { return exp2f( x ); }
ASVGF_DEVICE float flg2( float x ) { return log2f( x ); }
ASVGF_DEVICE float frcp( float x ) { return 1.0f / x; }
ASVGF_DEVICE float frsqrt( float x ) { return 1.0f / sqrtf( x ); }
ASVGF_DEVICE float fsqrt( float x ) { return sqrtf( x ); }
#endif
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

// E5B9G9R9 encoder (Utils.h:162-192) for components already clamped to [0, 65408], NaN-free.
//   shared exponent: floor(log2 max) + 1 + 15, one more when the 9-bit mantissa of max rounds up to 512, i.e. when the
//   top ten bits of max's fraction are all ones and the next is set: adding 0x4000 to the bit pattern carries into the exponent.
//   mantissas: x * 2^(24-e) rounded to nearest even = the low bits of x + 1.5 * 2^(e-1) (the sum's ulp is 2^(e-24)).
ASVGF_DEVICE uint32_t encodeE5Magic( float r, float g, float b )
{
    const float maxC = fmaxf( r, fmaxf( g, b ) );
    if( maxC == 0.0f ) return 0u;
    const uint32_t e     = max( ( __float_as_uint( maxC ) + 0x4000u ) >> 23, 111u ) - 111u; // max(-16, floor(log2)) + 16, bumped
    const uint32_t mbits = ( ( e + 126u ) << 23 ) | 0x400000u;                               // 1.5 * 2^(e-1)
    const float    m     = __uint_as_float( mbits );
    const uint32_t ri = __float_as_uint( __fadd_rn( r, m ) ) - mbits, gi = __float_as_uint( __fadd_rn( g, m ) ) - mbits, bi = __float_as_uint( __fadd_rn( b, m ) ) - mbits;
    return ( e << 27 ) | ( bi << 18 ) | ( gi << 9 ) | ri;
}
ASVGF_DEVICE uint32_t encodeE5MagicClamped( float3 c ) // NaN -> 0 like the oracle's clamp
{
    return encodeE5Magic( fminf( fmaxf( c.x, 0.0f ), 65408.0f ), fminf( fmaxf( c.y, 0.0f ), 65408.0f ), fminf( fmaxf( c.z, 0.0f ), 65408.0f ) );
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
// G: gradient samples.  Bit-exact: slot, positions and the unorm8 alphas.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__( kBlockW* kBlockH ) k_gradient_samples_fast( const __grid_constant__ GradientSamplesArgs a, const __grid_constant__ Uniform u )
{
    pdlLaunchDependents();
    const int gx = blockIdx.x * kBlockW + threadIdx.x;
    const int gy = a.gyBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( gx >= a.disPing.w || gy >= a.gyEnd ) return;
    constexpr int S = 3;
    if( gx * S >= u.W || gy * S >= u.H ) return;
    pdlWait();

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
        if( a.haloMiss && ( unsigned )ppx[ k ] < ( unsigned )u.W && ( unsigned )ppy[ k ] < ( unsigned )u.H && !full.in( ppx[ k ], ppy[ k ] ) ) atomicOr( a.haloMiss, 1u );
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

    // the three anti-lag alphas in the oracle's explicitly rounded operation order (RtGradients.rgen:139-177): the unorm8 codes
    // stored here feed (1 - alpha)^10 in the temporal pass, where one code of difference is a 4 % change of the history length --
    // so this pass is bit-exact, like the gradient a-trous passes behind it (1/9 of the pixels: the IEEE divisions cost nothing)
    float gD = antilagAlpha( subr( fwd.x, prevLum.x ), fmaxf( fwd.x, prevLum.x ) );
    float gI;
    {
        float       gradSample = fmaxf( 0.0f, subr( prevLum.x, fwd.x ) );
        const float brightness = luminance( fetchE5( a.indirAccumPrev, prevX, prevY ) );
        gradSample             = mulr( gradSample, subr( 1.0f, smoothstepr( 1.0f, 10.0f, brightness ) ) );
        gI                     = antilagAlpha( gradSample, fmaxf( fwd.x, prevLum.x ) );
    }
    float gS = antilagAlpha( subr( fwd.y, prevLum.y ), fmaxf( fwd.y, prevLum.y ) );
    gD = mulr( gD, u.gradientMult[ 0 ] ); gI = mulr( gI, u.gradientMult[ 1 ] ); gS = mulr( gS, u.gradientMult[ 2 ] );
    {
        const uint32_t h = fetchU32( a.disHistory, prevX / S, prevY / S ); // prevX,prevY >= 0 here
        gD = mixr( gD, unorm8ToFloat( h & 0xffu ), 0.5f );
        gI = mixr( gI, unorm8ToFloat( ( h >> 8 ) & 0xffu ), 0.5f );
        gS = mixr( gS, unorm8ToFloat( ( h >> 16 ) & 0xffu ), 0.5f );
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
    pdlLaunchDependents();
    const int gx = blockIdx.x * kBlockW + threadIdx.x;
    const int gy = a.gyBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( gx >= a.dst.w || gy >= a.gyEnd ) return;
    constexpr int S = 3;
    if( gx * S >= u.W || gy * S >= u.H ) return;
    const int sep = u.W / 2, areaH = u.H / S;
    const int ax0 = ( gx * S >= sep ? sep : 0 ) / S, ax1 = ( gx * S >= sep ? 2 * sep : sep ) / S;
    const int step = 1 << a.iteration;
    const int sy0 = max( a.src.row0, 0 ), sy1 = min( a.src.row0 + a.src.rows, a.src.h );
    pdlWait();

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
// T: temporal accumulation -- the kernel lives in asvgf_temporal.cuh; shared colour helper:
// ------------------------------------------------------------------------------------------------------------
ASVGF_DEVICE float specAccumSpeedFast( float minFrames, float Amax, float roughness, float powA, float parallax ) // :80-92, powA = pow(sat(1-NoV), 0.66)
{
    const float b                   = fmaf( roughness, roughness, 1.1f );
    const float parallaxSensitivity = ( b + powA ) * frcp( b - powA );
    const float powerScale          = fmaf( parallax, parallaxSensitivity, 1.0f );
    float       f                   = 1.0f - fex2( -200.0f * roughness * roughness );
    f *= fpow( saturatef( roughness ), 0.5f * powerScale );
    return clampf( 256.0f * f, minFrames, Amax );
}

// ------------------------------------------------------------------------------------------------------------
// F: anti-firefly.  One CTA = one 64x64 tile of ONE of the three E5 images (blockIdx.z).  Every E5B9G9R9 component
// is exactly representable in fp16, so the tile is decoded once into fp16 planes and the 5x5-minus-centre min/max
// runs as packed half2 min/max: separable (row windows of 5, and of 4 without the centre, for two adjacent output
// columns per lane from three 8-byte shared loads; then a sliding vertical combine in registers).
// ------------------------------------------------------------------------------------------------------------
#ifndef ASVGF_AF_MIN_CTAS
    #define ASVGF_AF_MIN_CTAS 4
#endif
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

__global__ void __launch_bounds__( kAfThreads, ASVGF_AF_MIN_CTAS ) k_antifirefly_fast( const __grid_constant__ AntiFireflyArgs a, const __grid_constant__ Uniform u )
{
    __shared__ __align__( 16 ) uint32_t sRG[ kAfT ][ kAfT ];
    __shared__ __align__( 16 ) uint16_t sB[ kAfT ][ kAfBPitch ];
    pdlLaunchDependents();
    const Img& src = a.src[ blockIdx.z ];
    const Img& dst = a.dst[ blockIdx.z ];
    const Rows rows = rowsOf( src );
    const int  bx = blockIdx.x * kAfTile - kAfR, by = u.yBegin + blockIdx.y * kAfTile - kAfR;
    pdlWait();

#ifdef ASVGF_AF_SCALAR_STAGING
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
#else
    // Staging by texel PAIRS (ncu of the one-texel form: 57 instructions per staged texel, more than half of the kernel -- a
    // division by the tile pitch, four bounds tests and a branch around every load).  The tile origin is even and W is even, so a
    // pair is inside or outside together and its address is 8-byte aligned: lane -> pair column, warp -> row, nine sweeps of
    // eight rows with all loads in flight, the column test hoisted, one row-pointer increment per sweep; the two pair columns
    // beyond the 32 lanes (tile columns 64..67) go to threads 0..135.  Out-of-image texels stage as zeros (the pinned fetch).
    {
        static_assert( kAfT == 68 && kAfThreads == 256 && ( kAfBPitch % 2 ) == 0, "staging geometry" );
        constexpr int SWEEPS = ( kAfT + 7 ) / 8; // 9
        const int  tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        const int  gxp = bx + 2 * tx;
        const bool okx = ( unsigned )gxp < ( unsigned )rows.w;
        const int  er = threadIdx.x >> 1, ec = 64 + 2 * ( threadIdx.x & 1 ); // the extra pair of threads 0..135
        const int  gxe = bx + ec;
        uint2      raw[ SWEEPS ], rawE = make_uint2( 0u, 0u );
#pragma unroll
        for( int k = 0; k < SWEEPS; k++ )
        {
            const int  r = 8 * k + ty, gy = by + r;
            const bool ok = okx && r < kAfT && gy >= rows.y0 && gy < rows.y1;
            raw[ k ] = make_uint2( 0u, 0u );
            if( ok ) raw[ k ] = __ldg( reinterpret_cast< const uint2* >( at< uint32_t >( src, gxp, gy ) ) );
        }
        {
            const int gy = by + er;
            if( er < kAfT && ( unsigned )gxe < ( unsigned )rows.w && gy >= rows.y0 && gy < rows.y1 ) rawE = __ldg( reinterpret_cast< const uint2* >( at< uint32_t >( src, gxe, gy ) ) );
        }
        auto put = [ & ]( int r, int c, const uint2 q ) {
            const float3  v0 = decodeE5( q.x ), v1 = decodeE5( q.y );
            const __half2 rg0 = __floats2half2_rn( v0.x, v0.y ), rg1 = __floats2half2_rn( v1.x, v1.y ), bb = __floats2half2_rn( v0.z, v1.z );
            *reinterpret_cast< uint2* >( &sRG[ r ][ c ] )   = make_uint2( *reinterpret_cast< const uint32_t* >( &rg0 ), *reinterpret_cast< const uint32_t* >( &rg1 ) );
            *reinterpret_cast< uint32_t* >( &sB[ r ][ c ] ) = *reinterpret_cast< const uint32_t* >( &bb );
        };
#pragma unroll
        for( int k = 0; k < SWEEPS; k++ )
        {
            const int r = 8 * k + ty;
            if( r < kAfT ) put( r, 2 * tx, raw[ k ] );
        }
        if( er < kAfT ) put( er, ec, rawE );
    }
#endif
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
                const uint2  o  = make_uint2( encodeE5Magic( f0.x, f0.y, fb.x ), encodeE5Magic( f1.x, f1.y, fb.y ) );
                *reinterpret_cast< uint2* >( atW< uint32_t >( dst, gx, gy ) ) = o;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// V: variance estimation
// ------------------------------------------------------------------------------------------------------------
// the spatial estimate of CmSVGFEstimateVariance.comp:163-205 for a pixel with too little history: 3x3 bilateral moments; also
// filters the colour
ASVGF_DEVICE float varianceSpatial( const VarianceArgs& a, const Rows& rows, int px, int py, float pixDepth, float historyLength, float2 temporalMoments, float3& color )
{
    float variance;
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
    return variance;
}

// One thread per pixel.  (Measured and dropped in round 2: four pixels per thread with 16-byte loads, 0.102 ms instead of 0.082 ms at
// 4K; queueing the short-history pixels for a second, compacted kernel, 0.096 ms.)
__global__ void __launch_bounds__( kBlockW* kBlockH ) k_variance_fast( const __grid_constant__ VarianceArgs a, const __grid_constant__ Uniform u )
{
    pdlLaunchDependents();
    const int px = blockIdx.x * kBlockW + threadIdx.x;
    const int py = u.yBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( px >= u.W || py >= u.yEnd ) return;
    const Rows rows = rowsOf( a.depth );
    pdlWait();

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
        if( historyLength > 4.0f ) // :153-161; the cancellation keeps the oracle's rounding
            variance = fmaxf( 0.0f, subr( temporalMoments.y, mulr( temporalMoments.x, temporalMoments.x ) ) ) * frcp( fmaxf( 1.0f, historyLength - 4.0f ) );
        else variance = varianceSpatial( a, rows, px, py, pixDepth, historyLength, temporalMoments, color );
    }
    const __half2 o0 = __floats2half2_rn( color.x, color.y ), o1 = __floats2half2_rn( color.z, variance );
    *atW< uint2 >( a.out, px, py ) = make_uint2( *reinterpret_cast< const uint32_t* >( &o0 ), *reinterpret_cast< const uint32_t* >( &o1 ) );
}

} // namespace asvgf
