// asvgf_temporal.cuh -- T: temporal accumulation (CmSVGFTemporalAccumulation.comp:58-385), sm_100a.
//
// One thread per pixel.  ncu of the round-1 kernel: 1528 instructions per pixel at 74 % issue, DRAM 26 %, 64 registers with
// ~60 local-memory instructions per pixel -- instruction-bound, and a quarter of the instructions were addressing and control
// flow.  This version keeps every expression that decides an integer / boolean output (tap positions, bilinear weights, the
// depth / normal / render-area tests, weightSum > 1e-4) in the oracle's explicitly rounded operation order, and removes
// instructions around them:
//
//   * ONE TEXEL OFFSET.  When the 18 checkerboard-layout images share their pitch in texels (the library's own allocation
//     does: every row holds align256(W) texels), a pixel's offset y*P + x is computed once and every image is addressed as
//     base + offset * sizeof(texel); the four taps of a 2x2 gather are offset, +1, +P, +P+1.  (UNIFORM = false: caller-bound
//     images with other pitches take one multiply-add per image.)
//   * BRANCH-FREE TAPS.  A rejected tap is loaded as zeros under its predicate and enters the sums with weight 0 -- adding
//     +0 is exact, so weightSum keeps the oracle's bits -- instead of a branch per tap.
//   * SPARSE PROJECTION.  GLSL's proj * view * p is 2 x (16 mul + 12 add) per matrix pair in a fixed order.  For the usual
//     perspective matrix (Source/Matrix.cpp:380-395: only m0, m5, m8..m11, m14 non-zero) the products with the zero
//     coefficients are exact zeros whenever the view-space point is finite, and x + 0 = x exactly: clip.x, clip.y, clip.w
//     need 5 multiplies and 2 adds.  The host checks the coefficients, the kernel checks finiteness; anything else takes
//     the full product.
//   * E5B9G9R9 encode by one magic-number add per component instead of a scale multiply and a float->int conversion.
#pragma once

#include "asvgf_fast.cuh"

namespace asvgf
{

// 32 x 4 threads per CTA, eight CTAs per SM (64 registers).  Measured at 4K: 32x8 / 4 CTAs 0.385 ms, 32x4 / 8 CTAs 0.380 ms,
// 32x4 / 7 CTAs (72 registers, half the spills) 0.387 ms, 32x8 / 3 CTAs (80 registers, no spills) 0.423 ms.
#ifndef ASVGF_T_MIN_CTAS
    #define ASVGF_T_MIN_CTAS 8
#endif
#ifndef ASVGF_T_BLOCK_H
    #define ASVGF_T_BLOCK_H 4 // rows per CTA (32 x ASVGF_T_BLOCK_H threads)
#endif
constexpr int kTBlockH = ASVGF_T_BLOCK_H;

template< bool UNIFORM, typename T > ASVGF_DEVICE const T* atOff( const Img& im, uint32_t off, int x, int y )
{
    return UNIFORM ? reinterpret_cast< const T* >( im.p0 ) + off : at< T >( im, x, y );
}
template< bool UNIFORM, typename T > ASVGF_DEVICE T* atOffW( const Img& im, uint32_t off, int x, int y )
{
    return UNIFORM ? reinterpret_cast< T* >( im.p0 ) + off : atW< T >( im, x, y );
}
template< bool UNIFORM, typename T > ASVGF_DEVICE T ldOff( const Img& im, uint32_t off, int x, int y, bool ok, T zero )
{
    return ok ? __ldg( atOff< UNIFORM, T >( im, off, x, y ) ) : zero;
}

// the 2x2 consistency gather of CmSVGFTemporalAccumulation.comp:149-162 / :258-270 (exact).  Returns the four tap bits.
// encN = the centre pixel's packed normal: a tap with the same bits decodes to the same unit vector, whose dot product with
// itself is 1 up to rounding, i.e. > 0.95 -- when all four taps match (flat surfaces, most pixels) the sin/cos gathers are
// skipped; the result is identical by construction.
template< bool UNIFORM, bool BAND >
ASVGF_DEVICE uint32_t consistency2x2( const TemporalArgs& a, const Rows& rows, uint32_t pitchT, int bx, int by, int ax0, int ax1, int H, float depth, float motionZ,
                                      float3 normal, uint32_t encN, uint32_t& off00 )
{
    off00 = ( uint32_t )by * pitchT + ( uint32_t )bx; // (unsigned: a saturated tap coordinate wraps instead of overflowing; such a tap is never loaded)
    uint32_t dz[ 4 ], en[ 4 ];
    bool     in[ 4 ];
#pragma unroll
    for( int t = 0; t < 4; t++ )
    {
        const int x = ( int )( ( uint32_t )bx + ( t & 1 ) ), y = ( int )( ( uint32_t )by + ( t >> 1 ) ); // (INT_MAX + 1 wraps to INT_MIN: still outside)
        in[ t ] = y >= 0 && y < H && x >= ax0 && x < ax1 && y >= rows.y0 && y < rows.y1;
        if( BAND && !in[ t ] && y >= 0 && y < H && x >= ax0 && x < ax1 ) atomicOr( a.haloMiss, 1u ); // a tap of the frame this band does not hold
        const uint32_t o = off00 + ( t & 1 ) + ( t >> 1 ) * pitchT;
        dz[ t ] = ldOff< UNIFORM, uint16_t >( a.depthPrev, o, x, y, in[ t ], ( uint16_t )0 );
        en[ t ] = ldOff< UNIFORM, uint32_t >( a.normalPrev, o, x, y, in[ t ], encN ); // taps outside never count: let them match
    }
    uint32_t nOk = 15u;
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
    // a / b < 0.1f (ShaderCommonGLSLFunc.h:368-371) from an approximate quotient; only a quotient within 1e-5 of the threshold
    // (or a non-finite one) needs the IEEE division -- one branch for the four taps
    const float az = fabsf( depth ), raz = frcp( az );
    float       num[ 4 ], q[ 4 ];
    bool        near = false;
#pragma unroll
    for( int t = 0; t < 4; t++ )
    {
        num[ t ] = fabsf( addr( subr( depth, h2f( ( uint16_t )dz[ t ] ) ), motionZ ) );
        q[ t ]   = num[ t ] * raz;
        near     = near || !( fabsf( q[ t ] - 0.1f ) > 1.0e-5f ); // NaN counts as near
    }
    uint32_t zOk = 0u;
    if( !near )
    {
#pragma unroll
        for( int t = 0; t < 4; t++ ) zOk |= q[ t ] < 0.1f ? ( 1u << t ) : 0u;
    }
    else
    {
#pragma unroll
        for( int t = 0; t < 4; t++ ) zOk |= divr( num[ t ], az ) < 0.1f ? ( 1u << t ) : 0u;
    }
    uint32_t bits = zOk & nOk;
#pragma unroll
    for( int t = 0; t < 4; t++ ) bits &= in[ t ] ? ~0u : ~( 1u << t );
    return bits;
}

// float -> int of a tap coordinate.  The pin (oracle: saturate at +-1e9, NaN -> -1e9) only matters through "is the tap inside the
// image": the hardware conversion saturates huge values to INT_MIN / INT_MAX (outside, like +-1e9; x + 1 may wrap to INT_MIN:
// still outside), and NaN is sent outside explicitly (the hardware would give 0, a valid column).
ASVGF_DEVICE int f2iTap( float f ) { return f == f ? __float2int_rz( f ) : -1000000000; }

// exact reprojected tap position and bilinear weights (ShaderCommonGLSLFunc.h:271-277, CmSVGFTemporalAccumulation.comp:126-137)
ASVGF_DEVICE void reproject( const Uniform& u, int px, int py, float mx, float my, int& bx, int& by, float ( &bw )[ 4 ] )
{
    const float ppx = mulr( addr( mulr( addr( ( float )px, 0.5f ), u.invHalfW ), mx ), u.halfW );
    const float ppy = mulr( addr( mulr( addr( ( float )py, 0.5f ), u.invH ), my ), u.renderHeight );
    const float qx = subr( ppx, 0.5f ), qy = subr( ppy, 0.5f );
    const float fx = floorf( qx ), fy = floorf( qy );
    bx = f2iTap( fx ); by = f2iTap( fy );
    const float sx = subr( qx, fx ), sy = subr( qy, fy );
    const float ox = subr( 1.0f, sx ), oy = subr( 1.0f, sy );
    bw[ 0 ] = mulr( ox, oy ); bw[ 1 ] = mulr( sx, oy ); bw[ 2 ] = mulr( ox, sy ); bw[ 3 ] = mulr( sx, sy );
}

template< bool UNIFORM, bool BAND >
__global__ void __launch_bounds__( kBlockW* kTBlockH, ASVGF_T_MIN_CTAS ) k_temporal_fast( const __grid_constant__ TemporalArgs a, const __grid_constant__ Uniform u )
{
    pdlLaunchDependents();
    const int px = blockIdx.x * kBlockW + threadIdx.x;
    const int py = u.yBegin + blockIdx.y * kTBlockH + threadIdx.y;
    if( px >= u.W || py >= u.yEnd ) return;
    pdlWait(); // (the gradient image and, when the previous kernel of the stream is this handle's, the history it may still be writing)
    const Rows     rows   = rowsOf( a.depth );
    const uint32_t pitchT = a.depth.pitchT;
    const uint32_t o      = ( uint32_t )( py * ( int )pitchT + px );

    // ---- current pixel: all loads up front ----
    const uint32_t eD   = __ldg( atOff< UNIFORM, uint32_t >( a.unfDirect, o, px, py ) );
    const uint32_t eI   = __ldg( atOff< UNIFORM, uint32_t >( a.unfIndir, o, px, py ) );
    const uint32_t eS   = __ldg( atOff< UNIFORM, uint32_t >( a.unfSpec, o, px, py ) );
    const uint2    rv   = __ldg( atOff< UNIFORM, uint2 >( a.viewDir, o, px, py ) );
    const uint2    mv   = __ldg( atOff< UNIFORM, uint2 >( a.motion, o, px, py ) );
    const uint16_t dv   = __ldg( atOff< UNIFORM, uint16_t >( a.depth, o, px, py ) );
    const uint32_t nv32 = __ldg( atOff< UNIFORM, uint32_t >( a.normal, o, px, py ) );
    const uint16_t mr   = __ldg( atOff< UNIFORM, uint16_t >( a.metRough, o, px, py ) );
    const float4   sp4  = __ldg( atOff< UNIFORM, float4 >( a.surfPos, o, px, py ) );
    // the gradient image has (W+1)/3 x (H+1)/3 texels (Framebuffers.cpp:691-699): where W % 3 == 1 or H % 3 == 1 the last pixel column /
    // row maps to a stratum beyond it, and the fetch returns 0 like every other fetch outside an image (CmSVGFTemporalAccumulation.comp:185)
    const uint32_t gr   = ldIf< uint32_t >( a.disPing, px / 3, py / 3, px / 3 < a.disPing.w && py / 3 < a.disPing.h, 0u );

    const float2 mot01   = __half22float2( *reinterpret_cast< const __half2* >( &mv.x ) );
    const float  motionZ = __low2float( *reinterpret_cast< const __half2* >( &mv.y ) );
    const float  depth   = h2f( dv );
    const float3 normal  = decodeNormal( a.lut, nv32 );
    if( a.writeGuide ) // the decoded geometry the a-trous iterations stage instead of re-decoding it four times
        *atW< float4 >( a.guide, px, py ) = make_float4( normal.x, normal.y, normal.z, __uint_as_float( __float_as_uint( depth ) | ( ( uint32_t )mr >> 8 ) ) );

    const int sep = u.W / 2, ax0 = px >= sep ? sep : 0, ax1 = ax0 + sep;

    // ---- reprojection and the first 2x2 gather (exact) ----
    int   pixPrevX, pixPrevY;
    float bw[ 4 ];
    reproject( u, px, py, mot01.x, mot01.y, pixPrevX, pixPrevY, bw );
    uint32_t o00;
    uint32_t mask = consistency2x2< UNIFORM, BAND >( a, rows, pitchT, pixPrevX, pixPrevY, ax0, ax1, u.H, depth, motionZ, normal, nv32, o00 );

    // ---- history of the consistent taps: all in flight together; a rejected tap reads as zeros and weighs 0 ----
    uint2    hDiff[ 4 ], hLen[ 4 ];
    uint32_t hMom[ 4 ], hSpec[ 4 ], hInd[ 4 ];
#pragma unroll
    for( int t = 0; t < 4; t++ )
    {
        const int      x = ( int )( ( uint32_t )pixPrevX + ( t & 1 ) ), y = ( int )( ( uint32_t )pixPrevY + ( t >> 1 ) );
        const uint32_t ot = o00 + ( t & 1 ) + ( t >> 1 ) * pitchT;
        const bool     ok = ( mask >> t ) & 1u;
        hDiff[ t ] = ldOff< UNIFORM, uint2 >( a.diffHistory, ot, x, y, ok, make_uint2( 0u, 0u ) );
        hMom[ t ]  = ldOff< UNIFORM, uint32_t >( a.momentsPrev, ot, x, y, ok, 0u );
        hLen[ t ]  = ldOff< UNIFORM, uint2 >( a.histLenPrev, ot, x, y, ok, make_uint2( 0u, 0u ) );
        hSpec[ t ] = ldOff< UNIFORM, uint32_t >( a.specPrev, ot, x, y, ok, 0u );
        hInd[ t ]  = ldOff< UNIFORM, uint32_t >( a.indirPrev, ot, x, y, ok, 0u );
        bw[ t ]    = ok ? bw[ t ] : 0.0f;
    }
    const bool   sppOk = rows.in( pixPrevX, pixPrevY );
    if( BAND && !sppOk && ( unsigned )pixPrevX < ( unsigned )u.W && ( unsigned )pixPrevY < ( unsigned )u.H ) atomicOr( a.haloMiss, 1u );
    const float4 spp4  = ldOff< UNIFORM, float4 >( a.surfPosPrev, o00, pixPrevX, pixPrevY, sppOk, make_float4( 0.0f, 0.0f, 0.0f, 0.0f ) );

    float4 historyLengthDIS = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
    float3 diffPrev = make_float3( 0.0f, 0.0f, 0.0f ), specPrevOnSurf = diffPrev, indirPrev = diffPrev;
    float2 momentsPrev = make_float2( 0.0f, 0.0f );
    float  weightSum   = 0.0f;
#pragma unroll
    for( int t = 0; t < 4; t++ )
    {
        const float  w   = bw[ t ];
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
        weightSum = addr( weightSum, w ); // in tap order; + 0 for a rejected tap leaves the bits alone: decides weightSum > 1e-4
    }

    const float3 unfilteredDiff = decodeE5( eD ), unfilteredIndir = decodeE5( eI ), unfilteredSpec = decodeE5( eS );
    const float2 ray01 = __half22float2( *reinterpret_cast< const __half2* >( &rv.x ) );
    const float2 ray23 = __half22float2( *reinterpret_cast< const __half2* >( &rv.y ) );
    const float3 rayDir = make_float3( ray01.x, ray01.y, ray23.x );
    const float  rayHit = ray23.y;
    const float  roughness = unorm8ToFloatFast( ( uint32_t )mr >> 8 );
    const float  l  = luminance( unfilteredDiff );
    const float  l2 = mulr( l, l );

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
        float scx, scy, spx, spy;
        {
            // view-space x, y, z of both frames: (m0 x + m4 y) + (m8 z + m12 * 1); the product with w = 1 is exact
            auto viewRow = [ & ]( const float* m, int r ) {
                return addr( addr( mulr( m[ r ], virtualPos.x ), mulr( m[ 4 + r ], virtualPos.y ) ), addr( mulr( m[ 8 + r ], virtualPos.z ), m[ 12 + r ] ) );
            };
            const float cx = viewRow( u.view, 0 ), cy = viewRow( u.view, 1 ), cz = viewRow( u.view, 2 );
            const float qx = viewRow( u.viewPrev, 0 ), qy = viewRow( u.viewPrev, 1 ), qz = viewRow( u.viewPrev, 2 );
            const bool  finite = ( fabsf( cx ) + fabsf( cy ) + fabsf( cz ) + fabsf( qx ) + fabsf( qy ) + fabsf( qz ) ) < __uint_as_float( 0x7f800000u );
            float clipCx, clipCy, clipCw, clipPx, clipPy, clipPw;
            if( u.projSparse && finite )
            {
                // perspective matrices with zeros everywhere but m0, m5, m8..m11, m14: the dropped products are exact zeros
                const float* p = u.projection;
                const float* q = u.projectionPrev;
                clipCx = addr( mulr( p[ 0 ], cx ), mulr( p[ 8 ], cz ) ); clipCy = addr( mulr( p[ 5 ], cy ), mulr( p[ 9 ], cz ) ); clipCw = mulr( p[ 11 ], cz );
                clipPx = addr( mulr( q[ 0 ], qx ), mulr( q[ 8 ], qz ) ); clipPy = addr( mulr( q[ 5 ], qy ), mulr( q[ 9 ], qz ) ); clipPw = mulr( q[ 11 ], qz );
            }
            else
            {
                const float4 vp       = make_float4( virtualPos.x, virtualPos.y, virtualPos.z, 1.0f );
                const float4 clipCur  = mulMat( u.projection, mulMat( u.view, vp ) );
                const float4 clipPrev = mulMat( u.projectionPrev, mulMat( u.viewPrev, vp ) );
                clipCx = clipCur.x; clipCy = clipCur.y; clipCw = clipCur.w;
                clipPx = clipPrev.x; clipPy = clipPrev.y; clipPw = clipPrev.w;
            }
            scx = addr( mulr( divr( clipCx, clipCw ), 0.5f ), 0.5f ); scy = addr( mulr( divr( clipCy, clipCw ), 0.5f ), 0.5f );
            spx = addr( mulr( divr( clipPx, clipPw ), 0.5f ), 0.5f ); spy = addr( mulr( divr( clipPy, clipPw ), 0.5f ), 0.5f );
        }
        int   pixPrevSpecX, pixPrevSpecY;
        float bws[ 4 ];
        reproject( u, px, py, subr( spx, scx ), subr( spy, scy ), pixPrevSpecX, pixPrevSpecY, bws );
        // the same four texels as the first gather give the same test results (same centre depth / normal / motion)
        uint32_t oSpec;
        const uint32_t vbits = ( pixPrevSpecX == pixPrevX && pixPrevSpecY == pixPrevY )
                                   ? ( mask & 15u )
                                   : consistency2x2< UNIFORM, BAND >( a, rows, pitchT, pixPrevSpecX, pixPrevSpecY, ax0, ax1, u.H, depth, motionZ, normal, nv32, oSpec );
        float weightSumVirtual = 0.0f;
#pragma unroll
        for( int t = 0; t < 4; t++ ) weightSumVirtual = addr( weightSumVirtual, ( ( vbits >> t ) & 1u ) ? bws[ t ] : 0.0f );

        if( weightSumVirtual > 0.0001f ) // :282-307
        {
            mask |= 1u << 5;
            const float Avirt    = specAccumSpeedFast( minFrames, A, roughness, powA, 0.0f );
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

    *atOffW< UNIFORM, uint32_t >( a.outDiffTemporary, o, px, py ) = encodeE5MagicClamped( diffColorAccum );
    {
        const __half2 m = __floats2half2_rn( diffMomentsAccum.x, diffMomentsAccum.y );
        *atOffW< UNIFORM, uint32_t >( a.outMoments, o, px, py ) = *reinterpret_cast< const uint32_t* >( &m );
    }
    {
        const __half2 h0 = __floats2half2_rn( diffHistoryLengthAccum, indirHistoryLengthAccum ), h1 = __floats2half2_rn( specHistoryLengthAccum, specHitDistanceAccum );
        *atOffW< UNIFORM, uint2 >( a.outHistLen, o, px, py ) = make_uint2( *reinterpret_cast< const uint32_t* >( &h0 ), *reinterpret_cast< const uint32_t* >( &h1 ) );
    }
    *atOffW< UNIFORM, uint32_t >( a.outSpecAccum, o, px, py ) = encodeE5MagicClamped( specColorAccum );
    *atOffW< UNIFORM, uint32_t >( a.outIndirAccum, o, px, py ) = encodeE5MagicClamped( indirAccum );
    if( a.writeMask ) *atW< uint8_t >( a.outMask, px, py ) = ( uint8_t )mask;
}

} // namespace asvgf
