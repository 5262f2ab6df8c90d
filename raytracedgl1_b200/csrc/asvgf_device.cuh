// asvgf_device.cuh -- device-side primitives of the A-SVGF path: image views, storage-format codecs and the
// GLSL helper functions the reference shaders share.
//
// Reference: Source/Shaders/Utils.h:54-72,124-140,162-207 (luminance, normal codec, E5B9G9R9),
//            Source/Shaders/ShaderCommonGLSLFunc.h:267-390 (checkerboard, reprojection, disocclusion tests).
//
// Numerical contract (shared with the oracle, see DESIGN.md "Pins"):
//   * every expression whose result decides an integer/boolean output (tap positions, disocclusion tests, masks)
//     is evaluated with explicitly rounded single operations (__fmul_rn/__fadd_rn never contract into FMA)
//   * decodeNormal takes sin/cos from a 65536-entry table of correctly rounded values built on the host
//   * stores convert f32 -> f16/unorm8/uf11/uf10 with round-to-nearest-even; out-of-image fetches return 0
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace asvgf
{

#define ASVGF_DEVICE_FWD __device__ __forceinline__

// Linear pitched image.  (w,h) is the full frame extent of the image; rows [row0,row0+rows) are stored
// (a horizontal band plus halo when the frame is split across GPUs, the whole image otherwise).
struct Img
{
    uint8_t* p;     // address of stored row `row0`
    uint32_t pitch; // bytes
    int32_t  w, h;
    int32_t  row0, rows;
    // derived by the host (setImgStorage): lean addressing for the production kernels
    uint8_t* p0;     // address frame row 0 would have (p - row0*pitch; only dereferenced at stored rows)
    uint32_t pitchT; // pitch in texels (pitch / bytes per texel)
};

// texel (x,y) of an image whose texel type is T: one 32-bit multiply-add and one widening add
template< typename T > ASVGF_DEVICE_FWD const T* at( const Img& im, int x, int y )
{
    return reinterpret_cast< const T* >( im.p0 ) + ( uint32_t )( y * ( int )im.pitchT + x );
}
template< typename T > ASVGF_DEVICE_FWD T* atW( const Img& im, int x, int y )
{
    return reinterpret_cast< T* >( im.p0 ) + ( uint32_t )( y * ( int )im.pitchT + x );
}

struct Uniform // the members of ShGlobalUniform the path reads (ShaderCommonC.h:209-304), passed by value
{
    float    view[ 16 ], viewPrev[ 16 ], projection[ 16 ], projectionPrev[ 16 ];
    float    cameraPosition[ 3 ], cameraPositionPrev[ 3 ];
    float    renderWidth, renderHeight, timeDelta, lightmapScreenCoverage;
    float    gradientMult[ 3 ];
    uint32_t debugShowFlags, antiFireflyEnabled;
    int32_t  W, H;         // int(renderWidth), int(renderHeight)
    int32_t  yBegin, yEnd; // rows this launch computes (the owned band)
    // derived on the host with the same IEEE single operations the shaders use (uniform across the launch)
    float    halfW, invHalfW, invH; // renderWidth/2, 1/(renderWidth/2), 1/renderHeight  (ShaderCommonGLSLFunc.h:271-277)
    float    specMinFrames;         // clamp(0.2/timeDelta, 0, 128)                       (CmSVGFTemporalAccumulation.comp:70-74)
    float    cameraDelta[ 3 ];      // cameraPositionPrev - cameraPosition                 (:60)
    int32_t  projSparse;            // projection and projectionPrev are zero outside m0, m5, m8..m11, m14 (Matrix.cpp:380-395)
};

struct NormalLut // (sin, cos) of phi_j = j*pi/65535 and theta_k = k*2*pi/65536, correctly rounded fp32
{
    const float2* phi;
    const float2* theta;
};

#define ASVGF_DEVICE __device__ __forceinline__

// ---- programmatic dependent launch (sm_90+): inside asvgfDenoise the passes of a frame are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the CTAs of pass k+1 are scheduled while the last wave of pass k drains
// (launch latency, argument fetch and index arithmetic overlap the tail) and block in pdlWait() until pass k has completed and
// flushed.  Nothing written by an earlier pass of the frame is touched before pdlWait(); without the launch attribute both
// instructions do nothing. ----
#ifndef ASVGF_HOSTSIM
ASVGF_DEVICE void pdlLaunchDependents() { asm volatile( "griddepcontrol.launch_dependents;" ); }
ASVGF_DEVICE void pdlWait() { asm volatile( "griddepcontrol.wait;" ::: "memory" ); }
#else // tests/hostsim (kernel source compiled for the host by the CPU test suite): grids run one after the other
ASVGF_DEVICE void pdlLaunchDependents() {}
ASVGF_DEVICE void pdlWait() {}
#endif

// ---- exactly rounded arithmetic (never fused) ----
ASVGF_DEVICE float mulr( float a, float b ) { return __fmul_rn( a, b ); }
ASVGF_DEVICE float addr( float a, float b ) { return __fadd_rn( a, b ); }
ASVGF_DEVICE float subr( float a, float b ) { return __fsub_rn( a, b ); }
ASVGF_DEVICE float divr( float a, float b ) { return __fdiv_rn( a, b ); }
ASVGF_DEVICE float dot3r( float3 a, float3 b ) { return addr( addr( mulr( a.x, b.x ), mulr( a.y, b.y ) ), mulr( a.z, b.z ) ); }
ASVGF_DEVICE float mixr( float a, float b, float t ) { return addr( mulr( a, subr( 1.0f, t ) ), mulr( b, t ) ); } // a*(1-t) + b*t
ASVGF_DEVICE float3 mix3r( float3 a, float3 b, float t ) { return make_float3( mixr( a.x, b.x, t ), mixr( a.y, b.y, t ), mixr( a.z, b.z, t ) ); }
ASVGF_DEVICE float clampf( float x, float lo, float hi ) { return fminf( fmaxf( x, lo ), hi ); }
ASVGF_DEVICE float saturatef( float x ) { return clampf( x, 0.0f, 1.0f ); }
ASVGF_DEVICE float safePositiveRcp( float f ) { return f <= 0.0f ? 0.0f : divr( 1.0f, f ); } // Utils.h:69-72
ASVGF_DEVICE float luminance( float3 c ) { return addr( addr( mulr( 0.2125f, c.x ), mulr( 0.7154f, c.y ) ), mulr( 0.0721f, c.z ) ); } // Utils.h:54-57
ASVGF_DEVICE float smoothstepr( float e0, float e1, float x )
{
    float t = saturatef( divr( subr( x, e0 ), subr( e1, e0 ) ) );
    return mulr( mulr( t, t ), subr( 3.0f, mulr( 2.0f, t ) ) );
}
ASVGF_DEVICE int f2i( float f ) // saturating float->int, same pin as the oracle
{
    if( !( f > -1.0e9f ) ) return -1000000000;
    if( f > 1.0e9f ) return 1000000000;
    return __float2int_rz( f );
}
ASVGF_DEVICE float3 operator+( float3 a, float3 b ) { return make_float3( addr( a.x, b.x ), addr( a.y, b.y ), addr( a.z, b.z ) ); }
ASVGF_DEVICE float3 operator-( float3 a, float3 b ) { return make_float3( subr( a.x, b.x ), subr( a.y, b.y ), subr( a.z, b.z ) ); }
ASVGF_DEVICE float3 operator*( float3 a, float s ) { return make_float3( mulr( a.x, s ), mulr( a.y, s ), mulr( a.z, s ) ); }
ASVGF_DEVICE float3 operator*( float3 a, float3 b ) { return make_float3( mulr( a.x, b.x ), mulr( a.y, b.y ), mulr( a.z, b.z ) ); }
ASVGF_DEVICE float3 normalize3r( float3 v ) { float inv = divr( 1.0f, __fsqrt_rn( dot3r( v, v ) ) ); return v * inv; }

// column-major mat4 * vec4 in the pinned order (GLM's, as the oracle): (c0*x + c1*y) + (c2*z + c3*w)
ASVGF_DEVICE float4 mulMat( const float* m, float4 v )
{
    float4 r;
    r.x = addr( addr( mulr( m[ 0 ], v.x ), mulr( m[ 4 ], v.y ) ), addr( mulr( m[ 8 ], v.z ), mulr( m[ 12 ], v.w ) ) );
    r.y = addr( addr( mulr( m[ 1 ], v.x ), mulr( m[ 5 ], v.y ) ), addr( mulr( m[ 9 ], v.z ), mulr( m[ 13 ], v.w ) ) );
    r.z = addr( addr( mulr( m[ 2 ], v.x ), mulr( m[ 6 ], v.y ) ), addr( mulr( m[ 10 ], v.z ), mulr( m[ 14 ], v.w ) ) );
    r.w = addr( addr( mulr( m[ 3 ], v.x ), mulr( m[ 7 ], v.y ) ), addr( mulr( m[ 11 ], v.z ), mulr( m[ 15 ], v.w ) ) );
    return r;
}

// ---- storage formats ----
ASVGF_DEVICE float3 decodeE5( uint32_t p ) // Utils.h:194-207; scale 2^(e-24) built from bits, exact
{
    // (mantissa | 2^23 bits, then one FMA that takes the 2^23 back out, was tried instead of the int->float conversions: the mask
    // and the OR need two LOP3 -- one immediate per instruction -- so the count does not drop: 33 I2FP become 33 LOP3 + 11 IADD3 in T)
    const float s = __uint_as_float( ( ( p >> 27 ) + 103u ) << 23 );
    return make_float3( mulr( s, ( float )( p & 0x1ffu ) ), mulr( s, ( float )( ( p >> 9 ) & 0x1ffu ) ), mulr( s, ( float )( ( p >> 18 ) & 0x1ffu ) ) );
}
ASVGF_DEVICE uint32_t encodeE5( float3 c ) // Utils.h:162-192; floor(log2) = the exponent field, exact
{
    c.x = clampf( c.x, 0.0f, 65408.0f );
    c.y = clampf( c.y, 0.0f, 65408.0f );
    c.z = clampf( c.z, 0.0f, 65408.0f );
    const float maxC = fmaxf( c.x, fmaxf( c.y, c.z ) );
    if( maxC == 0.0f ) return 0u;
    const int ef         = ( int )( ( __float_as_uint( maxC ) >> 23 ) & 0xffu ) - 127; // denormals give -127 -> clamped below
    int       expShared  = max( -16, ef ) + 16;
    float     s          = __uint_as_float( ( uint32_t )( 151 - expShared ) << 23 ); // 2^(24-expShared)
    const int maxS       = __float2int_rn( mulr( maxC, s ) );
    if( maxS == 512 )
    {
        expShared += 1;
        s = __uint_as_float( ( uint32_t )( 151 - expShared ) << 23 );
    }
    const uint32_t r = ( uint32_t )__float2int_rn( mulr( c.x, s ) );
    const uint32_t g = ( uint32_t )__float2int_rn( mulr( c.y, s ) );
    const uint32_t b = ( uint32_t )__float2int_rn( mulr( c.z, s ) );
    return ( ( uint32_t )expShared << 27 ) | ( b << 18 ) | ( g << 9 ) | r;
}

ASVGF_DEVICE uint32_t floatToUfloat( float f, uint32_t mbits ) // RTNE; NaN, negatives -> 0; overflow -> max finite
{
    const uint32_t maxCode = ( 30u << mbits ) | ( ( 1u << mbits ) - 1u );
    if( !( f > 0.0f ) ) return 0u;
    const uint32_t x = __float_as_uint( f );
    if( x >= 0x47800000u ) return maxCode;
    const uint32_t drop = 23u - mbits;
    if( x < 0x38800000u )
    {
        const uint32_t e     = x >> 23;
        const uint32_t shift = 126u - e + ( 10u - mbits );
        if( shift > 24u ) return 0u;
        const uint32_t m    = ( x & 0x7fffffu ) | 0x800000u;
        uint32_t       r    = m >> shift;
        const uint32_t rem  = m & ( ( 1u << shift ) - 1u );
        const uint32_t half = 1u << ( shift - 1u );
        if( rem > half || ( rem == half && ( r & 1u ) ) ) r++;
        return r;
    }
    uint32_t       r    = ( x - 0x38000000u ) >> drop;
    const uint32_t rem  = x & ( ( 1u << drop ) - 1u );
    const uint32_t half = 1u << ( drop - 1u );
    if( rem > half || ( rem == half && ( r & 1u ) ) ) r++;
    return r > maxCode ? maxCode : r;
}
ASVGF_DEVICE float ufloatToFloat( uint32_t c, uint32_t mbits )
{
    const uint32_t e = c >> mbits, m = c & ( ( 1u << mbits ) - 1u );
    if( e == 0 ) return mulr( ( float )m, __uint_as_float( ( 113u - mbits ) << 23 ) ); // m * 2^(-14-mbits)
    if( e == 31 ) return m == 0 ? __uint_as_float( 0x7f800000u ) : __uint_as_float( 0x7fc00000u );
    return __uint_as_float( ( ( e + 112u ) << 23 ) | ( m << ( 23u - mbits ) ) );
}
ASVGF_DEVICE uint32_t packUfloat111110( float3 c ) { return floatToUfloat( c.x, 6 ) | ( floatToUfloat( c.y, 6 ) << 11 ) | ( floatToUfloat( c.z, 5 ) << 22 ); }
ASVGF_DEVICE float3 unpackUfloat111110( uint32_t p ) { return make_float3( ufloatToFloat( p & 0x7ffu, 6 ), ufloatToFloat( ( p >> 11 ) & 0x7ffu, 6 ), ufloatToFloat( p >> 22, 5 ) ); }

ASVGF_DEVICE uint32_t floatToUnorm8( float f ) { return __float2uint_rn( mulr( fminf( fmaxf( f, 0.0f ), 1.0f ), 255.0f ) ); } // fmaxf(NaN,0)=0
ASVGF_DEVICE float    unorm8ToFloat( uint32_t c ) { return divr( ( float )c, 255.0f ); }
ASVGF_DEVICE float    h2f( uint16_t h ) { return __half2float( __ushort_as_half( h ) ); }
ASVGF_DEVICE uint16_t f2h( float f ) { return __half_as_ushort( __float2half_rn( f ) ); }

ASVGF_DEVICE float3 decodeNormal( const NormalLut& lut, uint32_t p ) // Utils.h:124-140
{
    const float2 a = __ldg( lut.phi + ( p >> 16 ) );      // sin(phi), cos(phi)
    const float2 b = __ldg( lut.theta + ( p & 0xffffu ) ); // sin(theta), cos(theta)
    return make_float3( mulr( a.x, b.y ), mulr( a.x, b.x ), a.y );
}

// ---- image access: out-of-image fetch -> 0, out-of-image store discarded ----
// inside the frame AND inside the rows this handle stores (a band handle holds rows [row0, row0 + rows) only)
ASVGF_DEVICE bool inImage( const Img& im, int x, int y ) { return ( unsigned )x < ( unsigned )im.w && ( unsigned )y < ( unsigned )im.h && ( unsigned )( y - im.row0 ) < ( unsigned )im.rows; }
template< typename T > ASVGF_DEVICE const T* texel( const Img& im, int x, int y )
{
    return reinterpret_cast< const T* >( im.p + ( size_t )( y - im.row0 ) * im.pitch ) + x;
}
template< typename T > ASVGF_DEVICE T* texelW( const Img& im, int x, int y )
{
    return reinterpret_cast< T* >( im.p + ( size_t )( y - im.row0 ) * im.pitch ) + x;
}
ASVGF_DEVICE uint32_t fetchU32( const Img& im, int x, int y ) { return inImage( im, x, y ) ? __ldg( texel< uint32_t >( im, x, y ) ) : 0u; }
ASVGF_DEVICE uint32_t fetchU8( const Img& im, int x, int y ) { return inImage( im, x, y ) ? ( uint32_t )__ldg( texel< uint8_t >( im, x, y ) ) : 0u; }
ASVGF_DEVICE float    fetchR16F( const Img& im, int x, int y ) { return inImage( im, x, y ) ? h2f( __ldg( texel< uint16_t >( im, x, y ) ) ) : 0.0f; }
ASVGF_DEVICE float2   fetchRG16F( const Img& im, int x, int y )
{
    if( !inImage( im, x, y ) ) return make_float2( 0.0f, 0.0f );
    const uint32_t v = __ldg( texel< uint32_t >( im, x, y ) );
    return make_float2( h2f( ( uint16_t )( v & 0xffffu ) ), h2f( ( uint16_t )( v >> 16 ) ) );
}
ASVGF_DEVICE float4 fetchRGBA16F( const Img& im, int x, int y )
{
    if( !inImage( im, x, y ) ) return make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
    const uint2 v = __ldg( texel< uint2 >( im, x, y ) );
    return make_float4( h2f( ( uint16_t )( v.x & 0xffffu ) ), h2f( ( uint16_t )( v.x >> 16 ) ), h2f( ( uint16_t )( v.y & 0xffffu ) ), h2f( ( uint16_t )( v.y >> 16 ) ) );
}
ASVGF_DEVICE float4 fetchRGBA32F( const Img& im, int x, int y ) { return inImage( im, x, y ) ? __ldg( texel< float4 >( im, x, y ) ) : make_float4( 0.0f, 0.0f, 0.0f, 0.0f ); }
ASVGF_DEVICE float2 fetchRG8( const Img& im, int x, int y )
{
    if( !inImage( im, x, y ) ) return make_float2( 0.0f, 0.0f );
    const uint16_t v = __ldg( texel< uint16_t >( im, x, y ) );
    return make_float2( unorm8ToFloat( v & 0xffu ), unorm8ToFloat( v >> 8 ) );
}
ASVGF_DEVICE float4 fetchRGBA8( const Img& im, int x, int y )
{
    if( !inImage( im, x, y ) ) return make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
    const uint32_t v = __ldg( texel< uint32_t >( im, x, y ) );
    return make_float4( unorm8ToFloat( v & 0xffu ), unorm8ToFloat( ( v >> 8 ) & 0xffu ), unorm8ToFloat( ( v >> 16 ) & 0xffu ), unorm8ToFloat( v >> 24 ) );
}
ASVGF_DEVICE float3 fetchE5( const Img& im, int x, int y ) { return decodeE5( fetchU32( im, x, y ) ); }

ASVGF_DEVICE void storeU32( const Img& im, int x, int y, uint32_t v ) { if( inImage( im, x, y ) ) *texelW< uint32_t >( im, x, y ) = v; }
ASVGF_DEVICE void storeU8( const Img& im, int x, int y, uint32_t v ) { if( inImage( im, x, y ) ) *texelW< uint8_t >( im, x, y ) = ( uint8_t )v; }
ASVGF_DEVICE void storeR16F( const Img& im, int x, int y, float v ) { if( inImage( im, x, y ) ) *texelW< uint16_t >( im, x, y ) = f2h( v ); }
ASVGF_DEVICE void storeRG16F( const Img& im, int x, int y, float a, float b )
{
    if( inImage( im, x, y ) ) *texelW< uint32_t >( im, x, y ) = ( uint32_t )f2h( a ) | ( ( uint32_t )f2h( b ) << 16 );
}
ASVGF_DEVICE void storeRGBA16F( const Img& im, int x, int y, float4 v )
{
    if( inImage( im, x, y ) )
        *texelW< uint2 >( im, x, y ) = make_uint2( ( uint32_t )f2h( v.x ) | ( ( uint32_t )f2h( v.y ) << 16 ), ( uint32_t )f2h( v.z ) | ( ( uint32_t )f2h( v.w ) << 16 ) );
}
ASVGF_DEVICE void storeRGBA8( const Img& im, int x, int y, float4 v )
{
    if( inImage( im, x, y ) )
        *texelW< uint32_t >( im, x, y ) = floatToUnorm8( v.x ) | ( floatToUnorm8( v.y ) << 8 ) | ( floatToUnorm8( v.z ) << 16 ) | ( floatToUnorm8( v.w ) << 24 );
}

// ---- checkerboard & reprojection, ShaderCommonGLSLFunc.h:267-390 ----
struct Area { int x0, x1, h; };
ASVGF_DEVICE Area checkerboardedRenderArea( const Uniform& u, int cbx )
{
    const int sep = u.W / 2, isOdd = cbx >= sep ? 1 : 0;
    return Area{ isOdd * sep, ( isOdd + 1 ) * sep, u.H };
}
ASVGF_DEVICE bool testPixInRenderArea( int x, int y, const Area& a ) { return y >= 0 && y < a.h && x >= a.x0 && x < a.x1; }
ASVGF_DEVICE int2 regularPixFromCheckerboardPix( const Uniform& u, int cbx, int cby )
{
    const int sep = u.W / 2, isOdd = cbx >= sep ? 1 : 0, x = cbx - isOdd * sep;
    return make_int2( x * 2 + ( isOdd + cby ) % 2, cby );
}
ASVGF_DEVICE bool testReprojectedDepth( float z, float zPrev, float zMotion ) { return divr( fabsf( addr( subr( z, zPrev ), zMotion ) ), fabsf( z ) ) < 0.1f; }
ASVGF_DEVICE bool testReprojectedNormal( float3 n, float3 nPrev ) { return dot3r( n, nPrev ) > 0.95f; }
ASVGF_DEVICE float2 prevScreenPos( const Uniform& u, float mx, float my, int px, int py )
{
    const float sx = divr( u.renderWidth, 2.0f ), sy = u.renderHeight;
    const float ix = divr( 1.0f, sx ), iy = divr( 1.0f, sy );
    return make_float2( mulr( addr( mulr( addr( ( float )px, 0.5f ), ix ), mx ), sx ), mulr( addr( mulr( addr( ( float )py, 0.5f ), iy ), my ), sy ) );
}
ASVGF_DEVICE float antilagAlpha( float gradSample, float normFactor )
{
    const float lambda = normFactor > 0.0001f ? clampf( divr( fabsf( gradSample ), normFactor ), 0.0f, 1.0f ) : 0.0f;
    return clampf( lambda, 0.0f, 1.0f );
}

} // namespace asvgf
