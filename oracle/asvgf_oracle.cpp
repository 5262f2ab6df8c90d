/*
 * asvgf_oracle.cpp -- CPU oracle of RTGL1's A-SVGF denoiser path.  TEST INFRASTRUCTURE ONLY (see asvgf_oracle.h).
 *
 * A scalar fp32 restatement, pass by pass, of the reference's GLSL (paths relative to /root/reference):
 *   gradientSamples() ... Source/Shaders/RtGradients.rgen:44-181        (ray-traced re-shade :107-136 -> FwdLum image)
 *   gradientAtrous() .... Source/Shaders/CmASVGFGradientAtrous.comp:42-105
 *   temporal() .......... Source/Shaders/CmSVGFTemporalAccumulation.comp:58-385
 *   antiFirefly() ....... Source/Shaders/CmAntiFirefly.comp:47-159
 *   variance() .......... Source/Shaders/CmSVGFEstimateVariance.comp:74-205
 *   atrous() i==0 ....... Source/Shaders/CmSVGFAtrous_Iter0.comp:92-335
 *   atrous() i>=1 ....... Source/Shaders/CmSVGFAtrous.comp:65-349      (last iteration composes PreFinal)
 *   helpers ............. Source/Shaders/Utils.h:54-72,124-140,162-207; ShaderCommonGLSLFunc.h:46-47,267-390,492-495;
 *                         Source/Shaders/BRDF.h:27-29,49-64
 *   pass order/routing .. Source/Denoiser.cpp:87-254;  cur/prev swap: Source/Framebuffers.cpp:35-50
 *
 * PARITY PINNED: the reference ships no golden vectors, but its shaders are plain GLSL, and oracle/_ref executes their
 * UNMODIFIED text (compiled as C++ over the reference's vendored GLM, oracle/ref_glsl/build_ref.py).  Every dispatch of
 * Denoiser::Denoise and the post dispatches are bit-identical between this file and those shaders from identical state
 * (tests/test_ref_pins_oracle.py: sizes 6x7 ... 322x182, the light switch, every debug flag, NaN/Inf garbage inputs), and
 * the fixtures that executor generated are committed (tests/golden/asvgf_ref_96x54.npz).  Only gradientSamples() -- a
 * ray-generation shader with no compute form -- is pinned by restatement and known-answer tests alone.
 * Behaviours that GLSL/Vulkan leave implementation-defined are pinned as follows (oracle/_ref and the CUDA path pin the same):
 *   - out-of-bounds texelFetch/imageLoad returns an all-zero texel; out-of-bounds imageStore is discarded
 *   - f32 -> f16 / unorm8 / uf11 / uf10 on imageStore: round-to-nearest-even; uf1x: NaN and negatives -> 0, overflow -> max finite
 *   - GLSL round() = round-half-even; mix(a,b,t) = a*(1-t) + b*t; dot(a,b) = (ax*bx + ay*by) + az*bz;
 *     mat*vec = (c0*x + c1*y) + (c2*z + c3*w), GLM's order; no fused multiply-add anywhere (build with -ffp-contract=off)
 *   - sin/cos of decodeNormal: correctly rounded fp32 of the fp32 angle; pow/exp/exp2/log: libm fp32
 *   - floor(log2(x)) of the E5B9G9R9 encoder: the exact exponent of x (what exact arithmetic gives)
 *   - float -> int conversions saturate at +-1e9 (out-of-range is undefined in GLSL)
 */
#include "asvgf_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
    #include <omp.h>
#endif

namespace
{

// ---------------------------------------------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------------------------------------------
struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };
struct ivec2 { int x, y; };
struct ivec3 { int x, y, z; };

inline uint32_t f2u( float f ) { uint32_t u; std::memcpy( &u, &f, 4 ); return u; }
inline float    u2f( uint32_t u ) { float f; std::memcpy( &f, &u, 4 ); return f; }

inline vec3 operator+( vec3 a, vec3 b ) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline vec3 operator-( vec3 a, vec3 b ) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline vec3 operator*( vec3 a, float s ) { return { a.x * s, a.y * s, a.z * s }; }
inline vec3 operator*( float s, vec3 a ) { return { a.x * s, a.y * s, a.z * s }; }
inline vec3 operator*( vec3 a, vec3 b ) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline vec3 operator-( vec3 a ) { return { -a.x, -a.y, -a.z }; }
inline vec2 operator+( vec2 a, vec2 b ) { return { a.x + b.x, a.y + b.y }; }
inline vec2 operator*( vec2 a, float s ) { return { a.x * s, a.y * s }; }
inline vec4 operator*( vec4 a, float s ) { return { a.x * s, a.y * s, a.z * s, a.w * s }; }
inline vec4 operator+( vec4 a, vec4 b ) { return { a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w }; }

inline float dot3( vec3 a, vec3 b ) { return ( a.x * b.x + a.y * b.y ) + a.z * b.z; }
inline float clampf( float x, float lo, float hi ) { return fminf( fmaxf( x, lo ), hi ); }
inline float saturate( float x ) { return clampf( x, 0.0f, 1.0f ); }                  // Utils.h:59-62
inline float mixf( float a, float b, float t ) { return a * ( 1.0f - t ) + b * t; }
inline vec3  mix3( vec3 a, vec3 b, float t ) { return { mixf( a.x, b.x, t ), mixf( a.y, b.y, t ), mixf( a.z, b.z, t ) }; }
inline vec2  mix2( vec2 a, vec2 b, float t ) { return { mixf( a.x, b.x, t ), mixf( a.y, b.y, t ) }; }
inline float square( float x ) { return x * x; }                                      // Utils.h:74-77
inline float safePositiveRcp( float f ) { return f <= 0.0f ? 0.0f : 1.0f / f; }      // Utils.h:69-72
inline float getLuminance( vec3 c ) { return ( 0.2125f * c.x + 0.7154f * c.y ) + 0.0721f * c.z; } // Utils.h:54-57
inline vec3  normalize3( vec3 v ) { float inv = 1.0f / sqrtf( dot3( v, v ) ); return v * inv; }
inline float smoothstepf( float e0, float e1, float x )
{
    float t = saturate( ( x - e0 ) / ( e1 - e0 ) );
    return ( t * t ) * ( 3.0f - 2.0f * t );
}
inline int f2i( float f ) // saturating float->int (pinned, see header)
{
    if( !( f > -1.0e9f ) ) return -1000000000;
    if( f > 1.0e9f ) return 1000000000;
    return ( int )f;
}
inline float glslPow( float x, float y ) { return powf( x, y ); }

// column-major mat4 * vec4.  GLSL does not fix the order of the four-term sums; the pin is the order of the reference's
// own vendored GLM (Tests/Libs/glm/glm/detail/type_mat4x4.inl, mat * vec: (c0*x + c1*y) + (c2*z + c3*w)), which is what
// oracle/_ref executes
inline vec4 mulMat( const float* m, vec4 v )
{
    vec4 r;
    r.x = ( m[ 0 ] * v.x + m[ 4 ] * v.y ) + ( m[ 8 ] * v.z + m[ 12 ] * v.w );
    r.y = ( m[ 1 ] * v.x + m[ 5 ] * v.y ) + ( m[ 9 ] * v.z + m[ 13 ] * v.w );
    r.z = ( m[ 2 ] * v.x + m[ 6 ] * v.y ) + ( m[ 10 ] * v.z + m[ 14 ] * v.w );
    r.w = ( m[ 3 ] * v.x + m[ 7 ] * v.y ) + ( m[ 11 ] * v.z + m[ 15 ] * v.w );
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// storage-format conversions
// ---------------------------------------------------------------------------------------------------------------
uint16_t floatToHalf( float f ) // round-to-nearest-even
{
    uint32_t x    = f2u( f );
    uint32_t sign = ( x >> 16 ) & 0x8000u;
    x &= 0x7fffffffu;
    if( x >= 0x7f800000u )
    {
        return ( uint16_t )( sign | 0x7c00u | ( x > 0x7f800000u ? ( 0x200u | ( ( x >> 13 ) & 0x3ffu ) ) : 0u ) );
    }
    if( x >= 0x47800000u ) // >= 65536
    {
        return ( uint16_t )( sign | 0x7c00u );
    }
    if( x < 0x38800000u ) // below the smallest normal half 2^-14
    {
        if( x < 0x33000000u ) // < 2^-25 rounds to zero
        {
            return ( uint16_t )sign;
        }
        uint32_t e     = x >> 23;
        uint32_t m     = ( x & 0x7fffffu ) | 0x800000u;
        uint32_t shift = 126u - e; // 14..24
        uint32_t r     = m >> shift;
        uint32_t rem   = m & ( ( 1u << shift ) - 1u );
        uint32_t half  = 1u << ( shift - 1u );
        if( rem > half || ( rem == half && ( r & 1u ) ) ) r++;
        return ( uint16_t )( sign | r );
    }
    uint32_t r   = ( x - 0x38000000u ) >> 13;
    uint32_t rem = x & 0x1fffu;
    if( rem > 0x1000u || ( rem == 0x1000u && ( r & 1u ) ) ) r++; // a carry into the exponent is correct, up to inf
    return ( uint16_t )( sign | r );
}

float halfToFloat( uint16_t h )
{
    uint32_t sign = ( uint32_t )( h & 0x8000u ) << 16;
    uint32_t e    = ( h >> 10 ) & 0x1fu;
    uint32_t m    = h & 0x3ffu;
    if( e == 0 )
    {
        if( m == 0 ) return u2f( sign );
        float v = ( float )m * 5.9604644775390625e-08f; // m * 2^-24, exact
        return sign ? -v : v;
    }
    if( e == 31 ) return u2f( sign | 0x7f800000u | ( m << 13 ) );
    return u2f( sign | ( ( e + 112u ) << 23 ) | ( m << 13 ) );
}

// unsigned small floats of B10G11R11: 5 exponent bits (bias 15), mbits mantissa bits, no sign
uint32_t floatToUfloat( float f, uint32_t mbits )
{
    const uint32_t maxCode = ( 30u << mbits ) | ( ( 1u << mbits ) - 1u ); // largest finite
    if( !( f > 0.0f ) ) return 0u;                                         // NaN, negatives, zeros
    uint32_t x = f2u( f );
    if( x >= 0x47800000u ) return maxCode; // >= 65536, inf
    uint32_t drop = 23u - mbits;
    if( x < 0x38800000u ) // < 2^-14: denormal, unit 2^(-14-mbits)
    {
        uint32_t e     = x >> 23;
        uint32_t shift = 126u - e + ( 10u - mbits ); // as for half (mbits=10), coarser by 10-mbits
        if( shift > 24u ) return 0u;
        uint32_t m    = ( x & 0x7fffffu ) | 0x800000u;
        uint32_t r    = m >> shift;
        uint32_t rem  = m & ( ( 1u << shift ) - 1u );
        uint32_t half = 1u << ( shift - 1u );
        if( rem > half || ( rem == half && ( r & 1u ) ) ) r++;
        return r;
    }
    uint32_t r    = ( x - 0x38000000u ) >> drop;
    uint32_t rem  = x & ( ( 1u << drop ) - 1u );
    uint32_t half = 1u << ( drop - 1u );
    if( rem > half || ( rem == half && ( r & 1u ) ) ) r++;
    return r > maxCode ? maxCode : r;
}

float ufloatToFloat( uint32_t c, uint32_t mbits )
{
    uint32_t e = c >> mbits;
    uint32_t m = c & ( ( 1u << mbits ) - 1u );
    if( e == 0 ) return ( float )m * ldexpf( 1.0f, -14 - ( int )mbits );
    if( e == 31 ) return m == 0 ? INFINITY : NAN;
    return u2f( ( ( e + 112u ) << 23 ) | ( m << ( 23u - mbits ) ) );
}

uint32_t packUfloat111110( vec3 c ) // VK_FORMAT_B10G11R11_UFLOAT_PACK32: R bits 0-10, G 11-21, B 22-31
{
    return floatToUfloat( c.x, 6 ) | ( floatToUfloat( c.y, 6 ) << 11 ) | ( floatToUfloat( c.z, 5 ) << 22 );
}
vec3 unpackUfloat111110( uint32_t p )
{
    return { ufloatToFloat( p & 0x7ffu, 6 ), ufloatToFloat( ( p >> 11 ) & 0x7ffu, 6 ), ufloatToFloat( p >> 22, 5 ) };
}

uint8_t floatToUnorm8( float f )
{
    if( !( f > 0.0f ) ) return 0; // NaN, negatives
    if( f >= 1.0f ) return 255;
    return ( uint8_t )rintf( f * 255.0f );
}
inline float unorm8ToFloat( uint8_t c ) { return ( float )c / 255.0f; }

// Utils.h:162-192  (EXT_texture_shared_exponent reference algorithm)
uint32_t encodeE5B9G9R9( vec3 c )
{
    const int   N = 9, Np2 = 1 << N, B = 15;
    const float sharedExpMax = 65408.0f;
    c.x = clampf( c.x, 0.0f, sharedExpMax );
    c.y = clampf( c.y, 0.0f, sharedExpMax );
    c.z = clampf( c.z, 0.0f, sharedExpMax );
    float maxC = fmaxf( c.x, fmaxf( c.y, c.z ) );
    if( maxC == 0.0f ) return 0u;
    int expSharedP = std::max( -B - 1, ( int )ilogbf( maxC ) ) + 1 + B; // floor(log2(maxC)), exact
    int maxS       = ( int )rintf( maxC * ldexpf( 1.0f, -( expSharedP - B - N ) ) );
    int expShared  = maxS != Np2 ? expSharedP : expSharedP + 1;
    float    s     = ldexpf( 1.0f, -( expShared - B - N ) );
    uint32_t r = ( uint32_t )rintf( c.x * s ), g = ( uint32_t )rintf( c.y * s ), b = ( uint32_t )rintf( c.z * s );
    return ( ( uint32_t )expShared << 27 ) | ( b << 18 ) | ( g << 9 ) | r;
}

// Utils.h:194-207
vec3 decodeE5B9G9R9( uint32_t p )
{
    int   expShared = ( int )( p >> 27 );
    float s         = ldexpf( 1.0f, expShared - 15 - 9 );
    return { s * ( float )( p & 0x1ffu ), s * ( float )( ( p >> 9 ) & 0x1ffu ), s * ( float )( ( p >> 18 ) & 0x1ffu ) };
}

const float PI_F = 3.14159265358979323846f; // GLSL has no double literals here: M_PI is the nearest float

// Utils.h:108-122
uint32_t encodeNormal( vec3 n )
{
    float phi   = acosf( n.z );
    float theta = atan2f( n.y, n.x );
    theta       = theta < 0.0f ? theta + 2.0f * PI_F : theta;
    uint32_t j  = ( uint32_t )f2i( rintf( phi * 65535.0f / PI_F ) );
    uint32_t k  = ( uint32_t )f2i( rintf( theta * 65536.0f / ( 2.0f * PI_F ) ) ) % 65536u;
    return ( j << 16 ) | k;
}

// Utils.h:124-140
vec3 decodeNormal( uint32_t p )
{
    uint32_t j     = p >> 16;
    uint32_t k     = p & 0xffffu;
    float    phi   = ( float )j * PI_F / 65535.0f;
    float    theta = ( float )( k * 2u ) * PI_F / 65536.0f;
    float sp = ( float )sin( ( double )phi ), cp = ( float )cos( ( double )phi );
    float st = ( float )sin( ( double )theta ), ct = ( float )cos( ( double )theta );
    return { sp * ct, sp * st, cp };
}

// ---------------------------------------------------------------------------------------------------------------
// images
// ---------------------------------------------------------------------------------------------------------------
struct Img
{
    int                    w = 0, h = 0, bpp = 0;
    std::vector< uint8_t > d;
    void alloc( int w_, int h_, int bpp_ ) { w = w_; h = h_; bpp = bpp_; d.assign( ( size_t )w * h * bpp, 0 ); }
    bool inside( int x, int y ) const { return x >= 0 && y >= 0 && x < w && y < h; }
    const uint8_t* at( int x, int y ) const { return d.data() + ( ( size_t )y * w + x ) * bpp; }
    uint8_t*       at( int x, int y ) { return d.data() + ( ( size_t )y * w + x ) * bpp; }
};

inline uint32_t fetchU32( const Img& im, ivec2 p )
{
    if( !im.inside( p.x, p.y ) ) return 0u;
    uint32_t v; std::memcpy( &v, im.at( p.x, p.y ), 4 ); return v;
}
inline uint32_t fetchU8( const Img& im, ivec2 p ) { return im.inside( p.x, p.y ) ? *im.at( p.x, p.y ) : 0u; }
inline float fetchR16F( const Img& im, ivec2 p )
{
    if( !im.inside( p.x, p.y ) ) return 0.0f;
    uint16_t v; std::memcpy( &v, im.at( p.x, p.y ), 2 ); return halfToFloat( v );
}
inline vec2 fetchRG16F( const Img& im, ivec2 p )
{
    if( !im.inside( p.x, p.y ) ) return { 0, 0 };
    uint16_t v[ 2 ]; std::memcpy( v, im.at( p.x, p.y ), 4 ); return { halfToFloat( v[ 0 ] ), halfToFloat( v[ 1 ] ) };
}
inline vec4 fetchRGBA16F( const Img& im, ivec2 p )
{
    if( !im.inside( p.x, p.y ) ) return { 0, 0, 0, 0 };
    uint16_t v[ 4 ]; std::memcpy( v, im.at( p.x, p.y ), 8 );
    return { halfToFloat( v[ 0 ] ), halfToFloat( v[ 1 ] ), halfToFloat( v[ 2 ] ), halfToFloat( v[ 3 ] ) };
}
inline vec4 fetchRGBA32F( const Img& im, ivec2 p )
{
    if( !im.inside( p.x, p.y ) ) return { 0, 0, 0, 0 };
    vec4 v; std::memcpy( &v, im.at( p.x, p.y ), 16 ); return v;
}
inline vec2 fetchRG8( const Img& im, ivec2 p )
{
    if( !im.inside( p.x, p.y ) ) return { 0, 0 };
    const uint8_t* s = im.at( p.x, p.y ); return { unorm8ToFloat( s[ 0 ] ), unorm8ToFloat( s[ 1 ] ) };
}
inline vec4 fetchRGBA8( const Img& im, ivec2 p )
{
    if( !im.inside( p.x, p.y ) ) return { 0, 0, 0, 0 };
    const uint8_t* s = im.at( p.x, p.y );
    return { unorm8ToFloat( s[ 0 ] ), unorm8ToFloat( s[ 1 ] ), unorm8ToFloat( s[ 2 ] ), unorm8ToFloat( s[ 3 ] ) };
}
inline vec3 fetchE5( const Img& im, ivec2 p ) { return decodeE5B9G9R9( fetchU32( im, p ) ); }
inline vec3 fetchUf11( const Img& im, ivec2 p ) { return unpackUfloat111110( fetchU32( im, p ) ); }

inline void storeU32( Img& im, ivec2 p, uint32_t v ) { if( im.inside( p.x, p.y ) ) std::memcpy( im.at( p.x, p.y ), &v, 4 ); }
inline void storeU8( Img& im, ivec2 p, uint8_t v ) { if( im.inside( p.x, p.y ) ) *im.at( p.x, p.y ) = v; }
inline void storeR16F( Img& im, ivec2 p, float v )
{
    if( !im.inside( p.x, p.y ) ) return;
    uint16_t h = floatToHalf( v ); std::memcpy( im.at( p.x, p.y ), &h, 2 );
}
inline void storeRG16F( Img& im, ivec2 p, vec2 v )
{
    if( !im.inside( p.x, p.y ) ) return;
    uint16_t h[ 2 ] = { floatToHalf( v.x ), floatToHalf( v.y ) }; std::memcpy( im.at( p.x, p.y ), h, 4 );
}
inline void storeRGBA16F( Img& im, ivec2 p, vec4 v )
{
    if( !im.inside( p.x, p.y ) ) return;
    uint16_t h[ 4 ] = { floatToHalf( v.x ), floatToHalf( v.y ), floatToHalf( v.z ), floatToHalf( v.w ) };
    std::memcpy( im.at( p.x, p.y ), h, 8 );
}
inline void storeRGBA8( Img& im, ivec2 p, vec4 v )
{
    if( !im.inside( p.x, p.y ) ) return;
    uint8_t* d = im.at( p.x, p.y );
    d[ 0 ] = floatToUnorm8( v.x ); d[ 1 ] = floatToUnorm8( v.y ); d[ 2 ] = floatToUnorm8( v.z ); d[ 3 ] = floatToUnorm8( v.w );
}
inline void storeE5( Img& im, ivec2 p, vec3 c ) { storeU32( im, p, encodeE5B9G9R9( c ) ); }

// ---------------------------------------------------------------------------------------------------------------
// checkerboard / reprojection helpers, ShaderCommonGLSLFunc.h:267-390
// ---------------------------------------------------------------------------------------------------------------
struct Ctx
{
    const AsvgfGlobalUniform* u;
    int                       W, H; // int(renderWidth), int(renderHeight)
};

inline int   getCheckerboardSeparatorX( const Ctx& c ) { return c.W / 2; }                       // :306-309
inline ivec2 getCheckerboardPix( const Ctx& c, ivec2 pix )                                        // :321-329
{
    int isOdd = ( pix.x + pix.y % 2 ) % 2;
    return { isOdd * getCheckerboardSeparatorX( c ) + pix.x / 2, pix.y };
}
inline ivec2 getRegularPixFromCheckerboardPix( const Ctx& c, ivec2 cb )                           // :331-342
{
    int sep   = getCheckerboardSeparatorX( c );
    int isOdd = cb.x >= sep ? 1 : 0;
    int x     = cb.x - isOdd * sep;
    return { x * 2 + ( isOdd + cb.y ) % 2, cb.y };
}
inline ivec3 getCheckerboardedRenderArea( const Ctx& c, ivec2 cb )                                // :345-358
{
    int sep   = getCheckerboardSeparatorX( c );
    int isOdd = cb.x >= sep ? 1 : 0;
    return { ( isOdd + 0 ) * sep, ( isOdd + 1 ) * sep, c.H };
}
inline bool testPixInRenderArea( ivec2 p, ivec3 a )                                               // :361-366
{
    return p.y >= 0 && p.y < a.z && p.x >= a.x && p.x < a.y;
}
inline bool testReprojectedDepth( float z, float zPrev, float zMotion )                          // :368-371
{
    return fabsf( z - zPrev + zMotion ) / fabsf( z ) < 0.1f;
}
inline bool testReprojectedNormal( vec3 n, vec3 nPrev ) { return dot3( n, nPrev ) > 0.95f; }      // :373-376
inline float getAntilagAlpha( float gradSample, float normFactor )                               // :383-390
{
    float lambda = normFactor > 0.0001f ? clampf( fabsf( gradSample ) / normFactor, 0.0f, 1.0f ) : 0.0f;
    return clampf( lambda, 0.0f, 1.0f );
}
inline vec2 getPrevScreenPos( const Ctx& c, vec2 motion, ivec2 pix )                             // :271-277
{
    vec2 screenSize    = { c.u->renderWidth / 2.0f, c.u->renderHeight };
    vec2 invScreenSize = { 1.0f / screenSize.x, 1.0f / screenSize.y };
    return { ( ( ( float )pix.x + 0.5f ) * invScreenSize.x + motion.x ) * screenSize.x,
             ( ( ( float )pix.y + 0.5f ) * invScreenSize.y + motion.y ) * screenSize.y };
}
inline bool classicShading( const Ctx& c, ivec2 regularPix )                                      // :492-495
{
    return ( float )regularPix.x < c.u->lightmapScreenCoverage * c.u->renderWidth;
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------
// oracle state: the Framebuffers mirror
// ---------------------------------------------------------------------------------------------------------------
struct AsvgfOracle_T
{
    uint32_t W = 0, H = 0, iterations = 4;
    Img      img[ ASVGF_IMAGE_COUNT ];
    AsvgfTonemapping tm = {}; // the reference's tonemapping buffer (ShTonemapping, Tonemapping.cpp:183-191)

    // Framebuffers.cpp:35-50: images with a _PREV twin are addressed as (index + frameIndex) / (index - frameIndex)
    static int swapKind( uint32_t i ) // 0 none, 1 = cur of a pair, 2 = prev of a pair
    {
        switch( i )
        {
            case ASVGF_IMAGE_NORMAL: case ASVGF_IMAGE_METALLIC_ROUGHNESS: case ASVGF_IMAGE_DEPTH_WORLD:
            case ASVGF_IMAGE_SURFACE_POSITION: case ASVGF_IMAGE_ACCUM_HISTORY_LENGTH: case ASVGF_IMAGE_DIFF_ACCUM_COLOR:
            case ASVGF_IMAGE_DIFF_ACCUM_MOMENTS: case ASVGF_IMAGE_SPEC_ACCUM_COLOR: case ASVGF_IMAGE_INDIR_ACCUM:
            case ASVGF_IMAGE_GRADIENT_INPUTS: case ASVGF_IMAGE_SCATTERING: case ASVGF_IMAGE_SCATTERING_HISTORY: return 1;
            case ASVGF_IMAGE_NORMAL_PREV: case ASVGF_IMAGE_METALLIC_ROUGHNESS_PREV: case ASVGF_IMAGE_DEPTH_WORLD_PREV:
            case ASVGF_IMAGE_SURFACE_POSITION_PREV: case ASVGF_IMAGE_ACCUM_HISTORY_LENGTH_PREV:
            case ASVGF_IMAGE_DIFF_ACCUM_COLOR_PREV: case ASVGF_IMAGE_DIFF_ACCUM_MOMENTS_PREV:
            case ASVGF_IMAGE_SPEC_ACCUM_COLOR_PREV: case ASVGF_IMAGE_INDIR_ACCUM_PREV:
            case ASVGF_IMAGE_GRADIENT_INPUTS_PREV: case ASVGF_IMAGE_SCATTERING_PREV: case ASVGF_IMAGE_SCATTERING_HISTORY_PREV: return 2;
            default: return 0;
        }
    }
    Img& get( uint32_t index, uint32_t frameIndex )
    {
        int k = swapKind( index );
        if( k == 1 ) return img[ index + frameIndex ];
        if( k == 2 ) return img[ index - frameIndex ];
        return img[ index ];
    }
};

namespace
{

int bytesPerPixelOf( uint32_t i )
{
    switch( i )
    {
        case ASVGF_IMAGE_IS_SKY: case ASVGF_IMAGE_GRADIENT_PREV_PIX: case ASVGF_IMAGE_DEBUG_MASK: return 1;
        case ASVGF_IMAGE_METALLIC_ROUGHNESS: case ASVGF_IMAGE_METALLIC_ROUGHNESS_PREV: case ASVGF_IMAGE_DEPTH_WORLD:
        case ASVGF_IMAGE_DEPTH_WORLD_PREV: case ASVGF_IMAGE_DEPTH_GRAD: case ASVGF_IMAGE_ATROUS_FILTERED_VARIANCE:
        case ASVGF_IMAGE_HISTOGRAM_INPUT: case ASVGF_IMAGE_SCATTERING_HISTORY: case ASVGF_IMAGE_SCATTERING_HISTORY_PREV: return 2;
        case ASVGF_IMAGE_MOTION: case ASVGF_IMAGE_VIEW_DIRECTION: case ASVGF_IMAGE_THROUGHPUT:
        case ASVGF_IMAGE_ACCUM_HISTORY_LENGTH: case ASVGF_IMAGE_ACCUM_HISTORY_LENGTH_PREV:
        case ASVGF_IMAGE_DIFF_COLOR_HISTORY: case ASVGF_IMAGE_DIFF_PING_COLOR_AND_VARIANCE:
        case ASVGF_IMAGE_DIFF_PONG_COLOR_AND_VARIANCE: case ASVGF_IMAGE_SCATTERING: case ASVGF_IMAGE_SCATTERING_PREV:
        case ASVGF_IMAGE_SCATTER_CURRENT: return 8;
        case ASVGF_IMAGE_SURFACE_POSITION: case ASVGF_IMAGE_SURFACE_POSITION_PREV: return 16;
        case ASVGF_IMAGE_ALBEDO: case ASVGF_IMAGE_NORMAL: case ASVGF_IMAGE_NORMAL_PREV:
        case ASVGF_IMAGE_UNFILTERED_DIRECT: case ASVGF_IMAGE_UNFILTERED_SPECULAR: case ASVGF_IMAGE_UNFILTERED_INDIR:
        case ASVGF_IMAGE_ACID_FOG_RT: case ASVGF_IMAGE_ACID_FOG: case ASVGF_IMAGE_SCREEN_EMIS_RT: case ASVGF_IMAGE_SCREEN_EMISSION: case ASVGF_IMAGE_BLOOM_INPUT:
        case ASVGF_IMAGE_PRE_FINAL: case ASVGF_IMAGE_FINAL: case ASVGF_IMAGE_DIFF_TEMPORARY: case ASVGF_IMAGE_DIFF_ACCUM_COLOR:
        case ASVGF_IMAGE_DIFF_ACCUM_COLOR_PREV: case ASVGF_IMAGE_DIFF_ACCUM_MOMENTS: case ASVGF_IMAGE_DIFF_ACCUM_MOMENTS_PREV:
        case ASVGF_IMAGE_SPEC_ACCUM_COLOR: case ASVGF_IMAGE_SPEC_ACCUM_COLOR_PREV: case ASVGF_IMAGE_SPEC_PING_COLOR:
        case ASVGF_IMAGE_SPEC_PONG_COLOR: case ASVGF_IMAGE_INDIR_ACCUM: case ASVGF_IMAGE_INDIR_ACCUM_PREV:
        case ASVGF_IMAGE_INDIR_PING: case ASVGF_IMAGE_INDIR_PONG: case ASVGF_IMAGE_GRADIENT_INPUTS:
        case ASVGF_IMAGE_GRADIENT_INPUTS_PREV: case ASVGF_IMAGE_DIS_PING_GRADIENT: case ASVGF_IMAGE_DIS_PONG_GRADIENT:
        case ASVGF_IMAGE_DIS_GRADIENT_HISTORY: case ASVGF_IMAGE_FWD_LUM: return 4;
        default: return 0;
    }
}
bool isThirdRes( uint32_t i ) { return i >= ASVGF_IMAGE_DIS_PING_GRADIENT && i <= ASVGF_IMAGE_GRADIENT_PREV_PIX; }

// ---------------------------------------------------------------------------------------------------------------
// G: gradient samples, RtGradients.rgen:44-181
// ---------------------------------------------------------------------------------------------------------------
void gradientSamples( AsvgfOracle_T& o, uint32_t fi, const Ctx& c )
{
    Img&       gradientPrevPix    = o.get( ASVGF_IMAGE_GRADIENT_PREV_PIX, fi );
    Img&       disPing            = o.get( ASVGF_IMAGE_DIS_PING_GRADIENT, fi );
    const Img& disHistory         = o.get( ASVGF_IMAGE_DIS_GRADIENT_HISTORY, fi );
    const Img& motion             = o.get( ASVGF_IMAGE_MOTION, fi );
    const Img& gradientInputsPrev = o.get( ASVGF_IMAGE_GRADIENT_INPUTS_PREV, fi );
    const Img& isSky              = o.get( ASVGF_IMAGE_IS_SKY, fi );
    const Img& indirAccumPrev     = o.get( ASVGF_IMAGE_INDIR_ACCUM_PREV, fi );
    const Img& fwdLum             = o.get( ASVGF_IMAGE_FWD_LUM, fi );
    const int  S                  = ASVGF_STRATA_SIZE;

#pragma omp parallel for schedule( dynamic, 4 )
    for( int gy = 0; gy < disPing.h; gy++ )
    {
        for( int gx = 0; gx < disPing.w; gx++ )
        {
            const ivec2 gradPix = { gx, gy };
            if( gradPix.x * S >= c.W || gradPix.y * S >= c.H ) continue; // :49-52

            if( c.u->gradientMultDiffuse < 0.001f && c.u->gradientMultIndirect < 0.001f &&
                c.u->gradientMultSpecular < 0.001f ) // :54-61
            {
                storeRGBA8( disPing, gradPix, { 0, 0, 0, 0 } );
                storeU8( gradientPrevPix, gradPix, ASVGF_GRADIENT_PREV_PIX_NONE );
                continue;
            }

            const uint32_t prevStrataPos = fetchU8( gradientPrevPix, gradPix ); // :63

            // (a) instead of a random sample in the stratum, the brightest; :66-96
            ivec2    prevShadingPix = { 0, 0 };
            uint32_t curStrataPos   = 0;
            ivec2    curShadingPix  = { 0, 0 };
            vec2     prevLuminance  = { 0, 0 };
            for( int yy = 0; yy < S; yy++ )
            {
                for( int xx = 0; xx < S; xx++ )
                {
                    if( ( uint32_t )( yy * S + xx ) == prevStrataPos ) continue;

                    const ivec2 iterCurPix = { gradPix.x * S + xx, gradPix.y * S + yy };
                    const vec4  m          = fetchRGBA16F( motion, iterCurPix );
                    const vec2  pp         = getPrevScreenPos( c, { m.x, m.y }, iterCurPix );
                    const ivec2 iterPrevPix = { f2i( floorf( pp.x ) ), f2i( floorf( pp.y ) ) };
                    const vec2  iterPrevLum = fetchRG16F( gradientInputsPrev, iterPrevPix );

                    if( fmaxf( iterPrevLum.x, iterPrevLum.y ) > fmaxf( prevLuminance.x, prevLuminance.y ) )
                    {
                        prevShadingPix = iterPrevPix;
                        curShadingPix  = iterCurPix;
                        curStrataPos   = ( uint32_t )( yy * S + xx );
                        prevLuminance  = iterPrevLum;
                    }
                }
            }

            if( !testPixInRenderArea( prevShadingPix, getCheckerboardedRenderArea( c, curShadingPix ) ) ||
                fetchU8( isSky, curShadingPix ) != 0 || ( prevLuminance.x <= 0.0f && prevLuminance.y <= 0.0f ) ) // :98-105
            {
                storeRGBA8( disPing, gradPix, { 0, 0, 0, 0 } );
                storeU8( gradientPrevPix, gradPix, ASVGF_GRADIENT_PREV_PIX_NONE );
                continue;
            }

            // :107-136 replaced: luminance of the previous sample re-shaded in the current frame is an input image
            const vec2 fwd = fetchRG16F( fwdLum, curShadingPix );

            float gradDIS[ 3 ] = { 0, 0, 0 };
            { // :141-146
                float gradSample = fwd.x - prevLuminance.x;
                float normFactor = fmaxf( fwd.x, prevLuminance.x );
                gradDIS[ 0 ]     = getAntilagAlpha( gradSample, normFactor );
            }
            { // :147-157
                float gradSample = fmaxf( 0.0f, prevLuminance.x - fwd.x );
                float normFactor = fmaxf( fwd.x, prevLuminance.x );
                float brightness = getLuminance( fetchE5( indirAccumPrev, prevShadingPix ) );
                gradSample *= 1.0f - smoothstepf( 1.0f, 10.0f, brightness );
                gradDIS[ 1 ] = getAntilagAlpha( gradSample, normFactor );
            }
            { // :158-163
                float gradSample = fwd.y - prevLuminance.y;
                float normFactor = fmaxf( fwd.y, prevLuminance.y );
                gradDIS[ 2 ]     = getAntilagAlpha( gradSample, normFactor );
            }
            gradDIS[ 0 ] *= c.u->gradientMultDiffuse; // :166-168
            gradDIS[ 1 ] *= c.u->gradientMultIndirect;
            gradDIS[ 2 ] *= c.u->gradientMultSpecular;
            { // :171-177
                const ivec2 gradPixPrev = { prevShadingPix.x / S, prevShadingPix.y / S };
                const vec4  prev        = fetchRGBA8( disHistory, gradPixPrev );
                gradDIS[ 0 ]            = mixf( gradDIS[ 0 ], prev.x, 0.5f );
                gradDIS[ 1 ]            = mixf( gradDIS[ 1 ], prev.y, 0.5f );
                gradDIS[ 2 ]            = mixf( gradDIS[ 2 ], prev.z, 0.5f );
            }
            storeRGBA8( disPing, gradPix, { gradDIS[ 0 ], gradDIS[ 1 ], gradDIS[ 2 ], 0.0f } ); // :179-180
            storeU8( gradientPrevPix, gradPix, ( uint8_t )curStrataPos );
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// GA: gradient a-trous, CmASVGFGradientAtrous.comp:42-105
// ---------------------------------------------------------------------------------------------------------------
void gradientAtrous( AsvgfOracle_T& o, uint32_t fi, const Ctx& c, uint32_t iteration )
{
    const bool even = ( iteration % 2 ) == 0; // :89-104
    const Img& src  = o.get( even ? ASVGF_IMAGE_DIS_PING_GRADIENT : ASVGF_IMAGE_DIS_PONG_GRADIENT, fi );
    Img&       dst  = o.get( even ? ASVGF_IMAGE_DIS_PONG_GRADIENT : ASVGF_IMAGE_DIS_PING_GRADIENT, fi );
    Img&       hist = o.get( ASVGF_IMAGE_DIS_GRADIENT_HISTORY, fi );
    const int  S = ASVGF_STRATA_SIZE, step = 1 << iteration;
    const float waveletKernel[ 2 ][ 2 ] = { { 1.0f, 0.5f }, { 0.5f, 0.25f } };

#pragma omp parallel for schedule( dynamic, 4 )
    for( int gy = 0; gy < dst.h; gy++ )
    {
        for( int gx = 0; gx < dst.w; gx++ )
        {
            const ivec2 gradPix = { gx, gy };
            if( gradPix.x * S >= c.W || gradPix.y * S >= c.H ) continue; // :82-85

            ivec3 a = getCheckerboardedRenderArea( c, { gradPix.x * S, gradPix.y * S } ); // :45
            const ivec3 gradArea = { a.x / S, a.y / S, a.z / S };

            float out[ 3 ]  = { 0, 0, 0 };
            float weightSum = 0.0f;
            for( int yy = -1; yy <= 1; yy++ )
            {
                for( int xx = -1; xx <= 1; xx++ )
                {
                    const ivec2 q  = { gradPix.x + xx * step, gradPix.y + yy * step };
                    const vec4  g  = fetchRGBA8( src, q );
                    const float w  = waveletKernel[ abs( yy ) ][ abs( xx ) ] * ( testPixInRenderArea( q, gradArea ) ? 1.0f : 0.0f );
                    out[ 0 ] += g.x * w; out[ 1 ] += g.y * w; out[ 2 ] += g.z * w;
                    weightSum += w;
                }
            }
            const vec4 r = { out[ 0 ] / weightSum, out[ 1 ] / weightSum, out[ 2 ] / weightSum, 0.0f }; // :74
            storeRGBA8( dst, gradPix, r );
            if( iteration == 0 ) storeRGBA8( hist, gradPix, r ); // :99-100
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// T: temporal accumulation, CmSVGFTemporalAccumulation.comp:58-385
// ---------------------------------------------------------------------------------------------------------------
float getParallax( const Ctx& c, vec3 pos, vec3 posPrev ) // :58-69
{
    vec3 cameraDelta = { c.u->cameraPositionPrev[ 0 ] - c.u->cameraPosition[ 0 ],
                         c.u->cameraPositionPrev[ 1 ] - c.u->cameraPosition[ 1 ],
                         c.u->cameraPositionPrev[ 2 ] - c.u->cameraPosition[ 2 ] };
    vec3  v        = normalize3( pos );
    vec3  vPrev    = normalize3( posPrev - cameraDelta );
    float cosa     = saturate( dot3( v, vPrev ) );
    float parallax = sqrtf( 1.0f - cosa * cosa ) / fmaxf( cosa, 1e-6f );
    parallax *= 60.0f;
    return parallax;
}
float getSpecMinAccumFrames( const Ctx& c ) // :70-74
{
    float fps = safePositiveRcp( c.u->timeDelta );
    return clampf( 0.2f * fps, 0.0f, 128.0f );
}
float getSpecAccumSpeed( const Ctx& c, float Amax, float roughness, float NoV, float parallax ) // :80-92
{
    float acos01sq            = 1.0f - NoV;
    float a                   = glslPow( saturate( acos01sq ), 0.66f );
    float b                   = 1.1f + roughness * roughness;
    float parallaxSensitivity = ( b + a ) / ( b - a );
    float powerScale          = 1.0f + parallax * parallaxSensitivity;
    float f                   = 1.0f - exp2f( -200.0f * roughness * roughness );
    f *= glslPow( saturate( roughness ), 0.5f * powerScale );
    float A = 256.0f * f;
    return clampf( A, getSpecMinAccumFrames( c ), Amax );
}
float getSpecularDominantFactor( float nv, float roughness ) // :93-99
{
    float a = 0.298475f * logf( 39.4115f - 39.0029f * roughness );
    float f = glslPow( clampf( 1.0f - nv, 0.0f, 1.0f ), 10.8649f ) * ( 1.0f - a ) + a;
    return clampf( f, 0.0f, 1.0f );
}

void temporal( AsvgfOracle_T& o, uint32_t fi, const Ctx& c )
{
    const Img& unfDirect   = o.get( ASVGF_IMAGE_UNFILTERED_DIRECT, fi );
    const Img& unfIndir    = o.get( ASVGF_IMAGE_UNFILTERED_INDIR, fi );
    const Img& unfSpec     = o.get( ASVGF_IMAGE_UNFILTERED_SPECULAR, fi );
    const Img& viewDir     = o.get( ASVGF_IMAGE_VIEW_DIRECTION, fi );
    const Img& motionImg   = o.get( ASVGF_IMAGE_MOTION, fi );
    const Img& depthImg    = o.get( ASVGF_IMAGE_DEPTH_WORLD, fi );
    const Img& normalImg   = o.get( ASVGF_IMAGE_NORMAL, fi );
    const Img& metRough    = o.get( ASVGF_IMAGE_METALLIC_ROUGHNESS, fi );
    const Img& surfPosImg  = o.get( ASVGF_IMAGE_SURFACE_POSITION, fi );
    const Img& disPing     = o.get( ASVGF_IMAGE_DIS_PING_GRADIENT, fi );
    const Img& depthPrevI  = o.get( ASVGF_IMAGE_DEPTH_WORLD_PREV, fi );
    const Img& normalPrevI = o.get( ASVGF_IMAGE_NORMAL_PREV, fi );
    const Img& diffHistory = o.get( ASVGF_IMAGE_DIFF_COLOR_HISTORY, fi );
    const Img& momentsPrevI= o.get( ASVGF_IMAGE_DIFF_ACCUM_MOMENTS_PREV, fi );
    const Img& histLenPrev = o.get( ASVGF_IMAGE_ACCUM_HISTORY_LENGTH_PREV, fi );
    const Img& specPrevI   = o.get( ASVGF_IMAGE_SPEC_ACCUM_COLOR_PREV, fi );
    const Img& indirPrevI  = o.get( ASVGF_IMAGE_INDIR_ACCUM_PREV, fi );
    const Img& surfPosPrevI= o.get( ASVGF_IMAGE_SURFACE_POSITION_PREV, fi );
    Img& outDiffTemporary  = o.get( ASVGF_IMAGE_DIFF_TEMPORARY, fi );
    Img& outMoments        = o.get( ASVGF_IMAGE_DIFF_ACCUM_MOMENTS, fi );
    Img& outHistLen        = o.get( ASVGF_IMAGE_ACCUM_HISTORY_LENGTH, fi );
    Img& outSpecAccum      = o.get( ASVGF_IMAGE_SPEC_ACCUM_COLOR, fi );
    Img& outIndirAccum     = o.get( ASVGF_IMAGE_INDIR_ACCUM, fi );
    Img& outMask           = o.get( ASVGF_IMAGE_DEBUG_MASK, fi );

#pragma omp parallel for schedule( dynamic, 4 )
    for( int py = 0; py < c.H; py++ )
    {
        for( int px = 0; px < c.W; px++ )
        {
            const ivec2 pix = { px, py }; // bounds check :105-108 is the loop range

            const vec3 unfilteredDiff    = fetchE5( unfDirect, pix ); // :110-114
            const vec3 unfilteredIndir   = fetchE5( unfIndir, pix );
            const vec3 unfilteredSpec    = fetchE5( unfSpec, pix );
            const vec4 unfilteredRayInfo = fetchRGBA16F( viewDir, pix );
            const vec3 rayDir            = { unfilteredRayInfo.x, unfilteredRayInfo.y, unfilteredRayInfo.z };

            const ivec3 chRenderArea = getCheckerboardedRenderArea( c, pix ); // :116-120
            const vec4  motion4      = fetchRGBA16F( motionImg, pix );
            const float motionZ      = motion4.z;
            const float depth        = fetchR16F( depthImg, pix );
            const vec3  normal       = decodeNormal( fetchU32( normalImg, pix ) );
            const float roughness    = fetchRG8( metRough, pix ).y;

            const float l       = getLuminance( unfilteredDiff ); // :122-123
            const vec2  moments = { l, l * l };

            ivec2 pixPrev; vec2 subPix; // :126-137
            {
                const vec2 posPrev = getPrevScreenPos( c, { motion4.x, motion4.y }, pix );
                const float fx = floorf( posPrev.x - 0.5f ), fy = floorf( posPrev.y - 0.5f );
                pixPrev = { f2i( fx ), f2i( fy ) };
                subPix  = { ( posPrev.x - 0.5f ) - fx, ( posPrev.y - 0.5f ) - fy };
            }
            const float bilinearWeights[ 2 ][ 2 ] = {
                { ( 1.0f - subPix.x ) * ( 1.0f - subPix.y ), subPix.x * ( 1.0f - subPix.y ) },
                { ( 1.0f - subPix.x ) * subPix.y, subPix.x * subPix.y },
            };

            vec4  historyLengthDIS = { 0, 0, 0, 0 }; // :141-146
            vec3  diffPrev = { 0, 0, 0 };
            vec2  momentsPrev = { 0, 0 };
            vec3  specPrevOnSurf = { 0, 0, 0 };
            vec3  indirPrev = { 0, 0, 0 };
            float weightSum = 0.0f;
            uint32_t mask = 0;

            for( int yy = 0; yy <= 1; yy++ ) // :149-178
            {
                for( int xx = 0; xx <= 1; xx++ )
                {
                    const ivec2 xy         = { pixPrev.x + xx, pixPrev.y + yy };
                    const float depthPrev  = fetchR16F( depthPrevI, xy );
                    const vec3  normalPrev = decodeNormal( fetchU32( normalPrevI, xy ) );
                    const bool isConsistent = testPixInRenderArea( xy, chRenderArea ) &&
                                              testReprojectedDepth( depth, depthPrev, motionZ ) &&
                                              testReprojectedNormal( normal, normalPrev );
                    if( isConsistent )
                    {
                        const float weight = bilinearWeights[ yy ][ xx ];
                        const vec4  dh     = fetchRGBA16F( diffHistory, xy );
                        diffPrev           = diffPrev + weight * vec3{ dh.x, dh.y, dh.z };
                        momentsPrev        = momentsPrev + fetchRG16F( momentsPrevI, xy ) * weight;
                        historyLengthDIS   = historyLengthDIS + fetchRGBA16F( histLenPrev, xy ) * weight;
                        specPrevOnSurf     = specPrevOnSurf + weight * fetchE5( specPrevI, xy );
                        indirPrev          = indirPrev + weight * fetchE5( indirPrevI, xy );
                        weightSum += weight;
                        mask |= 1u << ( yy * 2 + xx );
                    }
                }
            }

            float antilagAlphaDiff, antilagAlphaIndir, antilagAlphaSpec; // :181-194
            {
                const vec4 gradDIS = fetchRGBA8( disPing, { pix.x / ASVGF_STRATA_SIZE, pix.y / ASVGF_STRATA_SIZE } );
                antilagAlphaDiff   = gradDIS.x;
                antilagAlphaIndir  = gradDIS.y;
                antilagAlphaSpec   = gradDIS.z;
            }

            vec3  specColorAccum;
            float specHistoryLengthAccum;
            float specHitDistanceAccum;

            if( weightSum > 0.0001f ) // :202-309
            {
                mask |= 1u << 4;
                specPrevOnSurf = { specPrevOnSurf.x / weightSum, specPrevOnSurf.y / weightSum, specPrevOnSurf.z / weightSum };
                const float specHistoryLength   = historyLengthDIS.z / weightSum;
                const float specHitDistancePrev = historyLengthDIS.w / weightSum;

                const vec4 sp4 = fetchRGBA32F( surfPosImg, pix );
                const vec4 spp4 = fetchRGBA32F( surfPosPrevI, pixPrev );
                const vec3 surfPos = { sp4.x, sp4.y, sp4.z };
                const vec3 surfPosPrev = { spp4.x, spp4.y, spp4.z };
                float nv = fmaxf( 0.0f, dot3( normal, -rayDir ) );

                const float A = clampf( 1.0f + specHistoryLength * glslPow( 1.0f - antilagAlphaSpec, 5.0f ),
                                        getSpecMinAccumFrames( c ), 256.0f );

                const float parallax = getParallax( c, surfPos, surfPosPrev );
                float       AOnSurf  = getSpecAccumSpeed( c, A, roughness, nv, parallax );

                const vec3  specOnSurf    = mix3( specPrevOnSurf, unfilteredSpec, 1.0f / ( 1.0f + AOnSurf ) );
                const float hitDistOnSurf = mixf( specHitDistancePrev, unfilteredRayInfo.w, 1.0f / ( 1.0f + AOnSurf ) );

                specColorAccum         = specOnSurf;
                specHistoryLengthAccum = AOnSurf;
                specHitDistanceAccum   = hitDistOnSurf;

                // :228  vec * float * float evaluates left to right
                const float domFactor = getSpecularDominantFactor( nv, roughness );
                const vec3  virtualPos = surfPos + ( rayDir * hitDistOnSurf ) * domFactor;

                ivec2 pixPrevSpec; vec2 subPixSpec; // :232-249
                {
                    const vec4 vp              = { virtualPos.x, virtualPos.y, virtualPos.z, 1.0f };
                    const vec4 viewSpacePosCur = mulMat( c.u->view, vp );
                    const vec4 viewSpacePosPrev= mulMat( c.u->viewPrev, vp );
                    const vec4 clipCur         = mulMat( c.u->projection, viewSpacePosCur );
                    const vec4 clipPrev        = mulMat( c.u->projectionPrev, viewSpacePosPrev );
                    const vec2 ndcCur          = { clipCur.x / clipCur.w, clipCur.y / clipCur.w };
                    const vec2 ndcPrev         = { clipPrev.x / clipPrev.w, clipPrev.y / clipPrev.w };
                    const vec2 screenCur       = { ndcCur.x * 0.5f + 0.5f, ndcCur.y * 0.5f + 0.5f };
                    const vec2 screenPrev      = { ndcPrev.x * 0.5f + 0.5f, ndcPrev.y * 0.5f + 0.5f };
                    const vec2 specMotion      = { screenPrev.x - screenCur.x, screenPrev.y - screenCur.y };
                    const vec2 specPosPrev     = getPrevScreenPos( c, specMotion, pix );
                    const float fx = floorf( specPosPrev.x - 0.5f ), fy = floorf( specPosPrev.y - 0.5f );
                    pixPrevSpec = { f2i( fx ), f2i( fy ) };
                    subPixSpec  = { ( specPosPrev.x - 0.5f ) - fx, ( specPosPrev.y - 0.5f ) - fy };
                }
                const float bilinearWeightsSpec[ 2 ][ 2 ] = {
                    { ( 1.0f - subPixSpec.x ) * ( 1.0f - subPixSpec.y ), subPixSpec.x * ( 1.0f - subPixSpec.y ) },
                    { ( 1.0f - subPixSpec.x ) * subPixSpec.y, subPixSpec.x * subPixSpec.y },
                };

                float weightSumVirtual = 0.0f; // :256-280 (specPrevVirtual is accumulated by the shader but never used)
                for( int yy = 0; yy <= 1; yy++ )
                {
                    for( int xx = 0; xx <= 1; xx++ )
                    {
                        const ivec2 xy         = { pixPrevSpec.x + xx, pixPrevSpec.y + yy };
                        const float depthPrev  = fetchR16F( depthPrevI, xy );
                        const vec3  normalPrev = decodeNormal( fetchU32( normalPrevI, xy ) );
                        const bool isConsistent = testPixInRenderArea( xy, chRenderArea ) &&
                                                  testReprojectedDepth( depth, depthPrev, motionZ ) &&
                                                  testReprojectedNormal( normal, normalPrev );
                        if( isConsistent ) weightSumVirtual += bilinearWeightsSpec[ yy ][ xx ];
                    }
                }

                if( weightSumVirtual > 0.0001f ) // :282-307
                {
                    mask |= 1u << 5;
                    float Avirt = getSpecAccumSpeed( c, A, roughness, nv, 0.0f );
                    float confidence = 1.0f;
                    float Amin = fminf( Avirt, 3.0f * sqrtf( roughness ) );
                    float a    = mixf( 1.0f / ( 1.0f + Amin ), 1.0f / ( 1.0f + Avirt ), confidence );
                    Avirt      = 1.0f / a - 1.0f;
                    float Ahitdist = fminf( Avirt, 16.0f );

                    vec3 specVirt = mix3( specPrevOnSurf, unfilteredSpec, 1.0f / ( 1.0f + Avirt ) ); // :298 (sic)

                    float amount = getSpecularDominantFactor( nv, roughness );
                    a            = mixf( 1.0f / ( 1.0f + AOnSurf ), 1.0f / ( 1.0f + Avirt ), amount );
                    specColorAccum         = mix3( specOnSurf, specVirt, amount );
                    specHistoryLengthAccum = 1.0f / a - 1.0f;
                    specHitDistanceAccum   = mixf( specHitDistancePrev, unfilteredRayInfo.w, 1.0f / ( 1.0f + Ahitdist ) );
                }
            }
            else // :310-315
            {
                specColorAccum         = unfilteredSpec;
                specHistoryLengthAccum = 1.0f;
                specHitDistanceAccum   = unfilteredRayInfo.w;
            }

            vec3  diffColorAccum; vec2 diffMomentsAccum; float diffHistoryLengthAccum;
            vec3  indirAccum; float indirHistoryLengthAccum;

            if( weightSum > 0.0001f ) // :326-367
            {
                const float invWeightSum = 1.0f / weightSum;
                diffPrev    = diffPrev * invWeightSum;
                momentsPrev = momentsPrev * invWeightSum;
                indirPrev   = indirPrev * invWeightSum;
                float diffHistoryLength  = historyLengthDIS.x * invWeightSum;
                float indirHistoryLength = historyLengthDIS.y * invWeightSum;
                {
                    diffHistoryLength *= glslPow( 1.0f - antilagAlphaDiff, 10.0f );
                    diffHistoryLength = clampf( diffHistoryLength + 1.0f, 1.0f, 256.0f );
                    const float minAlpha = 0.01f;
                    float alphaColor   = fmaxf( minAlpha, 1.0f / diffHistoryLength );
                    float alphaMoments = fmaxf( minAlpha, 1.0f / diffHistoryLength );
                    alphaColor   = mixf( alphaColor, 1.0f, antilagAlphaDiff );
                    alphaMoments = mixf( alphaMoments, 1.0f, antilagAlphaDiff );
                    diffColorAccum         = mix3( diffPrev, unfilteredDiff, alphaColor );
                    diffMomentsAccum       = mix2( momentsPrev, moments, alphaMoments );
                    diffHistoryLengthAccum = diffHistoryLength;
                }
                {
                    indirHistoryLength *= glslPow( 1.0f - antilagAlphaIndir, 10.0f );
                    indirHistoryLength = clampf( indirHistoryLength + 1.0f, 1.0f, 256.0f );
                    const float minAlpha = 0.01f;
                    float alphaColor = fmaxf( minAlpha, 1.0f / indirHistoryLength );
                    alphaColor       = mixf( alphaColor, 1.0f, antilagAlphaIndir );
                    indirAccum              = mix3( indirPrev, unfilteredIndir, alphaColor );
                    indirHistoryLengthAccum = indirHistoryLength;
                }
            }
            else // :368-376
            {
                diffColorAccum = unfilteredDiff; diffMomentsAccum = moments; diffHistoryLengthAccum = 1.0f;
                indirAccum = unfilteredIndir; indirHistoryLengthAccum = 1.0f;
            }

            storeE5( outDiffTemporary, pix, diffColorAccum ); // :378-384
            storeRG16F( outMoments, pix, diffMomentsAccum );
            storeRGBA16F( outHistLen, pix, { diffHistoryLengthAccum, indirHistoryLengthAccum, specHistoryLengthAccum, specHitDistanceAccum } );
            storeE5( outSpecAccum, pix, specColorAccum );
            storeE5( outIndirAccum, pix, indirAccum );
            storeU8( outMask, pix, ( uint8_t )mask );
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// F: anti-firefly, CmAntiFirefly.comp:96-159  (the shared-memory tile :47-92 equals direct fetches with OOB -> 0)
// ---------------------------------------------------------------------------------------------------------------
void antiFirefly( AsvgfOracle_T& o, uint32_t fi, const Ctx& c )
{
    const Img* src[ 3 ] = { &o.get( ASVGF_IMAGE_DIFF_TEMPORARY, fi ), &o.get( ASVGF_IMAGE_SPEC_ACCUM_COLOR, fi ),
                            &o.get( ASVGF_IMAGE_INDIR_ACCUM, fi ) };
    Img*       dst[ 3 ] = { &o.get( ASVGF_IMAGE_DIFF_ACCUM_COLOR, fi ), &o.get( ASVGF_IMAGE_SPEC_PING_COLOR, fi ),
                            &o.get( ASVGF_IMAGE_INDIR_PING, fi ) };
    const float MAXFLOAT_ = 1000000.0f;

#pragma omp parallel for schedule( dynamic, 4 )
    for( int py = 0; py < c.H; py++ )
    {
        for( int px = 0; px < c.W; px++ )
        {
            const ivec2 pix = { px, py };
            for( int ch = 0; ch < 3; ch++ )
            {
                vec3 center = fetchE5( *src[ ch ], pix );
                if( c.u->antiFireflyEnabled == 0 ) // :109-117
                {
                    storeE5( *dst[ ch ], pix, center );
                    continue;
                }
                vec3 smin = { MAXFLOAT_, MAXFLOAT_, MAXFLOAT_ }, smax = { -MAXFLOAT_, -MAXFLOAT_, -MAXFLOAT_ };
                for( int i = -2; i <= 2; i++ ) // :130-149
                {
                    for( int j = -2; j <= 2; j++ )
                    {
                        if( i == 0 && j == 0 ) continue;
                        const vec3 other = fetchE5( *src[ ch ], { pix.x + i, pix.y + j } );
                        smin = { fminf( smin.x, other.x ), fminf( smin.y, other.y ), fminf( smin.z, other.z ) };
                        smax = { fmaxf( smax.x, other.x ), fmaxf( smax.y, other.y ), fmaxf( smax.z, other.z ) };
                    }
                }
                center = { clampf( center.x, smin.x, smax.x ), clampf( center.y, smin.y, smax.y ), clampf( center.z, smin.z, smax.z ) };
                storeE5( *dst[ ch ], pix, center );
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// V: variance estimation, CmSVGFEstimateVariance.comp:129-205
// ---------------------------------------------------------------------------------------------------------------
void variance( AsvgfOracle_T& o, uint32_t fi, const Ctx& c )
{
    const Img& colorImg   = o.get( ASVGF_IMAGE_DIFF_ACCUM_COLOR, fi );
    const Img& normalImg  = o.get( ASVGF_IMAGE_NORMAL, fi );
    const Img& depthImg   = o.get( ASVGF_IMAGE_DEPTH_WORLD, fi );
    const Img& histLenImg = o.get( ASVGF_IMAGE_ACCUM_HISTORY_LENGTH, fi );
    const Img& momentsImg = o.get( ASVGF_IMAGE_DIFF_ACCUM_MOMENTS, fi );
    Img&       out        = o.get( ASVGF_IMAGE_DIFF_PING_COLOR_AND_VARIANCE, fi );
    const float HISTORY_LENGTH_THRESHOLD = 4.0f;

#pragma omp parallel for schedule( dynamic, 4 )
    for( int py = 0; py < c.H; py++ ) // no bounds check in the shader; its out-of-image stores are discarded
    {
        for( int px = 0; px < c.W; px++ )
        {
            const ivec2 pix = { px, py };
            const float    pixDepth   = fetchR16F( depthImg, pix );
            const uint32_t pixEncNorm = fetchU32( normalImg, pix );
            vec3           pixDataColor = fetchE5( colorImg, pix );

            if( pixDepth < 0.0f || pixDepth > ASVGF_MAX_RAY_LENGTH ) // :143-147
            {
                storeRGBA16F( out, pix, { pixDataColor.x, pixDataColor.y, pixDataColor.z, 0.0f } );
                continue;
            }
            const float historyLength   = fetchRGBA16F( histLenImg, pix ).x; // :149-150
            const vec2  temporalMoments = fetchRG16F( momentsImg, pix );

            if( historyLength > HISTORY_LENGTH_THRESHOLD ) // :153-161
            {
                float temporalVariance = fmaxf( 0.0f, temporalMoments.y - temporalMoments.x * temporalMoments.x );
                temporalVariance /= fmaxf( 1.0f, historyLength - HISTORY_LENGTH_THRESHOLD );
                storeRGBA16F( out, pix, { pixDataColor.x, pixDataColor.y, pixDataColor.z, temporalVariance } );
                continue;
            }

            const float l = getLuminance( pixDataColor ); // :163-166
            vec2  spatialMoments = { l + temporalMoments.x, l * l + temporalMoments.y };
            float weightSum      = 1.0f;
            const vec3 n = decodeNormal( pixEncNorm );

            for( int yy = -1; yy <= 1; yy++ ) // :168-196
            {
                for( int xx = -1; xx <= 1; xx++ )
                {
                    if( xx == 0 && yy == 0 ) continue;
                    const ivec2 q      = { pix.x + xx, pix.y + yy };
                    const vec3  colorQ = fetchE5( colorImg, q );
                    const float lQ     = getLuminance( colorQ );
                    const float depthQ = fetchR16F( depthImg, q );
                    const vec3  nQ     = decodeNormal( fetchU32( normalImg, q ) );

                    const float wZ = expf( -fabsf( pixDepth - depthQ ) * safePositiveRcp( fmaxf( pixDepth, depthQ ) ) );
                    const float wN = glslPow( fmaxf( 0.0f, dot3( n, nQ ) ), 128.0f );
                    const float w  = wZ * wN;
                    if( !std::isnan( w ) )
                    {
                        spatialMoments = { spatialMoments.x + lQ * w, spatialMoments.y + ( lQ * lQ ) * w };
                        pixDataColor   = pixDataColor + colorQ * w;
                        weightSum += w;
                    }
                }
            }
            const float invWeightSum = safePositiveRcp( weightSum ); // :198-203
            spatialMoments = spatialMoments * invWeightSum;
            pixDataColor   = pixDataColor * invWeightSum;
            const float spatialVariance = fmaxf( 0.0f, spatialMoments.y - spatialMoments.x * spatialMoments.x ) *
                                          ( 1.0f + 2.0f * ( 1.0f - historyLength / HISTORY_LENGTH_THRESHOLD ) );
            storeRGBA16F( out, pix, { pixDataColor.x, pixDataColor.y, pixDataColor.z, spatialVariance } );
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// A: edge-stopping a-trous.  i==0: CmSVGFAtrous_Iter0.comp:158-335; i>=1: CmSVGFAtrous.comp:65-349
// ---------------------------------------------------------------------------------------------------------------
float roughnessSquaredToSpecPower( float alpha ) { return fmaxf( 0.01f, 2.0f / ( square( alpha ) + 1e-4f ) - 2.0f ); } // BRDF.h:27-29
vec3  getSpecularColor( vec3 albedo, float metallic ) { return mix3( { 0.04f, 0.04f, 0.04f }, albedo, metallic ); }   // BRDF.h:49-53
float getMaterialAmbient( vec3 albedo )                                                                               // BRDF.h:57-64
{
    float l = getLuminance( albedo );
    return l > 0.02f ? 1.0f : 1.0f - square( ( l - 0.02f ) / 0.02f );
}

void atrous( AsvgfOracle_T& o, uint32_t fi, const Ctx& c, uint32_t iteration )
{
    const bool last = iteration + 1 == o.iterations && iteration > 0;
    // routing: Iter0.comp:312-334; CmSVGFAtrous.comp:237-276; generalised beyond 4 iterations as ping-pong
    uint32_t sDiff, sSpec, sIndir, dDiff, dSpec, dIndir;
    if( iteration == 0 )
    {
        sDiff = ASVGF_IMAGE_DIFF_PING_COLOR_AND_VARIANCE; sSpec = ASVGF_IMAGE_SPEC_PING_COLOR; sIndir = ASVGF_IMAGE_INDIR_PING;
        dDiff = ASVGF_IMAGE_DIFF_COLOR_HISTORY; dSpec = ASVGF_IMAGE_SPEC_PONG_COLOR; dIndir = ASVGF_IMAGE_INDIR_PONG;
    }
    else if( iteration % 2 == 1 )
    {
        sDiff = iteration == 1 ? ASVGF_IMAGE_DIFF_COLOR_HISTORY : ASVGF_IMAGE_DIFF_PONG_COLOR_AND_VARIANCE;
        sSpec = ASVGF_IMAGE_SPEC_PONG_COLOR; sIndir = ASVGF_IMAGE_INDIR_PONG;
        dDiff = ASVGF_IMAGE_DIFF_PING_COLOR_AND_VARIANCE; dSpec = ASVGF_IMAGE_SPEC_PING_COLOR; dIndir = ASVGF_IMAGE_INDIR_PING;
    }
    else
    {
        sDiff = ASVGF_IMAGE_DIFF_PING_COLOR_AND_VARIANCE; sSpec = ASVGF_IMAGE_SPEC_PING_COLOR; sIndir = ASVGF_IMAGE_INDIR_PING;
        dDiff = ASVGF_IMAGE_DIFF_PONG_COLOR_AND_VARIANCE; dSpec = ASVGF_IMAGE_SPEC_PONG_COLOR; dIndir = ASVGF_IMAGE_INDIR_PONG;
    }
    const Img& samplerDiff = o.get( sDiff, fi );
    const Img& samplerSpec = o.get( sSpec, fi );
    const Img& samplerIndir= o.get( sIndir, fi );
    Img& outDiffImg  = o.get( dDiff, fi );
    Img& outSpecImg  = o.get( dSpec, fi );
    Img& outIndirImg = o.get( dIndir, fi );
    const Img& depthImg    = o.get( ASVGF_IMAGE_DEPTH_WORLD, fi );
    const Img& depthGrad   = o.get( ASVGF_IMAGE_DEPTH_GRAD, fi );
    const Img& normalImg   = o.get( ASVGF_IMAGE_NORMAL, fi );
    const Img& metRough    = o.get( ASVGF_IMAGE_METALLIC_ROUGHNESS, fi );
    const Img& histLenImg  = o.get( ASVGF_IMAGE_ACCUM_HISTORY_LENGTH, fi );
    Img&       filteredVar = o.get( ASVGF_IMAGE_ATROUS_FILTERED_VARIANCE, fi );
    const Img& albedoImg   = o.get( ASVGF_IMAGE_ALBEDO, fi );
    const Img& throughputI = o.get( ASVGF_IMAGE_THROUGHPUT, fi );
    const Img& isSkyImg    = o.get( ASVGF_IMAGE_IS_SKY, fi );
    Img&       preFinal    = o.get( ASVGF_IMAGE_PRE_FINAL, fi );
    Img&       histogramIn = o.get( ASVGF_IMAGE_HISTOGRAM_INPUT, fi );

    const int   STEP_SIZE       = 1 << iteration;
    const float SIGMA_LUMINANCE = 4.0f;
    const float waveletKernel[ 2 ][ 2 ] = { { 1.0f, 0.5f }, { 0.5f, 0.25f } };
    const float gaussianKernel[ 2 ][ 2 ] = { { 1.0f / 4.0f, 1.0f / 8.0f }, { 1.0f / 8.0f, 1.0f / 16.0f } };

#pragma omp parallel for schedule( dynamic, 4 )
    for( int py = 0; py < c.H; py++ )
    {
        for( int px = 0; px < c.W; px++ )
        {
            const ivec2 pix          = { px, py };
            const ivec3 chRenderArea = getCheckerboardedRenderArea( c, pix );

            const vec4 cv = fetchRGBA16F( samplerDiff, pix );
            vec3  outDiff     = { cv.x, cv.y, cv.z };
            float outVariance = cv.w;
            vec3  outSpec     = fetchE5( samplerSpec, pix );
            vec3  outIndir    = fetchE5( samplerIndir, pix );

            const float depth     = fetchR16F( depthImg, pix );
            const float gradDepth = fetchR16F( depthGrad, pix );
            const bool  invalid   = depth < 0.0f || depth > ASVGF_MAX_RAY_LENGTH;

            if( invalid && iteration == 0 ) // Iter0.comp:208-216: zeros, prefiltered variance NOT written
            {
                outVariance = 0.0f; outDiff = { 0, 0, 0 }; outSpec = { 0, 0, 0 }; outIndir = { 0, 0, 0 };
            }
            else if( !invalid ) // CmSVGFAtrous.comp:96-99: invalid pixels pass the centre through
            {
                const uint32_t encNormal = fetchU32( normalImg, pix );
                const vec3     normal    = decodeNormal( encNormal );
                const float    l         = getLuminance( outDiff );

                float sigmaL;
                if( iteration == 0 ) // Iter0.comp:158-187
                {
                    float r = 0.0f;
                    for( int yy = -1; yy <= 1; yy++ )
                        for( int xx = -1; xx <= 1; xx++ )
                        {
                            const float v = fetchRGBA16F( samplerDiff, { pix.x + xx, pix.y + yy } ).w;
                            const float w = gaussianKernel[ abs( xx ) ][ abs( yy ) ];
                            r += v * w;
                        }
                    r = sqrtf( fmaxf( r, 0.0f ) );
                    storeR16F( filteredVar, pix, r );
                    sigmaL = r;
                }
                else
                {
                    sigmaL = fetchR16F( filteredVar, pix ); // CmSVGFAtrous.comp:65-68
                }
                const float wLumMultiplier = 1.0f / ( SIGMA_LUMINANCE * sigmaL + 0.00001f );

                const float roughness        = fetchRG8( metRough, pix ).y;
                const float wRoughMultiplier = clampf( roughness * 30.0f - ( float )iteration, 0.0f, 1.0f );

                float historyLengthSpec = fetchRGBA16F( histLenImg, pix ).z;
                float normalWeightScale = clampf( historyLengthSpec / 8.0f, 0.0f, 1.0f );
                float normalWeightSpec  = roughnessSquaredToSpecPower( roughness * roughness );
                normalWeightSpec        = clampf( normalWeightSpec, 8.0f, 1024.0f );
                normalWeightSpec *= normalWeightScale;

                float weightSum = 1.0f, weightSumSpec = 1.0f, weightSumIndir = 1.0f;

                for( int yy = -1; yy <= 1; yy++ )
                {
                    for( int xx = -1; xx <= 1; xx++ )
                    {
                        if( xx == 0 && yy == 0 ) continue;
                        const ivec2 pixQ = { pix.x + xx * STEP_SIZE, pix.y + yy * STEP_SIZE };

                        const float roughnessQ = fetchRG8( metRough, pixQ ).y;
                        const vec3  normalQ    = decodeNormal( fetchU32( normalImg, pixQ ) );
                        const float depthQ     = fetchR16F( depthImg, pixQ );
                        const vec3  specColorQ = fetchE5( samplerSpec, pixQ );
                        const vec4  cvQ        = fetchRGBA16F( samplerDiff, pixQ );
                        const vec3  diffColorQ = { cvQ.x, cvQ.y, cvQ.z };
                        const float varianceQ  = cvQ.w;
                        const float lQ         = getLuminance( diffColorQ );
                        const float nN         = fmaxf( 0.0f, dot3( normal, normalQ ) );
                        const vec3  indirQ     = fetchE5( samplerIndir, pixQ );

                        const float wZ = fabsf( depth - depthQ ) / fmaxf( gradDepth * ( float )( abs( xx ) + abs( yy ) ), 0.01f );
                        const float wN = glslPow( nN, 128.0f );
                        const float wL = fabsf( l - lQ ) * wLumMultiplier;

                        float wR = fmaxf( 0.0f, 1.0f - 10.0f * fabsf( roughness - roughnessQ ) ) * wRoughMultiplier;
                        if( normalWeightSpec > 0.0f ) wR *= glslPow( nN, normalWeightSpec );

                        const float waveletW = waveletKernel[ abs( yy ) ][ abs( xx ) ];
                        const float isInside = testPixInRenderArea( pixQ, chRenderArea ) ? 1.0f : 0.0f;

                        const float wBase      = expf( -wZ * wZ ) * wN * waveletW * isInside;
                        const float wDiff      = wBase * expf( -wL );
                        const float wSpec      = wBase * wR;
                        const float wDiffIndir = wBase;

                        outDiff  = outDiff + diffColorQ * wDiff;
                        outSpec  = outSpec + specColorQ * wSpec;
                        outIndir = outIndir + indirQ * wDiffIndir;
                        outVariance += varianceQ * wDiff * wDiff;
                        weightSum += wDiff; weightSumSpec += wSpec; weightSumIndir += wDiffIndir;
                    }
                }
                const float invWeightSum = 1.0f / weightSum, invWeightSumSpec = 1.0f / weightSumSpec, invWeightSumIndir = 1.0f / weightSumIndir;
                outDiff = outDiff * invWeightSum;
                outVariance *= invWeightSum * invWeightSum;
                outSpec  = outSpec * invWeightSumSpec;
                outIndir = outIndir * invWeightSumIndir;
            }

            if( !last )
            {
                storeRGBA16F( outDiffImg, pix, { outDiff.x, outDiff.y, outDiff.z, outVariance } );
                storeE5( outSpecImg, pix, outSpec );
                storeE5( outIndirImg, pix, outIndir );
                continue;
            }

            // composition, CmSVGFAtrous.comp:278-348
            const ivec2 reg    = getRegularPixFromCheckerboardPix( c, pix );
            vec3        albedo = fetchUf11( albedoImg, reg );
            const vec4  tp4    = fetchRGBA16F( throughputI, pix );
            const vec3  throughput = { tp4.x, tp4.y, tp4.z };
            if( ( c.u->debugShowFlags & ASVGF_DEBUG_SHOW_FLAG_ALBEDO_WHITE ) != 0 ) albedo = { 1, 1, 1 };

            vec3 illuminated;
            if( fetchU8( isSkyImg, pix ) == 0 )
            {
                const vec2  mr        = fetchRG8( metRough, pix );
                const float metallic  = mr.x, roughness = mr.y;
                const vec3  diffuse = outDiff, indirect = outIndir;
                const vec3  specular = mix3( outSpec, indirect, smoothstepf( 0.5f, 0.5f + 0.25f, roughness ) );
                const vec3  roS = getSpecularColor( albedo, metallic );
                const vec3  roD = albedo * ( 1.0f - metallic );
                illuminated = ( diffuse + indirect ) * roD + specular * roS;
                if( ( c.u->debugShowFlags & ASVGF_DEBUG_SHOW_FLAG_ONLY_DIRECT_DIFFUSE ) != 0 ) illuminated = diffuse * albedo;
                else if( ( c.u->debugShowFlags & ASVGF_DEBUG_SHOW_FLAG_ONLY_SPECULAR ) != 0 ) illuminated = specular * getSpecularColor( albedo, metallic );
                else if( ( c.u->debugShowFlags & ASVGF_DEBUG_SHOW_FLAG_ONLY_INDIRECT_DIFFUSE ) != 0 ) illuminated = indirect;
                illuminated = illuminated * throughput;
                illuminated = illuminated * getMaterialAmbient( albedo );
            }
            else
            {
                illuminated = albedo * throughput;
            }
            storeR16F( histogramIn, reg, fmaxf( 0.0f, getLuminance( illuminated ) ) );
            if( classicShading( c, reg ) ) illuminated = albedo * throughput;
            storeU32( preFinal, reg, packUfloat111110( { fmaxf( 0.0f, illuminated.x ), fmaxf( 0.0f, illuminated.y ), fmaxf( 0.0f, illuminated.z ) } ) );
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Volumetric scattering accumulation, CmScatterAccum.comp:96-163 (Volumetric::ProcessScattering, Volumetric.cpp:71-195):
// the frame's other temporal-reprojection pass, same disocclusion tests as the temporal accumulation.
// fializeScattering (:40-94): VOLUME_ENABLE_SIMPLE and NONE are restated; VOLUME_ENABLE_VOLUMETRIC samples a 3-D
// illumination volume with a blue-noise dither (not an image-space operation) and is read from ASVGF_IMAGE_SCATTER_CURRENT.
// Pin: exp() = correctly rounded fp32.
// ---------------------------------------------------------------------------------------------------------------
void scatterAccumulate( AsvgfOracle_T& o, uint32_t fi, const Ctx& c, float maxHistoryLength )
{
    Img& motion = o.get( ASVGF_IMAGE_MOTION, fi );
    Img& depthW = o.get( ASVGF_IMAGE_DEPTH_WORLD, fi );
    Img& normalI = o.get( ASVGF_IMAGE_NORMAL, fi );
    Img& depthPrevI = o.get( ASVGF_IMAGE_DEPTH_WORLD_PREV, fi );
    Img& normalPrevI = o.get( ASVGF_IMAGE_NORMAL_PREV, fi );
    Img& scatPrev = o.get( ASVGF_IMAGE_SCATTERING_PREV, fi );
    Img& histPrev = o.get( ASVGF_IMAGE_SCATTERING_HISTORY_PREV, fi );
    Img& scat = o.get( ASVGF_IMAGE_SCATTERING, fi );
    Img& hist = o.get( ASVGF_IMAGE_SCATTERING_HISTORY, fi );
    Img& current = o.get( ASVGF_IMAGE_SCATTER_CURRENT, fi );
#pragma omp parallel for schedule( static )
    for( int y = 0; y < c.H; y++ )
    {
        for( int x = 0; x < c.W; x++ )
        {
            const ivec2 pix = { x, y }; // a checkerboard pixel (:98-103)
            const vec4  mv      = fetchRGBA16F( motion, pix );
            const float motionZ = mv.z;
            const float depth   = fetchR16F( depthW, pix );
            const vec3  normal  = decodeNormal( fetchU32( normalI, pix ) );
            const vec2  posPrev = getPrevScreenPos( c, { mv.x, mv.y }, pix ); // :112
            const vec2  q       = { posPrev.x - 0.5f, posPrev.y - 0.5f };
            const vec2  fl      = { floorf( q.x ), floorf( q.y ) };
            const ivec2 pixPrev = { f2i( fl.x ), f2i( fl.y ) };
            const vec2  subPix  = { q.x - fl.x, q.y - fl.y };
            const float bw[ 2 ][ 2 ] = { { ( 1.0f - subPix.x ) * ( 1.0f - subPix.y ), subPix.x * ( 1.0f - subPix.y ) },
                                         { ( 1.0f - subPix.x ) * subPix.y, subPix.x * subPix.y } };
            vec4  accum      = { 0, 0, 0, 0 };
            float historyLen = 0, weightSum = 0;
            const ivec3 area = getCheckerboardedRenderArea( c, pix );
            for( int yy = 0; yy <= 1; yy++ )
                for( int xx = 0; xx <= 1; xx++ ) // :127-151
                {
                    const ivec2 xy         = { pixPrev.x + xx, pixPrev.y + yy };
                    const float depthPrev  = fetchR16F( depthPrevI, xy );
                    const vec3  normalPrev = decodeNormal( fetchU32( normalPrevI, xy ) );
                    if( testPixInRenderArea( xy, area ) && testReprojectedDepth( depth, depthPrev, motionZ ) && testReprojectedNormal( normal, normalPrev ) )
                    {
                        const float w   = bw[ yy ][ xx ];
                        const ivec2 reg = getRegularPixFromCheckerboardPix( c, xy );
                        const vec4  sp  = fetchRGBA16F( scatPrev, reg );
                        accum      = { accum.x + w * sp.x, accum.y + w * sp.y, accum.z + w * sp.z, accum.w + w * sp.w };
                        historyLen = historyLen + w * fetchR16F( histPrev, reg );
                        weightSum  = weightSum + w;
                    }
                }
            historyLen = weightSum > 0.0001f ? historyLen / weightSum : 0.0f; // :154-155
            accum      = weightSum > 0.0001f ? vec4{ accum.x / weightSum, accum.y / weightSum, accum.z / weightSum, accum.w / weightSum } : vec4{ 0, 0, 0, 0 };
            historyLen = fminf( fmaxf( historyLen + 1.0f, 1.0f ), fmaxf( 1.0f, maxHistoryLength ) ); // :157

            const ivec2 reg = getRegularPixFromCheckerboardPix( c, pix );
            vec4        cur; // fializeScattering, :40-94
            if( c.u->volumeEnableType == ASVGF_VOLUME_ENABLE_VOLUMETRIC ) cur = fetchRGBA16F( current, reg );
            else if( c.u->volumeEnableType == ASVGF_VOLUME_ENABLE_SIMPLE )
            {
                const float virtualdepth = fetchR16F( depthW, getCheckerboardPix( c, reg ) );
                const float density      = c.u->volumeScattering * 0.00005f;
                if( !( virtualdepth > ASVGF_MAX_RAY_LENGTH ) )
                {
                    const float f = ( float )std::exp( ( double )( -virtualdepth * density ) );
                    cur           = { c.u->volumeAmbient[ 0 ] * ( 1.0f - f ), c.u->volumeAmbient[ 1 ] * ( 1.0f - f ), c.u->volumeAmbient[ 2 ] * ( 1.0f - f ), f };
                }
                else cur = { 0, 0, 0, 1.0f };
            }
            else cur = { 0, 0, 0, 1.0f };
            const float t = 1.0f / historyLen; // :158
            accum         = { mixf( accum.x, cur.x, t ), mixf( accum.y, cur.y, t ), mixf( accum.z, cur.z, t ), mixf( accum.w, cur.w, t ) };
            storeRGBA16F( scat, reg, accum );
            storeR16F( hist, reg, historyLen );
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Checkerboard resolve of the denoiser's output, CmCheckerboard.comp:43-78 (PreFinal -> Final; the shader's two other
// image pairs, ScreenEmisRT and AcidFogRT, come from other producers and are outside this path)
// ---------------------------------------------------------------------------------------------------------------
void resolveCheckerboard( AsvgfOracle_T& o, uint32_t fi, const Ctx& c )
{
    // the shader's three image pairs (CmCheckerboard.comp:62-77)
    const uint32_t srcs[ 3 ] = { ASVGF_IMAGE_PRE_FINAL, ASVGF_IMAGE_SCREEN_EMIS_RT, ASVGF_IMAGE_ACID_FOG_RT };
    const uint32_t dsts[ 3 ] = { ASVGF_IMAGE_FINAL, ASVGF_IMAGE_SCREEN_EMISSION, ASVGF_IMAGE_ACID_FOG };
    Img& throughput = o.get( ASVGF_IMAGE_THROUGHPUT, fi );
    const int sep   = getCheckerboardSeparatorX( c );
    for( int i = 0; i < 3; i++ )
    {
        Img& src = o.get( srcs[ i ], fi );
        Img& dst = o.get( dsts[ i ], fi );
#pragma omp parallel for schedule( static )
        for( int y = 0; y < c.H; y++ )
        {
            for( int x = 0; x < c.W; x++ )
            {
                const ivec2 pix = { x, y };
                const ivec2 cb  = getCheckerboardPix( c, pix );               // :55
                vec3        hdr = unpackUfloat111110( fetchU32( src, pix ) ); // :62-64
                if( fetchRGBA16F( throughput, cb ).w > 0.5f )                 // needResolveCheckerboard, ShaderCommonGLSLFunc.h:434-444
                {
                    const float fl = ( cb.x == 0 || cb.x == sep ) ? 0.0f : 1.0f;               // !isLeftOutOfBounds  :29-34
                    const float fr = ( cb.x == sep || cb.x == ( int )c.u->renderWidth ) ? 0.0f : 1.0f; // !isRightOutOfBounds :36-41
                    const vec3 l = unpackUfloat111110( fetchU32( src, { x - 1, y } ) ), r = unpackUfloat111110( fetchU32( src, { x + 1, y } ) );
                    const vec3 d = unpackUfloat111110( fetchU32( src, { x, y + 1 } ) ), t = unpackUfloat111110( fetchU32( src, { x, y - 1 } ) );
                    auto cross = [ & ]( float vl, float vr, float vd, float vt ) { return 0.25f * vl * fl + 0.25f * vr * fr + 0.25f * vd + 0.25f * vt; }; // :45-49
                    const vec3 k = { cross( l.x, r.x, d.x, t.x ), cross( l.y, r.y, d.y, t.y ), cross( l.z, r.z, d.z, t.z ) };
                    hdr          = mix3( hdr, k, 0.5f );                      // :51
                }
                storeU32( dst, pix, packUfloat111110( hdr ) );               // :73-77
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Final prepare, CmPrepareFinal.comp:99-226 (ImageComposition::ApplyTonemapping, Source/ImageComposition.cpp:110-155): Final and
// the in-scattering are composed, exposed with the eye-adapted EV100 (Exposure.h:36-58) and tonemapped -- AMD's LPM for the
// ray-traced columns (LpmMap of LPM/ffx_lpm.h:727-831 in the configuration LpmFilter gets at CmPrepareFinal.comp:71:
// shoulder, con, soft, con2, clip, scaleOnly all false), Uncharted 2 for the classic-shading columns -- and the bloom input is
// written.  lpm = the 24 x uvec4 control block ImageComposition::SetupLpmParams computes on the host (ImageComposition.cpp:365-420).
// Pin: log2 / exp2 are the correctly rounded fp32 values, as in the exposure passes; pow() is libm's (glslPow).
// ---------------------------------------------------------------------------------------------------------------
struct LpmConstants // the members of the control block this configuration reads (ffx_lpm.h:913-938)
{
    vec3  saturation; float contrast; // map0
    float toneScaleBiasX, toneScaleBiasY; vec3 lumaT; // map1.rg, map1.ba + map2.r
    vec3  crosstalk; // map2.gba
    vec3  rcpLumaT;  // map3.rgb
};
LpmConstants lpmConstants( const uint32_t* lpm )
{
    auto f = [ & ]( int vec, int comp ) { float v; std::memcpy( &v, &lpm[ vec * 4 + comp ], 4 ); return v; };
    LpmConstants k;
    k.saturation = { f( 0, 0 ), f( 0, 1 ), f( 0, 2 ) }; k.contrast = f( 0, 3 );
    k.toneScaleBiasX = f( 1, 0 ); k.toneScaleBiasY = f( 1, 1 ); k.lumaT = { f( 1, 2 ), f( 1, 3 ), f( 2, 0 ) };
    k.crosstalk = { f( 2, 1 ), f( 2, 2 ), f( 2, 3 ) };
    k.rcpLumaT  = { f( 3, 0 ), f( 3, 1 ), f( 3, 2 ) };
    return k;
}
vec3 lpmMap( vec3 c, const LpmConstants& k ) // ffx_lpm.h:747-819 with every optional stage compiled out
{
    float colorR = c.x, colorG = c.y, colorB = c.z;
    float rcpMax = 1.0f / fmaxf( colorR, fmaxf( colorG, colorB ) );
    float ratioR = colorR * rcpMax, ratioG = colorG * rcpMax, ratioB = colorB * rcpMax;
    ratioR = glslPow( ratioR, k.saturation.x ); ratioG = glslPow( ratioG, k.saturation.y ); ratioB = glslPow( ratioB, k.saturation.z );
    float luma = colorG * k.lumaT.y + ( colorR * k.lumaT.x + ( colorB * k.lumaT.z ) );
    luma       = glslPow( luma, k.contrast );
    const float lumaShoulder = luma;
    luma = luma * ( 1.0f / ( lumaShoulder * k.toneScaleBiasX + k.toneScaleBiasY ) );
    const float lumaRatio  = ( ratioR * k.lumaT.x + ratioG * k.lumaT.y ) + ratioB * k.lumaT.z;
    const float ratioScale = saturate( luma * ( 1.0f / lumaRatio ) );
    colorR = saturate( ratioR * ratioScale ); colorG = saturate( ratioG * ratioScale ); colorB = saturate( ratioB * ratioScale );
    const float capR = ( -k.crosstalk.x ) * colorR + k.crosstalk.x;
    const float capG = ( -k.crosstalk.y ) * colorG + k.crosstalk.y;
    const float capB = ( -k.crosstalk.z ) * colorB + k.crosstalk.z;
    float lumaAdd = saturate( ( -colorB ) * k.lumaT.z + ( ( -colorR ) * k.lumaT.x + ( ( -colorG ) * k.lumaT.y + luma ) ) );
    const float t = lumaAdd * ( 1.0f / ( capG * k.lumaT.y + ( capR * k.lumaT.x + ( capB * k.lumaT.z ) ) ) );
    colorR = saturate( t * capR + colorR ); colorG = saturate( t * capG + colorG ); colorB = saturate( t * capB + colorB );
    lumaAdd = saturate( ( -colorB ) * k.lumaT.z + ( ( -colorR ) * k.lumaT.x + ( ( -colorG ) * k.lumaT.y + luma ) ) );
    colorR = saturate( lumaAdd * k.rcpLumaT.x + colorR );
    colorG = saturate( lumaAdd * k.rcpLumaT.y + colorG );
    colorB = saturate( lumaAdd * k.rcpLumaT.z + colorB );
    return { colorR, colorG, colorB };
}
float uncharted2Op( float x ) // CmPrepareFinal.comp:75-85
{
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ( ( x * ( A * x + C * B ) + D * E ) / ( x * ( A * x + B ) + D * F ) ) - E / F;
}
void prepareFinal( AsvgfOracle_T& o, uint32_t fi, const Ctx& c, const uint32_t* lpm )
{
    Img& final = o.get( ASVGF_IMAGE_FINAL, fi ); Img& emis = o.get( ASVGF_IMAGE_SCREEN_EMISSION, fi ); Img& scat = o.get( ASVGF_IMAGE_SCATTERING, fi );
    Img& fog = o.get( ASVGF_IMAGE_ACID_FOG, fi ); Img& bloomOut = o.get( ASVGF_IMAGE_BLOOM_INPUT, fi );
    Img& motion = o.get( ASVGF_IMAGE_MOTION, fi ); Img& disPing = o.get( ASVGF_IMAGE_DIS_PING_GRADIENT, fi ); Img& normal = o.get( ASVGF_IMAGE_NORMAL, fi );
    const uint32_t unf[ 3 ] = { ASVGF_IMAGE_UNFILTERED_DIRECT, ASVGF_IMAGE_UNFILTERED_SPECULAR, ASVGF_IMAGE_UNFILTERED_INDIR };
    const LpmConstants k = lpmConstants( lpm );
    // Exposure.h:36-58 (S = 100, K = 12.5)
    const float ev100 = ( float )std::log2( ( double )( fmaxf( 0.0f, o.tm.avgLuminance ) * 100.0f / 12.5f ) );
    auto luminousExposure = [ & ]( float ev ) { const float maxLuminance = 1.2f * ( float )std::exp2( ( double )ev ); return maxLuminance > 0.0f ? 1.0f / maxLuminance : 0.0f; };
    const float exposure    = luminousExposure( ev100 );
    const float bloomLum    = ( float )std::exp2( ( double )( ( ev100 + c.u->bloomEmissionMultiplier ) - 3.0f ) ); // ev100ToLuminance(evBloom)
    const float whiteScale  = uncharted2Op( c.u->luminanceWhitePoint );
    const uint32_t flags    = c.u->debugShowFlags;
#pragma omp parallel for schedule( static )
    for( int y = 0; y < c.H; y++ )
    {
        for( int x = 0; x < c.W; x++ )
        {
            const ivec2 pix = { x, y };
            vec3        hdr        = fetchUf11( final, pix );       // :152
            const vec3  screenEmis = fetchUf11( emis, pix );
            {
                const vec4 volumetric = fetchRGBA16F( scat, pix );  // :155-158
                hdr = hdr * volumetric.w + vec3{ volumetric.x, volumetric.y, volumetric.z };
            }
            vec3 color; // finalizeColor :107-119
            if( ( float )x < c.u->lightmapScreenCoverage * c.u->renderWidth ) // classicShading, ShaderCommonGLSLFunc.h:492-495
            {
                const vec3 h2 = 2.0f * hdr;
                color = { uncharted2Op( h2.x ) / whiteScale, uncharted2Op( h2.y ) / whiteScale, uncharted2Op( h2.z ) / whiteScale };
            }
            else color = lpmMap( hdr * exposure + screenEmis * c.u->emissionMaxScreenColor, k );
            vec3 bloom = ( hdr + screenEmis * bloomLum ) * exposure; // getBloomInput :123-140
            if( std::isnan( bloom.x ) || std::isnan( bloom.y ) || std::isnan( bloom.z ) || std::isinf( bloom.x ) || std::isinf( bloom.y ) || std::isinf( bloom.z ) )
                bloom = { 0.0f, 0.0f, 0.0f };
            bloom = bloom * smoothstepf( 0.0f, c.u->bloomThreshold, getLuminance( bloom ) );
            color = color + fetchUf11( fog, pix ); // SHIPPING_HACK (ShaderCommonGLSLFunc.h:49), :163-165
            if( flags != 0 ) // processDebug :175-226
            {
                const ivec2 cb = getCheckerboardPix( c, pix );
                if( flags & 1u ) { const vec4 m = fetchRGBA16F( motion, cb ); color = { fabsf( m.x ), fabsf( m.y ), 0.0f }; }
                else if( flags & ( 1u << 2 ) ) color = fetchE5( o.get( unf[ 0 ], fi ), cb );
                else if( flags & ( 1u << 3 ) ) color = fetchE5( o.get( unf[ 1 ], fi ), cb );
                else if( flags & ( 1u << 4 ) ) color = fetchE5( o.get( unf[ 2 ], fi ), cb );
                else if( flags & ( 1u << 8 ) ) { /* DEBUG_SHOW_FLAG_LIGHT_GRID needs the light grid of the ray tracer: not on this path (fallback kept) */ }
                else if( flags & ( 1u << 1 ) ) { const vec4 g = fetchRGBA8( disPing, { cb.x / ASVGF_STRATA_SIZE, cb.y / ASVGF_STRATA_SIZE } ); color = { g.x, g.y, g.z }; }
                else if( flags & ( 1u << 10 ) ) { const vec3 n = decodeNormal( fetchU32( normal, cb ) ); color = n * 0.5f + vec3{ 0.5f, 0.5f, 0.5f }; }
            }
            storeU32( final, pix, packUfloat111110( color ) );    // :170-171
            storeU32( bloomOut, pix, packUfloat111110( bloom ) );
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Exposure: CmLuminanceHistogram.comp:38-86 and CmLuminanceAvg.comp:37-89 (Tonemapping.cpp:59-175).
// Pin: log2 / exp2 / exp, whose GLSL precision is implementation-defined, are the correctly rounded fp32 values
// (evaluated in double, rounded once) -- the same convention as the sin/cos of the normal codec.
// ---------------------------------------------------------------------------------------------------------------
uint32_t getColorBin( float inputLum, float minLogLuminance, float maxLogLuminance ) // CmLuminanceHistogram.comp:38-55
{
    const float inverseLogLuminanceRange = 1.0f / ( maxLogLuminance - minLogLuminance );
    if( inputLum < 0.001f ) return 0; // LUMINANCE_EPS
    const float lg           = ( float )std::log2( ( double )inputLum );
    const float logLuminance = fminf( fmaxf( ( lg - minLogLuminance ) * inverseLogLuminanceRange, 0.0f ), 1.0f ); // clamp; NaN -> 0 (pinned)
    return ( uint32_t )f2i( logLuminance * 254.0f + 1.0f );
}
void luminanceHistogram( AsvgfOracle_T& o, uint32_t fi, const Ctx& c )
{
    Img& in = o.get( ASVGF_IMAGE_HISTOGRAM_INPUT, fi );
    for( int y = 0; y < c.H; y++ )
        for( int x = 0; x < c.W; x++ ) o.tm.histogram[ getColorBin( fetchR16F( in, { x, y } ), c.u->minLogLuminance, c.u->maxLogLuminance ) ]++;
}
void luminanceAverage( AsvgfOracle_T& o, const Ctx& c )
{
    uint32_t       finalWeightedCount = 0;
    const uint32_t countInBin0        = o.tm.histogram[ 0 ];
    for( uint32_t i = 0; i < ASVGF_LUM_HISTOGRAM_BIN_COUNT; i++ )
    {
        finalWeightedCount += o.tm.histogram[ i ] * i; // :43-59 (uint arithmetic; the tree reduction is order-independent)
        o.tm.histogram[ i ] = 0;                       // :48, for the next frame
    }
    if( c.u->stopEyeAdaptation != 0 ) return;          // :64-67
    const float minLogLuminance = c.u->minLogLuminance, maxLogLuminance = c.u->maxLogLuminance;
    const float logLuminanceRange = maxLogLuminance - minLogLuminance;
    const float pixelCount        = c.u->renderWidth * c.u->renderHeight;
    const float blackPixelCount   = ( float )countInBin0;
    const float weightedLogAverage   = ( ( float )finalWeightedCount / fmaxf( pixelCount - blackPixelCount, 1.0f ) ) - 1.0f;
    const float weightedAvgLuminance = ( float )std::exp2( ( double )( weightedLogAverage / 254.0f * logLuminanceRange + minLogLuminance ) );
    const float lastFrameLuminance   = o.tm.avgLuminance;
    const float tau = 1.1f;
    const float t   = 1.0f - ( float )std::exp( ( double )( -c.u->timeDelta * tau ) );
    const float adaptedLuminance = lastFrameLuminance + ( weightedAvgLuminance - lastFrameLuminance ) * t;
    o.tm.avgLuminance = fminf( adaptedLuminance, ( float )std::exp2( ( double )maxLogLuminance ) );
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------
// C API
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

AsvgfOracle asvgfOracleCreate( uint32_t W, uint32_t H, uint32_t iterations )
{
    if( W == 0 || H == 0 || ( W & 1u ) || iterations < 2 || iterations > ASVGF_MAX_ATROUS_ITERATIONS ) return nullptr;
    AsvgfOracle_T* o = new AsvgfOracle_T();
    o->W = W; o->H = H; o->iterations = iterations;
    for( uint32_t i = 0; i < ASVGF_IMAGE_COUNT; i++ )
    {
        int bpp = bytesPerPixelOf( i );
        if( bpp == 0 ) continue;
        if( isThirdRes( i ) ) o->img[ i ].alloc( std::max( 1u, ( W + 1 ) / 3 ), std::max( 1u, ( H + 1 ) / 3 ), bpp ); // Framebuffers.cpp:691-699
        else o->img[ i ].alloc( ( int )W, ( int )H, bpp );
    }
    return o;
}
void asvgfOracleDestroy( AsvgfOracle o ) { delete o; }

void asvgfOracleSetThreads( int threads )
{
#ifdef _OPENMP
    omp_set_num_threads( threads > 0 ? threads : omp_get_num_procs() );
#else
    ( void )threads;
#endif
}
int asvgfOracleGetThreads( void )
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int asvgfOracleGetImage( AsvgfOracle o, uint32_t index, uint32_t frameIndex, void** pp, uint32_t* w, uint32_t* h, uint32_t* bpp )
{
    if( !o || index >= ASVGF_IMAGE_COUNT || frameIndex > 1 || bytesPerPixelOf( index ) == 0 ) return 1;
    Img& im = o->get( index, frameIndex );
    if( pp ) *pp = im.d.data();
    if( w ) *w = ( uint32_t )im.w;
    if( h ) *h = ( uint32_t )im.h;
    if( bpp ) *bpp = ( uint32_t )im.bpp;
    return 0;
}
void asvgfOracleResetHistory( AsvgfOracle o )
{
    for( auto& im : o->img ) std::fill( im.d.begin(), im.d.end(), 0 );
    o->tm = AsvgfTonemapping{};
}

void asvgfOracleRunPass( AsvgfOracle o, uint32_t pass, uint32_t frameIndex, const AsvgfGlobalUniform* u )
{
    Ctx c = { u, ( int )u->renderWidth, ( int )u->renderHeight };
    if( pass == 0 ) gradientSamples( *o, frameIndex, c );
    else if( pass >= 1 && pass <= 4 ) gradientAtrous( *o, frameIndex, c, pass - 1 );
    else if( pass == 5 ) temporal( *o, frameIndex, c );
    else if( pass == 6 ) antiFirefly( *o, frameIndex, c );
    else if( pass == 7 ) variance( *o, frameIndex, c );
    else if( pass >= 8 && pass < 8 + o->iterations ) atrous( *o, frameIndex, c, pass - 8 );
}
void asvgfOracleGradientSamples( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u )
{
    asvgfOracleRunPass( o, 0, frameIndex, u );
}
void asvgfOracleDenoise( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u ) // Denoiser.cpp:87-254
{
    for( uint32_t p = 1; p < 8 + o->iterations; p++ ) asvgfOracleRunPass( o, p, frameIndex, u );
}

void asvgfOracleScatterAccumulate( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u, float maxHistoryLength )
{
    Ctx c = { u, ( int )u->renderWidth, ( int )u->renderHeight };
    scatterAccumulate( *o, frameIndex, c, maxHistoryLength );
}
void asvgfOracleResolveCheckerboard( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u ) // ImageComposition.cpp:157-222
{
    Ctx c = { u, ( int )u->renderWidth, ( int )u->renderHeight };
    resolveCheckerboard( *o, frameIndex, c );
}
void asvgfOraclePrepareFinal( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u, const uint32_t* lpmControl96 ) // ImageComposition.cpp:110-155
{
    Ctx c = { u, ( int )u->renderWidth, ( int )u->renderHeight };
    prepareFinal( *o, frameIndex, c, lpmControl96 );
}
void asvgfOracleLuminanceHistogram( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u )
{
    Ctx c = { u, ( int )u->renderWidth, ( int )u->renderHeight };
    luminanceHistogram( *o, frameIndex, c );
}
void asvgfOracleLuminanceAverage( AsvgfOracle o, const AsvgfGlobalUniform* u )
{
    Ctx c = { u, ( int )u->renderWidth, ( int )u->renderHeight };
    luminanceAverage( *o, c );
}
void asvgfOracleCalculateExposure( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u ) // Tonemapping.cpp:59-175
{
    asvgfOracleLuminanceHistogram( o, frameIndex, u );
    asvgfOracleLuminanceAverage( o, u );
}
AsvgfTonemapping* asvgfOracleGetTonemapping( AsvgfOracle o ) { return &o->tm; }

uint32_t asvgfOracleEncodeE5B9G9R9( float r, float g, float b ) { return encodeE5B9G9R9( { r, g, b } ); }
void     asvgfOracleDecodeE5B9G9R9( uint32_t p, float* rgb ) { vec3 v = decodeE5B9G9R9( p ); rgb[ 0 ] = v.x; rgb[ 1 ] = v.y; rgb[ 2 ] = v.z; }
uint32_t asvgfOracleEncodeNormal( float x, float y, float z ) { return encodeNormal( { x, y, z } ); }
void     asvgfOracleDecodeNormal( uint32_t p, float* xyz ) { vec3 v = decodeNormal( p ); xyz[ 0 ] = v.x; xyz[ 1 ] = v.y; xyz[ 2 ] = v.z; }
uint16_t asvgfOracleFloatToHalf( float f ) { return floatToHalf( f ); }
float    asvgfOracleHalfToFloat( uint16_t h ) { return halfToFloat( h ); }
uint32_t asvgfOraclePackUfloat111110( float r, float g, float b ) { return packUfloat111110( { r, g, b } ); }
void     asvgfOracleUnpackUfloat111110( uint32_t p, float* rgb ) { vec3 v = unpackUfloat111110( p ); rgb[ 0 ] = v.x; rgb[ 1 ] = v.y; rgb[ 2 ] = v.z; }
uint8_t  asvgfOracleFloatToUnorm8( float f ) { return floatToUnorm8( f ); }
void     asvgfOracleCheckerboardPix( int x, int y, int fullWidth, int* out )
{
    AsvgfGlobalUniform u{}; Ctx c = { &u, fullWidth, 0 };
    ivec2 r = getCheckerboardPix( c, { x, y } ); out[ 0 ] = r.x; out[ 1 ] = r.y;
}
void asvgfOracleRegularPix( int x, int y, int fullWidth, int* out )
{
    AsvgfGlobalUniform u{}; Ctx c = { &u, fullWidth, 0 };
    ivec2 r = getRegularPixFromCheckerboardPix( c, { x, y } ); out[ 0 ] = r.x; out[ 1 ] = r.y;
}
void asvgfOraclePrevScreenPos( float mx, float my, int x, int y, float renderWidth, float renderHeight, float* out )
{
    AsvgfGlobalUniform u{}; u.renderWidth = renderWidth; u.renderHeight = renderHeight;
    Ctx c = { &u, ( int )renderWidth, ( int )renderHeight };
    vec2 r = getPrevScreenPos( c, { mx, my }, { x, y } ); out[ 0 ] = r.x; out[ 1 ] = r.y;
}

} // extern "C"
