/*
 * glsl_shim.h -- the GLSL execution environment for oracle/_ref: TEST INFRASTRUCTURE ONLY.
 *
 * oracle/ref_glsl/build_ref.py turns the UNMODIFIED text of the reference's compute shaders
 * (/root/reference/Source/Shaders/*.comp with their #includes) into C++ translation units: it only rewrites what is
 * GLSL syntax and not C++ (layout(...) declarations, in/out/inout parameter qualifiers, `shared`, float literals
 * without suffix); every expression of the shaders is compiled as written.  This header supplies what the GLSL
 * language and the Vulkan binding model supply on a GPU:
 *
 *   vectors / matrices ... the reference's own vendored GLM (/root/reference/Tests/Libs/glm), swizzle operators on
 *   built-in functions ... thin non-template overloads over GLM / libm (GLSL converts int arguments implicitly, C++
 *                          template deduction does not)
 *   images / samplers .... texelFetch / imageLoad / imageStore on linear host arrays in the storage formats of
 *                          Source/Generated/ShaderCommonGLSL.h:487-575
 *   work groups .......... gl_*ID built-ins; `shared` variables are per-OS-thread statics and every invocation of a work
 *                          group is a ucontext fiber on that thread, so barrier() is a real rendezvous
 *
 * Behaviour that GLSL / Vulkan leave to the implementation is pinned exactly as in oracle/asvgf_oracle.cpp (and in the
 * CUDA path): out-of-range texelFetch/imageLoad returns zero, out-of-range imageStore is dropped; f32 -> f16 / uf11 /
 * uf10 / unorm8 conversions round to nearest even; round() is round-half-even; sin/cos are correctly rounded fp32;
 * pow/exp/exp2/log/log2/sqrt are libm fp32; min/max with a NaN operand return the other operand; mat * vec sums as GLM does.
 * (Out-of-range float -> int conversion is undefined in GLSL: here it is whatever GLM's static_cast yields on x86, INT_MIN;
 * the oracle saturates at +-1e9 -- both are far outside every image, so no fetch, store or test can tell them apart.)
 */
#pragma once

#define GLM_FORCE_SWIZZLE
#define GLM_FORCE_PURE          /* no SIMD paths: plain fp32 arithmetic, no FMA contraction (built with -ffp-contract=off) */
#ifndef _MSC_EXTENSIONS
    #define _MSC_EXTENSIONS 1   /* glm/detail/setup.hpp:75 enables the swizzle OPERATORS (v.xyz as in GLSL) only for the MS language extensions */
#endif
#include <glm/glm.hpp>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <ucontext.h>
#include <vector>

#ifdef M_PI
    #undef M_PI /* the shaders define their own (Utils.h:26) */
#endif

namespace glslshim
{

using glm::vec2; using glm::vec3; using glm::vec4;
using glm::ivec2; using glm::ivec3; using glm::ivec4;
using glm::uvec2; using glm::uvec3; using glm::uvec4;
using glm::bvec2; using glm::bvec3; using glm::bvec4;
using glm::mat2; using glm::mat3; using glm::mat4; using glm::mat3x4; using glm::mat4x3; using glm::mat3x2; using glm::mat2x3;
using glm::mat2x4; using glm::mat4x2;
typedef unsigned int uint;

// ------------------------------------------------------------------------------------------------------------------
// storage formats (Source/Generated/GenerateShaderCommon.py:802-887)
// ------------------------------------------------------------------------------------------------------------------
enum Format { FMT_NONE = 0, FMT_r8ui, FMT_r32ui, FMT_rgba32ui, FMT_rg8, FMT_rgba8, FMT_r8, FMT_r16f, FMT_rg16f, FMT_rgba16f, FMT_r32f, FMT_rgba32f, FMT_r11f_g11f_b10f };

struct ImageDesc
{
    void*    data;   // row-major, tightly packed
    int32_t  width, height;
    int32_t  format; // Format
};

inline uint32_t bitsOf( float f ) { uint32_t u; std::memcpy( &u, &f, 4 ); return u; }
inline float    floatOf( uint32_t u ) { float f; std::memcpy( &f, &u, 4 ); return f; }

// f32 -> small float with `ebits`=5 exponent bits (bias 15) and `mbits` mantissa bits, round to nearest even.
// signedFormat: fp16 (sign bit, inf/NaN kept).  Unsigned (uf11/uf10): NaN and negatives -> 0, overflow -> max finite.
inline uint32_t toSmallFloat( float f, int mbits, bool signedFormat )
{
    const uint32_t x    = bitsOf( f );
    const uint32_t sign = x >> 31;
    const uint32_t mag  = x & 0x7fffffffu;
    const uint32_t infCode = 31u << mbits, maxFinite = infCode - 1u;
    uint32_t       code;
    if( mag > 0x7f800000u ) // NaN
    {
        if( !signedFormat ) return 0u;
        code = infCode | ( 1u << ( mbits - 1 ) ) | ( ( mag >> ( 23 - mbits ) ) & ( ( 1u << mbits ) - 1u ) );
    }
    else
    {
        // value = m * 2^(e-127-23); small format: quantum 2^(-14-mbits) below 2^-14, else 2^(E-15-mbits)
        const int e = ( int )( mag >> 23 );
        uint64_t  m = ( mag & 0x7fffffu ) | ( e ? 0x800000u : 0u );
        int       exp2 = ( e ? e : 1 ) - 127 - 23; // value = m * 2^exp2
        int       E    = ( e ? e : 1 ) - 127;      // floor(log2 value) for normals
        int       q    = ( E < -14 ? -14 : E ) - mbits; // exponent of the target quantum
        // round m * 2^exp2 to a multiple of 2^q
        uint64_t  r;
        const int shift = q - exp2;
        if( shift <= 0 ) r = m << ( -shift );
        else if( shift > 40 ) r = 0;
        else
        {
            const uint64_t rem = m & ( ( 1ull << shift ) - 1ull ), half = 1ull << ( shift - 1 );
            r = m >> shift;
            if( rem > half || ( rem == half && ( r & 1ull ) ) ) r++;
        }
        // r * 2^q: r < 2^(mbits+1) (+carry); code = exponent field and mantissa
        if( r == 0 ) code = 0;
        else
        {
            int top = 63 - __builtin_clzll( r ); // r in [2^top, 2^(top+1))
            int Eq  = q + top;                   // floor(log2) of the rounded value
            if( Eq < -14 ) code = ( uint32_t )r; // subnormal of the small format (q = -14 - mbits)
            else
            {
                // normalise to mbits+1 significant bits
                uint64_t sig = top >= mbits ? r >> ( top - mbits ) : r << ( mbits - top );
                code         = ( ( uint32_t )( Eq + 15 ) << mbits ) | ( ( uint32_t )sig & ( ( 1u << mbits ) - 1u ) );
            }
        }
        if( code >= infCode ) code = signedFormat ? infCode : maxFinite;
        if( mag == 0x7f800000u ) code = signedFormat ? infCode : maxFinite;
    }
    if( !signedFormat ) return sign && mag != 0 ? 0u : code;
    return ( sign << ( 5 + mbits ) ) | code;
}
inline float fromSmallFloat( uint32_t code, int mbits, bool signedFormat )
{
    const uint32_t sign = signedFormat ? ( code >> ( 5 + mbits ) ) & 1u : 0u;
    const uint32_t e = ( code >> mbits ) & 31u, m = code & ( ( 1u << mbits ) - 1u );
    float          v;
    if( e == 0 ) v = std::ldexp( ( float )m, -14 - mbits );
    else if( e == 31 ) v = m ? floatOf( 0x7fc00000u ) : floatOf( 0x7f800000u );
    else v = std::ldexp( ( float )( m | ( 1u << mbits ) ), ( int )e - 15 - mbits );
    return sign ? -v : v;
}
inline uint32_t toUnorm8( float f )
{
    if( !( f > 0.0f ) ) return 0u; // NaN, negatives
    if( f > 1.0f ) f = 1.0f;
    return ( uint32_t )std::nearbyintf( f * 255.0f ); // default rounding mode: to nearest even
}

inline bool inside( const ImageDesc& d, ivec2 p ) { return p.x >= 0 && p.y >= 0 && p.x < d.width && p.y < d.height; }

inline vec4 loadFloat( const ImageDesc& d, ivec2 p )
{
    if( !d.data ) { std::fprintf( stderr, "glsl_shim: fetch from an unbound image\n" ); std::abort(); }
    if( !inside( d, p ) ) return vec4( 0.0f );
    const size_t i = ( size_t )p.y * ( size_t )d.width + ( size_t )p.x;
    switch( d.format )
    {
        case FMT_r16f: return vec4( fromSmallFloat( ( ( const uint16_t* )d.data )[ i ], 10, true ), 0.0f, 0.0f, 1.0f );
        case FMT_rg16f: { const uint16_t* t = ( const uint16_t* )d.data + 2 * i; return vec4( fromSmallFloat( t[ 0 ], 10, true ), fromSmallFloat( t[ 1 ], 10, true ), 0.0f, 1.0f ); }
        case FMT_rgba16f: { const uint16_t* t = ( const uint16_t* )d.data + 4 * i; return vec4( fromSmallFloat( t[ 0 ], 10, true ), fromSmallFloat( t[ 1 ], 10, true ), fromSmallFloat( t[ 2 ], 10, true ), fromSmallFloat( t[ 3 ], 10, true ) ); }
        case FMT_r32f: return vec4( ( ( const float* )d.data )[ i ], 0.0f, 0.0f, 1.0f );
        case FMT_rgba32f: { const float* t = ( const float* )d.data + 4 * i; return vec4( t[ 0 ], t[ 1 ], t[ 2 ], t[ 3 ] ); }
        case FMT_r8: return vec4( ( float )( ( const uint8_t* )d.data )[ i ] / 255.0f, 0.0f, 0.0f, 1.0f );
        case FMT_rg8: { const uint8_t* t = ( const uint8_t* )d.data + 2 * i; return vec4( ( float )t[ 0 ] / 255.0f, ( float )t[ 1 ] / 255.0f, 0.0f, 1.0f ); }
        case FMT_rgba8: { const uint8_t* t = ( const uint8_t* )d.data + 4 * i; return vec4( ( float )t[ 0 ] / 255.0f, ( float )t[ 1 ] / 255.0f, ( float )t[ 2 ] / 255.0f, ( float )t[ 3 ] / 255.0f ); }
        case FMT_r11f_g11f_b10f:
        {
            const uint32_t v = ( ( const uint32_t* )d.data )[ i ];
            return vec4( fromSmallFloat( v & 0x7ffu, 6, false ), fromSmallFloat( ( v >> 11 ) & 0x7ffu, 6, false ), fromSmallFloat( v >> 22, 5, false ), 1.0f );
        }
        default: std::fprintf( stderr, "glsl_shim: float fetch from an integer image (format %d)\n", d.format ); std::abort();
    }
}
inline uvec4 loadUint( const ImageDesc& d, ivec2 p )
{
    if( !d.data ) { std::fprintf( stderr, "glsl_shim: fetch from an unbound image\n" ); std::abort(); }
    if( !inside( d, p ) ) return uvec4( 0u );
    const size_t i = ( size_t )p.y * ( size_t )d.width + ( size_t )p.x;
    switch( d.format )
    {
        case FMT_r8ui: return uvec4( ( ( const uint8_t* )d.data )[ i ], 0u, 0u, 1u );
        case FMT_r32ui: return uvec4( ( ( const uint32_t* )d.data )[ i ], 0u, 0u, 1u );
        case FMT_rgba32ui: { const uint32_t* t = ( const uint32_t* )d.data + 4 * i; return uvec4( t[ 0 ], t[ 1 ], t[ 2 ], t[ 3 ] ); }
        default: std::fprintf( stderr, "glsl_shim: integer fetch from a float image (format %d)\n", d.format ); std::abort();
    }
}
inline void storeFloat( const ImageDesc& d, ivec2 p, vec4 v )
{
    if( !d.data ) { std::fprintf( stderr, "glsl_shim: store to an unbound image\n" ); std::abort(); }
    if( !inside( d, p ) ) return;
    const size_t i = ( size_t )p.y * ( size_t )d.width + ( size_t )p.x;
    switch( d.format )
    {
        case FMT_r16f: ( ( uint16_t* )d.data )[ i ] = ( uint16_t )toSmallFloat( v.x, 10, true ); break;
        case FMT_rg16f: { uint16_t* t = ( uint16_t* )d.data + 2 * i; t[ 0 ] = ( uint16_t )toSmallFloat( v.x, 10, true ); t[ 1 ] = ( uint16_t )toSmallFloat( v.y, 10, true ); break; }
        case FMT_rgba16f: { uint16_t* t = ( uint16_t* )d.data + 4 * i; for( int c = 0; c < 4; c++ ) t[ c ] = ( uint16_t )toSmallFloat( v[ c ], 10, true ); break; }
        case FMT_r32f: ( ( float* )d.data )[ i ] = v.x; break;
        case FMT_rgba32f: { float* t = ( float* )d.data + 4 * i; for( int c = 0; c < 4; c++ ) t[ c ] = v[ c ]; break; }
        case FMT_r8: ( ( uint8_t* )d.data )[ i ] = ( uint8_t )toUnorm8( v.x ); break;
        case FMT_rg8: { uint8_t* t = ( uint8_t* )d.data + 2 * i; t[ 0 ] = ( uint8_t )toUnorm8( v.x ); t[ 1 ] = ( uint8_t )toUnorm8( v.y ); break; }
        case FMT_rgba8: { uint8_t* t = ( uint8_t* )d.data + 4 * i; for( int c = 0; c < 4; c++ ) t[ c ] = ( uint8_t )toUnorm8( v[ c ] ); break; }
        case FMT_r11f_g11f_b10f: ( ( uint32_t* )d.data )[ i ] = toSmallFloat( v.x, 6, false ) | ( toSmallFloat( v.y, 6, false ) << 11 ) | ( toSmallFloat( v.z, 5, false ) << 22 ); break;
        default: std::fprintf( stderr, "glsl_shim: float store to an integer image (format %d)\n", d.format ); std::abort();
    }
}
inline void storeUint( const ImageDesc& d, ivec2 p, uvec4 v )
{
    if( !d.data ) { std::fprintf( stderr, "glsl_shim: store to an unbound image\n" ); std::abort(); }
    if( !inside( d, p ) ) return;
    const size_t i = ( size_t )p.y * ( size_t )d.width + ( size_t )p.x;
    switch( d.format )
    {
        case FMT_r8ui: ( ( uint8_t* )d.data )[ i ] = ( uint8_t )v.x; break;
        case FMT_r32ui: ( ( uint32_t* )d.data )[ i ] = v.x; break;
        case FMT_rgba32ui: { uint32_t* t = ( uint32_t* )d.data + 4 * i; for( int c = 0; c < 4; c++ ) t[ c ] = v[ c ]; break; }
        default: std::fprintf( stderr, "glsl_shim: integer store to a float image (format %d)\n", d.format ); std::abort();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// descriptor sets: set DESC_SET_FRAMEBUFFERS = 76 storage images + 76 samplers of the same images
// (ShaderCommonGLSL.h:487-667, binding = FB_IMAGE_INDEX_* and 76 + FB_IMAGE_INDEX_*); set DESC_SET_GLOBAL_UNIFORM = the UBO
// ------------------------------------------------------------------------------------------------------------------
constexpr int kFramebufferImageCount = 76;

struct DispatchContext
{
    const ImageDesc* images;      // kFramebufferImageCount entries, cur/prev swap already applied (Framebuffers.cpp:35-50)
    void*            buffers[ 8 ][ 4 ]; // uniform / storage blocks by [descriptor set][binding] (the UBO, the tonemapping buffer)
    uint32_t         specConstant[ 4 ];
    uint32_t         pushConstant[ 32 ];
};
extern thread_local const DispatchContext* g_ctx;

struct image2D { const ImageDesc* d; };
struct uimage2D { const ImageDesc* d; };
struct sampler2D { const ImageDesc* d; };
struct usampler2D { const ImageDesc* d; };

inline void* uniformBlock( int set, int binding )
{
    void* p = g_ctx->buffers[ set ][ binding ];
    if( !p ) { std::fprintf( stderr, "glsl_shim: no buffer bound at set %d binding %d\n", set, binding ); std::abort(); }
    return p;
}
inline void* storageBlock( int set, int binding ) { return uniformBlock( set, binding ); }

// 3-D volumetric images (ShaderCommonGLSLFunc.h:225-238): declared by the headers, not image-space -> any access aborts
struct image3D {}; struct sampler3D {}; struct texture2DArray {}; struct sampler2DArray {}; struct samplerCube {};
inline vec4 texelFetch( const texture2DArray&, ivec3, int ) { std::fprintf( stderr, "glsl_shim: blue-noise fetch is outside the image-space path\n" ); std::abort(); }
inline vec4 notImageSpace() { std::fprintf( stderr, "glsl_shim: 3-D volume access is outside the image-space path\n" ); std::abort(); }
inline vec4 imageLoad( const image3D&, ivec3 ) { return notImageSpace(); }
inline void imageStore( const image3D&, ivec3, vec4 ) { notImageSpace(); }
inline vec4 texture( const sampler3D&, vec3 ) { return notImageSpace(); }
inline vec4 textureLod( const sampler3D&, vec3, float ) { return notImageSpace(); }
inline vec4 texelFetch( const sampler3D&, ivec3, int ) { return notImageSpace(); }
inline ivec3 textureSize( const sampler3D&, int ) { notImageSpace(); return ivec3( 0 ); }
inline ivec3 imageSize( const image3D& ) { notImageSpace(); return ivec3( 0 ); }

template< typename T > inline T bindImage( int set, int binding, int declaredFormat )
{
    ( void )set; // the only descriptor set with images on this path is the framebuffers set
    const ImageDesc* d = &g_ctx->images[ binding % kFramebufferImageCount ];
    if( declaredFormat != FMT_NONE && d->data && d->format != declaredFormat )
    {
        std::fprintf( stderr, "glsl_shim: binding %d is declared with format %d but bound with format %d\n", binding, declaredFormat, d->format );
        std::abort();
    }
    return T{ d };
}
inline uint32_t specConstant( int id, uint32_t defaultValue ) { ( void )defaultValue; return g_ctx->specConstant[ id ]; }

inline vec4  texelFetch( const sampler2D& s, ivec2 p, int lod ) { ( void )lod; return loadFloat( *s.d, p ); }
inline uvec4 texelFetch( const usampler2D& s, ivec2 p, int lod ) { ( void )lod; return loadUint( *s.d, p ); }
inline vec4  imageLoad( const image2D& s, ivec2 p ) { return loadFloat( *s.d, p ); }
inline uvec4 imageLoad( const uimage2D& s, ivec2 p ) { return loadUint( *s.d, p ); }
inline void  imageStore( const image2D& s, ivec2 p, vec4 v ) { storeFloat( *s.d, p, v ); }
inline void  imageStore( const uimage2D& s, ivec2 p, uvec4 v ) { storeUint( *s.d, p, v ); }
inline ivec2 imageSize( const image2D& s ) { return ivec2( s.d->width, s.d->height ); }
inline ivec2 imageSize( const uimage2D& s ) { return ivec2( s.d->width, s.d->height ); }
inline ivec2 textureSize( const sampler2D& s, int lod ) { ( void )lod; return ivec2( s.d->width, s.d->height ); }

// ------------------------------------------------------------------------------------------------------------------
// built-in functions.  Non-template overloads: GLSL converts int arguments to float implicitly.
// ------------------------------------------------------------------------------------------------------------------
inline float pinnedSin( float x ) { return ( float )std::sin( ( double )x ); }
inline float pinnedCos( float x ) { return ( float )std::cos( ( double )x ); }
inline float pinnedRound( float x ) { return std::nearbyintf( x ); } // round-half-even (default rounding mode)

#define GLSLSHIM_UNARY( NAME, EXPR )                                                                                   \
    inline float NAME( float x ) { return EXPR; }                                                                      \
    inline vec2  NAME( const vec2& v ) { return vec2( NAME( v.x ), NAME( v.y ) ); }                                    \
    inline vec3  NAME( const vec3& v ) { return vec3( NAME( v.x ), NAME( v.y ), NAME( v.z ) ); }                       \
    inline vec4  NAME( const vec4& v ) { return vec4( NAME( v.x ), NAME( v.y ), NAME( v.z ), NAME( v.w ) ); }
GLSLSHIM_UNARY( sin, pinnedSin( x ) )
GLSLSHIM_UNARY( cos, pinnedCos( x ) )
GLSLSHIM_UNARY( tan, std::tan( x ) )
GLSLSHIM_UNARY( asin, std::asin( x ) )
GLSLSHIM_UNARY( acos, std::acos( x ) )
GLSLSHIM_UNARY( exp, std::exp( x ) )
GLSLSHIM_UNARY( log, std::log( x ) )
GLSLSHIM_UNARY( exp2, std::exp2( x ) )
GLSLSHIM_UNARY( log2, std::log2( x ) )
GLSLSHIM_UNARY( sqrt, std::sqrt( x ) )
GLSLSHIM_UNARY( inversesqrt, 1.0f / std::sqrt( x ) )
GLSLSHIM_UNARY( floor, std::floor( x ) )
GLSLSHIM_UNARY( ceil, std::ceil( x ) )
GLSLSHIM_UNARY( trunc, std::trunc( x ) )
GLSLSHIM_UNARY( round, pinnedRound( x ) )
GLSLSHIM_UNARY( fract, x - std::floor( x ) )
GLSLSHIM_UNARY( abs, std::fabs( x ) )
GLSLSHIM_UNARY( sign, ( x > 0.0f ? 1.0f : x < 0.0f ? -1.0f : 0.0f ) )
GLSLSHIM_UNARY( saturateShim_, ( x < 0.0f ? 0.0f : x > 1.0f ? 1.0f : x ) )
#undef GLSLSHIM_UNARY
inline int   abs( int x ) { return x < 0 ? -x : x; }
inline int   sign( int x ) { return x > 0 ? 1 : x < 0 ? -1 : 0; }
inline ivec2 abs( const ivec2& v ) { return ivec2( abs( v.x ), abs( v.y ) ); }
inline float atan( float y, float x ) { return std::atan2( y, x ); }
inline float atan( float x ) { return std::atan( x ); }

// min/max/clamp.  With a NaN operand GLSL leaves the result to the implementation; the pin (oracle and CUDA path alike) is
// IEEE-754 minNum/maxNum -- the non-NaN operand -- which is what NVIDIA's FMNMX instruction returns.  clamp = min(max(x, lo), hi).
inline float min( float a, float b ) { return std::fmin( a, b ); }
inline float max( float a, float b ) { return std::fmax( a, b ); }
inline int   min( int a, int b ) { return b < a ? b : a; }
inline int   max( int a, int b ) { return a < b ? b : a; }
inline uint  min( uint a, uint b ) { return b < a ? b : a; }
inline uint  max( uint a, uint b ) { return a < b ? b : a; }
inline float min( float a, int b ) { return min( a, ( float )b ); }
inline float max( float a, int b ) { return max( a, ( float )b ); }
inline float min( int a, float b ) { return min( ( float )a, b ); }
inline float max( int a, float b ) { return max( ( float )a, b ); }
inline int   min( int a, uint b ) { return min( a, ( int )b ); }
inline int   max( int a, uint b ) { return max( a, ( int )b ); }
inline uint  min( uint a, int b ) { return min( a, ( uint )b ); }
inline uint  max( uint a, int b ) { return max( a, ( uint )b ); }
inline uint  clamp( uint x, int lo, uint hi ) { return min( max( x, ( uint )lo ), hi ); }
inline uint  clamp( uint x, int lo, int hi ) { return min( max( x, ( uint )lo ), ( uint )hi ); }
inline int   clamp( int x, int lo, uint hi ) { return min( max( x, lo ), ( int )hi ); }
inline float clamp( float x, float lo, float hi ) { return min( max( x, lo ), hi ); }
inline int   clamp( int x, int lo, int hi ) { return min( max( x, lo ), hi ); }
inline uint  clamp( uint x, uint lo, uint hi ) { return min( max( x, lo ), hi ); }
inline float clamp( float x, int lo, int hi ) { return clamp( x, ( float )lo, ( float )hi ); }
inline float clamp( float x, float lo, int hi ) { return clamp( x, lo, ( float )hi ); }
inline float clamp( float x, int lo, float hi ) { return clamp( x, ( float )lo, hi ); }
#define GLSLSHIM_VEC_MINMAX( V, N )                                                                                    \
    inline V min( const V& a, const V& b ) { V r; for( int i = 0; i < N; i++ ) r[ i ] = min( a[ i ], b[ i ] ); return r; } \
    inline V max( const V& a, const V& b ) { V r; for( int i = 0; i < N; i++ ) r[ i ] = max( a[ i ], b[ i ] ); return r; } \
    inline V min( const V& a, float b ) { V r; for( int i = 0; i < N; i++ ) r[ i ] = min( a[ i ], b ); return r; }     \
    inline V max( const V& a, float b ) { V r; for( int i = 0; i < N; i++ ) r[ i ] = max( a[ i ], b ); return r; }     \
    inline V clamp( const V& x, const V& lo, const V& hi ) { return min( max( x, lo ), hi ); }                         \
    inline V clamp( const V& x, float lo, float hi ) { return min( max( x, lo ), hi ); }                               \
    inline V mix( const V& a, const V& b, float t ) { return a * ( 1.0f - t ) + b * t; }                               \
    inline V mix( const V& a, const V& b, const V& t ) { return a * ( V( 1.0f ) - t ) + b * t; }                       \
    inline float dot( const V& a, const V& b ) { float s = a[ 0 ] * b[ 0 ]; for( int i = 1; i < N; i++ ) s = s + a[ i ] * b[ i ]; return s; } \
    inline float length( const V& a ) { return std::sqrt( dot( a, a ) ); }                                             \
    inline float distance( const V& a, const V& b ) { return length( a - b ); }                                        \
    inline V normalize( const V& a ) { return a * ( 1.0f / std::sqrt( dot( a, a ) ) ); }                               \
    inline V pow( const V& a, const V& b ) { V r; for( int i = 0; i < N; i++ ) r[ i ] = std::pow( a[ i ], b[ i ] ); return r; } \
    inline V step( float e, const V& x ) { V r; for( int i = 0; i < N; i++ ) r[ i ] = x[ i ] < e ? 0.0f : 1.0f; return r; } \
    inline V mod( const V& a, float b ) { V r; for( int i = 0; i < N; i++ ) r[ i ] = a[ i ] - b * std::floor( a[ i ] / b ); return r; }
GLSLSHIM_VEC_MINMAX( vec2, 2 )
GLSLSHIM_VEC_MINMAX( vec3, 3 )
GLSLSHIM_VEC_MINMAX( vec4, 4 )
#undef GLSLSHIM_VEC_MINMAX
inline ivec2 min( const ivec2& a, const ivec2& b ) { return ivec2( min( a.x, b.x ), min( a.y, b.y ) ); }
inline ivec2 max( const ivec2& a, const ivec2& b ) { return ivec2( max( a.x, b.x ), max( a.y, b.y ) ); }
inline ivec2 clamp( const ivec2& x, const ivec2& lo, const ivec2& hi ) { return min( max( x, lo ), hi ); }
inline float mix( float a, float b, float t ) { return a * ( 1.0f - t ) + b * t; }
inline float pow( float a, float b ) { return std::pow( a, b ); }
inline float step( float e, float x ) { return x < e ? 0.0f : 1.0f; }
inline float mod( float a, float b ) { return a - b * std::floor( a / b ); }
inline float smoothstep( float e0, float e1, float x ) { const float t = clamp( ( x - e0 ) / ( e1 - e0 ), 0.0f, 1.0f ); return t * t * ( 3.0f - 2.0f * t ); }
inline float length( float a ) { return std::fabs( a ); }
inline vec3  cross( const vec3& a, const vec3& b ) { return vec3( a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y ); }
inline vec3  reflect( const vec3& I, const vec3& N ) { return I - N * ( 2.0f * dot( N, I ) ); }
inline bool  isnan( float x ) { return std::isnan( x ); }
inline bool  isinf( float x ) { return std::isinf( x ); }
inline bool  any( const bvec2& b ) { return b.x || b.y; }
inline bool  any( const bvec3& b ) { return b.x || b.y || b.z; }
inline bool  all( const bvec2& b ) { return b.x && b.y; }
inline bool  all( const bvec3& b ) { return b.x && b.y && b.z; }
inline uint  floatBitsToUint( float f ) { return bitsOf( f ); }
inline int   floatBitsToInt( float f ) { return ( int )bitsOf( f ); }
inline float uintBitsToFloat( uint u ) { return floatOf( u ); }
inline float intBitsToFloat( int i ) { return floatOf( ( uint32_t )i ); }
inline uvec3 floatBitsToUint( const vec3& v ) { return uvec3( bitsOf( v.x ), bitsOf( v.y ), bitsOf( v.z ) ); }
inline vec3  uintBitsToFloat( const uvec3& v ) { return vec3( floatOf( v.x ), floatOf( v.y ), floatOf( v.z ) ); }
inline uint  packHalf2x16( const vec2& v ) { return toSmallFloat( v.x, 10, true ) | ( toSmallFloat( v.y, 10, true ) << 16 ); }
inline vec2  unpackHalf2x16( uint p ) { return vec2( fromSmallFloat( p & 0xffffu, 10, true ), fromSmallFloat( p >> 16, 10, true ) ); }
inline uint  packUnorm4x8( const vec4& v ) { return toUnorm8( v.x ) | ( toUnorm8( v.y ) << 8 ) | ( toUnorm8( v.z ) << 16 ) | ( toUnorm8( v.w ) << 24 ); }
inline vec4  unpackUnorm4x8( uint p ) { return vec4( ( float )( p & 0xffu ) / 255.0f, ( float )( ( p >> 8 ) & 0xffu ) / 255.0f, ( float )( ( p >> 16 ) & 0xffu ) / 255.0f, ( float )( p >> 24 ) / 255.0f ); }
inline uint  bitfieldExtract( uint v, int offset, int bits ) { return bits >= 32 ? v >> offset : ( v >> offset ) & ( ( 1u << bits ) - 1u ); }
inline uint  bitfieldInsert( uint base, uint insert, int offset, int bits )
{
    const uint mask = ( bits >= 32 ? ~0u : ( ( 1u << bits ) - 1u ) ) << offset;
    return ( base & ~mask ) | ( ( insert << offset ) & mask );
}
inline int   findMSB( uint v ) { return v ? 31 - __builtin_clz( v ) : -1; }
inline int   bitCount( uint v ) { return __builtin_popcount( v ); }
inline mat4  inverse( const mat4& m ) { return glm::inverse( m ); }
inline mat3  inverse( const mat3& m ) { return glm::inverse( m ); }
inline mat4  transpose( const mat4& m ) { return glm::transpose( m ); }
inline mat3  transpose( const mat3& m ) { return glm::transpose( m ); }

// operators GLSL has on swizzles that GLM's swizzle proxies lack
template< int N, typename T, glm::qualifier Q, int E0, int E1, int E2, int E3 >
inline glm::vec< N, T, Q > operator-( const glm::detail::_swizzle< N, T, Q, E0, E1, E2, E3 >& s ) { return -s(); }
template< int N, typename T, glm::qualifier Q, int E0, int E1, int E2, int E3 >
inline void operator*=( glm::detail::_swizzle< N, T, Q, E0, E1, E2, E3 >& s, T f ) { s = s() * f; }
template< int N, typename T, glm::qualifier Q, int E0, int E1, int E2, int E3 >
inline void operator/=( glm::detail::_swizzle< N, T, Q, E0, E1, E2, E3 >& s, T f ) { s = s() / f; }
template< int N, typename T, glm::qualifier Q, int E0, int E1, int E2, int E3 >
inline void operator+=( glm::detail::_swizzle< N, T, Q, E0, E1, E2, E3 >& s, T f ) { s = s() + f; }
template< int N, typename T, glm::qualifier Q, int E0, int E1, int E2, int E3 >
inline void operator-=( glm::detail::_swizzle< N, T, Q, E0, E1, E2, E3 >& s, T f ) { s = s() - f; }

// vec (op)= swizzle: GLM's compound assignments take the swizzle proxy for a scalar
template< int N, typename T, glm::qualifier Q, int E0, int E1, int E2, int E3 >
inline glm::vec< N, T, Q >& operator+=( glm::vec< N, T, Q >& a, const glm::detail::_swizzle< N, T, Q, E0, E1, E2, E3 >& b ) { return a += b(); }
template< int N, typename T, glm::qualifier Q, int E0, int E1, int E2, int E3 >
inline glm::vec< N, T, Q >& operator-=( glm::vec< N, T, Q >& a, const glm::detail::_swizzle< N, T, Q, E0, E1, E2, E3 >& b ) { return a -= b(); }
template< int N, typename T, glm::qualifier Q, int E0, int E1, int E2, int E3 >
inline glm::vec< N, T, Q >& operator*=( glm::vec< N, T, Q >& a, const glm::detail::_swizzle< N, T, Q, E0, E1, E2, E3 >& b ) { return a *= b(); }

// atomics on shared / buffer memory (work groups run on several OS threads)
inline uint atomicAdd( uint& mem, uint v ) { return __atomic_fetch_add( &mem, v, __ATOMIC_RELAXED ); }
inline uint atomicAdd( uint& mem, int v ) { return __atomic_fetch_add( &mem, ( uint )v, __ATOMIC_RELAXED ); }

// swizzle (op) scalar: GLM's swizzle proxies only define these for '-' and '*' (glm/detail/_swizzle.hpp:277-278)
#define GLSLSHIM_SWIZZLE_SCALAR( OP )                                                                                  \
    template< int N, typename T, glm::qualifier Q, int E0, int E1, int E2, int E3 >                                    \
    inline glm::vec< N, T, Q > operator OP( const glm::detail::_swizzle< N, T, Q, E0, E1, E2, E3 >& a, const T& b ) { return a() OP b; } \
    template< int N, typename T, glm::qualifier Q, int E0, int E1, int E2, int E3 >                                    \
    inline glm::vec< N, T, Q > operator OP( const T& a, const glm::detail::_swizzle< N, T, Q, E0, E1, E2, E3 >& b ) { return a OP b(); }
GLSLSHIM_SWIZZLE_SCALAR( + )
GLSLSHIM_SWIZZLE_SCALAR( / )
#undef GLSLSHIM_SWIZZLE_SCALAR

// implicit int -> float promotions of GLSL expressions that GLM's operator templates do not deduce.  Constrained to
// INTEGER scalars: a float operand must never be converted to int on the way in.
#define GLSLSHIM_INT template< typename I, typename = std::enable_if_t< std::is_integral< I >::value > >
#define GLSLSHIM_MIXED_OPS( V, IV )                                                                                    \
    GLSLSHIM_INT inline V operator*( const V& a, I b ) { return a * ( float )b; }                                      \
    GLSLSHIM_INT inline V operator*( I a, const V& b ) { return ( float )a * b; }                                      \
    GLSLSHIM_INT inline V operator/( const V& a, I b ) { return a / ( float )b; }                                      \
    GLSLSHIM_INT inline V operator/( I a, const V& b ) { return ( float )a / b; }                                      \
    GLSLSHIM_INT inline V operator+( const V& a, I b ) { return a + ( float )b; }                                      \
    GLSLSHIM_INT inline V operator+( I a, const V& b ) { return ( float )a + b; }                                      \
    GLSLSHIM_INT inline V operator-( const V& a, I b ) { return a - ( float )b; }                                      \
    GLSLSHIM_INT inline V operator-( I a, const V& b ) { return ( float )a - b; }                                      \
    inline V operator+( const IV& a, const V& b ) { return V( a ) + b; }                                               \
    inline V operator+( const V& a, const IV& b ) { return a + V( b ); }                                               \
    inline V operator-( const IV& a, const V& b ) { return V( a ) - b; }                                               \
    inline V operator-( const V& a, const IV& b ) { return a - V( b ); }                                               \
    inline V operator*( const IV& a, const V& b ) { return V( a ) * b; }                                               \
    inline V operator*( const V& a, const IV& b ) { return a * V( b ); }                                               \
    inline V operator/( const IV& a, const V& b ) { return V( a ) / b; }                                               \
    inline V operator/( const V& a, const IV& b ) { return a / V( b ); }                                               \
    inline V operator+( const IV& a, float b ) { return V( a ) + b; }                                                  \
    inline V operator-( const IV& a, float b ) { return V( a ) - b; }                                                  \
    inline V operator*( const IV& a, float b ) { return V( a ) * b; }                                                  \
    inline V operator/( const IV& a, float b ) { return V( a ) / b; }
GLSLSHIM_MIXED_OPS( vec2, ivec2 )
GLSLSHIM_MIXED_OPS( vec3, ivec3 )
GLSLSHIM_MIXED_OPS( vec4, ivec4 )
#undef GLSLSHIM_MIXED_OPS

// ------------------------------------------------------------------------------------------------------------------
// invocations, work groups, barrier()
// ------------------------------------------------------------------------------------------------------------------
// gl_*ID are uvec3 in GLSL; the shaders write ivec2( gl_GlobalInvocationID ) and ivec2( gl_WorkGroupID.xy ), conversions
// GLM only offers between vectors of its own types -- so the built-ins get a type that converts to both signednesses
struct BuiltinXY
{
    uint x, y;
    operator uvec2() const { return uvec2( x, y ); }
    operator ivec2() const { return ivec2( ( int )x, ( int )y ); }
};
struct BuiltinUvec3
{
    union { struct { uint x, y, z; }; BuiltinXY xy; };
    BuiltinUvec3() : x( 0 ), y( 0 ), z( 0 ) {}
    BuiltinUvec3( const uvec3& v ) : x( v.x ), y( v.y ), z( v.z ) {}
    operator uvec3() const { return uvec3( x, y, z ); }
    operator ivec3() const { return ivec3( ( int )x, ( int )y, ( int )z ); }
    operator ivec2() const { return ivec2( ( int )x, ( int )y ); }
    operator uvec2() const { return uvec2( x, y ); }
};
struct Builtins
{
    uvec3 gl_NumWorkGroups, gl_WorkGroupID, gl_LocalInvocationID, gl_GlobalInvocationID, gl_WorkGroupSize;
    uint  gl_LocalInvocationIndex;
};
extern thread_local Builtins g_builtins; // of the invocation that is currently running on this OS thread

struct InvocationBase
{
    const BuiltinUvec3 gl_NumWorkGroups = g_builtins.gl_NumWorkGroups, gl_WorkGroupID = g_builtins.gl_WorkGroupID,
                       gl_LocalInvocationID = g_builtins.gl_LocalInvocationID, gl_GlobalInvocationID = g_builtins.gl_GlobalInvocationID,
                       gl_WorkGroupSize = g_builtins.gl_WorkGroupSize;
    const uint  gl_LocalInvocationIndex = g_builtins.gl_LocalInvocationIndex;
};

void barrier();                       // rendezvous of the invocations of the current work group
inline void memoryBarrierShared() {}  // invocations of a group run on one OS thread: memory is always coherent
inline void memoryBarrier() {}
// CmLuminanceHistogram.comp:67,81 separates "zero the bins / count / flush" with groupMemoryBarrier() only -- a memory
// barrier without execution barrier, i.e. formally a race that GPUs close by timing (the eight warps of a group run in
// near lockstep).  Pinned to the intended meaning, as in the oracle: a rendezvous of the work group.
inline void groupMemoryBarrier() { barrier(); }

typedef void ( *InvocationFn )(); // constructs the shader object (binding its resources) and calls main()
// runs groupsX x groupsY x 1 work groups of localX x localY x 1 invocations (vkCmdDispatch); work groups are spread over
// OpenMP threads, the invocations of one group run as fibers on one thread
void dispatch( const DispatchContext& ctx, InvocationFn fn, uint32_t localX, uint32_t localY, uint32_t groupsX, uint32_t groupsY, bool usesBarrier );

} // namespace glslshim
