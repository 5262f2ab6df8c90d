/*
 * asvgf_synth.cpp -- deterministic synthetic G-buffer generator for the A-SVGF denoiser path (CPU, OpenMP).
 *
 * It plays the role of the part of RTGL1 that is out of scope (the path tracer): for a frame number it fills
 * every INPUT image of the denoiser with the producer semantics documented in SURVEY.md section 8(a-0):
 *   Source/Shaders/RaygenPrimary.inl:98-123,325-338  (G-buffer stores, sky conventions)
 *   Source/Shaders/HitInfo.inl:271-336               (motion vectors, depth gradient, linear depth)
 *   Source/Shaders/RtRaygenDirect.rgen:63-67         (demodulated radiance, GradientInputs)
 *   Source/Matrix.cpp:380-395                        (projection convention)
 * Scene: ground plane, back wall, 6 spheres (2 moving), 2 boxes, one point light; pin-hole camera orbiting
 * 0.5 deg/frame with a 0.02-unit dolly.  1-spp-like lognormal noise + 0.1% fireflies, PCG-hashed per
 * (pixel, frame, channel, seed).  A light switch at frame 24 (x4 inside one floor rectangle, x0 inside another)
 * drives the gradient (anti-lag) path; FwdLum is the previous frame's sample re-shaded under current lighting.
 *
 * Stateless: any frame can be generated independently.  Not part of the product path and not the oracle.
 */
#include "../include/asvgf_layout.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace
{
struct V3 { float x, y, z; };
inline V3    operator+( V3 a, V3 b ) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3    operator-( V3 a, V3 b ) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3    operator*( V3 a, float s ) { return { a.x * s, a.y * s, a.z * s }; }
inline V3    operator*( V3 a, V3 b ) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline float dot( V3 a, V3 b ) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3    cross( V3 a, V3 b ) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline float length( V3 a ) { return sqrtf( dot( a, a ) ); }
inline V3    normalize( V3 a ) { return a * ( 1.0f / length( a ) ); }
inline float lum( V3 c ) { return 0.2125f * c.x + 0.7154f * c.y + 0.0721f * c.z; }

const float PI_F = 3.14159265358979323846f;

// ---- storage formats (producer side) ----
inline uint32_t f2u( float f ) { uint32_t u; std::memcpy( &u, &f, 4 ); return u; }
inline float    u2f( uint32_t u ) { float f; std::memcpy( &f, &u, 4 ); return f; }
uint16_t toHalf( float f )
{
    uint32_t x = f2u( f ), sign = ( x >> 16 ) & 0x8000u;
    x &= 0x7fffffffu;
    if( x >= 0x7f800000u ) return ( uint16_t )( sign | 0x7c00u | ( x > 0x7f800000u ? 0x200u : 0u ) );
    if( x >= 0x47800000u ) return ( uint16_t )( sign | 0x7c00u );
    if( x < 0x38800000u )
    {
        if( x < 0x33000000u ) return ( uint16_t )sign;
        uint32_t e = x >> 23, m = ( x & 0x7fffffu ) | 0x800000u, shift = 126u - e;
        uint32_t r = m >> shift, rem = m & ( ( 1u << shift ) - 1u ), half = 1u << ( shift - 1u );
        if( rem > half || ( rem == half && ( r & 1u ) ) ) r++;
        return ( uint16_t )( sign | r );
    }
    uint32_t r = ( x - 0x38000000u ) >> 13, rem = x & 0x1fffu;
    if( rem > 0x1000u || ( rem == 0x1000u && ( r & 1u ) ) ) r++;
    return ( uint16_t )( sign | r );
}
float fromHalf( uint16_t h )
{
    uint32_t sign = ( uint32_t )( h & 0x8000u ) << 16, e = ( h >> 10 ) & 0x1fu, m = h & 0x3ffu;
    if( e == 0 ) { float v = ( float )m * 5.9604644775390625e-08f; return sign ? -v : v; }
    if( e == 31 ) return u2f( sign | 0x7f800000u | ( m << 13 ) );
    return u2f( sign | ( ( e + 112u ) << 23 ) | ( m << 13 ) );
}
uint32_t toUfloat( float f, uint32_t mbits ) // truncating is fine for a producer
{
    if( !( f > 0.0f ) ) return 0u;
    uint32_t x = f2u( f );
    uint32_t maxCode = ( 30u << mbits ) | ( ( 1u << mbits ) - 1u );
    if( x >= 0x47800000u ) return maxCode;
    if( x < 0x38800000u ) return 0u;
    uint32_t r = ( x - 0x38000000u ) >> ( 23u - mbits );
    return r > maxCode ? maxCode : r;
}
uint32_t packUf( V3 c ) { return toUfloat( c.x, 6 ) | ( toUfloat( c.y, 6 ) << 11 ) | ( toUfloat( c.z, 5 ) << 22 ); }
uint32_t packE5( V3 c ) // shared-exponent RGB9E5
{
    auto cl = []( float v ) { return std::min( std::max( v, 0.0f ), 65408.0f ); };
    c = { cl( c.x ), cl( c.y ), cl( c.z ) };
    float m = std::max( c.x, std::max( c.y, c.z ) );
    if( !( m > 0.0f ) ) return 0u;
    int e  = std::max( -16, ( int )ilogbf( m ) ) + 1 + 15;
    int ms = ( int )rintf( m * ldexpf( 1.0f, -( e - 24 ) ) );
    if( ms == 512 ) e++;
    float s = ldexpf( 1.0f, -( e - 24 ) );
    return ( ( uint32_t )e << 27 ) | ( ( uint32_t )rintf( c.z * s ) << 18 ) | ( ( uint32_t )rintf( c.y * s ) << 9 ) | ( uint32_t )rintf( c.x * s );
}
uint32_t packNormal( V3 n ) // Utils.h:108-122
{
    float nz    = std::min( 1.0f, std::max( -1.0f, n.z ) );
    float phi   = acosf( nz );
    float theta = atan2f( n.y, n.x );
    if( theta < 0 ) theta += 2.0f * PI_F;
    uint32_t j = ( uint32_t )rintf( phi * 65535.0f / PI_F );
    uint32_t k = ( ( uint32_t )rintf( theta * 65536.0f / ( 2.0f * PI_F ) ) ) % 65536u;
    return ( std::min( j, 65535u ) << 16 ) | k;
}

// ---- hashing ----
inline uint32_t pcg( uint32_t v )
{
    uint32_t s = v * 747796405u + 2891336453u;
    uint32_t w = ( ( s >> ( ( s >> 28u ) + 4u ) ) ^ s ) * 277803737u;
    return ( w >> 22u ) ^ w;
}
inline uint32_t hash5( uint32_t x, uint32_t y, uint32_t frame, uint32_t channel, uint32_t seed )
{
    return pcg( x + pcg( y + pcg( frame + pcg( channel + pcg( seed ) ) ) ) );
}
inline float u01( uint32_t h ) { return ( ( h >> 8 ) + 0.5f ) * ( 1.0f / 16777216.0f ); }

// ---- scene ----
struct Material { V3 albedo; float roughness, metallic; };
struct Sphere { V3 c; float r; V3 vel; Material m; };
struct Box { V3 lo, hi; Material m; };

const Sphere kSpheres[ 6 ] = {
    { { -3.0f, 1.0f, -4.0f }, 1.0f, { 0.00f, 0.0f, 0.00f }, { { 0.80f, 0.20f, 0.20f }, 0.30f, 0.0f } },
    { { 0.0f, 1.2f, -6.0f }, 1.2f, { 0.00f, 0.0f, 0.00f }, { { 0.90f, 0.90f, 0.90f }, 0.05f, 1.0f } },
    { { 3.0f, 0.8f, -3.5f }, 0.8f, { -0.06f, 0.0f, 0.00f }, { { 0.20f, 0.70f, 0.30f }, 0.60f, 0.0f } },
    { { -1.5f, 0.5f, -2.0f }, 0.5f, { 0.00f, 0.0f, 0.00f }, { { 0.95f, 0.75f, 0.30f }, 0.30f, 1.0f } },
    { { 1.5f, 0.6f, -1.5f }, 0.6f, { 0.00f, 0.03f, -0.04f }, { { 0.25f, 0.35f, 0.85f }, 1.00f, 0.0f } },
    { { 5.0f, 2.0f, -8.0f }, 2.0f, { 0.00f, 0.0f, 0.00f }, { { 0.60f, 0.60f, 0.65f }, 0.60f, 0.0f } },
};
const Box kBoxes[ 2 ] = {
    { { -6.0f, 0.0f, -9.0f }, { -4.0f, 3.0f, -7.0f }, { { 0.70f, 0.55f, 0.40f }, 1.00f, 0.0f } },
    { { 1.0f, 0.0f, -10.0f }, { 2.5f, 1.5f, -8.5f }, { { 0.40f, 0.40f, 0.45f }, 0.30f, 0.0f } },
};
const Material kGround = { { 0.55f, 0.55f, 0.50f }, 0.60f, 0.0f };
const Material kWall   = { { 0.35f, 0.45f, 0.60f }, 1.00f, 0.0f };
const float    kWallZ  = -12.0f;
const V3       kLightPos = { 2.0f, 6.0f, 1.0f };
const float    kLightIntensity = 60.0f;
const V3       kSkyColor = { 0.45f, 0.60f, 0.90f };
const uint32_t kSwitchFrame = 24;

struct Camera { V3 pos, r, u, f; }; // right, up, forward
Camera cameraAt( int frame )
{
    float  ang = ( 0.5f * PI_F / 180.0f ) * ( float )frame;
    float  rad = 9.0f - 0.02f * ( float )frame;
    V3     target = { 0.0f, 1.0f, -5.0f };
    Camera c;
    c.pos = { target.x + rad * sinf( ang ), 2.5f, target.z + rad * cosf( ang ) };
    c.f   = normalize( target - c.pos );
    c.r   = normalize( cross( c.f, { 0, 1, 0 } ) );
    c.u   = cross( c.r, c.f );
    return c;
}
void viewMatrix( const Camera& c, float* m ) // column-major, world -> view (camera looks down -z)
{
    V3 rows[ 3 ] = { c.r, c.u, c.f * -1.0f };
    for( int r = 0; r < 3; r++ )
    {
        m[ 0 * 4 + r ] = rows[ r ].x; m[ 1 * 4 + r ] = rows[ r ].y; m[ 2 * 4 + r ] = rows[ r ].z;
        m[ 3 * 4 + r ] = -dot( rows[ r ], c.pos );
    }
    m[ 3 ] = m[ 7 ] = m[ 11 ] = 0.0f; m[ 15 ] = 1.0f;
}
void invViewMatrix( const Camera& c, float* m )
{
    V3 cols[ 3 ] = { c.r, c.u, c.f * -1.0f };
    for( int k = 0; k < 3; k++ ) { m[ k * 4 + 0 ] = cols[ k ].x; m[ k * 4 + 1 ] = cols[ k ].y; m[ k * 4 + 2 ] = cols[ k ].z; m[ k * 4 + 3 ] = 0; }
    m[ 12 ] = c.pos.x; m[ 13 ] = c.pos.y; m[ 14 ] = c.pos.z; m[ 15 ] = 1.0f;
}
const float kFovY = 75.0f * PI_F / 180.0f, kNear = 0.1f, kFar = 1000.0f;
void projectionMatrix( float aspect, float* m ) // Matrix.cpp:380-395
{
    float t = tanf( kFovY * 0.5f );
    std::memset( m, 0, 64 );
    m[ 0 ] = 1.0f / ( aspect * t ); m[ 5 ] = -1.0f / t; m[ 10 ] = kFar / ( kNear - kFar ); m[ 11 ] = -1.0f; m[ 14 ] = ( kFar * kNear ) / ( kNear - kFar );
}
void invProjectionMatrix( float aspect, float* m )
{
    float t = tanf( kFovY * 0.5f ), a = 1.0f / ( aspect * t ), b = -1.0f / t, c = kFar / ( kNear - kFar ), d = ( kFar * kNear ) / ( kNear - kFar );
    std::memset( m, 0, 64 );
    m[ 0 ] = 1.0f / a; m[ 5 ] = 1.0f / b; m[ 11 ] = 1.0f / d; m[ 14 ] = -1.0f; m[ 15 ] = c / d;
}
struct Proj { float a, b, c, d; }; // x' = a x, y' = b y, z' = c z + d, w' = -z
inline V3 toView( const Camera& c, V3 p ) { V3 q = p - c.pos; return { dot( c.r, q ), dot( c.u, q ), -dot( c.f, q ) }; }
inline void toScreen( const Proj& pr, V3 v, float& sx, float& sy ) // HitInfo.inl:286-290
{
    float w = -v.z;
    sx = ( pr.a * v.x / w ) * 0.5f + 0.5f;
    sy = ( pr.b * v.y / w ) * 0.5f + 0.5f;
}

struct Hit { bool sky; V3 p, n, vel; Material m; };
Hit trace( V3 o, V3 d, int frame )
{
    Hit   h; h.sky = true; h.vel = { 0, 0, 0 };
    float best = 1e30f;
    if( d.y < -1e-6f ) // ground disc of radius 60
    {
        float t = -o.y / d.y;
        V3    q = o + d * t;
        if( t > 1e-3f && t < best && q.x * q.x + q.z * q.z < 3600.0f ) { best = t; h.sky = false; h.n = { 0, 1, 0 }; h.m = kGround; h.vel = { 0, 0, 0 }; }
    }
    if( d.z < -1e-6f ) // finite back wall: sky is visible above and beside it
    {
        float t = ( kWallZ - o.z ) / d.z;
        V3    q = o + d * t;
        if( t > 1e-3f && t < best && fabsf( q.x ) < 14.0f && q.y < 5.0f ) { best = t; h.sky = false; h.n = { 0, 0, 1 }; h.m = kWall; h.vel = { 0, 0, 0 }; }
    }
    for( const Sphere& s : kSpheres )
    {
        V3    c  = s.c + s.vel * ( float )frame;
        V3    oc = o - c;
        float b = dot( oc, d ), cc = dot( oc, oc ) - s.r * s.r, disc = b * b - cc;
        if( disc <= 0 ) continue;
        float t = -b - sqrtf( disc );
        if( t > 1e-3f && t < best ) { best = t; h.sky = false; h.n = normalize( o + d * t - c ); h.m = s.m; h.vel = s.vel; }
    }
    for( const Box& bx : kBoxes )
    {
        float t0 = 0, t1 = 1e30f; int axis = 0; float sgn = 1; bool ok = true;
        const float oo[ 3 ] = { o.x, o.y, o.z }, dd[ 3 ] = { d.x, d.y, d.z }, lo[ 3 ] = { bx.lo.x, bx.lo.y, bx.lo.z }, hi[ 3 ] = { bx.hi.x, bx.hi.y, bx.hi.z };
        for( int a = 0; a < 3 && ok; a++ )
        {
            if( fabsf( dd[ a ] ) < 1e-8f ) { if( oo[ a ] < lo[ a ] || oo[ a ] > hi[ a ] ) ok = false; continue; }
            float inv = 1.0f / dd[ a ], ta = ( lo[ a ] - oo[ a ] ) * inv, tb = ( hi[ a ] - oo[ a ] ) * inv, s = -1;
            if( ta > tb ) { std::swap( ta, tb ); s = 1; }
            if( ta > t0 ) { t0 = ta; axis = a; sgn = s; }
            t1 = std::min( t1, tb );
            if( t0 > t1 ) ok = false;
        }
        if( ok && t0 > 1e-3f && t0 < best ) { best = t0; h.sky = false; h.n = { 0, 0, 0 }; ( axis == 0 ? h.n.x : axis == 1 ? h.n.y : h.n.z ) = sgn; h.m = bx.m; h.vel = { 0, 0, 0 }; }
    }
    if( !h.sky ) h.p = o + d * best;
    return h;
}

inline float lightMultiplier( V3 p, uint32_t frame ) // the light switch that the gradient pass must detect
{
    if( frame < kSwitchFrame ) return 1.0f;
    if( p.x > -2.5f && p.x < 0.5f && p.z > -4.0f && p.z < -1.0f ) return 4.0f;
    if( p.x > 1.0f && p.x < 4.0f && p.z > -7.0f && p.z < -4.5f ) return 0.0f;
    return 1.0f;
}

struct Shaded { bool sky; V3 diff, spec, indir; float hitDist; Hit h; V3 dir; };

inline void ckb( int x, int y, int W, int& cx, int& cy ) { int odd = ( x + y % 2 ) % 2; cx = odd * ( W / 2 ) + x / 2; cy = y; } // ShaderCommonGLSLFunc.h:321-329
inline void reg( int cx, int cy, int W, int& x, int& y ) { int sep = W / 2, odd = cx >= sep; x = ( cx - odd * sep ) * 2 + ( odd + cy ) % 2; y = cy; }   // :331-342

V3 rayDir( const Camera& c, int x, int y, int W, int H, float dx, float dy )
{
    float t = tanf( kFovY * 0.5f ), aspect = ( float )W / ( float )H;
    float nx = ( ( x + 0.5f + dx ) / W ) * 2.0f - 1.0f, ny = ( ( y + 0.5f + dy ) / H ) * 2.0f - 1.0f;
    V3    dc = { nx * aspect * t, -ny * t, -1.0f };
    return normalize( c.r * dc.x + c.u * dc.y + c.f * 1.0f * ( -dc.z ) );
}

// radiance of checkerboard pixel (cx,cy) at `frame` with the noise of (cx,cy,frame), lit with the lighting of `lightFrame`
Shaded shade( int cx, int cy, int W, int H, uint32_t frame, uint32_t lightFrame, uint32_t seed )
{
    Shaded s;
    int x, y; reg( cx, cy, W, x, y );
    Camera cam = cameraAt( ( int )frame );
    s.dir = rayDir( cam, x, y, W, H, 0, 0 );
    s.h   = trace( cam.pos, s.dir, ( int )frame );
    s.sky = s.h.sky;
    s.diff = s.spec = s.indir = { 0, 0, 0 };
    s.hitDist = 0.0f;
    if( s.sky ) return s;
    const Hit& h = s.h;
    V3    L = kLightPos - h.p; float d2 = dot( L, L ); L = L * ( 1.0f / sqrtf( d2 ) );
    float ndl = std::max( 0.0f, dot( h.n, L ) );
    float e   = kLightIntensity * lightMultiplier( h.p, lightFrame ) / d2;
    V3    Hh  = normalize( L - s.dir );
    float a2  = std::max( 1e-3f, h.m.roughness * h.m.roughness );
    float sp  = powf( std::max( 0.0f, dot( h.n, Hh ) ), std::min( 2.0f / ( a2 * a2 ) - 2.0f + 1e-3f, 2000.0f ) ) * ( 1.0f / ( 4.0f * a2 + 0.1f ) );
    auto noise = [ & ]( uint32_t ch ) {
        float u1 = u01( hash5( cx, cy, frame, ch * 2 + 0, seed ) ), u2 = u01( hash5( cx, cy, frame, ch * 2 + 1, seed ) );
        float g  = sqrtf( -2.0f * logf( u1 ) ) * cosf( 2.0f * PI_F * u2 );
        float n  = expf( 1.0f * g - 0.5f );
        if( ( hash5( cx, cy, frame, 16 + ch, seed ) % 1000u ) == 0u ) n *= 50.0f; // fireflies
        return n;
    };
    float nd = noise( 0 ), ns = noise( 1 ), ni = noise( 2 );
    s.diff  = V3{ 1.0f, 0.97f, 0.92f } * ( e * ndl * nd );
    s.spec  = V3{ 1.0f, 0.97f, 0.92f } * ( e * ndl * sp * ns ) + V3{ 0.30f, 0.40f, 0.55f } * ( 0.15f * ns );
    s.indir = V3{ 0.30f, 0.33f, 0.40f } * ( ( 0.35f + 0.25f * h.n.y ) * ni );
    s.hitDist = 0.5f + 19.5f * u01( hash5( cx, cy, frame, 40, seed ) );
    return s;
}

struct Out
{
    uint8_t* p[ ASVGF_IMAGE_COUNT ];
};
template< typename T > inline void put( uint8_t* base, size_t idx, const T& v ) { if( base ) std::memcpy( base + idx * sizeof( T ), &v, sizeof( T ) ); }
struct H2 { uint16_t a, b; };
struct H4 { uint16_t a, b, c, d; };
struct F4 { float a, b, c, d; };
struct B2 { uint8_t a, b; };

} // namespace

extern "C" {

// Camera + per-frame parameters exactly as VulkanDevice::FillUniform would leave them (VulkanDevice.cpp:142-482).
void asvgfSynthFillUniform( AsvgfGlobalUniform* u, uint32_t W, uint32_t H, uint32_t frame )
{
    std::memset( u, 0, sizeof( *u ) );
    Camera cur = cameraAt( ( int )frame ), prev = cameraAt( ( int )frame - 1 );
    float  aspect = ( float )W / ( float )H;
    viewMatrix( cur, u->view ); invViewMatrix( cur, u->invView ); viewMatrix( prev, u->viewPrev );
    projectionMatrix( aspect, u->projection ); projectionMatrix( aspect, u->projectionPrev ); invProjectionMatrix( aspect, u->invProjection );
    u->renderWidth = ( float )W; u->renderHeight = ( float )H;
    u->frameId = frame; u->timeDelta = 1.0f / 60.0f;
    u->cameraPosition[ 0 ] = cur.pos.x; u->cameraPosition[ 1 ] = cur.pos.y; u->cameraPosition[ 2 ] = cur.pos.z; u->cameraPosition[ 3 ] = 1.0f;
    u->cameraPositionPrev[ 0 ] = prev.pos.x; u->cameraPositionPrev[ 1 ] = prev.pos.y; u->cameraPositionPrev[ 2 ] = prev.pos.z; u->cameraPositionPrev[ 3 ] = 1.0f;
    u->gradientMultDiffuse = 0.5f; u->gradientMultIndirect = 0.2f; u->gradientMultSpecular = 0.5f; // RTGL1.h:680-686 defaults
    u->antiFireflyEnabled = 1; u->lightmapScreenCoverage = 0.0f; u->debugShowFlags = 0;
    // exposure range as VulkanDevice.cpp:205-210 derives it from ev100Min = 0, ev100Max = 10: log2( 2^ev * 12.5 / 100 )
    u->minLogLuminance = -3.0f; u->maxLogLuminance = 7.0f; u->luminanceWhitePoint = 10.0f; u->stopEyeAdaptation = 0;
}

// images[i] = tightly packed host buffer for image index i (asvgf_layout.h) or NULL to skip.  Fills the INPUTS of
// frame `frame`: Albedo, IsSky, Normal, MetallicRoughness, DepthWorld, DepthGrad, Motion, Unfiltered{Direct,Specular,
// Indir}, SurfacePosition, ViewDirection, Throughput, GradientInputs, FwdLum.
// Rows [y0,y1) only: images[i] then holds just those rows (a band plus its halo, for the multi-GPU band mode).
void asvgfSynthGenerateRows( uint32_t W, uint32_t H, uint32_t frame, uint32_t seed, void** images, uint32_t y0, uint32_t y1 )
{
    Out o; for( int i = 0; i < ASVGF_IMAGE_COUNT; i++ ) o.p[ i ] = ( uint8_t* )images[ i ];
    const Camera cur = cameraAt( ( int )frame ), prev = cameraAt( ( int )frame - 1 );
    const float  aspect = ( float )W / ( float )H, t = tanf( kFovY * 0.5f );
    const Proj   pr = { 1.0f / ( aspect * t ), -1.0f / t, kFar / ( kNear - kFar ), ( kFar * kNear ) / ( kNear - kFar ) };
    const int    iW = ( int )W, iH = ( int )H;

#pragma omp parallel for schedule( dynamic, 4 )
    for( int cy = ( int )y0; cy < ( int )y1; cy++ )
    {
        for( int cx = 0; cx < iW; cx++ )
        {
            const size_t idx = ( size_t )( cy - ( int )y0 ) * W + cx;
            int x, y; reg( cx, cy, iW, x, y );
            const size_t ridx = ( size_t )( y - ( int )y0 ) * W + x;
            Shaded s = shade( cx, cy, iW, iH, frame, frame, seed );
            const Hit& h = s.h;

            put( o.p[ ASVGF_IMAGE_THROUGHPUT ], idx, H4{ toHalf( 1 ), toHalf( 1 ), toHalf( 1 ), 0 } );
            if( s.sky )
            {
                // RaygenPrimary.inl:98-123
                float sxc, syc, sxp, syp;
                V3 far = cur.pos + s.dir * 1.0e6f;
                toScreen( pr, toView( cur, far ), sxc, syc ); toScreen( pr, toView( prev, far ), sxp, syp );
                put( o.p[ ASVGF_IMAGE_IS_SKY ], idx, ( uint8_t )1 );
                put( o.p[ ASVGF_IMAGE_ALBEDO ], ridx, packUf( kSkyColor ) );
                put( o.p[ ASVGF_IMAGE_NORMAL ], idx, 0x80000000u ); // encodeNormal(vec3(0)): phi = pi/2, theta = 0
                put( o.p[ ASVGF_IMAGE_METALLIC_ROUGHNESS ], idx, B2{ 0, 0 } );
                put( o.p[ ASVGF_IMAGE_DEPTH_WORLD ], idx, toHalf( 2.0f * ASVGF_MAX_RAY_LENGTH ) );
                put( o.p[ ASVGF_IMAGE_DEPTH_GRAD ], idx, toHalf( 0.0f ) );
                put( o.p[ ASVGF_IMAGE_MOTION ], idx, H4{ toHalf( sxp - sxc ), toHalf( syp - syc ), 0, 0 } );
                put( o.p[ ASVGF_IMAGE_SURFACE_POSITION ], idx, F4{ 1.0e7f, 1.0e7f, 1.0e7f, 1.0e7f } );
                put( o.p[ ASVGF_IMAGE_VIEW_DIRECTION ], idx, H4{ toHalf( s.dir.x ), toHalf( s.dir.y ), toHalf( s.dir.z ), 0 } );
                put( o.p[ ASVGF_IMAGE_UNFILTERED_DIRECT ], idx, 0u );
                put( o.p[ ASVGF_IMAGE_UNFILTERED_SPECULAR ], idx, 0u );
                put( o.p[ ASVGF_IMAGE_UNFILTERED_INDIR ], idx, 0u );
                put( o.p[ ASVGF_IMAGE_GRADIENT_INPUTS ], idx, H2{ 0, 0 } );
            }
            else
            {
                // HitInfo.inl:271-336
                V3    vc = toView( cur, h.p ), vp = toView( prev, h.p - h.vel );
                float sxc, syc, sxp, syp;
                toScreen( pr, vc, sxc, syc ); toScreen( pr, vp, sxp, syp );
                float depthLinear = length( vc );
                float clipZ = pr.c * vc.z + pr.d;
                float gz[ 2 ];
                for( int a = 0; a < 2; a++ ) // ray differentials on the tangent plane
                {
                    V3    d2  = rayDir( cur, x, y, iW, iH, a == 0 ? 1.0f : 0.0f, a == 1 ? 1.0f : 0.0f );
                    float den = dot( d2, h.n );
                    float tt  = fabsf( den ) > 1e-6f ? dot( h.p - cur.pos, h.n ) / den : length( h.p - cur.pos );
                    V3    va  = toView( cur, cur.pos + d2 * tt );
                    gz[ a ]   = ( pr.c * va.z + pr.d ) - clipZ;
                }
                put( o.p[ ASVGF_IMAGE_IS_SKY ], idx, ( uint8_t )0 );
                put( o.p[ ASVGF_IMAGE_ALBEDO ], ridx, packUf( h.m.albedo ) );
                put( o.p[ ASVGF_IMAGE_NORMAL ], idx, packNormal( h.n ) );
                put( o.p[ ASVGF_IMAGE_METALLIC_ROUGHNESS ], idx, B2{ ( uint8_t )rintf( h.m.metallic * 255.0f ), ( uint8_t )rintf( h.m.roughness * 255.0f ) } );
                put( o.p[ ASVGF_IMAGE_DEPTH_WORLD ], idx, toHalf( depthLinear ) );
                put( o.p[ ASVGF_IMAGE_DEPTH_GRAD ], idx, toHalf( sqrtf( gz[ 0 ] * gz[ 0 ] + gz[ 1 ] * gz[ 1 ] ) ) );
                put( o.p[ ASVGF_IMAGE_MOTION ], idx, H4{ toHalf( sxp - sxc ), toHalf( syp - syc ), toHalf( length( vp ) - depthLinear ), 0 } );
                put( o.p[ ASVGF_IMAGE_SURFACE_POSITION ], idx, F4{ h.p.x, h.p.y, h.p.z, 0.0f } );
                put( o.p[ ASVGF_IMAGE_VIEW_DIRECTION ], idx, H4{ toHalf( s.dir.x ), toHalf( s.dir.y ), toHalf( s.dir.z ), toHalf( s.hitDist ) } );
                put( o.p[ ASVGF_IMAGE_UNFILTERED_DIRECT ], idx, packE5( s.diff ) );
                put( o.p[ ASVGF_IMAGE_UNFILTERED_SPECULAR ], idx, packE5( s.spec ) );
                put( o.p[ ASVGF_IMAGE_UNFILTERED_INDIR ], idx, packE5( s.indir ) );
                put( o.p[ ASVGF_IMAGE_GRADIENT_INPUTS ], idx, H2{ toHalf( lum( s.diff ) ), toHalf( lum( s.spec ) ) } );
            }

            // FwdLum: the sample that the gradient pass would forward-project into this pixel, re-shaded with the
            // lighting of the current frame (RtGradients.rgen:82,107-136).  prevPix as in RtGradients.rgen:82.
            if( o.p[ ASVGF_IMAGE_FWD_LUM ] )
            {
                H2 fl = { 0, 0 };
                if( !s.sky && frame > 0 )
                {
                    V3    vc = toView( cur, h.p ), vp = toView( prev, h.p - h.vel );
                    float sxc, syc, sxp, syp;
                    toScreen( pr, vc, sxc, syc ); toScreen( pr, vp, sxp, syp );
                    float mx = fromHalf( toHalf( sxp - sxc ) ), my = fromHalf( toHalf( syp - syc ) );
                    float ssx = ( float )W / 2.0f, ssy = ( float )H;
                    int   ppx = ( int )floorf( ( ( cx + 0.5f ) * ( 1.0f / ssx ) + mx ) * ssx ), ppy = ( int )floorf( ( ( cy + 0.5f ) * ( 1.0f / ssy ) + my ) * ssy );
                    if( ppx >= 0 && ppy >= 0 && ppx < iW && ppy < iH )
                    {
                        Shaded ps = shade( ppx, ppy, iW, iH, frame - 1, frame, seed );
                        if( !ps.sky ) fl = { toHalf( lum( ps.diff ) ), toHalf( lum( ps.spec ) ) };
                    }
                }
                put( o.p[ ASVGF_IMAGE_FWD_LUM ], idx, fl );
            }
        }
    }
}

void asvgfSynthGenerate( uint32_t W, uint32_t H, uint32_t frame, uint32_t seed, void** images )
{
    asvgfSynthGenerateRows( W, H, frame, seed, images, 0, H );
}

} // extern "C"
