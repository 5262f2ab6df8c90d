// asvgf_atrous_tiled.cuh -- what the a-trous kernel (asvgf_atrous_half.cuh: CmSVGFAtrous_Iter0.comp for iteration 0, CmSVGFAtrous.comp
// for iterations >= 1 and the composition) is built from, sm_100a: the geometry of the dilated tile and its shared-memory layout,
// the per-frame guide plane, the state of a pixel while its taps are evaluated, and the stores / composition of a finished pixel.
//
// Arithmetic: CmSVGFAtrous_Iter0.comp:158-335 / CmSVGFAtrous.comp:65-349.  Structure:
//
//  * GUIDE PLANE.  Normal, depth and roughness do not change between the iterations of a frame, yet every iteration used to
//    re-decode them per staged texel (two dependent sin/cos-table gathers per normal).  They are decoded ONCE per frame -- by the
//    temporal pass, which has them in registers anyway (k_guide only when that pass did not cover the rows) -- into a float4
//    plane (n.xyz, depth); roughness rides in the low byte of the depth word (an fp16 depth leaves the low 13 mantissa bits of
//    its fp32 form zero) and is split off while staging.  A tile's geometry costs one 16-byte load per texel.
//  * CENTRE PLANE.  The four weight constants of a pixel that do not depend on the iteration (luminance scale, the two depth
//    scales, the specular power) are derived by iteration 0 from three more images and published; iterations >= 1 read them.
//  * SKY TILES.  A tile without a single valid surface passes its centre texels through (the shader's early return) and
//    skips decode and taps.
//  * DILATED TILE.  For stride s a pixel only touches rows y, y+-s: a CTA owns 64 columns of 8 rows spaced s apart (the s
//    residues go to s different CTAs); halo = one dilated row above/below, max(s,4) columns left/right.  Read amplification
//    1.4-1.56 at every stride, every global access a full-line row segment.
//  * VECTOR STAGING (strides >= 4).  A thread stages four consecutive texels: 4+2+1+1 16-byte loads, one address computation,
//    decode, stores into quarter-row slots (conflict-free for the staging stores and for the tap loads, see below).
//  * WEIGHTS THROUGH ONE LOG.  w_n = (n.n')^128 and the specular lobe (n.n')^P share L = lg2(n.n'); the wavelet weight and the
//    render-area test are folded into the exponent (lk = log2 k, or -inf outside):
//      wBase = ex2(128 L + lk - wz^2 log2e),  wDiff = ex2(128 L + lk - (wz^2 + wl) log2e),  wSpec = r * ex2(. + P L)
//    1 lg2 + 3 ex2 per tap.  Tiles whose taps all stay inside the render area (all but the frame / checkerboard-separator
//    borders) take a copy of the tap loop compiled with immediate lk; the choice is per CTA.
//  * 64 registers: four CTAs = 32 warps per SM.
// (The kernel that kept the tile as four fp32 planes, and its two-pixel form, are in experiments/asvgf_atrous_fp32.cuh.)
#pragma once

#include <type_traits>

#include "asvgf_fast.cuh"

namespace asvgf
{

constexpr int kAtTileW   = 64;  // columns per CTA
constexpr int kAtTileH   = 8;   // dilated rows per CTA
constexpr int kAtThreads = 256; // 2 pixels per thread
#ifndef ASVGF_AT_MIN_CTAS
    #define ASVGF_AT_MIN_CTAS 4
#endif

constexpr int    atrousHalo( int step ) { return step < 4 ? 4 : step; } // halo columns: multiples of 4 keep 16-byte row segments
// Shared-memory layout of one staged row: QUARTER-ROW SLOTS.  Column c lives in slot (c & 3) * GP + (c >> 2).  The staging thread
// that holds columns 4g..4g+3 stores them to slots g, GP + g, 2 GP + g, 3 GP + g: consecutive lanes hit consecutive slots, and
// a quarter warp of the tap loop (eight consecutive columns) hits slots {0, GP, 2GP, 3GP, 1, GP+1, ...} + const, which are
// distinct modulo 8 (= distinct 16-byte bank groups) when GP = 2 or 6 (mod 8).  Both sides of the tile are conflict-free.
// (Storing four consecutive texels per lane into a plain row layout was measured: 4-way conflicts on every STS.128, 18 M
// conflict cycles per launch, the kernel 25 % slower in spite of 23 % fewer instructions.)
constexpr int    atrousGroupPitch( int step ) { int g = ( kAtTileW + 2 * atrousHalo( step ) ) / 4; while( g % 8 != 2 && g % 8 != 6 ) g++; return g; }
constexpr size_t atrousTiledSmemBytes( int step ) { return ( size_t )( 4 * atrousGroupPitch( step ) ) * ( kAtTileH + 2 ) * 64; }

ASVGF_DEVICE float ex2Approx( float x ) { return fex2( x ); }

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

// ------------------------------------------------------------------------------------------------------------
// guide plane: (n.x, n.y, n.z, depth | roughness byte) of every stored pixel of the current frame
// ------------------------------------------------------------------------------------------------------------
struct GuideArgs
{
    Img       normal, depth, metRough, guide;
    NormalLut lut;
    int32_t   yBegin, yEnd; // stored rows
};

ASVGF_DEVICE float4 guideTexel( const NormalLut& lut, uint32_t encNormal, uint16_t depthHalf, uint16_t metRough )
{
    const float3   n = decodeNormal( lut, encNormal );
    const uint32_t w = __float_as_uint( h2f( depthHalf ) ) | ( ( uint32_t )metRough >> 8 ); // fp16 -> fp32 leaves the low 13 bits zero
    return make_float4( n.x, n.y, n.z, __uint_as_float( w ) );
}
ASVGF_DEVICE float guideDepth( float w ) { return __uint_as_float( __float_as_uint( w ) & 0xffffff00u ); }
ASVGF_DEVICE float guideRoughByte( float w ) { return ( float )( __float_as_uint( w ) & 0xffu ); } // 0..255

__global__ void __launch_bounds__( 256 ) k_guide( const __grid_constant__ GuideArgs a )
{
    // one thread = four consecutive pixels of a row (W is even; a partial group at the right edge is handled per pair)
    pdlLaunchDependents();
    const int x0 = ( blockIdx.x * 64 + ( threadIdx.x & 63 ) ) * 4;
    const int y  = a.yBegin + blockIdx.y * 4 + ( threadIdx.x >> 6 );
    if( x0 >= a.guide.w || y >= a.yEnd ) return;
    pdlWait();
    const bool full = x0 + 3 < a.guide.w;
    uint32_t   n[ 4 ];
    uint16_t   d[ 4 ], mr[ 4 ];
    if( full )
    {
        const uint4 nn = __ldg( reinterpret_cast< const uint4* >( at< uint32_t >( a.normal, x0, y ) ) );
        const uint2 dd = __ldg( reinterpret_cast< const uint2* >( at< uint16_t >( a.depth, x0, y ) ) );
        const uint2 rr = __ldg( reinterpret_cast< const uint2* >( at< uint16_t >( a.metRough, x0, y ) ) );
        n[ 0 ] = nn.x; n[ 1 ] = nn.y; n[ 2 ] = nn.z; n[ 3 ] = nn.w;
        d[ 0 ] = ( uint16_t )dd.x; d[ 1 ] = ( uint16_t )( dd.x >> 16 ); d[ 2 ] = ( uint16_t )dd.y; d[ 3 ] = ( uint16_t )( dd.y >> 16 );
        mr[ 0 ] = ( uint16_t )rr.x; mr[ 1 ] = ( uint16_t )( rr.x >> 16 ); mr[ 2 ] = ( uint16_t )rr.y; mr[ 3 ] = ( uint16_t )( rr.y >> 16 );
    }
    else
    {
#pragma unroll
        for( int k = 0; k < 4; k++ )
        {
            const bool ok = x0 + k < a.guide.w;
            n[ k ]  = ldIf< uint32_t >( a.normal, x0 + k, y, ok, 0u );
            d[ k ]  = ldIf< uint16_t >( a.depth, x0 + k, y, ok, ( uint16_t )0 );
            mr[ k ] = ldIf< uint16_t >( a.metRough, x0 + k, y, ok, ( uint16_t )0 );
        }
    }
#pragma unroll
    for( int k = 0; k < 4; k++ )
        if( x0 + k < a.guide.w ) *atW< float4 >( a.guide, x0 + k, y ) = guideTexel( a.lut, n[ k ], d[ k ], mr[ k ] );
}

// ------------------------------------------------------------------------------------------------------------
// the filter
// ------------------------------------------------------------------------------------------------------------
// Geometry of the staged tile.  STEP >= 4: four-texel vector staging into quarter-row slots (above).  STEP < 4: the side taps
// sit at columns c +- 1 / c +- 2, whose eight-lane groups are not 4-aligned -- the slot layout would serve them with 2-way
// conflicts (measured: 19 M conflict cycles) -- so those strides stage one texel per lane into a plain row layout, which is
// conflict-free for stores and for loads at any alignment.
template< int STEP > struct AtrousGeom
{
    static constexpr bool VEC    = STEP >= 4;
    static constexpr int  HALO   = VEC ? atrousHalo( STEP ) : STEP;
    static constexpr int  TW     = kAtTileW + 2 * HALO; // staged columns
    static constexpr int  TROWS  = kAtTileH + 2;
    static constexpr int  GROUPS = TW / 4;
    static constexpr int  GP     = VEC ? atrousGroupPitch( STEP ) : 0;
    static constexpr int  RP     = VEC ? 4 * GP : TW;   // slots per staged row
    static constexpr int  NPX    = RP * TROWS;
    // slot of staged column `col` (0 .. TW-1) of staged row `row`
    static __device__ __forceinline__ int slot( int row, int col ) { return VEC ? row * RP + ( col & 3 ) * GP + ( col >> 2 ) : row * RP + col; }
    static constexpr int dCol = VEC ? STEP / 4 : STEP;  // slot distance of the tap columns c +- STEP
};
constexpr size_t atrousSmemBytes( int step )
{
    return step >= 4 ? atrousTiledSmemBytes( step ) : ( size_t )( kAtTileW + 2 * step ) * ( kAtTileH + 2 ) * 64;
}

// state of one output pixel while its taps are evaluated
struct AtrousPix
{
    float3 n;
    float  depth, l, roughByte;
    float  kLum, wRough, wR10, P, invK1, invK2;
    float3 diff, spec, indir;
    float  variance, wSum, wSumSpec, wSumIndir;
    bool   invalid;

    // centre texel planes + the centre-only inputs (CmSVGFAtrous.comp:86-118, CmSVGFAtrous_Iter0.comp:201-234)
    ASVGF_DEVICE void init( const float4 cA, const float4 cB, const float4 cC, const float4 cD, float gradDepth, float sigmaL, float historyLengthSpec, int iteration )
    {
        constexpr float LOG2E = 1.4426950408889634f;
        n = make_float3( cA.x, cA.y, cA.z ); depth = cA.w; l = cC.w; roughByte = cD.w;
        diff = make_float3( cB.x, cB.y, cB.z ); spec = make_float3( cC.x, cC.y, cC.z ); indir = make_float3( cD.x, cD.y, cD.z );
        variance = cB.w; wSum = wSumSpec = wSumIndir = 1.0f;
        invalid  = depth < 0.0f || depth > 10000.0f;
        const float roughness = roughByte * ( 1.0f / 255.0f );
        kLum   = LOG2E * frcp( fmaf( 4.0f, sigmaL, 0.00001f ) );
        // w_r = max(0, 1 - 10 |dr|) * wRough with r = byte / 255: the byte difference times wRough * 10/255
        wRough = saturatef( fmaf( roughness, 30.0f, -( float )iteration ) );
        wR10   = wRough * ( 10.0f / 255.0f );
        const float alpha = roughness * roughness;
        P = fmaxf( 0.01f, fmaf( 2.0f, frcp( fmaf( alpha, alpha, 1e-4f ) ), -2.0f ) ); // BRDF.h:27-29
        P = clampf( P, 8.0f, 1024.0f ) * saturatef( historyLengthSpec * 0.125f );
        // 1/max(g,.01) = min(1/g,100) and 1/max(2g,.01) = min(.5/g,100) from one MUFU (g >= 0; 1/0 = inf), pre-scaled by
        // sqrt(log2 e) so that -wZ*wZ is already the base-2 exponent
        constexpr float SQRT_LOG2E = 1.2011224087864498f;
        const float rgd = frcp( gradDepth );
        invK1 = SQRT_LOG2E * fminf( rgd, 100.0f ); invK2 = SQRT_LOG2E * fminf( 0.5f * rgd, 100.0f );
    }
    // iterations >= 1 with the centre plane of iteration 0: only the roughness weight depends on the iteration
    ASVGF_DEVICE void initFromCentre( const float4 cA, const float4 cB, const float4 cC, const float4 cD, const float4 k, int iteration )
    {
        n = make_float3( cA.x, cA.y, cA.z ); depth = cA.w; l = cC.w; roughByte = cD.w;
        diff = make_float3( cB.x, cB.y, cB.z ); spec = make_float3( cC.x, cC.y, cC.z ); indir = make_float3( cD.x, cD.y, cD.z );
        variance = cB.w; wSum = wSumSpec = wSumIndir = 1.0f;
        invalid  = depth < 0.0f || depth > 10000.0f;
        wRough = saturatef( fmaf( roughByte, 30.0f / 255.0f, -( float )iteration ) );
        wR10   = wRough * ( 10.0f / 255.0f );
        kLum = k.x; invK1 = k.y; invK2 = k.z; P = k.w;
    }
    ASVGF_DEVICE void normalise()
    {
        const float inv = frcp( wSum ), invSpec = frcp( wSumSpec ), invIndir = frcp( wSumIndir );
        diff     = make_float3( diff.x * inv, diff.y * inv, diff.z * inv );
        variance = variance * ( inv * inv );
        spec     = make_float3( spec.x * invSpec, spec.y * invSpec, spec.z * invSpec );
        indir    = make_float3( indir.x * invIndir, indir.y * invIndir, indir.z * invIndir );
    }
};

// stores of one pixel: the three colour images, or (LAST) the composition of CmSVGFAtrous.comp:278-348
template< bool LAST >
ASVGF_DEVICE void atrousStore( const AtrousArgs& a, const Uniform& u, int px, int py, const float3 outDiff, const float outVariance, const float3 outSpec, const float3 outIndir )
{
    if( !LAST )
    {
        const __half2 o0 = __floats2half2_rn( outDiff.x, outDiff.y ), o1 = __floats2half2_rn( outDiff.z, outVariance );
        *atW< uint2 >( a.dstDiff, px, py ) = make_uint2( *reinterpret_cast< const uint32_t* >( &o0 ), *reinterpret_cast< const uint32_t* >( &o1 ) );
        *atW< uint32_t >( a.dstSpec, px, py )  = encodeE5MagicClamped( outSpec );
        *atW< uint32_t >( a.dstIndir, px, py ) = encodeE5MagicClamped( outIndir );
        return;
    }
    // (the extra centre-only inputs are loaded here, one pixel at a time: the registers of the tap loop stay free)
    const int      sep   = u.W / 2;
    const int      isOdd = px >= sep ? 1 : 0;
    const int      regX  = ( px - isOdd * sep ) * 2 + ( ( isOdd + py ) & 1 ); // ShaderCommonGLSLFunc.h:331-342
    const uint32_t albedoBits = __ldg( at< uint32_t >( a.albedo, regX, py ) );
    const uint2    tp     = __ldg( at< uint2 >( a.throughput, px, py ) );
    const uint32_t sky    = __ldg( at< uint8_t >( a.isSky, px, py ) );
    const uint32_t metBits = __ldg( at< uint16_t >( a.metRough, px, py ) );
    float3        albedo = unpackUfloat111110Fast( albedoBits );
    const float2  tp01   = __half22float2( *reinterpret_cast< const __half2* >( &tp.x ) );
    const float3  throughput = make_float3( tp01.x, tp01.y, __low2float( *reinterpret_cast< const __half2* >( &tp.y ) ) );
    if( ( u.debugShowFlags & ( 1u << 9 ) ) != 0 ) albedo = make_float3( 1.0f, 1.0f, 1.0f );

    float3 illuminated;
    if( sky == 0u )
    {
        const float metallic = unorm8ToFloatFast( metBits & 0xffu );
        const float rough    = unorm8ToFloatFast( metBits >> 8 );
        const float t  = saturatef( ( rough - 0.5f ) * 4.0f );            // smoothstep(0.5, 0.75, roughness)
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

} // namespace asvgf
