// asvgf_atrous_tiled.cuh -- the a-trous kernel (CmSVGFAtrous_Iter0.comp for iteration 0, CmSVGFAtrous.comp for iterations >= 1 and
// the composition), sm_100a, plus the per-frame guide plane it reads.
//
// Arithmetic: CmSVGFAtrous_Iter0.comp:158-335 / CmSVGFAtrous.comp:65-349.  Structure (round 2; ncu of the round-1 kernel:
// 578 instructions per pixel, issue 73 %, shared-memory pipe 60 %, ALU pipe 41 % -- instruction- and LDS-bound, DRAM 30 %):
//
//  * GUIDE PLANE.  Normal, depth and roughness do not change between the iterations of a frame, yet every iteration used to
//    re-decode them per staged texel (two dependent sin/cos-table gathers per normal).  They are decoded ONCE per frame -- by the
//    temporal pass, which has them in registers anyway (k_guide only when that pass did not cover the rows) -- into a float4
//    plane (n.xyz, depth); roughness rides in the low byte of the depth word (an fp16 depth leaves the low 13 mantissa bits of
//    its fp32 form zero) and is split off while staging.  A tile's geometry costs one 16-byte load per texel.
//  * CENTRE PLANE.  The four weight constants of a pixel that do not depend on the iteration (luminance scale, the two depth
//    scales, the specular power) are derived by iteration 0 from three more images and published; iterations >= 1 read them.
//  * SKY TILES.  A tile without a single valid surface passes its centre texels through (the shader's early return): the CTA
//    votes on the centre depths first and skips staging and taps.
//  * DILATED TILE.  For stride s a pixel only touches rows y, y+-s: a CTA owns 64 columns of 8 rows spaced s apart (the s
//    residues go to s different CTAs); halo = one dilated row above/below, max(s,4) columns left/right.  Read amplification
//    1.4-1.56 at every stride, every global access a full-line row segment.
//  * VECTOR STAGING.  A thread stages four consecutive texels: 4+2+1+1 16-byte loads, one address computation per image,
//    decode, 16 STS.128 into quarter-row slots (conflict-free for the staging stores and for the tap loads, see below).
//    No per-texel index arithmetic, no table gathers.
//  * WEIGHTS THROUGH ONE LOG.  w_n = (n.n')^128 and the specular lobe (n.n')^P share L = lg2(n.n'); the wavelet weight and the
//    render-area test are folded into the exponent (lk = log2 k, or -inf outside):
//      wBase = ex2(128 L + lk - wz^2 log2e),  wDiff = ex2(128 L + lk - (wz^2 + wl) log2e),  wSpec = r * ex2(. + P L)
//    1 lg2 + 3 ex2 and ~30 FMA-pipe instructions per tap.  Tiles whose taps all stay inside the render area (all but the
//    frame / checkerboard-separator borders) take a copy of the tap loop compiled with immediate lk; the choice is per CTA.
//  * 64 registers, 41-51 KB of shared memory: four CTAs = 32 warps per SM.
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
            AtrousPix s;
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
        AtrousPix  s0, s1;
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
