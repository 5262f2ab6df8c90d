// asvgf_atrous_half.cuh -- the a-trous kernel on a PACKED tile (CmSVGFAtrous_Iter0.comp / CmSVGFAtrous.comp), sm_100a.
//
// Same filter, same arithmetic and the same results bit for bit as the kernel that kept the tile as four fp32 planes
// (experiments/asvgf_atrous_fp32.cuh); what changes is what a tap reads from shared memory.  ncu of that kernel: 36 + 4 LDS.128 per
// pixel = 160 shared-memory wavefronts per warp-pixel; four schedulers at ~525 issue slots per warp-pixel ask the 128 B/clk pipe
// for 640 clk of loads per 525 clk of issue -- the tap loop was bound by shared-memory bandwidth (pipe 65-72 % busy, issue
// 53-62 %), and trimming instructions elsewhere did not move it (staging cut from 90 to 64 instructions per texel: 0.3 % of the
// frame).  So the tile shrinks instead:
//
//   * every colour the filter reads is EXACT in fp16 -- the diffuse image is rgba16f, E5B9G9R9 components are 9-bit mantissas
//     times 2^(e-24) with e <= 31 (all representable, subnormals included), depth is r16f, roughness a byte -- so a texel is kept as
//       plane A (16 B): n.x, n.y, n.z (fp32), half2( depth, roughness byte )
//       plane B (16 B): diff.r, diff.g, diff.b, variance (fp32)
//       plane C (16 B): luminance (fp32), half2( spec.r, spec.g ), half2( spec.b, indir.r ), half2( indir.g, indir.b )
//     48 bytes instead of 64: a tap costs 12 wavefronts instead of 16, and the six packed colours are widened in registers
//     (HADD2.F32; the depth and roughness differences take the half operand directly: FHADD, sm_100's mixed-precision add).
//     Widening is exact, the sums and products see the same fp32 values as before: identical output.
//   * -DASVGF_AT_ALL_HALF packs the diffuse colour too (plane B: half2( r, g ), half2( b, variance ), luminance, half2( spec.r,
//     spec.g ); plane C 8 bytes with its own quarter-row pitch: 40 B per texel, 10 wavefronts and 10 widenings per tap).  Measured
//     back to back at 4K: fp32 planes 1.3865 ms per frame, all-half 1.3160, this layout 1.3107 -- once the shared-memory pipe
//     is no longer the limit (38-54 % busy) the widenings cost what the wavefronts save.
//   * the sky vote of iterations >= 1 comes from the staged guide depths (one memory round trip per tile, not two).
#pragma once

#include "asvgf_atrous_tiled.cuh"

namespace asvgf
{

// quarter-row pitch of the 8-byte plane: the 16 slots a half warp touches, q * GP8 + j (q, j = 0..3), must differ modulo 16
constexpr int    atrousGroupPitch8( int step ) { int g = ( kAtTileW + 2 * atrousHalo( step ) ) / 4; while( g % 16 != 4 && g % 16 != 12 ) g++; return g; }
#ifndef ASVGF_AT_ALL_HALF
    #define ASVGF_AT_HYBRID // the default layout (above)
#endif
#ifdef ASVGF_AT_HYBRID
using AtC = uint4;
#else
using AtC = uint2;
#endif
template< int STEP > struct AtrousGeomH : AtrousGeom< STEP >
{
    using G = AtrousGeom< STEP >;
#ifdef ASVGF_AT_HYBRID
    static constexpr int GP8  = G::GP;
#else
    static constexpr int GP8  = G::VEC ? atrousGroupPitch8( STEP ) : 0;
#endif
    static constexpr int RP8  = G::VEC ? 4 * GP8 : G::TW;
    static constexpr int NPX8 = RP8 * G::TROWS;
    static __device__ __forceinline__ int slot8( int row, int col ) { return G::VEC ? row * RP8 + ( col & 3 ) * GP8 + ( col >> 2 ) : row * RP8 + col; }
};
constexpr size_t atrousHalfSmemBytes( int step )
{
    const int tw  = kAtTileW + 2 * ( step >= 4 ? atrousHalo( step ) : step );
    const int npx = ( step >= 4 ? 4 * atrousGroupPitch( step ) : tw ) * ( kAtTileH + 2 ), npx8 = ( step >= 4 ? 4 * atrousGroupPitch8( step ) : tw ) * ( kAtTileH + 2 );
#ifdef ASVGF_AT_HYBRID
    return ( size_t )npx * 48 + ( step == 1 ? ( size_t )npx * 4 : 0 ); // iteration 0 keeps the variances once more as a compact plane
#else
    return ( size_t )npx * 32 + ( size_t )npx8 * 8;
#endif
}

ASVGF_DEVICE uint32_t packHalf2( float lo, float hi )
{
    const __half2 h = __floats2half2_rn( lo, hi );
    return *reinterpret_cast< const uint32_t* >( &h );
}
ASVGF_DEVICE float2 unpackHalf2( uint32_t v ) { return __half22float2( *reinterpret_cast< const __half2* >( &v ) ); }
// hint only: the composition / centre texels of a pixel are requested before the tile is staged
#ifndef ASVGF_HOSTSIM
ASVGF_DEVICE void prefetchL1( const void* p ) { asm volatile( "prefetch.global.L1 [%0];" ::"l"( p ) ); }
#else
ASVGF_DEVICE void prefetchL1( const void* ) {}
#endif
// f32 = half - f32 in one instruction (FHADD): exact widening inside the add
#ifndef ASVGF_HOSTSIM
ASVGF_DEVICE float halfLoMinus( uint32_t v, float f )
{
    float r;
    asm( "{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; sub.rn.f32.f16 %0, lo, %2; }" : "=f"( r ) : "r"( v ), "f"( f ) );
    return r;
}
ASVGF_DEVICE float halfHiMinus( uint32_t v, float f )
{
    float r;
    asm( "{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; sub.rn.f32.f16 %0, hi, %2; }" : "=f"( r ) : "r"( v ), "f"( f ) );
    return r;
}
#else // tests/hostsim (host compile of this file by the CPU test suite): the same exactly rounded operation, spelled in C
ASVGF_DEVICE float halfLoMinus( uint32_t v, float f ) { return __fsub_rn( unpackHalf2( v ).x, f ); }
ASVGF_DEVICE float halfHiMinus( uint32_t v, float f ) { return __fsub_rn( unpackHalf2( v ).y, f ); }
#endif

// one staged texel -> the three planes
ASVGF_DEVICE void atrousPackTexel( const float4 ga, const uint32_t d0, const uint32_t d1, const uint32_t encSpec, const uint32_t encIndir, float4& pA, uint4& pB, AtC& pC )
{
    const float2 rg = unpackHalf2( d0 ), bv = unpackHalf2( d1 );
    const float3 s3 = decodeE5( encSpec ), i3 = decodeE5( encIndir );
    // depth came from an fp16 image and the roughness byte is an integer <= 255: both convert back exactly
    pA = make_float4( ga.x, ga.y, ga.z, __uint_as_float( packHalf2( guideDepth( ga.w ), guideRoughByte( ga.w ) ) ) );
#ifdef ASVGF_AT_HYBRID
    pB = make_uint4( __float_as_uint( rg.x ), __float_as_uint( rg.y ), __float_as_uint( bv.x ), __float_as_uint( bv.y ) );
    pC = make_uint4( __float_as_uint( luminance( make_float3( rg.x, rg.y, bv.x ) ) ), packHalf2( s3.x, s3.y ), packHalf2( s3.z, i3.x ), packHalf2( i3.y, i3.z ) );
#else
    pB = make_uint4( d0, d1, __float_as_uint( luminance( make_float3( rg.x, rg.y, bv.x ) ) ), packHalf2( s3.x, s3.y ) );
    pC = make_uint2( packHalf2( s3.z, i3.x ), packHalf2( i3.y, i3.z ) );
#endif
}

// Loads the tile into registers, takes the CTA's vote "does the tile hold a valid surface among its own pixels" from the guide
// depths just loaded (ncu of the form that voted on a separate depth load first: a fifth of the warps' time went into that extra
// round trip), and only then decodes into the planes.  Returns the vote; a sky-only tile leaves the planes untouched.  Every thread
// of the CTA reaches the vote (no early exit before it).  FUSED_VOTE is off for the strides whose tile takes more than one
// group per thread (>= 32): those keep the separate vote of the kernel body.
template< int STEP > constexpr bool atrousFusedVote() { return !AtrousGeom< STEP >::VEC || AtrousGeom< STEP >::GROUPS * AtrousGeom< STEP >::TROWS <= kAtThreads; }

template< int STEP, bool UNI, bool INTERIOR, bool VOTE >
ASVGF_DEVICE bool atrousStageHalfImpl( const AtrousArgs& a, const Uniform& u, const Rows& rows, int xBase, int yBase, float4* sA, uint4* sB, AtC* sC, float* sV )
{
    using G = AtrousGeomH< STEP >;
    constexpr bool FUSED = VOTE && atrousFusedVote< STEP >();
    bool mine = false;
    auto validDepth = []( float w ) { const float dz = guideDepth( w ); return !( dz < 0.0f || dz > 10000.0f ); };
    const uint32_t pitchT = a.guide.pitchT;
    const uint32_t off0 = ( uint32_t )( ( yBase - STEP ) * ( int )pitchT + xBase - G::HALO ); // only dereferenced under `ok`
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
            const int  g = g0 + threadIdx.x;
            const bool act = NGRP % kAtThreads == 0 || g < NGRP;
            if( !FUSED && !act ) break;
            const int  r = g / G::GROUPS, cg = g - r * G::GROUPS;
            const int  gx = xBase - G::HALO + 4 * cg, gy = yBase + ( r - 1 ) * STEP;
            const bool rowOk = INTERIOR || ( gy >= rows.y0 && gy < rows.y1 );
            const bool ok01 = act && ( INTERIOR || ( rowOk && gx >= 0 && gx < u.W ) ), ok23 = act && ( INTERIOR || ( rowOk && gx >= 0 && gx + 2 < u.W ) ); // gx is a multiple of 4, W is even
            float4 ga[ 4 ];
            uint4  d01, d23, sp, in;
            {
                const uint32_t off = off0 + ( uint32_t )r * ( ( uint32_t )STEP * pitchT ) + 4u * ( uint32_t )cg;
                const float4*  pg = UNI ? bG + off : at< float4 >( a.guide, gx, gy );
                const uint4*   pd = UNI ? reinterpret_cast< const uint4* >( bD + off ) : reinterpret_cast< const uint4* >( at< uint2 >( a.srcDiff, gx, gy ) );
                const uint4*   ps = UNI ? reinterpret_cast< const uint4* >( bS + off ) : reinterpret_cast< const uint4* >( at< uint32_t >( a.srcSpec, gx, gy ) );
                const uint4*   pi = UNI ? reinterpret_cast< const uint4* >( bI + off ) : reinterpret_cast< const uint4* >( at< uint32_t >( a.srcIndir, gx, gy ) );
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
            if constexpr( FUSED )
            {
                // the group's texels are the tile's own pixels in rows 1..kAtTileH, columns HALO..HALO+kAtTileW-1 (HALO is a multiple of 4)
                if( act && r >= 1 && r <= kAtTileH && cg >= G::HALO / 4 && cg < ( G::HALO + kAtTileW ) / 4 && gy < u.yEnd )
                {
#pragma unroll
                    for( int k = 0; k < 4; k++ ) mine = mine || ( gx + k < u.W && validDepth( ga[ k ].w ) );
                }
                if( !__syncthreads_or( mine ) ) return false;
                if( !act ) break;
            }
            const int      e0 = r * G::RP + cg, e8 = r * G::RP8 + cg; // texel k of the group -> quarter row k
            const uint32_t dd[ 8 ] = { d01.x, d01.y, d01.z, d01.w, d23.x, d23.y, d23.z, d23.w };
            const uint32_t ss[ 4 ] = { sp.x, sp.y, sp.z, sp.w }, ii[ 4 ] = { in.x, in.y, in.z, in.w };
#pragma unroll
            for( int k = 0; k < 4; k++ )
            {
                float4 pA; uint4 pB; AtC pC;
                atrousPackTexel( ga[ k ], dd[ 2 * k ], dd[ 2 * k + 1 ], ss[ k ], ii[ k ], pA, pB, pC );
                sA[ e0 + k * G::GP ]  = pA;
                sB[ e0 + k * G::GP ]  = pB;
                sC[ e8 + k * G::GP8 ] = pC;
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
            // one of the tile's own pixels?  (rows 1..kAtTileH, columns HALO..HALO+kAtTileW-1, inside the launch's rows and the image)
            if constexpr( FUSED ) mine = mine || ( ok && r >= 1 && r <= kAtTileH && c >= G::HALO && c < G::HALO + kAtTileW && gx < u.W && gy < u.yEnd && validDepth( ga[ it ].w ) );
        }
        if constexpr( FUSED )
            if( !__syncthreads_or( mine ) ) return false;
#pragma unroll
        for( int it = 0; it < NIT; it++ )
        {
            const int e = threadIdx.x + it * kAtThreads;
            if( ( it + 1 ) * kAtThreads > G::NPX && e >= G::NPX ) break;
            float4 pA; uint4 pB; AtC pC;
            atrousPackTexel( ga[ it ], dd[ it ].x, dd[ it ].y, ss[ it ], ii[ it ], pA, pB, pC );
            sA[ e ] = pA; sB[ e ] = pB; sC[ e ] = pC;
#ifdef ASVGF_AT_HYBRID
            if constexpr( STEP == 1 ) sV[ e ] = __uint_as_float( pB.w ); // the 3x3 variance filter of iteration 0 reads these: 4-byte stride, conflict-free
#endif
        }
    }
    return true;
}

template< int STEP, bool VOTE >
ASVGF_DEVICE bool atrousStageHalf( const AtrousArgs& a, const Uniform& u, const Rows& rows, int xBase, int yBase, float4* sA, uint4* sB, AtC* sC, float* sV )
{
    using G = AtrousGeom< STEP >;
    const bool interior = xBase - G::HALO >= 0 && xBase + kAtTileW + G::HALO <= u.W && yBase - STEP >= rows.y0 && yBase + kAtTileH * STEP < rows.y1;
    if( a.uniformPitch )
    {
        if( interior ) return atrousStageHalfImpl< STEP, true, true, VOTE >( a, u, rows, xBase, yBase, sA, sB, sC, sV );
        return atrousStageHalfImpl< STEP, true, false, VOTE >( a, u, rows, xBase, yBase, sA, sB, sC, sV );
    }
    return atrousStageHalfImpl< STEP, false, false, VOTE >( a, u, rows, xBase, yBase, sA, sB, sC, sV );
}

// the packed texel, widened: what AtrousPix::init / initFromCentre / finish take for the centre
struct AtrousWide
{
    float4 A, B, C, D; // the four planes of k_atrous_tiled: (n, depth), (diff, variance), (spec, luminance), (indir, roughness byte)
};
ASVGF_DEVICE AtrousWide atrousWiden( const float4 pA, const uint4 pB, const AtC pC )
{
    const float2 dr = unpackHalf2( __float_as_uint( pA.w ) );
#ifdef ASVGF_AT_HYBRID
    const float2 rg = make_float2( __uint_as_float( pB.x ), __uint_as_float( pB.y ) ), bv = make_float2( __uint_as_float( pB.z ), __uint_as_float( pB.w ) );
    const float2 s01 = unpackHalf2( pC.y ), s2i0 = unpackHalf2( pC.z ), i12 = unpackHalf2( pC.w );
    const float  lum = __uint_as_float( pC.x );
#else
    const float2 rg = unpackHalf2( pB.x ), bv = unpackHalf2( pB.y ), s01 = unpackHalf2( pB.w ), s2i0 = unpackHalf2( pC.x ), i12 = unpackHalf2( pC.y );
    const float  lum = __uint_as_float( pB.z );
#endif
    AtrousWide w;
    w.A = make_float4( pA.x, pA.y, pA.z, dr.x );
    w.B = make_float4( rg.x, rg.y, bv.x, bv.y );
    w.C = make_float4( s01.x, s01.y, s2i0.x, lum );
    w.D = make_float4( s2i0.y, i12.x, i12.y, dr.y );
    return w;
}

// one tap from the packed planes: AtrousPix::tap with the widening folded in (same operations on the same fp32 values)
ASVGF_DEVICE void atrousTapHalf( AtrousPix& s, const float4 qA, const uint4 qB, const AtC qC, const bool diag, const float lk )
{
    const float TINY = __uint_as_float( 0x00800000u );
    const float nN    = fmaxf( TINY, fmaf( s.n.z, qA.z, fmaf( s.n.y, qA.y, s.n.x * qA.x ) ) );
    const float L     = flg2( nN );
    const uint32_t dr = __float_as_uint( qA.w );
    const float wZ    = fabsf( halfLoMinus( dr, s.depth ) ) * ( diag ? s.invK2 : s.invK1 ); // |q.depth - depth| = |depth - q.depth|
    const float eN    = fmaf( 128.0f, L, lk );
    const float eBase = fmaf( -wZ, wZ, eN );
#ifdef ASVGF_AT_HYBRID
    const float qLum = __uint_as_float( qC.x );
#else
    const float qLum = __uint_as_float( qB.z );
#endif
    const float eDiff = fmaf( fabsf( s.l - qLum ), -s.kLum, eBase );
    const float wBase = fex2( eBase );
    const float wDiff = fex2( eDiff );
    float       rfac;
#ifndef ASVGF_HOSTSIM
    asm( "fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"( rfac ) : "f"( fabsf( halfHiMinus( dr, s.roughByte ) ) ), "f"( -s.wR10 ), "f"( s.wRough ) );
#else
    rfac = fminf( fmaxf( fmaf( fabsf( halfHiMinus( dr, s.roughByte ) ), -s.wR10, s.wRough ), 0.0f ), 1.0f ); // (.sat: NaN -> 0, as fmaxf gives)
#endif
    const float wSpec = rfac * fex2( fmaf( s.P, L, eBase ) );
#ifdef ASVGF_AT_HYBRID
    const float2 rg = make_float2( __uint_as_float( qB.x ), __uint_as_float( qB.y ) ), bv = make_float2( __uint_as_float( qB.z ), __uint_as_float( qB.w ) );
    const float2 s01 = unpackHalf2( qC.y ), s2i0 = unpackHalf2( qC.z ), i12 = unpackHalf2( qC.w );
#else
    const float2 rg = unpackHalf2( qB.x ), bv = unpackHalf2( qB.y ), s01 = unpackHalf2( qB.w ), s2i0 = unpackHalf2( qC.x ), i12 = unpackHalf2( qC.y );
#endif
    s.diff.x = fmaf( rg.x, wDiff, s.diff.x ); s.diff.y = fmaf( rg.y, wDiff, s.diff.y ); s.diff.z = fmaf( bv.x, wDiff, s.diff.z );
    s.spec.x = fmaf( s01.x, wSpec, s.spec.x ); s.spec.y = fmaf( s01.y, wSpec, s.spec.y ); s.spec.z = fmaf( s2i0.x, wSpec, s.spec.z );
    s.indir.x = fmaf( s2i0.y, wBase, s.indir.x ); s.indir.y = fmaf( i12.x, wBase, s.indir.y ); s.indir.z = fmaf( i12.y, wBase, s.indir.z );
    s.wSum += wDiff; s.wSumSpec += wSpec; s.wSumIndir += wBase;
    s.variance = fmaf( __fmul_rn( bv.y, wDiff ), wDiff, s.variance );
}

template< int STEP, bool ITER0, bool LAST, bool CENTRE >
__global__ void __launch_bounds__( kAtThreads, ASVGF_AT_MIN_CTAS ) k_atrous_half( const __grid_constant__ AtrousArgs a, const __grid_constant__ Uniform u )
{
    pdlLaunchDependents();
#ifndef ASVGF_HOSTSIM
    extern __shared__ __align__( 128 ) uint8_t smemRaw[];
#else
    uint8_t* const smemRaw = hostsim::dynamicSmem(); // tests/hostsim: the launcher's per-grid allocation
#endif
    using G = AtrousGeomH< STEP >;
    constexpr int RP = G::RP, RP8 = G::RP8, NPX = G::NPX, HALO = G::HALO;
    constexpr int PIX = ( kAtTileW * kAtTileH ) / kAtThreads; // 2
    static_assert( PIX == 2, "two pixels per thread" );
    float4* sA = reinterpret_cast< float4* >( smemRaw );
    uint4*  sB = reinterpret_cast< uint4* >( sA + NPX );
    AtC*    sC = reinterpret_cast< AtC* >( sB + NPX );
    float*  sV = reinterpret_cast< float* >( sC + G::NPX8 ); // iteration 0 only (atrousHalfSmemBytes)

    const int band  = blockIdx.y / STEP, res = blockIdx.y - band * STEP;
    const int yBase = u.yBegin + band * ( kAtTileH * STEP ) + res;
    const int xBase = blockIdx.x * kAtTileW;
    const Rows rows = rowsOf( a.depth );
    const int c  = threadIdx.x & ( kAtTileW - 1 ), m = threadIdx.x / kAtTileW; // pixels: column c, tile rows m and m + 4
    auto      jOf = [ & ]( int k ) { return m + 4 * k; };
    const int px = xBase + c;

    // sky-only tiles pass the centre texels through, no neighbourhood (CmSVGFAtrous.comp:96-99, Iter0.comp:208-216)
    auto skyTile = [ & ]() {
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
                const float2 rg = unpackHalf2( dd.x ), bv = unpackHalf2( dd.y );
                oDiff = make_float3( rg.x, rg.y, bv.x ); oVar = bv.y;
                oSpec  = decodeE5( __ldg( at< uint32_t >( a.srcSpec, px, py ) ) );
                oIndir = decodeE5( __ldg( at< uint32_t >( a.srcIndir, px, py ) ) );
            }
            atrousStore< LAST >( a, u, px, py, oDiff, oVar, oSpec, oIndir );
        }
    };
    pdlWait(); // everything below reads this frame's passes
    // Iteration 0 votes on the depth image first: its sky tiles store zeros and need nothing else, so loading their tile would be
    // pure waste (measured: 0.160 -> 0.168 ms with the fused vote); the other iterations take the vote from the staged tile.
    constexpr bool FUSED_VOTE = !ITER0 && atrousFusedVote< STEP >();
    if constexpr( !FUSED_VOTE )
    {
        bool anyValid = false;
#pragma unroll
        for( int k = 0; k < PIX; k++ )
        {
            const int  py = yBase + jOf( k ) * STEP;
            const bool ok = px < u.W && py < u.yEnd;
            const float dz = h2f( ldIf< uint16_t >( a.depth, px, py, ok, ( uint16_t )0x7c00 ) );
            anyValid = anyValid || ( ok && !( dz < 0.0f || dz > 10000.0f ) );
        }
        if( !__syncthreads_or( anyValid ) ) { skyTile(); return; }
    }

    static_assert( !( ITER0 && CENTRE ), "iteration 0 writes the centre plane" );
    uint16_t cGrad[ PIX ], cSigma[ PIX ], cSpecLen[ PIX ];
    if constexpr( CENTRE )
    {
#pragma unroll
        for( int k = 0; k < PIX; k++ )
        {
            const int py = yBase + jOf( k ) * STEP;
            if( px < u.W && py < u.yEnd ) prefetchL1( at< float4 >( a.centre, px, py ) );
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
#pragma unroll
        for( int k = 0; k < PIX; k++ )
        {
            const int py = yBase + jOf( k ) * STEP;
            if( px < u.W && py < u.yEnd )
            {
                const int sepx = u.W / 2, isOdd = px >= sepx ? 1 : 0, regX = ( px - isOdd * sepx ) * 2 + ( ( isOdd + py ) & 1 );
                prefetchL1( at< uint32_t >( a.albedo, regX, py ) );
                prefetchL1( at< uint2 >( a.throughput, px, py ) );
                prefetchL1( at< uint16_t >( a.metRough, px, py ) );
                prefetchL1( at< uint8_t >( a.isSky, px, py ) );
            }
        }
    }
    if( !atrousStageHalf< STEP, FUSED_VOTE >( a, u, rows, xBase, yBase, sA, sB, sC, sV ) ) { skyTile(); return; } // (the vote: uniform per CTA)
    __syncthreads();

    const int   sep  = u.W / 2;
    const float NINF = __uint_as_float( 0xff800000u );
    const bool tileInterior = ( xBase >= sep ? xBase - STEP >= sep : xBase + kAtTileW + STEP <= sep ) && xBase - STEP >= 0 && xBase + kAtTileW + STEP <= u.W &&
                              yBase - STEP >= 0 && yBase + kAtTileH * STEP < u.H;
    const int  ax0 = px >= sep ? sep : 0, ax1 = ax0 + sep;
    const bool okL = px - STEP >= ax0, okR = px + STEP < ax1;
    constexpr int dC = G::dCol;

#pragma unroll
    for( int k = 0; k < PIX; k++ )
    {
        const int j = jOf( k ), py = yBase + j * STEP;
        if( px >= u.W || py >= u.yEnd ) continue;
        const int ce = G::slot( j + 1, c + HALO ), ce8 = G::slot8( j + 1, c + HALO );
        AtrousPix        s;
        const AtrousWide cw = atrousWiden( sA[ ce ], sB[ ce ], sC[ ce8 ] );
        if constexpr( CENTRE )
        {
            s.initFromCentre( cw.A, cw.B, cw.C, cw.D, __ldg( at< float4 >( a.centre, px, py ) ), a.iteration );
        }
        else
        {
            float sigmaL;
            if( !ITER0 ) sigmaL = h2f( cSigma[ k ] );
            else
            {
                // Iter0.comp:158-187: 3x3 Gaussian {1/4,1/8,1/16} of the variance (products by powers of two are exact)
                float r = 0.0f;
#pragma unroll
                for( int yy = -1; yy <= 1; yy++ )
#pragma unroll
                    for( int xx = -1; xx <= 1; xx++ )
#ifdef ASVGF_AT_HYBRID
                        r = fmaf( sV[ ce + yy * RP + xx * dC ], ( xx == 0 ? 0.5f : 0.25f ) * ( yy == 0 ? 0.5f : 0.25f ), r ); // (plane B's .w at a 16-byte stride: 4-way conflicts)
#else
                        r = fmaf( unpackHalf2( sB[ ce + yy * RP + xx * dC ].y ).y, ( xx == 0 ? 0.5f : 0.25f ) * ( yy == 0 ? 0.5f : 0.25f ), r );
#endif
                sigmaL = __fsqrt_rn( fmaxf( r, 0.0f ) );
            }
            s.init( cw.A, cw.B, cw.C, cw.D, h2f( cGrad[ k ] ), sigmaL, h2f( cSpecLen[ k ] ), a.iteration );
            if( ITER0 )
            {
                if( !s.invalid ) *atW< uint16_t >( a.filteredVariance, px, py ) = f2h( sigmaL );
                *atW< float4 >( a.centre, px, py ) = make_float4( s.kLum, s.invK1, s.invK2, s.P );
            }
        }
        if( !s.invalid )
        {
            auto taps = [ & ]( auto interior ) {
                constexpr bool INTERIOR = decltype( interior )::value;
                const bool okT = py - STEP >= 0, okB = py + STEP < u.H;
#pragma unroll
                for( int yy = -1; yy <= 1; yy++ )
#pragma unroll
                    for( int xx = -1; xx <= 1; xx++ )
                    {
                        if( xx == 0 && yy == 0 ) continue;
                        const bool diag = xx != 0 && yy != 0;
                        const int  qe = ce + yy * RP + xx * dC, qe8 = ce8 + yy * RP8 + xx * dC;
                        const bool in = INTERIOR || ( ( xx < 0 ? okL : xx > 0 ? okR : true ) && ( yy < 0 ? okT : yy > 0 ? okB : true ) );
                        atrousTapHalf( s, sA[ qe ], sB[ qe ], sC[ qe8 ], diag, in ? ( diag ? -2.0f : -1.0f ) : NINF );
                    }
            };
            if( tileInterior ) taps( std::true_type{} );
            else taps( std::false_type{} );
            s.normalise();
        }
        else // :96-99 pass-through of the centre (iteration 0, Iter0.comp:208-216: zeros)
        {
            s.diff = make_float3( cw.B.x, cw.B.y, cw.B.z ); s.variance = cw.B.w; s.spec = make_float3( cw.C.x, cw.C.y, cw.C.z ); s.indir = make_float3( cw.D.x, cw.D.y, cw.D.z );
            if( ITER0 ) { s.variance = 0.0f; s.diff = s.spec = s.indir = make_float3( 0.0f, 0.0f, 0.0f ); }
        }
        atrousStore< LAST >( a, u, px, py, s.diff, s.variance, s.spec, s.indir );
    }
}

} // namespace asvgf
