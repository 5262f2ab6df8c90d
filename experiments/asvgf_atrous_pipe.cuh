// asvgf_atrous_pipe.cuh -- EXPERIMENT, not part of libasvgf.so: the a-trous kernel as a PERSISTENT, ASYNCHRONOUSLY STAGED pipeline
// (sm_100a).  Parity-green and measured 10-25 % slower than the CTA-per-tile kernel: 4K frame 1.4530 ms against
// 1.3245 ms; ncu (profile of the four launches): +11 % instructions, warps-active 40 % instead of 50 % (static tile schedule),
// three CTAs per SM at stride 8.  As measured it was built on the all-half tile (-DASVGF_AT_ALL_HALF: plane C = uint2) and
// launched from asvgf.cu with grid = min(tiles, SMs x atrousPipeCtasPerSm(step)) and the two extra arguments (tilesX, numTiles).
//
// Same filter and the same packed tile as k_atrous_half (asvgf_atrous_half.cuh), identical results.  What changes is when the tile
// arrives.  ncu of k_atrous_half: the largest stall left is the long scoreboard (3.0-3.5 warps per issue slot, a third of all
// stalled warps): every CTA loads its tile into registers, waits one DRAM round trip with all eight warps idle, decodes, and
// only then filters; four CTAs per SM hide part of that, the rest is exposed.  Here
//
//   * a CTA is persistent: grid = SMs x resident CTAs, CTA i filters tiles i, i + grid, i + 2 grid, ...;
//   * the RAW texels of the NEXT tile (guide 16 B, diffuse 8 B, specular 4 B, indirect 4 B: the images as stored) are moved
//     global -> shared by asynchronous copies (cp.async = LDGSTS: no registers, no waiting warp; out-of-image texels arrive as
//     zeros through the copy's zero fill -- the pinned fetch) while the current tile is being filtered;
//   * at the top of the loop the raw tile is decoded shared -> shared into the packed planes (the conversion the tap loop must
//     not pay eight times per pixel -- the reason a raw TMA box cannot feed the taps directly), the sky vote is taken from the
//     staged guide depths, and the raw buffer is free for the tile after next.
//
// Shared memory: packed planes (40 B per texel) + one raw buffer (32 B per texel): 47-53 KB for strides 1-4 (four CTAs per SM),
// 60 KB and more from stride 8 (three).
#pragma once

#include "../raytracedgl1_b200/csrc/asvgf_atrous_half.cuh"

namespace asvgf
{

ASVGF_DEVICE uint32_t smemAddr( const void* p ) { return ( uint32_t )__cvta_generic_to_shared( p ); }
// asynchronous copy of BYTES (4, 8 or 16) global -> shared; ok = false copies nothing and fills the destination with zeros
template< int BYTES > ASVGF_DEVICE void cpAsyncZfill( void* dstShared, const void* srcGlobal, bool ok )
{
    const uint32_t n = ok ? ( uint32_t )BYTES : 0u;
    if constexpr( BYTES == 16 ) // .cg: through L2 only -- the tile is read once per CTA, and L1 is what shared memory leaves over
        asm volatile( "cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"( smemAddr( dstShared ) ), "l"( srcGlobal ), "r"( n ) : "memory" );
    else
        asm volatile( "cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"( smemAddr( dstShared ) ), "l"( srcGlobal ), "n"( BYTES ), "r"( n ) : "memory" );
}
ASVGF_DEVICE void cpAsyncCommit() { asm volatile( "cp.async.commit_group;" ::: "memory" ); }
ASVGF_DEVICE void cpAsyncWaitAll() { asm volatile( "cp.async.wait_group 0;" ::: "memory" ); }

template< int STEP > struct AtrousGeomP : AtrousGeomH< STEP >
{
    using G = AtrousGeomH< STEP >;
    static constexpr int NRAW = G::TW * G::TROWS; // raw texels, row-major, unpadded
};
constexpr size_t atrousPipeSmemBytes( int step )
{
    const int tw = kAtTileW + 2 * ( step >= 4 ? atrousHalo( step ) : step );
    return atrousHalfSmemBytes( step ) + ( size_t )tw * ( kAtTileH + 2 ) * 32;
}
constexpr int atrousPipeCtasPerSm( int step ) { return atrousPipeSmemBytes( step ) + 1024 <= ( 227 * 1024 ) / 4 ? 4 : atrousPipeSmemBytes( step ) + 1024 <= ( 227 * 1024 ) / 3 ? 3 : atrousPipeSmemBytes( step ) + 1024 <= ( 227 * 1024 ) / 2 ? 2 : 1; }

struct AtrousTile
{
    int xBase, yBase;
};
template< int STEP > ASVGF_DEVICE AtrousTile atrousTileOf( const Uniform& u, int t, int tilesX )
{
    const int by = t / tilesX, bx = t - by * tilesX;
    const int band = by / STEP, res = by - band * STEP;
    return AtrousTile{ bx * kAtTileW, u.yBegin + band * ( kAtTileH * STEP ) + res };
}

// issue the asynchronous copies of one tile's raw texels (no wait)
template< int STEP, bool UNI >
ASVGF_DEVICE void atrousIssueRaw( const AtrousArgs& a, const Uniform& u, const Rows& rows, const AtrousTile tile, float4* rawG, uint2* rawD, uint32_t* rawS, uint32_t* rawI )
{
    using G = AtrousGeomP< STEP >;
    const uint32_t pitchT = a.guide.pitchT;
    const uint32_t off0 = ( uint32_t )( ( tile.yBase - STEP ) * ( int )pitchT + tile.xBase - G::HALO ); // only used under `ok`
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
            const int  gx = tile.xBase - G::HALO + 4 * cg, gy = tile.yBase + ( r - 1 ) * STEP;
            const bool rowOk = gy >= rows.y0 && gy < rows.y1;
            const bool ok01 = rowOk && gx >= 0 && gx < u.W, ok23 = rowOk && gx >= 0 && gx + 2 < u.W; // gx is a multiple of 4, W is even
            const uint32_t off = ok01 ? off0 + ( uint32_t )r * ( ( uint32_t )STEP * pitchT ) + 4u * ( uint32_t )cg : 0u; // (a valid address either way)
            const float4*   pg = UNI ? bG + off : ( ok01 ? at< float4 >( a.guide, gx, gy ) : bG );
            const uint2*    pd = UNI ? bD + off : ( ok01 ? at< uint2 >( a.srcDiff, gx, gy ) : bD );
            const uint32_t* ps = UNI ? bS + off : ( ok01 ? at< uint32_t >( a.srcSpec, gx, gy ) : bS );
            const uint32_t* pi = UNI ? bI + off : ( ok01 ? at< uint32_t >( a.srcIndir, gx, gy ) : bI );
            const int e = r * G::TW + 4 * cg;
            cpAsyncZfill< 16 >( rawG + e, pg, ok01 );
            cpAsyncZfill< 16 >( rawG + e + 1, pg + 1, ok01 );
            cpAsyncZfill< 16 >( rawG + e + 2, pg + 2, ok23 );
            cpAsyncZfill< 16 >( rawG + e + 3, pg + 3, ok23 );
            cpAsyncZfill< 16 >( rawD + e, pd, ok01 );
            cpAsyncZfill< 16 >( rawD + e + 2, pd + 2, ok23 );
            cpAsyncZfill< 16 >( rawS + e, ps, ok01 );
            cpAsyncZfill< 16 >( rawI + e, pi, ok01 );
        }
    }
    else
    {
        constexpr int NIT = ( G::NRAW + kAtThreads - 1 ) / kAtThreads;
#pragma unroll
        for( int it = 0; it < NIT; it++ )
        {
            const int e = threadIdx.x + it * kAtThreads;
            if( ( it + 1 ) * kAtThreads > G::NRAW && e >= G::NRAW ) break;
            const int  r = e / G::TW, c = e - r * G::TW;
            const int  gx = tile.xBase - G::HALO + c, gy = tile.yBase + ( r - 1 ) * STEP;
            const bool ok = rows.in( gx, gy );
            const uint32_t off = ok ? off0 + ( uint32_t )r * ( ( uint32_t )STEP * pitchT ) + ( uint32_t )c : 0u;
            cpAsyncZfill< 16 >( rawG + e, UNI ? bG + off : ( ok ? at< float4 >( a.guide, gx, gy ) : bG ), ok );
            cpAsyncZfill< 8 >( rawD + e, UNI ? bD + off : ( ok ? at< uint2 >( a.srcDiff, gx, gy ) : bD ), ok );
            cpAsyncZfill< 4 >( rawS + e, UNI ? bS + off : ( ok ? at< uint32_t >( a.srcSpec, gx, gy ) : bS ), ok );
            cpAsyncZfill< 4 >( rawI + e, UNI ? bI + off : ( ok ? at< uint32_t >( a.srcIndir, gx, gy ) : bI ), ok );
        }
    }
}

// raw -> packed planes (shared -> shared).  Returns whether this thread saw a valid surface among the tile's own pixels.
template< int STEP >
ASVGF_DEVICE bool atrousDecodeRaw( const Uniform& u, const AtrousTile tile, const float4* rawG, const uint2* rawD, const uint32_t* rawS, const uint32_t* rawI, float4* sA, uint4* sB, uint2* sC )
{
    using G = AtrousGeomP< STEP >;
    bool anyValid = false;
    // a staged texel (row r, column c) is one of the tile's own pixels when 1 <= r <= kAtTileH and HALO <= c < HALO + kAtTileW
    auto own = [ & ]( int r, int c, float depth ) {
        const int px = tile.xBase + c - G::HALO, py = tile.yBase + ( r - 1 ) * STEP;
        return r >= 1 && r <= kAtTileH && c >= G::HALO && c < G::HALO + kAtTileW && px < u.W && py < u.yEnd && !( depth < 0.0f || depth > 10000.0f );
    };
    // one texel per lane in both layouts: consecutive lanes read consecutive raw texels (conflict-free) and store to consecutive
    // columns' slots (conflict-free by the choice of the quarter-row pitches)
    constexpr int NIT = ( G::NRAW + kAtThreads - 1 ) / kAtThreads;
#pragma unroll
    for( int it = 0; it < NIT; it++ )
    {
        const int e = threadIdx.x + it * kAtThreads;
        if( ( it + 1 ) * kAtThreads > G::NRAW && e >= G::NRAW ) break;
        const int    r = e / G::TW, c = e - r * G::TW;
        const float4 ga = rawG[ e ];
        const uint2  dd = rawD[ e ];
        float4 pA; uint4 pB; uint2 pC;
        atrousPackTexel( ga, dd.x, dd.y, rawS[ e ], rawI[ e ], pA, pB, pC );
        const int s16 = G::VEC ? G::slot( r, c ) : e, s8 = G::VEC ? G::slot8( r, c ) : e;
        sA[ s16 ] = pA; sB[ s16 ] = pB; sC[ s8 ] = pC;
        anyValid = anyValid || own( r, c, guideDepth( ga.w ) );
    }
    return anyValid;
}

template< int STEP, bool ITER0, bool LAST, bool CENTRE >
__global__ void __launch_bounds__( kAtThreads, atrousPipeCtasPerSm( STEP ) ) k_atrous_pipe( const __grid_constant__ AtrousArgs a, const __grid_constant__ Uniform u, const int tilesX, const int numTiles )
{
    pdlLaunchDependents();
    extern __shared__ __align__( 128 ) uint8_t smemRaw[];
    using G = AtrousGeomP< STEP >;
    constexpr int RP = G::RP, RP8 = G::RP8, NPX = G::NPX, NPX8 = G::NPX8, HALO = G::HALO, NRAW = G::NRAW;
    constexpr int PIX = ( kAtTileW * kAtTileH ) / kAtThreads; // 2
    static_assert( PIX == 2, "two pixels per thread" );
    static_assert( !( ITER0 && CENTRE ), "iteration 0 writes the centre plane" );
    float4*   sA   = reinterpret_cast< float4* >( smemRaw );
    uint4*    sB   = reinterpret_cast< uint4* >( sA + NPX );
    uint2*    sC   = reinterpret_cast< uint2* >( sB + NPX );
    float4*   rawG = reinterpret_cast< float4* >( sC + NPX8 );
    uint2*    rawD = reinterpret_cast< uint2* >( rawG + NRAW );
    uint32_t* rawS = reinterpret_cast< uint32_t* >( rawD + NRAW );
    uint32_t* rawI = rawS + NRAW;

    const Rows rows = rowsOf( a.depth );
    const int  c = threadIdx.x & ( kAtTileW - 1 ), m = threadIdx.x / kAtTileW; // pixels: column c, tile rows m and m + 4
    const int  sep = u.W / 2;
    const float NINF = __uint_as_float( 0xff800000u );
    constexpr int dC = G::dCol;
    const bool uni = a.uniformPitch != 0;

    pdlWait(); // the tile images are this frame's passes
    int t = blockIdx.x;
    if( t < numTiles )
    {
        const AtrousTile first = atrousTileOf< STEP >( u, t, tilesX );
        if( uni ) atrousIssueRaw< STEP, true >( a, u, rows, first, rawG, rawD, rawS, rawI );
        else atrousIssueRaw< STEP, false >( a, u, rows, first, rawG, rawD, rawS, rawI );
    }
    cpAsyncCommit();

    for( ; t < numTiles; t += gridDim.x )
    {
        const AtrousTile tile = atrousTileOf< STEP >( u, t, tilesX );
        const int xBase = tile.xBase, yBase = tile.yBase, px = xBase + c;

        // centre-only inputs of this thread's pixels: on their way to L1 while the tile is decoded
#pragma unroll
        for( int k = 0; k < PIX; k++ )
        {
            const int py = yBase + ( m + 4 * k ) * STEP;
            if( px < u.W && py < u.yEnd )
            {
                if constexpr( CENTRE ) asm volatile( "prefetch.global.L1 [%0];" ::"l"( at< float4 >( a.centre, px, py ) ) );
                else
                {
                    asm volatile( "prefetch.global.L1 [%0];" ::"l"( at< uint16_t >( a.depthGrad, px, py ) ) );
                    asm volatile( "prefetch.global.L1 [%0];" ::"l"( at< uint2 >( a.histLen, px, py ) ) );
                    if constexpr( !ITER0 ) asm volatile( "prefetch.global.L1 [%0];" ::"l"( at< uint16_t >( a.filteredVariance, px, py ) ) );
                }
                if constexpr( LAST )
                {
                    const int isOdd = px >= sep ? 1 : 0, regX = ( px - isOdd * sep ) * 2 + ( ( isOdd + py ) & 1 );
                    asm volatile( "prefetch.global.L1 [%0];" ::"l"( at< uint32_t >( a.albedo, regX, py ) ) );
                    asm volatile( "prefetch.global.L1 [%0];" ::"l"( at< uint2 >( a.throughput, px, py ) ) );
                    asm volatile( "prefetch.global.L1 [%0];" ::"l"( at< uint16_t >( a.metRough, px, py ) ) );
                    asm volatile( "prefetch.global.L1 [%0];" ::"l"( at< uint8_t >( a.isSky, px, py ) ) );
                }
            }
        }

        cpAsyncWaitAll();
        __syncthreads(); // every thread's copies have landed; every thread is done with the previous tile's planes
        const bool mine = atrousDecodeRaw< STEP >( u, tile, rawG, rawD, rawS, rawI, sA, sB, sC );
        const bool anyValid = __syncthreads_or( mine ); // planes complete, raw buffer free

        // the next tile's raw texels travel while this one is filtered
        {
            const int tn = t + ( int )gridDim.x;
            if( tn < numTiles )
            {
                const AtrousTile next = atrousTileOf< STEP >( u, tn, tilesX );
                if( uni ) atrousIssueRaw< STEP, true >( a, u, rows, next, rawG, rawD, rawS, rawI );
                else atrousIssueRaw< STEP, false >( a, u, rows, next, rawG, rawD, rawS, rawI );
            }
            cpAsyncCommit();
        }

        const bool tileInterior = ( xBase >= sep ? xBase - STEP >= sep : xBase + kAtTileW + STEP <= sep ) && xBase - STEP >= 0 && xBase + kAtTileW + STEP <= u.W &&
                                  yBase - STEP >= 0 && yBase + kAtTileH * STEP < u.H;
        const int  ax0 = px >= sep ? sep : 0, ax1 = ax0 + sep;
        const bool okL = px - STEP >= ax0, okR = px + STEP < ax1;

#pragma unroll
        for( int k = 0; k < PIX; k++ )
        {
            const int j = m + 4 * k, py = yBase + j * STEP;
            if( px >= u.W || py >= u.yEnd ) continue;
            const int ce = G::slot( j + 1, c + HALO ), ce8 = G::slot8( j + 1, c + HALO );
            const AtrousWide cw = atrousWiden( sA[ ce ], sB[ ce ], sC[ ce8 ] );
            if( !anyValid ) // sky-only tile: pass the centre through (CmSVGFAtrous.comp:96-99; iteration 0 stores zeros, Iter0.comp:208-216)
            {
                float3 oDiff = make_float3( 0.0f, 0.0f, 0.0f ), oSpec = oDiff, oIndir = oDiff;
                float  oVar  = 0.0f;
                if( !ITER0 ) { oDiff = make_float3( cw.B.x, cw.B.y, cw.B.z ); oVar = cw.B.w; oSpec = make_float3( cw.C.x, cw.C.y, cw.C.z ); oIndir = make_float3( cw.D.x, cw.D.y, cw.D.z ); }
                atrousStore< LAST >( a, u, px, py, oDiff, oVar, oSpec, oIndir );
                continue;
            }
            AtrousPix s;
            if constexpr( CENTRE )
            {
                s.initFromCentre( cw.A, cw.B, cw.C, cw.D, __ldg( at< float4 >( a.centre, px, py ) ), a.iteration );
            }
            else
            {
                const float gradDepth = h2f( __ldg( at< uint16_t >( a.depthGrad, px, py ) ) );
                const float specLen   = h2f( __ldg( reinterpret_cast< const uint16_t* >( at< uint2 >( a.histLen, px, py ) ) + 2 ) ); // .b of rgba16f
                float       sigmaL;
                if( !ITER0 ) sigmaL = h2f( __ldg( at< uint16_t >( a.filteredVariance, px, py ) ) );
                else
                {
                    // Iter0.comp:158-187: 3x3 Gaussian {1/4,1/8,1/16} of the variance (products by powers of two are exact)
                    float r = 0.0f;
#pragma unroll
                    for( int yy = -1; yy <= 1; yy++ )
#pragma unroll
                        for( int xx = -1; xx <= 1; xx++ )
                            r = fmaf( unpackHalf2( sB[ ce + yy * RP + xx * dC ].y ).y, ( xx == 0 ? 0.5f : 0.25f ) * ( yy == 0 ? 0.5f : 0.25f ), r );
                    sigmaL = __fsqrt_rn( fmaxf( r, 0.0f ) );
                }
                s.init( cw.A, cw.B, cw.C, cw.D, gradDepth, sigmaL, specLen, a.iteration );
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
    cpAsyncWaitAll(); // (nothing outstanding: the last iteration issued no copies)
}

} // namespace asvgf
