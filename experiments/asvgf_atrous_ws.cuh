// asvgf_atrous_ws.cuh -- the SELECTABLE a-trous kernels (ASVGF_ATROUS=ws|t2|t3), sm_100a.  The default is k_atrous_tiled
// (asvgf_atrous_tiled.cuh); these share its arithmetic and its dilated tile (shaders CmSVGFAtrous_Iter0.comp:158-335 and
// CmSVGFAtrous.comp:65-349) and explore how the tile is staged and how the SM is kept busy.  All three are parity-tested
// (tests/test_parity_gpu.py::test_atrous_kernel_variants); DESIGN.md section 3 has the measurements and why none of
// them beats the default yet.
//
// Common to the three: STAGING in 4-pixel vector loads (uint4 / uint2 per lane) into QUARTER-ROW SLOTS (tile column x ->
// slot (x & 3) * GP + (x >> 2)), which keeps the four stores of a staging lane and the LDS.128 of a filtering quarter
// warp (lanes take columns 4 apart) conflict-free at every stride.
//
//   k_atrous_t2<.., TH, PAIR, CTAS> .. one tile per CTA, every thread stages then filters.  PAIR: a thread filters a
//       vertical pair of output pixels, so 12 tile entries serve 16 tap evaluations (2/3 of the LDS bytes) ("t2": 64x16
//       tiles, 2 CTAs/SM).  !PAIR: one pixel per tap loop at 64 registers ("t3": 64x8 tiles, 4 CTAs/SM).
//   k_atrous_ws ..................... ONE persistent CTA per SM with split roles:
//     * PRODUCER warps stream tile t+1 from HBM (full-line row segments of the six tapped images; the packed normals
//       one tile further ahead, so that the sin/cos table lookups ride in the same memory round trip), unpack every
//       tile pixel once into one of two shared-memory tile buffers, and stage the three centre-only inputs of each
//       output pixel next to it;
//     * CONSUMER warps filter tile t out of the other buffer (pixel pairs): nothing in their instruction stream waits
//       on HBM (only the composing iteration reads its four extra centre-only images directly, ~500 instructions
//       ahead of use).
//     Hand-off is two pairs of named barriers (full[b]: producers arrive / consumers sync; empty[b]: the reverse), so
//     the two roles never meet at a CTA-wide __syncthreads.  Tiles are dealt round-robin to the CTAs (tile = blockIdx.x
//     + k * gridDim.x) in row-major order, so the halos of neighbouring tiles are read while still in L2.
#pragma once

#include "asvgf_atrous_tiled.cuh"

namespace asvgf
{

#ifndef ASVGF_WS_CONSUMERS
    #define ASVGF_WS_CONSUMERS 512
#endif
#ifndef ASVGF_WS_PRODUCERS
    #define ASVGF_WS_PRODUCERS 256
#endif
constexpr int kWsConsumers = ASVGF_WS_CONSUMERS;
constexpr int kWsProducers = ASVGF_WS_PRODUCERS;
constexpr int kWsThreads   = kWsConsumers + kWsProducers;
// dilated rows per tile: 16 up to stride 8; 8 beyond, so that two tile buffers still fit the 227 KB of one SM
constexpr int atrousWsTileH( int step ) { return step <= 8 ? 16 : 8; }
static_assert( ( kAtTileW * 8 ) % kWsConsumers == 0, "consumer count must divide the tile" );

// Tile geometry.  The producers move four pixels per load (uint4 / uint2), so the halo is rounded up to 4 columns and a
// tile row holds G = (64 + 2*halo)/4 groups.  Shared-memory slot of tile column x in a row: (x & 3) * GP + (x >> 2) --
// the four pixels a producer lane unpacks go to four different quarter rows, so its stores are conflict-free, and GP
// is padded to 2 or 6 (mod 8) so that the consumers' LDS.128 of 8 consecutive columns hit 8 different bank groups.
// For strides >= 4 a tap at x +- STEP is the same quarter row at +- STEP/4: still an immediate offset.
constexpr int atrousWsHalo( int step ) { return step < 4 ? 4 : step; }
constexpr int atrousWsGroups( int step ) { return ( kAtTileW + 2 * atrousWsHalo( step ) ) / 4; }
constexpr int atrousWsGroupPitch( int step )
{
    int g = atrousWsGroups( step );
    while( g % 8 != 2 && g % 8 != 6 ) g++;
    return g;
}
constexpr size_t atrousWsBufferBytes( int step )
{
    return ( size_t )( 4 * atrousWsGroupPitch( step ) ) * ( atrousWsTileH( step ) + 2 ) * kAtBytesPerTilePixel + ( size_t )kAtTileW * atrousWsTileH( step ) * 8;
}
constexpr size_t atrousWsSmemBytes( int step ) { return 2 * atrousWsBufferBytes( step ); }

ASVGF_DEVICE void wsBarSync( int id ) { asm volatile( "bar.sync %0, %1;" ::"r"( id ), "n"( kWsThreads ) : "memory" ); }
ASVGF_DEVICE void wsBarArrive( int id ) { asm volatile( "bar.arrive %0, %1;" ::"r"( id ), "n"( kWsThreads ) : "memory" ); }
ASVGF_DEVICE float fmaSat( float a, float b, float c ) { float r; asm( "fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"( r ) : "f"( a ), "f"( b ), "f"( c ) ); return r; }

struct AtrousWsTile { int xBase, yBase; };
template< int STEP > ASVGF_DEVICE AtrousWsTile atrousWsTile( int tile, int tilesX, int yBegin )
{
    const int ty = tile / tilesX, tx = tile - ty * tilesX;
    const int band = ty / STEP, res = ty - band * STEP;
    return AtrousWsTile{ tx * kAtTileW, yBegin + band * ( atrousWsTileH( STEP ) * STEP ) + res };
}

// per-output-pixel state of the tap loop
struct AtrousPx
{
    float4 cA;                                                    // centre normal.xyz, depth
    float  l, roughness, kLum, wR, wR10, P, invK1, invK2;         // centre luminance / roughness and the weight constants derived from the centre
    float3 oD, oS, oI;                                            // running sums (start at the centre texel, weight 1)
    float  oV, wsD, wsS, wsI;
};

// one tap.  DIAG: xx != 0 && yy != 0 (wavelet weight 1/4 instead of 1/2, depth-gradient span 2 instead of 1).
// MASKED: the tap may lie outside the checkerboarded render area (ShaderCommonGLSLFunc.h:345-366) -> exponent -inf.
template< bool DIAG, bool MASKED >
ASVGF_DEVICE void atrousWsTap( AtrousPx& s, const float4 qA, const float4 qB, const float4 qC, const float4 qD, bool in )
{
    const float NINF = __uint_as_float( 0xff800000u ), PINF = __uint_as_float( 0x7f800000u );
    const float nN = fmaxf( 0.0f, fmaf( s.cA.z, qA.z, fmaf( s.cA.y, qA.y, s.cA.x * qA.x ) ) );
    const float L  = fmaxf( flg2( nN ), -126.0f );
    const float wZ = fabsf( s.cA.w - qA.w ) * ( DIAG ? s.invK2 : s.invK1 ); // already scaled by sqrt(log2 e)
    float       lk = DIAG ? -2.0f : -1.0f;                                   // log2 of the wavelet weight
    if( MASKED ) lk = in ? lk : NINF;
    const float e0    = fmaf( 128.0f, L, lk );
    const float eBase = fmaf( -wZ, wZ, e0 );
    const float eDiff = fmaf( fabsf( s.l - qC.w ), -s.kLum, eBase );
    const float wBase = fex2( eBase );
    const float wDiff = fex2( eDiff );
    const float rfac  = fmaSat( fabsf( s.roughness - qD.w ), -s.wR10, s.wR ); // max(0, 1 - 10|dr|) * wRough, both factors in [0,1]
    const float wSpec = rfac * fex2( fmaf( s.P, L, eBase ) );

    s.oD.x = fmaf( qB.x, wDiff, s.oD.x ); s.oD.y = fmaf( qB.y, wDiff, s.oD.y ); s.oD.z = fmaf( qB.z, wDiff, s.oD.z );
    s.oS.x = fmaf( qC.x, wSpec, s.oS.x ); s.oS.y = fmaf( qC.y, wSpec, s.oS.y ); s.oS.z = fmaf( qC.z, wSpec, s.oS.z );
    s.oI.x = fmaf( qD.x, wBase, s.oI.x ); s.oI.y = fmaf( qD.y, wBase, s.oI.y ); s.oI.z = fmaf( qD.z, wBase, s.oI.z );
    s.wsD += wDiff; s.wsS += wSpec; s.wsI += wBase;
    // (var*w)*w.  MUFU.EX2 flushes a denormal weight to zero where the shader's exp() keeps it; that only shows when the
    // tap's fp16 variance is +inf (inf*denormal = inf, inf*0 = NaN), so only that case pays for the denormal-preserving
    // exponential
    if( qB.w == PINF )
    {
        const float wd = ex2ApproxDenorm( eDiff );
        s.oV           = fmaf( __fmul_rn( qB.w, wd ), wd, s.oV );
    }
    else
    {
        s.oV = fmaf( __fmul_rn( qB.w, wDiff ), wDiff, s.oV );
    }
}

// the ten taps a vertical pair of output pixels (dilated rows j, j+1 of one column) reads besides each other's centre:
// rows r-1 (upper pixel only), r and r+1 (both), r+2 (lower pixel only) at columns -STEP, 0, +STEP.  Twelve tile entries
// serve sixteen tap evaluations, so the pair reads 2/3 of the shared-memory bytes two separate pixels would.
// rowA = slot of the upper pixel's tile row, RS = slots per tile row, sx[] = in-row slots of columns x-STEP, x, x+STEP
template< int RS, bool MASKED >
ASVGF_DEVICE void atrousWsPairTaps( const float4* __restrict__ sA, const float4* __restrict__ sB, const float4* __restrict__ sC, const float4* __restrict__ sD, int rowA,
                                    const int ( &sx )[ 3 ], AtrousPx& A, AtrousPx& B, bool okL, bool okR, bool okTA, bool okBA, bool okBB )
{
#pragma unroll
    for( int xx = -1; xx <= 1; xx++ ) // row above the upper pixel
    {
        const int    qe = rowA - RS + sx[ xx + 1 ];
        const float4 qA = sA[ qe ], qB = sB[ qe ], qC = sC[ qe ], qD = sD[ qe ];
        const bool   inx = xx < 0 ? okL : xx > 0 ? okR : true;
        if( xx != 0 ) atrousWsTap< true, MASKED >( A, qA, qB, qC, qD, inx && okTA );
        else atrousWsTap< false, MASKED >( A, qA, qB, qC, qD, okTA );
    }
#pragma unroll
    for( int xx = -1; xx <= 1; xx += 2 ) // the upper pixel's own row: straight for it, diagonal for the lower pixel
    {
        const int    qe = rowA + sx[ xx + 1 ];
        const float4 qA = sA[ qe ], qB = sB[ qe ], qC = sC[ qe ], qD = sD[ qe ];
        const bool   inx = xx < 0 ? okL : okR;
        atrousWsTap< false, MASKED >( A, qA, qB, qC, qD, inx );
        atrousWsTap< true, MASKED >( B, qA, qB, qC, qD, inx ); // the row above the lower pixel is the upper pixel's row: always inside
    }
#pragma unroll
    for( int xx = -1; xx <= 1; xx += 2 ) // the lower pixel's own row
    {
        const int    qe = rowA + RS + sx[ xx + 1 ];
        const float4 qA = sA[ qe ], qB = sB[ qe ], qC = sC[ qe ], qD = sD[ qe ];
        const bool   inx = xx < 0 ? okL : okR;
        atrousWsTap< true, MASKED >( A, qA, qB, qC, qD, inx && okBA );
        atrousWsTap< false, MASKED >( B, qA, qB, qC, qD, inx );
    }
#pragma unroll
    for( int xx = -1; xx <= 1; xx++ ) // row below the lower pixel
    {
        const int    qe = rowA + 2 * RS + sx[ xx + 1 ];
        const float4 qA = sA[ qe ], qB = sB[ qe ], qC = sC[ qe ], qD = sD[ qe ];
        const bool   inx = xx < 0 ? okL : xx > 0 ? okR : true;
        if( xx != 0 ) atrousWsTap< true, MASKED >( B, qA, qB, qC, qD, inx && okBB );
        else atrousWsTap< false, MASKED >( B, qA, qB, qC, qD, okBB );
    }
}

// weight constants of one output pixel from its centre texel (Iter0.comp:224-234, CmSVGFAtrous.comp:100-112)
ASVGF_DEVICE void atrousWsSetup( AtrousPx& s, const float4 cA, const float4 cB, const float4 cC, const float4 cD, const uint2 cen, float sigmaL, int iteration )
{
    constexpr float LOG2E = 1.4426950408889634f, SQRT_LOG2E = 1.2011224087864498f;
    s.cA = cA; s.l = cC.w; s.roughness = cD.w;
    s.oD = make_float3( cB.x, cB.y, cB.z ); s.oV = cB.w;
    s.oS = make_float3( cC.x, cC.y, cC.z ); s.oI = make_float3( cD.x, cD.y, cD.z );
    s.wsD = s.wsS = s.wsI = 1.0f;
    const float gradDepth         = h2f( ( uint16_t )( cen.x & 0xffffu ) );
    const float historyLengthSpec = h2f( ( uint16_t )( cen.y & 0xffffu ) );
    s.kLum = LOG2E * frcp( fmaf( 4.0f, sigmaL, 0.00001f ) );
    s.wR   = saturatef( fmaf( s.roughness, 30.0f, -( float )iteration ) );
    s.wR10 = s.wR * 10.0f;
    const float alpha = s.roughness * s.roughness;
    float       P     = fmaxf( 0.01f, fmaf( 2.0f, frcp( fmaf( alpha, alpha, 1e-4f ) ), -2.0f ) ); // BRDF.h:27-29
    s.P               = clampf( P, 8.0f, 1024.0f ) * saturatef( historyLengthSpec * 0.125f );
    const float rg = frcp( gradDepth ); // 1/max(g,.01) = min(1/g,100) and 1/max(2g,.01) = min(.5/g,100) from one MUFU (g >= 0; 1/0 = inf)
    s.invK1 = SQRT_LOG2E * fminf( rg, 100.0f ); s.invK2 = SQRT_LOG2E * fminf( 0.5f * rg, 100.0f );
}
ASVGF_DEVICE void atrousWsNormalise( AtrousPx& s )
{
    const float inv = frcp( s.wsD ), invSpec = frcp( s.wsS ), invIndir = frcp( s.wsI );
    s.oD = make_float3( s.oD.x * inv, s.oD.y * inv, s.oD.z * inv );
    s.oV = s.oV * ( inv * inv );
    s.oS = make_float3( s.oS.x * invSpec, s.oS.y * invSpec, s.oS.z * invSpec );
    s.oI = make_float3( s.oI.x * invIndir, s.oI.y * invIndir, s.oI.z * invIndir );
}
ASVGF_DEVICE void atrousWsStore( const AtrousArgs& a, int px, int py, const AtrousPx& s )
{
    const __half2 o0 = __floats2half2_rn( s.oD.x, s.oD.y ), o1 = __floats2half2_rn( s.oD.z, s.oV );
    *atW< uint2 >( a.dstDiff, px, py ) = make_uint2( *reinterpret_cast< const uint32_t* >( &o0 ), *reinterpret_cast< const uint32_t* >( &o1 ) );
    *atW< uint32_t >( a.dstSpec, px, py )  = encodeE5Fast( s.oS );
    *atW< uint32_t >( a.dstIndir, px, py ) = encodeE5Fast( s.oI );
}

// composition of the last iteration, CmSVGFAtrous.comp:278-348
struct AtrousComposeRaw { uint32_t albedo, sky, metRough; uint2 tp; int regX; };
ASVGF_DEVICE AtrousComposeRaw atrousComposeLoad( const AtrousArgs& a, const Uniform& u, int px, int py, bool ok )
{
    AtrousComposeRaw r;
    const int sep = u.W / 2, isOdd = px >= sep ? 1 : 0;
    r.regX     = ( px - isOdd * sep ) * 2 + ( ( isOdd + py ) & 1 ); // ShaderCommonGLSLFunc.h:331-342
    r.albedo   = ldIf< uint32_t >( a.albedo, r.regX, py, ok, 0u );
    r.tp       = ldIf< uint2 >( a.throughput, px, py, ok, make_uint2( 0u, 0u ) );
    r.sky      = ldIf< uint8_t >( a.isSky, px, py, ok, ( uint8_t )0 );
    r.metRough = ldIf< uint16_t >( a.metRough, px, py, ok, ( uint16_t )0 );
    return r;
}
ASVGF_DEVICE void atrousCompose( const AtrousArgs& a, const Uniform& u, int px, int py, const AtrousPx& s, const AtrousComposeRaw& raw )
{
    const float3 outDiff = s.oD, outSpec = s.oS, outIndir = s.oI;
    float3       albedo  = unpackUfloat111110Fast( raw.albedo );
    const float2 tp01    = __half22float2( *reinterpret_cast< const __half2* >( &raw.tp.x ) );
    const float3 throughput = make_float3( tp01.x, tp01.y, __low2float( *reinterpret_cast< const __half2* >( &raw.tp.y ) ) );
    if( ( u.debugShowFlags & ( 1u << 9 ) ) != 0 ) albedo = make_float3( 1.0f, 1.0f, 1.0f );

    float3 illuminated;
    if( raw.sky == 0u )
    {
        const float metallic = unorm8ToFloatFast( raw.metRough & 0xffu );
        const float tt = saturatef( ( s.roughness - 0.5f ) * 4.0f ); // smoothstep(0.5, 0.75, roughness)
        const float sm = tt * tt * ( 3.0f - 2.0f * tt );
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
    *atW< uint16_t >( a.histogramInput, raw.regX, py ) = f2h( fmaxf( 0.0f, flum( illuminated ) ) );
    if( ( float )raw.regX < mulr( u.lightmapScreenCoverage, u.renderWidth ) ) illuminated = make_float3( albedo.x * throughput.x, albedo.y * throughput.y, albedo.z * throughput.z ); // classicShading
    *atW< uint32_t >( a.preFinal, raw.regX, py ) =
        packUfloat111110Fast( make_float3( fmaxf( 0.0f, illuminated.x ), fmaxf( 0.0f, illuminated.y ), fmaxf( 0.0f, illuminated.z ) ) );
}

// One vertical pair of output pixels (dilated rows 2*jp, 2*jp+1 of tile column c) out of a staged tile.
// cenA / cenB: (depth gradient | sigma << 16, specular history length) as fp16 bits.
template< int STEP, bool ITER0, bool LAST >
ASVGF_DEVICE void atrousPairUnit( const AtrousArgs& a, const Uniform& u, const float4* __restrict__ sA, const float4* __restrict__ sB, const float4* __restrict__ sC,
                                  const float4* __restrict__ sD, const AtrousWsTile t, int jp, int c, const uint2 cenA, const uint2 cenB )
{
    constexpr int HALO = atrousWsHalo( STEP ), GP = atrousWsGroupPitch( STEP ), RS = 4 * GP;
    const int     sep  = u.W / 2;
    const int px = t.xBase + c, pyA = t.yBase + 2 * jp * STEP, pyB = pyA + STEP;
    if( px >= u.W || pyA >= u.yEnd ) return;
    const bool haveB = pyB < u.yEnd;
    const int  x  = c + HALO;                                   // tile column of the pair
    int        sx[ 3 ];                                         // in-row slots of columns x-STEP, x, x+STEP
    sx[ 1 ] = ( x & 3 ) * GP + ( x >> 2 );
    if( STEP >= 4 ) { sx[ 0 ] = sx[ 1 ] - STEP / 4; sx[ 2 ] = sx[ 1 ] + STEP / 4; }
    else { sx[ 0 ] = ( ( x - STEP ) & 3 ) * GP + ( ( x - STEP ) >> 2 ); sx[ 2 ] = ( ( x + STEP ) & 3 ) * GP + ( ( x + STEP ) >> 2 ); }
    const int  rowA = ( 2 * jp + 1 ) * RS;                      // tile row of the upper pixel
    const int  ceA = rowA + sx[ 1 ], ceB = ceA + RS;            // the pair's own tile entries

    // the composing iteration's extra centre-only inputs: in flight behind the taps
    AtrousComposeRaw rawA, rawB;
    if( LAST )
    {
        rawA = atrousComposeLoad( a, u, px, pyA, true );
        rawB = atrousComposeLoad( a, u, px, pyB, haveB );
    }

    AtrousPx A, B;
    bool     validA, validB;
    {
        const float4 aA = sA[ ceA ], aB = sB[ ceA ], aC = sC[ ceA ], aD = sD[ ceA ];
        const float4 bA = sA[ ceB ], bB = sB[ ceB ], bC = sC[ ceB ], bD = sD[ ceB ];
        validA = !( aA.w < 0.0f || aA.w > 10000.0f );
        validB = haveB && !( bA.w < 0.0f || bA.w > 10000.0f );
        if( !validA && !validB )
        {
            // Iter0.comp:208-216 writes zeros, CmSVGFAtrous.comp:96-99 passes the centre through
            A.oD = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : make_float3( aB.x, aB.y, aB.z ); A.oV = ITER0 ? 0.0f : aB.w;
            A.oS = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : make_float3( aC.x, aC.y, aC.z );
            A.oI = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : make_float3( aD.x, aD.y, aD.z );
            B.oD = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : make_float3( bB.x, bB.y, bB.z ); B.oV = ITER0 ? 0.0f : bB.w;
            B.oS = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : make_float3( bC.x, bC.y, bC.z );
            B.oI = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : make_float3( bD.x, bD.y, bD.z );
            A.roughness = aD.w; B.roughness = bD.w;
        }
        else
        {
            float sigmaA, sigmaB;
            if( ITER0 ) // Iter0.comp:158-187: 3x3 Gaussian {1/4,1/8,1/16} of the variance (products by powers of two are exact)
            {
                float rA = 0.0f, rB = 0.0f;
#pragma unroll
                for( int yy = -1; yy <= 1; yy++ )
#pragma unroll
                    for( int xx = -1; xx <= 1; xx++ )
                    {
                        const float k = ( xx == 0 ? 0.5f : 0.25f ) * ( yy == 0 ? 0.5f : 0.25f );
                        rA = fmaf( sB[ rowA + yy * RS + sx[ xx + 1 ] ].w, k, rA );
                        rB = fmaf( sB[ rowA + ( yy + 1 ) * RS + sx[ xx + 1 ] ].w, k, rB );
                    }
                sigmaA = __fsqrt_rn( fmaxf( rA, 0.0f ) ); sigmaB = __fsqrt_rn( fmaxf( rB, 0.0f ) );
                if( validA ) *atW< uint16_t >( a.filteredVariance, px, pyA ) = f2h( sigmaA );
                if( validB ) *atW< uint16_t >( a.filteredVariance, px, pyB ) = f2h( sigmaB );
            }
            if( !ITER0 ) { sigmaA = h2f( ( uint16_t )( cenA.x >> 16 ) ); sigmaB = h2f( ( uint16_t )( cenB.x >> 16 ) ); }
            atrousWsSetup( A, aA, aB, aC, aD, cenA, sigmaA, a.iteration );
            atrousWsSetup( B, bA, bB, bC, bD, cenB, sigmaB, a.iteration );
            // each other's centre is the straight vertical tap (always inside the render area for the upper pixel's
            // row; the lower pixel's row is inside iff it exists in the frame)
            const int  ax0 = px >= sep ? sep : 0, ax1 = ax0 + sep;
            const bool okL = px - STEP >= ax0, okR = px + STEP < ax1, okTA = pyA - STEP >= 0, okBA = pyB < u.H, okBB = pyB + STEP < u.H;
            if( okL && okR && okTA && okBB )
            {
                atrousWsTap< false, false >( A, bA, bB, bC, bD, true );
                atrousWsTap< false, false >( B, aA, aB, aC, aD, true );
                atrousWsPairTaps< RS, false >( sA, sB, sC, sD, rowA, sx, A, B, true, true, true, true, true );
            }
            else
            {
                atrousWsTap< false, true >( A, bA, bB, bC, bD, okBA );
                atrousWsTap< false, false >( B, aA, aB, aC, aD, true );
                atrousWsPairTaps< RS, true >( sA, sB, sC, sD, rowA, sx, A, B, okL, okR, okTA, okBA, okBB );
            }
            atrousWsNormalise( A ); atrousWsNormalise( B );
            if( !validA || !validB ) // one of the two is sky / invalid: it takes the pass-through (or zero) result
            {
                const int    ce = validA ? ceB : ceA;
                AtrousPx&    X  = validA ? B : A;
                const float4 xB = sB[ ce ], xC = sC[ ce ], xD = sD[ ce ];
                X.oD = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : make_float3( xB.x, xB.y, xB.z ); X.oV = ITER0 ? 0.0f : xB.w;
                X.oS = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : make_float3( xC.x, xC.y, xC.z );
                X.oI = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : make_float3( xD.x, xD.y, xD.z );
            }
        }
    }

    if( !LAST )
    {
        atrousWsStore( a, px, pyA, A );
        if( haveB ) atrousWsStore( a, px, pyB, B );
    }
    else
    {
        atrousCompose( a, u, px, pyA, A, rawA );
        if( haveB ) atrousCompose( a, u, px, pyB, B, rawB );
    }
}

template< int STEP, bool ITER0, bool LAST >
__global__ void __launch_bounds__( kWsThreads, 1 ) k_atrous_ws( const __grid_constant__ AtrousArgs a, const __grid_constant__ Uniform u )
{
    extern __shared__ __align__( 16 ) uint8_t smemRaw[];
    constexpr int    HALO  = atrousWsHalo( STEP );
    constexpr int    G     = atrousWsGroups( STEP );      // 4-pixel groups per tile row
    constexpr int    GP    = atrousWsGroupPitch( STEP );  // padded
    constexpr int    RS    = 4 * GP;                      // slots per tile row
    constexpr int    TH    = atrousWsTileH( STEP );
    constexpr int    TROWS = TH + 2;
    constexpr int    NPX   = RS * TROWS;                  // slots per plane
    constexpr int    TILE_PIX = kAtTileW * TH;            // output pixels per tile
    constexpr size_t BUF   = atrousWsBufferBytes( STEP );
    constexpr int    TUNITS = TROWS * G;                                         // 4-pixel groups of the tile
    constexpr int    NR    = ( TUNITS + kWsProducers - 1 ) / kWsProducers;        // groups per producer thread
    constexpr int    CUNITS = TILE_PIX / 4;                                      // 4-pixel groups of output pixels (centre-only inputs)
    constexpr int    CR    = ( CUNITS + kWsProducers - 1 ) / kWsProducers;
    // named barriers: 1 + b = full[b], 3 + b = empty[b]

    const int tilesX = ( u.W + kAtTileW - 1 ) / kAtTileW;
    const int nTiles = tilesX * a.tilesPerCta; // tilesPerCta carries the number of dilated row sets (bands * STEP) for this kernel
    const int nLocal = ( nTiles - ( int )blockIdx.x + ( int )gridDim.x - 1 ) / ( int )gridDim.x;

    if( threadIdx.x >= kWsConsumers )
    {
        // ------------------------------------------------ producers ------------------------------------------------
        const int  ptid = threadIdx.x - kWsConsumers;
        const Rows rows = rowsOf( a.depth );
        // the packed normals are fetched one tile ahead, so that the sin/cos table entries they select can be requested
        // together with the rest of the tile: one memory round trip per tile instead of two dependent ones
        auto loadNormals = [ & ]( int it, uint4( &dst )[ NR ] )
        {
            const AtrousWsTile t = atrousWsTile< STEP >( ( int )blockIdx.x + it * ( int )gridDim.x, tilesX, u.yBegin );
#pragma unroll
            for( int k = 0; k < NR; k++ )
            {
                const int  un = ptid + k * kWsProducers;
                const int  r = un / G, g = un - r * G;
                const int  gx = t.xBase - HALO + 4 * g, gy = t.yBase + ( r - 1 ) * STEP;
                const bool ok = it < nLocal && un < TUNITS && rows.in( gx, gy );
                dst[ k ] = ok ? __ldg( reinterpret_cast< const uint4* >( at< uint32_t >( a.normal, gx, gy ) ) ) : make_uint4( 0u, 0u, 0u, 0u );
            }
        };
        uint4 rnNext[ NR ];
        loadNormals( 0, rnNext );
        for( int it = 0; it < nLocal; it++ )
        {
            const int b = it & 1;
            uint8_t*  buf = smemRaw + ( size_t )b * BUF;
            float4*   sA = reinterpret_cast< float4* >( buf ); // normal.xyz, depth
            float4*   sB = sA + NPX;                           // diffuse rgb, variance
            float4*   sC = sB + NPX;                           // specular rgb, luminance of the diffuse rgb
            float4*   sD = sC + NPX;                           // indirect rgb, roughness
            uint2*    sCen = reinterpret_cast< uint2* >( sD + NPX );
            const AtrousWsTile t = atrousWsTile< STEP >( ( int )blockIdx.x + it * ( int )gridDim.x, tilesX, u.yBegin );

            // every load of the tile is issued before the first use.  One lane moves four pixels per load.
            uint4  rn[ NR ], rd0[ NR ], rd1[ NR ], rs[ NR ], ri[ NR ];
            uint2  rz[ NR ], rm[ NR ];
            float2 lp[ NR ][ 4 ], lt[ NR ][ 4 ];
#pragma unroll
            for( int k = 0; k < NR; k++ )
            {
                rn[ k ] = rnNext[ k ];
                const uint32_t n4[ 4 ] = { rn[ k ].x, rn[ k ].y, rn[ k ].z, rn[ k ].w };
#pragma unroll
                for( int q = 0; q < 4; q++ )
                {
                    lp[ k ][ q ] = __ldg( a.lut.phi + ( n4[ q ] >> 16 ) );
                    lt[ k ][ q ] = __ldg( a.lut.theta + ( n4[ q ] & 0xffffu ) );
                }
            }
#pragma unroll
            for( int k = 0; k < NR; k++ )
            {
                const int  un = ptid + k * kWsProducers;
                const int  r = un / G, g = un - r * G;
                const int  gx = t.xBase - HALO + 4 * g, gy = t.yBase + ( r - 1 ) * STEP;
                const bool ok = un < TUNITS && rows.in( gx, gy ); // W is a multiple of 4: a group is inside or outside as a whole
                const uint4 z4 = make_uint4( 0u, 0u, 0u, 0u );
                const uint2 z2 = make_uint2( 0u, 0u );
                rz[ k ]  = ok ? __ldg( reinterpret_cast< const uint2* >( at< uint16_t >( a.depth, gx, gy ) ) ) : z2;
                rm[ k ]  = ok ? __ldg( reinterpret_cast< const uint2* >( at< uint16_t >( a.metRough, gx, gy ) ) ) : z2;
                rd0[ k ] = ok ? __ldg( reinterpret_cast< const uint4* >( at< uint2 >( a.srcDiff, gx, gy ) ) ) : z4;
                rd1[ k ] = ok ? __ldg( reinterpret_cast< const uint4* >( at< uint2 >( a.srcDiff, gx, gy ) ) + 1 ) : z4;
                rs[ k ]  = ok ? __ldg( reinterpret_cast< const uint4* >( at< uint32_t >( a.srcSpec, gx, gy ) ) ) : z4;
                ri[ k ]  = ok ? __ldg( reinterpret_cast< const uint4* >( at< uint32_t >( a.srcIndir, gx, gy ) ) ) : z4;
            }
            loadNormals( it + 1, rnNext );
            // centre-only inputs of the output pixels: depth gradient, filtered sigma, specular history length (fp16 bits)
            uint2 cg[ CR ], cs[ CR ];
            uint4 ch0[ CR ], ch1[ CR ];
#pragma unroll
            for( int k = 0; k < CR; k++ )
            {
                const int  cu = ptid + k * kWsProducers;
                const int  j = cu / ( kAtTileW / 4 ), c4 = ( cu - j * ( kAtTileW / 4 ) ) * 4;
                const int  px = t.xBase + c4, py = t.yBase + j * STEP;
                const bool ok = cu < CUNITS && px < u.W && py < u.yEnd;
                const uint4 z4 = make_uint4( 0u, 0u, 0u, 0u );
                const uint2 z2 = make_uint2( 0u, 0u );
                cg[ k ]  = ok ? __ldg( reinterpret_cast< const uint2* >( at< uint16_t >( a.depthGrad, px, py ) ) ) : z2;
                cs[ k ]  = ( !ITER0 && ok ) ? __ldg( reinterpret_cast< const uint2* >( at< uint16_t >( a.filteredVariance, px, py ) ) ) : z2;
                ch0[ k ] = ok ? __ldg( reinterpret_cast< const uint4* >( at< uint2 >( a.histLen, px, py ) ) ) : z4;
                ch1[ k ] = ok ? __ldg( reinterpret_cast< const uint4* >( at< uint2 >( a.histLen, px, py ) ) + 1 ) : z4;
            }
            if( it >= 2 ) wsBarSync( 3 + b ); // the consumers are done with this buffer
#pragma unroll
            for( int k = 0; k < CR; k++ )
            {
                const int cu = ptid + k * kWsProducers;
                if( cu >= CUNITS ) break;
                // slot p = j*64 + c: (gradDepth | sigma << 16, histLen.z); histLen texel = (x: diff|indir, y: spec|hitDist)
                uint4* dst = reinterpret_cast< uint4* >( sCen + cu * 4 );
                dst[ 0 ] = make_uint4( ( cg[ k ].x & 0xffffu ) | ( cs[ k ].x << 16 ), ch0[ k ].y & 0xffffu, ( cg[ k ].x >> 16 ) | ( cs[ k ].x & 0xffff0000u ), ch0[ k ].w & 0xffffu );
                dst[ 1 ] = make_uint4( ( cg[ k ].y & 0xffffu ) | ( cs[ k ].y << 16 ), ch1[ k ].y & 0xffffu, ( cg[ k ].y >> 16 ) | ( cs[ k ].y & 0xffff0000u ), ch1[ k ].w & 0xffffu );
            }
#pragma unroll
            for( int k = 0; k < NR; k++ )
            {
                const int un = ptid + k * kWsProducers;
                if( un >= TUNITS ) break;
                const int r = un / G, g = un - r * G;
                const uint32_t d4[ 8 ] = { rd0[ k ].x, rd0[ k ].y, rd0[ k ].z, rd0[ k ].w, rd1[ k ].x, rd1[ k ].y, rd1[ k ].z, rd1[ k ].w };
                const uint32_t s4[ 4 ] = { rs[ k ].x, rs[ k ].y, rs[ k ].z, rs[ k ].w }, i4[ 4 ] = { ri[ k ].x, ri[ k ].y, ri[ k ].z, ri[ k ].w };
                const uint32_t z4[ 2 ] = { rz[ k ].x, rz[ k ].y }, m4[ 2 ] = { rm[ k ].x, rm[ k ].y };
#pragma unroll
                for( int q = 0; q < 4; q++ )
                {
                    const int    e  = r * RS + q * GP + g;
                    const float3 n  = make_float3( mulr( lp[ k ][ q ].x, lt[ k ][ q ].y ), mulr( lp[ k ][ q ].x, lt[ k ][ q ].x ), lp[ k ][ q ].y ); // decodeNormal, Utils.h:124-140
                    const float2 rg = __half22float2( *reinterpret_cast< const __half2* >( &d4[ 2 * q ] ) );
                    const float2 bv = __half22float2( *reinterpret_cast< const __half2* >( &d4[ 2 * q + 1 ] ) );
                    const float3 sp = decodeE5( s4[ q ] ), in = decodeE5( i4[ q ] );
                    const uint32_t zq = ( z4[ q >> 1 ] >> ( 16 * ( q & 1 ) ) ) & 0xffffu, mq = ( m4[ q >> 1 ] >> ( 16 * ( q & 1 ) ) ) & 0xffffu;
                    sA[ e ] = make_float4( n.x, n.y, n.z, h2f( ( uint16_t )zq ) );
                    sB[ e ] = make_float4( rg.x, rg.y, bv.x, bv.y );
                    sC[ e ] = make_float4( sp.x, sp.y, sp.z, luminance( make_float3( rg.x, rg.y, bv.x ) ) );
                    sD[ e ] = make_float4( in.x, in.y, in.z, unorm8ToFloatFast( mq >> 8 ) );
                }
            }
            __threadfence_block();
            wsBarArrive( 1 + b );
        }
        return;
    }

    // -------------------------------------------------- consumers --------------------------------------------------
    // a consumer thread filters a vertical pair of output pixels: dilated rows 2*jp and 2*jp+1 of one tile column
    constexpr int UNITS = kAtTileW * ( TH / 2 );
    for( int it = 0; it < nLocal; it++ )
    {
        const int      b = it & 1;
        const uint8_t* buf = smemRaw + ( size_t )b * BUF;
        const float4*  sA = reinterpret_cast< const float4* >( buf );
        const float4*  sB = sA + NPX;
        const float4*  sC = sB + NPX;
        const float4*  sD = sC + NPX;
        const uint2*   sCen = reinterpret_cast< const uint2* >( sD + NPX );
        const AtrousWsTile t = atrousWsTile< STEP >( ( int )blockIdx.x + it * ( int )gridDim.x, tilesX, u.yBegin );
        wsBarSync( 1 + b );

#pragma unroll
        for( int unit = threadIdx.x; unit < UNITS; unit += kWsConsumers )
        {
            // lanes of a quarter warp take columns 4 apart (same quarter row of the tile, consecutive slots): every
            // LDS.128 of a tap is conflict-free at any stride; the warp still covers 32 consecutive columns
            const int jp = unit / kAtTileW, c0 = unit - jp * kAtTileW;
            const int c  = ( c0 & ~31 ) + ( ( c0 & 7 ) << 2 ) + ( ( c0 >> 3 ) & 3 );
            const int pA = 2 * jp * kAtTileW + c;
            atrousPairUnit< STEP, ITER0, LAST >( a, u, sA, sB, sC, sD, t, jp, c, sCen[ pA ], sCen[ pA + kAtTileW ] );
        }
        __syncwarp();
        if( it + 2 < nLocal ) wsBarArrive( 3 + b ); // hand the buffer back (nobody waits for the last two)
    }
}

// One output pixel (dilated row j of tile column c) out of a staged tile: the eight taps one by one.
template< int STEP, bool ITER0, bool LAST >
ASVGF_DEVICE void atrousSingleUnit( const AtrousArgs& a, const Uniform& u, const float4* __restrict__ sA, const float4* __restrict__ sB, const float4* __restrict__ sC,
                                    const float4* __restrict__ sD, const AtrousWsTile t, int j, int c, const uint2 cen )
{
    constexpr int HALO = atrousWsHalo( STEP ), GP = atrousWsGroupPitch( STEP ), RS = 4 * GP;
    const int     sep  = u.W / 2;
    const int px = t.xBase + c, py = t.yBase + j * STEP;
    if( px >= u.W || py >= u.yEnd ) return;
    const int x = c + HALO; // tile column
    int       sx[ 3 ];      // in-row slots of columns x-STEP, x, x+STEP
    sx[ 1 ] = ( x & 3 ) * GP + ( x >> 2 );
    if( STEP >= 4 ) { sx[ 0 ] = sx[ 1 ] - STEP / 4; sx[ 2 ] = sx[ 1 ] + STEP / 4; }
    else { sx[ 0 ] = ( ( x - STEP ) & 3 ) * GP + ( ( x - STEP ) >> 2 ); sx[ 2 ] = ( ( x + STEP ) & 3 ) * GP + ( ( x + STEP ) >> 2 ); }
    const int row = ( j + 1 ) * RS, ce = row + sx[ 1 ];

    AtrousComposeRaw raw;
    if( LAST ) raw = atrousComposeLoad( a, u, px, py, true ); // in flight behind the taps

    AtrousPx P;
    {
        const float4 cA = sA[ ce ], cB = sB[ ce ], cC = sC[ ce ], cD = sD[ ce ];
        if( cA.w < 0.0f || cA.w > 10000.0f )
        {
            // Iter0.comp:208-216 writes zeros, CmSVGFAtrous.comp:96-99 passes the centre through
            P.oD = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : make_float3( cB.x, cB.y, cB.z ); P.oV = ITER0 ? 0.0f : cB.w;
            P.oS = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : make_float3( cC.x, cC.y, cC.z );
            P.oI = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : make_float3( cD.x, cD.y, cD.z );
            P.roughness = cD.w;
        }
        else
        {
            float sigma;
            if( ITER0 ) // Iter0.comp:158-187: 3x3 Gaussian {1/4,1/8,1/16} of the variance (products by powers of two are exact)
            {
                float r = 0.0f;
#pragma unroll
                for( int yy = -1; yy <= 1; yy++ )
#pragma unroll
                    for( int xx = -1; xx <= 1; xx++ )
                        r = fmaf( sB[ row + yy * RS + sx[ xx + 1 ] ].w, ( xx == 0 ? 0.5f : 0.25f ) * ( yy == 0 ? 0.5f : 0.25f ), r );
                sigma = __fsqrt_rn( fmaxf( r, 0.0f ) );
                *atW< uint16_t >( a.filteredVariance, px, py ) = f2h( sigma );
            }
            else sigma = h2f( ( uint16_t )( cen.x >> 16 ) );
            atrousWsSetup( P, cA, cB, cC, cD, cen, sigma, a.iteration );
            const int  ax0 = px >= sep ? sep : 0, ax1 = ax0 + sep;
            const bool okL = px - STEP >= ax0, okR = px + STEP < ax1, okT = py - STEP >= 0, okB = py + STEP < u.H;
#pragma unroll
            for( int yy = -1; yy <= 1; yy++ )
#pragma unroll
                for( int xx = -1; xx <= 1; xx++ )
                {
                    if( xx == 0 && yy == 0 ) continue;
                    const int    qe = row + yy * RS + sx[ xx + 1 ];
                    const float4 qA = sA[ qe ], qB = sB[ qe ], qC = sC[ qe ], qD = sD[ qe ];
                    const bool   in = ( xx < 0 ? okL : xx > 0 ? okR : true ) && ( yy < 0 ? okT : yy > 0 ? okB : true );
                    if( xx != 0 && yy != 0 ) atrousWsTap< true, true >( P, qA, qB, qC, qD, in );
                    else atrousWsTap< false, true >( P, qA, qB, qC, qD, in );
                }
            atrousWsNormalise( P );
        }
    }
    if( !LAST ) atrousWsStore( a, px, py, P );
    else atrousCompose( a, u, px, py, P, raw );
}

// ------------------------------------------------------------------------------------------------------------
// k_atrous_t2: the same staged tile (4-pixel loads, quarter-row slots) and the same pair filter, but one tile per CTA
// and every thread in both roles: stage, __syncthreads, filter.  Two CTAs per SM overlap one's staging with the
// other's arithmetic.
// ------------------------------------------------------------------------------------------------------------
#ifndef ASVGF_T2_THREADS
    #define ASVGF_T2_THREADS 256
#endif
constexpr int kT2Threads = ASVGF_T2_THREADS;
#ifndef ASVGF_T2_MIN_CTAS
    #define ASVGF_T2_MIN_CTAS 2
#endif
constexpr size_t atrousT2SmemBytes( int step, int tileH ) { return ( size_t )( 4 * atrousWsGroupPitch( step ) ) * ( tileH + 2 ) * kAtBytesPerTilePixel; }
// "t3" = the t2 staging with 64 x 8 tiles and ONE pixel per tap loop at 64 registers: four CTAs (32 warps) per SM
constexpr int kT3TileH = 8, kT3MinCtas = 4;

// TH = dilated rows per tile, PAIR = a thread filters a vertical pixel pair (else single pixels), CTAS = CTAs per SM
template< int STEP, bool ITER0, bool LAST, int TH, bool PAIR, int CTAS >
__global__ void __launch_bounds__( kT2Threads, CTAS ) k_atrous_t2( const __grid_constant__ AtrousArgs a, const __grid_constant__ Uniform u )
{
    extern __shared__ __align__( 16 ) uint8_t smemRaw[];
    constexpr int HALO  = atrousWsHalo( STEP );
    constexpr int G     = atrousWsGroups( STEP );
    constexpr int GP    = atrousWsGroupPitch( STEP );
    constexpr int RS    = 4 * GP;
    constexpr int TROWS = TH + 2;
    constexpr int NPX   = RS * TROWS;
    constexpr int TUNITS = TROWS * G;
    constexpr int NR    = ( TUNITS + kT2Threads - 1 ) / kT2Threads;
    constexpr int UNITS = PAIR ? kAtTileW * ( TH / 2 ) : kAtTileW * TH;
    constexpr int NU    = ( UNITS + kT2Threads - 1 ) / kT2Threads;
    constexpr int PER   = PAIR ? 2 : 1; // output pixels per unit
    float4* sA = reinterpret_cast< float4* >( smemRaw ); // normal.xyz, depth
    float4* sB = sA + NPX;                               // diffuse rgb, variance
    float4* sC = sB + NPX;                               // specular rgb, luminance of the diffuse rgb
    float4* sD = sC + NPX;                               // indirect rgb, roughness

    const int band = blockIdx.y / STEP, res = blockIdx.y - band * STEP;
    const AtrousWsTile t{ ( int )blockIdx.x * kAtTileW, u.yBegin + band * ( TH * STEP ) + res };
    const Rows rows = rowsOf( a.depth );
    const int  tid  = threadIdx.x;

    // centre-only inputs of this thread's output pixels (fp16 bits), in flight behind the staging
    uint2 cen[ NU ][ PER ];
    int   ujp[ NU ], uc[ NU ];
#pragma unroll
    for( int k = 0; k < NU; k++ )
    {
        const int unit = tid + k * kT2Threads;
        const int jp = unit / kAtTileW, c0 = unit - jp * kAtTileW;
        // lanes of a quarter warp take columns 4 apart (same quarter row of the tile, consecutive slots): every LDS.128
        // of a tap is conflict-free at any stride; the warp still covers 32 consecutive columns
        const int c = ( c0 & ~31 ) + ( ( c0 & 7 ) << 2 ) + ( ( c0 >> 3 ) & 3 );
        ujp[ k ] = jp; uc[ k ] = c;
#pragma unroll
        for( int h = 0; h < PER; h++ )
        {
            const int  px = t.xBase + c, py = t.yBase + ( PER * jp + h ) * STEP;
            const bool ok = unit < UNITS && px < u.W && py < u.yEnd;
            const uint32_t gd = ldIf< uint16_t >( a.depthGrad, px, py, ok, ( uint16_t )0 );
            const uint32_t sg = ITER0 ? 0u : ( uint32_t )ldIf< uint16_t >( a.filteredVariance, px, py, ok, ( uint16_t )0 );
            const uint32_t hl = ok ? ( uint32_t )__ldg( reinterpret_cast< const uint16_t* >( at< uint2 >( a.histLen, px, py ) ) + 2 ) : 0u; // .z = specular history length
            cen[ k ][ h ] = make_uint2( gd | ( sg << 16 ), hl );
        }
    }

    // stage: one lane moves four pixels per load
    {
        uint4  rn[ NR ], rd0[ NR ], rd1[ NR ], rs[ NR ], ri[ NR ];
        uint2  rz[ NR ], rm[ NR ];
#pragma unroll
        for( int k = 0; k < NR; k++ )
        {
            const int  un = tid + k * kT2Threads;
            const int  r = un / G, g = un - r * G;
            const int  gx = t.xBase - HALO + 4 * g, gy = t.yBase + ( r - 1 ) * STEP;
            const bool ok = un < TUNITS && rows.in( gx, gy ); // W is a multiple of 4: a group is inside or outside as a whole
            const uint4 z4 = make_uint4( 0u, 0u, 0u, 0u );
            const uint2 z2 = make_uint2( 0u, 0u );
            rn[ k ]  = ok ? __ldg( reinterpret_cast< const uint4* >( at< uint32_t >( a.normal, gx, gy ) ) ) : z4;
            rz[ k ]  = ok ? __ldg( reinterpret_cast< const uint2* >( at< uint16_t >( a.depth, gx, gy ) ) ) : z2;
            rm[ k ]  = ok ? __ldg( reinterpret_cast< const uint2* >( at< uint16_t >( a.metRough, gx, gy ) ) ) : z2;
            rd0[ k ] = ok ? __ldg( reinterpret_cast< const uint4* >( at< uint2 >( a.srcDiff, gx, gy ) ) ) : z4;
            rd1[ k ] = ok ? __ldg( reinterpret_cast< const uint4* >( at< uint2 >( a.srcDiff, gx, gy ) ) + 1 ) : z4;
            rs[ k ]  = ok ? __ldg( reinterpret_cast< const uint4* >( at< uint32_t >( a.srcSpec, gx, gy ) ) ) : z4;
            ri[ k ]  = ok ? __ldg( reinterpret_cast< const uint4* >( at< uint32_t >( a.srcIndir, gx, gy ) ) ) : z4;
        }
#pragma unroll
        for( int k = 0; k < NR; k++ )
        {
            const int un = tid + k * kT2Threads;
            if( un >= TUNITS ) break;
            const int r = un / G, g = un - r * G;
            const uint32_t n4[ 4 ] = { rn[ k ].x, rn[ k ].y, rn[ k ].z, rn[ k ].w };
            float2 lp[ 4 ], lt[ 4 ];
#pragma unroll
            for( int q = 0; q < 4; q++ )
            {
                lp[ q ] = __ldg( a.lut.phi + ( n4[ q ] >> 16 ) );
                lt[ q ] = __ldg( a.lut.theta + ( n4[ q ] & 0xffffu ) );
            }
            const uint32_t d4[ 8 ] = { rd0[ k ].x, rd0[ k ].y, rd0[ k ].z, rd0[ k ].w, rd1[ k ].x, rd1[ k ].y, rd1[ k ].z, rd1[ k ].w };
            const uint32_t s4[ 4 ] = { rs[ k ].x, rs[ k ].y, rs[ k ].z, rs[ k ].w }, i4[ 4 ] = { ri[ k ].x, ri[ k ].y, ri[ k ].z, ri[ k ].w };
            const uint32_t z4[ 2 ] = { rz[ k ].x, rz[ k ].y }, m4[ 2 ] = { rm[ k ].x, rm[ k ].y };
#pragma unroll
            for( int q = 0; q < 4; q++ )
            {
                const int    e  = r * RS + q * GP + g;
                const float3 n  = make_float3( mulr( lp[ q ].x, lt[ q ].y ), mulr( lp[ q ].x, lt[ q ].x ), lp[ q ].y ); // decodeNormal, Utils.h:124-140
                const float2 rg = __half22float2( *reinterpret_cast< const __half2* >( &d4[ 2 * q ] ) );
                const float2 bv = __half22float2( *reinterpret_cast< const __half2* >( &d4[ 2 * q + 1 ] ) );
                const float3 sp = decodeE5( s4[ q ] ), in = decodeE5( i4[ q ] );
                const uint32_t zq = ( z4[ q >> 1 ] >> ( 16 * ( q & 1 ) ) ) & 0xffffu, mq = ( m4[ q >> 1 ] >> ( 16 * ( q & 1 ) ) ) & 0xffffu;
                sA[ e ] = make_float4( n.x, n.y, n.z, h2f( ( uint16_t )zq ) );
                sB[ e ] = make_float4( rg.x, rg.y, bv.x, bv.y );
                sC[ e ] = make_float4( sp.x, sp.y, sp.z, luminance( make_float3( rg.x, rg.y, bv.x ) ) );
                sD[ e ] = make_float4( in.x, in.y, in.z, unorm8ToFloatFast( mq >> 8 ) );
            }
        }
    }
    __syncthreads();

#pragma unroll
    for( int k = 0; k < NU; k++ )
    {
        if( tid + k * kT2Threads >= UNITS ) break;
        if constexpr( PAIR ) atrousPairUnit< STEP, ITER0, LAST >( a, u, sA, sB, sC, sD, t, ujp[ k ], uc[ k ], cen[ k ][ 0 ], cen[ k ][ PER - 1 ] );
        else atrousSingleUnit< STEP, ITER0, LAST >( a, u, sA, sB, sC, sD, t, ujp[ k ], uc[ k ], cen[ k ][ 0 ] );
    }
}

} // namespace asvgf
