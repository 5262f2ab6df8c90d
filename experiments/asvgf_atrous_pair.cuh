// asvgf_atrous_pair.cuh -- a-trous kernel, Blackwell packed-fp32 form (FFMA2/FMUL2/FADD2, sm_100a `*.f32x2`).
//
// Same arithmetic and the same dilated tile as k_atrous_tiled (asvgf_atrous_tiled.cuh; reference:
// CmSVGFAtrous_Iter0.comp:158-335, CmSVGFAtrous.comp:65-349).  The pass is bound by instruction issue, not by HBM,
// so this form halves the FMA-pipe instruction count: every thread filters TWO pixels at once -- (c, j) and (c, j+8)
// of the 64x16 dilated tile -- with each quantity held as a packed f32x2 register pair.  The shared-memory tile is
// laid out for it: entry (q, c) interleaves the 16 decoded floats of tile rows q-1 and q+7, so one LDS.128 yields two
// components of both pixels already paired, and the taps of the two pixels sit in the same entry for every stride
// (rows j+yy and j+8+yy are entry q+yy).  Tile rows 7 and 8 are stored twice (11 % more staging).
//   per tap pair: 8 LDS.128 + ~25 packed + ~10 scalar (max/abs) + 8 MUFU  vs  2 x (4 LDS.128 + ~40 + 4 MUFU).
#pragma once

#include "asvgf_atrous_tiled.cuh"

namespace asvgf
{

typedef unsigned long long f32x2; // lo = pixel A (tile row j), hi = pixel B (tile row j+8)

ASVGF_DEVICE f32x2 pk( float a, float b ) { f32x2 r; asm( "mov.b64 %0, {%1,%2};" : "=l"( r ) : "f"( a ), "f"( b ) ); return r; }
ASVGF_DEVICE float2 upk( f32x2 v ) { float2 r; asm( "mov.b64 {%0,%1}, %2;" : "=f"( r.x ), "=f"( r.y ) : "l"( v ) ); return r; }
ASVGF_DEVICE f32x2 fma2( f32x2 a, f32x2 b, f32x2 c ) { f32x2 d; asm( "fma.rn.f32x2 %0, %1, %2, %3;" : "=l"( d ) : "l"( a ), "l"( b ), "l"( c ) ); return d; }
ASVGF_DEVICE f32x2 mul2( f32x2 a, f32x2 b ) { f32x2 d; asm( "mul.rn.f32x2 %0, %1, %2;" : "=l"( d ) : "l"( a ), "l"( b ) ); return d; }
ASVGF_DEVICE f32x2 add2( f32x2 a, f32x2 b ) { f32x2 d; asm( "add.rn.f32x2 %0, %1, %2;" : "=l"( d ) : "l"( a ), "l"( b ) ); return d; }
ASVGF_DEVICE f32x2 sub2( f32x2 a, f32x2 b ) { f32x2 d; asm( "sub.rn.f32x2 %0, %1, %2;" : "=l"( d ) : "l"( a ), "l"( b ) ); return d; }

constexpr int kApPairRows = kAtTileH / 2 + 2; // entries per column: rows (q-1, q+7), q = 0..9
constexpr int kApHalf     = kAtTileH / 2;

struct AtrousDecoded // the 16 floats of one tile pixel
{
    float3 n; float depth; float3 diff; float var; float3 spec; float lum; float3 indir; float rough;
};

ASVGF_DEVICE AtrousDecoded atrousDecode( const NormalLut& lut, uint32_t encN, uint32_t dbits, uint32_t mr, uint2 dv, uint32_t es, uint32_t ei )
{
    AtrousDecoded o;
    o.n     = decodeNormal( lut, encN );
    o.depth = h2f( ( uint16_t )dbits );
    const float2 rg = __half22float2( *reinterpret_cast< const __half2* >( &dv.x ) );
    const float2 bv = __half22float2( *reinterpret_cast< const __half2* >( &dv.y ) );
    o.diff  = make_float3( rg.x, rg.y, bv.x );
    o.var   = bv.y;
    o.spec  = decodeE5( es );
    o.lum   = luminance( o.diff );
    o.indir = decodeE5( ei );
    o.rough = unorm8ToFloatFast( mr >> 8 );
    return o;
}

// lighting composition of one pixel, CmSVGFAtrous.comp:278-348
ASVGF_DEVICE void atrousCompose( const AtrousArgs& a, const Uniform& u, int px, int py, float3 outDiff, float3 outSpec, float3 outIndir, float roughness, uint32_t albedoBits,
                                 uint2 tp, uint32_t skyMetal )
{
    const int    sep   = u.W / 2;
    const int    isOdd = px >= sep ? 1 : 0;
    const int    regX  = ( px - isOdd * sep ) * 2 + ( ( isOdd + py ) & 1 ); // ShaderCommonGLSLFunc.h:331-342
    float3       albedo = unpackUfloat111110Fast( albedoBits );
    const float2 tp01   = __half22float2( *reinterpret_cast< const __half2* >( &tp.x ) );
    const float3 throughput = make_float3( tp01.x, tp01.y, __low2float( *reinterpret_cast< const __half2* >( &tp.y ) ) );
    if( ( u.debugShowFlags & ( 1u << 9 ) ) != 0 ) albedo = make_float3( 1.0f, 1.0f, 1.0f );

    float3 illuminated;
    if( ( skyMetal & 0xffu ) == 0u )
    {
        const float metallic = unorm8ToFloatFast( skyMetal >> 8 );
        const float t  = saturatef( ( roughness - 0.5f ) * 4.0f ); // smoothstep(0.5, 0.75, roughness)
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
    *atW< uint32_t >( a.preFinal, regX, py ) = packUfloat111110Fast( make_float3( fmaxf( 0.0f, illuminated.x ), fmaxf( 0.0f, illuminated.y ), fmaxf( 0.0f, illuminated.z ) ) );
}

template< int STEP, bool ITER0, bool LAST >
__global__ void __launch_bounds__( kAtThreads, 2 ) k_atrous_pair( const __grid_constant__ AtrousArgs a, const __grid_constant__ Uniform u )
{
    extern __shared__ __align__( 16 ) uint8_t smemRaw[];
    constexpr int TW = kAtTileW + 2 * STEP; // tile width incl. halo columns
    constexpr int NE = TW * kApPairRows;    // entries
    // plane i holds, per entry, (c[2i].A, c[2i].B, c[2i+1].A, c[2i+1].B) of the component list
    //   nx ny | nz depth | diff.r diff.g | diff.b var | spec.r spec.g | spec.b lum | indir.r indir.g | indir.b rough
    float4* pl = reinterpret_cast< float4* >( smemRaw );

    const int band  = blockIdx.y / STEP, res = blockIdx.y - band * STEP;
    const int yBase = u.yBegin + band * ( kAtTileH * STEP ) + res;
    const int xBase = blockIdx.x * kAtTileW;
    const Rows rows = rowsOf( a.depth );

    // ---- centre-only inputs of this thread's two pixel pairs: issued first so their latency hides behind the staging ----
    constexpr int PAIRS = ( kAtTileW * kApHalf ) / kAtThreads; // 2
    uint16_t cGradDepth[ PAIRS ][ 2 ], cHistLenSpec[ PAIRS ][ 2 ], cSigma[ PAIRS ][ 2 ];
    uint32_t cAlbedo[ PAIRS ][ 2 ], cSkyMetal[ PAIRS ][ 2 ];
    uint2    cThroughput[ PAIRS ][ 2 ];
#pragma unroll
    for( int k = 0; k < PAIRS; k++ )
#pragma unroll
        for( int h = 0; h < 2; h++ )
        {
            const int  p  = threadIdx.x + k * kAtThreads;
            const int  j  = p / kAtTileW + h * kApHalf, c = p % kAtTileW;
            const int  px = xBase + c, py = yBase + j * STEP;
            const bool ok = px < u.W && py < u.yEnd;
            cGradDepth[ k ][ h ]   = ldIf< uint16_t >( a.depthGrad, px, py, ok, ( uint16_t )0 );
            cHistLenSpec[ k ][ h ] = ( uint16_t )( ldIf< uint2 >( a.histLen, px, py, ok, make_uint2( 0u, 0u ) ).y & 0xffffu ); // .z of the rgba16f texel
            if( !ITER0 ) cSigma[ k ][ h ] = ldIf< uint16_t >( a.filteredVariance, px, py, ok, ( uint16_t )0 );
            if( LAST )
            {
                const int sepx = u.W / 2, isOdd = px >= sepx ? 1 : 0;
                const int regX = ( px - isOdd * sepx ) * 2 + ( ( isOdd + py ) & 1 );
                cAlbedo[ k ][ h ]     = ldIf< uint32_t >( a.albedo, regX, py, ok, 0u );
                cThroughput[ k ][ h ] = ldIf< uint2 >( a.throughput, px, py, ok, make_uint2( 0u, 0u ) );
                cSkyMetal[ k ][ h ]   = ( uint32_t )ldIf< uint8_t >( a.isSky, px, py, ok, ( uint8_t )0 ) | ( ( ( uint32_t )ldIf< uint16_t >( a.metRough, px, py, ok, ( uint16_t )0 ) & 0xffu ) << 8 );
            }
        }

    // ---- stage the tile: every entry unpacks the two pixels it interleaves ----
#pragma unroll
    for( int it = 0; it < ( NE + kAtThreads - 1 ) / kAtThreads; it++ )
    {
        const int e = threadIdx.x + it * kAtThreads;
        if( e >= NE ) break;
        const int q = e / TW, c = e - q * TW;
        const int gx = xBase - STEP + c;
        AtrousDecoded d[ 2 ];
#pragma unroll
        for( int h = 0; h < 2; h++ )
        {
            const int  gy = yBase + ( q - 1 + h * kApHalf ) * STEP;
            const bool ok = rows.in( gx, gy );
            d[ h ] = atrousDecode( a.lut, ldIf< uint32_t >( a.normal, gx, gy, ok, 0u ), ldIf< uint16_t >( a.depth, gx, gy, ok, ( uint16_t )0 ),
                                   ldIf< uint16_t >( a.metRough, gx, gy, ok, ( uint16_t )0 ), ldIf< uint2 >( a.srcDiff, gx, gy, ok, make_uint2( 0u, 0u ) ),
                                   ldIf< uint32_t >( a.srcSpec, gx, gy, ok, 0u ), ldIf< uint32_t >( a.srcIndir, gx, gy, ok, 0u ) );
        }
        pl[ 0 * NE + e ] = make_float4( d[ 0 ].n.x, d[ 1 ].n.x, d[ 0 ].n.y, d[ 1 ].n.y );
        pl[ 1 * NE + e ] = make_float4( d[ 0 ].n.z, d[ 1 ].n.z, d[ 0 ].depth, d[ 1 ].depth );
        pl[ 2 * NE + e ] = make_float4( d[ 0 ].diff.x, d[ 1 ].diff.x, d[ 0 ].diff.y, d[ 1 ].diff.y );
        pl[ 3 * NE + e ] = make_float4( d[ 0 ].diff.z, d[ 1 ].diff.z, d[ 0 ].var, d[ 1 ].var );
        pl[ 4 * NE + e ] = make_float4( d[ 0 ].spec.x, d[ 1 ].spec.x, d[ 0 ].spec.y, d[ 1 ].spec.y );
        pl[ 5 * NE + e ] = make_float4( d[ 0 ].spec.z, d[ 1 ].spec.z, d[ 0 ].lum, d[ 1 ].lum );
        pl[ 6 * NE + e ] = make_float4( d[ 0 ].indir.x, d[ 1 ].indir.x, d[ 0 ].indir.y, d[ 1 ].indir.y );
        pl[ 7 * NE + e ] = make_float4( d[ 0 ].indir.z, d[ 1 ].indir.z, d[ 0 ].rough, d[ 1 ].rough );
    }
    __syncthreads();

    const int       sep   = u.W / 2;
    constexpr float LOG2E = 1.4426950408889634f;
    const float     NINF  = __uint_as_float( 0xff800000u );

#pragma unroll
    for( int k = 0; k < PAIRS; k++ )
    {
        const int p  = threadIdx.x + k * kAtThreads;
        const int j  = p / kAtTileW, c = p - j * kAtTileW;
        const int px = xBase + c;
        const int py[ 2 ] = { yBase + j * STEP, yBase + ( j + kApHalf ) * STEP };
        if( px >= u.W || py[ 0 ] >= u.yEnd ) continue; // row A is the upper one: nothing to do when it is outside
        const bool okB = py[ 1 ] < u.yEnd;

        const int    ce = ( j + 1 ) * TW + ( c + STEP );
        const float4 c0 = pl[ 0 * NE + ce ], c1 = pl[ 1 * NE + ce ], c2 = pl[ 2 * NE + ce ], c3 = pl[ 3 * NE + ce ];
        const float4 c4 = pl[ 4 * NE + ce ], c5 = pl[ 5 * NE + ce ], c6 = pl[ 6 * NE + ce ], c7 = pl[ 7 * NE + ce ];
        const f32x2  cnx = pk( c0.x, c0.y ), cny = pk( c0.z, c0.w ), cnz = pk( c1.x, c1.y ), cdepth = pk( c1.z, c1.w );
        const f32x2  clum = pk( c5.z, c5.w ), crough = pk( c7.z, c7.w );
        const float  depth[ 2 ] = { c1.z, c1.w }, rough[ 2 ] = { c7.z, c7.w };

        // per-pixel scalars of the weights (centre only)
        float kC1[ 2 ], kC2[ 2 ], nkLum[ 2 ], Pw[ 2 ], wRM[ 2 ], c10[ 2 ], lky[ 2 ][ 3 ];
#pragma unroll
        for( int h = 0; h < 2; h++ )
        {
            const float gradDepth = h2f( cGradDepth[ k ][ h ] );
            float sigmaL;
            if( ITER0 ) // Iter0.comp:158-187: 3x3 Gaussian {1/4,1/8,1/16} of the variance (products by powers of two are exact)
            {
                float r = 0.0f;
#pragma unroll
                for( int yy = -1; yy <= 1; yy++ )
#pragma unroll
                    for( int xx = -1; xx <= 1; xx++ )
                    {
                        const float4 v = pl[ 3 * NE + ce + yy * TW + xx ];
                        r = fmaf( h == 0 ? v.z : v.w, ( xx == 0 ? 0.5f : 0.25f ) * ( yy == 0 ? 0.5f : 0.25f ), r );
                    }
                r = __fsqrt_rn( fmaxf( r, 0.0f ) );
                const bool valid = !( depth[ h ] < 0.0f || depth[ h ] > 10000.0f );
                if( valid && ( h == 0 || okB ) ) *atW< uint16_t >( a.filteredVariance, px, py[ h ] ) = f2h( r );
                sigmaL = r;
            }
            else
            {
                sigmaL = h2f( cSigma[ k ][ h ] );
            }
            nkLum[ h ] = -LOG2E * frcp( fmaf( 4.0f, sigmaL, 0.00001f ) );
            wRM[ h ]   = saturatef( fmaf( rough[ h ], 30.0f, -( float )a.iteration ) );
            c10[ h ]   = -10.0f * wRM[ h ];
            const float alpha = rough[ h ] * rough[ h ];
            float       P     = fmaxf( 0.01f, fmaf( 2.0f, frcp( fmaf( alpha, alpha, 1e-4f ) ), -2.0f ) ); // BRDF.h:27-29
            Pw[ h ]           = clampf( P, 8.0f, 1024.0f ) * saturatef( h2f( cHistLenSpec[ k ][ h ] ) * 0.125f );
            const float i1 = frcp( fmaxf( gradDepth, 0.01f ) ), i2 = frcp( fmaxf( gradDepth * 2.0f, 0.01f ) );
            kC1[ h ] = -LOG2E * i1 * i1; // exp(-wz^2) = ex2( dz^2 * kC )
            kC2[ h ] = -LOG2E * i2 * i2;
            lky[ h ][ 0 ] = py[ h ] - STEP >= 0 ? -1.0f : NINF; lky[ h ][ 1 ] = 0.0f; lky[ h ][ 2 ] = py[ h ] + STEP < u.H ? -1.0f : NINF;
        }
        // log2 of the wavelet weight per tap, -inf for taps outside the checkerboarded render area
        // (ShaderCommonGLSLFunc.h:345-366): taps never cross x = W/2 or leave the frame
        const int   ax0 = px >= sep ? sep : 0, ax1 = ax0 + sep;
        const float lkx[ 3 ] = { px - STEP >= ax0 ? -1.0f : NINF, 0.0f, px + STEP < ax1 ? -1.0f : NINF };
        const f32x2 kC1p = pk( kC1[ 0 ], kC1[ 1 ] ), kC2p = pk( kC2[ 0 ], kC2[ 1 ] ), Pp = pk( Pw[ 0 ], Pw[ 1 ] ), c128 = pk( 128.0f, 128.0f );

        f32x2 aD0 = pk( c2.x, c2.y ), aD1 = pk( c2.z, c2.w ), aD2 = pk( c3.x, c3.y ), aV = pk( c3.z, c3.w );
        f32x2 aS0 = pk( c4.x, c4.y ), aS1 = pk( c4.z, c4.w ), aS2 = pk( c5.x, c5.y );
        f32x2 aI0 = pk( c6.x, c6.y ), aI1 = pk( c6.z, c6.w ), aI2 = pk( c7.x, c7.y );
        f32x2 sumD = pk( 1.0f, 1.0f ), sumS = sumD, sumI = sumD;

        // sky / invalid-depth pixels take no taps (Iter0.comp:208-216, CmSVGFAtrous.comp:96-99): skip when both are
        const bool bothInvalid = ( depth[ 0 ] < 0.0f || depth[ 0 ] > 10000.0f ) && ( !okB || depth[ 1 ] < 0.0f || depth[ 1 ] > 10000.0f );
        if( !bothInvalid )
#pragma unroll
        for( int yy = -1; yy <= 1; yy++ )
        {
#pragma unroll
            for( int xx = -1; xx <= 1; xx++ )
            {
                if( xx == 0 && yy == 0 ) continue;
                const int    qe = ce + yy * TW + xx * STEP;
                const float4 t0 = pl[ 0 * NE + qe ], t1 = pl[ 1 * NE + qe ], t2 = pl[ 2 * NE + qe ], t3 = pl[ 3 * NE + qe ];
                const float4 t4 = pl[ 4 * NE + qe ], t5 = pl[ 5 * NE + qe ], t6 = pl[ 6 * NE + qe ], t7 = pl[ 7 * NE + qe ];

                const float2 dot = upk( fma2( cnz, pk( t1.x, t1.y ), fma2( cny, pk( t0.z, t0.w ), mul2( cnx, pk( t0.x, t0.y ) ) ) ) );
                const f32x2  L   = pk( fmaxf( flg2( fmaxf( dot.x, 0.0f ) ), -126.0f ), fmaxf( flg2( fmaxf( dot.y, 0.0f ) ), -126.0f ) );
                const f32x2  dz  = sub2( cdepth, pk( t1.z, t1.w ) );
                const f32x2  e0  = fma2( c128, L, pk( lkx[ xx + 1 ] + lky[ 0 ][ yy + 1 ], lkx[ xx + 1 ] + lky[ 1 ][ yy + 1 ] ) );
                const f32x2  eBp = fma2( mul2( dz, dz ), ( xx != 0 && yy != 0 ) ? kC2p : kC1p, e0 );
                const float2 eB  = upk( eBp );
                const float2 dl  = upk( sub2( clum, pk( t5.z, t5.w ) ) );
                const f32x2  wB  = pk( fex2( eB.x ), fex2( eB.y ) );
                const f32x2  wD  = pk( fex2( fmaf( fabsf( dl.x ), nkLum[ 0 ], eB.x ) ), fex2( fmaf( fabsf( dl.y ), nkLum[ 1 ], eB.y ) ) );
                const float2 dr  = upk( sub2( crough, pk( t7.z, t7.w ) ) );
                const float2 eS  = upk( fma2( Pp, L, eBp ) );
                const f32x2  wS  = mul2( pk( fmaxf( fmaf( fabsf( dr.x ), c10[ 0 ], wRM[ 0 ] ), 0.0f ), fmaxf( fmaf( fabsf( dr.y ), c10[ 1 ], wRM[ 1 ] ), 0.0f ) ),
                                         pk( fex2( eS.x ), fex2( eS.y ) ) );

                aD0 = fma2( pk( t2.x, t2.y ), wD, aD0 ); aD1 = fma2( pk( t2.z, t2.w ), wD, aD1 ); aD2 = fma2( pk( t3.x, t3.y ), wD, aD2 );
                aV  = fma2( mul2( pk( t3.z, t3.w ), wD ), wD, aV ); // (var*w)*w, like the shader
                aS0 = fma2( pk( t4.x, t4.y ), wS, aS0 ); aS1 = fma2( pk( t4.z, t4.w ), wS, aS1 ); aS2 = fma2( pk( t5.x, t5.y ), wS, aS2 );
                aI0 = fma2( pk( t6.x, t6.y ), wB, aI0 ); aI1 = fma2( pk( t6.z, t6.w ), wB, aI1 ); aI2 = fma2( pk( t7.x, t7.y ), wB, aI2 );
                sumD = add2( sumD, wD ); sumS = add2( sumS, wS ); sumI = add2( sumI, wB );
            }
        }
        const float2 sD = upk( sumD ), sS = upk( sumS ), sI = upk( sumI );
        const f32x2  invD = pk( frcp( sD.x ), frcp( sD.y ) ), invS = pk( frcp( sS.x ), frcp( sS.y ) ), invI = pk( frcp( sI.x ), frcp( sI.y ) );
        const float2 oD0 = upk( mul2( aD0, invD ) ), oD1 = upk( mul2( aD1, invD ) ), oD2 = upk( mul2( aD2, invD ) ), oV = upk( mul2( aV, mul2( invD, invD ) ) );
        const float2 oS0 = upk( mul2( aS0, invS ) ), oS1 = upk( mul2( aS1, invS ) ), oS2 = upk( mul2( aS2, invS ) );
        const float2 oI0 = upk( mul2( aI0, invI ) ), oI1 = upk( mul2( aI1, invI ) ), oI2 = upk( mul2( aI2, invI ) );

#pragma unroll
        for( int h = 0; h < 2; h++ )
        {
            if( h == 1 && !okB ) break;
            float3 outDiff  = h == 0 ? make_float3( oD0.x, oD1.x, oD2.x ) : make_float3( oD0.y, oD1.y, oD2.y );
            float3 outSpec  = h == 0 ? make_float3( oS0.x, oS1.x, oS2.x ) : make_float3( oS0.y, oS1.y, oS2.y );
            float3 outIndir = h == 0 ? make_float3( oI0.x, oI1.x, oI2.x ) : make_float3( oI0.y, oI1.y, oI2.y );
            float  outVariance = h == 0 ? oV.x : oV.y;
            if( depth[ h ] < 0.0f || depth[ h ] > 10000.0f ) // invalid depth: Iter0.comp:208-216 zeroes, CmSVGFAtrous.comp:96-99 passes the centre through
            {
                outDiff  = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : ( h == 0 ? make_float3( c2.x, c2.z, c3.x ) : make_float3( c2.y, c2.w, c3.y ) );
                outSpec  = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : ( h == 0 ? make_float3( c4.x, c4.z, c5.x ) : make_float3( c4.y, c4.w, c5.y ) );
                outIndir = ITER0 ? make_float3( 0.0f, 0.0f, 0.0f ) : ( h == 0 ? make_float3( c6.x, c6.z, c7.x ) : make_float3( c6.y, c6.w, c7.y ) );
                outVariance = ITER0 ? 0.0f : ( h == 0 ? c3.z : c3.w );
            }
            if( !LAST )
            {
                const __half2 o0 = __floats2half2_rn( outDiff.x, outDiff.y ), o1 = __floats2half2_rn( outDiff.z, outVariance );
                *atW< uint2 >( a.dstDiff, px, py[ h ] ) = make_uint2( *reinterpret_cast< const uint32_t* >( &o0 ), *reinterpret_cast< const uint32_t* >( &o1 ) );
                *atW< uint32_t >( a.dstSpec, px, py[ h ] )  = encodeE5Fast( outSpec );
                *atW< uint32_t >( a.dstIndir, px, py[ h ] ) = encodeE5Fast( outIndir );
            }
            else
            {
                atrousCompose( a, u, px, py[ h ], outDiff, outSpec, outIndir, rough[ h ], cAlbedo[ k ][ h ], cThroughput[ k ][ h ], cSkyMetal[ k ][ h ] );
            }
        }
    }
}

constexpr size_t atrousPairSmemBytes( int step ) { return ( size_t )( kAtTileW + 2 * step ) * kApPairRows * 128; }

} // namespace asvgf
