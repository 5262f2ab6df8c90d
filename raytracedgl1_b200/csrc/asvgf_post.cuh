// asvgf_post.cuh -- the two dispatches that consume the denoiser's outputs, sm_100a.
//
//   k_checkerboard_resolve ... CmCheckerboard.comp:43-78 (PreFinal -> Final)
//   k_lum_bin_table .......... getColorBin of CmLuminanceHistogram.comp:38-55 for every fp16 bit pattern
//   k_lum_histogram .......... CmLuminanceHistogram.comp:61-86
//   k_lum_average ............ CmLuminanceAvg.comp:37-89
//
// Both are integer/bit-exact against the oracle.  The histogram is HBM-bound (2 B per pixel): HistogramInput is fp16, so
// the bin of a texel is a function of its 16 bits; a 64 KB table of bins (rebuilt per call from the uniform's
// log-luminance range and a correctly rounded log2 table) sits in shared memory and a pixel costs one byte lookup and
// one warp-aggregated shared-memory atomic.
#pragma once

#include "../../include/asvgf_layout.h"
#include "asvgf_atrous_tiled.cuh"

namespace asvgf
{

struct ResolveArgs
{
    Img     src[ 3 ], dst[ 3 ]; // PreFinal -> Final, ScreenEmisRT -> ScreenEmission, AcidFogRT -> AcidFog (CmCheckerboard.comp:62-77)
    Img     throughput;
    int32_t pairs;              // 3, or 1 when only the denoiser's own image is resolved
};

// one thread per REGULAR pixel; arithmetic in the shader's order (products by 0.25 / 0.5 are exact, the three sums round)
__global__ void __launch_bounds__( kBlockW* kBlockH ) k_checkerboard_resolve( const __grid_constant__ ResolveArgs a, const __grid_constant__ Uniform u )
{
    const int x = blockIdx.x * kBlockW + threadIdx.x;
    const int y = u.yBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( x >= u.W || y >= u.yEnd ) return;
    const Rows rows = rowsOf( a.src[ 0 ] );
    const int  sep  = u.W / 2;
    const int  isOdd = ( x + y % 2 ) % 2;                 // ShaderCommonGLSLFunc.h:311-329
    const int  cbx   = isOdd * sep + x / 2;
    const uint16_t split  = __ldg( reinterpret_cast< const uint16_t* >( at< uint2 >( a.throughput, cbx, y ) ) + 3 ); // .a
    const bool     resolve = h2f( split ) > 0.5f;                       // wasSplit, ShaderCommonGLSLFunc.h:434-444
    const float    fl = ( cbx == 0 || cbx == sep ) ? 0.0f : 1.0f;       // !isLeftOutOfBounds
    const float    fr = ( cbx == sep || cbx == u.W ) ? 0.0f : 1.0f;     // !isRightOutOfBounds
    for( int i = 0; i < a.pairs; i++ )
    {
        const Img&     src    = a.src[ i ];
        const uint32_t centre = __ldg( at< uint32_t >( src, x, y ) );
        uint32_t       out    = centre;
        if( resolve )
        {
            const float3 c = unpackUfloat111110( centre );
            const float3 l = unpackUfloat111110( ldIf< uint32_t >( src, x - 1, y, rows.in( x - 1, y ), 0u ) );
            const float3 r = unpackUfloat111110( ldIf< uint32_t >( src, x + 1, y, rows.in( x + 1, y ), 0u ) );
            const float3 d = unpackUfloat111110( ldIf< uint32_t >( src, x, y + 1, rows.in( x, y + 1 ), 0u ) );
            const float3 t = unpackUfloat111110( ldIf< uint32_t >( src, x, y - 1, rows.in( x, y - 1 ), 0u ) );
            auto cross = [ & ]( float vl, float vr, float vd, float vt ) {
                return addr( addr( addr( mulr( mulr( 0.25f, vl ), fl ), mulr( mulr( 0.25f, vr ), fr ) ), mulr( 0.25f, vd ) ), mulr( 0.25f, vt ) );
            };
            const float3 k = make_float3( cross( l.x, r.x, d.x, t.x ), cross( l.y, r.y, d.y, t.y ), cross( l.z, r.z, d.z, t.z ) );
            out = packUfloat111110( make_float3( mixr( c.x, k.x, 0.5f ), mixr( c.y, k.y, 0.5f ), mixr( c.z, k.z, 0.5f ) ) );
        }
        *atW< uint32_t >( a.dst[ i ], x, y ) = out;
    }
}

// ---- exposure ----
struct LumArgs
{
    Img              histogramInput;
    const float*     log2Half;   // 65536 entries: correctly rounded log2 of every fp16 bit pattern (host-built)
    uint8_t*         binTable;   // 65536 entries, rebuilt per call
    AsvgfTonemapping* tm;
    float            minLogLuminance, maxLogLuminance, timeDelta, pixelCount;
    uint32_t         stopEyeAdaptation;
};

// getColorBin (CmLuminanceHistogram.comp:38-55) of every fp16 bit pattern
__global__ void __launch_bounds__( 256 ) k_lum_bin_table( const __grid_constant__ LumArgs a )
{
    const uint32_t h = blockIdx.x * 256u + threadIdx.x;
    if( h >= 65536u ) return;
    const float lum = h2f( ( uint16_t )h );
    uint32_t    bin;
    if( lum < 0.001f ) bin = 0u; // LUMINANCE_EPS; NaN falls through like the shader's comparison
    else
    {
        const float inv = divr( 1.0f, subr( a.maxLogLuminance, a.minLogLuminance ) );
        const float t   = fminf( fmaxf( mulr( subr( __ldg( a.log2Half + h ), a.minLogLuminance ), inv ), 0.0f ), 1.0f ); // clamp = min(max(x,0),1), NaN -> 0
        bin             = ( uint32_t )f2i( addr( mulr( t, 254.0f ), 1.0f ) );
    }
    a.binTable[ h ] = ( uint8_t )bin;
}

constexpr int kLumThreads = 256;

__global__ void __launch_bounds__( kLumThreads ) k_lum_histogram( const __grid_constant__ LumArgs a, const __grid_constant__ Uniform u )
{
#ifndef ASVGF_HOSTSIM
    extern __shared__ __align__( 16 ) uint8_t lumSmem[];
#else
    uint8_t* const lumSmem = hostsim::dynamicSmem();
#endif
    uint8_t*  sTable = lumSmem;                                            // 64 KB
    uint32_t* sHist  = reinterpret_cast< uint32_t* >( lumSmem + 65536 );   // one 256-bin histogram per warp
    constexpr int WARPS = kLumThreads / 32;
    // the 64 KB bin table comes in as ONE bulk asynchronous copy (TMA, cp.async.bulk: no registers, no per-thread loads --
    // ncu of the per-thread copy: the kernel spent most of its 35 us filling this table) signalled through an mbarrier,
    // while the threads zero the per-warp histograms
#ifndef ASVGF_HOSTSIM
    __shared__ __align__( 8 ) uint64_t tableBarrier;
    const uint32_t barAddr = ( uint32_t )__cvta_generic_to_shared( &tableBarrier ), dstAddr = ( uint32_t )__cvta_generic_to_shared( sTable );
    if( threadIdx.x == 0 )
    {
        asm volatile( "mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"( barAddr ) );
        asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" );
        asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( barAddr ), "r"( 65536u ) : "memory" );
        asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"( dstAddr ), "l"( a.binTable ), "r"( 65536u ), "r"( barAddr )
                      : "memory" );
    }
#else // tests/hostsim (host compile of this file by the CPU test suite): the bulk copy is a memcpy
    if( threadIdx.x == 0 ) memcpy( sTable, a.binTable, 65536 );
#endif
    for( int i = threadIdx.x; i < WARPS * 256; i += kLumThreads ) sHist[ i ] = 0u;
    __syncthreads(); // the barrier word is initialised and the histograms are zero
#ifndef ASVGF_HOSTSIM
    {
        uint32_t done = 0;
        while( !done )
            asm volatile( "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"( done ) : "r"( barAddr ) : "memory" );
    }
#endif

    uint32_t* myHist = sHist + ( threadIdx.x >> 5 ) * 256;
    auto count = [ & ]( uint32_t bin, bool valid ) {
#ifndef ASVGF_HOSTSIM
        // lanes with the same bin elect one of them to add their number: no same-address serialisation on flat images
        const uint32_t active = __ballot_sync( 0xffffffffu, valid );
        if( valid )
        {
            const uint32_t peers = __match_any_sync( active, bin );
            if( ( threadIdx.x & 31 ) == ( uint32_t )( __ffs( peers ) - 1 ) ) atomicAdd( myHist + bin, ( uint32_t )__popc( peers ) );
        }
#else // (no warps on the host: every thread adds its own count)
        if( valid ) atomicAdd( myHist + bin, 1u );
#endif
    };

    // rows of 8-texel groups (one uint4 per lane), dealt to the CTAs round-robin
    const int  rows = u.yEnd - u.yBegin;
    const int  gpr  = ( u.W + 7 ) / 8;
    const long total = ( long )rows * gpr;
    const bool vec  = ( a.histogramInput.pitch % 16 ) == 0 && ( ( uintptr_t )a.histogramInput.p0 % 16 ) == 0;
    // four groups per thread in flight (the pass is latency-bound otherwise: ncu showed 22 % issue, 6 % DRAM with one)
    constexpr int UNR = 4;
    const long stride = ( long )gridDim.x * kLumThreads;
    const long span   = ( ( total + stride * UNR - 1 ) / ( stride * UNR ) ) * ( stride * UNR ); // whole warps stay converged for the ballots
    for( long g0 = ( long )blockIdx.x * kLumThreads + threadIdx.x; g0 < span; g0 += stride * UNR )
    {
        uint32_t w[ UNR ][ 4 ];
        int      x0s[ UNR ];
        bool     ins[ UNR ];
#pragma unroll
        for( int q = 0; q < UNR; q++ )
        {
            const long g  = g0 + q * stride;
            const bool in = g < total;
            const int  r = in ? ( int )( g / gpr ) : 0, x0 = in ? ( int )( g - ( long )r * gpr ) * 8 : 0, y = u.yBegin + r;
            ins[ q ] = in; x0s[ q ] = x0;
            w[ q ][ 0 ] = w[ q ][ 1 ] = w[ q ][ 2 ] = w[ q ][ 3 ] = 0u;
            if( in && x0 + 8 <= u.W && vec )
            {
                const uint4 v = __ldg( reinterpret_cast< const uint4* >( at< uint16_t >( a.histogramInput, x0, y ) ) );
                w[ q ][ 0 ] = v.x; w[ q ][ 1 ] = v.y; w[ q ][ 2 ] = v.z; w[ q ][ 3 ] = v.w;
            }
            else if( in )
            {
#pragma unroll
                for( int k = 0; k < 8; k++ )
                    if( x0 + k < u.W ) w[ q ][ k >> 1 ] |= ( uint32_t )__ldg( at< uint16_t >( a.histogramInput, x0 + k, y ) ) << ( 16 * ( k & 1 ) );
            }
        }
#pragma unroll
        for( int q = 0; q < UNR; q++ )
#pragma unroll
            for( int k = 0; k < 8; k++ )
            {
                const uint32_t h = ( w[ q ][ k >> 1 ] >> ( 16 * ( k & 1 ) ) ) & 0xffffu;
                count( sTable[ h ], ins[ q ] && x0s[ q ] + k < u.W );
            }
    }
    __syncthreads();
    // CmLuminanceHistogram.comp:82-85: the work group's bins go to the global histogram
    for( int b = threadIdx.x; b < 256; b += kLumThreads )
    {
        uint32_t s = 0u;
#pragma unroll
        for( int wv = 0; wv < WARPS; wv++ ) s += sHist[ wv * 256 + b ];
        if( s ) atomicAdd( a.tm->histogram + b, s );
    }
}

// CmLuminanceAvg.comp:37-89, one work group of 256.  exp2 / exp are evaluated in double and rounded to fp32 (the pin
// "correctly rounded fp32" the oracle uses for transcendental functions whose GLSL precision is implementation-defined).
__global__ void __launch_bounds__( 256 ) k_lum_average( const __grid_constant__ LumArgs a )
{
    __shared__ uint32_t s[ 256 ];
    const uint32_t i = threadIdx.x;
    const uint32_t countInLocalBin = a.tm->histogram[ i ];
    s[ i ] = countInLocalBin * i;
    __syncthreads();
    a.tm->histogram[ i ] = 0u; // for the next frame
    for( uint32_t cutoff = 128u; cutoff > 0u; cutoff >>= 1 )
    {
        if( i < cutoff ) s[ i ] += s[ i + cutoff ];
        __syncthreads();
    }
    if( i == 0 )
    {
        if( a.stopEyeAdaptation != 0u ) return;
        const float logLuminanceRange = subr( a.maxLogLuminance, a.minLogLuminance );
        const float blackPixelCount   = ( float )countInLocalBin;
        const float weightedLogAverage = subr( divr( ( float )s[ 0 ], fmaxf( subr( a.pixelCount, blackPixelCount ), 1.0f ) ), 1.0f );
        const float e = addr( mulr( divr( weightedLogAverage, 254.0f ), logLuminanceRange ), a.minLogLuminance );
        const float weightedAvgLuminance = ( float )exp2( ( double )e );
        const float lastFrameLuminance   = a.tm->avgLuminance;
        const float tau = 1.1f;
        const float t   = subr( 1.0f, ( float )exp( ( double )mulr( -a.timeDelta, tau ) ) );
        const float adaptedLuminance = addr( lastFrameLuminance, mulr( subr( weightedAvgLuminance, lastFrameLuminance ), t ) );
        a.tm->avgLuminance = fminf( adaptedLuminance, ( float )exp2( ( double )a.maxLogLuminance ) );
    }
}

// ------------------------------------------------------------------------------------------------------------
// Volumetric scattering accumulation, CmScatterAccum.comp:96-163.  One thread per checkerboard
// pixel; the 2x2 disocclusion gather is the temporal pass's (exact operation order: which taps count is an integer
// decision), the blend is evaluated in the shader's order, so the fp16 outputs are bit-identical to the oracle's.
// ------------------------------------------------------------------------------------------------------------
struct ScatterArgs
{
    Img       motion, depth, normal, depthPrev, normalPrev, scatPrev, histPrev, scat, hist, current;
    NormalLut lut;
    float     maxHistoryLength, volumeScattering, volumeAmbient[ 3 ];
    uint32_t  volumeEnableType;
    // SIMPLE volume mode: exp(-depth * density) of every fp16 depth pattern for the current density (k_scatter_transmittance_table):
    // the depth is an fp16 texel, so the correctly rounded exp -- evaluated in double, the pin of this path -- is a function of 16 bits;
    // computing it per pixel was 60 % of the kernel's instructions
    float*    transmittance;
};

__global__ void __launch_bounds__( 256 ) k_scatter_transmittance_table( float* table, float density )
{
    const uint32_t bits = blockIdx.x * 256u + threadIdx.x; // 65536 entries
    table[ bits ] = ( float )exp( ( double )mulr( -h2f( ( uint16_t )bits ), density ) ); // pin: correctly rounded fp32
}

__global__ void __launch_bounds__( kBlockW* kBlockH ) k_scatter_accum( const __grid_constant__ ScatterArgs a, const __grid_constant__ Uniform u )
{
    const int px = blockIdx.x * kBlockW + threadIdx.x;
    const int py = u.yBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( px >= u.W || py >= u.yEnd ) return;
    const Rows rows = rowsOf( a.depth );
    const int  sep = u.W / 2, ax0 = px >= sep ? sep : 0, ax1 = ax0 + sep;

    const uint2    mv = __ldg( at< uint2 >( a.motion, px, py ) );
    const uint16_t dv = __ldg( at< uint16_t >( a.depth, px, py ) );
    const uint32_t nv = __ldg( at< uint32_t >( a.normal, px, py ) );
    const float2 mot01   = __half22float2( *reinterpret_cast< const __half2* >( &mv.x ) );
    const float  motionZ = __low2float( *reinterpret_cast< const __half2* >( &mv.y ) );
    const float  depth   = h2f( dv );
    const float3 normal  = decodeNormal( a.lut, nv );

    const float ppx = mulr( addr( mulr( addr( ( float )px, 0.5f ), u.invHalfW ), mot01.x ), u.halfW );        // getPrevScreenPos
    const float ppy = mulr( addr( mulr( addr( ( float )py, 0.5f ), u.invH ), mot01.y ), u.renderHeight );
    const float qx = subr( ppx, 0.5f ), qy = subr( ppy, 0.5f ), fx = floorf( qx ), fy = floorf( qy );
    const int   bx = f2i( fx ), by = f2i( fy );
    const float sx = subr( qx, fx ), sy = subr( qy, fy );
    const float bw[ 4 ] = { mulr( subr( 1.0f, sx ), subr( 1.0f, sy ) ), mulr( sx, subr( 1.0f, sy ) ), mulr( subr( 1.0f, sx ), sy ), mulr( sx, sy ) };

    uint32_t dz[ 4 ], en[ 4 ];
    bool     in[ 4 ];
#pragma unroll
    for( int t = 0; t < 4; t++ )
    {
        const int x = bx + ( t & 1 ), y = by + ( t >> 1 );
        in[ t ] = y >= 0 && y < u.H && x >= ax0 && x < ax1 && y >= rows.y0 && y < rows.y1;
        dz[ t ] = ldIf< uint16_t >( a.depthPrev, x, y, in[ t ], ( uint16_t )0 );
        en[ t ] = ldIf< uint32_t >( a.normalPrev, x, y, in[ t ], nv ); // taps outside never count: let them match
    }
    // a tap with the centre's packed normal decodes to the same unit vector (dot = 1 > 0.95): when all four match, the
    // sin/cos table gathers are skipped (same shortcut as the temporal pass, identical result by construction)
    uint32_t nOk = 15u;
    if( ( en[ 0 ] != nv ) | ( en[ 1 ] != nv ) | ( en[ 2 ] != nv ) | ( en[ 3 ] != nv ) )
    {
        nOk = 0u;
#pragma unroll
        for( int t = 0; t < 4; t++ ) nOk |= dot3r( normal, decodeNormal( a.lut, en[ t ] ) ) > 0.95f ? ( 1u << t ) : 0u;
    }
    float4 accum = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
    float  historyLen = 0.0f, weightSum = 0.0f;
    const float az = fabsf( depth );
#pragma unroll
    for( int t = 0; t < 4; t++ )
    {
        const float num = fabsf( addr( subr( depth, h2f( ( uint16_t )dz[ t ] ) ), motionZ ) );
        if( in[ t ] && ( ( nOk >> t ) & 1u ) && quotientBelowTenth( num, az ) )
        {
            const int x = bx + ( t & 1 ), y = by + ( t >> 1 );
            const int odd = x >= sep ? 1 : 0, rx = ( x - odd * sep ) * 2 + ( ( odd + y ) & 1 ); // getRegularPixFromCheckerboardPix
            const uint2    sp = __ldg( at< uint2 >( a.scatPrev, rx, y ) );
            const uint16_t hp = __ldg( at< uint16_t >( a.histPrev, rx, y ) );
            const float2 s01 = __half22float2( *reinterpret_cast< const __half2* >( &sp.x ) ), s23 = __half22float2( *reinterpret_cast< const __half2* >( &sp.y ) );
            const float  w = bw[ t ];
            accum = make_float4( addr( accum.x, mulr( w, s01.x ) ), addr( accum.y, mulr( w, s01.y ) ), addr( accum.z, mulr( w, s23.x ) ), addr( accum.w, mulr( w, s23.y ) ) );
            historyLen = addr( historyLen, mulr( w, h2f( hp ) ) );
            weightSum  = addr( weightSum, w );
        }
    }
    if( weightSum > 0.0001f )
    {
        historyLen = divr( historyLen, weightSum );
        accum = make_float4( divr( accum.x, weightSum ), divr( accum.y, weightSum ), divr( accum.z, weightSum ), divr( accum.w, weightSum ) );
    }
    else
    {
        historyLen = 0.0f;
        accum = make_float4( 0.0f, 0.0f, 0.0f, 0.0f );
    }
    historyLen = fminf( fmaxf( addr( historyLen, 1.0f ), 1.0f ), fmaxf( 1.0f, a.maxHistoryLength ) );

    const int odd = px >= sep ? 1 : 0, regX = ( px - odd * sep ) * 2 + ( ( odd + py ) & 1 );
    float4 cur = make_float4( 0.0f, 0.0f, 0.0f, 1.0f ); // fializeScattering, CmScatterAccum.comp:40-94
    if( a.volumeEnableType == ASVGF_VOLUME_ENABLE_VOLUMETRIC )
    {
        const uint2 c = __ldg( at< uint2 >( a.current, regX, py ) );
        const float2 c01 = __half22float2( *reinterpret_cast< const __half2* >( &c.x ) ), c23 = __half22float2( *reinterpret_cast< const __half2* >( &c.y ) );
        cur = make_float4( c01.x, c01.y, c23.x, c23.y );
    }
    else if( a.volumeEnableType == ASVGF_VOLUME_ENABLE_SIMPLE )
    {
        if( !( depth > ASVGF_MAX_RAY_LENGTH ) ) // the regular pixel maps back to this checkerboard pixel: virtualdepth = depth
        {
            const float f = __ldg( a.transmittance + dv ); // = (float)exp((double)(-depth * density)), correctly rounded
            const float g = subr( 1.0f, f );
            cur = make_float4( mulr( a.volumeAmbient[ 0 ], g ), mulr( a.volumeAmbient[ 1 ], g ), mulr( a.volumeAmbient[ 2 ], g ), f );
        }
    }
    const float t = divr( 1.0f, historyLen );
    const __half2 o0 = __floats2half2_rn( mixr( accum.x, cur.x, t ), mixr( accum.y, cur.y, t ) ), o1 = __floats2half2_rn( mixr( accum.z, cur.z, t ), mixr( accum.w, cur.w, t ) );
    *atW< uint2 >( a.scat, regX, py ) = make_uint2( *reinterpret_cast< const uint32_t* >( &o0 ), *reinterpret_cast< const uint32_t* >( &o1 ) );
    *atW< uint16_t >( a.hist, regX, py ) = f2h( historyLen );
}

// ------------------------------------------------------------------------------------------------------------
// Final prepare: CmPrepareFinal.comp:99-226 (ImageComposition::ApplyTonemapping, Source/ImageComposition.cpp:110-155).
// One thread per regular pixel: Final * transmittance + in-scattering, exposure from the eye-adapted average luminance
// (Exposure.h:36-58), AMD LPM tonemap of the ray-traced columns (LpmMap, LPM/ffx_lpm.h:747-819, in the 709 -> 709 configuration
// the shader instantiates: every optional stage off) or Uncharted 2 on the classic-shading columns, + AcidFog, the debug views,
// and the bloom input.  Arithmetic in the shader's order (explicitly rounded); pow() is ex2(y lg2 x).
// ------------------------------------------------------------------------------------------------------------
struct PrepareFinalArgs
{
    Img       final, screenEmission, scattering, acidFog, bloomInput;
    Img       motion, disPing, normal, unfiltered[ 3 ]; // the debug views (CmPrepareFinal.comp:175-226)
    NormalLut lut;
    const float* scalars; // luminous exposure, bloom luminance, Uncharted 2 white scale (k_prepare_final_scalars)
    float     saturation[ 3 ], contrast, toneScaleBias[ 2 ], lumaT[ 3 ], crosstalk[ 3 ], rcpLumaT[ 3 ]; // LPM control block, ffx_lpm.h:913-938
    float     luminanceWhitePoint, emissionMaxScreenColor, bloomThreshold, bloomEmissionMultiplier;
    float     classicEdge; // lightmapScreenCoverage * renderWidth (ShaderCommonGLSLFunc.h:492-495)
    uint32_t  debugShowFlags;
};

ASVGF_DEVICE float uncharted2Op( float x ) // CmPrepareFinal.comp:75-85
{
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    const float num = addr( mulr( x, addr( mulr( A, x ), mulr( C, B ) ) ), mulr( D, E ) );
    const float den = addr( mulr( x, addr( mulr( A, x ), B ) ), mulr( D, F ) );
    return subr( divr( num, den ), divr( E, F ) );
}

ASVGF_DEVICE float3 lpmMap709( const PrepareFinalArgs& a, float3 c )
{
    const float sat01lo = 0.0f;
    auto  sat = [ & ]( float x ) { return fminf( fmaxf( x, sat01lo ), 1.0f ); };
    float colorR = c.x, colorG = c.y, colorB = c.z;
    const float rcpMax = divr( 1.0f, fmaxf( colorR, fmaxf( colorG, colorB ) ) );
    // pow(x, y) = ex2(y * lg2 x) on the MUFU pipe (relative error ~2^-21: invisible behind the 6/6/5-bit mantissas of the output);
    // the precise powf was four fifths of this kernel's instructions
    float ratioR = fpow( mulr( colorR, rcpMax ), a.saturation[ 0 ] ), ratioG = fpow( mulr( colorG, rcpMax ), a.saturation[ 1 ] ), ratioB = fpow( mulr( colorB, rcpMax ), a.saturation[ 2 ] );
    float luma = addr( mulr( colorG, a.lumaT[ 1 ] ), addr( mulr( colorR, a.lumaT[ 0 ] ), mulr( colorB, a.lumaT[ 2 ] ) ) );
    luma       = fpow( luma, a.contrast );
    luma       = mulr( luma, divr( 1.0f, addr( mulr( luma, a.toneScaleBias[ 0 ] ), a.toneScaleBias[ 1 ] ) ) );
    const float lumaRatio  = addr( addr( mulr( ratioR, a.lumaT[ 0 ] ), mulr( ratioG, a.lumaT[ 1 ] ) ), mulr( ratioB, a.lumaT[ 2 ] ) );
    const float ratioScale = sat( mulr( luma, divr( 1.0f, lumaRatio ) ) );
    colorR = sat( mulr( ratioR, ratioScale ) ); colorG = sat( mulr( ratioG, ratioScale ) ); colorB = sat( mulr( ratioB, ratioScale ) );
    const float capR = addr( mulr( -a.crosstalk[ 0 ], colorR ), a.crosstalk[ 0 ] );
    const float capG = addr( mulr( -a.crosstalk[ 1 ], colorG ), a.crosstalk[ 1 ] );
    const float capB = addr( mulr( -a.crosstalk[ 2 ], colorB ), a.crosstalk[ 2 ] );
    auto lumaMissing = [ & ]() { return sat( addr( mulr( -colorB, a.lumaT[ 2 ] ), addr( mulr( -colorR, a.lumaT[ 0 ] ), addr( mulr( -colorG, a.lumaT[ 1 ] ), luma ) ) ) ); };
    float       lumaAdd = lumaMissing();
    const float t = mulr( lumaAdd, divr( 1.0f, addr( mulr( capG, a.lumaT[ 1 ] ), addr( mulr( capR, a.lumaT[ 0 ] ), mulr( capB, a.lumaT[ 2 ] ) ) ) ) );
    colorR = sat( addr( mulr( t, capR ), colorR ) ); colorG = sat( addr( mulr( t, capG ), colorG ) ); colorB = sat( addr( mulr( t, capB ), colorB ) );
    lumaAdd = lumaMissing();
    return make_float3( sat( addr( mulr( lumaAdd, a.rcpLumaT[ 0 ] ), colorR ) ), sat( addr( mulr( lumaAdd, a.rcpLumaT[ 1 ] ), colorG ) ), sat( addr( mulr( lumaAdd, a.rcpLumaT[ 2 ] ), colorB ) ) );
}

// the three scalars every pixel of the final prepare shares, at the pinned precision (log2 / exp2 correctly rounded: evaluated in
// double): ONE thread, once per call.  (Evaluating them in thread 0 of every CTA behind a barrier cost a third of the kernel.)
__global__ void k_prepare_final_scalars( const AsvgfTonemapping* tm, float bloomEmissionMultiplier, float luminanceWhitePoint, float* out )
{
    // Exposure.h:36-58: EV100 = log2(avgLuminance * S / K), S = 100, K = 12.5
    const float ev100 = ( float )log2( ( double )divr( mulr( fmaxf( 0.0f, tm->avgLuminance ), 100.0f ), 12.5f ) );
    const float maxLuminance = mulr( 1.2f, ( float )exp2( ( double )ev100 ) );
    out[ 0 ] = maxLuminance > 0.0f ? divr( 1.0f, maxLuminance ) : 0.0f;                            // ev100ToLuminousExposure
    out[ 1 ] = ( float )exp2( ( double )subr( addr( ev100, bloomEmissionMultiplier ), 3.0f ) );  // ev100ToLuminance(evBloom)
    out[ 2 ] = uncharted2Op( luminanceWhitePoint );
}

__global__ void __launch_bounds__( kBlockW* kBlockH ) k_prepare_final( const __grid_constant__ PrepareFinalArgs a, const __grid_constant__ Uniform u )
{
    const int x = blockIdx.x * kBlockW + threadIdx.x;
    const int y = u.yBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( x >= u.W || y >= u.yEnd ) return;
    const float exposure = __ldg( a.scalars ), bloomLum = __ldg( a.scalars + 1 ), whiteScale = __ldg( a.scalars + 2 );

    float3       hdr  = unpackUfloat111110( __ldg( at< uint32_t >( a.final, x, y ) ) );
    const float3 emis = unpackUfloat111110( __ldg( at< uint32_t >( a.screenEmission, x, y ) ) );
    {
        const uint2  v   = __ldg( at< uint2 >( a.scattering, x, y ) );
        const float2 v01 = __half22float2( *reinterpret_cast< const __half2* >( &v.x ) ), v23 = __half22float2( *reinterpret_cast< const __half2* >( &v.y ) );
        hdr = make_float3( addr( mulr( hdr.x, v23.y ), v01.x ), addr( mulr( hdr.y, v23.y ), v01.y ), addr( mulr( hdr.z, v23.y ), v23.x ) );
    }
    float3 color;
    if( ( float )x < a.classicEdge )
        color = make_float3( divr( uncharted2Op( mulr( 2.0f, hdr.x ) ), whiteScale ), divr( uncharted2Op( mulr( 2.0f, hdr.y ) ), whiteScale ), divr( uncharted2Op( mulr( 2.0f, hdr.z ) ), whiteScale ) );
    else
        color = lpmMap709( a, make_float3( addr( mulr( hdr.x, exposure ), mulr( emis.x, a.emissionMaxScreenColor ) ), addr( mulr( hdr.y, exposure ), mulr( emis.y, a.emissionMaxScreenColor ) ),
                                           addr( mulr( hdr.z, exposure ), mulr( emis.z, a.emissionMaxScreenColor ) ) ) );
    float3 bloom = make_float3( mulr( addr( hdr.x, mulr( emis.x, bloomLum ) ), exposure ), mulr( addr( hdr.y, mulr( emis.y, bloomLum ) ), exposure ), mulr( addr( hdr.z, mulr( emis.z, bloomLum ) ), exposure ) );
    if( !( fabsf( bloom.x ) < __uint_as_float( 0x7f800000u ) && fabsf( bloom.y ) < __uint_as_float( 0x7f800000u ) && fabsf( bloom.z ) < __uint_as_float( 0x7f800000u ) ) )
        bloom = make_float3( 0.0f, 0.0f, 0.0f ); // any NaN or Inf
    {
        const float k = smoothstepr( 0.0f, a.bloomThreshold, luminance( bloom ) );
        bloom = make_float3( mulr( bloom.x, k ), mulr( bloom.y, k ), mulr( bloom.z, k ) );
    }
    {
        const float3 fog = unpackUfloat111110( __ldg( at< uint32_t >( a.acidFog, x, y ) ) ); // SHIPPING_HACK, ShaderCommonGLSLFunc.h:49
        color = make_float3( addr( color.x, fog.x ), addr( color.y, fog.y ), addr( color.z, fog.z ) );
    }
    if( a.debugShowFlags != 0u )
    {
        const int sep = u.W / 2, isOdd = ( x + y % 2 ) % 2, cbx = isOdd * sep + x / 2; // getCheckerboardPix
        const uint32_t f = a.debugShowFlags;
        if( f & 1u )
        {
            const uint2  m   = __ldg( at< uint2 >( a.motion, cbx, y ) );
            const float2 m01 = __half22float2( *reinterpret_cast< const __half2* >( &m.x ) );
            color = make_float3( fabsf( m01.x ), fabsf( m01.y ), 0.0f );
        }
        else if( f & ( 1u << 2 ) ) color = decodeE5( __ldg( at< uint32_t >( a.unfiltered[ 0 ], cbx, y ) ) );
        else if( f & ( 1u << 3 ) ) color = decodeE5( __ldg( at< uint32_t >( a.unfiltered[ 1 ], cbx, y ) ) );
        else if( f & ( 1u << 4 ) ) color = decodeE5( __ldg( at< uint32_t >( a.unfiltered[ 2 ], cbx, y ) ) );
        else if( f & ( 1u << 8 ) ) { /* light grid: rejected by the host (not on this path) */ }
        else if( f & ( 1u << 1 ) )
        {
            // (the stratum of the last pixel row / column lies outside the (W+1)/3 x (H+1)/3 gradient image where W % 3 == 1 or H % 3 == 1: 0)
            const uint32_t g = ( cbx / 3 < a.disPing.w && y / 3 < a.disPing.h ) ? __ldg( at< uint32_t >( a.disPing, cbx / 3, y / 3 ) ) : 0u;
            color = make_float3( unorm8ToFloatFast( g & 0xffu ), unorm8ToFloatFast( ( g >> 8 ) & 0xffu ), unorm8ToFloatFast( ( g >> 16 ) & 0xffu ) );
        }
        else if( f & ( 1u << 10 ) )
        {
            const float3 n = decodeNormal( a.lut, __ldg( at< uint32_t >( a.normal, cbx, y ) ) );
            color = make_float3( addr( mulr( n.x, 0.5f ), 0.5f ), addr( mulr( n.y, 0.5f ), 0.5f ), addr( mulr( n.z, 0.5f ), 0.5f ) );
        }
    }
    *atW< uint32_t >( a.final, x, y )      = packUfloat111110( color );
    *atW< uint32_t >( a.bloomInput, x, y ) = packUfloat111110( bloom );
}

} // namespace asvgf
