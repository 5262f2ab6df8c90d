// asvgf_bandset.inl -- one frame split into horizontal bands across the GPUs of one box, behind the C ABI (included by asvgf.cu).
//
// The reference has no multi-GPU path; this is BASELINE config 4.  One band handle per entry of pDevices
// (a device may appear more than once: several bands on one GPU, which is how the single-GPU tests cover this code).  Every
// band handle is created with ASVGF_CREATE_BAND_APRON: each pass also computes the few rows beyond the band that the later
// passes of the frame read (asvgfBandApronRows), so results are bit-identical to the single-GPU frame and a frame needs halo
// data at only four points:
//
//     after the gradient samples ....... DISPingGradient, apron(GRADIENT_SAMPLES) gradient rows      (read by GA0 right away)
//     after the temporal pass .......... AccumHistoryLength, DiffAccumMoments, SpecAccumColor, IndirAccum (the next frame's
//                                        _Prev images) and DISGradientHistory, the whole halo           (read by frame t+1)
//     after a-trous iteration 0 ........ DiffColorHistory, the whole halo                             (read by frame t+1)
//
// EXCHANGE = PUSH, AND HIDDEN.  The band that owns the rows copies them into its neighbours' halo rows over NVLink
// (cudaMemcpy2DAsync between peer-enabled devices: copy engines, no SM time) on a side stream as soon as the producing pass
// has finished.  The state pushes of frame t therefore overlap the remaining passes of frame t (F, V, A0..A3: ~70 % of the
// frame), and frame t+1 starts without waiting; only the small gradient push (a few hundred KB) sits inside the frame.
// Ordering is by CUDA events; write-after-read hazards on the halo rows are closed the same way (a push of DiffColorHistory
// waits for the neighbours' temporal passes of the same frame, which are its only readers).
//
// HOST SIDE.  A frame is enqueued in three phases; within a phase the bands are enqueued by OpenMP threads (one per band, the
// ~40 API calls per band and frame would otherwise serialise to more than the GPU time at 8 GPUs), and the phase boundaries
// guarantee that an event has been recorded by its producer before a consumer enqueues a wait on it.
//
// MOTION BOUND.  A band holds bandHaloRows beyond its rows; the temporal pass and the gradient samples gather the previous
// frame at reprojected positions, so vertical motion beyond haloRows - apron(TEMPORAL) rows per frame would read rows the band
// does not hold.  The kernels detect exactly that (asvgf_args.cuh, haloMiss) and asvgfBandSetSynchronize returns
// ASVGF_HALO_EXCEEDED instead of silently diverging.

#include <omp.h>

struct AsvgfBandSet_T
{
    struct Band
    {
        AsvgfDenoiser_T* d = nullptr;
        int              device = 0;
        cudaStream_t     stream = nullptr, copy = nullptr;
        cudaEvent_t      evGradReady = nullptr;   // gradient samples of this frame written (stream)
        cudaEvent_t      evGradPushed = nullptr;  // ... and pushed into the neighbours' halos (copy)
        cudaEvent_t      evTemporal = nullptr;    // temporal pass of this frame done (stream)
        cudaEvent_t      evAtrous0 = nullptr;     // a-trous iteration 0 of this frame done (stream)
        cudaEvent_t      evStatePushed = nullptr; // every state push of this frame done (copy)
        cudaEvent_t      evFrame = nullptr;       // last pass of this frame done (stream)
        cudaEvent_t      evBegin = nullptr, evEnd = nullptr; // timing (asvgfBandSetBeginTiming / EndTiming)
    };
    std::vector< Band >     bands;
    std::vector< uint32_t > boundaries;
    uint32_t                W = 0, H = 0, iterations = 4, halo = 0;
    uint64_t                frames = 0;
    int                     threads = 1;
    bool                    exchange = true; // ASVGF_BAND_NO_EXCHANGE=1 (measurement only): no pushes, no waits -- wrong halos, pure kernel time
};

namespace
{

uint32_t defaultBandHalo( uint32_t iterations, uint32_t motionRows )
{
    const uint32_t need = ( uint32_t )std::max( 3 * apronRows( iterations, ASVGF_PASS_GRADIENT_SAMPLES ), apronRows( iterations, ASVGF_PASS_TEMPORAL ) + ( int )motionRows );
    return ( need + 2u ) / 3u * 3u;
}

// rows [r0, r1) (frame rows of this image's resolution) of `index` from band `src` into band `dst`, on `stream`
cudaError_t copyRows( const AsvgfBandSet_T::Band& src, const AsvgfBandSet_T::Band& dst, uint32_t index, uint32_t fi, int r0, int r1, cudaStream_t stream )
{
    const Img& a = src.d->get( index, fi );
    const Img& b = dst.d->get( index, fi );
    r0 = std::max( r0, std::max( a.row0, b.row0 ) );
    r1 = std::min( r1, std::min( a.row0 + a.rows, b.row0 + b.rows ) );
    if( r1 <= r0 || !a.p || !b.p ) return cudaSuccess;
    ImageDesc desc;
    describe( index, &desc );
    const size_t rowBytes = ( size_t )a.w * desc.bpp;
    return cudaMemcpy2DAsync( b.p + ( size_t )( r0 - b.row0 ) * b.pitch, b.pitch, a.p + ( size_t )( r0 - a.row0 ) * a.pitch, a.pitch, rowBytes, ( size_t )( r1 - r0 ),
                              cudaMemcpyDefault, stream );
}

// band k pushes the `depth` outermost rows it OWNS of the given images into the halo rows of its neighbours
cudaError_t pushEdges( AsvgfBandSet_T& set, int k, uint32_t fi, std::initializer_list< uint32_t > images, int depth, bool thirdRes )
{
    AsvgfBandSet_T::Band& me = set.bands[ k ];
    const int own0 = thirdRes ? me.d->gownBegin : me.d->ownBegin, own1 = thirdRes ? me.d->gownEnd : me.d->ownEnd;
    for( uint32_t index : images )
    {
        if( k > 0 )
        {
            cudaError_t e = copyRows( me, set.bands[ k - 1 ], index, fi, own0, std::min( own1, own0 + depth ), me.copy );
            if( e != cudaSuccess ) return e;
        }
        if( k + 1 < ( int )set.bands.size() )
        {
            cudaError_t e = copyRows( me, set.bands[ k + 1 ], index, fi, std::max( own0, own1 - depth ), own1, me.copy );
            if( e != cudaSuccess ) return e;
        }
    }
    return cudaSuccess;
}

} // namespace

extern "C" {

AsvgfResult asvgfBandSetCreate( const AsvgfBandSetCreateInfo* pInfo, AsvgfBandSet* pSet )
{
    if( !pInfo || !pSet ) return fail( ASVGF_WRONG_ARGUMENT, "null argument" );
    *pSet = nullptr;
    const uint32_t n = pInfo->deviceCount;
    if( n == 0 || n > 64 || !pInfo->pDevices ) return fail( ASVGF_WRONG_ARGUMENT, "deviceCount must be 1..64 with pDevices" );
    if( pInfo->atrousIterations < 2 || pInfo->atrousIterations > ASVGF_MAX_ATROUS_ITERATIONS ) return fail( ASVGF_WRONG_ARGUMENT, "atrousIterations must be 2..8" );
    const uint32_t H = pInfo->renderHeight;
    uint32_t halo = pInfo->bandHaloRows ? pInfo->bandHaloRows : defaultBandHalo( pInfo->atrousIterations, 48 );
    if( halo % 3 ) return fail( ASVGF_WRONG_ARGUMENT, "bandHaloRows must be a multiple of 3" );

    AsvgfBandSet_T* set = new( std::nothrow ) AsvgfBandSet_T();
    if( !set ) return fail( ASVGF_OUT_OF_MEMORY, "host allocation failed" );
    set->W = pInfo->renderWidth; set->H = H; set->iterations = pInfo->atrousIterations; set->halo = halo;
    set->boundaries.resize( n + 1 );
    for( uint32_t k = 0; k <= n; k++ )
    {
        uint32_t b = pInfo->pBandRowBoundaries ? pInfo->pBandRowBoundaries[ k ] : ( uint32_t )( ( ( uint64_t )k * H / n + 1 ) / 3 * 3 );
        if( k == n && !pInfo->pBandRowBoundaries ) b = H;
        set->boundaries[ k ] = b;
    }
    bool okRows = set->boundaries[ 0 ] == 0 && set->boundaries[ n ] == H;
    for( uint32_t k = 0; k < n && okRows; k++ )
        okRows = set->boundaries[ k + 1 ] > set->boundaries[ k ] && set->boundaries[ k ] % 3 == 0 && ( n == 1 || set->boundaries[ k + 1 ] - set->boundaries[ k ] >= halo );
    if( !okRows )
    {
        delete set;
        return fail( ASVGF_WRONG_ARGUMENT, "band boundaries must start at 0, end at renderHeight, be multiples of 3, and every band must be at least bandHaloRows tall (halos come from the direct neighbours)" );
    }

    set->bands.resize( n );
    AsvgfResult r = ASVGF_SUCCESS;
    for( uint32_t k = 0; k < n && r == ASVGF_SUCCESS; k++ )
    {
        AsvgfBandSet_T::Band& b = set->bands[ k ];
        b.device = pInfo->pDevices[ k ];
        AsvgfCreateInfo ci{};
        ci.renderWidth = pInfo->renderWidth; ci.renderHeight = H; ci.atrousIterations = pInfo->atrousIterations;
        ci.flags = ( pInfo->flags & ( ASVGF_CREATE_EXPORT_DEBUG_MASK | ASVGF_CREATE_EXTERNAL_IMAGES ) ) | ( n > 1 ? ASVGF_CREATE_BAND_APRON : 0u );
        ci.device = b.device;
        if( n > 1 ) { ci.bandRowBegin = set->boundaries[ k ]; ci.bandRowEnd = set->boundaries[ k + 1 ]; ci.bandHaloRows = halo; }
        r = asvgfCreate( &ci, &b.d );
        if( r != ASVGF_SUCCESS ) break;
        DeviceGuard guard( b.device );
        cudaError_t e = cudaStreamCreateWithFlags( &b.stream, cudaStreamNonBlocking );
        if( e == cudaSuccess ) e = cudaStreamCreateWithFlags( &b.copy, cudaStreamNonBlocking );
        for( cudaEvent_t* ev : { &b.evGradReady, &b.evGradPushed, &b.evTemporal, &b.evAtrous0, &b.evStatePushed, &b.evFrame } )
            if( e == cudaSuccess ) e = cudaEventCreateWithFlags( ev, cudaEventDisableTiming );
        if( e == cudaSuccess ) e = cudaEventCreate( &b.evBegin );
        if( e == cudaSuccess ) e = cudaEventCreate( &b.evEnd );
        if( e != cudaSuccess ) r = cudaFail( e, "cudaStreamCreate / cudaEventCreate" );
    }
    // neighbours on different devices: direct peer copies (NVLink); without peer access the copies stage through the host
    for( uint32_t k = 0; k + 1 < n && r == ASVGF_SUCCESS; k++ )
    {
        const int a = set->bands[ k ].device, b = set->bands[ k + 1 ].device;
        if( a == b ) continue;
        for( int dir = 0; dir < 2; dir++ )
        {
            const int from = dir ? b : a, to = dir ? a : b;
            int can = 0;
            if( cudaDeviceCanAccessPeer( &can, from, to ) == cudaSuccess && can )
            {
                DeviceGuard guard( from );
                cudaError_t e = cudaDeviceEnablePeerAccess( to, 0 );
                if( e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled ) r = cudaFail( e, "cudaDeviceEnablePeerAccess" );
                cudaGetLastError();
            }
        }
    }
    if( r != ASVGF_SUCCESS )
    {
        std::string msg = g_lastError;
        asvgfBandSetDestroy( set );
        return fail( r, msg );
    }
    set->threads = ( int )std::min< uint32_t >( n, ( uint32_t )std::max( 1, omp_get_max_threads() ) );
    if( const char* e = std::getenv( "ASVGF_BAND_THREADS" ) ) set->threads = std::max( 1, std::atoi( e ) );
    if( const char* e = std::getenv( "ASVGF_BAND_NO_EXCHANGE" ) ) set->exchange = std::atoi( e ) == 0;
    *pSet = set;
    return ASVGF_SUCCESS;
}

AsvgfResult asvgfBandSetDestroy( AsvgfBandSet set )
{
    if( !set ) return ASVGF_SUCCESS;
    for( auto& b : set->bands )
    {
        DeviceGuard guard( b.device );
        if( b.stream ) cudaStreamSynchronize( b.stream );
        if( b.copy ) cudaStreamSynchronize( b.copy );
        for( cudaEvent_t ev : { b.evGradReady, b.evGradPushed, b.evTemporal, b.evAtrous0, b.evStatePushed, b.evFrame, b.evBegin, b.evEnd } )
            if( ev ) cudaEventDestroy( ev );
        if( b.stream ) cudaStreamDestroy( b.stream );
        if( b.copy ) cudaStreamDestroy( b.copy );
        if( b.d ) asvgfDestroy( b.d );
    }
    delete set;
    return ASVGF_SUCCESS;
}

uint32_t asvgfBandSetGetBandCount( AsvgfBandSet set ) { return set ? ( uint32_t )set->bands.size() : 0u; }

AsvgfResult asvgfBandSetGetBand( AsvgfBandSet set, uint32_t band, AsvgfDenoiser* pDenoiser, AsvgfStream* pStream, uint32_t* pRowBegin, uint32_t* pRowEnd )
{
    if( !set || band >= set->bands.size() ) return fail( ASVGF_WRONG_ARGUMENT, "bad band index" );
    if( pDenoiser ) *pDenoiser = set->bands[ band ].d;
    if( pStream ) *pStream = ( AsvgfStream )set->bands[ band ].stream;
    if( pRowBegin ) *pRowBegin = set->boundaries[ band ];
    if( pRowEnd ) *pRowEnd = set->boundaries[ band + 1 ];
    return ASVGF_SUCCESS;
}

// = CalculateGradientsSamples + Denoiser::Denoise of the whole frame: every band on its own stream, halos pushed between them
AsvgfResult asvgfBandSetFrame( AsvgfBandSet set, uint32_t frameIndex, const AsvgfGlobalUniform* pUniform )
{
    if( !set ) return fail( ASVGF_WRONG_ARGUMENT, "null band set" );
    AsvgfResult r0 = checkCall( set->bands[ 0 ].d, frameIndex, pUniform );
    if( r0 != ASVGF_SUCCESS ) return r0;
    const int      n  = ( int )set->bands.size();
    const uint32_t fi = frameIndex, N = set->iterations;
    const bool     xch = set->exchange && n > 1;
    const bool     notFirst = set->frames > 0 && xch;
    std::vector< AsvgfResult > res( n, ASVGF_SUCCESS );
    std::vector< std::string > msg( n );
    auto note = [ & ]( int k, AsvgfResult r ) { if( r != ASVGF_SUCCESS && res[ k ] == ASVGF_SUCCESS ) { res[ k ] = r; msg[ k ] = g_lastError; } };
    auto cuda = [ & ]( int k, cudaError_t e, const char* what ) { if( e != cudaSuccess ) note( k, cudaFail( e, what ) ); };
    const int gradDepth = apronRows( N, ASVGF_PASS_GRADIENT_SAMPLES );

    // ---- phase A: gradient samples on the owned rows, pushed to the neighbours ----
#pragma omp parallel for num_threads( set->threads ) schedule( static, 1 )
    for( int k = 0; k < n; k++ )
    {
        AsvgfBandSet_T::Band& b = set->bands[ k ];
        DeviceGuard           guard( b.device );
        const Uniform         u = makeUniform( *b.d, *pUniform );
        if( notFirst ) // the previous frame's state rows of the neighbours are in this band's halo
            for( int q : { k - 1, k + 1 } )
                if( q >= 0 && q < n ) cuda( k, cudaStreamWaitEvent( b.stream, set->bands[ q ].evStatePushed, 0 ), "wait state push" );
        b.d->guideValid = b.d->centreValid = false;
        note( k, runPass( *b.d, b.stream, fi, u, ASVGF_PASS_GRADIENT_SAMPLES ) );
        if( xch )
        {
            cuda( k, cudaEventRecord( b.evGradReady, b.stream ), "record" );
            cuda( k, cudaStreamWaitEvent( b.copy, b.evGradReady, 0 ), "wait" );
            // (the neighbours finished reading their halo copy of last frame's gradients long ago: their temporal pass of the
            // previous frame precedes their state push, which this band's stream has just waited for)
            cuda( k, pushEdges( *set, k, fi, { ASVGF_IMAGE_DIS_PING_GRADIENT }, gradDepth, true ), "push gradient samples" );
            cuda( k, cudaEventRecord( b.evGradPushed, b.copy ), "record" );
        }
    }
    // ---- phase B: gradient a-trous x4 and the temporal pass (with the apron), then the push of the temporal state ----
#pragma omp parallel for num_threads( set->threads ) schedule( static, 1 )
    for( int k = 0; k < n; k++ )
    {
        AsvgfBandSet_T::Band& b = set->bands[ k ];
        DeviceGuard           guard( b.device );
        const Uniform         u = makeUniform( *b.d, *pUniform );
        for( int q : { k - 1, k + 1 } )
            if( xch && q >= 0 && q < n ) cuda( k, cudaStreamWaitEvent( b.stream, set->bands[ q ].evGradPushed, 0 ), "wait gradient push" );
        note( k, runPass( *b.d, b.stream, fi, u, ASVGF_PASS_GRADIENT_ATROUS_ALL ) );
        note( k, runPass( *b.d, b.stream, fi, u, ASVGF_PASS_TEMPORAL ) );
        cuda( k, cudaEventRecord( b.evTemporal, b.stream ), "record" );
    }
    // ---- phase C: state pushes behind the rest of the frame ----
#pragma omp parallel for num_threads( set->threads ) schedule( static, 1 )
    for( int k = 0; k < n; k++ )
    {
        AsvgfBandSet_T::Band& b = set->bands[ k ];
        DeviceGuard           guard( b.device );
        const Uniform         u = makeUniform( *b.d, *pUniform );
        if( xch )
        {
            cuda( k, cudaStreamWaitEvent( b.copy, b.evTemporal, 0 ), "wait" );
            cuda( k, pushEdges( *set, k, fi, { ASVGF_IMAGE_ACCUM_HISTORY_LENGTH, ASVGF_IMAGE_DIFF_ACCUM_MOMENTS, ASVGF_IMAGE_SPEC_ACCUM_COLOR, ASVGF_IMAGE_INDIR_ACCUM },
                                ( int )set->halo, false ), "push temporal state" );
            cuda( k, pushEdges( *set, k, fi, { ASVGF_IMAGE_DIS_GRADIENT_HISTORY }, ( int )set->halo / 3, true ), "push gradient history" );
        }
        for( uint32_t pass = ASVGF_PASS_ANTIFIREFLY; pass < ASVGF_PASS_ATROUS_0 + N; pass++ )
        {
            note( k, runPass( *b.d, b.stream, fi, u, pass ) );
            if( pass == ASVGF_PASS_ATROUS_0 && xch )
            {
                cuda( k, cudaEventRecord( b.evAtrous0, b.stream ), "record" );
                cuda( k, cudaStreamWaitEvent( b.copy, b.evAtrous0, 0 ), "wait" );
                // the neighbours read their halo copy of DiffColorHistory in their temporal pass of THIS frame: overwrite it after that
                for( int q : { k - 1, k + 1 } )
                    if( q >= 0 && q < n ) cuda( k, cudaStreamWaitEvent( b.copy, set->bands[ q ].evTemporal, 0 ), "wait" );
                cuda( k, pushEdges( *set, k, fi, { ASVGF_IMAGE_DIFF_COLOR_HISTORY }, ( int )set->halo, false ), "push colour history" );
                cuda( k, cudaEventRecord( b.evStatePushed, b.copy ), "record" );
            }
        }
        cuda( k, cudaEventRecord( b.evFrame, b.stream ), "record" );
    }
    set->frames++;
    for( int k = 0; k < n; k++ )
        if( res[ k ] != ASVGF_SUCCESS ) return fail( res[ k ], "band " + std::to_string( k ) + ": " + msg[ k ] );
    return ASVGF_SUCCESS;
}

AsvgfResult asvgfBandSetSynchronize( AsvgfBandSet set )
{
    if( !set ) return fail( ASVGF_WRONG_ARGUMENT, "null band set" );
    AsvgfResult out = ASVGF_SUCCESS;
    std::string msg;
    for( auto& b : set->bands )
    {
        DeviceGuard guard( b.device );
        cudaError_t e = cudaStreamSynchronize( b.copy );
        if( e != cudaSuccess ) return cudaFail( e, "cudaStreamSynchronize(copy)" );
        AsvgfResult r = asvgfSynchronize( b.d, ( AsvgfStream )b.stream );
        if( r != ASVGF_SUCCESS && out == ASVGF_SUCCESS ) { out = r; msg = g_lastError; }
    }
    return out == ASVGF_SUCCESS ? ASVGF_SUCCESS : fail( out, msg );
}

// Device-side timing of a run of frames: Begin records an event on every band's stream (after the streams have been brought to
// a common start by the caller's synchronize), End records another and returns the longest band interval in milliseconds.
AsvgfResult asvgfBandSetBeginTiming( AsvgfBandSet set )
{
    if( !set ) return fail( ASVGF_WRONG_ARGUMENT, "null band set" );
    for( auto& b : set->bands )
    {
        DeviceGuard guard( b.device );
        CUDA_TRY( cudaEventRecord( b.evBegin, b.stream ) );
    }
    return ASVGF_SUCCESS;
}
AsvgfResult asvgfBandSetEndTiming( AsvgfBandSet set, float* pMaxMilliseconds, float* pPerBandMilliseconds )
{
    if( !set || !pMaxMilliseconds ) return fail( ASVGF_WRONG_ARGUMENT, "null argument" );
    for( auto& b : set->bands )
    {
        DeviceGuard guard( b.device );
        CUDA_TRY( cudaStreamWaitEvent( b.stream, b.evStatePushed, 0 ) ); // (recorded only when there is more than one band)
        CUDA_TRY( cudaEventRecord( b.evEnd, b.stream ) );
    }
    float worst = 0.0f;
    for( size_t k = 0; k < set->bands.size(); k++ )
    {
        auto&       b = set->bands[ k ];
        DeviceGuard guard( b.device );
        CUDA_TRY( cudaEventSynchronize( b.evEnd ) );
        float ms = 0.0f;
        CUDA_TRY( cudaEventElapsedTime( &ms, b.evBegin, b.evEnd ) );
        if( pPerBandMilliseconds ) pPerBandMilliseconds[ k ] = ms;
        worst = std::max( worst, ms );
    }
    *pMaxMilliseconds = worst;
    return ASVGF_SUCCESS;
}

} // extern "C"
