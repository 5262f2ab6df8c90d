// asvgf.cu -- host side of libasvgf.so: the Framebuffers mirror (image table with the cur/prev swap), the
// Denoiser mirror (stream launcher of the pass kernels) and the C ABI of include/asvgf.h.
//
// Reference: Source/Denoiser.cpp:64-255 (dispatch order, ping-pong routing), Source/Framebuffers.cpp:35-50
// (FrameIndexToFBIndex), :639-700 (image extents), Source/PathTracer.cpp:196-219 (gradient samples dispatch).
#include "../../include/asvgf.h"

#include "asvgf_fast.cuh"
#include "asvgf_temporal.cuh"
#include "asvgf_atrous_tiled.cuh"
#include "asvgf_atrous_half.cuh"
#include "asvgf_post.cuh"
#ifdef ASVGF_WITH_LITERAL // expression-by-expression kernels for A/B numerics; not part of the release build
    #include "../../experiments/asvgf_literal_kernels.cuh"
#endif

#include <nvtx3/nvToolsExt.h> // header-only in CUDA 12: ranges are no-ops unless a profiler is attached

// tests/hostsim compiles this file with g++ -DASVGF_HOSTSIM against a stub of the CUDA runtime, so that the CPU test suite can
// drive the library's host logic and the kernels' source without a GPU (test infrastructure: tests/hostsim/README.md; the
// product is built by nvcc, where the macro is never defined and a launch is a launch)
#ifndef ASVGF_HOSTSIM
    #define ASVGF_LAUNCH( kernel, grid, block, smem, stream, ... ) kernel<<< grid, block, smem, stream >>>( __VA_ARGS__ )
#else
    #define ASVGF_LAUNCH( kernel, grid, block, smem, stream, ... ) ( ( void )( stream ), hostsim::launch( kernel, grid, block, smem, __VA_ARGS__ ) )
#endif

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace asvgf;

namespace
{

thread_local std::string g_lastError;

AsvgfResult fail( AsvgfResult r, const std::string& msg )
{
    g_lastError = msg;
    return r;
}
AsvgfResult cudaFail( cudaError_t e, const char* what )
{
    return fail( e == cudaErrorMemoryAllocation ? ASVGF_OUT_OF_MEMORY : ASVGF_CUDA_ERROR, std::string( what ) + ": " + cudaGetErrorString( e ) );
}
#define CUDA_TRY( expr )                                         \
    do                                                           \
    {                                                            \
        cudaError_t e_ = ( expr );                               \
        if( e_ != cudaSuccess ) return cudaFail( e_, #expr );    \
    } while( 0 )

struct ImageDesc { uint32_t format, bpp; bool thirdRes; };

bool describe( uint32_t i, ImageDesc* d )
{
    switch( i )
    {
        case ASVGF_IMAGE_IS_SKY: *d = { ASVGF_FORMAT_R8_UINT, 1, false }; return true;
        case ASVGF_IMAGE_DEBUG_MASK: *d = { ASVGF_FORMAT_R8_UINT, 1, false }; return true;
        case ASVGF_IMAGE_GRADIENT_PREV_PIX: *d = { ASVGF_FORMAT_R8_UINT, 1, true }; return true;
        case ASVGF_IMAGE_METALLIC_ROUGHNESS: case ASVGF_IMAGE_METALLIC_ROUGHNESS_PREV: *d = { ASVGF_FORMAT_R8G8_UNORM, 2, false }; return true;
        case ASVGF_IMAGE_DEPTH_WORLD: case ASVGF_IMAGE_DEPTH_WORLD_PREV: case ASVGF_IMAGE_DEPTH_GRAD:
        case ASVGF_IMAGE_ATROUS_FILTERED_VARIANCE: case ASVGF_IMAGE_HISTOGRAM_INPUT: case ASVGF_IMAGE_SCATTERING_HISTORY:
        case ASVGF_IMAGE_SCATTERING_HISTORY_PREV: *d = { ASVGF_FORMAT_R16_SFLOAT, 2, false }; return true;
        case ASVGF_IMAGE_DIFF_ACCUM_MOMENTS: case ASVGF_IMAGE_DIFF_ACCUM_MOMENTS_PREV: case ASVGF_IMAGE_GRADIENT_INPUTS:
        case ASVGF_IMAGE_GRADIENT_INPUTS_PREV: case ASVGF_IMAGE_FWD_LUM: *d = { ASVGF_FORMAT_R16G16_SFLOAT, 4, false }; return true;
        case ASVGF_IMAGE_MOTION: case ASVGF_IMAGE_VIEW_DIRECTION: case ASVGF_IMAGE_THROUGHPUT: case ASVGF_IMAGE_ACCUM_HISTORY_LENGTH:
        case ASVGF_IMAGE_ACCUM_HISTORY_LENGTH_PREV: case ASVGF_IMAGE_DIFF_COLOR_HISTORY: case ASVGF_IMAGE_DIFF_PING_COLOR_AND_VARIANCE:
        case ASVGF_IMAGE_DIFF_PONG_COLOR_AND_VARIANCE: case ASVGF_IMAGE_SCATTERING: case ASVGF_IMAGE_SCATTERING_PREV: case ASVGF_IMAGE_SCATTER_CURRENT:
            *d = { ASVGF_FORMAT_R16G16B16A16_SFLOAT, 8, false }; return true;
        case ASVGF_IMAGE_SURFACE_POSITION: case ASVGF_IMAGE_SURFACE_POSITION_PREV: *d = { ASVGF_FORMAT_R32G32B32A32_SFLOAT, 16, false }; return true;
        case ASVGF_IMAGE_ALBEDO: case ASVGF_IMAGE_PRE_FINAL: case ASVGF_IMAGE_FINAL: case ASVGF_IMAGE_ACID_FOG_RT: case ASVGF_IMAGE_ACID_FOG:
        case ASVGF_IMAGE_SCREEN_EMIS_RT: case ASVGF_IMAGE_SCREEN_EMISSION: case ASVGF_IMAGE_BLOOM_INPUT: *d = { ASVGF_FORMAT_B10G11R11_UFLOAT, 4, false }; return true;
        case ASVGF_IMAGE_NORMAL: case ASVGF_IMAGE_NORMAL_PREV: case ASVGF_IMAGE_UNFILTERED_DIRECT: case ASVGF_IMAGE_UNFILTERED_SPECULAR:
        case ASVGF_IMAGE_UNFILTERED_INDIR: case ASVGF_IMAGE_DIFF_TEMPORARY: case ASVGF_IMAGE_DIFF_ACCUM_COLOR: case ASVGF_IMAGE_DIFF_ACCUM_COLOR_PREV:
        case ASVGF_IMAGE_SPEC_ACCUM_COLOR: case ASVGF_IMAGE_SPEC_ACCUM_COLOR_PREV: case ASVGF_IMAGE_SPEC_PING_COLOR: case ASVGF_IMAGE_SPEC_PONG_COLOR:
        case ASVGF_IMAGE_INDIR_ACCUM: case ASVGF_IMAGE_INDIR_ACCUM_PREV: case ASVGF_IMAGE_INDIR_PING: case ASVGF_IMAGE_INDIR_PONG:
            *d = { ASVGF_FORMAT_R32_UINT, 4, false }; return true;
        case ASVGF_IMAGE_DIS_PING_GRADIENT: case ASVGF_IMAGE_DIS_PONG_GRADIENT: case ASVGF_IMAGE_DIS_GRADIENT_HISTORY:
            *d = { ASVGF_FORMAT_R8G8B8A8_UNORM, 4, true }; return true;
        default: return false;
    }
}

// Framebuffers.cpp:35-50: 1 = "cur" of a swappable pair (physical = index + frameIndex), 2 = its _PREV twin
int swapKind( uint32_t i )
{
    switch( i )
    {
        case ASVGF_IMAGE_NORMAL: case ASVGF_IMAGE_METALLIC_ROUGHNESS: case ASVGF_IMAGE_DEPTH_WORLD: case ASVGF_IMAGE_SURFACE_POSITION:
        case ASVGF_IMAGE_ACCUM_HISTORY_LENGTH: case ASVGF_IMAGE_DIFF_ACCUM_COLOR: case ASVGF_IMAGE_DIFF_ACCUM_MOMENTS:
        case ASVGF_IMAGE_SPEC_ACCUM_COLOR: case ASVGF_IMAGE_INDIR_ACCUM: case ASVGF_IMAGE_GRADIENT_INPUTS: case ASVGF_IMAGE_SCATTERING:
        case ASVGF_IMAGE_SCATTERING_HISTORY: return 1;
        case ASVGF_IMAGE_NORMAL_PREV: case ASVGF_IMAGE_METALLIC_ROUGHNESS_PREV: case ASVGF_IMAGE_DEPTH_WORLD_PREV: case ASVGF_IMAGE_SURFACE_POSITION_PREV:
        case ASVGF_IMAGE_ACCUM_HISTORY_LENGTH_PREV: case ASVGF_IMAGE_DIFF_ACCUM_COLOR_PREV: case ASVGF_IMAGE_DIFF_ACCUM_MOMENTS_PREV:
        case ASVGF_IMAGE_SPEC_ACCUM_COLOR_PREV: case ASVGF_IMAGE_INDIR_ACCUM_PREV: case ASVGF_IMAGE_GRADIENT_INPUTS_PREV:
        case ASVGF_IMAGE_SCATTERING_PREV: case ASVGF_IMAGE_SCATTERING_HISTORY_PREV: return 2;
        default: return 0;
    }
}

// ---- normal-codec trig table, one per device, shared by all handles ----
struct LutEntry { float2* phi = nullptr; float2* theta = nullptr; int refs = 0; };
std::mutex g_lutMutex;
LutEntry   g_lut[ 64 ];

AsvgfResult acquireLut( int device, NormalLut* out )
{
    std::lock_guard< std::mutex > lock( g_lutMutex );
    if( device < 0 || device >= 64 ) return fail( ASVGF_WRONG_ARGUMENT, "device ordinal out of range" );
    LutEntry& e = g_lut[ device ];
    if( e.refs == 0 )
    {
        // Utils.h:124-140: phi = j*pi/65535, theta = k*2*pi/65536 evaluated in fp32 as GLSL does, then sin/cos
        // correctly rounded (double libm -> float), the same pin the oracle uses.
        const float             PI_F = 3.14159265358979323846f;
        std::vector< float2 > phi( 65536 ), theta( 65536 );
        for( uint32_t j = 0; j < 65536; j++ )
        {
            const float p = ( float )j * PI_F / 65535.0f;
            const float t = ( float )( j * 2u ) * PI_F / 65536.0f;
            phi[ j ]      = make_float2( ( float )sin( ( double )p ), ( float )cos( ( double )p ) );
            theta[ j ]    = make_float2( ( float )sin( ( double )t ), ( float )cos( ( double )t ) );
        }
        CUDA_TRY( cudaMalloc( &e.phi, 65536 * sizeof( float2 ) ) );
        CUDA_TRY( cudaMalloc( &e.theta, 65536 * sizeof( float2 ) ) );
        CUDA_TRY( cudaMemcpy( e.phi, phi.data(), 65536 * sizeof( float2 ), cudaMemcpyHostToDevice ) );
        CUDA_TRY( cudaMemcpy( e.theta, theta.data(), 65536 * sizeof( float2 ), cudaMemcpyHostToDevice ) );
    }
    e.refs++;
    out->phi   = e.phi;
    out->theta = e.theta;
    return ASVGF_SUCCESS;
}
void releaseLut( int device )
{
    std::lock_guard< std::mutex > lock( g_lutMutex );
    LutEntry& e = g_lut[ device ];
    if( e.refs > 0 && --e.refs == 0 )
    {
        cudaFree( e.phi ); cudaFree( e.theta );
        e.phi = e.theta = nullptr;
    }
}

void setImgStorage( Img& im, uint8_t* p, uint32_t pitch, uint32_t bpp )
{
    im.p = p; im.pitch = pitch;
    im.pitchT = pitch / bpp;
    im.p0 = p ? p - ( ptrdiff_t )im.row0 * ( ptrdiff_t )pitch : nullptr;
}

// profiler ranges named like the reference's command-buffer labels (CmdLabel: Source/Denoiser.cpp:90,130,162,184,205,
// Source/PathTracer.cpp:198, Source/Tonemapping.cpp:63, Source/ImageComposition.cpp:162, Source/Volumetric.cpp:78)
struct NvtxRange
{
    explicit NvtxRange( const char* name ) { nvtxRangePushA( name ); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct DeviceGuard
{
    int  prev = -1;
    bool ok   = false;
    explicit DeviceGuard( int device )
    {
        if( cudaGetDevice( &prev ) != cudaSuccess ) return;
        ok = prev == device || cudaSetDevice( device ) == cudaSuccess;
    }
    ~DeviceGuard() { if( ok && prev >= 0 ) { int cur = -1; cudaGetDevice( &cur ); if( cur != prev ) cudaSetDevice( prev ); } }
};

} // namespace

struct AsvgfDenoiser_T
{
    AsvgfCreateInfo info{};
    int             device = 0;
    uint32_t        W = 0, H = 0, iterations = 4;
    // stored rows of full-resolution and 1/3-resolution images; owned rows are what passes compute
    int32_t   row0 = 0, rows = 0, ownBegin = 0, ownEnd = 0;
    int32_t   grow0 = 0, grows = 0, gownBegin = 0, gownEnd = 0;
    Img       img[ ASVGF_IMAGE_COUNT ]{};
    bool      owned[ ASVGF_IMAGE_COUNT ]{};
    NormalLut lut{};
    bool      haveLut     = false;
    // internal per-frame scratch: the decoded geometry of the current frame (asvgf_atrous_tiled.cuh, k_guide)
    Img       guide{}, centre{};
    uint32_t* haloMiss = nullptr;  // band handles: device flag set by a reprojected tap outside the stored rows (asvgf_args.cuh)
    bool      centreValid = false; // `centre` holds the weight constants of the frame being denoised (written by a-trous iteration 0)
    // cross-stream ordering of the gradient chain (asvgfGradientSamples may be given another stream than asvgfDenoise):
    // the gradient-sample pass of frame t+1 waits for the temporal pass of frame t, the temporal pass waits for the gradient chain
    cudaEvent_t evTemporalDone = nullptr, evGradientsDone = nullptr;
    bool        temporalRecorded = false, gradientsRecorded = false;
    bool        gradientAtrousDone = false; // ASVGF_CREATE_EAGER_GRADIENT_ATROUS: the chain already ran behind the gradient samples
    bool        guideValid = false;         // the guide plane matches the geometry images of the frame being denoised
    uint64_t  launchCount = 0;
    bool      atrousOptIn[ ASVGF_MAX_ATROUS_ITERATIONS ][ 4 ] = {}; // cudaFuncSetAttribute done for (iteration, composing + 2 * centre)
    bool      atrousCentre = true; // iterations >= 1 read the weight constants iteration 0 published (ASVGF_ATROUS_CENTRE=0: re-derive them)
    // Programmatic dependent launch (ASVGF_PDL=0 turns it off).  Inside asvgfGradientSamples / asvgfDenoise a kernel whose
    // predecessor on the stream is this call's previous kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization:
    // its CTAs are scheduled while the predecessor's last wave drains and block in griddepcontrol.wait (asvgf_device.cuh).  The
    // first kernel of a call, and a kernel behind an event record / wait, are launched normally.
    bool      pdl = true, pdlChain = false, pdlArmed = false;
    int       smCount     = 0;
    int32_t   passGyBegin = 0, passGyEnd = 0; // gradient rows the current gradient pass computes (runPass sets them)
    // exposure state (= the reference's tonemapping buffer) and the tables of the histogram pass
    AsvgfTonemapping* tm       = nullptr;
    float*            log2Half = nullptr; // correctly rounded log2 of every fp16 bit pattern
    uint8_t*          binTable = nullptr; // histogram bin of every fp16 bit pattern for the current uniform
    float*            transmittance = nullptr; // exp(-depth * density) of every fp16 depth for `transmittanceDensity` (asvgfScatterAccumulate)
    float             transmittanceDensity = -1.0f;
    float*            finalScalars = nullptr;  // asvgfPrepareFinal: exposure, bloom luminance, white scale of the call

    const Img& get( uint32_t index, uint32_t frameIndex ) const
    {
        const int k = swapKind( index );
        return img[ k == 1 ? index + frameIndex : k == 2 ? index - frameIndex : index ];
    }
};

namespace
{

Uniform makeUniform( const AsvgfDenoiser_T& d, const AsvgfGlobalUniform& g )
{
    Uniform u;
    std::memcpy( u.view, g.view, 64 ); std::memcpy( u.viewPrev, g.viewPrev, 64 );
    std::memcpy( u.projection, g.projection, 64 ); std::memcpy( u.projectionPrev, g.projectionPrev, 64 );
    for( int i = 0; i < 3; i++ ) { u.cameraPosition[ i ] = g.cameraPosition[ i ]; u.cameraPositionPrev[ i ] = g.cameraPositionPrev[ i ]; }
    u.renderWidth = g.renderWidth; u.renderHeight = g.renderHeight; u.timeDelta = g.timeDelta;
    u.lightmapScreenCoverage = g.lightmapScreenCoverage;
    u.gradientMult[ 0 ] = g.gradientMultDiffuse; u.gradientMult[ 1 ] = g.gradientMultIndirect; u.gradientMult[ 2 ] = g.gradientMultSpecular;
    u.debugShowFlags = g.debugShowFlags; u.antiFireflyEnabled = g.antiFireflyEnabled;
    u.W = ( int32_t )g.renderWidth; u.H = ( int32_t )g.renderHeight;
    u.yBegin = d.ownBegin; u.yEnd = d.ownEnd;
    u.halfW = g.renderWidth / 2.0f; u.invHalfW = 1.0f / u.halfW; u.invH = 1.0f / g.renderHeight;
    {
        const float rcp = g.timeDelta <= 0.0f ? 0.0f : 1.0f / g.timeDelta;
        u.specMinFrames = std::fmin( std::fmax( 0.2f * rcp, 0.0f ), 128.0f );
    }
    for( int i = 0; i < 3; i++ ) u.cameraDelta[ i ] = g.cameraPositionPrev[ i ] - g.cameraPosition[ i ];
    u.projSparse = 1;
    for( int i : { 1, 2, 3, 4, 6, 7, 12, 13, 15 } )
        if( g.projection[ i ] != 0.0f || g.projectionPrev[ i ] != 0.0f ) u.projSparse = 0;
    return u;
}

dim3 gridFor( int w, int rows ) { return dim3( ( unsigned )( ( w + kBlockW - 1 ) / kBlockW ), ( unsigned )( ( rows + kBlockH - 1 ) / kBlockH ), 1 ); }
const dim3 kBlock( kBlockW, kBlockH, 1 );

AsvgfResult need( const AsvgfDenoiser_T& d, uint32_t frameIndex, std::initializer_list< uint32_t > indices )
{
    for( uint32_t i : indices )
        if( d.get( i, frameIndex ).p == nullptr ) return fail( ASVGF_NOT_BOUND, "image " + std::to_string( i ) + " is neither owned nor bound" );
    return ASVGF_SUCCESS;
}
#define NEED( ... )                                                        \
    do                                                                     \
    {                                                                      \
        AsvgfResult r_ = need( d, fi, { __VA_ARGS__ } );                   \
        if( r_ != ASVGF_SUCCESS ) return r_;                               \
    } while( 0 )

AsvgfResult launchCheck( AsvgfDenoiser_T& d, const char* name )
{
    d.launchCount++;
    cudaError_t e = cudaGetLastError();
    if( e != cudaSuccess ) return cudaFail( e, name );
    return ASVGF_SUCCESS;
}

// every production kernel goes through here: <<<>>> semantics, plus the programmatic-serialization attribute when armed
template< typename... KArgs, typename... Args >
void launchK( AsvgfDenoiser_T& d, void ( *kernel )( KArgs... ), dim3 grid, dim3 block, size_t smem, cudaStream_t s, const Args&... args )
{
#ifndef ASVGF_HOSTSIM
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[ 1 ];
    attr[ 0 ].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[ 0 ].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = ( d.pdlChain && d.pdlArmed ) ? 1u : 0u;
    ( void )cudaLaunchKernelEx( &cfg, kernel, args... ); // the error is picked up by launchCheck
#else
    ( void )s;
    hostsim::launch( kernel, grid, block, smem, args... );
#endif
    d.pdlArmed = d.pdlChain;
}

#ifdef ASVGF_WITH_LITERAL
bool literal( const AsvgfDenoiser_T& d ) { return ( d.info.flags & ASVGF_CREATE_LITERAL_KERNELS ) != 0; }
    #define ASVGF_LITERAL_OR( d, literalLaunch, productionLaunch ) do { if( literal( d ) ) { literalLaunch; } else { productionLaunch; } } while( 0 )
#else
    #define ASVGF_LITERAL_OR( d, literalLaunch, productionLaunch ) do { productionLaunch; } while( 0 )
#endif

AsvgfResult runGradientSamples( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const Uniform& u )
{
    NEED( ASVGF_IMAGE_GRADIENT_PREV_PIX, ASVGF_IMAGE_DIS_PING_GRADIENT, ASVGF_IMAGE_DIS_GRADIENT_HISTORY, ASVGF_IMAGE_MOTION,
          ASVGF_IMAGE_GRADIENT_INPUTS_PREV, ASVGF_IMAGE_IS_SKY, ASVGF_IMAGE_INDIR_ACCUM_PREV, ASVGF_IMAGE_FWD_LUM );
    GradientSamplesArgs a;
    a.gradientPrevPix = d.get( ASVGF_IMAGE_GRADIENT_PREV_PIX, fi ); a.disPing = d.get( ASVGF_IMAGE_DIS_PING_GRADIENT, fi );
    a.disHistory = d.get( ASVGF_IMAGE_DIS_GRADIENT_HISTORY, fi ); a.motion = d.get( ASVGF_IMAGE_MOTION, fi );
    a.gradientInputsPrev = d.get( ASVGF_IMAGE_GRADIENT_INPUTS_PREV, fi ); a.isSky = d.get( ASVGF_IMAGE_IS_SKY, fi );
    a.indirAccumPrev = d.get( ASVGF_IMAGE_INDIR_ACCUM_PREV, fi ); a.fwdLum = d.get( ASVGF_IMAGE_FWD_LUM, fi );
    a.gyBegin = d.passGyBegin; a.gyEnd = d.passGyEnd;
    a.haloMiss = d.haloMiss;
    ASVGF_LITERAL_OR( d, ( k_gradient_samples<<< gridFor( a.disPing.w, a.gyEnd - a.gyBegin ), kBlock, 0, s >>>( a, u ) ),
                      ( launchK( d, k_gradient_samples_fast, gridFor( a.disPing.w, a.gyEnd - a.gyBegin ), kBlock, 0, s, a, u ) ) );
    return launchCheck( d, "k_gradient_samples" );
}

AsvgfResult runGradientAtrous( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const Uniform& u, uint32_t iteration )
{
    NEED( ASVGF_IMAGE_DIS_PING_GRADIENT, ASVGF_IMAGE_DIS_PONG_GRADIENT, ASVGF_IMAGE_DIS_GRADIENT_HISTORY );
    const bool         even = ( iteration % 2 ) == 0; // CmASVGFGradientAtrous.comp:89-104
    GradientAtrousArgs a;
    a.src = d.get( even ? ASVGF_IMAGE_DIS_PING_GRADIENT : ASVGF_IMAGE_DIS_PONG_GRADIENT, fi );
    a.dst = d.get( even ? ASVGF_IMAGE_DIS_PONG_GRADIENT : ASVGF_IMAGE_DIS_PING_GRADIENT, fi );
    a.history = d.get( ASVGF_IMAGE_DIS_GRADIENT_HISTORY, fi );
    a.iteration = ( int32_t )iteration; a.gyBegin = d.passGyBegin; a.gyEnd = d.passGyEnd;
    ASVGF_LITERAL_OR( d, ( k_gradient_atrous<<< gridFor( a.dst.w, a.gyEnd - a.gyBegin ), kBlock, 0, s >>>( a, u ) ),
                      ( launchK( d, k_gradient_atrous_fast, gridFor( a.dst.w, a.gyEnd - a.gyBegin ), kBlock, 0, s, a, u ) ) );
    return launchCheck( d, "k_gradient_atrous" );
}

AsvgfResult runTemporal( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const Uniform& u )
{
    NEED( ASVGF_IMAGE_UNFILTERED_DIRECT, ASVGF_IMAGE_UNFILTERED_INDIR, ASVGF_IMAGE_UNFILTERED_SPECULAR, ASVGF_IMAGE_VIEW_DIRECTION, ASVGF_IMAGE_MOTION,
          ASVGF_IMAGE_DEPTH_WORLD, ASVGF_IMAGE_NORMAL, ASVGF_IMAGE_METALLIC_ROUGHNESS, ASVGF_IMAGE_SURFACE_POSITION, ASVGF_IMAGE_DIS_PING_GRADIENT,
          ASVGF_IMAGE_DEPTH_WORLD_PREV, ASVGF_IMAGE_NORMAL_PREV, ASVGF_IMAGE_DIFF_COLOR_HISTORY, ASVGF_IMAGE_DIFF_ACCUM_MOMENTS_PREV,
          ASVGF_IMAGE_ACCUM_HISTORY_LENGTH_PREV, ASVGF_IMAGE_SPEC_ACCUM_COLOR_PREV, ASVGF_IMAGE_INDIR_ACCUM_PREV, ASVGF_IMAGE_SURFACE_POSITION_PREV,
          ASVGF_IMAGE_DIFF_TEMPORARY, ASVGF_IMAGE_DIFF_ACCUM_MOMENTS, ASVGF_IMAGE_ACCUM_HISTORY_LENGTH, ASVGF_IMAGE_SPEC_ACCUM_COLOR, ASVGF_IMAGE_INDIR_ACCUM );
    TemporalArgs a;
    a.unfDirect = d.get( ASVGF_IMAGE_UNFILTERED_DIRECT, fi ); a.unfIndir = d.get( ASVGF_IMAGE_UNFILTERED_INDIR, fi );
    a.unfSpec = d.get( ASVGF_IMAGE_UNFILTERED_SPECULAR, fi ); a.viewDir = d.get( ASVGF_IMAGE_VIEW_DIRECTION, fi );
    a.motion = d.get( ASVGF_IMAGE_MOTION, fi ); a.depth = d.get( ASVGF_IMAGE_DEPTH_WORLD, fi ); a.normal = d.get( ASVGF_IMAGE_NORMAL, fi );
    a.metRough = d.get( ASVGF_IMAGE_METALLIC_ROUGHNESS, fi ); a.surfPos = d.get( ASVGF_IMAGE_SURFACE_POSITION, fi );
    a.disPing = d.get( ASVGF_IMAGE_DIS_PING_GRADIENT, fi );
    a.depthPrev = d.get( ASVGF_IMAGE_DEPTH_WORLD_PREV, fi ); a.normalPrev = d.get( ASVGF_IMAGE_NORMAL_PREV, fi );
    a.diffHistory = d.get( ASVGF_IMAGE_DIFF_COLOR_HISTORY, fi ); a.momentsPrev = d.get( ASVGF_IMAGE_DIFF_ACCUM_MOMENTS_PREV, fi );
    a.histLenPrev = d.get( ASVGF_IMAGE_ACCUM_HISTORY_LENGTH_PREV, fi ); a.specPrev = d.get( ASVGF_IMAGE_SPEC_ACCUM_COLOR_PREV, fi );
    a.indirPrev = d.get( ASVGF_IMAGE_INDIR_ACCUM_PREV, fi ); a.surfPosPrev = d.get( ASVGF_IMAGE_SURFACE_POSITION_PREV, fi );
    a.outDiffTemporary = d.get( ASVGF_IMAGE_DIFF_TEMPORARY, fi ); a.outMoments = d.get( ASVGF_IMAGE_DIFF_ACCUM_MOMENTS, fi );
    a.outHistLen = d.get( ASVGF_IMAGE_ACCUM_HISTORY_LENGTH, fi ); a.outSpecAccum = d.get( ASVGF_IMAGE_SPEC_ACCUM_COLOR, fi );
    a.outIndirAccum = d.get( ASVGF_IMAGE_INDIR_ACCUM, fi ); a.outMask = d.get( ASVGF_IMAGE_DEBUG_MASK, fi );
    a.lut = d.lut;
    a.writeMask = ( ( d.info.flags & ASVGF_CREATE_EXPORT_DEBUG_MASK ) && a.outMask.p ) ? 1 : 0;
    // the a-trous iterations read the guide plane up to their reach beyond the rows they compute: this launch covers those rows
    // when it runs on the whole frame or with the band apron (18 rows >= 14 + 1); otherwise k_guide fills every stored row
    a.guide = d.guide;
    a.writeGuide = ( d.info.bandRowEnd == 0 || ( d.info.flags & ASVGF_CREATE_BAND_APRON ) ) ? 1 : 0;
    d.guideValid = a.writeGuide != 0;
    // one texel offset serves every image when they all hold the same number of texels per row (the library's own allocation)
    bool uniformPitch = true;
    for( const Img* im : { &a.unfDirect, &a.unfIndir, &a.unfSpec, &a.viewDir, &a.motion, &a.depth, &a.normal, &a.metRough, &a.surfPos, &a.depthPrev, &a.normalPrev,
                           &a.diffHistory, &a.momentsPrev, &a.histLenPrev, &a.specPrev, &a.indirPrev, &a.surfPosPrev, &a.outDiffTemporary, &a.outMoments, &a.outHistLen,
                           &a.outSpecAccum, &a.outIndirAccum } )
        uniformPitch = uniformPitch && im->pitchT == a.depth.pitchT;
    const dim3 grid = gridFor( ( int )d.W, u.yEnd - u.yBegin );
    a.haloMiss = d.haloMiss;
    auto production = [ & ]() {
        const dim3 tGrid( grid.x, ( unsigned )( ( u.yEnd - u.yBegin + kTBlockH - 1 ) / kTBlockH ), 1 ), tBlock( kBlockW, kTBlockH, 1 );
        if( d.haloMiss )
        {
            if( uniformPitch ) launchK( d, k_temporal_fast< true, true >, tGrid, tBlock, 0, s, a, u );
            else launchK( d, k_temporal_fast< false, true >, tGrid, tBlock, 0, s, a, u );
        }
        else
        {
            if( uniformPitch ) launchK( d, k_temporal_fast< true, false >, tGrid, tBlock, 0, s, a, u );
            else launchK( d, k_temporal_fast< false, false >, tGrid, tBlock, 0, s, a, u );
        }
    };
    ASVGF_LITERAL_OR( d, ( k_temporal<<< grid, kBlock, 0, s >>>( a, u ) ), production() );
    return launchCheck( d, "k_temporal" );
}

AsvgfResult runAntiFirefly( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const Uniform& u )
{
    NEED( ASVGF_IMAGE_DIFF_TEMPORARY, ASVGF_IMAGE_SPEC_ACCUM_COLOR, ASVGF_IMAGE_INDIR_ACCUM, ASVGF_IMAGE_DIFF_ACCUM_COLOR, ASVGF_IMAGE_SPEC_PING_COLOR,
          ASVGF_IMAGE_INDIR_PING );
    AntiFireflyArgs a;
    a.src[ 0 ] = d.get( ASVGF_IMAGE_DIFF_TEMPORARY, fi ); a.src[ 1 ] = d.get( ASVGF_IMAGE_SPEC_ACCUM_COLOR, fi ); a.src[ 2 ] = d.get( ASVGF_IMAGE_INDIR_ACCUM, fi );
    a.dst[ 0 ] = d.get( ASVGF_IMAGE_DIFF_ACCUM_COLOR, fi ); a.dst[ 1 ] = d.get( ASVGF_IMAGE_SPEC_PING_COLOR, fi ); a.dst[ 2 ] = d.get( ASVGF_IMAGE_INDIR_PING, fi );
    const dim3 grid( ( unsigned )( ( d.W + kAfTile - 1 ) / kAfTile ), ( unsigned )( ( u.yEnd - u.yBegin + kAfTile - 1 ) / kAfTile ), 3 );
    ASVGF_LITERAL_OR( d, ( k_antifirefly<<< gridFor( ( int )d.W, u.yEnd - u.yBegin ), kBlock, 0, s >>>( a, u ) ),
                      ( launchK( d, k_antifirefly_fast, grid, dim3( kAfThreads ), 0, s, a, u ) ) );
    return launchCheck( d, "k_antifirefly" );
}

AsvgfResult runVariance( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const Uniform& u )
{
    NEED( ASVGF_IMAGE_DIFF_ACCUM_COLOR, ASVGF_IMAGE_NORMAL, ASVGF_IMAGE_DEPTH_WORLD, ASVGF_IMAGE_ACCUM_HISTORY_LENGTH, ASVGF_IMAGE_DIFF_ACCUM_MOMENTS,
          ASVGF_IMAGE_DIFF_PING_COLOR_AND_VARIANCE );
    VarianceArgs a;
    a.color = d.get( ASVGF_IMAGE_DIFF_ACCUM_COLOR, fi ); a.normal = d.get( ASVGF_IMAGE_NORMAL, fi ); a.depth = d.get( ASVGF_IMAGE_DEPTH_WORLD, fi );
    a.histLen = d.get( ASVGF_IMAGE_ACCUM_HISTORY_LENGTH, fi ); a.moments = d.get( ASVGF_IMAGE_DIFF_ACCUM_MOMENTS, fi );
    a.out = d.get( ASVGF_IMAGE_DIFF_PING_COLOR_AND_VARIANCE, fi ); a.lut = d.lut;
    ASVGF_LITERAL_OR( d, ( k_variance<<< gridFor( ( int )d.W, u.yEnd - u.yBegin ), kBlock, 0, s >>>( a, u ) ),
                      ( launchK( d, k_variance_fast, gridFor( ( int )d.W, u.yEnd - u.yBegin ), kBlock, 0, s, a, u ) ) );
    return launchCheck( d, "k_variance" );
}

// decoded geometry of the frame (normal, depth, roughness) for the a-trous iterations, over every stored row
AsvgfResult runGuide( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi )
{
    NEED( ASVGF_IMAGE_NORMAL, ASVGF_IMAGE_DEPTH_WORLD, ASVGF_IMAGE_METALLIC_ROUGHNESS );
    GuideArgs a;
    a.normal = d.get( ASVGF_IMAGE_NORMAL, fi ); a.depth = d.get( ASVGF_IMAGE_DEPTH_WORLD, fi ); a.metRough = d.get( ASVGF_IMAGE_METALLIC_ROUGHNESS, fi );
    a.guide = d.guide; a.lut = d.lut;
    a.yBegin = d.row0; a.yEnd = d.row0 + d.rows;
    const dim3 grid( ( unsigned )( ( d.W + 255 ) / 256 ), ( unsigned )( ( d.rows + 3 ) / 4 ), 1 );
    launchK( d, k_guide, grid, dim3( 256 ), 0, s, a );
    return launchCheck( d, "k_guide" );
}

AsvgfResult runAtrous( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const Uniform& u, uint32_t iteration )
{
    if( iteration >= d.iterations ) return fail( ASVGF_WRONG_ARGUMENT, "a-trous iteration beyond atrousIterations" );
    const bool last = iteration + 1 == d.iterations && iteration > 0; // CmSVGFAtrous.comp:278: only CmSVGFAtrous.comp composes
    // routing: CmSVGFAtrous_Iter0.comp:312-334, CmSVGFAtrous.comp:237-276; beyond 4 iterations it keeps ping-ponging
    uint32_t sDiff, sSpec, sIndir, dDiff, dSpec, dIndir;
    if( iteration == 0 )
    {
        sDiff = ASVGF_IMAGE_DIFF_PING_COLOR_AND_VARIANCE; sSpec = ASVGF_IMAGE_SPEC_PING_COLOR; sIndir = ASVGF_IMAGE_INDIR_PING;
        dDiff = ASVGF_IMAGE_DIFF_COLOR_HISTORY; dSpec = ASVGF_IMAGE_SPEC_PONG_COLOR; dIndir = ASVGF_IMAGE_INDIR_PONG;
    }
    else if( iteration % 2 == 1 )
    {
        sDiff = iteration == 1 ? ASVGF_IMAGE_DIFF_COLOR_HISTORY : ASVGF_IMAGE_DIFF_PONG_COLOR_AND_VARIANCE;
        sSpec = ASVGF_IMAGE_SPEC_PONG_COLOR; sIndir = ASVGF_IMAGE_INDIR_PONG;
        dDiff = ASVGF_IMAGE_DIFF_PING_COLOR_AND_VARIANCE; dSpec = ASVGF_IMAGE_SPEC_PING_COLOR; dIndir = ASVGF_IMAGE_INDIR_PING;
    }
    else
    {
        sDiff = ASVGF_IMAGE_DIFF_PING_COLOR_AND_VARIANCE; sSpec = ASVGF_IMAGE_SPEC_PING_COLOR; sIndir = ASVGF_IMAGE_INDIR_PING;
        dDiff = ASVGF_IMAGE_DIFF_PONG_COLOR_AND_VARIANCE; dSpec = ASVGF_IMAGE_SPEC_PONG_COLOR; dIndir = ASVGF_IMAGE_INDIR_PONG;
    }
    NEED( sDiff, sSpec, sIndir, ASVGF_IMAGE_DEPTH_WORLD, ASVGF_IMAGE_DEPTH_GRAD, ASVGF_IMAGE_NORMAL, ASVGF_IMAGE_METALLIC_ROUGHNESS,
          ASVGF_IMAGE_ACCUM_HISTORY_LENGTH, ASVGF_IMAGE_ATROUS_FILTERED_VARIANCE );
    if( last ) NEED( ASVGF_IMAGE_ALBEDO, ASVGF_IMAGE_THROUGHPUT, ASVGF_IMAGE_IS_SKY, ASVGF_IMAGE_PRE_FINAL, ASVGF_IMAGE_HISTOGRAM_INPUT );
    else NEED( dDiff, dSpec, dIndir );

    AtrousArgs a;
    a.srcDiff = d.get( sDiff, fi ); a.srcSpec = d.get( sSpec, fi ); a.srcIndir = d.get( sIndir, fi );
    a.dstDiff = d.get( dDiff, fi ); a.dstSpec = d.get( dSpec, fi ); a.dstIndir = d.get( dIndir, fi );
    a.depth = d.get( ASVGF_IMAGE_DEPTH_WORLD, fi ); a.depthGrad = d.get( ASVGF_IMAGE_DEPTH_GRAD, fi ); a.normal = d.get( ASVGF_IMAGE_NORMAL, fi );
    a.metRough = d.get( ASVGF_IMAGE_METALLIC_ROUGHNESS, fi ); a.histLen = d.get( ASVGF_IMAGE_ACCUM_HISTORY_LENGTH, fi );
    a.filteredVariance = d.get( ASVGF_IMAGE_ATROUS_FILTERED_VARIANCE, fi );
    a.albedo = d.get( ASVGF_IMAGE_ALBEDO, fi ); a.throughput = d.get( ASVGF_IMAGE_THROUGHPUT, fi ); a.isSky = d.get( ASVGF_IMAGE_IS_SKY, fi );
    a.preFinal = d.get( ASVGF_IMAGE_PRE_FINAL, fi ); a.histogramInput = d.get( ASVGF_IMAGE_HISTOGRAM_INPUT, fi );
    a.lut = d.lut; a.iteration = ( int32_t )iteration;
    a.guide = d.guide;
    a.centre = d.centre; a.centreValid = d.centreValid ? 1 : 0;
    a.uniformPitch = ( a.guide.pitchT == a.srcDiff.pitchT && a.guide.pitchT == a.srcSpec.pitchT && a.guide.pitchT == a.srcIndir.pitchT ) ? 1 : 0;
    if( iteration == 0 ) d.centreValid = true; // this launch writes it

    const int step = 1 << iteration;
#ifdef ASVGF_WITH_LITERAL
    if( literal( d ) )
    {
        const dim3 grid = gridFor( ( int )d.W, u.yEnd - u.yBegin );
        if( iteration == 0 ) k_atrous< true, false ><<< grid, kBlock, 0, s >>>( a, u );
        else if( last ) k_atrous< false, true ><<< grid, kBlock, 0, s >>>( a, u );
        else k_atrous< false, false ><<< grid, kBlock, 0, s >>>( a, u );
        return launchCheck( d, "k_atrous" );
    }
#endif
    // dilated tiles: 64 columns x kAtTileH rows spaced `step` apart; grid.y enumerates (band of kAtTileH*step rows, residue)
    const int    rows   = u.yEnd - u.yBegin;
    const int    bands  = ( rows + kAtTileH * step - 1 ) / ( kAtTileH * step );
    const int    tilesX = ( int )( ( d.W + kAtTileW - 1 ) / kAtTileW );
    const size_t smem   = atrousHalfSmemBytes( step );
    const dim3   grid( ( unsigned )tilesX, ( unsigned )( bands * step ), 1 );
    // the opt-in to more than 48 KB of dynamic shared memory (strides >= 16) is per function and per device: once, not per launch
    auto launch = [ & ]( auto kernel, bool& optedIn ) -> cudaError_t {
        if( !optedIn )
        {
            cudaError_t e = cudaFuncSetAttribute( kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ( int )smem );
            if( e != cudaSuccess ) return e;
            optedIn = true;
        }
        launchK( d, kernel, grid, dim3( kAtThreads ), smem, s, a, u );
        return cudaSuccess;
    };
    cudaError_t e = cudaErrorInvalidValue;
    const bool  centre = a.centreValid != 0 && iteration > 0 && d.atrousCentre;
    bool&       opted = d.atrousOptIn[ iteration ][ ( last ? 1 : 0 ) + ( centre ? 2 : 0 ) ];
#define ASVGF_ATROUS_LAUNCH( S, I0, L ) ( centre ? launch( k_atrous_half< S, I0, L, true >, opted ) : launch( k_atrous_half< S, I0, L, false >, opted ) )
#define ASVGF_ATROUS_CASE( S ) \
    case S: e = last ? ASVGF_ATROUS_LAUNCH( S, false, true ) : ASVGF_ATROUS_LAUNCH( S, false, false ); break;
    switch( step )
    {
        case 1: e = launch( k_atrous_half< 1, true, false, false >, opted ); break; // iteration 0 is never the last (atrousIterations >= 2)
        ASVGF_ATROUS_CASE( 2 )
        ASVGF_ATROUS_CASE( 4 )
        ASVGF_ATROUS_CASE( 8 )
        ASVGF_ATROUS_CASE( 16 )
        ASVGF_ATROUS_CASE( 32 )
        ASVGF_ATROUS_CASE( 64 )
        ASVGF_ATROUS_CASE( 128 )
        default: break;
    }
#undef ASVGF_ATROUS_CASE
#undef ASVGF_ATROUS_LAUNCH
    if( e != cudaSuccess ) return cudaFail( e, "cudaFuncSetAttribute(k_atrous_half)" );
    return launchCheck( d, "k_atrous_half" );
}

AsvgfResult runResolveCheckerboard( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const Uniform& u )
{
    NEED( ASVGF_IMAGE_PRE_FINAL, ASVGF_IMAGE_THROUGHPUT, ASVGF_IMAGE_FINAL );
    const NvtxRange range( "Framebuf checkerboard interpolate" );
    ResolveArgs a;
    a.src[ 0 ] = d.get( ASVGF_IMAGE_PRE_FINAL, fi ); a.throughput = d.get( ASVGF_IMAGE_THROUGHPUT, fi ); a.dst[ 0 ] = d.get( ASVGF_IMAGE_FINAL, fi );
    // the shader's two other image pairs (CmCheckerboard.comp:63-64,70-77) when the caller keeps them in this table
    a.src[ 1 ] = d.get( ASVGF_IMAGE_SCREEN_EMIS_RT, fi ); a.dst[ 1 ] = d.get( ASVGF_IMAGE_SCREEN_EMISSION, fi );
    a.src[ 2 ] = d.get( ASVGF_IMAGE_ACID_FOG_RT, fi ); a.dst[ 2 ] = d.get( ASVGF_IMAGE_ACID_FOG, fi );
    a.pairs = ( a.src[ 1 ].p && a.dst[ 1 ].p && a.src[ 2 ].p && a.dst[ 2 ].p ) ? 3 : 1;
    ASVGF_LAUNCH( k_checkerboard_resolve, gridFor( ( int )d.W, u.yEnd - u.yBegin ), kBlock, 0, s, a, u );
    return launchCheck( d, "k_checkerboard_resolve" );
}

AsvgfResult runPrepareFinal( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const Uniform& u, const AsvgfGlobalUniform& g, const uint32_t* lpm )
{
    NEED( ASVGF_IMAGE_FINAL, ASVGF_IMAGE_SCREEN_EMISSION, ASVGF_IMAGE_SCATTERING, ASVGF_IMAGE_ACID_FOG, ASVGF_IMAGE_BLOOM_INPUT );
    if( g.debugShowFlags != 0 ) NEED( ASVGF_IMAGE_MOTION, ASVGF_IMAGE_DIS_PING_GRADIENT, ASVGF_IMAGE_NORMAL, ASVGF_IMAGE_UNFILTERED_DIRECT, ASVGF_IMAGE_UNFILTERED_SPECULAR, ASVGF_IMAGE_UNFILTERED_INDIR );
    const NvtxRange range( "Prefinal framebuf compose" ); // ImageComposition.cpp:116
    PrepareFinalArgs a;
    a.final = d.get( ASVGF_IMAGE_FINAL, fi ); a.screenEmission = d.get( ASVGF_IMAGE_SCREEN_EMISSION, fi ); a.scattering = d.get( ASVGF_IMAGE_SCATTERING, fi );
    a.acidFog = d.get( ASVGF_IMAGE_ACID_FOG, fi ); a.bloomInput = d.get( ASVGF_IMAGE_BLOOM_INPUT, fi );
    a.motion = d.get( ASVGF_IMAGE_MOTION, fi ); a.disPing = d.get( ASVGF_IMAGE_DIS_PING_GRADIENT, fi ); a.normal = d.get( ASVGF_IMAGE_NORMAL, fi );
    a.unfiltered[ 0 ] = d.get( ASVGF_IMAGE_UNFILTERED_DIRECT, fi ); a.unfiltered[ 1 ] = d.get( ASVGF_IMAGE_UNFILTERED_SPECULAR, fi ); a.unfiltered[ 2 ] = d.get( ASVGF_IMAGE_UNFILTERED_INDIR, fi );
    a.lut = d.lut;
    if( !d.finalScalars )
    {
        cudaError_t e = cudaMalloc( ( void** )&d.finalScalars, 4 * sizeof( float ) );
        if( e != cudaSuccess ) return cudaFail( e, "cudaMalloc(final-prepare scalars)" );
    }
    a.scalars = d.finalScalars;
    ASVGF_LAUNCH( k_prepare_final_scalars, dim3( 1 ), dim3( 1 ), 0, s, d.tm, g.bloomEmissionMultiplier, g.luminanceWhitePoint, d.finalScalars );
    {
        AsvgfResult r = launchCheck( d, "k_prepare_final_scalars" );
        if( r != ASVGF_SUCCESS ) return r;
    }
    // the members of the 24 x uvec4 control block LpmFilter hands to LpmMap in this configuration (LPM/ffx_lpm.h:913-938)
    auto f = [ & ]( int vec, int comp ) { float v; std::memcpy( &v, &lpm[ vec * 4 + comp ], 4 ); return v; };
    for( int i = 0; i < 3; i++ ) { a.saturation[ i ] = f( 0, i ); a.crosstalk[ i ] = f( 2, 1 + i ); a.rcpLumaT[ i ] = f( 3, i ); }
    a.contrast = f( 0, 3 ); a.toneScaleBias[ 0 ] = f( 1, 0 ); a.toneScaleBias[ 1 ] = f( 1, 1 );
    a.lumaT[ 0 ] = f( 1, 2 ); a.lumaT[ 1 ] = f( 1, 3 ); a.lumaT[ 2 ] = f( 2, 0 );
    a.luminanceWhitePoint = g.luminanceWhitePoint; a.emissionMaxScreenColor = g.emissionMaxScreenColor;
    a.bloomThreshold = g.bloomThreshold; a.bloomEmissionMultiplier = g.bloomEmissionMultiplier;
    a.classicEdge = g.lightmapScreenCoverage * g.renderWidth;
    a.debugShowFlags = g.debugShowFlags;
    ASVGF_LAUNCH( k_prepare_final, gridFor( ( int )d.W, u.yEnd - u.yBegin ), kBlock, 0, s, a, u );
    return launchCheck( d, "k_prepare_final" );
}

AsvgfResult runScatterAccumulate( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const Uniform& u, const AsvgfGlobalUniform& g, float maxHistoryLength )
{
    NEED( ASVGF_IMAGE_MOTION, ASVGF_IMAGE_DEPTH_WORLD, ASVGF_IMAGE_NORMAL, ASVGF_IMAGE_DEPTH_WORLD_PREV, ASVGF_IMAGE_NORMAL_PREV, ASVGF_IMAGE_SCATTERING_PREV,
          ASVGF_IMAGE_SCATTERING_HISTORY_PREV, ASVGF_IMAGE_SCATTERING, ASVGF_IMAGE_SCATTERING_HISTORY );
    if( g.volumeEnableType == ASVGF_VOLUME_ENABLE_VOLUMETRIC ) NEED( ASVGF_IMAGE_SCATTER_CURRENT );
    const NvtxRange range( "Volumetric Process" );
    ScatterArgs a;
    a.motion = d.get( ASVGF_IMAGE_MOTION, fi ); a.depth = d.get( ASVGF_IMAGE_DEPTH_WORLD, fi ); a.normal = d.get( ASVGF_IMAGE_NORMAL, fi );
    a.depthPrev = d.get( ASVGF_IMAGE_DEPTH_WORLD_PREV, fi ); a.normalPrev = d.get( ASVGF_IMAGE_NORMAL_PREV, fi );
    a.scatPrev = d.get( ASVGF_IMAGE_SCATTERING_PREV, fi ); a.histPrev = d.get( ASVGF_IMAGE_SCATTERING_HISTORY_PREV, fi );
    a.scat = d.get( ASVGF_IMAGE_SCATTERING, fi ); a.hist = d.get( ASVGF_IMAGE_SCATTERING_HISTORY, fi ); a.current = d.get( ASVGF_IMAGE_SCATTER_CURRENT, fi );
    a.lut = d.lut; a.maxHistoryLength = maxHistoryLength; a.volumeScattering = g.volumeScattering; a.volumeEnableType = g.volumeEnableType;
    a.transmittance = d.transmittance;
    if( g.volumeEnableType == ASVGF_VOLUME_ENABLE_SIMPLE )
    {
        const float density = g.volumeScattering * 0.00005f; // = mulr( volumeScattering, 0.00005f ) of the kernel (CmScatterAccum.comp:81)
        if( !d.transmittance )
        {
            cudaError_t e = cudaMalloc( ( void** )&d.transmittance, 65536 * sizeof( float ) );
            if( e != cudaSuccess ) return cudaFail( e, "cudaMalloc(transmittance table)" );
            a.transmittance = d.transmittance;
        }
        if( std::memcmp( &density, &d.transmittanceDensity, sizeof( float ) ) != 0 ) // rebuilt only when the density changes
        {
            ASVGF_LAUNCH( k_scatter_transmittance_table, dim3( 256 ), dim3( 256 ), 0, s, d.transmittance, density );
            AsvgfResult r = launchCheck( d, "k_scatter_transmittance_table" );
            if( r != ASVGF_SUCCESS ) return r;
            d.transmittanceDensity = density;
        }
    }
    for( int i = 0; i < 3; i++ ) a.volumeAmbient[ i ] = g.volumeAmbient[ i ];
    ASVGF_LAUNCH( k_scatter_accum, gridFor( ( int )d.W, u.yEnd - u.yBegin ), kBlock, 0, s, a, u );
    return launchCheck( d, "k_scatter_accum" );
}

LumArgs lumArgs( const AsvgfDenoiser_T& d, uint32_t fi, const AsvgfGlobalUniform& g )
{
    LumArgs a;
    a.histogramInput = d.get( ASVGF_IMAGE_HISTOGRAM_INPUT, fi );
    a.log2Half = d.log2Half; a.binTable = d.binTable; a.tm = d.tm;
    a.minLogLuminance = g.minLogLuminance; a.maxLogLuminance = g.maxLogLuminance; a.timeDelta = g.timeDelta;
    a.pixelCount = g.renderWidth * g.renderHeight; // CmLuminanceAvg.comp:68
    a.stopEyeAdaptation = g.stopEyeAdaptation;
    return a;
}

AsvgfResult runLuminanceHistogram( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const Uniform& u, const AsvgfGlobalUniform& g )
{
    NEED( ASVGF_IMAGE_HISTOGRAM_INPUT );
    const NvtxRange range( "Exposure" );
    const LumArgs a = lumArgs( d, fi, g );
    ASVGF_LAUNCH( k_lum_bin_table, dim3( 256 ), dim3( 256 ), 0, s, a );
    AsvgfResult r = launchCheck( d, "k_lum_bin_table" );
    if( r != ASVGF_SUCCESS ) return r;
    const size_t smem   = 65536 + ( kLumThreads / 32 ) * 256 * sizeof( uint32_t );
    const long   groups = ( long )( u.yEnd - u.yBegin ) * ( ( d.W + 7 ) / 8 );
    long         ctas   = ( groups + kLumThreads - 1 ) / kLumThreads;
    if( ctas > 3L * d.smCount ) ctas = 3L * d.smCount; // 73 KB of shared memory per CTA: three per SM
    if( ctas < 1 ) ctas = 1;
    cudaError_t e = cudaFuncSetAttribute( k_lum_histogram, cudaFuncAttributeMaxDynamicSharedMemorySize, ( int )smem );
    if( e != cudaSuccess ) return cudaFail( e, "cudaFuncSetAttribute(k_lum_histogram)" );
    ASVGF_LAUNCH( k_lum_histogram, dim3( ( unsigned )ctas ), dim3( kLumThreads ), smem, s, a, u );
    return launchCheck( d, "k_lum_histogram" );
}

AsvgfResult runLuminanceAverage( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const AsvgfGlobalUniform& g )
{
    const LumArgs a = lumArgs( d, fi, g );
    ASVGF_LAUNCH( k_lum_average, dim3( 1 ), dim3( 256 ), 0, s, a );
    return launchCheck( d, "k_lum_average" );
}

// ASVGF_CREATE_BAND_APRON: rows beyond the owned band that a pass computes redundantly, so that every later pass of the
// frame finds its inputs without a halo exchange.  A pass that reads +-r rows needs its producer to be valid r rows
// further out: a-trous i reads +-2^i (CmSVGFAtrous.comp:61,126-140), iteration 0 and the variance pass +-1
// (CmSVGFAtrous_Iter0.comp:121, CmSVGFEstimateVariance.comp:59), anti-firefly +-2 (CmAntiFirefly.comp:31); the temporal
// pass reads the gradient image at row/3, gradient a-trous i reads +-2^i gradient rows (CmASVGFGradientAtrous.comp:57-60).
// Gradient samples are computed on owned rows only (their output is exchanged once, asvgfBandApronRows(.., SAMPLES) deep).
int apronRows( uint32_t iterations, uint32_t pass )
{
    int extA[ ASVGF_MAX_ATROUS_ITERATIONS ] = {};
    for( int i = ( int )iterations - 1; i >= 1; i-- ) extA[ i - 1 ] = extA[ i ] + ( 1 << i );
    const int extV = extA[ 0 ] + 1, extF = extV + 1, extT = extF + 2;
    if( pass >= ASVGF_PASS_ATROUS_0 && pass < ASVGF_PASS_ATROUS_0 + iterations ) return extA[ pass - ASVGF_PASS_ATROUS_0 ];
    if( pass == ASVGF_PASS_VARIANCE ) return extV;
    if( pass == ASVGF_PASS_ANTIFIREFLY ) return extF;
    if( pass == ASVGF_PASS_TEMPORAL ) return extT;
    int g = ( extT + 2 ) / 3; // gradient rows the temporal pass touches beyond the band
    if( pass == ASVGF_PASS_GRADIENT_ATROUS_3 ) return g;
    for( int i = 3; i >= 1; i-- )
    {
        g += 1 << i;
        if( pass == ( uint32_t )( ASVGF_PASS_GRADIENT_ATROUS_0 + i - 1 ) ) return g;
    }
    if( pass == ASVGF_PASS_GRADIENT_SAMPLES ) return g + 1; // depth of the one exchange of its output
    return 0;
}

AsvgfResult runPassRows( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const Uniform& u, uint32_t pass );

AsvgfResult runPass( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const Uniform& u0, uint32_t pass )
{
    if( pass == ASVGF_PASS_GRADIENT_ATROUS_ALL )
    {
        for( uint32_t i = 0; i < ASVGF_GRADIENT_ATROUS_ITERATIONS; i++ )
        {
            AsvgfResult r = runPass( d, s, fi, u0, ASVGF_PASS_GRADIENT_ATROUS_0 + i );
            if( r != ASVGF_SUCCESS ) return r;
        }
        return ASVGF_SUCCESS;
    }
    Uniform u = u0;
    d.passGyBegin = d.gownBegin; d.passGyEnd = d.gownEnd;
    if( ( d.info.flags & ASVGF_CREATE_BAND_APRON ) && d.info.bandRowEnd != 0 )
    {
        const int ext = apronRows( d.iterations, pass );
        if( pass >= ASVGF_PASS_GRADIENT_ATROUS_0 && pass <= ASVGF_PASS_GRADIENT_ATROUS_3 )
        {
            d.passGyBegin = std::max( d.grow0, d.gownBegin - ext );
            d.passGyEnd   = std::min( d.grow0 + d.grows, d.gownEnd + ext );
        }
        else if( pass != ASVGF_PASS_GRADIENT_SAMPLES )
        {
            u.yBegin = std::max( d.row0, d.ownBegin - ext );
            u.yEnd   = std::min( d.row0 + d.rows, d.ownEnd + ext );
        }
    }
    return runPassRows( d, s, fi, u, pass );
}

AsvgfResult runPassRows( AsvgfDenoiser_T& d, cudaStream_t s, uint32_t fi, const Uniform& u, uint32_t pass )
{
    const NvtxRange range( pass == ASVGF_PASS_GRADIENT_SAMPLES ? "Gradient samples"
                           : pass <= ASVGF_PASS_GRADIENT_ATROUS_3 ? "Gradient Atrous"
                           : pass == ASVGF_PASS_TEMPORAL          ? "Temporal accumulation"
                           : pass == ASVGF_PASS_ANTIFIREFLY       ? "Antifirefly"
                           : pass == ASVGF_PASS_VARIANCE          ? "SVGF Variance estimation"
                                                                  : "SVGF Atrous" );
    if( pass == ASVGF_PASS_GRADIENT_SAMPLES ) return runGradientSamples( d, s, fi, u );
    if( pass >= ASVGF_PASS_GRADIENT_ATROUS_0 && pass <= ASVGF_PASS_GRADIENT_ATROUS_3 ) return runGradientAtrous( d, s, fi, u, pass - ASVGF_PASS_GRADIENT_ATROUS_0 );
    if( pass == ASVGF_PASS_TEMPORAL ) return runTemporal( d, s, fi, u );
    if( pass == ASVGF_PASS_ANTIFIREFLY ) return runAntiFirefly( d, s, fi, u );
    if( pass == ASVGF_PASS_VARIANCE ) return runVariance( d, s, fi, u );
    if( pass >= ASVGF_PASS_ATROUS_0 && pass < ASVGF_PASS_ATROUS_0 + d.iterations ) return runAtrous( d, s, fi, u, pass - ASVGF_PASS_ATROUS_0 );
    return fail( ASVGF_WRONG_ARGUMENT, "unknown pass " + std::to_string( pass ) );
}

// the launches of one API call form a chain of programmatic dependent launches (whole-frame handles only)
struct PdlScope
{
    AsvgfDenoiser_T& d;
    explicit PdlScope( AsvgfDenoiser_T& d_ ) : d( d_ ) { d.pdlChain = d.pdl && d.info.bandRowEnd == 0; d.pdlArmed = false; }
    ~PdlScope() { d.pdlChain = d.pdlArmed = false; }
};

AsvgfResult checkCall( AsvgfDenoiser d, uint32_t frameIndex, const AsvgfGlobalUniform* pUniform )
{
    if( !d ) return fail( ASVGF_WRONG_ARGUMENT, "null denoiser" );
    if( frameIndex > 1 ) return fail( ASVGF_WRONG_ARGUMENT, "frameIndex must be 0 or 1 (MAX_FRAMES_IN_FLIGHT)" );
    if( !pUniform ) return fail( ASVGF_WRONG_ARGUMENT, "null uniform" );
    if( ( uint32_t )pUniform->renderWidth != d->W || ( uint32_t )pUniform->renderHeight != d->H )
        return fail( ASVGF_WRONG_ARGUMENT, "uniform renderWidth/renderHeight differ from the created size" );
    return ASVGF_SUCCESS;
}

} // namespace

extern "C" {

AsvgfResult asvgfGetImageInfo( uint32_t imageIndex, uint32_t W, uint32_t H, uint32_t* pFormat, uint32_t* pWidth, uint32_t* pHeight, uint32_t* pBpp )
{
    ImageDesc desc;
    if( imageIndex >= ASVGF_IMAGE_COUNT || !describe( imageIndex, &desc ) ) return fail( ASVGF_WRONG_ARGUMENT, "image index not used by the denoiser path" );
    if( pFormat ) *pFormat = desc.format;
    if( pWidth ) *pWidth = desc.thirdRes ? ( ( W + 1 ) / 3 > 1 ? ( W + 1 ) / 3 : 1 ) : W; // Framebuffers.cpp:691-699
    if( pHeight ) *pHeight = desc.thirdRes ? ( ( H + 1 ) / 3 > 1 ? ( H + 1 ) / 3 : 1 ) : H;
    if( pBpp ) *pBpp = desc.bpp;
    return ASVGF_SUCCESS;
}

AsvgfResult asvgfCreate( const AsvgfCreateInfo* pInfo, AsvgfDenoiser* pDenoiser )
{
    if( !pInfo || !pDenoiser ) return fail( ASVGF_WRONG_ARGUMENT, "null argument" );
    *pDenoiser = nullptr;
    if( pInfo->renderWidth == 0 || pInfo->renderHeight == 0 || ( pInfo->renderWidth & 1u ) ) return fail( ASVGF_WRONG_ARGUMENT, "renderWidth must be even and non-zero" );
    if( pInfo->renderWidth > 32768 || pInfo->renderHeight > 32768 ) return fail( ASVGF_WRONG_ARGUMENT, "render size too large" );
    if( pInfo->atrousIterations < 2 || pInfo->atrousIterations > ASVGF_MAX_ATROUS_ITERATIONS ) return fail( ASVGF_WRONG_ARGUMENT, "atrousIterations must be 2..8" );

#ifndef ASVGF_WITH_LITERAL
    if( pInfo->flags & ASVGF_CREATE_LITERAL_KERNELS ) return fail( ASVGF_UNSUPPORTED, "the literal kernels are not part of this build (make LITERAL=1)" );
#endif
    if( ( pInfo->flags & ASVGF_CREATE_LITERAL_KERNELS ) && pInfo->bandRowEnd != 0 )
        return fail( ASVGF_UNSUPPORTED, "the literal kernels bounds-check against the frame, not the stored rows: no band handles" );
    if( pInfo->bandRowEnd != 0 )
    {
        // a pass reads up to 2^(N-1) rows beyond the rows it computes (the last a-trous iteration); with the apron schedule the
        // first pass of the chain computes apron(TEMPORAL) rows beyond the band and the gradient passes 3x their gradient rows
        uint32_t need = 1u << ( pInfo->atrousIterations - 1 );
        if( pInfo->flags & ASVGF_CREATE_BAND_APRON )
            need = std::max( need, ( uint32_t )std::max( apronRows( pInfo->atrousIterations, ASVGF_PASS_TEMPORAL ), 3 * apronRows( pInfo->atrousIterations, ASVGF_PASS_GRADIENT_SAMPLES ) ) );
        if( pInfo->bandHaloRows < need )
            return fail( ASVGF_WRONG_ARGUMENT, "bandHaloRows is smaller than the vertical reach of the passes (" + std::to_string( need ) + " rows)" );
    }

    int deviceCount = 0;
    if( cudaGetDeviceCount( &deviceCount ) != cudaSuccess || deviceCount == 0 )
    {
        cudaGetLastError();
        return fail( ASVGF_CUDA_ERROR, "no CUDA device: libasvgf has no CPU fallback" );
    }
    int device = pInfo->device;
    if( device < 0 ) CUDA_TRY( cudaGetDevice( &device ) );
    if( device >= deviceCount ) return fail( ASVGF_WRONG_ARGUMENT, "device ordinal out of range" );

    AsvgfDenoiser_T* d = new( std::nothrow ) AsvgfDenoiser_T();
    if( !d ) return fail( ASVGF_OUT_OF_MEMORY, "host allocation failed" );
    d->info = *pInfo; d->device = device;
    d->W = pInfo->renderWidth; d->H = pInfo->renderHeight; d->iterations = pInfo->atrousIterations;
    if( const char* e = std::getenv( "ASVGF_ATROUS_CENTRE" ) ) d->atrousCentre = std::atoi( e ) != 0;
    if( const char* e = std::getenv( "ASVGF_PDL" ) ) d->pdl = std::atoi( e ) != 0;
    const int32_t H = ( int32_t )d->H, gH = ( int32_t )( ( d->H + 1 ) / 3 > 1 ? ( d->H + 1 ) / 3 : 1 );

    if( pInfo->bandRowEnd == 0 )
    {
        d->ownBegin = 0; d->ownEnd = H; d->row0 = 0; d->rows = H;
        d->gownBegin = 0; d->gownEnd = gH; d->grow0 = 0; d->grows = gH;
    }
    else
    {
        if( pInfo->bandRowBegin >= pInfo->bandRowEnd || pInfo->bandRowEnd > d->H || pInfo->bandRowBegin % 3 || ( pInfo->bandRowEnd % 3 && pInfo->bandRowEnd != d->H ) ||
            pInfo->bandHaloRows % 3 )
        {
            delete d;
            return fail( ASVGF_WRONG_ARGUMENT, "band rows must be ordered, inside the frame and multiples of 3" );
        }
        const int32_t halo = ( int32_t )pInfo->bandHaloRows;
        d->ownBegin = ( int32_t )pInfo->bandRowBegin; d->ownEnd = ( int32_t )pInfo->bandRowEnd;
        d->row0 = d->ownBegin - halo > 0 ? d->ownBegin - halo : 0;
        const int32_t rend = d->ownEnd + halo < H ? d->ownEnd + halo : H;
        d->rows = rend - d->row0;
        d->gownBegin = d->ownBegin / 3;
        d->gownEnd   = d->ownEnd == H ? gH : d->ownEnd / 3;
        d->grow0     = d->gownBegin - halo / 3 > 0 ? d->gownBegin - halo / 3 : 0;
        const int32_t grend = d->gownEnd + halo / 3 < gH ? d->gownEnd + halo / 3 : gH;
        d->grows = grend - d->grow0;
    }

    DeviceGuard guard( device );
    if( !guard.ok ) { delete d; return fail( ASVGF_CUDA_ERROR, "cudaSetDevice failed" ); }

    AsvgfResult r = acquireLut( device, &d->lut );
    if( r != ASVGF_SUCCESS ) { delete d; return r; }
    d->haveLut = true;
    if( cudaDeviceGetAttribute( &d->smCount, cudaDevAttrMultiProcessorCount, device ) != cudaSuccess || d->smCount <= 0 ) d->smCount = 148;
    {
        // CmLuminanceHistogram.comp:50 takes log2 of an fp16 texel: 65536 possible arguments.  GLSL leaves the precision of
        // log2 to the implementation; the pin (shared with the oracle) is the correctly rounded fp32 value.
        std::vector< float > table( 65536 );
        for( uint32_t h = 0; h < 65536; h++ )
        {
            const uint32_t sign = h >> 15, e = ( h >> 10 ) & 31u, m = h & 1023u;
            double         v;
            if( e == 0 ) v = std::ldexp( ( double )m, -24 );
            else if( e == 31 ) v = m ? std::nan( "" ) : HUGE_VAL;
            else v = std::ldexp( 1.0 + ( double )m / 1024.0, ( int )e - 15 );
            if( sign ) v = -v;
            table[ h ] = ( float )std::log2( v );
        }
        cudaError_t e = cudaMalloc( ( void** )&d->log2Half, 65536 * sizeof( float ) );
        if( e == cudaSuccess ) e = cudaMalloc( ( void** )&d->binTable, 65536 );
        if( e == cudaSuccess ) e = cudaMalloc( ( void** )&d->tm, sizeof( AsvgfTonemapping ) );
        if( e == cudaSuccess ) e = cudaMemcpy( d->log2Half, table.data(), 65536 * sizeof( float ), cudaMemcpyHostToDevice );
        if( e == cudaSuccess ) e = cudaMemset( d->tm, 0, sizeof( AsvgfTonemapping ) );
        if( e != cudaSuccess )
        {
            AsvgfResult rr = cudaFail( e, "cudaMalloc(tonemapping)" );
            std::string msg = g_lastError;
            asvgfDestroy( d );
            return fail( rr, msg );
        }
    }

    for( uint32_t i = 0; i < ASVGF_IMAGE_COUNT; i++ )
    {
        ImageDesc desc;
        if( !describe( i, &desc ) ) continue;
        Img& im = d->img[ i ];
        uint32_t w, h;
        asvgfGetImageInfo( i, d->W, d->H, nullptr, &w, &h, nullptr );
        im.w = ( int32_t )w; im.h = ( int32_t )h;
        im.row0 = desc.thirdRes ? d->grow0 : d->row0;
        im.rows = desc.thirdRes ? d->grows : d->rows;
        // every image of one resolution holds the same number of texels per row, align256(width): the pitch is a multiple of 256 B for
        // every texel size, and kernels that touch many images at one pixel compute a single texel offset (asvgf_temporal.cuh)
        setImgStorage( im, nullptr, ( ( w + 255u ) & ~255u ) * desc.bpp, desc.bpp );
        if( pInfo->flags & ASVGF_CREATE_EXTERNAL_IMAGES ) continue;
        if( i == ASVGF_IMAGE_DEBUG_MASK && !( pInfo->flags & ASVGF_CREATE_EXPORT_DEBUG_MASK ) ) continue;
        const size_t bytes = ( size_t )im.pitch * ( size_t )im.rows;
        uint8_t*     mem   = nullptr;
        cudaError_t  e     = cudaMalloc( ( void** )&mem, bytes );
        if( e == cudaSuccess )
        {
            setImgStorage( im, mem, im.pitch, desc.bpp );
            d->owned[ i ] = true;
            e = cudaMemset( mem, 0, bytes );
        }
        if( e != cudaSuccess )
        {
            AsvgfResult rr = cudaFail( e, "cudaMalloc(image)" );
            std::string msg = g_lastError;
            asvgfDestroy( d );
            return fail( rr, msg );
        }
    }
    {
        Img& g = d->guide;
        g.w = ( int32_t )d->W; g.h = ( int32_t )d->H; g.row0 = d->row0; g.rows = d->rows;
        const uint32_t pitch = ( ( d->W + 255u ) & ~255u ) * 16u;
        uint8_t*       mem   = nullptr;
        cudaError_t    e     = cudaMalloc( ( void** )&mem, ( size_t )pitch * ( size_t )g.rows );
        if( e == cudaSuccess ) e = cudaMemset( mem, 0, ( size_t )pitch * ( size_t )g.rows );
        if( e == cudaSuccess ) e = cudaEventCreateWithFlags( &d->evTemporalDone, cudaEventDisableTiming );
        if( e == cudaSuccess ) e = cudaEventCreateWithFlags( &d->evGradientsDone, cudaEventDisableTiming );
        if( e != cudaSuccess )
        {
            if( mem ) cudaFree( mem );
            AsvgfResult rr = cudaFail( e, "cudaMalloc(guide plane)" );
            std::string msg = g_lastError;
            asvgfDestroy( d );
            return fail( rr, msg );
        }
        setImgStorage( g, mem, pitch, 16 );
        Img& c = d->centre;
        c = g;
        uint8_t* memC = nullptr;
        if( cudaMalloc( ( void** )&memC, ( size_t )pitch * ( size_t )g.rows ) != cudaSuccess || cudaMemset( memC, 0, ( size_t )pitch * ( size_t )g.rows ) != cudaSuccess )
        {
            if( memC ) cudaFree( memC );
            asvgfDestroy( d );
            return fail( ASVGF_OUT_OF_MEMORY, "cudaMalloc(centre plane)" );
        }
        setImgStorage( c, memC, pitch, 16 );
        if( pInfo->bandRowEnd != 0 )
        {
            if( cudaMalloc( ( void** )&d->haloMiss, sizeof( uint32_t ) ) != cudaSuccess || cudaMemset( d->haloMiss, 0, sizeof( uint32_t ) ) != cudaSuccess )
            {
                asvgfDestroy( d );
                return fail( ASVGF_OUT_OF_MEMORY, "cudaMalloc(halo flag)" );
            }
        }
    }
    *pDenoiser = d;
    return ASVGF_SUCCESS;
}

AsvgfResult asvgfDestroy( AsvgfDenoiser d )
{
    if( !d ) return ASVGF_SUCCESS;
    DeviceGuard guard( d->device );
    for( uint32_t i = 0; i < ASVGF_IMAGE_COUNT; i++ )
        if( d->owned[ i ] && d->img[ i ].p ) cudaFree( d->img[ i ].p );
    if( d->haveLut ) releaseLut( d->device );
    if( d->guide.p ) cudaFree( d->guide.p );
    if( d->centre.p ) cudaFree( d->centre.p );
    if( d->haloMiss ) cudaFree( d->haloMiss );
    if( d->evTemporalDone ) cudaEventDestroy( d->evTemporalDone );
    if( d->evGradientsDone ) cudaEventDestroy( d->evGradientsDone );
    if( d->log2Half ) cudaFree( d->log2Half );
    if( d->binTable ) cudaFree( d->binTable );
    if( d->transmittance ) cudaFree( d->transmittance );
    if( d->finalScalars ) cudaFree( d->finalScalars );
    if( d->tm ) cudaFree( d->tm );
    delete d;
    return ASVGF_SUCCESS;
}

AsvgfResult asvgfGetImage( AsvgfDenoiser d, uint32_t imageIndex, uint32_t frameIndex, void** ppDevice, size_t* pPitch, uint32_t* pWidth, uint32_t* pFirstRow,
                           uint32_t* pRows )
{
    ImageDesc desc;
    if( !d || frameIndex > 1 || imageIndex >= ASVGF_IMAGE_COUNT || !describe( imageIndex, &desc ) ) return fail( ASVGF_WRONG_ARGUMENT, "bad image index / frameIndex" );
    const Img& im = d->get( imageIndex, frameIndex );
    if( ppDevice ) *ppDevice = im.p;
    if( pPitch ) *pPitch = im.pitch;
    if( pWidth ) *pWidth = ( uint32_t )im.w;
    if( pFirstRow ) *pFirstRow = ( uint32_t )im.row0;
    if( pRows ) *pRows = ( uint32_t )im.rows;
    return ASVGF_SUCCESS;
}

AsvgfResult asvgfBindExternalImage( AsvgfDenoiser d, uint32_t imageIndex, void* pDevice, size_t pitchBytes )
{
    ImageDesc desc;
    if( !d || imageIndex >= ASVGF_IMAGE_COUNT || !describe( imageIndex, &desc ) ) return fail( ASVGF_WRONG_ARGUMENT, "bad image index" );
    Img& im = d->img[ imageIndex ];
    if( !pDevice || pitchBytes < ( size_t )im.w * desc.bpp || pitchBytes % desc.bpp || pitchBytes > 0xffffffffull )
        return fail( ASVGF_WRONG_ARGUMENT, "bad pointer or pitch" );
    // the kernels move 8- and 16-byte vectors (two E5 texels, four fp16 channels, TMA boxes): rows must keep that alignment
    if( ( ( uintptr_t )pDevice % 16 ) || ( pitchBytes % 16 ) ) return fail( ASVGF_WRONG_ARGUMENT, "pointer and pitch must be multiples of 16 bytes" );
    if( d->owned[ imageIndex ] && im.p )
    {
        DeviceGuard guard( d->device );
        cudaFree( im.p );
        d->owned[ imageIndex ] = false;
    }
    setImgStorage( im, ( uint8_t* )pDevice, ( uint32_t )pitchBytes, desc.bpp );
    d->guideValid = d->centreValid = false;
    return ASVGF_SUCCESS;
}

static AsvgfResult copyImage( AsvgfDenoiser d, AsvgfStream stream, uint32_t imageIndex, uint32_t frameIndex, void* pHost, size_t hostPitch, bool upload )
{
    ImageDesc desc;
    if( !d || frameIndex > 1 || !pHost || imageIndex >= ASVGF_IMAGE_COUNT || !describe( imageIndex, &desc ) ) return fail( ASVGF_WRONG_ARGUMENT, "bad argument" );
    const Img& im = d->get( imageIndex, frameIndex );
    if( !im.p ) return fail( ASVGF_NOT_BOUND, "image has no storage" );
    const size_t rowBytes = ( size_t )im.w * desc.bpp;
    if( hostPitch == 0 ) hostPitch = rowBytes;
    if( hostPitch < rowBytes ) return fail( ASVGF_WRONG_ARGUMENT, "host pitch smaller than a row" );
    DeviceGuard guard( d->device );
    if( upload ) d->guideValid = d->centreValid = false;
    if( upload ) CUDA_TRY( cudaMemcpy2DAsync( im.p, im.pitch, pHost, hostPitch, rowBytes, ( size_t )im.rows, cudaMemcpyHostToDevice, ( cudaStream_t )stream ) );
    else CUDA_TRY( cudaMemcpy2DAsync( pHost, hostPitch, im.p, im.pitch, rowBytes, ( size_t )im.rows, cudaMemcpyDeviceToHost, ( cudaStream_t )stream ) );
    return ASVGF_SUCCESS;
}
AsvgfResult asvgfUploadImage( AsvgfDenoiser d, AsvgfStream stream, uint32_t imageIndex, uint32_t frameIndex, const void* pHost, size_t hostPitch )
{
    return copyImage( d, stream, imageIndex, frameIndex, const_cast< void* >( pHost ), hostPitch, true );
}
AsvgfResult asvgfDownloadImage( AsvgfDenoiser d, AsvgfStream stream, uint32_t imageIndex, uint32_t frameIndex, void* pHost, size_t hostPitch )
{
    return copyImage( d, stream, imageIndex, frameIndex, pHost, hostPitch, false );
}

AsvgfResult asvgfResetHistory( AsvgfDenoiser d, AsvgfStream stream )
{
    if( !d ) return fail( ASVGF_WRONG_ARGUMENT, "null denoiser" );
    DeviceGuard guard( d->device );
    for( uint32_t i = 0; i < ASVGF_IMAGE_COUNT; i++ )
        if( d->owned[ i ] && d->img[ i ].p ) CUDA_TRY( cudaMemsetAsync( d->img[ i ].p, 0, ( size_t )d->img[ i ].pitch * d->img[ i ].rows, ( cudaStream_t )stream ) );
    d->guideValid = d->centreValid = false;
    return ASVGF_SUCCESS;
}

// ---- checkpoint / resume of the temporal state (the reference keeps it in GPU images only) ----
namespace
{
// Everything but the producer's per-frame inputs: both members of every cur/prev pair (history, and the previous frame's G-buffer),
// the a-trous colour history, the gradient history -- what frame t+1 reads of frame t -- and also the images the passes only
// partly overwrite (GradientPrevPix keeps the slot of strata without a sample, AtrousFilteredVariance the value of pixels that
// turned into sky), so that a resumed handle's whole table equals the one that kept running; plus the exposure state.
bool isStateImage( uint32_t i )
{
    switch( i )
    {
        case ASVGF_IMAGE_ALBEDO: case ASVGF_IMAGE_IS_SKY: case ASVGF_IMAGE_DEPTH_GRAD: case ASVGF_IMAGE_MOTION: case ASVGF_IMAGE_UNFILTERED_DIRECT:
        case ASVGF_IMAGE_UNFILTERED_SPECULAR: case ASVGF_IMAGE_UNFILTERED_INDIR: case ASVGF_IMAGE_VIEW_DIRECTION: case ASVGF_IMAGE_THROUGHPUT:
        case ASVGF_IMAGE_FWD_LUM: case ASVGF_IMAGE_SCATTER_CURRENT: case ASVGF_IMAGE_ACID_FOG_RT: case ASVGF_IMAGE_SCREEN_EMIS_RT: return false;
        default: return true;
    }
}
struct StateHeader { uint32_t magic, version, width, height, row0, rows, grow0, grows, imageCount, tonemappingBytes; };
constexpr uint32_t kStateMagic = 0x53565341u; // "ASVS"
size_t stateBytes( const AsvgfDenoiser_T& d, uint32_t* pCount )
{
    size_t   n = sizeof( StateHeader ) + sizeof( AsvgfTonemapping );
    uint32_t count = 0;
    ImageDesc desc;
    for( uint32_t i = 0; i < ASVGF_IMAGE_COUNT; i++ )
        if( isStateImage( i ) && describe( i, &desc ) && d.img[ i ].p ) { n += 2 * sizeof( uint32_t ) + ( size_t )d.img[ i ].w * desc.bpp * d.img[ i ].rows; count++; }
    if( pCount ) *pCount = count;
    return n;
}
}

AsvgfResult asvgfGetStateSize( AsvgfDenoiser d, size_t* pBytes )
{
    if( !d || !pBytes ) return fail( ASVGF_WRONG_ARGUMENT, "null argument" );
    *pBytes = stateBytes( *d, nullptr );
    return ASVGF_SUCCESS;
}

static AsvgfResult copyState( AsvgfDenoiser d, AsvgfStream stream, uint8_t* host, size_t bytes, bool save )
{
    if( !d || !host ) return fail( ASVGF_WRONG_ARGUMENT, "null argument" );
    uint32_t     count = 0;
    const size_t need  = stateBytes( *d, &count );
    if( bytes < need ) return fail( ASVGF_WRONG_ARGUMENT, "state buffer too small (" + std::to_string( need ) + " bytes needed)" );
    DeviceGuard        guard( d->device );
    const cudaStream_t s = ( cudaStream_t )stream;
    StateHeader        h = { kStateMagic, 1u, d->W, d->H, ( uint32_t )d->row0, ( uint32_t )d->rows, ( uint32_t )d->grow0, ( uint32_t )d->grows, count, ( uint32_t )sizeof( AsvgfTonemapping ) };
    if( save ) std::memcpy( host, &h, sizeof( h ) );
    else
    {
        StateHeader g;
        std::memcpy( &g, host, sizeof( g ) );
        if( std::memcmp( &g, &h, sizeof( h ) ) != 0 ) return fail( ASVGF_WRONG_ARGUMENT, "the state was saved by a handle of another size / band / image set" );
    }
    size_t off = sizeof( StateHeader );
    if( save ) CUDA_TRY( cudaMemcpyAsync( host + off, d->tm, sizeof( AsvgfTonemapping ), cudaMemcpyDeviceToHost, s ) );
    else CUDA_TRY( cudaMemcpyAsync( d->tm, host + off, sizeof( AsvgfTonemapping ), cudaMemcpyHostToDevice, s ) );
    off += sizeof( AsvgfTonemapping );
    ImageDesc desc;
    for( uint32_t i = 0; i < ASVGF_IMAGE_COUNT; i++ )
    {
        if( !isStateImage( i ) || !describe( i, &desc ) || !d->img[ i ].p ) continue;
        const Img&     im       = d->img[ i ];
        const size_t   rowBytes = ( size_t )im.w * desc.bpp;
        const uint32_t tag[ 2 ] = { i, ( uint32_t )( rowBytes * im.rows ) };
        if( save ) std::memcpy( host + off, tag, sizeof( tag ) );
        else if( std::memcmp( host + off, tag, sizeof( tag ) ) != 0 ) return fail( ASVGF_WRONG_ARGUMENT, "corrupt state buffer" );
        off += sizeof( tag );
        if( save ) CUDA_TRY( cudaMemcpy2DAsync( host + off, rowBytes, im.p, im.pitch, rowBytes, ( size_t )im.rows, cudaMemcpyDeviceToHost, s ) );
        else CUDA_TRY( cudaMemcpy2DAsync( im.p, im.pitch, host + off, rowBytes, rowBytes, ( size_t )im.rows, cudaMemcpyHostToDevice, s ) );
        off += rowBytes * im.rows;
    }
    if( !save ) d->guideValid = d->centreValid = false;
    return ASVGF_SUCCESS;
}
AsvgfResult asvgfSaveState( AsvgfDenoiser d, AsvgfStream stream, void* pHost, size_t bytes ) { return copyState( d, stream, ( uint8_t* )pHost, bytes, true ); }
AsvgfResult asvgfLoadState( AsvgfDenoiser d, AsvgfStream stream, const void* pHost, size_t bytes ) { return copyState( d, stream, ( uint8_t* )const_cast< void* >( pHost ), bytes, false ); }

AsvgfResult asvgfGradientSamples( AsvgfDenoiser d, AsvgfStream stream, uint32_t frameIndex, const AsvgfGlobalUniform* pUniform )
{
    AsvgfResult r = checkCall( d, frameIndex, pUniform );
    if( r != ASVGF_SUCCESS ) return r;
    DeviceGuard        guard( d->device );
    const cudaStream_t s = ( cudaStream_t )stream;
    const Uniform      u = makeUniform( *d, *pUniform );
    // this pass reads what the previous frame's temporal pass wrote (IndirAccum_Prev) and overwrites what it read
    // (DISPingGradient): when the two are enqueued on different streams the order must be made explicit
    if( d->temporalRecorded ) CUDA_TRY( cudaStreamWaitEvent( s, d->evTemporalDone, 0 ) );
    d->guideValid = d->centreValid = false; // a new frame starts
    const PdlScope pdl( *d );
    r = runPass( *d, s, frameIndex, u, ASVGF_PASS_GRADIENT_SAMPLES );
    if( r != ASVGF_SUCCESS ) return r;
    d->gradientAtrousDone = false;
    if( ( d->info.flags & ASVGF_CREATE_EAGER_GRADIENT_ATROUS ) && d->info.bandRowEnd == 0 )
    {
        r = runPass( *d, s, frameIndex, u, ASVGF_PASS_GRADIENT_ATROUS_ALL ); // Denoiser.cpp:87-120, issued early
        if( r != ASVGF_SUCCESS ) return r;
        d->gradientAtrousDone = true;
    }
    CUDA_TRY( cudaEventRecord( d->evGradientsDone, s ) );
    d->gradientsRecorded = true;
    return ASVGF_SUCCESS;
}

AsvgfResult asvgfDenoise( AsvgfDenoiser d, AsvgfStream stream, uint32_t frameIndex, const AsvgfGlobalUniform* pUniform )
{
    AsvgfResult r = checkCall( d, frameIndex, pUniform );
    if( r != ASVGF_SUCCESS ) return r;
    DeviceGuard        guard( d->device );
    const cudaStream_t s = ( cudaStream_t )stream;
    const Uniform      u = makeUniform( *d, *pUniform );
    d->guideValid = d->centreValid = false;
    const PdlScope pdl( *d );
    if( d->info.bandRowEnd != 0 && !( d->info.flags & ASVGF_CREATE_BAND_APRON ) )
    {
        r = runGuide( *d, s, frameIndex ); // per-pass band schedule: the temporal pass covers the owned rows only
        if( r != ASVGF_SUCCESS ) return r;
    }
    if( d->gradientsRecorded ) CUDA_TRY( cudaStreamWaitEvent( s, d->evGradientsDone, 0 ) );
    d->pdlArmed = false;
    if( !d->gradientAtrousDone )
    {
        r = runPass( *d, s, frameIndex, u, ASVGF_PASS_GRADIENT_ATROUS_ALL ); // Denoiser.cpp:87-120
        if( r != ASVGF_SUCCESS ) return r;
    }
    d->gradientAtrousDone = false;
    for( uint32_t pass = ASVGF_PASS_TEMPORAL; pass < ASVGF_PASS_ATROUS_0 + d->iterations; pass++ ) // Denoiser.cpp:123-254
    {
        r = runPass( *d, s, frameIndex, u, pass );
        if( r != ASVGF_SUCCESS ) return r;
        if( pass == ASVGF_PASS_TEMPORAL )
        {
            CUDA_TRY( cudaEventRecord( d->evTemporalDone, s ) );
            d->temporalRecorded = true;
            d->pdlArmed = false; // the anti-firefly pass sits behind the event record: launched normally
        }
    }
    return ASVGF_SUCCESS;
}

AsvgfResult asvgfDenoisePass( AsvgfDenoiser d, AsvgfStream stream, uint32_t frameIndex, const AsvgfGlobalUniform* pUniform, uint32_t pass )
{
    AsvgfResult r = checkCall( d, frameIndex, pUniform );
    if( r != ASVGF_SUCCESS ) return r;
    DeviceGuard guard( d->device );
    const cudaStream_t s = ( cudaStream_t )stream;
    if( pass == ASVGF_PASS_GUIDE )
    {
        r = runGuide( *d, s, frameIndex );
        d->guideValid = r == ASVGF_SUCCESS;
        return r;
    }
    const bool gradAtrous = ( pass >= ASVGF_PASS_GRADIENT_ATROUS_0 && pass <= ASVGF_PASS_GRADIENT_ATROUS_3 ) || pass == ASVGF_PASS_GRADIENT_ATROUS_ALL;
    if( gradAtrous && d->gradientAtrousDone ) return ASVGF_SUCCESS; // ASVGF_CREATE_EAGER_GRADIENT_ATROUS: already ran behind the gradient samples
    if( pass >= ASVGF_PASS_ATROUS_0 && pass < ASVGF_PASS_ATROUS_0 + d->iterations )
    {
        // the guide plane is refreshed by the first a-trous dispatch after anything that may have changed the geometry images
        if( !d->guideValid )
        {
            r = runGuide( *d, s, frameIndex );
            if( r != ASVGF_SUCCESS ) return r;
            d->guideValid = true;
        }
    }
    else if( pass == ASVGF_PASS_TEMPORAL )
    {
        d->guideValid = d->centreValid = false; // (the pass re-validates the guide when it covers the rows, runTemporal)
        d->gradientAtrousDone = false;
        if( d->gradientsRecorded ) CUDA_TRY( cudaStreamWaitEvent( s, d->evGradientsDone, 0 ) );
    }
    else if( pass < ASVGF_PASS_TEMPORAL )
    {
        d->guideValid = d->centreValid = false; // a new frame starts
    }
    r = runPass( *d, s, frameIndex, makeUniform( *d, *pUniform ), pass );
    if( r == ASVGF_SUCCESS && pass == ASVGF_PASS_TEMPORAL )
    {
        CUDA_TRY( cudaEventRecord( d->evTemporalDone, s ) );
        d->temporalRecorded = true;
    }
    return r;
}

AsvgfResult asvgfResolveCheckerboard( AsvgfDenoiser d, AsvgfStream stream, uint32_t frameIndex, const AsvgfGlobalUniform* pUniform )
{
    AsvgfResult r = checkCall( d, frameIndex, pUniform );
    if( r != ASVGF_SUCCESS ) return r;
    if( d->info.bandRowEnd != 0 ) return fail( ASVGF_UNSUPPORTED, "asvgfResolveCheckerboard on a band handle: the resolve reads PreFinal one row beyond the band, which no pass writes and no exchange fills (gather the frame on one device first)" );
    DeviceGuard guard( d->device );
    return runResolveCheckerboard( *d, ( cudaStream_t )stream, frameIndex, makeUniform( *d, *pUniform ) );
}

AsvgfResult asvgfPrepareFinal( AsvgfDenoiser d, AsvgfStream stream, uint32_t frameIndex, const AsvgfGlobalUniform* pUniform, const uint32_t* pLpmControl )
{
    AsvgfResult r = checkCall( d, frameIndex, pUniform );
    if( r != ASVGF_SUCCESS ) return r;
    if( !pLpmControl ) return fail( ASVGF_WRONG_ARGUMENT, "null LPM control block" );
    if( d->info.bandRowEnd != 0 ) return fail( ASVGF_UNSUPPORTED, "asvgfPrepareFinal on a band handle: Final is the resolved frame, which band handles do not produce (asvgfResolveCheckerboard)" );
    if( pUniform->debugShowFlags & ( 1u << 8 ) ) return fail( ASVGF_UNSUPPORTED, "DEBUG_SHOW_FLAG_LIGHT_GRID needs the ray tracer's light grid (LightGrid.h): not an image-space input of this library" );
    DeviceGuard guard( d->device );
    return runPrepareFinal( *d, ( cudaStream_t )stream, frameIndex, makeUniform( *d, *pUniform ), *pUniform, pLpmControl );
}

AsvgfResult asvgfScatterAccumulate( AsvgfDenoiser d, AsvgfStream stream, uint32_t frameIndex, const AsvgfGlobalUniform* pUniform, float maxHistoryLength )
{
    AsvgfResult r = checkCall( d, frameIndex, pUniform );
    if( r != ASVGF_SUCCESS ) return r;
    if( pUniform->volumeEnableType > ASVGF_VOLUME_ENABLE_VOLUMETRIC ) return fail( ASVGF_WRONG_ARGUMENT, "unknown volumeEnableType" );
    DeviceGuard guard( d->device );
    return runScatterAccumulate( *d, ( cudaStream_t )stream, frameIndex, makeUniform( *d, *pUniform ), *pUniform, maxHistoryLength );
}

AsvgfResult asvgfLuminanceHistogram( AsvgfDenoiser d, AsvgfStream stream, uint32_t frameIndex, const AsvgfGlobalUniform* pUniform )
{
    AsvgfResult r = checkCall( d, frameIndex, pUniform );
    if( r != ASVGF_SUCCESS ) return r;
    if( d->info.bandRowEnd != 0 ) return fail( ASVGF_UNSUPPORTED, "asvgfLuminanceHistogram on a band handle: a band would histogram its own rows only; the average needs the whole frame's histogram (gather the frame on one device first)" );
    DeviceGuard guard( d->device );
    return runLuminanceHistogram( *d, ( cudaStream_t )stream, frameIndex, makeUniform( *d, *pUniform ), *pUniform );
}

AsvgfResult asvgfLuminanceAverage( AsvgfDenoiser d, AsvgfStream stream, uint32_t frameIndex, const AsvgfGlobalUniform* pUniform )
{
    AsvgfResult r = checkCall( d, frameIndex, pUniform );
    if( r != ASVGF_SUCCESS ) return r;
    if( d->info.bandRowEnd != 0 ) return fail( ASVGF_UNSUPPORTED, "asvgfLuminanceAverage on a band handle: a band holds a partial histogram; the average needs the whole frame's (gather the frame on one device first)" );
    DeviceGuard guard( d->device );
    return runLuminanceAverage( *d, ( cudaStream_t )stream, frameIndex, *pUniform );
}

AsvgfResult asvgfCalculateExposure( AsvgfDenoiser d, AsvgfStream stream, uint32_t frameIndex, const AsvgfGlobalUniform* pUniform )
{
    AsvgfResult r = asvgfLuminanceHistogram( d, stream, frameIndex, pUniform ); // Tonemapping.cpp:59-115
    if( r != ASVGF_SUCCESS ) return r;
    return asvgfLuminanceAverage( d, stream, frameIndex, pUniform );           // Tonemapping.cpp:117-175
}

AsvgfResult asvgfGetTonemappingBuffer( AsvgfDenoiser d, void** ppDevice )
{
    if( !d || !ppDevice ) return fail( ASVGF_WRONG_ARGUMENT, "null argument" );
    *ppDevice = d->tm;
    return ASVGF_SUCCESS;
}
AsvgfResult asvgfUploadTonemapping( AsvgfDenoiser d, AsvgfStream stream, const AsvgfTonemapping* pHost )
{
    if( !d || !pHost ) return fail( ASVGF_WRONG_ARGUMENT, "null argument" );
    DeviceGuard guard( d->device );
    CUDA_TRY( cudaMemcpyAsync( d->tm, pHost, sizeof( AsvgfTonemapping ), cudaMemcpyHostToDevice, ( cudaStream_t )stream ) );
    return ASVGF_SUCCESS;
}
AsvgfResult asvgfDownloadTonemapping( AsvgfDenoiser d, AsvgfStream stream, AsvgfTonemapping* pHost )
{
    if( !d || !pHost ) return fail( ASVGF_WRONG_ARGUMENT, "null argument" );
    DeviceGuard guard( d->device );
    CUDA_TRY( cudaMemcpyAsync( pHost, d->tm, sizeof( AsvgfTonemapping ), cudaMemcpyDeviceToHost, ( cudaStream_t )stream ) );
    return ASVGF_SUCCESS;
}

AsvgfResult asvgfSynchronize( AsvgfDenoiser d, AsvgfStream stream )
{
    if( !d ) return fail( ASVGF_WRONG_ARGUMENT, "null denoiser" );
    DeviceGuard guard( d->device );
    CUDA_TRY( cudaStreamSynchronize( ( cudaStream_t )stream ) );
    if( d->haloMiss )
    {
        uint32_t miss = 0;
        CUDA_TRY( cudaMemcpy( &miss, d->haloMiss, sizeof( miss ), cudaMemcpyDeviceToHost ) );
        if( miss )
        {
            CUDA_TRY( cudaMemset( d->haloMiss, 0, sizeof( miss ) ) );
            return fail( ASVGF_HALO_EXCEEDED, "a motion vector reached beyond the band's halo rows: the banded frame differs from the single-GPU frame (recreate with a larger bandHaloRows)" );
        }
    }
    return ASVGF_SUCCESS;
}

uint64_t asvgfGetLaunchCount( AsvgfDenoiser d ) { return d ? d->launchCount : 0; }

const char* asvgfPassName( uint32_t pass )
{
    if( pass == ASVGF_PASS_GRADIENT_SAMPLES ) return "k_gradient_samples";
    if( pass >= ASVGF_PASS_GRADIENT_ATROUS_0 && pass <= ASVGF_PASS_GRADIENT_ATROUS_3 ) return "k_gradient_atrous";
    if( pass == ASVGF_PASS_GRADIENT_ATROUS_ALL ) return "k_gradient_atrous";
    if( pass == ASVGF_PASS_GUIDE ) return "k_guide";
    if( pass == ASVGF_PASS_TEMPORAL ) return "k_temporal";
    if( pass == ASVGF_PASS_ANTIFIREFLY ) return "k_antifirefly";
    if( pass == ASVGF_PASS_VARIANCE ) return "k_variance";
    if( pass >= ASVGF_PASS_ATROUS_0 && pass < ASVGF_PASS_COUNT ) return "k_atrous_half";
    return "?";
}

// kernel the default configuration launches for a pass of an N-iteration chain
const char* asvgfPassKernelName( uint32_t pass, uint32_t atrousIterations )
{
    ( void )atrousIterations; // every a-trous iteration runs k_atrous_half
    return asvgfPassName( pass );
}

// Algorithmic HBM bytes per full-resolution pixel: unique reads + writes in the reference's storage formats
// (BASELINE.md section 3).
double asvgfPassBytesPerPixel( uint32_t pass, uint32_t atrousIterations )
{
    if( pass == ASVGF_PASS_GRADIENT_SAMPLES ) return 12.8;
    if( pass >= ASVGF_PASS_GRADIENT_ATROUS_0 && pass <= ASVGF_PASS_GRADIENT_ATROUS_3 ) return 1.0;
    if( pass == ASVGF_PASS_GRADIENT_ATROUS_ALL ) return 4.0; // the reference's four dispatches (36 B per gradient pixel)
    if( pass == ASVGF_PASS_GUIDE ) return 0.0; // not a reference dispatch: its 24 B/px are extra traffic this design adds, not algorithmic bytes
    if( pass == ASVGF_PASS_TEMPORAL ) return 126.4;
    if( pass == ASVGF_PASS_ANTIFIREFLY ) return 24.0;
    if( pass == ASVGF_PASS_VARIANCE ) return 30.0;
    if( pass >= ASVGF_PASS_ATROUS_0 && pass < ASVGF_PASS_ATROUS_0 + atrousIterations ) return pass + 1 == ASVGF_PASS_ATROUS_0 + atrousIterations ? 55.0 : 52.0;
    return 0.0;
}

int32_t asvgfBandApronRows( uint32_t atrousIterations, uint32_t pass )
{
    if( atrousIterations < 2 || atrousIterations > ASVGF_MAX_ATROUS_ITERATIONS ) return -1;
    return apronRows( atrousIterations, pass );
}

const char* asvgfResultString( AsvgfResult r )
{
    switch( r )
    {
        case ASVGF_SUCCESS: return "ASVGF_SUCCESS";
        case ASVGF_WRONG_ARGUMENT: return "ASVGF_WRONG_ARGUMENT";
        case ASVGF_CUDA_ERROR: return "ASVGF_CUDA_ERROR";
        case ASVGF_OUT_OF_MEMORY: return "ASVGF_OUT_OF_MEMORY";
        case ASVGF_UNSUPPORTED: return "ASVGF_UNSUPPORTED";
        case ASVGF_NOT_BOUND: return "ASVGF_NOT_BOUND";
        case ASVGF_HALO_EXCEEDED: return "ASVGF_HALO_EXCEEDED";
        default: return "ASVGF_?";
    }
}
const char* asvgfLastErrorMessage( void ) { return g_lastError.c_str(); }

} // extern "C"

#include "asvgf_bandset.inl"
#include "asvgf_interop.inl"
