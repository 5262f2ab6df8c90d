/*
 * asvgf.h -- C ABI of the B200-native A-SVGF denoiser (libasvgf.so).
 *
 * Drop-in boundary: the pass sequence that RTGL1 records in
 *     RTGL1::Denoiser::Denoise( VkCommandBuffer, uint32_t frameIndex, shared_ptr<const GlobalUniform> )
 *                                                           (Source/Denoiser.h:45-47, Source/Denoiser.cpp:64-255)
 * plus the image-space half of
 *     RTGL1::PathTracer::CalculateGradientsSamples( const TraceParams& )            (Source/PathTracer.cpp:196-219)
 *
 * Mapping of the reference's interface onto this one
 *     VkCommandBuffer cmd            -> a CUDA stream; calls only ENQUEUE work, like vkCmd* recording
 *     descriptor set 0 (Framebuffers)-> the image table below, same indices (asvgf_layout.h), same cur/prev
 *                                        swap by frameIndex (Source/Framebuffers.cpp:35-50)
 *     descriptor set 1 (UBO)         -> const AsvgfGlobalUniform*, byte-compatible with ShGlobalUniform
 *     VK_CHECKERROR / RgException    -> AsvgfResult return codes, never throws (Source/RTGL1.cpp:150-168)
 *
 * Images are linear, pitched device memory (row pitch a multiple of 256 B) in the reference's storage formats.
 * A handle is not thread-safe; distinct handles are independent (one handle per view stream).
 * There is NO CPU fallback: every entry point that computes fails with ASVGF_CUDA_ERROR when no device is usable.
 */
#ifndef ASVGF_H_
#define ASVGF_H_

#include "asvgf_layout.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined( _WIN32 )
    #define ASVGF_API __declspec( dllexport )
#else
    #define ASVGF_API __attribute__( ( visibility( "default" ) ) )
#endif

typedef struct AsvgfDenoiser_T* AsvgfDenoiser;
typedef void*                   AsvgfStream; /* a cudaStream_t; NULL = the legacy default stream */

/* RgResult-style (Include/RTGL1/RTGL1.h:79-94) */
typedef enum AsvgfResult
{
    ASVGF_SUCCESS         = 0,
    ASVGF_WRONG_ARGUMENT  = 1,
    ASVGF_CUDA_ERROR      = 2,
    ASVGF_OUT_OF_MEMORY   = 3,
    ASVGF_UNSUPPORTED     = 4,
    ASVGF_NOT_BOUND       = 5, /* a pass needs an image that is neither owned nor bound */
    ASVGF_HALO_EXCEEDED   = 6  /* band handles: a reprojected history tap (temporal pass, gradient samples) lay inside the frame
                                  but outside the rows the band stores -- motion reached further than bandHaloRows allows, so the
                                  banded frame is NOT the single-GPU frame.  Reported by asvgfSynchronize / asvgfBandSetSynchronize;
                                  the motion budget of a halo is bandHaloRows - asvgfBandApronRows(N, ASVGF_PASS_TEMPORAL) rows */
} AsvgfResult;

typedef enum AsvgfCreateFlagBits
{
    ASVGF_CREATE_EXPORT_DEBUG_MASK = 1u << 0, /* temporal pass also writes ASVGF_IMAGE_DEBUG_MASK               */
    ASVGF_CREATE_EXTERNAL_IMAGES   = 1u << 1, /* allocate nothing: every image comes from asvgfBindExternalImage */
    ASVGF_CREATE_BAND_APRON        = 1u << 3, /* band handles: every pass also computes the rows beyond the band that the
                                                 later passes of the frame will read (asvgfBandApronRows), so one frame
                                                 needs two halo exchanges instead of one before each pass            */
    ASVGF_CREATE_EAGER_GRADIENT_ATROUS = 1u << 4, /* asvgfGradientSamples also enqueues, on ITS stream, the four gradient a-trous
                                                 dispatches that open the following asvgfDenoise (Denoiser.cpp:87-120), which
                                                 then skips them: a caller that gives the two calls different streams gets the
                                                 latency-bound gradient chain of frame t+1 overlapped with the a-trous
                                                 iterations of frame t.  The library orders the two streams itself (events);
                                                 results are identical.  Ignored on band handles.                     */
    ASVGF_CREATE_LITERAL_KERNELS   = 1u << 2  /* experiments/ build only (make LITERAL=1): one-thread-per-pixel kernels that follow
                                                 the shader text; ASVGF_UNSUPPORTED in the release build and on band handles */
} AsvgfCreateFlagBits;

/* = Denoiser ctor + Framebuffers::PrepareForSize (Source/Denoiser.cpp:24-55, Source/Framebuffers.cpp:216-230) */
typedef struct AsvgfCreateInfo
{
    uint32_t renderWidth;      /* full frame, must be even (Source/VulkanDevice.cpp:186-187)                     */
    uint32_t renderHeight;
    uint32_t atrousIterations; /* 4 = the reference; 2..8 accepted (5 = BASELINE config 3)                       */
    uint32_t flags;            /* AsvgfCreateFlagBits                                                            */
    int32_t  device;           /* CUDA ordinal, -1 = the calling thread's current device                        */
    /* Horizontal band owned by this handle, for one-process-per-GPU frame splitting.  bandRowEnd == 0 means the
       whole frame.  Rows are full-resolution rows and must be multiples of 3 (strata do not straddle bands).
       The handle stores rows [bandRowBegin-bandHaloRows, bandRowEnd+bandHaloRows) clipped to the frame; a pass
       COMPUTES only the owned rows and READS halo rows, which the host fills (NCCL/P2P) before the pass. */
    uint32_t bandRowBegin;
    uint32_t bandRowEnd;
    uint32_t bandHaloRows;     /* multiple of 3; >= 2^(atrousIterations-1), and with ASVGF_CREATE_BAND_APRON also
                                  >= max(apron(TEMPORAL), 3*apron(GRADIENT_SAMPLES)) (asvgfBandApronRows); checked     */
} AsvgfCreateInfo;

/* Passes in the reference's dispatch order (Source/Denoiser.cpp:87-254, Source/PathTracer.cpp:212-218). */
typedef enum AsvgfPass
{
    ASVGF_PASS_GRADIENT_SAMPLES  = 0,  /* RtGradients.rgen (image-space part)     */
    ASVGF_PASS_GRADIENT_ATROUS_0 = 1,  /* CmASVGFGradientAtrous.comp, iteration 0 */
    ASVGF_PASS_GRADIENT_ATROUS_1 = 2,
    ASVGF_PASS_GRADIENT_ATROUS_2 = 3,
    ASVGF_PASS_GRADIENT_ATROUS_3 = 4,
    ASVGF_PASS_TEMPORAL          = 5,  /* CmSVGFTemporalAccumulation.comp         */
    ASVGF_PASS_ANTIFIREFLY       = 6,  /* CmAntiFirefly.comp                      */
    ASVGF_PASS_VARIANCE          = 7,  /* CmSVGFEstimateVariance.comp             */
    ASVGF_PASS_ATROUS_0          = 8,  /* CmSVGFAtrous_Iter0.comp; ASVGF_PASS_ATROUS_0+i = CmSVGFAtrous.comp, i  */
    ASVGF_PASS_COUNT             = 8 + ASVGF_MAX_ATROUS_ITERATIONS,
    /* the four gradient a-trous dispatches in one call, as asvgfDenoise issues them (Denoiser.cpp:87-120) */
    ASVGF_PASS_GRADIENT_ATROUS_ALL = 100,
    /* not a reference dispatch: decodes the frame's normal / depth / roughness once into the plane the a-trous iterations
       read (asvgfDenoise issues it first; asvgfDenoisePass issues it on demand before an a-trous pass) */
    ASVGF_PASS_GUIDE = 101
} AsvgfPass;

ASVGF_API AsvgfResult asvgfCreate( const AsvgfCreateInfo* pInfo, AsvgfDenoiser* pDenoiser );
ASVGF_API AsvgfResult asvgfDestroy( AsvgfDenoiser denoiser );

/* Size/format of an image of the table for a given frame size: (W,H) or ((W+1)/3,(H+1)/3) (Framebuffers.cpp:639-700).
   Returns ASVGF_WRONG_ARGUMENT for indices the denoiser path does not use. */
ASVGF_API AsvgfResult asvgfGetImageInfo( uint32_t imageIndex, uint32_t renderWidth, uint32_t renderHeight,
                                         uint32_t* pFormat, uint32_t* pWidth, uint32_t* pHeight,
                                         uint32_t* pBytesPerPixel );

/* = Framebuffers::GetImage(fbImageIndex, frameIndex): applies the cur/prev swap.  *ppDevice addresses the first
   STORED row, whose frame row number is *pFirstRow (0 unless banded); *pRows rows are stored. */
ASVGF_API AsvgfResult asvgfGetImage( AsvgfDenoiser denoiser, uint32_t imageIndex, uint32_t frameIndex,
                                     void** ppDevice, size_t* pPitchBytes, uint32_t* pWidth, uint32_t* pFirstRow,
                                     uint32_t* pRows );

/* Caller-owned storage for one PHYSICAL image (no swap applied: bind X and X_PREV separately).  pDevice addresses
   the first stored row (same row range as asvgfGetImage reports); pDevice and pitchBytes must be multiples of 16 bytes
   (ASVGF_WRONG_ARGUMENT otherwise). */
ASVGF_API AsvgfResult asvgfBindExternalImage( AsvgfDenoiser denoiser, uint32_t imageIndex, void* pDevice,
                                              size_t pitchBytes );

/* Host <-> device copies of the stored rows of one image (swap applied), enqueued on `stream`.  The host buffer
   holds the stored rows back to back with hostPitchBytes per row (0 = tightly packed). */
ASVGF_API AsvgfResult asvgfUploadImage( AsvgfDenoiser denoiser, AsvgfStream stream, uint32_t imageIndex,
                                        uint32_t frameIndex, const void* pHost, size_t hostPitchBytes );
ASVGF_API AsvgfResult asvgfDownloadImage( AsvgfDenoiser denoiser, AsvgfStream stream, uint32_t imageIndex,
                                          uint32_t frameIndex, void* pHost, size_t hostPitchBytes );
/* Zero every owned image (= fresh images after a resolution change, Framebuffers.cpp:216-230). */
ASVGF_API AsvgfResult asvgfResetHistory( AsvgfDenoiser denoiser, AsvgfStream stream );
/* Checkpoint / resume of the temporal state.  The reference keeps it in GPU images only and loses it on a resolution change
   (Framebuffers.cpp:216-230); here it can be taken out and put back: both members of every cur/prev pair (accumulated colours,
   moments, history lengths, the previous frame's depth / normals / roughness / surface positions / gradient inputs, scattering),
   the a-trous colour history, the gradient history, the other images the passes write (some are only partly overwritten per frame)
   and the exposure state (AsvgfTonemapping) -- everything but the producer's per-frame inputs -- as tightly packed rows behind a
   small header.  Copies are enqueued on `stream` (pinned host memory makes them asynchronous); a frame computed after asvgfLoadState is
   bit-identical to the frame the saving handle would have computed.  Loading into a handle of another size, band or image set is
   ASVGF_WRONG_ARGUMENT. */
ASVGF_API AsvgfResult asvgfGetStateSize( AsvgfDenoiser denoiser, size_t* pBytes );
ASVGF_API AsvgfResult asvgfSaveState( AsvgfDenoiser denoiser, AsvgfStream stream, void* pHost, size_t bytes );
ASVGF_API AsvgfResult asvgfLoadState( AsvgfDenoiser denoiser, AsvgfStream stream, const void* pHost, size_t bytes );

/* = PathTracer::CalculateGradientsSamples, image-space part (RtGradients.rgen:44-105,139-181); the ray-traced
   re-shade is read from ASVGF_IMAGE_FWD_LUM.  Enqueues only.  `stream` may differ from the stream asvgfDenoise is given:
   the library makes this pass wait for the previous frame's temporal pass and the next temporal pass wait for it. */
ASVGF_API AsvgfResult asvgfGradientSamples( AsvgfDenoiser denoiser, AsvgfStream stream, uint32_t frameIndex,
                                            const AsvgfGlobalUniform* pUniform );
/* = Denoiser::Denoise.  Enqueues the whole chain (gradient a-trous x4, temporal, anti-firefly, variance,
   a-trous x N with the composition in the last iteration) and returns immediately. */
ASVGF_API AsvgfResult asvgfDenoise( AsvgfDenoiser denoiser, AsvgfStream stream, uint32_t frameIndex,
                                    const AsvgfGlobalUniform* pUniform );
/* One reference dispatch at a time (per-pass parity tests; banded multi-GPU runs interleave halo exchanges). */
ASVGF_API AsvgfResult asvgfDenoisePass( AsvgfDenoiser denoiser, AsvgfStream stream, uint32_t frameIndex,
                                        const AsvgfGlobalUniform* pUniform, uint32_t pass );

/* ---- the two dispatches that consume the denoiser's outputs right after it ---- */

/* = ImageComposition::PrepareForRaster -> ProcessCheckerboard (Source/ImageComposition.cpp:87-92, 157-222), shader
   CmCheckerboard.comp:43-78: ASVGF_IMAGE_PRE_FINAL -> ASVGF_IMAGE_FINAL with the cross-shaped resolve where
   ASVGF_IMAGE_THROUGHPUT.a marks a split (reflect/refract) pixel, and the same for the shader's two other image pairs,
   ASVGF_IMAGE_SCREEN_EMIS_RT -> ASVGF_IMAGE_SCREEN_EMISSION and ASVGF_IMAGE_ACID_FOG_RT -> ASVGF_IMAGE_ACID_FOG (written by the
   rasteriser / fog producers; skipped when those four images have no storage).  Enqueues only. */
ASVGF_API AsvgfResult asvgfResolveCheckerboard( AsvgfDenoiser denoiser, AsvgfStream stream, uint32_t frameIndex,
                                                const AsvgfGlobalUniform* pUniform );

/* = Tonemapping::CalculateExposure (Source/Tonemapping.cpp:59-175): the 256-bin log-luminance histogram of
   ASVGF_IMAGE_HISTOGRAM_INPUT (CmLuminanceHistogram.comp:41-86) followed by the eye-adaptation update of
   AsvgfTonemapping.avgLuminance, which also clears the histogram for the next frame (CmLuminanceAvg.comp:37-89).
   Enqueues only.  Whole-frame handles only: on a band handle this call, asvgfLuminanceHistogram / Average and
   asvgfResolveCheckerboard return ASVGF_UNSUPPORTED (a band holds part of the histogram, and the resolve reads one row of
   PreFinal beyond the band). */
ASVGF_API AsvgfResult asvgfCalculateExposure( AsvgfDenoiser denoiser, AsvgfStream stream, uint32_t frameIndex,
                                              const AsvgfGlobalUniform* pUniform );
/* The two dispatches of asvgfCalculateExposure separately (parity tests look at the histogram in between). */
ASVGF_API AsvgfResult asvgfLuminanceHistogram( AsvgfDenoiser denoiser, AsvgfStream stream, uint32_t frameIndex,
                                               const AsvgfGlobalUniform* pUniform );
ASVGF_API AsvgfResult asvgfLuminanceAverage( AsvgfDenoiser denoiser, AsvgfStream stream, uint32_t frameIndex,
                                             const AsvgfGlobalUniform* pUniform );
/* = ImageComposition::Finalize -> ApplyTonemapping (Source/ImageComposition.cpp:94-155, CmPrepareFinal.comp:144-172): the dispatch
   that turns the resolved frame into what the raster / bloom / upscale chain consumes.  In place on ASVGF_IMAGE_FINAL:
   Final * transmittance + in-scattering (ASVGF_IMAGE_SCATTERING), exposure from the eye-adapted average luminance this library's
   asvgfCalculateExposure keeps (Exposure.h:36-58), AMD LPM tonemap (Uncharted 2 on the classic-shading columns), + ASVGF_IMAGE_ACID_FOG,
   the debug views of debugShowFlags; writes ASVGF_IMAGE_BLOOM_INPUT.  pLpmControl = the 24 x uvec4 (96 x uint32) control block
   ImageComposition::SetupLpmParams computes on the host with AMD's LpmSetup (ImageComposition.cpp:365-420); read during the call.
   Call order of the reference frame: asvgfDenoise, asvgfScatterAccumulate, asvgfCalculateExposure, asvgfResolveCheckerboard, ...,
   asvgfPrepareFinal (VulkanDevice.cpp:583-590, 608-612).  ASVGF_UNSUPPORTED on band handles and for DEBUG_SHOW_FLAG_LIGHT_GRID. */
ASVGF_API AsvgfResult asvgfPrepareFinal( AsvgfDenoiser denoiser, AsvgfStream stream, uint32_t frameIndex,
                                         const AsvgfGlobalUniform* pUniform, const uint32_t* pLpmControl );

/* = Volumetric::ProcessScattering(cmd, frameIndex, uniform, ..., maxHistoryLength) (Source/Volumetric.cpp:71-195), shader
   CmScatterAccum.comp:96-163: the frame's other temporal-reprojection pass.  Reprojects
   ASVGF_IMAGE_SCATTERING_PREV / _HISTORY_PREV with the temporal pass's disocclusion tests and blends in this frame's
   scattering sample: computed in place for volumeEnableType NONE and SIMPLE (:75-93), read from
   ASVGF_IMAGE_SCATTER_CURRENT for VOLUMETRIC (the 3-D illumination-volume lookup of :41-73 is not image-space).
   maxHistoryLength is the shader's push constant.  Enqueues only. */
ASVGF_API AsvgfResult asvgfScatterAccumulate( AsvgfDenoiser denoiser, AsvgfStream stream, uint32_t frameIndex,
                                              const AsvgfGlobalUniform* pUniform, float maxHistoryLength );

/* Device address of the handle's AsvgfTonemapping (= the reference's tonemapping buffer, Tonemapping.cpp:183-191; zero-initialised here),
   and host copies of it enqueued on `stream`. */
ASVGF_API AsvgfResult asvgfGetTonemappingBuffer( AsvgfDenoiser denoiser, void** ppDevice );
ASVGF_API AsvgfResult asvgfUploadTonemapping( AsvgfDenoiser denoiser, AsvgfStream stream, const AsvgfTonemapping* pHost );
ASVGF_API AsvgfResult asvgfDownloadTonemapping( AsvgfDenoiser denoiser, AsvgfStream stream, AsvgfTonemapping* pHost );

/* Blocks the host until everything enqueued on `stream` has finished (= waiting for the frame fence,
   Source/VulkanDevice.cpp:791-796).  Surfaces asynchronous kernel faults as ASVGF_CUDA_ERROR. */
ASVGF_API AsvgfResult asvgfSynchronize( AsvgfDenoiser denoiser, AsvgfStream stream );

/* ---- Vulkan interop, CUDA half ----
   RTGL1 keeps the G-buffer in VkImages (Source/Framebuffers.cpp:702-804) and orders the denoiser with its command buffer and
   frame fence (Source/VulkanDevice.cpp:551-584, 791-796).  A host that keeps that producer exports each image's memory
   (VK_KHR_external_memory_fd, OPAQUE_FD; linear tiling, or a linear VkBuffer filled by vkCmdCopyImageToBuffer) and a semaphore per
   direction (VK_KHR_external_semaphore_fd); these calls import the file descriptors, map the memory as linear device memory for
   asvgfBindExternalImage, and wait / signal on a CUDA stream.  No Vulkan header is involved on this side.  On success the
   descriptor belongs to CUDA (do not close it).  device = -1: the calling thread's current device. */
typedef struct AsvgfExternalMemory_T*    AsvgfExternalMemory;
typedef struct AsvgfExternalSemaphore_T* AsvgfExternalSemaphore;
ASVGF_API AsvgfResult asvgfImportExternalMemoryFd( int32_t device, int fd, uint64_t allocationSize, uint32_t dedicated,
                                                   AsvgfExternalMemory* pMemory );
ASVGF_API AsvgfResult asvgfExternalMemoryGetBuffer( AsvgfExternalMemory memory, uint64_t offset, uint64_t size, void** ppDevice );
ASVGF_API AsvgfResult asvgfDestroyExternalMemory( AsvgfExternalMemory memory );
ASVGF_API AsvgfResult asvgfImportExternalSemaphoreFd( int32_t device, int fd, uint32_t timeline, AsvgfExternalSemaphore* pSemaphore );
ASVGF_API AsvgfResult asvgfWaitExternalSemaphore( AsvgfExternalSemaphore semaphore, AsvgfStream stream, uint64_t value );
ASVGF_API AsvgfResult asvgfSignalExternalSemaphore( AsvgfExternalSemaphore semaphore, AsvgfStream stream, uint64_t value );
ASVGF_API AsvgfResult asvgfDestroyExternalSemaphore( AsvgfExternalSemaphore semaphore );

/* ---- one frame across several GPUs (BASELINE config 4; the reference has no multi-GPU path) ----
   A band set owns one band handle per entry of pDevices -- a device may be listed more than once (several bands on one
   GPU) -- created with ASVGF_CREATE_BAND_APRON, a stream per band, and the halo exchange between neighbouring bands
   (peer copies over NVLink, issued behind the producing pass so that they overlap the rest of the frame).  The result is
   bit-identical to the single-GPU frame as long as per-frame vertical motion stays within the halo
   (ASVGF_HALO_EXCEEDED otherwise).  The caller fills each band's producer images through the band's handle
   (asvgfGetImage reports the rows the band stores: its rows plus the halo) on the band's stream. */
typedef struct AsvgfBandSet_T* AsvgfBandSet;
typedef struct AsvgfBandSetCreateInfo
{
    uint32_t        renderWidth, renderHeight, atrousIterations;
    uint32_t        flags;              /* ASVGF_CREATE_EXPORT_DEBUG_MASK, ASVGF_CREATE_EXTERNAL_IMAGES                    */
    uint32_t        deviceCount;        /* number of bands                                                               */
    const int32_t*  pDevices;           /* CUDA ordinal of every band, top to bottom                                     */
    const uint32_t* pBandRowBoundaries; /* deviceCount + 1 rows: 0 ... renderHeight, multiples of 3; NULL = equal bands  */
    uint32_t        bandHaloRows;       /* 0 = apron rows + 48 rows of motion; bands must be at least this tall          */
} AsvgfBandSetCreateInfo;
ASVGF_API AsvgfResult asvgfBandSetCreate( const AsvgfBandSetCreateInfo* pInfo, AsvgfBandSet* pSet );
ASVGF_API AsvgfResult asvgfBandSetDestroy( AsvgfBandSet set );
ASVGF_API uint32_t    asvgfBandSetGetBandCount( AsvgfBandSet set );
ASVGF_API AsvgfResult asvgfBandSetGetBand( AsvgfBandSet set, uint32_t band, AsvgfDenoiser* pDenoiser, AsvgfStream* pStream,
                                           uint32_t* pRowBegin, uint32_t* pRowEnd );
/* = CalculateGradientsSamples + Denoiser::Denoise of the whole frame.  Enqueues only. */
ASVGF_API AsvgfResult asvgfBandSetFrame( AsvgfBandSet set, uint32_t frameIndex, const AsvgfGlobalUniform* pUniform );
ASVGF_API AsvgfResult asvgfBandSetSynchronize( AsvgfBandSet set );
/* device-side timing of the frames enqueued between the two calls: the longest band interval (and each band's) in ms */
ASVGF_API AsvgfResult asvgfBandSetBeginTiming( AsvgfBandSet set );
ASVGF_API AsvgfResult asvgfBandSetEndTiming( AsvgfBandSet set, float* pMaxMilliseconds, float* pPerBandMilliseconds );

/* ASVGF_CREATE_BAND_APRON: rows (full-resolution rows; gradient rows for the gradient a-trous passes) a pass computes
   beyond the owned band.  For ASVGF_PASS_GRADIENT_SAMPLES (computed on owned rows only) it is the depth in gradient
   rows of the one exchange of ASVGF_IMAGE_DIS_PING_GRADIENT that follows it.  -1 for a bad iteration count. */
ASVGF_API int32_t     asvgfBandApronRows( uint32_t atrousIterations, uint32_t pass );

/* Number of CUDA kernels this handle has launched since creation. */
ASVGF_API uint64_t    asvgfGetLaunchCount( AsvgfDenoiser denoiser );
/* Name of the kernel that implements a pass and its algorithmic HBM bytes per full-resolution pixel. */
ASVGF_API const char* asvgfPassName( uint32_t pass );
ASVGF_API const char* asvgfPassKernelName( uint32_t pass, uint32_t atrousIterations ); /* default kernel choice per iteration */
ASVGF_API double      asvgfPassBytesPerPixel( uint32_t pass, uint32_t atrousIterations );
ASVGF_API const char* asvgfResultString( AsvgfResult result );
ASVGF_API const char* asvgfLastErrorMessage( void ); /* thread-local detail of the last failure */

#ifdef __cplusplus
}
#endif

#endif /* ASVGF_H_ */
