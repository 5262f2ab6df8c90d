/*
 * asvgf_layout.h -- the data contract of the A-SVGF denoiser path: image table indices, storage formats
 * and the GlobalUniform byte layout.  These are FACTS about the reference's interface (numbers, orders,
 * byte offsets), restated so that an RTGL1 host could hand its own tables over unchanged.
 *
 * Reference (paths relative to the RTGL1 source tree):
 *   image indices ........ Source/Generated/ShaderCommonGLSL.h:408-483   (FB_IMAGE_INDEX_*)
 *   image formats ........ Source/Generated/GenerateShaderCommon.py:802-887
 *   1/3-resolution rule .. Source/Framebuffers.cpp:665-668,691-699       ((W+1)/3, (H+1)/3)
 *   cur/prev swap rule ... Source/Framebuffers.cpp:35-50
 *   uniform layout ....... Source/Generated/ShaderCommonC.h:209-304      (sizeof == 2096)
 *   constants ............ Source/Generated/ShaderCommonGLSL.h:136-158
 */
#ifndef ASVGF_LAYOUT_H_
#define ASVGF_LAYOUT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- image table: same numbering as FB_IMAGE_INDEX_* so descriptor tables map 1:1 ---- */
enum AsvgfImageIndex
{
    ASVGF_IMAGE_ALBEDO                       = 0,  /* B10G11R11_UFLOAT, regular layout            */
    ASVGF_IMAGE_IS_SKY                       = 1,  /* R8_UINT                                     */
    ASVGF_IMAGE_NORMAL                       = 2,  /* R32_UINT (16-bit phi | 16-bit theta)        */
    ASVGF_IMAGE_NORMAL_PREV                  = 3,
    ASVGF_IMAGE_METALLIC_ROUGHNESS           = 4,  /* R8G8_UNORM                                  */
    ASVGF_IMAGE_METALLIC_ROUGHNESS_PREV      = 5,
    ASVGF_IMAGE_DEPTH_WORLD                  = 6,  /* R16_SFLOAT                                  */
    ASVGF_IMAGE_DEPTH_WORLD_PREV             = 7,
    ASVGF_IMAGE_DEPTH_GRAD                   = 8,  /* R16_SFLOAT                                  */
    ASVGF_IMAGE_MOTION                       = 10, /* R16G16B16A16_SFLOAT                         */
    ASVGF_IMAGE_UNFILTERED_DIRECT            = 11, /* E5B9G9R9 in R32_UINT                        */
    ASVGF_IMAGE_UNFILTERED_SPECULAR          = 12,
    ASVGF_IMAGE_UNFILTERED_INDIR             = 13,
    ASVGF_IMAGE_SURFACE_POSITION             = 14, /* R32G32B32A32_SFLOAT                         */
    ASVGF_IMAGE_SURFACE_POSITION_PREV        = 15,
    ASVGF_IMAGE_VIEW_DIRECTION               = 18, /* R16G16B16A16_SFLOAT                         */
    ASVGF_IMAGE_THROUGHPUT                   = 20, /* R16G16B16A16_SFLOAT                         */
    ASVGF_IMAGE_PRE_FINAL                    = 21, /* B10G11R11_UFLOAT, regular layout  (OUTPUT)  */
    ASVGF_IMAGE_FINAL                        = 22, /* B10G11R11_UFLOAT, regular layout: PreFinal after the checkerboard resolve
                                                      (CmCheckerboard.comp; written by asvgfResolveCheckerboard)           */
    ASVGF_IMAGE_ACCUM_HISTORY_LENGTH         = 27, /* R16G16B16A16_SFLOAT                         */
    ASVGF_IMAGE_ACCUM_HISTORY_LENGTH_PREV    = 28,
    ASVGF_IMAGE_DIFF_TEMPORARY               = 29, /* E5                                          */
    ASVGF_IMAGE_DIFF_ACCUM_COLOR             = 30, /* E5                                          */
    ASVGF_IMAGE_DIFF_ACCUM_COLOR_PREV        = 31,
    ASVGF_IMAGE_DIFF_ACCUM_MOMENTS           = 32, /* R16G16_SFLOAT                               */
    ASVGF_IMAGE_DIFF_ACCUM_MOMENTS_PREV      = 33,
    ASVGF_IMAGE_DIFF_COLOR_HISTORY           = 34, /* R16G16B16A16_SFLOAT                         */
    ASVGF_IMAGE_DIFF_PING_COLOR_AND_VARIANCE = 35, /* R16G16B16A16_SFLOAT                         */
    ASVGF_IMAGE_DIFF_PONG_COLOR_AND_VARIANCE = 36,
    ASVGF_IMAGE_SPEC_ACCUM_COLOR             = 37, /* E5                                          */
    ASVGF_IMAGE_SPEC_ACCUM_COLOR_PREV        = 38,
    ASVGF_IMAGE_SPEC_PING_COLOR              = 39, /* E5                                          */
    ASVGF_IMAGE_SPEC_PONG_COLOR              = 40,
    ASVGF_IMAGE_INDIR_ACCUM                  = 41, /* E5                                          */
    ASVGF_IMAGE_INDIR_ACCUM_PREV             = 42,
    ASVGF_IMAGE_INDIR_PING                   = 43, /* E5                                          */
    ASVGF_IMAGE_INDIR_PONG                   = 44,
    ASVGF_IMAGE_ATROUS_FILTERED_VARIANCE     = 45, /* R16_SFLOAT                                  */
    ASVGF_IMAGE_HISTOGRAM_INPUT              = 46, /* R16_SFLOAT, regular layout        (OUTPUT)  */
    ASVGF_IMAGE_SCATTERING                   = 48, /* R16G16B16A16_SFLOAT, regular layout: accumulated in-scattering rgb + transmittance (asvgfScatterAccumulate) */
    ASVGF_IMAGE_SCATTERING_PREV              = 49,
    ASVGF_IMAGE_SCATTERING_HISTORY           = 50, /* R16_SFLOAT, regular layout: its history length */
    ASVGF_IMAGE_SCATTERING_HISTORY_PREV      = 51,
    ASVGF_IMAGE_ACID_FOG_RT                  = 52, /* B10G11R11_UFLOAT, regular layout: the fog producer's checkerboarded image ...  */
    ASVGF_IMAGE_ACID_FOG                     = 53, /* ... and its resolved form (asvgfResolveCheckerboard, CmCheckerboard.comp:65-77) */
    ASVGF_IMAGE_SCREEN_EMIS_RT               = 54, /* B10G11R11_UFLOAT, regular layout: screen-space emission, checkerboarded ...      */
    ASVGF_IMAGE_SCREEN_EMISSION              = 55, /* ... and resolved                                                                 */
    ASVGF_IMAGE_BLOOM_INPUT                  = 56, /* B10G11R11_UFLOAT, regular layout: what the bloom chain reads (asvgfPrepareFinal)          */
    ASVGF_IMAGE_GRADIENT_INPUTS              = 70, /* R16G16_SFLOAT                               */
    ASVGF_IMAGE_GRADIENT_INPUTS_PREV         = 71,
    ASVGF_IMAGE_DIS_PING_GRADIENT            = 72, /* R8G8B8A8_UNORM, 1/3 resolution              */
    ASVGF_IMAGE_DIS_PONG_GRADIENT            = 73,
    ASVGF_IMAGE_DIS_GRADIENT_HISTORY         = 74,
    ASVGF_IMAGE_GRADIENT_PREV_PIX            = 75, /* R8_UINT, 1/3 resolution                     */
    ASVGF_IMAGE_REFERENCE_COUNT              = 76, /* ShFramebuffers_Count                        */

    /* extensions that have no counterpart in the reference's table */
    ASVGF_IMAGE_FWD_LUM    = 76, /* R16G16_SFLOAT: (lum diffuse, lum specular) of the forward-projected, re-shaded
                                    sample; stands in for the ray-traced step RtGradients.rgen:107-136            */
    ASVGF_IMAGE_DEBUG_MASK = 77, /* R8_UINT: bit0..3 per-tap isConsistent, bit4 weightSum>1e-4,
                                    bit5 weightSumVirtual>1e-4 (CmSVGFTemporalAccumulation.comp:159-176,202,282)  */
    ASVGF_IMAGE_SCATTER_CURRENT = 78, /* R16G16B16A16_SFLOAT, regular layout: this frame's scattering sample when
                                    volumeEnableType == VOLUMETRIC; stands in for the 3-D illumination-volume lookup
                                    volume_sampleDithered of CmScatterAccum.comp:41-73 (not an image-space operation) */
    ASVGF_IMAGE_COUNT      = 79
};

/* ShGlobalUniform.volumeEnableType (ShaderCommonC.h:195-197) */
#define ASVGF_VOLUME_ENABLE_NONE 0u
#define ASVGF_VOLUME_ENABLE_SIMPLE 1u
#define ASVGF_VOLUME_ENABLE_VOLUMETRIC 2u

enum AsvgfFormat
{
    ASVGF_FORMAT_UNDEFINED = 0,
    ASVGF_FORMAT_R8_UINT,
    ASVGF_FORMAT_R8G8_UNORM,
    ASVGF_FORMAT_R8G8B8A8_UNORM,
    ASVGF_FORMAT_R16_SFLOAT,
    ASVGF_FORMAT_R16G16_SFLOAT,
    ASVGF_FORMAT_R16G16B16A16_SFLOAT,
    ASVGF_FORMAT_R32_UINT,            /* encoded normal or E5B9G9R9 colour */
    ASVGF_FORMAT_R32G32B32A32_SFLOAT,
    ASVGF_FORMAT_B10G11R11_UFLOAT
};

/* constants of the path (ShaderCommonGLSL.h:136-158) */
#define ASVGF_STRATA_SIZE 3
#define ASVGF_GRADIENT_ATROUS_ITERATIONS 4
#define ASVGF_REFERENCE_ATROUS_ITERATIONS 4
#define ASVGF_MAX_ATROUS_ITERATIONS 8
#define ASVGF_MAX_RAY_LENGTH 10000.0f
#define ASVGF_GRADIENT_PREV_PIX_NONE 10 /* STRATA*STRATA+1, RtGradients.rgen:59 */

#define ASVGF_DEBUG_SHOW_FLAG_ONLY_DIRECT_DIFFUSE (1u << 5)
#define ASVGF_DEBUG_SHOW_FLAG_ONLY_SPECULAR (1u << 6)
#define ASVGF_DEBUG_SHOW_FLAG_ONLY_INDIRECT_DIFFUSE (1u << 7)
#define ASVGF_DEBUG_SHOW_FLAG_ALBEDO_WHITE (1u << 9)

/*
 * Byte-compatible with ShGlobalUniform (ShaderCommonC.h:209-304).  Only the members the denoiser path reads
 * are named; everything else is reserved padding so a pointer to the reference's own struct can be passed.
 * Matrices are column-major (m[col*4+row]) and multiply column vectors.
 */
typedef struct AsvgfGlobalUniform
{
    float    view[ 16 ];           /*    0 */
    float    invView[ 16 ];        /*   64 */
    float    viewPrev[ 16 ];       /*  128 */
    float    projection[ 16 ];     /*  192 */
    float    invProjection[ 16 ];  /*  256 */
    float    projectionPrev[ 16 ]; /*  320 */
    uint32_t _reserved0[ 65 ];     /*  384 */
    float    lightmapScreenCoverage; /* 644 */
    uint32_t _reserved1;           /*  648 */
    float    renderWidth;          /*  652  (a float in the reference too) */
    float    renderHeight;         /*  656 */
    uint32_t frameId;              /*  660 */
    float    timeDelta;            /*  664 */
    float    minLogLuminance;      /*  668  exposure histogram range (CmLuminanceHistogram.comp, CmLuminanceAvg.comp) */
    float    maxLogLuminance;      /*  672 */
    float    luminanceWhitePoint;  /*  676  asvgfPrepareFinal, classic-shading columns (CmPrepareFinal.comp:89-93) */
    uint32_t stopEyeAdaptation;    /*  680 */
    uint32_t _reserved2[ 9 ];      /*  684 */
    float    cameraPosition[ 4 ];  /*  720 */
    float    cameraPositionPrev[ 4 ]; /* 736 */
    uint32_t debugShowFlags;       /*  752 */
    uint32_t _reserved3[ 4 ];      /*  756 */
    float    emissionMaxScreenColor; /* 772  asvgfPrepareFinal (CmPrepareFinal.comp:119) */
    uint32_t _reserved3a[ 7 ];     /*  776 */
    float    bloomThreshold;       /*  804  asvgfPrepareFinal (CmPrepareFinal.comp:139) */
    float    bloomEmissionMultiplier; /* 808  (CmPrepareFinal.comp:128) */
    uint32_t _reserved3c[ 5 ];     /*  812 */
    uint32_t volumeEnableType;     /*  832  ASVGF_VOLUME_ENABLE_* (asvgfScatterAccumulate) */
    float    volumeScattering;     /*  836 */
    uint32_t _reserved3b[ 23 ];    /*  840 */
    float    jitterX;              /*  932 */
    float    jitterY;              /*  936 */
    uint32_t _reserved4[ 6 ];      /*  940 */
    float    gradientMultDiffuse;  /*  964 */
    float    gradientMultIndirect; /*  968 */
    float    gradientMultSpecular; /*  972 */
    uint32_t _reserved5[ 3 ];      /*  976 */
    uint32_t antiFireflyEnabled;   /*  988 */
    float    volumeAmbient[ 4 ];   /*  992 */
    uint32_t _reserved6[ 272 ];    /* 1008 */
} AsvgfGlobalUniform;              /* 2096 */

/* Byte-compatible with ShTonemapping (ShaderCommonC.h:344-348): the exposure state asvgfCalculateExposure keeps. */
#define ASVGF_LUM_HISTOGRAM_BIN_COUNT 256 /* COMPUTE_LUM_HISTOGRAM_BIN_COUNT, ShaderCommonC.h:138 */
typedef struct AsvgfTonemapping
{
    uint32_t histogram[ ASVGF_LUM_HISTOGRAM_BIN_COUNT ];
    float    avgLuminance;
} AsvgfTonemapping;

#ifdef __cplusplus
}
static_assert( sizeof( AsvgfGlobalUniform ) == 2096, "must match ShGlobalUniform" );
static_assert( offsetof( AsvgfGlobalUniform, minLogLuminance ) == 668, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, stopEyeAdaptation ) == 680, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, volumeEnableType ) == 832, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, emissionMaxScreenColor ) == 772, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, bloomThreshold ) == 804, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, bloomEmissionMultiplier ) == 808, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, luminanceWhitePoint ) == 676, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, volumeAmbient ) == 992, "layout" );
static_assert( sizeof( AsvgfTonemapping ) == 1028, "must match ShTonemapping" );
static_assert( offsetof( AsvgfGlobalUniform, viewPrev ) == 128, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, projectionPrev ) == 320, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, lightmapScreenCoverage ) == 644, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, renderWidth ) == 652, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, timeDelta ) == 664, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, cameraPosition ) == 720, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, cameraPositionPrev ) == 736, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, debugShowFlags ) == 752, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, jitterX ) == 932, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, gradientMultDiffuse ) == 964, "layout" );
static_assert( offsetof( AsvgfGlobalUniform, antiFireflyEnabled ) == 988, "layout" );
#endif

#endif /* ASVGF_LAYOUT_H_ */
