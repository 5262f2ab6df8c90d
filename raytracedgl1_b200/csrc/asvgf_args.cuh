// asvgf_args.cuh -- launch geometry and kernel argument blocks of the A-SVGF pass kernels: one block per reference
// dispatch (Source/Denoiser.cpp:87-254), holding the views of the images that dispatch binds (descriptor set 0).
#pragma once

#include "asvgf_device.cuh"

namespace asvgf
{

constexpr int kBlockW = 32; // one-thread-per-pixel kernels: 32 consecutive x per warp (full 128-byte row segments)
constexpr int kBlockH = 8;

// G: gradient samples, RtGradients.rgen:44-181 (image-space part; the re-shade is read from FwdLum)
struct GradientSamplesArgs
{
    Img     gradientPrevPix, disPing, disHistory, motion, gradientInputsPrev, isSky, indirAccumPrev, fwdLum;
    int32_t gyBegin, gyEnd; // gradient rows to compute
    // band handles: set to 1 when a reprojected tap lies inside the frame but outside the rows this band stores (the motion
    // reached beyond the halo: the band would silently diverge from the single-GPU frame); nullptr on whole-frame handles
    uint32_t* haloMiss;
};

// GA: gradient a-trous, CmASVGFGradientAtrous.comp:42-105
struct GradientAtrousArgs
{
    Img     src, dst, history;
    int32_t iteration;
    int32_t gyBegin, gyEnd;
};

// T: temporal accumulation, CmSVGFTemporalAccumulation.comp:58-385
struct TemporalArgs
{
    Img unfDirect, unfIndir, unfSpec, viewDir, motion, depth, normal, metRough, surfPos, disPing;
    Img depthPrev, normalPrev, diffHistory, momentsPrev, histLenPrev, specPrev, indirPrev, surfPosPrev;
    Img outDiffTemporary, outMoments, outHistLen, outSpecAccum, outIndirAccum, outMask;
    Img guide;               // (n.xyz, depth | roughness byte) of the frame for the a-trous iterations (asvgf_atrous_tiled.cuh)
    NormalLut lut;
    int32_t   writeMask;
    int32_t   writeGuide;    // this launch covers every row the a-trous iterations will read: write the guide plane here
    uint32_t* haloMiss;      // band handles: see GradientSamplesArgs
};

// F: anti-firefly, CmAntiFirefly.comp:47-159
struct AntiFireflyArgs
{
    Img src[ 3 ]; // DiffTemporary, SpecAccumColor, IndirAccum
    Img dst[ 3 ]; // DiffAccumColor, SpecPingColor, IndirPing
};

// V: variance estimation, CmSVGFEstimateVariance.comp:74-205
struct VarianceArgs
{
    Img       color, normal, depth, histLen, moments, out;
    NormalLut lut;
};

// A: a-trous, CmSVGFAtrous_Iter0.comp:92-335 (iteration 0) and CmSVGFAtrous.comp:65-349 (the last one composes PreFinal)
struct AtrousArgs
{
    Img       srcDiff, srcSpec, srcIndir, dstDiff, dstSpec, dstIndir;
    Img       depth, depthGrad, normal, metRough, histLen, filteredVariance;
    Img       albedo, throughput, isSky, preFinal, histogramInput;
    Img       guide;       // (n.xyz, depth | roughness byte) of the current frame, written by the temporal pass or k_guide
    Img       centre;      // (kLum, invK1, invK2, P) per pixel: the iteration-independent weight constants, written by iteration 0
    int32_t   centreValid; // iterations >= 1: read `centre` instead of re-deriving the constants from three more images
    int32_t   uniformPitch; // guide, srcDiff, srcSpec and srcIndir hold the same number of texels per row: one texel offset per staged texel
    NormalLut lut;
    int32_t   iteration;
};

} // namespace asvgf
