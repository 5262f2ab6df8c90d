// asvgf_kernels.cuh -- the A-SVGF pass kernels (sm_100a), one per reference dispatch ("unfused" set).
//
//   k_gradient_samples ... Source/Shaders/RtGradients.rgen:44-181 (image-space part; re-shade read from FwdLum)
//   k_gradient_atrous .... Source/Shaders/CmASVGFGradientAtrous.comp:42-105
//   k_temporal ........... Source/Shaders/CmSVGFTemporalAccumulation.comp:58-385
//   k_antifirefly ........ Source/Shaders/CmAntiFirefly.comp:47-159
//   k_variance ........... Source/Shaders/CmSVGFEstimateVariance.comp:74-205
//   k_atrous<ITER0,LAST> . Source/Shaders/CmSVGFAtrous_Iter0.comp:92-335, CmSVGFAtrous.comp:65-349
//
// Every kernel is HBM-bound image work: one thread per pixel, 32 consecutive x per warp so that each image row
// segment a warp touches is one or two full 128-byte lines; packed storage formats are kept end to end.
#pragma once

#include "../raytracedgl1_b200/csrc/asvgf_args.cuh"

namespace asvgf
{


// ------------------------------------------------------------------------------------------------------------
// G: gradient samples
// ------------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__( kBlockW* kBlockH ) k_gradient_samples( const __grid_constant__ GradientSamplesArgs a, const __grid_constant__ Uniform u )
{
    const int gx = blockIdx.x * kBlockW + threadIdx.x;
    const int gy = a.gyBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( gx >= a.disPing.w || gy >= a.gyEnd ) return;
    constexpr int S = 3;
    if( gx * S >= u.W || gy * S >= u.H ) return; // :49-52

    if( u.gradientMult[ 0 ] < 0.001f && u.gradientMult[ 1 ] < 0.001f && u.gradientMult[ 2 ] < 0.001f ) // :54-61
    {
        storeU32( a.disPing, gx, gy, 0u );
        storeU8( a.gradientPrevPix, gx, gy, 10u );
        return;
    }
    const uint32_t prevStrataPos = fetchU8( a.gradientPrevPix, gx, gy ); // :63

    int      prevX = 0, prevY = 0, curX = 0, curY = 0; // :66-96
    uint32_t curStrataPos = 0;
    float2   prevLum      = make_float2( 0.0f, 0.0f );
#pragma unroll
    for( int yy = 0; yy < S; yy++ )
    {
#pragma unroll
        for( int xx = 0; xx < S; xx++ )
        {
            if( ( uint32_t )( yy * S + xx ) == prevStrataPos ) continue;
            const int    cx = gx * S + xx, cy = gy * S + yy;
            const float4 m  = fetchRGBA16F( a.motion, cx, cy );
            const float2 pp = prevScreenPos( u, m.x, m.y, cx, cy );
            const int    px = f2i( floorf( pp.x ) ), py = f2i( floorf( pp.y ) );
            const float2 pl = fetchRG16F( a.gradientInputsPrev, px, py );
            if( fmaxf( pl.x, pl.y ) > fmaxf( prevLum.x, prevLum.y ) )
            {
                prevX = px; prevY = py; curX = cx; curY = cy;
                curStrataPos = ( uint32_t )( yy * S + xx );
                prevLum      = pl;
            }
        }
    }

    if( !testPixInRenderArea( prevX, prevY, checkerboardedRenderArea( u, curX ) ) || fetchU8( a.isSky, curX, curY ) != 0u ||
        ( prevLum.x <= 0.0f && prevLum.y <= 0.0f ) ) // :98-105
    {
        storeU32( a.disPing, gx, gy, 0u );
        storeU8( a.gradientPrevPix, gx, gy, 10u );
        return;
    }

    const float2 fwd = fetchRG16F( a.fwdLum, curX, curY ); // stands in for :107-136

    float gD, gI, gS;
    {
        gD = antilagAlpha( subr( fwd.x, prevLum.x ), fmaxf( fwd.x, prevLum.x ) ); // :141-146
    }
    {
        float       gradSample = fmaxf( 0.0f, subr( prevLum.x, fwd.x ) ); // :147-157
        const float brightness = luminance( fetchE5( a.indirAccumPrev, prevX, prevY ) );
        gradSample             = mulr( gradSample, subr( 1.0f, smoothstepr( 1.0f, 10.0f, brightness ) ) );
        gI                     = antilagAlpha( gradSample, fmaxf( fwd.x, prevLum.x ) );
    }
    {
        gS = antilagAlpha( subr( fwd.y, prevLum.y ), fmaxf( fwd.y, prevLum.y ) ); // :158-163
    }
    gD = mulr( gD, u.gradientMult[ 0 ] ); // :166-168
    gI = mulr( gI, u.gradientMult[ 1 ] );
    gS = mulr( gS, u.gradientMult[ 2 ] );
    {
        const float4 prev = fetchRGBA8( a.disHistory, prevX / S, prevY / S ); // :171-177
        gD = mixr( gD, prev.x, 0.5f );
        gI = mixr( gI, prev.y, 0.5f );
        gS = mixr( gS, prev.z, 0.5f );
    }
    storeRGBA8( a.disPing, gx, gy, make_float4( gD, gI, gS, 0.0f ) ); // :179-180
    storeU8( a.gradientPrevPix, gx, gy, curStrataPos );
}

// ------------------------------------------------------------------------------------------------------------
// GA: gradient a-trous (3x3, weights {1,.5,.25}, stride 1<<iteration, re-quantised to unorm8 every iteration)
// ------------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__( kBlockW* kBlockH ) k_gradient_atrous( const __grid_constant__ GradientAtrousArgs a, const __grid_constant__ Uniform u )
{
    const int gx = blockIdx.x * kBlockW + threadIdx.x;
    const int gy = a.gyBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( gx >= a.dst.w || gy >= a.gyEnd ) return;
    constexpr int S = 3;
    if( gx * S >= u.W || gy * S >= u.H ) return; // :82-85

    const Area full = checkerboardedRenderArea( u, gx * S ); // :45, integer division of each bound
    const Area area = Area{ full.x0 / S, full.x1 / S, full.h / S };
    const int  step = 1 << a.iteration;

    float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f, weightSum = 0.0f;
#pragma unroll
    for( int yy = -1; yy <= 1; yy++ )
    {
#pragma unroll
        for( int xx = -1; xx <= 1; xx++ )
        {
            const int    qx = gx + xx * step, qy = gy + yy * step;
            const float4 g  = fetchRGBA8( a.src, qx, qy );
            const float  k  = ( yy == 0 ? 1.0f : 0.5f ) * ( xx == 0 ? 1.0f : 0.5f );
            const float  w  = mulr( k, testPixInRenderArea( qx, qy, area ) ? 1.0f : 0.0f );
            o0        = addr( o0, mulr( g.x, w ) );
            o1        = addr( o1, mulr( g.y, w ) );
            o2        = addr( o2, mulr( g.z, w ) );
            weightSum = addr( weightSum, w );
        }
    }
    const float4 r = make_float4( divr( o0, weightSum ), divr( o1, weightSum ), divr( o2, weightSum ), 0.0f ); // :74
    storeRGBA8( a.dst, gx, gy, r );
    if( a.iteration == 0 ) storeRGBA8( a.history, gx, gy, r ); // :99-100
}

// ------------------------------------------------------------------------------------------------------------
// T: temporal accumulation
// ------------------------------------------------------------------------------------------------------------

ASVGF_DEVICE float specAccumSpeed( float minFrames, float Amax, float roughness, float NoV, float parallax ) // :80-92
{
    const float acos01sq            = subr( 1.0f, NoV );
    const float a                   = powf( saturatef( acos01sq ), 0.66f );
    const float b                   = addr( 1.1f, mulr( roughness, roughness ) );
    const float parallaxSensitivity = divr( addr( b, a ), subr( b, a ) );
    const float powerScale          = addr( 1.0f, mulr( parallax, parallaxSensitivity ) );
    float       f                   = subr( 1.0f, exp2f( mulr( mulr( -200.0f, roughness ), roughness ) ) );
    f                               = mulr( f, powf( saturatef( roughness ), mulr( 0.5f, powerScale ) ) );
    const float A                   = mulr( 256.0f, f );
    return clampf( A, minFrames, Amax );
}
ASVGF_DEVICE float specularDominantFactor( float nv, float roughness ) // :93-99
{
    const float a = mulr( 0.298475f, logf( subr( 39.4115f, mulr( 39.0029f, roughness ) ) ) );
    const float f = addr( mulr( powf( clampf( subr( 1.0f, nv ), 0.0f, 1.0f ), 10.8649f ), subr( 1.0f, a ) ), a );
    return clampf( f, 0.0f, 1.0f );
}

__global__ void __launch_bounds__( kBlockW* kBlockH ) k_temporal( const __grid_constant__ TemporalArgs a, const __grid_constant__ Uniform u )
{
    const int px = blockIdx.x * kBlockW + threadIdx.x;
    const int py = u.yBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( px >= u.W || py >= u.yEnd ) return; // :105-108

    const float3 unfilteredDiff    = fetchE5( a.unfDirect, px, py ); // :110-114
    const float3 unfilteredIndir   = fetchE5( a.unfIndir, px, py );
    const float3 unfilteredSpec    = fetchE5( a.unfSpec, px, py );
    const float4 unfilteredRayInfo = fetchRGBA16F( a.viewDir, px, py );
    const float3 rayDir            = make_float3( unfilteredRayInfo.x, unfilteredRayInfo.y, unfilteredRayInfo.z );

    const Area   chRenderArea = checkerboardedRenderArea( u, px ); // :116-120
    const float4 motion4      = fetchRGBA16F( a.motion, px, py );
    const float  motionZ      = motion4.z;
    const float  depth        = fetchR16F( a.depth, px, py );
    const float3 normal       = decodeNormal( a.lut, fetchU32( a.normal, px, py ) );
    const float  roughness    = fetchRG8( a.metRough, px, py ).y;

    const float l  = luminance( unfilteredDiff ); // :122-123
    const float l2 = mulr( l, l );

    int   pixPrevX, pixPrevY; // :126-137
    float subX, subY;
    {
        const float2 posPrev = prevScreenPos( u, motion4.x, motion4.y, px, py );
        const float  qx = subr( posPrev.x, 0.5f ), qy = subr( posPrev.y, 0.5f );
        const float  fx = floorf( qx ), fy = floorf( qy );
        pixPrevX = f2i( fx ); pixPrevY = f2i( fy );
        subX = subr( qx, fx ); subY = subr( qy, fy );
    }
    const float bw[ 4 ] = { mulr( subr( 1.0f, subX ), subr( 1.0f, subY ) ), mulr( subX, subr( 1.0f, subY ) ),
                            mulr( subr( 1.0f, subX ), subY ), mulr( subX, subY ) };

    float4 historyLengthDIS = make_float4( 0.0f, 0.0f, 0.0f, 0.0f ); // :141-146
    float3 diffPrev = make_float3( 0.0f, 0.0f, 0.0f ), specPrevOnSurf = diffPrev, indirPrev = diffPrev;
    float2 momentsPrev = make_float2( 0.0f, 0.0f );
    float  weightSum   = 0.0f;
    uint32_t mask = 0;

#pragma unroll
    for( int t = 0; t < 4; t++ ) // :149-178
    {
        const int    x = pixPrevX + ( t & 1 ), y = pixPrevY + ( t >> 1 );
        const float  depthPrev  = fetchR16F( a.depthPrev, x, y );
        const float3 normalPrev = decodeNormal( a.lut, fetchU32( a.normalPrev, x, y ) );
        const bool   isConsistent = testPixInRenderArea( x, y, chRenderArea ) && testReprojectedDepth( depth, depthPrev, motionZ ) &&
                                  testReprojectedNormal( normal, normalPrev );
        if( isConsistent )
        {
            const float  w  = bw[ t ];
            const float4 dh = fetchRGBA16F( a.diffHistory, x, y );
            diffPrev        = diffPrev + make_float3( dh.x, dh.y, dh.z ) * w;
            const float2 mp = fetchRG16F( a.momentsPrev, x, y );
            momentsPrev     = make_float2( addr( momentsPrev.x, mulr( mp.x, w ) ), addr( momentsPrev.y, mulr( mp.y, w ) ) );
            const float4 hl = fetchRGBA16F( a.histLenPrev, x, y );
            historyLengthDIS = make_float4( addr( historyLengthDIS.x, mulr( hl.x, w ) ), addr( historyLengthDIS.y, mulr( hl.y, w ) ),
                                            addr( historyLengthDIS.z, mulr( hl.z, w ) ), addr( historyLengthDIS.w, mulr( hl.w, w ) ) );
            specPrevOnSurf  = specPrevOnSurf + fetchE5( a.specPrev, x, y ) * w;
            indirPrev       = indirPrev + fetchE5( a.indirPrev, x, y ) * w;
            weightSum       = addr( weightSum, w );
            mask |= 1u << t;
        }
    }

    const float4 gradDIS           = fetchRGBA8( a.disPing, px / 3, py / 3 ); // :181-194
    const float  antilagAlphaDiff  = gradDIS.x;
    const float  antilagAlphaIndir = gradDIS.y;
    const float  antilagAlphaSpec  = gradDIS.z;

    float3 specColorAccum;
    float  specHistoryLengthAccum, specHitDistanceAccum;

    if( weightSum > 0.0001f ) // :202-309
    {
        mask |= 1u << 4;
        specPrevOnSurf = make_float3( divr( specPrevOnSurf.x, weightSum ), divr( specPrevOnSurf.y, weightSum ), divr( specPrevOnSurf.z, weightSum ) );
        const float specHistoryLength   = divr( historyLengthDIS.z, weightSum );
        const float specHitDistancePrev = divr( historyLengthDIS.w, weightSum );

        const float4 sp4  = fetchRGBA32F( a.surfPos, px, py );
        const float4 spp4 = fetchRGBA32F( a.surfPosPrev, pixPrevX, pixPrevY );
        const float3 surfPos = make_float3( sp4.x, sp4.y, sp4.z ), surfPosPrev = make_float3( spp4.x, spp4.y, spp4.z );
        const float  nv = fmaxf( 0.0f, dot3r( normal, make_float3( -rayDir.x, -rayDir.y, -rayDir.z ) ) );

        const float minFrames = clampf( mulr( 0.2f, safePositiveRcp( u.timeDelta ) ), 0.0f, 128.0f ); // :70-74
        const float A = clampf( addr( 1.0f, mulr( specHistoryLength, powf( subr( 1.0f, antilagAlphaSpec ), 5.0f ) ) ), minFrames, 256.0f );

        float parallax; // :58-69
        {
            const float3 cameraDelta = make_float3( subr( u.cameraPositionPrev[ 0 ], u.cameraPosition[ 0 ] ), subr( u.cameraPositionPrev[ 1 ], u.cameraPosition[ 1 ] ),
                                                    subr( u.cameraPositionPrev[ 2 ], u.cameraPosition[ 2 ] ) );
            const float3 v     = normalize3r( surfPos );
            const float3 vPrev = normalize3r( surfPosPrev - cameraDelta );
            const float  cosa  = saturatef( dot3r( v, vPrev ) );
            parallax           = divr( __fsqrt_rn( subr( 1.0f, mulr( cosa, cosa ) ) ), fmaxf( cosa, 1e-6f ) );
            parallax           = mulr( parallax, 60.0f );
        }
        const float AOnSurf = specAccumSpeed( minFrames, A, roughness, nv, parallax );

        const float  aOnSurf       = divr( 1.0f, addr( 1.0f, AOnSurf ) );
        const float3 specOnSurf    = mix3r( specPrevOnSurf, unfilteredSpec, aOnSurf );
        const float  hitDistOnSurf = mixr( specHitDistancePrev, unfilteredRayInfo.w, aOnSurf );

        specColorAccum         = specOnSurf;
        specHistoryLengthAccum = AOnSurf;
        specHitDistanceAccum   = hitDistOnSurf;

        const float  domFactor  = specularDominantFactor( nv, roughness );
        const float3 virtualPos = surfPos + ( rayDir * hitDistOnSurf ) * domFactor; // :228

        int   pixPrevSpecX, pixPrevSpecY; // :232-249
        float subSX, subSY;
        {
            const float4 vp        = make_float4( virtualPos.x, virtualPos.y, virtualPos.z, 1.0f );
            const float4 viewCur   = mulMat( u.view, vp );
            const float4 viewPrev  = mulMat( u.viewPrev, vp );
            const float4 clipCur   = mulMat( u.projection, viewCur );
            const float4 clipPrev  = mulMat( u.projectionPrev, viewPrev );
            const float  scx = addr( mulr( divr( clipCur.x, clipCur.w ), 0.5f ), 0.5f ), scy = addr( mulr( divr( clipCur.y, clipCur.w ), 0.5f ), 0.5f );
            const float  spx = addr( mulr( divr( clipPrev.x, clipPrev.w ), 0.5f ), 0.5f ), spy = addr( mulr( divr( clipPrev.y, clipPrev.w ), 0.5f ), 0.5f );
            const float2 specPosPrev = prevScreenPos( u, subr( spx, scx ), subr( spy, scy ), px, py );
            const float  qx = subr( specPosPrev.x, 0.5f ), qy = subr( specPosPrev.y, 0.5f );
            const float  fx = floorf( qx ), fy = floorf( qy );
            pixPrevSpecX = f2i( fx ); pixPrevSpecY = f2i( fy );
            subSX = subr( qx, fx ); subSY = subr( qy, fy );
        }
        const float bws[ 4 ] = { mulr( subr( 1.0f, subSX ), subr( 1.0f, subSY ) ), mulr( subSX, subr( 1.0f, subSY ) ),
                                 mulr( subr( 1.0f, subSX ), subSY ), mulr( subSX, subSY ) };
        float weightSumVirtual = 0.0f; // :256-280; the shader's specPrevVirtual is dead code (it blends specPrevOnSurf at :298)
#pragma unroll
        for( int t = 0; t < 4; t++ )
        {
            const int    x = pixPrevSpecX + ( t & 1 ), y = pixPrevSpecY + ( t >> 1 );
            const float  depthPrev  = fetchR16F( a.depthPrev, x, y );
            const float3 normalPrev = decodeNormal( a.lut, fetchU32( a.normalPrev, x, y ) );
            const bool   isConsistent = testPixInRenderArea( x, y, chRenderArea ) && testReprojectedDepth( depth, depthPrev, motionZ ) &&
                                      testReprojectedNormal( normal, normalPrev );
            if( isConsistent ) weightSumVirtual = addr( weightSumVirtual, bws[ t ] );
        }

        if( weightSumVirtual > 0.0001f ) // :282-307
        {
            mask |= 1u << 5;
            float       Avirt    = specAccumSpeed( minFrames, A, roughness, nv, 0.0f );
            const float Amin     = fminf( Avirt, mulr( 3.0f, __fsqrt_rn( roughness ) ) );
            float       al       = mixr( divr( 1.0f, addr( 1.0f, Amin ) ), divr( 1.0f, addr( 1.0f, Avirt ) ), 1.0f );
            Avirt                = subr( divr( 1.0f, al ), 1.0f );
            const float Ahitdist = fminf( Avirt, 16.0f );
            const float3 specVirt = mix3r( specPrevOnSurf, unfilteredSpec, divr( 1.0f, addr( 1.0f, Avirt ) ) );
            const float  amount   = domFactor;
            al                     = mixr( aOnSurf, divr( 1.0f, addr( 1.0f, Avirt ) ), amount );
            specColorAccum         = mix3r( specOnSurf, specVirt, amount );
            specHistoryLengthAccum = subr( divr( 1.0f, al ), 1.0f );
            specHitDistanceAccum   = mixr( specHitDistancePrev, unfilteredRayInfo.w, divr( 1.0f, addr( 1.0f, Ahitdist ) ) );
        }
    }
    else // :310-315
    {
        specColorAccum         = unfilteredSpec;
        specHistoryLengthAccum = 1.0f;
        specHitDistanceAccum   = unfilteredRayInfo.w;
    }

    float3 diffColorAccum, indirAccum;
    float2 diffMomentsAccum;
    float  diffHistoryLengthAccum, indirHistoryLengthAccum;
    if( weightSum > 0.0001f ) // :326-367
    {
        const float inv = divr( 1.0f, weightSum );
        diffPrev        = diffPrev * inv;
        momentsPrev     = make_float2( mulr( momentsPrev.x, inv ), mulr( momentsPrev.y, inv ) );
        indirPrev       = indirPrev * inv;
        float diffHistoryLength  = mulr( historyLengthDIS.x, inv );
        float indirHistoryLength = mulr( historyLengthDIS.y, inv );
        {
            diffHistoryLength = mulr( diffHistoryLength, powf( subr( 1.0f, antilagAlphaDiff ), 10.0f ) );
            diffHistoryLength = clampf( addr( diffHistoryLength, 1.0f ), 1.0f, 256.0f );
            float alpha       = fmaxf( 0.01f, divr( 1.0f, diffHistoryLength ) );
            alpha             = mixr( alpha, 1.0f, antilagAlphaDiff );
            diffColorAccum    = mix3r( diffPrev, unfilteredDiff, alpha );
            diffMomentsAccum  = make_float2( mixr( momentsPrev.x, l, alpha ), mixr( momentsPrev.y, l2, alpha ) );
            diffHistoryLengthAccum = diffHistoryLength;
        }
        {
            indirHistoryLength = mulr( indirHistoryLength, powf( subr( 1.0f, antilagAlphaIndir ), 10.0f ) );
            indirHistoryLength = clampf( addr( indirHistoryLength, 1.0f ), 1.0f, 256.0f );
            float alpha        = fmaxf( 0.01f, divr( 1.0f, indirHistoryLength ) );
            alpha              = mixr( alpha, 1.0f, antilagAlphaIndir );
            indirAccum         = mix3r( indirPrev, unfilteredIndir, alpha );
            indirHistoryLengthAccum = indirHistoryLength;
        }
    }
    else // :368-376
    {
        diffColorAccum = unfilteredDiff; diffMomentsAccum = make_float2( l, l2 ); diffHistoryLengthAccum = 1.0f;
        indirAccum = unfilteredIndir; indirHistoryLengthAccum = 1.0f;
    }

    storeU32( a.outDiffTemporary, px, py, encodeE5( diffColorAccum ) ); // :378-384
    storeRG16F( a.outMoments, px, py, diffMomentsAccum.x, diffMomentsAccum.y );
    storeRGBA16F( a.outHistLen, px, py, make_float4( diffHistoryLengthAccum, indirHistoryLengthAccum, specHistoryLengthAccum, specHitDistanceAccum ) );
    storeU32( a.outSpecAccum, px, py, encodeE5( specColorAccum ) );
    storeU32( a.outIndirAccum, px, py, encodeE5( indirAccum ) );
    if( a.writeMask ) storeU8( a.outMask, px, py, mask );
}

// ------------------------------------------------------------------------------------------------------------
// F: anti-firefly.  Clamp each colour component to the [min,max] of its 24 neighbours in 5x5 (centre excluded).
// Tile + 2-pixel halo of the three E5 images staged in shared memory; zeros outside the image participate.
// ------------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__( kBlockW* kBlockH ) k_antifirefly( const __grid_constant__ AntiFireflyArgs a, const __grid_constant__ Uniform u )
{
    constexpr int R = 2, TW = kBlockW + 2 * R, TH = kBlockH + 2 * R;
    __shared__ uint32_t tile[ 3 ][ TH ][ TW ];

    const int bx = blockIdx.x * kBlockW, by = u.yBegin + blockIdx.y * kBlockH;
    const int tid = threadIdx.y * kBlockW + threadIdx.x;
    for( int i = tid; i < TW * TH; i += kBlockW * kBlockH )
    {
        const int tx = i % TW, ty = i / TW;
        const int gx = bx + tx - R, gy = by + ty - R;
#pragma unroll
        for( int c = 0; c < 3; c++ ) tile[ c ][ ty ][ tx ] = fetchU32( a.src[ c ], gx, gy );
    }
    __syncthreads();

    const int px = bx + threadIdx.x, py = by + threadIdx.y;
    if( px >= u.W || py >= u.yEnd ) return;
    const int cx = threadIdx.x + R, cy = threadIdx.y + R;

#pragma unroll
    for( int c = 0; c < 3; c++ )
    {
        if( u.antiFireflyEnabled == 0 ) // :109-117 (decode -> encode round trip, as the shader does)
        {
            storeU32( a.dst[ c ], px, py, encodeE5( decodeE5( tile[ c ][ cy ][ cx ] ) ) );
            continue;
        }
        float3 smin = make_float3( 1000000.0f, 1000000.0f, 1000000.0f ), smax = make_float3( -1000000.0f, -1000000.0f, -1000000.0f );
#pragma unroll
        for( int j = -R; j <= R; j++ )
        {
#pragma unroll
            for( int i = -R; i <= R; i++ )
            {
                if( i == 0 && j == 0 ) continue;
                const float3 o = decodeE5( tile[ c ][ cy + j ][ cx + i ] );
                smin = make_float3( fminf( smin.x, o.x ), fminf( smin.y, o.y ), fminf( smin.z, o.z ) );
                smax = make_float3( fmaxf( smax.x, o.x ), fmaxf( smax.y, o.y ), fmaxf( smax.z, o.z ) );
            }
        }
        float3 v = decodeE5( tile[ c ][ cy ][ cx ] );
        v        = make_float3( clampf( v.x, smin.x, smax.x ), clampf( v.y, smin.y, smax.y ), clampf( v.z, smin.z, smax.z ) );
        storeU32( a.dst[ c ], px, py, encodeE5( v ) );
    }
}

// ------------------------------------------------------------------------------------------------------------
// V: variance estimation
// ------------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__( kBlockW* kBlockH ) k_variance( const __grid_constant__ VarianceArgs a, const __grid_constant__ Uniform u )
{
    const int px = blockIdx.x * kBlockW + threadIdx.x;
    const int py = u.yBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( px >= u.W || py >= u.yEnd ) return;

    const float pixDepth = fetchR16F( a.depth, px, py );
    float3      color    = fetchE5( a.color, px, py );

    if( pixDepth < 0.0f || pixDepth > 10000.0f ) // :143-147
    {
        storeRGBA16F( a.out, px, py, make_float4( color.x, color.y, color.z, 0.0f ) );
        return;
    }
    const float  historyLength   = fetchRGBA16F( a.histLen, px, py ).x; // :149-150
    const float2 temporalMoments = fetchRG16F( a.moments, px, py );

    if( historyLength > 4.0f ) // :153-161
    {
        float temporalVariance = fmaxf( 0.0f, subr( temporalMoments.y, mulr( temporalMoments.x, temporalMoments.x ) ) );
        temporalVariance       = divr( temporalVariance, fmaxf( 1.0f, subr( historyLength, 4.0f ) ) );
        storeRGBA16F( a.out, px, py, make_float4( color.x, color.y, color.z, temporalVariance ) );
        return;
    }

    const float  l  = luminance( color ); // :163-166
    float        m1 = addr( l, temporalMoments.x ), m2 = addr( mulr( l, l ), temporalMoments.y );
    float        weightSum = 1.0f;
    const float3 n = decodeNormal( a.lut, fetchU32( a.normal, px, py ) );

#pragma unroll
    for( int yy = -1; yy <= 1; yy++ ) // :168-196; out-of-image and cross-separator neighbours participate
    {
#pragma unroll
        for( int xx = -1; xx <= 1; xx++ )
        {
            if( xx == 0 && yy == 0 ) continue;
            const int    qx = px + xx, qy = py + yy;
            const float3 colorQ = fetchE5( a.color, qx, qy );
            const float  lQ     = luminance( colorQ );
            const float  depthQ = fetchR16F( a.depth, qx, qy );
            const float3 nQ     = decodeNormal( a.lut, fetchU32( a.normal, qx, qy ) );
            const float  wZ     = expf( mulr( -fabsf( subr( pixDepth, depthQ ) ), safePositiveRcp( fmaxf( pixDepth, depthQ ) ) ) );
            const float  wN     = powf( fmaxf( 0.0f, dot3r( n, nQ ) ), 128.0f );
            const float  w      = mulr( wZ, wN );
            if( !isnan( w ) )
            {
                m1        = addr( m1, mulr( lQ, w ) );
                m2        = addr( m2, mulr( mulr( lQ, lQ ), w ) );
                color     = color + colorQ * w;
                weightSum = addr( weightSum, w );
            }
        }
    }
    const float inv = safePositiveRcp( weightSum ); // :198-203
    m1    = mulr( m1, inv );
    m2    = mulr( m2, inv );
    color = color * inv;
    const float spatialVariance = mulr( fmaxf( 0.0f, subr( m2, mulr( m1, m1 ) ) ), addr( 1.0f, mulr( 2.0f, subr( 1.0f, divr( historyLength, 4.0f ) ) ) ) );
    storeRGBA16F( a.out, px, py, make_float4( color.x, color.y, color.z, spatialVariance ) );
}

// ------------------------------------------------------------------------------------------------------------
// A: edge-stopping a-trous wavelet filter (3x3 at stride 1<<iteration) over diffuse+variance, specular, indirect.
// ITER0 additionally computes/stores the 3x3-Gaussian-prefiltered luminance sigma; LAST composes PreFinal.
// ------------------------------------------------------------------------------------------------------------

template< bool ITER0, bool LAST >
__global__ void __launch_bounds__( kBlockW* kBlockH ) k_atrous( const __grid_constant__ AtrousArgs a, const __grid_constant__ Uniform u )
{
    const int px = blockIdx.x * kBlockW + threadIdx.x;
    const int py = u.yBegin + blockIdx.y * kBlockH + threadIdx.y;
    if( px >= u.W || py >= u.yEnd ) return;

    const Area chRenderArea = checkerboardedRenderArea( u, px );
    const int  step         = ITER0 ? 1 : ( 1 << a.iteration );

    const float4 cv = fetchRGBA16F( a.srcDiff, px, py );
    float3 outDiff     = make_float3( cv.x, cv.y, cv.z );
    float  outVariance = cv.w;
    float3 outSpec     = fetchE5( a.srcSpec, px, py );
    float3 outIndir    = fetchE5( a.srcIndir, px, py );

    const float depth   = fetchR16F( a.depth, px, py );
    const bool  invalid = depth < 0.0f || depth > 10000.0f;

    if( invalid )
    {
        if( ITER0 ) // Iter0.comp:208-216 (the prefiltered variance image keeps its stale value)
        {
            outVariance = 0.0f;
            outDiff = outSpec = outIndir = make_float3( 0.0f, 0.0f, 0.0f );
        }
        // CmSVGFAtrous.comp:96-99: centre passes through
    }
    else
    {
        const float  gradDepth = fetchR16F( a.depthGrad, px, py );
        const float3 normal    = decodeNormal( a.lut, fetchU32( a.normal, px, py ) );
        const float  l         = luminance( outDiff );

        float sigmaL;
        if( ITER0 ) // Iter0.comp:158-187
        {
            float r = 0.0f;
#pragma unroll
            for( int yy = -1; yy <= 1; yy++ )
#pragma unroll
                for( int xx = -1; xx <= 1; xx++ )
                {
                    const float v = ( xx == 0 && yy == 0 ) ? cv.w : fetchRGBA16F( a.srcDiff, px + xx, py + yy ).w;
                    const float w = ( xx == 0 ? 0.5f : 0.25f ) * ( yy == 0 ? 0.5f : 0.25f ); // {1/4,1/8,1/16}
                    r             = addr( r, mulr( v, w ) );
                }
            r = __fsqrt_rn( fmaxf( r, 0.0f ) );
            storeR16F( a.filteredVariance, px, py, r );
            sigmaL = r;
        }
        else
        {
            sigmaL = fetchR16F( a.filteredVariance, px, py ); // CmSVGFAtrous.comp:65-68
        }
        const float wLumMultiplier = divr( 1.0f, addr( mulr( 4.0f, sigmaL ), 0.00001f ) );

        const float roughness        = fetchRG8( a.metRough, px, py ).y;
        const float wRoughMultiplier = clampf( subr( mulr( roughness, 30.0f ), ( float )a.iteration ), 0.0f, 1.0f );

        const float historyLengthSpec = fetchRGBA16F( a.histLen, px, py ).z;
        const float normalWeightScale = clampf( divr( historyLengthSpec, 8.0f ), 0.0f, 1.0f );
        const float alpha             = mulr( roughness, roughness );
        float       normalWeightSpec  = fmaxf( 0.01f, subr( divr( 2.0f, addr( mulr( alpha, alpha ), 1e-4f ) ), 2.0f ) ); // BRDF.h:27-29
        normalWeightSpec              = mulr( clampf( normalWeightSpec, 8.0f, 1024.0f ), normalWeightScale );

        float weightSum = 1.0f, weightSumSpec = 1.0f, weightSumIndir = 1.0f;

#pragma unroll
        for( int yy = -1; yy <= 1; yy++ )
        {
#pragma unroll
            for( int xx = -1; xx <= 1; xx++ )
            {
                if( xx == 0 && yy == 0 ) continue;
                const int qx = px + xx * step, qy = py + yy * step;

                const float  roughnessQ = fetchRG8( a.metRough, qx, qy ).y;
                const float3 normalQ    = decodeNormal( a.lut, fetchU32( a.normal, qx, qy ) );
                const float  depthQ     = fetchR16F( a.depth, qx, qy );
                const float3 specQ      = fetchE5( a.srcSpec, qx, qy );
                const float4 cvQ        = fetchRGBA16F( a.srcDiff, qx, qy );
                const float3 diffQ      = make_float3( cvQ.x, cvQ.y, cvQ.z );
                const float  lQ         = luminance( diffQ );
                const float  nN         = fmaxf( 0.0f, dot3r( normal, normalQ ) );
                const float3 indirQ     = fetchE5( a.srcIndir, qx, qy );

                const float wZ = divr( fabsf( subr( depth, depthQ ) ), fmaxf( mulr( gradDepth, ( float )( abs( xx ) + abs( yy ) ) ), 0.01f ) );
                const float wN = powf( nN, 128.0f );
                const float wL = mulr( fabsf( subr( l, lQ ) ), wLumMultiplier );

                float wR = mulr( fmaxf( 0.0f, subr( 1.0f, mulr( 10.0f, fabsf( subr( roughness, roughnessQ ) ) ) ) ), wRoughMultiplier );
                if( normalWeightSpec > 0.0f ) wR = mulr( wR, powf( nN, normalWeightSpec ) );

                const float waveletW = ( yy == 0 ? 1.0f : 0.5f ) * ( xx == 0 ? 1.0f : 0.5f );
                const float isInside = testPixInRenderArea( qx, qy, chRenderArea ) ? 1.0f : 0.0f;

                const float wBase = mulr( mulr( mulr( expf( mulr( -wZ, wZ ) ), wN ), waveletW ), isInside );
                const float wDiff = mulr( wBase, expf( -wL ) );
                const float wSpec = mulr( wBase, wR );

                outDiff     = outDiff + diffQ * wDiff;
                outSpec     = outSpec + specQ * wSpec;
                outIndir    = outIndir + indirQ * wBase;
                outVariance = addr( outVariance, mulr( mulr( cvQ.w, wDiff ), wDiff ) );
                weightSum      = addr( weightSum, wDiff );
                weightSumSpec  = addr( weightSumSpec, wSpec );
                weightSumIndir = addr( weightSumIndir, wBase );
            }
        }
        const float inv = divr( 1.0f, weightSum ), invSpec = divr( 1.0f, weightSumSpec ), invIndir = divr( 1.0f, weightSumIndir );
        outDiff     = outDiff * inv;
        outVariance = mulr( outVariance, mulr( inv, inv ) );
        outSpec     = outSpec * invSpec;
        outIndir    = outIndir * invIndir;
    }

    if( !LAST )
    {
        storeRGBA16F( a.dstDiff, px, py, make_float4( outDiff.x, outDiff.y, outDiff.z, outVariance ) );
        storeU32( a.dstSpec, px, py, encodeE5( outSpec ) );
        storeU32( a.dstIndir, px, py, encodeE5( outIndir ) );
        return;
    }

    // composition, CmSVGFAtrous.comp:278-348
    const int2   reg        = regularPixFromCheckerboardPix( u, px, py );
    float3       albedo     = unpackUfloat111110( fetchU32( a.albedo, reg.x, reg.y ) );
    const float4 tp4        = fetchRGBA16F( a.throughput, px, py );
    const float3 throughput = make_float3( tp4.x, tp4.y, tp4.z );
    if( ( u.debugShowFlags & ( 1u << 9 ) ) != 0 ) albedo = make_float3( 1.0f, 1.0f, 1.0f );

    float3 illuminated;
    if( fetchU8( a.isSky, px, py ) == 0u )
    {
        const float2 mr        = fetchRG8( a.metRough, px, py );
        const float  metallic  = mr.x, roughness = mr.y;
        const float3 specular  = mix3r( outSpec, outIndir, smoothstepr( 0.5f, 0.5f + 0.25f, roughness ) );
        const float3 roS       = mix3r( make_float3( 0.04f, 0.04f, 0.04f ), albedo, metallic ); // BRDF.h:49-53
        const float3 roD       = albedo * subr( 1.0f, metallic );
        illuminated            = ( outDiff + outIndir ) * roD + specular * roS;
        if( ( u.debugShowFlags & ( 1u << 5 ) ) != 0 ) illuminated = outDiff * albedo;
        else if( ( u.debugShowFlags & ( 1u << 6 ) ) != 0 ) illuminated = specular * roS;
        else if( ( u.debugShowFlags & ( 1u << 7 ) ) != 0 ) illuminated = outIndir;
        illuminated = illuminated * throughput;
        const float la = luminance( albedo ); // BRDF.h:57-64
        float       ambient = 1.0f;
        if( !( la > 0.02f ) ) { const float q = divr( subr( la, 0.02f ), 0.02f ); ambient = subr( 1.0f, mulr( q, q ) ); }
        illuminated = illuminated * ambient;
    }
    else
    {
        illuminated = albedo * throughput;
    }
    storeR16F( a.histogramInput, reg.x, reg.y, fmaxf( 0.0f, luminance( illuminated ) ) );
    if( ( float )reg.x < mulr( u.lightmapScreenCoverage, u.renderWidth ) ) illuminated = albedo * throughput; // classicShading
    storeU32( a.preFinal, reg.x, reg.y, packUfloat111110( make_float3( fmaxf( 0.0f, illuminated.x ), fmaxf( 0.0f, illuminated.y ), fmaxf( 0.0f, illuminated.z ) ) ) );
}

} // namespace asvgf
