/*
 * asvgf_oracle.h -- C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a scalar fp32 restatement of RTGL1's denoiser shaders
 * (see asvgf_oracle.cpp for the file:line map).  It may be loaded only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs -- never by the product path (libasvgf.so), which has
 * no CPU fallback.
 *
 * PARITY PINNED BY THE REFERENCE'S OWN SHADER TEXT: RTGL1 ships no golden vectors and no Vulkan runs here, but
 * oracle/_ref compiles the unmodified Source/Shaders/Cm*.comp as C++ (oracle/ref_glsl/build_ref.py) and this oracle is
 * bit-identical to it on every dispatch (tests/test_ref_pins_oracle.py, tests/golden/asvgf_ref_96x54.npz).  The
 * gradient-sample stage (RtGradients.rgen, ray generation) is pinned by restatement and known-answer tests only.
 */
#ifndef ASVGF_ORACLE_H_
#define ASVGF_ORACLE_H_

#include "../include/asvgf_layout.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct AsvgfOracle_T* AsvgfOracle;

AsvgfOracle asvgfOracleCreate( uint32_t renderWidth, uint32_t renderHeight, uint32_t atrousIterations );
void        asvgfOracleDestroy( AsvgfOracle o );
void        asvgfOracleSetThreads( int threads ); /* <=0: all cores */
int         asvgfOracleGetThreads( void );

/* host pointer of an image, cur/prev swap applied (Framebuffers.cpp:35-50); rows are tightly packed */
int asvgfOracleGetImage( AsvgfOracle o, uint32_t imageIndex, uint32_t frameIndex, void** ppHost, uint32_t* pWidth,
                         uint32_t* pHeight, uint32_t* pBytesPerPixel );
void asvgfOracleResetHistory( AsvgfOracle o );

/* pass numbering = AsvgfPass of include/asvgf.h */
void asvgfOracleRunPass( AsvgfOracle o, uint32_t pass, uint32_t frameIndex, const AsvgfGlobalUniform* u );
void asvgfOracleGradientSamples( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u );
void asvgfOracleDenoise( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u );

/* the two dispatches after the denoiser: CmCheckerboard.comp (PreFinal -> Final) and
   CmLuminanceHistogram.comp + CmLuminanceAvg.comp on the oracle's own AsvgfTonemapping state */
void asvgfOracleResolveCheckerboard( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u );
void asvgfOracleLuminanceHistogram( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u );
void asvgfOracleLuminanceAverage( AsvgfOracle o, const AsvgfGlobalUniform* u );
void asvgfOracleCalculateExposure( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u );
AsvgfTonemapping* asvgfOracleGetTonemapping( AsvgfOracle o );
/* CmPrepareFinal.comp (second half): exposure + LPM / Uncharted 2 tonemap of Final, bloom input;
   lpmControl96 = the 24 x uvec4 block of ImageComposition::SetupLpmParams; reads the oracle's own AsvgfTonemapping state */
void asvgfOraclePrepareFinal( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u, const uint32_t* lpmControl96 );
/* CmScatterAccum.comp: temporal accumulation of the volumetric scattering */
void asvgfOracleScatterAccumulate( AsvgfOracle o, uint32_t frameIndex, const AsvgfGlobalUniform* u, float maxHistoryLength );

/* helper primitives exposed for unit tests */
uint32_t asvgfOracleEncodeE5B9G9R9( float r, float g, float b );
void     asvgfOracleDecodeE5B9G9R9( uint32_t packed, float* rgb );
uint32_t asvgfOracleEncodeNormal( float x, float y, float z );
void     asvgfOracleDecodeNormal( uint32_t packed, float* xyz );
uint16_t asvgfOracleFloatToHalf( float f );
float    asvgfOracleHalfToFloat( uint16_t h );
uint32_t asvgfOraclePackUfloat111110( float r, float g, float b );
void     asvgfOracleUnpackUfloat111110( uint32_t packed, float* rgb );
uint8_t  asvgfOracleFloatToUnorm8( float f );
void     asvgfOracleCheckerboardPix( int x, int y, int fullWidth, int* outXY );          /* regular -> checkerboard */
void     asvgfOracleRegularPix( int x, int y, int fullWidth, int* outXY );               /* checkerboard -> regular */
void     asvgfOraclePrevScreenPos( float mx, float my, int x, int y, float renderWidth, float renderHeight,
                                   float* outXY );

#ifdef __cplusplus
}
#endif
#endif
