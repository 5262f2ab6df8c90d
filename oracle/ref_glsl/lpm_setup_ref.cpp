// lpm_setup_ref.cpp -- TEST INFRASTRUCTURE ONLY (part of oracle/_ref/libasvgf_ref.so).
//
// The LPM tonemapper of CmPrepareFinal.comp reads a 24 x uvec4 control block that RTGL1 computes on the HOST with AMD's own
// setup code (Source/ImageComposition.cpp:339-400: ffx_a.h in A_CPU mode + ffx_lpm.h + LpmSetupCustom.inl).  This file compiles
// those three reference headers where they lie under /root/reference (-I <reference>/Source/Shaders/LPM) and calls LpmSetup with the
// arguments ImageComposition::SetupLpmParams passes, so that the tests feed the shader, the oracle and the CUDA kernel the block
// the reference would have produced.  Only the output callback below is written here (it is the reference's LpmSetupOut).
#include <cassert>
#include <cstdint>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define A_CPU 1
#include "ffx_a.h"

namespace
{
void LpmSetupOut( AU1, inAU4 ) { assert( 0 ); } // the context-less form must not be reached (ImageComposition.cpp:343-347)
void LpmSetupOut( void* pContext, AU1 i, const inAU4 value )
{
    AU1* ctl = static_cast< AU1* >( pContext );
    assert( i < 24 );
    for( int k = 0; k < 4; k++ ) ctl[ i * 4 + k ] = value[ k ];
}
}

#include "ffx_lpm.h"
#include "LpmSetupCustom.inl"

// ImageComposition.cpp:376-394: Rec.709 in, Rec.709 out, hdrMax 256, exposure 8 stops, contrast 0.1, no shoulder
extern "C" void asvgfRefLpmSetup( uint32_t* control96, const float* saturation3, const float* crosstalk3 )
{
    for( int i = 0; i < 96; i++ ) control96[ i ] = 0u;
    varAF3( saturation ) = initAF3( saturation3[ 0 ], saturation3[ 1 ], saturation3[ 2 ] );
    varAF3( crosstalk )  = initAF3( crosstalk3[ 0 ], crosstalk3[ 1 ], crosstalk3[ 2 ] );
    LpmSetup( control96, false, LPM_CONFIG_709_709, LPM_COLORS_709_709, 0.0f, 256.0f, 8.0f, 0.1f, 1.0f, saturation, crosstalk );
}
