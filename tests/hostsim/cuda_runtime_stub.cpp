// cuda_runtime_stub.cpp -- TEST INFRASTRUCTURE ONLY (tests/hostsim).  The few CUDA runtime entry points the host side of
// libasvgf (raytracedgl1_b200/csrc/asvgf.cu) calls, as synchronous host functions: "device memory" is the heap, a stream
// executes at enqueue time, events and waits are therefore no-ops, HOSTSIM_DEVICES (default 2) pretend devices share the one
// address space.  Linked INSTEAD of libcudart into tests/hostsim/libasvgf_hostsim.so, never into the product.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

namespace
{
thread_local int g_device = 0;
int deviceCount()
{
    const char* e = std::getenv( "HOSTSIM_DEVICES" );
    const int   n = e ? std::atoi( e ) : 2;
    return n < 1 ? 1 : n;
}
}

extern "C" {

// what raytracedgl1_b200.denoiser looks for to REFUSE this library outside the test suite (ASVGF_LIB cannot route the product to the CPU)
int asvgfHostBuildMarker( void ) { return 1; }

cudaError_t cudaGetDeviceCount( int* n ) { *n = deviceCount(); return cudaSuccess; }
cudaError_t cudaGetDevice( int* d ) { *d = g_device; return cudaSuccess; }
cudaError_t cudaSetDevice( int d ) { if( d < 0 || d >= deviceCount() ) return cudaErrorInvalidDevice; g_device = d; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute( int* v, cudaDeviceAttr attr, int ) { *v = attr == cudaDevAttrMultiProcessorCount ? 4 : 0; return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer( int* can, int, int ) { *can = 1; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess( int, unsigned ) { return cudaSuccess; }
cudaError_t cudaGetLastError( void ) { return cudaSuccess; }
const char* cudaGetErrorString( cudaError_t e ) { return e == cudaSuccess ? "no error" : "hostsim: CUDA runtime stub error"; }

cudaError_t cudaMalloc( void** p, size_t bytes )
{
    *p = nullptr;
    if( posix_memalign( p, 256, bytes ? bytes : 1 ) != 0 ) return cudaErrorMemoryAllocation;
    return cudaSuccess;
}
cudaError_t cudaFree( void* p ) { free( p ); return cudaSuccess; }
cudaError_t cudaMemset( void* p, int v, size_t n ) { std::memset( p, v, n ); return cudaSuccess; }
cudaError_t cudaMemsetAsync( void* p, int v, size_t n, cudaStream_t ) { std::memset( p, v, n ); return cudaSuccess; }
cudaError_t cudaMemcpy( void* d, const void* s, size_t n, cudaMemcpyKind ) { std::memmove( d, s, n ); return cudaSuccess; }
cudaError_t cudaMemcpyAsync( void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t ) { std::memmove( d, s, n ); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync( void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t )
{
    if( width > dpitch || width > spitch ) return cudaErrorInvalidPitchValue;
    for( size_t y = 0; y < height; y++ ) std::memmove( ( char* )d + y * dpitch, ( const char* )s + y * spitch, width );
    return cudaSuccess;
}

cudaError_t cudaStreamCreateWithFlags( cudaStream_t* s, unsigned ) { *s = ( cudaStream_t )std::malloc( 8 ); return *s ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaStreamDestroy( cudaStream_t s ) { std::free( s ); return cudaSuccess; }
cudaError_t cudaStreamSynchronize( cudaStream_t ) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent( cudaStream_t, cudaEvent_t, unsigned ) { return cudaSuccess; }
cudaError_t cudaEventCreate( cudaEvent_t* e ) { *e = ( cudaEvent_t )std::malloc( 8 ); return *e ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaEventCreateWithFlags( cudaEvent_t* e, unsigned ) { return cudaEventCreate( e ); }
cudaError_t cudaEventDestroy( cudaEvent_t e ) { std::free( e ); return cudaSuccess; }
cudaError_t cudaEventRecord( cudaEvent_t, cudaStream_t ) { return cudaSuccess; }
cudaError_t cudaEventSynchronize( cudaEvent_t ) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime( float* ms, cudaEvent_t, cudaEvent_t ) { *ms = 0.0f; return cudaSuccess; }
cudaError_t cudaFuncSetAttribute( const void*, cudaFuncAttribute, int ) { return cudaSuccess; }

// the Vulkan / driver interop has no host equivalent
cudaError_t cudaImportExternalMemory( cudaExternalMemory_t*, const cudaExternalMemoryHandleDesc* ) { return cudaErrorNotSupported; }
cudaError_t cudaExternalMemoryGetMappedBuffer( void**, cudaExternalMemory_t, const cudaExternalMemoryBufferDesc* ) { return cudaErrorNotSupported; }
cudaError_t cudaDestroyExternalMemory( cudaExternalMemory_t ) { return cudaSuccess; }
cudaError_t cudaImportExternalSemaphore( cudaExternalSemaphore_t*, const cudaExternalSemaphoreHandleDesc* ) { return cudaErrorNotSupported; }
cudaError_t cudaDestroyExternalSemaphore( cudaExternalSemaphore_t ) { return cudaSuccess; }
cudaError_t cudaWaitExternalSemaphoresAsync_v2( const cudaExternalSemaphore_t*, const cudaExternalSemaphoreWaitParams*, unsigned, cudaStream_t ) { return cudaErrorNotSupported; }
cudaError_t cudaSignalExternalSemaphoresAsync_v2( const cudaExternalSemaphore_t*, const cudaExternalSemaphoreSignalParams*, unsigned, cudaStream_t ) { return cudaErrorNotSupported; }

} // extern "C"
