// cuda_host_shim.h -- TEST INFRASTRUCTURE ONLY (tests/hostsim): lets g++ compile the product's kernel SOURCE
// (raytracedgl1_b200/csrc/*.cuh) for the host, so that the `-m "not gpu"` suite can execute the kernels' logic thread by thread
// on the CPU and compare it with the oracle where no GPU exists.  It is not a CPU path of the product: nothing under
// raytracedgl1_b200/ includes it, libasvgf.so does not contain it, and the numbers it produces are never shipped or timed.
//
// What is emulated: the launch geometry (threadIdx / blockIdx as thread-local variables; a grid is executed block after block,
// a block thread after thread -- or, as soon as a kernel reaches a barrier, by one host thread per CUDA thread), shared memory
// (static storage: one block runs at a time), __syncthreads / __syncthreads_or, the correctly rounded intrinsics (__fmul_rn ...
// are IEEE single operations: the host's, compiled with -ffp-contract=off), conversions, __ldg, atomics.
// What is NOT: warp operations, TMA / mbarrier and the MUFU approximations -- the few places that use them carry an
// `#ifdef ASVGF_HOSTSIM` spelling in C (libm instead of MUFU: colour values agree with the GPU's to rounding, integer outputs
// exactly), and what the GPU alone can show (the instructions themselves, timing, races) stays with the `-m gpu` tests.
#pragma once

#include <cuda_runtime.h> // vector types, dim3; under g++ the __device__ / __global__ / __launch_bounds__ qualifiers expand to nothing
#include <cuda_fp16.h>    // __half / __half2 with their host implementations

#include <cmath>
#include <condition_variable>
#include <csetjmp>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
    #error "cuda_host_shim.h is for host compilers only"
#endif

#ifndef __grid_constant__
    #define __grid_constant__
#endif
#ifndef __launch_bounds__
    #define __launch_bounds__( ... )
#endif
using std::isnan;

namespace hostsim
{
inline thread_local uint3 tIdx = { 0, 0, 0 }, bIdx = { 0, 0, 0 };
inline thread_local dim3  bDim( 1, 1, 1 ), gDim( 1, 1, 1 );

// barrier of the block that is running in threaded mode; exited threads no longer count (as on the GPU)
struct BlockBarrier
{
    std::mutex              m;
    std::condition_variable cv;
    int                     live = 0, waiting = 0, acc = 0, result = 0;
    unsigned                gen = 0;
    void release() { result = acc; acc = 0; waiting = 0; gen++; cv.notify_all(); }
    int  arrive( int pred )
    {
        std::unique_lock< std::mutex > lock( m );
        acc |= pred != 0;
        if( ++waiting == live ) { release(); return result; }
        const unsigned g = gen;
        cv.wait( lock, [ & ] { return gen != g; } );
        return result;
    }
    template< typename F > bool arriveLast( F hook ) // the last thread to arrive runs `hook` before anyone continues
    {
        std::unique_lock< std::mutex > lock( m );
        if( ++waiting == live ) { hook(); release(); return true; }
        const unsigned g = gen;
        cv.wait( lock, [ & ] { return gen != g; } );
        return false;
    }
    void leave()
    {
        std::unique_lock< std::mutex > lock( m );
        if( --live > 0 && waiting == live ) release();
    }
};
// a barrier reached in sequential mode jumps back into the launcher (kernels hold no objects with destructors), which then
// runs this block, and the rest of the grid, with one host thread per CUDA thread
inline thread_local jmp_buf* tRestart = nullptr;
inline thread_local BlockBarrier* tBarrier = nullptr;
inline uint8_t*                   gDynamicSmem = nullptr;
inline uint8_t*                   dynamicSmem() { return gDynamicSmem; }
inline std::mutex&                launchMutex() { static std::mutex m; return m; }
inline int sync( int pred )
{
    if( !tBarrier ) longjmp( *tRestart, 1 );
    return tBarrier->arrive( pred );
}

template< typename... KArgs, typename... Args >
void launch( void ( *kernel )( KArgs... ), dim3 grid, dim3 block, size_t smemBytes, const Args&... args )
{
    std::lock_guard< std::mutex > lock( launchMutex() ); // static "shared memory": one grid, one block at a time
    void* smem = nullptr;
    if( smemBytes && posix_memalign( &smem, 128, smemBytes ) != 0 ) std::abort();
    gDynamicSmem = ( uint8_t* )smem;
    const unsigned nThreads = block.x * block.y * block.z;
    auto setIdx = [ & ]( unsigned t, uint3 b ) {
        tIdx = { t % block.x, ( t / block.x ) % block.y, t / ( block.x * block.y ) };
        bIdx = b; bDim = block; gDim = grid;
    };
    volatile bool threaded = false;
    for( unsigned bz = 0; bz < grid.z; bz++ )
        for( unsigned by = 0; by < grid.y; by++ )
            for( unsigned bx = 0; bx < grid.x; bx++ )
            {
                const uint3 b = { bx, by, bz };
                if( !threaded )
                {
                    jmp_buf restart;
                    tRestart = &restart;
                    if( setjmp( restart ) == 0 )
                    {
                        for( volatile unsigned t = 0; t < nThreads; t = t + 1 ) { setIdx( t, b ); kernel( args... ); }
                        continue;
                    }
                    threaded = true; // (what the block's first threads wrote before the barrier is written again below)
                }
                // threaded mode, from this block to the end of the grid: one host thread per CUDA thread, created once; the
                // blocks still run one after the other (an all-thread barrier between them: static "shared memory")
                const unsigned first = ( bz * grid.y + by ) * grid.x + bx, nBlocks = grid.x * grid.y * grid.z;
                BlockBarrier   bar, between;
                bar.live = between.live = ( int )nThreads;
                std::vector< std::thread > pool;
                pool.reserve( nThreads );
                for( unsigned t = 0; t < nThreads; t++ )
                    pool.emplace_back( [ &, t ] {
                        tBarrier = &bar;
                        for( unsigned k = first; k < nBlocks; k++ )
                        {
                            setIdx( t, uint3{ k % grid.x, ( k / grid.x ) % grid.y, k / ( grid.x * grid.y ) } );
                            kernel( args... );
                            bar.leave();
                            between.arriveLast( [ & ] { bar.live = ( int )nThreads; bar.waiting = 0; bar.acc = 0; } ); // next block
                        }
                    } );
                for( auto& th : pool ) th.join();
                goto done;
            }
done:
    gDynamicSmem = nullptr;
    free( smem );
}
}
#define threadIdx ( ::hostsim::tIdx )
#define blockIdx ( ::hostsim::bIdx )
#define blockDim ( ::hostsim::bDim )
#define gridDim ( ::hostsim::gDim )

inline float    __fmul_rn( float a, float b ) { return a * b; }
inline float    __fadd_rn( float a, float b ) { return a + b; }
inline float    __fsub_rn( float a, float b ) { return a - b; }
inline float    __fdiv_rn( float a, float b ) { return a / b; }
inline float    __fsqrt_rn( float a ) { return std::sqrt( a ); }
inline uint32_t __float_as_uint( float f ) { uint32_t u; std::memcpy( &u, &f, 4 ); return u; }
inline float    __uint_as_float( uint32_t u ) { float f; std::memcpy( &f, &u, 4 ); return f; }
inline int      __float2int_rz( float f ) { return f != f ? 0 : f >= 2147483648.0f ? 2147483647 : f <= -2147483648.0f ? ( -2147483647 - 1 ) : ( int )f; }
inline int      __float2int_rn( float f ) { return f != f ? 0 : f >= 2147483648.0f ? 2147483647 : f <= -2147483648.0f ? ( -2147483647 - 1 ) : ( int )std::nearbyint( f ); }
inline uint32_t __float2uint_rn( float f ) { return !( f > 0.0f ) ? 0u : f >= 4294967296.0f ? 0xffffffffu : ( uint32_t )std::nearbyint( f ); }
inline uint32_t __byte_perm( uint32_t a, uint32_t b, uint32_t s )
{
    const uint64_t v = ( ( uint64_t )b << 32 ) | a;
    uint32_t       r = 0;
    for( int i = 0; i < 4; i++ ) r |= ( uint32_t )( ( v >> ( 8 * ( ( s >> ( 4 * i ) ) & 7u ) ) ) & 0xffu ) << ( 8 * i );
    return r;
}
inline int      __ffs( int v ) { return v ? __builtin_ctz( ( unsigned )v ) + 1 : 0; }
inline int      __popc( uint32_t v ) { return __builtin_popcount( v ); }
template< typename T > inline T __ldg( const T* p ) { return *p; }
inline uint32_t atomicOr( uint32_t* p, uint32_t v ) { return __atomic_fetch_or( p, v, __ATOMIC_RELAXED ); }
inline uint32_t atomicAdd( uint32_t* p, uint32_t v ) { return __atomic_fetch_add( p, v, __ATOMIC_RELAXED ); }
inline int      max( int a, int b ) { return a > b ? a : b; }
inline int      min( int a, int b ) { return a < b ? a : b; }
inline void __syncthreads() { ::hostsim::sync( 0 ); }
inline int  __syncthreads_or( int pred ) { return ::hostsim::sync( pred ); }
// __shared__ variables become static storage (cuda_runtime.h made the qualifier expand to nothing)
#undef __shared__
#define __shared__ static
// cuda_runtime.h declares the typed wrapper for nvcc only
template< typename... KArgs > inline cudaError_t cudaFuncSetAttribute( void ( *kernel )( KArgs... ), cudaFuncAttribute attr, int value )
{
    return ::cudaFuncSetAttribute( reinterpret_cast< const void* >( kernel ), attr, value );
}
