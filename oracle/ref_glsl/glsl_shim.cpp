/*
 * glsl_shim.cpp -- work-group execution for oracle/_ref (see glsl_shim.h): TEST INFRASTRUCTURE ONLY.
 *
 * vkCmdDispatch(groupsX, groupsY, 1) of a compute shader with local size (localX, localY, 1):
 *   - work groups are independent -> spread over OpenMP threads;
 *   - the invocations of ONE work group share `shared` variables (thread_local statics of the generated shader class) and
 *     may meet in barrier(): each invocation is a ucontext fiber on the group's OS thread, barrier() switches back to
 *     the scheduler, which resumes every fiber in turn until all have passed the barrier (or returned).
 *   Shaders without barrier() run their invocations as plain calls.
 */
#include "glsl_shim.h"

#ifdef _OPENMP
    #include <omp.h>
#endif

namespace glslshim
{

thread_local const DispatchContext* g_ctx = nullptr;
thread_local Builtins               g_builtins;

namespace
{

constexpr size_t kFiberStack = 256 * 1024;

struct Fiber
{
    ucontext_t ctx;
    Builtins   builtins;
    bool       finished = false;
    char*      stack    = nullptr;
};

struct GroupScheduler
{
    ucontext_t           main;
    std::vector< Fiber > fibers;
    Fiber*               current = nullptr;
    InvocationFn         fn      = nullptr;
};
thread_local GroupScheduler* g_sched = nullptr;

void fiberEntry()
{
    GroupScheduler* s = g_sched;
    s->fn();
    s->current->finished = true;
    swapcontext( &s->current->ctx, &s->main ); // never resumed
}

} // namespace

void barrier()
{
    GroupScheduler* s = g_sched;
    if( !s || !s->current ) return; // shader dispatched without fibers: a single pass is already sequential
    Fiber* f = s->current;
    swapcontext( &f->ctx, &s->main );
    g_builtins = f->builtins; // resumed: this fiber's built-ins are current again
}

void dispatch( const DispatchContext& ctx, InvocationFn fn, uint32_t localX, uint32_t localY, uint32_t groupsX, uint32_t groupsY, bool usesBarrier )
{
    const long nGroups = ( long )groupsX * ( long )groupsY;
    const uint32_t nLocal = localX * localY;
#pragma omp parallel
    {
        g_ctx = &ctx;
        GroupScheduler sched;
        sched.fn = fn;
        if( usesBarrier )
        {
            sched.fibers.resize( nLocal );
            for( Fiber& f : sched.fibers ) f.stack = ( char* )std::malloc( kFiberStack );
        }
        g_sched = &sched;
#pragma omp for schedule( dynamic, 8 )
        for( long g = 0; g < nGroups; g++ )
        {
            Builtins b;
            b.gl_NumWorkGroups = uvec3( groupsX, groupsY, 1u );
            b.gl_WorkGroupSize = uvec3( localX, localY, 1u );
            b.gl_WorkGroupID   = uvec3( ( uint32_t )( g % groupsX ), ( uint32_t )( g / groupsX ), 0u );
            if( !usesBarrier )
            {
                for( uint32_t l = 0; l < nLocal; l++ )
                {
                    b.gl_LocalInvocationID    = uvec3( l % localX, l / localX, 0u );
                    b.gl_LocalInvocationIndex = l;
                    b.gl_GlobalInvocationID   = b.gl_WorkGroupID * b.gl_WorkGroupSize + b.gl_LocalInvocationID;
                    g_builtins                = b;
                    fn();
                }
                continue;
            }
            for( uint32_t l = 0; l < nLocal; l++ )
            {
                Fiber& f                  = sched.fibers[ l ];
                b.gl_LocalInvocationID    = uvec3( l % localX, l / localX, 0u );
                b.gl_LocalInvocationIndex = l;
                b.gl_GlobalInvocationID   = b.gl_WorkGroupID * b.gl_WorkGroupSize + b.gl_LocalInvocationID;
                f.builtins                = b;
                f.finished                = false;
                getcontext( &f.ctx );
                f.ctx.uc_stack.ss_sp   = f.stack;
                f.ctx.uc_stack.ss_size = kFiberStack;
                f.ctx.uc_link          = nullptr;
                makecontext( &f.ctx, fiberEntry, 0 );
            }
            bool anyAlive = true;
            while( anyAlive ) // one sweep = run every live invocation up to its next barrier()
            {
                anyAlive = false;
                for( Fiber& f : sched.fibers )
                {
                    if( f.finished ) continue;
                    sched.current = &f;
                    g_builtins    = f.builtins;
                    swapcontext( &sched.main, &f.ctx );
                    if( !f.finished ) anyAlive = true;
                }
            }
            sched.current = nullptr;
        }
        for( Fiber& f : sched.fibers ) std::free( f.stack );
        g_sched = nullptr;
        g_ctx   = nullptr;
    }
}

} // namespace glslshim
