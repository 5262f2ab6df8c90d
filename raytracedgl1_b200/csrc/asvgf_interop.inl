// asvgf_interop.inl -- the CUDA half of the Vulkan interop (included by asvgf.cu).
//
// In RTGL1 the G-buffer lives in VkImages owned by Framebuffers (Source/Framebuffers.cpp:702-804, dedicated allocations), and
// VulkanDevice::Render orders the denoiser against its producers with the frame's command buffer and fence
// (Source/VulkanDevice.cpp:551-584, 791-796).  A host that keeps that producer exports
//   * the memory of each image (VK_KHR_external_memory_fd, VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT; the image created
//     with VK_IMAGE_TILING_LINEAR, or an optimal-tiled image copied into a linear VkBuffer by vkCmdCopyImageToBuffer), and
//   * one semaphore per direction (VK_KHR_external_semaphore_fd, binary or timeline),
// and this file is what the library does with those file descriptors: import, map as linear device memory (to be handed
// to asvgfBindExternalImage), wait / signal on a CUDA stream.  No Vulkan header is needed on this side of the boundary.
// The Vulkan half (creating the exportable objects) belongs to the host: this image has no Vulkan SDK to build it against.

struct AsvgfExternalMemory_T
{
    cudaExternalMemory_t mem = nullptr;
    int                  device = 0;
    uint64_t             size = 0;
};
struct AsvgfExternalSemaphore_T
{
    cudaExternalSemaphore_t sem = nullptr;
    int                     device = 0;
    bool                    timeline = false;
};

extern "C" {

AsvgfResult asvgfImportExternalMemoryFd( int32_t device, int fd, uint64_t allocationSize, uint32_t dedicated, AsvgfExternalMemory* pMemory )
{
    if( !pMemory || fd < 0 || allocationSize == 0 ) return fail( ASVGF_WRONG_ARGUMENT, "bad fd / size / output pointer" );
    *pMemory = nullptr;
    if( device < 0 ) CUDA_TRY( cudaGetDevice( &device ) );
    DeviceGuard guard( device );
    if( !guard.ok ) return fail( ASVGF_CUDA_ERROR, "cudaSetDevice failed" );
    cudaExternalMemoryHandleDesc desc{};
    desc.type      = cudaExternalMemoryHandleTypeOpaqueFd;
    desc.handle.fd = fd; // ownership passes to CUDA on success (as in vkGetMemoryFdKHR -> import)
    desc.size      = allocationSize;
    desc.flags     = dedicated ? cudaExternalMemoryDedicated : 0u;
    AsvgfExternalMemory_T* m = new( std::nothrow ) AsvgfExternalMemory_T();
    if( !m ) return fail( ASVGF_OUT_OF_MEMORY, "host allocation failed" );
    cudaError_t e = cudaImportExternalMemory( &m->mem, &desc );
    if( e != cudaSuccess )
    {
        delete m;
        cudaGetLastError();
        return cudaFail( e, "cudaImportExternalMemory(opaque fd)" );
    }
    m->device = device; m->size = allocationSize;
    *pMemory = m;
    return ASVGF_SUCCESS;
}

AsvgfResult asvgfExternalMemoryGetBuffer( AsvgfExternalMemory memory, uint64_t offset, uint64_t size, void** ppDevice )
{
    if( !memory || !ppDevice || size == 0 || offset + size > memory->size ) return fail( ASVGF_WRONG_ARGUMENT, "bad range" );
    DeviceGuard guard( memory->device );
    cudaExternalMemoryBufferDesc desc{};
    desc.offset = offset; desc.size = size;
    CUDA_TRY( cudaExternalMemoryGetMappedBuffer( ppDevice, memory->mem, &desc ) );
    return ASVGF_SUCCESS;
}

AsvgfResult asvgfDestroyExternalMemory( AsvgfExternalMemory memory )
{
    if( !memory ) return ASVGF_SUCCESS;
    DeviceGuard guard( memory->device );
    cudaDestroyExternalMemory( memory->mem ); // also releases every buffer mapped from it
    delete memory;
    return ASVGF_SUCCESS;
}

AsvgfResult asvgfImportExternalSemaphoreFd( int32_t device, int fd, uint32_t timeline, AsvgfExternalSemaphore* pSemaphore )
{
    if( !pSemaphore || fd < 0 ) return fail( ASVGF_WRONG_ARGUMENT, "bad fd / output pointer" );
    *pSemaphore = nullptr;
    if( device < 0 ) CUDA_TRY( cudaGetDevice( &device ) );
    DeviceGuard guard( device );
    cudaExternalSemaphoreHandleDesc desc{};
    desc.type      = timeline ? cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd : cudaExternalSemaphoreHandleTypeOpaqueFd;
    desc.handle.fd = fd;
    AsvgfExternalSemaphore_T* s = new( std::nothrow ) AsvgfExternalSemaphore_T();
    if( !s ) return fail( ASVGF_OUT_OF_MEMORY, "host allocation failed" );
    cudaError_t e = cudaImportExternalSemaphore( &s->sem, &desc );
    if( e != cudaSuccess )
    {
        delete s;
        cudaGetLastError();
        return cudaFail( e, "cudaImportExternalSemaphore(fd)" );
    }
    s->device = device; s->timeline = timeline != 0;
    *pSemaphore = s;
    return ASVGF_SUCCESS;
}

// the producer's "G-buffer written" signal: everything enqueued on `stream` afterwards waits for it
AsvgfResult asvgfWaitExternalSemaphore( AsvgfExternalSemaphore semaphore, AsvgfStream stream, uint64_t value )
{
    if( !semaphore ) return fail( ASVGF_WRONG_ARGUMENT, "null semaphore" );
    DeviceGuard guard( semaphore->device );
    cudaExternalSemaphoreWaitParams p{};
    p.params.fence.value = semaphore->timeline ? value : 0;
    CUDA_TRY( cudaWaitExternalSemaphoresAsync( &semaphore->sem, &p, 1, ( cudaStream_t )stream ) );
    return ASVGF_SUCCESS;
}

// "PreFinal written": signalled when everything enqueued on `stream` so far has finished (= the reference's frame fence)
AsvgfResult asvgfSignalExternalSemaphore( AsvgfExternalSemaphore semaphore, AsvgfStream stream, uint64_t value )
{
    if( !semaphore ) return fail( ASVGF_WRONG_ARGUMENT, "null semaphore" );
    DeviceGuard guard( semaphore->device );
    cudaExternalSemaphoreSignalParams p{};
    p.params.fence.value = semaphore->timeline ? value : 0;
    CUDA_TRY( cudaSignalExternalSemaphoresAsync( &semaphore->sem, &p, 1, ( cudaStream_t )stream ) );
    return ASVGF_SUCCESS;
}

AsvgfResult asvgfDestroyExternalSemaphore( AsvgfExternalSemaphore semaphore )
{
    if( !semaphore ) return ASVGF_SUCCESS;
    DeviceGuard guard( semaphore->device );
    cudaDestroyExternalSemaphore( semaphore->sem );
    delete semaphore;
    return ASVGF_SUCCESS;
}

} // extern "C"
