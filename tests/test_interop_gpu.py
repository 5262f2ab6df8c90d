"""The CUDA half of the Vulkan interop (include/asvgf.h, asvgfImportExternalMemoryFd ...).

RTGL1 would export each framebuffer image's memory with vkGetMemoryFdKHR (Source/Framebuffers.cpp:702-804 allocates them) and
hand the descriptor over.  This image has no Vulkan loader, so the exporting side of the test is the CUDA driver's own
virtual-memory API: cuMemCreate with a POSIX-fd shareable handle type + cuMemExportToShareableHandle gives the same kind of
descriptor an NVIDIA Vulkan driver returns for OPAQUE_FD.  The fd is imported through the C ABI, mapped as linear device
memory, bound as producer images with asvgfBindExternalImage, and the frame computed from them must equal the frame computed
from library-owned images bit for bit.  Semaphores cannot be exported by CUDA itself: only the argument checks run here."""
import ctypes
import os

import numpy as np
import pytest

import synth
from raytracedgl1_b200 import layout

pytestmark = pytest.mark.gpu


def _drv():
    from cuda.bindings import driver
    return driver


def _ck(res):
    err, rest = res[0], res[1:]
    if int(err) != 0:
        raise RuntimeError("CUDA driver call failed: %s" % (err,))
    return rest[0] if len(rest) == 1 else rest


class _Exported:
    """One physical allocation of the driver's VMM API with a POSIX fd; also mapped locally so the test can fill it."""

    def __init__(self, nbytes, device=0):
        d = _drv()
        import torch
        torch.cuda.init()
        torch.zeros(1, device="cuda:%d" % device)  # primary context current on this thread
        prop = d.CUmemAllocationProp()
        prop.type = d.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        prop.location.type = d.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        prop.location.id = device
        prop.requestedHandleTypes = d.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        gran = _ck(d.cuMemGetAllocationGranularity(prop, d.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
        self.size = (nbytes + gran - 1) // gran * gran
        self.handle = _ck(d.cuMemCreate(self.size, prop, 0))
        self.va = _ck(d.cuMemAddressReserve(self.size, 0, 0, 0))
        _ck(d.cuMemMap(self.va, self.size, 0, self.handle, 0))
        acc = d.CUmemAccessDesc()
        acc.location.type = d.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        acc.location.id = device
        acc.flags = d.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
        _ck(d.cuMemSetAccess(self.va, self.size, [acc], 1))
        self.fd = int(_ck(d.cuMemExportToShareableHandle(self.handle, d.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0)))

    def write(self, offset, host_bytes):
        d = _drv()
        buf = np.ascontiguousarray(host_bytes).view(np.uint8).reshape(-1)
        _ck(d.cuMemcpyHtoD(int(self.va) + offset, buf.ctypes.data, buf.size) + (None,))

    def close(self):
        d = _drv()
        d.cuMemUnmap(self.va, self.size)
        d.cuMemAddressFree(self.va, self.size)
        d.cuMemRelease(self.handle)


def _pitched(arr, pitch):
    h = arr.shape[0]
    row = arr.reshape(h, -1).view(np.uint8)
    out = np.zeros((h, pitch), np.uint8)
    out[:, :row.shape[1]] = row
    return out


def test_frame_from_imported_memory_equals_frame_from_owned_images():
    from raytracedgl1_b200 import denoiser as dn
    W, H = 320, 180
    owned = dn.Denoiser(W, H, 4, flags=dn.CREATE_EXPORT_DEBUG_MASK)
    ext = dn.Denoiser(W, H, 4, flags=dn.CREATE_EXPORT_DEBUG_MASK)
    exported, imported = [], []
    try:
        for frame in range(3):
            fi = frame & 1
            g = synth.generate(W, H, frame)
            u = synth.uniform(W, H, frame)
            # one exported allocation per frame holding every producer image at a 256-byte aligned offset (one VkDeviceMemory
            # with several bound images is what a Vulkan host would export)
            offsets, total = {}, 0
            for name, arr in g.items():
                pitch = (arr.shape[1] * arr.dtype.itemsize * (arr.shape[2] if arr.ndim == 3 else 1) + 255) // 256 * 256
                offsets[name] = (total, pitch)
                total += pitch * arr.shape[0]
            ex = _Exported(total)
            exported.append(ex)
            for name, arr in g.items():
                off, pitch = offsets[name]
                ex.write(off, _pitched(arr, pitch))
            mem = dn.ExternalMemory(ex.fd, ex.size, device=0)
            imported.append(mem)
            base = mem.buffer()
            assert base and base != int(ex.va), "the import maps the allocation at an address of its own"
            for name, arr in g.items():
                off, pitch = offsets[name]
                # asvgfBindExternalImage addresses PHYSICAL images: slot 1 of a cur/prev pair is X_Prev (Framebuffers.cpp:35-50)
                phys = name + "_Prev" if (fi == 1 and name + "_Prev" in layout.IMAGES) else name
                ext.bind_external(phys, base + off, pitch)
                owned.upload(name, fi, arr)
            for d in (owned, ext):
                d.gradient_samples(fi, u)
                d.denoise(fi, u)
                d.synchronize()
            for name in ("PreFinal", "HistogramInput", "DiffColorHistory", "SpecAccumColor", "IndirAccum", "DebugMask", "GradientPrevPix"):
                a, b = owned.download(name, fi), ext.download(name, fi)
                assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), "frame %d: %s" % (frame, name)
    finally:
        ext.close()
        owned.close()
        for m in imported:
            m.close()
        for e in exported:
            e.close()


def test_imported_memory_range_and_argument_checks():
    from raytracedgl1_b200 import denoiser as dn
    ex = _Exported(1 << 20)
    try:
        mem = dn.ExternalMemory(ex.fd, ex.size, device=0)
        with pytest.raises(dn.AsvgfError):
            mem.buffer(offset=ex.size - 16, size=64)      # past the end of the allocation
        with pytest.raises(dn.AsvgfError):
            mem.buffer(offset=0, size=0)
        assert mem.buffer(offset=4096, size=4096)
        mem.close()
    finally:
        ex.close()
    with pytest.raises(dn.AsvgfError):
        dn.ExternalMemory(-1, 4096)
    with pytest.raises(dn.AsvgfError):
        dn.ExternalMemory(os.open("/dev/null", os.O_RDONLY), 4096)   # a descriptor that is not an exported allocation


def test_semaphore_import_rejects_what_is_not_a_semaphore():
    from raytracedgl1_b200 import denoiser as dn
    with pytest.raises(dn.AsvgfError):
        dn.ExternalSemaphore(-1)
    fd = os.open("/dev/null", os.O_RDONLY)
    try:
        with pytest.raises(dn.AsvgfError):
            dn.ExternalSemaphore(fd, timeline=True)
    finally:
        try:
            os.close(fd)
        except OSError:
            pass
    L = dn.lib()
    assert L.asvgfWaitExternalSemaphore(None, None, 0) != 0 and L.asvgfSignalExternalSemaphore(None, None, 0) != 0
    assert L.asvgfDestroyExternalSemaphore(None) == 0 and L.asvgfDestroyExternalMemory(None) == 0
