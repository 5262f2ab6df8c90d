"""Wider sweeps of the host-compiled library (tests/hostsim, test infrastructure) under AddressSanitizer: run as
    LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 HS_LIB=<sanitizer build> \
        python tools/hostsim_sweep.py tight|chain|band|tsan
(the sanitizer build: the g++ line of tests/test_hostsim.py::test_kernel_addressing_under_address_sanitizer).
Results of the last run: profiles/r02_hostsim_memcheck.md."""
import os, sys, ctypes
import numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"tests"))
from raytracedgl1_b200 import denoiser as dn, layout
dn._LIB_PATH = os.environ.get("HS_LIB", "/tmp/hs/libasvgf_hostsim_asan.so")
dn._TEST_ONLY_ALLOW_HOST_BUILD = True
from _harness import Pair, assert_all_ok
import synth

def tight_external(W,H,N=4, extra=0):
    """every image caller-owned with a TIGHT pitch (16-byte aligned row bytes [+extra]): the non-uniform-pitch kernels"""
    own = dn.Denoiser(W,H,N, flags=dn.CREATE_EXPORT_DEBUG_MASK)
    ext = dn.Denoiser(W,H,N, flags=dn.CREATE_EXTERNAL_IMAGES | dn.CREATE_EXPORT_DEBUG_MASK)
    keep={}
    for name in layout.IMAGES:
        idx=layout.INDEX[name]
        f,w,h,b=[ctypes.c_uint32() for _ in range(4)]
        assert dn.lib().asvgfGetImageInfo(idx,W,H,ctypes.byref(f),ctypes.byref(w),ctypes.byref(h),ctypes.byref(b))==0
        pitch=((w.value*b.value+15)//16)*16 + 16*extra*(1 + idx%3)
        buf=np.zeros(pitch*h.value+16,dtype=np.uint8)
        off=(-buf.ctypes.data)%16
        keep[name]=buf
        ext.bind_external(name, buf.ctypes.data+off, pitch)
    for frame in range(3):
        fi=frame&1
        g=synth.generate(W,H,frame); u=synth.uniform(W,H,frame)
        for d in (own,ext):
            for name,arr in g.items(): d.upload(name,fi,arr)
            d.gradient_samples(fi,u); d.denoise(fi,u); d.synchronize()
        for name in layout.IMAGES:
            a=own.download(name,fi); b=ext.download(name,fi)
            assert np.array_equal(a.view(np.uint8),b.view(np.uint8)), (W,H,frame,name)
    own.close(); ext.close()
    print("tight external",W,H,N,extra,"identical",flush=True)

def chain(W,H,N,flags=None,frames=3,mutate=None):
    p=Pair(W,H,iterations=N,flags=flags)
    for frame in range(frames):
        fi,u=p.run_frame(frame,mutate=mutate)
        res=p.compare(fi); assert_all_ok(res,"%dx%d N%d f%d"%(W,H,N,frame))
        p.sync_state_to_cuda()
    p.close(); print("chain",W,H,N,flags,"ok",flush=True)

def band_handle(W,H,lo,hi,halo,apron):
    d=dn.Denoiser(W,H,4,flags=dn.CREATE_EXPORT_DEBUG_MASK|(dn.CREATE_BAND_APRON if apron else 0),band=(lo,hi),halo_rows=halo)
    for frame in range(3):
        fi=frame&1
        g=synth.generate(W,H,frame); u=synth.uniform(W,H,frame)
        for name,arr in g.items():
            _,_,w,r0,rows=d.image(name,fi)
            d.upload(name,fi,np.ascontiguousarray(arr[r0:r0+rows]))
        d.gradient_samples(fi,u); d.denoise(fi,u)
        try: d.synchronize()
        except dn.AsvgfError as e: assert e.code==dn.HALO_EXCEEDED
    d.close(); print("band handle",W,H,lo,hi,halo,apron,"ran",flush=True)

what=sys.argv[1]
if what=="tight":
    for (W,H) in [(34,10),(66,35),(130,49),(62,31)]:
        tight_external(W,H); tight_external(W,H,extra=1)
if what=="chain":
    for (W,H,N) in [(70,100,4),(258,20,4),(126,67,5),(130,49,2),(66,140,8),(2,1,8),(34,10,8)]:
        chain(W,H,N)
    chain(130,49,4,flags=dn.CREATE_EAGER_GRADIENT_ATROUS|dn.CREATE_EXPORT_DEBUG_MASK)
if what=="band":
    band_handle(34,300,99,201,66,True); band_handle(34,300,0,99,66,True); band_handle(34,300,201,300,66,True)
    band_handle(34,301,201,301,66,True); band_handle(34,100,33,66,9,False); band_handle(66,100,0,33,12,False); band_handle(66,100,66,100,30,False)
if what=="tsan":      # racecheck: HS_LIB = a -fsanitize=thread build, LD_PRELOAD = libtsan.so (profiles/r02_hostsim_memcheck.md)
    for (W,H,N) in [(130,20,4),(66,40,6)]:
        d=dn.Denoiser(W,H,N,flags=dn.CREATE_EXPORT_DEBUG_MASK)
        for frame in range(2):
            fi=frame&1; g=synth.generate(W,H,frame); u=synth.uniform(W,H,frame)
            for name,arr in g.items(): d.upload(name,fi,arr)
            d.gradient_samples(fi,u); d.denoise(fi,u); d.synchronize()
        d.calculate_exposure(fi,u); d.resolve_checkerboard(fi,u); d.synchronize()
        d.close(); print("tsan run",W,H,N,"done",flush=True)
