"""The library's own source on the CPU (tests/hostsim, TEST INFRASTRUCTURE): raytracedgl1_b200/csrc/asvgf.cu and every kernel it
includes are compiled for the host by g++ -- cuda_host_shim.h supplies threadIdx / __syncthreads / the rounded intrinsics and runs
a grid block by block, cuda_runtime_stub.cpp stands in for the runtime -- and a child pytest runs a selection of the `-m gpu`
tests, unchanged, against that build.  What this covers where no GPU exists: the host logic behind the C ABI (image table,
dispatch order, ping-pong routing, band schedule and halo pushes, state save / load, argument errors) and the arithmetic of the
kernels' source against the oracle.  What it cannot cover stays with `-m gpu`: the instructions nvcc generates (MUFU, packed
half, TMA), races, timing.  A second build under AddressSanitizer is the memcheck of the kernels' addressing (it found the
out-of-bounds gradient fetch of frames with H % 3 == 1 that the GPU run had read through silently).
"""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HS = os.path.join(ROOT, "tests", "hostsim")

# `-m gpu` tests that fit a CPU budget at their own sizes, plus the small-size cases of tests/hostsim_cases.py
SELECTION = [
    "tests/test_parity_gpu.py::test_tiny_and_ragged_sizes",
    "tests/test_parity_gpu.py::test_degenerate_inputs",
    "tests/test_parity_gpu.py::test_parameter_edges[two_iterations]",
    "tests/test_parity_gpu.py::test_abi_error_behaviour_on_gpu",
    "tests/test_parity_gpu.py::test_state_checkpoint_resumes_bit_identically",
    "tests/test_golden.py::test_cuda_path_matches_reference_shader_fixtures",
    "tests/test_post_passes.py::test_scatter_accumulation_parity",
    "tests/test_post_passes.py::test_histogram_every_half_pattern_and_flat_image",
    "tests/test_post_passes.py::test_post_passes_parity[size0]",
    "tests/test_bandset_gpu.py::test_band_set_argument_errors",
    "tests/hostsim_cases.py",
]


def _child_env(**extra):
    env = dict(os.environ, ASVGF_TEST_HOSTSIM="1", HOSTSIM_DEVICES="4", OMP_NUM_THREADS="4")
    env.update(extra)
    return env


def test_host_build_exports_the_whole_abi():
    import ctypes
    import hostsim
    from raytracedgl1_b200 import denoiser
    L = ctypes.CDLL(hostsim.build())
    assert not [s for s in denoiser.EXPORTED_SYMBOLS if not hasattr(L, s)]
    # and it is a different file from the product, which nothing under raytracedgl1_b200/ refers to
    assert os.path.realpath(hostsim.LIB_PATH) != os.path.realpath(os.path.join(ROOT, "raytracedgl1_b200", "libasvgf.so"))
    for dirpath, _, files in os.walk(os.path.join(ROOT, "raytracedgl1_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "libasvgf_hostsim" not in open(os.path.join(dirpath, f)).read(), f      # (the loader only knows how to refuse it)


def test_the_product_loader_refuses_the_host_build():
    """ASVGF_LIB (the A/B switch of the Python mirror) cannot be used to run the product on the CPU: the loader recognises the host
    build by its marker symbol and raises, as it does when libasvgf.so is missing."""
    import hostsim
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from raytracedgl1_b200 import denoiser\n"
            "try:\n    denoiser.lib()\nexcept ImportError as e:\n    print('refused:', e)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, ASVGF_LIB=hostsim.build()), capture_output=True, text=True, timeout=300)
    assert "refused:" in r.stdout and "no CPU fallback" in r.stdout, r.stdout + r.stderr


def test_gpu_tests_pass_on_the_host_compiled_library():
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu or not gpu", "-p", "no:cacheprovider"] + SELECTION,
                       cwd=ROOT, env=_child_env(), capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail


ASAN_DRIVER = r"""
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
from raytracedgl1_b200 import denoiser
denoiser._LIB_PATH = %(lib)r
denoiser._TEST_ONLY_ALLOW_HOST_BUILD = True
from _harness import Pair
for (W, H, N) in [(34, 10, 4), (130, 49, 5), (4, 3, 2)]:
    p = Pair(W, H, iterations=N)
    for frame in range(2):
        fi, u = p.run_frame(frame)
    u.volumeEnableType, u.volumeScattering = 1, 1500.0
    for side in (p.oracle, p.cuda):
        side.resolve_checkerboard(fi, u); side.luminance_histogram(fi, u); side.scatter_accumulate(fi, u, 16.0)
    p.cuda.synchronize()
    p.close()
from test_bandset_gpu import _run
_run(34, 201, [0, 1, 2], [0, 66, 135, 201], frames=2)
from hostsim_cases import run_tight_external
run_tight_external(34, 10, extra=1)
print("memcheck sweep done")
"""


def test_kernel_addressing_under_address_sanitizer(tmp_path):
    """The same source built with -fsanitize=address,alignment: every image is its own heap block, so a kernel that reads or writes
    one row past an image (the class of bug compute-sanitizer memcheck finds on the GPU) aborts here with the source line, and so
    does a vector access (uint2 / uint4 / float4: the host headers give them the GPU's alignment) at a misaligned address."""
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    libasan = subprocess.run([gxx, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not libasan or not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("no libasan for " + gxx)
    lib = str(tmp_path / "libasvgf_hostsim_asan.so")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.run([gxx, "-O0", "-g1", "-fsanitize=address,alignment", "-fno-sanitize-recover=alignment", "-fno-omit-frame-pointer", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off",
                    "-fno-fast-math", "-DASVGF_HOSTSIM", "-Wno-unknown-pragmas", "-I" + HS, "-I" + cuda + "/include", "-shared", "-o", lib,
                    "-include", "cuda_host_shim.h", "-x", "c++", os.path.join(ROOT, "raytracedgl1_b200", "csrc", "asvgf.cu"),
                    "-x", "c++", os.path.join(HS, "cuda_runtime_stub.cpp"), "-Wl,-Bsymbolic", "-ldl", "-lpthread"], check=True, cwd=HS)
    env = _child_env(LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    r = subprocess.run([sys.executable, "-c", ASAN_DRIVER % dict(root=ROOT, lib=lib)], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=1500)
    out = r.stdout + r.stderr
    assert "AddressSanitizer" not in out and "runtime error" not in out, out[-4000:]
    assert r.returncode == 0 and "memcheck sweep done" in out, out[-3000:]
