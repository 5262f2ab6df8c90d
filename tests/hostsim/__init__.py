"""TEST INFRASTRUCTURE ONLY: builds tests/hostsim/libasvgf_hostsim.so -- raytracedgl1_b200/csrc/asvgf.cu and its kernels
compiled for the host by g++ (cuda_host_shim.h, cuda_runtime_stub.cpp) -- and returns its path.  See README.md."""
import os
import subprocess

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libasvgf_hostsim.so")


def build():
    subprocess.run(["make", "-s", "-C", _DIR, "libasvgf_hostsim.so"], check=True)
    return LIB_PATH
