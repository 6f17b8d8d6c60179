"""Builds video_codecs_b200/libhmb200.so (sm_100a only) with nvcc, in-tree."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhmb200.so")
SRC = [os.path.join(HERE, "csrc", "hmb200_api.cu")]
DEPS = SRC + [os.path.join(HERE, "csrc", f) for f in
              ("hmb200_device.cuh", "hmb200_generic.cuh", "hmb200_search8.cuh", "hmb200_search8_cu.cuh", "hmb200_search16_cu.cuh", "hmb200_frac.cuh", "hmb200_tz.cuh", "hmb200_intra.cuh", "hmb200_mc_cand.cuh", "hmb200_one.cuh")] + \
       [os.path.join(HERE, "..", "include", "hmb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-cudart", "static", "--expt-relaxed-constexpr"]


def is_stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not is_stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + SRC
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
