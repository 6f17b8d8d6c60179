"""One host process, several hmb200 contexts on their own threads (tests/cpp/multi_ctx_test.cpp, plain C++ over the C-ABI):
the lookahead sharding a C++ encoder would do.  The binary is built by __graft_entry__.build() (g++, no CUDA needed)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "multi_ctx_test")


def build_binary():
    src = os.path.join(ROOT, "tests", "cpp", "multi_ctx_test.cpp")
    libdir = os.path.join(ROOT, "video_codecs_b200")
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", src, "-I", os.path.join(ROOT, "include"), "-L", libdir, "-lhmb200",
                               "-Wl,-rpath," + libdir, "-Wl,-rpath,$ORIGIN/../../../video_codecs_b200", "-o", BIN])
    return BIN


def test_binary_builds_and_links():
    assert os.path.exists(build_binary())


@pytest.mark.gpu
def test_two_contexts_two_threads_match_single_context():
    import torch
    n_dev = max(1, torch.cuda.device_count())
    for workers in (2, 3):
        r = subprocess.run([build_binary(), str(workers), str(min(n_dev, workers))], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        assert r.stdout.startswith("ok:"), r.stdout
