"""CPU suite: the C-ABI library loads without a GPU, exports every symbol include/hmb200.h declares, fails loudly
when asked to compute without a device, and its host-side window / job-list logic matches the oracle."""
import os
import re

import numpy as np
import pytest

from common import ROOT
from video_codecs_b200 import HMB200, HMB200Error


@pytest.fixture(scope="module")
def lib():
    return HMB200()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "hmb200.h")).read()
    declared = set(re.findall(r"\b(hmb200_[a-z0-9_]+)\s*\(", header))
    declared -= {"hmb200_prepared"}
    assert len(declared) >= 23
    import ctypes
    L = ctypes.CDLL(os.path.join(ROOT, "video_codecs_b200", "libhmb200.so"))
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert missing == []
    assert set(lib.exported) == declared


def test_no_libcuda_link_dependency():
    import subprocess
    out = subprocess.run(["ldd", os.path.join(ROOT, "video_codecs_b200", "libhmb200.so")], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out


def test_compute_without_device_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(HMB200Error, match="no CUDA device|CUDA"):
        lib.init(0)
    with pytest.raises(HMB200Error):
        lib.me_jobs(0, 1, np.zeros(0, dtype=lib.build_canonical_jobs(64, 64).dtype))


def test_search_range_matches_oracle(lib, oracle):
    rng = np.random.default_rng(9)
    for _ in range(3000):
        pw, ph = int(rng.integers(2, 61)) * 8, int(rng.integers(2, 35)) * 8
        cu = (int(rng.integers(0, pw // 8)) * 8, int(rng.integers(0, ph // 8)) * 8)
        pred = (int(rng.integers(-3000, 3001)), int(rng.integers(-3000, 3001)))
        sr = int(rng.choice([1, 4, 8, 64, 96, 128, 256]))
        assert lib.set_search_range(pred, sr, cu, (pw, ph)) == oracle.search_range(pred, sr, cu, (pw, ph))
    # far-out predictors are clipped first, then the corners
    assert lib.set_search_range((32000, -32000), 64, (0, 0), (416, 240)) == oracle.search_range((32000, -32000), 64, (0, 0), (416, 240))


def test_canonical_job_list_shape(lib):
    jobs = lib.build_canonical_jobs(128, 64, search_range=64, lambda_cost=1234)
    assert len(jobs) == 2 * 593
    area = (jobs["w"].astype(np.int64) * jobs["h"]).sum()
    assert area == 24 * 2 * 4096                       # area multiplier 24.0
    fen_area = (jobs["w"].astype(np.int64) * np.where(jobs["h"] > 8, jobs["h"] // 2, jobs["h"])).sum()
    assert fen_area * 2 == 29 * 2 * 4096               # 14.5 with FEN
    sizes = set(zip(jobs["w"].tolist(), jobs["h"].tolist()))
    assert len(sizes) == 24 and (4, 8) in sizes and (64, 48) in sizes and (12, 16) in sizes and (4, 4) not in sizes
    # windows near the picture corner are clipped to (maxCU + 8 - 1) to the left / top of the CU
    j0 = jobs[0]
    assert (j0["lt_x"], j0["lt_y"], j0["rb_x"], j0["rb_y"]) == (-64, -64, 64, 64)
    full = lib.build_canonical_jobs(1920, 1088)
    assert len(full) == 510 * 593
    part = lib.build_canonical_jobs(1920, 1088, ctu_first=30, ctu_count=30)
    assert len(part) == 30 * 593 and part["pu_y"].min() == 64 and part["pu_y"].max() < 128
    # partial CTUs: only CUs inside the picture (416x240 is 6.5 x 3.75 CTUs)
    small = lib.build_canonical_jobs(416, 240)
    assert (small["pu_x"] + small["w"]).max() <= 416 and (small["pu_y"] + small["h"]).max() <= 240
    assert len(small) < 24.4 * 593
    last_row = small[small["pu_y"] >= 192]
    assert last_row["rb_y"].max() == 240 + 8 - 192 - 1      # clipMv: picH + 8 - cuY - 1 (TComDataCU.cpp:2795)


def test_window_clipping_at_borders(lib, oracle):
    jobs = lib.build_canonical_jobs(416, 240, search_range=64)
    for j in jobs[::97]:
        size = 64
        # the CU that owns the PU is unknown from the job alone; the builder uses the true CU, and any CU containing the
        # PU gives a window that contains the job's window when clipped less: check only the invariants
        assert j["lt_x"] <= j["rb_x"] and j["lt_y"] <= j["rb_y"]
        assert j["pu_x"] + j["lt_x"] >= -(size + 8) and j["pu_x"] + j["w"] - 1 + j["rb_x"] <= 416 + 8 + size
        assert j["pu_y"] + j["lt_y"] >= -(size + 8) and j["pu_y"] + j["h"] - 1 + j["rb_y"] <= 240 + 8 + size


def test_motion_lambda_cost_matches_reference(lib, reference):
    """TComRdCost::setLambda -> m_uiCost, for the lambdas HM derives from QP 0..51 at depth 0 and deeper (x4)."""
    import ctypes as C
    f = reference.lib.hmref_motion_lambda_cost
    f.restype, f.argtypes = C.c_uint32, [C.c_double, C.c_int]
    for qp in range(0, 52):
        for scale in (0.4624, 0.578, 0.57, 1.0):
            lam = scale * 2 ** ((qp - 12) / 3.0)
            assert lib.motion_lambda_cost(lam) == f(lam, 8)
    assert lib.motion_lambda_cost(0.4624 * 2 ** ((35 - 12) / 3.0)) == 635239


def test_python_canonical_job_list_equals_the_library_builder():
    """oracle.pyoracle.py_canonical_jobs (what the reference arm of bench.py uses, so that it never loads libhmb200) against
    hmb200_build_canonical_jobs: same PUs, windows and order, incl. clipped windows and a non-zero predictor."""
    from oracle.pyoracle import py_canonical_jobs
    from video_codecs_b200 import HMB200
    hm = HMB200()
    for (w, h, sr, pred) in [(416, 240, 64, (0, 0)), (320, 256, 128, (6, -3)), (200, 136, 16, (-300, 500))]:
        assert np.array_equal(hm.build_canonical_jobs(w, h, sr, 777, pred=pred), py_canonical_jobs(w, h, sr, 777, pred))
    assert np.array_equal(hm.build_canonical_jobs(416, 240, 64, 5, ctu_first=9, ctu_count=2), py_canonical_jobs(416, 240, 64, 5, ctu_first=9, ctu_count=2))


def test_whole_cu_partition_geometry_matches_the_reference_table(tmp_path):
    """one_cu_part / one_cu_pus (hmb200_one.cuh), which the whole-CU kernels and the host-side answer cache share, against an
    independent restatement of TComDataCU::getPartIndexAndSize (tests/cpp/one_cu_part_test.cu): host code only, built with nvcc."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "one_cu_part_test")
    subprocess.check_call([nvcc, "-std=c++17", "-I", os.path.join(ROOT, "video_codecs_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                           "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, os.path.join(ROOT, "tests", "cpp", "one_cu_part_test.cu")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr
