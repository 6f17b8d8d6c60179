"""Planar-YUV luma ingest: TVideoIOYuv::read for COMPONENT_Y — conformance padding and bit-depth
scaling — restated in the oracle, pinned against the reference's own reader, and done on the GPU straight from the
file bytes by hmb200_register_plane_yuv."""
import os

import numpy as np
import pytest

from common import ROOT

CASES = [  # width, height, pad_x, pad_y, file bit depth, internal bit depth
    (64, 40, 0, 0, 8, 8), (52, 36, 4, 4, 8, 8), (64, 24, 0, 8, 8, 10), (40, 24, 8, 0, 10, 10), (48, 32, 0, 0, 10, 8),
    (32, 16, 8, 8, 12, 10), (24, 16, 0, 0, 16, 12)]


def file_bytes(w, h, fbd, seed):
    rng = np.random.default_rng(seed)
    if fbd == 8:
        return rng.integers(0, 256, size=w * h, dtype=np.uint8).tobytes()
    return rng.integers(0, 1 << fbd, size=w * h).astype("<u2").tobytes()


@pytest.mark.parametrize("case", CASES)
def test_oracle_reader_matches_reference(oracle, reference, tmp_path, case):
    w, h, px, py, fbd, ibd = case
    data = file_bytes(w, h, fbd, 5)
    path = str(tmp_path / "luma.yuv")
    open(path, "wb").write(data)
    assert np.array_equal(oracle.read_luma(data, w, h, px, py, fbd, ibd), reference.read_luma(path, w, h, px, py, fbd, ibd))


def test_oracle_reader_golden(oracle):
    """Vectors produced by the reference reader in the build container (tests/golden/make_golden.py --ingest)."""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "yuv_ingest_golden.npz"))
    for i, (w, h, px, py, fbd, ibd) in enumerate(gold["cases"]):
        got = oracle.read_luma(gold[f"file_{i}"].tobytes(), int(w), int(h), int(px), int(py), int(fbd), int(ibd))
        assert np.array_equal(got, gold[f"plane_{i}"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_ingest_matches_oracle(hm, oracle, case):
    w, h, px, py, fbd, ibd = case
    data = file_bytes(w, h, fbd, 11)
    mx, my = 16, 24
    pid = hm.register_plane_yuv(data, w, h, px, py, fbd, ibd, mx, my)
    try:
        got = hm.read_plane(pid, w + px, h + py, mx, my)
    finally:
        hm.release_plane(pid)
    exp = np.pad(oracle.read_luma(data, w, h, px, py, fbd, ibd), ((my, my), (mx, mx)), mode="edge")
    assert np.array_equal(got, exp)


@pytest.mark.gpu
def test_gpu_ingest_feeds_the_search(hm):
    """1080 rows padded to 1088 like ConformanceWindowMode=1: ingest from file bytes == register_plane_u8 of the padded frame."""
    from video_codecs_b200 import synth, FLAG_FEN, FLAG_HADME, FLAG_FRAC
    W, H = 256, 120
    f0, f1 = synth.luma_frame(W, H, 0, seed=3), synth.luma_frame(W, H, 1, seed=3)
    pad = lambda f: np.pad(f, ((0, 8), (0, 0)), mode="edge")
    jobs = hm.build_canonical_jobs(W, H + 8, 64, 50000)[::7]
    a = [hm.register_plane_yuv(f.tobytes(), W, H, 0, 8, 8, 8, 80, 80) for f in (f1, f0)]
    b = [hm.register_plane_u8(pad(f), 80, 80) for f in (f1, f0)]
    try:
        ra = hm.me_jobs(a[0], a[1], jobs, FLAG_FEN | FLAG_HADME | FLAG_FRAC)
        rb = hm.me_jobs(b[0], b[1], jobs, FLAG_FEN | FLAG_HADME | FLAG_FRAC)
    finally:
        for p in a + b:
            hm.release_plane(p)
    assert np.array_equal(ra, rb)
