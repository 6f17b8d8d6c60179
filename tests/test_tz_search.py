"""TZ fast search: TEncSearch::xTZSearch, FastSearch = 1 — the encoder's default integer search.
CPU part: the oracle's restatement against the unmodified reference (xTZSearch driven through oracle/ref_harness.cpp
with a bare TComDataCU for clipMv) and against golden vectors; GPU part: hmb200_tz_jobs against the oracle."""
import os

import numpy as np
import pytest

from common import PU_SIZES, MARGIN, ROOT, padded
from video_codecs_b200 import synth

W, H = 256, 192


def make_cases(rng, n, hm_range=None):
    """(px, py, w, h, pred, lam, sr, imv) with HM-consistent windows computed by the caller."""
    cases = []
    for i in range(n):
        w, h = PU_SIZES[int(rng.integers(0, len(PU_SIZES)))]
        s = max(w, h)
        cu_x, cu_y = int(rng.integers(0, (W - s) // s + 1)) * s, int(rng.integers(0, (H - s) // s + 1)) * s
        px = cu_x + (int(rng.integers(0, (s - w) // 4 + 1)) * 4 if w < s else 0)
        py = cu_y + (int(rng.integers(0, (s - h) // 4 + 1)) * 4 if h < s else 0)
        kind = i % 4
        pred = (0, 0) if kind == 0 else (int(rng.integers(-120, 121)), int(rng.integers(-120, 121))) if kind < 3 else \
            (int(rng.integers(-900, 901)), int(rng.integers(-700, 701)))
        lam = int(rng.choice([0, 40000, 635239, 4000000]))
        sr = int(rng.choice([8, 16, 64]))
        imv = None if i % 3 else (int(rng.integers(-40, 41)), int(rng.integers(-30, 31)))
        cases.append((px, py, w, h, cu_x, cu_y, pred, lam, sr, imv))
    return cases


@pytest.fixture(scope="module")
def planes():
    f0 = synth.luma_frame(W, H, 0, seed=17)
    f1 = synth.luma_frame(W, H, 3, seed=17)          # three frames apart: larger motion, raster + star stages get used
    cur, o0, stride = padded(f1)
    ref, _, _ = padded(f0)
    return f0, f1, cur, ref, o0, stride


@pytest.mark.parametrize("fen", [1, 0])
def test_oracle_tz_matches_reference(planes, fen):
    from oracle.pyoracle import Oracle, Reference
    try:
        R = Reference(fen=fen, hadme=1)
    except (FileNotFoundError, OSError) as e:
        pytest.skip(str(e))
    O = Oracle(fen=fen, hadme=1)
    f0, f1, cur, ref, o0, stride = planes
    rng = np.random.default_rng(100 + fen)
    stages = set()
    for (px, py, w, h, cu_x, cu_y, pred, lam, sr, imv) in make_cases(rng, 400):
        lt_rb = O.search_range(pred, sr, (cu_x, cu_y), (W, H))
        off = o0 + py * stride + px
        a = O.tz_search((cur, off, stride), w, h, (ref, off, stride), lt_rb[:2], lt_rb[2:], lam, pred, (cu_x, cu_y), (W, H), sr, imv)
        b = R.tz_search((cur, off, stride), w, h, (ref, off, stride), lt_rb[:2], lt_rb[2:], lam, pred, (cu_x, cu_y), (W, H), sr, imv)
        assert a == b, (px, py, w, h, pred, lam, sr, imv)
        stages.add(a[0])
    assert len(stages) > 20            # the MVs are not trivial


def test_oracle_tz_golden(planes):
    from oracle.pyoracle import Oracle
    gold = np.load(os.path.join(ROOT, "tests", "golden", "tz_golden.npz"))
    f0, f1, cur, ref, o0, stride = planes
    assert np.array_equal(gold["frame0"], f0) and np.array_equal(gold["frame1"], f1)
    O = Oracle(fen=1, hadme=1)
    for c, e in zip(gold["cases"], gold["expected"]):
        px, py, w, h, cu_x, cu_y, pdx, pdy, lam, sr, has_imv, ix, iy, ltx, lty, rbx, rby = [int(v) for v in c]
        off = o0 + py * stride + px
        got = O.tz_search((cur, off, stride), w, h, (ref, off, stride), (ltx, lty), (rbx, rby), lam, (pdx, pdy), (cu_x, cu_y), (W, H), sr,
                          (ix, iy) if has_imv else None)
        assert (got[0][0], got[0][1], got[1]) == tuple(int(v) for v in e)
