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


@pytest.mark.parametrize("fen,stop", [(1, 1), (0, 1), (1, 0), (0, 0)])
def test_oracle_tz_matches_reference(planes, fen, stop):
    from oracle.pyoracle import Oracle, Reference
    try:
        R = Reference(fen=fen, hadme=1)
    except (FileNotFoundError, OSError) as e:
        pytest.skip(str(e))
    O = Oracle(fen=fen, hadme=1)
    f0, f1, cur, ref, o0, stride = planes
    rng = np.random.default_rng(100 + fen + 2 * stop)
    stages = set()
    for (px, py, w, h, cu_x, cu_y, pred, lam, sr, imv) in make_cases(rng, 400):
        lt_rb = O.search_range(pred, sr, (cu_x, cu_y), (W, H))
        off = o0 + py * stride + px
        a = O.tz_search((cur, off, stride), w, h, (ref, off, stride), lt_rb[:2], lt_rb[2:], lam, pred, (cu_x, cu_y), (W, H), sr, imv,
                        first_search_stop=stop)
        b = R.tz_search((cur, off, stride), w, h, (ref, off, stride), lt_rb[:2], lt_rb[2:], lam, pred, (cu_x, cu_y), (W, H), sr, imv,
                        first_search_stop=stop)
        assert a == b, (px, py, w, h, pred, lam, sr, imv)
        stages.add(a[0])
    assert len(stages) > 20            # the MVs are not trivial


def test_oracle_tz_golden(planes):
    from oracle.pyoracle import Oracle
    gold = np.load(os.path.join(ROOT, "tests", "golden", "tz_golden.npz"))
    f0, f1, cur, ref, o0, stride = planes
    assert np.array_equal(gold["frame0"], f0) and np.array_equal(gold["frame1"], f1)
    O = Oracle(fen=1, hadme=1)
    for stop in (1, 0):
        for c, e in zip(gold["cases"], gold[f"expected_stop{stop}"]):
            px, py, w, h, cu_x, cu_y, pdx, pdy, lam, sr, has_imv, ix, iy, ltx, lty, rbx, rby = [int(v) for v in c]
            off = o0 + py * stride + px
            got = O.tz_search((cur, off, stride), w, h, (ref, off, stride), (ltx, lty), (rbx, rby), lam, (pdx, pdy), (cu_x, cu_y), (W, H), sr,
                              (ix, iy) if has_imv else None, first_search_stop=stop)
            assert (got[0][0], got[0][1], got[1]) == tuple(int(v) for v in e)


# ---------------------------------------------------------------------------------------------------------------------
# GPU: hmb200_tz_jobs through the C-ABI against the oracle
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_tz_jobs(O, cur, ref, o0, stride, jobs, extra, pic_wh, sr, bit_depth=8, stop=1):
    from video_codecs_b200 import RESULT_DTYPE
    out = np.zeros(len(jobs), dtype=RESULT_DTYPE)
    for i, (j, e) in enumerate(zip(jobs, extra)):
        off = o0 + int(j["pu_y"]) * stride + int(j["pu_x"])
        imv = (int(e["imv_x"]), int(e["imv_y"])) if e["has_imv"] else None
        mv, sad = O.tz_search((cur, off, stride), int(j["w"]), int(j["h"]), (ref, off, stride), (int(j["lt_x"]), int(j["lt_y"])),
                              (int(j["rb_x"]), int(j["rb_y"])), int(j["lambda_cost"]), (int(j["pred_x"]), int(j["pred_y"])),
                              (int(e["cu_x"]), int(e["cu_y"])), pic_wh, sr, imv, bit_depth, first_search_stop=stop)
        half, qter, cost = O.pattern_search_frac((cur, off, stride), int(j["w"]), int(j["h"]), (ref, off, stride), mv, int(j["lambda_cost"]),
                                                 (int(j["pred_x"]), int(j["pred_y"])), bit_depth)
        out[i] = (mv[0], mv[1], sad, half[0], half[1], qter[0], qter[1], cost)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("fen,stop", [(1, 1), (0, 1), (1, 0)])
def test_gpu_tz_random_cases_vs_oracle(hm, planes, fen, stop):
    from oracle.pyoracle import Oracle
    from video_codecs_b200 import JOB_DTYPE, TZ_EXTRA_DTYPE, FLAG_FEN, FLAG_HADME, FLAG_FRAC, FLAG_TZ_STOP
    f0, f1, cur, ref, o0, stride = planes
    O = Oracle(fen=fen, hadme=1)
    idc, idr = hm.register_plane_u8(f1, MARGIN, MARGIN, kind=0), hm.register_plane_u8(f0, MARGIN, MARGIN, kind=1)
    try:
        for sr in (8, 16, 64):
            cases = [c for c in make_cases(np.random.default_rng(500 + sr + fen), 360) if c[8] == sr]
            jobs, extra = np.zeros(len(cases), dtype=JOB_DTYPE), np.zeros(len(cases), dtype=TZ_EXTRA_DTYPE)
            for i, (px, py, w, h, cu_x, cu_y, pred, lam, _, imv) in enumerate(cases):
                lt_rb = O.search_range(pred, sr, (cu_x, cu_y), (W, H))
                jobs[i] = (px, py, w, h) + tuple(lt_rb) + (pred[0], pred[1], lam, 0)
                extra[i] = (cu_x, cu_y, 0 if imv is None else 1, 0 if imv is None else imv[0], 0 if imv is None else imv[1], (0, 0, 0))
            got = hm.tz_jobs(idc, idr, jobs, extra, (W, H), sr, (FLAG_FEN if fen else 0) | FLAG_HADME | FLAG_FRAC | (FLAG_TZ_STOP if stop else 0))
            exp = _oracle_tz_jobs(O, cur, ref, o0, stride, jobs, extra, (W, H), sr, stop=stop)
            bad = [i for i in range(len(jobs)) if got[i] != exp[i]]
            assert bad == [], (sr, bad[:5], got[bad[:1]], exp[bad[:1]], jobs[bad[:1]], extra[bad[:1]])
    finally:
        hm.release_plane(idc)
        hm.release_plane(idr)


@pytest.mark.gpu
def test_gpu_tz_canonical_and_golden(hm, planes):
    """Canonical list of a whole small picture (every PU shape, clipped windows) and the reference's golden vectors."""
    from oracle.pyoracle import Oracle
    from video_codecs_b200 import JOB_DTYPE, TZ_EXTRA_DTYPE, FLAG_FEN, FLAG_HADME, FLAG_FRAC, FLAG_TZ_STOP
    f0, f1, cur, ref, o0, stride = planes
    O = Oracle(fen=1, hadme=1)
    idc, idr = hm.register_plane_u8(f1, MARGIN, MARGIN, kind=0), hm.register_plane_u8(f0, MARGIN, MARGIN, kind=1)
    try:
        jobs = hm.build_canonical_jobs(W, H, 64, 635239)[::5]
        extra = hm.canonical_tz_extra(jobs)
        got = hm.tz_jobs(idc, idr, jobs, extra, (W, H), 64, FLAG_FEN | FLAG_HADME | FLAG_FRAC | FLAG_TZ_STOP)
        exp = _oracle_tz_jobs(O, cur, ref, o0, stride, jobs, extra, (W, H), 64)
        assert np.array_equal(got, exp)
        gold = np.load(os.path.join(ROOT, "tests", "golden", "tz_golden.npz"))
        c = gold["cases"]
        gj, ge = np.zeros(len(c), dtype=JOB_DTYPE), np.zeros(len(c), dtype=TZ_EXTRA_DTYPE)
        for i, r in enumerate(c):
            px, py, w, h, cu_x, cu_y, pdx, pdy, lam, sr, has_imv, ix, iy, ltx, lty, rbx, rby = [int(v) for v in r]
            gj[i] = (px, py, w, h, ltx, lty, rbx, rby, pdx, pdy, lam, 0)
            ge[i] = (cu_x, cu_y, has_imv, ix, iy, (0, 0, 0))
        for stop in (1, 0):
            for sr in (8, 16, 64):
                sel = np.nonzero(c[:, 9] == sr)[0]
                res = hm.tz_jobs(idc, idr, gj[sel], ge[sel], (W, H), sr, FLAG_FEN | (FLAG_TZ_STOP if stop else 0))
                e = gold[f"expected_stop{stop}"][sel]
                assert np.array_equal(np.stack([res["mv_x"], res["mv_y"], res["sad"]], 1).astype(np.int64), e)
    finally:
        hm.release_plane(idc)
        hm.release_plane(idr)
