"""CPU suite (build container only): the C oracle against the unmodified reference compiled from /root/reference
into oracle/_ref/libhmref.so.  Skipped where that library does not exist."""
import numpy as np
import pytest

from common import PU_SIZES, padded
from oracle.pyoracle import Oracle, Reference, JOB_DTYPE
from video_codecs_b200 import synth


def test_eg_bits_and_cost_exhaustive_small(oracle, reference):
    for scale in (0, 1, 2):
        for pred in ((0, 0), (-7, 13), (255, -256)):
            for v in list(range(-70, 71)) + [-1000, 999, 4095, -4096]:
                assert oracle.mv_bits(v, -v, pred, scale) == reference.mv_bits(v, -v, pred, scale)
    for lam in (0, 1, 65535, 65536, 4037017, 9000000, 0xFFFFFFFF):
        for (x, y) in ((0, 0), (64, -64), (-3, 17)):
            assert oracle.mv_cost(lam, x, y, (5, -9), 2) == reference.mv_cost(lam, x, y, (5, -9), 2)


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_distortion_all_pu_sizes_random_strides(oracle, reference, bd):
    rng = np.random.default_rng(100 + bd)
    hi = 1 << bd
    for (w, h) in PU_SIZES + [(4, 4), (2, 2), (6, 2), (20, 8)]:
        for rep in range(6):
            so, sc = int(rng.integers(w, 97)), int(rng.integers(w, 131))
            org = rng.integers(0, hi, size=(h + 3) * so + 8).astype(np.int16)
            cur = rng.integers(0, hi, size=(h + 3) * sc + 8).astype(np.int16) if rep % 2 else \
                np.clip(org[: (h + 3) * so + 8].astype(np.int32) + rng.integers(-9, 10, size=(h + 3) * so + 8), 0, hi - 1).astype(np.int16)
            if rep % 2 == 0:
                sc = so
            oo, co = int(rng.integers(0, 8)), int(rng.integers(0, 8))
            for kind, ss in ((0, 0), (0, 1), (1, 0), (2, 0), (3, 0)):
                if kind == 2 and (w % 2 or h % 2):
                    continue
                a = oracle.dist(kind, (org, oo, so), (cur, co, sc), w, h, bd, ss)
                b = reference.dist(kind, (org, oo, so), (cur, co, sc), w, h, bd, ss)
                assert a == b, (w, h, bd, kind, ss)


def test_bipred_style_out_of_range_pattern(oracle, reference):
    """2*org - otherPred can leave the pixel range (TComYuv.cpp:440-450): signed 16-bit patterns must work."""
    rng = np.random.default_rng(5)
    org = rng.integers(-255, 511, size=64 * 64).astype(np.int16)
    cur = rng.integers(0, 256, size=64 * 64).astype(np.int16)
    for (w, h) in ((8, 8), (16, 16), (64, 64), (12, 16)):
        for kind in (0, 1, 2):
            assert oracle.dist(kind, (org, 0, 64), (cur, 0, 64), w, h, 8, 0) == reference.dist(kind, (org, 0, 64), (cur, 0, 64), w, h, 8, 0)


@pytest.mark.parametrize("fen,had", [(1, 1), (0, 0)])
@pytest.mark.parametrize("bd", [8, 10])
def test_search_and_refinement_every_pu_size(fen, had, bd):
    O, R = Oracle(fen=fen, hadme=had), Reference(fen=fen, hadme=had)
    W, H = 192, 128
    f0 = synth.luma_frame(W, H, 0, seed=11, bit_depth=bd)
    f1 = synth.luma_frame(W, H, 1, seed=11, bit_depth=bd, vx=-1.25, vy=0.75)
    cur, o0, stride = padded(f1)
    ref, _, _ = padded(f0)
    rng = np.random.default_rng(3)
    jobs = np.zeros(len(PU_SIZES) * 2, dtype=JOB_DTYPE)
    for i, (w, h) in enumerate(PU_SIZES * 2):
        px, py = int(rng.integers(0, (W - w) // 4 + 1)) * 4, int(rng.integers(0, (H - h) // 4 + 1)) * 4
        pred = (int(rng.integers(-30, 31)), int(rng.integers(-30, 31)))
        sr = 10
        cx, cy = pred[0] >> 2, pred[1] >> 2
        jobs[i] = (px, py, w, h, cx - sr, cy - sr, cx + sr, cy + sr, pred[0], pred[1], int(rng.integers(1000, 9000000)), 0)
    a, _ = O.run_jobs((cur, o0, stride), (ref, o0, stride), jobs, bit_depth=bd)
    b, _ = R.run_jobs((cur, o0, stride), (ref, o0, stride), jobs, bit_depth=bd)
    assert np.array_equal(a, b)


def test_refinement_all_81_paths(oracle, reference):
    """Force every (half, quarter) branch of xExtDIFUpSamplingQ by refining around each candidate start with a
    zero lambda on content shifted by every quarter-pel phase."""
    W, H = 96, 96
    seen = set()
    for t, (vx, vy) in enumerate([(a / 4.0, b / 4.0) for a in range(-3, 4) for b in range(-3, 4)]):
        f0 = synth.luma_frame(W, H, 0, seed=21, n_rect=0)
        f1 = synth.luma_frame(W, H, 1, seed=21, vx=vx, vy=vy, theta_deg=0.0, n_rect=0)
        cur, o0, stride = padded(f1)
        ref, _, _ = padded(f0)
        for (w, h) in ((16, 16), (8, 4), (12, 16)):
            off = o0 + 32 * stride + 32
            a = oracle.pattern_search_frac((cur, off, stride), w, h, (ref, off, stride), (round(vx), round(vy)), 0, (0, 0))
            b = reference.pattern_search_frac((cur, off, stride), w, h, (ref, off, stride), (round(vx), round(vy)), 0, (0, 0))
            assert a == b, (vx, vy, w, h)
            seen.add((a[0], a[1]))
    assert len(seen) >= 30      # many distinct (half, quarter) combinations are actually reached


def test_tie_break_is_raster_first(oracle, reference):
    """Flat content: every SAD ties; the winner is decided by MV cost, then by raster order."""
    cur = np.full((200, 200), 90, dtype=np.int16)
    ref = np.full((200, 200), 90, dtype=np.int16)
    off = 100 * 200 + 100
    for lam, pred in ((0, (0, 0)), (65536 * 4, (9, -6)), (65536 * 4, (2, 2))):
        a = oracle.pattern_search((cur, off, 200), 8, 8, (ref, off, 200), (-5, -4), (6, 7), lam, pred)
        b = reference.pattern_search((cur, off, 200), 8, 8, (ref, off, 200), (-5, -4), (6, 7), lam, pred)
        assert a == b
    assert oracle.pattern_search((cur, off, 200), 8, 8, (ref, off, 200), (-5, -4), (6, 7), 0, (0, 0))[0] == (-5, -4)


def _intra_lines(rng, n, bd, kind):
    hi = (1 << bd) - 1
    if kind == 0:
        top, left = rng.integers(0, hi + 1, 2 * n + 1), rng.integers(0, hi + 1, 2 * n + 1)
    elif kind == 1:      # smooth ramp + small noise (what smoothed reference samples look like)
        top = np.clip(np.linspace(rng.integers(0, hi), rng.integers(0, hi), 2 * n + 1) + rng.integers(-3, 4, 2 * n + 1), 0, hi)
        left = np.clip(np.linspace(top[0], rng.integers(0, hi), 2 * n + 1) + rng.integers(-3, 4, 2 * n + 1), 0, hi)
    else:                # extremes
        top, left = rng.choice([0, hi], 2 * n + 1), rng.choice([0, hi], 2 * n + 1)
    top = np.asarray(top, dtype=np.int16); left = np.asarray(left, dtype=np.int16)
    left[0] = top[0]
    return top, left


def test_intra_predictions_and_mode_distortions(oracle, reference):
    """All 35 luma intra predictions (planar, DC + edge filter, 33 angles incl. the filtered pure vertical / horizontal),
    the smoothed-reference decision and the per-mode Hadamard distortion of the first pass, restatement vs reference."""
    O, R = oracle, reference
    rng = np.random.default_rng(77)
    for n in (4, 8, 16, 32, 64):
        for mode in range(35):
            assert O.intra_use_filtered(mode, n) == R.intra_use_filtered(mode, n), (mode, n)
        for bd in (8, 10):
            for kind in range(3):
                top, left = _intra_lines(rng, n, bd, kind)
                for mode in range(35):
                    for above_ok, left_ok, ef in ((1, 1, 1), (1, 1, 0)) + (((1, 0, 1), (0, 1, 1), (0, 0, 1)) if mode == 1 else ()):
                        a = O.intra_predict(mode, top, left, n, bd, above_ok, left_ok, ef)
                        b = R.intra_predict(mode, top, left, n, bd, above_ok, left_ok, ef)
                        assert np.array_equal(a, b), (n, bd, kind, mode, above_ok, left_ok, ef)
                top2, left2 = _intra_lines(rng, n, bd, 1)
                org = rng.integers(0, 1 << bd, (n, n)).astype(np.int16)
                assert np.array_equal(O.intra_modes_had((org, 0, n), top, left, top2, left2, n, bd),
                                      R.intra_modes_had((org, 0, n), top, left, top2, left2, n, bd)), (n, bd, kind)
