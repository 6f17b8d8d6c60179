"""Intra first pass: the 35 luma predictions of a block and their Hadamard distortions,
TEncSearch::estIntraPredQT's mode loop (TLibEncoder/TEncSearch.cpp:2270-2296).  CPU: the oracle against golden vectors
generated from the unmodified reference (tests/golden/make_intra_golden.py); GPU: the CUDA path through the C-ABI against
the same vectors and against the oracle on random inputs.  Bar: bit-exact."""
import os

import numpy as np
import pytest

from common import ROOT

GOLD = os.path.join(ROOT, "tests", "golden", "intra_golden.npz")


def _lines(refs, off, n):
    L = 2 * n + 1
    return [refs[off + k * L: off + (k + 1) * L] for k in range(4)]


@pytest.mark.parametrize("bd", [8, 10])
def test_oracle_intra_modes_vs_golden(oracle, bd):
    g = np.load(GOLD)
    plane, blocks, refs, exp = g[f"plane{bd}"], g[f"blocks{bd}"], g[f"refs{bd}"], g[f"had{bd}"]
    W = plane.shape[1]
    for i, (x, y, n, off, _, _) in enumerate(blocks):
        ln = _lines(refs, off, n)
        got = oracle.intra_modes_had((np.ascontiguousarray(plane), y * W + x, W), ln[0], ln[1], ln[2], ln[3], n, bd)
        assert np.array_equal(got, exp[i]), (bd, i, n)
    assert len(set(exp[:, 0].tolist())) > 3


def test_oracle_intra_predictions_vs_golden(oracle):
    g = np.load(GOLD)
    for n in (8, 32):
        ln = _lines(g[f"pred_lines{n}"], 0, n)
        exp = g[f"preds{n}"].reshape(35, n, n)
        for mode in range(35):
            f = oracle.intra_use_filtered(mode, n)
            got = oracle.intra_predict(mode, ln[2] if f else ln[0], ln[3] if f else ln[1], n, 8)
            assert np.array_equal(got, exp[mode]), (n, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("bd", [8, 10])
def test_gpu_intra_modes_vs_golden(hm, bd):
    from video_codecs_b200 import INTRA_BLOCK_DTYPE
    g = np.load(GOLD)
    plane, blocks, refs, exp = g[f"plane{bd}"], g[f"blocks{bd}"], g[f"refs{bd}"], g[f"had{bd}"]
    H, W = plane.shape
    if bd == 8:
        pid = hm.register_plane_u8(plane.astype(np.uint8), 16, 16, kind=0)
    else:
        from video_codecs_b200 import synth
        pid = hm.register_plane(synth.pad_plane(plane, 16, 16), W, H, 16, 16, bd, kind=0)
    try:
        rec = np.zeros(len(blocks), dtype=INTRA_BLOCK_DTYPE)
        for k, name in enumerate(("x", "y", "n", "ref_off", "flags", "reserved")):
            rec[name] = blocks[:, k]
        got = hm.intra_modes_had_batch(pid, rec, refs)
    finally:
        hm.release_plane(pid)
    assert np.array_equal(got, exp)


@pytest.mark.gpu
def test_gpu_intra_one_to_one_vs_oracle(hm, oracle):
    """The 1:1 entry with the reference's own buffer layout ((2n+1)-strided predictor buffers), random and extreme
    reference samples, every block size, 8 / 10 / 12 bit, and the bAbove / bLeft combinations of the DC mode."""
    rng = np.random.default_rng(5)
    for n in (4, 8, 16, 32, 64):
        L = 2 * n + 1
        for bd in (8, 10, 12):
            hi = (1 << bd) - 1
            for kind in range(3):
                bufs = []
                for _ in range(2):
                    b = np.zeros((L, L), dtype=np.int16)
                    if kind == 2:
                        b[0, :] = rng.choice([0, hi], L); b[:, 0] = rng.choice([0, hi], L)
                    else:
                        b[0, :] = rng.integers(0, hi + 1, L); b[:, 0] = rng.integers(0, hi + 1, L)
                    bufs.append(b)
                org = rng.integers(0, hi + 1, (n + 2, n + 5)).astype(np.int16)
                for above, left in ((1, 1), (1, 0), (0, 1), (0, 0)) if kind == 0 else ((1, 1),):
                    got = hm.intra_modes_had((org, 3, n + 5), bufs[0], bufs[1], n, bd, above, left)
                    lines = [np.ascontiguousarray(bufs[0][0, :]), np.ascontiguousarray(bufs[0][:, 0]),
                             np.ascontiguousarray(bufs[1][0, :]), np.ascontiguousarray(bufs[1][:, 0])]
                    exp = np.zeros(35, dtype=np.uint32)
                    o = (org, 3, n + 5)
                    if (above, left) == (1, 1):
                        exp = oracle.intra_modes_had(o, *lines, n, bd)
                    else:       # only the DC mode depends on the flags
                        exp = oracle.intra_modes_had(o, *lines, n, bd)
                        pred = oracle.intra_predict(1, lines[0], lines[1], n, bd, above, left, 1)
                        exp[1] = oracle.had((org, 3, n + 5), (pred, 0, n), n, n, bd)
                    assert np.array_equal(got, exp), (n, bd, kind, above, left)
