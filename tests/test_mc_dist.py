"""Motion-compensated prediction distortion: the compute part of xGetTemplateCost (AMVP candidate
cost: xPredInterBlk + SAD) and xGetInterPredictionError (merge candidate cost: MC + HADs) for uni-directional
prediction.  CPU: the oracle's xPredInterBlk restatement against the reference's own xPredInterBlk (driven through
oracle/ref_harness.cpp) and golden vectors; GPU: hmb200_mc_dist_batch against the oracle."""
import os

import numpy as np
import pytest

from common import PU_SIZES, MARGIN, ROOT, padded
from video_codecs_b200 import synth

W, H = 192, 128


def cases(rng, n):
    out = []
    for i in range(n):
        w, h = PU_SIZES[int(rng.integers(0, len(PU_SIZES)))]
        px, py = int(rng.integers(0, (W - w) // 4 + 1)) * 4, int(rng.integers(0, (H - h) // 4 + 1)) * 4
        mv = (int(rng.integers(-4 * 60, 4 * 60 + 1)), int(rng.integers(-4 * 60, 4 * 60 + 1)))
        if i % 5 == 0:
            mv = (mv[0] & ~3, mv[1])            # integer x
        if i % 7 == 0:
            mv = (mv[0], mv[1] & ~3)            # integer y
        out.append((px, py, w, h, mv, 0 if i % 2 else 2))
    return out


def frames(bd):
    f0 = synth.luma_frame(W, H, 0, seed=23, bit_depth=bd)
    f1 = synth.luma_frame(W, H, 1, seed=23, bit_depth=bd)
    cur, o0, stride = padded(f1)
    ref, _, _ = padded(f0)
    return f0, f1, cur, ref, o0, stride


@pytest.mark.parametrize("bd", [8, 10])
def test_oracle_mc_matches_reference(oracle, reference, bd):
    f0, f1, cur, ref, o0, stride = frames(bd)
    for (px, py, w, h, mv, kind) in cases(np.random.default_rng(3 + bd), 200):
        off = o0 + py * stride + px
        a = oracle.mc_dist(kind, (cur, off, stride), w, h, (ref, off, stride), mv, bd)
        b = reference.mc_dist(kind, (cur, off, stride), w, h, (ref, o0, stride), (W, H), MARGIN, (px, py), mv, bd)
        assert a == b, (px, py, w, h, mv, kind)


def test_oracle_mc_golden(oracle):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "mc_golden.npz"))
    for bd in (8, 10):
        f0, f1, cur, ref, o0, stride = frames(bd)
        assert np.array_equal(gold[f"frame0_{bd}"], f0)
        for c, e in zip(gold[f"cases_{bd}"], gold[f"expected_{bd}"]):
            px, py, w, h, mx, my, kind = [int(v) for v in c]
            off = o0 + py * stride + px
            assert oracle.mc_dist(kind, (cur, off, stride), w, h, (ref, off, stride), (mx, my), bd) == int(e)


@pytest.mark.gpu
@pytest.mark.parametrize("bd", [8, 10])
def test_gpu_mc_dist_vs_oracle(hm, oracle, bd):
    from video_codecs_b200 import MC_DESC_DTYPE, DF_SAD, DF_HADS
    f0, f1, cur, ref, o0, stride = frames(bd)
    if bd == 8:
        idc, idr = hm.register_plane_u8(f1, MARGIN, MARGIN, kind=0), hm.register_plane_u8(f0, MARGIN, MARGIN, kind=1)
    else:
        idc, idr = hm.register_plane(cur, W, H, MARGIN, MARGIN, bd, kind=0), hm.register_plane(ref, W, H, MARGIN, MARGIN, bd, kind=1)
    try:
        cs = cases(np.random.default_rng(77 + bd), 600)
        for kind, func in ((0, DF_SAD), (2, DF_HADS)):
            sel = [c for c in cs if c[5] == kind]
            descs = np.array([(px, py, w, h, mv[0], mv[1]) for (px, py, w, h, mv, _) in sel], dtype=MC_DESC_DTYPE)
            got = hm.mc_dist_batch(idc, idr, func, descs)
            exp = np.array([oracle.mc_dist(kind, (cur, o0 + py * stride + px, stride), w, h, (ref, o0 + py * stride + px, stride), mv, bd)
                            for (px, py, w, h, mv, _) in sel], dtype=np.uint32)
            assert np.array_equal(got, exp), kind
    finally:
        hm.release_plane(idc)
        hm.release_plane(idr)
