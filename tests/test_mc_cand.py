"""Merge / AMVP candidate costs (the rest of it): bi-prediction (xPredInterBlk with bi = true per list +
TComYuv::addAvg) next to the uni-directional prediction, the candidate loop of xMergeEstimation (error + getCost(bits), HADs)
and of xEstimateMvPredAMVP over xGetTemplateCost (SAD + calcRdCost).
CPU: the oracle's restatement against the reference's own xPredInterBlk / xWeightedAverage / calcRdCost (oracle/ref_harness.cpp)
and golden vectors from the reference; GPU: hmb200_mc_cand_dist_batch / hmb200_merge_estimation_batch /
hmb200_amvp_estimation_batch against the oracle."""
import os

import numpy as np
import pytest

from common import PU_SIZES, MARGIN, ROOT, padded
from video_codecs_b200 import synth

W, H = 192, 128
GOLD = os.path.join(ROOT, "tests", "golden", "mc_cand_golden.npz")


def frames(bd):
    f = [synth.luma_frame(W, H, t, seed=29, bit_depth=bd) for t in range(3)]
    pads = [padded(x) for x in f]
    return f, [p[0] for p in pads], pads[0][1], pads[0][2]


def bi_cases(rng, n):
    out = []
    for i in range(n):
        w, h = PU_SIZES[int(rng.integers(0, len(PU_SIZES)))]
        px, py = int(rng.integers(0, (W - w) // 4 + 1)) * 4, int(rng.integers(0, (H - h) // 4 + 1)) * 4
        mv0 = [int(rng.integers(-4 * 50, 4 * 50 + 1)), int(rng.integers(-4 * 50, 4 * 50 + 1))]
        mv1 = [int(rng.integers(-4 * 50, 4 * 50 + 1)), int(rng.integers(-4 * 50, 4 * 50 + 1))]
        if i % 5 == 0:
            mv0[0] &= ~3
        if i % 7 == 0:
            mv1[1] &= ~3
        if i % 11 == 0:
            mv0 = [mv0[0] & ~3, mv0[1] & ~3]
        if i % 13 == 0:
            mv1 = [mv1[0] & ~3, mv1[1] & ~3]
        out.append((px, py, w, h, mv0[0], mv0[1], mv1[0], mv1[1], 0 if i % 2 else 2))
    return out


@pytest.mark.parametrize("bd", [8, 10])
def test_oracle_bi_prediction_matches_reference(oracle, reference, bd):
    f, pads, o0, stride = frames(bd)
    cur, ra, rb = pads[1], pads[0], pads[2]
    for (px, py, w, h, ax, ay, bx, by, kind) in bi_cases(np.random.default_rng(5 + bd), 160):
        off = o0 + py * stride + px
        a = oracle.mc_cand_dist(kind, (cur, off, stride), w, h, 3, (ra, off, stride), (ax, ay), (rb, off, stride), (bx, by), bd)
        b = reference.mc_bi_dist(kind, (cur, off, stride), w, h, (ra, o0, stride), (rb, o0, stride), (W, H), MARGIN, (px, py), (ax, ay), (bx, by), bd)
        assert a == b, (px, py, w, h, ax, ay, bx, by, kind)


def test_oracle_bi_prediction_golden(oracle):
    gold = np.load(GOLD)
    for bd in (8, 10):
        f, pads, o0, stride = frames(bd)
        assert np.array_equal(gold[f"frame2_{bd}"], f[2])
        cur, ra, rb = pads[1], pads[0], pads[2]
        for c, e in zip(gold[f"cases_{bd}"], gold[f"expected_{bd}"]):
            px, py, w, h, ax, ay, bx, by, kind = [int(v) for v in c]
            off = o0 + py * stride + px
            assert oracle.mc_cand_dist(kind, (cur, off, stride), w, h, 3, (ra, off, stride), (ax, ay), (rb, off, stride), (bx, by), bd) == int(e)


def test_oracle_candidate_loops(oracle, reference):
    """The rate terms: xGetTemplateCost's (UInt) calcRdCost(bits, SAD, false, DF_SAD) against the reference's own TComRdCost, and
    the two argmin loops (strict '<' / first minimum, UInt wrap of getCost)."""
    rng = np.random.default_rng(1)
    for qp in (22, 35, 51):
        lam = 0.4624 * 2 ** ((qp - 12) / 3.0)
        L = int(np.floor(65536.0 * np.sqrt(lam)))
        for _ in range(50):
            bits, dist = int(rng.integers(0, 12)), int(rng.integers(0, 1 << 22))
            assert oracle.amvp_pick([dist], [bits], L)[1] == reference.template_rd_cost(bits, dist, lam)
    assert oracle.amvp_pick([10, 10, 9, 9], [1, 1, 1, 1], 0) == (2, 9)
    assert oracle.merge_pick([500, 400, 400], [1, 2, 3], 65536 * 100) == (0, 600)
    assert oracle.merge_pick([7, 7], [1, 1], 0xFFFFFFFF) == (0, 7 + (0xFFFFFFFF >> 16))
    assert oracle.merge_pick([7, 7], [1, 2], 0x80000001) == (1, 7)            # (lambda * 2) wraps in UInt to 2: the rate term vanishes


def _cands(rng, n_pu, planes, bd_unused=None):
    """n_pu PUs with 1..5 candidates each: a mix of list-0, list-1 and bi-directional candidates (merge), quarter-pel MVs."""
    from video_codecs_b200 import MC_CAND_DTYPE
    first, rows = [0], []
    for i in range(n_pu):
        w, h = PU_SIZES[int(rng.integers(0, len(PU_SIZES)))]
        px, py = int(rng.integers(0, (W - w) // 4 + 1)) * 4, int(rng.integers(0, (H - h) // 4 + 1)) * 4
        nc = int(rng.integers(1, 6))
        for k in range(nc):
            d = int(rng.integers(1, 4))
            mv0 = (int(rng.integers(-160, 161)), int(rng.integers(-160, 161)))
            mv1 = (int(rng.integers(-160, 161)), int(rng.integers(-160, 161)))
            if k == 1:
                mv0 = (mv0[0] & ~3, mv0[1] & ~3)
            r0, r1 = planes[int(rng.integers(0, len(planes)))], planes[int(rng.integers(0, len(planes)))]
            if d == 3 and k == 2:
                r1, mv1 = r0, mv0                                   # identical motion: predicted from list 0 alone
            bits = k + 1 - (1 if k == 4 else 0)                     # uiBitsCand with MaxNumMergeCand = 5
            rows.append((px, py, w, h, d, mv0[0], mv0[1], r0, mv1[0], mv1[1], r1, bits))
        first.append(len(rows))
    return np.array(first, dtype=np.int32), np.array(rows, dtype=MC_CAND_DTYPE)


@pytest.mark.gpu
@pytest.mark.parametrize("bd", [8, 10])
def test_gpu_candidate_costs_vs_oracle(hm, oracle, bd):
    from video_codecs_b200 import DF_SAD, DF_HADS
    f, pads, o0, stride = frames(bd)
    if bd == 8:
        ids = [hm.register_plane_u8(x, MARGIN, MARGIN, kind=1, poc=i) for i, x in enumerate(f)]
    else:
        ids = [hm.register_plane_u16(x, bd, MARGIN, MARGIN, kind=1, poc=i) for i, x in enumerate(f)]
    cur_id, cur = ids[1], pads[1]
    pad_of = {ids[0]: pads[0], ids[1]: pads[1], ids[2]: pads[2]}
    try:
        rng = np.random.default_rng(91 + bd)
        first, cands = _cands(rng, 120, [ids[0], ids[2]])

        def expect(kind):
            out = []
            for c in cands:
                off = o0 + int(c["pu_y"]) * stride + int(c["pu_x"])
                same = int(c["ref0_plane"] == c["ref1_plane"])
                out.append(oracle.mc_cand_dist(kind, (cur, off, stride), int(c["w"]), int(c["h"]), int(c["inter_dir"]),
                                               (pad_of[int(c["ref0_plane"])], off, stride), (int(c["mv0_x"]), int(c["mv0_y"])),
                                               (pad_of[int(c["ref1_plane"])], off, stride), (int(c["mv1_x"]), int(c["mv1_y"])), bd, same))
            return np.array(out, dtype=np.uint32)

        exp_had, exp_sad = expect(2), expect(0)
        assert np.array_equal(hm.mc_cand_dist_batch(cur_id, DF_HADS, cands), exp_had)
        assert np.array_equal(hm.mc_cand_dist_batch(cur_id, DF_SAD, cands), exp_sad)
        assert (cands["inter_dir"] == 3).sum() > 50 and len(np.unique(exp_had)) > 100
        # xMergeEstimation's loop, with and without HadamardME; a lambda that makes the rate term matter
        for use_had, exp in ((1, exp_had), (0, exp_sad)):
            lam = 65536 * 40
            best, cost, dist = hm.merge_estimation_batch(cur_id, first, cands, use_had, lam)
            assert np.array_equal(dist, exp)
            for i in range(len(first) - 1):
                a, b = first[i], first[i + 1]
                assert (int(best[i]), int(cost[i])) == oracle.merge_pick(exp[a:b], cands["bits"][a:b], lam), i
        # xEstimateMvPredAMVP's loop over xGetTemplateCost: uni-directional candidates, SAD + calcRdCost
        amvp = cands.copy()
        amvp["inter_dir"] = 1
        amvp["bits"] = 1
        exp_t = np.array([oracle.mc_cand_dist(0, (cur, o0 + int(c["pu_y"]) * stride + int(c["pu_x"]), stride), int(c["w"]), int(c["h"]), 1,
                                              (pad_of[int(c["ref0_plane"])], o0 + int(c["pu_y"]) * stride + int(c["pu_x"]), stride),
                                              (int(c["mv0_x"]), int(c["mv0_y"])),
                                              (pad_of[int(c["ref0_plane"])], o0, stride), (0, 0), bd) for c in amvp], dtype=np.uint32)
        L = 635000
        best, cost, dist = hm.amvp_estimation_batch(cur_id, first, amvp, L)
        assert np.array_equal(dist, exp_t)
        for i in range(len(first) - 1):
            a, b = first[i], first[i + 1]
            assert (int(best[i]), int(cost[i])) == oracle.amvp_pick(exp_t[a:b], amvp["bits"][a:b], L), i
        # error behaviour: a bi-directional AMVP candidate and an MV that leaves the padded plane are refused
        bad = amvp[:1].copy(); bad["inter_dir"] = 3
        with pytest.raises(Exception):
            hm.amvp_estimation_batch(cur_id, np.array([0, 1], dtype=np.int32), bad, L)
        far = cands[:1].copy(); far["inter_dir"] = 1; far["mv0_x"] = 4 * 1000
        with pytest.raises(Exception):
            hm.mc_cand_dist_batch(cur_id, DF_SAD, far)
    finally:
        for i in ids:
            hm.release_plane(i)
