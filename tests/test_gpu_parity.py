"""GPU suite: the CUDA path, called through the C-ABI (include/hmb200.h), against the oracle on the same seeded
inputs and against the golden vectors of the unmodified reference.  Integer work: the bar is bit-exact."""
import numpy as np
import pytest

from common import PU_SIZES, MARGIN, load_golden, padded, results_equal
from oracle.pyoracle import Oracle
from video_codecs_b200 import (FLAG_FEN, FLAG_HADME, FLAG_FRAC, DF_SAD, DF_SSE, DF_HADS, DF_SADS, JOB_DTYPE, DIST_DESC_DTYPE,
                               synth)

pytestmark = pytest.mark.gpu


def flags_of(fen, had, frac=True):
    return (FLAG_FEN if fen else 0) | (FLAG_HADME if had else 0) | (FLAG_FRAC if frac else 0)


@pytest.fixture(scope="module")
def gold():
    return load_golden()


# ---------------------------------------------------------------------------------------------------------------------
# distortion table (BASELINE config 2)
# ---------------------------------------------------------------------------------------------------------------------
def test_dist_single_call_golden(hm, gold):
    cases, exp = gold["dist_cases"], gold["dist_expected"]
    for (bd, kind, ss, w, h, oo, co), e in list(zip(cases, exp))[::3]:
        got = hm.dist(int(kind), (gold[f"dist_org_{bd}"], int(oo), 192), (gold[f"dist_cur_{bd}"], int(co), 192),
                      int(w), int(h), int(bd), int(ss))
        assert got == int(e), (bd, kind, ss, w, h)


@pytest.mark.parametrize("bd", [8, 10])
def test_dist_batch_every_pu_size(hm, oracle, bd):
    rng = np.random.default_rng(40 + bd)
    W, H = 320, 192
    hi = 1 << bd
    a = rng.integers(0, hi, size=(H, W)).astype(np.int16)
    b = np.clip(a.astype(np.int32) + rng.integers(-20, 21, size=a.shape), 0, hi - 1).astype(np.int16)
    pa, o0, stride = padded(a.astype(np.uint16))
    pb, _, _ = padded(b.astype(np.uint16))
    ida = hm.register_plane(pa, W, H, MARGIN, MARGIN, bd)
    idb = hm.register_plane(pb, W, H, MARGIN, MARGIN, bd)
    try:
        for func in (DF_SAD, DF_SSE, DF_HADS, DF_SADS):
            descs, expect = [], []
            for (w, h) in PU_SIZES:
                for rep in range(40):
                    ox, oy = int(rng.integers(-8, W - w + 8)), int(rng.integers(-8, H - h + 8))
                    cx, cy = int(rng.integers(-8, W - w + 8)), int(rng.integers(-8, H - h + 8))
                    ss = int(rng.integers(0, 2)) if func in (DF_SAD, DF_SADS) else 0
                    descs.append((ida, ox, oy, idb, cx, cy, w, h, ss, 0))
                    expect.append(oracle.dist(func, (pa, o0 + oy * stride + ox, stride), (pb, o0 + cy * stride + cx, stride), w, h, bd, ss))
            got = hm.dist_batch(func, bd, np.array(descs, dtype=DIST_DESC_DTYPE))
            assert np.array_equal(got, np.array(expect, dtype=np.uint32)), func
    finally:
        hm.release_plane(ida)
        hm.release_plane(idb)


@pytest.mark.parametrize("bd", [8, 10])
def test_dist_table_ten_thousand_blocks_per_size(hm, oracle, bd):
    """BASELINE.json configs[1] at the size SURVEY.md 8d asks for: 10^4 random block pairs for each of the 24 PU sizes, every
    family (SAD with random iSubShift, SSE, HADs), 8- and 10-bit, random offsets incl. the margins - bit-exact."""
    rng = np.random.default_rng(140 + bd)
    W, H, N = 640, 384, 10000
    hi = 1 << bd
    a = rng.integers(0, hi, size=(H, W)).astype(np.int16)
    b = np.clip(a.astype(np.int32) + rng.integers(-40, 41, size=a.shape), 0, hi - 1).astype(np.int16)
    pa, o0, stride = padded(a.astype(np.uint16))
    pb, _, _ = padded(b.astype(np.uint16))
    ida = hm.register_plane(pa, W, H, MARGIN, MARGIN, bd)
    idb = hm.register_plane(pb, W, H, MARGIN, MARGIN, bd)
    try:
        blocks = np.zeros((len(PU_SIZES) * N, 7), dtype=np.int32)
        for i, (w, h) in enumerate(PU_SIZES):
            v = blocks[i * N:(i + 1) * N]
            v[:, 0], v[:, 1] = rng.integers(-8, W - w + 8, N), rng.integers(-8, H - h + 8, N)
            v[:, 2], v[:, 3] = rng.integers(-8, W - w + 8, N), rng.integers(-8, H - h + 8, N)
            v[:, 4], v[:, 5], v[:, 6] = w, h, rng.integers(0, 2, N)
        descs = np.zeros(len(blocks), dtype=DIST_DESC_DTYPE)
        descs["org_plane"], descs["cur_plane"] = ida, idb
        for f, c in (("org_x", 0), ("org_y", 1), ("cur_x", 2), ("cur_y", 3), ("w", 4), ("h", 5)):
            descs[f] = blocks[:, c]
        for func in (DF_SAD, DF_SSE, DF_HADS):
            blk = blocks.copy()
            if func != DF_SAD:
                blk[:, 6] = 0
            descs["sub_shift"] = blk[:, 6]
            got = hm.dist_batch(func, bd, descs)
            exp = oracle.dist_batch(func, (pa, o0, stride), (pb, o0, stride), blk, bd)
            assert np.array_equal(got, exp), func
    finally:
        hm.release_plane(ida)
        hm.release_plane(idb)


@pytest.mark.parametrize("bd", [8, 10])
def test_dist_generic_shapes_2x2_hadamard_and_unsized_sad(hm, oracle, bd):
    """Shapes outside the 24 PU sizes: xGetHADs falls back to 2x2 tiles when a dimension is not a multiple of 4
    (xCalcHADs2x2, TComRdCost.cpp:1310-1330, 1573-1586) and the generic xGetSAD (:461-487) ignores iSubShift for widths
    without an unrolled variant; both through the batched entry (registered planes) and the 1:1 FpDistFunc-style call."""
    rng = np.random.default_rng(11 + bd)
    W, H = 128, 96
    hi = 1 << bd
    a = rng.integers(0, hi, size=(H, W)).astype(np.int16)
    b = rng.integers(0, hi, size=(H, W)).astype(np.int16)
    pa, o0, stride = padded(a.astype(np.uint16))
    pb, _, _ = padded(b.astype(np.uint16))
    shapes = [(2, 2), (6, 2), (2, 6), (6, 6), (10, 4), (4, 10), (20, 10), (2, 8), (36, 18), (64, 2), (6, 64), (20, 16), (40, 8), (36, 36)]
    ida = hm.register_plane(pa, W, H, MARGIN, MARGIN, bd)
    idb = hm.register_plane(pb, W, H, MARGIN, MARGIN, bd)
    try:
        for func in (DF_HADS, DF_SAD, DF_SSE):
            blocks = []
            for (w, h) in shapes:
                for rep in range(25):
                    blocks.append((int(rng.integers(0, W - w)), int(rng.integers(0, H - h)), int(rng.integers(0, W - w)), int(rng.integers(0, H - h)),
                                   w, h, int(rng.integers(0, 2)) if func == DF_SAD else 0))
            blocks = np.array(blocks, dtype=np.int32)
            descs = np.zeros(len(blocks), dtype=DIST_DESC_DTYPE)
            descs["org_plane"], descs["cur_plane"] = ida, idb
            for f, c in (("org_x", 0), ("org_y", 1), ("cur_x", 2), ("cur_y", 3), ("w", 4), ("h", 5), ("sub_shift", 6)):
                descs[f] = blocks[:, c]
            exp = oracle.dist_batch(func, (pa, o0, stride), (pb, o0, stride), blocks, bd)
            assert np.array_equal(hm.dist_batch(func, bd, descs), exp), func
            for k in range(0, len(blocks), 9):         # the same evaluations through hmb200_dist (host Pel* in, one call each)
                ox, oy, cx, cy, w, h, ss = [int(v) for v in blocks[k]]
                got = hm.dist(func, (pa, o0 + oy * stride + ox, stride), (pb, o0 + cy * stride + cx, stride), w, h, bd, ss)
                assert got == int(exp[k]), (func, w, h)
    finally:
        hm.release_plane(ida)
        hm.release_plane(idb)


def test_dist_signed_pattern(hm, oracle):
    rng = np.random.default_rng(6)
    org = rng.integers(-255, 511, size=64 * 64).astype(np.int16)
    cur = rng.integers(0, 256, size=64 * 64).astype(np.int16)
    for (w, h) in ((8, 8), (64, 64), (12, 16), (4, 8)):
        for func in (DF_SAD, DF_SSE, DF_HADS):
            assert hm.dist(func, (org, 0, 64), (cur, 0, 64), w, h, 8, 0) == oracle.dist(func, (org, 0, 64), (cur, 0, 64), w, h, 8, 0)


# ---------------------------------------------------------------------------------------------------------------------
# plane ingest
# ---------------------------------------------------------------------------------------------------------------------
def test_plane_round_trip_and_border_extension(hm, oracle):
    f = synth.luma_frame(200, 72, 3, seed=2)
    pid = hm.register_plane_u8(f, 80, 80)
    try:
        back = hm.read_plane(pid, 200, 72, 80, 80)
        assert np.array_equal(back, synth.pad_plane(f, 80, 80))
        manual = np.zeros((72 + 160, 200 + 160), dtype=np.int16)
        manual[80:-80, 80:-80] = f
        oracle.extend_border(manual, 80 * manual.shape[1] + 80, manual.shape[1], 200, 72, 80, 80)
        assert np.array_equal(back, manual)
    finally:
        hm.release_plane(pid)
    f10 = synth.luma_frame(64, 64, 0, seed=3, bit_depth=10)
    p10 = synth.pad_plane(f10, 16, 24)
    pid = hm.register_plane(p10, 64, 64, 16, 24, 10)
    try:
        assert np.array_equal(hm.read_plane(pid, 64, 64, 16, 24), p10)
    finally:
        hm.release_plane(pid)


# ---------------------------------------------------------------------------------------------------------------------
# batched search against the reference's golden vectors and against the oracle
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bd", [8, 10])
@pytest.mark.parametrize("fen,had", [(1, 1), (0, 0), (1, 0), (0, 1)])
def test_me_jobs_golden(hm, gold, bd, fen, had):
    tag = f"search_bd{bd}_fen{fen}_had{had}"
    jobs, exp = gold[tag + "_jobs"], gold[tag + "_results"]
    f0, f1 = gold[f"frame0_{bd}"], gold[f"frame1_{bd}"]
    H, W = f0.shape
    cur, _, _ = padded(f1)
    ref, _, _ = padded(f0)
    idc = hm.register_plane(cur, W, H, MARGIN, MARGIN, bd, kind=0)
    idr = hm.register_plane(ref, W, H, MARGIN, MARGIN, bd, kind=1)
    try:
        got = hm.me_jobs(idc, idr, jobs, flags_of(fen, had))
        assert results_equal(got, exp) == []
    finally:
        hm.release_plane(idc)
        hm.release_plane(idr)


@pytest.mark.parametrize("fen", [1, 0])
def test_me_canonical_ctus_vs_oracle(hm, fen):
    """Every PU of a few CTUs (interior, picture corner, last partial row) of a 416x240 pair, +-64, with refinement."""
    W, H = 416, 240
    f0 = synth.luma_frame(W, H, 0)
    f1 = synth.luma_frame(W, H, 1)
    cur, o0, stride = padded(f1)
    ref, _, _ = padded(f0)
    lam = int(np.floor(65536.0 * np.sqrt(0.4624 * 2 ** ((35 - 12) / 3.0))))
    jobs = np.concatenate([hm.build_canonical_jobs(W, H, 64, lam, ctu_first=c, ctu_count=1) for c in (0, 9, 27)])
    jobs = jobs[:: 3 if fen else 5]
    idc = hm.register_plane_u8(f1, MARGIN, MARGIN, kind=0)
    idr = hm.register_plane_u8(f0, MARGIN, MARGIN, kind=1)
    try:
        got = hm.me_jobs(idc, idr, jobs, flags_of(fen, 1))
    finally:
        hm.release_plane(idc)
        hm.release_plane(idr)
    exp, _ = Oracle(fen=fen, hadme=1).run_jobs((cur, o0, stride), (ref, o0, stride), jobs, 8, True)
    assert results_equal(got, exp) == []
    assert len(set(zip(got["mv_x"].tolist(), got["mv_y"].tolist()))) >= 2     # the MV field is not trivial


def test_me_random_predictors_and_windows(hm):
    """Per-PU predictors and asymmetric, clipped, tiny and single-candidate windows."""
    W, H = 256, 192
    rng = np.random.default_rng(12)
    f0 = synth.luma_frame(W, H, 0, seed=5)
    f1 = synth.luma_frame(W, H, 2, seed=5)
    cur, o0, stride = padded(f1)
    ref, _, _ = padded(f0)
    jobs = np.zeros(len(PU_SIZES) * 4, dtype=JOB_DTYPE)
    for i, (w, h) in enumerate(PU_SIZES * 4):
        px, py = int(rng.integers(0, (W - w) // 4 + 1)) * 4, int(rng.integers(0, (H - h) // 4 + 1)) * 4
        pred = (int(rng.integers(-60, 61)), int(rng.integers(-60, 61)))
        kind = i % 4
        if kind == 0:      # HM-style window
            sr = int(rng.choice([7, 16, 33]))
            lt = (max((pred[0] >> 2) - sr, -(px + 64 + 7)), max((pred[1] >> 2) - sr, -(py + 64 + 7)))
            rb = (min((pred[0] >> 2) + sr, W + 7 - px), min((pred[1] >> 2) + sr, H + 7 - py))
        elif kind == 1:    # single candidate
            lt = rb = (int(rng.integers(-5, 6)), int(rng.integers(-5, 6)))
        elif kind == 2:    # one row / one column
            lt = (-9, 3)
            rb = (22, 3) if i % 8 == 2 else (-9, 30)
        else:              # wide and flat / odd sizes
            lt = (-int(rng.integers(1, 70)), -int(rng.integers(1, 6)))
            rb = (int(rng.integers(1, 70)), int(rng.integers(1, 6)))
        jobs[i] = (px, py, w, h, lt[0], lt[1], rb[0], rb[1], pred[0], pred[1], int(rng.integers(0, 9000000)), 0)
    idc = hm.register_plane_u8(f1, MARGIN, MARGIN, kind=0)
    idr = hm.register_plane_u8(f0, MARGIN, MARGIN, kind=1)
    try:
        for fen, had in ((1, 1), (0, 0)):
            got = hm.me_jobs(idc, idr, jobs, flags_of(fen, had))
            exp, _ = Oracle(fen=fen, hadme=had).run_jobs((cur, o0, stride), (ref, o0, stride), jobs, 8, True)
            assert results_equal(got, exp) == [], (fen, had)
    finally:
        hm.release_plane(idc)
        hm.release_plane(idr)


def test_tie_break_flat_content(hm):
    """All SADs tie: the winner is decided by the MV cost and then by raster order (strict '<', first wins)."""
    W, H = 128, 128
    flat = np.full((H, W), 77, dtype=np.uint8)
    cur, o0, stride = padded(flat)
    jobs = np.zeros(6, dtype=JOB_DTYPE)
    jobs[0] = (32, 32, 8, 8, -64, -64, 64, 64, 0, 0, 0, 0)                # lambda 0: everything ties -> (-64,-64)
    jobs[1] = (32, 32, 16, 16, -64, -64, 64, 64, 9, -6, 65536 * 4, 0)
    jobs[2] = (0, 0, 64, 64, -64, -64, 64, 64, 2, 2, 65536 * 4, 0)
    jobs[3] = (64, 64, 4, 8, -3, -3, 3, 3, 1, 1, 4037017, 0)
    jobs[4] = (64, 64, 32, 32, -64, -64, 64, 64, -255, 255, 0xFFFFFFFF, 0)   # 32-bit wrap of m_uiCost * bits
    jobs[5] = (64, 0, 64, 32, -64, -64, 64, 64, 0, 0, 1, 0)
    idc = hm.register_plane_u8(flat, MARGIN, MARGIN)
    try:
        got = hm.me_jobs(idc, idc, jobs, flags_of(1, 1))
    finally:
        hm.release_plane(idc)
    exp, _ = Oracle(fen=1, hadme=1).run_jobs((cur, o0, stride), (cur, o0, stride), jobs, 8, True)
    assert results_equal(got, exp) == []
    assert (got[0]["mv_x"], got[0]["mv_y"]) == (-64, -64)


# ---------------------------------------------------------------------------------------------------------------------
# 1:1 entries (the TEncSearch signatures)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bd", [8, 10])
def test_pattern_search_and_frac_one_to_one(hm, bd):
    W, H = 192, 128
    f0 = synth.luma_frame(W, H, 0, seed=31, bit_depth=bd)
    f1 = synth.luma_frame(W, H, 1, seed=31, bit_depth=bd)
    ref, o0, stride = padded(f0)
    O = Oracle(fen=1, hadme=1)
    rng = np.random.default_rng(8)
    idr = hm.register_plane(ref, W, H, MARGIN, MARGIN, bd)
    try:
        for (w, h) in PU_SIZES[::2]:
            # CU-local original buffer with its own stride, as TEncCu hands it over (stride 64 >> depth)
            px, py = int(rng.integers(0, (W - w) // 4 + 1)) * 4, int(rng.integers(0, (H - h) // 4 + 1)) * 4
            org = np.ascontiguousarray(f1[py:py + h, px:px + w].astype(np.int16))
            if w == 16:     # bi-pred style pattern: 2*org - pred may leave the pixel range
                org = (2 * org - rng.integers(0, 1 << bd, size=org.shape)).astype(np.int16)
            pred = (int(rng.integers(-20, 21)), int(rng.integers(-20, 21)))
            lam = int(rng.integers(100000, 6000000))
            roff = o0 + py * stride + px
            lt, rb = (-12 + (pred[0] >> 2), -9 + (pred[1] >> 2)), (12 + (pred[0] >> 2), 9 + (pred[1] >> 2))
            mv, sad = hm.pattern_search((org, 0, w), w, h, (ref, roff, stride), lt, rb, lam, pred, bd, FLAG_FEN)
            assert (mv, sad) == O.pattern_search((org, 0, w), w, h, (ref, roff, stride), lt, rb, lam, pred, bd)
            got = hm.pattern_search_frac((org, 0, w), w, h, (ref, roff, stride), mv, lam, pred, bd, FLAG_HADME)
            assert got == O.pattern_search_frac((org, 0, w), w, h, (ref, roff, stride), mv, lam, pred, bd)
    finally:
        hm.release_plane(idr)


@pytest.mark.parametrize("bd", [8, 10])
def test_one_pu_kernels_every_size_full_window(hm, bd):
    """The 1:1 entries' own kernels (hmb200_one.cuh: a CTA per candidate row, byte-SIMD when pattern and plane are 8-bit, the
    last CTA decodes; refinement with a thread per tile column; result reported through mapped host memory): every PU size,
    +-64 windows at every byte alignment, FEN on / off, Hadamard and SAD refinement, a pattern that leaves the sample range
    (scalar path on an 8-bit plane), the fused search + refinement entry, and flat content (the first candidate in raster
    order must win) - against the oracle."""
    W, H = 416, 240
    f0 = synth.luma_frame(W, H, 0, seed=52, bit_depth=bd)
    f1 = synth.luma_frame(W, H, 1, seed=52, bit_depth=bd)
    ref, o0, stride = padded(f0)
    rng = np.random.default_rng(19 + bd)
    idr = hm.register_plane(ref, W, H, MARGIN, MARGIN, bd)
    try:
        for i, (w, h) in enumerate(PU_SIZES):
            fen, had = i & 1, (i >> 1) & 1
            O = Oracle(fen=fen, hadme=had)
            px, py = int(rng.integers(8, (W - w - 8) // 4)) * 4, int(rng.integers(4, (H - h - 8) // 4)) * 4
            org = np.ascontiguousarray(f1[py:py + h, px:px + w].astype(np.int16))
            if i % 5 == 4:
                org = (2 * org - rng.integers(0, 1 << bd, size=org.shape)).astype(np.int16)
            pred = (int(rng.integers(-40, 41)), int(rng.integers(-40, 41)))
            lam = int(rng.integers(100000, 6000000))
            roff = o0 + py * stride + px
            # +-64 around the integer predictor, clipped so that window + refinement stay inside the padded plane
            cx, cy = pred[0] >> 2, pred[1] >> 2
            lt = (max(cx - 64, -px - MARGIN + 8 + i % 4), max(cy - 64, -py - MARGIN + 8))
            rb = (min(cx + 64, W - px - w + MARGIN - 8), min(cy + 64, H - py - h + MARGIN - 8))
            flags = (FLAG_FEN if fen else 0) | (FLAG_HADME if had else 0)
            mv, sad = hm.pattern_search((org, 0, w), w, h, (ref, roff, stride), lt, rb, lam, pred, bd, flags)
            assert (mv, sad) == O.pattern_search((org, 0, w), w, h, (ref, roff, stride), lt, rb, lam, pred, bd), (w, h)
            exp = O.pattern_search_frac((org, 0, w), w, h, (ref, roff, stride), mv, lam, pred, bd)
            assert hm.pattern_search_frac((org, 0, w), w, h, (ref, roff, stride), mv, lam, pred, bd, flags) == exp, (w, h)
            both = hm.pattern_search_and_refine((org, 0, w), w, h, (ref, roff, stride), lt, rb, lam, pred, bd, flags)
            assert both == (mv, sad) + exp, (w, h)
        # flat content: every candidate has the same SAD; with lambda 0 the first one in raster order wins
        flat = np.full((H, W), 77, dtype=np.uint16 if bd > 8 else np.uint8)
        pf, _, _ = padded(flat)
        idf = hm.register_plane(pf, W, H, MARGIN, MARGIN, bd)
        try:
            org = np.full((16, 16), 70, dtype=np.int16)
            roff = o0 + 96 * stride + 160
            mv, sad = hm.pattern_search((org, 0, 16), 16, 16, (pf, roff, stride), (-61, -64), (64, 63), 0, (0, 0), bd, FLAG_FEN)
            assert (mv, sad) == ((-61, -64), (7 * 256) >> (bd - 8))
        finally:
            hm.release_plane(idf)
    finally:
        hm.release_plane(idr)


def test_one_pu_kernels_equal_round1_kernels(hm, monkeypatch):
    """The same 1:1 calls through a second context created with HMB200_NO_ONE_FAST=1 (round 1's kernels: k_search_split +
    finalize + k_frac_generic, D2H copy + stream synchronisation): vectors, SADs and refinement costs must be identical -
    including sizes whose pattern travels in the kernel arguments and sizes that go through the record."""
    W, H = 416, 240
    f0, f1 = synth.luma_frame(W, H, 0, seed=66), synth.luma_frame(W, H, 1, seed=66)
    ref, o0, stride = padded(f0)
    rng = np.random.default_rng(4)
    calls = []
    for (w, h) in PU_SIZES[::3] + [(8, 8), (16, 16)]:
        px, py = int(rng.integers(8, (W - w - 8) // 4)) * 4, int(rng.integers(4, (H - h - 8) // 4)) * 4
        org = np.ascontiguousarray(f1[py:py + h, px:px + w].astype(np.int16))
        pred = (int(rng.integers(-30, 31)), int(rng.integers(-30, 31)))
        calls.append((org, w, h, o0 + py * stride + px, pred, int(rng.integers(100000, 6000000))))

    def run_all():
        idr = hm.register_plane(ref, W, H, MARGIN, MARGIN, 8)
        try:
            out = []
            for org, w, h, roff, pred, lam in calls:
                lt, rb = (-40 + (pred[0] >> 2), -40 + (pred[1] >> 2)), (40 + (pred[0] >> 2), 40 + (pred[1] >> 2))
                both = hm.pattern_search_and_refine((org, 0, w), w, h, (ref, roff, stride), lt, rb, lam, pred, 8, FLAG_FEN | FLAG_HADME)
                mv, sad = hm.pattern_search((org, 0, w), w, h, (ref, roff, stride), lt, rb, lam, pred, 8, FLAG_FEN)
                fr = hm.pattern_search_frac((org, 0, w), w, h, (ref, roff, stride), mv, lam, pred, 8, FLAG_HADME)
                assert both == (mv, sad) + fr
                out.append(both)
            return out
        finally:
            hm.release_plane(idr)

    fast = run_all()
    monkeypatch.setenv("HMB200_NO_ONE_FAST", "1")
    ctx = hm.ctx_create(0)                       # the knob is read when a context is created; the new context is current
    try:
        slow = run_all()
    finally:
        hm.ctx_destroy(ctx)
        hm.ctx_set_current(None)
        monkeypatch.delenv("HMB200_NO_ONE_FAST")
    assert fast == slow


def _cu_parts(S):
    q, h = S // 4, S // 2
    parts = [(0, 0, S, S), (0, 0, S, h), (0, h, S, h), (0, 0, h, S), (h, 0, h, S)]
    if S > 8:
        parts += [(0, 0, S, q), (0, q, S, S - q), (0, 0, S, S - q), (0, S - q, S, q),
                  (0, 0, q, S), (q, 0, S - q, S), (0, 0, S - q, S), (S - q, 0, q, S)]
    if S == 16:      # HM goes on with the four 8x8 child CUs; the 16x16 launch carries their PUs as well
        for cy in (0, 8):
            for cx in (0, 8):
                parts += [(cx + ox, cy + oy, w, h) for (ox, oy, w, h) in _cu_parts(8)]
    return parts


@pytest.mark.parametrize("bd", [8, 10])
@pytest.mark.parametrize("fen,had", [(1, 1), (0, 1), (1, 0)])
def test_whole_cu_speculation_is_exact(hm, fen, had, bd):
    """hmb200_pattern_search_and_refine in HM's call order: the 2Nx2N PU of a CU, then its 2NxN / Nx2N / AMP partitions.  The 2Nx2N
    call searches and refines the whole CU; the following calls are answered from it when window, predictor, lambda, position and
    pattern samples match - every answer must equal the oracle's for that PU.  Calls that do not match (another predictor, a changed
    pattern) must take the per-PU path and be exact as well; the counters show which path ran.  8-bit planes run the byte-SIMD
    kernels, 10-bit planes the scalar 16-bit ones."""
    W, H = 416, 240
    f0, f1 = synth.luma_frame(W, H, 0, seed=91, bit_depth=bd), synth.luma_frame(W, H, 1, seed=91, bit_depth=bd)
    ref, o0, stride = padded(f0)
    O = Oracle(fen=fen, hadme=had)
    flags = (FLAG_FEN if fen else 0) | (FLAG_HADME if had else 0)
    rng = np.random.default_rng(6 + fen + 2 * had)
    idr = hm.register_plane(ref, W, H, MARGIN, MARGIN, bd)
    try:
        for S, (cx, cy) in ((8, (72, 40)), (16, (160, 96)), (32, (224, 128)), (64, (128, 64)), (16, (400, 224)), (64, (0, 0))):
            cu = np.ascontiguousarray(f1[cy:cy + S, cx:cx + S].astype(np.int16))        # the CU's original block, stride S (TEncCu's buffer)
            pred = (int(rng.integers(-30, 31)), int(rng.integers(-30, 31)))
            lam = int(rng.integers(100000, 6000000))
            px, py = pred[0] >> 2, pred[1] >> 2
            lt = (max(px - 64, -cx - MARGIN + 8), max(py - 64, -cy - MARGIN + 8))
            rb = (min(px + 64, W - cx - S + MARGIN - 8), min(py + 64, H - cy - S + MARGIN - 8))
            c0, l0, h0 = hm.one_call_stats()
            for (ox, oy, w, h) in _cu_parts(S):
                roff = o0 + (cy + oy) * stride + cx + ox
                got = hm.pattern_search_and_refine((cu, oy * S + ox, S), w, h, (ref, roff, stride), lt, rb, lam, pred, bd, flags)
                org = np.ascontiguousarray(cu[oy:oy + h, ox:ox + w])
                mv, sad = O.pattern_search((org, 0, w), w, h, (ref, roff, stride), lt, rb, lam, pred, bd)
                assert got == (mv, sad) + O.pattern_search_frac((org, 0, w), w, h, (ref, roff, stride), mv, lam, pred, bd), (S, ox, oy, w, h)
            c1, l1, h1 = hm.one_call_stats()
            n = len(_cu_parts(S))
            assert (c1 - c0, l1 - l0, h1 - h0) == (n, 1, n - 1), (S, c1 - c0, l1 - l0, h1 - h0)
            # a partition with its own predictor, and one whose pattern changed since the 2Nx2N call: per-PU path, still exact
            ox, oy, w, h = _cu_parts(S)[2]
            roff = o0 + (cy + oy) * stride + cx + ox
            org = np.ascontiguousarray(cu[oy:oy + h, ox:ox + w])
            pred2 = (pred[0] + 3, pred[1] - 2)
            got = hm.pattern_search_and_refine((cu, oy * S + ox, S), w, h, (ref, roff, stride), lt, rb, lam, pred2, bd, flags)
            mv, sad = O.pattern_search((org, 0, w), w, h, (ref, roff, stride), lt, rb, lam, pred2, bd)
            assert got == (mv, sad) + O.pattern_search_frac((org, 0, w), w, h, (ref, roff, stride), mv, lam, pred2, bd)
            cu[oy + 1, ox + 1] ^= 0x3f
            org = np.ascontiguousarray(cu[oy:oy + h, ox:ox + w])
            got = hm.pattern_search_and_refine((cu, oy * S + ox, S), w, h, (ref, roff, stride), lt, rb, lam, pred, bd, flags)
            mv, sad = O.pattern_search((org, 0, w), w, h, (ref, roff, stride), lt, rb, lam, pred, bd)
            assert got == (mv, sad) + O.pattern_search_frac((org, 0, w), w, h, (ref, roff, stride), mv, lam, pred, bd)
            assert hm.one_call_stats()[2] == h1                  # neither was answered from the CU launch
    finally:
        hm.release_plane(idr)


def test_error_behaviour(hm):
    from video_codecs_b200 import HMB200Error
    with pytest.raises(HMB200Error):
        hm.me_jobs(991, 992, np.zeros(1, dtype=JOB_DTYPE))
    flat = np.zeros((64, 64), dtype=np.uint8)
    pid = hm.register_plane_u8(flat, 80, 80)
    try:
        bad = np.zeros(1, dtype=JOB_DTYPE)
        bad[0] = (0, 0, 5, 8, -1, -1, 1, 1, 0, 0, 0, 0)        # width 5 is not a PU size
        with pytest.raises(HMB200Error):
            hm.me_jobs(pid, pid, bad)
        bad[0] = (0, 0, 8, 8, 2, 0, 1, 0, 0, 0, 0, 0)          # empty window
        with pytest.raises(HMB200Error):
            hm.me_jobs(pid, pid, bad)
        assert len(hm.me_jobs(pid, pid, np.zeros(0, dtype=JOB_DTYPE))) == 0      # empty job list is fine
    finally:
        hm.release_plane(pid)


# ---------------------------------------------------------------------------------------------------------------------
# full-size, size-independent properties (BASELINE.json configs[2] / [4] geometry; the oracle only samples)
# ---------------------------------------------------------------------------------------------------------------------
def _run_full(hm, cur, ref, jobs, flags):
    idc = hm.register_plane_u8(cur, MARGIN, MARGIN, kind=0)
    idr = hm.register_plane_u8(ref, MARGIN, MARGIN, kind=1)
    try:
        prep = hm.prepare_jobs(jobs, flags, 8)
        prep.run(idc, idr)
        out = prep.fetch()
        work = prep.work()
        prep.free()
        return out, work
    finally:
        hm.release_plane(idc)
        hm.release_plane(idr)


def test_pipelined_fetch_equals_blocking_fetch(hm):
    """hmb200_fetch_results_async / hmb200_fetch_wait with two alternating prepared handles (what bench.py's end-to-end
    leg does): every frame pair's MV field equals the one a blocking run + fetch gives; a handle that is run again while
    its fetch is pending is ordered after the copy."""
    from video_codecs_b200 import RESULT_DTYPE
    W, H = 256, 192
    frames = [synth.luma_frame(W, H, t, seed=31) for t in range(4)]
    lam = int(np.floor(65536.0 * np.sqrt(0.4624 * 2 ** ((35 - 12) / 3.0))))
    jobs = hm.build_canonical_jobs(W, H, 64, lam)[::5]
    flags = flags_of(1, 1)
    preps = [hm.prepare_jobs(jobs, flags, 8), hm.prepare_jobs(jobs, flags, 8)]
    outs = [hm.host_array(len(jobs), RESULT_DTYPE), hm.host_array(len(jobs), RESULT_DTYPE)]
    pinned = []
    for f in frames:
        a = hm.host_array(f.size, np.uint8).reshape(f.shape)
        a[...] = f
        pinned.append(a)
    got = []
    try:
        for k in range(6):
            p, o = preps[k % 2], outs[k % 2]
            p.fetch_wait()
            if k >= 2:
                got.append(o.copy())
            idc = hm.register_plane_u8(pinned[(k + 1) % 4], MARGIN, MARGIN, kind=0)
            idr = hm.register_plane_u8(pinned[k % 4], MARGIN, MARGIN, kind=1)
            p.run(idc, idr)
            p.fetch_async(o)
            hm.release_plane(idc)
            hm.release_plane(idr)
        for k in (4, 5):
            preps[k % 2].fetch_wait()
            got.append(outs[k % 2].copy())
        for k in range(6):
            exp, _ = _run_full(hm, frames[(k + 1) % 4], frames[k % 4], jobs, flags)
            assert results_equal(got[k], exp) == [], k
        with pytest.raises(Exception):
            preps[0].fetch_async(np.zeros(len(jobs), dtype=RESULT_DTYPE))      # pageable memory is refused
    finally:
        for p in preps:
            p.free()


def test_1080p_fused_path_equals_per_pu_path_and_oracle_samples(hm, monkeypatch):
    """The CU-fused kernels (one pass per CU for all 13 partitions) and the per-PU kernels must give the same MV field,
    SADs and refinement for the whole 1080p canonical list; 400 sampled PUs are also checked against the oracle."""
    W, H = 1920, 1088
    f0, f1 = synth.luma_frame(W, H, 0, seed=21), synth.luma_frame(W, H, 1, seed=21)
    lam = int(np.floor(65536.0 * np.sqrt(0.4624 * 2 ** ((35 - 12) / 3.0))))
    jobs = hm.build_canonical_jobs(W, H, 64, lam)
    assert len(jobs) == 302430
    fused, work = _run_full(hm, f1, f0, jobs, flags_of(1, 1))
    assert work["pus_fused"] == len(jobs) and work["abs_diffs_executed"] < work["abs_diffs"]
    monkeypatch.setenv("HMB200_NO_CU_FUSION", "1")
    plain, work2 = _run_full(hm, f1, f0, jobs, flags_of(1, 1))
    monkeypatch.delenv("HMB200_NO_CU_FUSION")
    assert work2["pus_fused"] == 0 and work2["abs_diffs_executed"] == work2["abs_diffs"] == work["abs_diffs"]
    assert results_equal(fused, plain) == []
    # every MV lies in its window
    assert np.all((fused["mv_x"] >= jobs["lt_x"]) & (fused["mv_x"] <= jobs["rb_x"]) &
                  (fused["mv_y"] >= jobs["lt_y"]) & (fused["mv_y"] <= jobs["rb_y"]))
    rng = np.random.default_rng(3)
    pick = np.sort(rng.choice(len(jobs), 400, replace=False))
    cur, o0, stride = padded(f1)
    ref, _, _ = padded(f0)
    exp, _ = Oracle(fen=1, hadme=1).run_jobs((cur, o0, stride), (ref, o0, stride), jobs[pick], 8, True)
    assert results_equal(fused[pick], exp) == []


def test_2160p_translation_property(hm):
    """cur(x, y) = ref(x + 7, y - 5): every PU whose displaced block stays inside the picture must find MV (7, -5) with
    SAD 0 and keep it through the refinement (half = quarter = 0); 3840x2160, the largest geometry of BASELINE.json."""
    W, H, dx, dy = 3840, 2160, 7, -5
    big = synth.luma_frame(W + 32, H + 32, 0, seed=5, n_rect=0)
    ref = np.ascontiguousarray(big[16:16 + H, 16:16 + W])
    cur = np.ascontiguousarray(big[16 + dy:16 + dy + H, 16 + dx:16 + dx + W])
    jobs = hm.build_canonical_jobs(W, H, 64, 40000)
    res, work = _run_full(hm, cur, ref, jobs, flags_of(1, 1))
    inside = (jobs["pu_x"] + dx >= 0) & (jobs["pu_y"] + dy >= 0) & (jobs["pu_x"] + jobs["w"] + dx <= W) & \
             (jobs["pu_y"] + jobs["h"] + dy <= H)
    assert inside.sum() > 0.97 * len(jobs)
    r = res[inside]
    assert np.all(r["mv_x"] == dx) and np.all(r["mv_y"] == dy) and np.all(r["sad"] == 0)
    assert np.all(r["half_x"] == 0) and np.all(r["half_y"] == 0) and np.all(r["qter_x"] == 0) and np.all(r["qter_y"] == 0)
    assert work["cand_sads"] > 1.9e10


def test_search_range_128_vs_oracle(hm):
    """+-128 windows (66049 candidates, BASELINE.json configs[3]'s range): window splitting across shared-memory groups."""
    W, H = 320, 256
    f0, f1 = synth.luma_frame(W, H, 0, seed=13), synth.luma_frame(W, H, 2, seed=13)
    jobs = hm.build_canonical_jobs(W, H, 128, 123456, pred=(6, -3), ctu_first=6, ctu_count=1)[::9]
    jobs = np.concatenate([jobs, hm.build_canonical_jobs(W, H, 128, 123456, ctu_first=0, ctu_count=1)[:13]])
    cur, o0, stride = padded(f1, 144)
    ref, _, _ = padded(f0, 144)
    idc = hm.register_plane_u8(f1, 144, 144, kind=0)
    idr = hm.register_plane_u8(f0, 144, 144, kind=1)
    try:
        got = hm.me_jobs(idc, idr, jobs, flags_of(1, 1))
    finally:
        hm.release_plane(idc)
        hm.release_plane(idr)
    exp, _ = Oracle(fen=1, hadme=1).run_jobs((cur, o0, stride), (ref, o0, stride), jobs, 8, True)
    assert results_equal(got, exp) == []


@pytest.mark.parametrize("lam,pred", [(0, (0, 0)), (65536 * 4, (9, -6)), (1, (-30, 21))])
def test_fused_tie_break_flat_content(hm, lam, pred):
    """Flat content through the CU-fused kernels (every PU of two CTUs): all SADs tie, so the MV cost and then the raster
    order decide for each of the 13 partitions of every CU; a lane sees several tiles of one candidate-row group."""
    W, H = 128, 128
    flat = np.full((H, W), 131, dtype=np.uint8)
    cur, o0, stride = padded(flat)
    jobs = np.concatenate([hm.build_canonical_jobs(W, H, 64, lam, pred=pred, ctu_first=c, ctu_count=1) for c in (0, 3)])
    idc = hm.register_plane_u8(flat, MARGIN, MARGIN)
    try:
        got = hm.me_jobs(idc, idc, jobs, flags_of(1, 1, frac=False))
    finally:
        hm.release_plane(idc)
    exp, _ = Oracle(fen=1, hadme=1).run_jobs((cur, o0, stride), (cur, o0, stride), jobs, 8, False)
    assert results_equal(got, exp, ("mv_x", "mv_y", "sad")) == []
    if lam == 0:
        assert np.array_equal(got["mv_x"], jobs["lt_x"]) and np.array_equal(got["mv_y"], jobs["lt_y"])


@pytest.mark.parametrize("fen", [1, 0])
def test_fused_windows_at_every_alignment(hm, fen):
    """CU-fused kernels stage 16-column candidate blocks aligned in shared memory and mask the columns of the first and
    last block that lie outside the window: a shared predictor per run moves the window's left edge through all 16
    residues (and narrow windows of 1..40 columns through every block count); every PU of a CTU against the oracle."""
    W, H = 192, 128
    f0, f1 = synth.luma_frame(W, H, 0, seed=5), synth.luma_frame(W, H, 1, seed=5)
    cur, o0, stride = padded(f1)
    ref, _, _ = padded(f0)
    idc = hm.register_plane_u8(f1, MARGIN, MARGIN, kind=0)
    idr = hm.register_plane_u8(f0, MARGIN, MARGIN, kind=1)
    rng = np.random.default_rng(11)
    try:
        for r in range(16):
            pred = (4 * r + int(rng.integers(0, 4)), -4 * (r % 5) + int(rng.integers(0, 4)))
            sr = int(rng.integers(1, 21))
            jobs = hm.build_canonical_jobs(W, H, sr, 30000 + 977 * r, pred=pred, ctu_first=1 + (r % 2) * 3, ctu_count=1)
            jobs = jobs[:: 2 if r % 3 else 1]
            got = hm.me_jobs(idc, idr, jobs, flags_of(fen, 1, frac=False))
            exp, _ = Oracle(fen=fen, hadme=1).run_jobs((cur, o0, stride), (ref, o0, stride), jobs, 8, False)
            assert results_equal(got, exp, ("mv_x", "mv_y", "sad")) == [], (r, pred, sr)
    finally:
        hm.release_plane(idc)
        hm.release_plane(idr)


def test_fused_small_cus_wide_windows_index_flush(hm):
    """+-128 windows for every 8x8 and 16x16 CU of a CTU: a warp's run of 8x8-CU tiles covers more candidate rows than the
    local index of the argmin key can number, so the kernel has to flush mid-CU (and a lane sees up to three tiles per
    row group); half of the runs use flat content so that only the raster order decides."""
    W, H = 320, 256
    lam = 77777
    for flat in (False, True):
        f0 = np.full((H, W), 90, dtype=np.uint8) if flat else synth.luma_frame(W, H, 0, seed=17)
        f1 = f0 if flat else synth.luma_frame(W, H, 2, seed=17)
        jobs = hm.build_canonical_jobs(W, H, 128, 0 if flat else lam, pred=(0, 0) if flat else (-9, 14), ctu_first=6, ctu_count=1)
        jobs = jobs[np.maximum(jobs["w"], jobs["h"]) <= 16]
        jobs = jobs[: 13 * 4 + 5 * 24] if flat else jobs[::2]
        cur, o0, stride = padded(f1, 144)
        ref, _, _ = padded(f0, 144)
        idc = hm.register_plane_u8(f1, 144, 144, kind=0)
        idr = hm.register_plane_u8(f0, 144, 144, kind=1)
        try:
            got = hm.me_jobs(idc, idr, jobs, flags_of(1, 1, frac=False))
        finally:
            hm.release_plane(idc)
            hm.release_plane(idr)
        exp, _ = Oracle(fen=1, hadme=1).run_jobs((cur, o0, stride), (ref, o0, stride), jobs, 8, False)
        assert results_equal(got, exp, ("mv_x", "mv_y", "sad")) == [], flat


@pytest.mark.parametrize("fen", [1, 0])
def test_child_fold_and_edge_items_equal_separate_kernels(hm, monkeypatch, fen):
    """A 16x16 CU's pass also yields the PUs of its four 8x8 child CUs (CHILD kernels), and a window whose last column
    starts a 16-column block runs that column as edge items (a lane per candidate row).  Both are scheduling choices: the MV
    field must equal the one of the separate 8x8-CU kernel with masked last blocks (knobs off), for the whole 416x240 list,
    for a thinned list (CUs with missing partitions / missing children) and for flat content (raster order alone decides,
    and the last column precedes the next rows); sampled PUs are checked against the oracle."""
    W, H = 416, 240
    lam = int(np.floor(65536.0 * np.sqrt(0.4624 * 2 ** ((35 - 12) / 3.0))))
    full = hm.build_canonical_jobs(W, H, 64, lam)
    rng = np.random.default_rng(77)
    thin = full[np.sort(rng.choice(len(full), len(full) // 2, replace=False))]
    flat = np.full((H, W), 55, dtype=np.uint8)
    cases = [("full", synth.luma_frame(W, H, 1, seed=9), synth.luma_frame(W, H, 0, seed=9), full),
             ("thin", synth.luma_frame(W, H, 2, seed=9), synth.luma_frame(W, H, 0, seed=9), thin),
             ("flat", flat, flat, hm.build_canonical_jobs(W, H, 64, 0, ctu_first=8, ctu_count=2))]
    # flat content, windows kept but the predictor moved next to the last column: every PU's winner is the edge candidate
    # (64, 0), which ties with (64, 1) - decoded from an edge item's local index
    ew = hm.build_canonical_jobs(W, H, 64, 65536 * 4, ctu_first=8, ctu_count=1).copy()
    ew["pred_x"], ew["pred_y"] = 258, 2
    cases.append(("edgewin", flat, flat, ew))
    for name, f1, f0, jobs in cases:
        fused, work = _run_full(hm, f1, f0, jobs, flags_of(fen, 1, frac=False))
        monkeypatch.setenv("HMB200_NO_CHILD_FOLD", "1")
        monkeypatch.setenv("HMB200_NO_EDGE_ITEMS", "1")
        plain, work2 = _run_full(hm, f1, f0, jobs, flags_of(fen, 1, frac=False))
        monkeypatch.delenv("HMB200_NO_CHILD_FOLD")
        monkeypatch.delenv("HMB200_NO_EDGE_ITEMS")
        assert work["abs_diffs_executed"] < work2["abs_diffs_executed"], name
        assert results_equal(fused, plain, ("mv_x", "mv_y", "sad")) == [], name
        pick = np.sort(rng.choice(len(jobs), 150, replace=False))
        cur, o0, stride = padded(f1)
        ref, _, _ = padded(f0)
        exp, _ = Oracle(fen=fen, hadme=1).run_jobs((cur, o0, stride), (ref, o0, stride), jobs[pick], 8, False)
        assert results_equal(fused[pick], exp, ("mv_x", "mv_y", "sad")) == [], name
        if name == "flat":
            assert np.array_equal(fused["mv_x"], jobs["lt_x"]) and np.array_equal(fused["mv_y"], jobs["lt_y"])
        if name == "edgewin":
            assert np.all(fused["mv_x"] == 64) and np.all(fused["mv_y"] == 0)


def test_tensor_map_window_load_equals_row_copies(hm, monkeypatch):
    """The 8-bit CU kernels stage the most common window geometry of a launch with one tensor-map load (UTMALDG) and every
    other window with one bulk copy per row (knob off: all windows).  Same MV field either way - on a picture whose right /
    bottom CTUs have clipped (smaller) windows, so that both paths run in the same launch - and equal to the oracle."""
    W, H = 416, 240
    lam = int(np.floor(65536.0 * np.sqrt(0.4624 * 2 ** ((35 - 12) / 3.0))))
    jobs = hm.build_canonical_jobs(W, H, 64, lam)
    f1, f0 = synth.luma_frame(W, H, 1, seed=21), synth.luma_frame(W, H, 0, seed=21)
    mapped, _ = _run_full(hm, f1, f0, jobs, flags_of(1, 1, frac=False))
    monkeypatch.setenv("HMB200_NO_TENSOR_MAP", "1")
    rows, _ = _run_full(hm, f1, f0, jobs, flags_of(1, 1, frac=False))
    monkeypatch.delenv("HMB200_NO_TENSOR_MAP")
    assert results_equal(mapped, rows, ("mv_x", "mv_y", "sad")) == []
    pick = np.sort(np.random.default_rng(3).choice(len(jobs), 150, replace=False))
    cur, o0, stride = padded(f1)
    ref, _, _ = padded(f0)
    exp, _ = Oracle(fen=1, hadme=1).run_jobs((cur, o0, stride), (ref, o0, stride), jobs[pick], 8, False)
    assert results_equal(mapped[pick], exp, ("mv_x", "mv_y", "sad")) == []


def test_refinement_on_unique_tiles_equals_per_instance_refinement(hm, monkeypatch):
    """The quarter-pel refinement computes each (original tile, reference tile at the integer MV, half offset) once for all
    PUs that share it (hash pass + gather).  Same MV field, half / quarter vectors and costs as one SATD per tile instance
    (knob off), for coherent motion (most tiles shared), for content whose MVs differ from PU to PU, and with SAD instead of
    the Hadamard distortion; sampled PUs against the oracle."""
    W, H = 416, 240
    lam = int(np.floor(65536.0 * np.sqrt(0.4624 * 2 ** ((35 - 12) / 3.0))))
    jobs = hm.build_canonical_jobs(W, H, 64, lam)
    rng = np.random.default_rng(5)
    noise0, noise1 = rng.integers(0, 256, size=(H, W)).astype(np.uint8), rng.integers(0, 256, size=(H, W)).astype(np.uint8)
    cases = [("coherent", synth.luma_frame(W, H, 1, seed=3), synth.luma_frame(W, H, 0, seed=3), 1),
             ("noise", noise1, noise0, 1), ("sad", synth.luma_frame(W, H, 2, seed=3), synth.luma_frame(W, H, 0, seed=3), 0)]
    for name, f1, f0, had in cases:
        got, _ = _run_full(hm, f1, f0, jobs, flags_of(1, had))
        monkeypatch.setenv("HMB200_NO_FRAC_DEDUPE", "1")
        plain, _ = _run_full(hm, f1, f0, jobs, flags_of(1, had))
        monkeypatch.delenv("HMB200_NO_FRAC_DEDUPE")
        assert results_equal(got, plain) == [], name
        pick = np.sort(rng.choice(len(jobs), 120, replace=False))
        cur, o0, stride = padded(f1)
        ref, _, _ = padded(f0)
        exp, _ = Oracle(fen=1, hadme=had).run_jobs((cur, o0, stride), (ref, o0, stride), jobs[pick], 8, True)
        assert results_equal(got[pick], exp) == [], name


# ---------------------------------------------------------------------------------------------------------------------
# 10-bit content: CU-fused 16-bit kernels (packed 16x2 arithmetic, distortion precision shift)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fen,sr", [(1, 64), (0, 64), (1, 128)])
def test_me_canonical_ctus_10bit_vs_oracle(hm, fen, sr):
    """Whole CTUs of a 10-bit pair (interior, corner, last partial row); +-128 exercises row-split windows."""
    W, H = (416, 240) if sr == 64 else (320, 256)
    f0 = synth.luma_frame(W, H, 0, seed=41, bit_depth=10)
    f1 = synth.luma_frame(W, H, 1, seed=41, bit_depth=10)
    margin = 80 if sr == 64 else 144
    cur, o0, stride = padded(f1, margin)
    ref, _, _ = padded(f0, margin)
    lam = int(np.floor(65536.0 * np.sqrt(0.4624 * 2 ** ((35 - 12) / 3.0))))
    ctus = (0, 9, 27) if sr == 64 else (6,)
    jobs = np.concatenate([hm.build_canonical_jobs(W, H, sr, lam, pred=(3, -2), ctu_first=c, ctu_count=1) for c in ctus])
    jobs = jobs[::4] if sr == 64 else jobs[::2]
    idc = hm.register_plane(cur, W, H, margin, margin, 10, kind=0)
    idr = hm.register_plane(ref, W, H, margin, margin, 10, kind=1)
    try:
        prep = hm.prepare_jobs(jobs, flags_of(fen, 1), 10)
        prep.run(idc, idr)
        got = prep.fetch()
        work = prep.work()
        prep.free()
    finally:
        hm.release_plane(idc)
        hm.release_plane(idr)
    assert work["pus_fused"] > 0.3 * len(jobs)          # the fused 16-bit kernels did run
    exp, _ = Oracle(fen=fen, hadme=1).run_jobs((cur, o0, stride), (ref, o0, stride), jobs, 10, True)
    assert results_equal(got, exp) == []


@pytest.mark.parametrize("bd,fen", [(12, 1), (14, 1), (14, 0), (9, 1)])
def test_me_deep_bit_depths_extreme_samples_vs_oracle(hm, bd, fen):
    """The 16-bit kernels compute SAD = sum(org) + sum(ref) - 2 sum(min(org, ref)) with packed 16x2 minima and signed
    16-bit dot products: exercise every bit depth class with samples at both ends of the range (0 and 2^bd - 1 in
    large flat patches next to noise), where a sign or overflow slip in that form would show."""
    W, H = 192, 128
    rng = np.random.default_rng(1000 + bd)
    top = (1 << bd) - 1
    def frame():
        f = rng.integers(0, top + 1, size=(H, W)).astype(np.uint16)
        f[16:48, 20:90] = top
        f[60:100, 100:170] = 0
        f[70:90, 10:60] = top
        return f
    f0 = frame()
    f1 = np.roll(f0, (3, -5), axis=(0, 1))
    f1[::7, ::5] = rng.integers(0, top + 1, size=f1[::7, ::5].shape).astype(np.uint16)
    cur, o0, stride = padded(f1, MARGIN)
    ref, _, _ = padded(f0, MARGIN)
    lam = int(np.floor(65536.0 * np.sqrt(0.4624 * 2 ** ((35 - 12) / 3.0))))
    jobs = hm.build_canonical_jobs(W, H, 64, lam, pred=(-6, 9))
    jobs = jobs[::3]
    idc = hm.register_plane(cur, W, H, MARGIN, MARGIN, bd, kind=0)
    idr = hm.register_plane(ref, W, H, MARGIN, MARGIN, bd, kind=1)
    try:
        prep = hm.prepare_jobs(jobs, flags_of(fen, 1, frac=False), bd)
        prep.run(idc, idr)
        got = prep.fetch()
        work = prep.work()
        prep.free()
    finally:
        hm.release_plane(idc)
        hm.release_plane(idr)
    assert work["pus_fused"] > 0.3 * len(jobs)          # the fused 16-bit kernels did run
    exp, _ = Oracle(fen=fen, hadme=1).run_jobs((cur, o0, stride), (ref, o0, stride), jobs, bd, False)
    assert results_equal(got, exp, ("mv_x", "mv_y", "sad")) == []


def test_me_ctu_row_entry(hm):
    """hmb200_me_ctu_row: the per-CTU-row launch granularity of the in-encoder frontend; rows must agree with one
    whole-picture call, and a job outside the row is rejected."""
    from video_codecs_b200 import HMB200Error
    W, H = 256, 192
    f0, f1 = synth.luma_frame(W, H, 0, seed=61), synth.luma_frame(W, H, 1, seed=61)
    jobs = hm.build_canonical_jobs(W, H, 64, 300000)
    idc = hm.register_plane_u8(f1, MARGIN, MARGIN, kind=0)
    idr = hm.register_plane_u8(f0, MARGIN, MARGIN, kind=1)
    try:
        whole = hm.me_jobs(idc, idr, jobs, flags_of(1, 1))
        for row in range(3):
            sel = jobs["pu_y"] // 64 == row
            got = hm.me_ctu_row(idc, idr, row, jobs[sel], flags_of(1, 1))
            assert np.array_equal(got, whole[sel])
        with pytest.raises(HMB200Error):
            hm.me_ctu_row(idc, idr, 0, jobs[jobs["pu_y"] // 64 == 1][:4], flags_of(1, 1))
    finally:
        hm.release_plane(idc)
        hm.release_plane(idr)


@pytest.mark.parametrize("bd,sr,n_cols", [(8, 64, 3), (10, 128, 2)])
def test_tile_column_shards_with_halo_crops_equal_unsharded(hm, bd, sr, n_cols):
    """BASELINE.json configs[3]: every tile-column shard uploads only its crop of both planes (column +- search range +
    interpolation reach) and searches its own jobs in crop coordinates; the shards merged in job order must equal the MV
    field of the whole picture searched at once (integer MV, SAD, half / quarter MV and cost)."""
    from video_codecs_b200 import shard
    W, H = 704, 192
    margin = 80 if sr <= 64 else 144
    f0, f1 = synth.luma_frame(W, H, 0, seed=41, bit_depth=bd), synth.luma_frame(W, H, 1, seed=41, bit_depth=bd)
    lam = 51234

    def reg(f, kind):
        return hm.register_plane_u8(f, margin, margin, kind=kind) if bd == 8 else hm.register_plane_u16(f, bd, margin, margin, kind=kind)

    def run(fc, fr, jobs):
        idc, idr = reg(np.ascontiguousarray(fc), 0), reg(np.ascontiguousarray(fr), 1)
        try:
            prep = hm.prepare_jobs(jobs, flags_of(1, 1), bd)
            prep.run(idc, idr)
            out, work = prep.fetch(), prep.work()
            prep.free()
            return out, work
        finally:
            hm.release_plane(idc)
            hm.release_plane(idr)

    full = hm.build_canonical_jobs(W, H, sr, lam)
    whole, work_whole = run(f1, f0, full)
    parts, results, uploaded, fused = [], [], 0, 0
    for c in range(n_cols):
        x0, x1 = hm.tile_column_range(W, n_cols, c)
        jobs = shard.tile_column_jobs(hm, W, H, n_cols, c, sr, lam)
        c0, c1 = shard.tile_column_crop(W, x0, x1, sr)
        assert c0 % 64 == 0 and 0 <= c0 < c1 <= W
        uploaded += c1 - c0
        res, work = run(f1[:, c0:c1], f0[:, c0:c1], shard.shift_jobs(jobs, c0))
        fused += work["pus_fused"]
        results.append(res)
        parts.append(jobs)
    merged = shard.merge_shards(parts, results, full)
    assert results_equal(merged, whole) == []
    assert uploaded < n_cols * W                       # the crops are smaller than whole planes per rank
    assert fused == work_whole["pus_fused"] > 0.9 * len(full)      # a crop keeps every CU on the CU-fused kernels
    assert len(set(zip(whole["mv_x"].tolist(), whole["mv_y"].tolist()))) >= 2


@pytest.mark.parametrize("bd,fen,sr", [(10, 1, 64), (10, 0, 64), (12, 1, 128)])
def test_child_fold_16bit_equals_separate_kernels(hm, monkeypatch, bd, fen, sr):
    """16-bit planes: a 16x16 CU's pass also yields its four 8x8 children (child minima kept in shared memory).  Same MV field
    as with the separate 8x8-CU kernel (knob HMB200_NO_CHILD_FOLD16), whole canonical list and a thinned one (missing partitions /
    children), flat content for the raster-order tie-break; sampled PUs against the oracle."""
    W, H = 320, 192
    margin = 80 if sr <= 64 else 144
    lam = 51000
    full = hm.build_canonical_jobs(W, H, sr, lam)
    rng = np.random.default_rng(7 + bd + fen)
    thin = full[np.sort(rng.choice(len(full), len(full) // 2, replace=False))]
    flat = np.full((H, W), 300, dtype=np.uint16)
    cases = [("full", synth.luma_frame(W, H, 1, seed=19, bit_depth=bd), synth.luma_frame(W, H, 0, seed=19, bit_depth=bd), full),
             ("thin", synth.luma_frame(W, H, 2, seed=19, bit_depth=bd), synth.luma_frame(W, H, 0, seed=19, bit_depth=bd), thin),
             ("flat", flat, flat, hm.build_canonical_jobs(W, H, sr, 0, ctu_first=6, ctu_count=1))]

    def run(f1, f0, jobs):
        idc = hm.register_plane_u16(np.ascontiguousarray(f1), bd, margin, margin, kind=0)
        idr = hm.register_plane_u16(np.ascontiguousarray(f0), bd, margin, margin, kind=1)
        try:
            prep = hm.prepare_jobs(jobs, flags_of(fen, 1, frac=False), bd)
            prep.run(idc, idr)
            out, work = prep.fetch(), prep.work()
            prep.free()
            return out, work
        finally:
            hm.release_plane(idc)
            hm.release_plane(idr)

    for name, f1, f0, jobs in cases:
        fused, work = run(f1, f0, jobs)
        monkeypatch.setenv("HMB200_NO_CHILD_FOLD16", "1")
        plain, work2 = run(f1, f0, jobs)
        monkeypatch.delenv("HMB200_NO_CHILD_FOLD16")
        assert work["abs_diffs_executed"] < work2["abs_diffs_executed"], name
        assert results_equal(fused, plain, ("mv_x", "mv_y", "sad")) == [], name
        pick = np.sort(rng.choice(len(jobs), 100, replace=False))
        cur, o0, stride = padded(f1, margin)
        ref, _, _ = padded(f0, margin)
        exp, _ = Oracle(fen=fen, hadme=1).run_jobs((cur, o0, stride), (ref, o0, stride), jobs[pick], bd, False)
        assert results_equal(fused[pick], exp, ("mv_x", "mv_y", "sad")) == [], name
        if name == "flat":
            assert np.array_equal(fused["mv_x"], jobs["lt_x"]) and np.array_equal(fused["mv_y"], jobs["lt_y"])
