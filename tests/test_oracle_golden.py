"""CPU suite: the C oracle (oracle/hm_oracle.c) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle on machines without /root/reference."""
import numpy as np
import pytest

from common import load_golden, padded, results_equal
from oracle.pyoracle import Oracle


@pytest.fixture(scope="module")
def gold():
    return load_golden()


def test_rate_bits_and_cost(gold, oracle):
    for (x, y, px, py, s, lam), eb, ec in zip(gold["rate_cases"], gold["rate_bits"], gold["rate_cost"]):
        assert oracle.mv_bits(int(x), int(y), (int(px), int(py)), int(s)) == int(eb)
        assert oracle.mv_cost(int(lam), int(x), int(y), (int(px), int(py)), int(s)) == int(ec)


def test_distortion_table(gold, oracle):
    cases, exp = gold["dist_cases"], gold["dist_expected"]
    assert len(cases) == 24 * 2 * 5 * 2
    for (bd, kind, ss, w, h, oo, co), e in zip(cases, exp):
        got = oracle.dist(int(kind), (gold[f"dist_org_{bd}"], int(oo), 192), (gold[f"dist_cur_{bd}"], int(co), 192),
                          int(w), int(h), int(bd), int(ss))
        assert got == int(e), (bd, kind, ss, w, h)


@pytest.mark.parametrize("bd", [8, 10])
@pytest.mark.parametrize("fen,had", [(1, 1), (0, 0), (1, 0), (0, 1)])
def test_search_jobs(gold, bd, fen, had):
    tag = f"search_bd{bd}_fen{fen}_had{had}"
    jobs, exp = gold[tag + "_jobs"], gold[tag + "_results"]
    cur, o0, stride = padded(gold[f"frame1_{bd}"])
    ref, _, _ = padded(gold[f"frame0_{bd}"])
    got, _ = Oracle(fen=fen, hadme=had).run_jobs((cur, o0, stride), (ref, o0, stride), jobs, bit_depth=bd, do_frac=True)
    assert results_equal(got, exp) == []


def test_golden_covers_all_fractional_outcomes(gold):
    """The fixture must reach every half-pel offset and every quarter-pel offset at least once."""
    halves, qters = set(), set()
    for k in gold.files:
        if k.endswith("_results"):
            r = gold[k]
            halves |= set(zip(r["half_x"].tolist(), r["half_y"].tolist()))
            qters |= set(zip(r["qter_x"].tolist(), r["qter_y"].tolist()))
    assert len(halves) == 9 and len(qters) == 9


@pytest.mark.parametrize("bd", [9, 10, 12, 14])
def test_sad_sum_of_minima_identity_matches_oracle(bd):
    """The arithmetic the 16-bit CUDA search relies on (hmb200_search16_cu.cuh): SAD = sum(org) + sum(ref) - 2 sum(min),
    accumulated from packed half-word pairs by a signed 16-bit dot product, and one shift instead of
    (sum << iSubShift) >> (bitDepth - 8) -- checked here against the oracle's xGetSAD restatement, FEN row sub-sampling
    included, with samples at both ends of the range."""
    from oracle.pyoracle import Oracle
    o = Oracle()
    rng = np.random.default_rng(bd)
    top = (1 << bd) - 1
    for (w, h, ss) in [(8, 8, 0), (16, 16, 1), (16, 4, 0), (32, 24, 1), (64, 64, 1), (64, 16, 1)]:
        org = rng.integers(0, top + 1, size=(h, 80)).astype(np.int16)
        ref = rng.integers(0, top + 1, size=(h, 80)).astype(np.int16)
        org[: h // 2, :5] = top
        ref[: h // 2, :5] = 0
        ref[h // 2:, 3:9] = top
        rows = np.arange(0, h, 1 << ss)
        a = org[rows, :w].astype(np.int64)
        b = ref[rows, :w].astype(np.int64)
        # packed pairs as the kernel sees them: signed 16-bit halves times taps (1, 1)
        halves = np.minimum(a, b).astype(np.uint16).view(np.int16).astype(np.int64)
        m = int(halves.sum())
        total = int(a.sum() + b.sum() - 2 * m)
        shr = bd - 8
        assert shr >= ss
        assert total >> (shr - ss) == (total << ss) >> shr
        assert total >> (shr - ss) == o.sad((org, 0, 80), (ref, 0, 80), w, h, bd, ss)
