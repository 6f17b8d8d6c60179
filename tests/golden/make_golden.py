"""Generates tests/golden/hm_golden.npz from the UNMODIFIED reference (oracle/_ref/libhmref.so, built from
/root/reference/hm-16.5rc1 by oracle/Makefile.ref).  Run in the build container only:

    python tests/golden/make_golden.py

The .npz holds the inputs (small planes, job lists) and the reference's outputs, so that the oracle and the CUDA
path can be checked on machines where /root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.pyoracle import Reference, JOB_DTYPE  # noqa: E402
from video_codecs_b200 import synth  # noqa: E402

PU_SIZES = [(64, 64), (64, 32), (32, 64), (64, 16), (64, 48), (16, 64), (48, 64), (32, 32), (32, 16), (16, 32), (32, 8),
            (32, 24), (8, 32), (24, 32), (16, 16), (16, 8), (8, 16), (16, 4), (16, 12), (4, 16), (12, 16), (8, 8), (8, 4),
            (4, 8)]
MARGIN = 80
W, H = 256, 128


def make_dist_cases(rng):
    """(bit_depth, kind, sub_shift, w, h, org_off, cur_off) over two random planes per bit depth."""
    cases = []
    for bd in (8, 10):
        for (w, h) in PU_SIZES:
            for kind, ss in ((0, 0), (0, 1), (1, 0), (2, 0), (3, 0)):
                for _ in range(2):
                    oy, ox = rng.integers(0, 192 - h), rng.integers(0, 192 - w)
                    cy, cx = rng.integers(0, 192 - h), rng.integers(0, 192 - w)
                    cases.append((bd, kind, ss, w, h, oy * 192 + ox, cy * 192 + cx))
    return np.array(cases, dtype=np.int32)


def make_jobs(rng, n_per_size, search_range, lam_choices):
    jobs = []
    for (w, h) in PU_SIZES:
        for k in range(n_per_size):
            if k == 0:   # picture corner: exercises window clipping against the margins
                px, py = (0, 0) if (w + h) % 16 == 0 else (W - w, H - h)
            else:
                px, py = int(rng.integers(0, (W - w) // 4 + 1)) * 4, int(rng.integers(0, (H - h) // 4 + 1)) * 4
            pred = (int(rng.integers(-40, 41)), int(rng.integers(-40, 41))) if k % 2 else (0, 0)
            jobs.append((px, py, w, h, pred, int(lam_choices[rng.integers(0, len(lam_choices))])))
    return jobs


def window(pred, rng_, cu_x, cu_y):
    """xSetSearchRange restated for the generator (TEncSearch.cpp:3765-3781, TComDataCU.cpp:2788-2801); the CU is
    taken to be the 64-aligned block containing the PU."""
    def clip(x, y):
        hmax, hmin = (W + 8 - cu_x - 1) * 4, (-64 - 8 - cu_x + 1) * 4
        vmax, vmin = (H + 8 - cu_y - 1) * 4, (-64 - 8 - cu_y + 1) * 4
        return min(hmax, max(hmin, x)), min(vmax, max(vmin, y))
    px, py = clip(*pred)
    lx, ly = clip(px - rng_ * 4, py - rng_ * 4)
    rx, ry = clip(px + rng_ * 4, py + rng_ * 4)
    return lx >> 2, ly >> 2, rx >> 2, ry >> 2


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    # ---- distortion table ---------------------------------------------------------------------------------------
    ref_fen = Reference(fen=1, hadme=1)
    for bd in (8, 10):
        hi = 1 << bd
        a = rng.integers(0, hi, size=(192, 192)).astype(np.int16)
        b = np.clip(a + rng.integers(-24, 25, size=a.shape), 0, hi - 1).astype(np.int16)
        b[:96] = rng.integers(0, hi, size=(96, 192)).astype(np.int16)          # half correlated, half not
        out[f"dist_org_{bd}"], out[f"dist_cur_{bd}"] = a, b
    cases = make_dist_cases(rng)
    vals = []
    for (bd, kind, ss, w, h, oo, co) in cases:
        vals.append(ref_fen.dist(int(kind), (out[f"dist_org_{bd}"], int(oo), 192), (out[f"dist_cur_{bd}"], int(co), 192),
                                 int(w), int(h), int(bd), int(ss)))
    out["dist_cases"], out["dist_expected"] = cases, np.array(vals, dtype=np.uint32)
    # ---- MV rate ---------------------------------------------------------------------------------------------------
    bits_cases = []
    for _ in range(400):
        bits_cases.append((int(rng.integers(-300, 301)), int(rng.integers(-300, 301)), int(rng.integers(-200, 201)),
                           int(rng.integers(-200, 201)), int(rng.integers(0, 3)), int(rng.integers(0, 1 << 23))))
    bits_cases.append((0, 0, 0, 0, 2, 4037017))
    bits_cases.append((64, -64, -3, 5, 2, 0xFFFFFFFF))      # 32-bit wrap of m_uiCost * bits
    bc = np.array(bits_cases, dtype=np.int64)
    out["rate_cases"] = bc
    out["rate_bits"] = np.array([ref_fen.mv_bits(int(x), int(y), (int(px), int(py)), int(s)) for (x, y, px, py, s, lam) in bc], dtype=np.uint32)
    out["rate_cost"] = np.array([ref_fen.mv_cost(int(lam), int(x), int(y), (int(px), int(py)), int(s)) for (x, y, px, py, s, lam) in bc], dtype=np.uint32)
    # ---- searches ------------------------------------------------------------------------------------------------
    lam = [int(np.floor(65536.0 * np.sqrt(0.4624 * 2 ** ((qp - 12) / 3.0)))) for qp in (22, 32, 35, 40)]
    for bd in (8, 10):
        f0 = synth.luma_frame(W, H, 0, seed=77, bit_depth=bd)
        f1 = synth.luma_frame(W, H, 1, seed=77, bit_depth=bd, vx=2.75, vy=-1.25)
        out[f"frame0_{bd}"], out[f"frame1_{bd}"] = f0, f1
        cur = synth.pad_plane(f1, MARGIN, MARGIN)
        refp = synth.pad_plane(f0, MARGIN, MARGIN)
        stride = cur.shape[1]
        o0 = MARGIN * stride + MARGIN
        for fen, hadme, sr, nps in ((1, 1, 64, 3), (0, 0, 12, 3), (1, 0, 20, 1), (0, 1, 9, 1)):
            spec = make_jobs(rng, nps, sr, lam)
            jobs = np.zeros(len(spec), dtype=JOB_DTYPE)
            for i, (px, py, w, h, pred, lc) in enumerate(spec):
                lt_x, lt_y, rb_x, rb_y = window(pred, sr, (px // 64) * 64, (py // 64) * 64)
                jobs[i] = (px, py, w, h, lt_x, lt_y, rb_x, rb_y, pred[0], pred[1], lc, 0)
            R = Reference(fen=fen, hadme=hadme)
            res, _ = R.run_jobs((cur, o0, stride), (refp, o0, stride), jobs, bit_depth=bd, do_frac=True)
            tag = f"search_bd{bd}_fen{fen}_had{hadme}"
            out[tag + "_jobs"], out[tag + "_results"] = jobs, res
            print(tag, len(jobs), "half!=0:", int(np.count_nonzero((res["half_x"] != 0) | (res["half_y"] != 0))),
                  "qter!=0:", int(np.count_nonzero((res["qter_x"] != 0) | (res["qter_y"] != 0))))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hm_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def make_ingest_golden():
    """tests/golden/yuv_ingest_golden.npz: file images and the luma planes the reference's TVideoIOYuv reads from them."""
    import tempfile
    ref = Reference(fen=1, hadme=1)
    cases = [(64, 40, 0, 0, 8, 8), (52, 36, 4, 4, 8, 8), (64, 24, 0, 8, 8, 10), (40, 24, 8, 0, 10, 10), (48, 32, 0, 0, 10, 8),
             (32, 16, 8, 8, 12, 10), (24, 16, 0, 0, 16, 12)]
    rng = np.random.default_rng(2024)
    out = {"cases": np.array(cases, dtype=np.int32)}
    with tempfile.TemporaryDirectory() as d:
        for i, (w, h, px, py, fbd, ibd) in enumerate(cases):
            data = (rng.integers(0, 256, size=w * h, dtype=np.uint8).tobytes() if fbd == 8
                    else rng.integers(0, 1 << fbd, size=w * h).astype("<u2").tobytes())
            path = os.path.join(d, f"c{i}.yuv")
            open(path, "wb").write(data)
            out[f"file_{i}"] = np.frombuffer(data, dtype=np.uint8)
            out[f"plane_{i}"] = ref.read_luma(path, w, h, px, py, fbd, ibd)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "yuv_ingest_golden.npz"), **out)
    print("wrote yuv_ingest_golden.npz")


def make_tz_golden():
    """tests/golden/tz_golden.npz: xTZSearch results of the unmodified reference (FEN = 1) for the cases of
    tests/test_tz_search.py::make_cases, with the windows of xSetSearchRange."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_tz_search as T
    from common import padded
    from oracle.pyoracle import Oracle
    ref_impl, O = Reference(fen=1, hadme=1), Oracle(fen=1, hadme=1)
    f0, f1 = synth.luma_frame(T.W, T.H, 0, seed=17), synth.luma_frame(T.W, T.H, 3, seed=17)
    cur, o0, stride = padded(f1)
    ref, _, _ = padded(f0)
    cases, exp, exp0 = [], [], []
    for (px, py, w, h, cu_x, cu_y, pred, lam, sr, imv) in T.make_cases(np.random.default_rng(7), 300):
        lt_rb = O.search_range(pred, sr, (cu_x, cu_y), (T.W, T.H))
        off = o0 + py * stride + px
        mv, sad = ref_impl.tz_search((cur, off, stride), w, h, (ref, off, stride), lt_rb[:2], lt_rb[2:], lam, pred, (cu_x, cu_y), (T.W, T.H), sr, imv,
                                     first_search_stop=1)
        mv0, sad0 = ref_impl.tz_search((cur, off, stride), w, h, (ref, off, stride), lt_rb[:2], lt_rb[2:], lam, pred, (cu_x, cu_y), (T.W, T.H), sr,
                                       imv, first_search_stop=0)
        exp0.append((mv0[0], mv0[1], sad0))
        cases.append((px, py, w, h, cu_x, cu_y, pred[0], pred[1], lam, sr, 0 if imv is None else 1, 0 if imv is None else imv[0],
                      0 if imv is None else imv[1]) + tuple(lt_rb))
        exp.append((mv[0], mv[1], sad))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tz_golden.npz"), frame0=f0, frame1=f1,
                        cases=np.array(cases, dtype=np.int64), expected_stop1=np.array(exp, dtype=np.int64),
                        expected_stop0=np.array(exp0, dtype=np.int64))
    print("wrote tz_golden.npz", len(cases))


def make_mc_golden():
    """tests/golden/mc_golden.npz: xPredInterBlk + SAD / HADs of the unmodified reference for tests/test_mc_dist.py::cases."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_mc_dist as T
    ref_impl = Reference(fen=1, hadme=1)
    out = {}
    for bd in (8, 10):
        f0, f1, cur, ref, o0, stride = T.frames(bd)
        cs = T.cases(np.random.default_rng(11 + bd), 240)
        out[f"frame0_{bd}"] = f0
        out[f"cases_{bd}"] = np.array([(px, py, w, h, mv[0], mv[1], kind) for (px, py, w, h, mv, kind) in cs], dtype=np.int64)
        out[f"expected_{bd}"] = np.array([ref_impl.mc_dist(kind, (cur, o0 + py * stride + px, stride), w, h, (ref, o0, stride), (T.W, T.H),
                                                           T.MARGIN, (px, py), mv, bd) for (px, py, w, h, mv, kind) in cs], dtype=np.int64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mc_golden.npz"), **out)
    print("wrote mc_golden.npz")


def make_mc_cand_golden():
    """tests/golden/mc_cand_golden.npz: bi-prediction (xPredInterBlk bi = true x 2 + xWeightedAverage) + SAD / HADs of the unmodified
    reference for tests/test_mc_cand.py::bi_cases."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_mc_cand as T
    ref_impl = Reference(fen=1, hadme=1)
    out = {}
    for bd in (8, 10):
        f, pads, o0, stride = T.frames(bd)
        cur, ra, rb = pads[1], pads[0], pads[2]
        cs = T.bi_cases(np.random.default_rng(17 + bd), 200)
        out[f"frame2_{bd}"] = f[2]
        out[f"cases_{bd}"] = np.array(cs, dtype=np.int64)
        out[f"expected_{bd}"] = np.array([ref_impl.mc_bi_dist(kind, (cur, o0 + py * stride + px, stride), w, h, (ra, o0, stride), (rb, o0, stride),
                                                              (T.W, T.H), T.MARGIN, (px, py), (ax, ay), (bx, by), bd)
                                          for (px, py, w, h, ax, ay, bx, by, kind) in cs], dtype=np.int64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mc_cand_golden.npz"), **out)
    print("wrote mc_cand_golden.npz")


if __name__ == "__main__":
    if "--mc-cand" in sys.argv:
        make_mc_cand_golden()
        sys.exit(0)
    if "--mc" in sys.argv:
        make_mc_golden()
        sys.exit(0)
    if "--tz" in sys.argv:
        make_tz_golden()
        sys.exit(0)
    if "--ingest" in sys.argv:
        make_ingest_golden()
        sys.exit(0)
    main()
