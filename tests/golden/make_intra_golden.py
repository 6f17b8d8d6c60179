"""Generates tests/golden/intra_golden.npz from the UNMODIFIED reference (oracle/_ref/libhmref.so): inputs (a small
picture, block list, reference lines) and the reference's 35 first-pass Hadamard distortions per block, plus a few full
predictions.  Run in the build container only:

    python tests/golden/make_intra_golden.py

Reference lines are what a lookahead on original samples would pass: the unfiltered lines are the picture's own
neighbours (border blocks replicate the nearest sample), the smoothed lines are their [1 2 1] / 4 filter with the end
samples kept - plausible inputs only; the functions under test take the lines as given.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.pyoracle import Reference  # noqa: E402
from video_codecs_b200 import synth  # noqa: E402

W, H = 192, 128


def lines_for(plane, x, y, n, rng):
    """(top_unf, left_unf, top_flt, left_flt), corner at index 0, from the picture's own samples (clamped at the borders)."""
    def px(xx, yy):
        return int(plane[min(max(yy, 0), plane.shape[0] - 1), min(max(xx, 0), plane.shape[1] - 1)])
    top = np.array([px(x - 1 + i, y - 1) for i in range(2 * n + 1)], dtype=np.int32)
    left = np.array([px(x - 1, y - 1 + i) for i in range(2 * n + 1)], dtype=np.int32)

    def smooth(a, corner_nb):
        f = a.copy()
        f[0] = (corner_nb + 2 * a[0] + a[1] + 2) >> 2
        f[1:-1] = (a[:-2] + 2 * a[1:-1] + a[2:] + 2) >> 2
        return f
    tf, lf = smooth(top, left[1]), smooth(left, top[1])
    lf[0] = tf[0]
    return [v.astype(np.int16) for v in (top, left, tf, lf)]


def main():
    R = Reference()
    rng = np.random.default_rng(2026)
    out = {}
    for bd in (8, 10):
        f = synth.luma_frame(W, H, 3, seed=9).astype(np.int16)
        if bd == 10:
            f = ((f.astype(np.int32) << 2) | rng.integers(0, 4, f.shape)).astype(np.int16)
        blocks, refs, sads = [], [], []
        off = 0
        for n in (4, 8, 16, 32, 64):
            for _ in range(6 if n < 64 else 3):
                x = int(rng.integers(0, (W - n) // 4 + 1)) * 4
                y = int(rng.integers(0, (H - n) // 4 + 1)) * 4
                ln = lines_for(f, x, y, n, rng)
                blocks.append((x, y, n, off, 3, 0))
                refs.extend(ln)
                off += 4 * (2 * n + 1)
                sads.append(R.intra_modes_had((np.ascontiguousarray(f), y * W + x, W), ln[0], ln[1], ln[2], ln[3], n, bd))
        out[f"plane{bd}"] = f
        out[f"blocks{bd}"] = np.array(blocks, dtype=np.int32)
        out[f"refs{bd}"] = np.concatenate(refs)
        out[f"had{bd}"] = np.array(sads, dtype=np.uint32)
    # a few whole predictions (every mode of one 8x8 and one 32x32 block, 8-bit)
    preds = []
    for n in (8, 32):
        ln = lines_for(out["plane8"], 64, 32, n, rng)
        for mode in range(35):
            f_ = R.intra_use_filtered(mode, n)
            preds.append(R.intra_predict(mode, ln[2] if f_ else ln[0], ln[3] if f_ else ln[1], n, 8).ravel())
        out[f"pred_lines{n}"] = np.concatenate(ln)
    out["preds8"] = np.concatenate(preds[:35])
    out["preds32"] = np.concatenate(preds[35:])
    path = os.path.join(ROOT, "tests", "golden", "intra_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
