"""Times the intra first pass (hmb200_intra_modes_had_batch) on every block of a 1080p picture's CU quadtree
(64x64 .. 8x8 CUs as 2Nx2N, 8x8 CUs also as four 4x4: 341 blocks per CTU, 173 910 blocks), reference lines taken from
the original picture's own neighbours, and the oracle (CPU, one core) on a sample.  Prints one JSON object.

    python tests/measure_intra.py            (needs a B200; the oracle leg is test infrastructure, used as the checker)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tests/ -> repo root
sys.path.insert(0, ROOT)

from video_codecs_b200 import HMB200, INTRA_BLOCK_DTYPE, synth  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

W, H = 1920, 1088


def build(plane):
    P = np.pad(plane.astype(np.int32), ((1, 129), (1, 129)), mode="edge")       # P[y+1, x+1] = plane[y, x], clamped
    blocks, refs, off = [], [], 0
    for cy in range(0, H, 64):
        for cx in range(0, W, 64):
            for n in (64, 32, 16, 8, 4):
                for y in range(cy, cy + 64, n):
                    for x in range(cx, cx + 64, n):
                        top = P[y, x: x + 2 * n + 1]
                        left = P[y: y + 2 * n + 1, x]
                        tf, lf = top.copy(), left.copy()
                        tf[1:-1] = (top[:-2] + 2 * top[1:-1] + top[2:] + 2) >> 2
                        lf[1:-1] = (left[:-2] + 2 * left[1:-1] + left[2:] + 2) >> 2
                        tf[0] = lf[0] = (left[1] + 2 * top[0] + top[1] + 2) >> 2
                        blocks.append((x, y, n, off, 3, 0))
                        refs.append(np.concatenate([top, left, tf, lf]).astype(np.int16))
                        off += 4 * (2 * n + 1)
    rec = np.array(blocks, dtype=np.int32).view(INTRA_BLOCK_DTYPE).ravel()
    return rec, np.concatenate(refs)


def main():
    f = synth.luma_frame(W, H, 2, seed=4)
    blocks, refs = build(f)
    hm = HMB200(); hm.init(0)
    pid = hm.register_plane_u8(f, 80, 80, kind=0)
    got = hm.intra_modes_had_batch(pid, blocks, refs)
    l0 = hm.launch_count()
    t0 = time.perf_counter()
    K = 5
    for _ in range(K):
        got = hm.intra_modes_had_batch(pid, blocks, refs)
    ms = 1e3 * (time.perf_counter() - t0) / K
    launches = (hm.launch_count() - l0) // K
    O = Oracle()
    plane16 = np.ascontiguousarray(f.astype(np.int16))
    pick = np.arange(0, len(blocks), 97)
    t0 = time.perf_counter()
    for i in pick:
        b = blocks[i]; n = int(b["n"]); L = 2 * n + 1; o = int(b["ref_off"])
        exp = O.intra_modes_had((plane16, int(b["y"]) * W + int(b["x"]), W), refs[o:o + L], refs[o + L:o + 2 * L], refs[o + 2 * L:o + 3 * L],
                                refs[o + 3 * L:o + 4 * L], n, 8)
        assert np.array_equal(exp, got[i]), i
    cpu_s = time.perf_counter() - t0
    px = float(np.sum(blocks["n"].astype(np.int64) ** 2))
    cpu_px = float(np.sum(blocks["n"][pick].astype(np.int64) ** 2))
    print(json.dumps({"what": "intra first pass, 35 modes x Hadamard, 1080p CU quadtree (5 depths incl. 4x4)", "blocks": int(len(blocks)),
                      "block_pixels": px, "ms_per_picture_e2e": ms, "gpu_launches_per_picture": int(launches),
                      "mode_pixels_per_s": 35 * px / (ms / 1e3), "h2d_bytes": int(refs.nbytes + blocks.nbytes), "d2h_bytes": int(got.nbytes),
                      "checked_blocks_bit_exact": int(len(pick)),
                      "cpu_oracle_one_core": {"blocks": int(len(pick)), "seconds": cpu_s, "mode_pixels_per_s": 35 * cpu_px / cpu_s}}))
    hm.release_plane(pid); hm.shutdown()


if __name__ == "__main__":
    main()
