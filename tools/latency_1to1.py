"""Round-trip latency of the 1:1 entries (the calls the in-encoder forwarders make), per PU size.

    python tools/latency_1to1.py [--calls 1500]

Prints microseconds per call of hmb200_pattern_search / hmb200_pattern_search_frac / hmb200_pattern_search_and_refine for one
PU at the centre of a 1080p plane, +-64 window, FEN + Hadamard, and of the cheapest possible round trip (a 4x4 SAD through
hmb200_dist).  ctypes adds ~4 us per call.  A search_and_refine call for a square PU searches the whole CU;
HMB200_NO_SPECULATION=1 times the per-PU path for those sizes.  Needs a GPU."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_codecs_b200 import api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--calls", type=int, default=1500)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    hm = api.HMB200()
    hm.init(0)
    W, H, M = 1920, 1080, 80
    f0, f1 = synth.luma_frame(W, H, 0), synth.luma_frame(W, H, 1)
    pad = lambda f: np.ascontiguousarray(np.pad(f.astype(np.int16), M, mode="edge"))
    ref, cur = pad(f0), pad(f1)
    stride = W + 2 * M
    idr = hm.register_plane(ref, W, H, M, M, 8, kind=1)
    lam = int(np.floor(65536.0 * np.sqrt(0.4624 * 2 ** ((35 - 12) / 3.0))))
    out = {}

    import itertools
    tick = itertools.count(1)

    def timed(fn):
        for _ in range(50):
            fn()
        t0 = time.perf_counter()
        for _ in range(args.calls):
            fn()
        return 1e6 * (time.perf_counter() - t0) / args.calls

    x0, y0 = 960, 512
    off = (y0 + M) * stride + x0 + M
    out["dist_4x4_sad"] = timed(lambda: hm.dist(0, (cur, off, stride), (ref, off, stride), 4, 4))
    for w, h in ((8, 8), (16, 16), (32, 32), (64, 64), (64, 32), (16, 4)):
        org, rf = (cur, off, stride), (ref, off, stride)
        mv = hm.pattern_search(org, w, h, rf, (-64, -64), (64, 64), lam, (0, 0))[0]
        out[f"{w}x{h}"] = {
            "search": timed(lambda: hm.pattern_search(org, w, h, rf, (-64, -64), (64, 64), lam, (0, 0))),
            "frac": timed(lambda: hm.pattern_search_frac(org, w, h, rf, mv, lam, (0, 0))),
            # a new lambda per call: an identical repeat would be answered from the previous call's whole-CU launch (square PUs)
            "search_and_refine": timed(lambda: hm.pattern_search_and_refine(org, w, h, rf, (-64, -64), (64, 64), lam + next(tick), (0, 0))),
        }
    hm.release_plane(idr)
    print(json.dumps(out, indent=1))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
