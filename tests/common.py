"""Shared helpers for the test-suite (inputs only; no arithmetic of the path lives here)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "hm_golden.npz")
MARGIN = 80

PU_SIZES = [(64, 64), (64, 32), (32, 64), (64, 16), (64, 48), (16, 64), (48, 64), (32, 32), (32, 16), (16, 32), (32, 8),
            (32, 24), (8, 32), (24, 32), (16, 16), (16, 8), (8, 16), (16, 4), (16, 12), (4, 16), (12, 16), (8, 8), (8, 4),
            (4, 8)]


def load_golden():
    return np.load(GOLDEN)


def padded(frame, margin=MARGIN):
    from video_codecs_b200 import synth
    p = synth.pad_plane(frame, margin, margin)
    return p, margin * p.shape[1] + margin, p.shape[1]


def results_equal(a, b, fields=None):
    fields = fields or a.dtype.names
    bad = [f for f in fields if not np.array_equal(a[f], b[f])]
    return bad


def first_diff(a, b):
    for i in range(len(a)):
        if a[i] != b[i]:
            return i, a[i], b[i]
    return None
