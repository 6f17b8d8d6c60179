"""TEST INFRASTRUCTURE — ctypes front-ends for the two checkers, with one API:

  Oracle()     -> oracle/_build/libhmoracle.so   (C restatement, oracle/hm_oracle.c)
  Reference()  -> oracle/_ref/libhmref.so        (unmodified HM-16.5 behind oracle/ref_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
All sample buffers are numpy int16 ("Pel", TLibCommon/TypeDef.h:219); jobs/results use JOB_DTYPE/RESULT_DTYPE,
which match hmb200_pu_job / hmb200_pu_result in include/hmb200.h field for field.
"""
import ctypes as C
import os

import numpy as np

from . import build_oracle as _b

JOB_DTYPE = np.dtype([("pu_x", "<i4"), ("pu_y", "<i4"), ("w", "<i4"), ("h", "<i4"),
                      ("lt_x", "<i4"), ("lt_y", "<i4"), ("rb_x", "<i4"), ("rb_y", "<i4"),
                      ("pred_x", "<i4"), ("pred_y", "<i4"), ("lambda_cost", "<u4"), ("reserved", "<i4")])
RESULT_DTYPE = np.dtype([("mv_x", "<i4"), ("mv_y", "<i4"), ("sad", "<u4"),
                         ("half_x", "<i4"), ("half_y", "<i4"), ("qter_x", "<i4"), ("qter_y", "<i4"),
                         ("frac_cost", "<u4")])

_p16 = C.POINTER(C.c_int16)
_pi = C.POINTER(C.c_int)
_pu = C.POINTER(C.c_uint32)


def _ptr(a, off=0):
    """Pointer to element `off` of a C-contiguous int16 array (off may address the interior of a padded plane)."""
    assert a.dtype == np.int16 and a.flags["C_CONTIGUOUS"]
    return C.cast(a.ctypes.data + 2 * int(off), _p16)




def _intra_bind(lib, prefix, with_handle):
    h = [C.c_void_p] if with_handle else []
    f = getattr(lib, prefix + "intra_use_filtered"); f.restype = C.c_int; f.argtypes = [C.c_int, C.c_int]
    f = getattr(lib, prefix + "intra_predict"); f.restype = None
    f.argtypes = h + [C.c_int, _p16, _p16, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p16, C.c_int]
    f = getattr(lib, prefix + "intra_modes_had"); f.restype = None
    f.argtypes = h + [_p16, C.c_int, _p16, _p16, _p16, _p16, C.c_int, C.c_int] + ([] if with_handle else [C.c_int, C.c_int]) + [_pu]


class _IntraMixin:
    """35 luma intra predictions + Hadamard distortions (TEncSearch::estIntraPredQT first pass).  top / left: int16 lines
    of 2n+1 reference samples with the corner at index 0."""

    def intra_use_filtered(self, mode, n):
        return int(getattr(self.lib, self._pfx + "intra_use_filtered")(int(mode), int(n)))

    def intra_predict(self, mode, top, left, n, bit_depth=8, above_ok=1, left_ok=1, edge_filters=1):
        top = np.ascontiguousarray(top, dtype=np.int16); left = np.ascontiguousarray(left, dtype=np.int16)
        out = np.zeros((n, n), dtype=np.int16)
        args = [int(mode), _ptr(top), _ptr(left), int(n), int(bit_depth), int(above_ok), int(left_ok), int(edge_filters), _ptr(out), int(n)]
        getattr(self.lib, self._pfx + "intra_predict")(*(self._h() + args))
        return out

    def intra_modes_had(self, org, top_unf, left_unf, top_flt, left_flt, n, bit_depth=8):
        (oa, oo, os_) = org
        arrs = [np.ascontiguousarray(a, dtype=np.int16) for a in (top_unf, left_unf, top_flt, left_flt)]
        out = np.zeros(35, dtype=np.uint32)
        args = [_ptr(oa, oo), int(os_)] + [_ptr(a) for a in arrs] + [int(n), int(bit_depth)] + ([] if self._h() else [1, 1]) + \
               [out.ctypes.data_as(_pu)]
        getattr(self.lib, self._pfx + "intra_modes_had")(*(self._h() + args))
        return out


class _Base(_IntraMixin):
    """Shared Python surface.  `org`/`cur`/`ref` arguments are (array, element offset, stride) triples."""

    def dist(self, kind, org, cur, w, h, bit_depth=8, sub_shift=0):
        raise NotImplementedError

    def sad(self, org, cur, w, h, bit_depth=8, sub_shift=0):
        return self.dist(0, org, cur, w, h, bit_depth, sub_shift)

    def sse(self, org, cur, w, h, bit_depth=8):
        return self.dist(1, org, cur, w, h, bit_depth, 0)

    def had(self, org, cur, w, h, bit_depth=8):
        return self.dist(2, org, cur, w, h, bit_depth, 0)


class Oracle(_Base):
    def __init__(self, fen=1, hadme=1):
        self.fen, self.hadme = int(fen), int(hadme)
        L = self.lib = C.CDLL(_b.build_oracle())
        self._pfx = "hmo_"; self._h = lambda: []
        _intra_bind(L, "hmo_", False)
        L.hmo_eg_bits.restype = C.c_uint32
        L.hmo_eg_bits.argtypes = [C.c_int32]
        L.hmo_mv_bits.restype = C.c_uint32
        L.hmo_mv_bits.argtypes = [C.c_int] * 5
        L.hmo_mv_cost.restype = C.c_uint32
        L.hmo_mv_cost.argtypes = [C.c_uint32, C.c_uint32]
        L.hmo_dist.restype = C.c_uint32
        L.hmo_dist.argtypes = [C.c_int, _p16, C.c_int, _p16, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.hmo_search_range.restype = None
        L.hmo_search_range.argtypes = [C.c_int] * 9 + [_pi] * 4
        L.hmo_pattern_search.restype = None
        L.hmo_pattern_search.argtypes = [_p16, C.c_int, C.c_int, C.c_int, C.c_int, _p16, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                         _pi, _pi, _pu]
        L.hmo_frac_search.restype = None
        L.hmo_frac_search.argtypes = [_p16, C.c_int, C.c_int, C.c_int, C.c_int, _p16, C.c_int, C.c_int, C.c_int,
                                      C.c_uint32, C.c_int, C.c_int, C.c_int, _pi, _pi, _pi, _pi, _pu]
        L.hmo_extend_border.restype = None
        L.hmo_extend_border.argtypes = [_p16, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.hmo_run_jobs.restype = C.c_double
        L.hmo_run_jobs.argtypes = [_p16, C.c_int, _p16, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_void_p]

    def eg_bits(self, v):
        return self.lib.hmo_eg_bits(int(v))

    def mv_bits(self, x, y, pred, scale):
        return self.lib.hmo_mv_bits(int(x), int(y), int(pred[0]), int(pred[1]), int(scale))

    def mv_cost(self, lambda_cost, x, y, pred, scale):
        return self.lib.hmo_mv_cost(int(lambda_cost), self.mv_bits(x, y, pred, scale))

    def dist(self, kind, org, cur, w, h, bit_depth=8, sub_shift=0):
        (oa, oo, os_), (ca, co, cs) = org, cur
        return self.lib.hmo_dist(kind, _ptr(oa, oo), os_, _ptr(ca, co), cs, w, h, bit_depth, sub_shift)

    def dist_batch(self, kind, org, cur, blocks, bit_depth=8):
        """org / cur: (padded int16 array, offset of sample (0,0), stride); blocks: int32 array (n, 7) of
        org_x, org_y, cur_x, cur_y, w, h, sub_shift."""
        (oa, oo, os_), (ca, co, cs) = org, cur
        blocks = np.ascontiguousarray(blocks, dtype=np.int32)
        out = np.zeros(len(blocks), dtype=np.uint32)
        self.lib.hmo_dist_batch.restype = None
        self.lib.hmo_dist_batch.argtypes = [C.c_int, _p16, C.c_int, _p16, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.hmo_dist_batch(kind, _ptr(oa, oo), os_, _ptr(ca, co), cs, bit_depth, len(blocks), blocks.ctypes.data, out.ctypes.data)
        return out

    def search_range(self, pred, rng, cu_xy, pic_wh, max_cu=64):
        o = [C.c_int() for _ in range(4)]
        self.lib.hmo_search_range(pred[0], pred[1], rng, cu_xy[0], cu_xy[1], pic_wh[0], pic_wh[1], max_cu, max_cu,
                                  *[C.byref(v) for v in o])
        return tuple(v.value for v in o)

    def pattern_search(self, org, w, h, ref, lt, rb, lambda_cost, pred, bit_depth=8):
        (oa, oo, os_), (ra, ro, rs) = org, ref
        mx, my, sad = C.c_int(), C.c_int(), C.c_uint32()
        self.lib.hmo_pattern_search(_ptr(oa, oo), os_, w, h, bit_depth, _ptr(ra, ro), rs, lt[0], lt[1], rb[0], rb[1],
                                    int(lambda_cost), pred[0], pred[1], self.fen, C.byref(mx), C.byref(my), C.byref(sad))
        return (mx.value, my.value), sad.value

    def pattern_search_frac(self, org, w, h, ref, mv_int, lambda_cost, pred, bit_depth=8, lossless=0):
        (oa, oo, os_), (ra, ro, rs) = org, ref
        v = [C.c_int() for _ in range(4)]
        cost = C.c_uint32()
        use_had = 1 if (self.hadme and not lossless) else 0
        self.lib.hmo_frac_search(_ptr(oa, oo), os_, w, h, bit_depth, _ptr(ra, ro), rs, mv_int[0], mv_int[1],
                                 int(lambda_cost), pred[0], pred[1], use_had, *[C.byref(x) for x in v], C.byref(cost))
        return (v[0].value, v[1].value), (v[2].value, v[3].value), cost.value

    def extend_border(self, plane, origin_off, stride, w, h, mx, my):
        self.lib.hmo_extend_border(_ptr(plane, origin_off), stride, w, h, mx, my)

    def mc_dist(self, kind, org, w, h, ref, mv, bit_depth=8):
        """Distortion (0 = SAD, 2 = HADs) between the original PU and its motion-compensated prediction at quarter-pel mv.
        org / ref: (array, offset of the PU's top-left / of the co-located reference sample, stride)."""
        (oa, oo, os_), (ra, ro, rs) = org, ref
        f = self.lib.hmo_mc_dist
        f.restype = C.c_uint32
        f.argtypes = [C.c_int, _p16, C.c_int, _p16, C.c_int] + [C.c_int] * 5
        return f(kind, _ptr(oa, oo), os_, _ptr(ra, ro), rs, w, h, mv[0], mv[1], bit_depth)

    def mc_cand_dist(self, kind, org, w, h, inter_dir, ref0, mv0, ref1, mv1, bit_depth=8, same_picture=0):
        """Prediction error of one merge / AMVP candidate: inter_dir 1 / 2 uni-directional, 3 bi-prediction (addAvg).
        org / ref0 / ref1: (array, offset of the PU's top-left / co-located sample, stride)."""
        (oa, oo, os_), (ra, ro, rs), (rb, rbo, rbs) = org, ref0, ref1
        f = self.lib.hmo_mc_cand_dist
        f.restype = C.c_uint32
        f.argtypes = [C.c_int, _p16, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p16, C.c_int, C.c_int, C.c_int, _p16, C.c_int, C.c_int, C.c_int, C.c_int]
        return f(kind, _ptr(oa, oo), os_, w, h, bit_depth, inter_dir, _ptr(ra, ro), rs, mv0[0], mv0[1], _ptr(rb, rbo), rbs, mv1[0], mv1[1], same_picture)

    def merge_pick(self, dist, bits, lambda_cost):
        d, b = np.ascontiguousarray(dist, dtype=np.uint32), np.ascontiguousarray(bits, dtype=np.uint32)
        cost = C.c_uint32()
        self.lib.hmo_merge_pick.restype = C.c_int
        self.lib.hmo_merge_pick.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_uint32)]
        i = self.lib.hmo_merge_pick(d.ctypes.data, b.ctypes.data, len(d), int(lambda_cost), C.byref(cost))
        return i, cost.value

    def amvp_pick(self, sad, bits, lambda_motion_sad):
        d, b = np.ascontiguousarray(sad, dtype=np.uint32), np.ascontiguousarray(bits, dtype=np.uint32)
        cost = C.c_uint32()
        self.lib.hmo_amvp_pick.restype = C.c_int
        self.lib.hmo_amvp_pick.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_uint32)]
        i = self.lib.hmo_amvp_pick(d.ctypes.data, b.ctypes.data, len(d), int(lambda_motion_sad), C.byref(cost))
        return i, cost.value

    def tz_search(self, org, w, h, ref, lt, rb, lambda_cost, pred, cu_xy, pic_wh, search_range=64, imv=None, bit_depth=8, max_cu=64,
                  first_search_stop=1):
        """xTZSearch (FastSearch=1).  imv: pIntegerMv2Nx2NPred (integer pel) or None; first_search_stop:
        FastMEAssumingSmootherMVEnabled (the encoder's default is on)."""
        (oa, oo, os_), (ra, ro, rs) = org, ref
        mx, my, sad = C.c_int(), C.c_int(), C.c_uint32()
        f = self.lib.hmo_tz_search
        f.restype = None
        f.argtypes = [_p16, C.c_int, C.c_int, C.c_int, C.c_int, _p16, C.c_int] + [C.c_int] * 4 + [C.c_uint32] + [C.c_int] * 13 + [_pi, _pi, _pu]
        f(_ptr(oa, oo), os_, w, h, bit_depth, _ptr(ra, ro), rs, lt[0], lt[1], rb[0], rb[1], int(lambda_cost), pred[0], pred[1],
          self.fen, cu_xy[0], cu_xy[1], pic_wh[0], pic_wh[1], max_cu, search_range, int(first_search_stop),
          0 if imv is None else 1, 0 if imv is None else imv[0], 0 if imv is None else imv[1], C.byref(mx), C.byref(my), C.byref(sad))
        return (mx.value, my.value), sad.value

    def read_luma(self, file_bytes, width, height, pad_x=0, pad_y=0, file_bit_depth=8, internal_bit_depth=8):
        """Luma plane (coded size, no margins) as TVideoIOYuv::read delivers it from a planar file image."""
        buf = np.frombuffer(file_bytes, dtype=np.uint8)
        out = np.zeros((height + pad_y, width + pad_x), dtype=np.int16)
        self.lib.hmo_read_luma.restype = None
        self.lib.hmo_read_luma.argtypes = [C.c_void_p] + [C.c_int] * 7 + [_p16, C.c_int]
        self.lib.hmo_read_luma(buf.ctypes.data, 1 if file_bit_depth > 8 else 0, width, height, pad_x, pad_y,
                               internal_bit_depth - file_bit_depth, internal_bit_depth, _ptr(out), out.shape[1])
        return out

    def run_jobs(self, cur, ref, jobs, bit_depth=8, do_frac=True):
        """cur/ref: (array, offset of sample (0,0), stride).  Returns (results, cpu_seconds)."""
        (ca, co, cs), (ra, ro, rs) = cur, ref
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        out = np.zeros(len(jobs), dtype=RESULT_DTYPE)
        t = self.lib.hmo_run_jobs(_ptr(ca, co), cs, _ptr(ra, ro), rs, bit_depth, jobs.ctypes.data, len(jobs),
                                  self.fen, self.hadme, int(bool(do_frac)), out.ctypes.data)
        return out, t


class Reference(_Base):
    """The unmodified reference.  Raises FileNotFoundError when oracle/_ref/libhmref.so is absent."""

    def __init__(self, fen=1, hadme=1):
        so = _b.build_reference()
        if so is None or not os.path.exists(so):
            raise FileNotFoundError("oracle/_ref/libhmref.so not built (needs /root/reference)")
        self.fen, self.hadme = int(fen), int(hadme)
        L = self.lib = C.CDLL(so)
        self._pfx = "hmref_"; self._h = lambda: [self.h]
        _intra_bind(L, "hmref_", True)
        L.hmref_create.restype = C.c_void_p
        L.hmref_create.argtypes = [C.c_int, C.c_int]
        L.hmref_destroy.argtypes = [C.c_void_p]
        L.hmref_dist.restype = C.c_uint32
        L.hmref_dist.argtypes = [C.c_void_p, C.c_int, _p16, C.c_int, _p16, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.hmref_get_cost.restype = C.c_uint32
        L.hmref_get_cost.argtypes = [C.c_void_p, C.c_uint32] + [C.c_int] * 5
        L.hmref_get_bits.restype = C.c_uint32
        L.hmref_get_bits.argtypes = [C.c_void_p] + [C.c_int] * 5
        L.hmref_pattern_search.restype = None
        L.hmref_pattern_search.argtypes = [C.c_void_p, _p16, C.c_int, C.c_int, C.c_int, C.c_int, _p16, C.c_int,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int,
                                           _pi, _pi, _pu]
        L.hmref_pattern_search_frac.restype = None
        L.hmref_pattern_search_frac.argtypes = [C.c_void_p, C.c_int, _p16, C.c_int, C.c_int, C.c_int, C.c_int, _p16,
                                                C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int,
                                                _pi, _pi, _pi, _pi, _pu]
        L.hmref_run_jobs.restype = C.c_double
        L.hmref_run_jobs.argtypes = [C.c_void_p, _p16, C.c_int, _p16, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                     C.c_void_p]
        self.h = L.hmref_create(self.fen, self.hadme)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.hmref_destroy(self.h)
            self.h = None

    def mv_bits(self, x, y, pred, scale):
        return self.lib.hmref_get_bits(self.h, pred[0], pred[1], scale, x, y)

    def mv_cost(self, lambda_cost, x, y, pred, scale):
        return self.lib.hmref_get_cost(self.h, int(lambda_cost), pred[0], pred[1], scale, x, y)

    def dist(self, kind, org, cur, w, h, bit_depth=8, sub_shift=0):
        (oa, oo, os_), (ca, co, cs) = org, cur
        return self.lib.hmref_dist(self.h, kind, _ptr(oa, oo), os_, _ptr(ca, co), cs, w, h, bit_depth, sub_shift)

    def pattern_search(self, org, w, h, ref, lt, rb, lambda_cost, pred, bit_depth=8):
        (oa, oo, os_), (ra, ro, rs) = org, ref
        mx, my, sad = C.c_int(), C.c_int(), C.c_uint32()
        self.lib.hmref_pattern_search(self.h, _ptr(oa, oo), os_, w, h, bit_depth, _ptr(ra, ro), rs,
                                      lt[0], lt[1], rb[0], rb[1], int(lambda_cost), pred[0], pred[1],
                                      C.byref(mx), C.byref(my), C.byref(sad))
        return (mx.value, my.value), sad.value

    def pattern_search_frac(self, org, w, h, ref, mv_int, lambda_cost, pred, bit_depth=8, lossless=0):
        (oa, oo, os_), (ra, ro, rs) = org, ref
        v = [C.c_int() for _ in range(4)]
        cost = C.c_uint32()
        self.lib.hmref_pattern_search_frac(self.h, lossless, _ptr(oa, oo), os_, w, h, bit_depth, _ptr(ra, ro), rs,
                                           mv_int[0], mv_int[1], int(lambda_cost), pred[0], pred[1],
                                           *[C.byref(x) for x in v], C.byref(cost))
        return (v[0].value, v[1].value), (v[2].value, v[3].value), cost.value

    def mc_dist(self, kind, org, w, h, ref_plane0, pic_wh, margin, pu_xy, mv, bit_depth=8):
        """The reference's own xPredInterBlk + distortion.  ref_plane0: (padded array, offset of sample (0,0), stride)."""
        (oa, oo, os_), (ra, ro, rs) = org, ref_plane0
        f = self.lib.hmref_mc_dist
        f.restype = C.c_uint32
        f.argtypes = [C.c_void_p, C.c_int, _p16, C.c_int, C.c_int, C.c_int, C.c_int, _p16, C.c_int] + [C.c_int] * 7
        return f(self.h, kind, _ptr(oa, oo), os_, w, h, bit_depth, _ptr(ra, ro), rs, pic_wh[0], pic_wh[1], margin, pu_xy[0], pu_xy[1], mv[0], mv[1])

    def mc_bi_dist(self, kind, org, w, h, ref_a0, ref_b0, pic_wh, margin, pu_xy, mv0, mv1, bit_depth=8):
        """The reference's own xPredInterBlk(bi = true) x 2 + xWeightedAverage (TComYuv::addAvg) + distortion.  ref_a0 / ref_b0:
        (padded array, offset of sample (0,0), stride) of the list-0 / list-1 pictures (same stride)."""
        (oa, oo, os_), (ra, ro, rs), (rb, rbo, _) = org, ref_a0, ref_b0
        f = self.lib.hmref_mc_bi_dist
        f.restype = C.c_uint32
        f.argtypes = [C.c_void_p, C.c_int, _p16, C.c_int, C.c_int, C.c_int, C.c_int, _p16, _p16, C.c_int] + [C.c_int] * 9
        return f(self.h, kind, _ptr(oa, oo), os_, w, h, bit_depth, _ptr(ra, ro), _ptr(rb, rbo), rs, pic_wh[0], pic_wh[1], margin, pu_xy[0], pu_xy[1],
                 mv0[0], mv0[1], mv1[0], mv1[1])

    def template_rd_cost(self, bits, dist, lam, bit_depth=8):
        f = self.lib.hmref_template_rd_cost
        f.restype = C.c_uint32
        f.argtypes = [C.c_uint32, C.c_uint32, C.c_double, C.c_int]
        return f(int(bits), int(dist), float(lam), bit_depth)

    def tz_search(self, org, w, h, ref, lt, rb, lambda_cost, pred, cu_xy, pic_wh, search_range=64, imv=None, bit_depth=8, max_cu=64,
                  first_search_stop=1):
        (oa, oo, os_), (ra, ro, rs) = org, ref
        mx, my, sad = C.c_int(), C.c_int(), C.c_uint32()
        f = self.lib.hmref_tz_search
        f.restype = None
        f.argtypes = [C.c_void_p, _p16, C.c_int, C.c_int, C.c_int, C.c_int, _p16, C.c_int] + [C.c_int] * 4 + [C.c_uint32] + [C.c_int] * 12 + [_pi, _pi, _pu]
        f(self.h, _ptr(oa, oo), os_, w, h, bit_depth, _ptr(ra, ro), rs, lt[0], lt[1], rb[0], rb[1], int(lambda_cost), pred[0], pred[1],
          cu_xy[0], cu_xy[1], pic_wh[0], pic_wh[1], max_cu, search_range, int(first_search_stop),
          0 if imv is None else 1, 0 if imv is None else imv[0], 0 if imv is None else imv[1], C.byref(mx), C.byref(my), C.byref(sad))
        return (mx.value, my.value), sad.value

    def read_luma(self, path, width, height, pad_x=0, pad_y=0, file_bit_depth=8, internal_bit_depth=8):
        """The reference's own reader (TVideoIOYuv) on a planar 4:0:0 file."""
        out = np.zeros((height + pad_y, width + pad_x), dtype=np.int16)
        self.lib.hmref_read_luma.restype = C.c_int
        self.lib.hmref_read_luma.argtypes = [C.c_char_p] + [C.c_int] * 6 + [_p16, C.c_int]
        rc = self.lib.hmref_read_luma(path.encode(), width, height, pad_x, pad_y, file_bit_depth, internal_bit_depth, _ptr(out), out.shape[1])
        if rc != 0:
            raise IOError(f"reference reader failed on {path}")
        return out

    def run_jobs(self, cur, ref, jobs, bit_depth=8, do_frac=True):
        (ca, co, cs), (ra, ro, rs) = cur, ref
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        out = np.zeros(len(jobs), dtype=RESULT_DTYPE)
        t = self.lib.hmref_run_jobs(self.h, _ptr(ca, co), cs, _ptr(ra, ro), rs, bit_depth, jobs.ctypes.data, len(jobs),
                                    int(bool(do_frac)), out.ctypes.data)
        return out, t


# ---------------------------------------------------------------------------------------------------------------------
# Canonical all-PU job list in plain Python - so that the reference arm of bench.py and the CPU tests can
# build job lists without loading the product library.  Restates TComDataCU::clipMv (TLibCommon/TComDataCU.cpp:2788-2801),
# TEncSearch::xSetSearchRange (TLibEncoder/TEncSearch.cpp:3765-3781) and the partition geometry of
# TComDataCU::getPartIndexAndSize (TLibCommon/TComDataCU.cpp:1893-1931).
# ---------------------------------------------------------------------------------------------------------------------
def _clip_mv(x, y, cu_x, cu_y, pic_w, pic_h, max_cu):
    off = 8
    hmax, hmin = (pic_w + off - cu_x - 1) * 4, (-max_cu - off - cu_x + 1) * 4
    vmax, vmin = (pic_h + off - cu_y - 1) * 4, (-max_cu - off - cu_y + 1) * 4
    return min(hmax, max(hmin, x)), min(vmax, max(vmin, y))


def _s16(v):
    return ((v + 0x8000) & 0xFFFF) - 0x8000          # TComMv components are Shorts


def py_search_range(pred, search_range, cu_xy, pic_wh, max_cu=64):
    px, py = _clip_mv(pred[0], pred[1], cu_xy[0], cu_xy[1], pic_wh[0], pic_wh[1], max_cu)
    lx, ly = _s16(px - (search_range << 2)), _s16(py - (search_range << 2))
    rx, ry = _s16(px + (search_range << 2)), _s16(py + (search_range << 2))
    lx, ly = _clip_mv(lx, ly, cu_xy[0], cu_xy[1], pic_wh[0], pic_wh[1], max_cu)
    rx, ry = _clip_mv(rx, ry, cu_xy[0], cu_xy[1], pic_wh[0], pic_wh[1], max_cu)
    return lx >> 2, ly >> 2, rx >> 2, ry >> 2


def py_canonical_jobs(pic_w, pic_h, search_range=64, lambda_cost=0, pred=(0, 0), max_cu=64, ctu_first=0, ctu_count=-1):
    """Every CU of depth 0..3 inside the picture; 2Nx2N, 2NxN, Nx2N at every depth plus the four AMP modes at depths 0..2:
    593 PUs per 64x64 CTU, in the order hmb200_build_canonical_jobs emits them."""
    ctus_x, ctus_y = (pic_w + max_cu - 1) // max_cu, (pic_h + max_cu - 1) // max_cu
    n_ctus = ctus_x * ctus_y
    ctu_end = n_ctus if ctu_count < 0 else min(n_ctus, ctu_first + ctu_count)
    rows = []
    for ctu in range(max(0, ctu_first), ctu_end):
        ox, oy = (ctu % ctus_x) * max_cu, (ctu // ctus_x) * max_cu
        s = max_cu
        while s >= 8:
            for cy in range(oy, oy + max_cu, s):
                for cx in range(ox, ox + max_cu, s):
                    if cx + s > pic_w or cy + s > pic_h:
                        continue
                    q = s >> 2
                    pus = [(cx, cy, s, s), (cx, cy, s, s // 2), (cx, cy + s // 2, s, s // 2), (cx, cy, s // 2, s), (cx + s // 2, cy, s // 2, s)]
                    if s > 8:
                        pus += [(cx, cy, s, q), (cx, cy + q, s, s - q), (cx, cy, s, s - q), (cx, cy + s - q, s, q),
                                (cx, cy, q, s), (cx + q, cy, s - q, s), (cx, cy, s - q, s), (cx + s - q, cy, q, s)]
                    lt_x, lt_y, rb_x, rb_y = py_search_range(pred, search_range, (cx, cy), (pic_w, pic_h), max_cu)
                    rows += [(x, y, w, h, lt_x, lt_y, rb_x, rb_y, pred[0], pred[1], int(lambda_cost), 0) for (x, y, w, h) in pus]
            s >>= 1
    return np.array(rows, dtype=JOB_DTYPE) if rows else np.zeros(0, dtype=JOB_DTYPE)
