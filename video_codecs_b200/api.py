"""ctypes mirror of include/hmb200.h.  Names follow the header one to one."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

FLAG_FEN, FLAG_HADME, FLAG_FRAC, FLAG_TZ, FLAG_TZ_STOP = 1, 2, 4, 8, 16
DF_SAD, DF_SSE, DF_HADS, DF_SADS = 0, 1, 2, 3
PLANE_ORG, PLANE_REC = 0, 1

JOB_DTYPE = np.dtype([("pu_x", "<i4"), ("pu_y", "<i4"), ("w", "<i4"), ("h", "<i4"),
                      ("lt_x", "<i4"), ("lt_y", "<i4"), ("rb_x", "<i4"), ("rb_y", "<i4"),
                      ("pred_x", "<i4"), ("pred_y", "<i4"), ("lambda_cost", "<u4"), ("reserved", "<i4")])
RESULT_DTYPE = np.dtype([("mv_x", "<i4"), ("mv_y", "<i4"), ("sad", "<u4"),
                         ("half_x", "<i4"), ("half_y", "<i4"), ("qter_x", "<i4"), ("qter_y", "<i4"),
                         ("frac_cost", "<u4")])
DIST_DESC_DTYPE = np.dtype([("org_plane", "<i4"), ("org_x", "<i4"), ("org_y", "<i4"),
                            ("cur_plane", "<i4"), ("cur_x", "<i4"), ("cur_y", "<i4"),
                            ("w", "<i4"), ("h", "<i4"), ("sub_shift", "<i4"), ("reserved", "<i4")])
MC_DESC_DTYPE = np.dtype([("pu_x", "<i4"), ("pu_y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("mv_x", "<i4"), ("mv_y", "<i4")])
MC_CAND_DTYPE = np.dtype([("pu_x", "<i4"), ("pu_y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("inter_dir", "<i4"), ("mv0_x", "<i4"), ("mv0_y", "<i4"),
                          ("ref0_plane", "<i4"), ("mv1_x", "<i4"), ("mv1_y", "<i4"), ("ref1_plane", "<i4"), ("bits", "<i4")])
assert MC_CAND_DTYPE.itemsize == 48
INTRA_BLOCK_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("n", "<i4"), ("ref_off", "<i4"), ("flags", "<i4"), ("reserved", "<i4")])
TZ_EXTRA_DTYPE = np.dtype([("cu_x", "<i4"), ("cu_y", "<i4"), ("has_imv", "<i4"), ("imv_x", "<i4"), ("imv_y", "<i4"),
                           ("reserved", "<i4", (3,))])
RESULT16_DTYPE = np.dtype([("mv_x", "<i2"), ("mv_y", "<i2"), ("sad", "<u4"), ("half_x", "i1"), ("half_y", "i1"), ("qter_x", "i1"),
                           ("qter_y", "i1"), ("frac_cost", "<u4")])
assert TZ_EXTRA_DTYPE.itemsize == 32 and RESULT16_DTYPE.itemsize == 16
assert JOB_DTYPE.itemsize == 48 and RESULT_DTYPE.itemsize == 32 and DIST_DESC_DTYPE.itemsize == 40


class HMB200Error(RuntimeError):
    pass


class _Mv(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32)]


class _DistParam(C.Structure):
    _fields_ = [("pOrg", C.c_void_p), ("pCur", C.c_void_p), ("iStrideOrg", C.c_int32), ("iStrideCur", C.c_int32),
                ("iRows", C.c_int32), ("iCols", C.c_int32), ("iStep", C.c_int32), ("func", C.c_int32),
                ("bitDepth", C.c_int32), ("bApplyWeight", C.c_int32), ("iSubShift", C.c_int32)]


class _Pattern(C.Structure):
    _fields_ = [("roi", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("stride", C.c_int32),
                ("bit_depth", C.c_int32)]


class _CostState(C.Structure):
    _fields_ = [("lambda_cost", C.c_uint32), ("pred", _Mv)]


def lib_path():
    """video_codecs_b200/libhmb200.so; HMB200_LIB names another build of the same sources (A/B experiments with compile-time knobs)."""
    return os.environ.get("HMB200_LIB") or os.path.join(HERE, "libhmb200.so")


def _load():
    path = lib_path()
    if not os.path.exists(path):
        raise HMB200Error(f"{path} is missing: build it with `python -m video_codecs_b200.build` "
                          "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp, i32, u32 = C.c_void_p, C.c_int, C.c_uint32
    sig = {
        "hmb200_init": (i32, [i32]), "hmb200_shutdown": (None, []), "hmb200_last_error": (C.c_char_p, []),
        "hmb200_launch_count": (C.c_uint64, []),
        "hmb200_one_call_stats": (None, [C.POINTER(C.c_uint64)] * 3),
        "hmb200_host_alloc": (vp, [C.c_size_t]), "hmb200_host_free": (None, [vp]),
        "hmb200_motion_lambda_cost": (u32, [C.c_double]),
        "hmb200_set_search_range": (None, [_Mv, i32, i32, i32, i32, i32, i32, i32, C.POINTER(_Mv), C.POINTER(_Mv)]),
        "hmb200_build_canonical_jobs": (i32, [i32, i32, i32, i32, u32, _Mv, i32, i32, vp, i32]),
        "hmb200_build_canonical_jobs_rect": (i32, [i32, i32, i32, i32, u32, _Mv, i32, i32, i32, i32, vp, i32]),
        "hmb200_tile_column_range": (i32, [i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]),
        "hmb200_register_plane": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32]),
        "hmb200_register_plane_u8": (i32, [vp, i32, i32, i32, i32, i32, i32, i32]),
        "hmb200_register_plane_u16": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32]),
        "hmb200_register_plane_yuv": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32]),
        "hmb200_read_plane": (i32, [i32, vp, i32]), "hmb200_release_plane": (None, [i32]),
        "hmb200_dist": (u32, [C.POINTER(_DistParam)]),
        "hmb200_dist_batch": (i32, [i32, i32, i32, vp, vp]),
        "hmb200_mc_dist_batch": (i32, [i32, i32, i32, i32, vp, vp]),
        "hmb200_mc_cand_dist_batch": (i32, [i32, i32, i32, vp, vp]),
        "hmb200_merge_estimation_batch": (i32, [i32, i32, vp, vp, i32, u32, vp, vp, vp]),
        "hmb200_amvp_estimation_batch": (i32, [i32, i32, vp, vp, u32, vp, vp, vp]),
        "hmb200_intra_modes_had_batch": (i32, [i32, i32, vp, vp, i32, vp]),
        "hmb200_intra_modes_had": (i32, [vp, i32, vp, vp, i32, i32, i32, i32, vp]),
        "hmb200_pattern_search": (i32, [C.POINTER(_Pattern), vp, i32, _Mv, _Mv, C.POINTER(_CostState), i32,
                                        C.POINTER(_Mv), C.POINTER(u32)]),
        "hmb200_pattern_search_and_refine": (i32, [C.POINTER(_Pattern), vp, i32, _Mv, _Mv, C.POINTER(_CostState), i32,
                                                   C.POINTER(_Mv), C.POINTER(u32), C.POINTER(_Mv), C.POINTER(_Mv), C.POINTER(u32)]),
        "hmb200_pattern_search_tz": (i32, [C.POINTER(_Pattern), vp, i32, _Mv, _Mv, C.POINTER(_CostState), i32, vp, i32, i32, i32, i32,
                                           C.POINTER(_Mv), C.POINTER(u32)]),
        "hmb200_pattern_search_tz_and_refine": (i32, [C.POINTER(_Pattern), vp, i32, _Mv, _Mv, C.POINTER(_CostState), i32, vp, i32, i32, i32, i32,
                                                      C.POINTER(_Mv), C.POINTER(u32), C.POINTER(_Mv), C.POINTER(_Mv), C.POINTER(u32)]),
        "hmb200_pattern_search_frac": (i32, [i32, C.POINTER(_Pattern), vp, i32, _Mv, C.POINTER(_CostState), i32,
                                             C.POINTER(_Mv), C.POINTER(_Mv), C.POINTER(u32)]),
        "hmb200_me_jobs": (i32, [i32, i32, vp, i32, i32, vp]),
        "hmb200_me_ctu_row": (i32, [i32, i32, i32, i32, vp, i32, i32, vp]),
        "hmb200_prepare_jobs": (vp, [vp, i32, i32, i32]), "hmb200_free_prepared": (None, [vp]),
        "hmb200_run_prepared": (i32, [vp, i32, i32]), "hmb200_fetch_results": (i32, [vp, vp]),
        "hmb200_fetch_results_async": (i32, [vp, vp]), "hmb200_fetch_wait": (i32, [vp]),
        "hmb200_sync": (i32, []),
        "hmb200_last_timing": (i32, [C.POINTER(C.c_float)] * 3),
        "hmb200_prepared_work": (i32, [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "hmb200_canonical_tz_extra": (None, [vp, i32, vp]),
        "hmb200_prepared_set_tz": (i32, [vp, vp, i32, i32, i32, i32]),
        "hmb200_tz_jobs": (i32, [i32, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
        "hmb200_prepared_executed_work": (i32, [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "hmb200_prepared_unique_work": (i32, [vp, C.POINTER(C.c_uint64)]),
        "hmb200_fetch_results16_async": (i32, [vp, vp]),
        "hmb200_ctx_create": (vp, [i32]), "hmb200_ctx_destroy": (None, [vp]), "hmb200_ctx_set_current": (i32, [vp]),
        "hmb200_ctx_get_current": (vp, []), "hmb200_ctx_device": (i32, [vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)          # AttributeError here == the library does not export what the header declares
        f.restype, f.argtypes = res, args
    return L, sorted(sig)


def _addr(a, off=0):
    return a.ctypes.data + a.dtype.itemsize * int(off)


def widen_results16(r16):
    """RESULT16_DTYPE -> RESULT_DTYPE (same values)."""
    out = np.zeros(len(r16), dtype=RESULT_DTYPE)
    for f in RESULT_DTYPE.names:
        out[f] = r16[f]
    return out


class HMB200:
    """Thin wrapper; all state lives in the C library's contexts (default context: init(); more: ctx_create())."""

    def __init__(self):
        self.lib, self.exported = _load()

    # -- lifetime --------------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc < 0:
            raise HMB200Error(f"hmb200 error {rc}: {self.lib.hmb200_last_error().decode()}")
        return rc

    def init(self, device=0):
        self._check(self.lib.hmb200_init(int(device)))

    def shutdown(self):
        self.lib.hmb200_shutdown()

    # -- contexts: one per GPU, each bound to the host thread that drives it (hmb200_ctx_*) ------------------------------
    def ctx_create(self, device):
        h = self.lib.hmb200_ctx_create(int(device))
        if not h:
            raise HMB200Error(self.lib.hmb200_last_error().decode())
        return h

    def ctx_set_current(self, ctx):
        self._check(self.lib.hmb200_ctx_set_current(ctx))

    def ctx_destroy(self, ctx):
        self.lib.hmb200_ctx_destroy(ctx)

    def host_array(self, n, dtype):
        """numpy array of n records in page-locked memory (hmb200_host_alloc); keep the returned array alive, free with host_free."""
        dtype = np.dtype(dtype)
        ptr = self.lib.hmb200_host_alloc(max(1, n) * dtype.itemsize)
        if not ptr:
            raise HMB200Error(self.lib.hmb200_last_error().decode())
        buf = (C.c_uint8 * (max(1, n) * dtype.itemsize)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=n)
        arr.flags.writeable = True
        self._host_ptrs = getattr(self, "_host_ptrs", {})
        self._host_ptrs[arr.ctypes.data] = ptr
        return arr

    def host_free(self, arr):
        ptr = getattr(self, "_host_ptrs", {}).pop(arr.ctypes.data, None)
        if ptr:
            self.lib.hmb200_host_free(ptr)

    def launch_count(self):
        return int(self.lib.hmb200_launch_count())

    def one_call_stats(self):
        """(hmb200_pattern_search_and_refine calls, whole-CU launches among them, calls answered from a whole-CU launch)."""
        v = [C.c_uint64() for _ in range(3)]
        self.lib.hmb200_one_call_stats(*[C.byref(x) for x in v])
        return tuple(int(x.value) for x in v)

    # -- host logic ------------------------------------------------------------------------------------------------
    def motion_lambda_cost(self, lam):
        return int(self.lib.hmb200_motion_lambda_cost(float(lam)))

    def set_search_range(self, pred, search_range, cu_xy, pic_wh, max_cu=64):
        lt, rb = _Mv(), _Mv()
        self.lib.hmb200_set_search_range(_Mv(*pred), search_range, cu_xy[0], cu_xy[1], pic_wh[0], pic_wh[1], max_cu, max_cu,
                                         C.byref(lt), C.byref(rb))
        return (lt.x, lt.y, rb.x, rb.y)

    def build_canonical_jobs(self, pic_w, pic_h, search_range=64, lambda_cost=0, pred=(0, 0), max_cu=64,
                             ctu_first=0, ctu_count=-1):
        n = self._check(self.lib.hmb200_build_canonical_jobs(pic_w, pic_h, max_cu, search_range, int(lambda_cost), _Mv(*pred),
                                                             ctu_first, ctu_count, None, 0))
        jobs = np.zeros(n, dtype=JOB_DTYPE)
        self._check(self.lib.hmb200_build_canonical_jobs(pic_w, pic_h, max_cu, search_range, int(lambda_cost), _Mv(*pred),
                                                         ctu_first, ctu_count, jobs.ctypes.data, n))
        return jobs

    def tile_column_range(self, pic_w, n_columns, column, max_cu=64):
        a, b = C.c_int(), C.c_int()
        self._check(self.lib.hmb200_tile_column_range(pic_w, max_cu, n_columns, column, C.byref(a), C.byref(b)))
        return a.value, b.value

    def build_canonical_jobs_rect(self, pic_w, pic_h, ctu_x, ctu_y, search_range=64, lambda_cost=0, pred=(0, 0), max_cu=64):
        """ctu_x / ctu_y: (first, end) CTU ranges."""
        args = (pic_w, pic_h, max_cu, search_range, int(lambda_cost), _Mv(*pred), ctu_x[0], ctu_x[1], ctu_y[0], ctu_y[1])
        n = self._check(self.lib.hmb200_build_canonical_jobs_rect(*args, None, 0))
        jobs = np.zeros(n, dtype=JOB_DTYPE)
        self._check(self.lib.hmb200_build_canonical_jobs_rect(*args, jobs.ctypes.data, n))
        return jobs

    # -- planes ----------------------------------------------------------------------------------------------------
    def register_plane(self, padded, width, height, margin_x, margin_y, bit_depth=8, kind=PLANE_REC, poc=0):
        """padded: C-contiguous int16 array of shape (height + 2*margin_y, width + 2*margin_x) (a TComPicYuv luma
        buffer).  The caller must keep it alive while 1:1 entries refer to it by pointer."""
        assert padded.dtype == np.int16 and padded.flags["C_CONTIGUOUS"]
        stride = padded.shape[1]
        assert padded.shape[0] == height + 2 * margin_y and stride >= width + 2 * margin_x
        origin = _addr(padded, margin_y * stride + margin_x)
        return self._check(self.lib.hmb200_register_plane(origin, stride, width, height, margin_x, margin_y, bit_depth, kind, poc))

    def register_plane_u8(self, samples, margin_x=80, margin_y=80, kind=PLANE_REC, poc=0):
        assert samples.dtype == np.uint8 and samples.ndim == 2 and samples.strides[1] == 1
        h, w = samples.shape
        return self._check(self.lib.hmb200_register_plane_u8(samples.ctypes.data, samples.strides[0], w, h, margin_x, margin_y, kind, poc))

    def register_plane_u16(self, samples, bit_depth, margin_x=80, margin_y=80, kind=PLANE_REC, poc=0):
        assert samples.dtype == np.uint16 and samples.ndim == 2 and samples.strides[1] == 2
        h, w = samples.shape
        return self._check(self.lib.hmb200_register_plane_u16(samples.ctypes.data, samples.strides[0] // 2, w, h, margin_x, margin_y,
                                                              bit_depth, kind, poc))

    def register_plane_yuv(self, file_luma, width, height, pad_x=0, pad_y=0, file_bit_depth=8, internal_bit_depth=8,
                           margin_x=80, margin_y=80, kind=PLANE_ORG, poc=0):
        """file_luma: bytes-like / uint8 array holding width*height luma samples of a planar file (16-bit LE when
        file_bit_depth > 8)."""
        buf = np.frombuffer(file_luma, dtype=np.uint8) if not isinstance(file_luma, np.ndarray) else np.ascontiguousarray(file_luma).view(np.uint8)
        is16 = 1 if file_bit_depth > 8 else 0
        assert buf.size >= width * height * (2 if is16 else 1)
        return self._check(self.lib.hmb200_register_plane_yuv(buf.ctypes.data, is16, width, height, pad_x, pad_y, file_bit_depth,
                                                              internal_bit_depth, margin_x, margin_y, kind, poc))

    def read_plane(self, plane_id, width, height, margin_x, margin_y):
        out = np.zeros((height + 2 * margin_y, width + 2 * margin_x), dtype=np.int16)
        self._check(self.lib.hmb200_read_plane(plane_id, _addr(out, margin_y * out.shape[1] + margin_x), out.shape[1]))
        return out

    def release_plane(self, plane_id):
        self.lib.hmb200_release_plane(int(plane_id))

    # -- distortion table ------------------------------------------------------------------------------------------
    def dist(self, func, org, cur, w, h, bit_depth=8, sub_shift=0):
        """org/cur: (int16 array, element offset, stride) — the FpDistFunc-compatible single call."""
        (oa, oo, os_), (ca, co, cs) = org, cur
        p = _DistParam(_addr(oa, oo), _addr(ca, co), os_, cs, h, w, 1, func, bit_depth, 0, sub_shift)
        return int(self.lib.hmb200_dist(C.byref(p)))

    def dist_batch(self, func, bit_depth, descs):
        descs = np.ascontiguousarray(descs, dtype=DIST_DESC_DTYPE)
        out = np.zeros(len(descs), dtype=np.uint32)
        self._check(self.lib.hmb200_dist_batch(func, bit_depth, len(descs), descs.ctypes.data, out.ctypes.data))
        return out

    def mc_dist_batch(self, cur_plane, ref_plane, func, descs):
        descs = np.ascontiguousarray(descs, dtype=MC_DESC_DTYPE)
        out = np.zeros(len(descs), dtype=np.uint32)
        self._check(self.lib.hmb200_mc_dist_batch(cur_plane, ref_plane, func, len(descs), descs.ctypes.data, out.ctypes.data))
        return out

    def mc_cand_dist_batch(self, cur_plane, func, cands):
        cands = np.ascontiguousarray(cands, dtype=MC_CAND_DTYPE)
        out = np.zeros(len(cands), dtype=np.uint32)
        self._check(self.lib.hmb200_mc_cand_dist_batch(cur_plane, func, len(cands), cands.ctypes.data, out.ctypes.data))
        return out

    def _estimation(self, fn, cur_plane, cand_first, cands, *mid):
        cands = np.ascontiguousarray(cands, dtype=MC_CAND_DTYPE)
        first = np.ascontiguousarray(cand_first, dtype=np.int32)
        n_pu = len(first) - 1
        best, cost, dist = np.zeros(n_pu, dtype=np.uint32), np.zeros(n_pu, dtype=np.uint32), np.zeros(len(cands), dtype=np.uint32)
        self._check(fn(cur_plane, n_pu, first.ctypes.data, cands.ctypes.data, *mid, best.ctypes.data, cost.ctypes.data, dist.ctypes.data))
        return best, cost, dist

    def merge_estimation_batch(self, cur_plane, cand_first, cands, use_hadme, lambda_cost):
        """xMergeEstimation's candidate loop for many PUs: returns (merge index, cost, per-candidate error)."""
        return self._estimation(self.lib.hmb200_merge_estimation_batch, cur_plane, cand_first, cands, int(use_hadme), int(lambda_cost))

    def amvp_estimation_batch(self, cur_plane, cand_first, cands, lambda_motion_sad):
        """xEstimateMvPredAMVP's candidate loop (xGetTemplateCost) for many PUs: returns (MVP index, cost, per-candidate SAD)."""
        return self._estimation(self.lib.hmb200_amvp_estimation_batch, cur_plane, cand_first, cands, int(lambda_motion_sad))

    # -- intra first pass ------------------------------------------------------------------------------------------
    def intra_modes_had_batch(self, org_plane, blocks, refs):
        """blocks: INTRA_BLOCK_DTYPE records; refs: int16 array holding every block's four reference lines (see the
        header).  Returns a (len(blocks), 35) uint32 array: the first-pass Hadamard distortion of every mode."""
        blocks = np.ascontiguousarray(blocks, dtype=INTRA_BLOCK_DTYPE)
        refs = np.ascontiguousarray(refs, dtype=np.int16)
        out = np.zeros((len(blocks), 35), dtype=np.uint32)
        self._check(self.lib.hmb200_intra_modes_had_batch(org_plane, len(blocks), blocks.ctypes.data, refs.ctypes.data, refs.size,
                                                          out.ctypes.data))
        return out

    def intra_modes_had(self, org, ref_unf, ref_flt, n, bit_depth=8, above=1, left=1):
        """1:1 form: org = (array, offset, stride); ref_unf / ref_flt: (2n+1) x (2n+1) int16 predictor buffers."""
        (oa, oo, os_) = org
        ref_unf = np.ascontiguousarray(ref_unf, dtype=np.int16); ref_flt = np.ascontiguousarray(ref_flt, dtype=np.int16)
        out = np.zeros(35, dtype=np.uint32)
        self._check(self.lib.hmb200_intra_modes_had(_addr(oa, oo), os_, ref_unf.ctypes.data, ref_flt.ctypes.data, n, bit_depth,
                                                    above, left, out.ctypes.data))
        return out

    # -- 1:1 searches ----------------------------------------------------------------------------------------------
    def pattern_search(self, org, w, h, ref, lt, rb, lambda_cost, pred, bit_depth=8, flags=FLAG_FEN):
        """org: (array, offset, stride) host pattern; ref: (registered padded array, offset of the co-located sample, stride)."""
        (oa, oo, os_), (ra, ro, rs) = org, ref
        key = _Pattern(_addr(oa, oo), w, h, os_, bit_depth)
        cs = _CostState(int(lambda_cost), _Mv(*pred))
        mv, sad = _Mv(), C.c_uint32()
        self._check(self.lib.hmb200_pattern_search(C.byref(key), _addr(ra, ro), rs, _Mv(*lt), _Mv(*rb), C.byref(cs), flags,
                                                   C.byref(mv), C.byref(sad)))
        return (mv.x, mv.y), sad.value

    def pattern_search_and_refine(self, org, w, h, ref, lt, rb, lambda_cost, pred, bit_depth=8, flags=FLAG_FEN | FLAG_HADME):
        """xPatternSearch + xPatternSearchFracDIF of one PU in one device round trip: ((mv), sad, (half), (qter), frac_cost)."""
        (oa, oo, os_), (ra, ro, rs) = org, ref
        key = _Pattern(_addr(oa, oo), w, h, os_, bit_depth)
        cs = _CostState(int(lambda_cost), _Mv(*pred))
        mv, sad, half, qter, cost = _Mv(), C.c_uint32(), _Mv(), _Mv(), C.c_uint32()
        self._check(self.lib.hmb200_pattern_search_and_refine(C.byref(key), _addr(ra, ro), rs, _Mv(*lt), _Mv(*rb), C.byref(cs), flags,
                                                              C.byref(mv), C.byref(sad), C.byref(half), C.byref(qter), C.byref(cost)))
        return (mv.x, mv.y), sad.value, (half.x, half.y), (qter.x, qter.y), cost.value

    def pattern_search_tz(self, org, w, h, ref, lt, rb, lambda_cost, pred, cu_xy, pic_wh, search_range=64, imv=None, bit_depth=8,
                          flags=FLAG_FEN, max_cu=64):
        (oa, oo, os_), (ra, ro, rs) = org, ref
        key = _Pattern(_addr(oa, oo), w, h, os_, bit_depth)
        cs = _CostState(int(lambda_cost), _Mv(*pred))
        extra = np.zeros(1, dtype=TZ_EXTRA_DTYPE)
        extra[0] = (cu_xy[0], cu_xy[1], 0 if imv is None else 1, 0 if imv is None else imv[0], 0 if imv is None else imv[1], (0, 0, 0))
        mv, sad = _Mv(), C.c_uint32()
        self._check(self.lib.hmb200_pattern_search_tz(C.byref(key), _addr(ra, ro), rs, _Mv(*lt), _Mv(*rb), C.byref(cs), flags,
                                                      extra.ctypes.data, pic_wh[0], pic_wh[1], max_cu, search_range,
                                                      C.byref(mv), C.byref(sad)))
        return (mv.x, mv.y), sad.value

    def pattern_search_frac(self, org, w, h, ref, mv_int, lambda_cost, pred, bit_depth=8, flags=FLAG_HADME, lossless=0):
        (oa, oo, os_), (ra, ro, rs) = org, ref
        key = _Pattern(_addr(oa, oo), w, h, os_, bit_depth)
        cs = _CostState(int(lambda_cost), _Mv(*pred))
        half, qter, cost = _Mv(), _Mv(), C.c_uint32()
        self._check(self.lib.hmb200_pattern_search_frac(lossless, C.byref(key), _addr(ra, ro), rs, _Mv(*mv_int), C.byref(cs), flags,
                                                        C.byref(half), C.byref(qter), C.byref(cost)))
        return (half.x, half.y), (qter.x, qter.y), cost.value

    # -- batched searches ------------------------------------------------------------------------------------------
    def me_jobs(self, cur_plane, ref_plane, jobs, flags=FLAG_FEN | FLAG_HADME | FLAG_FRAC):
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        out = np.zeros(len(jobs), dtype=RESULT_DTYPE)
        self._check(self.lib.hmb200_me_jobs(cur_plane, ref_plane, jobs.ctypes.data, len(jobs), flags, out.ctypes.data))
        return out

    def me_ctu_row(self, cur_plane, ref_plane, ctu_row, jobs, flags=FLAG_FEN | FLAG_HADME | FLAG_FRAC, max_cu=64):
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        out = np.zeros(len(jobs), dtype=RESULT_DTYPE)
        self._check(self.lib.hmb200_me_ctu_row(cur_plane, ref_plane, ctu_row, max_cu, jobs.ctypes.data, len(jobs), flags, out.ctypes.data))
        return out

    def canonical_tz_extra(self, jobs):
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        extra = np.zeros(len(jobs), dtype=TZ_EXTRA_DTYPE)
        self.lib.hmb200_canonical_tz_extra(jobs.ctypes.data, len(jobs), extra.ctypes.data)
        return extra

    def tz_jobs(self, cur_plane, ref_plane, jobs, extra, pic_wh, search_range=64, flags=FLAG_FEN | FLAG_HADME | FLAG_FRAC, max_cu=64):
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        extra = np.ascontiguousarray(extra, dtype=TZ_EXTRA_DTYPE)
        assert len(extra) == len(jobs)
        out = np.zeros(len(jobs), dtype=RESULT_DTYPE)
        self._check(self.lib.hmb200_tz_jobs(cur_plane, ref_plane, jobs.ctypes.data, extra.ctypes.data, len(jobs), pic_wh[0], pic_wh[1],
                                            max_cu, search_range, flags, out.ctypes.data))
        return out

    def prepare_jobs(self, jobs, flags=FLAG_FEN | FLAG_HADME | FLAG_FRAC, bit_depth=8):
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        h = self.lib.hmb200_prepare_jobs(jobs.ctypes.data, len(jobs), flags, bit_depth)
        if not h:
            raise HMB200Error(self.lib.hmb200_last_error().decode())
        return Prepared(self, h, len(jobs))

    def sync(self):
        self._check(self.lib.hmb200_sync())

    def last_timing(self):
        """CUDA-event times of the last hmb200_run_prepared (total / search / refinement) or hmb200_dist_batch (total = its kernel)."""
        t = [C.c_float() for _ in range(3)]
        self._check(self.lib.hmb200_last_timing(*[C.byref(x) for x in t]))
        return {"total_ms": t[0].value, "search_ms": t[1].value, "frac_ms": t[2].value}


class Prepared:
    def __init__(self, owner, handle, n):
        self.o, self.h, self.n = owner, handle, n

    def set_tz(self, extra, pic_wh, search_range=64, max_cu=64):
        extra = np.ascontiguousarray(extra, dtype=TZ_EXTRA_DTYPE)
        assert len(extra) == self.n
        self.o._check(self.o.lib.hmb200_prepared_set_tz(self.h, extra.ctypes.data, pic_wh[0], pic_wh[1], max_cu, search_range))

    def run(self, cur_plane, ref_plane):
        self.o._check(self.o.lib.hmb200_run_prepared(self.h, cur_plane, ref_plane))

    def fetch(self, out=None):
        if out is None:
            out = np.zeros(self.n, dtype=RESULT_DTYPE)
        self.o._check(self.o.lib.hmb200_fetch_results(self.h, out.ctypes.data))
        return out

    def fetch_async(self, out):
        """D2H of the last run's results on the copy stream; `out` must be a HMB200.host_array."""
        self.o._check(self.o.lib.hmb200_fetch_results_async(self.h, out.ctypes.data))

    def fetch16_async(self, out):
        """Same with 16-byte records (RESULT16_DTYPE), packed on the device behind the run."""
        assert out.dtype == RESULT16_DTYPE
        self.o._check(self.o.lib.hmb200_fetch_results16_async(self.h, out.ctypes.data))

    def fetch_wait(self):
        self.o._check(self.o.lib.hmb200_fetch_wait(self.h))

    def timing(self):
        t = [C.c_float() for _ in range(3)]
        self.o._check(self.o.lib.hmb200_last_timing(*[C.byref(x) for x in t]))
        return {"total_ms": t[0].value, "search_ms": t[1].value, "frac_ms": t[2].value}

    def work(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.o._check(self.o.lib.hmb200_prepared_work(self.h, C.byref(a), C.byref(b)))
        c, d = C.c_uint64(), C.c_uint64()
        self.o._check(self.o.lib.hmb200_prepared_executed_work(self.h, C.byref(c), C.byref(d)))
        e = C.c_uint64()
        self.o._check(self.o.lib.hmb200_prepared_unique_work(self.h, C.byref(e)))
        return {"cand_sads": a.value, "abs_diffs": b.value, "abs_diffs_executed": c.value, "pus_fused": d.value,
                "abs_diffs_unique": e.value}

    def free(self):
        if self.h:
            self.o.lib.hmb200_free_prepared(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
