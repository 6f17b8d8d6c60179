"""video_codecs_b200 — B200-native HM-16.5 integer-pel motion search + block distortion path.

The product is video_codecs_b200/libhmb200.so (CUDA, sm_100a) behind the C-ABI of include/hmb200.h.  This package
is a thin ctypes mirror of that ABI for tests and benchmarks; it never computes anything itself and never falls
back to a CPU implementation: if the shared library is missing, importing `video_codecs_b200.api` raises.
"""
from .api import (HMB200, HMB200Error, JOB_DTYPE, RESULT_DTYPE, DIST_DESC_DTYPE,  # noqa: F401
                  FLAG_FEN, FLAG_HADME, FLAG_FRAC, FLAG_TZ, FLAG_TZ_STOP, TZ_EXTRA_DTYPE, MC_DESC_DTYPE, INTRA_BLOCK_DTYPE, DF_SAD, DF_SSE, DF_HADS, DF_SADS, lib_path,
                  RESULT16_DTYPE, widen_results16, MC_CAND_DTYPE)
