// Intra first pass: the 35 luma predictions of a block and their Hadamard distortions, for many blocks per launch.
//
// Reference: the mode loop of TEncSearch::estIntraPredQT (TLibEncoder/TEncSearch.cpp:2270-2296): per mode
// filteringIntraReferenceSamples (TLibCommon/TComPattern.cpp:544-568), predIntraAng (TComPrediction.cpp:412-494 ->
// xPredIntraPlanar :756-816, xPredIntraAng :250-410, predIntraGetPredValDC :183-222, xDCPredFiltering :819-848) and
// distParam.DistFunc = xGetHADs (TComRdCost.cpp:380-392, 1526-1593).  The reference samples (availability substitution
// and smoothing, initIntraPatternChType, TComPattern.cpp:115-320) depend on the encoder's reconstruction state and stay on
// the host: a block arrives with its two reference-sample sets as lines.
//
// The reference predicts a whole block per mode with sequential row / column recurrences; here every predicted sample is
// a closed form of the reference lines (the projected side samples of the negative angles included), so the unit of
// work is one (mode, Hadamard tile): a thread predicts the 8x8 (4x4 for 4x4 blocks) samples of its tile, subtracts the
// original, transforms, and adds the rounded tile SATD to the mode's sum in shared memory.  One CTA per block; lines
// and original tile live in shared memory.
#pragma once
#include "hmb200_device.cuh"

namespace hmb200 {

constexpr int INTRA_THREADS = 128;
constexpr int INTRA_MODES = 35;

struct IntraBlockDev {        // 24 bytes
  int32_t x, y;               // block position in the original plane
  int32_t n;                  // 4, 8, 16, 32, 64
  int32_t ref_off;            // first sample of this block's lines in the reference-sample array:
                              // top_unf[2n+1], left_unf[2n+1], top_flt[2n+1], left_flt[2n+1]; index 0 of each = corner
  int32_t flags;              // bit 0: bAbove, bit 1: bLeft (predIntraAng's arguments; the encoder passes both true)
  int32_t pad;
};

__device__ __forceinline__ bool intra_use_filtered(int mode, int log2n) {
  // TComPrediction::m_aucIntraFilter, luma (TComPrediction.cpp:50-57): 4x4 10, 8x8 7, 16x16 1, 32x32 0, 64x64 10
  const int thr = log2n == 2 ? 10 : log2n == 3 ? 7 : log2n == 4 ? 1 : log2n == 5 ? 0 : 10;
  if (mode == 1) return false;
  return min(abs(mode - 10), abs(mode - 26)) > thr;
}

// one predicted sample of an angular mode; mainl / side: lines with the corner at index 0
__device__ __forceinline__ int intra_ang_ref(const int16_t* mainl, const int16_t* side, int t, int inv) {
  return t >= 0 ? (int)mainl[t] : (int)side[(128 - t * inv) >> 8];       // projected side sample (TComPrediction.cpp:318-324)
}

template <int TN, typename OrgT>
__global__ void __launch_bounds__(INTRA_THREADS)
k_intra_modes_had(const IntraBlockDev* __restrict__ blocks, const int32_t* __restrict__ order, const int16_t* __restrict__ refs,
                  uint32_t* __restrict__ out, DevPlane org_plane, int bit_depth) {
  __shared__ int16_t s_line[4][2 * 64 + 2];
  __shared__ int16_t s_org[64 * 64];
  __shared__ uint32_t s_sum[INTRA_MODES];
  __shared__ int s_dc;
  const int bi = order[blockIdx.x];                                      // blocks of one tile size per launch (4x4 blocks use 4x4 tiles)
  const IntraBlockDev b = blocks[bi];
  const int n = b.n, L = 2 * n + 1;
  const int log2n = 31 - __clz(n);
  for (int i = threadIdx.x; i < 4 * L; i += INTRA_THREADS) s_line[i / L][i % L] = refs[b.ref_off + i];
  for (int i = threadIdx.x; i < n * n; i += INTRA_THREADS) {
    const int y = i >> log2n, x = i & (n - 1);
    s_org[i] = (int16_t)*plane_at<OrgT>(org_plane, b.x + x, b.y + y);
  }
  if (threadIdx.x < INTRA_MODES) s_sum[threadIdx.x] = 0;
  __syncthreads();
  const bool above_ok = b.flags & 1, left_ok = (b.flags >> 1) & 1;
  if (threadIdx.x == 0) {                                                // predIntraGetPredValDC on the unfiltered lines
    int sum = 0;
    if (above_ok) for (int i = 0; i < n; i++) sum += s_line[0][1 + i];
    if (left_ok)  for (int i = 0; i < n; i++) sum += s_line[1][1 + i];
    s_dc = (above_ok && left_ok) ? (sum + n) / (2 * n) : (above_ok || left_ok) ? (sum + n / 2) / n : (int)s_line[1][1];
  }
  __syncthreads();
  const int tpr = n / TN, tiles = tpr * tpr;                             // Hadamard tiles per row / per block
  const int maxv = (1 << bit_depth) - 1;
  for (int u = threadIdx.x; u < INTRA_MODES * tiles; u += INTRA_THREADS) {
    const int mode = u / tiles, tile = u - mode * tiles;
    const int tx = (tile % tpr) * TN, ty = (tile / tpr) * TN;
    const bool flt = intra_use_filtered(mode, log2n);
    const int16_t* top = s_line[flt ? 2 : 0];
    const int16_t* left = s_line[flt ? 3 : 1];
    int d[TN * TN];
    if (mode == 0) {                                                     // planar
      const int tr = top[1 + n], bl = left[1 + n];
#pragma unroll
      for (int yy = 0; yy < TN; yy++)
#pragma unroll
        for (int xx = 0; xx < TN; xx++) {
          const int x = tx + xx, y = ty + yy;
          const int v = ((n - 1 - x) * left[1 + y] + (x + 1) * tr + (n - 1 - y) * top[1 + x] + (y + 1) * bl + n) >> (log2n + 1);
          d[yy * TN + xx] = (int)s_org[(y << log2n) + x] - v;
        }
    } else if (mode == 1) {                                              // DC (+ edge filter up to 16x16 when both neighbours exist)
      const int dc = s_dc;
      const bool ef = above_ok && left_ok && n <= 16;
#pragma unroll
      for (int yy = 0; yy < TN; yy++)
#pragma unroll
        for (int xx = 0; xx < TN; xx++) {
          const int x = tx + xx, y = ty + yy;
          int v = dc;
          if (ef) {
            if (x == 0 && y == 0) v = (top[1] + left[1] + 2 * dc + 2) >> 2;
            else if (y == 0)      v = (top[1 + x] + 3 * dc + 2) >> 2;
            else if (x == 0)      v = (left[1 + y] + 3 * dc + 2) >> 2;
          }
          d[yy * TN + xx] = (int)s_org[(y << log2n) + x] - v;
        }
    } else {                                                             // angular
      const bool ver = mode >= 18;
      const int am = ver ? mode - 26 : 10 - mode, aabs = abs(am);
      const int atab = aabs == 0 ? 0 : aabs == 1 ? 2 : aabs == 2 ? 5 : aabs == 3 ? 9 : aabs == 4 ? 13 : aabs == 5 ? 17 : aabs == 6 ? 21 : aabs == 7 ? 26 : 32;
      const int inv = aabs == 0 ? 0 : aabs == 1 ? 4096 : aabs == 2 ? 1638 : aabs == 3 ? 910 : aabs == 4 ? 630 : aabs == 5 ? 482 : aabs == 6 ? 390 : aabs == 7 ? 315 : 256;
      const int angle = am < 0 ? -atab : atab;
      const int16_t* mainl = ver ? top : left;
      const int16_t* side = ver ? left : top;
      const bool ef = angle == 0 && n <= 16;                             // pure vertical / horizontal: first column / row follows the side gradient
#pragma unroll
      for (int yy = 0; yy < TN; yy++)
#pragma unroll
        for (int xx = 0; xx < TN; xx++) {
          const int x = tx + xx, y = ty + yy;
          const int i = ver ? x : y, j = ver ? y : x;                    // i along the main reference, j across
          const int pos = (j + 1) * angle, ip = pos >> 5, fr = pos & 31;
          const int t = i + ip + 1;
          const int a = intra_ang_ref(mainl, side, t, inv);
          const int c = intra_ang_ref(mainl, side, min(t + 1, 2 * n), inv);
          int v = ((32 - fr) * a + fr * c + 16) >> 5;
          if (ef && i == 0) v = min(max(v + (((int)side[j + 1] - (int)side[0]) >> 1), 0), maxv);
          d[yy * TN + xx] = (int)s_org[(y << log2n) + x] - v;
        }
    }
    uint32_t s;
    if constexpr (TN == 8) s = (had8x8_abs(d) + 2) >> 2;                 // TComRdCost.cpp:1520
    else                   s = (had4x4_abs(d) + 1) >> 1;                 // TComRdCost.cpp:1423
    atomicAdd(&s_sum[mode], s);
  }
  __syncthreads();
  if (threadIdx.x < INTRA_MODES) out[(size_t)bi * INTRA_MODES + threadIdx.x] = s_sum[threadIdx.x] >> (bit_depth - 8);
}

}  // namespace hmb200
