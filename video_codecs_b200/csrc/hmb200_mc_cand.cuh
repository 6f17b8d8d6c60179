// Prediction error of merge / AMVP candidates, batched: what TEncSearch::xGetInterPredictionError (TLibEncoder/
// TEncSearch.cpp:2805-2826) and TEncSearch::xGetTemplateCost (:3619-3658) compute one candidate at a time.
//
//   uni-directional candidate   TComPrediction::xPredInterBlk, bi = false (TLibCommon/TComPrediction.cpp:668-706): pixels
//   bi-directional candidate    xPredInterBlk with bi = true from each list (the block stays in the 14-bit intermediate
//                               domain: every filter call has isLast = false, TComInterpolationFilter.cpp:196-251), then
//                               TComYuv::addAvg (TLibCommon/TComYuv.cpp:352-407): clip((a + b + offset) >> shift),
//                               shift = max(2, 14 - bitDepth) + 1, offset = 2^(shift-1) + 2 * 8192
//   distortion                  HADs (xGetHADs, per-tile rounding) or SAD of original - prediction
//
// Unit of work = one (candidate, Hadamard tile), like the refinement kernels: a thread interpolates the samples of its tile
// (separable 8-tap filter; the pass-through taps {0,0,0,64,0,0,0,0} reproduce filterCopy and the single-pass branches exactly,
// see DESIGN.md 3.5), transforms and adds the tile's distortion to dist[candidate].  Candidates may refer to different
// reference planes, so every candidate carries device pointers to its co-located samples.
#pragma once
#include "hmb200_frac.cuh"

namespace hmb200 {

struct McCandDev {            // 64 bytes
  const void* org;            // original PU, top-left sample
  const void* ref0;           // list-0 / list-1 reference sample co-located with the PU's top-left (nullptr: list unused)
  const void* ref1;
  int32_t org_pitch, ref0_pitch, ref1_pitch;    // in samples
  int32_t mv0_x, mv0_y, mv1_x, mv1_y;           // quarter pel, already clipped by the caller (TComDataCU::clipMv)
  int32_t w, h;
  int32_t pad;
};
static_assert(sizeof(McCandDev) == 64, "McCandDev layout");

// MODE 0: out = org - clip(uni prediction);  MODE 1: out = bi intermediate of this list;
// MODE 2: out = org - clip((out + bi intermediate of this list + offset) >> shift)      (addAvg, then the difference)
template <typename T, int N, int MODE>
__device__ __forceinline__ void mc_tile_pass(const T* ref, int ref_pitch, const T* org, int org_pitch, int fx, int fy, int head,
                                             int maxv, int (&out)[N * N]) {
  int th[8], tv[8];
  frac_load_taps(fx, th);
  frac_load_taps(fy, tv);
  const int hshift = 6 - head, hoff = 8192 << hshift;
  const int vshift = 6 + head, voff = (1 << (vshift - 1)) + (8192 << 6);          // uni: filter<> isLast after an isFirst pass
  const int ashift = head + 1, aoff = (1 << (ashift - 1)) + 2 * 8192;             // addAvg
#pragma unroll 1
  for (int cs = 0; cs < N; cs += 4) {
    int hh[N + 7][4];
#pragma unroll
    for (int r = 0; r < N + 7; r++) {
      int px[11];
      frac_load_row<T>(ref + (ptrdiff_t)(r - 3) * ref_pitch + (cs - 3), px);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        int sum = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) sum += px[c + t] * th[t];
        hh[r][c] = (int16_t)((sum - hoff) >> hshift);                    // isFirst: offset -8192 << shift
      }
      if (r >= 7) {
        const int y = r - 7;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          int sum = 0;
#pragma unroll
          for (int t = 0; t < 8; t++) sum += hh[y + t][c] * tv[t];
          const int i = y * N + cs + c;
          if (MODE == 0) {
            const int val = min(max((int)(int16_t)((sum + voff) >> vshift), 0), maxv);
            out[i] = (int)org[(ptrdiff_t)y * org_pitch + cs + c] - val;
          } else {
            const int inter = (int)(int16_t)(sum >> 6);                 // isFirst = isLast = false: shift 6, no offset
            if (MODE == 1) out[i] = inter;
            else {
              const int val = min(max((out[i] + inter + aoff) >> ashift, 0), maxv);
              out[i] = (int)org[(ptrdiff_t)y * org_pitch + cs + c] - val;
            }
          }
        }
      }
    }
  }
}

// tiles: candidate (24 bits) | tile column << 24 | tile row << 28 (frac_pack_tile)
template <typename T, int N, bool HAD>
__global__ void __launch_bounds__(128)
k_mc_cand_tiles(const McCandDev* __restrict__ cands, const uint32_t* __restrict__ tiles, int n_tiles, uint32_t* __restrict__ dist, int bit_depth) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tiles) return;
  const uint32_t tile = tiles[i];
  const uint32_t ci = tile & 0xffffffu;
  const int tx = (tile >> 24) & 15, ty = tile >> 28;
  const McCandDev c = cands[ci];
  const int head = max(2, 14 - bit_depth), maxv = (1 << bit_depth) - 1;
  const T* org = reinterpret_cast<const T*>(c.org) + (ptrdiff_t)(ty * N) * c.org_pitch + tx * N;
  auto at = [&](const void* base, int pitch, int mvx, int mvy) {
    return reinterpret_cast<const T*>(base) + (ptrdiff_t)(ty * N + (mvy >> 2)) * pitch + tx * N + (mvx >> 2);
  };
  int d[N * N];
  if (c.ref0 && c.ref1) {
    mc_tile_pass<T, N, 1>(at(c.ref0, c.ref0_pitch, c.mv0_x, c.mv0_y), c.ref0_pitch, org, c.org_pitch, c.mv0_x & 3, c.mv0_y & 3, head, maxv, d);
    mc_tile_pass<T, N, 2>(at(c.ref1, c.ref1_pitch, c.mv1_x, c.mv1_y), c.ref1_pitch, org, c.org_pitch, c.mv1_x & 3, c.mv1_y & 3, head, maxv, d);
  } else if (c.ref0) {
    mc_tile_pass<T, N, 0>(at(c.ref0, c.ref0_pitch, c.mv0_x, c.mv0_y), c.ref0_pitch, org, c.org_pitch, c.mv0_x & 3, c.mv0_y & 3, head, maxv, d);
  } else {
    mc_tile_pass<T, N, 0>(at(c.ref1, c.ref1_pitch, c.mv1_x, c.mv1_y), c.ref1_pitch, org, c.org_pitch, c.mv1_x & 3, c.mv1_y & 3, head, maxv, d);
  }
  uint32_t s;
  if (HAD) {
    if constexpr (N == 8) s = (had8x8_abs(d) + 2) >> 2;                  // TComRdCost.cpp:1520
    else                  s = (had4x4_abs(d) + 1) >> 1;                  // TComRdCost.cpp:1423
  } else {
    s = 0;
#pragma unroll
    for (int k = 0; k < N * N; k++) s += (uint32_t)abs(d[k]);
  }
  atomicAdd(&dist[ci], s);
}

__global__ void k_mc_cand_finish(uint32_t* __restrict__ dist, int n, int bit_depth) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dist[i] >>= (bit_depth - 8);                               // distortion precision adjustment (TComRdCost.cpp:517, 1592)
}

}  // namespace hmb200
