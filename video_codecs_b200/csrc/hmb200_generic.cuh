// Generic (any sample type, any PU size, any window) kernels of the path.  They are the always-correct baseline:
// 16-bit content, bi-pred style int16 patterns and the 1:1 per-call entries run here; 8-bit batched searches run
// in the tiled byte-SIMD kernel (hmb200_search8.cuh) and use these only for shapes it does not cover.
#pragma once
#include "hmb200_device.cuh"

namespace hmb200 {

// A PU search as the kernels see it: pointers already resolved by the host frontend.
struct SearchTask {
  int32_t org_x, org_y;  // PU top-left in the current plane (or in an uploaded pattern buffer)
  int32_t ref_x, ref_y;  // co-located sample in the padded reference plane
  int32_t w, h;
  int32_t lt_x, lt_y, rb_x, rb_y;
  int32_t pred_x, pred_y;
  uint32_t lambda_cost;
  int32_t sub_shift;     // 1 when FEN && h > 8 (TEncSearch.cpp:3804-3810)
};

struct DistTask {
  const void* org; const void* cur;
  int32_t org_stride, cur_stride, w, h, sub_shift, reserved;
};

// ---------------------------------------------------------------------------------------------------------------
// integer full search, one CTA per PU, one candidate per thread-iteration.
// TLibEncoder/TEncSearch.cpp:3786-3843 (xPatternSearch) + TComRdCost.cpp:489-953 (xGetSADnn with iSubShift)
// ---------------------------------------------------------------------------------------------------------------
template <typename RefT, typename OrgT>
__global__ void __launch_bounds__(256) k_search_generic(const SearchTask* __restrict__ tasks, hmb200_pu_result* __restrict__ out,
                                                        DevPlane cur_plane, DevPlane ref_plane, const int* __restrict__ index) {
  __shared__ int16_t s_org[64 * 64];
  __shared__ unsigned long long s_best[8];
  const int ti = index ? index[blockIdx.x] : (int)blockIdx.x;     // optional indirection: a subset of the task list
  const SearchTask t = tasks[ti];
  const OrgT* org = plane_at<OrgT>(cur_plane, t.org_x, t.org_y);
  const RefT* ref = plane_at<RefT>(ref_plane, t.ref_x, t.ref_y);
  const int org_stride = cur_plane.pitch, ref_stride = ref_plane.pitch, bit_depth = ref_plane.bit_depth;
  const int step = 1 << t.sub_shift;
  const int rows = t.h >> t.sub_shift;                  // rows actually visited: 0, step, 2*step, ...
  for (int i = threadIdx.x; i < rows * t.w; i += blockDim.x) {
    int r = i / t.w, c = i - r * t.w;
    s_org[r * t.w + c] = (int16_t)org[(size_t)(r * step) * org_stride + c];
  }
  __syncthreads();
  const int nx = t.rb_x - t.lt_x + 1, ny = t.rb_y - t.lt_y + 1;
  unsigned long long best = ~0ull;
  for (int idx = threadIdx.x; idx < nx * ny; idx += blockDim.x) {
    int cy = idx / nx, cx = idx - cy * nx;
    int x = t.lt_x + cx, y = t.lt_y + cy;
    const RefT* p = ref + (ptrdiff_t)y * ref_stride + x;
    uint32_t sum = 0;
    for (int r = 0; r < rows; r++) {
      const RefT* q = p + (ptrdiff_t)(r * step) * ref_stride;
      for (int c = 0; c < t.w; c++) sum += (uint32_t)abs((int)s_org[r * t.w + c] - (int)q[c]);
    }
    sum = (sum << t.sub_shift) >> (bit_depth - 8);
    uint32_t cost = sum + mv_cost(t.lambda_cost, mv_bits(x, y, t.pred_x, t.pred_y, 2));
    unsigned long long key = make_key(cost, (uint32_t)idx);
    best = key < best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) best = s_best[w] < best ? s_best[w] : best;
    uint32_t idx = (uint32_t)(best & 0xffffffffu), cost = (uint32_t)(best >> 32);
    int cy = idx / nx, cx = idx - cy * nx;
    int x = t.lt_x + cx, y = t.lt_y + cy;
    hmb200_pu_result r = out[ti];
    r.mv_x = x; r.mv_y = y;
    r.sad = cost - mv_cost(t.lambda_cost, mv_bits(x, y, t.pred_x, t.pred_y, 2));   // TEncSearch.cpp:3841
    out[ti] = r;
  }
}

// Same search with the candidates of one PU split over gridDim.y CTAs (the 1:1 entries: one PU per call, the whole
// GPU is idle otherwise).  Each CTA folds its slice into keys[task] with atomicMin; k_search8_finalize decodes the key.
template <typename RefT, typename OrgT>
__global__ void __launch_bounds__(256) k_search_split(const SearchTask* __restrict__ tasks, unsigned long long* __restrict__ keys,
                                                      DevPlane cur_plane, DevPlane ref_plane) {
  __shared__ int16_t s_org[64 * 64];
  __shared__ unsigned long long s_best[8];
  const SearchTask t = tasks[blockIdx.x];
  const OrgT* org = plane_at<OrgT>(cur_plane, t.org_x, t.org_y);
  const RefT* ref = plane_at<RefT>(ref_plane, t.ref_x, t.ref_y);
  const int org_stride = cur_plane.pitch, ref_stride = ref_plane.pitch, bit_depth = ref_plane.bit_depth;
  const int step = 1 << t.sub_shift, rows = t.h >> t.sub_shift;
  for (int i = threadIdx.x; i < rows * t.w; i += blockDim.x) {
    int r = i / t.w, c = i - r * t.w;
    s_org[r * t.w + c] = (int16_t)org[(size_t)(r * step) * org_stride + c];
  }
  __syncthreads();
  const int nx = t.rb_x - t.lt_x + 1, ny = t.rb_y - t.lt_y + 1, total = nx * ny;
  const int chunk = (total + gridDim.y - 1) / gridDim.y;
  const int first = blockIdx.y * chunk, last = min(total, first + chunk);
  unsigned long long best = ~0ull;
  for (int idx = first + threadIdx.x; idx < last; idx += blockDim.x) {
    int cy = idx / nx, cx = idx - cy * nx;
    int x = t.lt_x + cx, y = t.lt_y + cy;
    const RefT* p = ref + (ptrdiff_t)y * ref_stride + x;
    uint32_t sum = 0;
    for (int r = 0; r < rows; r++) {
      const RefT* q = p + (ptrdiff_t)(r * step) * ref_stride;
      for (int c = 0; c < t.w; c++) sum += (uint32_t)abs((int)s_org[r * t.w + c] - (int)q[c]);
    }
    sum = (sum << t.sub_shift) >> (bit_depth - 8);
    const unsigned long long key = make_key(sum + mv_cost(t.lambda_cost, mv_bits(x, y, t.pred_x, t.pred_y, 2)), (uint32_t)idx);
    best = key < best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) best = s_best[w] < best ? s_best[w] : best;
    if (best != ~0ull) atomicMin(&keys[blockIdx.x], best);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// quarter-pel refinement, one CTA per PU.
// TLibEncoder/TEncSearch.cpp:4240-4276 (xPatternSearchFracDIF), :808-861 (xPatternRefinement),
// :5338-5539 (xExtDIFUpSamplingH/Q) — each pre-filtered plane + offset of the reference is the block interpolated
// at quarter-pel displacement 2*half (+ qter) from the integer MV, built here per candidate from shared
// horizontally-filtered columns.  Distortion: xGetHADs tiles (TComRdCost.cpp:1526-1593) or SAD.
// ---------------------------------------------------------------------------------------------------------------
constexpr int FRAC_THREADS = 128;
constexpr int FRAC_REF_W = 64 + 8, FRAC_REF_H = 64 + 8;
// dynamic shared memory: ref (72*72 int16) + 3 hor planes (3*72*64 int16) + org (64*64 int16)
constexpr int FRAC_SMEM_BYTES = (FRAC_REF_W * FRAC_REF_H + 3 * FRAC_REF_H * 64 + 64 * 64) * 2;

template <typename RefT, typename OrgT>
__global__ void __launch_bounds__(FRAC_THREADS) k_frac_generic(const SearchTask* __restrict__ tasks,
                                                               hmb200_pu_result* __restrict__ out,
                                                               DevPlane cur_plane, DevPlane ref_plane, int use_had) {
  extern __shared__ __align__(16) int16_t smem16[];
  int16_t* s_ref = smem16;                                  // [H+8][W+8], origin at (-4,-4) of the MC block
  int16_t* s_hor = s_ref + FRAC_REF_W * FRAC_REF_H;         // [3][H+8][W]
  int16_t* s_org = s_hor + 3 * FRAC_REF_H * 64;             // [H][W]
  __shared__ uint32_t s_dist[9];
  __shared__ int s_sel[2];

  const SearchTask t = tasks[blockIdx.x];
  const OrgT* org = plane_at<OrgT>(cur_plane, t.org_x, t.org_y);
  const RefT* ref = plane_at<RefT>(ref_plane, t.ref_x, t.ref_y);
  const int org_stride = cur_plane.pitch, ref_stride = ref_plane.pitch, bit_depth = ref_plane.bit_depth;
  const int W = t.w, H = t.h, RW = W + 8, RH = H + 8;
  hmb200_pu_result res = out[blockIdx.x];
  const int mvx = res.mv_x, mvy = res.mv_y;
  const int head = max(2, 14 - bit_depth), maxv = (1 << bit_depth) - 1;

  for (int i = threadIdx.x; i < RW * RH; i += FRAC_THREADS) {
    int r = i / RW, c = i - r * RW;
    s_ref[r * RW + c] = (int16_t)ref[(ptrdiff_t)(mvy + r - 4) * ref_stride + (mvx + c - 4)];
  }
  for (int i = threadIdx.x; i < W * H; i += FRAC_THREADS) {
    int r = i / W, c = i - r * W;
    s_org[r * W + c] = (int16_t)org[(size_t)r * org_stride + c];
  }
  const int n = (!use_had) ? 4 : ((W % 8 == 0 && H % 8 == 0) ? 8 : 4);   // tile edge
  const int tiles_x = W / n, tiles = tiles_x * (H / n);

  int base_qx = 0, base_qy = 0;            // 2*half after stage 1
  for (int stage = 0; stage < 2; stage++) {
    const int stepq = (stage == 0) ? 2 : 1;
    if (threadIdx.x < 9) s_dist[threadIdx.x] = 0;
    __syncthreads();
    // horizontal pass for the three distinct quarter-pel x displacements of this stage
    for (int i = threadIdx.x; i < 3 * RH * W; i += FRAC_THREADS) {
      int k = i / (RH * W), rem = i - k * RH * W, r = rem / W, c = rem - r * W;
      int qx = base_qx + (k - 1) * stepq;
      int ix = floor_div4(qx), fx = qx & 3;
      int s[8];
#pragma unroll
      for (int tp = 0; tp < 8; tp++) s[tp] = s_ref[r * RW + (c + 4 + ix + tp - 3)];
      s_hor[(k * RH + r) * W + c] = interp_h(s, fx, head);
    }
    __syncthreads();
    // vertical pass + distortion, one (candidate, tile) per thread-iteration
    for (int task = threadIdx.x; task < 9 * tiles; task += FRAC_THREADS) {
      int cand = task / tiles, tile = task - cand * tiles;
      int ty = tile / tiles_x, tx = tile - ty * tiles_x;
      int cx = (stage == 0) ? k_refine_h[cand][0] : k_refine_q[cand][0];
      int cy = (stage == 0) ? k_refine_h[cand][1] : k_refine_q[cand][1];
      int qy = base_qy + cy * stepq;
      int iy = floor_div4(qy), fy = qy & 3;
      const int16_t* hp = s_hor + ((cx + 1) * RH) * W;
      uint32_t d = 0;
      if (n == 8) {
        int diff[64];
#pragma unroll
        for (int c = 0; c < 8; c++) {
          int col[15];
#pragma unroll
          for (int r = 0; r < 15; r++) col[r] = hp[(ty * 8 + 4 + iy + r - 3) * W + tx * 8 + c];
#pragma unroll
          for (int r = 0; r < 8; r++) {
            int cc[8];
#pragma unroll
            for (int tp = 0; tp < 8; tp++) cc[tp] = col[r + tp];
            diff[r * 8 + c] = (int)s_org[(ty * 8 + r) * W + tx * 8 + c] - interp_v(cc, fy, head, maxv);
          }
        }
        d = (had8x8_abs(diff) + 2) >> 2;                                   // TComRdCost.cpp:1520
      } else {
        int diff[16];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          int col[11];
#pragma unroll
          for (int r = 0; r < 11; r++) col[r] = hp[(ty * 4 + 4 + iy + r - 3) * W + tx * 4 + c];
#pragma unroll
          for (int r = 0; r < 4; r++) {
            int cc[8];
#pragma unroll
            for (int tp = 0; tp < 8; tp++) cc[tp] = col[r + tp];
            diff[r * 4 + c] = (int)s_org[(ty * 4 + r) * W + tx * 4 + c] - interp_v(cc, fy, head, maxv);
          }
        }
        if (use_had) d = (had4x4_abs(diff) + 1) >> 1;                      // TComRdCost.cpp:1423
        else {
#pragma unroll
          for (int i = 0; i < 16; i++) d += (uint32_t)abs(diff[i]);
        }
      }
      atomicAdd(&s_dist[cand], d);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t best = 0xffffffffu; int bi = 0;
      for (int i = 0; i < 9; i++) {
        int cx = (stage == 0) ? k_refine_h[i][0] : k_refine_q[i][0];
        int cy = (stage == 0) ? k_refine_h[i][1] : k_refine_q[i][1];
        uint32_t bits = (stage == 0)
            ? mv_bits(cx + 2 * mvx, cy + 2 * mvy, t.pred_x, t.pred_y, 1)                                   // :3746
            : mv_bits(cx + 2 * (2 * mvx) + base_qx, cy + 2 * (2 * mvy) + base_qy, t.pred_x, t.pred_y, 0);  // :4267
        uint32_t c = (s_dist[i] >> (bit_depth - 8)) + mv_cost(t.lambda_cost, bits);
        if (c < best) { best = c; bi = i; }
      }
      s_sel[0] = bi; s_sel[1] = (int)best;
    }
    __syncthreads();
    int bi = s_sel[0];
    if (stage == 0) {
      res.half_x = k_refine_h[bi][0]; res.half_y = k_refine_h[bi][1];
      base_qx = 2 * res.half_x; base_qy = 2 * res.half_y;
    } else {
      res.qter_x = k_refine_q[bi][0]; res.qter_y = k_refine_q[bi][1];
      res.frac_cost = (uint32_t)s_sel[1];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = res;
}

// ---------------------------------------------------------------------------------------------------------------
// distortion table entries, one warp per evaluation.
// TComRdCost.cpp:489-953 (SAD), :959-1304 (SSE), :1526-1593 (HADs)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) k_dist_generic(const DistTask* __restrict__ tasks, uint32_t* __restrict__ out,
                                                      int n_tasks, int func, int bit_depth) {
  int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= n_tasks) return;
  const DistTask t = tasks[wid];
  const T* org = reinterpret_cast<const T*>(t.org);
  const T* cur = reinterpret_cast<const T*>(t.cur);
  uint32_t sum = 0;
  if (func == HMB200_DF_SAD || func == HMB200_DF_SADS) {
    const bool sized = (t.w == 4 || t.w == 8 || t.w == 12 || t.w == 16 || t.w == 24 || t.w == 32 || t.w == 48 || t.w == 64);
    const int ss = sized ? t.sub_shift : 0;             // the generic xGetSAD ignores iSubShift (TComRdCost.cpp:461-487)
    const int rows = (t.h + (1 << ss) - 1) >> ss;
    for (int i = lane; i < rows * t.w; i += 32) {
      int r = (i / t.w) << ss, c = i % t.w;
      sum += (uint32_t)abs((int)org[(size_t)r * t.org_stride + c] - (int)cur[(size_t)r * t.cur_stride + c]);
    }
    sum <<= ss;
  } else if (func == HMB200_DF_SSE) {
    const int sh = (bit_depth - 8) << 1;
    for (int i = lane; i < t.h * t.w; i += 32) {
      int r = i / t.w, c = i % t.w;
      int d = (int)org[(size_t)r * t.org_stride + c] - (int)cur[(size_t)r * t.cur_stride + c];
      sum += (uint32_t)((d * d) >> sh);
    }
  } else {
    const int n = (t.w % 8 == 0 && t.h % 8 == 0) ? 8 : ((t.w % 4 == 0 && t.h % 4 == 0) ? 4 : 2);
    const int tiles_x = t.w / n, tiles = tiles_x * (t.h / n);
    for (int tile = lane; tile < tiles; tile += 32) {
      int ty = tile / tiles_x, tx = tile - ty * tiles_x;
      const T* o = org + (size_t)(ty * n) * t.org_stride + tx * n;
      const T* c = cur + (size_t)(ty * n) * t.cur_stride + tx * n;
      if (n == 8) {
        int d[64];
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int k = 0; k < 8; k++) d[r * 8 + k] = (int)o[(size_t)r * t.org_stride + k] - (int)c[(size_t)r * t.cur_stride + k];
        sum += (had8x8_abs(d) + 2) >> 2;
      } else if (n == 4) {
        int d[16];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
          for (int k = 0; k < 4; k++) d[r * 4 + k] = (int)o[(size_t)r * t.org_stride + k] - (int)c[(size_t)r * t.cur_stride + k];
        sum += (had4x4_abs(d) + 1) >> 1;
      } else {
        sum += had2x2_abs((int)o[0] - (int)c[0], (int)o[1] - (int)c[1],
                          (int)o[t.org_stride] - (int)c[t.cur_stride], (int)o[t.org_stride + 1] - (int)c[t.cur_stride + 1]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) out[wid] = (func == HMB200_DF_SSE) ? sum : (sum >> (bit_depth - 8));
}

// ---------------------------------------------------------------------------------------------------------------
// plane ingest: narrow Pel (int16) rows to the device sample type, or synthesise the margins like
// TComPicYuv::extendPicBorder (TLibCommon/TComPicYuv.cpp:197-242).
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_narrow_plane(const int16_t* __restrict__ src, int src_stride, T* __restrict__ dst, int dst_pitch,
                               int total_w, int total_h) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < total_w && y < total_h) dst[(size_t)y * dst_pitch + x] = (T)src[(size_t)y * src_stride + x];
}
template <typename T>
__global__ void k_widen_plane(const T* __restrict__ src, int src_pitch, int16_t* __restrict__ dst, int dst_stride,
                              int total_w, int total_h) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < total_w && y < total_h) dst[(size_t)y * dst_stride + x] = (int16_t)src[(size_t)y * src_pitch + x];
}
// src: tightly packed w x h samples (u8 or u16); dst: padded plane; every destination sample (margins included) is the
// clamped-coordinate source sample — identical to replicating edges.
template <typename T>
__global__ void k_pad_plane(const T* __restrict__ src, int src_stride, T* __restrict__ dst, int dst_pitch,
                            int w, int h, int mx, int my) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < w + 2 * mx && y < h + 2 * my) {
    int sx = min(max(x - mx, 0), w - 1), sy = min(max(y - my, 0), h - 1);
    dst[(size_t)y * dst_pitch + x] = src[(size_t)sy * src_stride + sx];
  }
}

// Planar-file luma ingest: TVideoIOYuv::read for the luma component (TLibVideoIO/TVideoIOYuv.cpp:247-377 readPlane,
// :70-99 scalePlane) fused with TComPicYuv::extendPicBorder (TLibCommon/TComPicYuv.cpp:197-242).  src: width x height
// file samples (bytes, or 16-bit little endian); the coded picture is (width + pad_x) x (height + pad_y) with the last
// column / row replicated; every sample is scaled by 2^shift; margins replicate the coded picture's edges.
template <typename T>
__global__ void k_ingest_luma(const uint8_t* __restrict__ src, int is16, int width, int height, int pad_x, int pad_y, int shift,
                              int maxval, T* __restrict__ dst, int dst_pitch, int mx, int my) {
  const int cw = width + pad_x, chh = height + pad_y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= cw + 2 * mx || y >= chh + 2 * my) return;
  const int sx = min(min(max(x - mx, 0), cw - 1), width - 1), sy = min(min(max(y - my, 0), chh - 1), height - 1);
  const size_t o = (size_t)sy * width + sx;
  int v = is16 ? (int)(int16_t)((uint16_t)src[2 * o] | ((uint16_t)src[2 * o + 1] << 8)) : (int)src[o];
  if (shift > 0) v = (int)(int16_t)(v << shift);
  else if (shift < 0) { v = (int)(int16_t)((int16_t)(v + (1 << (-shift - 1))) >> (-shift)); v = min(max(v, 0), maxval); }
  dst[(size_t)y * dst_pitch + x] = (T)v;
}

} // namespace hmb200
