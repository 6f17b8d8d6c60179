// One PU per call: the kernels behind the 1:1 entries (hmb200_pattern_search, hmb200_pattern_search_frac,
// hmb200_pattern_search_and_refine, hmb200_pattern_search_tz*), which the in-encoder forwarders call once per
// TEncSearch::xPatternSearch / xPatternSearchFracDIF (TLibEncoder/TEncSearch.cpp:3796-3841, :4240-4276).
//
// A call is latency: the host waits for this PU's answer before HM can form the next PU's predictor.  Measured on B200
// (tools/latency_1to1.py, profiles/r02_latency_1to1.txt) the floor of "copy in, K kernels, copy out, stream sync" is
// 16 us + 2.7 us per extra kernel; "copy in, K kernels, result + sequence flag written straight into mapped host memory,
// host spins on the flag" is 9 us + 2.7 us per extra kernel.  So:
//   * k_one_search  one CTA per candidate ROW of the window (129 CTAs at +-64, one wave); the CTA stages the reference
//                   rows its candidates touch (aligned 16-byte loads) and the pattern in shared memory, a thread per
//                   candidate runs byte-SIMD SADs (VABSDIFF4 on funnel-shifted words) when pattern and plane are 8-bit
//                   (scalar 16-bit otherwise: 10-bit planes, the signed 2*org - pred pattern of bi-prediction);
//                   the LAST CTA to finish (ticket counter) decodes the 64-bit argmin key - no finalize launch - and, for
//                   PUs up to 16x16 (nine calls in ten inside the encoder), refines the vector itself: one launch per call;
//   * k_one_frac    larger PUs / refinement-only calls: one CTA, up to 1024 threads, a thread per COLUMN of a (candidate,
//                   Hadamard tile) instead of a thread per tile (an 8x8 PU has 9 tiles per stage - 9 busy threads); the row
//                   transform runs across the 4 / 8 lanes with shuffles.  The sum of absolute coefficients does not depend
//                   on the butterfly order, the per-tile rounding is kept;
//   * OneBack       whichever kernel ends the call stores the result into mapped page-locked host memory as two 16-byte
//                   records that carry the call's sequence number in their last word; the host spins on the number;
//   * the task and the caller's result record travel as kernel arguments, and so does the pattern of a PU up to 16x16
//     (OnePattern, 512 bytes): such a call is ONE launch and no copy in either direction.
// Arithmetic, candidate order and tie-breaks are those of k_search_split / k_frac_generic (hmb200_generic.cuh), which stay
// as the fallback for windows that do not fit shared memory and as the A/B reference of tests/test_gpu_parity.py.
#pragma once
#include "hmb200_generic.cuh"

namespace hmb200 {

// Where the call's last kernel reports to: two 16-byte records in mapped page-locked host memory, each written by ONE vector store
// whose last word is the call's sequence number - {mv_x, mv_y, sad, seq} and {half | qter (4 x int8), frac_cost, 0, seq}.  A 16-byte
// store crosses PCIe as one write in address order, so a record whose last word shows the number is complete: no system-scope fence
// (2.3 us of a 10.9 us round trip, tools/roundtrip_report_probe.cu).  host_a == nullptr: report nowhere.
struct OneBack {
  uint4* host_a;
  uint4* host_b;
  uint32_t seq;
};

__device__ __forceinline__ void one_report(const OneBack& back, const hmb200_pu_result& r) {
  if (!back.host_a) return;
  const uint32_t frac = ((uint32_t)r.half_x & 0xffu) | (((uint32_t)r.half_y & 0xffu) << 8) | (((uint32_t)r.qter_x & 0xffu) << 16) |
                        (((uint32_t)r.qter_y & 0xffu) << 24);
  asm volatile("st.volatile.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(back.host_a), "r"((uint32_t)r.mv_x), "r"((uint32_t)r.mv_y), "r"(r.sad), "r"(back.seq) : "memory");
  asm volatile("st.volatile.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(back.host_b), "r"(frac), "r"(r.frac_cost), "r"(0u), "r"(back.seq) : "memory");
}

// a PU of up to 16x16 samples travels as a kernel argument (no H2D copy in the call at all)
constexpr int ONE_ARG_SAMPLES = 256;
struct OnePattern { int16_t px[ONE_ARG_SAMPLES]; };

// ---------------------------------------------------------------------------------------------------------------
// quarter-pel refinement of one PU (xPatternSearchFracDIF): k_frac_generic's stages with the vertical pass + distortion
// spread over N threads per (candidate, tile).
// ---------------------------------------------------------------------------------------------------------------
constexpr int ONE_FRAC_THREADS_MAX = 1024;

// column c of an N x N difference tile (d[r], r = 0..N-1) -> this lane's share of the sum of absolute Hadamard coefficients
template <int N>
__device__ __forceinline__ uint32_t had_columns_abs(int (&d)[N], int lane_in_group) {
  // column transform in registers
#pragma unroll
  for (int len = 1; len < N; len <<= 1)
#pragma unroll
    for (int b = 0; b < N; b += 2 * len)
#pragma unroll
      for (int k = 0; k < len; k++) { const int a0 = d[b + k], a1 = d[b + k + len]; d[b + k] = a0 + a1; d[b + k + len] = a0 - a1; }
  // row transform across the N lanes that hold the tile's columns
#pragma unroll
  for (int m = 1; m < N; m <<= 1) {
    const bool upper = (lane_in_group & m) != 0;
#pragma unroll
    for (int r = 0; r < N; r++) {
      const int other = __shfl_xor_sync(0xffffffffu, d[r], m);
      d[r] = upper ? other - d[r] : d[r] + other;
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int r = 0; r < N; r++) s += (uint32_t)abs(d[r]);
  return s;
}

template <int N>
__device__ __forceinline__ void one_frac_tiles(const int16_t* __restrict__ s_hor, const int16_t* __restrict__ s_org, uint32_t* s_dist,
                                               int W, int H, int RH, int stage, int base_qy, int stepq, int use_had, int head, int maxv) {
  const int tiles_x = W / N, tiles = tiles_x * (H / N), total = 9 * tiles * N;
  for (int base = 0; base < total; base += blockDim.x) {
    const int tt = base + (int)threadIdx.x;
    const bool live = tt < total;                                           // total and blockDim are multiples of N: a lane group is live as a whole
    const int g = (live ? tt : total - 1) / N, c = threadIdx.x & (N - 1);
    const int cand = g / tiles, tile = g - cand * tiles;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int cx = (stage == 0) ? k_refine_h[cand][0] : k_refine_q[cand][0];
    const int cy = (stage == 0) ? k_refine_h[cand][1] : k_refine_q[cand][1];
    const int qy = base_qy + cy * stepq;
    const int iy = floor_div4(qy), fy = qy & 3;
    const int16_t* hp = s_hor + ((cx + 1) * RH) * W;
    int col[N + 7];
#pragma unroll
    for (int r = 0; r < N + 7; r++) col[r] = hp[(ty * N + 4 + iy + r - 3) * W + tx * N + c];
    int d[N];
#pragma unroll
    for (int r = 0; r < N; r++) {
      int cc[8];
#pragma unroll
      for (int tp = 0; tp < 8; tp++) cc[tp] = col[r + tp];
      d[r] = (int)s_org[(ty * N + r) * W + tx * N + c] - interp_v(cc, fy, head, maxv);
    }
    uint32_t s;
    if (use_had) s = had_columns_abs<N>(d, c);
    else {
      s = 0;
#pragma unroll
      for (int r = 0; r < N; r++) s += (uint32_t)abs(d[r]);
    }
#pragma unroll
    for (int m = 1; m < N; m <<= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (live && c == 0) {
      if (use_had) s = (N == 8) ? (s + 2) >> 2 : (s + 1) >> 1;             // TComRdCost.cpp:1520, :1423
      atomicAdd(&s_dist[cand], s);
    }
  }
}

// dynamic shared memory (bytes) of one_frac_body for a PU: reference block with its 4-sample apron, three horizontally filtered
// planes, the pattern
__host__ __device__ inline int one_frac_smem(int w, int h) { return ((w + 8) * (h + 8) + 3 * (h + 8) * w + w * h) * 2; }

// xPatternSearchFracDIF of the PU at res.mv_x / mv_y, by every thread of the calling CTA (blockDim.x a multiple of 32); fills
// res.half_*, res.qter_*, res.frac_cost in every thread.  pattern: int16 rows, pattern_pitch samples apart (0: dense, pitch = w).
template <typename RefT>
__device__ __forceinline__ void one_frac_body(const SearchTask& t, hmb200_pu_result& res, const int16_t* __restrict__ pattern,
                                              const DevPlane& ref_plane, int use_had, int16_t* smem16, int pattern_pitch = 0) {
  const int W = t.w, H = t.h, RW = W + 8, RH = H + 8;
  int16_t* s_ref = smem16;                                  // [H+8][W+8], origin at (-4,-4) of the MC block
  int16_t* s_hor = s_ref + RW * RH;                         // [3][H+8][W]
  int16_t* s_org = s_hor + 3 * RH * W;                      // [H][W]
  __shared__ uint32_t s_dist[9];
  __shared__ int s_sel[2];

  const RefT* ref = plane_at<RefT>(ref_plane, t.ref_x, t.ref_y);
  const int ref_stride = ref_plane.pitch, bit_depth = ref_plane.bit_depth;
  const int mvx = res.mv_x, mvy = res.mv_y;
  const int head = max(2, 14 - bit_depth), maxv = (1 << bit_depth) - 1;

  for (int i = threadIdx.x; i < RW * RH; i += blockDim.x) {
    const int r = i / RW, c = i - r * RW;
    s_ref[i] = (int16_t)ref[(ptrdiff_t)(mvy + r - 4) * ref_stride + (mvx + c - 4)];
  }
  if (pattern_pitch == 0) pattern_pitch = W;              // dense rows unless the PU is a part of a larger block (one_cu_*)
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) { const int r = i / W; s_org[i] = pattern[r * pattern_pitch + (i - r * W)]; }
  const int n = (!use_had) ? 4 : ((W % 8 == 0 && H % 8 == 0) ? 8 : 4);   // tile edge

  int base_qx = 0, base_qy = 0;            // 2*half after stage 1
  for (int stage = 0; stage < 2; stage++) {
    const int stepq = (stage == 0) ? 2 : 1;
    if (threadIdx.x < 9) s_dist[threadIdx.x] = 0;
    __syncthreads();
    // horizontal pass for the three distinct quarter-pel x displacements of this stage
    for (int i = threadIdx.x; i < 3 * RH * W; i += blockDim.x) {
      const int k = i / (RH * W), rem = i - k * RH * W, r = rem / W, c = rem - r * W;
      const int qx = base_qx + (k - 1) * stepq;
      const int ix = floor_div4(qx), fx = qx & 3;
      int s[8];
#pragma unroll
      for (int tp = 0; tp < 8; tp++) s[tp] = s_ref[r * RW + (c + 4 + ix + tp - 3)];
      s_hor[i] = interp_h(s, fx, head);
    }
    __syncthreads();
    if (n == 8) one_frac_tiles<8>(s_hor, s_org, s_dist, W, H, RH, stage, base_qy, stepq, use_had, head, maxv);
    else        one_frac_tiles<4>(s_hor, s_org, s_dist, W, H, RH, stage, base_qy, stepq, use_had, head, maxv);
    __syncthreads();
    if (threadIdx.x < 32) {
      // cost of the nine candidates, one per lane; the first minimum wins (strict '<' in the reference's candidate order)
      const int i = min((int)threadIdx.x, 8);
      const int cx = (stage == 0) ? k_refine_h[i][0] : k_refine_q[i][0];
      const int cy = (stage == 0) ? k_refine_h[i][1] : k_refine_q[i][1];
      const uint32_t bits = (stage == 0)
          ? mv_bits(cx + 2 * mvx, cy + 2 * mvy, t.pred_x, t.pred_y, 1)                                   // :3746
          : mv_bits(cx + 2 * (2 * mvx) + base_qx, cy + 2 * (2 * mvy) + base_qy, t.pred_x, t.pred_y, 0);  // :4267
      const uint32_t c = (s_dist[i] >> (bit_depth - 8)) + mv_cost(t.lambda_cost, bits);
      unsigned long long k = threadIdx.x < 9 ? make_key(c, (uint32_t)i) : ~0ull;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
        k = other < k ? other : k;
      }
      if (threadIdx.x == 0) { s_sel[0] = (int)(k & 0xffffffffu); s_sel[1] = (int)(uint32_t)(k >> 32); }
    }
    __syncthreads();
    const int bi = s_sel[0];
    if (stage == 0) {
      res.half_x = k_refine_h[bi][0]; res.half_y = k_refine_h[bi][1];
      base_qx = 2 * res.half_x; base_qy = 2 * res.half_y;
    } else {
      res.qter_x = k_refine_q[bi][0]; res.qter_y = k_refine_q[bi][1];
      res.frac_cost = (uint32_t)s_sel[1];
    }
    __syncthreads();
  }
}

// seed: the caller's result record; mv_from_device: take the integer MV from out[0] (a search kernel ran before) instead
template <typename RefT>
__global__ void __launch_bounds__(ONE_FRAC_THREADS_MAX)
k_one_frac(const SearchTask t, hmb200_pu_result seed, int mv_from_device, hmb200_pu_result* __restrict__ out, const int16_t* __restrict__ pattern,
           DevPlane ref_plane, int use_had, OneBack back) {
  extern __shared__ __align__(16) int16_t one_smem16[];
  hmb200_pu_result res = seed;
  if (mv_from_device) res = out[0];
  one_frac_body<RefT>(t, res, pattern, ref_plane, use_had, one_smem16);
  if (threadIdx.x == 0) { out[0] = res; one_report(back, res); }
}
// refinement-only call of a small PU: everything in the arguments
template <typename RefT>
__global__ void __launch_bounds__(ONE_FRAC_THREADS_MAX)
k_one_frac_args(const SearchTask t, hmb200_pu_result seed, const __grid_constant__ OnePattern pat, DevPlane ref_plane, int use_had, OneBack back) {
  extern __shared__ __align__(16) int16_t one_smem16[];
  hmb200_pu_result res = seed;
  one_frac_body<RefT>(t, res, pat.px, ref_plane, use_had, one_smem16);
  if (threadIdx.x == 0) one_report(back, res);
}

constexpr int ONE_SEARCH_THREADS = 160;          // 129 candidates of a +-64 window row: five warps
constexpr int ONE_SEARCH_THREADS_MAX = 512;      // ... more when the last CTA also refines (a thread per tile column of a stage)
constexpr int ONE_SMEM_MAX = 200 * 1024;
constexpr int ONE_FUSE_FRAC = 1, ONE_FUSE_HAD = 2;  // k_one_search's fuse argument
constexpr int ONE_FUSE_MAX_SAMPLES = 256;           // PUs up to 16x16 are refined by the search kernel's last CTA

// dynamic shared memory of k_one_search for a task (host and device agree on the layout)
__host__ __device__ inline int one_search_row_bytes(bool bytes, int col0, int nx, int w) {
  if (bytes) return (((col0 & 15) + nx + w - 1 + 15) / 16 + 1) * 16;      // 16-byte vectors covering the row, one more for the word past it
  return ((nx + w - 1 + 7) & ~7) * 2;
}
__host__ __device__ inline int one_search_smem(bool bytes, int col0, int nx, int w, int rows) {
  return rows * one_search_row_bytes(bytes, col0, nx, w) + rows * w * (bytes ? 1 : 2);
}

// pattern: dense int16 rows (pitch = w).  BYTES: every pattern sample is in 0..255 and the plane holds bytes.
// key / ticket: the argmin key (~0 between calls) and the count of finished CTAs (0 between calls); the last CTA restores both.
template <bool BYTES, typename RefT>
__device__ __forceinline__ void one_search_body(const SearchTask& t, const hmb200_pu_result& seed, unsigned long long* __restrict__ key,
                                                uint32_t* __restrict__ ticket, hmb200_pu_result* __restrict__ out,
                                                const int16_t* __restrict__ pattern, const DevPlane& ref_plane, int fuse, const OneBack& back) {
  extern __shared__ __align__(16) uint8_t one_smem[];
  __shared__ unsigned long long s_best[ONE_SEARCH_THREADS_MAX / 32];
  __shared__ int s_last;
  __shared__ unsigned long long s_key;
  const int nx = t.rb_x - t.lt_x + 1;
  const int step = 1 << t.sub_shift, rows = t.h >> t.sub_shift;
  const int cy = blockIdx.x, y = t.lt_y + cy;
  const int col0 = t.ref_x + ref_plane.margin_x + t.lt_x;                  // plane column of candidate 0's first sample
  const int row0 = t.ref_y + ref_plane.margin_y + y;
  const int rp = one_search_row_bytes(BYTES, col0, nx, t.w);
  uint8_t* s_org8 = one_smem + rows * rp;
  unsigned long long best = ~0ull;

  if (BYTES) {
    const int nvec = rp / 16 - 1, vec0 = col0 >> 4;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(ref_plane.base);
    for (int i = threadIdx.x; i < rows * (nvec + 1); i += blockDim.x) {
      const int r = i / (nvec + 1), v = i - r * (nvec + 1);
      uint4 q = make_uint4(0, 0, 0, 0);
      if (v < nvec && (vec0 + v) * 16 < ref_plane.pitch)                   // the spare vector and anything past the row: zeros, never in a result
        q = *reinterpret_cast<const uint4*>(base + (size_t)(row0 + r * step) * ref_plane.pitch + (size_t)(vec0 + v) * 16);
      *reinterpret_cast<uint4*>(one_smem + r * rp + v * 16) = q;
    }
    for (int i = threadIdx.x; i < rows * t.w; i += blockDim.x) {
      const int r = i / t.w, c = i - r * t.w;
      s_org8[i] = (uint8_t)pattern[(r * step) * t.w + c];
    }
    __syncthreads();
    const int ww = t.w >> 2;
    const uint32_t* org32 = reinterpret_cast<const uint32_t*>(s_org8);
    for (int cx = threadIdx.x; cx < nx; cx += blockDim.x) {
      const int b = (col0 & 15) + cx, sh = (b & 3) * 8;
      const uint32_t* row = reinterpret_cast<const uint32_t*>(one_smem) + (b >> 2);
      uint32_t sum = 0;
      for (int r = 0; r < rows; r++, row += rp / 4) {
        uint32_t lo = row[0];
        for (int g = 0; g < ww; g++) {
          const uint32_t hi = row[g + 1];
          sum = sad4_acc(__funnelshift_r(lo, hi, sh), org32[r * ww + g], sum);
          lo = hi;
        }
      }
      const int x = t.lt_x + cx;
      sum <<= t.sub_shift;                                                  // bit depth 8: no precision shift
      const unsigned long long k = make_key(sum + mv_cost(t.lambda_cost, mv_bits(x, y, t.pred_x, t.pred_y, 2)), (uint32_t)(cy * nx + cx));
      best = k < best ? k : best;
    }
  } else {
    int16_t* s_ref = reinterpret_cast<int16_t*>(one_smem);
    int16_t* s_org = reinterpret_cast<int16_t*>(s_org8);
    const int rp16 = rp / 2, span = nx + t.w - 1;
    const RefT* base = reinterpret_cast<const RefT*>(ref_plane.base);
    for (int i = threadIdx.x; i < rows * span; i += blockDim.x) {
      const int r = i / span, c = i - r * span;
      s_ref[r * rp16 + c] = (int16_t)base[(size_t)(row0 + r * step) * ref_plane.pitch + col0 + c];
    }
    for (int i = threadIdx.x; i < rows * t.w; i += blockDim.x) {
      const int r = i / t.w, c = i - r * t.w;
      s_org[i] = pattern[(r * step) * t.w + c];
    }
    __syncthreads();
    for (int cx = threadIdx.x; cx < nx; cx += blockDim.x) {
      uint32_t sum = 0;
      for (int r = 0; r < rows; r++) {
        const int16_t* q = s_ref + r * rp16 + cx;
        const int16_t* o = s_org + r * t.w;
        for (int c = 0; c < t.w; c++) sum += (uint32_t)abs((int)o[c] - (int)q[c]);
      }
      const int x = t.lt_x + cx;
      sum = (sum << t.sub_shift) >> (ref_plane.bit_depth - 8);
      const unsigned long long k = make_key(sum + mv_cost(t.lambda_cost, mv_bits(x, y, t.pred_x, t.pred_y, 2)), (uint32_t)(cy * nx + cx));
      best = k < best ? k : best;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) best = s_best[w] < best ? s_best[w] : best;
    atomicMin(key, best);
    __threadfence();                                                        // this CTA's minimum is in the key before its ticket
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  // every CTA's minimum has arrived: key -> (rcMv, ruiSAD), TEncSearch.cpp:3839-3841
  if (threadIdx.x == 0) { __threadfence(); s_key = atomicExch(key, ~0ull); *ticket = 0; }
  __syncthreads();
  const unsigned long long k = s_key;
  const uint32_t idx = (uint32_t)(k & 0xffffffffu), cost = (uint32_t)(k >> 32);
  const int by = idx / nx, bx = idx - by * nx;
  hmb200_pu_result r = seed;
  r.mv_x = t.lt_x + bx; r.mv_y = t.lt_y + by;
  r.sad = cost - mv_cost(t.lambda_cost, mv_bits(r.mv_x, r.mv_y, t.pred_x, t.pred_y, 2));
  if (fuse & ONE_FUSE_FRAC)            // small PUs: this CTA refines the vector it just decoded - no second launch
    one_frac_body<RefT>(t, r, pattern, ref_plane, (fuse & ONE_FUSE_HAD) ? 1 : 0, reinterpret_cast<int16_t*>(one_smem));
  if (threadIdx.x == 0) { if (out) out[0] = r; one_report(back, r); }
}

template <bool BYTES, typename RefT>
__global__ void __launch_bounds__(ONE_SEARCH_THREADS_MAX)
k_one_search(const SearchTask t, hmb200_pu_result seed, unsigned long long* __restrict__ key, uint32_t* __restrict__ ticket,
             hmb200_pu_result* __restrict__ out, const int16_t* __restrict__ pattern, DevPlane ref_plane, int fuse, OneBack back) {
  one_search_body<BYTES, RefT>(t, seed, key, ticket, out, pattern, ref_plane, fuse, back);
}
// small PU: the pattern is an argument; nothing of the call lives in device memory but key and ticket
template <bool BYTES, typename RefT>
__global__ void __launch_bounds__(ONE_SEARCH_THREADS_MAX)
k_one_search_args(const SearchTask t, hmb200_pu_result seed, const __grid_constant__ OnePattern pat, unsigned long long* __restrict__ key,
                  uint32_t* __restrict__ ticket, DevPlane ref_plane, int fuse, OneBack back) {
  one_search_body<BYTES, RefT>(t, seed, key, ticket, nullptr, pat.px, ref_plane, fuse, back);
}

// ---------------------------------------------------------------------------------------------------------------
// One CU per call: when HM asks for the 2Nx2N PU of a CU, the other partitions of that CU (2NxN, Nx2N and the four AMP
// splits: 13 PUs, 5 for an 8x8 CU; TLibCommon/TComDataCU.cpp:1893-1931) are searched and refined in the same round trip
// with the 2Nx2N call's window, predictor and lambda.  The host keeps the answers and serves the following calls from them
// when - and only when - their window, predictor, lambda, flags, position and pattern samples are the ones used here
// (hmb200_api.cu, CuCache): exact, and HM's decision order is untouched.  In the 1080p I + P test clip 161,706 of the
// 180,580 non-2Nx2N calls qualify (profiles/r02_latency_1to1.txt).
//
// Every partition is a union of cells of a 4 x 4 grid over the CU (2 x 2 for an 8x8 CU), so a candidate's 13 SADs come
// from 16 cell sums - kept apart for even and odd rows, because under FEN a PU taller than 8 rows counts its even rows
// twice and a shorter one counts all rows (TEncSearch.cpp:3804-3810; PU offsets are multiples of 4, so PU-relative and
// CU-relative row parity agree).  BYTES: 8-bit plane, VABSDIFF4 on funnel-shifted words; otherwise (9..14-bit planes) scalar
// 16-bit samples.  Patterns that leave the sample range (bi-prediction's 2*org - pred) take the per-PU path.
// ---------------------------------------------------------------------------------------------------------------
// A 16x16 CU also carries the 5 PUs of each of its four 8x8 child CUs (its 4x4-sample cells are theirs too): in the test clip 26,743
// of the 32,400 2Nx2N calls of 8x8 CUs arrive with their parent's window, predictor and lambda.
constexpr int ONE_CU_MAX_PUS = 33;
__host__ __device__ constexpr int one_cu_pus(int S) { return S == 8 ? 5 : S == 16 ? 33 : 13; }
// partition p of an S x S CU: offset and size (p >= 13, S = 16: partition (p - 13) % 5 of child CU (p - 13) / 5)
__host__ __device__ inline void one_cu_part(int S, int p, int* ox, int* oy, int* w, int* h) {
  int bx = 0, by = 0;
  if (p >= 13) { const int c = (p - 13) / 5; p = (p - 13) % 5; bx = (c & 1) * (S / 2); by = (c >> 1) * (S / 2); S = S / 2; }
  const int H2 = S / 2, Q = S / 4;
  int x = bx, y = by, ww = S, hh = S;
  switch (p) {
    case 0: break;
    case 1: hh = H2; break;                 case 2: y += H2; hh = H2; break;             // 2NxN
    case 3: ww = H2; break;                 case 4: x += H2; ww = H2; break;             // Nx2N
    case 5: hh = Q; break;                  case 6: y += Q; hh = S - Q; break;           // 2NxnU
    case 7: hh = S - Q; break;              case 8: y += S - Q; hh = Q; break;           // 2NxnD
    case 9: ww = Q; break;                  case 10: x += Q; ww = S - Q; break;          // nLx2N
    case 11: ww = S - Q; break;             default: x += S - Q; ww = Q; break;          // nRx2N
  }
  *ox = x; *oy = y; *w = ww; *h = hh;
}

// key[p] / ticket as in one_search_body; out[p]: (rcMv, ruiSAD) of partition p.  t: the 2Nx2N task (w = h = S); fen: FLAG_FEN.
template <int S, bool BYTES>
__device__ __forceinline__ void one_cu_search_body(const SearchTask& t, int fen, unsigned long long* __restrict__ keys, uint32_t* __restrict__ ticket,
                                                   hmb200_pu_result* __restrict__ out, const int16_t* __restrict__ pattern, const DevPlane& ref_plane) {
  constexpr int NB = (S == 8) ? 2 : 4;           // cells per CU edge
  constexpr int BS = S / NB;                     // cell edge in samples (4, 4, 8, 16)
  constexpr int BW = BS / 4;                     // ... in 32-bit words
  constexpr int WW = S / 4;
  constexpr int NP = one_cu_pus(S);
  extern __shared__ __align__(16) uint8_t one_smem[];
  __shared__ unsigned long long s_best[NP][ONE_SEARCH_THREADS_MAX / 32];
  __shared__ int s_last;
  const int nx = t.rb_x - t.lt_x + 1;
  const bool odd_rows = !(S == 64 && fen);       // a 64x64 CU under FEN has no PU that reads odd rows
  const int step = odd_rows ? 1 : 2, rows = S / step;
  const int cy = blockIdx.x, y = t.lt_y + cy;
  const int col0 = t.ref_x + ref_plane.margin_x + t.lt_x;
  const int row0 = t.ref_y + ref_plane.margin_y + y;
  const int rp = one_search_row_bytes(BYTES, col0, nx, S);
  uint8_t* s_org8 = one_smem + rows * rp;
  if (BYTES) {
    const int nvec = rp / 16 - 1, vec0 = col0 >> 4;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(ref_plane.base);
    for (int i = threadIdx.x; i < rows * (nvec + 1); i += blockDim.x) {
      const int r = i / (nvec + 1), v = i - r * (nvec + 1);
      uint4 q = make_uint4(0, 0, 0, 0);
      if (v < nvec && (vec0 + v) * 16 < ref_plane.pitch)
        q = *reinterpret_cast<const uint4*>(base + (size_t)(row0 + r * step) * ref_plane.pitch + (size_t)(vec0 + v) * 16);
      *reinterpret_cast<uint4*>(one_smem + r * rp + v * 16) = q;
    }
    for (int i = threadIdx.x; i < rows * S; i += blockDim.x) {
      const int r = i / S, c = i - r * S;
      s_org8[i] = (uint8_t)pattern[(r * step) * S + c];
    }
  } else {
    int16_t* s_ref = reinterpret_cast<int16_t*>(one_smem);
    int16_t* s_org = reinterpret_cast<int16_t*>(s_org8);
    const int span = nx + S - 1;
    const int16_t* base = reinterpret_cast<const int16_t*>(ref_plane.base);
    for (int i = threadIdx.x; i < rows * span; i += blockDim.x) {
      const int r = i / span, c = i - r * span;
      s_ref[r * (rp / 2) + c] = base[(size_t)(row0 + r * step) * ref_plane.pitch + col0 + c];
    }
    for (int i = threadIdx.x; i < rows * S; i += blockDim.x) {
      const int r = i / S, c = i - r * S;
      s_org[i] = pattern[(r * step) * S + c];
    }
  }
  __syncthreads();
  unsigned long long best[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) best[p] = ~0ull;
  const uint32_t* org32 = reinterpret_cast<const uint32_t*>(s_org8);
  for (int cx = threadIdx.x; cx < nx; cx += blockDim.x) {
    uint32_t ce[NB][NB], co[NB][NB];             // cell sums over even / odd rows
#pragma unroll
    for (int i = 0; i < NB; i++)
#pragma unroll
      for (int j = 0; j < NB; j++) { ce[i][j] = 0; co[i][j] = 0; }
    if (BYTES) {
      const int b = (col0 & 15) + cx, sh = (b & 3) * 8;
      const uint32_t* row = reinterpret_cast<const uint32_t*>(one_smem) + (b >> 2);
#pragma unroll
      for (int rb = 0; rb < NB; rb++) {
#pragma unroll 1
        for (int rr = 0; rr < BS; rr += 2) {
#pragma unroll
          for (int par = 0; par < 2; par++) {
            if (par == 1 && !odd_rows) continue;
            const int r = (rb * BS + rr + par) / step;                    // staged row index
            const uint32_t* q = row + r * (rp / 4);
            const uint32_t* o = org32 + r * WW;
            uint32_t lo = q[0];
#pragma unroll
            for (int g = 0; g < WW; g++) {
              const uint32_t hi = q[g + 1];
              const uint32_t v = __funnelshift_r(lo, hi, sh);
              if (par == 0) ce[rb][g / BW] = sad4_acc(v, o[g], ce[rb][g / BW]);
              else          co[rb][g / BW] = sad4_acc(v, o[g], co[rb][g / BW]);
              lo = hi;
            }
          }
        }
      }
    } else {
      const int16_t* s_ref = reinterpret_cast<const int16_t*>(one_smem) + cx;
      const int16_t* s_org = reinterpret_cast<const int16_t*>(s_org8);
#pragma unroll
      for (int rb = 0; rb < NB; rb++) {
#pragma unroll 1
        for (int rr = 0; rr < BS; rr += 2) {
#pragma unroll
          for (int par = 0; par < 2; par++) {
            if (par == 1 && !odd_rows) continue;
            const int r = (rb * BS + rr + par) / step;
            const int16_t* q = s_ref + r * (rp / 2);
            const int16_t* o = s_org + r * S;
#pragma unroll
            for (int cb = 0; cb < NB; cb++) {
              uint32_t a = 0;
#pragma unroll 8
              for (int c = 0; c < BS; c++) a += (uint32_t)abs((int)o[cb * BS + c] - (int)q[cb * BS + c]);
              if (par == 0) ce[rb][cb] += a; else co[rb][cb] += a;
            }
          }
        }
      }
    }
    const int x = t.lt_x + cx;
    const uint32_t mvc = mv_cost(t.lambda_cost, mv_bits(x, y, t.pred_x, t.pred_y, 2));
    const uint32_t idx = (uint32_t)(cy * nx + cx);
#pragma unroll
    for (int p = 0; p < NP; p++) {
      int ox, oy, w, h;
      one_cu_part(S, p, &ox, &oy, &w, &h);
      uint32_t e = 0, od = 0;
#pragma unroll
      for (int i = 0; i < NB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++)
          if (i * BS >= oy && i * BS < oy + h && j * BS >= ox && j * BS < ox + w) { e += ce[i][j]; od += co[i][j]; }
      const uint32_t sad = ((fen && h > 8) ? (e << 1) : (e + od)) >> (ref_plane.bit_depth - 8);
      const unsigned long long k = make_key(sad + mvc, idx);
      best[p] = k < best[p] ? k : best[p];
    }
  }
#pragma unroll
  for (int p = 0; p < NP; p++) {
    unsigned long long v = best[p];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
      v = other < v ? other : v;
    }
    if ((threadIdx.x & 31) == 0) s_best[p][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < NP) {
    unsigned long long v = s_best[threadIdx.x][0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) v = s_best[threadIdx.x][w] < v ? s_best[threadIdx.x][w] : v;
    atomicMin(&keys[threadIdx.x], v);
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x < NP) {
    __threadfence();
    const unsigned long long k = atomicExch(&keys[threadIdx.x], ~0ull);
    const uint32_t idx = (uint32_t)(k & 0xffffffffu), cost = (uint32_t)(k >> 32);
    const int by = idx / nx, bx = idx - by * nx;
    hmb200_pu_result r;
    r.mv_x = t.lt_x + bx; r.mv_y = t.lt_y + by;
    r.sad = cost - mv_cost(t.lambda_cost, mv_bits(r.mv_x, r.mv_y, t.pred_x, t.pred_y, 2));
    r.half_x = r.half_y = r.qter_x = r.qter_y = 0; r.frac_cost = 0;
    out[threadIdx.x] = r;
  }
  if (threadIdx.x == 0) *ticket = 0;
}

template <int S, bool BYTES>
__global__ void __launch_bounds__(ONE_SEARCH_THREADS_MAX)
k_one_cu_search(const SearchTask t, int fen, unsigned long long* __restrict__ keys, uint32_t* __restrict__ ticket, hmb200_pu_result* __restrict__ out,
                const int16_t* __restrict__ pattern, DevPlane ref_plane) {
  one_cu_search_body<S, BYTES>(t, fen, keys, ticket, out, pattern, ref_plane);
}
template <int S, bool BYTES>
__global__ void __launch_bounds__(ONE_SEARCH_THREADS_MAX)
k_one_cu_search_args(const SearchTask t, int fen, const __grid_constant__ OnePattern pat, unsigned long long* __restrict__ keys,
                     uint32_t* __restrict__ ticket, hmb200_pu_result* __restrict__ out, DevPlane ref_plane) {
  one_cu_search_body<S, BYTES>(t, fen, keys, ticket, out, pat.px, ref_plane);
}

// refinement of every partition: CTA p refines partition p at the vector k_one_cu_search left in out[p] and reports to host slot p
// (two 16-byte records, 32 bytes per slot)
template <typename RefT>
__device__ __forceinline__ void one_cu_frac_body(const SearchTask& t, int S, const int16_t* __restrict__ pattern, hmb200_pu_result* __restrict__ out,
                                                 const DevPlane& ref_plane, int use_had, const OneBack& back) {
  extern __shared__ __align__(16) int16_t one_smem16[];
  const int p = blockIdx.x;
  int ox, oy, w, h;
  one_cu_part(S, p, &ox, &oy, &w, &h);
  SearchTask tp = t;
  tp.ref_x += ox; tp.ref_y += oy; tp.w = w; tp.h = h;
  hmb200_pu_result res = out[p];
  one_frac_body<RefT>(tp, res, pattern + oy * S + ox, ref_plane, use_had, one_smem16, S);
  if (threadIdx.x == 0) {
    OneBack slot = back;
    slot.host_a = back.host_a + 2 * p; slot.host_b = back.host_a + 2 * p + 1;
    one_report(slot, res);
  }
}
template <typename RefT>
__global__ void __launch_bounds__(ONE_FRAC_THREADS_MAX)
k_one_cu_frac(const SearchTask t, int S, const int16_t* __restrict__ pattern, hmb200_pu_result* __restrict__ out, DevPlane ref_plane, int use_had, OneBack back) {
  one_cu_frac_body<RefT>(t, S, pattern, out, ref_plane, use_had, back);
}
template <typename RefT>
__global__ void __launch_bounds__(ONE_FRAC_THREADS_MAX)
k_one_cu_frac_args(const SearchTask t, int S, const __grid_constant__ OnePattern pat, hmb200_pu_result* __restrict__ out, DevPlane ref_plane, int use_had,
                   OneBack back) {
  one_cu_frac_body<RefT>(t, S, pat.px, out, ref_plane, use_had, back);
}

// ends a call whose last kernel is one of the older ones (TZ search without refinement): one thread reports the record
__global__ void k_one_report(const hmb200_pu_result* __restrict__ out, OneBack back) { one_report(back, out[0]); }

}  // namespace hmb200
