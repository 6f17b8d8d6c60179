// Device-side arithmetic shared by all kernels of the HM-16.5 motion-search path.
// Every helper cites the reference lines (hm-16.5rc1/source/Lib/...) whose result it reproduces bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/hmb200.h"

namespace hmb200 {

// ---------------------------------------------------------------------------------------------------------------
// records shared with the host frontend
// ---------------------------------------------------------------------------------------------------------------

struct DevPlane {
  void*   base;        // first byte of the padded buffer (row -margin_y, column -margin_x)
  int32_t pitch;       // elements per row (multiple of 128 bytes)
  int32_t width, height;
  int32_t margin_x, margin_y;
  int32_t bytes_per_sample;   // 1 (8-bit content) or 2
  int32_t bit_depth;
};

template <typename T>
__device__ __forceinline__ const T* plane_at(const DevPlane& p, int x, int y) {
  return reinterpret_cast<const T*>(p.base) + (size_t)(y + p.margin_y) * p.pitch + (x + p.margin_x);
}

// ---------------------------------------------------------------------------------------------------------------
// MV rate: TLibCommon/TComRdCost.cpp:279-292 (xGetExpGolombNumberOfBits), TComRdCost.h:172-189 (getCost/getBits)
// ---------------------------------------------------------------------------------------------------------------

// 1 + 2*floor(log2(v <= 0 ? -2v+1 : 2v))
__device__ __forceinline__ uint32_t eg_bits(int v) {
  uint32_t t = (v <= 0) ? (((uint32_t)(-v)) << 1) + 1u : ((uint32_t)v << 1);
  return 2u * (31u - (uint32_t)__clz(t)) + 1u;
}
__device__ __forceinline__ uint32_t mv_bits(int x, int y, int pred_x, int pred_y, int scale) {
  return eg_bits(x * (1 << scale) - pred_x) + eg_bits(y * (1 << scale) - pred_y);
}
// (m_uiCost * bits) >> 16 in 32-bit unsigned arithmetic (wraps like the reference's UInt)
__device__ __forceinline__ uint32_t mv_cost(uint32_t lambda_cost, uint32_t bits) {
  return (lambda_cost * bits) >> 16;
}

// ---------------------------------------------------------------------------------------------------------------
// byte-SIMD SAD primitives (SASS: VABSDIFF4.U8.ACC with the accumulate fused, measured 64 lane-ops/clk/SM)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// ---------------------------------------------------------------------------------------------------------------
// Hadamard SATD tiles: TLibCommon/TComRdCost.cpp:1332-1523.  Any un-normalised 2-D Hadamard gives the same
// sum of absolute coefficients; only the per-tile rounding is normative.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t had8x8_abs(int (&d)[64]) {
#pragma unroll
  for (int y = 0; y < 8; y++) {
    int* r = &d[y * 8];
#pragma unroll
    for (int len = 1; len < 8; len <<= 1)
#pragma unroll
      for (int b = 0; b < 8; b += 2 * len)
#pragma unroll
        for (int k = 0; k < len; k++) { int a0 = r[b + k], a1 = r[b + k + len]; r[b + k] = a0 + a1; r[b + k + len] = a0 - a1; }
  }
  uint32_t s = 0;
#pragma unroll
  for (int x = 0; x < 8; x++) {
#pragma unroll
    for (int len = 1; len < 8; len <<= 1)
#pragma unroll
      for (int b = 0; b < 8; b += 2 * len)
#pragma unroll
        for (int k = 0; k < len; k++) {
          int a0 = d[(b + k) * 8 + x], a1 = d[(b + k + len) * 8 + x];
          d[(b + k) * 8 + x] = a0 + a1; d[(b + k + len) * 8 + x] = a0 - a1;
        }
#pragma unroll
    for (int y = 0; y < 8; y++) s += (uint32_t)abs(d[y * 8 + x]);
  }
  return s;
}

__device__ __forceinline__ uint32_t had4x4_abs(int (&d)[16]) {
#pragma unroll
  for (int y = 0; y < 4; y++) {
    int a = d[y * 4 + 0], b = d[y * 4 + 1], c = d[y * 4 + 2], e = d[y * 4 + 3];
    int s0 = a + b, s1 = a - b, s2 = c + e, s3 = c - e;
    d[y * 4 + 0] = s0 + s2; d[y * 4 + 1] = s1 + s3; d[y * 4 + 2] = s0 - s2; d[y * 4 + 3] = s1 - s3;
  }
  uint32_t s = 0;
#pragma unroll
  for (int x = 0; x < 4; x++) {
    int a = d[0 + x], b = d[4 + x], c = d[8 + x], e = d[12 + x];
    int s0 = a + b, s1 = a - b, s2 = c + e, s3 = c - e;
    s += (uint32_t)abs(s0 + s2) + (uint32_t)abs(s1 + s3) + (uint32_t)abs(s0 - s2) + (uint32_t)abs(s1 - s3);
  }
  return s;
}

__device__ __forceinline__ uint32_t had2x2_abs(int a, int b, int c, int e) {
  int m0 = a + c, m1 = b + e, m2 = a - c, m3 = b - e;
  return (uint32_t)abs(m0 + m1) + (uint32_t)abs(m0 - m1) + (uint32_t)abs(m2 + m3) + (uint32_t)abs(m2 - m3);
}

// ---------------------------------------------------------------------------------------------------------------
// luma interpolation: TLibCommon/TComInterpolationFilter.cpp:57-63 (taps), :94-154 (filterCopy), :172-257 (filter<>)
// ---------------------------------------------------------------------------------------------------------------
__device__ __constant__ const int8_t k_luma_taps[4][8] = {
  {  0, 0,   0, 64,  0,   0, 0,  0 },
  { -1, 4, -10, 58, 17,  -5, 1,  0 },
  { -1, 4, -11, 40, 40, -11, 4, -1 },
  {  0, 1,  -5, 17, 58, -10, 4, -1 } };

// first (horizontal) pass into the 14-bit intermediate domain; s[0..7] are the samples at columns -3..+4
__device__ __forceinline__ int16_t interp_h(const int (&s)[8], int fx, int head) {
  if (fx == 0) return (int16_t)((int16_t)(s[3] << head) - 8192);
  int sum = 0;
#pragma unroll
  for (int t = 0; t < 8; t++) sum += s[t] * (int)k_luma_taps[fx][t];
  int shift = 6 - head;
  return (int16_t)((sum - (8192 << shift)) >> shift);
}
// second (vertical) pass back to the pixel domain with clipping; c[0..7] are intermediates at rows -3..+4
__device__ __forceinline__ int interp_v(const int (&c)[8], int fy, int head, int maxv) {
  int val;
  if (fy == 0) val = (int16_t)((c[3] + 8192 + (1 << (head - 1))) >> head);
  else {
    int sum = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) sum += c[t] * (int)k_luma_taps[fy][t];
    int shift = 6 + head;
    val = (int16_t)((sum + (1 << (shift - 1)) + (8192 << 6)) >> shift);
  }
  return min(max(val, 0), maxv);
}

// TLibEncoder/TEncSearch.cpp:51-75
__device__ __constant__ const int8_t k_refine_h[9][2] = { {0,0},{0,-1},{0,1},{-1,0},{1,0},{-1,-1},{1,-1},{-1,1},{1,1} };
__device__ __constant__ const int8_t k_refine_q[9][2] = { {0,0},{0,-1},{0,1},{-1,-1},{1,-1},{-1,0},{1,0},{-1,1},{1,1} };

__device__ __forceinline__ int floor_div4(int v) { return v >> 2; }   // arithmetic shift

// lexicographic (cost, raster index) key: strict '<' + raster scan order == first-wins (TEncSearch.cpp:3813-3835)
__device__ __forceinline__ unsigned long long make_key(uint32_t cost, uint32_t idx) {
  return ((unsigned long long)cost << 32) | idx;
}

} // namespace hmb200
