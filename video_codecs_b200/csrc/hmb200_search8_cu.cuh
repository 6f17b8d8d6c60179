// CU-fused integer full search for 8-bit planes: all partitions of one CU in one pass over its samples.
//
// Inside HM every PU of a CU is searched separately (TEncCu::xCheckRDCostInter -> predInterSearch -> xPatternSearch,
// TLibEncoder/TEncCu.cpp:459-626, TEncSearch.cpp:2912-3063), so a sample of the CU is visited once per partition mode:
// 2Nx2N, 2NxN, Nx2N and the four AMP modes (TLibCommon/TComDataCU.cpp:1893-1931) = 7 times (3 for 8x8 CUs).  When
// those PUs share the search window and the MV predictor (xSetSearchRange clips by CU, TEncSearch.cpp:3765-3781; the
// canonical job list gives every PU of a CU the same predictor) their SADs at one displacement are sums of the same
// partial SADs.  This kernel computes, per candidate, a 4x4 grid of partial SADs over the CU (granule S/4 — every PU
// boundary of every partition mode lies on that grid) and derives all 13 PU SADs from row / column sums:
//
//   FEN (iSubShift, TEncSearch.cpp:3804-3810, TComRdCost.cpp:564-595): PUs with more than 8 rows visit only even
//   rows and shift the sum left by one; PUs with <= 8 rows visit every row.  Even rows accumulate into the grid E,
//   odd rows into per-strip sums O that only the <= 8-row PUs add (16x8, 16x4 in a 16x16 CU; 32x8 in a 32x32 CU).
//
// Every PU keeps its own running argmin key, so results are bit-identical to thirteen separate searches; what changes
// is the number of executed abs-diffs (1 pass over the CU instead of 3.5-4.25), not the arithmetic.
// Tile: a quad of lanes owns 16 candidate columns (lane s: columns s, s+4, s+8, s+12, one funnel shift per word) of
// KY candidate rows.  Windows whose width is not a multiple of 16 get a last block that overlaps its neighbour.
#pragma once
#include <cstdlib>
#include <map>
#include "hmb200_search8.cuh"

namespace hmb200 {

constexpr int CU_SLOTS = 13;
// A 16x16 CU whose four 8x8 child CUs share its window, predictor and lambda carries their PUs along (CHILD kernels):
// 4 children x (8x8, 8x4 top, 8x4 bottom, 4x8 left, 4x8 right)
constexpr int CU_CHILD_SLOTS = 20;
constexpr int CU_SLOTS_ALL = CU_SLOTS + CU_CHILD_SLOTS;
// 8-bit CU-fused kernels: ONE CTA of 16 warps per SM, because the window is staged four times (byte phases 0..3, see
// cu_load_ref): 4 x 37 KB for a CTU at +-64.  The 16-bit kernels keep two CTAs of 8 warps.
constexpr int CU8_THREADS = 512;
constexpr int CU8_WARPS = CU8_THREADS / 32;
constexpr int CU_PX_BLOCKS = 16;             // windows up to 256 columns keep their per-column MV-cost terms in shared memory

struct S8Bundle {             // 256 bytes; the 16-bit kernels read the first 128 only
  int32_t org_off, win_off;   // byte offsets of the CU's top-left sample / of candidate (lt_x, lt_y) in the staged tiles
  int32_t nx, ny;
  int32_t lt_x, lt_y, pred_x, pred_y;
  uint32_t lambda;
  int32_t n_blk;              // blocks per candidate row that run as block items: ceil((xal + nx) / 16) for 8-bit planes (one less when
                              // the last column runs as edge items), ceil(nx / 8) for 16-bit
  int32_t item_start, n_items;
  int32_t out_idx[CU_SLOTS];  // task index per partition slot (-1: that PU is not in the job list)
  int32_t n_rowgroups;        // ceil(ny / KY)
  int32_t cy_first;           // first candidate row of this bundle in the PU's window (tall windows are split by rows)
  int32_t shr;                // bitDepth - 8 (distortion precision adjustment); 0 for 8-bit planes
  int32_t xal;                // 8-bit planes: bytes between the 16-byte boundary below candidate column 0 and that column (blocks are
                              // aligned in shared memory; columns of the first / last block outside the window are masked)
  int32_t step_g, step_blk;   // (quads per item step) / n_blk and % n_blk: how far a quad moves per item step
  int32_t rank_bits;          // 8x8 CUs (four candidate rows per tile): bits that number a lane's tiles inside one row group
  // ---- 8-bit kernels only ----
  int32_t child_idx[CU_CHILD_SLOTS];   // CHILD kernels: task index per PU of the four 8x8 child CUs (-1: not in the job list)
  int32_t n_main;             // block items; the items behind them (n_items - n_main) are edge items
  int32_t edge;               // 1: (xal + nx) % 16 == 1, i.e. the window's last column would be a block of its own with 15 of 16
                              // columns masked (129 = 8 * 16 + 1): it runs as edge items instead, a lane per candidate row
  int32_t pad_[10];
};
static_assert(sizeof(S8Bundle) == 256, "S8Bundle is loaded as 64 ints per warp");

// partition slots: 0 2Nx2N | 1,2 2NxN top,bottom | 3,4 Nx2N left,right | 5,6 2NxnU | 7,8 2NxnD | 9,10 nLx2N | 11,12 nRx2N
__host__ __device__ constexpr int cu_slot_h(int S, int slot) {
  return (slot == 1 || slot == 2) ? S / 2 : (slot == 5 || slot == 8) ? S / 4 : (slot == 6 || slot == 7) ? 3 * S / 4 : S;
}
__host__ __device__ constexpr int cu_slot_w(int S, int slot) {
  return (slot == 3 || slot == 4) ? S / 2 : (slot == 9 || slot == 12) ? S / 4 : (slot == 10 || slot == 11) ? 3 * S / 4 : S;
}
__host__ __device__ constexpr int cu_slot_x(int S, int slot) { return slot == 4 ? S / 2 : slot == 10 ? S / 4 : slot == 12 ? 3 * S / 4 : 0; }
__host__ __device__ constexpr int cu_slot_y(int S, int slot) { return slot == 2 ? S / 2 : slot == 6 ? S / 4 : slot == 8 ? 3 * S / 4 : 0; }
__host__ __device__ constexpr int cu_ky(int S) { return S == 8 ? 4 : 1; }

template <int S, bool FEN> struct CuTraits {
  static constexpr int WW = S / 4;                 // 32-bit words per CU row
  static constexpr int N = (S == 8) ? 2 : 4;       // grid is N x N cells
  static constexpr int G = S / N;                  // rows per strip
  static constexpr int WC = WW / N;                // words per cell column
  static constexpr int KY = cu_ky(S);
  static constexpr bool PARITY = FEN && S >= 16;   // even / odd rows are told apart
  static constexpr bool ODD_ALL = PARITY && S == 16;   // odd rows needed in every strip
  static constexpr bool ODD_EDGE = PARITY && S == 32;  // odd rows needed in strips 0 and 3 only
};

// Shared-memory access by 32-bit shared address.  Row addresses are base + row * pitch with the row an immediate: written
// as mad.lo so that they run on the FMA pipe (IMAD) - the ALU pipe is the one VABSDIFF4 needs, and pointer increments or
// LEA would sit on it.  The loads are plain ld.shared (no volatile: the tiles are read-only after the staging barrier).
__device__ __forceinline__ uint32_t cu_row_addr(uint32_t base, int row, uint32_t pitch) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(pitch), "r"(row), "r"(base));
  return r;
}
__device__ __forceinline__ uint4 cu_lds128(uint32_t a) {
  uint4 v;
  asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 cu_lds64(uint32_t a) {
  uint2 v;
  asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t cu_lds32(uint32_t a) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}

// broadcast load of one original row (WW words) from shared memory; the CU is S-byte aligned in the tile
template <int WW>
__device__ __forceinline__ void cu_load_org(uint32_t p, uint32_t (&o)[WW]) {
  if constexpr (WW >= 4) {
#pragma unroll
    for (int i = 0; i < WW / 4; i++) {
      const uint4 v = cu_lds128(p + 16 * i);
      o[4 * i] = v.x; o[4 * i + 1] = v.y; o[4 * i + 2] = v.z; o[4 * i + 3] = v.w;
    }
  } else {
    const uint2 v = cu_lds64(p);
    o[0] = v.x; o[1] = v.y;
  }
}

// The staged window exists four times in shared memory: copy s holds the rows shifted left by s bytes (built once per
// CTA after the bulk copies land).  A lane whose candidates start s bytes past a 16-byte boundary reads copy s with
// aligned LDS.128 and gets its reference words ready to use: no per-word funnel shift in the inner loop (the ALU pipe
// that executes VABSDIFF4 also executes SHF, and it was the bound).  NW = WW + KC - 1 words serve the lane's KC candidates.
template <int NW>
__device__ __forceinline__ void cu_load_ref(uint32_t p, uint32_t (&w)[NW]) {
#pragma unroll
  for (int i = 0; i < NW / 4; i++) {
    const uint4 v = cu_lds128(p + 16 * i);
    w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
  }
  if constexpr (NW % 4 == 3) {                           // WW = 4, 8, 16 with four candidates: 7, 11, 19 words -> the last LDS.128 carries one spare word
    const uint4 v = cu_lds128(p + 16 * (NW / 4));
    w[NW - 3] = v.x; w[NW - 2] = v.y; w[NW - 1] = v.z;
  } else if constexpr (NW % 4 == 2) {
    const uint2 v = cu_lds64(p + 16 * (NW / 4));
    w[NW - 2] = v.x; w[NW - 1] = v.y;
  } else if constexpr (NW % 4 == 1) {                    // WW = 2: five words
    w[NW - 1] = cu_lds32(p + 16 * (NW / 4));
  }
}

// one reference row against one original row: the words of cell column c feed acc[c][k] (k = candidate column, 4 bytes apart)
template <int WW, int NC, int KC>
__device__ __forceinline__ void cu_row(uint32_t rp8, const uint32_t (&o)[WW], uint32_t (&acc)[NC][KC]) {
  uint32_t w[WW + KC - 1];
  cu_load_ref<WW + KC - 1>(rp8, w);
#pragma unroll
  for (int j = 0; j < WW + KC - 1; j++) {
#pragma unroll
    for (int k = 0; k < KC; k++) {
      const int i = j - k;
      if (i >= 0 && i < WW) acc[i / (WW / NC)][k] = sad4_acc(w[j], o[i], acc[i / (WW / NC)][k]);
    }
  }
}

// same, every word into one accumulator per candidate column (odd-row strips)
template <int WW, int KC>
__device__ __forceinline__ void cu_row1(uint32_t rp8, const uint32_t (&o)[WW], uint32_t (&acc)[KC]) {
  uint32_t w[WW + KC - 1];
  cu_load_ref<WW + KC - 1>(rp8, w);
#pragma unroll
  for (int j = 0; j < WW + KC - 1; j++) {
#pragma unroll
    for (int k = 0; k < KC; k++) {
      const int i = j - k;
      if (i >= 0 && i < WW) acc[k] = sad4_acc(w[j], o[i], acc[k]);
    }
  }
}

// Per-lane running argmin as ONE 32-bit key: (cost << CU_LOCAL_BITS) | local index.  cost < 2^21 for 8-bit content
// (SAD <= 64*32*255*2, MV cost < 2^16); the local index numbers this lane's candidates of the current CU in the order
// it visits them (= raster order inside the lane), so an unsigned min is "strict '<', first wins" (TEncSearch.cpp:
// 3813-3835).  Every warp takes a run of consecutive items; the tile loop says how the local index is formed.
constexpr int CU_LOCAL_BITS = 11;
// Block columns outside the window (blocks are aligned in shared memory, not in the window) must never win.  For CUs up
// to 32x32 their keys start at CU_KEY_MASKED = 2^32 - 1 - maxSAD * 2^11: key = sad * 2^11 + base cannot wrap, and no
// real key (cost < 32*32*255 + 2^16) reaches it.  64x64 CUs have no such gap in 32 bits and skip the update instead.
__host__ __device__ constexpr uint32_t cu_key_masked(int S) { return 0xffffffffu - ((uint32_t)(S * S * 255) << CU_LOCAL_BITS); }

// keys of candidate columns ka, kb -> running minima (VIMNMX3 when both columns are given)
__device__ __forceinline__ void cu_min2(uint32_t& best, uint32_t a, uint32_t b) { best = min(min(best, a), b); }

// S >= 16: per candidate column k, P[r][c] = partial SAD of cell column c over the strips 0..r (prefix sums come for
// free: VABSDIFF4 takes its addend from the strip above), OP[r] = odd rows of the strips 0..r (FEN only; for S = 32
// only strips 0 and 3 have odd rows that any PU needs, OP[3] = strip 0 + strip 3).  Derives the 13 PU keys of the
// columns KA and KB (KB < 0: one column) and updates the argmins.  key(x, M) = x * M + base: one IMAD per PU; a PU
// that is the complement of another inside the CU takes key(total) - x * M (still one IMAD, no sum of its own).
#ifndef CU_ADD_ON_FMA
#define CU_ADD_ON_FMA 0
#endif
template <int S, bool FEN, int KC>
struct CuKeys {
  uint32_t v[CU_SLOTS];
  // a + b as a * one + b with a run-time 1: IMAD on the FMA pipe instead of IADD3 on the ALU pipe
  static __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b, uint32_t one) { return CU_ADD_ON_FMA ? a * one + b : a + b; }
  static __device__ __forceinline__ uint32_t add4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t one) {
    return CU_ADD_ON_FMA ? add(add(a, b, one), add(c, d, one), one) : a + b + c + d;
  }
  __device__ __forceinline__ CuKeys(const uint32_t (&P)[4][4][KC], const uint32_t (&OP)[4][KC], int k, uint32_t base, uint32_t one) {
    // a PU with <= 8 rows under FEN visits every row: add the odd rows of its strips; all others shift by iSubShift
    constexpr bool f1 = FEN && cu_slot_h(S, 1) <= 8;      // 2NxN halves (S == 16)
    constexpr bool f5 = FEN && cu_slot_h(S, 5) <= 8;      // AMP quarter strips (S == 16, 32)
    constexpr uint32_t M1 = 1u << CU_LOCAL_BITS;          // key scale of a full-row PU
    constexpr uint32_t MS = M1 << (FEN ? 1 : 0);          // ... of a PU with iSubShift = FEN
    const uint32_t r0 = add4(P[0][0][k], P[0][1][k], P[0][2][k], P[0][3][k], one);      // strip 0
    const uint32_t r2 = add4(P[2][0][k], P[2][1][k], P[2][2][k], P[2][3][k], one);      // strips 0..2
    const uint32_t c0 = P[3][0][k], c3 = P[3][3][k], left = add(c0, P[3][1][k], one);
    const uint32_t tot = add(left, add(P[3][2][k], c3, one), one);
    const uint32_t k0 = tot * MS + base;
    v[0] = k0;
    if constexpr (f1) {          // S == 16 under FEN: 16x8 halves visit every row
      const uint32_t a = add(add4(P[1][0][k], P[1][1][k], P[1][2][k], P[1][3][k], one), OP[1][k], one);
      const uint32_t kb = OP[3][k] * M1 + (tot * M1 + base);
      v[1] = a * M1 + base;
      v[2] = kb - a * M1;
      v[5] = OP[0][k] * M1 + (r0 * M1 + base);
      v[8] = (kb - r2 * M1) - OP[2][k] * M1;
    } else {
      const uint32_t r1 = add4(P[1][0][k], P[1][1][k], P[1][2][k], P[1][3][k], one);    // strips 0..1
      v[1] = r1 * MS + base;
      v[2] = k0 - r1 * MS;
      if constexpr (f5) {        // S == 32 under FEN: 32x8 strips visit every row
        const uint32_t kb = OP[3][k] * M1 + (tot * M1 + base);
        v[5] = OP[0][k] * M1 + (r0 * M1 + base);
        v[8] = (kb - r2 * M1) - OP[0][k] * M1;
      } else {
        v[5] = r0 * MS + base;
        v[8] = k0 - r2 * MS;
      }
    }
    v[3] = left * MS + base;
    v[4] = k0 - left * MS;
    v[6] = k0 - r0 * MS;
    v[7] = r2 * MS + base;
    v[9] = c0 * MS + base;
    v[10] = k0 - c0 * MS;
    v[11] = k0 - c3 * MS;
    v[12] = c3 * MS + base;
  }
};

// One tile of a 16x16 / 32x32 / 64x64 CU: KC candidate columns (4 bytes apart) of one candidate row against the whole CU,
// all 13 partitions.  base[k]: MV cost and local index of column k, pre-shifted (KEY_NONE for a column that takes no part
// when the mask is part of the key, else valid[k] says so).
template <int S, bool FEN, int KC, bool MASK_BY_KEY, int NS>
__device__ __forceinline__ void cu_tile(uint32_t refp, uint32_t orgp, uint32_t rpitch, uint32_t opitch, const uint32_t (&base)[KC],
                                        const bool (&valid)[KC], uint32_t one, uint32_t (&best)[NS]) {
  typedef CuTraits<S, FEN> T;
  uint32_t P[4][4][KC], OP[4][KC];
#pragma unroll
  for (int r = 0; r < 4; r++) {                                      // strips: static, so the grid row index is too
    const bool odd_here = T::ODD_ALL || (T::ODD_EDGE && (r == 0 || r == 3));
    constexpr int OPREV[4] = {0, 0, 1, T::ODD_ALL ? 2 : 0};          // the strip whose odd-row sums strip r continues
#pragma unroll
    for (int k = 0; k < KC; k++) {
#pragma unroll
      for (int c = 0; c < 4; c++) P[r][c][k] = r ? P[r - 1][c][k] : 0u;
      OP[r][k] = r ? OP[OPREV[r]][k] : 0u;
    }
    if constexpr (T::PARITY) {
#pragma unroll 4
      for (int rr = 0; rr < T::G; rr += 2) {
        const int row = r * T::G + rr;
        uint32_t o[T::WW];
        cu_load_org<T::WW>(cu_row_addr(orgp, row, opitch), o);
        cu_row<T::WW, 4, KC>(cu_row_addr(refp, row, rpitch), o, P[r]);
        if (odd_here) {
          uint32_t o1[T::WW];
          cu_load_org<T::WW>(cu_row_addr(orgp, row + 1, opitch), o1);
          cu_row1<T::WW, KC>(cu_row_addr(refp, row + 1, rpitch), o1, OP[r]);
        }
      }
    } else {
#pragma unroll 4
      for (int rr = 0; rr < T::G; rr++) {
        const int row = r * T::G + rr;
        uint32_t o[T::WW];
        cu_load_org<T::WW>(cu_row_addr(orgp, row, opitch), o);
        cu_row<T::WW, 4, KC>(cu_row_addr(refp, row, rpitch), o, P[r]);
      }
    }
  }
  if constexpr (MASK_BY_KEY && KC >= 2) {
#pragma unroll
    for (int kk = 0; kk < KC; kk += 2) {
      const CuKeys<S, FEN, KC> a(P, OP, kk, base[kk], one), b(P, OP, kk + 1, base[kk + 1], one);
#pragma unroll
      for (int s = 0; s < CU_SLOTS; s++) cu_min2(best[s], a.v[s], b.v[s]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < KC; k++)
      if (MASK_BY_KEY || valid[k]) {
        const CuKeys<S, FEN, KC> a(P, OP, k, base[k], one);
#pragma unroll
        for (int s = 0; s < CU_SLOTS; s++) best[s] = min(best[s], a.v[s]);
      }
  }
}

// One tile of a 16x16 CU that carries its four 8x8 child CUs: every row is visited once (the children's PUs have at most 8
// rows, so none of them is sub-sampled: TEncSearch.cpp:3804-3810) and the CU is walked in two halves of two 4-row strips.
// Per half and candidate column: E[q][c] = even rows (all rows without FEN) of cell column c over the half's strips 0..q,
// O[q][c] = odd rows (FEN).  A child (half h, column pair j) is the sum of four cells; the 16x16 CU's own partitions need
// from the upper half only R0, R1 (strip / half sums of the even rows) and the column sums C0, C0+C1, C3.
// best[13 + 5 * (2 h + j) + {0: 8x8, 1: 8x4 top, 2: 8x4 bottom, 3: 4x8 left, 4: 4x8 right}] are the children's argmins.
template <bool FEN, int KC, int NS>
__device__ __forceinline__ void cu16_child_tile(uint32_t refp, uint32_t orgp, uint32_t rpitch, uint32_t opitch,
                                                const uint32_t (&base)[KC], uint32_t (&best)[NS]) {
  static_assert(NS == CU_SLOTS_ALL, "child tiles keep 33 argmins");
  constexpr uint32_t M1 = 1u << CU_LOCAL_BITS;          // key scale of a PU that visits every row (<= 8 rows)
  constexpr uint32_t MS = M1 << (FEN ? 1 : 0);          // ... of a PU with iSubShift = FEN (> 8 rows)
  constexpr int KP = KC >= 2 ? 2 : 1;                   // candidate columns whose keys are formed together (one VIMNMX3 per pair)
  uint32_t R0[KC], R1[KC], C0[KC], C01[KC], C3[KC];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    uint32_t E[2][4][KC], O[2][4][KC];
#pragma unroll
    for (int q = 0; q < 2; q++) {
#pragma unroll
      for (int c = 0; c < 4; c++)
#pragma unroll
        for (int k = 0; k < KC; k++) { E[q][c][k] = q ? E[0][c][k] : 0u; O[q][c][k] = q ? O[0][c][k] : 0u; }
#pragma unroll
      for (int rr = 0; rr < 4; rr++) {
        const int row = 8 * h + 4 * q + rr;
        uint32_t o[4];
        cu_load_org<4>(cu_row_addr(orgp, row, opitch), o);
        if (FEN && (rr & 1)) cu_row<4, 4, KC>(cu_row_addr(refp, row, rpitch), o, O[q]);
        else                 cu_row<4, 4, KC>(cu_row_addr(refp, row, rpitch), o, E[q]);
      }
    }
#pragma unroll
    for (int kk = 0; kk < KC; kk += KP) {
      uint32_t kc[KP][10], kh[KP][2], kf[KP][9];
#pragma unroll
      for (int u = 0; u < KP; u++) {
        const int k = kk + u;
        const uint32_t b = base[k];
        uint32_t T0[4], T1[4];
#pragma unroll
        for (int c = 0; c < 4; c++) { T0[c] = FEN ? E[0][c][k] + O[0][c][k] : E[0][c][k]; T1[c] = FEN ? E[1][c][k] + O[1][c][k] : E[1][c][k]; }
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const uint32_t s88 = T1[2 * j] + T1[2 * j + 1], top = T0[2 * j] + T0[2 * j + 1], left = T1[2 * j];
          const uint32_t k88 = s88 * M1 + b;
          kc[u][5 * j + 0] = k88;
          kc[u][5 * j + 1] = top * M1 + b;
          kc[u][5 * j + 2] = k88 - top * M1;
          kc[u][5 * j + 3] = left * M1 + b;
          kc[u][5 * j + 4] = k88 - left * M1;
        }
        const uint32_t all1 = (T1[0] + T1[1]) + (T1[2] + T1[3]);      // every row of the half: a 16x8 PU (2NxN), never sub-sampled
        const uint32_t all0 = (T0[0] + T0[1]) + (T0[2] + T0[3]);      // every row of the half's first strip
        const uint32_t e0 = (E[0][0][k] + E[0][1][k]) + (E[0][2][k] + E[0][3][k]);   // even rows of the first strip
        const uint32_t e1 = (E[1][0][k] + E[1][1][k]) + (E[1][2][k] + E[1][3][k]);   // even rows of the half
        if (h == 0) {
          kh[u][0] = all1 * M1 + b;                     // slot 1: 2NxN top
          kh[u][1] = all0 * M1 + b;                     // slot 5: 2NxnU top (16x4)
          R0[k] = e0; R1[k] = e1; C0[k] = E[1][0][k]; C01[k] = E[1][0][k] + E[1][1][k]; C3[k] = E[1][3][k];
        } else {
          kh[u][0] = all1 * M1 + b;                     // slot 2: 2NxN bottom
          kh[u][1] = (all1 - all0) * M1 + b;            // slot 8: 2NxnD bottom (16x4)
          const uint32_t tot = R1[k] + e1, k0 = tot * MS + b;
          const uint32_t left = C01[k] + E[1][0][k] + E[1][1][k], c0 = C0[k] + E[1][0][k], c3 = C3[k] + E[1][3][k];
          kf[u][0] = k0;                                // slot 0: 2Nx2N
          kf[u][1] = left * MS + b;                     // slot 3: Nx2N left
          kf[u][2] = k0 - left * MS;                    // slot 4: Nx2N right
          kf[u][3] = k0 - R0[k] * MS;                   // slot 6: 2NxnU bottom (16x12)
          kf[u][4] = (R1[k] + e0) * MS + b;             // slot 7: 2NxnD top (16x12)
          kf[u][5] = c0 * MS + b;                       // slot 9: nLx2N left
          kf[u][6] = k0 - c0 * MS;                      // slot 10
          kf[u][7] = k0 - c3 * MS;                      // slot 11: nRx2N left
          kf[u][8] = c3 * MS + b;                       // slot 12
        }
      }
      auto upd = [&](uint32_t& bst, const uint32_t a0, const uint32_t a1) { if constexpr (KP == 2) cu_min2(bst, a0, a1); else bst = min(bst, a0); };
#pragma unroll
      for (int i = 0; i < 10; i++) upd(best[CU_SLOTS + 10 * h + i], kc[0][i], kc[KP - 1][i]);
      if (h == 0) {
        upd(best[1], kh[0][0], kh[KP - 1][0]);
        upd(best[5], kh[0][1], kh[KP - 1][1]);
      } else {
        constexpr int FSLOT[9] = {0, 3, 4, 6, 7, 9, 10, 11, 12};
        upd(best[2], kh[0][0], kh[KP - 1][0]);
        upd(best[8], kh[0][1], kh[KP - 1][1]);
#pragma unroll
        for (int i = 0; i < 9; i++) upd(best[FSLOT[i]], kf[0][i], kf[KP - 1][i]);
      }
    }
  }
}

template <int S, bool FEN, bool CHILD>
__global__ void __launch_bounds__(CU8_THREADS, 1)
k_search8_cu(const S8Unit* __restrict__ units, const S8Bundle* __restrict__ bundles, unsigned long long* __restrict__ keys,
             DevPlane cur_plane, DevPlane ref_plane, const __grid_constant__ CUtensorMap ref_map) {
  typedef CuTraits<S, FEN> T;
  static_assert(!CHILD || S == 16, "only 16x16 CUs carry child CUs");
  constexpr int NSLOT = (S == 8) ? 5 : CHILD ? CU_SLOTS_ALL : CU_SLOTS;
  constexpr bool MASK_BY_KEY = S <= 32;                 // see cu_key_masked
  constexpr uint32_t KEY_NONE = MASK_BY_KEY ? cu_key_masked(S) : 0xffffffffu;   // best >= KEY_NONE: no candidate yet
  extern __shared__ __align__(128) uint8_t s8_smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ S8Bundle s_bd[CU8_WARPS];
  __shared__ __align__(16) uint32_t s_px[CU8_WARPS][CU_PX_BLOCKS * 16];   // per warp: lambda * bits of the x component of its CU's columns

  const S8Unit un = units[blockIdx.x];
  uint8_t* s_ref = s8_smem;
  uint8_t* s_org = s8_smem + un.org_smem_off;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // host: this unit's window has the geometry the launch's tensor map was built for.  The CHILD kernel keeps the row copies: with
  // the second staging path compiled in, its 14 k-instruction body ran 3 % slower (1304 -> 1343 us, profiles/r02_launches_final.txt)
  const bool use_map = !CHILD && (un.variant & 0x100) != 0;
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();                                      // the initialised barrier is visible before its first use
  if (threadIdx.x == 0) mbar_expect_tx(&s_bar, (uint32_t)(un.ref_pitch * un.ref_rows + un.org_pitch * un.org_rows));
  __syncthreads();                                      // ... and armed before any copy can complete on it
  {
    if (use_map) {
      // tensor-map TMA (SASS UTMALDG): the whole window is one 2-D box of the reference plane
      if (threadIdx.x == 0) tma_load_2d(s_ref, &ref_map, un.ref_bx + ref_plane.margin_x, un.ref_by + ref_plane.margin_y, &s_bar);
    } else {
      // every thread issues the bulk copies of its rows (one warp alone serialises ~250 UBLKCP issues: ncu showed a
      // fifth of the stall samples on the mbarrier spin)
      const uint8_t* gref = reinterpret_cast<const uint8_t*>(ref_plane.base) +
                            (size_t)(un.ref_by + ref_plane.margin_y) * ref_plane.pitch + (un.ref_bx + ref_plane.margin_x);
      for (int r = threadIdx.x; r < un.ref_rows; r += CU8_THREADS)
        bulk_g2s(s_ref + r * un.ref_pitch, gref + (size_t)r * ref_plane.pitch, (uint32_t)un.ref_pitch, &s_bar);
    }
    const uint8_t* gorg = reinterpret_cast<const uint8_t*>(cur_plane.base) +
                          (size_t)(un.org_by + cur_plane.margin_y) * cur_plane.pitch + (un.org_bx + cur_plane.margin_x);
    for (int r = (int)threadIdx.x - 256; r < un.org_rows; r += CU8_THREADS)
      if (r >= 0) bulk_g2s(s_org + r * un.org_pitch, gorg + (size_t)r * cur_plane.pitch, (uint32_t)un.org_pitch, &s_bar);
  }
  mbar_wait(&s_bar, 0);
  {
    // copies 1..3: the window shifted left by 1..3 bytes (the word past the last row belongs to the slack rows).  These cannot be
    // loaded: bulk and tensor copies start on 16-byte boundaries of global memory.
    const int nvec = (un.ref_pitch * un.ref_rows) >> 4;
    const int cs = un.copy_stride;
    for (int i = threadIdx.x; i < nvec; i += CU8_THREADS) {
      const uint4 a = reinterpret_cast<const uint4*>(s_ref)[i];
      const uint32_t e = reinterpret_cast<const uint32_t*>(s_ref)[4 * i + 4];
#pragma unroll
      for (int c = 1; c < 4; c++) {
        uint4 o;
        o.x = __funnelshift_r(a.x, a.y, 8 * c); o.y = __funnelshift_r(a.y, a.z, 8 * c);
        o.z = __funnelshift_r(a.z, a.w, 8 * c); o.w = __funnelshift_r(a.w, e, 8 * c);
        reinterpret_cast<uint4*>(s_ref + c * cs)[i] = o;
      }
    }
  }
  __syncthreads();

  int bslot = un.job_first;
  S8Bundle& bd = s_bd[warp];
  auto load_bundle = [&]() {
    __syncwarp();
    reinterpret_cast<int32_t*>(&bd)[lane] = reinterpret_cast<const int32_t*>(&bundles[bslot])[lane];            // 64 ints
    reinterpret_cast<int32_t*>(&bd)[lane + 32] = reinterpret_cast<const int32_t*>(&bundles[bslot])[lane + 32];
    __syncwarp();
    // x part of the MV cost of every column of the CU's window, once per CU and warp instead of four times per tile;
    // stored in the order the lanes read it: [block][lane & 3][k] = column block * 16 + (lane & 3) + 4 k
    if (bd.n_blk <= CU_PX_BLOCKS) {
      for (int i = lane; i < bd.n_blk * 16; i += 32) {
        const int col = (i & ~15) + ((i >> 2) & 3) + 4 * (i & 3) - bd.xal;
        s_px[warp][i] = bd.lambda * eg_bits(((bd.lt_x + col) << 2) - bd.pred_x);
      }
      __syncwarp();
    }
  };
  constexpr int LK = (S == 8) ? 4 : 2;                // bits of the within-tile candidate index
  constexpr int TILE_LIMIT = 1 << (CU_LOCAL_BITS - LK);   // tiles a lane may see between two flushes
  const int quad = lane >> 2, sub = lane & 3;
  const uint32_t s_copy = smem_u32(s_ref) + (uint32_t)(sub * un.copy_stride);   // this lane's byte phase (shared address)
  const uint32_t one = blockDim.x >> 9;               // 1, but not to the compiler: see CuKeys::add
  uint32_t best[NSLOT];
#pragma unroll
  for (int s = 0; s < NSLOT; s++) best[s] = 0xffffffffu;
  // every warp takes a contiguous run of the unit's items (round-robin dealing made every warp visit, and flush, every CU)
  const int per_warp = (un.item_last - un.item_first + CU8_WARPS - 1) / CU8_WARPS;
  const int w_first = un.item_first + warp * per_warp, w_last = min(w_first + per_warp, un.item_last);
  int first_item = w_first;                           // first item since the last flush (decodes local indices)
  int g = 0, blk = 0, g0 = 0;                         // row group / block of the lane's quad; g0: row group at the last flush (8x8 CUs)
  bool edge_mode = false;                             // the candidates since the last flush came from edge items (S >= 16, tile-numbered keys)
  // S >= 16, bd.edge == 2: ROW-INDEXED local keys, (candidate row << 3) | (4 for the last column, else k).  With eight blocks per
  // candidate row a lane keeps its block (blk = quad), so this orders every candidate a lane can see - block items and edge items
  // alike - in raster order, whatever order they are visited in: one running minimum, no flush between the two kinds of item.
  auto flush = [&]() {
    const bool rowkeys = S >= 16 && bd.edge == 2;
#pragma unroll
    for (int s = 0; s < NSLOT; s++) {
      // lanes number their candidates locally: reduce the cost first, then the raster index among the lanes that hold it
      const bool have = best[s] < KEY_NONE;
      const uint32_t cost = have ? (best[s] >> CU_LOCAL_BITS) : 0xffffffffu;
      const uint32_t cmin = __reduce_min_sync(0xffffffffu, cost);
      uint32_t idx = 0xffffffffu;
      if (have && cost == cmin) {
        const uint32_t local = best[s] & ((1u << CU_LOCAL_BITS) - 1u);
        int cyi, cxi;
        if constexpr (S == 8) {      // (candidate row since g0, tile rank inside the row group, column): see the tile loop
          const int rb = bd.rank_bits, rowrel = (int)(local >> (2 + rb)), rank = (int)(local >> 2) & ((1 << rb) - 1);
          const int gg = g0 + (rowrel >> 2), bb = ((quad - gg * bd.n_blk) & 7) + 8 * rank;
          cyi = bd.cy_first + gg * T::KY + (rowrel & 3); cxi = bb * 16 + sub + 4 * (int)(local & 3u) - bd.xal;
        } else if (rowkeys) {
          cyi = bd.cy_first + (int)(local >> 3);
          cxi = (local & 4u) ? bd.nx - 1 : quad * 16 + sub + 4 * (int)(local & 3u) - bd.xal;
        } else if (edge_mode) {      // edge item number since the last flush; the lane is the candidate row, the column is the last one
          cyi = bd.cy_first + (first_item + (int)(local >> LK) - bd.item_start - bd.n_main) * 32 + lane; cxi = bd.nx - 1;
        } else {                     // tile number since the last flush, column
          const int q = (first_item + (int)(local >> LK) - bd.item_start) * 8 + quad;
          const int gg = q / bd.n_blk, bb = q - gg * bd.n_blk;
          cyi = bd.cy_first + gg; cxi = bb * 16 + sub + 4 * (int)(local & 3u) - bd.xal;
        }
        idx = (uint32_t)(cyi * bd.nx + cxi);
      }
      const uint32_t imin = __reduce_min_sync(0xffffffffu, idx);
      const int out = (s < CU_SLOTS) ? bd.out_idx[s < CU_SLOTS ? s : 0] : bd.child_idx[s < CU_SLOTS ? 0 : s - CU_SLOTS];
      if (lane == 0 && cmin != 0xffffffffu && out >= 0) atomicMin(&keys[out], make_key(cmin, imin));
      best[s] = 0xffffffffu;
    }
  };
  // (row group, block) of this lane's quad in the current item; the next item is 8 quads further
  auto locate = [&](int item) {
    const int q = (item - bd.item_start) * 8 + quad;
    g = q / bd.n_blk; blk = q - g * bd.n_blk; g0 = g;
  };
  const uint32_t c16 = one << 16;                                       // x >> 16 as umulhi(x, 2^16): IMAD.HI, not SHF
  const uint32_t rpitch = (uint32_t)un.ref_pitch, opitch = (uint32_t)un.org_pitch;

  // ---- one block item: 8 quads x 16 candidate columns x KY candidate rows -------------------------------------------------
  auto block_item = [&](int item, bool rowkeys) {
    // Local index of a candidate inside the key: must grow in raster order along the candidates ONE LANE sees between
    // two flushes.  S >= 16 (one candidate row per tile): the lane's tiles come in raster order, so the tile number does
    // (or the row itself, see rowkeys).  S == 8 (four rows per tile): the lane may see two tiles of one row group (8 blocks
    // apart), so the index is (candidate row since g0, rank of the tile inside its row group = blk / 8, column).
    const uint32_t rowunit = 1u << (2 + bd.rank_bits);                  // S == 8: index step of one candidate row
    const uint32_t tile_local = (S == 8) ? (uint32_t)((g - g0) * 4) * rowunit + ((uint32_t)(blk >> 3) << 2)
                                         : rowkeys ? (uint32_t)g << 3 : (uint32_t)(item - first_item) << LK;
    const uint32_t orgp = smem_u32(s_org) + (uint32_t)bd.org_off;
    if (g < bd.n_rowgroups) {
    const int cyi0 = g * T::KY;
    const int cx0 = blk * 16 + sub - bd.xal;                            // window column of candidate k = 0 (may be < 0)
    const uint32_t refp = s_copy + (uint32_t)(bd.win_off - bd.xal + cyi0 * un.ref_pitch + blk * 16);     // 16-byte aligned
    // key base of candidate (jy, k) = MV cost * mk[k] + jy * ru[k] + idx0[k]: IMADs on the FMA pipe.  Block columns outside
    // the window take no part: their mk and ru are 0 and idx0 is KEY_NONE (MASK_BY_KEY), so the base is exactly KEY_NONE.
    uint32_t px[4], mk[4], idx0[4], ru[4];
    bool valid[4];
    if (bd.n_blk <= CU_PX_BLOCKS) {
      const uint4 v = *reinterpret_cast<const uint4*>(&s_px[warp][blk * 16 + sub * 4]);
      px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) px[k] = bd.lambda * eg_bits(((bd.lt_x + cx0 + 4 * k) << 2) - bd.pred_x);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      valid[k] = (unsigned)(cx0 + 4 * k) < (unsigned)bd.nx;
      const bool on = valid[k] || !MASK_BY_KEY;
      mk[k] = on ? (1u << CU_LOCAL_BITS) : 0u;
      idx0[k] = on ? (tile_local | (uint32_t)k) : KEY_NONE;
      ru[k] = on ? rowunit : 0u;
    }

    if constexpr (S == 8) {
      // four candidate rows per tile, whole CU (8 rows x 2 words) in registers; per candidate: Q[0][c] = rows 0..3 of
      // column half c, Q[1][c] = rows 0..7 (the lower half continues the upper half's accumulators)
      uint32_t o[8][2];
#pragma unroll
      for (int r = 0; r < 8; r++) cu_load_org<2>(cu_row_addr(orgp, r, opitch), o[r]);
      uint32_t Q[T::KY][2][2][4];
#pragma unroll
      for (int a = 0; a < T::KY; a++)
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
          for (int k = 0; k < 4; k++) Q[a][0][c][k] = 0;
#pragma unroll
      for (int r = 0; r < 8 + T::KY - 1; r++) {
        uint32_t w[5];
        cu_load_ref<5>(cu_row_addr(refp, r, rpitch), w);
#pragma unroll
        for (int jy = 0; jy < T::KY; jy++) {
          const int orow = r - jy;
          if (orow >= 0 && orow < 8) {
#pragma unroll
            for (int j = 0; j < 5; j++)
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const int i = j - k;
                if (i >= 0 && i < 2) {
                  const uint32_t addend = (orow == 4) ? Q[jy][0][i][k] : Q[jy][orow >> 2][i][k];
                  Q[jy][orow >> 2][i][k] = sad4_acc(w[j], o[orow][i], addend);
                }
              }
          }
        }
      }
#pragma unroll
      for (int jy = 0; jy < T::KY; jy++) {
        const int cyi = bd.cy_first + cyi0 + jy;
        if (cyi0 + jy < bd.ny) {
          const uint32_t py = bd.lambda * eg_bits(((bd.lt_y + cyi) << 2) - bd.pred_y);
          constexpr uint32_t M1 = 1u << CU_LOCAL_BITS;
          uint32_t key[4][5];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const uint32_t base = __umulhi(px[k] + py, c16) * mk[k] + ((uint32_t)jy * ru[k] + idx0[k]);
            const uint32_t t = Q[jy][0][0][k] * one + Q[jy][0][1][k], l = Q[jy][1][0][k];
            const uint32_t k0 = Q[jy][1][1][k] * M1 + (l * M1 + base);
            key[k][0] = k0;                       // 8x8  (no PU of an 8x8 CU has more than 8 rows: iSubShift = 0)
            key[k][1] = t * M1 + base;            // 8x4 top
            key[k][2] = k0 - t * M1;              // 8x4 bottom = whole - top
            key[k][3] = l * M1 + base;            // 4x8 left
            key[k][4] = k0 - l * M1;              // 4x8 right = whole - left
          }
#pragma unroll
          for (int s = 0; s < 5; s++) { cu_min2(best[s], key[0][s], key[1][s]); cu_min2(best[s], key[2][s], key[3][s]); }
        }
      }
    } else {
      const uint32_t py = bd.lambda * eg_bits(((bd.lt_y + bd.cy_first + cyi0) << 2) - bd.pred_y);
      uint32_t base[4];
#pragma unroll
      for (int k = 0; k < 4; k++) base[k] = __umulhi(px[k] + py, c16) * mk[k] + idx0[k];
      if constexpr (CHILD) cu16_child_tile<FEN, 4, NSLOT>(refp, orgp, rpitch, opitch, base, best);
      else                 cu_tile<S, FEN, 4, MASK_BY_KEY, NSLOT>(refp, orgp, rpitch, opitch, base, valid, one, best);
    }
    }   // row group in range
    blk += bd.step_blk; g += bd.step_g;                                  // 8 quads further
    if (blk >= bd.n_blk) { blk -= bd.n_blk; g++; }
  };

  // ---- one edge item: the window's last column, one candidate row per lane (32 rows per item) ---------------------------------
  // (xal + nx - 1) % 16 == 0, so the column starts a 16-byte block of copy 0 and every lane reads aligned words of its own row.
  auto edge_item = [&](int item, bool rowkeys) {
    if constexpr (S >= 16) {
      const int row = (item - bd.item_start - bd.n_main) * 32 + lane;
      const bool ok = row < bd.ny;
      const int rowc = min(row, bd.ny - 1);
      const uint32_t orgp = smem_u32(s_org) + (uint32_t)bd.org_off;
      const uint32_t refp = smem_u32(s_ref) + (uint32_t)(bd.win_off + rowc * un.ref_pitch + bd.nx - 1);
      const uint32_t px = bd.lambda * eg_bits(((bd.lt_x + bd.nx - 1) << 2) - bd.pred_x);
      const uint32_t py = bd.lambda * eg_bits(((bd.lt_y + bd.cy_first + rowc) << 2) - bd.pred_y);
      const uint32_t tile_local = rowkeys ? ((uint32_t)rowc << 3) | 4u : (uint32_t)(item - first_item) << LK;
      uint32_t base[1];
      bool valid[1] = {ok};
      base[0] = (ok || !MASK_BY_KEY) ? (__umulhi(px + py, c16) << CU_LOCAL_BITS) + tile_local : KEY_NONE;
      if constexpr (CHILD) cu16_child_tile<FEN, 1, NSLOT>(refp, orgp, rpitch, opitch, base, best);
      else                 cu_tile<S, FEN, 1, MASK_BY_KEY, NSLOT>(refp, orgp, rpitch, opitch, base, valid, one, best);
    }
  };

  if (w_first < w_last) {
    // the CU that holds this warp's first item: 32 descriptors per probe instead of a walk of dependent loads
    for (;; bslot += 32) {
      const int j = bslot + lane;
      const bool past = j >= un.job_first + un.job_count || w_first < bundles[j].item_start + bundles[j].n_items;
      const unsigned m = __ballot_sync(0xffffffffu, past);
      if (m) { bslot += __ffs(m) - 1; break; }
    }
    load_bundle();
    int item = w_first;
    for (;;) {                                          // one pass per CU of this warp's run: its block items, then its edge items
      const bool rowkeys = S >= 16 && bd.edge == 2;
      const int main_end = min(w_last, bd.item_start + bd.n_main), cu_end = min(w_last, bd.item_start + bd.n_items);
      bool dirty = false;                               // candidates of this CU are waiting in `best` (tile-numbered keys)
      if (item < main_end) {
        first_item = item;
        locate(item);
        edge_mode = false;
        dirty = true;
        for (; item < main_end; item++) {
          if (!rowkeys && (S == 8 ? __any_sync(0xffffffffu, (g - g0) >= (1 << (CU_LOCAL_BITS - 4 - bd.rank_bits))) : (item - first_item >= TILE_LIMIT))) {
            flush();                                    // the local index would run out of bits
            first_item = item; g0 = g;
          }
          block_item(item, rowkeys);
        }
      }
      if (S >= 16 && item < cu_end) {
        if (!rowkeys && dirty) flush();                 // tile-numbered keys: the last column's candidates precede later rows' in raster
        first_item = item;                              // order, so a lane's local order holds only inside one kind of item
        edge_mode = true;
        for (; item < cu_end; item++) {
          if (!rowkeys && item - first_item >= TILE_LIMIT) { flush(); first_item = item; }
          edge_item(item, rowkeys);
        }
      }
      flush();                                          // this warp's candidates of the CU: one atomicMin per PU
      if (item >= w_last) break;
      do { bslot++; } while (item >= bundles[bslot].item_start + bundles[bslot].n_items);
      load_bundle();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// CUV_16C_*: 16x16 CUs that carry their four 8x8 child CUs (8-bit planes only; the 16-bit tables end at CUV_BASE_COUNT)
enum : int { CUV_8 = 0, CUV_16_F0, CUV_16_F1, CUV_32_F0, CUV_32_F1, CUV_64_F0, CUV_64_F1, CUV_BASE_COUNT, CUV_16C_F0 = CUV_BASE_COUNT, CUV_16C_F1, CUV_COUNT };
constexpr int CUV_MAX = 2 * CUV_BASE_COUNT + 4;  // schedule tables: 16-bit planes number their aligned variants 7..13 and their child variants 14..17 (hmb200_search16_cu.cuh)
constexpr int CU16_CHILD_STATIC_SMEM = 24 * 1024;  // k_search16_cu<16,.,.,CHILD>: 20 KB of per-thread child minima + descriptors (two CTAs per SM must still fit)
typedef void (*S8CuKernel)(const S8Unit*, const S8Bundle*, unsigned long long*, DevPlane, DevPlane, const CUtensorMap);
inline const S8CuKernel* search8_cu_kernels() {
  static const S8CuKernel table[CUV_COUNT] = { k_search8_cu<8, false, false>, k_search8_cu<16, false, false>, k_search8_cu<16, true, false>,
                                               k_search8_cu<32, false, false>, k_search8_cu<32, true, false>, k_search8_cu<64, false, false>,
                                               k_search8_cu<64, true, false>, k_search8_cu<16, false, true>, k_search8_cu<16, true, true> };
  return table;
}
inline int cu_variant(int S, bool fen, bool child = false) {
  if (child) return fen ? CUV_16C_F1 : CUV_16C_F0;
  return S == 8 ? CUV_8 : S == 16 ? (fen ? CUV_16_F1 : CUV_16_F0) : S == 32 ? (fen ? CUV_32_F1 : CUV_32_F0) : (fen ? CUV_64_F1 : CUV_64_F0);
}

struct CuSchedule {
  int n_units = 0, n_bundles = 0;
  unsigned long long executed_abs_diffs = 0;     // byte abs-diffs the fused kernels execute (incl. overlapped last blocks)
  unsigned long long unique_abs_diffs = 0;       // ... they cannot avoid: every visited sample of every CU once per candidate (no masked lanes)
  unsigned long long fused_tasks = 0;
  int bps = 1;                                   // bytes per sample of the planes this schedule was built for
  S8Unit* d_units = nullptr;
  S8Bundle* d_bundles = nullptr;
  int unit_first[CUV_MAX] = {0}, unit_count[CUV_MAX] = {0}, smem_of[CUV_MAX] = {0};
  // 8-bit planes: the units of a variant whose window has the variant's most common geometry are staged through a tensor map
  // (box tma_w bytes x tma_h rows); tma_units of the variant's units take that path (0: none)
  int tma_w[CUV_MAX] = {0}, tma_h[CUV_MAX] = {0}, tma_units[CUV_MAX] = {0};
  S8Box rbox{0, 0, 0, 0}, obox{0, 0, 0, 0};
};

inline void cu_free_schedule(CuSchedule* s) {
  if (s->d_units) cudaFree(s->d_units);
  if (s->d_bundles) cudaFree(s->d_bundles);
  *s = CuSchedule();
}

inline int cu_configure(std::string* err) {
  const S8CuKernel* k = search8_cu_kernels();
  for (int v = 0; v < CUV_COUNT; v++) {
    cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(k[v]), cudaFuncAttributeMaxDynamicSharedMemorySize, S8_SMEM_MAX);
    if (e != cudaSuccess) { if (err) *err = std::string("cudaFuncSetAttribute(k_search8_cu): ") + cudaGetErrorString(e); return HMB200_ERR_CUDA; }
  }
  return HMB200_OK;
}

// geometry of the lane layout per sample size: 8-bit: quads of lanes own 16 columns, 8 quads per warp-item;
// 16-bit: pairs of lanes own 8 columns, 16 pairs per warp-item (hmb200_search16_cu.cuh)
struct CuGeom {
  int bps, blkw, groups, warps, smem_group;     // smem_group: dynamic shared memory a group may use
  int ky(int S) const { return bps == 1 ? cu_ky(S) : (S == 8 ? 2 : 1); }
  int lk(int S) const { return S == 8 ? 4 : 2; }
};
inline CuGeom cu_geom(int bps) { return bps == 1 ? CuGeom{1, 16, 8, CU8_WARPS, S8_SMEM_MAX} : CuGeom{2, 8, 16, S8_WARPS, S8_SMEM_SHARED2}; }

// Dynamic shared memory of a CU-fused group.  8-bit planes: four byte-shifted copies of the window (see cu_load_ref),
// copy_stride = 32 mod 128 so that the four lanes of a quad (one copy each) and the next quad (16 bytes further) hit
// disjoint banks; a few slack rows behind the window for masked candidates; then the original tile.
constexpr int CU8_SLACK_ROWS = 4;
// 16-bit planes, CUs up to 16x16 (cu16_two_phase): the staged window plus a second copy shifted by one sample.
inline bool cu16_two_phase(int S) { return S <= 16; }
inline int cu_smem_need(const S8Box& rb, const S8Box& ob, int* org_off, int* copy_stride, int bps, bool two_phase = false) {
  if (bps != 1 && !two_phase) { if (copy_stride) *copy_stride = 0; return s8_smem_need(rb, ob, org_off, bps); }
  if (bps != 1) {
    // hmb200_search16_cu.cuh: lanes whose candidates start at an odd sample read aligned words from the second copy.
    // Its offset is 8 bytes past a multiple of 128 so that the two lanes of a pair land in different shared-memory banks.
    const int rp = (s8_ce(rb.x1) - s8_fl(rb.x0)) * bps, rr = rb.y1 - rb.y0 + S8_SLACK_ROWS;
    const int op = (s8_ce(ob.x1) - s8_fl(ob.x0)) * bps, orr = ob.y1 - ob.y0;
    const int cs = (((rp * rr + 16) + 127) & ~127) + 8;
    const int ro = (2 * cs + 127) & ~127;
    if (copy_stride) *copy_stride = cs;
    if (org_off) *org_off = ro;
    return ro + op * orr + 16;
  }
  const int rp = s8_ce(rb.x1) - s8_fl(rb.x0), rr = rb.y1 - rb.y0 + CU8_SLACK_ROWS;
  const int op = s8_ce(ob.x1) - s8_fl(ob.x0), orr = ob.y1 - ob.y0;
  const int cs = ((rp * rr + 16 + 127) & ~127) + 32;
  const int ro = (4 * cs + 127) & ~127;
  if (copy_stride) *copy_stride = cs;
  if (org_off) *org_off = ro;
  return ro + op * orr + 16;
}

// Finds CU bundles: PUs that are partitions of the same aligned S x S CU and share window, predictor and lambda.
// taken[i] = 1 for every bundled task (the per-PU schedule skips those).
struct CuBundleHost {
  int S; bool fen; int cu_x, cu_y; int slot_task[CU_SLOTS]; int first_task;
  bool child = false;                       // 16x16 CU that carries its 8x8 child CUs (8-bit planes)
  int child_task[CU_CHILD_SLOTS];           // task per child PU: 5 * (2 * row + column) + {8x8, 8x4 top, 8x4 bottom, 4x8 left, 4x8 right}
};
inline void cu_extract_bundles(const std::vector<SearchTask>& tasks, int bps, std::vector<char>& taken, std::vector<CuBundleHost>& out) {
  const CuGeom G = cu_geom(bps);
  struct Key {
    int v[10];
    bool operator<(const Key& o) const { for (int i = 0; i < 10; i++) if (v[i] != o.v[i]) return v[i] < o.v[i]; return false; }
  };
  std::map<Key, CuBundleHost> found;
  for (int i = 0; i < (int)tasks.size(); i++) {
    const SearchTask& t = tasks[i];
    const int S = std::max(t.w, t.h);
    if (!(S == 8 || S == 16 || S == 32 || S == 64) || t.ref_x != t.org_x || t.ref_y != t.org_y || t.org_x < 0 || t.org_y < 0) continue;
    const int nx = t.rb_x - t.lt_x + 1, ny = t.rb_y - t.lt_y + 1;
    if (nx < (bps == 1 ? 1 : G.blkw) || ny < 1) continue;    // 8-bit blocks are masked per column; 16-bit blocks overlap
    const int cx = t.org_x - t.org_x % S, cy = t.org_y - t.org_y % S;
    // even a one-row-group slice of the window must fit in shared memory, and the per-lane local candidate index of
    // a slice that fills it must fit CU_LOCAL_BITS (see cu_min)
    if (cu_smem_need(S8Box{cx + t.lt_x, cy + t.lt_y, cx + t.rb_x + S, cy + t.lt_y + G.ky(S) - 1 + S}, S8Box{cx, cy, cx + S, cy + S}, nullptr, nullptr, bps,
                     bps != 1 && cu16_two_phase(S)) > S8_SMEM_MAX) continue;
    {
      const int n_blk = (nx + (bps == 1 ? 15 : 0) + G.blkw - 1) / G.blkw, nrg = (ny + G.ky(S) - 1) / G.ky(S);   // worst alignment
      const int n_items = (n_blk * nrg + G.groups - 1) / G.groups;
      if (bps != 1 && n_items / G.warps + 2 >= (1 << (CU_LOCAL_BITS - G.lk(S)))) continue;   // 8-bit kernels flush when the index runs out
    }
    int slot = -1;
    for (int s = 0; s < (S == 8 ? 5 : CU_SLOTS); s++)
      if (cu_slot_x(S, s) == t.org_x - cx && cu_slot_y(S, s) == t.org_y - cy && cu_slot_w(S, s) == t.w && cu_slot_h(S, s) == t.h) { slot = s; break; }
    if (slot < 0) continue;
    Key k = {{cx, cy, S, t.lt_x, t.lt_y, t.rb_x, t.rb_y, t.pred_x, t.pred_y, (int)t.lambda_cost}};
    auto it = found.find(k);
    if (it == found.end()) {
      CuBundleHost b; b.S = S; b.fen = false; b.cu_x = cx; b.cu_y = cy; b.first_task = i;
      for (int s = 0; s < CU_SLOTS; s++) b.slot_task[s] = -1;
      it = found.insert(std::make_pair(k, b)).first;
    }
    if (it->second.slot_task[slot] >= 0) continue;            // the same PU twice: the second one stays a plain job
    it->second.slot_task[slot] = i;
  }
  taken.assign(tasks.size(), 0);
  std::map<Key, int> where;                 // bundle key -> index in out (to find a 16x16 CU's children)
  std::vector<Key> key_of;
  for (auto& kv : found) {
    CuBundleHost& b = kv.second;
    int n = 0, fen_seen = -1; bool consistent = true;
    for (int s = 0; s < CU_SLOTS; s++) {
      if (b.slot_task[s] < 0) continue;
      n++;
      const SearchTask& t = tasks[b.slot_task[s]];
      if (t.h > 8) { if (fen_seen < 0) fen_seen = t.sub_shift; else if (fen_seen != t.sub_shift) consistent = false; }
      else if (t.sub_shift != 0) consistent = false;
    }
    if (n < 2 || !consistent) continue;
    b.fen = fen_seen == 1;
    for (int s = 0; s < CU_CHILD_SLOTS; s++) b.child_task[s] = -1;
    for (int s = 0; s < CU_SLOTS; s++) if (b.slot_task[s] >= 0) taken[b.slot_task[s]] = 1;
    where[kv.first] = (int)out.size();
    key_of.push_back(kv.first);
    out.push_back(b);
  }
  // 8-bit planes: a 16x16 CU takes the PUs of its four 8x8 child CUs along when they share its window (as displacements),
  // predictor and lambda - the child SADs are sums of the 4x4 cells the 16x16 pass computes anyway (hmb200 CHILD kernels).
  std::vector<char> absorbed(out.size(), 0);
  if (!getenv("HMB200_NO_CHILD_FOLD") && (bps == 1 || !getenv("HMB200_NO_CHILD_FOLD16"))) {
    for (size_t i = 0; i < out.size(); i++) {
      if (out[i].S != 16) continue;
      for (int c = 0; c < 4; c++) {
        Key k = key_of[i];
        k.v[0] = out[i].cu_x + 8 * (c & 1); k.v[1] = out[i].cu_y + 8 * (c >> 1); k.v[2] = 8;
        auto it = where.find(k);
        if (it == where.end() || absorbed[it->second]) continue;
        const CuBundleHost& ch = out[it->second];
        for (int s = 0; s < 5; s++) out[i].child_task[5 * c + s] = ch.slot_task[s];
        out[i].child = true;
        absorbed[it->second] = 1;
      }
    }
    std::vector<CuBundleHost> kept;
    for (size_t i = 0; i < out.size(); i++) if (!absorbed[i]) kept.push_back(out[i]);
    out.swap(kept);
  }
  std::sort(out.begin(), out.end(), [](const CuBundleHost& a, const CuBundleHost& b) { return a.first_task < b.first_task; });
}

inline bool cu_build_schedule(const std::vector<SearchTask>& tasks, const std::vector<CuBundleHost>& hb, int bps, int bit_depth,
                              int sm_count, cudaStream_t stream, CuSchedule* out, std::string* err) {
  if (hb.empty()) return true;
  CuGeom GM = cu_geom(bps);
  out->bps = bps;
  if (bps == 2)                                       // the child kernels keep 23 KB of static shared memory; two CTAs per SM must still fit
    for (const CuBundleHost& b : hb) if (b.child) { GM.smem_group -= CU16_CHILD_STATIC_SMEM; break; }
  // entity = a bundle, or a horizontal slice of its window when the whole window does not let two CTAs share an SM
  struct Ent { int b, cy_first, ny; S8Box rb, ob; bool two; };
  std::vector<Ent> ents;
  auto any_task = [&](const CuBundleHost& b) -> const SearchTask& {
    for (int s = 0; s < CU_SLOTS; s++) if (b.slot_task[s] >= 0) return tasks[b.slot_task[s]];
    return tasks[0];
  };
  for (size_t i = 0; i < hb.size(); i++) {
    const CuBundleHost& b = hb[i];
    const SearchTask& t = any_task(b);
    const int ny = t.rb_y - t.lt_y + 1, ky = GM.ky(b.S);
    const S8Box ob{b.cu_x, b.cu_y, b.cu_x + b.S, b.cu_y + b.S};
    auto rbox_of = [&](int r0, int n) { return S8Box{b.cu_x + t.lt_x, b.cu_y + t.lt_y + r0, b.cu_x + t.rb_x + b.S, b.cu_y + t.lt_y + r0 + n - 1 + b.S}; };
    const bool two = bps != 1 && cu16_two_phase(b.S);
    int parts = 1, rows = ny;
    while (cu_smem_need(rbox_of(0, rows), ob, nullptr, nullptr, bps, two) > GM.smem_group && rows > ky) {
      parts++;
      rows = (((ny + parts - 1) / parts + ky - 1) / ky) * ky;
    }
    for (int r0 = 0; r0 < ny; r0 += rows) {
      const int n = std::min(rows, ny - r0);
      ents.push_back(Ent{(int)i, r0, n, rbox_of(r0, n), ob, two});
    }
  }
  std::stable_sort(ents.begin(), ents.end(), [&](const Ent& a, const Ent& b) {
    const int ka = a.ob.y0 >> 6, kb = b.ob.y0 >> 6;
    if (ka != kb) return ka < kb;
    if ((a.ob.x0 >> 6) != (b.ob.x0 >> 6)) return (a.ob.x0 >> 6) < (b.ob.x0 >> 6);
    return a.two < b.two;                                  // groups hold one shared-memory layout
  });
  struct Group { int first, count; S8Box rb, ob; bool two; };
  std::vector<Group> groups;
  for (size_t p = 0; p < ents.size(); p++) {
    if (!groups.empty()) {
      Group& g = groups.back();
      const S8Box nr = s8_union(g.rb, ents[p].rb), no = s8_union(g.ob, ents[p].ob);
      if (g.count < 4096 && g.two == ents[p].two && no.x1 - no.x0 <= 128 && no.y1 - no.y0 <= 128 &&
          cu_smem_need(nr, no, nullptr, nullptr, bps, g.two) <= GM.smem_group) {
        g.rb = nr; g.ob = no; g.count++;
        continue;
      }
    }
    if (cu_smem_need(ents[p].rb, ents[p].ob, nullptr, nullptr, bps, ents[p].two) > S8_SMEM_MAX) { if (err) *err = "cu_build_schedule: window too large"; return false; }
    groups.push_back(Group{(int)p, 1, ents[p].rb, ents[p].ob, ents[p].two});
  }
  // 8-bit planes: variants 0..8 (7, 8 = 16x16 CUs with children); 16-bit planes: 0..6 and 7..13 = the same with aligned 64-bit loads
  const int S_of_variant[CUV_MAX] = {8, 16, 16, 32, 32, 64, 64, bps == 1 ? 16 : 8, 16, 16, 32, 32, 64, 64, 16, 16, 16, 16};
  const bool F_of_variant[CUV_MAX] = {false, false, true, false, true, false, true, bps == 1 ? false : false, bps == 1 ? true : false, true, false, true, false, true,
                                      false, true, false, true};
  const bool use_a8 = bps == 2 && !getenv("HMB200_NO_LDS64");
  // 16-bit planes: the unit can read its window rows with 64-bit loads when candidate column 0 sits on an 8-byte boundary of the
  // staged window (whose origin is 16-sample aligned): (cu_x + lt_x) % 4 == 0
  auto variant_of = [&](const CuBundleHost& b, const SearchTask& t) {
    if (bps == 1) return cu_variant(b.S, b.fen, b.child);
    const bool a8 = use_a8 && ((b.cu_x + t.lt_x) & 3) == 0;
    if (b.child) return 2 * CUV_BASE_COUNT + (b.fen ? 1 : 0) + (a8 ? 2 : 0);
    return cu_variant(b.S, b.fen, false) + (a8 ? CUV_BASE_COUNT : 0);
  };
  auto variant_is_child = [&](int v) { return bps == 1 ? v >= CUV_BASE_COUNT : v >= 2 * CUV_BASE_COUNT; };
  auto rows_visited = [](int S, bool fen) { return (fen && S >= 16) ? (S == 16 ? 16 : S == 32 ? 24 : 32) : S; };
  auto item_cost = [&](int S, bool fen, bool child = false) -> long long {
    return (long long)GM.ky(S) * (rows_visited(S, fen) * (S / 4) * (bps == 1 ? 1 : 6) + (child ? 100 : 40));
  };
  auto vcost = [&](int v) { return item_cost(S_of_variant[v], F_of_variant[v], variant_is_child(v)) * (bps == 2 && variant_is_child(v) ? 2 : 1); };
  const bool use_edge = bps == 1 && !getenv("HMB200_NO_EDGE_ITEMS");
  std::vector<S8Bundle> bundles; bundles.reserve(ents.size());
  std::vector<int> bvar; bvar.reserve(ents.size());
  std::vector<std::pair<int, int> > group_range(groups.size());
  long long variant_cost[CUV_MAX] = {0};
  S8Box all_r{1 << 30, 1 << 30, -(1 << 30), -(1 << 30)}, all_o = all_r;
  for (size_t gi = 0; gi < groups.size(); gi++) {
    Group& g = groups[gi];
    std::vector<int> ids;
    for (int k = 0; k < g.count; k++) ids.push_back(g.first + k);
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) {
      return variant_of(hb[ents[a].b], any_task(hb[ents[a].b])) > variant_of(hb[ents[b].b], any_task(hb[ents[b].b]));
    });
    const int rx0 = s8_fl(g.rb.x0), ox0 = s8_fl(g.ob.x0);
    const int rpitch = (s8_ce(g.rb.x1) - rx0) * bps, opitch = (s8_ce(g.ob.x1) - ox0) * bps;      // bytes
    group_range[gi] = std::make_pair((int)bundles.size(), g.count);
    int item = 0;
    for (int id : ids) {
      const Ent& e = ents[id];
      const CuBundleHost& b = hb[e.b];
      const SearchTask& t = any_task(b);
      S8Bundle d{};
      d.org_off = (b.cu_y - g.ob.y0) * opitch + (b.cu_x - ox0) * bps;
      d.win_off = (b.cu_y + t.lt_y + e.cy_first - g.rb.y0) * rpitch + (b.cu_x + t.lt_x - rx0) * bps;
      d.nx = t.rb_x - t.lt_x + 1; d.ny = e.ny; d.cy_first = e.cy_first; d.shr = bit_depth - 8;
      d.lt_x = t.lt_x; d.lt_y = t.lt_y; d.pred_x = t.pred_x; d.pred_y = t.pred_y; d.lambda = t.lambda_cost;
      d.xal = (bps == 1) ? (d.win_off & 15) : 0;           // window pitch and origin are multiples of 16 bytes
      d.n_blk = (d.xal + d.nx + GM.blkw - 1) / GM.blkw;
      // 129 = 8 * 16 + 1: a last block with a single column in it runs as edge items (a lane per candidate row) instead
      d.edge = (use_edge && b.S >= 16 && d.n_blk >= 2 && ((d.xal + d.nx) & 15) == 1) ? 1 : 0;
      if (d.edge) d.n_blk--;
      // eight blocks per candidate row and at most 256 rows: row-indexed argmin keys (see k_search8_cu), no flush between the
      // block items and the edge items of a CU
      if (d.edge && d.n_blk == 8 && d.ny <= 256 && !getenv("HMB200_NO_ROW_KEYS")) d.edge = 2;
      d.step_g = GM.groups / d.n_blk; d.step_blk = GM.groups % d.n_blk;      // 8-bit kernels: consecutive items per warp
      d.rank_bits = 0;
      while ((8 << d.rank_bits) < d.n_blk) d.rank_bits++;
      d.n_rowgroups = (d.ny + GM.ky(b.S) - 1) / GM.ky(b.S);
      d.n_main = (d.n_blk * d.n_rowgroups + GM.groups - 1) / GM.groups;
      const int n_edge = d.edge ? (d.ny + 31) / 32 : 0;
      d.n_items = d.n_main + n_edge;
      d.item_start = item; item += d.n_items;
      for (int s = 0; s < CU_SLOTS; s++) d.out_idx[s] = b.slot_task[s];
      for (int s = 0; s < CU_CHILD_SLOTS; s++) d.child_idx[s] = b.child ? b.child_task[s] : -1;
      const int v = variant_of(b, t);
      const int rows_v = b.child ? 16 : rows_visited(b.S, b.fen);         // child tiles visit every row
      variant_cost[v] += vcost(v) * d.n_main + (vcost(v) / 3) * n_edge;
      out->executed_abs_diffs += ((unsigned long long)d.n_blk * GM.blkw * (unsigned long long)(d.n_rowgroups * GM.ky(b.S)) + (unsigned long long)n_edge * 32) *
                                 rows_v * b.S;
      out->unique_abs_diffs += (unsigned long long)d.nx * d.ny * rows_v * b.S;
      if (e.cy_first == 0) {
        for (int s2 = 0; s2 < CU_SLOTS; s2++) if (b.slot_task[s2] >= 0) out->fused_tasks++;
        if (b.child) for (int s2 = 0; s2 < CU_CHILD_SLOTS; s2++) if (b.child_task[s2] >= 0) out->fused_tasks++;
      }
      bundles.push_back(d); bvar.push_back(v);
    }
    all_r = s8_union(all_r, g.rb); all_o = s8_union(all_o, g.ob);
  }
  long long target[CUV_MAX];
  int per_slot = bps == 1 ? 3 : 6;                     // units per resident CTA slot and variant (knob: HMB200_UNITS_PER_SLOT); 8-bit: one CTA per SM
  if (const char* e = getenv("HMB200_UNITS_PER_SLOT")) per_slot = std::max(1, atoi(e));
  for (int v = 0; v < CUV_MAX; v++) target[v] = std::max<long long>(variant_cost[v] / std::max(1, sm_count * (bps == 1 ? 1 : 2) * per_slot), 4000);
  std::vector<S8Unit> units;
  for (size_t gi = 0; gi < groups.size(); gi++) {
    const Group& g = groups[gi];
    S8Unit u{};
    const int rx0 = s8_fl(g.rb.x0), ox0 = s8_fl(g.ob.x0);
    u.ref_bx = rx0; u.ref_by = g.rb.y0; u.ref_pitch = (s8_ce(g.rb.x1) - rx0) * bps; u.ref_rows = g.rb.y1 - g.rb.y0;
    u.org_bx = ox0; u.org_by = g.ob.y0; u.org_pitch = (s8_ce(g.ob.x1) - ox0) * bps; u.org_rows = g.ob.y1 - g.ob.y0;
    int org_off = 0, copy_stride = 0;
    u.smem_need = cu_smem_need(g.rb, g.ob, &org_off, &copy_stride, bps, g.two);
    u.org_smem_off = org_off; u.copy_stride = copy_stride;
    const int bfirst = group_range[gi].first, bcount = group_range[gi].second;
    long long acc = 0;
    int ufirst_item = 0, ufirst_b = 0;
    auto emit = [&](int last_local, int item_last) {
      u.job_first = bfirst + ufirst_b; u.item_first = ufirst_item; u.item_last = item_last; u.job_count = last_local - ufirst_b + 1;
      u.variant = bvar[u.job_first];
      units.push_back(u);
    };
    for (int bl = 0; bl < bcount; bl++) {
      const S8Bundle& b = bundles[bfirst + bl];
      const int v = bvar[bfirst + bl];
      const long long per_item = vcost(v);
      if (acc > 0 && v != bvar[bfirst + bl - 1]) { emit(bl - 1, b.item_start); acc = 0; ufirst_item = b.item_start; ufirst_b = bl; }
      int done = 0;
      while (done < b.n_items) {
        const long long room = target[v] - acc;
        const int take = (int)std::min<long long>(b.n_items - done, std::max<long long>(1, (room + per_item - 1) / per_item));
        acc += take * per_item; done += take;
        if (acc >= target[v]) {
          emit(bl, b.item_start + done);
          acc = 0; ufirst_item = b.item_start + done; ufirst_b = (done == b.n_items) ? bl + 1 : bl;
        }
      }
    }
    if (acc > 0) { const S8Bundle& last = bundles[bfirst + bcount - 1]; emit(bcount - 1, last.item_start + last.n_items); }
  }
  std::stable_sort(units.begin(), units.end(), [](const S8Unit& a, const S8Unit& b) { return a.variant < b.variant; });
  for (size_t i = 0; i < units.size(); i++) {
    const int v = units[i].variant;
    if (out->unit_count[v]++ == 0) out->unit_first[v] = (int)i;
    out->smem_of[v] = std::max(out->smem_of[v], units[i].smem_need);
  }
  if (bps == 1 && !getenv("HMB200_NO_TENSOR_MAP")) {
    // Most common window geometry per variant -> that window is staged by ONE tensor-map load (box = pitch x rows, both <= 256)
    // instead of one bulk copy per row; flag 0x100 in the unit's variant word (read by the kernel only).  Only the unshifted copy
    // can come from the TMA unit: a box must start on a 16-byte boundary of the global row (tools/tma_probe.cu).
    for (int v = 0; v < CUV_COUNT; v++) {
      if (variant_is_child(v)) continue;                 // see k_search8_cu: the child kernel stages row by row
      std::map<std::pair<int, int>, int> freq;
      for (int i = out->unit_first[v]; i < out->unit_first[v] + out->unit_count[v]; i++) {
        const S8Unit& u = units[(size_t)i];
        if (u.ref_pitch <= 256 && u.ref_rows <= 256) freq[std::make_pair(u.ref_pitch, u.ref_rows)]++;
      }
      std::pair<int, int> best(0, 0); int n = 0;
      for (auto& kv : freq) if (kv.second > n) { n = kv.second; best = kv.first; }
      if (n == 0) continue;
      out->tma_w[v] = best.first; out->tma_h[v] = best.second; out->tma_units[v] = n;
      if (getenv("HMB200_DEBUG_TMA")) fprintf(stderr, "[hmb200] variant %d: %d of %d units staged by a %d x %d tensor-map box\n", v, n, out->unit_count[v], best.first, best.second);
      for (int i = out->unit_first[v]; i < out->unit_first[v] + out->unit_count[v]; i++)
        if (units[(size_t)i].ref_pitch == best.first && units[(size_t)i].ref_rows == best.second) units[(size_t)i].variant |= 0x100;
    }
  }
  out->n_units = (int)units.size(); out->n_bundles = (int)bundles.size();
  out->rbox = S8Box{s8_fl(all_r.x0), all_r.y0, s8_ce(all_r.x1), all_r.y1};
  out->obox = S8Box{s8_fl(all_o.x0), all_o.y0, s8_ce(all_o.x1), all_o.y1};
  bool ok = cudaMalloc((void**)&out->d_units, units.size() * sizeof(S8Unit)) == cudaSuccess &&
            cudaMalloc((void**)&out->d_bundles, bundles.size() * sizeof(S8Bundle)) == cudaSuccess &&
            cudaMemcpyAsync(out->d_units, units.data(), units.size() * sizeof(S8Unit), cudaMemcpyHostToDevice, stream) == cudaSuccess &&
            cudaMemcpyAsync(out->d_bundles, bundles.data(), bundles.size() * sizeof(S8Bundle), cudaMemcpyHostToDevice, stream) == cudaSuccess &&
            cudaStreamSynchronize(stream) == cudaSuccess;
  if (!ok) { if (err) *err = std::string("cu_build_schedule: ") + cudaGetErrorString(cudaGetLastError()); cu_free_schedule(out); }
  return ok;
}

}  // namespace hmb200
