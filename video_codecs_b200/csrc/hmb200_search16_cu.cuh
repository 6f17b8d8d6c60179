// CU-fused integer full search for 16-bit planes (bit depths 9..14): the hmb200_search8_cu.cuh scheme with packed
// 16x2 arithmetic.  There is no 16-bit SIMD SAD instruction on sm_100a; the direct form costs three instructions per
// 32-bit word (two samples) and candidate -- VIADD.16x2 (o - r), VIADDMNMX.S16x2 (max(r - o, o - r)), IDP.2A.LO (sum of
// the half-words) -- plus hoisted negations, two of the three on the ALU pipe (measured 15.9 T abs-diff/s register-only).
// The kernels use the sum-of-minima form below instead: one ALU-pipe and one FMA-pipe instruction per word and
// candidate (33.3 T abs-diff/s register-only, profiles/r01_microbench_int16.json).
// Distortion precision: (sum << iSubShift) >> (bitDepth - 8)  (TComRdCost.cpp:505-517, DISTORTION_PRECISION_ADJUSTMENT).
// Lane layout: a PAIR of lanes owns 8 candidate columns (lane parity = sample alignment inside the word, four candidates
// two samples apart per lane), 16 pairs per warp-item.  The window is kept twice in shared memory, as staged and shifted
// by one sample (built by the CTA after the bulk copies land), in the kernels for CUs up to 16x16, so that both lanes of a
// pair read aligned words; the windows of 32x32 and 64x64 CUs at +-128 would leave too few candidate rows per unit
// that way and keep one copy and a funnel shift per word.
#pragma once
#include "hmb200_search8_cu.cuh"

namespace hmb200 {

__host__ __device__ constexpr int cu16_ky(int S) { return S == 8 ? 2 : 1; }

// ---- sum-of-minima form --------------------------------------------------------------------------------------------
//     |o - r| = o + r - 2 min(o, r)   =>   SAD(cell, candidate) = A(cell) + B(cell, candidate) - 2 M(cell, candidate)
// M = sum of min(o, r): VIMNMX.U16x2 (ALU pipe) + IDP.2A.LO (FMA pipe) per two samples, no negated operands -- the
// register-only rate of this pair is 2.1x that of the three-instruction |o - r| sequence (profiles/r01_microbench_int16.json).
// B = box sum of the reference under the cell: one IDP.2A per shifted reference word and row into per-word column
// accumulators W[j] (shared by the lane's four candidates and all cells of the chunk), folded once per strip.
// A = sum of the original cell: once per CU and warp, kept in shared memory.
// Samples are < 2^15 (bit depth <= 14), so unsigned minima and the signed 16-bit dot product agree.
template <int WW>
__device__ __forceinline__ void cu16_load_org_raw(const uint8_t* p, uint32_t (&o)[WW]) {
#pragma unroll
  for (int i = 0; i < WW / 4; i++) {
    const uint4 v = reinterpret_cast<const uint4*>(p)[i];
    o[4 * i] = v.x; o[4 * i + 1] = v.y; o[4 * i + 2] = v.z; o[4 * i + 3] = v.w;
  }
}

// one reference row against WW original words: M[c][k] += sum of minima, W[j] += sum of the j-th reference word.
// SHIFT: rp8 is word-aligned and the lane's words start `sh` bits in (one funnel shift per word); otherwise rp8 points
// into the window copy of the lane's sample phase and the words are read as they are.
// A8: rp8 is 8-byte aligned (host: every bundle of the unit has win_off % 8 == 0 and blocks start on multiples of 8 samples):
// the row is read with 64-bit loads.  A pair of lanes is 16 bytes apart from the next pair, so a 32-bit load puts pairs q and
// q + 8 on the same bank (two wavefronts per load: 42 % of the shared-memory wavefronts were conflicts in round 1); a 64-bit
// load is served per half-warp, whose eight pairs (and the two window copies, 8 bytes apart modulo 128) fall on 32 distinct banks.
template <int WW, int NC, bool SHIFT, bool A8>
__device__ __forceinline__ void cu16_row_min(const uint8_t* rp8, const uint32_t (&o)[WW], uint32_t sh, uint32_t (*M)[4],
                                             uint32_t (&W)[WW + 3]) {
  constexpr int NW = SHIFT ? WW + 4 : WW + 3;
  uint32_t rw[(NW + 1) & ~1];
  if constexpr (A8) {
#pragma unroll
    for (int i = 0; i < (NW + 1) / 2; i++) {
      const uint2 v = reinterpret_cast<const uint2*>(rp8)[i];
      rw[2 * i] = v.x; rw[2 * i + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NW; i++) rw[i] = reinterpret_cast<const uint32_t*>(rp8)[i];
  }
  uint32_t lo = SHIFT ? rw[0] : 0u;
#pragma unroll
  for (int j = 0; j < WW + 3; j++) {
    uint32_t sw;
    if constexpr (SHIFT) {
      const uint32_t hi = rw[j + 1];
      sw = __funnelshift_r(lo, hi, sh);
      lo = hi;
    } else {
      sw = rw[j];
    }
    W[j] = (uint32_t)__dp2a_lo((int)sw, 0x0101, (int)W[j]);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int i = j - k;
      if (i >= 0 && i < WW) M[i / (WW / NC)][k] = (uint32_t)__dp2a_lo((int)__vminu2(sw, o[i]), 0x0101, (int)M[i / (WW / NC)][k]);
    }
  }
}

// M[c][k] <- A[c] + B(c, k) - 2 M[c][k]  (= the SAD) with B(c, k) = W[c*cw + k] + ... + W[c*cw + k + cw - 1]
template <int WW, int NC>
__device__ __forceinline__ void cu16_fold(uint32_t (*M)[4], const uint32_t (&W)[WW + 3], const uint32_t* asum) {
  constexpr int CWD = WW / NC;
  if constexpr (CWD == 2) {                          // 16x16 CUs: one three-input add per cell and candidate
#pragma unroll
    for (int c = 0; c < NC; c++) {
      const uint32_t a = asum[c];
#pragma unroll
      for (int k = 0; k < 4; k++) M[c][k] = (W[c * 2 + k] + W[c * 2 + k + 1] + a) - 2u * M[c][k];
    }
  } else {
    uint32_t P[WW + 4];
    P[0] = 0;
#pragma unroll
    for (int j = 0; j < WW + 3; j++) P[j + 1] = P[j] + W[j];
#pragma unroll
    for (int c = 0; c < NC; c++) {
      const uint32_t a = asum[c];
#pragma unroll
      for (int k = 0; k < 4; k++) M[c][k] = (P[c * CWD + k + CWD] - P[c * CWD + k] + a) - 2u * M[c][k];
    }
  }
}

// rows row0, row0 + rstep, ... (NROWS of them) of one chunk of WW words: acc[c][k] (zero on entry) <- SAD over those rows
template <int WW, int NC, int NROWS, bool SHIFT, bool A8>
__device__ __forceinline__ void cu16_strip_min(const uint8_t* refp, int ref_pitch, const uint8_t* orgp, int org_pitch, int row0,
                                               int rstep, uint32_t sh, uint32_t (*acc)[4], const uint32_t* asum) {
  uint32_t W[WW + 3];
#pragma unroll
  for (int j = 0; j < WW + 3; j++) W[j] = 0;
  if constexpr (NROWS <= 2) {                       // 16x16 CUs: two rows per strip and parity, straight-line code
#pragma unroll
    for (int i = 0; i < NROWS; i++) {
      uint32_t o[WW];
      cu16_load_org_raw<WW>(orgp + (row0 + i * rstep) * org_pitch, o);
      cu16_row_min<WW, NC, SHIFT, A8>(refp + (row0 + i * rstep) * ref_pitch, o, sh, acc, W);
    }
  } else {
#pragma unroll 1
    for (int i = 0, row = row0; i < NROWS; i++, row += rstep) {
      uint32_t o[WW];
      cu16_load_org_raw<WW>(orgp + row * org_pitch, o);
      cu16_row_min<WW, NC, SHIFT, A8>(refp + row * ref_pitch, o, sh, acc, W);
    }
  }
  cu16_fold<WW, NC>(acc, W, asum);
}

// A sums of one CU for one warp: dst[r*4 + c] = cell (r, c) over the rows the search visits, dst[16 + r] = the odd rows
// of strip r over the whole CU width (FEN partitions of height <= 8 read them, TComRdCost.cpp:566-571)
template <int S, bool PARITY>
__device__ __forceinline__ void cu16_org_sums(const uint8_t* orgp, int org_pitch, int lane, uint32_t* dst) {
  constexpr int WW = S / 2, G = S / 4, CWD = WW / 4, RSTEP = PARITY ? 2 : 1;
  if (lane < 20) {
    const bool odd = lane >= 16;
    const int r = odd ? lane - 16 : lane >> 2, c = odd ? 0 : lane & 3;
    const int nw = odd ? WW : CWD;
    uint32_t a = 0;
    if (!odd || PARITY) {
      for (int rr = 0; rr < G; rr += RSTEP) {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(orgp + (r * G + rr + (odd ? 1 : 0)) * org_pitch) + c * CWD;
        for (int j = 0; j < nw; j++) a = (uint32_t)__dp2a_lo((int)row[j], 0x0101, (int)a);
      }
    }
    dst[lane] = a;
  }
}

// CHILD tiles (a 16x16 CU that carries its 8x8 children): per-cell sums of the rows the even-row pass visits (every row without
// FEN) in dst[r*4 + c], of the odd rows (FEN) in dst[16 + r*4 + c]
template <bool PARITY>
__device__ __forceinline__ void cu16_org_sums_child(const uint8_t* orgp, int org_pitch, int lane, uint32_t* dst) {
  const bool odd = lane >= 16;
  const int r = (lane & 15) >> 2, c = lane & 3;
  uint32_t a = 0;
  if (!odd || PARITY) {
    for (int rr = odd ? 1 : 0; rr < 4; rr += PARITY ? 2 : 1) {
      const uint32_t* row = reinterpret_cast<const uint32_t*>(orgp + (r * 4 + rr) * org_pitch) + c * 2;
      a = (uint32_t)__dp2a_lo((int)row[0], 0x0101, (int)a);
      a = (uint32_t)__dp2a_lo((int)row[1], 0x0101, (int)a);
    }
  }
  dst[lane] = a;
}

// key of one PU: ((sum << ss) >> shr) scaled into the cost field
// (16-bit planes have bit depth >= 9, so shr >= 1 >= ss and the two shifts are one: sums are below 2^31)
__device__ __forceinline__ void cu16_min(uint32_t& best, uint32_t sum, int ss, int shr, uint32_t base) {
  best = min(best, ((sum >> (shr - ss)) << CU_LOCAL_BITS) + base);
}

template <int S, bool FEN>
__device__ __forceinline__ void cu16_epilogue(const uint32_t (&E)[4][4][4], const uint32_t (&O)[4][4], int k, int shr, uint32_t base,
                                              uint32_t (&best)[CU_SLOTS]) {
  uint32_t er[4], ec[4];
#pragma unroll
  for (int r = 0; r < 4; r++) er[r] = E[r][0][k] + E[r][1][k] + E[r][2][k] + E[r][3][k];
#pragma unroll
  for (int c = 0; c < 4; c++) ec[c] = E[0][c][k] + E[1][c][k] + E[2][c][k] + E[3][c][k];
  constexpr bool f1 = FEN && cu_slot_h(S, 1) <= 8;
  constexpr bool f5 = FEN && cu_slot_h(S, 5) <= 8;
  constexpr int ss = FEN ? 1 : 0;
  const uint32_t top = er[0] + er[1], bot = er[2] + er[3];
  cu16_min(best[0], top + bot, ss, shr, base);
  if (f1) { cu16_min(best[1], top + O[0][k] + O[1][k], 0, shr, base); cu16_min(best[2], bot + O[2][k] + O[3][k], 0, shr, base); }
  else    { cu16_min(best[1], top, ss, shr, base);                    cu16_min(best[2], bot, ss, shr, base); }
  cu16_min(best[3], ec[0] + ec[1], ss, shr, base);
  cu16_min(best[4], ec[2] + ec[3], ss, shr, base);
  if (f5) { cu16_min(best[5], er[0] + O[0][k], 0, shr, base); cu16_min(best[8], er[3] + O[3][k], 0, shr, base); }
  else    { cu16_min(best[5], er[0], ss, shr, base);          cu16_min(best[8], er[3], ss, shr, base); }
  cu16_min(best[6], er[1] + bot, ss, shr, base);
  cu16_min(best[7], top + er[2], ss, shr, base);
  cu16_min(best[9], ec[0], ss, shr, base);
  cu16_min(best[10], ec[1] + ec[2] + ec[3], ss, shr, base);
  cu16_min(best[11], ec[0] + ec[1] + ec[2], ss, shr, base);
  cu16_min(best[12], ec[3], ss, shr, base);
}

// CHILD (S == 16): the CU's pass also yields the PUs of its four 8x8 child CUs, as in k_search8_cu.  33 running minima do not fit
// next to the sum-of-minima accumulators, so the 20 child minima live in shared memory, one word per thread and PU: a tile folds
// its four candidate columns first and touches each word once.
template <int S, bool FEN, bool A8, bool CHILD = false>
__global__ void __launch_bounds__(S8_THREADS, 2)
k_search16_cu(const S8Unit* __restrict__ units, const S8Bundle* __restrict__ bundles, unsigned long long* __restrict__ keys,
              DevPlane cur_plane, DevPlane ref_plane) {
  static_assert(!CHILD || S == 16, "only 16x16 CUs carry child CUs");
  constexpr int NSLOT = (S == 8) ? 5 : CU_SLOTS;
  constexpr int KY = cu16_ky(S);
  constexpr int WW = S / 2;                            // 32-bit words per CU row
  constexpr int G = S / 4;                             // rows per strip (S >= 16)
  constexpr int CH = WW > 16 ? 16 : WW;                // words per row chunk (keeps the original row in <= 32 registers)
  constexpr bool PARITY = FEN && S >= 16, ODD_ALL = PARITY && S == 16, ODD_EDGE = PARITY && S == 32;
  constexpr bool TWO = S <= 16;                        // two window copies (host: cu16_two_phase)
  extern __shared__ __align__(128) uint8_t s8_smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ S8Bundle s_bd[S8_WARPS];
  __shared__ uint32_t s_asum[S8_WARPS][CHILD ? 32 : 20];
  __shared__ uint32_t s_cbest[CHILD ? CU_CHILD_SLOTS : 1][CHILD ? S8_THREADS : 1];

  const S8Unit un = units[blockIdx.x];
  uint8_t* s_ref = s8_smem;
  uint8_t* s_org = s8_smem + un.org_smem_off;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();                                      // the initialised barrier is visible before its first use
  if (threadIdx.x == 0) mbar_expect_tx(&s_bar, (uint32_t)(un.ref_pitch * un.ref_rows + un.org_pitch * un.org_rows));
  __syncthreads();                                      // ... and armed before any copy can complete on it
  {
    const uint8_t* gref = reinterpret_cast<const uint8_t*>(ref_plane.base) +
                          ((size_t)(un.ref_by + ref_plane.margin_y) * ref_plane.pitch + (un.ref_bx + ref_plane.margin_x)) * 2;
    for (int r = threadIdx.x; r < un.ref_rows; r += S8_THREADS)
      bulk_g2s(s_ref + r * un.ref_pitch, gref + (size_t)r * ref_plane.pitch * 2, (uint32_t)un.ref_pitch, &s_bar);
    const uint8_t* gorg = reinterpret_cast<const uint8_t*>(cur_plane.base) +
                          ((size_t)(un.org_by + cur_plane.margin_y) * cur_plane.pitch + (un.org_bx + cur_plane.margin_x)) * 2;
    for (int r = (int)threadIdx.x - 128; r < un.org_rows; r += S8_THREADS)
      if (r >= 0) bulk_g2s(s_org + r * un.org_pitch, gorg + (size_t)r * cur_plane.pitch * 2, (uint32_t)un.org_pitch, &s_bar);
  }
  mbar_wait(&s_bar, 0);
  if constexpr (TWO) {
    // second copy of the window, one sample to the left: word w = samples (2w + 1, 2w + 2)
    const uint32_t* p0 = reinterpret_cast<const uint32_t*>(s_ref);
    uint32_t* p1 = reinterpret_cast<uint32_t*>(s_ref + un.copy_stride);
    const int nw = (un.ref_pitch * un.ref_rows) >> 2;
    for (int w = threadIdx.x; w < nw; w += S8_THREADS) p1[w] = __funnelshift_r(p0[w], p0[w + 1], 16);
    __syncthreads();
  }

  int bslot = un.job_first;
  S8Bundle& bd = s_bd[warp];
  auto load_bundle = [&]() {
    __syncwarp();
    reinterpret_cast<int32_t*>(&bd)[lane] = reinterpret_cast<const int32_t*>(&bundles[bslot])[lane];
    if constexpr (CHILD) reinterpret_cast<int32_t*>(&bd)[lane + 32] = reinterpret_cast<const int32_t*>(&bundles[bslot])[lane + 32];
    __syncwarp();
    if constexpr (CHILD) {
      cu16_org_sums_child<PARITY>(s_org + bd.org_off, un.org_pitch, lane, s_asum[warp]);
    } else if constexpr (S >= 16) {
      cu16_org_sums<S, PARITY>(s_org + bd.org_off, un.org_pitch, lane, s_asum[warp]);
    } else if (lane < 4) {                              // 8x8 CU: four 4x4 quadrants
      uint32_t a = 0;
#pragma unroll
      for (int rr = 0; rr < 4; rr++) {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(s_org + bd.org_off + ((lane >> 1) * 4 + rr) * un.org_pitch) + (lane & 1) * 2;
        a = (uint32_t)__dp2a_lo((int)row[0], 0x0101, (int)a);
        a = (uint32_t)__dp2a_lo((int)row[1], 0x0101, (int)a);
      }
      s_asum[warp][lane] = a;
    }
    __syncwarp();
  };
  load_bundle();
  constexpr int LK = (S == 8) ? 4 : 2;
  uint32_t best[CU_SLOTS];
#pragma unroll
  for (int s = 0; s < CU_SLOTS; s++) best[s] = 0xffffffffu;
  volatile uint32_t* cbest = &s_cbest[0][CHILD ? threadIdx.x : 0];      // this thread's child minima: cbest[slot * S8_THREADS]
  if constexpr (CHILD) {
#pragma unroll
    for (int s = 0; s < CU_CHILD_SLOTS; s++) cbest[s * S8_THREADS] = 0xffffffffu;
  }
  int first_item = un.item_first + warp;
  auto flush_one = [&](uint32_t v, int out) {
    unsigned long long b = ~0ull;
    if (v != 0xffffffffu) {
      const uint32_t local = v & ((1u << CU_LOCAL_BITS) - 1u);
      const int it = first_item + (int)(local >> LK) * S8_WARPS;
      const int q = (it - bd.item_start) * 16 + (lane >> 1);
      const int g = q / bd.n_blk, blk = q - g * bd.n_blk;
      const int w = (int)(local & ((1u << LK) - 1u));
      const int cyi = bd.cy_first + g * KY + (w >> 2), cxi = (A8 ? blk * 8 : min(blk * 8, bd.nx - 8)) + (lane & 1) + 2 * (w & 3);
      b = make_key(v >> CU_LOCAL_BITS, (uint32_t)(cyi * bd.nx + cxi));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, b, o);
      b = other < b ? other : b;
    }
    if (lane == 0 && b != ~0ull && out >= 0) atomicMin(&keys[out], b);
  };
  auto flush = [&]() {
#pragma unroll
    for (int s = 0; s < NSLOT; s++) { flush_one(best[s], bd.out_idx[s]); best[s] = 0xffffffffu; }
    if constexpr (CHILD) {
#pragma unroll 1
      for (int s = 0; s < CU_CHILD_SLOTS; s++) { flush_one(cbest[s * S8_THREADS], bd.child_idx[s]); cbest[s * S8_THREADS] = 0xffffffffu; }
    }
  };

  for (int item = un.item_first + warp; item < un.item_last; item += S8_WARPS) {
    if (item >= bd.item_start + bd.n_items) {
      flush();
      do { bslot++; } while (item >= bundles[bslot].item_start + bundles[bslot].n_items);
      load_bundle();
      first_item = item;
    }
    const uint32_t tile_local = (uint32_t)((item - first_item) / S8_WARPS) << LK;
    const int q = (item - bd.item_start) * 16 + (lane >> 1);
    if (q < bd.n_blk * bd.n_rowgroups) {
      const int g = q / bd.n_blk, blk = q - g * bd.n_blk;
      const int cyl0 = g * KY;                                            // candidate row inside this (row-split) bundle
      // the last block overlaps its neighbour - or (A8: blocks stay on multiples of 8 samples) runs past the window and is masked
      const int cxi0 = (A8 ? blk * 8 : min(blk * 8, bd.nx - 8)) + (lane & 1);
      const int off = bd.win_off + cyl0 * un.ref_pitch + cxi0 * 2;        // bytes
      const uint8_t* refp = s_ref + ((TWO && (off & 2)) ? un.copy_stride : 0) + (off & ~3);
      const uint32_t sh = (uint32_t)(off & 3) * 8u;      // used by the single-copy kernels only
      const uint8_t* orgp = s_org + bd.org_off;
      const int shr = bd.shr;
      uint32_t px[4];
#pragma unroll
      for (int k = 0; k < 4; k++) px[k] = bd.lambda * eg_bits(((bd.lt_x + cxi0 + 2 * k) << 2) - bd.pred_x);

      if constexpr (S == 8) {
        // sum-of-minima form with two candidate rows per lane: reference row r meets original row r - jy
        uint32_t o[8][4];
#pragma unroll
        for (int r = 0; r < 8; r++) cu16_load_org_raw<4>(orgp + r * un.org_pitch, o[r]);
        uint32_t Q[KY][2][2][4], W[KY][2][7];
#pragma unroll
        for (int a = 0; a < KY; a++)
#pragma unroll
          for (int b = 0; b < 2; b++) {
#pragma unroll
            for (int j = 0; j < 7; j++) W[a][b][j] = 0;
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
              for (int k = 0; k < 4; k++) Q[a][b][c][k] = 0;
          }
#pragma unroll
        for (int r = 0; r < 8 + KY - 1; r++) {
          uint32_t rw[8];
          if constexpr (A8) {
#pragma unroll
            for (int i = 0; i < 4; i++) { const uint2 v = reinterpret_cast<const uint2*>(refp + r * un.ref_pitch)[i]; rw[2 * i] = v.x; rw[2 * i + 1] = v.y; }
          } else {
#pragma unroll
            for (int i = 0; i < 7; i++) rw[i] = reinterpret_cast<const uint32_t*>(refp + r * un.ref_pitch)[i];
          }
#pragma unroll
          for (int j = 0; j < 4 + 3; j++) {
            const uint32_t sw = rw[j];
#pragma unroll
            for (int jy = 0; jy < KY; jy++) {
              if (r - jy >= 0 && r - jy < 8) {
                W[jy][(r - jy) >> 2][j] = (uint32_t)__dp2a_lo((int)sw, 0x0101, (int)W[jy][(r - jy) >> 2][j]);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                  const int i = j - k;
                  if (i >= 0 && i < 4)
                    Q[jy][(r - jy) >> 2][i >> 1][k] =
                        (uint32_t)__dp2a_lo((int)__vminu2(sw, o[r - jy][i]), 0x0101, (int)Q[jy][(r - jy) >> 2][i >> 1][k]);
                }
              }
            }
          }
        }
        {
          const uint32_t* asum = s_asum[warp];
#pragma unroll
          for (int b = 0; b < 2; b++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
              const uint32_t a = asum[b * 2 + c];
#pragma unroll
              for (int jy = 0; jy < KY; jy++)
#pragma unroll
                for (int k = 0; k < 4; k++)
                  Q[jy][b][c][k] = a + W[jy][b][2 * c + k] + W[jy][b][2 * c + k + 1] - 2u * Q[jy][b][c][k];
            }
        }
#pragma unroll
        for (int jy = 0; jy < KY; jy++) {
          const int cyi = bd.cy_first + cyl0 + jy;
          if (cyl0 + jy < bd.ny) {
            const uint32_t py = bd.lambda * eg_bits(((bd.lt_y + cyi) << 2) - bd.pred_y);
#pragma unroll
            for (int k = 0; k < 4; k++) {
              if (A8 && cxi0 + 2 * k >= bd.nx) continue;                  // masked column of the last block
              const uint32_t base = (((px[k] + py) >> 16) << CU_LOCAL_BITS) | tile_local | (uint32_t)(jy * 4 + k);
              const uint32_t t = Q[jy][0][0][k] + Q[jy][0][1][k], b = Q[jy][1][0][k] + Q[jy][1][1][k];
              const uint32_t l = Q[jy][0][0][k] + Q[jy][1][0][k], r = Q[jy][0][1][k] + Q[jy][1][1][k];
              cu16_min(best[0], t + b, 0, shr, base);
              cu16_min(best[1], t, 0, shr, base);
              cu16_min(best[2], b, 0, shr, base);
              cu16_min(best[3], l, 0, shr, base);
              cu16_min(best[4], r, 0, shr, base);
            }
          }
        }
      } else if constexpr (CHILD) {
        // two halves of two 4-row strips; per strip the per-cell SADs of the even rows (every row without FEN) and of the odd rows
        const uint32_t* asum = s_asum[warp];
        const bool row_ok = cyl0 < bd.ny;
        const uint32_t py = bd.lambda * eg_bits(((bd.lt_y + bd.cy_first + cyl0) << 2) - bd.pred_y);
        uint32_t base[4];
        bool valid[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          base[k] = (((px[k] + py) >> 16) << CU_LOCAL_BITS) | tile_local | (uint32_t)k;
          valid[k] = row_ok && (!A8 || cxi0 + 2 * k < bd.nx);
        }
        constexpr int ss = FEN ? 1 : 0;
        // all candidate columns of one child PU -> its minimum in shared memory
        auto child_min = [&](int slot, const uint32_t (&sum)[4]) {
          uint32_t m = 0xffffffffu;
#pragma unroll
          for (int k = 0; k < 4; k++) if (valid[k]) m = min(m, ((sum[k] >> shr) << CU_LOCAL_BITS) + base[k]);
          cbest[slot * S8_THREADS] = min(cbest[slot * S8_THREADS], m);
        };
        auto own_min = [&](int slot, const uint32_t (&sum)[4], int sub) {
#pragma unroll
          for (int k = 0; k < 4; k++) if (valid[k]) cu16_min(best[slot], sum[k], sub, shr, base[k]);
        };
        uint32_t ec[4][4], P[4], P1[4], T0[4][4], all0[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { P[k] = 0; P1[k] = 0; all0[k] = 0;
#pragma unroll
          for (int c = 0; c < 4; c++) { ec[c][k] = 0; T0[c][k] = 0; } }
#pragma unroll
        for (int r = 0; r < 4; r++) {
          uint32_t Tc[4][4], er[4], alls[4];
          {
            uint32_t Ec[4][4];
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
              for (int k = 0; k < 4; k++) Ec[c][k] = 0;
            cu16_strip_min<8, 4, FEN ? 2 : 4, false, A8>(refp, un.ref_pitch, orgp, un.org_pitch, r * 4, FEN ? 2 : 1, sh, Ec, asum + r * 4);
#pragma unroll
            for (int k = 0; k < 4; k++) {
              er[k] = (Ec[0][k] + Ec[1][k]) + (Ec[2][k] + Ec[3][k]);
#pragma unroll
              for (int c = 0; c < 4; c++) { ec[c][k] += Ec[c][k]; Tc[c][k] = Ec[c][k]; }
            }
          }
          if constexpr (FEN) {
            uint32_t Oc[4][4];
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
              for (int k = 0; k < 4; k++) Oc[c][k] = 0;
            cu16_strip_min<8, 4, 2, false, A8>(refp, un.ref_pitch, orgp, un.org_pitch, r * 4 + 1, 2, sh, Oc, asum + 16 + r * 4);
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
              for (int k = 0; k < 4; k++) Tc[c][k] += Oc[c][k];
          }
#pragma unroll
          for (int k = 0; k < 4; k++) { alls[k] = (Tc[0][k] + Tc[1][k]) + (Tc[2][k] + Tc[3][k]); P[k] += er[k]; }
          if (r == 0) { own_min(5, alls, 0);                                  // 2NxnU top: 16x4, every row
#pragma unroll
            for (int k = 0; k < 4; k++) P1[k] = P[k]; }
          if (r == 2) {                                                       // 2NxnD top: rows 0..11 (sub-sampled under FEN)
            own_min(7, P, ss);
          }
          if (r == 3) own_min(8, alls, 0);                                    // 2NxnD bottom: 16x4, every row
          if ((r & 1) == 0) {
#pragma unroll
            for (int k = 0; k < 4; k++) { all0[k] = alls[k];
#pragma unroll
              for (int c = 0; c < 4; c++) T0[c][k] = Tc[c][k]; }
          } else {
            const int h = r >> 1;
#pragma unroll
            for (int j = 0; j < 2; j++) {
              uint32_t s88[4], top[4], bot[4], lft[4], rgt[4];
#pragma unroll
              for (int k = 0; k < 4; k++) {
                top[k] = T0[2 * j][k] + T0[2 * j + 1][k]; bot[k] = Tc[2 * j][k] + Tc[2 * j + 1][k];
                lft[k] = T0[2 * j][k] + Tc[2 * j][k];     rgt[k] = T0[2 * j + 1][k] + Tc[2 * j + 1][k];
                s88[k] = top[k] + bot[k];
              }
              const int cs = 5 * (2 * h + j);
              child_min(cs + 0, s88); child_min(cs + 1, top); child_min(cs + 2, bot); child_min(cs + 3, lft); child_min(cs + 4, rgt);
            }
            uint32_t half[4];
#pragma unroll
            for (int k = 0; k < 4; k++) half[k] = all0[k] + alls[k];
            own_min(h == 0 ? 1 : 2, half, 0);                                 // 2NxN: 16x8, every row
          }
        }
        {
          uint32_t t[4];
          own_min(0, P, ss);
#pragma unroll
          for (int k = 0; k < 4; k++) t[k] = ec[0][k] + ec[1][k];
          own_min(3, t, ss);
#pragma unroll
          for (int k = 0; k < 4; k++) t[k] = ec[2][k] + ec[3][k];
          own_min(4, t, ss);
#pragma unroll
          for (int k = 0; k < 4; k++) t[k] = P[k] - P1[k];
          own_min(6, t, ss);
          own_min(9, ec[0], ss);
#pragma unroll
          for (int k = 0; k < 4; k++) t[k] = ec[1][k] + ec[2][k] + ec[3][k];
          own_min(10, t, ss);
#pragma unroll
          for (int k = 0; k < 4; k++) t[k] = ec[0][k] + ec[1][k] + ec[2][k];
          own_min(11, t, ss);
          own_min(12, ec[3], ss);
        }
      } else {
        uint32_t E[4][4][4], O[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int k = 0; k < 4; k++) {
            O[a][k] = 0;
#pragma unroll
            for (int c = 0; c < 4; c++) E[a][c][k] = 0;
          }
        constexpr int NCH = WW / CH;                // chunks per row
        constexpr int CPC = 4 / NCH;                // cell columns per chunk
        constexpr int RSTEP = PARITY ? 2 : 1;
        const uint32_t* asum = s_asum[warp];
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const bool odd_here = ODD_ALL || (ODD_EDGE && (r == 0 || r == 3));
#pragma unroll
          for (int ch = 0; ch < NCH; ch++)
            cu16_strip_min<CH, CPC, G / RSTEP, !TWO, A8>(refp + ch * CH * 4, un.ref_pitch, orgp + ch * CH * 4, un.org_pitch, r * G, RSTEP, sh,
                                                     &E[r][ch * CPC], asum + r * 4 + ch * CPC);
          if (PARITY && odd_here) {
            static_assert(!PARITY || S == 64 || NCH == 1, "odd rows: one chunk per row");
            cu16_strip_min<CH, 1, G / 2, !TWO, A8>(refp, un.ref_pitch, orgp, un.org_pitch, r * G + 1, 2, sh, &O[r], asum + 16 + r);
          }
        }
        if (cyl0 < bd.ny) {
          const int cyi = bd.cy_first + cyl0;
          const uint32_t py = bd.lambda * eg_bits(((bd.lt_y + cyi) << 2) - bd.pred_y);
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (!A8 || cxi0 + 2 * k < bd.nx)
              cu16_epilogue<S, FEN>(E, O, k, shr, (((px[k] + py) >> 16) << CU_LOCAL_BITS) | tile_local | (uint32_t)k, best);
        }
      }
    }
  }
  flush();
}

// variants 0..6 as cu_variant(); 7..13 = the same with A8 (64-bit loads, masked last block) for units whose windows are 8-byte aligned;
// 14..17 = 16x16 CUs that carry their 8x8 children: 14 + FEN + 2 * A8
constexpr int CUV16_CHILD = 2 * CUV_BASE_COUNT;
constexpr int CUV16_COUNT = CUV16_CHILD + 4;
typedef void (*S16CuKernel)(const S8Unit*, const S8Bundle*, unsigned long long*, DevPlane, DevPlane);
inline const S16CuKernel* search16_cu_kernels() {
  static const S16CuKernel table[CUV16_COUNT] = {
      k_search16_cu<8, false, false>, k_search16_cu<16, false, false>, k_search16_cu<16, true, false>, k_search16_cu<32, false, false>,
      k_search16_cu<32, true, false>, k_search16_cu<64, false, false>, k_search16_cu<64, true, false>,
      k_search16_cu<8, false, true>, k_search16_cu<16, false, true>, k_search16_cu<16, true, true>, k_search16_cu<32, false, true>,
      k_search16_cu<32, true, true>, k_search16_cu<64, false, true>, k_search16_cu<64, true, true>,
      k_search16_cu<16, false, false, true>, k_search16_cu<16, true, false, true>, k_search16_cu<16, false, true, true>,
      k_search16_cu<16, true, true, true> };
  return table;
}
inline int cu16_configure(std::string* err) {
  const S16CuKernel* k = search16_cu_kernels();
  for (int v = 0; v < CUV16_COUNT; v++) {
    cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(k[v]), cudaFuncAttributeMaxDynamicSharedMemorySize, S8_SMEM_MAX);
    if (e != cudaSuccess) { if (err) *err = std::string("cudaFuncSetAttribute(k_search16_cu): ") + cudaGetErrorString(e); return HMB200_ERR_CUDA; }
  }
  return HMB200_OK;
}

}  // namespace hmb200
