// Tiled byte-SIMD integer full search for 8-bit planes (the hot kernel of the path).
//
// Reproduces TEncSearch::xPatternSearch (TLibEncoder/TEncSearch.cpp:3786-3843) with the xGetSADnn family
// (TLibCommon/TComRdCost.cpp:489-953, iSubShift included) and TComRdCost::getCost (TComRdCost.h:172-189) for many
// PUs per launch.
//
// Work decomposition (host side, search8_build_schedule):
//   group  = PUs whose reference footprints share one shared-memory window (canonical job list: the 593 PUs of a CTU,
//            window = CTU +- search range) and whose originals share one original tile.
//   unit   = a contiguous slice of a group's warp-item list; one CTA per unit.  The CTA stages the window and the
//            original tile with bulk-async (TMA) copies, then its warps pull warp-items from a shared counter.
//   warp-item (quad mode) = 8 quads; a quad owns a 16-candidate-wide, KY-candidate-tall tile of one PU: lane s of the
//            quad owns the columns s, s+4, s+8, s+12 (same byte alignment => one funnel shift per reference word serves
//            all four), so every reference word fetched from shared memory feeds 4*KY VABSDIFF4.ACC.
//   warp-item (column mode) = 32 lanes, each a (leftover column, row group) pair: windows are 129 = 8*16 + 1 wide.
// Under FEN (iSubShift = 1) a PU only visits its even rows, so candidates of equal row parity share reference rows:
// row groups are formed per parity and the tile walks the window with a two-row stride.
// Argmin: per-thread running minimum of the 64-bit key (cost << 32 | raster index); lexicographic order on
// (cost, raster index) == strict '<' in raster scan order (TEncSearch.cpp:3813-3835).  Warps flush their minimum to
// a global key per PU with atomicMin; k_search8_finalize turns keys into (rcMv, ruiSAD).
#pragma once
#include <string>
#include <vector>
#include <algorithm>
#include <cuda.h>
#include "hmb200_device.cuh"
#include "hmb200_generic.cuh"

namespace hmb200 {

constexpr int S8_THREADS = 256;
constexpr int S8_WARPS = S8_THREADS / 32;
constexpr int S8_SLACK_ROWS = 16;            // rows past the window that masked candidates may touch
constexpr int S8_SMEM_SHARED2 = 110 * 1024;  // per-CTA dynamic smem that still lets two CTAs share an SM
constexpr int S8_SMEM_MAX = 200 * 1024;

struct S8Job {              // 80 bytes
  int32_t out_idx;          // index into keys / results
  int32_t org_off;          // byte offset of the PU's top-left sample in the staged original tile
  int32_t win_off;          // byte offset of candidate (lt_x, lt_y)'s top-left sample in the staged window
  int32_t nx, ny;           // candidates per row / rows of candidates
  int32_t ww, hn;           // PU width in 32-bit words, rows visited (h >> ss)
  int32_t ss;               // iSubShift
  int32_t lt_x, lt_y, pred_x, pred_y;
  uint32_t lambda;
  int32_t ky;               // candidate rows per tile
  int32_t n_g0, n_g;        // row groups of parity 0 (all of them when ss == 0) / in total
  int32_t item_start;       // first warp-item within the group's list
  int32_t n_qitems;         // quad-mode warp-items
  int32_t n_items;          // all warp-items
  int32_t variant;          // tile instantiation
};

struct S8Unit {             // 64 bytes
  int32_t ref_bx, ref_by;   // window origin in picture coordinates (ref_bx % 16 == 0; planes need margin_x % 16 == 0)
  int32_t ref_pitch, ref_rows;
  int32_t org_bx, org_by, org_pitch, org_rows;
  int32_t job_first, job_count;
  int32_t item_first, item_last;
  int32_t org_smem_off;     // byte offset of the original tile in dynamic smem
  int32_t variant;          // tile variant of every PU of this unit
  int32_t smem_need;        // dynamic shared memory of this unit
  int32_t copy_stride;      // CU-fused 8-bit kernels: bytes between the four byte-shifted copies of the window
};

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + bulk-async copy (SASS: SYNCS.*, UBLKCP)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// one 2-D box of a tensor map (cuTensorMapEncodeTiled on the host) into shared memory; SASS: UTMALDG.  dst is 128-byte aligned; x
// (the coordinate of the contiguous dimension) times the element size must be a multiple of 16 - otherwise the instruction
// traps as illegal (tools/tma_probe.cu); elements outside the tensor read as 0.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// SAD tiles.  refp: word-aligned address at or below the lane's first candidate (row 0 of the tile); sh: bit shift
// (8 * byte misalignment); rows advance by ref_step / org_step bytes (pitch << ss).  acc[jy][k]: candidate row jy
// (tile rows are ref_step apart), column k*4 bytes to the right.
// ---------------------------------------------------------------------------------------------------------------

// whole PU in registers, fully unrolled: no masked work (rows visited <= 8, width <= 32)
template <int WW, int HN, int KY>
__device__ __forceinline__ void tile_static(const uint8_t* refp, int ref_step, const uint8_t* orgp, int org_step, uint32_t sh,
                                            uint32_t (&acc)[KY][4]) {
  uint32_t o[HN][WW];
#pragma unroll
  for (int r = 0; r < HN; r++) {
    const uint32_t* op = reinterpret_cast<const uint32_t*>(orgp + r * org_step);
#pragma unroll
    for (int i = 0; i < WW; i++) o[r][i] = op[i];
  }
#pragma unroll
  for (int r = 0; r < HN + KY - 1; r++) {
    const uint32_t* rp = reinterpret_cast<const uint32_t*>(refp + r * ref_step);
    uint32_t lo = rp[0];
#pragma unroll
    for (int j = 0; j < WW + 3; j++) {
      const uint32_t hi = rp[j + 1];
      const uint32_t sw = __funnelshift_r(lo, hi, sh);
      lo = hi;
#pragma unroll
      for (int jy = 0; jy < KY; jy++) {
        if (r - jy >= 0 && r - jy < HN) {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int i = j - k;
            if (i >= 0 && i < WW) acc[jy][k] = sad4_acc(sw, o[r - jy][i], acc[jy][k]);
          }
        }
      }
    }
  }
}

// KY original rows rolling through registers; STEADY iterations carry no masks
template <int WW, int KY, bool STEADY>
__device__ __forceinline__ void tile_rolling_rows(const uint8_t* refp, int ref_step, const uint8_t* orgp, int org_step, int r0,
                                                  int hn, uint32_t sh, uint32_t (&o)[KY][WW], uint32_t (&acc)[KY][4]) {
#pragma unroll
  for (int u = 0; u < KY; u++) {
    const int r = r0 + u;
    if (STEADY || r < hn + KY - 1) {
      if (STEADY || r < hn) {
        const uint32_t* op = reinterpret_cast<const uint32_t*>(orgp + r * org_step);
#pragma unroll
        for (int i = 0; i < WW; i++) o[u][i] = op[i];
      }
      const uint32_t* rp = reinterpret_cast<const uint32_t*>(refp + r * ref_step);
      uint32_t lo = rp[0];
#pragma unroll
      for (int j = 0; j < WW + 3; j++) {
        const uint32_t hi = rp[j + 1];
        const uint32_t sw = __funnelshift_r(lo, hi, sh);
        lo = hi;
#pragma unroll
        for (int jy = 0; jy < KY; jy++) {
          const int slot = (u - jy + KY) % KY;
          if (STEADY || (unsigned)(r - jy) < (unsigned)hn) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const int i = j - k;
              if (i >= 0 && i < WW) acc[jy][k] = sad4_acc(sw, o[slot][i], acc[jy][k]);
            }
          }
        }
      }
    }
  }
}

template <int WW, int KY>
__device__ __forceinline__ void tile_rolling(const uint8_t* refp, int ref_step, const uint8_t* orgp, int org_step, int hn,
                                             uint32_t sh, uint32_t (&acc)[KY][4]) {
  uint32_t o[KY][WW];
  for (int r0 = 0; r0 < hn + KY - 1; r0 += KY) {
    if (r0 >= KY && r0 + KY <= hn) tile_rolling_rows<WW, KY, true>(refp, ref_step, orgp, org_step, r0, hn, sh, o, acc);
    else                           tile_rolling_rows<WW, KY, false>(refp, ref_step, orgp, org_step, r0, hn, sh, o, acc);
  }
}

// one candidate, any shape (column mode: the leftover nx % 16 columns, < 1 % of a 129-wide window)
__device__ __forceinline__ uint32_t sad_candidate(const uint8_t* refp, int ref_step, const uint8_t* orgp, int org_step, int ww,
                                                  int hn, uint32_t sh) {
  uint32_t s = 0;
  for (int r = 0; r < hn; r++) {
    const uint32_t* rp = reinterpret_cast<const uint32_t*>(refp + r * ref_step);
    const uint32_t* op = reinterpret_cast<const uint32_t*>(orgp + r * org_step);
    uint32_t lo = rp[0];
    for (int i = 0; i < ww; i++) {
      const uint32_t hi = rp[i + 1];
      s = sad4_acc(__funnelshift_r(lo, hi, sh), op[i], s);
      lo = hi;
    }
  }
  return s;
}

// variant ids: static tiles 0..7, rolling tiles 8..15 (host table s8_pick_variant must match)
enum : int { S8V_S_1_8 = 0, S8V_S_2_4, S8V_S_2_8, S8V_S_3_8, S8V_S_4_4, S8V_S_4_6, S8V_S_4_8, S8V_S_8_8,
             S8V_R_1, S8V_R_2, S8V_R_3, S8V_R_4, S8V_R_6, S8V_R_8, S8V_R_12, S8V_R_16, S8V_COUNT };

struct S8VariantInfo { int ky; };
inline void s8_pick_variant(int ww, int hn, int* variant, int* ky) {
  struct E { int ww, hn, v, ky; };
  static const E stat[] = { {1, 8, S8V_S_1_8, 8}, {2, 4, S8V_S_2_4, 8}, {2, 8, S8V_S_2_8, 8}, {3, 8, S8V_S_3_8, 8},
                            {4, 4, S8V_S_4_4, 8}, {4, 6, S8V_S_4_6, 8}, {4, 8, S8V_S_4_8, 8}, {8, 8, S8V_S_8_8, 4} };
  for (const E& e : stat) if (e.ww == ww && e.hn == hn) { *variant = e.v; *ky = e.ky; return; }
  switch (ww) {
    case 1:  *variant = S8V_R_1;  *ky = 8; break;
    case 2:  *variant = S8V_R_2;  *ky = 8; break;
    case 3:  *variant = S8V_R_3;  *ky = 8; break;
    case 4:  *variant = S8V_R_4;  *ky = 8; break;
    case 6:  *variant = S8V_R_6;  *ky = 4; break;
    case 8:  *variant = S8V_R_8;  *ky = 4; break;
    case 12: *variant = S8V_R_12; *ky = 2; break;
    default: *variant = S8V_R_16; *ky = 2; break;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// epilogue of one lane tile: MV cost + running argmin.  KX columns 4 apart starting at candidate column cxi0; KY rows
// rstep apart starting at candidate row cyi0.
// ---------------------------------------------------------------------------------------------------------------
template <int KY, int KX>
__device__ __forceinline__ void tile_epilogue(const S8Job& jd, const uint32_t (&acc)[KY][4], int cxi0, int cyi0, int rstep,
                                              unsigned long long& best) {
  uint32_t px[KX];
#pragma unroll
  for (int k = 0; k < KX; k++) px[k] = jd.lambda * eg_bits(((jd.lt_x + cxi0 + 4 * k) << 2) - jd.pred_x);
#pragma unroll
  for (int jy = 0; jy < KY; jy++) {
    const int cyi = cyi0 + jy * rstep;
    if (cyi < jd.ny) {
      const uint32_t py = jd.lambda * eg_bits(((jd.lt_y + cyi) << 2) - jd.pred_y);
      const uint32_t row_idx = (uint32_t)(cyi * jd.nx + cxi0);
#pragma unroll
      for (int k = 0; k < KX; k++) {
        const uint32_t cost = (acc[jy][k] << jd.ss) + ((px[k] + py) >> 16);      // TComRdCost.h:177, u32 wrap as UInt
        const unsigned long long key = make_key(cost, row_idx + 4 * k);
        best = key < best ? key : best;
      }
    }
  }
}

template <int KY>
__device__ __forceinline__ void zero_acc(uint32_t (&acc)[KY][4]) {
#pragma unroll
  for (int a = 0; a < KY; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0;
}

// One lane tile = SAD accumulation + epilogue.
struct S8TileArgs {
  const uint8_t* refp; const uint8_t* orgp;
  int ref_step, org_step;
  uint32_t sh;
  int cxi0, cyi0, rstep;
};
template <int WW, int HN, int KY>
__device__ __forceinline__ unsigned long long run_static(const S8Job& jd, const S8TileArgs& a, unsigned long long best) {
  uint32_t acc[KY][4]; zero_acc<KY>(acc);
  tile_static<WW, HN, KY>(a.refp, a.ref_step, a.orgp, a.org_step, a.sh, acc);
  tile_epilogue<KY, 4>(jd, acc, a.cxi0, a.cyi0, a.rstep, best);
  return best;
}
template <int WW, int KY>
__device__ __forceinline__ unsigned long long run_rolling(const S8Job& jd, const S8TileArgs& a, unsigned long long best) {
  uint32_t acc[KY][4]; zero_acc<KY>(acc);
  tile_rolling<WW, KY>(a.refp, a.ref_step, a.orgp, a.org_step, jd.hn, a.sh, acc);
  tile_epilogue<KY, 4>(jd, acc, a.cxi0, a.cyi0, a.rstep, best);
  return best;
}

// One instantiation per tile variant: a unit only holds PUs of one variant.  (A single kernel switching between the
// sixteen tiles made ptxas keep the rolling tiles' register arrays live together: 254 registers against <= 96 for any
// tile alone.)
template <int V>
__global__ void __launch_bounds__(S8_THREADS, 2)
k_search8(const S8Unit* __restrict__ units, const S8Job* __restrict__ jobs, unsigned long long* __restrict__ keys,
          DevPlane cur_plane, DevPlane ref_plane) {
  extern __shared__ __align__(128) uint8_t s8_smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_next;
  __shared__ S8Job s_jd[S8_WARPS];        // each warp's current PU descriptor (kept out of registers)

  const S8Unit un = units[blockIdx.x];
  uint8_t* s_ref = s8_smem;
  uint8_t* s_org = s8_smem + un.org_smem_off;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) { mbar_init(&s_bar, 1); s_next = un.item_first; }
  __syncthreads();                                      // the initialised barrier is visible before its first use
  if (threadIdx.x == 0) mbar_expect_tx(&s_bar, (uint32_t)(un.ref_pitch * un.ref_rows + un.org_pitch * un.org_rows));
  __syncthreads();                                      // ... and armed before any copy can complete on it
  {
    // stage the window and the original tile: one bulk-async copy per row, issued by all threads, one mbarrier
    const uint8_t* gref = reinterpret_cast<const uint8_t*>(ref_plane.base) +
                          (size_t)(un.ref_by + ref_plane.margin_y) * ref_plane.pitch + (un.ref_bx + ref_plane.margin_x);
    for (int r = threadIdx.x; r < un.ref_rows; r += S8_THREADS)
      bulk_g2s(s_ref + r * un.ref_pitch, gref + (size_t)r * ref_plane.pitch, (uint32_t)un.ref_pitch, &s_bar);
    const uint8_t* gorg = reinterpret_cast<const uint8_t*>(cur_plane.base) +
                          (size_t)(un.org_by + cur_plane.margin_y) * cur_plane.pitch + (un.org_bx + cur_plane.margin_x);
    for (int r = (int)threadIdx.x - 128; r < un.org_rows; r += S8_THREADS)
      if (r >= 0) bulk_g2s(s_org + r * un.org_pitch, gorg + (size_t)r * cur_plane.pitch, (uint32_t)un.org_pitch, &s_bar);
  }
  mbar_wait(&s_bar, 0);

  int jslot = un.job_first;
  S8Job& jd = s_jd[warp];
  if (lane < (int)(sizeof(S8Job) / 4)) reinterpret_cast<int32_t*>(&jd)[lane] = reinterpret_cast<const int32_t*>(&jobs[jslot])[lane];
  __syncwarp();
  unsigned long long best = ~0ull;

  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(&s_next, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= un.item_last) break;
    if (item >= jd.item_start + jd.n_items) {
      // this warp moves on to another PU: publish its minimum for the old one
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
      }
      if (lane == 0 && best != ~0ull) atomicMin(&keys[jd.out_idx], best);
      best = ~0ull;
      do { jslot++; } while (item >= jobs[jslot].item_start + jobs[jslot].n_items);
      __syncwarp();
      if (lane < (int)(sizeof(S8Job) / 4)) reinterpret_cast<int32_t*>(&jd)[lane] = reinterpret_cast<const int32_t*>(&jobs[jslot])[lane];
      __syncwarp();
    }
    const int t = item - jd.item_start;
    const int n_blk = jd.nx >> 4;
    const int ref_step = un.ref_pitch << jd.ss, org_step = un.org_pitch << jd.ss;
    const int rstep = 1 << jd.ss;
    const uint8_t* orgp = s_org + jd.org_off;

    if (t < jd.n_qitems) {
      const int q = t * 8 + (lane >> 2);
      if (q < n_blk * jd.n_g) {
        const int g = q / n_blk, blk = q - g * n_blk;
        // first candidate row of the group: parity-0 groups first, then parity 1 (ss == 1); plain groups when ss == 0
        const int cyi0 = (g < jd.n_g0) ? (g * jd.ky) << jd.ss : (((g - jd.n_g0) * jd.ky) << 1) + 1;
        const int cxi0 = blk * 16 + (lane & 3);
        const int off = jd.win_off + cyi0 * un.ref_pitch + cxi0;
        const S8TileArgs ta{s_ref + (off & ~3), orgp, ref_step, org_step, (uint32_t)(off & 3) * 8u, cxi0, cyi0, rstep};
        if constexpr (V == S8V_S_1_8) best = run_static<1, 8, 8>(jd, ta, best);
        else if constexpr (V == S8V_S_2_4) best = run_static<2, 4, 8>(jd, ta, best);
        else if constexpr (V == S8V_S_2_8) best = run_static<2, 8, 8>(jd, ta, best);
        else if constexpr (V == S8V_S_3_8) best = run_static<3, 8, 8>(jd, ta, best);
        else if constexpr (V == S8V_S_4_4) best = run_static<4, 4, 8>(jd, ta, best);
        else if constexpr (V == S8V_S_4_6) best = run_static<4, 6, 8>(jd, ta, best);
        else if constexpr (V == S8V_S_4_8) best = run_static<4, 8, 8>(jd, ta, best);
        else if constexpr (V == S8V_S_8_8) best = run_static<8, 8, 4>(jd, ta, best);
        else if constexpr (V == S8V_R_1)   best = run_rolling<1, 8>(jd, ta, best);
        else if constexpr (V == S8V_R_2)   best = run_rolling<2, 8>(jd, ta, best);
        else if constexpr (V == S8V_R_3)   best = run_rolling<3, 8>(jd, ta, best);
        else if constexpr (V == S8V_R_4)   best = run_rolling<4, 8>(jd, ta, best);
        else if constexpr (V == S8V_R_6)   best = run_rolling<6, 4>(jd, ta, best);
        else if constexpr (V == S8V_R_8)   best = run_rolling<8, 4>(jd, ta, best);
        else if constexpr (V == S8V_R_12)  best = run_rolling<12, 2>(jd, ta, best);
        else                               best = run_rolling<16, 2>(jd, ta, best);
      }
    } else {
      // column mode: lane = (leftover column, row group)
      const int n_rem = jd.nx & 15;
      const int e = (t - jd.n_qitems) * 32 + lane;
      if (e < n_rem * jd.n_g) {
        const int g = e / n_rem, c = e - g * n_rem;
        const int cyi0 = (g < jd.n_g0) ? (g * jd.ky) << jd.ss : (((g - jd.n_g0) * jd.ky) << 1) + 1;
        const int cxi = n_blk * 16 + c;
        const uint32_t px = jd.lambda * eg_bits(((jd.lt_x + cxi) << 2) - jd.pred_x);
        for (int jy = 0; jy < jd.ky; jy++) {
          const int cyi = cyi0 + jy * rstep;
          if (cyi >= jd.ny) break;
          const int off = jd.win_off + cyi * un.ref_pitch + cxi;
          const uint32_t sad = sad_candidate(s_ref + (off & ~3), ref_step, orgp, org_step, jd.ww, jd.hn, (uint32_t)(off & 3) * 8u);
          const uint32_t py = jd.lambda * eg_bits(((jd.lt_y + cyi) << 2) - jd.pred_y);
          const unsigned long long key = make_key((sad << jd.ss) + ((px + py) >> 16), (uint32_t)(cyi * jd.nx + cxi));
          best = key < best ? key : best;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if (lane == 0 && best != ~0ull) atomicMin(&keys[jd.out_idx], best);
}

// keys -> (rcMv, ruiSAD): TEncSearch.cpp:3839-3841.  Keys left at ~0 belong to PUs searched by the generic kernel.
__global__ void k_search8_finalize(const SearchTask* __restrict__ tasks, const unsigned long long* __restrict__ keys,
                                   hmb200_pu_result* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = keys[i];
  if (key == ~0ull) return;
  const SearchTask t = tasks[i];
  const int nx = t.rb_x - t.lt_x + 1;
  const uint32_t idx = (uint32_t)(key & 0xffffffffu), cost = (uint32_t)(key >> 32);
  const int cy = idx / nx, cx = idx - cy * nx;
  const int x = t.lt_x + cx, y = t.lt_y + cy;
  out[i].mv_x = x;
  out[i].mv_y = y;
  out[i].sad = cost - mv_cost(t.lambda_cost, mv_bits(x, y, t.pred_x, t.pred_y, 2));
}

// ---------------------------------------------------------------------------------------------------------------
// host side: schedule
// ---------------------------------------------------------------------------------------------------------------
struct Search8Schedule {
  int n_units = 0, n_jobs = 0, n_leftover = 0;
  int smem_bytes = 0;
  S8Unit* d_units = nullptr;
  S8Job* d_jobs = nullptr;
  int* d_leftover = nullptr;          // task indices the tiled kernel does not cover (generic kernel)
  unsigned long long* d_keys = nullptr;
  int n_tasks = 0;
  unsigned long long executed_abs_diffs = 0;     // byte abs-diffs the per-PU tiles execute (== algorithmic for these)
  unsigned long long unique_abs_diffs = 0;       // same thing: a PU searched alone shares nothing
  int unit_first[S8V_COUNT] = {0}, unit_count[S8V_COUNT] = {0};   // units are bucketed by tile variant
  int smem_of[S8V_COUNT] = {0};
  // footprint of the scheduled jobs in picture coordinates (validated against the planes at run time)
  int min_x = 0, min_y = 0, max_x = 0, max_y = 0;       // reference samples touched: [min, max)
  int omin_x = 0, omin_y = 0, omax_x = 0, omax_y = 0;   // original samples touched
};

typedef void (*S8Kernel)(const S8Unit*, const S8Job*, unsigned long long*, DevPlane, DevPlane);
template <int V> struct S8Table { static void fill(S8Kernel* t) { t[V] = k_search8<V>; S8Table<V + 1>::fill(t); } };
template <> struct S8Table<S8V_COUNT> { static void fill(S8Kernel*) {} };
inline const S8Kernel* search8_kernels() {
  static S8Kernel table[S8V_COUNT];
  static bool filled = false;
  if (!filled) { S8Table<0>::fill(table); filled = true; }
  return table;
}

inline int  search8_configure(std::string* err) {
  const S8Kernel* k = search8_kernels();
  for (int v = 0; v < S8V_COUNT; v++) {
    cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(k[v]), cudaFuncAttributeMaxDynamicSharedMemorySize, S8_SMEM_MAX);
    if (e != cudaSuccess) { if (err) *err = std::string("cudaFuncSetAttribute(k_search8): ") + cudaGetErrorString(e); return HMB200_ERR_CUDA; }
  }
  return HMB200_OK;
}

inline void search8_free_schedule(Search8Schedule* s) {
  if (s->d_units) cudaFree(s->d_units);
  if (s->d_jobs) cudaFree(s->d_jobs);
  if (s->d_leftover) cudaFree(s->d_leftover);
  if (s->d_keys) cudaFree(s->d_keys);
  *s = Search8Schedule();
}

inline int s8_floor16(int v) { return v & ~15; }
inline int s8_ceil16(int v) { return (v + 15) & ~15; }
struct S8Box { int x0, y0, x1, y1; };
inline S8Box s8_union(const S8Box& a, const S8Box& b) { return S8Box{std::min(a.x0, b.x0), std::min(a.y0, b.y0), std::max(a.x1, b.x1), std::max(a.y1, b.y1)}; }
inline int s8_fl(int x) { return s8_floor16(x + 4096) - 4096; }      // 16-aligned in picture coordinates
inline int s8_ce(int x) { return s8_ceil16(x + 4096) - 4096; }
// dynamic shared memory of a group: window rows (+ slack) then the original tile, 128-byte aligned
inline int s8_smem_need(const S8Box& rb, const S8Box& ob, int* org_off, int bps = 1) {
  const int rp = (s8_ce(rb.x1) - s8_fl(rb.x0)) * bps, rr = rb.y1 - rb.y0 + S8_SLACK_ROWS;
  const int op = (s8_ce(ob.x1) - s8_fl(ob.x0)) * bps, orr = ob.y1 - ob.y0;
  const int ro = ((rp * rr + 16) + 127) & ~127;
  if (org_off) *org_off = ro;
  return ro + op * orr + 16;
}

// Window origins are multiples of 16 in picture coordinates: 16-byte aligned in the buffer when margin_x % 16 == 0.
// skip[i] != 0: task i is searched elsewhere (CU-fused kernel) and gets neither a job nor a leftover entry.
// tiled == false: no per-PU tiles at all (16-bit planes): every task that is not skipped becomes a leftover.
inline bool search8_build_schedule(const std::vector<SearchTask>& tasks, const std::vector<char>& skip, int sm_count, cudaStream_t stream,
                                   Search8Schedule* out, std::string* err, bool tiled = true) {
  typedef S8Box Box;
  auto uni = [](const Box& a, const Box& b) { return s8_union(a, b); };
  auto fl = [](int x) { return s8_fl(x); };
  auto ce = [](int x) { return s8_ce(x); };
  auto smem_need = [](const Box& rb, const Box& ob, int* org_off) { return s8_smem_need(rb, ob, org_off); };
  const int n = (int)tasks.size();
  std::vector<int> elig, left;
  std::vector<Box> rbox(n), obox(n);
  for (int i = 0; i < n; i++) {
    if (!skip.empty() && skip[i]) continue;
    const SearchTask& t = tasks[i];
    const int nx = t.rb_x - t.lt_x + 1, ny = t.rb_y - t.lt_y + 1;
    rbox[i] = Box{t.ref_x + t.lt_x, t.ref_y + t.lt_y, t.ref_x + t.rb_x + t.w, t.ref_y + t.rb_y + t.h};
    obox[i] = Box{t.org_x, t.org_y, t.org_x + t.w, t.org_y + t.h};
    const bool shape_ok = (t.w % 4 == 0) && t.w >= 4 && t.w <= 64 && t.h >= 1 && t.h <= 64 && (t.sub_shift == 0 || t.h % 2 == 0) &&
                          (t.w == 4 || t.w == 8 || t.w == 12 || t.w == 16 || t.w == 24 || t.w == 32 || t.w == 48 || t.w == 64);
    if (tiled && shape_ok && nx >= 1 && ny >= 1 && smem_need(rbox[i], obox[i], nullptr) <= S8_SMEM_MAX) elig.push_back(i);
    else left.push_back(i);
  }
  // group: stable sort by CTU, then greedy extension while the staged regions still let two CTAs share an SM
  std::stable_sort(elig.begin(), elig.end(), [&](int a, int b) {
    const SearchTask &ta = tasks[a], &tb = tasks[b];
    const int ka = (ta.org_y >> 6), kb = (tb.org_y >> 6);
    if (ka != kb) return ka < kb;
    return (ta.org_x >> 6) < (tb.org_x >> 6);
  });
  struct Group { int first, count; Box rb, ob; };
  std::vector<Group> groups;
  for (size_t p = 0; p < elig.size(); p++) {
    const int i = elig[p];
    if (!groups.empty()) {
      Group& g = groups.back();
      const Box nr = uni(g.rb, rbox[i]), no = uni(g.ob, obox[i]);
      if (g.count < 4096 && no.x1 - no.x0 <= 128 && no.y1 - no.y0 <= 128 && smem_need(nr, no, nullptr) <= S8_SMEM_SHARED2) {
        g.rb = nr; g.ob = no; g.count++;
        continue;
      }
    }
    groups.push_back(Group{(int)p, 1, rbox[i], obox[i]});
  }
  // per group: order PUs by descending tile cost, build job descriptors + warp-item prefix; then cut units
  std::vector<S8Job> jobs; jobs.reserve(elig.size());
  std::vector<S8Unit> units;
  struct GroupItems { int job_first, job_count; long long cost; };
  std::vector<GroupItems> gi(groups.size());
  std::vector<long long> job_item_cost; job_item_cost.reserve(elig.size());
  long long total_cost = 0;
  int smem_max = 0;
  Box all_r{1 << 30, 1 << 30, -(1 << 30), -(1 << 30)}, all_o = all_r;
  for (size_t gidx = 0; gidx < groups.size(); gidx++) {
    Group& g = groups[gidx];
    std::vector<int> ids(elig.begin() + g.first, elig.begin() + g.first + g.count);
    auto variant_of = [&](const SearchTask& t) { int v, ky; s8_pick_variant(t.w / 4, t.h >> t.sub_shift, &v, &ky); return v; };
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) {
      const SearchTask &ta = tasks[a], &tb = tasks[b];
      const int va = variant_of(ta), vb = variant_of(tb);
      if (va != vb) return va > vb;                 // wide rolling tiles first
      return ta.w * (ta.h >> ta.sub_shift) > tb.w * (tb.h >> tb.sub_shift);
    });
    const int rx0 = fl(g.rb.x0), ox0 = fl(g.ob.x0);
    const int rpitch = ce(g.rb.x1) - rx0, opitch = ce(g.ob.x1) - ox0;
    gi[gidx].job_first = (int)jobs.size(); gi[gidx].job_count = g.count; gi[gidx].cost = 0;
    int item = 0;
    for (int i : ids) {
      const SearchTask& t = tasks[i];
      S8Job j{};
      j.out_idx = i;
      j.org_off = (t.org_y - g.ob.y0) * opitch + (t.org_x - ox0);
      j.win_off = (t.ref_y + t.lt_y - g.rb.y0) * rpitch + (t.ref_x + t.lt_x - rx0);
      j.nx = t.rb_x - t.lt_x + 1; j.ny = t.rb_y - t.lt_y + 1;
      j.ww = t.w / 4; j.ss = t.sub_shift; j.hn = t.h >> t.sub_shift;
      j.lt_x = t.lt_x; j.lt_y = t.lt_y; j.pred_x = t.pred_x; j.pred_y = t.pred_y; j.lambda = t.lambda_cost;
      s8_pick_variant(j.ww, j.hn, &j.variant, &j.ky);
      if (j.ss == 0) { j.n_g0 = (j.ny + j.ky - 1) / j.ky; j.n_g = j.n_g0; }
      else {
        const int n0 = (j.ny + 1) / 2, n1 = j.ny / 2;
        j.n_g0 = (n0 + j.ky - 1) / j.ky; j.n_g = j.n_g0 + (n1 + j.ky - 1) / j.ky;
      }
      const int n_blk = j.nx / 16, n_rem = j.nx % 16;
      j.n_qitems = (n_blk * j.n_g + 7) / 8;
      j.n_items = j.n_qitems + (n_rem * j.n_g + 31) / 32;
      j.item_start = item;
      item += j.n_items;
      out->executed_abs_diffs += (unsigned long long)j.nx * j.ny * (unsigned long long)(4 * j.ww * j.hn);
      out->unique_abs_diffs += (unsigned long long)j.nx * j.ny * (unsigned long long)(4 * j.ww * j.hn);
      const long long c = (long long)(j.ww * j.hn + 6) * j.ky;
      job_item_cost.push_back(c);
      gi[gidx].cost += c * j.n_items;
      jobs.push_back(j);
    }
    total_cost += gi[gidx].cost;
    all_r = uni(all_r, g.rb); all_o = uni(all_o, g.ob);
  }
  // unit size per tile variant: every variant kernel gets ~12 units per resident CTA slot (2 per SM)
  long long variant_cost[S8V_COUNT] = {0}, variant_target[S8V_COUNT];
  for (size_t j = 0; j < jobs.size(); j++) variant_cost[jobs[j].variant] += job_item_cost[j] * jobs[j].n_items;
  for (int v = 0; v < S8V_COUNT; v++) variant_target[v] = std::max<long long>(variant_cost[v] / std::max(1, sm_count * 2 * 12), 4000);
  for (size_t gidx = 0; gidx < groups.size(); gidx++) {
    const Group& g = groups[gidx];
    const GroupItems& G = gi[gidx];
    S8Unit u{};
    const int rx0 = fl(g.rb.x0), ox0 = fl(g.ob.x0);
    u.ref_bx = rx0; u.ref_by = g.rb.y0;
    u.ref_pitch = ce(g.rb.x1) - rx0; u.ref_rows = g.rb.y1 - g.rb.y0;
    u.org_bx = ox0; u.org_by = g.ob.y0;
    u.org_pitch = ce(g.ob.x1) - ox0; u.org_rows = g.ob.y1 - g.ob.y0;
    int org_off = 0;
    const int need = smem_need(g.rb, g.ob, &org_off);
    u.org_smem_off = org_off;
    u.smem_need = need;
    smem_max = std::max(smem_max, need);
    // walk the item list, cutting at ~per cost and wherever the tile variant changes
    long long acc = 0;
    int ufirst_item = 0, ufirst_job = 0;
    auto emit = [&](int last_job_local, int item_last) {
      u.job_first = G.job_first + ufirst_job; u.item_first = ufirst_item; u.item_last = item_last;
      u.job_count = last_job_local - ufirst_job + 1;
      u.variant = jobs[u.job_first].variant;
      units.push_back(u);
    };
    for (int jl = 0; jl < G.job_count; jl++) {
      const S8Job& j = jobs[G.job_first + jl];
      const long long c = job_item_cost[G.job_first + jl];
      const long long per = variant_target[j.variant];
      if (acc > 0 && j.variant != jobs[G.job_first + jl - 1].variant) {
        emit(jl - 1, j.item_start);
        acc = 0; ufirst_item = j.item_start; ufirst_job = jl;
      }
      int done = 0;
      while (done < j.n_items) {
        const long long room = per - acc;
        const int take = (int)std::min<long long>(j.n_items - done, std::max<long long>(1, (room + c - 1) / c));
        acc += take * c; done += take;
        if (acc >= per) {
          emit(jl, j.item_start + done);
          acc = 0; ufirst_item = j.item_start + done; ufirst_job = (done == j.n_items) ? jl + 1 : jl;
        }
      }
    }
    if (acc > 0) {
      const S8Job& last = jobs[G.job_first + G.job_count - 1];
      emit(G.job_count - 1, last.item_start + last.n_items);
    }
  }
  // bucket the units by variant (stable: keeps the big-first order inside a bucket)
  std::stable_sort(units.begin(), units.end(), [](const S8Unit& a, const S8Unit& b) { return a.variant < b.variant; });
  for (size_t i = 0; i < units.size(); i++) {
    const int v = units[i].variant;
    if (out->unit_count[v]++ == 0) out->unit_first[v] = (int)i;
    out->smem_of[v] = std::max(out->smem_of[v], units[i].smem_need);
  }
  out->n_units = (int)units.size(); out->n_jobs = (int)jobs.size(); out->n_leftover = (int)left.size();
  out->smem_bytes = smem_max; out->n_tasks = n;
  if (!jobs.empty()) {
    out->min_x = fl(all_r.x0); out->min_y = all_r.y0; out->max_x = ce(all_r.x1); out->max_y = all_r.y1;
    out->omin_x = fl(all_o.x0); out->omin_y = all_o.y0; out->omax_x = ce(all_o.x1); out->omax_y = all_o.y1;
  }
  auto up = [&](void** d, const void* h, size_t bytes) {
    if (bytes == 0) return true;
    if (cudaMalloc(d, bytes) != cudaSuccess) return false;
    return cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, stream) == cudaSuccess;
  };
  bool ok = up((void**)&out->d_units, units.data(), units.size() * sizeof(S8Unit)) &&
            up((void**)&out->d_jobs, jobs.data(), jobs.size() * sizeof(S8Job)) &&
            up((void**)&out->d_leftover, left.data(), left.size() * sizeof(int)) &&
            (n == 0 || cudaMalloc((void**)&out->d_keys, (size_t)n * sizeof(unsigned long long)) == cudaSuccess) &&
            cudaStreamSynchronize(stream) == cudaSuccess;
  if (!ok) {
    if (err) *err = std::string("search8_build_schedule: ") + cudaGetErrorString(cudaGetLastError());
    search8_free_schedule(out);
    return false;
  }
  return true;
}

}  // namespace hmb200
