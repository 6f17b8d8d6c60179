// Batched quarter-pel refinement: TEncSearch::xPatternSearchFracDIF (TLibEncoder/TEncSearch.cpp:4240-4276) for many PUs.
//
// The reference builds whole-PU half/quarter-pel planes (xExtDIFUpSamplingH/Q, :5338-5539) and then evaluates nine
// candidates per stage with xGetHADs (TLibCommon/TComRdCost.cpp:1526-1593) or the SAD.  A candidate's distortion is a sum
// over independent Hadamard tiles (8x8 when both PU dimensions are multiples of 8, else 4x4; per-tile rounding), so
// here the unit of work is one (PU, candidate, tile): a thread interpolates exactly the samples its tile needs
// (separable 8-tap, 14-bit intermediates, TComInterpolationFilter.cpp:172-257), transforms and adds the rounded tile
// SATD to dist[PU][candidate] with one atomicAdd.  No shared memory, no block-level synchronisation, full occupancy for
// 4x8 PUs and 64x64 PUs alike.  Two stages (half, quarter) with a per-PU argmin kernel after each (xPatternRefinement,
// TEncSearch.cpp:808-861: candidate order of s_acMvRefineH/Q, strict '<', cost = dist + getCost at scale 1 / 0).
// The centre candidate of the quarter stage is the half stage's winner: its distortion is reused, not recomputed.
#pragma once
#include <string>
#include <vector>
#include "hmb200_device.cuh"
#include "hmb200_generic.cuh"

namespace hmb200 {

constexpr int FRAC_TILE_THREADS = 128;

// tile table entry: PU index (24 bits) | tile column (4 bits) << 24 | tile row (4 bits) << 28
__host__ __device__ __forceinline__ uint32_t frac_pack_tile(uint32_t pu, uint32_t tx, uint32_t ty) { return pu | (tx << 24) | (ty << 28); }

__device__ __forceinline__ void frac_load_taps(int f, int (&t)[8]) {
  // TComInterpolationFilter.cpp:57-63 (m_lumaFilter); f == 0 is the pass-through row {0,0,0,64,0,0,0,0}
  t[0] = (f == 1 || f == 2) ? -1 : 0;
  t[1] = (f == 0) ? 0 : (f == 3 ? 1 : 4);
  t[2] = (f == 0) ? 0 : (f == 1 ? -10 : (f == 2 ? -11 : -5));
  t[3] = (f == 0) ? 64 : (f == 1 ? 58 : (f == 2 ? 40 : 17));
  t[4] = (f == 0) ? 0 : (f == 1 ? 17 : (f == 2 ? 40 : 58));
  t[5] = (f == 0) ? 0 : (f == 1 ? -5 : (f == 2 ? -11 : -10));
  t[6] = (f == 0) ? 0 : (f == 3 ? 4 : (f == 2 ? 4 : 1));
  t[7] = (f == 2 || f == 3) ? -1 : 0;
}

// 11 consecutive samples of one row starting at p (any alignment)
template <typename T>
__device__ __forceinline__ void frac_load_row(const T* p, int (&px)[11]) {
  if constexpr (sizeof(T) == 1) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8u;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
    const uint32_t s0 = __funnelshift_r(w0, w1, sh), s1 = __funnelshift_r(w1, w2, sh), s2 = __funnelshift_r(w2, w3, sh);
#pragma unroll
    for (int i = 0; i < 4; i++) px[i] = (int)((s0 >> (8 * i)) & 0xffu);
#pragma unroll
    for (int i = 0; i < 4; i++) px[4 + i] = (int)((s1 >> (8 * i)) & 0xffu);
#pragma unroll
    for (int i = 0; i < 3; i++) px[8 + i] = (int)((s2 >> (8 * i)) & 0xffu);
  } else {
#pragma unroll
    for (int i = 0; i < 11; i++) px[i] = (int)__ldg(p + i);
  }
}

// Interpolated-minus-original differences of one N x N tile at quarter-pel displacement (fx, fy) (fractions; the
// integer parts are already folded into `ref`, which points at the tile's top-left reference sample).
// HSKIP / VSKIP: warp-uniform knowledge that fx == 0 / fy == 0 (pass-through taps), which removes the multiplies.
template <typename RefT, typename OrgT, int N, bool HSKIP, bool VSKIP>
__device__ __forceinline__ void frac_tile_diff(const RefT* ref, int ref_pitch, const OrgT* org, int org_pitch, int fx, int fy,
                                               int head, int maxv, int (&d)[N * N]) {
  int th[8], tv[8];
  frac_load_taps(fx, th);
  frac_load_taps(fy, tv);
  const int hshift = 6 - head, vshift = 6 + head;
  const int hoff = 8192 << hshift, voff = (1 << (vshift - 1)) + (8192 << 6);
  constexpr int R0 = VSKIP ? 3 : 0, R1 = VSKIP ? N + 3 : N + 7;          // rows of intermediates that are needed
#pragma unroll
  for (int cs = 0; cs < N; cs += 4) {                                    // four columns at a time: 8 x 4 live intermediates
    int hh[N + 7][4];
#pragma unroll
    for (int r = R0; r < R1; r++) {
      int px[11];
      frac_load_row<RefT>(ref + (ptrdiff_t)(r - 3) * ref_pitch + (cs - 3), px);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        int v;
        if (HSKIP) v = (int16_t)(px[c + 3] << head) - 8192;             // filterCopy isFirst (TComInterpolationFilter.cpp:113-124)
        else {
          int sum = 0;
#pragma unroll
          for (int t = 0; t < 8; t++) sum += px[c + t] * th[t];
          v = (int16_t)((sum - hoff) >> hshift);                         // filter<> isFirst: offset -8192 << shift (:196-251)
        }
        hh[r][c] = v;
      }
      if (r >= R1 - N) {                                                 // an output row is complete
        const int y = r - (R1 - N);
#pragma unroll
        for (int c = 0; c < 4; c++) {
          int val;
          if (VSKIP) val = (int16_t)((hh[y + 3][c] + 8192 + (1 << (head - 1))) >> head);     // filterCopy isLast (:126-141)
          else {
            int sum = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) sum += hh[y + t][c] * tv[t];
            val = (int16_t)((sum + voff) >> vshift);                     // filter<> isLast
          }
          val = min(max(val, 0), maxv);
          d[y * N + cs + c] = (int)org[(ptrdiff_t)y * org_pitch + cs + c] - val;
        }
      }
    }
  }
}

// ---- 8-bit planes: dot-product form ---------------------------------------------------------------------------------
// taps of one fraction as two packed signed-byte words (IDP.4A.U8.S8 operands)
__device__ __forceinline__ void frac_packed_taps(int f, int& lo, int& hi) {
  // {0,0,0,64 | 0,0,0,0}  {-1,4,-10,58 | 17,-5,1,0}  {-1,4,-11,40 | 40,-11,4,-1}  {0,1,-5,17 | 58,-10,4,-1}
  lo = (f == 0) ? 0x40000000 : (f == 1) ? 0x3AF604FF : (f == 2) ? 0x28F504FF : 0x11FB0100;
  hi = (f == 0) ? 0x00000000 : (f == 1) ? 0x0001FB11 : (f == 2) ? (int)0xFF04F528 : (int)0xFF04F63A;
}

// unsigned samples x signed taps (SASS IDP.4A.U8.S8)
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// Same result as frac_tile_diff for 8-bit samples, one code path for every fraction (the pass-through taps
// {0,0,0,64,0,0,0,0} give exactly filterCopy's values: 64*s >> 0 - 8192 and (64*(c + 8192) + 2^(5+head)) >> (6+head)).
// Horizontal pass: each output is two IDP.4A on byte-shifted words of the row; vertical pass: 8 IMAD.  Column strips of
// four share one rolled loop body, which keeps the kernel at ~20 KB of code (the fully specialised version thrashed
// the instruction cache: ncu 'no_instruction' was the top stall).
template <int N>
__device__ __forceinline__ void frac_tile_diff_u8(const uint8_t* ref, int ref_pitch, const uint8_t* org, int org_pitch, int fx, int fy,
                                                  int (&d)[N * N]) {
  constexpr int head = 6, vshift = 12;                                   // bit depth 8: headroom 14 - 8
  constexpr int voff = (1 << (vshift - 1)) + (8192 << 6);
  int t0, t1, tv[8];
  frac_packed_taps(fx, t0, t1);
  frac_load_taps(fy, tv);
  int dA[N][4], dB[N][4];
#pragma unroll 1
  for (int cs = 0; cs < N; cs += 4) {
    int hh[N + 7][4];
#pragma unroll
    for (int r = 0; r < N + 7; r++) {
      // 11 samples starting at column cs-3 of row r-3, as three words s0..s2 aligned to the first sample
      const uintptr_t a = reinterpret_cast<uintptr_t>(ref + (ptrdiff_t)(r - 3) * ref_pitch + (cs - 3));
      const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
      const uint32_t sh = (uint32_t)(a & 3) * 8u;
      const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
      const uint32_t s0 = __funnelshift_r(w0, w1, sh), s1 = __funnelshift_r(w1, w2, sh), s2 = __funnelshift_r(w2, w3, sh);
      // W[k] = samples k..k+3
      const uint32_t W0 = s0, W1 = __funnelshift_r(s0, s1, 8), W2 = __funnelshift_r(s0, s1, 16), W3 = __funnelshift_r(s0, s1, 24);
      const uint32_t W4 = s1, W5 = __funnelshift_r(s1, s2, 8), W6 = __funnelshift_r(s1, s2, 16), W7 = __funnelshift_r(s1, s2, 24);
      hh[r][0] = dp4a_us(W4, t1, dp4a_us(W0, t0, -8192));                  // filter<> isFirst at bit depth 8: shift 0, offset -8192
      hh[r][1] = dp4a_us(W5, t1, dp4a_us(W1, t0, -8192));
      hh[r][2] = dp4a_us(W6, t1, dp4a_us(W2, t0, -8192));
      hh[r][3] = dp4a_us(W7, t1, dp4a_us(W3, t0, -8192));
      if (r >= 7) {
        const int y = r - 7;
        const uint8_t* op = org + (ptrdiff_t)y * org_pitch + cs;
        const uint32_t ow = ((reinterpret_cast<uintptr_t>(op) & 3) == 0)
                                ? __ldg(reinterpret_cast<const uint32_t*>(op))
                                : ((uint32_t)op[0] | ((uint32_t)op[1] << 8) | ((uint32_t)op[2] << 16) | ((uint32_t)op[3] << 24));
#pragma unroll
        for (int c = 0; c < 4; c++) {
          int sum = voff;
#pragma unroll
          for (int t = 0; t < 8; t++) sum += hh[y + t][c] * tv[t];
          const int val = min(max(sum >> vshift, 0), 255);               // filter<> isLast + clip
          const int dv = (int)((ow >> (8 * c)) & 0xffu) - val;
          if (cs == 0) dA[y][c] = dv; else dB[y][c] = dv;
        }
      }
    }
  }
#pragma unroll
  for (int y = 0; y < N; y++)
#pragma unroll
    for (int c = 0; c < N; c++) d[y * N + c] = (c < 4) ? dA[y][c] : dB[y][c & 3];
  (void)head;
}

// one unique tile of the refinement (see "unique tiles" below): patch origin (tile - 4), original tile, half offset; 16 bytes
struct FracUTile { int16_t rx, ry, ox, oy; int16_t hx, hy; int16_t pad[2]; };

// stage 0: half-pel candidates around the integer MV; stage 1: quarter-pel candidates 1..8 around the best half.
// dist: [n_pu][9] accumulators (zeroed by the host; dist[.][0] of stage 1 is written by k_frac_argmin<0>).
template <typename RefT, typename OrgT, int N, bool HAD>
__global__ void __launch_bounds__(FRAC_TILE_THREADS, N == 8 ? 4 : 8)
k_frac_tiles(int stage, const SearchTask* __restrict__ tasks, const hmb200_pu_result* __restrict__ results,
             const uint32_t* __restrict__ tiles, int n_tiles, uint32_t* __restrict__ dist, DevPlane cur_plane, DevPlane ref_plane,
             const FracUTile* __restrict__ utiles = nullptr, const uint32_t* __restrict__ n_unique = nullptr) {
  const int nc = stage == 0 ? 9 : (stage == 1 ? 8 : 1);     // stage 2: one explicit quarter-pel MV per PU (motion compensation)
  const long long t = (long long)blockIdx.x * FRAC_TILE_THREADS + threadIdx.x;
  const int chunk = (int)(t / (32 * nc));
  const int within = (int)(t - (long long)chunk * 32 * nc);
  const int cand = within >> 5;                                          // warp-uniform
  const int slot = chunk * 32 + (within & 31);
  if (utiles) n_tiles = (int)*n_unique;                                  // unique tiles: the grid covers the worst case
  const bool active = slot < n_tiles;
  int fx = 0, fy = 0;
  uint32_t pu = 0;
  const RefT* ref = nullptr;
  const OrgT* org = nullptr;
  int ci = 0;
  if (active && utiles) {
    const FracUTile u = utiles[slot];
    pu = (uint32_t)slot;                                                 // row of the unique-tile distortion table
    int qx, qy;
    if (stage == 0) { ci = cand; qx = 2 * k_refine_h[ci][0]; qy = 2 * k_refine_h[ci][1]; }
    else            { ci = cand + 1; qx = 2 * u.hx + k_refine_q[ci][0]; qy = 2 * u.hy + k_refine_q[ci][1]; }
    fx = qx & 3; fy = qy & 3;
    ref = plane_at<RefT>(ref_plane, u.rx + 4 + (qx >> 2), u.ry + 4 + (qy >> 2));
    org = plane_at<OrgT>(cur_plane, u.ox, u.oy);
  } else if (active) {
    const uint32_t tile = tiles[slot];
    pu = tile & 0xffffffu;
    const int tx = (tile >> 24) & 15, ty = tile >> 28;
    const SearchTask tk = tasks[pu];
    const hmb200_pu_result rs = results[pu];
    int qx, qy;
    if (stage == 0)      { ci = cand; qx = 2 * k_refine_h[ci][0]; qy = 2 * k_refine_h[ci][1]; }
    else if (stage == 1) { ci = cand + 1; qx = 2 * rs.half_x + k_refine_q[ci][0]; qy = 2 * rs.half_y + k_refine_q[ci][1]; }
    else                 { ci = 0; qx = rs.qter_x; qy = rs.qter_y; }      // xPredInterBlk: the whole MV in quarter pel, mv_x/y = 0
    fx = qx & 3; fy = qy & 3;
    ref = plane_at<RefT>(ref_plane, tk.ref_x + rs.mv_x + (qx >> 2) + tx * N, tk.ref_y + rs.mv_y + (qy >> 2) + ty * N);
    org = plane_at<OrgT>(cur_plane, tk.org_x + tx * N, tk.org_y + ty * N);
  }
  const bool hskip = __all_sync(0xffffffffu, fx == 0), vskip = __all_sync(0xffffffffu, fy == 0);
  if (!active) return;
  const int bit_depth = ref_plane.bit_depth;
  const int head = max(2, 14 - bit_depth), maxv = (1 << bit_depth) - 1;
  int d[N * N];
  if constexpr (sizeof(RefT) == 1 && sizeof(OrgT) == 1) {
    (void)hskip; (void)vskip;
    frac_tile_diff_u8<N>(reinterpret_cast<const uint8_t*>(ref), ref_plane.pitch, reinterpret_cast<const uint8_t*>(org), cur_plane.pitch, fx, fy, d);
  } else
  if (hskip && vskip)  frac_tile_diff<RefT, OrgT, N, true, true>(ref, ref_plane.pitch, org, cur_plane.pitch, fx, fy, head, maxv, d);
  else if (hskip)      frac_tile_diff<RefT, OrgT, N, true, false>(ref, ref_plane.pitch, org, cur_plane.pitch, fx, fy, head, maxv, d);
  else if (vskip)      frac_tile_diff<RefT, OrgT, N, false, true>(ref, ref_plane.pitch, org, cur_plane.pitch, fx, fy, head, maxv, d);
  else                 frac_tile_diff<RefT, OrgT, N, false, false>(ref, ref_plane.pitch, org, cur_plane.pitch, fx, fy, head, maxv, d);
  uint32_t s;
  if (HAD) {
    if constexpr (N == 8) s = (had8x8_abs(d) + 2) >> 2;                  // TComRdCost.cpp:1520
    else                  s = (had4x4_abs(d) + 1) >> 1;                  // TComRdCost.cpp:1423
  } else {
    s = 0;
#pragma unroll
    for (int i = 0; i < N * N; i++) s += (uint32_t)abs(d[i]);
  }
  if (utiles) dist[(size_t)pu * 9 + ci] = s;                              // one thread per (unique tile, candidate)
  else        atomicAdd(&dist[(size_t)pu * 9 + ci], s);
}

// ---- unique tiles -------------------------------------------------------------------------------------------------------
// The PUs of a job list overlap: every sample of a CTU lies in 18 PUs whose Hadamard tiles are 8x8 (all partitions of the
// 64x64 and 32x32 CUs, three of the 16x16 CU's, the 8x8 PU) and in 6 whose tiles are 4x4, and those tiles sit on one 8 / 4
// sample grid because CU origins do.  Two PUs that found the same integer MV (and, for the quarter stage, the same half-pel
// offset) therefore ask for the SATD of exactly the same (original tile, interpolated tile) pairs.  The refinement works on
// UNIQUE tiles: a hash pass keys every tile instance by (original position, reference position incl. the integer MV, half
// offset), the SATD kernels run once per unique tile and candidate and store plain words, and a gather pass adds each
// instance's nine (eight) values to its PU.  Same integers, added in a different order - sums of uint32 are exact.
constexpr unsigned long long FRAC_KEY_EMPTY = ~0ull;

__device__ __forceinline__ uint32_t frac_hash64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (uint32_t)k;
}

// One thread per tile instance: inserts its key, the first inserter of a key allocates the dense unique id and writes the
// descriptor.  inst_slot[i] = hash slot of instance i (k_frac_gather turns it into the id after this kernel has finished).
template <int N>
__global__ void k_frac_unique(int stage, const SearchTask* __restrict__ tasks, const hmb200_pu_result* __restrict__ results,
                              const uint32_t* __restrict__ tiles, int n_tiles, unsigned long long* __restrict__ hkeys, uint32_t mask,
                              uint32_t* __restrict__ hval, uint32_t* __restrict__ count, FracUTile* __restrict__ utiles,
                              uint32_t* __restrict__ inst_slot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tiles) return;
  const uint32_t tile = tiles[i];
  const int pu = (int)(tile & 0xffffffu), tx = (tile >> 24) & 15, ty = tile >> 28;
  const SearchTask tk = tasks[pu];
  const hmb200_pu_result rs = results[pu];
  FracUTile u;
  u.rx = (int16_t)(tk.ref_x + rs.mv_x + tx * N - 4); u.ry = (int16_t)(tk.ref_y + rs.mv_y + ty * N - 4);
  u.ox = (int16_t)(tk.org_x + tx * N);               u.oy = (int16_t)(tk.org_y + ty * N);
  u.hx = (int16_t)(stage ? rs.half_x : 0);           u.hy = (int16_t)(stage ? rs.half_y : 0);
  u.pad[0] = u.pad[1] = 0;
  // 14 bits per coordinate (planes up to 8192 + margins around zero), 2 bits per half offset: 60 bits, never FRAC_KEY_EMPTY
  const unsigned long long key = ((unsigned long long)((uint32_t)(u.rx + 4096) & 0x3fffu)) | ((unsigned long long)((uint32_t)(u.ry + 4096) & 0x3fffu) << 14) |
                                 ((unsigned long long)((uint32_t)(u.ox + 4096) & 0x3fffu) << 28) | ((unsigned long long)((uint32_t)(u.oy + 4096) & 0x3fffu) << 42) |
                                 ((unsigned long long)((uint32_t)(u.hx + 1) & 3u) << 56) | ((unsigned long long)((uint32_t)(u.hy + 1) & 3u) << 58);
  uint32_t slot = frac_hash64(key) & mask;
  for (;;) {
    // most instances find their tile already there (4-8 instances per unique tile on coherent content): a plain load first,
    // the compare-and-swap only for slots that still look empty
    unsigned long long prev = *reinterpret_cast<volatile unsigned long long*>(&hkeys[slot]);
    if (prev == FRAC_KEY_EMPTY) prev = atomicCAS(&hkeys[slot], FRAC_KEY_EMPTY, key);
    if (prev == FRAC_KEY_EMPTY) {                        // first instance of this tile
      const uint32_t id = atomicAdd(count, 1u);
      hval[slot] = id;
      utiles[id] = u;
      break;
    }
    if (prev == key) break;
    slot = (slot + 1) & mask;
  }
  inst_slot[i] = slot;
}

// dist[pu][c] = sum over the PU's tile instances of udist[unique id][c]: one thread per (PU, candidate), plain stores.  The tile
// instances of a PU are contiguous in its tile table (frac_build_schedule emits them PU by PU): pu_first[pu] .. pu_first[pu + 1];
// a PU has tiles of one size only, so every (PU, candidate) is written by exactly one of the two launches.
__global__ void k_frac_gather(int stage, const uint32_t* __restrict__ pu_first, int n_pu, const uint32_t* __restrict__ inst_slot,
                              const uint32_t* __restrict__ hval, const uint32_t* __restrict__ udist, uint32_t* __restrict__ dist) {
  const int nc = stage == 0 ? 9 : 8;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int pu = (int)(t / nc), cand = (int)(t - (long long)pu * nc);
  if (pu >= n_pu) return;
  const uint32_t a = pu_first[pu], b = pu_first[pu + 1];
  if (a == b) return;                                  // this PU's tiles are of the other size
  const int ci = stage == 0 ? cand : cand + 1;
  uint32_t sum = 0;
  for (uint32_t i = a; i < b; i++) sum += udist[(size_t)hval[inst_slot[i]] * 9 + ci];
  dist[(size_t)pu * 9 + ci] = sum;
}

// ---- 8-bit planes, stages 0 / 1: reference patches staged in shared memory ------------------------------------------
// All candidates of a stage read the same 16 x 16 (8x8 tiles) or 12 x 16 (4x4 tiles) block of integer samples around
// the tile at the integer MV: columns -4..11, rows -4..N+3 (the candidates' integer parts are -1 or 0, the filter
// reaches 3 samples left / up and 4 right / down).  k_frac_tiles fetched those rows per (candidate, strip) with four
// unaligned global loads each and was bound by their latency (ncu: long_scoreboard the top stall, issue 0.55).  Here a
// CTA stages the patches of its 14 (half stage, 9 candidates) or 16 (quarter stage, 8 candidates) tiles once - one
// 16-byte row per thread and step, written with the patch's own alignment - and every (tile, candidate) thread reads
// whole rows with one LDS.128.  Arithmetic is frac_tile_diff_u8's, bit for bit.
__device__ __forceinline__ uint32_t frac_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 frac_lds128(uint32_t a) {
  uint4 v;
  asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t frac_lds32(uint32_t a) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}

// prow: shared address of the patch row that holds reference row (-3 + dy) of the tile (dy = integer part of the
// candidate's y offset); sh: 8 * (1 + dx), the byte phase of column (-3 + dx) inside the patch row; orgp: the tile's
// original samples, N bytes per row.
template <int N>
__device__ __forceinline__ void frac_tile_diff_patch(uint32_t prow, uint32_t sh, uint32_t orgp, int fx, int fy, int (&d)[N * N]) {
  constexpr int vshift = 12;                                             // bit depth 8: headroom 14 - 8
  constexpr int voff = (1 << (vshift - 1)) + (8192 << 6);
  int t0, t1, tv[8];
  frac_packed_taps(fx, t0, t1);
  frac_load_taps(fy, tv);
  int dA[N][4], dB[N][4];
#pragma unroll 1
  for (int cs = 0; cs < N; cs += 4) {
    int hh[N + 7][4];
#pragma unroll
    for (int r = 0; r < N + 7; r++) {
      const uint4 v = frac_lds128(prow + 16 * r);                        // patch columns 0..15 = tile columns -4..11
      const uint32_t a0 = cs ? v.y : v.x, a1 = cs ? v.z : v.y, a2 = cs ? v.w : v.z, a3 = 0u;   // the 11th sample ends in a2
      const uint32_t s0 = __funnelshift_r(a0, a1, sh), s1 = __funnelshift_r(a1, a2, sh), s2 = __funnelshift_r(a2, a3, sh);
      const uint32_t W0 = s0, W1 = __funnelshift_r(s0, s1, 8), W2 = __funnelshift_r(s0, s1, 16), W3 = __funnelshift_r(s0, s1, 24);
      const uint32_t W4 = s1, W5 = __funnelshift_r(s1, s2, 8), W6 = __funnelshift_r(s1, s2, 16), W7 = __funnelshift_r(s1, s2, 24);
      hh[r][0] = dp4a_us(W4, t1, dp4a_us(W0, t0, -8192));                  // filter<> isFirst at bit depth 8: shift 0, offset -8192
      hh[r][1] = dp4a_us(W5, t1, dp4a_us(W1, t0, -8192));
      hh[r][2] = dp4a_us(W6, t1, dp4a_us(W2, t0, -8192));
      hh[r][3] = dp4a_us(W7, t1, dp4a_us(W3, t0, -8192));
      if (r >= 7) {
        const int y = r - 7;
        const uint32_t ow = frac_lds32(orgp + y * N + cs);
#pragma unroll
        for (int c = 0; c < 4; c++) {
          int sum = voff;
#pragma unroll
          for (int t = 0; t < 8; t++) sum += hh[y + t][c] * tv[t];
          const int val = min(max(sum >> vshift, 0), 255);               // filter<> isLast + clip
          const int dv = (int)((ow >> (8 * c)) & 0xffu) - val;
          if (cs == 0) dA[y][c] = dv; else dB[y][c] = dv;
        }
      }
    }
  }
#pragma unroll
  for (int y = 0; y < N; y++)
#pragma unroll
    for (int c = 0; c < N; c++) d[y * N + c] = (c < 4) ? dA[y][c] : dB[y][c & 3];
}

template <int N> struct FracPatch {
  static constexpr int ROWS = N + 8;                                     // reference rows -4 .. N+3
  static constexpr int STRIDE = ROWS * 16 + 16;                          // 272 / 208 bytes: tiles of one warp fall on disjoint banks
  static constexpr int ORG = N * N;
  static constexpr int MAX_TILES = 16;
};

template <int N, bool HAD>
__global__ void __launch_bounds__(FRAC_TILE_THREADS, N == 8 ? 4 : 8)
k_frac_patch(int stage, const SearchTask* __restrict__ tasks, const hmb200_pu_result* __restrict__ results,
             const uint32_t* __restrict__ tiles, int n_tiles, uint32_t* __restrict__ dist, DevPlane cur_plane, DevPlane ref_plane,
             const FracUTile* __restrict__ utiles, const uint32_t* __restrict__ n_unique) {
  typedef FracPatch<N> P;
  __shared__ __align__(16) uint8_t s_patch[P::MAX_TILES * P::STRIDE];
  __shared__ __align__(16) uint8_t s_org[P::MAX_TILES * P::ORG];
  __shared__ int s_pu[P::MAX_TILES], s_rx[P::MAX_TILES], s_ry[P::MAX_TILES], s_ox[P::MAX_TILES], s_oy[P::MAX_TILES];
  __shared__ int s_hx[P::MAX_TILES], s_hy[P::MAX_TILES];
  const int nc = stage == 0 ? 9 : 8;
  const int tpc = FRAC_TILE_THREADS / nc;                                // tiles per CTA: 14 / 16
  const int slot0 = blockIdx.x * tpc;
  if (utiles) n_tiles = (int)*n_unique;                                  // unique tiles: the grid covers the worst case, the count is the device's
  const int nt = min(tpc, n_tiles - slot0);
  if (nt <= 0) return;
  if ((int)threadIdx.x < nt) {
    if (utiles) {
      const FracUTile u = utiles[slot0 + threadIdx.x];
      s_pu[threadIdx.x] = slot0 + (int)threadIdx.x;                      // row of the unique-tile distortion table
      s_rx[threadIdx.x] = u.rx; s_ry[threadIdx.x] = u.ry; s_ox[threadIdx.x] = u.ox; s_oy[threadIdx.x] = u.oy;
      s_hx[threadIdx.x] = u.hx; s_hy[threadIdx.x] = u.hy;
    } else {
      const uint32_t tile = tiles[slot0 + threadIdx.x];
      const int pu = (int)(tile & 0xffffffu), tx = (tile >> 24) & 15, ty = tile >> 28;
      const SearchTask tk = tasks[pu];
      const hmb200_pu_result rs = results[pu];
      s_pu[threadIdx.x] = pu;
      s_rx[threadIdx.x] = tk.ref_x + rs.mv_x + tx * N - 4; s_ry[threadIdx.x] = tk.ref_y + rs.mv_y + ty * N - 4;
      s_ox[threadIdx.x] = tk.org_x + tx * N;               s_oy[threadIdx.x] = tk.org_y + ty * N;
      s_hx[threadIdx.x] = rs.half_x;                       s_hy[threadIdx.x] = rs.half_y;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nt * P::ROWS; i += FRAC_TILE_THREADS) {   // one 16-byte patch row per step
    const int tl = i / P::ROWS, row = i - tl * P::ROWS;
    const uintptr_t a = reinterpret_cast<uintptr_t>(plane_at<uint8_t>(ref_plane, s_rx[tl], s_ry[tl] + row));
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8u;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3), w4 = __ldg(w + 4);
    uint4 o;
    o.x = __funnelshift_r(w0, w1, sh); o.y = __funnelshift_r(w1, w2, sh); o.z = __funnelshift_r(w2, w3, sh); o.w = __funnelshift_r(w3, w4, sh);
    *reinterpret_cast<uint4*>(s_patch + tl * P::STRIDE + row * 16) = o;
  }
  for (int i = threadIdx.x; i < nt * N; i += FRAC_TILE_THREADS) {          // original rows
    const int tl = i / N, row = i - tl * N;
    const uintptr_t a = reinterpret_cast<uintptr_t>(plane_at<uint8_t>(cur_plane, s_ox[tl], s_oy[tl] + row));   // any alignment
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8u;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1);
    if constexpr (N == 8) {
      const uint32_t w2 = __ldg(w + 2);
      *reinterpret_cast<uint2*>(s_org + tl * P::ORG + row * 8) = make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
    } else {
      *reinterpret_cast<uint32_t*>(s_org + tl * P::ORG + row * 4) = __funnelshift_r(w0, w1, sh);
    }
  }
  __syncthreads();
  const int tl = threadIdx.x / nc, cand = threadIdx.x - tl * nc;
  if (tl >= nt) return;
  int ci, qx, qy;
  if (stage == 0) { ci = cand; qx = 2 * k_refine_h[ci][0]; qy = 2 * k_refine_h[ci][1]; }
  else            { ci = cand + 1; qx = 2 * s_hx[tl] + k_refine_q[ci][0]; qy = 2 * s_hy[tl] + k_refine_q[ci][1]; }
  const int dx = qx >> 2, dy = qy >> 2;                                  // -1 or 0
  const uint32_t prow = frac_smem_addr(s_patch) + (uint32_t)(tl * P::STRIDE + (1 + dy) * 16);
  int d[N * N];
  frac_tile_diff_patch<N>(prow, (uint32_t)(1 + dx) * 8u, frac_smem_addr(s_org) + (uint32_t)(tl * P::ORG), qx & 3, qy & 3, d);
  uint32_t sv;
  if (HAD) {
    if constexpr (N == 8) sv = (had8x8_abs(d) + 2) >> 2;                 // TComRdCost.cpp:1520
    else                  sv = (had4x4_abs(d) + 1) >> 1;                 // TComRdCost.cpp:1423
  } else {
    sv = 0;
#pragma unroll
    for (int i = 0; i < N * N; i++) sv += (uint32_t)abs(d[i]);
  }
  if (utiles) dist[(size_t)s_pu[tl] * 9 + ci] = sv;                     // one thread per (unique tile, candidate): plain store
  else        atomicAdd(&dist[(size_t)s_pu[tl] * 9 + ci], sv);
}

// ---- 8-bit planes, stages 0 / 1: horizontal pass shared by the candidates, vertical pass as 16-bit dot products ------
// The nine (eight) candidates of a stage use only three horizontal positions (stage 0: -1/2, 0, +1/2 around the integer
// MV; stage 1: -1/4, 0, +1/4 around the best half-pel), so the first filter pass (TComInterpolationFilter.cpp:172-257,
// isFirst) is computed once per (tile, horizontal position) for the rows -4..N+3 that the two integer row offsets need,
// by the whole CTA, from the staged patches (k_frac_patch's layout).  The 14-bit intermediates are stored as PAIRS of
// vertically adjacent rows in one 32-bit word - pairs (0,1),(2,3).. and pairs (1,2),(3,4).. - which is the operand
// layout of IDP.2A.S16.S8: the second pass (isLast) of an output is four dot products of two rows x two taps instead of
// eight IMADs.  Same integers as k_frac_tiles: the products and the 32-bit sum are exact either way.
template <int N> struct FracHV {
  typedef FracPatch<N> P;
  static constexpr int NPE = P::ROWS / 2;                 // pair rows (2i, 2i+1): 8 / 6
  static constexpr int NPO = P::ROWS / 2 - 1;             // pair rows (2i+1, 2i+2): 7 / 5
  static constexpr int PRB = N * 4;                       // bytes of one pair row (N columns)
  static constexpr int OOFF = NPE * PRB + 48;             // odd pairs behind the even ones, shifted by 12 banks
  static constexpr int VSTR = OOFF + NPO * PRB;           // one horizontal position: 528 / 224 bytes
  static constexpr int TSTR = 3 * VSTR + 64;              // one tile
};

__device__ __forceinline__ int dp2a_lo_ss(uint32_t a, int b, int c) {
  int d;
  asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi_ss(uint32_t a, int b, int c) {
  int d;
  asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// first-pass intermediates of patch row `prow` (shared address) for N columns at horizontal position (dx, fx)
template <int N>
__device__ __forceinline__ void frac_hrow(uint32_t prow, uint32_t sh, int t0, int t1, int (&h)[N]) {
  const uint4 v = frac_lds128(prow);
#pragma unroll
  for (int cs = 0; cs < N; cs += 4) {
    const uint32_t a0 = cs ? v.y : v.x, a1 = cs ? v.z : v.y, a2 = cs ? v.w : v.z;
    const uint32_t s0 = __funnelshift_r(a0, a1, sh), s1 = __funnelshift_r(a1, a2, sh), s2 = __funnelshift_r(a2, 0u, sh);
    const uint32_t W1 = __funnelshift_r(s0, s1, 8), W2 = __funnelshift_r(s0, s1, 16), W3 = __funnelshift_r(s0, s1, 24);
    const uint32_t W5 = __funnelshift_r(s1, s2, 8), W6 = __funnelshift_r(s1, s2, 16), W7 = __funnelshift_r(s1, s2, 24);
    h[cs + 0] = dp4a_us(s1, t1, dp4a_us(s0, t0, -8192));                 // filter<> isFirst at bit depth 8: shift 0, offset -8192
    h[cs + 1] = dp4a_us(W5, t1, dp4a_us(W1, t0, -8192));
    h[cs + 2] = dp4a_us(W6, t1, dp4a_us(W2, t0, -8192));
    h[cs + 3] = dp4a_us(W7, t1, dp4a_us(W3, t0, -8192));
  }
}

#ifndef HV8_MIN_CTAS
#define HV8_MIN_CTAS 5            // 96 registers, 5 CTAs per SM: 1.358 ms per 1080p frame pair vs 1.378 at 4 (112 registers) and 1.478 at 6 (80, spills)
#endif
template <int N, bool HAD>
__global__ void __launch_bounds__(FRAC_TILE_THREADS, N == 8 ? HV8_MIN_CTAS : 8)
k_frac_hv(int stage, const SearchTask* __restrict__ tasks, const hmb200_pu_result* __restrict__ results,
          const uint32_t* __restrict__ tiles, int n_tiles, uint32_t* __restrict__ dist, DevPlane cur_plane, DevPlane ref_plane,
          const FracUTile* __restrict__ utiles, const uint32_t* __restrict__ n_unique) {
  typedef FracPatch<N> P;
  typedef FracHV<N> H;
  __shared__ __align__(16) uint8_t s_patch[P::MAX_TILES * P::STRIDE];
  __shared__ __align__(16) uint8_t s_org[P::MAX_TILES * P::ORG];
  __shared__ __align__(16) uint8_t s_h[P::MAX_TILES * H::TSTR];
  __shared__ int s_pu[P::MAX_TILES], s_rx[P::MAX_TILES], s_ry[P::MAX_TILES], s_ox[P::MAX_TILES], s_oy[P::MAX_TILES];
  __shared__ int s_hx[P::MAX_TILES], s_hy[P::MAX_TILES];
  const int nc = stage == 0 ? 9 : 8;
  const int tpc = FRAC_TILE_THREADS / nc;                                // tiles per CTA: 14 / 16
  const int slot0 = blockIdx.x * tpc;
  if (utiles) n_tiles = (int)*n_unique;                                  // unique tiles: the grid covers the worst case, the count is the device's
  const int nt = min(tpc, n_tiles - slot0);
  if (nt <= 0) return;
  if ((int)threadIdx.x < nt) {
    if (utiles) {
      const FracUTile u = utiles[slot0 + threadIdx.x];
      s_pu[threadIdx.x] = slot0 + (int)threadIdx.x;                      // row of the unique-tile distortion table
      s_rx[threadIdx.x] = u.rx; s_ry[threadIdx.x] = u.ry; s_ox[threadIdx.x] = u.ox; s_oy[threadIdx.x] = u.oy;
      s_hx[threadIdx.x] = u.hx; s_hy[threadIdx.x] = u.hy;
    } else {
      const uint32_t tile = tiles[slot0 + threadIdx.x];
      const int pu = (int)(tile & 0xffffffu), tx = (tile >> 24) & 15, ty = tile >> 28;
      const SearchTask tk = tasks[pu];
      const hmb200_pu_result rs = results[pu];
      s_pu[threadIdx.x] = pu;
      s_rx[threadIdx.x] = tk.ref_x + rs.mv_x + tx * N - 4; s_ry[threadIdx.x] = tk.ref_y + rs.mv_y + ty * N - 4;
      s_ox[threadIdx.x] = tk.org_x + tx * N;               s_oy[threadIdx.x] = tk.org_y + ty * N;
      s_hx[threadIdx.x] = rs.half_x;                       s_hy[threadIdx.x] = rs.half_y;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nt * P::ROWS; i += FRAC_TILE_THREADS) {   // one 16-byte patch row per step
    const int tl = i / P::ROWS, row = i - tl * P::ROWS;
    const uintptr_t a = reinterpret_cast<uintptr_t>(plane_at<uint8_t>(ref_plane, s_rx[tl], s_ry[tl] + row));
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8u;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3), w4 = __ldg(w + 4);
    uint4 o;
    o.x = __funnelshift_r(w0, w1, sh); o.y = __funnelshift_r(w1, w2, sh); o.z = __funnelshift_r(w2, w3, sh); o.w = __funnelshift_r(w3, w4, sh);
    *reinterpret_cast<uint4*>(s_patch + tl * P::STRIDE + row * 16) = o;
  }
  for (int i = threadIdx.x; i < nt * N; i += FRAC_TILE_THREADS) {          // original rows
    const int tl = i / N, row = i - tl * N;
    const uintptr_t a = reinterpret_cast<uintptr_t>(plane_at<uint8_t>(cur_plane, s_ox[tl], s_oy[tl] + row));   // any alignment
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8u;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1);
    if constexpr (N == 8) {
      const uint32_t w2 = __ldg(w + 2);
      *reinterpret_cast<uint2*>(s_org + tl * P::ORG + row * 8) = make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
    } else {
      *reinterpret_cast<uint32_t*>(s_org + tl * P::ORG + row * 4) = __funnelshift_r(w0, w1, sh);
    }
  }
  __syncthreads();
  // first pass: unit = (tile, horizontal position j, pair index m) -> rows 2m, 2m+1, 2m+2 -> pair rows E[m], O[m]
  for (int u = threadIdx.x; u < nt * 3 * H::NPE; u += FRAC_TILE_THREADS) {
    const int tl = u / (3 * H::NPE), rem = u - tl * (3 * H::NPE), j = rem / H::NPE, m = rem - j * H::NPE;
    const int qx = stage == 0 ? 2 * (j - 1) : 2 * s_hx[tl] + (j - 1);
    int t0, t1;
    frac_packed_taps(qx & 3, t0, t1);
    const uint32_t sh = (uint32_t)(1 + (qx >> 2)) * 8u;
    const uint32_t prow = frac_smem_addr(s_patch) + (uint32_t)(tl * P::STRIDE + 2 * m * 16);
    int h0[N], h1[N], h2[N];
    frac_hrow<N>(prow, sh, t0, t1, h0);
    frac_hrow<N>(prow + 16, sh, t0, t1, h1);
    uint8_t* dst = s_h + tl * H::TSTR + j * H::VSTR + m * H::PRB;
    uint32_t e[N], o[N];
#pragma unroll
    for (int c = 0; c < N; c++) e[c] = __byte_perm((uint32_t)h0[c], (uint32_t)h1[c], 0x5410);
#pragma unroll
    for (int c = 0; c < N; c += 4) *reinterpret_cast<uint4*>(dst + 4 * c) = make_uint4(e[c], e[c + 1], e[c + 2], e[c + 3]);
    if (m < H::NPO) {
      frac_hrow<N>(prow + 32, sh, t0, t1, h2);
#pragma unroll
      for (int c = 0; c < N; c++) o[c] = __byte_perm((uint32_t)h1[c], (uint32_t)h2[c], 0x5410);
#pragma unroll
      for (int c = 0; c < N; c += 4) *reinterpret_cast<uint4*>(dst + H::OOFF + 4 * c) = make_uint4(o[c], o[c + 1], o[c + 2], o[c + 3]);
    }
  }
  __syncthreads();
  const int tl = threadIdx.x / nc, cand = threadIdx.x - tl * nc;
  if (tl >= nt) return;
  int ci, j, qy;
  if (stage == 0) { ci = cand; j = k_refine_h[ci][0] + 1; qy = 2 * k_refine_h[ci][1]; }
  else            { ci = cand + 1; j = k_refine_q[ci][0] + 1; qy = 2 * s_hy[tl] + k_refine_q[ci][1]; }
  const int dy = qy >> 2;                                                // -1 or 0
  int tl4, th4;
  frac_packed_taps(qy & 3, tl4, th4);
  // output row y needs intermediate rows y+dy+1 .. y+dy+8 (patch row numbering): even y -> E (dy = -1) or O (dy = 0)
  // from pair y/2; odd y -> O from pair (y-1)/2 (dy = -1) or E from pair (y+1)/2 (dy = 0)
  const uint32_t hb = frac_smem_addr(s_h) + (uint32_t)(tl * H::TSTR + j * H::VSTR);
  const uint32_t pe = hb + (dy ? 0u : (uint32_t)H::OOFF), po = hb + (dy ? (uint32_t)H::OOFF : (uint32_t)H::PRB);
  const uint32_t orgp = frac_smem_addr(s_org) + (uint32_t)(tl * P::ORG);
  constexpr int vshift = 12, voff = (1 << (vshift - 1)) + (8192 << 6);   // bit depth 8: headroom 14 - 8
  constexpr int NPR = N / 2 + 3;                                         // pair rows per parity that a tile touches
  int dA[N][4], dB[N][4];
#pragma unroll 1
  for (int cs = 0; cs < N; cs += 4) {
    uint4 PE[NPR], PO[NPR];
#pragma unroll
    for (int i = 0; i < NPR; i++) { PE[i] = frac_lds128(pe + i * H::PRB + cs * 4); PO[i] = frac_lds128(po + i * H::PRB + cs * 4); }
#pragma unroll
    for (int y = 0; y < N; y++) {
      const uint32_t ow = frac_lds32(orgp + y * N + cs);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        int sum = voff;
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const uint4 q = (y & 1) ? PO[(y >> 1) + t] : PE[(y >> 1) + t];
          const uint32_t w = c == 0 ? q.x : c == 1 ? q.y : c == 2 ? q.z : q.w;
          const int taps = t < 2 ? tl4 : th4;
          sum = (t & 1) ? dp2a_hi_ss(w, taps, sum) : dp2a_lo_ss(w, taps, sum);
        }
        const int val = min(max(sum >> vshift, 0), 255);                 // filter<> isLast + clip
        const int dv = (int)((ow >> (8 * c)) & 0xffu) - val;
        if (cs == 0) dA[y][c] = dv; else dB[y][c] = dv;
      }
    }
  }
  int d[N * N];
#pragma unroll
  for (int y = 0; y < N; y++)
#pragma unroll
    for (int c = 0; c < N; c++) d[y * N + c] = (c < 4) ? dA[y][c] : dB[y][c & 3];
  uint32_t sv;
  if (HAD) {
    if constexpr (N == 8) sv = (had8x8_abs(d) + 2) >> 2;                 // TComRdCost.cpp:1520
    else                  sv = (had4x4_abs(d) + 1) >> 1;                 // TComRdCost.cpp:1423
  } else {
    sv = 0;
#pragma unroll
    for (int i = 0; i < N * N; i++) sv += (uint32_t)abs(d[i]);
  }
  if (utiles) dist[(size_t)s_pu[tl] * 9 + ci] = sv;                     // one thread per (unique tile, candidate): plain store
  else        atomicAdd(&dist[(size_t)s_pu[tl] * 9 + ci], sv);
}

// xPatternRefinement's argmin (TEncSearch.cpp:808-861).  STAGE 0 writes rcMvHalf and seeds the quarter stage's
// centre distortion; STAGE 1 writes rcMvQter and ruiCost.
template <int STAGE>
__global__ void k_frac_argmin(const SearchTask* __restrict__ tasks, hmb200_pu_result* __restrict__ results,
                              const uint32_t* __restrict__ dist_in, uint32_t* __restrict__ dist_next, int n, int bit_depth) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const SearchTask t = tasks[i];
  hmb200_pu_result r = results[i];
  uint32_t best = 0xffffffffu;
  int bi = 0;
#pragma unroll
  for (int c = 0; c < 9; c++) {
    uint32_t bits;
    if (STAGE == 0) bits = mv_bits(k_refine_h[c][0] + 2 * r.mv_x, k_refine_h[c][1] + 2 * r.mv_y, t.pred_x, t.pred_y, 1);      // :3746
    else            bits = mv_bits(k_refine_q[c][0] + 4 * r.mv_x + 2 * r.half_x, k_refine_q[c][1] + 4 * r.mv_y + 2 * r.half_y,
                                   t.pred_x, t.pred_y, 0);                                                                    // :4267
    const uint32_t cost = (dist_in[(size_t)i * 9 + c] >> (bit_depth - 8)) + mv_cost(t.lambda_cost, bits);
    if (cost < best) { best = cost; bi = c; }
  }
  if (STAGE == 0) {
    r.half_x = k_refine_h[bi][0]; r.half_y = k_refine_h[bi][1];
    dist_next[(size_t)i * 9] = dist_in[(size_t)i * 9 + bi];              // the quarter stage's candidate 0 is this block
  } else {
    r.qter_x = k_refine_q[bi][0]; r.qter_y = k_refine_q[bi][1];
    r.frac_cost = best;
  }
  results[i] = r;
}

// dist[pu][0] >> (bitDepth - 8): the distortion of one motion-compensated prediction per PU (stage 2 of k_frac_tiles)
__global__ void k_mc_collect(const uint32_t* __restrict__ dist, uint32_t* __restrict__ out, int n, int bit_depth) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dist[(size_t)i * 9] >> (bit_depth - 8);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
struct FracSchedule {
  int n_pu = 0, n_tiles8 = 0, n_tiles4 = 0;
  uint32_t* d_tiles8 = nullptr;
  uint32_t* d_tiles4 = nullptr;
  uint32_t* d_dist = nullptr;        // [2][n_pu][9]
  // unique-tile pass (8-bit planes): per tile size a hash table (keys, dense ids), the unique descriptors, per-instance slots
  // and the unique tiles' distortions; d_count[2 * stage + (N == 4)]
  unsigned long long* d_hkeys[2] = {nullptr, nullptr};
  uint32_t* d_hval[2] = {nullptr, nullptr};
  uint32_t* d_inst[2] = {nullptr, nullptr};
  FracUTile* d_utiles[2] = {nullptr, nullptr};
  uint32_t* d_udist[2] = {nullptr, nullptr};
  uint32_t* d_count = nullptr;
  uint32_t* d_pu_first[2] = {nullptr, nullptr};   // [n_pu + 1] first tile instance of every PU in d_tiles8 / d_tiles4
  uint32_t hmask[2] = {0, 0};
};

inline void frac_free_schedule(FracSchedule* s) {
  if (s->d_tiles8) cudaFree(s->d_tiles8);
  if (s->d_tiles4) cudaFree(s->d_tiles4);
  if (s->d_dist) cudaFree(s->d_dist);
  for (int k = 0; k < 2; k++) {
    if (s->d_hkeys[k]) cudaFree(s->d_hkeys[k]);
    if (s->d_hval[k]) cudaFree(s->d_hval[k]);
    if (s->d_inst[k]) cudaFree(s->d_inst[k]);
    if (s->d_utiles[k]) cudaFree(s->d_utiles[k]);
    if (s->d_udist[k]) cudaFree(s->d_udist[k]);
    if (s->d_pu_first[k]) cudaFree(s->d_pu_first[k]);
  }
  if (s->d_count) cudaFree(s->d_count);
  *s = FracSchedule();
}

// Buffers of the unique-tile pass, allocated on first use (8-bit planes with more than a handful of tiles).
inline bool frac_alloc_unique(FracSchedule* fs) {
  if (fs->d_count) return true;
  const int n[2] = {fs->n_tiles8, fs->n_tiles4};
  bool ok = cudaMalloc((void**)&fs->d_count, 4 * sizeof(uint32_t)) == cudaSuccess;
  for (int k = 0; k < 2 && ok; k++) {
    if (n[k] == 0) continue;
    uint32_t cap = 1024;
    while (cap < 2u * (uint32_t)n[k]) cap <<= 1;
    fs->hmask[k] = cap - 1;
    ok = cudaMalloc((void**)&fs->d_hkeys[k], (size_t)cap * sizeof(unsigned long long)) == cudaSuccess &&
         cudaMalloc((void**)&fs->d_hval[k], (size_t)cap * sizeof(uint32_t)) == cudaSuccess &&
         cudaMalloc((void**)&fs->d_inst[k], (size_t)n[k] * sizeof(uint32_t)) == cudaSuccess &&
         cudaMalloc((void**)&fs->d_utiles[k], (size_t)n[k] * sizeof(FracUTile)) == cudaSuccess &&
         cudaMalloc((void**)&fs->d_udist[k], (size_t)n[k] * 9 * sizeof(uint32_t)) == cudaSuccess;
  }
  return ok;
}

inline bool frac_build_schedule(const std::vector<SearchTask>& tasks, cudaStream_t stream, FracSchedule* out, std::string* err) {
  const int n = (int)tasks.size();
  if (n >= (1 << 24)) { if (err) *err = "frac_build_schedule: more than 2^24 PUs in one batch"; return false; }
  std::vector<uint32_t> t8, t4, f8((size_t)n + 1), f4((size_t)n + 1);
  for (int i = 0; i < n; i++) {
    const SearchTask& t = tasks[i];
    const int e = (t.w % 8 == 0 && t.h % 8 == 0) ? 8 : 4;               // xGetHADs tile choice (TComRdCost.cpp:1544-1572)
    std::vector<uint32_t>& dst = (e == 8) ? t8 : t4;
    f8[(size_t)i] = (uint32_t)t8.size(); f4[(size_t)i] = (uint32_t)t4.size();
    for (int ty = 0; ty < t.h / e; ty++)
      for (int tx = 0; tx < t.w / e; tx++) dst.push_back(frac_pack_tile((uint32_t)i, (uint32_t)tx, (uint32_t)ty));
  }
  f8[(size_t)n] = (uint32_t)t8.size(); f4[(size_t)n] = (uint32_t)t4.size();
  out->n_pu = n; out->n_tiles8 = (int)t8.size(); out->n_tiles4 = (int)t4.size();
  auto up = [&](uint32_t** d, const std::vector<uint32_t>& h) {
    if (h.empty()) return true;
    if (cudaMalloc((void**)d, h.size() * sizeof(uint32_t)) != cudaSuccess) return false;
    return cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream) == cudaSuccess;
  };
  bool ok = up(&out->d_tiles8, t8) && up(&out->d_tiles4, t4) && (n == 0 || (up(&out->d_pu_first[0], f8) && up(&out->d_pu_first[1], f4))) &&
            (n == 0 || cudaMalloc((void**)&out->d_dist, (size_t)n * 18 * sizeof(uint32_t)) == cudaSuccess) &&
            cudaStreamSynchronize(stream) == cudaSuccess;
  if (!ok) {
    if (err) *err = std::string("frac_build_schedule: ") + cudaGetErrorString(cudaGetLastError());
    frac_free_schedule(out);
  }
  return ok;
}

// Enqueues the whole refinement on `stream`; returns the number of kernel launches, or -1 on a launch error.
template <typename RefT, typename OrgT>
inline int frac_launch(const FracSchedule& fs, const SearchTask* d_tasks, hmb200_pu_result* d_results, const DevPlane& cur,
                       const DevPlane& ref, bool use_had, cudaStream_t stream, cudaStream_t side = nullptr, cudaEvent_t ev_fork = nullptr,
                       cudaEvent_t ev_join = nullptr) {
  if (fs.n_pu == 0) return 0;
  int launches = 0;
  // unique-tile keys hold 14-bit coordinates (-4096 .. 12287): larger planes keep one SATD per tile instance
  const bool plane_fits_i16 = std::max(cur.width, ref.width) + 2 * std::max(cur.margin_x, ref.margin_x) < 12000 &&
                              std::max(cur.height, ref.height) + 2 * std::max(cur.margin_y, ref.margin_y) < 12000;
  uint32_t* dist0 = fs.d_dist;
  uint32_t* dist1 = fs.d_dist + (size_t)fs.n_pu * 9;
  if (cudaMemsetAsync(fs.d_dist, 0, (size_t)fs.n_pu * 18 * sizeof(uint32_t), stream) != cudaSuccess) return -1;
  for (int stage = 0; stage < 2; stage++) {
    uint32_t* dist = stage == 0 ? dist0 : dist1;
    const int nc = stage == 0 ? 9 : 8;
    auto blocks = [&](int n_tiles) { return (int)((((long long)(n_tiles + 31) / 32) * 32 * nc + FRAC_TILE_THREADS - 1) / FRAC_TILE_THREADS); };
    constexpr bool PATCH = sizeof(RefT) == 1 && sizeof(OrgT) == 1;        // 8-bit planes: shared-memory patches (k_frac_hv / k_frac_patch)
    const bool old_path = getenv("HMB200_FRAC_PATCH") != nullptr;         // A/B knob: per-candidate first pass for 8x8 tiles too
    const bool hv4 = getenv("HMB200_FRAC_HV4") != nullptr;                // 4x4 tiles: the shared first pass costs more in barriers than
                                                                          // it saves (ncu: 263 vs 225 us per stage), so it is opt-in
    const int tpc = FRAC_TILE_THREADS / nc;
    const bool no_unique = getenv("HMB200_NO_FRAC_DEDUPE") != nullptr;     // A/B knob: every tile instance on its own
    const bool unique = !no_unique && plane_fits_i16 && frac_alloc_unique(const_cast<FracSchedule*>(&fs));
    if (unique && stage == 0 && cudaMemsetAsync(fs.d_count, 0, 4 * sizeof(uint32_t), stream) != cudaSuccess) return -1;
    // unique-tile pass of one tile size (k = 0: 8x8 tiles, 1: 4x4): keys + descriptors before the SATD kernel, gather after it
    // The 8x8-tile and the 4x4-tile pipelines of a stage are independent (disjoint PUs, own hash tables) and each is a chain of
    // small kernels: with a side stream they run next to each other and meet again in front of the argmin.
    const bool two = side && ev_fork && ev_join && fs.n_tiles8 > 0 && fs.n_tiles4 > 0 && !getenv("HMB200_FRAC_ONE_STREAM");
    if (two && (cudaEventRecord(ev_fork, stream) != cudaSuccess || cudaStreamWaitEvent(side, ev_fork, 0) != cudaSuccess)) return -1;
    cudaStream_t st_k = stream;
    auto unique_begin = [&](int k) -> bool {
      const int n = k ? fs.n_tiles4 : fs.n_tiles8;
      if (cudaMemsetAsync(fs.d_hkeys[k], 0xff, ((size_t)fs.hmask[k] + 1) * sizeof(unsigned long long), st_k) != cudaSuccess) return false;
      if (k == 0) k_frac_unique<8><<<(n + 255) / 256, 256, 0, st_k>>>(stage, d_tasks, d_results, fs.d_tiles8, n, fs.d_hkeys[0], fs.hmask[0], fs.d_hval[0],
                                                                       fs.d_count + 2 * stage, fs.d_utiles[0], fs.d_inst[0]);
      else        k_frac_unique<4><<<(n + 255) / 256, 256, 0, st_k>>>(stage, d_tasks, d_results, fs.d_tiles4, n, fs.d_hkeys[1], fs.hmask[1], fs.d_hval[1],
                                                                       fs.d_count + 2 * stage + 1, fs.d_utiles[1], fs.d_inst[1]);
      launches++;
      return true;
    };
    auto unique_end = [&](int k) {
      const int n = k ? fs.n_tiles4 : fs.n_tiles8;
      (void)n;
      k_frac_gather<<<(int)(((long long)fs.n_pu * nc + 255) / 256), 256, 0, st_k>>>(stage, fs.d_pu_first[k], fs.n_pu, fs.d_inst[k], fs.d_hval[k], fs.d_udist[k], dist);
      launches++;
    };
    for (int k = 0; k < 2; k++) {
      const int n = k ? fs.n_tiles4 : fs.n_tiles8;
      if (n == 0) continue;
      st_k = (k == 1 && two) ? side : stream;
      const uint32_t* d_tl = k ? fs.d_tiles4 : fs.d_tiles8;
      const FracUTile* ut = nullptr; const uint32_t* cnt = nullptr; uint32_t* out = dist;
      if (unique) {
        if (!unique_begin(k)) return -1;
        ut = fs.d_utiles[k]; cnt = fs.d_count + 2 * stage + k; out = fs.d_udist[k];
      }
      if constexpr (PATCH) {
        const int nb = (n + tpc - 1) / tpc;
        if (k == 0) {
          if (!use_had)      k_frac_patch<8, false><<<nb, FRAC_TILE_THREADS, 0, st_k>>>(stage, d_tasks, d_results, d_tl, n, out, cur, ref, ut, cnt);
          else if (old_path) k_frac_patch<8, true><<<nb, FRAC_TILE_THREADS, 0, st_k>>>(stage, d_tasks, d_results, d_tl, n, out, cur, ref, ut, cnt);
          else               k_frac_hv<8, true><<<nb, FRAC_TILE_THREADS, 0, st_k>>>(stage, d_tasks, d_results, d_tl, n, out, cur, ref, ut, cnt);
        } else {
          if (!use_had)      k_frac_patch<4, false><<<nb, FRAC_TILE_THREADS, 0, st_k>>>(stage, d_tasks, d_results, d_tl, n, out, cur, ref, ut, cnt);
          else if (!hv4)     k_frac_patch<4, true><<<nb, FRAC_TILE_THREADS, 0, st_k>>>(stage, d_tasks, d_results, d_tl, n, out, cur, ref, ut, cnt);
          else               k_frac_hv<4, true><<<nb, FRAC_TILE_THREADS, 0, st_k>>>(stage, d_tasks, d_results, d_tl, n, out, cur, ref, ut, cnt);
        }
      } else {
        const int nb = blocks(n);
        if (k == 0) {
          if (use_had) k_frac_tiles<RefT, OrgT, 8, true><<<nb, FRAC_TILE_THREADS, 0, st_k>>>(stage, d_tasks, d_results, d_tl, n, out, cur, ref, ut, cnt);
          else         k_frac_tiles<RefT, OrgT, 8, false><<<nb, FRAC_TILE_THREADS, 0, st_k>>>(stage, d_tasks, d_results, d_tl, n, out, cur, ref, ut, cnt);
        } else {
          if (use_had) k_frac_tiles<RefT, OrgT, 4, true><<<nb, FRAC_TILE_THREADS, 0, st_k>>>(stage, d_tasks, d_results, d_tl, n, out, cur, ref, ut, cnt);
          else         k_frac_tiles<RefT, OrgT, 4, false><<<nb, FRAC_TILE_THREADS, 0, st_k>>>(stage, d_tasks, d_results, d_tl, n, out, cur, ref, ut, cnt);
        }
      }
      launches++;
      if (unique) unique_end(k);
    }
    if (two && (cudaEventRecord(ev_join, side) != cudaSuccess || cudaStreamWaitEvent(stream, ev_join, 0) != cudaSuccess)) return -1;
    const int nb = (fs.n_pu + 255) / 256;
    if (stage == 0) k_frac_argmin<0><<<nb, 256, 0, stream>>>(d_tasks, d_results, dist0, dist1, fs.n_pu, ref.bit_depth);
    else            k_frac_argmin<1><<<nb, 256, 0, stream>>>(d_tasks, d_results, dist1, nullptr, fs.n_pu, ref.bit_depth);
    launches++;
  }
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}

// Motion-compensated distortion of n PUs: tiles of one candidate each (k_frac_tiles stage 2), then k_mc_collect.
template <typename RefT, typename OrgT>
inline int mc_launch(const FracSchedule& fs, const SearchTask* d_tasks, const hmb200_pu_result* d_mv, uint32_t* d_out, const DevPlane& cur,
                     const DevPlane& ref, bool use_had, cudaStream_t stream) {
  if (fs.n_pu == 0) return 0;
  int launches = 0;
  if (cudaMemsetAsync(fs.d_dist, 0, (size_t)fs.n_pu * 9 * sizeof(uint32_t), stream) != cudaSuccess) return -1;
  auto blocks = [&](int n_tiles) { return (int)((((long long)(n_tiles + 31) / 32) * 32 + FRAC_TILE_THREADS - 1) / FRAC_TILE_THREADS); };
  if (fs.n_tiles8 > 0) {
    if (use_had) k_frac_tiles<RefT, OrgT, 8, true><<<blocks(fs.n_tiles8), FRAC_TILE_THREADS, 0, stream>>>(2, d_tasks, d_mv, fs.d_tiles8, fs.n_tiles8, fs.d_dist, cur, ref);
    else         k_frac_tiles<RefT, OrgT, 8, false><<<blocks(fs.n_tiles8), FRAC_TILE_THREADS, 0, stream>>>(2, d_tasks, d_mv, fs.d_tiles8, fs.n_tiles8, fs.d_dist, cur, ref);
    launches++;
  }
  if (fs.n_tiles4 > 0) {
    if (use_had) k_frac_tiles<RefT, OrgT, 4, true><<<blocks(fs.n_tiles4), FRAC_TILE_THREADS, 0, stream>>>(2, d_tasks, d_mv, fs.d_tiles4, fs.n_tiles4, fs.d_dist, cur, ref);
    else         k_frac_tiles<RefT, OrgT, 4, false><<<blocks(fs.n_tiles4), FRAC_TILE_THREADS, 0, stream>>>(2, d_tasks, d_mv, fs.d_tiles4, fs.n_tiles4, fs.d_dist, cur, ref);
    launches++;
  }
  k_mc_collect<<<(fs.n_pu + 255) / 256, 256, 0, stream>>>(fs.d_dist, d_out, fs.n_pu, ref.bit_depth);
  launches++;
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}

}  // namespace hmb200
