// Batched TZ fast search: TEncSearch::xTZSearch (TLibEncoder/TEncSearch.cpp:3881-4083), the encoder's default integer
// search (FastSearch = 1), reached through xPatternSearchFast (:3847) — with xTZSearchHelp (:332-436, non-selective
// branch), xTZ8PointDiamondSearch (:626-800), xTZ2PointSearch (:438-567) under TZ_SEARCH_CONFIGURATION (:297-313).
//
// Every stage of the reference issues xTZSearchHelp calls whose positions depend only on the stage's start point, so the
// sequential strict-'<' updates equal "first minimum of an ordered point list".  One warp owns one PU: lanes evaluate
// the points of a stage in parallel (a full SAD + MV cost per lane: VABSDIFF4 on byte-realigned words read straight
// from the padded plane through L1/L2 — a PU touches a few hundred scattered candidates, not a window worth staging),
// then an ordered warp minimum (cost << 32 | position in call order) reproduces the sequential result.
// A point is tested only against the window edges it moved towards — the rule behind every range check of the
// reference's diamond and two-point code.  uiBestRound (diamond calls since the last improvement) only feeds the first
// search's stop criterion (bFirstSearchStop = FastMEAssumingSmootherMVEnabled, on by default): with it, the first three
// diamonds (the earliest possible stop) are one batch with per-diamond minima, later diamonds go one at a time.
#pragma once
#include "hmb200_device.cuh"
#include "hmb200_generic.cuh"

namespace hmb200 {

constexpr int TZ_WARPS = 4;

struct TzParams { int32_t pic_w, pic_h, max_cu, search_range, first_stop; };

// TComDataCU::clipMv (TLibCommon/TComDataCU.cpp:2788-2801), quarter-pel
__device__ __forceinline__ void tz_clip_mv(int& x, int& y, int cu_x, int cu_y, const TzParams& P) {
  const int hmax = (P.pic_w + 8 - cu_x - 1) * 4, hmin = (-P.max_cu - 8 - cu_x + 1) * 4;
  const int vmax = (P.pic_h + 8 - cu_y - 1) * 4, vmin = (-P.max_cu - 8 - cu_y + 1) * 4;
  x = min(hmax, max(hmin, x));
  y = min(vmax, max(vmin, y));
}

struct TzPoint { int x, y, nr, dist; bool valid; };

// moved-towards rule: (x, y) relative to the stage's start (sx, sy)
__device__ __forceinline__ bool tz_in_range(int x, int y, int sx, int sy, int L, int T, int R, int B) {
  return (x >= sx || x >= L) && (x <= sx || x <= R) && (y >= sy || y >= T) && (y <= sy || y <= B);
}

// idx-th point of the diamond sequence d = 1, 2, 4, ... <= search_range around (sx, sy), in the reference's call order
__device__ __forceinline__ TzPoint tz_diamond_point(int idx, int sx, int sy, int search_range, int L, int T, int R, int B) {
  TzPoint p{0, 0, 0, 0, false};
  int d = 1;
  for (;; d <<= 1) {
    if (d > search_range) return p;
    const int n = d == 1 ? 4 : (d <= 8 ? 8 : 16);
    if (idx < n) break;
    idx -= n;
  }
  int dx, dy;
  if (d == 1) {
    dx = idx == 1 ? -1 : (idx == 2 ? 1 : 0);  dy = idx == 0 ? -1 : (idx == 3 ? 1 : 0);
    p.nr = idx == 0 ? 2 : (idx == 1 ? 4 : (idx == 2 ? 5 : 7));  p.dist = 1;
  } else if (d <= 8) {
    const int h = d >> 1;
    //            top    (l2,t2) (r2,t2) left   right  (l2,b2) (r2,b2) bottom
    const int ax[8] = {0, -h, h, -d, d, -h, h, 0}, ay[8] = {-d, -h, -h, 0, 0, h, h, d};
    const int an[8] = {2, 1, 3, 4, 5, 6, 8, 7};
    dx = ax[idx]; dy = ay[idx]; p.nr = an[idx];
    p.dist = (idx == 0 || idx == 3 || idx == 4 || idx == 7) ? d : h;
  } else {
    if (idx < 4) { dx = idx == 1 ? -d : (idx == 2 ? d : 0); dy = idx == 0 ? -d : (idx == 3 ? d : 0); }
    else {
      const int i = ((idx - 4) >> 2) + 1, j = (idx - 4) & 3, q = (d >> 2) * i;
      dx = (j & 1) ? q : -q;                       // XL, XR, XL, XR
      dy = (j & 2) ? (d - q) : -(d - q);           // YT = top + q, YB = bottom - q
    }
    p.nr = 0; p.dist = d;
  }
  p.x = sx + dx; p.y = sy + dy;
  p.valid = tz_in_range(p.x, p.y, sx, sy, L, T, R, B);
  return p;
}

__device__ __forceinline__ int tz_diamond_count(int search_range) {
  int n = 0;
  for (int d = 1; d <= search_range; d <<= 1) n += d == 1 ? 4 : (d <= 8 ? 8 : 16);
  return n;
}

// xGetSADnn with iSubShift (TComRdCost.cpp:489-953) at integer displacement (x, y) + getCost (TComRdCost.h:172-189)
template <typename T, typename OrgT>
__device__ __forceinline__ uint32_t tz_cost(const uint8_t* s_org, const T* ref0, int ref_pitch, const SearchTask& t, int bit_depth, int x, int y) {
  const int step = 1 << t.sub_shift, rows = t.h >> t.sub_shift;
  uint32_t sum = 0;
  if constexpr (sizeof(T) == 1 && sizeof(OrgT) == 1) {
    const int ww = t.w >> 2;
    for (int r = 0; r < rows; r++) {
      const uint32_t* ow = reinterpret_cast<const uint32_t*>(s_org + (r * step) * t.w);
      const uintptr_t a = reinterpret_cast<uintptr_t>(ref0 + (ptrdiff_t)(y + r * step) * ref_pitch + x);
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
      const uint32_t sh = (uint32_t)(a & 3) * 8u;
      uint32_t lo = __ldg(rw);
      for (int i = 0; i < ww; i++) {
        const uint32_t hi = __ldg(rw + i + 1);
        sum = sad4_acc(__funnelshift_r(lo, hi, sh), ow[i], sum);
        lo = hi;
      }
    }
  } else {
    const OrgT* so = reinterpret_cast<const OrgT*>(s_org);      // 16-bit planes, or a signed Pel pattern (1:1 entry)
    for (int r = 0; r < rows; r++) {
      const T* rr = ref0 + (ptrdiff_t)(y + r * step) * ref_pitch + x;
      for (int c = 0; c < t.w; c++) sum += (uint32_t)abs((int)so[(r * step) * t.w + c] - (int)rr[c]);
    }
  }
  sum = (sum << t.sub_shift) >> (bit_depth - 8);
  return sum + mv_cost(t.lambda_cost, mv_bits(x, y, t.pred_x, t.pred_y, 2));
}

struct TzState { uint32_t cost; int x, y, dist, nr; };

// lanes hold one candidate each (order = lane order); folds the warp's first minimum into the state (strict '<');
// returns whether the state improved (warp-uniform)
__device__ __forceinline__ bool tz_fold(TzState& st, bool valid, uint32_t cost, const TzPoint& p, int lane) {
  unsigned long long key = valid ? make_key(cost, (uint32_t)lane) : ~0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other < key ? other : key;
  }
  if (key == ~0ull) return false;
  const uint32_t c = (uint32_t)(key >> 32);
  const int src = (int)(key & 31u);
  const int bx = __shfl_sync(0xffffffffu, p.x, src), by = __shfl_sync(0xffffffffu, p.y, src);
  const int bn = __shfl_sync(0xffffffffu, p.nr, src), bd = __shfl_sync(0xffffffffu, p.dist, src);
  if (c < st.cost) { st.cost = c; st.x = bx; st.y = by; st.nr = bn; st.dist = bd; return true; }
  return false;
}

template <typename T, typename OrgT>
__global__ void __launch_bounds__(TZ_WARPS * 32)
k_tz_search(const SearchTask* __restrict__ tasks, const hmb200_tz_extra* __restrict__ extra, hmb200_pu_result* __restrict__ out, int n,
            DevPlane cur_plane, DevPlane ref_plane, TzParams P) {
  __shared__ __align__(16) uint8_t s_org_all[TZ_WARPS][64 * 64 * sizeof(OrgT)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pu = blockIdx.x * TZ_WARPS + warp;
  if (pu >= n) return;
  const SearchTask t = tasks[pu];
  const hmb200_tz_extra ex = extra[pu];
  uint8_t* s_org = s_org_all[warp];
  {   // the PU's original block, row-major with stride w
    const OrgT* org = plane_at<OrgT>(cur_plane, t.org_x, t.org_y);
    OrgT* so = reinterpret_cast<OrgT*>(s_org);
    for (int i = lane; i < t.w * t.h; i += 32) { const int r = i / t.w, c = i - r * t.w; so[i] = org[(size_t)r * cur_plane.pitch + c]; }
  }
  __syncwarp();
  const T* ref0 = plane_at<T>(ref_plane, t.ref_x, t.ref_y);
  const int pitch = ref_plane.pitch, bd = ref_plane.bit_depth;
  const int L = t.lt_x, Tp = t.lt_y, R = t.rb_x, B = t.rb_y;
  int rl = L, rt = Tp, rr = R, rb = B;                       // range of the raster stage
  TzState st{0xffffffffu, 0, 0, 0, 0};

  // ---- start candidates: clipped predictor, zero vector, integer MV of the 2Nx2N PU (:3905-3946) -------------------
  {
    int sx = t.pred_x, sy = t.pred_y;
    tz_clip_mv(sx, sy, ex.cu_x, ex.cu_y, P);
    int ix = (int)(int16_t)(ex.imv_x << 2), iy = (int)(int16_t)(ex.imv_y << 2);
    tz_clip_mv(ix, iy, ex.cu_x, ex.cu_y, P);
    TzPoint p{0, 0, 0, 0, false};
    if (lane == 0) { p.x = sx >> 2; p.y = sy >> 2; p.valid = true; }
    if (lane == 1) { p.valid = true; }
    if (lane == 2 && ex.has_imv) { p.x = ix >> 2; p.y = iy >> 2; p.valid = true; }
    const uint32_t c = p.valid ? tz_cost<T, OrgT>(s_org, ref0, pitch, t, bd, p.x, p.y) : 0u;
    tz_fold(st, p.valid, c, p, lane);
    if (ex.has_imv) {    // only the raster stage sees the range re-centred on the best start (:3935-3946)
      int cx = (int)(int16_t)(st.x << 2), cy = (int)(int16_t)(st.y << 2);
      tz_clip_mv(cx, cy, ex.cu_x, ex.cu_y, P);
      int lx = (int)(int16_t)(cx - (P.search_range << 2)), ly = (int)(int16_t)(cy - (P.search_range << 2));
      int rx = (int)(int16_t)(cx + (P.search_range << 2)), ry = (int)(int16_t)(cy + (P.search_range << 2));
      tz_clip_mv(lx, ly, ex.cu_x, ex.cu_y, P);
      tz_clip_mv(rx, ry, ex.cu_x, ex.cu_y, P);
      rl = lx >> 2; rt = ly >> 2; rr = rx >> 2; rb = ry >> 2;
    }
  }
  const int n_diamond = tz_diamond_count(P.search_range);
  auto diamonds = [&](int sx, int sy) {
    for (int base = 0; base < n_diamond; base += 32) {
      const TzPoint p = tz_diamond_point(base + lane, sx, sy, P.search_range, L, Tp, R, B);
      const uint32_t c = p.valid ? tz_cost<T, OrgT>(s_org, ref0, pitch, t, bd, p.x, p.y) : 0u;
      tz_fold(st, p.valid, c, p, lane);
    }
  };
  auto two_point = [&]() {
    // the two untested neighbours of the best point, by the diamond position it came from (xTZ2PointSearch)
    const int n8 = st.nr;
    const int ox0[9] = {0, -1, -1, 0, -1, 1, -1, -1, 1}, oy0[9] = {0, 0, -1, -1, 1, -1, 0, 1, 0};
    const int ox1[9] = {0, 0, 1, 1, -1, 1, 0, 1, 0},     oy1[9] = {0, -1, -1, 0, -1, 1, 1, 1, 1};
    TzPoint p{0, 0, 0, 2, false};
    if (lane < 2 && n8 >= 1 && n8 <= 8) {
      p.x = st.x + (lane == 0 ? ox0[n8] : ox1[n8]);
      p.y = st.y + (lane == 0 ? oy0[n8] : oy1[n8]);
      p.valid = tz_in_range(p.x, p.y, st.x, st.y, L, Tp, R, B);
    }
    const uint32_t c = p.valid ? tz_cost<T, OrgT>(s_org, ref0, pitch, t, bd, p.x, p.y) : 0u;
    tz_fold(st, p.valid, c, p, lane);
  };

  // ---- first search, two-point completion, raster, star refinement (:3948-4078) ------------------------------------
  if (!P.first_stop) diamonds(st.x, st.y);
  else {
    // :3949-3966 with bFirstSearchStop: stop after uiFirstSearchRounds = 3 diamonds in a row without improvement
    const int sx = st.x, sy = st.y;
    int rounds = 0;
    {
      const TzPoint p = tz_diamond_point(lane, sx, sy, P.search_range, L, Tp, R, B);      // d = 1, 2, 4: 4 + 8 + 8 points
      const bool in_a = lane < 20 && p.valid;
      const uint32_t c = in_a ? tz_cost<T, OrgT>(s_org, ref0, pitch, t, bd, p.x, p.y) : 0u;
      const int k = lane < 4 ? 0 : (lane < 12 ? 1 : 2);
#pragma unroll
      for (int kk = 0; kk < 3; kk++) {
        if ((1 << kk) <= P.search_range) {
          const bool improved = tz_fold(st, in_a && k == kk, c, p, lane);
          rounds = improved ? 0 : rounds + 1;
        }
      }
    }
    int base = 20;
    for (int d = 8; d <= P.search_range && rounds < 3; d <<= 1) {
      const int nd = d <= 8 ? 8 : 16;
      TzPoint p{0, 0, 0, 0, false};
      if (lane < nd) p = tz_diamond_point(base + lane, sx, sy, P.search_range, L, Tp, R, B);
      const uint32_t c = p.valid ? tz_cost<T, OrgT>(s_org, ref0, pitch, t, bd, p.x, p.y) : 0u;
      const bool improved = tz_fold(st, p.valid, c, p, lane);
      rounds = improved ? 0 : rounds + 1;
      base += nd;
    }
  }
  if (st.dist == 1) { st.dist = 0; two_point(); }
  if (st.dist > 5) {
    st.dist = 5;
    const int nxr = rr >= rl ? (rr - rl) / 5 + 1 : 0, nyr = rb >= rt ? (rb - rt) / 5 + 1 : 0;
    for (int base = 0; base < nxr * nyr; base += 32) {
      const int i = base + lane;
      TzPoint p{0, 0, 0, 5, i < nxr * nyr};
      if (p.valid) { const int iy = i / nxr; p.x = rl + 5 * (i - iy * nxr); p.y = rt + 5 * iy; }
      const uint32_t c = p.valid ? tz_cost<T, OrgT>(s_org, ref0, pitch, t, bd, p.x, p.y) : 0u;
      tz_fold(st, p.valid, c, p, lane);
    }
  }
  while (st.dist > 0) {
    const int bx = st.x, by = st.y;
    st.dist = 0; st.nr = 0;
    diamonds(bx, by);
    if (st.dist == 1) { st.dist = 0; if (st.nr != 0) two_point(); }
  }
  if (lane == 0) {
    hmb200_pu_result r = out[pu];
    r.mv_x = st.x; r.mv_y = st.y;
    r.sad = st.cost - mv_cost(t.lambda_cost, mv_bits(st.x, st.y, t.pred_x, t.pred_y, 2));
    out[pu] = r;
  }
}

}  // namespace hmb200
