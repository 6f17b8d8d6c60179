// Reference-side binding of libhmb200 for HM-16.5 (xkfz007/video_codecs, hm-16.5rc1): the forwarders that
// integration/build_shim.py injects into TEncSearch::xPatternSearch, ::xPatternSearchFracDIF and
// ::xMotionEstimation call the three functions below.  OUR code; it includes the reference's headers from where they
// lie under /root/reference at build time and is compiled into integration/_build/ only (never copied into the repo).
//
// HMB200_SHIM=gpu (default)  every integer search and fractional refinement runs on the GPU through the C-ABI
// HMB200_SHIM=verify         the GPU result is checked call by call against the reference body (abort on mismatch)
// HMB200_SHIM=off            the reference bodies run (the stock encoder)
// HMB200_SHIM_LOG=<file>     appends one record per call (PU, window, predictor, MV, SAD / half, quarter, cost)
// HMB200_SHIM_TABLE=<n>|all  installs forwarders in TComRdCost::m_afpDistortFunc (the FpDistFunc table, TComRdCost.cpp:224-276):
//                            every n-th call of each family (SAD / SADS / SSE / HADS) is also evaluated by hmb200_dist and compared
//                            (abort on mismatch), the reference value is returned; `all` returns the GPU value of every call
//                            (per-call host copies: tiny clips only).  Unset: the table is left alone.
//
// The private cost state of TComRdCost (m_uiCost, m_mvPredictor; TLibCommon/TComRdCost.h:118-130) is read with the
// access-specifier trick below; a maintainer would add three public getters instead.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
// every standard header the reference pulls in, before the access-specifier trick (their include guards then keep
// the trick away from libstdc++)
#include <algorithm>
#include <cassert>
#include <cmath>
#include <deque>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <utility>
#include <vector>
#include <math.h>
#include <stdint.h>
#include <time.h>

#define private public
#define protected public
#include "TLibEncoder/TEncSearch.h"
#include "TLibCommon/TComRdCost.h"
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComPicYuv.h"
#include "TLibCommon/TComDataCU.h"
#undef private
#undef protected

#include "hmb200.h"

namespace {

enum Mode { OFF, GPU, VERIFY };
struct Shim {
  Mode mode = GPU;
  bool ready = false;
  bool in_reference = false;          // re-entrancy guard of verify mode
  FILE* log = nullptr;
  std::map<const Pel*, std::pair<int, int> > planes;   // buffer origin -> (plane id, POC)
  unsigned long long n_search = 0, n_frac = 0, n_upload = 0, n_frac_fused = 0, n_merge = 0, n_merge_cands = 0;
  double seconds = 0.0;                              // wall time spent inside the forwarders (host copies, launches, waits)
  // xMotionEstimation calls xPatternSearch and then xPatternSearchFracDIF on the MV it found (TEncSearch.cpp:3728, 3749): the
  // search forwarder runs both in one device round trip and keeps the refinement for the forwarder that follows
  struct Fused {
    bool valid = false;
    const Pel* roi = nullptr; const Pel* ref = nullptr;
    int w = 0, h = 0, flags_hadme = 0;
    hmb200_mv mv{0, 0}, half{0, 0}, qter{0, 0};
    hmb200_cost_state cs{0, {0, 0}};
    uint32_t cost = 0;
  } fused;
  // distortion-table hook
  FpDistFunc orig[DF_TOTAL_FUNCTIONS];
  unsigned long long table_period = 0;               // 0: not installed; 1: every call goes to the GPU
  unsigned long long table_calls[4] = {0, 0, 0, 0}, table_checked[4] = {0, 0, 0, 0};
};
Shim g_shim;

struct Stopwatch {                                   // adds a forwarder's wall time to g_shim.seconds
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  ~Stopwatch();
};

Stopwatch::~Stopwatch() { g_shim.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }

void die(const char* what) {
  fprintf(stderr, "hmb200 shim: %s: %s\n", what, hmb200_last_error());
  abort();                            // like the reference: assert/exit, no CPU fallback
}

void at_exit() {
  fprintf(stderr, "hmb200 shim: %llu integer searches, %llu fractional refinements (%llu of them served by the search's round trip), %llu plane uploads, %llu kernel launches, %.2f s inside the forwarders\n",
          g_shim.n_search, g_shim.n_frac, g_shim.n_frac_fused, g_shim.n_upload, (unsigned long long)hmb200_launch_count(), g_shim.seconds);
  {
    uint64_t calls = 0, cu_launches = 0, cu_served = 0;
    hmb200_one_call_stats(&calls, &cu_launches, &cu_served);
    fprintf(stderr, "hmb200 shim: %llu search+refinement calls: %llu launched a whole-CU search, %llu were answered from one\n",
            (unsigned long long)calls, (unsigned long long)cu_launches, (unsigned long long)cu_served);
  }
  if (g_shim.n_merge)
    fprintf(stderr, "hmb200 shim: %llu merge estimations (%llu candidates) through hmb200_merge_estimation_batch\n", g_shim.n_merge, g_shim.n_merge_cands);
  if (g_shim.table_period)
    fprintf(stderr, "hmb200 shim: distortion-table hook: SAD %llu/%llu SADS %llu/%llu SSE %llu/%llu HADS %llu/%llu calls evaluated on the GPU (all equal)\n",
            g_shim.table_checked[0], g_shim.table_calls[0], g_shim.table_checked[3], g_shim.table_calls[3], g_shim.table_checked[1],
            g_shim.table_calls[1], g_shim.table_checked[2], g_shim.table_calls[2]);
  if (g_shim.log) fclose(g_shim.log);
}

bool active() {
  if (!g_shim.ready) {
    const char* m = getenv("HMB200_SHIM");
    g_shim.mode = (m && !strcmp(m, "off")) ? OFF : (m && !strcmp(m, "verify")) ? VERIFY : GPU;
    if (const char* l = getenv("HMB200_SHIM_LOG")) g_shim.log = fopen(l, "w");
    if (g_shim.mode != OFF) {
      const char* d = getenv("HMB200_DEVICE");
      if (hmb200_init(d ? atoi(d) : 0) != HMB200_OK) die("hmb200_init");
      atexit(at_exit);
    }
    g_shim.ready = true;
  }
  return g_shim.mode != OFF && !g_shim.in_reference;
}

hmb200_cost_state cost_state(TComRdCost* rd) {
  hmb200_cost_state cs;
  cs.lambda_cost = rd->m_uiCost;
  cs.pred.x = rd->m_mvPredictor.getHor();
  cs.pred.y = rd->m_mvPredictor.getVer();
  return cs;
}

hmb200_pattern pattern_of(TComPattern* key) {
  hmb200_pattern p;
  p.roi = key->getROIY();
  p.width = key->getROIYWidth();
  p.height = key->getROIYHeight();
  p.stride = key->getPatternLStride();
  p.bit_depth = key->getBitDepthY();
  return p;
}

}  // namespace

// Injected at the top of xMotionEstimation: makes sure the reconstructed reference picture is resident on the GPU.
// Reconstructed planes are final and border-extended when a picture enters a reference list
// (TLibCommon/TComSlice.cpp:351-377); a picture buffer is re-used for later pictures, hence the POC check.
static int plane_of(TComPicYuv* yuv, int poc, int bit_depth, int kind) {
  const Pel* origin = yuv->getAddr(COMPONENT_Y);
  std::map<const Pel*, std::pair<int, int> >::iterator it = g_shim.planes.find(origin);
  if (it != g_shim.planes.end()) {
    if (it->second.second == poc) return it->second.first;
    hmb200_release_plane(it->second.first);
    g_shim.planes.erase(it);
  }
  const int mx = yuv->getMarginX(COMPONENT_Y), my = yuv->getMarginY(COMPONENT_Y);
  const int id = hmb200_register_plane(origin, yuv->getStride(COMPONENT_Y), yuv->getWidth(COMPONENT_Y), yuv->getHeight(COMPONENT_Y), mx, my,
                                       bit_depth, kind, poc);
  if (id < 0) die("hmb200_register_plane");
  g_shim.planes[origin] = std::make_pair(id, poc);
  g_shim.n_upload++;
  return id;
}

void hmb200_shim_ref_plane(TComPic* pic) {
  if (!active()) return;
  Stopwatch sw;
  plane_of(pic->getPicYuvRec(), pic->getPOC(), pic->getPicSym()->getSPS().getBitDepth(CHANNEL_TYPE_LUMA), HMB200_PLANE_REC);
}

// TEncSearch::xPatternSearch (TLibEncoder/TEncSearch.cpp:3786-3843).  Returns true when the call was served.
bool hmb200_shim_pattern_search(TEncSearch* self, TComPattern* key, Pel* piRefY, Int iRefStride, TComMv* lt, TComMv* rb,
                                TComMv& rcMv, Distortion& ruiSAD) {
  if (!active()) return false;
  Stopwatch sw;
  if (self->m_cDistParam.bApplyWeight) { fprintf(stderr, "hmb200 shim: weighted prediction is out of scope\n"); abort(); }
  const hmb200_pattern p = pattern_of(key);
  const hmb200_cost_state cs = cost_state(self->m_pcRdCost);
  hmb200_mv l = {lt->getHor(), lt->getVer()}, r = {rb->getHor(), rb->getVer()}, mv;
  uint32_t sad = 0;
  const int flags = self->m_pcEncCfg->getUseFastEnc() ? HMB200_FLAG_FEN : 0;
  static const bool fuse = getenv("HMB200_SHIM_NO_FUSE") == nullptr;
  g_shim.fused.valid = false;
  if (fuse) {
    const int hadme = self->m_pcEncCfg->getUseHADME() ? HMB200_FLAG_HADME : 0;
    Shim::Fused& f = g_shim.fused;
    if (hmb200_pattern_search_and_refine(&p, piRefY, iRefStride, l, r, &cs, flags | hadme, &mv, &sad, &f.half, &f.qter, &f.cost) != HMB200_OK)
      die("hmb200_pattern_search_and_refine");
    f.valid = true; f.roi = p.roi; f.ref = piRefY; f.w = p.width; f.h = p.height; f.flags_hadme = hadme; f.mv = mv; f.cs = cs;
  } else if (hmb200_pattern_search(&p, piRefY, iRefStride, l, r, &cs, flags, &mv, &sad) != HMB200_OK) die("hmb200_pattern_search");
  g_shim.n_search++;
  if (g_shim.mode == VERIFY) {
    TComMv ref_mv; Distortion ref_sad = 0;
    g_shim.in_reference = true;
    self->xPatternSearch(key, piRefY, iRefStride, lt, rb, ref_mv, ref_sad);
    g_shim.in_reference = false;
    if (ref_mv.getHor() != mv.x || ref_mv.getVer() != mv.y || ref_sad != sad) {
      fprintf(stderr, "hmb200 shim: xPatternSearch mismatch %dx%d: gpu (%d,%d) %u, reference (%d,%d) %u\n", p.width, p.height,
              mv.x, mv.y, sad, ref_mv.getHor(), ref_mv.getVer(), (unsigned)ref_sad);
      abort();
    }
  }
  if (g_shim.log)
    fprintf(g_shim.log, "I %dx%d lt %d %d rb %d %d pred %d %d lam %u -> %d %d %u\n", p.width, p.height, l.x, l.y, r.x, r.y, cs.pred.x,
            cs.pred.y, cs.lambda_cost, mv.x, mv.y, sad);
  rcMv.set((Short)mv.x, (Short)mv.y);
  ruiSAD = sad;
  return true;
}

// TEncSearch::xPatternSearchFast -> xTZSearch (TLibEncoder/TEncSearch.cpp:3847-3875, 3881-4083), FastSearch = 1 only;
// any other fast-search mode keeps the reference body.
bool hmb200_shim_pattern_search_fast(TEncSearch* self, TComDataCU* pcCU, TComPattern* key, Pel* piRefY, Int iRefStride, TComMv* lt,
                                     TComMv* rb, TComMv& rcMv, Distortion& ruiSAD, const TComMv* pIntegerMv2Nx2NPred) {
  if (!active()) return false;
  Stopwatch sw;
  g_shim.fused.valid = false;
  if (self->m_iFastSearch != 1) return false;
  if (self->m_cDistParam.bApplyWeight) { fprintf(stderr, "hmb200 shim: weighted prediction is out of scope\n"); abort(); }
  const hmb200_pattern p = pattern_of(key);
  const hmb200_cost_state cs = cost_state(self->m_pcRdCost);
  if (cs.pred.x != rcMv.getHor() || cs.pred.y != rcMv.getVer()) return false;      // start vector != cost predictor: not the xMotionEstimation call
  hmb200_tz_extra ex;
  memset(&ex, 0, sizeof(ex));
  ex.cu_x = (int)pcCU->getCUPelX(); ex.cu_y = (int)pcCU->getCUPelY();
  if (pIntegerMv2Nx2NPred) { ex.has_imv = 1; ex.imv_x = pIntegerMv2Nx2NPred->getHor(); ex.imv_y = pIntegerMv2Nx2NPred->getVer(); }
  const TComSPS& sps = *pcCU->getSlice()->getSPS();
  hmb200_mv l = {lt->getHor(), lt->getVer()}, r = {rb->getHor(), rb->getVer()}, mv;
  uint32_t sad = 0;
  const int flags = (self->m_pcEncCfg->getUseFastEnc() ? HMB200_FLAG_FEN : 0) |
                    (self->m_pcEncCfg->getFastMEAssumingSmootherMVEnabled() ? HMB200_FLAG_TZ_STOP : 0);
  static const bool fuse = getenv("HMB200_SHIM_NO_FUSE") == nullptr;
  if (fuse) {
    const int hadme = self->m_pcEncCfg->getUseHADME() ? HMB200_FLAG_HADME : 0;
    Shim::Fused& f = g_shim.fused;
    if (hmb200_pattern_search_tz_and_refine(&p, piRefY, iRefStride, l, r, &cs, flags | hadme, &ex, (int)sps.getPicWidthInLumaSamples(),
                                            (int)sps.getPicHeightInLumaSamples(), (int)sps.getMaxCUWidth(), self->m_iSearchRange, &mv, &sad,
                                            &f.half, &f.qter, &f.cost) != HMB200_OK)
      die("hmb200_pattern_search_tz_and_refine");
    f.valid = true; f.roi = p.roi; f.ref = piRefY; f.w = p.width; f.h = p.height; f.flags_hadme = hadme; f.mv = mv; f.cs = cs;
  } else if (hmb200_pattern_search_tz(&p, piRefY, iRefStride, l, r, &cs, flags, &ex, (int)sps.getPicWidthInLumaSamples(),
                               (int)sps.getPicHeightInLumaSamples(), (int)sps.getMaxCUWidth(), self->m_iSearchRange, &mv, &sad) != HMB200_OK)
    die("hmb200_pattern_search_tz");
  g_shim.n_search++;
  if (g_shim.mode == VERIFY) {
    TComMv ref_mv = rcMv; Distortion ref_sad = 0;
    g_shim.in_reference = true;
    self->xPatternSearchFast(pcCU, key, piRefY, iRefStride, lt, rb, ref_mv, ref_sad, pIntegerMv2Nx2NPred);
    g_shim.in_reference = false;
    if (ref_mv.getHor() != mv.x || ref_mv.getVer() != mv.y || ref_sad != sad) {
      fprintf(stderr, "hmb200 shim: xTZSearch mismatch %dx%d: gpu (%d,%d) %u, reference (%d,%d) %u\n", p.width, p.height, mv.x, mv.y, sad,
              ref_mv.getHor(), ref_mv.getVer(), (unsigned)ref_sad);
      abort();
    }
  }
  if (g_shim.log)
    fprintf(g_shim.log, "T %dx%d lt %d %d rb %d %d pred %d %d lam %u -> %d %d %u\n", p.width, p.height, l.x, l.y, r.x, r.y, cs.pred.x, cs.pred.y,
            cs.lambda_cost, mv.x, mv.y, sad);
  rcMv.set((Short)mv.x, (Short)mv.y);
  ruiSAD = sad;
  return true;
}

// TEncSearch::xPatternSearchFracDIF (TLibEncoder/TEncSearch.cpp:4240-4276).
bool hmb200_shim_pattern_search_frac(TEncSearch* self, Bool lossless, TComPattern* key, Pel* piRefY, Int iRefStride, TComMv* mvInt,
                                     TComMv& rcMvHalf, TComMv& rcMvQter, Distortion& ruiCost) {
  if (!active()) return false;
  Stopwatch sw;
  const hmb200_pattern p = pattern_of(key);
  const hmb200_cost_state cs = cost_state(self->m_pcRdCost);
  hmb200_mv mi = {mvInt->getHor(), mvInt->getVer()}, half, qter;
  uint32_t cost = 0;
  const int flags = self->m_pcEncCfg->getUseHADME() ? HMB200_FLAG_HADME : 0;
  {
    Shim::Fused& f = g_shim.fused;
    const bool hit = f.valid && !lossless && f.roi == p.roi && f.ref == piRefY && f.w == p.width && f.h == p.height && f.flags_hadme == flags &&
                     f.mv.x == mi.x && f.mv.y == mi.y && f.cs.lambda_cost == cs.lambda_cost && f.cs.pred.x == cs.pred.x && f.cs.pred.y == cs.pred.y;
    f.valid = false;                                   // one use: only the call that directly follows the search
    if (hit) { half = f.half; qter = f.qter; cost = f.cost; g_shim.n_frac_fused++; }
    else if (hmb200_pattern_search_frac(lossless ? 1 : 0, &p, piRefY, iRefStride, mi, &cs, flags, &half, &qter, &cost) != HMB200_OK)
      die("hmb200_pattern_search_frac");
  }
  g_shim.n_frac++;
  if (g_shim.mode == VERIFY) {
    TComMv rh, rq; Distortion rc = 0;
    const Int scale = self->m_pcRdCost->m_iCostScale;
    g_shim.in_reference = true;
    self->xPatternSearchFracDIF(lossless, key, piRefY, iRefStride, mvInt, rh, rq, rc);
    g_shim.in_reference = false;
    self->m_pcRdCost->setCostScale(scale);
    if (rh.getHor() != half.x || rh.getVer() != half.y || rq.getHor() != qter.x || rq.getVer() != qter.y || rc != cost) {
      fprintf(stderr, "hmb200 shim: xPatternSearchFracDIF mismatch %dx%d: gpu h(%d,%d) q(%d,%d) %u, reference h(%d,%d) q(%d,%d) %u\n",
              p.width, p.height, half.x, half.y, qter.x, qter.y, cost, rh.getHor(), rh.getVer(), rq.getHor(), rq.getVer(), (unsigned)rc);
      abort();
    }
  }
  if (g_shim.log)
    fprintf(g_shim.log, "F %dx%d int %d %d -> %d %d %d %d %u\n", p.width, p.height, mi.x, mi.y, half.x, half.y, qter.x, qter.y, cost);
  rcMvHalf.set((Short)half.x, (Short)half.y);
  rcMvQter.set((Short)qter.x, (Short)qter.y);
  ruiCost = cost;
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// The distortion function-pointer table (row a2 of the scope table): TComRdCost::init() fills m_afpDistortFunc with the
// xGetSAD* / xGetSSE* / xGetHADs members (TLibCommon/TComRdCost.cpp:224-276); build_shim.py appends one call to the function
// below at the end of init().  A forwarder has the FpDistFunc signature (TComRdCost.h:60) and knows its table index.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

int family_of(int idx) {          // HMB200_DF_* of a DFunc index (TLibCommon/TypeDef.h:334-378)
  if (idx >= DF_SSE && idx <= DF_SSE16N) return HMB200_DF_SSE;
  if ((idx >= DF_SAD && idx <= DF_SAD16N) || (idx >= DF_SAD12 && idx <= DF_SAD48)) return HMB200_DF_SAD;
  if ((idx >= DF_SADS && idx <= DF_SADS16N) || (idx >= DF_SADS12 && idx <= DF_SADS48)) return HMB200_DF_SADS;
  if (idx >= DF_HADS && idx <= DF_HADS16N) return HMB200_DF_HADS;
  return -1;
}

Distortion table_forward(int idx, DistParam* p) {
  const int fam = family_of(idx);
  const Distortion ref = g_shim.orig[idx](p);
  if (fam < 0 || p->bApplyWeight || p->iStep != 1 || p->iCols > 64 || p->iRows > 64) return ref;    // outside the path: reference body only
  const unsigned long long n = ++g_shim.table_calls[fam];
  if (n % g_shim.table_period != 0) return ref;
  if (!active()) return ref;
  hmb200_dist_param dp;
  dp.pOrg = p->pOrg; dp.pCur = p->pCur; dp.iStrideOrg = p->iStrideOrg; dp.iStrideCur = p->iStrideCur;
  dp.iRows = p->iRows; dp.iCols = p->iCols; dp.iStep = p->iStep; dp.func = fam; dp.bitDepth = p->bitDepth;
  dp.bApplyWeight = 0; dp.iSubShift = p->iSubShift;
  const uint32_t gpu = hmb200_dist(&dp);
  g_shim.table_checked[fam]++;
  if ((Distortion)gpu != ref) {
    fprintf(stderr, "hmb200 shim: m_afpDistortFunc[%d] mismatch %dx%d subShift %d bitDepth %d: gpu %u, reference %u\n", idx, p->iCols, p->iRows,
            p->iSubShift, p->bitDepth, gpu, (unsigned)ref);
    abort();
  }
  return (Distortion)gpu;
}

template <int I> Distortion table_entry(DistParam* p) { return table_forward(I, p); }
template <int I> struct TableFill {
  static void run(FpDistFunc* t) { if (t[I]) t[I] = table_entry<I>; TableFill<I + 1>::run(t); }
};
template <> struct TableFill<DF_TOTAL_FUNCTIONS> { static void run(FpDistFunc*) {} };

}  // namespace

void hmb200_shim_dist_table(FpDistFunc* table, int n) {
  const char* m = getenv("HMB200_SHIM_TABLE");
  if (!m || !*m || n != DF_TOTAL_FUNCTIONS) return;
  g_shim.table_period = !strcmp(m, "all") ? 1ull : (unsigned long long)atoll(m);
  if (g_shim.table_period == 0) return;
  for (int i = 0; i < DF_TOTAL_FUNCTIONS; i++) g_shim.orig[i] = table[i];
  TableFill<0>::run(table);
}

// ---------------------------------------------------------------------------------------------------------------------
// TEncSearch::xMergeEstimation's candidate loop (TLibEncoder/TEncSearch.cpp:2868-2892): build_shim.py injects a call to this
// function right after xRestrictBipredMergeCand.  All candidates of the PU go to the device in one hmb200_merge_estimation_batch
// call: motion compensation (uni- or bi-directional) + HADs / SAD of every candidate, the argmin with getCost(bits) on the way back.
// Opt-in: HMB200_SHIM_MERGE=1 (default: the reference loop).  Returns true when the call was served.
// ---------------------------------------------------------------------------------------------------------------------
bool hmb200_shim_merge_estimation(TEncSearch* self, TComDataCU* pcCU, TComYuv* pcYuvOrg, Int iPUIdx, UInt uiAbsPartIdx, Int iWidth, Int iHeight,
                                  TComMvField* nb, UChar* dirs, Int numValid, UInt& uiInterDir, TComMvField* pacMvField, UInt& uiMergeIndex,
                                  Distortion& ruiCost) {
  if (!active() || numValid <= 0) return false;
  // opt-in (HMB200_SHIM_MERGE=1): exact, but one more device round trip per PU where the host needs ~5 small motion compensations -
  // measured on 1080p I+P it adds 5.7 s (191 164 calls x 30 us) to a 28.9 s encode; the batched entry pays off from a frontend that
  // hands over the candidates of many PUs at once
  static const bool enabled = getenv("HMB200_SHIM_MERGE") && strcmp(getenv("HMB200_SHIM_MERGE"), "0") != 0;
  if (!enabled) return false;
  TComSlice* slice = pcCU->getSlice();
  if (slice->getPPS()->getUseWP() || slice->getPPS()->getWPBiPred()) return false;       // weighted prediction is out of scope
  Stopwatch sw;
  g_shim.fused.valid = false;
  const int bd = slice->getSPS()->getBitDepth(CHANNEL_TYPE_LUMA);
  TComPic* pic = pcCU->getPic();
  TComPicYuv* org = pic->getPicYuvOrg();
  const int org_plane = plane_of(org, pic->getPOC(), bd, HMB200_PLANE_ORG);
  const ptrdiff_t off = org->getAddr(COMPONENT_Y, pcCU->getCtuRsAddr(), pcCU->getZorderIdxInCtu() + uiAbsPartIdx) - org->getAddr(COMPONENT_Y);
  const int stride = org->getStride(COMPONENT_Y);
  const int pu_y = (int)(off / stride), pu_x = (int)(off - (ptrdiff_t)pu_y * stride);     // positions inside the picture: 0 <= x < width <= stride
  hmb200_mc_cand c[MRG_MAX_NUM_CANDS];
  const int maxc = (int)self->m_pcEncCfg->getMaxNumMergeCand();
  for (int i = 0; i < numValid; i++) {
    const TComMvField &f0 = nb[2 * i], &f1 = nb[2 * i + 1];
    hmb200_mc_cand& k = c[i];
    memset(&k, 0, sizeof(k));
    k.pu_x = pu_x; k.pu_y = pu_y; k.w = iWidth; k.h = iHeight;
    int dir = 0;
    if (f0.getRefIdx() >= 0) {
      TComMv m = f0.getMv(); pcCU->clipMv(m);
      TComPic* rp = slice->getRefPic(REF_PIC_LIST_0, f0.getRefIdx());
      k.mv0_x = m.getHor(); k.mv0_y = m.getVer(); k.ref0_plane = plane_of(rp->getPicYuvRec(), rp->getPOC(), bd, HMB200_PLANE_REC);
      dir |= 1;
    }
    if (f1.getRefIdx() >= 0) {
      TComMv m = f1.getMv(); pcCU->clipMv(m);
      TComPic* rp = slice->getRefPic(REF_PIC_LIST_1, f1.getRefIdx());
      k.mv1_x = m.getHor(); k.mv1_y = m.getVer(); k.ref1_plane = plane_of(rp->getPicYuvRec(), rp->getPOC(), bd, HMB200_PLANE_REC);
      dir |= 2;
    }
    if (dir == 0) return false;                                                            // not a usable candidate: leave it to the reference
    if (dir == 3) {
      // xCheckIdenticalMotion (TLibCommon/TComPrediction.cpp:496-517) looks at the UNCLIPPED vectors and the POCs
      const bool same = slice->isInterB() && slice->getRefPic(REF_PIC_LIST_0, f0.getRefIdx())->getPOC() == slice->getRefPic(REF_PIC_LIST_1, f1.getRefIdx())->getPOC() &&
                        f0.getMv() == f1.getMv();
      dir = same ? 1 : (3 | HMB200_INTER_DIR_NO_IDENTICAL_CHECK);
    }
    k.inter_dir = dir;
    k.bits = i + 1 - (i == maxc - 1 ? 1 : 0);
  }
  const int32_t first[2] = {0, (int32_t)numValid};
  uint32_t best = 0, cost = 0, dist[MRG_MAX_NUM_CANDS];
  const int use_had = self->m_pcEncCfg->getUseHADME() && (pcCU->getCUTransquantBypass(iPUIdx) == 0);
  if (hmb200_merge_estimation_batch(org_plane, 1, first, c, use_had, self->m_pcRdCost->m_uiCost, &best, &cost, dist) != HMB200_OK)
    die("hmb200_merge_estimation_batch");
  g_shim.n_merge++; g_shim.n_merge_cands += (unsigned long long)numValid;
  const PartSize ePartSize = pcCU->getPartitionSize(0);
  if (g_shim.mode == VERIFY) {
    // the reference loop, candidate by candidate
    for (int i = 0; i < numValid; i++) {
      Distortion e = 0;
      pcCU->getCUMvField(REF_PIC_LIST_0)->setAllMvField(nb[2 * i], ePartSize, uiAbsPartIdx, 0, iPUIdx);
      pcCU->getCUMvField(REF_PIC_LIST_1)->setAllMvField(nb[2 * i + 1], ePartSize, uiAbsPartIdx, 0, iPUIdx);
      g_shim.in_reference = true;
      self->xGetInterPredictionError(pcCU, pcYuvOrg, iPUIdx, e, self->m_pcEncCfg->getUseHADME());
      g_shim.in_reference = false;
      if ((uint32_t)e != dist[i]) {
        fprintf(stderr, "hmb200 shim: merge candidate %d of a %dx%d PU at (%d,%d): gpu %u, reference %u (dir %d)\n", i, iWidth, iHeight, pu_x, pu_y,
                dist[i], (unsigned)e, c[i].inter_dir);
        abort();
      }
    }
  }
  // what the reference loop leaves behind: the CU's MV fields hold the last candidate
  pcCU->getCUMvField(REF_PIC_LIST_0)->setAllMvField(nb[2 * (numValid - 1)], ePartSize, uiAbsPartIdx, 0, iPUIdx);
  pcCU->getCUMvField(REF_PIC_LIST_1)->setAllMvField(nb[2 * (numValid - 1) + 1], ePartSize, uiAbsPartIdx, 0, iPUIdx);
  ruiCost = cost;
  uiMergeIndex = best;
  uiInterDir = dirs[best];
  pacMvField[0] = nb[2 * best];
  pacMvField[1] = nb[2 * best + 1];
  return true;
}
