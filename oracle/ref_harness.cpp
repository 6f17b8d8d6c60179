// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C-ABI harness around the UNMODIFIED reference (xkfz007/video_codecs, hm-16.5rc1).  It is compiled by
// oracle/Makefile.ref against objects built from the sources where they lie under /root/reference and
// lands in oracle/_ref/libhmref.so.  Nothing of the reference is copied here: this file only *calls*
//   TComRdCost::setDistParam / DistParam::DistFunc / getDistPart     (TLibCommon/TComRdCost.h:153-158,221)
//   TEncSearch::xPatternSearch                                       (TLibEncoder/TEncSearch.cpp:3786)
//   TEncSearch::xPatternSearchFracDIF                                (TLibEncoder/TEncSearch.cpp:4240)
// through a subclass (both are `protected`, TLibEncoder/TEncSearch.h:413-430).
//
// Users: tests/ (pins oracle/hm_oracle.c and generates tests/golden/*), bench.py `--impl reference` and the
// `cpu_baseline` leg.  The product library (libhmb200.so) never links or loads this.
#include <cstdint>
#include <cstring>
#include <ctime>
// Standard headers first, then the access-specifier trick: the harness has to place a bare TComDataCU at a CU position
// (m_uiCUPelX/Y, m_pcSlice are private, TLibCommon/TComDataCU.h:73-82) to drive xTZSearch, which calls pcCU->clipMv.
// Layout is unaffected (GCC lays members out in declaration order regardless of access).
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
#include <map>
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
#include "TLibCommon/TComDataCU.h"
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComRdCost.h"
#undef private
#undef protected
#include "TLibCommon/TComRom.h"
#include "TLibCommon/TComRdCost.h"
#include "TLibCommon/TComPattern.h"
#include "TLibEncoder/TEncCfg.h"
#include "TLibEncoder/TEncSearch.h"
#include "TLibVideoIO/TVideoIOYuv.h"
#include "TLibCommon/TComPicYuv.h"

namespace {

struct Harness : public TEncSearch
{
  TEncCfg    cfg;
  TComRdCost rd;

  Harness(int fen, int hadme)
  {
    cfg.setUseFastEnc(fen != 0);
    cfg.setUseHADME(hadme != 0);
    m_pcEncCfg = &cfg;
    m_pcRdCost = &rd;
    initTempBuff(CHROMA_420);              // allocates m_filteredBlock / m_filteredBlockTmp (TComPrediction.cpp:126)
    m_cDistParam.bApplyWeight = false;     // what setWpScalingDistParam does when weighted prediction is off
    BitDepths bd; bd.recon[0] = bd.recon[1] = 8;
#if O0043_BEST_EFFORT_DECODING
    bd.stream[0] = bd.stream[1] = 8;
#endif
    rd.setLambda(0.0, bd);                 // m_uiLambdaMotionSAD[0] = 0, so getMotionCost(true, add, false) sets m_uiCost = add
  }
  ~Harness() { m_pcEncCfg = NULL; }        // the base destructor dereferences m_pcEncCfg (TEncSearch.cpp:152-178)

  void setCost(uint32_t uiCost, int predx, int predy, int scale)
  {
    rd.getMotionCost(true, (Int)uiCost, false);
    TComMv p((Short)predx, (Short)predy);
    rd.setPredictor(p);
    rd.setCostScale(scale);
  }

  void search(Pel* org, int orgStride, int w, int h, int bitDepth, Pel* refAtPu, int refStride,
              int ltx, int lty, int rbx, int rby, uint32_t uiCost, int predx, int predy,
              int* mvx, int* mvy, uint32_t* sad)
  {
    TComPattern pat;
    pat.initPattern(org, w, h, orgStride, bitDepth);
    setCost(uiCost, predx, predy, 2);      // TEncSearch.cpp:3719-3722
    TComMv lt((Short)ltx, (Short)lty), rb((Short)rbx, (Short)rby), mv;
    Distortion d = 0;
    xPatternSearch(&pat, refAtPu, refStride, &lt, &rb, mv, d);
    *mvx = mv.getHor(); *mvy = mv.getVer(); *sad = d;
  }

  void frac(int lossless, Pel* org, int orgStride, int w, int h, int bitDepth, Pel* refAtPu, int refStride,
            int mvx, int mvy, uint32_t uiCost, int predx, int predy,
            int* hx, int* hy, int* qx, int* qy, uint32_t* cost)
  {
    TComPattern pat;
    pat.initPattern(org, w, h, orgStride, bitDepth);
    setCost(uiCost, predx, predy, 1);      // TEncSearch.cpp:3745-3746
    TComMv mvInt((Short)mvx, (Short)mvy), half, qter;
    Distortion d = 0;
    xPatternSearchFracDIF(lossless != 0, &pat, refAtPu, refStride, &mvInt, half, qter, d);
    *hx = half.getHor(); *hy = half.getVer(); *qx = qter.getHor(); *qy = qter.getVer(); *cost = d;
  }

  // TComPrediction::xPredInterBlk (TComPrediction.cpp:668) for COMPONENT_Y, bi = false, on a reference picture we fill
  // from the caller's padded plane; then the distortion the way xGetTemplateCost (getDistPart DF_SAD) or
  // xGetInterPredictionError (setDistParam + bHadamard) take it.  The CU sits at the picture origin; the PU position is
  // folded into the motion vector (xPredInterBlk only adds (mv >> 2) to the block address).
  uint32_t mcDist(int kind, Pel* org, int orgStride, int w, int h, int bitDepth, const Pel* ref0, int refStride, int picW, int picH,
                  int margin, int pux, int puy, int mvx, int mvy)
  {
    TComPicYuv refPic;
    refPic.create(picW, picH, CHROMA_400, 64, 64, 4, true);               // margins 64 + 16 = 80 (TComPicYuv.cpp:93-94)
    const int m = refPic.getMarginX(COMPONENT_Y);
    for (int y = -m; y < picH + m; y++)
      for (int x = -m; x < picW + m; x++)
      {
        const int sy = y < -margin ? -margin : (y >= picH + margin ? picH + margin - 1 : y);
        const int sx = x < -margin ? -margin : (x >= picW + margin ? picW + margin - 1 : x);
        refPic.getAddr(COMPONENT_Y)[y * refPic.getStride(COMPONENT_Y) + x] = ref0[sy * refStride + sx];
      }
    TComPic pic;
    pic.m_apcPicYuv[TComPic::PIC_YUV_REC] = &refPic;
    cu.m_pcPic = &pic; cu.m_ctuRsAddr = 0; cu.m_absZIdxInCtu = 0;
    TComYuv dst;
    dst.create(64, 64, CHROMA_400);
    TComMv mv((Short)(mvx + 4 * pux), (Short)(mvy + 4 * puy));
    xPredInterBlk(COMPONENT_Y, &cu, &refPic, 0, &mv, w, h, &dst, false, bitDepth);
    uint32_t d;
    if (kind == 0)
      d = rd.getDistPart(bitDepth, dst.getAddr(COMPONENT_Y, 0), dst.getStride(COMPONENT_Y), org, orgStride, w, h, COMPONENT_Y, DF_SAD);
    else
    {
      DistParam dp; dp.bApplyWeight = false;
      rd.setDistParam(dp, bitDepth, org, orgStride, dst.getAddr(COMPONENT_Y, 0), dst.getStride(COMPONENT_Y), w, h, true);
      d = dp.DistFunc(&dp);
    }
    dst.destroy();
    pic.m_apcPicYuv[TComPic::PIC_YUV_REC] = NULL;
    cu.m_pcPic = NULL;
    refPic.destroy();
    return d;
  }

  // Bi-prediction the way TComPrediction::xPredInterBi builds it (TComPrediction.cpp:609-652): xPredInterBlk with bi = true from
  // each list into its own TComYuv, then TComYuv::addAvg through xWeightedAverage (:708-724); distortion as in mcDist.
  void fillRef(TComPicYuv& refPic, const Pel* ref0, int refStride, int picW, int picH, int margin)
  {
    refPic.create(picW, picH, CHROMA_400, 64, 64, 4, true);
    const int m = refPic.getMarginX(COMPONENT_Y);
    for (int y = -m; y < picH + m; y++)
      for (int x = -m; x < picW + m; x++)
      {
        const int sy = y < -margin ? -margin : (y >= picH + margin ? picH + margin - 1 : y);
        const int sx = x < -margin ? -margin : (x >= picW + margin ? picW + margin - 1 : x);
        refPic.getAddr(COMPONENT_Y)[y * refPic.getStride(COMPONENT_Y) + x] = ref0[sy * refStride + sx];
      }
  }
  uint32_t mcBiDist(int kind, Pel* org, int orgStride, int w, int h, int bitDepth, const Pel* refA, const Pel* refB, int refStride,
                    int picW, int picH, int margin, int pux, int puy, int mv0x, int mv0y, int mv1x, int mv1y)
  {
    TComPicYuv picA, picB;
    fillRef(picA, refA, refStride, picW, picH, margin);
    fillRef(picB, refB, refStride, picW, picH, margin);
    TComPic pic;
    cu.m_pcPic = &pic; cu.m_ctuRsAddr = 0; cu.m_absZIdxInCtu = 0;
    TComYuv a, b, dst;
    a.create(64, 64, CHROMA_400); b.create(64, 64, CHROMA_400); dst.create(64, 64, CHROMA_400);
    TComMv m0((Short)(mv0x + 4 * pux), (Short)(mv0y + 4 * puy)), m1((Short)(mv1x + 4 * pux), (Short)(mv1y + 4 * puy));
    pic.m_apcPicYuv[TComPic::PIC_YUV_REC] = &picA;
    xPredInterBlk(COMPONENT_Y, &cu, &picA, 0, &m0, w, h, &a, true, bitDepth);
    pic.m_apcPicYuv[TComPic::PIC_YUV_REC] = &picB;
    xPredInterBlk(COMPONENT_Y, &cu, &picB, 0, &m1, w, h, &b, true, bitDepth);
    BitDepths bds; bds.recon[0] = bds.recon[1] = bitDepth;
#if O0043_BEST_EFFORT_DECODING
    bds.stream[0] = bds.stream[1] = bitDepth;
#endif
    xWeightedAverage(&a, &b, 0, 0, 0, w, h, &dst, bds);
    uint32_t d;
    if (kind == 0)
      d = rd.getDistPart(bitDepth, dst.getAddr(COMPONENT_Y, 0), dst.getStride(COMPONENT_Y), org, orgStride, w, h, COMPONENT_Y, DF_SAD);
    else
    {
      DistParam dp; dp.bApplyWeight = false;
      rd.setDistParam(dp, bitDepth, org, orgStride, dst.getAddr(COMPONENT_Y, 0), dst.getStride(COMPONENT_Y), w, h, true);
      d = dp.DistFunc(&dp);
    }
    a.destroy(); b.destroy(); dst.destroy();
    pic.m_apcPicYuv[TComPic::PIC_YUV_REC] = NULL;
    cu.m_pcPic = NULL;
    picA.destroy(); picB.destroy();
    return d;
  }

  // TEncSearch::xTZSearch (TEncSearch.cpp:3881) as xMotionEstimation / xPatternSearchFast reach it with FastSearch = 1.
  // The CU only has to answer clipMv(): picture size and max CU size through its slice's SPS, and its own position.
  TComSPS    sps;
  TComSlice  slice;
  TComDataCU cu;
  void tz(Pel* org, int orgStride, int w, int h, int bitDepth, Pel* refAtPu, int refStride,
          int ltx, int lty, int rbx, int rby, uint32_t uiCost, int predx, int predy,
          int cux, int cuy, int picw, int pich, int maxcu, int searchRange, int firstSearchStop, int hasImv, int imvx, int imvy,
          int* mvx, int* mvy, uint32_t* sad)
  {
    sps.setPicWidthInLumaSamples(picw); sps.setPicHeightInLumaSamples(pich);
    sps.setMaxCUWidth(maxcu); sps.setMaxCUHeight(maxcu);
    slice.setSPS(&sps);
    cu.m_pcSlice = &slice; cu.m_uiCUPelX = cux; cu.m_uiCUPelY = cuy;
    cfg.setFastSearch(1);
    cfg.setFastMEAssumingSmootherMVEnabled(firstSearchStop != 0);
    m_iFastSearch = 1;
    m_iSearchRange = searchRange;
    TComPattern pat;
    pat.initPattern(org, w, h, orgStride, bitDepth);
    setCost(uiCost, predx, predy, 2);
    TComMv lt((Short)ltx, (Short)lty), rb((Short)rbx, (Short)rby), mv((Short)predx, (Short)predy), imv((Short)imvx, (Short)imvy);
    Distortion d = 0;
    xTZSearch(&cu, &pat, refAtPu, refStride, &lt, &rb, mv, d, hasImv ? &imv : NULL);
    *mvx = mv.getHor(); *mvy = mv.getVer(); *sad = d;
  }

  // one luma intra prediction the way TComPrediction::predIntraAng (TComPrediction.cpp:412-494) produces it, from a
  // (2n+1)-strided reference-sample buffer laid out like m_piYuvExt (row 0 = corner / above / above-right, column 0 =
  // left / below-left).  predIntraAng itself needs a TComTU; its body for the non-DPCM case is these three protected calls.
  void intraPredict(int mode, const Pel* buf, int n, int bitDepth, bool above, bool left, bool edgeFilters, Pel* dst, int dstStride)
  {
    const Int sw = 2 * n + 1;
    const Pel* src = buf + sw + 1;
    if (mode == PLANAR_IDX) xPredIntraPlanar(src, sw, dst, dstStride, n, n);
    else
    {
      xPredIntraAng(bitDepth, src, sw, dst, dstStride, n, n, CHANNEL_TYPE_LUMA, mode, above, left, edgeFilters);
      if (mode == DC_IDX && above && left) xDCPredFiltering(src, sw, dst, dstStride, n, n, CHANNEL_TYPE_LUMA);
    }
  }
};

bool g_rom = false;

} // namespace

extern "C" {

// Job / result records shared with oracle/hm_oracle.c, include/hmb200.h (same field order) so that one
// numpy dtype serves the reference, the oracle and the CUDA path.
struct hmref_job    { int32_t pu_x, pu_y, w, h, lt_x, lt_y, rb_x, rb_y, pred_x, pred_y; uint32_t lambda_cost; int32_t reserved; };
struct hmref_result { int32_t mv_x, mv_y; uint32_t sad; int32_t half_x, half_y, qter_x, qter_y; uint32_t frac_cost; };

void* hmref_create(int fen, int hadme)
{
  if (!g_rom) { initROM(); g_rom = true; }
  return new Harness(fen, hadme);
}
void hmref_destroy(void* h) { delete static_cast<Harness*>(h); }

// kind: 0 = SAD via the integer-ME setDistParam (TComRdCost.cpp:306) with iSubShift = sub_shift,
//       1 = SSE via getDistPart(DF_SSE) (TComRdCost.cpp:429),
//       2 = HAD via the sub-pel setDistParam with bHADME (TComRdCost.cpp:338),
//       3 = SAD via the sub-pel setDistParam without HAD (DF_SADS*)
uint32_t hmref_dist(void* hv, int kind, const int16_t* org, int org_stride, const int16_t* cur, int cur_stride,
                    int w, int h, int bit_depth, int sub_shift)
{
  Harness* H = static_cast<Harness*>(hv);
  Pel* o = const_cast<Pel*>(org); Pel* c = const_cast<Pel*>(cur);
  if (kind == 1)
    return H->rd.getDistPart(bit_depth, c, cur_stride, o, org_stride, w, h, COMPONENT_Y, DF_SSE);
  TComPattern pat; pat.initPattern(o, w, h, org_stride, bit_depth);
  DistParam dp; dp.bApplyWeight = false;
  if (kind == 0)      H->rd.setDistParam(&pat, c, cur_stride, dp);
  else                H->rd.setDistParam(&pat, c, cur_stride, 1, dp, kind == 2);
  dp.iSubShift = sub_shift;
  dp.bitDepth  = bit_depth;
  dp.pCur      = c;
  return dp.DistFunc(&dp);
}

uint32_t hmref_get_cost(void* hv, uint32_t ui_cost, int pred_x, int pred_y, int scale, int x, int y)
{
  Harness* H = static_cast<Harness*>(hv);
  H->setCost(ui_cost, pred_x, pred_y, scale);
  return H->rd.getCost(x, y);
}
uint32_t hmref_get_bits(void* hv, int pred_x, int pred_y, int scale, int x, int y)
{
  Harness* H = static_cast<Harness*>(hv);
  H->setCost(0, pred_x, pred_y, scale);
  return H->rd.getBits(x, y);
}

void hmref_pattern_search(void* hv, const int16_t* org, int org_stride, int w, int h, int bit_depth,
                          const int16_t* ref_at_pu, int ref_stride, int lt_x, int lt_y, int rb_x, int rb_y,
                          uint32_t ui_cost, int pred_x, int pred_y, int* mv_x, int* mv_y, uint32_t* sad)
{
  static_cast<Harness*>(hv)->search(const_cast<Pel*>(org), org_stride, w, h, bit_depth, const_cast<Pel*>(ref_at_pu),
                                    ref_stride, lt_x, lt_y, rb_x, rb_y, ui_cost, pred_x, pred_y, mv_x, mv_y, sad);
}

void hmref_pattern_search_frac(void* hv, int lossless, const int16_t* org, int org_stride, int w, int h, int bit_depth,
                               const int16_t* ref_at_pu, int ref_stride, int mv_x, int mv_y,
                               uint32_t ui_cost, int pred_x, int pred_y,
                               int* half_x, int* half_y, int* qter_x, int* qter_y, uint32_t* cost)
{
  static_cast<Harness*>(hv)->frac(lossless, const_cast<Pel*>(org), org_stride, w, h, bit_depth,
                                  const_cast<Pel*>(ref_at_pu), ref_stride, mv_x, mv_y, ui_cost, pred_x, pred_y,
                                  half_x, half_y, qter_x, qter_y, cost);
}

// Runs xPatternSearch (+ xPatternSearchFracDIF when do_frac) over a job list on padded Pel planes.
// cur0 / ref0 point at luma sample (0,0) of the current (original) and reference planes.
// Returns elapsed CPU seconds (clock()), used as the reference arm of bench.py.
double hmref_run_jobs(void* hv, const int16_t* cur0, int cur_stride, const int16_t* ref0, int ref_stride,
                      int bit_depth, const hmref_job* jobs, int njobs, int do_frac, hmref_result* out)
{
  Harness* H = static_cast<Harness*>(hv);
  clock_t t0 = clock();
  for (int i = 0; i < njobs; i++)
  {
    const hmref_job& j = jobs[i];
    Pel* org = const_cast<Pel*>(cur0) + j.pu_y * cur_stride + j.pu_x;
    Pel* ref = const_cast<Pel*>(ref0) + j.pu_y * ref_stride + j.pu_x;
    hmref_result r; memset(&r, 0, sizeof(r));
    H->search(org, cur_stride, j.w, j.h, bit_depth, ref, ref_stride, j.lt_x, j.lt_y, j.rb_x, j.rb_y,
              j.lambda_cost, j.pred_x, j.pred_y, &r.mv_x, &r.mv_y, &r.sad);
    if (do_frac)
      H->frac(0, org, cur_stride, j.w, j.h, bit_depth, ref, ref_stride, r.mv_x, r.mv_y,
              j.lambda_cost, j.pred_x, j.pred_y, &r.half_x, &r.half_y, &r.qter_x, &r.qter_y, &r.frac_cost);
    out[i] = r;
  }
  return double(clock() - t0) / CLOCKS_PER_SEC;
}

// TComRdCost::setLambda + getMotionCost(true, 0, false): the m_uiCost the searches multiply MV bits with
uint32_t hmref_motion_lambda_cost(double lambda, int bit_depth)
{
  TComRdCost rd;
  BitDepths bd; bd.recon[0] = bd.recon[1] = bit_depth;
#if O0043_BEST_EFFORT_DECODING
  bd.stream[0] = bd.stream[1] = bit_depth;
#endif
  rd.setLambda(lambda, bd);
  rd.getMotionCost(true, 0, false);
  return rd.m_uiCost;
}

uint32_t hmref_mc_dist(void* hv, int kind, const int16_t* org, int org_stride, int w, int h, int bit_depth, const int16_t* ref0,
                       int ref_stride, int pic_w, int pic_h, int margin, int pu_x, int pu_y, int mv_x, int mv_y)
{
  return static_cast<Harness*>(hv)->mcDist(kind, const_cast<Pel*>(org), org_stride, w, h, bit_depth, ref0, ref_stride, pic_w, pic_h,
                                           margin, pu_x, pu_y, mv_x, mv_y);
}

uint32_t hmref_mc_bi_dist(void* hv, int kind, const int16_t* org, int org_stride, int w, int h, int bit_depth, const int16_t* ref_a,
                          const int16_t* ref_b, int ref_stride, int pic_w, int pic_h, int margin, int pu_x, int pu_y, int mv0_x, int mv0_y,
                          int mv1_x, int mv1_y)
{
  return static_cast<Harness*>(hv)->mcBiDist(kind, const_cast<Pel*>(org), org_stride, w, h, bit_depth, ref_a, ref_b, ref_stride, pic_w, pic_h,
                                             margin, pu_x, pu_y, mv0_x, mv0_y, mv1_x, mv1_y);
}

// (UInt) calcRdCost(bits, dist, false, DF_SAD) with m_uiLambdaMotionSAD[0] = lambda_motion_sad: what xGetTemplateCost adds to the SAD
uint32_t hmref_template_rd_cost(uint32_t bits, uint32_t dist, double lambda, int bit_depth)
{
  TComRdCost rd;
  rd.init();
  BitDepths bd; bd.recon[0] = bd.recon[1] = bit_depth;
#if O0043_BEST_EFFORT_DECODING
  bd.stream[0] = bd.stream[1] = bit_depth;
#endif
  rd.setLambda(lambda, bd);
  return (uint32_t)rd.calcRdCost(bits, dist, false, DF_SAD);
}

void hmref_tz_search(void* hv, const int16_t* org, int org_stride, int w, int h, int bit_depth,
                     const int16_t* ref_at_pu, int ref_stride, int lt_x, int lt_y, int rb_x, int rb_y,
                     uint32_t ui_cost, int pred_x, int pred_y, int cu_x, int cu_y, int pic_w, int pic_h, int max_cu,
                     int search_range, int first_search_stop, int has_imv, int imv_x, int imv_y, int* mv_x, int* mv_y, uint32_t* sad)
{
  static_cast<Harness*>(hv)->tz(const_cast<Pel*>(org), org_stride, w, h, bit_depth, const_cast<Pel*>(ref_at_pu), ref_stride,
                                lt_x, lt_y, rb_x, rb_y, ui_cost, pred_x, pred_y, cu_x, cu_y, pic_w, pic_h, max_cu, search_range,
                                first_search_stop, has_imv, imv_x, imv_y, mv_x, mv_y, sad);
}

// TVideoIOYuv::open + read (TLibVideoIO/TVideoIOYuv.cpp:118-245, 680-741) on a planar 4:0:0 file: returns the luma
// plane the encoder would see (conformance padding pad_x / pad_y, bit-depth scaling), without margins.
int hmref_read_luma(const char* path, int width, int height, int pad_x, int pad_y, int file_bit_depth, int internal_bit_depth,
                    int16_t* dst, int dst_stride)
{
  if (!g_rom) { initROM(); g_rom = true; }
  TVideoIOYuv io;
  Int fileBD[MAX_NUM_CHANNEL_TYPE], msbBD[MAX_NUM_CHANNEL_TYPE], intBD[MAX_NUM_CHANNEL_TYPE];
  for (int c = 0; c < MAX_NUM_CHANNEL_TYPE; c++) { fileBD[c] = file_bit_depth; msbBD[c] = file_bit_depth; intBD[c] = internal_bit_depth; }
  io.open(const_cast<char*>(path), false, fileBD, msbBD, intBD);
  TComPicYuv pic;
  pic.create(width + pad_x, height + pad_y, CHROMA_400, 64, 64, 4, true);
  Int pad[2] = { pad_x, pad_y };
  const bool ok = io.read(&pic, &pic, IPCOLOURSPACE_UNCHANGED, pad, CHROMA_400, false);
  if (ok)
    for (int y = 0; y < height + pad_y; y++)
      memcpy(dst + y * dst_stride, pic.getAddr(COMPONENT_Y) + y * pic.getStride(COMPONENT_Y), sizeof(int16_t) * (size_t)(width + pad_x));
  pic.destroy();
  io.close();
  return ok ? 0 : -1;
}


// 35 luma intra predictions + Hadamard distortions of one block (TEncSearch::estIntraPredQT first pass, TEncSearch.cpp:
// 2270-2296).  top / left: reference lines with the corner at index 0 (2n+1 samples each), unfiltered and smoothed.
int hmref_intra_use_filtered(int mode, int n)
{
  return TComPrediction::filteringIntraReferenceSamples(COMPONENT_Y, mode, n, n, CHROMA_420, false) ? 1 : 0;
}
void hmref_intra_predict(void* hv, int mode, const int16_t* top, const int16_t* left, int n, int bit_depth, int above_ok, int left_ok,
                         int edge_filters, int16_t* dst, int dst_stride)
{
  Harness* H = static_cast<Harness*>(hv);
  const int sw = 2 * n + 1;
  std::vector<Pel> buf((size_t)sw * sw, 0);
  for (int i = 0; i < sw; i++) buf[i] = top[i];
  for (int i = 1; i < sw; i++) buf[(size_t)i * sw] = left[i];
  H->intraPredict(mode, buf.data(), n, bit_depth, above_ok != 0, left_ok != 0, edge_filters != 0, dst, dst_stride);
}
void hmref_intra_modes_had(void* hv, const int16_t* org, int org_stride, const int16_t* top_unf, const int16_t* left_unf,
                           const int16_t* top_flt, const int16_t* left_flt, int n, int bit_depth, uint32_t* out)
{
  Harness* H = static_cast<Harness*>(hv);
  std::vector<Pel> pred((size_t)n * n);
  DistParam dp;
  H->rd.setDistParam(dp, bit_depth, const_cast<Pel*>(org), org_stride, pred.data(), n, n, n, true);      // TEncSearch.cpp:2268
  dp.bApplyWeight = false;
  for (int mode = 0; mode < 35; mode++)
  {
    const bool f = hmref_intra_use_filtered(mode, n) != 0;
    hmref_intra_predict(hv, mode, f ? top_flt : top_unf, f ? left_flt : left_unf, n, bit_depth, 1, 1, 1, pred.data(), n);
    out[mode] = dp.DistFunc(&dp);
  }
}

} // extern "C"
