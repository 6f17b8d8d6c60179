/* hmb200.h — C-ABI of the B200-native HM-16.5 integer-pel motion search + block-distortion path.
 *
 * This is the drop-in boundary (plain pointers and sizes, no C++/torch types).  Every entry point names the
 * reference interface it replaces; paths are relative to /root/reference/hm-16.5rc1/source/Lib/.
 * The reference has no plugin API: the boundary is three C++ member functions plus the distortion function
 * pointer table, so the entry points mirror those signatures with POD stand-ins for DistParam / TComPattern /
 * TComMv.  INTEGRATION.md shows the forwarders a maintainer adds to TComRdCost.cpp / TEncSearch.cpp.
 *
 * There is NO CPU fallback: every compute entry returns HMB200_ERR_CUDA (and hmb200_last_error() explains) when
 * no sm_100 device is usable.  Loading the library and querying symbols does not need a GPU.
 */
#ifndef HMB200_H
#define HMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMB200_OK            0
#define HMB200_ERR_CUDA     -1   /* CUDA runtime/driver failure or no device */
#define HMB200_ERR_ARG      -2   /* bad argument (unsupported size, unknown plane, pointer outside planes ...) */
#define HMB200_ERR_STATE    -3   /* hmb200_init not called */

/* ---- flags of the search entries (TEncCfg getters the callee reads, TLibEncoder/TEncCfg.h:570,577) ---- */
#define HMB200_FLAG_FEN      1   /* getUseFastEnc(): iSubShift = 1 for PUs with more than 8 rows (TEncSearch.cpp:3804-3810) */
#define HMB200_FLAG_HADME    2   /* getUseHADME(): sub-pel refinement uses xGetHADs (TComRdCost.cpp:355-374)           */
#define HMB200_FLAG_FRAC     4   /* also run the quarter-pel refinement (xPatternSearchFracDIF)                          */
#define HMB200_FLAG_TZ       8   /* integer search = xTZSearch (FastSearch = 1) instead of the full search                */
#define HMB200_FLAG_TZ_STOP 16   /* getFastMEAssumingSmootherMVEnabled(): first search stops after 3 idle diamonds (TEncSearch.cpp:304,3962) */

/* ---- distortion function families, the rows of m_afpDistortFunc (TLibCommon/TypeDef.h:334-378) ---- */
#define HMB200_DF_SAD        0   /* DF_SAD*  : xGetSAD4..64/12/24/48 honouring iSubShift (TComRdCost.cpp:489-953) */
#define HMB200_DF_SSE        1   /* DF_SSE*  : xGetSSE* (TComRdCost.cpp:959-1304)                                 */
#define HMB200_DF_HADS       2   /* DF_HADS* : xGetHADs (TComRdCost.cpp:1526-1593)                                */
#define HMB200_DF_SADS       3   /* DF_SADS* : same functions as DF_SAD (TComRdCost.cpp:248-254)                  */

#define HMB200_PLANE_ORG     0
#define HMB200_PLANE_REC     1

/* TComMv (TLibCommon/TComMv.h:50-55): two Shorts; widened to int32 at the boundary. */
typedef struct { int32_t x, y; } hmb200_mv;

/* POD mirror of DistParam (TLibCommon/TComRdCost.h:67-101).  bApplyWeight must be 0 (weighted prediction is out of
 * scope: WeightedPredP defaults to false, App/TAppEncoder/TAppEncCfg.cpp:897); iStep must be 1 (asserted by the
 * reference, TComRdCost.cpp:1314,1338,1433). */
typedef struct {
  const int16_t* pOrg;       /* Pel* */
  const int16_t* pCur;
  int32_t iStrideOrg, iStrideCur;
  int32_t iRows, iCols;
  int32_t iStep;
  int32_t func;              /* HMB200_DF_* — stands in for the DistFunc pointer */
  int32_t bitDepth;
  int32_t bApplyWeight;
  int32_t iSubShift;
} hmb200_dist_param;

/* One distortion evaluation between two blocks of REGISTERED planes (device-resident; no host copies). */
typedef struct {
  int32_t org_plane, org_x, org_y;
  int32_t cur_plane, cur_x, cur_y;
  int32_t w, h;
  int32_t sub_shift;
  int32_t reserved;
} hmb200_dist_desc;

/* POD mirror of TComPattern's luma ROI (TLibCommon/TComPattern.h:54-99). */
typedef struct {
  const int16_t* roi;        /* getROIY()            */
  int32_t width, height;     /* getROIYWidth/Height  */
  int32_t stride;            /* getPatternLStride()  */
  int32_t bit_depth;         /* getBitDepthY()       */
} hmb200_pattern;

/* The TComRdCost state xPatternSearch / xPatternRefinement read through getCost()
 * (TLibCommon/TComRdCost.h:118-130,172-189): m_uiCost, m_mvPredictor.  The cost scale is fixed by the callee
 * (2 integer, 1 half, 0 quarter: TEncSearch.cpp:3722,3746,4267). */
typedef struct {
  uint32_t  lambda_cost;     /* m_uiCost after getMotionCost(true, 0, ...) */
  hmb200_mv pred;            /* m_mvPredictor, quarter-pel                 */
} hmb200_cost_state;

/* One PU search of the batched form.  pu_x/pu_y: luma position of the PU in the picture; the window corners are
 * xSetSearchRange's outputs (integer pel, inclusive; TEncSearch.cpp:3765-3781).  48 bytes. */
typedef struct {
  int32_t pu_x, pu_y, w, h;
  int32_t lt_x, lt_y, rb_x, rb_y;
  int32_t pred_x, pred_y;    /* quarter-pel, unclipped AMVP predictor (TEncSearch.cpp:3721) */
  uint32_t lambda_cost;
  int32_t reserved;
} hmb200_pu_job;

/* 32 bytes.  mv/sad = xPatternSearch outputs (rcMv, ruiSAD without MV cost); half/qter/frac_cost =
 * xPatternSearchFracDIF outputs (rcMvHalf, rcMvQter, ruiCost with MV bits). */
typedef struct {
  int32_t mv_x, mv_y;
  uint32_t sad;
  int32_t half_x, half_y, qter_x, qter_y;
  uint32_t frac_cost;
} hmb200_pu_result;

/* ------------------------------------------------------------------ lifetime ---------------------------------- */

/* Replaces nothing in the reference; called from TEncSearch::init (TEncSearch.cpp:201) where picture and CTU size
 * are known.  Selects the device, creates streams and pinned staging.  Idempotent per device.  This is the process-wide
 * DEFAULT context: the single-GPU encoder needs nothing else. */
int  hmb200_init(int device);
void hmb200_shutdown(void);                 /* tears down the calling thread's current context's device state */
const char* hmb200_last_error(void);        /* last error of the calling thread */

/* Multi-GPU inside one process (lookahead frame pairs / tile columns, "one host thread and stream set per
 * GPU"; the reference's frame loop is TAppEncTop::encode, App/TAppEncoder/TAppEncTop.cpp:478-520).  A context owns one
 * device's streams, planes and staging; every entry point of this header works on the CALLING THREAD'S CURRENT context.
 * hmb200_ctx_create makes the new context current on the calling thread; hmb200_ctx_set_current(ctx) binds one to another
 * thread (NULL = the default context of hmb200_init) and selects its device.  A context is used by one thread at a time;
 * different contexts run concurrently on their own threads.  Plane ids and prepared handles belong to the context they
 * were made in (a handle used under another context fails with HMB200_ERR_STATE). */
typedef struct hmb200_ctx hmb200_ctx;
hmb200_ctx* hmb200_ctx_create(int device);
void hmb200_ctx_destroy(hmb200_ctx* ctx);
int  hmb200_ctx_set_current(hmb200_ctx* ctx);
hmb200_ctx* hmb200_ctx_get_current(void);
int  hmb200_ctx_device(const hmb200_ctx* ctx);
/* Page-locked host memory for result arrays (hmb200_fetch_results copies straight into it, without the staging copy it
 * needs for pageable memory) and for planes handed to hmb200_register_plane*. */
void* hmb200_host_alloc(size_t bytes);
void  hmb200_host_free(void* p);
/* Number of kernel launches issued by this library since init (bench.py's gpu_launches). */
uint64_t hmb200_launch_count(void);
/* 1:1 entries of the calling thread's context: hmb200_pattern_search_and_refine calls so far, how many of them launched a whole-CU
 * search (a 2Nx2N PU: the CU's other partitions - 2NxN, Nx2N, AMP; TLibCommon/TComDataCU.cpp:1893-1931 - are searched and refined in
 * the same round trip with that call's window, predictor and lambda), and how many were answered from such a launch because their
 * window, predictor, lambda, flags, position and pattern samples were exactly those (HMB200_NO_SPECULATION=1 switches this off). */
void hmb200_one_call_stats(uint64_t* calls, uint64_t* cu_launches, uint64_t* served_from_cu);

/* ------------------------------------------------------------------ host-side window / job-list logic ---------- */

/* m_uiCost of a slice: TComRdCost::setLambda (TLibCommon/TComRdCost.cpp:195-220) turns the slice's motion lambda into the
 * fixed-point multiplier floor(65536 * sqrt(lambda)) (m_uiLambdaMotionSAD[0]) that getMotionCost(true, 0, ...) installs
 * and getCost() multiplies the MV bits with; it is the lambda_cost of hmb200_cost_state / hmb200_pu_job. */
uint32_t hmb200_motion_lambda_cost(double lambda);
/* TEncSearch::xSetSearchRange (TLibEncoder/TEncSearch.cpp:3765-3781) with TComDataCU::clipMv
 * (TLibCommon/TComDataCU.cpp:2788-2801): pred is the quarter-pel predictor, (cu_x, cu_y) the luma origin of the CU
 * that owns the PU.  Pure host arithmetic (no GPU needed). */
void hmb200_set_search_range(hmb200_mv pred, int search_range, int cu_x, int cu_y, int pic_w, int pic_h,
                             int max_cu_w, int max_cu_h, hmb200_mv* lt, hmb200_mv* rb);
/* Canonical all-PU job list of a picture: every CU of depth 0..3 that lies inside the
 * picture, partition modes 2Nx2N / 2NxN / Nx2N at every depth plus the four AMP modes at depths 0..2
 * (TLibCommon/TComDataCU.cpp:1893-1931 getPartIndexAndSize), 593 PUs per 64x64 CTU, predictor `pred` for every PU,
 * window from hmb200_set_search_range.  Writes at most `capacity` jobs; returns the total count (call with
 * capacity 0 to size the buffer).  ctu_first/ctu_count select a raster range of CTUs (ctu_count < 0: to the end). */
int  hmb200_build_canonical_jobs(int pic_w, int pic_h, int max_cu, int search_range, uint32_t lambda_cost, hmb200_mv pred,
                                 int ctu_first, int ctu_count, hmb200_pu_job* jobs, int capacity);

/* Same for a rectangle of CTUs [ctu_x0, ctu_x1) x [ctu_y0, ctu_y1) (CTU units): one tile column / CTU row of a picture.
 * Jobs come out in CTU raster order inside the rectangle. */
int  hmb200_build_canonical_jobs_rect(int pic_w, int pic_h, int max_cu, int search_range, uint32_t lambda_cost, hmb200_mv pred,
                                      int ctu_x0, int ctu_x1, int ctu_y0, int ctu_y1, hmb200_pu_job* jobs, int capacity);
/* CTU column range of tile column `column` of `n_columns` uniformly spaced tile columns
 * (TileUniformSpacing, TLibCommon/TComPicSym.cpp:217-229): the multi-GPU tile-column shard of a picture. */
int  hmb200_tile_column_range(int pic_w, int max_cu, int n_columns, int column, int* ctu_x0, int* ctu_x1);

/* ------------------------------------------------------------------ planes ------------------------------------ */

/* Uploads one luma plane once per frame.  host_origin points at sample (0,0) of a TComPicYuv luma buffer
 * (Pel = int16, TLibCommon/TComPicYuv.cpp:81-143: stride = width + 2*margin_x); the margins must already be
 * extended (TComPicYuv::extendPicBorder, :197-242) when they will be searched.  8-bit planes are narrowed to uint8
 * on upload, deeper ones to uint16.  The host range [origin - margin, ...] is remembered so that the 1:1 search
 * entries can translate a `Pel*` into (plane, x, y).  Returns a plane id >= 0 or an error code.
 * Call sites: after TComSlice::setRefPicList (TLibCommon/TComSlice.cpp:351-377) for reconstructed pictures,
 * TEncTop::encode (TLibEncoder/TEncTop.cpp:323-325) for originals. */
int  hmb200_register_plane(const int16_t* host_origin, int stride, int width, int height,
                           int margin_x, int margin_y, int bit_depth, int kind, int poc);
/* Same, from 8-bit samples without margins (planar YUV luma as read by TVideoIOYuv); the margins are synthesised on
 * the device exactly like extendPicBorder.  A tightly packed frame in page-locked memory (hmb200_host_alloc) is uploaded
 * on a separate stream and the call returns at once - the copy then runs behind the kernels of the frame pair before it;
 * every consumer of the plane is ordered after it on the device, the source must stay untouched until hmb200_sync (or a
 * blocking call that uses the plane) returns.  hmb200_release_plane is stream-ordered as well: the buffer is recycled
 * behind its last reader, without a host synchronisation. */
int  hmb200_register_plane_u8(const uint8_t* host_samples, int stride, int width, int height,
                              int margin_x, int margin_y, int kind, int poc);
/* Same for 9..14-bit content held as 16-bit samples (planar 16-bit YUV luma as TVideoIOYuv reads it for Main10,
 * TLibVideoIO/TVideoIOYuv.cpp:247-377); stride in samples. */
int  hmb200_register_plane_u16(const uint16_t* host_samples, int stride, int width, int height,
                               int margin_x, int margin_y, int bit_depth, int kind, int poc);
/* Direct ingest of one frame's luma from a planar YUV file image: TVideoIOYuv::read for COMPONENT_Y
 * (TLibVideoIO/TVideoIOYuv.cpp:680-741 -> readPlane :247-377 -> scalePlane :70-99) fused with extendPicBorder.
 * file_luma: width x height samples, bytes or (file_is16) 16-bit little endian; pad_x / pad_y: conformance padding
 * (aiPad, replicates the last column / row; e.g. 1080 -> 1088 rows); samples are scaled by
 * 2^(internal_bit_depth - file_bit_depth) (negative: round, shift, clip).  The registered plane has the coded size
 * (width + pad_x) x (height + pad_y).  No host-side Pel conversion, one H2D copy of the file bytes. */
int  hmb200_register_plane_yuv(const void* file_luma, int file_is16, int width, int height, int pad_x, int pad_y,
                               int file_bit_depth, int internal_bit_depth, int margin_x, int margin_y, int kind, int poc);
/* Reads a registered plane back (including margins) as Pel samples; dst_stride >= width + 2*margin_x. */
int  hmb200_read_plane(int plane_id, int16_t* dst_origin, int dst_stride);
void hmb200_release_plane(int plane_id);

/* ------------------------------------------------------------------ distortion table -------------------------- */

/* Signature-compatible with FpDistFunc (TLibCommon/TComRdCost.h:60): Distortion (*)(DistParam*).  Host buffers are
 * copied to the device per call (latency-bound; for parity and for the table hook of TComRdCost::init,
 * TComRdCost.cpp:224-276).  Aborts the process on CUDA failure like the reference aborts on assert. */
uint32_t hmb200_dist(const hmb200_dist_param* p);
/* n evaluations of one family between registered planes, one launch. */
int  hmb200_dist_batch(int func, int bit_depth, int n, const hmb200_dist_desc* descs, uint32_t* out);

/* One motion-compensated prediction: the PU and its quarter-pel MV as handed to xPredInterBlk (the caller clips it,
 * TComDataCU::clipMv).  24 bytes. */
typedef struct { int32_t pu_x, pu_y, w, h, mv_x, mv_y; } hmb200_mc_desc;
/* Distortion between each original PU (cur_plane) and its uni-directional motion-compensated prediction from ref_plane:
 * TComPrediction::xPredInterBlk(COMPONENT_Y, ..., bi = false) (TLibCommon/TComPrediction.cpp:668-706) followed by the
 * distortion function — func = HMB200_DF_SAD: what TEncSearch::xGetTemplateCost takes through getDistPart(DF_SAD)
 * (AMVP candidate selection, TLibEncoder/TEncSearch.cpp:3619-3658); func = HMB200_DF_HADS: what
 * TEncSearch::xGetInterPredictionError takes (merge candidate cost, :2809-2830).  One launch set for the whole batch;
 * the MV-index bit cost and calcRdCost stay on the host. */
int  hmb200_mc_dist_batch(int cur_plane, int ref_plane, int func, int n, const hmb200_mc_desc* descs, uint32_t* out);

/* One merge / AMVP candidate of a PU.  inter_dir: 1 = list 0, 2 = list 1, 3 = bi-prediction (the interDir of
 * TComDataCU::getInterMergeCandidates); mv*: quarter pel, clipped by the caller (TComDataCU::clipMv, as xPredInterUni does);
 * ref*_plane: registered planes of the two reference pictures (ignored for an unused list); bits: the candidate's rate term
 * (uiBitsCand of xMergeEstimation / m_auiMVPIdxCost[idx][AMVP_MAX_NUM_CANDS] of xGetTemplateCost).  48 bytes. */
#define HMB200_INTER_DIR_NO_IDENTICAL_CHECK 4   /* or'ed into inter_dir = 3: the caller has already applied xCheckIdenticalMotion (on the
                                                 * unclipped MVs and the POCs, as the reference does); always average the two lists */
typedef struct {
  int32_t pu_x, pu_y, w, h;
  int32_t inter_dir;
  int32_t mv0_x, mv0_y, ref0_plane;
  int32_t mv1_x, mv1_y, ref1_plane;
  int32_t bits;
} hmb200_mc_cand;
/* Prediction error of n candidates in one launch set: TComPrediction::motionCompensation (TLibCommon/TComPrediction.cpp:
 * 539-586) - uni-directional xPredInterBlk, or xPredInterBi: xPredInterBlk(bi = true) per list + TComYuv::addAvg
 * (TComPrediction.cpp:609-652, 708-724; TComYuv.cpp:352-407) - followed by the distortion: func = HMB200_DF_HADS as
 * TEncSearch::xGetInterPredictionError (TLibEncoder/TEncSearch.cpp:2805-2826), HMB200_DF_SAD as xGetTemplateCost (:3619-3658).
 * A bi-directional candidate whose lists name the same plane with the same MV is predicted from list 0 alone
 * (xCheckIdenticalMotion, TComPrediction.cpp:496-517).  Weighted prediction is out of scope. */
int  hmb200_mc_cand_dist_batch(int cur_plane, int func, int n, const hmb200_mc_cand* cands, uint32_t* dist);
/* The candidate loop of TEncSearch::xMergeEstimation (TLibEncoder/TEncSearch.cpp:2868-2892) for n_pu PUs at once: the candidates
 * of PU i are cands[cand_first[i] .. cand_first[i+1]) in merge-index order (cand_first has n_pu + 1 entries);
 * cost = error + getCost(bits) = error + ((lambda_cost * bits) >> 16) in UInt arithmetic, strict '<'.  use_hadme = getUseHADME().
 * best_cand[i] = merge index inside the PU's list, best_cost[i] = ruiCost.  cand_dist (optional, may be NULL) receives every
 * candidate's error.  getInterMergeCandidates / xRestrictBipredMergeCand stay on the host (they read the CU's neighbours). */
int  hmb200_merge_estimation_batch(int cur_plane, int n_pu, const int32_t* cand_first, const hmb200_mc_cand* cands, int use_hadme,
                                   uint32_t lambda_cost, uint32_t* best_cand, uint32_t* best_cost, uint32_t* cand_dist);
/* The candidate loop of TEncSearch::xEstimateMvPredAMVP (:3457-3469) over xGetTemplateCost: uni-directional prediction from
 * the candidate MV (list 0 fields of hmb200_mc_cand, inter_dir = 1), SAD, cost = (UInt) calcRdCost(bits, SAD, false, DF_SAD)
 * = SAD + ((bits * lambda_motion_sad) >> 16) (TLibCommon/TComRdCost.cpp:67-73, 99-108; lambda_motion_sad = m_uiLambdaMotionSAD[0]),
 * `uiBestCost > uiTmpCost` keeps the first minimum.  best_cand[i] = iBestIdx, best_cost[i] = uiBestCost (= *puiDistBiP). */
int  hmb200_amvp_estimation_batch(int cur_plane, int n_pu, const int32_t* cand_first, const hmb200_mc_cand* cands,
                                  uint32_t lambda_motion_sad, uint32_t* best_cand, uint32_t* best_cost, uint32_t* cand_dist);

/* ------------------------------------------------------------------ intra first pass -------------------------- */

/* One block of the intra mode pre-selection.  (x, y): position in the registered ORIGINAL plane; n = 4, 8, 16, 32 or 64;
 * ref_off: index (in samples) of the block's reference lines in `refs`: top_unf[2n+1], left_unf[2n+1], top_flt[2n+1],
 * left_flt[2n+1], each with the corner sample at index 0 - row 0 and column 0 of the reference's unfiltered / smoothed
 * (2n+1)-strided buffers (m_piYuvExt[COMPONENT_Y][PRED_BUF_UNFILTERED / _FILTERED] after
 * TComPrediction::initIntraPatternChType, TLibCommon/TComPattern.cpp:115-320, which stays on the host: availability and
 * reconstruction state are the encoder's).  flags: bit 0 bAbove, bit 1 bLeft as passed to predIntraAng (both set by
 * initIntraPatternChType, :155-156).  24 bytes. */
typedef struct { int32_t x, y, n, ref_off, flags, reserved; } hmb200_intra_block;
#define HMB200_INTRA_MODES 35
/* The mode loop of TEncSearch::estIntraPredQT (TLibEncoder/TEncSearch.cpp:2270-2296) for nblocks blocks in one launch
 * set: for every mode 0..34 filteringIntraReferenceSamples (TComPattern.cpp:544-568) + predIntraAng (TComPrediction.cpp:
 * 412-494: planar :756-816, DC :183-222 + :819-848, angular :250-410 with the edge filters) + distParam.DistFunc =
 * xGetHADs (TComRdCost.cpp:380-392, 1526-1593).  out[i * 35 + mode] = uiSad of block i; the mode bits
 * (xModeBitsIntra), the cost and the candidate list (:2282-2294) stay on the host. */
int  hmb200_intra_modes_had_batch(int org_plane, int nblocks, const hmb200_intra_block* blocks, const int16_t* refs,
                                  int n_ref_samples, uint32_t* out);
/* 1:1 form with the reference's own buffers: org = piOrg (stride org_stride), ref_unf / ref_flt = the two (2n+1)-strided
 * predictor buffers (getPredictorPtr(COMPONENT_Y, false / true), TComPrediction.h:121), above / left = bAbove / bLeft. */
int  hmb200_intra_modes_had(const int16_t* org, int org_stride, const int16_t* ref_unf, const int16_t* ref_flt, int n,
                            int bit_depth, int above, int left, uint32_t out[HMB200_INTRA_MODES]);

/* ------------------------------------------------------------------ searches, 1:1 ----------------------------- */

/* TEncSearch::xPatternSearch(TComPattern*, Pel* piRefY, Int iRefStride, TComMv* LT, TComMv* RB, TComMv& rcMv,
 * Distortion& ruiSAD)  (TLibEncoder/TEncSearch.h:413-419, .cpp:3786-3843).  ref_at_pu must point into a registered
 * plane.  flags: HMB200_FLAG_FEN. */
int  hmb200_pattern_search(const hmb200_pattern* key, const int16_t* ref_at_pu, int ref_stride,
                           hmb200_mv lt, hmb200_mv rb, const hmb200_cost_state* cs, int flags,
                           hmb200_mv* mv_out, uint32_t* sad_out);
/* TEncSearch::xPatternSearchFast -> xTZSearch (TLibEncoder/TEncSearch.cpp:3847-3875, 3881-4083) for m_iFastSearch == 1,
 * one call = one reference call.  cs->pred is both the MV-cost predictor and the start vector (rcMv on entry);
 * extra: owning CU position and pIntegerMv2Nx2NPred; pic_w/pic_h/max_cu: what clipMv reads from the SPS;
 * search_range: m_iSearchRange.  flags: HMB200_FLAG_FEN | HMB200_FLAG_TZ_STOP.  (hmb200_tz_extra is declared further down.) */
struct hmb200_tz_extra_s;
int  hmb200_pattern_search_tz(const hmb200_pattern* key, const int16_t* ref_at_pu, int ref_stride, hmb200_mv lt, hmb200_mv rb,
                              const hmb200_cost_state* cs, int flags, const struct hmb200_tz_extra_s* extra, int pic_w, int pic_h,
                              int max_cu, int search_range, hmb200_mv* mv_out, uint32_t* sad_out);
/* TEncSearch::xPatternSearchFracDIF(Bool bIsLosslessCoded, TComPattern*, Pel*, Int, TComMv* pcMvInt,
 * TComMv& rcMvHalf, TComMv& rcMvQter, Distortion& ruiCost)  (TEncSearch.h:421-430, .cpp:4240-4276).
 * flags: HMB200_FLAG_HADME. */
int  hmb200_pattern_search_frac(int lossless, const hmb200_pattern* key, const int16_t* ref_at_pu, int ref_stride,
                                hmb200_mv mv_int, const hmb200_cost_state* cs, int flags,
                                hmb200_mv* half_out, hmb200_mv* qter_out, uint32_t* cost_out);

/* Both calls of TEncSearch::xMotionEstimation for one PU in ONE device round trip (TLibEncoder/TEncSearch.cpp:3728 and :3749:
 * xPatternSearch, then xPatternSearchFracDIF on the integer MV it found, same pattern, reference, predictor and m_uiCost).
 * The in-encoder frontend calls this from the xPatternSearch forwarder and hands the refinement's outputs to the
 * xPatternSearchFracDIF forwarder that follows (integration/hm_shim.cpp); flags: HMB200_FLAG_FEN | HMB200_FLAG_HADME.
 * Results are those of the two separate entries, always.  Internally a call for a square PU of 8 / 16 / 32 / 64 samples
 * also searches and refines the other partitions of a CU of that size at that position - and of the four 8x8 child CUs of a
 * 16x16 CU - (the PUs HM's partition loop asks for next, TLibCommon/TComDataCU.cpp:1893-1931) and a later call is answered from that launch iff its window, predictor,
 * lambda, flags, position and pattern samples are identical to the launch's; hmb200_one_call_stats counts both. */
int  hmb200_pattern_search_and_refine(const hmb200_pattern* key, const int16_t* ref_at_pu, int ref_stride,
                                      hmb200_mv lt, hmb200_mv rb, const hmb200_cost_state* cs, int flags,
                                      hmb200_mv* mv_out, uint32_t* sad_out, hmb200_mv* half_out, hmb200_mv* qter_out, uint32_t* frac_cost_out);

/* The same for FastSearch = 1: xTZSearch, then xPatternSearchFracDIF on its MV.  flags: FEN | TZ_STOP | HADME. */
int  hmb200_pattern_search_tz_and_refine(const hmb200_pattern* key, const int16_t* ref_at_pu, int ref_stride, hmb200_mv lt, hmb200_mv rb,
                                         const hmb200_cost_state* cs, int flags, const struct hmb200_tz_extra_s* extra, int pic_w, int pic_h,
                                         int max_cu, int search_range, hmb200_mv* mv_out, uint32_t* sad_out, hmb200_mv* half_out,
                                         hmb200_mv* qter_out, uint32_t* frac_cost_out);

/* ------------------------------------------------------------------ searches, batched ------------------------- */

/* Many PUs of one (current, reference) plane pair.  Jobs are grouped by CTU and search window on the host side; each
 * group's window is staged once into shared memory with TMA.  flags: FEN | HADME | FRAC.
 * Results are written in job order.  Blocking. */
int  hmb200_me_jobs(int cur_plane, int ref_plane, const hmb200_pu_job* jobs, int njobs, int flags,
                    hmb200_pu_result* results);
/* One CTU row of a picture (the launch granularity the in-encoder frontend uses): jobs must all lie in CTU row
 * `ctu_row` (pu_y / max_cu == ctu_row). */
int  hmb200_me_ctu_row(int cur_plane, int ref_plane, int ctu_row, int max_cu, const hmb200_pu_job* jobs, int njobs,
                       int flags, hmb200_pu_result* results);

/* Device-resident variant used by benchmarks: jobs/results are DEVICE pointers, no host copies, asynchronous on
 * the library's stream; hmb200_sync() waits.  prepared = handle from hmb200_prepare_jobs (host-side grouping and
 * upload of the schedule are done once). */
typedef struct hmb200_prepared hmb200_prepared;
hmb200_prepared* hmb200_prepare_jobs(const hmb200_pu_job* jobs, int njobs, int flags, int bit_depth);
void hmb200_free_prepared(hmb200_prepared* p);
int  hmb200_run_prepared(hmb200_prepared* p, int cur_plane, int ref_plane);
int  hmb200_fetch_results(hmb200_prepared* p, hmb200_pu_result* results);   /* D2H + sync */
/* Transport only (no reference counterpart): the D2H copy of the last run's results on a second stream, so that it runs
 * behind the next frame pair's upload and kernels (alternate two prepared handles).  results must be page-locked
 * (hmb200_host_alloc).  hmb200_fetch_wait blocks until they have arrived; a new hmb200_run_prepared on the same handle
 * is ordered after a pending fetch on the device. */
int  hmb200_fetch_results_async(hmb200_prepared* p, hmb200_pu_result* results);
int  hmb200_fetch_wait(hmb200_prepared* p);
/* The same record in 16 bytes for the D2H link (the MV field is two thirds of a frame pair's host traffic): TComMv
 * components are Shorts (TLibCommon/TComMv.h:53-54), the half / quarter offsets lie in -1..1 (s_acMvRefineH/Q,
 * TLibEncoder/TEncSearch.cpp:51-75).  Packed on the device behind the run, then copied like hmb200_fetch_results_async. */
typedef struct {
  int16_t mv_x, mv_y;
  uint32_t sad;
  int8_t half_x, half_y, qter_x, qter_y;
  uint32_t frac_cost;
} hmb200_pu_result16;
int  hmb200_fetch_results16_async(hmb200_prepared* p, hmb200_pu_result16* results);
int  hmb200_sync(void);
/* CUDA-event timing of the last hmb200_run_prepared: total and per-kernel milliseconds. */
int  hmb200_last_timing(float* total_ms, float* search_ms, float* frac_ms);
/* Algorithmic work of a prepared job list: candidate-SADs and byte abs-diffs as HM would execute them
 * (SURVEY.md section 8d). */
int  hmb200_prepared_work(const hmb200_prepared* p, uint64_t* cand_sads, uint64_t* abs_diffs);

/* ------------------------------------------------------------------ TZ fast search, batched ------------------- */

/* What xTZSearch reads besides the PU job: the CU that owns the PU (pcCU->clipMv of the predictor and of the re-centred
 * raster range, TLibCommon/TComDataCU.cpp:2788-2801) and pIntegerMv2Nx2NPred (integer pel; TEncSearch.cpp:3926-3946). */
typedef struct hmb200_tz_extra_s {
  int32_t cu_x, cu_y;
  int32_t has_imv, imv_x, imv_y;
  int32_t reserved[3];
} hmb200_tz_extra;

/* cu_x / cu_y of canonical-list PUs (the aligned S x S CU with S = max(w, h)); has_imv = 0. */
void hmb200_canonical_tz_extra(const hmb200_pu_job* jobs, int njobs, hmb200_tz_extra* extra);
/* Turns a list prepared with HMB200_FLAG_TZ into a TZ search: TEncSearch::xTZSearch (TLibEncoder/TEncSearch.cpp:
 * 3881-4083) under TZ_SEARCH_CONFIGURATION (:297-313) (HMB200_FLAG_TZ_STOP = FastMEAssumingSmootherMVEnabled); jobs[i].pred is the AMVP
 * predictor the search starts from (rcMv on entry), jobs[i].lt/rb the window of xSetSearchRange(pred),
 * search_range = m_iSearchRange.  hmb200_run_prepared then runs it (followed by the refinement with FLAG_FRAC). */
int  hmb200_prepared_set_tz(hmb200_prepared* p, const hmb200_tz_extra* extra, int pic_w, int pic_h, int max_cu, int search_range);
/* prepare + set_tz + run + fetch in one blocking call. */
int  hmb200_tz_jobs(int cur_plane, int ref_plane, const hmb200_pu_job* jobs, const hmb200_tz_extra* extra, int njobs,
                    int pic_w, int pic_h, int max_cu, int search_range, int flags, hmb200_pu_result* results);

/* Byte abs-diffs the tiled kernels actually execute for a prepared list, and how many PUs run CU-fused (partial SADs
 * of a CU shared by all its partitions): executed < algorithmic when fusion applies.  Reported next to the roofline
 * so that algorithmic throughput and pipe utilisation are not conflated. */
int  hmb200_prepared_executed_work(const hmb200_prepared* p, uint64_t* abs_diffs_executed, uint64_t* pus_fused);
/* The roofline numerator: byte abs-diffs the shipped algorithm cannot avoid - every visited sample of every CU once per
 * candidate of its window for the CU-fused kernels (no masked lanes, no padding), W*H per candidate for PUs searched alone. */
int  hmb200_prepared_unique_work(const hmb200_prepared* p, uint64_t* abs_diffs_unique);

#ifdef __cplusplus
}
#endif
#endif /* HMB200_H */
