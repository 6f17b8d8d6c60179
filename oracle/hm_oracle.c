/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product library (libhmb200.so) never links, loads or calls it.
 *
 * Plain-C restatement of the HM-16.5 (xkfz007/video_codecs, hm-16.5rc1) integer-pel full search, quarter-pel
 * refinement and block-distortion arithmetic.  Written from the behaviour of the reference, not from its text;
 * every function cites the reference lines whose results it must reproduce (paths relative to
 * hm-16.5rc1/source/Lib/).  PARITY IS PINNED: tests/test_oracle_vs_reference.py checks every function below
 * against the unmodified reference compiled into oracle/_ref/libhmref.so, and tests/golden/ holds vectors
 * generated from that reference (tests/golden/make_golden.py) for machines where /root/reference is absent.
 *
 * Types follow the default reference build (TLibCommon/TypeDef.h:219-230): Pel = int16, Distortion = uint32.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { int32_t pu_x, pu_y, w, h, lt_x, lt_y, rb_x, rb_y, pred_x, pred_y; uint32_t lambda_cost; int32_t reserved; } hmo_job;
typedef struct { int32_t mv_x, mv_y; uint32_t sad; int32_t half_x, half_y, qter_x, qter_y; uint32_t frac_cost; } hmo_result;

/* ---------------------------------------------------------------- MV rate ------------------------------------- */

/* TLibCommon/TComRdCost.cpp:279-292  xGetExpGolombNumberOfBits: 1 + 2*floor(log2(v<=0 ? -2v+1 : 2v)) */
uint32_t hmo_eg_bits(int32_t v)
{
  uint32_t t = (v <= 0) ? (((uint32_t)(-v)) << 1) + 1u : ((uint32_t)v << 1);
  uint32_t len = 1;
  while (t != 1u) { t >>= 1; len += 2; }
  return len;
}

/* TLibCommon/TComRdCost.h:184-188  getBits(x, y) with m_mvPredictor, m_iCostScale */
uint32_t hmo_mv_bits(int x, int y, int pred_x, int pred_y, int scale)
{
  return hmo_eg_bits((int32_t)(x * (1 << scale)) - pred_x) + hmo_eg_bits((int32_t)(y * (1 << scale)) - pred_y);
}

/* TLibCommon/TComRdCost.h:172-183  getCost: (m_uiCost * bits) >> 16, all in 32-bit unsigned (wraps) */
uint32_t hmo_mv_cost(uint32_t lambda_cost, uint32_t bits)
{
  return (uint32_t)(lambda_cost * bits) >> 16;
}

/* ---------------------------------------------------------------- SAD / SSE / HAD ----------------------------- */

static int is_sized_sad_width(int w)
{
  return w == 4 || w == 8 || w == 12 || w == 16 || w == 24 || w == 32 || w == 48 || w == 64;
}

/* TLibCommon/TComRdCost.cpp:489-953 (xGetSAD4..64, 12/24/48): rows 0, 2^s, 2*2^s, ...; (sum << s) >> (bd-8).
 * TLibCommon/TComRdCost.cpp:461-487 (generic xGetSAD, other widths): every row, sub_shift ignored. */
uint32_t hmo_sad(const int16_t* org, int so, const int16_t* cur, int sc, int w, int h, int bit_depth, int sub_shift)
{
  uint32_t sum = 0;
  int step = is_sized_sad_width(w) ? (1 << sub_shift) : 1;
  int shl  = is_sized_sad_width(w) ? sub_shift : 0;
  for (int y = 0; y < h; y += step)
    for (int x = 0; x < w; x++)
      sum += (uint32_t)abs((int)org[y * so + x] - (int)cur[y * sc + x]);
  sum <<= shl;
  return sum >> (bit_depth - 8);
}

/* TLibCommon/TComRdCost.cpp:959-1304: per-sample ((d*d) >> 2(bd-8)) summed in uint32 */
uint32_t hmo_sse(const int16_t* org, int so, const int16_t* cur, int sc, int w, int h, int bit_depth)
{
  uint32_t sum = 0, sh = (uint32_t)((bit_depth - 8) << 1);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
    {
      int32_t d = (int32_t)org[y * so + x] - (int32_t)cur[y * sc + x];
      sum += (uint32_t)((d * d) >> sh);
    }
  return sum;
}

/* Un-normalised n x n Hadamard of the difference block, sum of absolute coefficients.  SURVEY.md App. A.10:
 * the reference's butterfly order (TComRdCost.cpp:1332-1523) only permutes / sign-flips coefficients. */
static uint32_t hadamard_abs_sum(const int16_t* org, int so, const int16_t* cur, int sc, int n)
{
  int32_t m[64];
  for (int y = 0; y < n; y++)
    for (int x = 0; x < n; x++)
      m[y * n + x] = (int32_t)org[y * so + x] - (int32_t)cur[y * sc + x];
  for (int len = 1; len < n; len <<= 1)                 /* rows */
    for (int y = 0; y < n; y++)
      for (int b = 0; b < n; b += 2 * len)
        for (int k = 0; k < len; k++)
        {
          int32_t a0 = m[y * n + b + k], a1 = m[y * n + b + k + len];
          m[y * n + b + k] = a0 + a1; m[y * n + b + k + len] = a0 - a1;
        }
  for (int len = 1; len < n; len <<= 1)                 /* columns */
    for (int x = 0; x < n; x++)
      for (int b = 0; b < n; b += 2 * len)
        for (int k = 0; k < len; k++)
        {
          int32_t a0 = m[(b + k) * n + x], a1 = m[(b + k + len) * n + x];
          m[(b + k) * n + x] = a0 + a1; m[(b + k + len) * n + x] = a0 - a1;
        }
  uint32_t s = 0;
  for (int i = 0; i < n * n; i++) s += (uint32_t)abs(m[i]);
  return s;
}

/* TLibCommon/TComRdCost.cpp:1526-1593 xGetHADs: 8x8 tiles iff both dims %8==0 ((s+2)>>2 per tile, :1520),
 * else 4x4 tiles iff %4 ((s+1)>>1, :1423), else 2x2 (no rounding, :1310-1330); total >> (bd-8). */
uint32_t hmo_had(const int16_t* org, int so, const int16_t* cur, int sc, int w, int h, int bit_depth)
{
  uint32_t sum = 0;
  int n = (w % 8 == 0 && h % 8 == 0) ? 8 : (w % 4 == 0 && h % 4 == 0) ? 4 : 2;
  for (int y = 0; y < h; y += n)
    for (int x = 0; x < w; x += n)
    {
      uint32_t s = hadamard_abs_sum(org + y * so + x, so, cur + y * sc + x, sc, n);
      sum += (n == 8) ? ((s + 2) >> 2) : (n == 4) ? ((s + 1) >> 1) : s;
    }
  return sum >> (bit_depth - 8);
}

/* kind: 0 SAD (integer ME), 1 SSE, 2 HAD, 3 SAD (sub-pel, DF_SADS*: same arithmetic as kind 0) */
uint32_t hmo_dist(int kind, const int16_t* org, int so, const int16_t* cur, int sc, int w, int h, int bit_depth, int sub_shift)
{
  if (kind == 1) return hmo_sse(org, so, cur, sc, w, h, bit_depth);
  if (kind == 2) return hmo_had(org, so, cur, sc, w, h, bit_depth);
  return hmo_sad(org, so, cur, sc, w, h, bit_depth, sub_shift);
}

/* n evaluations between two padded planes (test convenience: what hmb200_dist_batch computes on the device).
 * descs: n x {org_x, org_y, cur_x, cur_y, w, h, sub_shift}; org0 / cur0 point at sample (0, 0). */
void hmo_dist_batch(int kind, const int16_t* org0, int so, const int16_t* cur0, int sc, int bit_depth, int n,
                    const int32_t* descs, uint32_t* out)
{
  for (int i = 0; i < n; i++)
  {
    const int32_t* d = descs + 7 * i;
    out[i] = hmo_dist(kind, org0 + d[1] * so + d[0], so, cur0 + d[3] * sc + d[2], sc, d[4], d[5], bit_depth, d[6]);
  }
}

/* ---------------------------------------------------------------- search window -------------------------------- */

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* TLibCommon/TComDataCU.cpp:2788-2801 clipMv: quarter-pel MV clipped against the picture +- (8, maxCU+8) relative
 * to the CU origin (cu_x, cu_y). */
void hmo_clip_mv(int* mv_x, int* mv_y, int cu_x, int cu_y, int pic_w, int pic_h, int max_cu_w, int max_cu_h)
{
  int hmax = (pic_w + 8 - cu_x - 1) * 4, hmin = (-max_cu_w - 8 - cu_x + 1) * 4;
  int vmax = (pic_h + 8 - cu_y - 1) * 4, vmin = (-max_cu_h - 8 - cu_y + 1) * 4;
  *mv_x = clampi(*mv_x, hmin, hmax);
  *mv_y = clampi(*mv_y, vmin, vmax);
}

static int asr2(int v) { return (v >= 0) ? (v >> 2) : -((-v + 3) >> 2); }   /* floor(v / 4): arithmetic shift of a Short */

/* TLibEncoder/TEncSearch.cpp:3765-3781 xSetSearchRange: clip(pred) -+ (range << 2), clip both corners, >> 2 */
void hmo_search_range(int pred_x, int pred_y, int range, int cu_x, int cu_y, int pic_w, int pic_h, int max_cu_w, int max_cu_h,
                      int* lt_x, int* lt_y, int* rb_x, int* rb_y)
{
  int px = pred_x, py = pred_y;
  hmo_clip_mv(&px, &py, cu_x, cu_y, pic_w, pic_h, max_cu_w, max_cu_h);
  int lx = (int16_t)(px - (range << 2)), ly = (int16_t)(py - (range << 2));
  int rx = (int16_t)(px + (range << 2)), ry = (int16_t)(py + (range << 2));
  hmo_clip_mv(&lx, &ly, cu_x, cu_y, pic_w, pic_h, max_cu_w, max_cu_h);
  hmo_clip_mv(&rx, &ry, cu_x, cu_y, pic_w, pic_h, max_cu_w, max_cu_h);
  *lt_x = asr2(lx); *lt_y = asr2(ly); *rb_x = asr2(rx); *rb_y = asr2(ry);
}

/* ---------------------------------------------------------------- integer full search -------------------------- */

/* TLibEncoder/TEncSearch.cpp:3786-3843 xPatternSearch.  ref_at_pu = reference sample co-located with the PU's
 * top-left.  fen != 0 reproduces getUseFastEnc(): iSubShift = 1 when the PU has more than 8 rows (:3804-3810).
 * Scan y-outer / x-inner, strict '<' (first wins); cost scale 2 (:3722); returned SAD excludes the MV cost (:3841). */
void hmo_pattern_search(const int16_t* org, int so, int w, int h, int bit_depth,
                        const int16_t* ref_at_pu, int sr, int lt_x, int lt_y, int rb_x, int rb_y,
                        uint32_t lambda_cost, int pred_x, int pred_y, int fen,
                        int* mv_x, int* mv_y, uint32_t* sad_out)
{
  uint32_t best = 0xFFFFFFFFu; int bx = 0, by = 0;
  int sub = (fen && h > 8) ? 1 : 0;
  for (int y = lt_y; y <= rb_y; y++)
    for (int x = lt_x; x <= rb_x; x++)
    {
      uint32_t c = hmo_sad(org, so, ref_at_pu + y * sr + x, sr, w, h, bit_depth, sub)
                 + hmo_mv_cost(lambda_cost, hmo_mv_bits(x, y, pred_x, pred_y, 2));
      if (c < best) { best = c; bx = x; by = y; }
    }
  *mv_x = bx; *mv_y = by;
  *sad_out = best - hmo_mv_cost(lambda_cost, hmo_mv_bits(bx, by, pred_x, pred_y, 2));
}

/* ---------------------------------------------------------------- TZ fast search ------------------------------- */

/* TLibEncoder/TEncSearch.cpp:3881-4083 xTZSearch (FastSearch = 1) with its helpers xTZSearchHelp (:332-436, the
 * non-selective branch: SAD with FEN sub-sampling + getCost, strict '<'), xTZ8PointDiamondSearch (:626-800) and
 * xTZ2PointSearch (:438-567), under TZ_SEARCH_CONFIGURATION (:297-313): raster step 5, zero vector tested, diamond first
 * search over distances 1, 2, 4 .. <= search range without early stop (FastMEAssumingSmootherMV off), raster search
 * when the best distance exceeds 5, star refinement until the best distance is 0.
 *
 * Restated as "evaluate an ordered list of points, keep the first minimum": every stage of the reference issues
 * xTZSearchHelp calls whose positions depend only on the stage's start point, so the sequential strict-'<' updates
 * equal taking the first minimal cost in call order.  uiBestRound (diamond calls since the last improvement) is only
 * read by the first search's stop criterion, bFirstSearchStop = FastMEAssumingSmootherMVEnabled (on by default in
 * TAppEncCfg.cpp:808): the first search ends after uiFirstSearchRounds = 3 diamonds in a row without improvement. */
typedef struct { uint32_t cost; int x, y, dist, nr; } tz_state;
typedef struct { int x, y, nr, dist; } tz_point;

typedef struct {
  const int16_t* org; int so; const int16_t* ref; int sr; int w, h, bit_depth, sub;
  uint32_t lambda_cost; int pred_x, pred_y;
} tz_ctx;

static int tz_try(const tz_ctx* c, tz_state* st, int x, int y, int nr, int dist)
{
  uint32_t cost = hmo_sad(c->org, c->so, c->ref + y * c->sr + x, c->sr, c->w, c->h, c->bit_depth, c->sub)
                + hmo_mv_cost(c->lambda_cost, hmo_mv_bits(x, y, c->pred_x, c->pred_y, 2));
  if (cost < st->cost) { st->cost = cost; st->x = x; st->y = y; st->dist = dist; st->nr = nr; return 1; }
  return 0;
}

/* the points one diamond call visits, in call order; L/T/R/B = search range (inclusive).  Returns the count. */
static int tz_diamond(int sx, int sy, int d, int L, int T, int R, int B, tz_point* out)
{
  int n = 0;
  const int top = sy - d, bot = sy + d, left = sx - d, right = sx + d;
#define TZ_ADD(cond, px, py, pnr, pd) do { if (cond) { out[n].x = (px); out[n].y = (py); out[n].nr = (pnr); out[n].dist = (pd); n++; } } while (0)
  if (d == 1)
  {
    TZ_ADD(top >= T, sx, top, 2, d); TZ_ADD(left >= L, left, sy, 4, d); TZ_ADD(right <= R, right, sy, 5, d); TZ_ADD(bot <= B, sx, bot, 7, d);
  }
  else if (d <= 8)
  {
    const int h = d >> 1, t2 = sy - h, b2 = sy + h, l2 = sx - h, r2 = sx + h;
    /* inside the range every condition below holds, which is the reference's unconditional branch (same order) */
    TZ_ADD(top >= T, sx, top, 2, d);
    TZ_ADD(t2 >= T && l2 >= L, l2, t2, 1, h);
    TZ_ADD(t2 >= T && r2 <= R, r2, t2, 3, h);
    TZ_ADD(left >= L, left, sy, 4, d);
    TZ_ADD(right <= R, right, sy, 5, d);
    TZ_ADD(b2 <= B && l2 >= L, l2, b2, 6, h);
    TZ_ADD(b2 <= B && r2 <= R, r2, b2, 8, h);
    TZ_ADD(bot <= B, sx, bot, 7, d);
  }
  else
  {
    TZ_ADD(top >= T, sx, top, 0, d); TZ_ADD(left >= L, left, sy, 0, d); TZ_ADD(right <= R, right, sy, 0, d); TZ_ADD(bot <= B, sx, bot, 0, d);
    for (int i = 1; i < 4; i++)
    {
      const int q = (d >> 2) * i, yt = top + q, yb = bot - q, xl = sx - q, xr = sx + q;
      TZ_ADD(yt >= T && xl >= L, xl, yt, 0, d); TZ_ADD(yt >= T && xr <= R, xr, yt, 0, d);
      TZ_ADD(yb <= B && xl >= L, xl, yb, 0, d); TZ_ADD(yb <= B && xr <= R, xr, yb, 0, d);
    }
  }
#undef TZ_ADD
  return n;
}

/* xTZ2PointSearch: the two untested neighbours of the best point, selected by the diamond position it came from */
static void tz_two_point(const tz_ctx* c, tz_state* st, int L, int T, int R, int B)
{
  const int x = st->x, y = st->y;
  switch (st->nr)
  {
    case 1: if (x - 1 >= L) tz_try(c, st, x - 1, y, 0, 2);     if (y - 1 >= T) tz_try(c, st, x, y - 1, 0, 2); break;
    case 2: if (y - 1 >= T) { if (x - 1 >= L) tz_try(c, st, x - 1, y - 1, 0, 2); if (x + 1 <= R) tz_try(c, st, x + 1, y - 1, 0, 2); } break;
    case 3: if (y - 1 >= T) tz_try(c, st, x, y - 1, 0, 2);     if (x + 1 <= R) tz_try(c, st, x + 1, y, 0, 2); break;
    case 4: if (x - 1 >= L) { if (y + 1 <= B) tz_try(c, st, x - 1, y + 1, 0, 2); if (y - 1 >= T) tz_try(c, st, x - 1, y - 1, 0, 2); } break;
    case 5: if (x + 1 <= R) { if (y - 1 >= T) tz_try(c, st, x + 1, y - 1, 0, 2); if (y + 1 <= B) tz_try(c, st, x + 1, y + 1, 0, 2); } break;
    case 6: if (x - 1 >= L) tz_try(c, st, x - 1, y, 0, 2);     if (y + 1 <= B) tz_try(c, st, x, y + 1, 0, 2); break;
    case 7: if (y + 1 <= B) { if (x - 1 >= L) tz_try(c, st, x - 1, y + 1, 0, 2); if (x + 1 <= R) tz_try(c, st, x + 1, y + 1, 0, 2); } break;
    case 8: if (x + 1 <= R) tz_try(c, st, x + 1, y, 0, 2);     if (y + 1 <= B) tz_try(c, st, x, y + 1, 0, 2); break;
    default: break;   /* the reference asserts; the callers below never get here with nr == 0 */
  }
}

/* (pred_x, pred_y): the AMVP predictor in quarter pel (rcMv on entry and m_mvPredictor); (lt, rb): the window of
 * xSetSearchRange(pred); the CU geometry feeds clipMv (TLibCommon/TComDataCU.cpp:2788-2801); has_imv: pIntegerMv2Nx2NPred. */
void hmo_tz_search(const int16_t* org, int so, int w, int h, int bit_depth, const int16_t* ref_at_pu, int sr,
                   int lt_x, int lt_y, int rb_x, int rb_y, uint32_t lambda_cost, int pred_x, int pred_y, int fen,
                   int cu_x, int cu_y, int pic_w, int pic_h, int max_cu, int search_range, int first_search_stop,
                   int has_imv, int imv_x, int imv_y, int* mv_x, int* mv_y, uint32_t* sad_out)
{
  tz_ctx c = { org, so, ref_at_pu, sr, w, h, bit_depth, (fen && h > 8) ? 1 : 0, lambda_cost, pred_x, pred_y };
  tz_state st = { 0xFFFFFFFFu, 0, 0, 0, 0 };
  tz_point pts[16];
  int rl = lt_x, rt = lt_y, rr = rb_x, rb = rb_y;             /* range of the raster search (reset below) */

  int sx = pred_x, sy = pred_y;
  hmo_clip_mv(&sx, &sy, cu_x, cu_y, pic_w, pic_h, max_cu, max_cu);
  tz_try(&c, &st, asr2(sx), asr2(sy), 0, 0);                  /* :3908 predictor, clipped, integer pel */
  tz_try(&c, &st, 0, 0, 0, 0);                                /* :3924 zero vector */
  if (has_imv)
  {
    int ix = (int16_t)(imv_x << 2), iy = (int16_t)(imv_y << 2);
    hmo_clip_mv(&ix, &iy, cu_x, cu_y, pic_w, pic_h, max_cu, max_cu);
    tz_try(&c, &st, asr2(ix), asr2(iy), 0, 0);
    /* :3935-3946 only the raster search sees the re-centred range; the diamonds keep the caller's range */
    hmo_search_range((int16_t)(st.x << 2), (int16_t)(st.y << 2), search_range, cu_x, cu_y, pic_w, pic_h, max_cu, max_cu, &rl, &rt, &rr, &rb);
  }

  const int start_x = st.x, start_y = st.y;
  int rounds = 0;                                             /* uiBestRound: 0 after the start candidates (:3908 always improves) */
  for (int d = 1; d <= search_range; d *= 2)                  /* first search: every diamond around the same start */
  {
    int n = tz_diamond(start_x, start_y, d, lt_x, lt_y, rb_x, rb_y, pts);
    rounds++;
    for (int i = 0; i < n; i++) if (tz_try(&c, &st, pts[i].x, pts[i].y, pts[i].nr, pts[i].dist)) rounds = 0;
    if (first_search_stop && rounds >= 3) break;              /* :3962 bFirstSearchStop, uiFirstSearchRounds = 3 */
  }
  if (st.dist == 1) { st.dist = 0; tz_two_point(&c, &st, lt_x, lt_y, rb_x, rb_y); }
  if (st.dist > 5)                                            /* raster search, step iRaster = 5 */
  {
    st.dist = 5;
    for (int y = rt; y <= rb; y += 5)
      for (int x = rl; x <= rr; x += 5) tz_try(&c, &st, x, y, 0, 5);
  }
  while (st.dist > 0)                                         /* star refinement */
  {
    const int bx = st.x, by = st.y;
    st.dist = 0; st.nr = 0;
    for (int d = 1; d < search_range + 1; d *= 2)
    {
      int n = tz_diamond(bx, by, d, lt_x, lt_y, rb_x, rb_y, pts);
      for (int i = 0; i < n; i++) tz_try(&c, &st, pts[i].x, pts[i].y, pts[i].nr, pts[i].dist);
    }
    if (st.dist == 1) { st.dist = 0; if (st.nr != 0) tz_two_point(&c, &st, lt_x, lt_y, rb_x, rb_y); }
  }
  *mv_x = st.x; *mv_y = st.y;
  *sad_out = st.cost - hmo_mv_cost(lambda_cost, hmo_mv_bits(st.x, st.y, pred_x, pred_y, 2));
}

/* ---------------------------------------------------------------- interpolation -------------------------------- */

/* TLibCommon/TComInterpolationFilter.cpp:57-63 m_lumaFilter */
static const int k_luma_taps[4][8] = {
  {  0, 0,   0, 64,  0,   0, 0,  0 },
  { -1, 4, -10, 58, 17,  -5, 1,  0 },
  { -1, 4, -11, 40, 40, -11, 4, -1 },
  {  0, 1,  -5, 17, 58, -10, 4, -1 } };

/* One sample of the separable luma interpolation at integer position p (pointer into a padded plane) and
 * fractional phase (fx, fy) in quarter samples.  Horizontal pass first into the 14-bit intermediate domain, then
 * vertical pass back to pixels, exactly as the reference chains filterHor(isLast=false) and
 * filterVer(isFirst=false, isLast=true) for every plane it builds, including phase 0
 * (TLibCommon/TComInterpolationFilter.cpp:94-154 filterCopy, :172-257 filter<>, headRoom = max(2, 14 - bd),
 *  IF_INTERNAL_OFFS = 8192, TLibCommon/TComInterpolationFilter.h:47-51). */
static int16_t interp_sample(const int16_t* p, int stride, int fx, int fy, int bit_depth)
{
  int head = 14 - bit_depth; if (head < 2) head = 2;
  int16_t col[8];
  int r0 = (fy == 0) ? 0 : -3, r1 = (fy == 0) ? 0 : 4;
  for (int r = r0; r <= r1; r++)
  {
    const int16_t* q = p + r * stride;
    int16_t v;
    if (fx == 0)
      v = (int16_t)((int16_t)(q[0] << head) - 8192);                       /* filterCopy, isFirst && !isLast */
    else
    {
      int sum = 0;
      for (int t = 0; t < 8; t++) sum += q[t - 3] * k_luma_taps[fx][t];
      int shift = 6 - head;
      v = (int16_t)((sum + (-8192 * (1 << shift))) >> shift);              /* filter<8,false,true,false> */
    }
    col[r + 3] = v;
  }
  int maxv = (1 << bit_depth) - 1;
  int val;
  if (fy == 0)
    val = (int16_t)((col[3] + 8192 + (1 << (head - 1))) >> head);          /* filterCopy, !isFirst && isLast */
  else
  {
    int sum = 0;
    for (int t = 0; t < 8; t++) sum += col[t] * k_luma_taps[fy][t];
    int shift = 6 + head;
    val = (int16_t)((sum + (1 << (shift - 1)) + (8192 << 6)) >> shift);    /* filter<8,true,false,true> */
  }
  if (val < 0) val = 0;
  if (val > maxv) val = maxv;
  return (int16_t)val;
}

/* Interpolated W x H block at quarter-pel displacement (qx, qy) from `base` (integer-MV-compensated block origin). */
static void interp_block(const int16_t* base, int stride, int qx, int qy, int w, int h, int bit_depth, int16_t* dst /* w*h */)
{
  int ix = asr2(qx), iy = asr2(qy), fx = qx & 3, fy = qy & 3;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      dst[y * w + x] = interp_sample(base + (y + iy) * stride + x + ix, stride, fx, fy, bit_depth);
}

/* TLibEncoder/TEncSearch.cpp:51-75 candidate orders */
static const int k_refine_h[9][2] = { {0,0},{0,-1},{0,1},{-1,0},{1,0},{-1,-1},{1,-1},{-1,1},{1,1} };
static const int k_refine_q[9][2] = { {0,0},{0,-1},{0,1},{-1,-1},{1,-1},{-1,0},{1,0},{-1,1},{1,1} };

/* TLibEncoder/TEncSearch.cpp:4240-4276 xPatternSearchFracDIF with :808-861 xPatternRefinement and the planes of
 * :5338-5539 xExtDIFUpSamplingH/Q.  Each of the reference's pre-filtered planes, at the offset xPatternRefinement
 * applies to it (:830-840), is the block interpolated at quarter-pel displacement 2*half (stage 1) or
 * 2*half + qter (stage 2) from the integer MV; here each candidate block is interpolated directly.
 * Cost: distortion + getCost at scale 1 (half, :3746) / scale 0 (quarter, :4267) of (candidate + accumulated MV);
 * strict '<', first in table order wins.  Distortion is HAD when use_had (HadamardME && !lossless), else SAD. */
void hmo_frac_search(const int16_t* org, int so, int w, int h, int bit_depth,
                     const int16_t* ref_at_pu, int sr, int mv_x, int mv_y,
                     uint32_t lambda_cost, int pred_x, int pred_y, int use_had,
                     int* half_x, int* half_y, int* qter_x, int* qter_y, uint32_t* cost_out)
{
  const int16_t* base = ref_at_pu + mv_y * sr + mv_x;
  int16_t* blk = (int16_t*)malloc(sizeof(int16_t) * (size_t)w * (size_t)h);
  uint32_t best = 0xFFFFFFFFu; int bi = 0;
  for (int i = 0; i < 9; i++)
  {
    int cx = k_refine_h[i][0], cy = k_refine_h[i][1];
    interp_block(base, sr, 2 * cx, 2 * cy, w, h, bit_depth, blk);
    uint32_t d = use_had ? hmo_had(org, so, blk, w, w, h, bit_depth) : hmo_sad(org, so, blk, w, w, h, bit_depth, 0);
    d += hmo_mv_cost(lambda_cost, hmo_mv_bits(cx + 2 * mv_x, cy + 2 * mv_y, pred_x, pred_y, 1));
    if (d < best) { best = d; bi = i; }
  }
  int hx = k_refine_h[bi][0], hy = k_refine_h[bi][1];
  best = 0xFFFFFFFFu; bi = 0;
  for (int i = 0; i < 9; i++)
  {
    int cx = k_refine_q[i][0], cy = k_refine_q[i][1];
    interp_block(base, sr, 2 * hx + cx, 2 * hy + cy, w, h, bit_depth, blk);
    uint32_t d = use_had ? hmo_had(org, so, blk, w, w, h, bit_depth) : hmo_sad(org, so, blk, w, w, h, bit_depth, 0);
    d += hmo_mv_cost(lambda_cost, hmo_mv_bits(cx + 2 * (2 * mv_x + hx), cy + 2 * (2 * mv_y + hy), pred_x, pred_y, 0));
    if (d < best) { best = d; bi = i; }
  }
  free(blk);
  *half_x = hx; *half_y = hy; *qter_x = k_refine_q[bi][0]; *qter_y = k_refine_q[bi][1]; *cost_out = best;
}

/* ---------------------------------------------------------------- planes --------------------------------------- */

/* TLibCommon/TComPicYuv.cpp:197-242 extendPicBorder: replicate edge samples into the margins.
 * plane0 points at sample (0,0); the buffer holds margin_x / margin_y samples on every side. */
void hmo_extend_border(int16_t* plane0, int stride, int w, int h, int margin_x, int margin_y)
{
  for (int y = 0; y < h; y++)
  {
    int16_t* row = plane0 + y * stride;
    for (int x = 1; x <= margin_x; x++) { row[-x] = row[0]; row[w - 1 + x] = row[w - 1]; }
  }
  for (int y = 1; y <= margin_y; y++)
  {
    memcpy(plane0 - y * stride - margin_x, plane0 - margin_x, sizeof(int16_t) * (size_t)(w + 2 * margin_x));
    memcpy(plane0 + (h - 1 + y) * stride - margin_x, plane0 + (h - 1) * stride - margin_x, sizeof(int16_t) * (size_t)(w + 2 * margin_x));
  }
}

/* TLibVideoIO/TVideoIOYuv.cpp:247-377 (readPlane, luma component) followed by :70-99 (scalePlane) as driven by
 * TVideoIOYuv::read (:680-741): `file` holds width x height luma samples of one frame of a planar file (bytes, or
 * 16-bit little endian when is16); dst receives (width + pad_x) x (height + pad_y) Pels: right / lower padding
 * replicate the last column / row, then every sample is scaled by 2^shift (shift = internal - file bit depth;
 * negative: round, shift, clip to [0, 2^internal - 1]). */
void hmo_read_luma(const uint8_t* file, int is16, int width, int height, int pad_x, int pad_y, int shift,
                   int internal_bit_depth, int16_t* dst, int dst_stride)
{
  const int fw = width + pad_x, fh = height + pad_y;
  for (int y = 0; y < height; y++)
  {
    int16_t* row = dst + y * dst_stride;
    const uint8_t* src = file + (size_t)y * width * (is16 ? 2 : 1);
    for (int x = 0; x < width; x++)
      row[x] = is16 ? (int16_t)((uint16_t)src[2 * x] | ((uint16_t)src[2 * x + 1] << 8)) : (int16_t)src[x];
    for (int x = width; x < fw; x++) row[x] = row[width - 1];
  }
  for (int y = height; y < fh; y++) memcpy(dst + y * dst_stride, dst + (y - 1) * dst_stride, sizeof(int16_t) * (size_t)fw);
  if (shift > 0)
  {
    for (int y = 0; y < fh; y++) for (int x = 0; x < fw; x++) dst[y * dst_stride + x] = (int16_t)(dst[y * dst_stride + x] << shift);
  }
  else if (shift < 0)
  {
    const int s = -shift, rounding = 1 << (s - 1), maxval = (1 << internal_bit_depth) - 1;
    for (int y = 0; y < fh; y++)
      for (int x = 0; x < fw; x++)
      {
        int v = (int16_t)((int16_t)(dst[y * dst_stride + x] + rounding) >> s);      /* Pel arithmetic, like the reference */
        dst[y * dst_stride + x] = (int16_t)(v < 0 ? 0 : (v > maxval ? maxval : v));
      }
  }
}

/* ---------------------------------------------------------------- motion compensation + distortion ------------- */

/* TLibCommon/TComPrediction.cpp:668-706 xPredInterBlk for COMPONENT_Y with bi == false: the prediction block at
 * quarter-pel MV (mv_x, mv_y).  Three branches: yFrac == 0 -> one horizontal pass straight to pixels (filterHor with
 * isLast = true: shift 6, offset 32, clip; xFrac == 0 is a copy, TComInterpolationFilter.cpp:94-154); xFrac == 0 -> one
 * vertical pass; else horizontal into the 14-bit domain (rows -3..+4) then vertical (:172-257).  ref_at_pu points at the
 * reference sample co-located with the PU's top-left. */
static int16_t mc_clip(int v, int maxv) { return (int16_t)(v < 0 ? 0 : (v > maxv ? maxv : v)); }

void hmo_mc_block(const int16_t* ref_at_pu, int sr, int w, int h, int mv_x, int mv_y, int bit_depth, int16_t* dst, int sd)
{
  const int16_t* ref = ref_at_pu + asr2(mv_x) + asr2(mv_y) * sr;
  const int fx = mv_x & 3, fy = mv_y & 3, maxv = (1 << bit_depth) - 1;
  int head = 14 - bit_depth; if (head < 2) head = 2;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
    {
      const int16_t* p = ref + y * sr + x;
      if (fy == 0)
      {
        if (fx == 0) { dst[y * sd + x] = p[0]; continue; }
        int sum = 0;
        for (int t = 0; t < 8; t++) sum += k_luma_taps[fx][t] * p[t - 3];
        dst[y * sd + x] = mc_clip((int16_t)((sum + 32) >> 6), maxv);
      }
      else if (fx == 0)
      {
        int sum = 0;
        for (int t = 0; t < 8; t++) sum += k_luma_taps[fy][t] * p[(t - 3) * sr];
        dst[y * sd + x] = mc_clip((int16_t)((sum + 32) >> 6), maxv);
      }
      else
      {
        const int sh1 = 6 - head, sh2 = 6 + head;
        int col[8];
        for (int r = 0; r < 8; r++)
        {
          int sum = 0;
          for (int t = 0; t < 8; t++) sum += k_luma_taps[fx][t] * p[(r - 3) * sr + t - 3];
          col[r] = (int16_t)((sum - (8192 << sh1)) >> sh1);
        }
        int sum = 0;
        for (int t = 0; t < 8; t++) sum += k_luma_taps[fy][t] * col[t];
        dst[y * sd + x] = mc_clip((int16_t)((sum + (1 << (sh2 - 1)) + (8192 << 6)) >> sh2), maxv);
      }
    }
}

/* Distortion of that prediction against the original: kind 0 = SAD as xGetTemplateCost computes it through
 * getDistPart(DF_SAD) (TLibEncoder/TEncSearch.cpp:3619-3658), kind 2 = HADs as xGetInterPredictionError does
 * (:2809-2830, setDistParam with bHadamard). */
uint32_t hmo_mc_dist(int kind, const int16_t* org, int so, const int16_t* ref_at_pu, int sr, int w, int h, int mv_x, int mv_y, int bit_depth)
{
  int16_t pred[64 * 64];
  hmo_mc_block(ref_at_pu, sr, w, h, mv_x, mv_y, bit_depth, pred, 64);
  return hmo_dist(kind, org, so, pred, 64, w, h, bit_depth, 0);
}

/* xPredInterBlk with bi == true (TLibCommon/TComPrediction.cpp:668-706: every filter call gets isLast = !bi = false): the
 * block stays in the 14-bit intermediate domain.  yFrac == 0: one horizontal pass, isFirst (shift 6 - headroom, offset
 * -8192 << shift; xFrac == 0 is filterCopy's (s << headroom) - 8192, TComInterpolationFilter.cpp:113-124); xFrac == 0: one
 * vertical pass with the same shift and offset; else horizontal isFirst, then vertical with isFirst = isLast = false:
 * shift 6, no offset, no clip (:196-251). */
void hmo_mc_block_bi(const int16_t* ref_at_pu, int sr, int w, int h, int mv_x, int mv_y, int bit_depth, int16_t* dst, int sd)
{
  const int16_t* ref = ref_at_pu + asr2(mv_x) + asr2(mv_y) * sr;
  const int fx = mv_x & 3, fy = mv_y & 3;
  int head = 14 - bit_depth; if (head < 2) head = 2;
  const int sh1 = 6 - head;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
    {
      const int16_t* p = ref + y * sr + x;
      if (fy == 0)
      {
        if (fx == 0) { dst[y * sd + x] = (int16_t)((p[0] << head) - 8192); continue; }
        int sum = 0;
        for (int t = 0; t < 8; t++) sum += k_luma_taps[fx][t] * p[t - 3];
        dst[y * sd + x] = (int16_t)((sum - (8192 << sh1)) >> sh1);
      }
      else if (fx == 0)
      {
        int sum = 0;
        for (int t = 0; t < 8; t++) sum += k_luma_taps[fy][t] * p[(t - 3) * sr];
        dst[y * sd + x] = (int16_t)((sum - (8192 << sh1)) >> sh1);
      }
      else
      {
        int col[8];
        for (int r = 0; r < 8; r++)
        {
          int sum = 0;
          for (int t = 0; t < 8; t++) sum += k_luma_taps[fx][t] * p[(r - 3) * sr + t - 3];
          col[r] = (int16_t)((sum - (8192 << sh1)) >> sh1);
        }
        int sum = 0;
        for (int t = 0; t < 8; t++) sum += k_luma_taps[fy][t] * col[t];
        dst[y * sd + x] = (int16_t)(sum >> 6);
      }
    }
}

/* TComYuv::addAvg (TLibCommon/TComYuv.cpp:352-407) through TComPrediction::xWeightedAverage (TComPrediction.cpp:708-724):
 * clip((a + b + offset) >> shift), shift = max(2, 14 - bitDepth) + 1, offset = (1 << (shift - 1)) + 2 * 8192. */
void hmo_add_avg(const int16_t* a, int sa, const int16_t* b, int sb, int w, int h, int bit_depth, int16_t* dst, int sd)
{
  int head = 14 - bit_depth; if (head < 2) head = 2;
  const int shift = head + 1, offset = (1 << (shift - 1)) + 2 * 8192, maxv = (1 << bit_depth) - 1;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      dst[y * sd + x] = mc_clip((a[y * sa + x] + b[y * sb + x] + offset) >> shift, maxv);
}

/* Prediction error of one merge / AMVP candidate (TEncSearch::xGetInterPredictionError, TLibEncoder/TEncSearch.cpp:2805-2826,
 * after TComPrediction::motionCompensation, TComPrediction.cpp:539-586): inter_dir 1 / 2 = uni-directional from list 0 / 1,
 * 3 = bi-prediction average - unless both lists point at the same picture with the same MV (xCheckIdenticalMotion, :496-517),
 * which the caller signals with same_picture and which predicts from list 0 alone. */
uint32_t hmo_mc_cand_dist(int kind, const int16_t* org, int so, int w, int h, int bit_depth, int inter_dir,
                          const int16_t* ref0_at_pu, int sr0, int mv0_x, int mv0_y,
                          const int16_t* ref1_at_pu, int sr1, int mv1_x, int mv1_y, int same_picture)
{
  int16_t pred[64 * 64];
  if (inter_dir == 3 && same_picture && mv0_x == mv1_x && mv0_y == mv1_y) inter_dir = 1;
  if (inter_dir == 3)
  {
    int16_t a[64 * 64], b[64 * 64];
    hmo_mc_block_bi(ref0_at_pu, sr0, w, h, mv0_x, mv0_y, bit_depth, a, 64);
    hmo_mc_block_bi(ref1_at_pu, sr1, w, h, mv1_x, mv1_y, bit_depth, b, 64);
    hmo_add_avg(a, 64, b, 64, w, h, bit_depth, pred, 64);
  }
  else if (inter_dir == 1) hmo_mc_block(ref0_at_pu, sr0, w, h, mv0_x, mv0_y, bit_depth, pred, 64);
  else                     hmo_mc_block(ref1_at_pu, sr1, w, h, mv1_x, mv1_y, bit_depth, pred, 64);
  return hmo_dist(kind, org, so, pred, 64, w, h, bit_depth, 0);
}

/* The candidate loop of TEncSearch::xMergeEstimation (TLibEncoder/TEncSearch.cpp:2868-2892): cost = error + getCost(bits)
 * = error + ((m_uiCost * bits) >> 16) in UInt arithmetic (TLibCommon/TComRdCost.h:177), strict '<' in candidate order.
 * dist / bits: n candidates of one PU.  Returns the merge index, *cost = its cost. */
int hmo_merge_pick(const uint32_t* dist, const uint32_t* bits, int n, uint32_t lambda_cost, uint32_t* cost)
{
  uint32_t best = 0xffffffffu; int bi = 0;
  for (int i = 0; i < n; i++)
  {
    const uint32_t c = dist[i] + ((uint32_t)(lambda_cost * bits[i]) >> 16);
    if (c < best) { best = c; bi = i; }
  }
  *cost = best;
  return bi;
}

/* The candidate loop of TEncSearch::xEstimateMvPredAMVP (TLibEncoder/TEncSearch.cpp:3457-3469) over xGetTemplateCost
 * (:3619-3658): cost = (UInt) calcRdCost(bits, SAD, false, DF_SAD) = floor(SAD + floor(bits * lambda + 0.5) / 65536) with
 * lambda = m_uiLambdaMotionSAD[0] as a double (TLibCommon/TComRdCost.cpp:67-73, 99-108) - an integer product, so
 * SAD + ((bits * lambda) >> 16) in 64 bits; `uiBestCost > uiTmpCost` keeps the first minimum. */
int hmo_amvp_pick(const uint32_t* sad, const uint32_t* bits, int n, uint32_t lambda_motion_sad, uint32_t* cost)
{
  uint32_t best = 0xffffffffu; int bi = 0;
  for (int i = 0; i < n; i++)
  {
    const uint32_t c = (uint32_t)((uint64_t)sad[i] + (((uint64_t)bits[i] * (uint64_t)lambda_motion_sad) >> 16));
    if (best > c) { best = c; bi = i; }
  }
  *cost = best;
  return bi;
}

/* ---------------------------------------------------------------- job lists ------------------------------------ */

/* Executes a job list the way TEncSearch::xMotionEstimation (TLibEncoder/TEncSearch.cpp:3663-3760) drives the two
 * searches: integer search at cost scale 2, then quarter-pel refinement.  cur0 / ref0 point at luma sample (0,0) of
 * the current (original) and padded reference planes.  Returns CPU seconds (clock()). */
double hmo_run_jobs(const int16_t* cur0, int cur_stride, const int16_t* ref0, int ref_stride, int bit_depth,
                    const hmo_job* jobs, int njobs, int fen, int use_had, int do_frac, hmo_result* out)
{
  clock_t t0 = clock();
  for (int i = 0; i < njobs; i++)
  {
    const hmo_job* j = &jobs[i];
    const int16_t* org = cur0 + j->pu_y * cur_stride + j->pu_x;
    const int16_t* ref = ref0 + j->pu_y * ref_stride + j->pu_x;
    hmo_result r; memset(&r, 0, sizeof(r));
    int mx, my;
    hmo_pattern_search(org, cur_stride, j->w, j->h, bit_depth, ref, ref_stride, j->lt_x, j->lt_y, j->rb_x, j->rb_y,
                       j->lambda_cost, j->pred_x, j->pred_y, fen, &mx, &my, &r.sad);
    r.mv_x = mx; r.mv_y = my;
    if (do_frac)
    {
      int hx, hy, qx, qy;
      hmo_frac_search(org, cur_stride, j->w, j->h, bit_depth, ref, ref_stride, mx, my, j->lambda_cost, j->pred_x, j->pred_y,
                      use_had, &hx, &hy, &qx, &qy, &r.frac_cost);
      r.half_x = hx; r.half_y = hy; r.qter_x = qx; r.qter_y = qy;
    }
    out[i] = r;
  }
  return (double)(clock() - t0) / CLOCKS_PER_SEC;
}

/* ---------------------------------------------------------------- intra first pass ----------------------------- */
/* The 35 luma intra predictions of one block and their Hadamard distortions: the loop of TEncSearch::estIntraPredQT
 * (TLibEncoder/TEncSearch.cpp:2270-2296: predIntraAng + distParam.DistFunc per mode).  Reference samples arrive the way
 * the encoder has them after initIntraPatternChType: two (2N+1) x (2N+1)-strided buffers (unfiltered / smoothed), of
 * which only row 0 (corner, above, above-right) and column 0 (left, below-left) are defined
 * (TLibCommon/TComPattern.cpp:115-175; TComPrediction.cpp:412-494 reads them at ptrSrc + sw + 1).  Here they are two
 * lines instead: top[0..2N] = corner, above, above-right and left[0..2N] = corner, left, below-left. */

static int clip_pel(int v, int bit_depth) { int m = (1 << bit_depth) - 1; return v < 0 ? 0 : (v > m ? m : v); }
static int log2i(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

/* TLibCommon/TComPattern.cpp:544-568 filteringIntraReferenceSamples (luma; thresholds TComPrediction.cpp:50-57) */
int hmo_intra_use_filtered(int mode, int n)
{
  static const int thr[5] = {10, 7, 1, 0, 10};               /* 4, 8, 16, 32, 64 */
  if (mode == 1) return 0;                                    /* DC: never */
  int d1 = abs(mode - 10), d2 = abs(mode - 26);               /* HOR_IDX, VER_IDX; planar (0): min(10, 26) = 10 */
  return (d1 < d2 ? d1 : d2) > thr[log2i(n) - 2];
}

/* one prediction: TComPrediction.cpp:412-494 (predIntraAng, luma, no lossless DPCM), :756-816 (planar), :183-222 + :819-848
 * (DC and its edge filter), :250-410 (angular, edge filter of the pure vertical / horizontal modes).
 * above_ok / left_ok: bAbove / bLeft of predIntraAng (always true after initIntraPatternChType, :155-156). */
void hmo_intra_predict(int mode, const int16_t* top, const int16_t* left, int n, int bit_depth, int above_ok, int left_ok,
                       int edge_filters, int16_t* dst, int sd)
{
  if (mode == 0)                                              /* planar */
  {
    int sh = log2i(n);
    for (int y = 0; y < n; y++)
      for (int x = 0; x < n; x++)
      {
        int hor = (n - 1 - x) * left[1 + y] + (x + 1) * top[1 + n];
        int ver = (n - 1 - y) * top[1 + x] + (y + 1) * left[1 + n];
        dst[y * sd + x] = (int16_t)((hor + ver + n) >> (sh + 1));
      }
    return;
  }
  if (mode == 1)                                              /* DC */
  {
    int sum = 0, dc;
    if (above_ok) for (int i = 0; i < n; i++) sum += top[1 + i];
    if (left_ok)  for (int i = 0; i < n; i++) sum += left[1 + i];
    if (above_ok && left_ok) dc = (sum + n) / (2 * n);
    else if (above_ok || left_ok) dc = (sum + n / 2) / n;
    else dc = left[1];                                        /* pSrc[-1] (:218) */
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) dst[y * sd + x] = (int16_t)dc;
    if (above_ok && left_ok && n <= 16)                       /* xDCPredFiltering, luma, blocks up to 16x16 */
    {
      dst[0] = (int16_t)((top[1] + left[1] + 2 * dc + 2) >> 2);
      for (int x = 1; x < n; x++) dst[x] = (int16_t)((top[1 + x] + 3 * dc + 2) >> 2);
      for (int y = 1; y < n; y++) dst[y * sd] = (int16_t)((left[1 + y] + 3 * dc + 2) >> 2);
    }
    return;
  }
  static const int ang_tab[9] = {0, 2, 5, 9, 13, 17, 21, 26, 32};
  static const int inv_tab[9] = {0, 4096, 1638, 910, 630, 482, 390, 315, 256};
  const int ver = mode >= 18;
  const int am = ver ? mode - 26 : -(mode - 10);
  const int aabs = am < 0 ? -am : am;
  const int angle = (am < 0 ? -1 : 1) * ang_tab[aabs];
  const int16_t* mainl = ver ? top : left;                    /* index 0 = corner */
  const int16_t* side = ver ? left : top;
  int16_t buf[3 * 64 + 2];
  int16_t* ref = buf + 64;                                    /* ref[-n .. 2n], ref[0] = corner */
  if (angle < 0)
  {
    for (int i = 0; i <= n; i++) ref[i] = mainl[i];
    int last = (n * angle) >> 5, acc = 128;
    for (int k = -1; k > last; k--) { acc += inv_tab[aabs]; ref[k] = side[acc >> 8]; }
  }
  else
    for (int i = 0; i <= 2 * n; i++) ref[i] = mainl[i];
  for (int j = 0; j < n; j++)                                 /* j runs along the prediction direction's minor axis */
  {
    int pos = (j + 1) * angle, ip = pos >> 5, fr = pos & 31;
    for (int i = 0; i < n; i++)
    {
      int v;
      if (angle == 0) v = ref[i + 1];
      else if (fr)    v = ((32 - fr) * ref[i + ip + 1] + fr * ref[i + ip + 2] + 16) >> 5;
      else            v = ref[i + ip + 1];
      if (angle == 0 && i == 0 && edge_filters && n <= 16) v = clip_pel(v + ((side[j + 1] - side[0]) >> 1), bit_depth);
      if (ver) dst[j * sd + i] = (int16_t)v; else dst[i * sd + j] = (int16_t)v;
    }
  }
}

/* out[35]: Hadamard distortion of every mode, the uiSad of TEncSearch.cpp:2280 (setDistParam ..., bUseHadamard at :2268;
 * TComRdCost.cpp:380-392 selects DF_HADS*, i.e. xGetHADs). */
void hmo_intra_modes_had(const int16_t* org, int so, const int16_t* top_unf, const int16_t* left_unf, const int16_t* top_flt,
                         const int16_t* left_flt, int n, int bit_depth, int above_ok, int left_ok, uint32_t* out)
{
  int16_t pred[64 * 64];
  for (int mode = 0; mode < 35; mode++)
  {
    int f = hmo_intra_use_filtered(mode, n);
    hmo_intra_predict(mode, f ? top_flt : top_unf, f ? left_flt : left_unf, n, bit_depth, above_ok, left_ok, 1, pred, n);
    out[mode] = hmo_had(org, so, pred, n, n, n, bit_depth);
  }
}
