// C-ABI frontend of libhmb200.so (include/hmb200.h): plane registry, job scheduling, kernel launches.
// One caller thread per process (the reference path is non-reentrant too: shared m_cDistParam / m_filteredBlock,
// TLibEncoder/TEncSearch.h:113).  No CPU fallback anywhere: without a usable sm_100 device every compute entry
// fails with HMB200_ERR_CUDA.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "hmb200_device.cuh"
#include "hmb200_generic.cuh"
#include "hmb200_search8.cuh"
#include "hmb200_search8_cu.cuh"
#include "hmb200_search16_cu.cuh"
#include "hmb200_intra.cuh"
#include "hmb200_frac.cuh"
#include "hmb200_tz.cuh"

using namespace hmb200;

namespace {

struct Plane {
  bool used = false;
  DevPlane d{};
  const int16_t* host_lo = nullptr;   // first Pel of the registered host buffer (row -my, col -mx)
  const int16_t* host_hi = nullptr;   // one past the last Pel
  int host_stride = 0;
  int kind = 0, poc = 0;
  size_t bytes = 0;
  cudaEvent_t ev_ready = nullptr;     // uploaded on the upload stream: recorded behind the border extension, consumed by get_plane
  cudaEvent_t ev_reuse = nullptr;     // the recycled buffer's last reader on the compute stream (taken from the pool entry)
};

struct PoolBuf { size_t bytes; void* base; cudaEvent_t ev_free; };   // ev_free: behind the last kernel that may read the buffer

constexpr int N_SIDE = 4;

struct State {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};     // side streams for concurrent variant kernels
  cudaStream_t copy = nullptr;                                     // D2H of results behind the next frame pair's kernels (hmb200_fetch_results_async)
  cudaStream_t up = nullptr;                                       // H2D + border extension of page-locked frames behind the current pair's kernels
  void* upstage = nullptr; size_t upstage_bytes = 0;               // device staging of the upload stream (g.dstage belongs to the compute stream)
  std::vector<cudaEvent_t> free_events;
  cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<Plane> planes;
  void* pinned = nullptr; size_t pinned_bytes = 0;
  void* dstage = nullptr; size_t dstage_bytes = 0;     // device staging for plane uploads / per-call blocks
  Plane pattern;                                       // 64x64 int16 pattern buffer for the 1:1 entries
  std::vector<PoolBuf> pool;                           // released plane buffers, recycled by size (no malloc/free per frame)
  uint64_t launches = 0;
  float last_total_ms = 0, last_search_ms = 0, last_frac_ms = 0;
};

State g;
std::string g_err;

int g_err_code = HMB200_OK;
int fail(int code, const std::string& msg) { g_err = msg; g_err_code = code; return code; }
#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(HMB200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                 \
  } while (0)
#define NEED_READY() do { if (!g.ready) return fail(HMB200_ERR_STATE, "hmb200_init has not succeeded"); } while (0)

int ensure_pinned(size_t bytes) {
  if (bytes <= g.pinned_bytes) return HMB200_OK;
  if (g.pinned) cudaFreeHost(g.pinned);
  g.pinned = nullptr; g.pinned_bytes = 0;
  CUDA_TRY(cudaMallocHost(&g.pinned, bytes));
  g.pinned_bytes = bytes;
  return HMB200_OK;
}
int ensure_dstage(size_t bytes) {
  if (bytes <= g.dstage_bytes) return HMB200_OK;
  if (g.dstage) cudaFree(g.dstage);
  g.dstage = nullptr; g.dstage_bytes = 0;
  CUDA_TRY(cudaMalloc(&g.dstage, bytes));
  g.dstage_bytes = bytes;
  return HMB200_OK;
}

cudaEvent_t take_event() {
  if (!g.free_events.empty()) { cudaEvent_t e = g.free_events.back(); g.free_events.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  return e;
}
void give_event(cudaEvent_t e) { if (e) g.free_events.push_back(e); }

int alloc_plane_slot() {
  for (size_t i = 0; i < g.planes.size(); i++) if (!g.planes[i].used) return (int)i;
  g.planes.emplace_back();
  return (int)g.planes.size() - 1;
}

int make_plane(Plane& p, int width, int height, int mx, int my, int bit_depth) {
  int bps = bit_depth > 8 ? 2 : 1;
  int total_w = width + 2 * mx, total_h = height + 2 * my;
  int pitch_bytes = ((total_w * bps + 127) / 128) * 128;
  p.d.pitch = pitch_bytes / bps;
  p.d.width = width; p.d.height = height; p.d.margin_x = mx; p.d.margin_y = my;
  p.d.bytes_per_sample = bps; p.d.bit_depth = bit_depth;
  p.bytes = (size_t)pitch_bytes * total_h;
  p.d.base = nullptr;
  for (size_t i = 0; i < g.pool.size(); i++)
    if (g.pool[i].bytes == p.bytes) { p.d.base = g.pool[i].base; p.ev_reuse = g.pool[i].ev_free; g.pool.erase(g.pool.begin() + i); break; }
  if (!p.d.base) CUDA_TRY(cudaMalloc(&p.d.base, p.bytes));
  p.used = true;
  return HMB200_OK;
}

// Every consumer of a plane on the compute stream goes through here: a plane that was uploaded on the upload stream
// becomes visible to the compute stream (and the side streams forked from it) by one event wait.
Plane* get_plane(int id) {
  if (id < 0 || id >= (int)g.planes.size() || !g.planes[id].used) return nullptr;
  Plane& p = g.planes[id];
  if (p.ev_ready) { cudaStreamWaitEvent(g.stream, p.ev_ready, 0); give_event(p.ev_ready); p.ev_ready = nullptr; }
  return &p;
}

// which registered plane does a host Pel* fall into?  (1:1 entries hand us raw pointers into TComPicYuv buffers)
Plane* find_plane_by_host(const int16_t* ptr, int* x, int* y) {
  for (auto& p : g.planes) {
    if (!p.used || !p.host_lo) continue;
    if (ptr >= p.host_lo && ptr < p.host_hi) {
      ptrdiff_t off = ptr - p.host_lo;
      *y = (int)(off / p.host_stride) - p.d.margin_y;
      *x = (int)(off % p.host_stride) - p.d.margin_x;
      return &p;
    }
  }
  return nullptr;
}

bool supported_pu(int w, int h) { return w >= 4 && h >= 4 && w <= 64 && h <= 64 && (w % 4) == 0 && (h % 2) == 0; }

template <typename RefT, typename OrgT>
void launch_generic(const SearchTask* d_tasks, hmb200_pu_result* d_res, int n, const DevPlane& cur, const DevPlane& ref,
                    int flags, bool do_search, cudaEvent_t mid, const int* d_index = nullptr, int n_index = 0) {
  if (do_search) {
    const int blocks = d_index ? n_index : n;
    if (blocks > 0) {
      k_search_generic<RefT, OrgT><<<blocks, 256, 0, g.stream>>>(d_tasks, d_res, cur, ref, d_index);
      g.launches++;
    }
  }
  if (mid) cudaEventRecord(mid, g.stream);
  if (flags & HMB200_FLAG_FRAC) {
    k_frac_generic<RefT, OrgT><<<n, FRAC_THREADS, FRAC_SMEM_BYTES, g.stream>>>(d_tasks, d_res, cur, ref,
                                                                               (flags & HMB200_FLAG_HADME) ? 1 : 0);
    g.launches++;
  }
}

void dispatch_generic(const SearchTask* d_tasks, hmb200_pu_result* d_res, int n, const DevPlane& cur, const DevPlane& ref,
                      int flags, bool do_search, cudaEvent_t mid) {
  if (n <= 0) { if (mid) cudaEventRecord(mid, g.stream); return; }
  bool r8 = ref.bytes_per_sample == 1, o8 = cur.bytes_per_sample == 1;
  if (r8 && o8)        launch_generic<uint8_t, uint8_t>(d_tasks, d_res, n, cur, ref, flags, do_search, mid);
  else if (r8 && !o8)  launch_generic<uint8_t, int16_t>(d_tasks, d_res, n, cur, ref, flags, do_search, mid);
  else if (!r8 && o8)  launch_generic<int16_t, uint8_t>(d_tasks, d_res, n, cur, ref, flags, do_search, mid);
  else                 launch_generic<int16_t, int16_t>(d_tasks, d_res, n, cur, ref, flags, do_search, mid);
}

} // namespace

// ------------------------------------------------------------------------------------------------------------------
struct hmb200_prepared {
  int n = 0, flags = 0, bit_depth = 8;
  std::vector<SearchTask> tasks;
  SearchTask* d_tasks = nullptr;
  hmb200_pu_result* d_results = nullptr;
  uint64_t cand_sads = 0, abs_diffs = 0;
  Search8Schedule sched;          // tiled 8-bit kernel schedule (empty when not applicable)
  CuSchedule cu;                  // CU-fused bundles (PUs of one CU sharing window and predictor)
  FracSchedule frac;              // tile tables of the batched quarter-pel refinement
  hmb200_tz_extra* d_tz = nullptr;    // HMB200_FLAG_TZ: per-PU CU geometry / 2Nx2N integer MV
  TzParams tz{0, 0, 0, 0, 0};
  cudaEvent_t ev_done = nullptr, ev_fetched = nullptr;   // end of the last run on the compute stream / of the last asynchronous fetch
  bool fetch_pending = false;
};

extern "C" {

const char* hmb200_last_error(void) { return g_err.c_str(); }
uint64_t hmb200_launch_count(void) { return g.launches; }

int hmb200_init(int device) {
  if (g.ready && g.device == device) return HMB200_OK;
  if (g.ready) hmb200_shutdown();
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(HMB200_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)");
  if (device < 0 || device >= count) return fail(HMB200_ERR_ARG, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(HMB200_ERR_CUDA, "libhmb200 is built for sm_100a only; found " + std::string(prop.name));
  g.sm_count = prop.multiProcessorCount;
  CUDA_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  for (auto& ev : g.ev) CUDA_TRY(cudaEventCreate(&ev));
  for (auto& st : g.side) CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&g.copy, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&g.up, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&g.ev_fork, cudaEventDisableTiming));
  for (auto& ev : g.ev_join) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  int rc = search8_configure(&g_err);
  if (rc != HMB200_OK) return rc;
  if ((rc = cu_configure(&g_err)) != HMB200_OK) return rc;
  if ((rc = cu16_configure(&g_err)) != HMB200_OK) return rc;
  g.device = device;
  g.ready = true;
  // pattern buffer for the 1:1 entries: 64x64 int16, no margins
  g.pattern = Plane();
  rc = make_plane(g.pattern, 64, 64, 0, 0, 16);
  if (rc != HMB200_OK) { g.ready = false; return rc; }
  g.launches = 0;
  return HMB200_OK;
}

void hmb200_shutdown(void) {
  if (!g.ready) return;
  cudaStreamSynchronize(g.stream);
  for (auto& p : g.planes) if (p.used && p.d.base) cudaFree(p.d.base);
  g.planes.clear();
  for (auto& b : g.pool) { cudaFree(b.base); if (b.ev_free) cudaEventDestroy(b.ev_free); }
  g.pool.clear();
  for (auto& pl : g.planes) { if (pl.ev_ready) cudaEventDestroy(pl.ev_ready); if (pl.ev_reuse) cudaEventDestroy(pl.ev_reuse); pl.ev_ready = pl.ev_reuse = nullptr; }
  for (auto e : g.free_events) cudaEventDestroy(e);
  g.free_events.clear();
  if (g.upstage) { cudaFree(g.upstage); g.upstage = nullptr; g.upstage_bytes = 0; }
  if (g.pattern.d.base) cudaFree(g.pattern.d.base);
  g.pattern = Plane();
  if (g.pinned) cudaFreeHost(g.pinned);
  if (g.dstage) cudaFree(g.dstage);
  g.pinned = nullptr; g.pinned_bytes = 0; g.dstage = nullptr; g.dstage_bytes = 0;
  for (auto& ev : g.ev) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
  for (auto& st : g.side) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); st = nullptr; }
  if (g.copy) { cudaStreamSynchronize(g.copy); cudaStreamDestroy(g.copy); g.copy = nullptr; }
  if (g.up) { cudaStreamSynchronize(g.up); cudaStreamDestroy(g.up); g.up = nullptr; }
  if (g.ev_fork) { cudaEventDestroy(g.ev_fork); g.ev_fork = nullptr; }
  for (auto& ev : g.ev_join) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
  cudaStreamDestroy(g.stream); g.stream = nullptr;
  g.ready = false; g.device = -1;
}

void* hmb200_host_alloc(size_t bytes) {
  if (!g.ready) { fail(HMB200_ERR_STATE, "hmb200_init has not succeeded"); return nullptr; }
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) { fail(HMB200_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(cudaGetLastError())); return nullptr; }
  return p;
}
void hmb200_host_free(void* p) { if (p) cudaFreeHost(p); }

int hmb200_sync(void) {
  NEED_READY();
  CUDA_TRY(cudaStreamSynchronize(g.up));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.copy));
  return HMB200_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// host-side window / job-list logic (no GPU)
// ------------------------------------------------------------------------------------------------------------------
static void clip_mv(int& x, int& y, int cu_x, int cu_y, int pic_w, int pic_h, int max_cu_w, int max_cu_h) {
  // TLibCommon/TComDataCU.cpp:2788-2801
  const int off = 8;
  int hmax = (pic_w + off - cu_x - 1) * 4, hmin = (-max_cu_w - off - cu_x + 1) * 4;
  int vmax = (pic_h + off - cu_y - 1) * 4, vmin = (-max_cu_h - off - cu_y + 1) * 4;
  x = std::min(hmax, std::max(hmin, x));
  y = std::min(vmax, std::max(vmin, y));
}

uint32_t hmb200_motion_lambda_cost(double lambda) {
  // TLibCommon/TComRdCost.cpp:195-220 setLambda: m_uiLambdaMotionSAD[0]; getMotionCost(true, 0, false) copies it to m_uiCost
  return (uint32_t)std::floor(65536.0 * std::sqrt(lambda));
}

void hmb200_set_search_range(hmb200_mv pred, int search_range, int cu_x, int cu_y, int pic_w, int pic_h, int max_cu_w, int max_cu_h,
                             hmb200_mv* lt, hmb200_mv* rb) {
  int px = pred.x, py = pred.y;
  clip_mv(px, py, cu_x, cu_y, pic_w, pic_h, max_cu_w, max_cu_h);
  // TComMv components are Shorts: the sums wrap to 16 bits before the second clip (TLibCommon/TComMv.h:53-54)
  int lx = (int16_t)(px - (search_range << 2)), ly = (int16_t)(py - (search_range << 2));
  int rx = (int16_t)(px + (search_range << 2)), ry = (int16_t)(py + (search_range << 2));
  clip_mv(lx, ly, cu_x, cu_y, pic_w, pic_h, max_cu_w, max_cu_h);
  clip_mv(rx, ry, cu_x, cu_y, pic_w, pic_h, max_cu_w, max_cu_h);
  lt->x = lx >> 2; lt->y = ly >> 2; rb->x = rx >> 2; rb->y = ry >> 2;      // arithmetic shifts of Shorts
}

static int canonical_jobs_of_ctus(int pic_w, int pic_h, int max_cu, int search_range, uint32_t lambda_cost, hmb200_mv pred,
                                  int cx0, int cx1, int cy0, int cy1, int raster_first, int raster_end,
                                  hmb200_pu_job* jobs, int capacity) {
  const int ctus_x = (pic_w + max_cu - 1) / max_cu;
  int count = 0;
  auto emit = [&](int cu_x, int cu_y, int px, int py, int w, int h) {
    if (count < capacity && jobs) {
      hmb200_pu_job j;
      j.pu_x = px; j.pu_y = py; j.w = w; j.h = h;
      hmb200_mv lt, rb;
      hmb200_set_search_range(pred, search_range, cu_x, cu_y, pic_w, pic_h, max_cu, max_cu, &lt, &rb);
      j.lt_x = lt.x; j.lt_y = lt.y; j.rb_x = rb.x; j.rb_y = rb.y;
      j.pred_x = pred.x; j.pred_y = pred.y; j.lambda_cost = lambda_cost; j.reserved = 0;
      jobs[count] = j;
    }
    count++;
  };
  for (int cty = cy0; cty < cy1; cty++)
    for (int ctx = cx0; ctx < cx1; ctx++) {
      const int ctu = cty * ctus_x + ctx;
      if (ctu < raster_first || ctu >= raster_end) continue;
      const int ox = ctx * max_cu, oy = cty * max_cu;
      for (int s = max_cu; s >= 8; s >>= 1) {
        for (int cy = oy; cy < oy + max_cu; cy += s)
          for (int cx = ox; cx < ox + max_cu; cx += s) {
            if (cx + s > pic_w || cy + s > pic_h) continue;          // the encoder only tests CUs inside the picture
            const int q = s >> 2;
            emit(cx, cy, cx, cy, s, s);                                              // SIZE_2Nx2N
            emit(cx, cy, cx, cy, s, s / 2);  emit(cx, cy, cx, cy + s / 2, s, s / 2);  // SIZE_2NxN
            emit(cx, cy, cx, cy, s / 2, s);  emit(cx, cy, cx + s / 2, cy, s / 2, s);  // SIZE_Nx2N
            if (s > 8) {                                                              // AMP: not for the smallest CU
              emit(cx, cy, cx, cy, s, q);      emit(cx, cy, cx, cy + q, s, s - q);    // SIZE_2NxnU
              emit(cx, cy, cx, cy, s, s - q);  emit(cx, cy, cx, cy + s - q, s, q);    // SIZE_2NxnD
              emit(cx, cy, cx, cy, q, s);      emit(cx, cy, cx + q, cy, s - q, s);    // SIZE_nLx2N
              emit(cx, cy, cx, cy, s - q, s);  emit(cx, cy, cx + s - q, cy, q, s);    // SIZE_nRx2N
            }
          }
      }
    }
  return count;
}

int hmb200_build_canonical_jobs(int pic_w, int pic_h, int max_cu, int search_range, uint32_t lambda_cost, hmb200_mv pred,
                                int ctu_first, int ctu_count, hmb200_pu_job* jobs, int capacity) {
  if (pic_w <= 0 || pic_h <= 0 || max_cu < 8 || max_cu > 64 || (max_cu & (max_cu - 1))) return HMB200_ERR_ARG;
  const int ctus_x = (pic_w + max_cu - 1) / max_cu, ctus_y = (pic_h + max_cu - 1) / max_cu;
  const int n_ctus = ctus_x * ctus_y;
  if (ctu_first < 0) ctu_first = 0;
  const int ctu_end = (ctu_count < 0) ? n_ctus : std::min(n_ctus, ctu_first + ctu_count);
  return canonical_jobs_of_ctus(pic_w, pic_h, max_cu, search_range, lambda_cost, pred, 0, ctus_x, 0, ctus_y, ctu_first, ctu_end, jobs, capacity);
}

int hmb200_tile_column_range(int pic_w, int max_cu, int n_columns, int column, int* ctu_x0, int* ctu_x1) {
  if (pic_w <= 0 || max_cu <= 0 || n_columns <= 0 || column < 0 || column >= n_columns || !ctu_x0 || !ctu_x1) return HMB200_ERR_ARG;
  const int ctus_x = (pic_w + max_cu - 1) / max_cu;
  if (n_columns > ctus_x) return HMB200_ERR_ARG;
  *ctu_x0 = (column * ctus_x) / n_columns;               // TLibCommon/TComPicSym.cpp:217-229 (uniform tile spacing)
  *ctu_x1 = ((column + 1) * ctus_x) / n_columns;
  return HMB200_OK;
}

int hmb200_build_canonical_jobs_rect(int pic_w, int pic_h, int max_cu, int search_range, uint32_t lambda_cost, hmb200_mv pred,
                                     int ctu_x0, int ctu_x1, int ctu_y0, int ctu_y1, hmb200_pu_job* jobs, int capacity) {
  if (pic_w <= 0 || pic_h <= 0 || max_cu < 8 || max_cu > 64 || (max_cu & (max_cu - 1))) return HMB200_ERR_ARG;
  const int ctus_x = (pic_w + max_cu - 1) / max_cu, ctus_y = (pic_h + max_cu - 1) / max_cu;
  if (ctu_x0 < 0 || ctu_y0 < 0 || ctu_x1 > ctus_x || ctu_y1 > ctus_y || ctu_x0 > ctu_x1 || ctu_y0 > ctu_y1) return HMB200_ERR_ARG;
  return canonical_jobs_of_ctus(pic_w, pic_h, max_cu, search_range, lambda_cost, pred, ctu_x0, ctu_x1, ctu_y0, ctu_y1, 0, ctus_x * ctus_y, jobs, capacity);
}

// ------------------------------------------------------------------------------------------------------------------
// planes
// ------------------------------------------------------------------------------------------------------------------
int hmb200_register_plane(const int16_t* host_origin, int stride, int width, int height, int margin_x, int margin_y,
                          int bit_depth, int kind, int poc) {
  NEED_READY();
  if (!host_origin || width <= 0 || height <= 0 || margin_x < 0 || margin_y < 0 || stride < width + 2 * margin_x ||
      bit_depth < 8 || bit_depth > 16)
    return fail(HMB200_ERR_ARG, "hmb200_register_plane: bad geometry");
  int id = alloc_plane_slot();
  Plane& p = g.planes[id];
  p = Plane();
  int rc = make_plane(p, width, height, margin_x, margin_y, bit_depth);
  if (rc != HMB200_OK) return rc;
  int total_w = width + 2 * margin_x, total_h = height + 2 * margin_y;
  const int16_t* src = host_origin - (ptrdiff_t)margin_y * stride - margin_x;
  size_t bytes = (size_t)total_w * total_h * sizeof(int16_t);
  if ((rc = ensure_pinned(bytes)) != HMB200_OK || (rc = ensure_dstage(bytes)) != HMB200_OK) { hmb200_release_plane(id); return rc; }
  int16_t* pin = reinterpret_cast<int16_t*>(g.pinned);
  for (int y = 0; y < total_h; y++) memcpy(pin + (size_t)y * total_w, src + (size_t)y * stride, (size_t)total_w * sizeof(int16_t));
  CUDA_TRY(cudaMemcpyAsync(g.dstage, pin, bytes, cudaMemcpyHostToDevice, g.stream));
  dim3 grid((total_w + 255) / 256, total_h);
  if (p.d.bytes_per_sample == 1)
    k_narrow_plane<uint8_t><<<grid, 256, 0, g.stream>>>(reinterpret_cast<const int16_t*>(g.dstage), total_w,
                                                        reinterpret_cast<uint8_t*>(p.d.base), p.d.pitch, total_w, total_h);
  else
    k_narrow_plane<uint16_t><<<grid, 256, 0, g.stream>>>(reinterpret_cast<const int16_t*>(g.dstage), total_w,
                                                         reinterpret_cast<uint16_t*>(p.d.base), p.d.pitch, total_w, total_h);
  g.launches++;
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  p.host_lo = src; p.host_hi = src + (size_t)(total_h - 1) * stride + total_w; p.host_stride = stride;
  p.kind = kind; p.poc = poc;
  return id;
}

int hmb200_register_plane_u8(const uint8_t* host_samples, int stride, int width, int height, int margin_x, int margin_y,
                             int kind, int poc) {
  NEED_READY();
  if (!host_samples || width <= 0 || height <= 0 || stride < width || margin_x < 0 || margin_y < 0)
    return fail(HMB200_ERR_ARG, "hmb200_register_plane_u8: bad geometry");
  int id = alloc_plane_slot();
  Plane& p = g.planes[id];
  p = Plane();
  int rc = make_plane(p, width, height, margin_x, margin_y, 8);
  if (rc != HMB200_OK) return rc;
  size_t bytes = (size_t)width * height;
  cudaPointerAttributes attr;
  const bool user_pinned = stride == width && cudaPointerGetAttributes(&attr, host_samples) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (user_pinned) {
    // page-locked, tightly packed frame: H2D + border extension on the upload stream, behind whatever the compute stream is
    // doing; the source stays the caller's until the copy has run (hmb200_sync or any later fetch / blocking call orders it)
    if (bytes > g.upstage_bytes) {
      CUDA_TRY(cudaStreamSynchronize(g.up));
      if (g.upstage) cudaFree(g.upstage);
      g.upstage = nullptr; g.upstage_bytes = 0;
      CUDA_TRY(cudaMalloc(&g.upstage, bytes));
      g.upstage_bytes = bytes;
    }
    if (p.ev_reuse) { CUDA_TRY(cudaStreamWaitEvent(g.up, p.ev_reuse, 0)); give_event(p.ev_reuse); p.ev_reuse = nullptr; }
    CUDA_TRY(cudaMemcpyAsync(g.upstage, host_samples, bytes, cudaMemcpyHostToDevice, g.up));
    dim3 ugrid((width + 2 * margin_x + 255) / 256, height + 2 * margin_y);
    k_pad_plane_u8<<<ugrid, 256, 0, g.up>>>(reinterpret_cast<const uint8_t*>(g.upstage), width,
                                            reinterpret_cast<uint8_t*>(p.d.base), p.d.pitch, width, height, margin_x, margin_y);
    g.launches++;
    p.ev_ready = take_event();
    CUDA_TRY(cudaEventRecord(p.ev_ready, g.up));
    p.kind = kind; p.poc = poc;
    return id;
  }
  if ((rc = ensure_dstage(bytes)) != HMB200_OK) { hmb200_release_plane(id); return rc; }
  const uint8_t* src = host_samples;                      // page-locked, tightly packed frames go to the device without a staging copy
  if (!user_pinned) {
    if ((rc = ensure_pinned(bytes)) != HMB200_OK) { hmb200_release_plane(id); return rc; }
    uint8_t* pin = reinterpret_cast<uint8_t*>(g.pinned);
    for (int y = 0; y < height; y++) memcpy(pin + (size_t)y * width, host_samples + (size_t)y * stride, (size_t)width);
    src = pin;
  }
  CUDA_TRY(cudaMemcpyAsync(g.dstage, src, bytes, cudaMemcpyHostToDevice, g.stream));
  dim3 grid((width + 2 * margin_x + 255) / 256, height + 2 * margin_y);
  k_pad_plane_u8<<<grid, 256, 0, g.stream>>>(reinterpret_cast<const uint8_t*>(g.dstage), width,
                                             reinterpret_cast<uint8_t*>(p.d.base), p.d.pitch, width, height, margin_x, margin_y);
  g.launches++;
  // a page-locked source stays the caller's until the copy has run (hmb200_sync / any fetch orders it); the library's own
  // staging buffer is reused by the next call, so that path waits here
  if (!user_pinned) CUDA_TRY(cudaStreamSynchronize(g.stream));
  p.kind = kind; p.poc = poc;
  return id;
}

int hmb200_register_plane_yuv(const void* file_luma, int file_is16, int width, int height, int pad_x, int pad_y,
                              int file_bit_depth, int internal_bit_depth, int margin_x, int margin_y, int kind, int poc) {
  NEED_READY();
  if (!file_luma || width <= 0 || height <= 0 || pad_x < 0 || pad_y < 0 || margin_x < 0 || margin_y < 0 || file_bit_depth < 8 ||
      file_bit_depth > 16 || internal_bit_depth < 8 || internal_bit_depth > 14 || (!file_is16 && file_bit_depth > 8))
    return fail(HMB200_ERR_ARG, "hmb200_register_plane_yuv: bad geometry or bit depths");
  int id = alloc_plane_slot();
  Plane& p = g.planes[id];
  p = Plane();
  const int cw = width + pad_x, ch = height + pad_y;
  int rc = make_plane(p, cw, ch, margin_x, margin_y, internal_bit_depth);
  if (rc != HMB200_OK) return rc;
  const size_t bytes = (size_t)width * height * (file_is16 ? 2 : 1);
  if ((rc = ensure_pinned(bytes)) != HMB200_OK || (rc = ensure_dstage(bytes)) != HMB200_OK) { hmb200_release_plane(id); return rc; }
  memcpy(g.pinned, file_luma, bytes);
  CUDA_TRY(cudaMemcpyAsync(g.dstage, g.pinned, bytes, cudaMemcpyHostToDevice, g.stream));
  const int shift = internal_bit_depth - file_bit_depth, maxval = (1 << internal_bit_depth) - 1;
  dim3 grid((cw + 2 * margin_x + 255) / 256, ch + 2 * margin_y);
  if (p.d.bytes_per_sample == 1)
    k_ingest_luma<uint8_t><<<grid, 256, 0, g.stream>>>(reinterpret_cast<const uint8_t*>(g.dstage), file_is16, width, height, pad_x, pad_y,
                                                       shift, maxval, reinterpret_cast<uint8_t*>(p.d.base), p.d.pitch, margin_x, margin_y);
  else
    k_ingest_luma<uint16_t><<<grid, 256, 0, g.stream>>>(reinterpret_cast<const uint8_t*>(g.dstage), file_is16, width, height, pad_x, pad_y,
                                                        shift, maxval, reinterpret_cast<uint16_t*>(p.d.base), p.d.pitch, margin_x, margin_y);
  g.launches++;
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  p.kind = kind; p.poc = poc;
  return id;
}

int hmb200_read_plane(int plane_id, int16_t* dst_origin, int dst_stride) {
  NEED_READY();
  Plane* p = get_plane(plane_id);
  if (!p || !dst_origin) return fail(HMB200_ERR_ARG, "hmb200_read_plane: unknown plane");
  int total_w = p->d.width + 2 * p->d.margin_x, total_h = p->d.height + 2 * p->d.margin_y;
  if (dst_stride < total_w) return fail(HMB200_ERR_ARG, "hmb200_read_plane: dst_stride too small");
  size_t bytes = (size_t)total_w * total_h * sizeof(int16_t);
  int rc;
  if ((rc = ensure_pinned(bytes)) != HMB200_OK || (rc = ensure_dstage(bytes)) != HMB200_OK) return rc;
  dim3 grid((total_w + 255) / 256, total_h);
  if (p->d.bytes_per_sample == 1)
    k_widen_plane<uint8_t><<<grid, 256, 0, g.stream>>>(reinterpret_cast<const uint8_t*>(p->d.base), p->d.pitch,
                                                       reinterpret_cast<int16_t*>(g.dstage), total_w, total_w, total_h);
  else
    k_widen_plane<uint16_t><<<grid, 256, 0, g.stream>>>(reinterpret_cast<const uint16_t*>(p->d.base), p->d.pitch,
                                                        reinterpret_cast<int16_t*>(g.dstage), total_w, total_w, total_h);
  g.launches++;
  CUDA_TRY(cudaMemcpyAsync(g.pinned, g.dstage, bytes, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  const int16_t* pin = reinterpret_cast<const int16_t*>(g.pinned);
  int16_t* dst = dst_origin - (ptrdiff_t)p->d.margin_y * dst_stride - p->d.margin_x;
  for (int y = 0; y < total_h; y++) memcpy(dst + (size_t)y * dst_stride, pin + (size_t)y * total_w, (size_t)total_w * sizeof(int16_t));
  return HMB200_OK;
}

void hmb200_release_plane(int plane_id) {
  Plane* p = get_plane(plane_id);            // also orders a still pending upload of this plane before the compute stream's tail
  if (!p) return;
  if (p->d.base) {
    if (g.pool.size() < 16) {
      // no host synchronisation: the buffer goes back to the pool with an event behind the last kernel that may read it
      // (side streams are joined into the compute stream before a run ends); whoever reuses it on the upload stream waits on it
      cudaEvent_t ev = p->ev_reuse ? p->ev_reuse : take_event();
      p->ev_reuse = nullptr;
      cudaEventRecord(ev, g.stream);
      g.pool.push_back(PoolBuf{p->bytes, p->d.base, ev});
    } else {
      cudaStreamSynchronize(g.up);
      cudaStreamSynchronize(g.stream);
      cudaFree(p->d.base);
    }
  }
  give_event(p->ev_reuse);
  *p = Plane();
}

// ------------------------------------------------------------------------------------------------------------------
// distortion table
// ------------------------------------------------------------------------------------------------------------------
uint32_t hmb200_dist(const hmb200_dist_param* dp) {
  auto die = [](const char* m) { fprintf(stderr, "hmb200_dist: %s (%s)\n", m, g_err.c_str()); abort(); };
  if (!g.ready) die("library not initialised");
  if (!dp || !dp->pOrg || !dp->pCur) die("null DistParam");
  if (dp->bApplyWeight) die("weighted prediction is out of scope (bApplyWeight must be false)");
  if (dp->iStep != 1) die("iStep must be 1");
  int w = dp->iCols, h = dp->iRows;
  if (w <= 0 || h <= 0 || w > 64 || h > 64) die("block size out of range");
  size_t n = (size_t)w * h;
  if (ensure_pinned(2 * n * sizeof(int16_t) + 64) != HMB200_OK || ensure_dstage(2 * n * sizeof(int16_t) + sizeof(DistTask) + 64) != HMB200_OK)
    die("allocation failed");
  int16_t* pin = reinterpret_cast<int16_t*>(g.pinned);
  for (int y = 0; y < h; y++) {
    memcpy(pin + (size_t)y * w, dp->pOrg + (ptrdiff_t)y * dp->iStrideOrg, (size_t)w * sizeof(int16_t));
    memcpy(pin + n + (size_t)y * w, dp->pCur + (ptrdiff_t)y * dp->iStrideCur, (size_t)w * sizeof(int16_t));
  }
  int16_t* d = reinterpret_cast<int16_t*>(g.dstage);
  size_t task_off = ((2 * n * sizeof(int16_t) + 15) / 16) * 16;
  DistTask t{d, d + n, w, w, w, h, dp->iSubShift, 0};
  memcpy(reinterpret_cast<char*>(g.pinned) + task_off, &t, sizeof(t));
  uint32_t* d_out = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(g.dstage) + task_off + sizeof(DistTask));
  if (cudaMemcpyAsync(g.dstage, g.pinned, task_off + sizeof(DistTask), cudaMemcpyHostToDevice, g.stream) != cudaSuccess) die("H2D failed");
  k_dist_generic<int16_t><<<1, 128, 0, g.stream>>>(reinterpret_cast<const DistTask*>(reinterpret_cast<char*>(g.dstage) + task_off),
                                                   d_out, 1, dp->func, dp->bitDepth);
  g.launches++;
  uint32_t out = 0;
  if (cudaMemcpyAsync(&out, d_out, sizeof(out), cudaMemcpyDeviceToHost, g.stream) != cudaSuccess ||
      cudaStreamSynchronize(g.stream) != cudaSuccess) { g_err = cudaGetErrorString(cudaGetLastError()); die("kernel failed"); }
  return out;
}

int hmb200_dist_batch(int func, int bit_depth, int n, const hmb200_dist_desc* descs, uint32_t* out) {
  NEED_READY();
  if (n <= 0) return HMB200_OK;
  if (!descs || !out || func < 0 || func > 3) return fail(HMB200_ERR_ARG, "hmb200_dist_batch: bad arguments");
  std::vector<DistTask> tasks((size_t)n);
  int bps = 0;
  for (int i = 0; i < n; i++) {
    const hmb200_dist_desc& d = descs[i];
    Plane* po = get_plane(d.org_plane); Plane* pc = get_plane(d.cur_plane);
    if (!po || !pc) return fail(HMB200_ERR_ARG, "hmb200_dist_batch: unknown plane");
    if (po->d.bytes_per_sample != pc->d.bytes_per_sample) return fail(HMB200_ERR_ARG, "hmb200_dist_batch: planes differ in sample size");
    if (bps && bps != po->d.bytes_per_sample) return fail(HMB200_ERR_ARG, "hmb200_dist_batch: mixed sample sizes in one batch");
    bps = po->d.bytes_per_sample;
    if (d.w <= 0 || d.h <= 0 || d.w > 64 || d.h > 64) return fail(HMB200_ERR_ARG, "hmb200_dist_batch: block size out of range");
    auto addr = [&](Plane* p, int x, int y) {
      return reinterpret_cast<char*>(p->d.base) + ((size_t)(y + p->d.margin_y) * p->d.pitch + (x + p->d.margin_x)) * bps;
    };
    tasks[i] = DistTask{addr(po, d.org_x, d.org_y), addr(pc, d.cur_x, d.cur_y), po->d.pitch, pc->d.pitch, d.w, d.h, d.sub_shift, 0};
  }
  size_t tb = tasks.size() * sizeof(DistTask), ob = (size_t)n * sizeof(uint32_t);
  int rc;
  if ((rc = ensure_dstage(tb + ob)) != HMB200_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(g.dstage, tasks.data(), tb, cudaMemcpyHostToDevice, g.stream));
  uint32_t* d_out = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(g.dstage) + tb);
  int blocks = (n + 3) / 4;
  if (bps == 1) k_dist_generic<uint8_t><<<blocks, 128, 0, g.stream>>>(reinterpret_cast<const DistTask*>(g.dstage), d_out, n, func, bit_depth);
  else          k_dist_generic<int16_t><<<blocks, 128, 0, g.stream>>>(reinterpret_cast<const DistTask*>(g.dstage), d_out, n, func, bit_depth);
  g.launches++;
  CUDA_TRY(cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  return HMB200_OK;
}

int hmb200_mc_dist_batch(int cur_plane, int ref_plane, int func, int n, const hmb200_mc_desc* descs, uint32_t* out) {
  NEED_READY();
  if (n <= 0) return HMB200_OK;
  Plane* pc = get_plane(cur_plane); Plane* pr = get_plane(ref_plane);
  if (!pc || !pr) return fail(HMB200_ERR_ARG, "hmb200_mc_dist_batch: unknown plane");
  if (!descs || !out || (func != HMB200_DF_SAD && func != HMB200_DF_SADS && func != HMB200_DF_HADS))
    return fail(HMB200_ERR_ARG, "hmb200_mc_dist_batch: bad arguments (func must be SAD or HADS)");
  if (pc->d.bytes_per_sample != pr->d.bytes_per_sample) return fail(HMB200_ERR_ARG, "hmb200_mc_dist_batch: planes differ in sample size");
  std::vector<SearchTask> tasks((size_t)n);
  std::vector<hmb200_pu_result> mv((size_t)n);
  for (int i = 0; i < n; i++) {
    const hmb200_mc_desc& d = descs[i];
    if (!supported_pu(d.w, d.h)) return fail(HMB200_ERR_ARG, "hmb200_mc_dist_batch: unsupported PU size in descriptor " + std::to_string(i));
    tasks[i] = SearchTask{d.pu_x, d.pu_y, d.pu_x, d.pu_y, d.w, d.h, 0, 0, 0, 0, 0, 0, 0u, 0};
    mv[i] = hmb200_pu_result{0, 0, 0u, 0, 0, d.mv_x, d.mv_y, 0u};          // k_frac_tiles stage 2 reads the MV from qter_x / qter_y
  }
  FracSchedule fs;
  std::string why;
  if (!frac_build_schedule(tasks, g.stream, &fs, &why)) return fail(HMB200_ERR_CUDA, why);
  const size_t tb = (size_t)n * sizeof(SearchTask), rb = (size_t)n * sizeof(hmb200_pu_result), ob = (size_t)n * sizeof(uint32_t);
  int rc = ensure_dstage(tb + rb + ob + 256);
  if (rc != HMB200_OK) { frac_free_schedule(&fs); return rc; }
  char* base = reinterpret_cast<char*>(g.dstage);
  SearchTask* d_tasks = reinterpret_cast<SearchTask*>(base);
  hmb200_pu_result* d_mv = reinterpret_cast<hmb200_pu_result*>(base + ((tb + 63) & ~(size_t)63));
  uint32_t* d_out = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(d_mv) + ((rb + 63) & ~(size_t)63));
  cudaError_t e = cudaMemcpyAsync(d_tasks, tasks.data(), tb, cudaMemcpyHostToDevice, g.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_mv, mv.data(), rb, cudaMemcpyHostToDevice, g.stream);
  int nl = -1;
  if (e == cudaSuccess) {
    const bool had = func == HMB200_DF_HADS;
    nl = pr->d.bytes_per_sample == 1 ? mc_launch<uint8_t, uint8_t>(fs, d_tasks, d_mv, d_out, pc->d, pr->d, had, g.stream)
                                     : mc_launch<int16_t, int16_t>(fs, d_tasks, d_mv, d_out, pc->d, pr->d, had, g.stream);
  }
  if (nl >= 0) { g.launches += (uint64_t)nl; e = cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, g.stream); }
  if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
  frac_free_schedule(&fs);
  if (nl < 0 || e != cudaSuccess) return fail(HMB200_ERR_CUDA, std::string("hmb200_mc_dist_batch: ") + cudaGetErrorString(cudaGetLastError()));
  return HMB200_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// intra first pass
// ------------------------------------------------------------------------------------------------------------------
static int intra_run(const DevPlane& org, int nblocks, const hmb200_intra_block* blocks, const int16_t* refs, int n_ref_samples,
                     const void* org_upload, size_t org_upload_bytes, uint32_t* out) {
  std::vector<IntraBlockDev> hb((size_t)nblocks);
  std::vector<int32_t> ord4, ord8;
  for (int i = 0; i < nblocks; i++) {
    const hmb200_intra_block& b = blocks[i];
    if (!(b.n == 4 || b.n == 8 || b.n == 16 || b.n == 32 || b.n == 64) || b.ref_off < 0 || (int64_t)b.ref_off + 4 * (2 * b.n + 1) > n_ref_samples)
      return fail(HMB200_ERR_ARG, "hmb200_intra_modes_had: bad block size or reference-line offset in block " + std::to_string(i));
    if (!org_upload && (b.x < -org.margin_x || b.y < -org.margin_y || b.x + b.n > org.width + org.margin_x || b.y + b.n > org.height + org.margin_y))
      return fail(HMB200_ERR_ARG, "hmb200_intra_modes_had: block " + std::to_string(i) + " leaves the padded plane");
    hb[(size_t)i] = IntraBlockDev{b.x, b.y, b.n, b.ref_off, b.flags, 0};
    (b.n == 4 ? ord4 : ord8).push_back(i);
  }
  auto up64 = [](size_t v) { return (v + 63) & ~(size_t)63; };
  const size_t bb = up64(hb.size() * sizeof(IntraBlockDev)), ob = up64((size_t)nblocks * sizeof(int32_t)), rb = up64((size_t)n_ref_samples * 2),
               ub = up64(org_upload_bytes), outb = (size_t)nblocks * INTRA_MODES * sizeof(uint32_t);
  int rc = ensure_dstage(bb + ob + rb + ub + outb + 256);
  if (rc != HMB200_OK) return rc;
  char* base = reinterpret_cast<char*>(g.dstage);
  IntraBlockDev* d_blocks = reinterpret_cast<IntraBlockDev*>(base);
  int32_t* d_ord = reinterpret_cast<int32_t*>(base + bb);
  int16_t* d_refs = reinterpret_cast<int16_t*>(base + bb + ob);
  void* d_org = base + bb + ob + rb;
  uint32_t* d_out = reinterpret_cast<uint32_t*>(base + bb + ob + rb + ub);
  std::vector<int32_t> ord(ord8);
  ord.insert(ord.end(), ord4.begin(), ord4.end());
  CUDA_TRY(cudaMemcpyAsync(d_blocks, hb.data(), hb.size() * sizeof(IntraBlockDev), cudaMemcpyHostToDevice, g.stream));
  CUDA_TRY(cudaMemcpyAsync(d_ord, ord.data(), ord.size() * sizeof(int32_t), cudaMemcpyHostToDevice, g.stream));
  CUDA_TRY(cudaMemcpyAsync(d_refs, refs, (size_t)n_ref_samples * 2, cudaMemcpyHostToDevice, g.stream));
  DevPlane pl = org;
  if (org_upload) {
    CUDA_TRY(cudaMemcpyAsync(d_org, org_upload, org_upload_bytes, cudaMemcpyHostToDevice, g.stream));
    pl.base = d_org;
  }
  const int bd = pl.bit_depth;
  if (!ord8.empty()) {
    if (pl.bytes_per_sample == 1) k_intra_modes_had<8, uint8_t><<<(int)ord8.size(), INTRA_THREADS, 0, g.stream>>>(d_blocks, d_ord, d_refs, d_out, pl, bd);
    else                          k_intra_modes_had<8, int16_t><<<(int)ord8.size(), INTRA_THREADS, 0, g.stream>>>(d_blocks, d_ord, d_refs, d_out, pl, bd);
    g.launches++;
  }
  if (!ord4.empty()) {
    if (pl.bytes_per_sample == 1) k_intra_modes_had<4, uint8_t><<<(int)ord4.size(), INTRA_THREADS, 0, g.stream>>>(d_blocks, d_ord + ord8.size(), d_refs, d_out, pl, bd);
    else                          k_intra_modes_had<4, int16_t><<<(int)ord4.size(), INTRA_THREADS, 0, g.stream>>>(d_blocks, d_ord + ord8.size(), d_refs, d_out, pl, bd);
    g.launches++;
  }
  CUDA_TRY(cudaMemcpyAsync(out, d_out, outb, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  return HMB200_OK;
}

int hmb200_intra_modes_had_batch(int org_plane, int nblocks, const hmb200_intra_block* blocks, const int16_t* refs, int n_ref_samples,
                                 uint32_t* out) {
  NEED_READY();
  if (nblocks <= 0) return HMB200_OK;
  Plane* po = get_plane(org_plane);
  if (!po) return fail(HMB200_ERR_ARG, "hmb200_intra_modes_had_batch: unknown plane");
  if (!blocks || !refs || !out || n_ref_samples <= 0) return fail(HMB200_ERR_ARG, "hmb200_intra_modes_had_batch: bad arguments");
  return intra_run(po->d, nblocks, blocks, refs, n_ref_samples, nullptr, 0, out);
}

int hmb200_intra_modes_had(const int16_t* org, int org_stride, const int16_t* ref_unf, const int16_t* ref_flt, int n, int bit_depth,
                           int above, int left, uint32_t* out) {
  NEED_READY();
  if (!org || !ref_unf || !ref_flt || !out || org_stride < n || bit_depth < 8 || bit_depth > 14 ||
      !(n == 4 || n == 8 || n == 16 || n == 32 || n == 64))
    return fail(HMB200_ERR_ARG, "hmb200_intra_modes_had: bad arguments");
  const int L = 2 * n + 1;
  std::vector<int16_t> lines((size_t)4 * L), blk((size_t)n * n);
  for (int i = 0; i < L; i++) {                               // row 0 and column 0 of the (2n+1)-strided predictor buffers
    lines[i] = ref_unf[i];             lines[L + i] = ref_unf[(size_t)i * L];
    lines[2 * L + i] = ref_flt[i];     lines[3 * L + i] = ref_flt[(size_t)i * L];
  }
  for (int y = 0; y < n; y++) memcpy(&blk[(size_t)y * n], org + (size_t)y * org_stride, (size_t)n * 2);
  DevPlane pl{};
  pl.base = nullptr; pl.pitch = n; pl.width = n; pl.height = n; pl.margin_x = 0; pl.margin_y = 0; pl.bytes_per_sample = 2; pl.bit_depth = bit_depth;
  hmb200_intra_block b{0, 0, n, 0, (above ? 1 : 0) | (left ? 2 : 0), 0};
  return intra_run(pl, 1, &b, lines.data(), 4 * L, blk.data(), blk.size() * 2, out);
}

// ------------------------------------------------------------------------------------------------------------------
// batched searches
// ------------------------------------------------------------------------------------------------------------------
hmb200_prepared* hmb200_prepare_jobs(const hmb200_pu_job* jobs, int njobs, int flags, int bit_depth) {
  if (!g.ready) { fail(HMB200_ERR_STATE, "hmb200_init has not succeeded"); return nullptr; }
  if (njobs < 0 || (njobs > 0 && !jobs)) { fail(HMB200_ERR_ARG, "hmb200_prepare_jobs: bad arguments"); return nullptr; }
  auto* p = new hmb200_prepared();
  p->n = njobs; p->flags = flags; p->bit_depth = bit_depth;
  p->tasks.resize((size_t)njobs);
  for (int i = 0; i < njobs; i++) {
    const hmb200_pu_job& j = jobs[i];
    if (!supported_pu(j.w, j.h) || j.rb_x < j.lt_x || j.rb_y < j.lt_y ||
        (int64_t)(j.rb_x - j.lt_x + 1) * (j.rb_y - j.lt_y + 1) > (int64_t)1 << 24) {
      fail(HMB200_ERR_ARG, "hmb200_prepare_jobs: unsupported PU size or empty/oversized window in job " + std::to_string(i));
      delete p; return nullptr;
    }
    int ss = ((flags & HMB200_FLAG_FEN) && j.h > 8) ? 1 : 0;
    p->tasks[i] = SearchTask{j.pu_x, j.pu_y, j.pu_x, j.pu_y, j.w, j.h, j.lt_x, j.lt_y, j.rb_x, j.rb_y, j.pred_x, j.pred_y, j.lambda_cost, ss};
    uint64_t nc = (uint64_t)(j.rb_x - j.lt_x + 1) * (uint64_t)(j.rb_y - j.lt_y + 1);
    p->cand_sads += nc;
    p->abs_diffs += nc * (uint64_t)j.w * (uint64_t)(j.h >> ss);
  }
  if (njobs > 0) {
    if (cudaMalloc(&p->d_tasks, (size_t)njobs * sizeof(SearchTask)) != cudaSuccess ||
        cudaMalloc(&p->d_results, (size_t)njobs * sizeof(hmb200_pu_result)) != cudaSuccess ||
        cudaMemcpyAsync(p->d_tasks, p->tasks.data(), (size_t)njobs * sizeof(SearchTask), cudaMemcpyHostToDevice, g.stream) != cudaSuccess ||
        cudaMemsetAsync(p->d_results, 0, (size_t)njobs * sizeof(hmb200_pu_result), g.stream) != cudaSuccess ||
        cudaStreamSynchronize(g.stream) != cudaSuccess) {
      fail(HMB200_ERR_CUDA, std::string("hmb200_prepare_jobs: ") + cudaGetErrorString(cudaGetLastError()));
      hmb200_free_prepared(p); return nullptr;
    }
    std::string why;
    if (!(flags & HMB200_FLAG_TZ) && bit_depth >= 8 && bit_depth <= 14) {
      // 8-bit: per-PU tiles + CU-fused bundles; 9..14-bit: CU-fused bundles (packed 16x2 arithmetic), the rest generic
      const int bps = bit_depth > 8 ? 2 : 1;
      std::vector<char> bundled;
      std::vector<CuBundleHost> hb;
      if (!getenv("HMB200_NO_CU_FUSION")) cu_extract_bundles(p->tasks, bps, bundled, hb);
      if (!search8_build_schedule(p->tasks, bundled, g.sm_count, g.stream, &p->sched, &why, /*tiled=*/bps == 1) ||
          !cu_build_schedule(p->tasks, hb, bps, bit_depth, g.sm_count, g.stream, &p->cu, &why)) {
        fail(HMB200_ERR_CUDA, why); hmb200_free_prepared(p); return nullptr;
      }
    }
    if ((flags & HMB200_FLAG_FRAC) && !frac_build_schedule(p->tasks, g.stream, &p->frac, &why)) {
      fail(HMB200_ERR_CUDA, why); hmb200_free_prepared(p); return nullptr;
    }
  }
  return p;
}

void hmb200_free_prepared(hmb200_prepared* p) {
  if (!p) return;
  if (g.ready) { cudaStreamSynchronize(g.stream); if (g.copy) cudaStreamSynchronize(g.copy); }
  if (p->ev_done) cudaEventDestroy(p->ev_done);
  if (p->ev_fetched) cudaEventDestroy(p->ev_fetched);
  if (p->d_tasks) cudaFree(p->d_tasks);
  if (p->d_results) cudaFree(p->d_results);
  search8_free_schedule(&p->sched);
  cu_free_schedule(&p->cu);
  frac_free_schedule(&p->frac);
  if (p->d_tz) cudaFree(p->d_tz);
  delete p;
}

int hmb200_prepared_work(const hmb200_prepared* p, uint64_t* cand_sads, uint64_t* abs_diffs) {
  if (!p) return fail(HMB200_ERR_ARG, "null handle");
  if (cand_sads) *cand_sads = p->cand_sads;
  if (abs_diffs) *abs_diffs = p->abs_diffs;
  return HMB200_OK;
}

void hmb200_canonical_tz_extra(const hmb200_pu_job* jobs, int njobs, hmb200_tz_extra* extra) {
  for (int i = 0; i < njobs; i++) {
    const int s = std::max(jobs[i].w, jobs[i].h);
    hmb200_tz_extra e{};
    e.cu_x = jobs[i].pu_x - jobs[i].pu_x % s;
    e.cu_y = jobs[i].pu_y - jobs[i].pu_y % s;
    extra[i] = e;
  }
}

int hmb200_prepared_set_tz(hmb200_prepared* p, const hmb200_tz_extra* extra, int pic_w, int pic_h, int max_cu, int search_range) {
  NEED_READY();
  if (!p || !(p->flags & HMB200_FLAG_TZ)) return fail(HMB200_ERR_ARG, "hmb200_prepared_set_tz: the list was not prepared with HMB200_FLAG_TZ");
  if ((p->n > 0 && !extra) || pic_w <= 0 || pic_h <= 0 || max_cu <= 0 || search_range < 1 || search_range > 4096)
    return fail(HMB200_ERR_ARG, "hmb200_prepared_set_tz: bad arguments");
  if (p->n > 0) {
    if (!p->d_tz) CUDA_TRY(cudaMalloc((void**)&p->d_tz, (size_t)p->n * sizeof(hmb200_tz_extra)));
    CUDA_TRY(cudaMemcpyAsync(p->d_tz, extra, (size_t)p->n * sizeof(hmb200_tz_extra), cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
  }
  p->tz = TzParams{pic_w, pic_h, max_cu, search_range, (p->flags & HMB200_FLAG_TZ_STOP) ? 1 : 0};
  return HMB200_OK;
}

int hmb200_tz_jobs(int cur_plane, int ref_plane, const hmb200_pu_job* jobs, const hmb200_tz_extra* extra, int njobs,
                   int pic_w, int pic_h, int max_cu, int search_range, int flags, hmb200_pu_result* results) {
  NEED_READY();
  Plane* pr = get_plane(ref_plane);
  if (!pr) return fail(HMB200_ERR_ARG, "hmb200_tz_jobs: unknown reference plane");
  hmb200_prepared* p = hmb200_prepare_jobs(jobs, njobs, flags | HMB200_FLAG_TZ, pr->d.bit_depth);
  if (!p) return g_err_code != HMB200_OK ? g_err_code : HMB200_ERR_ARG;
  int rc = hmb200_prepared_set_tz(p, extra, pic_w, pic_h, max_cu, search_range);
  if (rc == HMB200_OK) rc = hmb200_run_prepared(p, cur_plane, ref_plane);
  if (rc == HMB200_OK) rc = hmb200_fetch_results(p, results);
  hmb200_free_prepared(p);
  return rc;
}

int hmb200_prepared_executed_work(const hmb200_prepared* p, uint64_t* abs_diffs_executed, uint64_t* pus_fused) {
  if (!p) return fail(HMB200_ERR_ARG, "null handle");
  // leftover (generic-kernel) PUs execute exactly their algorithmic work; they are not counted here separately
  if (abs_diffs_executed) *abs_diffs_executed = p->sched.executed_abs_diffs + p->cu.executed_abs_diffs;
  if (pus_fused) *pus_fused = p->cu.fused_tasks;
  return HMB200_OK;
}

int hmb200_run_prepared(hmb200_prepared* p, int cur_plane, int ref_plane) {
  NEED_READY();
  if (!p) return fail(HMB200_ERR_ARG, "null handle");
  Plane* pc = get_plane(cur_plane); Plane* pr = get_plane(ref_plane);
  if (!pc || !pr) return fail(HMB200_ERR_ARG, "hmb200_run_prepared: unknown plane");
  if (pr->d.bit_depth != p->bit_depth) return fail(HMB200_ERR_ARG, "hmb200_run_prepared: bit depth differs from prepare_jobs");
  if (p->n == 0) return HMB200_OK;
  if (p->fetch_pending) CUDA_TRY(cudaStreamWaitEvent(g.stream, p->ev_fetched, 0));     // the results of the previous run are still being copied out
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  const Search8Schedule& sc = p->sched;
  const CuSchedule& cu = p->cu;
  const int bps = pr->d.bytes_per_sample;
  if (p->flags & HMB200_FLAG_TZ) {
    if (!p->d_tz) return fail(HMB200_ERR_STATE, "hmb200_run_prepared: HMB200_FLAG_TZ list without hmb200_prepared_set_tz");
    if (pc->d.bytes_per_sample != bps) return fail(HMB200_ERR_ARG, "hmb200_run_prepared: TZ search needs planes of one sample size");
    const int blocks = (p->n + TZ_WARPS - 1) / TZ_WARPS;
    if (bps == 1) k_tz_search<uint8_t, uint8_t><<<blocks, TZ_WARPS * 32, 0, g.stream>>>(p->d_tasks, p->d_tz, p->d_results, p->n, pc->d, pr->d, p->tz);
    else          k_tz_search<int16_t, int16_t><<<blocks, TZ_WARPS * 32, 0, g.stream>>>(p->d_tasks, p->d_tz, p->d_results, p->n, pc->d, pr->d, p->tz);
    g.launches++;
  }
  bool fast = !(p->flags & HMB200_FLAG_TZ) && pc->d.bytes_per_sample == bps && ((bps == 1 && (sc.n_jobs > 0 || cu.n_bundles > 0)) || (bps == 2 && cu.n_bundles > 0)) &&
              cu.bps == bps && pc->d.margin_x % 16 == 0 && pr->d.margin_x % 16 == 0 && sc.d_keys != nullptr;
  if (fast) {
    // every staged byte must lie inside the padded buffers (the reference would read outside its planes too)
    auto inside = [](const DevPlane& d, int x0, int y0, int x1, int y1) {
      return x0 + d.margin_x >= 0 && x1 + d.margin_x <= d.pitch && y0 + d.margin_y >= 0 && y1 + d.margin_y <= d.height + 2 * d.margin_y;
    };
    if ((sc.n_jobs > 0 && !inside(pr->d, sc.min_x, sc.min_y, sc.max_x, sc.max_y)) ||
        (cu.n_bundles > 0 && !inside(pr->d, cu.rbox.x0, cu.rbox.y0, cu.rbox.x1, cu.rbox.y1)))
      return fail(HMB200_ERR_ARG, "hmb200_run_prepared: a search window leaves the padded reference plane");
    if ((sc.n_jobs > 0 && !inside(pc->d, sc.omin_x, sc.omin_y, sc.omax_x, sc.omax_y)) ||
        (cu.n_bundles > 0 && !inside(pc->d, cu.obox.x0, cu.obox.y0, cu.obox.x1, cu.obox.y1)))
      return fail(HMB200_ERR_ARG, "hmb200_run_prepared: a PU leaves the padded current plane");
    CUDA_TRY(cudaMemsetAsync(sc.d_keys, 0xff, (size_t)sc.n_tasks * sizeof(unsigned long long), g.stream));
    // one launch per tile variant present, spread over side streams so that their tails overlap
    const S8Kernel* kern = search8_kernels();
    const S8CuKernel* cukern = bps == 1 ? search8_cu_kernels() : search16_cu_kernels();
    CUDA_TRY(cudaEventRecord(g.ev_fork, g.stream));
    int order[S8V_COUNT + CUV_COUNT], used = 0;       // >= 0: per-PU variant, < 0: CU-fused variant ~v
    if (bps == 1) { for (int v = 0; v < CUV_COUNT; v++) if (cu.unit_count[v] > 0) order[used++] = ~v; }       // longest CTAs first: 8x8 CUs
    else for (int v = CUV_COUNT - 1; v >= 0; v--) if (cu.unit_count[v] > 0) order[used++] = ~v;
    for (int v = S8V_COUNT - 1; v >= 0; v--) if (sc.unit_count[v] > 0) order[used++] = v;     // wide tiles first
    for (int k = 0; k < used; k++) {
      const int v = order[k];
      cudaStream_t st = (k == 0) ? g.stream : g.side[(k - 1) % N_SIDE];
      if (k >= 1 && k <= N_SIDE) CUDA_TRY(cudaStreamWaitEvent(st, g.ev_fork, 0));
      if (v >= 0)
        kern[v]<<<sc.unit_count[v], S8_THREADS, sc.smem_of[v], st>>>(sc.d_units + sc.unit_first[v], sc.d_jobs, sc.d_keys, pc->d, pr->d);
      else
        cukern[~v]<<<cu.unit_count[~v], bps == 1 ? CU8_THREADS : S8_THREADS, cu.smem_of[~v], st>>>(cu.d_units + cu.unit_first[~v], cu.d_bundles, sc.d_keys, pc->d, pr->d);
      g.launches++;
    }
    for (int k = 0; k < N_SIDE && k + 1 < used; k++) {
      CUDA_TRY(cudaEventRecord(g.ev_join[k], g.side[k]));
      CUDA_TRY(cudaStreamWaitEvent(g.stream, g.ev_join[k], 0));
    }
    if (sc.n_leftover > 0) {    // shapes / windows the tiled kernels do not cover
      if (bps == 1) k_search_generic<uint8_t, uint8_t><<<sc.n_leftover, 256, 0, g.stream>>>(p->d_tasks, p->d_results, pc->d, pr->d, sc.d_leftover);
      else          k_search_generic<int16_t, int16_t><<<sc.n_leftover, 256, 0, g.stream>>>(p->d_tasks, p->d_results, pc->d, pr->d, sc.d_leftover);
      g.launches++;
    }
    k_search8_finalize<<<(sc.n_tasks + 255) / 256, 256, 0, g.stream>>>(p->d_tasks, sc.d_keys, p->d_results, sc.n_tasks);
    g.launches++;
  } else if (!(p->flags & HMB200_FLAG_TZ)) {
    dispatch_generic(p->d_tasks, p->d_results, p->n, pc->d, pr->d, p->flags & ~HMB200_FLAG_FRAC, /*do_search=*/true, nullptr);
  }
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  if (p->flags & HMB200_FLAG_FRAC) {
    const bool had = (p->flags & HMB200_FLAG_HADME) != 0;
    int nl;
    if (pc->d.bytes_per_sample == 1 && pr->d.bytes_per_sample == 1)
      nl = frac_launch<uint8_t, uint8_t>(p->frac, p->d_tasks, p->d_results, pc->d, pr->d, had, g.stream);
    else if (pc->d.bytes_per_sample == 2 && pr->d.bytes_per_sample == 2)
      nl = frac_launch<int16_t, int16_t>(p->frac, p->d_tasks, p->d_results, pc->d, pr->d, had, g.stream);
    else { dispatch_generic(p->d_tasks, p->d_results, p->n, pc->d, pr->d, p->flags, /*do_search=*/false, nullptr); nl = 0; }
    if (nl < 0) return fail(HMB200_ERR_CUDA, std::string("frac_launch: ") + cudaGetErrorString(cudaGetLastError()));
    g.launches += (uint64_t)nl;
  }
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  if (!p->ev_done) CUDA_TRY(cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(p->ev_done, g.stream));
  CUDA_TRY(cudaGetLastError());
  return HMB200_OK;
}

int hmb200_last_timing(float* total_ms, float* search_ms, float* frac_ms) {
  NEED_READY();
  CUDA_TRY(cudaEventSynchronize(g.ev[2]));
  CUDA_TRY(cudaEventElapsedTime(&g.last_total_ms, g.ev[0], g.ev[2]));
  CUDA_TRY(cudaEventElapsedTime(&g.last_search_ms, g.ev[0], g.ev[1]));
  CUDA_TRY(cudaEventElapsedTime(&g.last_frac_ms, g.ev[1], g.ev[2]));
  if (total_ms) *total_ms = g.last_total_ms;
  if (search_ms) *search_ms = g.last_search_ms;
  if (frac_ms) *frac_ms = g.last_frac_ms;
  return HMB200_OK;
}

int hmb200_fetch_results_async(hmb200_prepared* p, hmb200_pu_result* results) {
  NEED_READY();
  if (!p || (p->n > 0 && !results)) return fail(HMB200_ERR_ARG, "hmb200_fetch_results_async: bad arguments");
  if (p->n == 0) return HMB200_OK;
  if (!p->ev_done) return fail(HMB200_ERR_STATE, "hmb200_fetch_results_async: nothing has run on this handle");
  cudaPointerAttributes attr;
  const bool pinned = cudaPointerGetAttributes(&attr, results) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (!pinned) return fail(HMB200_ERR_ARG, "hmb200_fetch_results_async: results must come from hmb200_host_alloc (page-locked)");
  if (!p->ev_fetched) CUDA_TRY(cudaEventCreateWithFlags(&p->ev_fetched, cudaEventDisableTiming));
  CUDA_TRY(cudaStreamWaitEvent(g.copy, p->ev_done, 0));
  CUDA_TRY(cudaMemcpyAsync(results, p->d_results, (size_t)p->n * sizeof(hmb200_pu_result), cudaMemcpyDeviceToHost, g.copy));
  CUDA_TRY(cudaEventRecord(p->ev_fetched, g.copy));
  p->fetch_pending = true;
  return HMB200_OK;
}

int hmb200_fetch_wait(hmb200_prepared* p) {
  NEED_READY();
  if (!p) return fail(HMB200_ERR_ARG, "null handle");
  if (p->fetch_pending) { CUDA_TRY(cudaEventSynchronize(p->ev_fetched)); p->fetch_pending = false; }
  return HMB200_OK;
}

int hmb200_fetch_results(hmb200_prepared* p, hmb200_pu_result* results) {
  NEED_READY();
  if (!p || (p->n > 0 && !results)) return fail(HMB200_ERR_ARG, "hmb200_fetch_results: bad arguments");
  if (p->n == 0) return HMB200_OK;
  if (p->fetch_pending) { CUDA_TRY(cudaEventSynchronize(p->ev_fetched)); p->fetch_pending = false; }
  const size_t bytes = (size_t)p->n * sizeof(hmb200_pu_result);
  cudaPointerAttributes attr;
  const bool user_pinned = cudaPointerGetAttributes(&attr, results) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();                                                                              // clear the "not registered" status
  if (user_pinned) {                                                                               // hmb200_host_alloc'ed: straight D2H
    CUDA_TRY(cudaMemcpyAsync(results, p->d_results, bytes, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    CUDA_TRY(cudaGetLastError());
    return HMB200_OK;
  }
  int rc = ensure_pinned(bytes);
  if (rc != HMB200_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(g.pinned, p->d_results, bytes, cudaMemcpyDeviceToHost, g.stream));    // pinned staging: full PCIe rate
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  memcpy(results, g.pinned, bytes);
  return HMB200_OK;
}

int hmb200_me_jobs(int cur_plane, int ref_plane, const hmb200_pu_job* jobs, int njobs, int flags, hmb200_pu_result* results) {
  NEED_READY();
  Plane* pr = get_plane(ref_plane);
  if (!pr) return fail(HMB200_ERR_ARG, "hmb200_me_jobs: unknown reference plane");
  hmb200_prepared* p = hmb200_prepare_jobs(jobs, njobs, flags, pr->d.bit_depth);
  if (!p) return g_err_code != HMB200_OK ? g_err_code : HMB200_ERR_ARG;
  int rc = hmb200_run_prepared(p, cur_plane, ref_plane);
  if (rc == HMB200_OK) rc = hmb200_fetch_results(p, results);
  hmb200_free_prepared(p);
  return rc;
}

int hmb200_me_ctu_row(int cur_plane, int ref_plane, int ctu_row, int max_cu, const hmb200_pu_job* jobs, int njobs, int flags,
                      hmb200_pu_result* results) {
  NEED_READY();
  if (max_cu <= 0) return fail(HMB200_ERR_ARG, "hmb200_me_ctu_row: max_cu must be positive");
  for (int i = 0; i < njobs; i++)
    if (jobs[i].pu_y / max_cu != ctu_row) return fail(HMB200_ERR_ARG, "hmb200_me_ctu_row: job outside the CTU row");
  return hmb200_me_jobs(cur_plane, ref_plane, jobs, njobs, flags, results);
}

// ------------------------------------------------------------------------------------------------------------------
// 1:1 entries (per call, synchronous): exact inside the real encoder where predictors arrive one PU at a time
// ------------------------------------------------------------------------------------------------------------------
static int upload_pattern(const hmb200_pattern* key) {
  if (!key || !key->roi || !supported_pu(key->width, key->height)) return fail(HMB200_ERR_ARG, "unsupported pattern");
  int rc = ensure_pinned(64 * 64 * sizeof(int16_t));
  if (rc != HMB200_OK) return rc;
  int16_t* pin = reinterpret_cast<int16_t*>(g.pinned);
  for (int y = 0; y < key->height; y++)
    memcpy(pin + (size_t)y * key->width, key->roi + (ptrdiff_t)y * key->stride, (size_t)key->width * sizeof(int16_t));
  CUDA_TRY(cudaMemcpy2DAsync(g.pattern.d.base, (size_t)g.pattern.d.pitch * sizeof(int16_t), pin, (size_t)key->width * sizeof(int16_t),
                             (size_t)key->width * sizeof(int16_t), key->height, cudaMemcpyHostToDevice, g.stream));
  return HMB200_OK;
}

static int run_single(const hmb200_pattern* key, const int16_t* ref_at_pu, const SearchTask& proto, int flags, bool do_search,
                      hmb200_pu_result* io) {
  int rx, ry;
  Plane* pr = find_plane_by_host(ref_at_pu, &rx, &ry);
  if (!pr) return fail(HMB200_ERR_ARG, "reference pointer does not fall into a registered plane");
  if (pr->d.bit_depth != key->bit_depth) return fail(HMB200_ERR_ARG, "pattern bit depth differs from the reference plane");
  int rc = upload_pattern(key);
  if (rc != HMB200_OK) return rc;
  SearchTask t = proto;
  t.org_x = 0; t.org_y = 0; t.ref_x = rx; t.ref_y = ry; t.w = key->width; t.h = key->height;
  size_t need = sizeof(SearchTask) + sizeof(hmb200_pu_result) + 128;
  if ((rc = ensure_dstage(need)) != HMB200_OK) return rc;
  SearchTask* d_t = reinterpret_cast<SearchTask*>(g.dstage);
  hmb200_pu_result* d_r = reinterpret_cast<hmb200_pu_result*>(reinterpret_cast<char*>(g.dstage) + 64);
  unsigned long long* d_key = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(g.dstage) + 64 + sizeof(hmb200_pu_result));
  CUDA_TRY(cudaMemcpyAsync(d_t, &t, sizeof(t), cudaMemcpyHostToDevice, g.stream));
  CUDA_TRY(cudaMemcpyAsync(d_r, io, sizeof(*io), cudaMemcpyHostToDevice, g.stream));
  if (do_search) {
    // one PU: spread its candidates over the whole GPU (one CTA per ~256 candidates), fold with atomicMin, decode
    const long long total = (long long)(t.rb_x - t.lt_x + 1) * (t.rb_y - t.lt_y + 1);
    const int splits = (int)std::max<long long>(1, std::min<long long>(4 * g.sm_count, (total + 255) / 256));
    CUDA_TRY(cudaMemsetAsync(d_key, 0xff, sizeof(unsigned long long), g.stream));
    const dim3 grid(1, splits);
    if (pr->d.bytes_per_sample == 1) k_search_split<uint8_t, int16_t><<<grid, 256, 0, g.stream>>>(d_t, d_key, g.pattern.d, pr->d);
    else                             k_search_split<int16_t, int16_t><<<grid, 256, 0, g.stream>>>(d_t, d_key, g.pattern.d, pr->d);
    k_search8_finalize<<<1, 32, 0, g.stream>>>(d_t, d_key, d_r, 1);
    g.launches += 2;
  }
  dispatch_generic(d_t, d_r, 1, g.pattern.d, pr->d, flags, /*do_search=*/false, nullptr);
  CUDA_TRY(cudaMemcpyAsync(io, d_r, sizeof(*io), cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  return HMB200_OK;
}

int hmb200_pattern_search(const hmb200_pattern* key, const int16_t* ref_at_pu, int ref_stride, hmb200_mv lt, hmb200_mv rb,
                          const hmb200_cost_state* cs, int flags, hmb200_mv* mv_out, uint32_t* sad_out) {
  NEED_READY();
  (void)ref_stride;
  if (!key || !cs || !mv_out || !sad_out || rb.x < lt.x || rb.y < lt.y) return fail(HMB200_ERR_ARG, "hmb200_pattern_search: bad arguments");
  SearchTask t{};
  t.lt_x = lt.x; t.lt_y = lt.y; t.rb_x = rb.x; t.rb_y = rb.y; t.pred_x = cs->pred.x; t.pred_y = cs->pred.y;
  t.lambda_cost = cs->lambda_cost;
  t.sub_shift = ((flags & HMB200_FLAG_FEN) && key->height > 8) ? 1 : 0;
  hmb200_pu_result r{};
  int rc = run_single(key, ref_at_pu, t, flags & ~HMB200_FLAG_FRAC, true, &r);
  if (rc != HMB200_OK) return rc;
  mv_out->x = r.mv_x; mv_out->y = r.mv_y; *sad_out = r.sad;
  return HMB200_OK;
}

int hmb200_pattern_search_tz(const hmb200_pattern* key, const int16_t* ref_at_pu, int ref_stride, hmb200_mv lt, hmb200_mv rb,
                             const hmb200_cost_state* cs, int flags, const hmb200_tz_extra* extra, int pic_w, int pic_h, int max_cu,
                             int search_range, hmb200_mv* mv_out, uint32_t* sad_out) {
  NEED_READY();
  (void)ref_stride;
  if (!key || !cs || !extra || !mv_out || !sad_out || rb.x < lt.x || rb.y < lt.y || search_range < 1)
    return fail(HMB200_ERR_ARG, "hmb200_pattern_search_tz: bad arguments");
  int rx, ry;
  Plane* pr = find_plane_by_host(ref_at_pu, &rx, &ry);
  if (!pr) return fail(HMB200_ERR_ARG, "reference pointer does not fall into a registered plane");
  if (pr->d.bit_depth != key->bit_depth) return fail(HMB200_ERR_ARG, "pattern bit depth differs from the reference plane");
  int rc = upload_pattern(key);
  if (rc != HMB200_OK) return rc;
  SearchTask t{};
  t.org_x = 0; t.org_y = 0; t.ref_x = rx; t.ref_y = ry; t.w = key->width; t.h = key->height;
  t.lt_x = lt.x; t.lt_y = lt.y; t.rb_x = rb.x; t.rb_y = rb.y; t.pred_x = cs->pred.x; t.pred_y = cs->pred.y;
  t.lambda_cost = cs->lambda_cost;
  t.sub_shift = ((flags & HMB200_FLAG_FEN) && key->height > 8) ? 1 : 0;
  // one staging record: [SearchTask | tz_extra | result]
  struct Rec { SearchTask t; hmb200_tz_extra e; hmb200_pu_result r; } rec{t, *extra, hmb200_pu_result{}};
  if ((rc = ensure_dstage(sizeof(Rec) + 64)) != HMB200_OK) return rc;
  Rec* d = reinterpret_cast<Rec*>(g.dstage);
  CUDA_TRY(cudaMemcpyAsync(d, &rec, sizeof(rec), cudaMemcpyHostToDevice, g.stream));
  const TzParams P{pic_w, pic_h, max_cu, search_range, (flags & HMB200_FLAG_TZ_STOP) ? 1 : 0};
  if (pr->d.bytes_per_sample == 1) k_tz_search<uint8_t, int16_t><<<1, TZ_WARPS * 32, 0, g.stream>>>(&d->t, &d->e, &d->r, 1, g.pattern.d, pr->d, P);
  else                             k_tz_search<int16_t, int16_t><<<1, TZ_WARPS * 32, 0, g.stream>>>(&d->t, &d->e, &d->r, 1, g.pattern.d, pr->d, P);
  g.launches++;
  hmb200_pu_result r{};
  CUDA_TRY(cudaMemcpyAsync(&r, &d->r, sizeof(r), cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaGetLastError());
  mv_out->x = r.mv_x; mv_out->y = r.mv_y; *sad_out = r.sad;
  return HMB200_OK;
}

int hmb200_pattern_search_frac(int lossless, const hmb200_pattern* key, const int16_t* ref_at_pu, int ref_stride, hmb200_mv mv_int,
                               const hmb200_cost_state* cs, int flags, hmb200_mv* half_out, hmb200_mv* qter_out, uint32_t* cost_out) {
  NEED_READY();
  (void)ref_stride;
  if (!key || !cs || !half_out || !qter_out || !cost_out) return fail(HMB200_ERR_ARG, "hmb200_pattern_search_frac: bad arguments");
  SearchTask t{};
  t.pred_x = cs->pred.x; t.pred_y = cs->pred.y; t.lambda_cost = cs->lambda_cost;
  hmb200_pu_result r{};
  r.mv_x = mv_int.x; r.mv_y = mv_int.y;
  int f = HMB200_FLAG_FRAC | (((flags & HMB200_FLAG_HADME) && !lossless) ? HMB200_FLAG_HADME : 0);
  int rc = run_single(key, ref_at_pu, t, f, false, &r);
  if (rc != HMB200_OK) return rc;
  half_out->x = r.half_x; half_out->y = r.half_y; qter_out->x = r.qter_x; qter_out->y = r.qter_y; *cost_out = r.frac_cost;
  return HMB200_OK;
}

} // extern "C"
