// C-ABI frontend of libhmb200.so (include/hmb200.h): plane registry, job scheduling, kernel launches.
// One caller thread per CONTEXT (the reference path is non-reentrant too: shared m_cDistParam / m_filteredBlock,
// TLibEncoder/TEncSearch.h:113); several contexts - one per GPU - may run concurrently on their own threads.
// No CPU fallback anywhere: without a usable sm_100 device every compute entry
// fails with HMB200_ERR_CUDA.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>
#include <cuda_runtime.h>
#include "hmb200_device.cuh"
#include "hmb200_generic.cuh"
#include "hmb200_search8.cuh"
#include "hmb200_search8_cu.cuh"
#include "hmb200_search16_cu.cuh"
#include "hmb200_intra.cuh"
#include "hmb200_frac.cuh"
#include "hmb200_tz.cuh"
#include "hmb200_mc_cand.cuh"
#include "hmb200_one.cuh"
#include <atomic>

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
// layout of the per-call record of the 1:1 entries (host page-locked copy and device copy): offsets in bytes
constexpr size_t ONE_KEY = 0, ONE_TICKET = 8, ONE_RESULT = 16, ONE_TASK = 64, ONE_TZ = 128, ONE_HEAD = 256;
// the two 16-byte records the call's last kernel writes into mapped host memory (State::one_back, hmb200_one.cuh OneBack)
constexpr size_t BACK_A = 0, BACK_B = 16, BACK_CU = 64, BACK_BYTES = 16384;
// after the pattern rows of the per-call record: argmin keys and integer results of a speculated CU's partitions (hmb200_one.cuh)
constexpr size_t ONE_CU_KEYS = ONE_HEAD + 64 * 64 * sizeof(int16_t), ONE_CU_OUT = ONE_CU_KEYS + 512, ONE_DEV_BYTES = ONE_CU_OUT + ONE_CU_MAX_PUS * 32;
constexpr int CU_CACHE_ENTRIES = 8;                    // one per reference picture a PU loop walks over
static_assert(BACK_CU + CU_CACHE_ENTRIES * ONE_CU_MAX_PUS * 32 <= BACK_BYTES && ONE_CU_MAX_PUS * 8 <= 512, "1:1 report / key areas");

// What a speculated CU launch computed (hmb200_one.cuh, k_one_cu_search / k_one_cu_frac): every partition of the CU searched and
// refined with the 2Nx2N call's window, predictor and lambda.  A later 1:1 call is served from it iff plane, position, window,
// predictor, lambda, flags and the pattern SAMPLES are the ones the launch used.
struct CuCacheEntry {
  bool valid = false;
  int plane = -1, rx = 0, ry = 0, S = 0, flags = 0;
  int lt_x = 0, lt_y = 0, rb_x = 0, rb_y = 0, pred_x = 0, pred_y = 0;
  uint32_t lambda = 0, seq = 0;
  bool got[ONE_CU_MAX_PUS] = {false};
  hmb200_pu_result res[ONE_CU_MAX_PUS];
  int16_t pat[64 * 64];                                // the CU's original block, dense rows of S samples
};

struct State {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};     // side streams for concurrent variant kernels
  cudaStream_t copy = nullptr;                                     // D2H of results behind the next frame pair's kernels (hmb200_fetch_results_async)
  cudaStream_t up = nullptr;                                       // H2D + border extension of page-locked frames behind the current pair's kernels
  void* upstage = nullptr; size_t upstage_bytes = 0;               // device staging of the upload stream (G.dstage belongs to the compute stream)
  std::vector<cudaEvent_t> free_events;
  cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<Plane> planes;
  void* pinned = nullptr; size_t pinned_bytes = 0;
  void* dstage = nullptr; size_t dstage_bytes = 0;     // device staging for plane uploads / per-call blocks
  Plane pattern;                                       // 64x64 int16 pattern buffer for the 1:1 entries (inside one_dev)
  // 1:1 entries: ONE page-locked record up (argmin key | result | task | TZ extra | pattern rows) and one result down per call
  char* one_host = nullptr; char* one_dev = nullptr;
  char* one_back = nullptr;                            // mapped page-locked: the call's last kernel stores result + sequence number here
  uint32_t one_seq = 0;
  bool one_fast = true;                                // hmb200_one.cuh kernels + flag spin (HMB200_NO_ONE_FAST=1: round-1 path)
  bool speculate = true;                               // a 2Nx2N call searches the CU's other partitions too (HMB200_NO_SPECULATION=1: off)
  std::vector<CuCacheEntry> cu_cache;
  int cu_cache_next = 0;
  uint64_t one_calls = 0, cu_launches = 0, cu_hits = 0;
  std::vector<PoolBuf> pool;                           // released plane buffers, recycled by size (no malloc/free per frame)
  uint64_t launches = 0;
  uint64_t pool_mallocs = 0, pool_frees = 0, pool_sync_frees = 0;   // HMB200_DEBUG_POOL: printed when the context goes away
  float last_total_ms = 0, last_search_ms = 0, last_frac_ms = 0;
};

// One State per context.  The process-wide default context is what hmb200_init creates (the single-GPU encoder: one
// caller thread, nothing to bind); hmb200_ctx_create makes further ones - one per GPU, each driven by its own host thread
// ("one host thread and stream set per GPU").  Every entry point works on the calling thread's current
// context; a context must not be used by two threads at the same time.
State g_default;
thread_local State* tl_ctx = nullptr;            // nullptr: the default context
inline State& cur_state() { return tl_ctx ? *tl_ctx : g_default; }
#define G cur_state()
thread_local std::string g_err;                  // last error of the calling thread
std::mutex g_init_mutex;                         // context creation / destruction is serialised (kernel attribute tables, device selection)

thread_local int g_err_code = HMB200_OK;
int fail(int code, const std::string& msg) { g_err = msg; g_err_code = code; return code; }
#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(HMB200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                 \
  } while (0)
#define NEED_READY() do { if (!G.ready) return fail(HMB200_ERR_STATE, "hmb200_init has not succeeded"); } while (0)

int ensure_pinned(size_t bytes) {
  if (bytes <= G.pinned_bytes) return HMB200_OK;
  if (G.pinned) cudaFreeHost(G.pinned);
  G.pinned = nullptr; G.pinned_bytes = 0;
  CUDA_TRY(cudaMallocHost(&G.pinned, bytes));
  G.pinned_bytes = bytes;
  return HMB200_OK;
}
int ensure_dstage(size_t bytes) {
  if (bytes <= G.dstage_bytes) return HMB200_OK;
  if (G.dstage) cudaFree(G.dstage);
  G.dstage = nullptr; G.dstage_bytes = 0;
  CUDA_TRY(cudaMalloc(&G.dstage, bytes));
  G.dstage_bytes = bytes;
  return HMB200_OK;
}

cudaEvent_t take_event() {
  if (!G.free_events.empty()) { cudaEvent_t e = G.free_events.back(); G.free_events.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  return e;
}
void give_event(cudaEvent_t e) { if (e) G.free_events.push_back(e); }

int alloc_plane_slot() {
  for (size_t i = 0; i < G.planes.size(); i++) if (!G.planes[i].used) return (int)i;
  G.planes.emplace_back();
  return (int)G.planes.size() - 1;
}

int make_plane(Plane& p, int width, int height, int mx, int my, int bit_depth) {
  for (auto& e : G.cu_cache) e.valid = false;          // a plane id / host range may now mean other samples
  int bps = bit_depth > 8 ? 2 : 1;
  int total_w = width + 2 * mx, total_h = height + 2 * my;
  int pitch_bytes = ((total_w * bps + 127) / 128) * 128;
  p.d.pitch = pitch_bytes / bps;
  p.d.width = width; p.d.height = height; p.d.margin_x = mx; p.d.margin_y = my;
  p.d.bytes_per_sample = bps; p.d.bit_depth = bit_depth;
  p.bytes = (size_t)pitch_bytes * total_h;
  p.d.base = nullptr;
  // Recycle a released buffer of this size.  A buffer whose last reader has finished is taken at once; one that is still being
  // read (its event has not fired) would make this plane's upload wait for that kernel - with two planes per frame pair in
  // circulation every upload would queue behind the previous pair's kernels instead of running next to them - so up to
  // POOL_KEEP busy buffers of a size are left alone and a new one is allocated instead.
  static const int POOL_KEEP = getenv("HMB200_POOL_KEEP") ? std::max(1, atoi(getenv("HMB200_POOL_KEEP"))) : 4;
  int busy = 0, oldest = -1;
  for (size_t i = 0; i < G.pool.size() && !p.d.base; i++) {
    if (G.pool[i].bytes != p.bytes) continue;
    if (!G.pool[i].ev_free || cudaEventQuery(G.pool[i].ev_free) == cudaSuccess) {
      p.d.base = G.pool[i].base; p.ev_reuse = G.pool[i].ev_free; G.pool.erase(G.pool.begin() + i);
    } else {
      if (oldest < 0) oldest = (int)i;
      busy++;
    }
  }
  cudaGetLastError();                                  // cudaErrorNotReady from the queries is not an error
  if (!p.d.base && busy >= POOL_KEEP) {
    p.d.base = G.pool[oldest].base; p.ev_reuse = G.pool[oldest].ev_free; G.pool.erase(G.pool.begin() + oldest);
  }
  if (!p.d.base) {
    // A geometry the pool holds no buffer of (the first plane after a change of picture size / bit depth): this registration
    // allocates anyway, so it is also the moment to drop the idle buffers of other geometries - never on the release path of a
    // running pipeline, where a cudaFree would wait for the kernels in flight.
    bool same = false;
    for (auto& b : G.pool) same = same || b.bytes == p.bytes;
    if (!same)
      for (size_t i = 0; i < G.pool.size();) {
        if (!G.pool[i].ev_free || cudaEventQuery(G.pool[i].ev_free) == cudaSuccess) {
          give_event(G.pool[i].ev_free);
          cudaFree(G.pool[i].base);
          G.pool.erase(G.pool.begin() + i);
          G.pool_frees++;
        } else i++;
      }
    cudaGetLastError();
    CUDA_TRY(cudaMalloc(&p.d.base, p.bytes));
    G.pool_mallocs++;
  }
  p.used = true;
  return HMB200_OK;
}

// Every consumer of a plane on the compute stream goes through here: a plane that was uploaded on the upload stream
// becomes visible to the compute stream (and the side streams forked from it) by one event wait.
Plane* get_plane(int id) {
  if (id < 0 || id >= (int)G.planes.size() || !G.planes[id].used) return nullptr;
  Plane& p = G.planes[id];
  if (p.ev_ready) { cudaStreamWaitEvent(G.stream, p.ev_ready, 0); give_event(p.ev_ready); p.ev_ready = nullptr; }
  return &p;
}

// which registered plane does a host Pel* fall into?  (1:1 entries hand us raw pointers into TComPicYuv buffers)
Plane* find_plane_by_host(const int16_t* ptr, int* x, int* y) {
  for (auto& p : G.planes) {
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

// The PU shapes HM produces (TLibCommon/TComDataCU.cpp:1893-1931: 2Nx2N .. nRx2N of 8..64 CUs, plus 4x4): widths and heights
// of 4, 8, 12, 16, 24, 32, 48, 64.  Other shapes would need the 2x2 Hadamard tiles / generic-width SAD of the distortion
// table (hmb200_dist / hmb200_dist_batch cover those); the search and refinement kernels reject them instead of guessing.
bool pu_dim_ok(int v) { return v == 4 || v == 8 || v == 12 || v == 16 || v == 24 || v == 32 || v == 48 || v == 64; }
bool supported_pu(int w, int h) { return pu_dim_ok(w) && pu_dim_ok(h); }

// Footprints in picture coordinates, [x0, x1) x [y0, y1).  Every launch is preceded by a host-side check that the samples it
// may read lie inside the padded plane: an out-of-plane window would be an out-of-bounds global read (the reference would
// read outside its picture buffer in the same situation; we reject with HMB200_ERR_ARG instead).
struct Box { int x0, y0, x1, y1; };
inline Box box_empty() { return Box{1 << 30, 1 << 30, -(1 << 30), -(1 << 30)}; }
inline void box_add(Box& b, int x0, int y0, int x1, int y1) {
  b.x0 = std::min(b.x0, x0); b.y0 = std::min(b.y0, y0); b.x1 = std::max(b.x1, x1); b.y1 = std::max(b.y1, y1);
}
inline bool box_inside(const DevPlane& d, const Box& b) {
  if (b.x0 > b.x1) return true;                       // empty
  return b.x0 >= -d.margin_x && b.x1 <= d.width + d.margin_x && b.y0 >= -d.margin_y && b.y1 <= d.height + d.margin_y;
}
constexpr int FRAC_REACH = 4;                         // 8-tap filter: 3 samples before, 4 after (TComInterpolationFilter.cpp:57-63)

template <typename RefT, typename OrgT>
void launch_generic(const SearchTask* d_tasks, hmb200_pu_result* d_res, int n, const DevPlane& cur, const DevPlane& ref,
                    int flags, bool do_search, cudaEvent_t mid, const int* d_index = nullptr, int n_index = 0) {
  if (do_search) {
    const int blocks = d_index ? n_index : n;
    if (blocks > 0) {
      k_search_generic<RefT, OrgT><<<blocks, 256, 0, G.stream>>>(d_tasks, d_res, cur, ref, d_index);
      G.launches++;
    }
  }
  if (mid) cudaEventRecord(mid, G.stream);
  if (flags & HMB200_FLAG_FRAC) {
    k_frac_generic<RefT, OrgT><<<n, FRAC_THREADS, FRAC_SMEM_BYTES, G.stream>>>(d_tasks, d_res, cur, ref,
                                                                               (flags & HMB200_FLAG_HADME) ? 1 : 0);
    G.launches++;
  }
}

// 2-D tensor map over a padded 8-bit plane (rows of `pitch` bytes) with a box of box_w bytes x box_h rows, no swizzle, zero fill
// outside: what k_search8_cu's tensor-map staging loads (cp.async.bulk.tensor.2d).  cuTensorMapEncodeTiled is a driver entry
// point; it is resolved through the runtime so that the library does not link libcuda itself.
bool encode_plane_map(const DevPlane& d, int box_w, int box_h, CUtensorMap* out) {
  typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Encode encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return false;
    encode = reinterpret_cast<Encode>(fn);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)d.pitch, (cuuint64_t)(d.height + 2 * d.margin_y)};
  const cuuint64_t gstride[1] = {(cuuint64_t)d.pitch};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
  const cuuint32_t estr[2] = {1, 1};
  return encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d.base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void dispatch_generic(const SearchTask* d_tasks, hmb200_pu_result* d_res, int n, const DevPlane& cur, const DevPlane& ref,
                      int flags, bool do_search, cudaEvent_t mid) {
  if (n <= 0) { if (mid) cudaEventRecord(mid, G.stream); return; }
  bool r8 = ref.bytes_per_sample == 1, o8 = cur.bytes_per_sample == 1;
  if (r8 && o8)        launch_generic<uint8_t, uint8_t>(d_tasks, d_res, n, cur, ref, flags, do_search, mid);
  else if (r8 && !o8)  launch_generic<uint8_t, int16_t>(d_tasks, d_res, n, cur, ref, flags, do_search, mid);
  else if (!r8 && o8)  launch_generic<int16_t, uint8_t>(d_tasks, d_res, n, cur, ref, flags, do_search, mid);
  else                 launch_generic<int16_t, int16_t>(d_tasks, d_res, n, cur, ref, flags, do_search, mid);
}

} // namespace

__global__ void k_pack_results16(const hmb200_pu_result* __restrict__ in, hmb200_pu_result16* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const hmb200_pu_result r = in[i];
  hmb200_pu_result16 o;
  o.mv_x = (int16_t)r.mv_x; o.mv_y = (int16_t)r.mv_y; o.sad = r.sad;
  o.half_x = (int8_t)r.half_x; o.half_y = (int8_t)r.half_y; o.qter_x = (int8_t)r.qter_x; o.qter_y = (int8_t)r.qter_y;
  o.frac_cost = r.frac_cost;
  out[i] = o;
}

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
  hmb200_pu_result16* d_packed = nullptr;   // hmb200_fetch_results16_async: the 16-byte records
  State* owner = nullptr;         // the context the handle was prepared in (its streams, its device)
  Box foot_ref = box_empty(), foot_org = box_empty();    // samples any kernel of this list may read (validated against the planes per run)
};
#define NEED_OWNER(p) do { if ((p)->owner != &G) return fail(HMB200_ERR_STATE, "the prepared handle belongs to another context"); } while (0)

extern "C" {

const char* hmb200_last_error(void) { return g_err.c_str(); }
uint64_t hmb200_launch_count(void) { return G.launches; }

// ---- context lifetime ------------------------------------------------------------------------------------------
static void teardown_state(State& st) {
  // called with the state's device current; safe on a partially initialised state (init failure paths come here too)
  if (st.stream) cudaStreamSynchronize(st.stream);
  if (st.up) cudaStreamSynchronize(st.up);
  if (st.copy) cudaStreamSynchronize(st.copy);
  for (auto& p : st.planes) {
    if (p.used && p.d.base) cudaFree(p.d.base);
    if (p.ev_ready) cudaEventDestroy(p.ev_ready);
    if (p.ev_reuse) cudaEventDestroy(p.ev_reuse);
  }
  st.planes.clear();
  if (getenv("HMB200_DEBUG_POOL"))
    fprintf(stderr, "[hmb200] plane pool: %llu mallocs, %llu frees of other geometries, %llu synchronising frees, %zu buffers held\n",
            (unsigned long long)st.pool_mallocs, (unsigned long long)st.pool_frees, (unsigned long long)st.pool_sync_frees, st.pool.size());
  for (auto& b : st.pool) { cudaFree(b.base); if (b.ev_free) cudaEventDestroy(b.ev_free); }
  st.pool.clear();
  for (auto e : st.free_events) cudaEventDestroy(e);
  st.free_events.clear();
  if (st.upstage) { cudaFree(st.upstage); st.upstage = nullptr; st.upstage_bytes = 0; }
  if (st.one_dev) cudaFree(st.one_dev);
  if (st.one_host) cudaFreeHost(st.one_host);
  if (st.one_back) cudaFreeHost(st.one_back);
  st.one_dev = st.one_host = st.one_back = nullptr;
  st.pattern = Plane();
  if (st.pinned) cudaFreeHost(st.pinned);
  if (st.dstage) cudaFree(st.dstage);
  st.pinned = nullptr; st.pinned_bytes = 0; st.dstage = nullptr; st.dstage_bytes = 0;
  for (auto& ev : st.ev) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
  for (auto& s : st.side) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); s = nullptr; }
  if (st.copy) { cudaStreamDestroy(st.copy); st.copy = nullptr; }
  if (st.up) { cudaStreamDestroy(st.up); st.up = nullptr; }
  if (st.ev_fork) { cudaEventDestroy(st.ev_fork); st.ev_fork = nullptr; }
  for (auto& ev : st.ev_join) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
  if (st.stream) { cudaStreamDestroy(st.stream); st.stream = nullptr; }
  st.ready = false; st.device = -1;
}

// Brings `st` up on `device`.  The caller holds g_init_mutex and has made `st` the calling thread's current context (the
// helpers below allocate through G).
static int init_state_body(State& st, int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(HMB200_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)");
  if (device < 0 || device >= count) return fail(HMB200_ERR_ARG, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(HMB200_ERR_CUDA, "libhmb200 is built for sm_100a only; found " + std::string(prop.name));
  st.sm_count = prop.multiProcessorCount;
  st.device = device;
  CUDA_TRY(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
  for (auto& ev : st.ev) CUDA_TRY(cudaEventCreate(&ev));
  for (auto& s : st.side) CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&st.copy, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&st.up, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&st.ev_fork, cudaEventDisableTiming));
  for (auto& ev : st.ev_join) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  int rc = search8_configure(&g_err);               // kernel attributes are per device: set for every context
  if (rc != HMB200_OK) return rc;
  if ((rc = cu_configure(&g_err)) != HMB200_OK) return rc;
  if ((rc = cu16_configure(&g_err)) != HMB200_OK) return rc;
  st.ready = true;
  // pattern buffer for the 1:1 entries: 64x64 int16, no margins
  st.pattern = Plane();
  CUDA_TRY(cudaMalloc((void**)&st.one_dev, ONE_DEV_BYTES));
  CUDA_TRY(cudaMallocHost((void**)&st.one_host, ONE_HEAD + 64 * 64 * sizeof(int16_t)));
  CUDA_TRY(cudaHostAlloc((void**)&st.one_back, BACK_BYTES, cudaHostAllocMapped));
  memset(st.one_back, 0, BACK_BYTES);
  {
    // argmin key = ~0, ticket = 0: the state k_one_search leaves behind after every call
    unsigned long long init[2] = {~0ull, 0ull};
    CUDA_TRY(cudaMemcpy(st.one_dev + ONE_KEY, init, sizeof(init), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(st.one_dev + ONE_CU_KEYS, 0xff, 512));
  }
  st.cu_cache.assign(CU_CACHE_ENTRIES, CuCacheEntry());
  st.cu_cache_next = 0;
  st.one_calls = st.cu_launches = st.cu_hits = 0;
  CUDA_TRY(cudaFuncSetAttribute(k_one_cu_search<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  CUDA_TRY(cudaFuncSetAttribute(k_one_cu_search<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  CUDA_TRY(cudaFuncSetAttribute(k_one_cu_search_args<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  CUDA_TRY(cudaFuncSetAttribute(k_one_cu_search_args<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  CUDA_TRY(cudaFuncSetAttribute(k_one_cu_search<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  CUDA_TRY(cudaFuncSetAttribute(k_one_cu_search<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  CUDA_TRY(cudaFuncSetAttribute(k_one_cu_search_args<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  CUDA_TRY(cudaFuncSetAttribute(k_one_cu_search_args<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  st.one_seq = 0;
  st.one_fast = getenv("HMB200_NO_ONE_FAST") == nullptr;
  st.speculate = st.one_fast && getenv("HMB200_NO_SPECULATION") == nullptr;
  CUDA_TRY(cudaFuncSetAttribute(k_one_search<true, uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  CUDA_TRY(cudaFuncSetAttribute(k_one_search_args<true, uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  CUDA_TRY(cudaFuncSetAttribute(k_one_search<false, uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  CUDA_TRY(cudaFuncSetAttribute(k_one_search_args<false, uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  CUDA_TRY(cudaFuncSetAttribute(k_one_search<false, int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  CUDA_TRY(cudaFuncSetAttribute(k_one_search_args<false, int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONE_SMEM_MAX));
  st.pattern.used = true;
  st.pattern.d.base = st.one_dev + ONE_HEAD; st.pattern.d.pitch = 64; st.pattern.d.width = 64; st.pattern.d.height = 64;
  st.pattern.d.margin_x = 0; st.pattern.d.margin_y = 0; st.pattern.d.bytes_per_sample = 2; st.pattern.d.bit_depth = 16;
  st.launches = 0;
  return HMB200_OK;
}
static int init_state(State& st, int device) {
  const int rc = init_state_body(st, device);
  if (rc != HMB200_OK) { const std::string why = g_err; teardown_state(st); g_err = why; }   // no half-built context, nothing leaked
  return rc;
}

int hmb200_init(int device) {
  std::lock_guard<std::mutex> lock(g_init_mutex);
  tl_ctx = nullptr;                                  // the default context becomes the calling thread's current one
  if (g_default.ready && g_default.device == device) { cudaSetDevice(device); return HMB200_OK; }
  if (g_default.ready) { cudaSetDevice(g_default.device); teardown_state(g_default); }
  return init_state(g_default, device);
}

void hmb200_shutdown(void) {
  std::lock_guard<std::mutex> lock(g_init_mutex);
  State& st = G;
  if (!st.ready) return;
  cudaSetDevice(st.device);
  teardown_state(st);
}

struct hmb200_ctx { State st; };

hmb200_ctx* hmb200_ctx_create(int device) {
  std::lock_guard<std::mutex> lock(g_init_mutex);
  hmb200_ctx* c = new hmb200_ctx();
  State* prev = tl_ctx;
  tl_ctx = &c->st;
  if (init_state(c->st, device) != HMB200_OK) { tl_ctx = prev; delete c; return nullptr; }
  return c;                                          // current on the calling thread
}

int hmb200_ctx_set_current(hmb200_ctx* ctx) {
  State& st = ctx ? ctx->st : g_default;
  if (!st.ready) return fail(HMB200_ERR_STATE, "hmb200_ctx_set_current: the context is not initialised");
  CUDA_TRY(cudaSetDevice(st.device));
  tl_ctx = ctx ? &ctx->st : nullptr;
  return HMB200_OK;
}

hmb200_ctx* hmb200_ctx_get_current(void) { return tl_ctx ? reinterpret_cast<hmb200_ctx*>(tl_ctx) : nullptr; }

int hmb200_ctx_device(const hmb200_ctx* ctx) { return ctx ? ctx->st.device : g_default.device; }

void hmb200_ctx_destroy(hmb200_ctx* ctx) {
  if (!ctx) return;
  std::lock_guard<std::mutex> lock(g_init_mutex);
  if (ctx->st.ready) {
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->st.device);
    teardown_state(ctx->st);
    if (prev >= 0) cudaSetDevice(prev);
  }
  if (tl_ctx == &ctx->st) tl_ctx = nullptr;
  delete ctx;
}

void* hmb200_host_alloc(size_t bytes) {
  if (!G.ready) { fail(HMB200_ERR_STATE, "hmb200_init has not succeeded"); return nullptr; }
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) { fail(HMB200_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(cudaGetLastError())); return nullptr; }
  return p;
}
void hmb200_host_free(void* p) { if (p) cudaFreeHost(p); }

int hmb200_sync(void) {
  NEED_READY();
  CUDA_TRY(cudaStreamSynchronize(G.up));
  CUDA_TRY(cudaStreamSynchronize(G.stream));
  CUDA_TRY(cudaStreamSynchronize(G.copy));
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
  Plane& p = G.planes[id];
  p = Plane();
  int rc = make_plane(p, width, height, margin_x, margin_y, bit_depth);
  if (rc != HMB200_OK) return rc;
  int total_w = width + 2 * margin_x, total_h = height + 2 * margin_y;
  const int16_t* src = host_origin - (ptrdiff_t)margin_y * stride - margin_x;
  size_t bytes = (size_t)total_w * total_h * sizeof(int16_t);
  if ((rc = ensure_pinned(bytes)) != HMB200_OK || (rc = ensure_dstage(bytes)) != HMB200_OK) { hmb200_release_plane(id); return rc; }
  int16_t* pin = reinterpret_cast<int16_t*>(G.pinned);
  for (int y = 0; y < total_h; y++) memcpy(pin + (size_t)y * total_w, src + (size_t)y * stride, (size_t)total_w * sizeof(int16_t));
  CUDA_TRY(cudaMemcpyAsync(G.dstage, pin, bytes, cudaMemcpyHostToDevice, G.stream));
  dim3 grid((total_w + 255) / 256, total_h);
  if (p.d.bytes_per_sample == 1)
    k_narrow_plane<uint8_t><<<grid, 256, 0, G.stream>>>(reinterpret_cast<const int16_t*>(G.dstage), total_w,
                                                        reinterpret_cast<uint8_t*>(p.d.base), p.d.pitch, total_w, total_h);
  else
    k_narrow_plane<uint16_t><<<grid, 256, 0, G.stream>>>(reinterpret_cast<const int16_t*>(G.dstage), total_w,
                                                         reinterpret_cast<uint16_t*>(p.d.base), p.d.pitch, total_w, total_h);
  G.launches++;
  CUDA_TRY(cudaStreamSynchronize(G.stream));
  p.host_lo = src; p.host_hi = src + (size_t)(total_h - 1) * stride + total_w; p.host_stride = stride;
  p.kind = kind; p.poc = poc;
  return id;
}

// Tightly packed samples without margins (bytes for 8-bit content, 16-bit words for deeper content); the margins are
// synthesised on the device exactly like extendPicBorder.  stride in samples.
static int register_packed(const void* host_samples, int stride, int width, int height, int margin_x, int margin_y, int bit_depth,
                           int kind, int poc, const char* who) {
  NEED_READY();
  if (!host_samples || width <= 0 || height <= 0 || stride < width || margin_x < 0 || margin_y < 0 || bit_depth < 8 || bit_depth > 14)
    return fail(HMB200_ERR_ARG, std::string(who) + ": bad geometry");
  int id = alloc_plane_slot();
  Plane& p = G.planes[id];
  p = Plane();
  int rc = make_plane(p, width, height, margin_x, margin_y, bit_depth);
  if (rc != HMB200_OK) return rc;
  const int bps = p.d.bytes_per_sample;
  const size_t bytes = (size_t)width * height * bps;
  cudaPointerAttributes attr;
  const bool user_pinned = stride == width && cudaPointerGetAttributes(&attr, host_samples) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  auto pad_launch = [&](const void* d_src, cudaStream_t st) {
    dim3 grid((width + 2 * margin_x + 255) / 256, height + 2 * margin_y);
    if (bps == 1) k_pad_plane<uint8_t><<<grid, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(d_src), width, reinterpret_cast<uint8_t*>(p.d.base),
                                                             p.d.pitch, width, height, margin_x, margin_y);
    else          k_pad_plane<uint16_t><<<grid, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(d_src), width, reinterpret_cast<uint16_t*>(p.d.base),
                                                              p.d.pitch, width, height, margin_x, margin_y);
    G.launches++;
  };
  if (user_pinned) {
    // page-locked, tightly packed frame: H2D + border extension on the upload stream, behind whatever the compute stream is
    // doing; the source stays the caller's until the copy has run (hmb200_sync or any later fetch / blocking call orders it)
    auto upload = [&]() -> int {
      if (bytes > G.upstage_bytes) {
        CUDA_TRY(cudaStreamSynchronize(G.up));
        if (G.upstage) cudaFree(G.upstage);
        G.upstage = nullptr; G.upstage_bytes = 0;
        CUDA_TRY(cudaMalloc(&G.upstage, bytes));
        G.upstage_bytes = bytes;
      }
      if (p.ev_reuse) { CUDA_TRY(cudaStreamWaitEvent(G.up, p.ev_reuse, 0)); give_event(p.ev_reuse); p.ev_reuse = nullptr; }
      CUDA_TRY(cudaMemcpyAsync(G.upstage, host_samples, bytes, cudaMemcpyHostToDevice, G.up));
      pad_launch(G.upstage, G.up);
      p.ev_ready = take_event();
      CUDA_TRY(cudaEventRecord(p.ev_ready, G.up));
      return HMB200_OK;
    };
    if ((rc = upload()) != HMB200_OK) { const std::string why = g_err; hmb200_release_plane(id); g_err = why; return rc; }
    p.kind = kind; p.poc = poc;
    return id;
  }
  // pageable (or strided) source: through the library's page-locked staging buffer, which the next call reuses - so this path waits
  if ((rc = ensure_dstage(bytes)) != HMB200_OK || (rc = ensure_pinned(bytes)) != HMB200_OK) { hmb200_release_plane(id); return rc; }
  uint8_t* pin = reinterpret_cast<uint8_t*>(G.pinned);
  for (int y = 0; y < height; y++)
    memcpy(pin + (size_t)y * width * bps, reinterpret_cast<const uint8_t*>(host_samples) + (size_t)y * stride * bps, (size_t)width * bps);
  auto upload = [&]() -> int {
    CUDA_TRY(cudaMemcpyAsync(G.dstage, pin, bytes, cudaMemcpyHostToDevice, G.stream));
    pad_launch(G.dstage, G.stream);
    CUDA_TRY(cudaStreamSynchronize(G.stream));
    return HMB200_OK;
  };
  if ((rc = upload()) != HMB200_OK) { const std::string why = g_err; hmb200_release_plane(id); g_err = why; return rc; }
  p.kind = kind; p.poc = poc;
  return id;
}

int hmb200_register_plane_u8(const uint8_t* host_samples, int stride, int width, int height, int margin_x, int margin_y,
                             int kind, int poc) {
  return register_packed(host_samples, stride, width, height, margin_x, margin_y, 8, kind, poc, "hmb200_register_plane_u8");
}

int hmb200_register_plane_u16(const uint16_t* host_samples, int stride, int width, int height, int margin_x, int margin_y,
                              int bit_depth, int kind, int poc) {
  if (bit_depth <= 8) return fail(HMB200_ERR_ARG, "hmb200_register_plane_u16: bit_depth must be 9..14 (8-bit content uses hmb200_register_plane_u8)");
  return register_packed(host_samples, stride, width, height, margin_x, margin_y, bit_depth, kind, poc, "hmb200_register_plane_u16");
}

int hmb200_register_plane_yuv(const void* file_luma, int file_is16, int width, int height, int pad_x, int pad_y,
                              int file_bit_depth, int internal_bit_depth, int margin_x, int margin_y, int kind, int poc) {
  NEED_READY();
  if (!file_luma || width <= 0 || height <= 0 || pad_x < 0 || pad_y < 0 || margin_x < 0 || margin_y < 0 || file_bit_depth < 8 ||
      file_bit_depth > 16 || internal_bit_depth < 8 || internal_bit_depth > 14 || (!file_is16 && file_bit_depth > 8))
    return fail(HMB200_ERR_ARG, "hmb200_register_plane_yuv: bad geometry or bit depths");
  int id = alloc_plane_slot();
  Plane& p = G.planes[id];
  p = Plane();
  const int cw = width + pad_x, ch = height + pad_y;
  int rc = make_plane(p, cw, ch, margin_x, margin_y, internal_bit_depth);
  if (rc != HMB200_OK) return rc;
  const size_t bytes = (size_t)width * height * (file_is16 ? 2 : 1);
  if ((rc = ensure_pinned(bytes)) != HMB200_OK || (rc = ensure_dstage(bytes)) != HMB200_OK) { hmb200_release_plane(id); return rc; }
  memcpy(G.pinned, file_luma, bytes);
  CUDA_TRY(cudaMemcpyAsync(G.dstage, G.pinned, bytes, cudaMemcpyHostToDevice, G.stream));
  const int shift = internal_bit_depth - file_bit_depth, maxval = (1 << internal_bit_depth) - 1;
  dim3 grid((cw + 2 * margin_x + 255) / 256, ch + 2 * margin_y);
  if (p.d.bytes_per_sample == 1)
    k_ingest_luma<uint8_t><<<grid, 256, 0, G.stream>>>(reinterpret_cast<const uint8_t*>(G.dstage), file_is16, width, height, pad_x, pad_y,
                                                       shift, maxval, reinterpret_cast<uint8_t*>(p.d.base), p.d.pitch, margin_x, margin_y);
  else
    k_ingest_luma<uint16_t><<<grid, 256, 0, G.stream>>>(reinterpret_cast<const uint8_t*>(G.dstage), file_is16, width, height, pad_x, pad_y,
                                                        shift, maxval, reinterpret_cast<uint16_t*>(p.d.base), p.d.pitch, margin_x, margin_y);
  G.launches++;
  CUDA_TRY(cudaStreamSynchronize(G.stream));
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
    k_widen_plane<uint8_t><<<grid, 256, 0, G.stream>>>(reinterpret_cast<const uint8_t*>(p->d.base), p->d.pitch,
                                                       reinterpret_cast<int16_t*>(G.dstage), total_w, total_w, total_h);
  else
    k_widen_plane<uint16_t><<<grid, 256, 0, G.stream>>>(reinterpret_cast<const uint16_t*>(p->d.base), p->d.pitch,
                                                        reinterpret_cast<int16_t*>(G.dstage), total_w, total_w, total_h);
  G.launches++;
  CUDA_TRY(cudaMemcpyAsync(G.pinned, G.dstage, bytes, cudaMemcpyDeviceToHost, G.stream));
  CUDA_TRY(cudaStreamSynchronize(G.stream));
  const int16_t* pin = reinterpret_cast<const int16_t*>(G.pinned);
  int16_t* dst = dst_origin - (ptrdiff_t)p->d.margin_y * dst_stride - p->d.margin_x;
  for (int y = 0; y < total_h; y++) memcpy(dst + (size_t)y * dst_stride, pin + (size_t)y * total_w, (size_t)total_w * sizeof(int16_t));
  return HMB200_OK;
}

void hmb200_release_plane(int plane_id) {
  for (auto& e : G.cu_cache) e.valid = false;
  Plane* p = get_plane(plane_id);            // also orders a still pending upload of this plane before the compute stream's tail
  if (!p) return;
  if (p->d.base) {
    if (G.pool.size() >= 16) {
      // the pool is full: drop the oldest buffer of ANOTHER size (a geometry that is no longer in use would otherwise occupy the
      // pool for good and turn every release into a synchronise + free, every registration into a malloc)
      for (size_t i = 0; i < G.pool.size(); i++)
        if (G.pool[i].bytes != p->bytes) {
          if (G.pool[i].ev_free) { cudaEventSynchronize(G.pool[i].ev_free); give_event(G.pool[i].ev_free); }
          cudaFree(G.pool[i].base);
          G.pool.erase(G.pool.begin() + i);
          G.pool_frees++;
          break;
        }
    }
    if (G.pool.size() < 16) {
      // no host synchronisation: the buffer goes back to the pool with an event behind the last kernel that may read it
      // (side streams are joined into the compute stream before a run ends); whoever reuses it on the upload stream waits on it
      cudaEvent_t ev = p->ev_reuse ? p->ev_reuse : take_event();
      p->ev_reuse = nullptr;
      cudaEventRecord(ev, G.stream);
      G.pool.push_back(PoolBuf{p->bytes, p->d.base, ev});
    } else {
      cudaStreamSynchronize(G.up);
      cudaStreamSynchronize(G.stream);
      cudaFree(p->d.base);
      G.pool_sync_frees++;
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
  if (!G.ready) die("library not initialised");
  if (!dp || !dp->pOrg || !dp->pCur) die("null DistParam");
  if (dp->bApplyWeight) die("weighted prediction is out of scope (bApplyWeight must be false)");
  if (dp->iStep != 1) die("iStep must be 1");
  int w = dp->iCols, h = dp->iRows;
  if (w <= 0 || h <= 0 || w > 64 || h > 64) die("block size out of range");
  size_t n = (size_t)w * h;
  if (ensure_pinned(2 * n * sizeof(int16_t) + 64) != HMB200_OK || ensure_dstage(2 * n * sizeof(int16_t) + sizeof(DistTask) + 64) != HMB200_OK)
    die("allocation failed");
  int16_t* pin = reinterpret_cast<int16_t*>(G.pinned);
  for (int y = 0; y < h; y++) {
    memcpy(pin + (size_t)y * w, dp->pOrg + (ptrdiff_t)y * dp->iStrideOrg, (size_t)w * sizeof(int16_t));
    memcpy(pin + n + (size_t)y * w, dp->pCur + (ptrdiff_t)y * dp->iStrideCur, (size_t)w * sizeof(int16_t));
  }
  int16_t* d = reinterpret_cast<int16_t*>(G.dstage);
  size_t task_off = ((2 * n * sizeof(int16_t) + 15) / 16) * 16;
  DistTask t{d, d + n, w, w, w, h, dp->iSubShift, 0};
  memcpy(reinterpret_cast<char*>(G.pinned) + task_off, &t, sizeof(t));
  uint32_t* d_out = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(G.dstage) + task_off + sizeof(DistTask));
  if (cudaMemcpyAsync(G.dstage, G.pinned, task_off + sizeof(DistTask), cudaMemcpyHostToDevice, G.stream) != cudaSuccess) die("H2D failed");
  k_dist_generic<int16_t><<<1, 128, 0, G.stream>>>(reinterpret_cast<const DistTask*>(reinterpret_cast<char*>(G.dstage) + task_off),
                                                   d_out, 1, dp->func, dp->bitDepth);
  G.launches++;
  uint32_t out = 0;
  if (cudaMemcpyAsync(&out, d_out, sizeof(out), cudaMemcpyDeviceToHost, G.stream) != cudaSuccess ||
      cudaStreamSynchronize(G.stream) != cudaSuccess) { g_err = cudaGetErrorString(cudaGetLastError()); die("kernel failed"); }
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
    if (!box_inside(po->d, Box{d.org_x, d.org_y, d.org_x + d.w, d.org_y + d.h}) || !box_inside(pc->d, Box{d.cur_x, d.cur_y, d.cur_x + d.w, d.cur_y + d.h}))
      return fail(HMB200_ERR_ARG, "hmb200_dist_batch: block " + std::to_string(i) + " leaves its padded plane");
    auto addr = [&](Plane* p, int x, int y) {
      return reinterpret_cast<char*>(p->d.base) + ((size_t)(y + p->d.margin_y) * p->d.pitch + (x + p->d.margin_x)) * bps;
    };
    tasks[i] = DistTask{addr(po, d.org_x, d.org_y), addr(pc, d.cur_x, d.cur_y), po->d.pitch, pc->d.pitch, d.w, d.h, d.sub_shift, 0};
  }
  size_t tb = tasks.size() * sizeof(DistTask), ob = (size_t)n * sizeof(uint32_t);
  int rc;
  if ((rc = ensure_dstage(tb + ob)) != HMB200_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(G.dstage, tasks.data(), tb, cudaMemcpyHostToDevice, G.stream));
  uint32_t* d_out = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(G.dstage) + tb);
  int blocks = (n + 3) / 4;
  CUDA_TRY(cudaEventRecord(G.ev[0], G.stream));       // hmb200_last_timing: total = the kernel alone (descriptor / result copies outside)
  CUDA_TRY(cudaEventRecord(G.ev[1], G.stream));
  if (bps == 1) k_dist_generic<uint8_t><<<blocks, 128, 0, G.stream>>>(reinterpret_cast<const DistTask*>(G.dstage), d_out, n, func, bit_depth);
  else          k_dist_generic<int16_t><<<blocks, 128, 0, G.stream>>>(reinterpret_cast<const DistTask*>(G.dstage), d_out, n, func, bit_depth);
  CUDA_TRY(cudaEventRecord(G.ev[2], G.stream));
  G.launches++;
  CUDA_TRY(cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, G.stream));
  CUDA_TRY(cudaStreamSynchronize(G.stream));
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
    {
      const int ix = d.pu_x + (d.mv_x >> 2), iy = d.pu_y + (d.mv_y >> 2);
      if (!box_inside(pc->d, Box{d.pu_x, d.pu_y, d.pu_x + d.w, d.pu_y + d.h}) ||
          !box_inside(pr->d, Box{ix - FRAC_REACH, iy - FRAC_REACH, ix + d.w + FRAC_REACH, iy + d.h + FRAC_REACH}))
        return fail(HMB200_ERR_ARG, "hmb200_mc_dist_batch: descriptor " + std::to_string(i) + " (PU or its motion-compensated block with the 8-tap reach) leaves the padded plane");
    }
    tasks[i] = SearchTask{d.pu_x, d.pu_y, d.pu_x, d.pu_y, d.w, d.h, 0, 0, 0, 0, 0, 0, 0u, 0};
    mv[i] = hmb200_pu_result{0, 0, 0u, 0, 0, d.mv_x, d.mv_y, 0u};          // k_frac_tiles stage 2 reads the MV from qter_x / qter_y
  }
  FracSchedule fs;
  std::string why;
  if (!frac_build_schedule(tasks, G.stream, &fs, &why)) return fail(HMB200_ERR_CUDA, why);
  const size_t tb = (size_t)n * sizeof(SearchTask), rb = (size_t)n * sizeof(hmb200_pu_result), ob = (size_t)n * sizeof(uint32_t);
  int rc = ensure_dstage(tb + rb + ob + 256);
  if (rc != HMB200_OK) { frac_free_schedule(&fs); return rc; }
  char* base = reinterpret_cast<char*>(G.dstage);
  SearchTask* d_tasks = reinterpret_cast<SearchTask*>(base);
  hmb200_pu_result* d_mv = reinterpret_cast<hmb200_pu_result*>(base + ((tb + 63) & ~(size_t)63));
  uint32_t* d_out = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(d_mv) + ((rb + 63) & ~(size_t)63));
  cudaError_t e = cudaMemcpyAsync(d_tasks, tasks.data(), tb, cudaMemcpyHostToDevice, G.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_mv, mv.data(), rb, cudaMemcpyHostToDevice, G.stream);
  int nl = -1;
  if (e == cudaSuccess) {
    const bool had = func == HMB200_DF_HADS;
    nl = pr->d.bytes_per_sample == 1 ? mc_launch<uint8_t, uint8_t>(fs, d_tasks, d_mv, d_out, pc->d, pr->d, had, G.stream)
                                     : mc_launch<int16_t, int16_t>(fs, d_tasks, d_mv, d_out, pc->d, pr->d, had, G.stream);
  }
  if (nl >= 0) { G.launches += (uint64_t)nl; e = cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, G.stream); }
  if (e == cudaSuccess) e = cudaStreamSynchronize(G.stream);
  frac_free_schedule(&fs);
  if (nl < 0 || e != cudaSuccess) return fail(HMB200_ERR_CUDA, std::string("hmb200_mc_dist_batch: ") + cudaGetErrorString(cudaGetLastError()));
  return HMB200_OK;
}

// merge / AMVP candidates: uni- or bi-directional prediction + distortion per candidate, candidate loops on the host
static int mc_cand_run(int cur_plane, int func, int n, const hmb200_mc_cand* cands, uint32_t* dist, const char* who) {
  Plane* pc = get_plane(cur_plane);
  if (!pc) return fail(HMB200_ERR_ARG, std::string(who) + ": unknown current plane");
  if (!cands || !dist || (func != HMB200_DF_SAD && func != HMB200_DF_SADS && func != HMB200_DF_HADS))
    return fail(HMB200_ERR_ARG, std::string(who) + ": bad arguments (func must be SAD or HADS)");
  const int bps = pc->d.bytes_per_sample, bd = pc->d.bit_depth;
  std::vector<McCandDev> dev((size_t)n);
  std::vector<uint32_t> t8, t4;
  if (n >= (1 << 24)) return fail(HMB200_ERR_ARG, std::string(who) + ": more than 2^24 candidates in one batch");
  for (int i = 0; i < n; i++) {
    const hmb200_mc_cand& c = cands[i];
    const bool no_check = (c.inter_dir & HMB200_INTER_DIR_NO_IDENTICAL_CHECK) != 0;
    const int dir_in = c.inter_dir & 3;
    if (!supported_pu(c.w, c.h) || dir_in < 1 || (c.inter_dir & ~7) || (no_check && dir_in != 3))
      return fail(HMB200_ERR_ARG, std::string(who) + ": unsupported PU size or inter_dir in candidate " + std::to_string(i));
    if (!box_inside(pc->d, Box{c.pu_x, c.pu_y, c.pu_x + c.w, c.pu_y + c.h}))
      return fail(HMB200_ERR_ARG, std::string(who) + ": PU of candidate " + std::to_string(i) + " leaves the padded current plane");
    McCandDev d{};
    d.org = reinterpret_cast<const char*>(pc->d.base) + ((size_t)(c.pu_y + pc->d.margin_y) * pc->d.pitch + (c.pu_x + pc->d.margin_x)) * bps;
    d.org_pitch = pc->d.pitch; d.w = c.w; d.h = c.h;
    int dir = dir_in;
    if (dir == 3 && !no_check && c.ref0_plane == c.ref1_plane && c.mv0_x == c.mv1_x && c.mv0_y == c.mv1_y) dir = 1;     // xCheckIdenticalMotion
    for (int l = 0; l < 2; l++) {
      if (!(dir & (1 << l))) continue;
      Plane* pr = get_plane(l ? c.ref1_plane : c.ref0_plane);
      const int mvx = l ? c.mv1_x : c.mv0_x, mvy = l ? c.mv1_y : c.mv0_y;
      if (!pr) return fail(HMB200_ERR_ARG, std::string(who) + ": unknown reference plane in candidate " + std::to_string(i));
      if (pr->d.bytes_per_sample != bps || pr->d.bit_depth != bd)
        return fail(HMB200_ERR_ARG, std::string(who) + ": reference and current planes differ in bit depth");
      const int ix = c.pu_x + (mvx >> 2), iy = c.pu_y + (mvy >> 2);
      if (!box_inside(pr->d, Box{ix - FRAC_REACH, iy - FRAC_REACH, ix + c.w + FRAC_REACH, iy + c.h + FRAC_REACH}))
        return fail(HMB200_ERR_ARG, std::string(who) + ": the motion-compensated block of candidate " + std::to_string(i) +
                                        " (with the 8-tap reach) leaves the padded reference plane");
      const void* at = reinterpret_cast<const char*>(pr->d.base) + ((size_t)(c.pu_y + pr->d.margin_y) * pr->d.pitch + (c.pu_x + pr->d.margin_x)) * bps;
      if (l == 0) { d.ref0 = at; d.ref0_pitch = pr->d.pitch; d.mv0_x = mvx; d.mv0_y = mvy; }
      else        { d.ref1 = at; d.ref1_pitch = pr->d.pitch; d.mv1_x = mvx; d.mv1_y = mvy; }
    }
    dev[(size_t)i] = d;
    const int e = (c.w % 8 == 0 && c.h % 8 == 0) ? 8 : 4;                 // xGetHADs tile choice (TComRdCost.cpp:1544-1572); SAD: any tiling
    std::vector<uint32_t>& dst = (e == 8) ? t8 : t4;
    for (int ty = 0; ty < c.h / e; ty++)
      for (int tx = 0; tx < c.w / e; tx++) dst.push_back(frac_pack_tile((uint32_t)i, (uint32_t)tx, (uint32_t)ty));
  }
  auto up64 = [](size_t v) { return (v + 63) & ~(size_t)63; };
  const size_t cb = up64(dev.size() * sizeof(McCandDev)), b8 = up64(t8.size() * 4), b4 = up64(t4.size() * 4), ob = (size_t)n * sizeof(uint32_t);
  int rc = ensure_dstage(cb + b8 + b4 + ob + 256);
  if (rc != HMB200_OK) return rc;
  char* base = reinterpret_cast<char*>(G.dstage);
  McCandDev* d_c = reinterpret_cast<McCandDev*>(base);
  uint32_t* d_t8 = reinterpret_cast<uint32_t*>(base + cb);
  uint32_t* d_t4 = reinterpret_cast<uint32_t*>(base + cb + b8);
  uint32_t* d_out = reinterpret_cast<uint32_t*>(base + cb + b8 + b4);
  CUDA_TRY(cudaMemcpyAsync(d_c, dev.data(), dev.size() * sizeof(McCandDev), cudaMemcpyHostToDevice, G.stream));
  if (!t8.empty()) CUDA_TRY(cudaMemcpyAsync(d_t8, t8.data(), t8.size() * 4, cudaMemcpyHostToDevice, G.stream));
  if (!t4.empty()) CUDA_TRY(cudaMemcpyAsync(d_t4, t4.data(), t4.size() * 4, cudaMemcpyHostToDevice, G.stream));
  CUDA_TRY(cudaMemsetAsync(d_out, 0, ob, G.stream));
  const bool had = func == HMB200_DF_HADS;
  auto launch = [&](auto tag, int N_, const uint32_t* d_t, int nt) {
    typedef decltype(tag) T;
    if (nt == 0) return;
    const int blocks = (nt + 127) / 128;
    if (N_ == 8) { if (had) k_mc_cand_tiles<T, 8, true><<<blocks, 128, 0, G.stream>>>(d_c, d_t, nt, d_out, bd); else k_mc_cand_tiles<T, 8, false><<<blocks, 128, 0, G.stream>>>(d_c, d_t, nt, d_out, bd); }
    else         { if (had) k_mc_cand_tiles<T, 4, true><<<blocks, 128, 0, G.stream>>>(d_c, d_t, nt, d_out, bd); else k_mc_cand_tiles<T, 4, false><<<blocks, 128, 0, G.stream>>>(d_c, d_t, nt, d_out, bd); }
    G.launches++;
  };
  if (bps == 1) { launch(uint8_t(), 8, d_t8, (int)t8.size()); launch(uint8_t(), 4, d_t4, (int)t4.size()); }
  else          { launch(int16_t(), 8, d_t8, (int)t8.size()); launch(int16_t(), 4, d_t4, (int)t4.size()); }
  if (bd > 8) { k_mc_cand_finish<<<(n + 255) / 256, 256, 0, G.stream>>>(d_out, n, bd); G.launches++; }
  CUDA_TRY(cudaMemcpyAsync(dist, d_out, ob, cudaMemcpyDeviceToHost, G.stream));
  CUDA_TRY(cudaStreamSynchronize(G.stream));
  CUDA_TRY(cudaGetLastError());
  return HMB200_OK;
}

int hmb200_mc_cand_dist_batch(int cur_plane, int func, int n, const hmb200_mc_cand* cands, uint32_t* dist) {
  NEED_READY();
  if (n <= 0) return HMB200_OK;
  return mc_cand_run(cur_plane, func, n, cands, dist, "hmb200_mc_cand_dist_batch");
}

static int cand_ranges_ok(int n_pu, const int32_t* cand_first) {
  if (n_pu < 0 || !cand_first || cand_first[0] != 0) return 0;
  for (int i = 0; i < n_pu; i++) if (cand_first[i + 1] < cand_first[i]) return 0;
  return 1;
}

int hmb200_merge_estimation_batch(int cur_plane, int n_pu, const int32_t* cand_first, const hmb200_mc_cand* cands, int use_hadme,
                                  uint32_t lambda_cost, uint32_t* best_cand, uint32_t* best_cost, uint32_t* cand_dist) {
  NEED_READY();
  if (n_pu == 0) return HMB200_OK;
  if (!cand_ranges_ok(n_pu, cand_first) || !best_cand || !best_cost) return fail(HMB200_ERR_ARG, "hmb200_merge_estimation_batch: bad arguments");
  const int n = cand_first[n_pu];
  std::vector<uint32_t> dist((size_t)std::max(n, 1));
  if (n > 0) {
    const int rc = mc_cand_run(cur_plane, use_hadme ? HMB200_DF_HADS : HMB200_DF_SAD, n, cands, dist.data(), "hmb200_merge_estimation_batch");
    if (rc != HMB200_OK) return rc;
  }
  for (int i = 0; i < n_pu; i++) {                     // TEncSearch.cpp:2868-2892
    uint32_t best = 0xffffffffu, bi = 0;
    for (int k = cand_first[i]; k < cand_first[i + 1]; k++) {
      const uint32_t cost = dist[(size_t)k] + ((lambda_cost * (uint32_t)cands[k].bits) >> 16);
      if (cost < best) { best = cost; bi = (uint32_t)(k - cand_first[i]); }
    }
    best_cand[i] = bi; best_cost[i] = best;
  }
  if (cand_dist && n > 0) memcpy(cand_dist, dist.data(), (size_t)n * sizeof(uint32_t));
  return HMB200_OK;
}

int hmb200_amvp_estimation_batch(int cur_plane, int n_pu, const int32_t* cand_first, const hmb200_mc_cand* cands, uint32_t lambda_motion_sad,
                                 uint32_t* best_cand, uint32_t* best_cost, uint32_t* cand_dist) {
  NEED_READY();
  if (n_pu == 0) return HMB200_OK;
  if (!cand_ranges_ok(n_pu, cand_first) || !best_cand || !best_cost) return fail(HMB200_ERR_ARG, "hmb200_amvp_estimation_batch: bad arguments");
  const int n = cand_first[n_pu];
  for (int k = 0; k < n; k++)
    if (cands[k].inter_dir != 1) return fail(HMB200_ERR_ARG, "hmb200_amvp_estimation_batch: AMVP candidates are uni-directional (inter_dir = 1, list-0 fields)");
  std::vector<uint32_t> dist((size_t)std::max(n, 1));
  if (n > 0) {
    const int rc = mc_cand_run(cur_plane, HMB200_DF_SAD, n, cands, dist.data(), "hmb200_amvp_estimation_batch");
    if (rc != HMB200_OK) return rc;
  }
  for (int i = 0; i < n_pu; i++) {                     // TEncSearch.cpp:3457-3469 over xGetTemplateCost
    uint32_t best = 0xffffffffu, bi = 0;
    for (int k = cand_first[i]; k < cand_first[i + 1]; k++) {
      const uint32_t cost = (uint32_t)((uint64_t)dist[(size_t)k] + (((uint64_t)(uint32_t)cands[k].bits * (uint64_t)lambda_motion_sad) >> 16));
      if (best > cost) { best = cost; bi = (uint32_t)(k - cand_first[i]); }
    }
    best_cand[i] = bi; best_cost[i] = best;
  }
  if (cand_dist && n > 0) memcpy(cand_dist, dist.data(), (size_t)n * sizeof(uint32_t));
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
  char* base = reinterpret_cast<char*>(G.dstage);
  IntraBlockDev* d_blocks = reinterpret_cast<IntraBlockDev*>(base);
  int32_t* d_ord = reinterpret_cast<int32_t*>(base + bb);
  int16_t* d_refs = reinterpret_cast<int16_t*>(base + bb + ob);
  void* d_org = base + bb + ob + rb;
  uint32_t* d_out = reinterpret_cast<uint32_t*>(base + bb + ob + rb + ub);
  std::vector<int32_t> ord(ord8);
  ord.insert(ord.end(), ord4.begin(), ord4.end());
  CUDA_TRY(cudaMemcpyAsync(d_blocks, hb.data(), hb.size() * sizeof(IntraBlockDev), cudaMemcpyHostToDevice, G.stream));
  CUDA_TRY(cudaMemcpyAsync(d_ord, ord.data(), ord.size() * sizeof(int32_t), cudaMemcpyHostToDevice, G.stream));
  CUDA_TRY(cudaMemcpyAsync(d_refs, refs, (size_t)n_ref_samples * 2, cudaMemcpyHostToDevice, G.stream));
  DevPlane pl = org;
  if (org_upload) {
    CUDA_TRY(cudaMemcpyAsync(d_org, org_upload, org_upload_bytes, cudaMemcpyHostToDevice, G.stream));
    pl.base = d_org;
  }
  const int bd = pl.bit_depth;
  if (!ord8.empty()) {
    if (pl.bytes_per_sample == 1) k_intra_modes_had<8, uint8_t><<<(int)ord8.size(), INTRA_THREADS, 0, G.stream>>>(d_blocks, d_ord, d_refs, d_out, pl, bd);
    else                          k_intra_modes_had<8, int16_t><<<(int)ord8.size(), INTRA_THREADS, 0, G.stream>>>(d_blocks, d_ord, d_refs, d_out, pl, bd);
    G.launches++;
  }
  if (!ord4.empty()) {
    if (pl.bytes_per_sample == 1) k_intra_modes_had<4, uint8_t><<<(int)ord4.size(), INTRA_THREADS, 0, G.stream>>>(d_blocks, d_ord + ord8.size(), d_refs, d_out, pl, bd);
    else                          k_intra_modes_had<4, int16_t><<<(int)ord4.size(), INTRA_THREADS, 0, G.stream>>>(d_blocks, d_ord + ord8.size(), d_refs, d_out, pl, bd);
    G.launches++;
  }
  CUDA_TRY(cudaMemcpyAsync(out, d_out, outb, cudaMemcpyDeviceToHost, G.stream));
  CUDA_TRY(cudaStreamSynchronize(G.stream));
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
  if (!G.ready) { fail(HMB200_ERR_STATE, "hmb200_init has not succeeded"); return nullptr; }
  if (njobs < 0 || (njobs > 0 && !jobs)) { fail(HMB200_ERR_ARG, "hmb200_prepare_jobs: bad arguments"); return nullptr; }
  auto* p = new hmb200_prepared();
  p->n = njobs; p->flags = flags; p->bit_depth = bit_depth; p->owner = &G;
  const int reach = (flags & HMB200_FLAG_FRAC) ? FRAC_REACH : 0;
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
    box_add(p->foot_org, j.pu_x, j.pu_y, j.pu_x + j.w, j.pu_y + j.h);
    box_add(p->foot_ref, j.pu_x + j.lt_x - reach, j.pu_y + j.lt_y - reach, j.pu_x + j.rb_x + j.w + reach, j.pu_y + j.rb_y + j.h + reach);
  }
  if (njobs > 0) {
    if (cudaMalloc(&p->d_tasks, (size_t)njobs * sizeof(SearchTask)) != cudaSuccess ||
        cudaMalloc(&p->d_results, (size_t)njobs * sizeof(hmb200_pu_result)) != cudaSuccess ||
        cudaMemcpyAsync(p->d_tasks, p->tasks.data(), (size_t)njobs * sizeof(SearchTask), cudaMemcpyHostToDevice, G.stream) != cudaSuccess ||
        cudaMemsetAsync(p->d_results, 0, (size_t)njobs * sizeof(hmb200_pu_result), G.stream) != cudaSuccess ||
        cudaStreamSynchronize(G.stream) != cudaSuccess) {
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
      if (!search8_build_schedule(p->tasks, bundled, G.sm_count, G.stream, &p->sched, &why, /*tiled=*/bps == 1) ||
          !cu_build_schedule(p->tasks, hb, bps, bit_depth, G.sm_count, G.stream, &p->cu, &why)) {
        fail(HMB200_ERR_CUDA, why); hmb200_free_prepared(p); return nullptr;
      }
    }
    if ((flags & HMB200_FLAG_FRAC) && !frac_build_schedule(p->tasks, G.stream, &p->frac, &why)) {
      fail(HMB200_ERR_CUDA, why); hmb200_free_prepared(p); return nullptr;
    }
  }
  return p;
}

void hmb200_free_prepared(hmb200_prepared* p) {
  if (!p) return;
  State& st = p->owner ? *p->owner : G;           // stream waits and frees work from any thread / current device
  if (st.ready) { cudaStreamSynchronize(st.stream); if (st.copy) cudaStreamSynchronize(st.copy); }
  if (p->d_packed) cudaFree(p->d_packed);
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
  NEED_OWNER(p);
  // xTZSearch also visits the zero vector, the clipped predictor / 2Nx2N integer MV and a raster range re-centred on them:
  // all inside TComDataCU::clipMv's range for the owning CU
  const int reach = (p->flags & HMB200_FLAG_FRAC) ? FRAC_REACH : 0;
  for (int i = 0; i < p->n; i++) {
    const SearchTask& t = p->tasks[(size_t)i];
    const int xmin = std::min(0, -max_cu - 8 - extra[i].cu_x + 1), xmax = std::max(0, pic_w + 8 - extra[i].cu_x - 1);
    const int ymin = std::min(0, -max_cu - 8 - extra[i].cu_y + 1), ymax = std::max(0, pic_h + 8 - extra[i].cu_y - 1);
    box_add(p->foot_ref, t.ref_x + xmin - reach, t.ref_y + ymin - reach, t.ref_x + xmax + t.w + reach, t.ref_y + ymax + t.h + reach);
  }
  if (p->n > 0) {
    if (!p->d_tz) CUDA_TRY(cudaMalloc((void**)&p->d_tz, (size_t)p->n * sizeof(hmb200_tz_extra)));
    CUDA_TRY(cudaMemcpyAsync(p->d_tz, extra, (size_t)p->n * sizeof(hmb200_tz_extra), cudaMemcpyHostToDevice, G.stream));
    CUDA_TRY(cudaStreamSynchronize(G.stream));
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

int hmb200_prepared_unique_work(const hmb200_prepared* p, uint64_t* abs_diffs_unique) {
  if (!p) return fail(HMB200_ERR_ARG, "null handle");
  if (abs_diffs_unique) *abs_diffs_unique = p->sched.unique_abs_diffs + p->cu.unique_abs_diffs;
  return HMB200_OK;
}

int hmb200_run_prepared(hmb200_prepared* p, int cur_plane, int ref_plane) {
  NEED_READY();
  if (!p) return fail(HMB200_ERR_ARG, "null handle");
  Plane* pc = get_plane(cur_plane); Plane* pr = get_plane(ref_plane);
  if (!pc || !pr) return fail(HMB200_ERR_ARG, "hmb200_run_prepared: unknown plane");
  if (pr->d.bit_depth != p->bit_depth) return fail(HMB200_ERR_ARG, "hmb200_run_prepared: bit depth differs from prepare_jobs");
  NEED_OWNER(p);
  if (p->n == 0) return HMB200_OK;
  if (!box_inside(pr->d, p->foot_ref)) return fail(HMB200_ERR_ARG, "hmb200_run_prepared: a search window (plus the interpolation reach) leaves the padded reference plane");
  if (!box_inside(pc->d, p->foot_org)) return fail(HMB200_ERR_ARG, "hmb200_run_prepared: a PU leaves the padded current plane");
  if (p->fetch_pending) CUDA_TRY(cudaStreamWaitEvent(G.stream, p->ev_fetched, 0));     // the results of the previous run are still being copied out
  CUDA_TRY(cudaEventRecord(G.ev[0], G.stream));
  const Search8Schedule& sc = p->sched;
  const CuSchedule& cu = p->cu;
  const int bps = pr->d.bytes_per_sample;
  if (p->flags & HMB200_FLAG_TZ) {
    if (!p->d_tz) return fail(HMB200_ERR_STATE, "hmb200_run_prepared: HMB200_FLAG_TZ list without hmb200_prepared_set_tz");
    if (pc->d.bytes_per_sample != bps) return fail(HMB200_ERR_ARG, "hmb200_run_prepared: TZ search needs planes of one sample size");
    const int blocks = (p->n + TZ_WARPS - 1) / TZ_WARPS;
    if (bps == 1) k_tz_search<uint8_t, uint8_t><<<blocks, TZ_WARPS * 32, 0, G.stream>>>(p->d_tasks, p->d_tz, p->d_results, p->n, pc->d, pr->d, p->tz);
    else          k_tz_search<int16_t, int16_t><<<blocks, TZ_WARPS * 32, 0, G.stream>>>(p->d_tasks, p->d_tz, p->d_results, p->n, pc->d, pr->d, p->tz);
    G.launches++;
  }
  bool fast = !(p->flags & HMB200_FLAG_TZ) && pc->d.bytes_per_sample == bps && ((bps == 1 && (sc.n_jobs > 0 || cu.n_bundles > 0)) || (bps == 2 && cu.n_bundles > 0)) &&
              cu.bps == bps && pc->d.margin_x % 16 == 0 && pr->d.margin_x % 16 == 0 && sc.d_keys != nullptr;
  if (fast) {
    // the staged boxes are the footprints rounded out to 16 bytes: the round-up may reach into the row padding (pitch), never past it
    auto inside = [](const DevPlane& d, int x0, int y0, int x1, int y1) {
      return x0 + d.margin_x >= 0 && x1 + d.margin_x <= d.pitch && y0 + d.margin_y >= 0 && y1 + d.margin_y <= d.height + 2 * d.margin_y;
    };
    if ((sc.n_jobs > 0 && !inside(pr->d, sc.min_x, sc.min_y, sc.max_x, sc.max_y)) ||
        (cu.n_bundles > 0 && !inside(pr->d, cu.rbox.x0, cu.rbox.y0, cu.rbox.x1, cu.rbox.y1)))
      return fail(HMB200_ERR_ARG, "hmb200_run_prepared: a search window leaves the padded reference plane");
    if ((sc.n_jobs > 0 && !inside(pc->d, sc.omin_x, sc.omin_y, sc.omax_x, sc.omax_y)) ||
        (cu.n_bundles > 0 && !inside(pc->d, cu.obox.x0, cu.obox.y0, cu.obox.x1, cu.obox.y1)))
      return fail(HMB200_ERR_ARG, "hmb200_run_prepared: a PU leaves the padded current plane");
    CUDA_TRY(cudaMemsetAsync(sc.d_keys, 0xff, (size_t)sc.n_tasks * sizeof(unsigned long long), G.stream));
    // one launch per tile variant present, spread over side streams so that their tails overlap
    const S8Kernel* kern = search8_kernels();
    CUDA_TRY(cudaEventRecord(G.ev_fork, G.stream));
    int order[S8V_COUNT + CUV_MAX], used = 0;       // >= 0: per-PU variant, < 0: CU-fused variant ~v
    if (bps == 1) {                                   // longest CTAs first: 16x16 CUs with their children, then 8x8 CUs
      for (int v = CUV_BASE_COUNT; v < CUV_COUNT; v++) if (cu.unit_count[v] > 0) order[used++] = ~v;
      for (int v = 0; v < CUV_BASE_COUNT; v++) if (cu.unit_count[v] > 0) order[used++] = ~v;
    } else for (int v = CUV16_COUNT - 1; v >= 0; v--) if (cu.unit_count[v] > 0) order[used++] = ~v;
    for (int v = S8V_COUNT - 1; v >= 0; v--) if (sc.unit_count[v] > 0) order[used++] = v;     // wide tiles first
    for (int k = 0; k < used; k++) {
      const int v = order[k];
      cudaStream_t st = (k == 0) ? G.stream : G.side[(k - 1) % N_SIDE];
      if (k >= 1 && k <= N_SIDE) CUDA_TRY(cudaStreamWaitEvent(st, G.ev_fork, 0));
      if (v >= 0)
        kern[v]<<<sc.unit_count[v], S8_THREADS, sc.smem_of[v], st>>>(sc.d_units + sc.unit_first[v], sc.d_jobs, sc.d_keys, pc->d, pr->d);
      else if (bps == 1) {
        // the variant's common window geometry goes through a tensor map of the reference plane (the other units stage row by row)
        CUtensorMap map;
        memset(&map, 0, sizeof(map));
        if (cu.tma_w[~v] > 0 && !encode_plane_map(pr->d, cu.tma_w[~v], cu.tma_h[~v], &map))
          return fail(HMB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the reference plane");
        search8_cu_kernels()[~v]<<<cu.unit_count[~v], CU8_THREADS, cu.smem_of[~v], st>>>(cu.d_units + cu.unit_first[~v], cu.d_bundles, sc.d_keys,
                                                                                      pc->d, pr->d, map);
      } else
        search16_cu_kernels()[~v]<<<cu.unit_count[~v], S8_THREADS, cu.smem_of[~v], st>>>(cu.d_units + cu.unit_first[~v], cu.d_bundles, sc.d_keys, pc->d, pr->d);
      G.launches++;
    }
    for (int k = 0; k < N_SIDE && k + 1 < used; k++) {
      CUDA_TRY(cudaEventRecord(G.ev_join[k], G.side[k]));
      CUDA_TRY(cudaStreamWaitEvent(G.stream, G.ev_join[k], 0));
    }
    if (sc.n_leftover > 0) {    // shapes / windows the tiled kernels do not cover
      if (bps == 1) k_search_generic<uint8_t, uint8_t><<<sc.n_leftover, 256, 0, G.stream>>>(p->d_tasks, p->d_results, pc->d, pr->d, sc.d_leftover);
      else          k_search_generic<int16_t, int16_t><<<sc.n_leftover, 256, 0, G.stream>>>(p->d_tasks, p->d_results, pc->d, pr->d, sc.d_leftover);
      G.launches++;
    }
    k_search8_finalize<<<(sc.n_tasks + 255) / 256, 256, 0, G.stream>>>(p->d_tasks, sc.d_keys, p->d_results, sc.n_tasks);
    G.launches++;
  } else if (!(p->flags & HMB200_FLAG_TZ)) {
    dispatch_generic(p->d_tasks, p->d_results, p->n, pc->d, pr->d, p->flags & ~HMB200_FLAG_FRAC, /*do_search=*/true, nullptr);
  }
  CUDA_TRY(cudaEventRecord(G.ev[1], G.stream));
  if (p->flags & HMB200_FLAG_FRAC) {
    const bool had = (p->flags & HMB200_FLAG_HADME) != 0;
    int nl;
    if (pc->d.bytes_per_sample == 1 && pr->d.bytes_per_sample == 1)
      nl = frac_launch<uint8_t, uint8_t>(p->frac, p->d_tasks, p->d_results, pc->d, pr->d, had, G.stream, G.side[0], G.ev_fork, G.ev_join[0]);
    else if (pc->d.bytes_per_sample == 2 && pr->d.bytes_per_sample == 2)
      nl = frac_launch<int16_t, int16_t>(p->frac, p->d_tasks, p->d_results, pc->d, pr->d, had, G.stream, G.side[0], G.ev_fork, G.ev_join[0]);
    else { dispatch_generic(p->d_tasks, p->d_results, p->n, pc->d, pr->d, p->flags, /*do_search=*/false, nullptr); nl = 0; }
    if (nl < 0) return fail(HMB200_ERR_CUDA, std::string("frac_launch: ") + cudaGetErrorString(cudaGetLastError()));
    G.launches += (uint64_t)nl;
  }
  CUDA_TRY(cudaEventRecord(G.ev[2], G.stream));
  if (!p->ev_done) CUDA_TRY(cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(p->ev_done, G.stream));
  CUDA_TRY(cudaGetLastError());
  return HMB200_OK;
}

int hmb200_last_timing(float* total_ms, float* search_ms, float* frac_ms) {
  NEED_READY();
  CUDA_TRY(cudaEventSynchronize(G.ev[2]));
  CUDA_TRY(cudaEventElapsedTime(&G.last_total_ms, G.ev[0], G.ev[2]));
  CUDA_TRY(cudaEventElapsedTime(&G.last_search_ms, G.ev[0], G.ev[1]));
  CUDA_TRY(cudaEventElapsedTime(&G.last_frac_ms, G.ev[1], G.ev[2]));
  if (total_ms) *total_ms = G.last_total_ms;
  if (search_ms) *search_ms = G.last_search_ms;
  if (frac_ms) *frac_ms = G.last_frac_ms;
  return HMB200_OK;
}

int hmb200_fetch_results_async(hmb200_prepared* p, hmb200_pu_result* results) {
  NEED_READY();
  if (!p || (p->n > 0 && !results)) return fail(HMB200_ERR_ARG, "hmb200_fetch_results_async: bad arguments");
  NEED_OWNER(p);
  if (p->n == 0) return HMB200_OK;
  if (!p->ev_done) return fail(HMB200_ERR_STATE, "hmb200_fetch_results_async: nothing has run on this handle");
  cudaPointerAttributes attr;
  const bool pinned = cudaPointerGetAttributes(&attr, results) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (!pinned) return fail(HMB200_ERR_ARG, "hmb200_fetch_results_async: results must come from hmb200_host_alloc (page-locked)");
  if (!p->ev_fetched) CUDA_TRY(cudaEventCreateWithFlags(&p->ev_fetched, cudaEventDisableTiming));
  CUDA_TRY(cudaStreamWaitEvent(G.copy, p->ev_done, 0));
  CUDA_TRY(cudaMemcpyAsync(results, p->d_results, (size_t)p->n * sizeof(hmb200_pu_result), cudaMemcpyDeviceToHost, G.copy));
  CUDA_TRY(cudaEventRecord(p->ev_fetched, G.copy));
  p->fetch_pending = true;
  return HMB200_OK;
}

int hmb200_fetch_results16_async(hmb200_prepared* p, hmb200_pu_result16* results) {
  NEED_READY();
  if (!p || (p->n > 0 && !results)) return fail(HMB200_ERR_ARG, "hmb200_fetch_results16_async: bad arguments");
  NEED_OWNER(p);
  if (p->n == 0) return HMB200_OK;
  if (!p->ev_done) return fail(HMB200_ERR_STATE, "hmb200_fetch_results16_async: nothing has run on this handle");
  cudaPointerAttributes attr;
  const bool pinned = cudaPointerGetAttributes(&attr, results) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (!pinned) return fail(HMB200_ERR_ARG, "hmb200_fetch_results16_async: results must come from hmb200_host_alloc (page-locked)");
  if (!p->d_packed) CUDA_TRY(cudaMalloc((void**)&p->d_packed, (size_t)p->n * sizeof(hmb200_pu_result16)));
  if (!p->ev_fetched) CUDA_TRY(cudaEventCreateWithFlags(&p->ev_fetched, cudaEventDisableTiming));
  // pack on the copy stream behind the run: the compute stream goes on with the next frame pair
  CUDA_TRY(cudaStreamWaitEvent(G.copy, p->ev_done, 0));
  k_pack_results16<<<(p->n + 255) / 256, 256, 0, G.copy>>>(p->d_results, p->d_packed, p->n);
  G.launches++;
  CUDA_TRY(cudaMemcpyAsync(results, p->d_packed, (size_t)p->n * sizeof(hmb200_pu_result16), cudaMemcpyDeviceToHost, G.copy));
  CUDA_TRY(cudaEventRecord(p->ev_fetched, G.copy));
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
  NEED_OWNER(p);
  if (p->n == 0) return HMB200_OK;
  if (p->fetch_pending) { CUDA_TRY(cudaEventSynchronize(p->ev_fetched)); p->fetch_pending = false; }
  const size_t bytes = (size_t)p->n * sizeof(hmb200_pu_result);
  cudaPointerAttributes attr;
  const bool user_pinned = cudaPointerGetAttributes(&attr, results) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();                                                                              // clear the "not registered" status
  if (user_pinned) {                                                                               // hmb200_host_alloc'ed: straight D2H
    CUDA_TRY(cudaMemcpyAsync(results, p->d_results, bytes, cudaMemcpyDeviceToHost, G.stream));
    CUDA_TRY(cudaStreamSynchronize(G.stream));
    CUDA_TRY(cudaGetLastError());
    return HMB200_OK;
  }
  int rc = ensure_pinned(bytes);
  if (rc != HMB200_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(G.pinned, p->d_results, bytes, cudaMemcpyDeviceToHost, G.stream));    // pinned staging: full PCIe rate
  CUDA_TRY(cudaStreamSynchronize(G.stream));
  CUDA_TRY(cudaGetLastError());
  memcpy(results, G.pinned, bytes);
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
// Fills the host record: pattern rows (dense: pitch = PU width), task, result seed, argmin key, ticket; one H2D copy brings it to
// the device.  *fits_u8: every pattern sample is in 0..255 (not the case for the 2*org - pred pattern of bi-prediction).
static int stage_one_call(const hmb200_pattern* key, const SearchTask& t, const hmb200_pu_result& io, const hmb200_tz_extra* tz, bool* fits_u8) {
  if (!key || !key->roi || !supported_pu(key->width, key->height)) return fail(HMB200_ERR_ARG, "unsupported pattern");
  char* h = G.one_host;
  *reinterpret_cast<unsigned long long*>(h + ONE_KEY) = ~0ull;
  *reinterpret_cast<unsigned long long*>(h + ONE_TICKET) = 0;
  memcpy(h + ONE_RESULT, &io, sizeof(io));
  memcpy(h + ONE_TASK, &t, sizeof(t));
  if (tz) memcpy(h + ONE_TZ, tz, sizeof(*tz));
  int16_t* rows = reinterpret_cast<int16_t*>(h + ONE_HEAD);
  int all = 0;
  for (int y = 0; y < key->height; y++) {
    const int16_t* src = key->roi + (ptrdiff_t)y * key->stride;
    int16_t* dst = rows + (size_t)y * key->width;
    for (int x = 0; x < key->width; x++) { dst[x] = src[x]; all |= src[x]; }
  }
  if (fits_u8) *fits_u8 = (all & ~0xff) == 0;
  CUDA_TRY(cudaMemcpyAsync(G.one_dev, h, ONE_HEAD + (size_t)key->height * key->width * sizeof(int16_t), cudaMemcpyHostToDevice, G.stream));
  return HMB200_OK;
}
// the pattern buffer as the plane the round-1 kernels read their "original" from
static DevPlane pattern_plane(int width) { DevPlane d = G.pattern.d; d.pitch = width; return d; }
static int finish_one_call(hmb200_pu_result* io) {
  CUDA_TRY(cudaMemcpyAsync(G.one_host + ONE_RESULT, G.one_dev + ONE_RESULT, sizeof(*io), cudaMemcpyDeviceToHost, G.stream));
  CUDA_TRY(cudaStreamSynchronize(G.stream));
  CUDA_TRY(cudaGetLastError());
  memcpy(io, G.one_host + ONE_RESULT, sizeof(*io));
  return HMB200_OK;
}
// The call's last kernel stores the result in mapped host memory as two 16-byte records ending in this call's sequence number
// (OneBack); the host spins on the numbers instead of paying a D2H copy and a stream synchronisation (7 us of a 16 us round trip,
// profiles/r02_latency_1to1.txt).
static OneBack next_one_back() {
  G.one_seq++;
  if (G.one_seq == 0) G.one_seq = 1;
  return OneBack{reinterpret_cast<uint4*>(G.one_back + BACK_A), reinterpret_cast<uint4*>(G.one_back + BACK_B), G.one_seq};
}
static int wait_records(volatile uint32_t* a, volatile uint32_t* b, uint32_t seq, hmb200_pu_result* io);
static int wait_one_back(const OneBack& back, hmb200_pu_result* io) {
  CUDA_TRY(cudaGetLastError());                                   // a failed launch would never report
  return wait_records(reinterpret_cast<volatile uint32_t*>(G.one_back + BACK_A), reinterpret_cast<volatile uint32_t*>(G.one_back + BACK_B), back.seq, io);
}
// a PU of up to 16x16 samples as a kernel argument (dense rows); *fits_u8 as in stage_one_call
static int fill_one_pattern(const hmb200_pattern* key, OnePattern* pat, bool* fits_u8) {
  if (!key || !key->roi || !supported_pu(key->width, key->height)) return fail(HMB200_ERR_ARG, "unsupported pattern");
  int all = 0;
  for (int y = 0; y < key->height; y++) {
    const int16_t* src = key->roi + (ptrdiff_t)y * key->stride;
    int16_t* dst = pat->px + (size_t)y * key->width;
    for (int x = 0; x < key->width; x++) { dst[x] = src[x]; all |= src[x]; }
  }
  *fits_u8 = (all & ~0xff) == 0;
  return HMB200_OK;
}
// k_one_frac launch: a thread per column of every (candidate, tile) of a stage, capped at one full CTA
static void launch_one_frac(const hmb200_pattern* key, const SearchTask& t, const hmb200_pu_result& seed, bool mv_from_device,
                            const DevPlane& ref, int flags, const OneBack& back) {
  const int had = (flags & HMB200_FLAG_HADME) ? 1 : 0;
  const int n = !had ? 4 : ((key->width % 8 == 0 && key->height % 8 == 0) ? 8 : 4);
  const int want = 9 * (key->width / n) * (key->height / n) * n;
  const int threads = std::min(ONE_FRAC_THREADS_MAX, std::max(128, (want + 31) & ~31));
  const int smem = one_frac_smem(key->width, key->height);
  hmb200_pu_result* d_r = reinterpret_cast<hmb200_pu_result*>(G.one_dev + ONE_RESULT);
  const int16_t* d_pat = reinterpret_cast<const int16_t*>(G.one_dev + ONE_HEAD);
  if (ref.bytes_per_sample == 1) k_one_frac<uint8_t><<<1, threads, smem, G.stream>>>(t, seed, mv_from_device ? 1 : 0, d_r, d_pat, ref, had, back);
  else                           k_one_frac<int16_t><<<1, threads, smem, G.stream>>>(t, seed, mv_from_device ? 1 : 0, d_r, d_pat, ref, had, back);
  G.launches++;
}

static int run_single(const hmb200_pattern* key, const int16_t* ref_at_pu, const SearchTask& proto, int flags, bool do_search,
                      hmb200_pu_result* io) {
  int rx, ry;
  Plane* pr = find_plane_by_host(ref_at_pu, &rx, &ry);
  if (!pr) return fail(HMB200_ERR_ARG, "reference pointer does not fall into a registered plane");
  if (!key || pr->d.bit_depth != key->bit_depth) return fail(HMB200_ERR_ARG, "pattern bit depth differs from the reference plane");
  get_plane((int)(pr - &G.planes[0]));                // orders a pending upload of the plane before this call's kernels
  SearchTask t = proto;
  t.org_x = 0; t.org_y = 0; t.ref_x = rx; t.ref_y = ry; t.w = key->width; t.h = key->height;
  {
    const int reach = (flags & HMB200_FLAG_FRAC) ? FRAC_REACH : 0;
    const Box b = do_search ? Box{rx + t.lt_x - reach, ry + t.lt_y - reach, rx + t.rb_x + t.w + reach, ry + t.rb_y + t.h + reach}
                            : Box{rx + io->mv_x - reach, ry + io->mv_y - reach, rx + io->mv_x + t.w + reach, ry + io->mv_y + t.h + reach};
    if (!box_inside(pr->d, b)) return fail(HMB200_ERR_ARG, "the search window / refinement block leaves the padded reference plane");
  }
  const bool r8 = pr->d.bytes_per_sample == 1;
  const bool frac = (flags & HMB200_FLAG_FRAC) != 0;
  SearchTask* d_t = reinterpret_cast<SearchTask*>(G.one_dev + ONE_TASK);
  hmb200_pu_result* d_r = reinterpret_cast<hmb200_pu_result*>(G.one_dev + ONE_RESULT);
  unsigned long long* d_key = reinterpret_cast<unsigned long long*>(G.one_dev + ONE_KEY);
  uint32_t* d_ticket = reinterpret_cast<uint32_t*>(G.one_dev + ONE_TICKET);
  const int16_t* d_pat = reinterpret_cast<const int16_t*>(G.one_dev + ONE_HEAD);
  bool fits_u8 = false, staged = false;
  int rc;
  if (G.one_fast) {
    // hmb200_one.cuh: a CTA per candidate row, the last one decodes (and refines a small PU); larger PUs: refinement with a thread per
    // tile column in a second launch; result by flag.  Small PUs travel in the kernel arguments: no copy either way.
    const bool small = t.w * t.h <= ONE_ARG_SAMPLES;
    OnePattern arg;
    if (small) { if ((rc = fill_one_pattern(key, &arg, &fits_u8)) != HMB200_OK) return rc; }
    else       { if ((rc = stage_one_call(key, t, *io, nullptr, &fits_u8)) != HMB200_OK) return rc; staged = true; }
    const OneBack back = next_one_back(), none{nullptr, nullptr, 0};
    const int had = (flags & HMB200_FLAG_HADME) ? 1 : 0;
    bool searched = !do_search, fused = false;
    if (do_search) {
      const int nx = t.rb_x - t.lt_x + 1, ny = t.rb_y - t.lt_y + 1, rows = t.h >> t.sub_shift;
      const int col0 = rx + pr->d.margin_x + t.lt_x;
      const bool bytes = r8 && fits_u8 && pr->d.bit_depth == 8;
      fused = frac && t.w * t.h <= ONE_FUSE_MAX_SAMPLES;           // small PU: the search's last CTA refines, one launch per call
      const int smem = std::max(one_search_smem(bytes, col0, nx, t.w, rows), fused ? one_frac_smem(t.w, t.h) : 0);
      if (smem <= ONE_SMEM_MAX) {
        const OneBack& b = (frac && !fused) ? none : back;
        const int fuse = fused ? (ONE_FUSE_FRAC | (had ? ONE_FUSE_HAD : 0)) : 0;
        int threads = ONE_SEARCH_THREADS;
        if (fused) {                                               // enough threads for one pass over the refinement's tile columns
          const int n = !had ? 4 : ((t.w % 8 == 0 && t.h % 8 == 0) ? 8 : 4);
          threads = std::min(ONE_SEARCH_THREADS_MAX, std::max(threads, (9 * (t.w / n) * (t.h / n) * n + 31) & ~31));
        }
        if (small) {
          if (bytes)    k_one_search_args<true, uint8_t><<<ny, threads, smem, G.stream>>>(t, *io, arg, d_key, d_ticket, pr->d, fuse, b);
          else if (r8)  k_one_search_args<false, uint8_t><<<ny, threads, smem, G.stream>>>(t, *io, arg, d_key, d_ticket, pr->d, fuse, b);
          else          k_one_search_args<false, int16_t><<<ny, threads, smem, G.stream>>>(t, *io, arg, d_key, d_ticket, pr->d, fuse, b);
        } else {
          if (bytes)    k_one_search<true, uint8_t><<<ny, threads, smem, G.stream>>>(t, *io, d_key, d_ticket, d_r, d_pat, pr->d, fuse, b);
          else if (r8)  k_one_search<false, uint8_t><<<ny, threads, smem, G.stream>>>(t, *io, d_key, d_ticket, d_r, d_pat, pr->d, fuse, b);
          else          k_one_search<false, int16_t><<<ny, threads, smem, G.stream>>>(t, *io, d_key, d_ticket, d_r, d_pat, pr->d, fuse, b);
        }
        G.launches++;
        searched = true;
      }
    }
    if (searched) {
      if (frac && !fused) {
        if (small && !do_search) {                                 // refinement-only call of a small PU: arguments only
          const int n = !had ? 4 : ((t.w % 8 == 0 && t.h % 8 == 0) ? 8 : 4);
          const int threads = std::min(ONE_FRAC_THREADS_MAX, std::max(128, (9 * (t.w / n) * (t.h / n) * n + 31) & ~31));
          if (r8) k_one_frac_args<uint8_t><<<1, threads, one_frac_smem(t.w, t.h), G.stream>>>(t, *io, arg, pr->d, had, back);
          else    k_one_frac_args<int16_t><<<1, threads, one_frac_smem(t.w, t.h), G.stream>>>(t, *io, arg, pr->d, had, back);
          G.launches++;
        } else {
          launch_one_frac(key, t, *io, do_search, pr->d, flags, back);
        }
      }
      return wait_one_back(back, io);
    }
    // a window too large for shared memory: the round-1 kernels below
  }
  if (!staged && (rc = stage_one_call(key, t, *io, nullptr, &fits_u8)) != HMB200_OK) return rc;
  const DevPlane pat = pattern_plane(key->width);
  if (do_search) {
    // one PU: spread its candidates over the whole GPU (one CTA per ~256 candidates), fold with atomicMin, decode
    const long long total = (long long)(t.rb_x - t.lt_x + 1) * (t.rb_y - t.lt_y + 1);
    const int splits = (int)std::max<long long>(1, std::min<long long>(4 * G.sm_count, (total + 255) / 256));
    const dim3 grid(1, splits);
    if (r8) k_search_split<uint8_t, int16_t><<<grid, 256, 0, G.stream>>>(d_t, d_key, pat, pr->d);
    else    k_search_split<int16_t, int16_t><<<grid, 256, 0, G.stream>>>(d_t, d_key, pat, pr->d);
    k_search8_finalize<<<1, 32, 0, G.stream>>>(d_t, d_key, d_r, 1);
    G.launches += 2;
  }
  dispatch_generic(d_t, d_r, 1, pat, pr->d, flags, /*do_search=*/false, nullptr);
  return finish_one_call(io);
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

// waits for the two 16-byte records at a / b to show seq (see wait_one_back)
static int wait_records(volatile uint32_t* a, volatile uint32_t* b, uint32_t seq, hmb200_pu_result* io) {
  for (uint64_t spins = 1; a[3] != seq || b[3] != seq; spins++) {
    if ((spins & 0x3fff) == 0) {
      const cudaError_t e = cudaStreamQuery(G.stream);
      if (e == cudaSuccess) {
        if (a[3] != seq || b[3] != seq) return fail(HMB200_ERR_CUDA, "1:1 call: the stream finished without reporting a result");
        break;
      }
      if (e != cudaErrorNotReady) return fail(HMB200_ERR_CUDA, std::string("1:1 call: ") + cudaGetErrorString(e));
    }
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  const uint32_t fr = b[0];
  io->mv_x = (int32_t)a[0]; io->mv_y = (int32_t)a[1]; io->sad = a[2];
  io->half_x = (int8_t)(fr & 0xff); io->half_y = (int8_t)((fr >> 8) & 0xff);
  io->qter_x = (int8_t)((fr >> 16) & 0xff); io->qter_y = (int8_t)((fr >> 24) & 0xff);
  io->frac_cost = b[1];
  return HMB200_OK;
}

// The speculative CU path of hmb200_pattern_search_and_refine (see CuCacheEntry).  *served: *io holds the call's answer.
static int cu_speculation(const hmb200_pattern* key, const int16_t* ref_at_pu, const SearchTask& proto, int flags, hmb200_pu_result* io,
                          bool* served) {
  *served = false;
  int rx, ry;
  Plane* pr = find_plane_by_host(ref_at_pu, &rx, &ry);
  if (!pr || !key->roi || key->bit_depth != pr->d.bit_depth || !supported_pu(key->width, key->height))
    return HMB200_OK;                                   // the per-PU path handles (or rejects) it
  const bool bytes = pr->d.bytes_per_sample == 1;       // 8-bit plane: byte-SIMD kernels; 9..14-bit: the scalar 16-bit ones
  const int plane = (int)(pr - &G.planes[0]);
  const int w = key->width, h = key->height;
  // 1. an answer computed by an earlier CU launch?
  for (int ei = 0; ei < (int)G.cu_cache.size(); ei++) {
    CuCacheEntry& e = G.cu_cache[ei];
    if (!e.valid || e.plane != plane || e.flags != flags || e.lt_x != proto.lt_x || e.lt_y != proto.lt_y || e.rb_x != proto.rb_x ||
        e.rb_y != proto.rb_y || e.pred_x != proto.pred_x || e.pred_y != proto.pred_y || e.lambda != proto.lambda_cost) continue;
    const int dx = rx - e.rx, dy = ry - e.ry;
    if (dx < 0 || dy < 0 || dx + w > e.S || dy + h > e.S) continue;
    int part = -1;
    for (int p = 0; p < one_cu_pus(e.S) && part < 0; p++) {
      int ox, oy, pw, ph;
      one_cu_part(e.S, p, &ox, &oy, &pw, &ph);
      if (ox == dx && oy == dy && pw == w && ph == h) part = p;
    }
    if (part < 0) continue;
    bool same = true;
    for (int y = 0; y < h && same; y++)
      same = memcmp(key->roi + (ptrdiff_t)y * key->stride, e.pat + (size_t)(dy + y) * e.S + dx, (size_t)w * sizeof(int16_t)) == 0;
    if (!same) continue;
    if (!e.got[part]) {
      volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(G.one_back + BACK_CU + ((size_t)ei * ONE_CU_MAX_PUS + part) * 32);
      const int rc = wait_records(slot, slot + 4, e.seq, &e.res[part]);
      if (rc != HMB200_OK) return rc;
      e.got[part] = true;
    }
    *io = e.res[part];
    *served = true;
    G.cu_hits++;
    return HMB200_OK;
  }
  // 2. a 2Nx2N PU: search and refine the whole CU's partitions with this call's window, predictor and lambda
  if (w != h || (w != 8 && w != 16 && w != 32 && w != 64)) return HMB200_OK;
  const int S = w;
  SearchTask t = proto;
  t.org_x = 0; t.org_y = 0; t.ref_x = rx; t.ref_y = ry; t.w = S; t.h = S;
  if (!box_inside(pr->d, Box{rx + t.lt_x - FRAC_REACH, ry + t.lt_y - FRAC_REACH, rx + t.rb_x + S + FRAC_REACH, ry + t.rb_y + S + FRAC_REACH}))
    return HMB200_OK;                                   // run_single reports the error
  const int fen = (flags & HMB200_FLAG_FEN) ? 1 : 0, had = (flags & HMB200_FLAG_HADME) ? 1 : 0;
  const int nx = t.rb_x - t.lt_x + 1, ny = t.rb_y - t.lt_y + 1;
  const int rows = (S == 64 && fen) ? S / 2 : S;
  const int col0 = rx + pr->d.margin_x + t.lt_x;
  const int smem = rows * one_search_row_bytes(bytes, col0, nx, S) + rows * S * (bytes ? 1 : 2);
  if (smem > ONE_SMEM_MAX) return HMB200_OK;
  const int ei = G.cu_cache_next;
  CuCacheEntry& e = G.cu_cache[ei];
  e.valid = false;
  int all = 0;
  for (int y = 0; y < S; y++) {
    const int16_t* src = key->roi + (ptrdiff_t)y * key->stride;
    int16_t* dst = e.pat + (size_t)y * S;
    for (int x = 0; x < S; x++) { dst[x] = src[x]; all |= src[x]; }
  }
  if (all & ~((1 << pr->d.bit_depth) - 1)) return HMB200_OK;   // outside the sample range (the signed bi-prediction pattern): per-PU path
  get_plane(plane);                                     // orders a pending upload of the plane before this call's kernels
  const OneBack first = next_one_back();
  OneBack back = first;
  back.host_a = reinterpret_cast<uint4*>(G.one_back + BACK_CU + (size_t)ei * ONE_CU_MAX_PUS * 32);
  back.host_b = back.host_a + 1;
  unsigned long long* d_keys = reinterpret_cast<unsigned long long*>(G.one_dev + ONE_CU_KEYS);
  uint32_t* d_ticket = reinterpret_cast<uint32_t*>(G.one_dev + ONE_TICKET);
  hmb200_pu_result* d_out = reinterpret_cast<hmb200_pu_result*>(G.one_dev + ONE_CU_OUT);
  const int np = one_cu_pus(S);
  if (S <= 16) {
    OnePattern arg;
    memcpy(arg.px, e.pat, (size_t)S * S * sizeof(int16_t));
    if (S == 8 && bytes)       k_one_cu_search_args<8, true><<<ny, ONE_SEARCH_THREADS, smem, G.stream>>>(t, fen, arg, d_keys, d_ticket, d_out, pr->d);
    else if (S == 8)           k_one_cu_search_args<8, false><<<ny, ONE_SEARCH_THREADS, smem, G.stream>>>(t, fen, arg, d_keys, d_ticket, d_out, pr->d);
    else if (bytes)            k_one_cu_search_args<16, true><<<ny, ONE_SEARCH_THREADS, smem, G.stream>>>(t, fen, arg, d_keys, d_ticket, d_out, pr->d);
    else                       k_one_cu_search_args<16, false><<<ny, ONE_SEARCH_THREADS, smem, G.stream>>>(t, fen, arg, d_keys, d_ticket, d_out, pr->d);
    if (bytes) k_one_cu_frac_args<uint8_t><<<np, S == 8 ? 128 : 288, one_frac_smem(S, S), G.stream>>>(t, S, arg, d_out, pr->d, had, back);
    else       k_one_cu_frac_args<int16_t><<<np, S == 8 ? 128 : 288, one_frac_smem(S, S), G.stream>>>(t, S, arg, d_out, pr->d, had, back);
  } else {
    char* hrec = G.one_host;
    memcpy(hrec + ONE_HEAD, e.pat, (size_t)S * S * sizeof(int16_t));
    CUDA_TRY(cudaMemcpyAsync(G.one_dev + ONE_HEAD, hrec + ONE_HEAD, (size_t)S * S * sizeof(int16_t), cudaMemcpyHostToDevice, G.stream));
    const int16_t* d_pat = reinterpret_cast<const int16_t*>(G.one_dev + ONE_HEAD);
    if (S == 32 && bytes)      k_one_cu_search<32, true><<<ny, ONE_SEARCH_THREADS, smem, G.stream>>>(t, fen, d_keys, d_ticket, d_out, d_pat, pr->d);
    else if (S == 32)          k_one_cu_search<32, false><<<ny, ONE_SEARCH_THREADS, smem, G.stream>>>(t, fen, d_keys, d_ticket, d_out, d_pat, pr->d);
    else if (bytes)            k_one_cu_search<64, true><<<ny, ONE_SEARCH_THREADS, smem, G.stream>>>(t, fen, d_keys, d_ticket, d_out, d_pat, pr->d);
    else                       k_one_cu_search<64, false><<<ny, ONE_SEARCH_THREADS, smem, G.stream>>>(t, fen, d_keys, d_ticket, d_out, d_pat, pr->d);
    if (bytes) k_one_cu_frac<uint8_t><<<np, ONE_FRAC_THREADS_MAX, one_frac_smem(S, S), G.stream>>>(t, S, d_pat, d_out, pr->d, had, back);
    else       k_one_cu_frac<int16_t><<<np, ONE_FRAC_THREADS_MAX, one_frac_smem(S, S), G.stream>>>(t, S, d_pat, d_out, pr->d, had, back);
  }
  G.launches += 2;
  G.cu_launches++;
  CUDA_TRY(cudaGetLastError());
  e.plane = plane; e.rx = rx; e.ry = ry; e.S = S; e.flags = flags;
  e.lt_x = t.lt_x; e.lt_y = t.lt_y; e.rb_x = t.rb_x; e.rb_y = t.rb_y; e.pred_x = t.pred_x; e.pred_y = t.pred_y; e.lambda = t.lambda_cost;
  e.seq = first.seq;
  for (auto& g : e.got) g = false;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(back.host_a);
  const int rc = wait_records(slot, slot + 4, e.seq, &e.res[0]);
  if (rc != HMB200_OK) return rc;
  e.got[0] = true;
  e.valid = true;
  G.cu_cache_next = (ei + 1) % (int)G.cu_cache.size();
  *io = e.res[0];
  *served = true;
  return HMB200_OK;
}

void hmb200_one_call_stats(uint64_t* calls, uint64_t* cu_launches, uint64_t* served_from_cu) {
  if (calls) *calls = G.one_calls;
  if (cu_launches) *cu_launches = G.cu_launches;
  if (served_from_cu) *served_from_cu = G.cu_hits;
}

int hmb200_pattern_search_and_refine(const hmb200_pattern* key, const int16_t* ref_at_pu, int ref_stride, hmb200_mv lt, hmb200_mv rb,
                                     const hmb200_cost_state* cs, int flags, hmb200_mv* mv_out, uint32_t* sad_out, hmb200_mv* half_out,
                                     hmb200_mv* qter_out, uint32_t* frac_cost_out) {
  NEED_READY();
  (void)ref_stride;
  if (!key || !cs || !mv_out || !sad_out || !half_out || !qter_out || !frac_cost_out || rb.x < lt.x || rb.y < lt.y)
    return fail(HMB200_ERR_ARG, "hmb200_pattern_search_and_refine: bad arguments");
  SearchTask t{};
  t.lt_x = lt.x; t.lt_y = lt.y; t.rb_x = rb.x; t.rb_y = rb.y; t.pred_x = cs->pred.x; t.pred_y = cs->pred.y;
  t.lambda_cost = cs->lambda_cost;
  t.sub_shift = ((flags & HMB200_FLAG_FEN) && key->height > 8) ? 1 : 0;
  hmb200_pu_result r{};
  const int f = (flags & (HMB200_FLAG_FEN | HMB200_FLAG_HADME)) | HMB200_FLAG_FRAC;
  G.one_calls++;
  bool served = false;
  int rc = G.speculate ? cu_speculation(key, ref_at_pu, t, f, &r, &served) : HMB200_OK;
  if (rc != HMB200_OK) return rc;
  if (!served && (rc = run_single(key, ref_at_pu, t, f, true, &r)) != HMB200_OK) return rc;
  mv_out->x = r.mv_x; mv_out->y = r.mv_y; *sad_out = r.sad;
  half_out->x = r.half_x; half_out->y = r.half_y; qter_out->x = r.qter_x; qter_out->y = r.qter_y; *frac_cost_out = r.frac_cost;
  return HMB200_OK;
}

static int tz_single(const hmb200_pattern* key, const int16_t* ref_at_pu, hmb200_mv lt, hmb200_mv rb, const hmb200_cost_state* cs, int flags,
                     const hmb200_tz_extra* extra, int pic_w, int pic_h, int max_cu, int search_range, hmb200_pu_result* out) {
  NEED_READY();
  if (!key || !cs || !extra || !out || rb.x < lt.x || rb.y < lt.y || search_range < 1)
    return fail(HMB200_ERR_ARG, "hmb200_pattern_search_tz: bad arguments");
  int rx, ry;
  Plane* pr = find_plane_by_host(ref_at_pu, &rx, &ry);
  if (!pr) return fail(HMB200_ERR_ARG, "reference pointer does not fall into a registered plane");
  if (pr->d.bit_depth != key->bit_depth) return fail(HMB200_ERR_ARG, "pattern bit depth differs from the reference plane");
  SearchTask t{};
  t.org_x = 0; t.org_y = 0; t.ref_x = rx; t.ref_y = ry; t.w = key->width; t.h = key->height;
  t.lt_x = lt.x; t.lt_y = lt.y; t.rb_x = rb.x; t.rb_y = rb.y; t.pred_x = cs->pred.x; t.pred_y = cs->pred.y;
  {
    const int xmin = std::min(std::min(0, lt.x), -max_cu - 8 - extra->cu_x + 1), xmax = std::max(std::max(0, rb.x), pic_w + 8 - extra->cu_x - 1);
    const int ymin = std::min(std::min(0, lt.y), -max_cu - 8 - extra->cu_y + 1), ymax = std::max(std::max(0, rb.y), pic_h + 8 - extra->cu_y - 1);
    const int reach = (flags & HMB200_FLAG_FRAC) ? FRAC_REACH : 0;
    if (!box_inside(pr->d, Box{rx + xmin - reach, ry + ymin - reach, rx + xmax + t.w + reach, ry + ymax + t.h + reach}))
      return fail(HMB200_ERR_ARG, "hmb200_pattern_search_tz: the clipped search range leaves the padded reference plane");
  }
  t.lambda_cost = cs->lambda_cost;
  t.sub_shift = ((flags & HMB200_FLAG_FEN) && key->height > 8) ? 1 : 0;
  get_plane((int)(pr - &G.planes[0]));
  hmb200_pu_result r{};
  int rc = stage_one_call(key, t, r, extra, nullptr);
  if (rc != HMB200_OK) return rc;
  const DevPlane pat = pattern_plane(key->width);
  const TzParams P{pic_w, pic_h, max_cu, search_range, (flags & HMB200_FLAG_TZ_STOP) ? 1 : 0};
  const SearchTask* d_t = reinterpret_cast<const SearchTask*>(G.one_dev + ONE_TASK);
  const hmb200_tz_extra* d_e = reinterpret_cast<const hmb200_tz_extra*>(G.one_dev + ONE_TZ);
  hmb200_pu_result* d_r = reinterpret_cast<hmb200_pu_result*>(G.one_dev + ONE_RESULT);
  if (pr->d.bytes_per_sample == 1) k_tz_search<uint8_t, int16_t><<<1, TZ_WARPS * 32, 0, G.stream>>>(d_t, d_e, d_r, 1, pat, pr->d, P);
  else                             k_tz_search<int16_t, int16_t><<<1, TZ_WARPS * 32, 0, G.stream>>>(d_t, d_e, d_r, 1, pat, pr->d, P);
  G.launches++;
  if (G.one_fast) {
    const OneBack back = next_one_back();
    if (flags & HMB200_FLAG_FRAC) {                    // xPatternSearchFracDIF on the MV just found, same round trip
      launch_one_frac(key, t, r, true, pr->d, flags, back);
    } else {
      k_one_report<<<1, 1, 0, G.stream>>>(d_r, back);
      G.launches++;
    }
    if ((rc = wait_one_back(back, &r)) != HMB200_OK) return rc;
    *out = r;
    return HMB200_OK;
  }
  if (flags & HMB200_FLAG_FRAC)                       // xPatternSearchFracDIF on the MV just found, same round trip
    dispatch_generic(reinterpret_cast<const SearchTask*>(G.one_dev + ONE_TASK), d_r, 1, pat, pr->d, flags, /*do_search=*/false, nullptr);
  if ((rc = finish_one_call(&r)) != HMB200_OK) return rc;
  *out = r;
  return HMB200_OK;
}

int hmb200_pattern_search_tz(const hmb200_pattern* key, const int16_t* ref_at_pu, int ref_stride, hmb200_mv lt, hmb200_mv rb,
                             const hmb200_cost_state* cs, int flags, const hmb200_tz_extra* extra, int pic_w, int pic_h, int max_cu,
                             int search_range, hmb200_mv* mv_out, uint32_t* sad_out) {
  (void)ref_stride;
  if (!mv_out || !sad_out) return fail(HMB200_ERR_ARG, "hmb200_pattern_search_tz: bad arguments");
  hmb200_pu_result r{};
  const int rc = tz_single(key, ref_at_pu, lt, rb, cs, flags & ~(HMB200_FLAG_FRAC | HMB200_FLAG_HADME), extra, pic_w, pic_h, max_cu, search_range, &r);
  if (rc != HMB200_OK) return rc;
  mv_out->x = r.mv_x; mv_out->y = r.mv_y; *sad_out = r.sad;
  return HMB200_OK;
}

int hmb200_pattern_search_tz_and_refine(const hmb200_pattern* key, const int16_t* ref_at_pu, int ref_stride, hmb200_mv lt, hmb200_mv rb,
                                        const hmb200_cost_state* cs, int flags, const hmb200_tz_extra* extra, int pic_w, int pic_h, int max_cu,
                                        int search_range, hmb200_mv* mv_out, uint32_t* sad_out, hmb200_mv* half_out, hmb200_mv* qter_out,
                                        uint32_t* frac_cost_out) {
  (void)ref_stride;
  if (!mv_out || !sad_out || !half_out || !qter_out || !frac_cost_out) return fail(HMB200_ERR_ARG, "hmb200_pattern_search_tz_and_refine: bad arguments");
  hmb200_pu_result r{};
  const int rc = tz_single(key, ref_at_pu, lt, rb, cs, flags | HMB200_FLAG_FRAC, extra, pic_w, pic_h, max_cu, search_range, &r);
  if (rc != HMB200_OK) return rc;
  mv_out->x = r.mv_x; mv_out->y = r.mv_y; *sad_out = r.sad;
  half_out->x = r.half_x; half_out->y = r.half_y; qter_out->x = r.qter_x; qter_out->y = r.qter_y; *frac_cost_out = r.frac_cost;
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
