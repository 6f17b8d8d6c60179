// Two (or more) host threads inside ONE process, one hmb200 context each, searching disjoint frame pairs of a lookahead
// concurrently through the C-ABI - the way a C++ encoder / lookahead shards over the GPUs of a box (the frame
// loop it replaces is TAppEncTop::encode, App/TAppEncoder/TAppEncTop.cpp:478-520).  Contexts go to device (i mod
// device count): with one GPU they share it, with N they use N.  Every pair is then repeated on the default context by the
// main thread and must match byte for byte.  Also checks that a prepared handle is refused under a foreign context.
//
//   g++ -O2 -std=c++17 -pthread multi_ctx_test.cpp -I../../include -L../../video_codecs_b200 -lhmb200 -Wl,-rpath,...
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "hmb200.h"

namespace {

constexpr int W = 320, H = 192, SR = 32, MARGIN = 80, N_FRAMES = 7;

// smooth seeded texture, translated by (2t, -t) plus a little noise: non-trivial MV field
std::vector<uint8_t> make_frame(int t) {
  std::vector<uint8_t> f((size_t)W * H);
  auto tex = [](int x, int y) {
    uint32_t h = (uint32_t)(x >> 2) * 73856093u ^ (uint32_t)(y >> 2) * 19349663u;
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    return (int)(h & 127) + ((x * 3 + y * 5) & 63);
  };
  uint32_t lcg = 12345u + 977u * (uint32_t)t;
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      lcg = lcg * 1664525u + 1013904223u;
      int v = tex(x + 2 * t + 64, y - t + 64) + (int)((lcg >> 24) & 3);
      f[(size_t)y * W + x] = (uint8_t)(v > 255 ? 255 : v);
    }
  return f;
}

struct Shared {
  std::vector<std::vector<uint8_t>> frames;
  std::vector<hmb200_pu_job> jobs;
  std::vector<std::vector<hmb200_pu_result>> out;    // per frame pair
  int fail = 0;
};

bool run_pair(Shared& S, int pair, std::vector<hmb200_pu_result>& res) {
  const int idc = hmb200_register_plane_u8(S.frames[pair + 1].data(), W, W, H, MARGIN, MARGIN, HMB200_PLANE_ORG, pair + 1);
  const int idr = hmb200_register_plane_u8(S.frames[pair].data(), W, W, H, MARGIN, MARGIN, HMB200_PLANE_REC, pair);
  if (idc < 0 || idr < 0) { fprintf(stderr, "register_plane: %s\n", hmb200_last_error()); return false; }
  res.assign(S.jobs.size(), hmb200_pu_result{});
  const int rc = hmb200_me_jobs(idc, idr, S.jobs.data(), (int)S.jobs.size(), HMB200_FLAG_FEN | HMB200_FLAG_HADME | HMB200_FLAG_FRAC, res.data());
  hmb200_release_plane(idc);
  hmb200_release_plane(idr);
  if (rc != HMB200_OK) { fprintf(stderr, "me_jobs: %s\n", hmb200_last_error()); return false; }
  return true;
}

void worker(Shared* S, int index, int n_workers, int device, hmb200_prepared** foreign) {
  hmb200_ctx* ctx = hmb200_ctx_create(device);
  if (!ctx) { fprintf(stderr, "worker %d: ctx_create(%d): %s\n", index, device, hmb200_last_error()); S->fail = 1; return; }
  if (hmb200_ctx_get_current() != ctx || hmb200_ctx_device(ctx) != device) { fprintf(stderr, "worker %d: context not current\n", index); S->fail = 1; }
  for (int rep = 0; rep < 3; rep++)                  // several rounds so that the threads really overlap
    for (int pair = index; pair < N_FRAMES - 1; pair += n_workers)
      if (!run_pair(*S, pair, S->out[pair])) S->fail = 1;
  if (index == 0 && foreign) {                       // a handle prepared here is handed to the main thread's context below
    *foreign = hmb200_prepare_jobs(S->jobs.data(), 16, HMB200_FLAG_FEN, 8);
    if (!*foreign) S->fail = 1;
  } else {
    hmb200_ctx_destroy(ctx);
  }
}

}  // namespace

int main(int argc, char** argv) {
  const int n_workers = argc > 1 ? atoi(argv[1]) : 2;
  const int n_devices = argc > 2 ? atoi(argv[2]) : 1;
  if (hmb200_init(0) != HMB200_OK) { fprintf(stderr, "hmb200_init: %s\n", hmb200_last_error()); return 2; }
  Shared S;
  for (int t = 0; t < N_FRAMES; t++) S.frames.push_back(make_frame(t));
  const hmb200_mv zero{0, 0};
  const uint32_t lam = hmb200_motion_lambda_cost(0.4624 * 203.187);      // QP 35 lowdelay-P slice
  const int n = hmb200_build_canonical_jobs(W, H, 64, SR, lam, zero, 0, -1, nullptr, 0);
  S.jobs.resize((size_t)n);
  hmb200_build_canonical_jobs(W, H, 64, SR, lam, zero, 0, -1, S.jobs.data(), n);
  S.out.resize(N_FRAMES - 1);

  hmb200_prepared* foreign = nullptr;
  std::vector<std::thread> th;
  for (int i = 0; i < n_workers; i++) th.emplace_back(worker, &S, i, n_workers, i % n_devices, &foreign);
  for (auto& t : th) t.join();
  if (S.fail) { fprintf(stderr, "FAIL: a worker reported an error\n"); return 1; }

  // single-context reference run on the main thread (default context, device 0)
  if (hmb200_ctx_set_current(nullptr) != HMB200_OK) { fprintf(stderr, "set_current(NULL): %s\n", hmb200_last_error()); return 1; }
  int bad = 0, nontrivial = 0;
  for (int pair = 0; pair < N_FRAMES - 1; pair++) {
    std::vector<hmb200_pu_result> ref;
    if (!run_pair(S, pair, ref)) return 1;
    if (memcmp(ref.data(), S.out[pair].data(), ref.size() * sizeof(hmb200_pu_result)) != 0) { fprintf(stderr, "pair %d differs\n", pair); bad++; }
    for (auto& r : ref) if (r.mv_x != 0 || r.mv_y != 0) nontrivial++;
  }
  // the handle of worker 0's context must be refused here, not run on the wrong device's streams
  if (foreign) {
    const int rc = hmb200_run_prepared(foreign, 0, 0);
    if (rc != HMB200_ERR_STATE && rc != HMB200_ERR_ARG) { fprintf(stderr, "foreign handle was accepted (rc %d)\n", rc); bad++; }
    hmb200_free_prepared(foreign);
  }
  hmb200_shutdown();
  if (bad || !nontrivial) { fprintf(stderr, "FAIL: %d mismatching pairs, %d non-zero MVs\n", bad, nontrivial); return 1; }
  printf("ok: %d workers on %d device(s), %d pairs x %d PUs identical to the single-context run, %d non-zero MVs\n", n_workers, n_devices,
         N_FRAMES - 1, n, nontrivial);
  return 0;
}
