// How should a kernel hand a 32-byte result to a spinning host thread, and what does the input copy cost?
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o rr_probe tools/roundtrip_report_probe.cu && ./rr_probe
// mode 0: 256-byte H2D copy + kernel that stores 8 words, __threadfence_system, flag      mode 1: H2D copy + two 16-byte stores that
// carry the sequence number in their last word (no fence)      mode 2 / 3: the same without any copy      mode 4: a 512-byte struct
// as kernel argument, two 16-byte stores.  Counts records whose payload did not match their sequence number ("torn").
// Result on B200: profiles/r02_latency_1to1.txt.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <chrono>
struct Pat { int16_t px[256]; };
__global__ void k_fence(const int* in, int* out, volatile int* host_res, volatile int* host_flag, int seq) {
  if (threadIdx.x == 0) { int v = (in ? in[0] : seq) + 1; out[0] = v; for (int i = 0; i < 8; i++) host_res[i] = v + i; __threadfence_system(); host_flag[0] = seq; }
}
__global__ void k_vec(const int* in, int* out, uint4* host_a, uint4* host_b, int seq) {
  if (threadIdx.x == 0) { int v = (in ? in[0] : seq) + 1; out[0] = v; *host_a = make_uint4(v, v + 1, v + 2, seq); *host_b = make_uint4(v + 3, v + 4, 0, seq); }
}
__global__ void k_args(const __grid_constant__ Pat p, int idx, int* out, uint4* host_a, uint4* host_b, int seq) {
  if (threadIdx.x == 0) { int v = p.px[idx & 255] + 1; out[0] = v; *host_a = make_uint4(v, v + 1, v + 2, seq); *host_b = make_uint4(v + 3, v + 4, 0, seq); }
}
static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaSetDevice(0);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  char* h; cudaHostAlloc((void**)&h, 16384, cudaHostAllocMapped);
  char* d; cudaMalloc((void**)&d, 16384);
  int* dout; cudaMalloc((void**)&dout, 64);
  volatile int* flag = (volatile int*)(h + 8192 + 64); volatile int* res = (volatile int*)(h + 8192);
  volatile uint32_t* a = (volatile uint32_t*)(h + 8192 + 128); volatile uint32_t* b = (volatile uint32_t*)(h + 8192 + 144);
  Pat p; for (int i = 0; i < 256; i++) p.px[i] = (int16_t)i;
  const int N = 4000;
  for (int mode = 0; mode < 5; mode++) {
    double t0 = 0; long bad = 0;
    for (int it = -300; it < N; it++) {
      if (it == 0) t0 = now();
      const int seq = it + 1000;
      *(int*)h = it;
      if (mode < 2) cudaMemcpyAsync(d, h, 256, cudaMemcpyHostToDevice, st);
      if (mode == 0) k_fence<<<1, 32, 0, st>>>((const int*)d, dout, (volatile int*)res, (volatile int*)flag, seq);
      if (mode == 1) k_vec<<<1, 32, 0, st>>>((const int*)d, dout, (uint4*)a, (uint4*)b, seq);
      if (mode == 2) k_fence<<<1, 32, 0, st>>>(nullptr, dout, (volatile int*)res, (volatile int*)flag, seq);
      if (mode == 3) k_vec<<<1, 32, 0, st>>>(nullptr, dout, (uint4*)a, (uint4*)b, seq);
      if (mode == 4) k_args<<<1, 32, 0, st>>>(p, it, dout, (uint4*)a, (uint4*)b, seq);
      if (mode == 0 || mode == 2) { while (*flag != seq) { } }
      else { while (a[3] != (uint32_t)seq || b[3] != (uint32_t)seq) { } if (a[1] != a[0] + 1 || b[1] != b[0] + 1 || b[0] != a[0] + 3) bad++; }
    }
    const double us = (now() - t0) / N;
    cudaStreamSynchronize(st);
    printf("mode %d: %.2f us per round trip, %ld torn records\n", mode, us, bad);
  }
  return 0;
}
