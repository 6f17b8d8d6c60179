// round-trip probe: how fast can one tiny request go host -> GPU -> host?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <chrono>
#include <cstring>
__global__ void k_work(const int* in, int* out, int n) {   // ~trivial kernel, n CTAs
  __shared__ int s;
  if (threadIdx.x == 0) s = in[0];
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = s + 1;
}
__global__ void k_last(const int* in, int* out, volatile int* host_res, volatile int* host_flag, int seq) {
  if (threadIdx.x == 0) { int v = in[0] + 1; out[0] = v; if (host_res) { host_res[0] = v; __threadfence_system(); host_flag[0] = seq; } }
}
static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaSetDevice(0);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  char* h; cudaHostAlloc((void**)&h, 16384, cudaHostAllocMapped);
  char* hd; cudaHostGetDevicePointer((void**)&hd, h, 0);
  char* d; cudaMalloc((void**)&d, 16384);
  int* dout; cudaMalloc((void**)&dout, 64);
  volatile int* flag = (volatile int*)(h + 8192); volatile int* res = (volatile int*)(h + 8192 + 64);
  int* dflag = (int*)(hd + 8192); int* dres = (int*)(hd + 8192 + 64);
  const int N = 3000;
  for (int bytes : {256, 1024, 8448}) for (int nk : {1, 3}) for (int mode = 0; mode < 4; mode++) {
    // mode 0: H2D copy, nk kernels, D2H copy, stream sync      mode 1: H2D copy, kernels, mapped result + flag spin
    // mode 2: kernels read mapped host memory, D2H copy + sync  mode 3: all mapped, flag spin
    double t0 = 0;
    for (int it = -200; it < N; it++) {
      if (it == 0) t0 = now();
      const int seq = it + 1000;
      *(int*)h = it;
      const int* src = (mode >= 2) ? (const int*)hd : (const int*)d;
      if (mode < 2) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
      for (int k = 0; k + 1 < nk; k++) k_work<<<66, 256, 0, st>>>(src, dout, 66);
      const bool spin = (mode & 1);
      k_last<<<1, 32, 0, st>>>(nk > 1 ? dout : src, dout + 1, spin ? dres : nullptr, spin ? dflag : nullptr, seq);
      if (spin) { while (*flag != seq) { } }
      else { cudaMemcpyAsync(h + 4096, dout + 1, 48, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); }
    }
    const double us = (now() - t0) / N;
    cudaStreamSynchronize(st);
    printf("bytes %5d kernels %d mode %d: %.2f us per round trip\n", bytes, nk, mode, us);
  }
  return 0;
}
