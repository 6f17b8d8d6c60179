// Register-only calibration of the 16-bit (bit depths 9..14) SAD sequence of k_search16_cu
// (hmb200_search16_cu.cuh: VIADD.16x2 + VIADDMNMX.S16x2 + IDP.2A.LO per two samples) and of its parts, so that
// bench.py reports 10-bit content against its own integer-SIMD peak.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_int16 microbench_int16.cu ; prints one JSON object.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

constexpr int NACC = 16;
constexpr int ITER = 2048;

// mode 0: the kernel's sequence (3 instructions per 2 abs-diffs); 1: VIADD.16x2 only; 2: VIADDMNMX.S16x2 only;
// 3: IDP.2A.LO only; 4: VIMNMX.U16x2 + IDP.2A.LO (the sum-of-minima form, 2 instructions per 2 samples);
// 5: VIADD.16x2 + VIADDMNMX.U16x2 (sum of minima accumulated as packed half-words)
template<int MODE>
__global__ void __launch_bounds__(256) k_alu16(uint32_t* out, uint32_t seed){
  uint32_t acc[NACC], o[NACC], no[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++){ acc[i]=threadIdx.x+i; o[i]=(seed*(i+1)+threadIdx.x*0x00010001u)&0x03ff03ffu; no[i]=__vneg2(o[i]); }
  uint32_t r = (seed ^ 0x5a5a5a5au) & 0x03ff03ffu;
  for(int it=0; it<ITER; it++){
    const uint32_t nr = __vneg2(r);
    #pragma unroll
    for(int i=0;i<NACC;i++){
      if(MODE==0){ const uint32_t t=__vadd2(o[i],nr); const uint32_t m=__viaddmax_s16x2(r,no[i],t); acc[i]=(uint32_t)__dp2a_lo((int)m,0x0101,(int)acc[i]); }
      if(MODE==1){ acc[i]=__vadd2(acc[i],o[i]^r); }
      if(MODE==2){ acc[i]=__viaddmax_s16x2(r,no[i],acc[i]); }
      if(MODE==3){ acc[i]=(uint32_t)__dp2a_lo((int)(o[i]),(int)r,(int)acc[i]); }
      if(MODE==4){ const uint32_t m=__vminu2(o[i],r); acc[i]=(uint32_t)__dp2a_lo((int)m,0x0101,(int)acc[i]); }
      if(MODE==5){ const uint32_t t=__vadd2(acc[i],o[i]); acc[i]=__viaddmin_u16x2(acc[i],r,t); }
    }
    r = (r + 0x00010001u) & 0x03ff03ffu;
  }
  uint32_t s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

struct Ctx{ uint32_t* out; int grid; int which; };
static void launch(Ctx* c){
  switch(c->which){
    case 0: k_alu16<0><<<c->grid,256>>>(c->out,12345u); break;
    case 1: k_alu16<1><<<c->grid,256>>>(c->out,12345u); break;
    case 2: k_alu16<2><<<c->grid,256>>>(c->out,12345u); break;
    case 3: k_alu16<3><<<c->grid,256>>>(c->out,12345u); break;
    case 4: k_alu16<4><<<c->grid,256>>>(c->out,12345u); break;
    case 5: k_alu16<5><<<c->grid,256>>>(c->out,12345u); break;
  }}
static float timeit(Ctx* c){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for(int i=0;i<3;i++) launch(c);
  CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int rep=0;rep<5;rep++){
    CK(cudaEventRecord(e0)); launch(c); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms;
  }
  return best;
}

int main(){
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,0));
  const int sms=pr.multiProcessorCount;
  uint32_t* out; CK(cudaMalloc(&out, sizeof(uint32_t)*sms*8*256));
  printf("{\"gpu\":\"%s\",\"sms\":%d", pr.name, sms);
  const char* names[6]={"sad16x2_seq3","viadd16x2","viaddmnmx_s16x2","idp2a","vimnmx16x2_plus_idp2a","viadd16x2_plus_viaddmnmx_u16x2"};
  for(int ctas=2; ctas<=8; ctas*=2)
    for(int m=0;m<6;m++){
      Ctx c{out,sms*ctas,m};
      const float ms=timeit(&c);
      const double seqs=(double)sms*ctas*256*(double)ITER*NACC;      // lane-sequences (each covers two samples in modes 0, 4, 5)
      printf(",\"%s_ctas%d_Glaneseq_s\":%.2f", names[m], ctas, seqs/ms/1e6);
    }
  printf("}\n");
  return 0;
}
