// Register-only / shared-memory microbenchmarks that calibrate the integer-SIMD roofline
// used by bench.py ("peak is defined from that measurement, not a datasheet").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_int microbench_int.cu
// Prints one JSON object.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c){
  uint32_t r; asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%3;" : "=r"(r) : "r"(a),"r"(b),"r"(c)); return r;
}
__device__ __forceinline__ uint32_t sad1(uint32_t a, uint32_t b, uint32_t c){
  uint32_t r; asm volatile("vabsdiff.u32.u32.u32.add %0,%1,%2,%3;" : "=r"(r) : "r"(a),"r"(b),"r"(c)); return r;
}

constexpr int NACC = 16;
constexpr int ITER = 2048;

// mode 0: VABSDIFF4.ACC only; 1: scalar VABSDIFF; 2: VABSDIFF4 + IMAD interleaved 1:1; 3: IADD3 only (alu reference)
// 4: IMAD only (fma-pipe reference); 5: VABSDIFF4 + LOP3 1:1 (same pipe?); 6: dp4a only
template<int MODE>
__global__ void __launch_bounds__(256) k_alu(uint32_t* out, uint32_t seed){
  uint32_t acc[NACC], a[NACC], m[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++){ acc[i]=threadIdx.x+i; a[i]=seed*(i+1)+threadIdx.x*0x01010101u; m[i]=seed+i; }
  uint32_t b = seed ^ 0x5a5a5a5au;
  for(int it=0; it<ITER; it++){
    #pragma unroll
    for(int i=0;i<NACC;i++){
      if(MODE==0) acc[i]=sad4(a[i],b,acc[i]);
      if(MODE==1) acc[i]=sad1(a[i],b,acc[i]);
      if(MODE==2){ acc[i]=sad4(a[i],b,acc[i]); m[i]=m[i]*b+a[i]; }
      if(MODE==3){ asm volatile("add.u32 %0,%0,%1;" : "+r"(acc[i]) : "r"(a[i])); }
      if(MODE==4){ m[i]=m[i]*b+a[i]; }
      if(MODE==5){ acc[i]=sad4(a[i],b,acc[i]); asm volatile("xor.b32 %0,%0,%1;" : "+r"(m[i]) : "r"(a[i])); }
      if(MODE==6){ acc[i]=__dp4a(a[i],b,acc[i]); }
    }
    b += 0x01010101u;
  }
  uint32_t s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=acc[i]^m[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// shared-memory load rates. mode 0: LDS.32 consecutive words; 1: LDS.128 consecutive; 2: LDS.128 broadcast;
// 3: LDS.32 broadcast; 4: LDS.32 stride-odd rows (lane*33)
template<int MODE>
__global__ void __launch_bounds__(256) k_lds(uint32_t* out, int iters){
  extern __shared__ __align__(16) uint32_t sm[];
  for(int i=threadIdx.x;i<12288;i+=blockDim.x) sm[i]=i*2654435761u;
  __syncthreads();
  int lane=threadIdx.x&31, w=threadIdx.x>>5;
  uint32_t s0=0,s1=0,s2=0,s3=0;
  int base = w*64;
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int u=0;u<16;u++){
      int off=(base + u*36 + (it&7)*4) & 4095;
      if(MODE==0){ s0 += sm[off+lane]; }
      if(MODE==1){ uint4 v=*reinterpret_cast<const uint4*>(&sm[(off&~3)+lane*4]); s0+=v.x; s1+=v.y; s2+=v.z; s3+=v.w; }
      if(MODE==2){ uint4 v=*reinterpret_cast<const uint4*>(&sm[(off&~3)]); s0+=v.x; s1+=v.y; s2+=v.z; s3+=v.w; }
      if(MODE==3){ s0 += sm[off]; }
      if(MODE==4){ s0 += sm[off+lane*33]; }
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s0+s1+s2+s3;
}

// mixed: R VABSDIFF4 per LDS.32 (consecutive) to see whether LDS issue steals ALU issue slots
template<int R>
__global__ void __launch_bounds__(256) k_mix(uint32_t* out, int iters, uint32_t seed){
  extern __shared__ __align__(16) uint32_t sm[];
  for(int i=threadIdx.x;i<12288;i+=blockDim.x) sm[i]=i*2654435761u;
  __syncthreads();
  int lane=threadIdx.x&31, w=threadIdx.x>>5;
  uint32_t acc[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++) acc[i]=i;
  uint32_t b=seed;
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int u=0;u<NACC;u++){
      uint32_t v = sm[((w*64+u*36+(it&7)*4)&4095)+lane];
      #pragma unroll
      for(int r=0;r<R;r++) acc[(u+r)%NACC]=sad4(v,b+r,acc[(u+r)%NACC]);
    }
    b+=0x01010101u;
  }
  uint32_t s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

static float timeit(void(*launch)(void*), void* ctx){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for(int i=0;i<3;i++) launch(ctx);
  CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int rep=0;rep<5;rep++){
    CK(cudaEventRecord(e0)); launch(ctx); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms;
  }
  return best;
}

struct Ctx{ uint32_t* out; int grid; int block; int which; };

static void launch_alu(void* p){ Ctx* c=(Ctx*)p;
  switch(c->which){
    case 0: k_alu<0><<<c->grid,c->block>>>(c->out,12345u); break;
    case 1: k_alu<1><<<c->grid,c->block>>>(c->out,12345u); break;
    case 2: k_alu<2><<<c->grid,c->block>>>(c->out,12345u); break;
    case 3: k_alu<3><<<c->grid,c->block>>>(c->out,12345u); break;
    case 4: k_alu<4><<<c->grid,c->block>>>(c->out,12345u); break;
    case 5: k_alu<5><<<c->grid,c->block>>>(c->out,12345u); break;
    case 6: k_alu<6><<<c->grid,c->block>>>(c->out,12345u); break;
  }}
static const int LDS_ITERS=2048;
static void launch_lds(void* p){ Ctx* c=(Ctx*)p;
  size_t sh=49152;
  switch(c->which){
    case 0: k_lds<0><<<c->grid,c->block,sh>>>(c->out,LDS_ITERS); break;
    case 1: k_lds<1><<<c->grid,c->block,sh>>>(c->out,LDS_ITERS); break;
    case 2: k_lds<2><<<c->grid,c->block,sh>>>(c->out,LDS_ITERS); break;
    case 3: k_lds<3><<<c->grid,c->block,sh>>>(c->out,LDS_ITERS); break;
    case 4: k_lds<4><<<c->grid,c->block,sh>>>(c->out,LDS_ITERS); break;
  }}
static void launch_mix(void* p){ Ctx* c=(Ctx*)p;
  size_t sh=49152;
  switch(c->which){
    case 1: k_mix<1><<<c->grid,c->block,sh>>>(c->out,LDS_ITERS,99u); break;
    case 2: k_mix<2><<<c->grid,c->block,sh>>>(c->out,LDS_ITERS,99u); break;
    case 4: k_mix<4><<<c->grid,c->block,sh>>>(c->out,LDS_ITERS,99u); break;
    case 8: k_mix<8><<<c->grid,c->block,sh>>>(c->out,LDS_ITERS,99u); break;
  }}

int main(){
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,0));
  int sms=pr.multiProcessorCount;
  int clk_khz=0; CK(cudaDeviceGetAttribute(&clk_khz,cudaDevAttrClockRate,0));
  uint32_t* out; CK(cudaMalloc(&out, sizeof(uint32_t)*sms*8*256*2));
  printf("{\"gpu\":\"%s\",\"sms\":%d,\"clock_khz_attr\":%d", pr.name, sms, clk_khz);
  const char* alu_names[7]={"vabsdiff4_acc","vabsdiff_scalar","vabsdiff4_plus_imad","iadd","imad","vabsdiff4_plus_lop","dp4a"};
  for(int wps=2; wps<=8; wps*=2){          // CTAs of 256 threads per SM: 2 -> 16 warps/SM, 4 -> 32, 8 -> 64
    for(int m=0;m<7;m++){
      Ctx c{out,sms*wps,256,m};
      float ms=timeit(launch_alu,&c);
      double ops=(double)sms*wps*256*(double)ITER*NACC*((m==2||m==5)?1:1); // primary-op lane-instructions
      printf(",\"alu_%s_ctas%d_Glaneops_s\":%.2f", alu_names[m], wps, ops/ms/1e6);
    }
  }
  const char* lds_names[5]={"lds32_consec","lds128_consec","lds128_bcast","lds32_bcast","lds32_stride33"};
  for(int m=0;m<5;m++){
    Ctx c{out,sms*4,256,m};
    float ms=timeit(launch_lds,&c);
    double n=(double)sms*4*8*(double)LDS_ITERS*16; // warp-level LDS instructions
    printf(",\"%s_Gwarpinstr_s\":%.3f", lds_names[m], n/ms/1e6);
  }
  int rs[4]={1,2,4,8};
  for(int i=0;i<4;i++){
    Ctx c{out,sms*4,256,rs[i]};
    float ms=timeit(launch_mix,&c);
    double ops=(double)sms*4*256*(double)LDS_ITERS*NACC*rs[i];
    printf(",\"mix_lds32_per_%d_sad4_Glaneops_s\":%.2f", rs[i], ops/ms/1e6);
  }
  printf("}\n");
  return 0;
}
