// Microbenchmark: issue rate of packed FP32 (FFMA2/FADD2) against scalar FFMA/FADD on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(u64 r, float&a, float&b){asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(r));}
__device__ __forceinline__ u64 fma2(u64 a,u64 b,u64 c){u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ u64 add2(u64 a,u64 b){u64 r; asm("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}

template<int MODE>
__global__ void k(float* out, int iters, float s){
  float a[16]; for(int i=0;i<16;++i) a[i]=threadIdx.x*0.001f+i;
  if (MODE==0){ // scalar FFMA: 16 independent chains
    for(int it=0;it<iters;++it){
#pragma unroll
      for(int i=0;i<16;++i) a[i]=fmaf(a[i],s,1.0f);
    }
  } else if (MODE==1){ // packed: 8 chains of f32x2 (same flops)
    u64 p[8]; for(int i=0;i<8;++i) p[i]=pk(a[2*i],a[2*i+1]);
    u64 ss=pk(s,s), one=pk(1.f,1.f);
    for(int it=0;it<iters;++it){
#pragma unroll
      for(int i=0;i<8;++i) p[i]=fma2(p[i],ss,one);
    }
    for(int i=0;i<8;++i) upk(p[i],a[2*i],a[2*i+1]);
  } else if (MODE==2){ // scalar FADD
    for(int it=0;it<iters;++it){
#pragma unroll
      for(int i=0;i<16;++i) a[i]=a[i]+s;
    }
  } else if (MODE==3){
    u64 p[8]; for(int i=0;i<8;++i) p[i]=pk(a[2*i],a[2*i+1]);
    u64 ss=pk(s,s);
    for(int it=0;it<iters;++it){
#pragma unroll
      for(int i=0;i<8;++i) p[i]=add2(p[i],ss);
    }
    for(int i=0;i<8;++i) upk(p[i],a[2*i],a[2*i+1]);
  } else if (MODE==4){ // scalar FFMA mixed 1:1 with integer LOP3 (issue-bound mix)
    unsigned x[16]; for(int i=0;i<16;++i) x[i]=threadIdx.x+i;
    for(int it=0;it<iters;++it){
#pragma unroll
      for(int i=0;i<16;++i){ a[i]=fmaf(a[i],s,1.0f); x[i]=(x[i]^0x5bd1e995u)+ (x[i]>>3); }
    }
    for(int i=0;i<16;++i) a[i]+=x[i];
  } else if (MODE==5){
    unsigned x[16]; for(int i=0;i<16;++i) x[i]=threadIdx.x+i;
    u64 p[8]; for(int i=0;i<8;++i) p[i]=pk(a[2*i],a[2*i+1]);
    u64 ss=pk(s,s), one=pk(1.f,1.f);
    for(int it=0;it<iters;++it){
#pragma unroll
      for(int i=0;i<8;++i) p[i]=fma2(p[i],ss,one);
#pragma unroll
      for(int i=0;i<16;++i) x[i]=(x[i]^0x5bd1e995u)+ (x[i]>>3);
    }
    for(int i=0;i<8;++i) upk(p[i],a[2*i],a[2*i+1]);
    for(int i=0;i<16;++i) a[i]+=x[i];
  }
  float r=0; for(int i=0;i<16;++i) r+=a[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
template<int MODE> void run(const char* name){
  float* d; cudaMalloc(&d, 148*8*512*4);
  int iters=20000;
  k<MODE><<<148*2,512>>>(d,100,0.999f);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<148*2,512>>>(d,iters,0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double flop_elems = 148.0*2*512*16*iters; // fp32 element-ops
  printf("%-28s %.3f ms  %.1f Gelem-op/s  (%.1f elem-op/clk/SM at 1.965GHz)\n", name, ms, flop_elems/ms/1e6, flop_elems/ms/1e6/148/1.965);
  cudaFree(d);
}
int main(){
  run<0>("FFMA scalar"); run<1>("FFMA2 packed"); run<2>("FADD scalar"); run<3>("FADD2 packed");
  run<4>("FFMA + 2 int (scalar)"); run<5>("FFMA2 + 2 int");
  return 0;
}
