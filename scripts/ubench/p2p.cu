// Microbenchmark: SM-driven peer reads / writes over NVLink between GPU 0 and GPU 1 (both
// directions at once, like the x-line kernel), float2 and float4 accesses.
#include <cstdio>
#include <cuda_runtime.h>
#define OK(x) do{cudaError_t e=(x); if(e){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
template<typename T> __global__ void copyk(T* __restrict__ dst, const T* __restrict__ src, size_t n){
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x, s = (size_t)gridDim.x*blockDim.x;
  for(; i<n; i+=s) dst[i]=src[i];
}
int main(){
  int n=0; cudaGetDeviceCount(&n); if(n<2){printf("need 2 GPUs\n"); return 0;}
  const size_t bytes = 256u<<20;
  void *a[2], *b[2]; cudaStream_t st[2]; cudaEvent_t e0[2], e1[2];
  for(int d=0; d<2; ++d){ OK(cudaSetDevice(d)); OK(cudaDeviceEnablePeerAccess(1-d,0)); OK(cudaMalloc(&a[d],bytes)); OK(cudaMalloc(&b[d],bytes)); OK(cudaMemset(a[d],1,bytes)); OK(cudaStreamCreate(&st[d])); cudaEventCreate(&e0[d]); cudaEventCreate(&e1[d]); }
  for(int d=0; d<2; ++d){ cudaSetDevice(d); cudaDeviceSynchronize(); }
  const char* names[] = {"local copy", "peer READ  (both GPUs at once)", "peer WRITE (both GPUs at once)", "peer READ + WRITE each GPU (x-line pattern)"};
  for(int grid : {148, 148*4, 148*16}) for(int mode=0; mode<4; ++mode){
    float best=1e9;
    for(int rep=0; rep<4; ++rep){
      for(int d=0; d<2; ++d){ cudaSetDevice(d); cudaEventRecord(e0[d], st[d]);
        if(mode==0) copyk<float4><<<grid,512,0,st[d]>>>((float4*)b[d],(const float4*)a[d],bytes/16);
        if(mode==1) copyk<float4><<<grid,512,0,st[d]>>>((float4*)b[d],(const float4*)a[1-d],bytes/16);
        if(mode==2) copyk<float4><<<grid,512,0,st[d]>>>((float4*)b[1-d],(const float4*)a[d],bytes/16);
        if(mode==3){ copyk<float4><<<grid/2,512,0,st[d]>>>((float4*)b[d],(const float4*)a[1-d],bytes/32);
                     copyk<float4><<<grid/2,512,0,st[d]>>>((float4*)b[1-d]+bytes/32,(const float4*)a[d]+bytes/32,bytes/32); }
        cudaEventRecord(e1[d], st[d]); }
      float ms=0; for(int d=0; d<2; ++d){ cudaSetDevice(d); cudaEventSynchronize(e1[d]); float m; cudaEventElapsedTime(&m,e0[d],e1[d]); ms = m>ms?m:ms; }
      best = ms<best?ms:best;
    }
    printf("grid %5d  %-46s %.3f ms  %.0f GB/s per GPU\n", grid, names[mode], best, bytes/best/1e6);
  }
  return 0;
}
