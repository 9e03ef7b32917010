// Microbenchmark: copy-engine peer copies (contiguous and strided 2-D) between GPU 0 and GPU 1,
// both GPUs active, in + out concurrently (the staged-transpose pattern).
#include <cstdio>
#include <cuda_runtime.h>
#define OK(x) do{cudaError_t e=(x); if(e){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
int main(){
  int n=0; cudaGetDeviceCount(&n); if(n<2){printf("need 2 GPUs\n"); return 0;}
  const size_t rows=2048, w=8192*8, bytes=rows*w;   // 128 MB block = one peer's share at 2 GPUs
  char *T[2], *S[2]; cudaStream_t sin[2], sout[2]; cudaEvent_t e0[2], e1[2], e2[2];
  for(int d=0; d<2; ++d){ OK(cudaSetDevice(d)); OK(cudaDeviceEnablePeerAccess(1-d,0)); OK(cudaMalloc(&T[d],2*bytes)); OK(cudaMalloc(&S[d],2*bytes)); OK(cudaMemset(T[d],1,2*bytes));
    cudaStreamCreateWithFlags(&sin[d],cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sout[d],cudaStreamNonBlocking); cudaEventCreate(&e0[d]); cudaEventCreate(&e1[d]); cudaEventCreate(&e2[d]); }
  for(int d=0; d<2; ++d){ cudaSetDevice(d); cudaDeviceSynchronize(); }
  const char* names[]={"contiguous peer->local (in only)","2D strided peer->local (in only)","2D strided in + out concurrently","2D local->local strided","2D strided in, 8 chunks"};
  for(int mode=0; mode<5; ++mode){
    float best=1e9;
    for(int rep=0; rep<4; ++rep){
      for(int d=0; d<2; ++d){ cudaSetDevice(d); cudaEventRecord(e0[d], sin[d]); cudaEventRecord(e2[d], sout[d]);
        if(mode==0) cudaMemcpyAsync(S[d], T[1-d], bytes, cudaMemcpyDeviceToDevice, sin[d]);
        if(mode==1||mode==2) cudaMemcpy2DAsync(S[d], 2*w, T[1-d], w, w, rows, cudaMemcpyDeviceToDevice, sin[d]);
        if(mode==2) cudaMemcpy2DAsync(T[1-d]+bytes, w, S[d]+w, 2*w, w, rows, cudaMemcpyDeviceToDevice, sout[d]);
        if(mode==3) cudaMemcpy2DAsync(S[d], 2*w, T[d], w, w, rows, cudaMemcpyDeviceToDevice, sin[d]);
        if(mode==4) for(int c=0;c<8;++c) cudaMemcpy2DAsync(S[d]+c*(rows/8)*2*w, 2*w, T[1-d]+c*(rows/8)*w, w, w, rows/8, cudaMemcpyDeviceToDevice, sin[d]);
        cudaEventRecord(e1[d], sin[d]); cudaStreamWaitEvent(sin[d], e2[d], 0); }
      float ms=0; for(int d=0; d<2; ++d){ cudaSetDevice(d); cudaStreamSynchronize(sout[d]); cudaEventSynchronize(e1[d]); float m; cudaEventElapsedTime(&m,e0[d],e1[d]); ms = m>ms?m:ms; }
      for(int d=0; d<2; ++d){ cudaSetDevice(d); cudaDeviceSynchronize(); }
      best = ms<best?ms:best;
    }
    printf("%-40s %.3f ms  %.0f GB/s (in-stream bytes)\n", names[mode], best, bytes/best/1e6);
  }
  return 0;
}
