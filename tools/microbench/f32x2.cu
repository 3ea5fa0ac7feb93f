#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b){ unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b){ unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__global__ void k2(const unsigned long long* in, unsigned long long* out, int iters){
  unsigned long long acc[8]; unsigned long long x[8];
  for(int i=0;i<8;i++){acc[i]=0; x[i]=in[threadIdx.x*8+i];}
  unsigned long long c = in[1000];
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<8;i++) acc[i]=add2(acc[i], mul2(c, x[i]));
  }
  unsigned long long s=0; for(int i=0;i<8;i++) s^=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k1(const float* in, float* out, int iters){
  float acc[16]; float x[16];
  for(int i=0;i<16;i++){acc[i]=0; x[i]=in[threadIdx.x*16+i];}
  float c = in[2000];
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<16;i++) acc[i]=__fadd_rn(acc[i], __fmul_rn(c, x[i]));
  }
  float s=0; for(int i=0;i<16;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  float *in,*out; cudaMalloc(&in, 1<<20); cudaMalloc(&out, 1<<24); cudaMemset(in,0,1<<20);
  int iters=20000; int blocks=148*8, threads=256;
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); float ms;
  for(int rep=0;rep<2;rep++){
  k1<<<blocks,threads>>>(in,out,iters); cudaEventRecord(a); k1<<<blocks,threads>>>(in,out,iters); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
  printf("scalar: %.3f ms, %.2f T fp32 ops/s\n", ms, (double)blocks*threads*iters*32/ms/1e9);
  k2<<<blocks,threads>>>((unsigned long long*)in,(unsigned long long*)out,iters); cudaEventRecord(a); k2<<<blocks,threads>>>((unsigned long long*)in,(unsigned long long*)out,iters); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
  printf("f32x2 : %.3f ms, %.2f T fp32 ops/s\n", ms, (double)blocks*threads*iters*32/ms/1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
