// Microbenchmark: FP32 FMA issue/throughput on sm_100a for scalar FFMA, packed FFMA2, mixes, and MUFU.EX2 co-issue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_pipes fma_pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float a, float b){u64 r; asm("mov.b64 %0,{%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ u64 fma2(u64 a,u64 b,u64 c){u64 r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ float ffma(float a,float b,float c){float r; asm volatile("fma.rn.f32 %0,%1,%2,%3;":"=f"(r):"f"(a),"f"(b),"f"(c)); return r;}
__device__ __forceinline__ float ex2(float a){float r; asm volatile("ex2.approx.ftz.f32 %0,%1;":"=f"(r):"f"(a)); return r;}

// NS scalar chains, NP packed chains, NM mufu chains per thread; each loop iteration advances every chain once.
template<int NS,int NP,int NM>
__global__ void k(float* out, int iters, float seed){
  float s[NS>0?NS:1]; u64 p[NP>0?NP:1]; float m[NM>0?NM:1];
  for(int i=0;i<NS;i++) s[i]=seed+i;
  for(int i=0;i<NP;i++) p[i]=pack2(seed+i,seed-i);
  for(int i=0;i<NM;i++) m[i]=seed*0.001f*i;
  const float a=1.0000001f,b=1e-9f; const u64 a2=pack2(a,a),b2=pack2(b,b);
  for(int it=0;it<iters;++it){
#pragma unroll
    for(int i=0;i<NS;i++) s[i]=ffma(s[i],a,b);
#pragma unroll
    for(int i=0;i<NP;i++) p[i]=fma2(p[i],a2,b2);
#pragma unroll
    for(int i=0;i<NM;i++) m[i]=ex2(m[i]);
  }
  float acc=0; for(int i=0;i<NS;i++) acc+=s[i]; for(int i=0;i<NP;i++){acc+=__uint_as_float((unsigned)p[i]);} for(int i=0;i<NM;i++) acc+=m[i];
  if(acc==12345.678f) out[0]=acc;
}
template<int NS,int NP,int NM>
void run(const char* name, int warps_per_sm){
  int dev=0; cudaDeviceProp pr; cudaGetDeviceProperties(&pr,dev);
  float* out; cudaMalloc(&out,4);
  int iters=20000; int threads=warps_per_sm*32; int blocks=pr.multiProcessorCount;
  k<NS,NP,NM><<<blocks,threads>>>(out,100,1.f); cudaDeviceSynchronize();
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<NS,NP,NM><<<blocks,threads>>>(out,iters,1.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  int clk; cudaDeviceGetAttribute(&clk,cudaDevAttrClockRate,dev);
  double cycles=ms*1e-3*clk*1e3; // assuming max clock
  double warp_iters_per_smsp=(double)iters*warps_per_sm/4.0;
  double cyc_per_iter=cycles/warp_iters_per_smsp; // SMSP cycles per warp-iteration
  printf("%-28s warps/SM=%2d  NS=%2d NP=%2d NM=%2d : %.2f cyc per warp-iter per SMSP  => FMA lanes/clk/SMSP %.1f, MUFU lanes/clk/SMSP %.2f, inst/clk %.2f\n",
    name,warps_per_sm,NS,NP,NM,cyc_per_iter,(NS*32+NP*64)/cyc_per_iter,NM*32/cyc_per_iter,(NS+NP+NM)/cyc_per_iter);
  cudaFree(out);
}
int main(){
  for(int w: {8,16,32}){
    run<16,0,0>("scalar FFMA",w);
    run<0,16,0>("packed FFMA2",w);
    run<8,8,0>("8 scalar + 8 packed",w);
    run<8,16,0>("8 scalar + 16 packed",w);
    run<16,8,0>("16 scalar + 8 packed",w);
    run<0,0,8>("MUFU only",w);
    run<16,0,4>("16 scalar + 4 MUFU",w);
    run<0,16,4>("16 packed + 4 MUFU",w);
    run<0,16,8>("16 packed + 8 MUFU",w);
    run<0,12,6>("12 packed + 6 MUFU",w);
    run<8,12,6>("8s + 12 packed + 6 MUFU",w);
  }
  return 0;
}
