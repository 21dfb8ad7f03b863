// Microbenchmark: the selective-scan inner step (ScanDir<false>::step<NPOLY> of csrc/scan.cuh) in isolation --
// no global traffic, no chunk phases, no barriers -- for MUFU-flavoured warps, polynomial-flavoured warps and mixes.
// Reports SMSP cycles per warp-step (from clock64), to be compared with the pipe model (8 cyc per MUFU, 2 per FFMA2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I plantcaduceus_b200/csrc \
//        -o tools/ub/scan_step tools/ub/scan_step.cu
#include <cstdio>
#include "scan.cuh"
using namespace pcad;
#ifndef UB_BC_CONST
#define UB_BC_CONST 0   // 1: B|C rows come from __constant__ memory instead of shared memory (no LDS.128 in the step)
#endif
__constant__ __align__(16) float c_bc[16][32];

template <int NL, int NH, int WARPS, bool SOFTPLUS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(float* out, long long* cyc, int iters, int mod, int thr) {
  __shared__ __align__(16) float bc[16][32];
  __shared__ __align__(16) float ys[4][WARPS * 32];
  __shared__ __nv_bfloat16 dsm[16][256], usm[16][256];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16 * 32; i += blockDim.x) bc[i / 32][i % 32] = 0.01f * ((i * 7) % 13) - 0.05f;
  for (int j = 0; j < 16; ++j) {
    if (tid < 256) dsm[j][tid] = __float2bfloat16(0.1f * ((tid + j) % 7) - 0.3f);
    if (tid < 256) usm[j][tid] = __float2bfloat16(0.2f * ((tid * 3 + j) % 5) - 0.4f);
  }
  float A[16];
  for (int n = 0; n < 16; ++n) A[n] = -(n + 1.0f) * (1.0f + 1e-4f * (tid % 97)) * kLog2e;
  ScanDir<false> S;
  S.init(A, -4.0f);
  const bool heavy = mod > 0 && (warp % mod) < thr;
  __syncthreads();
  const long long t0 = clock64();
  auto run = [&](auto tag) {
    constexpr int NP = decltype(tag)::value;
    float uu = __bfloat162float(usm[0][tid & 255]);
    float dl = SOFTPLUS ? S.delta(__bfloat162float(dsm[0][tid & 255])) : S.delta_final(__bfloat162float(dsm[0][tid & 255]));
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
      for (int j = 0; j < 16; j += 4) {
        float yv[4];
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          const int jn = (j + k2 + 1) & 15;
          const float uu_n = __bfloat162float(usm[jn][tid & 255]);
          const float dr = __bfloat162float(dsm[jn][tid & 255]);
          const float dl_n = SOFTPLUS ? S.delta(dr) : S.delta_final(dr);
          yv[k2] = S.template step<NP>(dl, dl * uu, 0.5f * uu, UB_BC_CONST ? &c_bc[j + k2][0] : &bc[j + k2][0]);
          uu = uu_n; dl = dl_n;
        }
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) ys[k2][tid] = yv[k2];
      }
    }
  };
  if (heavy) run(IntTag<NH>()); else run(IntTag<NL>());
  const long long t1 = clock64();
  __syncthreads();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;   // warp 0's span; all warps run the same number of steps
  float acc = 0;
  for (int j = 0; j < 4; ++j) acc += ys[j][tid];
  if (acc == 12345.678f) out[0] = acc;
}

// ScanDir<false>::step() split in two for software pipelining: exps() only produces the 8 pair-decays of a step (all of
// its MUFU work), consume() only does the FMA-pipe work; the caller issues exps() of step j+1 before consume() of step j.
template <int NPOLY>
__device__ __forceinline__ void exps(const ScanDir<false>& S, float d, f32x2 (&dA)[kScanN / 2]) {
  const f32x2 dd = pack2(d, d);
  float dc = d;
  if (NPOLY > 0) dc = fminf(d, S.dclamp);
  const f32x2 ddc = pack2(dc, dc);
#pragma unroll
  for (int p = 0; p < kScanN / 2; ++p) {
    if (ScanDir<false>::is_poly<NPOLY>(p)) {
      dA[p] = exp2_prod_poly2(ddc, S.a[p]);
    } else {
      float x0, x1;
      unpack2(mul2(dd, S.a[p]), x0, x1);
      dA[p] = pack2(ex2_approx(x0), ex2_approx(x1));
    }
  }
}
__device__ __forceinline__ float consume(ScanDir<false>& S, const f32x2 (&dA)[kScanN / 2], float du, float y0, const float* bc) {
  const f32x2 duu = pack2(du, du);
  const ulonglong2* bc2 = reinterpret_cast<const ulonglong2*>(bc);
  f32x2 acc[2] = {pack2(y0, 0.f), pack2(0.f, 0.f)};
#pragma unroll
  for (int g = 0; g < kScanN / 4; ++g) {
    const ulonglong2 Bq = bc2[g], Cq = bc2[kScanN / 4 + g];
    const f32x2 Bp[2] = {Bq.x, Bq.y}, Cp[2] = {Cq.x, Cq.y};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int p = 2 * g + k;
      S.h[p] = fma2(dA[p], S.h[p], mul2(duu, Bp[k]));
      acc[k] = fma2(S.h[p], Cp[k], acc[k]);
    }
  }
  float s0, s1;
  unpack2(add2(acc[0], acc[1]), s0, s1);
  return s0 + s1;
}

// software-pipelined variant: MUFU work of step j+1 is issued before the FMA work of step j (16 more registers)
template <int NP, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) kp(float* out, long long* cyc, int iters) {
  __shared__ __align__(16) float bc[16][32];
  __shared__ __align__(16) float ys[4][WARPS * 32];
  __shared__ __nv_bfloat16 dsm[16][256], usm[16][256];
  const int tid = threadIdx.x;
  for (int i = tid; i < 16 * 32; i += blockDim.x) bc[i / 32][i % 32] = 0.01f * ((i * 7) % 13) - 0.05f;
  for (int j = 0; j < 16; ++j) {
    if (tid < 256) dsm[j][tid] = __float2bfloat16(0.1f * ((tid + j) % 7) - 0.3f);
    if (tid < 256) usm[j][tid] = __float2bfloat16(0.2f * ((tid * 3 + j) % 5) - 0.4f);
  }
  float A[16];
  for (int n = 0; n < 16; ++n) A[n] = -(n + 1.0f) * (1.0f + 1e-4f * (tid % 97)) * kLog2e;
  ScanDir<false> S;
  S.init(A, -4.0f);
  __syncthreads();
  const long long t0 = clock64();
  float uu = __bfloat162float(usm[0][tid & 255]);
  float dl = S.delta(__bfloat162float(dsm[0][tid & 255]));
  float dl1 = S.delta(__bfloat162float(dsm[1][tid & 255]));
  f32x2 dA[8];
  exps<NP>(S, dl, dA);
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int j = 0; j < 16; j += 4) {
      float yv[4];
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        const int jn = (j + k2 + 1) & 15, jn2 = (j + k2 + 2) & 15;
        const float uu_n = __bfloat162float(usm[jn][tid & 255]);
        const float dl2 = S.delta(__bfloat162float(dsm[jn2][tid & 255]));
        f32x2 dAn[8];
        exps<NP>(S, dl1, dAn);
        yv[k2] = consume(S, dA, dl * uu, 0.5f * uu, &bc[j + k2][0]);
#pragma unroll
        for (int p = 0; p < 8; ++p) dA[p] = dAn[p];
        uu = uu_n; dl = dl1; dl1 = dl2;
      }
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) ys[k2][tid] = yv[k2];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  float acc = 0;
  for (int j = 0; j < 4; ++j) acc += ys[j][tid];
  if (acc == 12345.678f) out[0] = acc;
}

template <int NP, int WARPS, int MINB>
void runp(const char* name) {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int blocks = pr.multiProcessorCount * MINB;
  float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, blocks * 8);
  const int iters = 400;
  kp<NP, WARPS, MINB><<<blocks, WARPS * 32>>>(out, cyc, 10);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kp<NP, WARPS, MINB><<<blocks, WARPS * 32>>>(out, cyc, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  long long* h = new long long[blocks]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
  const double steps_per_smsp = 16.0 * iters * WARPS * MINB / 4.0;
  printf("%-34s warps/SM=%2d poly=%d pipelined : %7.1f cyc per warp-step per SMSP  (%.3f ms, %s)\n", name, WARPS * MINB, NP,
         avg / steps_per_smsp, ms, cudaGetErrorString(err));
  cudaFree(out); cudaFree(cyc); delete[] h;
}

// two channels per thread sharing the B|C reads (the 2-CTA/SM layout sketched in DESIGN.md section 4): per step one set
// of 8 LDS.128 feeds two independent recurrences; 16 states x 2 + 16 A x 2 registers per thread -> 128-register budget
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k2ch(float* out, long long* cyc, int iters) {
  __shared__ __align__(16) float bc[16][32];
  __shared__ __align__(16) float ys[4][WARPS * 64];
  __shared__ __nv_bfloat16 dsm[16][256], usm[16][256];
  const int tid = threadIdx.x;
  for (int i = tid; i < 16 * 32; i += blockDim.x) bc[i / 32][i % 32] = 0.01f * ((i * 7) % 13) - 0.05f;
  for (int j = 0; j < 16; ++j)
    for (int i = tid; i < 256; i += blockDim.x) {
      dsm[j][i] = __float2bfloat16(0.1f * ((i + j) % 7) - 0.3f);
      usm[j][i] = __float2bfloat16(0.2f * ((i * 3 + j) % 5) - 0.4f);
    }
  float A[16];
  ScanDir<false> S0, S1;
  for (int n = 0; n < 16; ++n) A[n] = -(n + 1.0f) * (1.0f + 1e-4f * (tid % 97)) * kLog2e;
  S0.init(A, -4.0f);
  for (int n = 0; n < 16; ++n) A[n] *= 1.01f;
  S1.init(A, -4.1f);
  __syncthreads();
  const long long t0 = clock64();
  const int c2 = (tid * 2) & 255;
  auto ld2 = [&](const __nv_bfloat16 (*m)[256], int j, float& a, float& b) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&m[j][c2]);
    a = __uint_as_float(w << 16); b = __uint_as_float(w & 0xffff0000u);
  };
  float u0, u1, r0, r1;
  ld2(usm, 0, u0, u1); ld2(dsm, 0, r0, r1);
  float d0 = S0.delta(r0), d1 = S1.delta(r1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int j = 0; j < 16; j += 4) {
      float y0[4], y1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int jn = (j + k + 1) & 15;
        float un0, un1, rn0, rn1;
        ld2(usm, jn, un0, un1); ld2(dsm, jn, rn0, rn1);
        const float dn0 = S0.delta(rn0), dn1 = S1.delta(rn1);
        y0[k] = S0.template step<0>(d0, d0 * u0, 0.5f * u0, &bc[j + k][0]);
        y1[k] = S1.template step<0>(d1, d1 * u1, 0.5f * u1, &bc[j + k][0]);
        u0 = un0; u1 = un1; d0 = dn0; d1 = dn1;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<float2*>(&ys[k][tid * 2]) = make_float2(y0[k], y1[k]);
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  float acc = 0;
  for (int j = 0; j < 4; ++j) acc += ys[j][tid];
  if (acc == 12345.678f) out[0] = acc;
}

template <int WARPS, int MINB>
void run2ch(const char* name) {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int blocks = pr.multiProcessorCount * MINB;
  float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, blocks * 8);
  const int iters = 400;
  k2ch<WARPS, MINB><<<blocks, WARPS * 32>>>(out, cyc, 10);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k2ch<WARPS, MINB><<<blocks, WARPS * 32>>>(out, cyc, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  // channel-steps per SMSP: 2 per thread-step
  const double chsteps_per_smsp = 2.0 * 16.0 * iters * WARPS * MINB / 4.0;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-34s warps/SM=%2d 2 channels/thread : %.3f ms = %7.1f cyc per (warp, channel)-step per SMSP at %.2f GHz (%s)\n", name,
         WARPS * MINB, ms, ms * 1e-3 * clk * 1e3 / chsteps_per_smsp, clk / 1e6, cudaGetErrorString(err));
  cudaFree(out); cudaFree(cyc);
}

template <int NL, int NH, int WARPS, bool SOFTPLUS>
void run(const char* name, int mod, int thr) {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int blocks = pr.multiProcessorCount;
  float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, blocks * 8);
  const int iters = 400;   // x16 steps
  k<NL, NH, WARPS, SOFTPLUS><<<blocks, WARPS * 32>>>(out, cyc, 10, mod, thr);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<NL, NH, WARPS, SOFTPLUS><<<blocks, WARPS * 32>>>(out, cyc, iters, mod, thr);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  long long* h = new long long[blocks]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
  const double steps_per_smsp = 16.0 * iters * WARPS / 4.0;
  printf("%-34s warps/SM=%2d light=%d heavy=%d mix=%d/%d softplus=%d : %7.1f cyc per warp-step per SMSP  (%.3f ms, %s)\n", name, WARPS,
         NL, NH, thr, mod, (int)SOFTPLUS, avg / steps_per_smsp, ms, cudaGetErrorString(err));
  cudaFree(out); cudaFree(cyc); delete[] h;
}

int main() {
  printf("PCAD_SCAN_SPPOLY=%d PCAD_POLY_DEG=%d UB_BC_CONST=%d\n", PCAD_SCAN_SPPOLY, PCAD_POLY_DEG, UB_BC_CONST);
  {
    float hbc[16][32];
    for (int i = 0; i < 512; ++i) hbc[i / 32][i % 32] = 0.01f * ((i * 7) % 13) - 0.05f;
    cudaMemcpyToSymbol(c_bc, hbc, sizeof(hbc));
  }
  run<0, 8, 24, true>("all MUFU", 0, 0);
  run<0, 8, 24, false>("all MUFU, no softplus", 0, 0);
  run<1, 8, 24, true>("1 poly pair in every warp", 0, 0);
  run<2, 8, 24, true>("2 poly pairs in every warp", 0, 0);
  run<3, 8, 24, true>("3 poly pairs in every warp", 0, 0);
  run<0, 8, 24, true>("all polynomial", 1, 1);
  run<0, 8, 24, true>("1 of 5 warps polynomial", 5, 1);
  run<0, 6, 24, true>("1 of 3 warps 6-pair polynomial", 3, 1);
  run<0, 4, 24, true>("3 of 7 warps 4-pair", 7, 3);
  run<1, 4, 24, true>("1 pair everywhere, 3 of 7 warps 4-pair", 7, 3);
  run<1, 3, 24, true>("1 pair everywhere, 3 of 7 warps 3-pair", 7, 3);
  run2ch<16, 1>("2 ch/thread, 16 warps (128 regs)");
  run2ch<8, 2>("2 ch/thread, 2 x 8 warps");
  run2ch<12, 1>("2 ch/thread, 12 warps (168 regs)");
  run2ch<4, 4>("2 ch/thread, 4 x 4 warps");
  runp<0, 16, 1>("pipelined, 16 warps (128 regs)");
  runp<1, 16, 1>("pipelined, 16 warps (128 regs)");
  runp<2, 16, 1>("pipelined, 16 warps (128 regs)");
  runp<0, 8, 2>("pipelined, 2 x 8 warps (128 regs)");
  runp<1, 8, 2>("pipelined, 2 x 8 warps (128 regs)");
  runp<0, 20, 1>("pipelined, 20 warps (96 regs)");
  runp<1, 20, 1>("pipelined, 20 warps (96 regs)");
  runp<0, 24, 1>("pipelined, 24 warps (80 regs)");
  runp<1, 24, 1>("pipelined, 24 warps (80 regs)");
  run<0, 8, 16, true>("all MUFU, 16 warps", 0, 0);
  run<1, 8, 16, true>("1 poly pair, 16 warps", 0, 0);
  return 0;
}
