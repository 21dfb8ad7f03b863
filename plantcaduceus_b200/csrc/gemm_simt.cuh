// fp32 C[M,N] = A[M,K] * W[N,K]^T on CUDA cores.  Parity mode only (north star: fp32 logits within
// 1e-4 relative of the reference); the bf16 production path uses gemm_tcgen05.cuh.
#pragma once

#include "common.cuh"

namespace pcad {

constexpr int kSimtBM = 64, kSimtBN = 64, kSimtBK = 16;

__global__ void __launch_bounds__(256)
gemm_f32_simt_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ C, long long M, int N,
                     int K, long long lda, long long ldw, long long ldc) {
  __shared__ float As[kSimtBK][kSimtBM + 4];
  __shared__ float Ws[kSimtBK][kSimtBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  const long long m0 = static_cast<long long>(blockIdx.y) * kSimtBM;
  const int n0 = blockIdx.x * kSimtBN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: 64 rows x 16 k = 1024 elements, 4 per thread
  const int lr = tid >> 2;        // row 0..63
  const int lk = (tid & 3) * 4;   // k 0,4,8,12
  for (int k0 = 0; k0 < K; k0 += kSimtBK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + lk + j;
      const long long am = m0 + lr;
      const int wn = n0 + lr;
      As[lk + j][lr] = (am < M && k < K) ? A[am * lda + k] : 0.f;
      Ws[lk + j][lr] = (wn < N && k < K) ? W[static_cast<long long>(wn) * ldw + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSimtBK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) C[m * ldc + n] = acc[i][j];
    }
  }
}

inline cudaError_t gemm_f32_simt(const float* A, const float* W, float* C, long long M, int N, int K, long long lda,
                                 long long ldw, long long ldc, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  dim3 grid((N + kSimtBN - 1) / kSimtBN, static_cast<unsigned>((M + kSimtBM - 1) / kSimtBM));
  gemm_f32_simt_kernel<<<grid, 256, 0, stream>>>(A, W, C, M, N, K, lda, ldw, ldc);
  return cudaGetLastError();
}

}  // namespace pcad
