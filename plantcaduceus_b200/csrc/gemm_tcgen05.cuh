// C[M,N] (bf16) = A[M,K] (bf16, K-major) * W[N,K]^T (bf16, K-major), fp32 accumulation in TMEM.
//
// Persistent, warp-specialised sm_100a kernel:
//   warp 0     : TMA producer (one elected lane) -- cp.async.bulk.tensor 2D tiles, 128B swizzle,
//                STAGES-deep smem ring guarded by full/empty mbarriers
//   warp 1     : TMEM allocator + MMA issuer (one elected lane) -- tcgen05.mma.cta_group::1.kind::f16,
//                128 x BN x 16 per instruction, accumulator double-buffered in TMEM
//   warps 2..5 : epilogue -- tcgen05.ld 32x32b.x32, fp32 -> bf16, 128B-swizzled staging slab per warp in shared
//                memory (double-buffered), written back with cp.async.bulk.tensor (TMA store; M/N tails are
//                clipped by the tensor map)
// This is the F.linear of Mamba.in_proj / x_proj / dt_proj / out_proj ([EXT] mamba_ssm Mamba.forward),
// which the reference runs through cuBLAS.
#pragma once

#include "common.cuh"

namespace pcad {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;
  static constexpr int kBBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // B tiles must keep 1024-byte alignment inside the ring: pad each B slot to a multiple of 1024.
  static constexpr int kBSlot = (kBBytes + 1023) / 1024 * 1024;
  static constexpr int kSlot = kABytes + kBSlot;
  static constexpr int kStages = (BN >= 256) ? 4 : 6;
  static constexpr int kAccStride = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kBarBytes = 256;
  // epilogue staging: 4 warps x 2 buffers x (32 rows x 64 bf16 = 4 KB), each slab 1024-byte aligned
  static constexpr int kEpiSlab = 32 * 64 * 2;
  static constexpr int kEpiBytes = 4 * 2 * kEpiSlab;
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kSlot + kEpiBytes + kBarBytes;
};

// Epilogues: kEpiPlain stores the accumulator; kEpiSoftplus stores softplus(acc + bias[col]) (identity above 20),
// which is dt_proj with the selective scan's delta_softplus / delta_bias step moved up into the GEMM
// ([EXT] selective_scan_fn(..., delta_bias, delta_softplus=True)): the scan is bound by the MUFU pipe, this
// write-bound GEMM has it idle.
enum { kEpiPlain = 0, kEpiSoftplus = 1 };

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_c, long long M, int N, int K,
                         const float* __restrict__ bias) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* epi = smem + STAGES * Cfg::kSlot;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi + Cfg::kEpiBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (N + BN - 1) / BN;
  const long long tiles_m = (M + kGemmBM - 1) / kGemmBM;
  const long long num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + kGemmBK - 1) / kGemmBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = static_cast<int>(tile / tiles_n) * kGemmBM;
        const int n0 = static_cast<int>(tile % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = ring + stage * Cfg::kSlot;
          uint8_t* b_dst = a_dst + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(a_dst, &tmap_a, &full_bar[stage], kb * kGemmBK, m0);
          tma_load_2d(b_dst, &tmap_b, &full_bar[stage], kb * kGemmBK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kGemmBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(ring + stage * Cfg::kSlot);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(b_addr);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in 16-byte units
            umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem slot once the MMAs above have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // epilogue warps 2..5; TMEM lane quarter is fixed by warp id % 4
    const int q = warp & 3;
    uint8_t* slab = epi + q * 2 * Cfg::kEpiSlab;
    int buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = static_cast<int>(tile / tiles_n) * kGemmBM;
      const int n0 = static_cast<int>(tile % tiles_n) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::kAccStride;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 64) {
        if (n0 + c0 >= N) break;  // warp-uniform
        uint8_t* dst = slab + buf * Cfg::kEpiSlab;
        // the TMA store that read this buffer two chunks ago must have drained it
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(t_row + c0, r0);
        if (c0 + 32 < BN) tmem_ld_32x32b_x32(t_row + c0 + 32, r1);
        tmem_ld_wait();
        if constexpr (EPI == kEpiSoftplus) {
#pragma unroll
          for (int g = 0; g < 16; ++g) {
            const int col = n0 + c0 + g * 4;
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col + 4 <= N) b4 = *reinterpret_cast<const float4*>(bias + col);   // N % 4 == 0 is checked by the host
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
            uint32_t* r = (g < 8) ? (r0 + g * 4) : (r1 + (g - 8) * 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) r[k] = __float_as_uint(softplus<false>(__uint_as_float(r[k]) + bb[k]));
          }
        }
        // row `lane` of the slab: 8 chunks of 16 bytes, chunk j stored at (j ^ (lane & 7)) -- the 128B swizzle
        uint8_t* row = dst + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t* r = (j < 4) ? (r0 + j * 8) : (r1 + (j - 4) * 8);
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(r[0]), __uint_as_float(r[1]));
          v.y = pack_bf16x2(__uint_as_float(r[2]), __uint_as_float(r[3]));
          v.z = pack_bf16x2(__uint_as_float(r[4]), __uint_as_float(r[5]));
          v.w = pack_bf16x2(__uint_as_float(r[6]), __uint_as_float(r[7]));
          *reinterpret_cast<uint4*>(row + ((j ^ (lane & 7)) << 4)) = v;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap_c, dst, n0 + c0, m0 + q * 32);
          tma_store_commit();
        }
        buf ^= 1;
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2D bf16 row-major [rows, cols] with row pitch ld (elements); box = [box_rows, 64 cols], 128B swizzle.
// Loads (A, W) zero-fill out-of-range elements; stores (C) clip them.
inline bool make_tmap_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld,
                           int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kGemmBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

inline int pick_bn(int N) {
  if (N <= 64) return 64;
  if (N <= 80) return 80;
  if (N <= 96) return 96;
  if (N <= 128) return 128;
  const int waste256 = (N + 255) / 256 * 256 - N;
  const int waste128 = (N + 127) / 128 * 128 - N;
  return waste128 < waste256 ? 128 : 256;
}

template <int BN, int EPI>
inline cudaError_t launch_gemm_bn(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, long long M, int N,
                                  int K, const float* bias, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const long long tiles = ((M + kGemmBM - 1) / kGemmBM) * ((N + BN - 1) / BN);
  const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
  gemm_bf16_tcgen05_kernel<BN, EPI><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tc, M, N, K, bias);
  return cudaGetLastError();
}

// Returns cudaSuccess or an error; *why is set for non-CUDA failures.  bias != nullptr selects the softplus
// epilogue: C = softplus(A W^T + bias).
inline cudaError_t gemm_bf16_tcgen05(const bf16* A, const bf16* W, bf16* C, long long M, int N, int K, long long lda,
                                     long long ldw, long long ldc, int num_sms, cudaStream_t stream,
                                     const char** why, const float* bias = nullptr) {
  *why = nullptr;
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  if ((lda % 8) || (ldw % 8) || (ldc % 8) || (reinterpret_cast<uintptr_t>(A) & 15) ||
      (reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(C) & 15)) {
    *why = "gemm: pointers must be 16-byte aligned and row pitches multiples of 8 elements";
    return cudaErrorInvalidValue;
  }
  if (bias && ((N % 4) || (reinterpret_cast<uintptr_t>(bias) & 15))) {
    *why = "gemm: the softplus epilogue needs N % 4 == 0 and a 16-byte aligned bias";
    return cudaErrorInvalidValue;
  }
  const int BN = pick_bn(N);
  CUtensorMap ta, tb, tc;
  if (!make_tmap_bf16(&ta, A, M, K, lda, kGemmBM) || !make_tmap_bf16(&tb, W, N, K, ldw, BN) ||
      !make_tmap_bf16(&tc, C, M, N, ldc, 32)) {
    *why = "gemm: cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
#define PCAD_GEMM_CASE(BNV)                                                                                      \
  case BNV:                                                                                                      \
    return bias ? launch_gemm_bn<BNV, kEpiSoftplus>(ta, tb, tc, M, N, K, bias, num_sms, stream)                  \
                : launch_gemm_bn<BNV, kEpiPlain>(ta, tb, tc, M, N, K, nullptr, num_sms, stream);
  switch (BN) {
    PCAD_GEMM_CASE(64)
    PCAD_GEMM_CASE(80)
    PCAD_GEMM_CASE(96)
    PCAD_GEMM_CASE(128)
    default:
      return bias ? launch_gemm_bn<256, kEpiSoftplus>(ta, tb, tc, M, N, K, bias, num_sms, stream)
                  : launch_gemm_bn<256, kEpiPlain>(ta, tb, tc, M, N, K, nullptr, num_sms, stream);
  }
#undef PCAD_GEMM_CASE
}

}  // namespace pcad
