// Fused depthwise causal conv (k = 4) + SiLU for both scan directions AND both directions' x_proj GEMMs.
//
//   [EXT] causal_conv1d_fn(x, w, b, activation="silu")      -> xc_f (taps t-3..t), xc_r (taps t..t+3)
//   [EXT] Mamba.x_proj: (dt, B, C) = xc W_x^T               -> dbc_f, dbc_r   ([T, RP], RP = R + 32 padded to 16)
//
// Unfused, x_proj re-reads the 2 x 1.07 GB of conv output per l32 layer that the conv kernel has just written; here the
// conv output tile is written to global memory (the scan needs it) and, from the same registers, into the shared-memory
// A operand of a tcgen05 GEMM, so x_proj costs no HBM traffic of its own.
//
// STATUS: correct (tests/test_ops_gpu.py::test_conv_xproj_fused: conv outputs bit-identical to conv_silu_kernel) but
// measured SLOWER than the unfused pair on B200 (l32, B = 256: 1.47 ms vs 0.77 + 2 x 0.175 ms): with the K loop inside
// the tile a thread gets only 16 rows per k-block, so the first loads of every k-block are exposed and the MUFU-bound conv
// part loses its streaming prefetch.  The forward therefore uses it only when PCAD_FUSED_CONV_XPROJ=1.
//
// CTA = 576 threads, persistent over tiles of 128 consecutive tokens of one sequence (L % 128 == 0), K loop over the
// E channels in blocks of 64:
//   warps 0-15 : producers (the conv + SiLU part is bound by the MUFU pipe, so it gets 4 warps per scheduler), two groups
//                of eight that take alternate k-blocks; in a group, warp w owns tile rows [16w, 16w + 16) and lane l owns
//                channels 2l, 2l+1 of the k-block: sliding 7-row window in registers, conv + SiLU for both directions,
//                4-byte stores to xc_f / xc_r, and 4-byte stores into the 128B-swizzled K-major A_f / A_r tiles of the
//                stage (exactly the layout TMA would have produced);
//   warp 16    : one lane streams the W_x k-blocks [RP x 64] of both directions with TMA;
//   warp 17    : one lane issues tcgen05.mma (128 x RP x 16, kind::f16) into two TMEM accumulators and commits;
//   warps 0-3  : after the K loop, tcgen05.ld the two accumulators, convert to bf16, store dbc_f / dbc_r.
#pragma once

#include "common.cuh"
#include "gemm_tcgen05.cuh"

namespace pcad {

constexpr int kCxTile = 128;     // tokens per tile (UMMA M)
constexpr int kCxBK = 64;        // channels per k-block (one 128-byte swizzle row of bf16)
constexpr int kCxStages = 3;
constexpr int kCxThreads = 576;
constexpr int kCxRows = 16;      // tile rows per producer warp
constexpr int kCxABytes = kCxTile * kCxBK * 2;   // 16 KB per direction

template <int RP>
struct CxCfg {
  static constexpr int kBBytes = RP * kCxBK * 2;                 // multiple of 1024 for RP % 8 == 0
  static constexpr int kStageBytes = 2 * kCxABytes + 2 * kBBytes;
  static constexpr int kSmemBytes = 1024 + kCxStages * kStageBytes + 256;
  static constexpr int kTmemCols = 256;                          // D_f at column 0, D_r at column 128
};

template <int RP>
__global__ void __launch_bounds__(kCxThreads, 1)
conv_xproj_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ w_f, const float* __restrict__ b_f,
                  const float* __restrict__ w_r, const float* __restrict__ b_r, bf16* __restrict__ xc_f,
                  bf16* __restrict__ xc_r, const __grid_constant__ CUtensorMap tmap_wf,
                  const __grid_constant__ CUtensorMap tmap_wr, bf16* __restrict__ dbc_f, bf16* __restrict__ dbc_r,
                  long long T, int L, int E) {
  using Cfg = CxCfg<RP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kCxStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kCxStages;
  uint64_t* acc_full = empty_bar + kCxStages;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long num_tiles = T / kCxTile;
  const int nkb = E / kCxBK;

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmap_wf);
    tma_prefetch_desc(&tmap_wr);
    for (int s = 0; s < kCxStages; ++s) {
      mbar_init(&full_bar[s], 256 + 1);   // the 256 producer threads of the group + the TMA lane's expect_tx arrive
      mbar_init(&empty_bar[s], 1);        // tcgen05.commit
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 128);
    fence_mbar_init();
  }
  if (warp == 17) {
    tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 16) {
    // ===================== producers (+ epilogue on warps 0-3) =====================
    const int grp = warp >> 3;        // takes k-blocks kb % 2 == grp
    const int wr8 = warp & 7;         // rows [16 wr8, 16 wr8 + 16) of the tile
    const int wq = warp & 3;          // TMEM lane quarter for the epilogue (warps 0-3)
    uint32_t acc_phase = 0;
    long long it = 0;                 // tiles processed by this CTA (k-block counter base = it * nkb)
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const long long row_base = tile * kCxTile;                 // first token row of the tile
      const int t_tile = static_cast<int>(row_base % L);         // position of that row inside its sequence
      const long long seq_row0 = row_base - t_tile;
      const int t0 = t_tile + kCxRows * wr8;                     // this warp's first timestep
      for (int kb = grp; kb < nkb; kb += 2) {
        const long long g = it * nkb + kb;                       // global k-block counter -> stage / phase
        const int stage = static_cast<int>(g % kCxStages);
        const uint32_t phase = static_cast<uint32_t>((g / kCxStages) & 1);
        const int e0 = kb * kCxBK + 2 * lane;
        float wf[2][4], wr[2][4], bf[2], br[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float4 a = *reinterpret_cast<const float4*>(w_f + (e0 + c) * 4);
          const float4 b = *reinterpret_cast<const float4*>(w_r + (e0 + c) * 4);
          wf[c][0] = a.x; wf[c][1] = a.y; wf[c][2] = a.z; wf[c][3] = a.w;
          wr[c][0] = b.x; wr[c][1] = b.y; wr[c][2] = b.z; wr[c][3] = b.w;
          bf[c] = b_f[e0 + c];
          br[c] = b_r[e0 + c];
        }
        const bf16* xcol = x + e0;
        auto fetch = [&](int t) -> uint32_t {   // two channels of row t of this sequence, zero outside [0, L)
          return (t >= 0 && t < L) ? *reinterpret_cast<const uint32_t*>(xcol + (seq_row0 + t) * ldx) : 0u;
        };
        float win[7][2];
        uint32_t nxt[8];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const uint32_t r = fetch(t0 - 3 + j);
          win[j][0] = __uint_as_float(r << 16);
          win[j][1] = __uint_as_float(r & 0xffff0000u);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) nxt[i] = fetch(t0 + 3 + i);
        // the MMAs that read this stage's previous contents must have completed
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* a_f = smem + stage * Cfg::kStageBytes;
        uint8_t* a_r = a_f + kCxABytes;
#pragma unroll 1
        for (int blk = 0; blk < kCxRows; blk += 8) {
          uint32_t cur[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
          if (blk + 8 < kCxRows) {
#pragma unroll
            for (int i = 0; i < 8; ++i) nxt[i] = fetch(t0 + blk + 8 + 3 + i);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rloc = kCxRows * wr8 + blk + i;     // row inside the tile
            win[6][0] = __uint_as_float(cur[i] << 16);
            win[6][1] = __uint_as_float(cur[i] & 0xffff0000u);
            float of[2], orv[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              float a = bf[c];
              a = fmaf(wf[c][0], win[0][c], a);
              a = fmaf(wf[c][1], win[1][c], a);
              a = fmaf(wf[c][2], win[2][c], a);
              a = fmaf(wf[c][3], win[3][c], a);
              of[c] = silu<false>(a);
              float r = br[c];
              r = fmaf(wr[c][0], win[6][c], r);
              r = fmaf(wr[c][1], win[5][c], r);
              r = fmaf(wr[c][2], win[4][c], r);
              r = fmaf(wr[c][3], win[3][c], r);
              orv[c] = silu<false>(r);
            }
            const uint32_t pf = pack_bf16x2(of[0], of[1]), pr = pack_bf16x2(orv[0], orv[1]);
            const long long grow = row_base + rloc;
            *reinterpret_cast<uint32_t*>(xc_f + grow * E + e0) = pf;
            *reinterpret_cast<uint32_t*>(xc_r + grow * E + e0) = pr;
            // K-major, 128-byte swizzle: row rloc at rloc * 128, 16-byte chunk (lane / 4) stored at chunk ^ (rloc & 7)
            const uint32_t off = static_cast<uint32_t>(rloc) * 128u + ((static_cast<uint32_t>(lane >> 2) ^ (rloc & 7)) << 4) +
                                 ((lane & 3) << 2);
            *reinterpret_cast<uint32_t*>(a_f + off) = pf;
            *reinterpret_cast<uint32_t*>(a_r + off) = pr;
#pragma unroll
            for (int j = 0; j < 6; ++j) { win[j][0] = win[j + 1][0]; win[j][1] = win[j + 1][1]; }
          }
        }
        fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core's async proxy
        mbar_arrive(&full_bar[stage]);
      }
      if (warp < 4) {
        // ---- epilogue: dbc_f / dbc_r rows of this tile (TMEM lane = tile row)
        mbar_wait(acc_full, acc_phase);
        tc_fence_after();
        const long long grow = row_base + 32 * wq + lane;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(32 * wq) << 16);
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
          bf16* dst = (dir ? dbc_r : dbc_f) + grow * RP;
#pragma unroll
          for (int c0 = 0; c0 < RP; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_row + dir * 128 + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (c0 + j * 8 < RP) {
                uint4 v;
                v.x = pack_bf16x2(__uint_as_float(r[j * 8 + 0]), __uint_as_float(r[j * 8 + 1]));
                v.y = pack_bf16x2(__uint_as_float(r[j * 8 + 2]), __uint_as_float(r[j * 8 + 3]));
                v.z = pack_bf16x2(__uint_as_float(r[j * 8 + 4]), __uint_as_float(r[j * 8 + 5]));
                v.w = pack_bf16x2(__uint_as_float(r[j * 8 + 6]), __uint_as_float(r[j * 8 + 7]));
                *reinterpret_cast<uint4*>(dst + c0 + j * 8) = v;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(acc_empty);
        acc_phase ^= 1;
      }
    }
  } else if (warp == 16) {
    // ===================== W_x loader (TMA) =====================
    if (elect_one()) {
      long long g = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int stage = static_cast<int>(g % kCxStages);
          const uint32_t phase = static_cast<uint32_t>((g / kCxStages) & 1);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* b_f_s = smem + stage * Cfg::kStageBytes + 2 * kCxABytes;
          uint8_t* b_r_s = b_f_s + Cfg::kBBytes;
          mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kBBytes);
          tma_load_2d(b_f_s, &tmap_wf, &full_bar[stage], kb * kCxBK, 0);
          tma_load_2d(b_r_s, &tmap_wr, &full_bar[stage], kb * kCxBK, 0);
        }
      }
    }
  } else {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kCxTile, RP);
      long long g = 0;
      uint32_t acc_phase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(acc_empty, acc_phase ^ 1);       // the epilogue has drained the previous tile's accumulators
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int stage = static_cast<int>(g % kCxStages);
          const uint32_t phase = static_cast<uint32_t>((g / kCxStages) & 1);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_f_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t a_r_addr = a_f_addr + kCxABytes;
          const uint32_t b_f_addr = a_r_addr + kCxABytes;
          const uint32_t b_r_addr = b_f_addr + Cfg::kBBytes;
          const uint64_t daf = make_smem_desc_sw128(a_f_addr), dar = make_smem_desc_sw128(a_r_addr);
          const uint64_t dbf = make_smem_desc_sw128(b_f_addr), dbr = make_smem_desc_sw128(b_r_addr);
#pragma unroll
          for (int k = 0; k < kCxBK / 16; ++k) {
            umma_bf16_ss(tmem_base, daf + 2 * k, dbf + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_bf16_ss(tmem_base + 128, dar + 2 * k, dbr + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
        }
        umma_commit(acc_full);
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int RP>
inline cudaError_t launch_conv_xproj_rp(const bf16* x, long long ldx, const float* w_f, const float* b_f, const float* w_r,
                                        const float* b_r, bf16* xc_f, bf16* xc_r, const CUtensorMap& twf,
                                        const CUtensorMap& twr, bf16* dbc_f, bf16* dbc_r, long long T, int L, int E,
                                        int num_sms, cudaStream_t stream) {
  using Cfg = CxCfg<RP>;
  static unsigned long long attr_done = 0;
  cudaError_t e = ensure_dynamic_smem(conv_xproj_kernel<RP>, Cfg::kSmemBytes, attr_done);
  if (e != cudaSuccess) return e;
  const long long tiles = T / kCxTile;
  const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
  conv_xproj_kernel<RP><<<grid, kCxThreads, Cfg::kSmemBytes, stream>>>(x, ldx, w_f, b_f, w_r, b_r, xc_f, xc_r, twf, twr, dbc_f,
                                                                     dbc_r, T, L, E);
  return cudaGetLastError();
}

// True if the fused kernel covers this shape; otherwise the caller runs conv_silu_kernel + two x_proj GEMMs.
inline bool conv_xproj_supported(int L, int E, int RP) {
  return L > 0 && (L % kCxTile) == 0 && (E % kCxBK) == 0 && (RP == 64 || RP == 80 || RP == 96);
}

// wx_f / wx_r: [RP, E] bf16 (row pitch E), rows beyond R + 2N zero.
inline cudaError_t conv_xproj_bf16(const bf16* x, long long ldx, const float* w_f, const float* b_f, const float* w_r,
                                   const float* b_r, bf16* xc_f, bf16* xc_r, const bf16* wx_f, const bf16* wx_r,
                                   bf16* dbc_f, bf16* dbc_r, long long T, int L, int E, int RP, int num_sms,
                                   cudaStream_t stream, const char** why) {
  *why = nullptr;
  if (T <= 0) return cudaSuccess;
  if (!conv_xproj_supported(L, E, RP) || (T % L) != 0) {
    *why = "conv_xproj: needs L % 128 == 0, E % 64 == 0 and RP in {64, 80, 96}";
    return cudaErrorInvalidValue;
  }
  CUtensorMap twf, twr;
  if (!make_tmap_bf16(&twf, wx_f, RP, E, E, RP) || !make_tmap_bf16(&twr, wx_r, RP, E, E, RP)) {
    *why = "conv_xproj: cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
  switch (RP) {
    case 64: return launch_conv_xproj_rp<64>(x, ldx, w_f, b_f, w_r, b_r, xc_f, xc_r, twf, twr, dbc_f, dbc_r, T, L, E, num_sms, stream);
    case 80: return launch_conv_xproj_rp<80>(x, ldx, w_f, b_f, w_r, b_r, xc_f, xc_r, twf, twr, dbc_f, dbc_r, T, L, E, num_sms, stream);
    default: return launch_conv_xproj_rp<96>(x, ldx, w_f, b_f, w_r, b_r, xc_f, xc_r, twf, twr, dbc_f, dbc_r, T, L, E, num_sms, stream);
  }
}

}  // namespace pcad
