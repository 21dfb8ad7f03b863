// Bidirectional selective scan (Mamba-1, d_state = 16) with softplus(delta + bias), D skip and SiLU(z)
// gate; the forward-in-time and backward-in-time scans of BiMambaWrapper (strategy "add") run side by side in
// one CTA and meet in the middle, so their sum is formed without a flipped copy and y is written once.
//
//   [EXT] mamba_ssm selective_scan_fn(u, delta, A, B, C, D, z, delta_bias, delta_softplus=True)
//   [EXT] Caduceus BiMambaWrapper.forward: mamba_fwd(u) + flip_L(mamba_rev(flip_L(u)))
//
// Work decomposition: one CTA = one sequence x 128 channels; 256 threads: warps 0-3 run the forward scan,
// warps 4-7 the reverse scan, one thread = one (direction, channel) holding its 16 fp32 states and 16 A
// coefficients in registers.  (Keeping both directions in one thread needs ~126 registers, i.e. 4 warps per
// scheduler, and the kernel is latency-bound there; one direction per thread fits 6 warps per scheduler.)
// Step i advances the forward scan at t_f = i and the reverse scan at t_r = L-1-i.  Inputs are streamed in chunks
// of 16 steps through a 2-stage cp.async ring; the 32 B/C values per step are converted to fp32 once per chunk
// and broadcast-read.  Each thread leaves its un-gated y for the chunk in shared memory; a vectorised chunk
// epilogue then either parks the partial in y (the other direction has not reached that position yet) or adds the
// other direction's parked partial, applies SiLU(z) and stores the final value -- 16-byte global accesses only.
//
// The bf16 kernel is bound by the MUFU (ex2) and FMA pipes, not by HBM (ncu: profiles/).  Its inner loop
// (a) keeps (n, n+1) state pairs in 64-bit registers and uses the packed fp32x2 instructions of sm_100
// (FFMA2 / FMUL2: two lanes per issue slot), (b) can take kScanPoly of the 8 pair-exponentials from an FMA-pipe
// polynomial instead of MUFU.EX2 (default 0: measured on B200 at l32, B = 256: 0 -> 5.56 ms, 1 -> 5.57, 2 -> 5.91,
// 3 -> 6.34 per launch -- a polynomial exp costs as many FMA-pipe cycles as the MUFU cycles it saves and the two
// pipes share issue slots), (c) works in the log2 domain end to end: d' = log2(1 + 2^((delta + bias) log2 e)),
// exp(d A) = 2^(d' A), and the ln 2 that d = d' ln 2 owes to the input term is folded into B when B is converted.
#pragma once

#include <stdlib.h>

#include "common.cuh"

namespace pcad {

#ifndef PCAD_SCAN_TC
#define PCAD_SCAN_TC 16
#endif
constexpr int kScanTC = PCAD_SCAN_TC;      // timesteps per chunk
constexpr int kScanCH = 128;     // channels per CTA
constexpr int kScanThreads = 2 * kScanCH;
constexpr int kScanN = 16;       // d_state
#ifndef PCAD_SCAN_POLY
#define PCAD_SCAN_POLY 0
#endif
constexpr int kScanPoly = PCAD_SCAN_POLY;   // pairs (of 8) whose exp2 runs on the FMA pipe
#ifndef PCAD_SCAN_MINBLOCKS
#define PCAD_SCAN_MINBLOCKS 3
#endif

template <typename T>
struct ScanStage {
  T u[2][kScanTC][kScanCH];          // [direction][step][channel]
  T d[2][kScanTC][kScanCH];
  T bc_raw[2][kScanTC][2 * kScanN];
};

template <typename T>
struct ScanShared {
  ScanStage<T> st[2];
  float bc[2][kScanTC][2 * kScanN];   // fp32 B|C of the chunk being computed
  float ys[2][kScanTC][kScanCH];      // un-gated outputs of the chunk, per direction
  T pz[2][2][kScanTC][kScanCH];       // [partial | z][direction][step][channel]: prefetched for the chunk epilogue
};

// One direction's 16 states of one channel.  step() advances h <- exp(d*A) h + du*B and returns
// y0 + <C, h>;  bc points at this timestep's fp32 [B(16) | C(16)] row in shared memory (broadcast reads).
template <bool PRECISE> struct ScanDir;

template <> struct ScanDir<true> {
  float h[kScanN], a[kScanN];
  float bias;
  __device__ __forceinline__ void init(const float* A, float bias_) {
    bias = bias_;
#pragma unroll
    for (int n = 0; n < kScanN; ++n) { h[n] = 0.f; a[n] = A ? A[n] : 0.f; }
  }
  static __device__ __forceinline__ float b_scale() { return 1.0f; }
  // returns d (natural units)
  __device__ __forceinline__ float delta(float raw) const { return softplus<true>(raw + bias); }
  __device__ __forceinline__ float delta_final(float d) const { return d; }
  __device__ __forceinline__ float step(float d, float du, float y0, const float* bc) {
    float y = y0;
#pragma unroll
    for (int n = 0; n < kScanN; ++n) {
      h[n] = fmaf(expf(d * a[n]), h[n], du * bc[n]);
      y = fmaf(h[n], bc[kScanN + n], y);
    }
    return y;
  }
};

template <> struct ScanDir<false> {
  f32x2 h[kScanN / 2], a[kScanN / 2];   // a = A (log2 domain: multiplied by d' = d / ln 2)
  float bias_l2;   // bias * log2(e)
  __device__ __forceinline__ void init(const float* A, float bias_) {
    bias_l2 = bias_ * kLog2e;
#pragma unroll
    for (int p = 0; p < kScanN / 2; ++p) {
      h[p] = pack2(0.f, 0.f);
      a[p] = A ? pack2(A[2 * p], A[2 * p + 1]) : pack2(0.f, 0.f);
    }
  }
  static __device__ __forceinline__ float b_scale() { return kLn2; }   // B is pre-multiplied by ln 2
  // returns d' = softplus(raw + bias) / ln 2  (identity above 20, as the reference)
  __device__ __forceinline__ float delta(float raw) const {
    const float xl = fmaf(raw, kLog2e, bias_l2);
    const float sp = lg2_approx(1.0f + ex2_approx(xl));
    return xl > 20.0f * kLog2e ? xl : sp;
  }
  // delta already softplus'ed (dt_proj's softplus epilogue): only the change of units
  __device__ __forceinline__ float delta_final(float d) const { return d * kLog2e; }
  __device__ __forceinline__ float step(float d, float du, float y0, const float* bc) {
    const f32x2 dd = pack2(d, d), duu = pack2(du, du);
    const ulonglong2* bc2 = reinterpret_cast<const ulonglong2*>(bc);   // 16 bytes = two (n, n+1) pairs
    f32x2 acc[2] = {pack2(y0, 0.f), pack2(0.f, 0.f)};   // two chains: the FFMA2 -> FFMA2 latency is exposed otherwise
#pragma unroll
    for (int g = 0; g < kScanN / 4; ++g) {
      const ulonglong2 Bq = bc2[g], Cq = bc2[kScanN / 4 + g];
      const f32x2 Bp[2] = {Bq.x, Bq.y}, Cp[2] = {Cq.x, Cq.y};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int p = 2 * g + k;
        const f32x2 x2 = mul2(dd, a[p]);
        f32x2 dA;
        // the polynomial pairs are spread over the 8: p = 1, 4, 6, 3
        const bool poly = (kScanPoly >= 1 && p == 1) || (kScanPoly >= 2 && p == 4) || (kScanPoly >= 3 && p == 6) ||
                          (kScanPoly >= 4 && p == 3);
        if (poly) {
          dA = exp2_poly2(x2);
        } else {
          float x0, x1;
          unpack2(x2, x0, x1);
          dA = pack2(ex2_approx(x0), ex2_approx(x1));
        }
        h[p] = fma2(dA, h[p], mul2(duu, Bp[k]));
        acc[k] = fma2(h[p], Cp[k], acc[k]);
      }
    }
    float s0, s1;
    unpack2(add2(acc[0], acc[1]), s0, s1);
    return s0 + s1;
  }
};

// DFINAL: delta_* already hold softplus(dt_proj + bias) (the GEMM's softplus epilogue); bias_* are ignored.
template <typename T, bool PRECISE, bool DFINAL>
__global__ void __launch_bounds__(kScanThreads, PRECISE ? 1 : PCAD_SCAN_MINBLOCKS)
biscan_kernel(const T* __restrict__ u_f, const T* __restrict__ delta_f, const T* __restrict__ bc_f,
              const T* __restrict__ u_r, const T* __restrict__ delta_r, const T* __restrict__ bc_r, long long ldbc,
              int bc_off, const T* __restrict__ z, long long ldz, const float* __restrict__ A_f,
              const float* __restrict__ D_f, const float* __restrict__ bias_f, const float* __restrict__ A_r,
              const float* __restrict__ D_r, const float* __restrict__ bias_r, T* y, int L, int E) {
  extern __shared__ __align__(16) uint8_t scan_smem_raw[];
  ScanShared<T>& sm = *reinterpret_cast<ScanShared<T>*>(scan_smem_raw);

  const int tid = threadIdx.x;
  const int dir = tid >> 7;               // warp-uniform: warps 0-3 forward, 4-7 reverse
  const int ch = tid & (kScanCH - 1);
  const int e0 = blockIdx.x * kScanCH;
  const int e = e0 + ch;
  const bool active = e < E;
  const long long row0 = static_cast<long long>(blockIdx.y) * L;
  const int nch = (L + kScanTC - 1) / kScanTC;
  constexpr int VEC = 16 / sizeof(T);              // elements per 16-byte vector
  constexpr int SEGS = kScanCH / VEC;              // 16-byte segments per 128-channel row
  constexpr int BCSEGS = 2 * kScanN / VEC;         // 16-byte segments per B|C row

  // chunk c, stage row j: forward timestep 16c + j, reverse timestep L-1-(16c + j).
  auto issue = [&](int c, int stage) {
    ScanStage<T>& s = sm.st[stage];
    for (int idx = tid; idx < 2 * kScanTC * SEGS; idx += kScanThreads) {
      const int dd = idx / (kScanTC * SEGS);
      const int rem = idx - dd * (kScanTC * SEGS);
      const int j = rem / SEGS, seg = rem % SEGS;
      const int i = c * kScanTC + j;
      const int chn = e0 + seg * VEC;
      const bool ok = (i < L) && (chn < E);
      const long long r = row0 + (ok ? (dd ? (L - 1 - i) : i) : 0);
      const int chs = ok ? chn : 0;
      const int nb = ok ? 16 : 0;
      cp_async16(&s.u[dd][j][seg * VEC], (dd ? u_r : u_f) + r * E + chs, nb);
      cp_async16(&s.d[dd][j][seg * VEC], (dd ? delta_r : delta_f) + r * E + chs, nb);
    }
    for (int idx = tid; idx < 2 * kScanTC * BCSEGS; idx += kScanThreads) {
      const int dd = idx / (kScanTC * BCSEGS);
      const int rem = idx - dd * (kScanTC * BCSEGS);
      const int j = rem / BCSEGS, seg = rem % BCSEGS;
      const int i = c * kScanTC + j;
      const bool ok = i < L;
      const long long r = row0 + (ok ? (dd ? (L - 1 - i) : i) : 0);
      cp_async16(&s.bc_raw[dd][j][seg * VEC], (dd ? bc_r : bc_f) + r * ldbc + bc_off + seg * VEC, ok ? 16 : 0);
    }
  };

  ScanDir<PRECISE> S;
  {
    const float* A = dir ? A_r : A_f;
    const float* bias = dir ? bias_r : bias_f;
    S.init(active ? A + e * kScanN : nullptr, active ? bias[e] : 0.f);
  }
  const float Dskip = active ? (dir ? D_r : D_f)[e] : 0.f;
  const float bscale = ScanDir<PRECISE>::b_scale();

  issue(0, 0);
  cp_async_commit();
  for (int c = 0; c < nch; ++c) {
    const int stage = c & 1;
    if (c + 1 < nch) {
      issue(c + 1, stage ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();   // chunk c has landed; the previous chunk's epilogue is done with sm.ys / sm.pz
    const ScanStage<T>& s = sm.st[stage];
    const int i0 = c * kScanTC;
    const int nsteps = min(kScanTC, L - i0);
    // Positions of this chunk that the other direction visited in an EARLIER chunk will be finalised in the
    // epilogue: fetch their parked partials and z rows now (every earlier epilogue is complete and visible after
    // the barrier above), so that the epilogue does not wait on global memory.
    const bool has_final = (L - 1 - (i0 + nsteps - 1)) < i0 + nsteps;
    if (has_final) {
      for (int idx = tid; idx < 2 * kScanTC * SEGS; idx += kScanThreads) {
        const int dd = idx / (kScanTC * SEGS);
        const int rem = idx - dd * (kScanTC * SEGS);
        const int j = rem / SEGS, seg = rem % SEGS;
        const int i = i0 + j, io = L - 1 - i;
        const int chn = e0 + seg * VEC;
        const bool ok = (j < nsteps) && (chn < E) && (io < i0 + nsteps);
        const long long r = row0 + (ok ? (dd ? io : i) : 0);
        const int chs = ok ? chn : 0;
        if (ok && io < i0) cp_async16(&sm.pz[0][dd][j][seg * VEC], y + r * E + chs, 16);
        if (ok) cp_async16(&sm.pz[1][dd][j][seg * VEC], z + r * ldz + chs, 16);
      }
    }
    cp_async_commit();
    // B|C to fp32, once per chunk (B carries the ln 2 of the log2-domain delta on the fast path)
    for (int idx = tid; idx < 2 * kScanTC * 2 * kScanN; idx += kScanThreads) {
      const int dd = idx / (kScanTC * 2 * kScanN);
      const int rem = idx - dd * (kScanTC * 2 * kScanN);
      const int j = rem / (2 * kScanN), k = rem % (2 * kScanN);
      sm.bc[dd][j][k] = ActT<T>::to_f(s.bc_raw[dd][j][k]) * (k < kScanN ? bscale : 1.0f);
    }
    __syncthreads();

    if (active) {
      const T* up = &s.u[dir][0][ch];
      const T* dp = &s.d[dir][0][ch];
      const float* bcp = &sm.bc[dir][0][0];
      float* ysp = &sm.ys[dir][0][ch];
#pragma unroll 1
      for (int j = 0; j < nsteps; ++j) {
        const float uu = ActT<T>::to_f(up[j * kScanCH]);
        const float draw = ActT<T>::to_f(dp[j * kScanCH]);
        const float dl = DFINAL ? S.delta_final(draw) : S.delta(draw);
        ysp[j * kScanCH] = S.step(dl, dl * uu, Dskip * uu, bcp + j * 2 * kScanN);
      }
    }
    cp_async_wait<0>();
    __syncthreads();   // both directions' un-gated outputs of the chunk are in sm.ys, partials / z in sm.pz

    // ---- chunk epilogue: 16-byte vectors; item = (direction, step, segment of VEC channels)
    for (int idx = tid; idx < 2 * kScanTC * SEGS; idx += kScanThreads) {
      const int dd = idx / (kScanTC * SEGS);
      const int rem = idx - dd * (kScanTC * SEGS);
      const int j = rem / SEGS, seg = rem % SEGS;
      const int chn = e0 + seg * VEC;
      if (j >= nsteps || chn >= E) continue;
      const int i = i0 + j;                 // this direction's step index
      const int io = L - 1 - i;             // step at which the OTHER direction visits the same position
      const int t = dd ? io : i;            // the position itself
      float v[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] = sm.ys[dd][j][seg * VEC + k];
      T* yp = y + (row0 + t) * E + chn;
      if (io >= i0 + nsteps) {              // the other direction comes later: park the partial
        store16<T>(yp, v);
        continue;
      }
      if (io >= i0) {                       // both visits fall in this chunk: combine from shared memory, once
        if (io > i || (io == i && dd == 1)) continue;   // the later visitor (forward on a tie) writes
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] += sm.ys[dd ^ 1][io - i0][seg * VEC + k];
      } else {                              // parked in an earlier chunk (by another thread of this CTA)
        float p[VEC];
        load16<T>(&sm.pz[0][dd][j][seg * VEC], p);
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] += p[k];
      }
      float zz[VEC];
      load16<T>(&sm.pz[1][dd][j][seg * VEC], zz);
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] *= silu<PRECISE>(zz[k]);
      store16<T>(yp, v);
    }
    // no barrier here: the next iteration's first __syncthreads orders these reads of sm.ys and of the stage
    // against their next writers
  }
}

template <typename T, bool PRECISE, bool DFINAL>
inline cudaError_t launch_biscan(const T* u_f, const T* delta_f, const T* bc_f, const T* u_r, const T* delta_r,
                                 const T* bc_r, long long ldbc, int bc_off, const T* z, long long ldz,
                                 const float* A_f, const float* D_f, const float* bias_f, const float* A_r,
                                 const float* D_r, const float* bias_r, T* y, int S, int L, int E,
                                 cudaStream_t stream) {
  size_t smem = sizeof(ScanShared<T>);
  if (const char* ex = getenv("PCAD_SCAN_EXTRA_SMEM")) smem += static_cast<size_t>(atoi(ex));   // occupancy experiments
  static unsigned long long attr_done = 0;
  cudaError_t e1 = ensure_dynamic_smem(biscan_kernel<T, PRECISE, DFINAL>, static_cast<int>(smem), attr_done);
  if (e1 != cudaSuccess) return e1;
  dim3 grid((E + kScanCH - 1) / kScanCH, S);
  biscan_kernel<T, PRECISE, DFINAL><<<grid, kScanThreads, smem, stream>>>(u_f, delta_f, bc_f, u_r, delta_r, bc_r, ldbc, bc_off, z,
                                                                ldz, A_f, D_f, bias_f, A_r, D_r, bias_r, y, L, E);
  return cudaGetLastError();
}

}  // namespace pcad
