// Bidirectional selective scan (Mamba-1, d_state = 16) with softplus(delta + bias), D skip and SiLU(z)
// gate; the forward-in-time and backward-in-time scans of BiMambaWrapper (strategy "add") run in the
// same thread and meet in the middle, so their sum is formed without a flipped copy and written once.
//
//   [EXT] mamba_ssm selective_scan_fn(u, delta, A, B, C, D, z, delta_bias, delta_softplus=True)
//   [EXT] Caduceus BiMambaWrapper.forward: mamba_fwd(u) + flip_L(mamba_rev(flip_L(u)))
//
// Work decomposition: one CTA = one sequence s x 128 channels; one thread = one channel, holding the
// 2 x 16 fp32 states and the 2 x 16 A coefficients in registers.  Step i advances the forward scan
// at t_f = i and the reverse scan at t_r = L-1-i.  Until the two meet, each direction parks its un-gated
// partial output in y; after they cross, a direction reads the other's partial (written earlier by this
// same thread), adds its own, applies the gate and stores the final value.
// Inputs are streamed in chunks of 16 timesteps through a 2-stage cp.async ring (u, delta for both
// directions, z when a chunk finalises, and the 32 B/C values per timestep, which are converted to fp32
// once per chunk and then broadcast-read by all 128 threads).
//
// The bf16 kernel is bound by the MUFU (ex2) pipe and by issue slots, not by HBM (ncu: profiles/).  Its
// inner loop therefore (a) keeps (n, n+1) state pairs in 64-bit registers and uses the packed fp32x2
// instructions of sm_100 (FFMA2 / FMUL2: two lanes per issue slot), (b) takes kScanPoly of the 8
// pair-exponentials per direction from an FMA-pipe polynomial instead of MUFU.EX2, (c) works in the
// log2 domain end to end: d' = log2(1 + 2^((delta + bias) log2 e)), exp(d A) = 2^(d' A), and the ln 2 that
// d = d' ln 2 owes to the input term is folded into B when B is converted to fp32.
#pragma once

#include "common.cuh"

namespace pcad {

constexpr int kScanTC = 16;    // timesteps per chunk
constexpr int kScanCH = 128;   // channels (threads) per CTA
constexpr int kScanN = 16;     // d_state
#ifndef PCAD_SCAN_POLY
#define PCAD_SCAN_POLY 3
#endif
constexpr int kScanPoly = PCAD_SCAN_POLY;   // pairs (of 8) per direction whose exp2 runs on the FMA pipe

template <typename T>
struct ScanSmem {
  // per stage
  T uf[kScanTC][kScanCH];
  T df[kScanTC][kScanCH];
  T ur[kScanTC][kScanCH];
  T dr[kScanTC][kScanCH];
  T zf[kScanTC][kScanCH];
  T zr[kScanTC][kScanCH];
  T bcf_raw[kScanTC][2 * kScanN];
  T bcr_raw[kScanTC][2 * kScanN];
};

template <typename T>
struct ScanShared {
  ScanSmem<T> st[2];
  float bcf[kScanTC][2 * kScanN];  // fp32 B|C of the chunk being computed
  float bcr[kScanTC][2 * kScanN];
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
  f32x2 h[kScanN / 2], a[kScanN / 2];
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

enum ScanMode { kScanPark = 0, kScanFinal = 1, kScanMixed = 2 };

// One chunk of up to kScanTC steps.  PARK: every step has t_f < t_r (store un-gated partials).
// FINAL: every step has t_f > t_r (add the parked partial of the other direction, gate, store).
// MIXED: decide per step (the chunk that contains the meeting point when it is not chunk-aligned).
template <typename T, bool PRECISE, int MODE>
__device__ __forceinline__ void scan_chunk(ScanDir<PRECISE>& Sf, ScanDir<PRECISE>& Sr, const ScanSmem<T>& s,
                                           const float (*bcf)[2 * kScanN], const float (*bcr)[2 * kScanN], float Df,
                                           float Dr, T* yf_ptr, T* yr_ptr, long long E, int i0, int nsteps, int L,
                                           int tid) {
  // FINAL: the other direction's parked partials are fetched one step ahead and kept as raw bits; the
  // conversion sits at the point of use so the load has a whole step to land.
  T p1r = T(), p2r = T();
  if (MODE == kScanFinal) {
    p1r = *yf_ptr;
    p2r = *yr_ptr;
  }
#pragma unroll 1
  for (int j = 0; j < nsteps; ++j) {
    T p1n = T(), p2n = T();
    if (MODE == kScanFinal && j + 1 < nsteps) {
      p1n = yf_ptr[E];
      p2n = *(yr_ptr - E);
    }
    const float u1 = ActT<T>::to_f(s.uf[j][tid]);
    const float d1 = Sf.delta(ActT<T>::to_f(s.df[j][tid]));
    const float u2 = ActT<T>::to_f(s.ur[j][tid]);
    const float d2 = Sr.delta(ActT<T>::to_f(s.dr[j][tid]));
    const float y1 = Sf.step(d1, d1 * u1, Df * u1, bcf[j]);
    const float y2 = Sr.step(d2, d2 * u2, Dr * u2, bcr[j]);
    if (MODE == kScanPark) {
      *yf_ptr = ActT<T>::from_f(y1);
      *yr_ptr = ActT<T>::from_f(y2);
    } else if (MODE == kScanFinal) {
      const float z1 = ActT<T>::to_f(s.zf[j][tid]);
      const float z2 = ActT<T>::to_f(s.zr[j][tid]);
      *yf_ptr = ActT<T>::from_f((y1 + ActT<T>::to_f(p1r)) * silu<PRECISE>(z1));
      *yr_ptr = ActT<T>::from_f((y2 + ActT<T>::to_f(p2r)) * silu<PRECISE>(z2));
      p1r = p1n;
      p2r = p2n;
    } else {
      const int tf = i0 + j, tr = L - 1 - tf;
      if (tf < tr) {
        *yf_ptr = ActT<T>::from_f(y1);
        *yr_ptr = ActT<T>::from_f(y2);
      } else if (tf == tr) {
        const float zz = ActT<T>::to_f(s.zf[j][tid]);
        *yf_ptr = ActT<T>::from_f((y1 + y2) * silu<PRECISE>(zz));
      } else {
        const float q1 = ActT<T>::to_f(*yf_ptr);  // reverse-direction partial parked at tf
        const float q2 = ActT<T>::to_f(*yr_ptr);  // forward-direction partial parked at tr
        const float z1 = ActT<T>::to_f(s.zf[j][tid]);
        const float z2 = ActT<T>::to_f(s.zr[j][tid]);
        *yf_ptr = ActT<T>::from_f((y1 + q1) * silu<PRECISE>(z1));
        *yr_ptr = ActT<T>::from_f((y2 + q2) * silu<PRECISE>(z2));
      }
    }
    yf_ptr += E;
    yr_ptr -= E;
  }
}

template <typename T, bool PRECISE>
__global__ void __launch_bounds__(kScanCH)
biscan_kernel(const T* __restrict__ u_f, const T* __restrict__ delta_f, const T* __restrict__ bc_f,
              const T* __restrict__ u_r, const T* __restrict__ delta_r, const T* __restrict__ bc_r, long long ldbc,
              int bc_off, const T* __restrict__ z, long long ldz, const float* __restrict__ A_f,
              const float* __restrict__ D_f, const float* __restrict__ bias_f, const float* __restrict__ A_r,
              const float* __restrict__ D_r, const float* __restrict__ bias_r, T* y, int L, int E) {
  extern __shared__ __align__(16) uint8_t scan_smem_raw[];
  ScanShared<T>& sm = *reinterpret_cast<ScanShared<T>*>(scan_smem_raw);

  const int tid = threadIdx.x;
  const int e0 = blockIdx.x * kScanCH;
  const int e = e0 + tid;
  const bool active = e < E;
  const long long row0 = static_cast<long long>(blockIdx.y) * L;
  const int nch = (L + kScanTC - 1) / kScanTC;
  constexpr int VEC = 16 / sizeof(T);              // elements per 16-byte cp.async
  constexpr int SEGS = kScanCH / VEC;              // 16-byte segments per 128-channel row
  constexpr int BCSEGS = 2 * kScanN / VEC;         // 16-byte segments per B|C row

  // chunk c: forward rows [16c, 16c+16), reverse rows [L-16c-16, L-16c); row j of the stage buffers holds
  // forward timestep 16c + j and reverse timestep L-1-(16c + j).
  auto issue = [&](int c, int stage) {
    ScanSmem<T>& s = sm.st[stage];
    const bool fin = (2 * kScanTC * c + 2 * kScanTC - 1 >= L - 1);  // chunk may finalise -> needs z
    for (int idx = tid; idx < kScanTC * SEGS; idx += kScanCH) {
      const int j = idx / SEGS, seg = idx % SEGS;
      const int i = c * kScanTC + j;
      const int ch = e0 + seg * VEC;
      const bool ok = (i < L) && (ch < E);
      const long long rf = row0 + (ok ? i : 0);
      const long long rr = row0 + (ok ? (L - 1 - i) : 0);
      const int chs = ok ? ch : 0;
      const int nb = ok ? 16 : 0;
      cp_async16(&s.uf[j][seg * VEC], u_f + rf * E + chs, nb);
      cp_async16(&s.df[j][seg * VEC], delta_f + rf * E + chs, nb);
      cp_async16(&s.ur[j][seg * VEC], u_r + rr * E + chs, nb);
      cp_async16(&s.dr[j][seg * VEC], delta_r + rr * E + chs, nb);
      if (fin) {
        cp_async16(&s.zf[j][seg * VEC], z + rf * ldz + chs, nb);
        cp_async16(&s.zr[j][seg * VEC], z + rr * ldz + chs, nb);
      }
    }
    for (int idx = tid; idx < kScanTC * BCSEGS; idx += kScanCH) {
      const int j = idx / BCSEGS, seg = idx % BCSEGS;
      const int i = c * kScanTC + j;
      const bool ok = i < L;
      const long long rf = row0 + (ok ? i : 0);
      const long long rr = row0 + (ok ? (L - 1 - i) : 0);
      const int nb = ok ? 16 : 0;
      cp_async16(&s.bcf_raw[j][seg * VEC], bc_f + rf * ldbc + bc_off + seg * VEC, nb);
      cp_async16(&s.bcr_raw[j][seg * VEC], bc_r + rr * ldbc + bc_off + seg * VEC, nb);
    }
  };

  ScanDir<PRECISE> Sf, Sr;
  Sf.init(active ? A_f + e * kScanN : nullptr, active ? bias_f[e] : 0.f);
  Sr.init(active ? A_r + e * kScanN : nullptr, active ? bias_r[e] : 0.f);
  const float Df = active ? D_f[e] : 0.f, Dr = active ? D_r[e] : 0.f;
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
    __syncthreads();
    ScanSmem<T>& s = sm.st[stage];
    // B|C to fp32, once per chunk (B carries the ln 2 of the log2-domain delta on the fast path)
    for (int idx = tid; idx < kScanTC * 2 * kScanN; idx += kScanCH) {
      const int j = idx / (2 * kScanN), k = idx % (2 * kScanN);
      const float sc = k < kScanN ? bscale : 1.0f;
      sm.bcf[j][k] = ActT<T>::to_f(s.bcf_raw[j][k]) * sc;
      sm.bcr[j][k] = ActT<T>::to_f(s.bcr_raw[j][k]) * sc;
    }
    __syncthreads();

    if (active) {
      const int i0 = c * kScanTC;
      const int nsteps = min(kScanTC, L - i0);
      T* yf_ptr = y + (row0 + i0) * E + e;
      T* yr_ptr = y + (row0 + (L - 1 - i0)) * E + e;
      const int ilast = i0 + nsteps - 1;
      if (2 * ilast < L - 1)
        scan_chunk<T, PRECISE, kScanPark>(Sf, Sr, s, sm.bcf, sm.bcr, Df, Dr, yf_ptr, yr_ptr, E, i0, nsteps, L, tid);
      else if (2 * i0 > L - 1)
        scan_chunk<T, PRECISE, kScanFinal>(Sf, Sr, s, sm.bcf, sm.bcr, Df, Dr, yf_ptr, yr_ptr, E, i0, nsteps, L, tid);
      else
        scan_chunk<T, PRECISE, kScanMixed>(Sf, Sr, s, sm.bcf, sm.bcr, Df, Dr, yf_ptr, yr_ptr, E, i0, nsteps, L, tid);
    }
    __syncthreads();  // everyone is done with this stage (and sm.bc*) before it is refilled
  }
}

template <typename T, bool PRECISE>
inline cudaError_t launch_biscan(const T* u_f, const T* delta_f, const T* bc_f, const T* u_r, const T* delta_r,
                                 const T* bc_r, long long ldbc, int bc_off, const T* z, long long ldz,
                                 const float* A_f, const float* D_f, const float* bias_f, const float* A_r,
                                 const float* D_r, const float* bias_r, T* y, int S, int L, int E,
                                 cudaStream_t stream) {
  const size_t smem = sizeof(ScanShared<T>);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(biscan_kernel<T, PRECISE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem));
    if (e1 != cudaSuccess) return e1;
    attr_set = true;
  }
  dim3 grid((E + kScanCH - 1) / kScanCH, S);
  biscan_kernel<T, PRECISE><<<grid, kScanCH, smem, stream>>>(u_f, delta_f, bc_f, u_r, delta_r, bc_r, ldbc, bc_off, z,
                                                           ldz, A_f, D_f, bias_f, A_r, D_r, bias_r, y, L, E);
  return cudaGetLastError();
}

}  // namespace pcad
