// Bidirectional selective scan (Mamba-1, d_state = 16) with softplus(delta + bias), D skip and SiLU(z)
// gate; the forward-in-time and backward-in-time scans of BiMambaWrapper (strategy "add") run in the
// same thread and meet in the middle, so their sum is formed without a flipped copy and written once.
//
//   [EXT] mamba_ssm selective_scan_fn(u, delta, A, B, C, D, z, delta_bias, delta_softplus=True)
//   [EXT] Caduceus BiMambaWrapper.forward: mamba_fwd(u) + flip_L(mamba_rev(flip_L(u)))
//
// Work decomposition: one CTA = one sequence s x 128 channels; one thread = one channel, holding the
// 2 x 16 fp32 states and 2 x 16 (A * log2 e) coefficients in registers.  Step i advances the forward scan
// at t_f = i and the reverse scan at t_r = L-1-i.  Until the two meet, each direction parks its un-gated
// partial output in y; after they cross, a direction reads the other's partial (written earlier by this
// same thread), adds its own, applies the gate and stores the final value.
// Inputs are streamed in chunks of 16 timesteps through a 2-stage cp.async ring (u, delta for both
// directions, z when a chunk finalises, and the 32 B/C values per timestep, which are converted to fp32
// once per chunk and then broadcast-read by all 128 threads).
#pragma once

#include "common.cuh"

namespace pcad {

constexpr int kScanTC = 16;    // timesteps per chunk
constexpr int kScanCH = 128;   // channels (threads) per CTA
constexpr int kScanN = 16;     // d_state

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

  float hf[kScanN], hr[kScanN], af[kScanN], ar[kScanN];
  float Df = 0.f, Dr = 0.f, bf = 0.f, br = 0.f;
#pragma unroll
  for (int n = 0; n < kScanN; ++n) {
    hf[n] = 0.f;
    hr[n] = 0.f;
    const float a_f = active ? A_f[e * kScanN + n] : 0.f;
    const float a_r = active ? A_r[e * kScanN + n] : 0.f;
    af[n] = PRECISE ? a_f : a_f * kLog2e;
    ar[n] = PRECISE ? a_r : a_r * kLog2e;
  }
  if (active) { Df = D_f[e]; Dr = D_r[e]; bf = bias_f[e]; br = bias_r[e]; }

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
    // B|C to fp32, once per chunk
    for (int idx = tid; idx < kScanTC * 2 * kScanN; idx += kScanCH) {
      const int j = idx / (2 * kScanN), k = idx % (2 * kScanN);
      sm.bcf[j][k] = ActT<T>::to_f(s.bcf_raw[j][k]);
      sm.bcr[j][k] = ActT<T>::to_f(s.bcr_raw[j][k]);
    }
    __syncthreads();

    if (active) {
#pragma unroll 1
      for (int j = 0; j < kScanTC; ++j) {
        const int i = c * kScanTC + j;
        if (i >= L) break;
        const int tf = i, tr = L - 1 - i;
        // ---- forward direction at tf
        const float u1 = ActT<T>::to_f(s.uf[j][tid]);
        const float d1 = softplus<PRECISE>(ActT<T>::to_f(s.df[j][tid]) + bf);
        const float du1 = d1 * u1;
        float y1 = Df * u1;
        // ---- reverse direction at tr
        const float u2 = ActT<T>::to_f(s.ur[j][tid]);
        const float d2 = softplus<PRECISE>(ActT<T>::to_f(s.dr[j][tid]) + br);
        const float du2 = d2 * u2;
        float y2 = Dr * u2;
        const float4* bcf4 = reinterpret_cast<const float4*>(sm.bcf[j]);
        const float4* bcr4 = reinterpret_cast<const float4*>(sm.bcr[j]);
#pragma unroll
        for (int g = 0; g < kScanN / 4; ++g) {
          const float4 Bf = bcf4[g], Cf = bcf4[kScanN / 4 + g];
          const float4 Br = bcr4[g], Cr = bcr4[kScanN / 4 + g];
          const float bfv[4] = {Bf.x, Bf.y, Bf.z, Bf.w}, cfv[4] = {Cf.x, Cf.y, Cf.z, Cf.w};
          const float brv[4] = {Br.x, Br.y, Br.z, Br.w}, crv[4] = {Cr.x, Cr.y, Cr.z, Cr.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int n = g * 4 + k;
            const float dA1 = PRECISE ? expf(d1 * af[n]) : ex2_approx(d1 * af[n]);
            hf[n] = fmaf(dA1, hf[n], du1 * bfv[k]);
            y1 = fmaf(hf[n], cfv[k], y1);
            const float dA2 = PRECISE ? expf(d2 * ar[n]) : ex2_approx(d2 * ar[n]);
            hr[n] = fmaf(dA2, hr[n], du2 * brv[k]);
            y2 = fmaf(hr[n], crv[k], y2);
          }
        }
        // ---- combine: park partials before the directions meet, finalise after
        T* yf_ptr = y + (row0 + tf) * E + e;
        T* yr_ptr = y + (row0 + tr) * E + e;
        if (tf < tr) {
          *yf_ptr = ActT<T>::from_f(y1);
          *yr_ptr = ActT<T>::from_f(y2);
        } else if (tf == tr) {
          const float zz = ActT<T>::to_f(s.zf[j][tid]);
          *yf_ptr = ActT<T>::from_f((y1 + y2) * silu<PRECISE>(zz));
        } else {
          const float p1 = ActT<T>::to_f(*yf_ptr);  // reverse-direction partial parked at tf
          const float p2 = ActT<T>::to_f(*yr_ptr);  // forward-direction partial parked at tr
          const float z1 = ActT<T>::to_f(s.zf[j][tid]);
          const float z2 = ActT<T>::to_f(s.zr[j][tid]);
          *yf_ptr = ActT<T>::from_f((y1 + p1) * silu<PRECISE>(z1));
          *yr_ptr = ActT<T>::from_f((y2 + p2) * silu<PRECISE>(z2));
        }
      }
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
