// Final layer of the EEG towers (clip_model.py:439,472: nn.Linear(64 -> latent), latent = 8 by default) and its gradients.
// N = 8 is below every tensor-core tile and the generic fp32 GEMM spent ~95 us per launch on it (3 launches per step); the
// work is 21 MB of activations against 0.3 MFLOP per token, i.e. HBM-bound streaming kernels:
//   skinny_fwd    out[m][n]  = b[n] + sum_k x[m][k] W[n][k]          warp = 32 tokens staged through shared memory
//   skinny_dgrad  dx[m][k]   = sum_n dy[m][n] W[n][k]                thread = (token, 4 features), coalesced 128-bit stores
//   skinny_wgrad  dW[n][k]   = sum_m dy[m][n] x[m][k], db[n] = sum_m dy[m][n]   per-CTA partials over a token slice, summed in
//                 fixed order by lin_wgrad_reduce_kernel (deterministic, like every other weight gradient here)
// Exact fp32 FMA.  K = 64 (the towers' embedding width), NL = latent in {4, 8}.
#pragma once
#include "common.cuh"
#include "lin_tc.cuh"

namespace eegclip {
namespace skinny {

constexpr int SK = 64;
constexpr int WG_CTAS = 296;

template <int NL>
__global__ void __launch_bounds__(128) skinny_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                        const float* __restrict__ b, float* __restrict__ out, long M) {
  pdl_sync();
  __shared__ __align__(16) float sWt[SK][NL];          // W transposed: one 128-bit broadcast load gives 4 outputs of a feature
  __shared__ float sx[4][32][SK + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < NL * SK; i += 128) sWt[i % SK][i / SK] = W[i];
  __syncthreads();
  const long m0 = ((long)blockIdx.x * 4 + warp) * 32;
  if (m0 >= M) return;
  {
    const int c4 = (lane & 15) * 4, rsub = lane >> 4;
    float4 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const long m = m0 + j * 2 + rsub;
      v[j] = m < M ? __ldg(reinterpret_cast<const float4*>(x + m * SK + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float* d = &sx[warp][j * 2 + rsub][c4];
      d[0] = v[j].x; d[1] = v[j].y; d[2] = v[j].z; d[3] = v[j].w;
    }
  }
  __syncwarp();
  float acc[NL];
#pragma unroll
  for (int n = 0; n < NL; ++n) acc[n] = b ? __ldg(b + n) : 0.f;
#pragma unroll 8
  for (int k = 0; k < SK; ++k) {
    const float xv = sx[warp][lane][k];
#pragma unroll
    for (int n4 = 0; n4 < NL; n4 += 4) {
      const float4 w = *reinterpret_cast<const float4*>(&sWt[k][n4]);
      acc[n4] = fmaf(xv, w.x, acc[n4]); acc[n4 + 1] = fmaf(xv, w.y, acc[n4 + 1]);
      acc[n4 + 2] = fmaf(xv, w.z, acc[n4 + 2]); acc[n4 + 3] = fmaf(xv, w.w, acc[n4 + 3]);
    }
  }
  const long m = m0 + lane;
  if (m < M) {
#pragma unroll
    for (int n4 = 0; n4 < NL; n4 += 4)
      *reinterpret_cast<float4*>(out + m * NL + n4) = make_float4(acc[n4], acc[n4 + 1], acc[n4 + 2], acc[n4 + 3]);
  }
}

template <int NL>
__global__ void __launch_bounds__(256) skinny_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ W,
                                                          float* __restrict__ dx, long M) {
  pdl_sync();
  __shared__ __align__(16) float sW[NL][SK];
  for (int i = threadIdx.x; i < NL * SK; i += 256) sW[i / SK][i % SK] = W[i];
  __syncthreads();
  const long i = (long)blockIdx.x * 256 + threadIdx.x;   // over M * 16
  const long m = i >> 4;
  const int k4 = (int)(i & 15) * 4;
  if (m >= M) return;
  float g[NL];
#pragma unroll
  for (int n4 = 0; n4 < NL; n4 += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(dy + m * NL + n4));
    g[n4] = v.x; g[n4 + 1] = v.y; g[n4 + 2] = v.z; g[n4 + 3] = v.w;
  }
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int n = 0; n < NL; ++n) {
    const float4 w = *reinterpret_cast<const float4*>(&sW[n][k4]);
    r.x = fmaf(g[n], w.x, r.x); r.y = fmaf(g[n], w.y, r.y); r.z = fmaf(g[n], w.z, r.z); r.w = fmaf(g[n], w.w, r.w);
  }
  *reinterpret_cast<float4*>(dx + m * SK + k4) = r;
}

// partial[cta][NL * 64 + NL]: thread (k4 = tid & 15, slot = tid >> 4) walks tokens slot, slot + 16, ... of the CTA's slice
template <int NL>
__global__ void __launch_bounds__(256) skinny_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                          float* __restrict__ partial, long M) {
  pdl_sync();
  __shared__ float red[16][NL * SK + NL];
  const int tid = threadIdx.x, k4 = (tid & 15) * 4, slot = tid >> 4;
  const long per = (M + gridDim.x - 1) / gridDim.x;
  const long beg = (long)blockIdx.x * per, end = beg + per < M ? beg + per : M;
  float acc[NL][4], accb[NL];
#pragma unroll
  for (int n = 0; n < NL; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; accb[n] = 0.f; }
  for (long m = beg + slot; m < end; m += 16) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + m * SK + k4));
    float g[NL];
#pragma unroll
    for (int n4 = 0; n4 < NL; n4 += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(dy + m * NL + n4));
      g[n4] = v.x; g[n4 + 1] = v.y; g[n4 + 2] = v.z; g[n4 + 3] = v.w;
    }
#pragma unroll
    for (int n = 0; n < NL; ++n) {
      acc[n][0] = fmaf(g[n], xv.x, acc[n][0]); acc[n][1] = fmaf(g[n], xv.y, acc[n][1]);
      acc[n][2] = fmaf(g[n], xv.z, acc[n][2]); acc[n][3] = fmaf(g[n], xv.w, acc[n][3]);
      accb[n] += g[n];
    }
  }
#pragma unroll
  for (int n = 0; n < NL; ++n) {
#pragma unroll
    for (int e = 0; e < 4; ++e) red[slot][n * SK + k4 + e] = acc[n][e];
    if ((tid & 15) == 0) red[slot][NL * SK + n] = accb[n];
  }
  __syncthreads();
  float* p = partial + (long)blockIdx.x * (NL * SK + NL);
  for (int i = tid; i < NL * SK + NL; i += 256) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) s += red[g][i];
    p[i] = s;
  }
}

inline bool supported(int N, int K) { return K == SK && (N == 4 || N == 8); }
inline size_t partial_floats(int N) { return (size_t)WG_CTAS * (N * SK + N); }

inline int fwd(const float* x, const float* W, const float* b, float* out, long M, int N, cudaStream_t st) {
  const unsigned grid = (unsigned)((M + 127) / 128);
  if (N == 8) LAUNCH_PDL((skinny_fwd_kernel<8>), grid, 128, 0, st, x, W, b, out, M);
  else LAUNCH_PDL((skinny_fwd_kernel<4>), grid, 128, 0, st, x, W, b, out, M);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}
inline int dgrad(const float* dy, const float* W, float* dx, long M, int N, cudaStream_t st) {
  const unsigned grid = (unsigned)((M * 16 + 255) / 256);
  if (N == 8) LAUNCH_PDL((skinny_dgrad_kernel<8>), grid, 256, 0, st, dy, W, dx, M);
  else LAUNCH_PDL((skinny_dgrad_kernel<4>), grid, 256, 0, st, dy, W, dx, M);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}
// dW (N,64) and db (N) are OVERWRITTEN; partial: partial_floats(N) floats of scratch
inline int wgrad(const float* dy, const float* x, float* dW, float* db, float* partial, long M, int N, cudaStream_t st) {
  int ctas = (int)(M < WG_CTAS ? M : WG_CTAS);
  if (N == 8) LAUNCH_PDL((skinny_wgrad_kernel<8>), ctas, 256, 0, st, dy, x, partial, M);
  else LAUNCH_PDL((skinny_wgrad_kernel<4>), ctas, 256, 0, st, dy, x, partial, M);
  LAUNCH_CHECK();
  lintc::WgradReduceArgs r;
  r.partial = partial; r.ctas = ctas; r.Nout = N; r.Kin = SK; r.rows_per_dst = N; r.ldw = SK; r.log_scale = nullptr;
  for (int i = 0; i < 4; ++i) { r.dW[i] = nullptr; r.db[i] = nullptr; }
  r.dW[0] = dW; r.db[0] = db;
  const int total = N * SK + N;
  LAUNCH_PDL((lintc::lin_wgrad_reduce_kernel), dim3((total + 31) / 32, 1), 256, 0, st, r);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

}  // namespace skinny
}  // namespace eegclip
