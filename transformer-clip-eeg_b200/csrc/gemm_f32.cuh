// Generic strided fp32 GEMM on CUDA cores: the exact-fp32 companion of the tcgen05 kernels.
// It serves the small / irregular contractions of the path (latent projection N=8, bias-free
// views, weight-gradient reductions with a folded (batch,time) K index) and is the
// bit-conservative (fp32 FMA) implementation every tensor-core kernel is validated against on the GPU.
//
//   C[b][m][n] (+)= alpha * sum_k A[b][m][k] * B[b][k][n]      k = kb*KT + kt
//   addr(A) = A + b*a_bs + m*a_ms + kb*a_kbs + kt*a_ks   (likewise B with n), C + b*c_bs + m*c_ms + n*c_ns
//
// Overlapping rows (a_ms < KT*a_ks) are allowed: this is how Conv1d('same') over a zero-padded
// time-major buffer becomes a GEMM with K = taps*Cin without materialising im2col.
#pragma once
#include "common.cuh"

namespace eegclip {

struct GemmEpi {
  const float* bias_n = nullptr;      // + bias[n]
  int act = 0;                        // 1: GELU (pre-activation optionally stored to aux)
  float* aux = nullptr;               // same indexing as C
  const float* act_grad_src = nullptr;  // multiply by gelu'(src[m,n]) (same indexing as C)
  int drop_on = 0;                    // multiply by dropout keep/scale; index = (b*M + m)*N + n
  Drop drop;
  int drop_before_actgrad = 0;
  const float* residual = nullptr;    // + R[m,n] (same indexing as C)
  const float* residual2 = nullptr;   // + R2[m,n]
  float* copy_out = nullptr;          // second store of the final value (same indexing as C)
  int atomic = 0;                     // atomicAdd into C (split-K / batch-reduce)
  float alpha = 1.0f;
  // ---- InfoNCE modes (MODE 1: row log-sum-exp partials, MODE 2: gradient-weight tile G) ----
  const float* tau = nullptr;         // device scalar, logits = acc * exp(tau)
  float2* lse_part = nullptr;         // MODE 1: [M][gridDim.x] (max, sumexp) per row and column tile
  const float* lse_m = nullptr;       // MODE 2: log-sum-exp indexed by global row (m + m_off)
  const float* lse_n = nullptr;       // MODE 2: log-sum-exp indexed by global column (n + n_off)
  int m_off = 0, n_off = 0;           // global offsets of the tile rows / columns (for the diagonal)
  float inv_2b = 0.f;                 // 1 / (2 * global batch)
  float* dtau = nullptr;              // MODE 2: += sum G * L   (only when non-null)
  const float* alpha_dev = nullptr;   // MODE 2: device scalar multiplier (upstream gradient of the loss)
  int one_sided = 0;                  // MODE 2: G = (exp(L - lse_row) - delta) / B  (rows-only cross-entropy)
};

struct GemmArgs {
  const float* A = nullptr;
  const float* B = nullptr;
  float* C = nullptr;
  int M = 0, N = 0, K = 0, KT = 0;
  long a_ms = 0, a_ks = 0, a_kbs = 0, a_bs = 0;
  long b_ks = 0, b_ns = 0, b_kbs = 0, b_bs = 0;
  long c_ms = 0, c_ns = 0, c_bs = 0;
  int batch = 1;
  int splitk = 1;
  GemmEpi epi;
};

constexpr int GBM = 64, GBN = 64, GBK = 16;

template <int MODE>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmArgs g) {
  pdl_sync();
  __shared__ float As[GBK][GBM + 4];
  __shared__ float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int bz = blockIdx.z;
  const int b = bz / g.splitk;
  const int sk = bz % g.splitk;
  const int m0 = blockIdx.y * GBM;
  const int n0 = blockIdx.x * GBN;
  // K range of this split (in units of GBK tiles)
  const int ktiles = (g.K + GBK - 1) / GBK;
  const int per = (ktiles + g.splitk - 1) / g.splitk;
  const int kt_beg = sk * per;
  const int kt_end = min(ktiles, kt_beg + per);

  const float* Ab = g.A + (long)b * g.a_bs;
  const float* Bb = g.B + (long)b * g.b_bs;

  const bool a_kc = (g.a_ks == 1);  // k contiguous in A
  const bool b_kc = (g.b_ks == 1);  // k contiguous in B

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int tx = tid & 15;  // n direction
  const int ty = tid >> 4;  // m direction

  for (int kt = kt_beg; kt < kt_end; ++kt) {
    const int k0 = kt * GBK;
    // ---- load A tile (64 x 16) ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m, k;
      if (a_kc) { k = tid & 15; m = (tid >> 4) + 16 * i; }
      else      { m = tid & 63; k = (tid >> 6) + 4 * i; }
      int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < g.K) {
        int kb = gk / g.KT, kk = gk - kb * g.KT;
        v = __ldg(Ab + (long)gm * g.a_ms + (long)kb * g.a_kbs + (long)kk * g.a_ks);
      }
      As[k][m] = v;
    }
    // ---- load B tile (16 x 64) ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n, k;
      if (b_kc) { k = tid & 15; n = (tid >> 4) + 16 * i; }
      else      { n = tid & 63; k = (tid >> 6) + 4 * i; }
      int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < g.K) {
        int kb = gk / g.KT, kk = gk - kb * g.KT;
        v = __ldg(Bb + (long)gn * g.b_ns + (long)kb * g.b_kbs + (long)kk * g.b_ks);
      }
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  const GemmEpi& e = g.epi;
  if (MODE == 1) {
    // Row-wise (max, sum exp) over this tile's 64 columns; 16 consecutive lanes share a row group.
    const float scale = __expf(*e.tau);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int gm = m0 + ty * 4 + i;
      float v[4], mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int gn = n0 + tx * 4 + j;
        v[j] = gn < g.N ? acc[i][j] * scale : -INFINITY;
        mx = fmaxf(mx, v[j]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o, 16));
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) sum += (v[j] == -INFINITY) ? 0.f : __expf(v[j] - mx);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o, 16);
      if (tx == 0 && gm < g.M) e.lse_part[(long)gm * gridDim.x + blockIdx.x] = make_float2(mx, sum);
    }
    return;
  }
  if (MODE == 2) {
    // G[m][n] = (exp(L - lse_row) + exp(L - lse_col) - 2*delta) / (2B),  L = acc * exp(tau)
    __shared__ float red[8];
    const float scale = __expf(*e.tau);
    const float up = e.alpha_dev ? *e.alpha_dev : 1.f;
    float tsum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int gm = m0 + ty * 4 + i;
      if (gm >= g.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int gn = n0 + tx * 4 + j;
        if (gn >= g.N) continue;
        float L = acc[i][j] * scale;
        const bool dg = (gm + e.m_off) == (gn + e.n_off);
        float gval;
        if (e.one_sided) gval = (__expf(L - e.lse_m[gm + e.m_off]) - (dg ? 1.f : 0.f)) * (2.f * e.inv_2b) * up;
        else gval = (__expf(L - e.lse_m[gm + e.m_off]) + __expf(L - e.lse_n[gn + e.n_off]) - (dg ? 2.f : 0.f)) * e.inv_2b * up;
        g.C[(long)gm * g.c_ms + (long)gn * g.c_ns] = gval;
        tsum += gval * L;
      }
    }
    if (e.dtau) {
      tsum = warp_sum(tsum);
      if ((tid & 31) == 0) red[tid >> 5] = tsum;
      __syncthreads();
      if (tid == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += red[k];
        atomicAdd(e.dtau, t);
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      long ci = (long)b * g.c_bs + (long)gm * g.c_ms + (long)gn * g.c_ns;
      float v = acc[i][j] * e.alpha;
      if (e.tau) v *= __expf(*e.tau);
      if (e.atomic) {  // partial sums: epilogue terms are applied by split 0 only
        if (sk == 0 && e.bias_n) v += e.bias_n[gn];
        atomicAdd(g.C + ci, v);
        continue;
      }
      if (e.bias_n) v += e.bias_n[gn];
      if (e.act == 1) {
        if (e.aux) e.aux[ci] = v;
        v = gelu_f(v);
      }
      if (e.drop_on) v *= drop_mult(e.drop, ((uint64_t)b * g.M + gm) * (uint64_t)g.N + gn);
      if (e.act_grad_src) v *= gelu_grad_f(e.act_grad_src[ci]);
      if (e.residual) v += e.residual[ci];
      if (e.residual2) v += e.residual2[ci];
      g.C[ci] = v;
      if (e.copy_out) e.copy_out[ci] = v;
    }
  }
}

template <int MODE = 0>
inline int gemm_f32(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0 || g.batch <= 0) return EEGCLIP_OK;
  GemmArgs a = g;
  if (a.KT <= 0) a.KT = a.K;
  if (a.splitk < 1) a.splitk = 1;
  dim3 grid(ceil_div(a.N, GBN), ceil_div(a.M, GBM), a.batch * a.splitk);
  ProfScope prof(PROF_GEMM_F32, st);
  LAUNCH_PDL((gemm_f32_kernel<MODE>), grid, 256, 0, st, a);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

}  // namespace eegclip
