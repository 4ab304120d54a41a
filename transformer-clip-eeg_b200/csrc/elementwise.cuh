// HBM-bound kernels of the encoder: padding, LayerNorm([C,T])+activation, per-token LayerNorm,
// GELU/dropout, column sums, L2 normalisation.  All activations are time-major (B,T,C) fp32.
// Judged on achieved GB/s: every kernel reads/writes each element once, 128-bit where the
// layout allows, warp-shuffle reductions, no smem staging (no reuse to exploit).
#pragma once
#include "common.cuh"

namespace eegclip {

// ---------------------------------------------------------------------------------------------
// upad[b][PL + t][c] = x1[b][t][c] (+ x2[b][t][c]); rows [0,PL) and [PL+T, T+taps-1) are zero.
// Conv1d 'same' (clip_model.py:237): PL = (k-1)/2 for the forward, k-1-PL for the data gradient.
// ---------------------------------------------------------------------------------------------
__global__ void pad_add_kernel(const float* __restrict__ x1, const float* __restrict__ x2, float* __restrict__ out,
                               int T, int C, int PL, int TP, long total4) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  int c4 = C >> 2;
  long row = i / c4;           // b*TP + tp
  int cc = (int)(i - row * c4);
  long b = row / TP;
  int tp = (int)(row - b * TP);
  int t = tp - PL;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t >= 0 && t < T) {
    long src = ((b * T + t) * (long)C) / 4 + cc;
    v = reinterpret_cast<const float4*>(x1)[src];
    if (x2) {
      float4 w = reinterpret_cast<const float4*>(x2)[src];
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
  }
  reinterpret_cast<float4*>(out)[i] = v;
}

inline int pad_add(const float* x1, const float* x2, float* out, int B, int T, int C, int PL, int taps, cudaStream_t st) {
  int TP = T + taps - 1;
  long total4 = (long)B * TP * (C / 4);
  pad_add_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(x1, x2, out, T, C, PL, TP, total4);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// narrow-output convs on the 64-channel tensor-core kernels: channel padding of the output gradient and compaction (+ dropout,
// indexed in the compact (rows, Cs) layout like every other path) of the padded conv output
__global__ void pad_channels_kernel(const float* __restrict__ src, float* __restrict__ dst, long rows, int Cs, int Cd) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;      // one float4 of dst
  const int d4 = Cd >> 2;
  if (i >= rows * d4) return;
  const long r = i / d4;
  const int c = (int)(i - r * d4) * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < Cs) v = *reinterpret_cast<const float4*>(src + r * Cs + c);
  reinterpret_cast<float4*>(dst)[i] = v;
}
inline int pad_channels(const float* src, float* dst, long rows, int Cs, int Cd, cudaStream_t st) {
  const long n4 = rows * (Cd >> 2);
  pad_channels_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(src, dst, rows, Cs, Cd);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}
__global__ void compact_channels_drop_kernel(const float* __restrict__ src, float* __restrict__ dst, long rows, int Cs, int Cd, Drop d) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;      // one float4 of dst
  const int d4 = Cd >> 2;
  if (i >= rows * d4) return;
  const long r = i / d4;
  const int c = (int)(i - r * d4) * 4;
  float4 v = *reinterpret_cast<const float4*>(src + r * Cs + c);
  const float4 m = drop_mult4(d, (uint64_t)i * 4);
  v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
  reinterpret_cast<float4*>(dst)[i] = v;
}
inline int compact_channels_drop(const float* src, float* dst, long rows, int Cs, int Cd, const Drop& d, cudaStream_t st) {
  const long n4 = rows * (Cd >> 2);
  compact_channels_drop_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(src, dst, rows, Cs, Cd, d);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// out = a + b (b optional), float4
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long n4) {
  pdl_sync();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<const float4*>(a)[i];
  if (b) { float4 w = reinterpret_cast<const float4*>(b)[i]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
  reinterpret_cast<float4*>(out)[i] = v;
}
inline int add_f32(const float* a, const float* b, float* out, long n, cudaStream_t st) {
  long n4 = n / 4;
  LAUNCH_PDL((add_kernel), (unsigned)((n4 + 255) / 256), 256, 0, st, a, b, out, n4);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// out = a + b + c, float4 (the two skip-gradient contributions of an interleaved layer in one pass)
__global__ void add3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c, float* __restrict__ out,
                            long n4) {
  pdl_sync();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<const float4*>(a)[i];
  const float4 w = reinterpret_cast<const float4*>(b)[i], u = reinterpret_cast<const float4*>(c)[i];
  v.x += w.x + u.x; v.y += w.y + u.y; v.z += w.z + u.z; v.w += w.w + u.w;
  reinterpret_cast<float4*>(out)[i] = v;
}
inline int add3_f32(const float* a, const float* b, const float* c, float* out, long n, cudaStream_t st) {
  long n4 = n / 4;
  LAUNCH_PDL((add3_kernel), (unsigned)((n4 + 255) / 256), 256, 0, st, a, b, c, out, n4);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// out[i] = in[i] * dropmult(i)      (backward of a residual-branch dropout)
__global__ void drop_mul_kernel(const float* __restrict__ in, float* __restrict__ out, long n4, Drop d) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<const float4*>(in)[i];
  float4 m = drop_mult4(d, (uint64_t)i * 4);
  v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
  reinterpret_cast<float4*>(out)[i] = v;
}
inline int drop_mul(const float* in, float* out, long n, const Drop& d, cudaStream_t st) {
  long n4 = n / 4;
  drop_mul_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(in, out, n4, d);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// f = dropout(GELU(pre))   (recomputed in backward instead of being stored)
__global__ void gelu_drop_kernel(const float* __restrict__ pre, float* __restrict__ out, long n4, Drop d) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<const float4*>(pre)[i];
  float4 m = drop_mult4(d, (uint64_t)i * 4);
  v.x = gelu_f(v.x) * m.x; v.y = gelu_f(v.y) * m.y; v.z = gelu_f(v.z) * m.z; v.w = gelu_f(v.w) * m.w;
  reinterpret_cast<float4*>(out)[i] = v;
}
inline int gelu_drop(const float* pre, float* out, long n, const Drop& d, cudaStream_t st) {
  long n4 = n / 4;
  gelu_drop_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(pre, out, n4, d);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// dpre = dout * dropmult * GELU'(pre)   (backward of gelu_drop)
__global__ void gelu_drop_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ dout, float* __restrict__ dpre, long n4,
                                     Drop d) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<const float4*>(pre)[i], g = reinterpret_cast<const float4*>(dout)[i];
  float4 m = drop_mult4(d, (uint64_t)i * 4);
  g.x *= gelu_grad_f(v.x) * m.x; g.y *= gelu_grad_f(v.y) * m.y; g.z *= gelu_grad_f(v.z) * m.z; g.w *= gelu_grad_f(v.w) * m.w;
  reinterpret_cast<float4*>(dpre)[i] = g;
}
inline int gelu_drop_bwd(const float* pre, const float* dout, float* dpre, long n, const Drop& d, cudaStream_t st) {
  long n4 = n / 4;
  gelu_drop_bwd_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(pre, dout, dpre, n4, d);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the whole (C,T) sample with (C,T)-shaped affine + activation (+ skip).
// clip_model.py:239,247-248 (BasicBlock), vlaai.py:31,62.  One CTA per sample; the sample (T*C*4 B
// = 80 KB at T=320,C=64) is read twice (second read hits L1/L2).
//   y     : (B,T,C) conv output after bias+dropout
//   gamma : (C,T) reference layout (channel-major)
//   out   : act(gamma*xhat+beta) + skip
// ---------------------------------------------------------------------------------------------
// (C,T) -> (T,C) copies of the affine parameters: the activations are time-major, the checkpoint layout of the LayerNorm([C,T])
// affine is channel-major; reading it in place costs 32 sectors per warp load (measured: the LN kernels ran at 1.2-1.9 TB/s).
// One 164 KB transpose per block and direction makes every parameter access a coalesced float4.
constexpr int LNCT_GROUPS = 32;      // sample groups of the backward apply pass (affine-gradient partials per group)
// (gT, bT, 2 x LNCT_GROUPS affine-gradient partials) + the conv-bias partials of the fused column sum: [groups][CTAs along T*C][C]
inline size_t ln_ct_bias_part_floats(int T, int C) { return (size_t)(((long)T * C / 8 + 255) / 256) * LNCT_GROUPS * C; }
inline size_t ln_ct_scratch_floats(int T, int C) { return (size_t)(2 + 2 * LNCT_GROUPS) * T * C + ln_ct_bias_part_floats(T, C); }

__global__ void __launch_bounds__(256) ct_transpose2_kernel(const float* __restrict__ g, const float* __restrict__ b,
                                                           float* __restrict__ gT, float* __restrict__ bT, int C, int T) {
  pdl_sync();
  __shared__ float tile[2][32][33];
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, t = t0 + tx;
    if (c < C && t < T) { tile[0][r][tx] = g[(long)c * T + t]; tile[1][r][tx] = b[(long)c * T + t]; }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, c = c0 + tx;
    if (c < C && t < T) { gT[(long)t * C + c] = tile[0][tx][r]; bT[(long)t * C + c] = tile[1][tx][r]; }
  }
}

// dgamma[c][t] += sum_g part[g][0][t][c],  dbeta[c][t] += sum_g part[g][1][t][c]   (fixed order: deterministic)
// grid (T/8, C/32), block 256: thread (c = tid & 31, t = tid >> 5) sums its element over the groups with the loads unrolled
// (independent, coalesced), the 8 x 32 tile is transposed through shared memory for the (C,T) write.
__global__ void __launch_bounds__(256) ct_reduce_transpose_kernel(const float* __restrict__ part, float* __restrict__ dgamma,
                                                                 float* __restrict__ dbeta, int groups, int C, int T) {
  pdl_sync();
  __shared__ float tile[2][8][33];
  const int t0 = blockIdx.x * 8, c0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long n = (long)T * C;
  {
    const int t = t0 + ty, c = c0 + tx;
    float sg = 0.f, sb = 0.f;
    if (c < C && t < T) {
      const float* p0 = part + (long)t * C + c;
#pragma unroll 8
      for (int g = 0; g < groups; ++g) {
        sg += __ldg(p0 + (long)(2 * g) * n);
        sb += __ldg(p0 + (long)(2 * g + 1) * n);
      }
    }
    tile[0][ty][tx] = sg; tile[1][ty][tx] = sb;
  }
  __syncthreads();
  {
    const int c = c0 + (threadIdx.x >> 3), t = t0 + (threadIdx.x & 7);
    if (c < C && t < T) {
      dgamma[(long)c * T + t] += tile[0][threadIdx.x & 7][threadIdx.x >> 3];
      dbeta[(long)c * T + t] += tile[1][threadIdx.x & 7][threadIdx.x >> 3];
    }
  }
}

inline int ct_transpose2(const float* g, const float* b, float* gT, float* bT, int C, int T, cudaStream_t st) {
  dim3 grid((T + 31) / 32, (C + 31) / 32);
  LAUNCH_PDL((ct_transpose2_kernel), grid, 256, 0, st, g, b, gT, bT, C, T);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// gT / bT: (T,C) transposed affine (see above)
__global__ void __launch_bounds__(512) ln_ct_act_fwd_kernel(const float* __restrict__ y, const float* __restrict__ gT,
                                                           const float* __restrict__ bT, const float* __restrict__ skip,
                                                           float* __restrict__ out, float* __restrict__ stats, int T, int C,
                                                           int act, float eps) {
  pdl_sync();
  __shared__ float2 sh[33];
  const int b = blockIdx.x;
  const long n = (long)T * C;
  const float* yb = y + b * n;
  float s = 0.f, ss = 0.f;
  for (long i = threadIdx.x * 4L; i < n; i += blockDim.x * 4L) {
    float4 v = *reinterpret_cast<const float4*>(yb + i);
    s += v.x + v.y + v.z + v.w;
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  float2 r = block_sum2(s, ss, sh);
  float mean = r.x / (float)n;
  // second, numerically safer pass for the variance (values are L1/L2 resident)
  float sq = 0.f;
  for (long i = threadIdx.x * 4L; i < n; i += blockDim.x * 4L) {
    float4 v = *reinterpret_cast<const float4*>(yb + i);
    float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
    sq += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
  }
  r = block_sum2(sq, 0.f, sh);
  const float var = r.x / (float)n;
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) { stats[2 * b] = mean; stats[2 * b + 1] = rstd; }
  const float* sb = skip ? skip + b * n : nullptr;
  float* ob = out + b * n;
  for (long i = threadIdx.x * 4L; i < n; i += blockDim.x * 4L) {
    const float4 v = *reinterpret_cast<const float4*>(yb + i);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gT + i)), be = __ldg(reinterpret_cast<const float4*>(bT + i));
    float4 o;
    o.x = act_f((v.x - mean) * rstd * g.x + be.x, act);
    o.y = act_f((v.y - mean) * rstd * g.y + be.y, act);
    o.z = act_f((v.z - mean) * rstd * g.z + be.z, act);
    o.w = act_f((v.w - mean) * rstd * g.w + be.w, act);
    if (sb) { const float4 k = *reinterpret_cast<const float4*>(sb + i); o.x += k.x; o.y += k.y; o.z += k.z; o.w += k.w; }
    *reinterpret_cast<float4*>(ob + i) = o;
  }
}

// Register-resident variant for samples of up to 512 * 4 * NV floats (T = 320, C = 64: NV = 10): every thread issues all of its
// loads up front (NV independent 128-bit loads in flight per thread), keeps its slice in registers for the mean, the centred
// variance and the normalise pass -- one HBM read of y instead of one HBM + two L2 passes with a load-use chain per iteration.
template <int NV>
__global__ void __launch_bounds__(512) ln_ct_act_fwd_reg_kernel(const float* __restrict__ y, const float* __restrict__ gT,
                                                               const float* __restrict__ bT, const float* __restrict__ skip,
                                                               float* __restrict__ out, float* __restrict__ stats, int T, int C,
                                                               int act, float eps) {
  pdl_sync();
  __shared__ float2 sh[33];
  const int b = blockIdx.x;
  const int n = T * C;
  const float* yb = y + (long)b * n;
  float4 v[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = (k * 512 + threadIdx.x) * 4;
    v[k] = i < n ? *reinterpret_cast<const float4*>(yb + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  float2 r = block_sum2(s, 0.f, sh);
  const float mean = r.x / (float)n;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = (k * 512 + threadIdx.x) * 4;
    if (i < n) {
      const float a0 = v[k].x - mean, a1 = v[k].y - mean, a2 = v[k].z - mean, a3 = v[k].w - mean;
      sq += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
    }
  }
  r = block_sum2(sq, 0.f, sh);
  const float rstd = rsqrtf(r.x / (float)n + eps);
  if (threadIdx.x == 0) { stats[2 * b] = mean; stats[2 * b + 1] = rstd; }
  const float* sb = skip ? skip + (long)b * n : nullptr;
  float* ob = out + (long)b * n;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = (k * 512 + threadIdx.x) * 4;
    if (i < n) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gT + i)), be = __ldg(reinterpret_cast<const float4*>(bT + i));
      float4 o;
      o.x = act_f((v[k].x - mean) * rstd * g.x + be.x, act);
      o.y = act_f((v[k].y - mean) * rstd * g.y + be.y, act);
      o.z = act_f((v[k].z - mean) * rstd * g.z + be.z, act);
      o.w = act_f((v[k].w - mean) * rstd * g.w + be.w, act);
      if (sb) { const float4 kk = *reinterpret_cast<const float4*>(sb + i); o.x += kk.x; o.y += kk.y; o.z += kk.z; o.w += kk.w; }
      *reinterpret_cast<float4*>(ob + i) = o;
    }
  }
}

// lnscr: ln_ct_scratch_floats(T, C) floats
inline int ln_ct_act_fwd(const float* y, const float* gamma, const float* beta, const float* skip, float* out, float* stats,
                         float* lnscr, int B, int T, int C, int act, cudaStream_t st) {
  if (C & 3) return EEGCLIP_ERR_UNSUPPORTED;
  ProfScope prof(PROF_LNCT, st);
  float* gT = lnscr;
  float* bT = lnscr + (size_t)T * C;
  { int rc = ct_transpose2(gamma, beta, gT, bT, C, T, st); if (rc != EEGCLIP_OK) return rc; }
  if ((long)T * C <= 512L * 4 * 10) LAUNCH_PDL((ln_ct_act_fwd_reg_kernel<10>), B, 512, 0, st, y, gT, bT, skip, out, stats, T, C, act, 1e-5f);
  else LAUNCH_PDL((ln_ct_act_fwd_kernel), B, 512, 0, st, y, gT, bT, skip, out, stats, T, C, act, 1e-5f);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// Backward of the block above.  dout: gradient w.r.t. out (skip gradient handled by the caller).
// Writes dy (gradient w.r.t. the conv output *before* dropout, i.e. already multiplied by the conv
// dropout mask) into a zero-padded buffer dypad[b][PL + t][c] (rows outside are left untouched: the
// caller zeroes the buffer once), and accumulates dgamma/dbeta (C,T).
//   pass 1 (one CTA per sample): the two per-sample means  m1 = mean(dxh), m2 = mean(dxh * xhat)
//   pass 2 (thread = fixed (t, 8 channels), loop over a group of samples): dy, and dgamma/dbeta summed over the group in
//          registers -> one (T,C) partial per group, summed over the groups in fixed order and transposed by
//          ct_reduce_transpose_kernel (deterministic, no atomics)
__global__ void __launch_bounds__(512) ln_ct_bwd_stats_kernel(const float* __restrict__ dout, const float* __restrict__ y,
                                                             const float* __restrict__ stats, const float* __restrict__ gT,
                                                             const float* __restrict__ bT, float* __restrict__ m12, int T, int C,
                                                             int act) {
  pdl_sync();
  __shared__ float2 sh[33];
  const int b = blockIdx.x;
  const long n = (long)T * C;
  const float* yb = y + b * n;
  const float* db = dout + b * n;
  const float mean = stats[2 * b], rstd = stats[2 * b + 1];
  float s1 = 0.f, s2 = 0.f;
  constexpr int U = 1;                                   // (batching the loads of several iterations measured slower: 23.7 vs 15.4 us)
  for (long base = 0; base < n; base += (long)U * blockDim.x * 4L) {
    float4 v[U], d[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long i = base + ((long)k * blockDim.x + threadIdx.x) * 4L;
      if (i < n) { v[k] = *reinterpret_cast<const float4*>(yb + i); d[k] = *reinterpret_cast<const float4*>(db + i); }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long i = base + ((long)k * blockDim.x + threadIdx.x) * 4L;
      if (i < n) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gT + i)), b4 = __ldg(reinterpret_cast<const float4*>(bT + i));
        const float vv[4] = {v[k].x, v[k].y, v[k].z, v[k].w}, dd[4] = {d[k].x, d[k].y, d[k].z, d[k].w};
        const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xh = (vv[j] - mean) * rstd;
          const float dxh = dd[j] * act_grad_f(xh * gg[j] + bb[j], act) * gg[j];
          s1 += dxh; s2 += dxh * xh;
        }
      }
    }
  }
  const float2 r = block_sum2(s1, s2, sh);
  if (threadIdx.x == 0) { m12[2 * b] = r.x / (float)n; m12[2 * b + 1] = r.y / (float)n; }
}

// grid (T*C/8/256, groups), block 256; thread = (t, c8) position, samples b = group, group + groups, ...
// BIAS: the conv bias gradient (column sums of dy over samples and time) rides along: every thread sums its dy over its samples, the
// CTA folds its 2048 / C time rows through shared memory in row order and writes one C-vector per CTA to bpart[group][cta][C]
// (folded in fixed order by colsum_fold_kernel); needs 2048 % C == 0.
template <bool BIAS>
__global__ void __launch_bounds__(256) ln_ct_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ y,
                                                             const float* __restrict__ stats, const float* __restrict__ m12,
                                                             const float* __restrict__ gT, const float* __restrict__ bT,
                                                             float* __restrict__ dypad, float* __restrict__ part, int B, int T, int C,
                                                             int PL, int TP, int act, Drop drop, float* __restrict__ bpart) {
  pdl_sync();
  __shared__ __align__(16) float sbias[BIAS ? 256 * 8 : 8];
  const long n = (long)T * C;
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) {
    if (BIAS) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sbias[threadIdx.x * 8 + j] = 0.f;
      __syncthreads();
      if (threadIdx.x < C) {
        const int rows = 2048 / C, c8n = C >> 3;
        float acc = 0.f;
        for (int r = 0; r < rows; ++r) acc += sbias[(r * c8n + (threadIdx.x >> 3)) * 8 + (threadIdx.x & 7)];
        bpart[((long)blockIdx.y * gridDim.x + blockIdx.x) * C + threadIdx.x] = acc;
      }
    }
    return;
  }
  float g[8], be[8], ag[8], ab[8], ob[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gT + i)), g1 = __ldg(reinterpret_cast<const float4*>(gT + i) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bT + i)), b1 = __ldg(reinterpret_cast<const float4*>(bT + i) + 1);
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
    be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { ag[j] = 0.f; ab[j] = 0.f; ob[j] = 0.f; }
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const float mean = stats[2 * b], rstd = stats[2 * b + 1], m1 = m12[2 * b], m2 = m12[2 * b + 1];
    float vv[8], dd[8];
    ld256(y + b * n + i, vv);            // (i is a multiple of 8 floats and the tensors are 32-byte aligned: checked by the launcher)
    ld256(dout + b * n + i, dd);
    float mm[8], o[8];
    drop_mult8(drop, (uint64_t)b * n + i, mm);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (vv[j] - mean) * rstd;
      const float dl = dd[j] * act_grad_f(xh * g[j] + be[j], act);
      ag[j] += dl * xh; ab[j] += dl;
      o[j] = rstd * (dl * g[j] - m1 - xh * m2) * mm[j];
      if (BIAS) ob[j] += o[j];
    }
    st256(dypad + ((long)b * TP + PL) * C + i, o);
  }
  st256(part + (long)(2 * blockIdx.y) * n + i, ag);
  st256(part + (long)(2 * blockIdx.y + 1) * n + i, ab);
  if (BIAS) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sbias[threadIdx.x * 8 + j] = ob[j];
    __syncthreads();
    if (threadIdx.x < C) {       // thread tid of the CTA sits at time row tid / (C/8), channel chunk tid % (C/8)
      const int rows = 2048 / C, c8n = C >> 3;
      float acc = 0.f;
      for (int r = 0; r < rows; ++r) acc += sbias[(r * c8n + (threadIdx.x >> 3)) * 8 + (threadIdx.x & 7)];
      bpart[((long)blockIdx.y * gridDim.x + blockIdx.x) * C + threadIdx.x] = acc;
    }
  }
}
__global__ void __launch_bounds__(1024) colsum_fold_kernel(const float* __restrict__ part, float* __restrict__ out, int ctas, int N);

// m12: 2*B floats of scratch; lnscr: ln_ct_scratch_floats(T, C) floats
inline int ln_ct_act_bwd(const float* dout, const float* y, const float* stats, const float* gamma, const float* beta,
                         float* dypad, float* dgamma, float* dbeta, float* m12, float* lnscr, int B, int T, int C, int PL, int taps,
                         int act, const Drop& drop, cudaStream_t st, float* dbias = nullptr, bool* dbias_done = nullptr) {
  if ((C & 7) || ((((uintptr_t)dout | (uintptr_t)y | (uintptr_t)dypad | (uintptr_t)lnscr) & 31) != 0)) return EEGCLIP_ERR_UNSUPPORTED;   // 256-bit accesses
  ProfScope prof(PROF_LNCT, st);
  float* gT = lnscr;
  float* bT = lnscr + (size_t)T * C;
  float* part = lnscr + (size_t)2 * T * C;
  { int rc = ct_transpose2(gamma, beta, gT, bT, C, T, st); if (rc != EEGCLIP_OK) return rc; }
  LAUNCH_PDL((ln_ct_bwd_stats_kernel), B, 512, 0, st, dout, y, stats, gT, bT, m12, T, C, act);
  LAUNCH_CHECK();
  const long n8 = (long)T * C / 8;
  const int groups = B < LNCT_GROUPS ? B : LNCT_GROUPS;
  dim3 grid((unsigned)((n8 + 255) / 256), groups);
  // dbias (+= column sums of dy over samples and time) from the same pass when a CTA's 2048 floats are whole time rows
  const bool fuse_bias = dbias != nullptr && C <= 256 && (2048 % C) == 0;
  float* bpart = lnscr + (size_t)(2 + 2 * LNCT_GROUPS) * T * C;
  if (fuse_bias)
    LAUNCH_PDL((ln_ct_bwd_apply_kernel<true>), grid, 256, 0, st, dout, y, stats, m12, gT, bT, dypad, part, B, T, C, PL, T + taps - 1, act, drop, bpart);
  else
    LAUNCH_PDL((ln_ct_bwd_apply_kernel<false>), grid, 256, 0, st, dout, y, stats, m12, gT, bT, dypad, part, B, T, C, PL, T + taps - 1, act, drop, bpart);
  LAUNCH_CHECK();
  dim3 g2((T + 7) / 8, (C + 31) / 32);
  LAUNCH_PDL((ct_reduce_transpose_kernel), g2, 256, 0, st, part, dgamma, dbeta, groups, C, T);
  LAUNCH_CHECK();
  if (fuse_bias) {
    LAUNCH_PDL((colsum_fold_kernel), 1, 1024, 0, st, (const float*)bpart, dbias, (int)grid.x * groups, C);
    LAUNCH_CHECK();
  }
  if (dbias_done) *dbias_done = fuse_bias;
  return EEGCLIP_OK;
}

// ---------------------------------------------------------------------------------------------
// Per-token LayerNorm over 64 features (clip_model.py:84,89).  Half-warp per token (float4/lane).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln64_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                      const float* __restrict__ be, float* __restrict__ out, long rows) {
  pdl_sync();
  long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  int l = threadIdx.x & 15;
  if (row >= rows) return;
  float4 v = reinterpret_cast<const float4*>(x + row * 64)[l];
  float s = v.x + v.y + v.z + v.w;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
  float mean = s * (1.f / 64.f);
  float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
  float q = a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o, 16);
  float rstd = rsqrtf(q * (1.f / 64.f) + 1e-5f);
  float4 gg = reinterpret_cast<const float4*>(g)[l], bb = reinterpret_cast<const float4*>(be)[l];
  reinterpret_cast<float4*>(out + row * 64)[l] =
      make_float4(a0 * rstd * gg.x + bb.x, a1 * rstd * gg.y + bb.y, a2 * rstd * gg.z + bb.z, a3 * rstd * gg.w + bb.w);
}
inline int ln64_fwd(const float* x, const float* g, const float* b, float* out, long rows, cudaStream_t st) {
  long threads = rows * 16;
  LAUNCH_PDL((ln64_fwd_kernel), (unsigned)((threads + 255) / 256), 256, 0, st, x, g, b, out, rows);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// dx_out = resid + LNbwd(dh ; x, gamma) ; dgamma += sum dh*xhat ; dbeta += sum dh.
// Each CTA handles a contiguous chunk of rows, keeps its dgamma/dbeta partials in registers/smem
// and issues 128 atomics at the end.
__global__ void __launch_bounds__(256) ln64_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ x,
                                                      const float* __restrict__ g, const float* __restrict__ resid,
                                                      float* __restrict__ dx, float* __restrict__ dgamma,
                                                      float* __restrict__ dbeta, long rows, int rows_per_cta) {
  pdl_sync();
  __shared__ float sg[16][64], sb[16][64];
  const int l = threadIdx.x & 15, grp = threadIdx.x >> 4;  // 16 groups of 16 lanes
  float4 gg = reinterpret_cast<const float4*>(g)[l];
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = make_float4(0.f, 0.f, 0.f, 0.f);
  long r0 = (long)blockIdx.x * rows_per_cta;
  long r1 = min(rows, r0 + rows_per_cta);
  // R rows per iteration: their loads are issued together and the four 16-lane shuffle reductions of each row interleave
  // (one row at a time left the loop latency-bound on its 16 dependent shuffles)
  constexpr int R = 4;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long it0 = r0; it0 < r1; it0 += 16 * R) {       // CTA-uniform trip count: the shuffles below use the full mask
    const long rowb = it0 + grp;
    float4 v[R], d[R], rr[R];
    bool ok[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const long row = rowb + 16 * u;
      ok[u] = row < r1;
      v[u] = ok[u] ? reinterpret_cast<const float4*>(x + row * 64)[l] : z4;
      d[u] = ok[u] ? reinterpret_cast<const float4*>(dh + row * 64)[l] : z4;
      rr[u] = (ok[u] && resid) ? reinterpret_cast<const float4*>(resid + row * 64)[l] : z4;
    }
    float s[R], q[R], s1[R], s2[R];
#pragma unroll
    for (int u = 0; u < R; ++u) s[u] = v[u].x + v[u].y + v[u].z + v[u].w;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < R; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o, 16);
    float a[R][4];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const float mean = s[u] * (1.f / 64.f);
      a[u][0] = v[u].x - mean; a[u][1] = v[u].y - mean; a[u][2] = v[u].z - mean; a[u][3] = v[u].w - mean;
      q[u] = a[u][0] * a[u][0] + a[u][1] * a[u][1] + a[u][2] * a[u][2] + a[u][3] * a[u][3];
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < R; ++u) q[u] += __shfl_xor_sync(0xffffffffu, q[u], o, 16);
    float rstd[R], e[R][4];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      rstd[u] = rsqrtf(q[u] * (1.f / 64.f) + 1e-5f);
#pragma unroll
      for (int j = 0; j < 4; ++j) a[u][j] *= rstd[u];                   // xhat
      ag.x += d[u].x * a[u][0]; ag.y += d[u].y * a[u][1]; ag.z += d[u].z * a[u][2]; ag.w += d[u].w * a[u][3];
      ab.x += d[u].x; ab.y += d[u].y; ab.z += d[u].z; ab.w += d[u].w;
      e[u][0] = d[u].x * gg.x; e[u][1] = d[u].y * gg.y; e[u][2] = d[u].z * gg.z; e[u][3] = d[u].w * gg.w;
      s1[u] = e[u][0] + e[u][1] + e[u][2] + e[u][3];
      s2[u] = e[u][0] * a[u][0] + e[u][1] * a[u][1] + e[u][2] * a[u][2] + e[u][3] * a[u][3];
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < R; ++u) {
        s1[u] += __shfl_xor_sync(0xffffffffu, s1[u], o, 16);
        s2[u] += __shfl_xor_sync(0xffffffffu, s2[u], o, 16);
      }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const float m1 = s1[u] * (1.f / 64.f), m2 = s2[u] * (1.f / 64.f);
      float4 o4 = make_float4(rstd[u] * (e[u][0] - m1 - a[u][0] * m2), rstd[u] * (e[u][1] - m1 - a[u][1] * m2),
                              rstd[u] * (e[u][2] - m1 - a[u][2] * m2), rstd[u] * (e[u][3] - m1 - a[u][3] * m2));
      o4.x += rr[u].x; o4.y += rr[u].y; o4.z += rr[u].z; o4.w += rr[u].w;
      if (ok[u]) reinterpret_cast<float4*>(dx + (rowb + 16 * u) * 64)[l] = o4;
    }
  }
  reinterpret_cast<float4*>(&sg[grp][0])[l] = ag;
  reinterpret_cast<float4*>(&sb[grp][0])[l] = ab;
  __syncthreads();
  if (threadIdx.x < 128) {
    int c = threadIdx.x & 63;
    float acc = 0.f;
    if (threadIdx.x < 64) { for (int k = 0; k < 16; ++k) acc += sg[k][c]; atomicAdd(dgamma + c, acc); }
    else                  { for (int k = 0; k < 16; ++k) acc += sb[k][c]; atomicAdd(dbeta + c, acc); }
  }
}
inline int ln64_bwd(const float* dh, const float* x, const float* g, const float* resid, float* dx, float* dgamma, float* dbeta,
                    long rows, cudaStream_t st) {
  int rows_per_cta = g_tune[10] > 0 ? g_tune[10] : 128;   // 640 CTAs at 81 920 rows (256: 19.26 ms per step, 128: 19.04, 64: 19.10)
  LAUNCH_PDL((ln64_bwd_kernel), (unsigned)((rows + rows_per_cta - 1) / rows_per_cta), 256, 0, st, dh, x, g, resid, dx, dgamma, dbeta, rows, rows_per_cta);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// out[n] += sum_m X[m][n]   (bias gradients). N <= 256, N % 4 == 0.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, float* __restrict__ out, long M, int N,
                                                    int ld, int rows_per_cta) {
  pdl_sync();
  __shared__ float sh[256 * 4];
  const int n4 = N >> 2;                  // float4 columns
  const int lanes = n4;                   // threads across columns
  const int rgroups = 256 / lanes;        // row groups per CTA
  const int cl = threadIdx.x % lanes, rg = threadIdx.x / lanes;
  long r0 = (long)blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rg < rgroups)
    for (long r = r0 + rg; r < r1; r += rgroups) {
      float4 v = reinterpret_cast<const float4*>(X + r * ld)[cl];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  reinterpret_cast<float4*>(sh)[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x < N) {
    int c4 = threadIdx.x >> 2, j = threadIdx.x & 3;
    float acc = 0.f;
    for (int g = 0; g < rgroups; ++g) acc += sh[(g * lanes + c4) * 4 + j];
    atomicAdd(out + threadIdx.x, acc);
  }
}
// Deterministic form: per-CTA partials into `scratch` (colsum_det_ctas(M) x N floats), folded in CTA order by one CTA.
__global__ void __launch_bounds__(256) colsum_part_kernel(const float* __restrict__ X, float* __restrict__ part, long M, int N,
                                                         int ld, int rows_per_cta) {
  pdl_sync();
  __shared__ float sh[256 * 4];
  const int lanes = N >> 2;               // threads across float4 columns
  const int rgroups = 256 / lanes;        // row groups per CTA
  const int cl = threadIdx.x % lanes, rg = threadIdx.x / lanes;
  long r0 = (long)blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rg < rgroups)
    for (long r = r0 + rg; r < r1; r += rgroups) {
      float4 v = reinterpret_cast<const float4*>(X + r * ld)[cl];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  reinterpret_cast<float4*>(sh)[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x < N) {
    int c4 = threadIdx.x >> 2, j = threadIdx.x & 3;
    float acc = 0.f;
    for (int g = 0; g < rgroups; ++g) acc += sh[(g * lanes + c4) * 4 + j];
    part[(long)blockIdx.x * N + threadIdx.x] = acc;
  }
}
// out[n] += sum over c of part[c][n], in a fixed order: 1024 / N thread groups take the partials c = g, g + G, ... (four independent
// chains each, so ~16 loads per column are in flight: the first version walked them with one 64-thread CTA and four chains and took
// 23 us per call), the groups' sums are combined in group order through shared memory.
__global__ void __launch_bounds__(1024) colsum_fold_kernel(const float* __restrict__ part, float* __restrict__ out, int ctas, int N) {
  pdl_sync();
  __shared__ float sh[1024];
  const int G = 1024 / N;                          // N <= 256 -> G >= 4
  const int n = threadIdx.x % N, g = threadIdx.x / N;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (g < G) {
    int c = g;
    for (; c + 3 * G < ctas; c += 4 * G) {
      a0 += part[(long)c * N + n]; a1 += part[(long)(c + G) * N + n];
      a2 += part[(long)(c + 2 * G) * N + n]; a3 += part[(long)(c + 3 * G) * N + n];
    }
    for (; c < ctas; c += G) a0 += part[(long)c * N + n];
  }
  sh[threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (threadIdx.x < N) {
    float acc = 0.f;
    for (int k = 0; k < G; ++k) acc += sh[k * N + threadIdx.x];
    out[threadIdx.x] += acc;
  }
}
constexpr int COLSUM_DET_ROWS = 512;
inline int colsum_det_ctas(long M) { return (int)((M + COLSUM_DET_ROWS - 1) / COLSUM_DET_ROWS); }
inline int colsum_det(const float* X, float* out, long M, int N, int ld, float* scratch, cudaStream_t st) {
  if (N > 256 || (N & 3) || (256 % (N >> 2)) || (ld & 3)) return EEGCLIP_ERR_UNSUPPORTED;
  const int ctas = colsum_det_ctas(M);
  LAUNCH_PDL((colsum_part_kernel), (unsigned)ctas, 256, 0, st, X, scratch, M, N, ld, COLSUM_DET_ROWS);
  LAUNCH_CHECK();
  LAUNCH_PDL((colsum_fold_kernel), 1, 1024, 0, st, (const float*)scratch, out, ctas, N);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

inline int colsum(const float* X, float* out, long M, int N, int ld, cudaStream_t st) {
  if (N > 256 || (N & 3) || (256 % (N >> 2)) || (ld & 3)) return EEGCLIP_ERR_UNSUPPORTED;
  int rows_per_cta = g_tune[11] > 0 ? g_tune[11] : 512;
  LAUNCH_PDL((colsum_kernel), (unsigned)((M + rows_per_cta - 1) / rows_per_cta), 256, 0, st, X, out, M, N, ld, rows_per_cta);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// Column sum over a zero-padded (B,TP,C) buffer (conv bias gradient): pads are zero so they can be summed too.

// dW[co][ci][k] = tmp[co][k][ci]   (weight-gradient GEMM output -> reference Conv1d layout)
__global__ void wgrad_unpack_kernel(const float* __restrict__ tmp, float* __restrict__ dW, int Cout, int Cin, int K, long total) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int k = (int)(i % K);
  long r = i / K;
  int ci = (int)(r % Cin);
  int co = (int)(r / Cin);
  dW[i] = tmp[((long)co * K + k) * Cin + ci];
}
inline int wgrad_unpack(const float* tmp, float* dW, int Cout, int Cin, int K, cudaStream_t st) {
  long total = (long)Cout * Cin * K;
  wgrad_unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(tmp, dW, Cout, Cin, K, total);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

}  // namespace eegclip
