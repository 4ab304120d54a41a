// Short-sequence multi-head attention for the EEG encoder (clip_model.py:30-45):
//   8 heads x head_dim 8, T <= 512 tokens, softmax(Q K^T / sqrt(emb_size)) with dropout on the
//   probabilities, no mask.  The (B,8,T,T) energy / probability tensors of the reference (839 MB
//   each at B=256, T=320) never exist: scores are produced, exponentiated, masked and consumed in
//   registers; the backward recomputes them from q, k and the saved log-sum-exp.
// qkv is (B,T,192) = [q | k | v], head h owning columns h*8..h*8+7 of each third.
#pragma once
#include "common.cuh"

namespace eegclip {

constexpr int AH = 8;    // heads
constexpr int AD = 8;    // head dim
constexpr int AE = 64;   // embedding
constexpr int AQKV = 192;

__device__ __forceinline__ void load8(const float* p, float* r) {
  float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float* r) {
  reinterpret_cast<float4*>(p)[0] = make_float4(r[0], r[1], r[2], r[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(r[4], r[5], r[6], r[7]);
}
__device__ __forceinline__ float dot8(const float* a, const float* b) {
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < 8; ++d) s = fmaf(a[d], b[d], s);
  return s;
}

// grid: (B*H), block: 128. smem: K,V of the head (2*T*8 floats).
__global__ void __launch_bounds__(128) attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                      float* __restrict__ lse, int T, float scale, Drop drop) {
  extern __shared__ float sm[];
  float* Ks = sm;
  float* Vs = sm + (long)T * AD;
  const int bh = blockIdx.x, b = bh / AH, h = bh % AH;
  const float* base = qkv + (long)b * T * AQKV;
  for (int i = threadIdx.x; i < T * 2; i += blockDim.x) {  // 2 float4 per row per matrix
    int t = i >> 1, half = i & 1;
    reinterpret_cast<float4*>(Ks + t * AD)[half] = reinterpret_cast<const float4*>(base + (long)t * AQKV + 64 + h * AD)[half];
    reinterpret_cast<float4*>(Vs + t * AD)[half] = reinterpret_cast<const float4*>(base + (long)t * AQKV + 128 + h * AD)[half];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    float q[8], acc[8];
    load8(base + (long)i * AQKV + h * AD, q);
#pragma unroll
    for (int d = 0; d < 8; ++d) { q[d] *= scale; acc[d] = 0.f; }
    float m = -INFINITY, l = 0.f;
    const uint64_t row_idx = ((uint64_t)bh * T + i) * (uint64_t)T;
    for (int j0 = 0; j0 < T; j0 += 4) {
      float s[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) s[u] = dot8(q, Ks + (j0 + u) * AD);
      float mn = fmaxf(fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])), m);
      float corr = __expf(m - mn);
      l *= corr;
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] *= corr;
      float4 dm = drop_mult4(drop, row_idx + j0);
      float mm[4] = {dm.x, dm.y, dm.z, dm.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float p = __expf(s[u] - mn);
        l += p;
        float pd = p * mm[u];
        const float* v = Vs + (j0 + u) * AD;
#pragma unroll
        for (int d = 0; d < 8; ++d) acc[d] = fmaf(pd, v[d], acc[d]);
      }
      m = mn;
    }
    float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] *= inv;
    store8(out + ((long)b * T + i) * AE + h * AD, acc);
    lse[(long)bh * T + i] = m + __logf(l);
  }
}

// dQ pass: one thread per query row. grid (B*H), block 128, smem K,V.
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const float* __restrict__ qkv, const float* __restrict__ out,
                                                         const float* __restrict__ dout, const float* __restrict__ lse,
                                                         float* __restrict__ dqkv, int T, float scale, Drop drop) {
  extern __shared__ float sm[];
  float* Ks = sm;
  float* Vs = sm + (long)T * AD;
  const int bh = blockIdx.x, b = bh / AH, h = bh % AH;
  const float* base = qkv + (long)b * T * AQKV;
  for (int i = threadIdx.x; i < T * 2; i += blockDim.x) {
    int t = i >> 1, half = i & 1;
    reinterpret_cast<float4*>(Ks + t * AD)[half] = reinterpret_cast<const float4*>(base + (long)t * AQKV + 64 + h * AD)[half];
    reinterpret_cast<float4*>(Vs + t * AD)[half] = reinterpret_cast<const float4*>(base + (long)t * AQKV + 128 + h * AD)[half];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    float q[8], dO[8], o[8], dq[8];
    load8(base + (long)i * AQKV + h * AD, q);
    load8(dout + ((long)b * T + i) * AE + h * AD, dO);
    load8(out + ((long)b * T + i) * AE + h * AD, o);
    const float Di = dot8(dO, o);
    const float L = lse[(long)bh * T + i];
#pragma unroll
    for (int d = 0; d < 8; ++d) { q[d] *= scale; dq[d] = 0.f; }
    const uint64_t row_idx = ((uint64_t)bh * T + i) * (uint64_t)T;
    for (int j0 = 0; j0 < T; j0 += 4) {
      float4 dm = drop_mult4(drop, row_idx + j0);
      float mm[4] = {dm.x, dm.y, dm.z, dm.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* kj = Ks + (j0 + u) * AD;
        float p = __expf(dot8(q, kj) - L);
        float dP = dot8(dO, Vs + (j0 + u) * AD) * mm[u];
        float dS = p * (dP - Di);
#pragma unroll
        for (int d = 0; d < 8; ++d) dq[d] = fmaf(dS, kj[d], dq[d]);
      }
    }
#pragma unroll
    for (int d = 0; d < 8; ++d) dq[d] *= scale;
    store8(dqkv + ((long)b * T + i) * AQKV + h * AD, dq);
  }
}

// dK/dV pass: one thread per 4 consecutive keys. grid (B*H), block = multiple of 32 >= T/4.
// smem per query row: q*scale (8), dO (8), lse, D  -> 18 floats.
__global__ void attn_bwd_dkv_kernel(const float* __restrict__ qkv, const float* __restrict__ out, const float* __restrict__ dout,
                                    const float* __restrict__ lse, float* __restrict__ dqkv, int T, float scale, Drop drop) {
  extern __shared__ float sm[];
  float* Qs = sm;                       // T*8
  float* dOs = Qs + (long)T * AD;       // T*8
  float* Ls = dOs + (long)T * AD;       // T
  float* Ds = Ls + T;                   // T
  const int bh = blockIdx.x, b = bh / AH, h = bh % AH;
  const float* base = qkv + (long)b * T * AQKV;
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    float q[8], dO[8], o[8];
    load8(base + (long)i * AQKV + h * AD, q);
    load8(dout + ((long)b * T + i) * AE + h * AD, dO);
    load8(out + ((long)b * T + i) * AE + h * AD, o);
#pragma unroll
    for (int d = 0; d < 8; ++d) q[d] *= scale;
    store8(Qs + i * AD, q);
    store8(dOs + i * AD, dO);
    Ls[i] = lse[(long)bh * T + i];
    Ds[i] = dot8(dO, o);
  }
  __syncthreads();
  for (int j0 = threadIdx.x * 4; j0 < T; j0 += blockDim.x * 4) {
    float k[4][8], v[4][8], dk[4][8], dv[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      load8(base + (long)(j0 + u) * AQKV + 64 + h * AD, k[u]);
      load8(base + (long)(j0 + u) * AQKV + 128 + h * AD, v[u]);
#pragma unroll
      for (int d = 0; d < 8; ++d) { dk[u][d] = 0.f; dv[u][d] = 0.f; }
    }
    for (int i = 0; i < T; ++i) {
      const float* qi = Qs + i * AD;
      const float* dOi = dOs + i * AD;
      const float L = Ls[i], Di = Ds[i];
      float4 dm = drop_mult4(drop, ((uint64_t)bh * T + i) * (uint64_t)T + j0);
      float mm[4] = {dm.x, dm.y, dm.z, dm.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float p = __expf(dot8(qi, k[u]) - L);
        float pd = p * mm[u];
        float dP = dot8(dOi, v[u]) * mm[u];
        float dS = p * (dP - Di);
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          dv[u][d] = fmaf(pd, dOi[d], dv[u][d]);
          dk[u][d] = fmaf(dS, qi[d], dk[u][d]);   // qi already carries the 1/sqrt(emb) scale
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      store8(dqkv + ((long)b * T + j0 + u) * AQKV + 64 + h * AD, dk[u]);
      store8(dqkv + ((long)b * T + j0 + u) * AQKV + 128 + h * AD, dv[u]);
    }
  }
}

inline int attention_fwd(const float* qkv, float* out, float* lse, int B, int T, const Drop& drop, cudaStream_t st) {
  if (T % 4) return EEGCLIP_ERR_UNSUPPORTED;
  size_t smem = (size_t)T * AD * 2 * sizeof(float);
  ProfScope prof(PROF_ATTN_FWD, st);
  attn_fwd_kernel<<<B * AH, 128, smem, st>>>(qkv, out, lse, T, 0.125f, drop);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

inline int attention_bwd(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv, int B, int T,
                         const Drop& drop, cudaStream_t st) {
  if (T % 4) return EEGCLIP_ERR_UNSUPPORTED;
  size_t smem = (size_t)T * AD * 2 * sizeof(float);
  ProfScope prof(PROF_ATTN_BWD, st);
  attn_bwd_dq_kernel<<<B * AH, 128, smem, st>>>(qkv, out, dout, lse, dqkv, T, 0.125f, drop);
  LAUNCH_CHECK();
  int threads = ((T / 4 + 31) / 32) * 32;
  size_t smem2 = (size_t)T * (2 * AD + 2) * sizeof(float);
  attn_bwd_dkv_kernel<<<B * AH, threads, smem2, st>>>(qkv, out, dout, lse, dqkv, T, 0.125f, drop);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

}  // namespace eegclip
