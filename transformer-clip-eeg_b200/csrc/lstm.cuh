// Bidirectional single-layer LSTM of the default speech tower (clip_model.py:267-268, 322-323; nn.LSTM semantics: gate order
// i,f,g,o, zero initial state, batch_first, outputs of the two directions concatenated).
//
// The input projections x.W_ih^T (+ both biases) of ALL time steps are one token GEMM (lin_tc.cuh), and so are the weight
// gradients and the input gradient; only the recurrence itself is sequential.  The recurrence keeps W_hh on chip:
//   H = 128 (speech_lstm1): one CTA = 4 sequences of one direction, 512 threads = the 512 gate rows; each thread holds its
//       row of W_hh (64 weights in registers, 64 in shared memory), h_{t-1} of the 4 sequences is broadcast from shared
//       memory; exact fp32 FMA.  The backward runs the transposed product the same way (thread = (hidden unit, row quarter)).
//   H = 4 (speech_lstm2): one half-warp per (sequence, direction), lane = gate row, coalesced 64-byte gate rows, shuffles.
// Saved for the backward: post-activation gates (in place over the input projections), cell states, h_{t-1}.
#pragma once
#include "common.cuh"

namespace eegclip {
namespace lstm {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// packed fp32 FMA (sm_100 FFMA2): two independent fp32 FMAs per issue slot -- the recurrences are issue-bound on FFMA
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

// ------------------------------------------------------------------------------------------------
// H = 128 forward.  grid (ceil(B/4), 2 directions), block 512.
//   G   : (B*T, GS) rows; direction d owns columns [d*512, d*512+512) = [gate][unit]; in: x-projection + biases, out: gates
//   out : (B*T, 256)  h_t          Cs : (B*T, 256) c_t          Hp : (B*T, 256) h_{t-1} (the state the step started from)
// Recurrent product pre[q][r] = sum_k W[r][k] h[q][k] (512 rows x 4 sequences x K = 128), register-tiled: thread
// (rg = tid >> 2, kq = tid & 3) owns the 4 rows 4rg..4rg+3 over the k-quarter [32kq, 32kq+32): every h value it loads
// feeds 4 rows and every weight 4 sequences (48 shared-memory loads per 256 packed FMAs instead of 144), the four
// k-quarters are summed with two shuffles.  Weights: rows 0,1 of the tile in registers, rows 2,3 in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int LH = 128, LG = 512, LNB = 4;
constexpr int LHS = LNB * 4 * 9 * 4;                                  // h / da staging: [seq][k-quarter][8 float4 + 1 pad]
constexpr int L128_SMEM = (16 * LG * 4 + LHS + LNB * LG) * 4;         // W half + h + pre-activations

__global__ void __launch_bounds__(512, 1) lstm128_fwd_kernel(const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_r,
                                                            float* __restrict__ G, int GS, float* __restrict__ out,
                                                            float* __restrict__ Cs, float* __restrict__ Hp, int B, int T, unsigned long long* dbg) {
  pdl_sync();
  extern __shared__ __align__(16) float sml[];
  float4* Wsm = reinterpret_cast<float4*>(sml);          // [j = 16][tid 512] : rows 2,3 of the thread's tile
  // development timeline: thread 0 of CTA (0,0) stamps the phases of the first steps (tools/lstm_timeline.py)
  unsigned long long* dbgp = (dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) ? dbg : nullptr;
  int dbn = 0;
  auto stamp = [&](int ev) {
    if (dbgp && dbn < 120) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); dbgp[2 * dbn] = (unsigned long long)ev; dbgp[2 * dbn + 1] = t_; ++dbn; dbgp[255] = (unsigned long long)dbn; }
  };
  float* hs = sml + 16 * LG * 4;                         // h of the 4 sequences, padded (see hidx)
  float* pre = hs + LHS;                                 // [seq 4][512]
  const int dir = blockIdx.y, b0 = blockIdx.x * LNB;
  const int tid = threadIdx.x;
  const int rg = tid >> 2, kq = tid & 3;
  const float* W = (dir ? w_hh_r : w_hh_f) + (long)(rg * 4) * LH + kq * 32;   // tile origin: row 4rg, column 32kq
  float w[64];                                           // rows 0,1: w[row*32 + k]
#pragma unroll
  for (int rr = 0; rr < 2; ++rr)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(W + (long)rr * LH) + i);
      w[rr * 32 + 4 * i] = v.x; w[rr * 32 + 4 * i + 1] = v.y; w[rr * 32 + 4 * i + 2] = v.z; w[rr * 32 + 4 * i + 3] = v.w;
    }
#pragma unroll
  for (int j = 0; j < 16; ++j) Wsm[j * LG + tid] = __ldg(reinterpret_cast<const float4*>(W + (long)(2 + (j >> 3)) * LH) + (j & 7));
  // h[q][k] lives at hs[((q*4 + k/32)*9 + (k%32)/4)*4 + k%4]: the four k-quarters start 16 B apart modulo 128 B
  auto hidx = [](int q, int k) { return ((q * 4 + (k >> 5)) * 9 + ((k & 31) >> 2)) * 4 + (k & 3); };
  // gate-combine role: thread = (unit u, sequence s)
  const int u = tid & 127, s = tid >> 7;
  hs[hidx(s, u)] = 0.f;
  const int b = b0 + s;
  const bool live = b < B;
  float c = 0.f;
  __syncthreads();
  const float4* h4 = reinterpret_cast<const float4*>(hs);
  for (int step = 0; step < T; ++step) {
    const int t = dir ? T - 1 - step : step;
    const long row = (long)b * T + t;
    stamp(0);
    float gx[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
#pragma unroll
      for (int g = 0; g < 4; ++g) gx[g] = G[row * GS + dir * LG + g * LH + u];
    }
    // ---- recurrent product over this thread's k-quarter ----
    float2 acc[4][LNB];                                  // [row][seq], (even-k, odd-k) partial sums
#pragma unroll
    for (int rr = 0; rr < 4; ++rr)
#pragma unroll
      for (int q = 0; q < LNB; ++q) acc[rr][q] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 h[LNB];
#pragma unroll
      for (int q = 0; q < LNB; ++q) h[q] = h4[(q * 4 + kq) * 9 + i];
      const float4 w2 = Wsm[i * LG + tid], w3 = Wsm[(8 + i) * LG + tid];
#pragma unroll
      for (int q = 0; q < LNB; ++q) {
        const float2 hlo = make_float2(h[q].x, h[q].y), hhi = make_float2(h[q].z, h[q].w);
        acc[0][q] = ffma2(make_float2(w[4 * i], w[4 * i + 1]), hlo, acc[0][q]);
        acc[0][q] = ffma2(make_float2(w[4 * i + 2], w[4 * i + 3]), hhi, acc[0][q]);
        acc[1][q] = ffma2(make_float2(w[32 + 4 * i], w[32 + 4 * i + 1]), hlo, acc[1][q]);
        acc[1][q] = ffma2(make_float2(w[32 + 4 * i + 2], w[32 + 4 * i + 3]), hhi, acc[1][q]);
        acc[2][q] = ffma2(make_float2(w2.x, w2.y), hlo, acc[2][q]);
        acc[2][q] = ffma2(make_float2(w2.z, w2.w), hhi, acc[2][q]);
        acc[3][q] = ffma2(make_float2(w3.x, w3.y), hlo, acc[3][q]);
        acc[3][q] = ffma2(make_float2(w3.z, w3.w), hhi, acc[3][q]);
      }
    }
    // sum the four k-quarters (lanes kq = 0..3 of a quad); lane kq then stores sequence q = kq
    float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < LNB; ++q) {
      float v[4];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        float x = acc[rr][q].x + acc[rr][q].y;
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        v[rr] = x;
      }
      if (q == kq) mine = make_float4(v[0], v[1], v[2], v[3]);
    }
    stamp(1);
    *reinterpret_cast<float4*>(pre + kq * LG + rg * 4) = mine;
    __syncthreads();
    stamp(2);
    // ---- gates, state update (thread = (u, s)) ----
    const float hprev = hs[hidx(s, u)];
    const float ai = pre[s * LG + u] + gx[0], af = pre[s * LG + LH + u] + gx[1];
    const float ag = pre[s * LG + 2 * LH + u] + gx[2], ao = pre[s * LG + 3 * LH + u] + gx[3];
    const float gi = sigmoidf_(ai), gf = sigmoidf_(af), gg = tanhf(ag), go = sigmoidf_(ao);
    c = gf * c + gi * gg;
    const float h = go * tanhf(c);
    hs[hidx(s, u)] = h;                                   // only this thread read hs[s][u] since the product phase ended
    if (live) {
      float* gp = G + row * GS + dir * LG + u;
      gp[0] = gi; gp[LH] = gf; gp[2 * LH] = gg; gp[3 * LH] = go;
      out[row * 256 + dir * LH + u] = h;
      Cs[row * 256 + dir * LH + u] = c;
      Hp[row * 256 + dir * LH + u] = hprev;
    }
    stamp(3);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// H = 128 backward recurrence.  Same grid.  Walks the time steps in the opposite order of the forward, turns the saved
// gates into pre-activation gradients da (in place in G) and carries dh, dc.
//   dh_prev[q][k] = sum_r W[r][k] da[q][r]  (128 x 4 outputs, contraction 512): thread (kg = tid >> 4, rp = tid & 15) owns
//   the 4 outputs k = 4kg..4kg+3 over the row part [32rp, 32rp+32); the 16 parts are summed with four shuffles.
// ------------------------------------------------------------------------------------------------
constexpr int LDS_ = LNB * 16 * 9 * 4;                                 // da staging: [seq][row part 16][8 float4 + 1 pad]
constexpr int L128B_SMEM = (16 * LG * 4 + LDS_ + LNB * LH) * 4;

__global__ void __launch_bounds__(512, 1) lstm128_bwd_kernel(const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_r,
                                                            float* __restrict__ G, int GS, const float* __restrict__ dout,
                                                            const float* __restrict__ Cs, int B, int T) {
  pdl_sync();
  extern __shared__ __align__(16) float sml[];
  float4* Wsm = reinterpret_cast<float4*>(sml);          // [j = 16][tid 512] : W[32rp + 16 + j][4kg .. 4kg+3]
  float* das = sml + 16 * LG * 4;                        // da of the 4 sequences, padded (see didx)
  float* dhs = das + LDS_;                               // [seq 4][128]
  const int dir = blockIdx.y, b0 = blockIdx.x * LNB;
  const int tid = threadIdx.x;
  const int kg = tid >> 4, rp = tid & 15;
  const float* W = (dir ? w_hh_r : w_hh_f) + (long)(rp * 32) * LH + kg * 4;   // W[32rp + j][4kg ..]
  float4 w[16];                                          // rows j = 0..15 of the part
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    w[j] = __ldg(reinterpret_cast<const float4*>(W + (long)j * LH));
    Wsm[j * LG + tid] = __ldg(reinterpret_cast<const float4*>(W + (long)(16 + j) * LH));
  }
  // da[q][r] lives at das[((q*16 + r/32)*9 + (r%32)/4)*4 + r%4]
  auto didx = [](int q, int r) { return ((q * 16 + (r >> 5)) * 9 + ((r & 31) >> 2)) * 4 + (r & 3); };
  const int u = tid & 127, s = tid >> 7;                 // gate role: (unit, sequence)
  dhs[tid] = 0.f;
  const int b = b0 + s;
  const bool live = b < B;
  float dc = 0.f;
  __syncthreads();
  const float4* d4 = reinterpret_cast<const float4*>(das);
  // the saved state of step + 1 is fetched while step runs (seven loads per thread were issued and consumed in the same
  // step: one exposed L2 / DRAM round trip per time step)
  float n_g[4] = {0.f, 0.f, 0.f, 0.f}, n_ct = 0.f, n_cp = 0.f, n_dy = 0.f;
  auto fetch = [&](int step_) {
    if (live && step_ < T) {
      const int t_ = dir ? step_ : T - 1 - step_;
      const int tp_ = dir ? t_ + 1 : t_ - 1;
      const long row_ = (long)b * T + t_;
      const float* gp_ = G + row_ * GS + dir * LG + u;
      n_g[0] = gp_[0]; n_g[1] = gp_[LH]; n_g[2] = gp_[2 * LH]; n_g[3] = gp_[3 * LH];
      n_ct = Cs[row_ * 256 + dir * LH + u];
      n_cp = (tp_ >= 0 && tp_ < T) ? Cs[((long)b * T + tp_) * 256 + dir * LH + u] : 0.f;
      n_dy = dout[row_ * 256 + dir * LH + u];
    }
  };
  fetch(0);
  for (int step = 0; step < T; ++step) {
    const int t = dir ? step : T - 1 - step;             // reverse of the forward order
    const long row = (long)b * T + t;
    float da[4] = {0.f, 0.f, 0.f, 0.f};
    const float gi = n_g[0], gf = n_g[1], gg = n_g[2], go = n_g[3], ct = n_ct, cp = n_cp, dyv = n_dy;
    fetch(step + 1);
    if (live) {
      float* gp = G + row * GS + dir * LG + u;
      const float dh = dyv + dhs[s * LH + u];
      const float tc = tanhf(ct);
      const float dct = dc + dh * go * (1.f - tc * tc);
      da[0] = dct * gg * gi * (1.f - gi);
      da[1] = dct * cp * gf * (1.f - gf);
      da[2] = dct * gi * (1.f - gg * gg);
      da[3] = dh * tc * go * (1.f - go);
      dc = dct * gf;
      gp[0] = da[0]; gp[LH] = da[1]; gp[2 * LH] = da[2]; gp[3 * LH] = da[3];
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) das[didx(s, g * LH + u)] = da[g];
    __syncthreads();
    // ---- transposed recurrent product over this thread's row part ----
    float2 acc[2][LNB];                                  // [k pair][seq]: (k0,k1) and (k2,k3)
#pragma unroll
    for (int q = 0; q < LNB; ++q) { acc[0][q] = make_float2(0.f, 0.f); acc[1][q] = make_float2(0.f, 0.f); }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 d[LNB];
#pragma unroll
      for (int q = 0; q < LNB; ++q) d[q] = d4[(q * 16 + rp) * 9 + i];
      // rows 4i..4i+3 of the part: registers for i < 4, shared memory for i >= 4
      float4 wr[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) wr[e] = i < 4 ? w[4 * i + e] : Wsm[(4 * (i - 4) + e) * LG + tid];
#pragma unroll
      for (int q = 0; q < LNB; ++q) {
        const float dd[4] = {d[q].x, d[q].y, d[q].z, d[q].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 dv = make_float2(dd[e], dd[e]);
          acc[0][q] = ffma2(make_float2(wr[e].x, wr[e].y), dv, acc[0][q]);
          acc[1][q] = ffma2(make_float2(wr[e].z, wr[e].w), dv, acc[1][q]);
        }
      }
    }
    // sum the 16 row parts (lanes rp = 0..15 of a half-warp); lane rp < 4 then stores sequence q = rp
    float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < LNB; ++q) {
      float v[4] = {acc[0][q].x, acc[0][q].y, acc[1][q].x, acc[1][q].y};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) v[e] += __shfl_xor_sync(0xffffffffu, v[e], o);
      }
      if (q == rp) mine = make_float4(v[0], v[1], v[2], v[3]);
    }
    // all reads of dhs of this step happened before the barrier above
    if (rp < LNB) *reinterpret_cast<float4*>(dhs + rp * LH + kg * 4) = mine;
    __syncthreads();                                      // dhs visible to its (unit, sequence) readers; das free for the next step
  }
}

// ------------------------------------------------------------------------------------------------
// H = 4 (speech_lstm2): 16 gate rows = one half-warp per (sequence, direction); lane j of the group owns gate row j
// (gate j >> 2, unit j & 3), so the gate row of a time step is one coalesced 64-byte access.  State exchange by shuffles;
// the next step's input projection is prefetched while the current one is computed.
//   grid = ceil(2*B / 8) CTAs of 128 threads (8 groups); group id = dir * B + b.
// ------------------------------------------------------------------------------------------------
constexpr int LPF = 8;   // prefetch depth (time steps) of the H = 4 recurrences

__global__ void __launch_bounds__(128) lstm4_fwd_kernel(const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_r,
                                                       float* __restrict__ G, int GS, float* __restrict__ out, float* __restrict__ Cs,
                                                       float* __restrict__ Hp, int B, int T) {
  pdl_sync();
  const int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int j = threadIdx.x & 15;                       // gate row
  const bool live = gid < 2 * B;
  const int dir = live ? gid / B : 0, b = live ? gid - dir * B : 0;
  const unsigned gmask = 0xffffu << (threadIdx.x & 16); // the 16 lanes of this group
  const int lbase = threadIdx.x & 16;                   // first lane of the group inside the warp
  const float* W = (dir ? w_hh_r : w_hh_f) + j * 4;
  const float w0 = __ldg(W), w1 = __ldg(W + 1), w2 = __ldg(W + 2), w3 = __ldg(W + 3);
  float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f, c = 0.f;  // h replicated in every lane, c in the unit lanes (j < 4)
  const long rbase = (long)b * T;
  // The input projections of the next LPF steps are prefetched while the current LPF steps run: with a one-step look-ahead
  // every step waited ~0.5 us for its 64-byte gate row (the recurrence itself is ~200 dependent cycles per step).
  float gbuf[LPF];
#pragma unroll
  for (int u = 0; u < LPF; ++u) gbuf[u] = (live && u < T) ? G[(rbase + (dir ? T - 1 - u : u)) * GS + dir * 16 + j] : 0.f;
  for (int step0 = 0; step0 < T; step0 += LPF) {
    float gcur[LPF];
#pragma unroll
    for (int u = 0; u < LPF; ++u) gcur[u] = gbuf[u];
#pragma unroll
    for (int u = 0; u < LPF; ++u) {
      const int sn = step0 + LPF + u;
      if (live && sn < T) gbuf[u] = G[(rbase + (dir ? T - 1 - sn : sn)) * GS + dir * 16 + j];
    }
#pragma unroll
    for (int uu = 0; uu < LPF; ++uu) {
    const int step = step0 + uu;
    if (step >= T) break;
    const int t = dir ? T - 1 - step : step;
    const long row = rbase + t;
    const float gx = gcur[uu];
    float a = gx;
    a = fmaf(w0, h0, a); a = fmaf(w1, h1, a); a = fmaf(w2, h2, a); a = fmaf(w3, h3, a);
    const float act = (j >> 2) == 2 ? tanhf(a) : sigmoidf_(a);
    // unit lane u (= j & 3) gathers i, f, g, o of its unit
    const int u = j & 3;
    const float gi = __shfl_sync(gmask, act, lbase + u), gf = __shfl_sync(gmask, act, lbase + 4 + u);
    const float gg = __shfl_sync(gmask, act, lbase + 8 + u), go = __shfl_sync(gmask, act, lbase + 12 + u);
    c = gf * c + gi * gg;                               // identical in the 4 lanes sharing a unit; lanes j < 4 are authoritative
    const float hn = go * tanhf(c);
    if (live) {
      G[row * GS + dir * 16 + j] = act;
      if (j < 4) {
        const float hp = j == 0 ? h0 : j == 1 ? h1 : j == 2 ? h2 : h3;
        Hp[row * 8 + dir * 4 + j] = hp;
        out[row * 8 + dir * 4 + j] = hn;
        Cs[row * 8 + dir * 4 + j] = c;
      }
    }
    h0 = __shfl_sync(gmask, hn, lbase + 0); h1 = __shfl_sync(gmask, hn, lbase + 1);
    h2 = __shfl_sync(gmask, hn, lbase + 2); h3 = __shfl_sync(gmask, hn, lbase + 3);
    }
  }
}

// backward: lane j owns gate row j; also accumulates dW_hh[j][0..3] (reduced over the CTA, then atomics; pre-zeroed)
__global__ void __launch_bounds__(128) lstm4_bwd_kernel(const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_r,
                                                       float* __restrict__ G, int GS, const float* __restrict__ dout,
                                                       const float* __restrict__ Cs, const float* __restrict__ Hp,
                                                       float* __restrict__ dwhh_f, float* __restrict__ dwhh_r, int B, int T) {
  pdl_sync();
  __shared__ float red[2][64];
  const int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int j = threadIdx.x & 15;
  const bool live = gid < 2 * B;
  const int dir = live ? gid / B : 0, b = live ? gid - dir * B : 0;
  const unsigned gmask = 0xffffu << (threadIdx.x & 16);
  const int lbase = threadIdx.x & 16;
  const float* W = (dir ? w_hh_r : w_hh_f) + j * 4;
  const float w0 = __ldg(W), w1 = __ldg(W + 1), w2 = __ldg(W + 2), w3 = __ldg(W + 3);
  const int u = j & 3, gate = j >> 2;
  float dw0 = 0.f, dw1 = 0.f, dw2 = 0.f, dw3 = 0.f;
  float dh = 0.f, dc = 0.f;                              // of unit u, replicated in the 4 lanes sharing it
  const long rbase = (long)b * T;
  if (threadIdx.x < 128) { red[0][threadIdx.x & 63] = 0.f; red[1][threadIdx.x & 63] = 0.f; }
  __syncthreads();
  // block prefetch of the saved state (see lstm4_fwd_kernel): the five loads of a step used to be issued and consumed in the
  // same step -- one exposed L2 / DRAM round trip (~0.8 us) per time step
  float b_act[LPF], b_ct[LPF], b_cp[LPF], b_dy[LPF], b_hp[LPF];
  auto fetch = [&](int s_, float& act_, float& ct_, float& cp_, float& dy_, float& hp_) {
    act_ = 0.f; ct_ = 0.f; cp_ = 0.f; dy_ = 0.f; hp_ = 0.f;
    if (live && s_ < T) {
      const int t_ = dir ? s_ : T - 1 - s_;
      const int tp_ = dir ? t_ + 1 : t_ - 1;
      const long row_ = rbase + t_;
      act_ = G[row_ * GS + dir * 16 + j];
      ct_ = Cs[row_ * 8 + dir * 4 + u];
      cp_ = (tp_ >= 0 && tp_ < T) ? Cs[(rbase + tp_) * 8 + dir * 4 + u] : 0.f;
      dy_ = dout[row_ * 8 + dir * 4 + u];
      hp_ = Hp[row_ * 8 + dir * 4 + u];
    }
  };
#pragma unroll
  for (int q = 0; q < LPF; ++q) fetch(q, b_act[q], b_ct[q], b_cp[q], b_dy[q], b_hp[q]);
  for (int step0 = 0; step0 < T; step0 += LPF) {
    float c_act[LPF], c_ct[LPF], c_cp[LPF], c_dy[LPF], c_hp[LPF];
#pragma unroll
    for (int q = 0; q < LPF; ++q) { c_act[q] = b_act[q]; c_ct[q] = b_ct[q]; c_cp[q] = b_cp[q]; c_dy[q] = b_dy[q]; c_hp[q] = b_hp[q]; }
#pragma unroll
    for (int q = 0; q < LPF; ++q) fetch(step0 + LPF + q, b_act[q], b_ct[q], b_cp[q], b_dy[q], b_hp[q]);
#pragma unroll
    for (int q = 0; q < LPF; ++q) {
    const int step = step0 + q;
    if (step >= T) break;
    const int t = dir ? step : T - 1 - step;
    const long row = rbase + t;
    const float act = c_act[q], ct = c_ct[q], cp = c_cp[q], dy = c_dy[q], hp = c_hp[q];
    const float gi = __shfl_sync(gmask, act, lbase + u), gf = __shfl_sync(gmask, act, lbase + 4 + u);
    const float gg = __shfl_sync(gmask, act, lbase + 8 + u), go = __shfl_sync(gmask, act, lbase + 12 + u);
    const float dht = dy + dh;
    const float tc = tanhf(ct);
    const float dct = dc + dht * go * (1.f - tc * tc);
    // pre-activation gradient of THIS lane's gate row
    float da;
    if (gate == 0) da = dct * gg * gi * (1.f - gi);
    else if (gate == 1) da = dct * cp * gf * (1.f - gf);
    else if (gate == 2) da = dct * gi * (1.f - gg * gg);
    else da = dht * tc * go * (1.f - go);
    dc = dct * gf;
    if (live) G[row * GS + dir * 16 + j] = da;
    // dW_hh[j][k] += da * h_prev[k]  (h_prev[k] lives in the lanes with u == k)
    dw0 = fmaf(da, __shfl_sync(gmask, hp, lbase + 0), dw0); dw1 = fmaf(da, __shfl_sync(gmask, hp, lbase + 1), dw1);
    dw2 = fmaf(da, __shfl_sync(gmask, hp, lbase + 2), dw2); dw3 = fmaf(da, __shfl_sync(gmask, hp, lbase + 3), dw3);
    // dh_prev[k] = sum_j W[j][k] da[j] : reduce over the 16 lanes, lane ends up with the value of its unit u
    float p0 = w0 * da, p1 = w1 * da, p2 = w2 * da, p3 = w3 * da;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      p0 += __shfl_xor_sync(gmask, p0, o); p1 += __shfl_xor_sync(gmask, p1, o);
      p2 += __shfl_xor_sync(gmask, p2, o); p3 += __shfl_xor_sync(gmask, p3, o);
    }
    dh = u == 0 ? p0 : u == 1 ? p1 : u == 2 ? p2 : p3;
    }
  }
  // CTA-level reduction of dW_hh per direction (a CTA may straddle the two directions)
  if (live) {
    float* r = red[dir];
    atomicAdd(r + j * 4 + 0, dw0); atomicAdd(r + j * 4 + 1, dw1); atomicAdd(r + j * 4 + 2, dw2); atomicAdd(r + j * 4 + 3, dw3);
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    if (red[0][threadIdx.x] != 0.f) atomicAdd(dwhh_f + threadIdx.x, red[0][threadIdx.x]);
    if (red[1][threadIdx.x] != 0.f) atomicAdd(dwhh_r + threadIdx.x, red[1][threadIdx.x]);
  }
}

}  // namespace lstm
}  // namespace eegclip
