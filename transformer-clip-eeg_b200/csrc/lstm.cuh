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
// H = 128 on the warp-level tensor cores (mma.sync.m16n8k16, split-bf16: W_hi.h_hi + W_hi.h_lo + W_lo.h_hi, fp32 accumulate).
// The recurrent product of a time step is pre^T (512 gate rows x 8 sequences) = W_hh (512 x 128) . h^T (128 x 8): the gate
// rows are the M side (W_hh fragments stay ON CHIP for the whole sequence: the hi halves in registers, 128 per thread, the lo
// halves in 128 KB of shared memory in fragment order), the 8 sequences of a CTA are the N = 8 side, so no MMA row is padding.
// A per-step latency chain bounds a 320-step recurrence, not throughput: 8 sequences per CTA put the 2 x 256 sequences of
// the benchmarked batch on 64 SMs at ~1/3 of the FFMA2 kernel's time per step (2.8 us for 4 sequences, issue-bound on 256
// packed FMAs + 48 LDS per thread).  tcgen05 is not used on purpose: its smallest shapes (M = 128 x N = 16) cost ~31 cycles per
// instruction x 96 instructions per step on one issuing thread, plus a TMEM round trip per step.
//   block 256 = 8 warps.  Forward: warp w owns hidden units [16w, 16w+16) = 4 m-tiles (sub-block sb, gate pair): tile rows
//   0-7 = gate i (g) of units 16w+8sb+r, rows 8-15 = gate f (o) of the same units, so a thread's accumulators hold i, f, g, o
//   of ITS (unit, sequence) pairs and the cell update is thread-local (4 pairs per thread: unit 16w+8sb+(lane>>2), sequence
//   2(lane&3)+e).  h_t goes back to shared memory as bf16 hi / lo planes in B-fragment order (double buffered: one barrier
//   per step).  The contraction index is permuted (k-step s, fragment column 2t+e [+8] <-> unit 32t+4s+e [+2]) so that a
//   thread's B fragments of all 8 k-steps are 64 contiguous bytes.
//   Backward: warp w owns the m-tile of hidden units [16w, 16w+16) over all 512 gate rows (32 k-steps), so dh_{t-1} of a
//   thread's 4 (unit, sequence) pairs never leaves its registers; da goes to shared memory the same way (k-step s, column
//   2t+e [+8] <-> gate t, unit 4s+e [+2]).
// ------------------------------------------------------------------------------------------------
constexpr int LMS = 8;                                                // sequences per CTA (the MMA's N)
constexpr int LM_WLO = 8 * 32 * 32 * 16;                              // W_lo fragments: [warp 8][32 (tile, k-step)][lane 32] uint4

__device__ __forceinline__ void mma_bf16(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// split a pair of fp32 values into packed bf16 (hi, lo): element 0 in the low half
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float fa = __uint_as_float(hi << 16), fb = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - fb), "f"(a - fa));
}
__device__ __forceinline__ float ex2a(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcpa(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// four / five instructions each; absolute error ~1e-7 (ex2.approx 2^-22 relative, rcp.approx 1 ulp)
__device__ __forceinline__ float sigmoid_fast(float x) { return rcpa(1.f + ex2a(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(-2.f, rcpa(1.f + ex2a(2.8853900817779268f * x)), 1.f); }
constexpr int LGS = 8 * LH;                                          // row stride of the gate buffer at H = 128
// Global traffic goes through shared-memory stages so that every global access is a full 512-byte row piece per warp: an
// accumulator fragment owns (8 units) x (2 sequences) per register, i.e. 32-byte pieces of 4 different rows per store
// instruction -- 28 such stores per thread and step held the recurrence at 2.15 us per step (1.8 us without them).
//   forward : x-projections of step s+1 arrive by cp.async while step s runs; gates / h / c of step s are staged and written
//             out by (warp = sequence, lane = 4 units) at the start of step s+1 (h_{t-1} for Hp is that thread's previous h).
//   backward: saved gates by cp.async, da staged and written out under the MMA phase of the same step.
constexpr int LOS = 7 * LH + 4;                                       // forward out-stage row [i f g o h c h_prev][128]; +4: bank shift per sequence
constexpr int LXS = 4 * LH + 4;                                       // gate-row stage [4][128]
constexpr int LM_HB = 2 * 2 * 128 * 16;                               // h planes: [buffer][hi, lo][128 uint4]
constexpr int LM_DB = 2 * 2 * 512 * 16;                               // da planes
constexpr int L128M_SMEM = LM_WLO + LM_HB + 2 * LMS * LOS * 4 + 2 * LMS * LXS * 4;
constexpr int L128MB_SMEM = LM_WLO + LM_DB + 2 * 2 * LMS * LXS * 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// shared -> global through the TMA engine (1-D bulk copy): the stores of a step do not occupy LSU slots of the 8 warps (a warp
// sustains ~2.4 B/clk of st.global: the 28 KB of a step took 0.55 us as float4 stores)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"((uint32_t)__cvta_generic_to_shared(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(256, 1) lstm128_fwd_mma_kernel(const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_r,
                                                                float* __restrict__ G, float* __restrict__ out,
                                                                float* __restrict__ Cs, float* __restrict__ Hp, int B, int T, unsigned long long* dbg) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t smm[];
  // development timeline: thread 0 of CTA (0,0) stamps the phases of the first steps (tools/time_lstm.py timeline)
  unsigned long long* dbgp = (dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) ? dbg : nullptr;
  int dbn = 0;
  auto stamp = [&](int ev) {
    if (dbgp && dbn < 120) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); dbgp[2 * dbn] = (unsigned long long)ev; dbgp[2 * dbn + 1] = t_; ++dbn; dbgp[255] = (unsigned long long)dbn; }
  };
  uint4* Wlo = reinterpret_cast<uint4*>(smm);
  uint4* hbuf = reinterpret_cast<uint4*>(smm + LM_WLO);
  float* ostage = reinterpret_cast<float*>(smm + LM_WLO + LM_HB);      // [2][8][LOS]
  float* xstage = ostage + 2 * LMS * LOS;                             // [2][8][LXS]
  const int dir = blockIdx.y, b0 = blockIdx.x * LMS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const float* W = dir ? w_hh_r : w_hh_f;
  // ---- copy role: warp = sequence, lane = 4 consecutive units.  Sequences past the batch read sequence b0's rows (valid
  //      memory, results unused) and store nothing. ----
  const bool clive = b0 + warp < B;
  const int t0 = dir ? T - 1 : 0;
  const long dstep = dir ? -1 : 1;
  const long crow = (long)(clive ? b0 + warp : b0) * T + t0;          // row of step 0; step s is row crow + s * dstep
  auto fetch_x = [&](int step_) {
    const float* src = G + (crow + step_ * dstep) * LGS + dir * LG + 4 * lane;
    float* dst = xstage + ((step_ & 1) * LMS + warp) * LXS + 4 * lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) cp_async16(dst + k * LH, src + k * LH);
  };
  fetch_x(0);
  // ---- A fragments: tile tl = 2 sb + pair; a0/a2 = row (gate 2 pair, unit), a1/a3 = row (gate 2 pair + 1, unit) ----
  uint32_t whi[4][8][4];
#pragma unroll
  for (int tl = 0; tl < 4; ++tl) {
    const int unit = 16 * warp + 8 * (tl >> 1) + g;
    const float* r0p = W + (long)((2 * (tl & 1)) * LH + unit) * LH + 32 * tig;
    const float* r1p = r0p + (long)LH * LH;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const float4 r0 = __ldg(reinterpret_cast<const float4*>(r0p) + ks), r1 = __ldg(reinterpret_cast<const float4*>(r1p) + ks);
      uint4 lo;
      split_pair(r0.x, r0.y, whi[tl][ks][0], lo.x);
      split_pair(r1.x, r1.y, whi[tl][ks][1], lo.y);
      split_pair(r0.z, r0.w, whi[tl][ks][2], lo.z);
      split_pair(r1.z, r1.w, whi[tl][ks][3], lo.w);
      Wlo[((warp * 4 + tl) * 8 + ks) * 32 + lane] = lo;
    }
  }
  for (int i = tid; i < 2 * 2 * 128; i += 256) hbuf[i] = make_uint4(0u, 0u, 0u, 0u);
  // compute role: (unit u = 16w + 8sb + g, sequence n = 2tig + e).  h of (u, n) lives at bf16 index
  // (((u % 32) / 8 * 8 + n) * 4 + u / 32) * 8 + u % 8 of a plane: chunk 2(w&1) + sb, reader lane-quarter w >> 1, element g
  const int hoff = ((2 * (warp & 1) * 8 + 2 * tig) * 4 + (warp >> 1)) * 8 + g;      // + sb * 256 + e * 32 (bf16 elements)
  const int xo = 2 * tig * LXS + 16 * warp + g;                                    // + e * LXS + gate * LH + 8 sb
  const int oo = 2 * tig * LOS + 16 * warp + g;                                    // + e * LOS + array * LH + 8 sb
  float c[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float hprev[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  // lane 0 of warp w writes sequence w's staged rows of a step: gates (2 KB contiguous), h, c, h_{t-1}
  auto copy_out = [&](int step_) {
    if (lane == 0 && clive) {
      const float* src = ostage + ((step_ & 1) * LMS + warp) * LOS;
      const long row = crow + step_ * dstep;
      bulk_s2g(G + row * LGS + dir * LG, src, 4 * LH * 4);
      const long o = row * 256 + dir * LH;
      bulk_s2g(out + o, src + 4 * LH, LH * 4);
      bulk_s2g(Cs + o, src + 5 * LH, LH * 4);
      bulk_s2g(Hp + o, src + 6 * LH, LH * 4);
      bulk_commit();
    }
  };
  cp_async_wait_all();
  __syncthreads();
  for (int step = 0; step < T; ++step) {
    stamp(0);
    if (step + 1 < T) fetch_x(step + 1);
    stamp(5);
    const uint4* hh = hbuf + (step & 1) * 256;
    const uint4* hl = hh + 128;
    float acc[4][2][4];                                   // [tile][hi.hi (+ x-projection), hi.lo + lo.hi]
    {
      // accumulator j of tile tl = (gate 2 (tl & 1) + (j >> 1), sequence 2 tig + (j & 1))
      const float* xs = xstage + (step & 1) * LMS * LXS + xo;
#pragma unroll
      for (int tl = 0; tl < 4; ++tl)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[tl][0][j] = xs[(j & 1) * LXS + (2 * (tl & 1) + (j >> 1)) * LH + 8 * (tl >> 1)];
          acc[tl][1][j] = 0.f;
        }
    }
    if (dbgp) { if (acc[0][0][0] + acc[3][0][3] == 123.456f) dbgp[254] = 1; stamp(6); }
    if (step > 0) copy_out(step - 1);
    stamp(1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 vh = hh[(i * 8 + g) * 4 + tig], vl = hl[(i * 8 + g) * 4 + tig];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int ks = 2 * i + kk;
        const uint32_t bh0 = kk ? vh.z : vh.x, bh1 = kk ? vh.w : vh.y, bl0 = kk ? vl.z : vl.x, bl1 = kk ? vl.w : vl.y;
#pragma unroll
        for (int tl = 0; tl < 4; ++tl) {
          const uint4 al = Wlo[((warp * 4 + tl) * 8 + ks) * 32 + lane];
          const uint32_t alo[4] = {al.x, al.y, al.z, al.w};
          mma_bf16(acc[tl][0], whi[tl][ks], bh0, bh1);
          mma_bf16(acc[tl][1], whi[tl][ks], bl0, bl1);
          mma_bf16(acc[tl][1], alo, bh0, bh1);
        }
      }
    }
    // ---- gates and state of this thread's 4 (unit, sequence) pairs ----
    if (dbgp) { if (acc[0][0][0] + acc[3][1][3] == 123.456f) dbgp[254] = 1; stamp(2); }
    uint16_t* nh = reinterpret_cast<uint16_t*>(hbuf + ((step + 1) & 1) * 256) + hoff;
    float* os = ostage + (step & 1) * LMS * LOS + oo;
#pragma unroll
    for (int sb = 0; sb < 2; ++sb) {
      float hn[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float gi = sigmoid_fast(acc[2 * sb][0][e] + acc[2 * sb][1][e]);
        const float gf = sigmoid_fast(acc[2 * sb][0][2 + e] + acc[2 * sb][1][2 + e]);
        const float gg = tanh_fast(acc[2 * sb + 1][0][e] + acc[2 * sb + 1][1][e]);
        const float go = sigmoid_fast(acc[2 * sb + 1][0][2 + e] + acc[2 * sb + 1][1][2 + e]);
        const float cn = fmaf(gf, c[sb][e], gi * gg);
        c[sb][e] = cn;
        hn[e] = go * tanh_fast(cn);
        float* op = os + e * LOS + 8 * sb;
        op[0] = gi; op[LH] = gf; op[2 * LH] = gg; op[3 * LH] = go; op[4 * LH] = hn[e]; op[5 * LH] = cn; op[6 * LH] = hprev[sb][e];
        hprev[sb][e] = hn[e];
      }
      uint32_t hi, lo;
      split_pair(hn[0], hn[1], hi, lo);
      nh[sb * 256] = (uint16_t)(hi & 0xffffu);
      nh[sb * 256 + 32] = (uint16_t)(hi >> 16);
      nh[128 * 8 + sb * 256] = (uint16_t)(lo & 0xffffu);
      nh[128 * 8 + sb * 256 + 32] = (uint16_t)(lo >> 16);
    }
    stamp(3);
    fence_async_smem();                                   // staged rows -> visible to the bulk copies issued after the barrier
    cp_async_wait_all();
    if (lane == 0) bulk_wait_read();                      // the previous step's stage may be rewritten after the barrier
    stamp(4);
    __syncthreads();
  }
  copy_out(T - 1);
  if (lane == 0) bulk_wait_all();
}

__global__ void __launch_bounds__(256, 1) lstm128_bwd_mma_kernel(const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_r,
                                                                float* __restrict__ G, const float* __restrict__ dout,
                                                                const float* __restrict__ Cs, int B, int T) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t smm[];
  uint4* Wlo = reinterpret_cast<uint4*>(smm);
  uint4* dbuf = reinterpret_cast<uint4*>(smm + LM_WLO);
  float* gstage = reinterpret_cast<float*>(smm + LM_WLO + LM_DB);      // saved gates in:  [2][8][LXS]
  float* dstage = gstage + 2 * LMS * LXS;                             // da out:          [2][8][LXS]
  const int dir = blockIdx.y, b0 = blockIdx.x * LMS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const float* W = dir ? w_hh_r : w_hh_f;
  // ---- copy role (warp = sequence, lane = 4 units) ----
  const bool clive = b0 + warp < B;
  const int t0 = dir ? 0 : T - 1;                                                   // reverse of the forward order
  const long dstep = dir ? 1 : -1;
  const long crow = (long)(clive ? b0 + warp : b0) * T + t0;
  auto fetch_g = [&](int step_) {
    const float* src = G + (crow + step_ * dstep) * LGS + dir * LG + 4 * lane;
    float* dst = gstage + ((step_ & 1) * LMS + warp) * LXS + 4 * lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) cp_async16(dst + k * LH, src + k * LH);
  };
  fetch_g(0);
  // ---- A fragments of W^T: rows = units 16w+g (a0, a2) and 16w+8+g (a1, a3); k-step s column 2tig+e [+8] <-> gate row 128 tig + 4s + e [+2]
  uint32_t whi[32][4];
  {
    const float* wc = W + (long)(128 * tig) * LH + 16 * warp + g;
#pragma unroll
    for (int ks = 0; ks < 32; ++ks) {
      const float* p = wc + (long)(4 * ks) * LH;
      uint4 lo;
      split_pair(__ldg(p), __ldg(p + LH), whi[ks][0], lo.x);
      split_pair(__ldg(p + 8), __ldg(p + LH + 8), whi[ks][1], lo.y);
      split_pair(__ldg(p + 2 * LH), __ldg(p + 3 * LH), whi[ks][2], lo.z);
      split_pair(__ldg(p + 2 * LH + 8), __ldg(p + 3 * LH + 8), whi[ks][3], lo.w);
      Wlo[(warp * 32 + ks) * 32 + lane] = lo;
    }
  }
  // compute role: pairs (unit 16w + 8ub + g, sequence 2tig + e).  da of (gate gt, unit u, sequence n) lives at bf16 index
  // ((u / 8 * 8 + n) * 4 + gt) * 8 + u % 8 of a plane
  const int doff = ((2 * warp * 8 + 2 * tig) * 4) * 8 + g;                           // + ub * 256 + e * 32 + gt * 8 (bf16 elements)
  const int xo = 2 * tig * LXS + 16 * warp + g;                                     // + e * LXS + gate * LH + 8 ub
  const bool live0 = b0 + 2 * tig < B, live1 = b0 + 2 * tig + 1 < B;
  const int col = dir * LH + 16 * warp + g;
  long so[2] = {((long)(live0 ? b0 + 2 * tig : b0) * T + t0) * 256 + col, ((long)(live1 ? b0 + 2 * tig + 1 : b0) * T + t0) * 256 + col};
  float dh[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, dc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float n_cp[2][2], n_dy[2][2], ct[2][2];
  // dy and c_{t-1} of the step whose state row is o; `has_prev`: a forward-earlier step exists
  auto fetch_s = [&](const long* o, bool has_prev) {
#pragma unroll
    for (int ub = 0; ub < 2; ++ub)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        n_dy[ub][e] = dout[o[e] + 8 * ub];
        n_cp[ub][e] = has_prev ? Cs[o[e] + dstep * 256 + 8 * ub] : 0.f;
      }
  };
#pragma unroll
  for (int ub = 0; ub < 2; ++ub)
#pragma unroll
    for (int e = 0; e < 2; ++e) ct[ub][e] = Cs[so[e] + 8 * ub];
  fetch_s(so, T > 1);
  cp_async_wait_all();
  __syncthreads();
  for (int step = 0; step < T; ++step) {
    float dyv[2][2], cpv[2][2];
#pragma unroll
    for (int ub = 0; ub < 2; ++ub)
#pragma unroll
      for (int e = 0; e < 2; ++e) { dyv[ub][e] = n_dy[ub][e]; cpv[ub][e] = n_cp[ub][e]; }
    if (step + 1 < T) {                                    // uniform
      fetch_g(step + 1);
      so[0] += dstep * 256; so[1] += dstep * 256;
      fetch_s(so, step + 2 < T);
    }
    // ---- gate phase: da of this thread's 4 pairs ----
    uint16_t* dw = reinterpret_cast<uint16_t*>(dbuf + (step & 1) * 1024) + doff;
    const float* gs = gstage + (step & 1) * LMS * LXS + xo;
    float* ds = dstage + (step & 1) * LMS * LXS + xo;
    float da[2][2][4];
#pragma unroll
    for (int ub = 0; ub < 2; ++ub)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float* gp = gs + e * LXS + 8 * ub;
        const float gi = gp[0], gf = gp[LH], gg = gp[2 * LH], go = gp[3 * LH], cp = cpv[ub][e];
        const float dht = dyv[ub][e] + dh[ub][e];
        const float tc = tanh_fast(ct[ub][e]);
        const float dct = fmaf(dht * go, fmaf(-tc, tc, 1.f), dc[ub][e]);
        da[ub][e][0] = dct * gg * gi * (1.f - gi);
        da[ub][e][1] = dct * cp * gf * (1.f - gf);
        da[ub][e][2] = dct * gi * fmaf(-gg, gg, 1.f);
        da[ub][e][3] = dht * tc * go * (1.f - go);
        dc[ub][e] = dct * gf;
        ct[ub][e] = cp;                                    // c_{t-1} is the next step's c_t
        float* dp = ds + e * LXS + 8 * ub;
        dp[0] = da[ub][e][0]; dp[LH] = da[ub][e][1]; dp[2 * LH] = da[ub][e][2]; dp[3 * LH] = da[ub][e][3];
      }
#pragma unroll
    for (int ub = 0; ub < 2; ++ub)
#pragma unroll
      for (int gt = 0; gt < 4; ++gt) {
        uint32_t hi, lo;
        split_pair(da[ub][0][gt], da[ub][1][gt], hi, lo);
        dw[ub * 256 + gt * 8] = (uint16_t)(hi & 0xffffu);
        dw[ub * 256 + gt * 8 + 32] = (uint16_t)(hi >> 16);
        dw[512 * 8 + ub * 256 + gt * 8] = (uint16_t)(lo & 0xffffu);
        dw[512 * 8 + ub * 256 + gt * 8 + 32] = (uint16_t)(lo >> 16);
      }
    fence_async_smem();
    cp_async_wait_all();
    if (lane == 0) bulk_wait_read();                       // step - 1's da stage is rewritten in step + 1, after this barrier
    __syncthreads();
    // ---- da of this step out to G: one 2 KB bulk copy per sequence (lane 0 of warp = sequence), under the MMA phase ----
    if (lane == 0 && clive) {
      bulk_s2g(G + (crow + step * dstep) * LGS + dir * LG, dstage + ((step & 1) * LMS + warp) * LXS, 4 * LH * 4);
      bulk_commit();
    }
    // ---- dh_{t-1}^T (this warp's 16 units x 8 sequences) = W^T . da^T over the 512 gate rows ----
    const uint4* dah = dbuf + (step & 1) * 1024;
    const uint4* dal = dah + 512;
    float acc[4][4];                                       // hi.hi and (hi.lo + lo.hi), each over even / odd k-steps
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][j] = 0.f;
#pragma unroll
    for (int cch = 0; cch < 16; ++cch) {
      const uint4 vh = dah[(cch * 8 + g) * 4 + tig], vl = dal[(cch * 8 + g) * 4 + tig];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int ks = 2 * cch + kk;
        const uint32_t bh0 = kk ? vh.z : vh.x, bh1 = kk ? vh.w : vh.y, bl0 = kk ? vl.z : vl.x, bl1 = kk ? vl.w : vl.y;
        const uint4 al = Wlo[(warp * 32 + ks) * 32 + lane];
        const uint32_t alo[4] = {al.x, al.y, al.z, al.w};
        mma_bf16(acc[kk], whi[ks], bh0, bh1);
        mma_bf16(acc[2 + kk], whi[ks], bl0, bl1);
        mma_bf16(acc[2 + kk], alo, bh0, bh1);
      }
    }
    // accumulator j: unit 16w + 8 (j >> 1) + g, sequence 2 tig + (j & 1)
#pragma unroll
    for (int ub = 0; ub < 2; ++ub)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 2 * ub + e;
        dh[ub][e] = (acc[0][j] + acc[1][j]) + (acc[2][j] + acc[3][j]);
      }
  }
  if (lane == 0) bulk_wait_all();
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
