// Bidirectional single-layer LSTM of the default speech tower (clip_model.py:267-268, 322-323; nn.LSTM semantics: gate order
// i,f,g,o, zero initial state, batch_first, outputs of the two directions concatenated).
//
// The input projections x.W_ih^T (+ both biases) of ALL time steps are one token GEMM (lin_tc.cuh), and so are the weight
// gradients and the input gradient; only the recurrence itself is sequential.  The recurrence keeps W_hh on chip:
//   H = 128 (speech_lstm1): a cluster of two CTAs per 8 sequences of one direction on the warp-level tensor cores
//       (mma.sync.m16n8k16, split-bf16), W_hh register-resident -- see the banner below.
//   H = 4 (speech_lstm2): one half-warp per (sequence, direction), lane = gate row, coalesced 64-byte gate rows, shuffles.
// Saved for the backward: post-activation gates (in place over the input projections), cell states, h_{t-1}.
#pragma once
#include "common.cuh"

namespace eegclip {
namespace lstm {

constexpr int LH = 128, LG = 512;

// ------------------------------------------------------------------------------------------------
// H = 128 on the warp-level tensor cores (mma.sync.m16n8k16, split-bf16: W_hi.h_hi + W_hi.h_lo + W_lo.h_hi, fp32 accumulate;
// measured against the exact-fp32 recurrence this replaces: |dy| <= 1e-6, dx 2e-6 relative).
// The recurrent product of a time step is pre^T (512 gate rows x 8 sequences) = W_hh (512 x 128) . h^T (128 x 8): the gate
// rows are the M side, so W_hh fragments stay in registers for the whole sequence, and the 8 sequences of a cluster are the
// N = 8 side: no MMA row is padding.  A per-step latency chain bounds a 320-step recurrence, not throughput.
// History (B = 256, T = 320, both directions, tools/time_lstm.py; forward / backward recurrence):
//   fp32 FFMA2, 4 sequences per CTA, 128 CTAs (round 1)                                   906 / 885 us   (2.8 us per step)
//   mma.sync, 8 sequences per CTA, W_lo streamed from shared memory (128 KB per step)     707 / 880 us   shared-memory pipe bound
//   mma.sync, cluster of 2 CTAs per 8 sequences, everything in registers (this file)      443 / 530 us   (1.4 us per step)
//   (the same kernels without any global load / store: 313 / 377 us -- MMA phase 0.58 us + gates 0.19 us + barrier per step)
// tools/micro/hmma_probe.cu: mma.sync bf16 m16n8k16 has 20 clk dependent latency and 2048 dense FLOP/clk/SM (8 clk per
// instruction and SM sub-partition; TF32 m16n8k8 half of that).  tcgen05 is not used on purpose: its smallest shapes
// (M = 128 x N = 16) cost ~31 cycles per instruction x 96 instructions per step on one issuing thread, plus a TMEM round trip.
//   The contraction index is permuted (k-step s, fragment column 2t+e [+8] <-> unit 32t+4s+e [+2]) so that a thread's B
//   fragments of all 8 k-steps are 64 contiguous bytes of the bf16 hi / lo operand planes in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int LMS = 8;                                                // sequences per cluster (the MMA's N)

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

// ------------------------------------------------------------------------------------------------
// The same recurrence on a CLUSTER OF TWO CTAs per 8 sequences.  One SM cannot hold W_hh as bf16 hi + lo in registers (256 KB
// = the whole register file), so the single-CTA kernels above stream the lo halves from shared memory every step (128 KB per
// step: the shared-memory pipe, not the tensor pipe, bounds them at ~2.1 us per step).  Split over two SMs everything is
// register-resident (64 + 64 registers per thread), the MMA work per SM halves, and 128 instead of 64 SMs work:
//   forward : CTA r owns hidden units [64r, 64r+64) (all four gates: 256 gate rows, K = 128).  Each CTA needs the whole
//             h_{t-1}: a thread writes its h values (bf16 hi / lo) into its own AND the peer's operand planes (DSMEM, 2 KB per
//             step and CTA); one cluster barrier per step (planes double buffered).
//   backward: split over K: CTA r contracts over ITS 256 gate rows (its units' da, produced locally -- no da exchange) for
//             all 128 units and sends the half of the partial dh_{t-1} that the peer owns (2 KB per step, DSMEM).
// The hand-over is st.async + mbarrier (the DSMEM store itself signals the destination CTA's barrier with its byte count), one
// CTA-local __syncthreads and one mbarrier wait per step.  With st.shared::cluster + barrier.cluster the arrive's release
// waited for EVERY memory operation of the thread in flight, the step's global loads / stores included: 0.85 us of a 1.85 us
// step, 0.45 us of 1.4 us with the stores moved behind the arrive.  Global accesses are per-thread 32-byte pieces (8 loads +
// 14 stores per thread and step forward, 12 + 8 backward).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_peer(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_u16(uint32_t addr, uint32_t v) { asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ void st_cluster_f32x2(uint32_t addr, float a, float b) { asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory"); }
// DSMEM store that signals the DESTINATION CTA's mbarrier with its byte count (st.async): the hand-over needs no release
// fence on the sending thread, so the thread's global loads / stores in flight do not hold the step
__device__ __forceinline__ void st_async_b32(uint32_t addr, uint32_t v, uint32_t mbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(v), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async_f32x2(uint32_t addr, float a, float b, uint32_t mbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(addr), "f"(a), "f"(b), "r"(mbar) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive(); cluster_wait(); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
    lstm128_fwd_c2_kernel(const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_r, float* __restrict__ G, float* __restrict__ out,
                          float* __restrict__ Cs, float* __restrict__ Hp, int B, int T, unsigned long long* dbg) {
  pdl_sync();
  // development timeline: thread 0 of CTA (0,0) stamps the phases of the first steps (tools/time_lstm.py)
  unsigned long long* dbgp = (dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) ? dbg : nullptr;
  int dbn = 0;
  auto stamp = [&](int ev) {
    if (dbgp && dbn < 120) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); dbgp[2 * dbn] = (unsigned long long)ev; dbgp[2 * dbn + 1] = t_; ++dbn; dbgp[255] = (unsigned long long)dbn; }
  };
  __shared__ __align__(16) uint4 hbuf[2 * 2 * 128];        // h planes: [buffer][hi, lo][128 uint4], all 128 units
  __shared__ __align__(8) uint64_t hbar[2];                // per buffer: the peer's 2 KB of h have landed
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) { tc::mbar_init(&hbar[0], 1); tc::mbar_init(&hbar[1], 1); tc::mbar_fence_init(); }
  const int dir = blockIdx.y, b0 = (blockIdx.x >> 1) * LMS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const float* W = dir ? w_hh_r : w_hh_f;
  const int unit = 64 * (int)rank + 8 * warp + g;
  // ---- A fragments: tile p: rows 0-7 = gate 2p, rows 8-15 = gate 2p + 1 of this thread's unit ----
  uint32_t whi[2][8][4], wlo[2][8][4];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const float* r0p = W + (long)((2 * p) * LH + unit) * LH + 32 * tig;
    const float* r1p = r0p + (long)LH * LH;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const float4 r0 = __ldg(reinterpret_cast<const float4*>(r0p) + ks), r1 = __ldg(reinterpret_cast<const float4*>(r1p) + ks);
      split_pair(r0.x, r0.y, whi[p][ks][0], wlo[p][ks][0]);
      split_pair(r1.x, r1.y, whi[p][ks][1], wlo[p][ks][1]);
      split_pair(r0.z, r0.w, whi[p][ks][2], wlo[p][ks][2]);
      split_pair(r1.z, r1.w, whi[p][ks][3], wlo[p][ks][3]);
    }
  }
  for (int i = tid; i < 2 * 2 * 128; i += 256) hbuf[i] = make_uint4(0u, 0u, 0u, 0u);
  // h of (unit u, sequence n) lives at bf16 index (((u % 32) / 8 * 8 + n) * 4 + u / 32) * 8 + u % 8 of a plane.  Lane pairs
  // (g even / odd = adjacent units) swap one sequence each, so a thread writes ONE 32-bit word (units u & ~1, u | 1) per plane
  // of sequence 2tig + (g & 1) -- st.async moves 32-bit words
  const int hword = ((((warp & 3) * 8 + 2 * tig + (g & 1)) * 4 + 2 * (int)rank + (warp >> 2)) * 8 + (g & 6)) >> 1;   // 32-bit word index in a plane
  const uint32_t hpeer = map_peer(hbuf, rank ^ 1u) + 4u * (uint32_t)hword;
  const uint32_t bpeer = map_peer(hbar, rank ^ 1u);
  const bool live0 = b0 + 2 * tig < B, live1 = b0 + 2 * tig + 1 < B;
  const int t0 = dir ? T - 1 : 0;
  const long dstep = dir ? -1 : 1;
  float* gq0 = G + ((long)(live0 ? b0 + 2 * tig : b0) * T + t0) * LGS + dir * LG + unit;
  float* gq1 = G + ((long)(live1 ? b0 + 2 * tig + 1 : b0) * T + t0) * LGS + dir * LG + unit;
  long so0 = ((long)(live0 ? b0 + 2 * tig : b0) * T + t0) * 256 + dir * LH + unit;
  long so1 = ((long)(live1 ? b0 + 2 * tig + 1 : b0) * T + t0) * 256 + dir * LH + unit;
  float c[2] = {0.f, 0.f}, hprev[2] = {0.f, 0.f};
  // x-projection (+ biases) in accumulator order: [tile p][j] = (gate 2p + (j >> 1), sequence 2 tig + (j & 1))
  // Loaded TWO steps ahead into two register sets (even / odd steps): a step is ~1.3 us, a 32-byte global load under this
  // access pattern takes about as long, and with a one-step look-ahead every step waited for its x-projection.
  float gxa[2][4], gxb[2][4];
  auto fetch = [&](float (&dst)[2][4], const float* p0, const float* p1) {
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[p][j] = ((j & 1) ? p1 : p0)[(2 * p + (j >> 1)) * LH];
  };
  fetch(gxa, gq0, gq1);
  {
    const long nd = T > 1 ? dstep * LGS : 0;
    fetch(gxb, gq0 + nd, gq1 + nd);
  }
  cluster_sync_all();                                      // both CTAs' planes are zeroed before anyone writes remotely
  auto do_step = [&](const int step, float (&ngx)[2][4]) {
    stamp(0);
    if (tid == 0) tc::mbar_expect_tx(&hbar[(step + 1) & 1], 2048);   // this step's h of the peer: 256 threads x 2 words
    const uint4* hh = hbuf + (step & 1) * 256;
    const uint4* hl = hh + 128;
    float acc[2][2][4];                                    // [tile][hi.hi (+ x-projection), hi.lo + lo.hi]
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[p][0][j] = ngx[p][j]; acc[p][1][j] = 0.f; }
    {
      const long nd = step + 2 < T ? 2 * dstep * LGS : 0;  // past the end: re-read a valid row (unused)
      fetch(ngx, gq0 + nd, gq1 + nd);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 vh = hh[(i * 8 + g) * 4 + tig], vl = hl[(i * 8 + g) * 4 + tig];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int ks = 2 * i + kk;
        const uint32_t bh0 = kk ? vh.z : vh.x, bh1 = kk ? vh.w : vh.y, bl0 = kk ? vl.z : vl.x, bl1 = kk ? vl.w : vl.y;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          mma_bf16(acc[p][0], whi[p][ks], bh0, bh1);
          mma_bf16(acc[p][1], whi[p][ks], bl0, bl1);
          mma_bf16(acc[p][1], wlo[p][ks], bh0, bh1);
        }
      }
    }
    if (dbgp) { if (acc[0][0][0] + acc[1][1][3] == 123.456f) dbgp[254] = 1; stamp(2); }
    float gi[2], gf[2], gg[2], go[2], cn[2], hn[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      gi[e] = sigmoid_fast(acc[0][0][e] + acc[0][1][e]);
      gf[e] = sigmoid_fast(acc[0][0][2 + e] + acc[0][1][2 + e]);
      gg[e] = tanh_fast(acc[1][0][e] + acc[1][1][e]);
      go[e] = sigmoid_fast(acc[1][0][2 + e] + acc[1][1][2 + e]);
      cn[e] = fmaf(gf[e], c[e], gi[e] * gg[e]);
      c[e] = cn[e];
      hn[e] = go[e] * tanh_fast(cn[e]);
    }
    {
      uint32_t hi, lo;
      split_pair(hn[0], hn[1], hi, lo);
      const uint32_t phi = __shfl_xor_sync(0xffffffffu, hi, 4), plo = __shfl_xor_sync(0xffffffffu, lo, 4);
      // even g: sequence 2tig of units (u, u+1) = (own low half, partner's low half); odd g: sequence 2tig+1 of (u-1, u)
      const uint32_t whi = (g & 1) ? __byte_perm(phi, hi, 0x7632) : __byte_perm(hi, phi, 0x5410);
      const uint32_t wlo = (g & 1) ? __byte_perm(plo, lo, 0x7632) : __byte_perm(lo, plo, 0x5410);
      const int nbuf = (step + 1) & 1;
      uint32_t* nh = reinterpret_cast<uint32_t*>(hbuf + nbuf * 256) + hword;
      nh[0] = whi;
      nh[128 * 4] = wlo;
      const uint32_t pa = hpeer + (uint32_t)(nbuf * 256 * 16), pb = bpeer + 8u * (uint32_t)nbuf;
      st_async_b32(pa, whi, pb);
      st_async_b32(pa + 2048, wlo, pb);
    }
    stamp(3);
    // (history: with st.shared::cluster + barrier.cluster the arrive's release waited for every global load / store of the
    //  thread in flight: 0.45 us of a 1.4 us step even with the stores moved behind the arrive)
    if (live0) {
      gq0[0] = gi[0]; gq0[LH] = gf[0]; gq0[2 * LH] = gg[0]; gq0[3 * LH] = go[0];
      out[so0] = hn[0]; Cs[so0] = cn[0]; Hp[so0] = hprev[0];
    }
    if (live1) {
      gq1[0] = gi[1]; gq1[LH] = gf[1]; gq1[2 * LH] = gg[1]; gq1[3 * LH] = go[1];
      out[so1] = hn[1]; Cs[so1] = cn[1]; Hp[so1] = hprev[1];
    }
    hprev[0] = hn[0]; hprev[1] = hn[1];
    gq0 += dstep * LGS; gq1 += dstep * LGS; so0 += dstep * 256; so1 += dstep * 256;
    stamp(4);
    __syncthreads();                                       // this CTA's half of h_t is in its planes
    tc::mbar_wait(&hbar[(step + 1) & 1], (step >> 1) & 1);   // ... and the peer's half (each barrier completes every other step)
  };
  for (int step = 0; step < T; step += 2) {
    do_step(step, gxa);
    if (step + 1 < T) do_step(step + 1, gxb);
  }
  cluster_sync_all();                                      // neither CTA leaves while the other may still address it
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
    lstm128_bwd_c2_kernel(const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_r, float* __restrict__ G, const float* __restrict__ dout,
                          const float* __restrict__ Cs, int B, int T) {
  pdl_sync();
  __shared__ __align__(16) uint4 dbuf[2 * 256];            // da planes of this CTA's 256 gate rows: [hi, lo][256 uint4]
  __shared__ __align__(16) float dhx[2 * 2 * 64 * 8];      // partial dh_{t-1} of this CTA's units: [buffer][own, peer's][unit 64][sequence 8]
  __shared__ __align__(8) uint64_t dbar[2];                // per buffer: the peer's 2 KB of partial sums have landed
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) { tc::mbar_init(&dbar[0], 1); tc::mbar_init(&dbar[1], 1); tc::mbar_fence_init(); }
  const int dir = blockIdx.y, b0 = (blockIdx.x >> 1) * LMS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const float* W = dir ? w_hh_r : w_hh_f;
  // ---- A fragments of W^T over this CTA's gate rows: m-tile = warp (units 16w+g: a0, a2; 16w+8+g: a1, a3);
  //      k-step s column 2tig+e [+8] <-> gate row 128 tig + 64 rank + 4s + e [+2] ----
  uint32_t whi[16][4], wlo[16][4];
  {
    const float* wc = W + (long)(128 * tig + 64 * (int)rank) * LH + 16 * warp + g;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      const float* p = wc + (long)(4 * ks) * LH;
      split_pair(__ldg(p), __ldg(p + LH), whi[ks][0], wlo[ks][0]);
      split_pair(__ldg(p + 8), __ldg(p + LH + 8), whi[ks][1], wlo[ks][1]);
      split_pair(__ldg(p + 2 * LH), __ldg(p + 3 * LH), whi[ks][2], wlo[ks][2]);
      split_pair(__ldg(p + 2 * LH + 8), __ldg(p + 3 * LH + 8), whi[ks][3], wlo[ks][3]);
    }
  }
  for (int i = tid; i < 2 * 2 * 64 * 8; i += 256) dhx[i] = 0.f;
  // gate role: local unit lu = 8w + g (global unit 64 rank + lu), sequences 2tig + e.  da of (gate gt, lu, n) lives at bf16
  // index ((lu / 8 * 8 + n) * 4 + gt) * 8 + lu % 8 of a plane
  const int lu = 8 * warp + g, unit = 64 * (int)rank + lu;
  const int doff = ((warp * 8 + 2 * tig) * 4) * 8 + g;                               // + e * 32 + gt * 8 (bf16 elements)
  // partial-sum role: this warp's m-tile belongs to CTA (warp >> 2): rows 16 (w & 3) + g and + 8 of that CTA's dhx, float2 per row
  const bool mine = (uint32_t)(warp >> 2) == rank;
  const int poff = (mine ? 0 : 64 * 8) + (16 * (warp & 3) + g) * 8 + 2 * tig;          // + 64 for the second row; [own, peer's] halves
  const uint32_t ppeer = map_peer(dhx, rank ^ 1u) + 4u * (uint32_t)poff;
  const uint32_t bpeer = map_peer(dbar, rank ^ 1u);
  const bool live0 = b0 + 2 * tig < B, live1 = b0 + 2 * tig + 1 < B;
  const int t0 = dir ? 0 : T - 1;                                                   // reverse of the forward order
  const long dstep = dir ? 1 : -1;
  float* gq[2] = {G + ((long)(live0 ? b0 + 2 * tig : b0) * T + t0) * LGS + dir * LG + unit,
                  G + ((long)(live1 ? b0 + 2 * tig + 1 : b0) * T + t0) * LGS + dir * LG + unit};
  long so[2] = {((long)(live0 ? b0 + 2 * tig : b0) * T + t0) * 256 + dir * LH + unit,
                ((long)(live1 ? b0 + 2 * tig + 1 : b0) * T + t0) * 256 + dir * LH + unit};
  // saved state of a step: [e][gate 0..3, dy, c_{t-1}], loaded two steps ahead into two register sets (see the forward kernel);
  // `ahead` = steps past the cursors, `has_prev`: a forward-earlier step exists for that step
  float dc[2] = {0.f, 0.f}, ct[2], sva[2][6], svb[2][6];
  auto fetch = [&](float (&dst)[2][6], long ahead, bool has_prev) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
#pragma unroll
      for (int gt = 0; gt < 4; ++gt) dst[e][gt] = gq[e][ahead * dstep * LGS + gt * LH];
      dst[e][4] = dout[so[e] + ahead * dstep * 256];
      dst[e][5] = has_prev ? Cs[so[e] + (ahead + 1) * dstep * 256] : 0.f;
    }
  };
  ct[0] = Cs[so[0]]; ct[1] = Cs[so[1]];
  fetch(sva, 0, T > 1);
  fetch(svb, T > 1 ? 1 : 0, T > 2);
  cluster_sync_all();
  auto do_step = [&](const int step, float (&sv)[2][6]) {
    if (tid == 0) tc::mbar_expect_tx(&dbar[(step + 1) & 1], 2048);   // the peer's four non-owner warps x 32 lanes x 2 x 8 bytes
    // ---- gate phase: da of this thread's 2 pairs ----
    const float* dx = dhx + (step & 1) * 2 * 64 * 8 + lu * 8 + 2 * tig;
    const float2 d_own = *reinterpret_cast<const float2*>(dx), d_peer = *reinterpret_cast<const float2*>(dx + 64 * 8);
    const float dhv[2] = {d_own.x + d_peer.x, d_own.y + d_peer.y};
    float da[2][4];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float gi = sv[e][0], gf = sv[e][1], gg = sv[e][2], go = sv[e][3], cp = sv[e][5];
      const float dht = sv[e][4] + dhv[e];
      const float tc = tanh_fast(ct[e]);
      const float dct = fmaf(dht * go, fmaf(-tc, tc, 1.f), dc[e]);
      da[e][0] = dct * gg * gi * (1.f - gi);
      da[e][1] = dct * cp * gf * (1.f - gf);
      da[e][2] = dct * gi * fmaf(-gg, gg, 1.f);
      da[e][3] = dht * tc * go * (1.f - go);
      dc[e] = dct * gf;
      ct[e] = cp;                                          // c_{t-1} is the next step's c_t
    }
    if (live0) { gq[0][0] = da[0][0]; gq[0][LH] = da[0][1]; gq[0][2 * LH] = da[0][2]; gq[0][3 * LH] = da[0][3]; }
    if (live1) { gq[1][0] = da[1][0]; gq[1][LH] = da[1][1]; gq[1][2 * LH] = da[1][2]; gq[1][3 * LH] = da[1][3]; }
    fetch(sv, step + 2 < T ? 2 : 0, step + 3 < T);         // past the end: re-read a valid row (unused)
    gq[0] += dstep * LGS; gq[1] += dstep * LGS; so[0] += dstep * 256; so[1] += dstep * 256;
    {
      uint16_t* dw = reinterpret_cast<uint16_t*>(dbuf) + doff;
#pragma unroll
      for (int gt = 0; gt < 4; ++gt) {
        uint32_t hi, lo;
        split_pair(da[0][gt], da[1][gt], hi, lo);
        dw[gt * 8] = (uint16_t)(hi & 0xffffu); dw[gt * 8 + 32] = (uint16_t)(hi >> 16);
        dw[256 * 8 + gt * 8] = (uint16_t)(lo & 0xffffu); dw[256 * 8 + gt * 8 + 32] = (uint16_t)(lo >> 16);
      }
    }
    __syncthreads();
    // ---- partial dh_{t-1}^T (this warp's 16 units x 8 sequences) over this CTA's 256 gate rows ----
    float acc[4][4];                                       // hi.hi and (hi.lo + lo.hi), each over even / odd k-steps
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][j] = 0.f;
#pragma unroll
    for (int cch = 0; cch < 8; ++cch) {
      const uint4 vh = dbuf[(cch * 8 + g) * 4 + tig], vl = dbuf[256 + (cch * 8 + g) * 4 + tig];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int ks = 2 * cch + kk;
        const uint32_t bh0 = kk ? vh.z : vh.x, bh1 = kk ? vh.w : vh.y, bl0 = kk ? vl.z : vl.x, bl1 = kk ? vl.w : vl.y;
        mma_bf16(acc[kk], whi[ks], bh0, bh1);
        mma_bf16(acc[2 + kk], whi[ks], bl0, bl1);
        mma_bf16(acc[2 + kk], wlo[ks], bh0, bh1);
      }
    }
    // accumulator j: unit 16w + 8 (j >> 1) + g, sequence 2 tig + (j & 1)  ->  the owner's dhx of the NEXT step
    {
      const float p0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]), p1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
      const float p2 = (acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2]), p3 = (acc[0][3] + acc[1][3]) + (acc[2][3] + acc[3][3]);
      const int nb = ((step + 1) & 1) * 2 * 64 * 8;
      if (mine) {
        *reinterpret_cast<float2*>(dhx + nb + poff) = make_float2(p0, p1);
        *reinterpret_cast<float2*>(dhx + nb + poff + 64) = make_float2(p2, p3);
      } else {
        const uint32_t pb = bpeer + 8u * (uint32_t)((step + 1) & 1);
        st_async_f32x2(ppeer + 4u * (uint32_t)nb, p0, p1, pb);
        st_async_f32x2(ppeer + 4u * (uint32_t)(nb + 64), p2, p3, pb);
      }
    }
    __syncthreads();                                       // own partial sums written; the da planes may be rewritten
    tc::mbar_wait(&dbar[(step + 1) & 1], (step >> 1) & 1);   // the peer's partial sums have landed
  };
  for (int step = 0; step < T; step += 2) {
    do_step(step, sva);
    if (step + 1 < T) do_step(step + 1, svb);
  }
  cluster_sync_all();                                      // neither CTA leaves while the other may still address it
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
    const float act = (j >> 2) == 2 ? tanh_fast(a) : sigmoid_fast(a);
    // unit lane u (= j & 3) gathers i, f, g, o of its unit
    const int u = j & 3;
    const float gi = __shfl_sync(gmask, act, lbase + u), gf = __shfl_sync(gmask, act, lbase + 4 + u);
    const float gg = __shfl_sync(gmask, act, lbase + 8 + u), go = __shfl_sync(gmask, act, lbase + 12 + u);
    c = gf * c + gi * gg;                               // identical in the 4 lanes sharing a unit; lanes j < 4 are authoritative
    const float hn = go * tanh_fast(c);
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
                                                       float* __restrict__ dwhh_f, float* __restrict__ dwhh_r, int B, int T,
                                                       float* __restrict__ part /* [gridDim.x][2][64] or nullptr (atomics) */) {
  pdl_sync();
  __shared__ float red[2][64];
  __shared__ __align__(16) float slot[8][64];
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
    const float tc = tanh_fast(ct);
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
  if (part) {
    // fixed order: the 8 sequences of the CTA in sequence order, the CTAs in CTA order (lstm4_dw_fold_kernel)
    *reinterpret_cast<float4*>(&slot[threadIdx.x >> 4][j * 4]) = live ? make_float4(dw0, dw1, dw2, dw3) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int e = threadIdx.x & 63, dsel = threadIdx.x >> 6;
    float acc = 0.f;
    for (int sgrp = 0; sgrp < 8; ++sgrp) {
      const int gid_ = blockIdx.x * 8 + sgrp;
      if (gid_ < 2 * B && gid_ / B == dsel) acc += slot[sgrp][e];
    }
    part[(long)blockIdx.x * 128 + threadIdx.x] = acc;
    return;
  }
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

// dW_hh (both directions, 64 floats each) = sum of the per-CTA partials in CTA order
__global__ void __launch_bounds__(128) lstm4_dw_fold_kernel(const float* __restrict__ part, int ctas, float* __restrict__ dwhh_f,
                                                            float* __restrict__ dwhh_r) {
  pdl_sync();
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int c = 0;
  for (; c + 4 <= ctas; c += 4) {
    a0 += part[(long)(c + 0) * 128 + threadIdx.x]; a1 += part[(long)(c + 1) * 128 + threadIdx.x];
    a2 += part[(long)(c + 2) * 128 + threadIdx.x]; a3 += part[(long)(c + 3) * 128 + threadIdx.x];
  }
  for (; c < ctas; ++c) a0 += part[(long)c * 128 + threadIdx.x];
  const float v = (a0 + a1) + (a2 + a3);
  if (threadIdx.x < 64) dwhh_f[threadIdx.x] = v; else dwhh_r[threadIdx.x - 64] = v;
}

}  // namespace lstm
}  // namespace eegclip
