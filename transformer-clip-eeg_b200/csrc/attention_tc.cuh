// Tensor-core short-sequence attention (clip_model.py:30-45): 8 heads x head_dim 8, T <= 512, softmax(QK^T / sqrt(64)),
// dropout on the probabilities, no mask.  The (B,8,T,T) energy / probability tensors never exist.
//
// Warp-level mma.sync on fp16 operands with fp32 accumulation.  head_dim = 8 is the K extent of m16n8k8: one MMA produces a
// 16 x 8 score tile, and its accumulator fragment is re-used in place as the A fragment of the next MMA (P.V, P^T.dO, dS^T.Q:
// one cvt.rn.f16x2 per pair of scores) if the contracted indices are taken in the order the fragment already has them -- so
// probabilities never move between threads.  tcgen05 is not used here on purpose: the contraction dimension is 8,
// the work per score is exp2 + dropout + conversion (measured: the forward sits at 57 % of the MUFU pipe, 54 % of the ALU pipe,
// 68 % of the issue slots and only 35 % of the legacy tensor pipe), and the warp-level MMA keeps everything in registers.
// History (B = 256, T = 320, per layer): TF32 m16n8k8 with dS staged through shared memory for dQ: forward 118 us, backward
// 288 us; fp16 with m16n8k16 over queries / keys and movmatrix for dQ (this file): 96 / 174 us.  Every mma.sync shape costs the
// same 8 clk per SM sub-partition (tools/micro/hmma_probe.cu), so k = 16 forms do twice the work per tensor-pipe slot.
//
// Dropout grouping: one Philox call yields 8 decisions for 8 consecutive key indices of one query row (common.cuh), so
// the key <-> fragment-column assignment is permuted such that every thread owns 8 consecutive keys of its rows:
//   forward  (rows = queries): a 32-key block = 4 n-tiles X; tile X column c  <->  key 8*(c>>1) + 2X + (c&1)
//   backward (rows = keys)   : a warp owns 64 keys = 4 m-tiles Y; tile Y row g <-> key 8g+2Y, row g+8 <-> key 8g+2Y+1
// The B/A fragments of K and V are laid out in shared memory / registers in exactly that order.
//
// Numerics: Q, K, V, P, dO, dS are rounded to fp16 (11-bit significand, the precision of the TF32 form this replaced; dO is
// brought into fp16's range by a power-of-two scale per (batch, head)), accumulation and softmax statistics are fp32.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "attention.cuh"

namespace eegclip {
namespace attntc {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Dropout decisions of one (batch, head): the T x T keep-bit matrix in shared memory (1-bit mode, p == 0.5: T*T/8 bytes,
// 12.8 KB at T = 320), produced by T*T/128 Philox calls spread over the CTA -- the per-score cost of dropout becomes one
// shared-memory byte read instead of a tenth of a Philox call.  Word w of the matrix covers keys 32*(w % (T/32)).. of row w / (T/32).
__device__ __forceinline__ void build_keep_bits(uint32_t* smask, const Drop& drop, int bh, int T) {
  const int nblocks = (T * T) >> 7;
  const uint64_t block0 = ((uint64_t)bh * (uint64_t)T * (uint64_t)T) >> 7;
  for (int k = threadIdx.x; k < nblocks; k += blockDim.x) {
    const uint4 w = drop_words(drop, block0 + k);
    *reinterpret_cast<uint4*>(smask + 4 * k) = w;
  }
}

// ------------------------------------------------------------------------------------------------
// Backward on fp16 operands (same 11-bit significand as TF32; fp32 accumulation and softmax arithmetic unchanged).
// Every mma.sync costs the same 8 clk per SM sub-partition whatever its shape (tools/micro/hmma_probe.cu), so the products
// whose contraction runs over queries or keys use m16n8k16 (two query blocks per instruction), and 16-bit fragments make three
// things cheap that TF32 did not: accumulator -> A fragment is ONE cvt.rn.f16x2 per pair (TF32: add + and per element), the
// B fragments are single 32-bit shared loads, and dS^T reaches the dQ product through movmatrix (register 8x8 transpose)
// instead of a shared-memory staging tile (16 LDS + 8 STS + 8 half-filled MMAs per 8 queries before).
//   per 16 queries and 16-key tile: S^T - lse: 2 x m16n8k16 (lse in the spare contraction slots), dP^T: 2 x m16n8k8;
//   dV += Pd^T.dO, dK += dS^T.Q: 2 x m16n8k16;  dQ += dS.K': 1 x m16n8k16 (16 queries x 16 keys, no padding) + 4 movmatrix.
// Range: everything downstream of dO is linear in dO, so dO is scaled by a power of two per (batch, head) to [8, 16) max
// magnitude before it is rounded to fp16 and the three gradients are scaled back (gradients of 1e-6 would be fp16 subnormals).
// smem: Qh, dOh [T][8] and their transposes QT, dOT [8][T+8] (fp16), LB [T][4] (-lse' terms), Ds [T], per-warp dQ slots, keep bits.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void mma_h8(float* d, uint32_t a0, uint32_t a1, uint32_t b0, const float* c) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%8,%9,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
__device__ __forceinline__ void mma_h16(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// x as three fp16 terms (hi, mid | lo, 0): the row maximum / log-sum-exp rides in contraction slots 8..10 of a k = 16 score MMA
// against ones (every mma.sync costs the same 8 clk whatever its k), so S - max is produced by the MMA with a ZERO accumulator
// input -- as the accumulator input it cost four register moves per MMA (the C operand is a register quad)
__device__ __forceinline__ void split3_h(float x, uint32_t& w0, uint32_t& w1) {
  const __half h1 = __float2half_rn(x);
  const float r1 = x - __half2float(h1);
  const __half h2 = __float2half_rn(r1);
  const __half h3 = __float2half_rn(r1 - __half2float(h2));
  w0 = (uint32_t)__half_as_ushort(h1) | ((uint32_t)__half_as_ushort(h2) << 16);
  w1 = (uint32_t)__half_as_ushort(h3);
}
__device__ __forceinline__ void mma_h16z(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ uint32_t movm_t(uint32_t x) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}

// Forward on fp16 operands: S = Q.K^T as m16n8k8 (the contraction is the head dimension, 8), P.V as m16n8k16 -- the score
// accumulators of two 8-key tiles become one A fragment with four cvt.rn.f16x2 (the flash-attention register reuse), so the P.V
// product takes half the instructions of the TF32 form and its operand tables half the shared memory.  Two passes (row
// maximum first): fp16 probabilities need the true maximum -- an upper bound (|q| max|k|) would push them below 2^-14.
//   grid = B*H, block = (T/32) warps; warp w owns query rows [32w, 32w+32) as two 16-row tiles.
//   smem: Kf, Vf: uint4 [T/32 blocks][32 lanes] (K: tile X column g <-> key 8(g>>1)+2X+(g&1), dims 2tig, 2tig+1; V: keys
//   8tig..8tig+7 of the block for dim g), + keep bits.
__global__ void __launch_bounds__(512, 2) attn_fwd_h_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                           float* __restrict__ lse, int T, Drop drop) {
  pdl_sync();
  extern __shared__ uint4 smq[];
  uint4* Kf = smq;
  uint4* Vf = smq + T;
  uint32_t* smask = reinterpret_cast<uint32_t*>(Vf + T);
  const int bh = blockIdx.x, b = bh / AH, h = bh % AH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const float* base = qkv + (long)b * T * AQKV;
  const int nblk = T >> 5;
  for (int e = tid; e < 2 * T; e += blockDim.x) {
    const int isv = e >= T;
    const int r = isv ? e - T : e;
    const int blk = r >> 5, l = r & 31, eg = l >> 2, et = l & 3;
    uint4 v;
    if (!isv) {
      uint32_t w[4];
#pragma unroll
      for (int X = 0; X < 4; ++X) {
        const float2 kk = __ldg(reinterpret_cast<const float2*>(base + (long)(blk * 32 + 8 * (eg >> 1) + 2 * X + (eg & 1)) * AQKV + 64 + h * AD + 2 * et));
        w[X] = pack_h2(kk.x, kk.y);
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __ldg(base + (long)(blk * 32 + 8 * et + j) * AQKV + 128 + h * AD + eg);
      v = make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
    }
    (isv ? Vf : Kf)[blk * 32 + l] = v;
  }
  const bool bitmask = drop.enabled && drop.onebit;
  if (bitmask) build_keep_bits(smask, drop, bh, T);
  // ---- Q fragments of this warp (two row tiles), scaled by log2(e)/sqrt(64) ----
  const int i0 = warp * 32;
  const float qs = LOG2E * 0.125f;
  uint32_t qa[2][2];
#pragma unroll
  for (int R = 0; R < 2; ++R) {
    const float* q0p = base + (long)(i0 + 16 * R + g) * AQKV + h * AD + 2 * tig;
    const float2 a = __ldg(reinterpret_cast<const float2*>(q0p)), c = __ldg(reinterpret_cast<const float2*>(q0p + 8 * AQKV));
    qa[R][0] = pack_h2(a.x * qs, a.y * qs);
    qa[R][1] = pack_h2(c.x * qs, c.y * qs);
  }
  __syncthreads();
  const float zero[4] = {0.f, 0.f, 0.f, 0.f};
  // ---- pass 1: row maxima ----
  float mx[2][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}};
  for (int blk = 0; blk < nblk; ++blk) {
    const uint4 k4 = Kf[blk * 32 + lane];
    const uint32_t kb[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
    for (int X = 0; X < 4; ++X)
#pragma unroll
      for (int R = 0; R < 2; ++R) {
        float s[4];
        mma_h8(s, qa[R][0], qa[R][1], kb[X], zero);
        mx[R][0] = fmaxf(mx[R][0], fmaxf(s[0], s[1]));
        mx[R][1] = fmaxf(mx[R][1], fmaxf(s[2], s[3]));
      }
  }
#pragma unroll
  for (int R = 0; R < 2; ++R)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float m = mx[R][hh];
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      mx[R][hh] = m;
    }
  // ---- pass 2: probabilities, dropout, P.V ----
  // (measured and dropped here: the maximum in the spare contraction slots of a k = 16 score MMA, as the backward does with the
  //  log-sum-exp, plus packed f32x2 adds for the row sums: 99 against 96 us per launch -- this kernel is bound by the MUFU
  //  and ALU pipes at 55-57 % each, not by the register moves in front of the MMA)
  float l[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int blk = 0; blk < nblk; ++blk) {
    const uint4 k4 = Kf[blk * 32 + lane], v4 = Vf[blk * 32 + lane];
    const uint32_t kb[4] = {k4.x, k4.y, k4.z, k4.w}, vb[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int R = 0; R < 2; ++R) {
      const int r0 = i0 + 16 * R + g;
      uint32_t keep0, keep1;   // decisions of keys blk*32 + 8tig .. +7 for rows r0 and r0 + 8
      if (bitmask) {
        keep0 = (smask[r0 * nblk + blk] >> (8 * tig)) & 0xffu;
        keep1 = (smask[(r0 + 8) * nblk + blk] >> (8 * tig)) & 0xffu;
      } else {
        const uint64_t e0 = ((uint64_t)bh * T + (uint64_t)r0) * (uint64_t)T + (uint64_t)(blk * 32 + 8 * tig);
        keep0 = drop_bits8(drop, e0);
        keep1 = drop_bits8(drop, e0 + (uint64_t)8 * T);
      }
      // the row maximum enters as the accumulator input of the score MMA: S - max costs no arithmetic instruction
      const float negmx[4] = {-mx[R][0], -mx[R][0], -mx[R][1], -mx[R][1]};
#pragma unroll
      for (int XP = 0; XP < 2; ++XP) {
        uint32_t pa[4];
#pragma unroll
        for (int xx = 0; xx < 2; ++xx) {
          const int X = 2 * XP + xx;
          float s[4];
          mma_h8(s, qa[R][0], qa[R][1], kb[X], negmx);
          const float p0 = ex2(s[0]), p1 = ex2(s[1]), p2 = ex2(s[2]), p3 = ex2(s[3]);
          l[R][0] += p0 + p1; l[R][1] += p2 + p3;
          // this thread's keys of the block: 8tig + 2X (+1)
          pa[2 * xx] = pack_h2((keep0 >> (2 * X)) & 1u ? p0 : 0.f, (keep0 >> (2 * X + 1)) & 1u ? p1 : 0.f);
          pa[2 * xx + 1] = pack_h2((keep1 >> (2 * X)) & 1u ? p2 : 0.f, (keep1 >> (2 * X + 1)) & 1u ? p3 : 0.f);
        }
        mma_h16(o[R], pa[0], pa[1], pa[2], pa[3], vb[2 * XP], vb[2 * XP + 1]);
      }
    }
  }
#pragma unroll
  for (int R = 0; R < 2; ++R) {
    float l0 = l[R][0], l1 = l[R][1];
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float s0 = drop.scale / l0, s1 = drop.scale / l1;
    const int r0 = i0 + 16 * R + g;
    float* o0 = out + ((long)b * T + r0) * AE + h * AD + 2 * tig;
    *reinterpret_cast<float2*>(o0) = make_float2(o[R][0] * s0, o[R][1] * s0);
    *reinterpret_cast<float2*>(o0 + 8 * AE) = make_float2(o[R][2] * s1, o[R][3] * s1);
    if (tig == 0) {
      lse[(long)bh * T + r0] = (mx[R][0] + log2f(l0)) * LN2;
      lse[(long)bh * T + r0 + 8] = (mx[R][1] + log2f(l1)) * LN2;
    }
  }
}

__global__ void __maxnreg__(112) attn_bwd_h_kernel(const float* __restrict__ qkv, const float* __restrict__ out,
                                                           const float* __restrict__ dout, const float* __restrict__ lse,
                                                           float* __restrict__ dqkv, int T, Drop drop) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t smh[];
  const int TP = T + 8;
  __half* Qh = reinterpret_cast<__half*>(smh);
  __half* dOh = Qh + T * 8;
  __half* QT = dOh + T * 8;
  __half* dOT = QT + 8 * TP;
  uint32_t* LB = reinterpret_cast<uint32_t*>(dOT + 8 * TP);   // [T][4]: -lse' as three fp16 terms (words 0, 1), zeros (2, 3)
  float* Ds = reinterpret_cast<float*>(LB + 4 * T);
  float* dqp = Ds + T;                                     // [2][warps][16 x 8]: per-warp dQ partials of a 16-query block
  float* red = dqp + 2 * (T >> 6) * 128;
  uint32_t* smask = reinterpret_cast<uint32_t*>(red + 8);
  const bool bitmask = drop.enabled && drop.onebit;
  const int nblk = T >> 5;
  const int bh = blockIdx.x, b = bh / AH, h = bh % AH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const int nthr = blockDim.x;                            // T / 2: every thread stages rows tid and tid + T/2
  const float* base = qkv + (long)b * T * AQKV;
  // ---- stage Q (fp16, both layouts), lse', zero dQ; keep dO in registers until its scale is known ----
  float dOr[2][8], dsum[2];
  float amax = 0.f;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = tid + r * nthr;
    float q[8], o[8];
    load8(base + (long)i * AQKV + h * AD, q);
    load8(dout + ((long)b * T + i) * AE + h * AD, dOr[r]);
    load8(out + ((long)b * T + i) * AE + h * AD, o);
    float ds_ = 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      ds_ = fmaf(dOr[r][d], o[d], ds_);
      amax = fmaxf(amax, fabsf(dOr[r][d]));
      const __half qh = __float2half_rn(q[d]);
      Qh[i * 8 + d] = qh;
      QT[d * TP + i] = qh;
    }
    dsum[r] = ds_;
    uint32_t w0, w1;
    split3_h(-lse[(long)bh * T + i] * LOG2E, w0, w1);
    *reinterpret_cast<uint4*>(LB + 4 * i) = make_uint4(w0, w1, 0u, 0u);
  }
#pragma unroll
  for (int o_ = 16; o_ > 0; o_ >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o_));
  if (lane == 0) red[warp] = amax;
  if (bitmask) build_keep_bits(smask, drop, bh, T);
  __syncthreads();
  amax = red[0];
  for (int w_ = 1; w_ < (nthr >> 5); ++w_) amax = fmaxf(amax, red[w_]);
  // power-of-two scale that brings max |dO| into [8, 16)
  float scale = 1.f, inv = 1.f;
  if (amax > 1e-30f) {
    const int e_ = (int)((__float_as_uint(amax) >> 23) & 0xffu) - 127;
    scale = __uint_as_float((uint32_t)(127 + 3 - e_) << 23);
    inv = __uint_as_float((uint32_t)(127 - 3 + e_) << 23);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = tid + r * nthr;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const __half v = __float2half_rn(dOr[r][d] * scale);
      dOh[i * 8 + d] = v;
      dOT[d * TP + i] = v;
    }
    Ds[i] = dsum[r] * scale;
  }
  // ---- A fragments of this warp's 64 keys.  Tile Y: row g <-> key 8g+2Y, row g+8 <-> key 8g+2Y+1 ----
  const int j0 = warp * 64;
  const float ks = LOG2E * 0.125f;
  uint32_t ka[4][2], va[4][2], kt[4][2];
#pragma unroll
  for (int Y = 0; Y < 4; ++Y) {
    const float* k0 = base + (long)(j0 + 8 * g + 2 * Y) * AQKV + 64 + h * AD + 2 * tig;
    const float* k1 = k0 + AQKV;
    const float2 a = __ldg(reinterpret_cast<const float2*>(k0)), c = __ldg(reinterpret_cast<const float2*>(k1));
    ka[Y][0] = pack_h2(a.x * ks, a.y * ks); ka[Y][1] = pack_h2(c.x * ks, c.y * ks);
    const float2 va_ = __ldg(reinterpret_cast<const float2*>(k0 + 64)), vc_ = __ldg(reinterpret_cast<const float2*>(k1 + 64));
    va[Y][0] = pack_h2(va_.x, va_.y); va[Y][1] = pack_h2(vc_.x, vc_.y);
    // dQ = dS . K' : B fragment (n = d = g), contraction index = tile row: 2tig, 2tig+1 <-> keys 16tig+2Y, 16tig+8+2Y; +8 <-> the same + 1
    const float* kr = base + (long)(j0 + 16 * tig + 2 * Y) * AQKV + 64 + h * AD + g;
    kt[Y][0] = pack_h2(__ldg(kr) * ks, __ldg(kr + 8 * AQKV) * ks);
    kt[Y][1] = pack_h2(__ldg(kr + AQKV) * ks, __ldg(kr + 9 * AQKV) * ks);
  }
  float dk[4][4], dv[4][4];
#pragma unroll
  for (int Y = 0; Y < 4; ++Y)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dk[Y][e] = 0.f; dv[Y][e] = 0.f; }
  __syncthreads();
  const float zero[4] = {0.f, 0.f, 0.f, 0.f};
  const float dscale = drop.scale;
  const float sq = LN2 * inv;
  const int nwarp = nthr >> 5;
  const uint32_t kone = tig == 0 ? 0x3c003c00u : tig == 1 ? 0x00003c00u : 0u;   // ones in contraction slots 8, 9, 10 (see split3_h)
  const uint32_t* Qw = reinterpret_cast<const uint32_t*>(Qh);
  const uint32_t* Ow = reinterpret_cast<const uint32_t*>(dOh);
  for (int q0 = 0; q0 < T; q0 += 16) {
    // B fragments of the two query blocks X = 0, 1 (queries q0 + 8X + ...)
    uint32_t qb[2], ob[2], qt[2], ot[2], lb[2];
    float2 Dq[2];
    uint32_t keep[2][2];                                   // [block][query 2tig + e]: decisions of keys j0 + 8g .. +7
#pragma unroll
    for (int X = 0; X < 2; ++X) {
      const int qq = q0 + 8 * X;
      qb[X] = Qw[(qq + g) * 4 + tig];                      // (k = d 2tig, 2tig+1 ; n = query g)
      ob[X] = Ow[(qq + g) * 4 + tig];
      qt[X] = *reinterpret_cast<const uint32_t*>(QT + g * TP + qq + 2 * tig);    // (k = queries 2tig, 2tig+1 ; n = d g)
      ot[X] = *reinterpret_cast<const uint32_t*>(dOT + g * TP + qq + 2 * tig);
      lb[X] = LB[(qq + g) * 4 + tig];                      // (k = 8 + 2tig, 9 + 2tig ; n = query g): -lse' terms against the ones of `kone`
      Dq[X] = *reinterpret_cast<const float2*>(Ds + qq + 2 * tig);
      if (bitmask) {
        const int wcol = (j0 >> 5) + (g >> 2), sh = 8 * (g & 3);
        keep[X][0] = (smask[(qq + 2 * tig) * nblk + wcol] >> sh) & 0xffu;
        keep[X][1] = (smask[(qq + 2 * tig + 1) * nblk + wcol] >> sh) & 0xffu;
      } else {
        keep[X][0] = drop_bits8(drop, ((uint64_t)bh * T + (uint64_t)(qq + 2 * tig)) * (uint64_t)T + (uint64_t)(j0 + 8 * g));
        keep[X][1] = drop_bits8(drop, ((uint64_t)bh * T + (uint64_t)(qq + 2 * tig + 1)) * (uint64_t)T + (uint64_t)(j0 + 8 * g));
      }
    }
    float dq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int Y = 0; Y < 4; ++Y) {
      uint32_t pa[2][2], sa[2][2];                         // Pd and dS as A fragments: [block][rows g / g+8]
#pragma unroll
      for (int X = 0; X < 2; ++X) {
        float s[4], dp[4];
        mma_h16z(s, ka[Y][0], ka[Y][1], kone, kone, qb[X], lb[X]);
        mma_h8(dp, va[Y][0], va[Y][1], ob[X], zero);
        // element e: row (key) 8g+2Y+(e>>1), column (query) 2tig+(e&1)
        const float p0 = ex2(s[0]), p1 = ex2(s[1]), p2 = ex2(s[2]), p3 = ex2(s[3]);
        const float m0 = (keep[X][0] >> (2 * Y)) & 1u ? dscale : 0.f, m1 = (keep[X][1] >> (2 * Y)) & 1u ? dscale : 0.f;
        const float m2 = (keep[X][0] >> (2 * Y + 1)) & 1u ? dscale : 0.f, m3 = (keep[X][1] >> (2 * Y + 1)) & 1u ? dscale : 0.f;
        pa[X][0] = pack_h2(p0 * m0, p1 * m1);
        pa[X][1] = pack_h2(p2 * m2, p3 * m3);
        sa[X][0] = pack_h2(p0 * fmaf(dp[0], m0, -Dq[X].x), p1 * fmaf(dp[1], m1, -Dq[X].y));
        sa[X][1] = pack_h2(p2 * fmaf(dp[2], m2, -Dq[X].x), p3 * fmaf(dp[3], m3, -Dq[X].y));
      }
      mma_h16(dv[Y], pa[0][0], pa[0][1], pa[1][0], pa[1][1], ot[0], ot[1]);
      mma_h16(dk[Y], sa[0][0], sa[0][1], sa[1][0], sa[1][1], qt[0], qt[1]);
      // dQ (16 queries x 8 dims) += dS (queries x this tile's 16 keys) . K': the transposed dS fragments of both blocks fill all 16 rows
      mma_h16(dq, movm_t(sa[0][0]), movm_t(sa[1][0]), movm_t(sa[0][1]), movm_t(sa[1][1]), kt[Y][0], kt[Y][1]);
    }
    // dq[0], dq[1] = dQ'[query q0 + g][d = 2tig, 2tig+1];  dq[2], dq[3]: query q0 + 8 + g.  The block's dQ is complete once
    // every warp (= every key range) has contributed: the partials go to per-warp slots and 128 threads sum them IN WARP ORDER
    // and write the rows out -- no atomics, bitwise reproducible (the shared-memory float atomics this replaces were
    // compare-and-swap loops).  Slots are double buffered: one barrier per 16 queries.
    {
      float* slot = dqp + (((q0 >> 4) & 1) * nwarp + warp) * 128;
      *reinterpret_cast<float2*>(slot + g * 8 + 2 * tig) = make_float2(dq[0], dq[1]);
      *reinterpret_cast<float2*>(slot + (8 + g) * 8 + 2 * tig) = make_float2(dq[2], dq[3]);
      __syncthreads();
      for (int e = tid; e < 128; e += nthr) {              // (T = 128, 192 run 2, 3 warps: fewer threads than the 128 elements)
        const float* sl = dqp + ((q0 >> 4) & 1) * nwarp * 128 + e;
        float acc = sl[0];
        for (int w_ = 1; w_ < nwarp; ++w_) acc += sl[w_ * 128];
        dqkv[((long)b * T + q0 + (e >> 3)) * AQKV + h * AD + (e & 7)] = acc * sq;   // dQ = ln2 * sum (K' carries log2e / 8)
      }
    }
  }
  // ---- dK, dV of this warp's keys ----
  const float sk = 0.125f * inv;
#pragma unroll
  for (int Y = 0; Y < 4; ++Y) {
    float* r0 = dqkv + ((long)b * T + j0 + 8 * g + 2 * Y) * AQKV + h * AD + 2 * tig;
    float* r1 = r0 + AQKV;
    *reinterpret_cast<float2*>(r0 + 64) = make_float2(dk[Y][0] * sk, dk[Y][1] * sk);
    *reinterpret_cast<float2*>(r1 + 64) = make_float2(dk[Y][2] * sk, dk[Y][3] * sk);
    *reinterpret_cast<float2*>(r0 + 128) = make_float2(dv[Y][0] * inv, dv[Y][1] * inv);
    *reinterpret_cast<float2*>(r1 + 128) = make_float2(dv[Y][2] * inv, dv[Y][3] * inv);
  }
}

}  // namespace attntc

inline bool attention_tc_supported(int T) { return T >= 64 && (T % 64) == 0 && T <= 512; }

inline int attention_fwd_tc(const float* qkv, float* out, float* lse, int B, int T, const Drop& drop, cudaStream_t st) {
  const size_t smem = (size_t)T * 2 * sizeof(uint4) + (size_t)T * T / 8;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attntc::attn_fwd_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) != cudaSuccess) return EEGCLIP_ERR_CUDA;   // T = 512: 48 KB
    configured = true;
  }
  ProfScope prof(PROF_ATTN_FWD, st);
  LAUNCH_PDL((attntc::attn_fwd_h_kernel), B * AH, (T / 32) * 32, smem, st, qkv, out, lse, T, drop);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

inline int attention_bwd_tc(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv, int B, int T,
                            const Drop& drop, cudaStream_t st) {
  const int warps = T / 64;
  const size_t smem = (size_t)(16 * T + 16 * (T + 8)) * 2 + (size_t)(5 * T + 2 * (T / 64) * 128 + 8) * sizeof(float) + (size_t)T * T / 8;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attntc::attn_bwd_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess)   // T = 512: 92.4 KB
      return EEGCLIP_ERR_CUDA;
    configured = true;
  }
  ProfScope prof(PROF_ATTN_BWD, st);
  LAUNCH_PDL((attntc::attn_bwd_h_kernel), B * AH, warps * 32, smem, st, qkv, out, dout, lse, dqkv, T, drop);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

}  // namespace eegclip
