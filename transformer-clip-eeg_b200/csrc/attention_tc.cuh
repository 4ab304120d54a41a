// Tensor-core short-sequence attention (clip_model.py:30-45): 8 heads x head_dim 8, T <= 512, softmax(QK^T / sqrt(64)),
// dropout on the probabilities, no mask.  The (B,8,T,T) energy / probability tensors never exist.
//
// head_dim = 8 is exactly the K extent of mma.sync.m16n8k8 (TF32 operands, fp32 accumulate): one MMA produces a 16 x 8
// score tile, and its accumulator fragment can be re-used in place as the A fragment of the next MMA (P.V, P^T.dO,
// dS^T.Q) if the 8 contracted indices are taken in the order the fragment already has them -- so probabilities never
// move between threads.  tcgen05 is not used here on purpose: the contraction dimension is 8, the work per
// score is dominated by exp2 + Philox, not by the MMA, and the warp-level MMA keeps everything in registers.
//
// Dropout grouping: one Philox call yields 8 decisions for 8 consecutive key indices of one query row (common.cuh), so
// the key <-> fragment-column assignment is permuted such that every thread owns 8 consecutive keys of its rows:
//   forward  (rows = queries): a 32-key block = 4 n-tiles X; tile X column c  <->  key 8*(c>>1) + 2X + (c&1)
//   backward (rows = keys)   : a warp owns 64 keys = 4 m-tiles Y; tile Y row g <-> key 8g+2Y, row g+8 <-> key 8g+2Y+1
// The B/A fragments of K and V are laid out in shared memory / registers in exactly that order.
//
// Numerics: Q,K,V,P,dO,dS are rounded to TF32 (cvt.rna) before the MMAs, accumulation and softmax statistics are fp32.
#pragma once
#include "common.cuh"
#include "attention.cuh"

namespace eegclip {
namespace attntc {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int QS = 12;   // padded row stride (floats) of the Q / dO shared-memory copies: conflict-free for both B-fragment patterns

__device__ __forceinline__ uint32_t tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tf32f(float x) { return __uint_as_float(tf32(x)); }
// cvt.rna.tf32.f32 for FINITE inputs in two integer instructions: ptxas lowers the PTX conversion to IADD + FSETP(+inf) + SEL + LOP3
// (the compare / select only protect inf / NaN).  Probabilities and dS in the inner loops are finite; adding half an ulp of the 13
// dropped bits to the sign-magnitude pattern and truncating is round-to-nearest, ties away from zero -- exactly .rna.
__device__ __forceinline__ uint32_t tf32_fin(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

// D(16x8) = A(16x8, row) * B(8x8, col) + C
__device__ __forceinline__ void mma_tf32(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1,
                                         const float* c) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}

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
// Forward.  grid = B*H, block = (T/32) warps; warp w owns query rows [32w, 32w+32) as two 16-row tiles (two independent
// MMA / exp2 chains per warp).  smem: Kf / Vf fragment tables, float4 [T/32 blocks][2 halves][32 lanes] each, + keep bits.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 2) attn_fwd_tc_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                            float* __restrict__ lse, int T, Drop drop) {
  pdl_sync();
  extern __shared__ float4 smf[];
  float4* Kf = smf;
  float4* Vf = smf + (T / 32) * 64;
  uint32_t* smask = reinterpret_cast<uint32_t*>(Vf + (T / 32) * 64);
  const int bh = blockIdx.x, b = bh / AH, h = bh % AH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const float* base = qkv + (long)b * T * AQKV;
  const int nblk = T >> 5;
  // ---- build the fragment tables: one (block, lane) entry of K or V per iteration (2T entries) ----
  for (int e = tid; e < 2 * T; e += blockDim.x) {
    const int isv = e >= T;
    const int r = isv ? e - T : e;
    const int blk = r >> 5, l = r & 31, eg = l >> 2, et = l & 3;
    float v[8];
    if (!isv) {
      // S tile X, column eg  <->  key 8*(eg>>1) + 2X + (eg&1);  b0 = K[key][et], b1 = K[key][et+4]
#pragma unroll
      for (int X = 0; X < 4; ++X) {
        const float* kr = base + (long)(blk * 32 + 8 * (eg >> 1) + 2 * X + (eg & 1)) * AQKV + 64 + h * AD;
        v[2 * X] = tf32f(__ldg(kr + et));
        v[2 * X + 1] = tf32f(__ldg(kr + et + 4));
      }
    } else {
      // P.V k-index et <-> key 8et+2X, et+4 <-> key 8et+2X+1 ; b = V[key][d = eg]  => 8 consecutive keys
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = tf32f(__ldg(base + (long)(blk * 32 + 8 * et + j) * AQKV + 128 + h * AD + eg));
    }
    float4* dst = (isv ? Vf : Kf) + blk * 64 + l;
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[32] = make_float4(v[4], v[5], v[6], v[7]);
  }
  const bool bitmask = drop.enabled && drop.onebit;
  if (bitmask) build_keep_bits(smask, drop, bh, T);
  // ---- Q fragments of this warp (two row tiles), scaled by log2(e)/sqrt(64) ----
  const int i0 = warp * 32;
  const float qs = LOG2E * 0.125f;
  uint32_t qa[2][4];
#pragma unroll
  for (int R = 0; R < 2; ++R) {
    const float* q0p = base + (long)(i0 + 16 * R + g) * AQKV + h * AD;
    const float* q1p = q0p + 8 * AQKV;
    qa[R][0] = tf32(__ldg(q0p + tig) * qs); qa[R][1] = tf32(__ldg(q1p + tig) * qs);
    qa[R][2] = tf32(__ldg(q0p + tig + 4) * qs); qa[R][3] = tf32(__ldg(q1p + tig + 4) * qs);
  }
  __syncthreads();
  const float zero[4] = {0.f, 0.f, 0.f, 0.f};
  // ---- pass 1: row maxima ----
  float mx[2][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}};
  for (int blk = 0; blk < nblk; ++blk) {
    const float4 k0 = Kf[blk * 64 + lane], k1 = Kf[blk * 64 + 32 + lane];
    const float kb[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
    for (int X = 0; X < 4; ++X)
#pragma unroll
      for (int R = 0; R < 2; ++R) {
        float s[4];
        mma_tf32(s, qa[R][0], qa[R][1], qa[R][2], qa[R][3], __float_as_uint(kb[2 * X]), __float_as_uint(kb[2 * X + 1]), zero);
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
  float l[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int blk = 0; blk < nblk; ++blk) {
    const float4 k0 = Kf[blk * 64 + lane], k1 = Kf[blk * 64 + 32 + lane];
    const float4 v0 = Vf[blk * 64 + lane], v1 = Vf[blk * 64 + 32 + lane];
    const float kb[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    const float vb[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
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
      // the row maximum enters as the accumulator input of the score MMA: S - max costs no instruction
      const float negmx[4] = {-mx[R][0], -mx[R][0], -mx[R][1], -mx[R][1]};
#pragma unroll
      for (int X = 0; X < 4; ++X) {
        float s[4];
        mma_tf32(s, qa[R][0], qa[R][1], qa[R][2], qa[R][3], __float_as_uint(kb[2 * X]), __float_as_uint(kb[2 * X + 1]), negmx);
        const float p0 = ex2(s[0]), p1 = ex2(s[1]), p2 = ex2(s[2]), p3 = ex2(s[3]);
        l[R][0] += p0 + p1; l[R][1] += p2 + p3;
        // this thread's keys of the block: 8tig + 2X (+1)
        const uint32_t pa0 = (keep0 >> (2 * X)) & 1u ? tf32_fin(p0) : 0u;
        const uint32_t pa2 = (keep0 >> (2 * X + 1)) & 1u ? tf32_fin(p1) : 0u;
        const uint32_t pa1 = (keep1 >> (2 * X)) & 1u ? tf32_fin(p2) : 0u;
        const uint32_t pa3 = (keep1 >> (2 * X + 1)) & 1u ? tf32_fin(p3) : 0u;
        mma_tf32(o[R], pa0, pa1, pa2, pa3, __float_as_uint(vb[2 * X]), __float_as_uint(vb[2 * X + 1]), o[R]);
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

// ------------------------------------------------------------------------------------------------
// Backward.  grid = B*H, block = (T/64) warps; warp w owns keys [64w, 64w+64) and loops over all query tiles of 8.
//   S^T = K'.Q^T, dP^T = V.dO^T  (rows = keys)   ->   dV += Pd^T.dO, dK += dS^T.Q  (accumulators in registers)
//   dQ^T = K'^T.dS^T needs dS^T as a B fragment: staged through a per-warp 64 x 8 shared tile, accumulated over the
//   warps of the CTA in shared memory (red.shared), written once at the end.  (Measured alternative, round 2: a private
//   [T][8] dQ slab per warp summed in warp order is bitwise reproducible but needs 105 KB instead of 66 KB of shared memory at
//   T = 320 -- occupancy 3 -> 2 CTAs per SM -- and ran 2.86 -> 3.49 ms per step; the float atomics stay.)
// smem: Qs, dOs [T][QS] (TF32), Ls (lse * log2e), Ds (dO.O) [T], dQs [T][8], per-warp staging [64][8], keep bits [T][T/32].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) attn_bwd_tc_kernel(const float* __restrict__ qkv, const float* __restrict__ out,
                                                            const float* __restrict__ dout, const float* __restrict__ lse,
                                                            float* __restrict__ dqkv, int T, Drop drop) {
  pdl_sync();
  extern __shared__ float smb[];
  float* Qs = smb;
  float* dOs = Qs + T * QS;
  float* Ls = dOs + T * QS;
  float* Ds = Ls + T;
  float* dQs = Ds + T;
  float* stg_all = dQs + T * 8;
  uint32_t* smask = reinterpret_cast<uint32_t*>(stg_all + (blockDim.x >> 5) * 512);
  const bool bitmask = drop.enabled && drop.onebit;
  const int nblk = T >> 5;
  const int bh = blockIdx.x, b = bh / AH, h = bh % AH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const float* base = qkv + (long)b * T * AQKV;
  float* stg = stg_all + warp * 512;
  // ---- stage Q, dO (TF32), lse', D, zero dQ ----
  for (int i = tid; i < T; i += blockDim.x) {
    float q[8], dO[8], o[8];
    load8(base + (long)i * AQKV + h * AD, q);
    load8(dout + ((long)b * T + i) * AE + h * AD, dO);
    load8(out + ((long)b * T + i) * AE + h * AD, o);
    float dsum = 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      dsum = fmaf(dO[d], o[d], dsum);
      Qs[i * QS + d] = tf32f(q[d]);
      dOs[i * QS + d] = tf32f(dO[d]);
      dQs[i * 8 + d] = 0.f;
    }
    Ds[i] = dsum;
    Ls[i] = lse[(long)bh * T + i] * LOG2E;
  }
  if (bitmask) build_keep_bits(smask, drop, bh, T);
  // ---- A fragments of this warp's 64 keys ----
  const int j0 = warp * 64;
  const float ks = LOG2E * 0.125f;
  uint32_t ka[4][4], va[4][4], kt[8][2];
#pragma unroll
  for (int Y = 0; Y < 4; ++Y) {
    // tile Y: row g <-> key 8g+2Y, row g+8 <-> key 8g+2Y+1
    const float* k0 = base + (long)(j0 + 8 * g + 2 * Y) * AQKV + 64 + h * AD;
    const float* k1 = k0 + AQKV;
    ka[Y][0] = tf32(__ldg(k0 + tig) * ks); ka[Y][1] = tf32(__ldg(k1 + tig) * ks);
    ka[Y][2] = tf32(__ldg(k0 + tig + 4) * ks); ka[Y][3] = tf32(__ldg(k1 + tig + 4) * ks);
    const float* v0 = k0 + 64;
    const float* v1 = k1 + 64;
    va[Y][0] = tf32(__ldg(v0 + tig)); va[Y][1] = tf32(__ldg(v1 + tig));
    va[Y][2] = tf32(__ldg(v0 + tig + 4)); va[Y][3] = tf32(__ldg(v1 + tig + 4));
  }
  // dQ^T = K'^T . dS^T : A rows = d (g; rows 8..15 are zero), k-step s covers keys 8s..8s+7: a0 = K'[8s+tig][g], a2 = K'[8s+tig+4][g]
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const float* kr = base + (long)(j0 + 8 * s + tig) * AQKV + 64 + h * AD + g;
    kt[s][0] = tf32(__ldg(kr) * ks);
    kt[s][1] = tf32(__ldg(kr + 4 * AQKV) * ks);
  }
  float dk[4][4], dv[4][4];
#pragma unroll
  for (int Y = 0; Y < 4; ++Y)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dk[Y][e] = 0.f; dv[Y][e] = 0.f; }
  __syncthreads();
  const float zero[4] = {0.f, 0.f, 0.f, 0.f};
  const float dscale = drop.scale;
  // Cost split of this loop (timing experiments at B = 256, depth 10): the whole dQ product below (8 half-filled MMAs, 16 LDS,
  // staging stores, the compare-and-swap adds) is 0.77 ms of the 2.88 ms per step; the compare-and-swap adds alone 0.12 ms.
  // (measured and rejected: starting every warp at a different query block so that the shared-memory compare-and-swap adds of
  //  dQ do not collide -- 3.16 ms per step against 2.88 ms for the lock-step order)
  for (int q0 = 0; q0 < T; q0 += 8) {
    // B fragments: Q^T / dO^T (k = d, n = query)  and  Q / dO (k = query pair of this thread, n = d)
    const uint32_t qb0 = __float_as_uint(Qs[(q0 + g) * QS + tig]), qb1 = __float_as_uint(Qs[(q0 + g) * QS + tig + 4]);
    const uint32_t ob0 = __float_as_uint(dOs[(q0 + g) * QS + tig]), ob1 = __float_as_uint(dOs[(q0 + g) * QS + tig + 4]);
    const uint32_t qc0 = __float_as_uint(Qs[(q0 + 2 * tig) * QS + g]), qc1 = __float_as_uint(Qs[(q0 + 2 * tig + 1) * QS + g]);
    const uint32_t oc0 = __float_as_uint(dOs[(q0 + 2 * tig) * QS + g]), oc1 = __float_as_uint(dOs[(q0 + 2 * tig + 1) * QS + g]);
    const float2 L = *reinterpret_cast<const float2*>(Ls + q0 + 2 * tig);
    const float2 Dq = *reinterpret_cast<const float2*>(Ds + q0 + 2 * tig);
    // dropout decisions of this thread's 8 keys (8g .. 8g+7 of the warp's 64) for its two queries
    uint32_t keepa, keepb;
    if (bitmask) {
      const int wcol = (j0 >> 5) + (g >> 2), sh = 8 * (g & 3);
      keepa = (smask[(q0 + 2 * tig) * nblk + wcol] >> sh) & 0xffu;
      keepb = (smask[(q0 + 2 * tig + 1) * nblk + wcol] >> sh) & 0xffu;
    } else {
      keepa = drop_bits8(drop, ((uint64_t)bh * T + (uint64_t)(q0 + 2 * tig)) * (uint64_t)T + (uint64_t)(j0 + 8 * g));
      keepb = drop_bits8(drop, ((uint64_t)bh * T + (uint64_t)(q0 + 2 * tig + 1)) * (uint64_t)T + (uint64_t)(j0 + 8 * g));
    }
    // -lse' of the two query columns enters as the accumulator input of the score MMA (S - lse' costs no instruction)
    const float negL[4] = {-L.x, -L.y, -L.x, -L.y};
#pragma unroll
    for (int Y = 0; Y < 4; ++Y) {
      float s[4], dp[4];
      mma_tf32(s, ka[Y][0], ka[Y][1], ka[Y][2], ka[Y][3], qb0, qb1, negL);
      mma_tf32(dp, va[Y][0], va[Y][1], va[Y][2], va[Y][3], ob0, ob1, zero);
      // fragment element e: row (key) 8g+2Y+(e>>1), column (query) 2tig+(e&1)
      const float p0 = ex2(s[0]), p1 = ex2(s[1]), p2 = ex2(s[2]), p3 = ex2(s[3]);
      const float k0m = (keepa >> (2 * Y)) & 1u ? dscale : 0.f, k1m = (keepb >> (2 * Y)) & 1u ? dscale : 0.f;
      const float k2m = (keepa >> (2 * Y + 1)) & 1u ? dscale : 0.f, k3m = (keepb >> (2 * Y + 1)) & 1u ? dscale : 0.f;
      const float pd0 = p0 * k0m, pd1 = p1 * k1m, pd2 = p2 * k2m, pd3 = p3 * k3m;
      const float ds0 = p0 * (dp[0] * k0m - Dq.x), ds1 = p1 * (dp[1] * k1m - Dq.y);
      const float ds2 = p2 * (dp[2] * k2m - Dq.x), ds3 = p3 * (dp[3] * k3m - Dq.y);
      // accumulator fragment -> A fragment (k-index tig <-> query 2tig, tig+4 <-> query 2tig+1)
      mma_tf32(dv[Y], tf32_fin(pd0), tf32_fin(pd2), tf32_fin(pd1), tf32_fin(pd3), oc0, oc1, dv[Y]);
      const uint32_t t0 = tf32_fin(ds0), t1 = tf32_fin(ds1), t2 = tf32_fin(ds2), t3 = tf32_fin(ds3);
      mma_tf32(dk[Y], t0, t2, t1, t3, qc0, qc1, dk[Y]);
      // stage dS^T[key][query] for the dQ product
      *reinterpret_cast<float2*>(stg + (8 * g + 2 * Y) * 8 + 2 * tig) = make_float2(__uint_as_float(t0), __uint_as_float(t1));
      *reinterpret_cast<float2*>(stg + (8 * g + 2 * Y + 1) * 8 + 2 * tig) = make_float2(__uint_as_float(t2), __uint_as_float(t3));
    }
    __syncwarp();
    float dq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s8 = 0; s8 < 8; ++s8) {
      const uint32_t b0 = __float_as_uint(stg[(8 * s8 + tig) * 8 + g]), b1 = __float_as_uint(stg[(8 * s8 + tig + 4) * 8 + g]);
      mma_tf32(dq, kt[s8][0], 0u, kt[s8][1], 0u, b0, b1, dq);
    }
    __syncwarp();
    // dq[0], dq[1] = dQ'[query 2tig, 2tig+1][d = g] (rows 8..15 of the tile are padding)
    atomicAdd(dQs + (q0 + 2 * tig) * 8 + g, dq[0]);
    atomicAdd(dQs + (q0 + 2 * tig + 1) * 8 + g, dq[1]);
  }
  // ---- dK, dV of this warp's keys ----
#pragma unroll
  for (int Y = 0; Y < 4; ++Y) {
    float* r0 = dqkv + ((long)b * T + j0 + 8 * g + 2 * Y) * AQKV + h * AD + 2 * tig;
    float* r1 = r0 + AQKV;
    *reinterpret_cast<float2*>(r0 + 64) = make_float2(dk[Y][0] * 0.125f, dk[Y][1] * 0.125f);
    *reinterpret_cast<float2*>(r1 + 64) = make_float2(dk[Y][2] * 0.125f, dk[Y][3] * 0.125f);
    *reinterpret_cast<float2*>(r0 + 128) = make_float2(dv[Y][0], dv[Y][1]);
    *reinterpret_cast<float2*>(r1 + 128) = make_float2(dv[Y][2], dv[Y][3]);
  }
  __syncthreads();
  // dQ = ln2 * sum over warps (K' carries log2e/8)
  for (int i = tid; i < T * 2; i += blockDim.x) {
    const int t = i >> 1, half = i & 1;
    float4 v = *reinterpret_cast<const float4*>(dQs + t * 8 + half * 4);
    v.x *= LN2; v.y *= LN2; v.z *= LN2; v.w *= LN2;
    *reinterpret_cast<float4*>(dqkv + ((long)b * T + t) * AQKV + h * AD + half * 4) = v;
  }
}

}  // namespace attntc

inline bool attention_tc_supported(int T) { return T >= 64 && (T % 64) == 0 && T <= 512; }

inline int attention_fwd_tc(const float* qkv, float* out, float* lse, int B, int T, const Drop& drop, cudaStream_t st) {
  const size_t smem = (size_t)(T / 32) * 64 * 2 * sizeof(float4) + (size_t)T * T / 8;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attntc::attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess)
      return EEGCLIP_ERR_CUDA;
    configured = true;
  }
  ProfScope prof(PROF_ATTN_FWD, st);
  LAUNCH_PDL((attntc::attn_fwd_tc_kernel), B * AH, (T / 32) * 32, smem, st, qkv, out, lse, T, drop);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

inline int attention_bwd_tc(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv, int B, int T,
                            const Drop& drop, cudaStream_t st) {
  const int warps = T / 64;
  const size_t smem = ((size_t)T * (2 * attntc::QS + 2 + 8) + (size_t)warps * 512) * sizeof(float) + (size_t)T * T / 8;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attntc::attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024) != cudaSuccess)   // T = 512: 118.8 KB
      return EEGCLIP_ERR_CUDA;
    configured = true;
  }
  ProfScope prof(PROF_ATTN_BWD, st);
  LAUNCH_PDL((attntc::attn_bwd_tc_kernel), B * AH, warps * 32, smem, st, qkv, out, dout, lse, dqkv, T, drop);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

}  // namespace eegclip
