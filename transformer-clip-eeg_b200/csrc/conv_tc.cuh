// tcgen05 implicit-GEMM for Conv1d(64 -> 64, k = 64, 'same') on time-major activations: forward, data gradient and
// weight gradient (clip_model.py:237,245 -- 74 % of the EEG tower's FLOPs).
//
// Forward / data gradient (conv64_tc_kernel): one CTA per sample keeps the whole zero-padded activation tile
// (T+63 rows x 64 channels, bf16 hi+lo planes, chunk-major -- see tc_common.cuh) resident in shared memory.  Tap k of
// the convolution is the SAME tile read through a descriptor whose start address is shifted by k rows (+16 B), so the
// K = 4096 contraction is 64 taps x 4 K16-steps of tcgen05.mma with no im2col and no re-staging.  Weights (16 KB per
// tap, pre-split and pre-arranged in the UMMA layout by pack_conv_weights) stream through a 4-stage ring filled by
// 1-D bulk async copies (TMA engine) signalled on mbarriers.  Accumulators for all time tiles (128,128,64 rows at
// T=320) live in TMEM for the whole kernel; the epilogue reads them with tcgen05.ld, adds bias, applies the Philox
// dropout mask and stores fp32.
// The data gradient is the same kernel over the zero-padded output gradient with flipped / transposed weights.
//
// Weight gradient (wgrad64_tc_kernel): dW[co][ci][k] = sum_{b,t} dy[b,t,co] * u[b,t+k-31,ci].  Both operands are
// MN-major views of the same chunk-major tiles (K = time); each CTA owns 16 taps (16 M=64 x N=64 accumulators =
// all 512 TMEM columns, two per column block in the two 16-lane halves) and a slice of the batch, accumulating over its
// samples in TMEM; per-CTA partials are reduced by wgrad_reduce_kernel.
//
// Arithmetic: NTERMS = 3 -> split-bf16 (hi*hi + hi*lo + lo*hi, fp32 accumulate, ~2^-16 relative, meets the 1e-3
// gradient tolerance); NTERMS = 1 -> plain bf16 (fast mode, does not meet it).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace eegclip {
namespace convtc {

constexpr int CH = 64;               // channels in == out
constexpr int MAX_TAPS = 64;         // kernel sizes 16..64 (multiples of 16): 64 for the EEG tower, 32 for the speech BasicBlock
constexpr int NSTAGE = 4;            // weight ring depth
constexpr int W_PLANE_BYTES = CH * CH * 2;      // 8 KB: one tap, one plane, [ci chunk][co][8]
constexpr int W_TAP_BYTES = 2 * W_PLANE_BYTES;  // hi + lo
constexpr int WG_TAPS = 16;          // taps per weight-gradient CTA
constexpr int WG_MAX_GROUPS = 37;    // 4 tap groups x 37 sample groups = 148 CTAs

// ------------------------------------------------------------------------------------------------
// Weight packing: fp32 W[co][ci][k]  ->  bf16, per tap [k-chunk][n' = 128][8] with n' < 64 the hi part of output n' and
// n' >= 64 the lo part of output n' - 64: ONE N=128 MMA computes A_hi.W_hi and A_hi.W_lo (the activation tile, the larger
// operand, is read once for both terms -- the kernel is shared-memory-bandwidth bound at N = 64), a second N=64 MMA over the
// first 64 rows of the same operand adds A_lo.W_hi.
//   mode 0 (forward) : n = co, contraction index = ci, tap = k
//   mode 1 (dgrad)   : n = ci, contraction index = co, tap k' holds W[..][..][63-k']
// ------------------------------------------------------------------------------------------------
__global__ void pack_conv_weights_kernel(const float* __restrict__ W, uint8_t* __restrict__ out, int mode, int TAPS) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // over taps * 8 chunks * 64 n
  if (i >= TAPS * 8 * CH) return;
  int n = i & 63, ch = (i >> 6) & 7, tap = i >> 9;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    int kk = ch * 8 + e;  // contraction index
    v[e] = mode == 0 ? W[((long)n * CH + kk) * TAPS + tap] : W[((long)kk * CH + n) * TAPS + (TAPS - 1 - tap)];
  }
  uint4 hi, lo;
  tc::split8(v, hi, lo);
  uint8_t* base = out + (long)tap * W_TAP_BYTES + ch * (2 * CH * 16) + n * 16;
  *reinterpret_cast<uint4*>(base) = hi;
  *reinterpret_cast<uint4*>(base + CH * 16) = lo;
}

struct ConvTcArgs {
  const float* src;       // (B, src_rows, 64) fp32
  const float* skip;      // optional addend, same shape as src
  const uint8_t* wpacked; // TAPS * W_TAP_BYTES
  const float* bias;      // optional
  float* out;             // (B, T, 64)
  int T;                  // output rows
  int src_rows;           // valid source rows
  int row_off;            // smem row r holds src row r - row_off (zero outside)
  int taps;               // kernel size
  Drop drop;
  unsigned long long* dbg; // development timeline (CTA 0): start, staged, mma-done, end (globaltimer ns)
  int exp_mode;            // development experiments (g_tune[5]); 0 = shipped path
};

__host__ __device__ inline uint32_t conv_smem_bytes(int T, int taps) {
  uint32_t TP = T + taps - 1;
  return 2u * 8u * TP * 16u + NSTAGE * W_TAP_BYTES + 128;
}

template <int NTERMS>
__global__ void __launch_bounds__(256, 1) conv64_tc_kernel(const ConvTcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const int T = a.T, TAPS = a.taps, TP = T + TAPS - 1;
  const uint32_t CS = (uint32_t)TP * 16u;   // chunk stride (bytes)
  const uint32_t PS = 8u * CS;              // plane stride
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2u * PS;             // 2*PS = 256*TP: 128-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NSTAGE * W_TAP_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + NSTAGE;
  uint64_t* accfull = bars + 2 * NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 1);

  const int ntiles = (T + 127) >> 7;
  const bool last64 = (T & 127) != 0 && !(a.exp_mode & 1);   // T % 128 == 64 -> last tile has M = 64 (exp bit 0: run it as M = 128 over-read)
  constexpr uint32_t TCOLS = NTERMS > 1 ? 128u : 64u;   // TMEM columns per time tile: [hi.hi + lo.hi | hi.lo]
  uint32_t ncols = 64;
  while (ncols < (uint32_t)ntiles * TCOLS) ncols <<= 1;

  auto stamp = [&](int slot) {
    if (a.dbg && blockIdx.x == 0 && tid == 128) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      a.dbg[slot] = t;
    }
  };
  stamp(0);
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(accfull, 1);
    tc::mbar_fence_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, ncols);

  // ---- stage the activation tile: fp32 (+skip) -> bf16 hi/lo, chunk-major, zero padded ----
  // Loads are issued STAGE_U items deep before the first conversion: the tile is 82-164 KB per CTA and a load-convert-store
  // loop exposes one DRAM round trip per iteration (ncu: long_scoreboard was the top stall of this kernel).
  {
    const float* sb = a.src + (long)b * a.src_rows * CH;
    const float* kb = a.skip ? a.skip + (long)b * a.src_rows * CH : nullptr;
    constexpr int STAGE_U = 6;
    const int total = TP * 8;
    for (int base = 0; base < total; base += 256 * STAGE_U) {
      float4 x[STAGE_U][2], y[STAGE_U][2];
#pragma unroll
      for (int u = 0; u < STAGE_U; ++u) {
        const int idx = base + u * 256 + tid;
        const int r = idx >> 3, ch = idx & 7;
        const int t = r - a.row_off;
        x[u][0] = x[u][1] = y[u][0] = y[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < total && t >= 0 && t < a.src_rows) {
          const float4* p = reinterpret_cast<const float4*>(sb + (long)t * CH + ch * 8);
          x[u][0] = __ldg(p); x[u][1] = __ldg(p + 1);
          if (kb) {
            const float4* q = reinterpret_cast<const float4*>(kb + (long)t * CH + ch * 8);
            y[u][0] = __ldg(q); y[u][1] = __ldg(q + 1);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < STAGE_U; ++u) {
        const int idx = base + u * 256 + tid;
        if (idx < total) {
          const int r = idx >> 3, ch = idx & 7;
          const float v[8] = {x[u][0].x + y[u][0].x, x[u][0].y + y[u][0].y, x[u][0].z + y[u][0].z, x[u][0].w + y[u][0].w,
                              x[u][1].x + y[u][1].x, x[u][1].y + y[u][1].y, x[u][1].z + y[u][1].z, x[u][1].w + y[u][1].w};
          uint4 hi, lo;
          tc::split8(v, hi, lo);
          uint8_t* d = sA + ch * CS + r * 16;
          *reinterpret_cast<uint4*>(d) = hi;
          if (NTERMS > 1) *reinterpret_cast<uint4*>(d + PS) = lo;
        }
      }
    }
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  stamp(1);

  if (warp == 0 && lane == 0) {
    // ===== weight producer: one 1-D bulk copy per tap into the ring =====
    const uint32_t bytes = W_TAP_BYTES;
    for (int tap = 0; tap < TAPS; ++tap) {
      const int s = tap % NSTAGE;
      const uint32_t ph = (tap / NSTAGE) & 1;
      tc::mbar_wait(&empty[s], ph ^ 1);
      if ((a.exp_mode & 2) && tap >= NSTAGE) { tc::mbar_arrive(&full[s]); continue; }   // timing experiment: no weight re-fetch
      tc::mbar_expect_tx(&full[s], bytes);
      tc::bulk_g2s(sB + s * W_TAP_BYTES, a.wpacked + (long)tap * W_TAP_BYTES, bytes, &full[s]);
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop (descriptors stay in uniform registers), one elected lane issues =====
    const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB);
    // idesc[m64][wide]: wide = N 128 (hi.hi | hi.lo in one MMA), narrow = N 64
    const uint32_t idw128 = tc::idesc_bf16(128, 2 * CH, 0, 0), idw64 = tc::idesc_bf16(64, 2 * CH, 0, 0);
    const uint32_t idn128 = tc::idesc_bf16(128, CH, 0, 0), idn64 = tc::idesc_bf16(64, CH, 0, 0);
    for (int tap = 0; tap < TAPS; ++tap) {
      const int s = tap % NSTAGE;
      const uint32_t ph = (tap / NSTAGE) & 1;
      tc::mbar_wait(&full[s], ph);
      tc::tc_fence_after();
      const uint32_t wb = sB_u + s * W_TAP_BYTES;
      const uint64_t b_w = tc::smem_desc(wb, 2 * CH * 16, 128);   // chunk stride 2048 B: rows 0-63 hi, 64-127 lo
      if (tc::elect_one()) {
        for (int tile = 0; tile < ntiles; ++tile) {
          const bool m64 = last64 && tile == ntiles - 1;
          const uint32_t idw = m64 ? idw64 : idw128, idn = m64 ? idn64 : idn128;
          const uint32_t d = tmem + tile * TCOLS;
          const uint32_t arow = sA_u + (uint32_t)(tile * 128 + tap) * 16u;
          const uint64_t a_hi = tc::smem_desc(arow, CS, 128), a_lo = tc::smem_desc(arow + PS, CS, 128);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t da = (uint64_t)((2 * ks * CS) >> 4);           // start-address field is in 16-byte units
            const uint64_t db = (uint64_t)((2 * ks * (2 * CH * 16)) >> 4);
            if (NTERMS > 1) {
              tc::mma_bf16(d, a_hi + da, b_w + db, idw, (tap | ks) != 0);   // [hi.hi | hi.lo]
              if (!(a.exp_mode & 4)) tc::mma_bf16(d, a_lo + da, b_w + db, idn, 1);   // += lo.hi (first 64 columns)
            } else {
              tc::mma_bf16(d, a_hi + da, b_w + db, idn, (tap | ks) != 0);
            }
          }
        }
        tc::tc_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::tc_commit(accfull);
    __syncwarp();
  }
  __syncwarp();
  {
    // ===== epilogue (all 8 warps): TMEM -> registers -> bias / dropout -> global.  Warp w reads TMEM lane quarter w % 4;
    // warps 4-7 take output channels 0-31, warps 0-3 (done with their producer / MMA roles by now) channels 32-63 =====
    const int q = warp & 3;
    const int half = warp < 4 ? 1 : 0;
    tc::mbar_wait(accfull, 0);
    tc::tc_fence_after();
    stamp(2);
    float4 bb[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bb[c] = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + half * 32) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int tile = 0; tile < ntiles; ++tile) {
      const bool m64 = last64 && tile == ntiles - 1;
      const int row = m64 ? tile * 128 + q * 16 + lane : tile * 128 + q * 32 + lane;
      const bool valid = m64 ? (lane < 16) : true;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + tile * TCOLS;
      float v[32];
      tc::tmem_ld32(taddr + half * 32, v);
      if (NTERMS > 1) {
        float v2[32];
        tc::tmem_ld32(taddr + CH + half * 32, v2);
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] += v2[c];
      }
      if (valid && row < T) {
        float* o = a.out + ((long)b * T + row) * CH + half * 32;
        const uint64_t didx = ((uint64_t)b * T + row) * CH + half * 32;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float4 r = make_float4(v[c] + bb[c >> 2].x, v[c + 1] + bb[c >> 2].y, v[c + 2] + bb[c >> 2].z, v[c + 3] + bb[c >> 2].w);
          float4 m = drop_mult4(a.drop, didx + c);
          r.x *= m.x; r.y *= m.y; r.z *= m.z; r.w *= m.w;
          *reinterpret_cast<float4*>(o + c) = r;
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  stamp(3);
  if (warp == 2) tc::tmem_dealloc(tmem, ncols);
}

// ------------------------------------------------------------------------------------------------
// Weight gradient
// ------------------------------------------------------------------------------------------------
struct WgradTcArgs {
  const float* xin;     // (B,T,64) conv input
  const float* skip;    // optional addend
  const float* dypad;   // (B,TP,64) zero-padded output gradient, valid rows [PLb, PLb+T)
  float* partial;       // [groups][taps][64 co][64 ci]
  int B, T, PL, PLb, groups, taps;
};

__host__ __device__ inline uint32_t wgrad_smem_bytes(int T) {
  return 2u * 8u * (uint32_t)(T + WG_TAPS - 1) * 16u + 2u * 8u * (uint32_t)T * 16u + 128;
}

template <int NTERMS>
__global__ void __launch_bounds__(256, 1) wgrad64_tc_kernel(const WgradTcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tg = blockIdx.x, grp = blockIdx.y;
  const int T = a.T, TAPS = a.taps, TU = T + WG_TAPS - 1;
  const int k0 = tg * WG_TAPS;
  const uint32_t CSU = (uint32_t)TU * 16u, PSU = 8u * CSU;
  const uint32_t CSD = (uint32_t)T * 16u, PSD = 8u * CSD;
  uint8_t* sU = smem;
  uint8_t* sD = smem + 2u * PSU;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sD + 2u * PSD);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
  if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = tc::idesc_bf16(64, CH, 1, 1);   // both operands MN-major (K = time)
  uint32_t phase = 0;
  bool first = true;
  for (int b = grp; b < a.B; b += a.groups) {
    // ---- stage u rows [k0 - PL, k0 - PL + TU) of the padded input and the T rows of dy (loads issued WG_U deep) ----
    const float* xb = a.xin + (long)b * T * CH;
    const float* kb = a.skip ? a.skip + (long)b * T * CH : nullptr;
    const float* db = a.dypad + ((long)b * (T + TAPS - 1) + a.PLb) * CH;
    constexpr int WG_U = 6;
    const int nu = TU * 8, total = nu + T * 8;       // items [0, nu): u tile; [nu, total): dy tile
    for (int base = 0; base < total; base += 256 * WG_U) {
      float4 x[WG_U][2], y[WG_U][2];
#pragma unroll
      for (int u = 0; u < WG_U; ++u) {
        const int idx = base + u * 256 + tid;
        x[u][0] = x[u][1] = y[u][0] = y[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < nu) {
          const int r = idx >> 3, ch = idx & 7;
          const int t = r + k0 - a.PL;
          if (t >= 0 && t < T) {
            const float4* p = reinterpret_cast<const float4*>(xb + (long)t * CH + ch * 8);
            x[u][0] = __ldg(p); x[u][1] = __ldg(p + 1);
            if (kb) {
              const float4* q = reinterpret_cast<const float4*>(kb + (long)t * CH + ch * 8);
              y[u][0] = __ldg(q); y[u][1] = __ldg(q + 1);
            }
          }
        } else if (idx < total) {
          const int i2 = idx - nu;
          const float4* p = reinterpret_cast<const float4*>(db + (long)(i2 >> 3) * CH + (i2 & 7) * 8);
          x[u][0] = __ldg(p); x[u][1] = __ldg(p + 1);
        }
      }
#pragma unroll
      for (int u = 0; u < WG_U; ++u) {
        const int idx = base + u * 256 + tid;
        if (idx < total) {
          const float v[8] = {x[u][0].x + y[u][0].x, x[u][0].y + y[u][0].y, x[u][0].z + y[u][0].z, x[u][0].w + y[u][0].w,
                              x[u][1].x + y[u][1].x, x[u][1].y + y[u][1].y, x[u][1].z + y[u][1].z, x[u][1].w + y[u][1].w};
          uint4 hi, lo;
          tc::split8(v, hi, lo);
          const bool isu = idx < nu;
          const int i2 = isu ? idx : idx - nu;
          uint8_t* d = isu ? sU + (i2 & 7) * CSU + (i2 >> 3) * 16 : sD + (i2 & 7) * CSD + (i2 >> 3) * 16;
          *reinterpret_cast<uint4*>(d) = hi;
          if (NTERMS > 1) *reinterpret_cast<uint4*>(d + (isu ? PSU : PSD)) = lo;
        }
      }
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (warp == 0) {
      const uint32_t sU_u = tc::smem_u32(sU), sD_u = tc::smem_u32(sD);
      const int ksteps = T >> 4;
      const uint32_t acc0 = first ? 0u : 1u;
      if (tc::elect_one()) {
        for (int j = 0; j < WG_TAPS; ++j) {
          const uint32_t d = tmem + (uint32_t)(j >> 1) * 64u + ((uint32_t)((j & 1) * 16) << 16);
          // A = dy, B = u (both MN-major, K = time): hi*hi, hi*lo, lo*hi
          const uint64_t a_hi = tc::smem_desc(sD_u, 128, CSD), a_lo = tc::smem_desc(sD_u + PSD, 128, CSD);
          const uint64_t b_hi = tc::smem_desc(sU_u + (uint32_t)j * 16u, 128, CSU), b_lo = tc::smem_desc(sU_u + PSU + (uint32_t)j * 16u, 128, CSU);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t dk = (uint64_t)(ks * 16);                     // 16 time rows = 256 bytes = 16 address units
            tc::mma_bf16(d, a_hi + dk, b_hi + dk, idesc, acc0 | (uint32_t)(ks != 0));
            if (NTERMS > 1) {
              tc::mma_bf16(d, a_hi + dk, b_lo + dk, idesc, 1);
              tc::mma_bf16(d, a_lo + dk, b_hi + dk, idesc, 1);
            }
          }
        }
        tc::tc_commit(bar);
      }
      __syncwarp();
    }
    tc::mbar_wait(bar, phase);   // all MMAs that read this sample's tiles are done
    tc::tc_fence_after();
    phase ^= 1;
    first = false;
    __syncthreads();
  }
  // ---- epilogue: 16 accumulators -> partial[grp][tap][co][ci] ----
  if (warp >= 4 && !first) {
    const int q = warp - 4;
    const int co = q * 16 + (lane & 15);
    for (int cb = 0; cb < 8; ++cb) {
      const int tap = k0 + cb * 2 + (lane >> 4);
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + cb * 64;
      float* o = a.partial + (((long)grp * TAPS + tap) * CH + co) * CH;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[32];
        tc::tmem_ld32(taddr + half * 32, v);
#pragma unroll
        for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(o + half * 32 + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem, 512);
}

// dW[co][ci][k] = sum_g partial[g][k][co][ci]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, int groups, int TAPS) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // over k*4096 + co*64 + ci
  if (i >= TAPS * CH * CH) return;
  float s = 0.f;
  for (int g = 0; g < groups; ++g) s += partial[(long)g * TAPS * CH * CH + i];
  int ci = i & 63, co = (i >> 6) & 63, k = i >> 12;
  dW[((long)co * CH + ci) * TAPS + k] = s;
}

}  // namespace convtc

// ------------------------------------------------------------------------------------------------
// Interface used by tower.cu
// ------------------------------------------------------------------------------------------------
inline bool conv_tc_supported(int Cin, int Cout, int taps, int T) {
  return Cin == 64 && Cout == 64 && taps >= 16 && taps <= convtc::MAX_TAPS && (taps % 16) == 0 && T >= 64 && (T % 64) == 0 && T <= 512;
}

// scratch: packed weights (1 MB) + weight-gradient partials (groups MB)
inline size_t conv_tc_scratch_bytes(int B, int T, int taps) {
  (void)T;
  if (taps > convtc::MAX_TAPS || (taps % 16)) return 0;
  int groups = B < convtc::WG_MAX_GROUPS ? B : convtc::WG_MAX_GROUPS;
  return (size_t)taps * convtc::W_TAP_BYTES + (size_t)groups * taps * 64 * 64 * sizeof(float) + 256;
}

template <int NTERMS>
inline int conv_tc_launch(const convtc::ConvTcArgs& a, int B, cudaStream_t st) {
  static bool configured = false;
  uint32_t smem = convtc::conv_smem_bytes(a.T, a.taps);
  if (!configured) {
    if (cudaFuncSetAttribute(convtc::conv64_tc_kernel<NTERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return EEGCLIP_ERR_CUDA;
    configured = true;
  }
  ProfScope prof(PROF_CONV_TC, st);
  convtc::conv64_tc_kernel<NTERMS><<<B, 256, smem, st>>>(a);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

inline int conv_tc_forward(int math, const float* xin, const float* skip_in, const float* w, const float* bias, float* y, int B, int T,
                           int taps, int PL, const Drop& drop, void* scratch, cudaStream_t st) {
  uint8_t* wp = (uint8_t*)scratch;
  convtc::pack_conv_weights_kernel<<<(taps * 8 * 64 + 255) / 256, 256, 0, st>>>(w, wp, 0, taps);
  LAUNCH_CHECK();
  convtc::ConvTcArgs a;
  a.src = xin; a.skip = skip_in; a.wpacked = wp; a.bias = bias; a.out = y; a.T = T; a.src_rows = T; a.row_off = PL; a.taps = taps; a.drop = drop; a.dbg = g_dbg_buf; a.exp_mode = g_tune[5];
  return math == EEGCLIP_MATH_BF16 ? conv_tc_launch<1>(a, B, st) : conv_tc_launch<3>(a, B, st);
}

template <int NTERMS>
inline int wgrad_tc_launch(const convtc::WgradTcArgs& a, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(convtc::wgrad64_tc_kernel<NTERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return EEGCLIP_ERR_CUDA;
    configured = true;
  }
  dim3 grid(a.taps / convtc::WG_TAPS, a.groups);
  ProfScope prof(PROF_WGRAD_TC, st);
  convtc::wgrad64_tc_kernel<NTERMS><<<grid, 256, convtc::wgrad_smem_bytes(a.T), st>>>(a);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// du = dgrad(dypad, w) ; dw = wgrad(dypad, xin + skip_in)
inline int conv_tc_backward(int math, const float* xin, const float* skip_in, const float* w, const float* dypad, int taps, int PLb,
                            float* du, float* dw, int B, int T, void* scratch, cudaStream_t st) {
  uint8_t* wp = (uint8_t*)scratch;
  float* partial = (float*)(wp + (size_t)taps * convtc::W_TAP_BYTES);
  const int TP = T + taps - 1;
  convtc::pack_conv_weights_kernel<<<(taps * 8 * 64 + 255) / 256, 256, 0, st>>>(w, wp, 1, taps);
  LAUNCH_CHECK();
  convtc::ConvTcArgs a;
  a.src = dypad; a.skip = nullptr; a.wpacked = wp; a.bias = nullptr; a.out = du; a.T = T; a.src_rows = TP; a.row_off = 0; a.taps = taps; a.dbg = g_dbg_buf; a.exp_mode = g_tune[5];
  a.drop = make_drop(0, 0, 0, 0.f, 0);
  int rc = math == EEGCLIP_MATH_BF16 ? conv_tc_launch<1>(a, B, st) : conv_tc_launch<3>(a, B, st);
  if (rc != EEGCLIP_OK) return rc;
  convtc::WgradTcArgs g;
  g.xin = xin; g.skip = skip_in; g.dypad = dypad; g.partial = partial; g.B = B; g.T = T; g.PL = taps - 1 - PLb; g.PLb = PLb; g.taps = taps;
  g.groups = B < convtc::WG_MAX_GROUPS ? B : convtc::WG_MAX_GROUPS;
  rc = math == EEGCLIP_MATH_BF16 ? wgrad_tc_launch<1>(g, st) : wgrad_tc_launch<3>(g, st);
  if (rc != EEGCLIP_OK) return rc;
  convtc::wgrad_reduce_kernel<<<(taps * 64 * 64 + 255) / 256, 256, 0, st>>>(partial, dw, g.groups, taps);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

}  // namespace eegclip
