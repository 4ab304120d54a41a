// tcgen05 implicit-GEMM for Conv1d(Cin -> Cout, k = 16..64, 'same') on time-major activations: forward, data gradient
// and weight gradient (clip_model.py:237,245 -- 74 % of the EEG tower's FLOPs; vlaai.py:27-33,55 -- the
// 64..256-channel stacks of VLAAI).  Channels are processed in blocks of 64: an output block is a CTA (grid.y), input
// blocks are an outer loop that re-stages the activation tile and keeps accumulating into the same TMEM columns.
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

// ------------------------------------------------------------------------------------------------
// Weight packing: fp32 W[co][ci][k]  ->  bf16, per tap [k-chunk][n' = 128][8] with n' < 64 the hi part of output n' and
// n' >= 64 the lo part of output n' - 64: ONE N=128 MMA computes A_hi.W_hi and A_hi.W_lo (the activation tile, the larger
// operand, is read once for both terms -- the kernel is shared-memory-bandwidth bound at N = 64), a second N=64 MMA over the
// first 64 rows of the same operand adds A_lo.W_hi.
//   mode 0 (forward) : n = co, contraction index = ci, tap = k
//   mode 1 (dgrad)   : n = ci, contraction index = co, tap k' holds W[..][..][63-k']
// ------------------------------------------------------------------------------------------------
__global__ void pack_conv_weights_kernel(const float* __restrict__ W, uint8_t* __restrict__ out, int mode, int TAPS, int Cin, int Cout) {
  pdl_sync();
  // packed block (nb, kb) = 64 outputs x 64 contraction channels, all taps: out + ((nb * nkb + kb) * TAPS + tap) * W_TAP_BYTES
  const int nN = (mode == 0 ? Cout : Cin) / CH, nK = (mode == 0 ? Cin : Cout) / CH;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // over blocks * taps * 8 chunks * 64 n
  if (i >= (long)nN * nK * TAPS * 8 * CH) return;
  const int n = i & 63, ch = (i >> 6) & 7;
  const long rest = i >> 9;
  const int tap = rest % TAPS;
  const int blk = rest / TAPS, kb = blk % nK, nb = blk / nK;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int kk = kb * CH + ch * 8 + e, nn = nb * CH + n;  // contraction / output channel
    v[e] = mode == 0 ? W[((long)nn * Cin + kk) * TAPS + tap] : W[((long)kk * Cin + nn) * TAPS + (TAPS - 1 - tap)];
  }
  uint4 hi, lo;
  tc::split8(v, hi, lo);
  uint8_t* base = out + ((long)blk * TAPS + tap) * W_TAP_BYTES + ch * (2 * CH * 16) + n * 16;
  *reinterpret_cast<uint4*>(base) = hi;
  *reinterpret_cast<uint4*>(base + CH * 16) = lo;
}

struct ConvTcArgs {
  const float* src;       // (B, src_rows, src_ld) fp32
  const float* skip;      // optional addend, same shape as src
  const uint8_t* wpacked; // [n block][k block][tap] W_TAP_BYTES
  const float* bias;      // optional (out_ld)
  float* out;             // (B, T, out_ld)
  int src_ld, out_ld;     // channels of src / out (multiples of 64); grid.y = out_ld / 64, the kernel loops over src_ld / 64
  int T;                  // output rows
  int src_rows;           // valid source rows
  int row_off;            // smem row r holds src row r - row_off (zero outside)
  int taps;               // kernel size
  Drop drop;
  unsigned long long* dbg; // development timeline (CTA 0): start, staged, mma-done, end (globaltimer ns)
};

__host__ __device__ inline uint32_t conv_smem_bytes(int T, int taps) {
  uint32_t TP = T + taps - 1;
  return 2u * 8u * TP * 16u + NSTAGE * W_TAP_BYTES + 128;
}

template <int NTERMS>
__global__ void __launch_bounds__(256, 1) conv64_tc_kernel(const ConvTcArgs a) {
  pdl_trigger();
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x, nb = blockIdx.y;
  const int nkb = a.src_ld / CH;
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
  const bool last64 = (T & 127) != 0;       // T % 128 == 64 -> last tile has M = 64
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
  uint32_t tmem = 0;
  pdl_wait();   // global memory is read from here on

  for (int kb = 0; kb < nkb; ++kb) {
  // ---- stage the activation tile: fp32 (+skip) -> bf16 hi/lo, chunk-major, zero padded ----
  // Loads are issued STAGE_U items deep before the first conversion: the tile is 82-164 KB per CTA and a load-convert-store
  // loop exposes one DRAM round trip per iteration (ncu: long_scoreboard was the top stall of this kernel).
  {
    const int ld = a.src_ld;
    const float* sb = a.src + (long)b * a.src_rows * ld + kb * CH;
    const float* kp = a.skip ? a.skip + (long)b * a.src_rows * ld + kb * CH : nullptr;
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
          const float4* p = reinterpret_cast<const float4*>(sb + (long)t * ld + ch * 8);
          x[u][0] = __ldg(p); x[u][1] = __ldg(p + 1);
          if (kp) {
            const float4* q = reinterpret_cast<const float4*>(kp + (long)t * ld + ch * 8);
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
  tmem = *tmem_slot;
  if (kb == 0) stamp(1);
  const uint8_t* wblk = a.wpacked + ((long)nb * nkb + kb) * TAPS * W_TAP_BYTES;
  const int g0 = kb * TAPS;                 // global tap counter of this block's first tap (ring slot / phase bookkeeping)

  if (warp == 0 && lane == 0) {
    // ===== weight producer: one 1-D bulk copy per tap into the ring =====
    const uint32_t bytes = W_TAP_BYTES;
    for (int tap = 0; tap < TAPS; ++tap) {
      const int s = (g0 + tap) % NSTAGE;
      const uint32_t ph = ((g0 + tap) / NSTAGE) & 1;
      tc::mbar_wait(&empty[s], ph ^ 1);
      tc::mbar_expect_tx(&full[s], bytes);
      tc::bulk_g2s(sB + s * W_TAP_BYTES, wblk + (long)tap * W_TAP_BYTES, bytes, &full[s]);
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop (descriptors stay in uniform registers), one elected lane issues =====
    const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB);
    // idesc[m64][wide]: wide = N 128 (hi.hi | hi.lo in one MMA), narrow = N 64
    const uint32_t idw128 = tc::idesc_bf16(128, 2 * CH, 0, 0), idw64 = tc::idesc_bf16(64, 2 * CH, 0, 0);
    const uint32_t idn128 = tc::idesc_bf16(128, CH, 0, 0), idn64 = tc::idesc_bf16(64, CH, 0, 0);
    for (int tap = 0; tap < TAPS; ++tap) {
      const int s = (g0 + tap) % NSTAGE;
      const uint32_t ph = ((g0 + tap) / NSTAGE) & 1;
      const uint32_t acc = (uint32_t)((kb | tap) != 0);
      tc::mbar_wait(&full[s], ph);
      tc::tc_fence_after();
      const uint32_t wb = sB_u + s * W_TAP_BYTES;
      const uint64_t b_w = tc::smem_desc(wb, 2 * CH * 16, 128);   // chunk stride 2048 B: rows 0-63 hi, 64-127 lo
      if (tc::elect_one()) {
        // Tile-major issue order.  Measured alternatives (tools/conv_timeline.py, 64 taps, T = 320, per window): K-step-major
        // order (consecutive MMAs on different accumulators) 67.7 us vs 62.7 us here; operand starts aligned to 128 B instead
        // of tap * 16 B: no change; no weight re-fetch: -2 us; without the lo.hi MMAs: -6 us.  The pipe is paced per
        // instruction at ~21 + 0.63 * N cycles (M = 128, K = 16) in both the SS form used here and the TS form below.
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
              tc::mma_bf16(d, a_hi + da, b_w + db, idw, acc | (uint32_t)(ks != 0));   // [hi.hi | hi.lo]
              tc::mma_bf16(d, a_lo + da, b_w + db, idn, 1);                            // += lo.hi (first 64 columns)
            } else {
              tc::mma_bf16(d, a_hi + da, b_w + db, idn, acc | (uint32_t)(ks != 0));
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
  // every MMA that reads this input block's tile has completed (also: the accumulators are final after the last block)
  tc::mbar_wait(accfull, (uint32_t)(kb & 1));
  tc::tc_fence_after();
  }  // kb
  {
    // ===== epilogue (all 8 warps): TMEM -> registers -> bias / dropout -> global.  Warp w reads TMEM lane quarter w % 4;
    // warps 4-7 take output channels 0-31, warps 0-3 (done with their producer / MMA roles by now) channels 32-63 =====
    const int q = warp & 3;
    const int half = warp < 4 ? 1 : 0;
    stamp(2);
    float4 bb[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bb[c] = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + nb * CH + half * 32) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
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
        float* o = a.out + ((long)b * T + row) * a.out_ld + nb * CH + half * 32;
        const uint64_t didx = ((uint64_t)b * T + row) * a.out_ld + nb * CH + half * 32;
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
  const float* xin;     // (B,T,Cin) conv input
  const float* skip;    // optional addend
  const float* dypad;   // (B,TP,Cout) zero-padded output gradient, valid rows [PLb, PLb+T)
  float* partial;       // [groups][taps][Cout][Cin]
  int B, T, PL, PLb, groups, taps;
  int Cin, Cout;        // multiples of 64; grid.x = (taps / 16) * (Cout / 64) * (Cin / 64)
};


// ------------------------------------------------------------------------------------------------
// Weight gradient, double-buffered over time slices.  (Round 1's first version staged a whole sample -- u and dy tiles, 168 KB at
// T = 320 -- issued its MMAs, waited, staged the next one: ~3.5 us of staging exposed per 24 us of MMAs, 2.07 ms per step against
// 1.64 ms for this kernel; that code is gone, the measurement is in DESIGN.md.)  Here the
// contraction over time is cut into S slices of TH = T / S rows (S = 2, 4 above T = 384): two slice buffers (86 KB each at
// T = 320), seven producer warps stage slice i+1 while warp 0 issues the MMAs of slice i (mbarrier full / empty per buffer).
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline int wgrad_db_slices(int T) { return T <= 384 ? 2 : 4; }
__host__ __device__ inline uint32_t wgrad_db_smem_bytes(int T) {
  const uint32_t TH = (uint32_t)T / wgrad_db_slices(T);
  return 2u * (2u * 8u * (TH + WG_TAPS - 1) * 16u + 2u * 8u * TH * 16u) + 128;
}

template <int NTERMS>
__global__ void __launch_bounds__(256, 1) wgrad64_db_kernel(const WgradTcArgs a) {
  pdl_trigger();
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntg = a.taps / WG_TAPS, nci = a.Cin / CH;
  const int tg = blockIdx.x % ntg, cib = (blockIdx.x / ntg) % nci, cob = blockIdx.x / (ntg * nci), grp = blockIdx.y;
  const int T = a.T, TAPS = a.taps;
  const int S = wgrad_db_slices(T), TH = T / S, TUH = TH + WG_TAPS - 1;
  const int k0 = tg * WG_TAPS;
  const uint32_t CSU = (uint32_t)TUH * 16u, PSU = 8u * CSU;
  const uint32_t CSD = (uint32_t)TH * 16u, PSD = 8u * CSD;
  const uint32_t BUF = 2u * PSU + 2u * PSD;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2u * BUF);
  uint64_t* full = bars;        // [2] producers -> MMA (7 warp arrivals)
  uint64_t* empty = bars + 2;   // [2] MMA -> producers (tcgen05.commit)
  uint64_t* accfull = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&full[i], 7); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(accfull, 1);
    tc::mbar_fence_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();   // global memory is read from here on
  const int nsamp = grp < a.B ? (a.B - grp + a.groups - 1) / a.groups : 0;
  const int nitems = nsamp * S;                      // item i = (sample grp + (i / S) * groups, time slice i % S)

  if (warp == 0) {
    // ===== MMA issuer =====
    const uint32_t idesc = tc::idesc_bf16(64, CH, 1, 1);   // both operands MN-major (K = time)
    const int ksteps = TH >> 4;
    for (int i = 0; i < nitems; ++i) {
      const int s = i & 1;
      tc::mbar_wait(&full[s], (i >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t sU_u = tc::smem_u32(smem + s * BUF), sD_u = sU_u + 2u * PSU;
      if (tc::elect_one()) {
        for (int j = 0; j < WG_TAPS; ++j) {
          const uint32_t d = tmem + (uint32_t)(j >> 1) * 64u + ((uint32_t)((j & 1) * 16) << 16);
          // A = dy, B = u (both MN-major, K = time): hi*hi, hi*lo, lo*hi
          const uint64_t a_hi = tc::smem_desc(sD_u, 128, CSD), a_lo = tc::smem_desc(sD_u + PSD, 128, CSD);
          const uint64_t b_hi = tc::smem_desc(sU_u + (uint32_t)j * 16u, 128, CSU), b_lo = tc::smem_desc(sU_u + PSU + (uint32_t)j * 16u, 128, CSU);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t dk = (uint64_t)(ks * 16);                     // 16 time rows = 256 bytes = 16 address units
            tc::mma_bf16(d, a_hi + dk, b_hi + dk, idesc, (uint32_t)((i | ks) != 0));
            if (NTERMS > 1) {
              tc::mma_bf16(d, a_hi + dk, b_lo + dk, idesc, 1);
              tc::mma_bf16(d, a_lo + dk, b_hi + dk, idesc, 1);
            }
          }
        }
        tc::tc_commit(&empty[s]);
        if (i == nitems - 1) tc::tc_commit(accfull);
      }
      __syncwarp();
    }
  } else {
    // ===== producers (warps 1-7): stage u rows [t0 + k0 - PL, + TUH) and dy rows [t0, t0 + TH) of the item's sample =====
    const int ptid = tid - 32;
    const int ldx = a.Cin, ldy = a.Cout;
    constexpr int WG_U = 6, NPT = 224;
    const int nu = TUH * 8, total = nu + TH * 8;     // work units [0, nu): u tile; [nu, total): dy tile
    for (int i = 0; i < nitems; ++i) {
      const int s = i & 1;
      const int b = grp + (i / S) * a.groups, t0 = (i % S) * TH;
      const float* xb = a.xin + (long)b * T * ldx + cib * CH;
      const float* kb = a.skip ? a.skip + (long)b * T * ldx + cib * CH : nullptr;
      const float* db = a.dypad + ((long)b * (T + TAPS - 1) + a.PLb + t0) * ldy + cob * CH;
      uint8_t* sU = smem + s * BUF;
      uint8_t* sD = sU + 2u * PSU;
      tc::mbar_wait(&empty[s], ((i >> 1) & 1) ^ 1);
      for (int base = 0; base < total; base += NPT * WG_U) {
        float4 x[WG_U][2], y[WG_U][2];
#pragma unroll
        for (int u = 0; u < WG_U; ++u) {
          const int idx = base + u * NPT + ptid;
          x[u][0] = x[u][1] = y[u][0] = y[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (idx < nu) {
            const int r = idx >> 3, ch = idx & 7;
            const int t = t0 + r + k0 - a.PL;
            if (t >= 0 && t < T) {
              const float4* p = reinterpret_cast<const float4*>(xb + (long)t * ldx + ch * 8);
              x[u][0] = __ldg(p); x[u][1] = __ldg(p + 1);
              if (kb) {
                const float4* q = reinterpret_cast<const float4*>(kb + (long)t * ldx + ch * 8);
                y[u][0] = __ldg(q); y[u][1] = __ldg(q + 1);
              }
            }
          } else if (idx < total) {
            const int i2 = idx - nu;
            const float4* p = reinterpret_cast<const float4*>(db + (long)(i2 >> 3) * ldy + (i2 & 7) * 8);
            x[u][0] = __ldg(p); x[u][1] = __ldg(p + 1);
          }
        }
#pragma unroll
        for (int u = 0; u < WG_U; ++u) {
          const int idx = base + u * NPT + ptid;
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
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full[s]);
    }
  }
  __syncthreads();
  // ---- epilogue: 16 accumulators -> partial[grp][tap][co][ci] ----
  if (warp >= 4 && nitems > 0) {
    tc::mbar_wait(accfull, 0);
    tc::tc_fence_after();
    const int q = warp - 4;
    const int co = q * 16 + (lane & 15);
    for (int cb = 0; cb < 8; ++cb) {
      const int tap = k0 + cb * 2 + (lane >> 4);
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + cb * 64;
      float* o = a.partial + (((long)grp * TAPS + tap) * a.Cout + cob * CH + co) * a.Cin + cib * CH;
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

// dW[co][ci][k] = sum_g partial[g][k][co][ci]   (fixed order: deterministic)
// grid (TAPS / 16, Cin / 32, Cout), block 256: a 16-tap x 32-channel tile is summed over the groups with coalesced reads (ci
// fastest) and transposed through shared memory so that the (co, ci, k) weight-gradient layout is written with k fastest.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, int groups, int TAPS,
                                                          int Cin, int Cout) {
  pdl_sync();
  __shared__ float tile[16][33];
  const int k0 = blockIdx.x * 16, ci0 = blockIdx.y * 32, co = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // ty: 0..7
  const long per = (long)TAPS * Cout * Cin;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int k = k0 + ty + 8 * h;
    const float* p = partial + ((long)k * Cout + co) * Cin + ci0 + tx;
    float s = 0.f;
#pragma unroll 4
    for (int g = 0; g < groups; ++g) s += __ldg(p + (long)g * per);
    tile[ty + 8 * h][tx] = s;
  }
  __syncthreads();
  // 512 outputs of the tile: thread -> (ci = t >> 4, k = t & 15), two passes
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int t = threadIdx.x + 256 * h;
    const int ci = t >> 4, kk = t & 15;
    dW[((long)co * Cin + ci0 + ci) * TAPS + k0 + kk] = tile[kk][ci];
  }
}

}  // namespace convtc

// ------------------------------------------------------------------------------------------------
// Interface used by tower.cu
// ------------------------------------------------------------------------------------------------
inline bool conv_tc_supported(int Cin, int Cout, int taps, int T) {
  return Cin >= 64 && Cout >= 64 && (Cin % 64) == 0 && (Cout % 64) == 0 && Cin <= 1024 && Cout <= 512 && taps >= 16 &&
         taps <= convtc::MAX_TAPS && (taps % 16) == 0 && T >= 64 && (T % 64) == 0 && T <= 512;
}
// narrow outputs (SpeechSmallConv: 1024 -> 8 channels, clip_model.py:204-232) run on the same kernels with the output channels
// zero-padded to one 64-channel block (tower.cu::conv_block_fwd / _bwd)
inline bool conv_tc_padded_ok(int Cin, int Cout, int taps, int T) {
  return Cout < 64 && (Cout % 4) == 0 && conv_tc_supported(Cin, 64, taps, T);
}

// weight-gradient sample groups: (tap groups x channel blocks) x groups CTAs ~ one wave of 148 SMs
inline int conv_tc_wgrad_groups(int B, int taps, int Cin, int Cout) {
  const int gx = (taps / convtc::WG_TAPS) * (Cin / 64) * (Cout / 64);
  int g = 148 / (gx > 0 ? gx : 1);
  if (g < 1) g = 1;
  return B < g ? B : g;
}

// scratch: packed weights (16 KB per tap and 64x64 channel block) + weight-gradient partials (groups x |W|)
inline size_t conv_tc_scratch_bytes(int B, int T, int taps, int Cin, int Cout) {
  (void)T;
  if (taps > convtc::MAX_TAPS || (taps % 16) || (Cin % 64) || (Cout % 64)) return 0;
  const size_t blocks = (size_t)(Cin / 64) * (Cout / 64);
  return align_up(blocks * taps * convtc::W_TAP_BYTES, 256) +
         (size_t)conv_tc_wgrad_groups(B, taps, Cin, Cout) * taps * Cin * Cout * sizeof(float) + 256;
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
  LAUNCH_PDL((convtc::conv64_tc_kernel<NTERMS>), dim3(B, a.out_ld / convtc::CH), 256, smem, st, a);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

inline int conv_tc_pack(const float* w, uint8_t* wp, int mode, int taps, int Cin, int Cout, int T, cudaStream_t st) {
  const long n = (long)(Cin / 64) * (Cout / 64) * taps * 8 * 64;
  (void)T;
  LAUNCH_PDL((convtc::pack_conv_weights_kernel), (unsigned)((n + 255) / 256), 256, 0, st, w, wp, mode, taps, Cin, Cout);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

inline int conv_tc_forward(int math, const float* xin, const float* skip_in, const float* w, const float* bias, float* y, int B, int T,
                           int Cin, int Cout, int taps, int PL, const Drop& drop, void* scratch, cudaStream_t st) {
  uint8_t* wp = (uint8_t*)scratch;
  { int prc = conv_tc_pack(w, wp, 0, taps, Cin, Cout, T, st); if (prc != EEGCLIP_OK) return prc; }
  convtc::ConvTcArgs a;
  a.src = xin; a.skip = skip_in; a.wpacked = wp; a.bias = bias; a.out = y; a.src_ld = Cin; a.out_ld = Cout;
  a.T = T; a.src_rows = T; a.row_off = PL; a.taps = taps; a.drop = drop; a.dbg = g_dbg_buf;
  return math == EEGCLIP_MATH_BF16 ? conv_tc_launch<1>(a, B, st) : conv_tc_launch<3>(a, B, st);
}

template <int NTERMS>
inline int wgrad_tc_launch(const convtc::WgradTcArgs& a, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(convtc::wgrad64_db_kernel<NTERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return EEGCLIP_ERR_CUDA;
    configured = true;
  }
  dim3 grid((a.taps / convtc::WG_TAPS) * (a.Cin / convtc::CH) * (a.Cout / convtc::CH), a.groups);
  ProfScope prof(PROF_WGRAD_TC, st);
  LAUNCH_PDL((convtc::wgrad64_db_kernel<NTERMS>), grid, 256, convtc::wgrad_db_smem_bytes(a.T), st, a);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// du = dgrad(dypad, w) ; dw = wgrad(dypad, xin + skip_in)
inline int conv_tc_backward(int math, const float* xin, const float* skip_in, const float* w, const float* dypad, int Cin, int Cout,
                            int taps, int PLb, float* du, float* dw, int B, int T, void* scratch, cudaStream_t st) {
  uint8_t* wp = (uint8_t*)scratch;
  const size_t blocks = (size_t)(Cin / 64) * (Cout / 64);
  float* partial = (float*)(wp + align_up(blocks * taps * convtc::W_TAP_BYTES, 256));
  const int TP = T + taps - 1;
  { int prc = conv_tc_pack(w, wp, 1, taps, Cin, Cout, T, st); if (prc != EEGCLIP_OK) return prc; }
  convtc::ConvTcArgs a;
  a.src = dypad; a.skip = nullptr; a.wpacked = wp; a.bias = nullptr; a.out = du; a.src_ld = Cout; a.out_ld = Cin;
  a.T = T; a.src_rows = TP; a.row_off = 0; a.taps = taps; a.dbg = g_dbg_buf;
  a.drop = make_drop(0, 0, 0, 0.f, 0);
  int rc = math == EEGCLIP_MATH_BF16 ? conv_tc_launch<1>(a, B, st) : conv_tc_launch<3>(a, B, st);
  if (rc != EEGCLIP_OK) return rc;
  convtc::WgradTcArgs g;
  g.xin = xin; g.skip = skip_in; g.dypad = dypad; g.partial = partial; g.B = B; g.T = T; g.PL = taps - 1 - PLb; g.PLb = PLb; g.taps = taps;
  g.Cin = Cin; g.Cout = Cout;
  g.groups = conv_tc_wgrad_groups(B, taps, Cin, Cout);
  rc = math == EEGCLIP_MATH_BF16 ? wgrad_tc_launch<1>(g, st) : wgrad_tc_launch<3>(g, st);
  if (rc != EEGCLIP_OK) return rc;
  LAUNCH_PDL((convtc::wgrad_reduce_kernel), dim3(taps / 16, Cin / 32, Cout), 256, 0, st, partial, dw, g.groups, taps, Cin, Cout);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

}  // namespace eegclip
