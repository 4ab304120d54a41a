// tcgen05 similarity kernels of the contrastive head (clip_model.py:675-693, 913-930): logits = S.E^T * exp(tau) are
// produced tile by tile in TMEM and consumed in the epilogue -- the B x B logit matrix never exists in HBM.
//
//   epack_kernel        normalised embeddings (R x D fp32) -> bf16 hi/lo operand blocks, one 32 KB block per
//                       (128-row block, 64-column chunk):  [plane][c8][row][8]   (tc_common.cuh chunk-major layout)
//                       so that a GEMM stage is ONE contiguous bulk async copy (TMA engine) per operand block.
//   logits_tc_kernel    CTA (m block of 128 rows) x (n tile of 256 columns): TMA producer thread + MMA thread (K loop over
//                       D in 64-column chunks, 2-stage ring) + 8 epilogue warps (one thread per row and 128-column half;
//                       with 4 the exp / pack work of MODE 2 was 32 us of latency-bound epilogue per 49 us of MMAs):
//       MODE 1  per-row (max, sum exp) partial of the tile + the diagonal logit        (forward: row / column LSE)
//       MODE 2  G = (exp(L - lse_m) + exp(L - lse_n) - 2 delta) / (2B) * upstream, written straight into the packed bf16 hi/lo
//               A-operand layout of the next GEMM (rows m, contraction index n) + sum G.L for d tau   (backward: CE gradient)
//       MODE 3  dot products (x exp(tau) when tau is given) stored row-major: match-mismatch bank scoring
//               (train_clip_helper_functions.py:182) and the two products of the backward
//   The products of the backward, dS = exp(tau) G.E and dE = exp(tau) G^T.S, are contractions over the batch index: the same
//   kernel in MODE 3 with A = packed G (K = batch) and B = the TRANSPOSED embeddings packed by epack_t_kernel (rows = features).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace eegclip {
namespace headtc {

constexpr int RB = 128;                        // rows per operand block
constexpr int KC = 64;                         // columns per operand block
constexpr int PLANE = 8 * RB * 16;             // 16 KB
constexpr int BLK = 2 * PLANE;                 // hi + lo
constexpr int NT = 256;                        // logit columns per CTA (two operand blocks)
constexpr int NSTAGE = 2;
constexpr int STAGE = 3 * BLK;                 // A block + 2 B blocks = 96 KB
constexpr float LOG2E = 1.4426950408889634f;

// Tile width.  NTC = 256: two whole 128-row operand blocks of the column operand per stage (large batches).  NTC = 64: small
// batches -- at B = 256 the 256-wide form is 2 CTAs on 148 SMs running 480 N=128 MMAs per 64-feature chunk each (the in-step head
// was 5 launches of ~50 us on two SMs); 64-column tiles are 4 x as many CTAs with a quarter of the MMA work and a 48 KB stage
// (four stages in flight).  The 64-row sub-block of an operand block is eight 1 KB pieces per plane (one per 8-feature chunk).
template <int NTC>
struct Tile {
  static constexpr bool SUB = NTC < RB;                                  // column operand is a sub-block of one operand block
  static constexpr int NBLK = SUB ? 1 : NTC / RB;
  static constexpr uint32_t BPLANE = SUB ? 8u * NTC * 16u : (uint32_t)PLANE;
  static constexpr uint32_t BBYTES = SUB ? 2u * BPLANE : (uint32_t)(NBLK * BLK);
  static constexpr uint32_t STAGE_B = (uint32_t)BLK + BBYTES;
  static constexpr int NSTG = SUB ? 4 : 2;
  static constexpr int HALF = NTC / 2 < 32 ? 32 : NTC / 2;               // columns per epilogue warp set
};
// tiles of the 256-wide form below which the 64-wide form is launched (its CTA count is 4 x that)
constexpr int SMALL_TILES = 37;
inline int logits_ntc(int mode, int M, int N) {
  const int t256 = ((N + NT - 1) / NT) * ((M + RB - 1) / RB);
  return (mode != 3 && t256 <= SMALL_TILES && g_tune[14] == 0) ? 64 : NT;   // (g_tune[14] = 1: always 256-wide, A/B timing)
}

inline size_t epack_bytes(int R, int D) { return (size_t)((R + RB - 1) / RB) * (D / KC) * BLK; }

// grid: (D/64 chunks, row blocks); block 256: lane -> (row = lane & 7, c8 = lane >> 3 (+4 per half)) as in lin_tc stage_chunk
__global__ void __launch_bounds__(256) epack_kernel(const float* __restrict__ X, uint8_t* __restrict__ P, int R, int D) {
  pdl_sync();
  const int kc = blockIdx.x, rb = blockIdx.y, nkc = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* blk = P + ((size_t)rb * nkc + kc) * BLK;
  // 16 row groups of 8 x 2 chunk halves = 32 combos over 8 warps
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int combo = it * 8 + warp;
    const int rgrp = combo >> 1, half = combo & 1;
    const int r = rgrp * 8 + (lane & 7), c8 = half * 4 + (lane >> 3);
    const long row = (long)rb * RB + r;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (row < R) {
      const float4* p = reinterpret_cast<const float4*>(X + row * D + kc * KC + c8 * 8);
      const float4 a = __ldg(p), b = __ldg(p + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    uint8_t* d = blk + (c8 * RB + r) * 16;
    *reinterpret_cast<uint4*>(d) = hi;
    *reinterpret_cast<uint4*>(d + PLANE) = lo;
  }
}

// Transposed pack: X (R x D fp32) -> operand blocks of X^T (D rows x R contraction columns), same block layout as epack_kernel.
// grid (ceil(R/64) chunks, ceil(D/128) row blocks); the 64 x 128 fp32 tile goes through shared memory (coalesced reads along D,
// conflict-free column reads).
__global__ void __launch_bounds__(256) epack_t_kernel(const float* __restrict__ X, uint8_t* __restrict__ P, int R, int D) {
  pdl_sync();
  __shared__ float tile[KC][RB + 1];
  const int kc = blockIdx.x, rb = blockIdx.y, nkc = gridDim.x;
  const int r0 = kc * KC, d0 = rb * RB;
  for (int i = threadIdx.x; i < KC * (RB / 4); i += 256) {
    const int r = i / (RB / 4), d4 = (i % (RB / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < R && d0 + d4 < D) v = __ldg(reinterpret_cast<const float4*>(X + (long)(r0 + r) * D + d0 + d4));   // D % 4 == 0
    tile[r][d4] = v.x; tile[r][d4 + 1] = v.y; tile[r][d4 + 2] = v.z; tile[r][d4 + 3] = v.w;
  }
  __syncthreads();
  uint8_t* blk = P + ((size_t)rb * nkc + kc) * BLK;
  for (int u = threadIdx.x; u < 8 * RB; u += 256) {
    const int c8 = u / RB, row = u % RB;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = tile[c8 * 8 + e][row];
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    uint8_t* d = blk + (c8 * RB + row) * 16;
    *reinterpret_cast<uint4*>(d) = hi;
    *reinterpret_cast<uint4*>(d + PLANE) = lo;
  }
}

struct LogitsArgs {
  const uint8_t* Ap;      // packed row operand (m), block index = (m0/128 + blockIdx.y) * nkc + kc
  const uint8_t* Bp;      // packed column operand (n)
  int a_blk0;             // first row block of A for blockIdx.y == 0 (local rows inside the gathered matrix)
  int M, N, D;            // valid rows of this call (m), valid columns (n), features
  const float* tau;       // device scalar: logits = dot * exp(tau)
  int m_off, n_off;       // global indices of m = 0 / n = 0 (diagonal where m + m_off == n + n_off)
  // MODE 1
  float2* part;           // [M][ntiles]
  float* diag;            // [M] (written by the tile that holds the diagonal)
  // MODE 2
  const float* lse_m;     // indexed by m + m_off
  const float* lse_n;     // indexed by n + n_off   (unused when one_sided)
  const float* up;        // device scalar: upstream gradient of the loss
  float inv_2b;
  int one_sided;          // 1: G = (exp(L - lse_m) - delta) / B ; 2: G = (exp(L - lse_n) - delta) / B ; 0: symmetric
  uint8_t* Gp;            // packed G: operand blocks [m / 128][n / 64], the A operand of the backward contractions over n
  int nkc_g;              // 64-column chunks per row block of Gp (= ceil(N / 64))
  float* dtau;            // += sum G * L (nullptr: skip) -- float atomics unless dtau_slots is given
  float* dtau_slots;      // [gridDim.x][8]: per-(CTA, epilogue warp) sums of G * L, folded in fixed order by dtau_fold (dtau then unused)
  // MODE 3
  float* out;             // [M][ldo] dot products (x exp(tau) when tau != nullptr)
  long ldo;
};

// Persistent: CTA c works on tiles c, c + gridDim.x, ... (n tile fastest, so neighbouring CTAs share the A row block in L2); the
// accumulator is double-buffered in TMEM (2 x 256 columns), so the epilogue of tile i overlaps the K loop of tile i + 1.
template <int MODE, int NTERMS, int NTC>
__global__ void __launch_bounds__(320, 1) logits_tc_kernel(const LogitsArgs a) {
  pdl_sync();
  using TL = Tile<NTC>;
  constexpr int NT = NTC;                       // (shadows the namespace constant: tile width of this instantiation)
  constexpr int NSTAGE = TL::NSTG;
  constexpr uint32_t STAGE = TL::STAGE_B;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE);
  uint64_t* full = bars;                  // [NSTAGE]  TMA -> MMA
  uint64_t* empty = bars + NSTAGE;        // [NSTAGE]  MMA -> TMA   (tcgen05.commit)
  uint64_t* accfull = bars + 2 * NSTAGE;  // [2]       MMA -> epilogue
  uint64_t* accempty = accfull + 2;       // [2]       epilogue -> MMA (8 warps x 32 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);
  const int nkc = a.D / KC;
  const int ntn = (a.N + NT - 1) / NT, ntm = (a.M + RB - 1) / RB;
  const int ntiles = ntn * ntm;
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&accfull[i], 1); tc::mbar_init(&accempty[i], 8 * 32); }
    tc::mbar_fence_init();
  }
  if (warp == 5) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ===== TMA producer: one bulk copy per operand block =====
    if (lane == 0) {
      uint32_t c = 0;                      // chunk counter across tiles: ring slot and phase
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tn = tile % ntn, tm = tile / ntn;
        const uint32_t one = NTERMS > 1 ? BLK : PLANE;
        const uint8_t* ab = a.Ap + (size_t)(a.a_blk0 + tm) * nkc * BLK;
        if (TL::SUB) {
          // rows [r0, r0 + NT) of operand block tn * NT / 128: per plane eight contiguous NT x 16 byte pieces (one per chunk)
          const uint8_t* bb0 = a.Bp + (size_t)((tn * NT) / RB) * nkc * BLK + (size_t)((tn * NT) % RB) * 16;
          constexpr uint32_t piece = (uint32_t)NT * 16u;
          const uint32_t bytes = one + (NTERMS > 1 ? 16u : 8u) * piece;
          for (int kc = 0; kc < nkc; ++kc, ++c) {
            const int s = c % NSTAGE;
            tc::mbar_wait(&empty[s], ((c / NSTAGE) & 1) ^ 1);
            tc::mbar_expect_tx(&full[s], bytes);
            uint8_t* st = smem + s * STAGE;
            tc::bulk_g2s(st, ab + (size_t)kc * BLK, one, &full[s]);
            const uint8_t* src = bb0 + (size_t)kc * BLK;
#pragma unroll
            for (int pl = 0; pl < (NTERMS > 1 ? 2 : 1); ++pl)
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8)
                tc::bulk_g2s(st + BLK + pl * TL::BPLANE + c8 * piece, src + pl * PLANE + c8 * (RB * 16), piece, &full[s]);
          }
        } else {
        const int nb_blocks = (a.N - tn * NT > RB) ? 2 : 1;            // valid column blocks of this tile
        const uint32_t bytes = (uint32_t)(1 + nb_blocks) * one;
        const uint8_t* bb0 = a.Bp + (size_t)(tn * 2) * nkc * BLK;
        for (int kc = 0; kc < nkc; ++kc, ++c) {
          const int s = c % NSTAGE;
          tc::mbar_wait(&empty[s], ((c / NSTAGE) & 1) ^ 1);
          tc::mbar_expect_tx(&full[s], bytes);
          uint8_t* st = smem + s * STAGE;
          tc::bulk_g2s(st, ab + (size_t)kc * BLK, one, &full[s]);
          for (int j = 0; j < nb_blocks; ++j) tc::bulk_g2s(st + (1 + j) * BLK, bb0 + ((size_t)j * nkc + kc) * BLK, one, &full[s]);
        }
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issue (whole warp runs the loop, one elected lane issues) =====
    const uint32_t base = tc::smem_u32(smem);
    constexpr int NCOL = TL::SUB ? NT : RB;                  // columns per MMA (one operand block, or the sub-block)
    const uint32_t idesc = tc::idesc_bf16(128, NCOL, 0, 0);
    uint32_t c = 0, t = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
      const int tn = tile % ntn;
      const int nb_blocks = TL::SUB ? 1 : ((a.N - tn * NT > RB) ? 2 : 1);
      const uint32_t buf = t & 1;
      tc::mbar_wait(&accempty[buf], ((t >> 1) & 1) ^ 1);
      tc::tc_fence_after();
      for (int kc = 0; kc < nkc; ++kc, ++c) {
        const int s = c % NSTAGE;
        tc::mbar_wait(&full[s], (c / NSTAGE) & 1);
        tc::tc_fence_after();
        const uint32_t st = base + s * STAGE;
        const uint64_t a_hi = tc::smem_desc(st, RB * 16, 128), a_lo = tc::smem_desc(st + PLANE, RB * 16, 128);
        if (tc::elect_one()) {
          for (int j = 0; j < nb_blocks; ++j) {
            const uint32_t bs = st + (1 + j) * BLK;
            const uint64_t b_hi = tc::smem_desc(bs, NCOL * 16, 128), b_lo = tc::smem_desc(bs + TL::BPLANE, NCOL * 16, 128);
            const uint32_t d = tmem + buf * NT + j * 128;
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks) {
              const uint64_t dk = (uint64_t)((2 * ks * RB * 16) >> 4);         // A: 16 features = two chunks of RB rows
              const uint64_t dkb = (uint64_t)((2 * ks * NCOL * 16) >> 4);      // B: two chunks of NCOL rows
              tc::mma_bf16(d, a_hi + dk, b_hi + dkb, idesc, (kc | ks) != 0);
              if (NTERMS > 1) {
                tc::mma_bf16(d, a_hi + dk, b_lo + dkb, idesc, 1);
                tc::mma_bf16(d, a_lo + dk, b_hi + dkb, idesc, 1);
              }
            }
          }
          tc::tc_commit(&empty[s]);
          if (kc == nkc - 1) tc::tc_commit(&accfull[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: thread = logit row m; warps 0-3 take columns [0,128) of the tile, warps 6-9 columns [128,256) =====
    const int q = warp & 3, chalf = warp >= 6 ? 1 : 0;      // TMEM lane quarter of a warp is warp % 4
    const float scale = a.tau ? __expf(*a.tau) : 1.f;
    float tsum_cta = 0.f;                                   // MODE 2 with dtau_slots: this thread's sum of G * L over all its tiles
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
    const int tn = tile % ntn, tm = tile / ntn;
    const int n0 = tn * NT, m0 = tm * RB;
    const uint32_t buf = t & 1;
    const uint32_t tacc = tmem + buf * NT + ((uint32_t)(q * 32) << 16);
    const int m = m0 + q * 32 + lane;
    const bool mv = m < a.M;
    tc::mbar_wait(&accfull[buf], (t >> 1) & 1);
    tc::tc_fence_after();
    const int ncols = min(NT, a.N - n0);
    const int c_lo = chalf * TL::HALF, c_hi = min(ncols, c_lo + TL::HALF);   // this warp's columns (empty when c_lo >= ncols)
    const int gm = m + a.m_off;
    if (MODE == 1) {
      const float sc2 = scale * LOG2E;                     // work in base 2
      float mx = -INFINITY, sum = 0.f;
      for (int cb = c_lo; cb < c_hi; cb += 32) {
        float v[32];
        tc::tmem_ld32(tacc + cb, v);
        float cm = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = (cb + j < ncols) ? v[j] * sc2 : -INFINITY;
          cm = fmaxf(cm, v[j]);
        }
        const float nm = fmaxf(mx, cm);
        float cs = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) cs += exp2f(v[j] - nm);
        sum = sum * exp2f(mx - nm) + cs;
        mx = nm;
        const int dj = gm - (n0 + a.n_off) - cb;            // column of the diagonal inside this chunk
        if (mv && dj >= 0 && dj < 32 && a.diag) {
          float dv = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) dv = (j == dj) ? v[j] : dv;
          a.diag[m] = dv * (1.0f / LOG2E);
        }
      }
      // natural-log convention of the partials: (max, sum exp(L - max))
      // one partial per (tile, column half); an empty half leaves (-inf, 0), which the combine ignores
      if (mv) a.part[(long)m * (2 * ntn) + 2 * tn + chalf] = make_float2(mx * (1.0f / LOG2E), sum);
    } else if (MODE == 3) {
      for (int cb = c_lo; cb < c_hi; cb += 32) {
        float v[32];
        tc::tmem_ld32(tacc + cb, v);
        if (mv) {
          float* o = a.out + (long)m * a.ldo + n0 + cb;
          if (cb + 32 <= ncols && (a.ldo & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(v[j] * scale, v[j + 1] * scale, v[j + 2] * scale, v[j + 3] * scale);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cb + j < ncols) o[j] = v[j] * scale;
          }
        }
      }
    } else {
      const float up = a.up ? *a.up : 1.f;
      const bool use_m = a.one_sided != 2, use_n = a.one_sided != 1;
      const float lm = (mv && use_m) ? a.lse_m[gm] : 0.f;
      const float kk = (a.one_sided ? 2.f * a.inv_2b : a.inv_2b) * up;
      const float dsub = a.one_sided ? 1.f : 2.f;
      float tsum = 0.f;
      // whole 64-column contraction chunks are written (zeros beyond the last valid column: the consumer reads full chunks)
      const int ccols = min(NT, ((ncols + KC - 1) / KC) * KC);
      uint8_t* grow = a.Gp + (size_t)tm * a.nkc_g * BLK + (size_t)(q * 32 + lane) * 16;
      for (int cb = c_lo; cb < min(ccols, c_lo + TL::HALF); cb += 32) {
        float v[32];
        tc::tmem_ld32(tacc + cb, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n0 + cb + j;
          float g = 0.f;
          if (cb + j < ncols) {
            const float Lg = v[j] * scale;
            g = use_m ? __expf(Lg - lm) : 0.f;
            if (use_n) g += __expf(Lg - __ldg(a.lse_n + n + a.n_off));
            g = (g - (gm == n + a.n_off ? dsub : 0.f)) * kk;
            if (mv) tsum += g * Lg;
          }
          v[j] = mv ? g : 0.f;
        }
        uint8_t* gblk = grow + (size_t)((n0 + cb) / KC) * BLK + (size_t)(((n0 + cb) % KC) / 8) * (RB * 16);
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint4 hi, lo;
          tc::split8(v + 8 * j8, hi, lo);
          *reinterpret_cast<uint4*>(gblk + j8 * (RB * 16)) = hi;
          *reinterpret_cast<uint4*>(gblk + j8 * (RB * 16) + PLANE) = lo;
        }
      }
      if (a.dtau_slots) {
        tsum_cta += tsum;
      } else if (a.dtau) {
        tsum = warp_sum(tsum);
        if (lane == 0) atomicAdd(a.dtau, tsum);
      }
    }
    tc::tc_fence_before();
    tc::mbar_arrive(&accempty[buf]);        // every epilogue thread: this buffer's columns have been read
    }
    if (MODE == 2 && a.dtau_slots) {
      tsum_cta = warp_sum(tsum_cta);
      if (lane == 0) a.dtau_slots[blockIdx.x * 8 + q + 4 * chalf] = tsum_cta;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tmem, 512);
}

inline bool head_tc_supported(int b, int Bg, int D) {
  return (D % KC) == 0 && D >= KC && b >= 1 && Bg >= b;
}

inline int epack(const float* X, uint8_t* P, int R, int D, cudaStream_t st) {
  dim3 grid(D / KC, (R + RB - 1) / RB);
  LAUNCH_PDL((epack_kernel), grid, 256, 0, st, X, P, R, D);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

inline int epack_t(const float* X, uint8_t* P, int R, int D, cudaStream_t st) {
  dim3 grid((R + KC - 1) / KC, (D + RB - 1) / RB);
  LAUNCH_PDL((epack_t_kernel), grid, 256, 0, st, X, P, R, D);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}
inline size_t epack_t_bytes(int R, int D) { return (size_t)((D + RB - 1) / RB) * ((R + KC - 1) / KC) * BLK; }

// d tau = sum of the per-(CTA, warp) slots of a MODE-2 launch, in slot order (one warp; bitwise reproducible)
__global__ void __launch_bounds__(32) dtau_fold_kernel(const float* __restrict__ slots, int n, float* __restrict__ dtau) {
  pdl_sync();
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) acc += slots[i];
  acc = warp_sum(acc);
  if (threadIdx.x == 0) *dtau = acc;
}
inline int logits_grid(const LogitsArgs& a, int ntc = NT) {
  const int ntiles = ((a.N + ntc - 1) / ntc) * ((a.M + RB - 1) / RB);
  return ntiles < 148 ? ntiles : 148;                     // persistent: one CTA per SM at most
}
inline int dtau_fold(const float* slots, int grid, float* dtau, cudaStream_t st) {
  LAUNCH_PDL((dtau_fold_kernel), 1, 32, 0, st, slots, grid * 8, dtau);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

template <int MODE, int NTERMS, int NTC>
inline int logits_launch_t(const LogitsArgs& a, cudaStream_t st) {
  static bool configured = false;
  const uint32_t smem = Tile<NTC>::NSTG * Tile<NTC>::STAGE_B + 256;
  if (!configured) {
    if (cudaFuncSetAttribute(logits_tc_kernel<MODE, NTERMS, NTC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return EEGCLIP_ERR_CUDA;
    configured = true;
  }
  const int grid = logits_grid(a, NTC);
  ProfScope prof(PROF_GEMM_F32, st);
  if (g_tune[5] == 0) pdl_break();   // plain stream order into the similarity kernels (measured -0.04 ms per step; see STREAM_BREAK, tower.cu)
  LAUNCH_PDL((logits_tc_kernel<MODE, NTERMS, NTC>), grid, 320, smem, st, a);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}
// the tile width is logits_ntc(MODE, a.M, a.N): callers size the MODE-1 partials (2 per tile) and the d tau slots with it
template <int MODE>
inline int logits_launch(int math, const LogitsArgs& a, cudaStream_t st) {
  if (MODE != 3 && logits_ntc(MODE, a.M, a.N) == 64)
    return math == EEGCLIP_MATH_BF16 ? logits_launch_t<MODE == 3 ? 1 : MODE, 1, 64>(a, st) : logits_launch_t<MODE == 3 ? 1 : MODE, 3, 64>(a, st);
  return math == EEGCLIP_MATH_BF16 ? logits_launch_t<MODE, 1, NT>(a, st) : logits_launch_t<MODE, 3, NT>(a, st);
}

}  // namespace headtc
}  // namespace eegclip
