// tcgen05 token GEMMs of the Transformer block (clip_model.py:24-26,31-33,43,63-66): QKV / out-projection / FFN
// linears, their data gradients and their weight gradients, on time-major (tokens, features) fp32 activations.
//
// All of these have a tiny contraction (K = 64..256) and a huge token count (M = B*T = 81 920 at B=256): they are
// HBM-bound (AI ~ 24..48 FLOP/B in fp32 storage), so the kernels are organised around streaming the activations once:
//
//   lin_tc_kernel   C[m][n] = epi( sum_k pro(A[m][k]) * W[n][k] )        (forward and data gradient: W or W^T packed)
//     persistent CTAs, warp-specialised:
//       warps 0-7  epilogue : TMEM -> registers -> per-warp smem transpose -> coalesced 128-bit global stores with
//                             bias / GELU(+pre-activation save) / Philox dropout / GELU' / residual fused
//                             (warp w: TMEM lane quarter w%4, column half w/4)
//       warps 8-11 producers: coalesced 128-bit loads of a 128-token x 64-feature fp32 chunk (+ optional dropout /
//                             GELU.dropout prologue), split into bf16 hi/lo planes in the chunk-major UMMA layout
//                             (tc_common.cuh), 3-stage ring signalled on mbarriers
//       warp  12   MMA      : one elected lane issues tcgen05.mma (M=128, N = full output width <= 256) into one of two
//                             TMEM accumulator buffers, so the epilogue of tile i overlaps the MMAs of tile i+1
//     (the erf / Philox work of the epilogues needs 8 warps: with 4 it was issue-bound at ~100 us per launch)
//     the packed weights (<= 64 KB) are fetched once per CTA by a 1-D bulk async copy (TMA engine) and stay resident.
//
//   lin_wgrad_tc_kernel   dW[n][k] = sum_m pro(dy[m][n]) * pro(x[m][k]),  db[n] = sum_m pro(dy[m][n])
//     contraction over tokens: both operands are MN-major views of the same chunk-major tiles (as the conv weight
//     gradient); each CTA accumulates its token slice in TMEM and writes one partial; lin_wgrad_reduce_kernel sums the
//     partials in a fixed order (deterministic gradients, no atomics).
//
// Arithmetic: NTERMS = 3 -> split-bf16 (hi*hi + hi*lo + lo*hi, fp32 accumulate); NTERMS = 1 -> plain bf16.
#pragma once
#include <string.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace eegclip {
namespace lintc {

constexpr int BM = 128;                  // tokens per tile
constexpr int KC = 64;                   // features per pipeline chunk
constexpr int NSTAGE = 3;                // maximum ring depth (runtime depth in LinTcArgs::nstage)
constexpr int A_CS = BM * 16;            // stride between 8-feature chunks inside a plane (bytes)
constexpr int A_PLANE = (KC / 8) * A_CS; // 16 KB
constexpr int A_STAGE = 2 * A_PLANE;     // hi + lo
constexpr int EPI_LD = 36;               // floats per staged row (32 + pad: 16 B aligned, conflict-free 128-bit access)
constexpr int EPI_WARP_FLOATS = 32 * EPI_LD;
// 13 warps (<= 4 per scheduler -> 128 registers each): NE epilogue warps + (12 - NE) producer warps + 1 MMA warp.
//   NE = 8: heavy epilogues (GELU, wide outputs).   NE = 4: deep contractions / narrow outputs -- 8 producer warps, each with
//   the loads of the NEXT chunk in flight while it converts the current one (the K >= 192 shapes are load-latency bound).
constexpr int NWARPS = 13;
constexpr int NTHREADS = NWARPS * 32;

enum : int { PRO_NONE = 0, PRO_DROP = 1, PRO_GELU_DROP = 2 };
// Compile-time feature masks of the token-GEMM epilogue.  The kernels are instantiated per (prologue, epilogue mask) actually
// used by the encoder, so each variant carries only its own code: the all-features kernel was > 64 KB of SASS, twice the
// instruction cache, and every role ran ~4x slower than its instruction count predicts.  Mask -1 = generic (runtime flags).
enum : int { EF_BIAS = 1, EF_ACT1 = 2, EF_ACT2 = 4, EF_DROP = 8, EF_ACTGRAD = 16, EF_MULSRC = 32, EF_RES = 64, EF_LNBWD = 128 };
// EF_LNBWD (N = 64, four epilogue warps): the GEMM result is dh, the gradient w.r.t. a per-token LayerNorm(64) output; the epilogue
// runs the LayerNorm backward in place of the plain store:  C = residual + LNbwd(dh ; x = lnb_x, gamma = lnb_gamma)  and leaves the
// CTA's column sums of dh * xhat and dh in lnb_partial (folded by lin_wgrad_reduce).  Replaces a data-gradient launch that wrote dh
// plus an ln64_bwd launch that read it back with x and the residual (clip_model.py:84,89 backward).
constexpr int LNB_R = 2;                // rows whose loads are in flight together in that epilogue
constexpr int LNB_LD = 68;               // floats per staged row of the LayerNorm-backward epilogue (64 + pad, 16 B aligned)

// streaming 128-bit load.  The CTAs here keep > 160 KB of shared memory, which leaves only a few KB of L1: plain
// (allocating) loads then throttle on free L1 lines long before HBM saturates, so activations bypass L1 allocation.
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
// development timeline (tools/lin_timeline.py): role `who` of CTA 0 records (event id, globaltimer ns) pairs
__device__ __forceinline__ void dbg_mark(unsigned long long* dbg, int who, int& n, int ev) {
  if (dbg && n < 120) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    dbg[who * 256 + 2 * n] = (unsigned long long)ev;
    dbg[who * 256 + 2 * n + 1] = t;
    ++n;
    dbg[who * 256 + 255] = (unsigned long long)n;
  }
}
__device__ __forceinline__ float4 ld_act(const float4* p, int policy) { return policy == 1 ? __ldg(p) : ld_stream(p); }
// 32 contiguous bytes per lane.  As TWO 128-bit loads every warp instruction touches 32-byte sectors it uses only half of, and with
// L1::no_allocate the second instruction fetches the same sectors from L2 again: the producers' lane mapping (8 rows x four 32-byte
// pieces per warp instruction) streamed at 2.46 TB/s that way and at 5.43 TB/s with ONE 256-bit load per lane
// (tools/micro/load_probe.cu, one persistent CTA per SM) -- the reason every register-staged token kernel sat at 45-60 % of the
// copy bandwidth in round 1.  `p` must be 32-byte aligned for the 256-bit form (checked per launch: LinTcArgs::vec8).
__device__ __forceinline__ void ld_act8(const float* p, int policy, bool vec8, float4& a, float4& b) {
  if (vec8) {
    if (policy == 1)
      asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
    else
      asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
  } else {
    a = ld_act(reinterpret_cast<const float4*>(p), policy);
    b = ld_act(reinterpret_cast<const float4*>(p) + 1, policy);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight packing.  Packed operand B (Ntot x Ktot, "n" = output feature, "k" = contraction index):
//   plane p (0 hi, 1 lo), element (n,k) at  p*Ntot*Ktot*2 + (k/8)*(Ntot*16) + n*16 + (k%8)*2   bytes.
// A job copies an (n_cnt x k_cnt) block at (n_off,k_off) from a strided fp32 source: src[n*sn + k*sk].
// kind 1 jobs copy n_cnt floats (bias concatenation).
// ------------------------------------------------------------------------------------------------
struct PackJob {
  const float* src;
  uint8_t* dst;
  int kind, Ntot, Ktot, n_off, k_off, n_cnt, k_cnt, sn, sk;
};
constexpr int MAX_PACK_JOBS = 16;
struct PackJobs { PackJob j[MAX_PACK_JOBS]; int n; };

static __global__ void __launch_bounds__(256) pack_lin_weights_kernel(const PackJobs jobs) {
  pdl_sync();
  const PackJob& J = jobs.j[blockIdx.y];
  if (J.kind == 1) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < J.n_cnt; i += gridDim.x * blockDim.x)
      reinterpret_cast<float*>(J.dst)[J.n_off + i] = J.src[i];
    return;
  }
  const int kch = J.k_cnt >> 3;
  const long plane = (long)J.Ntot * J.Ktot * 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < J.n_cnt * kch; i += gridDim.x * blockDim.x) {
    const int n = i % J.n_cnt, c = i / J.n_cnt;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = J.src[(long)n * J.sn + (long)(c * 8 + e) * J.sk];
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    uint8_t* d = J.dst + (long)((J.k_off >> 3) + c) * (J.Ntot * 16) + (long)(J.n_off + n) * 16;
    *reinterpret_cast<uint4*>(d) = hi;
    *reinterpret_cast<uint4*>(d + plane) = lo;
  }
}

inline int pack_launch(const PackJobs& jobs, cudaStream_t st) {
  if (jobs.n <= 0) return EEGCLIP_OK;
  LAUNCH_PDL((pack_lin_weights_kernel), dim3(8, jobs.n), 256, 0, st, jobs);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}
inline void add_pack(PackJobs& J, const float* src, uint8_t* dst, int Ntot, int Ktot, int n_off, int k_off, int n_cnt, int k_cnt,
                     int sn, int sk) {
  PackJob& j = J.j[J.n++];
  j.src = src; j.dst = dst; j.kind = 0; j.Ntot = Ntot; j.Ktot = Ktot; j.n_off = n_off; j.k_off = k_off; j.n_cnt = n_cnt; j.k_cnt = k_cnt;
  j.sn = sn; j.sk = sk;
}
inline void add_copy(PackJobs& J, const float* src, float* dst, int off, int cnt) {
  PackJob& j = J.j[J.n++];
  j.src = src; j.dst = reinterpret_cast<uint8_t*>(dst); j.kind = 1; j.Ntot = j.Ktot = 0; j.n_off = off; j.k_off = 0; j.n_cnt = cnt;
  j.k_cnt = 0; j.sn = j.sk = 0;
}
inline size_t packed_bytes(int N, int K) { return (size_t)N * K * 4; }

// ------------------------------------------------------------------------------------------------
// Shared producer: stage ROWS x 64 fp32 features (row-major, leading dimension ld) as bf16 hi/lo chunk-major planes.
//   lane -> (row = lane & 7, chunk = lane >> 3) inside an 8-row x 4-chunk block: each row contributes 128 contiguous
//   bytes to a request (full lines) and the eight 16-byte smem stores of a quarter-warp fill one 128-byte wavefront.
// Every thread keeps a FIXED chunk ((pw & 1) * 4 + (lane >> 3)), so column sums can live in registers.
// ------------------------------------------------------------------------------------------------
template <int ROWS, int NPW>
struct ChunkRegs { float4 x0[(ROWS / 8) * 2 / NPW], x1[(ROWS / 8) * 2 / NPW]; };

// issue the global loads of this thread's share of a ROWS x 64 chunk
template <int ROWS, int NPW>
__device__ __forceinline__ void load_chunk(ChunkRegs<ROWS, NPW>& R, const float* __restrict__ src, long ld, long row0, long rows_total,
                                           int col0, int pw, int lane, int policy, bool vec8 = false) {
  constexpr int ITERS = (ROWS / 8) * 2 / NPW;
  static_assert(((ROWS / 8) * 2) % NPW == 0, "producer warps must divide the chunk");
  const int ch = (pw & 1) * 4 + (lane >> 3);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int rb = it * (NPW / 2) + (pw >> 1);
    const long r = row0 + rb * 8 + (lane & 7);
    if (r < rows_total) {
      ld_act8(src + r * ld + col0 + ch * 8, policy, vec8, R.x0[it], R.x1[it]);
    } else {
      R.x0[it] = make_float4(0.f, 0.f, 0.f, 0.f); R.x1[it] = R.x0[it];
    }
  }
}

// prologue + bf16 hi/lo split + shared-memory store of the loaded share
// nxt_src != nullptr: as soon as iteration `it` of the chunk has been converted, the loads of iteration `it` of a LATER chunk (rows
// nxt_row0.., columns nxt_col0..) are issued into the same registers -- a rolling prefetch that keeps two chunks of loads in
// flight per thread without a third register set.
template <int ROWS, int NTERMS, int NPW, int PRO /* -1: runtime `pro` */>
__device__ __forceinline__ void convert_chunk(ChunkRegs<ROWS, NPW>& R, long row0, long rows_total, int col0, int ncols_total,
                                              uint8_t* dst, uint32_t CS, uint32_t PS, int pw, int lane, int pro, const Drop& drop,
                                              float* colsum /* nullptr or 8 running sums */, const float* __restrict__ nxt_src = nullptr,
                                              long nxt_ld = 0, long nxt_row0 = 0, int nxt_col0 = 0, int policy = 0, bool vec8 = false) {
  constexpr int ITERS = (ROWS / 8) * 2 / NPW;
  const int ch = (pw & 1) * 4 + (lane >> 3);
  const int prog = PRO < 0 ? pro : PRO;
  // 1-bit dropout: a warp's share of the chunk is ITERS x 8 rows x 32 columns = one 32-bit Philox word per row.  Lane l draws
  // the words of rows (l >> 3, l & 7) of iterations 4j + (l >> 3) once, the iterations fetch theirs by shuffle: ITERS / 4
  // Philox calls per thread and chunk instead of ITERS (the prologue was most of the producers' instruction count).
  const bool fast = prog != PRO_NONE && drop.enabled && drop.onebit && (ncols_total & 31) == 0;
  constexpr int NW = (ITERS + 3) / 4;
  uint32_t wown[NW];
  if (fast) {
#pragma unroll
    for (int j = 0; j < NW; ++j) {
      const int it_ = 4 * j + (lane >> 3);
      const int rb_ = it_ * (NPW / 2) + (pw >> 1);
      const uint64_t idx0 = (uint64_t)(row0 + rb_ * 8 + (lane & 7)) * (uint64_t)ncols_total + (uint64_t)(col0 + (pw & 1) * 32);
      const uint4 w4 = drop_words(drop, idx0 >> 7);
      wown[j] = word_of(w4, (uint32_t)(idx0 >> 5) & 3u);
    }
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int rb = it * (NPW / 2) + (pw >> 1);
    const int rl = rb * 8 + (lane & 7);
    float v[8] = {R.x0[it].x, R.x0[it].y, R.x0[it].z, R.x0[it].w, R.x1[it].x, R.x1[it].y, R.x1[it].z, R.x1[it].w};
    if (fast) {
      const uint32_t w = __shfl_sync(0xffffffffu, wown[it >> 2], ((it & 3) << 3) | (lane & 7));
      const uint32_t bits = w >> ((lane >> 3) * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (prog == PRO_GELU_DROP ? gelu_f(v[e]) : v[e]) * ((bits >> e) & 1u ? drop.scale : 0.f);
    } else if (prog != PRO_NONE) {
      const long r = row0 + rl;
      if (r < rows_total) {
        const uint64_t idx = (uint64_t)r * (uint64_t)ncols_total + (uint64_t)(col0 + ch * 8);
        float mm[8];
        drop_mult8(drop, idx, mm);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (prog == PRO_GELU_DROP ? gelu_f(v[e]) : v[e]) * mm[e];
      }
    }
    if (colsum) {
#pragma unroll
      for (int e = 0; e < 8; ++e) colsum[e] += v[e];
    }
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    uint8_t* d = dst + ch * CS + rl * 16;
    *reinterpret_cast<uint4*>(d) = hi;
    if (NTERMS > 1) *reinterpret_cast<uint4*>(d + PS) = lo;
    if (nxt_src) {
      const long r = nxt_row0 + rl;
      if (r < rows_total) {
        ld_act8(nxt_src + r * nxt_ld + nxt_col0 + ch * 8, policy, vec8, R.x0[it], R.x1[it]);
      } else {
        R.x0[it] = make_float4(0.f, 0.f, 0.f, 0.f); R.x1[it] = R.x0[it];
      }
    }
  }
}

template <int ROWS, int NTERMS, int NPW /* producer warps */, int PRO /* -1: runtime `pro` */>
__device__ __forceinline__ void stage_chunk(const float* __restrict__ src, long ld, long row0, long rows_total, int col0, int ncols_total,
                                            uint8_t* dst, uint32_t CS, uint32_t PS, int pw, int lane, int pro, const Drop& drop,
                                            float* colsum /* nullptr or 8 running sums */, int policy, bool vec8 = false) {
  ChunkRegs<ROWS, NPW> R;
  load_chunk<ROWS, NPW>(R, src, ld, row0, rows_total, col0, pw, lane, policy, vec8);
  convert_chunk<ROWS, NTERMS, NPW, PRO>(R, row0, rows_total, col0, ncols_total, dst, CS, PS, pw, lane, pro, drop, colsum);
}

// ------------------------------------------------------------------------------------------------
// Forward / data-gradient GEMM
// ------------------------------------------------------------------------------------------------
struct LinTcArgs {
  const float* A; long lda;       // (M, K) fp32
  const uint8_t* wpacked;         // packed (N x K) operand, see pack_lin_weights_kernel
  float* C; long ldc;             // (M, N) fp32; aux / act_grad_src / residual share C's indexing
  int M, N, K;
  int pro; Drop pro_drop;         // prologue on A (element index m*K + k)
  const float* bias;              // + bias[n]
  int act; float* aux;            // act == 1: aux = v (pre-activation, optional); v = GELU(v)
                                  // act == 2: v = GELU(pre), aux = GELU'(pre), BOTH multiplied by the dropout mask below
  const float* mul_src;           // v *= mul_src[m][n]
  int drop_on; Drop drop;         // v *= dropmult(m*N + n)
  const float* act_grad_src;      // v *= GELU'(src[m][n])
  const float* residual;          // v += residual[m][n]
  const float* lnb_x;             // EF_LNBWD: LayerNorm input rows (M, 64)
  const float* lnb_gamma;         // EF_LNBWD: LayerNorm weight (64)
  float* lnb_partial;             // EF_LNBWD: [gridDim.x][130]: sum dh*xhat (64), sum dh (64), 2 pad
  int nstage;                     // ring depth (2..NSTAGE)
  int policy;                     // load policy of the streamed activations (0: L1 no-allocate, 1: __ldg)
  int vec8;                       // A (and the LayerNorm-backward side streams) are 32-byte aligned: 256-bit loads
  int late_trigger;               // g_tune[2]: griddepcontrol.launch_dependents after the last MMA instead of first thing (A/B timing)
  unsigned long long* dbg;        // development timeline buffer (nullptr in production)
};

inline uint32_t lin_smem_bytes(int N, int K, int nstage = NSTAGE, int nepi = 8) {
  return (uint32_t)packed_bytes(N, K) + nstage * A_STAGE + nepi * EPI_WARP_FLOATS * 4 + 256 + 1024;   // + barriers + bias[256]
}

template <int NTERMS, int PRO, int EF, int NEPI>
__global__ void __launch_bounds__(NTHREADS, 1) lin_tc_kernel(const LinTcArgs a) {
  if (!a.late_trigger) pdl_trigger();
  constexpr int NPROD = NWARPS - 1 - NEPI;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, K = a.K;
  const uint32_t WP = (uint32_t)N * K * 2;                 // weight plane bytes
  uint8_t* sW = smem;
  uint8_t* sA = smem + 2 * WP;
  const int nstage = a.nstage;
  float* sE = reinterpret_cast<float*>(sA + nstage * A_STAGE);
  constexpr bool LNB = EF >= 0 && (EF & EF_LNBWD) != 0;
  constexpr int EPI_SLOTS = LNB ? 8 : NEPI;       // the LayerNorm-backward epilogue stages 32 x 64 per warp: two slots each
  static_assert(!LNB || NEPI == 4, "EF_LNBWD runs with four epilogue warps");
  uint64_t* bars = reinterpret_cast<uint64_t*>(sE + EPI_SLOTS * EPI_WARP_FLOATS);
  uint64_t* full = bars;                  // [NSTAGE] producers -> MMA   (NPROD*32 arrivals)
  uint64_t* empty = bars + NSTAGE;        // [NSTAGE] MMA -> producers   (tcgen05.commit)
  uint64_t* accfull = bars + 2 * NSTAGE;  // [2] MMA -> epilogue
  uint64_t* accempty = accfull + 2;       // [2] epilogue -> MMA         (NEPI*32 arrivals)
  uint64_t* wfull = accempty + 2;         // weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
  float* sBias = reinterpret_cast<float*>(bars) + 64;      // 256 B past the barrier block: bias[N] (zeros when absent)
  constexpr int MMA_WARP = NEPI + NPROD;

  uint32_t ncols = 32;
  while (ncols < 2u * (uint32_t)N) ncols <<= 1;
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&full[i], NPROD); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&accfull[i], 1); tc::mbar_init(&accempty[i], NEPI * 32); }
    tc::mbar_init(wfull, 1);
    tc::mbar_fence_init();
  }
  if (warp == MMA_WARP) tc::tmem_alloc(tmem_slot, ncols);
  pdl_wait();   // everything above touches only this CTA's shared memory / TMEM; global memory is read from here on
  for (int n = tid; n < N; n += NTHREADS) sBias[n] = a.bias ? __ldg(a.bias + n) : 0.f;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ntiles = (a.M + BM - 1) / BM;
  const int nchunk = K / KC;

  if (warp >= NEPI && warp < MMA_WARP) {
    // ===== producers =====
    const int pw = warp - NEPI;
    int s = 0; uint32_t ph = 0;
    unsigned long long* dbg = (blockIdx.x == 0 && pw == 0 && lane == 0) ? a.dbg : nullptr;
    int dn = 0;
    dbg_mark(dbg, 0, dn, 0);
    if (NPROD >= 8) {
      // software-pipelined: chunk c+1's loads are issued before chunk c is converted (2 x 8 float4 per thread in flight)
      const int my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const int total = my_tiles * nchunk;
      // two chunks of loads in flight per thread at all times: R[0], R[1] are loaded up front, and while chunk c is converted
      // its registers are refilled, iteration by iteration, with chunk c + 2 (in-kernel timeline: with one-chunk look-ahead the
      // chunks arrived in pairs with ~2 us of exposed latency between the pairs)
      ChunkRegs<BM, NPROD> R[2];
      auto chunk_pos = [&](int c_, int& tile_, int& kc_) { tile_ = blockIdx.x + (c_ / nchunk) * gridDim.x; kc_ = c_ % nchunk; };
      const bool v8 = a.vec8 != 0;
      if (total > 0) load_chunk<BM, NPROD>(R[0], a.A, a.lda, (long)blockIdx.x * BM, a.M, 0, pw, lane, a.policy, v8);
      if (total > 1) { int t1, k1; chunk_pos(1, t1, k1); load_chunk<BM, NPROD>(R[1], a.A, a.lda, (long)t1 * BM, a.M, k1 * KC, pw, lane, a.policy, v8); }
      int tile = blockIdx.x, kc = 0;
      for (int c = 0; c < total; c += 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (c + u < total) {
            int ntile = tile, nkc = kc + 1;
            if (nkc == nchunk) { nkc = 0; ntile += gridDim.x; }
            int t2 = 0, k2 = 0;
            const bool more = c + u + 2 < total;
            if (more) chunk_pos(c + u + 2, t2, k2);
            tc::mbar_wait(&empty[s], ph ^ 1);
            dbg_mark(dbg, 0, dn, 1);
            convert_chunk<BM, NTERMS, NPROD, PRO>(R[u], (long)tile * BM, a.M, kc * KC, K, sA + s * A_STAGE, A_CS, A_PLANE, pw, lane, a.pro,
                                                  a.pro_drop, nullptr, more ? a.A : nullptr, a.lda, (long)t2 * BM, k2 * KC, a.policy, v8);
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full[s]);
            dbg_mark(dbg, 0, dn, 2);
            if (++s == nstage) { s = 0; ph ^= 1; }
            tile = ntile; kc = nkc;
          }
        }
      }
    } else {
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int kc = 0; kc < nchunk; ++kc) {
          tc::mbar_wait(&empty[s], ph ^ 1);
          dbg_mark(dbg, 0, dn, 1);
          stage_chunk<BM, NTERMS, NPROD, PRO>(a.A, a.lda, (long)tile * BM, a.M, kc * KC, K, sA + s * A_STAGE, A_CS, A_PLANE, pw, lane, a.pro,
                                              a.pro_drop, nullptr, a.policy, a.vec8 != 0);
          tc::fence_async_smem();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&full[s]);
          dbg_mark(dbg, 0, dn, 2);
          if (++s == nstage) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===== weights (once) + MMA issue: the whole warp runs the loop, one elected lane issues =====
    const uint32_t wbytes = NTERMS > 1 ? 2 * WP : WP;
    if (tc::elect_one()) {
      tc::mbar_expect_tx(wfull, wbytes);
      tc::bulk_g2s(sW, a.wpacked, wbytes, wfull);
    }
    __syncwarp();
    tc::mbar_wait(wfull, 0);
    unsigned long long* dbg = (blockIdx.x == 0 && lane == 0) ? a.dbg : nullptr;
    int dn = 0;
    dbg_mark(dbg, 1, dn, 10);
    const uint32_t sA_u = tc::smem_u32(sA), sW_u = tc::smem_u32(sW);
    const uint32_t idesc = tc::idesc_bf16(128, N, 0, 0);
    const uint32_t nb16 = (uint32_t)N * 16u;
    uint32_t t = 0, ph = 0;
    int s = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
      const uint32_t buf = t & 1;
      tc::mbar_wait(&accempty[buf], ((t >> 1) & 1) ^ 1);
      tc::tc_fence_after();
      dbg_mark(dbg, 1, dn, 11);
      const uint32_t d = tmem + buf * (uint32_t)N;
      for (int kc = 0; kc < nchunk; ++kc) {
        tc::mbar_wait(&full[s], ph);
        tc::tc_fence_after();
        dbg_mark(dbg, 1, dn, 12);
        const uint64_t a_hi = tc::smem_desc(sA_u + s * A_STAGE, A_CS, 128);
        const uint64_t a_lo = tc::smem_desc(sA_u + s * A_STAGE + A_PLANE, A_CS, 128);
        const uint64_t b_hi = tc::smem_desc(sW_u + (uint32_t)(kc * 8) * nb16, nb16, 128);
        const uint64_t b_lo = tc::smem_desc(sW_u + WP + (uint32_t)(kc * 8) * nb16, nb16, 128);
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks) {
            const uint64_t da = (uint64_t)((2 * ks * A_CS) >> 4);          // descriptor start-address field is in 16-byte units
            const uint64_t db = (uint64_t)((2 * ks * nb16) >> 4);
            tc::mma_bf16(d, a_hi + da, b_hi + db, idesc, (kc | ks) != 0);   // hi*hi
            if (NTERMS > 1) {
              tc::mma_bf16(d, a_hi + da, b_lo + db, idesc, 1);              // hi*lo
              tc::mma_bf16(d, a_lo + da, b_hi + db, idesc, 1);              // lo*hi
            }
          }
          tc::tc_commit(&empty[s]);
        }
        __syncwarp();
        if (++s == nstage) { s = 0; ph ^= 1; }
      }
      if (tc::elect_one()) tc::tc_commit(&accfull[buf]);
      __syncwarp();
    }
    if (a.late_trigger) pdl_trigger();    // experiment: dependents may launch once every CTA has issued its last MMA
  } else {
    // ===== epilogue: warp w reads TMEM lane quarter w%4, columns [ (w/4)*N/2, (w/4+1)*N/2 ) =====
    if constexpr (LNB) {
      // ---- LayerNorm(64) backward epilogue: warp q owns the 32 rows of TMEM lane quarter q, all 64 columns ----
      const int q = warp & 3;
      float* stg = sE + warp * 2 * EPI_WARP_FLOATS;                 // [32][LNB_LD]
      const int c8 = (lane & 7) * 8, rq = lane >> 3;                // coalesced phase: 8 lanes per row, 4 rows per step
      float g8[8], ag[8], ab[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { g8[j] = __ldg(a.lnb_gamma + c8 + j); ag[j] = 0.f; ab[j] = 0.f; }
      uint32_t t = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
        const uint32_t buf = t & 1;
        tc::mbar_wait(&accfull[buf], (t >> 1) & 1);
        tc::tc_fence_after();
        const long mrow0 = (long)tile * BM + q * 32;
#pragma unroll
        for (int cb = 0; cb < 64; cb += 32) {
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)N + cb, v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(stg + lane * LNB_LD + cb + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        tc::tc_fence_before();
        tc::mbar_arrive(&accempty[buf]);                            // the accumulator is free as soon as it sits in shared memory
        __syncwarp();
        // LNB_R rows in flight per thread: their loads are issued together (one exposed memory latency per pair)
#pragma unroll 1
        for (int i0 = 0; i0 < 8; i0 += LNB_R) {
          float xv[LNB_R][8], rv[LNB_R][8], dh[LNB_R][8];
          bool ok[LNB_R];
#pragma unroll
          for (int u = 0; u < LNB_R; ++u) {
            const int rl = (i0 + u) * 4 + rq;
            const long m = mrow0 + rl;
            ok[u] = m < a.M;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 x0 = z4, x1 = z4, r0 = z4, r1 = z4;
            if (ok[u]) {
              ld_act8(a.lnb_x + m * 64 + c8, a.policy, a.vec8 != 0, x0, x1);
              if (a.residual) ld_act8(a.residual + m * a.ldc + c8, a.policy, a.vec8 != 0, r0, r1);
            }
            const float4 d0 = *reinterpret_cast<const float4*>(stg + rl * LNB_LD + c8);
            const float4 d1 = *reinterpret_cast<const float4*>(stg + rl * LNB_LD + c8 + 4);
            xv[u][0] = x0.x; xv[u][1] = x0.y; xv[u][2] = x0.z; xv[u][3] = x0.w; xv[u][4] = x1.x; xv[u][5] = x1.y; xv[u][6] = x1.z; xv[u][7] = x1.w;
            rv[u][0] = r0.x; rv[u][1] = r0.y; rv[u][2] = r0.z; rv[u][3] = r0.w; rv[u][4] = r1.x; rv[u][5] = r1.y; rv[u][6] = r1.z; rv[u][7] = r1.w;
            dh[u][0] = d0.x; dh[u][1] = d0.y; dh[u][2] = d0.z; dh[u][3] = d0.w; dh[u][4] = d1.x; dh[u][5] = d1.y; dh[u][6] = d1.z; dh[u][7] = d1.w;
          }
#pragma unroll
          for (int u = 0; u < LNB_R; ++u) {
            // row statistics over the 8 lanes that share the row (same arithmetic as ln64_fwd / ln64_bwd)
            float sx = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) sx += xv[u][j];
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) sx += __shfl_xor_sync(0xffffffffu, sx, o, 8);
            const float mean = sx * (1.f / 64.f);
            float qv = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { xv[u][j] -= mean; qv = fmaf(xv[u][j], xv[u][j], qv); }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) qv += __shfl_xor_sync(0xffffffffu, qv, o, 8);
            const float rstd = rsqrtf(qv * (1.f / 64.f) + 1e-5f);
            float s1 = 0.f, s2 = 0.f, e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              xv[u][j] *= rstd;                                     // xhat
              e[j] = dh[u][j] * g8[j];
              s1 += e[j];
              s2 = fmaf(e[j], xv[u][j], s2);
              ag[j] = fmaf(dh[u][j], xv[u][j], ag[j]);              // rows beyond M staged zeros: they add nothing
              ab[j] += dh[u][j];
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
              s1 += __shfl_xor_sync(0xffffffffu, s1, o, 8);
              s2 += __shfl_xor_sync(0xffffffffu, s2, o, 8);
            }
            const float m1 = s1 * (1.f / 64.f), m2 = s2 * (1.f / 64.f);
            if (ok[u]) {
              float o8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) o8[j] = rv[u][j] + rstd * (e[j] - m1 - xv[u][j] * m2);
              float* op = a.C + (mrow0 + (i0 + u) * 4 + rq) * a.ldc + c8;
              *reinterpret_cast<float4*>(op) = make_float4(o8[0], o8[1], o8[2], o8[3]);
              *reinterpret_cast<float4*>(op + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
            }
          }
        }
        __syncwarp();
      }
      // column sums of this warp: fold the four row groups (lanes with equal lane & 7), park them for the CTA-wide fold below
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ag[j] += __shfl_xor_sync(0xffffffffu, ag[j], 8); ag[j] += __shfl_xor_sync(0xffffffffu, ag[j], 16);
        ab[j] += __shfl_xor_sync(0xffffffffu, ab[j], 8); ab[j] += __shfl_xor_sync(0xffffffffu, ab[j], 16);
      }
      if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { stg[c8 + j] = ag[j]; stg[64 + c8 + j] = ab[j]; }
      }
    } else {
    const int q = warp & 3, chalf = warp >> 2;
    float* stg = sE + warp * EPI_WARP_FLOATS;
    const int cq = (lane & 7) * 4, rq = lane >> 3;      // coalesced phase: 4 columns x (4 rows per iteration)
    const int ncol_w = NEPI == 8 ? (N >> 1) : N;          // columns of this warp: half of N (8 warps) or all of it (4 warps)
    const int cb0 = chalf * ncol_w, cb1 = cb0 + ncol_w;
    uint32_t t = 0;
    unsigned long long* dbg = (blockIdx.x == 0 && warp == 0 && lane == 0) ? a.dbg : nullptr;
    int dn = 0;
    dbg_mark(dbg, 2, dn, 20);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
      const uint32_t buf = t & 1;
      tc::mbar_wait(&accfull[buf], (t >> 1) & 1);
      tc::tc_fence_after();
      dbg_mark(dbg, 2, dn, 21);
      const long mrow0 = (long)tile * BM + q * 32;
      for (int cb = cb0; cb < cb1; cb += 32) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)N + cb, v);
        dbg_mark(dbg, 2, dn, 23);
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(stg + lane * EPI_LD + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        dbg_mark(dbg, 2, dn, 24);
        const int n = cb + cq;
        const bool f_bias = EF < 0 ? a.bias != nullptr : (EF & EF_BIAS) != 0;
        const bool f_act1 = EF < 0 ? a.act == 1 : (EF & EF_ACT1) != 0;
        const bool f_act2 = EF < 0 ? a.act == 2 : (EF & EF_ACT2) != 0;
        const bool f_drop = EF < 0 ? a.drop_on != 0 : (EF & EF_DROP) != 0;
        const bool f_ag = EF < 0 ? a.act_grad_src != nullptr : (EF & EF_ACTGRAD) != 0;
        const bool f_mul = EF < 0 ? a.mul_src != nullptr : (EF & EF_MULSRC) != 0;
        const bool f_res = EF < 0 ? a.residual != nullptr : (EF & EF_RES) != 0;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f_bias) bb = *reinterpret_cast<const float4*>(sBias + n);
        // the side stream of the epilogue (residual, or the saved multiplier) is fetched for all 8 row groups before any
        // of it is used: one exposed memory latency per 32-column block instead of one per row group
        const float* side_p = f_res ? a.residual : (f_mul ? a.mul_src : nullptr);
        // 1-bit dropout (p = 0.5, every dropout of the transformer block): one Philox call yields 128 decisions of ONE row, so
        // lane l draws the word of row mrow0 + l for this 32-column block once and the coalesced phase fetches it by shuffle --
        // one call per block and warp instead of one per float4 (32x redundant: the FFN1 epilogue was issue-bound on it)
        const bool drop_fast = f_drop && a.drop.enabled && a.drop.onebit;
        uint32_t wown = 0u;
        if (drop_fast) {
          const uint64_t idx0 = (uint64_t)(mrow0 + lane) * (uint64_t)N + (uint64_t)cb;
          const uint4 w4 = drop_words(a.drop, idx0 >> 7);
          wown = word_of(w4, (uint32_t)(idx0 >> 5) & 3u);
        }
        // two halves of four row groups, each fully unrolled so that side[] stays in registers (one array of 8 indexed by a
        // partially unrolled loop lived in local memory: 128-byte stack frames in every residual / multiplier variant)
        for (int ih = 0; ih < 2; ++ih) {
        float4 side[4];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const long m = mrow0 + (ih * 4 + i4) * 4 + rq;
          side[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (side_p && m < a.M) side[i4] = *reinterpret_cast<const float4*>(side_p + m * a.ldc + n);
        }
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int i = ih * 4 + i4;
          const int rl = i * 4 + rq;
          const long m = mrow0 + rl;
          const uint32_t wrow = __shfl_sync(0xffffffffu, wown, rl);
          if (m < a.M) {
            float4 r = *reinterpret_cast<const float4*>(stg + rl * EPI_LD + cq);
            const long ci = m * a.ldc + n;
            r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
            float4 gd = make_float4(0.f, 0.f, 0.f, 0.f);
            if (f_act1) {
              if (a.aux) *reinterpret_cast<float4*>(a.aux + ci) = r;
              r.x = gelu_f(r.x); r.y = gelu_f(r.y); r.z = gelu_f(r.z); r.w = gelu_f(r.w);
            } else if (f_act2) {
              gelu_both(r.x, r.x, gd.x); gelu_both(r.y, r.y, gd.y); gelu_both(r.z, r.z, gd.z); gelu_both(r.w, r.w, gd.w);
            }
            if (f_drop) {
              float4 mk;
              if (drop_fast) {
                const uint32_t bits = wrow >> cq;
                mk = make_float4(bits & 1u ? a.drop.scale : 0.f, bits & 2u ? a.drop.scale : 0.f, bits & 4u ? a.drop.scale : 0.f,
                                 bits & 8u ? a.drop.scale : 0.f);
              } else {
                mk = drop_mult4(a.drop, (uint64_t)m * (uint64_t)N + (uint64_t)n);
              }
              r.x *= mk.x; r.y *= mk.y; r.z *= mk.z; r.w *= mk.w;
              gd.x *= mk.x; gd.y *= mk.y; gd.z *= mk.z; gd.w *= mk.w;
            }
            if (f_act2) *reinterpret_cast<float4*>(a.aux + ci) = gd;
            if (f_mul) {
              const float4 s4 = f_res ? ld_act(reinterpret_cast<const float4*>(a.mul_src + ci), a.policy) : side[i4];
              r.x *= s4.x; r.y *= s4.y; r.z *= s4.z; r.w *= s4.w;
            }
            if (f_ag) {
              const float4 s4 = ld_act(reinterpret_cast<const float4*>(a.act_grad_src + ci), a.policy);
              r.x *= gelu_grad_f(s4.x); r.y *= gelu_grad_f(s4.y); r.z *= gelu_grad_f(s4.z); r.w *= gelu_grad_f(s4.w);
            }
            if (f_res) {
              // (plain coherent loads above: the residual may alias C in the K-split accumulation passes)
              const float4 s4 = side[i4];
              r.x += s4.x; r.y += s4.y; r.z += s4.z; r.w += s4.w;
            }
            *reinterpret_cast<float4*>(a.C + ci) = r;
          }
        }
        }
        __syncwarp();
        dbg_mark(dbg, 2, dn, 25);
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&accempty[buf]);
      dbg_mark(dbg, 2, dn, 22);
    }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tc::tmem_dealloc(tmem, ncols);
  if constexpr (LNB) {
    // CTA-wide fold of the four epilogue warps' column sums (fixed order) -> this CTA's partial
    if (tid < 128) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) sum += sE[w * 2 * EPI_WARP_FLOATS + tid];
      a.lnb_partial[(long)blockIdx.x * 130 + tid] = sum;
    } else if (tid < 130) {
      a.lnb_partial[(long)blockIdx.x * 130 + tid] = 0.f;
    }
  }
}

inline bool lin_tc_supported(long M, int N, int K) {
  return M >= 1 && (N == 64 || N == 128 || N == 192 || N == 256) && (K % KC) == 0 && K >= KC &&
         lin_smem_bytes(N, K, 2) <= 227u * 1024u;
}

template <int NTERMS, int PRO, int EF, int NEPI>
inline int lin_tc_launch_w(const LinTcArgs& a, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(lin_tc_kernel<NTERMS, PRO, EF, NEPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return EEGCLIP_ERR_CUDA;
    configured = true;
  }
  constexpr int slots = (EF >= 0 && (EF & EF_LNBWD)) ? 8 : NEPI;
  LAUNCH_PDL((lin_tc_kernel<NTERMS, PRO, EF, NEPI>), grid, NTHREADS, lin_smem_bytes(a.N, a.K, a.nstage, slots), st, a);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}
// deep contractions with narrow outputs run the producer-heavy split, everything else the epilogue-heavy one
template <int NTERMS, int PRO, int EF>
inline int lin_tc_launch_v(const LinTcArgs& a, int grid, uint32_t, cudaStream_t st) {
  constexpr bool light = (EF >= 0) && (EF & (EF_ACT1 | EF_ACT2 | EF_ACTGRAD)) == 0;
  // g_tune[4]: 0 shipped rule, 1 never, 2 every narrow output, 3 every light epilogue (development sweeps)
  const bool narrow = g_tune[4] == 3 ? true : a.N <= 64;
  const bool deep = g_tune[4] >= 2 ? true : a.K >= 128;
  if (light && narrow && deep && g_tune[4] != 1) return lin_tc_launch_w<NTERMS, PRO, EF, 4>(a, grid, st);
  return lin_tc_launch_w<NTERMS, PRO, EF, 8>(a, grid, st);
}

// CTAs of a token-GEMM launch over M rows (persistent: one per SM at most) -- also the number of EF_LNBWD partials
inline int lin_tc_grid(long M) {
  const long ntiles = (M + BM - 1) / BM;
  return (int)(ntiles < 148 ? ntiles : 148);
}

template <int NTERMS>
inline int lin_tc_launch_t(LinTcArgs a, cudaStream_t st) {
  const int grid = lin_tc_grid(a.M);
  a.nstage = g_tune[0] >= 2 && g_tune[0] <= NSTAGE ? g_tune[0] : 2;
  while (a.nstage > 2 && lin_smem_bytes(a.N, a.K, a.nstage) > 227u * 1024u) --a.nstage;
  a.policy = g_tune[1];
  a.dbg = g_dbg_buf;
  // 256-bit loads need 32-byte aligned rows (g_tune[13] = 1 forces the 2 x 128-bit form for A/B timing)
  a.late_trigger = g_tune[2] & 1;
  a.vec8 = (g_tune[13] == 0 && ((uintptr_t)a.A & 31) == 0 && (a.lda & 7) == 0 &&
            (!a.lnb_x || ((((uintptr_t)a.lnb_x | (uintptr_t)a.residual) & 31) == 0 && (a.ldc & 7) == 0))) ? 1 : 0;
  const uint32_t smem = 0;
  const int mask = (a.bias ? EF_BIAS : 0) | (a.act == 1 ? EF_ACT1 : 0) | (a.act == 2 ? EF_ACT2 : 0) | (a.drop_on ? EF_DROP : 0) |
                   (a.act_grad_src ? EF_ACTGRAD : 0) | (a.mul_src ? EF_MULSRC : 0) | (a.residual ? EF_RES : 0);
  ProfScope prof(PROF_LIN_TC, st);
  if (a.lnb_x) {          // data gradient + LayerNorm backward in one launch (plain prologue, N = 64)
    if (a.N != 64 || a.pro != PRO_NONE || !a.lnb_gamma || !a.lnb_partial || a.bias || a.act || a.drop_on || a.act_grad_src || a.mul_src ||
        lin_smem_bytes(a.N, a.K, a.nstage, 8) > 227u * 1024u)
      return EEGCLIP_ERR_UNSUPPORTED;
    return lin_tc_launch_w<NTERMS, PRO_NONE, EF_LNBWD, 4>(a, grid, st);
  }
  if (g_tune[3] == 0) {   // g_tune[3] != 0 forces the generic kernel (development)
    // the variants the encoder launches: QKV / plain, out-proj + FFN2, FFN1, and the four data gradients
    if (a.pro == PRO_NONE) {
      if (mask == 0) return lin_tc_launch_v<NTERMS, PRO_NONE, 0>(a, grid, smem, st);
      if (mask == EF_BIAS) return lin_tc_launch_v<NTERMS, PRO_NONE, EF_BIAS>(a, grid, smem, st);
      if (mask == EF_RES) return lin_tc_launch_v<NTERMS, PRO_NONE, EF_RES>(a, grid, smem, st);
      if (mask == (EF_BIAS | EF_RES)) return lin_tc_launch_v<NTERMS, PRO_NONE, EF_BIAS | EF_RES>(a, grid, smem, st);
      if (mask == (EF_BIAS | EF_DROP | EF_RES)) return lin_tc_launch_v<NTERMS, PRO_NONE, EF_BIAS | EF_DROP | EF_RES>(a, grid, smem, st);
      if (mask == (EF_BIAS | EF_ACT2 | EF_DROP)) return lin_tc_launch_v<NTERMS, PRO_NONE, EF_BIAS | EF_ACT2 | EF_DROP>(a, grid, smem, st);
      if (mask == (EF_BIAS | EF_ACT2)) return lin_tc_launch_v<NTERMS, PRO_NONE, EF_BIAS | EF_ACT2>(a, grid, smem, st);
      if (mask == EF_MULSRC) return lin_tc_launch_v<NTERMS, PRO_NONE, EF_MULSRC>(a, grid, smem, st);
    } else if (a.pro == PRO_DROP) {
      if (mask == 0) return lin_tc_launch_v<NTERMS, PRO_DROP, 0>(a, grid, smem, st);
      if (mask == EF_MULSRC) return lin_tc_launch_v<NTERMS, PRO_DROP, EF_MULSRC>(a, grid, smem, st);
    }
  }
  return lin_tc_launch_v<NTERMS, -1, -1>(a, grid, smem, st);
}
inline int lin_tc_launch(int math, const LinTcArgs& a, cudaStream_t st) {
  return math == EEGCLIP_MATH_BF16 ? lin_tc_launch_t<1>(a, st) : lin_tc_launch_t<3>(a, st);
}

// ------------------------------------------------------------------------------------------------
// Weight gradient (contraction over tokens)
// ------------------------------------------------------------------------------------------------
constexpr int WT = 64;                   // tokens per stage
constexpr int W_CS = WT * 16;            // chunk stride (bytes)
constexpr int WG_NPROD = 16;             // producer warps (they are the ALU-heavy role: convert + GELU/Philox prologues)
constexpr int WG_THREADS = WG_NPROD * 32;   // warp 0 also issues the MMAs (16 warps: 128 registers each)
constexpr int WG_MAX_CTAS = 148;
constexpr int WG_MAXC = 5;               // 64-feature chunks of dy plus x per stage (Nout + Kin <= 320)

struct LinWgradArgs {
  const float* dy; long lddy; int Nout;   // (M, Nout): rows of dW
  const float* x; long ldx; int Kin;      // (M, Kin) : columns of dW
  int M;
  int pro_dy; Drop drop_dy;               // prologue on dy (element index m*Nout + n)
  int pro_x; Drop drop_x;                 // prologue on x  (element index m*Kin + k)
  float* partial;                         // [kin blocks][ctas][Nout*Kin + Nout]
  int want_db;
  int late_trigger;                       // (g_tune[2] & 2) griddepcontrol.launch_dependents after the last MMA (A/B timing)
  int policy;                             // load policy of the streamed activations (0: L1 no-allocate, 1: __ldg)
  unsigned long long* dbg;                // development timeline buffer (nullptr in production)
};

inline uint32_t wgrad_stage_bytes(int Nout, int Kin) { return (uint32_t)(Nout + Kin) * WT * 4; }
inline uint32_t wgrad_lin_smem_bytes(int Nout, int Kin) { return 2 * wgrad_stage_bytes(Nout, Kin) + 8 * 256 * 4 + 256; }

// PDY / PX: compile-time prologues of dy / x (-1 = runtime), WDB: bias-gradient column sums (-1 = runtime)
template <int NTERMS, int PDY, int PX, int WDB>
__global__ void __launch_bounds__(WG_THREADS, 1) lin_wgrad_tc_kernel(const LinWgradArgs a) {
  pdl_trigger();
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Nout = a.Nout, Kin = a.Kin;
  const bool want_db = WDB < 0 ? a.want_db != 0 : WDB != 0;
  const uint32_t PSD = (uint32_t)(Nout / 8) * W_CS, PSX = (uint32_t)(Kin / 8) * W_CS;   // plane strides
  const uint32_t STAGE = 2 * PSD + 2 * PSX;
  float* sCol = reinterpret_cast<float*>(smem + 2 * STAGE);     // [8 row blocks][256] column-sum staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(sCol + 8 * 256);
  uint64_t* full = bars;          // [2]
  uint64_t* empty = bars + 2;     // [2]
  uint64_t* accfull = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int nmt = (Nout + 127) / 128;                 // M tiles (rows of dW)
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(nmt * Kin)) ncols <<= 1;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&full[i], WG_NPROD); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(accfull, 1);
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, ncols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();   // global memory is read from here on

  // token range of this CTA, in units of WT-token stages
  const int nst_total = (a.M + WT - 1) / WT;
  const int per = (nst_total + gridDim.x - 1) / gridDim.x;
  const int st_beg = blockIdx.x * per;
  const int st_end = min(nst_total, st_beg + per);
  const int nst = max(0, st_end - st_beg);
  const int ndc = Nout / KC, nxc = Kin / KC;          // 64-feature chunks per operand
  const int ntot = ndc + nxc;

  {
    // ===== producers: warp pw owns row block pw>>1 (8 tokens) and chunk half pw&1 of every 64-feature chunk =====
    const int pw = warp;
    const int ch = (pw & 1) * 4 + (lane >> 3);
    const int rl = (pw >> 1) * 8 + (lane & 7);
    const uint32_t base = tc::smem_u32(smem);
    float cs[4][8];                                   // column sums of dy chunks
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) cs[i][e] = 0.f;
    // (measured and rejected: issuing the loads of stage it+1 before / while stage it is converted -- whole-stage double
    //  buffering with 32-token stages 2.60 ms per step, chunk-level rolling reissue 2.75 ms, this loop 2.44 ms)
    unsigned long long* dbgp = (blockIdx.x == 0 && blockIdx.y == 0 && warp == 3 && lane == 0) ? a.dbg : nullptr;
    unsigned long long* dbgm = (blockIdx.x == 0 && blockIdx.y == 0 && warp == 0 && lane == 0) ? a.dbg : nullptr;
    int dnp = 0, dnm = 0;
    dbg_mark(dbgp, 0, dnp, 0);
    // one 256-bit load per lane when both operands have 32-byte aligned rows (see ld_act8)
    const bool v8w = ((((uintptr_t)a.dy | (uintptr_t)a.x) & 31) == 0) && ((a.lddy | a.ldx) & 7) == 0;
    for (int it = 0; it < nst; ++it) {
      const int s = it & 1;
      uint8_t* sb = smem + s * STAGE;
      const long r = (long)(st_beg + it) * WT + rl;
      const bool ok = r < a.M;
      float4 x0[WG_MAXC], x1[WG_MAXC];
#pragma unroll
      for (int c = 0; c < WG_MAXC; ++c) {
        x0[c] = make_float4(0.f, 0.f, 0.f, 0.f); x1[c] = x0[c];
        if (c < ntot && ok) {
          const float* src = (c < ndc) ? a.dy + r * a.lddy + c * KC + ch * 8
                                       : a.x + r * a.ldx + (long)blockIdx.y * Kin + (c - ndc) * KC + ch * 8;
          x0[c] = ld_act(reinterpret_cast<const float4*>(src), a.policy); x1[c] = ld_act(reinterpret_cast<const float4*>(src) + 1, a.policy);
        }
      }
      dbg_mark(dbgp, 0, dnp, 1);                       // loads issued
      tc::mbar_wait(&empty[s], ((it >> 1) & 1) ^ 1);
      dbg_mark(dbgp, 0, dnp, 2);                       // slot free
#pragma unroll
      for (int c = 0; c < WG_MAXC; ++c) {
        if (c < ntot) {
          const bool is_dy = c < ndc;
          float v[8] = {x0[c].x, x0[c].y, x0[c].z, x0[c].w, x1[c].x, x1[c].y, x1[c].z, x1[c].w};
          const int pro = is_dy ? (PDY < 0 ? a.pro_dy : PDY) : (PX < 0 ? a.pro_x : PX);
          if (pro != PRO_NONE && ok) {
            const int col = (is_dy ? c : c - ndc) * KC + ch * 8;
            const uint64_t idx = (uint64_t)r * (uint64_t)(is_dy ? Nout : Kin) + (uint64_t)col;
            const Drop& dr = is_dy ? a.drop_dy : a.drop_x;
            float mm[8];
            drop_mult8(dr, idx, mm);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (pro == PRO_GELU_DROP ? gelu_f(v[e]) : v[e]) * mm[e];
          }
          if (c < 4 && is_dy && want_db) {
#pragma unroll
            for (int e = 0; e < 8; ++e) cs[c < 4 ? c : 0][e] += v[e];
          }
          uint4 hi, lo;
          tc::split8(v, hi, lo);
          uint8_t* d = sb + (is_dy ? (uint32_t)(c * 8) * W_CS : 2 * PSD + (uint32_t)((c - ndc) * 8) * W_CS) + ch * W_CS + rl * 16;
          *reinterpret_cast<uint4*>(d) = hi;
          if (NTERMS > 1) *reinterpret_cast<uint4*>(d + (is_dy ? PSD : PSX)) = lo;
        }
      }
      tc::fence_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full[s]);          // one arrival per warp (512 per-thread arrivals serialise on the barrier)
      dbg_mark(dbgp, 0, dnp, 3);                       // converted + arrived
      if (warp == 0) {
        // ===== MMA (warp 0, after its own share of the stage): uniform descriptors, one elected lane issues =====
        tc::mbar_wait(&full[s], (it >> 1) & 1);
        tc::tc_fence_after();
        dbg_mark(dbgm, 1, dnm, 12);                    // stage full
        const uint32_t sD = base + s * STAGE, sX = sD + 2 * PSD;
        const uint64_t b_hi = tc::smem_desc(sX, 128, W_CS), b_lo = tc::smem_desc(sX + PSX, 128, W_CS);
        if (tc::elect_one()) {
          for (int mt = 0; mt < nmt; ++mt) {
            const int mrows = min(128, Nout - mt * 128);          // 128 or 64
            const uint32_t idesc = tc::idesc_bf16(mrows, Kin, 1, 1);
            const uint32_t d = tmem + (uint32_t)(mt * Kin);
            // A = dy, B = x (both MN-major, K = tokens): hi*hi, hi*lo, lo*hi
            const uint64_t a_hi = tc::smem_desc(sD + (uint32_t)(mt * 16) * W_CS, 128, W_CS);
            const uint64_t a_lo = tc::smem_desc(sD + PSD + (uint32_t)(mt * 16) * W_CS, 128, W_CS);
#pragma unroll
            for (int ks = 0; ks < WT / 16; ++ks) {
              const uint64_t dk = (uint64_t)(ks * 16);            // 16 token rows = 256 bytes = 16 address units
              tc::mma_bf16(d, a_hi + dk, b_hi + dk, idesc, (it | ks) != 0);
              if (NTERMS > 1) {
                tc::mma_bf16(d, a_hi + dk, b_lo + dk, idesc, 1);
                tc::mma_bf16(d, a_lo + dk, b_hi + dk, idesc, 1);
              }
            }
          }
          tc::tc_commit(&empty[s]);
          if (it == nst - 1) tc::tc_commit(accfull);
        }
        __syncwarp();
        dbg_mark(dbgm, 1, dnm, 13);                    // MMAs issued
      }
    }
    dbg_mark(dbgp, 0, dnp, 4);
    if (want_db) {
      // reduce over the 8 token lanes of a warp (lane & 7); the 8 row-block warps of a chunk half are summed below
#pragma unroll
      for (int dc = 0; dc < 4; ++dc)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float v = cs[dc][e];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          cs[dc][e] = v;
        }
      if ((lane & 7) == 0) {
#pragma unroll
        for (int dc = 0; dc < 4; ++dc)
          if (dc < ndc)
#pragma unroll
            for (int e = 0; e < 8; ++e) sCol[(pw >> 1) * 256 + dc * 64 + ch * 8 + e] = cs[dc][e];
      }
    }
  }
  __syncthreads();   // column sums staged; all roles done issuing
  // blockIdx.y selects a Kin-wide column block of x (wide layers: one launch covers all blocks)
  float* part = a.partial + ((long)blockIdx.y * gridDim.x + blockIdx.x) * ((long)Nout * Kin + Nout);
  if (warp < 4) {
    // ===== epilogue: accumulators -> partial[cta][n][k] =====
    const int q = warp;
    unsigned long long* dbge = (blockIdx.x == 0 && blockIdx.y == 0 && warp == 0 && lane == 0) ? a.dbg : nullptr;
    int dne = 0;
    dbg_mark(dbge, 2, dne, 20);
    if (nst > 0) {
      tc::mbar_wait(accfull, 0);
      tc::tc_fence_after();
    }
    dbg_mark(dbge, 2, dne, 21);
    for (int mt = 0; mt < nmt; ++mt) {
      const int mrows = min(128, Nout - mt * 128);
      // M = 128: row = q*32 + lane ; M = 64: row = q*16 + lane (lanes 0-15 of each sub-partition)
      const int row = mrows == 128 ? q * 32 + lane : q * 16 + lane;
      const bool valid = mrows == 128 || lane < 16;
      for (int cb = 0; cb < Kin; cb += 32) {
        float v[32];
        if (nst > 0) tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * Kin + cb), v);
        else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (valid) {
          float* o = part + (long)(mt * 128 + row) * Kin + cb;
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
    if (want_db) {
      float* pb = part + (long)Nout * Kin;
      for (int n = tid; n < Nout; n += 128) {
        float sum = 0.f;
        if (nst > 0)
#pragma unroll
          for (int g = 0; g < 8; ++g) sum += sCol[g * 256 + n];
        pb[n] = sum;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, ncols);
}

// ------------------------------------------------------------------------------------------------
// Weight gradient, bulk-copy fed (dense operands: lddy == Nout, ldx == Kin -- every weight gradient of the transformer block).
// The in-kernel timeline of the kernel above (tools/wgrad_timeline.py) showed 3.6 us per 64-token stage: 0.8 us to ISSUE the
// stage's 160 warp-level 128-bit loads, ~1.5 us of exposed latency, ~1 us of conversion, and warp 0 -- producer and MMA issuer
// at once -- held every stage back by another 0.9 us.  Here a stage (32 tokens) is TWO 1-D bulk async copies (the dy rows and
// the x rows are contiguous; one copy per row was request-bound in the copy engine: 64 requests ~ 1 us), three stages in flight,
// no registers and no LSU issue slots; sixteen producer warps convert landing -> bf16 hi/lo operand tiles; one warp issues
// the copies, one the MMAs.
//   landing: dense row-major fp32, read with consecutive lanes on consecutive 16-byte groups (conflict-free);
//   operand tiles: chunk-major with the chunk stride padded to 32*16 + 16 bytes, which makes the transposing 8-byte stores of
//   a half-warp (8 chunks x 2 halves) hit 32 different banks; the padding only changes the descriptors' stride field.
//   thread -> (fixed 4-feature group, row slot): column sums for the bias gradient stay in registers.
// ------------------------------------------------------------------------------------------------
constexpr int WT2 = 32;                  // tokens per stage
constexpr int W2_CS = WT2 * 16 + 16;     // padded chunk stride of the operand tiles (bytes)
constexpr int W2_NL = 3;                 // landing stages
constexpr int W2_NO = 2;                 // operand stages
constexpr int W2_TMA_WARP = WG_NPROD, W2_MMA_WARP = WG_NPROD + 1;
constexpr int W2_THREADS = (WG_NPROD + 2) * 32;

inline uint32_t wgrad2_smem_bytes(int Nout, int Kin) {
  return W2_NL * (uint32_t)(Nout + Kin) * WT2 * 4u + W2_NO * 2u * (uint32_t)((Nout + Kin) / 8) * W2_CS + 2048 * 4 + 256;
}

template <int NTERMS, int PDY, int PX, int WDB>
__global__ void __launch_bounds__(W2_THREADS, 1) lin_wgrad_tma_kernel(const LinWgradArgs a, const __grid_constant__ CUtensorMap tm_dy,
                                                                      const __grid_constant__ CUtensorMap tm_x, const int use_tm) {
  if (!a.late_trigger) pdl_trigger();
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Nout = a.Nout, Kin = a.Kin;
  const bool want_db = WDB < 0 ? a.want_db != 0 : WDB != 0;
  const uint32_t LD_BYTES = (uint32_t)Nout * WT2 * 4u, LX_BYTES = (uint32_t)Kin * WT2 * 4u;   // landing blocks of a stage
  const uint32_t LSTAGE = LD_BYTES + LX_BYTES;
  const uint32_t PSD = (uint32_t)(Nout / 8) * W2_CS, PSX = (uint32_t)(Kin / 8) * W2_CS;       // plane strides
  const uint32_t STAGE = 2 * PSD + 2 * PSX;
  uint8_t* sL = smem;
  uint8_t* sO = smem + W2_NL * LSTAGE;
  float* sCol = reinterpret_cast<float*>(sO + W2_NO * STAGE);     // [row slots][Nout] column-sum staging (<= 2048 floats)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sCol + 2048);
  uint64_t* lfull = bars;                 // [W2_NL] copies landed (expect_tx)
  uint64_t* lempty = bars + W2_NL;        // [W2_NL] producers done reading (16 warp arrivals)
  uint64_t* ofull = bars + 2 * W2_NL;     // [W2_NO] operand tile written (16 warp arrivals)
  uint64_t* oempty = ofull + W2_NO;       // [W2_NO] MMAs done (tcgen05.commit)
  uint64_t* accfull = oempty + W2_NO;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);

  const int nmt = (Nout + 127) / 128;                 // M tiles (rows of dW)
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(nmt * Kin)) ncols <<= 1;
  if (tid == 0) {
    for (int i = 0; i < W2_NL; ++i) { tc::mbar_init(&lfull[i], 1); tc::mbar_init(&lempty[i], WG_NPROD); }
    for (int i = 0; i < W2_NO; ++i) { tc::mbar_init(&ofull[i], WG_NPROD); tc::mbar_init(&oempty[i], 1); }
    tc::mbar_init(accfull, 1);
    tc::mbar_fence_init();
  }
  if (warp == W2_MMA_WARP) tc::tmem_alloc(tmem_slot, ncols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();   // global memory is read from here on

  // token range of this CTA, in units of WT2-token stages
  const int nst_total = (a.M + WT2 - 1) / WT2;
  const int per = (nst_total + gridDim.x - 1) / gridDim.x;
  const int st_beg = blockIdx.x * per;
  const int st_end = min(nst_total, st_beg + per);
  const int nst = max(0, st_end - st_beg);
  // producer geometry: a row of an operand is n4 = N / 4 float4 groups; rpp rows are converted per pass of the 512 threads
  const int n4d = Nout >> 2, n4x = Kin >> 2;
  const int rpp_d = min(WT2, (WG_NPROD * 32) / n4d), rpp_x = min(WT2, (WG_NPROD * 32) / n4x);

  if (warp == W2_TMA_WARP) {
    // ===== copy issuer: two bulk copies per stage =====
    const bool dense_dy = a.lddy == Nout, dense_x = a.ldx == Kin && gridDim.y == 1;
    unsigned long long* dbgt = (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) ? a.dbg : nullptr;
    int dnt = 0;
    for (int it = 0; it < nst; ++it) {
      const int ls = it % W2_NL;
      tc::mbar_wait(&lempty[ls], ((it / W2_NL) & 1) ^ 1);
      dbg_mark(dbgt, 2, dnt, 30);
      const long r0 = (long)(st_beg + it) * WT2;
      const int rows = (int)min((long)WT2, (long)a.M - r0);
      // a dense operand (leading dimension == width) is one 1-D bulk copy per stage; a strided one (column block of a wider
      // tensor: speech features, LSTM gate / state buffers) is ONE tensor-map copy of the 32-row box (rows past M arrive as
      // zeros and count for the transaction bytes).  (Without a tensor map: one copy per row, 32 requests per stage -- the copy
      // engine is request-bound there, and with both operands strided the register-staged kernel had to take over.)
      const bool tm_d = (use_tm & 1) != 0, tm_xx = (use_tm & 2) != 0;
      if (lane == 0) tc::mbar_expect_tx(&lfull[ls], (uint32_t)(tm_d ? WT2 : rows) * (uint32_t)Nout * 4u + (uint32_t)(tm_xx ? WT2 : rows) * (uint32_t)Kin * 4u);
      __syncwarp();
      uint8_t* ld = sL + ls * LSTAGE;
      if (tm_d) { if (lane == 0) tc::tma_load_2d(ld, &tm_dy, 0, (int)r0, &lfull[ls]); }
      else if (dense_dy) { if (lane == 0) tc::bulk_g2s(ld, a.dy + r0 * Nout, (uint32_t)rows * (uint32_t)Nout * 4u, &lfull[ls]); }
      else if (lane < rows) tc::bulk_g2s(ld + (uint32_t)lane * Nout * 4u, a.dy + (r0 + lane) * a.lddy, (uint32_t)Nout * 4u, &lfull[ls]);
      if (tm_xx) { if (lane == 0) tc::tma_load_2d(ld + LD_BYTES, &tm_x, (int)blockIdx.y * Kin, (int)r0, &lfull[ls]); }
      else if (dense_x) { if (lane == 0) tc::bulk_g2s(ld + LD_BYTES, a.x + r0 * Kin, (uint32_t)rows * (uint32_t)Kin * 4u, &lfull[ls]); }
      else if (lane < rows) tc::bulk_g2s(ld + LD_BYTES + (uint32_t)lane * Kin * 4u, a.x + (r0 + lane) * a.ldx + (long)blockIdx.y * Kin, (uint32_t)Kin * 4u, &lfull[ls]);
      __syncwarp();
    }
  } else if (warp == W2_MMA_WARP) {
    // ===== MMA issuer: uniform descriptors, one elected lane issues =====
    const uint32_t base = tc::smem_u32(sO);
    unsigned long long* dbgm = (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) ? a.dbg : nullptr;
    int dnm = 0;
    for (int it = 0; it < nst; ++it) {
      const int s = it % W2_NO;
      tc::mbar_wait(&ofull[s], (it / W2_NO) & 1);
      tc::tc_fence_after();
      dbg_mark(dbgm, 1, dnm, 12);
      const uint32_t sD = base + s * STAGE, sX = sD + 2 * PSD;
      const uint64_t b_hi = tc::smem_desc(sX, 128, W2_CS), b_lo = tc::smem_desc(sX + PSX, 128, W2_CS);
      if (tc::elect_one()) {
        for (int mt = 0; mt < nmt; ++mt) {
          const int mrows = min(128, Nout - mt * 128);          // 128 or 64
          const uint32_t idesc = tc::idesc_bf16(mrows, Kin, 1, 1);
          const uint32_t d = tmem + (uint32_t)(mt * Kin);
          // A = dy, B = x (both MN-major, K = tokens): hi*hi, hi*lo, lo*hi
          const uint64_t a_hi = tc::smem_desc(sD + (uint32_t)(mt * 16) * W2_CS, 128, W2_CS);
          const uint64_t a_lo = tc::smem_desc(sD + PSD + (uint32_t)(mt * 16) * W2_CS, 128, W2_CS);
#pragma unroll
          for (int ks = 0; ks < WT2 / 16; ++ks) {
            const uint64_t dk = (uint64_t)(ks * 16);            // 16 token rows = 256 bytes = 16 address units
            tc::mma_bf16(d, a_hi + dk, b_hi + dk, idesc, (it | ks) != 0);
            if (NTERMS > 1) {
              tc::mma_bf16(d, a_hi + dk, b_lo + dk, idesc, 1);
              tc::mma_bf16(d, a_lo + dk, b_hi + dk, idesc, 1);
            }
          }
        }
        tc::tc_commit(&oempty[s]);
        if (it == nst - 1) tc::tc_commit(accfull);
      }
      __syncwarp();
      dbg_mark(dbgm, 1, dnm, 13);
    }
    if (a.late_trigger) pdl_trigger();    // experiment (g_tune[2] & 2): dependents may launch once every CTA has issued its last MMA
  } else {
    // ===== producers =====
    unsigned long long* dbgp = (blockIdx.x == 0 && blockIdx.y == 0 && warp == 3 && lane == 0) ? a.dbg : nullptr;
    int dnp = 0;
    dbg_mark(dbgp, 0, dnp, 0);
    const int fd = tid % n4d, sd = tid / n4d;          // dy: float4 group of a row, row slot (active if sd < rpp_d)
    const int fx = tid % n4x, sx = tid / n4x;          // x : likewise
    const int pro_dy = PDY < 0 ? a.pro_dy : PDY, pro_x = PX < 0 ? a.pro_x : PX;
    float cs[4] = {0.f, 0.f, 0.f, 0.f};                // column sums of this thread's 4 dy features
    auto put = [&](uint8_t* plane_hi, uint32_t PS, int f4, int row, const float4& q) {
      uint2 h, l;
      tc::split2(q.x, q.y, h.x, l.x);
      tc::split2(q.z, q.w, h.y, l.y);
      uint8_t* d = plane_hi + (uint32_t)(f4 >> 1) * W2_CS + row * 16 + (f4 & 1) * 8;
      *reinterpret_cast<uint2*>(d) = h;
      if (NTERMS > 1) *reinterpret_cast<uint2*>(d + PS) = l;
    };
    for (int it = 0; it < nst; ++it) {
      const int ls = it % W2_NL, s = it % W2_NO;
      const long r0 = (long)(st_beg + it) * WT2;
      const float4* Ld = reinterpret_cast<const float4*>(sL + ls * LSTAGE);
      const float4* Lx = reinterpret_cast<const float4*>(sL + ls * LSTAGE + LD_BYTES);
      uint8_t* sb = sO + s * STAGE;
      tc::mbar_wait(&lfull[ls], (it / W2_NL) & 1);
      dbg_mark(dbgp, 0, dnp, 1);
      tc::mbar_wait(&oempty[s], ((it / W2_NO) & 1) ^ 1);
      dbg_mark(dbgp, 0, dnp, 2);
      if (sd < rpp_d) {
        for (int row = sd; row < WT2; row += rpp_d) {
          float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r0 + row < a.M) {
            q = Ld[row * n4d + fd];
            if (pro_dy != PRO_NONE) {
              const float4 m = drop_mult4(a.drop_dy, (uint64_t)(r0 + row) * (uint64_t)Nout + (uint64_t)(fd * 4));
              if (pro_dy == PRO_GELU_DROP) { q.x = gelu_f(q.x); q.y = gelu_f(q.y); q.z = gelu_f(q.z); q.w = gelu_f(q.w); }
              q.x *= m.x; q.y *= m.y; q.z *= m.z; q.w *= m.w;
            }
          }
          if (want_db) { cs[0] += q.x; cs[1] += q.y; cs[2] += q.z; cs[3] += q.w; }
          put(sb, PSD, fd, row, q);
        }
      }
      if (sx < rpp_x) {
        for (int row = sx; row < WT2; row += rpp_x) {
          float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r0 + row < a.M) {
            q = Lx[row * n4x + fx];
            if (pro_x != PRO_NONE) {
              const float4 m = drop_mult4(a.drop_x, (uint64_t)(r0 + row) * (uint64_t)Kin + (uint64_t)(fx * 4));
              if (pro_x == PRO_GELU_DROP) { q.x = gelu_f(q.x); q.y = gelu_f(q.y); q.z = gelu_f(q.z); q.w = gelu_f(q.w); }
              q.x *= m.x; q.y *= m.y; q.z *= m.z; q.w *= m.w;
            }
          }
          put(sb + 2 * PSD, PSX, fx, row, q);
        }
      }
      tc::fence_async_smem();
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive(&ofull[s]); tc::mbar_arrive(&lempty[ls]); }
      dbg_mark(dbgp, 0, dnp, 3);
    }
    if (want_db && sd < rpp_d) {
      float* o = sCol + sd * Nout + fd * 4;
      o[0] = cs[0]; o[1] = cs[1]; o[2] = cs[2]; o[3] = cs[3];
    }
  }
  __syncthreads();   // column sums staged; all roles done issuing
  // blockIdx.y selects a Kin-wide column block of x (wide layers: one launch covers all blocks)
  float* part = a.partial + ((long)blockIdx.y * gridDim.x + blockIdx.x) * ((long)Nout * Kin + Nout);
  if (warp < 4) {
    // ===== epilogue: accumulators -> partial[cta][n][k] =====
    const int q = warp;
    unsigned long long* dbge = (blockIdx.x == 0 && blockIdx.y == 0 && warp == 0 && lane == 0) ? a.dbg : nullptr;
    int dne = 0;
    if (nst > 0) {
      tc::mbar_wait(accfull, 0);
      tc::tc_fence_after();
    }
    for (int mt = 0; mt < nmt; ++mt) {
      const int mrows = min(128, Nout - mt * 128);
      // M = 128: row = q*32 + lane ; M = 64: row = q*16 + lane (lanes 0-15 of each sub-partition)
      const int row = mrows == 128 ? q * 32 + lane : q * 16 + lane;
      const bool valid = mrows == 128 || lane < 16;
      for (int cb = 0; cb < Kin; cb += 32) {
        float v[32];
        if (nst > 0) tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * Kin + cb), v);
        else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (valid) {
          float* o = part + (long)(mt * 128 + row) * Kin + cb;
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
    if (want_db) {
      float* pb = part + (long)Nout * Kin;
      for (int n = tid; n < Nout; n += 128) {
        float sum = 0.f;
        if (nst > 0)
#pragma unroll
          for (int g = 0; g < rpp_d; ++g) sum += sCol[g * Nout + n];
        pb[n] = sum;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == W2_MMA_WARP) tc::tmem_dealloc(tmem, ncols);
}


// dW (split over up to 3 destinations of rows_per_dst rows each) = sum over CTAs of the partials, fixed order.
struct WgradReduceArgs {
  const float* partial; int ctas, Nout, Kin, rows_per_dst;   // blockIdx.y = Kin block: partial and dW advance per block
  long ldw;                       // row stride of the destination(s) (>= Kin: column blocks of a wider dW)
  const float* log_scale;         // optional device scalar: results are multiplied by exp(*log_scale)
  float* dW[4]; float* db[4];     // destination d covers rows [d*rows_per_dst, (d+1)*rows_per_dst); null = discard (padding rows)
};
__device__ __forceinline__ void wgrad_reduce_body(const WgradReduceArgs& a, int bx, int by) {
  // 32 consecutive outputs x 8 partial groups per CTA; 8 independent loads in flight per thread (latency-bound otherwise)
  __shared__ float sh[8][33];
  const int total = a.Nout * a.Kin + a.Nout;
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = bx * 32 + lane;
  float s = 0.f;
  if (i < total) {
    const float* p = a.partial + (long)by * a.ctas * total + i;
    int c = g;
    for (; c + 56 < a.ctas; c += 64) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = p[(long)(c + 8 * u) * total];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; c < a.ctas; c += 8) s += p[(long)c * total];
  }
  sh[g][lane] = s;
  __syncthreads();
  if (g == 0 && i < total) {
    s = ((sh[0][lane] + sh[1][lane]) + (sh[2][lane] + sh[3][lane])) + ((sh[4][lane] + sh[5][lane]) + (sh[6][lane] + sh[7][lane]));
    if (a.log_scale) s *= __expf(*a.log_scale);
    if (i < a.Nout * a.Kin) {
      const int n = i / a.Kin, k = i - n * a.Kin;
      const int d = n / a.rows_per_dst;
      if (d < 4 && a.dW[d]) a.dW[d][(long)(n - d * a.rows_per_dst) * a.ldw + (long)by * a.Kin + k] = s;
    } else {
      const int n = i - a.Nout * a.Kin;
      const int d = n / a.rows_per_dst;
      if (d < 4 && a.db[d] && by == 0) a.db[d][n - d * a.rows_per_dst] = s;
    }
  }
}
static __global__ void __launch_bounds__(256) lin_wgrad_reduce_kernel(const WgradReduceArgs a) {
  pdl_sync();
  wgrad_reduce_body(a, blockIdx.x, blockIdx.y);
}
// Several reductions in one launch (blockIdx.z = job): the four weight gradients of a transformer block write their partials to
// separate regions and are folded together at the end of the block's backward (one launch instead of four ~10 us ones).
constexpr int MAX_REDUCE_JOBS = 6;
struct WgradReduceBatch { WgradReduceArgs j[MAX_REDUCE_JOBS]; int kin_blocks[MAX_REDUCE_JOBS]; int n; };
static __global__ void __launch_bounds__(256) lin_wgrad_reduce_batch_kernel(const WgradReduceBatch b) {
  pdl_sync();
  const WgradReduceArgs& a = b.j[blockIdx.z];
  const int total = a.Nout * a.Kin + a.Nout;
  if ((int)blockIdx.y >= b.kin_blocks[blockIdx.z] || (int)blockIdx.x * 32 >= total) return;
  wgrad_reduce_body(a, blockIdx.x, blockIdx.y);
}
inline int lin_wgrad_reduce_flush(WgradReduceBatch& b, cudaStream_t st) {
  if (b.n <= 0) return EEGCLIP_OK;
  int gx = 1, gy = 1;
  for (int i = 0; i < b.n; ++i) {
    const int total = b.j[i].Nout * b.j[i].Kin + b.j[i].Nout;
    gx = max(gx, (total + 31) / 32);
    gy = max(gy, b.kin_blocks[i]);
  }
  LAUNCH_PDL((lin_wgrad_reduce_batch_kernel), dim3(gx, gy, b.n), 256, 0, st, b);
  LAUNCH_CHECK();
  b.n = 0;
  return EEGCLIP_OK;
}

inline bool lin_wgrad_tc_supported(long M, int Nout, int Kin) {
  return M >= 1 && (Nout == 64 || Nout == 128 || Nout == 192 || Nout == 256) && (Kin == 64 || Kin == 128 || Kin == 192 || Kin == 256) &&
         ((Nout + 127) / 128) * Kin <= 512 && (Nout + Kin) / KC <= WG_MAXC && wgrad_lin_smem_bytes(Nout, Kin) <= 227u * 1024u;
}
// CTAs along the token dimension: all SMs for a single column block, ~2 waves in total when blockIdx.y multiplies the grid
inline int wgrad_token_ctas(int nst, int kin_blocks) {
  int ctas = nst < WG_MAX_CTAS ? nst : WG_MAX_CTAS;
  if (kin_blocks > 1) ctas = max(1, min(ctas, (2 * WG_MAX_CTAS) / kin_blocks));
  return ctas;
}
inline size_t lin_wgrad_partial_bytes(int Nout, int Kin, int kin_blocks = 1) {
  return (size_t)kin_blocks * wgrad_token_ctas(WG_MAX_CTAS, kin_blocks) * ((size_t)Nout * Kin + Nout) * sizeof(float);
}

template <int NTERMS, int PDY, int PX, int WDB>
inline int lin_wgrad_launch_v(const LinWgradArgs& a, dim3 grid, uint32_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(lin_wgrad_tc_kernel<NTERMS, PDY, PX, WDB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return EEGCLIP_ERR_CUDA;
    configured = true;
  }
  LAUNCH_PDL((lin_wgrad_tc_kernel<NTERMS, PDY, PX, WDB>), grid, WG_THREADS, smem, st, a);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

// the bulk-copy kernel: dense operands arrive as one 1-D bulk copy per stage, strided ones through a 2-D tensor map;
// g_tune[9] = 1 forces the register-staged kernel, g_tune[12] = 1 the per-row copies (A/B timing)
inline bool lin_wgrad_tma_ok(const LinWgradArgs& a) {
  return g_tune[9] == 0 && ((uintptr_t)a.dy & 15) == 0 && ((uintptr_t)a.x & 15) == 0 && (a.lddy & 3) == 0 && (a.ldx & 3) == 0 &&
         wgrad2_smem_bytes(a.Nout, a.Kin) <= 227u * 1024u;
}
template <int NTERMS, int PDY, int PX, int WDB>
inline int lin_wgrad_tma_launch_v(const LinWgradArgs& a, dim3 grid, cudaStream_t st, const CUtensorMap& tm_dy, const CUtensorMap& tm_x, int use_tm) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(lin_wgrad_tma_kernel<NTERMS, PDY, PX, WDB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return EEGCLIP_ERR_CUDA;
    configured = true;
  }
  LinWgradArgs al = a;
  al.late_trigger = (g_tune[2] >> 1) & 1;
  LAUNCH_PDL((lin_wgrad_tma_kernel<NTERMS, PDY, PX, WDB>), grid, W2_THREADS, wgrad2_smem_bytes(a.Nout, a.Kin), st, al, tm_dy, tm_x, use_tm);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

template <int NTERMS>
inline int lin_wgrad_launch_t(LinWgradArgs a, float* const dW[3], float* const db[3], int rows_per_dst, long ldw, cudaStream_t st,
                              const float* log_scale = nullptr, int kin_blocks = 1, WgradReduceBatch* defer = nullptr) {
  if (!lin_wgrad_tc_supported(a.M, a.Nout, a.Kin)) return EEGCLIP_ERR_UNSUPPORTED;
  const int nst = (a.M + WT - 1) / WT;
  const int ctas = wgrad_token_ctas(nst, kin_blocks);
  a.want_db = (db[0] != nullptr) ? 1 : 0;
  a.policy = g_tune[1];
  a.dbg = g_dbg_buf;
  {
    ProfScope prof(PROF_LIN_WGRAD, st);
    const dim3 grid(ctas, kin_blocks);
    const uint32_t smem = wgrad_lin_smem_bytes(a.Nout, a.Kin);
    int rc;
    // strided operands: one tensor map each (host-side encode, no device work); both strided without maps -> register-staged kernel
    alignas(64) CUtensorMap tm_dy, tm_x;
    memset(&tm_dy, 0, sizeof(tm_dy)); memset(&tm_x, 0, sizeof(tm_x));
    int use_tm = 0;
    bool tma = lin_wgrad_tma_ok(a);
    if (tma) {
      const bool dense_dy = a.lddy == a.Nout, dense_x = a.ldx == a.Kin && kin_blocks == 1;
      if (g_tune[12] == 0) {
        if (!dense_dy && tc::make_tmap_2d_f32(&tm_dy, a.dy, (uint64_t)a.Nout, (uint64_t)a.M, (uint64_t)a.lddy, (uint32_t)a.Nout, WT2) == 0) use_tm |= 1;
        if (!dense_x && tc::make_tmap_2d_f32(&tm_x, a.x, (uint64_t)a.Kin * kin_blocks, (uint64_t)a.M, (uint64_t)a.ldx, (uint32_t)a.Kin, WT2) == 0) use_tm |= 2;
      }
      const bool row_dy = !dense_dy && !(use_tm & 1), row_x = !dense_x && !(use_tm & 2);
      if (row_dy && row_x) tma = false;                    // per-row copies of BOTH operands: 64 requests per stage
    }
    if (tma) {
      if (a.pro_x == PRO_NONE && a.pro_dy == PRO_NONE && !a.want_db) rc = lin_wgrad_tma_launch_v<NTERMS, PRO_NONE, PRO_NONE, 0>(a, grid, st, tm_dy, tm_x, use_tm);
      else if (a.pro_x == PRO_NONE && a.pro_dy == PRO_NONE && a.want_db) rc = lin_wgrad_tma_launch_v<NTERMS, PRO_NONE, PRO_NONE, 1>(a, grid, st, tm_dy, tm_x, use_tm);
      else if (a.pro_x == PRO_NONE && a.pro_dy == PRO_DROP && a.want_db) rc = lin_wgrad_tma_launch_v<NTERMS, PRO_DROP, PRO_NONE, 1>(a, grid, st, tm_dy, tm_x, use_tm);
      else rc = lin_wgrad_tma_launch_v<NTERMS, -1, -1, -1>(a, grid, st, tm_dy, tm_x, use_tm);
    } else
    if (g_tune[3] == 0 && a.pro_x == PRO_NONE && a.pro_dy == PRO_NONE && !a.want_db) rc = lin_wgrad_launch_v<NTERMS, PRO_NONE, PRO_NONE, 0>(a, grid, smem, st);
    else if (g_tune[3] == 0 && a.pro_x == PRO_NONE && a.pro_dy == PRO_NONE && a.want_db) rc = lin_wgrad_launch_v<NTERMS, PRO_NONE, PRO_NONE, 1>(a, grid, smem, st);
    else if (g_tune[3] == 0 && a.pro_x == PRO_NONE && a.pro_dy == PRO_DROP && a.want_db) rc = lin_wgrad_launch_v<NTERMS, PRO_DROP, PRO_NONE, 1>(a, grid, smem, st);
    else rc = lin_wgrad_launch_v<NTERMS, -1, -1, -1>(a, grid, smem, st);
    if (rc != EEGCLIP_OK) return rc;
  }
  WgradReduceArgs r;
  r.partial = a.partial; r.ctas = ctas; r.Nout = a.Nout; r.Kin = a.Kin; r.rows_per_dst = rows_per_dst; r.ldw = ldw > 0 ? ldw : a.Kin; r.log_scale = log_scale;
  for (int i = 0; i < 3; ++i) { r.dW[i] = dW[i]; r.db[i] = db[i]; }
  r.dW[3] = nullptr; r.db[3] = nullptr;
  if (defer && defer->n < MAX_REDUCE_JOBS) {          // the caller folds this reduction into a later batched launch
    defer->j[defer->n] = r; defer->kin_blocks[defer->n] = kin_blocks; ++defer->n;
    return EEGCLIP_OK;
  }
  const int total = a.Nout * a.Kin + a.Nout;
  LAUNCH_PDL((lin_wgrad_reduce_kernel), dim3((total + 31) / 32, kin_blocks), 256, 0, st, r);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}
inline int lin_wgrad_launch(int math, const LinWgradArgs& a, float* const dW[3], float* const db[3], int rows_per_dst, cudaStream_t st,
                            long ldw = 0, const float* log_scale = nullptr, int kin_blocks = 1, WgradReduceBatch* defer = nullptr) {
  return math == EEGCLIP_MATH_BF16 ? lin_wgrad_launch_t<1>(a, dW, db, rows_per_dst, ldw, st, log_scale, kin_blocks, defer)
                                   : lin_wgrad_launch_t<3>(a, dW, db, rows_per_dst, ldw, st, log_scale, kin_blocks, defer);
}

// ------------------------------------------------------------------------------------------------
// Generic linear layers on the kernels above (nn.Linear / 1x1 Conv1d with N in {64,128,192,256}, K % 64 == 0):
// K wider than the resident-weight budget is split into passes that accumulate through the residual input
// (speech tower: 1024 -> 64 in four K = 256 passes).
// ------------------------------------------------------------------------------------------------
inline int lin_pass_width(int N, int K) {
  int kp = K < 256 ? K : 256;
  while (kp > KC && ((long)N * kp > 16384 || K % kp)) kp -= KC;
  return kp;
}
inline bool linear_tc_ok(long M, int N, int K) {
  if (!(N == 64 || N == 128 || N == 192 || N == 256) || K < KC || (K % KC)) return false;
  const int kp = lin_pass_width(N, K);
  return (long)N * kp <= 16384 && K % kp == 0 && lin_tc_supported(M, N, kp);
}
inline size_t linear_tc_scratch_bytes(int N, int K) {
  size_t a = packed_bytes(N, K) + 256;                                     // packed weights (all passes)
  size_t b = lin_wgrad_partial_bytes(N, K < 256 ? K : 256, K < 256 ? 1 : K / 256) + 256;   // weight-gradient partials
  size_t c = packed_bytes(K, N) + 256;                                     // transposed pack for the data gradient
  return a + b + c;
}

// out[m][n] = sum_k x[m][k] W[n][k] + b[n]
inline int linear_tc_fwd(int math, const float* x, long ldx, const float* W, const float* b, float* out, long ldo, long M, int N, int K,
                         uint8_t* wp, cudaStream_t st) {
  const int kp = lin_pass_width(N, K), npass = K / kp;
  for (int p0 = 0; p0 < npass; p0 += MAX_PACK_JOBS) {
    PackJobs J; J.n = 0;
    for (int p = p0; p < npass && p < p0 + MAX_PACK_JOBS; ++p)
      add_pack(J, W + (long)p * kp, wp + (size_t)p * packed_bytes(N, kp), N, kp, 0, 0, N, kp, K, 1);
    int rc = pack_launch(J, st);
    if (rc != EEGCLIP_OK) return rc;
  }
  for (int p = 0; p < npass; ++p) {
    LinTcArgs a{};
    a.A = x + (long)p * kp; a.lda = ldx; a.wpacked = wp + (size_t)p * packed_bytes(N, kp); a.C = out; a.ldc = ldo;
    a.M = (int)M; a.N = N; a.K = kp;
    a.pro = PRO_NONE; a.pro_drop = make_drop(0, 0, 0, 0.f, 0); a.drop = a.pro_drop;
    a.bias = p == 0 ? b : nullptr;
    a.residual = p == 0 ? nullptr : out;
    int rc = lin_tc_launch(math, a, st);
    if (rc != EEGCLIP_OK) return rc;
  }
  return EEGCLIP_OK;
}

// dW[n][k] = sum_m dy[m][n] x[m][k] ; db[n] = sum_m dy[m][n]   (overwrites)
inline int linear_tc_wgrad(int math, const float* dy, long lddy, const float* x, long ldx, float* dW, float* db, long M, int N, int K,
                           float* partial, cudaStream_t st) {
  const int kb = K < 256 ? K : 256;
  LinWgradArgs a{};
  a.dy = dy; a.lddy = lddy; a.Nout = N; a.x = x; a.ldx = ldx; a.Kin = kb; a.M = (int)M;
  a.drop_dy = make_drop(0, 0, 0, 0.f, 0); a.drop_x = a.drop_dy; a.partial = partial;
  float* dWs[3] = {dW, nullptr, nullptr};
  float* dbs[3] = {db, nullptr, nullptr};
  return lin_wgrad_launch(math, a, dWs, dbs, N, st, K, nullptr, K / kb);
}
inline bool linear_tc_wgrad_ok(long M, int N, int K) {
  const int kb = K < 256 ? K : 256;
  return (K % kb) == 0 && lin_wgrad_tc_supported(M, N, kb);
}

// dx[m][k] = sum_n dy[m][n] W[n][k]      (output width K in {64..256}, contraction N)
inline bool linear_tc_dgrad_ok(long M, int N, int K) { return linear_tc_ok(M, K, N) && lin_pass_width(K, N) == N; }
inline int linear_tc_dgrad(int math, const float* dy, long lddy, const float* W, float* dx, long lddx, long M, int N, int K, uint8_t* wp,
                           cudaStream_t st) {
  PackJobs J; J.n = 0;
  add_pack(J, W, wp, K, N, 0, 0, K, N, 1, K);        // operand (n' = k, k' = n) = W[n][k]
  int rc = pack_launch(J, st);
  if (rc != EEGCLIP_OK) return rc;
  LinTcArgs a{};
  a.A = dy; a.lda = lddy; a.wpacked = wp; a.C = dx; a.ldc = lddx; a.M = (int)M; a.N = K; a.K = N;
  a.pro = PRO_NONE; a.pro_drop = make_drop(0, 0, 0, 0.f, 0); a.drop = a.pro_drop;
  return lin_tc_launch(math, a, st);
}

}  // namespace lintc
}  // namespace eegclip
