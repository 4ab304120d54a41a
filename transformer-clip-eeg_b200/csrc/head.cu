// Contrastive head: L2 normalise, symmetric InfoNCE (local or sharded), memory bank, AdamW, match-mismatch scoring.
// Replaces clip_model.py:675-693 / 909-944 (loss), :731-745 (memoryBank), train_clip_final.py:409-413,492 (AdamW),
// train_clip_helper_functions.py:153-163,176-187 (scoring).
#include "../../include/eegclip.h"
#include "common.cuh"
#include "gemm_f32.cuh"
#include "lin_tc.cuh"
#include "head_tc.cuh"

using namespace eegclip;

#define TRY(x) do { int _r = (x); if (_r != EEGCLIP_OK) return _r; } while (0)

// ---- launch accounting / per-kernel-class timing state (declared in common.cuh) ---------------------------
namespace eegclip {
long long g_launch_count = 0;
int g_tune[16] = {0};
thread_local int g_pdl_break = 0;
unsigned long long* g_dbg_buf = nullptr;
constexpr int PROF_MAX = 8192;
static bool g_prof_on = false;
static int g_prof_n = 0;
static cudaEvent_t g_prof_ev[PROF_MAX][2];
static int g_prof_cls[PROF_MAX];
static bool g_prof_init = false;
// g_tune[8]: bit mask of the kernel classes that record events (0 = all).  bench.py times its steps with the roofline
// kernel's class only, so that the event records do not break up programmatic dependent launch between the other kernels.
static inline bool prof_cls_on(int cls) { return g_tune[8] == 0 || ((g_tune[8] >> cls) & 1); }
void prof_begin(int cls, cudaStream_t st) {
  if (!g_prof_on || g_prof_n >= PROF_MAX || !prof_cls_on(cls)) return;
  g_prof_cls[g_prof_n] = cls;
  cudaEventRecord(g_prof_ev[g_prof_n][0], st);
}
void prof_end(int cls, cudaStream_t st) {
  if (!g_prof_on || g_prof_n >= PROF_MAX || !prof_cls_on(cls)) return;
  cudaEventRecord(g_prof_ev[g_prof_n][1], st);
  ++g_prof_n;
}
}  // namespace eegclip

namespace {

// ---- L2 normalise rows: xn = x / max(||x||, 1e-12)  (F.normalize, clip_model.py:675-676) -----------
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ xn,
                                                        float* __restrict__ inv_norm, int D) {
  pdl_sync();
  __shared__ float2 sh[33];
  const long row = blockIdx.x;
  const float* xr = x + row * D;
  float s = 0.f;
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(xr + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  float2 r = block_sum2(s, 0.f, sh);
  float inv = 1.f / fmaxf(sqrtf(r.x), 1e-12f);
  if (threadIdx.x == 0 && inv_norm) inv_norm[row] = inv;
  float* o = xn + row * D;
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(xr + i);
    *reinterpret_cast<float4*>(o + i) = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
  }
}

// dx = inv * (dxn - xn * <xn, dxn>)
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* __restrict__ xn, const float* __restrict__ inv_norm,
                                                        const float* __restrict__ dxn, float* __restrict__ dx, int D) {
  pdl_sync();
  __shared__ float2 sh[33];
  const long row = blockIdx.x;
  const float* a = xn + row * D;
  const float* g = dxn + row * D;
  float s = 0.f;
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(a + i), w = *reinterpret_cast<const float4*>(g + i);
    s += v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
  }
  float2 r = block_sum2(s, 0.f, sh);
  const float dot = r.x, inv = inv_norm[row];
  float* o = dx + row * D;
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(a + i), w = *reinterpret_cast<const float4*>(g + i);
    *reinterpret_cast<float4*>(o + i) =
        make_float4(inv * (w.x - v.x * dot), inv * (w.y - v.y * dot), inv * (w.z - v.z * dot), inv * (w.w - v.w * dot));
  }
}

// combine per-tile (max,sumexp) partials into a log-sum-exp per row
__global__ void lse_combine_kernel(const float2* __restrict__ part, int ntiles, float* __restrict__ lse, int rows) {
  pdl_sync();
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float m = -INFINITY;
  for (int t = 0; t < ntiles; ++t) m = fmaxf(m, part[(long)r * ntiles + t].x);
  float s = 0.f;
  for (int t = 0; t < ntiles; ++t) { float2 p = part[(long)r * ntiles + t]; s += p.y * __expf(p.x - m); }
  lse[r] = m + __logf(s);
}

// diag[i] = exp(tau) * <S[row0+i], E[row0+i]>   (one warp per row)
__global__ void diag_kernel(const float* __restrict__ S, const float* __restrict__ E, const float* __restrict__ tau,
                            float* __restrict__ diag, int b, int row0, int D) {
  pdl_sync();
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
  if (w >= b) return;
  const float* s = S + (long)(row0 + w) * D;
  const float* e = E + (long)(row0 + w) * D;
  float a = 0.f;
  for (int i = l * 4; i < D; i += 128) {
    float4 x = *reinterpret_cast<const float4*>(s + i), y = *reinterpret_cast<const float4*>(e + i);
    a += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
  }
  a = warp_sum(a);
  if (l == 0) diag[w] = a * __expf(*tau);
}

__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ lr, const float* __restrict__ lc,
                                                  const float* __restrict__ dg, int Bg, int one_sided, float* __restrict__ loss) {
  pdl_sync();
  __shared__ float2 sh[33];
  float s = 0.f;
  for (int i = threadIdx.x; i < Bg; i += blockDim.x) s += one_sided ? (lr[i] - dg[i]) : (lr[i] - dg[i]) + (lc[i] - dg[i]);
  float2 r = block_sum2(s, 0.f, sh);
  if (threadIdx.x == 0) *loss = r.x / ((one_sided ? 1.f : 2.f) * (float)Bg);
}

// memoryBank.forward (clip_model.py:731-745) in two launches, so that every old row is read before any row is written (the
// reference gathers with index_select before index_copy_): duplicate ids in one batch all see the pre-batch row, and the
// LAST occurrence of an id is the one whose update lands (index_copy_ on CPU writes in order).  Ids outside [0, bank_rows)
// -- an IndexError in the reference -- write nothing and return NaN rows, so the loss turns NaN instead of corrupting memory.
__global__ void __launch_bounds__(256) membank_gather_kernel(const float* __restrict__ mem, const int64_t* __restrict__ idx,
                                                            float* __restrict__ old_out, int D, long bank_rows) {
  pdl_sync();
  const long r = blockIdx.x;
  const long id = idx[r];
  const bool ok = id >= 0 && id < bank_rows;
  const float* m = mem + (ok ? id : 0) * (long)D;
  float* o = old_out + r * D;
  const float nan = __int_as_float(0x7fc00000);
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(o + i) = ok ? *reinterpret_cast<const float4*>(m + i) : make_float4(nan, nan, nan, nan);
}

__global__ void __launch_bounds__(256) membank_scatter_kernel(float* __restrict__ mem, const int64_t* __restrict__ idx,
                                                             const float* __restrict__ data, const float* __restrict__ old,
                                                             int rows, int D, long bank_rows, float momentum, float om) {
  pdl_sync();
  __shared__ int later;
  const int r = blockIdx.x;
  const long id = idx[r];
  if (id < 0 || id >= bank_rows) return;
  if (threadIdx.x == 0) later = 0;
  __syncthreads();
  for (int j = r + 1 + threadIdx.x; j < rows; j += blockDim.x)
    if (idx[j] == id) later = 1;                       // a later row of this batch carries the same id: it wins
  __syncthreads();
  if (later) return;
  float* m = mem + id * (long)D;
  const float* d = data + (long)r * D;
  const float* o = old + (long)r * D;
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    float4 a = *reinterpret_cast<const float4*>(o + i), x = *reinterpret_cast<const float4*>(d + i);
    // new = old*momentum + data*(1-momentum), same operation order as the reference (mul_ then add_)
    *reinterpret_cast<float4*>(m + i) = make_float4(a.x * momentum + x.x * om, a.y * momentum + x.y * om,
                                                    a.z * momentum + x.z * om, a.w * momentum + x.w * om);
  }
}

// AdamW / Adam, one launch for all tensors: blockIdx.y = tensor, grid-stride over its elements.  torch.optim semantics:
// decoupled decay (AdamW) multiplies the parameter by (1 - lr*wd); coupled decay (Adam) adds wd*p to the gradient;
// amsgrad (entry.vmax != NULL) keeps the running maximum of exp_avg_sq and divides by it.
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float* vmax, float lr, float b1, float b2, float eps,
                                            float wd, int coupled, float step_size, float bc2_sqrt) {
  if (coupled) g += wd * p; else p *= (1.f - lr * wd);
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  float vv = v;
  if (vmax) { vv = fmaxf(*vmax, v); *vmax = vv; }
  p -= step_size * m / (sqrtf(vv) / bc2_sqrt + eps);
}

__global__ void __launch_bounds__(256) adamw_kernel(const eegclip_adamw_entry* __restrict__ tab, float lr, float b1, float b2,
                                                   float eps, float wd, int coupled, float bc1, float bc2_sqrt) {
  pdl_sync();
  const eegclip_adamw_entry e = tab[blockIdx.y];
  float* p = (float*)e.p; const float* g = (const float*)e.g; float* m = (float*)e.m; float* v = (float*)e.v;
  float* vm = (float*)e.vmax;
  const long n = e.numel;
  const long stride = (long)gridDim.x * blockDim.x;
  const float step_size = lr / bc1;
  const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)vm) & 15) == 0);
  const long n4 = vec ? (n >> 2) : 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<const float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    float4 X = vm ? reinterpret_cast<float4*>(vm)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float pp[4] = {P.x, P.y, P.z, P.w}, gg[4] = {G.x, G.y, G.z, G.w}, mm[4] = {M.x, M.y, M.z, M.w}, vv[4] = {V.x, V.y, V.z, V.w};
    float xx[4] = {X.x, X.y, X.z, X.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) adam_update(pp[j], gg[j], mm[j], vv[j], vm ? &xx[j] : nullptr, lr, b1, b2, eps, wd, coupled, step_size, bc2_sqrt);
    reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (vm) reinterpret_cast<float4*>(vm)[i] = make_float4(xx[0], xx[1], xx[2], xx[3]);
  }
  for (long i = n4 * 4 + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float P = p[i], M = m[i], V = v[i];
    adam_update(P, g[i], M, V, vm ? vm + i : nullptr, lr, b1, b2, eps, wd, coupled, step_size, bc2_sqrt);
    p[i] = P; m[i] = M; v[i] = V;
  }
}

// scores[k][n] = <eeg[n], cand[n][k]> and choice[n] = argmax_k (first max wins, as torch.argmax)
__global__ void __launch_bounds__(128) mm_rowdots_kernel(const float* __restrict__ eeg, const float* __restrict__ cand,
                                                        float* __restrict__ scores, int64_t* __restrict__ choice, int N, int K,
                                                        int D) {
  __shared__ float2 sh[33];
  const int n = blockIdx.x;
  const float* e = eeg + (long)n * D;
  float best = -INFINITY;
  int bi = 0;
  for (int k = 0; k < K; ++k) {
    const float* c = cand + ((long)n * K + k) * D;
    float a = 0.f;
    for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
      float4 x = *reinterpret_cast<const float4*>(e + i), y = *reinterpret_cast<const float4*>(c + i);
      a += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
    float2 r = block_sum2(a, 0.f, sh);
    if (threadIdx.x == 0) scores[(long)k * N + n] = r.x;
    if (r.x > best) { best = r.x; bi = k; }
  }
  if (threadIdx.x == 0 && choice) choice[n] = bi;
}

}  // namespace

extern "C" {

int eegclip_abi_version(void) { return EEGCLIP_ABI_VERSION; }

long long eegclip_launch_count(void) { return eegclip::g_launch_count; }

int eegclip_debug_buffer(void* dev_ptr) {
  eegclip::g_dbg_buf = (unsigned long long*)dev_ptr;
  return EEGCLIP_OK;
}

int eegclip_tune_set(int32_t key, int32_t value) {
  if (key < 0 || key >= 16) return EEGCLIP_ERR_ARG;
  eegclip::g_tune[key] = value;
  return EEGCLIP_OK;
}

int eegclip_profile_begin(void) {
  using namespace eegclip;
  if (!g_prof_init) {
    for (int i = 0; i < PROF_MAX; ++i)
      for (int j = 0; j < 2; ++j)
        if (cudaEventCreate(&g_prof_ev[i][j]) != cudaSuccess) return EEGCLIP_ERR_CUDA;
    g_prof_init = true;
  }
  g_prof_n = 0;
  g_prof_on = true;
  return EEGCLIP_OK;
}

int eegclip_profile_end(double* ms_by_class, long long* launches_by_class, int32_t n_classes) {
  using namespace eegclip;
  g_prof_on = false;
  if (!ms_by_class || !launches_by_class || n_classes < PROF_NCLASS) return EEGCLIP_ERR_ARG;
  for (int c = 0; c < n_classes; ++c) { ms_by_class[c] = 0.0; launches_by_class[c] = 0; }
  if (cudaDeviceSynchronize() != cudaSuccess) return EEGCLIP_ERR_CUDA;
  for (int i = 0; i < g_prof_n; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof_ev[i][0], g_prof_ev[i][1]) != cudaSuccess) return EEGCLIP_ERR_CUDA;
    ms_by_class[g_prof_cls[i]] += ms;
    launches_by_class[g_prof_cls[i]] += 1;
  }
  g_prof_n = 0;
  return EEGCLIP_OK;
}
const char* eegclip_build_info(void) { return "eegclip_b200 sm_100a " __DATE__ " " __TIME__; }

int eegclip_l2norm_forward(const float* x, float* xn, float* inv_norm, int32_t rows, int32_t D, void* stream) {
  if (!x || !xn || rows <= 0 || D <= 0 || (D & 3)) return EEGCLIP_ERR_ARG;
  LAUNCH_PDL((l2norm_fwd_kernel), rows, 256, 0, (cudaStream_t)stream, x, xn, inv_norm, D);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

int eegclip_l2norm_backward(const float* xn, const float* inv_norm, const float* dxn, float* dx, int32_t rows, int32_t D,
                            void* stream) {
  if (!xn || !inv_norm || !dxn || !dx || rows <= 0 || D <= 0 || (D & 3)) return EEGCLIP_ERR_ARG;
  LAUNCH_PDL((l2norm_bwd_kernel), rows, 256, 0, (cudaStream_t)stream, xn, inv_norm, dxn, dx, D);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

struct HeadScratch {
  size_t part1, part2, g1, g2, packS, packE, packST, packET, total;
};
static HeadScratch head_scratch(int b, int Bg, int D) {
  HeadScratch h;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  const size_t part = (size_t)b * 2 * ceil_div(Bg, GBN) * sizeof(float2);   // fp32 path: one per 64 columns; tcgen05 64-wide tiles: two
  h.part1 = take(part); h.part2 = take(part);
  const bool tcok = headtc::head_tc_supported(b, Bg, D);
  // G blocks: fp32 (b, Bg) on the CUDA-core path; packed bf16 hi/lo operand blocks [b/128][Bg/64] (same bytes, rounded up) on tcgen05
  const size_t gbytes = tcok ? headtc::epack_bytes(b, (int)align_up((size_t)Bg, headtc::KC)) : (size_t)b * Bg * sizeof(float);
  h.g1 = take(gbytes);
  h.g2 = take(gbytes);
  h.packS = take(tcok ? headtc::epack_bytes(Bg, D) : 0);
  h.packE = take(tcok ? headtc::epack_bytes(Bg, D) : 0);
  h.packST = take(tcok ? headtc::epack_t_bytes(Bg, D) : 0);     // S_all^T / E_all^T operand blocks of the backward contractions
  h.packET = take(tcok ? headtc::epack_t_bytes(Bg, D) : 0);
  h.total = o;
  return h;
}
// tensor-core head: D % 64 == 0, local rows start on a 128-row operand block, b splits into 64-row multiples
static bool head_use_tc(int math, int b, int row0, int Bg, int D) {
  return math != EEGCLIP_MATH_FP32 && headtc::head_tc_supported(b, Bg, D) && (row0 % headtc::RB) == 0 && (b % 64) == 0;
}

int eegclip_infonce_workspace(int32_t b, int32_t Bg, int32_t D, size_t* scratch_bytes) {
  if (b <= 0 || Bg < b || D <= 0 || !scratch_bytes) return EEGCLIP_ERR_ARG;
  *scratch_bytes = head_scratch(b, Bg, D).total;
  return EEGCLIP_OK;
}

int eegclip_infonce_lse(const float* S_all, const float* E_all, const float* tau, int32_t b, int32_t row0, int32_t Bg, int32_t D,
                        float* lse_row, float* lse_col, float* diag, int32_t math, int32_t one_sided, void* scratch, void* stream) {
  if (!S_all || !E_all || !tau || !lse_row || (!lse_col && !one_sided) || !diag || !scratch) return EEGCLIP_ERR_ARG;
  if (b <= 0 || row0 < 0 || row0 + b > Bg || D <= 0 || (D & 3)) return EEGCLIP_ERR_ARG;
  NvtxRange nvtx_("eegclip_infonce_lse");
  cudaStream_t st = (cudaStream_t)stream;
  const HeadScratch hs = head_scratch(b, Bg, D);
  char* sc = (char*)scratch;
  float2* part = (float2*)(sc + hs.part1);
  float2* part2 = (float2*)(sc + hs.part2);
  if (head_use_tc(math, b, row0, Bg, D)) {
    // ---- tcgen05: pack both gathered matrices once, then one logits pass per direction ----
    uint8_t* pS = (uint8_t*)(sc + hs.packS);
    uint8_t* pE = (uint8_t*)(sc + hs.packE);
    TRY(headtc::epack(S_all, pS, Bg, D, st));
    TRY(headtc::epack(E_all, pE, Bg, D, st));
    const int nt = 2 * ceil_div(Bg, headtc::logits_ntc(1, b, Bg));   // one (max, sum exp) partial per tile and column half
    headtc::LogitsArgs a{};
    a.Ap = pS; a.Bp = pE; a.a_blk0 = row0 / headtc::RB; a.M = b; a.N = Bg; a.D = D; a.tau = tau;
    a.m_off = row0; a.n_off = 0; a.part = part; a.diag = diag;
    TRY(headtc::logits_launch<1>(math, a, st));
    LAUNCH_PDL((lse_combine_kernel), ceil_div(b, 128), 128, 0, st, part, nt, lse_row, b);
    LAUNCH_CHECK();
    if (one_sided) return EEGCLIP_OK;
    a.Ap = pE; a.Bp = pS; a.part = part2; a.diag = nullptr;
    TRY(headtc::logits_launch<1>(math, a, st));
    LAUNCH_PDL((lse_combine_kernel), ceil_div(b, 128), 128, 0, st, part2, nt, lse_col, b);
    LAUNCH_CHECK();
    return EEGCLIP_OK;
  }
  const int ntiles = ceil_div(Bg, GBN);
  // rows: this rank's speech rows against every EEG column
  GemmArgs g;
  g.A = S_all + (long)row0 * D; g.B = E_all; g.C = nullptr;
  g.M = b; g.N = Bg; g.K = D; g.KT = D;
  g.a_ms = D; g.a_ks = 1; g.b_ks = 1; g.b_ns = D;
  g.epi.tau = tau; g.epi.lse_part = part;
  TRY(gemm_f32<1>(g, st));
  LAUNCH_PDL((lse_combine_kernel), ceil_div(b, 128), 128, 0, st, part, ntiles, lse_row, b);
  LAUNCH_CHECK();
  LAUNCH_PDL((diag_kernel), ceil_div(b * 32, 256), 256, 0, st, S_all, E_all, tau, diag, b, row0, D);
  LAUNCH_CHECK();
  if (one_sided) return EEGCLIP_OK;
  // columns: this rank's EEG rows against every speech row (the transposed block)
  g.A = E_all + (long)row0 * D; g.B = S_all;
  g.epi.lse_part = part2;
  TRY(gemm_f32<1>(g, st));
  LAUNCH_PDL((lse_combine_kernel), ceil_div(b, 128), 128, 0, st, part2, ntiles, lse_col, b);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

int eegclip_infonce_loss(const float* lse_row_all, const float* lse_col_all, const float* diag_all, int32_t Bg, int32_t one_sided,
                         float* loss, void* stream) {
  if (!lse_row_all || (!lse_col_all && !one_sided) || !diag_all || !loss || Bg <= 0) return EEGCLIP_ERR_ARG;
  LAUNCH_PDL((loss_kernel), 1, 256, 0, (cudaStream_t)stream, lse_row_all, lse_col_all, diag_all, Bg, one_sided, loss);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

int eegclip_infonce_backward(const float* S_all, const float* E_all, const float* tau, const float* lse_row_all,
                             const float* lse_col_all, int32_t b, int32_t row0, int32_t Bg, int32_t D, const float* dloss,
                             float* dS_loc, float* dE_loc, float* dtau_partial, int32_t math, int32_t one_sided, void* scratch,
                             void* stream) {
  if (!S_all || !E_all || !tau || !lse_row_all || !dE_loc || !dtau_partial || !scratch || !dloss) return EEGCLIP_ERR_ARG;
  if (!one_sided && (!lse_col_all || !dS_loc)) return EEGCLIP_ERR_ARG;
  if (b <= 0 || row0 < 0 || row0 + b > Bg || D <= 0 || (D & 3)) return EEGCLIP_ERR_ARG;
  NvtxRange nvtx_("eegclip_infonce_backward");
  cudaStream_t st = (cudaStream_t)stream;
  const HeadScratch hs = head_scratch(b, Bg, D);
  char* sc = (char*)scratch;
  float* Gr = (float*)(sc + hs.g1);   // fp32 path: (b, Bg) rows local, all columns ; tensor-core path: its transpose (Bg, b)
  float* Gc = (float*)(sc + hs.g2);   // (Bg, b): all rows, local columns
  CUDA_TRY(cudaMemsetAsync(dtau_partial, 0, sizeof(float), st));
  if (head_use_tc(math, b, row0, Bg, D)) {
    uint8_t* pS = (uint8_t*)(sc + hs.packS);
    uint8_t* pE = (uint8_t*)(sc + hs.packE);
    TRY(headtc::epack(S_all, pS, Bg, D, st));
    TRY(headtc::epack(E_all, pE, Bg, D, st));
    // Gr(i = local speech row, j = any EEG column), packed as the A operand (rows i, contraction j)  -> dS_loc = exp(tau) Gr . E_all
    const int nkc_g = ceil_div(Bg, headtc::KC);
    uint8_t* pGr = (uint8_t*)Gr;
    uint8_t* pGc = (uint8_t*)Gc;
    headtc::LogitsArgs a{};
    a.Ap = pS; a.Bp = pE; a.a_blk0 = row0 / headtc::RB; a.M = b; a.N = Bg; a.D = D; a.tau = tau;
    a.m_off = row0; a.n_off = 0; a.lse_m = lse_row_all; a.lse_n = lse_col_all; a.up = dloss; a.inv_2b = 0.5f / (float)Bg;
    a.one_sided = one_sided; a.Gp = pGr; a.nkc_g = nkc_g; a.dtau = dtau_partial;
    // d tau without float atomics: per-(CTA, epilogue warp) sums land in the (not yet written) second G buffer and are folded in
    // fixed order before the next launch overwrites it
    const int grid2 = headtc::logits_grid(a, headtc::logits_ntc(2, a.M, a.N));
    a.dtau_slots = (size_t)grid2 * 8 * sizeof(float) <= hs.g2 - hs.g1 ? (float*)pGc : nullptr;
    const bool need_ds = dS_loc && !one_sided;
    TRY(headtc::logits_launch<2>(math, a, st));   // (one-sided: only d tau comes out of this block, its G is not consumed)
    if (a.dtau_slots) TRY(headtc::dtau_fold(a.dtau_slots, grid2, dtau_partial, st));
    // Gc^T(j = local EEG column, i = any speech row): the (E_loc x S_all) product, packed with rows j and contraction i
    //   -> dE_loc = exp(tau) Gc^T . S_all.
    // One-sided (memory-bank term, clip_model.py:934-937): G(i,j) = (softmax_j L(i,.) - delta) / B, the statistics belong to
    // the rows i of X = S_all, which are the n side of this product (one_sided = 2); only E has a gradient.
    headtc::LogitsArgs c = a;
    c.Ap = pE; c.Bp = pS; c.Gp = pGc; c.dtau = nullptr; c.dtau_slots = nullptr;
    c.lse_m = one_sided ? nullptr : lse_col_all;
    c.lse_n = lse_row_all;
    c.one_sided = one_sided ? 2 : 0;
    TRY(headtc::logits_launch<2>(math, c, st));
    // the two contractions over the global batch index: out (b x D) = exp(tau) * Gpacked (b x Bg) . X_all (Bg x D), i.e. the
    // similarity kernel again with A = packed G and B = packed X^T (rows = features), dot products stored row-major
    auto contract = [&](const uint8_t* Gpk, const float* X, uint8_t* pXT, float* out) -> int {
      int rc = headtc::epack_t(X, pXT, Bg, D, st);
      if (rc != EEGCLIP_OK) return rc;
      headtc::LogitsArgs g{};
      g.Ap = Gpk; g.Bp = pXT; g.a_blk0 = 0; g.M = b; g.N = D; g.D = nkc_g * headtc::KC; g.tau = tau; g.out = out; g.ldo = D;
      return headtc::logits_launch<3>(math, g, st);
    };
    if (need_ds) TRY(contract(pGr, E_all, (uint8_t*)(sc + hs.packET), dS_loc));
    TRY(contract(pGc, S_all, (uint8_t*)(sc + hs.packST), dE_loc));
    return EEGCLIP_OK;
  }
  // G rows block, scaled by dloss * exp(tau) on the fly?  exp(tau) is a device scalar -> applied in the second GEMM.
  GemmArgs g;
  g.A = S_all + (long)row0 * D; g.B = E_all; g.C = Gr;
  g.M = b; g.N = Bg; g.K = D; g.KT = D;
  g.a_ms = D; g.a_ks = 1; g.b_ks = 1; g.b_ns = D;
  g.c_ms = Bg; g.c_ns = 1;
  g.epi.tau = tau; g.epi.lse_m = lse_row_all; g.epi.lse_n = lse_col_all; g.epi.m_off = row0; g.epi.n_off = 0;
  g.epi.inv_2b = 0.5f / (float)Bg; g.epi.alpha_dev = dloss; g.epi.one_sided = one_sided; g.epi.dtau = dtau_partial;
  TRY(gemm_f32<2>(g, st));
  const float* Gcols = Gr;
  long gc_ms = 1, gc_ks = Bg;   // A(m=j, k=i) = Gr[i][j] when the local block is the whole matrix
  if (b != Bg) {
    GemmArgs h;
    h.A = S_all; h.B = E_all + (long)row0 * D; h.C = Gc;
    h.M = Bg; h.N = b; h.K = D; h.KT = D;
    h.a_ms = D; h.a_ks = 1; h.b_ks = 1; h.b_ns = D;
    h.c_ms = b; h.c_ns = 1;
    h.epi.tau = tau; h.epi.lse_m = lse_row_all; h.epi.lse_n = lse_col_all; h.epi.m_off = 0; h.epi.n_off = row0;
    h.epi.inv_2b = 0.5f / (float)Bg; h.epi.alpha_dev = dloss; h.epi.one_sided = one_sided; h.epi.dtau = nullptr;
    TRY(gemm_f32<2>(h, st));
    Gcols = Gc; gc_ms = 1; gc_ks = b;
  }
  // dS_loc = exp(tau) * Gr . E_all      (b x D, K = Bg)
  // dE_loc = exp(tau) * Gcols^T . S_all (b x D, K = Bg)
  // exp(tau) is folded in by a tiny scale kernel-free trick: alpha is host-side, so read tau on device via MODE 0 epilogue
  // multiplier: we pre-multiply through `scale_by_exp_tau` below.
  GemmArgs s1;
  s1.A = Gr; s1.B = E_all; s1.C = dS_loc;
  s1.M = b; s1.N = D; s1.K = Bg; s1.KT = Bg;
  s1.a_ms = Bg; s1.a_ks = 1; s1.b_ks = D; s1.b_ns = 1; s1.c_ms = D; s1.c_ns = 1;
  s1.epi.tau = tau;   // MODE 0 multiplies by exp(tau) when tau != nullptr
  if (dS_loc) TRY(gemm_f32<0>(s1, st));
  GemmArgs s2;
  s2.A = Gcols; s2.B = S_all; s2.C = dE_loc;
  s2.M = b; s2.N = D; s2.K = Bg; s2.KT = Bg;
  s2.a_ms = gc_ms; s2.a_ks = gc_ks; s2.b_ks = D; s2.b_ns = 1; s2.c_ms = D; s2.c_ns = 1;
  s2.epi.tau = tau;
  TRY(gemm_f32<0>(s2, st));
  return EEGCLIP_OK;
}

int eegclip_membank_update(float* memory, int64_t bank_rows, const int64_t* idx, const float* data, float* old_out, int32_t rows,
                           int32_t D, float momentum, float one_minus_momentum, void* stream) {
  if (!memory || !idx || !data || !old_out || rows <= 0 || D <= 0 || (D & 3) || bank_rows <= 0) return EEGCLIP_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  LAUNCH_PDL((membank_gather_kernel), rows, 256, 0, st, (const float*)memory, idx, old_out, D, (long)bank_rows);
  LAUNCH_CHECK();
  LAUNCH_PDL((membank_scatter_kernel), rows, 256, 0, st, memory, idx, data, (const float*)old_out, rows, D, (long)bank_rows, momentum,
             one_minus_momentum);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

int eegclip_adamw_step(const eegclip_adamw_entry* table_dev, int32_t n_tensors, int64_t max_numel, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int32_t coupled_decay, int64_t step, void* stream) {
  if (!table_dev || n_tensors <= 0 || max_numel <= 0 || step <= 0) return EEGCLIP_ERR_ARG;
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  long per_cta = 256L * 4 * 4;
  int gx = (int)((max_numel + per_cta - 1) / per_cta);
  if (gx < 1) gx = 1;
  if (gx > 128) gx = 128;
  dim3 grid(gx, n_tensors);
  LAUNCH_PDL((adamw_kernel), grid, 256, 0, (cudaStream_t)stream, table_dev, lr, beta1, beta2, eps, weight_decay, (int)coupled_decay,
             (float)bc1, (float)sqrt(bc2));
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

int eegclip_mm_rowdots(const float* eeg, const float* cand, float* scores, int64_t* choice, int32_t N, int32_t K, int32_t D,
                       void* stream) {
  if (!eeg || !cand || !scores || N <= 0 || K <= 0 || D <= 0 || (D & 3)) return EEGCLIP_ERR_ARG;
  mm_rowdots_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(eeg, cand, scores, choice, N, K, D);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

int eegclip_mm_bank_workspace(int32_t N, int32_t M, int32_t D, size_t* scratch_bytes) {
  if (N <= 0 || M <= 0 || D <= 0 || !scratch_bytes) return EEGCLIP_ERR_ARG;
  *scratch_bytes = (D % headtc::KC) == 0 ? headtc::epack_bytes(N, D) + headtc::epack_bytes(M, D) + 512 : 256;
  return EEGCLIP_OK;
}

int eegclip_mm_bank_logits(const float* eeg, const float* bank, float* logits, int32_t N, int32_t M, int32_t D, int32_t math,
                           void* scratch, void* stream) {
  if (!eeg || !bank || !logits || N <= 0 || M <= 0 || D <= 0) return EEGCLIP_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (math != EEGCLIP_MATH_FP32 && scratch && (D % headtc::KC) == 0) {
    // tcgen05: both operands packed once, one 128 x 256 similarity tile per CTA, raw dots stored row-major
    uint8_t* pE = (uint8_t*)scratch;
    uint8_t* pB = pE + align_up(headtc::epack_bytes(N, D), 256);
    TRY(headtc::epack(eeg, pE, N, D, st));
    TRY(headtc::epack(bank, pB, M, D, st));
    headtc::LogitsArgs a{};
    a.Ap = pE; a.Bp = pB; a.a_blk0 = 0; a.M = N; a.N = M; a.D = D; a.tau = nullptr; a.out = logits; a.ldo = M;
    return headtc::logits_launch<3>(math, a, st);
  }
  GemmArgs g;
  g.A = eeg; g.B = bank; g.C = logits;
  g.M = N; g.N = M; g.K = D; g.KT = D;
  g.a_ms = D; g.a_ks = 1; g.b_ks = 1; g.b_ns = D; g.c_ms = M; g.c_ns = 1;
  return gemm_f32<0>(g, st);
}

}  // extern "C"
