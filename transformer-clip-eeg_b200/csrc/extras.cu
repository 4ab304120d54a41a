// Evaluation-side kernels of the EEG-CLIP path that are neither tower nor head:
//   * per-subject mean-variance normalisation of EEG windows (train_clip_helper_functions.py:133-136)
//   * per-row top-k of a similarity matrix (train_clip_helper_functions.py:182-187: torch.topk(logits, 100))
//   * the downstream regression head: Conv1d(latent -> n_out, k=32, 'same') + LeakyReLU and the Pearson loss
//     (train_clip_helper_functions.py:1107-1140, used by :620-640)
// All HBM/latency-bound integer / fp32 CUDA-core work: coalesced 128-bit accesses, fixed-order reductions (no float atomics).
#include "../../include/eegclip.h"
#include "common.cuh"

using namespace eegclip;

namespace {

// ------------------------------------------------------------------------------------------------
// Per-subject MVN: x (R, C) with R = N*T rows; per-channel mean / population std over all rows, y = (x - mean) / std.
// Two launches: (1) per-CTA partial sums of x and x^2 in fp64 (the reference reduces float64 numpy arrays);
// (2) every CTA folds the partials in fixed order and normalises its rows.
// ------------------------------------------------------------------------------------------------
constexpr int MVN_CTAS = 64;

__global__ void __launch_bounds__(256) mvn_stats_kernel(const float* __restrict__ x, double* __restrict__ partial, long R, int C) {
  pdl_sync();
  extern __shared__ double sh[];           // [256/C-groups][2][C]
  const int groups = 256 / C;              // C <= 256, 256 % C == 0
  const int c = threadIdx.x % C, g = threadIdx.x / C;
  double s = 0.0, q = 0.0;
  if (g < groups)
    for (long r = (long)blockIdx.x * groups + g; r < R; r += (long)gridDim.x * groups) {
      const double v = (double)x[r * C + c];
      s += v; q += v * v;
    }
  sh[(g * 2 + 0) * C + c] = s;
  sh[(g * 2 + 1) * C + c] = q;
  __syncthreads();
  if (threadIdx.x < C) {
    double ss = 0.0, qq = 0.0;
    for (int k = 0; k < groups; ++k) { ss += sh[(k * 2 + 0) * C + c]; qq += sh[(k * 2 + 1) * C + c]; }
    partial[((long)blockIdx.x * 2 + 0) * C + c] = ss;
    partial[((long)blockIdx.x * 2 + 1) * C + c] = qq;
  }
}

__global__ void __launch_bounds__(256) mvn_apply_kernel(const float* __restrict__ x, const double* __restrict__ partial,
                                                       float* __restrict__ y, long R, int C, int n_partial) {
  pdl_sync();
  extern __shared__ float shf[];           // mean[C], inv_std[C]
  if (threadIdx.x < C) {
    double s = 0.0, q = 0.0;
    for (int k = 0; k < n_partial; ++k) { s += partial[((long)k * 2 + 0) * C + threadIdx.x]; q += partial[((long)k * 2 + 1) * C + threadIdx.x]; }
    const double mean = s / (double)R;
    double var = q / (double)R - mean * mean;
    if (var < 0.0) var = 0.0;
    shf[threadIdx.x] = (float)mean;
    shf[C + threadIdx.x] = (float)(1.0 / sqrt(var));
  }
  __syncthreads();
  const long n = R * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    y[i] = (x[i] - shf[c]) * shf[C + c];
  }
}

// ------------------------------------------------------------------------------------------------
// Per-row top-k (k <= 1024) of x (N rows, M columns, row stride ld), one CTA per row, the row read twice (the second read
// hits L2).  Fast path (k <= 256): every thread takes the maximum of its strided share of the row; the k-th largest of
// the 256 thread maxima is a lower bound of the row's k-th largest value, and the elements at or above it (~1.3 k of them for
// k = 100, M = 1e5) are collected and sorted in shared memory -- no histogram, no hot shared-memory atomics.  Fallback
// (larger k, short rows, massive ties): radix select on order-preserving keys, 11 + 11 + 10 bits.  Output: values
// descending, ties broken by the lower column, so the result is deterministic.
// ------------------------------------------------------------------------------------------------
constexpr int TK_CAP = 2048;
constexpr int TK_THREADS = 256;

__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// f(key, column) over the row; 128-bit loads when the row is 16-byte aligned
template <class F>
__device__ __forceinline__ void scan_row(const float* __restrict__ row, int M, bool vec, F f) {
  if (vec) {
    const int m4 = M >> 2;
    for (int i = threadIdx.x; i < m4; i += TK_THREADS) {
      const float4 v = reinterpret_cast<const float4*>(row)[i];
      f(f2key(v.x), 4 * i); f(f2key(v.y), 4 * i + 1); f(f2key(v.z), 4 * i + 2); f(f2key(v.w), 4 * i + 3);
    }
    for (int i = (m4 << 2) + threadIdx.x; i < M; i += TK_THREADS) f(f2key(row[i]), i);
  } else {
    for (int i = threadIdx.x; i < M; i += TK_THREADS) f(f2key(row[i]), i);
  }
}

// bitonic sort of n2 (power of two) entries: descending by key, ascending by column on ties
__device__ __forceinline__ void bitonic_desc(uint32_t* ckey, int* cidx, int n2) {
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (n2 >> 1); t += TK_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint32_t ka = ckey[lo], kb = ckey[hi];
        const int ia = cidx[lo], ib = cidx[hi];
        const bool a_first = (ka > kb) || (ka == kb && ia < ib);    // a belongs before b in the final order
        if (a_first != desc) { ckey[lo] = kb; ckey[hi] = ka; cidx[lo] = ib; cidx[hi] = ia; }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(TK_THREADS) row_topk_kernel(const float* __restrict__ x, long ld, int M, int k,
                                                              float* __restrict__ vals, int64_t* __restrict__ idx, long col0) {
  __shared__ uint32_t hist[2048];
  __shared__ uint32_t tsum[TK_THREADS];
  __shared__ uint32_t ckey[TK_CAP];
  __shared__ int cidx[TK_CAP];
  __shared__ uint32_t s_bin, s_above, s_ncand;
  const float* row = x + (long)blockIdx.x * ld;
  const int tid = threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  bool done = false;
  if (k <= TK_THREADS && M >= 4 * TK_THREADS) {
    uint32_t mx = 0;
    scan_row(row, M, vec, [&](uint32_t key, int) { mx = max(mx, key); });
    ckey[tid] = mx; cidx[tid] = tid;
    if (tid == 0) s_ncand = 0;
    __syncthreads();
    bitonic_desc(ckey, cidx, TK_THREADS);
    const uint32_t t0 = ckey[k - 1];
    __syncthreads();
    scan_row(row, M, vec, [&](uint32_t key, int i) {
      if (key >= t0) {
        const uint32_t pos = atomicAdd(&s_ncand, 1u);
        if (pos < (uint32_t)TK_CAP) { ckey[pos] = key; cidx[pos] = i; }
      }
    });
    __syncthreads();
    done = s_ncand <= (uint32_t)TK_CAP;
    __syncthreads();
  }
  if (!done) {
    uint32_t prefix = 0;          // key bits fixed so far (value of the selected bins), the top `pbits` bits
    int pbits = 0;
    uint32_t need = (uint32_t)k;  // how many of the elements matching `prefix` are still wanted
    bool exact = false;
    for (int level = 0; level < 3; ++level) {
      const int bits = level < 2 ? 11 : 10;
      const int nb = 1 << bits;
      const int shift = 32 - pbits - bits;
      for (int i = tid; i < nb; i += TK_THREADS) hist[i] = 0;
      __syncthreads();
      scan_row(row, M, vec, [&](uint32_t key, int) {
        if (pbits == 0 || (key >> (32 - pbits)) == prefix) atomicAdd(&hist[(key >> shift) & (nb - 1)], 1u);
      });
      __syncthreads();
      const int per = nb / TK_THREADS;   // 8 or 4 bins per thread
      uint32_t mine = 0;
      for (int j = 0; j < per; ++j) mine += hist[tid * per + j];
      tsum[tid] = mine;
      __syncthreads();
      uint32_t above = 0;
      for (int t = tid + 1; t < TK_THREADS; ++t) above += tsum[t];
      uint32_t run = above;
      for (int j = per - 1; j >= 0; --j) {
        const uint32_t h = hist[tid * per + j];
        if (run < need && run + h >= need) { s_bin = (uint32_t)(tid * per + j); s_above = run; }
        run += h;
      }
      __syncthreads();
      const uint32_t bin = s_bin, ab = s_above;
      const uint32_t in_bin = hist[bin];
      __syncthreads();
      // (k - need) elements lie above the prefix, `ab` above the chosen bin inside it: all of those are selected for sure
      const uint32_t total_ge = ((uint32_t)k - need) + ab + in_bin;
      prefix = (prefix << bits) | bin;
      pbits += bits;
      need -= ab;
      if (total_ge <= (uint32_t)TK_CAP) break;
      if (level == 2) exact = true;      // prefix is a full key and too many elements equal it: take only `need` of them
    }
    if (tid == 0) s_ncand = 0;
    __syncthreads();
    scan_row(row, M, vec, [&](uint32_t key, int i) {
      const uint32_t top = pbits == 32 ? key : (key >> (32 - pbits));
      if (top > prefix || (!exact && top == prefix)) {
        const uint32_t pos = atomicAdd(&s_ncand, 1u);
        if (pos < (uint32_t)TK_CAP) { ckey[pos] = key; cidx[pos] = i; }
      }
    });
    __syncthreads();
    if (exact) {
      // massive tie at the k-th value: take the `need` LOWEST columns that equal it (ordered pass over the row, block-wide ranks)
      uint32_t taken = 0;
      const int lane = tid & 31, wid = tid >> 5;
      for (int base = 0; base < M && taken < need; base += TK_THREADS) {
        const int i = base + tid;
        const bool eq = i < M && f2key(row[i]) == prefix;
        const uint32_t bal = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) tsum[wid] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < TK_THREADS / 32; ++w) { if (w < wid) before += tsum[w]; total += tsum[w]; }
        const uint32_t rank = taken + before + __popc(bal & ((1u << lane) - 1u));
        if (eq && rank < need) {
          const uint32_t pos = atomicAdd(&s_ncand, 1u);
          if (pos < (uint32_t)TK_CAP) { ckey[pos] = prefix; cidx[pos] = i; }
        }
        taken += total;
        __syncthreads();
      }
    }
    __syncthreads();
  }
  const int n = (int)min(s_ncand, (uint32_t)TK_CAP);
  int n2 = 2;
  while (n2 < n) n2 <<= 1;
  for (int i = n + tid; i < n2; i += TK_THREADS) { ckey[i] = 0u; cidx[i] = 0x7fffffff; }
  __syncthreads();
  bitonic_desc(ckey, cidx, n2);
  for (int i = tid; i < k; i += TK_THREADS) {
    vals[(long)blockIdx.x * k + i] = i < n ? key2f(ckey[i]) : -INFINITY;
    idx[(long)blockIdx.x * k + i] = i < n ? (int64_t)cidx[i] + col0 : (int64_t)0;
  }
}

// ------------------------------------------------------------------------------------------------
// RegressionModel (train_clip_helper_functions.py:1132-1140): out = LeakyReLU(Conv1d(Cin -> Cout, K, 'same')(x)),
// channel-major x (B, Cin, T) -> out (B, Cout, T); 'same' pads (K-1)/2 on the left, the rest on the right (torch).
// grid (B, Cout); x window and the filter of this output channel in shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ out, int Cin,
                                                            int Cout, int T, int K) {
  pdl_sync();
  extern __shared__ float sm[];
  const int b = blockIdx.x, co = blockIdx.y, PL = (K - 1) / 2, TP = T + K - 1;
  float* xs = sm;                 // [Cin][TP]
  float* ws = sm + Cin * TP;      // [Cin][K]
  for (int i = threadIdx.x; i < Cin * TP; i += blockDim.x) {
    const int ci = i / TP, t = i % TP - PL;
    xs[i] = (t >= 0 && t < T) ? x[((long)b * Cin + ci) * T + t] : 0.f;
  }
  for (int i = threadIdx.x; i < Cin * K; i += blockDim.x) ws[i] = w[(long)co * Cin * K + i];
  __syncthreads();
  const float bb = bias ? bias[co] : 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float a = bb;
    for (int ci = 0; ci < Cin; ++ci)
      for (int kk = 0; kk < K; ++kk) a = fmaf(ws[ci * K + kk], xs[ci * TP + t + kk], a);
    out[((long)b * Cout + co) * T + t] = a > 0.f ? a : 0.01f * a;
  }
}

// per-(b, co) partial weight / bias gradients: part[b][co][ci*K + kk], pbias[b][co]; dpre = dout * LeakyReLU'(out)
__global__ void __launch_bounds__(256) conv_small_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ out,
                                                              const float* __restrict__ dout, float* __restrict__ part,
                                                              float* __restrict__ pbias, int Cin, int Cout, int T, int K) {
  pdl_sync();
  extern __shared__ float sm[];
  __shared__ float2 red[33];
  const int b = blockIdx.x, co = blockIdx.y, PL = (K - 1) / 2, TP = T + K - 1;
  float* xs = sm;                 // [Cin][TP]
  float* ds = sm + Cin * TP;      // [T]
  for (int i = threadIdx.x; i < Cin * TP; i += blockDim.x) {
    const int ci = i / TP, t = i % TP - PL;
    xs[i] = (t >= 0 && t < T) ? x[((long)b * Cin + ci) * T + t] : 0.f;
  }
  float sb = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const long o = ((long)b * Cout + co) * T + t;
    const float d = dout[o] * (out[o] > 0.f ? 1.f : 0.01f);
    ds[t] = d;
    sb += d;
  }
  const float2 r = block_sum2(sb, 0.f, red);
  if (threadIdx.x == 0) pbias[(long)b * Cout + co] = r.x;
  for (int i = threadIdx.x; i < Cin * K; i += blockDim.x) {
    const int ci = i / K, kk = i % K;
    float a = 0.f;
    for (int t = 0; t < T; ++t) a = fmaf(ds[t], xs[ci * TP + t + kk], a);
    part[((long)b * Cout + co) * Cin * K + i] = a;
  }
}

// dW[co][i] = sum_b part[b][co][i] (fixed order), db[co] = sum_b pbias[b][co]
__global__ void conv_small_reduce_kernel(const float* __restrict__ part, const float* __restrict__ pbias, float* __restrict__ dw,
                                         float* __restrict__ db, int B, int n_w, int Cout) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_w) {
    float a = 0.f;
    for (int b = 0; b < B; ++b) a += part[(long)b * n_w + i];
    dw[i] = a;
  }
  if (i < Cout && db) {
    float a = 0.f;
    for (int b = 0; b < B; ++b) a += pbias[(long)b * Cout + i];
    db[i] = a;
  }
}

// dx[b][ci][s] = sum_{co,kk} dpre[b][co][s - kk + PL] * w[co][ci][kk];  grid (B, Cin)
__global__ void __launch_bounds__(128) conv_small_dgrad_kernel(const float* __restrict__ out, const float* __restrict__ dout,
                                                              const float* __restrict__ w, float* __restrict__ dx, int Cin, int Cout,
                                                              int T, int K) {
  pdl_sync();
  extern __shared__ float sm[];
  const int b = blockIdx.x, ci = blockIdx.y, PL = (K - 1) / 2, TP = T + K - 1;
  float* ds = sm;                 // [Cout][TP], dpre shifted so that ds[co][s + (K-1-PL) ... ] lines up
  float* ws = sm + Cout * TP;     // [Cout][K]
  const int PR = K - 1 - PL;
  for (int i = threadIdx.x; i < Cout * TP; i += blockDim.x) {
    const int co = i / TP, t = i % TP - PR;
    float d = 0.f;
    if (t >= 0 && t < T) { const long o = ((long)b * Cout + co) * T + t; d = dout[o] * (out[o] > 0.f ? 1.f : 0.01f); }
    ds[i] = d;
  }
  for (int i = threadIdx.x; i < Cout * K; i += blockDim.x) ws[i] = w[((long)(i / K) * Cin + ci) * K + (i % K)];
  __syncthreads();
  for (int s = threadIdx.x; s < T; s += blockDim.x) {
    float a = 0.f;
    for (int co = 0; co < Cout; ++co)
      for (int kk = 0; kk < K; ++kk) a = fmaf(ws[co * K + kk], ds[co * TP + s + PR + PL - kk], a);   // t = s - kk + PL
    dx[((long)b * Cin + ci) * T + s] = a;
  }
}

// ------------------------------------------------------------------------------------------------
// PearsonLoss (train_clip_helper_functions.py:1107-1118): x, y (B, C, T);
//   r[b][c] = cos(x - mean_t x, y - mean_t y)  (CosineSimilarity(dim=2, eps=1e-6): each norm clamped at eps);
//   loss[c] = -mean_b r[b][c].   grid = C; warp w handles samples w, w+8, ...; thread 0 sums r over b in order.
// ------------------------------------------------------------------------------------------------
struct PStats { float mx, my, nx, ny, dot; };

__device__ __forceinline__ PStats pearson_stats(const float* __restrict__ xs, const float* __restrict__ ys, int T, int lane) {
  float sx = 0.f, sy = 0.f;
  for (int t = lane; t < T; t += 32) { sx += xs[t]; sy += ys[t]; }
  sx = warp_sum(sx); sy = warp_sum(sy);
  PStats p;
  p.mx = sx / (float)T; p.my = sy / (float)T;
  float xx = 0.f, yy = 0.f, xy = 0.f;
  for (int t = lane; t < T; t += 32) {
    const float a = xs[t] - p.mx, b = ys[t] - p.my;
    xx = fmaf(a, a, xx); yy = fmaf(b, b, yy); xy = fmaf(a, b, xy);
  }
  p.nx = fmaxf(sqrtf(warp_sum(xx)), 1e-6f);
  p.ny = fmaxf(sqrtf(warp_sum(yy)), 1e-6f);
  p.dot = warp_sum(xy);
  return p;
}

__global__ void __launch_bounds__(256) pearson_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         float* __restrict__ r_out, float* __restrict__ loss, int B, int C, int T) {
  pdl_sync();
  const int c = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = w; b < B; b += 8) {
    const PStats p = pearson_stats(x + ((long)b * C + c) * T, y + ((long)b * C + c) * T, T, lane);
    if (lane == 0) r_out[(long)b * C + c] = p.dot / (p.nx * p.ny);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += r_out[(long)b * C + c];
    loss[c] = -s / (float)B;
  }
}

// dx[b][c][t] = dloss[c] * (-1/B) * (yc_t / (nx ny) - r xc_t / nx^2)   (both terms are zero-mean: centering is its own adjoint)
__global__ void __launch_bounds__(256) pearson_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         const float* __restrict__ dloss, float* __restrict__ dx, int B, int C, int T) {
  pdl_sync();
  const int c = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float g = -dloss[c] / (float)B;
  for (int b = w; b < B; b += 8) {
    const float* xs = x + ((long)b * C + c) * T;
    const float* ys = y + ((long)b * C + c) * T;
    const PStats p = pearson_stats(xs, ys, T, lane);
    const float r = p.dot / (p.nx * p.ny);
    const float ia = g / (p.nx * p.ny), ib = g * r / (p.nx * p.nx);
    for (int t = lane; t < T; t += 32) dx[((long)b * C + c) * T + t] = ia * (ys[t] - p.my) - ib * (xs[t] - p.mx);
  }
}

}  // namespace

extern "C" {

int eegclip_mvn_workspace(int32_t C, size_t* scratch_bytes) {
  if (!scratch_bytes || C <= 0) return EEGCLIP_ERR_ARG;
  *scratch_bytes = (size_t)MVN_CTAS * 2 * C * sizeof(double);
  return EEGCLIP_OK;
}

int eegclip_mvn_normalize(const float* x, float* y, int64_t rows, int32_t C, void* scratch, void* stream) {
  if (!x || !y || !scratch || rows <= 0 || C <= 0) return EEGCLIP_ERR_ARG;
  if (C > 256 || (256 % C)) return EEGCLIP_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int groups = 256 / C;
  LAUNCH_PDL((mvn_stats_kernel), MVN_CTAS, 256, (size_t)groups * 2 * C * sizeof(double), st, x, (double*)scratch, (long)rows, (int)C);
  LAUNCH_CHECK();
  LAUNCH_PDL((mvn_apply_kernel), 148 * 4, 256, (size_t)2 * C * sizeof(float), st, x, (const double*)scratch, y, (long)rows, (int)C,
             MVN_CTAS);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

int eegclip_row_topk(const float* x, int64_t ld, int32_t N, int32_t M, int32_t k, int64_t col_offset, float* vals, int64_t* idx,
                     void* stream) {
  if (!x || !vals || !idx || N <= 0 || M <= 0 || k <= 0 || ld < M) return EEGCLIP_ERR_ARG;
  if (k > 1024 || k > M) return EEGCLIP_ERR_UNSUPPORTED;
  row_topk_kernel<<<N, TK_THREADS, 0, (cudaStream_t)stream>>>(x, (long)ld, M, k, vals, idx, (long)col_offset);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

int eegclip_conv_small_workspace(int32_t B, int32_t Cin, int32_t Cout, int32_t K, size_t* scratch_bytes) {
  if (!scratch_bytes || B <= 0 || Cin <= 0 || Cout <= 0 || K <= 0) return EEGCLIP_ERR_ARG;
  *scratch_bytes = ((size_t)B * Cout * Cin * K + (size_t)B * Cout) * sizeof(float);
  return EEGCLIP_OK;
}

static bool conv_small_ok(int Cin, int Cout, int T, int K) {
  const size_t smem = ((size_t)(Cin > Cout ? Cin : Cout) * (T + K - 1) + (size_t)(Cin > Cout ? Cin : Cout) * K + T) * sizeof(float);
  return smem <= 48 * 1024;
}

int eegclip_conv_small_forward(const float* x, const float* w, const float* bias, float* out, int32_t B, int32_t Cin, int32_t Cout,
                               int32_t T, int32_t K, void* stream) {
  if (!x || !w || !out || B <= 0 || Cin <= 0 || Cout <= 0 || T <= 0 || K <= 0) return EEGCLIP_ERR_ARG;
  if (!conv_small_ok(Cin, Cout, T, K)) return EEGCLIP_ERR_UNSUPPORTED;
  const size_t smem = ((size_t)Cin * (T + K - 1) + (size_t)Cin * K) * sizeof(float);
  LAUNCH_PDL((conv_small_fwd_kernel), dim3(B, Cout), 128, smem, (cudaStream_t)stream, x, w, bias, out, (int)Cin, (int)Cout, (int)T, (int)K);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

int eegclip_conv_small_backward(const float* x, const float* w, const float* out, const float* dout, float* dx, float* dw, float* db,
                                int32_t B, int32_t Cin, int32_t Cout, int32_t T, int32_t K, void* scratch, void* stream) {
  if (!x || !w || !out || !dout || !dw || !scratch || B <= 0 || Cin <= 0 || Cout <= 0 || T <= 0 || K <= 0) return EEGCLIP_ERR_ARG;
  if (!conv_small_ok(Cin, Cout, T, K)) return EEGCLIP_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  float* part = (float*)scratch;
  float* pbias = part + (size_t)B * Cout * Cin * K;
  const size_t smem_w = ((size_t)Cin * (T + K - 1) + T) * sizeof(float);
  LAUNCH_PDL((conv_small_wgrad_kernel), dim3(B, Cout), 256, smem_w, st, x, out, dout, part, pbias, (int)Cin, (int)Cout, (int)T, (int)K);
  LAUNCH_CHECK();
  const int n_w = Cout * Cin * K;
  LAUNCH_PDL((conv_small_reduce_kernel), (n_w + 255) / 256, 256, 0, st, (const float*)part, (const float*)pbias, dw, db, (int)B, n_w, (int)Cout);
  LAUNCH_CHECK();
  if (dx) {
    const size_t smem_d = ((size_t)Cout * (T + K - 1) + (size_t)Cout * K) * sizeof(float);
    LAUNCH_PDL((conv_small_dgrad_kernel), dim3(B, Cin), 128, smem_d, st, out, dout, w, dx, (int)Cin, (int)Cout, (int)T, (int)K);
    LAUNCH_CHECK();
  }
  return EEGCLIP_OK;
}

int eegclip_pearson_forward(const float* x, const float* y, float* r, float* loss, int32_t B, int32_t C, int32_t T, void* stream) {
  if (!x || !y || !r || !loss || B <= 0 || C <= 0 || T <= 0) return EEGCLIP_ERR_ARG;
  LAUNCH_PDL((pearson_fwd_kernel), C, 256, 0, (cudaStream_t)stream, x, y, r, loss, (int)B, (int)C, (int)T);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

int eegclip_pearson_backward(const float* x, const float* y, const float* dloss, float* dx, int32_t B, int32_t C, int32_t T,
                             void* stream) {
  if (!x || !y || !dloss || !dx || B <= 0 || C <= 0 || T <= 0) return EEGCLIP_ERR_ARG;
  LAUNCH_PDL((pearson_bwd_kernel), C, 256, 0, (cudaStream_t)stream, x, y, dloss, dx, (int)B, (int)C, (int)T);
  LAUNCH_CHECK();
  return EEGCLIP_OK;
}

}  // extern "C"
