// Shared device/host helpers for the EEG-CLIP B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges are no-ops unless a profiler injects itself
#include <stdint.h>
#include <math.h>

#define EEGCLIP_OK 0
#define EEGCLIP_ERR_ARG -1
#define EEGCLIP_ERR_CUDA -2
#define EEGCLIP_ERR_UNSUPPORTED -3

#define CUDA_TRY(expr)                                   \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return EEGCLIP_ERR_CUDA;      \
  } while (0)

#define LAUNCH_CHECK()                                   \
  do {                                                   \
    ++::eegclip::g_launch_count;                         \
    if (cudaPeekAtLastError() != cudaSuccess) return EEGCLIP_ERR_CUDA; \
  } while (0)

namespace eegclip {

// ---- launch accounting + optional per-kernel-class device timing (bench.py roofline; see eegclip_profile_*) ----
extern long long g_launch_count;
extern int g_tune[16];
extern thread_local int g_pdl_break;   // set by pdl_break(): the next LAUNCH_PDL of this thread omits the programmatic-launch attribute
extern unsigned long long* g_dbg_buf;   // development timeline buffer (eegclip_debug_buffer), nullptr in production   // development knobs (eegclip_tune_set): 0 lin ring depth, 1 streaming-load policy
enum : int { PROF_CONV_TC = 0, PROF_WGRAD_TC = 1, PROF_ATTN_FWD = 2, PROF_ATTN_BWD = 3, PROF_LNCT = 4, PROF_GEMM_F32 = 5, PROF_LIN_TC = 6, PROF_LIN_WGRAD = 7, PROF_LSTM = 8, PROF_NCLASS = 12 };
void prof_begin(int cls, cudaStream_t st);
void prof_end(int cls, cudaStream_t st);
struct ProfScope {
  int cls; cudaStream_t st;
  ProfScope(int c, cudaStream_t s) : cls(c), st(s) { prof_begin(c, s); }
  ~ProfScope() { prof_end(cls, st); }
};

// NVTX range around a host-side enqueue section (C-ABI entry points, per layer inside the towers): names the launch groups in
// Nsight Systems / ncu --nvtx timelines.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// Programmatic dependent launch.  A step is ~510 dependent kernel launches on one stream and the back-to-back launch gap
// (~2 us: drain, block scheduling, barrier / TMEM set-up of the next kernel) was ~5 % of the step.  Every hot kernel starts
// with pdl_sync(): `launch_dependents` lets the NEXT kernel's CTAs be scheduled as soon as all of this kernel's CTAs have
// started, `wait` then blocks until the PREVIOUS grid has completed and its memory is visible -- nothing of the kernel body
// runs before that, so the semantics are exactly stream order; only launch latency and CTA ramp-up overlap the predecessor's
// tail.  Launches go through LAUNCH_PDL (cudaLaunchKernelEx + programmatic stream serialization); g_tune[7] = 1 turns the
// attribute off (A/B timing).  A kernel launched without the attribute executes both instructions as no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_trigger(); pdl_wait(); }
// (the tensor-core kernels split the pair: barrier initialisation and the TMEM allocation, which touch no global memory, run
//  between pdl_trigger() and pdl_wait() and so overlap the predecessor's tail as well)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_tune[7] || g_pdl_break) ? 0 : 1;
  g_pdl_break = 0;

  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// pdl_break(): the NEXT launch of this thread goes out without the attribute, i.e. in plain stream order -- it is not scheduled
// before its predecessor has drained (see STREAM_BREAK in tower.cu for where that is faster and by how much)
inline void pdl_break() { g_pdl_break = 1; }
#define LAUNCH_PDL(kernel, grid, block, smem, st, ...) (void)::eegclip::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)

// dropout site ids; stream = layer * 16 + site  (oracle/philox_ref.py::stream_id)
enum : int { SITE_CONV = 0, SITE_ATTN = 1, SITE_PROJ = 2, SITE_FFN_HID = 3, SITE_FFN_OUT = 4 };

struct Drop {
  uint32_t seed_lo, seed_hi;
  uint32_t stream;
  uint32_t thresh;   // 16-bit mode: keep iff 16-bit draw >= thresh  (thresh = floor(p * 65536))
  float scale;       // 1/(1-p); p == 0 -> disabled (thresh 0, scale 1)
  int enabled;
  int onebit;        // p == 0.5 exactly: ONE bit per decision (128 decisions per Philox call), keep iff bit set
};

__host__ inline Drop make_drop(uint64_t seed, int layer, int site, float p, int train) {
  Drop d;
  d.seed_lo = (uint32_t)(seed & 0xffffffffull);
  d.seed_hi = (uint32_t)(seed >> 32);
  d.stream = (uint32_t)(layer * 16 + site);
  d.enabled = (train && p > 0.f) ? 1 : 0;
  d.onebit = (d.enabled && p == 0.5f) ? 1 : 0;
  double t = floor((double)p * 65536.0);
  if (t > 65535.0) t = 65535.0;
  d.thresh = d.enabled ? (uint32_t)t : 0u;
  d.scale = d.enabled ? 1.0f / (1.0f - p) : 1.0f;
  return d;
}

// not inlined on purpose: ~100 instructions per copy, and the fused kernels call it from many unrolled sites
static __device__ __noinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Dropout decisions (oracle/philox_ref.py::keep_mask).  Philox block `blk` = philox(ctr = (blk lo, blk hi, stream, 0), key = seed).
//   16-bit mode (p != 0.5): block = idx >> 3; element e = idx & 7 draws (word[e >> 1] >> (16 * (e & 1))) & 0xffff, keep iff >= thresh
//   1-bit  mode (p == 0.5): block = idx >> 7; element e = idx & 127 keeps iff bit (e & 31) of word[e >> 5] is set
__device__ __forceinline__ uint4 drop_words(const Drop& d, uint64_t block) {
  return philox4x32_10((uint32_t)(block & 0xffffffffull), (uint32_t)(block >> 32), d.stream, 0u, d.seed_lo, d.seed_hi);
}
__device__ __forceinline__ uint32_t word_of(const uint4& w, uint32_t i) { return i == 0 ? w.x : i == 1 ? w.y : i == 2 ? w.z : w.w; }

// keep bits (bit e <-> element idx + e) of the 8 consecutive elements starting at idx (idx % 8 == 0)
__device__ __forceinline__ uint32_t drop_bits8(const Drop& d, uint64_t idx) {
  if (!d.enabled) return 0xffu;
  if (d.onebit) {
    const uint4 w = drop_words(d, idx >> 7);
    const uint32_t e = (uint32_t)(idx & 127);
    return (word_of(w, e >> 5) >> (e & 31)) & 0xffu;
  }
  const uint4 w = drop_words(d, idx >> 3);
  const uint32_t t = d.thresh;
  uint32_t m = 0;
  m |= ((w.x & 0xffffu) >= t) ? 1u : 0u;   m |= ((w.x >> 16) >= t) ? 2u : 0u;
  m |= ((w.y & 0xffffu) >= t) ? 4u : 0u;   m |= ((w.y >> 16) >= t) ? 8u : 0u;
  m |= ((w.z & 0xffffu) >= t) ? 16u : 0u;  m |= ((w.z >> 16) >= t) ? 32u : 0u;
  m |= ((w.w & 0xffffu) >= t) ? 64u : 0u;  m |= ((w.w >> 16) >= t) ? 128u : 0u;
  return m;
}

// Multiplier (0 or 1/(1-p)) for a single element index.
__device__ __forceinline__ float drop_mult(const Drop& d, uint64_t idx) {
  if (!d.enabled) return 1.0f;
  const uint32_t bits = drop_bits8(d, idx & ~7ull);
  return (bits >> (uint32_t)(idx & 7)) & 1u ? d.scale : 0.0f;
}

// Multipliers for the 8 consecutive elements starting at idx (idx % 8 == 0).
__device__ __forceinline__ void drop_mult8(const Drop& d, uint64_t idx, float* m) {
  const uint32_t bits = drop_bits8(d, idx);
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = (bits >> e) & 1u ? d.scale : 0.f;
}

// Multipliers for the 4 consecutive elements starting at idx (idx % 4 == 0).
__device__ __forceinline__ float4 drop_mult4(const Drop& d, uint64_t idx) {
  if (!d.enabled) return make_float4(1.f, 1.f, 1.f, 1.f);
  const uint32_t bits = drop_bits8(d, idx & ~7ull) >> (uint32_t)(idx & 4);
  return make_float4(bits & 1u ? d.scale : 0.f, bits & 2u ? d.scale : 0.f, bits & 4u ? d.scale : 0.f, bits & 8u ? d.scale : 0.f);
}

// Exact (erf) GELU of nn.GELU() (clip_model.py:64,240).  erf via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, the level of erff's
// own fp32 rounding): one MUFU.RCP + one MUFU.EX2 + 5 FMA, and the exponential exp(-x^2/2) is shared with the Gaussian density
// that GELU' needs -- ~17 instructions for GELU and GELU' together against ~45 with erff + __expf (the FFN1 epilogue and the
// LayerNorm([C,T]) kernels were issue-bound on it).  Measured against fp64: |GELU error| <= 4.7e-7, |GELU' error| <= 3.2e-7.
__device__ __forceinline__ void gelu_core(float x, float& cdf, float& e) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, er;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-(z * z) * 1.4426950408889634f));   // exp(-x^2 / 2)
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  er = fmaf(-(p * t), e, 1.0f);                      // erf(|x| / sqrt 2)
  cdf = 0.5f + copysignf(0.5f * er, x);
}
__device__ __forceinline__ float gelu_f(float x) {
  float cdf, e;
  gelu_core(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float cdf, e;
  gelu_core(x, cdf, e);
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}
// GELU(x) and GELU'(x) together
__device__ __forceinline__ void gelu_both(float x, float& g, float& gd) {
  float cdf, e;
  gelu_core(x, cdf, e);
  g = x * cdf;
  gd = fmaf(x * 0.39894228040143267794f, e, cdf);
}
// activation ids for conv+LN blocks: 0 = GELU (BasicBlock), 1 = LeakyReLU(0.01) (VLAAI / SpeechSmallConv)
__device__ __forceinline__ float act_f(float x, int act) { return act == 0 ? gelu_f(x) : (x > 0.f ? x : 0.01f * x); }
__device__ __forceinline__ float act_grad_f(float x, int act) { return act == 0 ? gelu_grad_f(x) : (x > 0.f ? 1.f : 0.01f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of two values (blockDim.x multiple of 32, <= 1024). Result valid in all threads.
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* sh /* >= 33 entries */) {
  a = warp_sum(a); b = warp_sum(b);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = make_float2(a, b);
  __syncthreads();
  if (w == 0) {
    float2 v = l < nw ? sh[l] : make_float2(0.f, 0.f);
    v.x = warp_sum(v.x); v.y = warp_sum(v.y);
    if (l == 0) sh[32] = v;
  }
  __syncthreads();
  return sh[32];
}

// 256-bit global accesses (sm_100): a thread that owns 32 contiguous bytes moves them with one instruction, so a warp instruction
// covers whole 32-byte sectors (two 128-bit accesses per thread touch every sector twice, half of it each time)
__device__ __forceinline__ void ld256(const float* p, float* v) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void st256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
               "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace eegclip
