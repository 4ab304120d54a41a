// EEG tower forward / backward: host-side orchestration of the per-layer kernels behind one C-ABI call.
// Replaces EEGConformerInterleaved.forward / EEGConformer.forward and their autograd
// (/root/reference/clip_model.py:445-474, 373-398, 75-94, 30-45, 60-67, 234-249).
//
// One call enqueues the whole tower on the caller's stream: no Python between kernels, graph-capturable.
// Layout: every activation is time-major (B,T,64) fp32, so the reference's permutes (clip_model.py:446,458,461)
// vanish; LayerNorm([C,T]) affines are read transposed from their (C,T) checkpoint layout.
#include "../../include/eegclip.h"
#include "common.cuh"
#include "gemm_f32.cuh"
#include "elementwise.cuh"
#include "attention.cuh"
#include "attention_tc.cuh"
#include "conv_tc.cuh"
#include "lin_tc.cuh"
#include "skinny.cuh"

using namespace eegclip;

namespace {

constexpr int C = 64;      // channels == embedding
constexpr int FF = 256;    // FFN hidden
constexpr int NP_CONV = 4, NP_XF = 16;

struct ConvP { const float *w, *b, *g, *be; };
struct XfP { const float *ln1g, *ln1b, *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo, *ln2g, *ln2b, *w1, *b1, *w2, *b2; };
struct ConvG { float *w, *b, *g, *be; };
struct XfG { float *ln1g, *ln1b, *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo, *ln2g, *ln2b, *w1, *b1, *w2, *b2; };

template <class P, class T>
P conv_at(T* tab, int i) { T* t = tab + 2 + NP_CONV * i; return P{t[0], t[1], t[2], t[3]}; }
template <class P, class T>
P xf_at(T* tab, int n_conv, int j) {
  T* t = tab + 2 + NP_CONV * n_conv + NP_XF * j;
  return P{t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], t[10], t[11], t[12], t[13], t[14], t[15]};
}

// ---- saved-activation layout (floats) -------------------------------------------------------
struct SaveLayout {
  size_t eegx;                 // n*C
  size_t conv_stride, xf_stride, conv0, xf0, total;
  // within a conv block: y (n*C), out (n*C), stats (2B rounded to 4)
  size_t c_y, c_out, c_stats;
  // within a transformer block: qkv (n*192), o (n*C), lse (B*8*T), z1 (n*C), fpre (n*FF), gp (n*FF), zout (n*C),
  //   h1 = LN1(zin), h2 = LN2(z1) (n*C each; tcgen05 path: kept so that the backward's weight gradients need no LayerNorm recompute)
  //   fp32 path   : fpre = FFN pre-activation (GELU and mask recomputed in the backward), gp unused
  //   tcgen05 path: fpre slot holds f = dropout(GELU(pre)), gp = mask * GELU'(pre): the backward multiplies, no erf / Philox
  size_t x_qkv, x_o, x_lse, x_z1, x_fpre, x_gp, x_zout, x_h1, x_h2, x_wp;
  size_t map_wp;               // packed eeg_spatial_mapping weights (forward + data-gradient forms)
};

// packed tensor-core operands of one transformer block (byte offsets inside its x_wp region)
struct XfPacked {
  static constexpr size_t QKV_F = 0;                       // (192 x 64)  forward
  static constexpr size_t QKV_D = QKV_F + 192 * 64 * 4;    // (64 x 192)  data gradient
  static constexpr size_t WO_F = QKV_D + 64 * 192 * 4;     // (64 x 64)
  static constexpr size_t WO_D = WO_F + 64 * 64 * 4;
  static constexpr size_t W1_F = WO_D + 64 * 64 * 4;       // (256 x 64)
  static constexpr size_t W1_D = W1_F + 256 * 64 * 4;      // (64 x 256)
  static constexpr size_t W2_F = W1_D + 64 * 256 * 4;      // (64 x 256)
  static constexpr size_t W2_D = W2_F + 64 * 256 * 4;      // (256 x 64)
  static constexpr size_t BQKV = W2_D + 256 * 64 * 4;      // 192 floats
  static constexpr size_t BYTES = BQKV + 1024;
};
constexpr size_t MAP_WP_BYTES = 2 * 64 * 64 * 4;

SaveLayout save_layout(const eegclip_tower_desc& d) {
  SaveLayout L;
  size_t n = (size_t)d.B * d.T;
  size_t o = 0;
  L.eegx = o; o += n * C;
  L.c_y = 0; L.c_out = n * C; L.c_stats = 2 * n * C;
  L.conv_stride = 2 * n * C + align_up((size_t)2 * d.B, 8);      // (every saved tensor starts 32-byte aligned: 256-bit accesses)
  L.x_qkv = 0; L.x_o = n * AQKV; L.x_lse = L.x_o + n * C; L.x_z1 = L.x_lse + align_up((size_t)d.B * AH * d.T, 8);
  L.x_fpre = L.x_z1 + n * C; L.x_gp = L.x_fpre + n * FF; L.x_zout = L.x_gp + n * FF;
  L.x_h1 = L.x_zout + n * C; L.x_h2 = L.x_h1 + n * C;
  L.x_wp = L.x_h2 + n * C;
  L.xf_stride = L.x_wp + XfPacked::BYTES / sizeof(float);
  L.map_wp = o; o += MAP_WP_BYTES / sizeof(float);
  L.conv0 = o; o += L.conv_stride * d.n_conv;
  L.xf0 = o; o += L.xf_stride * d.depth;
  L.total = o;
  return L;
}

struct Scratch {
  float *upad, *dypad, *h, *f, *dfpre, *dqkv, *d_o, *dg, *dza, *dzb, *dzc, *deeg, *wtmp, *wgp;   // wgp: 4 partial regions
  size_t wgp_stride;
  float* lnp;                       // 2 x [148][130]: per-CTA column sums of the fused LayerNorm-backward epilogues
  void* tc;
  size_t total;
};

Scratch scratch_layout(const eegclip_tower_desc& d, float* base) {
  Scratch s;
  size_t n = (size_t)d.B * d.T, TP = d.T + d.taps - 1;
  size_t o = 0;
  auto take = [&](size_t cnt) { float* p = base ? base + o : nullptr; o += align_up(cnt, 64); return p; };
  s.upad = take((size_t)d.B * TP * C);
  s.dypad = take((size_t)d.B * TP * C);
  s.h = take(n * C);
  s.f = take(n * FF);
  s.dfpre = take(n * FF);
  s.dqkv = take(n * AQKV);
  s.d_o = take(n * C);
  s.dg = take(n * C);
  s.dza = take(n * C);
  s.dzb = take(n * C);
  s.dzc = take(n * C);
  s.deeg = take(n * C);
  s.wtmp = take((size_t)C * C * d.taps);
  s.wgp_stride = align_up(lintc::lin_wgrad_partial_bytes(256, 64) / sizeof(float), 64);
  s.wgp = take(4 * s.wgp_stride);    // one region per weight gradient of a transformer block (their reductions are batched)
  s.lnp = take(2 * 148 * 130);
  s.tc = take(align_up(ln_ct_scratch_floats(d.T, C), 64) + conv_tc_scratch_bytes(d.B, d.T, d.taps, C, C) / sizeof(float) + 64);
  s.total = o;
  return s;
}

bool desc_ok(const eegclip_tower_desc* d) {
  if (!d) return false;
  if (d->B <= 0 || d->T <= 0 || (d->T & 3) || d->depth < 0 || d->n_conv < 0 || d->taps <= 0 || d->latent <= 0) return false;
  if (d->kind == EEGCLIP_TOWER_INTERLEAVED && d->n_conv != d->depth) return false;
  if (d->kind != EEGCLIP_TOWER_INTERLEAVED && d->kind != EEGCLIP_TOWER_SEQUENTIAL) return false;
  if (d->latent & 3) return false;
  return true;
}

#define TRY(x) do { int _r = (x); if (_r != EEGCLIP_OK) return _r; } while (0)

// out[m][n] = x[m][:] . w[n][:] + b[n]  with optional epilogue
int linear_f32(const float* x, long ldx, const float* w, const float* b, float* out, long ldo, long M, int N, int K,
               GemmEpi epi, cudaStream_t st) {
  GemmArgs g;
  g.A = x; g.B = w; g.C = out;
  g.M = (int)M; g.N = N; g.K = K; g.KT = K;
  g.a_ms = ldx; g.a_ks = 1;
  g.b_ks = 1; g.b_ns = K;
  g.c_ms = ldo; g.c_ns = 1;
  epi.bias_n = b;
  g.epi = epi;
  return gemm_f32(g, st);
}

// dx[m][k] = sum_n dy[m][n] * w[n][k]   (w is (N,K) row-major)
int linear_dgrad_f32(const float* dy, long lddy, const float* w, float* dx, long lddx, long M, int N, int K, GemmEpi epi,
                     cudaStream_t st) {
  GemmArgs g;
  g.A = dy; g.B = w; g.C = dx;
  g.M = (int)M; g.N = K; g.K = N; g.KT = N;
  g.a_ms = lddy; g.a_ks = 1;
  g.b_ks = K; g.b_ns = 1;
  g.c_ms = lddx; g.c_ns = 1;
  g.epi = epi;
  return gemm_f32(g, st);
}

// dw[n][k] += sum_m dy[m][n] * x[m][k]   (atomic split over m; dw must be zeroed by the caller)
int linear_wgrad_f32(const float* dy, long lddy, const float* x, long ldx, float* dw, long M, int N, int K, cudaStream_t st) {
  GemmArgs g;
  g.A = dy; g.B = x; g.C = dw;
  g.M = N; g.N = K; g.K = (int)M; g.KT = (int)M;
  g.a_ms = 1; g.a_ks = lddy;
  g.b_ks = ldx; g.b_ns = 1;
  g.c_ms = K; g.c_ns = 1;
  int tiles = ceil_div(N, GBM) * ceil_div(K, GBN);
  int sk = max(1, min(ceil_div((int)M, 512), (148 * 4) / max(1, tiles)));
  g.splitk = sk;
  g.epi.atomic = 1;
  return gemm_f32(g, st);
}

// conv forward: y[b][t][co] = bias[co] + sum_{k,ci} upad[b][t+k][ci] * W[co][ci][k], then dropout
int conv_fwd_f32(const float* upad, const float* W, const float* bias, float* y, int B, int T, int Cin, int Cout, int taps,
                 const Drop& drop, cudaStream_t st) {
  GemmArgs g;
  int TP = T + taps - 1;
  g.A = upad; g.B = W; g.C = y;
  g.batch = B; g.M = T; g.N = Cout; g.K = taps * Cin; g.KT = Cin;
  g.a_bs = (long)TP * Cin; g.a_ms = Cin; g.a_kbs = Cin; g.a_ks = 1;
  g.b_kbs = 1; g.b_ks = taps; g.b_ns = (long)Cin * taps;
  g.c_bs = (long)T * Cout; g.c_ms = Cout; g.c_ns = 1;
  g.epi.bias_n = bias;
  g.epi.drop_on = drop.enabled; g.epi.drop = drop;
  return gemm_f32(g, st);
}

// conv data gradient: du[b][t][ci] = sum_{k',co} dypad[b][t+k'][co] * W[co][ci][taps-1-k']
int conv_dgrad_f32(const float* dypad, const float* W, float* du, int B, int T, int Cin, int Cout, int taps, cudaStream_t st) {
  GemmArgs g;
  int TP = T + taps - 1;
  g.A = dypad; g.B = W + (taps - 1); g.C = du;
  g.batch = B; g.M = T; g.N = Cin; g.K = taps * Cout; g.KT = Cout;
  g.a_bs = (long)TP * Cout; g.a_ms = Cout; g.a_kbs = Cout; g.a_ks = 1;
  g.b_kbs = -1; g.b_ks = (long)Cin * taps; g.b_ns = taps;
  g.c_bs = (long)T * Cin; g.c_ms = Cin; g.c_ns = 1;
  return gemm_f32(g, st);
}

// conv weight gradient into tmp[co][k][ci] (zeroed by caller), reduction over (b,t)
int conv_wgrad_f32(const float* dypad, int PLb, const float* upad, float* tmp, int B, int T, int Cin, int Cout, int taps,
                   cudaStream_t st) {
  GemmArgs g;
  int TP = T + taps - 1;
  g.A = dypad + (long)PLb * Cout; g.B = upad; g.C = tmp;
  g.M = Cout; g.N = taps * Cin; g.K = B * T; g.KT = T;
  g.a_ms = 1; g.a_ks = Cout; g.a_kbs = (long)TP * Cout;
  g.b_ks = Cin; g.b_kbs = (long)TP * Cin; g.b_ns = 1;
  g.c_ms = (long)taps * Cin; g.c_ns = 1;
  int tiles = ceil_div(g.M, GBM) * ceil_div(g.N, GBN);
  g.splitk = max(1, min(ceil_div(g.K, 256), (148 * 4) / max(1, tiles)));
  g.epi.atomic = 1;
  return gemm_f32(g, st);
}

// -------------------------------------------------------------------------------------------------
// BasicBlock forward on time-major data.  xin (+ skip_in) -> conv -> dropout -> LN([C,T]) -> act (+ skip_out)
// -------------------------------------------------------------------------------------------------
// scratch of the padded narrow-output path (floats): padded weights | padded bias | padded conv output / output gradient |
// padded weight gradient, then the tensor-core conv scratch for Cout = 64
struct PadScr { float *w, *b, *y, *dw; void* tcs; };
size_t pad_scr_floats(int B, int T, int Cin, int taps) {
  const size_t TP = T + taps - 1;
  return 2 * align_up((size_t)64 * Cin * taps, 64) + 64 + align_up((size_t)B * TP * 64, 64) +
         conv_tc_scratch_bytes(B, T, taps, Cin, 64) / sizeof(float) + 64;
}
PadScr pad_scr(float* base, int B, int T, int Cin, int taps) {
  const size_t TP = T + taps - 1;
  PadScr s;
  s.w = base; s.b = s.w + align_up((size_t)64 * Cin * taps, 64); s.y = s.b + 64; s.dw = s.y + align_up((size_t)B * TP * 64, 64);
  s.tcs = s.dw + align_up((size_t)64 * Cin * taps, 64);
  return s;
}

// Stream breaks.  pdl_break() makes the next launch go out in plain stream order (no programmatic dependent launch on that edge):
// the kernel is not scheduled before its predecessor has drained.  On three edges per layer that is FASTER than the overlapped
// launch -- found when hoisting the per-layer dypad memset out of the loop made the step slower (the memset had been acting as such a
// break), then mapped edge by edge (tools/time_step_ab.py 15 <mask + 1> ..., one box, B = 256, T = 320, two streams):
//   bit 1  start of a conv block's backward (ct_transpose2 / LayerNorm statistics after the transformer block's reductions)  -0.28 ms
//   bit 4  before the attention backward (after the out-projection data gradient)                                            -0.30 ms
//   bit 8  start of a conv block's forward (weight pack / conv after the FFN2 GEMM)                                          -0.10 ms
// together 15.02 -> 14.41 ms per step.  Every other launch edge of the transformer / conv blocks (18 positions: block starts, before
// and after the attention kernels, before each token GEMM / weight gradient / reduction, before the conv data / weight gradient, before
// the LayerNorm forward, before the skip sums) measured neutral or slower (+0.02 ... +0.07 ms), as did dropping the attribute on every
// large <-> small shared-memory transition; programmatic launch off everywhere (g_tune[7]) is 0.2 ms slower on one stream.
// A hint at the mechanism: with NO breaks, issuing griddepcontrol.launch_dependents in lin_tc_kernel after its last MMA instead of
// first thing (g_tune[2] = 1) recovers 0.53 of those 0.6 ms -- the loss comes with the EARLY launch of a token GEMM's dependents; with
// the breaks in place the early trigger is the faster one again (14.22 vs 14.30 ms), so it stays.
// g_tune[15]: 0 = the shipped set, k > 0 = the bit mask k - 1.  g_tune[6] = 1: the break is a 4-byte memset instead (same timing).
constexpr int DEFAULT_BREAKS = 1 | 4 | 8;
inline int stream_breaks() { return g_tune[15] ? g_tune[15] - 1 : DEFAULT_BREAKS; }
#define STREAM_BREAK(bit, ptr) do { if (stream_breaks() & (bit)) { if (g_tune[6] && (void*)(ptr) != nullptr) CUDA_TRY(cudaMemsetAsync((void*)(ptr), 0, 4, st)); else pdl_break(); } } while (0)

int conv_block_fwd(int math, const float* xin, const float* skip_in, const ConvP& p, const float* skip_out, float* y, float* stats,
                   float* out, float* upad, void* tcs, int B, int T, int Cin, int Cout, int taps, int act, const Drop& drop,
                   cudaStream_t st, float* padscr = nullptr) {
  const int PL = (taps - 1) / 2;
  float* lnscr = (float*)tcs;                      // scratch layout: [LN transposed affine + partials][conv tensor-core scratch]
  tcs = lnscr + align_up(ln_ct_scratch_floats(T, Cout), 64);
  STREAM_BREAK(8, lnscr);
  if (math != EEGCLIP_MATH_FP32 && conv_tc_supported(Cin, Cout, taps, T)) {
    TRY(conv_tc_forward(math, xin, skip_in, p.w, p.b, y, B, T, Cin, Cout, taps, PL, drop, tcs, st));
  } else if (math != EEGCLIP_MATH_FP32 && padscr && conv_tc_padded_ok(Cin, Cout, taps, T)) {
    // narrow output (SpeechSmallConv): weights / bias zero-padded to 64 output channels, conv on the tensor-core kernel without
    // dropout, then compaction to Cout channels with the dropout mask indexed in the compact layout
    PadScr ps = pad_scr(padscr, B, T, Cin, taps);
    const size_t wrow = (size_t)Cin * taps;
    CUDA_TRY(cudaMemsetAsync(ps.w, 0, (64 * wrow + 64 + 64) * sizeof(float), st));      // (w, and b right behind it)
    CUDA_TRY(cudaMemcpyAsync(ps.w, p.w, (size_t)Cout * wrow * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (p.b) CUDA_TRY(cudaMemcpyAsync(ps.b, p.b, (size_t)Cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TRY(conv_tc_forward(math, xin, skip_in, ps.w, ps.b, ps.y, B, T, Cin, 64, taps, PL, make_drop(0, 0, 0, 0.f, 0), ps.tcs, st));
    TRY(compact_channels_drop(ps.y, y, (long)B * T, 64, Cout, drop, st));
  } else {
    TRY(pad_add(xin, skip_in, upad, B, T, Cin, PL, taps, st));
    TRY(conv_fwd_f32(upad, p.w, p.b, y, B, T, Cin, Cout, taps, drop, st));
  }
  TRY(ln_ct_act_fwd(y, p.g, p.be, skip_out, out, stats, lnscr, B, T, Cout, act, st));
  return EEGCLIP_OK;
}

// Backward of the block: dout is the gradient w.r.t. `out` (excluding skip_out's own path).
// Produces du (gradient w.r.t. xin + skip_in) and fills the four parameter gradients (pre-zeroed).
int conv_block_bwd(int math, const float* xin, const float* skip_in, const ConvP& p, const ConvG& gr, const float* y,
                   const float* stats, const float* dout, float* du, float* upad, float* dypad, float* wtmp, void* tcs, int B, int T,
                   int Cin, int Cout, int taps, int act, const Drop& drop, cudaStream_t st, float* padscr = nullptr,
                   bool zero_dypad = true) {
  const int PL = (taps - 1) / 2, PLb = taps - 1 - PL, TP = T + taps - 1;
  float* lnscr = (float*)tcs;
  tcs = lnscr + align_up(ln_ct_scratch_floats(T, Cout), 64);
  // only the pad rows need the zeros (the LayerNorm backward writes every valid row); a tower whose conv blocks all share one
  // geometry zeroes the buffer for its first block only (zero_dypad = false afterwards: nothing ever writes the pad rows)
  if (zero_dypad) CUDA_TRY(cudaMemsetAsync(dypad, 0, (size_t)B * TP * Cout * sizeof(float), st));
  else STREAM_BREAK(1, wtmp);
  bool bias_done = false;   // the conv bias gradient (column sums of dy) comes out of the LayerNorm backward's apply pass when it can
  TRY(ln_ct_act_bwd(dout, y, stats, p.g, p.be, dypad, gr.g, gr.be, wtmp /* 2B floats of per-sample means */, lnscr, B, T, Cout, PLb, taps, act,
                    drop, st, gr.b, &bias_done));
  // otherwise: column sums of dypad in fixed order (the LayerNorm scratch is free again: stream order)
  if (bias_done) {
  } else if ((size_t)colsum_det_ctas((long)B * TP) * Cout <= ln_ct_scratch_floats(T, Cout)) TRY(colsum_det(dypad, gr.b, (long)B * TP, Cout, Cout, lnscr, st));
  else TRY(colsum(dypad, gr.b, (long)B * TP, Cout, Cout, st));
  if (math != EEGCLIP_MATH_FP32 && conv_tc_supported(Cin, Cout, taps, T)) {
    TRY(conv_tc_backward(math, xin, skip_in, p.w, dypad, Cin, Cout, taps, PLb, du, gr.w, B, T, tcs, st));
  } else if (math != EEGCLIP_MATH_FP32 && padscr && conv_tc_padded_ok(Cin, Cout, taps, T)) {
    PadScr ps = pad_scr(padscr, B, T, Cin, taps);
    const size_t wrow = (size_t)Cin * taps;
    CUDA_TRY(cudaMemsetAsync(ps.w, 0, 64 * wrow * sizeof(float), st));
    CUDA_TRY(cudaMemcpyAsync(ps.w, p.w, (size_t)Cout * wrow * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TRY(pad_channels(dypad, ps.y, (long)B * TP, Cout, 64, st));                  // output gradient, channels zero-padded to 64
    TRY(conv_tc_backward(math, xin, skip_in, ps.w, ps.y, Cin, 64, taps, PLb, du, ps.dw, B, T, ps.tcs, st));
    CUDA_TRY(cudaMemcpyAsync(gr.w, ps.dw, (size_t)Cout * wrow * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    TRY(pad_add(xin, skip_in, upad, B, T, Cin, PL, taps, st));
    CUDA_TRY(cudaMemsetAsync(wtmp, 0, (size_t)Cout * Cin * taps * sizeof(float), st));
    TRY(conv_wgrad_f32(dypad, PLb, upad, wtmp, B, T, Cin, Cout, taps, st));
    TRY(wgrad_unpack(wtmp, gr.w, Cout, Cin, taps, st));
    TRY(conv_dgrad_f32(dypad, p.w, du, B, T, Cin, Cout, taps, st));
  }
  return EEGCLIP_OK;
}

struct XfSave { float *qkv, *o, *lse, *z1, *fpre, *gp, *zout, *h1, *h2; uint8_t* wp; };
bool xf_tc_ok(const eegclip_tower_desc& d);
int xf_block_fwd_tc(const eegclip_tower_desc& d, int layer, const XfP& p, const float* zin, const XfSave& s, Scratch& w, cudaStream_t st);
int xf_block_bwd_tc(const eegclip_tower_desc& d, int layer, const XfP& p, const XfG& g, const float* zin, const XfSave& s,
                    const float* dzout, float* dzin, Scratch& w, cudaStream_t st);

// TransformerEncoderBlock forward (clip_model.py:75-94).  zin -> zout
int xf_block_fwd(const eegclip_tower_desc& d, int layer, const XfP& p, const float* zin, const XfSave& s, Scratch& w,
                 cudaStream_t st) {
  if (xf_tc_ok(d)) return xf_block_fwd_tc(d, layer, p, zin, s, w, st);
  const long n = (long)d.B * d.T;
  TRY(ln64_fwd(zin, p.ln1g, p.ln1b, w.h, n, st));
  GemmEpi none;
  TRY(linear_f32(w.h, C, p.wq, p.bq, s.qkv + 0, AQKV, n, C, C, none, st));
  TRY(linear_f32(w.h, C, p.wk, p.bk, s.qkv + 64, AQKV, n, C, C, none, st));
  TRY(linear_f32(w.h, C, p.wv, p.bv, s.qkv + 128, AQKV, n, C, C, none, st));
  TRY(attention_fwd(s.qkv, s.o, s.lse, d.B, d.T, make_drop(d.seed, layer, SITE_ATTN, d.p_attn, d.train), st));
  GemmEpi e1;
  e1.drop = make_drop(d.seed, layer, SITE_PROJ, d.p_proj, d.train); e1.drop_on = e1.drop.enabled;
  e1.residual = zin;
  TRY(linear_f32(s.o, C, p.wo, p.bo, s.z1, C, n, C, C, e1, st));
  TRY(ln64_fwd(s.z1, p.ln2g, p.ln2b, w.h, n, st));
  GemmEpi e2;
  e2.act = 1; e2.aux = s.fpre;
  e2.drop = make_drop(d.seed, layer, SITE_FFN_HID, d.p_ffn_hid, d.train); e2.drop_on = e2.drop.enabled;
  TRY(linear_f32(w.h, C, p.w1, p.b1, w.f, FF, n, FF, C, e2, st));
  GemmEpi e3;
  e3.drop = make_drop(d.seed, layer, SITE_FFN_OUT, d.p_ffn_out, d.train); e3.drop_on = e3.drop.enabled;
  e3.residual = s.z1;
  TRY(linear_f32(w.f, FF, p.w2, p.b2, s.zout, C, n, C, FF, e3, st));
  return EEGCLIP_OK;
}

// Backward: dzout -> dzin (both n*C). Parameter gradients accumulate into pre-zeroed buffers.
int xf_block_bwd(const eegclip_tower_desc& d, int layer, const XfP& p, const XfG& g, const float* zin, const XfSave& s,
                 const float* dzout, float* dzin, Scratch& w, cudaStream_t st) {
  if (xf_tc_ok(d)) return xf_block_bwd_tc(d, layer, p, g, zin, s, dzout, dzin, w, st);
  const long n = (long)d.B * d.T;
  // ---- FFN branch -------------------------------------------------------------------------
  Drop d_out = make_drop(d.seed, layer, SITE_FFN_OUT, d.p_ffn_out, d.train);
  const float* dg = dzout;
  if (d_out.enabled) { TRY(drop_mul(dzout, w.dg, n * C, d_out, st)); dg = w.dg; }
  TRY(colsum(dg, g.b2, n, C, C, st));
  Drop d_hid = make_drop(d.seed, layer, SITE_FFN_HID, d.p_ffn_hid, d.train);
  TRY(gelu_drop(s.fpre, w.f, n * FF, d_hid, st));
  TRY(linear_wgrad_f32(dg, C, w.f, FF, g.w2, n, C, FF, st));
  GemmEpi e;
  e.drop = d_hid; e.drop_on = d_hid.enabled; e.act_grad_src = s.fpre;
  TRY(linear_dgrad_f32(dg, C, p.w2, w.dfpre, FF, n, C, FF, e, st));
  TRY(colsum(w.dfpre, g.b1, n, FF, FF, st));
  TRY(ln64_fwd(s.z1, p.ln2g, p.ln2b, w.h, n, st));
  TRY(linear_wgrad_f32(w.dfpre, FF, w.h, C, g.w1, n, FF, C, st));
  GemmEpi none;
  TRY(linear_dgrad_f32(w.dfpre, FF, p.w1, w.d_o, C, n, FF, C, none, st));      // d_o reused as dh2
  TRY(ln64_bwd(w.d_o, s.z1, p.ln2g, dzout, w.dzc, g.ln2g, g.ln2b, n, st));     // dzc = dz1
  // ---- attention branch -------------------------------------------------------------------
  Drop d_proj = make_drop(d.seed, layer, SITE_PROJ, d.p_proj, d.train);
  const float* dp = w.dzc;
  if (d_proj.enabled) { TRY(drop_mul(w.dzc, w.dg, n * C, d_proj, st)); dp = w.dg; }
  TRY(colsum(dp, g.bo, n, C, C, st));
  TRY(linear_wgrad_f32(dp, C, s.o, C, g.wo, n, C, C, st));
  TRY(linear_dgrad_f32(dp, C, p.wo, w.d_o, C, n, C, C, none, st));             // d_o = grad wrt attention output
  TRY(attention_bwd(s.qkv, s.o, w.d_o, s.lse, w.dqkv, d.B, d.T, make_drop(d.seed, layer, SITE_ATTN, d.p_attn, d.train), st));
  TRY(colsum(w.dqkv + 0, g.bq, n, C, AQKV, st));
  TRY(colsum(w.dqkv + 64, g.bk, n, C, AQKV, st));
  TRY(colsum(w.dqkv + 128, g.bv, n, C, AQKV, st));
  TRY(ln64_fwd(zin, p.ln1g, p.ln1b, w.h, n, st));
  TRY(linear_wgrad_f32(w.dqkv + 0, AQKV, w.h, C, g.wq, n, C, C, st));
  TRY(linear_wgrad_f32(w.dqkv + 64, AQKV, w.h, C, g.wk, n, C, C, st));
  TRY(linear_wgrad_f32(w.dqkv + 128, AQKV, w.h, C, g.wv, n, C, C, st));
  TRY(linear_dgrad_f32(w.dqkv + 0, AQKV, p.wq, w.d_o, C, n, C, C, none, st));  // d_o reused as dh1
  GemmEpi acc; acc.residual = w.d_o;
  TRY(linear_dgrad_f32(w.dqkv + 64, AQKV, p.wk, w.d_o, C, n, C, C, acc, st));
  TRY(linear_dgrad_f32(w.dqkv + 128, AQKV, p.wv, w.d_o, C, n, C, C, acc, st));
  TRY(ln64_bwd(w.d_o, zin, p.ln1g, w.dzc, dzin, g.ln1g, g.ln1b, n, st));
  return EEGCLIP_OK;
}


// ---- tensor-core (tcgen05) token-GEMM path of the transformer block -----------------------------------------
bool xf_tc_ok(const eegclip_tower_desc& d) { return d.math != EEGCLIP_MATH_FP32; }

// fp32 weights of one block -> packed bf16 hi/lo operands (forward and data-gradient forms) + concatenated QKV bias
int xf_pack(const XfP& p, uint8_t* wp, cudaStream_t st) {
  using namespace lintc;
  PackJobs J; J.n = 0;
  const float* wqkv[3] = {p.wq, p.wk, p.wv};
  const float* bqkv[3] = {p.bq, p.bk, p.bv};
  for (int i = 0; i < 3; ++i) {
    add_pack(J, wqkv[i], wp + XfPacked::QKV_F, 192, 64, 64 * i, 0, 64, 64, 64, 1);   // out n = 64*i + row, contraction k = column
    add_pack(J, wqkv[i], wp + XfPacked::QKV_D, 64, 192, 0, 64 * i, 64, 64, 1, 64);   // out n' = column, contraction k' = 64*i + row
    add_copy(J, bqkv[i], (float*)(wp + XfPacked::BQKV), 64 * i, 64);
  }
  add_pack(J, p.wo, wp + XfPacked::WO_F, 64, 64, 0, 0, 64, 64, 64, 1);
  add_pack(J, p.wo, wp + XfPacked::WO_D, 64, 64, 0, 0, 64, 64, 1, 64);
  add_pack(J, p.w1, wp + XfPacked::W1_F, 256, 64, 0, 0, 256, 64, 64, 1);
  add_pack(J, p.w1, wp + XfPacked::W1_D, 64, 256, 0, 0, 64, 256, 1, 64);
  add_pack(J, p.w2, wp + XfPacked::W2_F, 64, 256, 0, 0, 64, 256, 256, 1);
  add_pack(J, p.w2, wp + XfPacked::W2_D, 256, 64, 0, 0, 256, 64, 1, 256);
  return pack_launch(J, st);
}

lintc::LinTcArgs lin_args(const float* A, long lda, const uint8_t* w, float* Cp, long ldc, long M, int N, int K) {
  lintc::LinTcArgs a{};
  a.A = A; a.lda = lda; a.wpacked = w; a.C = Cp; a.ldc = ldc; a.M = (int)M; a.N = N; a.K = K;
  a.pro = lintc::PRO_NONE; a.pro_drop = make_drop(0, 0, 0, 0.f, 0); a.drop = a.pro_drop;
  return a;
}

int xf_block_fwd_tc(const eegclip_tower_desc& d, int layer, const XfP& p, const float* zin, const XfSave& s, Scratch& w,
                    cudaStream_t st) {
  using namespace lintc;
  const long n = (long)d.B * d.T;
  TRY(xf_pack(p, s.wp, st));
  TRY(ln64_fwd(zin, p.ln1g, p.ln1b, s.h1, n, st));
  {
    LinTcArgs a = lin_args(s.h1, C, s.wp + XfPacked::QKV_F, s.qkv, AQKV, n, AQKV, C);
    a.bias = (const float*)(s.wp + XfPacked::BQKV);
    TRY(lin_tc_launch(d.math, a, st));
  }
  if (attention_tc_supported(d.T)) TRY(attention_fwd_tc(s.qkv, s.o, s.lse, d.B, d.T, make_drop(d.seed, layer, SITE_ATTN, d.p_attn, d.train), st));
  else TRY(attention_fwd(s.qkv, s.o, s.lse, d.B, d.T, make_drop(d.seed, layer, SITE_ATTN, d.p_attn, d.train), st));
  {
    LinTcArgs a = lin_args(s.o, C, s.wp + XfPacked::WO_F, s.z1, C, n, C, C);
    a.bias = p.bo; a.drop = make_drop(d.seed, layer, SITE_PROJ, d.p_proj, d.train); a.drop_on = a.drop.enabled; a.residual = zin;
    TRY(lin_tc_launch(d.math, a, st));
  }
  TRY(ln64_fwd(s.z1, p.ln2g, p.ln2b, s.h2, n, st));
  {
    // f = mask * GELU(pre) -> saved (W2 operand now, dW2 operand later); gp = mask * GELU'(pre) -> saved for the data gradient
    LinTcArgs a = lin_args(s.h2, C, s.wp + XfPacked::W1_F, s.fpre, FF, n, FF, C);
    a.bias = p.b1; a.act = 2; a.aux = s.gp;
    a.drop = make_drop(d.seed, layer, SITE_FFN_HID, d.p_ffn_hid, d.train); a.drop_on = a.drop.enabled;
    TRY(lin_tc_launch(d.math, a, st));
  }
  {
    LinTcArgs a = lin_args(s.fpre, FF, s.wp + XfPacked::W2_F, s.zout, C, n, C, FF);
    a.bias = p.b2; a.drop = make_drop(d.seed, layer, SITE_FFN_OUT, d.p_ffn_out, d.train); a.drop_on = a.drop.enabled; a.residual = s.z1;
    TRY(lin_tc_launch(d.math, a, st));
  }
  return EEGCLIP_OK;
}

int xf_block_bwd_tc(const eegclip_tower_desc& d, int layer, const XfP& p, const XfG& g, const float* zin, const XfSave& s,
                    const float* dzout, float* dzin, Scratch& w, cudaStream_t st) {
  using namespace lintc;
  const long n = (long)d.B * d.T;
  const Drop d_out = make_drop(d.seed, layer, SITE_FFN_OUT, d.p_ffn_out, d.train);
  const Drop d_hid = make_drop(d.seed, layer, SITE_FFN_HID, d.p_ffn_hid, d.train);
  const Drop d_proj = make_drop(d.seed, layer, SITE_PROJ, d.p_proj, d.train);
  WgradReduceBatch red; red.n = 0;          // the four partial reductions of this block run as ONE launch at the end
  // ---- FFN branch: dg = dzout * mask_out (never materialised: prologue of both consumers) ----
  {
    LinWgradArgs a{};
    a.dy = dzout; a.lddy = C; a.Nout = C; a.x = s.fpre; a.ldx = FF; a.Kin = FF; a.M = (int)n;
    a.pro_dy = d_out.enabled ? PRO_DROP : PRO_NONE; a.drop_dy = d_out;
    a.pro_x = PRO_NONE; a.drop_x = d_hid;                            // x = f = dropout(GELU(pre)), saved by the forward
    a.partial = w.wgp;
    float* dW[3] = {g.w2, nullptr, nullptr}; float* db[3] = {g.b2, nullptr, nullptr};
    TRY(lin_wgrad_launch(d.math, a, dW, db, C, st, 0, nullptr, 1, &red));
  }
  {
    LinTcArgs a = lin_args(dzout, C, s.wp + XfPacked::W2_D, w.dfpre, FF, n, FF, C);
    a.pro = d_out.enabled ? PRO_DROP : PRO_NONE; a.pro_drop = d_out;
    a.mul_src = s.gp;                                                // * mask * GELU'(pre), saved by the forward
    TRY(lin_tc_launch(d.math, a, st));
  }
  {
    LinWgradArgs a{};
    a.dy = w.dfpre; a.lddy = FF; a.Nout = FF; a.x = s.h2; a.ldx = C; a.Kin = C; a.M = (int)n;
    a.drop_dy = d_out; a.drop_x = d_out; a.partial = w.wgp + w.wgp_stride;
    float* dW[3] = {g.w1, nullptr, nullptr}; float* db[3] = {g.b1, nullptr, nullptr};
    TRY(lin_wgrad_launch(d.math, a, dW, db, FF, st, 0, nullptr, 1, &red));
  }
  // dh2 = dfpre . W1 and the LayerNorm-2 backward in ONE launch: dzc = dz1 = dzout + LNbwd(dh2 ; z1); the affine gradients come out
  // as per-CTA column sums and join the batched reduction at the end of the block
  auto ln_reduce_job = [&](float* partial, float* dgamma, float* dbeta) {
    WgradReduceArgs r{};
    r.partial = partial; r.ctas = lin_tc_grid(n); r.Nout = 2; r.Kin = 64; r.rows_per_dst = 1; r.ldw = 64; r.log_scale = nullptr;
    r.dW[0] = dgamma; r.dW[1] = dbeta; r.dW[2] = nullptr; r.dW[3] = nullptr;
    for (int i = 0; i < 4; ++i) r.db[i] = nullptr;
    red.j[red.n] = r; red.kin_blocks[red.n] = 1; ++red.n;
  };
  {
    LinTcArgs a = lin_args(w.dfpre, FF, s.wp + XfPacked::W1_D, w.dzc, C, n, C, FF);
    a.lnb_x = s.z1; a.lnb_gamma = p.ln2g; a.lnb_partial = w.lnp; a.residual = dzout;
    TRY(lin_tc_launch(d.math, a, st));
    ln_reduce_job(w.lnp, g.ln2g, g.ln2b);
  }
  // ---- attention branch: dp = dz1 * mask_proj (prologue) ----
  {
    LinWgradArgs a{};
    a.dy = w.dzc; a.lddy = C; a.Nout = C; a.x = s.o; a.ldx = C; a.Kin = C; a.M = (int)n;
    a.pro_dy = d_proj.enabled ? PRO_DROP : PRO_NONE; a.drop_dy = d_proj; a.drop_x = d_proj; a.partial = w.wgp + 2 * w.wgp_stride;
    float* dW[3] = {g.wo, nullptr, nullptr}; float* db[3] = {g.bo, nullptr, nullptr};
    TRY(lin_wgrad_launch(d.math, a, dW, db, C, st, 0, nullptr, 1, &red));
  }
  {
    LinTcArgs a = lin_args(w.dzc, C, s.wp + XfPacked::WO_D, w.d_o, C, n, C, C);         // grad wrt attention output
    a.pro = d_proj.enabled ? PRO_DROP : PRO_NONE; a.pro_drop = d_proj;
    TRY(lin_tc_launch(d.math, a, st));
  }
  STREAM_BREAK(4, w.dqkv);
  if (attention_tc_supported(d.T)) TRY(attention_bwd_tc(s.qkv, s.o, w.d_o, s.lse, w.dqkv, d.B, d.T, make_drop(d.seed, layer, SITE_ATTN, d.p_attn, d.train), st));
  else TRY(attention_bwd(s.qkv, s.o, w.d_o, s.lse, w.dqkv, d.B, d.T, make_drop(d.seed, layer, SITE_ATTN, d.p_attn, d.train), st));
  {
    LinWgradArgs a{};
    a.dy = w.dqkv; a.lddy = AQKV; a.Nout = AQKV; a.x = s.h1; a.ldx = C; a.Kin = C; a.M = (int)n;
    a.drop_dy = d_out; a.drop_x = d_out; a.partial = w.wgp + 3 * w.wgp_stride;
    float* dW[3] = {g.wq, g.wk, g.wv}; float* db[3] = {g.bq, g.bk, g.bv};
    TRY(lin_wgrad_launch(d.math, a, dW, db, C, st, 0, nullptr, 1, &red));
  }
  {
    // dh1 = dq.Wq + dk.Wk + dv.Wv and the LayerNorm-1 backward: dzin = dz1 + LNbwd(dh1 ; zin)
    LinTcArgs a = lin_args(w.dqkv, AQKV, s.wp + XfPacked::QKV_D, dzin, C, n, C, AQKV);
    a.lnb_x = zin; a.lnb_gamma = p.ln1g; a.lnb_partial = w.lnp + 148 * 130; a.residual = w.dzc;
    TRY(lin_tc_launch(d.math, a, st));
    ln_reduce_job(w.lnp + 148 * 130, g.ln1g, g.ln1b);
  }
  TRY(lin_wgrad_reduce_flush(red, st));
  return EEGCLIP_OK;
}

XfSave xf_save(float* save, const SaveLayout& L, int j) {
  float* b = save + L.xf0 + L.xf_stride * j;
  return XfSave{b + L.x_qkv, b + L.x_o, b + L.x_lse, b + L.x_z1, b + L.x_fpre, b + L.x_gp, b + L.x_zout, b + L.x_h1, b + L.x_h2,
                (uint8_t*)(b + L.x_wp)};
}

}  // namespace

extern "C" {

int eegclip_tower_workspace(const eegclip_tower_desc* d, size_t* save_bytes, size_t* scratch_bytes) {
  if (!desc_ok(d)) return EEGCLIP_ERR_ARG;
  if (save_bytes) *save_bytes = save_layout(*d).total * sizeof(float);
  if (scratch_bytes) *scratch_bytes = scratch_layout(*d, nullptr).total * sizeof(float);
  return EEGCLIP_OK;
}

int eegclip_tower_forward(const eegclip_tower_desc* dp, const float* const* params, const float* x, float* out, void* save_v,
                          void* scratch_v, void* stream) {
  if (!desc_ok(dp) || !params || !x || !out || !scratch_v || !save_v) return EEGCLIP_ERR_ARG;
  NvtxRange nvtx_("eegclip_tower_forward");
  const eegclip_tower_desc& d = *dp;
  cudaStream_t st = (cudaStream_t)stream;
  SaveLayout L = save_layout(d);
  float* save = (float*)save_v;
  Scratch w = scratch_layout(d, (float*)scratch_v);
  const long n = (long)d.B * d.T;
  float* eegx = save + L.eegx;
  GemmEpi none;
  // eeg_spatial_mapping: 1x1 conv == per-token linear (clip_model.py:447)
  if (xf_tc_ok(d) && lintc::linear_tc_ok(n, C, C)) TRY(lintc::linear_tc_fwd(d.math, x, C, params[0], params[1], eegx, C, n, C, C, (uint8_t*)(save + L.map_wp), st));
  else TRY(linear_f32(x, C, params[0], params[1], eegx, C, n, C, C, none, st));
  const float* xin = eegx;
  if (d.kind == EEGCLIP_TOWER_INTERLEAVED) {
    for (int i = 0; i < d.depth; ++i) {
      NvtxRange nvtx_layer("tower layer forward (conv block + transformer block)");
      ConvP cp = conv_at<ConvP>(params, i);
      float* cb = save + L.conv0 + L.conv_stride * i;
      const bool last = (i == d.depth - 1);
      // conv input is x + eeg_x (clip_model.py:459); transformer input adds eeg_x again unless last (:465-469)
      TRY(conv_block_fwd(d.math, xin, eegx, cp, last ? nullptr : eegx, cb + L.c_y, cb + L.c_stats, cb + L.c_out, w.upad, w.tc, d.B, d.T, C,
                         C, d.taps, 0, make_drop(d.seed, i, SITE_CONV, d.p_conv, d.train), st));
      XfSave xs = xf_save(save, L, i);
      TRY(xf_block_fwd(d, i, xf_at<XfP>(params, d.n_conv, i), cb + L.c_out, xs, w, st));
      xin = xs.zout;
    }
  } else {
    for (int i = 0; i < d.n_conv; ++i) {
      ConvP cp = conv_at<ConvP>(params, i);
      float* cb = save + L.conv0 + L.conv_stride * i;
      const bool last = (i == d.n_conv - 1);  // no input skip in the last conv block (clip_model.py:385-390)
      TRY(conv_block_fwd(d.math, xin, last ? nullptr : eegx, cp, nullptr, cb + L.c_y, cb + L.c_stats, cb + L.c_out, w.upad, w.tc, d.B, d.T,
                         C, C, d.taps, 0, make_drop(d.seed, i, SITE_CONV, d.p_conv, d.train), st));
      xin = cb + L.c_out;
    }
    for (int j = 0; j < d.depth; ++j) {
      XfSave xs = xf_save(save, L, j);
      TRY(xf_block_fwd(d, d.n_conv + j, xf_at<XfP>(params, d.n_conv, j), xin, xs, w, st));
      xin = xs.zout;
    }
  }
  const int fi = 2 + NP_CONV * d.n_conv + NP_XF * d.depth;
  if (d.math != EEGCLIP_MATH_FP32 && skinny::supported(d.latent, C)) TRY(skinny::fwd(xin, params[fi], params[fi + 1], out, n, d.latent, st));
  else TRY(linear_f32(xin, C, params[fi], params[fi + 1], out, d.latent, n, d.latent, C, none, st));
  return EEGCLIP_OK;
}

int eegclip_tower_backward(const eegclip_tower_desc* dp, const float* const* params, float* const* grads, void* grad_base,
                           size_t grad_bytes, const float* x, const float* dout, float* dx, const void* save_v, void* scratch_v,
                           void* stream) {
  if (!desc_ok(dp) || !params || !grads || !x || !dout || !save_v || !scratch_v) return EEGCLIP_ERR_ARG;
  NvtxRange nvtx_("eegclip_tower_backward");
  const eegclip_tower_desc& d = *dp;
  cudaStream_t st = (cudaStream_t)stream;
  SaveLayout L = save_layout(d);
  float* save = (float*)save_v;  // read-only use
  Scratch w = scratch_layout(d, (float*)scratch_v);
  const long n = (long)d.B * d.T;
  const float* eegx = save + L.eegx;
  if (grad_base && grad_bytes) CUDA_TRY(cudaMemsetAsync(grad_base, 0, grad_bytes, st));
  GemmEpi none;
  const int fi = 2 + NP_CONV * d.n_conv + NP_XF * d.depth;

  // last activation feeding the final linear
  const float* xlast = eegx;
  if (d.depth > 0) xlast = xf_save(save, L, d.depth - 1).zout;
  else if (d.n_conv > 0) xlast = save + L.conv0 + L.conv_stride * (d.n_conv - 1) + L.c_out;

  float* dz = w.dza;
  float* dz2 = w.dzb;
  if (d.math != EEGCLIP_MATH_FP32 && skinny::supported(d.latent, C)) {
    TRY(skinny::wgrad(dout, xlast, grads[fi], grads[fi + 1], w.wgp, n, d.latent, st));
    TRY(skinny::dgrad(dout, params[fi], dz, n, d.latent, st));
  } else {
    TRY(colsum(dout, grads[fi + 1], n, d.latent, d.latent, st));
    TRY(linear_wgrad_f32(dout, d.latent, xlast, C, grads[fi], n, d.latent, C, st));
    TRY(linear_dgrad_f32(dout, d.latent, params[fi], dz, C, n, d.latent, C, none, st));
  }
  CUDA_TRY(cudaMemsetAsync(w.deeg, 0, (size_t)n * C * sizeof(float), st));

  if (d.kind == EEGCLIP_TOWER_INTERLEAVED) {
    for (int i = d.depth - 1; i >= 0; --i) {
      float* cb = save + L.conv0 + L.conv_stride * i;
      const bool last = (i == d.depth - 1);
      XfSave xs = xf_save(save, L, i);
      TRY(xf_block_bwd(d, i, xf_at<XfP>(params, d.n_conv, i), xf_at<XfG>(grads, d.n_conv, i), cb + L.c_out, xs, dz, dz2, w, st));
      const float* xin = (i == 0) ? eegx : xf_save(save, L, i - 1).zout;
      TRY(conv_block_bwd(d.math, xin, eegx, conv_at<ConvP>(params, i), conv_at<ConvG>(grads, i), cb + L.c_y, cb + L.c_stats, dz2, dz,
                         w.upad, w.dypad, w.wtmp, w.tc, d.B, d.T, C, C, d.taps, 0, make_drop(d.seed, i, SITE_CONV, d.p_conv, d.train), st,
                         nullptr, last));
      // skip gradients into eeg_x: the transformer input (dz2, not on the last layer, clip_model.py:463-466) and the conv input (dz);
      // dz2 is still intact here (conv_block_bwd only reads it), so both are added in one pass.  (fp32 sum order: deeg + (dz2 + dz))
      if (!last) TRY(add3_f32(w.deeg, dz2, dz, w.deeg, n * C, st));
      else TRY(add_f32(w.deeg, dz, w.deeg, n * C, st));
    }
    TRY(add_f32(w.deeg, dz, w.deeg, n * C, st));                // x_0 == eeg_x itself
  } else {
    for (int j = d.depth - 1; j >= 0; --j) {
      XfSave xs = xf_save(save, L, j);
      const float* zin = (j == 0) ? (d.n_conv > 0 ? save + L.conv0 + L.conv_stride * (d.n_conv - 1) + L.c_out : eegx)
                                  : xf_save(save, L, j - 1).zout;
      TRY(xf_block_bwd(d, d.n_conv + j, xf_at<XfP>(params, d.n_conv, j), xf_at<XfG>(grads, d.n_conv, j), zin, xs, dz, dz2, w, st));
      float* t = dz; dz = dz2; dz2 = t;
    }
    for (int i = d.n_conv - 1; i >= 0; --i) {
      float* cb = save + L.conv0 + L.conv_stride * i;
      const bool last = (i == d.n_conv - 1);
      const float* xin = (i == 0) ? eegx : save + L.conv0 + L.conv_stride * (i - 1) + L.c_out;
      TRY(conv_block_bwd(d.math, xin, last ? nullptr : eegx, conv_at<ConvP>(params, i), conv_at<ConvG>(grads, i), cb + L.c_y,
                         cb + L.c_stats, dz, dz2, w.upad, w.dypad, w.wtmp, w.tc, d.B, d.T, C, C, d.taps, 0,
                         make_drop(d.seed, i, SITE_CONV, d.p_conv, d.train), st, nullptr, last));
      if (!last) TRY(add_f32(w.deeg, dz2, w.deeg, n * C, st));
      float* t = dz; dz = dz2; dz2 = t;
    }
    TRY(add_f32(w.deeg, dz, w.deeg, n * C, st));                // x_0 == eeg_x
  }
  // eeg_spatial_mapping backward
  if (xf_tc_ok(d) && lintc::linear_tc_wgrad_ok(n, C, C) && lintc::linear_tc_dgrad_ok(n, C, C)) {
    TRY(lintc::linear_tc_wgrad(d.math, w.deeg, C, x, C, grads[0], grads[1], n, C, C, w.wgp, st));
    if (dx) TRY(lintc::linear_tc_dgrad(d.math, w.deeg, C, params[0], dx, C, n, C, C, (uint8_t*)w.wtmp, st));
  } else {
    TRY(colsum(w.deeg, grads[1], n, C, C, st));
    TRY(linear_wgrad_f32(w.deeg, C, x, C, grads[0], n, C, C, st));
    if (dx) TRY(linear_dgrad_f32(w.deeg, C, params[0], dx, C, n, C, C, none, st));
  }
  return EEGCLIP_OK;
}

// -------------------------------------------------------------------------------------------------
// Stand-alone conv block (BasicBlock of the speech tower, VLAAI layers) and linear layers
// -------------------------------------------------------------------------------------------------
int eegclip_convblock_workspace(const eegclip_convblock_desc* d, size_t* save_bytes, size_t* scratch_bytes) {
  if (!d || d->B <= 0 || d->T <= 0 || (d->Cin & 3) || (d->Cout & 3) || d->taps <= 0) return EEGCLIP_ERR_ARG;
  size_t n = (size_t)d->B * d->T, TP = d->T + d->taps - 1;
  if (save_bytes) *save_bytes = (n * d->Cout + align_up((size_t)2 * d->B, 4)) * sizeof(float);
  if (scratch_bytes)
    *scratch_bytes = (align_up((size_t)d->B * TP * d->Cin, 64) + align_up((size_t)d->B * TP * d->Cout, 64) +
                      align_up((size_t)d->Cout * d->Cin * d->taps, 64)) * sizeof(float) + conv_tc_scratch_bytes(d->B, d->T, d->taps, d->Cin, d->Cout) +
                      align_up(ln_ct_scratch_floats(d->T, d->Cout), 64) * sizeof(float) + 256 +
                      (conv_tc_padded_ok(d->Cin, d->Cout, d->taps, d->T) ? pad_scr_floats(d->B, d->T, d->Cin, d->taps) * sizeof(float) + 256 : 0);
  return EEGCLIP_OK;
}
// start of the padded-path scratch inside a convblock scratch buffer (nullptr when the shape does not use it)
static float* convblock_padscr(const eegclip_convblock_desc* d, void* scratch) {
  if (!conv_tc_padded_ok(d->Cin, d->Cout, d->taps, d->T)) return nullptr;
  const size_t TP = d->T + d->taps - 1;
  const size_t front = (align_up((size_t)d->B * TP * d->Cin, 64) + align_up((size_t)d->B * TP * d->Cout, 64) +
                        align_up((size_t)d->Cout * d->Cin * d->taps, 64)) * sizeof(float) +
                       conv_tc_scratch_bytes(d->B, d->T, d->taps, d->Cin, d->Cout) +
                       align_up(ln_ct_scratch_floats(d->T, d->Cout), 64) * sizeof(float) + 256;
  return (float*)((char*)scratch + align_up(front, 256));
}

int eegclip_convblock_forward(const eegclip_convblock_desc* d, const float* x, const float* skip_in, const float* w,
                              const float* bias, const float* gamma, const float* beta, float* out, void* save, void* scratch,
                              void* stream) {
  if (!d || !x || !w || !gamma || !beta || !out || !save || !scratch) return EEGCLIP_ERR_ARG;
  if ((d->Cin & 3) || (d->Cout & 3) || (d->T & 3) || d->Cout > 256) return EEGCLIP_ERR_UNSUPPORTED;
  size_t n = (size_t)d->B * d->T;
  float* y = (float*)save;
  float* stats = y + n * d->Cout;
  float* upad = (float*)scratch;
  size_t TPf = d->T + d->taps - 1;
  void* tcs = upad + align_up((size_t)d->B * TPf * d->Cin, 64) + align_up((size_t)d->B * TPf * d->Cout, 64) +
              align_up((size_t)d->Cout * d->Cin * d->taps, 64);
  ConvP p{w, bias, gamma, beta};
  return conv_block_fwd(d->math, x, skip_in, p, nullptr, y, stats, out, upad, tcs, d->B, d->T, d->Cin, d->Cout, d->taps, d->act,
                        make_drop(d->seed, d->layer, SITE_CONV, d->p_drop, d->train), (cudaStream_t)stream, convblock_padscr(d, scratch));
}

int eegclip_convblock_backward(const eegclip_convblock_desc* d, const float* x, const float* skip_in, const float* w,
                               const float* gamma, const float* beta, const float* dout, float* dx, float* dw, float* dbias,
                               float* dgamma, float* dbeta, const void* save, void* scratch, void* stream) {
  if (!d || !x || !w || !gamma || !beta || !dout || !dx || !dw || !dbias || !dgamma || !dbeta || !save || !scratch)
    return EEGCLIP_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  size_t n = (size_t)d->B * d->T, TP = d->T + d->taps - 1;
  const float* y = (const float*)save;
  const float* stats = y + n * d->Cout;
  float* upad = (float*)scratch;
  float* dypad = upad + align_up((size_t)d->B * TP * d->Cin, 64);
  float* wtmp = dypad + align_up((size_t)d->B * TP * d->Cout, 64);
  void* tcs = wtmp + align_up((size_t)d->Cout * d->Cin * d->taps, 64);
  CUDA_TRY(cudaMemsetAsync(dbias, 0, d->Cout * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(dgamma, 0, (size_t)d->Cout * d->T * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(dbeta, 0, (size_t)d->Cout * d->T * sizeof(float), st));
  ConvP p{w, nullptr, gamma, beta};
  ConvG g{dw, dbias, dgamma, dbeta};
  return conv_block_bwd(d->math, x, skip_in, p, g, y, stats, dout, dx, upad, dypad, wtmp, tcs, d->B, d->T, d->Cin, d->Cout, d->taps, d->act,
                        make_drop(d->seed, d->layer, SITE_CONV, d->p_drop, d->train), st, convblock_padscr(d, scratch));
}

int eegclip_linear_workspace(int64_t M, int32_t N, int32_t K, size_t* scratch_bytes) {
  if (M <= 0 || N <= 0 || K <= 0 || !scratch_bytes) return EEGCLIP_ERR_ARG;
  *scratch_bytes = lintc::linear_tc_ok(M, N, K) || lintc::linear_tc_wgrad_ok(M, N, K) ? lintc::linear_tc_scratch_bytes(N, K) : 256;
  return EEGCLIP_OK;
}

int eegclip_linear_forward(const float* x, const float* w, const float* b, float* out, int64_t M, int32_t N, int32_t K, int32_t math,
                           void* scratch, void* stream) {
  if (!x || !w || !out || M <= 0 || N <= 0 || K <= 0) return EEGCLIP_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (math != EEGCLIP_MATH_FP32 && scratch && lintc::linear_tc_ok(M, N, K))
    return lintc::linear_tc_fwd(math, x, K, w, b, out, N, M, N, K, (uint8_t*)scratch, st);
  GemmEpi none;
  return linear_f32(x, K, w, b, out, N, M, N, K, none, st);
}

int eegclip_linear_backward(const float* x, const float* w, const float* dout, float* dx, float* dw, float* db, int64_t M, int32_t N,
                            int32_t K, int32_t math, void* scratch, void* stream) {
  if (!x || !w || !dout || M <= 0 || N <= 0 || K <= 0) return EEGCLIP_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  GemmEpi none;
  const bool tc = math != EEGCLIP_MATH_FP32 && scratch;
  uint8_t* sc = (uint8_t*)scratch;
  if (dx) {
    if (tc && lintc::linear_tc_dgrad_ok(M, N, K)) {
      uint8_t* wpT = sc + lintc::packed_bytes(N, K) + 256 + lintc::lin_wgrad_partial_bytes(N, K < 256 ? K : 256, K < 256 ? 1 : K / 256) + 256;
      TRY(lintc::linear_tc_dgrad(math, dout, N, w, dx, K, M, N, K, wpT, st));
    } else {
      TRY(linear_dgrad_f32(dout, N, w, dx, K, M, N, K, none, st));
    }
  }
  bool db_done = false;
  if (dw) {
    if (tc && lintc::linear_tc_wgrad_ok(M, N, K)) {
      float* partial = (float*)(sc + lintc::packed_bytes(N, K) + 256);
      TRY(lintc::linear_tc_wgrad(math, dout, N, x, K, dw, db, M, N, K, partial, st));
      db_done = db != nullptr;
    } else {
      CUDA_TRY(cudaMemsetAsync(dw, 0, (size_t)N * K * sizeof(float), st));
      TRY(linear_wgrad_f32(dout, N, x, K, dw, M, N, K, st));
    }
  }
  if (db && !db_done) {
    CUDA_TRY(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st));
    if ((N & 3) == 0 && N <= 256 && 256 % (N >> 2) == 0) TRY(colsum(dout, db, M, N, N, st));
    else return EEGCLIP_ERR_UNSUPPORTED;
  }
  return EEGCLIP_OK;
}

// -------------------------------------------------------------------------------------------------
// Stand-alone TransformerEncoderBlock
// -------------------------------------------------------------------------------------------------
static eegclip_tower_desc xf_as_tower(const eegclip_xfblock_desc* d) {
  eegclip_tower_desc t{};
  t.kind = EEGCLIP_TOWER_SEQUENTIAL; t.B = d->B; t.T = d->T; t.n_conv = 0; t.depth = 1; t.taps = 1; t.latent = 8;
  t.train = d->train; t.math = d->math;
  t.p_attn = d->p_attn; t.p_proj = d->p_proj; t.p_ffn_hid = d->p_ffn_hid; t.p_ffn_out = d->p_ffn_out; t.seed = d->seed;
  return t;
}

int eegclip_xfblock_workspace(const eegclip_xfblock_desc* d, size_t* save_bytes, size_t* scratch_bytes) {
  if (!d || d->B <= 0 || d->T <= 0 || (d->T & 3)) return EEGCLIP_ERR_ARG;
  eegclip_tower_desc t = xf_as_tower(d);
  SaveLayout L = save_layout(t);
  if (save_bytes) *save_bytes = L.xf_stride * sizeof(float);
  if (scratch_bytes) *scratch_bytes = scratch_layout(t, nullptr).total * sizeof(float);
  return EEGCLIP_OK;
}

int eegclip_xfblock_forward(const eegclip_xfblock_desc* d, const float* const* params, const float* zin, float* zout, void* save,
                            void* scratch, void* stream) {
  if (!d || !params || !zin || !zout || !save || !scratch || d->B <= 0 || d->T <= 0 || (d->T & 3)) return EEGCLIP_ERR_ARG;
  eegclip_tower_desc t = xf_as_tower(d);
  SaveLayout L = save_layout(t);
  float* b = (float*)save;
  XfSave xs{b + L.x_qkv, b + L.x_o, b + L.x_lse, b + L.x_z1, b + L.x_fpre, b + L.x_gp, zout, b + L.x_h1, b + L.x_h2, (uint8_t*)(b + L.x_wp)};
  Scratch w = scratch_layout(t, (float*)scratch);
  const float* const* tp = params - 2;  // xf_at() skips the two mapping entries
  return xf_block_fwd(t, d->layer, xf_at<XfP>(tp, 0, 0), zin, xs, w, (cudaStream_t)stream);
}

int eegclip_xfblock_backward(const eegclip_xfblock_desc* d, const float* const* params, float* const* grads, void* grad_base,
                             size_t grad_bytes, const float* zin, const float* dzout, float* dzin, const void* save, void* scratch,
                             void* stream) {
  if (!d || !params || !grads || !zin || !dzout || !dzin || !save || !scratch) return EEGCLIP_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  eegclip_tower_desc t = xf_as_tower(d);
  SaveLayout L = save_layout(t);
  float* b = (float*)save;
  XfSave xs{b + L.x_qkv, b + L.x_o, b + L.x_lse, b + L.x_z1, b + L.x_fpre, b + L.x_gp, nullptr, b + L.x_h1, b + L.x_h2, (uint8_t*)(b + L.x_wp)};
  Scratch w = scratch_layout(t, (float*)scratch);
  if (grad_base && grad_bytes) CUDA_TRY(cudaMemsetAsync(grad_base, 0, grad_bytes, st));
  return xf_block_bwd(t, d->layer, xf_at<XfP>(params - 2, 0, 0), xf_at<XfG>(grads - 2, 0, 0), zin, xs, dzout, dzin, w, st);
}

// -------------------------------------------------------------------------------------------------
// Stand-alone pieces of the transformer block (MultiHeadAttention.forward, ResidualAdd.forward; clip_model.py:30-57)
// -------------------------------------------------------------------------------------------------
int eegclip_attention_forward(const float* qkv, float* out, float* lse, int32_t B, int32_t T, float p_drop, int32_t train,
                              int32_t layer, uint64_t seed, int32_t math, void* stream) {
  if (!qkv || !out || !lse || B <= 0 || T <= 0 || (T & 3)) return EEGCLIP_ERR_ARG;
  const Drop drop = make_drop(seed, layer, SITE_ATTN, p_drop, train);
  if (math != EEGCLIP_MATH_FP32 && attention_tc_supported(T)) return attention_fwd_tc(qkv, out, lse, B, T, drop, (cudaStream_t)stream);
  return attention_fwd(qkv, out, lse, B, T, drop, (cudaStream_t)stream);
}

int eegclip_attention_backward(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv, int32_t B,
                               int32_t T, float p_drop, int32_t train, int32_t layer, uint64_t seed, int32_t math, void* stream) {
  if (!qkv || !out || !dout || !lse || !dqkv || B <= 0 || T <= 0 || (T & 3)) return EEGCLIP_ERR_ARG;
  const Drop drop = make_drop(seed, layer, SITE_ATTN, p_drop, train);
  if (math != EEGCLIP_MATH_FP32 && attention_tc_supported(T)) return attention_bwd_tc(qkv, out, dout, lse, dqkv, B, T, drop, (cudaStream_t)stream);
  return attention_bwd(qkv, out, dout, lse, dqkv, B, T, drop, (cudaStream_t)stream);
}

int eegclip_layernorm_forward(const float* x, const float* gamma, const float* beta, float* out, int64_t rows, int32_t C, void* stream) {
  if (!x || !gamma || !beta || !out || rows <= 0) return EEGCLIP_ERR_ARG;
  if (C != 64) return EEGCLIP_ERR_UNSUPPORTED;
  return ln64_fwd(x, gamma, beta, out, rows, (cudaStream_t)stream);
}

int eegclip_layernorm_backward(const float* dout, const float* x, const float* gamma, float* dx, float* dgamma, float* dbeta,
                               int64_t rows, int32_t C, void* stream) {
  if (!dout || !x || !gamma || !dx || !dgamma || !dbeta || rows <= 0) return EEGCLIP_ERR_ARG;
  if (C != 64) return EEGCLIP_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(dgamma, 0, 64 * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(dbeta, 0, 64 * sizeof(float), st));
  return ln64_bwd(dout, x, gamma, nullptr, dx, dgamma, dbeta, rows, st);
}

int eegclip_dropout(const float* in, float* out, int64_t n, float p, int32_t train, int32_t layer, int32_t site, uint64_t seed,
                    void* stream) {
  if (!in || !out || n <= 0 || (n & 3) || site < 0 || site > 15) return EEGCLIP_ERR_ARG;
  return drop_mul(in, out, n, make_drop(seed, layer, site, p, train), (cudaStream_t)stream);
}

// FeedForwardBlock's nn.GELU() -> nn.Dropout pair (clip_model.py:64-65) as one elementwise pass and its backward
int eegclip_gelu_dropout_forward(const float* pre, float* out, int64_t n, float p, int32_t train, int32_t layer, int32_t site,
                                 uint64_t seed, void* stream) {
  if (!pre || !out || n <= 0 || (n & 3) || site < 0 || site > 15) return EEGCLIP_ERR_ARG;
  return gelu_drop(pre, out, n, make_drop(seed, layer, site, p, train), (cudaStream_t)stream);
}

int eegclip_gelu_dropout_backward(const float* pre, const float* dout, float* dpre, int64_t n, float p, int32_t train, int32_t layer,
                                  int32_t site, uint64_t seed, void* stream) {
  if (!pre || !dout || !dpre || n <= 0 || (n & 3) || site < 0 || site > 15) return EEGCLIP_ERR_ARG;
  return gelu_drop_bwd(pre, dout, dpre, n, make_drop(seed, layer, site, p, train), (cudaStream_t)stream);
}

}  // extern "C"
