// Bidirectional LSTM entry points (speech tower, clip_model.py:267-268, 322-323): input projections / weight gradients /
// input gradient on the token-GEMM kernels (lin_tc.cuh), the recurrence on the kernels of lstm.cuh.
#include "../../include/eegclip.h"
#include "common.cuh"
#include "lin_tc.cuh"
#include "lstm.cuh"

using namespace eegclip;

#define TRY(x) do { int _r = (x); if (_r != EEGCLIP_OK) return _r; } while (0)

namespace {

struct LstmLayout {
  int GS;                         // row stride of the gate buffer: max(8H, 64)
  size_t g, cs, hp, save_total;   // save offsets (floats)
  size_t wp, wpt, bsum, wgp, scratch_total;   // scratch offsets (bytes)
};

bool lstm_ok(const eegclip_bilstm_desc* d) {
  if (!d || d->B <= 0 || d->T <= 0) return false;
  if (d->math == EEGCLIP_MATH_FP32) return false;                       // the GEMMs of this path are the tensor-core ones
  if (d->H == 128) return d->In == 64 || d->In == 128;
  if (d->H == 4) return d->In == 64 || d->In == 128 || d->In == 192 || d->In == 256;
  return false;
}

LstmLayout lstm_layout(const eegclip_bilstm_desc& d) {
  LstmLayout L;
  const size_t M = (size_t)d.B * d.T;
  L.GS = 8 * d.H < 64 ? 64 : 8 * d.H;
  size_t o = 0;
  L.g = o; o += M * L.GS;
  L.cs = o; o += M * 2 * d.H;
  L.hp = o; o += M * 2 * d.H;
  L.save_total = o;
  size_t s = 0;
  auto take = [&](size_t bytes) { size_t r = s; s += align_up(bytes, 256); return r; };
  L.wp = take((size_t)L.GS * d.In * 4);
  L.wpt = take((size_t)L.GS * d.In * 4);
  L.bsum = take((size_t)L.GS * 4);
  const size_t p1 = lintc::lin_wgrad_partial_bytes(d.H == 128 ? 256 : 64, d.In < 256 ? d.In : 256, 1);
  const size_t p2 = lintc::lin_wgrad_partial_bytes(256, 64, 2);
  L.wgp = take(p1 > p2 ? p1 : p2);
  L.scratch_total = s;
  return L;
}

__global__ void lstm_bias_kernel(const float* __restrict__ bif, const float* __restrict__ bhf, const float* __restrict__ bir,
                                 const float* __restrict__ bhr, float* __restrict__ out, int G4, int GS) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= GS) return;
  float v = 0.f;
  if (i < G4) v = bif[i] + bhf[i];
  else if (i < 2 * G4) v = bir[i - G4] + bhr[i - G4];
  out[i] = v;
}

lintc::LinTcArgs lin_plain(const float* A, long lda, const uint8_t* w, float* C, long ldc, long M, int N, int K) {
  lintc::LinTcArgs a{};
  a.A = A; a.lda = lda; a.wpacked = w; a.C = C; a.ldc = ldc; a.M = (int)M; a.N = N; a.K = K;
  a.pro = lintc::PRO_NONE; a.pro_drop = make_drop(0, 0, 0, 0.f, 0); a.drop = a.pro_drop;
  return a;
}

}  // namespace

extern "C" {

int eegclip_bilstm_supported(const eegclip_bilstm_desc* d) { return lstm_ok(d) ? 1 : 0; }

int eegclip_bilstm_workspace(const eegclip_bilstm_desc* d, size_t* save_bytes, size_t* scratch_bytes) {
  if (!lstm_ok(d)) return EEGCLIP_ERR_UNSUPPORTED;
  const LstmLayout L = lstm_layout(*d);
  if (save_bytes) *save_bytes = L.save_total * sizeof(float);
  if (scratch_bytes) *scratch_bytes = L.scratch_total;
  return EEGCLIP_OK;
}

// params: w_ih, w_hh, b_ih, b_hh, w_ih_reverse, w_hh_reverse, b_ih_reverse, b_hh_reverse
int eegclip_bilstm_forward(const eegclip_bilstm_desc* dp, const float* const* params, const float* x, float* out, void* save_v,
                           void* scratch_v, void* stream) {
  if (!lstm_ok(dp)) return EEGCLIP_ERR_UNSUPPORTED;
  if (!params || !x || !out || !save_v || !scratch_v) return EEGCLIP_ERR_ARG;
  NvtxRange nvtx_("eegclip_bilstm_forward");
  const eegclip_bilstm_desc& d = *dp;
  cudaStream_t st = (cudaStream_t)stream;
  const LstmLayout L = lstm_layout(d);
  float* save = (float*)save_v;
  uint8_t* sc = (uint8_t*)scratch_v;
  float* G = save + L.g;
  const long M = (long)d.B * d.T;
  const int G4 = 4 * d.H, In = d.In;
  float* bsum = (float*)(sc + L.bsum);
  LAUNCH_PDL((lstm_bias_kernel), ceil_div(L.GS, 256), 256, 0, st, params[2], params[3], params[6], params[7], bsum, G4, L.GS);
  LAUNCH_CHECK();
  // ---- input projections of every time step: G = x . [W_ih ; W_ih_reverse]^T + biases ----
  uint8_t* wp = sc + L.wp;
  const int NB = L.GS < 256 ? L.GS : 256;                  // output columns per GEMM (resident weights: NB x In)
  if (8 * d.H < L.GS) CUDA_TRY(cudaMemsetAsync(wp, 0, (size_t)L.GS * In * 4, st));
  {
    lintc::PackJobs J; J.n = 0;
    for (int dir = 0; dir < 2; ++dir)
      for (int r0 = 0; r0 < G4; r0 += NB) {
        const int n = dir * G4 + r0;                        // first packed row of this piece
        const int cnt = G4 - r0 < NB ? G4 - r0 : NB;
        lintc::add_pack(J, params[dir * 4] + (long)r0 * In, wp + (size_t)(n / NB) * lintc::packed_bytes(NB, In), NB, In, n % NB, 0, cnt, In,
                        In, 1);
      }
    TRY(lintc::pack_launch(J, st));
  }
  for (int nb = 0; nb < L.GS / NB; ++nb) {
    lintc::LinTcArgs a = lin_plain(x, In, wp + (size_t)nb * lintc::packed_bytes(NB, In), G + nb * NB, L.GS, M, NB, In);
    a.bias = bsum + nb * NB;
    TRY(lintc::lin_tc_launch(d.math, a, st));
  }
  // ---- recurrence ----
  if (d.H == 128) {
    ProfScope prof(PROF_LSTM, st);
    LAUNCH_PDL((lstm::lstm128_fwd_c2_kernel), dim3(2 * ceil_div(d.B, lstm::LMS), 2), 256, 0, st, params[1], params[5], G, out, save + L.cs, save + L.hp,
                                                                                d.B, d.T, g_dbg_buf);
    LAUNCH_CHECK();
  } else {
    ProfScope prof(PROF_LSTM, st);
    LAUNCH_PDL((lstm::lstm4_fwd_kernel), ceil_div(2 * d.B * 16, 128), 128, 0, st, params[1], params[5], G, L.GS, out, save + L.cs, save + L.hp, d.B,
                                                                        d.T);
    LAUNCH_CHECK();
  }
  return EEGCLIP_OK;
}

// grads: same order as params (all eight are overwritten).  dx may be NULL.  `save` is consumed (gates -> pre-activation gradients).
int eegclip_bilstm_backward(const eegclip_bilstm_desc* dp, const float* const* params, float* const* grads, const float* x,
                            const float* dout, float* dx, void* save_v, void* scratch_v, void* stream) {
  if (!lstm_ok(dp)) return EEGCLIP_ERR_UNSUPPORTED;
  if (!params || !grads || !x || !dout || !save_v || !scratch_v) return EEGCLIP_ERR_ARG;
  NvtxRange nvtx_("eegclip_bilstm_backward");
  const eegclip_bilstm_desc& d = *dp;
  cudaStream_t st = (cudaStream_t)stream;
  const LstmLayout L = lstm_layout(d);
  float* save = (float*)save_v;
  uint8_t* sc = (uint8_t*)scratch_v;
  float* G = save + L.g;
  const float* Cs = save + L.cs;
  const float* Hp = save + L.hp;
  const long M = (long)d.B * d.T;
  const int H = d.H, G4 = 4 * d.H, In = d.In;
  float* partial = (float*)(sc + L.wgp);
  const Drop nodrop = make_drop(0, 0, 0, 0.f, 0);
  // ---- recurrence: gates -> da (in place) ----
  if (H == 128) {
    ProfScope prof(PROF_LSTM, st);
    LAUNCH_PDL((lstm::lstm128_bwd_c2_kernel), dim3(2 * ceil_div(d.B, lstm::LMS), 2), 256, 0, st, params[1], params[5], G, dout, Cs, d.B, d.T);
    LAUNCH_CHECK();
  } else {
    // dW_hh: per-CTA partials in the (still unused) weight-gradient partial buffer, folded in CTA order; float atomics only if
    // that buffer were too small
    const int ctas = ceil_div(2 * d.B * 16, 128);
    const size_t p1 = lintc::lin_wgrad_partial_bytes(64, d.In < 256 ? d.In : 256, 1), p2 = lintc::lin_wgrad_partial_bytes(256, 64, 2);
    float* dwpart = (size_t)ctas * 128 * sizeof(float) <= (p1 > p2 ? p1 : p2) ? partial : nullptr;
    if (!dwpart) {
      CUDA_TRY(cudaMemsetAsync(grads[1], 0, (size_t)G4 * H * sizeof(float), st));
      CUDA_TRY(cudaMemsetAsync(grads[5], 0, (size_t)G4 * H * sizeof(float), st));
    }
    ProfScope prof(PROF_LSTM, st);
    LAUNCH_PDL((lstm::lstm4_bwd_kernel), ctas, 128, 0, st, params[1], params[5], G, L.GS, dout, Cs, Hp, grads[1], grads[5], d.B, d.T, dwpart);
    LAUNCH_CHECK();
    if (dwpart) {
      LAUNCH_PDL((lstm::lstm4_dw_fold_kernel), 1, 128, 0, st, (const float*)dwpart, ctas, grads[1], grads[5]);
      LAUNCH_CHECK();
    }
  }
  // ---- dW_ih = da^T . x, db = sum da ; dW_hh = da^T . h_prev ----
  if (H == 128) {
    for (int dir = 0; dir < 2; ++dir)
      for (int r0 = 0; r0 < G4; r0 += 256) {
        lintc::LinWgradArgs a{};
        a.dy = G + dir * G4 + r0; a.lddy = L.GS; a.Nout = 256; a.x = x; a.ldx = In; a.Kin = In; a.M = (int)M;
        a.drop_dy = nodrop; a.drop_x = nodrop; a.partial = partial;
        float* dW[3] = {grads[dir * 4] + (long)r0 * In, nullptr, nullptr};
        float* db[3] = {grads[dir * 4 + 2] + r0, nullptr, nullptr};
        TRY(lintc::lin_wgrad_launch(d.math, a, dW, db, 256, st));
        lintc::LinWgradArgs h{};
        h.dy = G + dir * G4 + r0; h.lddy = L.GS; h.Nout = 256; h.x = Hp + dir * H; h.ldx = 2 * H; h.Kin = 64; h.M = (int)M;
        h.drop_dy = nodrop; h.drop_x = nodrop; h.partial = partial;
        float* dWh[3] = {grads[dir * 4 + 1] + (long)r0 * H, nullptr, nullptr};
        float* none3[3] = {nullptr, nullptr, nullptr};
        TRY(lintc::lin_wgrad_launch(d.math, h, dWh, none3, 256, st, H, nullptr, H / 64));
      }
  } else {
    lintc::LinWgradArgs a{};
    a.dy = G; a.lddy = L.GS; a.Nout = L.GS; a.x = x; a.ldx = In; a.Kin = In; a.M = (int)M;
    a.drop_dy = nodrop; a.drop_x = nodrop; a.partial = partial;
    float* dW[3] = {grads[0], grads[4], nullptr};
    float* db[3] = {grads[2], grads[6], nullptr};
    TRY(lintc::lin_wgrad_launch(d.math, a, dW, db, G4, st));
  }
  for (int dir = 0; dir < 2; ++dir)   // b_hh receives the same gradient as b_ih
    CUDA_TRY(cudaMemcpyAsync(grads[dir * 4 + 3], grads[dir * 4 + 2], (size_t)G4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // ---- dx = da . [W_ih ; W_ih_reverse] ----
  if (dx) {
    uint8_t* wpt = sc + L.wpt;
    const int KP = L.GS < 256 ? L.GS : 256;                // contraction width per pass
    if (8 * H < L.GS) CUDA_TRY(cudaMemsetAsync(wpt, 0, (size_t)L.GS * In * 4, st));
    lintc::PackJobs J; J.n = 0;
    for (int dir = 0; dir < 2; ++dir)
      for (int r0 = 0; r0 < G4; r0 += KP) {
        const int kglob = dir * G4 + r0;
        const int cnt = G4 - r0 < KP ? G4 - r0 : KP;
        // operand (n = input feature, k = gate row) = W_ih[row][feature]
        lintc::add_pack(J, params[dir * 4] + (long)r0 * In, wpt + (size_t)(kglob / KP) * lintc::packed_bytes(In, KP), In, KP, 0, kglob % KP, In,
                        cnt, 1, In);
      }
    TRY(lintc::pack_launch(J, st));
    for (int p = 0; p < L.GS / KP; ++p) {
      lintc::LinTcArgs a = lin_plain(G + p * KP, L.GS, wpt + (size_t)p * lintc::packed_bytes(In, KP), dx, In, M, In, KP);
      a.residual = p == 0 ? nullptr : dx;
      TRY(lintc::lin_tc_launch(d.math, a, st));
    }
  }
  return EEGCLIP_OK;
}

}  // extern "C"
