/*
 * eegclip.h -- C ABI of the B200-native EEG-CLIP hot path (libeegclip_b200.so).
 *
 * The reference (mikiken/transformer-clip-eeg) has no FFI layer: its boundary is the Python
 * nn.Module surface of clip_model.py / vlaai.py / train_clip_final.py.  Each entry point below
 * replaces the ATen op stream that one reference function launches; the citation says which.
 * The Python mirror of the reference interface (transformer-clip-eeg_b200/clip_model.py, ...)
 * binds these symbols with ctypes; INTEGRATION.md shows the stub a reference maintainer adds.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer owned by the caller;
 *     parameter/gradient tables are HOST arrays of device pointers;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *   - no allocation inside: outputs, saved activations and scratch are caller-provided, sized by
 *     the *_workspace query;
 *   - returns 0 on success, <0 on error (EEGCLIP_ERR_*), never throws;
 *   - fp32 storage everywhere (the reference is fp32); activations are time-major (B,T,C).
 */
#ifndef EEGCLIP_H_
#define EEGCLIP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EEGCLIP_ABI_VERSION 4

#if defined(__GNUC__)
#define EEGCLIP_API __attribute__((visibility("default")))
#else
#define EEGCLIP_API
#endif

enum {
  EEGCLIP_OK = 0,
  EEGCLIP_ERR_ARG = -1,
  EEGCLIP_ERR_CUDA = -2,
  EEGCLIP_ERR_UNSUPPORTED = -3
};

/* Tower kinds */
enum {
  EEGCLIP_TOWER_INTERLEAVED = 0, /* EEGConformerInterleaved, clip_model.py:400-474 */
  EEGCLIP_TOWER_SEQUENTIAL = 1   /* EEGConformer,            clip_model.py:327-398 */
};

/* GEMM arithmetic of the tensor-core kernels (conv / projections / similarity).
 *   FP32   : CUDA-core fp32 FMA (exact-fp32 companion path)
 *   BF16X3 : tcgen05 kind::f16, operands split into bf16 hi+lo, 3 MMAs, fp32 accumulate (~2^-16)
 *   BF16   : tcgen05 kind::f16, single bf16 MMA (fast; does not meet the 1e-3 gradient tolerance) */
enum {
  EEGCLIP_MATH_FP32 = 0,
  EEGCLIP_MATH_BF16X3 = 1,
  EEGCLIP_MATH_BF16 = 2
};

typedef struct {
  int32_t kind;        /* EEGCLIP_TOWER_* */
  int32_t B, T;        /* batch (windows), time samples per window */
  int32_t n_conv;      /* BasicBlocks (== depth for INTERLEAVED) */
  int32_t depth;       /* transformer blocks */
  int32_t taps;        /* Conv1d kernel size (64) */
  int32_t latent;      /* final Linear output (8) */
  int32_t train;       /* 0: eval (dropout off), 1: train */
  int32_t math;        /* EEGCLIP_MATH_* */
  int32_t reserved;
  float p_conv;        /* BasicBlock dropout (0.2) */
  float p_attn;        /* attention-probability dropout (0.5) */
  float p_proj;        /* post-projection dropout (0.5) */
  float p_ffn_hid;     /* FFN hidden dropout (0.5) */
  float p_ffn_out;     /* post-FFN dropout (0.5) */
  float reserved_f;
  uint64_t seed;       /* Philox key for this call (forward and its backward must match) */
} eegclip_tower_desc;

/* Parameter table order for both tower kinds (host array of device pointers, fp32):
 *   [0] eeg_spatial_mapping.weight (64,64,1)   [1] eeg_spatial_mapping.bias (64)
 *   then per BasicBlock i (n_conv of them), 4 entries:
 *       conv.weight (64,64,taps), conv.bias (64), normalization.weight (64,T), normalization.bias (64,T)
 *   then per TransformerEncoderBlock j (depth of them), 16 entries:
 *       ln1.w, ln1.b, queries.w, queries.b, keys.w, keys.b, values.w, values.b, projection.w, projection.b,
 *       ln2.w, ln2.b, ffn0.w (256,64), ffn0.b, ffn3.w (64,256), ffn3.b
 *   then final_layer.weight (latent,64), final_layer.bias (latent)
 * The gradient table has the same order. */
#define EEGCLIP_TOWER_NPARAMS(n_conv, depth) (2 + 4 * (n_conv) + 16 * (depth) + 2)

EEGCLIP_API int eegclip_abi_version(void);
EEGCLIP_API const char* eegclip_build_info(void);

/* Measurement hooks (bench.py): number of kernels this library has launched so far, and optional CUDA-event timing of
 * the dominant kernel classes on the stream they are launched on.  Classes: 0 conv fwd/dgrad (tcgen05), 1 conv wgrad
 * (tcgen05), 2 attention fwd, 3 attention bwd, 4 LayerNorm([C,T]) fwd+bwd, 5 head similarity kernel (tcgen05) + exact-fp32 GEMMs, 6 token GEMMs
 * (tcgen05), 7 token weight-gradient GEMMs (tcgen05), 8 LSTM recurrences.  eegclip_profile_end synchronises the device and
 * returns summed milliseconds and launch counts per class (arrays of >= 12 entries). */
EEGCLIP_API long long eegclip_launch_count(void);
/* Development knobs for kernel tuning sweeps and A/B timing; 0 everywhere = shipped configuration.  Keys:
 *   0 token-GEMM ring depth (2..3)          1 activation load policy (1 = __ldg)      3 force the generic token-GEMM instantiation
 *   4 producer / epilogue warp split rule   7 1 = no programmatic dependent launch    8 bit mask of kernel classes that record
 *   profiling events (0 = all)              9 1 = register-staged token weight gradient
 *  10 ln64 backward rows per CTA           11 column-sum rows per CTA               12 1 = strided weight-gradient operands by
 *      per-row bulk copies instead of 2-D tensor maps (A/B timing)             13 1 = token-GEMM producers load 2 x 128 bit
 *      instead of 1 x 256 bit per lane (A/B timing)
 *   2 bit 0 / 1: lin_tc_kernel / lin_wgrad_tma_kernel issue griddepcontrol.launch_dependents after their last MMA (A/B timing)
 *   5 1 = programmatic launch also into the head's similarity kernels      6 1 = stream breaks as 4-byte memsets (as first found)
 *  14 1 = similarity kernel always with 256-column tiles (the 64-column small-batch form off)
 *  15 k > 0 = bit mask k - 1 of the launch edges that go out in plain stream order instead of programmatic dependent launch
 *      (1 conv block backward start, 4 before the attention backward, 8 conv block forward start; 0 = the shipped set 1|4|8)
 * (the Python layer reads EEGCLIP_TUNE="key=value,..." from the environment at load time).
 * THREADING: the knobs, the launch counter and the profiling state are PROCESS-GLOBAL development state, written without
 * synchronisation -- set them before compute starts, from one thread.  All compute entry points are reentrant per
 * (device, stream): they keep no state of their own besides one-time cudaFuncSetAttribute calls and read the knobs only. */
EEGCLIP_API int eegclip_tune_set(int32_t key, int32_t value);
/* Development: device buffer (>= 3*256 uint64) that CTA 0 of the token-GEMM kernel fills with a (event, globaltimer) timeline
 * of its producer / MMA / epilogue roles; NULL (default) disables it. */
EEGCLIP_API int eegclip_debug_buffer(void* dev_ptr);
EEGCLIP_API int eegclip_profile_begin(void);
EEGCLIP_API int eegclip_profile_end(double* ms_by_class, long long* launches_by_class, int32_t n_classes);

/* Bytes of saved activations (forward -> backward) and of scratch for one tower call. */
EEGCLIP_API int eegclip_tower_workspace(const eegclip_tower_desc* d, size_t* save_bytes, size_t* scratch_bytes);

/* Replaces EEGConformerInterleaved.forward / EEGConformer.forward (clip_model.py:445-474 / 373-398).
 * x (B,T,64) -> out (B,T,latent).  `save` may be NULL when no backward follows (inference). */
EEGCLIP_API int eegclip_tower_forward(const eegclip_tower_desc* d, const float* const* params, const float* x, float* out,
                          void* save, void* scratch, void* stream);

/* Autograd of the above: given dout (B,T,latent) fills every entry of `grads` (overwrites; `grad_base`/
 * `grad_bytes` describe one contiguous region covering all of them, zero-filled first) and, if dx != NULL,
 * dx (B,T,64). */
EEGCLIP_API int eegclip_tower_backward(const eegclip_tower_desc* d, const float* const* params, float* const* grads,
                           void* grad_base, size_t grad_bytes, const float* x, const float* dout, float* dx,
                           const void* save, void* scratch, void* stream);

/* One TransformerEncoderBlock (clip_model.py:75-94) on its own: zin (B,T,64) -> zout (B,T,64).
 * params/grads: the 16-entry block of the tower table (ln1.w ... ffn3.b).  Used by the stand-alone
 * TransformerEncoderBlock / TransformerEncoder modules; the towers call the same kernels internally. */
typedef struct {
  int32_t B, T, layer, train, math, reserved;
  float p_attn, p_proj, p_ffn_hid, p_ffn_out;
  uint64_t seed;
} eegclip_xfblock_desc;

EEGCLIP_API int eegclip_xfblock_workspace(const eegclip_xfblock_desc* d, size_t* save_bytes, size_t* scratch_bytes);
EEGCLIP_API int eegclip_xfblock_forward(const eegclip_xfblock_desc* d, const float* const* params, const float* zin, float* zout,
                            void* save, void* scratch, void* stream);
EEGCLIP_API int eegclip_xfblock_backward(const eegclip_xfblock_desc* d, const float* const* params, float* const* grads, void* grad_base,
                             size_t grad_bytes, const float* zin, const float* dzout, float* dzin, const void* save,
                             void* scratch, void* stream);

/* Stand-alone pieces of the transformer block, for callers that use the reference's sub-modules directly
 * (MultiHeadAttention.forward clip_model.py:30-45, ResidualAdd.forward :52-57, nn.LayerNorm(64) :84,89, nn.Dropout :86,91).
 *   attention : qkv (B,T,192) = [q | k | v], head h owning columns 8h..8h+7 of each third -> out (B,T,64); softmax(QK^T/sqrt(64)),
 *               dropout on the probabilities (Philox site ATTN of `layer`); lse (B,8,T) links forward and backward.
 *   layernorm : per-token LayerNorm over C = 64 features, eps 1e-5; dgamma/dbeta are overwritten.
 *   dropout   : out = in * keep/(1-p) with the Philox stream (layer, site); the same call is its own backward. */
EEGCLIP_API int eegclip_attention_forward(const float* qkv, float* out, float* lse, int32_t B, int32_t T, float p_drop, int32_t train,
                              int32_t layer, uint64_t seed, int32_t math, void* stream);
EEGCLIP_API int eegclip_attention_backward(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv, int32_t B,
                               int32_t T, float p_drop, int32_t train, int32_t layer, uint64_t seed, int32_t math, void* stream);
EEGCLIP_API int eegclip_layernorm_forward(const float* x, const float* gamma, const float* beta, float* out, int64_t rows, int32_t C,
                              void* stream);
EEGCLIP_API int eegclip_layernorm_backward(const float* dout, const float* x, const float* gamma, float* dx, float* dgamma, float* dbeta,
                               int64_t rows, int32_t C, void* stream);
EEGCLIP_API int eegclip_dropout(const float* in, float* out, int64_t n, float p, int32_t train, int32_t layer, int32_t site, uint64_t seed,
                    void* stream);
/* nn.GELU() followed by nn.Dropout (FeedForwardBlock, clip_model.py:64-65) in one pass: out = dropout(GELU(pre)); backward
 * dpre = dout * keep/(1-p) * GELU'(pre) (mask regenerated from the Philox stream (layer, site)). */
EEGCLIP_API int eegclip_gelu_dropout_forward(const float* pre, float* out, int64_t n, float p, int32_t train, int32_t layer, int32_t site,
                                 uint64_t seed, void* stream);
EEGCLIP_API int eegclip_gelu_dropout_backward(const float* pre, const float* dout, float* dpre, int64_t n, float p, int32_t train,
                                  int32_t layer, int32_t site, uint64_t seed, void* stream);

/* BasicBlock (clip_model.py:234-249) / VLAAI conv+LN+LeakyReLU (vlaai.py:29-35,60-72) on a time-major tensor:
 *   y = act(LayerNorm_[C,T](dropout(conv1d_same(x (+ skip_in)))))      act: 0 GELU, 1 LeakyReLU(0.01)
 * x (B,T,Cin) -> out (B,T,Cout).  w (Cout,Cin,taps), gamma/beta (Cout,T).
 * save: conv output y (B*T*Cout floats) followed by 2*B stats; scratch: see eegclip_convblock_workspace. */
typedef struct {
  int32_t B, T, Cin, Cout, taps, act, train, math;
  float p_drop;
  int32_t layer;       /* Philox stream = layer*16 + SITE_CONV */
  uint64_t seed;
} eegclip_convblock_desc;

EEGCLIP_API int eegclip_convblock_workspace(const eegclip_convblock_desc* d, size_t* save_bytes, size_t* scratch_bytes);
EEGCLIP_API int eegclip_convblock_forward(const eegclip_convblock_desc* d, const float* x, const float* skip_in, const float* w,
                              const float* bias, const float* gamma, const float* beta, float* out, void* save,
                              void* scratch, void* stream);
/* dw/dbias/dgamma/dbeta are overwritten. dx (B,T,Cin) is the gradient w.r.t. (x + skip_in). */
EEGCLIP_API int eegclip_convblock_backward(const eegclip_convblock_desc* d, const float* x, const float* skip_in, const float* w,
                               const float* gamma, const float* beta, const float* dout, float* dx, float* dw,
                               float* dbias, float* dgamma, float* dbeta, const void* save, void* scratch, void* stream);

/* Linear over tokens: out[m][n] = sum_k x[m][k] w[n][k] + b[n]  (1x1 Conv1d / nn.Linear; clip_model.py:421,439,267;
 * vlaai.py:18,91,94,104).  `scratch` (eegclip_linear_workspace bytes) holds the packed tensor-core operands and the
 * weight-gradient partials; with scratch == NULL or math == FP32 the exact-fp32 CUDA-core kernels run. */
EEGCLIP_API int eegclip_linear_workspace(int64_t M, int32_t N, int32_t K, size_t* scratch_bytes);
EEGCLIP_API int eegclip_linear_forward(const float* x, const float* w, const float* b, float* out, int64_t M, int32_t N, int32_t K,
                           int32_t math, void* scratch, void* stream);
EEGCLIP_API int eegclip_linear_backward(const float* x, const float* w, const float* dout, float* dx, float* dw, float* db, int64_t M,
                            int32_t N, int32_t K, int32_t math, void* scratch, void* stream);

/* Bidirectional single-layer LSTM with zero initial state (nn.LSTM(batch_first=True, bidirectional=True); speech tower,
 * clip_model.py:267-268, 322-323).  x (B,T,In) -> out (B,T,2H), forward direction in columns [0,H), reverse in [H,2H).
 * params / grads (host arrays of 8 device pointers): weight_ih_l0 (4H,In), weight_hh_l0 (4H,H), bias_ih_l0, bias_hh_l0,
 * then the four *_reverse tensors; gate order i,f,g,o.  Covered shapes: H = 128 with In in {64,128} (speech_lstm1) and
 * H = 4 with In in {64,...,256} (speech_lstm2); eegclip_bilstm_supported tells, other shapes are the caller's business
 * (the Python mirror keeps them on the cuDNN library call).  `save` links forward and backward and is consumed by the
 * backward; all eight gradients are overwritten; dx may be NULL. */
typedef struct {
  int32_t B, T, In, H, math, reserved;
} eegclip_bilstm_desc;

EEGCLIP_API int eegclip_bilstm_supported(const eegclip_bilstm_desc* d);
EEGCLIP_API int eegclip_bilstm_workspace(const eegclip_bilstm_desc* d, size_t* save_bytes, size_t* scratch_bytes);
EEGCLIP_API int eegclip_bilstm_forward(const eegclip_bilstm_desc* d, const float* const* params, const float* x, float* out, void* save,
                           void* scratch, void* stream);
EEGCLIP_API int eegclip_bilstm_backward(const eegclip_bilstm_desc* d, const float* const* params, float* const* grads, const float* x,
                            const float* dout, float* dx, void* save, void* scratch, void* stream);

/* Symmetric InfoNCE head (clip_model.py:675-693, 913-930), local or sharded.
 *   rows : this rank's raw (un-normalised) flattened embeddings, S_loc / E_loc (b,D)
 *   all  : the gathered NORMALISED embeddings S_all / E_all (Bg,D) (== the local ones when world==1)
 * Step 1  eegclip_l2norm_forward : raw (b,D) -> normalised (b,D) + inverse norms (b)
 * Step 2  eegclip_infonce_lse    : row LSE of this rank's speech rows vs all EEG, column LSE of this rank's
 *                                  EEG columns vs all speech, and the diagonal; logits never touch HBM.
 * Step 3  (caller all-gathers lse_row/lse_col/diag over ranks; no-op for world==1)
 * Step 4  eegclip_infonce_loss   : loss = ((lse_row-diag).mean + (lse_col-diag).mean)/2 on the full vectors
 * Step 5  eegclip_infonce_backward : dS_loc, dE_loc w.r.t. the NORMALISED local rows and dtau (partial sum
 *                                  over this rank's rows), recomputing logits tiles.
 * Step 6  eegclip_l2norm_backward : gradient w.r.t. the raw embeddings.
 * tau is the learnable log-scale (device scalar); logits = S.E^T * exp(tau). */
EEGCLIP_API int eegclip_l2norm_forward(const float* x, float* xn, float* inv_norm, int32_t rows, int32_t D, void* stream);
EEGCLIP_API int eegclip_l2norm_backward(const float* xn, const float* inv_norm, const float* dxn, float* dx, int32_t rows, int32_t D,
                            void* stream);
EEGCLIP_API int eegclip_infonce_workspace(int32_t b, int32_t Bg, int32_t D, size_t* scratch_bytes);
/* one_sided != 0 computes the one-directional cross-entropy CE(X.E^T * exp(tau), arange) used for the memory-bank term
 * (clip_model.py:934-937): S_all then holds the row operand X (no gradient), lse_col is neither written nor read. */
EEGCLIP_API int eegclip_infonce_lse(const float* S_all, const float* E_all, const float* tau, int32_t b, int32_t row0, int32_t Bg,
                        int32_t D, float* lse_row, float* lse_col, float* diag, int32_t math, int32_t one_sided,
                        void* scratch, void* stream);
EEGCLIP_API int eegclip_infonce_loss(const float* lse_row_all, const float* lse_col_all, const float* diag_all, int32_t Bg,
                         int32_t one_sided, float* loss, void* stream);
/* dloss: DEVICE scalar holding the upstream gradient of the loss (no host sync).  dS_loc may be NULL when one_sided. */
EEGCLIP_API int eegclip_infonce_backward(const float* S_all, const float* E_all, const float* tau, const float* lse_row_all,
                             const float* lse_col_all, int32_t b, int32_t row0, int32_t Bg, int32_t D, const float* dloss,
                             float* dS_loc, float* dE_loc, float* dtau_partial, int32_t math, int32_t one_sided,
                             void* scratch, void* stream);

/* memoryBank.forward (clip_model.py:731-745): old = memory[idx]; memory[idx] = m*old + (1-m)*data. idx int64, memory
 * (bank_rows, D).  Every old row is gathered before any row is written (index_select, then index_copy_): duplicate ids in a
 * batch all return the pre-batch row and the last occurrence's update is the one stored.  An id outside [0, bank_rows)
 * (IndexError in the reference) writes nothing and returns a NaN row.
 * one_minus_momentum is passed separately because the reference rounds (1 - m) from a Python double. */
EEGCLIP_API int eegclip_membank_update(float* memory, int64_t bank_rows, const int64_t* idx, const float* data, float* old_out,
                           int32_t rows, int32_t D, float momentum, float one_minus_momentum, void* stream);

/* AdamW / Adam over a table of tensors (train_clip_final.py:403-413,492; torch.optim.AdamW semantics with decoupled
 * weight decay, or torch.optim.Adam semantics -- decay added to the gradient -- when coupled_decay != 0; amsgrad when the
 * entry carries a vmax buffer).  `table_dev` is a DEVICE-resident array of n_tensors entries (the caller builds it on the
 * host and uploads it once; it stays valid while the parameter/gradient/state pointers do).  One launch updates
 * every tensor: 16 B loads/stores, 7 x 4 B of traffic per element (HBM-bound). */
typedef struct {
  void* p;          /* parameter, fp32, updated in place */
  const void* g;    /* gradient, fp32 */
  void* m;          /* exp_avg, fp32 */
  void* v;          /* exp_avg_sq, fp32 */
  int64_t numel;
  void* vmax;       /* max_exp_avg_sq, fp32 (amsgrad), or NULL */
} eegclip_adamw_entry;

EEGCLIP_API int eegclip_adamw_step(const eegclip_adamw_entry* table_dev, int32_t n_tensors, int64_t max_numel, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int32_t coupled_decay, int64_t step, void* stream);

/* Match-mismatch scoring (train_clip_helper_functions.py:153-163,176-187).
 *   eegclip_mm_rowdots : scores[k][n] = <eeg[n], cand[n][k]>  (replaces the N x N matmul + diag) and argmax over k
 *   eegclip_mm_bank_logits : logits (N,M) = eeg (N,D) . bank (M,D)^T  (top-k is taken by the caller); with `scratch`
 *                            (eegclip_mm_bank_workspace bytes) and math != FP32 the tcgen05 similarity kernel runs */
EEGCLIP_API int eegclip_mm_rowdots(const float* eeg, const float* cand, float* scores, int64_t* choice, int32_t N, int32_t K, int32_t D,
                       void* stream);
EEGCLIP_API int eegclip_mm_bank_workspace(int32_t N, int32_t M, int32_t D, size_t* scratch_bytes);
EEGCLIP_API int eegclip_mm_bank_logits(const float* eeg, const float* bank, float* logits, int32_t N, int32_t M, int32_t D, int32_t math,
                           void* scratch, void* stream);

/* Per-subject mean-variance normalisation of EEG windows (train_clip_helper_functions.py:133-136):
 * x (rows = N*T, C) -> y = (x - mean_c) / std_c with the per-channel mean and population std over all rows
 * (sums in fp64, fixed order).  scratch: eegclip_mvn_workspace bytes. */
EEGCLIP_API int eegclip_mvn_workspace(int32_t C, size_t* scratch_bytes);
EEGCLIP_API int eegclip_mvn_normalize(const float* x, float* y, int64_t rows, int32_t C, void* scratch, void* stream);

/* Per-row top-k of a similarity matrix (train_clip_helper_functions.py:182-187: torch.topk(logits, min(100, M))):
 * x (N, M) with row stride ld -> vals (N,k) descending, idx (N,k) = column + col_offset; ties go to the lower column.
 * k <= 1024, k <= M.  Reads every row twice (radix/threshold select), writes nothing else. */
EEGCLIP_API int eegclip_row_topk(const float* x, int64_t ld, int32_t N, int32_t M, int32_t k, int64_t col_offset, float* vals,
                     int64_t* idx, void* stream);

/* Downstream regression head (train_clip_helper_functions.py:1132-1140, 1107-1118; used by :620-640):
 *   conv_small : out = LeakyReLU_0.01(Conv1d(Cin -> Cout, K, padding='same')(x)), channel-major x (B,Cin,T) -> out (B,Cout,T);
 *                backward overwrites dw (Cout,Cin,K), db (Cout) (may be NULL) and, if dx != NULL, dx (B,Cin,T).
 *   pearson    : PearsonLoss: r[b][c] = cosine(x - mean_t x, y - mean_t y) (eps 1e-6), loss[c] = -mean_b r[b][c];
 *                backward gives dx (B,C,T) for an upstream gradient dloss (C) (device vector). */
EEGCLIP_API int eegclip_conv_small_workspace(int32_t B, int32_t Cin, int32_t Cout, int32_t K, size_t* scratch_bytes);
EEGCLIP_API int eegclip_conv_small_forward(const float* x, const float* w, const float* bias, float* out, int32_t B, int32_t Cin,
                               int32_t Cout, int32_t T, int32_t K, void* stream);
EEGCLIP_API int eegclip_conv_small_backward(const float* x, const float* w, const float* out, const float* dout, float* dx, float* dw,
                                float* db, int32_t B, int32_t Cin, int32_t Cout, int32_t T, int32_t K, void* scratch, void* stream);
EEGCLIP_API int eegclip_pearson_forward(const float* x, const float* y, float* r, float* loss, int32_t B, int32_t C, int32_t T,
                            void* stream);
EEGCLIP_API int eegclip_pearson_backward(const float* x, const float* y, const float* dloss, float* dx, int32_t B, int32_t C, int32_t T,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EEGCLIP_H_ */
