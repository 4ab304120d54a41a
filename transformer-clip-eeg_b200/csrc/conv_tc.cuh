// tcgen05 implicit-GEMM Conv1d(64,64,k=64,'same') -- interface used by tower.cu.
#pragma once
#include "common.cuh"

namespace eegclip {

inline bool conv_tc_supported(int Cin, int Cout, int taps, int T) { (void)Cin; (void)Cout; (void)taps; (void)T; return false; }
inline size_t conv_tc_scratch_bytes(int B, int T, int taps) { (void)B; (void)T; (void)taps; return 0; }
inline int conv_tc_forward(int math, const float* xin, const float* skip_in, const float* w, const float* bias, float* y, int B, int T,
                           int PL, const Drop& drop, float* scratch, cudaStream_t st) {
  return EEGCLIP_ERR_UNSUPPORTED;
}
inline int conv_tc_backward(int math, const float* xin, const float* skip_in, const float* w, const float* dypad, int PLb, float* du,
                            float* dw, int B, int T, float* scratch, float* wtmp, cudaStream_t st) {
  return EEGCLIP_ERR_UNSUPPORTED;
}

}  // namespace eegclip
