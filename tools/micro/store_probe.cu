// Micro-benchmark: how fast can 8 warps per SM (the token-GEMM epilogue's population) store a 128 x 256 fp32 tile stream?
//   mode 0: the epilogue's pattern -- STG.128, a warp instruction covers 4 rows x 128 B
//   mode 1: 256-bit stores (st.global.v8.f32), a warp instruction covers 8 rows x 128 B
//   mode 2: STG.128, a warp instruction covers ONE row x 512 B (fully contiguous)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_probe store_probe.cu ; run: ./store_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 1) k(float* out, int ntiles, int mode) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3, chalf = warp >> 2;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    float* base = out + (long)tile * 128 * 256;
    for (int cb = 0; cb < 4; ++cb) {
      const int col0 = chalf * 128 + cb * 32;
      if (mode == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = q * 32 + i * 4 + (lane >> 3), col = col0 + (lane & 7) * 4;
          *reinterpret_cast<float4*>(base + row * 256 + col) = v;
        }
      } else if (mode == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = q * 32 + i * 8 + (lane >> 2), col = col0 + (lane & 3) * 8;
          asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(base + row * 256 + col), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w),
                       "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      } else {
        // one row x 128 columns per warp instruction: this warp's 32 rows x (its 128-column half), 32 instructions per tile
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = q * 32 + cb * 8 + i, col = chalf * 128 + lane * 4;
          *reinterpret_cast<float4*>(base + row * 256 + col) = v;
        }
      }
    }
  }
}
int main() {
  const int ntiles = 640 * 8;   // 8 x the FFN1 output (671 MB)
  float* out; cudaMalloc(&out, (size_t)ntiles * 128 * 256 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int mode = 0; mode < 3; ++mode) {
    k<<<148, 256, 200 * 1024>>>(out, ntiles, mode); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<148, 256, 200 * 1024>>>(out, ntiles, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d: %.1f us, %.0f GB/s (%s)\n", mode, ms * 1e3, (double)ntiles * 128 * 256 * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
