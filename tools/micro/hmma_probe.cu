// Micro-benchmark: latency and throughput of the warp-level mma.sync.m16n8k16 (bf16) and m16n8k8 (tf32) on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_probe hmma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CH, int KIND>
__global__ void probe(float* out, long long* clk, int iters) {
  float acc[CH][4];
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[c][j] = (float)threadIdx.x;
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 2)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                     : "r"(a0), "r"(a1), "r"(b0));
      else if (KIND == 3) {
        uint32_t t;
        asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(t) : "r"(__float_as_uint(acc[c][0])));
        acc[c][0] = __uint_as_float(t);
      } else
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int CH, int KIND>
void run(int warps, float* out, long long* clk) {
  const int iters = 2000;
  probe<CH, KIND><<<148, warps * 32>>>(out, clk, iters);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / iters;
  const double flop = (KIND == 0 ? 4096.0 : 2048.0) * CH * warps;   // per SM per iteration
  if (KIND >= 2) { printf("%s chains/warp %2d warps/SM %2d: %.1f clk per instruction per warp\n", KIND == 2 ? "f16 m16n8k8" : "movmatrix  ", CH, warps, per / CH); return; }
  printf("%s chains/warp %2d warps/SM %2d: %.1f clk per round of %d MMAs per warp -> %.1f clk per MMA per warp, %.0f dense FLOP/clk/SM\n",
         KIND == 0 ? "bf16 m16n8k16" : "tf32 m16n8k8 ", CH, warps, per, CH, per / CH, flop / per);
}

int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  run<1, 0>(1, out, clk); run<2, 0>(1, out, clk); run<4, 0>(1, out, clk); run<8, 0>(1, out, clk); run<12, 0>(1, out, clk);
  run<1, 0>(4, out, clk); run<4, 0>(4, out, clk); run<8, 0>(4, out, clk); run<12, 0>(8, out, clk); run<8, 0>(16, out, clk); run<8, 0>(32, out, clk);
  run<1, 1>(1, out, clk); run<8, 1>(1, out, clk); run<8, 1>(8, out, clk); run<8, 1>(16, out, clk); run<8, 1>(32, out, clk);
  run<1, 2>(1, out, clk); run<8, 2>(1, out, clk); run<8, 2>(8, out, clk); run<8, 2>(16, out, clk);
  run<1, 3>(1, out, clk); run<8, 3>(1, out, clk); run<8, 3>(8, out, clk); run<8, 3>(16, out, clk);
  return 0;
}
