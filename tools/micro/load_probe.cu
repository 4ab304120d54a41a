// Micro-benchmark: how fast can ONE persistent CTA per SM (the token-GEMM kernels' population: ~200 KB of shared memory, so no
// second resident CTA) pull a stream of 128 x 64 fp32 tiles (32 KB, contiguous) out of HBM?
//   mode 0: register-staged -- 8 producer warps, 128-bit ld.global.nc.L1::no_allocate, CHUNKS tiles of loads in flight per thread
//           (the lin_tc_kernel producers keep 2), consumed by a dependent add
//   mode 1: cp.async.bulk (TMA engine, 1-D) into a ring of NBUF shared-memory landing buffers, consumed by a shared-memory read
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o load_probe load_probe.cu ; run: ./load_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
template <int CHUNKS>
__global__ void __launch_bounds__(256, 1) k_reg(const float* __restrict__ in, float* out, int ntiles) {
  extern __shared__ float sm[];
  float4 r[CHUNKS][8];
  float acc = 0.f;
  const int my = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](int slot, int i) {
    const float4* p = reinterpret_cast<const float4*>(in + (long)(blockIdx.x + (long)i * gridDim.x) * 8192);
#pragma unroll
    for (int j = 0; j < 8; ++j) r[slot][j] = ldg_stream(p + j * 256 + threadIdx.x);
  };
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) if (c < my) issue(c, c);
  for (int i = 0; i < my; i += CHUNKS) {
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      if (i + c < my) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += r[c][j].x + r[c][j].y + r[c][j].z + r[c][j].w;
        if (i + c + CHUNKS < my) issue(c, i + c + CHUNKS);
      }
    }
  }
  if (acc == 123.456f) out[threadIdx.x] = acc;
}
// mode 2: as mode 0, but the 32 KB tile is a 64-column strip of a (rows, 256) fp32 matrix: 128 pieces of 256 B at a 1 KB pitch --
// the access pattern of one K chunk of the K = 256 token GEMMs (FFN2, data gradients); the four strips of a 128-row block are read
// one after the other, as the kernel's chunk loop does
template <int CHUNKS>
__global__ void __launch_bounds__(256, 1) k_strip(const float* __restrict__ in, float* out, int ntiles) {
  extern __shared__ float sm[];
  float4 r[CHUNKS][8];
  float acc = 0.f;
  const int my = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](int slot, int i) {
    const long t = blockIdx.x + (long)i * gridDim.x;           // chunk index: row block t / 4, strip t % 4
    const float* base = in + (t >> 2) * (128L * 256) + (t & 3) * 64;
#pragma unroll
    for (int j = 0; j < 8; ++j) r[slot][j] = ldg_stream(reinterpret_cast<const float4*>(base + (long)(j * 16 + (threadIdx.x >> 4)) * 256) + (threadIdx.x & 15));
  };
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) if (c < my) issue(c, c);
  for (int i = 0; i < my; i += CHUNKS) {
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      if (i + c < my) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += r[c][j].x + r[c][j].y + r[c][j].z + r[c][j].w;
        if (i + c + CHUNKS < my) issue(c, i + c + CHUNKS);
      }
    }
  }
  if (acc == 123.456f) out[threadIdx.x] = acc;
}
// mode 3: the lin_tc_kernel producers' lane mapping -- lane -> (row = lane & 7, 32-byte piece = lane >> 3) of an 8-row x 128-byte
// block, fetched as TWO 128-bit loads per lane (V8 = 0: every warp instruction touches 32 sectors and uses half of each; with
// L1::no_allocate the second instruction fetches the same sectors again) or as ONE 256-bit load per lane (V8 = 1)
template <int V8>
__global__ void __launch_bounds__(256, 1) k_lane(const float* __restrict__ in, float* out, int ntiles) {
  extern __shared__ float sm[];
  float4 r[2][8][2];
  float acc = 0.f;
  const int my = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto issue = [&](int slot, int i) {
    const float* base = in + (long)(blockIdx.x + (long)i * gridDim.x) * 8192;      // 128 rows x 64 floats, dense
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = (it * 4 + (warp >> 1)) * 8 + (lane & 7), ch = (warp & 1) * 4 + (lane >> 3);
      const float* p = base + row * 64 + ch * 8;
      if (V8) {
        asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r[slot][it][0].x), "=f"(r[slot][it][0].y), "=f"(r[slot][it][0].z), "=f"(r[slot][it][0].w),
                       "=f"(r[slot][it][1].x), "=f"(r[slot][it][1].y), "=f"(r[slot][it][1].z), "=f"(r[slot][it][1].w) : "l"(p));
      } else {
        r[slot][it][0] = ldg_stream(reinterpret_cast<const float4*>(p));
        r[slot][it][1] = ldg_stream(reinterpret_cast<const float4*>(p) + 1);
      }
    }
  };
  issue(0, 0); if (my > 1) issue(1, 1);
  for (int i = 0; i < my; i += 2) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (i + c < my) {
#pragma unroll
        for (int it = 0; it < 4; ++it) acc += r[c][it][0].x + r[c][it][0].w + r[c][it][1].y + r[c][it][1].z;
        if (i + c + 2 < my) issue(c, i + c + 2);
      }
    }
  }
  if (acc == 123.456f) out[threadIdx.x] = acc;
}
template <int NBUF>
__global__ void __launch_bounds__(256, 1) k_bulk(const float* __restrict__ in, float* out, int ntiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NBUF * 32768);
  uint64_t* empty = full + NBUF;
  const int my = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NBUF; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[i])), "r"(7));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto wait = [&](uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\nD:\n\t}" ::"r"(s32(bar)), "r"(parity) : "memory");
  };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  if (warp == 0) {
    if (lane == 0)
      for (int i = 0; i < my; ++i) {
        const int b = i % NBUF;
        wait(&empty[b], ((i / NBUF) & 1) ^ 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[b])), "r"(32768) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + b * 32768)),
                     "l"(in + (long)(blockIdx.x + (long)i * gridDim.x) * 8192), "r"(32768), "r"(s32(&full[b])) : "memory");
      }
  } else {
    for (int i = 0; i < my; ++i) {
      const int b = i % NBUF;
      wait(&full[b], (i / NBUF) & 1);
      const float4* p = reinterpret_cast<const float4*>(smem + b * 32768);
      for (int j = (warp - 1) * 32 + lane; j < 2048; j += 224) { const float4 v = p[j]; acc += v.x + v.y + v.z + v.w; }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[b])) : "memory");
    }
  }
  if (acc == 123.456f) out[threadIdx.x] = acc;
}
template <class K> void run(const char* name, K kern, size_t smem, const float* in, float* out, int ntiles) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<148, 256, smem>>>(in, out, ntiles); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); kern<<<148, 256, smem>>>(in, out, ntiles); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-28s %.1f us, %.0f GB/s (%s)\n", name, ms * 1e3, (double)ntiles * 32768 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int ntiles = 640 * 32;   // 671 MB: 32 x the (81920, 64) fp32 activation of the benchmarked step
  float *in, *out; cudaMalloc(&in, (size_t)ntiles * 32768); cudaMalloc(&out, 4096); cudaMemset(in, 0, (size_t)ntiles * 32768);
  const size_t big = 200 * 1024;   // the token-GEMM kernels' footprint: one CTA per SM
  run("registers, 1 tile in flight", k_reg<1>, big, in, out, ntiles);
  run("registers, 2 tiles in flight", k_reg<2>, big, in, out, ntiles);
  run("registers, 3 tiles in flight", k_reg<3>, big, in, out, ntiles);
  run("256 B strips, 1 in flight", k_strip<1>, big, in, out, ntiles);
  run("256 B strips, 2 in flight", k_strip<2>, big, in, out, ntiles);
  run("lin_tc lanes, 2 x LDG.128", k_lane<0>, big, in, out, ntiles);
  run("lin_tc lanes, 1 x LDG.256", k_lane<1>, big, in, out, ntiles);
  run("bulk copy, 2 buffers", k_bulk<2>, big, in, out, ntiles);
  run("bulk copy, 3 buffers", k_bulk<3>, big, in, out, ntiles);
  run("bulk copy, 4 buffers", k_bulk<4>, big, in, out, ntiles);
  run("bulk copy, 6 buffers", k_bulk<6>, big, in, out, ntiles);
  return 0;
}
