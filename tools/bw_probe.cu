// HBM read-bandwidth probe for the weight-streaming design: persistent CTAs stream a large buffer through a
// shared-memory ring with 1-D bulk async copies (UBLKCP); no math.  Sweeps chunk size, ring depth, CTAs/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_probe tools/bw_probe.cu && ./bw_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "lvllm_b200/csrc/common.cuh"
using namespace b200;

__global__ void __launch_bounds__(64) probe_kernel(const uint8_t* __restrict__ src, size_t total, int chunk,
                                                   int stages, int split, unsigned long long* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * chunk);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const size_t n_chunks = total / chunk;
  const int warp = threadIdx.x >> 5;
  if (warp == 0 && threadIdx.x == 0) {
    const uint64_t pol = policy_evict_first();
    uint32_t it = 0;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
      const int s = it % stages;
      mbar_wait(&empty[s], ((it / stages) & 1) ^ 1);
      mbar_arrive_expect_tx(&full[s], chunk);
      const int sub = chunk / split;
      for (int k = 0; k < split; ++k)
        bulk_g2s_hint(smem + (size_t)s * chunk + k * sub, src + c * (size_t)chunk + k * sub, sub, &full[s], pol);
    }
  } else if (warp == 1 && threadIdx.x == 32) {
    uint32_t it = 0;
    unsigned long long acc = 0;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
      const int s = it % stages;
      mbar_wait(&full[s], (it / stages) & 1);
      acc += *reinterpret_cast<volatile uint32_t*>(smem + (size_t)s * chunk);
      mbar_arrive(&empty[s]);
    }
    if (acc == 0x1234567) *sink = acc;
  }
}

__global__ void __launch_bounds__(512) ldg_kernel(const uint4* __restrict__ src, size_t n16, unsigned long long* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 7 * stride < n16; i += 8 * stride) {
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(src + i + k * stride));
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc.x ^= v[k].x; acc.y ^= v[k].y; acc.z ^= v[k].z; acc.w ^= v[k].w; }
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x1234567) *sink = 1;
}

int main() {
  const size_t total = (size_t)4 << 30;
  uint8_t* buf;
  unsigned long long* sink;
  cudaMalloc(&buf, total);
  cudaMalloc(&sink, 8);
  cudaMemset(buf, 1, total);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  auto run = [&](int ctas, int chunk, int stages, int split) {
    size_t smem = (size_t)stages * chunk + stages * 16 + 1024;
    float best = 1e9;
    for (int r = 0; r < 3; ++r) {
      cudaEventRecord(e0);
      probe_kernel<<<ctas, 64, smem>>>(buf, total, chunk, stages, split, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("ctas=%4d chunk=%6d stages=%2d split=%d smem=%6zu : %8.1f GB/s %s\n", ctas, chunk, stages, split, smem,
           total / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  {
    float best = 1e9;
    for (int r = 0; r < 3; ++r) {
      cudaEventRecord(e0);
      ldg_kernel<<<148 * 4, 512>>>(reinterpret_cast<const uint4*>(buf), total / 16, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("LDG.128 x8 unrolled read-only: %8.1f GB/s\n", total / best / 1e6);
  }
  for (int ctas : {148, 296}) {
    const int budget = ctas == 148 ? 200 * 1024 : 100 * 1024;
    for (int chunk : {4096, 8192, 16384, 32768}) {
      for (int split : {1, 4}) {
        if (chunk / split < 2048) continue;
        run(ctas, chunk, budget / chunk, split);
      }
    }
  }
  run(128, 32768, 6, 1);
  run(148, 32768, 3, 1);
  run(148, 16384, 6, 1);
  run(148, 16384, 3, 1);
  run(592, 8192, 6, 1);
  return 0;
}
