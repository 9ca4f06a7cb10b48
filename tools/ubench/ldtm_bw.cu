// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM for different numbers of concurrently loading warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../animate_anything_b200/csrc ldtm_bw.cu -o ldtm_bw
#include "common.cuh"
#include <cstdio>
using namespace aab;

__global__ void __launch_bounds__(512, 1) ldtm_kernel(int iters, int active_warps, unsigned long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < active_warps) {
    const uint32_t addr = base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
      tmem_ld_32x32(addr + ((i * 32) & 511), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 8) acc += __uint_as_float(r[j]);
    }
    t1 = clock64();
  }
  if ((threadIdx.x & 31) == 0 && warp < active_warps) atomicMax(out, static_cast<unsigned long long>(t1 - t0));
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(base, 512);
  }
}

int main() {
  unsigned long long* d;
  float* sink;
  cudaMalloc(&d, 8);
  cudaMalloc(&sink, 4);
  const int iters = 4096;
  for (int aw : {1, 2, 4, 8, 16}) {
    cudaMemset(d, 0, 8);
    ldtm_kernel<<<1, 512>>>(iters, aw, d, sink);
    cudaDeviceSynchronize();
    unsigned long long c;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    const double bytes = double(aw) * iters * 4096.0;
    printf("warps=%2d  %8.1f clk per 32x32b.x32 load (per warp)  -> %6.1f B/clk/SM  (err=%s)\n", aw, double(c) / iters,
           bytes / double(c), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
