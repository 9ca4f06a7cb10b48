// Micro-benchmark: raw tcgen05.mma (kind::f16, cta_group::1, SS) issue rate per (M, N) -- how long does one K=16 step
// take for M in {64, 128} and N in {64, 128, 256}?  Round-1 finding (profiles/r01_igemm_issue_trace.md): with M=128
// the step costs ~128 clk whatever N.  This checks whether M=64 halves it (the premise of the transposed remainder
// tile and of the transposed PV product planned in DESIGN.md section 7b), and whether keeping A in tensor memory (TS mode,
// floor N/2 cycles according to /opt/skills/guides/B300_MICROARCH.md) removes the flat ~128 clk of the SS mode.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../animate_anything_b200/csrc umma_rate.cu -o umma_rate
#include "common.cuh"
#include <cstdio>
using namespace aab;

// A operand in tensor memory (TS mode): D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 2) umma_rate_kernel(int M, int N, int reps, int a_in_tmem, int nacc, int tmem_cols,
                                                           unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;                 // 128 rows x 128 B (K-major, SWIZZLE_128B layout; content irrelevant, zeroed)
  uint8_t* smB = smem + 16384;         // 256 rows x 128 B
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if ((threadIdx.x >> 5) == 0) {
    tmem_alloc(&slot, tmem_cols);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if ((threadIdx.x >> 5) == 0 && elect_one()) {   // warp-uniform branch + elect: no per-instruction waterfall loop in the SASS
    const uint32_t idesc = make_idesc_f16(1, M, N, 0, 0);
    const uint32_t a_addr = smem_u32(smA), b_addr = smem_u32(smB);
    const long long t0 = clock64();
    if (nacc > 1) {
      // `nacc` independent accumulators (TMEM columns j*N), consecutive instructions never share one: separates the
      // dependent-accumulate latency from the issue / operand-read rate
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t d = tmem + static_cast<uint32_t>(((r * 4 + k) % nacc) * N);
          umma_f16_ss(d, make_desc_kmajor_sw128(a_addr + k * 32), make_desc_kmajor_sw128(b_addr + k * 32), idesc,
                      (r * 4 + k >= nacc) ? 1u : 0u);
        }
      }
    } else
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (a_in_tmem)   // A (128 lanes x 8 columns of packed 16-bit pairs per K=16 step) at TMEM columns 256.., D at 0..255
          umma_f16_ts(tmem, tmem + 256 + k * 8, make_desc_kmajor_sw128(b_addr + k * 32), idesc, (r > 0 || k > 0) ? 1u : 0u);
        else
          umma_f16_ss(tmem, make_desc_kmajor_sw128(a_addr + k * 32), make_desc_kmajor_sw128(b_addr + k * 32), idesc,
                      (r > 0 || k > 0) ? 1u : 0u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, tmem_cols);
  }
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 8 * 512);
  const int smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 2048;   // x 4 instructions
  for (int ts : {0, 1})
  for (int M : {128, 64}) {
    for (int N : {256, 192, 128, 64, 32}) {
      cudaMemset(d, 0, 8);
      umma_rate_kernel<<<1, 128, smem>>>(M, N, reps, ts, 1, 512, d);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long c = 0;
      cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const double per = double(c) / (reps * 4.0);
      printf("%s M=%3d N=%3d  %7.1f clk per K=16 instruction  -> %7.0f FLOP/clk/SM  (%s)\n", ts ? "A in TMEM" : "A in SMEM", M, N, per,
             2.0 * M * N * 16 / per, cudaGetErrorString(e));
    }
  }
  for (int nacc : {2, 4})
    for (int N : {256, 128, 64}) {
      if (nacc * N > 512) continue;
      cudaMemset(d, 0, 8);
      umma_rate_kernel<<<1, 128, smem>>>(128, N, reps, 0, nacc, 512, d);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long c = 0;
      cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const double per = double(c) / (reps * 4.0);
      printf("A in SMEM M=128 N=%3d  %d independent accumulators  %7.1f clk per K=16 instruction  -> %7.0f FLOP/clk/SM  (%s)\n", N, nacc,
             per, 2.0 * 128 * N * 16 / per, cudaGetErrorString(e));
    }
  // two CTAs resident on every SM (grid = 2 x SM count, 256 TMEM columns each), one dependent accumulate chain per CTA: does the
  // tensor pipe interleave the two CTAs' instructions (each chain then sees the other's instruction between two of its own)?
  {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int N : {256, 128, 64}) {
      for (int ctas_per_sm : {1, 2}) {
        const int grid = sms * ctas_per_sm;
        cudaMemset(d, 0, 8 * 512);
        umma_rate_kernel<<<grid, 128, smem>>>(128, N, reps, 0, 1, 256, d);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long c[512];
        cudaMemcpy(c, d, 8 * grid, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0, mn = ~0ull;
        for (int i = 0; i < grid; ++i) { mx = c[i] > mx ? c[i] : mx; mn = c[i] < mn ? c[i] : mn; }
        printf("A in SMEM M=128 N=%3d  %d CTA(s) per SM, one chain each: %7.1f .. %7.1f clk per instruction per CTA -> %7.0f FLOP/clk/SM  (%s)\n",
               N, ctas_per_sm, double(mn) / (reps * 4.0), double(mx) / (reps * 4.0),
               ctas_per_sm * 2.0 * 128 * N * 16 / (double(mx) / (reps * 4.0)), cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
