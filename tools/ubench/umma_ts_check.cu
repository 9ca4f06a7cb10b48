// Correctness probe for the TS form of tcgen05.mma (A operand in tensor memory), the basis of keeping P = softmax(S)
// in TMEM for the P.V product.  Assumption under test: for kind::f16 the A operand of an
// M=128 instruction is read from TMEM with lane = row m and 32-bit column j holding the pair (A[m][2j], A[m][2j+1])
// (low half = even k), i.e. exactly what `tcgen05.st.32x32b` writes when thread m stores its row as packed pairs; a
// K=16 step consumes 8 consecutive columns.  D = A (128x64) x B^T (64x64), small integers, compared exactly.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../animate_anything_b200/csrc umma_ts_check.cu -o umma_ts_check
#include "common.cuh"
#include <cstdio>
using namespace aab;

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
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __host__ inline int a_val(int m, int k) { return (m + 2 * k) % 7 - 3; }
__device__ __host__ inline int b_val(int n, int k) { return (3 * n + 7 * k) % 11 - 5; }

constexpr int A_COL = 256;   // TMEM column where the A operand starts (D occupies columns 0..63)

__global__ void __launch_bounds__(128, 1) ts_check_kernel(int* mismatches, float* sample) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smB = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  const int m = threadIdx.x;           // row of A / D, TMEM lane
  const int warp = threadIdx.x >> 5;
  // B [64 n][64 k] bf16, K-major, 128-byte rows, SWIZZLE_128B: 16-byte chunk c of row n lives at chunk (c ^ (n & 7))
  if (m < 64) {
    for (int c = 0; c < 8; ++c) {
      uint32_t w[4];
      for (int e = 0; e < 4; ++e) w[e] = pack2(float(b_val(m, c * 8 + 2 * e)), float(b_val(m, c * 8 + 2 * e + 1)), true);
      *reinterpret_cast<uint4*>(smB + m * 128 + ((c ^ (m & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;   // this warp's TMEM lane quarter
  {
    uint32_t r[32];
    for (int j = 0; j < 32; ++j) r[j] = pack2(float(a_val(m, 2 * j)), float(a_val(m, 2 * j + 1)), true);
    tmem_st_32x32(tmem + lane_base + A_COL, r);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(1, 128, 64, 0, 0);
    const uint32_t b_addr = smem_u32(smB);
    for (int k = 0; k < 4; ++k)
      umma_f16_ts(tmem, tmem + A_COL + k * 8, make_desc_kmajor_sw128(b_addr + k * 32), idesc, k > 0 ? 1u : 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  int bad = 0;
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + lane_base + half * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) {
      const int n = half * 32 + j;
      int ref = 0;
      for (int k = 0; k < 64; ++k) ref += a_val(m, k) * b_val(n, k);
      if (__uint_as_float(r[j]) != float(ref)) ++bad;
      if (m == 5 && n == 9) { sample[0] = __uint_as_float(r[j]); sample[1] = float(ref); }
    }
  }
  if (bad) atomicAdd(mismatches, bad);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  int* d;
  float* s;
  cudaMalloc(&d, 4);
  cudaMalloc(&s, 8);
  cudaMemset(d, 0, 4);
  cudaMemset(s, 0, 8);
  const int smem = 64 * 128 + 1024;
  ts_check_kernel<<<1, 128, smem>>>(d, s);
  cudaError_t e = cudaDeviceSynchronize();
  int bad = -1;
  float hs[2] = {0, 0};
  cudaMemcpy(&bad, d, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hs, s, 8, cudaMemcpyDeviceToHost);
  printf("TS-mode A-in-TMEM check: %d mismatches of %d (%s); D[5][9] = %.1f, reference %.1f\n", bad, 128 * 64,
         cudaGetErrorString(e), hs[0], hs[1]);
  return bad != 0;
}
