// Implicit-GEMM on tcgen05 tensor cores for every dense contraction of the denoising path:
//   Linear / 1x1 conv (1 tap), conv3x3 (9 taps over H,W), temporal conv (3,1,1) (3 taps over T),
//   stride-2 conv (9 taps over a space-to-depth view), batched Q.K^T / P.V for the VAE attention.
//
//   D[pixels, N] = act( sum_taps A_tap[pixels, Kc] . B[N, tap*Kc : (tap+1)*Kc]^T + bias + bias2[sample] + residual ) * s
//
// * A (activations, channels-last) is never im2col'd: every tap is a TMA box load of the same 5-D tensor map at a
//   shifted pixel coordinate; out-of-bounds pixels are zero-filled by the TMA unit (= conv zero padding).
// * B (weights, [N, K] K-contiguous) is TMA-loaded; both land in shared memory in the 128B-swizzled K-major layout
//   that tcgen05.mma consumes directly.  One elected thread issues tcgen05.mma (M=128, N=BN, K=16), accumulators
//   live in TMEM (double buffered so the epilogue of tile i overlaps the main loop of tile i+1).
// * Persistent CTAs (one per SM) walk the tile list n-fastest so that the CTAs running concurrently share the same
//   A rows (L2 hits) while the weights stay L2-resident.
// * Epilogue warps: tcgen05.ld -> bias / per-sample bias (time embedding) / residual / activation / GEGLU gate ->
//   16-bit pack -> swizzled smem staging -> TMA store (hardware clips partial tiles).  A direct-store variant handles
//   N that is not a multiple of 8 (e.g. conv_out, N=4) and fp32 outputs.
//
// Replaces, for the reference path: cuDNN conv2d/conv3d and cuBLAS linear calls made by
// diffusers ResnetBlock2D / TemporalConvLayer / Transformer2DModel / TransformerTemporalModel / AutoencoderKL
// (models/unet_3d_blocks.py:262-306,425-467,564-583,660-701,794-813).
#include "common.cuh"
#include "igemm.h"

namespace aab {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;       // 16 KiB
constexpr int OUT_CHUNK_BYTES = BM * 32 * 2;     // one 32-column output chunk (64-byte rows, SWIZZLE_64B), 8 KiB
constexpr int NUM_EPI_GROUPS = 4;                // epilogue warpgroups (4 warps each): latency-bound chains, so more in parallel
constexpr int STORE_WARP = 2 + 4 * NUM_EPI_GROUPS;
constexpr int NUM_THREADS = 32 * (STORE_WARP + 1);  // warp0 TMA, warp1 MMA, warps 2..17 epilogue, warp 18 store

// DEEP = epilogue-heavy launches (few K iterations per tile): one pipeline stage less, staging ring twice as deep
// (residual prefetch distance / store slack 7 chunks instead of 3).
template <int BN, bool DEEP>
struct Cfg {
  static constexpr int STAGES_BASE = (BN == 256) ? 4 : (BN == 128) ? 6 : 8;
  static constexpr int STAGES = DEEP ? ((BN == 64 || BN == 32) ? STAGES_BASE - 2 : STAGES_BASE - 1) : STAGES_BASE;
  static constexpr int NBUF = DEEP ? 8 : 4;
  static constexpr int NBUF_LOG2 = DEEP ? 3 : 2;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;   // two accumulator stages
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + NBUF * OUT_CHUNK_BYTES +
                                    1024 /*align*/ + 512 /*barriers*/;
};

// profiling aid: cycles spent inside an mbarrier wait, accumulated per warp role when IgemmParams::dbg != NULL
struct WaitTimer {
  unsigned long long acc = 0;
  bool on;
  __device__ explicit WaitTimer(const void* dbg) : on(dbg != nullptr) {}
  __device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
    if (on) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      acc += static_cast<unsigned long long>(clock64() - t0);
    } else {
      mbar_wait(bar, parity);
    }
  }
  __device__ __forceinline__ void flush(unsigned long long* dbg, int slot) {
    if (on) atomicAdd(dbg + slot, acc);
  }
};

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == AAB_ACT_SILU) return silu_f(x);
  if (act == AAB_ACT_GELU) return gelu_erf_f(x);
  return x;
}

// EPI: 0 = 16-bit output through swizzled smem staging + TMA store (n_out % 32 == 0);
//      1 = same with the GEGLU gate (B tile = [values | gates], out = value * gelu(gate));
//      2 = generic direct-to-global store (any n_out, fp32 or 16-bit output, masked).
// The epilogue is written as compact loops (no full unrolling): its instruction footprint is executed once per tile by
// four warps, and a bloated epilogue thrashes the instruction cache when K is small.
template <int BN, int EPI, bool DEEP>
__global__ void __launch_bounds__(NUM_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
             const __grid_constant__ CUtensorMap tmR, const IgemmParams p) {
  using C = Cfg<BN, DEEP>;
  constexpr int STAGES = C::STAGES;
  constexpr int NUM_OUT_BUFS = C::NBUF;
  constexpr int NB_LOG2 = C::NBUF_LOG2;
  constexpr bool GEGLU = (EPI == 1);
  constexpr bool DIRECT = (EPI == 2);
  constexpr int OUT_BN = GEGLU ? BN / 2 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smA + STAGES * A_STAGE_BYTES;
  uint8_t* smO = smB + STAGES * C::B_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smO + NUM_OUT_BUFS * OUT_CHUNK_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint64_t* res_bar = bars + 2 * STAGES + 4;     // [NUM_OUT_BUFS] residual chunk landed in staging buffer
  uint64_t* ready_bar = res_bar + NUM_OUT_BUFS;  // [NUM_OUT_BUFS] output chunk written by 128 epilogue threads
  uint64_t* bfree_bar = ready_bar + NUM_OUT_BUFS;  // [NUM_OUT_BUFS] staging buffer drained by its TMA store
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfree_bar + NUM_OUT_BUFS);

  const int warp = threadIdx.x >> 5;
  const bool bf16 = (p.flags & AAB_F_BF16) != 0;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int kb_per_tap = p.kb_per_tap;
  const int k_iters = p.num_taps * kb_per_tap;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmA2);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmD);
    prefetch_tmap(&tmR);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull_bar[i], 1);
        mbar_init(&tempty_bar[i], 128 * NUM_EPI_GROUPS);
      }
      for (int i = 0; i < NUM_OUT_BUFS; ++i) {
        mbar_init(&res_bar[i], 1);
        mbar_init(&ready_bar[i], 128);
        mbar_init(&bfree_bar[i], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      WaitTimer w_empty(p.dbg);
      const long long t_start = clock64();
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = tile % p.num_n_tiles;
        int mt = tile / p.num_n_tiles;
        int cb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          cb[i] = (mt % p.tiles[i]) * p.box[i];
          mt /= p.tiles[i];
        }
        const int bbatch = (p.b_batch_dim >= 0) ? cb[p.b_batch_dim] : 0;
        const int n0 = nt * OUT_BN;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int o0 = p.tap_off[tap][0], o1 = p.tap_off[tap][1], o2 = p.tap_off[tap][2], o3 = p.tap_off[tap][3],
                    o4 = p.tap_off[tap][4];
          for (int kb = 0; kb < kb_per_tap; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            w_empty.wait(&empty_bar[s], ph ^ 1);
            mbar_arrive_expect_tx(&full_bar[s], A_STAGE_BYTES + C::B_STAGE_BYTES);
            const int kc = kb * BK;
            if (kc < p.Kc1)
              tma_load_5d(smA + s * A_STAGE_BYTES, &tmA, &full_bar[s], kc + o0, cb[0] + o1, cb[1] + o2, cb[2] + o3,
                          cb[3] + o4);
            else
              tma_load_5d(smA + s * A_STAGE_BYTES, &tmA2, &full_bar[s], kc - p.Kc1 + o0, cb[0] + o1, cb[1] + o2,
                          cb[2] + o3, cb[3] + o4);
            const int kg = tap * p.Kc + kc;
            tma_load_3d(smB + s * C::B_STAGE_BYTES, &tmB, &full_bar[s], kg, n0, bbatch);
            if (GEGLU)
              tma_load_3d(smB + s * C::B_STAGE_BYTES + (BN / 2) * 128, &tmB, &full_bar[s], kg, p.N / 2 + n0, bbatch);
          }
        }
      }
      w_empty.flush(p.dbg, 0);                                  // slot 0: producer waiting for a free stage
      if (p.dbg) atomicAdd(p.dbg + 15, static_cast<unsigned long long>(clock64() - t_start));   // slot 15: producer lifetime
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    WaitTimer w_full(p.dbg), w_tempty(p.dbg);
    uint32_t it = 0;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
      const uint32_t as = tl & 1;
      const uint32_t aph = (tl >> 1) & 1;
      // ragged last N tile: issue the MMA only over the valid columns (multiple of 16); the TMA box of B is zero-filled
      // beyond N, so no extra traffic either.  N = 320 -> tiles of 256 + 64 columns instead of 3 x 128.
      int n_mma = BN;
      if (!GEGLU) {
        const int nvalid = p.N - (tile % p.num_n_tiles) * BN;
        if (nvalid < BN) n_mma = (nvalid + 15) & ~15;
      }
      const uint32_t idesc = make_idesc_f16(bf16 ? 1 : 0, BM, n_mma, 0, 0);
      w_tempty.wait(&tempty_bar[as], aph ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int ki = 0; ki < k_iters; ++ki, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        w_full.wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smA + s * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smB + s * C::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_desc_kmajor_sw128(a_addr + k * 32);
            const uint64_t db = make_desc_kmajor_sw128(b_addr + k * 32);
            umma_f16_ss(tmem_d, da, db, idesc, (ki > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (ki == k_iters - 1) umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
      }
    }
    if (lane_id() == 0) {
      w_full.flush(p.dbg, 1);                                   // slot 1: MMA waiting for TMA data
      w_tempty.flush(p.dbg, 2);                                 // slot 2: MMA waiting for a drained accumulator
    }
  } else if (warp == STORE_WARP) {
    // ===================================================== store warp: drains the staging ring with TMA stores and
    // prefetches residual chunks (TMA loads) into freed staging buffers; epilogue warps never wait on a store.
    constexpr int CPT = OUT_BN / 32;
    if (!DIRECT && elect_one()) {
      const bool has_res = (p.residual != nullptr);
      auto issue_res_load = [&](uint32_t gg) {
        const uint32_t tseq = gg / CPT;
        const long tile2 = static_cast<long>(blockIdx.x) + static_cast<long>(tseq) * gridDim.x;
        if (tile2 >= num_tiles) return;
        const int nt2 = static_cast<int>(tile2 % p.num_n_tiles);
        int mt2 = static_cast<int>(tile2 / p.num_n_tiles);
        int c2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          c2[i] = (mt2 % p.tiles[i]) * p.box[i];
          mt2 /= p.tiles[i];
        }
        const int col2 = nt2 * OUT_BN + static_cast<int>(gg % CPT) * 32;
        if (col2 >= p.n_out) return;
        const uint32_t b2 = gg & (NUM_OUT_BUFS - 1);
        mbar_arrive_expect_tx(&res_bar[b2], OUT_CHUNK_BYTES);
        tma_load_5d(smO + b2 * OUT_CHUNK_BYTES, &tmR, &res_bar[b2], col2, c2[0], c2[1], c2[2], c2[3]);
      };
      // residual rows are pulled into L2 two tiles ahead (HBM latency is longer than the ring can cover),
      // the ring loads below then hit L2
      auto prefetch_res_tile = [&](long tile2) {
        if (tile2 >= num_tiles) return;
        const int nt2 = static_cast<int>(tile2 % p.num_n_tiles);
        int mt2 = static_cast<int>(tile2 / p.num_n_tiles);
        int c2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          c2[i] = (mt2 % p.tiles[i]) * p.box[i];
          mt2 /= p.tiles[i];
        }
        for (int ci = 0; ci < CPT; ++ci) {
          const int col2 = nt2 * OUT_BN + ci * 32;
          if (col2 < p.n_out) tma_prefetch_l2_5d(&tmR, col2, c2[0], c2[1], c2[2], c2[3]);
        }
      };
      if (has_res) {
        prefetch_res_tile(blockIdx.x);
        prefetch_res_tile(static_cast<long>(blockIdx.x) + gridDim.x);
        for (uint32_t g0 = 0; g0 < NUM_OUT_BUFS; ++g0) issue_res_load(g0);
      }
      WaitTimer w_ready(p.dbg);
      unsigned long long drain = 0;
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = tile % p.num_n_tiles;
        int mt = tile / p.num_n_tiles;
        int cb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          cb[i] = (mt % p.tiles[i]) * p.box[i];
          mt /= p.tiles[i];
        }
        if (has_res) prefetch_res_tile(static_cast<long>(tile) + 2L * gridDim.x);
        for (int ci = 0; ci < CPT; ++ci, ++g) {
          const uint32_t buf = g & (NUM_OUT_BUFS - 1);
          const int col = nt * OUT_BN + ci * 32;
          w_ready.wait(&ready_bar[buf], (g >> NB_LOG2) & 1);
          if (col < p.n_out) tma_store_5d(&tmD, smO + buf * OUT_CHUNK_BYTES, col, cb[0], cb[1], cb[2], cb[3]);
          tma_store_commit();                 // (an empty group for skipped chunks keeps the ring count exact)
          const long long td = p.dbg ? clock64() : 0;
          tma_store_wait_read<1>();           // every store but the newest has drained its staging buffer
          if (p.dbg) drain += static_cast<unsigned long long>(clock64() - td);
          if (g >= 1) {
            mbar_arrive(&bfree_bar[(g - 1) & (NUM_OUT_BUFS - 1)]);
            if (has_res) issue_res_load(g + NUM_OUT_BUFS - 1);
          }
        }
      }
      tma_store_wait_all<0>();
      w_ready.flush(p.dbg, 3);                                  // slot 3: store warp waiting for a written chunk
      if (p.dbg) atomicAdd(p.dbg + 4, drain);                   // slot 4: store warp waiting for TMA stores to drain
    }
  } else {
    // ===================================================== epilogue: warps 2..17 = four groups of four warps (TMEM lane
    // quarter = warp % 4).  Group eg takes the 32-column chunks with (chunk index % 4 == eg) of every tile and writes
    // them (16-bit, SWIZZLE_64B) into the staging ring; when a residual is added its chunk has been TMA-prefetched
    // into the same staging buffer, so each thread finds its residual piece exactly where it will write its output.
    const int q = warp & 3;
    const int eg = (warp - 2) >> 2;
    const int row = q * 32 + lane_id();           // row inside the 128-row tile == TMEM lane
    constexpr int CPT = OUT_BN / 32;              // chunks per tile
    const bool has_res = !DIRECT && (p.residual != nullptr);
    uint32_t tl = 0;
    uint32_t gbase = 0;                           // global chunk index of the first chunk of the current tile
    uint32_t res_phase = 0;                       // per staging buffer phase bits of res_bar
    WaitTimer w_tfull(p.dbg), w_bfree(p.dbg), w_res(p.dbg);
    unsigned long long t_tmem = 0, t_fence = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl, gbase += CPT) {
      const uint32_t as = tl & 1;
      const uint32_t aph = (tl >> 1) & 1;
      const int nt = tile % p.num_n_tiles;
      int mt = tile / p.num_n_tiles;
      int cb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cb[i] = (mt % p.tiles[i]) * p.box[i];
        mt /= p.tiles[i];
      }
      long grow = 0;
      bool rvalid = true;
      {
        int rr = row;
        int g[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          g[i] = cb[i] + rr % p.box[i];
          rr /= p.box[i];
          rvalid = rvalid && (g[i] < p.dimD[i]);
        }
        grow = ((static_cast<long>(g[3]) * p.dimD[2] + g[2]) * p.dimD[1] + g[1]) * p.dimD[0] + g[0];
      }
      const int n0 = nt * OUT_BN;
      const float* bias2row =
          (p.bias2 != nullptr && rvalid) ? p.bias2 + (grow / p.rows_per_bias2) * static_cast<long>(p.ld_bias2) : nullptr;

      w_tfull.wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);

#pragma unroll 1
      for (int ci = eg; ci < CPT; ci += NUM_EPI_GROUPS) {
        const int cc = ci * 32;
        const int col = n0 + cc;                  // global output column of this 32-wide chunk
        const uint32_t gch = gbase + ci;
        const uint32_t buf = gch & (NUM_OUT_BUFS - 1);
        uint8_t* stage_buf = smO + buf * OUT_CHUNK_BYTES;
        if (!DIRECT) w_bfree.wait(&bfree_bar[buf], ((gch >> NB_LOG2) & 1) ^ 1);
        if (col < p.n_out) {                      // warp-uniform
          // (prefetching the next chunk's accumulator columns before post-processing this one was measured: it costs
          //  32-64 registers and made the GEGLU epilogue 25 % slower; tcgen05.wait::ld is ~0 % of the epilogue time)
          float v[32];
          float gt[GEGLU ? 32 : 1];
          {
            uint32_t r[32];
            tmem_ld_32x32(tmem_acc + cc, r);
            if (GEGLU) {
              uint32_t r2[32];
              tmem_ld_32x32(tmem_acc + BN / 2 + cc, r2);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) gt[j] = __uint_as_float(r2[j]);
            } else {
              tmem_ld_wait();
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          }
          if (!DIRECT) {
            // ---------------- fast path: n_out % 32 == 0, everything vectorised
            if (p.bias != nullptr) {
              const float4* bp = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 b = __ldg(bp + j4);
                v[j4 * 4 + 0] += b.x; v[j4 * 4 + 1] += b.y; v[j4 * 4 + 2] += b.z; v[j4 * 4 + 3] += b.w;
              }
            }
            if (GEGLU) {
              const float4* gp = reinterpret_cast<const float4*>(p.bias + p.N / 2 + col);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias != nullptr) b = __ldg(gp + j4);
                v[j4 * 4 + 0] *= gelu_fast_f(gt[j4 * 4 + 0] + b.x);
                v[j4 * 4 + 1] *= gelu_fast_f(gt[j4 * 4 + 1] + b.y);
                v[j4 * 4 + 2] *= gelu_fast_f(gt[j4 * 4 + 2] + b.z);
                v[j4 * 4 + 3] *= gelu_fast_f(gt[j4 * 4 + 3] + b.w);
              }
            }
            if (bias2row != nullptr) {
              const float4* bp = reinterpret_cast<const float4*>(bias2row + col);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 b = __ldg(bp + j4);
                v[j4 * 4 + 0] += b.x; v[j4 * 4 + 1] += b.y; v[j4 * 4 + 2] += b.z; v[j4 * 4 + 3] += b.w;
              }
            }
            // this thread's 64-byte row of the staging tile; 16-byte piece c lives at ((c ^ ((row >> 1) & 3)) << 4)
            uint8_t* rowp = stage_buf + row * 64;
            const int sw = (row >> 1) & 3;
            if (has_res) {
              w_res.wait(&res_bar[buf], (res_phase >> buf) & 1);
              res_phase ^= (1u << buf);
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const uint4 u = *reinterpret_cast<const uint4*>(rowp + ((j4 ^ sw) << 4));
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = unpack2(w[e], bf16);
                  v[j4 * 8 + e * 2] += f.x;
                  v[j4 * 8 + e * 2 + 1] += f.y;
                }
              }
            }
            if (p.act == AAB_ACT_SILU) {          // (GELU as a plain activation takes the generic path)
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
            }
            if (p.out_scale != 1.0f) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
            }
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              uint4 u;
              u.x = pack2(v[j4 * 8 + 0], v[j4 * 8 + 1], bf16);
              u.y = pack2(v[j4 * 8 + 2], v[j4 * 8 + 3], bf16);
              u.z = pack2(v[j4 * 8 + 4], v[j4 * 8 + 5], bf16);
              u.w = pack2(v[j4 * 8 + 6], v[j4 * 8 + 7], bf16);
              *reinterpret_cast<uint4*>(rowp + ((j4 ^ sw) << 4)) = u;
            }
          } else {
            // ---------------- generic path: masked, any n_out, fp32 or 16-bit output, straight to global memory
            const uint8_t* resrow = (p.residual != nullptr && rvalid)
                                        ? reinterpret_cast<const uint8_t*>(p.residual) + grow * p.ld_res * 2
                                        : nullptr;
            const int nv = (p.n_out - col < 32) ? (p.n_out - col) : 32;
#pragma unroll 1
            for (int j = 0; j < 32; ++j) {
              // v[] is indexed dynamically here on purpose (compact code); it lives in local memory on this path
              if (j < nv) {
                float x = v[j];
                if (p.bias != nullptr) x += __ldg(p.bias + col + j);
                if (bias2row != nullptr) x += __ldg(bias2row + col + j);
                if (resrow != nullptr) x += load_elem(resrow, col + j, bf16);
                x = apply_act(x, p.act) * p.out_scale;
                if (rvalid) {
                  if (p.flags & AAB_F_OUT_F32) reinterpret_cast<float*>(p.out)[grow * p.ld_out + col + j] = x;
                  else store_elem(p.out, grow * p.ld_out + col + j, x, bf16);
                }
              }
            }
          }
        }
        if (!DIRECT) {
          const long long tf0 = p.dbg ? clock64() : 0;
          fence_proxy_async_smem();
          mbar_arrive(&ready_bar[buf]);           // non-blocking hand-off to the store warp
          if (p.dbg) t_fence += static_cast<unsigned long long>(clock64() - tf0);
        }
      }
      // all TMEM reads of this accumulator stage are complete -> hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
    }
    if ((threadIdx.x == 64 || threadIdx.x == 64 + 128) && eg < 2) {   // one thread of epilogue groups 0 and 1
      w_tfull.flush(p.dbg, 5 + 3 * eg);                          // slots 5/8: epilogue waiting for the accumulator
      w_bfree.flush(p.dbg, 6 + 3 * eg);                          // slots 6/9: waiting for a free staging buffer
      w_res.flush(p.dbg, 7 + 3 * eg);                            // slots 7/10: waiting for the residual chunk
      if (p.dbg && eg == 0) {
        atomicAdd(p.dbg + 11, t_tmem);                           // slot 11: tcgen05.wait::ld
        atomicAdd(p.dbg + 12, t_fence);                          // slot 12: fence.proxy.async + arrive
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

// rank-`rank` 16-bit tensor map, dims/strides innermost first (strides in elements, strides[0] must be 1)
int make_tmap_16(CUtensorMap* out, const void* base, int rank, const long* dims, const long* strides, const int* box,
                 int is_bf16, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return AAB_ERR_DRIVER;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = static_cast<cuuint64_t>(dims[i]);
    bx[i] = static_cast<cuuint32_t>(box[i]);
    es[i] = 1;
    if (i > 0) gstr[i - 1] = static_cast<cuuint64_t>(strides[i]) * 2;
  }
  CUresult r = enc(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? AAB_OK : AAB_ERR_DRIVER;
}

static int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int EPI, bool DEEP>
static int launch_cfg(const CUtensorMap& a, const CUtensorMap& a2, const CUtensorMap& b, const CUtensorMap& d,
                      const CUtensorMap& r, const IgemmParams& p, int max_ctas, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_kernel<BN, EPI, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg<BN, DEEP>::SMEM_BYTES);
    if (e != cudaSuccess) return AAB_ERR_CUDA;
    attr_set = true;
  }
  int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = tiles < num_sms() ? tiles : num_sms();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  igemm_kernel<BN, EPI, DEEP><<<grid, NUM_THREADS, Cfg<BN, DEEP>::SMEM_BYTES, stream>>>(a, a2, b, d, r, p);
  return cudaGetLastError() == cudaSuccess ? AAB_OK : AAB_ERR_CUDA;
}

template <int BN, int EPI>
static int launch_bn(const CUtensorMap& a, const CUtensorMap& a2, const CUtensorMap& b, const CUtensorMap& d,
                     const CUtensorMap& r, const IgemmParams& p, int max_ctas, cudaStream_t stream) {
  // The DEEP variant (8-buffer staging ring, one pipeline stage less) was measured on B200 and is NOT faster: with
  // K <= 768 the MMA warp then waits longer for TMA data (3 instead of 4 stages) than the epilogue gains from the deeper
  // ring (profiles/r01_igemm_roles.md).  It stays selectable for experiments through AAB_F_DEEP_RING.
  if (EPI != 2 && (p.flags & AAB_F_DEEP_RING)) return launch_cfg<BN, EPI, true>(a, a2, b, d, r, p, max_ctas, stream);
  return launch_cfg<BN, EPI, false>(a, a2, b, d, r, p, max_ctas, stream);
}

}  // namespace aab

using namespace aab;

extern "C" int aab_igemm(const AabIgemmDesc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!d || !d->a || !d->b || !d->out) return AAB_ERR_ARG;
  const int is_bf16 = (d->flags & AAB_F_BF16) ? 1 : 0;
  const bool geglu = (d->flags & AAB_F_GEGLU) != 0;
  int bn = d->block_n;
  if (bn != 32 && bn != 64 && bn != 128 && bn != 256) return AAB_ERR_ARG;
  bool direct = (d->flags & AAB_F_DIRECT) != 0;
  const int n_out = geglu ? d->n / 2 : d->n;
  if (bn == 32 || (d->ld_out % 8) != 0 || (n_out % 32) != 0 || (d->flags & AAB_F_OUT_F32)) direct = true;
  if (geglu && (bn < 128 || direct)) return AAB_ERR_ARG;   // output tile must cover whole 64-column store boxes
  if (d->residual && (d->ld_res % 8) != 0 && !direct) direct = true;
  if (d->act == AAB_ACT_GELU && !geglu) direct = true;
  if (d->num_taps < 1 || d->num_taps > AAB_MAX_TAPS) return AAB_ERR_ARG;
  if (d->kc % 8 != 0) return AAB_ERR_ARG;
  if (d->a2 && (d->kc1 % 64 != 0)) return AAB_ERR_ARG;

  IgemmParams p;
  memset(&p, 0, sizeof(p));
  long box_prod = 1;
  p.num_m_tiles = 1;
  for (int i = 0; i < 4; ++i) {
    p.dimD[i] = d->dim_d[i];
    p.box[i] = d->box[i];
    if (p.box[i] < 1 || p.dimD[i] < 1) return AAB_ERR_ARG;
    p.tiles[i] = (p.dimD[i] + p.box[i] - 1) / p.box[i];
    p.num_m_tiles *= p.tiles[i];
    box_prod *= p.box[i];
  }
  if (box_prod != BM) return AAB_ERR_ARG;
  const int out_bn = geglu ? bn / 2 : bn;
  p.num_n_tiles = (n_out + out_bn - 1) / out_bn;
  p.N = d->n;
  p.n_out = n_out;
  p.Kc = d->kc;
  p.Kc1 = d->a2 ? d->kc1 : (1 << 30);
  p.num_taps = d->num_taps;
  p.kb_per_tap = (d->kc + BK - 1) / BK;
  for (int t = 0; t < d->num_taps; ++t)
    for (int j = 0; j < 5; ++j) p.tap_off[t][j] = d->tap_off[t][j];
  p.b_batch_dim = d->b_batch_dim;
  p.bias = d->bias;
  p.bias2 = d->bias2;
  p.rows_per_bias2 = d->rows_per_bias2 > 0 ? d->rows_per_bias2 : 1;
  p.ld_bias2 = d->ld_bias2;
  p.residual = d->residual;
  p.ld_res = d->ld_res;
  p.out = d->out;
  p.ld_out = d->ld_out;
  p.out_scale = d->out_scale;
  p.act = d->act;
  p.flags = d->flags | (direct ? AAB_F_DIRECT : 0);
  p.dbg = d->debug_cycles;

  CUtensorMap tmA, tmA2, tmB, tmD, tmR;
  {
    int box[5] = {BK, d->box[0], d->box[1], d->box[2], d->box[3]};
    int r = make_tmap_16(&tmA, d->a, 5, d->a_dims, d->a_strides, box, is_bf16, 128);
    if (r) return r;
    if (d->a2) {
      r = make_tmap_16(&tmA2, d->a2, 5, d->a2_dims, d->a2_strides, box, is_bf16, 128);
      if (r) return r;
    } else {
      tmA2 = tmA;
    }
  }
  {
    long dims[3] = {static_cast<long>(d->num_taps) * d->kc, d->n, d->b_batch > 0 ? d->b_batch : 1};
    long strides[3] = {1, d->ld_b, d->b_batch_stride > 0 ? d->b_batch_stride : static_cast<long>(d->n) * d->ld_b};
    int box[3] = {BK, geglu ? bn / 2 : bn, 1};
    int r = make_tmap_16(&tmB, d->b, 3, dims, strides, box, is_bf16, 128);
    if (r) return r;
  }
  if (!direct) {
    // D (and the residual) viewed with the same pixel decomposition as the tile grid: [n_out, dimD0..3]; rows are
    // contiguous in pixel-linear order with a row stride of ld_out (ld_res) elements.  32-column boxes, 64B swizzle.
    long dims[5] = {n_out, d->dim_d[0], d->dim_d[1], d->dim_d[2], d->dim_d[3]};
    long strides[5];
    strides[0] = 1;
    strides[1] = d->ld_out;
    for (int i = 2; i < 5; ++i) strides[i] = strides[i - 1] * d->dim_d[i - 2];
    int box[5] = {32, d->box[0], d->box[1], d->box[2], d->box[3]};
    int r = make_tmap_16(&tmD, d->out, 5, dims, strides, box, is_bf16, 64);
    if (r) return r;
    if (d->residual) {
      strides[1] = d->ld_res;
      for (int i = 2; i < 5; ++i) strides[i] = strides[i - 1] * d->dim_d[i - 2];
      r = make_tmap_16(&tmR, d->residual, 5, dims, strides, box, is_bf16, 64);
      if (r) return r;
    } else {
      tmR = tmD;
    }
  } else {
    tmD = tmB;
    tmR = tmB;
  }
  if (direct) {
    switch (bn) {
      case 32: return launch_bn<32, 2>(tmA, tmA2, tmB, tmD, tmR, p, d->max_ctas, stream);
      case 64: return launch_bn<64, 2>(tmA, tmA2, tmB, tmD, tmR, p, d->max_ctas, stream);
      case 128: return launch_bn<128, 2>(tmA, tmA2, tmB, tmD, tmR, p, d->max_ctas, stream);
      default: return launch_bn<256, 2>(tmA, tmA2, tmB, tmD, tmR, p, d->max_ctas, stream);
    }
  }
  if (geglu) {
    if (bn == 128) return launch_bn<128, 1>(tmA, tmA2, tmB, tmD, tmR, p, d->max_ctas, stream);
    return launch_bn<256, 1>(tmA, tmA2, tmB, tmD, tmR, p, d->max_ctas, stream);
  }
  switch (bn) {
    case 64: return launch_bn<64, 0>(tmA, tmA2, tmB, tmD, tmR, p, d->max_ctas, stream);
    case 128: return launch_bn<128, 0>(tmA, tmA2, tmB, tmD, tmR, p, d->max_ctas, stream);
    default: return launch_bn<256, 0>(tmA, tmA2, tmB, tmD, tmR, p, d->max_ctas, stream);
  }
}

extern "C" int aab_num_sms(void) { return aab::num_sms(); }
