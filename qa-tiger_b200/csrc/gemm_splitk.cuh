// Split-K tcgen05 GEMM for the per-question contractions that come AFTER the router (M = batch rows, 256 .. 512):
//   C[M,N] = act(A[M,K] W[N,K]^T + bias) + res,  A / W as bf16 (hi, lo) plane pairs, D += Alo*Whi + Ahi*Wlo + Ahi*Whi.
// With so few rows the persistent kernel of gemm_tc.cuh puts 8 - 32 CTAs on the problem and every one of them pulls its whole
// K extent (up to 1.9 MB) through ONE SM's L2 port: measured 22 us for [512,1856]x[512,1856]^T, no better than the CUDA-core
// kernel, which is shared-memory bound at its 4 x 4 register tile (35 us).  Here the k-blocks of a 128 x 128 tile are cut into S
// slices (S <= 8) so that tiles * S CTAs cover the SMs; the S CTAs of a tile form a thread-block cluster: each parks its fp32
// partial tile in its own shared memory (the operand ring is free by then), and after a cluster barrier rank r sums rows
// [128 r / S, 128 (r + 1) / S) of the S partials over distributed shared memory in slice order (deterministic), applies bias /
// activation / residual and stores whichever planes the consumer reads.  No partial tile reaches HBM, one launch per GEMM.
//   warp 0      TMA producer (4 boxes per k-block: A hi, A lo, W hi, W lo; 128B swizzle; 3-stage ring)
//   warp 1      TMEM allocator (128 columns) + tcgen05.mma issuer (12 MMAs per k-block)
//   warps 2..5  accumulator -> fp32 partial tile in shared memory (one TMEM lane quarter each), then the cluster reduction
#pragma once
#include "gemm_tc.cuh"

namespace qat {
namespace tc {

constexpr int SK_STAGES = 3;
constexpr int SK_STAGE_BYTES = 4 * 16384;                 // A hi | A lo | W hi | W lo, [128 x 64] bf16 each
constexpr int SK_SMEM = SK_STAGES * SK_STAGE_BYTES + 128 + 1024;

constexpr int SK_LDP = 132;                               // floats per row of the parked partial tile (128 + 4: 16-byte aligned rows, 4 banks apart)

struct SplitKP {
  int tiles_m, tiles_n, S, nkb;
  int M, N;
  Epi epi;
  Act C;
};

__device__ __forceinline__ float4 ld_cluster_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

__global__ void __launch_bounds__(192, 1)
gemm_splitk_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo, const __grid_constant__ CUtensorMap mWhi,
                   const __grid_constant__ CUtensorMap mWlo, SplitKP p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SK_STAGES * SK_STAGE_BYTES);
  uint64_t* empty = full + SK_STAGES;
  uint64_t* tfull = empty + SK_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / p.S, slice = blockIdx.x % p.S;
  const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
  const int kb0 = (int)((long long)p.nkb * slice / p.S), kb1 = (int)((long long)p.nkb * (slice + 1) / p.S);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mAhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mWhi)) : "memory");
    for (int s = 0; s < SK_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tacc = *tmem_slot;

  if (warp == 0) {
    for (int kb = kb0, it = 0; kb < kb1; ++kb, ++it) {
      const int s = it % SK_STAGES;
      mbar_wait(&empty[s], ((it / SK_STAGES) & 1) ^ 1);
      unsigned char* st = smem + s * SK_STAGE_BYTES;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[s], SK_STAGE_BYTES);
        tma_load_2d(st, &mAhi, &full[s], kb * 64, tm * 128);
        tma_load_2d(st + 16384, &mAlo, &full[s], kb * 64, tm * 128);
        tma_load_2d(st + 32768, &mWhi, &full[s], kb * 64, tn * 128);
        tma_load_2d(st + 49152, &mWlo, &full[s], kb * 64, tn * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc(1, 128, 128);                   // bf16 operands
    for (int kb = kb0, it = 0; kb < kb1; ++kb, ++it) {
      const int s = it % SK_STAGES;
      mbar_wait(&full[s], (it / SK_STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_u32(smem + s * SK_STAGE_BYTES);
      const uint64_t dAh = umma_desc(a0), dAl = umma_desc(a0 + 16384), dWh = umma_desc(a0 + 32768), dWl = umma_desc(a0 + 49152);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adv = uint64_t(k * 2);
          umma<KIND_BF16X3>(tacc, dAl + adv, dWh + adv, idesc, (it | k) != 0);    // lo * hi
          umma<KIND_BF16X3>(tacc, dAh + adv, dWl + adv, idesc, 1);                // hi * lo
          umma<KIND_BF16X3>(tacc, dAh + adv, dWh + adv, idesc, 1);                // hi * hi
        }
        umma_commit(&empty[s]);
        if (kb == kb1 - 1) umma_commit(tfull);
      }
      __syncwarp();
    }
  } else {
    // every MMA has completed (tfull), so the operand ring is free: the partial tile is parked at its start
    const int lg = warp & 3, row = lg * 32 + lane;
    mbar_wait(tfull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* dst = reinterpret_cast<float*>(smem) + row * SK_LDP;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(tacc + (uint32_t(lg * 32) << 16) + uint32_t(c * 32), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(dst + c * 32 + 4 * q) = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                       __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  cluster_sync_all();                                     // every slice's partial tile is in its CTA's shared memory
  if (warp >= 2) {
    const int r0 = 128 * slice / p.S, r1 = 128 * (slice + 1) / p.S;     // the rows this rank combines
    const int t = threadIdx.x - 64, c4 = (t & 31) * 4;
    const uint32_t base = smem_u32(smem);
    for (int row = r0 + (t >> 5); row < r1; row += 4) {
      const long long gr = (long long)tm * 128 + row;
      if (gr >= p.M) break;
      const uint32_t off = base + uint32_t(row * SK_LDP + c4) * 4u;
      float4 a = ld_cluster_f4(mapa_rank(off, 0));
      for (int sl = 1; sl < p.S; ++sl) {
        const float4 b = ld_cluster_f4(mapa_rank(off, (uint32_t)sl));
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      const int gc = tn * 128 + c4;
      if (p.epi.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.epi.bias + gc));
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      a.x = epi_act(a.x, p.epi.relu); a.y = epi_act(a.y, p.epi.relu); a.z = epi_act(a.z, p.epi.relu); a.w = epi_act(a.w, p.epi.relu);
      if (p.epi.res) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.epi.res + gr * p.epi.ldr + gc));
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      act_store4(p.C, gr, gc, a);
    }
  }
  cluster_sync_all();                                     // no CTA leaves while a peer may still read its partial tile
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tacc), "r"(128) : "memory");
  }
}

}  // namespace tc

// usable when the contraction may run on 16-bit-mantissa operands (after the router), M <= 1024, N % 128 == 0, K % 64 == 0
inline bool gemm_splitk_ok(long long M, int N, int K, const bf16* Whi, const bf16* Wlo) {
  return M >= 1 && M <= 1024 && N % 128 == 0 && K % 64 == 0 && Whi && Wlo;
}

inline int gemm_splitk_bf16x3(Ctx& c, const bf16* Ahi, const bf16* Alo, int lda, long long M, const bf16* Whi, const bf16* Wlo, int N, int K,
                              const Epi& epi, Act C) {
  QAT_REQUIRE(gemm_splitk_ok(M, N, K, Whi, Wlo), "gemm_splitk: need M <= 1024, N %% 128 == 0, K %% 64 == 0 (M=%lld N=%d K=%d)", M, N, K);
  QAT_REQUIRE(!epi.coef && !epi.res_h && C.ld % 4 == 0 && (!epi.res || epi.ldr % 4 == 0), "gemm_splitk: unsupported epilogue");
  tc::SplitKP p;
  p.tiles_m = (int)((M + 127) / 128);
  p.tiles_n = N / 128;
  p.nkb = K / 64;
  p.M = (int)M;
  p.N = N;
  p.epi = epi;
  p.C = C;
  const int tiles = p.tiles_m * p.tiles_n;
  if (c.dry) return QAT_OK;
  QAT_REQUIRE(((reinterpret_cast<uintptr_t>(Ahi) | reinterpret_cast<uintptr_t>(Alo) | reinterpret_cast<uintptr_t>(Whi) | reinterpret_cast<uintptr_t>(Wlo)) & 15) == 0 &&
                  lda % 8 == 0, "gemm_splitk: operands must be 16-byte aligned with 16-byte row pitches");
  CUtensorMap mAhi, mAlo, mWhi, mWlo;
  QAT_TRY(tc::make_map(&mAhi, Ahi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, M, K, lda, 128));
  QAT_TRY(tc::make_map(&mAlo, Alo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, M, K, lda, 128));
  QAT_TRY(tc::make_map(&mWhi, Whi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, N, K, K, 128));
  QAT_TRY(tc::make_map(&mWlo, Wlo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, N, K, K, 128));
  QAT_CUDA(cudaFuncSetAttribute(tc::gemm_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SK_SMEM));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = tc::SK_SMEM;
  cfg.stream = c.stream;
  // S = the largest cluster size whose clusters are all co-resident (a cluster of 8 one-CTA-per-SM blocks needs 8 free SMs in one
  // GPC: 16 of them do not fit a B200 and the launch would run in two waves); a device property, looked up once per size
  static std::atomic<int> max_clusters[9];
  int S = 1;
  for (int cand = 8; cand >= 2; cand >>= 1) {
    if (cand > p.nkb) continue;
    int mc = max_clusters[cand].load(std::memory_order_relaxed);
    if (mc == 0) {
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = (unsigned)cand;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      cfg.gridDim = dim3((unsigned)cand);
      cfg.attrs = qa;
      cfg.numAttrs = 1;
      QAT_CUDA(cudaOccupancyMaxActiveClusters(&mc, tc::gemm_splitk_kernel, &cfg));
      if (mc < 1) mc = -1;
      max_clusters[cand].store(mc, std::memory_order_relaxed);
    }
    if (mc >= tiles) {
      S = cand;
      break;
    }
  }
  p.S = S;
  cfg.gridDim = dim3((unsigned)(tiles * S));
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)S;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  QAT_CUDA(cudaLaunchKernelEx(&cfg, tc::gemm_splitk_kernel, mAhi, mAlo, mWhi, mWlo, p));
  return check_launch("gemm_splitk_kernel");
}

}  // namespace qat
