// tcgen05 tensor-core GEMMs for sm_100a:  C[M,N] = act(A[M,K] W[N,K]^T + bias) + res, fp32 accumulate in TMEM.
//
//   KIND_TF32   : A, W fp32 in HBM, consumed as tf32 by tcgen05.mma.kind::tf32 (1 MMA per product).
//                 Used on the patch path (67% of the FLOPs; error budget in DESIGN.md).
//   KIND_BF16X3 : A, W as bf16 (hi, lo) plane pairs, x ~= hi + lo (16 mantissa bits);
//                 D += Alo*Whi + Ahi*Wlo + Ahi*Whi  -> 3 x tcgen05.mma.kind::f16 per k-step.
//                 Used on the serial chain that feeds the router.
//   KIND_F16    : A, W as one fp16 plane each (round-to-nearest, 11 significant bits = tf32 grade at twice the
//                 tf32 MMA rate and half the bytes).  Used inside PatchSelecter, whose activations live in HBM/L2
//                 as fp16 planes (error budget in DESIGN.md).
//
// Epilogue: every output plane (fp32 / bf16 hi, lo / fp16 / fp16 hi, lo) leaves through the TMA unit: an epilogue warp packs its
// 32 x 32 sub-tile into a swizzled smem box and one lane issues cp.async.bulk.tensor (row / column tails clipped
// by the TMA unit, no strided LSU stores).
//
// Structure (one 128 x BN output tile per CTA -- or one 256 x BN tile per CTA PAIR, `CG = 2`, cta_group::2 --, 320 threads):
//   warp 0      : TMA producer  - cp.async.bulk.tensor.2d (128B swizzle) into a STAGES-deep smem ring
//   warp 1      : TMEM allocator + tcgen05.mma issuer (warp-uniform loop, one elected lane); tcgen05.commit frees smem slots
//   warps 2..9  : epilogue - tcgen05.ld (32 lanes x 32 columns) -> bias/ReLU/residual in registers -> swizzled
//                 smem box -> TMA store per plane (two warps per TMEM lane quarter, one per column half)
// Both operands are K-major; smem tiles are rows of 128 bytes (32 tf32 / 64 bf16), 8-row 1024-byte swizzle atoms.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "gemm_simt.cuh"

namespace qat {
namespace tc {

enum { KIND_TF32 = 0, KIND_BF16X3 = 1, KIND_F16 = 2 };

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory operand descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) = 1 | SBO>>4 [32,46) = 1024>>4 | version [46,48) = 1 | layout [61,64) = 2
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(2) << 61);
}
// cute::UMMA::InstrDescriptor: c_format f32 [4,6)=1, a/b format [7,10),[10,13), K-major both, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t umma_idesc(int fmt, int M, int N) {
  return (1u << 4) | (uint32_t(fmt) << 7) | (uint32_t(fmt) << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// ---- CTA-pair (cta_group::2) helpers: the two CTAs of a cluster sit on the two SMs of one TPC; the leader (cluster rank 0)
// issues one M = 256 MMA over both CTAs' shared memory (own 128 A rows + own half of the W rows each) and both TMEMs ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `saddr` (a shared::cta address of the executing CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// one elected lane of a converged warp (the same lane on every call with the full mask).  The producer and MMA loops run
// warp-uniform with only the asynchronous instruction under this predicate: their operands then stay in uniform registers
// (UTCHMMA / UTMALDG straight from URs) instead of going through a per-instruction ELECT + R2UR + retry loop, which made
// the single-thread MMA issue stream (~110 SASS instructions per k-block) the pacing item next to a busy epilogue.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the load lands in the executing CTA's shared memory; its bytes are counted on `bar_cluster` (a shared::cluster address,
// the LEADER's full barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
// bounded wait with cluster-scope acquire (the arrivals come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 24)) __trap();
  }
}

template <int KIND, int CG = 1>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  if (CG == 2) {
    if (KIND == KIND_TF32) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
          : "memory");
    } else {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
          : "memory");
    }
    return;
  }
  if (KIND == KIND_TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// pair form: one arrival on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

#ifdef QAT_PROFILE_BUILD
#define QAT_DBG(p, bit) ((p).dbg & (bit))
#else
#define QAT_DBG(p, bit) 0
#endif

// column sums over the 32 lanes of a warp: lane L holds x[0..31] (one row, 32 columns); returns, in lane j, sum_L x_L[j].
// Recursive halving: at distance o a lane keeps the half of its columns that matches its lane bit and hands the other half to
// its partner -- 31 shuffles for 32 x 32 values, fixed summation tree (bit-reproducible), no shared memory.
__device__ __forceinline__ float warp_colsum32(float (&x)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? x[i] : x[i + o];
      const float keep = up ? x[i + o] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return x[0];
}

struct GemmP {
  int M, N, K;
  int tile_rows;   // rows of A a CTA's tile advances by: 128, or (128 / T) * T with the reduce epilogue (whole samples per tile)
  Epi epi;
  Act C;
  int pair_f16;    // 1: the (hi, lo) output pair is fp16 (Act::ph / pl, through the mChi / mClo maps) instead of bf16
  int tma_res;     // 1: the fp16 residual (epi.res_h) arrives through the mR tensor map into the warp's box (fp16-only output)
  int dbg;         // QAT_PROFILE_BUILD only (per-tile cost model, tests/prof_gemm_k.py; compiled out of the product library):
                   // 1 = epilogue releases the accumulator without reading it, 2 = MMA issuer skips the MMAs, 4 = no global stores
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

template <int KIND, int BN, int STAGES, int CG = 1>
struct Cfg {
  static constexpr int BM = 128;                                          // rows per CTA (the pair tile is CG * BM rows)
  static constexpr int ROW_BYTES = 128;                                   // one swizzle row
  static constexpr int BK = KIND == KIND_TF32 ? 32 : 64;                  // elements per k-block
  static constexpr int UK = KIND == KIND_TF32 ? 8 : 16;                   // elements per MMA
  static constexpr int PLANES = KIND == KIND_BF16X3 ? 2 : 1;
  static constexpr int A_BYTES = BM * ROW_BYTES;                          // per plane
  static constexpr int BN_CTA = BN / CG;                                  // W rows this CTA stages (the pair splits them)
  static constexpr int W_BYTES = BN_CTA * ROW_BYTES;
  static constexpr int STAGE_BYTES = PLANES * (A_BYTES + W_BYTES);
  static constexpr int EPI_BYTES = 8 * 4096;                              // 8 epilogue warps x one 4 KB TMA-store box
  static constexpr int BAR_BYTES = (2 * STAGES + 4 + 8) * 8 + 16;        // + one residual-box barrier per epilogue warp
  static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;   // + alignment slack
  static constexpr int TMEM_COLS = 2 * BN;                                // two accumulators: epilogue(i) overlaps mainloop(i+1)
  static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");
};

// Persistent: grid = min(#tiles, #SMs); CTA c walks tiles c, c+grid, ... (n fastest, so co-scheduled CTAs share
// the A rows through L2).  Three decoupled loops meet only through mbarriers:
//   full/empty[STAGES]   TMA <-> MMA    (smem ring, continuous across tiles)
//   tfull/tempty[2]      MMA <-> epilogue (the two TMEM accumulators)
// EPI specialises the epilogue on the set of output planes (a smaller chunk body schedules measurably better):
//   1 = one fp16 plane, residual (if any) through the TMA unit (the PatchSelecter GEMMs); 2 = one fp32 plane;
//   3 = bf16 (hi, lo) planes; 0 = any combination behind run-time flags.
// CG = 2: CTA pairs (cluster of 2 on one TPC, launched with a cluster dimension of 2).  The pair owns a 256 x BN tile: each
// CTA stages its own 128 A rows and HALF of the W rows, the leader issues tcgen05.mma.cta_group::2 (M = 256) over both
// shared memories and both TMEMs, so every SM's operand traffic through shared memory drops by a third (the single-CTA
// f16 kind at K = 512 is bound by it: TMA writes + MMA reads + epilogue boxes exceed the 128 B/clk port).  Barriers: `full`
// lives in the leader and counts both CTAs' bytes; `empty` / `tfull` are signalled in both CTAs by a multicast commit;
// `tempty` lives in the leader and collects the epilogue warps of both CTAs.
template <int KIND, int BN, int STAGES, int EPI, int CG = 1>
__global__ void __launch_bounds__(320, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
               const __grid_constant__ CUtensorMap mW0, const __grid_constant__ CUtensorMap mW1,
               const __grid_constant__ CUtensorMap mCf, const __grid_constant__ CUtensorMap mChi,
               const __grid_constant__ CUtensorMap mClo, const __grid_constant__ CUtensorMap mCh,
               const __grid_constant__ CUtensorMap mR, GemmP p) {
  typedef Cfg<KIND, BN, STAGES, CG> G;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* tiles = smem;                                               // STAGES x [A planes | W planes]
  float* epi_smem = reinterpret_cast<float*>(smem + STAGES * G::STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * G::STAGE_BYTES + G::EPI_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;      // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint64_t* rbar = tempty + 2;           // [8] residual boxes (TMA-loaded fp16 residual), one per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.K + G::BK - 1) / G::BK;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = ((p.M + p.tile_rows - 1) / p.tile_rows + CG - 1) / CG;  // pair tiles when CG == 2 (a tail pair may be half empty)
  const int ntiles = tiles_n * tiles_m;
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;
  const int unit0 = blockIdx.x / CG, nunits = gridDim.x / CG;                 // persistent walk, one unit = one CTA (pair)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mA0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mW0)) : "memory");
    if (G::PLANES == 2) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mA1)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mW1)) : "memory");
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8 * CG);     // one arrival per epilogue warp (of both CTAs of a pair)
    }
    for (int w = 0; w < 8; ++w) mbar_init(&rbar[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(G::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(G::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();        // the peer's barriers are initialised before anything signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    {
      uint32_t it = 0;                                   // running k-block counter across tiles
      for (int t = unit0; t < ntiles; t += nunits) {
        const int m0 = ((t / tiles_n) * CG + (int)crank) * p.tile_rows, n0 = (t % tiles_n) * BN + (int)crank * G::BN_CTA;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          unsigned char* st = tiles + s * G::STAGE_BYTES;
          const int k0 = kb * G::BK;
          if (CG == 2) {
            // both CTAs' bytes are counted on the leader's barrier (armed by the leader alone, for the pair)
            const uint32_t fb = mapa_rank(smem_u32(&full[s]), 0);
            if (elect_one()) {
              if (crank == 0) mbar_arrive_expect_tx(&full[s], 2 * G::STAGE_BYTES);
              tma_load_2d_pair(st, &mA0, fb, k0, m0);
              if (G::PLANES == 2) tma_load_2d_pair(st + G::A_BYTES, &mA1, fb, k0, m0);
              tma_load_2d_pair(st + G::PLANES * G::A_BYTES, &mW0, fb, k0, n0);
              if (G::PLANES == 2) tma_load_2d_pair(st + G::PLANES * G::A_BYTES + G::W_BYTES, &mW1, fb, k0, n0);
            }
          } else if (elect_one()) {
            mbar_arrive_expect_tx(&full[s], G::STAGE_BYTES);
            tma_load_2d(st, &mA0, &full[s], k0, m0);
            if (G::PLANES == 2) tma_load_2d(st + G::A_BYTES, &mA1, &full[s], k0, m0);
            tma_load_2d(st + G::PLANES * G::A_BYTES, &mW0, &full[s], k0, n0);
            if (G::PLANES == 2) tma_load_2d(st + G::PLANES * G::A_BYTES + G::W_BYTES, &mW1, &full[s], k0, n0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    if (crank == 0) {
      constexpr uint32_t idesc = umma_idesc(KIND == KIND_TF32 ? 2 : (KIND == KIND_BF16X3 ? 1 : 0), G::BM * CG, BN);
      uint32_t it = 0, lt = 0;                           // lt = local tile counter
      for (int t = unit0; t < ntiles; t += nunits, ++lt) {
        const uint32_t as = lt & 1;
        // the epilogue (of both CTAs) has drained this accumulator
        if (CG == 2) mbar_wait_cluster(&tempty[as], ((lt >> 1) & 1) ^ 1);
        else mbar_wait(&tempty[as], ((lt >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + as * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a0 = smem_u32(tiles + s * G::STAGE_BYTES);
          const uint32_t w0 = a0 + G::PLANES * G::A_BYTES;
          const uint64_t dA0 = umma_desc(a0), dW0 = umma_desc(w0);
          const uint64_t dA1 = umma_desc(a0 + G::A_BYTES), dW1 = umma_desc(w0 + G::W_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < G::BK / G::UK; ++k) {
              if (QAT_DBG(p, 2)) break;
              const uint64_t adv = uint64_t((k * G::UK * (KIND == KIND_TF32 ? 4 : 2)) >> 4);   // +32 bytes per MMA
              if (G::PLANES == 1) {
                umma<KIND, CG>(tacc, dA0 + adv, dW0 + adv, idesc, (kb | k) != 0);
              } else {
                umma<KIND, CG>(tacc, dA1 + adv, dW0 + adv, idesc, (kb | k) != 0);    // lo * hi
                umma<KIND, CG>(tacc, dA0 + adv, dW1 + adv, idesc, 1);                // hi * lo
                umma<KIND, CG>(tacc, dA0 + adv, dW0 + adv, idesc, 1);                // hi * hi
              }
            }
            // smem slot reusable once these MMAs have read it (in both CTAs of a pair)
            if (CG == 2) umma_commit_pair(&empty[s]);
            else umma_commit(&empty[s]);
          }
          __syncwarp();
        }
        // accumulator complete (tcgen05.commit covers the MMAs of the issuing thread: the same elected lane)
        if (elect_one()) {
          if (CG == 2) umma_commit_pair(&tfull[as]);
          else umma_commit(&tfull[as]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps 2..9 =====================
    // warp w may touch TMEM lanes 32*(w%4)..+31; warps 2..5 take columns [0, BN/2) of the tile, warps 6..9 the rest.
    // Per 32-column chunk: tcgen05.ld (lane = row) -> bias / ReLU / residual in registers -> the warp's 32 x 32 box in
    // shared memory in the TMA swizzle layout -> one lane issues cp.async.bulk.tensor stores (one per output plane);
    // the TMA unit clips row / column tails, so the LSU never sees the (row-strided) output.
    const bool out_f = EPI == 0 ? p.C.f != nullptr : EPI == 2, out_hi = EPI == 0 ? (p.C.hi != nullptr || p.C.ph != nullptr) : EPI == 3;
    const bool out_lo = EPI == 0 ? (p.C.lo != nullptr || p.C.pl != nullptr) : EPI == 3, out_h = EPI == 0 ? p.C.h != nullptr : EPI == 1;
    const float* const res_f = EPI == 1 ? nullptr : p.epi.res;
    const __half* const res_h = (EPI == 1 && !p.tma_res) ? nullptr : p.epi.res_h;
    const int ew = warp - 2;
    const int lg = warp & 3;
    const int half = ew >> 2;
    constexpr int HALF_N = BN / 2;
    constexpr int NCH = HALF_N / 32;
    unsigned char* box = reinterpret_cast<unsigned char*>(epi_smem) + ew * 4096;
    const uint32_t sbox = smem_u32(box);
    const uint32_t sw128 = uint32_t(lane & 7), sw64 = uint32_t((lane >> 1) & 3);
    uint32_t lt = 0, nres = 0;
    // accumulator hand-back: the leader's barrier (a remote arrival from the peer CTA)
    const uint32_t tempty_a0 = CG == 2 ? mapa_rank(smem_u32(&tempty[0]), 0) : smem_u32(&tempty[0]);
    const uint32_t tempty_a1 = CG == 2 ? mapa_rank(smem_u32(&tempty[1]), 0) : smem_u32(&tempty[1]);
    for (int t = unit0; t < ntiles; t += nunits, ++lt) {
      const int m0 = ((t / tiles_n) * CG + (int)crank) * p.tile_rows, n0 = (t % tiles_n) * BN;
      const uint32_t as = lt & 1;
      mbar_wait(&tfull[as], (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + as * BN + half * HALF_N + (uint32_t(lg * 32) << 16);
      const long long row = (long long)m0 + lg * 32 + lane;
      const bool row_ok = row < p.M;
      if (QAT_DBG(p, 1)) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster((as ? tempty_a1 : tempty_a0));
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((as ? tempty_a1 : tempty_a0)) : "memory");
        }
        continue;
      }
      if (EPI == 4) {
        // ---- reduce epilogue: Gaussian-weighted temporal reduction of the tile's samples (Epi::coef) ----------------------
        // The tile holds tile_rows / T whole samples.  A warp's 32 rows touch at most two of them (T >= 32): rows [0, split)
        // belong to the first, the rest to the next.  Per 32-column chunk every lane scales its row by coef[row, expert] and the
        // warp sums columns over its rows with warp_colsum32 (twice when it straddles two samples); the per-warp partial sums
        // (lane = column) are parked in the warp's box, double-buffered over tiles, and after ONE named barrier per tile the
        // 256 epilogue threads add the partials of the four row quarters in order and store sample x column sums -- no
        // atomics, bit-reproducible.
        const int TT = p.epi.coef_T, TR = p.tile_rows;
        const int r_in = lg * 32 + lane;
        const bool rvalid = r_in < TR && row_ok;
        const int samp0 = (lg * 32) / TT;
        const int split = (samp0 + 1) * TT - lg * 32;                 // >= 1
        const bool two = split < 32 && lg * 32 + split < TR;           // warp-uniform
        const bool inA = lane < split;
        const float* crow = p.epi.coef + (rvalid ? (row % p.epi.coef_mod) : 0) * p.epi.coef_E;
        float ps0[NCH], ps1[NCH];
        uint32_t v[32];
        if (n0 + half * HALF_N < p.N) tmem_ld32(tacc, v);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int col0 = n0 + half * HALF_N + c * 32;
          const bool live = col0 < p.N;                       // warp-uniform
          float cf = 0.f;
          float4 bb[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            bb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && p.epi.bias && col0 + 4 * q < p.N) bb[q] = __ldg(reinterpret_cast<const float4*>(p.epi.bias + col0 + 4 * q));
          }
          if (live && rvalid) cf = __ldg(crow + col0 / p.epi.coef_Ch);
          if (live) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float x[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            x[4 * q] = __uint_as_float(v[4 * q]) + bb[q].x;
            x[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + bb[q].y;
            x[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + bb[q].z;
            x[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + bb[q].w;
          }
          if (c + 1 < NCH && col0 + 32 < p.N) tmem_ld32(tacc + uint32_t((c + 1) * 32), v);
          if (c == NCH - 1) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) {
              if (CG == 2) mbar_arrive_cluster((as ? tempty_a1 : tempty_a0));
              else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((as ? tempty_a1 : tempty_a0)) : "memory");
            }
          }
          ps0[c] = 0.f;
          ps1[c] = 0.f;
          if (!live) continue;
          // rows beyond the tile's samples (or beyond M) contribute exact zeros, whatever their accumulator holds
          if (p.epi.relu == 1) {
#pragma unroll
            for (int e = 0; e < 32; ++e) x[e] = rvalid ? fmaxf(x[e], 0.f) * cf : 0.f;
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) x[e] = rvalid ? epi_act(x[e], p.epi.relu) * cf : 0.f;
          }
          if (two) {
            float y[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              y[e] = inA ? 0.f : x[e];
              x[e] = inA ? x[e] : 0.f;
            }
            ps0[c] = warp_colsum32(x, lane);
            ps1[c] = warp_colsum32(y, lane);
          } else {
            ps0[c] = warp_colsum32(x, lane);
          }
        }
        float* pb = reinterpret_cast<float*>(box + (lt & 1) * 1024);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          pb[c * 32 + lane] = ps0[c];
          pb[(NCH + c) * 32 + lane] = ps1[c];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");            // the eight epilogue warps
        {
          const int q = ew * 32 + lane;                             // column of the tile this thread combines
          const int hq = q / HALF_N, cq = (q % HALF_N) / 32, jq = q % 32;
          const int nsamp = TR / TT;
          for (int s = 0; s < nsamp; ++s) {
            float sum = 0.f;
#pragma unroll
            for (int g = 0; g < 4; ++g) {                           // row quarters in order
              const int f = (g * 32) / TT;
              const int sp = (f + 1) * TT - g * 32;
              int slot = -1;
              if (s == f) slot = 0;
              else if (s == f + 1 && sp < 32 && g * 32 + sp < TR) slot = 1;
              if (slot >= 0) {
                const int ewg = hq * 4 + ((g + 2) & 3);             // the warp of column half hq that owns TMEM lane quarter g
                const float* ob = reinterpret_cast<const float*>(reinterpret_cast<unsigned char*>(epi_smem) + ewg * 4096 + (lt & 1) * 1024);
                sum += ob[(slot * NCH + cq) * 32 + jq];
              }
            }
            const long long srow = (long long)m0 / TT + s;
            if (srow * TT < p.M && n0 + q < p.N) {
              p.C.f[srow * p.C.ld + n0 + q] = sum;
              if (p.C.hi) {                                         // planes for a tensor-core consumer of the reduced tile
                bf16 sh, sl;
                split_bf16(sum, sh, sl);
                p.C.hi[srow * p.C.ld + n0 + q] = sh;
                p.C.lo[srow * p.C.ld + n0 + q] = sl;
              }
            }
          }
        }
        continue;
      }
      // the TMEM load of chunk c + 1 is issued as soon as chunk c has left its registers, so it is in flight while chunk c
      // is packed, parked in shared memory and handed to the TMA unit
      uint32_t v[32];
      if (n0 + half * HALF_N < p.N) tmem_ld32(tacc, v);
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int col0 = n0 + half * HALF_N + c * 32;
        const bool live = col0 < p.N;                       // warp-uniform
        if (live && p.tma_res) {
          // fp16 residual box (32 rows x 32 columns, 64B swizzle) into the upper half of the warp's slot; rows / columns
          // beyond the matrix are zero-filled by the TMA unit
          // (the previous chunk's lanes finished reading this half before their fence.proxy.async + __syncwarp below;
          //  the output box is the lower half, so an in-flight store of the previous chunk does not touch it)
          if (elect_one()) {
            mbar_arrive_expect_tx(&rbar[ew], 2048);
            tma_load_2d(box + 2048, &mR, &rbar[ew], col0, m0 + lg * 32);
          }
        }
        // operands of the epilogue are fetched while the TMEM load is in flight
        float4 bb[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          bb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live && p.epi.bias && col0 + 4 * q < p.N) bb[q] = __ldg(reinterpret_cast<const float4*>(p.epi.bias + col0 + 4 * q));
        }
        float4 rr[8];
        if (res_f) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            rr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && row_ok && col0 + 4 * q < p.N) rr[q] = __ldg(reinterpret_cast<const float4*>(res_f + row * p.epi.ldr + col0 + 4 * q));
          }
        } else if (EPI != 1 && res_h && !p.tma_res) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            rr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && row_ok && col0 + 4 * q < p.N) {
              const uint2 u = __ldg(reinterpret_cast<const uint2*>(res_h + row * p.epi.ldr + col0 + 4 * q));
              const float2 r0 = unpack_half2(u.x), r1 = unpack_half2(u.y);
              rr[q] = make_float4(r0.x, r0.y, r1.x, r1.y);
            }
          }
        }
        if (live) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float x[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          x[4 * q] = __uint_as_float(v[4 * q]) + bb[q].x;
          x[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + bb[q].y;
          x[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + bb[q].z;
          x[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + bb[q].w;
        }
        if (c + 1 < NCH && col0 + 32 < p.N) tmem_ld32(tacc + uint32_t((c + 1) * 32), v);
        if (c == NCH - 1) {
          // last TMEM read of this tile by this warp: hand the accumulator back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster((as ? tempty_a1 : tempty_a0));
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((as ? tempty_a1 : tempty_a0)) : "memory");
          }
        }
        if (!live) continue;
        if (p.epi.relu == 1) {
#pragma unroll
          for (int e = 0; e < 32; ++e) x[e] = fmaxf(x[e], 0.f);
        } else if (p.epi.relu != 0) {
#pragma unroll
          for (int e = 0; e < 32; ++e) x[e] = epi_act(x[e], p.epi.relu);
        }
        if (p.tma_res) {
          mbar_wait(&rbar[ew], nres & 1);
          ++nres;
#pragma unroll
          for (uint32_t q = 0; q < 4; ++q) {
            const uint4 u = *reinterpret_cast<const uint4*>(box + 2048 + lane * 64 + ((q ^ sw64) << 4));
            const float2 r0 = unpack_half2(u.x), r1 = unpack_half2(u.y), r2 = unpack_half2(u.z), r3 = unpack_half2(u.w);
            x[8 * q] += r0.x; x[8 * q + 1] += r0.y; x[8 * q + 2] += r1.x; x[8 * q + 3] += r1.y;
            x[8 * q + 4] += r2.x; x[8 * q + 5] += r2.y; x[8 * q + 6] += r3.x; x[8 * q + 7] += r3.y;
          }
        } else if (EPI != 1 && (res_f || res_h)) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            x[4 * q] += rr[q].x; x[4 * q + 1] += rr[q].y; x[4 * q + 2] += rr[q].z; x[4 * q + 3] += rr[q].w;
          }
        }
        const int crow = m0 + lg * 32;
        if (out_f) {
          if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the box is free again
          __syncwarp();
#pragma unroll
          for (uint32_t q = 0; q < 8; ++q)
            sts128(sbox + lane * 128 + ((q ^ sw128) << 4), __float_as_uint(x[4 * q]), __float_as_uint(x[4 * q + 1]),
                   __float_as_uint(x[4 * q + 2]), __float_as_uint(x[4 * q + 3]));
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (elect_one() && !QAT_DBG(p, 4)) {
            tma_store_2d(&mCf, box, col0, crow);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (out_hi) {
          if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (uint32_t q = 0; q < 4; ++q) {
            uint32_t hw[4], lw[4];
            if (p.pair_f16) {
#pragma unroll
              for (int e = 0; e < 4; ++e) split_half2(x[8 * q + 2 * e], x[8 * q + 2 * e + 1], hw[e], lw[e]);
            } else
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              bf16 h0, h1, l0, l1;
              split_bf16(x[8 * q + 2 * e], h0, l0);
              split_bf16(x[8 * q + 2 * e + 1], h1, l1);
              __nv_bfloat162 hh = __halves2bfloat162(h0, h1), ll = __halves2bfloat162(l0, l1);
              hw[e] = *reinterpret_cast<uint32_t*>(&hh);
              lw[e] = *reinterpret_cast<uint32_t*>(&ll);
            }
            sts128(sbox + lane * 64 + ((q ^ sw64) << 4), hw[0], hw[1], hw[2], hw[3]);
            sts128(sbox + 2048 + lane * 64 + ((q ^ sw64) << 4), lw[0], lw[1], lw[2], lw[3]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (elect_one() && !QAT_DBG(p, 4)) {
            tma_store_2d(&mChi, box, col0, crow);
            if (out_lo) tma_store_2d(&mClo, box + 2048, col0, crow);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (out_h) {
          // an fp16 box is 2 KB: when it is the only plane the two halves of the warp's 4 KB slot alternate
          const bool solo = !out_f && !out_hi && !p.tma_res;
          const uint32_t hb = solo ? uint32_t(c & 1) * 2048u : 0u;
          if (elect_one()) {
            if (solo) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          __syncwarp();
#pragma unroll
          for (uint32_t q = 0; q < 4; ++q)
            sts128(sbox + hb + lane * 64 + ((q ^ sw64) << 4), pack_half2(x[8 * q], x[8 * q + 1]), pack_half2(x[8 * q + 2], x[8 * q + 3]),
                   pack_half2(x[8 * q + 4], x[8 * q + 5]), pack_half2(x[8 * q + 6], x[8 * q + 7]));
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (elect_one() && !QAT_DBG(p, 4)) {
            tma_store_2d(&mCh, box + hb, col0, crow);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
    }
    if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (CG == 2) cluster_sync_all();        // no CTA of the pair leaves (or frees TMEM) while the other may still signal / be read
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G::TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G::TMEM_COLS) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------------
// rows x K matrix, K contiguous, leading dimension ld (elements); box = 128 bytes x box_rows, 128B swizzle
inline int make_map(CUtensorMap* m, const void* base, CUtensorMapDataType dt, long long rows, int K, long long ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) QAT_FAIL(QAT_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  const int elem_bytes = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, dt, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) QAT_FAIL(QAT_ECUDA, "cuTensorMapEncodeTiled failed with code %d (rows=%lld K=%d ld=%lld)", (int)r, rows, K, ld);
  return QAT_OK;
}

// a [rows, cols] output plane with leading dimension ld; box = 32 rows x 32 columns (128 B rows: SWIZZLE_128B for 4-byte
// elements, 64 B rows: SWIZZLE_64B for 2-byte elements)
inline int make_store_map(CUtensorMap* m, const void* base, CUtensorMapDataType dt, long long rows, int cols, long long ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) QAT_FAIL(QAT_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  const int eb = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 2;
  QAT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * eb) % 16 == 0,
              "gemm_tc: output planes must be 16-byte aligned with a 16-byte multiple row pitch (ld=%lld)", ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * eb};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, dt, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  eb == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) QAT_FAIL(QAT_ECUDA, "cuTensorMapEncodeTiled (store map) failed with code %d (rows=%lld cols=%d ld=%lld)", (int)r, rows, cols, ld);
  return QAT_OK;
}

template <int KIND, int BN, int STAGES, int CG = 1>
inline int launch(Ctx& c, const void* A0, const void* A1, long long lda, long long M, const void* W0, const void* W1, int N, int K,
                  const Epi& epi, const Act& C) {
  typedef Cfg<KIND, BN, STAGES, CG> G;
  static_assert(G::SMEM <= 227 * 1024, "gemm_tc: shared-memory budget exceeded");
  const CUtensorMapDataType dt = KIND == KIND_TF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : (KIND == KIND_BF16X3 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  CUtensorMap mA0, mA1, mW0, mW1, mCf, mChi, mClo, mCh, mR;
  QAT_TRY(make_map(&mA0, A0, dt, M, K, lda, G::BM));
  QAT_TRY(make_map(&mW0, W0, dt, N, K, K, G::BN_CTA));
  if (G::PLANES == 2) {
    QAT_TRY(make_map(&mA1, A1, dt, M, K, lda, G::BM));
    QAT_TRY(make_map(&mW1, W1, dt, N, K, K, G::BN_CTA));
  } else {
    mA1 = mA0;
    mW1 = mW0;
  }
  QAT_REQUIRE(C.f || C.hi || C.h || C.ph, "gemm_tc: no output plane");
  const Act Cstore = epi.coef ? Act() : C;      // the reduce epilogue stores through the LSU ([samples, N], not [M, N])
  QAT_REQUIRE(!C.lo || C.hi, "gemm_tc: a lo plane needs its hi plane");
  QAT_REQUIRE((C.ph != nullptr) == (C.pl != nullptr) && !(C.ph && C.hi), "gemm_tc: the fp16 pair needs both planes and excludes the bf16 pair");
  mCf = mChi = mClo = mCh = mA0;
  if (Cstore.f) QAT_TRY(make_store_map(&mCf, C.f, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, M, N, C.ld));
  if (C.hi && !epi.coef) QAT_TRY(make_store_map(&mChi, C.hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, M, N, C.ld));
  if (C.lo && !epi.coef) QAT_TRY(make_store_map(&mClo, C.lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, M, N, C.ld));
  if (C.h) QAT_TRY(make_store_map(&mCh, C.h, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, M, N, C.ld));
  if (C.ph) {
    QAT_TRY(make_store_map(&mChi, C.ph, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, M, N, C.ld));
    QAT_TRY(make_store_map(&mClo, C.pl, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, M, N, C.ld));
  }
  GemmP p;
  p.pair_f16 = C.ph ? 1 : 0;
  p.M = (int)M; p.N = N; p.K = K; p.epi = epi; p.C = C;
  p.tile_rows = G::BM;
  if (epi.coef) {
    QAT_REQUIRE(BN == 256 && C.f && (C.hi != nullptr) == (C.lo != nullptr) && !C.h && !C.ph && !epi.res && !epi.res_h,
                "gemm_tc: the reduce epilogue writes one fp32 plane (+ optional bf16 hi / lo planes) from 256-wide tiles");
    QAT_REQUIRE(epi.coef_T >= 32 && epi.coef_T <= 128 && M % epi.coef_T == 0 && epi.coef_Ch % 32 == 0 && epi.coef_E >= 1 && epi.coef_mod >= 1,
                "gemm_tc: reduce epilogue needs 32 <= T <= 128, M %% T == 0 and Ch %% 32 == 0 (T=%d Ch=%d)", epi.coef_T, epi.coef_Ch);
    p.tile_rows = (G::BM / epi.coef_T) * epi.coef_T;
  }
  // fp16 residual next to an fp16-only output: through the TMA unit as well (lane-per-row LSU loads of a row-strided
  // residual cost 32 L1 wavefronts per instruction and made the PatchSelecter out-proj epilogue-bound)
  mR = mA0;
  p.tma_res = 0;
  if (epi.res_h && !epi.res && C.h && !C.f && !C.hi && !C.ph && (reinterpret_cast<uintptr_t>(epi.res_h) & 15) == 0 && epi.ldr % 8 == 0) {
    QAT_TRY(make_store_map(&mR, epi.res_h, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, M, N, epi.ldr));
    p.tma_res = 1;
  }
  p.dbg = 0;
#ifdef QAT_PROFILE_BUILD
  if (const char* dbg = getenv("QAT_GEMM_DBG")) p.dbg = atoi(dbg);
#endif
  // persistent units: CTAs, or CTA pairs walking 256-row pair tiles
  const long long ntiles = (long long)((N + BN - 1) / BN) * (((M + p.tile_rows - 1) / p.tile_rows + CG - 1) / CG);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long units = sms / CG;
  const unsigned grid = (unsigned)(ntiles < units ? ntiles : units) * CG;
  int mode = 0;
  if (KIND != KIND_BF16X3 && C.h && !C.f && !C.hi && !C.ph && !epi.res && (!epi.res_h || p.tma_res)) mode = 1;
  else if (C.f && !C.hi && !C.lo && !C.h && !C.ph) mode = 2;
  else if (KIND == KIND_BF16X3 && ((C.hi && C.lo) || C.ph) && !C.f && !C.h) mode = 3;
#define QAT_GEMM_LAUNCH(EPI_)                                                                                 \
  do {                                                                                                        \
    auto kern = gemm_tc_kernel<KIND, BN, STAGES, EPI_, CG>;                                                   \
    QAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));               \
    cudaLaunchConfig_t cfg = {};                                                                              \
    cfg.gridDim = dim3(grid);                                                                                 \
    cfg.blockDim = dim3(320);                                                                                 \
    cfg.dynamicSmemBytes = G::SMEM;                                                                           \
    cfg.stream = c.stream;                                                                                    \
    cudaLaunchAttribute at[1];                                                                                \
    at[0].id = cudaLaunchAttributeClusterDimension;                                                           \
    at[0].val.clusterDim.x = CG;                                                                              \
    at[0].val.clusterDim.y = 1;                                                                               \
    at[0].val.clusterDim.z = 1;                                                                               \
    cfg.attrs = at;                                                                                           \
    cfg.numAttrs = CG == 2 ? 1 : 0;                                                                           \
    QAT_CUDA(cudaLaunchKernelEx(&cfg, kern, mA0, mA1, mW0, mW1, mCf, mChi, mClo, mCh, mR, p));                \
  } while (0)
  if (epi.coef) {
    // instantiated for the kind / tile the expert GEMM uses (bf16x3, 256-wide tiles, single CTA or CTA pair)
    if (KIND == KIND_BF16X3 && BN == 256) QAT_GEMM_LAUNCH((KIND == KIND_BF16X3 && BN == 256 ? 4 : 0));
    else QAT_FAIL(QAT_EINVAL, "gemm_tc: the reduce epilogue is built for the bf16x3 kind with 256-wide tiles");
  } else if (mode == 1) QAT_GEMM_LAUNCH((KIND == KIND_BF16X3 ? 0 : 1));
  else if (mode == 2) QAT_GEMM_LAUNCH(2);
  else if (mode == 3) QAT_GEMM_LAUNCH((KIND == KIND_BF16X3 ? 3 : 0));
  else QAT_GEMM_LAUNCH(0);
#undef QAT_GEMM_LAUNCH
  return check_launch("gemm_tc_kernel");
}

inline int check_common(long long M, int N, int K, long long lda, const Epi& epi, const Act& C, int eb) {
  QAT_REQUIRE(M > 0 && M < (1ll << 31), "gemm_tc: M=%lld out of range", M);
  QAT_REQUIRE(N % 4 == 0, "gemm_tc: N=%d must be a multiple of 4", N);
  QAT_REQUIRE((lda * eb) % 16 == 0 && (K * eb) % 16 == 0, "gemm_tc: row strides must be multiples of 16 bytes");
  QAT_REQUIRE(C.ld % 4 == 0 && ((epi.res == nullptr && epi.res_h == nullptr) || epi.ldr % 4 == 0),
              "gemm_tc: output/residual leading dims must be multiples of 4");
  return QAT_OK;
}

inline bool wide_tile(const Ctx& c, int N) {
  if (c.opt.v[OPT_GEMM_BN] > 0) return c.opt.v[OPT_GEMM_BN] == 256;
  return N % 256 == 0;
}
// CTA pairs (cta_group::2) for the wide tile; option gemm_cg = 1 / 2 is the A/B switch.  Measured on B200 (tests/prof_gemm.py):
// bf16x3 +12 % (475 vs 425 nominal TFLOP/s at [15360,512]x[3584,512]^T), tf32 +4..10 %, f16 -3..7 % (its MMA time per tile is
// close to its epilogue time, and a pair couples the two CTAs' epilogues to one accumulator hand-back), hence the defaults.
inline bool pair_tile(const Ctx& c, long long M, bool dflt) {
  if (c.opt.v[OPT_GEMM_CG] > 0) return c.opt.v[OPT_GEMM_CG] == 2;
  return dflt && M > 128;
}

}  // namespace tc

inline int gemm_tc_tf32(Ctx& c, const float* A, int lda, long long M, const float* W, int N, int K, const Epi& epi, Act C) {
  if (M == 0 || N == 0) return QAT_OK;
  QAT_TRY(tc::check_common(M, N, K, lda, epi, C, 4));
  QAT_REQUIRE((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) % 16 == 0, "gemm_tc: operands must be 16-byte aligned");
  if (c.dry) return QAT_OK;
  if (tc::wide_tile(c, N) && tc::pair_tile(c, M, true)) return tc::launch<tc::KIND_TF32, 256, 6, 2>(c, A, nullptr, lda, M, W, nullptr, N, K, epi, C);
  if (tc::wide_tile(c, N)) return tc::launch<tc::KIND_TF32, 256, 4>(c, A, nullptr, lda, M, W, nullptr, N, K, epi, C);
  return tc::launch<tc::KIND_TF32, 128, 6>(c, A, nullptr, lda, M, W, nullptr, N, K, epi, C);
}

inline int gemm_tc_bf16x3(Ctx& c, const bf16* Ahi, const bf16* Alo, int lda, long long M, const bf16* Whi, const bf16* Wlo, int N, int K,
                          const Epi& epi, Act C) {
  if (M == 0 || N == 0) return QAT_OK;
  QAT_TRY(tc::check_common(M, N, K, lda, epi, C, 2));
  if (c.dry) return QAT_OK;
  QAT_REQUIRE((reinterpret_cast<uintptr_t>(Ahi) | reinterpret_cast<uintptr_t>(Alo) | reinterpret_cast<uintptr_t>(Whi) |
               reinterpret_cast<uintptr_t>(Wlo)) % 16 == 0, "gemm_tc: operands must be 16-byte aligned");
  if (tc::wide_tile(c, N) && tc::pair_tile(c, M, true)) return tc::launch<tc::KIND_BF16X3, 256, 3, 2>(c, Ahi, Alo, lda, M, Whi, Wlo, N, K, epi, C);
  if (tc::wide_tile(c, N)) return tc::launch<tc::KIND_BF16X3, 256, 2>(c, Ahi, Alo, lda, M, Whi, Wlo, N, K, epi, C);
  return tc::launch<tc::KIND_BF16X3, 128, 3>(c, Ahi, Alo, lda, M, Whi, Wlo, N, K, epi, C);
}

inline int gemm_tc_f16(Ctx& c, const __half* A, int lda, long long M, const __half* W, int N, int K, const Epi& epi, Act C) {
  if (M == 0 || N == 0) return QAT_OK;
  QAT_TRY(tc::check_common(M, N, K, lda, epi, C, 2));
  if (c.dry) return QAT_OK;
  QAT_REQUIRE((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) % 16 == 0, "gemm_tc: operands must be 16-byte aligned");
  if (tc::wide_tile(c, N) && tc::pair_tile(c, M, false)) return tc::launch<tc::KIND_F16, 256, 6, 2>(c, A, nullptr, lda, M, W, nullptr, N, K, epi, C);
  if (tc::wide_tile(c, N)) return tc::launch<tc::KIND_F16, 256, 4>(c, A, nullptr, lda, M, W, nullptr, N, K, epi, C);
  return tc::launch<tc::KIND_F16, 128, 6>(c, A, nullptr, lda, M, W, nullptr, N, K, epi, C);
}

}  // namespace qat
