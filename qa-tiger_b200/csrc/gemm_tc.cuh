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
// Epilogues: the generic one (any mix of fp32 / bf16 hi,lo / fp16 output planes, smem transpose -> 16-byte
// stores) and, when the only output is an fp16 plane, a TMA-store one: the warp packs its 32 x 64 sub-tile into
// a 128B-swizzled smem box and one lane issues cp.async.bulk.tensor (row / column tails clipped by the TMA unit).
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0      : TMA producer  - cp.async.bulk.tensor.2d (128B swizzle) into a STAGES-deep smem ring
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit frees smem slots
//   warps 2..5  : epilogue - tcgen05.ld (32 lanes x 32 columns) -> smem transpose -> bias/ReLU/residual
//                 -> coalesced 16-byte stores of fp32 and/or bf16 (hi, lo) planes
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

template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
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

struct GemmP {
  int M, N, K;
  Epi epi;
  Act C;
  int tma_store;   // 1: the only output is C.h and it leaves through the mC tensor map
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

template <int KIND, int BN, int STAGES>
struct Cfg {
  static constexpr int BM = 128;
  static constexpr int ROW_BYTES = 128;                                   // one swizzle row
  static constexpr int BK = KIND == KIND_TF32 ? 32 : 64;                  // elements per k-block
  static constexpr int UK = KIND == KIND_TF32 ? 8 : 16;                   // elements per MMA
  static constexpr int PLANES = KIND == KIND_BF16X3 ? 2 : 1;
  static constexpr int A_BYTES = BM * ROW_BYTES;                          // per plane
  static constexpr int W_BYTES = BN * ROW_BYTES;
  static constexpr int STAGE_BYTES = PLANES * (A_BYTES + W_BYTES);
  static constexpr int EPI_BYTES = 4 * 2 * 4096;                          // generic: 4 warps x [32][36] floats; TMA store: 4 warps x 2 boxes x 4 KB
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;   // + alignment slack
  static constexpr int TMEM_COLS = 2 * BN;                                // two accumulators: epilogue(i) overlaps mainloop(i+1)
  static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");
};

// Persistent: grid = min(#tiles, #SMs); CTA c walks tiles c, c+grid, ... (n fastest, so co-scheduled CTAs share
// the A rows through L2).  Three decoupled loops meet only through mbarriers:
//   full/empty[STAGES]   TMA <-> MMA    (smem ring, continuous across tiles)
//   tfull/tempty[2]      MMA <-> epilogue (the two TMEM accumulators)
template <int KIND, int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
               const __grid_constant__ CUtensorMap mW0, const __grid_constant__ CUtensorMap mW1,
               const __grid_constant__ CUtensorMap mC, GemmP p) {
  typedef Cfg<KIND, BN, STAGES> G;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* tiles = smem;                                               // STAGES x [A planes | W planes]
  float* epi_smem = reinterpret_cast<float*>(smem + STAGES * G::STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * G::STAGE_BYTES + G::EPI_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;      // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.K + G::BK - 1) / G::BK;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + G::BM - 1) / G::BM;
  const int ntiles = tiles_n * tiles_m;

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
      mbar_init(&tempty[a], 4);          // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(G::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;                                   // running k-block counter across tiles
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * G::BM, n0 = (t % tiles_n) * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          unsigned char* st = tiles + s * G::STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], G::STAGE_BYTES);
          const int k0 = kb * G::BK;
          tma_load_2d(st, &mA0, &full[s], k0, m0);
          if (G::PLANES == 2) tma_load_2d(st + G::A_BYTES, &mA1, &full[s], k0, m0);
          tma_load_2d(st + G::PLANES * G::A_BYTES, &mW0, &full[s], k0, n0);
          if (G::PLANES == 2) tma_load_2d(st + G::PLANES * G::A_BYTES + G::W_BYTES, &mW1, &full[s], k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(KIND == KIND_TF32 ? 2 : (KIND == KIND_BF16X3 ? 1 : 0), G::BM, BN);
      uint32_t it = 0, lt = 0;                           // lt = local tile counter
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++lt) {
        const uint32_t as = lt & 1;
        mbar_wait(&tempty[as], ((lt >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
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
#pragma unroll
          for (int k = 0; k < G::BK / G::UK; ++k) {
            const uint64_t adv = uint64_t((k * G::UK * (KIND == KIND_TF32 ? 4 : 2)) >> 4);   // +32 bytes per MMA
            if (G::PLANES == 1) {
              umma<KIND>(tacc, dA0 + adv, dW0 + adv, idesc, (kb | k) != 0);
            } else {
              umma<KIND>(tacc, dA1 + adv, dW0 + adv, idesc, (kb | k) != 0);    // lo * hi
              umma<KIND>(tacc, dA0 + adv, dW1 + adv, idesc, 1);                // hi * lo
              umma<KIND>(tacc, dA0 + adv, dW0 + adv, idesc, 1);                // hi * hi
            }
          }
          umma_commit(&empty[s]);            // smem slot reusable once these MMAs have read it
        }
        umma_commit(&tfull[as]);             // accumulator complete
      }
    }
  } else {
    // ===================== epilogue warps 2..5: TMEM lanes 32*(warp%4) =====================
    const int lg = warp & 3;
    float* tile = epi_smem + (warp - 2) * (32 * 36);
    const int rq = lane >> 3, cq = lane & 7;
    uint32_t lt = 0;
    if (p.tma_store) {
      // ---- fp16-only output through the TMA unit: lane = row of the warp's 32 x 64 box ----
      unsigned char* stg = reinterpret_cast<unsigned char*>(epi_smem) + (warp - 2) * 8192;
      uint32_t nbox = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++lt) {
        const int m0 = (t / tiles_n) * G::BM, n0 = (t % tiles_n) * BN;
        const uint32_t as = lt & 1;
        mbar_wait(&tfull[as], (lt >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + as * BN + (uint32_t(lg * 32) << 16);
        const long long row = (long long)m0 + lg * 32 + lane;
        for (int c0 = 0; c0 < BN; c0 += 64) {
          if (n0 + c0 >= p.N) break;
          uint32_t v[32], u[32];
          tmem_ld32(tacc + uint32_t(c0), v);
          tmem_ld32(tacc + uint32_t(c0 + 32), u);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c0 + 64 >= BN || n0 + c0 + 64 >= p.N) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[as])) : "memory");
          }
          unsigned char* box = stg + (nbox & 1) * 4096;
          ++nbox;
          // the box written two stores ago must have been read out by the TMA unit
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          const int colb = n0 + c0;
#pragma unroll
          for (int q = 0; q < 8; ++q) {          // 8 columns = one 16-byte chunk of the row
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(q < 4 ? v[8 * q + e] : u[8 * (q - 4) + e]);
            const int col = colb + 8 * q;
            if (col < p.N) {                     // N is a multiple of 8 on this path (host check)
              if (p.epi.bias) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.epi.bias + col));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.epi.bias + col + 4));
                x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
                x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
              }
              if (p.epi.relu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = fmaxf(x[e], 0.f);
              }
              if (p.epi.res_h && row < p.M) {
                const uint4 r = __ldg(reinterpret_cast<const uint4*>(p.epi.res_h + row * p.epi.ldr + col));
                const float2 r0 = unpack_half2(r.x), r1 = unpack_half2(r.y), r2 = unpack_half2(r.z), r3 = unpack_half2(r.w);
                x[0] += r0.x; x[1] += r0.y; x[2] += r1.x; x[3] += r1.y;
                x[4] += r2.x; x[5] += r2.y; x[6] += r3.x; x[7] += r3.y;
              }
            }
            const uint4 o = make_uint4(pack_half2(x[0], x[1]), pack_half2(x[2], x[3]), pack_half2(x[4], x[5]), pack_half2(x[6], x[7]));
            *reinterpret_cast<uint4*>(box + lane * 128 + ((q ^ (lane & 7)) << 4)) = o;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&mC, box, colb, m0 + lg * 32);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    } else
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++lt) {
      const int m0 = (t / tiles_n) * G::BM, n0 = (t % tiles_n) * BN;
      const uint32_t as = lt & 1;
      mbar_wait(&tfull[as], (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + as * BN + (uint32_t(lg * 32) << 16);
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= p.N) break;
        uint32_t v[32];
        tmem_ld32(tacc + uint32_t(c0), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 32 >= BN || n0 + c0 + 32 >= p.N) {
          // last TMEM read of this tile: hand the accumulator back to the MMA warp before the stores
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[as])) : "memory");
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(tile + lane * 36 + q * 4) =
              make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        __syncwarp();
        const int col = n0 + c0 + cq * 4;
        if (col < p.N) {   // N is a multiple of 4 (checked on the host)
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.epi.bias) bv = __ldg(reinterpret_cast<const float4*>(p.epi.bias + col));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + rq;
            const long long row = (long long)m0 + lg * 32 + r;
            if (row < p.M) {
              float4 x = *reinterpret_cast<const float4*>(tile + r * 36 + cq * 4);
              x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
              if (p.epi.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
              if (p.epi.res) {
                float4 rr = __ldg(reinterpret_cast<const float4*>(p.epi.res + row * p.epi.ldr + col));
                x.x += rr.x; x.y += rr.y; x.z += rr.z; x.w += rr.w;
              } else if (p.epi.res_h) {
                const uint2 rr = __ldg(reinterpret_cast<const uint2*>(p.epi.res_h + row * p.epi.ldr + col));
                const float2 r0 = unpack_half2(rr.x), r1 = unpack_half2(rr.y);
                x.x += r0.x; x.y += r0.y; x.z += r1.x; x.w += r1.y;
              }
              act_store4(p.C, row, col, x);
            }
          }
        }
        __syncwarp();
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G::TMEM_COLS) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;   // idempotent one-time lookup (immutable afterwards)
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

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

inline bool tma_store_ok(const Act& C) {
  return C.h && !C.f && !C.hi && !C.lo && C.ld % 8 == 0 && (reinterpret_cast<uintptr_t>(C.h) & 15) == 0;
}

template <int KIND, int BN, int STAGES>
inline int launch(Ctx& c, const void* A0, const void* A1, long long lda, long long M, const void* W0, const void* W1, int N, int K,
                  const Epi& epi, const Act& C) {
  typedef Cfg<KIND, BN, STAGES> G;
  const CUtensorMapDataType dt = KIND == KIND_TF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : (KIND == KIND_BF16X3 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  CUtensorMap mA0, mA1, mW0, mW1, mC;
  QAT_TRY(make_map(&mA0, A0, dt, M, K, lda, G::BM));
  QAT_TRY(make_map(&mW0, W0, dt, N, K, K, BN));
  if (G::PLANES == 2) {
    QAT_TRY(make_map(&mA1, A1, dt, M, K, lda, G::BM));
    QAT_TRY(make_map(&mW1, W1, dt, N, K, K, BN));
  } else {
    mA1 = mA0;
    mW1 = mW0;
  }
  GemmP p;
  p.M = (int)M; p.N = N; p.K = K; p.epi = epi; p.C = C;
  p.tma_store = (tma_store_ok(C) && N % 8 == 0 && (epi.res == nullptr) && (epi.res_h == nullptr || epi.ldr % 8 == 0)) ? 1 : 0;
  const char* no_tma = getenv("QAT_GEMM_NO_TMA_STORE");     // A/B switch for profiling
  if (no_tma && no_tma[0] == '1') p.tma_store = 0;
  if (p.tma_store) QAT_TRY(make_map(&mC, C.h, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, M, N, C.ld, 32));
  else mC = mA0;
  auto kern = gemm_tc_kernel<KIND, BN, STAGES>;
  QAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
  const long long ntiles = (long long)((N + BN - 1) / BN) * ((M + G::BM - 1) / G::BM);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
  kern<<<grid, 192, G::SMEM, c.stream>>>(mA0, mA1, mW0, mW1, mC, p);
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

inline bool wide_tile(int N) {
  const char* e = getenv("QAT_GEMM_BN");
  if (e) return atoi(e) == 256;
  return N % 256 == 0;
}

}  // namespace tc

inline int gemm_tc_tf32(Ctx& c, const float* A, int lda, long long M, const float* W, int N, int K, const Epi& epi, Act C) {
  if (M == 0 || N == 0) return QAT_OK;
  QAT_TRY(tc::check_common(M, N, K, lda, epi, C, 4));
  QAT_REQUIRE((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) % 16 == 0, "gemm_tc: operands must be 16-byte aligned");
  if (c.dry) return QAT_OK;
  if (tc::wide_tile(N)) return tc::launch<tc::KIND_TF32, 256, 4>(c, A, nullptr, lda, M, W, nullptr, N, K, epi, C);
  return tc::launch<tc::KIND_TF32, 128, 6>(c, A, nullptr, lda, M, W, nullptr, N, K, epi, C);
}

inline int gemm_tc_bf16x3(Ctx& c, const bf16* Ahi, const bf16* Alo, int lda, long long M, const bf16* Whi, const bf16* Wlo, int N, int K,
                          const Epi& epi, Act C) {
  if (M == 0 || N == 0) return QAT_OK;
  QAT_TRY(tc::check_common(M, N, K, lda, epi, C, 2));
  if (c.dry) return QAT_OK;
  QAT_REQUIRE((reinterpret_cast<uintptr_t>(Ahi) | reinterpret_cast<uintptr_t>(Alo) | reinterpret_cast<uintptr_t>(Whi) |
               reinterpret_cast<uintptr_t>(Wlo)) % 16 == 0, "gemm_tc: operands must be 16-byte aligned");
  if (tc::wide_tile(N)) return tc::launch<tc::KIND_BF16X3, 256, 2>(c, Ahi, Alo, lda, M, Whi, Wlo, N, K, epi, C);
  return tc::launch<tc::KIND_BF16X3, 128, 3>(c, Ahi, Alo, lda, M, Whi, Wlo, N, K, epi, C);
}

inline int gemm_tc_f16(Ctx& c, const __half* A, int lda, long long M, const __half* W, int N, int K, const Epi& epi, Act C) {
  if (M == 0 || N == 0) return QAT_OK;
  QAT_TRY(tc::check_common(M, N, K, lda, epi, C, 2));
  if (c.dry) return QAT_OK;
  QAT_REQUIRE((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) % 16 == 0, "gemm_tc: operands must be 16-byte aligned");
  if (tc::wide_tile(N)) return tc::launch<tc::KIND_F16, 256, 4>(c, A, nullptr, lda, M, W, nullptr, N, K, epi, C);
  return tc::launch<tc::KIND_F16, 128, 6>(c, A, nullptr, lda, M, W, nullptr, N, K, epi, C);
}

}  // namespace qat
