// PatchSelecter fused projection + attention kernels, A-resident form (modules.py:688, 692-698).
//
// The single-CTA kernels of gemm_attn.cuh re-stage the 128-row activation tile once per head and feed both MMA operands from
// shared memory: per head they push 320 KB of TMA writes + 320 KB of tensor-core reads + the epilogue's tile traffic through the
// SM's shared-memory pipe (measured 117 - 127 B/clk: that pipe, not the tensor pipe, paces them -- tensor pipe 49 % after the
// epilogue's mma.sync share is taken out).  Here the A operand lives in TENSOR MEMORY (tcgen05.mma with [a_tmem]): a row tile
// [128, K <= 512] of fp16 activations is written ONCE into 256 TMEM columns (two fp16 per 32-bit column, lane = row) and reused by
// every head and every (q,) k, v block of the row tile; shared memory only carries the weight rows (8 KB per 64 x 64 block) and the
// attention tiles.  The 64-column blocks (q,) k, v of the heads of a row tile form one sequence that is contracted two blocks at
// a time (N = 128: q k | v q' | k' v' for a head pair of the self form, k | v for a head of the cross form) into two 128-column
// TMEM slots, so A (256 columns) + accumulators (2 x 128) fill the 512 columns and the drain of pair i overlaps the MMAs of pair
// i+1.  (A first version with N = 64 per MMA was issue-bound: ~110 SASS instructions of the single issuing warp per 256 tensor
// clocks; ncu: tensor pipe 46 %.)
//
//   warp 0        TMA producer of the weight blocks (stage = one k-block of a block pair: two 64-row boxes = 16 KB, 128B swizzle)
//   warp 1        TMEM allocator + tcgen05.mma issuer (kind::f16, M = 128, N = 128, A from TMEM, B from shared memory)
//   warps 4..7    A loaders: TMA the row tile's k-blocks into a two-deep staging ring (they issue their own loads), read them back
//                 row-per-lane (the 128B swizzle makes the 16-byte reads conflict-free) and tcgen05.st them into the A columns
//   warps 8..11   drain: accumulator slot -> +bias -> fp16 [126][(q |) k | v] tile in shared memory (one per TMEM lane quarter)
//   warps 2,3,12..18  attention: one (frame, head) at a time on mma.sync out of the finished tile (ga_attend_frame); nine warps,
//                 so the nine frames of a P = 14 tile take one round per head
// Work split: the (row tile, head) -- self form: (row tile, head pair) -- units in row-tile-major order are cut into gridDim.x equal contiguous ranges, so every CTA
// reloads A at most once more than the minimum and no CTA idles in a tail round.
#pragma once
#include <type_traits>

#include "gemm_attn.cuh"

namespace qat {
namespace tc {

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31])
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16, one CTA
__device__ __forceinline__ void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void lds128(uint32_t saddr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(saddr) : "memory");
}

#ifdef QAT_PROFILE_BUILD
#define GA_DBG(p, bit) ((p).flags & (bit))
#else
#define GA_DBG(p, bit) 0
#endif

template <bool CROSS>
struct GtCfg {
  static constexpr int NB = CROSS ? 2 : 3;                 // 64-column blocks of a head: (q,) k, v
  static constexpr int BN = 64 * NB;
  static constexpr int UG = CROSS ? 1 : 3;                 // block pairs per work unit: one head (k | v) / two heads (q k | v q' | k' v')
  static constexpr int STAGE = 16384;                      // one k-block of a block pair (two [64 x 64] fp16 weight boxes) or of the row tile ([128 x 64])
  static constexpr int NKB = 8;                            // k-blocks of 64: K = 512
  static constexpr int NST = NKB;                          // ring depth = k-blocks per group, so the stage of a k-block is the k-block: every
                                                           // barrier / descriptor offset of the unrolled issue loop is a compile-time constant
  static constexpr int NAW = 9;                            // attention warps
  static constexpr int THREADS = (10 + NAW) * 32;
  static constexpr int STRIDE = (BN + 8) * 2;              // bytes per tile row: conflict-free ldmatrix / 16-byte stores
  static constexpr int TILE_BYTES = GA_QKV_ROWS * STRIDE;
  static constexpr int NBAR = 2 * NST + 8 + 8 + 2 + 2 + 2 + 2;
  static constexpr int BAR_BYTES = NBAR * 8 + 16;
  static constexpr int SMEM = NST * STAGE + 2 * TILE_BYTES + BAR_BYTES;   // no alignment slack: the kernel has no static shared memory,
                                                                          // so the dynamic window starts 1024-byte aligned (checked, traps otherwise)
  static constexpr int A_COL = 0, ACC_COL = 256;           // TMEM: A [128 x 512] fp16 in 256 columns, two 128-column accumulator slots
  static_assert(SMEM <= 232448, "gemm_attn_ts: shared-memory budget");
};

template <bool CROSS>
__global__ void __launch_bounds__(GtCfg<CROSS>::THREADS, 1)
gemm_attn_ts_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mW, GemmAttnP p) {
  typedef GtCfg<CROSS> G;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* sR = smem_dyn;                                            // the ring (128B-swizzled boxes need 1024-byte alignment)
  if ((smem_u32(smem_dyn) & 1023u) != 0u) __trap();
  unsigned char* sT = sR + G::NST * G::STAGE;                              // two attention tiles
  uint64_t* full = reinterpret_cast<uint64_t*>(sT + 2 * G::TILE_BYTES);
  uint64_t* empty = full + G::NST;
  uint64_t* aready = empty + G::NST;      // [8] k-block of the row tile sits in TMEM (the four loader warps)
  uint64_t* afree = aready + 8;           // [8] the last MMA of the row tile that reads the k-block has completed: its A columns may be rewritten
  uint64_t* tfull = afree + 8;            // [2] accumulator slot complete
  uint64_t* tempty = tfull + 2;           // [2] slot drained (the four drain warps)
  uint64_t* kvfull = tempty + 2;          // [2] attention tile written (the four drain warps)
  uint64_t* kvempty = kvfull + 2;         // [2] attention tile consumed (the attention warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kvempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = p.P, H = p.H, D = H * 64;
  const int FPT = 128 / P, RPT = FPT * P;                                 // frames / rows per tile
  constexpr int nkb = G::NKB;
  const int NPT = H * G::NB / 2;                                          // block pairs per row tile
  const int tiles_m = (p.nframes + FPT - 1) / FPT;
  const long long U = (long long)tiles_m * (NPT / G::UG);                 // work units, row-tile major
  const int g0 = (int)(U * blockIdx.x / gridDim.x) * G::UG, g1 = (int)(U * (blockIdx.x + 1) / gridDim.x) * G::UG;   // this CTA's block pairs
  // Every role walks the block pairs gp = g0 .. g1-1 and the SAME sequence of stage GROUPS (one group = the ring once round, stage
  // kb = k-block kb): at the first pair of a row tile (gp == g0 or gp % NPT == 0) the row tile's group (consumed by the loader
  // warps), then one weight group per block pair (consumed by the MMA issuer).  Barrier parities follow the group count.

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mW)) : "memory");
    for (int s = 0; s < G::NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int k = 0; k < 8; ++k) {
      mbar_init(&aready[k], 4);
      mbar_init(&afree[k], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
      mbar_init(&kvfull[a], 4);
      mbar_init(&kvempty[a], G::NAW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer: row-tile k-blocks and weight k-blocks through one ring =====================
    uint32_t grp = 0;
    for (int gp = g0; gp < g1; ++gp) {
      const int lp = gp % NPT;
      if (gp == g0 || lp == 0) {
        const int m0 = (gp / NPT) * RPT;
#pragma unroll
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[kb], (grp & 1) ^ 1);
          if (elect_one()) {
            if (GA_DBG(p, 32)) {
              asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[kb])) : "memory");
            } else {
              mbar_arrive_expect_tx(&full[kb], G::STAGE);
              tma_load_2d(sR + kb * G::STAGE, &mA, &full[kb], kb * 64, m0);
            }
          }
          __syncwarp();
        }
        ++grp;
      }
      const int b0 = 2 * lp;                                              // blocks b0, b0 + 1 of the row tile: block = head * NB + (q,) k, v
      const int w0 = (b0 % G::NB) * D + (b0 / G::NB) * 64, w1 = ((b0 + 1) % G::NB) * D + ((b0 + 1) / G::NB) * 64;   // their first rows in W
#pragma unroll
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[kb], (grp & 1) ^ 1);
        unsigned char* st = sR + kb * G::STAGE;
        if (elect_one()) {
          if (GA_DBG(p, 32)) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[kb])) : "memory");
          } else {
            mbar_arrive_expect_tx(&full[kb], G::STAGE);
            tma_load_2d(st, &mW, &full[kb], kb * 64, w0);
            tma_load_2d(st + 8192, &mW, &full[kb], kb * 64, w1);
          }
        }
        __syncwarp();
      }
      ++grp;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one N = 128 accumulator (a block pair) at a time =====================
    // The single issuing warp is the pacing item of this kernel (a stage is only 4 x 64 tensor clocks): the k loop is fully
    // unrolled with constant barrier / descriptor / TMEM offsets, and the first pair of a row tile (which also polls the
    // loaders' barriers) runs its own instance, so a steady-state step is two barrier polls, eight UTCHMMA and the commits.
    constexpr uint32_t idesc = umma_idesc(0, 128, 128);
    const uint64_t d0 = umma_desc(smem_u32(sR));
    const uint32_t ta = tmem_base + G::A_COL;
    uint32_t pc = 0, rtc = 0, grp = 0;
    auto issue_pair = [&](auto first_tag, uint32_t tacc, uint64_t* tfull_slot, bool last) {
      constexpr bool FIRST = decltype(first_tag)::value;
      constexpr int KS = FIRST ? 1 : 2;                                   // k-blocks per elected issue region
      const uint32_t wph = grp & 1, aph = rtc & 1;
#pragma unroll
      for (int kb0 = 0; kb0 < nkb; kb0 += KS) {
#pragma unroll
        for (int j = 0; j < KS; ++j) {
          if (FIRST) mbar_wait(&aready[kb0 + j], aph);
          mbar_wait(&full[kb0 + j], wph);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < KS; ++j) {
            const int kb = kb0 + j;
            const uint64_t dW = d0 + uint64_t(kb * (G::STAGE >> 4));
            if (!GA_DBG(p, 16)) {
              umma_ts_f16(tacc, ta + kb * 32, dW, idesc, kb != 0);
              umma_ts_f16(tacc, ta + kb * 32 + 8, dW + 2, idesc, 1);
              umma_ts_f16(tacc, ta + kb * 32 + 16, dW + 4, idesc, 1);
              umma_ts_f16(tacc, ta + kb * 32 + 24, dW + 6, idesc, 1);
            }
            umma_commit(&empty[kb]);
            if (last) umma_commit(&afree[kb]);       // the row tile's last reader of this k-block: the loaders may refill its columns
            if (kb == nkb - 1) umma_commit(tfull_slot);
          }
        }
        __syncwarp();
      }
    };
    for (int gp = g0; gp < g1; ++gp, ++pc) {
      const int lp = gp % NPT;
      const bool first = gp == g0 || lp == 0, last = gp == g1 - 1 || lp == NPT - 1;
      if (first) ++grp;                                                   // the row tile's own group belongs to the loader warps
      const uint32_t slot = pc & 1;
      mbar_wait(&tempty[slot], ((pc >> 1) & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + G::ACC_COL + slot * 128;
      if (first) issue_pair(std::true_type(), tacc, &tfull[slot], last);
      else issue_pair(std::false_type(), tacc, &tfull[slot], last);
      ++grp;
      if (last) ++rtc;
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== A loaders: the row tile's ring stages -> TMEM =====================
    const int q = warp & 3, row = q * 32 + lane;
    uint32_t rtc = 0, grp = 0;
    for (int gp = g0; gp < g1; ++gp) {
      if (gp == g0 || gp % NPT == 0) {
#pragma unroll
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&afree[kb], (rtc & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          mbar_wait(&full[kb], grp & 1);
          const uint32_t rb = smem_u32(sR + kb * G::STAGE) + row * 128;
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) lds128(rb + ((j ^ (row & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          tmem_st32(tmem_base + (uint32_t(q * 32) << 16) + G::A_COL + kb * 32, v);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&aready[kb])) : "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");                  // the four loader warps have read the stage
          if (q == 0 && lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[kb])) : "memory");
        }
        ++rtc;
        ++grp;
      }
      ++grp;                                                              // the pair's weight group belongs to the MMA issuer
    }
  } else if (warp >= 8 && warp < 12) {
    // ===================== drain: accumulator halves -> (+bias) -> fp16 tile columns of the block =====================
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    uint32_t pc = 0, hc = 0;                                            // block pairs / heads this CTA has drained
    uint32_t st = smem_u32(sT);
    for (int gp = g0; gp < g1; ++gp, ++pc) {
      const uint32_t slot = pc & 1;
      const int b0 = 2 * (gp % NPT);
      mbar_wait(&tfull[slot], (pc >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + G::ACC_COL + slot * 128 + (uint32_t(lg * 32) << 16);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int h = (b0 + half) / G::NB, blk = (b0 + half) % G::NB;
        if (blk == 0) {
          mbar_wait(&kvempty[hc & 1], ((hc >> 1) & 1) ^ 1);             // the attention warps are done with this buffer
          st = smem_u32(sT + (hc & 1) * G::TILE_BYTES);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          if (!GA_DBG(p, 8)) tmem_ld32(tacc + uint32_t(half * 64 + c * 32), v);
          const float* bptr = p.bias + blk * D + h * 64 + c * 32;
          float4 bb[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) bb[i] = __ldg(reinterpret_cast<const float4*>(bptr) + i);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (half == 1 && c == 1) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[slot])) : "memory");
          }
          if (row < RPT && !GA_DBG(p, 8)) {                             // rows beyond the tile's frames belong to the next tile
#pragma unroll
            for (uint32_t i = 0; i < 4; ++i) {
              const float4 b0v = bb[2 * i], b1v = bb[2 * i + 1];
              sts128(st + row * G::STRIDE + (blk * 64 + c * 32) * 2 + i * 16,
                     pack_half2(__uint_as_float(v[8 * i]) + b0v.x, __uint_as_float(v[8 * i + 1]) + b0v.y),
                     pack_half2(__uint_as_float(v[8 * i + 2]) + b0v.z, __uint_as_float(v[8 * i + 3]) + b0v.w),
                     pack_half2(__uint_as_float(v[8 * i + 4]) + b1v.x, __uint_as_float(v[8 * i + 5]) + b1v.y),
                     pack_half2(__uint_as_float(v[8 * i + 6]) + b1v.z, __uint_as_float(v[8 * i + 7]) + b1v.w));
            }
          }
        }
        if (blk == G::NB - 1) {
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&kvfull[hc & 1])) : "memory");
          ++hc;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ===================== attention warps 2, 3, 12..: one (frame, head) at a time on mma.sync =====================
    const int aw = warp < 4 ? warp - 2 : warp - 10;
    const int gh0 = 2 * g0 / G::NB, gh1 = 2 * g1 / G::NB;               // heads of this CTA (row-tile major)
    uint32_t lt = 0;
    for (int gh = gh0; gh < gh1; ++gh, ++lt) {
      const int mt = gh / H, h = gh % H;
      const uint32_t as = lt & 1, ph = (lt >> 1) & 1;
      const uint32_t st = smem_u32(sT + as * G::TILE_BYTES);
      bool waited = false;
      for (int f = aw; f < FPT; f += G::NAW) {
        const long long frame = (long long)mt * FPT + f;
        if (frame >= p.nframes || GA_DBG(p, 4)) break;
        ga_attend_frame<CROSS, G::STRIDE>(st + (f * P) * G::STRIDE, p, frame, h, &kvfull[as], ph, waited);
      }
      if (!waited) mbar_wait(&kvfull[as], ph);           // a warp without a frame in this tile still follows the phases
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&kvempty[as])) : "memory");
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <bool CROSS>
inline int gemm_attn_ts_launch(Ctx& c, const CUtensorMap& mA, const CUtensorMap& mW, const GemmAttnP& p, long long row_tiles) {
  typedef GtCfg<CROSS> G;
  auto kern = gemm_attn_ts_kernel<CROSS>;
  QAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long units = row_tiles * (p.H * G::NB / 2 / G::UG);
  const unsigned grid = (unsigned)(units < sms ? units : sms);
  kern<<<grid, G::THREADS, G::SMEM, c.stream>>>(mA, mW, p);
  return check_launch(CROSS ? "gemm_attn_ts_kernel<cross>" : "gemm_attn_ts_kernel<self>");
}

// the A-resident form needs the whole row tile in 256 TMEM columns and whole stage pairs
inline bool gemm_attn_ts_ok(const Ctx& c, int K, int H) {
  return c.opt.v[OPT_GEMM_ATTN_NO_TS] != 1 && K == 512 && H % 2 == 0;
}

}  // namespace tc

// O[rows, D] (fp16) = per-frame softmax((X Wq^T + bq)(X Wk^T + bk)^T / 8)(X Wv^T + bv), X [nframes * P, K] fp16,
// W [3D, K] fp16 = in_proj_weight, bias [3D] fp32
inline int gemm_attn_f16(Ctx& c, const __half* X, int ldx, int nframes, int P, int H, int K, const __half* W, const float* bias, float scale,
                         __half* O, int ldo) {
  QAT_REQUIRE(P >= 1 && P <= 16 && H >= 1 && K % 64 == 0 && bias, "gemm_attn: need P <= 16, K %% 64 == 0 and a bias (P=%d K=%d)", P, K);
  QAT_REQUIRE((128 / P) * P <= tc::GA_QKV_ROWS, "gemm_attn: P=%d fills more than %d rows of a tile", P, tc::GA_QKV_ROWS);
  QAT_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0, "gemm_attn: row pitches must be multiples of 16 bytes");
  if (c.dry || nframes == 0) return QAT_OK;
  QAT_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(O) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
              "gemm_attn: tensors must be 16-byte aligned");
  const int D = H * 64, FPT = 128 / P;
  const long long M = (long long)nframes * P;
  CUtensorMap mA, mW;
  QAT_TRY(tc::make_map(&mA, X, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, M, K, ldx, 128));
  QAT_TRY(tc::make_map(&mW, W, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3 * D, K, K, 32));      // 32-row weight pieces
  tc::GemmAttnP p;
  p.M = (int)M; p.nframes = nframes; p.P = P; p.H = H; p.K = K; p.bias = bias; p.scale = scale;
  p.o = O;
  p.ldo = ldo;
  const long long row_tiles = (nframes + FPT - 1) / FPT;
  if (tc::gemm_attn_ts_ok(c, K, H)) {
    QAT_TRY(tc::make_map(&mW, W, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3 * D, K, K, 64));    // 64-row weight blocks
    p.flags = c.opt.v[OPT_GEMM_ATTN_TS_FLAGS];
    return tc::gemm_attn_ts_launch<false>(c, mA, mW, p, row_tiles);
  }
  return tc::gemm_attn_pipe_launch<false>(c, mA, mW, p, row_tiles);
}

// O[nframes * 2, D] (fp16) = per-frame softmax(Q (X Wk^T + bk)^T / 8)(X Wv^T + bv) for the two projected query rows
// Q [nframes, 2, D] (fp16) of every frame; X [nframes * P, K] fp16, W [2D, K] fp16 = in_proj_weight[D:], bias [2D] fp32
inline int gemm_attn_cross_f16(Ctx& c, const __half* X, int ldx, int nframes, int P, int H, int K, const __half* W, const float* bias,
                               float scale, const __half* Q, __half* O, int ldo) {
  QAT_REQUIRE(P >= 1 && P <= 16 && H >= 1 && K % 64 == 0 && bias, "gemm_attn_cross: need P <= 16, K %% 64 == 0 and a bias (P=%d K=%d)", P, K);
  QAT_REQUIRE((128 / P) * P <= tc::GA_QKV_ROWS, "gemm_attn_cross: P=%d fills more than %d rows of a tile", P, tc::GA_QKV_ROWS);
  QAT_REQUIRE(ldx % 8 == 0 && ldo % 2 == 0, "gemm_attn_cross: row pitches must be multiples of 16 (X) / 4 (O) bytes");
  if (c.dry || nframes == 0) return QAT_OK;
  QAT_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0 &&
                  ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(O)) & 3) == 0,
              "gemm_attn_cross: tensors must be 16-byte (X, W, bias) / 4-byte (Q, O) aligned");
  const int D = H * 64, FPT = 128 / P;
  const long long M = (long long)nframes * P;
  CUtensorMap mA, mW;
  QAT_TRY(tc::make_map(&mA, X, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, M, K, ldx, 128));
  QAT_TRY(tc::make_map(&mW, W, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2 * D, K, K, 32));      // 32-row weight pieces
  tc::GemmAttnP p;
  p.M = (int)M; p.nframes = nframes; p.P = P; p.H = H; p.K = K; p.bias = bias; p.scale = scale;
  p.q = Q; p.o = O; p.ldo = ldo;
  const long long row_tiles = (nframes + FPT - 1) / FPT;
  if (tc::gemm_attn_ts_ok(c, K, H)) {
    QAT_TRY(tc::make_map(&mW, W, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2 * D, K, K, 64));    // 64-row weight blocks
    p.flags = c.opt.v[OPT_GEMM_ATTN_TS_FLAGS];
    return tc::gemm_attn_ts_launch<true>(c, mA, mW, p, row_tiles);
  }
  return tc::gemm_attn_pipe_launch<true>(c, mA, mW, p, row_tiles);
}

}  // namespace qat
