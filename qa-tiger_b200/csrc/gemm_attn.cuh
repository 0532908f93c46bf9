// PatchSelecter self-attention (modules.py:688) as ONE kernel: the packed in-projection [q | k | v] = p W_in^T + b on the
// tcgen05 tensor cores, and softmax(q k^T / 8) v over the P (<= 16) tokens of every frame in the epilogue, so the
// [rows, 3D] projection never reaches HBM (it is 0.66 GB written + 0.66 GB read back per 256-question batch otherwise).
//
// Tile = FPT whole frames (FPT = 128 / P; 9 frames = 126 rows for P = 14) x one head: the B operand of the tile is the
// head's 64 q rows, 64 k rows and 64 v rows of W_in (three TMA boxes -> one 192-row K-major tile), the accumulator is
// 128 x 192 fp32 in TMEM (two of them, so the epilogue of tile i overlaps the main loop of tile i+1).
//   warp 0      TMA producer (A: fp16 activations, W: fp16 weights; 128B swizzle, 4-stage ring)
//   warp 1      TMEM allocator + tcgen05.mma issuer (kind::f16, M = 128, N = 192)
//   warps 2..9  epilogue: TMEM -> +bias -> fp16 [126][192] tile in shared memory -> (named barrier) -> one warp per
//               frame: mma.sync m16n8k16 attention out of that tile (ldmatrix, softmax in the accumulator fragments,
//               P re-used as the A fragment) -> O rows in a 128B-swizzled box -> (named barrier) -> one TMA store.
//   warps 10,11 attention only: with 9 frames per tile, ten attention warps finish in one round instead of two
#pragma once
#include "attn_core.cuh"
#include "gemm_tc.cuh"

namespace qat {
namespace tc {

struct GemmAttnP {
  int M;            // rows = nframes * P
  int nframes, P, H, K;
  const float* bias;   // [3 * H * 64]  (cross form: [2 * H * 64], k then v)
  float scale;
  const __half* q = nullptr;   // cross form: projected queries [nframes, 2, H * 64]
  __half* o = nullptr;         // cross form: output [nframes * 2, ldo]
  int ldo = 0;
  int flags = 0;               // A-resident form, QAT_PROFILE_BUILD only (per-stage cost model): 4 no attention, 8 no drain, 16 no MMAs, 32 no loads
};

constexpr int GA_A_BYTES = 128 * 128;                  // 128 rows x 64 fp16
constexpr int GA_QKV_ROWS = 126;                       // rows of the tile kept for the attention (9 frames x 14 tokens): 128 rows would
                                                       // push the ring + the two tile buffers past the 227 KB of shared memory

// ---- both attentions of the block with a PIPELINED epilogue ----------------------------------------------------------------
// CROSS = false: the self-attention above.  CROSS = true: the cross-attention of the same block (modules.py:692-698): the
// contraction is the packed K | V projection of the frame's tokens (B tile = the head's 64 k rows and 64 v rows, N = 128), the two
// queries of the frame ([video; audio], already projected) come from global memory, and the epilogue runs softmax(q k^T / 8) v for
// those two rows, so the [rows, 2D] K | V projection (0.44 GB written + read back per 256-question batch) never reaches HBM either.
// The epilogue is split into two warp groups that work on different tiles at the same time: four drain warps (one per TMEM lane
// quarter) turn accumulator i+1 into the fp16 [126][(q |) k | v] tile while six attention warps run the attention of tile i out of
// the other buffer; the output rows leave as plain 32-bit stores.  Cross form: the N = 128 tile's MMA time is 1.8 us; with a serial
// drain -> barrier -> attention -> barrier epilogue it ran at 3.1 us per tile, with this one at 2.8 us -- an N = 128 tile moves
// 256 KB of operands per 16.8 MFLOP (ncu: 14 TB/s of L2 -> SM traffic, tensor pipe 55 % active), so operand delivery is next.
// Shared memory: STAGES x (A 16 KB + 8 KB per W box) + two [126][BN + 8] fp16 tiles.
// A CTA-pair form (cta_group::2, the pair owning two row tiles of one head and splitting the head's weight rows) was built and
// measured in round 2: bit-identical, 30 % less L2 -> SM traffic, 1 % SLOWER (the pair couples two drain / attention pipelines to
// one accumulator hand-back); dropped in favour of the A-resident kernels of gemm_attn_ts.cuh, which these kernels back up for
// the shapes those do not take (K != 512, odd head counts).
template <bool CROSS>
struct GxCfg {
  static constexpr int NW = CROSS ? 2 : 3;                 // 64-row blocks of the head's weights: (q,) k, v
  static constexpr int BN = 64 * NW;
  static constexpr int WP = BN / 32;                       // 32-row weight pieces staged per k-block
  static constexpr int STAGES = CROSS ? 4 : 3;
  static constexpr int STAGE_BYTES = GA_A_BYTES + WP * 4096;
  static constexpr int STRIDE = (BN + 8) * 2;              // bytes per tile row: conflict-free ldmatrix / 16-byte stores
  static constexpr int TILE_BYTES = ((GA_QKV_ROWS * STRIDE + 1023) / 1024) * 1024;
  static constexpr int ACC_COLS = CROSS ? 128 : 256;       // TMEM columns per accumulator (two of them)
  static constexpr int BAR_BYTES = (2 * STAGES + 8) * 8 + 16;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 2 * TILE_BYTES + BAR_BYTES + 1024;
  static_assert(SMEM <= 232448, "gemm_attn_pipe: shared-memory budget");
};

// softmax(q k^T * scale) v of ONE (frame, head) out of the fp16 [(q |) k | v] tile rows at shared-memory address `fb` (row pitch
// STRIDE bytes), on mma.sync m16n8k16: ldmatrix fragments, softmax in the accumulator fragments, P re-used as the A fragment of
// the P V product, output rows as plain 32-bit stores.  Cross form: the two query rows come from global memory and are fetched
// before the tile is awaited (`kvfull`, parity `ph`; `waited` records that this warp has passed the barrier for this tile).
template <bool CROSS, int STRIDE>
__device__ __forceinline__ void ga_attend_frame(uint32_t fb, const GemmAttnP& p, long long frame, int h, uint64_t* kvfull, uint32_t ph, bool& waited) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3, m8 = lane >> 3, r8 = lane & 7;
  const int P = p.P, D = p.H * 64;
  constexpr int KOFF = CROSS ? 0 : 64, VOFF = KOFF + 64;                  // tile columns of k and v (q: [0,64) in the self form)
  // A fragments of the 16 x 64 query tile.  Cross form: rows 0, 1 are the frame's queries, fetched from global memory
  // before the tile is awaited, the other rows are zero.  Self form: the frame's own q rows, out of the tile.
  uint32_t qa[4][4];
  if constexpr (CROSS) {
    const __half* qrow = p.q + (frame * 2 + g) * D + h * 64 + 2 * tq;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      qa[ks][0] = g < 2 ? __ldg(reinterpret_cast<const uint32_t*>(qrow + ks * 16)) : 0u;
      qa[ks][2] = g < 2 ? __ldg(reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8)) : 0u;
      qa[ks][1] = 0u;
      qa[ks][3] = 0u;
    }
  }
  if (!waited) {
    mbar_wait(kvfull, ph);
    waited = true;
  }
  if constexpr (!CROSS) {
    const int qrow = min(r8 + (m8 & 1) * 8, P - 1), qcol = (m8 >> 1) * 8;
    const uint32_t qaddr = fb + qrow * STRIDE + qcol * 2;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], qaddr + ks * 32);
  }
  float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  {
    const int krow = min(r8 + (m8 >> 1) * 8, P - 1), kcol = (m8 & 1) * 8;
    const uint32_t ka = fb + krow * STRIDE + (KOFF + kcol) * 2;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t b[4];
      ldsm_x4(b, ka + ks * 32);
      mma_f16(s[0], qa[ks], b[0], b[1]);
      mma_f16(s[1], qa[ks], b[2], b[3]);
    }
  }
  // rows g (and, self form, g + 8) of the score tile: masked softmax over the P keys
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool ok = nt * 8 + 2 * tq + e < P;
      s[nt][e] = ok ? s[nt][e] * p.scale : -INFINITY;
      s[nt][2 + e] = ok ? s[nt][2 + e] * p.scale : -INFINITY;
      mx0 = fmaxf(mx0, s[nt][e]);
      mx1 = fmaxf(mx1, s[nt][2 + e]);
    }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      s[nt][e] = expf(s[nt][e] - mx0);
      s[nt][2 + e] = expf(s[nt][2 + e] - mx1);
      s0 += s[nt][e];
      s1 += s[nt][2 + e];
    }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
  s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float i0 = 1.f / s0, i1 = 1.f / s1;
  uint32_t pa[4];
  pa[0] = pack_half2(s[0][0] * i0, s[0][1] * i0);
  pa[1] = CROSS ? 0u : pack_half2(s[0][2] * i1, s[0][3] * i1);
  pa[2] = pack_half2(s[1][0] * i0, s[1][1] * i0);
  pa[3] = CROSS ? 0u : pack_half2(s[1][2] * i1, s[1][3] * i1);
  float o[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;
  {
    const int vrow = min(r8 + (m8 & 1) * 8, P - 1), vcol = (m8 >> 1) * 8;
    const uint32_t va = fb + vrow * STRIDE + (VOFF + vcol) * 2;
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldsm_x4_t(b, va + dp * 32);
      mma_f16(o[2 * dp], pa, b[0], b[1]);
      mma_f16(o[2 * dp + 1], pa, b[2], b[3]);
    }
  }
  // output rows: cross form (frame, query g < 2); self form (frame, token g) and (frame, token g + 8)
  const int nq = CROSS ? 2 : P;
  if (g < nq) {
    __half* orow = p.o + (frame * nq + g) * p.ldo + h * 64 + 2 * tq;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_half2(o[dt][0], o[dt][1]);
  }
  if (!CROSS && g + 8 < nq) {
    __half* orow = p.o + (frame * nq + g + 8) * p.ldo + h * 64 + 2 * tq;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) *reinterpret_cast<uint32_t*>(orow + dt * 8) = pack_half2(o[dt][2], o[dt][3]);
  }
}

template <bool CROSS>
__global__ void __launch_bounds__(384, 1)
gemm_attn_pipe_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mW, GemmAttnP p) {
  typedef GxCfg<CROSS> G;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* tiles = smem;
  unsigned char* sT = smem + G::STAGES * G::STAGE_BYTES;                   // two tiles
  uint64_t* full = reinterpret_cast<uint64_t*>(sT + 2 * G::TILE_BYTES);
  uint64_t* empty = full + G::STAGES;
  uint64_t* tfull = empty + G::STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* kvfull = tempty + 2;
  uint64_t* kvempty = kvfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kvempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = p.P, H = p.H, D = H * 64;
  const int FPT = 128 / P, RPT = FPT * P;                                 // frames / rows per tile
  const int nkb = p.K / 64;
  const int tiles_m = (p.nframes + FPT - 1) / FPT;
  const int ntiles = tiles_m * H;
  const int unit0 = blockIdx.x, nunits = gridDim.x;                       // persistent walk

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mW)) : "memory");
    for (int s = 0; s < G::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);          // the four drain warps
      mbar_init(&kvfull[a], 4);          // the four drain warps
      mbar_init(&kvempty[a], 6);         // the six attention warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * G::ACC_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    uint32_t it = 0;
    for (int t = unit0; t < ntiles; t += nunits) {
      const int m0 = (t / H) * RPT, h = t % H;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % G::STAGES;
        mbar_wait(&empty[s], ((it / G::STAGES) & 1) ^ 1);
        unsigned char* st = tiles + s * G::STAGE_BYTES;
        const int k0 = kb * 64;
        // weight pieces of 32 rows: piece w of the head's [BN, K] block = rows (w / 2) * D + h * 64 + (w % 2) * 32 of W
        // ((q,) k, v blocks are D rows apart)
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[s], G::STAGE_BYTES);
          tma_load_2d(st, &mA, &full[s], k0, m0);
#pragma unroll
          for (int w = 0; w < G::WP; ++w) tma_load_2d(st + GA_A_BYTES + w * 4096, &mW, &full[s], k0, (w >> 1) * D + h * 64 + (w & 1) * 32);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = umma_idesc(0, 128, G::BN);
      uint32_t it = 0, lt = 0;
      for (int t = unit0; t < ntiles; t += nunits, ++lt) {
        const uint32_t as = lt & 1;
        mbar_wait(&tempty[as], ((lt >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + as * G::ACC_COLS;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % G::STAGES;
          mbar_wait(&full[s], (it / G::STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a0 = smem_u32(tiles + s * G::STAGE_BYTES);
          const uint64_t dA = umma_desc(a0), dW = umma_desc(a0 + GA_A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma<KIND_F16>(tacc, dA + uint64_t(k * 2), dW + uint64_t(k * 2), idesc, (kb | k) != 0);
            umma_commit(&empty[s]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&tfull[as]);
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ===================== drain warps 2..5: accumulator -> (+bias) -> fp16 [126][(q |) k | v] tile =====================
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    uint32_t lt = 0;
    const uint32_t tempty_a0 = smem_u32(&tempty[0]), tempty_a1 = smem_u32(&tempty[1]);
    for (int t = unit0; t < ntiles; t += nunits, ++lt) {
      const int h = t % H;
      const uint32_t as = lt & 1, ph = (lt >> 1) & 1;
      mbar_wait(&tfull[as], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mbar_wait(&kvempty[as], ph ^ 1);                                  // the attention warps are done with this buffer
      const uint32_t tacc = tmem_base + as * G::ACC_COLS + (uint32_t(lg * 32) << 16);
      const uint32_t st = smem_u32(sT + as * G::TILE_BYTES);
#pragma unroll 1
      for (int c = 0; c < G::BN / 32; ++c) {
        const int acol = c * 32;                                        // blocks of 64: (q,) k, v
        uint32_t v[32];
        tmem_ld32(tacc + uint32_t(acol), v);
        const float* bptr = p.bias + (acol >> 6) * D + h * 64 + (acol & 63);
        float4 bb[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) bb[q] = __ldg(reinterpret_cast<const float4*>(bptr) + q);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c == G::BN / 32 - 1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(as ? tempty_a1 : tempty_a0) : "memory");
        }
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
          if (row >= RPT) break;                                        // rows beyond the tile's frames belong to the next tile
          const float4 b0 = bb[2 * q], b1 = bb[2 * q + 1];
          sts128(st + row * G::STRIDE + acol * 2 + q * 16,
                 pack_half2(__uint_as_float(v[8 * q]) + b0.x, __uint_as_float(v[8 * q + 1]) + b0.y),
                 pack_half2(__uint_as_float(v[8 * q + 2]) + b0.z, __uint_as_float(v[8 * q + 3]) + b0.w),
                 pack_half2(__uint_as_float(v[8 * q + 4]) + b1.x, __uint_as_float(v[8 * q + 5]) + b1.y),
                 pack_half2(__uint_as_float(v[8 * q + 6]) + b1.z, __uint_as_float(v[8 * q + 7]) + b1.w));
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&kvfull[as])) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ===================== attention warps 6..11: one (frame, head) at a time on mma.sync =====================
    const int aw = warp - 6;
    uint32_t lt = 0;
    for (int t = unit0; t < ntiles; t += nunits, ++lt) {
      const int mt = t / H, h = t % H;
      const uint32_t as = lt & 1, ph = (lt >> 1) & 1;
      const uint32_t st = smem_u32(sT + as * G::TILE_BYTES);
      bool waited = false;
      for (int f = aw; f < FPT; f += 6) {
        const long long frame = (long long)mt * FPT + f;
        if (frame >= p.nframes) break;
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * G::ACC_COLS) : "memory");
  }
}

template <bool CROSS>
inline int gemm_attn_pipe_launch(Ctx& c, const CUtensorMap& mA, const CUtensorMap& mW, const GemmAttnP& p, long long row_tiles) {
  typedef GxCfg<CROSS> G;
  auto kern = gemm_attn_pipe_kernel<CROSS>;
  QAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long ntiles = row_tiles * p.H;
  kern<<<(unsigned)(ntiles < sms ? ntiles : sms), 384, G::SMEM, c.stream>>>(mA, mW, p);
  return check_launch(CROSS ? "gemm_attn_pipe_kernel<cross>" : "gemm_attn_pipe_kernel<self>");
}

}  // namespace tc

}  // namespace qat
