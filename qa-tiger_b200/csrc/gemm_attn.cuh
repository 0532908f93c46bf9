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
// CG = 2: CTA pairs (cluster of two CTAs on one TPC, tcgen05.mma.cta_group::2, M = 256): the pair owns TWO row tiles of one head;
// each CTA stages its own 128 A rows and HALF of the head's weight rows, the leader issues one MMA over both shared memories and
// both TMEMs, and every CTA ends up with the accumulator of ITS row tile x all BN columns -- the drain / attention half of the
// kernel is unchanged, while the weight traffic through L2 and shared memory halves (ncu on the single-CTA form: 12.5 - 14 TB/s
// of L2 -> SM traffic, of which 60 % weight rows re-read by every row tile; tensor pipe 55 - 58 %).
template <bool CROSS, int CG = 1>
struct GxCfg {
  static constexpr int NW = CROSS ? 2 : 3;                 // 64-row blocks of the head's weights: (q,) k, v
  static constexpr int BN = 64 * NW;
  static constexpr int WP = BN / 32 / CG;                  // 32-row weight pieces this CTA stages per k-block
  static constexpr int STAGES = CG == 2 ? (CROSS ? 5 : 4) : (CROSS ? 4 : 3);
  static constexpr int STAGE_BYTES = GA_A_BYTES + WP * 4096;
  static constexpr int STRIDE = (BN + 8) * 2;              // bytes per tile row: conflict-free ldmatrix / 16-byte stores
  static constexpr int TILE_BYTES = ((GA_QKV_ROWS * STRIDE + 1023) / 1024) * 1024;
  static constexpr int ACC_COLS = CROSS ? 128 : 256;       // TMEM columns per accumulator (two of them)
  static constexpr int BAR_BYTES = (2 * STAGES + 8) * 8 + 16;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 2 * TILE_BYTES + BAR_BYTES + 1024;
  static_assert(SMEM <= 232448, "gemm_attn_pipe: shared-memory budget");
};

template <bool CROSS, int CG>
__global__ void __launch_bounds__(384, 1)
gemm_attn_pipe_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mW, GemmAttnP p) {
  typedef GxCfg<CROSS, CG> G;
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
  const int tiles_m = ((p.nframes + FPT - 1) / FPT + CG - 1) / CG;         // (pairs of) row tiles; a tail pair may be half empty
  const int ntiles = tiles_m * H;
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;
  const int unit0 = blockIdx.x / CG, nunits = gridDim.x / CG;             // persistent walk, one unit = one CTA (pair)
  constexpr int KOFF = CROSS ? 0 : 64, VOFF = KOFF + 64;                  // tile columns of k and v (q: [0,64) in the self form)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mW)) : "memory");
    for (int s = 0; s < G::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4 * CG);     // the four drain warps (of both CTAs of a pair)
      mbar_init(&kvfull[a], 4);          // the four drain warps
      mbar_init(&kvempty[a], 6);         // the six attention warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * G::ACC_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * G::ACC_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();        // the peer's barriers are initialised before anything signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    uint32_t it = 0;
    for (int t = unit0; t < ntiles; t += nunits) {
      const int m0 = ((t / H) * CG + (int)crank) * RPT, h = t % H;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % G::STAGES;
        mbar_wait(&empty[s], ((it / G::STAGES) & 1) ^ 1);
        unsigned char* st = tiles + s * G::STAGE_BYTES;
        const int k0 = kb * 64;
        // weight pieces of 32 rows: piece pi of the head's [BN, K] block = rows (pi / 2) * D + h * 64 + (pi % 2) * 32 of W
        // ((q,) k, v blocks are D rows apart); a CTA of a pair stages the pieces of its half of the BN columns
        if (CG == 2) {
          const uint32_t fb = mapa_rank(smem_u32(&full[s]), 0);           // both CTAs' bytes are counted on the leader's barrier
          if (elect_one()) {
            if (crank == 0) mbar_arrive_expect_tx(&full[s], 2 * G::STAGE_BYTES);
            tma_load_2d_pair(st, &mA, fb, k0, m0);
#pragma unroll
            for (int w = 0; w < G::WP; ++w) {
              const int pi = (int)crank * G::WP + w;
              tma_load_2d_pair(st + GA_A_BYTES + w * 4096, &mW, fb, k0, (pi >> 1) * D + h * 64 + (pi & 1) * 32);
            }
          }
        } else if (elect_one()) {
          mbar_arrive_expect_tx(&full[s], G::STAGE_BYTES);
          tma_load_2d(st, &mA, &full[s], k0, m0);
#pragma unroll
          for (int w = 0; w < G::WP; ++w) tma_load_2d(st + GA_A_BYTES + w * 4096, &mW, &full[s], k0, (w >> 1) * D + h * 64 + (w & 1) * 32);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {                                                     // the leader issues for the pair
      constexpr uint32_t idesc = umma_idesc(0, 128 * CG, G::BN);
      uint32_t it = 0, lt = 0;
      for (int t = unit0; t < ntiles; t += nunits, ++lt) {
        const uint32_t as = lt & 1;
        if (CG == 2) mbar_wait_cluster(&tempty[as], ((lt >> 1) & 1) ^ 1);
        else mbar_wait(&tempty[as], ((lt >> 1) & 1) ^ 1);
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
            for (int k = 0; k < 4; ++k) umma<KIND_F16, CG>(tacc, dA + uint64_t(k * 2), dW + uint64_t(k * 2), idesc, (kb | k) != 0);
            if (CG == 2) umma_commit_pair(&empty[s]);
            else umma_commit(&empty[s]);
          }
          __syncwarp();
        }
        if (elect_one()) {
          if (CG == 2) umma_commit_pair(&tfull[as]);
          else umma_commit(&tfull[as]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ===================== drain warps 2..5: accumulator -> (+bias) -> fp16 [126][(q |) k | v] tile =====================
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    uint32_t lt = 0;
    // accumulator hand-back: the leader's barrier (a remote arrival from the peer CTA)
    const uint32_t tempty_a0 = CG == 2 ? mapa_rank(smem_u32(&tempty[0]), 0) : smem_u32(&tempty[0]);
    const uint32_t tempty_a1 = CG == 2 ? mapa_rank(smem_u32(&tempty[1]), 0) : smem_u32(&tempty[1]);
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
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(as ? tempty_a1 : tempty_a0);
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(as ? tempty_a1 : tempty_a0) : "memory");
          }
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
    const int g = lane >> 2, tq = lane & 3, m8 = lane >> 3, r8 = lane & 7;
    uint32_t lt = 0;
    for (int t = unit0; t < ntiles; t += nunits, ++lt) {
      const int mt = (t / H) * CG + (int)crank, h = t % H;
      const uint32_t as = lt & 1, ph = (lt >> 1) & 1;
      const uint32_t st = smem_u32(sT + as * G::TILE_BYTES);
      bool waited = false;
      for (int f = aw; f < FPT; f += 6) {
        const long long frame = (long long)mt * FPT + f;
        if (frame >= p.nframes) break;
        const uint32_t fb = st + (f * P) * G::STRIDE;
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
          mbar_wait(&kvfull[as], ph);
          waited = true;
        }
        if constexpr (!CROSS) {
          const int qrow = min(r8 + (m8 & 1) * 8, P - 1), qcol = (m8 >> 1) * 8;
          const uint32_t qaddr = fb + qrow * G::STRIDE + qcol * 2;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], qaddr + ks * 32);
        }
        float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        {
          const int krow = min(r8 + (m8 >> 1) * 8, P - 1), kcol = (m8 & 1) * 8;
          const uint32_t ka = fb + krow * G::STRIDE + (KOFF + kcol) * 2;
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
          const uint32_t va = fb + vrow * G::STRIDE + (VOFF + vcol) * 2;
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
      if (!waited) mbar_wait(&kvfull[as], ph);           // a warp without a frame in this tile still follows the phases
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&kvempty[as])) : "memory");
    }
  }
  __syncthreads();
  if (CG == 2) cluster_sync_all();        // no CTA of the pair leaves (or frees TMEM) while the other may still signal / be read
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * G::ACC_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * G::ACC_COLS) : "memory");
  }
}

// launch: `cg` = 1 single CTAs, 2 = CTA pairs (cluster dimension 2)
template <bool CROSS, int CG>
inline int gemm_attn_pipe_launch(Ctx& c, const CUtensorMap& mA, const CUtensorMap& mW, const GemmAttnP& p, long long row_tiles) {
  typedef GxCfg<CROSS, CG> G;
  auto kern = gemm_attn_pipe_kernel<CROSS, CG>;
  QAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long ntiles = ((row_tiles + CG - 1) / CG) * p.H;
  const long long units = sms / CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ntiles < units ? ntiles : units) * CG);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = G::SMEM;
  cfg.stream = c.stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = CG == 2 ? 1 : 0;
  QAT_CUDA(cudaLaunchKernelEx(&cfg, kern, mA, mW, p));
  return check_launch(CROSS ? "gemm_attn_pipe_kernel<cross>" : "gemm_attn_pipe_kernel<self>");
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
  const int D = H * 64, FPT = 128 / P, RPT = FPT * P;
  const long long M = (long long)nframes * P;
  CUtensorMap mA, mW;
  QAT_TRY(tc::make_map(&mA, X, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, M, K, ldx, 128));
  QAT_TRY(tc::make_map(&mW, W, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3 * D, K, K, 32));      // 32-row weight pieces
  tc::GemmAttnP p;
  p.M = (int)M; p.nframes = nframes; p.P = P; p.H = H; p.K = K; p.bias = bias; p.scale = scale;
  p.o = O;
  p.ldo = ldo;
  const long long row_tiles = (nframes + FPT - 1) / FPT;
  if (row_tiles >= 2 && c.opt.v[OPT_GEMM_ATTN_CG] == 2) return tc::gemm_attn_pipe_launch<false, 2>(c, mA, mW, p, row_tiles);
  return tc::gemm_attn_pipe_launch<false, 1>(c, mA, mW, p, row_tiles);
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
  if (row_tiles >= 2 && c.opt.v[OPT_GEMM_ATTN_CG] == 2) return tc::gemm_attn_pipe_launch<true, 2>(c, mA, mW, p, row_tiles);
  return tc::gemm_attn_pipe_launch<true, 1>(c, mA, mW, p, row_tiles);
}

}  // namespace qat
