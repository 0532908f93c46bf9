// Softmax(QK^T)V cores for the short sequences of the hot path: lengths 1/2/14/60/77/120,
// head dim 64.  These are <1% of the FLOPs; they read strided views of the projection GEMM outputs (no head split /
// transpose is ever materialised): the 60..128-token cores on warp-level mma.sync with fp16 (hi, lo) operand pairs
// (attn_mma16_kernel), the rest on CUDA cores in fp32.
// Semantics: torch MHA slow path, F.multi_head_attention_forward -- q is scaled by hd^-0.5 BEFORE the
// dot product, softmax = exp(s - max)/sum, no masks (modules.py:217-221,338,562,688,698).
#pragma once
#include <stdlib.h>

#include "qat_common.cuh"

namespace qat {

struct AttnP {
  const float* q;
  const float* k;
  const float* v;
  long long q_sn, q_si;  // float strides: per sequence, per query row
  long long k_sn, k_si;  // per kv-sequence, per key row (k and v share them)
  int k_shift, k_mod;    // kv sequence of query sequence n is (n + k_shift) % k_mod
  Act o;                 // output, row = n*o_rn + i*o_ri, col = h*64 + d
  long long o_rn, o_ri;
  float* attw;           // optional [nseq, Lq, Lk] head-averaged weights (zero-initialised by the caller)
  int nseq, H, Lq, Lk;
  float scale;
  int causal = 0;        // 1: key j attends only to queries i >= j (CLIP text tower; attn_mma16_kernel only)
  // optional: Q / K / V as fp16 (hi, lo) plane pairs written by the in-projection GEMM (same strides, in elements); when
  // set, q / k / v are ignored and the call runs on attn_mma16_kernel<NT16, true>
  const __half* qh = nullptr;
  const __half* ql = nullptr;
  const __half* kh = nullptr;
  const __half* kl = nullptr;
  const __half* vh = nullptr;
  const __half* vl = nullptr;
  // optional additive per-key bias [n_kv_sequences, Lk] (ToMe's proportional attention: log(size), timm.py:94-95); long-sequence
  // core only (attn_long.cuh)
  const float* kbias = nullptr;
};

constexpr int HD = 64;
// floats taken by a [Lk][65] K tile, rounded up so the V tile behind it stays 16-byte aligned
__host__ __device__ inline int ks_floats(int Lk) { return (Lk * 65 + 3) & ~3; }

// ---- one CTA (4 warps) per (sequence, head); Lk <= 128 --------------------------------------------
__global__ void __launch_bounds__(128) attn_block_kernel(AttnP p) {
  extern __shared__ __align__(16) float smem[];
  const int Lk = p.Lk;
  float* Ks = smem;                   // [Lk][65]
  float* Vs = Ks + ks_floats(Lk);     // [Lk][64]
  float* Ps = Vs + Lk * 64;           // [4][128]
  const int n = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int kn = (n + p.k_shift) % p.k_mod;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* kb = p.k + kn * p.k_sn + h * HD;
  const float* vb = p.v + kn * p.k_sn + h * HD;
  for (int f = tid; f < Lk * 16; f += 128) {
    int r = f >> 4, c4 = f & 15;
    float4 kv = __ldg(reinterpret_cast<const float4*>(kb + r * p.k_si) + c4);
    float4 vv = __ldg(reinterpret_cast<const float4*>(vb + r * p.k_si) + c4);
    float* kd = Ks + r * 65 + c4 * 4;
    kd[0] = kv.x; kd[1] = kv.y; kd[2] = kv.z; kd[3] = kv.w;
    *reinterpret_cast<float4*>(Vs + r * 64 + c4 * 4) = vv;
  }
  __syncthreads();
  float* Pw = Ps + w * 128;
  const float invH = 1.f / p.H;
  for (int i = w; i < p.Lq; i += 4) {
    const float* qr = p.q + n * p.q_sn + i * p.q_si + h * HD;
    float q0 = __ldg(qr + lane) * p.scale, q1 = __ldg(qr + 32 + lane) * p.scale;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int d = 0; d < 64; ++d) {
      float qd = __shfl_sync(0xffffffffu, d < 32 ? q0 : q1, d & 31);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        int j = lane + 32 * m;
        if (j < Lk) s[m] = fmaf(qd, Ks[j * 65 + d], s[m]);
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int m = 0; m < 4; ++m)
      if (lane + 32 * m < Lk) mx = fmaxf(mx, s[m]);
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      s[m] = (lane + 32 * m < Lk) ? expf(s[m] - mx) : 0.f;
      sum += s[m];
    }
    sum = warp_sum(sum);
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      int j = lane + 32 * m;
      if (j < Lk) {
        float pj = s[m] / sum;
        Pw[j] = pj;
        if (p.attw) atomicAdd(p.attw + ((long long)n * p.Lq + i) * Lk + j, pj * invH);
      }
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < Lk; ++j) {
      float pj = Pw[j];
      o0 = fmaf(pj, Vs[j * 64 + lane], o0);
      o1 = fmaf(pj, Vs[j * 64 + 32 + lane], o1);
    }
    long long row = n * p.o_rn + i * p.o_ri;
    act_store1(p.o, row, h * HD + lane, o0);
    act_store1(p.o, row, h * HD + 32 + lane, o1);
  }
}

// ---- one warp per (sequence, head); Lk <= 32; a CTA carries WPB consecutive (seq, head) problems ----
template <int WPB>
__global__ void __launch_bounds__(32 * WPB) attn_warp_kernel(AttnP p) {
  extern __shared__ __align__(16) float smem[];
  const int Lk = p.Lk;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long prob = (long long)blockIdx.x * WPB + w;
  if (prob >= (long long)p.nseq * p.H) return;
  const int n = (int)(prob / p.H), h = (int)(prob % p.H);
  const int kn = (n + p.k_shift) % p.k_mod;
  float* Ks = smem + w * (ks_floats(Lk) + Lk * 64);  // [Lk][65]
  float* Vs = Ks + ks_floats(Lk);                    // [Lk][64]
  const float* kb = p.k + kn * p.k_sn + h * HD;
  const float* vb = p.v + kn * p.k_sn + h * HD;
  for (int f = lane; f < Lk * 16; f += 32) {
    int r = f >> 4, c4 = f & 15;
    float4 kv = __ldg(reinterpret_cast<const float4*>(kb + r * p.k_si) + c4);
    float4 vv = __ldg(reinterpret_cast<const float4*>(vb + r * p.k_si) + c4);
    float* kd = Ks + r * 65 + c4 * 4;
    kd[0] = kv.x; kd[1] = kv.y; kd[2] = kv.z; kd[3] = kv.w;
    *reinterpret_cast<float4*>(Vs + r * 64 + c4 * 4) = vv;
  }
  __syncwarp();
  const float invH = 1.f / p.H;
  for (int i = 0; i < p.Lq; ++i) {
    const float* qr = p.q + n * p.q_sn + i * p.q_si + h * HD;
    float q0 = __ldg(qr + lane) * p.scale, q1 = __ldg(qr + 32 + lane) * p.scale;
    float s = 0.f;
    const int j = lane < Lk ? lane : 0;
#pragma unroll 16
    for (int d = 0; d < 64; ++d) {
      float qd = __shfl_sync(0xffffffffu, d < 32 ? q0 : q1, d & 31);
      s = fmaf(qd, Ks[j * 65 + d], s);
    }
    float mx = warp_max(lane < Lk ? s : -INFINITY);
    float e = lane < Lk ? expf(s - mx) : 0.f;
    float sum = warp_sum(e);
    float pj = e / sum;
    if (p.attw && lane < Lk) atomicAdd(p.attw + ((long long)n * p.Lq + i) * Lk + lane, pj * invH);
    float o0 = 0.f, o1 = 0.f;
    for (int jj = 0; jj < Lk; ++jj) {
      float pb = __shfl_sync(0xffffffffu, pj, jj);
      o0 = fmaf(pb, Vs[jj * 64 + lane], o0);
      o1 = fmaf(pb, Vs[jj * 64 + 32 + lane], o1);
    }
    long long row = n * p.o_rn + i * p.o_ri;
    act_store1(p.o, row, h * HD + lane, o0);
    act_store1(p.o, row, h * HD + 32 + lane, o1);
  }
}

// ---- one THREAD per query row; short sequences (Lq, Lk <= 16), e.g. the 14 patch tokens of a frame ----
// A CTA owns one sequence and all its heads: K and V of every head are staged once in shared memory with
// cp.async; a thread keeps its q row, its <=16 scores and its output row in registers, so the softmax
// needs no shuffles and every K/V shared-memory read is a 128-bit broadcast.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
constexpr int ROWS_MAXL = 16;

__global__ void __launch_bounds__(128) attn_rows_kernel(AttnP p) {
  extern __shared__ __align__(16) float smem[];
  const int Lq = p.Lq, Lk = p.Lk, H = p.H;
  const int n = blockIdx.x;
  const int kn = (n + p.k_shift) % p.k_mod;
  const int blk = Lk * HD + 4;                       // +4 floats: heads land on different banks
  float* Ks = smem;                                  // [H][Lk][64] (+pad)
  float* Vs = Ks + H * blk;
  const int tid = threadIdx.x;
  const float* kb = p.k + kn * p.k_sn;
  const float* vb = p.v + kn * p.k_sn;
  for (int f = tid; f < H * Lk * 16; f += blockDim.x) {
    const int c4 = f & 15, hj = f >> 4;
    const int j = hj % Lk, h = hj / Lk;
    cp_async16(Ks + h * blk + j * HD + c4 * 4, kb + j * p.k_si + h * HD + c4 * 4);
    cp_async16(Vs + h * blk + j * HD + c4 * 4, vb + j * p.k_si + h * HD + c4 * 4);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // blockDim >= H*Lq (host check): one row per thread, idle threads only help with the staging above
  const int r = tid;
  const bool active = r < H * Lq;
  const int h = active ? r / Lq : 0, i = active ? r % Lq : 0;
  float q[HD];
  if (active) {
    const float4* qr = reinterpret_cast<const float4*>(p.q + n * p.q_sn + i * p.q_si + h * HD);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float4 t = __ldg(qr + c);
      q[4 * c] = t.x * p.scale; q[4 * c + 1] = t.y * p.scale; q[4 * c + 2] = t.z * p.scale; q[4 * c + 3] = t.w * p.scale;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (active) {
    float sc[ROWS_MAXL];
    const float* kh = Ks + h * blk;
#pragma unroll
    for (int j = 0; j < ROWS_MAXL; ++j) {
      if (j < Lk) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float4* kr = reinterpret_cast<const float4*>(kh + j * HD);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float4 t = kr[c];
          a0 = fmaf(q[4 * c], t.x, a0); a1 = fmaf(q[4 * c + 1], t.y, a1);
          a2 = fmaf(q[4 * c + 2], t.z, a2); a3 = fmaf(q[4 * c + 3], t.w, a3);
        }
        sc[j] = (a0 + a1) + (a2 + a3);
      } else {
        sc[j] = -INFINITY;
      }
    }
    float mx = sc[0];
#pragma unroll
    for (int j = 1; j < ROWS_MAXL; ++j) mx = fmaxf(mx, sc[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < ROWS_MAXL; ++j) {
      sc[j] = j < Lk ? expf(sc[j] - mx) : 0.f;
      sum += sc[j];
    }
    const float inv = 1.f / sum;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    const float* vh = Vs + h * blk;
#pragma unroll
    for (int j = 0; j < ROWS_MAXL; ++j) {
      if (j < Lk) {
        const float pj = sc[j] * inv;
        if (p.attw) atomicAdd(p.attw + ((long long)n * Lq + i) * Lk + j, pj / H);
        const float4* vr = reinterpret_cast<const float4*>(vh + j * HD);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float4 t = vr[c];
          o[4 * c] = fmaf(pj, t.x, o[4 * c]); o[4 * c + 1] = fmaf(pj, t.y, o[4 * c + 1]);
          o[4 * c + 2] = fmaf(pj, t.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(pj, t.w, o[4 * c + 3]);
        }
      }
    }
    const long long row = n * p.o_rn + i * p.o_ri;
#pragma unroll
    for (int c = 0; c < 16; ++c) act_store4(p.o, row, h * HD + 4 * c, make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]));
  }
}

// ---- one WARP per (sequence, head), LQ <= 2 query rows, any Lk: streaming, single pass -------------
// Lane l owns dims (2l, 2l+1): K and V rows are read once, coalesced (256 B per row), scores meet through
// xor-shuffles, the softmax is the online (running max / running sum) form, exact up to fp32 rounding.
template <int LQ>
__global__ void __launch_bounds__(256) attn_fewq_kernel(AttnP p) {
  const int lane = threadIdx.x & 31;
  const long long prob = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (prob >= (long long)p.nseq * p.H) return;
  const int n = (int)(prob / p.H), h = (int)(prob % p.H);
  const int kn = (n + p.k_shift) % p.k_mod;
  const float* kb = p.k + kn * p.k_sn + h * HD + 2 * lane;
  const float* vb = p.v + kn * p.k_sn + h * HD + 2 * lane;
  float2 q[LQ], o[LQ];
  float m[LQ], l[LQ];
#pragma unroll
  for (int i = 0; i < LQ; ++i) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p.q + n * p.q_sn + i * p.q_si + h * HD + 2 * lane));
    q[i] = make_float2(t.x * p.scale, t.y * p.scale);
    o[i] = make_float2(0.f, 0.f);
    m[i] = -INFINITY;
    l[i] = 0.f;
  }
  const int Lk = p.Lk;
  constexpr int UN = LQ == 1 ? 16 : 8;   // independent K/V row loads in flight per batch (latency-bound kernel)
  for (int j0 = 0; j0 < Lk; j0 += UN) {
    float2 kk[UN], vv[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u < Lk ? j0 + u : Lk - 1;
      kk[u] = __ldg(reinterpret_cast<const float2*>(kb + j * p.k_si));
      vv[u] = __ldg(reinterpret_cast<const float2*>(vb + j * p.k_si));
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (j0 + u < Lk) {
#pragma unroll
        for (int i = 0; i < LQ; ++i) {
          float s = warp_sum(fmaf(q[i].x, kk[u].x, q[i].y * kk[u].y));
          const float mn = fmaxf(m[i], s);
          const float corr = expf(m[i] - mn);        // exp(-inf) = 0 on the first key
          const float e = expf(s - mn);
          l[i] = l[i] * corr + e;
          o[i].x = fmaf(e, vv[u].x, o[i].x * corr);
          o[i].y = fmaf(e, vv[u].y, o[i].y * corr);
          m[i] = mn;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < LQ; ++i) {
    const long long row = n * p.o_rn + i * p.o_ri;
    const float inv = 1.f / l[i];
    act_store1(p.o, row, h * HD + 2 * lane, o[i].x * inv);
    act_store1(p.o, row, h * HD + 2 * lane + 1, o[i].y * inv);
  }
}

// ---- fp16 tensor-core fragments (mma.sync m16n8k16, fp32 accumulate) -----------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D += (ah + al) (bh + bl) without the lo*lo term
__device__ __forceinline__ void mma3_f16(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1,
                                         uint32_t bl0, uint32_t bl1) {
  mma_f16(d, al, bh0, bh1);
  mma_f16(d, ah, bl0, bl1);
  mma_f16(d, ah, bh0, bh1);
}

// ---- tensor-core core for the 60..128-token sequences of AVQCrossAttn, second generation --------------------
// mma.sync m16n8k16 on fp16 (hi, lo) pairs (x = hi + lo keeps 22 bits: fp32-faithful, the serial chain into the
// router needs it).  K and V sit in shared memory as fp16 planes (144-byte rows: conflict-free ldmatrix),
// P never leaves registers (the S accumulator fragments of two key tiles are exactly the A fragment of one P*V k-step).
// One CTA (4 warps) per (sequence, head); warp w owns query rows 16w..16w+15 of each 64-row pass.
//   PL = false: fp32 operands; K / V are split ONCE while they are staged, Q is split in registers (fp32 mode, the exported
//               qat_attention_fwd, the text tower);
//   PL = true : the operands ARRIVE as fp16 (hi, lo) plane pairs written by the in-projection GEMM's epilogue (same bytes
//               as fp32): K / V are 16-byte async copies, Q fragments 32-bit loads issued while those are in flight -- one
//               memory round trip per CTA, no conversion work.
// The kernel is issue / latency bound (ncu: 2 460 warp instructions per warp, 51 % issue utilisation, tensor pipe 32 %), so
// the softmax multiplies by one reciprocal per row instead of dividing every element, masks only the key tiles that reach
// past the limit, and the output addresses are formed once per row.
// 2^x on the SFU (ex2.approx.ftz: 2 ulp); the scores carry log2(e) in their scale, so softmax needs no further multiply -- the
// accurate expf() was a quarter of this kernel's issue slots
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NT16, bool PL>   // key tiles of 16: Lk <= 16*NT16
__global__ void __launch_bounds__(128, 4) attn_mma16_kernel(AttnP p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int LKP = 16 * NT16, RSB = 144;              // padded key count, row stride in bytes (72 halves)
  const int Lq = p.Lq, Lk = p.Lk;
  unsigned char* Kh = smraw;
  unsigned char* Kl = Kh + LKP * RSB;
  unsigned char* Vh = Kl + LKP * RSB;
  unsigned char* Vl = Vh + LKP * RSB;
  const int n = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int kn = (n + p.k_shift) % p.k_mod;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  if constexpr (PL) {
    const long long kbase = kn * p.k_sn + h * HD;
    const __half* kbh = p.kh + kbase;
    const __half* kbl = p.kl + kbase;
    const __half* vbh = p.vh + kbase;
    const __half* vbl = p.vl + kbase;
    for (int f = tid; f < LKP * 8; f += 128) {
      const int j = f >> 3, c8 = f & 7;
      const bool ok = j < Lk;                              // rows beyond Lk are zero-filled
      const long long off = ok ? j * p.k_si + c8 * 8 : 0;
      cp_async16_zfill(Kh + j * RSB + c8 * 16, kbh + off, ok);
      cp_async16_zfill(Kl + j * RSB + c8 * 16, kbl + off, ok);
      cp_async16_zfill(Vh + j * RSB + c8 * 16, vbh + off, ok);
      cp_async16_zfill(Vl + j * RSB + c8 * 16, vbl + off, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  } else {
    const float* kb = p.k + kn * p.k_sn + h * HD;
    const float* vb = p.v + kn * p.k_sn + h * HD;
    // staging in batches of 4 iterations: all 8 loads of a batch are issued before the first is consumed (prefetching the
    // Q fragments as well was measured slower -- it costs the fourth resident CTA)
    constexpr int ITERS = LKP * 16 / 128;
#pragma unroll
    for (int it0 = 0; it0 < ITERS; it0 += 4) {
      float4 kv[4], vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = tid + (it0 + u) * 128;
        const int j = f >> 4, c4 = f & 15;
        kv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        vv[u] = kv[u];
        if (it0 + u < ITERS && j < Lk) {
          kv[u] = __ldg(reinterpret_cast<const float4*>(kb + j * p.k_si) + c4);
          vv[u] = __ldg(reinterpret_cast<const float4*>(vb + j * p.k_si) + c4);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (it0 + u >= ITERS) break;
        const int f = tid + (it0 + u) * 128;
        const int j = f >> 4, c4 = f & 15;
        uint2 hi, lo;
        split_half2(kv[u].x, kv[u].y, hi.x, lo.x);
        split_half2(kv[u].z, kv[u].w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(Kh + j * RSB + c4 * 8) = hi;
        *reinterpret_cast<uint2*>(Kl + j * RSB + c4 * 8) = lo;
        split_half2(vv[u].x, vv[u].y, hi.x, lo.x);
        split_half2(vv[u].z, vv[u].w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(Vh + j * RSB + c4 * 8) = hi;
        *reinterpret_cast<uint2*>(Vl + j * RSB + c4 * 8) = lo;
      }
    }
    __syncthreads();
  }
  bool staged = !PL;
  const uint32_t sKh = tc::smem_u32(Kh), sKl = tc::smem_u32(Kl), sVh = tc::smem_u32(Vh), sVl = tc::smem_u32(Vl);
  const int m8 = lane >> 3, r8 = lane & 7;
  const uint32_t koff = uint32_t((r8 + (m8 >> 1) * 8) * RSB + (m8 & 1) * 16);   // + np*16 rows, + ks*32 bytes
  const uint32_t voff = uint32_t((r8 + (m8 & 1) * 8) * RSB + (m8 >> 1) * 16);   // + kk*16 rows, + dp*32 bytes
  const float invH = 1.f / p.H;
  const float scale2 = p.scale * 1.4426950408889634f;        // scores in units of log2: softmax = 2^(s - max) / sum
  // PL: every warp passes the staging barrier exactly once, inside its first pass (a warp without query rows still takes it)
  for (int r0 = warp * 16; r0 < Lq || !staged; r0 += 64) {
    // ---- Q fragments, rows r0+g / r0+g+8 ----
    const int i0 = r0 + g, i1 = r0 + g + 8;
    uint32_t qh[4][4], ql[4][4];
    if constexpr (PL) {
      // the scale (a power of two on this path) is applied to S
      const long long qo0 = n * p.q_sn + (long long)min(i0, Lq - 1) * p.q_si + h * HD + 2 * t;
      const long long qo1 = n * p.q_sn + (long long)min(i1, Lq - 1) * p.q_si + h * HD + 2 * t;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        qh[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(p.qh + qo0 + ks * 16));
        qh[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(p.qh + qo1 + ks * 16));
        qh[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(p.qh + qo0 + ks * 16 + 8));
        qh[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(p.qh + qo1 + ks * 16 + 8));
        ql[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(p.ql + qo0 + ks * 16));
        ql[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(p.ql + qo1 + ks * 16));
        ql[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(p.ql + qo0 + ks * 16 + 8));
        ql[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(p.ql + qo1 + ks * 16 + 8));
      }
      if (!staged) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        staged = true;
      }
    } else {
      const float* q0 = p.q + n * p.q_sn + (long long)min(i0, Lq - 1) * p.q_si + h * HD + 2 * t;
      const float* q1 = p.q + n * p.q_sn + (long long)min(i1, Lq - 1) * p.q_si + h * HD + 2 * t;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float2 a0 = __ldg(reinterpret_cast<const float2*>(q0 + ks * 16));
        const float2 a1 = __ldg(reinterpret_cast<const float2*>(q1 + ks * 16));
        const float2 a2 = __ldg(reinterpret_cast<const float2*>(q0 + ks * 16 + 8));
        const float2 a3 = __ldg(reinterpret_cast<const float2*>(q1 + ks * 16 + 8));
        split_half2(a0.x * scale2, a0.y * scale2, qh[ks][0], ql[ks][0]);
        split_half2(a1.x * scale2, a1.y * scale2, qh[ks][1], ql[ks][1]);
        split_half2(a2.x * scale2, a2.y * scale2, qh[ks][2], ql[ks][2]);
        split_half2(a3.x * scale2, a3.y * scale2, qh[ks][3], ql[ks][3]);
      }
    }
    // ---- S = Q K^T ----
    float s[2 * NT16][4];
#pragma unroll
    for (int nt = 0; nt < 2 * NT16; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < NT16; ++np) {
        uint32_t bh[4], bl[4];
        ldsm_x4(bh, sKh + koff + np * 16 * RSB + ks * 32);
        ldsm_x4(bl, sKl + koff + np * 16 * RSB + ks * 32);
        mma3_f16(s[2 * np], qh[ks], ql[ks], bh[0], bh[1], bl[0], bl[1]);
        mma3_f16(s[2 * np + 1], qh[ks], ql[ks], bh[2], bh[3], bl[2], bl[3]);
      }
    }
    // ---- softmax over the two rows this thread shares with its quad ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
    const int lim0 = p.causal ? min(Lk, i0 + 1) : Lk, lim1 = p.causal ? min(Lk, i1 + 1) : Lk;
    const int lim_min = p.causal ? min(Lk, r0 + 1) : Lk;     // warp-uniform: key tiles wholly below it need no mask
#pragma unroll
    for (int nt = 0; nt < 2 * NT16; ++nt) {
      if constexpr (PL) {
        s[nt][0] *= scale2; s[nt][1] *= scale2; s[nt][2] *= scale2; s[nt][3] *= scale2;
      }
      if (nt * 8 + 8 > lim_min) {
        const int c = nt * 8 + 2 * t;
        if (c >= lim0) s[nt][0] = -INFINITY;
        if (c >= lim1) s[nt][2] = -INFINITY;
        if (c + 1 >= lim0) s[nt][1] = -INFINITY;
        if (c + 1 >= lim1) s[nt][3] = -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2 * NT16; ++nt) {
      s[nt][0] = ex2_fast(s[nt][0] - mx0);
      s[nt][1] = ex2_fast(s[nt][1] - mx0);
      s[nt][2] = ex2_fast(s[nt][2] - mx1);
      s[nt][3] = ex2_fast(s[nt][3] - mx1);
      s0 += s[nt][0] + s[nt][1];
      s1 += s[nt][2] + s[nt][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float r0inv = 1.f / s0, r1inv = 1.f / s1;        // one reciprocal per row (<= 1 ulp from the per-element division)
#pragma unroll
    for (int nt = 0; nt < 2 * NT16; ++nt) {
      s[nt][0] *= r0inv; s[nt][1] *= r0inv; s[nt][2] *= r1inv; s[nt][3] *= r1inv;
    }
    if (p.attw) {
      float* w0 = p.attw + ((long long)n * Lq + i0) * Lk;
      float* w1 = p.attw + ((long long)n * Lq + i1) * Lk;
#pragma unroll
      for (int nt = 0; nt < 2 * NT16; ++nt) {
        const int c = nt * 8 + 2 * t;
        if (i0 < Lq && c < Lk) atomicAdd(w0 + c, s[nt][0] * invH);
        if (i0 < Lq && c + 1 < Lk) atomicAdd(w0 + c + 1, s[nt][1] * invH);
        if (i1 < Lq && c < Lk) atomicAdd(w1 + c, s[nt][2] * invH);
        if (i1 < Lq && c + 1 < Lk) atomicAdd(w1 + c + 1, s[nt][3] * invH);
      }
    }
    // ---- O = P V ----
    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT16; ++kk) {
      uint32_t ph[4], pl[4];
      split_half2(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
      split_half2(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
      split_half2(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
      split_half2(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t bh[4], bl[4];
        ldsm_x4_t(bh, sVh + voff + kk * 16 * RSB + dp * 32);
        ldsm_x4_t(bl, sVl + voff + kk * 16 * RSB + dp * 32);
        mma3_f16(o[2 * dp], ph, pl, bh[0], bh[1], bl[0], bl[1]);
        mma3_f16(o[2 * dp + 1], ph, pl, bh[2], bh[3], bl[2], bl[3]);
      }
    }
    // the row's output planes, advanced to (row, h*64 + 2t) once: the eight stores of a row differ by a constant column
    if (i0 < Lq) {
      const Act orow = p.o.rows(n * p.o_rn + i0 * p.o_ri).cols(h * HD + 2 * t);
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) act_store2(orow, 0, dt * 8, o[dt][0], o[dt][1]);
    }
    if (i1 < Lq) {
      const Act orow = p.o.rows(n * p.o_rn + i1 * p.o_ri).cols(h * HD + 2 * t);
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) act_store2(orow, 0, dt * 8, o[dt][2], o[dt][3]);
    }
  }
}

template <int NT16, bool PL>
inline int attn_mma16_launch(Ctx& c, const AttnP& p) {
  const size_t smem = (size_t)4 * 16 * NT16 * 144;
  auto kern = attn_mma16_kernel<NT16, PL>;
  if (smem > 48 * 1024) QAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)(p.nseq * p.H), 128, smem, c.stream>>>(p);
  return check_launch("attn_mma16_kernel");
}
template <bool PL>
inline int attn_mma16_dispatch(Ctx& c, const AttnP& p) {
  if (p.Lk <= 64) return attn_mma16_launch<4, PL>(c, p);
  if (p.Lk <= 80) return attn_mma16_launch<5, PL>(c, p);
  if (p.Lk <= 96) return attn_mma16_launch<6, PL>(c, p);
  return attn_mma16_launch<8, PL>(c, p);
}

inline int attention(Ctx& c, const AttnP& p) {
  QAT_REQUIRE(p.Lk >= 1 && p.Lk <= 128, "attention: key length %d outside [1,128]", p.Lk);
  QAT_REQUIRE(p.Lq >= 1 && p.nseq >= 0 && p.H >= 1, "attention: bad shape");
  if (c.dry || p.nseq == 0) return QAT_OK;
  if (p.qh) {
    QAT_REQUIRE(p.ql && p.kh && p.kl && p.vh && p.vl, "attention: fp16 plane pairs need all six planes");
    QAT_REQUIRE(p.q_si % 8 == 0 && p.k_si % 8 == 0 && p.q_sn % 8 == 0 && p.k_sn % 8 == 0, "attention: plane strides must be multiples of 8 elements");
    return attn_mma16_dispatch<true>(c, p);
  }
  if (p.causal) {
    QAT_REQUIRE(p.attw == nullptr, "attention: causal attention has no weights output");
    return attn_mma16_dispatch<false>(c, p);
  }
  if (p.Lq <= 2 && p.attw == nullptr) {
    const long long probs = (long long)p.nseq * p.H;
    const unsigned grid = (unsigned)((probs + 7) / 8);
    if (p.Lq == 1) attn_fewq_kernel<1><<<grid, 256, 0, c.stream>>>(p);
    else attn_fewq_kernel<2><<<grid, 256, 0, c.stream>>>(p);
    return check_launch("attn_fewq_kernel");
  }
  if (p.Lk <= ROWS_MAXL && p.Lq <= ROWS_MAXL && p.H * p.Lq <= 128) {
    size_t smem = size_t(2) * p.H * (p.Lk * HD + 4) * 4;
    if (smem > 48 * 1024) QAT_CUDA(cudaFuncSetAttribute(attn_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nt = ((p.H * p.Lq + 31) / 32) * 32;
    attn_rows_kernel<<<(unsigned)p.nseq, nt, smem, c.stream>>>(p);
    return check_launch("attn_rows_kernel");
  }
  if (p.Lk > 32 || p.Lq >= 32) {
    return attn_mma16_dispatch<false>(c, p);
  }
  if (p.Lk <= 32) {
    constexpr int WPB = 8;
    size_t smem = size_t(WPB) * (ks_floats(p.Lk) + p.Lk * 64) * 4;
    if (smem > 48 * 1024)
      QAT_CUDA(cudaFuncSetAttribute(attn_warp_kernel<WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long probs = (long long)p.nseq * p.H;
    attn_warp_kernel<WPB><<<(unsigned)((probs + WPB - 1) / WPB), 32 * WPB, smem, c.stream>>>(p);
    return check_launch("attn_warp_kernel");
  }
  size_t smem = (size_t(ks_floats(p.Lk)) + size_t(p.Lk) * 64 + 4 * 128) * 4;
  if (smem > 48 * 1024)
    QAT_CUDA(cudaFuncSetAttribute(attn_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_block_kernel<<<(unsigned)(p.nseq * p.H), 128, smem, c.stream>>>(p);
  return check_launch("attn_block_kernel");
}

}  // namespace qat
