// Attention cores of the PatchSelecter's fp16-plane path (bf16 mode): Q/K/V arrive as fp16 planes straight out of
// the tcgen05 GEMM, softmax in fp32, outputs as fp16 planes (or any Act for the 2-query cross-attention).
#pragma once
#include "attn_core.cuh"

namespace qat {

struct AttnHP {
  const __half* q;       // [nseq, Lq, ldq] rows; head h at columns h*64
  const __half* kv;      // [nseq, Lk, ldkv] rows: K at columns [0, D), V at [D, 2D) of the row (K | V packed)
  long long ldq, ldkv;   // row strides (elements)
  __half* o;             // [nseq, Lq, ldo]
  long long ldo;
  int nseq, H, Lq, Lk;
  float scale;
};

// ---- attention over <= 16 queries and <= 16 keys (modules.py:688 the 14 patch tokens of a frame attending to each
// other; :698 the frame's [video; audio] query pair attending to them) -----------------------------------------
// One CTA per sequence, one warp per head.  The sequence's fp16 rows (Q: [Lq, D], K|V: [Lk, 2D]) are staged once
// with cp.async (row strides padded by 16 bytes: conflict-free ldmatrix); S = Q K^T is 4 k-steps x 2 key tiles of
// mma.m16n8k16 (fp32 accumulate), the softmax lives in the C fragments (a row is spread over a quad), P goes
// back in as the A fragment (fp16) and O = P V is 8 dim tiles x 1 k-step with V through ldmatrix.trans.
// Rows / keys beyond the sequence are clamped to the last row when addressing shared memory; padded keys are
// masked to -inf, padded query rows are computed and dropped.
template <int H>
__global__ void __launch_bounds__(32 * H) attn_small_f16_kernel(AttnHP p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int D = H * HD;
  constexpr int rsq = (D + 8) * 2, rskv = (2 * D + 8) * 2;      // smem row strides in bytes
  const int Lq = p.Lq, Lk = p.Lk;
  unsigned char* sq = smem_raw;
  unsigned char* skv = smem_raw + Lq * rsq;
  const int tid = threadIdx.x, lane = tid & 31, h = tid >> 5;
  const long long n = blockIdx.x;
  {
    constexpr int qch = D / 8, kch = 2 * D / 8;       // 16-byte chunks per row
    const __half* qb = p.q + n * Lq * p.ldq;
    const __half* kb = p.kv + n * Lk * p.ldkv;
    for (int f = tid; f < Lq * qch; f += 32 * H) {
      const int r = f / qch, c = f % qch;
      cp_async16(sq + r * rsq + c * 16, qb + r * p.ldq + c * 8);
    }
    for (int f = tid; f < Lk * kch; f += 32 * H) {
      const int r = f / kch, c = f % kch;
      cp_async16(skv + r * rskv + c * 16, kb + r * p.ldkv + c * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const uint32_t sqb = tc::smem_u32(sq), skb = tc::smem_u32(skv);
  const int g = lane >> 2, t = lane & 3;
  const int m = lane >> 3, r8 = lane & 7;
  float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  {
    const int qrow = min(r8 + (m & 1) * 8, Lq - 1), qcol = (m >> 1) * 8;
    const int krow = min(r8 + (m >> 1) * 8, Lk - 1), kcol = (m & 1) * 8;
    const uint32_t qa = sqb + qrow * rsq + (h * HD + qcol) * 2;
    const uint32_t ka = skb + krow * rskv + (h * HD + kcol) * 2;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4], b[4];
      ldsm_x4(a, qa + ks * 32);
      ldsm_x4(b, ka + ks * 32);
      mma_f16(s[0], a, b[0], b[1]);
      mma_f16(s[1], a, b[2], b[3]);
    }
  }
  // softmax over the keys of rows g (elements 0,1) and g+8 (elements 2,3)
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool ok = nt * 8 + 2 * t + e < Lk;
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
      s[nt][e] = expf(s[nt][e] - mx0);               // exp(-inf) = 0 for the padded keys
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
  pa[1] = pack_half2(s[0][2] * i1, s[0][3] * i1);
  pa[2] = pack_half2(s[1][0] * i0, s[1][1] * i0);
  pa[3] = pack_half2(s[1][2] * i1, s[1][3] * i1);
  float o[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;
  {
    const int vrow = min(r8 + (m & 1) * 8, Lk - 1), vcol = (m >> 1) * 8;
    const uint32_t va = skb + vrow * rskv + (D + h * HD + vcol) * 2;
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldsm_x4_t(b, va + dp * 32);
      mma_f16(o[2 * dp], pa, b[0], b[1]);
      mma_f16(o[2 * dp + 1], pa, b[2], b[3]);
    }
  }
  // O of this head overwrites its own Q columns (only this warp reads them, and it is done), then the CTA
  // copies whole 2*D-byte rows out with 16-byte stores.
  __syncwarp();
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int col = h * HD + dt * 8 + 2 * t;
    if (g < Lq) *reinterpret_cast<uint32_t*>(sq + g * rsq + col * 2) = pack_half2(o[dt][0], o[dt][1]);
    if (g + 8 < Lq) *reinterpret_cast<uint32_t*>(sq + (g + 8) * rsq + col * 2) = pack_half2(o[dt][2], o[dt][3]);
  }
  __syncthreads();
  {
    constexpr int dch = D / 8;
    for (int f = tid; f < Lq * dch; f += 32 * H) {
      const int r = f / dch, c = f % dch;
      *reinterpret_cast<uint4*>(p.o + (n * Lq + r) * p.ldo + c * 8) = *reinterpret_cast<const uint4*>(sq + r * rsq + c * 16);
    }
  }
}

inline int attention_small_f16(Ctx& c, const AttnHP& p) {
  QAT_REQUIRE(p.Lq >= 1 && p.Lq <= 16 && p.Lk >= 1 && p.Lk <= 16 && p.H >= 1 && p.H <= 8,
              "attention_small_f16: need Lq, Lk <= 16 and H <= 8 (Lq=%d Lk=%d H=%d)", p.Lq, p.Lk, p.H);
  QAT_REQUIRE(p.ldq % 8 == 0 && p.ldkv % 8 == 0 && p.ldo % 8 == 0, "attention_small_f16: row strides must be multiples of 8 elements");
  if (c.dry || p.nseq == 0) return QAT_OK;
  QAT_REQUIRE(((reinterpret_cast<uintptr_t>(p.q) | reinterpret_cast<uintptr_t>(p.kv) | reinterpret_cast<uintptr_t>(p.o)) & 15) == 0,
              "attention_small_f16: tensors must be 16-byte aligned");
  const int D = p.H * HD;
  const size_t smem = (size_t)p.Lq * (D + 8) * 2 + (size_t)p.Lk * (2 * D + 8) * 2;
  switch (p.H) {
#define QAT_SMALL_CASE(HH)                                                                                                    \
  case HH:                                                                                                                    \
    if (smem > 48 * 1024) QAT_CUDA(cudaFuncSetAttribute(attn_small_f16_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    attn_small_f16_kernel<HH><<<(unsigned)p.nseq, 32 * HH, smem, c.stream>>>(p);                                              \
    break;
    QAT_SMALL_CASE(1) QAT_SMALL_CASE(2) QAT_SMALL_CASE(3) QAT_SMALL_CASE(4) QAT_SMALL_CASE(5) QAT_SMALL_CASE(6) QAT_SMALL_CASE(7) QAT_SMALL_CASE(8)
#undef QAT_SMALL_CASE
    default: QAT_FAIL(QAT_EINVAL, "attention_small_f16: H=%d unsupported", p.H);
  }
  return check_launch("attn_small_f16_kernel");
}

// ---- <= 2 fp32 query rows against fp16 K/V rows (PatchSelecter cross-attention, modules.py:698) -------------
// Same streaming / online-softmax structure as attn_fewq_kernel; lane l owns dims (2l, 2l+1) = one half2.
template <int LQ>
__global__ void __launch_bounds__(256) attn_fewq_f16_kernel(AttnP p, const __half* __restrict__ kh, const __half* __restrict__ vh) {
  const int lane = threadIdx.x & 31;
  const long long prob = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (prob >= (long long)p.nseq * p.H) return;
  const int n = (int)(prob / p.H), h = (int)(prob % p.H);
  const int kn = (n + p.k_shift) % p.k_mod;
  const __half* kb = kh + kn * p.k_sn + h * HD + 2 * lane;
  const __half* vb = vh + kn * p.k_sn + h * HD + 2 * lane;
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
  constexpr int UN = 16;            // the 14 patch tokens of a frame go out as ONE batch of 2 x 14 independent loads
  for (int j0 = 0; j0 < Lk; j0 += UN) {
    uint32_t kk[UN], vv[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u < Lk ? j0 + u : Lk - 1;
      kk[u] = __ldg(reinterpret_cast<const uint32_t*>(kb + j * p.k_si));
      vv[u] = __ldg(reinterpret_cast<const uint32_t*>(vb + j * p.k_si));
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (j0 + u < Lk) {
        const float2 kf = unpack_half2(kk[u]), vf = unpack_half2(vv[u]);
#pragma unroll
        for (int i = 0; i < LQ; ++i) {
          float s = warp_sum(fmaf(q[i].x, kf.x, q[i].y * kf.y));
          const float mn = fmaxf(m[i], s);
          const float corr = expf(m[i] - mn);
          const float e = expf(s - mn);
          l[i] = l[i] * corr + e;
          o[i].x = fmaf(e, vf.x, o[i].x * corr);
          o[i].y = fmaf(e, vf.y, o[i].y * corr);
          m[i] = mn;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < LQ; ++i) {
    const long long row = n * p.o_rn + i * p.o_ri;
    const float inv = 1.f / l[i];
    act_store2(p.o, row, h * HD + 2 * lane, o[i].x * inv, o[i].y * inv);
  }
}

// p.k / p.v are ignored; strides k_sn / k_si are in fp16 elements
inline int attention_fewq_f16(Ctx& c, const AttnP& p, const __half* kh, const __half* vh) {
  QAT_REQUIRE(p.Lq >= 1 && p.Lq <= 2 && p.Lk >= 1 && p.attw == nullptr, "attention_fewq_f16: need 1 <= Lq <= 2 and no weights output");
  if (c.dry || p.nseq == 0) return QAT_OK;
  const long long probs = (long long)p.nseq * p.H;
  const unsigned grid = (unsigned)((probs + 7) / 8);
  if (p.Lq == 1) attn_fewq_f16_kernel<1><<<grid, 256, 0, c.stream>>>(p, kh, vh);
  else attn_fewq_f16_kernel<2><<<grid, 256, 0, c.stream>>>(p, kh, vh);
  return check_launch("attn_fewq_f16_kernel");
}

}  // namespace qat
