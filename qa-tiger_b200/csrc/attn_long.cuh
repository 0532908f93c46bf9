// Softmax(Q K^T * scale + key_bias) V for sequences longer than the 128-key cores of attn_core.cuh -- the ViT-L attention of the
// ToMe extractor (577 -> 27 tokens, 16 heads of 64; /root/reference/src/tome/patch/timm.py:68-108), whose proportional attention
// adds log(size[key]) to every score (timm.py:94-95).  Flash-style: one CTA (4 warps) per (sequence, head, block of 64 queries);
// keys / values stream through shared memory in blocks of 64 as fp16 (hi, lo) planes (x = hi + lo keeps 22 bits: fp32-faithful,
// like attn_mma16_kernel -- the attention output feeds the residual stream the NEXT block's merge metric is computed from), S and
// O live in mma.sync m16n8k16 accumulator fragments, the softmax is the online (running max / running sum) form.
#pragma once
#include "attn_core.cuh"

namespace qat {

__global__ void __launch_bounds__(128, 4) attn_long_kernel(AttnP p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int KB = 64, NT16 = KB / 16, RSB = 144;
  const int Lq = p.Lq, Lk = p.Lk;
  unsigned char* Kh = smraw;
  unsigned char* Kl = Kh + KB * RSB;
  unsigned char* Vh = Kl + KB * RSB;
  unsigned char* Vl = Vh + KB * RSB;
  float* s_bias = reinterpret_cast<float*>(Vl + KB * RSB);      // [KB]
  const int nqb = (Lq + 63) / 64;
  const int qb = blockIdx.x % nqb, nh = blockIdx.x / nqb;
  const int n = nh / p.H, h = nh % p.H;
  const int kn = (n + p.k_shift) % p.k_mod;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float* kb = p.k + kn * p.k_sn + h * HD;
  const float* vb = p.v + kn * p.k_sn + h * HD;
  const float* bias = p.kbias ? p.kbias + (long long)kn * Lk : nullptr;
  const uint32_t sKh = tc::smem_u32(Kh), sKl = tc::smem_u32(Kl), sVh = tc::smem_u32(Vh), sVl = tc::smem_u32(Vl);
  const int m8 = lane >> 3, r8 = lane & 7;
  const uint32_t koff = uint32_t((r8 + (m8 >> 1) * 8) * RSB + (m8 & 1) * 16);
  const uint32_t voff = uint32_t((r8 + (m8 & 1) * 8) * RSB + (m8 >> 1) * 16);
  // ---- Q fragments (scaled), rows r0+g / r0+g+8 ----
  const int r0 = qb * 64 + warp * 16;
  const int i0 = r0 + g, i1 = r0 + g + 8;
  uint32_t qh[4][4], ql[4][4];
  {
    const float* q0 = p.q + n * p.q_sn + (long long)min(i0, Lq - 1) * p.q_si + h * HD + 2 * t;
    const float* q1 = p.q + n * p.q_sn + (long long)min(i1, Lq - 1) * p.q_si + h * HD + 2 * t;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const float2 a0 = __ldg(reinterpret_cast<const float2*>(q0 + ks * 16));
      const float2 a1 = __ldg(reinterpret_cast<const float2*>(q1 + ks * 16));
      const float2 a2 = __ldg(reinterpret_cast<const float2*>(q0 + ks * 16 + 8));
      const float2 a3 = __ldg(reinterpret_cast<const float2*>(q1 + ks * 16 + 8));
      split_half2(a0.x * p.scale, a0.y * p.scale, qh[ks][0], ql[ks][0]);
      split_half2(a1.x * p.scale, a1.y * p.scale, qh[ks][1], ql[ks][1]);
      split_half2(a2.x * p.scale, a2.y * p.scale, qh[ks][2], ql[ks][2]);
      split_half2(a3.x * p.scale, a3.y * p.scale, qh[ks][3], ql[ks][3]);
    }
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;

  for (int k0 = 0; k0 < Lk; k0 += KB) {
    __syncthreads();                                         // every warp is done with the previous block
    // ---- stage keys / values k0 .. k0+63 (rows beyond Lk zero-filled), split once into fp16 (hi, lo) ----
    constexpr int ITERS = KB * 16 / 128;
#pragma unroll
    for (int it0 = 0; it0 < ITERS; it0 += 4) {
      float4 kv[4], vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = tid + (it0 + u) * 128;
        const int j = f >> 4, c4 = f & 15;
        kv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        vv[u] = kv[u];
        if (k0 + j < Lk) {
          kv[u] = __ldg(reinterpret_cast<const float4*>(kb + (long long)(k0 + j) * p.k_si) + c4);
          vv[u] = __ldg(reinterpret_cast<const float4*>(vb + (long long)(k0 + j) * p.k_si) + c4);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
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
    if (tid < KB) s_bias[tid] = (bias && k0 + tid < Lk) ? __ldg(bias + k0 + tid) : 0.f;
    __syncthreads();
    // ---- S = Q K^T (+ key bias), keys beyond Lk masked ----
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
    float bm0 = -INFINITY, bm1 = -INFINITY;
    const int lim = Lk - k0;                                 // keys of this block that exist
#pragma unroll
    for (int nt = 0; nt < 2 * NT16; ++nt) {
      const int c = nt * 8 + 2 * t;
      const float b0 = s_bias[c], b1 = s_bias[c + 1];
      s[nt][0] = c < lim ? s[nt][0] + b0 : -INFINITY;
      s[nt][2] = c < lim ? s[nt][2] + b0 : -INFINITY;
      s[nt][1] = c + 1 < lim ? s[nt][1] + b1 : -INFINITY;
      s[nt][3] = c + 1 < lim ? s[nt][3] + b1 : -INFINITY;
      bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
      bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    // ---- online softmax: rescale what has been accumulated to the new running maximum ----
    const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);   // finite: every block holds at least one existing key
    const float c0 = expf(m0 - mn0), c1 = expf(m1 - mn1);     // exp(-inf) = 0 on the first block
    m0 = mn0;
    m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2 * NT16; ++nt) {
      s[nt][0] = expf(s[nt][0] - mn0);
      s[nt][1] = expf(s[nt][1] - mn0);
      s[nt][2] = expf(s[nt][2] - mn1);
      s[nt][3] = expf(s[nt][3] - mn1);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
    }
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1);
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1;
    }
    // ---- O += P V ----
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
  }
  const float r0inv = 1.f / l0, r1inv = 1.f / l1;
  if (i0 < Lq) {
    const Act orow = p.o.rows(n * p.o_rn + i0 * p.o_ri).cols(h * HD + 2 * t);
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) act_store2(orow, 0, dt * 8, o[dt][0] * r0inv, o[dt][1] * r0inv);
  }
  if (i1 < Lq) {
    const Act orow = p.o.rows(n * p.o_rn + i1 * p.o_ri).cols(h * HD + 2 * t);
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) act_store2(orow, 0, dt * 8, o[dt][2] * r1inv, o[dt][3] * r1inv);
  }
}

inline int attention_long(Ctx& c, const AttnP& p) {
  QAT_REQUIRE(p.q && p.k && p.v, "attention (long): fp32 q / k / v required");
  QAT_REQUIRE(!p.causal && !p.attw, "attention (long): no causal mask / weights output on sequences longer than 128 keys");
  QAT_REQUIRE(p.q_si % 2 == 0 && p.q_sn % 2 == 0 && p.k_si % 4 == 0 && p.k_sn % 4 == 0, "attention (long): q strides must be even, k / v strides multiples of 4");
  if (c.dry || p.nseq == 0) return QAT_OK;
  const int nqb = (p.Lq + 63) / 64;
  const long long grid = (long long)p.nseq * p.H * nqb;
  QAT_REQUIRE(grid < (1ll << 31), "attention (long): grid too large");
  const size_t smem = (size_t)4 * 64 * 144 + 64 * 4;
  attn_long_kernel<<<(unsigned)grid, 128, smem, c.stream>>>(p);
  return check_launch("attn_long_kernel");
}

}  // namespace qat
