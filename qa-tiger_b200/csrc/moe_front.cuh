// TempMoE front end (modules.py:562-578): temp_w = qst_attn(q = question, kv = data) -> router logits | gauss_pred, for up to two
// TempMoE modules (at_aggregator, vt_aggregator) per launch.
//
// The reference projects all T rows of `data` to K and V (two [B*T, D] x [D, D] contractions per module) to attend to them with
// ONE query row per sample, then out-projects the result and applies two more Linear layers.  With a single query the projections
// commute with the attention:
//     score[b,t,h] = q_h[b] . (Wk_h x[b,t] + bk_h) = x[b,t] . u_h[b] + const(b,h),   u_h[b] = Wk_h^T q_h[b]      (const drops out
//                                                                                                         of the softmax over t)
//     o_h[b] = sum_t p[b,t,h] (Wv_h x[b,t] + bv_h) = Wv_h xbar_h[b] + bv_h,           xbar_h[b] = sum_t p[b,t,h] x[b,t]
//     [router | gauss_pred](Wo o + bo) = sum_h Wrg''_h xbar_h[b] + brg'',              Wrg''_h = (Wrg Wo)[:, h] Wv_h   (host, fp64)
// so the K | V projection (16 GFLOP nominal per module and batch, a bf16x3 tensor-core GEMM before), the out-projection and the
// router / gauss_pred contractions collapse into: u (a [B, 64] x [64, D] contraction per head), two passes over the fp32 rows of
// `data` (scores, then the p-weighted row sums) and a [3E, H*D] matrix-vector product per sample -- all fp32 FFMA, which is also
// what the rank-critical path wants.  Exact algebra; only the summation order differs from the reference.
#pragma once
#include "qat_common.cuh"

namespace qat {

constexpr int MF_MAX = 2;   // modules per launch

struct MoeFrontP {
  int n, B, T, D, H, E3;               // E3 = 3 * n_experts (router E | gauss_pred 2E)
  float scale;
  const float* q[MF_MAX];               // projected question query attn_q(qst) [B, ldq]
  int ldq[MF_MAX];
  const float* wk[MF_MAX];              // in_proj_weight rows [D, 2D) = Wk, [D, D] row-major
  const float* x[MF_MAX];               // data [B*T, D] fp32
  const float* wrg[MF_MAX];             // Wrg'' [E3, H*D]
  const float* brg[MF_MAX];             // brg'' [E3]
  float* u[MF_MAX];                     // scratch [B, H*D]
  float* rg[MF_MAX];                    // out [B, E3]
};

// u[b, h*D + c] = sum_d q[b, h*64 + d] Wk[h*64 + d, c]: one CTA per (16 samples, head, module); thread = two channels c
constexpr int MF_UB = 16;
__global__ void __launch_bounds__(256) moe_u_kernel(MoeFrontP p) {
  const int m = blockIdx.z, h = blockIdx.y, b0 = blockIdx.x * MF_UB;
  const int D = p.D, tid = threadIdx.x;
  __shared__ float sq[MF_UB][64];
  for (int f = tid; f < MF_UB * 64; f += 256) {
    const int s = f >> 6, d = f & 63;
    sq[s][d] = b0 + s < p.B ? __ldg(p.q[m] + (long long)(b0 + s) * p.ldq[m] + h * 64 + d) : 0.f;
  }
  __syncthreads();
  for (int c0 = 0; c0 < D; c0 += 512) {
    const int c = c0 + tid * 2;
    if (c >= D) break;
    float2 acc[MF_UB];
#pragma unroll
    for (int s = 0; s < MF_UB; ++s) acc[s] = make_float2(0.f, 0.f);
    const float* w = p.wk[m] + (long long)h * 64 * D + c;
#pragma unroll 4
    for (int d = 0; d < 64; ++d) {
      const float2 wv = __ldg(reinterpret_cast<const float2*>(w + (long long)d * D));
#pragma unroll
      for (int s = 0; s < MF_UB; ++s) {
        acc[s].x = fmaf(sq[s][d], wv.x, acc[s].x);
        acc[s].y = fmaf(sq[s][d], wv.y, acc[s].y);
      }
    }
#pragma unroll
    for (int s = 0; s < MF_UB; ++s)
      if (b0 + s < p.B) *reinterpret_cast<float2*>(p.u[m] + ((long long)(b0 + s) * p.H + h) * D + c) = acc[s];
  }
}

// one CTA per (sample, module): scores -> softmax over t -> xbar -> [router | gauss_pred] logits.  D == 512, H == 8.
constexpr int MF_NT = 256;
__global__ void __launch_bounds__(MF_NT) moe_attn_rg_kernel(MoeFrontP p) {
  extern __shared__ __align__(16) float mf_smem[];
  const int m = blockIdx.y, b = blockIdx.x, T = p.T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int D = 512, H = 8;
  float* s_u = mf_smem;                 // [H][D]
  float* s_xb = s_u + H * D;            // [H][D]
  float* s_p = s_xb + H * D;            // [T][H]
  const float* xb = p.x[m] + (long long)b * T * D;
  {
    const float4* src = reinterpret_cast<const float4*>(p.u[m] + (long long)b * H * D);
    float4* dst = reinterpret_cast<float4*>(s_u);
    for (int f = tid; f < H * D / 4; f += MF_NT) dst[f] = __ldg(src + f);
  }
  __syncthreads();
  // pass 1: score[t][h] = scale * x[b,t] . u_h   (warp per row; lane holds channels 4*lane + 128*k)
  for (int t = warp; t < T; t += MF_NT / 32) {
    float4 xv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xv[k] = __ldg(reinterpret_cast<const float4*>(xb + (long long)t * D + 128 * k + 4 * lane));
    float d[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 uv = *reinterpret_cast<const float4*>(s_u + h * D + 128 * k + 4 * lane);
        a = fmaf(xv[k].x, uv.x, a);
        a = fmaf(xv[k].y, uv.y, a);
        a = fmaf(xv[k].z, uv.z, a);
        a = fmaf(xv[k].w, uv.w, a);
      }
      d[h] = a;
    }
#pragma unroll
    for (int h = 0; h < H; ++h) d[h] = warp_sum(d[h]);
    if (lane < H) {
      float v = d[0];
#pragma unroll
      for (int h = 1; h < H; ++h) v = lane == h ? d[h] : v;
      s_p[t * H + lane] = v * p.scale;
    }
  }
  __syncthreads();
  // softmax over t, one warp per head
  if (warp < H) {
    float mx = -INFINITY;
    for (int t = lane; t < T; t += 32) mx = fmaxf(mx, s_p[t * H + warp]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float e = expf(s_p[t * H + warp] - mx);
      s_p[t * H + warp] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    for (int t = lane; t < T; t += 32) s_p[t * H + warp] = s_p[t * H + warp] / sum;
  }
  __syncthreads();
  // pass 2: xbar[h][c] = sum_t p[t][h] x[b,t,c]   (thread = channels 2*tid, 2*tid + 1; rows re-read from L2)
  {
    float2 acc[H];
#pragma unroll
    for (int h = 0; h < H; ++h) acc[h] = make_float2(0.f, 0.f);
    const float2* xr = reinterpret_cast<const float2*>(xb) + tid;
    int t = 0;
    for (; t + 4 <= T; t += 4) {
      float2 xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = __ldg(xr + (long long)(t + i) * (D / 2));
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float pw = s_p[(t + i) * H + h];
          acc[h].x = fmaf(pw, xv[i].x, acc[h].x);
          acc[h].y = fmaf(pw, xv[i].y, acc[h].y);
        }
    }
    for (; t < T; ++t) {
      const float2 xv = __ldg(xr + (long long)t * (D / 2));
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float pw = s_p[t * H + h];
        acc[h].x = fmaf(pw, xv.x, acc[h].x);
        acc[h].y = fmaf(pw, xv.y, acc[h].y);
      }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) *reinterpret_cast<float2*>(s_xb + h * D + 2 * tid) = acc[h];
  }
  __syncthreads();
  // [router | gauss_pred][n] = brg''[n] + Wrg''[n, :] . xbar   (warp per output; 128-bit loads of the weight row)
  for (int n = warp; n < p.E3; n += MF_NT / 32) {
    const float4* w = reinterpret_cast<const float4*>(p.wrg[m] + (long long)n * H * D);
    const float4* xv = reinterpret_cast<const float4*>(s_xb);
    float a = 0.f;
#pragma unroll 8
    for (int f = lane; f < H * D / 4; f += 32) {
      const float4 wv = __ldg(w + f);
      const float4 x4 = xv[f];
      a = fmaf(wv.x, x4.x, a);
      a = fmaf(wv.y, x4.y, a);
      a = fmaf(wv.z, x4.z, a);
      a = fmaf(wv.w, x4.w, a);
    }
    a = warp_sum(a);
    if (lane == 0) p.rg[m][(long long)b * p.E3 + n] = a + __ldg(p.brg[m] + n);
  }
}

inline int moe_front_launch(Ctx& c, const MoeFrontP& p) {
  QAT_REQUIRE(p.n >= 1 && p.n <= MF_MAX && p.B >= 1 && p.T >= 1 && p.T <= 1024, "moe_front: bad shape");
  QAT_REQUIRE(p.D == 512 && p.H == 8, "moe_front: built for d_model = 512, 8 heads (got D=%d H=%d)", p.D, p.H);
  QAT_REQUIRE(p.E3 >= 3 && p.E3 <= 96, "moe_front: 3*E=%d outside [3, 96]", p.E3);
  if (c.dry) return QAT_OK;
  for (int m = 0; m < p.n; ++m) {
    QAT_REQUIRE(p.q[m] && p.wk[m] && p.x[m] && p.wrg[m] && p.brg[m] && p.u[m] && p.rg[m], "moe_front: NULL tensor (module %d)", m);
    QAT_REQUIRE(((reinterpret_cast<uintptr_t>(p.x[m]) | reinterpret_cast<uintptr_t>(p.wrg[m]) | reinterpret_cast<uintptr_t>(p.u[m]) |
                  reinterpret_cast<uintptr_t>(p.wk[m])) & 15) == 0, "moe_front: tensors must be 16-byte aligned");
  }
  dim3 gu((p.B + MF_UB - 1) / MF_UB, p.H, p.n);
  moe_u_kernel<<<gu, 256, 0, c.stream>>>(p);
  QAT_TRY(check_launch("moe_u_kernel"));
  const size_t smem = ((size_t)2 * 8 * 512 + (size_t)p.T * 8) * 4;
  if (smem > 48 * 1024) QAT_CUDA(cudaFuncSetAttribute(moe_attn_rg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 ga(p.B, p.n);
  moe_attn_rg_kernel<<<ga, MF_NT, smem, c.stream>>>(p);
  return check_launch("moe_attn_rg_kernel");
}

}  // namespace qat
