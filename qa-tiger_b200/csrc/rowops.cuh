// Row-wise helpers: LayerNorm (eps 1e-5, biased variance, affine) over rows of D<=1024, fused with
// the small linear combinations that precede it on the hot path, plus add / split / copy kernels.
#pragma once
#include "qat_common.cuh"

namespace qat {

// y = LN(c0*x0 + c1*x1 + x2) * gamma + beta ; x1, x2 optional.  One warp per row.
// x0 + x1 is formed first, then scaled by c01, then x2 is added -- the order used by
// QstGrounding (`data.mean(0) + mlp(.)`, modules.py:344) and by plain LN (x1=x2=null, c01=1).
struct LnP {
  const float* x0; long long ld0;
  const float* x1; long long ld1;
  const float* x2; long long ld2;
  float c01;
  const float* gamma;
  const float* beta;
  Act out;             // row r -> out row r
  const float* post = nullptr;   // optional: out = post[r, :] + LN(...)  (the `data + sub_data` of TempMoE, modules.py:587,597,
  long long ldp = 0;             //           fused into the LayerNorm that produces sub_data)
  long long rows;
  int D;
  float eps;
};

template <int VPL>  // float4 per lane: D = 128*VPL
__global__ void __launch_bounds__(256) ln_kernel(LnP p) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  float4 v[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    int c4 = lane + 32 * i;
    float4 a = __ldg(reinterpret_cast<const float4*>(p.x0 + row * p.ld0) + c4);
    if (p.x1) {
      float4 b = __ldg(reinterpret_cast<const float4*>(p.x1 + row * p.ld1) + c4);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      a.x *= p.c01; a.y *= p.c01; a.z *= p.c01; a.w *= p.c01;
    }
    if (p.x2) {
      float4 b = __ldg(reinterpret_cast<const float4*>(p.x2 + row * p.ld2) + c4);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    v[i] = a;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / p.D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = rsqrtf(warp_sum(q) / p.D + p.eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    int c4 = lane + 32 * i;
    float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma) + c4);
    float4 b = __ldg(reinterpret_cast<const float4*>(p.beta) + c4);
    float4 o;
    o.x = v[i].x * rstd * g.x + b.x;
    o.y = v[i].y * rstd * g.y + b.y;
    o.z = v[i].z * rstd * g.z + b.z;
    o.w = v[i].w * rstd * g.w + b.w;
    if (p.post) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p.post + row * p.ldp) + c4);
      o.x = a.x + o.x; o.y = a.y + o.y; o.z = a.z + o.z; o.w = a.w + o.w;
    }
    act_store4(p.out, row, c4 * 4, o);
  }
}

inline int layer_norm(Ctx& c, const LnP& p) {
  QAT_REQUIRE(p.D % 128 == 0 && p.D <= 1024, "layer_norm: D=%d must be a multiple of 128 and <= 1024", p.D);
  if (c.dry || p.rows == 0) return QAT_OK;
  unsigned grid = (unsigned)((p.rows + 7) / 8);
  switch (p.D / 128) {
    case 1: ln_kernel<1><<<grid, 256, 0, c.stream>>>(p); break;
    case 2: ln_kernel<2><<<grid, 256, 0, c.stream>>>(p); break;
    case 3: ln_kernel<3><<<grid, 256, 0, c.stream>>>(p); break;
    case 4: ln_kernel<4><<<grid, 256, 0, c.stream>>>(p); break;
    case 5: ln_kernel<5><<<grid, 256, 0, c.stream>>>(p); break;
    case 6: ln_kernel<6><<<grid, 256, 0, c.stream>>>(p); break;
    case 7: ln_kernel<7><<<grid, 256, 0, c.stream>>>(p); break;
    default: ln_kernel<8><<<grid, 256, 0, c.stream>>>(p); break;
  }
  return check_launch("ln_kernel");
}

// out[r, :] = a[r, :] + b[r, :]  (b optional) -- also the fp32 -> (fp32, hi, lo) plane converter
__global__ void __launch_bounds__(256) add_rows_kernel(const float* __restrict__ a, long long lda,
                                                       const float* __restrict__ b, long long ldb, Act out,
                                                       long long rows, int cols4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = rows * cols4;
  for (; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / cols4;
    int c4 = (int)(i % cols4);
    float4 x = __ldg(reinterpret_cast<const float4*>(a + r * lda) + c4);
    if (b) {
      float4 y = __ldg(reinterpret_cast<const float4*>(b + r * ldb) + c4);
      x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    }
    act_store4(out, r, c4 * 4, x);
  }
}

inline int add_rows(Ctx& c, const float* a, long long lda, const float* b, long long ldb, Act out, long long rows, int cols) {
  QAT_REQUIRE(cols % 4 == 0, "add_rows: cols=%d must be a multiple of 4", cols);
  if (c.dry || rows == 0) return QAT_OK;
  long long total = rows * (cols / 4);
  unsigned grid = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  add_rows_kernel<<<grid, 256, 0, c.stream>>>(a, lda, b, ldb, out, rows, cols / 4);
  return check_launch("add_rows_kernel");
}

// fp16 -> fp32 widening of a dense array (n8 groups of 8 elements)
__global__ void __launch_bounds__(256) widen_f16_kernel(const __half* __restrict__ x, float* __restrict__ y, long long n8) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x) + i);
    const float2 a = unpack_half2(u.x), b = unpack_half2(u.y), c = unpack_half2(u.z), d = unpack_half2(u.w);
    reinterpret_cast<float4*>(y)[2 * i] = make_float4(a.x, a.y, b.x, b.y);
    reinterpret_cast<float4*>(y)[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
  }
}
inline int widen_f16(Ctx& c, const __half* x, float* y, long long n) {
  QAT_REQUIRE(n % 8 == 0, "widen_f16: element count must be a multiple of 8");
  if (c.dry || n == 0) return QAT_OK;
  const long long n8 = n / 8;
  unsigned grid = (unsigned)((n8 + 255) / 256 > 148 * 16 ? 148 * 16 : (n8 + 255) / 256);
  widen_f16_kernel<<<grid, 256, 0, c.stream>>>(x, y, n8);
  return check_launch("widen_f16_kernel");
}

// out[b, :] = mean_s x[b, s, :]  (x [B, S, D] contiguous; the sum runs over s in order, then one division)
__global__ void __launch_bounds__(256) mean_seq_kernel(const float* __restrict__ x, float* __restrict__ out, long long B, int S, int D4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D4) return;
  const long long b = i / D4;
  const int c4 = (int)(i % D4);
  const float4* p = reinterpret_cast<const float4*>(x + b * S * D4 * 4) + c4;
  float4 a = __ldg(p);
  for (int s = 1; s < S; ++s) {
    const float4 v = __ldg(p + (long long)s * D4);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  const float inv = (float)S;
  reinterpret_cast<float4*>(out)[i] = make_float4(a.x / inv, a.y / inv, a.z / inv, a.w / inv);
}
inline int mean_seq(Ctx& c, const float* x, float* out, long long B, int S, int D) {
  QAT_REQUIRE(D % 4 == 0 && S >= 1, "mean_seq: D=%d must be a multiple of 4", D);
  if (c.dry || B == 0) return QAT_OK;
  const long long n = B * (D / 4);
  mean_seq_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(x, out, B, S, D / 4);
  return check_launch("mean_seq_kernel");
}

__global__ void __launch_bounds__(256) relu_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaxf(x[i], 0.f);
}
inline int relu_rows(Ctx& c, const float* x, float* y, long long n) {
  if (c.dry || n == 0) return QAT_OK;
  relu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(x, y, n);
  return check_launch("relu_kernel");
}

// x[b*L + l, :] = token_embedding[tokens[b,l], :] + positional_embedding[l, :]   (clip_base.py:390-391)
__global__ void __launch_bounds__(256) clip_embed_kernel(const int* __restrict__ tokens, const float* __restrict__ tok_emb,
                                                         const float* __restrict__ pos_emb, float* __restrict__ x, long long rows, int L,
                                                         int W4, int vocab) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * W4) return;
  const long long r = i / W4;
  const int c4 = (int)(i % W4);
  const int id = tokens[r];
  if (id < 0 || id >= vocab) {
    // nn.Embedding raises on a bad id (a corrupt / mis-typed `quest` tensor).  The library cannot raise from the device without
    // synchronising, so the row is poisoned instead: NaN propagates to that question's features and logits -- loud, never a
    // plausible answer from a clamped id.  Host-side callers validate ids before the call (encoders.py).
    const float q = __int_as_float(0x7fc00000);
    reinterpret_cast<float4*>(x)[i] = make_float4(q, q, q, q);
    return;
  }
  const float4 a = __ldg(reinterpret_cast<const float4*>(tok_emb + (size_t)id * W4 * 4) + c4);
  const float4 b = __ldg(reinterpret_cast<const float4*>(pos_emb + (size_t)(r % L) * W4 * 4) + c4);
  reinterpret_cast<float4*>(x)[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// out[b, :] = x[b*L + argmax_l tokens[b, l], :]  (first maximum, as torch.argmax; clip_base.py:396)
__global__ void __launch_bounds__(256) clip_eot_gather_kernel(const int* __restrict__ tokens, const float* __restrict__ x,
                                                              float* __restrict__ out, int L, int W4) {
  const int b = blockIdx.x;
  __shared__ int s_eot;
  if (threadIdx.x == 0) {
    int best = tokens[(size_t)b * L], bi = 0;
    for (int l = 1; l < L; ++l) {
      const int v = tokens[(size_t)b * L + l];
      if (v > best) { best = v; bi = l; }
    }
    s_eot = bi;
  }
  __syncthreads();
  const float4* src = reinterpret_cast<const float4*>(x) + ((size_t)b * L + s_eot) * W4;
  float4* dst = reinterpret_cast<float4*>(out) + (size_t)b * W4;
  for (int c = threadIdx.x; c < W4; c += blockDim.x) dst[c] = src[c];
}

__global__ void __launch_bounds__(256) ones_kernel(float* __restrict__ y, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = 1.f;
}

}  // namespace qat
