// fp32 FFMA GEMM on CUDA cores: C[M,N] = act(A[M,K] W[N,K]^T + bias) + res.
// This is the QAT_PREC_FP32 path (1e-5 parity gate) and the path of the tiny fp32-only
// contractions (router / gauss_pred / head) in every mode.
#pragma once
#include "qat_common.cuh"

namespace qat {

struct Epi {
  const float* bias = nullptr;  // [N]
  int relu = 0;                 // 0 none, 1 ReLU, 2 QuickGELU x*sigmoid(1.702 x) (clip_base.py:162-164), 3 exact GELU
  const float* res = nullptr;   // [M, ldr] added after the activation
  int ldr = 0;
  const __half* res_h = nullptr;  // the same residual as an fp16 plane (tensor-core GEMMs only; res wins if both are set)
  // tcgen05 kinds only -- the "reduce" epilogue (M4 of SURVEY 8a folded into the expert first-layer GEMM): rows are (sample, t)
  // pairs, T rows per sample; instead of storing act(A W^T + b) the epilogue writes, per sample s and column n,
  //   C.f[s, n] = sum_t coef[(s*T + t) % coef_mod, n / coef_Ch] * act(...)[s*T + t, n]
  // i.e. the Gaussian-weighted temporal reduction of expert n / coef_Ch (coef from moe_coef_kernel)
  const float* coef = nullptr;    // [coef_mod, coef_E]
  int coef_E = 0, coef_T = 0, coef_Ch = 0;
  long long coef_mod = 0;
};

__device__ __forceinline__ float epi_act(float x, int mode) {
  if (mode == 1) return fmaxf(x, 0.f);
  if (mode == 2) return x / (1.f + expf(-1.702f * x));
  if (mode == 3) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));   // nn.GELU() (timm Mlp of the ToMe-patched ViT block)
  return x;
}

// Tile = (8*TY) x (8*TX), BK = 16, TY*TX threads; each thread owns two 4-row and two 4-col strips
// (rows ty*4+i and BM/2+ty*4+i) so that every shared-memory float4 read is conflict-free.
template <int TY, int TX>
__global__ void __launch_bounds__(TY* TX) gemm_simt_kernel(const float* __restrict__ A, int lda,
                                                           const float* __restrict__ W, int ldw, Epi epi, Act C,
                                                           int M, int N, int K) {
  constexpr int BM = 8 * TY, BN = 8 * TX, BK = 16, NT = TY * TX;
  constexpr int PAD = 4;
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Ws[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const long long m0 = (long long)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;

  constexpr int A_LD = BM * 4 / NT;  // float4 loads per thread per k-tile
  constexpr int W_LD = BN * 4 / NT;
  float4 ra[A_LD], rw[W_LD];

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int f = tid + i * NT;
      int r = f >> 2, kq = f & 3;
      long long gr = m0 + r;
      ra[i] = (gr < M) ? __ldg(reinterpret_cast<const float4*>(A + gr * lda + k0 + kq * 4)) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < W_LD; ++i) {
      int f = tid + i * NT;
      int r = f >> 2, kq = f & 3;
      int gn = n0 + r;
      rw[i] = (gn < N) ? __ldg(reinterpret_cast<const float4*>(W + (long long)gn * ldw + k0 + kq * 4)) : make_float4(0, 0, 0, 0);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int f = tid + i * NT;
      int r = f >> 2, kq = f & 3;
      As[buf][kq * 4 + 0][r] = ra[i].x;
      As[buf][kq * 4 + 1][r] = ra[i].y;
      As[buf][kq * 4 + 2][r] = ra[i].z;
      As[buf][kq * 4 + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < W_LD; ++i) {
      int f = tid + i * NT;
      int r = f >> 2, kq = f & 3;
      Ws[buf][kq * 4 + 0][r] = rw[i].x;
      Ws[buf][kq * 4 + 1][r] = rw[i].y;
      Ws[buf][kq * 4 + 2][r] = rw[i].z;
      Ws[buf][kq * 4 + 3][r] = rw[i].w;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = K / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][BM / 2 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][k][BN / 2 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  const bool vec_ok = ((C.ld & 3) == 0) && (epi.res == nullptr || (epi.ldr & 3) == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long gr = m0 + (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
    if (gr >= M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int gc = n0 + (h == 0 ? tx * 4 : BN / 2 + tx * 4);
      if (gc >= N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = acc[i][h * 4 + j];
        if (gc + j < N) {
          if (epi.bias) x += __ldg(epi.bias + gc + j);
          x = epi_act(x, epi.relu);
          if (epi.res) x += __ldg(epi.res + gr * epi.ldr + gc + j);
        }
        v[j] = x;
      }
      if (vec_ok && gc + 3 < N) {
        act_store4(C, gr, gc, make_float4(v[0], v[1], v[2], v[3]));
      } else {
        for (int j = 0; j < 4 && gc + j < N; ++j) act_store1(C, gr, gc + j, v[j]);
      }
    }
  }
}

// Small-M variant (M < 1024: the per-question contractions, M = batch): 32x32 output tiles so N/32 * M/32 CTAs
// cover the SMs; the K loop is split over KG groups of 64 threads inside the CTA, each group streaming its
// k-tiles (g, g+KG, ...) through its own 3-stage cp.async ring (named barrier per group), so an SM keeps
// KG * 2 tiles of global loads in flight instead of one.  Partial tiles meet in shared memory and are summed in
// fixed group order (deterministic), then bias / ReLU / residual and coalesced 16-byte stores.
constexpr int SMALL_KG = 4;
constexpr int SMALL_NS = 3;
constexpr int SMALL_BK = 32;
constexpr int SMALL_LD = 36;   // floats per staged row (32 + 4 pad: rows r, r+1, .. land 4 banks apart)
constexpr int SMALL_SMEM = SMALL_KG * SMALL_NS * 2 * 32 * SMALL_LD * 4;


__global__ void __launch_bounds__(64 * SMALL_KG) gemm_small_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                                                                    Epi epi, Act C, int M, int N, int K) {
  constexpr int NT = 64, LD = SMALL_LD, BK = SMALL_BK, NS = SMALL_NS;
  extern __shared__ __align__(16) float sm_small[];
  const int g = threadIdx.x / NT, tid = threadIdx.x % NT, tx = tid & 7, ty = tid >> 3;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float* gbase = sm_small + g * (NS * 2 * 32 * LD);
  auto As = [&](int s) { return gbase + s * (2 * 32 * LD); };
  auto Ws = [&](int s) { return gbase + s * (2 * 32 * LD) + 32 * LD; };
  const int nk = (K + BK - 1) / BK;
  const int my = (nk - g + SMALL_KG - 1) / SMALL_KG;      // k-tiles g, g+KG, ... of this group
  auto load = [&](int it) {
    const int k0 = (g + it * SMALL_KG) * BK;
    float* as = As(it % NS);
    float* ws = Ws(it % NS);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = tid + i * NT, r = f >> 3, kq = f & 7;
      const bool kin = k0 + kq * 4 < K;
      const bool av = kin && m0 + r < M, wv = kin && n0 + r < N;
      cp_async16_zfill(as + r * LD + kq * 4, av ? A + (long long)(m0 + r) * lda + k0 + kq * 4 : A, av);
      cp_async16_zfill(ws + r * LD + kq * 4, wv ? W + (long long)(n0 + r) * ldw + k0 + kq * 4 : W, wv);
    }
  };
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(NT) : "memory"); };
  // thread (ty, tx) owns rows ty + 8i and columns tx + 8j: its four A rows / W rows sit 4 banks apart per lane group,
  // so the 128-bit reads below are conflict-free
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int s = 0; s < NS - 1; ++s) {
    if (s < my) load(s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int it = 0; it < my; ++it) {
    asm volatile("cp.async.wait_group %0;" ::"n"(NS - 2) : "memory");
    group_sync();
    if (it + NS - 1 < my) load(it + NS - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float* as = As(it % NS);
    const float* ws = Ws(it % NS);
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      float4 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(as + (ty + 8 * i) * LD + k4 * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(ws + (tx + 8 * j) * LD + k4 * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // fixed-order reduction over the K groups through shared memory (each group parks its tile in its own ring)
  __syncthreads();
  float* red = gbase;                                     // [32][33]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[(ty + 8 * i) * 33 + tx + 8 * j] = acc[i][j];
  __syncthreads();
  const int r = threadIdx.x >> 3, c4 = (threadIdx.x & 7) * 4;
  const long long gr = m0 + r;
  const int gc = n0 + c4;
  if (gr >= M || gc >= N) return;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int gg = 0; gg < SMALL_KG; ++gg) {
    const float* rg = sm_small + gg * (NS * 2 * 32 * LD) + r * 33 + c4;
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += rg[j];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (gc + j < N) {
      if (epi.bias) v[j] += __ldg(epi.bias + gc + j);
      v[j] = epi_act(v[j], epi.relu);
      if (epi.res) v[j] += __ldg(epi.res + gr * epi.ldr + gc + j);
    }
  }
  const bool vec_ok = ((C.ld & 3) == 0) && gc + 3 < N;
  if (vec_ok) act_store4(C, gr, gc, make_float4(v[0], v[1], v[2], v[3]));
  else
    for (int j = 0; j < 4 && gc + j < N; ++j) act_store1(C, gr, gc + j, v[j]);
}

inline int gemm_simt(Ctx& c, const float* A, int lda, long long M, const float* W, int N, int K, const Epi& epi, Act C) {
  if (M == 0 || N == 0) return QAT_OK;
  QAT_REQUIRE(K % 4 == 0 && lda % 4 == 0 && (M < 1024 || K % 16 == 0), "gemm_simt: K=%d must be a multiple of 4 (16 for M >= 1024) and lda=%d of 4", K, lda);
  QAT_REQUIRE(M < (1ll << 31), "gemm_simt: M too large");
  if (c.dry) return QAT_OK;
  if (M >= 1024) {
    dim3 grid((N + 127) / 128, (unsigned)((M + 127) / 128));
    QAT_REQUIRE(grid.y <= 65535, "gemm_simt: M=%lld exceeds the grid limit; chunk the rows", M);
    gemm_simt_kernel<16, 16><<<grid, 256, 0, c.stream>>>(A, lda, W, K, epi, C, (int)M, N, K);
  } else {
    dim3 grid((N + 31) / 32, (unsigned)((M + 31) / 32));
    QAT_CUDA(cudaFuncSetAttribute(gemm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_SMEM));
    QAT_REQUIRE((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) % 16 == 0, "gemm_simt: operands must be 16-byte aligned");
    gemm_small_kernel<<<grid, 64 * SMALL_KG, SMALL_SMEM, c.stream>>>(A, lda, W, K, epi, C, (int)M, N, K);
  }
  return check_launch("gemm_simt_kernel");
}

}  // namespace qat
