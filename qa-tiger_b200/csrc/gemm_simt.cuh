// fp32 FFMA GEMM on CUDA cores: C[M,N] = act(A[M,K] W[N,K]^T + bias) + res.
// This is the QAT_PREC_FP32 path (1e-5 parity gate) and the path of the tiny fp32-only
// contractions (router / gauss_pred / head) in every mode.
#pragma once
#include "qat_common.cuh"

namespace qat {

struct Epi {
  const float* bias = nullptr;  // [N]
  int relu = 0;
  const float* res = nullptr;   // [M, ldr] added after the activation
  int ldr = 0;
  const __half* res_h = nullptr;  // the same residual as an fp16 plane (tensor-core GEMMs only; res wins if both are set)
};

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
          if (epi.relu) x = fmaxf(x, 0.f);
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
// cover the SMs, and the K loop is split over KG groups of 64 threads inside the CTA (named barriers per group,
// fixed-order reduction through shared memory => deterministic) so each SM holds 8 warps of independent work.
constexpr int SMALL_KG = 4;
__global__ void __launch_bounds__(64 * SMALL_KG) gemm_small_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                                                                    Epi epi, Act C, int M, int N, int K) {
  constexpr int BM = 32, BN = 32, BK = 16, NT = 64, PAD = 4;
  __shared__ __align__(16) float As[SMALL_KG][2][BK][BM + PAD];
  __shared__ __align__(16) float Ws[SMALL_KG][2][BK][BN + PAD];
  const int g = threadIdx.x / NT, tid = threadIdx.x % NT, tx = tid & 7, ty = tid >> 3;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float4 ra[2], rw[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = tid + i * NT, r = f >> 2, kq = f & 3;
      const bool kin = k0 + kq * 4 < K;
      ra[i] = (m0 + r < M && kin) ? __ldg(reinterpret_cast<const float4*>(A + (long long)(m0 + r) * lda + k0 + kq * 4)) : make_float4(0, 0, 0, 0);
      rw[i] = (n0 + r < N && kin) ? __ldg(reinterpret_cast<const float4*>(W + (long long)(n0 + r) * ldw + k0 + kq * 4)) : make_float4(0, 0, 0, 0);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = tid + i * NT, r = f >> 2, kq = f & 3;
      As[g][buf][kq * 4 + 0][r] = ra[i].x; As[g][buf][kq * 4 + 1][r] = ra[i].y; As[g][buf][kq * 4 + 2][r] = ra[i].z; As[g][buf][kq * 4 + 3][r] = ra[i].w;
      Ws[g][buf][kq * 4 + 0][r] = rw[i].x; Ws[g][buf][kq * 4 + 1][r] = rw[i].y; Ws[g][buf][kq * 4 + 2][r] = rw[i].z; Ws[g][buf][kq * 4 + 3][r] = rw[i].w;
    }
  };
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(NT) : "memory"); };
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int nk = (K + BK - 1) / BK;
  const int my = (nk - g + SMALL_KG - 1) / SMALL_KG;      // k-tiles g, g+KG, ... of this group
  if (my > 0) {
    gload(g * BK);
    sstore(0);
    group_sync();
    for (int it = 0; it < my; ++it) {
      const int buf = it & 1;
      if (it + 1 < my) gload((g + (it + 1) * SMALL_KG) * BK);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[g][buf][k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Ws[g][buf][k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (it + 1 < my) {
        sstore(buf ^ 1);
        group_sync();
      }
    }
  }
  // fixed-order reduction over the K groups: groups 1.. park their tiles in their own (now idle) A buffers
  __syncthreads();
  if (g > 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(&As[g][0][0][0] + (ty * 4 + i) * 32 + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  __syncthreads();
  if (g > 0) return;
#pragma unroll
  for (int gg = 1; gg < SMALL_KG; ++gg)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(&As[gg][0][0][0] + (ty * 4 + i) * 32 + tx * 4);
      acc[i][0] += t.x; acc[i][1] += t.y; acc[i][2] += t.z; acc[i][3] += t.w;
    }
  const bool vec_ok = ((C.ld & 3) == 0) && (epi.res == nullptr || (epi.ldr & 3) == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long gr = m0 + ty * 4 + i;
    const int gc = n0 + tx * 4;
    if (gr >= M || gc >= N) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = acc[i][j];
      if (gc + j < N) {
        if (epi.bias) x += __ldg(epi.bias + gc + j);
        if (epi.relu) x = fmaxf(x, 0.f);
        if (epi.res) x += __ldg(epi.res + gr * epi.ldr + gc + j);
      }
      v[j] = x;
    }
    if (vec_ok && gc + 3 < N) act_store4(C, gr, gc, make_float4(v[0], v[1], v[2], v[3]));
    else
      for (int j = 0; j < 4 && gc + j < N; ++j) act_store1(C, gr, gc + j, v[j]);
  }
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
    gemm_small_kernel<<<grid, 64 * SMALL_KG, 0, c.stream>>>(A, lda, W, K, epi, C, (int)M, N, K);
  }
  return check_launch("gemm_simt_kernel");
}

}  // namespace qat
