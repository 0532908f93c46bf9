// The question-aware Gaussian mixture of temporal experts, M2-M4 of SURVEY 8a, as ONE kernel:
//   router softmax over E -> top-K (descending, lowest index on ties) -> renormalise
//   centre/width from gauss_pred (tanh*margin, sigmoid) -> clamp -> Gaussian over t/(T-1)
//   -> divide by the max over t (peak normalisation, modules.py:467-469)
//   -> out[b] = sum_t sum_j coef[t][j] * X[t,b,j,:],  coef[t][j] = sum_{i: perm_t[i]=j} p[i] G[i][t]
// where perm_t is the rank list of sample (b*T+t) mod B -- the `topk_inds.repeat(T,1)` batch-coupling
// quirk of modules.py:502, reproduced on purpose.
//
// HBM-bound: X is read exactly once.  Each (b,t) row (E*C contiguous elements) is pulled by ONE 1-D bulk
// copy (cp.async.bulk -> UBLKCP) into a ring of shared-memory stages guarded by mbarriers; the first RING
// copies are issued before the router/Gaussian prologue so HBM is busy from the first cycle, and the bytes
// in flight per SM (ring depth x row bytes x resident CTAs) do not depend on registers or occupancy.
// The t axis of one sample can be split over a thread-block cluster whose partial sums meet in
// distributed shared memory in rank order (=> bit-reproducible).
#pragma once
#include <cooperative_groups.h>

#include "qat_common.cuh"

namespace qat {
namespace cg = cooperative_groups;

struct GaussP {
  const void* x;
  long long st, sb;   // element strides of X over t and b; (e, c) are contiguous: e*C + c
  const float* rl;    // router logits, row b at rl + b * rl_ld (dense [B,E]: rl_ld = E)
  const float* gr;    // gauss_pred output, row b at gr + b * gr_ld (dense [B,2E]: gr_ld = 2E)
  int rl_ld = 0, gr_ld = 0;   // 0 = dense
  int B, T, E, K, C;
  float sigma;
  int per_expert;     // 0: out[b, c] summed over experts (reference get_output)
                      // 1: out[b, e*C + c] kept per expert and out[b, E*C + e] = sum_t coef[t][e]
                      //    (the M4 refactor: reduce the hidden activations before the 2nd expert GEMM)
  Act out;
  long long* topk_inds;  // optional [B,K]
  float* topk_probs;     // optional [B,K]
  float* gauss_weight;   // optional [B,K,T]
  int ring;              // shared-memory stages (rows in flight per CTA), set by the launcher
};

__device__ __forceinline__ float linspace_at(float start, float end, int steps, int i) {
  // torch.linspace (fp32): symmetric evaluation around the midpoint
  if (steps == 1) return start;
  float step = (end - start) / (float)(steps - 1);
  return (i < steps / 2) ? start + step * (float)i : end - step * (float)(steps - 1 - i);
}

constexpr int GAUSS_MAX_E = 32;

template <typename VecT, int VEC>
__device__ __forceinline__ void fma_vec(float (&acc)[VEC], const VecT& v, float cf);

template <>
__device__ __forceinline__ void fma_vec<float4, 4>(float (&acc)[4], const float4& v, float cf) {
  acc[0] = fmaf(cf, v.x, acc[0]);
  acc[1] = fmaf(cf, v.y, acc[1]);
  acc[2] = fmaf(cf, v.z, acc[2]);
  acc[3] = fmaf(cf, v.w, acc[3]);
}
template <>
__device__ __forceinline__ void fma_vec<uint4, 8>(float (&acc)[8], const uint4& v, float cf) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float lo = __uint_as_float(u[i] << 16);          // bf16 -> fp32 is a shift
    float hi = __uint_as_float(u[i] & 0xffff0000u);
    acc[2 * i] = fmaf(cf, lo, acc[2 * i]);
    acc[2 * i + 1] = fmaf(cf, hi, acc[2 * i + 1]);
  }
}

// shared-memory arrays of the router / Gaussian prologue
struct GaussSm {
  float* coef;          // [T][E]
  float* G;             // [K][T]
  float* p;             // [K] gate of rank i (renormalised)
  float* c;             // [K] centre
  float* w;             // [K] width
  int* idx;             // [K] expert of rank i (own sample)
  unsigned char* perm;  // [T][K] expert of rank i of sample (b*T+t) % B
};

// M2 + M3 for sample b, shared by the streaming kernel below and by moe_coef_kernel (the GEMM-epilogue form of M4):
//   phase 0  rank lists.  One thread per (row t, expert e) -- and per (own sample, e): the rank of e is the number of experts
//            that beat it (higher softmax probability, lower index on ties = torch.topk's order), so no thread walks a
//            top-K selection loop over a locally indexed array (that serial form cost ~10 us per launch)
//   phase 1  Gaussian masks, one warp per selected expert, peak-normalised over t
//   phase 2  coef[t][j] = sum_{i: perm_t[i] = j} p[i] G[i][t]
// Ends with a __syncthreads(): every array of `s` is complete.
template <int NT>
__device__ __forceinline__ void gauss_prologue(const GaussP& p, int b, const GaussSm& s) {
  const int T = p.T, E = p.E, K = p.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rl_ld = p.rl_ld ? p.rl_ld : E;
  for (int item = tid; item < (T + 1) * E; item += NT) {
    const int t = item / E, e = item - t * E;
    const bool own = (t == T);
    const int o = own ? b : (int)(((long long)b * T + t) % p.B);
    const float* l = p.rl + (long long)o * rl_ld;
    float mx = -INFINITY;
    for (int j = 0; j < E; ++j) mx = fmaxf(mx, __ldg(l + j));
    float sum = 0.f;
    for (int j = 0; j < E; ++j) sum += expf(__ldg(l + j) - mx);
    const float pe = expf(__ldg(l + e) - mx) / sum;
    int rank = 0;
    for (int j = 0; j < E; ++j) {
      const float pj = expf(__ldg(l + j) - mx) / sum;
      rank += (pj > pe || (pj == pe && j < e)) ? 1 : 0;
    }
    if (rank < K) {
      if (own) {
        s.idx[rank] = e;
        s.p[rank] = pe;
      } else {
        s.perm[t * K + rank] = (unsigned char)e;
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    // renormalise the K gates (sum in rank order, as modules.py:570) and derive centre / width of the selected experts
    float psum = 0.f;
    for (int i = 0; i < K; ++i) psum += s.p[i];
    __syncwarp();
    if (lane < K) {
      const int e = s.idx[lane];
      s.p[lane] = s.p[lane] / psum;
      const float margin = 1.0f / (float)(E * 2);
      const long long gro = (long long)b * (p.gr_ld ? p.gr_ld : 2 * E);
      const float g0 = __ldg(p.gr + gro + 2 * e);
      const float g1 = __ldg(p.gr + gro + 2 * e + 1);
      const float centre = linspace_at(margin, 1.f - margin, E, e) + tanhf(g0) * margin;
      const float width = 1.f / (1.f + expf(-g1));
      s.c[lane] = fminf(fmaxf(centre, 0.f), 1.f);
      s.w[lane] = fmaxf(width, 0.09f) / p.sigma;
    }
  }
  __syncthreads();
  for (int i = warp; i < K; i += NT / 32) {
    const float c = s.c[i], w = s.w[i];
    const float pref = 0.3989422804014327f / w;
    const float den = 2.f * (w * w);
    float mx = -INFINITY;
    for (int t = lane; t < T; t += 32) {
      float d = linspace_at(0.f, 1.f, T, t) - c;
      float g = pref * expf(-(d * d) / den);
      s.G[i * T + t] = g;
      mx = fmaxf(mx, g);
    }
    mx = warp_max(mx);
    for (int t = lane; t < T; t += 32) s.G[i * T + t] = s.G[i * T + t] / mx;
  }
  __syncthreads();
  for (int t = tid; t < T; t += NT) {
    for (int e = 0; e < E; ++e) s.coef[t * E + e] = 0.f;
    for (int i = 0; i < K; ++i) s.coef[t * E + s.perm[t * K + i]] += s.p[i] * s.G[i * T + t];
  }
  __syncthreads();
}

// optional rank / gate / mask outputs of sample b (after gauss_prologue)
template <int NT>
__device__ __forceinline__ void gauss_write_aux(const GaussP& p, int b, const GaussSm& s) {
  const int tid = threadIdx.x, K = p.K, T = p.T;
  if (p.topk_inds && tid < K) p.topk_inds[(long long)b * K + tid] = s.idx[tid];
  if (p.topk_probs && tid < K) p.topk_probs[(long long)b * K + tid] = s.p[tid];
  if (p.gauss_weight)
    for (int f = tid; f < K * T; f += NT) p.gauss_weight[(long long)b * K * T + f] = s.G[f];
}

// VecT = float4 (fp32 X) or uint4 (bf16 X); SLOTS = vectors per thread per time step.
template <typename VecT, int VEC, int SLOTS, int NT, bool CLUSTER>
__global__ void __launch_bounds__(NT) gauss_reduce_kernel(GaussP p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = p.T, E = p.E, K = p.K, C = p.C;
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int rank = CLUSTER ? blockIdx.y : 0;
  const int nrank = CLUSTER ? gridDim.y : 1;
  const int nvec = E * C / VEC;   // vectors per (b,t) row
  const int cvec = C / VEC;       // vectors per expert

  // ---- shared layout -------------------------------------------------------------------------
  const int ring = p.ring;
  const uint32_t row_bytes = (uint32_t)nvec * sizeof(VecT);
  VecT* s_ring = reinterpret_cast<VecT*>(smem_raw);                                   // [ring][nvec]
  uint64_t* s_full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)ring * row_bytes); // [ring]
  float* s_part = reinterpret_cast<float*>(s_full + ring);       // [nvec*VEC] partial sums (DSMEM-visible)
  GaussSm sm;
  sm.coef = s_part + (size_t)nvec * VEC;                          // [T][E]
  sm.G = sm.coef + T * E;                                         // [K][T]
  sm.p = sm.G + K * T;                                            // [K]
  sm.c = sm.p + K;                                                // [K]
  sm.w = sm.c + K;                                                // [K]
  sm.idx = reinterpret_cast<int*>(sm.w + K);                      // [K]
  sm.perm = reinterpret_cast<unsigned char*>(sm.idx + K);         // [T][K]
  float* s_coef = sm.coef;

  // ---- start streaming before anything else: arm the ring and issue the first `ring` row copies ----
  const int t0 = (int)((long long)T * rank / nrank), t1 = (int)((long long)T * (rank + 1) / nrank);
  const char* xrow0 = reinterpret_cast<const char*>(p.x) + ((long long)b * p.sb) * (long long)(sizeof(VecT) / VEC);
  const long long row_stride = p.st * (long long)(sizeof(VecT) / VEC);   // bytes between consecutive t
  if (tid == 0) {
    for (int s = 0; s < ring; ++s) tc::mbar_init(&s_full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int s = 0; s < ring && t0 + s < t1; ++s) {
      tc::mbar_arrive_expect_tx(&s_full[s], row_bytes);
      tc::bulk_load(s_ring + (size_t)s * nvec, xrow0 + (long long)(t0 + s) * row_stride, row_bytes, &s_full[s]);
    }
  }

  // ---- router / Gaussian prologue (HBM is already streaming) ----------------------------------
  gauss_prologue<NT>(p, b, sm);
  if (rank == 0) gauss_write_aux<NT>(p, b, sm);

  // ---- main loop: consume ring stages in order, refill each stage as soon as the CTA is done with it ----
  float acc[SLOTS][VEC];
  int je[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[s][v] = 0.f;
    int f = tid + s * NT;
    je[s] = f < nvec ? f / cvec : 0;
  }
  for (int t = t0, it = 0; t < t1; ++t, ++it) {
    const int st = it % ring;
    tc::mbar_wait(&s_full[st], (uint32_t)(it / ring) & 1u);
    const VecT* xr = s_ring + (size_t)st * nvec;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      int f = tid + s * NT;
      if (f < nvec) fma_vec<VecT, VEC>(acc[s], xr[f], s_coef[t * E + je[s]]);
    }
    __syncthreads();                       // every thread has read stage `st`
    if (tid == 0 && t + ring < t1) {
      tc::mbar_arrive_expect_tx(&s_full[st], row_bytes);
      tc::bulk_load(s_ring + (size_t)st * nvec, xrow0 + (long long)(t + ring) * row_stride, row_bytes, &s_full[st]);
    }
  }
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    int f = tid + s * NT;
    if (f < nvec) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) s_part[(size_t)f * VEC + v] = acc[s][v];
    }
  }

  // ---- combine the cluster's partial sums in rank order (deterministic) ----------------------
  if (CLUSTER) {
    cg::this_cluster().sync();
  } else {
    __syncthreads();
  }
  {
    // every rank combines an interleaved share of the outputs (always in rank order); four outputs at a time, so that the
    // distributed-shared-memory loads of a batch are in flight together (one output at a time on rank 0 alone cost ~8 us of
    // back-to-back remote-load latency per launch)
    const int nout = p.per_expert ? E * C : C;
    constexpr int U = 4;
    const int stride = NT * nrank;
    for (int o0 = rank * NT + tid; o0 < nout; o0 += stride * U) {
      float sum[U];
#pragma unroll
      for (int u = 0; u < U; ++u) sum[u] = 0.f;
      for (int r = 0; r < nrank; ++r) {
        const float* part = s_part;
        if (CLUSTER) part = cg::this_cluster().map_shared_rank(s_part, r);
        if (p.per_expert) {
          float v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) v[u] = (o0 + u * stride < nout) ? part[o0 + u * stride] : 0.f;
#pragma unroll
          for (int u = 0; u < U; ++u) sum[u] += v[u];
        } else {
          for (int e = 0; e < E; ++e) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = (o0 + u * stride < nout) ? part[e * C + o0 + u * stride] : 0.f;
#pragma unroll
            for (int u = 0; u < U; ++u) sum[u] += v[u];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (o0 + u * stride < nout) act_store1(p.out, b, o0 + u * stride, sum[u]);
    }
  }
  if (rank == 0) {
    if (p.per_expert) {
      for (int e = tid; e < 64; e += NT) {
        float sum = 0.f;
        if (e < E)
          for (int t = 0; t < T; ++t) sum += s_coef[t * E + e];
        act_store1(p.out, b, E * C + e, sum);
      }
    }
  }
  if (CLUSTER) cg::this_cluster().sync();   // keep every CTA's smem alive until all ranks have read it
}

// The GEMM-epilogue form of M4: the expert first-layer GEMM multiplies its relu'd tile by coef[b,t,e] and sums
// over the rows of a sample in its epilogue (gemm_tc.cuh, EPI = 4), so the hidden activations [T,B,E,C/2] never reach HBM.  This
// kernel provides what that epilogue needs: coef[b*T + t][e] for every row of the batch, the bias column block
// hb[b, E*Ch + e] = sum_t coef[b,t,e] (e < 64, zero beyond E) of the second-layer contraction -- in `nrep` row blocks of hb,
// B rows apart (the two vt branches share one router) -- and the optional rank / gate / mask outputs.  One CTA per sample.
struct MoeCoefP {
  GaussP g;          // rl / gr / B / T / E / K / sigma / aux outputs (x, out unused)
  float* coef;       // [B*T, E]
  float* hb;         // [nrep*B, ldh]
  bf16* hbh = nullptr;   // the same block as bf16 (hi, lo) planes (optional: the split-K tensor-core second layer reads those)
  bf16* hbl = nullptr;
  int ldh, col0, nrep;
};
struct MoeCoefPack {
  MoeCoefP m[2];     // blockIdx.y selects the module (at_aggregator / vt_aggregator share a launch)
};

constexpr int MOE_COEF_NT = 128;

__global__ void __launch_bounds__(MOE_COEF_NT) moe_coef_kernel(MoeCoefPack pk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const MoeCoefP& q = pk.m[blockIdx.y];
  const GaussP& p = q.g;
  const int T = p.T, E = p.E, K = p.K, tid = threadIdx.x, b = blockIdx.x;
  GaussSm sm;
  sm.coef = reinterpret_cast<float*>(smem_raw);
  sm.G = sm.coef + T * E;
  sm.p = sm.G + K * T;
  sm.c = sm.p + K;
  sm.w = sm.c + K;
  sm.idx = reinterpret_cast<int*>(sm.w + K);
  sm.perm = reinterpret_cast<unsigned char*>(sm.idx + K);
  gauss_prologue<MOE_COEF_NT>(p, b, sm);
  gauss_write_aux<MOE_COEF_NT>(p, b, sm);
  for (int f = tid; f < T * E; f += MOE_COEF_NT) q.coef[(long long)b * T * E + f] = sm.coef[f];
  for (int e = tid; e < 64; e += MOE_COEF_NT) {
    float sum = 0.f;
    if (e < E)
      for (int t = 0; t < T; ++t) sum += sm.coef[t * E + e];
    bf16 sh, sl;
    split_bf16(sum, sh, sl);
    for (int r = 0; r < q.nrep; ++r) {
      const long long o = ((long long)r * p.B + b) * q.ldh + q.col0 + e;
      q.hb[o] = sum;
      if (q.hbh) {
        q.hbh[o] = sh;
        q.hbl[o] = sl;
      }
    }
  }
}

// n modules (1 or 2) with the same B / T / E / K in one launch
inline int moe_coef(Ctx& c, const MoeCoefPack& pk, int n) {
  const GaussP& p = pk.m[0].g;
  QAT_REQUIRE(n >= 1 && n <= 2, "moe_coef: 1 or 2 modules per launch");
  QAT_REQUIRE(p.E >= 1 && p.E <= GAUSS_MAX_E && p.K >= 1 && p.K <= p.E && p.T >= 1 && p.T <= 4096 && p.B >= 1, "moe_coef: bad shape");
  for (int m = 1; m < n; ++m)
    QAT_REQUIRE(pk.m[m].g.B == p.B && pk.m[m].g.T == p.T && pk.m[m].g.E == p.E && pk.m[m].g.K == p.K, "moe_coef: modules of one launch must share B / T / E / K");
  if (c.dry) return QAT_OK;
  const size_t smem = ((size_t)p.T * p.E + (size_t)p.K * p.T + 3 * (size_t)p.K) * 4 + (size_t)p.K * 4 + (size_t)p.T * p.K + 16;
  QAT_REQUIRE(smem <= 200 * 1024, "moe_coef: T*E too large for one CTA");
  if (smem > 48 * 1024) QAT_CUDA(cudaFuncSetAttribute(moe_coef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  moe_coef_kernel<<<dim3(p.B, n), MOE_COEF_NT, smem, c.stream>>>(pk);
  return check_launch("moe_coef_kernel");
}

inline size_t gauss_smem_bytes(const GaussP& p, size_t row_bytes) {
  size_t f = (size_t)p.E * p.C + (size_t)p.T * p.E + (size_t)p.K * p.T + 3 * (size_t)p.K;
  return (size_t)p.ring * (row_bytes + 8) + f * 4 + (size_t)p.K * 4 + (size_t)p.T * p.K + 16;
}

template <typename VecT, int VEC, int SLOTS, int NT>
inline int gauss_launch(Ctx& c, const GaussP& p_in, int tsplit) {
  GaussP p = p_in;
  const size_t row_bytes = (size_t)p.E * p.C * (sizeof(VecT) / VEC);
  // ring depth: all CTAs should be co-resident (one wave), so the per-CTA shared-memory budget is the SM's
  // 227 KB divided by the CTAs each SM has to host; 2..8 stages, never more than the slice length
  const long long ctas = (long long)p.B * tsplit;
  int per_sm = (int)((ctas + 147) / 148);
  per_sm = per_sm > 8 ? 8 : per_sm;
  p.ring = 0;
  const size_t fixed = gauss_smem_bytes(p, row_bytes) + 1024;
  const size_t budget = (size_t)(225 * 1024) / per_sm;
  int ring = budget > fixed ? (int)((budget - fixed) / (row_bytes + 8)) : 2;
  const int slice = (p.T + tsplit - 1) / tsplit;
  ring = ring > 8 ? 8 : ring;
  ring = ring > slice ? slice : ring;
  ring = ring < 2 ? 2 : ring;
  p.ring = ring;
  size_t smem = gauss_smem_bytes(p, row_bytes);
  QAT_REQUIRE(smem <= 220 * 1024, "gauss_reduce: row of %zu bytes does not fit the shared-memory ring", row_bytes);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.B, tsplit, 1);
  cfg.blockDim = dim3(NT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c.stream;
  cudaLaunchAttribute attr[1];
  if (tsplit > 1) {
    auto kern = gauss_reduce_kernel<VecT, VEC, SLOTS, NT, true>;
    QAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = tsplit;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    QAT_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  } else {
    auto kern = gauss_reduce_kernel<VecT, VEC, SLOTS, NT, false>;
    QAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QAT_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  }
  return check_launch("gauss_reduce_kernel");
}

// dtype: QAT_DT_F32 / QAT_DT_BF16.  tsplit: CTAs per sample (cluster size), 0 = auto.
inline int gauss_reduce(Ctx& c, const GaussP& p, int dtype, int tsplit) {
  QAT_REQUIRE(p.E >= 1 && p.E <= GAUSS_MAX_E && p.K >= 1 && p.K <= p.E, "gauss_reduce: need 1<=K<=E<=%d (E=%d K=%d)", GAUSS_MAX_E, p.E, p.K);
  QAT_REQUIRE(p.T >= 1 && p.T <= 4096 && p.B >= 1, "gauss_reduce: bad B/T");
  const int VEC = dtype == QAT_DT_BF16 ? 8 : 4;
  QAT_REQUIRE(dtype == QAT_DT_F32 || dtype == QAT_DT_BF16, "gauss_reduce: unknown dtype %d", dtype);
  QAT_REQUIRE(p.C % VEC == 0 && p.st % VEC == 0 && p.sb % VEC == 0, "gauss_reduce: C and strides must be multiples of %d elements", VEC);
  QAT_REQUIRE(reinterpret_cast<uintptr_t>(p.x) % 16 == 0, "gauss_reduce: experts tensor must be 16-byte aligned");
  QAT_REQUIRE((size_t)p.E * p.C <= 16384, "gauss_reduce: E*C=%d exceeds 16384 (one CTA holds a full (b,t) row in flight)", p.E * p.C);
  if (c.dry) return QAT_OK;
  if (tsplit <= 0) {
    // measured at B=256,T=60,E=7 (CUDA-graph replay): t_split 1 -> 4.44 TB/s, 2 -> 4.75 TB/s, 4 -> 3.3 TB/s:
    // ~4 co-resident CTAs per SM with a short ring beat 2 CTAs with a deep one; cluster size <= 8, slices >= ~8 steps
    tsplit = 1;
    while (tsplit < 8 && (long long)p.B * tsplit * 2 <= 148 * 4 && p.T / (tsplit * 2) >= 8) tsplit *= 2;
  }
  QAT_REQUIRE(tsplit >= 1 && tsplit <= 8, "gauss_reduce: t_split=%d outside [1,8]", tsplit);
  const int nvec = p.E * p.C / VEC;
  // SLOTS*NT >= nvec
  if (dtype == QAT_DT_F32) {
    if (nvec <= 4 * 128) return gauss_launch<float4, 4, 4, 128>(c, p, tsplit);
    if (nvec <= 8 * 128) return gauss_launch<float4, 4, 8, 128>(c, p, tsplit);
    if (nvec <= 8 * 256) return gauss_launch<float4, 4, 8, 256>(c, p, tsplit);
    return gauss_launch<float4, 4, 16, 256>(c, p, tsplit);
  } else {
    if (nvec <= 4 * 128) return gauss_launch<uint4, 8, 4, 128>(c, p, tsplit);
    if (nvec <= 8 * 128) return gauss_launch<uint4, 8, 8, 128>(c, p, tsplit);
    return gauss_launch<uint4, 8, 8, 256>(c, p, tsplit);
  }
}

}  // namespace qat
