// The question-aware Gaussian mixture of temporal experts, M2-M4 of SURVEY 8a, as ONE kernel:
//   router softmax over E -> top-K (descending, lowest index on ties) -> renormalise
//   centre/width from gauss_pred (tanh*margin, sigmoid) -> clamp -> Gaussian over t/(T-1)
//   -> divide by the max over t (peak normalisation, modules.py:467-469)
//   -> out[b] = sum_t sum_j coef[t][j] * X[t,b,j,:],  coef[t][j] = sum_{i: perm_t[i]=j} p[i] G[i][t]
// where perm_t is the rank list of sample (b*T+t) mod B -- the `topk_inds.repeat(T,1)` batch-coupling
// quirk of modules.py:502, reproduced on purpose.
//
// HBM-bound: X is read exactly once, by TMA box loads (cp.async.bulk.tensor.4d) into a ring of shared-memory stages guarded
// by mbarriers; the first stages are issued before the router/Gaussian prologue so HBM is busy from the first cycle, and
// the bytes in flight per SM (ring bytes x resident CTAs) do not depend on registers or occupancy.  The CTAs of one sample
// cut the channel axis, so every output element is one thread's sum over t in ascending order (=> bit-reproducible,
// independent of the slice count).
#pragma once
#include "qat_common.cuh"

namespace qat {

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
  // set by the launcher:
  int ring = 0;          // shared-memory stages in flight per CTA
  int tg = 0;            // time steps per stage
  int split = 0;         // CTAs per sample (column slices)
  int cw = 0;            // 16-byte vectors per expert and slice
  int csz = 1;           // CTAs sharing one prologue = cluster size (split, or 1 when split > 8)
  int t_outer = 0;       // 1: tensor-map axes (c, e, b, t); 0: (c, e, t, b)
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
  float* prob;          // [T+1][E] router logits, then probabilities, of the T borrowed samples and (row T) of the own sample
  float* graw;          // [2E] gauss_pred row of the own sample
  float* p;             // [K] gate of rank i as the softmax gives it
  float* pn;            // [K] gate of rank i, renormalised over the K selected experts
  int* idx;             // [K] expert of rank i (own sample)
  unsigned char* rk;    // [T+1][E] rank of expert e in sample (b*T+t) % B (row T: own sample)
};
// bytes of the arrays above (a multiple of 16)
__host__ __device__ inline size_t gauss_sm_bytes(int T, int E, int K) {
  return ((size_t)T * E + (size_t)K * T + (size_t)(T + 1) * E + 2 * (size_t)E + 3 * (size_t)K) * 4 + (((size_t)(T + 1) * E + 15) & ~(size_t)15);
}
__device__ __forceinline__ GaussSm gauss_sm_carve(unsigned char* base, int T, int E, int K) {
  GaussSm s;
  s.coef = reinterpret_cast<float*>(base);
  s.G = s.coef + T * E;
  s.prob = s.G + K * T;
  s.graw = s.prob + (T + 1) * E;
  s.p = s.graw + 2 * E;
  s.pn = s.p + K;
  s.idx = reinterpret_cast<int*>(s.pn + K);
  s.rk = reinterpret_cast<unsigned char*>(s.idx + K);
  return s;
}

// M2 + M3 for sample b, shared by the streaming kernel below and by moe_coef_kernel (the GEMM-epilogue form of M4).  The
// prologue sits in front of every CTA's stream, so it is laid out for latency: every global load of the call is issued in
// phase A (one element per thread, one round trip), the phases between the barriers are a few hundred cycles each.
//   A  router logits of the T borrowed samples (b*T+t) % B and of the own sample, gauss_pred row of the own sample -> shared
//   B  router softmax, one thread per row
//   C  ranks, one thread per (row, expert): the rank of e is the number of experts that beat it (higher probability, lower
//      index on ties = torch.topk's order) -- no thread walks a top-K selection loop (that serial form cost ~10 us)
//   D  one warp per selected expert: renormalised gate (sum in rank order, modules.py:570), centre / width, Gaussian mask,
//      peak-normalised over t
//   E  coef[t][e] = p[r] G[r][t] with r = rank of e in the row-t sample (0 beyond the top K), one thread per (t, e)
// Ends with a __syncthreads(): every array of `s` is complete.
template <int NT>
__device__ __forceinline__ void gauss_prologue(const GaussP& p, int b, const GaussSm& s) {
  const int T = p.T, E = p.E, K = p.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rl_ld = p.rl_ld ? p.rl_ld : E;
  for (int item = tid; item < (T + 1) * E; item += NT) {
    const int t = item / E, e = item - t * E;
    const int o = (t == T) ? b : (int)(((long long)b * T + t) % p.B);
    s.prob[item] = __ldg(p.rl + (long long)o * rl_ld + e);
  }
  for (int i = tid; i < 2 * E; i += NT) s.graw[i] = __ldg(p.gr + (long long)b * (p.gr_ld ? p.gr_ld : 2 * E) + i);
  __syncthreads();
  for (int t = tid; t <= T; t += NT) {
    float* l = s.prob + t * E;
    float mx = -INFINITY;
    for (int j = 0; j < E; ++j) mx = fmaxf(mx, l[j]);
    float sum = 0.f;
    for (int j = 0; j < E; ++j) sum += expf(l[j] - mx);
    for (int j = 0; j < E; ++j) l[j] = expf(l[j] - mx) / sum;
  }
  __syncthreads();
  for (int item = tid; item < (T + 1) * E; item += NT) {
    const int t = item / E, e = item - t * E;
    const float* pr = s.prob + t * E;
    const float pe = pr[e];
    int rank = 0;
    for (int j = 0; j < E; ++j) {
      const float pj = pr[j];
      rank += (pj > pe || (pj == pe && j < e)) ? 1 : 0;
    }
    s.rk[item] = (unsigned char)rank;
    if (t == T && rank < K) {
      s.idx[rank] = e;
      s.p[rank] = pe;
    }
  }
  __syncthreads();
  for (int i = warp; i < K; i += NT / 32) {
    float psum = 0.f;
    for (int k = 0; k < K; ++k) psum += s.p[k];
    const int e = s.idx[i];
    const float margin = 1.0f / (float)(E * 2);
    const float centre = linspace_at(margin, 1.f - margin, E, e) + tanhf(s.graw[2 * e]) * margin;
    const float width = 1.f / (1.f + expf(-s.graw[2 * e + 1]));
    const float c = fminf(fmaxf(centre, 0.f), 1.f);
    const float w = fmaxf(width, 0.09f) / p.sigma;
    if (lane == 0) s.pn[i] = s.p[i] / psum;
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
  for (int item = tid; item < T * E; item += NT) {
    const int t = item / E;
    const int r = s.rk[item];
    s.coef[item] = r < K ? s.pn[r] * s.G[r * T + t] : 0.f;
  }
  __syncthreads();
}

// optional rank / gate / mask outputs of sample b (after gauss_prologue)
template <int NT>
__device__ __forceinline__ void gauss_write_aux(const GaussP& p, int b, const GaussSm& s) {
  const int tid = threadIdx.x, K = p.K, T = p.T;
  if (p.topk_inds && tid < K) p.topk_inds[(long long)b * K + tid] = s.idx[tid];
  if (p.topk_probs && tid < K) p.topk_probs[(long long)b * K + tid] = s.pn[tid];
  if (p.gauss_weight)
    for (int f = tid; f < K * T; f += NT) p.gauss_weight[(long long)b * K * T + f] = s.G[f];
}

// ---- the prologue of the streaming kernel: M2 + M3 ONCE per sample, shared by the CTAs of its cluster -------------------------
// Every CTA of a sample needs coef[t][e] for all t, and with one prologue per CTA the prologue was the kernel's largest
// fixed cost: ~7 300 warp instructions per CTA, issue-bound at seven CTAs per SM (6-8 us of a 27 us launch; with the
// prologue switched off the same stream ran at 0.83 / 0.96 of the HBM peak).  Now the CTA of cluster rank r handles the rows
// t = r, r + S, ... (router softmax and ranks of the borrowed sample, coef of that row) and stores its rows of coef into
// the shared memory of every CTA of the cluster (st.shared::cluster), one cluster barrier later all of them hold [T][E].
// The own sample's top-K (one softmax row, K gates / centres / widths) is recomputed by every CTA -- it is tiny -- and the
// peak of a Gaussian mask is taken over the grid points next to its centre instead of over all T (the mask is unimodal;
// when the grid is so fine against the width that neighbouring values could tie within rounding, all T are scanned), so no
// [K][T] mask array is built: coef[t][e] = pn[r] * (g_r(t) / peak_r) evaluates the same expressions, bit for bit, as the
// array form (gauss_prologue above, kept for moe_coef_kernel).
struct GaussCl {
  float* coef;          // [T][E] (complete after the cluster barrier)
  float* lg;            // [rows+1][E] router logits: this CTA's borrowed rows, then the own sample
  float* prob;          // [rows+1][E] softmax numerators exp(l - max) of the same rows
  float* graw;          // [2E]
  float *p, *pn, *c, *pref, *den, *mx;   // [K] each
  int* idx;             // [K]
  unsigned char* rk;    // [rows+1][E]
};
__host__ __device__ inline int gauss_cl_rows(int T, int csz) { return (T + csz - 1) / csz; }
__host__ __device__ inline size_t gauss_cl_bytes(int T, int E, int K, int csz) {
  const size_t rows = (size_t)gauss_cl_rows(T, csz) + 1;
  return ((size_t)T * E + 2 * rows * E + 2 * (size_t)E + 7 * (size_t)K) * 4 + ((rows * E + 15) & ~(size_t)15);
}
__device__ __forceinline__ GaussCl gauss_cl_carve(unsigned char* base, int T, int E, int K, int csz) {
  const int rows = gauss_cl_rows(T, csz) + 1;
  GaussCl s;
  s.coef = reinterpret_cast<float*>(base);
  s.lg = s.coef + T * E;
  s.prob = s.lg + rows * E;
  s.graw = s.prob + rows * E;
  s.p = s.graw + 2 * E;
  s.pn = s.p + K;
  s.c = s.pn + K;
  s.pref = s.c + K;
  s.den = s.pref + K;
  s.mx = s.den + K;
  s.idx = reinterpret_cast<int*>(s.mx + K);
  s.rk = reinterpret_cast<unsigned char*>(s.idx + K);
  return s;
}
__device__ __forceinline__ float gauss_at(float pref, float den, float c, int T, int t) {
  const float d = linspace_at(0.f, 1.f, T, t) - c;
  return pref * expf(-(d * d) / den);
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// cr = rank of this CTA among the csz CTAs that share the prologue.
// Phase A, by the first NTL threads of the CTA: every global load of the prologue, one element per thread and one round trip.
// Called BEFORE the ring's first TMA loads are issued: behind ~150 KB of bulk traffic per SM these few loads waited for
// microseconds.
template <int NTL>
__device__ __forceinline__ void gauss_prologue_load(const GaussP& p, int b, int cr, int csz, const GaussCl& s) {
  const int T = p.T, E = p.E;
  const int tid = threadIdx.x;
  if (tid >= NTL) return;
  const int rl_ld = p.rl_ld ? p.rl_ld : E;
  const int nmy = cr < T ? (T - cr + csz - 1) / csz : 0;   // borrowed rows of this CTA: t = cr + k*csz; local row nmy = own sample
  const bool small = (long long)p.B * T < (1ll << 31);
  const int nitem = (nmy + 1) * E;
  auto fetch = [&](int item) {
    const int k = item / E, e = item - k * E;
    const int t = cr + k * csz;
    const int o = (k == nmy) ? b : (small ? (int)((unsigned)(b * T + t) % (unsigned)p.B) : (int)(((long long)b * T + t) % p.B));
    return __ldg(p.rl + (long long)o * rl_ld + e);
  };
  // first element per thread: issued, the producer warp released (named barrier 2), then stored
  const float v0 = tid < nitem ? fetch(tid) : 0.f;
  const float g0 = tid < 2 * E ? __ldg(p.gr + (long long)b * (p.gr_ld ? p.gr_ld : 2 * E) + tid) : 0.f;
  asm volatile("bar.arrive 2, %0;" ::"n"(NTL + 32) : "memory");
  if (tid < nitem) s.lg[tid] = v0;
  if (tid < 2 * E) s.graw[tid] = g0;
  for (int item = tid + NTL; item < nitem; item += NTL) s.lg[item] = fetch(item);
}

// The rest of the prologue, by all NT threads.  The caller has executed cluster_arrive() once at kernel entry (every CTA of
// the cluster is known to run before its shared memory is written).  Ends with every array complete.
template <int NT>
__device__ __forceinline__ void gauss_prologue_cluster(const GaussP& p, int b, int cr, int csz, const GaussCl& s) {
  const int T = p.T, E = p.E, K = p.K;
  const int tid = threadIdx.x;
  const int nmy = cr < T ? (T - cr + csz - 1) / csz : 0;
  __syncthreads();
  // softmax numerators, one thread per (row, expert)
  const int nitem = (nmy + 1) * E;
  for (int item = tid; item < nitem; item += NT) {
    const int k = item / E;
    const float* l = s.lg + k * E;
    float mx = -INFINITY;
    for (int j = 0; j < E; ++j) mx = fmaxf(mx, l[j]);
    s.prob[item] = expf(l[item - k * E] - mx);
  }
  __syncthreads();
  // probabilities (sum in index order, as one thread per row would) and the rank of every expert in its row
  for (int item = tid; item < nitem; item += NT) {
    const int k = item / E, e = item - k * E;
    const float* ex = s.prob + k * E;
    float sum = 0.f;
    for (int j = 0; j < E; ++j) sum += ex[j];
    const float pe = ex[e] / sum;
    int rank = 0;
    for (int j = 0; j < E; ++j) {
      const float pj = ex[j] / sum;
      rank += (pj > pe || (pj == pe && j < e)) ? 1 : 0;
    }
    s.rk[item] = (unsigned char)rank;
    if (k == nmy && rank < K) {
      s.idx[rank] = e;
      s.p[rank] = pe;
    }
  }
  __syncthreads();
  // selected experts of the own sample: four lanes per expert evaluate the mask at the grid points around its centre
  for (int i0 = 0; i0 < K; i0 += NT / 4) {
    const int i = i0 + (tid >> 2), q = tid & 3;
    const bool act = i < K;
    float c = 0.f, pref = 0.f, den = 1.f, mx = -INFINITY, psum = 0.f;
    bool scan = false;
    if (act) {
      for (int k = 0; k < K; ++k) psum += s.p[k];
      const int e = s.idx[i];
      const float margin = 1.0f / (float)(E * 2);
      const float centre = linspace_at(margin, 1.f - margin, E, e) + tanhf(s.graw[2 * e]) * margin;
      const float width = 1.f / (1.f + expf(-s.graw[2 * e + 1]));
      c = fminf(fmaxf(centre, 0.f), 1.f);
      const float w = fmaxf(width, 0.09f) / p.sigma;
      pref = 0.3989422804014327f / w;
      den = 2.f * (w * w);
      const float step = T > 1 ? 1.f / (float)(T - 1) : 1.f;
      scan = !(step * step > 1e-4f * den);
      if (!scan) {
        const int t = (int)floorf(c * (float)(T - 1)) - 1 + q;
        if (t >= 0 && t < T) mx = gauss_at(pref, den, c, T, t);
      } else {
        for (int t = q; t < T; t += 4) mx = fmaxf(mx, gauss_at(pref, den, c, T, t));
      }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    if (act && q == 0) {
      s.pn[i] = s.p[i] / psum;
      s.c[i] = c;
      s.pref[i] = pref;
      s.den[i] = den;
      s.mx[i] = mx;
    }
  }
  __syncthreads();
  cluster_wait();     // every CTA of the cluster has started
  const uint32_t coef_sa = tc::smem_u32(s.coef);
  for (int item = tid; item < nmy * E; item += NT) {
    const int k = item / E, e = item - k * E;
    const int t = cr + k * csz;
    const int r = s.rk[item];
    const float v = r < K ? s.pn[r] * (gauss_at(s.pref[r], s.den[r], s.c[r], T, t) / s.mx[r]) : 0.f;
    const uint32_t la = coef_sa + (uint32_t)(t * E + e) * 4u;
    for (int q = 0; q < csz; ++q) {
      uint32_t ra;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(q));
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
    }
  }
  __syncwarp();
  cluster_arrive();
  cluster_wait();
}

// optional rank / gate / mask outputs of sample b (after gauss_prologue_cluster)
template <int NT>
__device__ __forceinline__ void gauss_write_aux_cluster(const GaussP& p, int b, const GaussCl& s) {
  const int tid = threadIdx.x, K = p.K, T = p.T;
  if (p.topk_inds && tid < K) p.topk_inds[(long long)b * K + tid] = s.idx[tid];
  if (p.topk_probs && tid < K) p.topk_probs[(long long)b * K + tid] = s.pn[tid];
  if (p.gauss_weight)
    for (int f = tid; f < K * T; f += NT) {
      const int i = f / T, t = f - i * T;
      p.gauss_weight[(long long)b * K * T + f] = gauss_at(s.pref[i], s.den[i], s.c[i], T, t) / s.mx[i];
    }
}

// The streaming kernel.  VecT = float4 (fp32 X) or uint4 (bf16 X); SLOTS = 16-byte vectors per consumer thread per time step.
//
// One CTA per (sample b, column slice s): the S slices of a sample cut the channel axis, so a CTA owns `cw` 16-byte vectors of
// every expert for ALL T steps -- no partial sums ever cross a CTA boundary (the first version cut the t axis over a cluster
// and met in distributed shared memory), B*S CTAs spread evenly over the SMs (S = 4 at B = 256: 6.9 CTAs per SM; the
// two-CTA clusters left 3 or 4 per SM, a 16 % imbalance), and every sum over t is one thread's left-to-right chain.
// X is one 4-D tensor map (c, e, then b / t in stride order); a stage of the shared-memory ring = the box [TG steps][E][cw] of
// one TMA load.  Warp 4 is the producer (first `ring` stages in flight before the prologue runs, then one refill per
// `empty` completion), warps 0-3 consume: wait `full`, FMA from shared memory, one `empty` arrival per warp -- no CTA-wide
// barrier in the loop.
constexpr int GAUSS_NC = 128;          // consumer threads
constexpr int GAUSS_NT = GAUSS_NC + 32;

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(tc::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <typename VecT, int VEC, int SLOTS>
__global__ void __launch_bounds__(GAUSS_NT) gauss_reduce_kernel(const __grid_constant__ CUtensorMap mX, GaussP p) {
  extern __shared__ __align__(128) unsigned char gauss_smem[];
  const int T = p.T, E = p.E, C = p.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = p.split, TG = p.tg, ring = p.ring;
  const int b = blockIdx.x / S, s = blockIdx.x - b * S;
  const int cvec = C / VEC;          // vectors per expert
  const int cw = p.cw;               // vectors per expert owned by this CTA (the last slice may be clipped)
  const int nloc = E * cw;           // vectors per time step in this CTA's box
  const int nst = (T + TG - 1) / TG; // stages to stream

  // ---- shared layout: ring | barriers | prologue arrays ----------------------------------------------
  const uint32_t stage_bytes = (uint32_t)TG * nloc * sizeof(VecT);                              // bytes one box load delivers
  const uint32_t svec = (stage_bytes + 127u) / 128u * 128u / (uint32_t)sizeof(VecT);            // stage pitch (TMA: 128-byte aligned)
  VecT* s_ring = reinterpret_cast<VecT*>(gauss_smem);                                             // [ring][TG][E][cw]
  uint64_t* s_full = reinterpret_cast<uint64_t*>(gauss_smem + (size_t)ring * svec * sizeof(VecT)); // [ring]
  uint64_t* s_empty = s_full + ring;                                                            // [ring]
  const int csz = p.csz, cr = csz > 1 ? s : 0;   // CTAs that share one prologue (the sample's cluster), rank among them
  const GaussCl sm = gauss_cl_carve(reinterpret_cast<unsigned char*>(s_empty + ring), T, E, p.K, csz);
  const float* s_coef = sm.coef;
  cluster_arrive();
  gauss_prologue_load<GAUSS_NC>(p, b, cr, csz, sm);

  // ---- start streaming before anything else ------------------------------------------------------------
  const int c0 = s * cw * VEC;
  if (warp == GAUSS_NC / 32) asm volatile("bar.sync 2, %0;" ::"n"(GAUSS_NT) : "memory");   // the prologue's loads are on their way
  if (tid == GAUSS_NC) {
    for (int i = 0; i < ring; ++i) {
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&s_empty[i], GAUSS_NC / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int i = 0; i < ring && i < nst; ++i) {
      tc::mbar_arrive_expect_tx(&s_full[i], stage_bytes);
      if (p.t_outer) tma_load_4d(s_ring + (size_t)i * svec, &mX, &s_full[i], c0, 0, b, i * TG);
      else           tma_load_4d(s_ring + (size_t)i * svec, &mX, &s_full[i], c0, 0, i * TG, b);
    }
  }

  // ---- router / Gaussian prologue (HBM is already streaming) ----------------------------------
  gauss_prologue_cluster<GAUSS_NT>(p, b, cr, csz, sm);
  if (s == 0) gauss_write_aux_cluster<GAUSS_NT>(p, b, sm);

  if (warp == GAUSS_NC / 32) {
    // ---- producer: refill a stage as soon as the four consumer warps have released it ----
    if (lane == 0) {
      int st = 0, ph = 0;
      for (int it = ring; it < nst; ++it) {
        tc::mbar_wait(&s_empty[st], (uint32_t)ph);
        tc::mbar_arrive_expect_tx(&s_full[st], stage_bytes);
        if (p.t_outer) tma_load_4d(s_ring + (size_t)st * svec, &mX, &s_full[st], c0, 0, b, it * TG);
        else           tma_load_4d(s_ring + (size_t)st * svec, &mX, &s_full[st], c0, 0, it * TG, b);
        if (++st == ring) { st = 0; ph ^= 1; }
      }
    }
    return;
  }

  // ---- consumers: thread owns local vectors j = tid + k*GAUSS_NC of every time step ----------------------
  float acc[SLOTS][VEC];
  int je[SLOTS];
#pragma unroll
  for (int k = 0; k < SLOTS; ++k) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[k][v] = 0.f;
    const int j = tid + k * GAUSS_NC;
    je[k] = j < nloc ? j / cw : 0;
  }
  {
    int st = 0, ph = 0;
    for (int it = 0, t = 0; it < nst; ++it) {
      tc::mbar_wait(&s_full[st], (uint32_t)ph);
      const VecT* xs = s_ring + (size_t)st * svec;
      for (int g = 0; g < TG && t < T; ++g, ++t, xs += nloc) {
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
          const int j = tid + k * GAUSS_NC;
          if (j < nloc) fma_vec<VecT, VEC>(acc[k], xs[j], s_coef[t * E + je[k]]);
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&s_empty[st])) : "memory");
      if (++st == ring) { st = 0; ph ^= 1; }
    }
  }

  if (p.per_expert) {
    // out[b, e*C + c] per expert; the slice of rank 0 adds out[b, E*C + e] = sum_t coef[t][e] (zero beyond E, 64 columns)
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      const int j = tid + k * GAUSS_NC;
      const int cvi = s * cw + (j - je[k] * cw);
      if (j < nloc && cvi < cvec) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) act_store1(p.out, b, je[k] * C + cvi * VEC + v, acc[k][v]);
      }
    }
    if (s == 0 && tid < 64) {
      float sum = 0.f;
      if (tid < E)
        for (int t = 0; t < T; ++t) sum += s_coef[t * E + tid];
      act_store1(p.out, b, E * C + tid, sum);
    }
  } else {
    // out[b, c] = sum over the experts in index order: the per-expert sums meet in shared memory (the ring is free: every
    // stage has been consumed, the producer has no load outstanding)
    float* s_part = reinterpret_cast<float*>(gauss_smem);   // [E][cw*VEC]
    asm volatile("bar.sync 1, %0;" ::"n"(GAUSS_NC) : "memory");
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      const int j = tid + k * GAUSS_NC;
      if (j < nloc) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) s_part[(size_t)j * VEC + v] = acc[k][v];
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(GAUSS_NC) : "memory");
    const int wcols = cw * VEC;
    for (int o = tid; o < wcols; o += GAUSS_NC) {
      const int col = s * wcols + o;
      if (col >= C) break;
      float sum = 0.f;
      for (int e = 0; e < E; ++e) sum += s_part[e * wcols + o];
      act_store1(p.out, b, col, sum);
    }
  }
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
  const GaussSm sm = gauss_sm_carve(smem_raw, T, E, K);
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
  const size_t smem = gauss_sm_bytes(p.T, p.E, p.K);
  QAT_REQUIRE(smem <= 200 * 1024, "moe_coef: T*E too large for one CTA");
  if (smem > 48 * 1024) QAT_CUDA(cudaFuncSetAttribute(moe_coef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  moe_coef_kernel<<<dim3(p.B, n), MOE_COEF_NT, smem, c.stream>>>(pk);
  return check_launch("moe_coef_kernel");
}

// S = CTAs per sample (0 = auto), cut along the channel axis.
template <typename VecT, int VEC, int SLOTS>
inline int gauss_launch(Ctx& c, const GaussP& p_in, CUtensorMapDataType dt) {
  GaussP p = p_in;
  const int elem = (int)sizeof(VecT) / VEC;
  const int nloc = p.E * p.cw;
  const size_t rowb = (size_t)nloc * sizeof(VecT);
  // ring: all CTAs co-resident (one wave), so the per-CTA budget is the SM's shared memory over the CTAs it hosts
  const long long ctas = (long long)p.B * p.split;
  int per_sm = (int)((ctas + 147) / 148) + (p.csz > 1 ? 1 : 0);   // one slot of slack: clusters are placed GPC by GPC
  per_sm = per_sm > 8 ? 8 : per_sm;
  const size_t fixed = gauss_cl_bytes(p.T, p.E, p.K, p.csz) + 2 * 8 * 8 + 1024 + 128;
  const size_t budget = (size_t)(225 * 1024) / per_sm;
  const size_t ring_bytes = budget > fixed + 2 * rowb ? budget - fixed : 2 * rowb;
  int tg = (int)(7168 / rowb);
  tg = tg < 1 ? 1 : (tg > 8 ? 8 : tg);
  tg = tg > p.T ? p.T : tg;
  while (tg > 1 && ring_bytes / (tg * rowb) < 3) --tg;
  const int nst = (p.T + tg - 1) / tg;
  int ring = (int)(ring_bytes / (tg * rowb));
  ring = ring > 8 ? 8 : ring;
  ring = ring > nst ? nst : ring;
  ring = ring < 2 ? 2 : ring;   // >= 2 stages: the per-expert partial sums of the reference-layout form reuse the ring
  p.tg = tg;
  p.ring = ring;
  const size_t part = p.per_expert ? 0 : (size_t)nloc * VEC * 4;
  size_t ring_sz = (size_t)ring * ((tg * rowb + 127) / 128 * 128);
  QAT_REQUIRE(ring_sz >= part, "gauss_reduce: internal: ring smaller than the partial-sum tile");
  const size_t smem = ring_sz + 2 * (size_t)ring * 8 + gauss_cl_bytes(p.T, p.E, p.K, p.csz) + 16;
  QAT_REQUIRE(smem <= 220 * 1024, "gauss_reduce: a time step of %zu bytes per CTA does not fit the shared-memory ring", rowb);
  // X as a 4-D tensor: (c, e) contiguous, then the b and t axes in stride order
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) QAT_FAIL(QAT_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  p.t_outer = p.st >= p.sb ? 1 : 0;
  CUtensorMap mX;
  {
    const cuuint64_t d2 = p.t_outer ? (cuuint64_t)p.B : (cuuint64_t)p.T, d3 = p.t_outer ? (cuuint64_t)p.T : (cuuint64_t)p.B;
    const long long s2 = p.t_outer ? p.sb : p.st, s3 = p.t_outer ? p.st : p.sb;
    cuuint64_t dims[4] = {(cuuint64_t)p.C, (cuuint64_t)p.E, d2, d3};
    // a size-1 axis may carry any stride; keep the encoder's "multiple of 16, non-zero" rule satisfied
    cuuint64_t strides[3] = {(cuuint64_t)p.C * elem, (cuuint64_t)(s2 > 0 ? s2 : (long long)p.E * p.C) * elem,
                             (cuuint64_t)(s3 > 0 ? s3 : (long long)p.E * p.C) * elem};
    cuuint32_t box[4] = {(cuuint32_t)(p.cw * VEC), (cuuint32_t)p.E, p.t_outer ? 1u : (cuuint32_t)tg, p.t_outer ? (cuuint32_t)tg : 1u};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(&mX, dt, 4, const_cast<void*>(p.x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      QAT_FAIL(QAT_ECUDA, "gauss_reduce: cuTensorMapEncodeTiled failed with code %d (C=%d E=%d B=%d T=%d st=%lld sb=%lld)", (int)r, p.C, p.E,
               p.B, p.T, p.st, p.sb);
  }
  auto kern = gauss_reduce_kernel<VecT, VEC, SLOTS>;
  if (smem > 48 * 1024) QAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas, 1, 1);
  cfg.blockDim = dim3(GAUSS_NT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.csz;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  QAT_CUDA(cudaLaunchKernelEx(&cfg, kern, mX, p));
  return check_launch("gauss_reduce_kernel");
}

// dtype: QAT_DT_F32 / QAT_DT_BF16.  split: CTAs per sample (column slices), 0 = auto.
inline int gauss_reduce(Ctx& c, const GaussP& p_in, int dtype, int split) {
  GaussP p = p_in;
  QAT_REQUIRE(p.E >= 1 && p.E <= GAUSS_MAX_E && p.K >= 1 && p.K <= p.E, "gauss_reduce: need 1<=K<=E<=%d (E=%d K=%d)", GAUSS_MAX_E, p.E, p.K);
  QAT_REQUIRE(p.T >= 1 && p.T <= 4096 && p.B >= 1, "gauss_reduce: bad B/T");
  const int VEC = dtype == QAT_DT_BF16 ? 8 : 4;
  QAT_REQUIRE(dtype == QAT_DT_F32 || dtype == QAT_DT_BF16, "gauss_reduce: unknown dtype %d", dtype);
  QAT_REQUIRE(p.C % VEC == 0 && p.st % VEC == 0 && p.sb % VEC == 0, "gauss_reduce: C and strides must be multiples of %d elements", VEC);
  QAT_REQUIRE(reinterpret_cast<uintptr_t>(p.x) % 16 == 0, "gauss_reduce: experts tensor must be 16-byte aligned");
  QAT_REQUIRE((size_t)p.E * p.C <= 16384, "gauss_reduce: E*C=%d exceeds 16384", p.E * p.C);
  QAT_REQUIRE(split >= 0 && split <= 64, "gauss_reduce: t_split=%d outside [0,64]", split);
  if (c.dry) return QAT_OK;
  const int cvec = p.C / VEC;
  // a box row holds at most 256 elements; a CTA at most 8 vectors per consumer thread and time step
  int smin = (p.C + 255) / 256;
  while ((long long)p.E * ((cvec + smin - 1) / smin) > 8 * GAUSS_NC) ++smin;
  int smax = cvec < 8 ? cvec : 8;
  smax = smax < smin ? smin : smax;
  if (split <= 0) {
    // the slice count whose B*S CTAs fill the 148 SMs most evenly (all co-resident: at most 8 per SM), smaller on ties
    double best = 1e30;
    for (int s = smin; s <= smax; ++s) {
      const double per = (double)p.B * s / 148.0;
      if (per > 8.0 && s > smin) break;
      const double cost = (double)(long long)(per + 0.999999) / per;
      if (cost < best - 1e-9) {
        best = cost;
        split = s;
      }
    }
  }
  split = split < smin ? smin : (split > cvec ? cvec : split);
  p.split = split;
  p.cw = (cvec + split - 1) / split;
  p.split = (cvec + p.cw - 1) / p.cw;   // drop slices that would be empty
  p.csz = p.split <= 8 ? p.split : 1;   // portable cluster size; wider splits (C > 2048) run one prologue per CTA
  const int nloc = p.E * p.cw;
  const CUtensorMapDataType dt = dtype == QAT_DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#define QAT_GAUSS_GO(VT, V)                                                      \
  do {                                                                           \
    if (nloc <= 1 * GAUSS_NC) return gauss_launch<VT, V, 1>(c, p, dt);           \
    if (nloc <= 2 * GAUSS_NC) return gauss_launch<VT, V, 2>(c, p, dt);           \
    if (nloc <= 4 * GAUSS_NC) return gauss_launch<VT, V, 4>(c, p, dt);           \
    return gauss_launch<VT, V, 8>(c, p, dt);                                     \
  } while (0)
  if (dtype == QAT_DT_F32) QAT_GAUSS_GO(float4, 4);
  QAT_GAUSS_GO(uint4, 8);
#undef QAT_GAUSS_GO
}

}  // namespace qat
