// ToMe bipartite soft matching (T1), size-weighted merge (T2), source tracking as group ids (T3) and the
// stand-in metric of the 577->14 schedule (T0).  Reference: src/tome/merge.py:18-97, 210-241.
//
// Bit-exactness: the reference's indices are a function of fp32 roundings, so the
// arithmetic ORDER of torch's CPU kernels is reproduced, not just their math.  Orders were pinned
// empirically against torch 2.11 CPU in the build container (0 mismatching floats):
//   norm   : 8 interleaved partial sums acc[l] += x[d+l]*x[d+l] (separate mul and add), then
//            ((acc0+acc1)+acc2)+...; sqrt; every element divided by the norm        (merge.py:50)
//   scores : one sequential FMA chain over the channel index, starting from 0        (merge.py:52)
//   max    : first index among equal maxima; argsort descending, equal keys keep index order (:59-60)
//   merge  : dst += src in ascending order of the sorted edge list; x*size first, /size last (:70-80,221-224)
//   metric : ((x_0 + x_1) + ...) + x_15 over the 16 channel groups, then /16          
// All of it is plain IEEE fp32 (__fmul_rn/__fadd_rn/__fmaf_rn/__fdiv_rn/__fsqrt_rn: no contraction).
#pragma once
#include "qat_common.cuh"

namespace qat {

// `class_token` throughout this file is the protected-token flag word of the C ABI: bit 0 = token 0 is a class token
// (merge.py:54-55, 66-68), bit 1 = token 1 is a distillation token (merge.py:56-57, 77-78)
__host__ __device__ inline int tome_clamp_r(int t, int r, int class_token) {
  int prot = (class_token & 1) + ((class_token >> 1) & 1);
  int m = (t - prot) / 2;
  r = r < m ? r : m;
  return r > 0 ? r : 0;
}

constexpr int TOME_MAX_C = 64;   // metric channels held in registers per thread

// Output row layout of merge(): [unm | dst], or with a distillation token [unm_0, dst_0, unm_1.., dst_1..] (merge.py:77-80).
__device__ __forceinline__ int tome_pos_unm(int k, int distill) { return distill && k > 0 ? k + 1 : k; }
__device__ __forceinline__ int tome_pos_dst(int j, int n_unm, int distill) { return distill ? (j == 0 ? 1 : n_unm + j) : n_unm + j; }
// row -> (is it an unmerged A token?, index into unm_idx / B)
__device__ __forceinline__ bool tome_row_is_unm(int row, int n_unm, int distill, int& idx) {
  if (!distill) {
    idx = row < n_unm ? row : row - n_unm;
    return row < n_unm;
  }
  if (row == 0) { idx = 0; return true; }
  if (row == 1) { idx = 0; return false; }
  if (row <= n_unm) { idx = row - 1; return true; }
  idx = row - n_unm;
  return false;
}

// ---- T1 -------------------------------------------------------------------------------------------
// One CTA per frame.  Thread (i, half) owns row i of set A (even tokens) and scans half of set B.
template <int C>
__global__ void __launch_bounds__(640) tome_match_kernel(const float* __restrict__ metric, int t, int r, int class_token,
                                                         int* __restrict__ unm_idx, int* __restrict__ src_idx,
                                                         int* __restrict__ dst_idx) {
  extern __shared__ __align__(16) float smem[];
  const int ta = (t + 1) / 2, tb = t / 2;
  float* Bs = smem;                                  // [tb][C] normalised odd tokens
  float* nmax = Bs + (size_t)tb * C;                 // [ta]
  int* nidx = reinterpret_cast<int*>(nmax + ta);     // [ta]
  int* rank = nidx + ta;                             // [ta]
  float* hmax = reinterpret_cast<float*>(rank + ta); // [ta] second-half partials
  int* hidx = reinterpret_cast<int*>(hmax + ta);     // [ta]
  const int n = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
  const float* mf = metric + (size_t)n * t * C;

  // out <- row / ||row||, in place (keeps one C-float array live)
  auto normalise = [&](const float* row, float (&out)[C]) {
#pragma unroll
    for (int d = 0; d < C; d += 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(row + d));
      out[d] = v.x; out[d + 1] = v.y; out[d + 2] = v.z; out[d + 3] = v.w;
    }
    float acc[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) acc[l] = 0.f;
#pragma unroll
    for (int d = 0; d < C; d += 8)
#pragma unroll
      for (int l = 0; l < 8; ++l) acc[l] = __fadd_rn(acc[l], __fmul_rn(out[d + l], out[d + l]));
    float s = acc[0];
#pragma unroll
    for (int l = 1; l < 8; ++l) s = __fadd_rn(s, acc[l]);
    const float nrm = __fsqrt_rn(s);
#pragma unroll
    for (int d = 0; d < C; ++d) out[d] = __fdiv_rn(out[d], nrm);
  };

  // normalised B rows -> shared
  for (int j = tid; j < tb; j += NT) {
    float o[C];
    normalise(mf + (size_t)(2 * j + 1) * C, o);
#pragma unroll
    for (int d = 0; d < C; d += 4) *reinterpret_cast<float4*>(Bs + (size_t)j * C + d) = make_float4(o[d], o[d + 1], o[d + 2], o[d + 3]);
  }
  __syncthreads();

  const int halfT = NT / 2;
  const int half = tid / halfT;
  const bool cls = (class_token & 1) != 0, dis = (class_token & 2) != 0;
  const int jlo = half == 0 ? 0 : (tb + 1) / 2, jhi = half == 0 ? (tb + 1) / 2 : tb;
  for (int i = tid % halfT; i < ta; i += halfT) {
    float a[C];
    normalise(mf + (size_t)(2 * i) * C, a);
    float best = -INFINITY;
    int bidx = jlo;
    int j = (dis && jlo == 0) ? 1 : jlo;                       // merge.py:57: column 0 (the distillation token) is -inf
    for (; j + 4 <= jhi; j += 4) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      const float4* b0 = reinterpret_cast<const float4*>(Bs + (size_t)j * C);
#pragma unroll
      for (int d4 = 0; d4 < C / 4; ++d4) {
        const float4 v0 = b0[d4], v1 = b0[C / 4 + d4], v2 = b0[2 * (C / 4) + d4], v3 = b0[3 * (C / 4) + d4];
        const int d = d4 * 4;
        s0 = __fmaf_rn(a[d], v0.x, s0); s0 = __fmaf_rn(a[d + 1], v0.y, s0); s0 = __fmaf_rn(a[d + 2], v0.z, s0); s0 = __fmaf_rn(a[d + 3], v0.w, s0);
        s1 = __fmaf_rn(a[d], v1.x, s1); s1 = __fmaf_rn(a[d + 1], v1.y, s1); s1 = __fmaf_rn(a[d + 2], v1.z, s1); s1 = __fmaf_rn(a[d + 3], v1.w, s1);
        s2 = __fmaf_rn(a[d], v2.x, s2); s2 = __fmaf_rn(a[d + 1], v2.y, s2); s2 = __fmaf_rn(a[d + 2], v2.z, s2); s2 = __fmaf_rn(a[d + 3], v2.w, s2);
        s3 = __fmaf_rn(a[d], v3.x, s3); s3 = __fmaf_rn(a[d + 1], v3.y, s3); s3 = __fmaf_rn(a[d + 2], v3.z, s3); s3 = __fmaf_rn(a[d + 3], v3.w, s3);
      }
      if (s0 > best) { best = s0; bidx = j; }
      if (s1 > best) { best = s1; bidx = j + 1; }
      if (s2 > best) { best = s2; bidx = j + 2; }
      if (s3 > best) { best = s3; bidx = j + 3; }
    }
    for (; j < jhi; ++j) {
      float s0 = 0.f;
      const float* b0 = Bs + (size_t)j * C;
#pragma unroll
      for (int d = 0; d < C; ++d) s0 = __fmaf_rn(a[d], b0[d], s0);
      if (s0 > best) { best = s0; bidx = j; }
    }
    if (half == 0) { nmax[i] = best; nidx[i] = bidx; }
    else { hmax[i] = best; hidx[i] = bidx; }
  }
  __syncthreads();
  for (int i = tid; i < ta; i += NT) {
    float m0 = nmax[i];
    int i0 = nidx[i];
    // the upper half of B only wins with a strictly larger score (first index on ties, merge.py:59)
    if (hmax[i] > m0) { m0 = hmax[i]; i0 = hidx[i]; }
    if (cls && i == 0) { m0 = -INFINITY; i0 = 0; }             // merge.py:55: whole row -inf -> argmax 0
    nmax[i] = m0;
    nidx[i] = i0;
  }
  __syncthreads();
  // descending, index-stable rank by counting
  for (int i = tid; i < ta; i += NT) {
    const float v = nmax[i];
    int rk = 0;
    for (int k = 0; k < ta; ++k) {
      float u = nmax[k];
      rk += (u > v) || (u == v && k < i);
    }
    rank[i] = rk;
  }
  __syncthreads();
  const int n_unm = ta - r;
  int* unm = unm_idx + (size_t)n * n_unm;
  int* src = src_idx + (size_t)n * r;
  int* dst = dst_idx + (size_t)n * r;
  for (int i = tid; i < ta; i += NT) {
    const int rk = rank[i];
    if (rk < r) {
      src[rk] = i;
      dst[rk] = nidx[i];
    } else if (!(class_token & 1)) {
      unm[rk - r] = i;                                         // stays in sorted-score order
    } else {
      int pos = 0;                                             // merge.py:66-68: ascending token order
      for (int k = 0; k < i; ++k) pos += rank[k] >= r;
      unm[pos] = i;
    }
  }
}

// ---- T1, register-tiled -----------------------------------------------------------------------------
// One CTA per frame.  All t rows are normalised ONCE into shared memory (rows of 68 floats: consecutive rows land 16 bytes
// apart modulo 128, so a warp's 128-bit reads of 32 consecutive rows are conflict-free), then the ta x tb score matrix is
// computed in R x 8 register tiles (R = 2, 3 or 4, whichever pads the A rows least): lane l of an A-warp owns rows l, l+32, ..
// of its 32R-row band and walks one segment of the B rows eight at a time; the B reads of a warp are broadcasts.  4R + 8
// shared-memory wavefronts per 32R FMA instructions keeps the FMA pipe the bound; the class-token row, whose scores are
// discarded (merge.py:55), is not computed (the row-per-thread kernel above re-read B once per A row: 1 wavefront per 4
// FMAs, and spent as long again normalising its A rows twice).  Every score is still ONE sequential FMA chain over the
// channel index and the winners are compared in ascending j, so the indices are bit-identical to the reference's.
//
// Round 2 (ncu stall samples of the first form at t = 577: 20 % of the samples in the normalisation -- thread-per-row 128-bit
// global loads touch 32 lines per instruction --, 67 % in the score tiles with the FMA pipe 41 % active over the launch -- 18
// warps sit 5 / 5 / 4 / 4 on the four schedulers --, 12 % in the rank / index lists with half the warps idle):
//   * the raw rows arrive by 16-byte cp.async copies (consecutive threads = consecutive 16 bytes of the frame) and are
//     normalised in place, thread per row, two passes over shared memory -- no 64-register row copy;
//   * each warp walks a contiguous range of 8-column blocks; the segment count is the one whose busiest scheduler (warp w
//     issues on scheduler w % 4) walks the fewest blocks, up to 24 warps;
//   * ranks are counted by all threads (NT / ta threads per row, integer atomics: order-free), the ascending list of the
//     unmerged tokens (class-token form, merge.py:66-68) by a ballot prefix sum instead of a loop per token.
// Measured: the 22 match launches of the 577 -> 14 chain 2.41 -> 2.30 ms per 592 frames (step 0: 247 -> 220-239 us); the samples
// overstated what the side phases cost on the critical path.
constexpr int TM_LD = 68;
constexpr int TM_NT_MAX = 768;

template <int C, int R>
__global__ void __launch_bounds__(TM_NT_MAX) tome_match_tiled_kernel(const float* __restrict__ metric, int t, int r, int class_token, int nseg,
                                                                     int ncb, int* __restrict__ unm_idx, int* __restrict__ src_idx,
                                                                     int* __restrict__ dst_idx) {
  static_assert(C == 64, "metric channels");
  extern __shared__ __align__(16) float smem[];
  const int ta = (t + 1) / 2, tb = t / 2, tb_pad = ncb * 8;
  float* An = smem;                                          // [ta][TM_LD] normalised even tokens
  float* Bn = An + (size_t)ta * TM_LD;                       // [tb_pad][TM_LD] normalised odd tokens (+ zero rows)
  float* segmax = Bn + (size_t)tb_pad * TM_LD;               // [nseg][ta]
  int* segidx = reinterpret_cast<int*>(segmax + (size_t)nseg * ta);
  float* nmax = reinterpret_cast<float*>(segidx + (size_t)nseg * ta);   // [ta]
  int* nidx = reinterpret_cast<int*>(nmax + ta);             // [ta]
  int* rank = nidx + ta;                                     // [ta]
  int* wcnt = rank + ta;                                     // [32] per-warp counts of the unmerged-token prefix sum
  const int n = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
  const float* mf = metric + (size_t)n * t * C;

  // ---- phase 0: the frame's raw rows -> shared memory, 16 bytes per copy, consecutive threads = consecutive addresses ----
  for (int q = tid; q < t * (C / 4); q += NT) {
    const int row = q / (C / 4), d4 = q - row * (C / 4);
    float* dst = ((row & 1) ? Bn : An) + (size_t)(row >> 1) * TM_LD + d4 * 4;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(mf + (size_t)q * 4) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int q = tid; q < (tb_pad - tb) * (C / 4); q += NT)    // zero rows behind B: read by the last tile, never compared
    reinterpret_cast<float4*>(Bn + (size_t)(tb + q / (C / 4)) * TM_LD)[q % (C / 4)] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- phase 1: row / ||row|| (merge.py:50) in the reference's summation order, in place; threads [0, ta) own the A rows,
  //      [ta, t) the B rows (a warp's rows are consecutive in one matrix: conflict-free 128-bit accesses) ----
  for (int i = tid; i < t; i += NT) {
    float4* row = reinterpret_cast<float4*>(i < ta ? An + (size_t)i * TM_LD : Bn + (size_t)(i - ta) * TM_LD);
    float acc[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) acc[l] = 0.f;
#pragma unroll 4
    for (int d = 0; d < C / 4; d += 2) {
      const float4 v0 = row[d], v1 = row[d + 1];
      acc[0] = __fadd_rn(acc[0], __fmul_rn(v0.x, v0.x));
      acc[1] = __fadd_rn(acc[1], __fmul_rn(v0.y, v0.y));
      acc[2] = __fadd_rn(acc[2], __fmul_rn(v0.z, v0.z));
      acc[3] = __fadd_rn(acc[3], __fmul_rn(v0.w, v0.w));
      acc[4] = __fadd_rn(acc[4], __fmul_rn(v1.x, v1.x));
      acc[5] = __fadd_rn(acc[5], __fmul_rn(v1.y, v1.y));
      acc[6] = __fadd_rn(acc[6], __fmul_rn(v1.z, v1.z));
      acc[7] = __fadd_rn(acc[7], __fmul_rn(v1.w, v1.w));
    }
    float sq = acc[0];
#pragma unroll
    for (int l = 1; l < 8; ++l) sq = __fadd_rn(sq, acc[l]);
    const float nrm = __fsqrt_rn(sq);
#pragma unroll 4
    for (int d = 0; d < C / 4; ++d) {
      const float4 v = row[d];
      row[d] = make_float4(__fdiv_rn(v.x, nrm), __fdiv_rn(v.y, nrm), __fdiv_rn(v.z, nrm), __fdiv_rn(v.w, nrm));
    }
  }
  for (int i = tid; i < ta; i += NT) rank[i] = 0;
  __syncthreads();

  // ---- phase 2: scores = a . b (merge.py:52) and the running (max, first argmax) of every A row within a B segment ----
  {
    const int warp = tid >> 5, lane = tid & 31;
    const int row0 = (class_token & 1) ? 1 : 0;
    const bool dis = (class_token & 2) != 0;
    const int aw_n = (ta - row0 + 32 * R - 1) / (32 * R);
    const int seg = warp / aw_n, aw = warp % aw_n;
    if (seg < nseg) {
      int ri[R];
      const float* ap[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        ri[i] = row0 + aw * 32 * R + lane + 32 * i;
        ap[i] = An + (size_t)min(ri[i], ta - 1) * TM_LD;
      }
      float best[R];
      int bidx[R];
      const int j_beg = (int)((long long)seg * ncb / nseg) * 8, j_end = (int)((long long)(seg + 1) * ncb / nseg) * 8;
#pragma unroll
      for (int i = 0; i < R; ++i) { best[i] = -INFINITY; bidx[i] = j_beg; }
      for (int j0 = j_beg; j0 < j_end; j0 += 8) {
        float sc[R][8];
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) sc[i][jj] = 0.f;
        const float* bp = Bn + (size_t)j0 * TM_LD;
        // NOT unrolled: a fully unrolled tile lets the scheduler hoist all 192 loads ahead of the 2 048 FMAs (spills)
#pragma unroll 1
        for (int d4 = 0; d4 < C / 4; ++d4) {
          float4 a[R];
#pragma unroll
          for (int i = 0; i < R; ++i) a[i] = *reinterpret_cast<const float4*>(ap[i] + d4 * 4);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const float4 b = *reinterpret_cast<const float4*>(bp + jj * TM_LD + d4 * 4);
#pragma unroll
            for (int i = 0; i < R; ++i) {
              sc[i][jj] = __fmaf_rn(a[i].x, b.x, sc[i][jj]);
              sc[i][jj] = __fmaf_rn(a[i].y, b.y, sc[i][jj]);
              sc[i][jj] = __fmaf_rn(a[i].z, b.z, sc[i][jj]);
              sc[i][jj] = __fmaf_rn(a[i].w, b.w, sc[i][jj]);
            }
          }
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          if (j0 + jj < tb && !(dis && j0 + jj == 0)) {        // merge.py:57: column 0 (the distillation token) is -inf
#pragma unroll
            for (int i = 0; i < R; ++i)
              if (sc[i][jj] > best[i]) { best[i] = sc[i][jj]; bidx[i] = j0 + jj; }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < R; ++i)
        if (ri[i] < ta) { segmax[(size_t)seg * ta + ri[i]] = best[i]; segidx[(size_t)seg * ta + ri[i]] = bidx[i]; }
    }
  }
  __syncthreads();
  for (int i = tid; i < ta; i += NT) {
    float m0 = -INFINITY;
    int i0 = 0;
    if (!((class_token & 1) && i == 0)) {                      // merge.py:55: the class token's row is -inf -> argmax 0
      m0 = segmax[i];
      i0 = segidx[i];
      // a later segment only wins with a strictly larger score (first index on ties, merge.py:59)
      for (int sg = 1; sg < nseg; ++sg) {
        const float m = segmax[(size_t)sg * ta + i];
        if (m > m0) { m0 = m; i0 = segidx[(size_t)sg * ta + i]; }
      }
    }
    nmax[i] = m0;
    nidx[i] = i0;
  }
  __syncthreads();
  // descending, index-stable rank by counting; `parts` threads share a row's comparisons (integer atomics: order-free)
  {
    const int parts = NT / ta > 0 ? NT / ta : 1;
    for (int w = tid; w < ta * parts; w += NT) {
      const int part = w / ta, i = w - part * ta;
      const int k0 = (int)((long long)part * ta / parts), k1 = (int)((long long)(part + 1) * ta / parts);
      const float v = nmax[i];
      int rk = 0;
      for (int k = k0; k < k1; ++k) {
        const float u = nmax[k];
        rk += (u > v) || (u == v && k < i);
      }
      if (parts > 1) atomicAdd(&rank[i], rk);
      else rank[i] = rk;
    }
  }
  __syncthreads();
  const int n_unm = ta - r;
  int* unm = unm_idx + (size_t)n * n_unm;
  int* src = src_idx + (size_t)n * r;
  int* dst = dst_idx + (size_t)n * r;
  if (!(class_token & 1)) {
    for (int i = tid; i < ta; i += NT) {
      const int rk = rank[i];
      if (rk < r) {
        src[rk] = i;
        dst[rk] = nidx[i];
      } else {
        unm[rk - r] = i;                                       // stays in sorted-score order
      }
    }
  } else {
    // merge.py:66-68: the unmerged tokens in ascending token order = an exclusive prefix count of (rank >= r)
    const int warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
    int carry = 0;
    for (int base = 0; base < ta; base += NT) {                // one round for ta <= NT
      const int i = base + tid;
      const int rk = i < ta ? rank[i] : 0;
      const bool keep = i < ta && rk >= r;
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) wcnt[warp] = __popc(bal);
      __syncthreads();
      int pos = carry;
      for (int w = 0; w < warp; ++w) pos += wcnt[w];
      pos += __popc(bal & ((1u << lane) - 1u));
      if (i < ta) {
        if (keep) unm[pos] = i;
        else { src[rk] = i; dst[rk] = nidx[i]; }
      }
      int tot = 0;
      for (int w = 0; w < nw; ++w) tot += wcnt[w];
      carry += tot;
      __syncthreads();
    }
  }
}

// ---- T2 + T3 ------------------------------------------------------------------------------------
// grid (N, RS): CTA (n, s) writes output rows s, s+RS, ... of frame n.  A dst row scans the (<= r) edge
// list held in shared memory for its sources, in ascending edge order (the CPU scatter order).
// mode: QAT_TOME_WAVG  x*size summed, divided by the summed size (merge_wavg, merge.py:210-225)
//       QAT_TOME_SUM / _MEAN / _AMAX  the bare `merge(x, mode)` closure (:70-80; mean = sum / (1 + #sources))
enum { QAT_TOME_WAVG = 0, QAT_TOME_SUM = 1, QAT_TOME_MEAN = 2, QAT_TOME_AMAX = 3 };

template <int V> struct TVec;
template <> struct TVec<4> {
  typedef float4 T;
  static __device__ __forceinline__ T ld(const float* p) { return ldg_stream(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ void st(float* p, const T& v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct TVec<1> {
  typedef float T;
  static __device__ __forceinline__ T ld(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void st(float* p, const T& v) { *p = v; }
};
template <typename F> __device__ __forceinline__ float4 vmap(const float4& a, F f) { return make_float4(f(a.x), f(a.y), f(a.z), f(a.w)); }
template <typename F> __device__ __forceinline__ float vmap(const float& a, F f) { return f(a); }
template <typename F> __device__ __forceinline__ float4 vzip(const float4& a, const float4& b, F f) { return make_float4(f(a.x, b.x), f(a.y, b.y), f(a.z, b.z), f(a.w, b.w)); }
template <typename F> __device__ __forceinline__ float vzip(const float& a, const float& b, F f) { return f(a, b); }

template <int V>
__global__ void __launch_bounds__(256) tome_merge_kernel(const float* __restrict__ x, const float* __restrict__ size, int t, int C, int r, int mode,
                                                         const int* __restrict__ unm_idx, const int* __restrict__ src_idx,
                                                         const int* __restrict__ dst_idx, float* __restrict__ x_out,
                                                         float* __restrict__ size_out, int* __restrict__ group, int T0, int distill) {
  extern __shared__ __align__(16) int smi[];
  const int ta = (t + 1) / 2, tb = t / 2, n_unm = ta - r, t_out = t - r;
  int* s_src = smi;               // [r]
  int* s_dst = s_src + r;         // [r]
  int* s_newpos = s_dst + r;      // [t]   (only when group != null)
  const int n = blockIdx.x, tid = threadIdx.x;
  for (int k = tid; k < r; k += blockDim.x) {
    s_src[k] = src_idx[(size_t)n * r + k];
    s_dst[k] = dst_idx[(size_t)n * r + k];
  }
  __syncthreads();
  const float* xf = x + (size_t)n * t * C;
  const bool wavg = mode == QAT_TOME_WAVG;
  const float* sf = (wavg && size) ? size + (size_t)n * t : nullptr;
  const int CV = C / V;
  const int lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  typedef typename TVec<V>::T VT;
  for (int row = blockIdx.y * nwarp + warp; row < t_out; row += gridDim.y * nwarp) {
    float* xo = x_out + ((size_t)n * t_out + row) * C;
    int ri;
    if (tome_row_is_unm(row, n_unm, distill, ri)) {
      const int tok = 2 * unm_idx[(size_t)n * n_unm + ri];
      const float s = sf ? sf[tok] : 1.f;
      for (int c = lane; c < CV; c += 32) {
        VT v = TVec<V>::ld(xf + (size_t)tok * C + c * V);
        if (wavg) v = vmap(v, [s](float a) { return __fdiv_rn(__fmul_rn(a, s), s); });
        TVec<V>::st(xo + c * V, v);
      }
      if (wavg && lane == 0) size_out[(size_t)n * t_out + row] = s;
    } else {
      const int j = ri;
      const int tok = 2 * j + 1;
      const float s0 = sf ? sf[tok] : 1.f;
      float stot = s0;
      for (int k = 0; k < r; ++k)
        if (s_dst[k] == j) stot = __fadd_rn(stot, sf ? sf[2 * s_src[k]] : 1.f);
      for (int c = lane; c < CV; c += 32) {
        VT v = TVec<V>::ld(xf + (size_t)tok * C + c * V);
        if (wavg) v = vmap(v, [s0](float a) { return __fmul_rn(a, s0); });
        for (int k = 0; k < r; ++k) {
          if (s_dst[k] == j) {
            const int st = 2 * s_src[k];
            VT u = TVec<V>::ld(xf + (size_t)st * C + c * V);
            if (mode == QAT_TOME_AMAX) {
              v = vzip(v, u, [](float a, float b) { return fmaxf(a, b); });
            } else {
              const float ss = sf ? sf[st] : 1.f;
              if (wavg) u = vmap(u, [ss](float a) { return __fmul_rn(a, ss); });
              v = vzip(v, u, [](float a, float b) { return __fadd_rn(a, b); });
            }
          }
        }
        if (wavg || mode == QAT_TOME_MEAN) v = vmap(v, [stot](float a) { return __fdiv_rn(a, stot); });
        TVec<V>::st(xo + c * V, v);
      }
      if (wavg && lane == 0) size_out[(size_t)n * t_out + row] = stot;
    }
  }
  if (group && blockIdx.y == 0) {
    for (int j = tid; j < tb; j += blockDim.x) s_newpos[2 * j + 1] = tome_pos_dst(j, n_unm, distill);
    for (int k = tid; k < n_unm; k += blockDim.x) s_newpos[2 * unm_idx[(size_t)n * n_unm + k]] = tome_pos_unm(k, distill);
    for (int k = tid; k < r; k += blockDim.x) s_newpos[2 * s_src[k]] = tome_pos_dst(s_dst[k], n_unm, distill);
    __syncthreads();
    int* g = group + (size_t)n * T0;
    for (int o = tid; o < T0; o += blockDim.x) g[o] = s_newpos[g[o]];
  }
}

// ---- T2 for the 577->14 chain: size-weighted merge of C = 64*G channel rows (G <= 16) with the NEXT step's stand-in
// metric fused into the output pass.  A half-warp owns one output row: lane l holds float4 chunks l + 16 g (g < G), i.e.
// all G channel groups of 4 metric channels, so the metric ((x_0 + x_1) + ...) / G falls out of its registers in the
// reference's summation order and the merged tokens are never re-read.  All loads of a row are issued back to back.
// Arithmetic order is the one of tome_merge_kernel (bit-exact with torch CPU, see the header of this file).
template <int G>
__global__ void __launch_bounds__(256) tome_merge_wavg_kernel(const float* __restrict__ x, const float* __restrict__ size, int t, int r,
                                                              const int* __restrict__ unm_idx, const int* __restrict__ src_idx,
                                                              const int* __restrict__ dst_idx, float* __restrict__ x_out,
                                                              float* __restrict__ size_out, float* __restrict__ metric_out,
                                                              int* __restrict__ group, int T0, int distill) {
  extern __shared__ __align__(16) int smi[];
  constexpr int C = 64 * G;
  const int ta = (t + 1) / 2, tb = t / 2, n_unm = ta - r, t_out = t - r;
  int* s_src = smi;               // [r]
  int* s_dst = s_src + r;         // [r]
  int* s_newpos = s_dst + r;      // [t]   (only when group != null)
  const int n = blockIdx.x, tid = threadIdx.x;
  for (int k = tid; k < r; k += blockDim.x) {
    s_src[k] = src_idx[(size_t)n * r + k];
    s_dst[k] = dst_idx[(size_t)n * r + k];
  }
  __syncthreads();
  const float* xf = x + (size_t)n * t * C;
  const float* sf = size ? size + (size_t)n * t : nullptr;
  const int l16 = tid & 15, hw = tid >> 4, nhw = blockDim.x >> 4;
  const unsigned hmask = 0xFFFFu << (tid & 16);               // the lanes of this half-warp
  auto load_row = [&](int tok, float4 (&v)[G]) {
    const float4* p = reinterpret_cast<const float4*>(xf + (size_t)tok * C) + l16;
#pragma unroll
    for (int g = 0; g < G; ++g) v[g] = ldg_stream(p + 16 * g);
  };
  for (int row = blockIdx.y * nhw + hw; row < t_out; row += gridDim.y * nhw) {
    float4 v[G];
    float stot;
    int ri;
    if (tome_row_is_unm(row, n_unm, distill, ri)) {
      const int tok = 2 * unm_idx[(size_t)n * n_unm + ri];
      load_row(tok, v);
      stot = sf ? sf[tok] : 1.f;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        v[g].x = __fmul_rn(v[g].x, stot); v[g].y = __fmul_rn(v[g].y, stot); v[g].z = __fmul_rn(v[g].z, stot); v[g].w = __fmul_rn(v[g].w, stot);
      }
    } else {
      const int j = ri;
      const int tok = 2 * j + 1;
      load_row(tok, v);
      const float s0 = sf ? sf[tok] : 1.f;
      stot = s0;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        v[g].x = __fmul_rn(v[g].x, s0); v[g].y = __fmul_rn(v[g].y, s0); v[g].z = __fmul_rn(v[g].z, s0); v[g].w = __fmul_rn(v[g].w, s0);
      }
      // sources of this dst row, in ascending edge order (the CPU scatter order): 16 edges per ballot
      for (int k0 = 0; k0 < r; k0 += 16) {
        const int k = k0 + l16;
        const bool hit = k < r && s_dst[k] == j;
        unsigned m = (__ballot_sync(hmask, hit) >> (tid & 16)) & 0xFFFFu;
        while (m) {
          const int kk = k0 + __ffs(m) - 1;
          m &= m - 1;
          const int st = 2 * s_src[kk];
          const float ss = sf ? sf[st] : 1.f;
          float4 u[G];
          load_row(st, u);
          stot = __fadd_rn(stot, ss);
#pragma unroll
          for (int g = 0; g < G; ++g) {
            v[g].x = __fadd_rn(v[g].x, __fmul_rn(u[g].x, ss)); v[g].y = __fadd_rn(v[g].y, __fmul_rn(u[g].y, ss));
            v[g].z = __fadd_rn(v[g].z, __fmul_rn(u[g].z, ss)); v[g].w = __fadd_rn(v[g].w, __fmul_rn(u[g].w, ss));
          }
        }
      }
    }
    float4* xo = reinterpret_cast<float4*>(x_out + ((size_t)n * t_out + row) * C) + l16;
    float4 ms = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      v[g].x = __fdiv_rn(v[g].x, stot); v[g].y = __fdiv_rn(v[g].y, stot); v[g].z = __fdiv_rn(v[g].z, stot); v[g].w = __fdiv_rn(v[g].w, stot);
      xo[16 * g] = v[g];
      ms.x = __fadd_rn(ms.x, v[g].x); ms.y = __fadd_rn(ms.y, v[g].y); ms.z = __fadd_rn(ms.z, v[g].z); ms.w = __fadd_rn(ms.w, v[g].w);
    }
    if (metric_out) {
      const float d = (float)G;
      ms.x = __fdiv_rn(ms.x, d); ms.y = __fdiv_rn(ms.y, d); ms.z = __fdiv_rn(ms.z, d); ms.w = __fdiv_rn(ms.w, d);
      reinterpret_cast<float4*>(metric_out + ((size_t)n * t_out + row) * 64)[l16] = ms;
    }
    if (l16 == 0) size_out[(size_t)n * t_out + row] = stot;
  }
  if (group && blockIdx.y == 0) {
    for (int j = tid; j < tb; j += blockDim.x) s_newpos[2 * j + 1] = tome_pos_dst(j, n_unm, distill);
    for (int k = tid; k < n_unm; k += blockDim.x) s_newpos[2 * unm_idx[(size_t)n * n_unm + k]] = tome_pos_unm(k, distill);
    for (int k = tid; k < r; k += blockDim.x) s_newpos[2 * s_src[k]] = tome_pos_dst(s_dst[k], n_unm, distill);
    __syncthreads();
    int* g = group + (size_t)n * T0;
    for (int o = tid; o < T0; o += blockDim.x) g[o] = s_newpos[g[o]];
  }
}

// ---- stand-in metric: mean over the C/64 channel groups --------------------
// With ldx = 3C and x pointing at the K third of a packed [tokens, 3C] qkv projection this is the REAL metric of the ToMe-patched
// ViT, k.mean(heads) (timm.py:107): torch sums the heads left to right and divides by their number, exactly as below.
__global__ void __launch_bounds__(256) tome_metric_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ metric, long long tokens, int C) {
  const int G = C / 64;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one float4 of one token's 64-vector
  if (i >= tokens * 16) return;
  const long long tok = i >> 4;
  const int c4 = (int)(i & 15);
  const float4* xr = reinterpret_cast<const float4*>(x + tok * ldx) + c4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int g = 0; g < G; ++g) {
    float4 v = __ldg(xr + g * 16);
    s.x = __fadd_rn(s.x, v.x); s.y = __fadd_rn(s.y, v.y); s.z = __fadd_rn(s.z, v.z); s.w = __fadd_rn(s.w, v.w);
  }
  const float d = (float)G;
  s.x = __fdiv_rn(s.x, d); s.y = __fdiv_rn(s.y, d); s.z = __fdiv_rn(s.z, d); s.w = __fdiv_rn(s.w, d);
  reinterpret_cast<float4*>(metric + tok * 64)[c4] = s;
}

__global__ void iota_rows_kernel(int* g, long long rows, int cols) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * cols) g[i] = (int)(i % cols);
}

inline size_t tome_match_smem(int t, int C) {
  const int ta = (t + 1) / 2, tb = t / 2;
  return (size_t)tb * C * 4 + (size_t)ta * 5 * 4 + 16;
}

inline int tome_match(Ctx& c, const float* metric, int N, int t, int C, int r, int class_token, int* unm, int* src, int* dst) {
  QAT_REQUIRE(C == 64, "tome_match: metric channels C=%d unsupported (the ViT-L head dim 64 is built)", C);
  QAT_REQUIRE(r > 0 && r <= tome_clamp_r(t, r, class_token), "tome_match: r=%d not clamped for t=%d", r, t);
  // register-tiled kernel whenever the frame's normalised rows fit one CTA's shared memory (t <= ~700)
  {
    const int ta = (t + 1) / 2, tb = t / 2;
    const int rows = ta - ((class_token & 1) ? 1 : 0);
    int R = 4;                                                // the tile height that pads the A rows least (ties: the taller)
    for (int cand = 3; cand >= 2; --cand)
      if ((rows + 32 * cand - 1) / (32 * cand) * cand < (rows + 32 * R - 1) / (32 * R) * R) R = cand;
    const int aw_n = rows > 0 ? (rows + 32 * R - 1) / (32 * R) : 1;
    const int ncb = (tb + 7) / 8;                             // 8-column blocks of B
    // segments: warps = aw_n * nseg <= 24; warp w issues on scheduler w % 4, so pick the count whose busiest scheduler walks
    // the fewest 8-column blocks (ties: more warps, i.e. more latency hiding)
    int nseg = 1;
    {
      int best_load = 1 << 30;
      for (int cand = 1; cand * aw_n * 32 <= TM_NT_MAX && cand <= ncb; ++cand) {
        int load[4] = {0, 0, 0, 0};
        for (int sg = 0; sg < cand; ++sg) {
          const int nb = (int)((long long)(sg + 1) * ncb / cand) - (int)((long long)sg * ncb / cand);
          for (int aw = 0; aw < aw_n; ++aw) load[(sg * aw_n + aw) & 3] += nb;
        }
        int mx = load[0];
        for (int q = 1; q < 4; ++q) mx = load[q] > mx ? load[q] : mx;
        if (mx <= best_load) { best_load = mx; nseg = cand; }
      }
    }
    if (nseg > ncb) nseg = ncb;
    if (nseg < 1) nseg = 1;
    const size_t smem_t = ((size_t)(ta + ncb * 8) * TM_LD + (size_t)2 * nseg * ta + (size_t)3 * ta + 32) * 4 + 16;
    if (rows > 0 && aw_n * nseg * 32 <= TM_NT_MAX && smem_t <= 220 * 1024 && c.opt.v[OPT_TOME_MATCH_ROWS] != 1) {
      if (c.dry || N == 0) return QAT_OK;
      auto kern = R == 4 ? tome_match_tiled_kernel<64, 4> : (R == 3 ? tome_match_tiled_kernel<64, 3> : tome_match_tiled_kernel<64, 2>);
      if (smem_t > 48 * 1024) QAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
      kern<<<N, 32 * aw_n * nseg, smem_t, c.stream>>>(metric, t, r, class_token, nseg, ncb, unm, src, dst);
      return check_launch("tome_match_tiled_kernel");
    }
  }
  size_t smem = tome_match_smem(t, C);
  QAT_REQUIRE(smem <= 200 * 1024, "tome_match: t=%d too long for one CTA", t);
  if (c.dry || N == 0) return QAT_OK;
  if (smem > 48 * 1024) QAT_CUDA(cudaFuncSetAttribute(tome_match_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ta = (t + 1) / 2;
  int nt = ta <= 64 ? 128 : (ta <= 128 ? 256 : (ta <= 192 ? 384 : 640));
  tome_match_kernel<64><<<N, nt, smem, c.stream>>>(metric, t, r, class_token, unm, src, dst);
  return check_launch("tome_match_kernel");
}

inline int tome_merge(Ctx& c, const float* x, const float* size, int N, int t, int C, int r, int mode, const int* unm, const int* src,
                      const int* dst, float* x_out, float* size_out, int* group, int T0, int distill = 0) {
  QAT_REQUIRE(mode >= QAT_TOME_WAVG && mode <= QAT_TOME_AMAX, "tome_merge: unknown mode %d", mode);
  QAT_REQUIRE(r > 0 && r <= t / 2, "tome_merge: bad r=%d for t=%d", r, t);
  QAT_REQUIRE(mode != QAT_TOME_WAVG || size_out, "tome_merge: size_out missing");
  if (c.dry || N == 0) return QAT_OK;
  int rs = 1;
  while ((long long)N * rs < 148 * 4 && rs < 64) rs *= 2;
  size_t smem = ((size_t)2 * r + t) * 4 + 16;
  dim3 grid(N, rs);
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x_out)) % 16 == 0);
  if (vec)
    tome_merge_kernel<4><<<grid, 256, smem, c.stream>>>(x, size, t, C, r, mode, unm, src, dst, x_out, size_out, group, T0, distill);
  else
    tome_merge_kernel<1><<<grid, 256, smem, c.stream>>>(x, size, t, C, r, mode, unm, src, dst, x_out, size_out, group, T0, distill);
  return check_launch("tome_merge_kernel");
}

// chain step: merge_wavg + the next step's metric (metric_out may be NULL).  C must be 64*G with G in {4, 8, 16}.
inline bool tome_merge_wavg_supported(int C) { return C == 256 || C == 512 || C == 1024; }
inline int tome_merge_wavg(Ctx& c, const float* x, const float* size, int N, int t, int C, int r, const int* unm, const int* src,
                           const int* dst, float* x_out, float* size_out, float* metric_out, int* group, int T0, int distill = 0) {
  QAT_REQUIRE(tome_merge_wavg_supported(C), "tome_merge_wavg: C=%d unsupported", C);
  QAT_REQUIRE(r > 0 && r <= t / 2 && size_out, "tome_merge_wavg: bad r=%d for t=%d or size_out missing", r, t);
  QAT_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x_out) | reinterpret_cast<uintptr_t>(metric_out)) % 16 == 0,
              "tome_merge_wavg: tensors must be 16-byte aligned");
  if (c.dry || N == 0) return QAT_OK;
  int rs = 1;
  while ((long long)N * rs < 148 * 4 && rs < 64) rs *= 2;
  const size_t smem = ((size_t)2 * r + t) * 4 + 16;
  dim3 grid(N, rs);
  if (C == 1024) tome_merge_wavg_kernel<16><<<grid, 256, smem, c.stream>>>(x, size, t, r, unm, src, dst, x_out, size_out, metric_out, group, T0, distill);
  else if (C == 512) tome_merge_wavg_kernel<8><<<grid, 256, smem, c.stream>>>(x, size, t, r, unm, src, dst, x_out, size_out, metric_out, group, T0, distill);
  else tome_merge_wavg_kernel<4><<<grid, 256, smem, c.stream>>>(x, size, t, r, unm, src, dst, x_out, size_out, metric_out, group, T0, distill);
  return check_launch("tome_merge_wavg_kernel");
}

inline int tome_metric(Ctx& c, const float* x, float* metric, long long tokens, int C, long long ldx = 0) {
  QAT_REQUIRE(C % 64 == 0 && ldx % 4 == 0, "tome_metric: C=%d must be a multiple of 64 (and the row pitch of 4)", C);
  if (c.dry || tokens == 0) return QAT_OK;
  long long th = tokens * 16;
  tome_metric_kernel<<<(unsigned)((th + 255) / 256), 256, 0, c.stream>>>(x, ldx ? ldx : C, metric, tokens, C);
  return check_launch("tome_metric_kernel");
}

}  // namespace qat
