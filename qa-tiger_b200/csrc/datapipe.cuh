// batch assembly out of the device-resident per-video feature cache.
//
// The reference loads the three feature files of a question's video per QUESTION (dataset.py:132-159) and copies every batch to
// the device (trainutils.py:223-236); MUSIC-AVQA test has 9 129 questions over 6 399 videos, so 30 % of those bytes cross the host
// link twice.  qa_tiger_b200/datapipe.py keeps every video that has been copied once in HBM (slots of three cache tensors) and
// assembles a batch with ONE launch of this kernel: out_k[i, :] = cache_k[slot[i], :] for the tensors of a question (audio, video,
// patch).  Pure HBM copy: 128-bit loads, eight in flight per thread before the first store, no shared-memory staging needed
// (rows are contiguous and megabytes long; both sides coalesce perfectly).
#pragma once
#include "qat_common.cuh"

namespace qat {

struct GatherP {
  const uint4* src[QAT_GATHER_MAX];
  uint4* dst[QAT_GATHER_MAX];
  long long vec[QAT_GATHER_MAX];      // 16-byte vectors per row of tensor k
  long long svec[QAT_GATHER_MAX];     // 16-byte vectors between consecutive cache rows of tensor k (>= vec[k])
  long long vec_end[QAT_GATHER_MAX];  // prefix sums: tensor k owns [vec_end[k-1], vec_end[k]) of the concatenated row
  const int* slots;
  int n_tensors, n_slots;
};

constexpr int GATHER_NT = 256;
constexpr int GATHER_U = 8;   // vectors per thread per pass: 32 KB per CTA pass

__global__ void __launch_bounds__(GATHER_NT) gather_rows_kernel(GatherP p) {
  const int i = blockIdx.y;
  const int slot = __ldg(p.slots + i);
  const bool ok = slot >= 0 && slot < p.n_slots;
  const long long total = p.vec_end[p.n_tensors - 1];
  const long long v0 = (long long)blockIdx.x * (GATHER_NT * GATHER_U) + threadIdx.x;
  uint4 r[GATHER_U];
  int tk[GATHER_U];
  long long off[GATHER_U];
#pragma unroll
  for (int u = 0; u < GATHER_U; ++u) {
    const long long v = v0 + (long long)u * GATHER_NT;
    tk[u] = -1;
    if (v < total) {
      const int k = (int)(v >= p.vec_end[0]) + (int)(v >= p.vec_end[1]) + (int)(v >= p.vec_end[2]);   // vec_end[k] repeats past n_tensors
      tk[u] = k;
      off[u] = v - (k ? p.vec_end[k - 1] : 0);
      // a slot outside the cache (a host-side bookkeeping bug) poisons the row with NaN bit patterns instead of reading out of bounds
      r[u] = ok ? ldg_stream(p.src[k] + (long long)slot * p.svec[k] + off[u]) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    }
  }
#pragma unroll
  for (int u = 0; u < GATHER_U; ++u)
    if (tk[u] >= 0) p.dst[tk[u]][(long long)i * p.vec[tk[u]] + off[u]] = r[u];
}

inline int gather_rows(Ctx& c, const QatGatherArgs* a) {
  QAT_REQUIRE(a->n >= 0 && a->n_tensors >= 1 && a->n_tensors <= QAT_GATHER_MAX && a->n_slots >= 1, "gather_rows: bad n / n_tensors / n_slots");
  QAT_REQUIRE(a->slots, "gather_rows: slots is NULL");
  GatherP p;
  long long acc = 0;
  for (int k = 0; k < QAT_GATHER_MAX; ++k) {
    p.src[k] = nullptr; p.dst[k] = nullptr; p.vec[k] = 0; p.svec[k] = 0; p.vec_end[k] = acc;
    if (k >= a->n_tensors) { p.vec_end[k] = 0x7fffffffffffffffll; continue; }
    QAT_REQUIRE(a->src[k] && a->dst[k], "gather_rows: tensor %d is NULL", k);
    QAT_REQUIRE(a->row_bytes[k] > 0 && a->row_bytes[k] % 16 == 0, "gather_rows: row_bytes[%d]=%zu must be a positive multiple of 16", k, a->row_bytes[k]);
    QAT_REQUIRE(((reinterpret_cast<uintptr_t>(a->src[k]) | reinterpret_cast<uintptr_t>(a->dst[k])) & 15) == 0, "gather_rows: tensor %d must be 16-byte aligned", k);
    p.src[k] = reinterpret_cast<const uint4*>(a->src[k]);
    p.dst[k] = reinterpret_cast<uint4*>(a->dst[k]);
    p.vec[k] = (long long)(a->row_bytes[k] / 16);
    const size_t ss = a->src_stride[k] ? a->src_stride[k] : a->row_bytes[k];
    QAT_REQUIRE(ss % 16 == 0 && ss >= a->row_bytes[k], "gather_rows: src_stride[%d]=%zu must be a multiple of 16 and >= row_bytes", k, ss);
    p.svec[k] = (long long)(ss / 16);
    acc += p.vec[k];
    p.vec_end[k] = acc;
  }
  p.slots = a->slots; p.n_tensors = a->n_tensors; p.n_slots = a->n_slots;
  if (a->n == 0) return QAT_OK;
  QAT_REQUIRE(a->n <= 65535, "gather_rows: n=%d exceeds 65535 rows per call", a->n);
  const long long per = (long long)GATHER_NT * GATHER_U;
  dim3 grid((unsigned)((acc + per - 1) / per), (unsigned)a->n);
  gather_rows_kernel<<<grid, GATHER_NT, 0, c.stream>>>(p);
  return check_launch("gather_rows_kernel");
}

}  // namespace qat
