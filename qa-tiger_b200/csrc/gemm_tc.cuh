// tcgen05 tensor-core GEMMs (placeholder until the TMA/tcgen05 pipeline lands in this file).
#pragma once
#include "gemm_simt.cuh"
namespace qat {
inline int gemm_tc_tf32(Ctx&, const float*, int, long long, const float*, int, int, const Epi&, Act) {
  QAT_FAIL(QAT_EINVAL, "tensor-core GEMM not built");
}
inline int gemm_tc_bf16x3(Ctx&, const bf16*, const bf16*, int, long long, const bf16*, const bf16*, int, int, const Epi&, Act) {
  QAT_FAIL(QAT_EINVAL, "tensor-core GEMM not built");
}
}  // namespace qat
