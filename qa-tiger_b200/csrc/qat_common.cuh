// Shared host/device helpers for libqat_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/qat.h"

namespace qat {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------
// error plumbing: thread-local message, no exceptions across the C ABI
// ------------------------------------------------------------------------------------------------
extern thread_local char g_err[512];
extern thread_local long long g_launches;

// A/B switches (profiling aids; every default is the measured-faster path).  Process-wide, initialised ONCE from the environment
// (QAT_<NAME>) when the library is loaded and changed afterwards only through qat_set_option(); each entry point snapshots
// them into its Ctx, so the workspace dry run and the kernels of one call always see the same values and no getenv() runs on
// the hot path.
enum {
  OPT_GEMM_CG = 0,            // 1 | 2: single-CTA / CTA-pair tiles for every tcgen05 kind (0 = per-kind default)
  OPT_GEMM_BN,                // 128 | 256: tile width (0 = 256 when N % 256 == 0)
  OPT_ATTN_FP32_IN,           // 1: fp32 instead of fp16 (hi, lo) plane-pair in-projections into the AVQ attention cores
  OPT_PATCH_CHUNK,            // frames per PatchSelecter pass (0 = 16384)
  OPT_PATCH_FP32_PLANES,      // 1: fp32 intermediates on the patch path
  OPT_PATCH_NO_FUSED_ATTN,    // 1: in-projection GEMM + separate self-attention kernel
  OPT_PATCH_NO_FUSED_CROSS,   // 1: K | V GEMM + separate cross-attention kernel
  OPT_TOME_MATCH_ROWS,        // 1: the row-per-thread ToMe match kernel (the fallback of frames too long for the tiled one)
  OPT_MOE_NO_FUSED_REDUCE,    // 1: expert first-layer GEMM writes the hidden tensor, separate gauss_reduce launches
  OPT_GEMM_ATTN_NO_TS,        // 1: the fused PatchSelecter kernels with both operands from shared memory (gemm_attn.cuh) instead of the
                              //    A-resident form (A operand in tensor memory, gemm_attn_ts.cuh)
  OPT_GEMM_ATTN_TS_FLAGS,     // per-stage cost-model bits of the A-resident kernels (GemmAttnP::flags); read by a -DQAT_PROFILE_BUILD library only
  OPT_MOE_W2_SIMT,            // 1: the post-router contractions that default to the split-K bf16x3 tensor-core path (TempMoE's second expert layer,
                              //    QstGrounding's K | V projection) on the fp32 CUDA-core GEMM
  OPT_COUNT
};
struct Options {
  int v[OPT_COUNT];
};
extern std::atomic<int> g_opt[OPT_COUNT];

#define QAT_FAIL(code, ...)                     \
  do {                                          \
    snprintf(qat::g_err, sizeof(qat::g_err), __VA_ARGS__); \
    return (code);                              \
  } while (0)

#define QAT_REQUIRE(cond, ...)                  \
  do {                                          \
    if (!(cond)) QAT_FAIL(QAT_EINVAL, __VA_ARGS__); \
  } while (0)

#define QAT_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      QAT_FAIL(QAT_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define QAT_TRY(expr)        \
  do {                       \
    int _r = (expr);         \
    if (_r != QAT_OK) return _r; \
  } while (0)

// ------------------------------------------------------------------------------------------------
// An activation as the kernels see it: fp32 and/or a bf16 (hi, lo) pair and/or one fp16 plane, all with
// the same leading dimension.  x ~= hi + lo carries 16 mantissa bits, which is what the bf16x3 tensor-core
// GEMMs eat; the fp16 plane (11 significant bits, round-to-nearest) is the storage format of the patch path; the fp16
// (ph, pl) pair (x ~= ph + pl, 22 significant bits) is what the mma.sync attention cores on the router path eat -- the
// in-projection GEMM writes it instead of fp32 (same bytes), so the attention kernel stages K / V with plain async copies.
// ------------------------------------------------------------------------------------------------
struct Act {
  float* f = nullptr;
  bf16* hi = nullptr;
  bf16* lo = nullptr;
  __half* h = nullptr;
  __half* ph = nullptr;
  __half* pl = nullptr;
  int ld = 0;
  __host__ __device__ Act cols(int c0) const {
    Act a = *this;
    if (a.f) a.f += c0;
    if (a.hi) a.hi += c0;
    if (a.lo) a.lo += c0;
    if (a.h) a.h += c0;
    if (a.ph) a.ph += c0;
    if (a.pl) a.pl += c0;
    return a;
  }
  __host__ __device__ Act rows(long long r0) const {
    Act a = *this;
    if (a.f) a.f += r0 * ld;
    if (a.hi) a.hi += r0 * ld;
    if (a.lo) a.lo += r0 * ld;
    if (a.h) a.h += r0 * ld;
    if (a.ph) a.ph += r0 * ld;
    if (a.pl) a.pl += r0 * ld;
    return a;
  }
};

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t u) {
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

// x ~= hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 significant bits
__device__ __forceinline__ void split_half2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// store 4 consecutive values of one row into whichever planes exist (col multiple of 4)
__device__ __forceinline__ void act_store4(const Act& a, long long row, int col, float4 v) {
  long long off = row * a.ld + col;
  if (a.f) *reinterpret_cast<float4*>(a.f + off) = v;
  if (a.hi) {
    bf16 h0, h1, h2, h3, l0, l1, l2, l3;
    split_bf16(v.x, h0, l0);
    split_bf16(v.y, h1, l1);
    split_bf16(v.z, h2, l2);
    split_bf16(v.w, h3, l3);
    __nv_bfloat162 a0 = __halves2bfloat162(h0, h1), a1 = __halves2bfloat162(h2, h3);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a0);
    u.y = *reinterpret_cast<uint32_t*>(&a1);
    *reinterpret_cast<uint2*>(a.hi + off) = u;
    if (a.lo) {
      __nv_bfloat162 b0 = __halves2bfloat162(l0, l1), b1 = __halves2bfloat162(l2, l3);
      u.x = *reinterpret_cast<uint32_t*>(&b0);
      u.y = *reinterpret_cast<uint32_t*>(&b1);
      *reinterpret_cast<uint2*>(a.lo + off) = u;
    }
  }
  if (a.h) *reinterpret_cast<uint2*>(a.h + off) = make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
}

// store 2 consecutive values (col even)
__device__ __forceinline__ void act_store2(const Act& a, long long row, int col, float v0, float v1) {
  long long off = row * a.ld + col;
  if (a.f) *reinterpret_cast<float2*>(a.f + off) = make_float2(v0, v1);
  if (a.hi) {
    bf16 h0, h1, l0, l1;
    split_bf16(v0, h0, l0);
    split_bf16(v1, h1, l1);
    *reinterpret_cast<__nv_bfloat162*>(a.hi + off) = __halves2bfloat162(h0, h1);
    if (a.lo) *reinterpret_cast<__nv_bfloat162*>(a.lo + off) = __halves2bfloat162(l0, l1);
  }
  if (a.h) *reinterpret_cast<uint32_t*>(a.h + off) = pack_half2(v0, v1);
}

__device__ __forceinline__ void act_store1(const Act& a, long long row, int col, float v) {
  long long off = row * a.ld + col;
  if (a.f) a.f[off] = v;
  if (a.hi) {
    bf16 h, l;
    split_bf16(v, h, l);
    a.hi[off] = h;
    if (a.lo) a.lo[off] = l;
  }
  if (a.h) a.h[off] = __float2half_rn(v);
}

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const int n = valid ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled, the source is not read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(n) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}


// ------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA) primitives shared by the GEMM and the streaming kernels
// ------------------------------------------------------------------------------------------------
namespace tc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> cudaErrorLaunchFailure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
// 1-D bulk copy global -> shared (bytes % 16 == 0, both addresses 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- host side: the driver's tensor-map encoder, looked up once through the runtime (no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;   // idempotent one-time lookup (immutable afterwards)
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------
// Per-call context: stream, precision, bump allocator over the caller's workspace.
// `dry` = size the workspace without touching the GPU (qat_workspace_bytes runs the same code path).
// ------------------------------------------------------------------------------------------------
struct Ctx {
  cudaStream_t stream = nullptr;
  int prec = QAT_PREC_FP32;
  bool dry = false;
  Options opt;          // snapshot of the A/B switches taken when the call entered the library
  char* base = nullptr;
  size_t cap = 0;
  size_t off = 0;
  size_t peak = 0;

  void* alloc(size_t bytes) {
    size_t a = (off + 255) & ~size_t(255);
    off = a + bytes;
    if (off > peak) peak = off;
    if (dry) return reinterpret_cast<void*>(size_t(256));  // never dereferenced
    if (off > cap) return nullptr;
    return base + a;
  }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

// NVTX range per fused stage ("NVTX ranges per fused stage"): visible on an nsys / ncu --nvtx timeline, a no-op
// without a tool attached (header-only nvtx3: the injection library is looked up once).  Not opened by the sizing dry run.
struct StageRange {
  bool on;
  StageRange(const Ctx& c, const char* name) : on(!c.dry) {
    if (on) nvtxRangePushA(name);
  }
  ~StageRange() {
    if (on) nvtxRangePop();
  }
  StageRange(const StageRange&) = delete;
  StageRange& operator=(const StageRange&) = delete;
};

inline int ctx_init(Ctx& c, void* stream, int prec, void* ws, size_t ws_bytes, bool dry) {
  c.stream = reinterpret_cast<cudaStream_t>(stream);
  c.prec = prec;
  c.dry = dry;
  c.base = reinterpret_cast<char*>(ws);
  c.cap = ws_bytes;
  for (int i = 0; i < OPT_COUNT; ++i) c.opt.v[i] = g_opt[i].load(std::memory_order_relaxed);
  QAT_REQUIRE(prec == QAT_PREC_FP32 || prec == QAT_PREC_BF16, "unknown precision mode %d", prec);
  if (!dry) QAT_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
  return QAT_OK;
}

// allocate an activation [rows, ld] with the planes a consumer needs
inline int act_alloc(Ctx& c, Act& a, long long rows, int ld, bool want_f, bool want_planes) {
  a = Act();
  a.ld = ld;
  if (want_f) {
    a.f = reinterpret_cast<float*>(c.alloc(size_t(rows) * ld * 4));
    if (!a.f) QAT_FAIL(QAT_EWORKSPACE, "workspace too small (need > %zu bytes, have %zu)", c.off, c.cap);
  }
  if (want_planes) {
    a.hi = reinterpret_cast<bf16*>(c.alloc(size_t(rows) * ld * 2));
    a.lo = reinterpret_cast<bf16*>(c.alloc(size_t(rows) * ld * 2));
    if (!a.hi || !a.lo) QAT_FAIL(QAT_EWORKSPACE, "workspace too small (need > %zu bytes, have %zu)", c.off, c.cap);
  }
  return QAT_OK;
}

// a single fp16 plane [rows, ld] (patch-path storage format)
inline int act_alloc_h(Ctx& c, Act& a, long long rows, int ld) {
  a = Act();
  a.ld = ld;
  a.h = reinterpret_cast<__half*>(c.alloc(size_t(rows) * ld * 2));
  if (!a.h) QAT_FAIL(QAT_EWORKSPACE, "workspace too small (need > %zu bytes, have %zu)", c.off, c.cap);
  return QAT_OK;
}

// an fp16 (hi, lo) pair [rows, ld] (operands of the mma.sync attention cores)
inline int act_alloc_pp(Ctx& c, Act& a, long long rows, int ld) {
  a = Act();
  a.ld = ld;
  a.ph = reinterpret_cast<__half*>(c.alloc(size_t(rows) * ld * 2));
  a.pl = reinterpret_cast<__half*>(c.alloc(size_t(rows) * ld * 2));
  if (!a.ph || !a.pl) QAT_FAIL(QAT_EWORKSPACE, "workspace too small (need > %zu bytes, have %zu)", c.off, c.cap);
  return QAT_OK;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) QAT_FAIL(QAT_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  ++g_launches;
  return QAT_OK;
}

}  // namespace qat
