// libqat_b200: C-ABI entry points (include/qat.h) and the host-side orchestration of the sm_100a kernels.
// Every entry point: validate -> carve the caller's workspace -> enqueue kernels on the caller's stream.
// No allocation, no synchronisation, no global mutable state.
#include "attn_core.cuh"
#include "attn_f16.cuh"
#include "attn_long.cuh"
#include "datapipe.cuh"
#include "gauss.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gemm_attn_ts.cuh"
#include "gemm_splitk.cuh"
#include "moe_front.cuh"
#include "qat_common.cuh"
#include "rowops.cuh"
#include "tome.cuh"

namespace qat {

thread_local char g_err[512] = "";
thread_local long long g_launches = 0;

static const char* const kOptNames[OPT_COUNT] = {"gemm_cg", "gemm_bn", "attn_fp32_in", "patch_chunk", "patch_fp32_planes",
                                                 "patch_no_fused_attn", "patch_no_fused_cross", "tome_match_rows",
                                                 "moe_no_fused_reduce", "gemm_attn_no_ts", "gemm_attn_ts_flags", "moe_w2_simt"};
std::atomic<int> g_opt[OPT_COUNT];
// defaults from the environment, read once at load time (QAT_GEMM_CG=1 etc.)
static struct OptInit {
  OptInit() {
    for (int i = 0; i < OPT_COUNT; ++i) {
      char env[64] = "QAT_";
      size_t n = 4;
      for (const char* s = kOptNames[i]; *s && n + 1 < sizeof(env); ++s) env[n++] = (*s >= 'a' && *s <= 'z') ? char(*s - 32) : *s;
      env[n] = 0;
      const char* e = getenv(env);
      g_opt[i].store(e ? atoi(e) : 0, std::memory_order_relaxed);
    }
  }
} g_opt_init;

// rows [n0, n0+n) of one nn.Linear
struct WView {
  const float* w = nullptr;
  const bf16* hi = nullptr;
  const bf16* lo = nullptr;
  const __half* h16 = nullptr;
  const float* bias = nullptr;
  int n = 0, k = 0;
};

static WView wview(const QatLinearW& W, int n0 = 0, int n = -1) {
  WView v;
  if (n < 0) n = W.n - n0;
  v.n = n;
  v.k = W.k;
  v.w = W.w ? W.w + (size_t)n0 * W.k : nullptr;
  v.hi = W.w_hi ? reinterpret_cast<const bf16*>(W.w_hi) + (size_t)n0 * W.k : nullptr;
  v.lo = W.w_lo ? reinterpret_cast<const bf16*>(W.w_lo) + (size_t)n0 * W.k : nullptr;
  v.h16 = W.w_f16 ? reinterpret_cast<const __half*>(W.w_f16) + (size_t)n0 * W.k : nullptr;
  v.bias = W.bias ? W.bias + n0 : nullptr;
  return v;
}

static int check_w(const QatLinearW& W, int n, int k, const char* name, int prec) {
  QAT_REQUIRE(W.w != nullptr, "%s: weight pointer is NULL", name);
  QAT_REQUIRE(W.n == n && W.k == k, "%s: expected [%d,%d], got [%d,%d]", name, n, k, W.n, W.k);
  (void)prec;
  return QAT_OK;
}

// Which engine runs a given contraction (DESIGN.md "precision policy").
enum GemmKind { G_SIMT = 0, G_TF32 = 1, G_BF16X3 = 2, G_F16 = 3 };

static GemmKind pick_kind(const Ctx& c, long long M, int N, int K, bool patch_path, const WView& W) {
  if (c.prec != QAT_PREC_BF16) return G_SIMT;
  if (M < 1024 || N < 64 || (N % 8) != 0) return G_SIMT;     // tiny / fp32-only contractions stay on CUDA cores
  if (patch_path) return (K % 32 == 0) ? G_TF32 : G_SIMT;
  if (K % 64 == 0 && W.hi && W.lo) return G_BF16X3;
  return (K % 32 == 0) ? G_TF32 : G_SIMT;
}

static bool reads_planes(GemmKind k) { return k == G_BF16X3; }

// allocate an activation in the storage format the consuming engine reads (fp32 / bf16 hi,lo planes / fp16 plane)
static int act_alloc_for(Ctx& c, Act& a, long long rows, int ld, GemmKind consumer) {
  if (consumer == G_F16) return act_alloc_h(c, a, rows, ld);
  return act_alloc(c, a, rows, ld, !reads_planes(consumer), reads_planes(consumer));
}

// make sure `x` (which always carries its fp32 view) also has the storage planes engine `k` reads; a conversion pass runs
// only when the producer did not already emit them (inside qat_forward_fwd the producers do)
static int ensure_operand(Ctx& c, Act& x, long long rows, int cols, GemmKind k) {
  const bool need_hl = k == G_BF16X3 && !(x.hi && x.lo), need_h = k == G_F16 && !x.h;
  if (!need_hl && !need_h) return QAT_OK;
  QAT_REQUIRE(c.dry || x.f, "operand: fp32 view missing");
  QAT_REQUIRE(x.ld == cols, "operand: planes need a dense fp32 view (ld=%d cols=%d)", x.ld, cols);
  Act pl;
  QAT_TRY(act_alloc_for(c, pl, rows, cols, k));
  QAT_TRY(add_rows(c, x.f, x.ld, nullptr, 0, pl, rows, cols));
  if (need_hl) { x.hi = pl.hi; x.lo = pl.lo; }
  if (need_h) x.h = pl.h;
  return QAT_OK;
}

static int gemm(Ctx& c, GemmKind kind, const Act& A, long long M, const WView& W, Epi epi, Act C) {
  if (!epi.bias) epi.bias = W.bias;
  switch (kind) {
    case G_SIMT:
      QAT_REQUIRE(c.dry || A.f, "gemm: fp32 operand missing");
      QAT_REQUIRE(!C.ph, "gemm: the fp16 (hi, lo) output pair is written by the tcgen05 epilogue only");
      return gemm_simt(c, A.f, A.ld, M, W.w, W.n, W.k, epi, C);
    case G_TF32:
      QAT_REQUIRE(c.dry || A.f, "gemm: fp32 operand missing");
      return gemm_tc_tf32(c, A.f, A.ld, M, W.w, W.n, W.k, epi, C);
    case G_BF16X3:
      QAT_REQUIRE(c.dry || (A.hi && A.lo && W.hi && W.lo), "gemm: bf16 hi/lo operand planes missing");
      return gemm_tc_bf16x3(c, A.hi, A.lo, A.ld, M, W.hi, W.lo, W.n, W.k, epi, C);
    case G_F16:
      QAT_REQUIRE(c.dry || (A.h && W.h16), "gemm: fp16 operand plane missing");
      return gemm_tc_f16(c, A.h, A.ld, M, W.h16, W.n, W.k, epi, C);
  }
  return QAT_EINVAL;
}

static Act act_of(float* p, int ld) {
  Act a;
  a.f = p;
  a.ld = ld;
  return a;
}
static Act act_of(const float* p, int ld) { return act_of(const_cast<float*>(p), ld); }


static int ln(Ctx& c, const float* x0, long long ld0, const float* x1, long long ld1, const float* x2, long long ld2, float c01,
              const QatLayerNormW& w, Act out, long long rows, int D, const float* post = nullptr, long long ldp = 0, float eps = 1e-5f) {
  QAT_REQUIRE(w.gamma && w.beta, "layer_norm: affine parameters missing");
  LnP p;
  p.x0 = x0; p.ld0 = ld0; p.x1 = x1; p.ld1 = ld1; p.x2 = x2; p.ld2 = ld2;
  p.c01 = c01; p.gamma = w.gamma; p.beta = w.beta; p.out = out; p.rows = rows; p.D = D; p.eps = eps;
  p.post = post; p.ldp = ldp;
  return layer_norm(c, p);
}

// =================================================================================================
// P1  Projection
// =================================================================================================
static int linear_core(Ctx& c, const QatLinearW& w, Act X, long long rows, int relu, Act Y) {
  StageRange nvtx_(c, "qat.projection");
  QAT_REQUIRE(rows >= 0 && w.n > 0 && w.k > 0, "linear: bad shape");
  QAT_TRY(check_w(w, w.n, w.k, "linear.w", c.prec));
  WView W = wview(w);
  GemmKind kind = pick_kind(c, rows, W.n, W.k, false, W);
  QAT_TRY(ensure_operand(c, X, rows, W.k, kind));
  Epi e;
  e.relu = relu;
  return gemm(c, kind, X, rows, W, e, Y);
}

static int run_linear(const QatLinearArgs* a, Ctx& c) {
  return linear_core(c, a->w, act_of(a->x, a->w.k), a->rows, a->relu, act_of(a->y, a->w.n));
}

static int run_gemm_attn(const QatGemmAttnArgs* a, Ctx& c) {
  QAT_REQUIRE(a->x_f16 && a->w_f16 && a->bias && a->o_f16 && (!a->cross || a->q_f16), "gemm_attn: NULL tensor");
  QAT_REQUIRE(a->nframes >= 0 && a->P >= 1 && a->H >= 1, "gemm_attn: bad shape");
  const __half* X = reinterpret_cast<const __half*>(a->x_f16);
  const __half* W = reinterpret_cast<const __half*>(a->w_f16);
  __half* O = reinterpret_cast<__half*>(a->o_f16);
  if (a->cross)
    return gemm_attn_cross_f16(c, X, a->ldx, a->nframes, a->P, a->H, a->K, W, a->bias, a->scale, reinterpret_cast<const __half*>(a->q_f16), O, a->ldo);
  return gemm_attn_f16(c, X, a->ldx, a->nframes, a->P, a->H, a->K, W, a->bias, a->scale, O, a->ldo);
}

static int run_gemm(const QatGemmArgs* a, Ctx& c) {
  QAT_REQUIRE(a->kind >= G_SIMT && a->kind <= G_F16, "gemm: unknown kind %d", a->kind);
  QAT_REQUIRE(a->M >= 0 && a->w.n > 0 && a->w.k > 0 && a->lda >= a->w.k && a->ldc >= a->w.n, "gemm: bad shape");
  Act A, C;
  A.f = const_cast<float*>(a->a);
  A.hi = reinterpret_cast<bf16*>(const_cast<void*>(a->a_hi));
  A.lo = reinterpret_cast<bf16*>(const_cast<void*>(a->a_lo));
  A.h = reinterpret_cast<__half*>(const_cast<void*>(a->a_f16));
  A.ld = a->lda;
  C.f = a->c;
  C.hi = reinterpret_cast<bf16*>(a->c_hi);
  C.lo = reinterpret_cast<bf16*>(a->c_lo);
  C.h = reinterpret_cast<__half*>(a->c_f16);
  C.ld = a->ldc;
  Epi e;
  e.relu = a->relu;
  e.res = a->res;
  e.res_h = reinterpret_cast<const __half*>(a->res_f16);
  e.ldr = a->ldr;
  QAT_REQUIRE(a->kind != G_SIMT || !e.res_h, "gemm: the SIMT engine takes fp32 residuals only");
  if (a->coef) {
    QAT_REQUIRE(a->kind == G_BF16X3, "gemm: the reduce epilogue is built for the bf16x3 kind");
    e.coef = a->coef; e.coef_E = a->coef_E; e.coef_T = a->coef_T; e.coef_Ch = a->coef_Ch; e.coef_mod = a->coef_mod;
  }
  return gemm(c, (GemmKind)a->kind, A, a->M, wview(a->w), e, C);
}

static int run_attention(const QatAttnArgs* a, Ctx& c) {
  QAT_REQUIRE(a->nseq >= 0 && a->H >= 1 && a->Lq >= 1 && a->Lk >= 1, "attention: bad shape");
  if (a->kind == QAT_ATTN_F16_SMALL) {
    QAT_REQUIRE(a->q && a->k && a->o_f16, "attention: NULL tensor");
    AttnHP p;
    p.q = reinterpret_cast<const __half*>(a->q); p.kv = reinterpret_cast<const __half*>(a->k);
    p.ldq = a->q_si; p.ldkv = a->k_si; p.o = reinterpret_cast<__half*>(a->o_f16); p.ldo = a->ldo;
    p.nseq = a->nseq; p.H = a->H; p.Lq = a->Lq; p.Lk = a->Lk; p.scale = a->scale;
    return attention_small_f16(c, p);
  }
  QAT_REQUIRE(a->kind == QAT_ATTN_F32, "attention: unknown kind %d", a->kind);
  QAT_REQUIRE(a->q && a->k && a->v && a->o && a->k_mod >= 1, "attention: NULL tensor or k_mod < 1");
  QAT_REQUIRE(!a->causal || a->Lk > 32 || a->Lq >= 32, "attention: the causal core serves Lk > 32 or Lq >= 32");
  const bool long_core = a->Lk > 128 || a->kbias != nullptr;
  AttnP p;
  p.q = reinterpret_cast<const float*>(a->q); p.k = reinterpret_cast<const float*>(a->k); p.v = reinterpret_cast<const float*>(a->v);
  p.q_sn = a->q_sn; p.q_si = a->q_si; p.k_sn = a->k_sn; p.k_si = a->k_si; p.k_shift = a->k_shift; p.k_mod = a->k_mod;
  p.o = act_of(a->o, a->ldo); p.o_rn = a->o_rn; p.o_ri = a->o_ri; p.attw = a->attw;
  p.nseq = a->nseq; p.H = a->H; p.Lq = a->Lq; p.Lk = a->Lk; p.scale = a->scale; p.causal = a->causal;
  p.kbias = a->kbias;
  return long_core ? attention_long(c, p) : attention(c, p);
}

// =================================================================================================
// A1-A3  AVQCrossAttn
// =================================================================================================
// Xa / Xv / Xw carry the fp32 views (residuals) and whichever planes their producer emitted; out_a / out_v receive every
// plane they point at (fp32 always; inside qat_forward_fwd also the planes the next blocks read)
static int avq_core(const QatAvqArgs* a, int B, int T, int L, Ctx& c, Act Xa, Act Xv, Act Xw, Act out_a, Act out_v, float* attw_audio,
                    float* attw_video) {
  StageRange nvtx_(c, "qat.avq_cross_attn");
  const int D = a->D, H = a->H;
  QAT_REQUIRE(B >= 1 && T >= 1 && L >= 1 && H >= 1 && D == H * 64, "avq: need D == H*64 (D=%d H=%d)", D, H);
  // sequences up to 128 tokens run on the register-resident mma.sync cores (fp16 plane pairs in bf16 mode); longer ones on the
  // flash-style long-sequence core (attn_long.cuh), which reads fp32 q / k / v and has no weights output
  QAT_REQUIRE(T <= 4096 && L <= 4096, "avq: sequence lengths T=%d L=%d must be <= 4096", T, L);
  const bool longseq = T > 128 || L > 128;
  QAT_REQUIRE(!(L > 128 && (attw_audio || attw_video)), "avq: attention weights (visualize) need L=%d <= 128 words", L);
  QAT_TRY(check_w(a->in_src, 7 * D, D, "avq.in_src", c.prec));
  QAT_TRY(check_w(a->in_words, 2 * D, D, "avq.in_words", c.prec));
  QAT_TRY(check_w(a->out_cat, D, 3 * D, "avq.out_cat", c.prec));
  QAT_TRY(check_w(a->ffn1, D, D, "avq.ffn1", c.prec));
  QAT_TRY(check_w(a->ffn2, D, D, "avq.ffn2", c.prec));
  const long long BT = (long long)B * T, BL = (long long)B * L;
  const WView Wsrc = wview(a->in_src), Wwords = wview(a->in_words), Wout = wview(a->out_cat), W1 = wview(a->ffn1), W2 = wview(a->ffn2);
  const GemmKind k_src = pick_kind(c, BT, 7 * D, D, false, Wsrc), k_words = pick_kind(c, BL, 2 * D, D, false, Wwords);
  const GemmKind k_out = pick_kind(c, BT, D, 3 * D, false, Wout), k_f1 = pick_kind(c, BT, D, D, false, W1), k_f2 = pick_kind(c, BT, D, D, false, W2);

  QAT_TRY(ensure_operand(c, Xa, BT, D, k_src));
  QAT_TRY(ensure_operand(c, Xv, BT, D, k_src));
  QAT_TRY(ensure_operand(c, Xw, BL, D, k_words));
  // S = in-projections of both modalities, rows [0,BT) audio, [BT,2BT) video; cols as in_src
  // bf16 mode: the in-projections leave as fp16 (hi, lo) plane pairs (same bytes as fp32), the form the attention cores eat
  const bool pp = k_src == G_BF16X3 && k_words == G_BF16X3 && D % 8 == 0 && c.opt.v[OPT_ATTN_FP32_IN] != 1 && !longseq;
  auto attn = [&](const AttnP& q) { return q.Lk > 128 ? attention_long(c, q) : attention(c, q); };
  Act S, KVw, O, X1, Hh, Y;
  if (pp) {
    QAT_TRY(act_alloc_pp(c, S, 2 * BT, 7 * D));
    QAT_TRY(act_alloc_pp(c, KVw, BL, 2 * D));
  } else {
    QAT_TRY(act_alloc(c, S, 2 * BT, 7 * D, true, false));
    QAT_TRY(act_alloc(c, KVw, BL, 2 * D, true, false));
  }
  // the two modalities share every weight of the block: when their rows sit back to back (qat_forward_fwd lays them out
  // so) each per-modality contraction is ONE launch over 2*BT rows (one ragged last wave instead of two)
  auto behind = [&](const Act& x, const Act& y) {
    return x.ld == D && y.ld == D && (!x.f || y.f == x.f + BT * D) && (!x.hi || y.hi == x.hi + BT * D) && (!x.lo || y.lo == x.lo + BT * D) &&
           (x.f != nullptr) == (y.f != nullptr) && (x.hi != nullptr) == (y.hi != nullptr) && (x.lo != nullptr) == (y.lo != nullptr);
  };
  const bool stacked = !c.dry ? behind(Xa, Xv) : false;
  if (stacked) {
    QAT_TRY(gemm(c, k_src, Xa, 2 * BT, Wsrc, Epi(), S));
  } else {
    QAT_TRY(gemm(c, k_src, Xa, BT, Wsrc, Epi(), S));
    QAT_TRY(gemm(c, k_src, Xv, BT, Wsrc, Epi(), S.rows(BT)));
  }
  QAT_TRY(gemm(c, k_words, Xw, BL, Wwords, Epi(), KVw));
  // O = [o_qst | o_slf | o_crs] per row, both directions
  QAT_TRY(act_alloc(c, O, 2 * BT, 3 * D, !reads_planes(k_out), reads_planes(k_out)));
  const float scale = 0.125f;  // 64^-0.5
  AttnP p;
  p.nseq = 2 * B; p.H = H; p.Lq = T; p.scale = scale;
  p.q_sn = (long long)T * 7 * D; p.q_si = 7 * D;
  p.o_rn = T; p.o_ri = 1;
  // qst: keys/values = words of sample n % B
  // q / k / v of the three attentions are column slices of S (and of KVw): the same offsets in fp32 or in the plane pairs
  auto set_q = [&](AttnP& a, long long off) { a.q = pp ? nullptr : S.f + off; a.qh = pp ? S.ph + off : nullptr; a.ql = pp ? S.pl + off : nullptr; };
  auto set_kv = [&](AttnP& a, const Act& src, long long koff, long long voff) {
    a.k = pp ? nullptr : src.f + koff; a.v = pp ? nullptr : src.f + voff;
    a.kh = pp ? src.ph + koff : nullptr; a.kl = pp ? src.pl + koff : nullptr;
    a.vh = pp ? src.ph + voff : nullptr; a.vl = pp ? src.pl + voff : nullptr;
  };
  set_q(p, 0); set_kv(p, KVw, 0, D); p.k_sn = (long long)L * 2 * D; p.k_si = 2 * D; p.k_shift = 0; p.k_mod = B; p.Lk = L;
  p.o = O.cols(0); p.attw = nullptr;
  if (attw_audio || attw_video) {
    QAT_REQUIRE(attw_audio && attw_video, "avq: pass both attention-weight outputs or neither");
    // the two outputs are separate tensors: run the two directions as separate launches
    AttnP pa = p;
    pa.nseq = B; pa.attw = attw_audio;
    QAT_TRY(attn(pa));
    AttnP pv = p;
    pv.nseq = B; pv.attw = attw_video; set_q(pv, BT * 7 * D); pv.o = O.rows(BT).cols(0);
    QAT_TRY(attn(pv));
  } else {
    QAT_TRY(attn(p));
  }
  // slf: own keys/values
  p.attw = nullptr;
  set_q(p, D); set_kv(p, S, 2 * D, 3 * D); p.k_sn = p.q_sn; p.k_si = 7 * D; p.k_shift = 0; p.k_mod = 2 * B; p.Lk = T;
  p.o = O.cols(D);
  QAT_TRY(attn(p));
  // crs: keys/values of the OTHER modality (direction 2 sees the original audio, modules.py:257-259)
  set_q(p, 4 * D); set_kv(p, S, 5 * D, 6 * D); p.k_shift = B; p.k_mod = 2 * B;
  p.o = O.cols(2 * D);
  QAT_TRY(attn(p));
  // x1 = LN1(src + [o_qst|o_slf|o_crs] Wout^T + (b_qst+b_slf+b_crs))
  QAT_TRY(act_alloc(c, Y, 2 * BT, D, true, false));
  {
    Epi e;
    e.res = Xa.f; e.ldr = Xa.ld;
    if (stacked) {
      QAT_TRY(gemm(c, k_out, O, 2 * BT, Wout, e, Y));
    } else {
      QAT_TRY(gemm(c, k_out, O, BT, Wout, e, Y));
      e.res = Xv.f; e.ldr = Xv.ld;
      QAT_TRY(gemm(c, k_out, O.rows(BT), BT, Wout, e, Y.rows(BT)));
    }
  }
  QAT_TRY(act_alloc(c, X1, 2 * BT, D, true, reads_planes(k_f1)));
  QAT_TRY(ln(c, Y.f, D, nullptr, 0, nullptr, 0, 1.f, a->norm1, X1, 2 * BT, D));
  // FFN
  QAT_TRY(act_alloc(c, Hh, 2 * BT, D, !reads_planes(k_f2), reads_planes(k_f2)));
  {
    Epi e;
    e.relu = 1;
    QAT_TRY(gemm(c, k_f1, X1, 2 * BT, W1, e, Hh));
    Epi e2;
    e2.res = X1.f; e2.ldr = D;
    QAT_TRY(gemm(c, k_f2, Hh, 2 * BT, W2, e2, Y));
  }
  QAT_TRY(ln(c, Y.f, D, nullptr, 0, nullptr, 0, 1.f, a->norm2, out_a, BT, D));
  QAT_TRY(ln(c, Y.f + BT * D, D, nullptr, 0, nullptr, 0, 1.f, a->norm2, out_v, BT, D));
  return QAT_OK;
}

static int run_avq(const QatAvqArgs* a, Ctx& c) {
  const int D = a->D;
  return avq_core(a, a->B, a->T, a->L, c, act_of(a->audio, D), act_of(a->video, D), act_of(a->words, D), act_of(a->out_audio, D),
                  act_of(a->out_video, D), a->attw_audio, a->attw_video);
}

// =================================================================================================
// P1(patch) + S1  PatchSelecter
// =================================================================================================
// post (optional, qat_forward_fwd): the outputs leave as post + anorm(a) / post + vnorm(v) -- the expert inputs `data + sub_data[i]`
// of the vt TempMoE (modules.py:587,597), whose only consumer they are
static int patch_core(const QatPatchArgs* a, int B, int T, int P, Ctx& c, const float* patch, Act Xaud, Act Xvid, Act out_a, Act out_v,
                      const float* post = nullptr, long long ldp = 0) {
  StageRange nvtx_(c, "qat.patch_selecter");
  const int D = a->D, H = a->H;
  const __half* patch_h = reinterpret_cast<const __half*>(a->patch_f16);
  QAT_REQUIRE(c.dry || patch || patch_h, "patch_select: patch tensor missing");
  QAT_REQUIRE(B >= 1 && T >= 1 && P >= 1 && P <= 32 && D == H * 64, "patch_select: need P<=32 and D==H*64 (P=%d D=%d H=%d)", P, D, H);
  if (a->has_proj) QAT_TRY(check_w(a->proj, D, a->Cin, "patch.proj", c.prec));
  QAT_TRY(check_w(a->slf_in, 3 * D, D, "patch.slf_in", c.prec));
  QAT_TRY(check_w(a->slf_out, D, D, "patch.slf_out", c.prec));
  QAT_TRY(check_w(a->crs_q, D, D, "patch.crs_q", c.prec));
  QAT_TRY(check_w(a->crs_kv, 2 * D, D, "patch.crs_kv", c.prec));
  QAT_TRY(check_w(a->crs_out, D, D, "patch.crs_out", c.prec));
  QAT_TRY(check_w(a->mlp0, D / 2, D, "patch.mlp0", c.prec));
  QAT_TRY(check_w(a->mlp2, D, D / 2, "patch.mlp2", c.prec));
  const long long BT = (long long)B * T;
  const int Cin = a->has_proj ? a->Cin : D;
  // frames per pass: 16 384 (B = 256, T = 60 is ONE pass: fewer launches and ragged last waves than two passes of 8 192, measured
  // 2 % of the step; 1.7 GB of workspace).  Smaller, L2-sized passes were measured slower still
  long long chunk = a->chunk_frames > 0 ? a->chunk_frames : (c.opt.v[OPT_PATCH_CHUNK] > 0 ? c.opt.v[OPT_PATCH_CHUNK] : 16384);
  if (chunk > BT) chunk = BT;
  const float scale = 0.125f;
  const WView Wp = wview(a->proj), Wsi = wview(a->slf_in), Wso = wview(a->slf_out), Wq = wview(a->crs_q), Wkv = wview(a->crs_kv);
  const WView Wco = wview(a->crs_out), Wm0 = wview(a->mlp0), Wm2 = wview(a->mlp2);

  // The whole block is off the router path (its outputs reach the logits only through the vt experts), so in bf16 mode
  // its remaining contractions run single-pass fp16 as well when the fp16 weights are supplied.
  const bool sel16 = c.prec == QAT_PREC_BF16 && BT >= 512 && D % 64 == 0 && c.opt.v[OPT_PATCH_FP32_PLANES] != 1;
  // per-frame query rows [video; audio] -> q2 [BT, 2, D]  (modules.py:692: video first)
  Act Q2, O2, O3, Hm, Mo;
  {
    GemmKind kq = pick_kind(c, BT, D, D, false, Wq);
    if (sel16 && Wq.h16) kq = G_F16;
    // queries as an fp16 plane when the mma.sync cross-attention core will read them (fp16 path, P <= 16)
    const bool q16 = kq == G_F16 && sel16 && Wco.h16 && P <= 16 && H <= 8 && Wsi.h16 && Wso.h16 && Wkv.h16;
    if (q16) QAT_TRY(act_alloc_h(c, Q2, BT, 2 * D));
    else QAT_TRY(act_alloc(c, Q2, BT, 2 * D, true, false));
    QAT_TRY(ensure_operand(c, Xvid, BT, D, kq));
    QAT_TRY(ensure_operand(c, Xaud, BT, D, kq));
    QAT_TRY(gemm(c, kq, Xvid, BT, Wq, Epi(), Q2.cols(0)));
    QAT_TRY(gemm(c, kq, Xaud, BT, Wq, Epi(), Q2.cols(D)));
  }
  GemmKind k_co = pick_kind(c, 2 * BT, D, D, false, Wco), k_m0 = pick_kind(c, 2 * BT, D / 2, D, false, Wm0);
  GemmKind k_m2 = pick_kind(c, 2 * BT, D, D / 2, false, Wm2);
  if (sel16 && Wco.h16) k_co = G_F16;
  if (sel16 && Wm0.h16) k_m0 = G_F16;
  if (sel16 && Wm2.h16) k_m2 = G_F16;
  QAT_TRY(act_alloc_for(c, O2, 2 * BT, D, k_co));

  // bf16 mode, P <= 16: activations between the kernels of a pass are single fp16 planes (half the bytes of
  // fp32, tf32-grade operands at twice the tf32 MMA rate), attention through mma.sync; DESIGN.md section 3.
  const bool f16_path = c.prec == QAT_PREC_BF16 && P <= 16 && H <= 8 && D % 64 == 0 && Wsi.h16 && Wso.h16 && Wkv.h16 &&
                        c.opt.v[OPT_PATCH_FP32_PLANES] != 1;
  const size_t mark = c.mark();
  for (long long f0 = 0; f0 < BT; f0 += chunk) {
    const long long nf = (BT - f0 < chunk) ? BT - f0 : chunk;
    const long long rows = nf * P;
    c.release(mark);
    const float* pin = patch ? patch + (size_t)f0 * P * Cin : nullptr;
    const __half* pin_h = patch_h ? patch_h + (size_t)f0 * P * Cin : nullptr;
    AttnP q;
    q.nseq = (int)nf; q.H = H; q.Lq = 2; q.Lk = P; q.scale = scale;
    q.q = Q2.f + (size_t)f0 * 2 * D; q.q_sn = 2 * D; q.q_si = D;
    q.k_sn = (long long)P * 2 * D; q.k_si = 2 * D; q.k_shift = 0; q.k_mod = (int)nf;
    q.o = O2.rows(2 * f0); q.o_rn = 2; q.o_ri = 1; q.attw = nullptr;
    if (f16_path) {
      Act Pp, QKV, Oa, P2, KV;
      if (pin_h && !a->has_proj) {
        Pp = Act();                                    // fp16 patches that are already projected: the plane is the input itself
        Pp.h = const_cast<__half*>(pin_h);
        Pp.ld = D;
      } else {
        QAT_TRY(act_alloc_h(c, Pp, rows, D));
        if (pin_h) {
          // fp16 feature store: patch_proj as a single kind::f16 pass straight on the stored patches
          QAT_REQUIRE(c.dry || Wp.h16, "patch_select: fp16 patches need patch_proj's fp16 weights (proj.w_f16)");
          QAT_REQUIRE(Cin % 64 == 0, "patch_select: fp16 patches need Cin %% 64 == 0 (Cin=%d)", Cin);
          Act Xh;
          Xh.h = const_cast<__half*>(pin_h);
          Xh.ld = Cin;
          QAT_TRY(gemm(c, G_F16, Xh, rows, Wp, Epi(), Pp));
        } else if (a->has_proj) {
          const GemmKind kp = (Cin % 32 == 0) ? G_TF32 : G_SIMT;
          QAT_TRY(gemm(c, kp, act_of(pin, Cin), rows, Wp, Epi(), Pp));
        } else {
          QAT_TRY(add_rows(c, pin, D, nullptr, 0, Pp, rows, D));
        }
      }
      QAT_TRY(act_alloc_h(c, Oa, rows, D));
      if (Wsi.bias && (128 / P) * P <= tc::GA_QKV_ROWS && c.opt.v[OPT_PATCH_NO_FUSED_ATTN] != 1) {
        // in-projection and the 14-token self-attention in one kernel: [q | k | v] never reaches HBM
        QAT_TRY(gemm_attn_f16(c, Pp.h, Pp.ld, (int)nf, P, H, D, Wsi.h16, Wsi.bias, scale, Oa.h, D));
      } else {
        QAT_TRY(act_alloc_h(c, QKV, rows, 3 * D));
        QAT_TRY(gemm(c, G_F16, Pp, rows, Wsi, Epi(), QKV));
        AttnHP hp;
        hp.q = QKV.h; hp.kv = QKV.h + D; hp.ldq = 3 * D; hp.ldkv = 3 * D;
        hp.o = Oa.h; hp.ldo = D; hp.nseq = (int)nf; hp.H = H; hp.Lq = P; hp.Lk = P; hp.scale = scale;
        QAT_TRY(attention_small_f16(c, hp));
      }
      // p' = p + slf_attn(p)   (:688, no LayerNorm)
      QAT_TRY(act_alloc_h(c, P2, rows, D));
      {
        Epi e;
        e.res_h = Pp.h; e.ldr = Pp.ld;
        QAT_TRY(gemm(c, G_F16, Oa, rows, Wso, e, P2));
      }
      if (Q2.h && O2.h && !O2.f && !O2.hi && Wkv.bias && (128 / P) * P <= tc::GA_QKV_ROWS && c.opt.v[OPT_PATCH_NO_FUSED_CROSS] != 1) {
        // K | V projection and the two-query cross-attention in one kernel: [k | v] never reaches HBM
        QAT_TRY(gemm_attn_cross_f16(c, P2.h, P2.ld, (int)nf, P, H, D, Wkv.h16, Wkv.bias, scale, Q2.h + (size_t)f0 * 2 * D,
                                    O2.h + (size_t)f0 * 2 * D, D));
        continue;
      }
      QAT_TRY(act_alloc_h(c, KV, rows, 2 * D));
      QAT_TRY(gemm(c, G_F16, P2, rows, Wkv, Epi(), KV));
      if (Q2.h && O2.h && !O2.f && !O2.hi) {
        // [video; audio] query pair of every frame against the frame's K | V rows, all fp16 planes, same mma.sync core
        AttnHP cp;
        cp.q = Q2.h + (size_t)f0 * 2 * D; cp.ldq = D; cp.kv = KV.h; cp.ldkv = 2 * D;
        cp.o = O2.h + (size_t)f0 * 2 * D; cp.ldo = D; cp.nseq = (int)nf; cp.H = H; cp.Lq = 2; cp.Lk = P; cp.scale = scale;
        QAT_TRY(attention_small_f16(c, cp));
      } else {
        QAT_TRY(attention_fewq_f16(c, q, KV.h, KV.h + D));
      }
      continue;
    }
    Act Pp, QKV, Oa, P2, KV;
    if (pin_h) {                                       // fp32 mode (or P > 16): widen the fp16 patches of this pass first
      float* wide = reinterpret_cast<float*>(c.alloc((size_t)rows * Cin * 4));
      if (!c.dry && !wide) QAT_FAIL(QAT_EWORKSPACE, "workspace too small (need > %zu bytes, have %zu)", c.off, c.cap);
      QAT_TRY(widen_f16(c, pin_h, wide, rows * (long long)Cin));
      pin = wide;
    }
    if (a->has_proj) {
      QAT_TRY(act_alloc(c, Pp, rows, D, true, false));
      QAT_TRY(gemm(c, pick_kind(c, rows, D, Cin, true, Wp), act_of(pin, Cin), rows, Wp, Epi(), Pp));
    } else {
      Pp = act_of(pin, D);
    }
    QAT_TRY(act_alloc(c, QKV, rows, 3 * D, true, false));
    QAT_TRY(gemm(c, pick_kind(c, rows, 3 * D, D, true, Wsi), Pp, rows, Wsi, Epi(), QKV));
    QAT_TRY(act_alloc(c, Oa, rows, D, true, false));
    AttnP p;
    p.nseq = (int)nf; p.H = H; p.Lq = P; p.Lk = P; p.scale = scale;
    p.q = QKV.f; p.k = QKV.f + D; p.v = QKV.f + 2 * D;
    p.q_sn = (long long)P * 3 * D; p.q_si = 3 * D; p.k_sn = p.q_sn; p.k_si = 3 * D; p.k_shift = 0; p.k_mod = (int)nf;
    p.o = Oa; p.o_rn = P; p.o_ri = 1; p.attw = nullptr;
    QAT_TRY(attention(c, p));
    // p' = p + slf_attn(p)   (:688, no LayerNorm)
    QAT_TRY(act_alloc(c, P2, rows, D, true, false));
    {
      Epi e;
      e.res = Pp.f; e.ldr = Pp.ld;
      QAT_TRY(gemm(c, pick_kind(c, rows, D, D, true, Wso), Oa, rows, Wso, e, P2));
    }
    QAT_TRY(act_alloc(c, KV, rows, 2 * D, true, false));
    QAT_TRY(gemm(c, pick_kind(c, rows, 2 * D, D, true, Wkv), P2, rows, Wkv, Epi(), KV));
    q.k = KV.f; q.v = KV.f + D;
    QAT_TRY(attention(c, q));
  }
  c.release(mark);
  QAT_TRY(act_alloc_for(c, O3, 2 * BT, D, k_m0));
  QAT_TRY(gemm(c, k_co, O2, 2 * BT, Wco, Epi(), O3));
  QAT_TRY(act_alloc_for(c, Hm, 2 * BT, D / 2, k_m2));
  {
    Epi e;
    e.relu = 1;
    QAT_TRY(gemm(c, k_m0, O3, 2 * BT, Wm0, e, Hm));
  }
  QAT_TRY(act_alloc(c, Mo, 2 * BT, D, true, false));
  QAT_TRY(gemm(c, k_m2, Hm, 2 * BT, Wm2, Epi(), Mo));
  // rows (bt,0) = video-related, (bt,1) = audio-related (:706); returned as [anorm(a), vnorm(v)] (:713-716)
  QAT_TRY(ln(c, Mo.f + D, 2 * D, nullptr, 0, nullptr, 0, 1.f, a->anorm, out_a, BT, D, post, ldp));
  QAT_TRY(ln(c, Mo.f, 2 * D, nullptr, 0, nullptr, 0, 1.f, a->vnorm, out_v, BT, D, post, ldp));
  return QAT_OK;
}

static int run_patch(const QatPatchArgs* a, Ctx& c) {
  QAT_REQUIRE(c.dry || ((a->patch || a->patch_f16) && a->audio && a->video && a->out_a && a->out_v), "patch_select: NULL tensor");
  return patch_core(a, a->B, a->T, a->P, c, a->patch, act_of(a->audio, a->D), act_of(a->video, a->D), act_of(a->out_a, a->D),
                    act_of(a->out_v, a->D));
}

// =================================================================================================
// M2-M4 standalone
// =================================================================================================
static int run_gauss(const QatGaussArgs* a, Ctx& c) {
  StageRange nvtx_(c, "qat.gauss_reduce");
  QAT_REQUIRE(a->experts && a->router_logits && a->gauss_raw && a->out, "gauss_reduce: NULL tensor");
  GaussP p;
  p.x = a->experts; p.st = a->stride_t; p.sb = a->stride_b;
  p.rl = a->router_logits; p.gr = a->gauss_raw;
  p.B = a->B; p.T = a->T; p.E = a->E; p.K = a->K; p.C = a->C; p.sigma = a->sigma;
  p.per_expert = 0;
  p.out = act_of(a->out, a->C);
  p.topk_inds = reinterpret_cast<long long*>(a->topk_inds);
  p.topk_probs = a->topk_probs;
  p.gauss_weight = a->gauss_weight;
  QAT_REQUIRE(a->sigma > 0.f, "gauss_reduce: sigma must be positive");
  return gauss_reduce(c, p, a->dtype, a->t_split);
}

// =================================================================================================
// M1-M4 (+LN)  TempMoE
// =================================================================================================
// One TempMoE module inside a front / back pair.  xe_a / xe_v (optional, qat_forward_fwd): the expert inputs data + sub_data[i]
// already formed by the producer of sub_data (sub_a / sub_v are then only markers for "two branches" and are not read);
// xe_stacked: xe_v's rows follow xe_a's in memory, so the two branches, which share every weight, run their expert GEMM as a
// single launch over 2*BT rows.
struct MoeMod {
  const QatMoeArgs* a = nullptr;
  const float* qst = nullptr;      // [B, D] projected question (input of attn_q) ...
  const float* qpre = nullptr;     // ... or attn_q(qst) computed by the caller, row pitch ldq
  int ldq = 0;
  Act Xd;                          // data [B*T, D]: fp32 view + whatever planes its producer emitted
  const float* sub_a = nullptr;
  const float* sub_v = nullptr;
  const Act* xe_a = nullptr;
  const Act* xe_v = nullptr;
  bool xe_stacked = false;
  float* out_a = nullptr;
  float* out_v = nullptr;
  // set by moe_front
  Act RG;                          // [B, 3E] router logits | gauss_pred
  bool fused = false;              // M4 runs in the expert GEMM's epilogue
  float* coef = nullptr;           // [B*T, E]   (fused)
  Act Hb;                          // [nbranch*B, E*Ch + 64]  (fused: bias-column block already written)
  bool w2_tc = false;              // second layer on the split-K tensor-core path (Hb carries bf16 hi / lo planes)
};

// temp_w -> router | gauss_pred for up to two modules per launch sequence (moe_front.cuh), then the per-row coefficients of
// the modules whose M4 runs in the GEMM epilogue
static int moe_front(Ctx& c, MoeMod* mods, int n, int B, int T) {
  StageRange nvtx_(c, "qat.temp_moe.front");
  QAT_REQUIRE(n >= 1 && n <= MF_MAX, "temp_moe: 1 or 2 modules per front end");
  const long long BT = (long long)B * T;
  MoeFrontP fp;
  fp.n = n; fp.B = B; fp.T = T; fp.scale = 0.125f;
  for (int m = 0; m < n; ++m) {
    MoeMod& M = mods[m];
    const QatMoeArgs* a = M.a;
    const int D = a->D, H = a->H, E = a->E, K = a->K;
    QAT_REQUIRE(B >= 1 && T >= 1 && T <= 1024 && D == H * 64, "temp_moe: need T<=1024 and D==H*64 (T=%d D=%d H=%d)", T, D, H);
    QAT_REQUIRE(E >= 1 && E <= GAUSS_MAX_E && K >= 1 && K <= E, "temp_moe: need 1<=K<=E<=%d", GAUSS_MAX_E);
    QAT_REQUIRE((M.sub_a == nullptr) == (M.sub_v == nullptr), "temp_moe: sub_data needs both tensors or none");
    QAT_REQUIRE(m == 0 || (a->D == mods[0].a->D && a->H == mods[0].a->H && a->E == mods[0].a->E && a->K == mods[0].a->K),
                "temp_moe: modules of one front end must share D / H / E / K");
    const int Ch = D / 2, KE = E * Ch + 64;
    QAT_TRY(check_w(a->attn_q, D, D, "moe.attn_q", c.prec));
    QAT_TRY(check_w(a->attn_kv, 2 * D, D, "moe.attn_kv", c.prec));
    QAT_TRY(check_w(a->attn_rg, 3 * E, H * D, "moe.attn_rg", c.prec));
    QAT_REQUIRE(c.dry || a->attn_rg.bias, "moe.attn_rg: the composed bias is missing");
    QAT_TRY(check_w(a->w1_cat, E * Ch, D, "moe.w1_cat", c.prec));
    QAT_TRY(check_w(a->w2_ext, D, KE, "moe.w2_ext", c.prec));
    QAT_REQUIRE(c.dry || (M.out_a && (!M.sub_a || M.out_v)), "temp_moe: output pointer missing");
    QAT_REQUIRE(c.dry || M.Xd.f, "temp_moe: data needs its fp32 view");
    QAT_REQUIRE(M.Xd.ld == D, "temp_moe: data must be dense [B*T, D]");
    Act Q, U;
    if (M.qpre) {
      Q = act_of(M.qpre, M.ldq);
    } else {
      QAT_TRY(act_alloc(c, Q, B, D, true, false));
      QAT_TRY(gemm(c, G_SIMT, act_of(M.qst, D), B, wview(a->attn_q), Epi(), Q));
    }
    QAT_TRY(act_alloc(c, U, B, H * D, true, false));
    QAT_TRY(act_alloc(c, M.RG, B, 3 * E, true, false));
    fp.D = D; fp.H = H; fp.E3 = 3 * E;
    fp.q[m] = Q.f; fp.ldq[m] = Q.ld; fp.wk[m] = a->attn_kv.w; fp.x[m] = M.Xd.f;
    fp.wrg[m] = a->attn_rg.w; fp.brg[m] = a->attn_rg.bias; fp.u[m] = U.f; fp.rg[m] = M.RG.f;
  }
  QAT_TRY(moe_front_launch(c, fp));
  // M4 in the expert GEMM's epilogue: the first-layer GEMM scales its relu'd tile by coef[b,t,e] and sums over
  // the rows of a sample before anything is stored, so the hidden activations [rows, E*Ch] (3 x 110 MB written + read back per
  // 256-question batch) never reach HBM and the gauss_reduce launches disappear.  bf16 mode, 32 <= T <= 128, 256-wide tiles.
  MoeCoefPack pk;
  int nf = 0;
  for (int m = 0; m < n; ++m) {
    MoeMod& M = mods[m];
    const QatMoeArgs* a = M.a;
    const int D = a->D, E = a->E, Ch = D / 2, KE = E * Ch + 64;
    const bool two = M.sub_a != nullptr;
    const GemmKind k_w1 = pick_kind(c, BT, E * Ch, D, false, wview(a->w1_cat));
    M.fused = k_w1 == G_BF16X3 && T >= 32 && T <= 128 && Ch % 32 == 0 && (E * Ch) % 256 == 0 && c.opt.v[OPT_MOE_NO_FUSED_REDUCE] != 1 &&
              c.opt.v[OPT_GEMM_BN] != 128;
    if (!M.fused) continue;
    const int nb = two ? 2 : 1;
    M.coef = reinterpret_cast<float*>(c.alloc((size_t)BT * E * 4));
    if (!c.dry && !M.coef) QAT_FAIL(QAT_EWORKSPACE, "workspace too small (need > %zu bytes, have %zu)", c.off, c.cap);
    // the second layer runs on the split-K tensor-core path when its weights carry planes: its operand then leaves the first
    // layer's epilogue (and the bias block below) as bf16 (hi, lo) planes as well
    M.w2_tc = gemm_splitk_ok((long long)nb * B, D, KE, wview(a->w2_ext).hi, wview(a->w2_ext).lo) && c.opt.v[OPT_MOE_W2_SIMT] != 1;
    QAT_TRY(act_alloc(c, M.Hb, (long long)nb * B, KE, true, M.w2_tc));
    MoeCoefP& q = pk.m[nf++];
    q.g = GaussP();
    q.g.x = nullptr; q.g.st = 0; q.g.sb = 0;
    q.g.rl = M.RG.f; q.g.gr = M.RG.f + E; q.g.rl_ld = 3 * E; q.g.gr_ld = 3 * E; q.g.B = B; q.g.T = T; q.g.E = E; q.g.K = a->K; q.g.C = Ch;
    q.g.sigma = a->sigma; q.g.per_expert = 1; q.g.ring = 0;
    q.g.topk_inds = reinterpret_cast<long long*>(a->topk_inds); q.g.topk_probs = a->topk_probs; q.g.gauss_weight = a->gauss_weight;
    q.coef = M.coef; q.hb = M.Hb.f; q.hbh = M.Hb.hi; q.hbl = M.Hb.lo; q.ldh = KE; q.col0 = E * Ch; q.nrep = nb;
  }
  if (nf) QAT_TRY(moe_coef(c, pk, nf));
  return QAT_OK;
}

// experts -> M4 -> second-layer contraction -> LayerNorm of one module (modules.py:585-612)
static int moe_back(Ctx& c, MoeMod& M, int B, int T) {
  StageRange nvtx_(c, "qat.temp_moe.experts");
  const QatMoeArgs* a = M.a;
  const int D = a->D, E = a->E, K = a->K;
  const int Ch = D / 2, KE = E * Ch + 64;
  const long long BT = (long long)B * T;
  const bool two = M.sub_a != nullptr;
  const WView W1 = wview(a->w1_cat), W2 = wview(a->w2_ext);
  const GemmKind k_w1 = pick_kind(c, BT, E * Ch, D, false, W1);
  const Act* xe_a = M.xe_a;
  const Act* xe_v = M.xe_v;
  const float* sub_a = M.sub_a;
  const float* sub_v = M.sub_v;
  float* out_a = M.out_a;
  float* out_v = M.out_v;
  const Act& RG = M.RG;
  Act Xd = M.Xd;
  if (!two) QAT_TRY(ensure_operand(c, Xd, BT, D, k_w1));
  if (M.fused) {
    const int nb = two ? 2 : 1;
    const bool stacked = two && xe_a && xe_v && M.xe_stacked;
    Act Hb = M.Hb, Ob;
    Epi e;
    e.relu = 1;
    e.coef = M.coef; e.coef_E = E; e.coef_T = T; e.coef_Ch = Ch; e.coef_mod = BT;
    if (stacked) {
      QAT_TRY(gemm(c, k_w1, *xe_a, 2 * BT, W1, e, Hb));
    } else {
      for (int br = 0; br < nb; ++br) {
        const size_t mark = c.mark();
        Act Xe = Xd;
        if (two && xe_a && xe_v) {
          Xe = br == 0 ? *xe_a : *xe_v;
        } else if (two) {
          QAT_TRY(act_alloc(c, Xe, BT, D, false, true));
          QAT_TRY(add_rows(c, Xd.f, Xd.ld, br == 0 ? sub_a : sub_v, D, Xe, BT, D));     // :587,597
        }
        QAT_TRY(gemm(c, k_w1, Xe, BT, W1, e, Hb.rows((long long)br * B)));
        c.release(mark);
      }
    }
    QAT_TRY(act_alloc(c, Ob, (long long)nb * B, D, true, false));
    {
      // second expert layer + biases as one [nb*B, E*Ch+64] x [D, E*Ch+64]^T contraction.  It comes after the router, so 16-bit
      // operands are enough; at K = 1856 the CUDA-core kernel (shared-memory bound at its 4 x 4 register tile) took 23 + 35 us
      // per step for the two modules
      Epi e2;
      WView w2 = W2;
      w2.bias = nullptr;
      if (M.w2_tc) QAT_TRY(gemm_splitk_bf16x3(c, Hb.hi, Hb.lo, Hb.ld, (long long)nb * B, w2.hi, w2.lo, w2.n, w2.k, e2, Ob));
      else QAT_TRY(gemm(c, G_SIMT, Hb, (long long)nb * B, w2, e2, Ob));
    }
    QAT_TRY(ln(c, Ob.f, D, nullptr, 0, nullptr, 0, 1.f, a->norm_a, act_of(out_a, D), B, D));
    if (two) QAT_TRY(ln(c, Ob.f + (size_t)B * D, D, nullptr, 0, nullptr, 0, 1.f, a->norm_v, act_of(out_v, D), B, D));
    return QAT_OK;
  }
  // fp32 mode (and shapes outside the fused epilogue): first-layer GEMM writes the hidden activations, the streaming kernel reduces them
  const int nbranch = two ? 2 : 1;
  for (int br = 0; br < nbranch; ++br) {
    const size_t mark = c.mark();
    Act Xe = Xd;
    if (two && xe_a && xe_v) {
      Xe = br == 0 ? *xe_a : *xe_v;
    } else if (two) {
      QAT_TRY(act_alloc(c, Xe, BT, D, !reads_planes(k_w1), reads_planes(k_w1)));
      QAT_TRY(add_rows(c, Xd.f, Xd.ld, br == 0 ? sub_a : sub_v, D, Xe, BT, D));     // :587,597
    }
    // hidden = relu(W1_e x + b1_e) for every expert: [BT, E*Ch]
    Act Hd, Hb, Ob;
    QAT_TRY(act_alloc(c, Hd, BT, E * Ch, true, false));
    {
      Epi e;
      e.relu = 1;
      QAT_TRY(gemm(c, k_w1, Xe, BT, W1, e, Hd));
    }
    // hbar[b, e, :] = sum_t coef[b,t,e] hidden[b,t,e,:]  |  csum[b,e] = sum_t coef[b,t,e]
    QAT_TRY(act_alloc(c, Hb, B, KE, true, false));
    GaussP g;
    g.x = Hd.f; g.st = (long long)E * Ch; g.sb = (long long)T * E * Ch;
    g.rl = RG.f; g.gr = RG.f + E; g.rl_ld = 3 * E; g.gr_ld = 3 * E; g.B = B; g.T = T; g.E = E; g.K = K; g.C = Ch; g.sigma = a->sigma;
    g.per_expert = 1; g.out = Hb;
    g.topk_inds = br == 0 ? reinterpret_cast<long long*>(a->topk_inds) : nullptr;
    g.topk_probs = br == 0 ? a->topk_probs : nullptr;
    g.gauss_weight = br == 0 ? a->gauss_weight : nullptr;
    QAT_TRY(gauss_reduce(c, g, QAT_DT_F32, 0));
    // out = sum_e W2_e hbar_e + b2_e csum_e : one [B, E*Ch+64] x [D, E*Ch+64]^T contraction, no bias
    QAT_TRY(act_alloc(c, Ob, B, D, true, false));
    {
      Epi e;
      WView w2 = W2;
      w2.bias = nullptr;
      QAT_TRY(gemm(c, G_SIMT, Hb, B, w2, e, Ob));
    }
    const QatLayerNormW& nw = (br == 0) ? a->norm_a : a->norm_v;
    QAT_TRY(ln(c, Ob.f, D, nullptr, 0, nullptr, 0, 1.f, nw, act_of(br == 0 ? out_a : out_v, D), B, D));
    c.release(mark);
  }
  return QAT_OK;
}

static int moe_core(const QatMoeArgs* a, int B, int T, Ctx& c, const float* qst, Act Xd, const float* sub_a, const float* sub_v, float* out_a,
                    float* out_v) {
  MoeMod M;
  M.a = a; M.qst = qst; M.Xd = Xd; M.sub_a = sub_a; M.sub_v = sub_v; M.out_a = out_a; M.out_v = out_v;
  QAT_TRY(moe_front(c, &M, 1, B, T));
  return moe_back(c, M, B, T);
}

static int run_moe(const QatMoeArgs* a, Ctx& c) {
  QAT_REQUIRE(c.dry || (a->qst && a->data), "temp_moe: NULL tensor");
  return moe_core(a, a->B, a->T, c, a->qst, act_of(a->data, a->D), a->sub_a, a->sub_v, a->out_a, a->out_v);
}

// =================================================================================================
// G1 (+H1)  QstGrounding on a length-2 sequence, optional head
// =================================================================================================
// qpre (optional): the query projection attn_q(qst) computed by the caller, row pitch ldq (qat_forward_fwd batches the
// query projections of every block that attends from the question into one contraction)
static int ground_core(const QatGroundArgs* a, int B, Ctx& c, const float* qst, const float* d0, const float* d1, float* out, bool with_head,
                       float* logits, const float* qpre = nullptr, int ldq = 0) {
  StageRange nvtx_(c, "qat.qst_grounding");
  const int D = a->D, H = a->H;
  QAT_REQUIRE(B >= 1 && D == H * 64, "qst_grounding: need D==H*64");
  QAT_TRY(check_w(a->attn_q, D, D, "ground.attn_q", c.prec));
  QAT_TRY(check_w(a->attn_kv, 2 * D, D, "ground.attn_kv", c.prec));
  QAT_TRY(check_w(a->attn_out, D, D, "ground.attn_out", c.prec));
  QAT_TRY(check_w(a->mlp0, D / 2, D, "ground.mlp0", c.prec));
  QAT_TRY(check_w(a->mlp2, D, D / 2, "ground.mlp2", c.prec));
  Act KV, Q, O, A1, Hm, Mo;
  QAT_TRY(act_alloc(c, KV, 2ll * B, 2 * D, true, false));
  AttnP p;
  if (d1 == d0 + (size_t)B * D) {
    // the two length-1 inputs sit back to back: one contraction over [2B, D]; key i of sample b is row i*B + b.  The block comes
    // after the router, so in bf16 mode its largest contraction runs on the split-K tensor-core path (16-bit-mantissa operands;
    // 24 -> 12 us per pass at B = 256), after a plane-conversion pass over the 2B input rows
    const WView Wkv = wview(a->attn_kv);
    if (c.prec == QAT_PREC_BF16 && gemm_splitk_ok(2ll * B, 2 * D, D, Wkv.hi, Wkv.lo) && 2ll * B >= 256 && c.opt.v[OPT_MOE_W2_SIMT] != 1) {
      Act In = act_of(d0, D);
      QAT_TRY(ensure_operand(c, In, 2ll * B, D, G_BF16X3));
      Epi e;
      e.bias = Wkv.bias;
      QAT_TRY(gemm_splitk_bf16x3(c, In.hi, In.lo, D, 2ll * B, Wkv.hi, Wkv.lo, Wkv.n, Wkv.k, e, KV));
    } else {
      QAT_TRY(gemm(c, G_SIMT, act_of(d0, D), 2ll * B, Wkv, Epi(), KV));
    }
    p.k_sn = 2 * D; p.k_si = (long long)B * 2 * D;
  } else {
    Act kv0 = KV, kv1 = KV;                                      // [B, 2, 2D]
    kv0.ld = 4 * D;
    kv1.ld = 4 * D;
    kv1.f += 2 * D;
    QAT_TRY(gemm(c, G_SIMT, act_of(d0, D), B, wview(a->attn_kv), Epi(), kv0));
    QAT_TRY(gemm(c, G_SIMT, act_of(d1, D), B, wview(a->attn_kv), Epi(), kv1));
    p.k_sn = 4 * D; p.k_si = 2 * D;
  }
  if (qpre) {
    Q = act_of(qpre, ldq);
  } else {
    QAT_TRY(act_alloc(c, Q, B, D, true, false));
    QAT_TRY(gemm(c, G_SIMT, act_of(qst, D), B, wview(a->attn_q), Epi(), Q));
  }
  QAT_TRY(act_alloc(c, O, B, D, true, false));
  p.nseq = B; p.H = H; p.Lq = 1; p.Lk = 2; p.scale = 0.125f;
  p.q = Q.f; p.q_sn = Q.ld; p.q_si = Q.ld;
  p.k = KV.f; p.v = KV.f + D; p.k_shift = 0; p.k_mod = B;
  p.o = O; p.o_rn = 1; p.o_ri = 1; p.attw = nullptr;
  QAT_TRY(attention(c, p));
  QAT_TRY(act_alloc(c, A1, B, D, true, false));
  QAT_TRY(gemm(c, G_SIMT, O, B, wview(a->attn_out), Epi(), A1));
  QAT_TRY(act_alloc(c, Hm, B, D / 2, true, false));
  {
    Epi e;
    e.relu = 1;
    QAT_TRY(gemm(c, G_SIMT, A1, B, wview(a->mlp0), e, Hm));
  }
  QAT_TRY(act_alloc(c, Mo, B, D, true, false));
  QAT_TRY(gemm(c, G_SIMT, Hm, B, wview(a->mlp2), Epi(), Mo));
  // feat = mean_seq(data) + mlp(attn)  -> LN   (:344-346)
  QAT_TRY(ln(c, d0, D, d1, D, Mo.f, D, 0.5f, a->norm, act_of(out, D), B, D));
  if (with_head) {
    QAT_REQUIRE((c.dry || logits) && a->head.w && a->head.k == D, "qst_grounding: head weights/logits missing");
    // head(relu(fusion)) net.py:141-142: relu folded into a tiny elementwise pass on the LN output
    Act R;
    QAT_TRY(act_alloc(c, R, B, D, true, false));
    if (!c.dry) {
      QAT_TRY(relu_rows(c, out, R.f, (long long)B * D));
    }
    QAT_TRY(gemm(c, G_SIMT, R, B, wview(a->head), Epi(), act_of(logits, a->head.n)));
  }
  return QAT_OK;
}

// QstGrounding over any sequence: data [B, S, D] (the concatenated list of modules.py:318-326).  Off the eval path (both call
// sites have S = 2, ground_core above): plain composition of the exported pieces.
static int ground_general(const QatGroundArgs* a, Ctx& c) {
  StageRange nvtx_(c, "qat.qst_grounding");
  const int B = a->B, S = a->S, D = a->D, H = a->H;
  QAT_REQUIRE(B >= 1 && S >= 1 && D == H * 64, "qst_grounding: need B, S >= 1 and D==H*64");
  QAT_TRY(check_w(a->attn_q, D, D, "ground.attn_q", c.prec));
  QAT_TRY(check_w(a->attn_kv, 2 * D, D, "ground.attn_kv", c.prec));
  QAT_TRY(check_w(a->attn_out, D, D, "ground.attn_out", c.prec));
  QAT_TRY(check_w(a->mlp0, D / 2, D, "ground.mlp0", c.prec));
  QAT_TRY(check_w(a->mlp2, D, D / 2, "ground.mlp2", c.prec));
  const long long BS = (long long)B * S;
  Act KV, Q, O, A1, Hm, Mo, Mean;
  QAT_TRY(act_alloc(c, KV, BS, 2 * D, true, false));
  QAT_TRY(linear_core(c, a->attn_kv, act_of(a->data, D), BS, 0, KV));
  QAT_TRY(act_alloc(c, Q, B, D, true, false));
  QAT_TRY(gemm(c, G_SIMT, act_of(a->qst, D), B, wview(a->attn_q), Epi(), Q));
  QAT_TRY(act_alloc(c, O, B, D, true, false));
  AttnP p;
  p.nseq = B; p.H = H; p.Lq = 1; p.Lk = S; p.scale = 0.125f;
  p.q = Q.f; p.q_sn = Q.ld; p.q_si = Q.ld;
  p.k = KV.f; p.v = KV.f + D; p.k_sn = (long long)S * 2 * D; p.k_si = 2 * D; p.k_shift = 0; p.k_mod = B;
  p.o = O; p.o_rn = 1; p.o_ri = 1; p.attw = nullptr;
  QAT_TRY(attention(c, p));
  QAT_TRY(act_alloc(c, A1, B, D, true, false));
  QAT_TRY(gemm(c, G_SIMT, O, B, wview(a->attn_out), Epi(), A1));
  QAT_TRY(act_alloc(c, Hm, B, D / 2, true, false));
  {
    Epi e;
    e.relu = 1;
    QAT_TRY(gemm(c, G_SIMT, A1, B, wview(a->mlp0), e, Hm));
  }
  QAT_TRY(act_alloc(c, Mo, B, D, true, false));
  QAT_TRY(gemm(c, G_SIMT, Hm, B, wview(a->mlp2), Epi(), Mo));
  // feat = mean_seq(data) + mlp(attn) -> LN   (:344-346)
  QAT_TRY(act_alloc(c, Mean, B, D, true, false));
  QAT_TRY(mean_seq(c, a->data, Mean.f, B, S, D));
  QAT_TRY(ln(c, Mean.f, D, nullptr, 0, Mo.f, D, 1.f, a->norm, act_of(a->out, D), B, D));
  if (a->has_head) {
    QAT_REQUIRE((c.dry || a->logits) && a->head.w && a->head.k == D, "qst_grounding: head weights/logits missing");
    Act R;
    QAT_TRY(act_alloc(c, R, B, D, true, false));
    if (!c.dry) QAT_TRY(relu_rows(c, a->out, R.f, (long long)B * D));
    QAT_TRY(gemm(c, G_SIMT, R, B, wview(a->head), Epi(), act_of(a->logits, a->head.n)));
  }
  return QAT_OK;
}

static int run_ground(const QatGroundArgs* a, Ctx& c) {
  if (a->data) {
    QAT_REQUIRE(c.dry || (a->qst && a->out), "qst_grounding: NULL tensor");
    return ground_general(a, c);
  }
  QAT_REQUIRE(c.dry || (a->qst && a->d0 && a->d1 && a->out), "qst_grounding: NULL tensor");
  return ground_core(a, a->B, c, a->qst, a->d0, a->d1, a->out, a->has_head != 0, a->logits);
}

// =================================================================================================
// The whole hot path in ONE call: net.py:120-142.  Intermediates live in the workspace and travel between the blocks
// with the storage planes their consumers read (no conversion passes at module boundaries, one launch sequence).
// =================================================================================================
static int run_forward(const QatForwardArgs* a, Ctx& c) {
  StageRange nvtx_(c, "qat.forward");
  const int B = a->B, T = a->T, L = a->L, P = a->P, D = a->avq.D;
  QAT_REQUIRE(B >= 1 && T >= 1 && L >= 1 && P >= 1, "forward: bad shape");
  QAT_REQUIRE(c.dry || (a->quest && a->words && a->audio && a->video && (a->patch || a->patch_sel.patch_f16) && a->logits), "forward: NULL tensor");
  QAT_REQUIRE(a->patch_sel.D == D && a->at.D == D && a->vt.D == D && a->ground.D == D && a->audio_proj.n == D && a->video_proj.n == D &&
                  a->words_proj.n == D && a->quest_proj.n == D, "forward: every block must share d_model=%d", D);
  QAT_REQUIRE(a->patch_sel.has_proj, "forward: patch_proj weights missing (patch_sel.has_proj)");
  const long long BT = (long long)B * T, BL = (long long)B * L;
  const bool tc = c.prec == QAT_PREC_BF16;
  // projected inputs: fp32 (residuals) + bf16 hi/lo planes (the AVQ in-projections read those in bf16 mode)
  Act A0, V0, W0, Q0, Ax, Vx, Ap, Vp;
  // audio rows then video rows of ONE activation: the AVQ block then runs its per-modality contractions as single launches
  {
    Act AV0;
    QAT_TRY(act_alloc(c, AV0, 2 * BT, D, true, tc));
    A0 = AV0;
    V0 = AV0.rows(BT);
  }
  QAT_TRY(act_alloc(c, W0, BL, D, true, tc));
  QAT_TRY(act_alloc(c, Q0, B, D, true, false));
  QAT_TRY(linear_core(c, a->audio_proj, act_of(a->audio, a->audio_proj.k), BT, 0, A0));
  QAT_TRY(linear_core(c, a->video_proj, act_of(a->video, a->video_proj.k), BT, 0, V0));
  QAT_TRY(linear_core(c, a->words_proj, act_of(a->words, a->words_proj.k), BL, 0, W0));
  QAT_TRY(linear_core(c, a->quest_proj, act_of(a->quest, a->quest_proj.k), B, 0, Q0));
  // AVQ outputs: fp32 + bf16 planes (TempMoE K/V and expert GEMMs) + fp16 plane (PatchSelecter query projection)
  QAT_TRY(act_alloc(c, Ax, BT, D, true, tc));
  QAT_TRY(act_alloc(c, Vx, BT, D, true, tc));
  if (tc) {
    Act h;
    QAT_TRY(act_alloc_h(c, h, BT, D)); Ax.h = h.h;
    QAT_TRY(act_alloc_h(c, h, BT, D)); Vx.h = h.h;
  }
  // PatchSelecter outputs: their only consumer is the vt TempMoE, as `video + patch feature` (modules.py:587,597), so the two
  // LayerNorms that end the block write that sum, in the storage form the expert GEMM reads
  const GemmKind k_w1vt = pick_kind(c, BT, a->vt.E * (D / 2), D, false, wview(a->vt.w1_cat));
  {
    Act APV;                                              // audio-related rows, then video-related rows
    QAT_TRY(act_alloc(c, APV, 2 * BT, D, !reads_planes(k_w1vt), reads_planes(k_w1vt)));
    Ap = APV;
    Vp = APV.rows(BT);
  }
  // ap_global | vp_global | fusion1 | a_global | fusion2: both QstGrounding passes find their two inputs back to back
  float* glob = reinterpret_cast<float*>(c.alloc((size_t)5 * B * D * 4));
  if (!c.dry && !glob) QAT_FAIL(QAT_EWORKSPACE, "workspace too small (need > %zu bytes, have %zu)", c.off, c.cap);
  float* ap_global = glob;
  float* vp_global = glob + (size_t)B * D;
  float* fusion1 = glob + (size_t)2 * B * D;
  float* a_global = glob + (size_t)3 * B * D;
  float* fusion2 = glob + (size_t)4 * B * D;
  // the three blocks that attend FROM the question project it with different weights: one [B, D] x [3D, D]^T contraction
  // (q_cat = at.attn_q | vt.attn_q | ground.attn_q) instead of four launches (ground runs twice on the same query)
  const float* q_at = nullptr;
  const float* q_vt = nullptr;
  const float* q_gr = nullptr;
  int ldq = 0;
  if (a->q_cat.w) {
    QAT_TRY(check_w(a->q_cat, 3 * D, D, "forward.q_cat", c.prec));
    Act Qc;
    QAT_TRY(act_alloc(c, Qc, B, 3 * D, true, false));
    QAT_TRY(gemm(c, G_SIMT, Q0, B, wview(a->q_cat), Epi(), Qc));
    q_at = Qc.f; q_vt = Qc.f + D; q_gr = Qc.f + 2 * D; ldq = 3 * D;
  }
  const size_t mark = c.mark();
  QAT_TRY(avq_core(&a->avq, B, T, L, c, A0, V0, W0, Ax, Vx, nullptr, nullptr));
  c.release(mark);
  QAT_TRY(patch_core(&a->patch_sel, B, T, P, c, a->patch, Ax, Vx, Ap, Vp, Vx.f, Vx.ld));
  c.release(mark);
  {
    // both TempMoE modules share one front end (question attention -> router | gauss_pred -> coefficients: three launches)
    MoeMod mm[2];
    mm[0].a = &a->at; mm[0].qst = Q0.f; mm[0].qpre = q_at; mm[0].ldq = ldq; mm[0].Xd = Ax; mm[0].out_a = a_global;
    mm[1].a = &a->vt; mm[1].qst = Q0.f; mm[1].qpre = q_vt; mm[1].ldq = ldq; mm[1].Xd = Vx; mm[1].sub_a = Vx.f; mm[1].sub_v = Vx.f;
    mm[1].xe_a = &Ap; mm[1].xe_v = &Vp; mm[1].xe_stacked = true; mm[1].out_a = ap_global; mm[1].out_v = vp_global;
    if (a->at.E == a->vt.E && a->at.K == a->vt.K) {
      QAT_TRY(moe_front(c, mm, 2, B, T));
    } else {
      QAT_TRY(moe_front(c, &mm[0], 1, B, T));
      QAT_TRY(moe_front(c, &mm[1], 1, B, T));
    }
    QAT_TRY(moe_back(c, mm[0], B, T));
    QAT_TRY(moe_back(c, mm[1], B, T));
  }
  c.release(mark);
  QAT_TRY(ground_core(&a->ground, B, c, Q0.f, ap_global, vp_global, fusion1, false, nullptr, q_gr, ldq));
  c.release(mark);
  QAT_TRY(ground_core(&a->ground, B, c, Q0.f, fusion1, a_global, fusion2, true, a->logits, q_gr, ldq));
  return QAT_OK;
}


// =================================================================================================
// N1  CLIP text tower (frozen), encoders.py:30-37 / clip_base.py:186-254
// =================================================================================================
static int run_clip_text(const QatClipTextArgs* a, Ctx& c) {
  StageRange nvtx_(c, "qat.clip_text");
  const int B = a->B, L = a->L, W = a->W, H = a->H, E = a->E;
  QAT_REQUIRE(B >= 1 && L >= 1 && L <= 128 && H >= 1 && W == H * 64 && E >= 1, "clip_text: need W == H*64 and L <= 128 (W=%d H=%d L=%d)", W, H, L);
  QAT_REQUIRE(a->n_layers >= 0 && (a->n_layers == 0 || a->layers), "clip_text: layer weights missing");
  QAT_REQUIRE(c.dry || (a->tokens && a->token_embedding && a->positional_embedding && a->quest && a->words), "clip_text: NULL tensor");
  QAT_TRY(check_w(a->text_projection, E, W, "clip_text.text_projection", c.prec));
  const long long BL = (long long)B * L;
  Act X, Y, QKV, O, Hh, G;
  QAT_TRY(act_alloc(c, X, BL, W, true, false));
  if (!c.dry) {
    const long long n4 = BL * (W / 4);
    clip_embed_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, c.stream>>>(a->tokens, a->token_embedding, a->positional_embedding, X.f, BL, L,
                                                                           W / 4, a->vocab);
    QAT_TRY(check_launch("clip_embed_kernel"));
  }
  for (int l = 0; l < a->n_layers; ++l) {
    const QatClipLayerW& w = a->layers[l];
    QAT_TRY(check_w(w.in_proj, 3 * W, W, "clip_text.in_proj", c.prec));
    QAT_TRY(check_w(w.out_proj, W, W, "clip_text.out_proj", c.prec));
    QAT_TRY(check_w(w.c_fc, 4 * W, W, "clip_text.c_fc", c.prec));
    QAT_TRY(check_w(w.c_proj, W, 4 * W, "clip_text.c_proj", c.prec));
    const WView Wi = wview(w.in_proj), Wo = wview(w.out_proj), Wf = wview(w.c_fc), Wp = wview(w.c_proj);
    const GemmKind k_i = pick_kind(c, BL, 3 * W, W, false, Wi), k_o = pick_kind(c, BL, W, W, false, Wo);
    const GemmKind k_f = pick_kind(c, BL, 4 * W, W, false, Wf), k_p = pick_kind(c, BL, W, 4 * W, false, Wp);
    const size_t mark = c.mark();
    // x = x + out_proj(attn(ln_1 x))
    QAT_TRY(act_alloc_for(c, Y, BL, W, k_i));
    QAT_TRY(ln(c, X.f, W, nullptr, 0, nullptr, 0, 1.f, w.ln_1, Y, BL, W));
    // bf16 mode: [q | k | v] leaves the GEMM as an fp16 (hi, lo) plane pair, the form the attention core eats
    const bool pp = k_i == G_BF16X3 && W % 8 == 0;
    if (pp) QAT_TRY(act_alloc_pp(c, QKV, BL, 3 * W));
    else QAT_TRY(act_alloc(c, QKV, BL, 3 * W, true, false));
    QAT_TRY(gemm(c, k_i, Y, BL, Wi, Epi(), QKV));
    QAT_TRY(act_alloc_for(c, O, BL, W, k_o));
    AttnP p;
    p.nseq = B; p.H = H; p.Lq = L; p.Lk = L; p.scale = 0.125f; p.causal = 1;
    if (pp) {
      p.q = p.k = p.v = nullptr;
      p.qh = QKV.ph; p.ql = QKV.pl; p.kh = QKV.ph + W; p.kl = QKV.pl + W; p.vh = QKV.ph + 2 * W; p.vl = QKV.pl + 2 * W;
    } else {
      p.q = QKV.f; p.k = QKV.f + W; p.v = QKV.f + 2 * W;
    }
    p.q_sn = (long long)L * 3 * W; p.q_si = 3 * W; p.k_sn = p.q_sn; p.k_si = 3 * W; p.k_shift = 0; p.k_mod = B;
    p.o = O; p.o_rn = L; p.o_ri = 1; p.attw = nullptr;
    QAT_TRY(attention(c, p));
    {
      Epi e;
      e.res = X.f; e.ldr = W;
      QAT_TRY(gemm(c, k_o, O, BL, Wo, e, X));           // in place: a tile's residual is read before the tile is stored
    }
    c.release(mark);
    // x = x + c_proj(QuickGELU(c_fc(ln_2 x)))
    QAT_TRY(act_alloc_for(c, Y, BL, W, k_f));
    QAT_TRY(ln(c, X.f, W, nullptr, 0, nullptr, 0, 1.f, w.ln_2, Y, BL, W));
    QAT_TRY(act_alloc_for(c, Hh, BL, 4 * W, k_p));
    {
      Epi e;
      e.relu = 2;
      QAT_TRY(gemm(c, k_f, Y, BL, Wf, e, Hh));
      Epi e2;
      e2.res = X.f; e2.ldr = W;
      QAT_TRY(gemm(c, k_p, Hh, BL, Wp, e2, X));
    }
    c.release(mark);
  }
  // words = ln_final(x); quest = words[eot] @ text_projection
  QAT_TRY(ln(c, X.f, W, nullptr, 0, nullptr, 0, 1.f, a->ln_final, act_of(a->words, W), BL, W));
  QAT_TRY(act_alloc(c, G, B, W, true, false));
  if (!c.dry) {
    clip_eot_gather_kernel<<<B, 256, 0, c.stream>>>(a->tokens, a->words, G.f, L, W / 4);
    QAT_TRY(check_launch("clip_eot_gather_kernel"));
  }
  {
    WView wp = wview(a->text_projection);
    wp.bias = nullptr;
    QAT_TRY(gemm(c, G_SIMT, G, B, wp, Epi(), act_of(a->quest, E)));
  }
  return QAT_OK;
}

// =================================================================================================
// T1-T3, T0
// =================================================================================================
static int run_tome_match(const QatTomeMatchArgs* a, Ctx& c) {
  StageRange nvtx_(c, "qat.tome.match");
  const int r = tome_clamp_r(a->t, a->r, a->class_token);
  QAT_REQUIRE(r > 0, "tome_match: nothing to merge (r clamps to 0 for t=%d)", a->t);
  QAT_REQUIRE(a->metric && a->unm_idx && a->src_idx && a->dst_idx, "tome_match: NULL tensor");
  return tome_match(c, a->metric, a->N, a->t, a->C, r, a->class_token, a->unm_idx, a->src_idx, a->dst_idx);
}

static int run_tome_merge(const QatTomeMergeArgs* a, Ctx& c) {
  StageRange nvtx_(c, "qat.tome.merge");
  QAT_REQUIRE(a->x && a->unm_idx && a->src_idx && a->dst_idx && a->x_out, "tome_merge: NULL tensor");
  return tome_merge(c, a->x, a->size, a->N, a->t, a->C, a->r, a->mode, a->unm_idx, a->src_idx, a->dst_idx, a->x_out,
                    a->size_out, a->group, a->T0, a->distill_token != 0);
}

static int run_tome_chain(const QatTomeChainArgs* a, Ctx& c) {
  StageRange nvtx_(c, "qat.tome.chain");
  const int N = a->N, C = a->C;
  QAT_REQUIRE(a->n_steps >= 0 && (a->n_steps == 0 || a->r_list), "tome_chain: r_list missing");
  QAT_REQUIRE(C % 64 == 0, "tome_chain: C=%d must be a multiple of 64 (stand-in metric)", C);
  int t = a->t0;
  // ping-pong buffers sized for the first step's output
  float* xb[2] = {nullptr, nullptr};
  float* sb[2] = {nullptr, nullptr};
  for (int i = 0; i < 2; ++i) {
    xb[i] = reinterpret_cast<float*>(c.alloc((size_t)N * a->t0 * C * 4));
    sb[i] = reinterpret_cast<float*>(c.alloc((size_t)N * a->t0 * 4));
  }
  float* metric[2];
  for (int i = 0; i < 2; ++i) metric[i] = reinterpret_cast<float*>(c.alloc((size_t)N * a->t0 * 64 * 4));
  int mp = 0;
  bool have_metric = false;
  int* idx = reinterpret_cast<int*>(c.alloc((size_t)N * ((a->t0 + 1) / 2 + 2 * (a->t0 / 2)) * 4));
  if (!c.dry && (!xb[1] || !sb[1] || !metric[1] || !idx)) QAT_FAIL(QAT_EWORKSPACE, "tome_chain: workspace too small (need %zu)", c.peak);
  if (a->group && !c.dry) {
    long long tot = (long long)N * a->t0;
    iota_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c.stream>>>(a->group, N, a->t0);
    QAT_TRY(check_launch("iota_rows_kernel"));
  }
  const float* xcur = a->x;
  const float* scur = nullptr;
  int pp = 0;
  size_t trace_off = 0;
  for (int s = 0; s < a->n_steps; ++s) {
    const int r = tome_clamp_r(t, a->r_list[s], a->class_token);
    if (r <= 0) continue;
    const int ta = (t + 1) / 2, n_unm = ta - r;
    int* unm = idx;
    int* src = unm + (size_t)N * n_unm;
    int* dst = src + (size_t)N * r;
    if (a->idx_trace && !c.dry) {
      unm = a->idx_trace + trace_off;
      src = unm + (size_t)N * n_unm;
      dst = src + (size_t)N * r;
    }
    trace_off += (size_t)N * (n_unm + 2 * r);
    // the metric of step s+1 comes out of step s's merge pass (fused); only the first step reads x for it
    const bool fused = tome_merge_wavg_supported(C);
    if (!fused || !have_metric) QAT_TRY(tome_metric(c, xcur, metric[mp], (long long)N * t, C));
    QAT_TRY(tome_match(c, metric[mp], N, t, 64, r, a->class_token, unm, src, dst));
    if (fused) {
      QAT_TRY(tome_merge_wavg(c, xcur, scur, N, t, C, r, unm, src, dst, xb[pp], sb[pp], metric[mp ^ 1], a->group, a->t0, (a->class_token >> 1) & 1));
      mp ^= 1;
      have_metric = true;
    } else {
      QAT_TRY(tome_merge(c, xcur, scur, N, t, C, r, QAT_TOME_WAVG, unm, src, dst, xb[pp], sb[pp], a->group, a->t0, (a->class_token >> 1) & 1));
    }
    xcur = xb[pp];
    scur = sb[pp];
    pp ^= 1;
    t -= r;
  }
  if (!c.dry) {
    QAT_REQUIRE(a->x_out && a->size_out, "tome_chain: output pointer missing");
    if (scur == nullptr) {
      QAT_CUDA(cudaMemcpyAsync(a->x_out, a->x, (size_t)N * t * C * 4, cudaMemcpyDeviceToDevice, c.stream));
      ones_kernel<<<(unsigned)(((long long)N * t + 255) / 256), 256, 0, c.stream>>>(a->size_out, (long long)N * t);
      QAT_TRY(check_launch("ones_kernel"));
    } else {
      QAT_CUDA(cudaMemcpyAsync(a->x_out, xcur, (size_t)N * t * C * 4, cudaMemcpyDeviceToDevice, c.stream));
      QAT_CUDA(cudaMemcpyAsync(a->size_out, scur, (size_t)N * t * 4, cudaMemcpyDeviceToDevice, c.stream));
    }
  }
  return QAT_OK;
}

// =================================================================================================
// N4  one ToMe-patched ViT block (tome/patch/timm.py:35-108)
// =================================================================================================
__global__ void __launch_bounds__(256) log_rows_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = logf(x[i]);
}

// attn_only: ToMeAttention.forward alone (timm.py:75-108): x_out [N,t,C] = proj(attention), no residual; metric is required
static int tome_block_core(const QatTomeBlockArgs* a, Ctx& c, bool attn_only) {
  StageRange nvtx_(c, attn_only ? "qat.tome.attention" : "qat.tome.block");
  const int N = a->N, t = a->t, C = a->C, H = a->H;
  QAT_REQUIRE(N >= 1 && t >= 1 && H >= 1 && C == H * 64, "tome_block: need C == H*64 (C=%d H=%d)", C, H);
  QAT_REQUIRE(c.dry || (a->x && a->x_out && (attn_only ? a->metric != nullptr : a->size_out != nullptr)), "tome_block: NULL tensor");
  QAT_REQUIRE((a->unm_idx != nullptr) == (a->src_idx != nullptr) && (a->unm_idx != nullptr) == (a->dst_idx != nullptr),
              "tome_block: pass all three index outputs or none");
  QAT_TRY(check_w(a->qkv, 3 * C, C, "tome_block.qkv", c.prec));
  QAT_TRY(check_w(a->proj, C, C, "tome_block.proj", c.prec));
  QAT_REQUIRE(attn_only || (a->fc1.w && a->fc2.w && a->fc1.k == C && a->fc2.n == C && a->fc2.k == a->fc1.n),
              "tome_block: mlp weights do not chain (fc1 [%d,%d], fc2 [%d,%d])", a->fc1.n, a->fc1.k, a->fc2.n, a->fc2.k);
  const int r = (a->r > 0 && !attn_only) ? tome_clamp_r(t, a->r, a->class_token) : 0;
  const long long rows = (long long)N * t, rows2 = (long long)N * (t - r);
  const int Ch = a->fc1.n;
  const float eps = a->eps > 0.f ? a->eps : 1e-6f;
  const WView Wqkv = wview(a->qkv), Wp = wview(a->proj), W1 = wview(a->fc1), W2 = wview(a->fc2);
  const GemmKind k_qkv = pick_kind(c, rows, 3 * C, C, false, Wqkv), k_p = pick_kind(c, rows, C, C, false, Wp);
  const GemmKind k_1 = attn_only ? G_SIMT : pick_kind(c, rows2, Ch, C, false, W1), k_2 = attn_only ? G_SIMT : pick_kind(c, rows2, C, Ch, false, W2);
  Act Y, QKV, O, X1, X2, Y2, Hh;
  // x_attn, metric = attn(norm1(x), size)   (:40)
  if (a->norm1.gamma || !attn_only) {
    QAT_TRY(act_alloc_for(c, Y, rows, C, k_qkv));
    QAT_TRY(ln(c, a->x, C, nullptr, 0, nullptr, 0, 1.f, a->norm1, Y, rows, C, nullptr, 0, eps));
  } else {
    Y = act_of(a->x, C);                       // ToMeAttention.forward gets norm1(x) from its caller
    QAT_TRY(ensure_operand(c, Y, rows, C, k_qkv));
  }
  QAT_TRY(act_alloc(c, QKV, rows, 3 * C, true, false));
  QAT_TRY(gemm(c, k_qkv, Y, rows, Wqkv, Epi(), QKV));
  float* ls = nullptr;
  if (a->prop_attn && a->size) {
    ls = reinterpret_cast<float*>(c.alloc((size_t)rows * 4));
    if (!c.dry && !ls) QAT_FAIL(QAT_EWORKSPACE, "workspace too small (need > %zu bytes, have %zu)", c.off, c.cap);
    if (!c.dry) {
      log_rows_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, c.stream>>>(a->size, ls, rows);
      QAT_TRY(check_launch("log_rows_kernel"));
    }
  }
  QAT_TRY(act_alloc_for(c, O, rows, C, k_p));
  {
    AttnP p;
    p.q = QKV.f; p.k = QKV.f + C; p.v = QKV.f + 2 * C;
    p.q_sn = (long long)t * 3 * C; p.q_si = 3 * C; p.k_sn = p.q_sn; p.k_si = 3 * C; p.k_shift = 0; p.k_mod = N;
    p.o = O; p.o_rn = t; p.o_ri = 1; p.attw = nullptr; p.kbias = ls;
    p.nseq = N; p.H = H; p.Lq = t; p.Lk = t; p.scale = 0.125f; p.causal = 0;
    QAT_TRY(attention_long(c, p));
  }
  // x = x + proj(x_attn)   (:41)
  if (attn_only) {
    QAT_TRY(gemm(c, k_p, O, rows, Wp, Epi(), act_of(a->x_out, C)));
    return tome_metric(c, QKV.f + C, a->metric, rows, C, 3 * C);
  }
  QAT_TRY(act_alloc(c, X1, rows, C, true, false));
  {
    Epi e;
    e.res = a->x; e.ldr = C;
    QAT_TRY(gemm(c, k_p, O, rows, Wp, e, X1));
  }
  // metric = k.mean(heads)   (:107)
  float* metric = a->metric;
  if (!metric) {
    metric = reinterpret_cast<float*>(c.alloc((size_t)rows * 64 * 4));
    if (!c.dry && !metric) QAT_FAIL(QAT_EWORKSPACE, "workspace too small (need > %zu bytes, have %zu)", c.off, c.cap);
  }
  QAT_TRY(tome_metric(c, QKV.f + C, metric, rows, C, 3 * C));
  // merge (:45-58)
  if (r > 0) {
    const int ta = (t + 1) / 2, n_unm = ta - r;
    int* unm = a->unm_idx;
    int* src = a->src_idx;
    int* dst = a->dst_idx;
    if (!unm) {
      unm = reinterpret_cast<int*>(c.alloc((size_t)N * (n_unm + 2 * r) * 4));
      if (!c.dry && !unm) QAT_FAIL(QAT_EWORKSPACE, "workspace too small (need > %zu bytes, have %zu)", c.off, c.cap);
      src = unm + (size_t)N * n_unm;
      dst = src + (size_t)N * r;
    }
    QAT_TRY(tome_match(c, metric, N, t, 64, r, a->class_token, unm, src, dst));
    QAT_TRY(act_alloc(c, X2, rows2, C, true, false));
    if (tome_merge_wavg_supported(C)) {
      QAT_TRY(tome_merge_wavg(c, X1.f, a->size, N, t, C, r, unm, src, dst, X2.f, a->size_out, nullptr, a->group, a->T0, (a->class_token >> 1) & 1));
    } else {
      QAT_TRY(tome_merge(c, X1.f, a->size, N, t, C, r, QAT_TOME_WAVG, unm, src, dst, X2.f, a->size_out, a->group, a->T0, (a->class_token >> 1) & 1));
    }
  } else {
    X2 = X1;
    if (!c.dry) {
      if (a->size) {
        QAT_CUDA(cudaMemcpyAsync(a->size_out, a->size, (size_t)rows * 4, cudaMemcpyDeviceToDevice, c.stream));
      } else {
        ones_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, c.stream>>>(a->size_out, rows);
        QAT_TRY(check_launch("ones_kernel"));
      }
    }
  }
  // x = x + mlp(norm2(x))   (:62)
  QAT_TRY(act_alloc_for(c, Y2, rows2, C, k_1));
  QAT_TRY(ln(c, X2.f, C, nullptr, 0, nullptr, 0, 1.f, a->norm2, Y2, rows2, C, nullptr, 0, eps));
  QAT_TRY(act_alloc_for(c, Hh, rows2, Ch, k_2));
  {
    Epi e;
    e.relu = 3;
    QAT_TRY(gemm(c, k_1, Y2, rows2, W1, e, Hh));
    Epi e2;
    e2.res = X2.f; e2.ldr = C;
    QAT_TRY(gemm(c, k_2, Hh, rows2, W2, e2, act_of(a->x_out, C)));
  }
  return QAT_OK;
}

static int run_tome_block(const QatTomeBlockArgs* a, Ctx& c) { return tome_block_core(a, c, false); }
static int run_tome_attention(const QatTomeBlockArgs* a, Ctx& c) { return tome_block_core(a, c, true); }

template <typename Args, typename Fn>
static int dispatch(const Args* a, void* stream, int prec, void* ws, size_t wsb, Fn fn) {
  if (!a) QAT_FAIL(QAT_EINVAL, "args is NULL");
  Ctx c;
  QAT_TRY(ctx_init(c, stream, prec, ws, wsb, false));
  int dev = 0;
  QAT_CUDA(cudaGetDevice(&dev));
  int major = 0;
  QAT_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) QAT_FAIL(QAT_EARCH, "libqat_b200 is built for sm_100a only; device %d has compute capability %d.x", dev, major);
  return fn(a, c);
}

template <typename Args, typename Fn>
static size_t dry_bytes(const Args* a, int prec, Fn fn) {
  Ctx c;
  if (ctx_init(c, nullptr, prec, nullptr, 0, true) != QAT_OK) return 0;
  if (fn(a, c) != QAT_OK) return 0;
  return c.peak + 256;
}

}  // namespace qat

using namespace qat;

extern "C" {

int qat_gemm(const QatGemmArgs* a) { return dispatch(a, a ? a->stream : nullptr, QAT_PREC_FP32, nullptr, 0, run_gemm); }
int qat_gemm_attn_fwd(const QatGemmAttnArgs* a) { return dispatch(a, a ? a->stream : nullptr, QAT_PREC_BF16, nullptr, 0, run_gemm_attn); }
int qat_attention_fwd(const QatAttnArgs* a) { return dispatch(a, a ? a->stream : nullptr, QAT_PREC_FP32, nullptr, 0, run_attention); }
int qat_linear_fwd(const QatLinearArgs* a) { return dispatch(a, a ? a->stream : nullptr, a ? a->prec : 0, a ? a->workspace : nullptr, a ? a->workspace_bytes : 0, run_linear); }
int qat_avq_cross_attn_fwd(const QatAvqArgs* a) { return dispatch(a, a ? a->stream : nullptr, a ? a->prec : 0, a ? a->workspace : nullptr, a ? a->workspace_bytes : 0, run_avq); }
int qat_patch_select_fwd(const QatPatchArgs* a) { return dispatch(a, a ? a->stream : nullptr, a ? a->prec : 0, a ? a->workspace : nullptr, a ? a->workspace_bytes : 0, run_patch); }
int qat_gauss_reduce_fwd(const QatGaussArgs* a) { return dispatch(a, a ? a->stream : nullptr, QAT_PREC_FP32, nullptr, 0, run_gauss); }
int qat_temp_moe_fwd(const QatMoeArgs* a) { return dispatch(a, a ? a->stream : nullptr, a ? a->prec : 0, a ? a->workspace : nullptr, a ? a->workspace_bytes : 0, run_moe); }
int qat_qst_grounding_head_fwd(const QatGroundArgs* a) { return dispatch(a, a ? a->stream : nullptr, a ? a->prec : 0, a ? a->workspace : nullptr, a ? a->workspace_bytes : 0, run_ground); }
int qat_tome_match(const QatTomeMatchArgs* a) { return dispatch(a, a ? a->stream : nullptr, QAT_PREC_FP32, a ? a->workspace : nullptr, a ? a->workspace_bytes : 0, run_tome_match); }
int qat_tome_merge(const QatTomeMergeArgs* a) { return dispatch(a, a ? a->stream : nullptr, QAT_PREC_FP32, a ? a->workspace : nullptr, a ? a->workspace_bytes : 0, run_tome_merge); }
int qat_tome_chain(const QatTomeChainArgs* a) { return dispatch(a, a ? a->stream : nullptr, QAT_PREC_FP32, a ? a->workspace : nullptr, a ? a->workspace_bytes : 0, run_tome_chain); }

int qat_clip_text_fwd(const QatClipTextArgs* a) { return dispatch(a, a ? a->stream : nullptr, a ? a->prec : 0, a ? a->workspace : nullptr, a ? a->workspace_bytes : 0, run_clip_text); }

int qat_forward_fwd(const QatForwardArgs* a) { return dispatch(a, a ? a->stream : nullptr, a ? a->prec : 0, a ? a->workspace : nullptr, a ? a->workspace_bytes : 0, run_forward); }

int qat_tome_block_fwd(const QatTomeBlockArgs* a) { return dispatch(a, a ? a->stream : nullptr, a ? a->prec : 0, a ? a->workspace : nullptr, a ? a->workspace_bytes : 0, run_tome_block); }

int qat_tome_attention_fwd(const QatTomeBlockArgs* a) { return dispatch(a, a ? a->stream : nullptr, a ? a->prec : 0, a ? a->workspace : nullptr, a ? a->workspace_bytes : 0, run_tome_attention); }

int qat_gather_rows(const QatGatherArgs* a) {
  return dispatch(a, a ? a->stream : nullptr, QAT_PREC_FP32, nullptr, 0, [](const QatGatherArgs* g, Ctx& c) { return gather_rows(c, g); });
}

int qat_tome_clamp_r(int t, int r, int class_token) { return tome_clamp_r(t, r, class_token); }

int qat_tome_chain_t_end(int t0, int n_steps, const int32_t* r_list, int class_token) {
  int t = t0;
  for (int s = 0; s < n_steps; ++s) t -= tome_clamp_r(t, r_list[s], class_token);
  return t;
}

size_t qat_tome_chain_trace_len(int N, int t0, int n_steps, const int32_t* r_list, int class_token) {
  size_t tot = 0;
  int t = t0;
  for (int s = 0; s < n_steps; ++s) {
    int r = tome_clamp_r(t, r_list[s], class_token);
    if (r <= 0) continue;
    tot += (size_t)N * ((t + 1) / 2 - r + 2 * r);
    t -= r;
  }
  return tot;
}

size_t qat_workspace_bytes(int op, const void* args) {
  if (!args) return 0;
  switch (op) {
    case QAT_OP_LINEAR: return dry_bytes(reinterpret_cast<const QatLinearArgs*>(args), reinterpret_cast<const QatLinearArgs*>(args)->prec, run_linear);
    case QAT_OP_AVQ: return dry_bytes(reinterpret_cast<const QatAvqArgs*>(args), reinterpret_cast<const QatAvqArgs*>(args)->prec, run_avq);
    case QAT_OP_PATCH: return dry_bytes(reinterpret_cast<const QatPatchArgs*>(args), reinterpret_cast<const QatPatchArgs*>(args)->prec, run_patch);
    case QAT_OP_MOE: return dry_bytes(reinterpret_cast<const QatMoeArgs*>(args), reinterpret_cast<const QatMoeArgs*>(args)->prec, run_moe);
    case QAT_OP_GROUND: return dry_bytes(reinterpret_cast<const QatGroundArgs*>(args), reinterpret_cast<const QatGroundArgs*>(args)->prec, run_ground);
    case QAT_OP_CLIP_TEXT: return dry_bytes(reinterpret_cast<const QatClipTextArgs*>(args), reinterpret_cast<const QatClipTextArgs*>(args)->prec, run_clip_text);
    case QAT_OP_FORWARD: return dry_bytes(reinterpret_cast<const QatForwardArgs*>(args), reinterpret_cast<const QatForwardArgs*>(args)->prec, run_forward);
    case QAT_OP_TOME_MATCH: return 256;
    case QAT_OP_TOME_MERGE: return 256;
    case QAT_OP_TOME_CHAIN: return dry_bytes(reinterpret_cast<const QatTomeChainArgs*>(args), QAT_PREC_FP32, run_tome_chain);
    case QAT_OP_TOME_ATTENTION: return dry_bytes(reinterpret_cast<const QatTomeBlockArgs*>(args), reinterpret_cast<const QatTomeBlockArgs*>(args)->prec, run_tome_attention);
    case QAT_OP_TOME_BLOCK: return dry_bytes(reinterpret_cast<const QatTomeBlockArgs*>(args), reinterpret_cast<const QatTomeBlockArgs*>(args)->prec, run_tome_block);
    default: return 0;
  }
}

int qat_set_option(const char* name, int value) {
  if (!name) QAT_FAIL(QAT_EINVAL, "qat_set_option: name is NULL");
  for (int i = 0; i < OPT_COUNT; ++i)
    if (strcmp(name, kOptNames[i]) == 0) {
      g_opt[i].store(value, std::memory_order_relaxed);
      return QAT_OK;
    }
  QAT_FAIL(QAT_EINVAL, "qat_set_option: unknown option '%s'", name);
}
int qat_get_option(const char* name, int* value) {
  if (!name || !value) QAT_FAIL(QAT_EINVAL, "qat_get_option: NULL argument");
  for (int i = 0; i < OPT_COUNT; ++i)
    if (strcmp(name, kOptNames[i]) == 0) {
      *value = g_opt[i].load(std::memory_order_relaxed);
      return QAT_OK;
    }
  QAT_FAIL(QAT_EINVAL, "qat_get_option: unknown option '%s'", name);
}

const char* qat_last_error(void) { return g_err; }
int qat_abi_version(void) { return QAT_ABI_VERSION; }
long long qat_launch_count(int reset) {
  long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

}  // extern "C"
