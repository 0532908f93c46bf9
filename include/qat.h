/*
 * qat.h -- C ABI of libqat_b200.so: the QA-TIGER batched-eval hot path on B200 (sm_100a).
 *
 * The reference (/root/reference, 100% Python/PyTorch) has no FFI; the boundary it offers is the
 * set of nn.Module.forward() signatures in src/models/modules.py and src/models/net.py.  Each
 * entry point below replaces the body of one of those forwards (file:line cited per function) and
 * is called from the Python mirror in qa-tiger_b200/models/ through ctypes (INTEGRATION.md shows
 * the stub a maintainer would add to the reference's own modules.py).
 *
 * Conventions
 *  - plain C structs of DEVICE pointers + int sizes + a cudaStream_t passed as void*.
 *  - the CALLER owns every buffer: inputs, outputs and the workspace (sized with
 *    qat_workspace_bytes).  The library never allocates or frees device memory, never
 *    synchronises the device or the stream, never changes the current device.
 *  - every tensor is dense, row-major, in the reference's [B,T,...] order, fp32 unless noted.
 *  - return 0 on success; otherwise a QAT_E* code, with a thread-local message in
 *    qat_last_error().  Unsupported shapes / arch (!= sm_100) are errors, never fallbacks.
 *  - re-entrant: no global mutable state on the data path (nn.DataParallel drives one thread per GPU); the only process-wide
 *    state is the A/B option table (qat_set_option), atomics that each call snapshots on entry.
 */
#ifndef QAT_H_
#define QAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QAT_ABI_VERSION 4

enum {
  QAT_OK = 0,
  QAT_EINVAL = 1,     /* bad argument / unsupported shape */
  QAT_EWORKSPACE = 2, /* workspace too small */
  QAT_ECUDA = 3,      /* CUDA runtime / driver error (message has the CUDA string) */
  QAT_EARCH = 4       /* device is not sm_100 */
};

/* Arithmetic mode of the dense contractions (error budget in DESIGN.md). */
enum {
  QAT_PREC_FP32 = 0, /* fp32 FFMA on CUDA cores: the 1e-5 parity gate */
  QAT_PREC_BF16 = 1  /* "bf16 mode" (16-bit tensor-core operands): tcgen05, fp32 accumulate in TMEM;
                        bf16 hi+lo pairs (3 MMAs per product) on the serial chain that feeds the router,
                        tf32 for patch_proj (straight from the fp32 input) and single fp16 planes
                        (round-to-nearest) inside PatchSelecter; router / Gaussian / rank path stays
                        fp32: the 1e-3 gate */
};

/* Element type of the streamed tensor of qat_gauss_reduce_fwd. */
enum { QAT_DT_F32 = 0, QAT_DT_BF16 = 1 };

/* One nn.Linear: y = x W^T + b.  `w` is the reference parameter itself ([n,k] row-major fp32);
 * `w_hi`/`w_lo` are its bf16 split (w ~= hi + lo) and `w_f16` its fp16 rounding; they are read only
 * with QAT_PREC_BF16 (w_f16 only by the PatchSelecter weights slf_in / slf_out / crs_kv; without it
 * that block runs tf32 on fp32 activations). */
typedef struct {
  const float* w;
  const void* w_hi;
  const void* w_lo;
  const void* w_f16;
  const float* bias; /* [n] or NULL */
  int32_t n, k;
} QatLinearW;

typedef struct {
  const float* gamma;
  const float* beta;
} QatLayerNormW;

/* ---- P1: Projection.forward, modules.py:35-43 (net.py:120-124) ------------------------------ */
typedef struct {
  int32_t rows;     /* product of the leading dims */
  int32_t prec;
  void* stream;
  const float* x;   /* [rows, w.k] */
  QatLinearW w;
  float* y;         /* [rows, w.n] */
  int32_t relu;     /* 0 none, 1 ReLU, 2 QuickGELU x*sigmoid(1.702x) */
  void* workspace;
  size_t workspace_bytes;
} QatLinearArgs;
int qat_linear_fwd(const QatLinearArgs* a);

/* ---- one dense contraction on a chosen engine (building block of every entry point below; exported for
 *      parity tests and roofline timing of the GEMM kernels in isolation) ------------------------- */
enum { QAT_GEMM_SIMT = 0,   /* fp32 FFMA, reads a / w.w */
       QAT_GEMM_TF32 = 1,   /* tcgen05 kind::tf32, reads a / w.w */
       QAT_GEMM_BF16X3 = 2, /* tcgen05 kind::f16 x3, reads (a_hi, a_lo) / (w.w_hi, w.w_lo) */
       QAT_GEMM_F16 = 3     /* tcgen05 kind::f16, reads a_f16 / w.w_f16 */ };
typedef struct {
  int32_t kind;
  int32_t M;
  void* stream;
  const float* a;      /* [M, lda] fp32 (SIMT, TF32) */
  const void* a_hi;    /* [M, lda] bf16 planes (BF16X3) */
  const void* a_lo;
  const void* a_f16;   /* [M, lda] fp16 plane (F16) */
  int32_t lda;
  QatLinearW w;        /* [n, k] */
  int32_t relu;
  const float* res;    /* optional residual [M, ldr], added after the activation */
  const void* res_f16; /* the same as an fp16 plane (tensor-core kinds only; `res` wins if both are set) */
  int32_t ldr;
  float* c;            /* any of the outputs may be NULL */
  void* c_hi;
  void* c_lo;
  void* c_f16;         /* fp16 plane; when it is the only output the tile leaves through a TMA store */
  int32_t ldc;
  /* reduce epilogue (QAT_GEMM_BF16X3 only; M4 of SURVEY 8a folded into the expert first-layer GEMM, modules.py:473-534): rows of
   * A are (sample, t) pairs, coef_T rows per sample; instead of act(A W^T + b) the call writes, per sample s and column n,
   *   c[s, n] = sum_t coef[(s*coef_T + t) % coef_mod, n / coef_Ch] * act(...)[s*coef_T + t, n]       (c is [M / coef_T, ldc])
   * 32 <= coef_T <= 128, M % coef_T == 0, coef_Ch % 32 == 0, w.n % 256 == 0; coef = NULL: the plain epilogue above */
  const float* coef;   /* [coef_mod, coef_E] */
  int32_t coef_E, coef_T, coef_Ch;
  int64_t coef_mod;
} QatGemmArgs;
int qat_gemm(const QatGemmArgs* a);

/* ---- PatchSelecter's fused projection + attention kernels (modules.py:688 and :692-698), exported for parity tests and roofline
 *      timing.  fp16 planes in and out, fp32 bias, head dim 64, P <= 14 tokens per frame (whole frames per 128-row tile).
 *      cross = 0:  O[nframes*P, ldo] = per-frame softmax((X Wq^T + bq)(X Wk^T + bk)^T * scale)(X Wv^T + bv),  W = in_proj [3D, K]
 *      cross = 1:  O[nframes*2, ldo] = per-frame softmax(Q (X Wk^T + bk)^T * scale)(X Wv^T + bv) for the two projected query rows
 *                  Q [nframes, 2, D] of every frame,  W = in_proj[D:] [2D, K]
 *      K = 512 and an even head count run the A-resident kernels (A operand in tensor memory), other shapes the round-1 form ---- */
typedef struct {
  int32_t cross, nframes, P, H, K;
  float scale;
  void* stream;
  const void* x_f16;       /* [nframes*P, ldx] */
  int32_t ldx;
  const void* w_f16;       /* [(3 - cross) * H*64, K] */
  const float* bias;       /* [(3 - cross) * H*64] */
  const void* q_f16;       /* cross only */
  void* o_f16;
  int32_t ldo;
} QatGemmAttnArgs;
int qat_gemm_attn_fwd(const QatGemmAttnArgs* a);

/* ---- softmax(q k^T * scale) v on head-dim-64 heads: the attention cores every block above is built from (torch MHA
 *      slow path semantics, SURVEY 8a-A3: no masks except `causal`); exported for parity tests and roofline timing ---- */
enum { QAT_ATTN_F32 = 0,       /* fp32 q/k/v, generic strided views; the library picks the core from (Lq, Lk) */
       QAT_ATTN_F16_SMALL = 1  /* fp16 planes, Lq, Lk <= 16, H <= 8: one CTA per sequence on mma.sync (PatchSelecter) */ };
typedef struct {
  int32_t kind;
  int32_t nseq, H, Lq, Lk;
  float scale;
  int32_t causal;          /* F32 only, Lk > 32 or Lq >= 32 */
  void* stream;
  const void* q;           /* F32: element (n, i, h, d) at n*q_sn + i*q_si + h*64 + d;  F16_SMALL: rows [nseq*Lq, q_si] */
  const void* k;           /* F32: kv sequence of query sequence n is (n + k_shift) % k_mod; F16_SMALL: rows [nseq*Lk, k_si] = K | V packed */
  const void* v;           /* F32 only */
  int64_t q_sn, q_si, k_sn, k_si;
  int32_t k_shift, k_mod;
  float* o;                /* F32: row n*o_rn + i*o_ri of a [rows, ldo] matrix, col h*64 + d */
  void* o_f16;             /* F16_SMALL: rows [nseq*Lq, ldo] */
  int64_t o_rn, o_ri;
  int32_t ldo;
  float* attw;             /* F32 only: optional [nseq, Lq, Lk] head-averaged weights, zero-initialised by the caller */
  const float* kbias;      /* F32 only, optional [k_mod, Lk]: added to every score of key j AFTER the scaling (ToMe's proportional
                              attention, log(size): tome/patch/timm.py:94-95).  Sequences with Lk > 128 or a bias run on the
                              long-sequence (online-softmax) core: no causal mask, no attw */
} QatAttnArgs;
int qat_attention_fwd(const QatAttnArgs* a);

/* ---- A1-A3: AVQCrossAttn.forward, modules.py:239-265 (sub_forward :194-237) ----------------- */
typedef struct {
  int32_t B, T, L, D, H; /* T frames, L words; D = H*64 */
  int32_t prec;
  void* stream;
  const float* audio;    /* src_q [B,T,D] */
  const float* video;    /* src_v [B,T,D] */
  const float* words;    /* query [B,L,D] */
  /* rows: q_qst | q_slf k_slf v_slf | q_crs k_crs v_crs  = cat(qst.in_proj[0:D], slf.in_proj, crs.in_proj) */
  QatLinearW in_src;     /* [7D, D] */
  QatLinearW in_words;   /* [2D, D] = qst.in_proj[D:3D] */
  QatLinearW out_cat;    /* [D, 3D] = [Wo_qst | Wo_slf | Wo_crs], bias = sum of the three out_proj biases */
  QatLinearW ffn1, ffn2; /* linear1, linear2 */
  QatLayerNormW norm1, norm2;
  float* out_audio;      /* [B,T,D] */
  float* out_video;      /* [B,T,D] */
  float* attw_audio;     /* optional [B,T,L] head-averaged qst_attn weights (visualize=True), else NULL; must be zeroed */
  float* attw_video;
  void* workspace;
  size_t workspace_bytes;
} QatAvqArgs;
int qat_avq_cross_attn_fwd(const QatAvqArgs* a);

/* ---- P1(patch) + S1: patch_proj (net.py:124) + PatchSelecter.forward, modules.py:656-716 ---- */
typedef struct {
  int32_t B, T, P, D, H, Cin; /* Cin = raw patch dim (1024) when patch_proj is fused, else ignored */
  int32_t prec;
  void* stream;
  const float* patch;    /* [B,T,P,Cin] if has_proj else [B,T,P,D] */
  const void* patch_f16; /* the same tensor as fp16 (half-size feature store / H2D); when set, `patch` is ignored.
                            bf16 mode feeds it straight to the tensor cores (patch_proj then needs proj.w_f16), fp32 mode widens it */
  const float* audio;    /* [B,T,D] */
  const float* video;    /* [B,T,D] */
  int32_t has_proj;
  QatLinearW proj;       /* patch_proj.proj */
  QatLinearW slf_in;     /* [3D,D] */
  QatLinearW slf_out;    /* [D,D] */
  QatLinearW crs_q;      /* crs.in_proj[0:D] */
  QatLinearW crs_kv;     /* crs.in_proj[D:3D] */
  QatLinearW crs_out;
  QatLinearW mlp0, mlp2;
  QatLayerNormW anorm, vnorm;
  float* out_a;          /* [B,T,D] anorm(a) */
  float* out_v;          /* [B,T,D] vnorm(v) */
  int32_t chunk_frames;  /* frames (b,t pairs) per pass, 0 = library default (16 384: B = 256, T = 60 is one pass) */
  void* workspace;
  size_t workspace_bytes;
} QatPatchArgs;
int qat_patch_select_fwd(const QatPatchArgs* a);

/* ---- M2-M4 standalone: router softmax + top-K + Gaussian masks + weighted temporal reduction.
 *      TempMoE.forward :565-578, generate_gaussian :420-471, get_output :473-534 ---------------- */
typedef struct {
  int32_t B, T, E, K, C;
  int32_t dtype;           /* QAT_DT_* of `experts` */
  float sigma;             /* 9 */
  void* stream;
  const void* experts;     /* experts_logits: element (t,b,e,c) at t*stride_t + b*stride_b + e*C + c */
  int64_t stride_t, stride_b; /* reference layout [T,B,E,C]: stride_t = B*E*C, stride_b = E*C */
  const float* router_logits; /* [B,E]  pre-softmax */
  const float* gauss_raw;     /* [B,2E] pre tanh/sigmoid, interleaved (centre, width) per expert */
  float* out;              /* [B,C] (pre-LayerNorm) */
  int64_t* topk_inds;      /* optional [B,K] */
  float* topk_probs;       /* optional [B,K] renormalised */
  float* gauss_weight;     /* optional [B,K,T] peak-normalised */
  int32_t t_split;         /* CTAs per sample (column slices of one cluster) 0 = auto */
} QatGaussArgs;
int qat_gauss_reduce_fwd(const QatGaussArgs* a);

/* ---- M1-M4 (+LN): TempMoE.forward, modules.py:536-612 -------------------------------------- */
typedef struct {
  int32_t B, T, D, H, E, K;
  int32_t prec;
  float sigma;
  void* stream;
  const float* qst;        /* [B,D] */
  const float* data;       /* [B,T,D] */
  const float* sub_a;      /* optional sub_data[0] [B,T,D] */
  const float* sub_v;      /* optional sub_data[1] [B,T,D] */
  QatLinearW attn_q;       /* qst_attn.in_proj[0:D] */
  QatLinearW attn_kv;      /* qst_attn.in_proj[D:3D] (fp32; only the K rows [0, D) are read, see attn_rg) */
  QatLinearW attn_rg;      /* [3E, H*D] fp32 + bias [3E]: the value projection, the out-projection, router.0 and gauss_pred.0 of
                              modules.py:562-578 composed on the host (one query row per sample: the projections commute with the
                              attention, csrc/moe_front.cuh).  With Wrg = cat(router.0.weight, gauss_pred.0.weight) [3E, D],
                              Wo / bo = qst_attn.out_proj, Wv / bv = qst_attn.in_proj[2D:3D], head h = rows / columns [64h, 64h+64):
                                attn_rg.w[:, h*D:(h+1)*D] = (Wrg Wo)[:, 64h:64h+64] Wv[64h:64h+64, :]
                                attn_rg.bias             = Wrg (Wo bv + bo) + cat(router.0.bias, gauss_pred.0.bias)
                              (compute in fp64, round once to fp32; qa-tiger_b200/models/modules.py TempMoE._prepare) */
  QatLinearW w1_cat;       /* [E*D/2, D] = cat_e experts.e.0 */
  QatLinearW w2_ext;       /* [D, E*D/2 + 64]: cols e*D/2+c = experts.e.2.weight[:,c]; col E*D/2+e = experts.e.2.bias; 0 after */
  QatLayerNormW norm_a;    /* `norm` (single branch) or `anorm` */
  QatLayerNormW norm_v;    /* `vnorm` (two branches) */
  float* out_a;            /* [B,D]  (norm(main) or anorm(a)) */
  float* out_v;            /* [B,D]  or NULL when sub_* are NULL */
  int64_t* topk_inds;      /* optional debug/parity outputs, as in QatGaussArgs */
  float* topk_probs;
  float* gauss_weight;
  void* workspace;
  size_t workspace_bytes;
} QatMoeArgs;
int qat_temp_moe_fwd(const QatMoeArgs* a);

/* ---- G1 (x1) : QstGrounding.forward, modules.py:304-347; optionally followed by H1: head(relu(.)) net.py:141-142.
 *      Two forms: the length-2 data list of both call sites (net.py:136, 138) as d0 / d1, or any concatenated sequence as
 *      `data` [B,S,D] with S >= 1 (modules.py:318-326 accept any list of [B,S_i,D] tensors) ---------------------------------- */
typedef struct {
  int32_t B, D, H;
  int32_t prec;
  void* stream;
  const float* qst;        /* [B,D] */
  const float* d0;         /* [B,D] first element of the data list */
  const float* d1;         /* [B,D] second */
  const float* data;       /* general form: [B,S,D] contiguous (d0 / d1 ignored); NULL selects the d0 / d1 form */
  int32_t S;
  QatLinearW attn_q, attn_kv, attn_out, mlp0, mlp2;
  QatLayerNormW norm;
  float* out;              /* [B,D] */
  int32_t has_head;
  QatLinearW head;         /* [n_ans, D] */
  float* logits;           /* [B,n_ans] when has_head */
  void* workspace;
  size_t workspace_bytes;
} QatGroundArgs;
int qat_qst_grounding_head_fwd(const QatGroundArgs* a);

/* ---- T1: bipartite_soft_matching, tome/merge.py:18-97 (index form) -------------------------- */
typedef struct {
  int32_t N, t, C;          /* metric [N,t,C] */
  int32_t r;                /* requested; clamped as merge.py:43-44 */
  int32_t class_token;      /* protected-token flags: bit 0 = token 0 is a class token (merge.py:54-55, 66-68), bit 1 = token 1 is a
                               distillation token (merge.py:56-57); the same flag word wherever `class_token` appears below */
  void* stream;
  const float* metric;
  int32_t* unm_idx;         /* [N, ceil(t/2)-r'] */
  int32_t* src_idx;         /* [N, r'] */
  int32_t* dst_idx;         /* [N, r'] */
  void* workspace;
  size_t workspace_bytes;
} QatTomeMatchArgs;
int qat_tome_match(const QatTomeMatchArgs* a);
int qat_tome_clamp_r(int t, int r, int class_token);

/* ---- T2/T3: merge_wavg :210-225 and merge_source :228-241 (as group ids) --------------------- */
enum { QAT_TOME_MODE_WAVG = 0, /* merge_wavg: x*size summed / summed size */
       QAT_TOME_MODE_SUM = 1, QAT_TOME_MODE_MEAN = 2, QAT_TOME_MODE_AMAX = 3 /* the bare merge(x, mode) closure */ };
typedef struct {
  int32_t N, t, C, r;       /* r = clamped */
  int32_t mode;             /* QAT_TOME_MODE_* */
  int32_t distill_token;    /* 1: output rows ordered [unm_0, dst_0, unm_1.., dst_1..] (merge.py:77-78) instead of [unm | dst] */
  void* stream;
  const float* x;           /* [N,t,C] */
  const float* size;        /* [N,t] or NULL (= ones); WAVG only */
  const int32_t* unm_idx;
  const int32_t* src_idx;
  const int32_t* dst_idx;
  float* x_out;             /* [N,t-r,C] */
  float* size_out;          /* [N,t-r]; WAVG only */
  int32_t* group;           /* optional in/out [N,T0]: current position of every original token */
  int32_t T0;
  void* workspace;
  size_t workspace_bytes;
} QatTomeMergeArgs;
int qat_tome_merge(const QatTomeMergeArgs* a);

/* ---- T0: the schedule driver tome/patch/timm.py:35-65 with the stand-in metric of SURVEY 8d --- */
typedef struct {
  int32_t N, t0, C;
  int32_t n_steps;
  const int32_t* r_list;    /* HOST pointer, n_steps entries */
  int32_t class_token;
  void* stream;
  const float* x;           /* [N,t0,C] */
  float* x_out;             /* [N,t_end,C] */
  float* size_out;          /* [N,t_end] */
  int32_t* group;           /* optional [N,t0] */
  int32_t* idx_trace;       /* optional: per step unm|src|dst concatenated, see qat_tome_chain_trace_len */
  void* workspace;
  size_t workspace_bytes;
} QatTomeChainArgs;
int qat_tome_chain(const QatTomeChainArgs* a);
int qat_tome_chain_t_end(int t0, int n_steps, const int32_t* r_list, int class_token);
size_t qat_tome_chain_trace_len(int N, int t0, int n_steps, const int32_t* r_list, int class_token);

/* ---- the whole hot path in one call: QA_TIGER.forward net.py:120-142 on the text tower's outputs -----------------
 * The embedded per-block structs carry WEIGHTS and hyper-parameters only (D, H, E, K, sigma, Cin, has_proj, chunk_frames,
 * head); their tensor pointers, B/T/L/P, prec, stream and workspace fields are ignored.  at.topk_inds / topk_probs /
 * gauss_weight (and vt.*) are honoured as optional outputs.  Intermediates live in the workspace and move between the
 * blocks with the storage planes their consumers read, so no conversion pass runs at a module boundary. */
typedef struct {
  int32_t B, T, L, P;
  int32_t prec;
  void* stream;
  const float* quest;     /* [B, quest_proj.k]   CLIP question feature */
  const float* words;     /* [B, L, words_proj.k] */
  const float* audio;     /* [B, T, audio_proj.k] */
  const float* video;     /* [B, T, video_proj.k] */
  const float* patch;     /* [B, T, P, patch_sel.Cin]; or NULL with patch_sel.patch_f16 set (the one tensor pointer of the
                             embedded structs that IS honoured) */
  QatLinearW audio_proj, video_proj, words_proj, quest_proj;   /* net.py:120-123; patch_proj is patch_sel.proj */
  QatAvqArgs avq;
  QatPatchArgs patch_sel;
  QatMoeArgs at, vt;
  QatGroundArgs ground;   /* both QstGrounding passes share these weights; `head` is applied after the second */
  QatLinearW q_cat;       /* optional [3D, D] = cat(at.attn_q, vt.attn_q, ground.attn_q) (+ biases): the query projections of
                             the blocks that attend from the question as one contraction; w = NULL -> each block does its own */
  float* logits;          /* [B, ground.head.n] */
  void* workspace;
  size_t workspace_bytes;
} QatForwardArgs;
int qat_forward_fwd(const QatForwardArgs* a);

/* ---- N1: frozen CLIP text tower, encoders.py:30-37 + base/clip_base.py:186-254 (ResidualAttentionBlock
 *      x + attn(ln_1 x), x + c_proj(QuickGELU(c_fc(ln_2 x))), causal mask), ln_final, EOT-row @ text_projection ---- */
typedef struct {
  QatLayerNormW ln_1, ln_2;
  QatLinearW in_proj;       /* attn.in_proj_{weight,bias} [3W, W] */
  QatLinearW out_proj;      /* attn.out_proj [W, W] */
  QatLinearW c_fc;          /* mlp.c_fc [4W, W] */
  QatLinearW c_proj;        /* mlp.c_proj [W, 4W] */
} QatClipLayerW;
typedef struct {
  int32_t B, L, W, H, E;    /* L tokens per question (77), width W = H*64, embed dim E */
  int32_t n_layers, vocab;
  int32_t prec;             /* QAT_PREC_FP32: fp32 FFMA; QAT_PREC_BF16: bf16 hi+lo x3 (fp32-faithful: the tower feeds the router) */
  void* stream;
  const int32_t* tokens;    /* [B, L] CLIP token ids; the EOT row is the FIRST argmax of the ids (clip_base.py:396) */
  const float* token_embedding;      /* [vocab, W] */
  const float* positional_embedding; /* [L, W] */
  const QatClipLayerW* layers;       /* HOST pointer, n_layers entries */
  QatLayerNormW ln_final;
  QatLinearW text_projection;        /* [E, W] = text_projection^T, no bias */
  float* quest;             /* [B, E] */
  float* words;             /* [B, L, W] */
  void* workspace;
  size_t workspace_bytes;
} QatClipTextArgs;
int qat_clip_text_fwd(const QatClipTextArgs* a);

/* ---- N4: one ToMe-patched ViT block = ToMeBlock.forward, tome/patch/timm.py:35-65, with ToMeAttention.forward
 *      :68-108 inside:  x += proj(softmax(q k^T * hd^-0.5 + log(size)) v)   (proportional attention :94-95; q, k, v = qkv(norm1 x))
 *                       metric = k.mean(heads) :107  ->  bipartite_soft_matching(metric, r) -> merge_wavg (+ source group ids)
 *                       x += fc2(GELU(fc1(norm2 x))).
 *      So T1 runs on the REAL metric form, not the stand-in of qat_tome_chain.  Parameter names are timm's (norm1, attn.qkv,
 *      attn.proj, norm2, mlp.fc1, mlp.fc2). ---- */
typedef struct {
  int32_t N, t, C, H;        /* frames, tokens per frame, width C = H*64, heads */
  int32_t r;                 /* tokens to merge after the attention (clamped as merge.py:43-44); 0 = none */
  int32_t class_token;       /* token 0 is the class token (protected) */
  int32_t prop_attn;         /* 1: add log(size) to the scores (needs `size`; size = NULL means all ones: no bias) */
  int32_t prec;
  float eps;                 /* LayerNorm eps (timm ViT: 1e-6) */
  void* stream;
  const float* x;            /* [N,t,C] */
  const float* size;         /* [N,t] token sizes, or NULL */
  QatLayerNormW norm1, norm2;
  QatLinearW qkv;            /* [3C, C] */
  QatLinearW proj;           /* [C, C] */
  QatLinearW fc1;            /* [4C, C] (any hidden width) */
  QatLinearW fc2;            /* [C, 4C] */
  float* x_out;              /* [N,t-r',C], r' = the clamped r (qat_tome_clamp_r) */
  float* size_out;           /* [N,t-r'] */
  float* metric;             /* optional [N,t,64] */
  int32_t* unm_idx;          /* optional merge indices of this block, as qat_tome_match (all three or none) */
  int32_t* src_idx;
  int32_t* dst_idx;
  int32_t* group;            /* optional in/out [N,T0]: merged-token position of every original token (merge_source :228-241) */
  int32_t T0;
  void* workspace;
  size_t workspace_bytes;
} QatTomeBlockArgs;
int qat_tome_block_fwd(const QatTomeBlockArgs* a);
/* ToMeAttention.forward alone (timm.py:75-108) on the same struct: x = norm1 output is NOT applied here -- `x` is the attention's
 * input as the caller passes it (the reference calls attn(norm1(x), size)); x_out [N,t,C] = proj(attention) without residual,
 * metric [N,t,64] = k.mean(heads) (required).  norm1 is applied iff norm1.gamma != NULL; r, norm2, fc1, fc2, size_out are ignored. */
int qat_tome_attention_fwd(const QatTomeBlockArgs* a);

/* ---- N3: batch assembly from a device-resident per-video feature cache.  dataset.py:132-159 loads the three
 *      feature files of a question's video per QUESTION and trainutils.py:223-236 copies every batch to the device; with the
 *      cache (qa_tiger_b200/datapipe.py) a video crosses the host link once and a batch is gathered in HBM:
 *      dst[k][i, :] = src[k][slots[i], :] for k < n_tensors, one launch.  Rows are dense; row_bytes[k] % 16 == 0. ---- */
#define QAT_GATHER_MAX 4
typedef struct {
  int32_t n;                    /* rows to gather (questions of the batch), <= 65535 */
  int32_t n_tensors;            /* 1..QAT_GATHER_MAX (audio, video, patch) */
  int32_t n_slots;              /* cache capacity in rows; a slot id outside [0, n_slots) fills the row with NaN bit patterns */
  void* stream;
  const int32_t* slots;         /* [n] device: cache row of each question's video */
  const void* src[QAT_GATHER_MAX];   /* cache tensors: row s of tensor k starts at src[k] + s * src_stride[k] */
  void* dst[QAT_GATHER_MAX];         /* batch tensors [n, row_bytes[k]], dense */
  size_t row_bytes[QAT_GATHER_MAX];
  size_t src_stride[QAT_GATHER_MAX]; /* bytes between cache rows (0 = dense = row_bytes[k]); the cache keeps a video's audio |
                                        video | patch in ONE row so that a video is one host->device copy */
} QatGatherArgs;
int qat_gather_rows(const QatGatherArgs* a);

/* ---- bookkeeping ---------------------------------------------------------------------------- */
enum { QAT_OP_LINEAR = 0, QAT_OP_AVQ = 1, QAT_OP_PATCH = 2, QAT_OP_MOE = 3, QAT_OP_GROUND = 4,
       QAT_OP_TOME_MATCH = 5, QAT_OP_TOME_MERGE = 6, QAT_OP_TOME_CHAIN = 7, QAT_OP_CLIP_TEXT = 8, QAT_OP_FORWARD = 9,
       QAT_OP_TOME_BLOCK = 10, QAT_OP_TOME_ATTENTION = 11 };
/* Pass the same args struct the op will be called with (pointers may be NULL). */
size_t qat_workspace_bytes(int op, const void* args);
const char* qat_last_error(void);
int qat_abi_version(void);
/* number of kernel launches issued by this thread since the last reset (bench.py's gpu_launches) */
long long qat_launch_count(int reset);
/* A/B switches (profiling aids; defaults are the measured-faster paths): "gemm_cg", "gemm_bn", "attn_fp32_in", "patch_chunk",
 * "patch_fp32_planes", "patch_no_fused_attn", "patch_no_fused_cross", "tome_match_rows", "moe_no_fused_reduce", "gemm_attn_no_ts", "moe_w2_simt", "gemm_attn_ts_flags" (the last one only acts in a -DQAT_PROFILE_BUILD library).  Process-wide;
 * initialised once from the environment (QAT_<UPPER-CASE NAME>) when the library is loaded, never read from the environment
 * again; every entry point snapshots them on entry, so sizing and execution of one call agree.  Set them between calls. */
int qat_set_option(const char* name, int value);
int qat_get_option(const char* name, int* value);

#ifdef __cplusplus
}
#endif
#endif /* QAT_H_ */
