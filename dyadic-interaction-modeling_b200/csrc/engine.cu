// Model-level entry points: a handle owns the weight registry (reference state_dict keys -> borrowed device pointers),
// re-packed weights, and the launch sequences of
//   * VQAutoEncoder.encode / decode           (/root/reference/code/models/stage1_BIWI.py:22-37, :307-317, :376-393)
//   * SLMFT.forward_encoder + context concat   (/root/reference/code/seq2seq_pretrain.py:431-446)
//   * decoder_joint.generate                   (seq2seq_pretrain.py:450; x-transformers 1.30.16)
// Everything is enqueued on the caller's stream; no host synchronisation inside.
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "attention.cuh"
#include "common.cuh"
#include "decode_mk.cuh"
#include "gemm_f32.cuh"
#include "gemm_tc.cuh"
#include "rowops.cuh"
#include "vq.cuh"

using namespace dimb;

namespace {

bool g_mk_trace_on = false;
int g_decode_impl = 0;                                    // 0: persistent kernel when supported; 1: per-kernel CUDA-graph path
std::vector<unsigned long long> g_mk_trace_host;     // last collected trace, ns per phase
std::vector<int> g_mk_trace_types;

struct TensorRef {
  const void* p = nullptr;
  int dtype = DIM_DTYPE_F32;
  std::vector<int64_t> shape;
};

// Tensor-core dispatch state of one built model: planes == 0 -> fp32 FFMA kernels; 1 -> plain bf16 operands;
// 3 -> exact 3-way bf16 split of fp32 operands (6 products, fp32-grade results).  wmap: fp32 weight -> its plane matrix.
struct TcCtx {
  int planes = 0;
  std::unordered_map<const float*, const __nv_bfloat16*> wmap;
};

struct VqLayer {
  const float *ln1_g, *ln1_b, *wqkv, *wo, *bo, *ln2_g, *ln2_b, *w1, *b1, *w2, *b2;
};

struct VqModel {
  dim_vq_config cfg{};
  int precision = DIM_PREC_FP32;
  TcCtx tc;
  const float *map_w, *map_b, *conv_b, *emb_w, *emb_b, *pe, *post_w, *post_b;
  float* conv_wr = nullptr;
  std::vector<VqLayer> enc;
  const float *pre_w, *pre_b, *dconv_b, *demb_w, *demb_b, *dpe, *rev_w;
  float* dconv_wr = nullptr;
  std::vector<VqLayer> dec;
  const float* codebook;
  bool has_enc = true, has_dec = true;
  int fqn = 1, out_dim = 0;            // codes per frame; decoder output channels
};

struct XtAttn {
  const float *norm_g = nullptr, *norm_b = nullptr, *wq = nullptr, *wo = nullptr;
  float* wqkv = nullptr;   // owned: cat(to_q, to_k, to_v) [3*inner, dim]   (self attention)
  float* wkv = nullptr;    // owned: cat(to_k, to_v)       [2*inner, dim]   (cross attention)
};
struct XtFF {
  const float *norm_g = nullptr, *norm_b = nullptr, *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;
};
struct XtEncoder {
  int dim_in = 0;
  const float *proj_w = nullptr, *proj_b = nullptr, *pos_emb = nullptr, *final_g = nullptr, *final_b = nullptr;
  std::vector<XtAttn> attn;
  std::vector<XtFF> ff;
};
struct S2SModel {
  dim_s2s_config cfg{};
  int precision = DIM_PREC_FP32;
  TcCtx tc;
  XtEncoder enc_s, enc_joint, enc_l;           // enc_l: SLM pre-training only (optional); enc_joint absent in single-encoder models
  bool has_enc_l = false, has_enc_joint = false;
  const float *patch_s = nullptr, *patch_dec_s = nullptr, *norm_s_g = nullptr, *norm_s_b = nullptr;
  const float *norm_l_g = nullptr, *norm_l_b = nullptr, *norm_j_g = nullptr, *norm_j_b = nullptr;     // SLM: norm_l, norm
  const float* dec_pos_emb = nullptr;          // decoder_joint.net.pos_emb (SLM keeps use_abs_pos_emb=True; SLMFT has none)
  const float* token_emb = nullptr;
  std::vector<XtAttn> self_attn, cross_attn;
  std::vector<XtFF> ff;
  const float *final_g = nullptr, *final_b = nullptr, *logits_w = nullptr, *logits_b = nullptr;
  // CUDA graph of ONE decode step (the step index lives in device memory, so every step replays the same graph)
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;      // one decode step
    cudaGraphExec_t exec_u = nullptr;    // kUnroll decode steps back to back (fewer graph launches per generate)
    std::vector<uintptr_t> key;
    cudaStream_t stream = nullptr;       // side stream the graph runs on (capture is not allowed on the legacy stream)
    cudaEvent_t fork = nullptr, join = nullptr;
  };
  mutable std::vector<StepGraph> graphs;
  mutable cudaEvent_t fork_ev = nullptr;
};

}  // namespace

struct dim_handle_s {
  int device = 0;
  std::unordered_map<std::string, TensorRef> tensors;
  std::vector<std::unique_ptr<VqModel>> vq;
  std::vector<std::unique_ptr<S2SModel>> s2s;
  std::vector<void*> owned;
};

namespace {

// Look a float tensor up by name and check its shape.  required=false returns nullptr quietly when absent.
int lookup(dim_handle_s* h, const std::string& name, std::initializer_list<int64_t> shape, bool required,
           const float** out) {
  *out = nullptr;
  auto it = h->tensors.find(name);
  if (it == h->tensors.end()) {
    if (!required) return DIM_OK;
    return fail(DIM_EMISSING, "weight not registered: " + name);
  }
  const TensorRef& t = it->second;
  if (t.dtype != DIM_DTYPE_F32) return fail(DIM_EINVAL, "weight is not fp32: " + name);
  std::vector<int64_t> want(shape);
  if (t.shape != want) {
    std::string got, exp;
    for (auto v : t.shape) got += std::to_string(v) + ",";
    for (auto v : want) exp += std::to_string(v) + ",";
    return fail(DIM_EINVAL, "shape mismatch for " + name + ": got (" + got + ") expected (" + exp + ")");
  }
  *out = static_cast<const float*>(t.p);
  return DIM_OK;
}

#define LOOKUP(dst, name, req, ...)                                             \
  do {                                                                          \
    if (int _e = lookup(h, (name), {__VA_ARGS__}, (req), &(dst))) return _e;     \
  } while (0)

int owned_alloc(dim_handle_s* h, size_t bytes, float** out) {
  void* p = nullptr;
  DIM_CHECK_CUDA(cudaMalloc(&p, bytes));
  h->owned.push_back(p);
  *out = static_cast<float*>(p);
  return DIM_OK;
}

int planes_of(int precision) {
  return precision == DIM_PREC_BF16 ? 1 : (precision == DIM_PREC_FP32_TC ? 3 : 0);
}

// Split an fp32 weight [N,K] into its bf16 plane matrix (owned by the handle) and remember it.
int tc_add_weight(dim_handle_s* h, TcCtx& tc, const float* W, int N, int K) {
  if (tc.planes == 0 || W == nullptr || tc.wmap.count(W)) return DIM_OK;
  const int kp = tc_round_k(K);
  float* buf = nullptr;
  if (int e = owned_alloc(h, (size_t)N * tc.planes * kp * sizeof(__nv_bfloat16), &buf)) return e;
  GemmArgs a;
  a.A = W; a.lda = K; a.M = N; a.K = K;
  if (int e = launch_split_planes(a, reinterpret_cast<__nv_bfloat16*>(buf), kp, tc.planes, nullptr)) return e;
  tc.wmap[W] = reinterpret_cast<const __nv_bfloat16*>(buf);
  return DIM_OK;
}

// One Linear: tensor cores (operand split + tcgen05 GEMM) whenever the model is built for them and the problem is beyond the
// GEMV kernel's range (M > 8): rows past M inside the 128-row MMA tile are zero-filled by TMA and never stored.  (The threshold
// used to be 64 rows; at 9..63 rows -- e.g. 32 LM-Listener chunks per batch -- the decode then streamed fp32 weights through the
// tiled FFMA kernel: 3.1 s per 1023-step batch, profiles/r01_notes.md.)
inline bool tc_on(const TcCtx& tc, int M) { return tc.planes > 0 && M > 8; }

// split_hint: DIM_SPLIT_NEVER for every prefill GEMM (M = clips x frames), DIM_SPLIT_DECODE for the per-step GEMMs (M = clips).
// The split-K factor then depends on the call site and on (N, K) only -- never on how many clips share the batch -- so a
// clip's bits do not depend on the batch it is decoded in (test_concurrent_group_decoding_is_bit_identical).
int run_gemm(const TcCtx& tc, const GemmArgs& a_in, __nv_bfloat16* scratch, cudaStream_t s, int split_hint = DIM_SPLIT_NEVER) {
  GemmArgs a = a_in;
  a.split_hint = split_hint;
  auto it = tc.wmap.find(a.W);
  const bool tc_ok = tc_on(tc, a.M) && it != tc.wmap.end() && (scratch != nullptr || a.Ap != nullptr);
  if (!tc_ok) {
    if (a.Ap != nullptr || a.Cp != nullptr) return fail(DIM_EINVAL, "run_gemm: plane operands without a tensor-core path");
    // <= 8 rows in bf16 mode: the GEMV reads the bf16 weight image (half the bytes of the stream that bounds a small-batch decode
    // step; VERDICT r01: "B <= 8 ignores the precision mode").  DIM_GEMV_FP32=1 keeps the fp32 weights (A/B hook).
    static const bool gemv_fp32 = getenv("DIM_GEMV_FP32") != nullptr;
    if (tc.planes == 1 && a.M <= 8 && a.conv_T == 0 && a.K % 8 == 0 && it != tc.wmap.end() && !gemv_fp32)
      return launch_gemv_bf16w(a, it->second, tc_round_k(a.K), s);
    return launch_gemm_f32(a, s);
  }
  const int kp = tc_round_k(a.K);
  if (a.Ap != nullptr) return launch_gemm_tc(a, a.Ap, it->second, kp, tc.planes, s);
  // 5-tap conv: implicit GEMM over the padded frame planes (five row-shifted TMA boxes per k-range) when the channel count allows;
  // DIM_CONV_IM2COL=1 keeps the explicit im2col plane rows (A/B hook; bit-identical results: same products, same order)
  static const bool im2col = getenv("DIM_CONV_IM2COL") != nullptr;
  if (a.conv_T > 0 && a.conv_C % 64 == 0 && a.M % a.conv_T == 0 && !im2col) {
    if (int e = launch_split_conv_pad(a, scratch, tc.planes, s)) return e;
    return launch_gemm_tc(a, scratch, it->second, kp, tc.planes, s);
  }
  GemmArgs flat = a;                                    // explicit path: the GEMM sees a plain [M, 5C] operand
  flat.conv_T = 0;
  if (int e = launch_split_planes(a, scratch, kp, tc.planes, s)) return e;
  return launch_gemm_tc(flat, scratch, it->second, kp, tc.planes, s);
}

int build_vq_stack(dim_handle_s* h, const std::string& p, const dim_vq_config& c, std::vector<VqLayer>& out) {
  const int64_t H = c.hidden, F = c.ffn;
  out.resize(c.layers);
  for (int l = 0; l < c.layers; ++l) {
    VqLayer& L = out[l];
    std::string a = p + ".net." + std::to_string(2 * l) + ".fn", m = p + ".net." + std::to_string(2 * l + 1) + ".fn";
    LOOKUP(L.ln1_g, a + ".norm.weight", true, H);
    LOOKUP(L.ln1_b, a + ".norm.bias", true, H);
    LOOKUP(L.wqkv, a + ".fn.to_qkv.weight", true, 3 * H, H);
    LOOKUP(L.wo, a + ".fn.to_out.weight", true, H, H);
    LOOKUP(L.bo, a + ".fn.to_out.bias", true, H);
    LOOKUP(L.ln2_g, m + ".norm.weight", true, H);
    LOOKUP(L.ln2_b, m + ".norm.bias", true, H);
    LOOKUP(L.w1, m + ".fn.l1.weight", true, F, H);
    LOOKUP(L.b1, m + ".fn.l1.bias", true, F);
    LOOKUP(L.w2, m + ".fn.l2.weight", true, H, F);
    LOOKUP(L.b2, m + ".fn.l2.bias", true, H);
  }
  return DIM_OK;
}

// ---- VQ-VAE workspace layout (floats per frame row) ---------------------------------------------------------------
struct VqWs {
  float *h0, *h1, *x, *ln, *qkv, *att, *ff, *z;
  __nv_bfloat16 *ap, *ap2;
  int64_t* idx;
  size_t bytes;
};
VqWs carve_vq(const dim_vq_config& c, int planes, int B, int T, void* base) {
  size_t R = (size_t)B * T;
  char* p = static_cast<char*>(base);
  VqWs w{};
  auto take = [&](size_t nfloat) {
    float* r = reinterpret_cast<float*>(p);
    p += align_up(nfloat * sizeof(float), 256);
    return r;
  };
  w.h0 = take(R * c.hidden); w.h1 = take(R * c.hidden); w.x = take(R * c.hidden); w.ln = take(R * c.hidden);
  w.qkv = take(R * 3 * c.hidden); w.att = take(R * c.hidden); w.ff = take(R * c.ffn);
  const size_t fq = (size_t)std::max(1, c.fqn);
  w.z = take(R * fq * c.zdim);
  w.idx = reinterpret_cast<int64_t*>(take(R * fq * 2));
  const size_t kmax = (size_t)tc_round_k(std::max(5 * c.hidden, c.ffn));        // widest A operand: the conv's im2col rows
  w.ap = planes ? reinterpret_cast<__nv_bfloat16*>(take(R * planes * kmax / 2 + 64)) : nullptr;
  w.ap2 = planes ? reinterpret_cast<__nv_bfloat16*>(take(R * planes * (size_t)tc_round_k(c.ffn) / 2 + 64)) : nullptr;
  w.bytes = (size_t)(p - static_cast<char*>(base));
  return w;
}

// conv(k5, replicate) + LeakyReLU + InstanceNorm, Linear + pe[batch], transformer stack: shared by encoder and decoder.
int vq_trunk(const VqModel& m, const std::vector<VqLayer>& layers, const float* conv_wr, const float* conv_b,
             const float* emb_w, const float* emb_b, const float* pe, VqWs& w, const int32_t* lens,
             const int32_t* batch_index, int B, int T, cudaStream_t s) {
  const dim_vq_config& c = m.cfg;
  const int R = B * T, H = c.hidden;
  {  // h1 = leaky(conv5(h0) + b)
    GemmArgs a;
    a.A = w.h0; a.lda = H; a.W = conv_wr; a.bias = conv_b; a.C = w.h1; a.ldc = H; a.M = R; a.N = H; a.K = 5 * H;
    a.act = DIM_ACT_LEAKY; a.slope = c.neg_slope; a.conv_T = T; a.conv_C = H; a.lens = lens;
    if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
  }
  if (int e = launch_instance_norm(w.h1, lens, B, T, H, 1e-5f, s)) return e;
  {  // x = h1 @ emb^T + b + pe[batch_index[b]]
    GemmArgs a;
    a.A = w.h1; a.lda = H; a.W = emb_w; a.bias = emb_b; a.C = w.x; a.ldc = H; a.M = R; a.N = H; a.K = H;
    a.tab = pe; a.ldtab = H; a.tab_mode = 1; a.tab_index = batch_index; a.tab_T = T;
    if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
  }
  const int Dh = H / c.heads;
  const bool tcp = tc_on(m.tc, R);                      // plane-fused path: LN / attention / FF1 emit bf16 planes directly
  const int P = m.tc.planes;
  for (const VqLayer& L : layers) {
    if (int e = launch_layer_norm(w.x, L.ln1_g, L.ln1_b, tcp ? nullptr : w.ln, nullptr, R, H, 1e-5f, s, tcp ? w.ap : nullptr, P, H))
      return e;
    static const bool qkv_fp32 = getenv("DIM_ATTN_QKV_FP32") != nullptr;
    const bool q16 = tcp && P == 1 && !qkv_fp32;        // plain-bf16 mode: bf16 QKV straight into the attention kernel (see xt_encoder_forward)
    __nv_bfloat16* qkv16 = reinterpret_cast<__nv_bfloat16*>(w.qkv);
    {
      GemmArgs a;
      a.A = w.ln; a.lda = H; a.Ap = tcp ? w.ap : nullptr; a.W = L.wqkv; a.M = R; a.N = 3 * H; a.K = H;
      if (q16) { a.Cb = qkv16; a.ldcb = 3 * H; } else { a.C = w.qkv; a.ldc = 3 * H; }
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
    {
      AttnArgs a;                      // 'b n (qkv h d)': q at col 0, k at H, v at 2H
      a.q = w.qkv; a.k = w.qkv + H; a.v = w.qkv + 2 * H; a.ldq = a.ldk = a.ldv = 3 * H;
      if (q16) {
        a.in_bf16 = 1;
        a.q = reinterpret_cast<const float*>(qkv16); a.k = reinterpret_cast<const float*>(qkv16 + H);
        a.v = reinterpret_cast<const float*>(qkv16 + 2 * H);
      }
      a.out = tcp ? nullptr : w.att; a.ldo = H; a.out_p = tcp ? w.ap : nullptr; a.planes = P; a.kp = H;
      a.lens = lens; a.B = B; a.H = c.heads; a.Tq = T; a.Tk = T; a.Dh = Dh;
      a.scale = 1.0f / sqrtf((float)H);          // hidden_size**-0.5, not head_dim
      if (int e = launch_attention_prefill(a, s)) return e;
    }
    {
      GemmArgs a;
      a.A = w.att; a.lda = H; a.Ap = tcp ? w.ap : nullptr; a.W = L.wo; a.bias = L.bo; a.residual = w.x; a.ldr = H; a.C = w.x;
      a.ldc = H; a.M = R; a.N = H; a.K = H;
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
    if (int e = launch_layer_norm(w.x, L.ln2_g, L.ln2_b, tcp ? nullptr : w.ln, nullptr, R, H, 1e-5f, s, tcp ? w.ap : nullptr, P, H))
      return e;
    {
      GemmArgs a;
      a.A = w.ln; a.lda = H; a.Ap = tcp ? w.ap : nullptr; a.W = L.w1; a.bias = L.b1; a.M = R; a.N = c.ffn; a.K = H;
      a.act = DIM_ACT_GELU_TANH;
      if (tcp) { a.Cp = w.ap2; a.cp_planes = P; a.cp_kp = c.ffn; } else { a.C = w.ff; a.ldc = c.ffn; }
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
    {
      GemmArgs a;
      a.A = w.ff; a.lda = c.ffn; a.Ap = tcp ? w.ap2 : nullptr; a.W = L.w2; a.bias = L.b2; a.residual = w.x; a.ldr = H; a.C = w.x;
      a.ldc = H; a.M = R; a.N = H; a.K = c.ffn;
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
  }
  return DIM_OK;
}

// ---- x-transformers pieces ----------------------------------------------------------------------------------------
int build_xt_attn(dim_handle_s* h, const std::string& lp, int dim, int ctx_dim, int inner, bool cross, XtAttn& A) {
  const float *wk, *wv;
  LOOKUP(A.norm_g, lp + ".0.0.weight", true, dim);
  LOOKUP(A.norm_b, lp + ".0.0.bias", false, dim);
  LOOKUP(A.wq, lp + ".1.to_q.weight", true, inner, dim);
  LOOKUP(wk, lp + ".1.to_k.weight", true, inner, ctx_dim);
  LOOKUP(wv, lp + ".1.to_v.weight", true, inner, ctx_dim);
  LOOKUP(A.wo, lp + ".1.to_out.weight", true, dim, inner);
  size_t qb = (size_t)inner * dim * sizeof(float), kb = (size_t)inner * ctx_dim * sizeof(float);
  if (!cross) {
    if (int e = owned_alloc(h, qb + 2 * kb, &A.wqkv)) return e;
    DIM_CHECK_CUDA(cudaMemcpy(A.wqkv, A.wq, qb, cudaMemcpyDeviceToDevice));
    DIM_CHECK_CUDA(cudaMemcpy(A.wqkv + (size_t)inner * dim, wk, kb, cudaMemcpyDeviceToDevice));
    DIM_CHECK_CUDA(cudaMemcpy(A.wqkv + (size_t)2 * inner * dim, wv, kb, cudaMemcpyDeviceToDevice));
  } else {
    if (int e = owned_alloc(h, 2 * kb, &A.wkv)) return e;
    DIM_CHECK_CUDA(cudaMemcpy(A.wkv, wk, kb, cudaMemcpyDeviceToDevice));
    DIM_CHECK_CUDA(cudaMemcpy(A.wkv + (size_t)inner * ctx_dim, wv, kb, cudaMemcpyDeviceToDevice));
  }
  return DIM_OK;
}

int build_xt_ff(dim_handle_s* h, const std::string& lp, int dim, int mult, XtFF& F) {
  LOOKUP(F.norm_g, lp + ".0.0.weight", true, dim);
  LOOKUP(F.norm_b, lp + ".0.0.bias", false, dim);
  LOOKUP(F.w1, lp + ".1.ff.0.0.weight", true, (int64_t)mult * dim, dim);
  LOOKUP(F.b1, lp + ".1.ff.0.0.bias", true, (int64_t)mult * dim);
  LOOKUP(F.w2, lp + ".1.ff.2.weight", true, dim, (int64_t)mult * dim);
  LOOKUP(F.b2, lp + ".1.ff.2.bias", true, dim);
  return DIM_OK;
}

int build_xt_encoder(dim_handle_s* h, const std::string& name, int dim_in, const dim_s2s_config& c, XtEncoder& E) {
  const int inner = c.heads * c.dim_head;
  E.dim_in = dim_in;
  LOOKUP(E.proj_w, name + ".project_in.weight", true, c.dim, dim_in);
  LOOKUP(E.proj_b, name + ".project_in.bias", false, c.dim);
  LOOKUP(E.pos_emb, name + ".pos_emb.emb.weight", true, c.max_seq_len, c.dim);
  E.attn.resize(c.depth);
  E.ff.resize(c.depth);
  for (int l = 0; l < c.depth; ++l) {
    std::string p = name + ".attn_layers.layers.";
    if (int e = build_xt_attn(h, p + std::to_string(2 * l), c.dim, c.dim, inner, false, E.attn[l])) return e;
    if (int e = build_xt_ff(h, p + std::to_string(2 * l + 1), c.dim, c.ff_mult, E.ff[l])) return e;
  }
  LOOKUP(E.final_g, name + ".attn_layers.final_norm.weight", true, c.dim);
  LOOKUP(E.final_b, name + ".attn_layers.final_norm.bias", false, c.dim);
  return DIM_OK;
}

struct CtxWs {
  float *x, *ln, *qkv, *att, *ff;
  __nv_bfloat16 *ap, *ap2;
  size_t bytes;
};
CtxWs carve_ctx(const dim_s2s_config& c, int planes, int B, int T, void* base) {
  size_t R = (size_t)B * T;
  const int inner = c.heads * c.dim_head;
  char* p = static_cast<char*>(base);
  CtxWs w{};
  auto take = [&](size_t nfloat) {
    float* r = reinterpret_cast<float*>(p);
    p += align_up(nfloat * sizeof(float), 256);
    return r;
  };
  w.x = take(R * c.dim); w.ln = take(R * c.dim); w.qkv = take(R * 3 * inner); w.att = take(R * inner);
  w.ff = take(R * c.ff_mult * c.dim);
  w.ap = planes ? reinterpret_cast<__nv_bfloat16*>(take(R * planes * (size_t)tc_round_k(c.ff_mult * c.dim) / 2 + 64)) : nullptr;
  w.ap2 = planes ? reinterpret_cast<__nv_bfloat16*>(take(R * planes * (size_t)tc_round_k(c.ff_mult * c.dim) / 2 + 64)) : nullptr;
  w.bytes = (size_t)(p - static_cast<char*>(base));
  return w;
}

// ContinuousTransformerWrapper(x, mask, attn_mask=causal, return_embeddings=True); result left in w.x
int xt_encoder_forward(const S2SModel& m, const XtEncoder& E, const float* in, const float* a_add, const uint8_t* mask,
                       CtxWs& w, int B, int T, cudaStream_t s, int causal = 1) {
  const dim_s2s_config& c = m.cfg;
  const int R = B * T, D = c.dim, inner = c.heads * c.dim_head, F = c.ff_mult * c.dim;
  {  // x = project_in(in + a_add) + pos_emb[t] * dim^-0.5
    GemmArgs a;
    a.A = in; a.lda = E.dim_in; a.W = E.proj_w; a.bias = E.proj_b; a.C = w.x; a.ldc = D; a.M = R; a.N = D; a.K = E.dim_in;
    a.a_add = a_add; a.tab = E.pos_emb; a.ldtab = D; a.tab_mode = 2; a.tab_T = T; a.tab_scale = 1.0f / sqrtf((float)D);
    if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
  }
  const bool tcp = tc_on(m.tc, R);
  const int P = m.tc.planes;
  for (int l = 0; l < c.depth; ++l) {
    const XtAttn& A = E.attn[l];
    const XtFF& FF = E.ff[l];
    if (int e = launch_layer_norm(w.x, A.norm_g, A.norm_b, tcp ? nullptr : w.ln, nullptr, R, D, 1e-5f, s, tcp ? w.ap : nullptr, P, D))
      return e;
    // plain-bf16 mode: the QKV GEMM writes bf16 and the attention kernel stages it with cp.async -- the same roundings the fp32
    // round trip produced inside the attention kernel, at half the bytes (DIM_ATTN_QKV_FP32=1: A/B hook)
    static const bool qkv_fp32 = getenv("DIM_ATTN_QKV_FP32") != nullptr;
    const bool q16 = tcp && P == 1 && !qkv_fp32;
    __nv_bfloat16* qkv16 = reinterpret_cast<__nv_bfloat16*>(w.qkv);
    {
      GemmArgs a;
      a.A = w.ln; a.lda = D; a.Ap = tcp ? w.ap : nullptr; a.W = A.wqkv; a.M = R; a.N = 3 * inner;
      a.K = D;
      if (q16) { a.Cb = qkv16; a.ldcb = 3 * inner; } else { a.C = w.qkv; a.ldc = 3 * inner; }
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
    {
      AttnArgs a;
      a.q = w.qkv; a.k = w.qkv + inner; a.v = w.qkv + 2 * inner; a.ldq = a.ldk = a.ldv = 3 * inner;
      if (q16) {
        a.in_bf16 = 1;
        a.q = reinterpret_cast<const float*>(qkv16); a.k = reinterpret_cast<const float*>(qkv16 + inner);
        a.v = reinterpret_cast<const float*>(qkv16 + 2 * inner);
      }
      a.out = tcp ? nullptr : w.att; a.ldo = inner; a.out_p = tcp ? w.ap : nullptr; a.planes = P; a.kp = inner;
      a.key_mask = mask; a.B = B; a.H = c.heads; a.Tq = T; a.Tk = T; a.Dh = c.dim_head;
      a.scale = 1.0f / sqrtf((float)c.dim_head); a.causal = causal;
      if (int e = launch_attention_prefill(a, s)) return e;
    }
    {
      GemmArgs a;
      a.A = w.att; a.lda = inner; a.Ap = tcp ? w.ap : nullptr; a.W = A.wo; a.residual = w.x; a.ldr = D; a.C = w.x; a.ldc = D;
      a.M = R; a.N = D; a.K = inner;
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
    if (int e = launch_layer_norm(w.x, FF.norm_g, FF.norm_b, tcp ? nullptr : w.ln, nullptr, R, D, 1e-5f, s, tcp ? w.ap : nullptr, P, D))
      return e;
    {
      GemmArgs a;
      a.A = w.ln; a.lda = D; a.Ap = tcp ? w.ap : nullptr; a.W = FF.w1; a.bias = FF.b1; a.M = R; a.N = F; a.K = D;
      a.act = DIM_ACT_GELU_ERF;
      if (tcp) { a.Cp = w.ap2; a.cp_planes = P; a.cp_kp = F; } else { a.C = w.ff; a.ldc = F; }
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
    {
      GemmArgs a;
      a.A = w.ff; a.lda = F; a.Ap = tcp ? w.ap2 : nullptr; a.W = FF.w2; a.bias = FF.b2; a.residual = w.x; a.ldr = D; a.C = w.x;
      a.ldc = D; a.M = R; a.N = D; a.K = F;
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
  }
  return launch_layer_norm(w.x, E.final_g, E.final_b, w.x, nullptr, R, D, 1e-5f, s);
}

struct GenWs {
  std::vector<float*> cross_kv;            // per layer, head-major: K[B,H,T,64] then V[B,H,T,64]
  std::vector<float*> self_k, self_v;      // per layer, head-major [B,H,steps+1,64]
  float* kv_tmp;                           // token-major [B*T, 2*inner] projection output, re-laid out into cross_kv[l]
  float *x, *ln, *qkv, *att, *ff, *logits;
  int64_t* tokens;                         // [B, steps+1]
  int* step;
  int32_t* key_valid;                      // [B] prefix length of each clip's key-padding mask, -1 = not a prefix mask (persistent kernel)
  uint8_t* mask_stage;                     // [B,T] copy of the caller's key-padding mask   } what the captured step graph reads:
  float* u_stage;                          // [R,steps] copy of the caller's uniforms       } the graph never holds caller pointers
  __nv_bfloat16 *ap, *ap2;                 // A-operand plane scratch for the tensor-core GEMMs
  // persistent decode kernel (decode_mk.cu): split-K partials [MK_PART_FLOATS per row], attention-output planes, grid barrier, trace
  float* mk_part;
  __nv_bfloat16* mk_attp;
  unsigned int* mk_bar;
  unsigned long long* mk_trace;
  size_t bytes;
};
// B clips, S samples per clip: the cross-attention K/V exist once per clip, everything else once per decode row (R = B*S).
GenWs carve_gen(const dim_s2s_config& c, int planes, bool kv_bf16, int B, int T, int steps, void* base, int S = 1) {
  const int inner = c.heads * c.dim_head, D = c.dim + c.dim_audio;
  const size_t R = (size_t)B * S;
  char* p = static_cast<char*>(base);
  GenWs w{};
  auto take = [&](size_t nfloat) {
    float* r = reinterpret_cast<float*>(p);
    p += align_up(nfloat * sizeof(float), 256);
    return r;
  };
  const size_t kvdiv = kv_bf16 ? 2 : 1;                 // bf16 caches take half the floats
  for (int l = 0; l < c.depth; ++l) w.cross_kv.push_back(take((size_t)B * T * 2 * inner / kvdiv));
  w.kv_tmp = take((size_t)B * T * 2 * inner / kvdiv);
  for (int l = 0; l < c.depth; ++l) {
    w.self_k.push_back(take(R * (steps + 1) * inner / kvdiv));
    w.self_v.push_back(take(R * (steps + 1) * inner / kvdiv));
  }
  w.x = take(R * D); w.ln = take(R * D); w.qkv = take(R * 3 * inner);
  w.att = take(R * inner); w.ff = take(R * c.ff_mult * D); w.logits = take(R * c.num_tokens);
  w.tokens = reinterpret_cast<int64_t*>(take(R * (steps + 1) * 2));
  w.step = reinterpret_cast<int*>(take(64));
  w.key_valid = reinterpret_cast<int32_t*>(take((size_t)B + 1));
  w.mask_stage = reinterpret_cast<uint8_t*>(take(((size_t)B * T + 3) / 4 + 1));
  w.u_stage = take(R * (size_t)steps + 1);
  {
    const size_t rows_ctx = (size_t)B * T, rows_step = R;
    const size_t need = std::max(rows_ctx * tc_round_k(D), rows_step * tc_round_k(c.ff_mult * D)) * planes;
    w.ap = planes ? reinterpret_cast<__nv_bfloat16*>(take(need / 2 + 64)) : nullptr;
    w.ap2 = planes ? reinterpret_cast<__nv_bfloat16*>(take(rows_step * tc_round_k(c.ff_mult * D) * planes / 2 + 64)) : nullptr;
  }
  w.mk_part = planes ? take(R * MK_PART_FLOATS) : nullptr;
  w.mk_attp = planes ? reinterpret_cast<__nv_bfloat16*>(take(R * (size_t)tc_round_k(inner) * planes / 2 + 64)) : nullptr;
  w.mk_bar = reinterpret_cast<unsigned int*>(take(64));
  w.mk_trace = reinterpret_cast<unsigned long long*>(take(2 * MK_MAX_PHASES));
  w.bytes = (size_t)(p - static_cast<char*>(base));
  return w;
}

}  // namespace

// ======================================================== C ABI ========================================================

extern "C" int dim_create(dim_handle_t* out, int device) {
  DIM_REQUIRE(out != nullptr, "dim_create: null out");
  DIM_CHECK_CUDA(cudaSetDevice(device));
  if (int e = ensure_device()) return e;
  auto* h = new dim_handle_s();
  h->device = device;
  *out = h;
  return DIM_OK;
}

extern "C" int dim_destroy(dim_handle_t h) {
  if (!h) return DIM_OK;
  cudaSetDevice(h->device);
  for (auto& m : h->s2s) {
    for (auto& G : m->graphs) {
      if (G.exec) cudaGraphExecDestroy(G.exec);
      if (G.exec_u) cudaGraphExecDestroy(G.exec_u);
      if (G.fork) cudaEventDestroy(G.fork);
      if (G.join) cudaEventDestroy(G.join);
      if (G.stream) cudaStreamDestroy(G.stream);
    }
    if (m->fork_ev) cudaEventDestroy(m->fork_ev);
  }
  for (void* p : h->owned) cudaFree(p);
  delete h;
  return DIM_OK;
}

extern "C" int dim_set_tensor(dim_handle_t h, const char* name, const void* ptr, int dtype, int ndim, const int64_t* shape) {
  DIM_REQUIRE(h && name && ptr && ndim >= 0 && ndim <= 8, "dim_set_tensor: bad argument");
  TensorRef t;
  t.p = ptr;
  t.dtype = dtype;
  t.shape.assign(shape, shape + ndim);
  h->tensors[name] = t;
  return DIM_OK;
}

extern "C" int dim_vqvae_build_parts(dim_handle_t h, const char* prefix_c, const char* enc_c, const char* dec_c, const dim_vq_config* cfg,
                                     int precision, int* model) {
  DIM_REQUIRE(h && cfg && model, "dim_vqvae_build: null argument");
  DIM_CHECK_CUDA(cudaSetDevice(h->device));            // the handle's device, whatever the caller's current device is
  DIM_REQUIRE(precision == DIM_PREC_FP32 || precision == DIM_PREC_FP32_TC || precision == DIM_PREC_BF16,
              "dim_vqvae_build: unknown precision");
  DIM_REQUIRE(enc_c || dec_c, "dim_vqvae_build: a model needs an encoder or a decoder");
  dim_vq_config c = *cfg;
  if (c.fqn <= 0) c.fqn = 1;
  if (c.out_dim <= 0) c.out_dim = c.in_dim;
  DIM_REQUIRE(c.hidden % 16 == 0 && c.hidden % c.heads == 0, "hidden must be a multiple of 16 and of heads");
  DIM_REQUIRE(c.hidden / c.heads == 48 || c.hidden / c.heads == 64 || c.hidden / c.heads == 96, "head dim must be 48, 64 or 96");
  DIM_REQUIRE(c.in_dim % 4 == 0 && c.out_dim % 4 == 0 && c.zdim % 4 == 0 && c.ffn % 4 == 0, "dims must be multiples of 4");
  std::string p = prefix_c ? prefix_c : "";
  auto m = std::make_unique<VqModel>();
  m->cfg = c;
  m->precision = precision;
  m->tc.planes = planes_of(precision);
  m->has_enc = enc_c != nullptr;
  m->has_dec = dec_c != nullptr;
  m->fqn = c.fqn;
  m->out_dim = c.out_dim;
  const int64_t H = c.hidden, Z = c.zdim, ZT = (int64_t)c.fqn * c.zdim;
  const float *conv_w = nullptr, *dconv_w = nullptr;
  const size_t cb = (size_t)H * H * 5 * sizeof(float);
  if (m->has_enc) {
    const std::string e = p + enc_c + ".";
    LOOKUP(m->map_w, e + "vertice_mapping.0.weight", true, H, c.in_dim);
    LOOKUP(m->map_b, e + "vertice_mapping.0.bias", true, H);
    LOOKUP(conv_w, e + "squasher.0.0.weight", true, H, H, 5);
    LOOKUP(m->conv_b, e + "squasher.0.0.bias", true, H);
    LOOKUP(m->emb_w, e + "encoder_linear_embedding.net.weight", true, H, H);
    LOOKUP(m->emb_b, e + "encoder_linear_embedding.net.bias", true, H);
    LOOKUP(m->pe, e + "encoder_pos_embedding.pe", true, c.pe_max_len, 1, H);
    LOOKUP(m->post_w, e + "encoder_linear_embedding_post.net.weight", true, ZT, H);
    LOOKUP(m->post_b, e + "encoder_linear_embedding_post.net.bias", true, ZT);
    if (int err = build_vq_stack(h, e + "encoder_transformer", c, m->enc)) return err;
    if (int err = owned_alloc(h, cb, &m->conv_wr)) return err;
    if (int err = dim_repack_conv_weight(conv_w, m->conv_wr, c.hidden, c.hidden, nullptr)) return err;
  }
  if (m->has_dec) {
    const std::string d = p + dec_c + ".";
    LOOKUP(m->pre_w, d + "decoder_linear_embedding_pre.net.weight", true, H, ZT);
    LOOKUP(m->pre_b, d + "decoder_linear_embedding_pre.net.bias", true, H);
    LOOKUP(dconv_w, d + "expander.0.0.weight", true, H, H, 5);
    LOOKUP(m->dconv_b, d + "expander.0.0.bias", true, H);
    LOOKUP(m->demb_w, d + "decoder_linear_embedding.net.weight", true, H, H);
    LOOKUP(m->demb_b, d + "decoder_linear_embedding.net.bias", true, H);
    LOOKUP(m->dpe, d + "decoder_pos_embedding.pe", true, c.pe_max_len, 1, H);
    LOOKUP(m->rev_w, d + "vertice_map_reverse.weight", true, c.out_dim, H);
    if (int err = build_vq_stack(h, d + "decoder_transformer", c, m->dec)) return err;
    if (int err = owned_alloc(h, cb, &m->dconv_wr)) return err;
    if (int err = dim_repack_conv_weight(dconv_w, m->dconv_wr, c.hidden, c.hidden, nullptr)) return err;
  }
  LOOKUP(m->codebook, p + "quantize.embedding.weight", true, c.n_embed, Z);
  {
    TcCtx& tc = m->tc;
    const int Hh = c.hidden;
    if (m->has_enc) {
      if (int e = tc_add_weight(h, tc, m->map_w, Hh, c.in_dim)) return e;
      if (int e = tc_add_weight(h, tc, m->conv_wr, Hh, 5 * Hh)) return e;
      if (int e = tc_add_weight(h, tc, m->emb_w, Hh, Hh)) return e;
      if (int e = tc_add_weight(h, tc, m->post_w, (int)ZT, Hh)) return e;
    }
    if (m->has_dec) {
      if (int e = tc_add_weight(h, tc, m->pre_w, Hh, (int)ZT)) return e;
      if (int e = tc_add_weight(h, tc, m->dconv_wr, Hh, 5 * Hh)) return e;
      if (int e = tc_add_weight(h, tc, m->demb_w, Hh, Hh)) return e;
      if (int e = tc_add_weight(h, tc, m->rev_w, c.out_dim, Hh)) return e;
    }
    for (auto* stack : {&m->enc, &m->dec})
      for (const VqLayer& L : *stack) {
        if (int e = tc_add_weight(h, tc, L.wqkv, 3 * Hh, Hh)) return e;
        if (int e = tc_add_weight(h, tc, L.wo, Hh, Hh)) return e;
        if (int e = tc_add_weight(h, tc, L.w1, c.ffn, Hh)) return e;
        if (int e = tc_add_weight(h, tc, L.w2, Hh, c.ffn)) return e;
      }
  }
  DIM_CHECK_CUDA(cudaStreamSynchronize(nullptr));
  h->vq.push_back(std::move(m));
  *model = (int)h->vq.size() - 1;
  return DIM_OK;
}

extern "C" int dim_vqvae_build(dim_handle_t h, const char* prefix_c, const dim_vq_config* cfg, int precision, int* model) {
  return dim_vqvae_build_parts(h, prefix_c, "encoder", "decoder", cfg, precision, model);
}

extern "C" size_t dim_vqvae_workspace_bytes(dim_handle_t h, int model, int B, int T) {
  if (!h || model < 0 || model >= (int)h->vq.size() || B <= 0 || T <= 0) return 0;
  return carve_vq(h->vq[model]->cfg, h->vq[model]->tc.planes, B, T, nullptr).bytes;
}

extern "C" int dim_vqvae_encode(dim_handle_t h, int model, const float* x, const int32_t* lens, const int32_t* batch_index,
                                int B, int T, int64_t* idx, float* z, float* quant_bcl, void* ws, size_t ws_bytes,
                                void* stream) {
  DIM_REQUIRE(h && model >= 0 && model < (int)h->vq.size(), "dim_vqvae_encode: bad model");
  DIM_CHECK_CUDA(cudaSetDevice(h->device));            // the handle's device, whatever the caller's current device is
  DIM_REQUIRE(x && B > 0 && T > 0 && (idx || z), "dim_vqvae_encode: bad argument");
  const VqModel& m = *h->vq[model];
  DIM_REQUIRE(m.has_enc, "dim_vqvae_encode: this model was built without an encoder");
  const dim_vq_config& c = m.cfg;
  const int ZT = m.fqn * c.zdim;
  DIM_REQUIRE(batch_index != nullptr || B <= c.pe_max_len, "batch larger than the positional table (SURVEY F4/H3)");
  VqWs w = carve_vq(c, m.tc.planes, B, T, ws);
  if (ws == nullptr || ws_bytes < w.bytes) return fail(DIM_EWORKSPACE, "dim_vqvae_encode: workspace too small");
  cudaStream_t s = as_stream(stream);
  const int R = B * T, H = c.hidden;
  {  // h0 = leaky(x @ map^T + b)
    GemmArgs a;
    a.A = x; a.lda = c.in_dim; a.W = m.map_w; a.bias = m.map_b; a.C = w.h0; a.ldc = H; a.M = R; a.N = H; a.K = c.in_dim;
    a.act = DIM_ACT_LEAKY; a.slope = c.neg_slope;
    if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
  }
  if (int e = vq_trunk(m, m.enc, m.conv_wr, m.conv_b, m.emb_w, m.emb_b, m.pe, w, lens, batch_index, B, T, s)) return e;
  float* zbuf = z ? z : w.z;
  {
    GemmArgs a;
    a.A = w.x; a.lda = H; a.W = m.post_w; a.bias = m.post_b; a.C = zbuf; a.ldc = ZT; a.M = R; a.N = ZT; a.K = H;
    if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
  }
  // h.view(B, -1, zquant_dim) (stage1_BIWI.py:24-25,152-153): a frame's fqn*zdim channels are fqn consecutive tokens
  int64_t* ibuf = idx ? idx : w.idx;
  if (idx || quant_bcl)
    if (int e = launch_vq_argmin(zbuf, m.codebook, ibuf, R * m.fqn, c.zdim, c.n_embed, s)) return e;
  if (quant_bcl)
    if (int e = launch_vq_gather_bcl(ibuf, m.codebook, quant_bcl, B, T * m.fqn, c.zdim, c.n_embed, s)) return e;
  return DIM_OK;
}

extern "C" int dim_vqvae_decode(dim_handle_t h, int model, const int64_t* codes, const float* quant_bcl,
                                const int32_t* batch_index, int B, int L, float* out, void* ws, size_t ws_bytes,
                                void* stream) {
  DIM_REQUIRE(h && model >= 0 && model < (int)h->vq.size(), "dim_vqvae_decode: bad model");
  DIM_CHECK_CUDA(cudaSetDevice(h->device));            // the handle's device, whatever the caller's current device is
  DIM_REQUIRE((codes != nullptr) != (quant_bcl != nullptr), "dim_vqvae_decode: pass exactly one of codes / quant");
  DIM_REQUIRE(out && B > 0 && L > 0, "dim_vqvae_decode: bad argument");
  const VqModel& m = *h->vq[model];
  DIM_REQUIRE(m.has_dec, "dim_vqvae_decode: this model was built without a decoder");
  const dim_vq_config& c = m.cfg;
  const int ZT = m.fqn * c.zdim;
  DIM_REQUIRE(batch_index != nullptr || B <= c.pe_max_len, "batch larger than the positional table (SURVEY F4/H3)");
  VqWs w = carve_vq(c, m.tc.planes, B, L, ws);
  if (ws == nullptr || ws_bytes < w.bytes) return fail(DIM_EWORKSPACE, "dim_vqvae_decode: workspace too small");
  cudaStream_t s = as_stream(stream);
  const int R = B * L, H = c.hidden;
  if (codes) {
    if (int e = launch_vq_gather(codes, m.codebook, w.z, R * m.fqn, c.zdim, c.n_embed, nullptr, s)) return e;
  } else {
    if (int e = launch_rows_from_bcl(quant_bcl, w.z, B, L * m.fqn, c.zdim, s)) return e;
  }
  {
    GemmArgs a;
    a.A = w.z; a.lda = ZT; a.W = m.pre_w; a.bias = m.pre_b; a.C = w.h0; a.ldc = H; a.M = R; a.N = H; a.K = ZT;
    if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
  }
  if (int e = vq_trunk(m, m.dec, m.dconv_wr, m.dconv_b, m.demb_w, m.demb_b, m.dpe, w, nullptr, batch_index, B, L, s))
    return e;
  {
    GemmArgs a;
    a.A = w.x; a.lda = H; a.W = m.rev_w; a.C = out; a.ldc = m.out_dim; a.M = R; a.N = m.out_dim; a.K = H;
    if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
  }
  return DIM_OK;
}

extern "C" int dim_slmft_build(dim_handle_t h, const dim_s2s_config* cfg, int precision, int* model) {
  DIM_REQUIRE(h && cfg && model, "dim_slmft_build: null argument");
  DIM_CHECK_CUDA(cudaSetDevice(h->device));            // the handle's device, whatever the caller's current device is
  DIM_REQUIRE(precision == DIM_PREC_FP32 || precision == DIM_PREC_FP32_TC || precision == DIM_PREC_BF16,
              "dim_slmft_build: unknown precision");
  const dim_s2s_config& c = *cfg;
  DIM_REQUIRE(c.dim_head == 64, "x-transformers dim_head must be 64");
  DIM_REQUIRE(c.dim % 4 == 0 && c.dim_in % 4 == 0 && c.dim_audio % 4 == 0, "dims must be multiples of 4");
  DIM_REQUIRE(c.num_tokens <= 1024, "num_tokens must be <= 1024");
  auto m = std::make_unique<S2SModel>();
  m->cfg = c;
  m->precision = precision;
  m->tc.planes = planes_of(precision);
  const int inner = c.heads * c.dim_head, D = c.dim + c.dim_audio;
  if (int e = build_xt_encoder(h, "encoder_s", c.dim_in, c, m->enc_s)) return e;
  {  // single-encoder models (the older ListenerGenerator, seq2seq.py:172-182, registered as encoder_s + decoder_joint) have no joint
     // encoder, patch embeddings or norm_s: dim_slmft_context needs them, dim_slmft_encode / generate / teacher_forced do not
    const float* probe = nullptr;
    LOOKUP(probe, "encoder_joint.project_in.weight", false, c.dim, c.dim);
    if (probe) {
      if (int e = build_xt_encoder(h, "encoder_joint", c.dim, c, m->enc_joint)) return e;
      m->has_enc_joint = true;
    }
  }
  LOOKUP(m->patch_s, "patch_embed_s", false, 1, 1, c.dim_in);
  LOOKUP(m->patch_dec_s, "patch_embed_dec_s", false, 1, 1, c.dim);
  LOOKUP(m->norm_s_g, "norm_s.weight", false, c.dim);
  LOOKUP(m->norm_s_b, "norm_s.bias", false, c.dim);
  LOOKUP(m->norm_l_g, "norm_l.weight", false, c.dim);
  LOOKUP(m->norm_l_b, "norm_l.bias", false, c.dim);
  LOOKUP(m->norm_j_g, "norm.weight", false, c.dim);
  LOOKUP(m->norm_j_b, "norm.bias", false, c.dim);
  {  // the listener encoder exists in every checkpoint but only the SLM pre-training forward runs it (seq2seq_pretrain.py:218)
    const float* probe = nullptr;
    LOOKUP(probe, "encoder_l.project_in.weight", false, c.dim, c.dim_in);
    if (probe) {
      if (int e = build_xt_encoder(h, "encoder_l", c.dim_in, c, m->enc_l)) return e;
      m->has_enc_l = true;
    }
  }
  const std::string dn = "decoder_joint.net";
  LOOKUP(m->token_emb, dn + ".token_emb.emb.weight", true, c.num_tokens, D);
  LOOKUP(m->dec_pos_emb, dn + ".pos_emb.emb.weight", false, c.max_seq_len, D);
  m->self_attn.resize(c.depth);
  m->cross_attn.resize(c.depth);
  m->ff.resize(c.depth);
  for (int l = 0; l < c.depth; ++l) {
    std::string p = dn + ".attn_layers.layers.";
    if (int e = build_xt_attn(h, p + std::to_string(3 * l), D, D, inner, false, m->self_attn[l])) return e;
    if (int e = build_xt_attn(h, p + std::to_string(3 * l + 1), D, D, inner, true, m->cross_attn[l])) return e;
    if (int e = build_xt_ff(h, p + std::to_string(3 * l + 2), D, c.ff_mult, m->ff[l])) return e;
  }
  LOOKUP(m->final_g, dn + ".attn_layers.final_norm.weight", true, D);
  LOOKUP(m->final_b, dn + ".attn_layers.final_norm.bias", false, D);
  LOOKUP(m->logits_w, dn + ".to_logits.weight", true, c.num_tokens, D);
  LOOKUP(m->logits_b, dn + ".to_logits.bias", false, c.num_tokens);
  {
    TcCtx& tc = m->tc;
    for (XtEncoder* E : {&m->enc_s, &m->enc_joint, &m->enc_l}) {
      if (E == &m->enc_l && !m->has_enc_l) continue;
      if (E == &m->enc_joint && !m->has_enc_joint) continue;
      if (int e = tc_add_weight(h, tc, E->proj_w, c.dim, E->dim_in)) return e;
      for (int l = 0; l < c.depth; ++l) {
        if (int e = tc_add_weight(h, tc, E->attn[l].wqkv, 3 * inner, c.dim)) return e;
        if (int e = tc_add_weight(h, tc, E->attn[l].wo, c.dim, inner)) return e;
        if (int e = tc_add_weight(h, tc, E->ff[l].w1, c.ff_mult * c.dim, c.dim)) return e;
        if (int e = tc_add_weight(h, tc, E->ff[l].w2, c.dim, c.ff_mult * c.dim)) return e;
      }
    }
    for (int l = 0; l < c.depth; ++l) {
      if (int e = tc_add_weight(h, tc, m->self_attn[l].wqkv, 3 * inner, D)) return e;
      if (int e = tc_add_weight(h, tc, m->self_attn[l].wo, D, inner)) return e;
      if (int e = tc_add_weight(h, tc, m->cross_attn[l].wq, inner, D)) return e;
      if (int e = tc_add_weight(h, tc, m->cross_attn[l].wkv, 2 * inner, D)) return e;
      if (int e = tc_add_weight(h, tc, m->cross_attn[l].wo, D, inner)) return e;
      if (int e = tc_add_weight(h, tc, m->ff[l].w1, c.ff_mult * D, D)) return e;
      if (int e = tc_add_weight(h, tc, m->ff[l].w2, D, c.ff_mult * D)) return e;
    }
    if (int e = tc_add_weight(h, tc, m->logits_w, c.num_tokens, D)) return e;
    DIM_CHECK_CUDA(cudaStreamSynchronize(nullptr));
  }
  h->s2s.push_back(std::move(m));
  *model = (int)h->s2s.size() - 1;
  return DIM_OK;
}

// Decoding is a chain of ~49 short, strictly dependent kernels per generated frame: one chain cannot fill the GPU.
// Clips are independent, so a batch is decoded as up to 4 concurrent groups of >= 128 clips (one 128-row MMA tile each) on
// side streams: their chains interleave on the SMs and hide each other's launch/pipeline-fill/epilogue latencies.
// Per-row results do not depend on the grouping (no cross-row arithmetic; split-K depends on K only).
constexpr int kMaxGroups = 8;
int plan_groups(const S2SModel& m, int B, int max_keys, int* begin /*[kMaxGroups+1]*/) {
  {  // the persistent decode kernel owns every SM: one chain
    const dim_s2s_config& c = m.cfg;
    const int D = c.dim + c.dim_audio;
    static const bool small_mk = getenv("DIM_SMALL_BATCH_MK") != nullptr;
    if (g_decode_impl == 0 && m.tc.planes > 0 && (B > 8 || small_mk) &&
        mk_supported(D, c.heads * c.dim_head, c.ff_mult * D, c.num_tokens, c.heads, m.tc.planes, max_keys)) {
      begin[0] = 0;
      begin[1] = B;
      return 1;
    }
  }
  // tuning hooks: DIM_GROUP_ROWS (rows per group, multiple of 64; default 128), DIM_MAX_GROUPS, DIM_NO_GROUPS.
  // Default: plain bf16 operands decode as ONE chain (with the cp.async attention kernel and the narrow-tile decode GEMMs a
  // single chain measured faster than 2 concurrent groups: 271 vs 282 ms per step); the fp32-grade mode (3-6x longer GEMM
  // main loops) keeps up to 4 concurrent groups (160 k vs 151 k frames/s).  profiles/r01_notes.md
  static const int kGroupRows = getenv("DIM_GROUP_ROWS") ? std::max(64, atoi(getenv("DIM_GROUP_ROWS")) / 64 * 64) : 128;
  static const int env_groups = getenv("DIM_MAX_GROUPS") ? std::min(kMaxGroups, std::max(1, atoi(getenv("DIM_MAX_GROUPS")))) : 0;
  const int max_groups = env_groups ? env_groups : (m.tc.planes == 1 ? 1 : 4);
  int ng = 1;
  if (tc_on(m.tc, B) && B >= 2 * kGroupRows) ng = std::min(max_groups, B / kGroupRows);
  static const bool no_groups = getenv("DIM_NO_GROUPS") != nullptr;
  if (no_groups) ng = 1;
  const int per = (B / ng + kGroupRows - 1) / kGroupRows * kGroupRows;      // multiples of the group size, remainder in the last
  ng = std::min(ng, (B + per - 1) / per);                                   // rounding up may leave fewer non-empty groups
  for (int g = 0; g <= ng; ++g) begin[g] = std::min(B, g * per);
  begin[ng] = B;
  return ng;
}

extern "C" size_t dim_slmft_workspace_bytes(dim_handle_t h, int model, int B, int T, int steps) {
  if (!h || model < 0 || model >= (int)h->s2s.size() || B <= 0 || T <= 0) return 0;
  const dim_s2s_config& c = h->s2s[model]->cfg;
  const int planes = h->s2s[model]->tc.planes;
  size_t a = carve_ctx(c, planes, B, T, nullptr).bytes;
  size_t b = 0;
  if (steps > 0) {
    int begin[kMaxGroups + 1];
    const int ng = plan_groups(*h->s2s[model], B, std::max(T, steps + 1), begin);
    for (int g = 0; g < ng; ++g)
      b += carve_gen(c, planes, h->s2s[model]->precision == DIM_PREC_BF16, begin[g + 1] - begin[g], T, steps, nullptr).bytes;
  }
  return a > b ? a : b;
}

extern "C" int dim_slmft_context(dim_handle_t h, int model, const float* v_speaker, const float* v_audio,
                                 const uint8_t* mask, int B, int T, float* ctx, float* x_s, void* ws, size_t ws_bytes,
                                 void* stream) {
  DIM_REQUIRE(h && model >= 0 && model < (int)h->s2s.size(), "dim_slmft_context: bad model");
  DIM_CHECK_CUDA(cudaSetDevice(h->device));            // the handle's device, whatever the caller's current device is
  DIM_REQUIRE(v_speaker && (ctx || x_s) && (v_audio || !ctx) && B > 0 && T > 0, "dim_slmft_context: bad argument");
  const S2SModel& m = *h->s2s[model];
  const dim_s2s_config& c = m.cfg;
  DIM_REQUIRE(T <= c.max_seq_len, "sequence longer than the positional table");
  DIM_REQUIRE(m.has_enc_joint && m.patch_s && m.patch_dec_s && m.norm_s_g && m.norm_s_b,
              "dim_slmft_context: the model has no encoder_joint / patch embeddings / norm_s (single-encoder model: use dim_slmft_encode)");
  CtxWs w = carve_ctx(c, m.tc.planes, B, T, ws);
  if (ws == nullptr || ws_bytes < w.bytes) return fail(DIM_EWORKSPACE, "dim_slmft_context: workspace too small");
  cudaStream_t s = as_stream(stream);
  if (int e = xt_encoder_forward(m, m.enc_s, v_speaker, m.patch_s, mask, w, B, T, s)) return e;
  // encoder_joint's project_in writes w.x, so its input (encoder_s's output in w.x) is staged in w.att, which the
  // first attention of encoder_joint overwrites only after project_in has consumed it (same stream).
  DIM_CHECK_CUDA(cudaMemcpyAsync(w.att, w.x, (size_t)B * T * c.dim * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (int e = xt_encoder_forward(m, m.enc_joint, w.att, nullptr, mask, w, B, T, s)) return e;
  if (int e = launch_layer_norm(w.x, m.norm_s_g, m.norm_s_b, w.ln, nullptr, B * T, c.dim, 1e-5f, s)) return e;
  if (x_s) DIM_CHECK_CUDA(cudaMemcpyAsync(x_s, w.ln, (size_t)B * T * c.dim * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (!ctx) return DIM_OK;
  return launch_build_context(w.ln, m.patch_dec_s, v_audio, ctx, nullptr, (size_t)B * T, c.dim, c.dim_audio, s);
}

// One ContinuousTransformerWrapper call of the SLM pre-training forward (seq2seq_pretrain.py:216-224): encoder `which`
// (0 encoder_s, 1 encoder_l, 2 encoder_joint) over x (B,T,dim_in of that encoder) with the key-padding mask, optional causal
// attn_mask, then an optional LayerNorm head (0 none, 1 norm_s, 2 norm_l, 3 norm).  out (B,T,dim).
extern "C" int dim_slmft_encode(dim_handle_t h, int model, int which, const float* x, const float* add, const uint8_t* mask, int causal,
                                int norm, int B, int T, float* out, void* ws, size_t ws_bytes, void* stream) {
  DIM_REQUIRE(h && model >= 0 && model < (int)h->s2s.size(), "dim_slmft_encode: bad model");
  DIM_CHECK_CUDA(cudaSetDevice(h->device));
  DIM_REQUIRE(x && out && B > 0 && T > 0 && which >= 0 && which <= 2 && norm >= 0 && norm <= 3, "dim_slmft_encode: bad argument");
  const S2SModel& m = *h->s2s[model];
  const dim_s2s_config& c = m.cfg;
  DIM_REQUIRE(T <= c.max_seq_len, "sequence longer than the positional table");
  DIM_REQUIRE(which != 1 || m.has_enc_l, "dim_slmft_encode: encoder_l is not registered");
  DIM_REQUIRE(which != 2 || m.has_enc_joint, "dim_slmft_encode: encoder_joint is not registered");
  const float *g = nullptr, *b = nullptr;
  if (norm == 1) { g = m.norm_s_g; b = m.norm_s_b; }
  if (norm == 2) { g = m.norm_l_g; b = m.norm_l_b; }
  if (norm == 3) { g = m.norm_j_g; b = m.norm_j_b; }
  DIM_REQUIRE(norm == 0 || (g && b), "dim_slmft_encode: that LayerNorm is not registered");
  CtxWs w = carve_ctx(c, m.tc.planes, B, T, ws);
  if (ws == nullptr || ws_bytes < w.bytes) return fail(DIM_EWORKSPACE, "dim_slmft_encode: workspace too small");
  cudaStream_t s = as_stream(stream);
  const XtEncoder& E = which == 0 ? m.enc_s : (which == 1 ? m.enc_l : m.enc_joint);
  if (int e = xt_encoder_forward(m, E, x, add, mask, w, B, T, s, causal ? 1 : 0)) return e;
  if (norm) return launch_layer_norm(w.x, g, b, out, nullptr, B * T, c.dim, 1e-5f, s);
  DIM_CHECK_CUDA(cudaMemcpyAsync(out, w.x, (size_t)B * T * c.dim * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return DIM_OK;
}

namespace {

// ---- persistent decode kernel: the phase program of ONE decode step (decode_mk.cu) -------------------------------------------
// Tile width and split-K factor of a decode-step GEMM: ~72 work units per 128-row tile of decode rows (144 at 256 rows: one per
// SM).  A function of (N, K, planes) only -- never of the number of rows -- so a row's bits do not depend on its batch.
void mk_gemm_cfg(int N, int K, int planes, int* bn, int* splits) {
  const int kp = tc_round_k(K), npairs = planes == 1 ? 1 : (planes == 2 ? 3 : 6);
  const int total_it = kp / 64 * npairs;
  *bn = (N >= 2048 || kp >= 2048) ? 128 : 64;
  // fp32-grade planes: 6 products per k-block make the phase ingest-bound (A tile 16 KB + W tile per 128 x bn x 64 product), so the
  // wider tile with a deeper K split wins from N = 1024 on (DIM_MK_BN128=0: A/B hook)
  static const bool bn128 = !(getenv("DIM_MK_BN128") != nullptr && atoi(getenv("DIM_MK_BN128")) == 0);
  if (planes >= 2 && N >= 1024 && bn128) *bn = 128;
  const int nt = cdiv(N, *bn);
  int sp = (72 + nt / 2) / nt;
  sp = std::max(1, std::min(sp, 9));
  sp = std::min(sp, total_it);
  while (sp > 1 && sp * N > MK_PART_FLOATS) --sp;
  *splits = sp;
}

struct MkBuilder {
  MkPlan& P;
  const S2SModel& m;
  int nmaps = 0;
  int rows;
  MkBuilder(MkPlan& p, const S2SModel& mm, int r) : P(p), m(mm), rows(r) {}
  int add_map(const __nv_bfloat16* ptr, int nrows, int cols, int box_rows, int* idx) {
    if (nmaps >= MK_MAX_MAPS) return fail(DIM_EINVAL, "decode plan: too many tensor maps");
    if (int e = tc_make_map(ptr, nrows, cols, cols, box_rows, &P.maps[nmaps])) return e;
    map_ptr[nmaps] = ptr;
    *idx = nmaps++;
    return DIM_OK;
  }
  MkPhase* next(int type) {
    if (P.nphases >= MK_MAX_PHASES) return nullptr;
    MkPhase* ph = &P.phases[P.nphases++];
    ph->type = type;
    ph->M = rows;
    return ph;
  }
  // part[z][rows][N] = A_planes . W^T ; returns the split factor
  const __nv_bfloat16* map_ptr[MK_MAX_MAPS] = {};       // the operand behind an A map (the GEMV phases read it directly)
  int gemm(int mapA, const float* W, int N, int K, float* part, float w_keep, int* splits_out) {
    auto it = m.tc.wmap.find(W);
    if (it == m.tc.wmap.end()) return fail(DIM_EINVAL, "decode plan: weight without bf16 planes");
    MkPhase* ph = next(MK_GEMM);
    if (!ph) return fail(DIM_EINVAL, "decode plan: too many phases");
    const int planes = m.tc.planes, kp = tc_round_k(K);
    ph->N = N; ph->kp = kp; ph->kblocks = kp / 64;
    ph->npairs = tc_pairs(planes, ph->pa, ph->pw);
    mk_gemm_cfg(N, K, planes, &ph->bn, &ph->splits);
    ph->mapA = mapA;
    if (int e = add_map(it->second, N, planes * kp, ph->bn, &ph->mapW)) return e;
    ph->part = part;
    ph->w_keep = w_keep;
    if (rows <= 8) {                                    // small batches: the phase streams weights as a GEMV (decode_mk.cu: mk_gemv)
      ph->gemv = 1; ph->splits = 1; ph->K = K; ph->a_planes = map_ptr[mapA]; ph->wb = it->second; ph->w32 = W;
    }
    *splits_out = ph->splits;
    return DIM_OK;
  }
};

// x-transformers Decoder layer order (a, c, f) x depth, final norm, to_logits, sampling: 12 phases per layer + 2.
int build_mk_plan(const S2SModel& m, const GenWs& w, int B, int Bc, int T, int steps, int samples, float temperature, int top_k,
                  const uint8_t* mask, const float* uniforms, float* logits_out, MkPlan& P) {
  const dim_s2s_config& c = m.cfg;
  const int inner = c.heads * c.dim_head, D = c.dim + c.dim_audio, F = c.ff_mult * D, V = c.num_tokens;
  const int planes = m.tc.planes;
  const bool kv16 = m.precision == DIM_PREC_BF16;
  memset(&P, 0, sizeof(P));
  MkBuilder b(P, m, B);
  int mapLn, mapAtt, mapFf;
  if (int e = b.add_map(w.ap, B, planes * D, 128, &mapLn)) return e;
  if (int e = b.add_map(w.mk_attp, B, planes * inner, 128, &mapAtt)) return e;
  if (int e = b.add_map(w.ap2, B, planes * F, 128, &mapFf)) return e;
  static const float w_keep_env = getenv("DIM_L2_WEIGHT_KEEP") ? (float)atof(getenv("DIM_L2_WEIGHT_KEEP")) : -1.f;
  const float w_keep = w_keep_env >= 0.f ? w_keep_env : (planes == 1 ? 0.7f : 0.f);
  const float scale = 1.0f / sqrtf((float)c.dim_head);
  auto resln = [&](int in_splits, const float* bias, const float* gain, const float* beta) -> int {
    MkPhase* ph = b.next(MK_ROW_RESLN);
    if (!ph) return fail(DIM_EINVAL, "decode plan: too many phases");
    ph->N = D; ph->part = w.mk_part; ph->in_splits = in_splits; ph->bias = bias; ph->x = w.x; ph->gain = gain; ph->beta = beta;
    ph->outp = w.ap; ph->out_kp = D;
    return DIM_OK;
  };
  for (int l = 0; l < c.depth; ++l) {
    const XtAttn& SA = m.self_attn[l];
    const XtAttn& CA = m.cross_attn[l];
    const XtFF& FF = m.ff[l];
    int sp = 1;
    // --- causal self attention with KV cache
    if (int e = b.gemm(mapLn, SA.wqkv, 3 * inner, D, w.mk_part, w_keep, &sp)) return e;
    {
      MkPhase* ph = b.next(MK_ATTN);
      if (!ph) return fail(DIM_EINVAL, "decode plan: too many phases");
      ph->part = w.mk_part; ph->q_splits = sp; ph->q_ld = 3 * inner; ph->q_col = 0; ph->k_col = inner; ph->v_col = 2 * inner;
      ph->append = 1; ph->Tk = 0; ph->kv_group = 1; ph->kcache = w.self_k[l]; ph->vcache = w.self_v[l];
      ph->kv_batch_stride = (size_t)(steps + 1) * inner; ph->kv_head_stride = (size_t)(steps + 1) * c.dim_head;
      ph->scale = scale; ph->outp = w.mk_attp; ph->out_kp = inner;
    }
    if (int e = b.gemm(mapAtt, SA.wo, D, inner, w.mk_part, w_keep, &sp)) return e;
    if (int e = resln(sp, nullptr, CA.norm_g, CA.norm_b)) return e;
    // --- cross attention over the once-projected context K/V
    if (int e = b.gemm(mapLn, CA.wq, inner, D, w.mk_part, w_keep, &sp)) return e;
    {
      MkPhase* ph = b.next(MK_ATTN);
      if (!ph) return fail(DIM_EINVAL, "decode plan: too many phases");
      ph->part = w.mk_part; ph->q_splits = sp; ph->q_ld = inner; ph->q_col = 0;
      ph->append = 0; ph->Tk = T; ph->kv_group = samples; ph->kcache = w.cross_kv[l];
      const size_t vplane = (size_t)Bc * T * inner;      // V block follows the K block (one per clip)
      ph->vcache = kv16 ? static_cast<void*>(reinterpret_cast<__nv_bfloat16*>(w.cross_kv[l]) + vplane)
                        : static_cast<void*>(w.cross_kv[l] + vplane);
      ph->kv_batch_stride = (size_t)T * inner; ph->kv_head_stride = (size_t)T * c.dim_head;
      ph->key_mask = mask; ph->key_valid = mask ? w.key_valid : nullptr; ph->scale = scale; ph->outp = w.mk_attp; ph->out_kp = inner;
    }
    if (int e = b.gemm(mapAtt, CA.wo, D, inner, w.mk_part, w_keep, &sp)) return e;
    if (int e = resln(sp, nullptr, FF.norm_g, FF.norm_b)) return e;
    // --- feed forward
    if (int e = b.gemm(mapLn, FF.w1, F, D, w.mk_part, w_keep, &sp)) return e;
    static const bool no_fuse = getenv("DIM_MK_NO_GELU_FUSE") != nullptr;     // measurement hook
    if (planes == 1 && cdiv(F, 64) >= 72 && !no_fuse) {        // a function of (N, K, planes) only, like every split decision
      // bf16 operands: FF1 has enough 64-wide output tiles to fill the SMs without splitting K, so bias + GELU + the bf16
      // rounding of FF2's A operand run in the GEMM's own epilogue and the GELU row phase (and its grid barrier) disappears
      MkPhase* ph = &P.phases[P.nphases - 1];
      ph->bn = 64; ph->splits = 1; ph->epi = 1; ph->bias = FF.b1; ph->outp = w.ap2; ph->out_kp = F;
      sp = 1;
      if (int e = tc_make_map(m.tc.wmap.find(FF.w1)->second, F, planes * tc_round_k(D), planes * tc_round_k(D), 64, &P.maps[ph->mapW])) return e;
    } else {
      MkPhase* ph = b.next(MK_ROW_GELU);
      if (!ph) return fail(DIM_EINVAL, "decode plan: too many phases");
      ph->N = F; ph->part = w.mk_part; ph->in_splits = sp; ph->bias = FF.b1; ph->outp = w.ap2; ph->out_kp = F;
    }
    if (int e = b.gemm(mapFf, FF.w2, D, F, w.mk_part, w_keep, &sp)) return e;
    const bool last = l + 1 == c.depth;
    if (int e = resln(sp, FF.b2, last ? m.final_g : m.self_attn[l + 1].norm_g, last ? m.final_b : m.self_attn[l + 1].norm_b)) return e;
  }
  {
    int sp = 1;
    if (int e = b.gemm(mapLn, m.logits_w, V, D, w.mk_part, w_keep, &sp)) return e;
    MkPhase* ph = b.next(MK_ROW_SAMPLE);
    if (!ph) return fail(DIM_EINVAL, "decode plan: too many phases");
    ph->N = V; ph->D = D; ph->part = w.mk_part; ph->in_splits = sp; ph->bias = m.logits_b; ph->emb = m.token_emb; ph->x = w.x;
    ph->pos = m.dec_pos_emb; ph->pos_scale = 1.0f / sqrtf((float)D);
    ph->gain = m.self_attn[0].norm_g; ph->beta = m.self_attn[0].norm_b; ph->outp = w.ap; ph->out_kp = D;
  }
  P.nmaps = b.nmaps;
  P.B = B; P.H = c.heads; P.planes = planes; P.kv_bf16 = kv16 ? 1 : 0; P.steps = steps;
  P.sc_floats = (std::max(T, steps + 1) + 3) / 4 * 4;
  {  // attention sub-group scratch (decode_mk.cu: MK_SUB_BYTES = 32 KB): a third ring slot when the scores fit beside it
    static const int st_env = getenv("DIM_MK_ATTN_STAGES") ? atoi(getenv("DIM_MK_ATTN_STAGES")) : 0;
    const size_t need3 = 3 * 64 * 144 + (size_t)P.sc_floats * 4 + 8 * 64 * 4 + 3 * 64 * 4 + 16;
    P.attn_stages = (need3 <= 32768 && st_env == 3) ? 3 : 2;     // measured: a third slot does not pay (341 vs 346 us of attention per step)
  }
  {  // measurement hooks: DIM_MK_ATTN_FFMA=1 keeps the FFMA attention items for bf16 caches; DIM_MK_NOPS=k appends k empty phases
    static const bool ffma = getenv("DIM_MK_ATTN_FFMA") != nullptr;
    static const int nops = getenv("DIM_MK_NOPS") ? atoi(getenv("DIM_MK_NOPS")) : 0;
    P.attn_mma = ffma ? 0 : 1;
    for (int i = 0; i < nops; ++i)
      if (!b.next(MK_NOP)) break;
  }
  {  // K/V prefetch windows: every non-attention phase serves the next attention phase of the program (cyclically), with a share of
     // the budget proportional to a rough duration weight
    static const double pf_mb = getenv("DIM_MK_PF_MB") ? atof(getenv("DIM_MK_PF_MB")) : 0.0;      // whole-GPU budget per attention phase
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    P.pf_budget = B > 8 ? (unsigned long long)(pf_mb * 1048576.0 / sms) : 0ull;
    const int n = P.nphases;
    auto weight = [&](const MkPhase& ph) -> float {
      if (ph.type == MK_GEMM) return (float)ph.N * (float)ph.kp >= 4.0e6f ? 2.f : 1.f;
      if (ph.type == MK_ROW_SAMPLE) return 3.f;
      return ph.type == MK_NOP ? 0.f : 1.f;
    };
    for (int i = 0; i < n; ++i) P.phases[i].pf_target = -1;
    for (int q = 0; q < n; ++q) {
      if (P.phases[q].type != MK_ATTN) continue;
      // window = the phases between the previous attention phase (cyclically) and q
      int first = q;
      float total = 0.f;
      for (int back = 1; back < n; ++back) {
        const int i = (q - back + n) % n;
        if (P.phases[i].type == MK_ATTN) break;
        first = i;
        total += weight(P.phases[i]);
      }
      float acc = 0.f;
      for (int i = first; i != q && total > 0.f; i = (i + 1) % n) {
        P.phases[i].pf_target = q;
        P.phases[i].pf_f0 = acc / total;
        acc += weight(P.phases[i]);
        P.phases[i].pf_f1 = acc / total;
      }
    }
  }
  {
    static const int dbg = getenv("DIM_MK_ATTN_DBG") ? atoi(getenv("DIM_MK_ATTN_DBG")) : 0;
    P.attn_dbg = dbg;
  }
  P.attn_nsub = P.attn_stages == 3 ? 6 : mk_attn_subgroups(P.kv_bf16, P.attn_mma, B <= 8 ? 1 : 0, std::max(T, steps + 1), B * c.heads);
  {
    static const int pre_rows = getenv("DIM_MK_ATTN_PRE") != nullptr ? std::max(0, atoi(getenv("DIM_MK_ATTN_PRE"))) : 0;      // measured: what the attention phase gains the barrier before it loses
    P.attn_pre = (P.kv_bf16 && P.attn_mma && B > 8) ? pre_rows : 0;
  }
  P.bar = w.mk_bar; P.trace = g_mk_trace_on ? w.mk_trace : nullptr;
  P.tokens = w.tokens; P.tok_stride = steps + 1;
  P.uniforms = uniforms; P.u_stride = steps;
  P.logits_out = logits_out; P.lo_stride = (long long)steps * V;
  P.temperature = temperature; P.top_k = top_k;
  return DIM_OK;
}

// Decode `B` clips on stream `s` (one group).  G: this group's cached step graph.
// samples > 1: every clip is decoded `samples` times (rows b*samples + j, their own uniforms) over ONE projection of its
// context: the cross-attention K/V of a clip are shared by its rows (the 10-sample best-of-N eval loop).
int generate_group(const S2SModel& m, S2SModel::StepGraph& G, const float* ctx, const uint8_t* mask, const int64_t* prompt,
                   int Bc, int T, int steps, float temperature, int top_k, const float* uniforms, int64_t* out_codes,
                   float* logits_out, void* ws, size_t ws_bytes, cudaStream_t s, int samples = 1) {
  const int B = Bc * samples;                        // decode rows
  const dim_s2s_config& c = m.cfg;
  const int inner = c.heads * c.dim_head, D = c.dim + c.dim_audio, F = c.ff_mult * D, V = c.num_tokens;
  const bool kv16 = m.precision == DIM_PREC_BF16;   // bf16 mode keeps both KV caches in bf16 (half the decode-attention bytes)
  GenWs w = carve_gen(c, m.tc.planes, kv16, Bc, T, steps, ws, samples);
  if (ws == nullptr || ws_bytes < w.bytes) return fail(DIM_EWORKSPACE, "dim_slmft_generate: workspace too small");
  const float scale = 1.0f / sqrtf((float)c.dim_head);

  for (int l = 0; l < c.depth; ++l) {  // cross-attention K/V of the whole context, once
    GemmArgs a;
    a.A = ctx; a.lda = D; a.W = m.cross_attn[l].wkv; a.M = Bc * T; a.N = 2 * inner; a.K = D;
    if (kv16) { a.Cb = reinterpret_cast<__nv_bfloat16*>(w.kv_tmp); a.ldcb = 2 * inner; }
    else { a.C = w.kv_tmp; a.ldc = 2 * inner; }
    if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    // head-major: each (clip, head) K / V block is one contiguous stream for the per-step attention (DRAM page locality)
    if (int e = launch_kv_head_major(w.kv_tmp, w.cross_kv[l], Bc, T, c.heads, kv16, s)) return e;
  }
  if (int e = launch_init_tokens(w.tokens, steps + 1, prompt, B, samples, s)) return e;      // tokens[r, 0] = prompt[r / samples]
  // the step graph reads the mask and the uniforms from the workspace, so a cached graph stays valid whatever buffers the
  // caller passes next time
  if (mask) DIM_CHECK_CUDA(cudaMemcpyAsync(w.mask_stage, mask, (size_t)Bc * T, cudaMemcpyDeviceToDevice, s));
  if (uniforms) DIM_CHECK_CUDA(cudaMemcpyAsync(w.u_stage, uniforms, (size_t)B * steps * sizeof(float), cudaMemcpyDeviceToDevice, s));
  const uint8_t* mask_g = mask ? w.mask_stage : nullptr;
  const float* uniforms_g = uniforms ? w.u_stage : nullptr;
  if (int e = launch_set_step(w.step, 0, s)) return e;

  const int max_keys = std::max(T, steps + 1);
  // <= 8 rows: the persistent kernel has a GEMV flavour of its GEMM phases (decode_mk.cu: mk_gemv), but measured SLOWER than the
  // per-kernel chain there (B = 1, bf16: 311 vs 249 us per step: 8 single-item attention phases at 10 us and 12 row phases at 3.5 us
  // pay a grid barrier each for work one CTA does) -- opt-in with DIM_SMALL_BATCH_MK=1 until its attention items split their keys
  static const bool small_mk = getenv("DIM_SMALL_BATCH_MK") != nullptr;
  if (g_decode_impl == 0 && m.tc.planes > 0 && (B > 8 || small_mk) && mk_supported(D, inner, F, V, c.heads, m.tc.planes, max_keys)) {
    // persistent decode kernel (more than 8 decode rows; fewer stay on the weight-streaming GEMV chain, which is faster there:
    // 77.9 vs 94 ms per single 300-frame clip, profiles/r02_notes.md): prompt embedding + layer 0's LayerNorm here, then every step inside ONE cooperative launch
    const XtAttn& SA0 = m.self_attn[0];
    if (int e = launch_embed_tokens(w.tokens, steps + 1, w.step, m.token_emb, w.x, B, D, V, s, m.dec_pos_emb, 1.0f / sqrtf((float)D))) return e;
    if (int e = launch_layer_norm(w.x, SA0.norm_g, SA0.norm_b, nullptr, nullptr, B, D, 1e-5f, s, w.ap, m.tc.planes, D)) return e;
    DIM_CHECK_CUDA(cudaMemsetAsync(w.mk_bar, 0, 256, s));
    if (g_mk_trace_on) DIM_CHECK_CUDA(cudaMemsetAsync(w.mk_trace, 0, MK_MAX_PHASES * sizeof(unsigned long long), s));
    if (mask) { if (int e = launch_mask_prefix(mask, Bc, T, w.key_valid, s)) return e; }
    static thread_local MkPlan plan;
    if (int e = build_mk_plan(m, w, B, Bc, T, steps, samples, temperature, top_k, mask, uniforms, logits_out, plan)) return e;
    if (int e = launch_decode_megakernel(plan, s)) return e;
    if (g_mk_trace_on) {
      g_mk_trace_host.assign(MK_MAX_PHASES, 0ull);
      g_mk_trace_types.assign(MK_MAX_PHASES, 0);
      for (int i = 0; i < plan.nphases; ++i) g_mk_trace_types[i] = plan.phases[i].type;
      DIM_CHECK_CUDA(cudaMemcpyAsync(g_mk_trace_host.data(), w.mk_trace, MK_MAX_PHASES * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    }
    DIM_CHECK_CUDA(cudaMemcpy2DAsync(out_codes, (size_t)steps * sizeof(int64_t), w.tokens + 1,
                                     (size_t)(steps + 1) * sizeof(int64_t), (size_t)steps * sizeof(int64_t), B,
                                     cudaMemcpyDeviceToDevice, s));
    return DIM_OK;
  }
  const bool tcp = tc_on(m.tc, B);
  const int P = m.tc.planes;
  // Head of the very first step: embedding of the prompt token and layer 0's self-attention LayerNorm.  Every later step gets
  // both from the tail kernel of the step before it (sample_next_kernel).
  auto enqueue_prologue = [&](cudaStream_t s) -> int {
    if (int e = launch_embed_tokens(w.tokens, steps + 1, w.step, m.token_emb, w.x, B, D, V, s, m.dec_pos_emb, 1.0f / sqrtf((float)D))) return e;
    const XtAttn& SA0 = m.self_attn[0];
    return launch_layer_norm(w.x, SA0.norm_g, SA0.norm_b, tcp ? nullptr : w.ln, nullptr, B, D, 1e-5f, s, tcp ? w.ap : nullptr, P, D);
  };
  auto enqueue_step = [&](int st, cudaStream_t s) -> int {
    for (int l = 0; l < c.depth; ++l) {
      const XtAttn& SA = m.self_attn[l];
      const XtAttn& CA = m.cross_attn[l];
      const XtFF& FF = m.ff[l];
      // --- causal self attention with KV cache (layer 0's LayerNorm was written by the previous step's tail / the prologue)
      if (l > 0)
        if (int e = launch_layer_norm(w.x, SA.norm_g, SA.norm_b, tcp ? nullptr : w.ln, nullptr, B, D, 1e-5f, s, tcp ? w.ap : nullptr, P, D))
          return e;
      {
        GemmArgs a;
        a.A = w.ln; a.lda = D; a.Ap = tcp ? w.ap : nullptr; a.W = SA.wqkv; a.C = w.qkv; a.ldc = 3 * inner; a.M = B; a.N = 3 * inner;
        a.K = D;
        if (int e = run_gemm(m.tc, a, w.ap, s, DIM_SPLIT_DECODE)) return e;
      }
      {
        DecodeAttnArgs a;
        a.q = w.qkv; a.ldq = 3 * inner; a.k = w.self_k[l]; a.v = w.self_v[l]; a.kv_bf16 = kv16;
        a.kv_batch_stride = (size_t)(steps + 1) * inner; a.kv_head_stride = (size_t)(steps + 1) * c.dim_head;
        a.kv_tok_stride = c.dim_head;
        a.k_new = w.qkv + inner; a.v_new = w.qkv + 2 * inner; a.ld_new = 3 * inner; a.append = 1; a.step = w.step;
        a.out = tcp ? nullptr : w.att; a.ldo = inner; a.out_p = tcp ? w.ap : nullptr; a.planes = P; a.kp = inner;
        a.B = B; a.H = c.heads; a.Tk = 0; a.scale = scale; a.prof_pos = st;
        if (int e = launch_attention_decode(a, max_keys, s)) return e;
      }
      {
        GemmArgs a;
        a.A = w.att; a.lda = inner; a.Ap = tcp ? w.ap : nullptr; a.W = SA.wo; a.residual = w.x; a.ldr = D; a.C = w.x; a.ldc = D;
        a.M = B; a.N = D; a.K = inner;
        if (int e = run_gemm(m.tc, a, w.ap, s, DIM_SPLIT_DECODE)) return e;
      }
      // --- cross attention over the cached context K/V
      if (int e = launch_layer_norm(w.x, CA.norm_g, CA.norm_b, tcp ? nullptr : w.ln, nullptr, B, D, 1e-5f, s, tcp ? w.ap : nullptr, P, D))
        return e;
      {
        GemmArgs a;
        a.A = w.ln; a.lda = D; a.Ap = tcp ? w.ap : nullptr; a.W = CA.wq; a.C = w.qkv; a.ldc = inner; a.M = B; a.N = inner; a.K = D;
        if (int e = run_gemm(m.tc, a, w.ap, s, DIM_SPLIT_DECODE)) return e;
      }
      {
        DecodeAttnArgs a;
        a.q = w.qkv; a.ldq = inner; a.kv_bf16 = kv16; a.k = w.cross_kv[l];
        const size_t vplane = (size_t)Bc * T * inner;    // V block follows the K block (one per clip)
        a.v = kv16 ? static_cast<void*>(reinterpret_cast<__nv_bfloat16*>(w.cross_kv[l]) + vplane) : static_cast<void*>(w.cross_kv[l] + vplane);
        a.kv_batch_stride = (size_t)T * inner; a.kv_head_stride = (size_t)T * c.dim_head; a.kv_tok_stride = c.dim_head;
        a.append = 0; a.step = w.step;
        a.key_mask = mask_g; a.out = tcp ? nullptr : w.att; a.ldo = inner; a.out_p = tcp ? w.ap : nullptr; a.planes = P;
        a.kp = inner; a.B = B; a.H = c.heads; a.Tk = T; a.scale = scale; a.kv_group = samples;
        if (int e = launch_attention_decode(a, max_keys, s)) return e;
      }
      {
        GemmArgs a;
        a.A = w.att; a.lda = inner; a.Ap = tcp ? w.ap : nullptr; a.W = CA.wo; a.residual = w.x; a.ldr = D; a.C = w.x; a.ldc = D;
        a.M = B; a.N = D; a.K = inner;
        if (int e = run_gemm(m.tc, a, w.ap, s, DIM_SPLIT_DECODE)) return e;
      }
      // --- feed forward
      if (int e = launch_layer_norm(w.x, FF.norm_g, FF.norm_b, tcp ? nullptr : w.ln, nullptr, B, D, 1e-5f, s, tcp ? w.ap : nullptr, P, D))
        return e;
      {
        GemmArgs a;
        a.A = w.ln; a.lda = D; a.Ap = tcp ? w.ap : nullptr; a.W = FF.w1; a.bias = FF.b1; a.M = B; a.N = F; a.K = D;
        a.act = DIM_ACT_GELU_ERF;
        if (tcp) { a.Cp = w.ap2; a.cp_planes = P; a.cp_kp = F; } else { a.C = w.ff; a.ldc = F; }
        if (int e = run_gemm(m.tc, a, w.ap, s, DIM_SPLIT_DECODE)) return e;
      }
      {
        GemmArgs a;
        a.A = w.ff; a.lda = F; a.Ap = tcp ? w.ap2 : nullptr; a.W = FF.w2; a.bias = FF.b2; a.residual = w.x; a.ldr = D; a.C = w.x;
        a.ldc = D; a.M = B; a.N = D; a.K = F;
        if (int e = run_gemm(m.tc, a, w.ap, s, DIM_SPLIT_DECODE)) return e;
      }
    }
    if (int e = launch_layer_norm(w.x, m.final_g, m.final_b, tcp ? nullptr : w.ln, nullptr, B, D, 1e-5f, s, tcp ? w.ap : nullptr, P, D))
      return e;
    {
      GemmArgs a;
      a.A = w.ln; a.lda = D; a.Ap = tcp ? w.ap : nullptr; a.W = m.logits_w; a.bias = m.logits_b; a.C = w.logits; a.ldc = V; a.M = B; a.N = V; a.K = D;
      if (int e = run_gemm(m.tc, a, w.ap, s, DIM_SPLIT_DECODE)) return e;
    }
    // tail: sample this step's token, advance the step counter, and produce the next step's embedding + layer-0 LayerNorm
    const XtAttn& SA0 = m.self_attn[0];
    if (int e = launch_sample_next(w.logits, B, V, temperature, top_k, uniforms_g, steps, w.step,
                                   reinterpret_cast<unsigned int*>(w.step + 16), w.tokens, steps + 1, 1, logits_out, steps * V,
                                   m.token_emb, w.x, D, SA0.norm_g, SA0.norm_b, tcp ? nullptr : w.ln, tcp ? w.ap : nullptr, P, 1e-5f, s, m.dec_pos_emb,
                                   1.0f / sqrtf((float)D)))
      return e;
    return DIM_OK;
  };

  // One decode step = 46 short, strictly dependent launches: replay it as a CUDA graph (captured once per distinct call
  // signature) to remove the per-launch submission cost; DIM_NO_GRAPH=1 or profiling falls back to plain launches.
  static const bool no_graph = getenv("DIM_NO_GRAPH") != nullptr;
  if (no_graph || g_prof_on) {
    if (int e = enqueue_prologue(s)) return e;
    for (int st = 0; st < steps; ++st)
      if (int e = enqueue_step(st, s)) return e;
  } else {
    std::vector<uintptr_t> key = {(uintptr_t)(mask != nullptr), (uintptr_t)(uniforms != nullptr), (uintptr_t)logits_out, (uintptr_t)ws,
                                  (uintptr_t)B, (uintptr_t)samples, (uintptr_t)T, (uintptr_t)steps, (uintptr_t)top_k,
                                  (uintptr_t)(temperature * 65536.0f)};
    if (!G.stream) {
      DIM_CHECK_CUDA(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
      DIM_CHECK_CUDA(cudaEventCreateWithFlags(&G.fork, cudaEventDisableTiming));
      DIM_CHECK_CUDA(cudaEventCreateWithFlags(&G.join, cudaEventDisableTiming));
    }
    // steps per graph launch (DIM_GRAPH_UNROLL, default 1).  Measured: 1, 13, 23 and 299 steps per graph all give 260.5-261.3 ms
    // per bench step -- the gap between consecutive graph launches is not what limits the decode (profiles/r01_notes.md).
    static const int unroll = getenv("DIM_GRAPH_UNROLL") ? std::max(1, atoi(getenv("DIM_GRAPH_UNROLL"))) : 1;
    if (G.exec == nullptr || G.key != key) {
      if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
      if (G.exec_u) { cudaGraphExecDestroy(G.exec_u); G.exec_u = nullptr; }
      if (int e = enqueue_prologue(s)) return e;
      if (int e = enqueue_step(0, s)) return e;                       // warm (lazy attribute setup, tensor maps) outside capture
      if (int e = launch_set_step(w.step, 0, s)) return e;            // ... and rewind the step counter it advanced
      DIM_CHECK_CUDA(cudaStreamSynchronize(s));
      for (int pass = 0; pass < 2; ++pass) {
        const int n = pass == 0 ? 1 : unroll;
        if (pass == 1 && (unroll <= 1 || steps < unroll)) break;
        cudaGraph_t graph = nullptr;
        DIM_CHECK_CUDA(cudaStreamBeginCapture(G.stream, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        for (int i = 0; i < n && rc == 0; ++i) rc = enqueue_step(i, G.stream);
        cudaError_t ce = cudaStreamEndCapture(G.stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        DIM_CHECK_CUDA(ce);
        DIM_CHECK_CUDA(cudaGraphInstantiate(pass == 0 ? &G.exec : &G.exec_u, graph, 0));
        cudaGraphDestroy(graph);
      }
      G.key = key;
    }
    DIM_CHECK_CUDA(cudaEventRecord(G.fork, s));
    DIM_CHECK_CUDA(cudaStreamWaitEvent(G.stream, G.fork, 0));
    if (int e = enqueue_prologue(G.stream)) return e;
    int st = 0;
    if (G.exec_u)
      for (; st + unroll <= steps; st += unroll) DIM_CHECK_CUDA(cudaGraphLaunch(G.exec_u, G.stream));
    for (; st < steps; ++st) DIM_CHECK_CUDA(cudaGraphLaunch(G.exec, G.stream));
    g_launches.fetch_add((uint64_t)steps * (uint64_t)(2 + c.depth * 11), std::memory_order_relaxed);
    DIM_CHECK_CUDA(cudaEventRecord(G.join, G.stream));
    DIM_CHECK_CUDA(cudaStreamWaitEvent(s, G.join, 0));
  }
  DIM_CHECK_CUDA(cudaMemcpy2DAsync(out_codes, (size_t)steps * sizeof(int64_t), w.tokens + 1,
                                   (size_t)(steps + 1) * sizeof(int64_t), (size_t)steps * sizeof(int64_t), B,
                                   cudaMemcpyDeviceToDevice, s));
  return DIM_OK;
}
}  // namespace

extern "C" int dim_slmft_generate(dim_handle_t h, int model, const float* ctx, const uint8_t* mask, const int64_t* prompt,
                                  int B, int T, int steps, float temperature, int top_k, const float* uniforms,
                                  int64_t* out_codes, float* logits_out, void* ws, size_t ws_bytes, void* stream) {
  DIM_REQUIRE(h && model >= 0 && model < (int)h->s2s.size(), "dim_slmft_generate: bad model");
  DIM_CHECK_CUDA(cudaSetDevice(h->device));            // the handle's device, whatever the caller's current device is
  DIM_REQUIRE(ctx && prompt && out_codes && B > 0 && T > 0 && steps > 0, "dim_slmft_generate: bad argument");
  DIM_REQUIRE(temperature >= 0.f, "temperature must be >= 0");
  DIM_REQUIRE(temperature == 0.f || (uniforms && top_k > 0), "sampling needs uniforms and top_k");
  const S2SModel& m = *h->s2s[model];
  DIM_REQUIRE(h->s2s[model]->dec_pos_emb == nullptr || steps + 1 <= h->s2s[model]->cfg.max_seq_len, "generate: longer than the decoder's positional table");
  const dim_s2s_config& c = m.cfg;
  const int D = c.dim + c.dim_audio, V = c.num_tokens;
  cudaStream_t s = as_stream(stream);
  int begin[kMaxGroups + 1];
  const int ng = plan_groups(m, B, std::max(T, steps + 1), begin);
  if ((int)m.graphs.size() < kMaxGroups + 1) m.graphs.resize(kMaxGroups + 1);     // last slot: multi-sample decoding
  if (ng == 1 || g_prof_on) {
    if (ng > 1) {   // profiling: same groups, sequentially on the caller's stream (events need one stream)
      char* wsp = static_cast<char*>(ws);
      for (int g = 0; g < ng; ++g) {
        const int b0 = begin[g], bg = begin[g + 1] - b0;
        const size_t need = carve_gen(c, m.tc.planes, m.precision == DIM_PREC_BF16, bg, T, steps, nullptr).bytes;
        if ((size_t)(wsp - static_cast<char*>(ws)) + need > ws_bytes) return fail(DIM_EWORKSPACE, "dim_slmft_generate: workspace too small");
        if (int e = generate_group(m, m.graphs[g], ctx + (size_t)b0 * T * D, mask ? mask + (size_t)b0 * T : nullptr, prompt + b0, bg, T,
                                   steps, temperature, top_k, uniforms ? uniforms + (size_t)b0 * steps : nullptr,
                                   out_codes + (size_t)b0 * steps, logits_out ? logits_out + (size_t)b0 * steps * V : nullptr, wsp,
                                   need, s))
          return e;
        wsp += need;
      }
      return DIM_OK;
    }
    return generate_group(m, m.graphs[0], ctx, mask, prompt, B, T, steps, temperature, top_k, uniforms, out_codes, logits_out, ws,
                          ws_bytes, s);
  }
  // fork: every group decodes on its own side stream, ordered after the caller's stream; join at the end
  if (!m.fork_ev) DIM_CHECK_CUDA(cudaEventCreateWithFlags(&m.fork_ev, cudaEventDisableTiming));
  DIM_CHECK_CUDA(cudaEventRecord(m.fork_ev, s));
  char* wsp = static_cast<char*>(ws);
  for (int g = 0; g < ng; ++g) {
    S2SModel::StepGraph& G = m.graphs[g];
    if (!G.stream) {
      DIM_CHECK_CUDA(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
      DIM_CHECK_CUDA(cudaEventCreateWithFlags(&G.fork, cudaEventDisableTiming));
      DIM_CHECK_CUDA(cudaEventCreateWithFlags(&G.join, cudaEventDisableTiming));
    }
    const int b0 = begin[g], bg = begin[g + 1] - b0;
    const size_t need = carve_gen(c, m.tc.planes, m.precision == DIM_PREC_BF16, bg, T, steps, nullptr).bytes;
    if ((size_t)(wsp - static_cast<char*>(ws)) + need > ws_bytes) return fail(DIM_EWORKSPACE, "dim_slmft_generate: workspace too small");
    DIM_CHECK_CUDA(cudaStreamWaitEvent(G.stream, m.fork_ev, 0));
    if (int e = generate_group(m, G, ctx + (size_t)b0 * T * D, mask ? mask + (size_t)b0 * T : nullptr, prompt + b0, bg, T, steps,
                               temperature, top_k, uniforms ? uniforms + (size_t)b0 * steps : nullptr,
                               out_codes + (size_t)b0 * steps, logits_out ? logits_out + (size_t)b0 * steps * V : nullptr, wsp, need,
                               G.stream))
      return e;
    DIM_CHECK_CUDA(cudaEventRecord(G.join, G.stream));
    DIM_CHECK_CUDA(cudaStreamWaitEvent(s, G.join, 0));
    wsp += need;
  }
  return DIM_OK;
}

// ---- several samples per clip over one context projection (evaluate_test_epoch's best-of-N loop) -----------
extern "C" size_t dim_slmft_samples_workspace_bytes(dim_handle_t h, int model, int B, int T, int steps, int samples) {
  if (!h || model < 0 || model >= (int)h->s2s.size() || B <= 0 || T <= 0 || steps <= 0 || samples <= 0) return 0;
  const S2SModel& m = *h->s2s[model];
  return carve_gen(m.cfg, m.tc.planes, m.precision == DIM_PREC_BF16, B, T, steps, nullptr, samples).bytes;
}

extern "C" int dim_slmft_generate_samples(dim_handle_t h, int model, const float* ctx, const uint8_t* mask, const int64_t* prompt,
                                          int B, int T, int steps, int samples, float temperature, int top_k,
                                          const float* uniforms, int64_t* out_codes, float* logits_out, void* ws,
                                          size_t ws_bytes, void* stream) {
  DIM_REQUIRE(h && model >= 0 && model < (int)h->s2s.size(), "dim_slmft_generate_samples: bad model");
  DIM_CHECK_CUDA(cudaSetDevice(h->device));            // the handle's device, whatever the caller's current device is
  DIM_REQUIRE(ctx && prompt && out_codes && B > 0 && T > 0 && steps > 0 && samples > 0, "dim_slmft_generate_samples: bad argument");
  DIM_REQUIRE(temperature >= 0.f, "temperature must be >= 0");
  DIM_REQUIRE(temperature == 0.f || (uniforms && top_k > 0), "sampling needs uniforms and top_k");
  const S2SModel& m = *h->s2s[model];
  DIM_REQUIRE(h->s2s[model]->dec_pos_emb == nullptr || steps + 1 <= h->s2s[model]->cfg.max_seq_len, "generate: longer than the decoder's positional table");
  if ((int)m.graphs.size() < kMaxGroups + 1) m.graphs.resize(kMaxGroups + 1);
  return generate_group(m, m.graphs[kMaxGroups], ctx, mask, prompt, B, T, steps, temperature, top_k, uniforms, out_codes, logits_out,
                        ws, ws_bytes, as_stream(stream), samples);
}

// ---- teacher-forced decoder forward: AutoregressiveWrapper.forward(..., return_outputs=True) ----------------
// seq2seq_pretrain.py:447-448 -> x-transformers TransformerWrapper over the WHOLE input sequence: token embedding, then per layer
// causal self attention (optionally with the random `self_attn_kv_mask` of mask_prob, passed in as kv_mask), cross attention
// over the context under the key-padding mask, feed forward; final norm; logits.  Forward only (no autograd).
namespace {
struct TfWs {
  float *x, *ln, *qkv, *att, *ff, *ckv;
  __nv_bfloat16 *ap, *ap2;
  size_t bytes;
};
TfWs carve_tf(const dim_s2s_config& c, int planes, int B, int T, int L, void* base) {
  const int inner = c.heads * c.dim_head, D = c.dim + c.dim_audio, F = c.ff_mult * D;
  const size_t R = (size_t)B * L, RC = (size_t)B * T;
  char* p = static_cast<char*>(base);
  TfWs w{};
  auto take = [&](size_t nfloat) {
    float* r = reinterpret_cast<float*>(p);
    p += align_up(nfloat * sizeof(float), 256);
    return r;
  };
  w.x = take(R * D); w.ln = take(R * D); w.qkv = take(R * 3 * inner); w.att = take(R * inner); w.ff = take(R * F);
  w.ckv = take(RC * 2 * inner);
  const size_t need = std::max(R * (size_t)tc_round_k(F), RC * (size_t)tc_round_k(D)) * planes;
  w.ap = planes ? reinterpret_cast<__nv_bfloat16*>(take(need / 2 + 64)) : nullptr;
  w.ap2 = planes ? reinterpret_cast<__nv_bfloat16*>(take(R * (size_t)tc_round_k(F) * planes / 2 + 64)) : nullptr;
  w.bytes = (size_t)(p - static_cast<char*>(base));
  return w;
}
}  // namespace

extern "C" size_t dim_slmft_teacher_forced_workspace_bytes(dim_handle_t h, int model, int B, int T, int L) {
  if (!h || model < 0 || model >= (int)h->s2s.size() || B <= 0 || T <= 0 || L <= 0) return 0;
  return carve_tf(h->s2s[model]->cfg, h->s2s[model]->tc.planes, B, T, L, nullptr).bytes;
}

extern "C" int dim_slmft_teacher_forced(dim_handle_t h, int model, const float* ctx, const uint8_t* mask, const int64_t* tokens,
                                        const uint8_t* kv_mask, int B, int T, int L, float* logits, void* ws, size_t ws_bytes,
                                        void* stream) {
  DIM_REQUIRE(h && model >= 0 && model < (int)h->s2s.size(), "dim_slmft_teacher_forced: bad model");
  DIM_CHECK_CUDA(cudaSetDevice(h->device));            // the handle's device, whatever the caller's current device is
  DIM_REQUIRE(ctx && tokens && logits && B > 0 && T > 0 && L > 0, "dim_slmft_teacher_forced: bad argument");
  const S2SModel& m = *h->s2s[model];
  const dim_s2s_config& c = m.cfg;
  const int inner = c.heads * c.dim_head, D = c.dim + c.dim_audio, F = c.ff_mult * D, V = c.num_tokens;
  TfWs w = carve_tf(c, m.tc.planes, B, T, L, ws);
  if (ws == nullptr || ws_bytes < w.bytes) return fail(DIM_EWORKSPACE, "dim_slmft_teacher_forced: workspace too small");
  cudaStream_t s = as_stream(stream);
  const int R = B * L;
  const bool tcp = tc_on(m.tc, R);
  const int P = m.tc.planes;
  const float scale = 1.0f / sqrtf((float)c.dim_head);
  // token embedding (TransformerWrapper; SLMFT has no positional embedding in the decoder, seq2seq_pretrain.py:386)
  if (int e = launch_vq_gather(tokens, m.token_emb, w.x, R, D, V, nullptr, s)) return e;
  if (m.dec_pos_emb) {                                  // SLM flavour: + pos_emb[t] * D^-0.5 (x-transformers AbsolutePositionalEmbedding)
    DIM_REQUIRE(L <= c.max_seq_len, "sequence longer than the decoder's positional table");
    if (int e = launch_add_pos_table(w.x, m.dec_pos_emb, 1.0f / sqrtf((float)D), B, L, D, s)) return e;
  }
  for (int l = 0; l < c.depth; ++l) {
    const XtAttn& SA = m.self_attn[l];
    const XtAttn& CA = m.cross_attn[l];
    const XtFF& FF = m.ff[l];
    // --- causal self attention over the whole sequence (+ optional key mask)
    if (int e = launch_layer_norm(w.x, SA.norm_g, SA.norm_b, tcp ? nullptr : w.ln, nullptr, R, D, 1e-5f, s, tcp ? w.ap : nullptr, P, D))
      return e;
    {
      GemmArgs a;
      a.A = w.ln; a.lda = D; a.Ap = tcp ? w.ap : nullptr; a.W = SA.wqkv; a.C = w.qkv; a.ldc = 3 * inner; a.M = R; a.N = 3 * inner; a.K = D;
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
    {
      AttnArgs a;
      a.q = w.qkv; a.k = w.qkv + inner; a.v = w.qkv + 2 * inner; a.ldq = a.ldk = a.ldv = 3 * inner;
      a.out = tcp ? nullptr : w.att; a.ldo = inner; a.out_p = tcp ? w.ap : nullptr; a.planes = P; a.kp = inner;
      a.key_mask = kv_mask; a.B = B; a.H = c.heads; a.Tq = L; a.Tk = L; a.Dh = c.dim_head; a.scale = scale; a.causal = 1;
      if (int e = launch_attention_prefill(a, s)) return e;
    }
    {
      GemmArgs a;
      a.A = w.att; a.lda = inner; a.Ap = tcp ? w.ap : nullptr; a.W = SA.wo; a.residual = w.x; a.ldr = D; a.C = w.x; a.ldc = D;
      a.M = R; a.N = D; a.K = inner;
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
    // --- cross attention over the context (K/V of this layer projected here; key-padding mask on the context)
    {
      GemmArgs a;
      a.A = ctx; a.lda = D; a.W = CA.wkv; a.C = w.ckv; a.ldc = 2 * inner; a.M = B * T; a.N = 2 * inner; a.K = D;
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
    if (int e = launch_layer_norm(w.x, CA.norm_g, CA.norm_b, tcp ? nullptr : w.ln, nullptr, R, D, 1e-5f, s, tcp ? w.ap : nullptr, P, D))
      return e;
    {
      GemmArgs a;
      a.A = w.ln; a.lda = D; a.Ap = tcp ? w.ap : nullptr; a.W = CA.wq; a.C = w.qkv; a.ldc = inner; a.M = R; a.N = inner; a.K = D;
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
    {
      AttnArgs a;
      a.q = w.qkv; a.ldq = inner; a.k = w.ckv; a.v = w.ckv + inner; a.ldk = a.ldv = 2 * inner;
      a.out = tcp ? nullptr : w.att; a.ldo = inner; a.out_p = tcp ? w.ap : nullptr; a.planes = P; a.kp = inner;
      a.key_mask = mask; a.B = B; a.H = c.heads; a.Tq = L; a.Tk = T; a.Dh = c.dim_head; a.scale = scale; a.causal = 0;
      if (int e = launch_attention_prefill(a, s)) return e;
    }
    {
      GemmArgs a;
      a.A = w.att; a.lda = inner; a.Ap = tcp ? w.ap : nullptr; a.W = CA.wo; a.residual = w.x; a.ldr = D; a.C = w.x; a.ldc = D;
      a.M = R; a.N = D; a.K = inner;
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
    // --- feed forward
    if (int e = launch_layer_norm(w.x, FF.norm_g, FF.norm_b, tcp ? nullptr : w.ln, nullptr, R, D, 1e-5f, s, tcp ? w.ap : nullptr, P, D))
      return e;
    {
      GemmArgs a;
      a.A = w.ln; a.lda = D; a.Ap = tcp ? w.ap : nullptr; a.W = FF.w1; a.bias = FF.b1; a.M = R; a.N = F; a.K = D;
      a.act = DIM_ACT_GELU_ERF;
      if (tcp) { a.Cp = w.ap2; a.cp_planes = P; a.cp_kp = F; } else { a.C = w.ff; a.ldc = F; }
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
    {
      GemmArgs a;
      a.A = w.ff; a.lda = F; a.Ap = tcp ? w.ap2 : nullptr; a.W = FF.w2; a.bias = FF.b2; a.residual = w.x; a.ldr = D; a.C = w.x;
      a.ldc = D; a.M = R; a.N = D; a.K = F;
      if (int e = run_gemm(m.tc, a, w.ap, s)) return e;
    }
  }
  if (int e = launch_layer_norm(w.x, m.final_g, m.final_b, tcp ? nullptr : w.ln, nullptr, R, D, 1e-5f, s, tcp ? w.ap : nullptr, P, D))
    return e;
  GemmArgs a;
  a.A = w.ln; a.lda = D; a.Ap = tcp ? w.ap : nullptr; a.W = m.logits_w; a.bias = m.logits_b; a.C = logits; a.ldc = V; a.M = R; a.N = V; a.K = D;
  return run_gemm(m.tc, a, w.ap, s);
}

// ---- persistent decode kernel: phase trace and implementation switch ----------------------------------------------------------
extern "C" int dim_decode_trace_enable(int on) {
  g_mk_trace_on = on != 0;
  return DIM_OK;
}

extern "C" int dim_decode_trace_collect(double* ms_per_phase, int32_t* type_per_phase, int max_phases, int* n_out) {
  DIM_REQUIRE(ms_per_phase && type_per_phase && n_out && max_phases > 0, "dim_decode_trace_collect: bad argument");
  if (cudaDeviceSynchronize() != cudaSuccess) return fail(DIM_ECUDA, "decode trace: device synchronize failed");
  int n = 0;
  for (size_t i = 0; i < g_mk_trace_types.size() && n < max_phases; ++i) {
    if (g_mk_trace_types[i] == 0) break;
    ms_per_phase[n] = (double)g_mk_trace_host[i] * 1e-6;
    type_per_phase[n] = g_mk_trace_types[i];
    ++n;
  }
  *n_out = n;
  return DIM_OK;
}

extern "C" int dim_decode_set_impl(int impl) {
  DIM_REQUIRE(impl == 0 || impl == 1, "dim_decode_set_impl: 0 = persistent kernel, 1 = per-kernel graph path");
  g_decode_impl = impl;
  return DIM_OK;
}
