/*
 * dimb200.h -- C ABI of libdimb200.so: the B200 (sm_100a) kernels behind the DIM inference hot path.
 *
 * The reference (Boese0601/Dyadic-Interaction-Modeling) is pure PyTorch: it has no FFI, its "operators" are ATen
 * calls inside Python modules.  Each entry point below therefore cites the reference Python lines whose arithmetic it
 * replaces; INTEGRATION.md shows the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with `h_`; plain sizes, no torch types;
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*); the caller owns every buffer,
 *     including workspaces (query the size first); weights registered with dim_set_tensor are BORROWED and must
 *     outlive the handle; derived (re-packed) weights are owned by the handle;
 *   - return value: 0 = OK, otherwise a DIM_E* code; dim_last_error() returns a thread-local message;
 *   - no CPU fallback: every entry point fails with DIM_ENODEVICE when no sm_100 device is usable;
 *   - a handle is not thread-safe; one process per GPU.
 *   - row-major everywhere; "frames" (B,T,C) means C fastest.
 */
#ifndef DIMB200_H
#define DIMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIM_OK 0
#define DIM_EINVAL 1      /* bad argument / shape */
#define DIM_ENODEVICE 2   /* no usable sm_100 GPU */
#define DIM_ECUDA 3       /* CUDA runtime error (message has the detail) */
#define DIM_EMISSING 4    /* a required weight was not registered */
#define DIM_EWORKSPACE 5  /* workspace too small */

/* activation selectors for fused GEMM epilogues */
#define DIM_ACT_NONE 0
#define DIM_ACT_LEAKY 1      /* LeakyReLU(slope)             stage1_BIWI.py:262,267 */
#define DIM_ACT_GELU_TANH 2  /* tanh-approximated GELU       utils/base_model_util.py:81-94 */
#define DIM_ACT_GELU_ERF 3   /* exact GELU (nn.GELU())       x-transformers FeedForward */

/* arithmetic modes of a built model */
#define DIM_PREC_FP32 0   /* fp32 storage + fp32 FFMA accumulation: the parity mode (<=1e-4 vs the reference) */
#define DIM_PREC_BF16 1   /* GEMM operands rounded to bf16, fp32 accumulate on tcgen05; fp32 softmax/LayerNorm/residual */
#define DIM_PREC_FP32_TC 2 /* fp32-accurate GEMMs on tcgen05: operands split exactly into 3 bf16 planes, 6 products summed in
                              fp32 (TMEM).  Same parity bar as DIM_PREC_FP32 at tensor-core speed. */

const char* dim_last_error(void);
int dim_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t dim_launch_count(void);

/* Per-kernel-category timing for bench.py's roofline: when enabled, every launch is bracketed by CUDA events on its own
 * stream; dim_profile_collect synchronises the device, sums (launches, ms, algorithmic bytes, algorithmic flops) per
 * category since the last collect and clears the log.  Enabling it serialises nothing but adds event overhead: never
 * leave it on inside a throughput measurement. */
typedef struct {
  int32_t category;
  int32_t launches;
  double ms, bytes, flops;
} dim_prof_entry;
int dim_profile_enable(int on);
int dim_profile_collect(dim_prof_entry* out, int max_entries, int* n_out);
const char* dim_profile_category_name(int category);

/* ------------------------------------------------------------------------------------------------------------
 * Operator level (one kernel each).  These are what the unit parity tests call.
 * ---------------------------------------------------------------------------------------------------------- */

/* Codebook nearest neighbour.  Replaces models/lib/quantizer.py:38-45 (d = |z|^2 + |e|^2 - 2 z E^T ; argmin, first
 * minimum wins).  z [N,D] fp32, codebook [K,D] fp32, idx [N] int64.  D % 4 == 0, D <= 256, K <= 4096. */
int dim_vq_argmin(const float* z, const float* codebook, int64_t* idx, int N, int D, int K, void* stream);

/* Codebook row gather.  Replaces the one-hot (N,K)x(K,D) matmul of quantizer.py:79-90 and
 * seq2seq_pretrain.py:457-461.  idx [N] int64 -> out [N,D] fp32 (rows).  Returns DIM_EINVAL semantics on the device
 * are not possible; out-of-range indices are clamped into [0,K) and counted in *bad (nullable, device int32). */
int dim_vq_gather(const int64_t* idx, const float* codebook, float* out, int N, int D, int K, int32_t* bad,
                  void* stream);

/* C[M,N] = act(A[M,K] @ W[N,K]^T + bias[N]) + residual[M,N]     (nn.Linear semantics: base_models.py:48-50,118-119)
 * fp32 in/out, fp32 FFMA accumulation.  bias/residual nullable.  K % 4 == 0, N % 4 == 0, lda/ldc in elements. */
int dim_linear_f32(const float* A, int lda, const float* W, const float* bias, const float* residual, int ldr,
                   float* C, int ldc, int M, int N, int K, int act, float slope, void* stream);

/* Tensor-core (tcgen05) path of the same Linear.  Operands are bf16 "plane" matrices [rows, planes*Kp], Kp = K rounded up
 * to 64 and zero padded: planes == 1 is plain bf16; planes == 2 / 3 hold the exact bf16 split x = h + m (+ l) of fp32
 * values, and the GEMM accumulates the 3 / 6 significant plane-pair products in fp32 (TMEM), i.e. an fp32-accurate
 * Linear at tensor-core speed.  dim_split_bf16_planes produces a plane matrix from fp32 rows. */
int dim_split_bf16_planes(const float* X, int ldx, int rows, int K, int planes, void* out_bf16, void* stream);
int dim_linear_bf16_planes(const void* A_planes, const void* W_planes, int K, int planes, const float* bias,
                           const float* residual, int ldr, float* C, int ldc, int M, int N, int act, float slope,
                           void* stream);

/* Conv1d(C,C,k=5,stride 1,padding 2 replicate) + bias + LeakyReLU(slope) on frames (B,T,C) -> (B,T,C).
 * Replaces stage1_BIWI.py:265-267 / :331-333.  Wr is the weight re-laid out as [Cout][5][Cin] (dim_repack_conv_weight).
 * lens [B] int32 (nullable): frames >= lens[b] are treated as absent (replicate padding happens at lens[b]-1). */
int dim_conv5_leaky_f32(const float* x, const float* Wr, const float* bias, const int32_t* lens, float* y, int B, int T,
                        int C, float slope, void* stream);
int dim_repack_conv_weight(const float* w_oik /*[Cout][Cin][5]*/, float* w_oki /*[Cout][5][Cin]*/, int Cout, int Cin,
                           void* stream);

/* InstanceNorm1d(affine=False, eps) over time, in place on frames (B,T,C): per (b,c) mean and biased variance over
 * t < lens[b] (or T).  Replaces stage1_BIWI.py:268 / :334. */
int dim_instance_norm_f32(float* x, const int32_t* lens, int B, int T, int C, float eps, void* stream);

/* Linear with NO alignment requirement on K, N, lda, ldw or ldc: C[M,N] = act(A[M,K] @ W[N,K]^T + bias).  For the 70110-wide
 * mesh layers of EmocaConverter / SpeakerSLMFT (nn.Linear(70110, 56) + LeakyReLU: seq2seq_pretrain.py:777, applied at :712;
 * nn.Linear(768, 70110): :803-807, applied at :658).  A wide-K, narrow-N problem is split over K into a workspace
 * (dim_linear_ragged_workspace_bytes; 0 = none needed) and reduced in a fixed order. */
size_t dim_linear_ragged_workspace_bytes(int M, int N, int K);
int dim_linear_ragged_f32(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M, int N,
                          int K, int act, float slope, void* ws, size_t ws_bytes, void* stream);

/* One nn.LSTM layer, batch_first, zero initial state, fp32: x (B,T,in_dim) -> out (B,T,ndir*H), forward direction in columns
 * [0,H), reverse (when w_ih_r != NULL) in [H,2H).  Weights in torch's layout: w_ih [4H,in_dim], w_hh [4H,H], b_* [4H], gate
 * order i,f,g,o.  Replaces EmocaConverter.vertice_map_reverse_lstm (seq2seq_pretrain.py:789-802; calls :657, :823): call once
 * per layer, feeding layer k's out to layer k+1.  in_dim % 4 == 0, H % 64 == 0.  ws: dim_lstm_layer_workspace_bytes. */
size_t dim_lstm_layer_workspace_bytes(int B, int T, int in_dim, int H, int ndir);
int dim_lstm_layer_f32(const float* x, int in_dim, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                       const float* w_ih_r, const float* w_hh_r, const float* b_ih_r, const float* b_hh_r, int B, int T, int H,
                       float* out, void* ws, size_t ws_bytes, void* stream);

/* Audio-feature resampling to the motion frame rate, (t,d) fp32 -> (new_t,d) fp32, d % 4 == 0.
 * mode 0: window mean, out[i] = mean(in[i*window : (i+1)*window])  -- vico_preprocessing.downsample_mean
 *         (code/vico_preprocessing.py:7-19: new_t = int(t*0.6), window = int(t/new_t), i.e. 1 for the 50->30 fps case);
 * mode 1: linear interpolation with align_corners=True -- dataset/l2l.downsample_mean (code/dataset/l2l.py:23-29). */
int dim_resample_features(const float* in, int t, int d, int new_t, int window, int mode, float* out, void* stream);

/* The loaders' collate step on the device (dataset/data_loader.py:138-152 ViCoDataset.__getitem__, :229-245
 * LmListenerDataset.__getitem__, :429-439 pad_collate, x_engine_pt.py:246-249 mask): clip b owns rows [offsets[b], offsets[b+1]) of the
 * packed (sum_len, .) device arrays; src (B,T,motion_dim+audio_dim) = speaker | audio, tgt (B,T,motion_dim) = listener, zero past
 * each clip's length, mask (B,T) uint8 = t < len (nullable).  speaker == NULL writes ones (the ViCo loader's torch.ones_like),
 * audio == NULL writes zeros (the LM-Listener loader's torch.zeros).  offsets: (B+1) int64 on the device. */
int dim_assemble_batch(const float* speaker, const float* audio, const float* listener, const int64_t* offsets, int B, int T,
                       int motion_dim, int audio_dim, float* src, float* tgt, uint8_t* mask, void* stream);

/* LayerNorm over the last dim (eps), gain, optional bias.  base_models.py:14 ; x-transformers bias-free LayerNorm. */
int dim_layer_norm_f32(const float* x, const float* gain, const float* bias, float* y, int rows, int dim, float eps,
                       void* stream);

/* Multi-head attention over frames, fp32.  q/k/v point at element (b=0,t=0,h=0,d=0) of tensors whose (b,t) rows are
 * ld* elements apart and whose head h occupies columns [h*Dh, (h+1)*Dh).  out (B,Tq,H*Dh).
 *   scores = (q.k) * scale ; masked scores are filled with -FLT_MAX (x-transformers) ; softmax fp32 ; out = P v.
 *   key_mask (B,Tk) uint8 nullable (1 = keep) ; lens (B) nullable (keys >= lens[b] masked) ; causal: key j > query i masked.
 * Replaces base_models.py:136-143 (VQ-VAE, Dh=48, scale=hidden**-0.5, no mask) and x-transformers Attend (Dh=64). */
int dim_attention_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out, int ldo,
                      const uint8_t* key_mask, const int32_t* lens, int B, int H, int Tq, int Tk, int Dh, float scale,
                      int causal, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Model level.  A handle holds registered weights (by reference state_dict key) and built models.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct dim_handle_s* dim_handle_t;

int dim_create(dim_handle_t* out, int device);
int dim_destroy(dim_handle_t h);

#define DIM_DTYPE_F32 0
#define DIM_DTYPE_I64 1
/* Register a weight under its reference state_dict key (e.g. "listener_vq.encoder.squasher.0.0.weight"). */
int dim_set_tensor(dim_handle_t h, const char* name, const void* ptr, int dtype, int ndim, const int64_t* shape);

typedef struct {
  int32_t in_dim, hidden, layers, heads, ffn, n_embed, zdim, pe_max_len;
  float neg_slope;
  int32_t fqn;          /* face_quan_num: codes per frame (the encoder emits fqn*zdim channels); 0 means 1 */
  int32_t out_dim;      /* decoder output channels; 0 means in_dim */
} dim_vq_config;        /* code/config.yaml:15-30, code/config_speaker_old.yaml:15-30 */

/* Build a VQ-VAE from the tensors registered under `prefix` ("" or "listener_vq.").  Returns a model id >= 0 in *model. */
int dim_vqvae_build(dim_handle_t h, const char* prefix, const dim_vq_config* cfg, int precision, int* model);
/* The same with named halves: `encoder` / `decoder` are the module names under `prefix` (NULL: that half is absent and the
 * matching call fails with DIM_EINVAL).  VQSpeakerAutoEncoder (stage1_BIWI.py:140-173: one encoder, decoder_v -> 56 channels,
 * decoder_a -> 768 channels, 8 codes per frame) is three such models over one codebook: ("encoder", NULL), (NULL, "decoder_v")
 * with out_dim 56 and (NULL, "decoder_a") with out_dim 768. */
int dim_vqvae_build_parts(dim_handle_t h, const char* prefix, const char* encoder, const char* decoder, const dim_vq_config* cfg,
                          int precision, int* model);
size_t dim_vqvae_workspace_bytes(dim_handle_t h, int model, int B, int T);

/* VQAutoEncoder.encode (stage1_BIWI.py:22-27): x (B,T,in_dim) -> idx (B*T*fqn) int64 [, z (B,T,fqn*zdim) pre-quantisation
 * latents, quant (B,zdim,T*fqn) = E[idx] channel-major like the reference's return].  lens/batch_index (B) int32 nullable:
 * batch_index[b] selects the positional-encoding row (base_models.py:271-273); default b. */
int dim_vqvae_encode(dim_handle_t h, int model, const float* x, const int32_t* lens, const int32_t* batch_index, int B,
                     int T, int64_t* idx, float* z, float* quant_bcl, void* ws, size_t ws_bytes, void* stream);
/* VQAutoEncoder.decode (stage1_BIWI.py:29-37) of either codes (B*L) int64 -- the gather of seq2seq_pretrain.py:454-463
 * fused in front -- or a (B,zdim,L*fqn) fp32 `quant` tensor (exactly one of codes/quant_bcl non-NULL; codes: (B*L*fqn))
 * -> out (B,L,out_dim). */
int dim_vqvae_decode(dim_handle_t h, int model, const int64_t* codes, const float* quant_bcl, const int32_t* batch_index,
                     int B, int L, float* out, void* ws, size_t ws_bytes, void* stream);

typedef struct {
  int32_t dim_in, dim, dim_audio, depth, heads, dim_head, max_seq_len, num_tokens, ff_mult;
} dim_s2s_config;       /* seq2seq_pretrain.py:369-386,413 */

int dim_slmft_build(dim_handle_t h, const dim_s2s_config* cfg, int precision, int* model);
size_t dim_slmft_workspace_bytes(dim_handle_t h, int model, int B, int T, int steps);

/* SLMFT.forward_encoder + context assembly (seq2seq_pretrain.py:431-446): v_speaker (B,T,56), v_audio (B,T,768),
 * mask (B,T) uint8 -> x_s (B,T,dim) = norm_s(encoder_joint(encoder_s(v+patch_embed_s)))            [forward_encoder's return]
 *                  -> ctx (B,T,dim+dim_audio) = cat(x_s + patch_embed_dec_s, audio)                 [generate's context]
 * Either output may be NULL (v_audio may be NULL when ctx is). */
int dim_slmft_context(dim_handle_t h, int model, const float* v_speaker, const float* v_audio, const uint8_t* mask, int B,
                      int T, float* ctx, float* x_s, void* ws, size_t ws_bytes, void* stream);

/* One encoder call of the SLM pre-training forward (seq2seq_pretrain.py:205-226, x-transformers ContinuousTransformerWrapper with
 * return_embeddings=True): which = 0 encoder_s, 1 encoder_l, 2 encoder_joint; x (B,T,dim_in of that encoder) [+ add (dim_in) on every
 * frame, nullable]; mask (B,T) uint8 key padding; causal != 0 adds the causal attn_mask SLMFT passes; norm = 0 none, 1 norm_s,
 * 2 norm_l, 3 norm (the nn.LayerNorm heads of :224).  out (B,T,dim).  Workspace: dim_slmft_workspace_bytes(B, T, 1). */
int dim_slmft_encode(dim_handle_t h, int model, int which, const float* x, const float* add, const uint8_t* mask, int causal,
                     int norm, int B, int T, float* out, void* ws, size_t ws_bytes, void* stream);

/* decoder_joint.generate (seq2seq_pretrain.py:450; x-transformers AutoregressiveWrapper.generate): KV-cached
 * autoregressive decoding of `steps` tokens from prompt (B) int64, cross-attending ctx (B,T,D) under mask (B,T).
 * temperature == 0 -> argmax.  temperature > 0 -> top-k filter (top_k logits kept), softmax(logits/temperature) and an
 * inverse-CDF draw with uniforms (B,steps) fp32 supplied by the caller.  out_codes (B,steps) int64.
 * logits_out (B,steps,num_tokens) fp32 nullable.  Cross-attention K/V are projected once. */
int dim_slmft_generate(dim_handle_t h, int model, const float* ctx, const uint8_t* mask, const int64_t* prompt, int B,
                       int T, int steps, float temperature, int top_k, const float* uniforms, int64_t* out_codes,
                       float* logits_out, void* ws, size_t ws_bytes, void* stream);

/* The same decoding for `samples` independent draws per clip over ONE projection of the clip's context (x_engine_pt.py:255-270
 * calls the model 10 times per batch and keeps the best sample; the context, and therefore the cross-attention K/V, are the
 * same in all 10 calls).  Row r = b*samples + j: uniforms (B*samples, steps), out_codes (B*samples, steps),
 * logits_out (B*samples, steps, num_tokens) nullable.  Every row equals what dim_slmft_generate returns for clip b with that
 * row's uniforms (bit for bit: no arithmetic crosses rows). */
size_t dim_slmft_samples_workspace_bytes(dim_handle_t h, int model, int B, int T, int steps, int samples);
int dim_slmft_generate_samples(dim_handle_t h, int model, const float* ctx, const uint8_t* mask, const int64_t* prompt, int B,
                               int T, int steps, int samples, float temperature, int top_k, const float* uniforms,
                               int64_t* out_codes, float* logits_out, void* ws, size_t ws_bytes, void* stream);

/* Teacher-forced decoder forward (seq2seq_pretrain.py:447-448 -> x-transformers AutoregressiveWrapper.forward(..., return_outputs=True)):
 * tokens (B,L) int64 = the decoder INPUT sequence (z_l[:, :-1] with ignore_index already replaced by pad_value), ctx (B,T,D) and
 * mask (B,T) as for generate, kv_mask (B,L) uint8 nullable = the `self_attn_kv_mask` upstream draws at random when mask_prob > 0
 * (1 = key kept).  logits (B,L,num_tokens) fp32.  Forward only. */
size_t dim_slmft_teacher_forced_workspace_bytes(dim_handle_t h, int model, int B, int T, int L);
int dim_slmft_teacher_forced(dim_handle_t h, int model, const float* ctx, const uint8_t* mask, const int64_t* tokens,
                             const uint8_t* kv_mask, int B, int T, int L, float* logits, void* ws, size_t ws_bytes, void* stream);

/* The decode loop of dim_slmft_generate* runs as ONE persistent cooperative kernel (csrc/decode_mk.cu) whenever the model was
 * built for the tensor cores; its launches-turned-phases cannot be timed from outside, so the kernel itself can record the time
 * CTA 0 spends in every phase of the step program (summed over the steps of a call).
 *   dim_decode_trace_enable(1) before the generate call; dim_decode_trace_collect afterwards (synchronises the device):
 *   ms_per_phase[i], type_per_phase[i] for the i-th phase of one step (1 GEMM, 2 attention, 3 residual+LayerNorm, 4 GELU,
 *   5 sampling + embedding + LayerNorm).
 * dim_decode_set_impl(1) selects the per-kernel CUDA-graph decode path instead (0 = persistent kernel, the default): the two
 * implementations are each other's cross-check in the parity tests. */
int dim_decode_trace_enable(int on);
int dim_decode_trace_collect(double* h_ms_per_phase, int32_t* h_type_per_phase, int max_phases, int* h_n_out);
int dim_decode_set_impl(int impl);

#ifdef __cplusplus
}
#endif
#endif /* DIMB200_H */
