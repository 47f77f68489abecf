// Row-wise kernels (LayerNorm, InstanceNorm, context assembly, token embedding, sampling).  See rowops.cu.
#pragma once
#include "common.cuh"
#include <algorithm>

namespace dimb {
// y (fp32), yb (bf16 copy) and yp (bf16 planes [rows, planes*kp], kp == dim) are each optional outputs.
int launch_layer_norm(const float* x, const float* gain, const float* bias, float* y, __nv_bfloat16* yb, int rows, int dim,
                      float eps, cudaStream_t s, __nv_bfloat16* yp = nullptr, int planes = 0, int kp = 0);
int launch_instance_norm(float* x, const int32_t* lens, int B, int T, int C, float eps, cudaStream_t s);
int launch_build_context(const float* xs, const float* pe_dec, const float* audio, float* ctx, __nv_bfloat16* ctxb,
                         size_t rows, int d1, int d2, cudaStream_t s);
// padded batch tensors from packed ragged clips (the loader's collate step, on the device)
int launch_assemble_batch(const float* speaker, const float* audio, const float* listener, const int64_t* offsets, int B, int T, int Dm,
                          int Da, float* src, float* tgt, uint8_t* mask, cudaStream_t s);
// x[b,t,:] += tab[t,:] * scale   (absolute positional embedding of a teacher-forced sequence)
int launch_add_pos_table(float* x, const float* tab, float scale, int B, int L, int D, cudaStream_t s);
int launch_embed_tokens(const int64_t* tok, int tok_stride, const int* step, const float* emb, float* x, int B, int D, int V,
                        cudaStream_t s, const float* pos = nullptr, float pos_scale = 0.f);
int launch_sample(const float* logits, int B, int V, float temperature, int top_k, const float* uniforms, int u_stride,
                  const int* step, int64_t* out, int out_stride, int out_offset, float* logits_out, int lo_stride,
                  cudaStream_t s);
// sample + step++ + next step's token embedding + layer-0 LayerNorm in one launch (see rowops.cu)
int launch_sample_next(const float* logits, int B, int V, float temperature, int top_k, const float* uniforms, int u_stride,
                       int* step, unsigned int* ticket, int64_t* out, int out_stride, int out_offset, float* logits_out,
                       int lo_stride, const float* emb, float* x, int D, const float* gain, const float* bias, float* y,
                       __nv_bfloat16* yp, int planes, float eps, cudaStream_t s, const float* pos = nullptr, float pos_scale = 0.f);
int launch_resample(const float* in, float* out, int t, int d, int new_t, int window, int mode, cudaStream_t s);
int launch_init_tokens(int64_t* tokens, int stride, const int64_t* prompt, int rows, int samples, cudaStream_t s);
// out[b] = number of kept keys when mask[b, :] is a prefix mask, -1 otherwise
int launch_mask_prefix(const uint8_t* mask, int B, int T, int32_t* out, cudaStream_t s);
int launch_advance_step(int* step, cudaStream_t s);
int launch_set_step(int* step, int v, cudaStream_t s);
}  // namespace dimb
