// Attention kernels (fp32 parity mode).  See attention.cu.
#pragma once
#include "common.cuh"

namespace dimb {

struct AttnArgs {
  const float *q = nullptr, *k = nullptr, *v = nullptr;   // element (b=0,t=0,h=0,d=0); rows ld* apart; head h at h*Dh
  int ldq = 0, ldk = 0, ldv = 0;
  int in_bf16 = 0;                                        // 1: q/k/v point to bf16 data (ld* in bf16 elements): the QKV GEMM's bf16 output,
                                                          //    plain-bf16 mode only (planes == 1, tensor-core kernel)
  float* out = nullptr; int ldo = 0;                      // (B,Tq,H*Dh)  (nullable when out_p is given)
  __nv_bfloat16* out_p = nullptr; int planes = 0, kp = 0; // optional bf16-plane copy [B*Tq, planes*kp] for the next GEMM
  const uint8_t* key_mask = nullptr;                      // (B,Tk) 1 = keep  (masked -> -FLT_MAX)
  const int32_t* lens = nullptr;                          // (B) keys >= lens[b] do not exist
  int B = 0, H = 0, Tq = 0, Tk = 0, Dh = 0;
  float scale = 1.f;
  int causal = 0;
};
int launch_attention_prefill(const AttnArgs& a, cudaStream_t s);

struct DecodeAttnArgs {
  const float* q = nullptr; int ldq = 0;                  // (B, H*64) this step's queries
  void *k = nullptr, *v = nullptr;                        // cache bases (fp32, or bf16 when kv_bf16): element (b,t,h,d) at
                                                          //   b*kv_batch_stride + h*kv_head_stride + t*kv_tok_stride + d (elements)
  int kv_bf16 = 0;                                        // head-major caches [B,H,tokens,64]: head_stride = tokens*64, tok_stride = 64
  size_t kv_batch_stride = 0, kv_head_stride = 64; int kv_tok_stride = 0;   // (token-major [B,tokens,H*64]: 64 / H*64)
  const float *k_new = nullptr, *v_new = nullptr; int ld_new = 0;   // rows appended at position *step when append != 0
  int append = 0;
  const int* step = nullptr;                              // device scalar: index of the token being decoded
  const uint8_t* key_mask = nullptr;                      // (B,Tk), cross attention only
  float* out = nullptr; int ldo = 0;                      // (B, H*64)  (nullable when out_p is given)
  __nv_bfloat16* out_p = nullptr; int planes = 0, kp = 0; // optional bf16-plane copy [B, planes*kp]
  int B = 0, H = 0, Tk = 0;                               // Tk: number of keys when append == 0
  int kv_group = 1;                                       // rows b*kv_group .. +kv_group-1 share cache row b (and its key mask):
                                                          //   several samples per clip over one cross-attention K/V
  float scale = 1.f;
  int sc_floats = 0;                                      // set by the launcher
  int kv_evict_first = 0;                                 // set by the launcher: K/V loads carry an L2 evict_first policy
  int prof_pos = 0;                                       // host copy of *step, for profiling byte counts only
};
int launch_attention_decode(DecodeAttnArgs a, int max_keys, cudaStream_t s);

// Token-major projected K|V rows [B*T, 2*H*64] (K at column 0, V at H*64; fp32 or bf16) -> head-major caches
// dst = K[B,H,T,64] followed by V[B,H,T,64]: every (b,h) head is then one contiguous stream for the decode kernel.
int launch_kv_head_major(const void* src, void* dst, int B, int T, int H, int bf16, cudaStream_t s);

}  // namespace dimb
