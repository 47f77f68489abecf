// fp32 GEMM family (parity mode): C = epilogue(A @ W^T), nn.Linear weight layout W[N,K].
//   * tiled register-blocked FFMA kernel for M > 8 (also serves the 5-tap replicate-padded Conv1d as an implicit GEMM)
//   * weight-streaming skinny kernel for M <= 8 (the autoregressive decode step at small batch: HBM-bound on W)
#pragma once
#include "common.cuh"

namespace dimb {

enum { DIM_SPLIT_AUTO = 0, DIM_SPLIT_NEVER = 1, DIM_SPLIT_DECODE = 2 };

struct GemmArgs {
  const float* A = nullptr; int lda = 0;      // [M,K] rows lda apart (conv mode: frames (B,T,Cin), lda = Cin)
  const float* W = nullptr;                   // [N,K] row-major
  const float* bias = nullptr;                // [N]
  const float* residual = nullptr; int ldr = 0;   // [M,N], added after the activation
  float* C = nullptr; int ldc = 0;            // [M,N]
  __nv_bfloat16* Cb = nullptr; int ldcb = 0;  // optional bf16 copy of C
  const __nv_bfloat16* Ap = nullptr;          // tensor-core path: A already split into bf16 planes [M, planes*Kp] (skips the split pass)
  __nv_bfloat16* Cp = nullptr;                // tensor-core path: also emit C as bf16 planes [M, cp_planes*cp_kp] for the next GEMM
  int cp_planes = 0, cp_kp = 0;
  int M = 0, N = 0, K = 0;
  int act = DIM_ACT_NONE; float slope = 0.f;
  const float* a_add = nullptr;               // [K] added to every row of A while loading (patch_embed_s)
  int conv_T = 0, conv_C = 0;                 // conv mode when conv_T > 0: K = 5*conv_C, row r=(b,t) gathers t-2..t+2 clamped
  const int32_t* lens = nullptr;              // [B] valid lengths for the conv clamp
  const float* tab = nullptr; int ldtab = 0;  // table added before the activation:
  int tab_mode = 0;                           //   1: row tab_index[r / tab_T] (or r / tab_T)  -- pe[batch]  (F4 quirk)
  const int32_t* tab_index = nullptr;         //   2: row (r % tab_T)                          -- pos_emb[t] * tab_scale
  int tab_T = 1; float tab_scale = 1.f;
  int split_hint = 0;                         // tensor-core path: DIM_SPLIT_AUTO (by M) / _NEVER / _DECODE (see launch_gemm_tc)
};

int launch_gemm_f32(const GemmArgs& a, cudaStream_t s);
// M <= 8 rows, weight row read as bf16 from plane 0 of the tensor-core weight image [N, kp] (DIM_PREC_BF16)
int launch_gemv_bf16w(const GemmArgs& a, const __nv_bfloat16* Wb, int kp, cudaStream_t s);

}  // namespace dimb
