// tcgen05 (5th-gen tensor core) GEMM on bf16 plane operands + the fp32 -> bf16-plane split.  See gemm_tc.cu.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_f32.cuh"

namespace dimb {

inline int tc_round_k(int K) { return (K + 63) / 64 * 64; }

// out[row, p*kp + k] = p-th bf16 part of A'(row,k) for k < K (0 for K <= k < kp), where A' is A with the GemmArgs
// prologue applied (a_add, conv-mode gather with replicate clamp / lens).  Uses a.A, a.lda, a.M, a.K, a.a_add, a.conv_*.
int launch_split_planes(const GemmArgs& a, __nv_bfloat16* out, int kp, int planes, cudaStream_t s);

// Conv operand for the implicit 5-tap conv GEMM: frames (B,T,C) fp32 -> padded plane matrix [B*(T+4), planes*C] (C % 64 == 0):
// row b*(T+4)+t' = frame clamp(t'-2, 0, lens[b]-1).  1.01x the activation bytes instead of the 5x im2col rows.
int launch_split_conv_pad(const GemmArgs& a, __nv_bfloat16* out, int planes, cudaStream_t s);

// C = epilogue(sum over plane pairs of Ap_i @ Wp_j^T); epilogue fields, M and N are taken from `e`.
// e.conv_T > 0: implicit conv -- Ap is the padded matrix of launch_split_conv_pad, kp = 5*C, e.M = B*T output rows.
int launch_gemm_tc(const GemmArgs& e, const __nv_bfloat16* Ap, const __nv_bfloat16* Wp, int kp, int planes, cudaStream_t s);

int tc_pairs(int planes, int* pa, int* pw);
// TMA descriptor of a bf16 matrix [rows, cols] with row pitch ld (elements): box = 64 columns x box_rows rows, 128-byte swizzle
int tc_make_map(const __nv_bfloat16* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out);
extern long long* g_tc_dbg;
extern int g_tc_force_splits;
extern int g_tc_force_bn;

}  // namespace dimb
