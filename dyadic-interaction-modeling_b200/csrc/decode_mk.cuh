// Persistent decode kernel ("megakernel"): every autoregressive step of decoder_joint.generate in ONE cooperative launch.
// See decode_mk.cu.  The host (engine.cu) fills an MkPlan -- a phase program for one decode step plus the TMA descriptors it
// needs -- and launches decode_megakernel with it as the kernel parameter.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dimb {

enum { MK_GEMM = 1, MK_ATTN = 2, MK_ROW_RESLN = 3, MK_ROW_GELU = 4, MK_ROW_SAMPLE = 5, MK_NOP = 6, MK_NTYPES = 7 };

constexpr int MK_MAX_PHASES = 64;
constexpr int MK_MAX_MAPS = 48;
constexpr int MK_PART_FLOATS = 9216;     // split-K partial columns per decode row (max over the GEMMs of splits * N)

struct MkPhase {
  int type;
  int M, N;                                  // rows (decode rows), output columns
  // ---- MK_GEMM: part[z][M][N] = A_planes[M, :] . W_planes[N, :]^T over the z-th slice of the (plane pair, k-block) iterations
  int mapA, mapW;                            // indices into MkPlan::maps
  int kblocks, npairs, splits, bn, kp;
  int pa[6], pw[6];
  float w_keep;                              // > 0: fraction of the weight tiles loaded with L2 evict_last
  int epi;                                   // 0: fp32 partials -> part; 1 (splits == 1 only): gelu_erf(acc + bias) -> planes outp
  // <= 8 decode rows: the phase runs as a GEMV (every warp streams whole weight rows; A staged in shared memory), splits == 1
  int gemv, K;
  const __nv_bfloat16* a_planes;             // the A operand behind mapA: [M, planes * kp] bf16
  const __nv_bfloat16* wb;                   // planes == 1: bf16 weight image [N, kp]
  const float* w32;                          // planes == 3: the fp32 weight [N, K] (A is reconstructed exactly from its planes)
  float* part;
  // ---- MK_ATTN: one query row per (decode row, head) over a head-major K/V cache
  int q_splits, q_ld, q_col, k_col, v_col;   // the producing GEMM's partials: pitch and column offsets of q / new k / new v
  int append, Tk, kv_group;
  void *kcache, *vcache;
  unsigned long long kv_batch_stride, kv_head_stride;
  const uint8_t* key_mask;
  const int32_t* key_valid;                  // nullable, per clip: >= 0 = key_mask[b] keeps exactly the first key_valid[b] keys (no byte loads)
  float scale;
  // ---- row phases (also the outputs of MK_ATTN): bf16 planes [M, planes * out_kp] = the next GEMM's A operand
  __nv_bfloat16* outp;
  int out_kp;
  int in_splits;                             // partial slices to sum (row phases)
  const float* bias;                         // [N] nullable
  float* x;                                  // residual stream [M, N]
  const float *gain, *beta;                  // LayerNorm (beta nullable)
  const float* emb;                          // MK_ROW_SAMPLE: token embedding [V, D]
  const float* pos;                          // MK_ROW_SAMPLE: nullable absolute positional table [max_seq_len, D] (+ pos[st + 1] * pos_scale)
  float pos_scale;
  int D;                                     // MK_ROW_SAMPLE: model width (N = vocabulary)
  // ---- any non-attention phase: while it runs (HBM nearly idle) the CTA asks for a slice of the NEXT attention phase's K/V blocks
  // to be brought into L2 (decode_mk.cu: mk_kv_prefetch).  pf_target = index of that attention phase (-1: none); the phase covers
  // the fraction [pf_f0, pf_f1) of the entries the byte budget allows.
  int pf_target;
  float pf_f0, pf_f1;
};

struct alignas(128) MkPlan {
  CUtensorMap maps[MK_MAX_MAPS];
  MkPhase phases[MK_MAX_PHASES];
  int nphases;
  int nmaps;
  int B;                                     // decode rows
  int H;
  int planes;
  int kv_bf16;
  int steps;
  int attn_stages;                           // cp.async ring slots of the mma attention items (3 when the scores fit beside them)
  int attn_mma;                              // bf16 caches: 1 = mma.sync attention items, 0 = FFMA items (A/B hook DIM_MK_ATTN_FFMA)
  int attn_pre;                              // mma items: rows of the first item's K block (half as many of V) prefetched into L2 before the
                                             // grid barrier that precedes an attention phase (DIM_MK_ATTN_PRE=rows, 0 = off)
  int attn_dbg;                              // timing ablations of the mma attention items (DIM_MK_ATTN_DBG bits: 1 no K/V loads, 2 no products, 4 no projection partials, 8 no mask bytes; results are wrong)
  unsigned long long pf_budget;              // K/V bytes per CTA and attention phase to prefetch into L2 (0 = off)
  int attn_nsub;                             // attention sub-groups per CTA: 6 (384 threads) or 8 (512-thread flavour, mk_attn_subgroups)
  int sc_floats;                             // score slots per attention work item (>= max keys, multiple of 4)
  unsigned int* bar;                         // grid barrier counter, zero at launch
  unsigned long long* trace;                 // nullable: [MK_MAX_PHASES] ns spent per phase (summed over steps), CTA 0's view
  int64_t* tokens;                           // [B, tok_stride]; column 0 = prompt, column st+1 = token of step st
  int tok_stride;
  const float* uniforms;                     // [B, u_stride] (sampling)
  int u_stride;
  float* logits_out;                         // nullable [B, lo_stride]
  long long lo_stride;
  float temperature;
  int top_k;
};

// True when the persistent kernel can decode this configuration (otherwise the caller keeps the per-kernel path).
bool mk_supported(int D, int inner, int F, int V, int H, int planes, int max_keys);
// Bytes of shared-memory scratch one attention sub-group needs for `max_keys` keys, and the sub-group count the launch will use.
size_t mk_attn_scratch(int max_keys);
int mk_attn_subgroups(int kv_bf16, int attn_mma, int small, int max_keys, int items);
// Cooperative launch of the persistent kernel; the plan travels as a __grid_constant__ kernel parameter (TMA descriptors
// included).  The barrier counter (plan.bar) must have been zeroed on the same stream.
int launch_decode_megakernel(const MkPlan& plan, cudaStream_t s);

}  // namespace dimb
