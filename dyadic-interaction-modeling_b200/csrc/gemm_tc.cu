// tcgen05 GEMM for sm_100a:  C[M,N] = epilogue( sum_pairs A_pa[M,Kp] @ W_pw[N,Kp]^T )   bf16 operands, fp32 accumulation in TMEM.
//
//   * operands are K-major bf16 "plane" matrices: Ap = [M, planes*Kp], Wp = [N, planes*Kp] (Kp = K rounded up to 64, zero
//     padded).  planes == 1 is a plain bf16 GEMM.  planes == 2/3 hold the bf16 split of fp32 values (x = h + m + l, 8 mantissa
//     bits each): summing the products of the listed plane pairs on the tensor cores reproduces an fp32 GEMM to ~2^-16
//     (3 pairs) or ~2^-23 (6 pairs) relative error per product -- this is how the fp32-parity mode reaches tensor-core speed.
//   * one 128 x BN output tile per CTA (or per CLUSTER when K is split); 192 threads: warp 0 = TMA producer, warp 1 = TMEM
//     allocator + MMA issuer (one elected lane issues tcgen05.mma; accumulator = 128 lanes x BN columns of TMEM),
//     warps 2-5 = epilogue.
//   * A and W tiles (128 x 64 and BN x 64 bf16) are staged by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a
//     STAGES-deep shared-memory ring guarded by full/empty mbarriers; tcgen05.commit releases slots and signals the epilogue.
//   * epilogue: tcgen05.ld 32x32b -> registers -> the CTA's (now idle) pipeline smem as an fp32 [128][BN+4] tile ->
//     whole rows are re-read by the warps so that bias / table / residual loads and the C stores are fully coalesced
//     128-bit accesses; bias, positional table, activation, residual and bf16 / bf16-plane outputs are fused here.
//   * split-K for short-M problems (the per-step decode GEMMs, M = batch, which cannot fill 148 SMs with output tiles):
//     the S CTAs of a thread-block cluster (1,1,S) each accumulate a K-slice of the same output tile, park their partial
//     tile in their own shared memory, and after a cluster barrier CTA r reduces rows [r*128/S, (r+1)*128/S) by reading
//     all S partial tiles over distributed shared memory in rank order (deterministic), then runs the fused epilogue.
//     No global workspace, no atomics.  S depends on K only, never on M: a row's bits do not depend on the batch it is in.
//   * smem is sized so that two CTAs fit per SM: one tile's epilogue overlaps the other's MMA main loop.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace dimb {

namespace {

constexpr int BM = 128, BKE = 64;            // tile rows, K elements per stage (64 bf16 = 128 B = one swizzle row)

struct TcParams {
  GemmArgs e;                 // epilogue fields + M, N (A/W/K of `e` are unused here)
  int kblocks;                // Kp / 64
  int npairs;
  int pa[6], pw[6];           // plane index of A / W for each accumulated pair
  int kp;                     // padded K (elements) = plane stride
  int splits;                 // split-K factor = cluster size along z (1, 2, 4 or 8)
  long long* dbg;             // optional timeline of CTA (0,0,0): clock64 stamps (debug / tuning only)
  float w_keep;               // > 0: fraction of the weight tiles loaded with an L2 evict_last policy (decode-step GEMMs:
                              //      keep part of the per-step weight stream L2-resident across steps), rest evict_first  // implicit 5-tap Conv1d (conv_cb > 0): A is the PADDED frame matrix [B*(T+4), planes*a_kp] (row b*(T+4)+t' = frame
  // clamp(t'-2, 0, L_b-1) of clip b), k-block kb = tap*conv_cb + c reads A rows m0+tap.. at columns c*64: the five taps are five
  // row-shifted TMA boxes of the same matrix, no im2col copy.  Tile rows live in the padded row space; the epilogue maps
  // m' = b*(T+4)+t back to output row b*T+t and drops t >= T.
  int conv_cb = 0, a_kp = 0, conv_T = 0, conv_Mp = 0;
};

template <int ACT>
__device__ __forceinline__ float act_fixed(float x, float slope) {
  if (ACT == DIM_ACT_LEAKY) return x > 0.f ? x : x * slope;
  if (ACT == DIM_ACT_GELU_TANH) {
    float u = 0.7978845608028654f * (x + 0.044715f * (x * x * x));
    return x * (0.5f * (1.0f + tanhf(u)));
  }
  if (ACT == DIM_ACT_GELU_ERF) return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
  return x;
}

// Rare epilogue features live out of line: the main loop has ONE warp per scheduler, so its code must stay small enough
// to sit in the instruction cache (a 45 KB epilogue ran at instruction-fetch speed).
__device__ __noinline__ float4 tc_tab_add(const GemmArgs& e, float4 o, int row, int col) {
  int g = e.tab_mode == 1 ? row / e.tab_T : row % e.tab_T;
  if (e.tab_mode == 1 && e.tab_index) g = __ldg(e.tab_index + g);
  const float4 t = __ldg(reinterpret_cast<const float4*>(e.tab + (size_t)g * e.ldtab + col));
  const float sc = e.tab_mode == 1 ? 1.0f : e.tab_scale;
  o.x = __fadd_rn(o.x, __fmul_rn(t.x, sc)); o.y = __fadd_rn(o.y, __fmul_rn(t.y, sc));
  o.z = __fadd_rn(o.z, __fmul_rn(t.z, sc)); o.w = __fadd_rn(o.w, __fmul_rn(t.w, sc));
  return o;
}
// plain bf16 copy of the result (cross-attention K/V rows; the bf16 QKV rows of the plain-bf16 prefill attention): inline -- six
// instructions; the plane split below stays out of line (instruction-cache footprint of the epilogue, see the header)
__device__ __forceinline__ void tc_emit_bf16(const GemmArgs& e, float4 o, int row, int col) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo);
  pk.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(e.Cb + (size_t)row * e.ldcb + col) = pk;
}
__device__ __noinline__ void tc_emit_narrow(const GemmArgs& e, float4 o, int row, int col) {
  if (e.Cp) {
    __nv_bfloat16* dst = e.Cp + (size_t)row * e.cp_planes * e.cp_kp + col;
    float v[4] = {o.x, o.y, o.z, o.w};
#pragma unroll 1
    for (int pl = 0; pl < e.cp_planes; ++pl) {
      unsigned short h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat16 b = __float2bfloat16_rn(v[j]);
        h[j] = __bfloat16_as_ushort(b);
        v[j] -= __bfloat162float(b);
      }
      uint2 pk;
      pk.x = (uint32_t)h[0] | ((uint32_t)h[1] << 16);
      pk.y = (uint32_t)h[2] | ((uint32_t)h[3] << 16);
      *reinterpret_cast<uint2*>(dst + (size_t)pl * e.cp_kp) = pk;
    }
  }
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  float4 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
  return v;
}

template <int BN, int STAGES, int ACT>
__global__ void __launch_bounds__(192) gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmW, const TcParams p) {
  constexpr uint32_t A_BYTES = BM * BKE * 2, W_BYTES = BN * BKE * 2, STAGE_BYTES = A_BYTES + W_BYTES;
  constexpr int CP = BN + 4;                           // fp32 staging-tile pitch (floats): conflict-free 128-bit accesses
  static_assert((size_t)BM * CP * 4 <= (size_t)STAGES * STAGE_BYTES, "staging tile must fit in the pipeline buffers");
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B tiles need 1024-byte alignment
  const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (dbg && threadIdx.x == 0) p.dbg[0] = clock64();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int total_kb = p.kblocks * p.npairs;
  const int S = p.splits;                              // == cluster size along z; blockIdx.z is the rank in the cluster
  // split-K: this CTA accumulates flattened (pair, k-block) iterations [it_begin, it_end)
  const int it_begin = (int)(((long)blockIdx.z * total_kb) / S);
  const int it_end = (int)(((long)(blockIdx.z + 1) * total_kb) / S);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                                   // whole warp allocates BN TMEM columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  // Everything above (barrier init, TMEM allocation, descriptor prefetch) touched no global data: with programmatic
  // dependent launch it overlapped the previous kernel's tail.  From here on operands written by that kernel are read.
  pdl_prologue();
  if (dbg && threadIdx.x == 0) p.dbg[1] = clock64();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint64_t wpol = 0;
      if (p.w_keep > 0.f)
        asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(wpol) : "f"(p.w_keep));
      for (int it = it_begin; it < it_end; ++it) {
        const int pair = it / p.kblocks, kb = it - pair * p.kblocks;
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_expect_tx(fb, STAGE_BYTES);
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        if (p.conv_cb > 0) {
          const int tap = kb / p.conv_cb, cb = kb - tap * p.conv_cb;
          tma_load_2d(sa, &tmA, p.pa[pair] * p.a_kp + cb * BKE, m0 + tap, fb);
        } else {
          tma_load_2d(sa, &tmA, p.pa[pair] * p.kp + kb * BKE, m0, fb);
        }
        if (p.w_keep > 0.f) tma_load_2d_hint(sa + A_BYTES, &tmW, p.pw[pair] * p.kp + kb * BKE, n0, fb, wpol);
        else tma_load_2d(sa + A_BYTES, &tmW, p.pw[pair] * p.kp + kb * BKE, n0, fb);
        if (dbg && it - it_begin < 16) p.dbg[8 + (it - it_begin)] = clock64();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor: D fp32 (1<<4), A bf16 (1<<7), B bf16 (1<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = it_begin; it < it_end; ++it) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg && it - it_begin < 16) p.dbg[24 + (it - it_begin)] = clock64();
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        const uint64_t adesc = make_sdesc(sa), bdesc = make_sdesc(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BKE / 16; ++k)            // UMMA_K = 16 bf16 = 32 bytes: +2 in the (addr >> 4) field
          umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it > it_begin || k > 0) ? 1u : 0u);
        umma_commit(smem_u32(&empty_bar[stage]));     // slot reusable once these MMAs have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(smem_u32(&accum_bar));              // accumulator complete (covers every MMA issued above)
    }
  } else {
    // ===== epilogue, phase 1 (warps 2..5; warp w may only touch TMEM lanes 32*(w%4) .. +31) =====
    // All MMAs are complete when accum_bar fires, hence every TMA write has been consumed: the pipeline buffers are free
    // and become the fp32 staging tile  stage_tile[row][col], pitch CP.
    const int q = warp & 3;
    mbar_wait(smem_u32(&accum_bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (dbg && threadIdx.x == 64) p.dbg[2] = clock64();
    float* stage_tile = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)));
    float* my_row = stage_tile + (size_t)(q * 32 + lane) * CP;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tmem_ld16(tlane + (uint32_t)c0, v);
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<float4*>(my_row + c0 + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
      if (dbg && threadIdx.x == 64 && c0 == 0) p.dbg[40] = clock64();
    }
    if (dbg && threadIdx.x == 64) p.dbg[5] = clock64();
  }
  // every partial tile of the cluster is now in its CTA's shared memory
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();                                        // re-converge warps 0/1 (single-lane roles) for the .aligned barrier
  if (S > 1) cluster_sync_all(); else __syncthreads();

  if (dbg && threadIdx.x == 64) p.dbg[6] = clock64();
  {
    // ===== epilogue, phase 2 (all 6 warps): this CTA finishes rows [rank*BM/S, (rank+1)*BM/S) of the tile =====
    // A lane owns 4 fixed columns (its bias is loaded once) and walks down the rows, RB independent rows per iteration so
    // that the smem/DSMEM reads, residual loads and stores of different rows overlap (one warp per scheduler: ILP matters).
    constexpr int LPR = BN / 4;                        // lanes per row (one float4 each)
    constexpr int RPI = 32 / LPR;                      // rows per warp instruction
    constexpr int RB = 4;                              // rows in flight per lane
    const GemmArgs& e = p.e;
    const int rows_here = BM / S, row_begin = (int)blockIdx.z * rows_here, row_end = row_begin + rows_here;
    const float* stage_tile = reinterpret_cast<const float*>(smem_raw + (smem_base - smem_u32(smem_raw)));
    const int c = (lane % LPR) * 4;
    const int col = n0 + c;
    if (col < e.N) {                                   // N % 4 == 0: a float4 is entirely inside or outside
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e.bias) bias4 = __ldg(reinterpret_cast<const float4*>(e.bias + col));
      const int stride = 6 * RPI;
#pragma unroll 1
      for (int r0 = row_begin + warp * RPI + lane / LPR; r0 < row_end; r0 += stride * RB) {
        float4 acc[RB], res[RB];
        bool ok[RB];
        int orow[RB];                                  // output row of tile row r
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          const int r = r0 + u * stride;
          orow[u] = m0 + r;
          ok[u] = r < row_end && orow[u] < e.M;
          if (p.conv_cb > 0) {                         // padded row space -> frame row (uniform branch)
            const int tp = p.conv_T + 4, b = orow[u] / tp, t = orow[u] - b * tp;
            ok[u] = r < row_end && orow[u] < p.conv_Mp && t < p.conv_T;
            orow[u] = b * p.conv_T + t;
          }
          acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          res[u] = acc[u];
        }
        if (e.residual) {                              // independent of the tile data: in flight under the DSMEM reads below
#pragma unroll
          for (int u = 0; u < RB; ++u)
            if (ok[u]) res[u] = *reinterpret_cast<const float4*>(e.residual + (size_t)orow[u] * e.ldr + col);
        }
        if (S == 1) {
#pragma unroll
          for (int u = 0; u < RB; ++u)
            if (ok[u]) acc[u] = *reinterpret_cast<const float4*>(stage_tile + (size_t)(r0 + u * stride) * CP + c);
        } else {
          // fixed rank order: deterministic sum.  The remote reads of 4 ranks x RB rows are issued back to back before the
          // first add (one DSMEM round trip per batch instead of one per rank: this loop was 5 k of the kernel's 15 k cycles).
#pragma unroll 1
          for (int z0 = 0; z0 < S; z0 += 4) {
            float4 v[4][RB];
#pragma unroll
            for (int zz = 0; zz < 4; ++zz) {          // branch-free: out-of-range ranks / rows read a valid address, dropped below
              const uint32_t rank = (uint32_t)min(z0 + zz, S - 1);
#pragma unroll
              for (int u = 0; u < RB; ++u)
                v[zz][u] = ld_dsmem_f4(smem_base + (uint32_t)((min(r0 + u * stride, BM - 1) * CP + c) * 4), rank);
            }
#pragma unroll
            for (int zz = 0; zz < 4; ++zz) {
              const bool zok = z0 + zz < S;
#pragma unroll
              for (int u = 0; u < RB; ++u) {
                const bool k = zok && ok[u];
                acc[u].x += k ? v[zz][u].x : 0.f; acc[u].y += k ? v[zz][u].y : 0.f;
                acc[u].z += k ? v[zz][u].z : 0.f; acc[u].w += k ? v[zz][u].w : 0.f;
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          if (!ok[u]) continue;
          const int row = orow[u];
          float4 o = acc[u];
          o.x += bias4.x; o.y += bias4.y; o.z += bias4.z; o.w += bias4.w;
          if (e.tab_mode != 0) o = tc_tab_add(e, o, row, col);
          o.x = act_fixed<ACT>(o.x, e.slope); o.y = act_fixed<ACT>(o.y, e.slope);
          o.z = act_fixed<ACT>(o.z, e.slope); o.w = act_fixed<ACT>(o.w, e.slope);
          o.x = __fadd_rn(o.x, res[u].x); o.y = __fadd_rn(o.y, res[u].y);
          o.z = __fadd_rn(o.z, res[u].z); o.w = __fadd_rn(o.w, res[u].w);
          if (e.C) *reinterpret_cast<float4*>(e.C + (size_t)row * e.ldc + col) = o;
          if (e.Cb) tc_emit_bf16(e, o, row, col);
          if (e.Cp) {
            if (e.cp_planes == 1) {                     // one plane = a plain bf16 row: the inline store (plain-bf16 FF1 -> FF2 operand)
              __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&lo);
              pk.y = *reinterpret_cast<uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(e.Cp + (size_t)row * e.cp_kp + col) = pk;
            } else {
              tc_emit_narrow(e, o, row, col);
            }
          }
        }
      }
    }
    if (dbg && threadIdx.x == 64) p.dbg[3] = clock64();
  }
  // peers may still be reading this CTA's staging tile over DSMEM: leave together
  __syncwarp();
  if (S > 1) cluster_sync_all(); else __syncthreads();
  if (dbg && threadIdx.x == 0) p.dbg[4] = clock64();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
  }
}

// ---- persistent flavour for the large-M (prefill) GEMMs ---------------------------------------------------------------------
// The one-tile-per-CTA kernel above pays barrier init + TMEM allocation + a cold pipeline per 128 x 128 tile and its epilogue
// only overlaps the OTHER resident CTA's main loop through the shared smem port: ncu put its tensor pipe at 28-66 %
// (profiles/r01z_ncu/summary.md).  Here one CTA per SM stays resident and walks the tile list (n fastest, so the CTAs of a wave
// share their A panel through L2 and W stays L2-resident):
//   * 128 x bn tiles, bn in {128, 192, 256} picked on the host so that N is covered without padding waste (1152 = 6 x 192):
//     a 256-wide tile moves 1.5x fewer operand bytes per flop from L2 than a 128-wide one;
//   * TWO accumulators in TMEM (2 x 256 of the 512 columns): the MMA warp starts tile i+1 while 8 epilogue warps drain tile i;
//     tmem_full / tmem_empty mbarriers hand the halves back and forth;
//   * the 4-slot TMA ring (16 KB A + up to 32 KB W per slot) never drains between tiles;
//   * every epilogue warp owns 32 TMEM lanes x bn/2 columns and transposes them through a PRIVATE 32 x 16 staging block
//     (__syncwarp only, no CTA-wide barrier anywhere in the steady state), then applies the same fused epilogue code, in the
//     same order, as the kernel above -- the accumulation order over K is unchanged too, so results are bit-identical.
constexpr int PG_THREADS = 320, PG_STAGES = 4, PG_EPI_WARPS = 8, PG_ACC_COLS = 256;
constexpr uint32_t PG_A_BYTES = BM * BKE * 2;                       // 16 KB
constexpr uint32_t PG_STAGE_BYTES = PG_A_BYTES + 256 * BKE * 2;     // + 32 KB W slot (a bn-row tile uses bn * 128 B of it)
constexpr int PG_SP = 20;                                           // staging pitch in floats: conflict-free float4 rows
constexpr uint32_t PG_STAGING = PG_EPI_WARPS * 32 * PG_SP * 4;
constexpr size_t PG_SMEM = (size_t)PG_STAGES * PG_STAGE_BYTES + PG_STAGING + 1024;

// bounded mbarrier wait: a protocol error traps after ~2 s instead of wedging the GPU
__device__ __forceinline__ void pg_wait(uint32_t bar, uint32_t parity) {
  unsigned long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((it & 1023) == 1023) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 2000000000ull) __trap();
    }
  }
}

template <int ACT>
__global__ void __launch_bounds__(PG_THREADS, 1) gemm_bf16_tcgen05_persist(const __grid_constant__ CUtensorMap tmA,
                                                                           const __grid_constant__ CUtensorMap tmW, const TcParams p,
                                                                           const int bn, const int pf, const int rpf) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[PG_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[PG_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];       // accumulator half complete (MMA -> epilogue)
  __shared__ __align__(8) uint64_t tempty_bar[2];      // accumulator half drained (8 epilogue warps -> MMA)
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_kb = p.kblocks * p.npairs;
  const int nt = (p.e.N + bn - 1) / bn, mt = (p.e.M + BM - 1) / BM, ntiles = mt * nt;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
#pragma unroll
    for (int s = 0; s < PG_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), PG_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                                     // the whole TMEM: two 256-column accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(2 * PG_ACC_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  pdl_prologue();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = PG_A_BYTES + (uint32_t)bn * (BKE * 2);
      // Optional L2 prefetch cursor, `pf` k-blocks ahead of the loads (DIM_GEMM_PF, default 0 = off).  Measured (B200, M = 76800): a
      // k-block costs ~600 + 1.3 bn cycles against 2 bn of MMA time, but an extra tensor prefetch per k-block made it WORSE
      // (1262 -> 835 TFLOP/s at N = 1536, K = 1152, 6 plane pairs; profiles/r02_prefill_gemm_pf.txt): the cost is per TMA
      // request, not HBM latency -- fewer operand rows per flop (wider tiles, CTA pairs) is what helps, not earlier requests.
      int pt = blockIdx.x, pit = 0;
      auto prefetch_next = [&]() {
        if (pt >= ntiles) return;
        const int pm = pt / nt, pn = pt - pm * nt;
        const int pair = pit / p.kblocks, kb = pit - pair * p.kblocks;
        tma_prefetch_2d(&tmA, p.pa[pair] * p.kp + kb * BKE, pm * BM);
        if (pt == (int)blockIdx.x) tma_prefetch_2d(&tmW, p.pw[pair] * p.kp + kb * BKE, pn * bn);
        if (++pit == total_kb) { pit = 0; pt += gridDim.x; }
      };
      for (int i = 0; i < pf; ++i) prefetch_next();
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int m_i = t / nt, n_i = t - m_i * nt;
        const int m0 = m_i * BM, n0 = n_i * bn;
        for (int it = 0; it < total_kb; ++it) {
          const int pair = it / p.kblocks, kb = it - pair * p.kblocks;
          if (pf > 0) prefetch_next();
          pg_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, stage_tx);
          const uint32_t sa = smem_base + stage * PG_STAGE_BYTES;
          tma_load_2d(sa, &tmA, p.pa[pair] * p.kp + kb * BKE, m0, fb);
          tma_load_2d(sa + PG_A_BYTES, &tmW, p.pw[pair] * p.kp + kb * BKE, n0, fb);
          if (++stage == PG_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int i = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
        const int a = i & 1;
        pg_wait(smem_u32(&tempty_bar[a]), (((uint32_t)i >> 1) & 1u) ^ 1u);      // the epilogue has drained this half (first use: free)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + (uint32_t)(a * PG_ACC_COLS);
        for (int it = 0; it < total_kb; ++it) {
          pg_wait(smem_u32(&full_bar[stage]), phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + stage * PG_STAGE_BYTES;
          const uint64_t adesc = make_sdesc(sa), bdesc = make_sdesc(sa + PG_A_BYTES);
#pragma unroll
          for (int k = 0; k < BKE / 16; ++k)
            umma_f16(acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == PG_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(smem_u32(&tfull_bar[a]));
      }
    }
  } else {
    // ===== epilogue: warp w drains TMEM lanes 32*(w%4) .. +31 (hardware rule), column half (w-2)/4 of every tile =====
    const int ew = warp - 2, q = warp & 3, hf = ew >> 2;
    float* st = reinterpret_cast<float*>(smem + (size_t)PG_STAGES * PG_STAGE_BYTES) + (size_t)ew * 32 * PG_SP;
    const int half = bn >> 1, cb = hf * half, ce = cb + half;
    const GemmArgs& e = p.e;
    const int rr = lane >> 2, cc = (lane & 3) * 4;      // re-read: 4 lanes per staged row, 8 rows per instruction
    int i = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
      const int a = i & 1;
      const int m_i = t / nt, n_i = t - m_i * nt;
      const int m0 = m_i * BM + q * 32, n0 = n_i * bn;
      if (e.residual && rpf && t + (int)gridDim.x < ntiles) {
        // residual rows of this CTA's NEXT tile -> L2 (one row segment per lane): the epilogue keeps only 4 float4 per lane in flight,
        // far too few to stream a cold residual tile at HBM latency
        const int t2 = t + (int)gridDim.x, m2 = t2 / nt, n2 = t2 - m2 * nt;
        const int row = m2 * BM + q * 32 + lane, c2 = n2 * bn + cb;
        const int cols = min(half, e.N - c2);
        if (row < e.M && cols > 0)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(e.residual + (size_t)row * e.ldr + c2), "r"(cols * 4) : "memory");
      }
      pg_wait(smem_u32(&tfull_bar[a]), ((uint32_t)i >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * PG_ACC_COLS);
      // software pipeline over the 16-column chunks: the TMEM load and the bias of chunk c+1 are requested before chunk c is
      // processed (ncu: 17 % of the epilogue's samples waited for tcgen05.ld, 22 % for the bias)
      uint32_t tv[16];
      tmem_ld16_issue(tl + (uint32_t)cb, tv);
      float4 bias_nx = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e.bias && n0 + cb + cc < e.N) bias_nx = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + cb + cc));
#pragma unroll 1
      for (int c0 = cb; c0 < ce; c0 += 16) {
        {
          tmem_ld16_wait(tv);
          float* w = st + lane * PG_SP;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<float4*>(w + g * 4) = make_float4(__uint_as_float(tv[g * 4]), __uint_as_float(tv[g * 4 + 1]),
                                                                __uint_as_float(tv[g * 4 + 2]), __uint_as_float(tv[g * 4 + 3]));
        }
        const float4 bias4 = bias_nx;
        if (c0 + 16 < ce) {
          tmem_ld16_issue(tl + (uint32_t)(c0 + 16), tv);
          if (e.bias && n0 + c0 + 16 + cc < e.N) bias_nx = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + c0 + 16 + cc));
        }
        __syncwarp();
        const int col = n0 + c0 + cc;
        if (col < e.N) {                                 // N % 4 == 0: a float4 is entirely inside or outside
          float4 acc[4], res[4];
          bool ok[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            ok[u] = m0 + u * 8 + rr < e.M;
            res[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (e.residual) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (ok[u]) res[u] = *reinterpret_cast<const float4*>(e.residual + (size_t)(m0 + u * 8 + rr) * e.ldr + col);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] = *reinterpret_cast<const float4*>(st + (u * 8 + rr) * PG_SP + cc);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (!ok[u]) continue;
            const int row = m0 + u * 8 + rr;
            float4 o = acc[u];
            o.x += bias4.x; o.y += bias4.y; o.z += bias4.z; o.w += bias4.w;
            if (e.tab_mode != 0) o = tc_tab_add(e, o, row, col);
            o.x = act_fixed<ACT>(o.x, e.slope); o.y = act_fixed<ACT>(o.y, e.slope);
            o.z = act_fixed<ACT>(o.z, e.slope); o.w = act_fixed<ACT>(o.w, e.slope);
            o.x = __fadd_rn(o.x, res[u].x); o.y = __fadd_rn(o.y, res[u].y);
            o.z = __fadd_rn(o.z, res[u].z); o.w = __fadd_rn(o.w, res[u].w);
            if (e.C) *reinterpret_cast<float4*>(e.C + (size_t)row * e.ldc + col) = o;
            if (e.Cb) tc_emit_bf16(e, o, row, col);
            if (e.Cp) {
              if (e.cp_planes == 1) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(e.Cp + (size_t)row * e.cp_kp + col) = pk;
              } else {
                tc_emit_narrow(e, o, row, col);
              }
            }
          }
        }
        __syncwarp();                                    // the staging block is rewritten by the next chunk
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty_bar[a])) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();                                          // re-converge the single-lane roles for the .aligned barrier
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * PG_ACC_COLS) : "memory");
  }
}

// ---- host: tensor maps ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 matrix [rows, cols] with row pitch ld (elements); box = 64 columns x box_rows rows, 128-byte swizzle.
int make_map(const __nv_bfloat16* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  using Key = std::tuple<const void*, int, int, int, int>;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  Key key{ptr, rows, cols, ld, box_rows};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return DIM_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(DIM_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {(cuuint32_t)BKE, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DIM_ECUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return DIM_OK;
}

template <int BN, int STAGES, int ACT>
int launch_tc_act(const CUtensorMap& tmA, const CUtensorMap& tmW, const TcParams& p, cudaStream_t s) {
  constexpr size_t smem = (size_t)STAGES * (BM * BKE * 2 + BN * BKE * 2) + 1024;
  static PerDeviceOnce once;
  if (once.first()) DIM_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05<BN, STAGES, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cdiv(p.e.N, BN), cdiv(p.conv_cb > 0 ? p.conv_Mp : p.e.M, BM), p.splits);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  // the S K-slices of one output tile form a cluster; un-split launches carry NO cluster attribute (plain CTA launch path)
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (p.splits > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = p.splits;
    ++na;
  }
  if (g_pdl_on) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  // algorithmic traffic: every operand plane once + the fp32 result; skinny (M < 2048: decode-step) launches are reported
  // separately because their roofline is HBM (weights streamed once per step), not the tensor pipe
  const int planes_n = p.npairs == 1 ? 1 : (p.npairs == 3 ? 2 : 3);
  ProfScope ps(p.e.M < 2048 ? CAT_GEMM_TC_SKINNY : CAT_GEMM_TC, s,
               2.0 * ((double)p.e.M + p.e.N) * p.kp * planes_n + 4.0 * p.e.M * p.e.N,
               2.0 * p.e.M * (double)p.e.N * p.kp * p.npairs);
  DIM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05<BN, STAGES, ACT>, tmA, tmW, p));
  DIM_LAUNCHED();
  return DIM_OK;
}

template <int ACT>
int launch_tc_persist_act(const CUtensorMap& tmA, const CUtensorMap& tmW, const TcParams& p, int bn, cudaStream_t s) {
  static PerDeviceOnce once;
  if (once.first())
    DIM_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_persist<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PG_SMEM));
  int dev = 0, sms = 0;
  DIM_CHECK_CUDA(cudaGetDevice(&dev));
  DIM_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ntiles = cdiv(p.e.M, BM) * cdiv(p.e.N, bn);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(std::min(sms, ntiles));
  cfg.blockDim = dim3(PG_THREADS);
  cfg.dynamicSmemBytes = PG_SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (g_pdl_on) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const int planes_n = p.npairs == 1 ? 1 : (p.npairs == 3 ? 2 : 3);
  ProfScope ps(CAT_GEMM_TC, s, 2.0 * ((double)p.e.M + p.e.N) * p.kp * planes_n + 4.0 * p.e.M * p.e.N,
               2.0 * p.e.M * (double)p.e.N * p.kp * p.npairs);
  static const int pf = getenv("DIM_GEMM_PF") ? std::max(0, atoi(getenv("DIM_GEMM_PF"))) : 0;      // operand L2 prefetch distance in k-blocks
  static const int rpf = getenv("DIM_GEMM_RPF") ? atoi(getenv("DIM_GEMM_RPF")) : 0;               // residual rows of the next tile -> L2 (measured: no gain)
  DIM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_persist<ACT>, tmA, tmW, p, bn, pf, rpf));
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_tc_persist(const CUtensorMap& tmA, const CUtensorMap& tmW, const TcParams& p, int bn, cudaStream_t s) {
  switch (p.e.act) {
    case DIM_ACT_LEAKY: return launch_tc_persist_act<DIM_ACT_LEAKY>(tmA, tmW, p, bn, s);
    case DIM_ACT_GELU_TANH: return launch_tc_persist_act<DIM_ACT_GELU_TANH>(tmA, tmW, p, bn, s);
    case DIM_ACT_GELU_ERF: return launch_tc_persist_act<DIM_ACT_GELU_ERF>(tmA, tmW, p, bn, s);
    default: return launch_tc_persist_act<DIM_ACT_NONE>(tmA, tmW, p, bn, s);
  }
}

template <int BN, int STAGES>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmW, const TcParams& p, cudaStream_t s) {
  switch (p.e.act) {
    case DIM_ACT_LEAKY: return launch_tc_act<BN, STAGES, DIM_ACT_LEAKY>(tmA, tmW, p, s);
    case DIM_ACT_GELU_TANH: return launch_tc_act<BN, STAGES, DIM_ACT_GELU_TANH>(tmA, tmW, p, s);
    case DIM_ACT_GELU_ERF: return launch_tc_act<BN, STAGES, DIM_ACT_GELU_ERF>(tmA, tmW, p, s);
    default: return launch_tc_act<BN, STAGES, DIM_ACT_NONE>(tmA, tmW, p, s);
  }
}

// fp32 -> bf16 planes.  One thread per 4 consecutive k of one row.
__global__ void __launch_bounds__(256) split_planes_kernel(const GemmArgs a, __nv_bfloat16* __restrict__ out, int kp,
                                                           int planes) {
  const int k4n = kp >> 2;
  const size_t total = (size_t)a.M * k4n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / k4n), k0 = (int)(i - (size_t)row * k4n) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (k0 < a.K) {
      const float* src;
      if (a.conv_T > 0) {
        int tap = k0 / a.conv_C, c = k0 - tap * a.conv_C;
        int b = row / a.conv_T, t = row - b * a.conv_T;
        int L = a.lens ? __ldg(a.lens + b) : a.conv_T;
        int ts = min(max(t + tap - 2, 0), L - 1);
        src = a.A + ((size_t)b * a.conv_T + ts) * a.lda + c;
      } else {
        src = a.A + (size_t)row * a.lda + k0;
      }
      float4 x = *reinterpret_cast<const float4*>(src);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
      if (a.a_add) {
        float4 e = __ldg(reinterpret_cast<const float4*>(a.a_add + k0));
        v[0] += e.x; v[1] += e.y; v[2] += e.z; v[3] += e.w;
      }
    }
    __nv_bfloat16* dst = out + (size_t)row * planes * kp + k0;
    for (int pl = 0; pl < planes; ++pl) {
      __nv_bfloat16 h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = __float2bfloat16_rn(v[j]);
        v[j] = v[j] - __bfloat162float(h[j]);          // exact: the remainder of an RN split is representable
      }
      uint2 pk;
      pk.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
      pk.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
      *reinterpret_cast<uint2*>(dst + (size_t)pl * kp) = pk;
    }
  }
}

// Conv operand: fp32 frames (B,T,C) -> padded bf16 plane matrix [B*(T+4), planes*cp]: row b*(T+4)+t' holds frame
// clamp(t'-2, 0, L_b-1) of clip b (replicate padding; lens as in the explicit gather), columns C..cp zero.
__global__ void __launch_bounds__(256) split_conv_pad_kernel(const GemmArgs a, __nv_bfloat16* __restrict__ out, int cp, int planes) {
  const int c4n = cp >> 2, T = a.conv_T, Tp = T + 4;
  const size_t total = (size_t)(a.M / T) * Tp * c4n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / c4n), c0 = (int)(i - (size_t)row * c4n) * 4;
    const int b = row / Tp, tq = row - b * Tp;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 < a.conv_C) {
      const int L = a.lens ? __ldg(a.lens + b) : T;
      const int ts = min(max(tq - 2, 0), L - 1);
      x = *reinterpret_cast<const float4*>(a.A + ((size_t)b * T + ts) * a.lda + c0);
    }
    store_planes4(out + (size_t)row * planes * cp + c0, x, planes, cp);
  }
}

}  // namespace

int launch_split_conv_pad(const GemmArgs& a, __nv_bfloat16* out, int planes, cudaStream_t s) {
  DIM_REQUIRE(a.conv_T > 0 && a.conv_C % 64 == 0 && a.M % a.conv_T == 0 && a.K == 5 * a.conv_C && a.a_add == nullptr,
              "split_conv_pad: needs C % 64 == 0 and whole clips");
  DIM_REQUIRE(planes >= 1 && planes <= 3, "split: planes must be 1..3");
  const size_t total = (size_t)(a.M / a.conv_T) * (a.conv_T + 4) * (a.conv_C / 4);
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 32);
  ProfScope ps(CAT_MISC, s, (double)a.M * a.conv_C * 4.0 + (double)total * 4 * planes * 2.0, 0);
  split_conv_pad_kernel<<<blocks, 256, 0, s>>>(a, out, a.conv_C, planes);
  DIM_LAUNCHED();
  return DIM_OK;
}

int tc_make_map(const __nv_bfloat16* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  return make_map(ptr, rows, cols, ld, box_rows, out);
}

int launch_split_planes(const GemmArgs& a, __nv_bfloat16* out, int kp, int planes, cudaStream_t s) {
  DIM_REQUIRE(a.M > 0 && a.K > 0 && a.K % 4 == 0 && kp % 64 == 0 && kp >= a.K, "split: bad K");
  DIM_REQUIRE(planes >= 1 && planes <= 3, "split: planes must be 1..3");
  if (a.conv_T > 0) DIM_REQUIRE(a.conv_C % 4 == 0 && a.K == 5 * a.conv_C, "split: conv mode needs K = 5*C");
  size_t total = (size_t)a.M * (kp / 4);
  int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 32);
  ProfScope ps(CAT_MISC, s, (double)a.M * a.K * 4.0 + (double)a.M * kp * planes * 2.0, 0);
  split_planes_kernel<<<blocks, 256, 0, s>>>(a, out, kp, planes);
  DIM_LAUNCHED();
  return DIM_OK;
}

long long* g_tc_dbg = nullptr;
int g_tc_force_splits = 0;
int g_tc_force_bn = 0;

int tc_pairs(int planes, int* pa, int* pw) {
  // plane 0 = high, 1 = middle, 2 = low part of the bf16 split.  Products kept: all with (index sum) < planes.
  int n = 0;
  for (int s = planes - 1; s >= 0; --s)        // smallest contributions first
    for (int i = 0; i <= s; ++i) {
      pa[n] = i;
      pw[n] = s - i;
      ++n;
    }
  return n;
}

int launch_gemm_tc(const GemmArgs& e, const __nv_bfloat16* Ap, const __nv_bfloat16* Wp, int kp, int planes, cudaStream_t s) {
  DIM_REQUIRE(e.M > 0 && e.N > 0 && kp > 0 && kp % BKE == 0, "gemm_tc: Kp must be a positive multiple of 64");
  DIM_REQUIRE(e.N % 4 == 0 && (e.C == nullptr || e.ldc % 4 == 0) && (e.Cb == nullptr || e.ldcb % 4 == 0),
              "gemm_tc: N and output pitches must be multiples of 4");
  DIM_REQUIRE(planes >= 1 && planes <= 3, "gemm_tc: planes must be 1..3");
  DIM_REQUIRE(e.C != nullptr || e.Cb != nullptr || e.Cp != nullptr, "gemm_tc: no output");
  DIM_REQUIRE(e.Cp == nullptr || (e.cp_planes >= 1 && e.cp_planes <= 3 && e.cp_kp >= e.N), "gemm_tc: bad plane output");
  DIM_REQUIRE(((uintptr_t)Ap & 15) == 0 && ((uintptr_t)Wp & 15) == 0, "gemm_tc: operands must be 16-byte aligned");
  TcParams p;
  p.e = e;
  p.kblocks = kp / BKE;
  p.kp = kp;
  p.npairs = tc_pairs(planes, p.pa, p.pw);
  p.splits = 1;
  p.dbg = g_tc_dbg;
  const bool conv = e.conv_T > 0;                       // Ap = padded frame planes (launch_split_conv_pad), kp = 5 * C
  if (conv) {
    DIM_REQUIRE(e.conv_C % BKE == 0 && kp == 5 * e.conv_C && e.M % e.conv_T == 0, "gemm_tc: conv mode needs C % 64 == 0, Kp = 5*C");
    p.conv_cb = e.conv_C / BKE;
    p.a_kp = e.conv_C;
    p.conv_T = e.conv_T;
    p.conv_Mp = e.M / e.conv_T * (e.conv_T + 4);
  }
  // Decode-step GEMMs re-read the same weights every step (144 MB of bf16 planes per step, more than the 126 MB L2): load 70 % of
  // the weight tiles with an L2 evict_last policy and the rest evict_first, so that most of the stream is served from L2 on the
  // next step instead of thrashing (measured: 261.9 -> 251.8 ms per bench step; 0.4: 254.5, 1.0: 256.4; profiles/r01_notes.md).
  // The K/V streams of the decode attention carry evict_first.  DIM_L2_WEIGHT_KEEP overrides (0 disables); plain-bf16 mode only.
  static const float w_keep_env = getenv("DIM_L2_WEIGHT_KEEP") ? (float)atof(getenv("DIM_L2_WEIGHT_KEEP")) : -1.f;
  const float w_keep_default = planes == 1 ? 0.7f : 0.f;
  p.w_keep = e.split_hint == DIM_SPLIT_DECODE ? (w_keep_env >= 0.f ? w_keep_env : w_keep_default) : 0.f;
  const int total_kb = p.kblocks * p.npairs;
  int bn = e.N >= 128 ? 128 : (e.N > 32 ? 64 : 32);
  int splits = 1;
  const bool skinny = e.split_hint == DIM_SPLIT_DECODE || (e.split_hint == DIM_SPLIT_AUTO && e.M < 2048);
  if (skinny) {
    if (planes == 1 && (e.split_hint == DIM_SPLIT_DECODE || e.M <= 512)) {
      // Decode-step GEMMs with plain bf16 operands (measured: profiles/r01g_tc_sweep_*.txt, r01l_graph_gap.txt): a CTA ingests
      // ~60 B/clk, so its main loop costs (128 + BN) * K * 2 B / 60 and the 128-row A tile dominates; the DSMEM reduction of a
      // split costs ~20 B/clk per SM and a cluster launch is slower to schedule.  Narrow tiles with the whole K per CTA and no
      // cluster win for K <= 1152; long K (FF2, 4608) takes 64-wide tiles split 4 ways.  Depends on K only, never on M or N.
      if (kp >= 2048) { bn = 64; splits = 4; }
      else { bn = 32; splits = 1; }
    } else if (planes == 3 && e.split_hint == DIM_SPLIT_DECODE && kp <= 2048) {
      // fp32-grade decode-step GEMMs, K <= 1152 (108 plane-pair k-blocks): 64-wide tiles split 4 ways beat 128-wide tiles
      // split 8 ways (profiles/r01x_tc_sweep3_m128.txt: 18.7 vs 22.9 us QKV, 12.9 vs 14.7 us out-proj, 20.9 vs 35.9 us FF1
      // with a 2-way split -- 288 clusters of 8 do not fit one wave).  Depends on (N, K) only.
      bn = 64;
      splits = e.N >= 4096 ? 2 : 4;
    } else {
      // fp32-grade plane products with long K, and mid-sized M: split K over a cluster so that every CTA owns >= ~4 k-blocks.
      const int want = total_kb / 4;
      splits = want >= 8 ? 8 : (want >= 4 ? 4 : (want >= 2 ? 2 : 1));
    }
  }
  if (g_tc_force_bn > 0) bn = g_tc_force_bn;
  p.splits = splits;
  if (g_tc_force_splits > 0) p.splits = g_tc_force_splits;
  while (p.splits > 1 && p.splits > total_kb) p.splits >>= 1;
  CUtensorMap tmA, tmW;
  // Large-M (prefill) GEMMs: the persistent kernel with a double-buffered TMEM accumulator.  Tile width = the widest of 256 / 192 /
  // 128 that covers N without padding waste -- a function of N only.  DIM_GEMM_PERSIST=0 keeps the one-tile-per-CTA kernel (A/B).
  static const bool persist_off = getenv("DIM_GEMM_PERSIST") != nullptr && atoi(getenv("DIM_GEMM_PERSIST")) == 0;
  if (!persist_off && !conv && !skinny && p.splits == 1 && e.N >= 128 && g_tc_force_bn <= 0 && p.dbg == nullptr) {
    int pbn = 256, best = cdiv(e.N, 256) * 256;
    if (cdiv(e.N, 192) * 192 < best) { pbn = 192; best = cdiv(e.N, 192) * 192; }
    if (cdiv(e.N, 128) * 128 < best) { pbn = 128; best = cdiv(e.N, 128) * 128; }
    // long main loops (>= 24 k-block iterations per tile) are bound by operand traffic, where the widest tile wins even with
    // 10 % of padded columns (N = 1152, 6 pairs, K = 384: 390 us at 256 vs 418 at 192); short ones by the epilogue, where padding loses
    if (total_kb >= 24 && e.N >= 1024 && cdiv(e.N, 256) * 256 * 8 <= e.N * 9) pbn = 256;
    if (g_tc_force_bn < 0) pbn = -g_tc_force_bn;          // tuning hook: dim_debug_tc_bn(-bn), bn % 32 == 0, 32 <= bn <= 256
    if (int err = make_map(Ap, e.M, planes * kp, planes * kp, BM, &tmA)) return err;
    if (int err = make_map(Wp, e.N, planes * kp, planes * kp, pbn, &tmW)) return err;
    return launch_tc_persist(tmA, tmW, p, pbn, s);
  }
  if (conv) {
    if (int err = make_map(Ap, p.conv_Mp, planes * p.a_kp, planes * p.a_kp, BM, &tmA)) return err;
  } else if (int err = make_map(Ap, e.M, planes * kp, planes * kp, BM, &tmA)) return err;
  if (int err = make_map(Wp, e.N, planes * kp, planes * kp, bn, &tmW)) return err;
  if (bn == 128) return launch_tc<128, 3>(tmA, tmW, p, s);
  if (bn == 64) return launch_tc<64, 4>(tmA, tmW, p, s);
  return launch_tc<32, 4>(tmA, tmW, p, s);
}

}  // namespace dimb

using namespace dimb;

extern "C" int dim_split_bf16_planes(const float* X, int ldx, int rows, int K, int planes, void* out, void* stream) {
  if (int e = ensure_device()) return e;
  GemmArgs a;
  a.A = X; a.lda = ldx; a.M = rows; a.K = K;
  return launch_split_planes(a, static_cast<__nv_bfloat16*>(out), tc_round_k(K), planes, as_stream(stream));
}

extern "C" int dim_linear_bf16_planes(const void* Ap, const void* Wp, int K, int planes, const float* bias,
                                      const float* residual, int ldr, float* C, int ldc, int M, int N, int act, float slope,
                                      void* stream) {
  if (int e = ensure_device()) return e;
  GemmArgs a;
  a.bias = bias; a.residual = residual; a.ldr = ldr; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K; a.act = act;
  a.slope = slope;
  return launch_gemm_tc(a, static_cast<const __nv_bfloat16*>(Ap), static_cast<const __nv_bfloat16*>(Wp), tc_round_k(K), planes,
                        as_stream(stream));
}

// tuning hooks (not part of the stable ABI): timeline buffer for CTA (0,0,0) and a split-K override
extern "C" int dim_debug_tc(void* dbg_buffer_64x8B, int force_splits) {
  dimb::g_tc_dbg = static_cast<long long*>(dbg_buffer_64x8B);
  dimb::g_tc_force_splits = force_splits;
  return DIM_OK;
}
extern "C" int dim_debug_tc_bn(int force_bn) {
  dimb::g_tc_force_bn = force_bn;
  return DIM_OK;
}
