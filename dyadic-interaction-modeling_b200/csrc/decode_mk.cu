// Persistent decode kernel for sm_100a: every autoregressive step of decoder_joint.generate (seq2seq_pretrain.py:450;
// x-transformers AutoregressiveWrapper.generate / Decoder) inside ONE cooperative launch.
//
// Why: one decode step used to be 46 dependent kernel launches (CUDA graph); each paid ~4-5 us of launch gap, barrier/TMEM
// prologue, first-TMA round trip and epilogue drain, so a 256-clip step took ~670 us against an HBM floor of ~240 us
// (profiles/r01z_graph_gap.txt).  Here the grid is one CTA per SM, resident for the whole generate call; the
// launches become PHASES of a small program interpreted by every CTA, separated by a grid-wide barrier (one atomic + an
// acquire spin, < 1 us).  mbarriers, TMEM and the TMA descriptors are set up once.
//
// Phase types (MkPhase::type):
//   MK_GEMM       split-K tcgen05 GEMM.  Work unit = (128-row tile of decode rows, bn-wide tile of output features, K slice z);
//                 units are dealt round-robin to the CTAs.  warp 0 = TMA producer (A and W tiles, 128B swizzle, 6-slot
//                 full/empty mbarrier ring), warp 1 = tcgen05.mma issuer (fp32 accumulator in TMEM), warps 4-7 move the
//                 accumulator TMEM -> smem, then all 12 warps store it as coalesced rows of the fp32 PARTIAL part[z][row][col].
//                 No epilogue arithmetic here: the consumer phase sums the K slices in a fixed order (deterministic, and a row's
//                 bits never depend on which other rows share the batch: the split factor depends on (N, K, planes) only).
//   MK_ATTN       decode attention, one (row, head) work item per 64-thread sub-group (6 per CTA), items of a CTA handed out
//                 through a shared-memory counter.  q and the step's new k, v rows are summed from the QKV partials, the new row is
//                 appended to the head-major cache, K then V tiles stream through a 2-stage cp.async ring (same algorithm as
//                 attn_decode_kernel in attention.cu, which stays the stand-alone / per-kernel-path version).
//   MK_ROW_RESLN  x += bias + sum_z part[z]; LayerNorm(x) -> bf16 planes (the next GEMM's A operand).  One warp per row.
//   MK_ROW_GELU   gelu_erf(bias + sum_z part[z]) -> bf16 planes (FF2's A operand).  Elementwise.
//   MK_ROW_SAMPLE logits = bias + sum_z part[z]; argmax or top-k / softmax / inverse-CDF draw (same arithmetic as
//                 sample_row in rowops.cu); token embedding -> x; layer 0's LayerNorm -> planes.  One warp per row.
//
// Every wait (mbarrier, grid barrier) is bounded: after ~4 s it traps, so a logic error surfaces as a launch failure instead
// of wedging the GPU.
#include <algorithm>
#include <cstdlib>

#include "decode_mk.cuh"
#include "kvio.cuh"
#include "tc_ptx.cuh"

namespace dimb {

namespace {

// Threads per CTA: 384 (12 warps, 6 attention sub-groups of 64 threads) or, for bf16 caches with the mma.sync attention items, 512
// (16 warps, 8 sub-groups: the attention phases are neither issue- nor HBM-bound with 6 -- 27 % of the issue slots, 64 % of the HBM
// peak -- so two more items in flight per SM shorten them, and 3072 items make 3 rounds of 1184 instead of 4 of 888).  The wide
// flavour is a separate instantiation limited to 128 registers; it leaves out the FFMA attention items (64 registers of q alone).
constexpr int MK_THREADS = 384, MK_THREADS_WIDE = 512;
constexpr int MK_MAX_WARPS = MK_THREADS_WIDE / 32;
constexpr int MK_MAX_NSUB = MK_THREADS_WIDE / 64;
#define MK_NTHREADS ((int)blockDim.x)
#define MK_NWARPS ((int)(blockDim.x >> 5))
constexpr int MK_STAGES = 6;
constexpr uint32_t MK_A_BYTES = 128 * 64 * 2;       // one A tile: 128 rows x 64 bf16 (128-byte swizzle rows)
constexpr uint32_t MK_SLOT = 2 * MK_A_BYTES;        // ring slot: A tile + W tile of up to 128 rows
constexpr uint32_t MK_RING = MK_STAGES * MK_SLOT;   // 192 KB, also the attention / staging scratch
constexpr int MK_TMEM_COLS = 128;
constexpr int MK_MAXS = 9;                          // largest split-K factor (engine.cu: mk_gemm_cfg)
constexpr unsigned long long MK_TIMEOUT_NS = 4000000000ull;

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// mbarrier wait with a watchdog
__device__ __forceinline__ void mbar_wait_guard(uint32_t bar, uint32_t parity) {
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
    if (it == 64) t0 = gtime_ns();
    if (it > 64 && (it & 255) == 0 && gtime_ns() - t0 > MK_TIMEOUT_NS) __trap();
  }
}

__device__ __forceinline__ void bar_sub(int sub) {   // named barrier of one 64-thread attention sub-group
  asm volatile("bar.sync %0, 64;" ::"r"(sub + 1) : "memory");
}

// Grid-wide barrier: CTA barrier, then thread 0 publishes with a gpu-scope RELEASE reduction (cumulative over the CTA's writes
// through the barrier before it) and spins with gpu-scope ACQUIRE loads; a second CTA barrier releases the other threads.
// (A variant in which the last arriver publishes a generation flag on its own line and everyone polls that line measured
// SLOWER: +1.2 us per barrier -- the extra release store and its propagation cost more than the contention they avoid.)
// to_tma: the next phase reads this phase's generic-proxy global writes through TMA (async proxy) -> proxy fences on both sides.
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int& target, bool to_tma) {
  if (to_tma) asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if ((int)(v - target) >= 0) break;
      if (it == 64) t0 = gtime_ns();
      if (it > 64 && (it & 63) == 0 && gtime_ns() - t0 > MK_TIMEOUT_NS) __trap();
    }
  }
  __syncthreads();
  if (to_tma) asm volatile("fence.proxy.async;" ::: "memory");
}

struct Pipe {            // ring / accumulator barrier positions; they persist across units, phases and steps
  uint32_t p_stage = 0, p_phase = 0;     // producer
  uint32_t c_stage = 0, c_phase = 0;     // MMA issuer
  uint32_t acc_phase = 0;                // accumulator barrier parity (all threads track it)
};

struct Ctl {             // static shared memory
  uint64_t full_bar[MK_STAGES];
  uint64_t empty_bar[MK_STAGES];
  uint64_t accum_bar;
  uint32_t tmem_slot;
  int attn_ctr;
  int sub_item[MK_MAX_NSUB];
  float red[2 * MK_MAX_WARPS];           // block reductions of the row phases
  int tok[2];
};

// ---- MK_GEMM ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mk_gemm(const MkPlan& P, const MkPhase& ph, uint8_t* smem, uint32_t smem_base, Ctl& ctl,
                                        Pipe& pipe, uint32_t tmem_base) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = ph.M, N = ph.N, bn = ph.bn, S = ph.splits;
  const int mt = (M + 127) >> 7, nt = (N + bn - 1) / bn;
  const int units = mt * nt * S;
  const int total_it = ph.kblocks * ph.npairs;
  const uint32_t stage_bytes = MK_A_BYTES + (uint32_t)bn * 128u;
  const int CP = bn + 4;                                // fp32 staging pitch (floats): conflict-free 128-bit accesses
  float* stage_tile = reinterpret_cast<float*>(smem);
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    // the m tiles that share a weight tile are neighbours in the unit order: they run at the same time on different SMs and the
    // second one finds the tile in L2
    const int m_i = u % mt, r1 = u / mt, n_i = r1 % nt, z = r1 / nt;
    const int m0 = m_i * 128, n0 = n_i * bn;
    const int it_begin = (int)(((long long)z * total_it) / S), it_end = (int)(((long long)(z + 1) * total_it) / S);
    if (warp == 0) {
      if (lane == 0) {
        const CUtensorMap* tmA = &P.maps[ph.mapA];
        const CUtensorMap* tmW = &P.maps[ph.mapW];
        uint64_t wpol = 0;
        const bool hint = ph.w_keep > 0.f;
        if (hint) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(wpol) : "f"(ph.w_keep));
        uint32_t stage = pipe.p_stage, phase = pipe.p_phase;
        for (int it = it_begin; it < it_end; ++it) {
          const int pair = it / ph.kblocks, kb = it - pair * ph.kblocks;
          mbar_wait_guard(smem_u32(&ctl.empty_bar[stage]), phase ^ 1u);
          const uint32_t fb = smem_u32(&ctl.full_bar[stage]);
          mbar_expect_tx(fb, stage_bytes);
          const uint32_t sa = smem_base + stage * MK_SLOT;
          if (hint) tma_load_2d_hint(sa + MK_A_BYTES, tmW, ph.pw[pair] * ph.kp + kb * 64, n0, fb, wpol);
          else tma_load_2d(sa + MK_A_BYTES, tmW, ph.pw[pair] * ph.kp + kb * 64, n0, fb);
          tma_load_2d(sa, tmA, ph.pa[pair] * ph.kp + kb * 64, m0, fb);
          if (++stage == MK_STAGES) { stage = 0; phase ^= 1u; }
        }
        pipe.p_stage = stage; pipe.p_phase = phase;
      }
      __syncwarp();
    } else if (warp == 1) {
      if (lane == 0) {
        // instruction descriptor: D fp32 (1<<4), A bf16 (1<<7), B bf16 (1<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint32_t stage = pipe.c_stage, phase = pipe.c_phase;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");     // the previous unit's TMEM reads are done (CTA barrier)
        for (int it = it_begin; it < it_end; ++it) {
          mbar_wait_guard(smem_u32(&ctl.full_bar[stage]), phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + stage * MK_SLOT;
          const uint64_t adesc = make_sdesc(sa), bdesc = make_sdesc(sa + MK_A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)                   // UMMA_K = 16 bf16 = 32 bytes: +2 in the (addr >> 4) field
            umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it > it_begin || k > 0) ? 1u : 0u);
          umma_commit(smem_u32(&ctl.empty_bar[stage]));
          if (++stage == MK_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(smem_u32(&ctl.accum_bar));
        pipe.c_stage = stage; pipe.c_phase = phase;
      }
      __syncwarp();
    } else if (warp >= 4 && warp < 8) {
      // accumulator -> staging tile (the ring is idle once accum_bar fires: every MMA, hence every TMA write, has completed)
      const int q = warp & 3;                            // a warp may only touch TMEM lanes 32*(warp%4) .. +31
      mbar_wait_guard(smem_u32(&ctl.accum_bar), pipe.acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* my_row = stage_tile + (size_t)(q * 32 + lane) * CP;
      const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < bn; c0 += 16) {
        float v[16];
        tmem_ld16(tlane + (uint32_t)c0, v);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<float4*>(my_row + c0 + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    pipe.acc_phase ^= 1u;
    __syncthreads();
    {  // staging tile -> part[z][m0 + r][n0 + c], whole rows per warp instruction
      const int LPR = bn >> 2, RPI = 32 / LPR;
      const int c = (lane % LPR) * 4, col = n0 + c;
      const int rows = min(128, M - m0);
      if (ph.epi == 1) {
        // unsplit GEMM with the consumer's row phase folded in: gelu_erf(acc + bias) -> bf16 planes (FF1 -> FF2's A operand)
        if (col < N) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ph.bias) bb = __ldg(reinterpret_cast<const float4*>(ph.bias + col));
          for (int r = warp * RPI + lane / LPR; r < rows; r += MK_NWARPS * RPI) {
            float4 v = *reinterpret_cast<const float4*>(stage_tile + (size_t)r * CP + c);
            v.x = act_apply(v.x + bb.x, DIM_ACT_GELU_ERF, 0.f); v.y = act_apply(v.y + bb.y, DIM_ACT_GELU_ERF, 0.f);
            v.z = act_apply(v.z + bb.z, DIM_ACT_GELU_ERF, 0.f); v.w = act_apply(v.w + bb.w, DIM_ACT_GELU_ERF, 0.f);
            store_planes4(ph.outp + (size_t)(m0 + r) * P.planes * ph.out_kp + col, v, P.planes, ph.out_kp);
          }
        }
      } else {
        float* dst = ph.part + ((size_t)z * M + m0) * N + col;
        if (col < N)
          for (int r = warp * RPI + lane / LPR; r < rows; r += MK_NWARPS * RPI)
            *reinterpret_cast<float4*>(dst + (size_t)r * N) = *reinterpret_cast<const float4*>(stage_tile + (size_t)r * CP + c);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic accesses to the ring before the next TMA writes
    __syncthreads();
  }
}

// ---- MK_GEMM for <= 8 decode rows: GEMV ------------------------------------------------------------------------------------
// A decode step of a single clip is one weight stream (144 MB in bf16 mode) against a handful of activation rows: tiles and tensor
// cores have nothing to offer.  Every warp of the grid owns whole output columns n = gw, gw + #warps, ...; the M activation rows are
// staged once per CTA in shared memory as fp32 (the sum of their bf16 planes, which is exact), the weight row streams through
// registers with 128-bit loads -- the first batch is requested BEFORE the staging, so the HBM round trip and the L2 round trip
// overlap -- and a warp reduction yields part[0][m][n] (or, epi == 1, gelu(acc + bias) as the next GEMV's bf16 planes).
// bf16 mode reads the bf16 weight image; the fp32-grade mode reads the original fp32 weights (fewer bytes than three planes).
template <bool W16>
__device__ __forceinline__ void mk_gemv(const MkPlan& P, const MkPhase& ph, uint8_t* smem) {
  constexpr int NB = 9;                                 // 128-bit weight vectors in flight per lane and batch
  constexpr int EPV = W16 ? 8 : 4;                      // weights per vector
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = ph.M, N = ph.N, kp = ph.kp;
  const int kw = W16 ? kp : ph.K;                       // weights per row actually streamed
  const int nvec = kw / EPV;
  float* a_s = reinterpret_cast<float*>(smem);          // [M][kp]
  const int gw = (int)blockIdx.x * MK_NWARPS + warp, nw = (int)gridDim.x * MK_NWARPS;
  const uint4* wrow0 = W16 ? reinterpret_cast<const uint4*>(ph.wb) : reinterpret_cast<const uint4*>(ph.w32);
  const size_t row_vecs = (size_t)(W16 ? kp : ph.K) / EPV;
  uint4 wv[NB];
  auto load_batch = [&](int n, int b0) {
    const uint4* wr = wrow0 + (size_t)n * row_vecs;
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int j = b0 + u * 32 + lane;
      wv[u] = j < nvec ? __ldcs(wr + j) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  int n = gw;
  if (n < N) load_batch(n, 0);
  {  // stage A: fp32 = sum of the planes in plane order (h + m, then + l: both sums exact)
    const int v8 = kp / 8;
    for (int i = threadIdx.x; i < M * v8; i += MK_NTHREADS) {
      const int m = i / v8, j = i - m * v8;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int pl = 0; pl < P.planes; ++pl) {
        const uint4 q = *reinterpret_cast<const uint4*>(ph.a_planes + ((size_t)m * P.planes + pl) * kp + j * 8);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) { f[2 * t] += __uint_as_float(w[t] << 16); f[2 * t + 1] += __uint_as_float(w[t] & 0xffff0000u); }
      }
      float4* d = reinterpret_cast<float4*>(a_s + (size_t)m * kp + j * 8);
      d[0] = make_float4(f[0], f[1], f[2], f[3]);
      d[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
  }
  __syncthreads();
  for (bool first = true; n < N; n += nw, first = false) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int b0 = 0; b0 < nvec; b0 += NB * 32) {
      if (!(first && b0 == 0)) load_batch(n, b0);
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int j = b0 + u * 32 + lane;
        if (j < nvec) {
          float wf[EPV];
          const uint32_t w[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w};
          if (W16) {
#pragma unroll
            for (int t = 0; t < 4; ++t) { wf[2 * t] = __uint_as_float(w[t] << 16); wf[(2 * t + 1) % EPV] = __uint_as_float(w[t] & 0xffff0000u); }
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) wf[t % EPV] = __uint_as_float(w[t]);
          }
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            if (m < M) {
              const float4* ar = reinterpret_cast<const float4*>(a_s + (size_t)m * kp + (size_t)j * EPV);
              const float4 a0 = ar[0];
              acc[m] = fmaf(a0.x, wf[0], acc[m]); acc[m] = fmaf(a0.y, wf[1], acc[m]);
              acc[m] = fmaf(a0.z, wf[2], acc[m]); acc[m] = fmaf(a0.w, wf[3], acc[m]);
              if (W16) {
                const float4 a1 = ar[1];
                acc[m] = fmaf(a1.x, wf[4 % EPV], acc[m]); acc[m] = fmaf(a1.y, wf[5 % EPV], acc[m]);
                acc[m] = fmaf(a1.z, wf[6 % EPV], acc[m]); acc[m] = fmaf(a1.w, wf[7 % EPV], acc[m]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 8; ++m)
      if (m < M) acc[m] = warp_sum(acc[m]);
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        if (m >= M) continue;
        if (ph.epi == 1) {
          const float v = act_apply(acc[m] + (ph.bias ? __ldg(ph.bias + n) : 0.f), DIM_ACT_GELU_ERF, 0.f);
          store_planes1(ph.outp + (size_t)m * P.planes * ph.out_kp + n, v, P.planes, ph.out_kp);
        } else {
          ph.part[(size_t)m * N + n] = acc[m];
        }
      }
    }
  }
  __syncthreads();                                       // a_s is the next phase's scratch
}

// ---- MK_ATTN ------------------------------------------------------------------------------------------------------------
template <bool BF16>
__device__ __forceinline__ void mk_attn_item(const MkPhase& ph, int Brows, int H, int planes, int sc_floats, uint8_t* reg, int sub,
                                             int tid, int item, int pos) {
  typedef typename KvIo<BF16>::T KT;
  constexpr int NT = 64, DH = 64, EPL = KvIo<BF16>::EPL, LPK = DH / EPL, NG = NT / LPK, CHUNK = BF16 ? 64 : 32;
  constexpr int ROWB = DH * (int)sizeof(KT), PITCH = ROWB + 16, STAGE = CHUNK * PITCH, KPG = CHUNK / NG;
  float* sc = reinterpret_cast<float*>(reg + 2 * STAGE);
  float* part = sc + sc_floats;                        // [NG][64]
  float* qs = part + NG * DH;                          // [64]
  float* red = qs + DH;                                // [4]
  float* knew = red + 4;                               // [64] this step's key row (append)
  float* vnew = knew + DH;                             // [64]
  const int lane = tid & 31, warp = tid >> 5;
  const int grp = tid / LPK, lk = tid % LPK;
  const int b = item / H, h = item - b * H;
  const int nkeys = ph.append ? pos + 1 : ph.Tk;     // keys attended, the step's own key included
  const int bkv = b / ph.kv_group;
  KT* khead = static_cast<KT*>(ph.kcache) + (size_t)bkv * ph.kv_batch_stride + (size_t)h * ph.kv_head_stride;
  KT* vhead = static_cast<KT*>(ph.vcache) + (size_t)bkv * ph.kv_batch_stride + (size_t)h * ph.kv_head_stride;
  const uint32_t ring_u = smem_u32(reg);
  // keys already in the cache; the key/value of THIS step (append) never takes the round trip through global memory: it is
  // written to the cache for the later steps and used here from shared memory (rounded to the cache's element type first)
  const int nold = ph.append ? pos : ph.Tk;
  const int nch = (nold + CHUNK - 1) / CHUNK;
  const int kval = (ph.key_mask && ph.key_valid) ? __ldg(ph.key_valid + bkv) : -1;      // requested now, used in the softmax
  uint64_t kvpol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(kvpol));
  auto issue = [&](const KT* head, int c, int stage) {                   // rows [c*CHUNK, ..) of a head block -> ring[stage]
    const int row0 = c * CHUNK, pieces = min(CHUNK, nold - row0) * LPK;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(head + (size_t)row0 * DH);
    const uint32_t dst = ring_u + stage * STAGE;
    for (int q = tid; q < pieces; q += NT) cp_async_16_hint(dst + (q / LPK) * PITCH + (q % LPK) * 16, src + (size_t)q * 16, kvpol);
    cp_async_commit();
  };
  if (nch > 0) issue(khead, 0, 0);                     // the first K tiles travel while the projection partials are summed
  if (nch > 1) issue(khead, 1, 1);
  {  // q (and this step's k, v) = sum of the K slices of the projection, in slice order; thread t owns element t of the head row.
     // All loads are issued before the first add (one L2 round trip, not one per slice).
    const size_t mn = (size_t)Brows * ph.q_ld;
    const float* base = reinterpret_cast<const float*>(ph.part) + (size_t)b * ph.q_ld + h * DH + tid;
    float pq[MK_MAXS], pk[MK_MAXS], pv[MK_MAXS];
#pragma unroll
    for (int z = 0; z < MK_MAXS; ++z) {
      const bool on = z < ph.q_splits;
      pq[z] = on ? base[z * mn + ph.q_col] : 0.f;
      pk[z] = (on && ph.append) ? base[z * mn + ph.k_col] : 0.f;
      pv[z] = (on && ph.append) ? base[z * mn + ph.v_col] : 0.f;
    }
    float qv = pq[0], kv = pk[0], vv = pv[0];
#pragma unroll
    for (int z = 1; z < MK_MAXS; ++z)
      if (z < ph.q_splits) { qv += pq[z]; kv += pk[z]; vv += pv[z]; }
    qs[tid] = qv;
    if (ph.append) {
      if (BF16) {
        const __nv_bfloat16 kb = __float2bfloat16_rn(kv), vb = __float2bfloat16_rn(vv);
        reinterpret_cast<__nv_bfloat16*>(khead)[(size_t)pos * DH + tid] = kb;
        reinterpret_cast<__nv_bfloat16*>(vhead)[(size_t)pos * DH + tid] = vb;
        kv = __bfloat162float(kb);
        vv = __bfloat162float(vb);
      } else {
        reinterpret_cast<float*>(khead)[(size_t)pos * DH + tid] = kv;
        reinterpret_cast<float*>(vhead)[(size_t)pos * DH + tid] = vv;
      }
      knew[tid] = kv;
      vnew[tid] = vv;
    }
  }
  bar_sub(sub);

  float q[DH];
#pragma unroll
  for (int i = 0; i < DH; i += 4) {
    const float4 x = *reinterpret_cast<const float4*>(qs + i);
    q[i] = x.x; q[i + 1] = x.y; q[i + 2] = x.z; q[i + 3] = x.w;
  }
  if (ph.append && tid == NT - 1) {                    // score of the step's own key, from shared memory
    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < DH; ++i) d[i & 3] = fmaf(q[i], knew[i], d[i & 3]);
    sc[pos] = ((d[0] + d[1]) + (d[2] + d[3])) * ph.scale;
  }
  // ---- pass 1: scores of the cached keys, one thread per key
  for (int c = 0; c < nch; ++c) {
    if (c + 1 < nch) cp_async_wait<1>(); else cp_async_wait<0>();
    bar_sub(sub);
    const int j = c * CHUNK + tid;
    if (tid < CHUNK && j < nold) {
      const uint4* row = reinterpret_cast<const uint4*>(reg + (c & 1) * STAGE + tid * PITCH);
      float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < LPK; ++i) d[i & 3] += KvIo<BF16>::dot(row[i], q + i * EPL);
      sc[j] = ((d[0] + d[1]) + (d[2] + d[3])) * ph.scale;
    }
    bar_sub(sub);
    if (c + 2 < nch) issue(khead, c + 2, c & 1);
  }
  if (nch > 0) issue(vhead, 0, 0);
  if (nch > 1) issue(vhead, 1, 1);
  bar_sub(sub);                                        // the own key's score is visible (covers the no-tile first step too)
  // key-padding mask (masked_fill(-finfo.max)) and softmax statistics
  const uint8_t* km = ph.key_mask ? ph.key_mask + (size_t)bkv * ph.Tk : nullptr;
  float mx = -INFINITY;
  for (int j = tid; j < nkeys; j += NT) {
    float v = sc[j];
    const bool drop = kval >= 0 ? j >= kval : (km && !km[j]);      // prefix mask: one length per clip, no bytes
    if (drop) { v = -FLT_MAX; sc[j] = v; }
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  bar_sub(sub);
  mx = fmaxf(red[0], red[1]);
  bar_sub(sub);
  float sum = 0.f;
  for (int j = tid; j < nkeys; j += NT) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  bar_sub(sub);
  const float inv = 1.f / (red[0] + red[1]);
  // ---- pass 2: out = P V, a group of LPK lanes per key
  float acc[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) acc[i] = 0.f;
  for (int c = 0; c < nch; ++c) {
    if (c + 1 < nch) cp_async_wait<1>(); else cp_async_wait<0>();
    bar_sub(sub);
    const uint8_t* tile = reg + (c & 1) * STAGE + lk * 16;
#pragma unroll
    for (int i = 0; i < KPG; ++i) {
      const int jl = grp + NG * i, j = c * CHUNK + jl;
      if (j < nold) KvIo<BF16>::axpy(*reinterpret_cast<const uint4*>(tile + jl * PITCH), sc[j], acc);
    }
    bar_sub(sub);
    if (c + 2 < nch) issue(vhead, c + 2, c & 1);
  }
#pragma unroll
  for (int i = 0; i < EPL; ++i) part[grp * DH + lk * EPL + i] = acc[i];
  bar_sub(sub);
  {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < NG; ++w) r += part[w * DH + tid];
    if (ph.append) r = fmaf(sc[pos], vnew[tid], r);
    store_planes1(ph.outp + (size_t)b * planes * ph.out_kp + h * DH + tid, r * inv, planes, ph.out_kp);
  }
}

// ---- MK_ATTN, bf16 caches: the two products of an item on the (legacy) tensor pipe -----------------------------------------
// The FFMA item above spends ~14 warp instructions per key and is issue-bound with the 12 warps of this kernel (51.8 us per
// 256-clip cross-attention phase against 36 us of HBM time).  With bf16 K/V rows the dot products map onto mma.sync m16n8k16:
//   S = q K^T : A = q (row 0: bf16 high part, row 1: low part -- q = hi + lo to 2^-17, the other 14 rows are zero),
//               B = a K tile straight from the padded smem rows (ldmatrix, conflict-free with the 144-byte pitch);
//   O = P V   : A = p (rows 0/1: high / low part of the fp32 probabilities), B = the V tile (ldmatrix.trans).
// fp32 accumulation; rows 0 and 1 of the result are added with one shuffle.  ~4 warp instructions per key.  Warp w of the
// sub-group owns keys [32w, 32w+32) of every 64-key tile.  Rows of a tile past the last cached key are ZERO-filled (cp.async
// src-size 0): the tensor core multiplies whole tiles, and 0 * stale-NaN would poison the sum.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst_smem, const void* src, uint32_t src_bytes, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(dst_smem), "l"(src), "r"(src_bytes), "l"(policy)
               : "memory");
}
// (x, y) -> packed bf16x2 of the high parts (sel == 0), of the low parts x - bf16(x) (sel == 1), or 0
__device__ __forceinline__ uint32_t pack_part(float x, float y, int sel) {
  __nv_bfloat16 bx = __float2bfloat16_rn(x), by = __float2bfloat16_rn(y);
  if (sel == 1) {
    bx = __float2bfloat16_rn(x - __bfloat162float(bx));
    by = __float2bfloat16_rn(y - __bfloat162float(by));
  }
  const uint32_t w = (uint32_t)__bfloat16_as_ushort(bx) | ((uint32_t)__bfloat16_as_ushort(by) << 16);
  return sel <= 1 ? w : 0u;
}

// rows [c*64, c*64+64) of a bf16 head block -> one ring slot (64 rows of 144 bytes).  WARP w of the sub-group copies rows 32w..32w+31
// -- exactly the keys whose scores and PV products it computes -- so a tile needs no barrier between the two warps: a lane waits for
// its own cp.async groups and a __syncwarp publishes them to the warp.  Lane -> 8 pieces of 16 bytes: rows 32w + lane/8 + 4i, chunk
// lane%8 (a warp instruction covers 4 whole rows = 512 contiguous bytes).  Rows past the last cached key (>= nold) are zero-filled.
__device__ __forceinline__ void mk_issue_kv_tile(const __nv_bfloat16* head, int c, int nold, uint32_t slot_addr, int tid, uint64_t kvpol) {
  const int row0 = c * 64, valid = nold - row0;
  const int r0 = (tid >> 5) * 32 + ((tid & 31) >> 3), ch = tid & 7;
  const uint8_t* src = reinterpret_cast<const uint8_t*>(head + (size_t)(row0 + r0) * 64) + ch * 16;
  const uint32_t dst = slot_addr + (uint32_t)(r0 * 144 + ch * 16);
  if (valid >= 64) {                                   // whole tile cached (all but the last tile of a block): no per-piece predicates
#pragma unroll
    for (int i = 0; i < 8; ++i) cp_async_16_hint(dst + i * 4 * 144, src + i * 512, kvpol);
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool ok = r0 + 4 * i < valid;
    cp_async_16_zfill(dst + i * 4 * 144, ok ? src + i * 512 : reinterpret_cast<const uint8_t*>(head), ok ? 16u : 0u, kvpol);
  }
}

// The first tiles of every sub-group's FIRST item do not depend on the projection GEMM that precedes an attention phase (they are
// cached keys of earlier steps / of the context): at the END of that GEMM phase one lane per sub-group asks for them with an L2
// bulk prefetch, so they come from HBM during the grid barrier and the cp.async stream after it starts on L2 hits.  (Requesting the
// tiles themselves with cp.async before the barrier was measured and lost: the barrier's gpu-scope release waits for the CTA's pending
// cp.async, +2.5-3 us per projection phase against -2.6 us per attention phase; profiles/r02_notes.md.)  The first nsub entries of a
// CTA's item list are therefore assigned statically (entry s to sub-group s); the rest are still handed out through the shared counter.
__device__ __forceinline__ void mk_attn_prefetch(const MkPlan& P, const MkPhase& ph, int pos) {
  const int sub = threadIdx.x >> 6, tid = threadIdx.x & 63;
  if (sub >= P.attn_nsub || tid >= 2) return;          // lane 0: K block, lane 1: V block
  const int items = P.B * P.H;
  const int mine = (int)blockIdx.x < items ? (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int nold = ph.append ? pos : ph.Tk;
  if (sub >= mine || nold <= 0) return;
  const int item = (int)blockIdx.x + sub * (int)gridDim.x;
  const int b = item / P.H, h = item - b * P.H, bkv = b / ph.kv_group;
  const size_t off = (size_t)bkv * ph.kv_batch_stride + (size_t)h * ph.kv_head_stride;
  const __nv_bfloat16* head = static_cast<const __nv_bfloat16*>(tid == 0 ? ph.kcache : ph.vcache) + off;
  const int rows = min(nold, tid == 0 ? P.attn_pre : P.attn_pre / 2);       // rows of 128 bytes
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(head), "r"(rows * 128) : "memory");
}

// K/V of the next attention phase -> L2 while a latency-bound phase runs.  The GEMM / row / sampling phases of a step take ~280 us
// during which HBM carries only the 144 MB of weights; the attention phases are the HBM-bound part.  Every non-attention phase
// therefore starts by asking (bulk L2 prefetch, no shared-memory destination, nothing to wait for) for the K and V blocks of the
// first entries of this CTA's item list of the NEXT attention phase -- the entries its sub-groups take first -- so that the
// attention phase begins on L2 hits and streams only the rest from HBM.  The byte budget per CTA (plan.pf_budget) is spread over
// the phases of the window by the host (pf_f0 .. pf_f1).  Issued by the CTA's last two warps, which have no producer / MMA role.
__device__ __forceinline__ void mk_kv_prefetch(const MkPlan& P, const MkPhase& ph, int phase_idx, int st) {
  const int t = (int)threadIdx.x - ((int)blockDim.x - 64);
  if (t < 0) return;
  const MkPhase& at = P.phases[ph.pf_target];
  const int pos = st + (ph.pf_target < phase_idx ? 1 : 0);       // a target earlier in the program belongs to the next step
  if (pos >= P.steps) return;
  const int nold = at.append ? pos : at.Tk;
  if (nold <= 0) return;
  const int items = P.B * P.H;
  const int mine = (int)blockIdx.x < items ? (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const unsigned esize = P.kv_bf16 ? 2u : 4u, rowb = 64u * esize;
  const unsigned long long per_item = 2ull * (unsigned)nold * rowb;
  const int E = (int)min((unsigned long long)mine, P.pf_budget / per_item);
  const int k0 = (int)(ph.pf_f0 * (float)E), k1 = (int)(ph.pf_f1 * (float)E);
  for (int j = t; j < 2 * (k1 - k0); j += 64) {
    const int item = (int)blockIdx.x + (k0 + (j >> 1)) * (int)gridDim.x;
    const int b = item / P.H, h = item - b * P.H, bkv = b / at.kv_group;
    const size_t off = ((size_t)bkv * at.kv_batch_stride + (size_t)h * at.kv_head_stride) * esize;
    const uint8_t* src = static_cast<const uint8_t*>((j & 1) ? at.vcache : at.kcache) + off;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((unsigned)nold * rowb) : "memory");
  }
}

// The K tiles then the V tiles of an item form one stream of 2*nch tiles through an NST-slot cp.async ring, and the stream runs on
// into the NEXT item of the sub-group (every item of a phase has the same tile count): while an item finishes (softmax tail, output
// reduction, store) and the next one sums its projection partials, the next item's first tiles are already in flight.  `issued`
// counts the cp.async groups committed by this sub-group in this phase (one per stream position, empty past the last item), `g0` is
// the stream position of this item's first tile; slot = position % NST.
template <int NST>
__device__ __forceinline__ void mk_attn_item_mma(const MkPhase& ph, int Brows, int H, int planes, int sc_floats, uint8_t* reg, int sub,
                                                 int tid, int item, int next_item, int pos, int& issued, int& g0, int dbg) {
  constexpr int NT = 64, DH = 64, CHUNK = 64, PITCH = 144, STAGE = CHUNK * PITCH;
  float* sc = reinterpret_cast<float*>(reg + NST * STAGE);
  float* part = sc + sc_floats;                        // [2][64] partial outputs of the two warps
  float* qs = part + 2 * DH;                           // [64]
  float* red = qs + DH;                                // [4]
  float* knew = red + 4;                               // [64] this step's key row (append)
  float* vnew = knew + DH;                             // [64]
  const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  const int b = item / H, h = item - b * H;
  const int nkeys = ph.append ? pos + 1 : ph.Tk;       // keys attended, the step's own key included
  const int nold = ph.append ? pos : ph.Tk;            // keys already in the cache
  const int bkv = b / ph.kv_group;
  __nv_bfloat16* khead = static_cast<__nv_bfloat16*>(ph.kcache) + (size_t)bkv * ph.kv_batch_stride + (size_t)h * ph.kv_head_stride;
  __nv_bfloat16* vhead = static_cast<__nv_bfloat16*>(ph.vcache) + (size_t)bkv * ph.kv_batch_stride + (size_t)h * ph.kv_head_stride;
  const uint32_t ring_u = smem_u32(reg);
  const int nch = (nold + CHUNK - 1) / CHUNK;
  uint64_t kvpol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(kvpol));
  const int n = 2 * nch;                                // tiles of one item's stream
  // head blocks of the next item (its first tiles are prefetched at the end of this one): computed once, not per tile
  const __nv_bfloat16 *knext = khead, *vnext = vhead;
  if (next_item >= 0) {
    const int nb = next_item / H, nh = next_item - nb * H, nbkv = nb / ph.kv_group;
    const size_t off = (size_t)nbkv * ph.kv_batch_stride + (size_t)nh * ph.kv_head_stride;
    knext = static_cast<const __nv_bfloat16*>(ph.kcache) + off;
    vnext = static_cast<const __nv_bfloat16*>(ph.vcache) + off;
  }
  auto issue_tile = [&](const __nv_bfloat16* head, int c, int slot) {     // rows [c*64, c*64+64) of a head block -> ring[slot]
    if (!(dbg & 1)) mk_issue_kv_tile(head, c, nold, ring_u + slot * STAGE, tid, kvpol);      // dbg bit 0: timing ablation without the K/V loads
  };
  int slot_next = issued % NST;                         // ring slot of stream position `issued`
  auto issue_upto = [&](int limit) {                    // commit one group per stream position below `limit`
    while (issued < limit) {
      const int idx = issued - g0;
      if (idx < nch) issue_tile(khead, idx, slot_next);
      else if (idx < n) issue_tile(vhead, idx - nch, slot_next);
      else if (next_item >= 0 && idx < n + nch) issue_tile(knext, idx - n, slot_next);
      else if (next_item >= 0 && idx < 2 * n) issue_tile(vnext, idx - n - nch, slot_next);
      cp_async_commit();
      ++issued;
      if (++slot_next == NST) slot_next = 0;
    }
  };
  // key-padding mask: prefix masks (the usual case, key_valid[clip] >= 0) need no bytes at all -- one int per clip, requested here and
  // first looked at after the projection partials have been requested (so its round trip hides behind theirs)
  const uint8_t* km = ph.key_mask ? ph.key_mask + (size_t)bkv * ph.Tk : nullptr;
  const int kval = (km && ph.key_valid && !(dbg & 8)) ? __ldg(ph.key_valid + bkv) : -1;
  uint32_t kmbits = 0xffffffffu;                       // arbitrary masks: bit i = key tid + 64 i is kept
  int slot_cur = g0 % NST;                             // ring slot of this item's next tile to consume
  issue_upto(g0 + NST);                                // the first tiles travel while the projection partials are summed (a no-op
                                                       // when the previous item of this sub-group has already sent them)
  {  // q (and this step's k, v) = sum of the K slices of the projection, in slice order; thread t owns element t of the head row
    const size_t mn = (size_t)Brows * ph.q_ld;
    const float* base = reinterpret_cast<const float*>(ph.part) + (size_t)b * ph.q_ld + h * DH + tid;
    const float *bq = base + ph.q_col, *bk = base + ph.k_col, *bv = base + ph.v_col;
    const int S = ph.q_splits;
    const bool app = ph.append != 0;
    float pq[MK_MAXS], pk[MK_MAXS], pv[MK_MAXS];
#pragma unroll
    for (int z = 0; z < MK_MAXS; ++z) {                 // all loads before the first add; one pointer step per slice
      const bool on = z < S && !(dbg & 4);             // dbg bit 2: without the projection partials
      pq[z] = on ? *bq : 0.f;
      pk[z] = (on && app) ? *bk : 0.f;
      pv[z] = (on && app) ? *bv : 0.f;
      bq += mn; bk += mn; bv += mn;
    }
    float qv = pq[0], kv = pk[0], vv = pv[0];
#pragma unroll
    for (int z = 1; z < MK_MAXS; ++z)
      if (z < S) { qv += pq[z]; kv += pk[z]; vv += pv[z]; }
    qs[tid] = qv;
    if (ph.append) {
      const __nv_bfloat16 kb = __float2bfloat16_rn(kv), vb = __float2bfloat16_rn(vv);
      khead[(size_t)pos * DH + tid] = kb;
      vhead[(size_t)pos * DH + tid] = vb;
      knew[tid] = __bfloat162float(kb);
      vnew[tid] = __bfloat162float(vb);
    }
  }
  if (km && kval < 0 && !(dbg & 8)) {                  // not a prefix mask: bytes of this thread's keys (tid, tid + 64, ...); dbg bit 3: none
    const int nw = min(32, (nkeys + 63) >> 6);
#pragma unroll 4
    for (int i = 0; i < nw; ++i)
      if (tid + 64 * i < nkeys && !km[tid + 64 * i]) kmbits &= ~(1u << i);
  }
  bar_sub(sub);
  // A fragments of q: lanes g == 0 carry the high parts (row 0), lanes g == 1 the low parts (row 1), every other row is zero
  uint32_t aq[4][4];
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    const float2 x0 = *reinterpret_cast<const float2*>(qs + kc * 16 + 2 * t4);
    const float2 x1 = *reinterpret_cast<const float2*>(qs + kc * 16 + 8 + 2 * t4);
    aq[kc][0] = pack_part(x0.x, x0.y, g);
    aq[kc][1] = 0u;
    aq[kc][2] = pack_part(x1.x, x1.y, g);
    aq[kc][3] = 0u;
  }
  if (ph.append && tid == NT - 1) {                    // score of the step's own key, from shared memory (fp32)
    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < DH; ++i) d[i & 3] = fmaf(qs[i], knew[i], d[i & 3]);
    sc[pos] = ((d[0] + d[1]) + (d[2] + d[3])) * ph.scale;
  }
  // per-lane ldmatrix offsets (bytes), as in attn_prefill_mma
  const uint32_t bk_off = (uint32_t)(((lane & 7) + (lane >> 4) * 8) * PITCH + ((lane >> 3) & 1) * 16);
  const uint32_t bv_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 16);
  // ---- pass 1: S = q K^T; warp w owns keys 32w .. 32w+31 of the tile
  for (int c = 0; c < nch; ++c) {
    cp_async_wait<NST - 1>();                          // the NST - 1 stream positions after this tile may still be in flight
    __syncwarp();                                      // this warp's rows of the tile (its own copies) are complete: no sub-group barrier
    const uint32_t tile = ring_u + slot_cur * STAGE;
    if (++slot_cur == NST) slot_cur = 0;
    // keys as the M dimension: A = a 16-key x 16-dim block of the K tile (ldmatrix, no transpose), B = q with column 0 = bf16 high
    // parts and column 1 = low parts (the other six columns are zero): 8 instead of 16 mma.sync per warp and tile, and the two parts
    // of a score land in ONE lane (t4 == 0: c0 + c1 for key g, c2 + c3 for key g + 8) -- no shuffle
    float s[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) s[mt][0] = s[mt][1] = s[mt][2] = s[mt][3] = 0.f;
    if (!(dbg & 2))                                    // dbg bit 1: timing ablation without the tensor-core products
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        uint32_t a[4];
        ldsm_x4(tile + (uint32_t)((warp * 32 + mt * 16) * PITCH + kc * 32) + bv_off, a[0], a[1], a[2], a[3]);
        mma_bf16(s[mt], a, aq[kc][0], aq[kc][2]);
      }
    }
    if (t4 == 0) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int j = c * CHUNK + warp * 32 + mt * 16 + g;
        if (j < nold) sc[j] = (s[mt][0] + s[mt][1]) * ph.scale;
        if (j + 8 < nold) sc[j + 8] = (s[mt][2] + s[mt][3]) * ph.scale;
      }
    }
    __syncwarp();                                      // every lane has read its rows: the slot may be refilled (by this warp only)
    issue_upto(g0 + c + 1 + NST);
  }
  bar_sub(sub);                                        // the own key's score is visible (covers the no-tile first step too)
  // key-padding mask (masked_fill(-finfo.max)) and softmax statistics
  float mx = -INFINITY;
  for (int j = tid, i = 0; j < nkeys; j += NT, ++i) {
    float v = sc[j];
    const bool keep = kval >= 0 ? j < kval : (i < 32 ? ((kmbits >> i) & 1u) != 0 : !(km && !km[j]));
    if (!keep) { v = -FLT_MAX; sc[j] = v; }
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  bar_sub(sub);
  mx = fmaxf(red[0], red[1]);
  bar_sub(sub);
  float sum = 0.f;
  for (int j = tid; j < nkeys; j += NT) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  bar_sub(sub);
  const float inv = 1.f / (red[0] + red[1]);
  // ---- pass 2: O = P V; warp w owns keys 32w .. 32w+31 of the tile (two 16-key k-steps), all 64 output columns
  // dims as the M dimension: A = a 16-dim x 16-key block of V^T (ldmatrix.trans of the V rows), B = p with column 0 = high parts and
  // column 1 = low parts: 8 instead of 16 mma.sync per warp and tile, 16 instead of 32 accumulator registers
  float o[4][4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f;
  for (int c = 0; c < nch; ++c) {
    cp_async_wait<NST - 1>();
    __syncwarp();
    const uint32_t tile = ring_u + slot_cur * STAGE;
    if (++slot_cur == NST) slot_cur = 0;
    if (!(dbg & 2))
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {
      const int kb = warp * 32 + k2 * 16, j0 = c * CHUNK + kb + 2 * t4;     // this lane's keys of the B operand: j0, j0+1, j0+8, j0+9
      const float p0 = j0 < nold ? sc[j0] : 0.f, p1 = j0 + 1 < nold ? sc[j0 + 1] : 0.f;
      const float p2 = j0 + 8 < nold ? sc[j0 + 8] : 0.f, p3 = j0 + 9 < nold ? sc[j0 + 9] : 0.f;
      const uint32_t bp0 = pack_part(p0, p1, g), bp1 = pack_part(p2, p3, g);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        uint32_t a[4];
        ldsm_x4_t(tile + (uint32_t)(kb * PITCH + mt * 32) + bk_off, a[0], a[1], a[2], a[3]);
        mma_bf16(o[mt], a, bp0, bp1);
      }
    }
    __syncwarp();
    issue_upto(g0 + nch + c + 1 + NST);                // past this item's last tile: the next item's first tiles
  }
  g0 += n;
  if (t4 == 0) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      part[warp * DH + mt * 16 + g] = o[mt][0] + o[mt][1];          // high part of p + low part
      part[warp * DH + mt * 16 + g + 8] = o[mt][2] + o[mt][3];
    }
  }
  bar_sub(sub);
  {
    float r = part[tid] + part[DH + tid];
    if (ph.append) r = fmaf(sc[pos], vnew[tid], r);
    store_planes1(ph.outp + (size_t)b * planes * ph.out_kp + h * DH + tid, r * inv, planes, ph.out_kp);
  }
}

template <int NST>
__device__ __forceinline__ void mk_attn_mma_loop(const MkPlan& P, const MkPhase& ph, uint8_t* reg, Ctl& ctl, int sub, int tid, int mine, int pos) {
  auto fetch = [&]() {                                 // next entry of this CTA's item list, handed out to whichever sub-group asks
    if (tid == 0) ctl.sub_item[sub] = atomicAdd(&ctl.attn_ctr, 1);
    bar_sub(sub);
    const int k = ctl.sub_item[sub];
    bar_sub(sub);
    return k;
  };
  // entry `sub` of the list is this sub-group's first item (mk_attn starts the counter at nsub): mk_attn_prefetch has asked for its
  // first tiles before the grid barrier
  int issued = 0, g0 = 0;
  int k = sub;
  while (k < mine) {
    const int kn = fetch();
    const int item = (int)blockIdx.x + k * (int)gridDim.x;
    const int next = kn < mine ? (int)blockIdx.x + kn * (int)gridDim.x : -1;
    mk_attn_item_mma<NST>(ph, P.B, P.H, P.planes, P.sc_floats, reg, sub, tid, item, next, pos, issued, g0, P.attn_dbg);
    bar_sub(sub);
    k = kn;
  }
  cp_async_wait<0>();                                  // only empty groups can be pending here
}

template <bool WIDE>
__device__ __forceinline__ void mk_attn(const MkPlan& P, const MkPhase& ph, uint8_t* smem, Ctl& ctl, int pos) {
  const int items = P.B * P.H;
  const int mine = (int)blockIdx.x < items ? (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const bool mma = WIDE || (P.kv_bf16 && P.attn_mma);
  if (threadIdx.x == 0) ctl.attn_ctr = mma ? P.attn_nsub : 0;      // mma items: the first nsub entries are assigned statically
  __syncthreads();
  const int sub = threadIdx.x >> 6, tid = threadIdx.x & 63;
  if (sub >= P.attn_nsub) return;                      // long key lists: fewer, larger scratch regions than 64-thread groups
  uint8_t* reg = smem + (size_t)sub * (MK_RING / (uint32_t)P.attn_nsub);
  if (mma) {
    if (P.attn_stages == 3) mk_attn_mma_loop<3>(P, ph, reg, ctl, sub, tid, mine, pos);
    else mk_attn_mma_loop<2>(P, ph, reg, ctl, sub, tid, mine, pos);
    return;
  }
  if constexpr (!WIDE) {
    for (;;) {
      if (tid == 0) ctl.sub_item[sub] = atomicAdd(&ctl.attn_ctr, 1);
      bar_sub(sub);
      const int k = ctl.sub_item[sub];
      if (k >= mine) break;
      const int item = (int)blockIdx.x + k * (int)gridDim.x;
      if (P.kv_bf16) mk_attn_item<true>(ph, P.B, P.H, P.planes, P.sc_floats, reg, sub, tid, item, pos);
      else mk_attn_item<false>(ph, P.B, P.H, P.planes, P.sc_floats, reg, sub, tid, item, pos);
      bar_sub(sub);
    }
  }
}

// ---- row phases ----------------------------------------------------------------------------------------------------------
// A row (<= 384 float4 = 1536 floats) is spread over the WHOLE CTA, one float4 per thread, and a CTA handles two rows at a
// time (rows r and r + gridDim.x: at 256 decode rows every CTA has at most two), so that every load of the phase -- all K
// slices of both rows -- is in flight before the first add: the phase costs one L2 round trip, not one per slice.

// block-wide sums of two values; every thread of the CTA must call it
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) { red[warp] = a; red[MK_MAX_WARPS + warp] = b; }
  __syncthreads();
  float sa = 0.f, sb = 0.f;
  const int nwarps = MK_NWARPS;
#pragma unroll 4
  for (int w = 0; w < nwarps; ++w) { sa += red[w]; sb += red[MK_MAX_WARPS + w]; }
  __syncthreads();
  a = sa;
  b = sb;
}

// sum of the K slices of one float4, slice order; all loads first
__device__ __forceinline__ float4 sum_slices(const float* p, size_t slice_stride, int splits) {
  float4 t[MK_MAXS];
#pragma unroll
  for (int z = 0; z < MK_MAXS; ++z)
    t[z] = z < splits ? *reinterpret_cast<const float4*>(p + z * slice_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 a = t[0];
#pragma unroll
  for (int z = 1; z < MK_MAXS; ++z)
    if (z < splits) { a.x += t[z].x; a.y += t[z].y; a.z += t[z].z; a.w += t[z].w; }
  return a;
}

// LayerNorm of two rows held one float4 per thread (thread c < n4 owns columns 4c..4c+3; on[k]: row k exists) -> bf16 planes
__device__ __forceinline__ void ln_pair(const float4 (&v)[2], const bool (&on)[2], int c, int n4, int dim, const float* gain,
                                        const float* beta, __nv_bfloat16* const (&out)[2], int planes, int kp, float* red) {
  const bool mine = c < n4;
  // gain / bias are requested before the two block reductions (loads do not move across __syncthreads on their own): their L2
  // round trip overlaps the reductions instead of following them
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f), bb = g;
  if (mine) {
    g = __ldg(reinterpret_cast<const float4*>(gain) + c);
    if (beta) bb = __ldg(reinterpret_cast<const float4*>(beta) + c);
  }
  float s0 = (mine && on[0]) ? (v[0].x + v[0].y) + (v[0].z + v[0].w) : 0.f;
  float s1 = (mine && on[1]) ? (v[1].x + v[1].y) + (v[1].z + v[1].w) : 0.f;
  block_sum2(s0, s1, red);
  const float mean[2] = {s0 / (float)dim, s1 / (float)dim};
  float q[2] = {0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (mine && on[k]) {
      const float a = v[k].x - mean[k], b = v[k].y - mean[k], cc = v[k].z - mean[k], d = v[k].w - mean[k];
      q[k] = (a * a + b * b) + (cc * cc + d * d);
    }
  block_sum2(q[0], q[1], red);
  if (!mine) return;
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (on[k]) {
      const float rstd = rsqrtf(q[k] / (float)dim + 1e-5f);
      float4 o;
      o.x = (v[k].x - mean[k]) * rstd * g.x; o.y = (v[k].y - mean[k]) * rstd * g.y;
      o.z = (v[k].z - mean[k]) * rstd * g.z; o.w = (v[k].w - mean[k]) * rstd * g.w;
      if (beta) { o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w; }
      store_planes4(out[k] + c * 4, o, planes, kp);
    }
}

__device__ __forceinline__ void mk_row_resln(const MkPlan& P, const MkPhase& ph, Ctl& ctl) {
  const int N = ph.N, n4 = N >> 2, c = threadIdx.x;
  const size_t mn = (size_t)ph.M * N;
  for (int r0 = (int)blockIdx.x; r0 < ph.M; r0 += 2 * (int)gridDim.x) {
    const int row[2] = {r0, r0 + (int)gridDim.x};
    const bool on[2] = {true, row[1] < ph.M};
    float4 v[2], res[2];
    v[0] = v[1] = res[0] = res[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < n4) {
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (on[k]) {
          res[k] = *reinterpret_cast<const float4*>(ph.x + (size_t)row[k] * N + c * 4);
          v[k] = sum_slices(ph.part + (size_t)row[k] * N + c * 4, mn, ph.in_splits);
        }
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ph.bias) bb = __ldg(reinterpret_cast<const float4*>(ph.bias) + c);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (on[k]) {
          float4 a = v[k];
          a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
          a.x = __fadd_rn(a.x, res[k].x); a.y = __fadd_rn(a.y, res[k].y); a.z = __fadd_rn(a.z, res[k].z); a.w = __fadd_rn(a.w, res[k].w);
          *reinterpret_cast<float4*>(ph.x + (size_t)row[k] * N + c * 4) = a;
          v[k] = a;
        }
    }
    __nv_bfloat16* const out[2] = {ph.outp + (size_t)row[0] * P.planes * ph.out_kp, ph.outp + (size_t)row[1] * P.planes * ph.out_kp};
    ln_pair(v, on, c, n4, N, ph.gain, ph.beta, out, P.planes, ph.out_kp, ctl.red);
  }
}

// gelu_erf(bias + sum of the K slices) -> planes.  U float4 per thread per batch, all loads of a batch in flight together.
__device__ __forceinline__ void mk_row_gelu(const MkPlan& P, const MkPhase& ph) {
  constexpr int U = 3;
  const int N = ph.N, n4 = N >> 2;
  const size_t mn = (size_t)ph.M * N;
  const int total = ph.M * n4, stride = (int)gridDim.x * MK_NTHREADS;
  for (int i0 = (int)blockIdx.x * MK_NTHREADS + (int)threadIdx.x; i0 < total; i0 += U * stride) {
    float4 a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * stride;
      if (i < total) a[u] = sum_slices(ph.part + (size_t)i * 4, mn, ph.in_splits);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * stride;
      if (i >= total) continue;
      const int row = i / n4, c = i - row * n4;
      float4 v = a[u];
      if (ph.bias) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(ph.bias) + c);
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
      }
      v.x = act_apply(v.x, DIM_ACT_GELU_ERF, 0.f); v.y = act_apply(v.y, DIM_ACT_GELU_ERF, 0.f);
      v.z = act_apply(v.z, DIM_ACT_GELU_ERF, 0.f); v.w = act_apply(v.w, DIM_ACT_GELU_ERF, 0.f);
      store_planes4(ph.outp + (size_t)row * P.planes * ph.out_kp + c * 4, v, P.planes, ph.out_kp);
    }
  }
}

__device__ __forceinline__ uint32_t order_key(float v) {      // unsigned key with the order of the floats (-0 == +0)
  const uint32_t u = __float_as_uint(v + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// One warp samples one row of logits held in shared memory.  Lane l owns the NPL = V/32 consecutive logits [l*NPL, (l+1)*NPL):
// index order is (lane, slot) order.  Same decisions as sample_row (rowops.cu): greedy = first maximal index; sampling = keep the
// top_k logits with ties toward the lower index, softmax(l / temperature) over them in fp32, inverse-CDF draw in index order
// with fp64 running sums whose association (chunks of V/256 entries, 8 chunks per lane, Hillis-Steele scan over lanes) equals
// the block kernel's.
template <int NPL>
__device__ __forceinline__ int mk_sample_row(const MkPlan& P, const float* lg, int row, int st, int lane) {
  constexpr int PER = NPL / 8;                           // entries per chunk of the block kernel (V / 256)
  constexpr int V = NPL * 32;
  float l[NPL];
#pragma unroll
  for (int j = 0; j < NPL; j += 4) {
    const float4 a = *reinterpret_cast<const float4*>(lg + lane * NPL + j);
    l[j] = a.x; l[j + 1] = a.y; l[j + 2] = a.z; l[j + 3] = a.w;
  }
  int tok;
  if (P.temperature == 0.f) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int i = lane * NPL + j;
      if (l[j] > best || (l[j] == best && i < bi)) { best = l[j]; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    tok = bi == 0x7fffffff ? 0 : bi;
  } else {
    // k-th largest key by bitwise binary search: T = max{t : #(key >= t) >= k}
    uint32_t key[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) key[j] = order_key(l[j]);
    const int k = min(P.top_k, V);
    uint32_t T = 0;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = T | (1u << bit);
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < NPL; ++j) cnt += key[j] >= cand;
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (cnt >= k) T = cand;
    }
    int gt = 0, eq = 0;
#pragma unroll
    for (int j = 0; j < NPL; ++j) { gt += key[j] > T; eq += key[j] == T; }
    const int need = k - __reduce_add_sync(0xffffffffu, gt);        // how many of the entries equal to T are kept (lowest indices)
    int eq_before = eq;                                              // exclusive prefix of eq over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, eq_before, o);
      if (lane >= o) eq_before += up;
    }
    eq_before -= eq;
    float sp[NPL];
    float mx = -INFINITY;
    int seen = 0;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      bool keep = key[j] > T;
      if (key[j] == T) { keep = eq_before + seen < need; ++seen; }
      sp[j] = keep ? l[j] / P.temperature : -INFINITY;
      if (keep) mx = fmaxf(mx, sp[j]);
    }
    mx = warp_max(mx);
#pragma unroll
    for (int j = 0; j < NPL; ++j) sp[j] = sp[j] == -INFINITY ? 0.f : expf(sp[j] - mx);
    double loc[8], run = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      double mine = 0.0;
#pragma unroll
      for (int e = 0; e < PER; ++e) mine += (double)sp[c * PER + e];
      loc[c] = run;
      run += mine;
    }
    double incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    const double excl = incl - run;
    const double total = __shfl_sync(0xffffffffu, incl, 31);
    const double target = (double)P.uniforms[(size_t)row * P.u_stride + st] * total;
    int pick = 0x7fffffff, last = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      double cum = excl + loc[c];
#pragma unroll
      for (int e = 0; e < PER; ++e) {
        const float pv = sp[c * PER + e];
        const int i = lane * NPL + c * PER + e;
        if (pv > 0.f) {
          last = max(last, i);
          cum += (double)pv;
          if (cum > target) pick = min(pick, i);
        }
      }
    }
    pick = __reduce_min_sync(0xffffffffu, pick);
    last = __reduce_max_sync(0xffffffffu, last);
    tok = pick == 0x7fffffff ? last : pick;
  }
  if (lane == 0) P.tokens[(size_t)row * P.tok_stride + st + 1] = tok;
  return tok;
}

// logits = bias + sum of the K slices (whole CTA, into shared memory) -> one warp per row samples -> token embedding -> x and
// layer 0's LayerNorm -> planes (whole CTA again).  Two rows per pass.
__device__ __forceinline__ void mk_row_sample(const MkPlan& P, const MkPhase& ph, float* lg /*[2][V] shared*/, Ctl& ctl, int st) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int V = ph.N, v4 = V >> 2, D = ph.D, n4 = D >> 2, c = threadIdx.x;
  const size_t mn = (size_t)ph.M * V;
  for (int r0 = (int)blockIdx.x; r0 < ph.M; r0 += 2 * (int)gridDim.x) {
    const int row[2] = {r0, r0 + (int)gridDim.x};
    const bool on[2] = {true, row[1] < ph.M};
    for (int f = c; f < 2 * v4; f += MK_NTHREADS) {
      const int k = f / v4, j = f - k * v4;
      if (!on[k]) continue;
      float4 a = sum_slices(ph.part + (size_t)row[k] * V + j * 4, mn, ph.in_splits);
      if (ph.bias) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(ph.bias) + j);
        a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
      }
      *reinterpret_cast<float4*>(lg + k * V + j * 4) = a;
      if (P.logits_out) *reinterpret_cast<float4*>(P.logits_out + (size_t)row[k] * P.lo_stride + (size_t)st * V + j * 4) = a;
    }
    __syncthreads();
    if (warp < 2 && on[warp]) {
      int tok;
      if (V == 512) tok = mk_sample_row<16>(P, lg + warp * V, row[warp], st, lane);
      else if (V == 1024) tok = mk_sample_row<32>(P, lg + warp * V, row[warp], st, lane);
      else tok = mk_sample_row<8>(P, lg + warp * V, row[warp], st, lane);
      if (lane == 0) ctl.tok[warp] = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
    }
    __syncthreads();
    // next step's input: token embedding -> x, layer 0's LayerNorm -> planes
    float4 v[2];
    v[0] = v[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < n4) {
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (on[k]) {
          v[k] = __ldg(reinterpret_cast<const float4*>(ph.emb + (size_t)ctl.tok[k] * D) + c);
          if (ph.pos) {                                   // absolute positional table: the token enters at position st + 1
            const float4 q = __ldg(reinterpret_cast<const float4*>(ph.pos + (size_t)(st + 1) * D) + c);
            v[k].x = fmaf(q.x, ph.pos_scale, v[k].x); v[k].y = fmaf(q.y, ph.pos_scale, v[k].y);
            v[k].z = fmaf(q.z, ph.pos_scale, v[k].z); v[k].w = fmaf(q.w, ph.pos_scale, v[k].w);
          }
          *reinterpret_cast<float4*>(ph.x + (size_t)row[k] * D + c * 4) = v[k];
        }
    }
    __nv_bfloat16* const out[2] = {ph.outp + (size_t)row[0] * P.planes * ph.out_kp, ph.outp + (size_t)row[1] * P.planes * ph.out_kp};
    ln_pair(v, on, c, n4, D, ph.gain, ph.beta, out, P.planes, ph.out_kp, ctl.red);
  }
}

// SMALL: the <= 8-row flavour (GEMV phases, no tile GEMM): a separate instantiation so that the GEMV's registers (9 weight vectors in
// flight per lane) do not spill the tile kernel's hot loops.
template <bool SMALL, bool WIDE>
__global__ void __launch_bounds__(WIDE ? MK_THREADS_WIDE : MK_THREADS, 1) decode_megakernel(const __grid_constant__ MkPlan P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) Ctl ctl;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < MK_STAGES; ++s) {
      mbar_init(smem_u32(&ctl.full_bar[s]), 1);
      mbar_init(smem_u32(&ctl.empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&ctl.accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl.tmem_slot)), "r"(MK_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = ctl.tmem_slot;

  Pipe pipe;
  unsigned int bar_target = 0;
  const int nph = P.nphases, steps = P.steps;
  if ((int)threadIdx.x < P.nmaps) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.maps[threadIdx.x]) : "memory");
  unsigned long long t_prev = 0;
  const bool tracing = P.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (tracing) t_prev = gtime_ns();
  for (int st = 0; st < steps; ++st) {
    for (int i = 0; i < nph; ++i) {
      const MkPhase& ph = P.phases[i];                 // kernel-parameter space: uniform constant-bank reads
      const int nxt = i + 1 < nph ? i + 1 : 0;
      if (threadIdx.x == 32 && P.phases[nxt].type == MK_GEMM && !P.phases[nxt].gemv) {     // descriptors of the next GEMM phase
        asm volatile("prefetch.tensormap [%0];" ::"l"(&P.maps[P.phases[nxt].mapA]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&P.maps[P.phases[nxt].mapW]) : "memory");
      }
      if (ph.pf_f1 > ph.pf_f0 && P.pf_budget > 0) mk_kv_prefetch(P, ph, i, st);
      switch (ph.type) {
        case MK_GEMM:
          if (!SMALL) mk_gemm(P, ph, smem, smem_base, ctl, pipe, tmem_base);
          else if (P.planes == 1) mk_gemv<true>(P, ph, smem);
          else mk_gemv<false>(P, ph, smem);
          break;
        case MK_ATTN: mk_attn<WIDE>(P, ph, smem, ctl, st); break;
        case MK_ROW_RESLN: mk_row_resln(P, ph, ctl); break;
        case MK_ROW_GELU: mk_row_gelu(P, ph); break;
        case MK_ROW_SAMPLE: mk_row_sample(P, ph, reinterpret_cast<float*>(smem), ctl, st); break;
        default: break;
      }
      if (!SMALL && P.attn_pre && ph.type == MK_GEMM && P.phases[nxt].type == MK_ATTN && nxt > i) {
        mk_attn_prefetch(P, P.phases[nxt], st);
      }
      grid_sync(P.bar, bar_target, P.phases[nxt].type == MK_GEMM && !P.phases[nxt].gemv);
      if (tracing) {
        const unsigned long long t = gtime_ns();
        P.trace[i] += t - t_prev;
        t_prev = t;
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(MK_TMEM_COLS) : "memory");
  }
}

}  // namespace

bool mk_supported(int D, int inner, int F, int V, int H, int planes, int max_keys) {
  static const bool off = getenv("DIM_DECODE_IMPL") != nullptr && std::string(getenv("DIM_DECODE_IMPL")) == "graph";
  if (off) return false;
  if (planes < 1 || planes > 3) return false;
  if (D % 64 || inner % 64 || F % 64 || D > 4 * MK_THREADS || inner != H * 64) return false;
  if (!(V == 256 || V == 512 || V == 1024)) return false;
  // per attention sub-group: 2-stage K/V ring + scores + partial outputs + q + reduction slots
  return mk_attn_scratch(max_keys) <= MK_RING / 6;
}

size_t mk_attn_scratch(int max_keys) { return 2 * 64 * 144 + (size_t)((max_keys + 3) / 4 * 4) * 4 + 8 * 64 * 4 + 3 * 64 * 4 + 16; }

int mk_attn_subgroups(int kv_bf16, int attn_mma, int small, int max_keys, int items) {
  static const bool wide_off = getenv("DIM_MK_WIDE") != nullptr && atoi(getenv("DIM_MK_WIDE")) == 0;      // A/B hook
  if (wide_off || !kv_bf16 || !attn_mma || small) return 6;
  // fewer (row, head) items than 6 sub-groups per SM can hold at once: the extra warps only slow the GEMM / row phases down
  // (32 LM-Listener chunks: 449.6 ms with 384 threads, 464.6 with 512).  Numerics do not depend on the choice.
  if (items <= 6 * 148) return 6;
  // the mma items keep [2][64] partial outputs, not [8][64]
  const size_t need = 2 * 64 * 144 + (size_t)((max_keys + 3) / 4 * 4) * 4 + 2 * 64 * 4 + 3 * 64 * 4 + 16;
  return need <= MK_RING / 8 ? 8 : 6;
}

int launch_decode_megakernel(const MkPlan& plan, cudaStream_t s) {
  static_assert(sizeof(MkPlan) <= 32000, "the plan travels as a kernel parameter (limit 32764 bytes)");
  DIM_REQUIRE(plan.nphases > 0 && plan.nphases <= MK_MAX_PHASES && plan.steps > 0 && plan.bar != nullptr, "decode megakernel: bad plan");
  constexpr size_t smem = MK_RING + 1024;
  int dev = 0, sms = 0;
  DIM_CHECK_CUDA(cudaGetDevice(&dev));
  DIM_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  bool small = false;
  for (int i = 0; i < plan.nphases; ++i) small = small || (plan.phases[i].type == MK_GEMM && plan.phases[i].gemv);
  DIM_REQUIRE(plan.attn_nsub == 6 || (plan.attn_nsub == 8 && !small && plan.kv_bf16 && plan.attn_mma), "decode megakernel: bad sub-group count");
  const bool wide = plan.attn_nsub == 8;
  const int threads = wide ? MK_THREADS_WIDE : MK_THREADS;
  auto kern = small ? decode_megakernel<true, false> : (wide ? decode_megakernel<false, true> : decode_megakernel<false, false>);
  DIM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  DIM_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  DIM_REQUIRE(per_sm >= 1, "decode megakernel: one CTA does not fit an SM");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;         // all CTAs co-resident: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope ps(CAT_DECODE_MK, s, 0, 0);
  DIM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, plan));
  DIM_LAUNCHED();
  return DIM_OK;
}

}  // namespace dimb
