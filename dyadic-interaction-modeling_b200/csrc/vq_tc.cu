// Codebook nearest neighbour on the tensor cores, bit-identical to the exact fp32 kernel (vq.cu: vq_argmin_f32).
//   reference: d_k = (|z|^2 + |e_k|^2) - 2 (z . e_k) in fp32, argmin with the first minimum winning (models/lib/quantizer.py:38-45).
//
// The exact kernel is bound by fp32 FFMA issue (131 kFLOP per token: 4.1 ms per 2^20 tokens = 2 % of the HBM rate of its 520
// algorithmic bytes per token).  Here the 512 dot products of a token run ONCE in fp16 on tcgen05 (128 tokens x 512 codes per tile,
// fp32 accumulators = all 512 TMEM columns) and only produce a SHORTLIST; the winner is decided by the exact fp32 expression.
//
// Shortlist bound.  zh = fp16(z), eh_k = fp16(e_k), dz = z - zh, de_k = e_k - eh_k (both differences are exact in fp32).
//   s_k = |e_k|^2 - 2 * dot_tc(zh, eh_k)                                  (approximate score, from the tensor core)
//   z.e_k - zh.eh_k = dz.eh_k + z.de_k    =>   |z.e_k - zh.eh_k| <= |dz| |eh|max + |z| |de|max          (Cauchy-Schwarz)
// |dz| is MEASURED per token by the converter warps, |eh|max and |de|max per codebook by vq_tc_prepare, so the bound holds for any
// data -- values outside the fp16 range make |dz| infinite, every code is shortlisted and the result is still exact, only slow.
//   eps = 2 (|dz| |eh|max + |z| |de|max)            operand rounding
//       + 2 * 2^-13 |zh| |eh|max                    the tensor core's fp32 accumulation of 128 exact fp16 products (8x margin)
//       + 2^-16 |z| |e|max + 2^-19 (|z|^2 + |e|^2max)   rounding inside the exact kernel's own fp32 expression (generous)
//   => |s_k - (d_k - |z|^2)| <= eps for the d_k the exact kernel computes, hence its argmin k* satisfies s_k* <= min_k s_k + 2 eps.
// Every code within 2 eps of the approximate minimum is re-evaluated with the SAME arithmetic as vq_argmin_f32 (|z|^2 by the same
// warp reduction, |e_k|^2 by the same warp reduction, the dot product as one sequential fmaf chain, (zz + ee) - 2 dot with each
// step rounded); the winner is the lexicographic minimum of (d, k) exactly like the exact kernel.  A lone candidate needs no
// distance.  On the benchmark distribution ~5 % of the tokens shortlist more than one code (bf16 operands: 50 %).
//
// Persistent CTAs (one per SM), 15 warps, one accumulator tile = 128 tokens x 512 codes = all of TMEM:
//   warps 0-7   epilogue: thread = token row (TMEM lane) x one 128-column quarter of each code half.  TMEM reads run at ~64 B/clk
//               per SM (measured), so the accumulator is read ONCE, 32 columns at a time through two register buffers (the next
//               load is in flight while a chunk is processed): running minimum, and every score within tau of the running minimum
//               (a superset of the final shortlist) is parked in a small per-thread list; a code half goes back to the MMA warp as
//               soon as it sits in registers; the parked scores are filtered against the final limit and the survivors go on a
//               (row, code) list;
//   warps 8-11  converters: fp32 z rows -> fp16, 128-byte-swizzled K-major smem tiles (double buffered), |z|^2 and |dz|^2;
//   warps 12-13 re-rank (even / odd tiles): rows with more than one listed code get exact distances (one pair per lane), then
//               idx is written for the whole tile;
//   warp  14    MMA issue (one lane): 16 x tcgen05.mma M=128 N=256 K=16 per tile, code half 0 first.
// The fp16 codebook (128 KB, swizzled) stays in shared memory for the whole kernel.  Measured (B200, 2^20 tokens): 389 us against
// 4113 us for the exact FFMA kernel (profiles/r02_vq_roofline.jsonl); timing ablations behind dim_debug_vq_argmin_impl(dbg << 8).
#include <cuda_fp16.h>

#include "tc_ptx.cuh"
#include "vq.cuh"

namespace dimb {

namespace {

constexpr int VT_D = 128, VT_K = 512, VT_TM = 128;                 // only this shape (the DIM codebook): 512 codes x 128 dims
constexpr uint32_t VT_E_BYTES = VT_K * VT_D * 2;                   // 128 KB: 2 k-blocks x [512 rows x 128 B]
constexpr uint32_t VT_Z_BYTES = VT_TM * VT_D * 2;                  // 32 KB per buffer: 2 k-blocks x [128 rows x 128 B]
constexpr uint32_t VT_SMEM = VT_E_BYTES + 2 * VT_Z_BYTES + 1024;   // + alignment slack
constexpr int VT_CAP = 1024;                                       // (row, code) pairs re-ranked cooperatively per tile
constexpr int VT_THREADS = 15 * 32, VT_PARK = 6;                 // parked scores per epilogue thread

// byte offset of element (row, col) inside a K-major SWIZZLE_128B operand image of `rows` rows: k-block (64 columns) major,
// 128-byte rows, the 16-byte chunk index XORed with (row & 7)  -- what TMA writes for a [rows x 64] box at a 1024-aligned base
__device__ __forceinline__ uint32_t swz_off(int rows, int row, int col) {
  const int kb = col >> 6, c = col & 63;
  return (uint32_t)(kb * rows * 128 + row * 128 + ((((c >> 3) ^ (row & 7)) << 4) | ((c & 7) << 1)));
}

// codebook constants: [0] max |fp16(e_k)|^2, [1] max |e_k - fp16(e_k)|^2, [2] max |e_k|^2 (non-negative floats order like ints)
// |e_k|^2 with the reduction order of vq_argmin_f32 (lane l: elements l, l+32, ...; xor tree); also the fp16 swizzled codebook image
__global__ void __launch_bounds__(256) vq_tc_prepare(const float* __restrict__ E, float* __restrict__ e2, uint8_t* __restrict__ Eimg,
                                                      float* __restrict__ cst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = blockIdx.x * 8 + warp; k < VT_K; k += gridDim.x * 8) {
    float s = 0.f, sh = 0.f, sd = 0.f;
    for (int c = lane; c < VT_D; c += 32) {
      const float v = E[(size_t)k * VT_D + c];
      s = fmaf(v, v, s);
      const __half h = __float2half_rn(v);
      const float hf = __half2float(h), r = v - hf;
      sh = fmaf(hf, hf, sh);
      sd = fmaf(r, r, sd);
      *reinterpret_cast<__half*>(Eimg + swz_off(VT_K, k, c)) = h;
    }
    s = warp_sum(s);
    sh = warp_sum(sh);
    sd = warp_sum(sd);
    if (lane == 0) {
      e2[k] = s;
      atomicMax(reinterpret_cast<int*>(cst + 0), __float_as_int(sh));
      atomicMax(reinterpret_cast<int*>(cst + 1), __float_as_int(sd));
      atomicMax(reinterpret_cast<int*>(cst + 2), __float_as_int(s));
    }
  }
}

// mbarrier wait with back-off: the polls of 18 waiting warps are shared-memory requests that compete with the tensor core's operand
// reads (SS-mode MMAs need 96 of the 128 B/clk) -- measured: tight polling made the MMAs of a tile 3-4x slower
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, unsigned ns) {
  for (;;) {
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
    __nanosleep(ns);
  }
}

__device__ __forceinline__ void bar_epi() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 256 epilogue threads only

__device__ __forceinline__ unsigned long long pack_dk(float d, int k) {      // orders like (d, k) lexicographically
  const uint32_t b = __float_as_uint(d);
  const uint32_t key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)key << 32) | (uint32_t)k;
}

// vq_argmin_f32's expression, operation for operation -- including its |z|^2: lane l of a warp chains fmaf over elements l, l+32,
// l+64, l+96, then an xor-butterfly over the 32 lanes (every lane of a pair adds the same two numbers, so one thread can replay it)
__device__ __forceinline__ float exact_dist(const float* __restrict__ z, const float* __restrict__ E, int t, int code, float ee) {
  const float4* zr = reinterpret_cast<const float4*>(z + (size_t)t * VT_D);
  const float4* er = reinterpret_cast<const float4*>(E + (size_t)code * VT_D);
  float s[32];
#pragma unroll
  for (int l = 0; l < 32; ++l) s[l] = 0.f;
#pragma unroll
  for (int d4 = 0; d4 < VT_D / 4; ++d4) {
    const float4 a = zr[d4];
    const int l = (4 * d4) & 31;
    s[l] = fmaf(a.x, a.x, s[l]); s[l + 1] = fmaf(a.y, a.y, s[l + 1]); s[l + 2] = fmaf(a.z, a.z, s[l + 2]); s[l + 3] = fmaf(a.w, a.w, s[l + 3]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int l = 0; l < o; ++l) s[l] = s[l] + s[l + o];
  const float zz = s[0];
  float dot = 0.f;
#pragma unroll 8
  for (int d4 = 0; d4 < VT_D / 4; ++d4) {
    const float4 a = zr[d4], e = __ldg(er + d4);
    dot = fmaf(a.x, e.x, dot); dot = fmaf(a.y, e.y, dot); dot = fmaf(a.z, e.z, dot); dot = fmaf(a.w, e.w, dot);
  }
  return __fsub_rn(__fadd_rn(zz, ee), __fmul_rn(2.f, dot));
}

__device__ __noinline__ float exact_dist_cold(const float* z, const float* E, int t, int code, float ee) {
  return exact_dist(z, E, t, code, ee);
}

// 64 accumulator columns of this thread's TMEM lane: two 32-column loads in flight, one wait (the wait names the registers so that
// nothing consumes them early)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float* v) {
  uint32_t r[64];
#pragma unroll
  for (int h = 0; h < 2; ++h)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[32 * h + 0]), "=r"(r[32 * h + 1]), "=r"(r[32 * h + 2]), "=r"(r[32 * h + 3]), "=r"(r[32 * h + 4]), "=r"(r[32 * h + 5]),
          "=r"(r[32 * h + 6]), "=r"(r[32 * h + 7]), "=r"(r[32 * h + 8]), "=r"(r[32 * h + 9]), "=r"(r[32 * h + 10]), "=r"(r[32 * h + 11]),
          "=r"(r[32 * h + 12]), "=r"(r[32 * h + 13]), "=r"(r[32 * h + 14]), "=r"(r[32 * h + 15]), "=r"(r[32 * h + 16]), "=r"(r[32 * h + 17]),
          "=r"(r[32 * h + 18]), "=r"(r[32 * h + 19]), "=r"(r[32 * h + 20]), "=r"(r[32 * h + 21]), "=r"(r[32 * h + 22]), "=r"(r[32 * h + 23]),
          "=r"(r[32 * h + 24]), "=r"(r[32 * h + 25]), "=r"(r[32 * h + 26]), "=r"(r[32 * h + 27]), "=r"(r[32 * h + 28]), "=r"(r[32 * h + 29]),
          "=r"(r[32 * h + 30]), "=r"(r[32 * h + 31])
        : "r"(taddr + 32u * h)
        : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    asm volatile("" : "+r"(r[j]));             // consumers are ordered after the wait
    v[j] = __uint_as_float(r[j]);
  }
}

// clock64 timeline of CTA 0 (debug): trace[role][it][event], 4 roles x 16 tiles x 12 events
#define VT_TRACE(role, ev) do { if (trace && blockIdx.x == 0 && lane == 0 && it < 16) trace[((role) * 16 + it) * 12 + (ev)] = clock64(); } while (0)

// 32 accumulator columns of this thread's TMEM lane; the wait names the registers so that nothing consumes them early
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) asm volatile("" : "+r"(r[j]));
}
__device__ __forceinline__ float sel32(const uint32_t (&v)[32], int j) {
  uint32_t a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 16) ? v[i + 16] : v[i];
#pragma unroll
  for (int w = 8; w > 0; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; ++i) a[i] = (j & w) ? a[i + w] : a[i];
  return __uint_as_float(a[0]);
}

// v[j] for a run-time j without indexing the register array: a 6-level select tree (63 selects)
__device__ __forceinline__ float sel64(const float (&v)[64], int j) {
  float a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = (j & 32) ? v[i + 32] : v[i];
#pragma unroll
  for (int w = 16; w > 0; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; ++i) a[i] = (j & w) ? a[i + w] : a[i];
  return a[0];
}

// Warp roles (15 warps, 128 registers each): 0-7 epilogue, 8-11 converters, 12-13 re-rank (even / odd tiles), 14 MMA issue.
__global__ void __launch_bounds__(VT_THREADS, 1) vq_argmin_tc(const float* __restrict__ z, const float* __restrict__ E,
                                                              const float* __restrict__ e2g, const uint8_t* __restrict__ Eimg,
                                                              const float* __restrict__ cst, int64_t* __restrict__ idx, int N,
                                                              int* __restrict__ stats, int dbg, long long* __restrict__ trace) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t zfull[2], zfree[2], acc_full[2], acc_free[2], list_full[2], list_free[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float e2s[VT_K];
  __shared__ float z2s[8][VT_TM], dz2s[8][VT_TM];   // |z|^2, |z - fp16(z)|^2 of the tiles in flight (a slot is rewritten 8 tiles later)
  __shared__ float minpart[2][2][VT_TM];
  __shared__ __align__(8) unsigned long long rowbest[2][VT_TM];
  __shared__ int rowcnt[2][VT_TM], rowfirst[2][VT_TM];
  __shared__ uint16_t pairs[2][VT_CAP];
  __shared__ float park_s[VT_PARK][256];
  __shared__ uint16_t park_k[VT_PARK][256];
  __shared__ int npairs[2];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t e_u = base, z_u = base + VT_E_BYTES;
  uint8_t* zbuf = smem + VT_E_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (N + VT_TM - 1) / VT_TM;

  // ---- set-up: codebook image and |e|^2 into shared memory, barriers, TMEM (all 512 columns)
  for (uint32_t i = tid; i < VT_E_BYTES / 16; i += VT_THREADS) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(Eimg)[i];
  for (int i = tid; i < VT_K; i += VT_THREADS) e2s[i] = e2g[i];
  if (tid < 2 * VT_TM) { (&rowbest[0][0])[tid] = ~0ull; (&rowcnt[0][0])[tid] = 0; (&rowfirst[0][0])[tid] = 0x7fffffff; }
  if (tid == 0) {
    npairs[0] = npairs[1] = 0;
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&zfull[i]), 4);           // one lane of each of the 4 converter warps arrives
      mbar_init(smem_u32(&zfree[i]), 1);           // tcgen05.commit arrives
      mbar_init(smem_u32(&acc_full[i]), 1);        // tcgen05.commit arrives (per code half)
      mbar_init(smem_u32(&acc_free[i]), 8);        // one lane of each of the 8 epilogue warps arrives (per code half)
      mbar_init(smem_u32(&list_full[i]), 8);       // one lane of each epilogue warp arrives (per list buffer)
      mbar_init(smem_u32(&list_free[i]), 1);       // one lane of the re-rank warp arrives
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the codebook image was written with generic stores, the MMA reads it
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (warp == 14) {
    // ===== MMA issue: one lane =====
    if (lane == 0) {
      // D[128 tokens x 256 codes] per instruction: M = 128, N = 256, fp16 x fp16 -> fp32, both operands K-major
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait_sleep(smem_u32(&zfull[b]), (it >> 1) & 1, 32);
        VT_TRACE(0, 0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait_sleep(smem_u32(&acc_free[h]), (it & 1) ^ 1, 20);   // the epilogue has drained this half of the previous tile
          VT_TRACE(0, 1 + h);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t adesc = make_sdesc(z_u + b * VT_Z_BYTES + kb * (VT_TM * 128));
            const uint64_t bdesc = make_sdesc(e_u + kb * (VT_K * 128) + h * (256 * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem + (uint32_t)(h * 256), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&acc_full[h]));
        }
        umma_commit(smem_u32(&zfree[b]));
      }
    }
    __syncwarp();
  } else if (warp >= 12) {
    // ===== re-rank: the shortlist of tile `it` (list buffer it & 1) -> exact distances -> idx =====
    const int rt = lane;
    int it = 0, n_multi = 0, n_cand = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int p = it & 1, t0 = tile * VT_TM;
      if (p != warp - 12) continue;                                   // one warp per list buffer: a re-rank may take two tile periods
      mbar_wait_sleep(smem_u32(&list_full[p]), (it >> 1) & 1, 100);
      if (warp == 12) VT_TRACE(1, 0);
      const int np = min(npairs[p], VT_CAP);
      // most rows shortlist one code and need no distance: compact the pairs of the other rows first (in place: the write index
      // never passes the read index), so that the ~10 real pairs of a tile are evaluated side by side in one round
      int nreal = 0;
      for (int i0 = 0; i0 < np; i0 += 32) {
        const int i = i0 + rt;
        const uint16_t pk = i < np ? pairs[p][i] : (uint16_t)0;
        const bool real = i < np && rowcnt[p][pk >> 9] > 1;
        const unsigned bal = __ballot_sync(0xffffffffu, real);
        __syncwarp();
        if (real) pairs[p][nreal + __popc(bal & ((1u << rt) - 1u))] = pk;
        nreal += __popc(bal);
        __syncwarp();
      }
      for (int i = rt; i < nreal && !(dbg & 2); i += 32) {
        const int pr = pairs[p][i] >> 9, k = pairs[p][i] & 511;
        atomicMin(&rowbest[p][pr], pack_dk(exact_dist(z, E, t0 + pr, k, e2s[k]), k));
      }
      if (warp == 12) VT_TRACE(1, 1);
      __syncwarp();
      if (warp == 12) VT_TRACE(1, 2);
#pragma unroll
      for (int r = rt; r < VT_TM; r += 32) {
        const int cnt = rowcnt[p][r], t = t0 + r;
        if (t < N) {
          idx[t] = cnt > 1 ? (int)(uint32_t)(rowbest[p][r] & 0xffffffffull) : rowfirst[p][r];
          n_multi += cnt > 1;
          n_cand += cnt;
        }
        rowbest[p][r] = ~0ull; rowcnt[p][r] = 0; rowfirst[p][r] = 0x7fffffff;
      }
      if (rt == 0) npairs[p] = 0;
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&list_free[p])) : "memory");
    }
    if (stats) {
      n_multi = __reduce_add_sync(0xffffffffu, n_multi);
      n_cand = __reduce_add_sync(0xffffffffu, n_cand);
      if (lane == 0) { atomicAdd(stats, n_multi); atomicAdd(stats + 1, n_cand); }
    }
  } else if (warp >= 8) {
    // ===== converters: tile t -> buffer t & 1.  A warp owns 32 rows of a tile; one 128-bit load instruction covers one 128-byte
    // segment of 4 rows (8 lanes per row), so a lane ends up with 16 elements of ONE row: its partial |z|^2 and |z - fp16(z)|^2
    // need only a 3-step reduction over the 8 lanes of the row.  Four batches of 8 rows through two register sets keep loads in
    // flight while the other set is converted.  (|z|^2 here only feeds the shortlist bound; the re-rank replays the exact order.) =====
    const int cw = warp - 8, part = lane & 7, rsub = lane >> 3;
    float4 va[2][4], vb[2][4];
    auto load8 = [&](float4 (&v)[2][4], int tile, int g0) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int t = tile * VT_TM + cw * 32 + (g0 + g) * 4 + rsub;
#pragma unroll
        for (int sg = 0; sg < 4; ++sg)
          v[g][sg] = (tile < ntiles && t < N && !(dbg & 4)) ? __ldcs(reinterpret_cast<const float4*>(z + (size_t)t * VT_D + sg * 32 + part * 4))
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto conv8 = [&](float4 (&v)[2][4], uint8_t* zb, int slot, int g0) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int r = cw * 32 + (g0 + g) * 4 + rsub;
        float sq = 0.f, rq = 0.f;
#pragma unroll
        for (int sg = 0; sg < 4; ++sg) {
          const float4 x = v[g][sg];
          const __half2 h01 = __floats2half2_rn(x.x, x.y), h23 = __floats2half2_rn(x.z, x.w);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const float r0 = x.x - f01.x, r1 = x.y - f01.y, r2 = x.z - f23.x, r3 = x.w - f23.y;
          sq = fmaf(x.x, x.x, sq); sq = fmaf(x.y, x.y, sq); sq = fmaf(x.z, x.z, sq); sq = fmaf(x.w, x.w, sq);
          rq = fmaf(r0, r0, rq); rq = fmaf(r1, r1, rq); rq = fmaf(r2, r2, rq); rq = fmaf(r3, r3, rq);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&h01);
          pk.y = *reinterpret_cast<const uint32_t*>(&h23);
          *reinterpret_cast<uint2*>(zb + swz_off(VT_TM, r, sg * 32 + part * 4)) = pk;
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          sq += __shfl_xor_sync(0xffffffffu, sq, o);
          rq += __shfl_xor_sync(0xffffffffu, rq, o);
        }
        if (part == 0) { z2s[slot][r] = sq; dz2s[slot][r] = rq; }
      }
    };
    int it = 0;
    load8(va, blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      uint8_t* zb = zbuf + b * VT_Z_BYTES;
      const bool ld = !(dbg & 16);
      if (ld || it == 0) load8(vb, tile, 2);
      mbar_wait_sleep(smem_u32(&zfree[b]), ((it >> 1) & 1) ^ 1, 200);   // the MMAs that read this buffer two tiles ago are done
      if (warp == 8) VT_TRACE(2, 0);
      conv8(va, zb, it & 7, 0);
      if (ld) load8(va, tile, 4);
      conv8(vb, zb, it & 7, 2);
      if (warp == 8) VT_TRACE(2, 1);
      if (ld) load8(vb, tile, 6);
      conv8(va, zb, it & 7, 4);
      if (ld) load8(va, tile + gridDim.x, 0);
      conv8(vb, zb, it & 7, 6);
      if (warp == 8) VT_TRACE(2, 2);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&zfull[b])) : "memory");
    }
  } else {
    // ===== epilogue: thread = token row (TMEM lane) x one 128-column quarter of each code half.  TMEM reads run at ~64 B/clk per
    // SM, so the accumulator is read ONCE: a thread keeps a running minimum and parks every score that is within tau of the
    // running minimum (a superset of the final shortlist: the final minimum is not larger) in a small per-thread list; each code
    // half of the accumulator goes back to the MMA warp as soon as it has been read; the parked scores are filtered against the
    // final limit and the survivors go on the (row, code) list of the re-rank warps =====
    const int q = warp & 3, cq = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    const float ehmax = sqrtf(cst[0]), demax = sqrtf(cst[1]), e2max = cst[2], emax = sqrtf(cst[2]);
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int p = it & 1, t0 = tile * VT_TM;
      const bool valid = t0 + row < N;                                // rows past the end of z shortlist nothing
      float m_run = INFINITY, tau = 0.f;
      int nt = 0;                                                     // parked scores of this thread
      bool cold = false;                                              // some candidate was evaluated on the spot (park list full)
      mbar_wait_sleep(smem_u32(&list_free[p]), ((it >> 1) & 1) ^ 1, 64);   // the re-rank warp is done with this list buffer (tile it - 2)
      // 8 chunks of 32 columns per thread (chunks 0-3: code half 0, 4-7: code half 1), two register buffers: the load of chunk
      // c + 1 is in flight while chunk c is processed
      auto col_of = [&](int c) { return (c >> 2) * 256 + cq * 128 + (c & 3) * 32; };
      auto process = [&](uint32_t (&r)[32], int c) {
        const int c0 = col_of(c);
        float m4[4] = {INFINITY, INFINITY, INFINITY, INFINITY};       // four independent chains
#pragma unroll
        for (int j = 0; j < 32; ++j) {                                // scores in place
          const float sc = fmaf(-2.f, __uint_as_float(r[j]), e2s[c0 + j]);
          r[j] = __float_as_uint(sc);
          m4[j & 3] = fminf(m4[j & 3], sc);
        }
        const float cm = fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3]));
        m_run = fminf(m_run, cm);
        const float lr = m_run + tau;
        // sign bit of (lr - score) = "not within tau"; a funnel shift per element collects the sign bits, four independent chains
        // of 8 (element 8q + i ends up in bit 7 - i of w[q]); NaN scores/limits have a clear sign bit here: parked (slow, exact)
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) w[qq] = __funnelshift_l(__float_as_uint(lr - __uint_as_float(r[8 * qq + i])), w[qq], 1);
        uint32_t mask = __brev(~((w[0] << 24) | ((w[1] & 0xffu) << 16) | ((w[2] & 0xffu) << 8) | (w[3] & 0xffu)));   // bit j = element j
        if (!valid || (dbg & 1)) mask = 0;
        // every candidate is parked with the chunk minimum cm as its score: a lower bound, so the final filter keeps a superset
        // (a candidate survives whenever the chunk minimum does) and no register has to be picked by a run-time index
        while (mask && !(dbg & 64)) {
          const int j = __ffs(mask) - 1;
          mask &= mask - 1;
          if (nt == VT_PARK) {                                        // full: drop what the current limit already rules out
            int keep = 0;
            for (int e = 0; e < VT_PARK; ++e) {
              const float ps = park_s[e][tid];
              const uint16_t pk = park_k[e][tid];
              if (!(ps > lr)) { park_s[keep][tid] = ps; park_k[keep][tid] = pk; ++keep; }
            }
            nt = keep;
          }
          if (nt < VT_PARK) { park_s[nt][tid] = cm; park_k[nt][tid] = (uint16_t)(c0 + j); ++nt; }
          else {                                                      // six live near-ties already (degenerate data): evaluate now
            atomicMin(&rowbest[p][row], pack_dk(exact_dist_cold(z, E, t0 + row, c0 + j, e2s[c0 + j]), c0 + j));
            cold = true;
          }
        }
      };
      auto release = [&](int h) {                                     // this warp has read code half h of the accumulator
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_free[h])) : "memory");
      };
      uint32_t ra[32], rb[32];
      mbar_wait_sleep(smem_u32(&acc_full[0]), it & 1, 32);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (warp == 0) VT_TRACE(3, 0);
      tmem_ld32_issue(tl + (uint32_t)col_of(0), ra);
      {
        const float zz = z2s[it & 7][row], dz2 = dz2s[it & 7][row];   // written by the converters before this tile's MMAs were issued
        const float zn = sqrtf(zz), dzn = sqrtf(dz2);
        const float eps = 2.f * (dzn * ehmax + zn * demax) + 2.44140625e-4f * (zn + dzn) * ehmax + 1.52587891e-5f * zn * emax +
                          1.90734863e-6f * (zz + e2max);
        tau = 2.01f * eps;          // 2 eps (header) + margin: rounding of this expression, summation order of |z|^2 here
      }
#pragma unroll 1
      for (int c = 0; c < 8; c += 2) {
        tmem_ld_wait(ra);
        tmem_ld32_issue(tl + (uint32_t)col_of(c + 1), rb);
        if (!(dbg & 8)) process(ra, c);
        tmem_ld_wait(rb);
        if (c == 2) {
          release(0);
          if (warp == 0) VT_TRACE(3, 1);
          mbar_wait_sleep(smem_u32(&acc_full[1]), it & 1, 32);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (warp == 0) VT_TRACE(3, 2);
        }
        if (c < 6) tmem_ld32_issue(tl + (uint32_t)col_of(c + 2), ra);
        else release(1);
        if (!(dbg & 8)) process(rb, c + 1);
      }
      if (warp == 0) VT_TRACE(3, 3);
      if (dbg & 32) nt = 0;
      minpart[p][cq][row] = m_run;
      bar_epi();
      if (warp == 0) VT_TRACE(3, 4);
      const float lim = fminf(minpart[p][0][row], minpart[p][1][row]) + tau;
      if (warp == 0) VT_TRACE(3, 5);
      // final filter of the parked scores; the survivors go on the (row, code) list (the warp reserves its slots with ONE atomic);
      // only when the list is full (degenerate data) a survivor is evaluated on the spot
      int cnt = 0, first = 0x7fffffff;
      unsigned live = 0;
      for (int e = 0; e < nt; ++e)
        if (!(park_s[e][tid] > lim)) {
          if (cnt == 0) first = park_k[e][tid];                       // parked in ascending code order
          live |= 1u << e;
          ++cnt;
        }
      {
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += u;
        }
        int basep = 0;
        if (lane == 31 && incl > 0) basep = atomicAdd(&npairs[p], incl);
        basep = __shfl_sync(0xffffffffu, basep, 31);
        int slot = basep + incl - cnt;
        if (cnt > 0 || cold) {
          atomicAdd(&rowcnt[p][row], cnt + (cold ? 2 : 0));           // cold: rowbest already holds a distance, it must decide
          if (cnt > 0) atomicMin(&rowfirst[p][row], first);
          while (live) {
            const int e = __ffs(live) - 1;
            live &= live - 1;
            const int k = park_k[e][tid];
            if (slot < VT_CAP) pairs[p][slot] = (uint16_t)((row << 9) | k);
            else atomicMin(&rowbest[p][row], pack_dk(exact_dist_cold(z, E, t0 + row, k, e2s[k]), k));
            ++slot;
          }
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&list_full[p])) : "memory");
      if (warp == 0) VT_TRACE(3, 8);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

struct TcCodebook {                      // derived images of one codebook pointer (owned, cached per device pointer)
  const float* E = nullptr;
  int device = -1;
  float* e2 = nullptr;
  uint8_t* img = nullptr;
  float* cst = nullptr;
};
TcCodebook g_cb[8];
int g_cb_n = 0;

}  // namespace

int g_vq_tc_dbg = 0;                    // timing ablations (dim_debug_vq_argmin_impl bits 8..): results are wrong when set
int g_vq_argmin_impl = 0;               // 0: tensor-core shortlist + exact re-rank when the shape allows; 1: exact FFMA kernel only

bool vq_argmin_tc_supported(int N, int D, int K) { return g_vq_argmin_impl == 0 && D == VT_D && K == VT_K && N >= 4 * VT_TM; }

// The codebook images are rebuilt on every call (4 us): the weights are borrowed pointers and may have been overwritten in place.
int launch_vq_argmin_tc(const float* z, const float* E, int64_t* idx, int N, int* stats, cudaStream_t s, long long* trace) {
  int dev = 0, sms = 0;
  DIM_CHECK_CUDA(cudaGetDevice(&dev));
  DIM_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TcCodebook* cb = nullptr;
  for (int i = 0; i < g_cb_n; ++i)
    if (g_cb[i].E == E && g_cb[i].device == dev) cb = &g_cb[i];
  if (!cb) {
    static int next = 0;
    if (g_cb_n < 8) cb = &g_cb[g_cb_n++];
    else cb = &g_cb[next++ & 7];                                   // tiny cache: recycle
    if (cb->e2 && cb->device != dev) {                             // buffers of another device: leave them, allocate here
      cb->e2 = nullptr; cb->img = nullptr; cb->cst = nullptr;
    }
    cb->E = E;
    cb->device = dev;
    if (!cb->e2) {
      DIM_CHECK_CUDA(cudaMalloc(&cb->e2, VT_K * sizeof(float)));
      DIM_CHECK_CUDA(cudaMalloc(&cb->img, VT_E_BYTES));
      DIM_CHECK_CUDA(cudaMalloc(&cb->cst, 4 * sizeof(float)));
    }
  }
  DIM_CHECK_CUDA(cudaMemsetAsync(cb->cst, 0, 4 * sizeof(float), s));
  vq_tc_prepare<<<16, 256, 0, s>>>(E, cb->e2, cb->img, cb->cst);
  DIM_LAUNCHED();
  DIM_CHECK_CUDA(cudaFuncSetAttribute(vq_argmin_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VT_SMEM));   // per device: cheap, unconditional
  const int ntiles = (N + VT_TM - 1) / VT_TM;
  ProfScope ps(CAT_VQ_ARGMIN, s, (double)N * (VT_D * 4.0 + 8.0), 2.0 * N * (double)VT_D * VT_K);
  vq_argmin_tc<<<std::min(sms, ntiles), VT_THREADS, VT_SMEM, s>>>(z, E, cb->e2, cb->img, cb->cst, idx, N, stats, g_vq_tc_dbg, trace);
  DIM_LAUNCHED();
  return DIM_OK;
}

}  // namespace dimb

// tuning / test hooks (not part of the stable ABI): force the exact FFMA kernel (1) or allow the tensor-core path (0);
// run the tensor-core kernel and return {tokens with more than one shortlisted code, shortlisted codes in total}
extern "C" int dim_debug_vq_argmin_impl(int impl) {
  dimb::g_vq_argmin_impl = impl & 0xff;
  dimb::g_vq_tc_dbg = impl >> 8;
  return DIM_OK;
}
extern "C" int dim_debug_vq_argmin_tc_trace(const float* z, const float* E, int64_t* idx, int N, long long* trace_dev, void* stream) {
  if (int e = dimb::ensure_device()) return e;
  return dimb::launch_vq_argmin_tc(z, E, idx, N, nullptr, dimb::as_stream(stream), trace_dev);
}
extern "C" int dim_debug_vq_argmin_tc_stats(const float* z, const float* E, int64_t* idx, int N, int* stats_dev2, void* stream) {
  if (int e = dimb::ensure_device()) return e;
  return dimb::launch_vq_argmin_tc(z, E, idx, N, stats_dev2, dimb::as_stream(stream));
}
