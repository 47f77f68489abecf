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
//       + 2^-16 |z| |e|max + 2^-21 (|z|^2 + |e|^2max)   rounding inside the exact kernel's own fp32 expression (generous)
//   => |s_k - (d_k - |z|^2)| <= eps for the d_k the exact kernel computes, hence its argmin k* satisfies s_k* <= min_k s_k + 2 eps.
// Every code within 2 eps of the approximate minimum is re-evaluated with the SAME arithmetic as vq_argmin_f32 (|z|^2 by the same
// warp reduction, |e_k|^2 by the same warp reduction, the dot product as one sequential fmaf chain, (zz + ee) - 2 dot with each
// step rounded); the winner is the lexicographic minimum of (d, k) exactly like the exact kernel.  A lone candidate needs no
// distance.  On the benchmark distribution ~4 % of the tokens shortlist more than one code (bf16 operands: 50 %).
//
// Persistent CTAs (one per SM), 16 warps:
//   warps 0-7  epilogue: thread = token row (TMEM lane) x one half of the codes (256 TMEM columns); pass 1 = approximate minimum,
//              pass 2 = shortlist into a shared (row, code) list; the accumulator is released; the list is then re-ranked by all
//              256 threads together (one exact distance per thread per round -- no divergent per-token serial loops);
//   warps 8-15 converters: fp32 z rows -> fp16, 128-byte-swizzled K-major smem tiles (double buffered), |z|^2 and |dz|^2;
//              thread 256 also issues the MMAs (code half 0 first, so its epilogue overlaps the MMAs of half 1).
// The fp16 codebook (128 KB, swizzled) stays in shared memory for the whole kernel.
#include <cuda_fp16.h>

#include "tc_ptx.cuh"
#include "vq.cuh"

namespace dimb {

namespace {

constexpr int VT_D = 128, VT_K = 512, VT_TM = 128;                 // only this shape (the DIM codebook): 512 codes x 128 dims
constexpr uint32_t VT_E_BYTES = VT_K * VT_D * 2;                   // 128 KB: 2 k-blocks x [512 rows x 128 B]
constexpr uint32_t VT_Z_BYTES = VT_TM * VT_D * 2;                  // 32 KB per buffer: 2 k-blocks x [128 rows x 128 B]
constexpr uint32_t VT_SMEM = VT_E_BYTES + 2 * VT_Z_BYTES + 1024;   // + alignment slack
constexpr int VT_CAP = 4096;                                       // (row, code) pairs re-ranked cooperatively per tile
constexpr int VT_THREADS = 512, VT_EPI = 256;

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

__device__ __forceinline__ void bar_epi() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 256 epilogue threads only

__device__ __forceinline__ unsigned long long pack_dk(float d, int k) {      // orders like (d, k) lexicographically
  const uint32_t b = __float_as_uint(d);
  const uint32_t key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)key << 32) | (uint32_t)k;
}

// vq_argmin_f32's expression, operation for operation
__device__ __forceinline__ float exact_dist(const float* __restrict__ z, const float* __restrict__ E, int t, int code, float zz, float ee) {
  const float4* zr = reinterpret_cast<const float4*>(z + (size_t)t * VT_D);
  const float4* er = reinterpret_cast<const float4*>(E + (size_t)code * VT_D);
  float dot = 0.f;
#pragma unroll 8
  for (int d4 = 0; d4 < VT_D / 4; ++d4) {
    const float4 a = zr[d4], e = __ldg(er + d4);
    dot = fmaf(a.x, e.x, dot); dot = fmaf(a.y, e.y, dot); dot = fmaf(a.z, e.z, dot); dot = fmaf(a.w, e.w, dot);
  }
  return __fsub_rn(__fadd_rn(zz, ee), __fmul_rn(2.f, dot));
}

__global__ void __launch_bounds__(VT_THREADS, 1) vq_argmin_tc(const float* __restrict__ z, const float* __restrict__ E,
                                                              const float* __restrict__ e2g, const uint8_t* __restrict__ Eimg,
                                                              const float* __restrict__ cst, int64_t* __restrict__ idx, int N,
                                                              int* __restrict__ stats) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t zfull[2], zfree[2], acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float e2s[VT_K];
  __shared__ float z2s[4][VT_TM], dz2s[4][VT_TM];   // |z|^2, |z - fp16(z)|^2 of the tiles in flight (a slot is rewritten 4 tiles later)
  __shared__ float minpart[2][VT_TM];
  __shared__ __align__(8) unsigned long long rowbest[VT_TM];
  __shared__ int rowcnt[VT_TM], rowfirst[VT_TM];
  __shared__ uint16_t pairs[VT_CAP];
  __shared__ int npairs;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t e_u = base, z_u = base + VT_E_BYTES;
  uint8_t* zbuf = smem + VT_E_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (N + VT_TM - 1) / VT_TM;

  // ---- set-up: codebook image and |e|^2 into shared memory, barriers, TMEM (all 512 columns)
  for (uint32_t i = tid; i < VT_E_BYTES / 16; i += VT_THREADS) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(Eimg)[i];
  for (int i = tid; i < VT_K; i += VT_THREADS) e2s[i] = e2g[i];
  if (tid < VT_TM) { rowbest[tid] = ~0ull; rowcnt[tid] = 0; rowfirst[tid] = 0x7fffffff; }
  if (tid == 0) {
    npairs = 0;
    mbar_init(smem_u32(&zfull[0]), 256); mbar_init(smem_u32(&zfull[1]), 256);      // 256 converter threads arrive
    mbar_init(smem_u32(&zfree[0]), 1); mbar_init(smem_u32(&zfree[1]), 1);          // tcgen05.commit arrives
    mbar_init(smem_u32(&acc_full[0]), 1); mbar_init(smem_u32(&acc_full[1]), 1);    // tcgen05.commit arrives (per code half)
    mbar_init(smem_u32(&acc_free[0]), 128); mbar_init(smem_u32(&acc_free[1]), 128);  // the 128 epilogue threads of a code half arrive
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

  if (warp >= 8) {
    // ===== converters: tile t -> buffer t & 1; thread 256 also issues the MMAs of the tile it has just helped to convert =====
    const int cw = warp - 8;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      mbar_wait(smem_u32(&zfree[b]), ((it >> 1) & 1) ^ 1);          // the MMAs that read this buffer two tiles ago are done
      uint8_t* zb = zbuf + b * VT_Z_BYTES;
      // one warp per token row (vq_argmin_f32's |z|^2 reduction order), 16 rows = 64 coalesced 128-byte loads in flight per warp
      float v[16][VT_D / 32];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int t = tile * VT_TM + cw + 8 * u;
#pragma unroll
        for (int j = 0; j < VT_D / 32; ++j) v[u][j] = t < N ? __ldcs(z + (size_t)t * VT_D + lane + 32 * j) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int r = cw + 8 * u;
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int j = 0; j < VT_D / 32; ++j) {
          s = fmaf(v[u][j], v[u][j], s);
          const __half h = __float2half_rn(v[u][j]);
          const float rem = v[u][j] - __half2float(h);
          q = fmaf(rem, rem, q);
          *reinterpret_cast<__half*>(zb + swz_off(VT_TM, r, lane + 32 * j)) = h;
        }
        s = warp_sum(s);
        q = warp_sum(q);
        if (lane == 0) { z2s[it & 3][r] = s; dz2s[it & 3][r] = q; }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic smem writes -> visible to the tensor core (async proxy)
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&zfull[b])) : "memory");
      if (tid == 256) {
        mbar_wait(smem_u32(&zfull[b]), (it >> 1) & 1);
        // D[128 tokens x 256 codes] per instruction: M = 128, N = 256, fp16 x fp16 -> fp32, both operands K-major
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(smem_u32(&acc_free[h]), (it & 1) ^ 1);          // the epilogue has drained this half of the previous tile
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
      __syncwarp();
    }
  } else {
    // ===== epilogue: thread = token row (TMEM lane) x code half =====
    const int q = warp & 3, hf = warp >> 2;
    const int row = q * 32 + lane, cbase = hf * 256;
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cbase;
    const float ehmax = sqrtf(cst[0]), demax = sqrtf(cst[1]), e2max = cst[2], emax = sqrtf(cst[2]);
    int it = 0;
    int n_multi = 0, n_cand = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int t0 = tile * VT_TM;
      mbar_wait(smem_u32(&acc_full[hf]), it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const float zz = z2s[it & 3][row], dz2 = dz2s[it & 3][row];
      const float zn = sqrtf(zz), dzn = sqrtf(dz2);
      const float eps = 2.f * (dzn * ehmax + zn * demax) + 2.44140625e-4f * (zn + dzn) * ehmax + 1.52587891e-5f * zn * emax +
                        4.76837158e-7f * (zz + e2max);
      const float tau = 2.002f * eps;                                 // 2 eps (header) + margin for the rounding of this expression
      // pass 1: the approximate minimum over this thread's 256 codes
      float m4[4] = {INFINITY, INFINITY, INFINITY, INFINITY};         // four independent chains: the minimum is latency-bound otherwise
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 16) {
        float v[16];
        tmem_ld16(tl + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) m4[j & 3] = fminf(m4[j & 3], fmaf(-2.f, v[j], e2s[cbase + c0 + j]));
      }
      minpart[hf][row] = fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3]));
      bar_epi();
      const float lim = fminf(minpart[0][row], minpart[1][row]) + tau;
      // pass 2: everything within tau of it goes on the list (NaN/inf scores or limits shortlist the code: slow, still exact)
      bool overflow = false;
      const bool valid = t0 + row < N;                                // rows past the end of z shortlist nothing
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 16) {
        float v[16];
        tmem_ld16(tl + (uint32_t)c0, v);
        unsigned mask = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) mask |= (!(fmaf(-2.f, v[j], e2s[cbase + c0 + j]) > lim) ? 1u : 0u) << j;
        if (!valid) mask = 0;
        while (mask) {
          const int j = __ffs(mask) - 1;
          mask &= mask - 1;
          const int k = cbase + c0 + j;
          atomicAdd(&rowcnt[row], 1);
          atomicMin(&rowfirst[row], k);
          const int slot = atomicAdd(&npairs, 1);
          if (slot < VT_CAP) pairs[slot] = (uint16_t)((row << 9) | k);
        }
      }
      bar_epi();
      const int np = npairs;
      overflow = np > VT_CAP;
      if (overflow) {
        // the list did not hold this tile's shortlist (degenerate data): every thread re-scans its columns and evaluates its own
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 16) {
          float v[16];
          tmem_ld16(tl + (uint32_t)c0, v);
          unsigned mask = 0;
#pragma unroll
          for (int j = 0; j < 16; ++j) mask |= (!(fmaf(-2.f, v[j], e2s[cbase + c0 + j]) > lim) ? 1u : 0u) << j;
          while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const int k = cbase + c0 + j;
            if (t0 + row < N && rowcnt[row] > 1) atomicMin(&rowbest[row], pack_dk(exact_dist(z, E, t0 + row, k, zz, e2s[k]), k));
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_free[hf])) : "memory");
      if (!overflow) {
        // cooperative exact re-rank: one (row, code) pair per thread per round
        for (int p = tid; p < np; p += VT_EPI) {
          const int pr = pairs[p] >> 9, k = pairs[p] & 511;
          if (rowcnt[pr] > 1 && t0 + pr < N)
            atomicMin(&rowbest[pr], pack_dk(exact_dist(z, E, t0 + pr, k, z2s[it & 3][pr], e2s[k]), k));
        }
      }
      bar_epi();
      if (hf == 0) {
        const int cnt = rowcnt[row], t = t0 + row;
        if (t < N) {
          idx[t] = cnt > 1 ? (int)(uint32_t)(rowbest[row] & 0xffffffffull) : rowfirst[row];
          n_multi += cnt > 1;
          n_cand += cnt;
        }
        rowbest[row] = ~0ull; rowcnt[row] = 0; rowfirst[row] = 0x7fffffff;
        if (row == 0) npairs = 0;
      }
      // the next tile's list writes come after its first bar_epi(), i.e. after every thread has passed this point
    }
    if (stats) {
      n_multi = __reduce_add_sync(0xffffffffu, n_multi);
      n_cand = __reduce_add_sync(0xffffffffu, n_cand);
      if (lane == 0) { atomicAdd(stats, n_multi); atomicAdd(stats + 1, n_cand); }
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

int g_vq_argmin_impl = 0;               // 0: tensor-core shortlist + exact re-rank when the shape allows; 1: exact FFMA kernel only

bool vq_argmin_tc_supported(int N, int D, int K) { return g_vq_argmin_impl == 0 && D == VT_D && K == VT_K && N >= 4 * VT_TM; }

// The codebook images are rebuilt on every call (4 us): the weights are borrowed pointers and may have been overwritten in place.
int launch_vq_argmin_tc(const float* z, const float* E, int64_t* idx, int N, int* stats, cudaStream_t s) {
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
  vq_argmin_tc<<<std::min(sms, ntiles), VT_THREADS, VT_SMEM, s>>>(z, E, cb->e2, cb->img, cb->cst, idx, N, stats);
  DIM_LAUNCHED();
  return DIM_OK;
}

}  // namespace dimb

// tuning / test hooks (not part of the stable ABI): force the exact FFMA kernel (1) or allow the tensor-core path (0);
// run the tensor-core kernel and return {tokens with more than one shortlisted code, shortlisted codes in total}
extern "C" int dim_debug_vq_argmin_impl(int impl) {
  dimb::g_vq_argmin_impl = impl;
  return DIM_OK;
}
extern "C" int dim_debug_vq_argmin_tc_stats(const float* z, const float* E, int64_t* idx, int N, int* stats_dev2, void* stream) {
  if (int e = dimb::ensure_device()) return e;
  return dimb::launch_vq_argmin_tc(z, E, idx, N, stats_dev2, dimb::as_stream(stream));
}
