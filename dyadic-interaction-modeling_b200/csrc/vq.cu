// VQ codebook kernels.
//   vq_argmin_f32 : exact fp32 nearest neighbour.  d_j = (|z|^2 + |e_j|^2) - 2 (z . e_j), first minimum wins
//                   (models/lib/quantizer.py:38-45).  64 tokens x 64 codes per CTA tile, fp32 FFMA.
//   vq_gather     : codes -> codebook rows (replaces the one-hot matmul of quantizer.py:79-90 and
//                   seq2seq_pretrain.py:457-461; also the decoder's token embedding): 8 B index in, D*4 B row out per code;
//                   pure HBM streaming.  Default: vq_gather_pf_kernel (index stream prefetched one chunk ahead, 128-bit
//                   loads/stores, 4 rows in flight per warp): 5.7-5.9 TB/s; variants kept behind dim_debug_vq_gather_mode.
//   vq_gather_bcl / vq_rows_from_bcl : the (B,D,L) channel-major layout VQAutoEncoder.encode returns / decode accepts.
#include "vq.cuh"

namespace dimb {

namespace {

constexpr int TT = 64, TC = 64;

template <int D>
__global__ void __launch_bounds__(256) vq_argmin_f32(const float* __restrict__ z, const float* __restrict__ E,
                                                     int64_t* __restrict__ idx, int N, int K) {
  constexpr int DP = D + 4;
  extern __shared__ __align__(16) float smem[];
  float* Zs = smem;                 // [TT][DP]
  float* Es = Zs + TT * DP;         // [TC][DP]
  float* zz = Es + TC * DP;         // [TT]
  float* ee = zz + TT;              // [TC]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
  const int t0 = blockIdx.x * TT;

  for (int i = tid; i < TT * (D / 4); i += 256) {
    int r = i / (D / 4), c = (i % (D / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t0 + r < N) v = *reinterpret_cast<const float4*>(z + (size_t)(t0 + r) * D + c);
    *reinterpret_cast<float4*>(Zs + r * DP + c) = v;
  }
  __syncthreads();
  for (int r = warp; r < TT; r += 8) {                       // |z|^2 per token (one warp per row)
    float s = 0.f;
    for (int c = lane; c < D; c += 32) { float v = Zs[r * DP + c]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) zz[r] = s;
  }

  float best[4];
  int bidx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { best[i] = INFINITY; bidx[i] = 0x7fffffff; }

  for (int c0 = 0; c0 < K; c0 += TC) {
    __syncthreads();
    for (int i = tid; i < TC * (D / 4); i += 256) {
      int r = i / (D / 4), c = (i % (D / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + r < K) v = __ldg(reinterpret_cast<const float4*>(E + (size_t)(c0 + r) * D + c));
      *reinterpret_cast<float4*>(Es + r * DP + c) = v;
    }
    __syncthreads();
    for (int r = warp; r < TC; r += 8) {
      float s = 0.f;
      for (int c = lane; c < D; c += 32) { float v = Es[r * DP + c]; s = fmaf(v, v, s); }
      s = warp_sum(s);
      if (lane == 0) ee[r] = s;
    }
    float dot[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dot[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; d += 4) {
      float4 a[4], e[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(Zs + (ty + 16 * i) * DP + d);
#pragma unroll
      for (int j = 0; j < 4; ++j) e[j] = *reinterpret_cast<const float4*>(Es + (tx + 16 * j) * DP + d);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dot[i][j] = fmaf(a[i].x, e[j].x, dot[i][j]);
          dot[i][j] = fmaf(a[i].y, e[j].y, dot[i][j]);
          dot[i][j] = fmaf(a[i].z, e[j].z, dot[i][j]);
          dot[i][j] = fmaf(a[i].w, e[j].w, dot[i][j]);
        }
    }
    __syncthreads();                                          // ee visible
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float zi = zz[ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {                           // ascending code index within the thread
        const int code = c0 + tx + 16 * j;
        // (|z|^2 + |e|^2) - 2*dot, each step rounded to fp32 like the reference expression
        const float dist = __fsub_rn(__fadd_rn(zi, ee[tx + 16 * j]), __fmul_rn(2.f, dot[i][j]));
        if (code < K && (dist < best[i] || (dist == best[i] && code < bidx[i]))) { best[i] = dist; bidx[i] = code; }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float bv = best[i];
    int bi = bidx[i];
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    const int t = t0 + ty + 16 * i;
    if (tx == 0 && t < N) idx[t] = bi == 0x7fffffff ? 0 : bi;
  }
}

__global__ void __launch_bounds__(256) vq_gather_kernel(const int64_t* __restrict__ idx, const float* __restrict__ E,
                                                        float* __restrict__ out, int N, int D4, int K,
                                                        int32_t* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  constexpr int R = 4;
  for (size_t r0 = warp * R; r0 < (size_t)N; r0 += nwarps * R) {
    int64_t code[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      int64_t c = (r0 + u < (size_t)N) ? __ldcs(idx + r0 + u) : 0;
      if (c < 0 || c >= K) {
        if (bad && lane == 0) atomicAdd(bad, 1);
        c = c < 0 ? 0 : K - 1;
      }
      code[u] = c;
    }
    for (int c4 = lane; c4 < D4; c4 += 32) {
      float4 v[R];
#pragma unroll
      for (int u = 0; u < R; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(E) + (size_t)code[u] * D4 + c4);
#pragma unroll
      for (int u = 0; u < R; ++u)
        if (r0 + u < (size_t)N) __stcs(reinterpret_cast<float4*>(out) + (r0 + u) * D4 + c4, v[u]);
    }
  }
}

// Same contract, staged through shared memory: a CTA assembles TR consecutive output rows (codebook rows come from L2) and
// hands each finished tile to the TMA engine as ONE contiguous bulk store (cp.async.bulk global <- shared, TR*D*4 bytes),
// NBUF tiles in flight per CTA.  The SM issues no per-row store instructions and HBM sees long full-line write bursts.
template <int TR, int NBUF>
__global__ void __launch_bounds__(256) vq_gather_bulk_kernel(const int64_t* __restrict__ idx, const float* __restrict__ E,
                                                             float* __restrict__ out, int N, int D4, int K,
                                                             int32_t* __restrict__ bad) {
  extern __shared__ __align__(128) float4 gtile[];       // [NBUF][TR * D4]
  const int tid = threadIdx.x;
  const int ntiles = (N + TR - 1) / TR;
  const int per_tile = TR * D4;
  int buf = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // the bulk store issued NBUF tiles ago has finished READING this buffer once <= NBUF-1 newer groups are pending
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
    __syncthreads();
    float4* dst = gtile + (size_t)buf * per_tile;
    const int r0 = t * TR, rows = min(TR, N - r0), n4 = rows * D4;
#pragma unroll 4
    for (int i = tid; i < n4; i += 256) {
      const int r = i / D4, c = i - r * D4;
      int64_t code = __ldg(idx + r0 + r);
      if (code < 0 || code >= K) {
        if (bad && c == 0) atomicAdd(bad, 1);
        code = code < 0 ? 0 : K - 1;
      }
      dst[i] = __ldg(reinterpret_cast<const float4*>(E) + (size_t)code * D4 + c);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the TMA engine
    __syncthreads();
    if (tid == 0) {
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(dst);
      float* g = out + (size_t)r0 * D4 * 4;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(src), "r"(n4 * 16) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    buf = buf + 1 == NBUF ? 0 : buf + 1;
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // smem must outlive the reads; writes complete
}

// Same contract, index stream decoupled from the row stream: a CTA walks chunks of CH codes; the int64 indices of the NEXT
// chunk are fetched into registers before the rows of the current chunk are written and parked in shared memory after it,
// so no row ever waits for an index read that queues behind the write stream in the memory controller.  Codebook rows come
// from L1/L2 (256 KB codebook), each warp keeps R rows (R x 512 B) in flight.
template <int CH, bool STREAM>
__global__ void __launch_bounds__(CH < 256 ? CH : 256) vq_gather_pf_kernel(const int64_t* __restrict__ idx, const float* __restrict__ E,
                                                           float* __restrict__ out, int N, int D4, int K,
                                                           int32_t* __restrict__ bad) {
  __shared__ int codes[2][CH];
  constexpr int NTH = CH < 256 ? CH : 256, PER = CH / NTH, R = 4, NWARP = NTH / 32;      // launched with NTH threads
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchunks = (N + CH - 1) / CH;
  int nbad = 0;
  auto fetch = [&](int chunk, int64_t* reg) {
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const size_t i = (size_t)chunk * CH + tid + NTH * u;
      reg[u] = i < (size_t)N ? __ldcs(idx + i) : 0;
    }
  };
  auto park = [&](int buf, const int64_t* reg) {
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      int64_t c = reg[u];
      if (c < 0 || c >= K) { ++nbad; c = c < 0 ? 0 : K - 1; }
      codes[buf][tid + NTH * u] = (int)c;
    }
  };
  int64_t reg[PER];
  int buf = 0;
  if ((int)blockIdx.x < nchunks) { fetch(blockIdx.x, reg); park(0, reg); }
  __syncthreads();
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int next = chunk + gridDim.x;
    if (next < nchunks) fetch(next, reg);                          // in flight while this chunk's rows are written
    const size_t r0 = (size_t)chunk * CH;
    const int rows = (int)min((size_t)CH, (size_t)N - r0);
    for (int r = warp * R; r < rows; r += NWARP * R) {
      for (int c4 = lane; c4 < D4; c4 += 32) {
        float4 v[R];
#pragma unroll
        for (int u = 0; u < R; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(E) + (size_t)codes[buf][min(r + u, rows - 1)] * D4 + c4);
#pragma unroll
        for (int u = 0; u < R; ++u)
          if (r + u < rows) {
            float4* dst = reinterpret_cast<float4*>(out) + (r0 + r + u) * D4 + c4;
            if (STREAM) __stcs(dst, v[u]); else *dst = v[u];
          }
      }
    }
    if (next < nchunks) park(buf ^ 1, reg);
    __syncthreads();
    buf ^= 1;
  }
  if (bad && nbad) atomicAdd(bad, nbad);
}

// out[b][d][l] = E[idx[b,l]][d]   (32 codes x 32 dims per tile, transposed through shared memory)
__global__ void __launch_bounds__(256) vq_gather_bcl_kernel(const int64_t* __restrict__ idx, const float* __restrict__ E,
                                                            float* __restrict__ out, int L, int D, int K) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    int l = l0 + r;
    float v = 0.f;
    if (l < L && d0 + tx < D) {
      int64_t c = idx[(size_t)b * L + l];
      c = c < 0 ? 0 : (c >= K ? K - 1 : c);
      v = __ldg(E + (size_t)c * D + d0 + tx);
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int d = d0 + r, l = l0 + tx;
    if (d < D && l < L) out[((size_t)b * D + d) * L + l] = tile[tx][r];
  }
}

// rows[b*L + l][d] = q[b][d][l]
__global__ void __launch_bounds__(256) rows_from_bcl_kernel(const float* __restrict__ q, float* __restrict__ rows, int L,
                                                            int D) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    int d = d0 + r, l = l0 + tx;
    tile[r][tx] = (d < D && l < L) ? q[((size_t)b * D + d) * L + l] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int l = l0 + r, d = d0 + tx;
    if (l < L && d < D) rows[((size_t)b * L + l) * D + d] = tile[tx][r];
  }
}

}  // namespace

int launch_vq_argmin(const float* z, const float* E, int64_t* idx, int N, int D, int K, cudaStream_t s) {
  DIM_REQUIRE(N > 0 && K > 0 && K <= 65536, "vq_argmin: bad sizes");
  DIM_REQUIRE(D == 128 || D == 64 || D == 256, "vq_argmin: D must be 64, 128 or 256");
  if (vq_argmin_tc_supported(N, D, K)) return launch_vq_argmin_tc(z, E, idx, N, nullptr, s);     // same indices, ~20x faster (vq_tc.cu)
  dim3 grid(cdiv(N, TT));
  size_t smem = ((size_t)(TT + TC) * (D + 4) + TT + TC) * sizeof(float);
  ProfScope ps(CAT_VQ_ARGMIN, s, (double)N * (4.0 * D + 8.0), 2.0 * N * (double)D * K);     // 520 B / token
#define DIM_ARGMIN_CASE(DD)                                                                                      \
  {                                                                                                              \
    static PerDeviceOnce once;                                                                                   \
    if (once.first())                                                                                            \
      DIM_CHECK_CUDA(cudaFuncSetAttribute(vq_argmin_f32<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                                                                                                                 \
    vq_argmin_f32<DD><<<grid, 256, smem, s>>>(z, E, idx, N, K);                                                  \
  }
  if (D == 64) DIM_ARGMIN_CASE(64) else if (D == 128) DIM_ARGMIN_CASE(128) else DIM_ARGMIN_CASE(256)
#undef DIM_ARGMIN_CASE
  DIM_LAUNCHED();
  return DIM_OK;
}

// tuning hook.  -1: automatic (256-code chunks with a prefetched index stream: 5.7-5.9 TB/s = 89-92 % of the measured copy peak
// at 1-4 M codes, profiles/r01o_vq_roofline.jsonl); 0: warp-per-row, index read in line (4.6 TB/s: every row waits for an
// index read that queues behind the write stream); 1: smem-staged TMA bulk stores (2.9 TB/s); 2..6: prefetched-index variants
int g_vq_gather_mode = -1;

int launch_vq_gather(const int64_t* idx, const float* E, float* out, int N, int D, int K, int32_t* bad, cudaStream_t s) {
  DIM_REQUIRE(N > 0 && D % 4 == 0 && K > 0, "vq_gather: bad sizes");
  ProfScope ps(CAT_VQ_GATHER, s, (double)N * (4.0 * D + 8.0), 0);                          // 520 B / code
  constexpr int TR = 32, NBUF = 4;
  const size_t smem = (size_t)NBUF * TR * D * sizeof(float);
  const bool bulk_ok = ((uintptr_t)out & 15) == 0 && smem <= 96 * 1024;
  const int mode = g_vq_gather_mode >= 0 ? g_vq_gather_mode : (N >= 1024 ? 6 : 0);
  if (mode >= 2) {
    // 2: 512-code chunks, streaming stores; 3: same, plain stores; 4 / 5 / 6: 128 / 64 / 256-code chunks (tuning sweep)
    const int ch = mode == 4 ? 128 : (mode == 5 ? 64 : (mode == 6 ? 256 : 512));
    const int blocks = std::min(cdiv(N, ch), 148 * 8);
    if (mode == 3) vq_gather_pf_kernel<512, false><<<std::min(cdiv(N, 512), 148 * 6), 256, 0, s>>>(idx, E, out, N, D / 4, K, bad);
    else if (mode == 4) vq_gather_pf_kernel<128, true><<<blocks, 128, 0, s>>>(idx, E, out, N, D / 4, K, bad);
    else if (mode == 5) vq_gather_pf_kernel<64, true><<<blocks, 64, 0, s>>>(idx, E, out, N, D / 4, K, bad);
    else if (mode == 6) vq_gather_pf_kernel<256, true><<<blocks, 256, 0, s>>>(idx, E, out, N, D / 4, K, bad);
    else vq_gather_pf_kernel<512, true><<<std::min(cdiv(N, 512), 148 * 6), 256, 0, s>>>(idx, E, out, N, D / 4, K, bad);
  } else if (mode == 1 && bulk_ok) {
    static size_t configured = 48 * 1024;
    if (smem > configured) {
      DIM_CHECK_CUDA(cudaFuncSetAttribute(vq_gather_bulk_kernel<TR, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    const int per_sm = std::max(1, (int)(200 * 1024 / smem));
    const int blocks = std::min(cdiv(N, TR), 148 * per_sm);
    vq_gather_bulk_kernel<TR, NBUF><<<blocks, 256, smem, s>>>(idx, E, out, N, D / 4, K, bad);
  } else {
    int warps_needed = cdiv(N, 4);
    int blocks = std::min(cdiv(warps_needed, 8), 148 * 8);
    vq_gather_kernel<<<blocks, 256, 0, s>>>(idx, E, out, N, D / 4, K, bad);
  }
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_vq_gather_bcl(const int64_t* idx, const float* E, float* out, int B, int L, int D, int K, cudaStream_t s) {
  ProfScope ps(CAT_VQ_GATHER, s, (double)B * L * (4.0 * D + 8.0), 0);
  vq_gather_bcl_kernel<<<dim3(cdiv(L, 32), cdiv(D, 32), B), 256, 0, s>>>(idx, E, out, L, D, K);
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_rows_from_bcl(const float* q, float* rows, int B, int L, int D, cudaStream_t s) {
  ProfScope ps(CAT_MISC, s, 8.0 * B * L * D, 0);
  rows_from_bcl_kernel<<<dim3(cdiv(L, 32), cdiv(D, 32), B), 256, 0, s>>>(q, rows, L, D);
  DIM_LAUNCHED();
  return DIM_OK;
}

}  // namespace dimb

using namespace dimb;

extern "C" int dim_vq_argmin(const float* z, const float* codebook, int64_t* idx, int N, int D, int K, void* stream) {
  if (int e = ensure_device()) return e;
  return launch_vq_argmin(z, codebook, idx, N, D, K, as_stream(stream));
}

// tuning hook (not part of the stable ABI): force a gather implementation (-1 = automatic)
extern "C" int dim_debug_vq_gather_mode(int mode) {
  dimb::g_vq_gather_mode = mode;
  return DIM_OK;
}

extern "C" int dim_vq_gather(const int64_t* idx, const float* codebook, float* out, int N, int D, int K, int32_t* bad,
                             void* stream) {
  if (int e = ensure_device()) return e;
  return launch_vq_gather(idx, codebook, out, N, D, K, bad, as_stream(stream));
}
