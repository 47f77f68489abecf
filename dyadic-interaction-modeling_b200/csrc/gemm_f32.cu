// fp32 FFMA GEMMs with fused prologue/epilogue (see gemm_f32.cuh).  sm_100a.
#include <algorithm>

#include "gemm_f32.cuh"

namespace dimb {

namespace {

constexpr int BK = 16;

__device__ __forceinline__ float epilogue_elem(const GemmArgs& p, float v, int row, int col) {
  if (p.bias) v += __ldg(p.bias + col);
  if (p.tab_mode == 1) {
    int g = row / p.tab_T;
    if (p.tab_index) g = __ldg(p.tab_index + g);
    v = __fadd_rn(v, __ldg(p.tab + (size_t)g * p.ldtab + col));
  } else if (p.tab_mode == 2) {
    int g = row % p.tab_T;
    v = __fadd_rn(v, __fmul_rn(__ldg(p.tab + (size_t)g * p.ldtab + col), p.tab_scale));
  }
  v = act_apply(v, p.act, p.slope);
  if (p.residual) v = __fadd_rn(v, p.residual[(size_t)row * p.ldr + col]);   // plain load: may alias C (in-place residual)
  return v;
}

// Source pointer of A(row, k0..k0+3); handles conv-mode gathering.  k0 % 4 == 0.
__device__ __forceinline__ const float* a_src(const GemmArgs& p, int row, int k0) {
  if (p.conv_T > 0) {
    int tap = k0 / p.conv_C, c = k0 - tap * p.conv_C;
    int b = row / p.conv_T, t = row - b * p.conv_T;
    int L = p.lens ? __ldg(p.lens + b) : p.conv_T;
    int ts = min(max(t + tap - 2, 0), L - 1);
    return p.A + ((size_t)b * p.conv_T + ts) * p.lda + c;
  }
  return p.A + (size_t)row * p.lda + k0;
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) gemm_f32_tiled(const GemmArgs p) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int TMG = TM > 4 ? 2 : 1, TMW = TM / TMG;   // row groups / width
  constexpr int TNG = TN > 4 ? 2 : 1, TNW = TN / TNG;
  static_assert(TNW == 4, "column micro-tile is one float4 per group");
  constexpr int LA = (BM * 4 + NT - 1) / NT, LW = (BN * 4 + NT - 1) / NT;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN + 4];

  pdl_prologue();
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[LA], rw[LW];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int idx = tid + i * NT;
      int r = idx >> 2, kq = (idx & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < BM * 4 && m0 + r < p.M && k0 + kq < p.K) {
        v = __ldg(reinterpret_cast<const float4*>(a_src(p, m0 + r, k0 + kq)));
        if (p.a_add) {
          float4 e = __ldg(reinterpret_cast<const float4*>(p.a_add + k0 + kq));
          v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
        }
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < LW; ++i) {
      int idx = tid + i * NT;
      int r = idx >> 2, kq = (idx & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < BN * 4 && n0 + r < p.N && k0 + kq < p.K)
        v = __ldg(reinterpret_cast<const float4*>(p.W + (size_t)(n0 + r) * p.K + k0 + kq));
      rw[i] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int idx = tid + i * NT;
      if (idx < BM * 4) {
        int r = idx >> 2, kq = (idx & 3) * 4;
        As[kq + 0][r] = ra[i].x; As[kq + 1][r] = ra[i].y; As[kq + 2][r] = ra[i].z; As[kq + 3][r] = ra[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < LW; ++i) {
      int idx = tid + i * NT;
      if (idx < BN * 4) {
        int r = idx >> 2, kq = (idx & 3) * 4;
        Ws[kq + 0][r] = rw[i].x; Ws[kq + 1][r] = rw[i].y; Ws[kq + 2][r] = rw[i].z; Ws[kq + 3][r] = rw[i].w;
      }
    }
  };

  const int nk = (p.K + BK - 1) / BK;
  load_tiles(0);
  store_tiles();
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], w[TN];
#pragma unroll
      for (int g = 0; g < TMG; ++g) {
        const float* src = &As[k][g * (BM / TMG) + ty * TMW];
        if (TMW == 4) {
          float4 v = *reinterpret_cast<const float4*>(src);
          a[g * TMW + 0] = v.x; a[g * TMW + 1] = v.y; a[g * TMW + 2] = v.z; a[g * TMW + 3] = v.w;
        } else {
#pragma unroll
          for (int i = 0; i < TMW; ++i) a[g * TMW + i] = src[i];
        }
      }
#pragma unroll
      for (int g = 0; g < TNG; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&Ws[k][g * (BN / TNG) + tx * TNW]);
        w[g * 4 + 0] = v.x; w[g * 4 + 1] = v.y; w[g * 4 + 2] = v.z; w[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
    if (kt + 1 < nk) {
      store_tiles();
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int row = m0 + (i / TMW) * (BM / TMG) + ty * TMW + (i % TMW);
    if (row >= p.M) continue;
#pragma unroll
    for (int g = 0; g < TNG; ++g) {
      int col = n0 + g * (BN / TNG) + tx * TNW;
      if (col >= p.N) continue;
      float4 o;
      o.x = epilogue_elem(p, acc[i][g * 4 + 0], row, col + 0);
      o.y = epilogue_elem(p, acc[i][g * 4 + 1], row, col + 1);
      o.z = epilogue_elem(p, acc[i][g * 4 + 2], row, col + 2);
      o.w = epilogue_elem(p, acc[i][g * 4 + 3], row, col + 3);
      if (p.C) *reinterpret_cast<float4*>(p.C + (size_t)row * p.ldc + col) = o;
      if (p.Cb) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(p.Cb + (size_t)row * p.ldcb + col) = pk;
      }
    }
  }
}

// GEMV-like GEMM (M <= 8 rows): one CTA per output column, all MT rows; W streamed once with 128-bit loads, A served by L1.
template <int MT>
__global__ void __launch_bounds__(128) gemm_f32_skinny(const GemmArgs p) {
  // One CTA per output column n: its 128 lanes split the K-long weight row, every lane keeps up to U 128-bit loads in flight
  // (the whole row of a K <= 2048 problem is requested before the first FMA), so the N * K * 4 B weight stream -- all there
  // is to a GEMV -- runs with N * K * 4 B / 16 independent requests instead of 4 per warp.  Partial sums: warp shuffle, then
  // 4 warps through shared memory.
  __shared__ float part[4][MT];
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int n = blockIdx.x;
  const int k4n = p.K >> 2;
  const float4* w = reinterpret_cast<const float4*>(p.W + (size_t)n * p.K);
  float acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.f;

  constexpr int U = 4;
  for (int k4 = tid; k4 < k4n; k4 += 128 * U) {
    float4 wv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int kk = k4 + u * 128;
      wv[u] = kk < k4n ? __ldcs(w + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int kk = k4 + u * 128;
      if (kk < k4n) {
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.a_add) e = __ldg(reinterpret_cast<const float4*>(p.a_add) + kk);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          if (m < p.M) {
            float4 av = __ldg(reinterpret_cast<const float4*>(p.A + (size_t)m * p.lda) + kk);
            acc[m] = fmaf(av.x + e.x, wv[u].x, acc[m]);
            acc[m] = fmaf(av.y + e.y, wv[u].y, acc[m]);
            acc[m] = fmaf(av.z + e.z, wv[u].z, acc[m]);
            acc[m] = fmaf(av.w + e.w, wv[u].w, acc[m]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    acc[m] = warp_sum(acc[m]);
    if (lane == 0) part[warp][m] = acc[m];
  }
  __syncthreads();
  if (tid < MT && tid < p.M) {
    const int m = tid;
    const float tot = (part[0][m] + part[1][m]) + (part[2][m] + part[3][m]);
    float o = epilogue_elem(p, tot, m, n);
    if (p.C) p.C[(size_t)m * p.ldc + n] = o;
    if (p.Cb) p.Cb[(size_t)m * p.ldcb + n] = __float2bfloat16_rn(o);
  }
}

// The same GEMV with the weight row read as bf16 (plane 0 of the tensor-core weight image, [N, kp] zero-padded): half the bytes of
// the stream that bounds a <= 8-row decode step.  Used in DIM_PREC_BF16 only (the mode's GEMM operands are bf16 by definition);
// activations stay fp32 here, accumulation fp32.
template <int MT>
__global__ void __launch_bounds__(128) gemm_bf16w_skinny(const GemmArgs p, const __nv_bfloat16* __restrict__ Wb, int kp) {
  __shared__ float part[4][MT];
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int n = blockIdx.x;
  const int k8n = p.K >> 3;                              // whole 8-element vectors; K % 8 == 0 is checked by the launcher
  const uint4* w = reinterpret_cast<const uint4*>(Wb + (size_t)n * kp);
  float acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.f;
  constexpr int U = 4;
  for (int k8 = tid; k8 < k8n; k8 += 128 * U) {
    uint4 wv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kk = k8 + u * 128;
      wv[u] = kk < k8n ? __ldcs(w + kk) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kk = k8 + u * 128;
      if (kk < k8n) {
        const uint32_t ww[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w};
        float wf[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) { wf[2 * i] = __uint_as_float(ww[i] << 16); wf[2 * i + 1] = __uint_as_float(ww[i] & 0xffff0000u); }
        float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0;
        if (p.a_add) { e0 = __ldg(reinterpret_cast<const float4*>(p.a_add) + 2 * kk); e1 = __ldg(reinterpret_cast<const float4*>(p.a_add) + 2 * kk + 1); }
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          if (m < p.M) {
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.A + (size_t)m * p.lda) + 2 * kk);
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.A + (size_t)m * p.lda) + 2 * kk + 1);
            acc[m] = fmaf(a0.x + e0.x, wf[0], acc[m]); acc[m] = fmaf(a0.y + e0.y, wf[1], acc[m]);
            acc[m] = fmaf(a0.z + e0.z, wf[2], acc[m]); acc[m] = fmaf(a0.w + e0.w, wf[3], acc[m]);
            acc[m] = fmaf(a1.x + e1.x, wf[4], acc[m]); acc[m] = fmaf(a1.y + e1.y, wf[5], acc[m]);
            acc[m] = fmaf(a1.z + e1.z, wf[6], acc[m]); acc[m] = fmaf(a1.w + e1.w, wf[7], acc[m]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    acc[m] = warp_sum(acc[m]);
    if (lane == 0) part[warp][m] = acc[m];
  }
  __syncthreads();
  if (tid < MT && tid < p.M) {
    const int m = tid;
    const float tot = (part[0][m] + part[1][m]) + (part[2][m] + part[3][m]);
    float o = epilogue_elem(p, tot, m, n);
    if (p.C) p.C[(size_t)m * p.ldc + n] = o;
    if (p.Cb) p.Cb[(size_t)m * p.ldcb + n] = __float2bfloat16_rn(o);
  }
}

__global__ void repack_conv_kernel(const float* __restrict__ w, float* __restrict__ o, int Cout, int Cin) {
  size_t n = (size_t)Cout * Cin * 5;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = i % Cin;
    int tap = (i / Cin) % 5;
    int co = i / ((size_t)Cin * 5);
    o[i] = w[((size_t)co * Cin + c) * 5 + tap];
  }
}

// Linear with no alignment requirement on K, N or the leading dimensions (the 70110-wide mesh layers of EmocaConverter,
// /root/reference/code/seq2seq_pretrain.py:777, :803-812): 64 x 64 tiles, scalar guarded loads and stores, ascending-k fmaf
// chains.  gridDim.z > 1 splits K: slice z writes its partial tile to part[z][M][N]; ragged_reduce_kernel adds the slices in
// ascending z, then bias and activation (deterministic).
template <int V>   // V consecutive k per load: 4 when every row of A and W is 16-byte aligned, 2 when 8-byte aligned, else 1
__global__ void __launch_bounds__(256) gemm_f32_ragged(const GemmArgs p, int ldw, int kslice, float* part) {
  constexpr int BM = 64, BN = 64, NL = 4 / V, KV = BK / V;
  __shared__ float As[BK][BM + 1];
  __shared__ float Ws[BK][BN + 1];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * kslice, ke = min(p.K, kb + kslice);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[NL][V], rw[NL][V];
  auto ldv = [&](const float* src, int k, float* dst) {       // V elements at src[k..k+V), zero beyond ke
    if (k + V <= ke) {
      if constexpr (V == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + k));
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
      } else if constexpr (V == 2) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(src + k));
        dst[0] = v.x; dst[1] = v.y;
      } else {
        dst[0] = __ldg(src + k);
      }
    } else {
#pragma unroll
      for (int e = 0; e < V; ++e) dst[e] = k + e < ke ? __ldg(src + k + e) : 0.f;
    }
  };
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const int idx = tid + i * 256, r = idx / KV, k = k0 + (idx % KV) * V;
#pragma unroll
      for (int e = 0; e < V; ++e) { ra[i][e] = 0.f; rw[i][e] = 0.f; }
      if (m0 + r < p.M) ldv(p.A + (size_t)(m0 + r) * p.lda, k, ra[i]);
      if (n0 + r < p.N) ldv(p.W + (size_t)(n0 + r) * ldw, k, rw[i]);
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const int idx = tid + i * 256, r = idx / KV, k = (idx % KV) * V;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        As[k + e][r] = ra[i][e];
        Ws[k + e][r] = rw[i][e];
      }
    }
  };
  load_tiles(kb);
  store_tiles();
  __syncthreads();
  for (int k0 = kb; k0 < ke; k0 += BK) {
    if (k0 + BK < ke) load_tiles(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; w[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
    if (k0 + BK < ke) {
      store_tiles();
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= p.N) continue;
      if (part) part[((size_t)blockIdx.z * p.M + row) * p.N + col] = acc[i][j];
      else p.C[(size_t)row * p.ldc + col] = epilogue_elem(p, acc[i][j], row, col);
    }
  }
}

__global__ void ragged_reduce_kernel(const GemmArgs p, const float* part, int splits) {
  const size_t n = (size_t)p.M * p.N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += part[(size_t)z * n + i];
    const int row = (int)(i / p.N), col = (int)(i - (size_t)row * p.N);
    p.C[(size_t)row * p.ldc + col] = epilogue_elem(p, v, row, col);
  }
}

int ragged_splits(int M, int N, int K) {
  const long tiles = (long)cdiv(M, 64) * cdiv(N, 64);
  if (tiles >= 148 || K < 2048) return 1;
  long s = (2 * 148 + tiles - 1) / tiles;
  return (int)std::min<long>(s, K / 512);
}

template <int BM, int BN, int TM, int TN>
int launch_tiled(const GemmArgs& a, cudaStream_t s) {
  dim3 grid(cdiv(a.N, BN), cdiv(a.M, BM));
  ProfScope ps(a.conv_T > 0 ? CAT_CONV : CAT_GEMM_TILED, s, 4.0 * ((double)a.M * (a.conv_T > 0 ? a.conv_C : a.K) + (double)a.N * a.K + (double)a.M * a.N),
               2.0 * a.M * (double)a.N * a.K);
  DIM_CHECK_CUDA(launch_k(gemm_f32_tiled<BM, BN, TM, TN>, grid, dim3((BM / TM) * (BN / TN)), 0, s, a));
  DIM_LAUNCHED();
  return DIM_OK;
}

}  // namespace

// <= 8 rows with bf16 weights (plane 0 of the tensor-core weight image)
int launch_gemv_bf16w(const GemmArgs& a, const __nv_bfloat16* Wb, int kp, cudaStream_t s) {
  DIM_REQUIRE(a.M > 0 && a.M <= 8 && a.N > 0 && a.K > 0 && a.K % 8 == 0 && a.conv_T == 0, "gemv_bf16w: bad problem");
  DIM_REQUIRE(a.lda % 4 == 0 && (a.C != nullptr || a.Cb != nullptr), "gemv_bf16w: bad operands");
  dim3 grid(a.N);
  ProfScope ps(CAT_GEMM_SKINNY, s, 4.0 * ((double)a.M * a.K + (double)a.M * a.N) + 2.0 * (double)a.N * a.K, 2.0 * a.M * (double)a.N * a.K);
  if (a.M == 1) DIM_CHECK_CUDA(launch_k(gemm_bf16w_skinny<1>, grid, dim3(128), 0, s, a, Wb, kp));
  else if (a.M == 2) DIM_CHECK_CUDA(launch_k(gemm_bf16w_skinny<2>, grid, dim3(128), 0, s, a, Wb, kp));
  else if (a.M <= 4) DIM_CHECK_CUDA(launch_k(gemm_bf16w_skinny<4>, grid, dim3(128), 0, s, a, Wb, kp));
  else DIM_CHECK_CUDA(launch_k(gemm_bf16w_skinny<8>, grid, dim3(128), 0, s, a, Wb, kp));
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_gemm_f32(const GemmArgs& a, cudaStream_t s) {
  DIM_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem");
  DIM_REQUIRE(a.K % 4 == 0 && a.N % 4 == 0, "gemm: K and N must be multiples of 4");
  DIM_REQUIRE(a.lda % 4 == 0 && (a.C == nullptr || a.ldc % 4 == 0), "gemm: leading dims must be multiples of 4");
  DIM_REQUIRE(a.C != nullptr || a.Cb != nullptr, "gemm: no output");
  if (a.conv_T > 0) DIM_REQUIRE(a.conv_C % 4 == 0 && a.K == 5 * a.conv_C, "gemm: conv mode needs Cin % 4 == 0");
  if (a.M <= 8 && a.conv_T == 0) {
    dim3 grid(a.N);                                    // one CTA per output column
    ProfScope ps(CAT_GEMM_SKINNY, s, 4.0 * ((double)a.M * a.K + (double)a.N * a.K + (double)a.M * a.N),
                 2.0 * a.M * (double)a.N * a.K);
    if (a.M == 1) DIM_CHECK_CUDA(launch_k(gemm_f32_skinny<1>, grid, dim3(128), 0, s, a));
    else if (a.M == 2) DIM_CHECK_CUDA(launch_k(gemm_f32_skinny<2>, grid, dim3(128), 0, s, a));
    else if (a.M <= 4) DIM_CHECK_CUDA(launch_k(gemm_f32_skinny<4>, grid, dim3(128), 0, s, a));
    else DIM_CHECK_CUDA(launch_k(gemm_f32_skinny<8>, grid, dim3(128), 0, s, a));
    DIM_LAUNCHED();
    return DIM_OK;
  }
  // Pick the tile so that the grid covers the 148 SMs when it can.
  long tiles128 = (long)cdiv(a.M, 128) * cdiv(a.N, 128);
  long tiles64 = (long)cdiv(a.M, 64) * cdiv(a.N, 64);
  if (tiles128 >= 148) return launch_tiled<128, 128, 8, 8>(a, s);
  if (tiles64 >= 96 || a.M > 32) return launch_tiled<64, 64, 4, 4>(a, s);
  return launch_tiled<32, 64, 2, 4>(a, s);
}

}  // namespace dimb

// ---- C ABI ------------------------------------------------------------------------------------------------------
using namespace dimb;

extern "C" int dim_linear_f32(const float* A, int lda, const float* W, const float* bias, const float* residual,
                              int ldr, float* C, int ldc, int M, int N, int K, int act, float slope, void* stream) {
  if (int e = ensure_device()) return e;
  GemmArgs a;
  a.A = A; a.lda = lda; a.W = W; a.bias = bias; a.residual = residual; a.ldr = ldr; a.C = C; a.ldc = ldc;
  a.M = M; a.N = N; a.K = K; a.act = act; a.slope = slope;
  return launch_gemm_f32(a, as_stream(stream));
}

extern "C" int dim_conv5_leaky_f32(const float* x, const float* Wr, const float* bias, const int32_t* lens, float* y,
                                   int B, int T, int C, float slope, void* stream) {
  if (int e = ensure_device()) return e;
  GemmArgs a;
  a.A = x; a.lda = C; a.W = Wr; a.bias = bias; a.C = y; a.ldc = C;
  a.M = B * T; a.N = C; a.K = 5 * C; a.act = DIM_ACT_LEAKY; a.slope = slope;
  a.conv_T = T; a.conv_C = C; a.lens = lens;
  return launch_gemm_f32(a, as_stream(stream));
}


extern "C" int dim_repack_conv_weight(const float* w_oik, float* w_oki, int Cout, int Cin, void* stream) {
  if (int e = ensure_device()) return e;
  DIM_REQUIRE(Cout > 0 && Cin > 0, "repack: bad shape");
  ProfScope ps(CAT_MISC, as_stream(stream), 8.0 * Cout * Cin * 5, 0);
  repack_conv_kernel<<<148 * 4, 256, 0, as_stream(stream)>>>(w_oik, w_oki, Cout, Cin);
  DIM_LAUNCHED();
  return DIM_OK;
}

extern "C" size_t dim_linear_ragged_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int sp = ragged_splits(M, N, K);
  return sp > 1 ? (size_t)sp * M * N * sizeof(float) : 0;
}

extern "C" int dim_linear_ragged_f32(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc,
                                     int M, int N, int K, int act, float slope, void* ws, size_t ws_bytes, void* stream) {
  if (int e = ensure_device()) return e;
  DIM_REQUIRE(A && W && C && M > 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldc >= N, "linear_ragged: bad operands");
  GemmArgs a;
  a.A = A; a.lda = lda; a.W = W; a.bias = bias; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K; a.act = act; a.slope = slope;
  cudaStream_t s = as_stream(stream);
  const int sp = ragged_splits(M, N, K);
  float* part = nullptr;
  int kslice = K;
  if (sp > 1) {
    DIM_REQUIRE(ws != nullptr && ws_bytes >= (size_t)sp * M * N * sizeof(float), "linear_ragged: workspace too small");
    part = static_cast<float*>(ws);
    kslice = cdiv(cdiv(K, sp), BK) * BK;
  }
  ProfScope ps(CAT_GEMM_TILED, s, 4.0 * ((double)M * K + (double)N * K + (double)M * N), 2.0 * M * (double)N * K);
  const dim3 grid(cdiv(N, 64), cdiv(M, 64), cdiv(K, kslice));
  const uintptr_t bits = reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | (uintptr_t)lda * 4 | (uintptr_t)ldw * 4;
  if (bits % 16 == 0) gemm_f32_ragged<4><<<grid, 256, 0, s>>>(a, ldw, kslice, part);        // kslice % 16 == 0: slices stay aligned
  else if (bits % 8 == 0) gemm_f32_ragged<2><<<grid, 256, 0, s>>>(a, ldw, kslice, part);
  else gemm_f32_ragged<1><<<grid, 256, 0, s>>>(a, ldw, kslice, part);
  DIM_LAUNCHED();
  if (part) {
    ragged_reduce_kernel<<<std::min(148 * 4, cdiv(M * N, 256)), 256, 0, s>>>(a, part, cdiv(K, kslice));
    DIM_LAUNCHED();
  }
  return DIM_OK;
}
