// Attention kernels.
//   attn_prefill_mma : prefill attention on the tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate) with Q, K, V and the
//                      probabilities split exactly into 1 (bf16 mode) or 3 (fp32-grade) bf16 planes; online softmax in
//                      registers, 64 x 64 tiles.  VQ-VAE layers (Dh 48, scale hidden^-0.5, no mask;
//                      models/lib/base_models.py:136-143), x-transformers encoders (Dh 64, causal + key-padding mask,
//                      -FLT_MAX fill) and the teacher-forced decoder (self: causal + kv mask; cross: Tq != Tk).
//   attn_prefill_f32 : the same contract on FFMA (DIM_PREC_FP32, operator-level ABI dim_attention_f32).
//   attn_decode_kernel / attn_decode_lanes : one query per (clip, head) against a HEAD-MAJOR K/V cache [B,H,tokens,64]
//                      (self attention with in-kernel append of the new key/value, or cross attention over the once-projected
//                      context, optionally shared by several sample rows of a clip); HBM-bound streaming kernels, see the
//                      comments at each kernel and DESIGN.md 5.3.
//   kv_head_major    : token-major projection output -> head-major caches.
#include "attention.cuh"
#include "kvio.cuh"

#include <algorithm>
#include <cstdlib>

namespace dimb {

namespace {

constexpr int TQ = 64, TKV = 64;

template <int DH>
__global__ void __launch_bounds__(256) attn_prefill_f32(const AttnArgs p) {
  constexpr int DP = DH + 4;            // padded row: conflict-free 128-bit reads at stride 16 rows
  constexpr int DJ = DH / 16;           // output columns per thread
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                     // [TQ][DP]
  float* Ks = Qs + TQ * DP;             // [TKV][DP]
  float* Vs = Ks + TKV * DP;            // [TKV][DH]
  float* Ps = Vs + TKV * DH;            // [TQ][TKV+4]
  constexpr int PP = TKV + 4;

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TQ;
  const float* qb = p.q + (size_t)b * p.Tq * p.ldq + h * DH;
  const float* kb = p.k + (size_t)b * p.Tk * p.ldk + h * DH;
  const float* vb = p.v + (size_t)b * p.Tk * p.ldv + h * DH;
  const int klen = p.lens ? min(__ldg(p.lens + b), p.Tk) : p.Tk;     // keys >= klen do not exist
  const uint8_t* km = p.key_mask ? p.key_mask + (size_t)b * p.Tk : nullptr;

  for (int i = tid; i < TQ * (DH / 4); i += 256) {
    int r = i / (DH / 4), c = (i % (DH / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < p.Tq) v = *reinterpret_cast<const float4*>(qb + (size_t)(q0 + r) * p.ldq + c);
    *reinterpret_cast<float4*>(Qs + r * DP + c) = v;
  }

  float m_run[4], l_run[4], o[4][DJ];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < DJ; ++j) o[i][j] = 0.f;
  }

  int nkt = (klen + TKV - 1) / TKV;
  if (p.causal) nkt = min(nkt, (min(q0 + TQ, p.Tq) - 1) / TKV + 1);   // tiles right of the diagonal contribute exp(-FLT_MAX - m) = 0

  for (int kt = 0; kt < nkt; ++kt) {
    const int k0 = kt * TKV;
    __syncthreads();                    // previous tile's Ps / Vs fully consumed (and Qs written, first time)
    for (int i = tid; i < TKV * (DH / 4); i += 256) {
      int r = i / (DH / 4), c = (i % (DH / 4)) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < klen) {
        kv = *reinterpret_cast<const float4*>(kb + (size_t)(k0 + r) * p.ldk + c);
        vv = *reinterpret_cast<const float4*>(vb + (size_t)(k0 + r) * p.ldv + c);
      }
      *reinterpret_cast<float4*>(Ks + r * DP + c) = kv;
      *reinterpret_cast<float4*>(Vs + r * DH + c) = vv;
    }
    __syncthreads();

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      float4 qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * DP + d);
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * DP + d);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(qv[i].x, kv[j].x, s[i][j]);
          s[i][j] = fmaf(qv[i].y, kv[j].y, s[i][j]);
          s[i][j] = fmaf(qv[i].z, kv[j].z, s[i][j]);
          s[i][j] = fmaf(qv[i].w, kv[j].w, s[i][j]);
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + ty + 16 * i;
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kj = k0 + tx + 16 * j;
        float v = s[i][j] * p.scale;
        if (kj >= klen) v = -INFINITY;                                  // not a key at all
        else if ((km && !km[kj]) || (p.causal && kj > qi)) v = -FLT_MAX;   // masked_fill(-finfo.max)
        s[i][j] = v;
        tmax = fmaxf(tmax, v);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
      const float m_new = fmaxf(m_run[i], tmax);
      const float alpha = (m_run[i] == -INFINITY) ? 0.f : expf(m_run[i] - m_new);
      float psum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float pv = (s[i][j] == -INFINITY) ? 0.f : expf(s[i][j] - m_new);
        Ps[(ty + 16 * i) * PP + tx + 16 * j] = pv;
        psum += pv;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, off);
      l_run[i] = l_run[i] * alpha + psum;
      m_run[i] = m_new;
#pragma unroll
      for (int j = 0; j < DJ; ++j) o[i][j] *= alpha;
    }
    __syncthreads();

#pragma unroll 4
    for (int kk = 0; kk < TKV; kk += 4) {
      float4 pv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * PP + kk);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float vv[DJ];
#pragma unroll
        for (int j = 0; j < DJ; ++j) vv[j] = Vs[(kk + u) * DH + tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float pe = u == 0 ? pv[i].x : (u == 1 ? pv[i].y : (u == 2 ? pv[i].z : pv[i].w));
#pragma unroll
          for (int j = 0; j < DJ; ++j) o[i][j] = fmaf(pe, vv[j], o[i][j]);
        }
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + ty + 16 * i;
    if (qi >= p.Tq) continue;
    const float inv = l_run[i] > 0.f ? 1.f / l_run[i] : 0.f;
    if (p.out) {
      float* orow = p.out + ((size_t)b * p.Tq + qi) * p.ldo + h * DH;
#pragma unroll
      for (int j = 0; j < DJ; ++j) orow[tx + 16 * j] = o[i][j] * inv;
    }
    if (p.out_p) {
      __nv_bfloat16* prow = p.out_p + ((size_t)b * p.Tq + qi) * p.planes * p.kp + h * DH;
#pragma unroll
      for (int j = 0; j < DJ; ++j) store_planes1(prow + tx + 16 * j, o[i][j] * inv, p.planes, p.kp);
    }
  }
}


// ---- prefill on the tensor cores -------------------------------------------------------------------------------------
// Same contract as attn_prefill_f32, arithmetic on mma.sync.m16n8k16 (bf16 operands, fp32 accumulate).  Q, K, V and the
// probabilities P are split EXACTLY into NP bf16 planes (x = h + m + l, round-to-nearest at each step); the plane products
// with index sum < NP are accumulated: NP = 3 keeps 24 mantissa bits per operand (fp32-grade scores and outputs, the parity
// mode of the tcgen05 GEMMs), NP = 1 is plain bf16.  64 queries x 64 keys per CTA tile, 4 warps x 16 query rows, online
// softmax in registers (the m16n8 accumulator layout of S is re-used as the A-operand layout of P), the next K/V tile is
// prefetched into registers while the current one is multiplied.
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
// fp32 pair -> NP packed bf16x2 plane words (low half = first element)
template <int NP>
__device__ __forceinline__ void split2(float x, float y, uint32_t* w) {
#pragma unroll
  for (int pl = 0; pl < NP; ++pl) {
    const __nv_bfloat16 bx = __float2bfloat16_rn(x), by = __float2bfloat16_rn(y);
    w[pl] = (uint32_t)__bfloat16_as_ushort(bx) | ((uint32_t)__bfloat16_as_ushort(by) << 16);
    x -= __bfloat162float(bx);
    y -= __bfloat162float(by);
  }
}
template <int NP, int PITCH>
__device__ __forceinline__ void split_store4(__nv_bfloat16* base, int r, int c, float4 v) {
  uint32_t lo[NP], hi[NP];
  split2<NP>(v.x, v.y, lo);
  split2<NP>(v.z, v.w, hi);
#pragma unroll
  for (int pl = 0; pl < NP; ++pl)
    *reinterpret_cast<uint2*>(base + ((size_t)pl * 64 + r) * PITCH + c) = make_uint2(lo[pl], hi[pl]);
}

__device__ __forceinline__ void cp_async_16_zf(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}

// IN16 (NP == 1 only): q/k/v are the bf16 output of the QKV GEMM.  The tiles go global -> shared memory with cp.async (no register
// staging, no conversion), K/V double-buffered: half the bytes of the fp32 round trip, ~64 registers fewer.  The values are the same
// bf16 roundings the fp32 path produces on the fly, so the result is bit-identical.
template <int DH, int NP, bool IN16>
__global__ void __launch_bounds__(128, (NP == 1 && IN16 && DH <= 64) ? 4 : 1) attn_prefill_mma(const AttnArgs p) {
  static_assert(!IN16 || NP == 1, "bf16 inputs are the plain-bf16 mode");
  constexpr int PITCH = DH + 8;                       // bf16 elements per smem row: 16-byte aligned, conflict-free ldmatrix
  constexpr int KC = DH / 16;                         // k-chunks of the QK^T product
  constexpr int ONT = DH / 8;                         // 8-column n-tiles of the output
  constexpr int F4 = IN16 ? 1 : 64 * DH / 4 / 128;    // float4 per thread per 64-row tile (fp32 inputs)
  constexpr int NBUF = IN16 ? 2 : 1;
  constexpr int TILE = NP * 64 * PITCH;               // elements of one staged tile
  extern __shared__ __align__(16) uint8_t smem_mma[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_mma);      // [NP][64][PITCH]
  __nv_bfloat16* Ks = Qs + TILE;                                       // [NBUF][NP][64][PITCH]
  __nv_bfloat16* Vs = Ks + NBUF * TILE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 64;
  const float* qb = p.q + (size_t)b * p.Tq * p.ldq + h * DH;
  const float* kb = p.k + (size_t)b * p.Tk * p.ldk + h * DH;
  const float* vb = p.v + (size_t)b * p.Tk * p.ldv + h * DH;
  const __nv_bfloat16* qh = reinterpret_cast<const __nv_bfloat16*>(p.q) + (size_t)b * p.Tq * p.ldq + h * DH;     // IN16 views
  const __nv_bfloat16* kh = reinterpret_cast<const __nv_bfloat16*>(p.k) + (size_t)b * p.Tk * p.ldk + h * DH;
  const __nv_bfloat16* vh = reinterpret_cast<const __nv_bfloat16*>(p.v) + (size_t)b * p.Tk * p.ldv + h * DH;
  const int klen = p.lens ? min(__ldg(p.lens + b), p.Tk) : p.Tk;
  const uint8_t* km = p.key_mask ? p.key_mask + (size_t)b * p.Tk : nullptr;

  int nkt = (klen + 63) / 64;
  if (p.causal) nkt = min(nkt, (min(q0 + 64, p.Tq) - 1) / 64 + 1);

  // IN16: 64 rows x DH bf16 of a row-major matrix -> a staged tile, rows >= limit zero-filled
  auto fill16 = [&](const __nv_bfloat16* src, int ld, int row0, int limit, __nv_bfloat16* dst) {
    constexpr int CPR = DH / 8;                       // 16-byte chunks per row
    for (int i = tid; i < 64 * CPR; i += 128) {
      const int r = i / CPR, c = (i - r * CPR) * 8;
      const bool ok = row0 + r < limit;
      cp_async_16_zf((uint32_t)__cvta_generic_to_shared(dst + r * PITCH + c), ok ? src + (size_t)(row0 + r) * ld + c : src, ok ? 16u : 0u);
    }
  };

  float4 kreg[F4], vreg[F4];
  auto load_kv = [&](int kt) {
#pragma unroll
    for (int u = 0; u < F4; ++u) {
      const int i = tid + 128 * u, r = i / (DH / 4), c = (i % (DH / 4)) * 4;
      kreg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      vreg[u] = kreg[u];
      if (kt * 64 + r < klen) {
        kreg[u] = *reinterpret_cast<const float4*>(kb + (size_t)(kt * 64 + r) * p.ldk + c);
        vreg[u] = *reinterpret_cast<const float4*>(vb + (size_t)(kt * 64 + r) * p.ldv + c);
      }
    }
  };
  if constexpr (IN16) {
    fill16(qh, p.ldq, q0, p.Tq, Qs);
    if (nkt > 0) { fill16(kh, p.ldk, 0, klen, Ks); fill16(vh, p.ldv, 0, klen, Vs); }
    cp_async_commit();
  } else {
    if (nkt > 0) load_kv(0);
#pragma unroll
    for (int u = 0; u < F4; ++u) {
      const int i = tid + 128 * u, r = i / (DH / 4), c = (i % (DH / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q0 + r < p.Tq) v = *reinterpret_cast<const float4*>(qb + (size_t)(q0 + r) * p.ldq + c);
      split_store4<NP, PITCH>(Qs, r, c, v);
    }
  }

  float o[ONT][4];
#pragma unroll
  for (int nt = 0; nt < ONT; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int qrow0 = q0 + warp * 16 + (lane >> 2);     // this thread's rows: qrow0 and qrow0 + 8
  const uint32_t qs_u = (uint32_t)__cvta_generic_to_shared(Qs), ks_u = (uint32_t)__cvta_generic_to_shared(Ks),
                 vs_u = (uint32_t)__cvta_generic_to_shared(Vs);
  // per-lane ldmatrix row/column offsets (elements)
  const int a_row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;      // A (Q): m0 r0-7/c0-7, m1 r8-15, m2 c8-15, m3
  const int bk_row = (lane & 7) + (lane >> 4) * 8, bk_col = ((lane >> 3) & 1) * 8;               // B of S (K rows = keys): two n-tiles per x4
  const int bv_row = (lane & 7) + ((lane >> 3) & 1) * 8, bv_col = (lane >> 4) * 8;               // B of PV (V rows = keys, transposed): two n-tiles per x4

  const uint32_t ks_base = ks_u, vs_base = vs_u;
  for (int kt = 0; kt < nkt; ++kt) {
    const int k0 = kt * 64;
    if constexpr (IN16) {
      cp_async_wait<0>();                             // tile kt (and Q) have landed
      __syncthreads();                                // ... for every thread; buffer (kt + 1) & 1 is no longer being read
      if (kt + 1 < nkt) {
        fill16(kh, p.ldk, (kt + 1) * 64, klen, Ks + ((kt + 1) & 1) * TILE);
        fill16(vh, p.ldv, (kt + 1) * 64, klen, Vs + ((kt + 1) & 1) * TILE);
        cp_async_commit();
      }
    } else {
      __syncthreads();                                // previous tile fully consumed (first pass: nothing to wait for)
#pragma unroll
      for (int u = 0; u < F4; ++u) {
        const int i = tid + 128 * u, r = i / (DH / 4), c = (i % (DH / 4)) * 4;
        split_store4<NP, PITCH>(Ks, r, c, kreg[u]);
        split_store4<NP, PITCH>(Vs, r, c, vreg[u]);
      }
      __syncthreads();
      if (kt + 1 < nkt) load_kv(kt + 1);              // in flight during the MMAs below
    }
    const uint32_t ks_u = ks_base + (IN16 ? (uint32_t)((kt & 1) * TILE * 2) : 0u), vs_u = vs_base + (IN16 ? (uint32_t)((kt & 1) * TILE * 2) : 0u);

    // ---- S = Q K^T : 16 rows x 64 keys per warp
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
      uint32_t a[NP][4];
#pragma unroll
      for (int pa = 0; pa < NP; ++pa)
        ldsm_x4(qs_u + (uint32_t)(((pa * 64 + a_row) * PITCH + kc * 16 + a_col) * 2), a[pa][0], a[pa][1], a[pa][2], a[pa][3]);
#pragma unroll
      for (int pw = NP - 1; pw >= 0; --pw) {
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(ks_u + (uint32_t)(((pw * 64 + n2 * 16 + bk_row) * PITCH + kc * 16 + bk_col) * 2), b0, b1, b2, b3);
#pragma unroll
          for (int pa = NP - 1 - pw; pa >= 0; --pa) {
            mma_bf16(s[2 * n2], a[pa], b0, b1);
            mma_bf16(s[2 * n2 + 1], a[pa], b2, b3);
          }
        }
      }
    }

    // ---- scale, mask, online softmax (rows qrow0 -> s[.][0..1], qrow0 + 8 -> s[.][2..3])
    // Interior tiles (all 64 keys exist, none padded, entirely below this warp's diagonal) skip the per-element tests; the
    // key-padding mask of the tile is fetched once per warp (2 bytes per lane) and turned into a 64-bit ballot.
    unsigned long long keep = ~0ull;                    // bit j: key k0 + j is not padding
    if (km) {
      const int ka = k0 + lane, kb2 = k0 + 32 + lane;
      const unsigned lo = __ballot_sync(0xffffffffu, ka < klen ? km[ka] != 0 : true);
      const unsigned hi = __ballot_sync(0xffffffffu, kb2 < klen ? km[kb2] != 0 : true);
      keep = (unsigned long long)lo | ((unsigned long long)hi << 32);
    }
    const bool interior = (k0 + 64 <= klen) && keep == ~0ull && (!p.causal || k0 + 63 <= q0 + warp * 16);
    float tmax[2] = {-INFINITY, -INFINITY};
    if (interior) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = s[nt][e] * p.scale;
          s[nt][e] = v;
          tmax[e >> 1] = fmaxf(tmax[e >> 1], v);
        }
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int jj = nt * 8 + 2 * (lane & 3) + (e & 1), kj = k0 + jj, qi = qrow0 + (e >> 1) * 8;
          float v = s[nt][e] * p.scale;
          if (kj >= klen) v = -INFINITY;                                      // not a key at all
          else if (!((keep >> jj) & 1ull) || (p.causal && kj > qi)) v = -FLT_MAX;    // masked_fill(-finfo.max)
          s[nt][e] = v;
          tmax[e >> 1] = fmaxf(tmax[e >> 1], v);
        }
      }
    }
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
      const float m_new = fmaxf(m_run[r], tmax[r]);
      alpha[r] = (m_run[r] == -INFINITY) ? 0.f : expf(m_run[r] - m_new);
      m_run[r] = m_new;
    }
    float psum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // NP == 1 is the plain-bf16 mode (P is rounded to bf16 right below): the 2-instruction exp is exact enough there
        const float x = s[nt][e] - m_run[e >> 1];
        const float pv = (s[nt][e] == -INFINITY) ? 0.f : (NP == 1 ? __expf(x) : expf(x));
        s[nt][e] = pv;
        psum[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      psum[r] += __shfl_xor_sync(0xffffffffu, psum[r], 1);
      psum[r] += __shfl_xor_sync(0xffffffffu, psum[r], 2);
      l_run[r] = l_run[r] * alpha[r] + psum[r];
    }
#pragma unroll
    for (int nt = 0; nt < ONT; ++nt) {
      o[nt][0] *= alpha[0]; o[nt][1] *= alpha[0];
      o[nt][2] *= alpha[1]; o[nt][3] *= alpha[1];
    }

    // ---- O += P V : P (registers) is the A operand, 16 keys per k-chunk
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      uint32_t a[NP][4], w[NP];
      split2<NP>(s[2 * kc][0], s[2 * kc][1], w);
#pragma unroll
      for (int pl = 0; pl < NP; ++pl) a[pl][0] = w[pl];
      split2<NP>(s[2 * kc][2], s[2 * kc][3], w);
#pragma unroll
      for (int pl = 0; pl < NP; ++pl) a[pl][1] = w[pl];
      split2<NP>(s[2 * kc + 1][0], s[2 * kc + 1][1], w);
#pragma unroll
      for (int pl = 0; pl < NP; ++pl) a[pl][2] = w[pl];
      split2<NP>(s[2 * kc + 1][2], s[2 * kc + 1][3], w);
#pragma unroll
      for (int pl = 0; pl < NP; ++pl) a[pl][3] = w[pl];
#pragma unroll
      for (int pw = NP - 1; pw >= 0; --pw) {
#pragma unroll
        for (int n2 = 0; n2 < ONT / 2; ++n2) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(vs_u + (uint32_t)(((pw * 64 + kc * 16 + bv_row) * PITCH + n2 * 16 + bv_col) * 2), b0, b1, b2, b3);
#pragma unroll
          for (int pa = NP - 1 - pw; pa >= 0; --pa) {
            mma_bf16(o[2 * n2], a[pa], b0, b1);
            mma_bf16(o[2 * n2 + 1], a[pa], b2, b3);
          }
        }
      }
    }
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qi = qrow0 + r * 8;
    if (qi >= p.Tq) continue;
    const float inv = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
    const int cbase = h * DH + 2 * (lane & 3);
    if (p.out) {
      float* orow = p.out + ((size_t)b * p.Tq + qi) * p.ldo + cbase;
#pragma unroll
      for (int nt = 0; nt < ONT; ++nt)
        *reinterpret_cast<float2*>(orow + nt * 8) = make_float2(o[nt][2 * r] * inv, o[nt][2 * r + 1] * inv);
    }
    if (p.out_p) {
      __nv_bfloat16* prow = p.out_p + ((size_t)b * p.Tq + qi) * p.planes * p.kp + cbase;
#pragma unroll
      for (int nt = 0; nt < ONT; ++nt) {
        float x = o[nt][2 * r] * inv, y = o[nt][2 * r + 1] * inv;
        for (int pl = 0; pl < p.planes; ++pl) {
          const __nv_bfloat16 bx = __float2bfloat16_rn(x), by = __float2bfloat16_rn(y);
          *reinterpret_cast<uint32_t*>(prow + (size_t)pl * p.kp + nt * 8) =
              (uint32_t)__bfloat16_as_ushort(bx) | ((uint32_t)__bfloat16_as_ushort(by) << 16);
          x -= __bfloat162float(bx);
          y -= __bfloat162float(by);
        }
      }
    }
  }
}

// ---- decode: one query row per (b,h) ----------------------------------------------------------------------------
// A group of LPK lanes owns whole keys: every lane moves one 128-bit vector per key (4 fp32 or 8 bf16), so a (b,h) head row
// -- 256 B fp32 / 128 B bf16, contiguous in the head-major cache -- is one fully used request and a warp reads 512 contiguous
// bytes per instruction (pass 2 of the kernel below; pass 1 reads one whole row per thread from shared memory).
// volatile: the U loads of a batch stay where they are written, back to back, ahead of the first use
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// Lane-group variant (kept for A/B measurements, DIM_ATTN_IMPL=lanes): LPK lanes per key, U keys in flight per lane.
template <bool BF16, int NT, int U>
__global__ void __launch_bounds__(NT) attn_decode_lanes(const DecodeAttnArgs p) {
  typedef typename KvIo<BF16>::T KT;
  constexpr int DH = 64, EPL = KvIo<BF16>::EPL, LPK = DH / EPL, NG = NT / LPK, NW = NT / 32;
  extern __shared__ __align__(16) float sc[];          // [nkeys_max] scores, then [NG][64] partial outputs
  __shared__ float red[8];
  pdl_prologue();
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid / LPK, lk = tid % LPK;           // key group of this lane, position inside the head row
  const int pos = p.step ? *p.step : 0;                                   // index of the token being decoded
  const int nkeys = p.append ? pos + 1 : p.Tk;
  const int bkv = b / p.kv_group;                      // cache row (samples of one clip share the cross-attention K/V)
  KT* kbase = static_cast<KT*>(p.k) + (size_t)bkv * p.kv_batch_stride + (size_t)h * p.kv_head_stride + lk * EPL;
  KT* vbase = static_cast<KT*>(p.v) + (size_t)bkv * p.kv_batch_stride + (size_t)h * p.kv_head_stride + lk * EPL;

  float q[EPL];
#pragma unroll
  for (int i = 0; i < EPL; i += 4) {
    const float4 x = *reinterpret_cast<const float4*>(p.q + (size_t)b * p.ldq + h * DH + lk * EPL + i);
    q[i] = x.x; q[i + 1] = x.y; q[i + 2] = x.z; q[i + 3] = x.w;
  }
  if (p.append) {                                                        // cache[pos] <- this step's k, v
    if (grp < 2) {
      const float* src = (grp == 0 ? p.k_new : p.v_new) + (size_t)b * p.ld_new + h * DH + lk * EPL;
      float v[EPL];
#pragma unroll
      for (int i = 0; i < EPL; i += 4) {
        const float4 x = *reinterpret_cast<const float4*>(src + i);
        v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
      }
      KvIo<BF16>::st((grp == 0 ? kbase : vbase) + (size_t)pos * p.kv_tok_stride, v);
    }
    __syncthreads();
  }
  const uint8_t* km = p.key_mask ? p.key_mask + (size_t)bkv * p.Tk : nullptr;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

  // scores: U keys in flight per group
  for (int base = 0; base < nkeys; base += NG * U) {   // warp-uniform trip count: the shuffles below need all lanes
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = base + grp + NG * u;
      raw[u] = zero4;
      if (j < nkeys) raw[u] = ld_stream16(kbase + (size_t)j * p.kv_tok_stride);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = base + grp + NG * u;
      float d = KvIo<BF16>::dot(raw[u], q);
#pragma unroll
      for (int off = LPK / 2; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
      if (lk == 0 && j < nkeys) sc[j] = d * p.scale;
    }
  }
  __syncthreads();
  // key-padding mask (masked_fill(-finfo.max), applied here with coalesced byte loads instead of one dependent load per key
  // inside the streaming loop) and softmax statistics
  float mx = -INFINITY;
  for (int j = tid; j < nkeys; j += NT) {
    float v = sc[j];
    if (km && !km[j]) { v = -FLT_MAX; sc[j] = v; }
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < nkeys; j += NT) {
    float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  float tot = red[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) tot += red[w];
  const float inv = 1.f / tot;

  // out = P V
  float acc[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) acc[i] = 0.f;
  for (int base = 0; base < nkeys; base += NG * U) {
    uint4 raw[U];
    float pj[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = base + grp + NG * u;
      const bool ok = j < nkeys;
      raw[u] = zero4;
      if (ok) raw[u] = ld_stream16(vbase + (size_t)j * p.kv_tok_stride);
      pj[u] = ok ? sc[j] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) KvIo<BF16>::axpy(raw[u], pj[u], acc);
  }
  float* part = sc + p.sc_floats;                       // [NG][64]
#pragma unroll
  for (int i = 0; i < EPL; ++i) part[grp * DH + lk * EPL + i] = acc[i];
  __syncthreads();
  if (tid < DH) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < NG; ++w) r += part[w * DH + tid];
    if (p.out) p.out[(size_t)b * p.ldo + h * DH + tid] = r * inv;
    if (p.out_p) store_planes1(p.out_p + (size_t)b * p.planes * p.kp + h * DH + tid, r * inv, p.planes, p.kp);
  }
}


// One CTA per (clip, head); the head's K block and V block are contiguous streams (head-major cache).  Keys are staged in
// CHUNK-row tiles through a 2-stage cp.async ring in shared memory (rows padded by 16 B: conflict-free for both access
// patterns below), so the bytes in flight per SM do not depend on registers or on instruction scheduling:
//   pass 1 (scores): one THREAD per key -- the whole 64-element dot product is local (no shuffles), q lives in registers;
//   softmax over the scores in shared memory (key-padding mask applied here), the first V tiles already in flight;
//   pass 2 (P.V)   : a group of LPK lanes per key, each lane owns one 16-byte slice of the head row and accumulates it over
//                    the keys of its group; groups are reduced through shared memory at the end.
// ~10 warp instructions per key (the lane-group kernel above spends ~30 and is issue-bound at ~4 TB/s on bf16 rows).
// Measured alone (scripts/attn_roofline.py, 256 clips): bf16 rows 5.2 TB/s at 300 keys, 6.2 TB/s at 1024 keys; fp32 rows
// 6.2-6.6 TB/s.  A variant with separate K and V rings (everything requested up front, 79 KB per CTA) was SLOWER (4.1 TB/s):
// resident CTAs per SM matter more than the second exposed round trip.
template <bool BF16, int NT, int CHUNK>
__global__ void __launch_bounds__(NT) attn_decode_kernel(const DecodeAttnArgs p) {
  typedef typename KvIo<BF16>::T KT;
  constexpr int DH = 64, EPL = KvIo<BF16>::EPL, LPK = DH / EPL, NG = NT / LPK, NW = NT / 32;
  constexpr int ROWB = DH * (int)sizeof(KT), PITCH = ROWB + 16, STAGE = CHUNK * PITCH, KPG = CHUNK / NG;
  static_assert(CHUNK <= NT && CHUNK % NG == 0, "chunk shape");
  extern __shared__ __align__(16) uint8_t dsm[];       // [2][STAGE] ring | scores [sc_floats] | partial outputs [NG][64]
  __shared__ float red[8];
  float* sc = reinterpret_cast<float*>(dsm + 2 * STAGE);
  float* part = sc + p.sc_floats;
  pdl_prologue();
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid / LPK, lk = tid % LPK;           // pass-2 role: key group, 16-byte slice inside the head row
  const int pos = p.step ? *p.step : 0;                                   // index of the token being decoded
  const int nkeys = p.append ? pos + 1 : p.Tk;
  const int bkv = b / p.kv_group;                      // cache row (samples of one clip share the cross-attention K/V)
  KT* khead = static_cast<KT*>(p.k) + (size_t)bkv * p.kv_batch_stride + (size_t)h * p.kv_head_stride;   // rows 64 elements apart
  KT* vhead = static_cast<KT*>(p.v) + (size_t)bkv * p.kv_batch_stride + (size_t)h * p.kv_head_stride;
  const uint32_t ring_u = (uint32_t)__cvta_generic_to_shared(dsm);

  if (p.append) {                                                        // cache[pos] <- this step's k, v
    if (grp < 2) {
      const float* src = (grp == 0 ? p.k_new : p.v_new) + (size_t)b * p.ld_new + h * DH + lk * EPL;
      float v[EPL];
#pragma unroll
      for (int i = 0; i < EPL; i += 4) {
        const float4 x = *reinterpret_cast<const float4*>(src + i);
        v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
      }
      KvIo<BF16>::st((grp == 0 ? khead : vhead) + (size_t)pos * DH + lk * EPL, v);
      __threadfence();                                                   // the row is read back below through cp.async (L2)
    }
    __syncthreads();
  }
  const int nch = (nkeys + CHUNK - 1) / CHUNK;
  uint64_t kvpol = 0;
  if (p.kv_evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(kvpol));
  auto issue = [&](const KT* head, int c, int stage) {                   // rows [c*CHUNK, ..) of a head block -> ring[stage]
    const int row0 = c * CHUNK, pieces = min(CHUNK, nkeys - row0) * LPK;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(head + (size_t)row0 * DH);
    const uint32_t dst = ring_u + stage * STAGE;
    if (p.kv_evict_first) {                                              // read-once stream: do not displace L2-resident weights
      for (int q = tid; q < pieces; q += NT) cp_async_16_hint(dst + (q / LPK) * PITCH + (q % LPK) * 16, src + (size_t)q * 16, kvpol);
    } else {
      for (int q = tid; q < pieces; q += NT) cp_async_16(dst + (q / LPK) * PITCH + (q % LPK) * 16, src + (size_t)q * 16);
    }
    cp_async_commit();
  };
  if (nch > 0) issue(khead, 0, 0);
  if (nch > 1) issue(khead, 1, 1);

  float q[DH];
#pragma unroll
  for (int i = 0; i < DH; i += 4) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(p.q + (size_t)b * p.ldq + h * DH + i));
    q[i] = x.x; q[i + 1] = x.y; q[i + 2] = x.z; q[i + 3] = x.w;
  }

  // ---- pass 1: scores
  for (int c = 0; c < nch; ++c) {
    if (c + 1 < nch) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    const int j = c * CHUNK + tid;
    if (tid < CHUNK && j < nkeys) {
      const uint4* row = reinterpret_cast<const uint4*>(dsm + (c & 1) * STAGE + tid * PITCH);
      float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < LPK; ++i) d[i & 3] += KvIo<BF16>::dot(row[i], q + i * EPL);
      sc[j] = ((d[0] + d[1]) + (d[2] + d[3])) * p.scale;
    }
    __syncthreads();
    if (c + 2 < nch) issue(khead, c + 2, c & 1);
  }
  // first V tiles in flight while the softmax statistics are computed
  if (nch > 0) issue(vhead, 0, 0);
  if (nch > 1) issue(vhead, 1, 1);

  // key-padding mask (masked_fill(-finfo.max)) and softmax statistics
  const uint8_t* km = p.key_mask ? p.key_mask + (size_t)bkv * p.Tk : nullptr;
  float mx = -INFINITY;
  for (int j = tid; j < nkeys; j += NT) {
    float v = sc[j];
    if (km && !km[j]) { v = -FLT_MAX; sc[j] = v; }
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < nkeys; j += NT) {
    float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  float tot = red[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) tot += red[w];
  const float inv = 1.f / tot;

  // ---- pass 2: out = P V
  float acc[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) acc[i] = 0.f;
  for (int c = 0; c < nch; ++c) {
    if (c + 1 < nch) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    const uint8_t* tile = dsm + (c & 1) * STAGE + lk * 16;
#pragma unroll
    for (int i = 0; i < KPG; ++i) {
      const int jl = grp + NG * i, j = c * CHUNK + jl;
      if (j < nkeys) KvIo<BF16>::axpy(*reinterpret_cast<const uint4*>(tile + jl * PITCH), sc[j], acc);
    }
    __syncthreads();
    if (c + 2 < nch) issue(vhead, c + 2, c & 1);
  }
#pragma unroll
  for (int i = 0; i < EPL; ++i) part[grp * DH + lk * EPL + i] = acc[i];
  __syncthreads();
  if (tid < DH) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < NG; ++w) r += part[w * DH + tid];
    if (p.out) p.out[(size_t)b * p.ldo + h * DH + tid] = r * inv;
    if (p.out_p) store_planes1(p.out_p + (size_t)b * p.planes * p.kp + h * DH + tid, r * inv, p.planes, p.kp);
  }
}

// 16-byte vectors: read token-major rows fully coalesced, write 64-element head rows (128 B bf16 / 256 B fp32 chunks).
template <int VE>
__global__ void __launch_bounds__(256) kv_head_major_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int B, int T,
                                                            int H) {
  const int vpr = 2 * H * 64 / VE;                      // vectors per source row
  const size_t total = (size_t)B * T * vpr, plane = (size_t)B * H * T * 64 / VE;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / vpr;
    const int cv = (int)(i - row * vpr), col = cv * VE;
    const int kv = col / (H * 64), hc = col - kv * H * 64, h = hc >> 6, d = hc & 63;
    const int b = (int)(row / T), t = (int)(row - (size_t)b * T);
    dst[kv * plane + ((((size_t)b * H + h) * T + t) * 64 + d) / VE] = __ldcs(src + i);
  }
}

}  // namespace

int launch_kv_head_major(const void* src, void* dst, int B, int T, int H, int bf16, cudaStream_t s) {
  DIM_REQUIRE(src && dst && B > 0 && T > 0 && H > 0, "kv_head_major: bad argument");
  const size_t total = (size_t)B * T * 2 * H * 64 / (bf16 ? 8 : 4);
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 16);
  ProfScope ps(CAT_MISC, s, (double)total * 32.0, 0);
  if (bf16) kv_head_major_kernel<8><<<blocks, 256, 0, s>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), B, T, H);
  else kv_head_major_kernel<4><<<blocks, 256, 0, s>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), B, T, H);
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_attention_prefill(const AttnArgs& a, cudaStream_t s) {
  DIM_REQUIRE(a.Dh == 48 || a.Dh == 64 || a.Dh == 96, "attention: head dim must be 48, 64 or 96");
  DIM_REQUIRE(a.B > 0 && a.H > 0 && a.Tq > 0 && a.Tk > 0, "attention: empty");
  DIM_REQUIRE(a.ldq % 4 == 0 && a.ldk % 4 == 0 && a.ldv % 4 == 0, "attention: leading dims must be multiples of 4");
  DIM_REQUIRE(!a.in_bf16 || (a.out_p != nullptr && a.planes == 1), "attention: bf16 inputs are served by the tensor-core kernel only");
  dim3 grid(cdiv(a.Tq, TQ), a.H, a.B);
  // algorithmic: read q,k,v once, write out once; QK^T and PV (causal: half)
  ProfScope ps(CAT_ATTN_PREFILL, s, 4.0 * a.B * a.H * a.Dh * (2.0 * a.Tq + 2.0 * a.Tk),
               4.0 * a.B * a.H * (double)a.Tq * a.Tk * a.Dh * (a.causal ? 0.5 : 1.0));
  // tensor-core path: whenever the caller runs the plane-fused tensor-core data flow (bf16-plane output requested) with
  // 1 plane (bf16 mode) or 3 planes (fp32-grade); DIM_ATTN_PREFILL=ffma forces the FFMA kernel (A/B tuning hook)
  static const bool force_ffma = getenv("DIM_ATTN_PREFILL") != nullptr && std::string(getenv("DIM_ATTN_PREFILL")) == "ffma";
  if (a.out_p != nullptr && (a.planes == 1 || a.planes == 3) && !force_ffma) {
    typedef void (*Kern)(const AttnArgs);
    const int np = a.planes;
    // 48: VQAutoEncoder (384 / 8 heads), 64: x-transformers, 96: VQSpeakerAutoEncoder (768 / 8 heads, stage1_BIWI.py:140)
    const bool in16 = a.in_bf16 != 0;
    DIM_REQUIRE(!in16 || np == 1, "attention: bf16 inputs need the plain-bf16 mode (planes == 1)");
    DIM_REQUIRE(!in16 || (a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0), "attention: bf16 rows must be 16-byte aligned");
    Kern kern = a.Dh == 48 ? (np == 1 ? (in16 ? (Kern)attn_prefill_mma<48, 1, true> : (Kern)attn_prefill_mma<48, 1, false>) : (Kern)attn_prefill_mma<48, 3, false>)
              : a.Dh == 64 ? (np == 1 ? (in16 ? (Kern)attn_prefill_mma<64, 1, true> : (Kern)attn_prefill_mma<64, 1, false>) : (Kern)attn_prefill_mma<64, 3, false>)
                           : (np == 1 ? (in16 ? (Kern)attn_prefill_mma<96, 1, true> : (Kern)attn_prefill_mma<96, 1, false>) : (Kern)attn_prefill_mma<96, 3, false>);
    const size_t smem = (size_t)(in16 ? 5 : 3) * np * 64 * (a.Dh + 8) * sizeof(__nv_bfloat16);
    static PerDeviceOnce configured[9];
    const int slot = (a.Dh == 48 ? 0 : a.Dh == 64 ? 3 : 6) + (np == 1 ? (in16 ? 2 : 0) : 1);
    if (configured[slot].first()) DIM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(cdiv(a.Tq, 64), a.H, a.B), 128, smem, s>>>(a);
    DIM_LAUNCHED();
    return DIM_OK;
  }
  if (a.Dh == 48) {
    constexpr size_t smem = (TQ * 52 + TKV * 52 + TKV * 48 + TQ * (TKV + 4)) * sizeof(float);
    static PerDeviceOnce once;
    if (once.first()) DIM_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_f32<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_prefill_f32<48><<<grid, 256, smem, s>>>(a);
  } else if (a.Dh == 64) {
    constexpr size_t smem = (TQ * 68 + TKV * 68 + TKV * 64 + TQ * (TKV + 4)) * sizeof(float);
    static PerDeviceOnce once;
    if (once.first()) DIM_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_f32<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_prefill_f32<64><<<grid, 256, smem, s>>>(a);
  } else {
    constexpr size_t smem = (TQ * 100 + TKV * 100 + TKV * 96 + TQ * (TKV + 4)) * sizeof(float);
    static PerDeviceOnce once;
    if (once.first()) DIM_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_f32<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_prefill_f32<96><<<grid, 256, smem, s>>>(a);
  }
  DIM_LAUNCHED();
  return DIM_OK;
}

int g_attn_impl = -1;      // -1: from DIM_ATTN_IMPL (default ring); 0: cp.async ring kernel; 1: lane-group kernel

int launch_attention_decode(DecodeAttnArgs a, int max_keys, cudaStream_t s) {
  DIM_REQUIRE(a.B > 0 && a.H > 0 && a.kv_group >= 1, "decode attention: empty");
  // K/V rows are read once per step: evict_first keeps them from displacing the L2-resident decode weights (see launch_gemm_tc)
  static const bool evict_first = !(getenv("DIM_L2_WEIGHT_KEEP") != nullptr && atof(getenv("DIM_L2_WEIGHT_KEEP")) <= 0.0);
  a.kv_evict_first = evict_first ? 1 : 0;
  DIM_REQUIRE(a.kv_group == 1 || a.append == 0, "decode attention: only read-only caches can be shared between rows");
  DIM_REQUIRE(a.kv_tok_stride == 64, "decode attention: the K/V caches must be head-major ([B,H,tokens,64])");
  const bool bf = a.kv_bf16 != 0;
  if (g_attn_impl < 0) {
    const char* e = getenv("DIM_ATTN_IMPL");
    g_attn_impl = (e && std::string(e) == "lanes") ? 1 : (e && atoi(e) > 0 ? atoi(e) : 0);
  }
  // impl 0 (default): bf16 rows: ring kernel, 64 threads, 64-key tiles (24 KB / CTA, ~10 CTAs per SM: fewest waves, 5.4 TB/s
  // at 300 keys); fp32 rows: lane-group kernel (6.2 TB/s, co-resides with the concurrent decode groups of the fp32-grade
  // mode).  1: lane-group kernel; 2, 3, 4: ring-kernel shapes kept
  // for A/B sweeps (scripts/attn_roofline.py; profiles/r01j_attn_roofline.jsonl)
  typedef void (*Kern)(const DecodeAttnArgs);
  Kern kern;
  int nt = 128;
  size_t ring = 0;
  switch (g_attn_impl) {
    case 1: kern = bf ? (Kern)attn_decode_lanes<true, 128, 8> : (Kern)attn_decode_lanes<false, 128, 8>; break;
    case 2: nt = 64; ring = bf ? 2 * 64 * 144 : 2 * 32 * 272;
            kern = bf ? (Kern)attn_decode_kernel<true, 64, 64> : (Kern)attn_decode_kernel<false, 64, 32>; break;
    case 3: ring = bf ? 2 * 64 * 144 : 2 * 32 * 272;
            kern = bf ? (Kern)attn_decode_kernel<true, 128, 64> : (Kern)attn_decode_kernel<false, 128, 32>; break;
    case 4: ring = bf ? 2 * 128 * 144 : 2 * 64 * 272;
            kern = bf ? (Kern)attn_decode_kernel<true, 128, 128> : (Kern)attn_decode_kernel<false, 128, 64>; break;
    default:
      if (bf) { nt = 64; ring = 2 * 64 * 144; kern = (Kern)attn_decode_kernel<true, 64, 64>; }
      else kern = (Kern)attn_decode_lanes<false, 128, 8>;     // fp32 rows: equal bandwidth (6.2 TB/s), 5 KB of smem per CTA
      break;
  }
  a.sc_floats = (max_keys + 3) / 4 * 4;
  const size_t smem = ring + (size_t)(a.sc_floats + 16 * 64) * sizeof(float);
  DIM_REQUIRE(smem <= 200 * 1024, "decode attention: too many keys for one CTA");
  static size_t configured[64][8];                                          // per device: largest size set so far (0 = the 48 KB default)
  const int slot = (bf ? 1 : 0) + 2 * (g_attn_impl == 4 ? 0 : (g_attn_impl & 3));
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (smem > 48 * 1024 && smem > configured[dev][slot]) {
    DIM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev][slot] = smem;
  }
  {
    const double keys = a.append ? (double)(a.prof_pos + 1) : (double)a.Tk;   // K and V rows actually read
    const double esz = bf ? 2.0 : 4.0;
    ProfScope ps(CAT_ATTN_DECODE, s, a.B * (double)a.H * 64.0 * (2.0 * keys * esz + 8.0), 4.0 * a.B * a.H * 64.0 * keys);
    DIM_CHECK_CUDA(launch_k(kern, dim3(a.B * a.H), dim3(nt), smem, s, a));
  }
  DIM_LAUNCHED();
  return DIM_OK;
}

}  // namespace dimb

using namespace dimb;

// tuning hook (not part of the stable ABI): one cross-attention-shaped decode launch on caller buffers.
// k, v: head-major [B,H,Tk,64] (fp32 or bf16); q, out: [B, H*64] fp32.  impl: 0 ring, 1 lanes.
extern "C" int dim_debug_attn_decode(int impl, void* k, void* v, const float* q, float* out, int B, int H, int Tk, int bf16,
                                     void* stream) {
  if (int e = ensure_device()) return e;
  if (impl >= 0) dimb::g_attn_impl = impl;          // impl < 0: whatever the library would use (default / DIM_ATTN_IMPL)
  DecodeAttnArgs a;
  a.q = q; a.ldq = H * 64; a.k = k; a.v = v; a.kv_bf16 = bf16;
  a.kv_batch_stride = (size_t)H * Tk * 64; a.kv_head_stride = (size_t)Tk * 64; a.kv_tok_stride = 64;
  a.append = 0; a.out = out; a.ldo = H * 64; a.B = B; a.H = H; a.Tk = Tk; a.scale = 0.125f;
  return launch_attention_decode(a, Tk, as_stream(stream));
}

extern "C" int dim_attention_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out,
                                 int ldo, const uint8_t* key_mask, const int32_t* lens, int B, int H, int Tq, int Tk, int Dh,
                                 float scale, int causal, void* stream) {
  if (int e = ensure_device()) return e;
  AttnArgs a;
  a.q = q; a.ldq = ldq; a.k = k; a.ldk = ldk; a.v = v; a.ldv = ldv; a.out = out; a.ldo = ldo;
  a.key_mask = key_mask; a.lens = lens; a.B = B; a.H = H; a.Tq = Tq; a.Tk = Tk; a.Dh = Dh; a.scale = scale;
  a.causal = causal;
  return launch_attention_prefill(a, as_stream(stream));
}
