// Row-wise kernels: LayerNorm, InstanceNorm over time, embedding/context assembly, logits sampling.  All HBM-bound:
// coalesced 128-bit accesses, warp-shuffle reductions, grids sized from the row count.
#include "rowops.cuh"

namespace dimb {

namespace {

// One warp per row.  dim % 4 == 0, dim <= 128 * MAXV (MAXV float4 per lane kept in registers: one HBM read per element).
template <int MAXV>
__global__ void __launch_bounds__(256) layer_norm_kernel(const float* __restrict__ x, const float* __restrict__ gain,
                                                         const float* __restrict__ bias, float* __restrict__ y,
                                                         __nv_bfloat16* __restrict__ yb, int rows, int dim, float eps,
                                                         __nv_bfloat16* __restrict__ yp, int planes, int kp) {
  pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int n4 = dim >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)warp * dim);
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = lane + 32 * i;
    if (c < n4) {
      v[i] = xr[c];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)dim;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = lane + 32 * i;
    if (c < n4) {
      float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)dim + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = lane + 32 * i;
    if (c < n4) {
      float4 g = __ldg(reinterpret_cast<const float4*>(gain) + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x; o.y = (v[i].y - mean) * rstd * g.y;
      o.z = (v[i].z - mean) * rstd * g.z; o.w = (v[i].w - mean) * rstd * g.w;
      if (bias) {
        float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c);
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
      }
      if (y) reinterpret_cast<float4*>(y + (size_t)warp * dim)[c] = o;
      if (yb) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        reinterpret_cast<uint2*>(yb + (size_t)warp * dim)[c] = pk;
      }
      if (yp) store_planes4(yp + (size_t)warp * planes * kp + c * 4, o, planes, kp);
    }
  }
}

// InstanceNorm1d over time on frames (B,T,C): block = 32 channels x 8 time lanes of one sample.
__global__ void __launch_bounds__(256) instance_norm_kernel(float* __restrict__ x, const int32_t* __restrict__ lens, int T,
                                                            int C, float eps) {
  __shared__ float red[8][33];
  __shared__ float stat[2][32];
  const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), tl = threadIdx.x >> 5;
  const int L = lens ? min(lens[b], T) : T;
  float* xb = x + (size_t)b * T * C;
  const bool ok = c < C;
  float s = 0.f;
  if (ok)
    for (int t = tl; t < L; t += 8) s += xb[(size_t)t * C + c];
  red[tl][threadIdx.x & 31] = s;
  __syncthreads();
  if (tl == 0) {
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) m += red[i][threadIdx.x];
    stat[0][threadIdx.x] = m / (float)L;
  }
  __syncthreads();
  const float mean = stat[0][threadIdx.x & 31];
  float q = 0.f;
  if (ok)
    for (int t = tl; t < L; t += 8) {
      float d = xb[(size_t)t * C + c] - mean;
      q += d * d;
    }
  red[tl][threadIdx.x & 31] = q;
  __syncthreads();
  if (tl == 0) {
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) m += red[i][threadIdx.x];
    stat[1][threadIdx.x] = rsqrtf(m / (float)L + eps);     // biased variance, like F.instance_norm
  }
  __syncthreads();
  const float rstd = stat[1][threadIdx.x & 31];
  if (ok)
    for (int t = tl; t < L; t += 8) {
      size_t o = (size_t)t * C + c;
      xb[o] = (xb[o] - mean) * rstd;
    }
}

// ctx[b,t,:] = cat(x_s[b,t,:] + pe_dec[:], audio[b,t,:])      seq2seq_pretrain.py:445-446
__global__ void build_context_kernel(const float* __restrict__ xs, const float* __restrict__ pe_dec,
                                     const float* __restrict__ audio, float* __restrict__ ctx, __nv_bfloat16* ctxb,
                                     size_t rows, int d1, int d2) {
  const int d4 = (d1 + d2) >> 2, a4 = d1 >> 2;
  size_t total = rows * d4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i / d4;
    int c = (int)(i - r * d4);
    float4 v;
    if (c < a4) {
      v = reinterpret_cast<const float4*>(xs + r * d1)[c];
      float4 e = __ldg(reinterpret_cast<const float4*>(pe_dec) + c);
      v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
    } else {
      v = reinterpret_cast<const float4*>(audio + r * d2)[c - a4];
    }
    if (ctx) reinterpret_cast<float4*>(ctx)[i] = v;
    if (ctxb) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(ctxb)[i] = pk;
    }
  }
}

// x[b,:] = token_emb[tok[b],:]   (TransformerWrapper token embedding; no positional term for SLMFT)
// pos != nullptr: + pos[position,:] * pos_scale (x-transformers AbsolutePositionalEmbedding; position = the step index)
__global__ void embed_tokens_kernel(const int64_t* __restrict__ tok, int tok_stride, const int* __restrict__ step,
                                    const float* __restrict__ emb, float* __restrict__ x, int B, int D, int V,
                                    const float* __restrict__ pos, float pos_scale) {
  pdl_prologue();
  const int b = blockIdx.x;
  const int st = step ? *step : 0;
  const int64_t* tp = tok + (size_t)b * tok_stride + st;
  int64_t t = *tp;
  t = t < 0 ? 0 : (t >= V ? V - 1 : t);
  const float4* src = reinterpret_cast<const float4*>(emb + (size_t)t * D);
  const float4* ps = pos ? reinterpret_cast<const float4*>(pos + (size_t)st * D) : nullptr;
  float4* dst = reinterpret_cast<float4*>(x + (size_t)b * D);
  for (int i = threadIdx.x; i < (D >> 2); i += blockDim.x) {
    float4 v = __ldg(src + i);
    if (ps) {
      const float4 q = __ldg(ps + i);
      v.x = fmaf(q.x, pos_scale, v.x); v.y = fmaf(q.y, pos_scale, v.y); v.z = fmaf(q.z, pos_scale, v.z); v.w = fmaf(q.w, pos_scale, v.w);
    }
    dst[i] = v;
  }
}

// One block (256 threads) per row of logits [V <= 1024].  Greedy: first maximal index (torch.argmax on CPU returns the
// first occurrence).  Sampling: keep the top_k largest logits (ties broken toward the lower index, like a stable
// descending sort), softmax(l / temperature) over them in fp32, inverse-CDF draw in index order with the supplied uniform.
// Returns the chosen token to EVERY thread of the block (and stores it at out[b, out_offset + st]).
__device__ __forceinline__ int sample_row(const float* __restrict__ logits, int V, float temperature, int top_k,
                                          const float* __restrict__ uniforms, int u_stride, int st, int64_t* __restrict__ out,
                                          int out_stride, int out_offset, float* __restrict__ logits_out, int lo_stride) {
  __shared__ __align__(16) float sl[1024];
  __shared__ __align__(16) float sp[1024];
  __shared__ float redf[8];
  __shared__ int redi[8];
  __shared__ int tok_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* lr = logits + (size_t)b * V;
  for (int i = tid; i < V; i += 256) {
    float v = lr[i];
    sl[i] = v;
    if (logits_out) logits_out[(size_t)b * lo_stride + (size_t)st * V + i] = v;
  }
  __syncthreads();
  int64_t* dst = out + (size_t)b * out_stride + out_offset + st;
  if (temperature == 0.f) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < V; i += 256) {
      float v = sl[i];
      if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { redf[warp] = best; redi[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (redf[w] > best || (redf[w] == best && redi[w] < bi)) { best = redf[w]; bi = redi[w]; }
      tok_s = bi == 0x7fffffff ? 0 : bi;
      *dst = tok_s;
    }
    __syncthreads();
    return tok_s;
  }
  // rank of element i = #{j : l_j > l_i or (l_j == l_i and j < i)} ; kept iff rank < top_k.
  // Each thread ranks its elements against the row held in shared memory, 4 comparisons per 128-bit read.
  float mx = -INFINITY;
  const float4* sl4 = reinterpret_cast<const float4*>(sl);
  for (int i = tid; i < V; i += 256) {
    const float v = sl[i];
    int rank = 0;
    for (int j4 = 0; j4 < (V >> 2); ++j4) {
      const float4 w = sl4[j4];
      const int j = j4 << 2;
      rank += (w.x > v) || (w.x == v && j < i);
      rank += (w.y > v) || (w.y == v && j + 1 < i);
      rank += (w.z > v) || (w.z == v && j + 2 < i);
      rank += (w.w > v) || (w.w == v && j + 3 < i);
    }
    for (int j = V & ~3; j < V; ++j) {
      const float w = sl[j];
      rank += (w > v) || (w == v && j < i);
    }
    const bool keep = rank < top_k;
    sp[i] = keep ? v / temperature : -INFINITY;
    if (keep) mx = fmaxf(mx, v / temperature);
  }
  mx = warp_max(mx);
  if (lane == 0) redf[warp] = mx;
  __syncthreads();
  mx = redf[0];
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, redf[w]);
  __syncthreads();
  for (int i = tid; i < V; i += 256) sp[i] = sp[i] == -INFINITY ? 0.f : expf(sp[i] - mx);
  __syncthreads();
  // Inverse CDF in index order (fp64, deterministic): thread t owns entries [t*per, (t+1)*per); warp 0 turns the 256 chunk
  // sums into exclusive prefixes; each thread then tests its own entries against target = u * total.
  __shared__ double chunk[256];
  __shared__ double total_s;
  __shared__ int pick_s, last_s;
  const int per = (V + 255) / 256;                                 // <= 4 for V <= 1024
  double mine = 0.0;
  for (int k = 0; k < per; ++k) {
    const int i = tid * per + k;
    if (i < V) mine += (double)sp[i];
  }
  chunk[tid] = mine;
  if (tid == 0) { pick_s = 0x7fffffff; last_s = 0; }
  __syncthreads();
  if (warp == 0) {                                                   // lane l scans chunks 8l .. 8l+7, then a warp scan
    double loc[8], run = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { loc[k] = run; run += chunk[lane * 8 + k]; }
    double incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    const double excl = incl - run;
#pragma unroll
    for (int k = 0; k < 8; ++k) chunk[lane * 8 + k] = excl + loc[k];
    if (lane == 31) total_s = incl;
  }
  __syncthreads();
  const double target = (double)uniforms[(size_t)b * u_stride + st] * total_s;
  double c = chunk[tid];
  for (int k = 0; k < per; ++k) {
    const int i = tid * per + k;
    if (i < V) {
      const float pv = sp[i];
      if (pv > 0.f) {
        atomicMax(&last_s, i);
        c += (double)pv;
        if (c > target) atomicMin(&pick_s, i);
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    tok_s = pick_s == 0x7fffffff ? last_s : pick_s;
    *dst = tok_s;
  }
  __syncthreads();
  return tok_s;
}

__global__ void __launch_bounds__(256) sample_kernel(const float* __restrict__ logits, int V, float temperature, int top_k,
                                                     const float* __restrict__ uniforms, int u_stride,
                                                     const int* __restrict__ step, int64_t* __restrict__ out,
                                                     int out_stride, int out_offset, float* __restrict__ logits_out,
                                                     int lo_stride) {
  pdl_prologue();
  sample_row(logits, V, temperature, top_k, uniforms, u_stride, step ? *step : 0, out, out_stride, out_offset, logits_out, lo_stride);
}

// The tail of decode step t and the head of step t+1 in one launch, one block per clip:
//   token = sample(logits[b])  ->  x[b,:] = token_emb[token]  ->  LayerNorm of layer 0's self-attention (fp32 and/or bf16 planes)
//   -> the last block to finish advances the device step counter.
// Replaces 4 launches of the per-step chain (sample, step++, embed, LayerNorm).  D <= 256 * 4 * MAXV.
template <int MAXV>
__global__ void __launch_bounds__(256) sample_next_kernel(const float* __restrict__ logits, int V, float temperature, int top_k,
                                                          const float* __restrict__ uniforms, int u_stride, int* step,
                                                          unsigned int* ticket, int64_t* __restrict__ out, int out_stride,
                                                          int out_offset, float* __restrict__ logits_out, int lo_stride,
                                                          const float* __restrict__ emb, float* __restrict__ x, int D,
                                                          const float* __restrict__ gain, const float* __restrict__ bias,
                                                          float* __restrict__ y, __nv_bfloat16* __restrict__ yp, int planes, float eps,
                                                          const float* __restrict__ pos, float pos_scale) {
  __shared__ float redn[8];
  pdl_prologue();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int st = *step;
  int tok = sample_row(logits, V, temperature, top_k, uniforms, u_stride, st, out, out_stride, out_offset, logits_out, lo_stride);
  tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
  const int n4 = D >> 2;
  const float4* src = reinterpret_cast<const float4*>(emb + (size_t)tok * D);
  const float4* ps = pos ? reinterpret_cast<const float4*>(pos + (size_t)(st + 1) * D) : nullptr;     // the token enters at position st + 1
  float4 v[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + 256 * i;
    if (c < n4) {
      v[i] = __ldg(src + c);
      if (ps) {
        const float4 q = __ldg(ps + c);
        v[i].x = fmaf(q.x, pos_scale, v[i].x); v[i].y = fmaf(q.y, pos_scale, v[i].y);
        v[i].z = fmaf(q.z, pos_scale, v[i].z); v[i].w = fmaf(q.w, pos_scale, v[i].w);
      }
      reinterpret_cast<float4*>(x + (size_t)b * D)[c] = v[i];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  sum = warp_sum(sum);
  if (lane == 0) redn[warp] = sum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += redn[w];
  const float mean = tot / (float)D;
  __syncthreads();
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + 256 * i;
    if (c < n4) {
      const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
      q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
  }
  q = warp_sum(q);
  if (lane == 0) redn[warp] = q;
  __syncthreads();
  float qt = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) qt += redn[w];
  const float rstd = rsqrtf(qt / (float)D + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + 256 * i;
    if (c < n4) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gain) + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x; o.y = (v[i].y - mean) * rstd * g.y;
      o.z = (v[i].z - mean) * rstd * g.z; o.w = (v[i].w - mean) * rstd * g.w;
      if (bias) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + c);
        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
      }
      if (y) reinterpret_cast<float4*>(y + (size_t)b * D)[c] = o;
      if (yp) store_planes4(yp + (size_t)b * planes * D + c * 4, o, planes, D);
    }
  }
  // every block has read *step long before the LAST one gets here
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *step = st + 1;
      *ticket = 0u;
    }
  }
}


// ---- input side: audio-feature resampling to the motion frame rate ------------------------------------------
// mode 0: vico_preprocessing.downsample_mean (code/vico_preprocessing.py:7-19): out[i] = mean(in[i*w : i*w + w]), w = int(t / new_t)
// mode 1: dataset/l2l.downsample_mean (code/dataset/l2l.py:23-29): F.interpolate(mode='linear', align_corners=True)
// One thread per 4 consecutive channels of one output frame: coalesced 128-bit loads and stores, pure HBM streaming.
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ in, float* __restrict__ out, int t, int d4,
                                                       int new_t, int window, int mode) {
  const size_t total = (size_t)new_t * d4;
  const float4* in4 = reinterpret_cast<const float4*>(in);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / d4), c = (int)(i - (size_t)row * d4);
    float4 o;
    if (mode == 0) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const int start = row * window;
      for (int k = 0; k < window; ++k) {
        const float4 v = __ldcs(in4 + (size_t)(start + k) * d4 + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      const float inv = 1.0f / (float)window;
      o = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    } else {
      // ATen area_pixel_compute_source_index(align_corners=True): src = scale * dst, scale = (t-1)/(new_t-1) in fp32
      const float scale = new_t > 1 ? (float)(t - 1) / (float)(new_t - 1) : 0.f;
      const float src = scale * (float)row;
      const int i0 = (int)src, i1 = i0 + (i0 < t - 1 ? 1 : 0);
      const float w1 = src - (float)i0, w0 = 1.0f - w1;
      const float4 a = __ldcs(in4 + (size_t)i0 * d4 + c), b = __ldcs(in4 + (size_t)i1 * d4 + c);
      o = make_float4(__fadd_rn(__fmul_rn(w0, a.x), __fmul_rn(w1, b.x)), __fadd_rn(__fmul_rn(w0, a.y), __fmul_rn(w1, b.y)),
                      __fadd_rn(__fmul_rn(w0, a.z), __fmul_rn(w1, b.z)), __fadd_rn(__fmul_rn(w0, a.w), __fmul_rn(w1, b.w)));
    }
    __stcs(reinterpret_cast<float4*>(out) + i, o);
  }
}

// tokens[r, 0] = prompt[r / samples]   (every sample of a clip starts from the clip's prompt token)
__global__ void init_tokens_kernel(int64_t* __restrict__ tokens, int stride, const int64_t* __restrict__ prompt, int rows, int samples) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) tokens[(size_t)r * stride] = prompt[r / samples];
}

// Key-padding masks are almost always prefix masks (keys 0 .. L_b-1 kept, seq2seq_pretrain.py:433-436 builds them from src_len):
// out[b] = L_b when clip b's mask is one, -1 otherwise (arbitrary mask: the attention items read its bytes).  One warp per clip.
__global__ void mask_prefix_kernel(const uint8_t* __restrict__ mask, int B, int T, int32_t* __restrict__ out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const uint8_t* m = mask + (size_t)b * T;
  int total = 0, first_zero = T;
  for (int j = lane; j < T; j += 32) {
    if (m[j]) ++total;
    else first_zero = min(first_zero, j);
  }
  total = __reduce_add_sync(0xffffffffu, total);
  first_zero = __reduce_min_sync(0xffffffffu, first_zero);
  if (lane == 0) out[b] = total == first_zero ? first_zero : -1;
}

__global__ void advance_step_kernel(int* step) {
  pdl_prologue();
  *step += 1;
}
__global__ void set_step_kernel(int* step, int v) {
  *step = v;
  step[16] = 0;          // the block ticket of sample_next_kernel lives 64 bytes behind the counter
}

}  // namespace

int launch_layer_norm(const float* x, const float* gain, const float* bias, float* y, __nv_bfloat16* yb, int rows, int dim,
                      float eps, cudaStream_t s, __nv_bfloat16* yp, int planes, int kp) {
  DIM_REQUIRE(yp == nullptr || (planes >= 1 && planes <= 3 && kp == dim), "layer_norm: plane output needs kp == dim");
  DIM_REQUIRE(rows > 0 && dim > 0 && dim % 4 == 0 && dim <= 4096, "layer_norm: dim must be a multiple of 4, <= 4096");
  ProfScope ps(CAT_LAYERNORM, s, (y ? 8.0 : 4.0) * rows * dim + (yb ? 2.0 * rows * dim : 0.0) + (yp ? 2.0 * planes * rows * dim : 0.0),
               8.0 * rows * dim);
  if (dim <= 384) DIM_CHECK_CUDA(launch_k(layer_norm_kernel<3>, dim3(cdiv(rows, 8)), dim3(256), 0, s, x, gain, bias, y, yb, rows, dim, eps, yp, planes, kp));
  else if (dim <= 1152) DIM_CHECK_CUDA(launch_k(layer_norm_kernel<9>, dim3(cdiv(rows, 8)), dim3(256), 0, s, x, gain, bias, y, yb, rows, dim, eps, yp, planes, kp));
  else DIM_CHECK_CUDA(launch_k(layer_norm_kernel<32>, dim3(cdiv(rows, 8)), dim3(256), 0, s, x, gain, bias, y, yb, rows, dim, eps, yp, planes, kp));
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_instance_norm(float* x, const int32_t* lens, int B, int T, int C, float eps, cudaStream_t s) {
  DIM_REQUIRE(B > 0 && T > 0 && C > 0, "instance_norm: empty");
  ProfScope ps(CAT_INSTNORM, s, 8.0 * B * T * C, 6.0 * B * T * C);
  instance_norm_kernel<<<dim3(cdiv(C, 32), B), 256, 0, s>>>(x, lens, T, C, eps);
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_build_context(const float* xs, const float* pe_dec, const float* audio, float* ctx, __nv_bfloat16* ctxb,
                         size_t rows, int d1, int d2, cudaStream_t s) {
  DIM_REQUIRE(d1 % 4 == 0 && d2 % 4 == 0, "context dims must be multiples of 4");
  size_t total = rows * ((d1 + d2) / 4);
  int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  ProfScope ps(CAT_MISC, s, 8.0 * rows * (d1 + d2), 0);
  build_context_kernel<<<blocks, 256, 0, s>>>(xs, pe_dec, audio, ctx, ctxb, rows, d1, d2);
  DIM_LAUNCHED();
  return DIM_OK;
}

// Collate on the device: clip b owns rows [offsets[b], offsets[b+1]) of the packed (sum_len, D) arrays.  One thread per float4 of an
// output row: src[b,t] = speaker[row] (or 1.0 when speaker == nullptr: ViCoDataset replaces the speaker motion by ones,
// data_loader.py:147) | audio[row] (or 0.0: LmListenerDataset's audio, data_loader.py:242); tgt[b,t] = listener[row]; rows past the
// clip's length are zero (pad_sequence padding_value 0, data_loader.py:432-433); mask[b,t] = t < len (x_engine_pt.py:246-249).
__global__ void __launch_bounds__(256) assemble_batch_kernel(const float* __restrict__ speaker, const float* __restrict__ audio,
                                                             const float* __restrict__ listener, const int64_t* __restrict__ offsets,
                                                             int B, int T, int Dm4, int Da4, float* __restrict__ src,
                                                             float* __restrict__ tgt, uint8_t* __restrict__ mask) {
  const int W4 = Dm4 + Da4 + Dm4;                       // float4 per (b, t): src motion | src audio | tgt
  const size_t total = (size_t)B * T * W4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bt = i / W4;
    const int c = (int)(i - bt * W4), b = (int)(bt / T), t = (int)(bt - (size_t)b * T);
    const int64_t r0 = offsets[b];
    const int len = (int)(offsets[b + 1] - r0);
    const bool on = t < len;
    const size_t row = (size_t)(r0 + t);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < Dm4) {
      if (on) v = speaker ? __ldcs(reinterpret_cast<const float4*>(speaker) + row * Dm4 + c) : make_float4(1.f, 1.f, 1.f, 1.f);
      __stcs(reinterpret_cast<float4*>(src) + bt * (Dm4 + Da4) + c, v);
      if (c == 0 && mask) mask[bt] = on ? 1 : 0;
    } else if (c < Dm4 + Da4) {
      if (on && audio) v = __ldcs(reinterpret_cast<const float4*>(audio) + row * Da4 + (c - Dm4));
      __stcs(reinterpret_cast<float4*>(src) + bt * (Dm4 + Da4) + c, v);
    } else {
      if (on) v = __ldcs(reinterpret_cast<const float4*>(listener) + row * Dm4 + (c - Dm4 - Da4));
      __stcs(reinterpret_cast<float4*>(tgt) + bt * Dm4 + (c - Dm4 - Da4), v);
    }
  }
}
int launch_assemble_batch(const float* speaker, const float* audio, const float* listener, const int64_t* offsets, int B, int T, int Dm,
                          int Da, float* src, float* tgt, uint8_t* mask, cudaStream_t s) {
  DIM_REQUIRE(B > 0 && T > 0 && Dm % 4 == 0 && Da % 4 == 0 && Dm > 0 && Da > 0, "assemble_batch: bad sizes");
  DIM_REQUIRE(listener && offsets && src && tgt, "assemble_batch: null argument");
  const size_t total = (size_t)B * T * (2 * Dm + Da) / 4;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 32);
  ProfScope ps(CAT_MISC, s, (double)total * 32.0, 0);
  assemble_batch_kernel<<<blocks, 256, 0, s>>>(speaker, audio, listener, offsets, B, T, Dm / 4, Da / 4, src, tgt, mask);
  DIM_LAUNCHED();
  return DIM_OK;
}

__global__ void __launch_bounds__(256) add_pos_table_kernel(float* __restrict__ x, const float* __restrict__ tab, float scale, int L, int D4,
                                                            size_t total4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / D4;
    const int c = (int)(i - row * D4), t = (int)(row % L);
    float4 v = reinterpret_cast<float4*>(x)[i];
    const float4 p = __ldg(reinterpret_cast<const float4*>(tab) + (size_t)t * D4 + c);
    v.x = fmaf(p.x, scale, v.x); v.y = fmaf(p.y, scale, v.y); v.z = fmaf(p.z, scale, v.z); v.w = fmaf(p.w, scale, v.w);
    reinterpret_cast<float4*>(x)[i] = v;
  }
}
int launch_add_pos_table(float* x, const float* tab, float scale, int B, int L, int D, cudaStream_t s) {
  DIM_REQUIRE(D % 4 == 0 && B > 0 && L > 0, "add_pos_table: bad sizes");
  const size_t total4 = (size_t)B * L * (D / 4);
  const int blocks = (int)std::min<size_t>((total4 + 255) / 256, (size_t)148 * 16);
  ProfScope ps(CAT_MISC, s, (double)total4 * 32.0, 0);
  add_pos_table_kernel<<<blocks, 256, 0, s>>>(x, tab, scale, L, D / 4, total4);
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_embed_tokens(const int64_t* tok, int tok_stride, const int* step, const float* emb, float* x, int B, int D, int V,
                        cudaStream_t s, const float* pos, float pos_scale) {
  ProfScope ps(CAT_MISC, s, 8.0 * B * D, 0);
  DIM_CHECK_CUDA(launch_k(embed_tokens_kernel, dim3(B), dim3(128), 0, s, tok, tok_stride, step, emb, x, B, D, V, pos, pos_scale));
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_sample(const float* logits, int B, int V, float temperature, int top_k, const float* uniforms, int u_stride,
                  const int* step, int64_t* out, int out_stride, int out_offset, float* logits_out, int lo_stride,
                  cudaStream_t s) {
  DIM_REQUIRE(V > 0 && V <= 1024, "sample: vocabulary must be <= 1024");
  DIM_REQUIRE(temperature == 0.f || (uniforms != nullptr && top_k > 0), "sample: sampling needs uniforms and top_k");
  ProfScope ps(CAT_SAMPLE, s, 4.0 * B * V + 8.0 * B, 0);
  DIM_CHECK_CUDA(launch_k(sample_kernel, dim3(B), dim3(256), 0, s, logits, V, temperature, top_k, uniforms, u_stride, step, out,
                          out_stride, out_offset, logits_out, lo_stride));
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_sample_next(const float* logits, int B, int V, float temperature, int top_k, const float* uniforms, int u_stride,
                       int* step, unsigned int* ticket, int64_t* out, int out_stride, int out_offset, float* logits_out,
                       int lo_stride, const float* emb, float* x, int D, const float* gain, const float* bias, float* y,
                       __nv_bfloat16* yp, int planes, float eps, cudaStream_t s, const float* pos, float pos_scale) {
  DIM_REQUIRE(V > 0 && V <= 1024, "sample: vocabulary must be <= 1024");
  DIM_REQUIRE(temperature == 0.f || (uniforms != nullptr && top_k > 0), "sample: sampling needs uniforms and top_k");
  DIM_REQUIRE(D % 4 == 0 && D <= 256 * 4 * 4, "sample_next: model dim must be a multiple of 4, <= 4096");
  ProfScope ps(CAT_SAMPLE, s, 4.0 * B * V + 8.0 * B + 12.0 * B * D, 8.0 * B * D);
  if (D <= 256 * 4 * 2)
    DIM_CHECK_CUDA(launch_k(sample_next_kernel<2>, dim3(B), dim3(256), 0, s, logits, V, temperature, top_k, uniforms, u_stride, step,
                            ticket, out, out_stride, out_offset, logits_out, lo_stride, emb, x, D, gain, bias, y, yp, planes, eps, pos, pos_scale));
  else
    DIM_CHECK_CUDA(launch_k(sample_next_kernel<4>, dim3(B), dim3(256), 0, s, logits, V, temperature, top_k, uniforms, u_stride, step,
                            ticket, out, out_stride, out_offset, logits_out, lo_stride, emb, x, D, gain, bias, y, yp, planes, eps, pos, pos_scale));
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_resample(const float* in, float* out, int t, int d, int new_t, int window, int mode, cudaStream_t s) {
  DIM_REQUIRE(in && out && t > 0 && d > 0 && d % 4 == 0 && new_t > 0, "resample: bad sizes (d must be a multiple of 4)");
  DIM_REQUIRE(mode == 1 || (window >= 1 && (long)new_t * window <= t), "resample: window mean reads past the input");
  const size_t total = (size_t)new_t * (d / 4);
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 16);
  ProfScope ps(CAT_MISC, s, 4.0 * d * ((mode == 0 ? (double)new_t * window : 2.0 * new_t) + new_t), 0);
  resample_kernel<<<blocks, 256, 0, s>>>(in, out, t, d / 4, new_t, window, mode);
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_init_tokens(int64_t* tokens, int stride, const int64_t* prompt, int rows, int samples, cudaStream_t s) {
  DIM_REQUIRE(rows > 0 && samples >= 1, "init_tokens: bad sizes");
  ProfScope ps(CAT_MISC, s, 16.0 * rows, 0);
  init_tokens_kernel<<<cdiv(rows, 256), 256, 0, s>>>(tokens, stride, prompt, rows, samples);
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_mask_prefix(const uint8_t* mask, int B, int T, int32_t* out, cudaStream_t s) {
  DIM_REQUIRE(mask && out && B > 0 && T > 0, "mask_prefix: bad arguments");
  ProfScope ps(CAT_MISC, s, (double)B * T + 4.0 * B, 0);
  mask_prefix_kernel<<<cdiv(B, 8), 256, 0, s>>>(mask, B, T, out);
  DIM_LAUNCHED();
  return DIM_OK;
}

int launch_advance_step(int* step, cudaStream_t s) {
  ProfScope ps(CAT_MISC, s, 4, 0);
  DIM_CHECK_CUDA(launch_k(advance_step_kernel, dim3(1), dim3(1), 0, s, step));
  DIM_LAUNCHED();
  return DIM_OK;
}
int launch_set_step(int* step, int v, cudaStream_t s) {
  ProfScope ps(CAT_MISC, s, 4, 0);
  set_step_kernel<<<1, 1, 0, s>>>(step, v);
  DIM_LAUNCHED();
  return DIM_OK;
}

}  // namespace dimb

using namespace dimb;

extern "C" int dim_resample_features(const float* in, int t, int d, int new_t, int window, int mode, float* out, void* stream) {
  if (int e = ensure_device()) return e;
  return launch_resample(in, out, t, d, new_t, window, mode, as_stream(stream));
}

extern "C" int dim_assemble_batch(const float* speaker, const float* audio, const float* listener, const int64_t* offsets, int B, int T,
                                  int motion_dim, int audio_dim, float* src, float* tgt, uint8_t* mask, void* stream) {
  if (int e = ensure_device()) return e;
  return launch_assemble_batch(speaker, audio, listener, offsets, B, T, motion_dim, audio_dim, src, tgt, mask, as_stream(stream));
}

extern "C" int dim_layer_norm_f32(const float* x, const float* gain, const float* bias, float* y, int rows, int dim,
                                  float eps, void* stream) {
  if (int e = ensure_device()) return e;
  return launch_layer_norm(x, gain, bias, y, nullptr, rows, dim, eps, as_stream(stream), nullptr, 0, 0);
}

extern "C" int dim_instance_norm_f32(float* x, const int32_t* lens, int B, int T, int C, float eps, void* stream) {
  if (int e = ensure_device()) return e;
  return launch_instance_norm(x, lens, B, T, C, eps, as_stream(stream));
}
