// One (bi)directional nn.LSTM layer, batch_first, zero initial state, fp32.  sm_100a.
// Replaces the `vertice_map_reverse_lstm` head of EmocaConverter / SpeakerSLMFT
//   (/root/reference/code/seq2seq_pretrain.py:789-802 construction, :657 and :823 call sites; nn.LSTM gate order i, f, g, o).
//
// Two stages per layer:
//   1. gx[d] = x @ W_ih[d]^T + b_ih[d] for every (b, t) at once, one GEMM per direction: the tiled FFMA GEMM (gemm_f32.cu) below
//      2048 rows, the tcgen05 GEMM on 3-plane bf16 splits of x and W_ih (fp32-grade products, gemm_tc.cu) from there on.
//   2. the recurrence: ONE cooperative launch for all T steps and both directions.  The hidden units are sliced across CTAs
//      (UPC units = 4*UPC gate rows of W_hh per CTA, resident in shared memory for the whole sequence); per step a CTA reads
//      h_{t-1} of every batch row from the layer output itself (L2), adds its slice of h W_hh^T to gx, applies the gates, keeps
//      its slice of c in global scratch only it touches, writes its slice of h_t straight into out[b, t, d*H + j], and meets the
//      other CTAs of ITS direction at a release/acquire counter barrier.  The two directions never wait for each other.
// Summation order of the K = H dot products: KPN interleaved partial sums (16 for <= 16 batch rows, 4 for <= 64, else 1), each
// ascending in k, then a butterfly: deterministic, and independent of the batch size within each of the three layouts.
#include "common.cuh"
#include "gemm_f32.cuh"
#include "gemm_tc.cuh"

namespace dimb {
namespace {

constexpr int LSTM_THREADS = 256;
constexpr int LSTM_KC = 64;                    // k-chunk of h staged in shared memory
constexpr unsigned long long LSTM_TIMEOUT_NS = 4000000000ull;

struct LstmArgs {
  const float* gx[2];      // [B*T, 4H]   x W_ih^T + b_ih
  const float* whh[2];     // [4H, H]
  const float* bhh[2];     // [4H]
  float* out;              // [B, T, ndir*H]
  float* c;                // [ndir, B, H] scratch
  unsigned int* bar;       // [ndir] zeroed counters, 128 B apart (32 uints)
  int B, T, H, ndir, cpd;  // cpd: CTAs per direction
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Thread layout of one pass over BP = 1024 / (4 * KPN) ... batch rows: tid = (rg * 4 + gg) * KPN + kp
//   rg: row group -- the thread owns rows rg + NRG*i, i < 4 (NRG = BP / 4);  gg: gate (i, f, g, o) -> UPC gate rows of W_hh;
//   kp: k lane -- the K = H dot products are split over KPN lanes (butterfly at the end).
// Every W_hh value read from shared memory feeds 4 FMAs (4 rows) and every h value UPC FMAs: the loop is bound by the FMA
// pipe and the 128 B/clk shared-memory port together instead of by the port alone.
template <int UPC, int KPN>
__global__ void __launch_bounds__(LSTM_THREADS) lstm_recurrence(const LstmArgs p) {
  constexpr int R = 4 * UPC, RT = 4;
  constexpr int NRG = LSTM_THREADS / (4 * KPN), BP = NRG * RT;
  constexpr int ITEMS = (BP * UPC + LSTM_THREADS - 1) / LSTM_THREADS;
  constexpr int HPAD = KPN == 4 ? 16 : 4;
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, B = p.B, T = p.T;
  const int kc = KPN == 1 ? LSTM_KC : H;               // h is staged whole for the small-batch layouts (one L2 round trip per step)
  const int ldw = H + 4, ldk = kc + HPAD, ldp = BP + 1;
  float* Wsm = smem;                                   // [R][ldw]
  float* hs = Wsm + (size_t)R * ldw;                   // [BP][ldk]
  float* pre = hs + (size_t)(KPN == 1 ? 2 : 1) * BP * ldk;   // [R][ldp]  (two h buffers when h is staged in chunks)
  const int d = blockIdx.x / p.cpd, slice = blockIdx.x % p.cpd, j0 = slice * UPC;
  const int tid = threadIdx.x;
  const int kp = tid % KPN, gg = (tid / KPN) % 4, rg = tid / (4 * KPN);

  for (int i = tid; i < R * H; i += LSTM_THREADS) {
    const int r = i / H, k = i - r * H, gate = r / UPC, u = r - gate * UPC;
    Wsm[(size_t)r * ldw + k] = __ldg(p.whh[d] + (size_t)(gate * H + j0 + u) * H + k);
  }
  __syncthreads();

  const size_t ldo = (size_t)p.ndir * H;
  const float* gxd = d == 0 ? p.gx[0] : p.gx[1];
  const float* bhd = d == 0 ? p.bhh[0] : p.bhh[1];
  unsigned int* bar = p.bar + 32 * d;
  unsigned int target = 0;
  const float* wrow = Wsm + (size_t)(gg * UPC) * ldw;

  for (int s = 0; s < T; ++s) {
    const int t = d == 0 ? s : T - 1 - s, tprev = d == 0 ? t - 1 : t + 1;
    for (int b0 = 0; b0 < B; b0 += BP) {
      const int nb = min(BP, B - b0);
      // this step's input gates do not depend on h: request them before the recurrence work
      float gxr[ITEMS][4];
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int idx = tid + it * LSTM_THREADS, bl = idx / UPC, u = idx - bl * UPC;
        if (bl < nb) {
          const float* gx = gxd + ((size_t)(b0 + bl) * T + t) * 4 * H + j0 + u;
#pragma unroll
          for (int g = 0; g < 4; ++g) gxr[it][g] = __ldg(gx + g * H);
        }
      }
      float acc[RT][UPC];
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int g = 0; g < UPC; ++g) acc[i][g] = 0.f;
      if (s > 0) {
        // h_{t-1} reaches shared memory through cp.async (L2 -> smem, no registers), one k-chunk ahead of the FMAs: two buffers
        const int q4 = kc / 4, total = nb * q4;
        auto stage = [&](int k0, float* dst) {
          for (int i = tid; i < total; i += LSTM_THREADS) {
            const int b = i / q4, q = (i - b * q4) * 4;
            const float* src = p.out + ((size_t)(b0 + b) * T + tprev) * ldo + (size_t)d * H + k0 + q;
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst + (size_t)b * ldk + q);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        };
        __syncthreads();                               // the previous pass / step has finished with both buffers
        stage(0, hs);
        int buf = 0;
        for (int k0 = 0; k0 < H; k0 += kc, buf ^= 1) {
          const float* hcur = hs + (size_t)buf * BP * ldk;
          if (k0 + kc < H) {
            stage(k0 + kc, hs + (size_t)(buf ^ 1) * BP * ldk);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
          } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
          }
          __syncthreads();
          for (int kk = kp * 4; kk < kc; kk += KPN * 4) {
            float4 h[RT];
#pragma unroll
            for (int i = 0; i < RT; ++i) {
              const int bl = rg + NRG * i;
              h[i] = bl < nb ? *reinterpret_cast<const float4*>(hcur + (size_t)bl * ldk + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int g = 0; g < UPC; ++g) {
              const float4 w = *reinterpret_cast<const float4*>(wrow + (size_t)g * ldw + k0 + kk);
#pragma unroll
              for (int i = 0; i < RT; ++i)
                acc[i][g] = fmaf(h[i].w, w.w, fmaf(h[i].z, w.z, fmaf(h[i].y, w.y, fmaf(h[i].x, w.x, acc[i][g]))));
            }
          }
          __syncthreads();                             // everyone is done with hcur before the chunk after next lands in it
        }
#pragma unroll
        for (int o = KPN >> 1; o > 0; o >>= 1) {
#pragma unroll
          for (int i = 0; i < RT; ++i)
#pragma unroll
            for (int g = 0; g < UPC; ++g) acc[i][g] += __shfl_xor_sync(0xffffffffu, acc[i][g], o);
        }
      }
      if (kp == 0) {
#pragma unroll
        for (int i = 0; i < RT; ++i)
#pragma unroll
          for (int g = 0; g < UPC; ++g) pre[(size_t)(gg * UPC + g) * ldp + rg + NRG * i] = acc[i][g];
      }
      __syncthreads();
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int idx = tid + it * LSTM_THREADS, bl = idx / UPC, u = idx - bl * UPC;
        if (bl < nb) {
          const int b = b0 + bl;
          // (x W_ih^T + b_ih) + (h W_hh^T + b_hh), as nn.LSTM groups it
          const float gi = gxr[it][0] + (pre[(size_t)(0 * UPC + u) * ldp + bl] + __ldg(bhd + 0 * H + j0 + u));
          const float gf = gxr[it][1] + (pre[(size_t)(1 * UPC + u) * ldp + bl] + __ldg(bhd + 1 * H + j0 + u));
          const float gc = gxr[it][2] + (pre[(size_t)(2 * UPC + u) * ldp + bl] + __ldg(bhd + 2 * H + j0 + u));
          const float go = gxr[it][3] + (pre[(size_t)(3 * UPC + u) * ldp + bl] + __ldg(bhd + 3 * H + j0 + u));
          float* cc = p.c + ((size_t)d * B + b) * H + j0 + u;
          const float cprev = s > 0 ? *cc : 0.f;
          const float cn = sigmoidf_(gf) * cprev + sigmoidf_(gi) * tanhf(gc);
          *cc = cn;
          __stcg(p.out + ((size_t)b * T + t) * ldo + (size_t)d * H + j0 + u, sigmoidf_(go) * tanhf(cn));
        }
      }
      // pre / hs are rewritten by the next pass only after its own __syncthreads()s
    }
    if (s + 1 < T) {                                   // h_t of this direction must be visible before anyone starts step s+1
      __threadfence();
      __syncthreads();
      target += p.cpd;
      if (tid == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned long long t0 = 0;
        for (uint32_t it = 0;; ++it) {
          unsigned int v;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
          if ((int)(v - target) >= 0) break;
          if (it == 64) t0 = gtime();
          if (it > 64 && (it & 63) == 0 && gtime() - t0 > LSTM_TIMEOUT_NS) __trap();
        }
      }
      __syncthreads();
    }
  }
}

int pick_kpn(int B) { return B <= 16 ? 16 : (B <= 64 ? 4 : 1); }

int pick_upc(int H, int ndir, int sms) {
  for (int upc : {4, 6, 8})
    if (H % upc == 0 && ndir * (H / upc) <= sms) return upc;
  return 0;
}

size_t rec_smem(int upc, int H, int B) {
  const int kpn = pick_kpn(B), bp = LSTM_THREADS / (4 * kpn) * 4;
  const int kc = kpn == 1 ? LSTM_KC : H, hpad = kpn == 4 ? 16 : 4;
  return ((size_t)4 * upc * (H + 4) + (size_t)(kpn == 1 ? 2 : 1) * bp * (kc + hpad) + (size_t)4 * upc * (bp + 1)) * sizeof(float);
}

template <int UPC>
void* rec_kernel(int kpn) {
  return kpn == 16 ? (void*)lstm_recurrence<UPC, 16> : kpn == 4 ? (void*)lstm_recurrence<UPC, 4> : (void*)lstm_recurrence<UPC, 1>;
}

constexpr int LSTM_TC_ROWS = 2048;     // from this many (clip, frame) rows on, the input GEMMs run on the tensor cores (3 bf16 planes:
                                       // fp32-grade products, gemm_tc.cu); below it the FFMA GEMM is launch-latency bound anyway
struct WsLayout {
  size_t gx[2], c, bar, ap, wp[2], total;
  int kp;                              // 0: FFMA input GEMMs
};
WsLayout ws_layout(int B, int T, int in_dim, int H, int ndir) {
  WsLayout w{};
  size_t off = 0;
  for (int d = 0; d < ndir; ++d) { w.gx[d] = off; off += align_up((size_t)B * T * 4 * H * sizeof(float), 256); }
  w.c = off; off += align_up((size_t)ndir * B * H * sizeof(float), 256);
  w.bar = off; off += 256;
  if ((long)B * T >= LSTM_TC_ROWS) {
    w.kp = tc_round_k(in_dim);
    w.ap = off; off += align_up((size_t)B * T * 3 * w.kp * sizeof(__nv_bfloat16), 256);
    for (int d = 0; d < ndir; ++d) { w.wp[d] = off; off += align_up((size_t)4 * H * 3 * w.kp * sizeof(__nv_bfloat16), 256); }
  }
  w.total = off;
  return w;
}

}  // namespace
}  // namespace dimb

using namespace dimb;

extern "C" size_t dim_lstm_layer_workspace_bytes(int B, int T, int in_dim, int H, int ndir) {
  if (B <= 0 || T <= 0 || in_dim <= 0 || H <= 0 || ndir < 1 || ndir > 2) return 0;
  return ws_layout(B, T, in_dim, H, ndir).total;
}

extern "C" int dim_lstm_layer_f32(const float* x, int in_dim, const float* w_ih, const float* w_hh, const float* b_ih,
                                  const float* b_hh, const float* w_ih_r, const float* w_hh_r, const float* b_ih_r,
                                  const float* b_hh_r, int B, int T, int H, float* out, void* ws, size_t ws_bytes,
                                  void* stream) {
  if (int e = ensure_device()) return e;
  const int ndir = w_ih_r ? 2 : 1;
  DIM_REQUIRE(x && w_ih && w_hh && b_ih && b_hh && out && ws, "lstm: null operand");
  DIM_REQUIRE(ndir == 1 || (w_hh_r && b_ih_r && b_hh_r), "lstm: incomplete reverse direction");
  DIM_REQUIRE(B > 0 && T > 0 && in_dim > 0 && in_dim % 4 == 0 && H > 0 && H % LSTM_KC == 0, "lstm: bad shape (in_dim % 4, H % 64)");
  const WsLayout L = ws_layout(B, T, in_dim, H, ndir);
  DIM_REQUIRE(ws_bytes >= L.total, "lstm: workspace too small");
  cudaStream_t s = as_stream(stream);
  int dev = 0, sms = 0;
  DIM_CHECK_CUDA(cudaGetDevice(&dev));
  DIM_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int upc = pick_upc(H, ndir, sms);
  DIM_REQUIRE(upc != 0, "lstm: hidden size cannot be sliced over the SMs (H % 4/6/8, ndir*H/upc <= #SM)");

  char* base = static_cast<char*>(ws);
  const float* wi[2] = {w_ih, w_ih_r};
  const float* bi[2] = {b_ih, b_ih_r};
  LstmArgs a{};
  __nv_bfloat16* ap = L.kp ? reinterpret_cast<__nv_bfloat16*>(base + L.ap) : nullptr;
  if (ap) {                                            // x -> bf16 planes once, shared by both directions
    GemmArgs sp;
    sp.A = x; sp.lda = in_dim; sp.M = B * T; sp.K = in_dim;
    if (int e = launch_split_planes(sp, ap, L.kp, 3, s)) return e;
  }
  for (int d = 0; d < ndir; ++d) {
    GemmArgs g;
    g.A = x; g.lda = in_dim; g.W = wi[d]; g.bias = bi[d];
    g.C = reinterpret_cast<float*>(base + L.gx[d]); g.ldc = 4 * H;
    g.M = B * T; g.N = 4 * H; g.K = in_dim;
    if (ap) {
      __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(base + L.wp[d]);
      GemmArgs sw;
      sw.A = wi[d]; sw.lda = in_dim; sw.M = 4 * H; sw.K = in_dim;
      if (int e = launch_split_planes(sw, wp, L.kp, 3, s)) return e;
      g.split_hint = DIM_SPLIT_NEVER;
      if (int e = launch_gemm_tc(g, ap, wp, L.kp, 3, s)) return e;
    } else if (int e = launch_gemm_f32(g, s)) {
      return e;
    }
    a.gx[d] = g.C;
  }
  a.whh[0] = w_hh; a.whh[1] = w_hh_r;
  a.bhh[0] = b_hh; a.bhh[1] = b_hh_r;
  a.out = out;
  a.c = reinterpret_cast<float*>(base + L.c);
  a.bar = reinterpret_cast<unsigned int*>(base + L.bar);
  a.B = B; a.T = T; a.H = H; a.ndir = ndir; a.cpd = H / upc;
  DIM_CHECK_CUDA(cudaMemsetAsync(a.bar, 0, 256, s));

  const size_t smem = rec_smem(upc, H, B);
  DIM_REQUIRE(smem <= 227 * 1024, "lstm: hidden size too large for the resident W_hh slice");
  const int kpn = pick_kpn(B);
  void* kern = upc == 4 ? rec_kernel<4>(kpn) : upc == 6 ? rec_kernel<6>(kpn) : rec_kernel<8>(kpn);
  DIM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(CAT_MISC, s, 4.0 * ndir * ((double)B * T * 5 * H + 4.0 * H * H), 8.0 * ndir * (double)B * T * H * H);
  void* kargs[] = {&a};
  DIM_CHECK_CUDA(cudaLaunchCooperativeKernel(kern, dim3(ndir * a.cpd), dim3(LSTM_THREADS), kargs, smem, s));
  DIM_LAUNCHED();
  return DIM_OK;
}
