// Error plumbing, device check, launch counter.
#include "common.cuh"
#include <cstdlib>
#include <vector>

namespace dimb {

static thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

int ensure_device() {
  static int cached = -1;   // 0 ok, else error code
  if (cached == 0) return DIM_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(DIM_ENODEVICE, "no CUDA device visible: libdimb200 has no CPU fallback");
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp p{};
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return fail(DIM_ENODEVICE, "cudaGetDeviceProperties failed");
  if (p.major != 10)
    return fail(DIM_ENODEVICE, std::string("device is sm_") + std::to_string(p.major) + std::to_string(p.minor) +
                                   "; libdimb200 is built for sm_100a only");
  cached = 0;
  return DIM_OK;
}

bool g_prof_on = false;
bool g_pdl_on = getenv("DIM_PDL") != nullptr;   // opt-in: measured slower on this workload (early-resident dependents hold SM slots)
namespace {
struct ProfRec { int cat; cudaEvent_t a, b; double bytes, flops; };
std::vector<ProfRec> g_recs;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
const char* kCatNames[CAT_COUNT] = {"gemm_f32_tiled", "gemm_f32_skinny", "conv5_implicit_gemm", "layer_norm",
                                    "instance_norm", "attn_prefill_f32", "attn_decode", "vq_argmin", "vq_gather",
                                    "sample", "misc", "gemm_bf16_tcgen05", "gemm_bf16_tcgen05_skinny",
                                    "decode_megakernel"};
}  // namespace
void prof_begin(int cat, cudaStream_t s, double bytes, double flops) {
  ProfRec r{cat, get_event(), get_event(), bytes, flops};
  cudaEventRecord(r.a, s);
  g_recs.push_back(r);
}
void prof_end(cudaStream_t s) { cudaEventRecord(g_recs.back().b, s); }

}  // namespace dimb

extern "C" int dim_profile_enable(int on) {
  dimb::g_prof_on = on != 0;
  return DIM_OK;
}
extern "C" const char* dim_profile_category_name(int cat) {
  return (cat >= 0 && cat < dimb::CAT_COUNT) ? dimb::kCatNames[cat] : "";
}
extern "C" int dim_profile_collect(dim_prof_entry* out, int max_entries, int* n_out) {
  using namespace dimb;
  if (cudaDeviceSynchronize() != cudaSuccess) return fail(DIM_ECUDA, "profile: device synchronize failed");
  dim_prof_entry acc[CAT_COUNT] = {};
  for (auto& r : g_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    acc[r.cat].category = r.cat;
    acc[r.cat].launches += 1;
    acc[r.cat].ms += ms;
    acc[r.cat].bytes += r.bytes;
    acc[r.cat].flops += r.flops;
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
  int n = 0;
  for (int c = 0; c < CAT_COUNT && n < max_entries; ++c)
    if (acc[c].launches) out[n++] = acc[c];
  if (n_out) *n_out = n;
  return DIM_OK;
}

extern "C" const char* dim_last_error(void) { return dimb::g_err.c_str(); }
extern "C" int dim_version(void) { return 100; }
extern "C" uint64_t dim_launch_count(void) { return dimb::g_launches.load(); }
