// Error plumbing, device check, launch counter.
#include "common.cuh"

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

}  // namespace dimb

extern "C" const char* dim_last_error(void) { return dimb::g_err.c_str(); }
extern "C" int dim_version(void) { return 100; }
extern "C" uint64_t dim_launch_count(void) { return dimb::g_launches.load(); }
