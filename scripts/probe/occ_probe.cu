// Which feature of the persistent decode kernel pins its occupancy at 1 CTA/SM?  nvcc -arch=sm_100a occ_probe.cu && ./a.out
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct Big { char b[25000]; };
struct Small { char b[3000]; };
__global__ void __launch_bounds__(256, 2) k_plain(int* o) { if (o) o[threadIdx.x] = 1; }
__global__ void __launch_bounds__(256, 2) k_bigparam(const __grid_constant__ Big p, int* o) { if (o) o[threadIdx.x] = p.b[threadIdx.x]; }
__global__ void __launch_bounds__(256, 2) k_smallparam(const __grid_constant__ Small p, int* o) { if (o) o[threadIdx.x] = p.b[threadIdx.x]; }
__global__ void __launch_bounds__(256, 2) k_tmem(int* o) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(128) : "memory");
  if (o) o[threadIdx.x] = 1;
}
__global__ void __launch_bounds__(256, 2) k_tmem_imm(int* o) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(slot) : "memory");
  if (o) o[threadIdx.x] = 1;
}
// real co-residency: every block allocates 128 TMEM columns, then waits (bounded) until `want` blocks are resident at once
__global__ void __launch_bounds__(256, 2) k_tmem_coresident(unsigned int* ctr, unsigned int want, int* result) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(ctr, 1u);
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int ok = 0;
    for (;;) {
      if (*(volatile unsigned int*)ctr >= want) { ok = 1; break; }
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > 200000000ull) break;            // 0.2 s
    }
    if (blockIdx.x == 0) *result = ok;
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(128) : "memory");
}
__global__ void __launch_bounds__(256, 2) k_namedbar(int* o) {
  asm volatile("bar.sync 1, 64;" ::: "memory");
  if (o) o[threadIdx.x] = 1;
}
__global__ void __launch_bounds__(256, 2) k_trap(int* o) {
  if (o && o[0] == 12345) __trap();
  if (o) o[threadIdx.x] = 1;
}
__global__ void __launch_bounds__(256, 2) k_timer(unsigned long long* o) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if (o) o[threadIdx.x] = t;
}
__global__ void __launch_bounds__(256, 2) k_mbar(int* o) {
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (o) o[threadIdx.x] = 1;
}
template <typename K> void report(const char* name, K k) {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 99328);
  cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int a = -1, b = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k, 256, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k, 256, 99328);
  cudaFuncAttributes fa{};
  cudaFuncGetAttributes(&fa, k);
  printf("%-14s regs=%3d  blocks/SM at smem 0: %d   at 99328: %d\n", name, fa.numRegs, a, b);
}
int main() {
  report("plain", k_plain);
  report("bigparam", k_bigparam);
  report("smallparam", k_smallparam);
  report("tmem", k_tmem);
  report("tmem_imm", k_tmem_imm);
  report("namedbar", k_namedbar);
  report("trap", k_trap);
  report("timer", k_timer);
  report("mbarrier", k_mbar);
  {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned int* ctr; int* res;
    cudaMalloc(&ctr, 4); cudaMalloc(&res, 4);
    for (int mult = 1; mult <= 2; ++mult) {
      cudaMemset(ctr, 0, 4); cudaMemset(res, 0xff, 4);
      k_tmem_coresident<<<sms * mult, 256, 32768>>>(ctr, (unsigned)(sms * mult), res);
      cudaError_t e = cudaDeviceSynchronize();
      int r = -2; cudaMemcpy(&r, res, 4, cudaMemcpyDeviceToHost);
      printf("plain launch of %d x SMs TMEM blocks (128 columns each): all co-resident = %d (%s)\n", mult, r, cudaGetErrorString(e));
    }
  }
  return 0;
}
