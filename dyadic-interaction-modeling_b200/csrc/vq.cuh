// VQ codebook kernels.  See vq.cu.
#pragma once
#include "common.cuh"
#include <algorithm>

namespace dimb {
int launch_vq_argmin(const float* z, const float* E, int64_t* idx, int N, int D, int K, cudaStream_t s);
// tensor-core shortlist + exact fp32 re-rank (vq_tc.cu): same indices as the exact kernel, D = 128, K = 512 only
bool vq_argmin_tc_supported(int N, int D, int K);
int launch_vq_argmin_tc(const float* z, const float* E, int64_t* idx, int N, int* stats, cudaStream_t s, long long* trace = nullptr);
int launch_vq_gather(const int64_t* idx, const float* E, float* out, int N, int D, int K, int32_t* bad, cudaStream_t s);
int launch_vq_gather_bcl(const int64_t* idx, const float* E, float* out, int B, int L, int D, int K, cudaStream_t s);
int launch_rows_from_bcl(const float* q, float* rows, int B, int L, int D, cudaStream_t s);
}  // namespace dimb
