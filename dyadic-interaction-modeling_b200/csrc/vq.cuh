// VQ codebook kernels.  See vq.cu.
#pragma once
#include "common.cuh"
#include <algorithm>

namespace dimb {
int launch_vq_argmin(const float* z, const float* E, int64_t* idx, int N, int D, int K, cudaStream_t s);
int launch_vq_gather(const int64_t* idx, const float* E, float* out, int N, int D, int K, int32_t* bad, cudaStream_t s);
int launch_vq_gather_bcl(const int64_t* idx, const float* E, float* out, int B, int L, int D, int K, cudaStream_t s);
int launch_rows_from_bcl(const float* q, float* rows, int B, int L, int D, cudaStream_t s);
}  // namespace dimb
