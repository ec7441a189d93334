#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace csr {
size_t metrics_scratch_bytes(int n, int h, int w);
// Enqueues the 4 kernels of the masked metric step; out = CSR_NUM_METRICS floats (see include/climsr_b200.h).
cudaError_t launch_masked_metrics(const float* sr, const float* hr, const float* orig, const float* mask, const double* mn, const double* mx,
                                  float zmean, float zstd, double ra, double rb, double eps, int n, int h, int w, float* out, void* scratch,
                                  cudaStream_t s, int* launches);
}  // namespace csr
