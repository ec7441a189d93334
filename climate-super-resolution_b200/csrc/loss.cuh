#pragma once
#include <cuda_runtime.h>

namespace csr {
int pixel_loss_blocks(long n);
// mode 0 = L1, 1 = MSE.  out[0] = mean loss; grad (nullable) = d loss / d sr; partial: pixel_loss_blocks(n) doubles.
cudaError_t launch_pixel_loss(int mode, const float* sr, const float* hr, float* grad, long n, float* out, double* partial, cudaStream_t s);
}  // namespace csr
