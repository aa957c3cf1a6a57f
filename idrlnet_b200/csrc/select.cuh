// Forward-only residuals and top-k selection (select.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pinn_b200.h"

namespace pinn {

constexpr int kTopkMax = 4096;  // survivors are sorted by one block in shared memory

// out[e * N + n] = residual of equation e at point n, from jets[(c * n_out + o) * N + n]
int residual_eval(const pinn_domain* dom, int n_out, const float* jets, float* out, cudaStream_t st);
int64_t topk_workspace_bytes(int k);
int topk_abs(const float* v, int64_t n, int k, int64_t* idx_out, float* val_out, void* ws, cudaStream_t st);

}  // namespace pinn
