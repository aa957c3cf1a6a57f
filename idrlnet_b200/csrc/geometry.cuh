// CSG sampling on the device (geometry.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pinn_b200.h"

namespace pinn {

int64_t csg_workspace_bytes(int64_t n_candidates);
int csg_sample(float* const* columns, int n_dims, const float* lo, const float* hi, int64_t n, int64_t capacity, uint64_t seed,
               uint64_t stream_id, uint64_t offset, const pinn_shape_op* prog, int n_ops, float* sdf, int64_t* count, void* ws,
               cudaStream_t st);

}  // namespace pinn
