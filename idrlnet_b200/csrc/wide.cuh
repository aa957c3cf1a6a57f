// Wide-net path (hidden width > 32): entry points used by abi.cu.
#pragma once
#include <cuda_runtime.h>

#include "pinn_common.h"

namespace pinn {
void count_launch(int n);
bool wide_supported(const pinn_mlp* m, const pinn_jets* j);
int64_t wide_workspace_bytes(const pinn_mlp* m, const pinn_jets* j, int n_domains);
int wide_step_fused(const pinn_mlp* m, const pinn_jets* j, const pinn_domain* dom, void* ws, int64_t ws_bytes,
                    int n_slots, int domain_slot, int accumulate, bool loss_only, cudaStream_t st);
int wide_step_finalize(const pinn_mlp* m, const pinn_jets* j, void* ws, int64_t ws_bytes, int n_domains,
                       float* flat_grad, int add_to_grad, float* losses, cudaStream_t st);
int wide_jet_forward(const pinn_mlp* m, const pinn_jets* j, int64_t n_points, const float* const* coords, float* out,
                     void* ws, int64_t ws_bytes, cudaStream_t st);
int wide_jet_backward(const pinn_mlp* m, const pinn_jets* j, int64_t n_points, const float* const* coords,
                      const float* out_bar, void* ws, int64_t ws_bytes, int accumulate, cudaStream_t st);
}  // namespace pinn
