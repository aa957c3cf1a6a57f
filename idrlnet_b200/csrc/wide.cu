// Wide-net path: placeholder until the tiled kernels land (returns "unsupported"
// so callers take the generic torch path; never a silent CPU fallback).
#include "wide.cuh"
namespace pinn {
bool wide_supported(const pinn_mlp*, const pinn_jets*) { return false; }
int64_t wide_workspace_bytes(const pinn_mlp*, const pinn_jets*, int) { return 0; }
int wide_step_fused(const pinn_mlp*, const pinn_jets*, const pinn_domain*, void*, int64_t, int, int, cudaStream_t) { return PINN_E_UNSUPPORTED; }
int wide_step_finalize(const pinn_mlp*, const pinn_jets*, void*, int64_t, int, float*, int, float*, cudaStream_t) { return PINN_E_UNSUPPORTED; }
int wide_jet_forward(const pinn_mlp*, const pinn_jets*, int64_t, const float* const*, float*, cudaStream_t) { return PINN_E_UNSUPPORTED; }
int wide_jet_backward(const pinn_mlp*, const pinn_jets*, int64_t, const float* const*, const float*, void*, int64_t, int, cudaStream_t) { return PINN_E_UNSUPPORTED; }
}  // namespace pinn
