// __global__ wrapper around the warp programs of narrow_warp.h and the
// per-translation-unit lookup tables.
#pragma once
#include <cuda_runtime.h>

#include "narrow_warp.h"

#ifndef PINN_NARROW_BOUNDS
#define PINN_NARROW_BOUNDS __launch_bounds__(NW::kMaxWarps * 32, 1)
#endif

namespace pinn {

typedef void (*narrow_kernel_t)(const NarrowParams, int);

struct NarrowKernelInfo {
  narrow_kernel_t fn;
  int H, C, max_warps;
  int (*warp_floats)(int L);
  int (*stash_floats)(int L);
};

template <class NW>
__global__ void PINN_NARROW_BOUNDS narrow_kernel(const __grid_constant__ NarrowParams p, int Hreal) {
  extern __shared__ __align__(16) float smem[];
  const int L = p.n_hidden;
  const Layout lay(NW::H, L);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;
  const int wf = NW::warp_floats(L);
  float* wm = smem + lay.size() + warp * wf;
  float* gs = p.stash + ((size_t)blockIdx.x * nwarps + warp) * NW::stash_floats(L);

  float* g = p.gacc + ((size_t)blockIdx.x * nwarps + warp) * lay.size();  // this warp's gradient accumulators

  for (int i = tid; i < lay.size(); i += blockDim.x) smem[i] = stage_weight(p, Hreal, lay, i);
  if (p.mode != MODE_FWD)  // (the jets-only mode has no workspace)
    for (int i = lane; i < lay.size(); i += 32) g[i] = 0.f;
  __syncthreads();

  typename NW::Lane ln;
  ln.loss = 0.f;
  // tile -> (CTA, warp) CTA-fastest: a small domain (fewer tiles than warps on the GPU) spreads over all SMs with few
  // active warps each instead of filling a few SMs with 8 warps (a tile's latency is ~0.55x with 2-3 warps per SM)
  for (int64_t tile = (int64_t)warp * gridDim.x + blockIdx.x; tile < p.n_tiles; tile += (int64_t)gridDim.x * nwarps) {
    NW::phase_forward(p, smem, wm, gs, lane, tile, ln);
    if (p.mode == MODE_FWD || p.mode == MODE_LOSS) continue;
    for (int l = L + 1; l >= 2; --l) {
      NW::phase_backward(p, smem, wm, gs, lane, l, ln);
      __syncwarp();
      if (l == L + 1) NW::phase_gemm_out(p, wm, g, lane);
      else NW::phase_gemm_hidden(p, wm, g, lane, l);
      __syncwarp();
      NW::phase_store_zk(wm, lane, ln);
    }
    __syncwarp();
    NW::phase_gemm_first(p, wm, g, lane);
    __syncwarp();
  }
  if (p.mode == MODE_FWD) return;

  wm[NW::off_xs() + lane] = ln.loss;
  __syncthreads();  // also orders this block's accumulator writes (global) before the reads below
  // per-CTA partial = sum over the CTA's warps, in warp order (deterministic)
  float* gp = p.gpart + (size_t)blockIdx.x * lay.size();
  for (int i = tid; i < lay.size() && p.mode != MODE_LOSS; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += p.gacc[((size_t)blockIdx.x * nwarps + w) * lay.size() + i];
    gp[i] += s;
  }
  if ((p.mode == MODE_FUSED || p.mode == MODE_LOSS) && tid == 0) {
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) {
      float sw_ = 0.f;
      for (int l2 = 0; l2 < 32; ++l2) sw_ += smem[lay.size() + w * wf + NW::off_xs() + l2];
      s += sw_;
    }
    p.lpart[(size_t)p.domain_slot * p.max_cta + blockIdx.x] += s;
  }
}

// One table per (H, ACT) translation unit: entries for every (NF, NS), plus the extra-channel programs
// (x: NarrowWarp's X) for the plans that occur: mixed with (2, 0) / (2, 1) / (2, 2), third and third + fourth order
// with (1, 1).
#define PINN_NARROW_ENTRY(H, NF, NS, ACT, X) \
  { narrow_kernel<NarrowWarpFor<H, NF, NS, ACT, X>>, H, NarrowWarpFor<H, NF, NS, ACT, X>::C, NarrowWarpFor<H, NF, NS, ACT, X>::kMaxWarps, \
    &NarrowWarpFor<H, NF, NS, ACT, X>::warp_floats, \
    &NarrowWarpFor<H, NF, NS, ACT, X>::stash_floats }

#define PINN_NARROW_TABLE(NAME, H, ACT)                                          \
  const NarrowKernelInfo* NAME(int nf, int ns, int x) {                          \
    static const NarrowKernelInfo t[4][4] = {                                    \
        {PINN_NARROW_ENTRY(H, 0, 0, ACT, 0), {0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}}, \
        {PINN_NARROW_ENTRY(H, 1, 0, ACT, 0), PINN_NARROW_ENTRY(H, 1, 1, ACT, 0), {0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}}, \
        {PINN_NARROW_ENTRY(H, 2, 0, ACT, 0), PINN_NARROW_ENTRY(H, 2, 1, ACT, 0), PINN_NARROW_ENTRY(H, 2, 2, ACT, 0), {0, 0, 0, 0, 0, 0}}, \
        {PINN_NARROW_ENTRY(H, 3, 0, ACT, 0), PINN_NARROW_ENTRY(H, 3, 1, ACT, 0), PINN_NARROW_ENTRY(H, 3, 2, ACT, 0), PINN_NARROW_ENTRY(H, 3, 3, ACT, 0)}}; \
    static const NarrowKernelInfo mixed20 = PINN_NARROW_ENTRY(H, 2, 0, ACT, 1), mixed21 = PINN_NARROW_ENTRY(H, 2, 1, ACT, 1), \
                                  mixed22 = PINN_NARROW_ENTRY(H, 2, 2, ACT, 1);  \
    static const NarrowKernelInfo third = PINN_NARROW_ENTRY(H, 1, 1, ACT, 2), fourth = PINN_NARROW_ENTRY(H, 1, 1, ACT, 3); \
    if (nf < 0 || nf > 3 || ns < 0 || ns > nf) return nullptr;                   \
    if (x == 1) return nf != 2 ? nullptr : ns == 0 ? &mixed20 : ns == 1 ? &mixed21 : &mixed22; \
    if (x == 2 || x == 3) return (nf == 1 && ns == 1) ? (x == 2 ? &third : &fourth) : nullptr; \
    if (x != 0) return nullptr;                                                  \
    return t[nf][ns].fn ? &t[nf][ns] : nullptr;                                  \
  }

const NarrowKernelInfo* narrow_table_h20_tanh(int nf, int ns, int x);
const NarrowKernelInfo* narrow_table_h20_silu(int nf, int ns, int x);
const NarrowKernelInfo* narrow_table_h20_sin(int nf, int ns, int x);
const NarrowKernelInfo* narrow_table_h32_tanh(int nf, int ns, int x);
const NarrowKernelInfo* narrow_table_h32_silu(int nf, int ns, int x);
const NarrowKernelInfo* narrow_table_h32_sin(int nf, int ns, int x);

}  // namespace pinn
