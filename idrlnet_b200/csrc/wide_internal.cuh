// Pieces shared by the wide-path translation units (wide.cu, wide_tc.cu).
#pragma once
#include <cuda_runtime.h>

#include "pinn_common.h"

namespace pinn {
void count_launch(int n);

constexpr int kMaxC = 7;
constexpr int kBlocks = 148;         // partial slots of the per-block reductions
constexpr int kBM = 32;              // points per GEMM tile
constexpr int kBN = 128;             // output columns per GEMM tile
constexpr int kBK = 16;
constexpr int64_t kChunkBudget = 1ll << 30;   // bytes of Z/H planes per chunk (large chunks: many tiles in flight)
constexpr int kMaxDomains = 64;       // loss slots

struct WNet {
  int n_in, n_out, L, H, act, C, NF, NS;
  int precision;  // 0: FP32 FMA kernels (parity path); 1: tcgen05 TF32 GEMMs
  int first_in[4];
  const float* W[PINN_MAX_LAYERS];
  const float* b[PINN_MAX_LAYERS];
};

// ---- activation jets with runtime channel counts ---------------------------
template <int ACT>
__device__ __forceinline__ void jet_fwd(int NF, int NS, const float* z, float* h) {
  float s0, s1, s2, s3;
  act_derivs<ACT>(z[0], s0, s1, s2, s3);
  h[0] = s0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (i < NF) h[1 + i] = s1 * z[1 + i];
#pragma unroll
  for (int s = 0; s < 3; ++s)
    if (s < NS) h[1 + NF + s] = s2 * z[1 + s] * z[1 + s] + s1 * z[1 + NF + s];
}

template <int ACT>
__device__ __forceinline__ void jet_bwd(int NF, int NS, const float* z, const float* hb, float* zb) {
  float s0, s1, s2, s3;
  act_derivs<ACT>(z[0], s0, s1, s2, s3);
  float zv = hb[0] * s1;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (i < NF) {
      zv += hb[1 + i] * s2 * z[1 + i];
      zb[1 + i] = hb[1 + i] * s1;
    }
#pragma unroll
  for (int s = 0; s < 3; ++s)
    if (s < NS) {
      const float zi = z[1 + s], zii = z[1 + NF + s], hbs = hb[1 + NF + s];
      zv += hbs * (s3 * zi * zi + s2 * zii);
      zb[1 + s] += 2.0f * hbs * s2 * zi;
      zb[1 + NF + s] = hbs * s1;
    }
  zb[0] = zv;
}


// tensor-core GEMMs (wide_tc.cu)
template <int ACT, int EPI>
int tc_launch_rows(const WNet& n, const float* Bsrc, const float* Aw, const float* bias, float* Zio, float* Hout, int m, int Mc,
                   cudaStream_t st);
int tc_launch_dw(const float* Zb, const float* Hb, int H, int C, int m, int Mc, float* part, int n_splits, cudaStream_t st);
int tc_dw_splits(int H);
int tc_transpose(const float* W, float* WT, int H, cudaStream_t st);

}  // namespace pinn
