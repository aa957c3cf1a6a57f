// Pieces shared by the wide-path translation units (wide.cu, wide_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "pinn_common.h"

namespace pinn {
void count_launch(int n);

constexpr int kMaxC = 7;
constexpr int kBlocks = 148;         // partial slots of the per-block reductions
constexpr int kBM = 32;              // points per GEMM tile
constexpr int kBN = 128;             // output columns per GEMM tile
constexpr int kBK = 16;
int64_t chunk_budget();              // bytes of Z/H planes (+ re-tiled operands) per chunk (wide.cu)
constexpr int kMaxDomains = 64;       // loss slots

struct WNet {
  int n_in, n_out, L, H, act, C, NF, NS;
  int precision;  // 0: FP32 FMA kernels (parity path); 1: tcgen05 TF32 GEMMs
  int first_in[4];
  const float* W[PINN_MAX_LAYERS];
  const float* b[PINN_MAX_LAYERS];
};

// ---- activation jets with runtime channel counts ---------------------------
// All array indices are compile-time constants (selection by predicate chains), so the jet
// arrays stay in registers even though NF / NS are only known at run time.
__device__ __forceinline__ float pick7(const float* a, int idx) {
  float r = a[0];
#pragma unroll
  for (int c = 1; c < 7; ++c) r = (idx == c) ? a[c] : r;
  return r;
}

template <int ACT>
__device__ __forceinline__ void jet_fwd(int NF, int NS, const float* z, float* h) {
  float s0, s1, s2, s3;
  act_derivs<ACT>(z[0], s0, s1, s2, s3);
  h[0] = s0;
#pragma unroll
  for (int c = 1; c < 7; ++c) {
    if (c <= NF) {
      h[c] = s1 * z[c];
    } else if (c <= NF + NS) {
      const float zi = pick7(z, c - NF);  // first-order channel of the same input
      h[c] = s2 * zi * zi + s1 * z[c];
    }
  }
}

template <int ACT>
__device__ __forceinline__ void jet_bwd(int NF, int NS, const float* z, const float* hb, float* zb) {
  float s0, s1, s2, s3;
  act_derivs<ACT>(z[0], s0, s1, s2, s3);
  float zv = hb[0] * s1;
#pragma unroll
  for (int c = 1; c < 7; ++c) {
    if (c <= NF) {
      float t = hb[c] * s1;
      zv += hb[c] * s2 * z[c];
      if (c <= NS) {  // this input also has a second-order channel at index c + NF
        const float hbs = pick7(hb, c + NF), zii = pick7(z, c + NF);
        zv += hbs * (s3 * z[c] * z[c] + s2 * zii);
        t += 2.0f * hbs * s2 * z[c];
      }
      zb[c] = t;
    } else if (c <= NF + NS) {
      zb[c] = hb[c] * s1;
    }
  }
  zb[0] = zv;
}

// compile-time channel structure (no dynamically indexed jet arrays -> everything stays in registers)
template <int ACT, int NF, int NS, bool FAST>
__device__ __forceinline__ void jet_fwd_t(const float (&z)[1 + NF + NS], float (&h)[1 + NF + NS]) {
  float s0, s1, s2, s3;
  if (FAST) act_derivs_fast<ACT>(z[0], s0, s1, s2, s3);
  else act_derivs<ACT>(z[0], s0, s1, s2, s3);
  h[0] = s0;
#pragma unroll
  for (int i = 0; i < NF; ++i) h[1 + i] = s1 * z[1 + i];
#pragma unroll
  for (int s = 0; s < NS; ++s) h[1 + NF + s] = s2 * z[1 + s] * z[1 + s] + s1 * z[1 + NF + s];
}

template <int ACT, int NF, int NS, bool FAST>
__device__ __forceinline__ void jet_bwd_t(const float (&z)[1 + NF + NS], const float (&hb)[1 + NF + NS],
                                          float (&zb)[1 + NF + NS]) {
  float s0, s1, s2, s3;
  if (FAST) act_derivs_fast<ACT>(z[0], s0, s1, s2, s3);
  else act_derivs<ACT>(z[0], s0, s1, s2, s3);
  float zv = hb[0] * s1;
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    zv += hb[1 + i] * s2 * z[1 + i];
    zb[1 + i] = hb[1 + i] * s1;
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const float zi = z[1 + s], zii = z[1 + NF + s], hbs = hb[1 + NF + s];
    zv += hbs * (s3 * zi * zi + s2 * zii);
    zb[1 + s] += 2.0f * hbs * s2 * zi;
    zb[1 + NF + s] = hbs * s1;
  }
  zb[0] = zv;
}

// residual polynomials, loss and its seed adjoint for one point (runtime descriptor, C channels)
template <int C>
__device__ __forceinline__ void residual_loss_rt(const pinn_domain& dm, int64_t n, const float* x, const float (*y)[4],
                                                 float (*yb)[4], float& loss_acc) {
  float sym[PINN_SYM_COUNT];
  float sg[PINN_SYM_COORD0];
  for (int i = 0; i < PINN_SYM_COUNT; ++i) sym[i] = 0.f;
  for (int i = 0; i < PINN_SYM_COORD0; ++i) sg[i] = 0.f;
  for (int c = 0; c < C; ++c)
    for (int o = 0; o < 4; ++o) sym[c * 4 + o] = y[c][o];
  for (int d = 0; d < 4; ++d) sym[PINN_SYM_COORD0 + d] = x[d];
  for (int a = 0; a < dm.n_aux; ++a) sym[PINN_SYM_AUX0 + a] = dm.aux[a][n];
  const float area = dm.area ? dm.area[n] : dm.area_const;
  float loss = 0.f;
  for (int e = 0; e < dm.n_eq; ++e) {
    const pinn_equation& q = dm.eq[e];
    float r = 0.f;
    for (int t = q.term_begin; t < q.term_end; ++t) {
      const pinn_term& tm = dm.terms[t];
      float prod = tm.coef;
      for (int f = 0; f < tm.n_fac; ++f) prod *= sym[tm.fac[f]];
      r += prod;
    }
    const float tgt = q.target ? q.target[n] : q.target_const;
    const float lam = q.lambda ? q.lambda[n] : q.lambda_const;
    const float diff = r - tgt, w = lam * area;
    float f0, f1;
    if (dm.loss_kind == PINN_LOSS_SQUARE) { f0 = diff * diff; f1 = 2.0f * diff; }
    else if (dm.loss_kind == PINN_LOSS_L1) { f0 = fabsf(diff); f1 = (diff > 0.f) ? 1.0f : ((diff < 0.f) ? -1.0f : 0.f); }
    else { f0 = diff; f1 = 1.0f; }
    loss += f0 * w;
    const float g = f1 * w * dm.sigma;
    for (int t = q.term_begin; t < q.term_end; ++t) {
      const pinn_term& tm = dm.terms[t];
      for (int f = 0; f < tm.n_fac; ++f) {
        const int id = tm.fac[f];
        if (id >= PINN_SYM_COORD0) continue;
        float prod = tm.coef * g;
        for (int f2 = 0; f2 < tm.n_fac; ++f2)
          if (f2 != f) prod *= sym[tm.fac[f2]];
        sg[id] += prod;
      }
    }
  }
  loss_acc += loss;
  for (int c = 0; c < C; ++c)
    for (int o = 0; o < 4; ++o) yb[c][o] = sg[c * 4 + o];
}

// whole-step-on-chip kernel for H <= 128 (wide_fused.cu)
bool fused_wide_supported(const WNet& n);
int64_t fused_wide_stash_floats(const WNet& n);
int64_t fused_wide_wprep_floats(const WNet& n);
int fused_wide_prep(const WNet& n, float* wprep, cudaStream_t st);
template <int ACT>
int fused_wide_launch(const WNet& n, const pinn_domain* dom, float* stash, const float* wt_base, float* dw_part, float* db_part,
                      float* first_part, float* head_part, int n_splits, bool loss_only, cudaStream_t st);

// tensor-core GEMMs (wide_tc.cu)
template <int ACT, int EPI>
int tc_launch_rows(const WNet& n, const float* Bsrc, const float* Aw, const float* bias, float* Zio, float* Hout, int m, int Mc,
                   bool round_out, float* colsum, cudaStream_t st);  // colsum (reverse, TF32 mode): [ceil(m/16)][H] sums of zbar's value channel
bool tc_encode_tiled(CUtensorMap* tm, int rank, const void* base, const cuuint64_t* gdim, const cuuint64_t* gstr, const cuuint32_t* box,
                     const cuuint32_t* est, int swizzle);  // 0 none, 1 SWIZZLE_128B, 2 SWIZZLE_128B_ATOM_32B
bool tc_dw_direct();  // tensor modes: dW operands straight from the planes (tc_launch_dw3) instead of re-tiled copies
int tc_launch_dw3(const float* Zb, const float* Hb, int H, int C, int m, int Mc, float* part, int n_splits, bool x3, cudaStream_t st);
int tc_colsum_finish(const float* colsum, int nrows, int H, float* db_slot, cudaStream_t st);
int tc_dw_splits(int H);
int tc_rows_points(int C);  // points per tile of the row GEMMs
int64_t tc_dw2_scratch_floats(int H, int C, int Mc, bool x3);
int tc_launch_dw2(const float* Zb, const float* Hb, int H, int C, int m, int Mc, float* part, int n_splits, float* scratch,
                  float* db_slot, bool x3, cudaStream_t st);
int tc_transpose(const float* W, float* WT, int H, bool round, cudaStream_t st);
int tc_round_copy(const float* W, float* Wr, int n, cudaStream_t st);  // Wr = tf32(W)

}  // namespace pinn
