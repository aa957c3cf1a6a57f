// Wide-net path (hidden width > 32; configs C/D/E: 128, 100, 512), FP32.
//
// Same mathematics as the narrow path (forward Taylor jets, fused residual/loss,
// reverse pass), organised as layer-streamed kernels over chunks
// of points, because at these widths a layer's work IS a dense contraction
// (rows = points x jet channels, K = N = H):
//
//   first_fwd                Z1 = W1 x + b1 (jets), H1 = act(Z1)
//   gemm_fwd   (l = 2..L)    Z_l = H_{l-1} W_l^T + b_l ; H_l = act(Z_l)   (act fused in the epilogue)
//   head                     y = H_L Wout^T + bout ; residual ; loss ; ybar ;
//                            zbar_L = act'(Z_L; Wout^T ybar) ; dWout, dbout
//   for l = L..2:
//     gemm_dw                dW_l  = zbar_l^T H_{l-1}            (split over points, per-split partials)
//     reduce_cols            db_l  = column sums of zbar_l (value channel)
//     gemm_hbar              zbar_{l-1} = act'(Z_{l-1}; zbar_l W_l)       (act' fused in the epilogue, in place)
//   reduce_cols              dW1, db1 from zbar_1 and the coordinates
//
// Jets live in a chunk workspace as channel planes [C][Mc][H] (Z_l, H_l for every
// layer of the chunk; Z_l is overwritten by zbar_l).  All reductions go through
// per-block / per-split partial slots that are summed in a fixed order by
// wide_finalize -> deterministic, no atomics.
//
// This FP32-FMA implementation is the parity path (1e-5); it keeps the tensor
// pipe idle.  See DESIGN.md for the tcgen05 plan that replaces the three GEMMs.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include "wide.cuh"
#include "wide_internal.cuh"

namespace pinn {

namespace {

#define ACT_SWITCH(act, CALL)                      \
  do {                                             \
    if ((act) == PINN_ACT_TANH) { CALL(PINN_ACT_TANH); } \
    else if ((act) == PINN_ACT_SILU) { CALL(PINN_ACT_SILU); } \
    else { CALL(PINN_ACT_SIN); }                   \
  } while (0)

// ---- first layer ------------------------------------------------------------
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

struct Coords {
  const float* p[PINN_MAX_IN];
};

constexpr int kFirstPPT = 8;  // points per thread of first_fwd_kernel
// thread = (group of kFirstPPT consecutive points, unit j): the unit's first-layer row and the derivative channels of Z_1 (the
// weights themselves) stay in registers over the group; compile-time channel structure (the runtime-indexed version kept its jets
// in local memory: 2.1 TB/s of plane stores)
template <int ACT, int NF, int NS>
__global__ void __launch_bounds__(256) first_fwd_kernel(WNet net, Coords xs, int64_t n0, int m, int Mc, float* __restrict__ Z,
                                                        float* __restrict__ Hb) {
  constexpr int C = 1 + NF + NS;
  const int H = net.H;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int groups = (m + kFirstPPT - 1) / kFirstPPT;
  if (idx >= (int64_t)groups * H) return;
  const int g = (int)(idx / H), j = (int)(idx - (int64_t)g * H);
  float w[PINN_MAX_IN];
#pragma unroll
  for (int d = 0; d < PINN_MAX_IN; ++d) w[d] = d < net.n_in ? net.W[0][j * net.n_in + d] : 0.f;
  float z[C], h[C];
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    const int fi = net.first_in[i];
    float wf = w[0];
#pragma unroll
    for (int d = 1; d < PINN_MAX_IN; ++d) wf = (fi == d) ? w[d] : wf;
    z[1 + i] = wf;
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) z[1 + NF + s] = 0.f;
  const float bj = net.b[0][j];
  const size_t plane = (size_t)Mc * H;
  const bool round = net.precision == 1 && net.L >= 2;  // TF32 mode: H_1 is a tensor-core operand, stored rounded
  const int nb = g * kFirstPPT, ne = (nb + kFirstPPT < m) ? nb + kFirstPPT : m;
  for (int n = nb; n < ne; ++n) {
    float v = bj;
#pragma unroll
    for (int d = 0; d < PINN_MAX_IN; ++d)
      if (d < net.n_in) v += w[d] * xs.p[d][n0 + n];
    z[0] = v;
    jet_fwd_t<ACT, NF, NS, false>(z, h);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      Z[c * plane + (size_t)n * H + j] = z[c];
      Hb[c * plane + (size_t)n * H + j] = round ? tf32_rna(h[c]) : h[c];
    }
  }
}

// ---- GEMM with per-point channel epilogue -----------------------------------
// out[(c,n)][j] = sum_k A[(c,n)][k] * Bop(k, j):   BT = true : B is W[j][k]  (forward, out = A W^T)
//                                                   BT = false: B is W[k][j]  (reverse, out = A W)
// tile: 32 points x C channels rows, 128 columns; 256 threads = 16 (cols) x 16 (points);
// thread tile: C channels x 2 points x 8 columns.
// EPI = 0: Zout = acc + bias ; Hout = act(Zout)                 (forward layer)
// EPI = 1: Zio  = act'(Zio; acc)   (in place: reads z, writes zbar)   (reverse layer)
template <int C, int ACT, bool BT, int EPI>
__global__ void __launch_bounds__(256) gemm_rows_kernel(WNet net, const float* __restrict__ A, const float* __restrict__ Wl,
                                                        const float* __restrict__ bias, float* __restrict__ Zio,
                                                        float* __restrict__ Hout, int m, int Mc) {
  const int H = net.H;
  __shared__ float As[kBK][C * kBM + 4];
  __shared__ __align__(16) float Bs[kBK][kBN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * kBM, j0 = blockIdx.y * kBN;
  const size_t plane = (size_t)Mc * H;
  float acc[C][2][8];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[c][r][t] = 0.f;

  for (int k0 = 0; k0 < H; k0 += kBK) {
    // A tile: rows (c, p), 16 k's each, loaded as float4 along k and stored transposed
    for (int i = tid; i < C * kBM * 4; i += 256) {
      const int row = i >> 2, kq = (i & 3) * 4;
      const int c = row / kBM, p = row % kBM;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + p < m && k0 + kq < H) v = *reinterpret_cast<const float4*>(A + c * plane + (size_t)(n0 + p) * H + k0 + kq);
      As[kq + 0][row] = v.x; As[kq + 1][row] = v.y; As[kq + 2][row] = v.z; As[kq + 3][row] = v.w;
    }
    if (BT) {  // B tile from W[j][k]: 128 j rows x 16 k, transposed into Bs[k][j]
      for (int i = tid; i < kBN * 4; i += 256) {
        const int j = i >> 2, kq = (i & 3) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j0 + j < H && k0 + kq < H) v = *reinterpret_cast<const float4*>(Wl + (size_t)(j0 + j) * H + k0 + kq);
        Bs[kq + 0][j] = v.x; Bs[kq + 1][j] = v.y; Bs[kq + 2][j] = v.z; Bs[kq + 3][j] = v.w;
      }
    } else {  // B tile from W[k][j]: 16 k rows x 128 j, direct
      for (int i = tid; i < kBK * (kBN / 4); i += 256) {
        const int kk = i / (kBN / 4), jq = (i % (kBN / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + kk < H && j0 + jq < H) v = *reinterpret_cast<const float4*>(Wl + (size_t)(k0 + kk) * H + j0 + jq);
        *reinterpret_cast<float4*>(&Bs[kk][jq]) = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float2 a = *reinterpret_cast<const float2*>(&As[kk][c * kBM + ty * 2]);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          acc[c][0][t] += a.x * bv[t];
          acc[c][1][t] += a.y * bv[t];
        }
      }
    }
    __syncthreads();
  }
  // epilogue: all C channels of a (point, column) are in this thread
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int n = n0 + ty * 2 + r;
    if (n >= m) continue;
#pragma unroll
    for (int tq = 0; tq < 2; ++tq) {
      const int j = j0 + tx * 8 + tq * 4;
      if (j >= H) continue;
      float4 outz[C], outh[C];
      float4 zin[C];
      if (EPI == 1) {
#pragma unroll
        for (int c = 0; c < C; ++c) zin[c] = *reinterpret_cast<const float4*>(Zio + c * plane + (size_t)n * H + j);
      }
      float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
      if (EPI == 0) bj = *reinterpret_cast<const float4*>(bias + j);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v[kMaxC] = {}, w[kMaxC] = {}, o[kMaxC] = {};
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = acc[c][r][tq * 4 + e];
        if (EPI == 0) {
          v[0] += (&bj.x)[e];
          jet_fwd<ACT>(net.NF, net.NS, v, o);
#pragma unroll
          for (int c = 0; c < C; ++c) {
            (&outz[c].x)[e] = v[c];
            (&outh[c].x)[e] = o[c];
          }
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            w[c] = (&zin[c].x)[e];
            o[c] = 0.f;
          }
          jet_bwd<ACT>(net.NF, net.NS, w, v, o);
#pragma unroll
          for (int c = 0; c < C; ++c) (&outz[c].x)[e] = o[c];
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        *reinterpret_cast<float4*>(Zio + c * plane + (size_t)n * H + j) = outz[c];
        if (EPI == 0) *reinterpret_cast<float4*>(Hout + c * plane + (size_t)n * H + j) = outh[c];
      }
    }
  }
}

// ---- dW_l[j][k] (+)= sum over rows (c,n) of zbar[(c,n)][j] * h[(c,n)][k] ----------
// grid (tiles_j * tiles_k, n_splits); 128 x 128 tile, thread tile 8 x 8.
__global__ void __launch_bounds__(256) gemm_dw_kernel(const float* __restrict__ Zb, const float* __restrict__ Hb, int H, int C,
                                                      int m, int Mc, float* __restrict__ part, int tiles_k) {
  __shared__ __align__(16) float As[kBK][kBN + 4];
  __shared__ __align__(16) float Bs[kBK][kBN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int j0 = (blockIdx.x / tiles_k) * kBN, k0 = (blockIdx.x % tiles_k) * kBN;
  const int split = blockIdx.y, n_splits = gridDim.y;
  const size_t plane = (size_t)Mc * H;
  const int64_t rows = (int64_t)C * m;
  int64_t per = (rows + n_splits - 1) / n_splits;
  per = (per + kBK - 1) / kBK * kBK;
  const int64_t r_begin = split * per, r_end = (r_begin + per < rows) ? r_begin + per : rows;
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += kBK) {
    for (int i = tid; i < kBK * (kBN / 4); i += 256) {
      const int kk = i / (kBN / 4), q = (i % (kBN / 4)) * 4;
      const int64_t r = r0 + kk;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (r < r_end) {
        const int c = (int)(r / m), n = (int)(r % m);
        const size_t base = c * plane + (size_t)n * H;
        if (j0 + q < H) va = *reinterpret_cast<const float4*>(Zb + base + j0 + q);
        if (k0 + q < H) vb = *reinterpret_cast<const float4*>(Hb + base + k0 + q);
      }
      *reinterpret_cast<float4*>(&As[kk][q]) = va;
      *reinterpret_cast<float4*>(&Bs[kk][q]) = vb;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] += av[a] * bv[b];
    }
    __syncthreads();
  }
  float* out = part + (size_t)split * H * H;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int j = j0 + ty * 8 + a;
    if (j >= H) continue;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int k = k0 + tx * 8 + b;
      if (k < H) out[(size_t)j * H + k] += acc[a][b];
    }
  }
}

// ---- column reductions: part[block][slot_q][j] += sum_n src[plane_q][n][j] * (w_q ? w_q[n0+n] : 1)
// (several terms may target the same slot, so the slot layout does not depend on the jet plan)
struct RedSpec {
  int nq;
  int plane[8];
  int slot[8];
  const float* w[8];
};
// grid (ceil(H/32), kBlocks); block = 32 columns x 8 row groups; blockIdx.y is the partial slot.
__global__ void __launch_bounds__(256) reduce_cols_kernel(const float* __restrict__ src, RedSpec rs, int H, int64_t n0, int m,
                                                           int Mc, float* __restrict__ part) {
  __shared__ float red[8][8][33];
  const size_t plane = (size_t)Mc * H;
  const int per = (m + gridDim.y - 1) / gridDim.y;
  const int nb = blockIdx.y * per, ne = (nb + per < m) ? nb + per : m;
  const int jl = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + jl;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (j < H) {
    for (int n = nb + rg; n < ne; n += 32) {  // 4 rows in flight per thread
      float v[4][8], w[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int nn = n + 8 * u;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          v[u][q] = 0.f;
          w[u][q] = 1.f;
          if (q < rs.nq && nn < ne) {
            v[u][q] = src[rs.plane[q] * plane + (size_t)nn * H + j];
            if (rs.w[q]) w[u][q] = rs.w[q][n0 + nn];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] += v[u][q] * w[u][q];
    }
  }
  float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (q < rs.nq) {
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (rs.slot[q] == t) o[t] += a[q];
    }
#pragma unroll
  for (int t = 0; t < 8; ++t) red[rg][t][jl] = o[t];
  __syncthreads();
  if (rg == 0 && j < H) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float sum = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) sum += red[g][t][jl];
      if (sum != 0.f) part[((size_t)blockIdx.y * 8 + t) * H + j] += sum;
    }
  }
}

// ---- head: output layer, residual, loss, seeds of the reverse pass --------------------
struct HeadArgs {
  int mode;  // MODE_*
  pinn_domain dom;
  Coords xs;
  int64_t n0, n_total;
  float* jets_out;
  const float* jets_bar;
  float* part;  // [block][O*H + O + 1]  (dWout, dbout, loss)
};
enum { W_MODE_FUSED = 0, W_MODE_FWD = 1, W_MODE_BWD = 2, W_MODE_LOSS = 3 };  // LOSS: FUSED without the reverse pass

// out_layer: y[n][c][o] = sum_k Wout[o][k] h_L[c][n][k] + bout[o]*(c==0), FP32 (the output values feed
// residuals with cancellations, so this small skinny product stays off the TF32 pipe).  One warp
// handles 4 rows (c, n) per iteration (16 loads in flight per lane), lanes stride over k.
template <int KI>
__global__ void __launch_bounds__(256) out_layer_kernel(WNet net, const float* __restrict__ HL, float* __restrict__ Yv, int m, int Mc) {
  extern __shared__ float sW[];  // Wout[O][H]
  const int H = net.H, O = net.n_out, C = net.C;
  for (int i = threadIdx.x; i < O * H; i += blockDim.x) sW[i] = net.W[net.L][i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const size_t plane = (size_t)Mc * H;
  const int rows = C * m;
  float w[4][KI];
#pragma unroll
  for (int o = 0; o < 4; ++o)
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const int k = lane + 32 * i;
      w[o][i] = (o < O && k < H) ? sW[o * H + k] : 0.f;
    }
  for (int r0 = (blockIdx.x * nwarps + warp) * 4; r0 < rows; r0 += gridDim.x * nwarps * 4) {
    float h[4][KI];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u;
      const int c = r / m, n = r - c * m;
#pragma unroll
      for (int i = 0; i < KI; ++i) {
        const int k = lane + 32 * i;
        h[u][i] = (r < rows && k < H) ? HL[c * plane + (size_t)n * H + k] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < KI; ++i)
#pragma unroll
        for (int o = 0; o < 4; ++o) s[o] += w[o][i] * h[u][i];
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s[o] += __shfl_xor_sync(0xffffffffu, s[o], d);
      const int r = r0 + u;
      if (lane == 0 && r < rows) {
        const int c = r / m, n = r - c * m;
        float4 out = make_float4(s[0], s[1], s[2], s[3]);
        if (c == 0) {
          out.x += net.b[net.L][0];
          if (O > 1) out.y += net.b[net.L][1];
          if (O > 2) out.z += net.b[net.L][2];
          if (O > 3) out.w += net.b[net.L][3];
        }
        *reinterpret_cast<float4*>(Yv + ((size_t)n * C + c) * 4) = out;
      }
    }
  }
}

// head_point: one THREAD per point: residual / loss / ybar (in place in Yv) from the output-layer
// values y[n][c][o]; dbout and loss to the block's partial slot.
template <int C>
__global__ void __launch_bounds__(256) head_point_kernel(WNet net, const __grid_constant__ HeadArgs ha, float* __restrict__ Yv, int m) {
  __shared__ float red[8][5];
  const int O = net.n_out, H = net.H;
  float gb[4] = {0.f, 0.f, 0.f, 0.f};
  float loss = 0.f;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < m; n += gridDim.x * blockDim.x) {
    const int64_t ng = ha.n0 + n;
    float y[C][4], yb[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 t = *reinterpret_cast<const float4*>(Yv + ((size_t)n * C + c) * 4);
      y[c][0] = t.x; y[c][1] = t.y; y[c][2] = t.z; y[c][3] = t.w;
#pragma unroll
      for (int o = 0; o < 4; ++o)
        if (o >= O) y[c][o] = 0.f;
    }
    if (ha.mode == W_MODE_FWD) {
      for (int c = 0; c < C; ++c)
        for (int o = 0; o < O; ++o) ha.jets_out[(int64_t)(c * O + o) * ha.n_total + ng] = y[c][o];
      continue;
    }
    if (ha.mode == W_MODE_FUSED) {
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      for (int d = 0; d < net.n_in; ++d) x[d] = ha.xs.p[d][ng];
      residual_loss_rt<C>(ha.dom, ng, x, y, yb, loss);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int o = 0; o < 4; ++o) yb[c][o] = (o < O) ? ha.jets_bar[(int64_t)(c * O + o) * ha.n_total + ng] : 0.f;
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) gb[o] += yb[0][o];
#pragma unroll
    for (int c = 0; c < C; ++c)
      *reinterpret_cast<float4*>(Yv + ((size_t)n * C + c) * 4) = make_float4(yb[c][0], yb[c][1], yb[c][2], yb[c][3]);
  }
  if (ha.mode == W_MODE_FWD) return;
  float vals[5] = {gb[0], gb[1], gb[2], gb[3], loss};
#pragma unroll
  for (int q = 0; q < 5; ++q) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) vals[q] += __shfl_xor_sync(0xffffffffu, vals[q], d);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int q = 0; q < 5; ++q) red[warp][q] = vals[q];
  __syncthreads();
  const int hstride = O * H + O + 1;
  if (threadIdx.x <= O) {
    const int q = threadIdx.x < O ? threadIdx.x : 4;
    float sum = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += red[w][q];
    ha.part[(size_t)blockIdx.x * hstride + O * H + threadIdx.x] += sum;
  }
}

// head_bwd: streaming over (point, unit): zbar_L = act'(Z_L; Wout^T ybar) in place, dWout += ybar^T h_L.
// grid (ceil(H/128), kBlocks); block = 128 units x 4 point lanes; blockIdx.y is the partial slot.
template <int ACT, int NF, int NS>
__global__ void __launch_bounds__(1024) head_bwd_kernel(WNet net, const float* __restrict__ Yb, float* __restrict__ ZL,
                                                        const float* __restrict__ HL, int m, int Mc, float* __restrict__ part,
                                                        float* __restrict__ dbp) {
  constexpr int C = 1 + NF + NS;  // compile-time channel structure: the jets stay in registers
  __shared__ float red[8][5][129];
  const int H = net.H, O = net.n_out;
  const int jl = threadIdx.x & 127, pl = threadIdx.x >> 7;
  const int j = blockIdx.x * 128 + jl;
  const size_t plane = (size_t)Mc * H;
  const int per = (m + gridDim.y - 1) / gridDim.y;
  const int nb = blockIdx.y * per, ne = (nb + per < m) ? nb + per : m;
  float gW[4] = {0.f, 0.f, 0.f, 0.f};
  float dbs = 0.f;  // db_L = sum of zbar_L's (unrounded) value channel, when the caller asks for it (dbp)
  const bool round = net.precision == 1 && net.L >= 2;  // TF32 mode: zbar_L is a tensor-core operand, stored rounded
  if (j < H) {
    float wo[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) wo[o] = (o < O) ? net.W[net.L][o * H + j] : 0.f;
    for (int n = nb + pl; n < ne; n += 16) {  // two points (n, n + 8) in flight
      float z[2][C], h[2][C];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int nn = n + 8 * u;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const bool ok = nn < ne;
          z[u][c] = ok ? ZL[c * plane + (size_t)nn * H + j] : 0.f;
          h[u][c] = ok ? HL[c * plane + (size_t)nn * H + j] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int nn = n + 8 * u;
        float hb[C], zb[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          // ybar of the point: the same 16 bytes for the 128 units of the point lane (a broadcast hit in L1)
          const float4 y4 = nn < ne ? *reinterpret_cast<const float4*>(Yb + ((size_t)nn * C + c) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          hb[c] = wo[0] * y4.x + wo[1] * y4.y + wo[2] * y4.z + wo[3] * y4.w;
          gW[0] += y4.x * h[u][c]; gW[1] += y4.y * h[u][c]; gW[2] += y4.z * h[u][c]; gW[3] += y4.w * h[u][c];
        }
        jet_bwd_t<ACT, NF, NS, false>(z[u], hb, zb);
        if (nn < ne) {
          dbs += zb[0];
#pragma unroll
          for (int c = 0; c < C; ++c) ZL[c * plane + (size_t)nn * H + j] = round ? tf32_rna(zb[c]) : zb[c];
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) red[pl][o][jl] = gW[o];
  red[pl][4][jl] = dbs;
  __syncthreads();
  if (pl == 0 && j < H) {
    const int hstride = O * H + O + 1;
    for (int o = 0; o < O; ++o) {
      float sum = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) sum += red[g][o][jl];
      part[(size_t)blockIdx.y * hstride + o * H + j] += sum;
    }
    if (dbp != nullptr) {  // this block's slot of the [kBlocks][8][H] partials, row 0
      float sum = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) sum += red[g][4][jl];
      dbp[(size_t)blockIdx.y * 8 * H + j] += sum;
    }
  }
}

// ---- finalize: partial slots -> flat gradient (pinn_param_count order) and losses -----------
struct WideLayout {  // workspace offsets in floats
  int64_t Z, Hb;            // L planes-sets each: [l][C][Mc][H]
  int64_t dw, db, first, head, loss_slots, acc_end, total;
  int64_t wt, wprep, wr, ybar, tdw;  // tdw: re-tiled operands of the pipelined TF32 dW GEMM; wr: tf32(W_l) (TF32 mode)
  int Mc, n_splits, fp32_splits, tiles, n_dom;
};

__global__ void wide_finalize_kernel(WNet net, const float* __restrict__ ws, WideLayout wl, int64_t n_params,
                                     float* __restrict__ flat, int add, float* __restrict__ losses, int n_dom) {
  const int H = net.H, L = net.L, O = net.n_out, D = net.n_in;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_params) {
    int64_t r = i;
    float s = 0.f;
    const float* fp = ws + wl.first;
    const int hstride = O * H + O + 1;
    if (r < (int64_t)H * D) {  // W1[j][d]
      const int j = (int)(r / D), d = (int)(r % D);
      for (int b = 0; b < kBlocks; ++b) s += fp[((size_t)b * 8 + d) * H + j];
    } else if ((r -= (int64_t)H * D) < H) {  // b1 (slot PINN_MAX_IN)
      for (int b = 0; b < kBlocks; ++b) s += fp[((size_t)b * 8 + PINN_MAX_IN) * H + r];
    } else {
      r -= H;
      bool done = false;
      for (int l = 2; l <= L && !done; ++l) {
        if (r < (int64_t)H * H) {
          const float* p = ws + wl.dw + (size_t)(l - 2) * wl.n_splits * H * H;
          for (int sp = 0; sp < wl.n_splits; ++sp) s += p[(size_t)sp * H * H + r];
          done = true;
        } else if ((r -= (int64_t)H * H) < H) {
          const float* p = ws + wl.db + (size_t)(l - 2) * kBlocks * 8 * H;
          for (int b = 0; b < kBlocks; ++b) s += p[(size_t)b * 8 * H + r];
          done = true;
        } else {
          r -= H;
        }
      }
      if (!done) {  // Wout[o][k], bout[o]
        const float* p = ws + wl.head;
        for (int b = 0; b < kBlocks; ++b) s += p[(size_t)b * hstride + r];
      }
    }
    flat[i] = add ? flat[i] + s : s;
  }
  if (blockIdx.x == 0 && threadIdx.x < n_dom && losses) {
    const int hstride = O * H + O + 1;
    float s = 0.f;
    for (int b = 0; b < kBlocks; ++b) s += ws[wl.loss_slots + (size_t)threadIdx.x * kBlocks + b];
    (void)hstride;
    losses[threadIdx.x] = s;
  }
}

// move this step's head-loss (slot O*H+O of every block partial) into the per-domain loss slots
__global__ void wide_collect_loss_kernel(float* ws, WideLayout wl, int hstride, int slot) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < kBlocks) {
    float* p = ws + wl.head + (size_t)b * hstride + (hstride - 1);
    ws[wl.loss_slots + (size_t)slot * kBlocks + b] += *p;
    *p = 0.f;
  }
}

// ---- host side -------------------------------------------------------------------
bool make_net(const pinn_mlp* m, const pinn_jets* j, WNet* n) {
  if (m->width <= 32 || m->width > 512 || (m->width % 4) != 0) return false;
  if (m->n_hidden < 1 || j->n_first > 3 || j->n_second > j->n_first) return false;
  if (j->n_mixed || j->n_high) return false;  // mixed / third / fourth-order channels: narrow kernels only
  memset(n, 0, sizeof(*n));
  n->n_in = m->n_in; n->n_out = m->n_out; n->L = m->n_hidden; n->H = m->width; n->act = m->act;
  n->NF = j->n_first; n->NS = j->n_second; n->C = 1 + n->NF + n->NS;
  n->precision = (m->precision == PINN_PREC_TF32 || m->precision == PINN_PREC_TF32_STREAMED) ? 1
                 : (m->precision == PINN_PREC_FP32X3) ? 2 : 0;  // 0: FP32 SIMT, 1: TF32, 2: 3xTF32 (tensor cores, FP32-grade)
  for (int i = 0; i < 4; ++i) n->first_in[i] = j->first_in[i];
  for (int l = 0; l <= m->n_hidden; ++l) { n->W[l] = m->W[l]; n->b[l] = m->b[l]; }
  return true;
}

}  // namespace

// Bytes of planes per chunk of the layer-streamed path: 1/24 of the device memory, between 1 and 8 GB (7.5 GB on a 180 GB
// B200; PINN_WIDE_CHUNK_MB overrides).  It was a fixed 1 GB: 4 736 points per chunk of mlp_xl = four tiles per CTA of the
// persistent row GEMMs, whose fill and drain (one mainloop + one epilogue not overlapped) were then a quarter of every launch,
// and 7 800 launches per 2^20-point step.  Evaluated once: the workspace size and every step must agree on it.
int64_t chunk_budget() {
  static const int64_t v = [] {
    if (const char* e = getenv("PINN_WIDE_CHUNK_MB")) {
      const long long mb = atoll(e);
      if (mb >= 64) return (int64_t)mb << 20;
    }
    size_t free_b = 0, total_b = 0;
    int64_t b = 1ll << 30;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) b = (int64_t)(total_b / 24);
    if (b < (1ll << 30)) b = 1ll << 30;
    if (b > (8ll << 30)) b = 8ll << 30;
    return b;
  }();
  return v;
}

namespace {
WideLayout make_layout(const WNet& n, int n_dom) {
  WideLayout wl;
  const int64_t per_point = (2ll * n.L + (n.precision == 2 ? 5 : n.precision ? 3 : 0)) * n.C * n.H * 4;  // planes (+ re-tiled dW operands)
  int64_t mc = chunk_budget() / per_point;
  mc = mc / kBM * kBM;
  if (mc < 256) mc = 256;
  if (mc > 65536) mc = 65536;
  if (n.precision) {
    // tensor-core row GEMMs: one CTA per (P points, 128 units) tile, two CTAs per SM -> a wave is 2 * kBlocks CTAs.
    // Shrink the chunk to a whole number of waves (4800 points of mlp_xl were 2.03 waves: a third of the last one idle).
    const int P = tc_rows_points(n.C), units = (n.H + 127) / 128, wave = 2 * kBlocks;
    int64_t tiles = mc / P * units;
    if (tiles > wave) {
      int64_t t = tiles / wave * wave;          // whole waves
      while (t % units) t -= wave;              // ... that are also whole point tiles
      if (t >= wave) mc = t / units * P;
    }
  }
  wl.Mc = (int)mc;
  wl.tiles = (n.H + kBN - 1) / kBN;
  int ns = tc_dw_splits(n.H);  // slot count (>= the FP32 kernel's split count)
  wl.n_splits = ns;
  wl.fp32_splits = kBlocks / (wl.tiles * wl.tiles) < 1 ? 1 : kBlocks / (wl.tiles * wl.tiles);
  if (wl.fp32_splits > ns) wl.fp32_splits = ns;
  wl.n_dom = n_dom;
  // accumulators first: their offsets must not depend on the jet plan (domains with
  // different channel counts accumulate into the same slots); chunk planes last.
  int64_t off = 0;
  const int64_t planes = (int64_t)n.L * n.C * mc * n.H;
  wl.dw = off; off += (int64_t)(n.L - 1) * ns * n.H * n.H;
  wl.db = off; off += (int64_t)(n.L - 1) * kBlocks * 8 * n.H;
  wl.first = off; off += (int64_t)kBlocks * 8 * n.H;
  wl.head = off; off += (int64_t)kBlocks * (n.n_out * n.H + n.n_out + 1);
  wl.loss_slots = off; off += (int64_t)kMaxDomains * kBlocks;
  off = (off + 3) / 4 * 4;
  // W_l^T for the streamed tensor-core reverse GEMM, then the prepared weight tiles of the on-chip kernel
  wl.wt = off; off += ((int64_t)(n.L - 1) * n.H * n.H + 3) / 4 * 4;
  wl.wprep = off; off += (n.H <= 128) ? (fused_wide_wprep_floats(n) + 3) / 4 * 4 : 0;
  wl.wr = off; off += (n.precision == 1) ? ((int64_t)(n.L - 1) * n.H * n.H + 3) / 4 * 4 : 0;
  wl.acc_end = off;
  wl.Z = off; off += planes;
  wl.Hb = off; off += planes;
  wl.ybar = off; off += (int64_t)mc * kMaxC * 4;
  wl.tdw = off; off += n.precision ? (tc_dw2_scratch_floats(n.H, n.C, (int)mc, n.precision == 2) + 3) / 4 * 4 : 0;
  wl.total = (off + 3) / 4 * 4;
  return wl;
}

#define CK(call)                                   \
  do {                                             \
    cudaError_t e_ = (call);                       \
    if (e_ != cudaSuccess) return (int)e_;         \
  } while (0)

template <int ACT, bool BT, int EPI>
int launch_gemm_rows(const WNet& n, const float* A, const float* Wl, const float* bias, float* Zio, float* Hout, int m, int Mc,
                     cudaStream_t st) {
  dim3 grid((m + kBM - 1) / kBM, (n.H + kBN - 1) / kBN);
  switch (n.C) {
#define CASE(Cc) case Cc: gemm_rows_kernel<Cc, ACT, BT, EPI><<<grid, 256, 0, st>>>(n, A, Wl, bias, Zio, Hout, m, Mc); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7)
#undef CASE
    default: return PINN_E_UNSUPPORTED;
  }
  count_launch(1);
  return (int)cudaGetLastError();
}

template <int ACT>
int run_chunks(const WNet& n, const WideLayout& wl, float* ws, int mode, const pinn_domain* dom, int64_t N,
               const float* const* coords, float* jets_out, const float* jets_bar, cudaStream_t st) {
  const int H = n.H, C = n.C, L = n.L;
  const size_t planes = (size_t)C * wl.Mc * H;
  // TF32 mode: the operands of the tensor-core GEMMs (H_1 .. H_{L-1}, zbar_L .. zbar_2, the weight copies) are stored
  // rounded by their producers; H_L (output layer) and zbar_1 (first-layer gradient) stay FP32 (wide_tc.cu)
  const bool tf32 = n.precision == 1;
  const bool direct = n.precision != 0 && tc_dw_direct();  // dW operands by TMA straight from the planes; db from the producers of zbar
  Coords xs;
  for (int d = 0; d < PINN_MAX_IN; ++d) xs.p[d] = d < n.n_in ? coords[d] : nullptr;
  for (int64_t n0 = 0; n0 < N; n0 += wl.Mc) {
    const int m = (int)((N - n0 < wl.Mc) ? N - n0 : wl.Mc);
    float* Z = ws + wl.Z;
    float* Hb = ws + wl.Hb;
    {
      const int64_t tot = (int64_t)((m + kFirstPPT - 1) / kFirstPPT) * H;
      const unsigned fgrid = (unsigned)((tot + 255) / 256);
      switch (n.NF * 4 + n.NS) {  // NS <= NF <= 3 (make_net)
#define FIRST(F, S_) case F * 4 + S_: first_fwd_kernel<ACT, F, S_><<<fgrid, 256, 0, st>>>(n, xs, n0, m, wl.Mc, Z, Hb); break;
        FIRST(0, 0) FIRST(1, 0) FIRST(1, 1) FIRST(2, 0) FIRST(2, 1) FIRST(2, 2) FIRST(3, 0) FIRST(3, 1) FIRST(3, 2) FIRST(3, 3)
#undef FIRST
        default: return PINN_E_UNSUPPORTED;
      }
      count_launch(1);
    }
    for (int l = 2; l <= L; ++l) {
      int rc = n.precision
                   ? tc_launch_rows<ACT, 0>(n, Hb + (l - 2) * planes, tf32 ? ws + wl.wr + (size_t)(l - 2) * H * H : n.W[l - 1],
                                            n.b[l - 1], Z + (l - 1) * planes, Hb + (l - 1) * planes, m, wl.Mc, tf32 && l < L, nullptr, st)
                   : launch_gemm_rows<ACT, true, 0>(n, Hb + (l - 2) * planes, n.W[l - 1], n.b[l - 1], Z + (l - 1) * planes,
                                                    Hb + (l - 1) * planes, m, wl.Mc, st);
      if (rc) return rc;
    }
    {
      HeadArgs ha;
      memset(&ha, 0, sizeof(ha));
      ha.mode = mode == W_MODE_LOSS ? W_MODE_FUSED : mode;  // the head evaluates residual and loss as in the full step
      if (dom) ha.dom = *dom;
      ha.xs = xs;
      ha.n0 = n0;
      ha.n_total = N;
      ha.jets_out = jets_out;
      ha.jets_bar = jets_bar;
      ha.part = ws + wl.head;
      float* zl = Z + (L - 1) * planes;
      const float* hl = Hb + (L - 1) * planes;
      float* yb = ws + wl.ybar;
      const dim3 bgrid((H + 127) / 128, kBlocks);
#define HBWD(F, S_) case F * 4 + S_: head_bwd_kernel<ACT, F, S_><<<bgrid, 1024, 0, st>>>(n, yb, zl, hl, m, wl.Mc, ws + wl.head, dbl); break;
#define HEAD(KI_, C_)                                                                                       \
  case C_: {                                                                                                \
    int ob = (C * m + 31) / 32;                                                                             \
    if (ob > kBlocks * 4) ob = kBlocks * 4;                                                                 \
    out_layer_kernel<KI_><<<ob, 256, (size_t)n.n_out * H * 4, st>>>(n, hl, yb, m, wl.Mc);                   \
    int pb = (m + 255) / 256;                                                                               \
    if (pb > kBlocks) pb = kBlocks;                                                                         \
    head_point_kernel<C_><<<pb, 256, 0, st>>>(n, ha, yb, m);                                                \
    count_launch(1);                                                                                        \
    if (mode != W_MODE_FWD && mode != W_MODE_LOSS) {                                                                 \
      float* dbl = (direct && L >= 2) ? ws + wl.db + (size_t)(L - 2) * kBlocks * 8 * H : nullptr;         \
      switch (n.NF * 4 + n.NS) {                                                                            \
        HBWD(0, 0) HBWD(1, 0) HBWD(1, 1) HBWD(2, 0) HBWD(2, 1) HBWD(2, 2) HBWD(3, 0) HBWD(3, 1) HBWD(3, 2) HBWD(3, 3) \
        default: return PINN_E_UNSUPPORTED;                                                                 \
      }                                                                                                     \
      count_launch(1);                                                                                      \
    }                                                                                                       \
  } break;
      if (H <= 128) {
        switch (C) { HEAD(4, 1) HEAD(4, 2) HEAD(4, 3) HEAD(4, 4) HEAD(4, 5) HEAD(4, 6) HEAD(4, 7) default: return PINN_E_UNSUPPORTED; }
      } else {
        switch (C) { HEAD(16, 1) HEAD(16, 2) HEAD(16, 3) HEAD(16, 4) HEAD(16, 5) HEAD(16, 6) HEAD(16, 7) default: return PINN_E_UNSUPPORTED; }
      }
#undef HEAD
#undef HBWD
      count_launch(1);
      CK(cudaGetLastError());
    }
    if (mode == W_MODE_FWD || mode == W_MODE_LOSS) continue;
    for (int l = L; l >= 2; --l) {
      float* zb = Z + (l - 1) * planes;        // zbar_l (in place of Z_l)
      const float* hprev = Hb + (l - 2) * planes;
      {
        float* dwp = ws + wl.dw + (size_t)(l - 2) * wl.n_splits * H * H;
        float* dbp = ws + wl.db + (size_t)(l - 2) * kBlocks * 8 * H;
        if (direct) {
          // zbar_L's column sums were added by head_bwd_kernel; zbar_l (l < L) left its per-group sums in the scratch
          if (l < L) {
            int rc = tc_colsum_finish(ws + wl.tdw, (m + 15) / 16, H, dbp, st);
            if (rc) return rc;
          }
          int rc = tc_launch_dw3(zb, hprev, H, C, m, wl.Mc, dwp, wl.n_splits, n.precision == 2, st);
          if (rc) return rc;
        } else if (n.precision) {  // db_l rides along with the re-tiling of zbar_l
          int rc = tc_launch_dw2(zb, hprev, H, C, m, wl.Mc, dwp, wl.n_splits, ws + wl.tdw, dbp, n.precision == 2, st);
          if (rc) return rc;
        } else {
          dim3 grid(wl.tiles * wl.tiles, wl.fp32_splits);
          gemm_dw_kernel<<<grid, 256, 0, st>>>(zb, hprev, H, C, m, wl.Mc, dwp, wl.tiles);
          RedSpec rs;
          memset(&rs, 0, sizeof(rs));
          rs.nq = 1;
          reduce_cols_kernel<<<dim3((H + 31) / 32, kBlocks), 256, 0, st>>>(zb, rs, H, n0, m, wl.Mc, dbp);
          count_launch(2);
        }
      }
      int rc = n.precision
                   ? tc_launch_rows<ACT, 1>(n, zb, ws + wl.wt + (size_t)(l - 2) * H * H, nullptr, Z + (l - 2) * planes, nullptr,
                                            m, wl.Mc, tf32 && l > 2, (direct && l > 2) ? ws + wl.tdw : nullptr, st)
                   : launch_gemm_rows<ACT, false, 1>(n, zb, n.W[l - 1], nullptr, Z + (l - 2) * planes, nullptr, m, wl.Mc, st);
      if (rc) return rc;
    }
    {
      RedSpec rs;
      memset(&rs, 0, sizeof(rs));
      int q = 0;
      for (int d = 0; d < n.n_in; ++d) { rs.plane[q] = 0; rs.w[q] = coords[d]; rs.slot[q] = d; ++q; }
      rs.plane[q] = 0; rs.w[q] = nullptr; rs.slot[q] = PINN_MAX_IN; ++q;  // db1
      for (int i = 0; i < n.NF; ++i) { rs.plane[q] = 1 + i; rs.w[q] = nullptr; rs.slot[q] = n.first_in[i]; ++q; }
      rs.nq = q;
      reduce_cols_kernel<<<dim3((H + 31) / 32, kBlocks), 256, 0, st>>>(Z, rs, H, n0, m, wl.Mc, ws + wl.first);
      count_launch(1);
      CK(cudaGetLastError());
    }
  }
  return 0;
}

int run_mode(const pinn_mlp* m, const pinn_jets* j, int mode, const pinn_domain* dom, int64_t N, const float* const* coords,
             float* jets_out, const float* jets_bar, void* ws, int64_t ws_bytes, int n_slots, int slot, int accumulate,
             cudaStream_t st) {
  WNet n;
  if (!make_net(m, j, &n)) return PINN_E_UNSUPPORTED;
  if (n.n_in + 1 + n.NF > 8 || n_slots > kMaxDomains) return PINN_E_UNSUPPORTED;
  const WideLayout wl = make_layout(n, n_slots);
  if (!ws || ws_bytes < wl.total * 4) return PINN_E_WORKSPACE;
  float* w = (float*)ws;
  if (!accumulate && mode != W_MODE_FWD)  // clear every partial slot (all domains' loss slots included)
    CK(cudaMemsetAsync(w + wl.dw, 0, (size_t)(wl.wt - wl.dw) * 4, st));  // partial slots only (not the weight tiles)
  int rc = 0;
  const bool stepping = mode == W_MODE_FUSED || mode == W_MODE_LOSS;
  // whole step on chip: TF32 (16-point tiles) and, where the doubled operand buffers fit, FP32-grade 3xTF32 (8-point
  // tiles); PINN_PREC_TF32_STREAMED and everything else take the layer-streamed kernels
  const bool on_chip = stepping && fused_wide_supported(n) &&
                       (m->precision == PINN_PREC_TF32 || m->precision == PINN_PREC_FP32X3) &&
                       fused_wide_stash_floats(n) <= (int64_t)2 * n.L * n.C * wl.Mc * n.H;
  // weight preparation once per step: later domains of the same step (accumulate != 0) reuse it
  const bool prep = !accumulate || !stepping;
  if (n.precision && mode != W_MODE_FWD && mode != W_MODE_LOSS && prep)
    for (int l = 2; l <= n.L; ++l) {
      rc = tc_transpose(n.W[l - 1], w + wl.wt + (size_t)(l - 2) * n.H * n.H, n.H, n.precision == 1, st);
      if (rc) return rc;
    }
  if (n.precision == 1 && prep)
    for (int l = 2; l <= n.L; ++l) {
      rc = tc_round_copy(n.W[l - 1], w + wl.wr + (size_t)(l - 2) * n.H * n.H, n.H * n.H, st);
      if (rc) return rc;
    }
  // (prepared whenever ANY domain of this net could take the on-chip kernel: whether one does depends on its channel
  // count, and later domains of the step reuse the first one's preparation)
  if (stepping && prep && (m->precision == PINN_PREC_TF32 || m->precision == PINN_PREC_FP32X3) && n.H <= 128 && n.L >= 2) {
    rc = fused_wide_prep(n, w + wl.wprep, st);
    if (rc) return rc;
  }
  if (on_chip) {
    // whole step on chip; same partial slots, same finalize
#define RUNF(A)                                                                                                   \
  rc = fused_wide_launch<A>(n, dom, w + wl.Z, w + wl.wprep, w + wl.dw, w + wl.db, w + wl.first, w + wl.head, wl.n_splits, mode == W_MODE_LOSS, st)
    ACT_SWITCH(n.act, RUNF);
#undef RUNF
    if (rc) return rc;
    wide_collect_loss_kernel<<<1, 256, 0, st>>>(w, wl, n.n_out * n.H + n.n_out + 1, slot);
    count_launch(1);
    return (int)cudaGetLastError();
  }
#define RUN(A) rc = run_chunks<A>(n, wl, w, mode, dom, N, coords, jets_out, jets_bar, st)
  ACT_SWITCH(n.act, RUN);
#undef RUN
  if (rc) return rc;
  if (stepping) {
    wide_collect_loss_kernel<<<1, 256, 0, st>>>(w, wl, n.n_out * n.H + n.n_out + 1, slot);
    count_launch(1);
  }
  return (int)cudaGetLastError();
}

}  // namespace

bool wide_supported(const pinn_mlp* m, const pinn_jets* j) {
  WNet n;
  return make_net(m, j, &n) && (n.n_in + 1 + n.NF <= 8);
}

int64_t wide_workspace_bytes(const pinn_mlp* m, const pinn_jets* j, int n_domains) {
  WNet n;
  if (!make_net(m, j, &n)) return 0;
  return make_layout(n, n_domains).total * 4;
}

int wide_step_fused(const pinn_mlp* m, const pinn_jets* j, const pinn_domain* dom, void* ws, int64_t ws_bytes, int n_slots,
                    int domain_slot, int accumulate, bool loss_only, cudaStream_t st) {
  const float* coords[PINN_MAX_IN];
  for (int d = 0; d < PINN_MAX_IN; ++d) coords[d] = dom->coords[d];
  return run_mode(m, j, loss_only ? W_MODE_LOSS : W_MODE_FUSED, dom, dom->n_points, coords, nullptr, nullptr, ws, ws_bytes, n_slots, domain_slot,
                  accumulate, st);
}

int wide_step_finalize(const pinn_mlp* m, const pinn_jets* j, void* ws, int64_t ws_bytes, int n_domains, float* flat_grad,
                       int add_to_grad, float* losses, cudaStream_t st) {
  WNet n;
  if (!make_net(m, j, &n)) return PINN_E_UNSUPPORTED;
  const WideLayout wl = make_layout(n, n_domains);
  if (!ws || ws_bytes < wl.total * 4) return PINN_E_WORKSPACE;
  const int64_t np = pinn_param_count(m);
  wide_finalize_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(n, (const float*)ws, wl, np, flat_grad, add_to_grad, losses,
                                                                     n_domains);
  count_launch(1);
  return (int)cudaGetLastError();
}

int wide_jet_forward(const pinn_mlp* m, const pinn_jets* j, int64_t n_points, const float* const* coords, float* out, void* ws,
                     int64_t ws_bytes, cudaStream_t st) {
  return run_mode(m, j, W_MODE_FWD, nullptr, n_points, coords, out, nullptr, ws, ws_bytes, 1, 0, 0, st);
}

int wide_jet_backward(const pinn_mlp* m, const pinn_jets* j, int64_t n_points, const float* const* coords, const float* out_bar,
                      void* ws, int64_t ws_bytes, int accumulate, cudaStream_t st) {
  return run_mode(m, j, W_MODE_BWD, nullptr, n_points, coords, nullptr, out_bar, ws, ws_bytes, 1, 0, accumulate, st);
}

}  // namespace pinn
