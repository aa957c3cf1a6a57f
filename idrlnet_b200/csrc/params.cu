// Parameter-sized helpers and the on-device sampler (see include/pinn_b200.h).
#include <cuda_runtime.h>
#include <stdio.h>

#include "pinn_common.h"
#include "philox.h"

namespace pinn {
void count_launch(int n);

// one warp per output row: W[j,:] = g[j] * v[j,:] / ||v[j,:]||
__device__ __forceinline__ void weight_norm_fwd_row(const float* __restrict__ g, const float* __restrict__ v,
                                                    float* __restrict__ W, int row, int in_dim, int lane) {
  const float* vr = v + (size_t)row * in_dim;
  float ss = 0.f;
  for (int k = lane; k < in_dim; k += 32) ss += vr[k] * vr[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float scale = g[row] / sqrtf(ss);
  for (int k = lane; k < in_dim; k += 32) W[(size_t)row * in_dim + k] = vr[k] * scale;
}

__global__ void weight_norm_fwd_kernel(const float* __restrict__ g, const float* __restrict__ v, float* __restrict__ W,
                                       int out_dim, int in_dim) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= out_dim) return;
  weight_norm_fwd_row(g, v, W, row, in_dim, threadIdx.x & 31);
}

// dg[j] = <dW_j, v_j>/||v_j|| ; dv[j,:] = g_j/||v_j|| * (dW_j - <dW_j, v_j>/||v_j||^2 * v_j)
__device__ __forceinline__ void weight_norm_bwd_row(const float* __restrict__ g, const float* __restrict__ v,
                                                    const float* __restrict__ dW, float* __restrict__ dg,
                                                    float* __restrict__ dv, int row, int in_dim, int lane) {
  const float* vr = v + (size_t)row * in_dim;
  const float* dr = dW + (size_t)row * in_dim;
  float ss = 0.f, dot = 0.f;
  for (int k = lane; k < in_dim; k += 32) {
    ss += vr[k] * vr[k];
    dot += dr[k] * vr[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  const float inv = rsqrtf(ss);
  if (lane == 0) dg[row] = dot * inv;
  const float a = g[row] * inv, b = dot / ss;
  for (int k = lane; k < in_dim; k += 32) dv[(size_t)row * in_dim + k] = a * (dr[k] - b * vr[k]);
}

__global__ void weight_norm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ v,
                                       const float* __restrict__ dW, float* __restrict__ dg, float* __restrict__ dv,
                                       int out_dim, int in_dim) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= out_dim) return;
  weight_norm_bwd_row(g, v, dW, dg, dv, row, in_dim, threadIdx.x & 31);
}

// all weight-norm layers of a net in one launch: warp -> (layer, row) through the row prefix sums
struct WnBatch {
  pinn_wn_layer l[PINN_MAX_LAYERS];
  int row_end[PINN_MAX_LAYERS];
  int n;
};

template <bool BWD>
__global__ void weight_norm_batch_kernel(const __grid_constant__ WnBatch b) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= b.row_end[b.n - 1]) return;
  int i = 0;
  while (row >= b.row_end[i]) ++i;
  if (i > 0) row -= b.row_end[i - 1];
  const pinn_wn_layer& y = b.l[i];
  if (BWD)
    weight_norm_bwd_row(y.g, y.v, y.dW, y.dg, y.dv, row, y.in_dim, threadIdx.x & 31);
  else
    weight_norm_fwd_row(y.g, y.v, y.W, row, y.in_dim, threadIdx.x & 31);
}

// FP32 FMA throughput probe: 16 independent chains per thread (roofline denominator
// for the narrow kernels; MEASURED_PEAKS.json only carries HBM and bf16 numbers).
__global__ void fma_probe_kernel(int iters, float* out) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0f + 1e-3f * (threadIdx.x + i);
  const float m = 1.0f - 1e-7f * (threadIdx.x & 3), c = 1e-9f * blockIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = a[i] * m + c;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678f) out[0] = s;  // keep the chains alive
}

// the same chains as packed pairs: fma.rn.f32x2 (SASS FFMA2), two FP32 FMAs per lane per instruction
__global__ void fma2_probe_kernel(int iters, float* out) {
  unsigned long long a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 v = make_float2(1.0f + 1e-3f * (threadIdx.x + 2 * i), 1.0f + 1e-3f * (threadIdx.x + 2 * i + 1));
    a[i] = *reinterpret_cast<const unsigned long long*>(&v);
  }
  const float mf = 1.0f - 1e-7f * (threadIdx.x & 3), cf = 1e-9f * blockIdx.x;
  const float2 m2 = make_float2(mf, mf), c2 = make_float2(cf, cf);
  const unsigned long long m = *reinterpret_cast<const unsigned long long*>(&m2), c = *reinterpret_cast<const unsigned long long*>(&c2);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(m), "l"(c));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 v = *reinterpret_cast<const float2*>(&a[i]);
    s += v.x + v.y;
  }
  if (s == 12345.678f) out[0] = s;
}

struct SampleArgs {
  float* col[PINN_MAX_IN];
  float lo[PINN_MAX_IN], hi[PINN_MAX_IN];
  float* sdf;
  int n_dims, sdf_dims;
  int64_t n;
  uint64_t seed, stream_id, offset;
};

__global__ void sample_box_kernel(const SampleArgs a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32_10(a.seed, a.stream_id, a.offset + (uint64_t)i, r);
    float xs[PINN_MAX_IN];
#pragma unroll
    for (int d = 0; d < PINN_MAX_IN; ++d) {
      if (d < a.n_dims) {
        xs[d] = box_point(r[d], a.lo[d], a.hi[d]);
        a.col[d][i] = xs[d];
      }
    }
    if (a.sdf) a.sdf[i] = box_sdf(xs, a.lo, a.hi, a.sdf_dims);
  }
}

}  // namespace pinn

using namespace pinn;

extern "C" {

int pinn_weight_norm_forward(const float* g, const float* v, float* W, int32_t out_dim, int32_t in_dim, void* stream) {
  if (!g || !v || !W || out_dim < 1 || in_dim < 1) return PINN_E_INVALID;
  const int wpb = 4;
  weight_norm_fwd_kernel<<<(out_dim + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(g, v, W, out_dim, in_dim);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  count_launch(1);
  return 0;
}

int pinn_weight_norm_backward(const float* g, const float* v, const float* dW, float* dg, float* dv, int32_t out_dim,
                              int32_t in_dim, void* stream) {
  if (!g || !v || !dW || !dg || !dv || out_dim < 1 || in_dim < 1) return PINN_E_INVALID;
  const int wpb = 4;
  weight_norm_bwd_kernel<<<(out_dim + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(g, v, dW, dg, dv, out_dim, in_dim);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  count_launch(1);
  return 0;
}

static int weight_norm_batch(const pinn_wn_layer* layers, int32_t n_layers, bool bwd, void* stream) {
  if (!layers || n_layers < 1 || n_layers > PINN_MAX_LAYERS) return PINN_E_INVALID;
  WnBatch b;
  int rows = 0;
  for (int i = 0; i < n_layers; ++i) {
    const pinn_wn_layer& y = layers[i];
    if (!y.g || !y.v || y.out_dim < 1 || y.in_dim < 1) return PINN_E_INVALID;
    if (bwd ? (!y.dW || !y.dg || !y.dv) : !y.W) return PINN_E_INVALID;
    b.l[i] = y;
    rows += y.out_dim;
    b.row_end[i] = rows;
  }
  b.n = n_layers;
  const int wpb = 4;
  if (bwd)
    weight_norm_batch_kernel<true><<<(rows + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(b);
  else
    weight_norm_batch_kernel<false><<<(rows + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  count_launch(1);
  return 0;
}

int pinn_weight_norm_forward_batch(const pinn_wn_layer* layers, int32_t n_layers, void* stream) {
  return weight_norm_batch(layers, n_layers, false, stream);
}

int pinn_weight_norm_backward_batch(const pinn_wn_layer* layers, int32_t n_layers, void* stream) {
  return weight_norm_batch(layers, n_layers, true, stream);
}

int64_t pinn_fp32_fma_probe(int32_t iters, int32_t blocks, float* scratch, void* stream) {
  if (iters < 1 || blocks < 1 || !scratch) return PINN_E_INVALID;
  fma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, scratch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -(int64_t)e;
  return (int64_t)2 * 16 * iters * 256 * blocks;  // flops issued
}

int64_t pinn_fp32_fma2_probe(int32_t iters, int32_t blocks, float* scratch, void* stream) {
  if (iters < 1 || blocks < 1 || !scratch) return PINN_E_INVALID;
  fma2_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, scratch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -(int64_t)e;
  return (int64_t)2 * 16 * iters * 256 * blocks;  // flops issued
}

int pinn_sample_box(float* const* columns, int32_t n_dims, const float* lo, const float* hi, int64_t n_points,
                    uint64_t seed, uint64_t stream_id, uint64_t offset, float* sdf, int32_t sdf_dims, void* stream) {
  if (!columns || !lo || !hi || n_dims < 1 || n_dims > PINN_MAX_IN || n_points < 0 || sdf_dims < 0 || sdf_dims > n_dims)
    return PINN_E_INVALID;
  if (n_points == 0) return 0;
  SampleArgs a;
  for (int d = 0; d < PINN_MAX_IN; ++d) {
    a.col[d] = d < n_dims ? columns[d] : nullptr;
    a.lo[d] = d < n_dims ? lo[d] : 0.f;
    a.hi[d] = d < n_dims ? hi[d] : 0.f;
    if (d < n_dims && !columns[d]) return PINN_E_INVALID;
  }
  a.sdf = sdf;
  a.n_dims = n_dims;
  a.sdf_dims = sdf_dims;
  a.n = n_points;
  a.seed = seed;
  a.stream_id = stream_id;
  a.offset = offset;
  int64_t blocks = (n_points + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sample_box_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  count_launch(1);
  return 0;
}

}  // extern "C"
