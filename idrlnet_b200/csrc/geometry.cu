// On-device sampling of CSG solids ("geometry sampling moved on-device").
//
// The reference draws N uniform points in the bounding box on the host, evaluates the SymPy signed
// distance and keeps sdf > 0 (idrlnet/geo_utils/geo.py:204-244).  Here one pass does the same on the
// GPU: candidate i is the Philox4x32-10 point of pinn_sample_box (same counter -> same point), its
// signed distance comes from a small stack program over primitives (idrlnet/geo_utils/geo_obj.py
// closed forms) and CSG operators (idrlnet/geo_utils/geo.py:246-300: union = max, difference =
// min(a, -b), intersection = min), and the survivors are compacted IN CANDIDATE ORDER (block counts,
// one-block scan, ordered write), so the output is a deterministic function of (seed, stream, offset).
#include <cuda_runtime.h>
#include <string.h>

#include "philox.h"
#include "wide_internal.cuh"

namespace pinn {
namespace {

constexpr int kMaxOps = PINN_MAX_SHAPE_OPS;
constexpr int kBlock = 256;

struct CsgArgs {
  float* col[PINN_MAX_IN];
  float lo[PINN_MAX_IN], hi[PINN_MAX_IN];
  float* sdf;
  int64_t* count;         // device: number of points kept
  unsigned int* block_n;  // [blocks + 1] per-block survivors, then their exclusive scan
  int n_dims, n_ops;
  int64_t n, capacity;
  uint64_t seed, stream_id, offset;
  pinn_shape_op prog[kMaxOps];
};

__device__ __forceinline__ float slab_sdf(const float* x, const float* lo, const float* hi, int dims) { return box_sdf(x, lo, hi, dims); }

// signed distance (positive inside) of the CSG program at x
__device__ float csg_sdf(const CsgArgs& a, const float* x) {
  float st[8];
  int sp = 0;
  for (int i = 0; i < a.n_ops; ++i) {
    const pinn_shape_op& o = a.prog[i];
    switch (o.op) {
      case PINN_SHAPE_INTERVAL: {  // p = (x0, x1), axis = p[2]
        const int ax = (int)o.p[2];
        const float h = 0.5f * (o.p[1] - o.p[0]);
        st[sp++] = h - fabsf(x[ax] - (o.p[0] + h));
      } break;
      case PINN_SHAPE_RECT: {  // p = (x0, y0, x1, y1)
        const float lo[2] = {o.p[0], o.p[1]}, hi[2] = {o.p[2], o.p[3]};
        st[sp++] = slab_sdf(x, lo, hi, 2);
      } break;
      case PINN_SHAPE_CIRCLE: {  // p = (cx, cy, r)
        const float dx = x[0] - o.p[0], dy = x[1] - o.p[1];
        st[sp++] = o.p[2] - sqrtf(dx * dx + dy * dy);
      } break;
      case PINN_SHAPE_CHANNEL2D: {  // p = (y0, y1): only the two walls bound it
        const float h = 0.5f * (o.p[1] - o.p[0]);
        st[sp++] = h - fabsf(x[1] - (o.p[0] + h));
      } break;
      case PINN_SHAPE_BOX: {  // p = (x0, y0, z0, x1, y1, z1)
        const float lo[3] = {o.p[0], o.p[1], o.p[2]}, hi[3] = {o.p[3], o.p[4], o.p[5]};
        st[sp++] = slab_sdf(x, lo, hi, 3);
      } break;
      case PINN_SHAPE_SPHERE: {  // p = (cx, cy, cz, r)
        const float dx = x[0] - o.p[0], dy = x[1] - o.p[1], dz = x[2] - o.p[2];
        st[sp++] = o.p[3] - sqrtf(dx * dx + dy * dy + dz * dz);
      } break;
      case PINN_CSG_UNION: --sp; st[sp - 1] = fmaxf(st[sp - 1], st[sp]); break;
      case PINN_CSG_DIFFERENCE: --sp; st[sp - 1] = fminf(st[sp - 1], -st[sp]); break;
      case PINN_CSG_INTERSECTION: --sp; st[sp - 1] = fminf(st[sp - 1], st[sp]); break;
      case PINN_CSG_COMPLEMENT: st[sp - 1] = -st[sp - 1]; break;
      default: break;
    }
  }
  return st[0];
}

__device__ __forceinline__ float candidate(const CsgArgs& a, int64_t i, float (&x)[PINN_MAX_IN]) {
  uint32_t r[4];
  philox4x32_10(a.seed, a.stream_id, a.offset + (uint64_t)i, r);
#pragma unroll
  for (int d = 0; d < PINN_MAX_IN; ++d) x[d] = d < a.n_dims ? box_point(r[d], a.lo[d], a.hi[d]) : 0.f;
  return csg_sdf(a, x);
}

// pass 1: survivors per block of kBlock consecutive candidates
__global__ void __launch_bounds__(kBlock) csg_count_kernel(const __grid_constant__ CsgArgs a) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  float x[PINN_MAX_IN];
  const bool keep = i < a.n && candidate(a, i, x) > 0.f;
  const int n = __syncthreads_count(keep);
  if (threadIdx.x == 0) a.block_n[blockIdx.x] = (unsigned int)n;
}

// pass 2: exclusive scan of the block counts in one block (blocks <= a few hundred thousand)
__global__ void __launch_bounds__(1024) csg_scan_kernel(unsigned int* block_n, int blocks, int64_t* count, int64_t capacity) {
  __shared__ unsigned int part[1024];
  const int per = (blocks + 1023) / 1024;
  const int b0 = threadIdx.x * per, b1 = min(b0 + per, blocks);
  unsigned int s = 0;
  for (int b = b0; b < b1; ++b) s += block_n[b];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int run = 0;
    for (int t = 0; t < 1024; ++t) {
      const unsigned int v = part[t];
      part[t] = run;
      run += v;
    }
    *count = (int64_t)run < capacity ? (int64_t)run : capacity;
    block_n[blocks] = run;
  }
  __syncthreads();
  unsigned int run = part[threadIdx.x];
  for (int b = b0; b < b1; ++b) {
    const unsigned int v = block_n[b];
    block_n[b] = run;
    run += v;
  }
}

// pass 3: recompute the candidate (cheaper than storing it) and write it at its ordered position
__global__ void __launch_bounds__(kBlock) csg_write_kernel(const __grid_constant__ CsgArgs a) {
  __shared__ unsigned int warp_n[kBlock / 32];
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  float x[PINN_MAX_IN];
  float sdf = 0.f;
  bool keep = false;
  if (i < a.n) {
    sdf = candidate(a, i, x);
    keep = sdf > 0.f;
  }
  const unsigned int ballot = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_n[warp] = __popc(ballot);
  __syncthreads();
  unsigned int base = a.block_n[blockIdx.x];
  for (int w = 0; w < warp; ++w) base += warp_n[w];
  const int64_t pos = (int64_t)base + __popc(ballot & ((1u << lane) - 1u));
  if (keep && pos < a.capacity) {
#pragma unroll
    for (int d = 0; d < PINN_MAX_IN; ++d)
      if (d < a.n_dims) a.col[d][pos] = x[d];
    if (a.sdf) a.sdf[pos] = sdf;
  }
}

}  // namespace

int64_t csg_workspace_bytes(int64_t n) { return (int64_t)((n + kBlock - 1) / kBlock + 2) * (int64_t)sizeof(unsigned int); }

int csg_sample(float* const* columns, int n_dims, const float* lo, const float* hi, int64_t n, int64_t capacity, uint64_t seed,
               uint64_t stream_id, uint64_t offset, const pinn_shape_op* prog, int n_ops, float* sdf, int64_t* count, void* ws,
               cudaStream_t st) {
  CsgArgs a;
  memset(&a, 0, sizeof(a));
  for (int d = 0; d < n_dims; ++d) {
    a.col[d] = columns[d];
    a.lo[d] = lo[d];
    a.hi[d] = hi[d];
  }
  a.sdf = sdf;
  a.count = count;
  a.block_n = (unsigned int*)ws;
  a.n_dims = n_dims;
  a.n_ops = n_ops;
  a.n = n;
  a.capacity = capacity;
  a.seed = seed;
  a.stream_id = stream_id;
  a.offset = offset;
  for (int i = 0; i < n_ops; ++i) a.prog[i] = prog[i];
  const int64_t blocks = (n + kBlock - 1) / kBlock;
  if (blocks == 0) return (int)cudaMemsetAsync(count, 0, sizeof(int64_t), st);
  csg_count_kernel<<<(unsigned)blocks, kBlock, 0, st>>>(a);
  csg_scan_kernel<<<1, 1024, 0, st>>>(a.block_n, (int)blocks, count, capacity);
  csg_write_kernel<<<(unsigned)blocks, kBlock, 0, st>>>(a);
  count_launch(3);
  return (int)cudaGetLastError();
}

}  // namespace pinn
