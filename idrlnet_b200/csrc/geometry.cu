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

// ---------------------------------------------------------------------------------------------------------
// Expression programs: the parametric boundary pieces of a solid (idrlnet/geo_utils/geo.py:38-108 `Edge`: coordinates,
// normals and the measure as SymPy functions of the edge parameters) and symbolic signed distances, evaluated on the
// device.  The host compiles every SymPy expression once into a postfix program over
//   variables = the Philox-drawn parameters of the point (edge parameter(s), time / design-parameter ranges)
// and one launch draws the parameters of `n` points and writes every output column (replaces the NumPy lambdify +
// per-key H2D copy of `Edge.sample`, idrlnet/geo_utils/geo.py:80-108).
namespace pinn {
namespace {

struct ExprArgs {
  float* out[PINN_EXPR_MAX_OUT];
  int32_t begin[PINN_EXPR_MAX_OUT + 1];
  float lo[PINN_EXPR_MAX_VARS], hi[PINN_EXPR_MAX_VARS];
  int32_t n_out, n_vars;
  int64_t n;
  uint64_t seed, stream_id, offset;
  pinn_expr_op ops[PINN_EXPR_MAX_OPS];
};

__global__ void __launch_bounds__(kBlock) expr_program_kernel(const __grid_constant__ ExprArgs a) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= a.n) return;
  float var[PINN_EXPR_MAX_VARS];
#pragma unroll
  for (int g = 0; g < PINN_EXPR_MAX_VARS / 4; ++g) {
    if (g * 4 < a.n_vars) {
      uint32_t r[4];
      philox4x32_10(a.seed, a.stream_id + (uint64_t)g, a.offset + (uint64_t)i, r);
#pragma unroll
      for (int k = 0; k < 4; ++k) var[g * 4 + k] = box_point(r[k], a.lo[g * 4 + k], a.hi[g * 4 + k]);
    }
  }
  for (int o = 0; o < a.n_out; ++o) {
    float st[PINN_EXPR_STACK];
    int sp = 0;
    for (int pc = a.begin[o]; pc < a.begin[o + 1]; ++pc) {
      const pinn_expr_op op = a.ops[pc];
      switch (op.op) {
        case PINN_EX_CONST: st[sp++] = op.imm; break;
        case PINN_EX_VAR: st[sp++] = var[(int)op.imm]; break;
        case PINN_EX_ADD: --sp; st[sp - 1] += st[sp]; break;
        case PINN_EX_SUB: --sp; st[sp - 1] -= st[sp]; break;
        case PINN_EX_MUL: --sp; st[sp - 1] *= st[sp]; break;
        case PINN_EX_DIV: --sp; st[sp - 1] /= st[sp]; break;
        case PINN_EX_MIN: --sp; st[sp - 1] = fminf(st[sp - 1], st[sp]); break;
        case PINN_EX_MAX: --sp; st[sp - 1] = fmaxf(st[sp - 1], st[sp]); break;
        case PINN_EX_POW: --sp; st[sp - 1] = powf(st[sp - 1], st[sp]); break;
        case PINN_EX_NEG: st[sp - 1] = -st[sp - 1]; break;
        case PINN_EX_ABS: st[sp - 1] = fabsf(st[sp - 1]); break;
        case PINN_EX_SQRT: st[sp - 1] = sqrtf(st[sp - 1]); break;
        case PINN_EX_SIN: st[sp - 1] = sinf(st[sp - 1]); break;
        case PINN_EX_COS: st[sp - 1] = cosf(st[sp - 1]); break;
        case PINN_EX_TAN: st[sp - 1] = tanf(st[sp - 1]); break;
        case PINN_EX_EXP: st[sp - 1] = expf(st[sp - 1]); break;
        case PINN_EX_LOG: st[sp - 1] = logf(st[sp - 1]); break;
        case PINN_EX_TANH: st[sp - 1] = tanhf(st[sp - 1]); break;
        case PINN_EX_SIGN: st[sp - 1] = (st[sp - 1] > 0.f) ? 1.f : ((st[sp - 1] < 0.f) ? -1.f : 0.f); break;
        case PINN_EX_HEAVISIDE: st[sp - 1] = (st[sp - 1] > 0.f) ? 1.f : ((st[sp - 1] < 0.f) ? 0.f : 0.5f); break;
        case PINN_EX_SQUARE: st[sp - 1] *= st[sp - 1]; break;
        case PINN_EX_RCP: st[sp - 1] = 1.0f / st[sp - 1]; break;
        case PINN_EX_ATAN2: --sp; st[sp - 1] = atan2f(st[sp - 1], st[sp]); break;
        case PINN_EX_SCALE: st[sp - 1] *= op.imm; break;    // * immediate
        case PINN_EX_OFFSET: st[sp - 1] += op.imm; break;   // + immediate
        case PINN_EX_ASIN: st[sp - 1] = asinf(fminf(fmaxf(st[sp - 1], -1.f), 1.f)); break;
        case PINN_EX_ACOS: st[sp - 1] = acosf(fminf(fmaxf(st[sp - 1], -1.f), 1.f)); break;
        case PINN_EX_ATAN: st[sp - 1] = atanf(st[sp - 1]); break;
        case PINN_EX_SINH: st[sp - 1] = sinhf(st[sp - 1]); break;
        case PINN_EX_COSH: st[sp - 1] = coshf(st[sp - 1]); break;
        default: break;
      }
    }
    a.out[o][i] = sp > 0 ? st[0] : 0.f;
  }
}

}  // namespace
}  // namespace pinn

extern "C" int pinn_sample_program(float* const* out, int32_t n_out, const pinn_expr_op* ops, const int32_t* begin,
                                   const float* lo, const float* hi, int32_t n_vars, int64_t n_points, uint64_t seed,
                                   uint64_t stream_id, uint64_t offset, void* stream) {
  using namespace pinn;
  if (!out || !ops || !begin || n_out < 1 || n_out > PINN_EXPR_MAX_OUT || n_vars < 0 || n_vars > PINN_EXPR_MAX_VARS || n_points < 0 ||
      (n_vars > 0 && (!lo || !hi)))
    return PINN_E_INVALID;
  if (begin[0] != 0 || begin[n_out] > PINN_EXPR_MAX_OPS) return PINN_E_INVALID;
  ExprArgs a;
  memset(&a, 0, sizeof(a));
  for (int o = 0; o < n_out; ++o) {
    if (!out[o] || begin[o + 1] < begin[o]) return PINN_E_INVALID;
    a.out[o] = out[o];
  }
  for (int o = 0; o <= n_out; ++o) a.begin[o] = begin[o];
  // validate the programs on the host: stack depth, variable indices, op codes
  for (int o = 0; o < n_out; ++o) {
    int sp = 0;
    for (int pc = begin[o]; pc < begin[o + 1]; ++pc) {
      const int op = ops[pc].op;
      int pop = 0, push = 0;
      if (op == PINN_EX_CONST) push = 1;
      else if (op == PINN_EX_VAR) {
        push = 1;
        if ((int)ops[pc].imm < 0 || (int)ops[pc].imm >= n_vars) return PINN_E_INVALID;
      } else if (op == PINN_EX_ADD || op == PINN_EX_SUB || op == PINN_EX_MUL || op == PINN_EX_DIV || op == PINN_EX_MIN ||
                 op == PINN_EX_MAX || op == PINN_EX_POW || op == PINN_EX_ATAN2) { pop = 2; push = 1; }
      else if (op >= PINN_EX_NEG && op <= PINN_EX_COSH) { pop = 1; push = 1; }
      else return PINN_E_INVALID;
      if (sp < pop) return PINN_E_INVALID;
      sp += push - pop;
      if (sp > PINN_EXPR_STACK) return PINN_E_UNSUPPORTED;
    }
    if (sp != 1) return PINN_E_INVALID;
    for (int pc = begin[o]; pc < begin[o + 1]; ++pc) a.ops[pc] = ops[pc];
  }
  for (int v = 0; v < n_vars; ++v) { a.lo[v] = lo[v]; a.hi[v] = hi[v]; }
  a.n_out = n_out; a.n_vars = n_vars; a.n = n_points;
  a.seed = seed; a.stream_id = stream_id; a.offset = offset;
  if (n_points == 0) return 0;
  const int64_t blocks = (n_points + kBlock - 1) / kBlock;
  expr_program_kernel<<<(unsigned)blocks, kBlock, 0, (cudaStream_t)stream>>>(a);
  count_launch(1);
  return (int)cudaGetLastError();
}
