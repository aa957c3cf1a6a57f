// Device side of adaptive re-sampling: forward-only residuals of a
// candidate point set and the indices of the k largest |r|.
//
// The reference does this on the host: `Solver.infer_step` (idrlnet/solver.py:410-420) returns
// the residual column, the example's receiver copies it to NumPy, `np.argsort`s |r| and keeps the
// largest 200 (examples/allen_cahn/allen_cahn.py:128-149).  Here the residual is evaluated from
// the jets of pinn_jet_forward by one elementwise kernel and the selection is a radix select
// over the 64-bit keys (|r| bits << 32 | ~index): all keys are distinct, so the result is unique
// -- descending |r|, ties by ascending index, which is what a stable argsort of -|r| gives.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "select.cuh"
#include "wide_internal.cuh"

namespace pinn {
namespace {

// ---- residuals ---------------------------------------------------------------------------------
struct ResidualArgs {
  pinn_domain dom;
  const float* jets;  // [(c * n_out + o) * N + n]
  float* out;         // [e * N + n]
  int64_t n;
  int32_t n_out;
};

__device__ __forceinline__ float load_symbol(const ResidualArgs& a, int id, int64_t n) {
  if (id < PINN_SYM_COORD0) return a.jets[(int64_t)((id >> 2) * a.n_out + (id & 3)) * a.n + n];
  if (id < PINN_SYM_AUX0) return a.dom.coords[id - PINN_SYM_COORD0][n];
  return a.dom.aux[id - PINN_SYM_AUX0][n];
}

// HBM-bound: every symbol column a term names is read once per point (repeats hit L1), one
// float per equation is written.  Term and factor order match residual_loss of the fused kernels.
__global__ void __launch_bounds__(256) residual_kernel(const __grid_constant__ ResidualArgs a) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < a.n; n += (int64_t)gridDim.x * blockDim.x) {
    for (int e = 0; e < a.dom.n_eq; ++e) {
      const pinn_equation& q = a.dom.eq[e];
      float r = 0.f;
      for (int t = q.term_begin; t < q.term_end; ++t) {
        const pinn_term& tm = a.dom.terms[t];
        float prod = tm.coef;
        for (int f = 0; f < tm.n_fac; ++f) prod *= load_symbol(a, tm.fac[f], n);
        r += prod;
      }
      a.out[(int64_t)e * a.n + n] = r;
    }
  }
}

// ---- top-k of |v| ----------------------------------------------------------------------------------
struct TopkState {
  unsigned int hist[256];
  unsigned long long prefix;  // the digits selected so far (most significant first)
  unsigned int k_rem;         // rank still to be located inside the current prefix bucket
  unsigned int ticket;        // blocks finished in the current pass
  unsigned int n_found;       // compaction cursor
  unsigned int pad;
};

__device__ __forceinline__ unsigned long long topk_key(float v, uint32_t i) {
  return ((unsigned long long)(__float_as_uint(v) & 0x7fffffffu) << 32) | (uint32_t)(~i);
}

// One 8-bit digit of the radix select, most significant first.  Every block histograms the keys
// that carry the prefix chosen so far; the block that finishes last walks the 256 bins from the top
// and fixes the next digit.  No block waits for another one.
__global__ void __launch_bounds__(256) topk_pass_kernel(const float* __restrict__ v, int64_t n, int pass, unsigned int k,
                                                        TopkState* s) {
  __shared__ unsigned int h[256];
  __shared__ bool is_last;
  const int tid = threadIdx.x;
  h[tid] = 0;
  __syncthreads();
  const int shift = 56 - 8 * pass;
  const unsigned long long prefix = s->prefix;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = topk_key(v[i], (uint32_t)i);
    if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(&h[(unsigned int)(key >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (h[tid]) atomicAdd(&s->hist[tid], h[tid]);
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd(&s->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  h[tid] = *((volatile unsigned int*)&s->hist[tid]);
  s->hist[tid] = 0;
  __syncthreads();
  if (tid == 0) {
    const unsigned int want = pass == 0 ? k : s->k_rem;
    unsigned int above = 0;
    int d = 255;
    for (; d > 0; --d) {
      if (above + h[d] >= want) break;
      above += h[d];
    }
    s->prefix = (prefix << 8) | (unsigned long long)d;
    s->k_rem = want - above;
    s->ticket = 0;
  }
}

// after the last pass `prefix` is the key of the k-th largest element: keep everything >= it
__global__ void __launch_bounds__(256) topk_gather_kernel(const float* __restrict__ v, int64_t n, unsigned int k, TopkState* s,
                                                          unsigned long long* __restrict__ cand) {
  const unsigned long long thr = s->prefix;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = topk_key(v[i], (uint32_t)i);
    if (key >= thr) {
      const unsigned int pos = atomicAdd(&s->n_found, 1u);
      if (pos < k) cand[pos] = key;
    }
  }
}

// one block: bitonic sort (descending) of the k survivors, then decode
__global__ void __launch_bounds__(1024) topk_sort_kernel(const float* __restrict__ v, const unsigned long long* __restrict__ cand,
                                                         unsigned int k, unsigned int k_pow2, int64_t* __restrict__ idx_out,
                                                         float* __restrict__ val_out) {
  extern __shared__ unsigned long long keys[];
  for (unsigned int i = threadIdx.x; i < k_pow2; i += blockDim.x) keys[i] = i < k ? cand[i] : 0ull;
  __syncthreads();
  for (unsigned int size = 2; size <= k_pow2; size <<= 1) {
    for (unsigned int stride = size >> 1; stride > 0; stride >>= 1) {
      for (unsigned int i = threadIdx.x; i < k_pow2; i += blockDim.x) {
        const unsigned int partner = i ^ stride;
        if (partner > i) {
          const bool descending = (i & size) == 0;
          const unsigned long long a = keys[i], b = keys[partner];
          if ((a < b) == descending) {
            keys[i] = b;
            keys[partner] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (unsigned int i = threadIdx.x; i < k; i += blockDim.x) {
    const uint32_t idx = ~(uint32_t)(keys[i] & 0xffffffffull);
    idx_out[i] = (int64_t)idx;
    if (val_out) val_out[i] = v[idx];
  }
}

int grid_for(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)kBlocks * 8;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

int residual_eval(const pinn_domain* dom, int n_out, const float* jets, float* out, cudaStream_t st) {
  ResidualArgs a;
  memset(&a, 0, sizeof(a));
  a.dom = *dom;
  a.jets = jets;
  a.out = out;
  a.n = dom->n_points;
  a.n_out = n_out;
  if (a.n == 0) return 0;
  residual_kernel<<<grid_for(a.n), 256, 0, st>>>(a);
  count_launch(1);
  return (int)cudaGetLastError();
}

int64_t topk_workspace_bytes(int k) { return (int64_t)sizeof(TopkState) + (int64_t)k * 8 + 256; }

int topk_abs(const float* v, int64_t n, int k, int64_t* idx_out, float* val_out, void* ws, cudaStream_t st) {
  TopkState* s = (TopkState*)ws;
  unsigned long long* cand = (unsigned long long*)((char*)ws + ((sizeof(TopkState) + 255) & ~(size_t)255));
  cudaError_t e = cudaMemsetAsync(s, 0, sizeof(TopkState), st);
  if (e != cudaSuccess) return (int)e;
  const int grid = grid_for(n);
  for (int pass = 0; pass < 8; ++pass) topk_pass_kernel<<<grid, 256, 0, st>>>(v, n, pass, (unsigned int)k, s);
  topk_gather_kernel<<<grid, 256, 0, st>>>(v, n, (unsigned int)k, s, cand);
  unsigned int p2 = 1;
  while (p2 < (unsigned int)k) p2 <<= 1;
  topk_sort_kernel<<<1, 1024, p2 * sizeof(unsigned long long), st>>>(v, cand, (unsigned int)k, p2, idx_out, val_out);
  count_launch(10);
  return (int)cudaGetLastError();
}

}  // namespace pinn
