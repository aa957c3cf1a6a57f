// Wide-net GEMMs on the 5th-generation tensor cores (tcgen05.mma kind::tf32, FP32
// accumulators in TMEM), hand-written PTX (tc_common.cuh).
//
// Orientation ("swapped"): the 128 TMEM lanes carry 128 hidden UNITS, the columns carry
// (channel, point) pairs, col = c*P + p.  One epilogue thread therefore owns one unit and
// sees every jet channel of a point in its own registers (the activation jet and its
// adjoint need all channels), and global stores are coalesced along the unit index.
//
//   tc_rows  EPI 0:  Z_l[c][n][j]  = sum_k W_l[j][k] H_{l-1}[c][n][k] + b ;  H_l = act(Z_l)
//            EPI 1:  zbar_{l-1}[c][n][k] = act'(Z_{l-1}; sum_j W_l^T[k][j] zbar_l[c][n][j])     (in place)
//   (the weight-gradient GEMM dW_l = zbar_l^T H_{l-1} lives in wide_dw.cu)
//
// Operands are staged by the CTA's threads into the no-swizzle K-major canonical layout
// (core matrix = 8 rows x 16 bytes), rounded to TF32 with cvt.rna; K is consumed in chunks of 64
// (32 with 3xTF32 splitting) through one smem stage released by tcgen05.commit -> mbarrier: the next
// chunk's global loads are in flight in registers while the tensor core works, and two CTAs per SM
// overlap each other's phases.  Inputs are FP32 in memory; only the multiplicands are rounded (10-bit
// mantissa, or hi + lo pairs in 3xTF32 mode), sums are FP32.  Stated tolerance of the TF32 mode 1e-2
// norm-wise (tests/test_gpu_wide_tc.py); the 3xTF32 mode is held to the FP32 bar (tests/test_gpu_wide.py).
#include <cuda_runtime.h>
#include <stdio.h>

#include "tc_common.cuh"
#include "wide_internal.cuh"

namespace pinn {
namespace {

__device__ __forceinline__ float4 sub4(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

template <int C>
struct TileCfg {
  static constexpr int P = (C == 1) ? 256 : (C == 2) ? 128 : (C == 3) ? 80 : (C == 4) ? 64 : (C == 5) ? 48 : 32;
  static constexpr int N = C * P;  // MMA N (multiple of 16, <= 256)
  static_assert(N % 16 == 0 && N <= 256 && P % 16 == 0, "tile");
};

// stage layout (floats): A[KT/4][128 + 1][4] then B[KT/4][NB + 1][4]: one 16-byte pad row per
// K-quad slab (LBO = (rows + 1) * 16 bytes) so that the staging stores of the K-quads of a row
// fall into different banks.  X3 (3xTF32, FP32-grade products): every operand is split into
// hi = tf32(x) and lo = tf32(x - hi), stored as two such blocks (hi first), and a K step issues
// lo*hi + hi*lo + hi*hi (the dropped lo*lo term and the split residuals are ~2^-22 relative).
template <int KT, bool X3>
__device__ __forceinline__ void issue_chunk(uint32_t tmem, uint32_t sA, uint32_t sB, int nb, uint32_t idesc, bool first) {
  const uint32_t a_lo = sA + (KT / 4) * 129 * 16, b_lo = sB + (KT / 4) * (nb + 1) * 16;
#pragma unroll
  for (int ks = 0; ks < KT / 8; ++ks) {
    const uint32_t oa = ks * 2 * (129 * 16), ob = ks * 2 * ((nb + 1) * 16);
    const uint64_t da = tc::smem_desc(sA + oa, 129 * 16, 128);
    const uint64_t db = tc::smem_desc(sB + ob, (nb + 1) * 16, 128);
    if (X3) {
      const uint64_t dal = tc::smem_desc(a_lo + oa, 129 * 16, 128);
      const uint64_t dbl = tc::smem_desc(b_lo + ob, (nb + 1) * 16, 128);
      tc::mma_tf32(tmem, dal, db, idesc, (first && ks == 0) ? 0u : 1u);
      tc::mma_tf32(tmem, da, dbl, idesc, 1u);
      tc::mma_tf32(tmem, da, db, idesc, 1u);
    } else {
      tc::mma_tf32(tmem, da, db, idesc, (first && ks == 0) ? 0u : 1u);
    }
  }
}

template <int C, int ACT, int EPI, bool X3>
__global__ void __launch_bounds__(256, 2) tc_rows_kernel(WNet net, const float* __restrict__ Bsrc, const float* __restrict__ Aw,
                                                      const float* __restrict__ bias, float* __restrict__ Zio,
                                                      float* __restrict__ Hout, int m, int Mc, int a_rows) {
  using T = TileCfg<C>;
  constexpr int P = T::P, N = T::N;
  constexpr int kKT = X3 ? 32 : 64;  // K elements per smem stage
  const int H = net.H;
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bar[1];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * P, j0 = blockIdx.y * 128;
  const size_t plane = (size_t)Mc * H;
  if (tid == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = tc::idesc_tf32(N, 0, 0);
  const int nch = (H + kKT - 1) / kKT;
  bool ok = true;
  float* sA = sm;
  float* sB = sA + (X3 ? 2 : 1) * (kKT / 4) * 129 * 4;
  constexpr int kLoA = (kKT / 4) * 129 * 4, kLoB = (kKT / 4) * (N + 1) * 4;  // offsets of the lo blocks (X3)
  constexpr int kQ = kKT / 4;                                                // K quads per stage
  for (int ch = 0; ch < nch; ++ch) {
    const int k0 = ch * kKT;
    // global -> registers first (many loads in flight), then wait for the previous chunk's MMAs
    // before overwriting the single smem stage.  Two CTAs per SM overlap each other's phases.
    constexpr int kItA = 128 * (kKT / 4) / 256, kItB = (N * (kKT / 4) + 255) / 256;
    float4 ra[kItA], rb[kItB];
#pragma unroll
    for (int u = 0; u < kItA; ++u) {
      const int i = tid + u * 256, row = i / kQ, kq = i % kQ;
      ra[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + row < a_rows && k0 + kq * 4 < H) ra[u] = *reinterpret_cast<const float4*>(Aw + (size_t)(j0 + row) * H + k0 + kq * 4);
    }
#pragma unroll
    for (int u = 0; u < kItB; ++u) {
      const int i = tid + u * 256, q = i / kQ, kq = i % kQ;
      const int c = q / P, p = q % P;
      rb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < N && n0 + p < m && k0 + kq * 4 < H)
        rb[u] = *reinterpret_cast<const float4*>(Bsrc + c * plane + (size_t)(n0 + p) * H + k0 + kq * 4);
    }
    if (ch >= 1) ok = tc::mbar_wait(&bar[0], (ch - 1) & 1) && ok;
#pragma unroll
    for (int u = 0; u < kItA; ++u) {
      const int i = tid + u * 256, row = i / kQ, kq = i % kQ;
      const float4 hi = tc::to_tf32(ra[u]);
      *reinterpret_cast<float4*>(sA + ((size_t)kq * 129 + row) * 4) = hi;
      if (X3) *reinterpret_cast<float4*>(sA + kLoA + ((size_t)kq * 129 + row) * 4) = tc::to_tf32(sub4(ra[u], hi));
    }
#pragma unroll
    for (int u = 0; u < kItB; ++u) {
      const int i = tid + u * 256, q = i / kQ, kq = i % kQ;
      if (q < N) {
        const float4 hi = tc::to_tf32(rb[u]);
        *reinterpret_cast<float4*>(sB + ((size_t)kq * (N + 1) + q) * 4) = hi;
        if (X3) *reinterpret_cast<float4*>(sB + kLoB + ((size_t)kq * (N + 1) + q) * 4) = tc::to_tf32(sub4(rb[u], hi));
      }
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      issue_chunk<kKT, X3>(tmem, tc::smem_u32(sA), tc::smem_u32(sB), N, idesc, ch == 0);
      tc::mma_commit(&bar[0]);
    }
  }
  ok = tc::mbar_wait(&bar[0], (nch - 1) & 1) && ok;
  if (!ok) __trap();
  tc::tc_fence_after();

  // epilogue: lane = hidden unit; 16-point groups alternate between the two warp quads
  const int j = j0 + (warp & 3) * 32 + lane;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const float bj = (EPI != 1 && j < a_rows) ? bias[j] : 0.f;
  for (int g = warp >> 2; g < P / 16; g += 2) {
    float v[C][16];
#pragma unroll
    for (int c = 0; c < C; ++c) tc::tmem_ld16(tmem + lane_base + c * P + g * 16, v[c]);
    if (EPI == 1) {
      // reverse layer: Z_{l-1} is fetched 4 points ahead (double-buffered registers) so the global
      // loads of the next quad are in flight while the current quad's adjoints are computed
      float zin[2][C][4];
      auto fetch = [&](int quad, float (&dst)[C][4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = n0 + g * 16 + quad * 4 + i;
#pragma unroll
          for (int c = 0; c < C; ++c) dst[c][i] = (j < H && n < m) ? Zio[c * plane + (size_t)n * H + j] : 0.f;
        }
      };
      fetch(0, zin[0]);
      tc::tmem_ld_wait();
#pragma unroll
      for (int quad = 0; quad < 4; ++quad) {
        if (quad + 1 < 4) fetch(quad + 1, zin[(quad + 1) & 1]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = n0 + g * 16 + quad * 4 + i;
          float a[kMaxC] = {}, b[kMaxC] = {}, o[kMaxC] = {};
#pragma unroll
          for (int c = 0; c < C; ++c) {
            a[c] = v[c][quad * 4 + i];
            b[c] = zin[quad & 1][c][i];
            o[c] = 0.f;
          }
          jet_bwd<ACT>(net.NF, net.NS, b, a, o);
          if (j < H && n < m) {
#pragma unroll
            for (int c = 0; c < C; ++c) Zio[c * plane + (size_t)n * H + j] = o[c];
          }
        }
      }
    } else {
      tc::tmem_ld_wait();
      if (j < a_rows) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = n0 + g * 16 + i;
          if (n < m) {
            float a[kMaxC] = {}, o[kMaxC] = {};
#pragma unroll
            for (int c = 0; c < C; ++c) a[c] = v[c][i];
            a[0] += bj;
            jet_fwd<ACT>(net.NF, net.NS, a, o);
#pragma unroll
            for (int c = 0; c < C; ++c) {
              Zio[c * plane + (size_t)n * H + j] = a[c];
              Hout[c * plane + (size_t)n * H + j] = o[c];
            }
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

__global__ void transpose_kernel(const float* __restrict__ W, float* __restrict__ WT, int H) {
  __shared__ float t[32][33];
  const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    if (x < H && y0 + r < H) t[r][threadIdx.x] = W[(size_t)(y0 + r) * H + x];
  __syncthreads();
  const int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    if (xo < H && yo0 + r < H) WT[(size_t)(yo0 + r) * H + xo] = t[threadIdx.x][r];
}

template <int C, int ACT, int EPI, bool X3>
int launch_rows_c(const WNet& n, const float* Bsrc, const float* Aw, const float* bias, float* Zio, float* Hout, int m, int Mc,
                  int a_rows, cudaStream_t st) {
  using T = TileCfg<C>;
  constexpr int KT = X3 ? 32 : 64;
  const size_t smem = (size_t)(X3 ? 2 : 1) * (KT / 4) * (129 + T::N + 1) * 4 * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute((const void*)tc_rows_kernel<C, ACT, EPI, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid((m + T::P - 1) / T::P, (a_rows + 127) / 128);
  tc_rows_kernel<C, ACT, EPI, X3><<<grid, 256, smem, st>>>(n, Bsrc, Aw, bias, Zio, Hout, m, Mc, a_rows);
  count_launch(1);
  return (int)cudaGetLastError();
}

}  // namespace

template <int ACT, int EPI>
int tc_launch_rows(const WNet& n, const float* Bsrc, const float* Aw, const float* bias, float* Zio, float* Hout, int m, int Mc,
                   cudaStream_t st) {
  const int a_rows = n.H;
  switch (n.C) {
#define CASE(Cc)                                                                                              \
  case Cc:                                                                                                    \
    return n.precision == 2 ? launch_rows_c<Cc, ACT, EPI, true>(n, Bsrc, Aw, bias, Zio, Hout, m, Mc, a_rows, st) \
                            : launch_rows_c<Cc, ACT, EPI, false>(n, Bsrc, Aw, bias, Zio, Hout, m, Mc, a_rows, st);
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7)
#undef CASE
  }
  return PINN_E_UNSUPPORTED;
}

#define INST(A)                                                                                                              \
  template int tc_launch_rows<A, 0>(const WNet&, const float*, const float*, const float*, float*, float*, int, int, cudaStream_t); \
  template int tc_launch_rows<A, 1>(const WNet&, const float*, const float*, const float*, float*, float*, int, int, cudaStream_t);
INST(PINN_ACT_TANH)
INST(PINN_ACT_SILU)
INST(PINN_ACT_SIN)
#undef INST

int tc_rows_points(int C) {
  return (C == 1) ? 256 : (C == 2) ? 128 : (C == 3) ? 80 : (C == 4) ? 64 : (C == 5) ? 48 : 32;  // TileCfg<C>::P
}

int tc_dw_splits(int H) {
  const int tiles = ((H + 127) / 128) * ((H + 255) / 256);
  int ns = 296 / tiles;  // two resident CTAs per SM
  return ns < 1 ? 1 : ns;
}

int tc_transpose(const float* W, float* WT, int H, cudaStream_t st) {
  dim3 grid((H + 31) / 32, (H + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, st>>>(W, WT, H);
  count_launch(1);
  return (int)cudaGetLastError();
}

}  // namespace pinn
