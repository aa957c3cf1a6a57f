// Self-test of the tcgen05 plumbing (descriptors, TMEM, mbarrier): one CTA computes a
// 128 x N x K TF32 GEMM from global operands in either operand majorness.
#include <stdio.h>

#include "pinn_common.h"
#include "tc_common.cuh"

namespace pinn {
void count_launch(int n);

// variant 0: A[128][K], B[N][K] (both K-major):   D[i][n] = sum_k A[i][k] * B[n][k]
// variant 1: A[K][128], B[K][N] (both MN-major):  D[i][n] = sum_k A[k][i] * B[k][n]
// variant 4: as 0, but the A operand is written to TMEM (tcgen05.st) and read from there by the MMA
__global__ void __launch_bounds__(128) tc_selftest_kernel(int variant, int N, int K, const float* __restrict__ A,
                                                          const float* __restrict__ B, float* __restrict__ D, int* status) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  float* sA = sm;
  float* sB = sm + 128 * K;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int amn = (variant == 1 || variant == 2) ? 1 : 0;   // A MN-major
  const int bmn = (variant == 1 || variant == 2 || variant == 3) ? 1 : 0;  // B MN-major
  const int swap = (variant == 2) ? 1 : 0;  // experiment: LBO/SBO roles swapped for MN-major
  if (variant == 4) {  // row = tid: K values into TMEM columns 256 .. 256 + K of this thread's lane
    for (int k0 = 0; k0 < K; k0 += 8) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = tc::to_tf32(A[(size_t)tid * K + k0 + i]);
      tc::tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 256 + k0, v);
    }
    tc::tmem_st_wait();
    tc::tc_fence_before();
  } else if (!amn) {
    // sX[k/4][row][4]
    for (int i = tid; i < 128 * (K / 4); i += 128) {
      const int row = i / (K / 4), kq = i % (K / 4);
      const float4 v = tc::to_tf32(*reinterpret_cast<const float4*>(A + (size_t)row * K + kq * 4));
      *reinterpret_cast<float4*>(sA + ((size_t)kq * 128 + row) * 4) = v;
    }
  } else {
    // sX[k/8][mn/4][k%8][4]
    for (int i = tid; i < K * 32; i += 128) {
      const int k = i / 32, mq = i % 32;
      const float4 v = tc::to_tf32(*reinterpret_cast<const float4*>(A + (size_t)k * 128 + mq * 4));
      *reinterpret_cast<float4*>(sA + (((size_t)(k / 8) * 32 + mq) * 8 + (k % 8)) * 4) = v;
    }
  }
  if (!bmn) {
    for (int i = tid; i < N * (K / 4); i += 128) {
      const int row = i / (K / 4), kq = i % (K / 4);
      const float4 v = tc::to_tf32(*reinterpret_cast<const float4*>(B + (size_t)row * K + kq * 4));
      *reinterpret_cast<float4*>(sB + ((size_t)kq * N + row) * 4) = v;
    }
  } else {
    for (int i = tid; i < K * (N / 4); i += 128) {
      const int k = i / (N / 4), nq = i % (N / 4);
      const float4 v = tc::to_tf32(*reinterpret_cast<const float4*>(B + (size_t)k * N + nq * 4));
      *reinterpret_cast<float4*>(sB + (((size_t)(k / 8) * (N / 4) + nq) * 8 + (k % 8)) * 4) = v;
    }
  }
  tc::fence_async_smem();
  __syncthreads();
  if (tid == 0) {
    tc::tc_fence_after();
    const uint32_t idesc = tc::idesc_tf32(N, amn, bmn);
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t da, db;
      if (!amn) da = tc::smem_desc(tc::smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
      else if (!swap) da = tc::smem_desc(tc::smem_u32(sA) + ks * (32 * 128), 32 * 128, 128);
      else da = tc::smem_desc(tc::smem_u32(sA) + ks * (32 * 128), 128, 32 * 128);
      if (!bmn) db = tc::smem_desc(tc::smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
      else if (!swap) db = tc::smem_desc(tc::smem_u32(sB) + ks * ((N / 4) * 128), (N / 4) * 128, 128);
      else db = tc::smem_desc(tc::smem_u32(sB) + ks * ((N / 4) * 128), 128, (N / 4) * 128);
      if (variant == 4) tc::mma_tf32_ts(tmem, tmem + 256 + ks * 8, db, idesc, ks > 0 ? 1u : 0u);
      else tc::mma_tf32(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
    }
    tc::mma_commit(&bar);
  }
  const bool ok = tc::mbar_wait(&bar, 0);
  if (!ok) {
    if (tid == 0) *status = 1;
  } else {
    tc::tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) D[(size_t)(warp * 32 + lane) * N + c0 + i] = v[i];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace pinn

extern "C" int pinn_tc_selftest(int variant, int N, int K, const float* A, const float* B, float* D, int* status,
                                          void* stream) {
  if (N % 16 || N < 16 || N > 256 || K % 8 || K < 8 || (variant == 4 && K > 256) || (128 + N) * K * 4 > 200 * 1024) return PINN_E_INVALID;
  const size_t smem = (size_t)(128 + N) * K * 4;
  cudaError_t e = cudaFuncSetAttribute((const void*)pinn::tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  pinn::tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(variant, N, K, A, B, D, status);
  pinn::count_launch(1);
  return (int)cudaGetLastError();
}

// ---- dense TF32 tcgen05 rate (roofline denominator of the wide tensor-core kernels) -------------------------
namespace pinn {
// One CTA per SM; thread 0 issues `iters` rounds of 8 MMAs (M = 128, N = 256, K = 8, kind::tf32, both operands
// resident in shared memory, K-major canonical layout) into two alternating TMEM accumulators: nothing but the
// tensor pipe and its shared-memory operand reads is exercised, which is the ceiling the wide kernels are held to.
__global__ void __launch_bounds__(128) tf32_mma_probe_kernel(int iters, float* __restrict__ scratch) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  constexpr int K = 64, N = 256;
  float* sA = sm;            // [K/4][128][4]
  float* sB = sm + 128 * K;  // [K/4][N][4]
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + N) * K; i += 128) {
    const uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    sm[i] = tc::to_tf32((float)((h >> 9) & 0xFFFF) * (1.0f / 65536.0f) - 0.5f);
  }
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_tf32(N, 0, 0);
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem + (uint32_t)(it & 1) * 256;
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t da = tc::smem_desc(tc::smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
        const uint64_t db = tc::smem_desc(tc::smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
        tc::mma_tf32(d, da, db, idesc, (it > 1 || ks > 0) ? 1u : 0u);
      }
    }
    tc::mma_commit(&bar);
  }
  const bool ok = tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  if (ok) {
    float v[4];
    tc::tmem_ld4(tmem + ((uint32_t)(warp * 32) << 16), v);
    tc::tmem_ld_wait();
    if (v[0] == 123.456f) scratch[blockIdx.x] = v[1];  // keep the result observable
  } else if (tid == 0) {
    scratch[blockIdx.x] = -1.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace pinn

extern "C" int64_t pinn_tf32_mma_probe(int32_t iters, float* scratch, void* stream) {
  if (iters < 1 || !scratch) return PINN_E_INVALID;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return PINN_E_NO_DEVICE;
  const size_t smem = (size_t)(128 + 256) * 64 * 4;
  cudaError_t e = cudaFuncSetAttribute((const void*)pinn::tf32_mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(int64_t)e;
  pinn::tf32_mma_probe_kernel<<<sms, 128, smem, (cudaStream_t)stream>>>(iters, scratch);
  pinn::count_launch(1);
  e = cudaGetLastError();
  if (e != cudaSuccess) return -(int64_t)e;
  return (int64_t)sms * iters * 8 * 2 * 128 * 256 * 8;
}
