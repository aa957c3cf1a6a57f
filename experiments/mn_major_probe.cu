// Probe (round 2): which TMA swizzle mode + tcgen05 shared-memory descriptor reads an MN-major TF32 operand correctly?
//   A[k][m] (k = 32 points, m = 128 units contiguous), B[k][n] (n = 64): D[m][n] = sum_k A[k][m] * B[k][n].
// The operands arrive by TMA boxes {32 units, 32 k} (one per 32-unit group) with the swizzle mode under test; four MMAs of
// K = 8 read them through descriptors (layout type, LBO, SBO, K-step offset) under test.  Small-integer data: exact in TF32.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../idrlnet_b200/csrc -o build/mn_probe mn_major_probe.cu
//   ./build/mn_probe <tma swizzle enum value>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "tc_common.cuh"
using namespace pinn;

constexpr int K = 32, M = 128, N = 64;

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)type << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, uint32_t type,
                                             uint32_t lbo, uint32_t sbo, uint32_t kstep, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t raw[];
  __shared__ uint64_t full, done;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t base = ((tc::smem_u32(raw) + 1023u) & ~1023u) + 16384;  // (slack on both sides)
  const uint32_t sA = base, sB = base + 4 * K * 128;
  if (tid == 0) {
    tc::mbar_init(&full, 1);
    tc::mbar_init(&done, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&slot, 64);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = slot;
  if (tid == 0) {
    tc::mbar_expect_tx(&full, (4 + N / 32) * K * 128);
    for (int g = 0; g < 4; ++g) tc::tma_load_2d(sA + g * K * 128, &tmA, g * 32, 0, &full);
    for (int g = 0; g < N / 32; ++g) tc::tma_load_2d(sB + g * K * 128, &tmB, g * 32, 0, &full);
    if (!tc::mbar_wait(&full, 0)) __trap();
    tc::tc_fence_after();
    const uint32_t idesc = tc::idesc_tf32(N, 1, 1);
    for (int ks = 0; ks < K / 8; ++ks)
      tc::mma_tf32(tmem, desc(sA + ks * kstep, lbo, sbo, type), desc(sB + ks * kstep, lbo, sbo, type), idesc, ks ? 1u : 0u);
    tc::mma_commit(&done);
  }
  __syncthreads();
  if (!tc::mbar_wait(&done, 0)) __trap();
  tc::tc_fence_after();
  for (int g = 0; g < N / 16; ++g) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + g * 16, v);
    tc::tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[tid * N + g * 16 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int swz = argc > 1 ? atoi(argv[1]) : 3;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return printf("no entry point\n"), 1;
  Fn enc = (Fn)p;
  std::vector<float> A(K * M), B(K * N), want(M * N), got(M * N);
  for (int k = 0; k < K; ++k) {
    for (int m = 0; m < M; ++m) A[k * M + m] = (float)((k * 7 + m * 3 + (k * m) % 5) % 11 - 5);
    for (int n = 0; n < N; ++n) B[k * N + n] = (float)((k * 5 + n * 2 + (k * n) % 3) % 13 - 6);
  }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[k * M + m] * B[k * N + n];
      want[m * N + n] = s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, got.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  {
    cuuint64_t gd[2] = {M, K}, gs[1] = {M * 4};
    cuuint32_t box[2] = {32, K}, es[2] = {1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swz,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t gd2[2] = {N, K}, gs2[1] = {N * 4};
    CUresult r2 = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, gd2, gs2, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swz,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return printf("swizzle %d: encode failed %d %d\n", swz, (int)r, (int)r2), 1;
  }
  const int smem = 16384 + (4 + N / 32) * K * 128 + 16384 + 1024;
  cudaFuncSetAttribute((const void*)probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const uint32_t types[] = {2, 1, 0, 4, 6};
  const uint32_t offs[] = {4096, 1024, 512, 256, 128, 2048};
  const uint32_t ksteps[] = {1024, 512, 256, 128, 2048};
  int found = 0;
  for (uint32_t type : types)
    for (uint32_t lbo : offs)
      for (uint32_t sbo : offs)
        for (uint32_t ks : ksteps) {
          cudaMemset(dD, 0xff, got.size() * 4);
          probe<<<1, 128, smem>>>(tmA, tmB, type, lbo, sbo, ks, dD);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("swizzle %d type %u lbo %u sbo %u kstep %u: CUDA error %s -- stopping\n", swz, type, lbo, sbo, ks, cudaGetErrorString(e));
            fflush(stdout);
            return 2;
          }
          cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost);
          int bad = 0, zero = 0;
          for (int i = 0; i < M * N; ++i) {
            bad += got[i] != want[i];
            zero += got[i] == 0.f;
          }
          if (bad == 0) {
            printf("MATCH swizzle %d type %u lbo %u sbo %u kstep %u\n", swz, type, lbo, sbo, ks);
            ++found;
          } else if (bad < M * N / 2) {
            printf("partial swizzle %d type %u lbo %u sbo %u kstep %u: %d of %d wrong (%d zeros)\n", swz, type, lbo, sbo, ks, bad, M * N, zero);
          }
          fflush(stdout);
        }
  printf("swizzle %d: %d matching descriptor forms\n", swz, found);
  return 0;
}
