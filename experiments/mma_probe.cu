// Probe: mma.sync.m16n8k8.tf32 throughput on sm_100a (legacy tensor path), alone and next to packed FP32 FMAs,
// and the shared-memory wavefront cost of LDS.128 with warp-uniform / few-distinct / all-distinct addresses.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/mma_probe experiments/mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// mode 0: all warps MMA; mode 1: all warps FFMA2; mode 2: even warps MMA, odd warps FFMA2
template <int NACC>
__global__ void probe(float* out, long long* cyc, int iters, int mode) {
  const int warp = threadIdx.x >> 5;
  float d[NACC][4];
  uint32_t a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
  for (int t = 0; t < NACC; ++t) for (int i = 0; i < 4; ++i) d[t][i] = 0.f;
  float2 f[16];
  for (int i = 0; i < 16; ++i) f[i] = make_float2(threadIdx.x * 1e-3f, i);
  float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(1e-5f, 1e-6f);
  const bool do_mma = (mode == 0) || (mode == 2 && (warp & 1) == 0);
  __syncthreads();
  long long t0 = clock64();
  if (do_mma) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int t = 0; t < NACC; ++t) mma_tf32(d[t], a, b);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(*reinterpret_cast<unsigned long long*>(&f[i]))
                     : "l"(*reinterpret_cast<unsigned long long*>(&m)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0.f;
  for (int t = 0; t < NACC; ++t) for (int i = 0; i < 4; ++i) s += d[t][i];
  for (int i = 0; i < 16; ++i) s += f[i].x + f[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + warp] = t1 - t0;
}

// shared-memory LDS.128 patterns: 0 = warp-uniform address, 1 = 5 distinct 16B chunks (lane % 5), 2 = lane-distinct rows,
// 3 = LDS.32 lane-distinct consecutive
__global__ void lds_probe(float* out, long long* cyc, int iters, int pattern) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 1e-4f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int off;
  if (pattern == 0) off = 0;
  else if (pattern == 1) off = (lane % 5) * 4;
  else if (pattern == 2) off = lane * 4;
  else if (pattern == 4) off = ((lane / 5) % 5) * 4 + 64;  // like the 'A' load of the 4x4 tile map
  else off = lane;
  const float* base = sm + warp * 512 + off;
  float4 acc = make_float4(0, 0, 0, 0);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (pattern == 3) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((uint32_t)__cvta_generic_to_shared(base + ((u * 32 + it) & 255))));
        acc.x += v;
      } else {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"((uint32_t)__cvta_generic_to_shared(base + (((u + it) & 15) * 128 & 255))));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  if (lane == 0) cyc[blockIdx.x * 32 + warp] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 32 * 8);
  long long h[148 * 32];
  const int iters = 4096;
  for (int warps : {1, 2, 4, 8, 16}) {
    for (int mode = 0; mode < 3; ++mode) {
      probe<6><<<148, warps * 32>>>(out, cyc, iters, mode);
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < warps; ++w) if (h[w] > mx) mx = h[w];
      int mma_warps = mode == 0 ? warps : mode == 2 ? (warps + 1) / 2 : 0;
      int fma_warps = warps - mma_warps;
      double mma_per_clk = (double)mma_warps * iters * 6 / mx;      // per SM
      double ffma2_per_clk = (double)fma_warps * iters * 16 / mx;   // warp instructions per clk per SM
      printf("warps %2d mode %d: %lld cyc; mma.m16n8k8/clk/SM %.3f (%.0f MAC/clk/SM, %.1f TFLOP/s chip) ; FFMA2 warp-inst/clk/SM %.3f\n",
             warps, mode, mx, mma_per_clk, mma_per_clk * 1024, mma_per_clk * 1024 * 2 * 148 * 1.965e9 / 1e12, ffma2_per_clk);
    }
  }
  for (int nacc : {1, 2, 3}) {
    if (nacc == 1) probe<1><<<148, 256>>>(out, cyc, iters, 0);
    if (nacc == 2) probe<2><<<148, 256>>>(out, cyc, iters, 0);
    if (nacc == 3) probe<3><<<148, 256>>>(out, cyc, iters, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("8 warps, %d accumulators/warp: %.2f cyc per dependent-chain MMA slot\n", nacc, (double)h[0] / iters / nacc);
  }
  for (int pattern = 0; pattern < 5; ++pattern) {
    for (int warps : {1, 8}) {
      lds_probe<<<148, warps * 32, 32768>>>(out, cyc, 1024, pattern);
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < warps; ++w) if (h[w] > mx) mx = h[w];
      printf("lds pattern %d warps %d: %.2f cyc per warp-load per SM\n", pattern, warps, (double)mx / (1024.0 * 16 * warps));
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
