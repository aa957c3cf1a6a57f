// Weight-gradient GEMM of the layer-streamed TF32 path, pipelined:
//   dW_l[j][k] (+)= sum over (channel c, point n) of zbar_l[c][n][j] * h_{l-1}[c][n][k].
//
// The contraction runs over points, so both operands are needed with K = point contiguous, i.e.
// transposed with respect to the [c][n][unit] planes the row GEMMs use.  `retile_kernel` writes
// each operand once in the tensor core's own shared-memory layout (no-swizzle K-major core
// matrices, TF32-rounded): tiles of 32 points x R units, tile(kb, ub) at ((kb * UB + ub) * 8 * R * 4)
// floats with element (unit row, point k) at ((k / 4) * R + row) * 4 + k % 4.  A tile is therefore one
// contiguous block that `cp.async.bulk` drops into shared memory unchanged, and the GEMM kernel has
// no staging code at all: one thread streams tiles through a 4-stage ring (mbarrier expect_tx),
// one thread issues tcgen05.mma (M = 128 units j, N = R_B units k, K = 8 points per instruction) and
// releases stages with tcgen05.commit, the accumulator stays in TMEM until the split's K range is
// exhausted, then 128 threads add it into the split's partial slot.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "tc_common.cuh"
#include "wide_internal.cuh"

namespace pinn {
namespace {

constexpr int kStages = 4;
constexpr int kKB = 32;  // points per K block (4 MMAs of K = 8)

// src: planes [c][n][j] (plane stride Mc * H), valid n < m, j < H.  grid (nblk, UB, C), 256 threads.
// colsum (optional, R = 128 only): colsum[nb][j] = sum over the block's points of the UNROUNDED value
// channel (c == 0) -- the bias gradient db_l = sum_n zbar_l[0][n][:] rides along with the re-tiling.
// X3: the tile is followed by its residual tile lo = tf32(x - tf32(x)) (3xTF32, see wide_tc.cu).
template <int R, bool X3>
__global__ void __launch_bounds__(256) retile_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int m, int Mc,
                                                     int nblk, int UB, float* __restrict__ colsum) {
  __shared__ float t[kKB][R + 1];
  __shared__ float red[256];
  const int nb = blockIdx.x, ub = blockIdx.y, c = blockIdx.z;
  const float* plane = src + (size_t)c * Mc * H;
  float acc = 0.f;
  for (int i = threadIdx.x; i < kKB * R; i += 256) {
    const int p = i / R, jj = i - p * R;
    const int n = nb * kKB + p, j = ub * R + jj;
    const float v = (n < m && j < H) ? plane[(size_t)n * H + j] : 0.f;
    acc += v;  // R divides 256: a thread always sees the same unit jj
    t[p][jj] = v;
  }
  if (R == 128 && colsum != nullptr && c == 0) {
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < 128 && ub * 128 + threadIdx.x < H)
      colsum[(size_t)nb * H + ub * 128 + threadIdx.x] = red[threadIdx.x] + red[threadIdx.x + 128];
  }
  __syncthreads();
  float* tile = dst + ((size_t)(c * nblk + nb) * UB + ub) * ((X3 ? 2 : 1) * kKB * R);
  for (int i = threadIdx.x; i < (kKB / 4) * R; i += 256) {
    const int kq = i / R, row = i - kq * R;
    const float4 x = make_float4(t[kq * 4][row], t[kq * 4 + 1][row], t[kq * 4 + 2][row], t[kq * 4 + 3][row]);
    const float4 hi = tc::to_tf32(x);
    *reinterpret_cast<float4*>(tile + (size_t)i * 4) = hi;
    if (X3)
      *reinterpret_cast<float4*>(tile + kKB * R + (size_t)i * 4) =
          tc::to_tf32(make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w));
  }
}

// grid (tiles_j * tiles_k, splits); 128 threads.  tz: A tiles (R = 128), th: B tiles (R = NB).
template <int NB, bool X3>
__global__ void __launch_bounds__(128, 1) tc_dw2_kernel(const float* __restrict__ tz, const float* __restrict__ th, int H, int kblocks,
                                                        int UBA, int UBB, int tiles_k, float* __restrict__ part, int dbg) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t full[kStages], empty[kStages], done;  // the first kS of each are used
  __shared__ uint32_t tmem_slot;
  constexpr int kS = X3 ? kStages / 2 : kStages;  // hi + lo tiles double a stage
  constexpr int kAFloats = (X3 ? 2 : 1) * kKB * 128, kBFloats = (X3 ? 2 : 1) * kKB * NB, kStageFloats = kAFloats + kBFloats;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tj = blockIdx.x / tiles_k, tk = blockIdx.x % tiles_k;
  const int per = (kblocks + gridDim.y - 1) / gridDim.y;
  const int kb0 = blockIdx.y * per, kb1 = (kb0 + per < kblocks) ? kb0 + per : kblocks;
  const int nit = kb1 - kb0;
  if (tid == 0) {
    for (int s = 0; s < kS; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(&done, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, NB);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (nit > 0) {
    if (warp == 0 && lane == 0) {  // producer: whole tiles by bulk copy
      for (int it = 0; it < nit; ++it) {
        const int s = it % kS;
        if (it >= kS && !tc::mbar_wait(&empty[s], ((it / kS) - 1) & 1)) __trap();
        float* sA = sm + (size_t)s * kStageFloats;
        const size_t kb = (size_t)(kb0 + it);
        tc::mbar_expect_tx(&full[s], (uint32_t)(kStageFloats * sizeof(float)));
        if (dbg & 2) {
          asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(&full[s])), "r"((uint32_t)(kStageFloats * sizeof(float))) : "memory");
          continue;
        }
        tc::bulk_copy_g2s(sA, tz + (kb * UBA + tj) * kAFloats, kAFloats * sizeof(float), &full[s]);
        tc::bulk_copy_g2s(sA + kAFloats, th + (kb * UBB + tk) * kBFloats, kBFloats * sizeof(float), &full[s]);
      }
    } else if (warp == 1 && lane == 0) {  // tensor-core issuer
      const uint32_t idesc = tc::idesc_tf32(NB, 0, 0);
      for (int it = 0; it < nit; ++it) {
        const int s = it % kS;
        if (!tc::mbar_wait(&full[s], (it / kS) & 1)) __trap();
        tc::tc_fence_after();
        const uint32_t a0 = tc::smem_u32(sm + (size_t)s * kStageFloats), b0 = a0 + kAFloats * sizeof(float);
#pragma unroll
        for (int ks = 0; ks < kKB / 8; ++ks) {
          if (dbg & 1) break;
          const uint64_t da = tc::smem_desc(a0 + ks * 2 * (128 * 16), 128 * 16, 128);
          const uint64_t db = tc::smem_desc(b0 + ks * 2 * (NB * 16), NB * 16, 128);
          if (X3) {  // lo*hi + hi*lo + hi*hi
            const uint64_t dal = tc::smem_desc(a0 + kKB * 128 * 4 + ks * 2 * (128 * 16), 128 * 16, 128);
            const uint64_t dbl = tc::smem_desc(b0 + kKB * NB * 4 + ks * 2 * (NB * 16), NB * 16, 128);
            tc::mma_tf32(tmem, dal, db, idesc, (it > 0 || ks > 0) ? 1u : 0u);
            tc::mma_tf32(tmem, da, dbl, idesc, 1u);
            tc::mma_tf32(tmem, da, db, idesc, 1u);
          } else {
            tc::mma_tf32(tmem, da, db, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          }
        }
        tc::mma_commit(&empty[s]);  // the stage is free once these MMAs have read it
      }
      tc::mma_commit(&done);
    }
    __syncwarp();
    if (!tc::mbar_wait(&done, 0)) __trap();
    tc::tc_fence_after();
    const int j = tj * 128 + tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float* out = part + (size_t)blockIdx.y * H * H;
    // flush: the split's slot is touched by this CTA only, once per launch, so `+=` is done as fire-and-forget vector
    // reductions (red.global.add.v4.f32: one add per address per launch, launches in stream order -> the same sums as a
    // read-modify-write, without its 16 serial round trips to a slot that is in DRAM by now: that chain was 54 of the 70 ms
    // this kernel took per mlp_xl step)
    for (int g = 0; g < NB / 16; ++g) {
      if (dbg & 4) break;
      float v[16];
      tc::tmem_ld16(tmem + lane_base + g * 16, v);
      tc::tmem_ld_wait();
      if (j < H) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const int k = tk * NB + g * 16 + i;
          if (k < H)  // H % 4 == 0: a quad is inside or outside as a whole
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + (size_t)j * H + k), "f"(v[i]), "f"(v[i + 1]),
                         "f"(v[i + 2]), "f"(v[i + 3])
                         : "memory");
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, NB);
}

// db slot (block 0, row 0 of the [kBlocks][8][H] partials) += sum over point blocks: 32 segments of the block range in index
// order each, combined in segment order (a fixed order).  One CTA = 32 units x 32 segments, so H / 32 CTAs share the work (one
// serial chain of `nblk` loads per unit took 31 us per call at 2 048 blocks; a chunk of allen_cahn has that many).
__global__ void __launch_bounds__(1024) colsum_finish_kernel(const float* __restrict__ colsum, int nblk, int H, float* __restrict__ db_slot) {
  __shared__ float red[32][33];
  const int jl = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + jl;
  const int per = (nblk + 31) / 32, nb0 = seg * per, nb1 = (nb0 + per < nblk) ? nb0 + per : nblk;
  float s = 0.f;
  if (j < H)
    for (int nb = nb0; nb < nb1; ++nb) s += colsum[(size_t)nb * H + j];
  red[seg][jl] = s;
  __syncthreads();
  if (seg == 0 && j < H) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 32; ++g) t += red[g][jl];
    db_slot[j] += t;
  }
}

// ---- TF32 mode: dW operands straight from the planes -------------------------------------------------------------------
// dW_l[j][k] += sum over (c, n) of zbar_l[c][n][j] * h_{l-1}[c][n][k]: both operands are wanted with the point index as K, and the
// planes [c][n][unit] have the UNIT index contiguous -- "MN-major" operands.  For 32-bit elements tcgen05 reads those in the
// 128-byte swizzle with 32-byte atoms (descriptor layout type 1; TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): groups of 4 K-rows x
// 128 bytes (32 units), SBO = 512 bytes between the two 4-row groups of one K = 8 MMA, LBO between 32-unit groups -- what a TMA
// box {32 units, KB points, 1 channel} leaves in shared memory with LBO = KB * 128.  The form was found by exhaustive search on the
// hardware (experiments/mn_major_probe.cu, profiles/r02/mn_major_probe.txt): the plain SWIZZLE_128B (16-byte atoms) matches no
// descriptor form for TF32.  So the re-tiling pass (retile_kernel: read + write of both operands, 26 % of an mlp_xl TF32 step)
// disappears: one thread issues 4 + NB / 32 boxes per 32-point K block into a 4-stage ring, one thread issues 4 MMAs (K = 8
// points each: start address + ks * 1024 bytes) per block.
// In TF32 mode the planes already hold TF32-rounded values (wide_tc.cu); rows >= m arrive as zeros, units >= H as zeros.
// db_l comes from the producers of zbar_l (head_bwd_kernel, rows_epilogue) instead of the re-tiling pass.
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;  // 128-byte swizzle, 32-byte atoms
  return d;
}

template <int NB>
__global__ void __launch_bounds__(128, 1) tc_dw3_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmH,
                                                        int H, int nblk, int kblocks, int tiles_k, float* __restrict__ part, int dbg) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  __shared__ uint64_t full[kStages], empty[kStages], done;
  __shared__ uint32_t tmem_slot;
  constexpr int kABytes = 4 * kKB * 128, kBBytes = (NB / 32) * kKB * 128, kStageBytes = kABytes + kBBytes;  // 16 KB + 32 KB
  constexpr uint32_t kLBO = kKB * 128;  // bytes between 32-unit groups
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tj = blockIdx.x / tiles_k, tk = blockIdx.x % tiles_k;
  const int per = (kblocks + gridDim.y - 1) / gridDim.y;
  const int kb0 = blockIdx.y * per, kb1 = (kb0 + per < kblocks) ? kb0 + per : kblocks;
  const int nit = kb1 - kb0;
  const uint32_t ring = (tc::smem_u32(sm_raw) + 1023u) & ~1023u;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(&done, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, NB);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (nit > 0) {
    if (warp == 0 && lane == 0) {  // producer
      tc::tma_prefetch_desc(&tmZ);
      tc::tma_prefetch_desc(&tmH);
      for (int it = 0; it < nit; ++it) {
        const int s = it % kStages;
        if (it >= kStages && !tc::mbar_wait(&empty[s], ((it / kStages) - 1) & 1)) __trap();
        const int kb = kb0 + it, c = kb / nblk, n0 = (kb - c * nblk) * kKB;
        const uint32_t sA = ring + (uint32_t)s * kStageBytes, sB = sA + kABytes;
        tc::mbar_expect_tx(&full[s], (uint32_t)kStageBytes);
        if (dbg & 2) {
          asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(&full[s])), "r"((uint32_t)kStageBytes) : "memory");
          continue;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) tc::tma_load_3d(sA + g * kLBO, &tmZ, tj * 128 + g * 32, n0, c, &full[s]);
#pragma unroll
        for (int g = 0; g < NB / 32; ++g) tc::tma_load_3d(sB + g * kLBO, &tmH, tk * NB + g * 32, n0, c, &full[s]);
      }
    } else if (warp == 1 && lane == 0) {  // tensor-core issuer
      const uint32_t idesc = tc::idesc_tf32(NB, 1, 1);  // both operands MN-major
      for (int it = 0; it < nit; ++it) {
        const int s = it % kStages;
        if (!tc::mbar_wait(&full[s], (it / kStages) & 1)) __trap();
        tc::tc_fence_after();
        const uint32_t sA = ring + (uint32_t)s * kStageBytes, sB = sA + kABytes;
#pragma unroll
        for (int ks = 0; ks < kKB / 8; ++ks) {
          if (dbg & 1) break;
          tc::mma_tf32(tmem, smem_desc_mn_sw128(sA + ks * 1024, kLBO), smem_desc_mn_sw128(sB + ks * 1024, kLBO), idesc,
                       (it > 0 || ks > 0) ? 1u : 0u);
        }
        tc::mma_commit(&empty[s]);
      }
      tc::mma_commit(&done);
    }
    __syncwarp();
    if (!tc::mbar_wait(&done, 0)) __trap();
    tc::tc_fence_after();
    const int j = tj * 128 + tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float* out = part + (size_t)blockIdx.y * H * H;
    for (int g = 0; g < NB / 16; ++g) {  // flush by vector reductions (see tc_dw2_kernel)
      if (dbg & 4) break;
      float v[16];
      tc::tmem_ld16(tmem + lane_base + g * 16, v);
      tc::tmem_ld_wait();
      if (j < H) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const int k = tk * NB + g * 16 + i;
          if (k < H)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + (size_t)j * H + k), "f"(v[i]), "f"(v[i + 1]),
                         "f"(v[i + 2]), "f"(v[i + 3])
                         : "memory");
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, NB);
}

// FP32-grade (3xTF32) mode, same idea: the planes hold FP32 values, so the split into hi = tf32(x) and lo = tf32(x - hi) happens
// in shared memory.  TMA delivers the raw boxes (full[s]); four converter warps rewrite them in place as hi and write lo into the
// second half of the stage -- elementwise, so the swizzled layout is simply carried over -- then publish the stage to the MMA
// thread (conv[s], one arrival per warp after fence.proxy.async); 12 MMAs per K block (lo*hi + hi*lo + hi*hi, the order of
// tc_dw2_kernel<., true>).  Per stage the converters move 96 KB through shared memory (~770 cycles at 128 B/clk) next to ~840
// cycles of MMAs at NB = 128: the pass hides behind the tensor pipe, and the re-tiled hi / lo copies (HBM write + read of twice
// the operands) are gone.  Stage = [A hi | B hi | A lo | B lo] = 64 KB, three stages.
constexpr int kStagesX3 = 3;

template <int NB>
__global__ void __launch_bounds__(192, 1) tc_dw3x_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmH,
                                                         int H, int nblk, int kblocks, int tiles_k, float* __restrict__ part) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  __shared__ uint64_t full[kStagesX3], conv[kStagesX3], empty[kStagesX3], done;
  __shared__ uint32_t tmem_slot;
  constexpr int kABytes = 4 * kKB * 128, kBBytes = (NB / 32) * kKB * 128, kRawBytes = kABytes + kBBytes, kStageBytes = 2 * kRawBytes;
  constexpr uint32_t kLBO = kKB * 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tj = blockIdx.x / tiles_k, tk = blockIdx.x % tiles_k;
  const int per = (kblocks + gridDim.y - 1) / gridDim.y;
  const int kb0 = blockIdx.y * per, kb1 = (kb0 + per < kblocks) ? kb0 + per : kblocks;
  const int nit = kb1 - kb0;
  const uint32_t ring = (tc::smem_u32(sm_raw) + 1023u) & ~1023u;
  if (tid == 0) {
    for (int s = 0; s < kStagesX3; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&conv[s], 4);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(&done, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, NB);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (nit > 0) {
    if (warp == 0) {
      if (lane == 0) {  // producer
        tc::tma_prefetch_desc(&tmZ);
        tc::tma_prefetch_desc(&tmH);
        for (int it = 0; it < nit; ++it) {
          const int s = it % kStagesX3;
          if (it >= kStagesX3 && !tc::mbar_wait(&empty[s], ((it / kStagesX3) - 1) & 1)) __trap();
          const int kb = kb0 + it, c = kb / nblk, n0 = (kb - c * nblk) * kKB;
          const uint32_t sA = ring + (uint32_t)s * kStageBytes, sB = sA + kABytes;
          tc::mbar_expect_tx(&full[s], (uint32_t)kRawBytes);
#pragma unroll
          for (int g = 0; g < 4; ++g) tc::tma_load_3d(sA + g * kLBO, &tmZ, tj * 128 + g * 32, n0, c, &full[s]);
#pragma unroll
          for (int g = 0; g < NB / 32; ++g) tc::tma_load_3d(sB + g * kLBO, &tmH, tk * NB + g * 32, n0, c, &full[s]);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {  // tensor-core issuer
        const uint32_t idesc = tc::idesc_tf32(NB, 1, 1);
        for (int it = 0; it < nit; ++it) {
          const int s = it % kStagesX3;
          if (!tc::mbar_wait(&conv[s], (it / kStagesX3) & 1)) __trap();
          tc::tc_fence_after();
          const uint32_t sA = ring + (uint32_t)s * kStageBytes, sB = sA + kABytes;
#pragma unroll
          for (int ks = 0; ks < kKB / 8; ++ks) {
            const uint64_t da = smem_desc_mn_sw128(sA + ks * 1024, kLBO), db = smem_desc_mn_sw128(sB + ks * 1024, kLBO);
            const uint64_t dal = smem_desc_mn_sw128(sA + kRawBytes + ks * 1024, kLBO), dbl = smem_desc_mn_sw128(sB + kRawBytes + ks * 1024, kLBO);
            tc::mma_tf32(tmem, dal, db, idesc, (it > 0 || ks > 0) ? 1u : 0u);
            tc::mma_tf32(tmem, da, dbl, idesc, 1u);
            tc::mma_tf32(tmem, da, db, idesc, 1u);
          }
          tc::mma_commit(&empty[s]);
        }
        tc::mma_commit(&done);
      }
    } else {  // converter warps 2 .. 5, then the flush (TMEM lane quarter = warp & 3)
      const int ct = tid - 64;
      for (int it = 0; it < nit; ++it) {
        const int s = it % kStagesX3;
        if (!tc::mbar_wait(&full[s], (it / kStagesX3) & 1)) __trap();
        const uint32_t base = ring + (uint32_t)s * kStageBytes + (uint32_t)ct * 16;
#pragma unroll 4
        for (int i = 0; i < kRawBytes / (128 * 16); ++i) {
          const uint32_t a = base + (uint32_t)i * (128 * 16);
          float4 x;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(a) : "memory");
          const float4 hi = tc::to_tf32(x);
          const float4 lo = tc::to_tf32(make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w));
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a + (uint32_t)kRawBytes), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w)
                       : "memory");
        }
        tc::fence_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&conv[s]);
      }
      if (!tc::mbar_wait(&done, 0)) __trap();
      tc::tc_fence_after();
      const int q = warp & 3;
      const int j = tj * 128 + q * 32 + lane;
      const uint32_t lane_base = (uint32_t)(q * 32) << 16;
      float* out = part + (size_t)blockIdx.y * H * H;
      for (int g = 0; g < NB / 16; ++g) {  // flush by vector reductions (see tc_dw2_kernel)
        float v[16];
        tc::tmem_ld16(tmem + lane_base + g * 16, v);
        tc::tmem_ld_wait();
        if (j < H) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int k = tk * NB + g * 16 + i;
            if (k < H)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + (size_t)j * H + k), "f"(v[i]), "f"(v[i + 1]),
                           "f"(v[i + 2]), "f"(v[i + 3])
                           : "memory");
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, NB);
}

int dw2_dbg() {  // timing experiments only (PINN_DW2_DBG: 1 no MMAs, 2 no copies, 4 no flush): results are garbage
  static const int v = [] { const char* e = getenv("PINN_DW2_DBG"); return e ? atoi(e) : 0; }();
  return v;
}

int rows_b(int H) { return H > 128 ? 256 : 128; }

template <int R, bool X3>
int launch_dw2(const float* Zb, const float* Hb, int H, int C, int m, int Mc, float* part, int n_splits, float* scratch,
               float* db_slot, cudaStream_t st) {
  constexpr int F = X3 ? 2 : 1;
  const int nblk = (m + kKB - 1) / kKB;
  const int UBA = (H + 127) / 128, UBB = (H + R - 1) / R;
  const size_t nblk_alloc = (size_t)(Mc + kKB - 1) / kKB;
  float* tz = scratch;
  float* th = tz + (size_t)F * C * nblk_alloc * UBA * (kKB * 128);
  float* colsum = th + (size_t)F * C * nblk_alloc * UBB * (kKB * R);
  retile_kernel<128, X3><<<dim3(nblk, UBA, C), 256, 0, st>>>(Zb, tz, H, m, Mc, nblk, UBA, colsum);
  colsum_finish_kernel<<<(H + 31) / 32, 1024, 0, st>>>(colsum, nblk, H, db_slot);
  retile_kernel<R, X3><<<dim3(nblk, UBB, C), 256, 0, st>>>(Hb, th, H, m, Mc, nblk, UBB, nullptr);
  const int tiles = UBA * UBB, kblocks = C * nblk;
  int splits = kBlocks / tiles;
  if (splits < 1) splits = 1;
  if (splits > n_splits) splits = n_splits;
  if (splits > kblocks) splits = kblocks;
  const size_t smem = (size_t)kStages * (kKB * 128 + kKB * R) * sizeof(float);  // X3: half the stages, twice the size
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute((const void*)tc_dw2_kernel<R, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    set = true;
  }
  tc_dw2_kernel<R, X3><<<dim3(tiles, splits), 128, smem, st>>>(tz, th, H, kblocks, UBA, UBB, UBB, part, dw2_dbg());
  count_launch(4);
  return (int)cudaGetLastError();
}

template <int NB>
int launch_dw3(const float* Zb, const float* Hb, int H, int C, int m, int Mc, float* part, int n_splits, cudaStream_t st) {
  const int nblk = (m + kKB - 1) / kKB;
  const int UBA = (H + 127) / 128, UBB = (H + NB - 1) / NB;
  CUtensorMap tmZ, tmH;
  const cuuint64_t gdim[3] = {(cuuint64_t)H, (cuuint64_t)m, (cuuint64_t)C}, gstr[2] = {(cuuint64_t)H * 4, (cuuint64_t)Mc * H * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)kKB, 1}, est[3] = {1, 1, 1};
  if (!tc_encode_tiled(&tmZ, 3, Zb, gdim, gstr, box, est, 2) || !tc_encode_tiled(&tmH, 3, Hb, gdim, gstr, box, est, 2))
    return PINN_E_UNSUPPORTED;
  const int tiles = UBA * UBB, kblocks = C * nblk;
  int splits = kBlocks / tiles;
  if (splits < 1) splits = 1;
  if (splits > n_splits) splits = n_splits;
  if (splits > kblocks) splits = kblocks;
  const int smem = kStages * (4 + NB / 32) * kKB * 128 + 1024;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute((const void*)tc_dw3_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    set = true;
  }
  tc_dw3_kernel<NB><<<dim3(tiles, splits), 128, smem, st>>>(tmZ, tmH, H, nblk, kblocks, UBB, part, dw2_dbg());
  count_launch(1);
  return (int)cudaGetLastError();
}

int launch_dw3x(const float* Zb, const float* Hb, int H, int C, int m, int Mc, float* part, int n_splits, cudaStream_t st) {
  constexpr int NB = 128;
  const int nblk = (m + kKB - 1) / kKB;
  const int UBA = (H + 127) / 128, UBB = (H + NB - 1) / NB;
  CUtensorMap tmZ, tmH;
  const cuuint64_t gdim[3] = {(cuuint64_t)H, (cuuint64_t)m, (cuuint64_t)C}, gstr[2] = {(cuuint64_t)H * 4, (cuuint64_t)Mc * H * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)kKB, 1}, est[3] = {1, 1, 1};
  if (!tc_encode_tiled(&tmZ, 3, Zb, gdim, gstr, box, est, 2) || !tc_encode_tiled(&tmH, 3, Hb, gdim, gstr, box, est, 2))
    return PINN_E_UNSUPPORTED;
  const int tiles = UBA * UBB, kblocks = C * nblk;
  int splits = kBlocks / tiles;
  if (splits < 1) splits = 1;
  if (splits > n_splits) splits = n_splits;
  if (splits > kblocks) splits = kblocks;
  const int smem = kStagesX3 * 2 * (4 + NB / 32) * kKB * 128 + 1024;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute((const void*)tc_dw3x_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    set = true;
  }
  tc_dw3x_kernel<NB><<<dim3(tiles, splits), 192, smem, st>>>(tmZ, tmH, H, nblk, kblocks, UBB, part);
  count_launch(1);
  return (int)cudaGetLastError();
}

}  // namespace

// floats of scratch for the two re-tiled operands of one layer (+ the per-block column sums)
int64_t tc_dw2_scratch_floats(int H, int C, int Mc, bool x3) {
  const int64_t nblk = (Mc + kKB - 1) / kKB;
  const int R = rows_b(H), f = x3 ? 2 : 1;
  return f * ((int64_t)C * nblk * ((H + 127) / 128) * (kKB * 128) + (int64_t)C * nblk * ((H + R - 1) / R) * (kKB * R)) + nblk * H;
}

int tc_launch_dw2(const float* Zb, const float* Hb, int H, int C, int m, int Mc, float* part, int n_splits, float* scratch,
                  float* db_slot, bool x3, cudaStream_t st) {
  if (rows_b(H) == 256)
    return x3 ? launch_dw2<256, true>(Zb, Hb, H, C, m, Mc, part, n_splits, scratch, db_slot, st)
              : launch_dw2<256, false>(Zb, Hb, H, C, m, Mc, part, n_splits, scratch, db_slot, st);
  return x3 ? launch_dw2<128, true>(Zb, Hb, H, C, m, Mc, part, n_splits, scratch, db_slot, st)
            : launch_dw2<128, false>(Zb, Hb, H, C, m, Mc, part, n_splits, scratch, db_slot, st);
}

int tc_launch_dw3(const float* Zb, const float* Hb, int H, int C, int m, int Mc, float* part, int n_splits, bool x3, cudaStream_t st) {
  if (x3) return launch_dw3x(Zb, Hb, H, C, m, Mc, part, n_splits, st);
  return rows_b(H) == 256 ? launch_dw3<256>(Zb, Hb, H, C, m, Mc, part, n_splits, st) : launch_dw3<128>(Zb, Hb, H, C, m, Mc, part, n_splits, st);
}

int tc_colsum_finish(const float* colsum, int nrows, int H, float* db_slot, cudaStream_t st) {
  colsum_finish_kernel<<<(H + 31) / 32, 1024, 0, st>>>(colsum, nrows, H, db_slot);
  count_launch(1);
  return (int)cudaGetLastError();
}

// PINN_DW_DIRECT=0 keeps the re-tiling pass of the tensor modes (A/B measurements)
bool tc_dw_direct() {
  static const bool v = [] { const char* e = getenv("PINN_DW_DIRECT"); return !(e && e[0] == '0'); }();
  return v;
}

}  // namespace pinn
