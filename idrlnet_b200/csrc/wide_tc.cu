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
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

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

__device__ __forceinline__ void release_tmem(uint64_t* bar, int lane) {
  tc::tc_fence_before();
  __syncwarp();
  if (lane == 0) tc::mbar_arrive(bar);
}

// Epilogue of one (P points, 128 units) accumulator tile: lane = hidden unit; the 16-point groups alternate between the two
// warp quads (warp = 0..7).  The channel structure (NF first-order, NS second-order channels) is a template parameter: with
// run-time counts the jets were predicate chains over seven-element arrays in local memory, and the row kernels, once fed by
// TMA, turned out to be bound by exactly this code (skipping it took the mlp_xl step from 319 to 238 ms).
// round_out: the values that are operands of a later tensor-core GEMM (H_l, zbar_l) are stored already rounded to TF32
// (cvt.rna, the rounding the staging code applies), so that tc_rows2_kernel can copy them into shared memory without touching
// them.  release: mbarrier to arrive on (one lane per warp) as soon as this warp has read its last accumulator columns
// (tc_rows2_kernel: the tensor core may overwrite the buffer).
template <int NF, int NS, int ACT, int EPI>
__device__ __forceinline__ void rows_epilogue(const WNet& net, uint32_t tmem, int warp, int lane, int n0, int j0, int m, size_t plane,
                                              int a_rows, const float* __restrict__ bias, float* __restrict__ Zio,
                                              float* __restrict__ Hout, bool round_out, uint64_t* release, int dbg = 0, float* __restrict__ colsum = nullptr) {
  constexpr int C = 1 + NF + NS;
  using T = TileCfg<C>;
  constexpr int P = T::P;
  const int H = net.H;
  const int j = j0 + (warp & 3) * 32 + lane;
  const bool jok = j < a_rows;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const float bj = (EPI != 1 && jok) ? bias[j] : 0.f;
  for (int g = warp >> 2; g < P / 16; g += 2) {
    const int nb = n0 + g * 16;
    const int left = jok ? m - nb : 0;  // points of this group that exist (>= 16: all of them)
    float* zp = Zio + (size_t)nb * H + j;
    if (EPI == 1) {
      // reverse layer: Z_{l-1} is fetched 4 points ahead (double-buffered registers) so the global loads of the next quad
      // are in flight while the current quad's adjoints are computed; the accumulator columns are read eight points at a
      // time (all sixteen at once cost 96 registers next to the 48 of the Z buffers: 570 bytes of spills per thread)
      float zin[2][C][4];
      auto fetch = [&](int quad, float (&dst)[C][4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ok = quad * 4 + i < left && !(dbg & 16);
#pragma unroll
          for (int c = 0; c < C; ++c) dst[c][i] = ok ? zp[c * plane + (size_t)(quad * 4 + i) * H] : 0.f;
        }
      };
      fetch(0, zin[0]);
      float dbsum = 0.f;  // db_{l-1} rides along: the group's sum of the unrounded value channel
#pragma unroll
      for (int oct = 0; oct < 2; ++oct) {
        float v[C][8];
#pragma unroll
        for (int c = 0; c < C; ++c) tc::tmem_ld8(tmem + lane_base + c * P + g * 16 + oct * 8, v[c]);
        tc::tmem_ld_wait();
        if (release != nullptr && oct == 1 && g + 2 >= P / 16) release_tmem(release, lane);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int quad = oct * 2 + q;
          if (quad + 1 < 4) fetch(quad + 1, zin[(quad + 1) & 1]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float hb[C], z[C], o[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
              hb[c] = v[c][q * 4 + i];
              z[c] = zin[quad & 1][c][i];
            }
            jet_bwd_t<ACT, NF, NS, false>(z, hb, o);
            if (quad * 4 + i < left && !(dbg & 8)) {
              dbsum += o[0];
#pragma unroll
              for (int c = 0; c < C; ++c) zp[c * plane + (size_t)(quad * 4 + i) * H] = round_out ? tc::to_tf32(o[c]) : o[c];
            }
          }
        }
      }
      if (colsum != nullptr && jok) colsum[(size_t)(nb >> 4) * H + j] = dbsum;
    } else {
      float v[C][16];
#pragma unroll
      for (int c = 0; c < C; ++c) tc::tmem_ld16(tmem + lane_base + c * P + g * 16, v[c]);
      tc::tmem_ld_wait();
      if (release != nullptr && g + 2 >= P / 16) release_tmem(release, lane);
      float* hp = Hout + (size_t)nb * H + j;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < left && !(dbg & 8)) {
          float z[C], o[C];
#pragma unroll
          for (int c = 0; c < C; ++c) z[c] = v[c][i];
          z[0] += bj;
          jet_fwd_t<ACT, NF, NS, false>(z, o);
#pragma unroll
          for (int c = 0; c < C; ++c) {
            zp[c * plane + (size_t)i * H] = z[c];
            hp[c * plane + (size_t)i * H] = round_out ? tc::to_tf32(o[c]) : o[c];
          }
        }
      }
    }
  }
}

template <int NF, int NS, int ACT, int EPI, bool X3>
__global__ void __launch_bounds__(256, 2) tc_rows_kernel(WNet net, const float* __restrict__ Bsrc, const float* __restrict__ Aw,
                                                      const float* __restrict__ bias, float* __restrict__ Zio,
                                                      float* __restrict__ Hout, int m, int Mc, int a_rows, int round_out, float* __restrict__ colsum) {
  constexpr int C = 1 + NF + NS;
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

  rows_epilogue<NF, NS, ACT, EPI>(net, tmem, warp, lane, n0, j0, m, plane, a_rows, bias, Zio, Hout, round_out != 0, nullptr, 0, colsum);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// ---- TF32 mode: persistent, warp-specialised row GEMM fed by TMA --------------------------------------------------
// tc_rows_kernel above spends nine tenths of a CTA's life outside the tensor core (profiles/r02/ncu_tc_rows_ns_xl.txt:
// a serial load -> convert -> store -> MMA -> wait chain per K chunk on ONE shared-memory stage, then an epilogue during
// which nothing is loaded; only the second CTA of the SM overlaps it).  In TF32 mode every GEMM operand is stored ALREADY
// ROUNDED by its producer (first_fwd_kernel, head_bwd_kernel, the epilogues here with round_out, the per-step weight
// copies), so staging is a pure copy and the TMA engine can do it: one persistent CTA per SM, 10 warps --
//   warp 8      one thread: per K chunk of 32 two tensor copies (cp.async.bulk.tensor, 128-byte-swizzled boxes: weights
//               {32 k, 128 units} from a 2-D map, jets {32 k, P points, C channels} from a 3-D map over the [c][n][k] planes;
//               out-of-range rows / k arrive as zeros) into a 4/5-stage ring, accounted on full[s] by expect_tx;
//   warp 9      one thread issues tcgen05.mma on SWIZZLE_128B K-major descriptors and releases stages (tcgen05.commit ->
//               empty[s]) and accumulators;
//   warps 0-7   epilogue of tile i (rows_epilogue) from one of TWO accumulator buffers in TMEM (columns 0 and 256),
//               while the copies and the tensor core already work on tile i+1.
// Tiles (point tile, unit tile) are dealt round-robin with the unit tile fastest, so the CTAs that share a point tile's
// B rows run at the same time (L2 hits).  Same products on the same rounded operands as tc_rows_kernel<.., false>.
// (A first version staged with 16-byte cp.async from two loader warps: 1.26x SLOWER than tc_rows_kernel -- 22 GB/s per SM,
// twice the L2 -> SM bytes of the tile (sector-granular 16-byte requests), experiments/tc_rows2_cpasync.cu.txt.)
template <int C, int EPI>
struct Rows2Cfg {
  static constexpr int N = TileCfg<C>::N;
  static constexpr int kKT = 32;
  static constexpr int kABytes = 128 * 128, kBBytes = N * 128, kStageBytes = kABytes + kBBytes;  // multiples of 1024
  // forward epilogue through shared memory + TMA stores: per warp quad two buffers of {Z, H} boxes [C][4 points][128 units]
  static constexpr bool kTmaEpi = (EPI == 0) && (C <= 6);
  static constexpr int kBoxFloats = C * 4 * 128;
  static constexpr int kStagingBytes = kTmaEpi ? 2 * 2 * 2 * kBoxFloats * 4 : 0;  // C x 16 KB
  static constexpr int kFit = (232448 - 2048 - kStagingBytes) / kStageBytes;
  static constexpr int kStages = kFit > 5 ? 5 : kFit;  // 5 x 40 KB at N = 192 without staging, 3 with it
  static_assert(kStages >= 3, "ring too shallow");
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024;  // + alignment slack
};

// Forward epilogue of one accumulator tile through shared memory (tc_rows2_kernel): the 16-point groups alternate between
// the two warp quads as in rows_epilogue; a quad of points at a time, the 128 threads of a warp quad write Z and H into
// boxes [C][4 points][128 units] (lanes = consecutive units: conflict-free, compile-time offsets) and one thread hands both
// boxes to the TMA engine (cp.async.bulk.tensor store; units >= H and points >= m are clipped by the tensor map).  Two
// buffers per warp quad: the store issued two quads ago must have been read (wait_group.read 1) before its buffer is reused.
// Replaces 12 STG.32 + their address arithmetic per thread and point (L1TEX was the busiest unit of the forward launch, 59 %).
template <int NF, int NS, int ACT>
__device__ __forceinline__ void rows_epilogue_fwd_tma(uint32_t tmem, int warp, int lane, int n0, int j0, int a_rows,
                                                      const float* __restrict__ bias, const CUtensorMap* tmZ, const CUtensorMap* tmH,
                                                      float* stage, uint32_t& phase, bool round_out, uint64_t* release) {
  constexpr int C = 1 + NF + NS;
  constexpr int P = TileCfg<C>::P;
  constexpr int kBox = C * 4 * 128;
  const int wq = warp & 3, grp = warp >> 2;
  const int jl = wq * 32 + lane, j = j0 + jl;
  const bool leader = (wq == 0 && lane == 0);
  const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
  const float bj = (j < a_rows) ? bias[j] : 0.f;
  float* mine = stage + (size_t)grp * (4 * kBox);
  for (int g = grp; g < P / 16; g += 2) {
    float v[C][16];
#pragma unroll
    for (int c = 0; c < C; ++c) tc::tmem_ld16(tmem + lane_base + c * P + g * 16, v[c]);
    tc::tmem_ld_wait();
    if (g + 2 >= P / 16) release_tmem(release, lane);
#pragma unroll
    for (int quad = 0; quad < 4; ++quad) {
      float* buf = mine + (phase & 1u) * (2 * kBox);
      if (leader) tc::bulk_wait_read<1>();
      tc::named_bar(1 + grp, 128);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float z[C], o[C];
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = v[c][quad * 4 + i];
        z[0] += bj;
        jet_fwd_t<ACT, NF, NS, false>(z, o);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          buf[(c * 4 + i) * 128 + jl] = z[c];
          buf[kBox + (c * 4 + i) * 128 + jl] = round_out ? tc::to_tf32(o[c]) : o[c];
        }
      }
      tc::fence_async_smem();
      tc::named_bar(1 + grp, 128);
      if (leader) {
        tc::tma_store_3d(tmZ, tc::smem_u32(buf), j0, n0 + g * 16 + quad * 4, 0);
        tc::tma_store_3d(tmH, tc::smem_u32(buf + kBox), j0, n0 + g * 16 + quad * 4, 0);
        tc::bulk_commit();
      }
      ++phase;
    }
  }
}

template <int NF, int NS, int ACT, int EPI>
__global__ void __launch_bounds__(320, 1) tc_rows2_kernel(WNet net, const __grid_constant__ CUtensorMap tmA,
                                                          const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmZ,
                                                          const __grid_constant__ CUtensorMap tmH, const float* __restrict__ bias,
                                                          float* __restrict__ Zio, float* __restrict__ Hout, int m, int Mc, int a_rows,
                                                          int tiles_j, int n_tiles, int round_out, int dbg, float* __restrict__ colsum) {
  constexpr int C = 1 + NF + NS;
  using T = TileCfg<C>;
  using R = Rows2Cfg<C, EPI>;
  constexpr int P = T::P, N = T::N, S = R::kStages, KT = R::kKT;
  const int H = net.H;
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  __shared__ uint64_t full[S], empty[S], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t plane = (size_t)Mc * H;
  const int nch = (H + KT - 1) / KT;
  const uint32_t ring = (tc::smem_u32(sm_raw) + 1023u) & ~1023u;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(&acc_full[a], 1);
      tc::mbar_init(&acc_empty[a], 8);
    }
    tc::fence_mbar_init();
  }
  if (warp == 9) tc::tmem_alloc(&tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp < 8) {  // ---- epilogue warps
    int ti = 0;
    uint32_t st_phase = 0;
    float* staging = reinterpret_cast<float*>(sm_raw + (ring - tc::smem_u32(sm_raw)) + S * R::kStageBytes);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
      const int a = ti & 1;
      if (!tc::mbar_wait(&acc_full[a], (ti >> 1) & 1)) __trap();
      tc::tc_fence_after();
      if (dbg & 4) {
        release_tmem(&acc_empty[a], lane);
        continue;
      }
      if constexpr (R::kTmaEpi) {
        rows_epilogue_fwd_tma<NF, NS, ACT>(tmem + a * 256, warp, lane, (t / tiles_j) * P, (t % tiles_j) * 128, a_rows, bias, &tmZ, &tmH,
                                           staging, st_phase, round_out != 0, &acc_empty[a]);
      } else {
        rows_epilogue<NF, NS, ACT, EPI>(net, tmem + a * 256, warp, lane, (t / tiles_j) * P, (t % tiles_j) * 128, m, plane, a_rows, bias,
                                        Zio, Hout, round_out != 0, &acc_empty[a], dbg, colsum);
      }
    }
    if constexpr (R::kTmaEpi) {
      if ((warp & 3) == 0 && lane == 0) tc::bulk_wait_read<0>();  // the boxes must outlive their stores
    }
  } else if (warp == 8) {  // ---- TMA producer
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmA);
      tc::tma_prefetch_desc(&tmB);
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int n0 = (t / tiles_j) * P, j0 = (t % tiles_j) * 128;
        for (int ch = 0; ch < nch; ++ch, ++it) {
          const int s = it % S;
          if (it >= S && !tc::mbar_wait(&empty[s], ((it / S) - 1) & 1)) __trap();
          const uint32_t sA = ring + (uint32_t)s * R::kStageBytes, sB = sA + R::kABytes;
          tc::mbar_expect_tx(&full[s], (uint32_t)R::kStageBytes);
          if (dbg & 2) {  // (timing experiment: no copies)
            asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(&full[s])), "r"((uint32_t)R::kStageBytes) : "memory");
            continue;
          }
          tc::tma_load_2d(sA, &tmA, ch * KT, j0, &full[s]);
          tc::tma_load_3d(sB, &tmB, ch * KT, n0, 0, &full[s]);
        }
      }
    }
  } else if (lane == 0) {  // ---- tensor-core issuer (warp 9)
    const uint32_t idesc = tc::idesc_tf32(N, 0, 0);
    int it = 0, ti = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
      const int a = ti & 1;
      if (ti >= 2 && !tc::mbar_wait(&acc_empty[a], ((ti >> 1) - 1) & 1)) __trap();
      tc::tc_fence_after();
      for (int ch = 0; ch < nch; ++ch, ++it) {
        const int s = it % S;
        if (!tc::mbar_wait(&full[s], (it / S) & 1)) __trap();
        tc::tc_fence_after();
        const uint32_t sA = ring + (uint32_t)s * R::kStageBytes, sB = sA + R::kABytes;
        if (!(dbg & 1)) {
#pragma unroll
          for (int ks = 0; ks < KT / 8; ++ks)
            tc::mma_tf32(tmem + a * 256, tc::smem_desc_sw128(sA + ks * 32), tc::smem_desc_sw128(sB + ks * 32), idesc,
                         (ch == 0 && ks == 0) ? 0u : 1u);
        }
        tc::mma_commit(&empty[s]);
      }
      tc::mma_commit(&acc_full[a]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 9) tc::tmem_dealloc(tmem, 512);
}

// elementwise TF32 rounding of a weight matrix (the per-step copy the forward row GEMMs of the TF32 mode read)
__global__ void round_copy_kernel(const float* __restrict__ W, float* __restrict__ Wr, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) Wr[i] = tc::to_tf32(W[i]);
}

__global__ void transpose_kernel(const float* __restrict__ W, float* __restrict__ WT, int H, int round) {
  __shared__ float t[32][33];
  const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    if (x < H && y0 + r < H) t[r][threadIdx.x] = W[(size_t)(y0 + r) * H + x];
  __syncthreads();
  const int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    if (xo < H && yo0 + r < H) WT[(size_t)(yo0 + r) * H + xo] = round ? tc::to_tf32(t[threadIdx.x][r]) : t[threadIdx.x][r];
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
bool encode_tiled(CUtensorMap* tm, int rank, const void* base, const cuuint64_t* gdim, const cuuint64_t* gstr, const cuuint32_t* box,
                  const cuuint32_t* est, int swizzle = 1) {  // 0 none, 1 SWIZZLE_128B, 2 SWIZZLE_128B_ATOM_32B
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static const Fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (Fn)p;
  }();
  if (!fn) return false;
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, box, est,
            CU_TENSOR_MAP_INTERLEAVE_NONE,
            swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int rows2_dbg() {  // timing experiments only (PINN_ROWS2_DBG: 1 no MMAs, 2 no loads, 4 no epilogue, 8 no epilogue stores, 16 no reverse Z loads): results are garbage
  static const int v = [] { const char* e = getenv("PINN_ROWS2_DBG"); return e ? atoi(e) : 0; }();
  return v;
}

template <int NF, int NS, int ACT, int EPI, bool X3>
int launch_rows_c(const WNet& n, const float* Bsrc, const float* Aw, const float* bias, float* Zio, float* Hout, int m, int Mc,
                  int a_rows, bool round_out, float* colsum, cudaStream_t st) {
  constexpr int C = 1 + NF + NS;
  using T = TileCfg<C>;
  if constexpr (!X3) {  // TF32 mode: operands are already rounded (tc_rows2_kernel, TMA-fed)
    using R = Rows2Cfg<C, EPI>;
    static bool attr_set2 = false;
    static int n_sm = 0;
    if (!attr_set2) {
      cudaError_t e = cudaFuncSetAttribute((const void*)tc_rows2_kernel<NF, NS, ACT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, R::kSmemBytes);
      if (e != cudaSuccess) return (int)e;
      int dev = 0;
      if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
      if ((e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
      attr_set2 = true;
    }
    CUtensorMap tmA, tmB;
    {
      // weights [a_rows][H] row-major: dims innermost first {k, unit}
      const cuuint64_t gdim[2] = {(cuuint64_t)n.H, (cuuint64_t)a_rows}, gstr[1] = {(cuuint64_t)n.H * 4};
      const cuuint32_t box[2] = {32, 128}, est[2] = {1, 1};
      if (!encode_tiled(&tmA, 2, Aw, gdim, gstr, box, est)) return PINN_E_UNSUPPORTED;
    }
    {
      // jet planes [C][Mc][H]: dims {k, point, channel}; only the m valid points are in bounds (the rest arrive as zeros)
      const cuuint64_t gdim[3] = {(cuuint64_t)n.H, (cuuint64_t)m, (cuuint64_t)C}, gstr[2] = {(cuuint64_t)n.H * 4, (cuuint64_t)Mc * n.H * 4};
      const cuuint32_t box[3] = {32, (cuuint32_t)T::P, (cuuint32_t)C}, est[3] = {1, 1, 1};
      if (!encode_tiled(&tmB, 3, Bsrc, gdim, gstr, box, est)) return PINN_E_UNSUPPORTED;
    }
    CUtensorMap tmZ = tmA, tmH = tmA;  // (placeholders unless the epilogue stores through TMA)
    if (R::kTmaEpi) {
      // output planes [C][Mc][H] as {unit, point, channel}; boxes {128 units, 4 points, C}, no swizzle; only valid elements are stored
      const cuuint64_t gdim[3] = {(cuuint64_t)n.H, (cuuint64_t)m, (cuuint64_t)C}, gstr[2] = {(cuuint64_t)n.H * 4, (cuuint64_t)Mc * n.H * 4};
      const cuuint32_t box[3] = {128, 4, (cuuint32_t)C}, est[3] = {1, 1, 1};
      if (!encode_tiled(&tmZ, 3, Zio, gdim, gstr, box, est, 0) || !encode_tiled(&tmH, 3, Hout, gdim, gstr, box, est, 0))
        return PINN_E_UNSUPPORTED;
    }
    const int tiles_j = (a_rows + 127) / 128, n_tiles = ((m + T::P - 1) / T::P) * tiles_j;
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    tc_rows2_kernel<NF, NS, ACT, EPI><<<grid, 320, R::kSmemBytes, st>>>(n, tmA, tmB, tmZ, tmH, bias, Zio, Hout, m, Mc, a_rows, tiles_j, n_tiles,
                                                                   round_out ? 1 : 0, rows2_dbg(), colsum);
    count_launch(1);
    return (int)cudaGetLastError();
  } else {
  constexpr int KT = X3 ? 32 : 64;
  const size_t smem = (size_t)(X3 ? 2 : 1) * (KT / 4) * (129 + T::N + 1) * 4 * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute((const void*)tc_rows_kernel<NF, NS, ACT, EPI, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid((m + T::P - 1) / T::P, (a_rows + 127) / 128);
  tc_rows_kernel<NF, NS, ACT, EPI, X3><<<grid, 256, smem, st>>>(n, Bsrc, Aw, bias, Zio, Hout, m, Mc, a_rows, round_out ? 1 : 0, colsum);
  count_launch(1);
  return (int)cudaGetLastError();
  }
}

}  // namespace

bool tc_encode_tiled(CUtensorMap* tm, int rank, const void* base, const cuuint64_t* gdim, const cuuint64_t* gstr, const cuuint32_t* box,
                     const cuuint32_t* est, int swizzle) {
  return encode_tiled(tm, rank, base, gdim, gstr, box, est, swizzle);
}

template <int ACT, int EPI>
int tc_launch_rows(const WNet& n, const float* Bsrc, const float* Aw, const float* bias, float* Zio, float* Hout, int m, int Mc,
                   bool round_out, float* colsum, cudaStream_t st) {
  const int a_rows = n.H;
  switch (n.NF * 4 + n.NS) {  // NS <= NF <= 3 (make_net)
#define CASE(F, S_)                                                                                                  \
  case F * 4 + S_:                                                                                                   \
    return n.precision == 2 ? launch_rows_c<F, S_, ACT, EPI, true>(n, Bsrc, Aw, bias, Zio, Hout, m, Mc, a_rows, false, colsum, st) \
                            : launch_rows_c<F, S_, ACT, EPI, false>(n, Bsrc, Aw, bias, Zio, Hout, m, Mc, a_rows, round_out, colsum, st);
    CASE(0, 0) CASE(1, 0) CASE(1, 1) CASE(2, 0) CASE(2, 1) CASE(2, 2) CASE(3, 0) CASE(3, 1) CASE(3, 2) CASE(3, 3)
#undef CASE
  }
  return PINN_E_UNSUPPORTED;
}

#define INST(A)                                                                                                              \
  template int tc_launch_rows<A, 0>(const WNet&, const float*, const float*, const float*, float*, float*, int, int, bool, float*, cudaStream_t); \
  template int tc_launch_rows<A, 1>(const WNet&, const float*, const float*, const float*, float*, float*, int, int, bool, float*, cudaStream_t);
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

int tc_transpose(const float* W, float* WT, int H, bool round, cudaStream_t st) {
  dim3 grid((H + 31) / 32, (H + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, st>>>(W, WT, H, round ? 1 : 0);
  count_launch(1);
  return (int)cudaGetLastError();
}

int tc_round_copy(const float* W, float* Wr, int n, cudaStream_t st) {
  round_copy_kernel<<<(n + 255) / 256, 256, 0, st>>>(W, Wr, n);
  count_launch(1);
  return (int)cudaGetLastError();
}

}  // namespace pinn
