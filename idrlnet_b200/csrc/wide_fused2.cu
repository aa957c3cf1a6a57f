// Wide nets, H <= 128, 2..4 hidden layers, TF32: the whole training step on chip, second generation.
//
// Same mathematics and the same partial-slot protocol as wide_fused.cu (so wide_finalize is shared), reorganised
// around what bounded that kernel (profiles/r01/ncu_wide_fused_ldc_v4_16warps.txt: a serial
// barrier -> MMA -> wait -> epilogue chain with ONE tile in flight, 43 KB of L2 traffic per point, 47 % of the
// shared-memory wavefronts bank-conflicted):
//
//   * the CTA is split into two HALVES of 8 compute warps; each half owns its own tile of P = 8 points and they
//     run half a phase apart, so one half's activation epilogue overlaps the other half's tensor-core GEMM;
//   * a 17th warp is the only one that talks to the TMA engine and the tensor cores: it streams the weight tiles
//     (double-buffered, one fetch serves BOTH halves -> half the weight traffic per point) and issues every
//     tcgen05.mma; compute warps only arrive on "operands ready" mbarriers and wait on "accumulator ready" ones;
//   * zbar_l, the A operand of the weight-gradient GEMM dW_l += zbar_l^T h_{l-1}, never goes to shared memory: the
//     epilogue thread that owns a hidden unit (= TMEM lane) writes it with tcgen05.st and the MMA reads it from TMEM
//     (the A-from-TMEM form); h_{l-1}, its B operand, is written with 64-bit stores (two points of a channel);
//   * the pre-activation jet z_l needed twice by the reverse sweep (adjoint of layer l, then h_l for dW_{l+1}) is
//     read from the L2-resident stash once and kept in registers across the step.
//
// TMEM columns: [0, N) and [N, 2N) accumulators of the two halves, [2N, 4N) their zbar operands, then (L-1) x N2
// weight-gradient accumulators that live for the whole kernel (N = C * 8, N2 = H rounded up to 16).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "tc_common.cuh"
#include "wide_internal.cuh"

namespace pinn {
namespace {

constexpr int kP = 8;          // points per half-tile
constexpr int kMaxHid = 4;
constexpr int kHalf = 256;     // compute threads per half
constexpr int kThreads = 2 * kHalf + 32;

struct Fused2Args {
  pinn_domain dom;
  const float* xs[PINN_MAX_IN];
  int64_t n_points;
  float* stash;        // [cta][half][(L-1)][C][kP][128]
  const float* wprep;  // [2*(l-2) + dir][Kq][128][4]: W_l (dir 0) / W_l^T (dir 1), TF32-rounded, K-major canonical tiles
  float* dw_part;      // [(l-2)][n_splits][H*H]
  float* db_part;      // [(l-2)][kBlocks][8][H]
  float* first_part;   // [kBlocks][8][H]
  float* head_part;    // [kBlocks][O*H + O + 1]
  int n_splits;
  int loss_only;
};

__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float (&v)[2]) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
  v[0] = __uint_as_float(r0);
  v[1] = __uint_as_float(r1);
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, float a, float b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b))
               : "memory");
}

struct Smem2 {  // offsets in floats
  int sW, sX, sHt, sY, sP, sXc, sG, sB, sW1, sWo, total;
};
constexpr int kSl = 6;  // unit slices of the SIMT output layer
__host__ __device__ inline Smem2 smem2_layout(int H, int C) {
  const int H8 = (H + 7) & ~7, Kq = H8 / 4, N = C * kP, N2 = (H + 15) & ~15;
  Smem2 s;
  int off = 0;
  s.sW = off; off += 2 * Kq * 128 * 4;            // two weight slots [Kq][128][4]
  s.sX = off; off += 2 * Kq * (N + 1) * 4;         // per half: B operand, K = unit     [Kq][N + 1][4]
  s.sHt = off; off += 2 * (N / 4) * (N2 + 1) * 4;  // per half: B operand of dW, K = (channel, point)  [N/4][N2 + 1][4]
  s.sY = off; off += 2 * 2 * N * 4;                // per half: y[N][4], ybar[N][4]
  s.sP = off; off += 2 * kSl * N * 4;              // per half: partial sums of the output layer
  s.sXc = off; off += 2 * kP * 4;                  // per half: coordinates
  s.sG = off; off += 2 * PINN_MAX_EQ * kP;         // per half: loss seeds
  s.sB = off; off += kMaxHid * 128;                // biases b_1 .. b_L by unit
  s.sW1 = off; off += 4 * 128;                     // W1[d][unit]
  s.sWo = off; off += 4 * 128;                     // Wout[o][unit]
  s.total = off;
  return s;
}

template <int NF, int NS, int ACT>
__global__ void __launch_bounds__(kThreads, 1) wide_fused2_kernel(WNet net, const __grid_constant__ Fused2Args fa) {
  constexpr int C = 1 + NF + NS;
  constexpr int N = C * kP;
  const int H = net.H, L = net.L, O = net.n_out;
  const int H8 = (H + 7) & ~7, Kq = H8 / 4, N2 = (H + 15) & ~15;
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t b_wfull[2], b_wfree[2], b_ready[2], b_done[2];
  __shared__ uint32_t tmem_slot;
  const Smem2 lay = smem2_layout(H, C);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = warp == 16;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&b_wfull[i], 1);
      tc::mbar_init(&b_wfree[i], 1);
      tc::mbar_init(&b_ready[i], kHalf);
      tc::mbar_init(&b_done[i], 1);
    }
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  for (int i = tid; i < 4 * 128; i += kThreads) {
    const int o = i >> 7, k = i & 127;
    sm[lay.sWo + i] = (o < O && k < H) ? net.W[L][o * H + k] : 0.f;
    sm[lay.sW1 + i] = (o < net.n_in && k < H) ? net.W[0][k * net.n_in + o] : 0.f;
    sm[lay.sB + i] = (o < L && k < H) ? net.b[o][k] : 0.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int64_t n_tiles = (fa.n_points + kP - 1) / kP;       // half-tiles
  const int64_t n_iter = (n_tiles + 2 * gridDim.x - 1) / (2 * gridDim.x);  // every CTA runs the same number of rounds
  const int n_seq = fa.loss_only ? (L - 1) : 2 * (L - 1);    // GEMM phases per tile: W_2..W_L, then W_L^T..W_2^T
  const uint32_t wbytes = (uint32_t)(Kq * 128 * 4 * sizeof(float));
  const uint32_t colD = 0, colA = 2 * N, colW = 4 * N;        // TMEM column bases

  // per-unit gradient accumulators of the compute threads (summed over the 4 thread groups of a unit at the end)
  float g_w1[4] = {0.f, 0.f, 0.f, 0.f}, g_b[kMaxHid + 1] = {0.f, 0.f, 0.f, 0.f, 0.f}, g_wo[4] = {0.f, 0.f, 0.f, 0.f};
  float g_bo_r = 0.f, loss = 0.f;  // threads (o, p) = tl < 32 / (e, p) = tl < 64 of each half

  // ================================================================ issuer warp ======================
  if (issuer) {
    if (lane == 0) {
      const uint32_t idesc_x = tc::idesc_tf32(N, 0, 0), idesc_w = tc::idesc_tf32(N2, 0, 0);
      const int64_t n_phase = n_iter * n_seq;
      uint32_t ph_ready[2] = {0, 0}, ph_wfull[2] = {0, 0}, ph_wfree[2] = {0, 0};
      auto weight_of = [&](int64_t g) -> const float* {  // weight tile of global phase g
        const int sidx = (int)(g % n_seq);
        const int l = sidx < L - 1 ? 2 + sidx : L - (sidx - (L - 1));
        const int dir = sidx < L - 1 ? 0 : 1;
        return fa.wprep + (size_t)(2 * (l - 2) + dir) * (wbytes / 4);
      };
      auto fetch = [&](int64_t g) {
        const int slot = (int)(g & 1);
        tc::mbar_expect_tx(&b_wfull[slot], wbytes);
        tc::bulk_copy_g2s(sm + lay.sW + (size_t)slot * (wbytes / 4), weight_of(g), wbytes, &b_wfull[slot]);
      };
      if (n_phase > 0) fetch(0);
      if (n_phase > 1) fetch(1);
      bool first_dw = true;
      for (int64_t g = 0; g < n_phase; ++g) {
        const int slot = (int)(g & 1);
        const int sidx = (int)(g % n_seq);
        const bool bwd = sidx >= L - 1;
        const int l = bwd ? L - (sidx - (L - 1)) : 2 + sidx;
        const uint32_t sWa = tc::smem_u32(sm + lay.sW + (size_t)slot * (wbytes / 4));
        if (!tc::mbar_wait(&b_wfull[slot], ph_wfull[slot] & 1)) __trap();
        ++ph_wfull[slot];
        for (int h = 0; h < 2; ++h) {
          if (!tc::mbar_wait(&b_ready[h], ph_ready[h] & 1)) __trap();
          ++ph_ready[h];
          tc::tc_fence_after();
          const uint32_t sXa = tc::smem_u32(sm + lay.sX + (size_t)h * Kq * (N + 1) * 4);
          for (int ks = 0; ks < H8 / 8; ++ks) {
            const uint64_t da = tc::smem_desc(sWa + ks * 2 * (128 * 16), 128 * 16, 128);
            const uint64_t db = tc::smem_desc(sXa + ks * 2 * ((N + 1) * 16), (N + 1) * 16, 128);
            tc::mma_tf32(tmem + colD + h * N, da, db, idesc_x, ks > 0 ? 1u : 0u);
          }
          if (bwd) {  // dW_l += zbar_l^T h_{l-1}: A = zbar_l from TMEM (lane = unit, column = (channel, point))
            const uint32_t sHa = tc::smem_u32(sm + lay.sHt + (size_t)h * (N / 4) * (N2 + 1) * 4);
            const uint32_t d2 = tmem + colW + (uint32_t)(l - 2) * N2;
            for (int ks = 0; ks < N / 8; ++ks) {
              const uint64_t db = tc::smem_desc(sHa + ks * 2 * ((N2 + 1) * 16), (N2 + 1) * 16, 128);
              tc::mma_tf32_ts(d2, tmem + colA + h * N + ks * 8, db, idesc_w, (first_dw && h == 0 && ks == 0) ? 0u : 1u);
            }
          }
          tc::mma_commit(&b_done[h]);
        }
        if (bwd && l == 2) first_dw = false;  // every dW accumulator has been initialised by the first tile
        // the slot is free once both halves' GEMMs have completed: refill it with the tile two phases ahead
        tc::mma_commit(&b_wfree[slot]);
        if (g + 2 < n_phase) {
          if (!tc::mbar_wait(&b_wfree[slot], ph_wfree[slot] & 1)) __trap();
          ++ph_wfree[slot];
          fetch(g + 2);
        }
      }
    }
    __syncwarp();
  } else {
    // ============================================================== compute warps ====================
    const int h = warp >> 3;                    // half
    const int tl = tid - h * kHalf;             // thread within the half
    const int j = (warp & 3) * 32 + lane;       // hidden unit owned by this thread (TMEM lane)
    const int grp = (warp >> 2) & 1;            // which 4 of the half-tile's 8 points
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const bool unit_ok = j < H;
    const int bar_id = 1 + h;
    float* sX = sm + lay.sX + (size_t)h * Kq * (N + 1) * 4;
    float* sHt = sm + lay.sHt + (size_t)h * (N / 4) * (N2 + 1) * 4;
    float* sY = sm + lay.sY + (size_t)h * 2 * N * 4;
    float* sYb = sY + N * 4;
    float* sXc = sm + lay.sXc + h * kP * 4;
    float* sG = sm + lay.sG + h * PINN_MAX_EQ * kP;
    const float* sB = sm + lay.sB;
    const float* sW1 = sm + lay.sW1;
    const float* sWo = sm + lay.sWo;
    const uint32_t tD = tmem + lane_base + colD + h * N;
    const uint32_t tA = tmem + lane_base + colA + h * N;
    float* const sx_row = sX + ((j >> 2) * (N + 1)) * 4 + (j & 3);  // + q * 4: element (row q, K = unit j)
    float* const sht_row = sHt + j * 4;                              // + (q >> 2) * (N2 + 1) * 4 + (q & 3)
    float* const stash = fa.stash + ((size_t)blockIdx.x * 2 + h) * (L - 1) * C * kP * 128 + j;
    uint32_t ph_done = 0;

    auto layer1 = [&](int p, float (&z)[C]) {
      float v = sB[j];
#pragma unroll
      for (int d = 0; d < 4; ++d) v += sW1[d * 128 + j] * sXc[p * 4 + d];
      z[0] = v;
#pragma unroll
      for (int i = 0; i < NF; ++i) z[1 + i] = sW1[net.first_in[i] * 128 + j];
#pragma unroll
      for (int s = 0; s < NS; ++s) z[1 + NF + s] = 0.f;
    };
    auto wait_done = [&]() {
      if (!tc::mbar_wait(&b_done[h], ph_done & 1)) __trap();
      ++ph_done;
      tc::tc_fence_after();
    };
    auto signal_ready = [&]() {  // this thread's operand writes (smem: generic -> async proxy; TMEM: st complete) are visible
      tc::fence_async_smem();
      tc::tc_fence_before();
      mbar_arrive(&b_ready[h]);
    };

    for (int64_t it = 0; it < n_iter; ++it) {
      const int64_t tile = (it * gridDim.x + blockIdx.x) * 2 + h;
      const int64_t n0 = tile * kP;
      auto sym_at = [&](int id, int p) -> float {
        if (id < PINN_SYM_COORD0) return sY[((id >> 2) * kP + p) * 4 + (id & 3)];
        if (id < PINN_SYM_AUX0) return sXc[p * 4 + (id - PINN_SYM_COORD0)];
        return (n0 + p < fa.n_points) ? fa.dom.aux[id - PINN_SYM_AUX0][n0 + p] : 0.f;
      };
      named_bar(bar_id, kHalf);  // previous tile of this half fully consumed
      if (tl < kP * 4) {
        const int p = tl >> 2, d = tl & 3;
        sXc[tl] = (d < net.n_in && n0 + p < fa.n_points) ? fa.xs[d][n0 + p] : 0.f;
      }
      named_bar(bar_id, kHalf);

      // ---- hidden layer 1 (SIMT): h_1 -> sX ----------------------------------------------------------
      if (j < H8) {
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const int p = grp * 4 + pp;
          float z[C], hv[C];
          layer1(p, z);
          jet_fwd_t<ACT, NF, NS, true>(z, hv);
#pragma unroll
          for (int c = 0; c < C; ++c) sx_row[(c * kP + p) * 4] = tc::to_tf32(hv[c]);
        }
      }
      signal_ready();
      // ---- hidden layers 2..L: tensor-core GEMM (issuer warp) + activation epilogue ----------------------
      float zc[C][4];  // z_L of this thread's four points stays in registers for the reverse sweep
      for (int l = 2; l <= L; ++l) {
        wait_done();
        const float bias = sB[(l - 1) * 128 + j];
#pragma unroll
        for (int c = 0; c < C; ++c) tc::tmem_ld4(tD + c * kP + grp * 4, zc[c]);
        tc::tmem_ld_wait();
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const int p = grp * 4 + pp;
          float z[C], hv[C];
#pragma unroll
          for (int c = 0; c < C; ++c) z[c] = zc[c][pp];
          z[0] += bias;
          zc[0][pp] = z[0];
          jet_fwd_t<ACT, NF, NS, true>(z, hv);
          if (j < H8) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
              if (!fa.loss_only && l < L) stash[(((l - 2) * C + c) * kP + p) * 128] = z[c];
              sx_row[(c * kP + p) * 4] = (l == L) ? hv[c] : tc::to_tf32(hv[c]);
            }
          }
        }
        if (l < L) signal_ready();
      }
      tc::tc_fence_before();
      named_bar(bar_id, kHalf);
      // ---- output layer (FP32 SIMT, h_L kept in FP32): partial sums over 6 slices of the units ------------
      {
        float* part = sm + lay.sP + (size_t)h * kSl * N * 4;
        const int per = (H + kSl - 1) / kSl;
        for (int t = tl; t < kSl * N; t += kHalf) {
          const int q = t % N, sl = t / N;
          const int k0 = sl * per, k1 = min(H, k0 + per);
          float y[4] = {0.f, 0.f, 0.f, 0.f};
          for (int k = k0; k < k1; ++k) {
            const float hv = sX[((k >> 2) * (N + 1) + q) * 4 + (k & 3)];
#pragma unroll
            for (int o = 0; o < 4; ++o) y[o] += sWo[o * 128 + k] * hv;
          }
          *reinterpret_cast<float4*>(part + (sl * N + q) * 4) = make_float4(y[0], y[1], y[2], y[3]);
        }
        named_bar(bar_id, kHalf);
        if (tl < N * 4) {
          const int q = tl >> 2, o = tl & 3;
          float y = 0.f;
#pragma unroll
          for (int sl = 0; sl < kSl; ++sl) y += part[(sl * N + q) * 4 + o];
          if (q < kP && o < O) y += net.b[L][o];
          sY[q * 4 + o] = y;
        }
        named_bar(bar_id, kHalf);
      }
      // ---- residual, loss, seeds ---------------------------------------------------------------------------
      if (tl < kP * PINN_MAX_EQ) {
        const int p = tl & (kP - 1), e = tl / kP;
        const int64_t ng = n0 + p;
        float g = 0.f;
        if (e < fa.dom.n_eq && ng < fa.n_points) {
          const pinn_equation& q = fa.dom.eq[e];
          float r = 0.f;
          for (int t = q.term_begin; t < q.term_end; ++t) {
            const pinn_term& tm = fa.dom.terms[t];
            float prod = tm.coef;
            for (int f = 0; f < tm.n_fac; ++f) prod *= sym_at(tm.fac[f], p);
            r += prod;
          }
          const float area = fa.dom.area ? fa.dom.area[ng] : fa.dom.area_const;
          const float tgt = q.target ? q.target[ng] : q.target_const;
          const float lam = q.lambda ? q.lambda[ng] : q.lambda_const;
          const float diff = r - tgt, w = lam * area;
          float f0, f1;
          if (fa.dom.loss_kind == PINN_LOSS_SQUARE) { f0 = diff * diff; f1 = 2.0f * diff; }
          else if (fa.dom.loss_kind == PINN_LOSS_L1) { f0 = fabsf(diff); f1 = (diff > 0.f) ? 1.0f : ((diff < 0.f) ? -1.0f : 0.f); }
          else { f0 = diff; f1 = 1.0f; }
          loss += f0 * w;
          g = f1 * w * fa.dom.sigma;
        }
        sG[tl] = g;  // [e][p]
      }
      named_bar(bar_id, kHalf);
      if (fa.loss_only) continue;  // uniform over the CTA
      for (int item = tl; item < C * 4 * kP; item += kHalf) {
        const int p = item & (kP - 1), id = item / kP;
        float acc = 0.f;
        for (int e = 0; e < fa.dom.n_eq; ++e) {
          const float g = sG[e * kP + p];
          const pinn_equation& q = fa.dom.eq[e];
          for (int t = q.term_begin; t < q.term_end; ++t) {
            const pinn_term& tm = fa.dom.terms[t];
            for (int f = 0; f < tm.n_fac; ++f) {
              if (tm.fac[f] != id) continue;
              float prod = tm.coef * g;
              for (int f2 = 0; f2 < tm.n_fac; ++f2)
                if (f2 != f) prod *= sym_at(tm.fac[f2], p);
              acc += prod;
            }
          }
        }
        sYb[((id >> 2) * kP + p) * 4 + (id & 3)] = acc;
        if (item < 4 * kP) g_bo_r += acc;  // value channel of output o = id: dL/d bout[o]  (item == tl here)
      }
      named_bar(bar_id, kHalf);

      // ---- reverse sweep ------------------------------------------------------------------------------------
      // step l: zbar_l -> sX (B of hbar_{l-1} = W_l^T zbar_l) and TMEM (A of dW_l); h_{l-1} -> sHt (B of dW_l)
      for (int l = L; l >= 1; --l) {
        float zn[C][4];  // z_{l-1}: h_{l-1} now, the adjoint's z in the next step
        if (l >= 3) {
#pragma unroll
          for (int pp = 0; pp < 4; ++pp)
#pragma unroll
            for (int c = 0; c < C; ++c)
              zn[c][pp] = (j < H8) ? stash[(((l - 3) * C + c) * kP + grp * 4 + pp) * 128] : 0.f;
        } else if (l == 2) {
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
            float z[C];
            layer1(grp * 4 + pp, z);
#pragma unroll
            for (int c = 0; c < C; ++c) zn[c][pp] = z[c];
          }
        }
        if (l < L) wait_done();
        float wo[4];
        if (l == L) {
#pragma unroll
          for (int o = 0; o < 4; ++o) wo[o] = sWo[o * 128 + j];
        }
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          float hbv[C][2];
          if (l < L) {
#pragma unroll
            for (int c = 0; c < C; ++c) tmem_ld2(tD + c * kP + grp * 4 + pr * 2, hbv[c]);
            tc::tmem_ld_wait();
          }
          float zbp[C][2], hpp[C][2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int pp = pr * 2 + e, p = grp * 4 + pp;
            float z[C], hb[C], zb[C];
#pragma unroll
            for (int c = 0; c < C; ++c) z[c] = zc[c][pp];
            if (l == L) {
              float hv[C];
              jet_fwd_t<ACT, NF, NS, true>(z, hv);  // h_L for dWout
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const float4 t = *reinterpret_cast<const float4*>(sYb + (c * kP + p) * 4);
                hb[c] = wo[0] * t.x + wo[1] * t.y + wo[2] * t.z + wo[3] * t.w;
                g_wo[0] += t.x * hv[c]; g_wo[1] += t.y * hv[c]; g_wo[2] += t.z * hv[c]; g_wo[3] += t.w * hv[c];
              }
            } else {
#pragma unroll
              for (int c = 0; c < C; ++c) hb[c] = hbv[c][e];
            }
#pragma unroll
            for (int c = 0; c < C; ++c) zb[c] = 0.f;
            jet_bwd_t<ACT, NF, NS, true>(z, hb, zb);
            if (l == 1) g_b[0] += zb[0];
            else if (l == 2) g_b[1] += zb[0];
            else if (l == 3) g_b[2] += zb[0];
            else g_b[3] += zb[0];
            if (l == 1) {
#pragma unroll
              for (int d = 0; d < 4; ++d) g_w1[d] += zb[0] * sXc[p * 4 + d];
#pragma unroll
              for (int i = 0; i < NF; ++i) {
                const int d = net.first_in[i];
                if (d == 0) g_w1[0] += zb[1 + i];
                else if (d == 1) g_w1[1] += zb[1 + i];
                else if (d == 2) g_w1[2] += zb[1 + i];
                else g_w1[3] += zb[1 + i];
              }
            } else {
              float zq[C], hp[C];
#pragma unroll
              for (int c = 0; c < C; ++c) zq[c] = zn[c][pp];
              jet_fwd_t<ACT, NF, NS, true>(zq, hp);
#pragma unroll
              for (int c = 0; c < C; ++c) {
                zbp[c][e] = tc::to_tf32(zb[c]);
                hpp[c][e] = tc::to_tf32(hp[c]);
              }
            }
          }
          if (l >= 2) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const int q = c * kP + grp * 4 + pr * 2;
              if (j < H8) {
                sx_row[q * 4] = zbp[c][0];
                sx_row[(q + 1) * 4] = zbp[c][1];
              }
              tmem_st2(tA + q, zbp[c][0], zbp[c][1]);
              if (j < N2) *reinterpret_cast<float2*>(sht_row + (q >> 2) * (N2 + 1) * 4 + (q & 3)) = make_float2(hpp[c][0], hpp[c][1]);
            }
          }
        }
        if (l >= 2) {
          tc::tmem_st_wait();
          signal_ready();
#pragma unroll
          for (int c = 0; c < C; ++c)
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) zc[c][pp] = zn[c][pp];
        }
      }
    }

    tc::tc_fence_before();
  }
  // ---- flush: TMEM dW accumulators and per-unit register accumulators -> partial slots ---------------------
  // (each half has waited for the commit of its own last GEMMs, dW included, before it got here)
  __syncthreads();
  tc::tc_fence_after();
  const bool any_tile = (int64_t)blockIdx.x * 2 < n_tiles;
  const int j = (warp & 3) * 32 + lane;
  const bool unit_ok = j < H && !issuer;
  const int grp4 = warp >> 2;  // 0..3: (half, point group); 4: issuer warp
  if (!issuer && any_tile && !fa.loss_only) {
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int l = 2; l <= L; ++l) {
      float* out = fa.dw_part + ((size_t)(l - 2) * fa.n_splits + blockIdx.x) * H * H;
      for (int g = grp4; g < N2 / 16; g += 4) {
        float v[16];
        tc::tmem_ld16(tmem + colW + (uint32_t)(l - 2) * N2 + lane_base + g * 16, v);
        tc::tmem_ld_wait();
        if (unit_ok) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = g * 16 + i;
            if (k < H) out[(size_t)j * H + k] += v[i];
          }
        }
      }
    }
  }
  // the 4 thread groups of a unit: add through shared memory (the weight slots are free now), group 0 writes
  float* red = sm + lay.sW;
  float acc[12] = {g_w1[0], g_w1[1], g_w1[2], g_w1[3], g_b[0], g_b[1], g_b[2], g_b[3], g_wo[0], g_wo[1], g_wo[2], g_wo[3]};
  if (!issuer && grp4 > 0) {
#pragma unroll
    for (int i = 0; i < 12; ++i) red[((grp4 - 1) * 12 + i) * 128 + j] = acc[i];
  }
  __syncthreads();
  if (grp4 == 0 && unit_ok && any_tile) {
#pragma unroll
    for (int g = 1; g < 4; ++g)
#pragma unroll
      for (int i = 0; i < 12; ++i) acc[i] += red[((g - 1) * 12 + i) * 128 + j];
    float* fp = fa.first_part + (size_t)blockIdx.x * 8 * H;
#pragma unroll
    for (int d = 0; d < 4; ++d) fp[d * H + j] += acc[d];
    fp[PINN_MAX_IN * H + j] += acc[4];
#pragma unroll
    for (int l = 2; l <= kMaxHid; ++l)
      if (l <= L) fa.db_part[((size_t)(l - 2) * kBlocks + blockIdx.x) * 8 * H + j] += acc[4 + l - 1];
    float* hp = fa.head_part + (size_t)blockIdx.x * (O * H + O + 1);
    for (int o = 0; o < O; ++o) hp[o * H + j] += acc[8 + o];
  }
  __syncthreads();
  if (!issuer) {
    const int hh = warp >> 3, tl = tid - hh * kHalf;
    if (tl < 4 * kP) red[hh * 128 + tl] = g_bo_r;                        // [half][o][p]
    if (tl < kP * PINN_MAX_EQ) red[256 + hh * 128 + tl] = loss;          // [half][e][p]
  }
  __syncthreads();
  if (tid <= O && any_tile) {
    float sacc = 0.f;
    for (int hh = 0; hh < 2; ++hh) {
      if (tid < O) {
        for (int p = 0; p < kP; ++p) sacc += red[hh * 128 + tid * kP + p];
      } else {
        for (int i = 0; i < kP * PINN_MAX_EQ; ++i) sacc += red[256 + hh * 128 + i];
      }
    }
    fa.head_part[(size_t)blockIdx.x * (O * H + O + 1) + O * H + tid] += sacc;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

// W (row-major [H][H]) -> K-major canonical tile [Kq][128][4] (no padding row: the tiles arrive by bulk copy),
// TF32-rounded, zero padded; transpose != 0 stores W^T (rows = input units of the layer, K = its output units)
__global__ void prep_weights2_kernel(const float* __restrict__ W, int H, int Kq, float* __restrict__ out, int transpose) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Kq * 128 * 4) return;
  const int e = i & 3, row = (i >> 2) & 127, kq = i >> 9;
  const int k = kq * 4 + e;
  float v = 0.f;
  if (row < H && k < H) v = tc::to_tf32(transpose ? W[(size_t)k * H + row] : W[(size_t)row * H + k]);
  out[i] = v;
}

int fused2_wide_prep(const WNet& n, float* wprep, cudaStream_t st) {
  const int Kq = ((n.H + 7) & ~7) / 4, tile = Kq * 128 * 4;
  for (int l = 2; l <= n.L; ++l)
    for (int dir = 0; dir < 2; ++dir) {
      prep_weights2_kernel<<<(tile + 255) / 256, 256, 0, st>>>(n.W[l - 1], n.H, Kq, wprep + (size_t)(2 * (l - 2) + dir) * tile, dir);
      count_launch(1);
    }
  return (int)cudaGetLastError();
}

bool fused2_wide_supported(const WNet& n) {
  if (n.precision != 1 || n.H > 128 || n.H < 8 || n.L < 2 || n.L > kMaxHid || n.n_in > 4) return false;
  const int N = n.C * kP, N2 = (n.H + 15) & ~15;
  if (4 * N + (n.L - 1) * N2 > 512) return false;  // TMEM columns
  return (size_t)smem2_layout(n.H, n.C).total * 4 <= 226 * 1024;
}

// the stash of the first-generation kernel ([cta][(L-1)][C][16][128]) has exactly the size of two half-tiles
template <int ACT>
int fused2_wide_launch(const WNet& n, const pinn_domain* dom, float* stash, const float* wt_base, float* dw_part, float* db_part,
                       float* first_part, float* head_part, int n_splits, bool loss_only, cudaStream_t st) {
  Fused2Args fa;
  memset(&fa, 0, sizeof(fa));
  fa.loss_only = loss_only ? 1 : 0;
  fa.dom = *dom;
  for (int d = 0; d < PINN_MAX_IN; ++d) fa.xs[d] = d < n.n_in ? dom->coords[d] : nullptr;
  fa.n_points = dom->n_points;
  fa.stash = stash;
  fa.wprep = wt_base;
  fa.dw_part = dw_part;
  fa.db_part = db_part;
  fa.first_part = first_part;
  fa.head_part = head_part;
  fa.n_splits = n_splits;
  if (n_splits < kBlocks) return PINN_E_UNSUPPORTED;
  const size_t smem = (size_t)smem2_layout(n.H, n.C).total * sizeof(float);
  const int64_t tiles = (dom->n_points + kP - 1) / kP, pairs = (tiles + 1) / 2;
  int grid = pairs < kBlocks ? (int)pairs : kBlocks;
  if (grid < 1) return 0;
#define CASE(NF_, NS_)                                                                                                 \
  if (n.NF == NF_ && n.NS == NS_) {                                                                                   \
    cudaError_t e = cudaFuncSetAttribute((const void*)wide_fused2_kernel<NF_, NS_, ACT>,                               \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                      \
    if (e != cudaSuccess) return (int)e;                                                                               \
    wide_fused2_kernel<NF_, NS_, ACT><<<grid, kThreads, smem, st>>>(n, fa);                                            \
    launched = true;                                                                                                   \
  }
  bool launched = false;
  CASE(0, 0) CASE(1, 0) CASE(1, 1) CASE(2, 0) CASE(2, 1) CASE(2, 2) CASE(3, 0) CASE(3, 1) CASE(3, 2) CASE(3, 3)
  if (!launched) return PINN_E_UNSUPPORTED;
#undef CASE
  count_launch(1);
  return (int)cudaGetLastError();
}

template int fused2_wide_launch<PINN_ACT_TANH>(const WNet&, const pinn_domain*, float*, const float*, float*, float*, float*, float*, int, bool, cudaStream_t);
template int fused2_wide_launch<PINN_ACT_SILU>(const WNet&, const pinn_domain*, float*, const float*, float*, float*, float*, float*, int, bool, cudaStream_t);
template int fused2_wide_launch<PINN_ACT_SIN>(const WNet&, const pinn_domain*, float*, const float*, float*, float*, float*, float*, int, bool, cudaStream_t);

}  // namespace pinn
