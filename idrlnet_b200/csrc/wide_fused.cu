// Wide nets, H <= 128, 2..4 hidden layers: the whole training step of a tile of 16 points on chip.
//
// One persistent CTA per SM walks over tiles of P = 16 collocation points.  The jets of the tile
// never leave the SM between layers: each layer's GEMM runs on the tensor cores (tcgen05.mma
// kind::tf32, accumulators in TMEM), in the swapped orientation of wide_tc.cu (TMEM lane = hidden
// unit, column = channel * 16 + point), and the epilogue thread that owns a unit writes the next
// layer's operand straight back into shared memory in the K-major canonical layout.  The only
// forward state kept for the reverse pass is the pre-activation jet of layers 2..L, stashed in a
// per-CTA global scratch that every tile overwrites (L2-resident, as in the narrow kernel).  The
// weight gradients dW_l = zbar_l^T h_{l-1} are tensor-core GEMMs contracting over the tile's
// (channel, point) pairs whose accumulators stay in TMEM for the whole kernel (3 x 128 columns)
// and are written once per CTA into the per-split partial slots of the layer-streamed path, so
// wide_finalize and the deterministic reduction are shared.  The output layer, the residual and
// the loss are FP32 SIMT code.  HBM traffic per point: the coordinates (+ targets/lambda columns).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "tc_common.cuh"
#include "wide_internal.cuh"

namespace pinn {
namespace {

constexpr int kP = 16;       // points per tile
constexpr int kMaxHid = 4;   // hidden layers supported (TMEM: 128 + 3 * 128 columns)

struct FusedArgs {
  pinn_domain dom;
  const float* xs[PINN_MAX_IN];
  int64_t n_points;
  float* stash;          // [cta][(L-1)][C][kP][128]
  const float* wprep;    // [2*(l-2) + dir][Kq*129*4]: W_l (dir 0) / W_l^T (dir 1), TF32-rounded, canonical padded tiles
  float* dw_part;        // [(l-2)][n_splits][H*H]
  float* db_part;        // [(l-2)][kBlocks][8][H]
  float* first_part;     // [kBlocks][8][H]
  float* head_part;      // [kBlocks][O*H + O + 1]
  int n_splits;
  int loss_only;         // forward + residual + loss only: no stash, no reverse sweep, no gradient partials
};

// G point groups: 128 * G threads, thread (unit j, group g) handles the kP / G points g * kP / G ... of the tile.
// G = 4 (16 warps, <= 128 registers) hides the latency of the serial layer chain better than G = 2 (8 warps, up to
// 234 registers): +8 % on the 5-channel LDC interior, and the value-only boundary domains are L2-latency bound.
template <int NF, int NS, int ACT, int G>
__global__ void __launch_bounds__(128 * G, 1) wide_fused_kernel(WNet net, const __grid_constant__ FusedArgs fa) {
  constexpr int C = 1 + NF + NS;
  constexpr int N = C * kP;
  constexpr int kT = 128 * G, kPP = kP / G;
  const int H = net.H, L = net.L, O = net.n_out;
  const int H8 = (H + 7) & ~7, Kq = H8 / 4, N2 = (H + 15) & ~15;
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bar[2];  // [0]: MMA completion, [1]: weight tile landed
  __shared__ uint32_t tmem_slot;
  float* sW = sm;                           // [Kq][129][4]      A operand: W_l or W_l^T
  float* sX = sW + Kq * 129 * 4;            // [Kq][N + 1][4]    B operand: h_{l-1} / zbar_l, K = unit
  float* sZt = sX + Kq * (N + 1) * 4;       // [N/4][129][4]     A operand of dW: zbar_l, K = (channel, point)
  float* sHt = sZt + (N / 4) * 129 * 4;     // [N/4][N2 + 1][4]  B operand of dW: h_{l-1}, K = (channel, point)
  float* sY = sHt + (N / 4) * (N2 + 1) * 4; // [N][4]  y, then ybar
  float* sXc = sY + 2 * N * 4;              // [kP][4] coordinates of the tile
  float* sWo = sXc + kP * 4;                // [4][128] Wout[o][k]
  float* sG = sWo + 4 * 128;                // [PINN_MAX_EQ][kP] loss seeds of the tile

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = (warp & 3) * 32 + lane;     // hidden unit owned by this thread (TMEM lane)
  const int grp = warp >> 2;                // which kPP of the tile's 16 points this thread handles
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const bool unit_ok = j < H;

  if (tid == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  for (int i = tid; i < 4 * 128; i += kT) {
    const int o = i >> 7, k = i & 127;
    sWo[i] = (o < O && k < H) ? net.W[L][o * H + k] : 0.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t phase = 0, wphase = 0;
  // weight tiles are consumed in a fixed cyclic order per point tile: W_2 .. W_L, W_L^T .. W_2^T
  const uint32_t wbytes = (uint32_t)(Kq * 129 * 4 * sizeof(float));
  const int n_seq = fa.loss_only ? (L - 1) : 2 * (L - 1);
  int wseq = 0;  // next tile to fetch (thread 0 only)
  auto fetch_weights = [&]() {  // thread 0: asynchronous bulk copy of the next weight tile into sW
    const int sidx = wseq % n_seq;
    const int l = sidx < L - 1 ? 2 + sidx : L - (sidx - (L - 1));
    const int dir = sidx < L - 1 ? 0 : 1;
    tc::mbar_expect_tx(&bar[1], wbytes);
    tc::bulk_copy_g2s(sW, fa.wprep + (size_t)(2 * (l - 2) + dir) * (wbytes / 4), wbytes, &bar[1]);
    ++wseq;
  };
  if (tid == 0 && blockIdx.x < (fa.n_points + kP - 1) / kP) fetch_weights();

  // per-unit constants in registers
  float w1[4] = {0.f, 0.f, 0.f, 0.f}, b1 = 0.f, bl[kMaxHid] = {0.f, 0.f, 0.f, 0.f}, wo[4];
  if (unit_ok) {
#pragma unroll
    for (int d = 0; d < 4; ++d)  // static indices keep w1 in registers
      if (d < net.n_in) w1[d] = net.W[0][j * net.n_in + d];
    b1 = net.b[0][j];
#pragma unroll
    for (int l = 2; l <= kMaxHid; ++l)
      if (l <= L) bl[l - 1] = net.b[l - 1][j];
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) wo[o] = sWo[o * 128 + j];
  float w1f[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    const int d = net.first_in[i];
    w1f[i] = d == 0 ? w1[0] : d == 1 ? w1[1] : d == 2 ? w1[2] : w1[3];
  }

  // per-unit gradient accumulators (both point halves accumulate; summed at the end)
  float g_w1[4] = {0.f, 0.f, 0.f, 0.f}, g_b[kMaxHid + 1] = {0.f, 0.f, 0.f, 0.f, 0.f}, g_wo[4] = {0.f, 0.f, 0.f, 0.f};
  float g_bo_r = 0.f, loss = 0.f;  // threads (o, p) = tid < 64 / threads (e, p) = tid < 128, summed at the end

  float* stash = fa.stash + (size_t)blockIdx.x * (L - 1) * C * kP * 128;
  const uint32_t idesc_x = tc::idesc_tf32(N, 0, 0), idesc_w = tc::idesc_tf32(N2, 0, 0);
  const int64_t n_tiles = (fa.n_points + kP - 1) / kP;
  bool first_tile = true;

  auto layer1 = [&](const float (&x)[4], float (&z)[C]) {
    float v = b1;
#pragma unroll
    for (int d = 0; d < 4; ++d) v += w1[d] * x[d];
    z[0] = v;
#pragma unroll
    for (int i = 0; i < NF; ++i) z[1 + i] = w1f[i];
#pragma unroll
    for (int s = 0; s < NS; ++s) z[1 + NF + s] = 0.f;
  };
  // element (row q = c*kP + p, K index = unit j) of the K = unit operand
  auto sx_at = [&](int q) -> float& { return sX[((j >> 2) * (N + 1) + q) * 4 + (j & 3)]; };
  // element (row = unit j, K index q) of the K = (channel, point) operands
  auto szt_at = [&](int q) -> float& { return sZt[((q >> 2) * 129 + j) * 4 + (q & 3)]; };
  auto sht_at = [&](int q) -> float& { return sHt[((q >> 2) * (N2 + 1) + j) * 4 + (q & 3)]; };

  int64_t n0 = 0;
  // residual symbol of point p of the current tile: output jets (sY), coordinates (sXc), aux columns (global)
  auto sym_at = [&](int id, int p) -> float {
    if (id < PINN_SYM_COORD0) return sY[((id >> 2) * kP + p) * 4 + (id & 3)];
    if (id < PINN_SYM_AUX0) return sXc[p * 4 + (id - PINN_SYM_COORD0)];
    return (n0 + p < fa.n_points) ? fa.dom.aux[id - PINN_SYM_AUX0][n0 + p] : 0.f;
  };

  auto run_mma = [&](bool with_dw, int l, bool more) {  // D = sW * sX^T  (+ optionally D2_l += sZt * sHt^T)
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      if (!tc::mbar_wait(&bar[1], wphase & 1)) __trap();  // weight tile landed
      ++wphase;
      tc::tc_fence_after();
      for (int ks = 0; ks < H8 / 8; ++ks) {
        const uint64_t da = tc::smem_desc(tc::smem_u32(sW) + ks * 2 * (129 * 16), 129 * 16, 128);
        const uint64_t db = tc::smem_desc(tc::smem_u32(sX) + ks * 2 * ((N + 1) * 16), (N + 1) * 16, 128);
        tc::mma_tf32(tmem, da, db, idesc_x, ks > 0 ? 1u : 0u);
      }
      if (with_dw) {
        const uint32_t d2 = tmem + 128 + (uint32_t)(l - 2) * 128;
        for (int ks = 0; ks < N / 8; ++ks) {
          const uint64_t da = tc::smem_desc(tc::smem_u32(sZt) + ks * 2 * (129 * 16), 129 * 16, 128);
          const uint64_t db = tc::smem_desc(tc::smem_u32(sHt) + ks * 2 * ((N2 + 1) * 16), (N2 + 1) * 16, 128);
          tc::mma_tf32(d2, da, db, idesc_w, (first_tile && ks == 0) ? 0u : 1u);
        }
      }
      tc::mma_commit(&bar[0]);
    }
    if (!tc::mbar_wait(&bar[0], phase & 1)) __trap();
    ++phase;
    tc::tc_fence_after();
    if (tid == 0 && more) fetch_weights();  // sW is free again: overlap the next tile's copy with the epilogue
  };

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    n0 = tile * kP;
    __syncthreads();  // previous tile fully consumed
    if (tid < kP * 4) {
      const int p = tid >> 2, d = tid & 3;
      sXc[tid] = (d < net.n_in && n0 + p < fa.n_points) ? fa.xs[d][n0 + p] : 0.f;
    }
    __syncthreads();

    // ---- hidden layer 1 (SIMT): h_1 -> sX -------------------------------------------------
#pragma unroll 2
    for (int pp = 0; pp < kPP; ++pp) {
      const int p = grp * kPP + pp;
      float x[4] = {sXc[p * 4], sXc[p * 4 + 1], sXc[p * 4 + 2], sXc[p * 4 + 3]};
      float z[C], h[C];
      layer1(x, z);
      jet_fwd_t<ACT, NF, NS, true>(z, h);
      if (j < H8) {
#pragma unroll
        for (int c = 0; c < C; ++c) sx_at(c * kP + p) = unit_ok ? tc::to_tf32(h[c]) : 0.f;
      }
    }
    // ---- hidden layers 2..L: tensor-core GEMM + activation epilogue -------------------------
    for (int l = 2; l <= L; ++l) {
      run_mma(false, 0, !(fa.loss_only && l == L && tile + gridDim.x >= n_tiles));
      float v[C][kPP];
#pragma unroll
      for (int c = 0; c < C; ++c) tc::tmem_ldn(tmem + lane_base + c * kP + grp * kPP, v[c]);
      tc::tmem_ld_wait();
      const float bias = l == 2 ? bl[1] : l == 3 ? bl[2] : bl[3];
#pragma unroll
      for (int pp = 0; pp < kPP; ++pp) {
        const int p = grp * kPP + pp;
        float z[C], h[C];
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = v[c][pp];
        z[0] += bias;
        jet_fwd_t<ACT, NF, NS, true>(z, h);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          if (!fa.loss_only) stash[(((size_t)(l - 2) * C + c) * kP + p) * 128 + j] = z[c];
          if (j < H8) sx_at(c * kP + p) = unit_ok ? (l == L ? h[c] : tc::to_tf32(h[c])) : 0.f;
        }
      }
      tc::tc_fence_before();
    }
    __syncthreads();
    // ---- output layer (FP32 SIMT): y[q][o] = sum_k Wout[o][k] h_L[q][k] (h_L kept in FP32) ------
    if (tid < N) {
      const int q = tid;
      float y[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < H; ++k) {
        const float hv = sX[((k >> 2) * (N + 1) + q) * 4 + (k & 3)];
#pragma unroll
        for (int o = 0; o < 4; ++o) y[o] += sWo[o * 128 + k] * hv;
      }
      if (q < kP) {
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (o < O) y[o] += net.b[L][o];
      }
      *reinterpret_cast<float4*>(sY + q * 4) = make_float4(y[0], y[1], y[2], y[3]);
    }
    __syncthreads();
    // ---- residual, loss, ybar: two short parallel phases instead of one thread per point ---------
    // R1: thread (point p, equation e) evaluates r_e, the loss term and the seed g = f'(r - target) * w * sigma
    if (tid < kP * PINN_MAX_EQ) {
      const int p = tid & (kP - 1), e = tid >> 4;
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
      sG[tid] = g;  // [e][p]
    }
    __syncthreads();
    if (fa.loss_only) continue;  // (uniform: every thread leaves together)
    // R2: thread (point p, output symbol id) gathers ybar[id] = sum_e g_e * d r_e / d sym[id] (same e, term, factor order
    // as the serial evaluation of the layer-streamed path)
    for (int item = tid; item < C * 4 * kP; item += kT) {
      const int p = item & (kP - 1), id = item >> 4;
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
      sY[N * 4 + ((id >> 2) * kP + p) * 4 + (id & 3)] = acc;
      if (item < 4 * kP) g_bo_r += acc;  // value channel of output o = id: dL/d bout[o]  (item == tid here)
    }
    __syncthreads();
    const float* sYb = sY + N * 4;

    // ---- reverse sweep -------------------------------------------------------------------
    // step "l": produce zbar_l (into sX and sZt) and h_{l-1} (into sHt); then the tensor cores
    // compute hbar_{l-1} = W_l^T zbar_l and dW_l += zbar_l^T h_{l-1}.
    bool last_weights = false;
    for (int l = L; l >= 1; --l) {
      // stash reads do not depend on the tensor cores: issue them first
      float zl[C][kPP], zp[C][kPP];
#pragma unroll
      for (int pp = 0; pp < kPP; ++pp) {
        const int p = grp * kPP + pp;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          zl[c][pp] = (l >= 2) ? stash[(((size_t)(l - 2) * C + c) * kP + p) * 128 + j] : 0.f;
          zp[c][pp] = (l >= 3) ? stash[(((size_t)(l - 3) * C + c) * kP + p) * 128 + j] : 0.f;
        }
      }
      float v[C][kPP];
      if (l < L) {
#pragma unroll
        for (int c = 0; c < C; ++c) tc::tmem_ldn(tmem + lane_base + c * kP + grp * kPP, v[c]);
        tc::tmem_ld_wait();
      }
#pragma unroll
      for (int pp = 0; pp < kPP; ++pp) {
        const int p = grp * kPP + pp;
        float x[4] = {sXc[p * 4], sXc[p * 4 + 1], sXc[p * 4 + 2], sXc[p * 4 + 3]};
        float z[C], hb[C], zb[C], h[C];
        if (l >= 2) {
#pragma unroll
          for (int c = 0; c < C; ++c) z[c] = zl[c][pp];
        } else {
          layer1(x, z);
        }
        // hbar_l: from ybar (l == L) or from the GEMM accumulators
        if (l == L) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float4 t = *reinterpret_cast<const float4*>(sYb + (c * kP + p) * 4);
            hb[c] = wo[0] * t.x + wo[1] * t.y + wo[2] * t.z + wo[3] * t.w;
          }
          jet_fwd_t<ACT, NF, NS, true>(z, h);  // h_L for dWout
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float4 t = *reinterpret_cast<const float4*>(sYb + (c * kP + p) * 4);
            g_wo[0] += t.x * h[c]; g_wo[1] += t.y * h[c]; g_wo[2] += t.z * h[c]; g_wo[3] += t.w * h[c];
          }
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) hb[c] = v[c][pp];
        }
#pragma unroll
        for (int c = 0; c < C; ++c) zb[c] = 0.f;
        jet_bwd_t<ACT, NF, NS, true>(z, hb, zb);
        if (!unit_ok) {
#pragma unroll
          for (int c = 0; c < C; ++c) zb[c] = 0.f;
        }
        if (l == 1) g_b[0] += zb[0];
        else if (l == 2) g_b[1] += zb[0];
        else if (l == 3) g_b[2] += zb[0];
        else g_b[3] += zb[0];
        if (l == 1) {
#pragma unroll
          for (int d = 0; d < 4; ++d) g_w1[d] += zb[0] * x[d];
#pragma unroll
          for (int i = 0; i < NF; ++i) {
            const int d = net.first_in[i];
            if (d == 0) g_w1[0] += zb[1 + i];
            else if (d == 1) g_w1[1] += zb[1 + i];
            else if (d == 2) g_w1[2] += zb[1 + i];
            else g_w1[3] += zb[1 + i];
          }
        } else {
          // operands of the next two GEMMs
          float zq[C], hp[C];
          if (l - 1 >= 2) {
#pragma unroll
            for (int c = 0; c < C; ++c) zq[c] = zp[c][pp];
          } else {
            layer1(x, zq);
          }
          jet_fwd_t<ACT, NF, NS, true>(zq, hp);
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int q = c * kP + p;
            const float zt = tc::to_tf32(zb[c]);
            if (j < H8) sx_at(q) = zt;
            szt_at(q) = zt;
            if (j < N2) sht_at(q) = unit_ok ? tc::to_tf32(hp[c]) : 0.f;
          }
        }
      }
      if (l >= 2) {
        tc::tc_fence_before();
        last_weights = (l == 2) && (tile + gridDim.x >= n_tiles);  // nothing left to fetch after this GEMM
        run_mma(true, l, !last_weights);
      }
    }
    first_tile = false;
  }

  // ---- flush: TMEM dW accumulators and per-unit register accumulators -> partial slots -----------
  __syncthreads();
  const bool any_tile = blockIdx.x < n_tiles;
  if (any_tile && !fa.loss_only) {
    tc::tc_fence_after();
    for (int l = 2; l <= L; ++l) {
      float* out = fa.dw_part + ((size_t)(l - 2) * fa.n_splits + blockIdx.x) * H * H;
      for (int g = grp; g < N2 / 16; g += G) {
        float v[16];
        tc::tmem_ld16(tmem + 128 + (uint32_t)(l - 2) * 128 + lane_base + g * 16, v);
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
  // the G point groups of a unit: add through shared memory, group 0 writes
  float* red = sX;  // free now: (G - 1) * 12 * 128 floats of scratch for the final reductions
  float acc[13] = {g_w1[0], g_w1[1], g_w1[2], g_w1[3], g_b[0], g_b[1], g_b[2], g_b[3], g_wo[0], g_wo[1], g_wo[2], g_wo[3], 0.f};
  __syncthreads();
  if (grp > 0) {
#pragma unroll
    for (int i = 0; i < 12; ++i) red[((grp - 1) * 12 + i) * 128 + j] = acc[i];
  }
  __syncthreads();
  if (grp == 0 && unit_ok) {
#pragma unroll
    for (int g = 1; g < G; ++g)
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
  if (tid < 4 * kP) red[tid] = g_bo_r;                   // [o][p]
  if (tid < kP * PINN_MAX_EQ) red[4 * kP + tid] = loss;  // [e][p]
  __syncthreads();
  if (tid <= O) {
    float s = 0.f;
    if (tid < O) {
      for (int p = 0; p < kP; ++p) s += red[tid * kP + p];
    } else {
      for (int i = 0; i < kP * PINN_MAX_EQ; ++i) s += red[4 * kP + i];
    }
    fa.head_part[(size_t)blockIdx.x * (O * H + O + 1) + O * H + tid] += s;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

// W (row-major [H][H]) -> canonical padded K-major tile [Kq][129][4], TF32-rounded, zero padded;
// transpose != 0 stores W^T (rows = input units of the layer, K = its output units)
__global__ void prep_weights_kernel(const float* __restrict__ W, int H, int Kq, float* __restrict__ out, int transpose) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Kq * 129 * 4) return;
  const int e = i & 3, row = (i >> 2) % 129, kq = (i >> 2) / 129;
  const int k = kq * 4 + e;
  float v = 0.f;
  if (row < H && k < H) v = tc::to_tf32(transpose ? W[(size_t)k * H + row] : W[(size_t)row * H + k]);
  out[i] = v;
}

int fused_wide_prep(const WNet& n, float* wprep, cudaStream_t st) {
  const int Kq = ((n.H + 7) & ~7) / 4, tile = Kq * 129 * 4;
  for (int l = 2; l <= n.L; ++l)
    for (int dir = 0; dir < 2; ++dir) {
      prep_weights_kernel<<<(tile + 255) / 256, 256, 0, st>>>(n.W[l - 1], n.H, Kq, wprep + (size_t)(2 * (l - 2) + dir) * tile, dir);
      count_launch(1);
    }
  return (int)cudaGetLastError();
}

int64_t fused_wide_wprep_floats(const WNet& n) { return (int64_t)2 * (n.L - 1) * (((n.H + 7) & ~7) / 4) * 129 * 4; }

bool fused_wide_supported(const WNet& n) {
  if (n.precision != 1 || n.H > 128 || n.L < 2 || n.L > kMaxHid || n.n_in > 4) return false;
  const int H8 = (n.H + 7) & ~7, Kq = H8 / 4, N = n.C * kP, N2 = (n.H + 15) & ~15;
  const size_t floats = (size_t)Kq * 129 * 4 + (size_t)Kq * (N + 1) * 4 + (size_t)(N / 4) * 129 * 4 +
                        (size_t)(N / 4) * (N2 + 1) * 4 + 2 * N * 4 + kP * 4 + 4 * 128 + kP * PINN_MAX_EQ;
  return floats * 4 <= 225 * 1024 && 3 * 12 * 128 <= Kq * (N + 1) * 4;
}

int64_t fused_wide_stash_floats(const WNet& n) { return (int64_t)kBlocks * (n.L - 1) * n.C * kP * 128; }

template <int ACT>
int fused_wide_launch(const WNet& n, const pinn_domain* dom, float* stash, const float* wt_base, float* dw_part, float* db_part,
                      float* first_part, float* head_part, int n_splits, bool loss_only, cudaStream_t st) {
  FusedArgs fa;
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
  const int H8 = (n.H + 7) & ~7, Kq = H8 / 4, N = n.C * kP, N2 = (n.H + 15) & ~15;
  const size_t smem = ((size_t)Kq * 129 * 4 + (size_t)Kq * (N + 1) * 4 + (size_t)(N / 4) * 129 * 4 +
                       (size_t)(N / 4) * (N2 + 1) * 4 + 2 * N * 4 + kP * 4 + 4 * 128 + kP * PINN_MAX_EQ) * sizeof(float);
  const int64_t tiles = (dom->n_points + kP - 1) / kP;
  int grid = tiles < kBlocks ? (int)tiles : kBlocks;
  if (grid < 1) return 0;
#define CASE(NF_, NS_)                                                                                                 \
  if (n.NF == NF_ && n.NS == NS_) {                                                                                   \
    constexpr int G_ = 4;                                                                                              \
    cudaError_t e = cudaFuncSetAttribute((const void*)wide_fused_kernel<NF_, NS_, ACT, G_>,                            \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                      \
    if (e != cudaSuccess) return (int)e;                                                                               \
    wide_fused_kernel<NF_, NS_, ACT, G_><<<grid, 128 * G_, smem, st>>>(n, fa);                                         \
    launched = true;                                                                                                   \
  }
  bool launched = false;
  CASE(0, 0) CASE(1, 0) CASE(1, 1) CASE(2, 0) CASE(2, 1) CASE(2, 2) CASE(3, 0) CASE(3, 1) CASE(3, 2) CASE(3, 3)
  if (!launched) return PINN_E_UNSUPPORTED;
#undef CASE
  count_launch(1);
  return (int)cudaGetLastError();
}

template int fused_wide_launch<PINN_ACT_TANH>(const WNet&, const pinn_domain*, float*, const float*, float*, float*, float*, float*, int, bool, cudaStream_t);
template int fused_wide_launch<PINN_ACT_SILU>(const WNet&, const pinn_domain*, float*, const float*, float*, float*, float*, float*, int, bool, cudaStream_t);
template int fused_wide_launch<PINN_ACT_SIN>(const WNet&, const pinn_domain*, float*, const float*, float*, float*, float*, float*, int, bool, cudaStream_t);

}  // namespace pinn
