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

// Two instances per jet plan: KP = 16 points per tile with single-pass TF32 multiplicands (PINN_PREC_TF32), and KP = 8 with
// 3xTF32 operand splitting (PINN_PREC_FP32X3, FP32-grade: every operand buffer holds a hi block followed by a lo block,
// hi = tf32(x), lo = x - hi, and a K step issues lo*hi + hi*lo + hi*hi).  The doubled buffers are why the FP32-grade
// instance works on half tiles: 220 KB of shared memory at H = 100, C = 5; it does not fit H = 128.
constexpr int kPMax = 16;    // points per tile of the TF32 instance (the stash is sized for it)
constexpr int kMaxHid = 4;   // hidden layers supported (TMEM: 128 + 3 * 128 columns)
constexpr int kSl = 6;       // unit slices of the SIMT output layer

struct FusedArgs {
  pinn_domain dom;
  const float* xs[PINN_MAX_IN];
  int64_t n_points;
  float* stash;          // [cta][(L-1)][C][KP][128]
  const float* wprep;    // [2*(l-2) + dir][Kq*128*4 (x2: hi, lo)]: W_l (dir 0) / W_l^T (dir 1), TF32-rounded, K-major canonical tiles
  float* dw_part;        // [(l-2)][n_splits][H*H]
  float* db_part;        // [(l-2)][kBlocks][8][H]
  float* first_part;     // [kBlocks][8][H]
  float* head_part;      // [kBlocks][O*H + O + 1]
  int n_splits;
  int loss_only;         // forward + residual + loss only: no stash, no reverse sweep, no gradient partials
  int dbl;               // two sHt buffers: h_{l-2} is prepared while the GEMMs of reverse step l run
};

struct FusedSmem {  // offsets in floats; with x3 every operand buffer is [hi block][lo block], lo at + its *_lo stride
  int sW, sX, sZt, sHt, sHt2, sY, sP, sXc, sWo, sG, total;
  int w_lo, x_lo, zt_lo, ht_lo;  // size of one block of each operand buffer (= offset of its lo block)
};
__host__ __device__ inline FusedSmem fused_smem(int H, int C, int kp, bool x3, bool dbl) {
  const int H8 = (H + 7) & ~7, Kq = H8 / 4, N = C * kp, N2 = (H + 15) & ~15, m = x3 ? 2 : 1;
  FusedSmem s;
  s.w_lo = Kq * 128 * 4;
  s.x_lo = Kq * (N + 1) * 4;
  s.zt_lo = (N / 4) * 129 * 4;
  s.ht_lo = (N / 4) * (N2 + 1) * 4;
  int off = 0;
  s.sW = off; off += m * s.w_lo;                 // [Kq][128][4]      A operand: W_l or W_l^T (arrives by bulk copy)
  s.sX = off; off += m * s.x_lo;                 // [Kq][N + 1][4]    B operand: h_{l-1} / zbar_l, K = unit
  s.sZt = off; off += m * s.zt_lo;               // [N/4][129][4]     A operand of dW: zbar_l, K = (channel, point)
  s.sHt = off; off += m * s.ht_lo;               // [N/4][N2 + 1][4]  B operand of dW: h_{l-1}, K = (channel, point)
  s.sHt2 = dbl ? off : s.sHt; off += dbl ? m * s.ht_lo : 0;
  s.sY = off; off += 2 * N * 4;                  // y[N][4], then ybar[N][4]
  s.sP = off; off += kSl * N * 4;                // partial sums of the output layer
  s.sXc = off; off += kp * 4;                    // coordinates of the tile
  s.sWo = off; off += 4 * 128;                   // Wout as [k][4 outputs]: one 128-bit load per unit
  s.sG = off; off += PINN_MAX_EQ * kp;           // loss seeds of the tile
  s.total = off;
  return s;
}

// round-to-nearest (ties away) to TF32: cvt.rna.tf32.f32 expands on sm_100a to {inf/nan test, add 0x1000 to the bit
// pattern, mask the 13 low mantissa bits}; the operands here are finite and kind::tf32 reads only the upper 19 bits of a
// 32-bit operand, so the add alone gives the tensor core the same multiplicand: 1 instruction instead of 3 (the parity
// tests of tests/test_gpu_wide_tc.py bound the result either way)
__device__ __forceinline__ float tf32r(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
// 3xTF32 operand split: hi = x rounded to TF32 (masked, so that lo = x - hi is exact), lo as it is (the tensor core reads
// its upper 19 bits): dropped terms <= 2^-21 |x| per operand and lo*lo <= 2^-22 |ab|
__device__ __forceinline__ void split3(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}

// KP * 32 threads: thread (unit j = TMEM lane, point group g) handles the four points 4g .. 4g+3 of the tile.
template <int NF, int NS, int ACT, int KP, bool X3>
__global__ void __launch_bounds__(KP * 32, 1) wide_fused_kernel(WNet net, const __grid_constant__ FusedArgs fa) {
  constexpr int C = 1 + NF + NS;
  constexpr int kP = KP;
  constexpr int N = C * kP;
  // MMA N of the layer GEMMs: M = 128 wants a multiple of 16.  For N = 40 (8-point tiles, C = 5) the MMA reads eight rows
  // past each K-quad slab of sX (the next slab's first rows / the following buffer): finite garbage into accumulator
  // columns N .. NB-1 that nobody reads (shared memory is cleared once so that it IS finite).
  constexpr int NB = (N + 15) & ~15;
  constexpr int kT = KP * 32, G = KP / 4;
  const int H = net.H, L = net.L, O = net.n_out;
  const int H8 = (H + 7) & ~7, Kq = H8 / 4, N2 = (H + 15) & ~15;
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bar[2];  // [0]: MMA completion, [1]: weight tile landed
  __shared__ uint32_t tmem_slot;
  const FusedSmem lay = fused_smem(H, C, kP, X3, fa.dbl != 0);
  float* const sW = sm + lay.sW;
  float* const sX = sm + lay.sX;
  float* const sZt = sm + lay.sZt;
  float* const sY = sm + lay.sY;
  float* const sYb = sY + N * 4;
  float* const sP = sm + lay.sP;
  float* const sXc = sm + lay.sXc;
  float* const sWo = sm + lay.sWo;
  float* const sG = sm + lay.sG;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = (warp & 3) * 32 + lane;     // hidden unit owned by this thread (TMEM lane)
  const int grp = warp >> 2;                // which four of the tile's 16 points this thread handles
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const bool unit_ok = j < H;               // only the final flush needs it: padded units carry exact zeros throughout
  const bool k_ok = j < H8, n2_ok = j < N2; // rows of the K = unit / K = (channel, point) operands this thread writes

  // residual evaluation tables (built once): entry i of the (term, factor) list, and its count
  __shared__ uint8_t tf_t[128], tf_f[128];
  __shared__ int n_tf_s;
  if (tid == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    tc::fence_mbar_init();
    int n = 0;
    for (int t = 0; t < fa.dom.n_terms; ++t)
      for (int f = 0; f < fa.dom.terms[t].n_fac; ++f) {
        if (n < 128) { tf_t[n] = (uint8_t)t; tf_f[n] = (uint8_t)f; }
        ++n;
      }
    n_tf_s = n;
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  if (NB > N || X3)  // operand rows the epilogues never write (pad rows, the lo block of h_L, what the N..NB-1 columns read)
    for (int i = tid; i < lay.sY; i += kT) sm[i] = 0.f;  // the operand buffers only (sWo is being filled by other threads)
  for (int i = tid; i < 4 * 128; i += kT) {
    const int o = i & 3, k = i >> 2;
    sWo[i] = (o < O && k < H) ? net.W[L][o * H + k] : 0.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int n_tf = n_tf_s;
  // wide residual phases (one thread per (point, term) and per (point, term, factor)) when their scratch fits into sP
  const bool wide_res = n_tf <= 128 && (fa.dom.n_terms + n_tf) * kP <= kSl * N * 4;
  uint32_t phase = 0, wphase = 0;
  // weight tiles are consumed in a fixed cyclic order per point tile: W_2 .. W_L, W_L^T .. W_2^T
  const uint32_t wbytes = (uint32_t)((X3 ? 2 : 1) * Kq * 128 * 4 * sizeof(float));
  const int n_seq = fa.loss_only ? (L - 1) : 2 * (L - 1);
  const int64_t n_tiles = (fa.n_points + kP - 1) / kP;
  const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  int64_t wleft = my_tiles * n_seq;  // weight tiles still to fetch (thread 0 only)
  int wseq = 0;
  auto fetch_weights = [&]() {  // thread 0: asynchronous bulk copy of the next weight tile into sW
    if (wleft <= 0) return;
    --wleft;
    const int sidx = wseq;
    wseq = (wseq + 1 == n_seq) ? 0 : wseq + 1;
    const int l = sidx < L - 1 ? 2 + sidx : L - (sidx - (L - 1));
    const int dir = sidx < L - 1 ? 0 : 1;
    tc::mbar_expect_tx(&bar[1], wbytes);
    tc::bulk_copy_g2s(sW, fa.wprep + (size_t)(2 * (l - 2) + dir) * (wbytes / 4), wbytes, &bar[1]);
  };
  if (tid == 0) fetch_weights();

  // per-unit constants in registers (zero for padded units: their pre-activations, activations and adjoints are then
  // exactly zero everywhere -- tanh(0) = sin(0) = silu(0) = 0 and every derivative channel is a multiple of a zero)
  float w1[4] = {0.f, 0.f, 0.f, 0.f}, b1 = 0.f, bl[kMaxHid] = {0.f, 0.f, 0.f, 0.f};
  if (unit_ok) {
#pragma unroll
    for (int d = 0; d < 4; ++d)  // static indices keep w1 in registers
      if (d < net.n_in) w1[d] = net.W[0][j * net.n_in + d];
    b1 = net.b[0][j];
#pragma unroll
    for (int l = 2; l <= kMaxHid; ++l)
      if (l <= L) bl[l - 1] = net.b[l - 1][j];
  }
  float w1f[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    const int d = net.first_in[i];
    w1f[i] = d == 0 ? w1[0] : d == 1 ? w1[1] : d == 2 ? w1[2] : w1[3];
  }

  // per-unit gradient accumulators (all four point groups accumulate; summed at the end)
  float g_w1[4] = {0.f, 0.f, 0.f, 0.f}, g_b[kMaxHid + 1] = {0.f, 0.f, 0.f, 0.f, 0.f}, g_wo[4] = {0.f, 0.f, 0.f, 0.f};
  float g_bo_r = 0.f, loss = 0.f;  // threads (o, p) = tid < 64 / threads (e, p) = tid < 128, summed at the end

  // per-thread bases: every element this thread touches is base + a compile-time offset (c, pp)
  const int p0 = grp * 4;
  // stash [(l-2)][c][point group][unit j][4 points]: the four points of a thread are one 128-bit access per channel
  float* const stash_t = fa.stash + (size_t)blockIdx.x * (L - 1) * C * kP * 128 + (grp * 128 + j) * 4;  // + ((l-2)*C + c)*kP*128
  float* const sx_t = sX + ((j >> 2) * (N + 1) + p0) * 4 + (j & 3);                              // + (c*kP + pp)*4
  float* const szt_t = sZt + (grp * 129 + j) * 4;                                                // + c*G*129*4   (float4 = 4 points)
  const int sht_off = (grp * (N2 + 1) + j) * 4;                                                  // + c*G*(N2+1)*4 (float4 = 4 points)
  const uint32_t t_acc = tmem + lane_base + p0;                                                  // + c*kP
  const uint32_t idesc_x = tc::idesc_tf32(NB, 0, 0), idesc_w = tc::idesc_tf32(N2, 0, 0);
  bool first_tile = true;

  auto layer1 = [&](int p, float (&z)[C]) {
    const float4 xv = *reinterpret_cast<const float4*>(sXc + p * 4);
    z[0] = b1 + w1[0] * xv.x + w1[1] * xv.y + w1[2] * xv.z + w1[3] * xv.w;
#pragma unroll
    for (int i = 0; i < NF; ++i) z[1 + i] = w1f[i];
#pragma unroll
    for (int s = 0; s < NS; ++s) z[1 + NF + s] = 0.f;
  };
  // z_m of this thread's four points: from the stash (m >= 2) or recomputed (m == 1)
  auto load_z4 = [&](int m, float (&z4)[C][4]) {
    if (m >= 2) {
      const float* sp = stash_t + (size_t)(m - 2) * C * kP * 128;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(sp + c * kP * 128);
        z4[c][0] = v.x; z4[c][1] = v.y; z4[c][2] = v.z; z4[c][3] = v.w;
      }
    } else {
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        float z[C];
        layer1(p0 + pp, z);
#pragma unroll
        for (int c = 0; c < C; ++c) z4[c][pp] = z[c];
      }
    }
  };
  // h_m (TF32) of this thread's four points -> the K = (channel, point) operand buffer `dst` (one 128-bit store per channel)
  // (X3: h arrives unrounded and is split here into the hi and lo blocks)
  auto store_ht = [&](float* dst, const float (&hq)[C][4]) {
    if (n2_ok) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        if (X3) {
          float hi[4], lo[4];
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) split3(hq[c][pp], hi[pp], lo[pp]);
          *reinterpret_cast<float4*>(dst + sht_off + c * G * (N2 + 1) * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(dst + lay.ht_lo + sht_off + c * G * (N2 + 1) * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        } else {
          *reinterpret_cast<float4*>(dst + sht_off + c * G * (N2 + 1) * 4) = make_float4(hq[c][0], hq[c][1], hq[c][2], hq[c][3]);
        }
      }
    }
  };
  // one multiplicand of the layer GEMMs -> this thread's slot of sX (hi block, and lo block with X3)
  auto store_x = [&](int c, int pp, float v) {
    if (X3) {
      float hi, lo;
      split3(v, hi, lo);
      sx_t[(c * kP + pp) * 4] = hi;
      sx_t[lay.x_lo + (c * kP + pp) * 4] = lo;
    } else {
      sx_t[(c * kP + pp) * 4] = tf32r(v);
    }
  };

  int64_t n0 = 0;
  // residual symbol of point p of the current tile: output jets (sY), coordinates (sXc), aux columns (global)
  auto sym_at = [&](int id, int p) -> float {
    if (id < PINN_SYM_COORD0) return sY[((id >> 2) * kP + p) * 4 + (id & 3)];
    if (id < PINN_SYM_AUX0) return sXc[p * 4 + (id - PINN_SYM_COORD0)];
    return (n0 + p < fa.n_points) ? fa.dom.aux[id - PINN_SYM_AUX0][n0 + p] : 0.f;
  };

  // D = sW * sX^T  (+ optionally D2_l += sZt * sHt^T): barrier, then ONE thread issues; nobody waits here
  auto issue_mma = [&](bool with_dw, int l, const float* sHt_cur) {
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      if (!tc::mbar_wait(&bar[1], wphase & 1)) __trap();  // weight tile landed
      ++wphase;
      tc::tc_fence_after();
      for (int ks = 0; ks < H8 / 8; ++ks) {
        const uint32_t oa = ks * 2 * (128 * 16), ob = ks * 2 * ((N + 1) * 16);
        const uint64_t da = tc::smem_desc(tc::smem_u32(sW) + oa, 128 * 16, 128);
        const uint64_t db = tc::smem_desc(tc::smem_u32(sX) + ob, (N + 1) * 16, 128);
        if (X3) {
          const uint64_t dal = tc::smem_desc(tc::smem_u32(sW + lay.w_lo) + oa, 128 * 16, 128);
          const uint64_t dbl = tc::smem_desc(tc::smem_u32(sX + lay.x_lo) + ob, (N + 1) * 16, 128);
          tc::mma_tf32(tmem, dal, db, idesc_x, ks > 0 ? 1u : 0u);
          tc::mma_tf32(tmem, da, dbl, idesc_x, 1u);
          tc::mma_tf32(tmem, da, db, idesc_x, 1u);
        } else {
          tc::mma_tf32(tmem, da, db, idesc_x, ks > 0 ? 1u : 0u);
        }
      }
      if (with_dw) {
        const uint32_t d2 = tmem + 128 + (uint32_t)(l - 2) * 128;
        for (int ks = 0; ks < N / 8; ++ks) {
          const uint32_t oa = ks * 2 * (129 * 16), ob = ks * 2 * ((N2 + 1) * 16);
          const uint64_t da = tc::smem_desc(tc::smem_u32(sZt) + oa, 129 * 16, 128);
          const uint64_t db = tc::smem_desc(tc::smem_u32(sHt_cur) + ob, (N2 + 1) * 16, 128);
          const uint32_t acc0 = (first_tile && ks == 0) ? 0u : 1u;
          if (X3) {
            const uint64_t dal = tc::smem_desc(tc::smem_u32(sZt + lay.zt_lo) + oa, 129 * 16, 128);
            const uint64_t dbl = tc::smem_desc(tc::smem_u32(sHt_cur + lay.ht_lo) + ob, (N2 + 1) * 16, 128);
            tc::mma_tf32(d2, dal, db, idesc_w, acc0);
            tc::mma_tf32(d2, da, dbl, idesc_w, 1u);
            tc::mma_tf32(d2, da, db, idesc_w, 1u);
          } else {
            tc::mma_tf32(d2, da, db, idesc_w, acc0);
          }
        }
      }
      tc::mma_commit(&bar[0]);
    }
  };
  auto wait_mma = [&]() {
    if (!tc::mbar_wait(&bar[0], phase & 1)) __trap();
    ++phase;
    tc::tc_fence_after();
    if (tid == 0) fetch_weights();  // sW is free again: the next tile's copy overlaps the epilogue
  };

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    n0 = tile * kP;
    __syncthreads();  // previous tile fully consumed
    if (tid < kP * 4) {
      const int p = tid >> 2, d = tid & 3;
      sXc[tid] = (d < net.n_in && n0 + p < fa.n_points) ? fa.xs[d][n0 + p] : 0.f;
    }
    __syncthreads();

    // ---- hidden layer 1 (SIMT): h_1 -> sX (and -> sHt when it is the dW operand of the first reverse step) ------
    {
      float hq[C][4];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        float z[C], h[C];
        layer1(p0 + pp, z);
        jet_fwd_t<ACT, NF, NS, !X3>(z, h);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          hq[c][pp] = X3 ? h[c] : tf32r(h[c]);
          if (k_ok) store_x(c, pp, h[c]);
        }
      }
      if (L == 2 && !fa.loss_only) store_ht(sm + lay.sHt, hq);
    }
    // ---- hidden layers 2..L: tensor-core GEMM + activation epilogue; the stash stores of layer l ride in the
    //      shadow of the GEMM of layer l + 1 ----------------------------------------------------------------------
    float zk[C][4];  // pre-activations of the layer just finished (kept for the deferred stash stores)
    for (int l = 2; l <= L; ++l) {
      issue_mma(false, 0, nullptr);
      if (l > 2 && !fa.loss_only && k_ok) {  // z_{l-1} -> stash while the GEMM of layer l runs
        float* sp = stash_t + (size_t)(l - 3) * C * kP * 128;
#pragma unroll
        for (int c = 0; c < C; ++c)
          *reinterpret_cast<float4*>(sp + c * kP * 128) = make_float4(zk[c][0], zk[c][1], zk[c][2], zk[c][3]);
      }
      wait_mma();
#pragma unroll
      for (int c = 0; c < C; ++c) tc::tmem_ld4(t_acc + c * kP, zk[c]);
      tc::tmem_ld_wait();
      const float bias = l == 2 ? bl[1] : l == 3 ? bl[2] : bl[3];
      float hq[C][4];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        float z[C], h[C];
        zk[0][pp] += bias;
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = zk[c][pp];
        jet_fwd_t<ACT, NF, NS, !X3>(z, h);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          hq[c][pp] = X3 ? h[c] : tf32r(h[c]);
          if (k_ok) {
            if (l == L) sx_t[(c * kP + pp) * 4] = h[c];  // h_L stays FP32 for the SIMT output layer
            else store_x(c, pp, h[c]);
          }
        }
      }
      if (l == L - 1 && !fa.loss_only) store_ht(sm + lay.sHt, hq);  // h_{L-1}: B operand of dW_L, first reverse step
      tc::tc_fence_before();
    }
    if (!fa.loss_only && k_ok) {  // z_L -> stash
      float* sp = stash_t + (size_t)(L - 2) * C * kP * 128;
#pragma unroll
      for (int c = 0; c < C; ++c)
        *reinterpret_cast<float4*>(sp + c * kP * 128) = make_float4(zk[c][0], zk[c][1], zk[c][2], zk[c][3]);
    }
    __syncthreads();
    // ---- output layer (FP32 SIMT, h_L in FP32): partial sums over kSl slices of the units, then a short reduction ----
    {
      const int per = (H + kSl - 1) / kSl;
      for (int t = tid; t < kSl * N; t += kT) {
        const int q = t % N, sl = t / N;
        const int k0 = sl * per, k1 = min(H, k0 + per);
        float y[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = k0; k < k1; ++k) {
          const float hv = sX[((k >> 2) * (N + 1) + q) * 4 + (k & 3)];
          const float4 w = *reinterpret_cast<const float4*>(sWo + k * 4);  // warp-uniform: one broadcast load
          y[0] += w.x * hv; y[1] += w.y * hv; y[2] += w.z * hv; y[3] += w.w * hv;
        }
        *reinterpret_cast<float4*>(sP + (sl * N + q) * 4) = make_float4(y[0], y[1], y[2], y[3]);
      }
      __syncthreads();
      if (tid < N * 4) {
        const int q = tid >> 2, o = tid & 3;
        float y = 0.f;
#pragma unroll
        for (int sl = 0; sl < kSl; ++sl) y += sP[(sl * N + q) * 4 + o];
        if (q < kP && o < O) y += net.b[L][o];
        sY[q * 4 + o] = y;
      }
      __syncthreads();
    }
    // ---- residual, loss, ybar ------------------------------------------------------------------------------
    // Four short, wide phases (the serial per-equation / per-symbol loops were 10 % of the kernel's stall samples, spent
    // by 464 threads waiting at the barriers for 48): R1a thread (point, term) -> the monomial; R1b thread (point,
    // equation) -> r_e, loss term, seed g_e = f'(r - target) * w * sigma; R2a thread (point, term, factor) -> g_e * d term /
    // d factor; R2b thread (point, output symbol) -> ybar.  Summation orders (terms within an equation; equations, terms
    // and factors within a symbol) are those of the serial evaluation of the layer-streamed path: bitwise the same sums.
    float* const sT = sP;                                // [n_terms][kP]
    float* const sPt = sP + fa.dom.n_terms * kP;         // [n_tf][kP]
    if (wide_res) {
      for (int item = tid; item < fa.dom.n_terms * kP; item += kT) {
        const int p = item % kP, t = item / kP;
        const pinn_term& tm = fa.dom.terms[t];
        float prod = tm.coef;
        for (int f = 0; f < tm.n_fac; ++f) prod *= sym_at(tm.fac[f], p);
        sT[item] = prod;
      }
      __syncthreads();
    }
    if (tid < kP * PINN_MAX_EQ) {
      const int p = tid % kP, e = tid / kP;
      const int64_t ng = n0 + p;
      float g = 0.f;
      if (e < fa.dom.n_eq && ng < fa.n_points) {
        const pinn_equation& q = fa.dom.eq[e];
        float r = 0.f;
        if (wide_res) {
          for (int t = q.term_begin; t < q.term_end; ++t) r += sT[t * kP + p];
        } else {
          for (int t = q.term_begin; t < q.term_end; ++t) {
            const pinn_term& tm = fa.dom.terms[t];
            float prod = tm.coef;
            for (int f = 0; f < tm.n_fac; ++f) prod *= sym_at(tm.fac[f], p);
            r += prod;
          }
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
    if (wide_res) {
      for (int item = tid; item < n_tf * kP; item += kT) {
        const int p = item % kP, i = item / kP;
        const int t = tf_t[i], f = tf_f[i];
        const pinn_term& tm = fa.dom.terms[t];
        int e = 0;
        while (e + 1 < fa.dom.n_eq && t >= fa.dom.eq[e].term_end) ++e;
        float prod = tm.coef * sG[e * kP + p];
        for (int f2 = 0; f2 < tm.n_fac; ++f2)
          if (f2 != f) prod *= sym_at(tm.fac[f2], p);
        sPt[item] = prod;
      }
      __syncthreads();
      for (int item = tid; item < C * 4 * kP; item += kT) {
        const int p = item % kP, id = item / kP;
        float acc = 0.f;
        for (int i = 0; i < n_tf; ++i)
          if (fa.dom.terms[tf_t[i]].fac[tf_f[i]] == id) acc += sPt[i * kP + p];
        sYb[((id >> 2) * kP + p) * 4 + (id & 3)] = acc;
        if (item < 4 * kP) g_bo_r += acc;  // value channel of output o = id: dL/d bout[o]  (item == tid here)
      }
    } else {
      for (int item = tid; item < C * 4 * kP; item += kT) {
        const int p = item % kP, id = item / kP;
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
        if (item < 4 * kP) g_bo_r += acc;  // value channel of output o = id: dL/d bout[o]  (item == tid here)
      }
    }
    __syncthreads();

    // ---- reverse sweep -------------------------------------------------------------------
    // step l: zbar_l -> sX and sZt; the tensor cores then compute hbar_{l-1} = W_l^T zbar_l and dW_l += zbar_l^T h_{l-1}
    // with h_{l-1} taken from an sHt buffer that was filled EARLIER: by the forward epilogue (l = L), or -- with two
    // buffers -- while the GEMMs of step l + 1 were running.
    int buf = 0;
    for (int l = L; l >= 1; --l) {
      float zl[C][4];  // z_l of the four points (stash reads do not depend on the tensor cores: issued before the wait)
      load_z4(l, zl);
      float wo[4];
      if (l == L) {
        const float4 w = *reinterpret_cast<const float4*>(sWo + j * 4);
        wo[0] = w.x; wo[1] = w.y; wo[2] = w.z; wo[3] = w.w;
      } else {
        wait_mma();
      }
      float zq[C][4];  // zbar_l (TF32) of the four points
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        const int p = p0 + pp;
        float z[C], hb[C], zb[C];
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = zl[c][pp];
        if (l == L) {
          float h[C];
          jet_fwd_t<ACT, NF, NS, !X3>(z, h);  // h_L for dWout
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float4 t = *reinterpret_cast<const float4*>(sYb + (c * kP + p) * 4);
            hb[c] = wo[0] * t.x + wo[1] * t.y + wo[2] * t.z + wo[3] * t.w;
            g_wo[0] += t.x * h[c]; g_wo[1] += t.y * h[c]; g_wo[2] += t.z * h[c]; g_wo[3] += t.w * h[c];
          }
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c)  // one TMEM column per load keeps the live set small; the C loads pipeline
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(hb[c]) : "r"(t_acc + c * kP + pp) : "memory");
          tc::tmem_ld_wait();
        }
#pragma unroll
        for (int c = 0; c < C; ++c) zb[c] = 0.f;
        jet_bwd_t<ACT, NF, NS, !X3>(z, hb, zb);
        if (l == 1) g_b[0] += zb[0];
        else if (l == 2) g_b[1] += zb[0];
        else if (l == 3) g_b[2] += zb[0];
        else g_b[3] += zb[0];
        if (l == 1) {
          const float4 xv = *reinterpret_cast<const float4*>(sXc + p * 4);
          g_w1[0] += zb[0] * xv.x; g_w1[1] += zb[0] * xv.y; g_w1[2] += zb[0] * xv.z; g_w1[3] += zb[0] * xv.w;
#pragma unroll
          for (int i = 0; i < NF; ++i) {
            const int d = net.first_in[i];
            if (d == 0) g_w1[0] += zb[1 + i];
            else if (d == 1) g_w1[1] += zb[1 + i];
            else if (d == 2) g_w1[2] += zb[1 + i];
            else g_w1[3] += zb[1 + i];
          }
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            zq[c][pp] = X3 ? zb[c] : tf32r(zb[c]);
            if (k_ok) store_x(c, pp, zb[c]);
          }
        }
      }
      if (l >= 2) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          if (X3) {
            float hi[4], lo[4];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) split3(zq[c][pp], hi[pp], lo[pp]);
            *reinterpret_cast<float4*>(szt_t + c * G * 129 * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(szt_t + lay.zt_lo + c * G * 129 * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          } else {
            *reinterpret_cast<float4*>(szt_t + c * G * 129 * 4) = make_float4(zq[c][0], zq[c][1], zq[c][2], zq[c][3]);
          }
        }
        if (!fa.dbl && l < L) {  // single sHt buffer: h_{l-1} now (the GEMMs of step l + 1 have completed)
          float hq[C][4], z4[C][4];
          load_z4(l - 1, z4);
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
            float z[C], h[C];
#pragma unroll
            for (int c = 0; c < C; ++c) z[c] = z4[c][pp];
            jet_fwd_t<ACT, NF, NS, !X3>(z, h);
#pragma unroll
            for (int c = 0; c < C; ++c) hq[c][pp] = X3 ? h[c] : tf32r(h[c]);
          }
          store_ht(sm + lay.sHt, hq);
        }
        tc::tc_fence_before();
        issue_mma(true, l, sm + (buf ? lay.sHt2 : lay.sHt));
        if (fa.dbl && l >= 3) {  // h_{l-2}, the dW operand of the NEXT step, in the shadow of this step's GEMMs
          float hq[C][4], z4[C][4];
          load_z4(l - 2, z4);
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
            float z[C], h[C];
#pragma unroll
            for (int c = 0; c < C; ++c) z[c] = z4[c][pp];
            jet_fwd_t<ACT, NF, NS, !X3>(z, h);
#pragma unroll
            for (int c = 0; c < C; ++c) hq[c][pp] = X3 ? h[c] : tf32r(h[c]);
          }
          buf ^= 1;
          store_ht(sm + (buf ? lay.sHt2 : lay.sHt), hq);
        }
      }
    }
    first_tile = false;
  }

  // ---- flush: TMEM dW accumulators and per-unit register accumulators -> partial slots -----------
  __syncthreads();
  const bool any_tile = blockIdx.x < n_tiles;
  if (any_tile && !fa.loss_only) {
    tc::tc_fence_after();
    // TMEM row = unit j (one per lane), 16 columns k per load; a 32 x 17 shared-memory tile per warp turns the
    // read-modify-write of dw_part[j][k] into row segments of 16 consecutive floats (two rows per warp access)
    // instead of 32 scattered words (the flush was 7 % of the kernel's stall samples, half of a boundary launch)
    float* tile = sW + warp * (32 * 17);  // sW and sX (adjacent) are free now; tiny nets that cannot hold a tile per warp write directly
    const bool use_tile = (kT / 32) * 32 * 17 <= lay.sZt - lay.sW;
    for (int l = 2; l <= L; ++l) {
      float* out = fa.dw_part + ((size_t)(l - 2) * fa.n_splits + blockIdx.x) * H * H;
      for (int g = grp; g < N2 / 16; g += G) {
        float v[16];
        tc::tmem_ld16(tmem + 128 + (uint32_t)(l - 2) * 128 + lane_base + g * 16, v);
        tc::tmem_ld_wait();
        if (use_tile) {
#pragma unroll
          for (int i = 0; i < 16; ++i) tile[lane * 17 + i] = v[i];
          __syncwarp();
          const int k = g * 16 + (lane & 15), jb = (warp & 3) * 32;
#pragma unroll
          for (int r = 0; r < 32; r += 2) {
            const int rr = r + (lane >> 4), jj = jb + rr;
            if (jj < H && k < H) out[(size_t)jj * H + k] += tile[rr * 17 + (lane & 15)];
          }
          __syncwarp();
        } else if (unit_ok) {
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
  float* red = sW;  // free now (no copy in flight): (G - 1) * 12 * 128 floats of scratch for the final reductions
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

// W (row-major [H][H]) -> K-major canonical tile [Kq][128][4], TF32-rounded, zero padded;
// transpose != 0 stores W^T (rows = input units of the layer, K = its output units); x3: the hi tile is followed by the
// residual tile lo = W - hi
__global__ void prep_weights_kernel(const float* __restrict__ W, int H, int Kq, float* __restrict__ out, int transpose, int x3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Kq * 128 * 4) return;
  const int e = i & 3, row = (i >> 2) & 127, kq = i >> 9;
  const int k = kq * 4 + e;
  float v = 0.f;
  if (row < H && k < H) v = transpose ? W[(size_t)k * H + row] : W[(size_t)row * H + k];
  const float hi = tc::to_tf32(v);
  out[i] = hi;
  if (x3) out[Kq * 128 * 4 + i] = v - hi;
}

int fused_wide_prep(const WNet& n, float* wprep, cudaStream_t st) {
  const int x3 = n.precision == 2 ? 1 : 0;
  const int Kq = ((n.H + 7) & ~7) / 4, tile = Kq * 128 * 4;
  for (int l = 2; l <= n.L; ++l)
    for (int dir = 0; dir < 2; ++dir) {
      prep_weights_kernel<<<(tile + 255) / 256, 256, 0, st>>>(n.W[l - 1], n.H, Kq, wprep + (size_t)(2 * (l - 2) + dir) * tile * (1 + x3), dir, x3);
      count_launch(1);
    }
  return (int)cudaGetLastError();
}

// (sized for the 3xTF32 instance: hi + lo tiles)
int64_t fused_wide_wprep_floats(const WNet& n) { return (int64_t)2 * 2 * (n.L - 1) * (((n.H + 7) & ~7) / 4) * 128 * 4; }

static bool fused_fits_kp(const WNet& n, int kp, bool dbl) {
  const FusedSmem lay = fused_smem(n.H, n.C, kp, n.precision == 2, dbl);
  // (the reductions of the flush reuse sW: (G - 1) * 12 * 128 floats; the dW GEMM contracts over N = C * kp in steps of 8)
  return (size_t)lay.total * 4 <= 226 * 1024 && (kp / 4 - 1) * 12 * 128 <= lay.sZt - lay.sW && (n.C * kp) % 8 == 0;
}

// tile size of the instance that serves this net: TF32 -> 16-point tiles; FP32-grade (3xTF32, doubled operand buffers) ->
// 8-point tiles.  (A 4-point instance fits H = 128 -- 186 KB -- but measured SLOWER than the layer-streamed 3xTF32 kernels
// on allen_cahn, 2.28e7 vs 2.60e7 points/s, so such nets stay on the streamed path.)
static int fused_kp(const WNet& n) { return n.precision == 2 ? 8 : kPMax; }

static bool fused_fits(const WNet& n, bool dbl) { return fused_fits_kp(n, fused_kp(n), dbl); }

bool fused_wide_supported(const WNet& n) {
  if ((n.precision != 1 && n.precision != 2) || n.H > 128 || n.L < 2 || n.L > kMaxHid || n.n_in > 4) return false;
  return fused_fits(n, false);
}

int64_t fused_wide_stash_floats(const WNet& n) { return (int64_t)kBlocks * (n.L - 1) * n.C * kPMax * 128; }

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
  fa.dbl = fused_fits(n, true) ? 1 : 0;
  const int kp = fused_kp(n);
  const bool x3 = n.precision == 2;
  const size_t smem = (size_t)fused_smem(n.H, n.C, kp, x3, fa.dbl != 0).total * sizeof(float);
  const int64_t tiles = (dom->n_points + kp - 1) / kp;
  int grid = tiles < kBlocks ? (int)tiles : kBlocks;
  if (grid < 1) return 0;
#define LAUNCH(NF_, NS_, KP_, X3_)                                                                                    \
  {                                                                                                                   \
    cudaError_t e = cudaFuncSetAttribute((const void*)wide_fused_kernel<NF_, NS_, ACT, KP_, X3_>,                      \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                      \
    if (e != cudaSuccess) return (int)e;                                                                               \
    wide_fused_kernel<NF_, NS_, ACT, KP_, X3_><<<grid, KP_ * 32, smem, st>>>(n, fa);                                   \
    launched = true;                                                                                                   \
  }
#define CASE(NF_, NS_)                                                                                                 \
  if (n.NF == NF_ && n.NS == NS_) {                                                                                   \
    if (x3) LAUNCH(NF_, NS_, 8, true)                                                                                 \
    else LAUNCH(NF_, NS_, kPMax, false)                                                                               \
  }
  bool launched = false;
  CASE(0, 0) CASE(1, 0) CASE(1, 1) CASE(2, 0) CASE(2, 1) CASE(2, 2) CASE(3, 0) CASE(3, 1) CASE(3, 2) CASE(3, 3)
  if (!launched) return PINN_E_UNSUPPORTED;
#undef CASE
#undef LAUNCH
  count_launch(1);
  return (int)cudaGetLastError();
}

template int fused_wide_launch<PINN_ACT_TANH>(const WNet&, const pinn_domain*, float*, const float*, float*, float*, float*, float*, int, bool, cudaStream_t);
template int fused_wide_launch<PINN_ACT_SILU>(const WNet&, const pinn_domain*, float*, const float*, float*, float*, float*, float*, int, bool, cudaStream_t);
template int fused_wide_launch<PINN_ACT_SIN>(const WNet&, const pinn_domain*, float*, const float*, float*, float*, float*, float*, int, bool, cudaStream_t);

}  // namespace pinn
