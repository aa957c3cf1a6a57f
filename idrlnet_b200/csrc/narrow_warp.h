// Fused PINN step for NARROW nets (hidden width H <= 32): FP32 FMA, one warp
// owns a tile of 32 collocation points from the forward pass through the
// reverse pass; one lane = one point.
//
// Replaces idrlnet/net.py:14-36 + idrlnet/variable.py:161-208 (forward Taylor
// jets instead of nested autograd.grad), idrlnet/node.py:105-109 (residual),
// idrlnet/variable.py:40-80 (weighted loss) and loss.backward()
// (idrlnet/solver.py:338).  Math: SURVEY.md section 8(a) "Jet mathematics".
//
// Data flow per warp tile (all on chip; HBM traffic = the coordinate columns,
// targets/lambda/area in, nothing per point out):
//   * the lane's layer-input jet h[c][k] lives in registers (C*H values);
//   * the (transposed) weights Wt[k][j] are shared by the CTA in smem and read
//     as warp-broadcast float4 -> 4*C FMAs per LDS.128;
//   * pre-activation jets z of layers 2..L are the only forward state kept for
//     the reverse pass (activations are recomputed).  They are stashed in a
//     per-warp scratch region in GLOBAL memory that every tile overwrites, so it
//     stays L2-resident (148 CTAs x 8 warps x 30 KB = 36 MB << 126 MB L2) and
//     costs no shared memory: 8 warps per SM instead of 4;
//   * per warp only two smem row buffers remain: A (zbar_l, or ybar) and
//     B (h_{l-1}), the operands of the warp-local GEMM
//     dW_l = sum over the 32 points and C channels of zbar_l^T h_{l-1}
//     (4x4 register tiles, accumulated in a per-warp smem copy of the
//     gradient -> no atomics, deterministic); row stride chosen so float4
//     accesses are conflict free;
//   * warps never synchronise with each other inside the tile loop.
//
// The code below is written as per-lane "phase" functions separated by
// __syncwarp(); tests/emu runs the same functions lane by lane on the CPU.
// NarrowWarp is the scalar-FMA program (any channel count); NarrowWarpPair (further down) is the
// same program on channel pairs and packed FMAs for even channel counts; NarrowWarpFor picks one.
#pragma once
#include "pinn_common.h"

#ifndef PINN_DW_UNROLL
#define PINN_DW_UNROLL 8  // k-steps of the tensor-path dW loop unrolled together (1 / 2 / 4 / 8 measured: profiles/r02/narrow_dw_variants.txt)
#endif
#ifndef PINN_PAIR_WARPS
#define PINN_PAIR_WARPS 8  // warps per CTA of the packed program (register budget: 65536 / (32 * warps))
#endif

namespace pinn {

enum { MODE_FUSED = 0, MODE_FWD = 1, MODE_BWD = 2, MODE_LOSS = 3 };  // LOSS: forward + residual + loss, no reverse pass

struct NarrowParams {
  // network
  int32_t n_in, n_out, n_hidden, mode;  // mode: MODE_*
  const float* W[PINN_MAX_LAYERS];
  const float* b[PINN_MAX_LAYERS];
  pinn_jets jets;
  // points
  int64_t n_points;
  const float* coords[PINN_MAX_IN];
  // fused mode
  pinn_domain dom;
  // generic modes
  float* jets_out;        // MODE_FWD : [(c*n_out+o)*N + n]
  const float* jets_bar;  // MODE_BWD : same layout
  // results
  float* gpart;  // [max_cta][PP] gradient partials (padded layout, see Layout)
  float* lpart;  // [n_slots][max_cta] loss partials
  float* stash;  // [cta][warp][(L-1)*C*H*32] pre-activation jets of layers 2..L (L2-resident scratch)
  float* gacc;   // [cta][warp][Layout] per-warp gradient accumulators (L2-resident scratch)
  int32_t max_cta, domain_slot;
  int32_t n_warps, n_tiles;
};

// Padded, transposed parameter layout used in smem for the weights, for the
// per-warp gradient accumulators and for the per-CTA gradient partials.
struct Layout {
  int H, L;
  PINN_HD Layout(int h, int l) : H(h), L(l) {}
  PINN_HD int w1() const { return 0; }                    // W1t[d][j]   4*H
  PINN_HD int b1() const { return 4 * H; }                 // b1[j]       H
  PINN_HD int wh(int l) const { return 5 * H + (l - 2) * (H * H + H); }  // Wt_l[k][j], l = 2..L
  PINN_HD int bh(int l) const { return wh(l) + H * H; }
  PINN_HD int wo() const { return 5 * H + (L - 1) * (H * H + H); }      // Woutt[k][o4]  4*H
  PINN_HD int bo() const { return wo() + 4 * H; }          // bout[4]
  PINN_HD int size() const { return bo() + 4; }
};

// X_ selects extra jet channels after the pure second-order ones: 1 = the mixed second derivative
// d2/d(first_in[0]) d(first_in[1]); 2 = the pure third derivative along first_in[0]; 3 = third and fourth (the chain of
// examples/euler_beam/euler_beam.py:47-54).
template <int H_, int NF_, int NS_, int ACT_, int X_ = 0>
struct NarrowWarp {
  static constexpr int H = H_, NF = NF_, NS = NS_, ACT = ACT_, X = X_;
  static constexpr int NM = (X == 1) ? 1 : 0;                 // mixed channels
  static constexpr int NH = (X == 2) ? 1 : (X == 3) ? 2 : 0;  // third / fourth order channels
  static constexpr int CM = 1 + NF + NS;                      // index of the mixed channel
  static constexpr int C3 = 1 + NF + NS + NM;                 // index of the third-order channel (fourth: C3 + 1)
  static constexpr int C = 1 + NF + NS + NM + NH;
  static_assert(X >= 0 && X <= 3, "extra-channel mode");
  static_assert(NM == 0 || NF >= 2, "the mixed channel needs two first-order channels");
  static_assert(NH == 0 || NS >= 1, "third/fourth order along first_in[0] need its second-order channel");
  static constexpr int H4 = H / 4;
  static constexpr int kMaxWarps = 8;  // (10 warps fit in shared memory now, but at <= 200 registers they lose: profiles/r01/narrow_warps_sweep.txt)
  static_assert(H % 4 == 0, "hidden width must be a multiple of 4");
  static_assert(NS <= NF, "second-order channels need their first-order channel");

  // ---- per-warp shared memory carve-up (in floats) -----------------------
  // [A: 32 rows of S][B: 32 rows of S][xs: 32 x 4]; the per-warp gradient accumulators (Layout) live in an
  // L2-resident global scratch so that shared memory holds 10 warps of row buffers
  PINN_HD static int S() { return row_stride(C * H); }
  PINN_HD static int off_a() { return 0; }
  PINN_HD static int off_b() { return 32 * S(); }
  PINN_HD static int off_xs() { return 2 * 32 * S(); }
  PINN_HD static int warp_floats(int) { return off_xs() + 32 * 4; }
  // per-warp global stash (floats): [(l-2)][c][k4][lane][4]
  PINN_HD static int stash_floats(int L) { return (L - 1) * C * H * 32; }
  PINN_HD static int stash_off(int l, int c, int k4, int lane) { return (((l - 2) * C + c) * H4 + k4) * 128 + lane * 4; }

  struct Lane {          // lane-private state crossing phase boundaries
    float loss;          // running un-weighted loss of this lane's points
    float zk[C][H];      // zbar of the layer just differentiated (registers on the GPU)
  };

  // ---- small pieces -------------------------------------------------------
  // sigma .. sigma^(5) at z (the orders above 3 only for the third/fourth-order channels)
  PINN_HD static void act_all(float z, float (&sd)[6]) {
    if (NH > 0) {
      act_derivs6<ACT>(z, sd);
    } else {
      act_derivs<ACT>(z, sd[0], sd[1], sd[2], sd[3]);
      sd[4] = 0.f; sd[5] = 0.f;
    }
  }

  // h = sigma(z) as a jet.  Third / fourth order along one input (Faa di Bruno; z1, z2, z3, z4 the derivatives of z):
  //   h3 = s3 z1^3 + 3 s2 z1 z2 + s1 z3;   h4 = s4 z1^4 + 6 s3 z1^2 z2 + 3 s2 z2^2 + 4 s2 z1 z3 + s1 z4;
  // mixed: h_ab = s2 z_a z_b + s1 z_ab.
  PINN_HD static void unit_fwd(const float (&z)[C], float (&h)[C]) {
    float sd[6];
    act_all(z[0], sd);
    const float s1 = sd[1], s2 = sd[2];
    h[0] = sd[0];
#pragma unroll
    for (int i = 0; i < NF; ++i) h[1 + i] = s1 * z[1 + i];
#pragma unroll
    for (int s = 0; s < NS; ++s) h[1 + NF + s] = s2 * z[1 + s] * z[1 + s] + s1 * z[1 + NF + s];
    if (NM > 0) h[CM] = s2 * z[1] * z[NF >= 2 ? 2 : 1] + s1 * z[CM];
    if (NH > 0) {
      const float z1 = z[1], z2 = z[1 + NF], z3 = z[C3];
      h[C3] = sd[3] * z1 * z1 * z1 + 3.0f * s2 * z1 * z2 + s1 * z3;
      if (NH > 1)
        h[C - 1] = sd[4] * (z1 * z1) * (z1 * z1) + 6.0f * sd[3] * z1 * z1 * z2 + s2 * (3.0f * z2 * z2 + 4.0f * z1 * z3) +
                   s1 * z[C - 1];
    }
  }

  // adjoint of unit_fwd: zb = (dh/dz)^T hb (every sigma^(n) becomes sigma^(n+1) in the value slot); also returns h
  PINN_HD static void unit_bwd(const float (&z)[C], const float (&hb)[C], float (&h)[C], float (&zb)[C]) {
    float sd[6];
    act_all(z[0], sd);
    const float s1 = sd[1], s2 = sd[2], s3 = sd[3];
    h[0] = sd[0];
    float zv = hb[0] * s1;
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      h[1 + i] = s1 * z[1 + i];
      zv += hb[1 + i] * s2 * z[1 + i];
      zb[1 + i] = hb[1 + i] * s1;
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const float zi = z[1 + s], zii = z[1 + NF + s], hbs = hb[1 + NF + s];
      h[1 + NF + s] = s2 * zi * zi + s1 * zii;
      zv += hbs * (s3 * zi * zi + s2 * zii);
      zb[1 + s] += 2.0f * hbs * s2 * zi;
      zb[1 + NF + s] = hbs * s1;
    }
    if (NM > 0) {
      constexpr int B = NF >= 2 ? 2 : 1;
      const float za = z[1], zc = z[B], zm = z[CM], hbm = hb[CM];
      h[CM] = s2 * za * zc + s1 * zm;
      zv += hbm * (s3 * za * zc + s2 * zm);
      zb[1] += hbm * s2 * zc;
      zb[B] += hbm * s2 * za;
      zb[CM] = hbm * s1;
    }
    if (NH > 0) {
      const float z1 = z[1], z2 = z[1 + NF], z3 = z[C3], hb3 = hb[C3], s4 = sd[4];
      h[C3] = s3 * z1 * z1 * z1 + 3.0f * s2 * z1 * z2 + s1 * z3;
      zv += hb3 * (s4 * z1 * z1 * z1 + 3.0f * s3 * z1 * z2 + s2 * z3);
      zb[1] += hb3 * (3.0f * s3 * z1 * z1 + 3.0f * s2 * z2);
      zb[1 + NF] += hb3 * 3.0f * s2 * z1;
      zb[C3] = hb3 * s1;
      if (NH > 1) {
        const float z4 = z[C - 1], hb4 = hb[C - 1], q1 = z1 * z1;
        h[C - 1] = s4 * q1 * q1 + 6.0f * s3 * q1 * z2 + s2 * (3.0f * z2 * z2 + 4.0f * z1 * z3) + s1 * z4;
        zv += hb4 * (sd[5] * q1 * q1 + 6.0f * s4 * q1 * z2 + s3 * (3.0f * z2 * z2 + 4.0f * z1 * z3) + s2 * z4);
        zb[1] += hb4 * (4.0f * s4 * q1 * z1 + 12.0f * s3 * z1 * z2 + 4.0f * s2 * z3);
        zb[1 + NF] += hb4 * (6.0f * s3 * q1 + 6.0f * s2 * z2);
        zb[C3] += hb4 * 4.0f * s2 * z1;
        zb[C - 1] = hb4 * s1;
      }
    }
    zb[0] = zv;
  }

  // pre-activation jet of hidden layer 1 for units k4*4..k4*4+3 (never stashed:
  // z_val = b1 + W1 x, z_i = W1[:, first_in[i]], z_ii = 0)
  PINN_HD static void layer1_tile(const float* sw, const Layout& lay, const NarrowParams& p, const float (&x)[4],
                                  int k4, f4 (&zt)[C]) {
    f4 w[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) w[d] = ld4(sw + lay.w1() + d * H + k4 * 4);
    f4 b = ld4(sw + lay.b1() + k4 * 4);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = f4_at(b, e);
#pragma unroll
      for (int d = 0; d < 4; ++d) v += f4_at(w[d], e) * x[d];
      f4_at(zt[0], e) = v;
    }
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      const int d = p.jets.first_in[i];
      f4 r = w[0];  // select without dynamic register indexing
      if (d == 1) r = w[1];
      if (d == 2) r = w[2];
      if (d == 3) r = w[3];
      zt[1 + i] = r;
    }
#pragma unroll
    for (int c = 1 + NF; c < C; ++c) zt[c] = f4{0.f, 0.f, 0.f, 0.f};  // z_1 is linear in x: every higher derivative is zero
  }

  PINN_HD static void load_row_tile(const float* row, int k4, f4 (&t)[C]) {
#pragma unroll
    for (int c = 0; c < C; ++c) t[c] = ld4(row + c * H + k4 * 4);
  }
  PINN_HD static void store_row_tile(float* row, int k4, const f4 (&t)[C]) {
#pragma unroll
    for (int c = 0; c < C; ++c) st4(row + c * H + k4 * 4, t[c]);
  }

  // hin[c][k] <- activation jet of the z tile
  PINN_HD static void act_tile_fwd(const f4 (&zt)[C], int k4, float (&hin)[C][H]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float z[C], h[C];
#pragma unroll
      for (int c = 0; c < C; ++c) z[c] = (&zt[c].x)[e];
      unit_fwd(z, h);
#pragma unroll
      for (int c = 0; c < C; ++c) hin[c][k4 * 4 + e] = h[c];
    }
  }

  // ---- residual, loss and its seed adjoint (FUSED mode) -------------------
  // y[c][o] are the output jets of this lane's point; returns ybar[c][o] =
  // d(sigma * loss)/dy and adds the point's un-weighted loss to lane.loss.
  PINN_HD static void residual_loss(const NarrowParams& p, int64_t n, bool valid, const float (&x)[4],
                                    const f4 (&y)[C], f4 (&yb)[C], Lane& lane) {
    const pinn_domain& dm = p.dom;
    float sym[PINN_SYM_COUNT];
    float sg[PINN_SYM_COORD0];
#pragma unroll
    for (int i = 0; i < PINN_SYM_COUNT; ++i) sym[i] = 0.f;
#pragma unroll
    for (int i = 0; i < PINN_SYM_COORD0; ++i) sg[i] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      sym[c * 4 + 0] = y[c].x; sym[c * 4 + 1] = y[c].y; sym[c * 4 + 2] = y[c].z; sym[c * 4 + 3] = y[c].w;
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) sym[PINN_SYM_COORD0 + d] = x[d];
    for (int a = 0; a < dm.n_aux; ++a) sym[PINN_SYM_AUX0 + a] = valid ? dm.aux[a][n] : 0.f;
    const float area = valid ? (dm.area ? dm.area[n] : dm.area_const) : 0.f;
    float loss = 0.f;
    for (int e = 0; e < dm.n_eq; ++e) {
      const pinn_equation& q = dm.eq[e];
      float r = 0.f;
      for (int t = q.term_begin; t < q.term_end; ++t) {
        const pinn_term& tm = dm.terms[t];
        float prod = tm.coef;
        for (int f = 0; f < tm.n_fac; ++f) prod *= sym[tm.fac[f]];
        r += prod;
      }
      const float tgt = q.target ? (valid ? q.target[n] : 0.f) : q.target_const;
      const float lam = q.lambda ? (valid ? q.lambda[n] : 0.f) : q.lambda_const;
      const float diff = r - tgt;
      const float w = lam * area;
      float f0, f1;
      if (dm.loss_kind == PINN_LOSS_SQUARE) {
        f0 = diff * diff; f1 = 2.0f * diff;
      } else if (dm.loss_kind == PINN_LOSS_L1) {
        f0 = fabsf(diff); f1 = (diff > 0.f) ? 1.0f : ((diff < 0.f) ? -1.0f : 0.f);
      } else {
        f0 = diff; f1 = 1.0f;
      }
      loss += f0 * w;
      const float g = f1 * w * dm.sigma;
      for (int t = q.term_begin; t < q.term_end; ++t) {
        const pinn_term& tm = dm.terms[t];
        for (int f = 0; f < tm.n_fac; ++f) {
          const int id = tm.fac[f];
          if (id >= PINN_SYM_COORD0) continue;  // only net outputs carry gradient
          float prod = tm.coef * g;
          for (int f2 = 0; f2 < tm.n_fac; ++f2)
            if (f2 != f) prod *= sym[tm.fac[f2]];
          sg[id] += prod;
        }
      }
    }
    lane.loss += loss;
#pragma unroll
    for (int c = 0; c < C; ++c) yb[c] = f4{sg[c * 4 + 0], sg[c * 4 + 1], sg[c * 4 + 2], sg[c * 4 + 3]};
  }

  // ---- phase F: forward, residual/loss; leaves ybar in this lane's A row ----
  // sw: CTA weights, wm: this warp's smem region, gs: this warp's global stash.
  PINN_HD static void phase_forward(const NarrowParams& p, const float* sw, float* wm, float* gs, int lane_id,
                                    int64_t tile, Lane& lane) {
    const int L = p.n_hidden;
    const Layout lay(H, L);
    const int s = S();
    const int64_t n = tile * 32 + lane_id;
    const bool valid = n < p.n_points;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (d < p.n_in && valid) x[d] = p.coords[d][n];
    st4(wm + off_xs() + lane_id * 4, f4{x[0], x[1], x[2], x[3]});
    float* a_row = wm + off_a() + lane_id * s;

    float hin[C][H];
#pragma unroll
    for (int k4 = 0; k4 < H4; ++k4) {  // hidden layer 1
      f4 zt[C];
      layer1_tile(sw, lay, p, x, k4, zt);
      act_tile_fwd(zt, k4, hin);
    }
    for (int l = 2; l <= L; ++l) {  // hidden layers 2..L
      const float* Wt = sw + lay.wh(l);
      const float* bl = sw + lay.bh(l);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
      for (int j4 = 0; j4 < H4; ++j4) {
        f4 acc[C];
        acc[0] = ld4(bl + j4 * 4);
#pragma unroll
        for (int c = 1; c < C; ++c) acc[c] = f4{0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const f4 w = ld4(Wt + k * H + j4 * 4);
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float hv = hin[c][k];
            acc[c].x += w.x * hv; acc[c].y += w.y * hv; acc[c].z += w.z * hv; acc[c].w += w.w * hv;
          }
        }
        if (p.mode == MODE_FUSED || p.mode == MODE_BWD) {  // the stash only feeds the reverse pass
#pragma unroll
          for (int c = 0; c < C; ++c) st4(gs + stash_off(l, c, j4, lane_id), acc[c]);
        }
        // activation of this tile inside the loop: its MUFU/ALU work overlaps the next tile's FMAs
        f4 ht[C];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float z[C], h[C];
#pragma unroll
          for (int c = 0; c < C; ++c) z[c] = (&acc[c].x)[e];
          unit_fwd(z, h);
#pragma unroll
          for (int c = 0; c < C; ++c) (&ht[c].x)[e] = h[c];
        }
        store_row_tile(a_row, j4, ht);  // scratch: hin is still being read by the remaining tiles
      }
#pragma unroll
      for (int k4 = 0; k4 < H4; ++k4) {
        f4 ht[C];
        load_row_tile(a_row, k4, ht);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          hin[c][k4 * 4 + 0] = ht[c].x; hin[c][k4 * 4 + 1] = ht[c].y; hin[c][k4 * 4 + 2] = ht[c].z; hin[c][k4 * 4 + 3] = ht[c].w;
        }
      }
    }
    // output layer: y[c][o] = sum_k Woutt[k][o] * h[c][k] (+ bout)
    f4 y[C];
    y[0] = ld4(sw + lay.bo());
#pragma unroll
    for (int c = 1; c < C; ++c) y[c] = f4{0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const f4 w = ld4(sw + lay.wo() + k * 4);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float hv = hin[c][k];
        y[c].x += w.x * hv; y[c].y += w.y * hv; y[c].z += w.z * hv; y[c].w += w.w * hv;
      }
    }

    if (p.mode == MODE_FWD) {
      if (valid) {
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < p.n_out) p.jets_out[(int64_t)(c * p.n_out + o) * p.n_points + n] = (&y[c].x)[o];
      }
      return;
    }

    f4 yb[C];
    if (p.mode == MODE_FUSED || p.mode == MODE_LOSS) {
      residual_loss(p, n, valid, x, y, yb, lane);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        yb[c] = f4{0.f, 0.f, 0.f, 0.f};
        if (valid) {
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < p.n_out) (&yb[c].x)[o] = p.jets_bar[(int64_t)(c * p.n_out + o) * p.n_points + n];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) st4(a_row + c * 4, yb[c]);  // A row now holds ybar[c][o4]
  }

  // ---- phase B_l (l = L+1 .. 2): adjoint through linear layer l, then through
  // activation l-1.  Reads this lane's A row (ybar for l == L+1, zbar_l
  // otherwise) and z_{l-1} (global stash, or recomputed for l-1 == 1); writes
  // h_{l-1} to this lane's B row and keeps zbar_{l-1} in lane.zk. ------------
  PINN_HD static void phase_backward(const NarrowParams& p, const float* sw, float* wm, const float* gs, int lane_id,
                                     int l, Lane& lane) {
    const int L = p.n_hidden;
    const Layout lay(H, L);
    const int s = S();
    const float* a_row = wm + off_a() + lane_id * s;
    float* b_row = wm + off_b() + lane_id * s;
    float (&hbar)[C][H] = lane.zk;  // adjoint of h_{l-1}; overwritten in place by zbar_{l-1}
    if (l == L + 1) {
      f4 yb[C];
#pragma unroll
      for (int c = 0; c < C; ++c) yb[c] = ld4(a_row + c * 4);
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const f4 w = ld4(sw + lay.wo() + k * 4);
#pragma unroll
        for (int c = 0; c < C; ++c) hbar[c][k] = w.x * yb[c].x + w.y * yb[c].y + w.z * yb[c].z + w.w * yb[c].w;
      }
    } else {
      const float* Wt = sw + lay.wh(l);
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int k = 0; k < H; ++k) hbar[c][k] = 0.f;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
      for (int j4 = 0; j4 < H4; ++j4) {
        f4 zb[C];
        load_row_tile(a_row, j4, zb);
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const f4 w = ld4(Wt + k * H + j4 * 4);
#pragma unroll
          for (int c = 0; c < C; ++c)
            hbar[c][k] += w.x * zb[c].x + w.y * zb[c].y + w.z * zb[c].z + w.w * zb[c].w;
        }
      }
    }
    // through activation l-1
    const int lm = l - 1;
    float x[4];
    {
      const f4 xv = ld4(wm + off_xs() + lane_id * 4);
      x[0] = xv.x; x[1] = xv.y; x[2] = xv.z; x[3] = xv.w;
    }
#pragma unroll
    for (int k4 = 0; k4 < H4; ++k4) {
      f4 zt[C], ht[C];
      if (lm >= 2) {
#pragma unroll
        for (int c = 0; c < C; ++c) zt[c] = ld4(gs + stash_off(lm, c, k4, lane_id));
      } else {
        layer1_tile(sw, lay, p, x, k4, zt);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float z[C], hb[C], h[C], zb[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          z[c] = (&zt[c].x)[e];
          hb[c] = hbar[c][k4 * 4 + e];
          zb[c] = 0.f;
        }
        unit_bwd(z, hb, h, zb);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          (&ht[c].x)[e] = h[c];
          hbar[c][k4 * 4 + e] = zb[c];
        }
      }
      store_row_tile(b_row, k4, ht);
    }
  }

  // after the GEMM that consumed A: A row <- zbar kept in registers
  PINN_HD static void phase_store_zk(float* wm, int lane_id, const Lane& lane) {
    float* a_row = wm + off_a() + lane_id * S();
#pragma unroll
    for (int k4 = 0; k4 < H4; ++k4)
#pragma unroll
      for (int c = 0; c < C; ++c)
        st4(a_row + c * H + k4 * 4,
            f4{lane.zk[c][k4 * 4], lane.zk[c][k4 * 4 + 1], lane.zk[c][k4 * 4 + 2], lane.zk[c][k4 * 4 + 3]});
  }

  // ---- GEMM phases (cross-lane; run after __syncwarp) ----------------------
  // dWout[o][k] += sum_{n,c} ybar[n][c][o] * h_L[n][c][k];  dbout[o] += sum_n ybar[n][0][o]
  PINN_HD static void phase_gemm_out(const NarrowParams& p, float* wm, float* g, int lane_id) {
    const int L = p.n_hidden;
    const Layout lay(H, L);
    const int s = S();
    const float* yb = wm + off_a();
    const float* hb = wm + off_b();
    const int T = p.n_out * H4;
    // (accumulator targets are requested before the reduction loops: their L2 round trip hides behind them)
    const float gb = lane_id < p.n_out ? g[lay.bo() + lane_id] : 0.f;
    for (int t = lane_id; t < T; t += 32) {
      const int o = t / H4, kb = t % H4;
      float* gw = g + lay.wo() + (kb * 4) * 4 + o;
      const float g0 = gw[0], g1 = gw[4], g2 = gw[8], g3 = gw[12];
      f4 acc = f4{0.f, 0.f, 0.f, 0.f};
      for (int n = 0; n < 32; ++n) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float a = yb[n * s + c * 4 + o];
          const f4 b = ld4(hb + n * s + c * H + kb * 4);
          acc.x += a * b.x; acc.y += a * b.y; acc.z += a * b.z; acc.w += a * b.w;
        }
      }
      gw[0] = g0 + acc.x; gw[4] = g1 + acc.y; gw[8] = g2 + acc.z; gw[12] = g3 + acc.w;
    }
    if (lane_id < p.n_out) {
      float a = 0.f;
      for (int n = 0; n < 32; ++n) a += yb[n * s + lane_id];
      g[lay.bo() + lane_id] = gb + a;
    }
  }

  // dW_l[j][k] += sum_{n,c} zbar_l[n][c][j] * h_{l-1}[n][c][k];  db_l[j] += sum_n zbar_l[n][0][j]
  PINN_HD static void phase_gemm_hidden_fma(const NarrowParams& p, float* wm, float* g, int lane_id, int l) {
    const int L = p.n_hidden;
    const Layout lay(H, L);
    const int s = S();
    const float* A = wm + off_a();  // zbar_l
    const float* B = wm + off_b();  // h_{l-1}
    for (int t = lane_id; t < H4 * H4; t += 32) {
      const int ja = t / H4, kb = t % H4;
      f4 acc[4];  // acc[kk] = (j0..j3) for k = kb*4+kk
      f4 gv[4];   // the accumulator tile is requested up front: its L2 latency hides behind the loop
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        acc[kk] = f4{0.f, 0.f, 0.f, 0.f};
        gv[kk] = ld4(g + lay.wh(l) + (kb * 4 + kk) * H + ja * 4);
      }
#if defined(__CUDA_ARCH__)
#pragma unroll 2
#endif
      for (int n = 0; n < 32; ++n) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const f4 a = ld4(A + n * s + c * H + ja * 4);
          const f4 b = ld4(B + n * s + c * H + kb * 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float bv = (&b.x)[kk];
            acc[kk].x += a.x * bv; acc[kk].y += a.y * bv; acc[kk].z += a.z * bv; acc[kk].w += a.w * bv;
          }
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float* gp = g + lay.wh(l) + (kb * 4 + kk) * H + ja * 4;
        f4 v = gv[kk];
        v.x += acc[kk].x; v.y += acc[kk].y; v.z += acc[kk].z; v.w += acc[kk].w;
        st4(gp, v);
      }
    }
    for (int t = lane_id; t < H4; t += 32) {
      f4 a = f4{0.f, 0.f, 0.f, 0.f};
      for (int n = 0; n < 32; ++n) {
        const f4 v = ld4(A + n * s + t * 4);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      float* gp = g + lay.bh(l) + t * 4;
      f4 v = ld4(gp);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      st4(gp, v);
    }
  }

  // ---- the same contraction on the warp-level tensor path (mma.sync m16n8k8, 3xTF32 operand split) -----------------
  // D (M = units j, N = units k) += A (M x K) * B (K x N), K = (point, channel); one k-step is EIGHT points of one channel:
  // column t <-> point n0 + 2t, column t + 4 <-> point n0 + 2t + 1.  Tile row g <-> unit 2u, row g + 8 <-> unit 2u + 1
  // (u = 8 mt + g): a 64-bit load per point is (a0, a1) / (a2, a3); on the B side a 64-bit load is b0 (or b1) of two n8
  // tiles (even / odd units of u = 8 nb + g).  Rows are [c][H] with an odd quad stride, so the half-warp of a 64-bit load
  // (four g, four t) covers sixteen distinct 8-byte banks.  H = 20: the four left-over units are one n8 tile from scalar
  // loads whose column 4 carries ones for the value channel (bias gradient).  See NarrowWarpPair for the packed layout.
  static constexpr int NPAIR = H / 2;
  static constexpr int MT = (NPAIR + 7) / 8;
  static constexpr int NB2 = NPAIR / 8;
  static constexpr int NPL = NPAIR % 8;
  static constexpr int NTL = NPL == 0 ? 0 : (NPL <= 2 ? 1 : 2);
  static constexpr int NT = 2 * NB2 + NTL;
  static constexpr bool kBiasInMma = (NTL == 1);

  // fragments of lane (g = lane / 4, t = lane % 4) for k-step (channel c, points n0 .. n0 + 7)
  PINN_HD static void load_a_frags(const float* A, int s, int c, int n0, int lane_id, float (&a)[MT][4]) {
    const int gq = lane_id >> 2, t = lane_id & 3;
    const float* r0 = A + (n0 + 2 * t) * s + c * H;
    const float* r1 = r0 + s;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int u = 8 * mt + gq;
      f2 v0 = f2{0.f, 0.f}, v1 = f2{0.f, 0.f};
      if (8 * mt + 7 < NPAIR || u < NPAIR) { v0 = ld2(r0 + 2 * u); v1 = ld2(r1 + 2 * u); }
      a[mt][0] = v0.x; a[mt][1] = v0.y; a[mt][2] = v1.x; a[mt][3] = v1.y;
    }
  }
  PINN_HD static void load_b_frags(const float* B, int s, int c, int n0, int lane_id, float (&b)[NT][2]) {
    const int gq = lane_id >> 2, t = lane_id & 3;
    const float* r0 = B + (n0 + 2 * t) * s + c * H;
    const float* r1 = r0 + s;
#pragma unroll
    for (int nb = 0; nb < NB2 + (NTL == 2 ? 1 : 0); ++nb) {
      const int u = 8 * nb + gq;
      f2 v0 = f2{0.f, 0.f}, v1 = f2{0.f, 0.f};
      if (nb < NB2 || u < NPAIR) { v0 = ld2(r0 + 2 * u); v1 = ld2(r1 + 2 * u); }
      b[2 * nb][0] = v0.x; b[2 * nb + 1][0] = v0.y; b[2 * nb][1] = v1.x; b[2 * nb + 1][1] = v1.y;
    }
    if (NTL == 1) {
      float v0 = 0.f, v1 = 0.f;
      if (gq < 2 * NPL) { v0 = r0[16 * NB2 + gq]; v1 = r1[16 * NB2 + gq]; }
      else if (gq == 4 && c == 0) { v0 = 1.0f; v1 = 1.0f; }  // bias column: ones against the value channel
      b[NT - 1][0] = v0; b[NT - 1][1] = v1;
    }
  }
  // float offset in the gradient accumulators of the element pair (rows g, g + 8; column cn) of tile (mt, nt); -1: padding
  PINN_HD static int dw_target(const Layout& lay, int l, int mt, int nt, int gq, int cn) {
    const int ua = 8 * mt + gq;
    if (ua >= NPAIR) return -1;
    if (NTL == 1 && nt == NT - 1) {
      if (cn < 2 * NPL) return lay.wh(l) + (16 * NB2 + cn) * H + 2 * ua;
      return cn == 4 ? lay.bh(l) + 2 * ua : -1;
    }
    const int ub = 8 * (nt >> 1) + cn;
    if (ub >= NPAIR) return -1;
    return lay.wh(l) + (2 * ub + (nt & 1)) * H + 2 * ua;
  }

  // The scalar program keeps the FFMA GEMM by default: with an odd channel count a k-step is eight points of ONE channel
  // (4 C k-steps of 18 MMAs per layer: 360 for C = 5 against 288 for the packed program's C = 4) and the Poisson step
  // measured 9 % SLOWER on the tensor path (1.947 vs 1.788 ms per 2^20-point launch); PINN_NARROW_SCALAR_DW_MMA selects it.
  PINN_HD static void phase_gemm_hidden(const NarrowParams& p, float* wm, float* g, int lane_id, int l) {
#if defined(PINN_NARROW_DW_FMA) || !defined(PINN_NARROW_SCALAR_DW_MMA)
    phase_gemm_hidden_fma(p, wm, g, lane_id, l);
#else
    const int L = p.n_hidden;
    const Layout lay(H, L);
    const int s = S();
    const float* A = wm + off_a();  // zbar_l
    const float* B = wm + off_b();  // h_{l-1}
    const int gq = lane_id >> 2, t = lane_id & 3;
    float acc[MT][NT][4];
    // this lane's accumulator targets are requested up front: their L2 latency hides behind the MMA loop (and the loads
    // cannot be ordered behind the stores of other targets, which the compiler must assume may alias)
    int goff[MT][NT][2];
    f2 gv[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          goff[mt][nt][w] = dw_target(lay, l, mt, nt, gq, 2 * t + w);
          gv[mt][nt][w] = goff[mt][nt][w] >= 0 ? ld2(g + goff[mt][nt][w]) : f2{0.f, 0.f};
        }
      }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
#pragma unroll 2
      for (int n0 = 0; n0 < 32; n0 += 8) {
        float a[MT][4], b[NT][2];
        load_a_frags(A, s, c, n0, lane_id, a);
        load_b_frags(B, s, c, n0, lane_id, b);
        uint32_t ah[MT][4], al[MT][4], bh[NT][2], bl[NT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int e = 0; e < 4; ++e) split_tf32(a[mt][e], ah[mt][e], al[mt][e]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) split_tf32(b[nt][e], bh[nt][e], bl[nt][e]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_m16n8k8_tf32(acc[mt][nt], al[mt], bh[nt]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_m16n8k8_tf32(acc[mt][nt], ah[mt], bl[nt]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_m16n8k8_tf32(acc[mt][nt], ah[mt], bh[nt]);
      }
    }
#else
    // host emulation (tests/emu): this lane's D elements from the fragments the OTHER lanes would hold
    for (int c = 0; c < C; ++c)
      for (int n0 = 0; n0 < 32; n0 += 8) {
        float af[4][MT][4], bf[2][4][NT][2];
        for (int tt = 0; tt < 4; ++tt) {
          load_a_frags(A, s, c, n0, gq * 4 + tt, af[tt]);
          for (int w = 0; w < 2; ++w) load_b_frags(B, s, c, n0, (2 * t + w) * 4 + tt, bf[w][tt]);
        }
        for (int mt = 0; mt < MT; ++mt)
          for (int nt = 0; nt < NT; ++nt)
            for (int e = 0; e < 4; ++e) {
              const int row_hi = e >> 1, w = e & 1;
              float d = 0.f;
              for (int kk = 0; kk < 8; ++kk)
                d += mul_tf32x3(af[kk & 3][mt][row_hi + 2 * (kk >> 2)], bf[w][kk & 3][nt][kk >> 2]);
              acc[mt][nt][e] += d;
            }
      }
#endif
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          if (goff[mt][nt][w] >= 0) {
            f2 v = gv[mt][nt][w];
            v.x += acc[mt][nt][w];
            v.y += acc[mt][nt][2 + w];
            st2(g + goff[mt][nt][w], v);
          }
        }
    if (!kBiasInMma) {
      for (int tq = lane_id; tq < H4; tq += 32) {
        f4 a = f4{0.f, 0.f, 0.f, 0.f};
        for (int n = 0; n < 32; ++n) {
          const f4 v = ld4(A + n * s + tq * 4);
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        float* gp = g + lay.bh(l) + tq * 4;
        f4 v = ld4(gp);
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        st4(gp, v);
      }
    }
#endif
  }

  // dW1[j][d] += sum_n zbar_val[n][j] x_d[n] + [d == first_in[i]] sum_n zbar_i[n][j];  db1[j] += sum_n zbar_val[n][j]
  PINN_HD static void phase_gemm_first(const NarrowParams& p, float* wm, float* g, int lane_id) {
    const int L = p.n_hidden;
    const Layout lay(H, L);
    const int s = S();
    const float* X = wm + off_a();  // zbar_1
    const float* xs = wm + off_xs();
    for (int j = lane_id; j < H; j += 32) {
      // targets first (five independent loads whose L2 latency hides behind the loop); the first-order sums are folded
      // into the column of their input in registers, so no read-modify-write is ordered behind another
      float gd[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) gd[d] = g[lay.w1() + d * H + j];
      const float gb = g[lay.b1() + j];
      float ad[4] = {0.f, 0.f, 0.f, 0.f};
      float af[NF > 0 ? NF : 1];
#pragma unroll
      for (int i = 0; i < NF; ++i) af[i] = 0.f;
      float ab = 0.f;
      for (int n = 0; n < 32; ++n) {
        const float zv = X[n * s + j];
        const f4 xv = ld4(xs + n * 4);
        ad[0] += zv * xv.x; ad[1] += zv * xv.y; ad[2] += zv * xv.z; ad[3] += zv * xv.w;
        ab += zv;
#pragma unroll
        for (int i = 0; i < NF; ++i) af[i] += X[n * s + (1 + i) * H + j];
      }
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        float v = gd[d] + ad[d];
#pragma unroll
        for (int i = 0; i < NF; ++i)
          if (p.jets.first_in[i] == d) v += af[i];
        g[lay.w1() + d * H + j] = v;
      }
      g[lay.b1() + j] = gb + ab;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// The same warp program for jets with an EVEN number of channels, with every jet stored as channel
// pairs (2cp, 2cp+1) -- in registers, in the shared-memory rows and in the L2 stash -- so that the
// three hot loops run on packed FMAs (FFMA2, half the FMA issue slots):
//   forward   acc[cp][j] += W[j][k] (scalar, broadcast) * h[cp][k]
//   reverse   hbar[cp][k] += W[j][k] (scalar) * zbar[cp][j]
//   dW GEMM   acc[j][k] (pair of per-channel partial sums, added at the end) += zbar[cp][j] * h[cp][k]
// Row layout (floats): [cp][half h][unit quad m][4] with the 4 floats = (unit 4m+2h: ch 2cp, ch 2cp+1,
// unit 4m+2h+1: ch 2cp, ch 2cp+1).  The 4-unit tile of a channel pair is the two 16-byte chunks (h = 0, 1);
// chunks of consecutive unit quads are adjacent, so the five (H = 20) or eight (H = 32) distinct chunks
// that the lanes of a GEMM load touch fall into different bank groups.
template <int H_, int NF_, int NS_, int ACT_, int X_ = 0>
struct NarrowWarpPair {
  using Base = NarrowWarp<H_, NF_, NS_, ACT_, X_>;
  static constexpr int H = H_, NF = NF_, NS = NS_, ACT = ACT_, X = X_;
  static constexpr int C = Base::C;
  static constexpr int CP = C / 2;
  static constexpr int H4 = H / 4;
  static constexpr int kMaxWarps = PINN_PAIR_WARPS;
  static_assert(C % 2 == 0, "channel pairs need an even channel count");

  PINN_HD static int S() { return row_stride(C * H); }
  PINN_HD static int off_a() { return 0; }
  PINN_HD static int off_b() { return 32 * S(); }
  PINN_HD static int off_xs() { return 2 * 32 * S(); }
  PINN_HD static int warp_floats(int) { return off_xs() + 32 * 4; }
  // per-warp global stash (floats): [(l-2)][cp][k4][half][lane][4]
  PINN_HD static int stash_floats(int L) { return (L - 1) * C * H * 32; }
  PINN_HD static int stash_off(int l, int cp, int k4, int half, int lane) {
    return ((((l - 2) * CP + cp) * H4 + k4) * 2 + half) * 128 + lane * 4;
  }
  PINN_HD static int chunk_off(int cp, int h, int m) { return ((cp * 2 + h) * H4 + m) * 4; }
  // float offset of (channel c, unit j) inside a row
  PINN_HD static int elem_off(int c, int j) { return chunk_off(c >> 1, (j >> 1) & 1, j >> 2) + (j & 1) * 2 + (c & 1); }

  struct Lane {
    float loss;
    f2 zk[CP][H];  // zbar of the layer just differentiated, as channel pairs
  };

  PINN_HD static float get(const f2& v, int half) { return half ? v.y : v.x; }

  // 4 units x one channel pair <-> the two chunks of a row
  PINN_HD static void load_pairs(const float* row, int cp, int m, f2 (&t)[4]) {
    const f4 a = ld4(row + chunk_off(cp, 0, m)), b = ld4(row + chunk_off(cp, 1, m));
    t[0] = f2{a.x, a.y}; t[1] = f2{a.z, a.w}; t[2] = f2{b.x, b.y}; t[3] = f2{b.z, b.w};
  }
  PINN_HD static void store_pairs(float* row, int cp, int m, const f2 (&t)[4]) {
    st4(row + chunk_off(cp, 0, m), f4{t[0].x, t[0].y, t[1].x, t[1].y});
    st4(row + chunk_off(cp, 1, m), f4{t[2].x, t[2].y, t[3].x, t[3].y});
  }

  PINN_HD static void residual_loss(const NarrowParams& p, int64_t n, bool valid, const float (&x)[4], const f4 (&y)[C],
                                    f4 (&yb)[C], Lane& lane) {
    typename Base::Lane bl;
    bl.loss = lane.loss;
    Base::residual_loss(p, n, valid, x, y, yb, bl);
    lane.loss = bl.loss;
  }

  PINN_HD static void phase_forward(const NarrowParams& p, const float* sw, float* wm, float* gs, int lane_id,
                                    int64_t tile, Lane& lane) {
    const int L = p.n_hidden;
    const Layout lay(H, L);
    const int s = S();
    const int64_t n = tile * 32 + lane_id;
    const bool valid = n < p.n_points;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (d < p.n_in && valid) x[d] = p.coords[d][n];
    st4(wm + off_xs() + lane_id * 4, f4{x[0], x[1], x[2], x[3]});
    float* a_row = wm + off_a() + lane_id * s;

    f2 hin[CP][H];
#pragma unroll
    for (int k4 = 0; k4 < H4; ++k4) {  // hidden layer 1
      f4 zt[C];
      Base::layer1_tile(sw, lay, p, x, k4, zt);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float z[C], h[C];
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = (&zt[c].x)[e];
        Base::unit_fwd(z, h);
#pragma unroll
        for (int cp = 0; cp < CP; ++cp) hin[cp][k4 * 4 + e] = f2{h[2 * cp], h[2 * cp + 1]};
      }
    }
    for (int l = 2; l <= L; ++l) {  // hidden layers 2..L
      const float* Wt = sw + lay.wh(l);
      const float* bl = sw + lay.bh(l);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
      for (int j4 = 0; j4 < H4; ++j4) {
        f2 acc[CP][4];  // acc[cp][e]: pre-activation of unit j4*4+e, channels (2cp, 2cp+1)
        {
          const f4 b = ld4(bl + j4 * 4);
          acc[0][0] = f2{b.x, 0.f}; acc[0][1] = f2{b.y, 0.f}; acc[0][2] = f2{b.z, 0.f}; acc[0][3] = f2{b.w, 0.f};
        }
#pragma unroll
        for (int cp = 1; cp < CP; ++cp)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[cp][e] = f2{0.f, 0.f};
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const f4 w = ld4(Wt + k * H + j4 * 4);
#pragma unroll
          for (int cp = 0; cp < CP; ++cp) {
            acc[cp][0] = fma2s(hin[cp][k], w.x, acc[cp][0]);
            acc[cp][1] = fma2s(hin[cp][k], w.y, acc[cp][1]);
            acc[cp][2] = fma2s(hin[cp][k], w.z, acc[cp][2]);
            acc[cp][3] = fma2s(hin[cp][k], w.w, acc[cp][3]);
          }
        }
        if (p.mode == MODE_FUSED || p.mode == MODE_BWD) {  // the stash only feeds the reverse pass
#pragma unroll
          for (int cp = 0; cp < CP; ++cp) {
            st4(gs + stash_off(l, cp, j4, 0, lane_id), f4{acc[cp][0].x, acc[cp][0].y, acc[cp][1].x, acc[cp][1].y});
            st4(gs + stash_off(l, cp, j4, 1, lane_id), f4{acc[cp][2].x, acc[cp][2].y, acc[cp][3].x, acc[cp][3].y});
          }
        }
        f2 ht[CP][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float z[C], h[C];
#pragma unroll
          for (int c = 0; c < C; ++c) z[c] = get(acc[c >> 1][e], c & 1);
          Base::unit_fwd(z, h);
#pragma unroll
          for (int cp = 0; cp < CP; ++cp) ht[cp][e] = f2{h[2 * cp], h[2 * cp + 1]};
        }
#pragma unroll
        for (int cp = 0; cp < CP; ++cp) store_pairs(a_row, cp, j4, ht[cp]);  // scratch: hin is still being read
      }
#pragma unroll
      for (int k4 = 0; k4 < H4; ++k4)
#pragma unroll
        for (int cp = 0; cp < CP; ++cp) {
          f2 t[4];
          load_pairs(a_row, cp, k4, t);
          hin[cp][k4 * 4 + 0] = t[0]; hin[cp][k4 * 4 + 1] = t[1]; hin[cp][k4 * 4 + 2] = t[2]; hin[cp][k4 * 4 + 3] = t[3];
        }
    }
    // output layer, packed over the output pairs (o0,o1), (o2,o3)
    f2 y2[C][2];
    {
      const f4 b = ld4(sw + lay.bo());
      y2[0][0] = f2{b.x, b.y}; y2[0][1] = f2{b.z, b.w};
    }
#pragma unroll
    for (int c = 1; c < C; ++c) { y2[c][0] = f2{0.f, 0.f}; y2[c][1] = f2{0.f, 0.f}; }
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const f4 w = ld4(sw + lay.wo() + k * 4);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float hv = get(hin[c >> 1][k], c & 1);
        y2[c][0] = fma2s(f2{w.x, w.y}, hv, y2[c][0]);
        y2[c][1] = fma2s(f2{w.z, w.w}, hv, y2[c][1]);
      }
    }
    f4 y[C];
#pragma unroll
    for (int c = 0; c < C; ++c) y[c] = f4{y2[c][0].x, y2[c][0].y, y2[c][1].x, y2[c][1].y};

    if (p.mode == MODE_FWD) {
      if (valid) {
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < p.n_out) p.jets_out[(int64_t)(c * p.n_out + o) * p.n_points + n] = (&y[c].x)[o];
      }
      return;
    }

    f4 yb[C];
    if (p.mode == MODE_FUSED || p.mode == MODE_LOSS) {
      residual_loss(p, n, valid, x, y, yb, lane);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        yb[c] = f4{0.f, 0.f, 0.f, 0.f};
        if (valid) {
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < p.n_out) (&yb[c].x)[o] = p.jets_bar[(int64_t)(c * p.n_out + o) * p.n_points + n];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) st4(a_row + c * 4, yb[c]);  // A row now holds ybar[c][o4]
  }

  PINN_HD static void phase_backward(const NarrowParams& p, const float* sw, float* wm, const float* gs, int lane_id,
                                     int l, Lane& lane) {
    const int L = p.n_hidden;
    const Layout lay(H, L);
    const int s = S();
    const float* a_row = wm + off_a() + lane_id * s;
    float* b_row = wm + off_b() + lane_id * s;
    f2 (&hbar)[CP][H] = lane.zk;  // adjoint of h_{l-1} (channel pairs); overwritten in place by zbar_{l-1}
    if (l == L + 1) {
      f4 yb[C];
#pragma unroll
      for (int c = 0; c < C; ++c) yb[c] = ld4(a_row + c * 4);
      f2 ybp[CP][4];  // (ybar[2cp][o], ybar[2cp+1][o])
#pragma unroll
      for (int cp = 0; cp < CP; ++cp)
#pragma unroll
        for (int o = 0; o < 4; ++o) ybp[cp][o] = f2{(&yb[2 * cp].x)[o], (&yb[2 * cp + 1].x)[o]};
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const f4 w = ld4(sw + lay.wo() + k * 4);
#pragma unroll
        for (int cp = 0; cp < CP; ++cp) {
          f2 t = f2{0.f, 0.f};
          t = fma2s(ybp[cp][0], w.x, t);
          t = fma2s(ybp[cp][1], w.y, t);
          t = fma2s(ybp[cp][2], w.z, t);
          t = fma2s(ybp[cp][3], w.w, t);
          hbar[cp][k] = t;
        }
      }
    } else {
      const float* Wt = sw + lay.wh(l);
#pragma unroll
      for (int cp = 0; cp < CP; ++cp)
#pragma unroll
        for (int k = 0; k < H; ++k) hbar[cp][k] = f2{0.f, 0.f};
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
      for (int j4 = 0; j4 < H4; ++j4) {
        f2 zb[CP][4];
#pragma unroll
        for (int cp = 0; cp < CP; ++cp) load_pairs(a_row, cp, j4, zb[cp]);
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const f4 w = ld4(Wt + k * H + j4 * 4);
#pragma unroll
          for (int cp = 0; cp < CP; ++cp) {
            f2 t = hbar[cp][k];
            t = fma2s(zb[cp][0], w.x, t);
            t = fma2s(zb[cp][1], w.y, t);
            t = fma2s(zb[cp][2], w.z, t);
            t = fma2s(zb[cp][3], w.w, t);
            hbar[cp][k] = t;
          }
        }
      }
    }
    // through activation l-1
    const int lm = l - 1;
    float x[4];
    {
      const f4 xv = ld4(wm + off_xs() + lane_id * 4);
      x[0] = xv.x; x[1] = xv.y; x[2] = xv.z; x[3] = xv.w;
    }
#pragma unroll
    for (int k4 = 0; k4 < H4; ++k4) {
      f4 zt[C];      // lm == 1: per-channel tiles from the first layer
      f2 zp[CP][4];  // lm >= 2: channel pairs from the stash
      if (lm >= 2) {
#pragma unroll
        for (int cp = 0; cp < CP; ++cp) {
          const f4 a = ld4(gs + stash_off(lm, cp, k4, 0, lane_id)), b = ld4(gs + stash_off(lm, cp, k4, 1, lane_id));
          zp[cp][0] = f2{a.x, a.y}; zp[cp][1] = f2{a.z, a.w}; zp[cp][2] = f2{b.x, b.y}; zp[cp][3] = f2{b.z, b.w};
        }
      } else {
        Base::layer1_tile(sw, lay, p, x, k4, zt);
      }
      f2 ht[CP][4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float z[C], hb[C], h[C], zb[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          z[c] = (lm >= 2) ? get(zp[c >> 1][e], c & 1) : (&zt[c].x)[e];
          hb[c] = get(hbar[c >> 1][k4 * 4 + e], c & 1);
          zb[c] = 0.f;
        }
        Base::unit_bwd(z, hb, h, zb);
#pragma unroll
        for (int cp = 0; cp < CP; ++cp) {
          ht[cp][e] = f2{h[2 * cp], h[2 * cp + 1]};
          hbar[cp][k4 * 4 + e] = f2{zb[2 * cp], zb[2 * cp + 1]};
        }
      }
#pragma unroll
      for (int cp = 0; cp < CP; ++cp) store_pairs(b_row, cp, k4, ht[cp]);
    }
  }

  PINN_HD static void phase_store_zk(float* wm, int lane_id, const Lane& lane) {
    float* a_row = wm + off_a() + lane_id * S();
#pragma unroll
    for (int k4 = 0; k4 < H4; ++k4)
#pragma unroll
      for (int cp = 0; cp < CP; ++cp) {
        const f2 t[4] = {lane.zk[cp][k4 * 4], lane.zk[cp][k4 * 4 + 1], lane.zk[cp][k4 * 4 + 2], lane.zk[cp][k4 * 4 + 3]};
        store_pairs(a_row, cp, k4, t);
      }
  }

  PINN_HD static void phase_gemm_out(const NarrowParams& p, float* wm, float* g, int lane_id) {
    const int L = p.n_hidden;
    const Layout lay(H, L);
    const int s = S();
    const float* yb = wm + off_a();
    const float* hb = wm + off_b();
    const int T = p.n_out * H4;
    const float gb = lane_id < p.n_out ? g[lay.bo() + lane_id] : 0.f;  // targets first: see NarrowWarp::phase_gemm_out
    for (int t = lane_id; t < T; t += 32) {
      const int o = t / H4, kb = t % H4;
      float* gw = g + lay.wo() + (kb * 4) * 4 + o;
      const float g0 = gw[0], g1 = gw[4], g2 = gw[8], g3 = gw[12];
      f2 acc[4];  // per unit k: partial sums of the two channels of a pair
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] = f2{0.f, 0.f};
      for (int n = 0; n < 32; ++n) {
#pragma unroll
        for (int cp = 0; cp < CP; ++cp) {
          const f2 a = f2{yb[n * s + (2 * cp) * 4 + o], yb[n * s + (2 * cp + 1) * 4 + o]};
          f2 b[4];
          load_pairs(hb + n * s, cp, kb, b);
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[e] = fma2(a, b[e], acc[e]);
        }
      }
      gw[0] = g0 + (acc[0].x + acc[0].y); gw[4] = g1 + (acc[1].x + acc[1].y);
      gw[8] = g2 + (acc[2].x + acc[2].y); gw[12] = g3 + (acc[3].x + acc[3].y);
    }
    if (lane_id < p.n_out) {
      float a = 0.f;
      for (int n = 0; n < 32; ++n) a += yb[n * s + lane_id];
      g[lay.bo() + lane_id] = gb + a;
    }
  }

  PINN_HD static void phase_gemm_hidden_fma(const NarrowParams& p, float* wm, float* g, int lane_id, int l) {
    const int L = p.n_hidden;
    const Layout lay(H, L);
    const int s = S();
    const float* A = wm + off_a();  // zbar_l
    const float* B = wm + off_b();  // h_{l-1}
    for (int t = lane_id; t < H4 * H4; t += 32) {
      const int ja = t / H4, kb = t % H4;
      f2 acc[4][4];  // acc[kk][jj]: per-channel partial sums of dW[ja*4+jj][kb*4+kk]
      f4 gv[4];      // the accumulator tile is requested up front: its L2 latency hides behind the loop
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        gv[kk] = ld4(g + lay.wh(l) + (kb * 4 + kk) * H + ja * 4);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[kk][jj] = f2{0.f, 0.f};
      }
#if defined(__CUDA_ARCH__)
#pragma unroll 2
#endif
      for (int n = 0; n < 32; ++n) {
#pragma unroll
        for (int cp = 0; cp < CP; ++cp) {
          f2 a[4], b[4];
          load_pairs(A + n * s, cp, ja, a);
          load_pairs(B + n * s, cp, kb, b);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[kk][jj] = fma2(a[jj], b[kk], acc[kk][jj]);
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float* gp = g + lay.wh(l) + (kb * 4 + kk) * H + ja * 4;
        f4 v = gv[kk];
        v.x += acc[kk][0].x + acc[kk][0].y; v.y += acc[kk][1].x + acc[kk][1].y;
        v.z += acc[kk][2].x + acc[kk][2].y; v.w += acc[kk][3].x + acc[kk][3].y;
        st4(gp, v);
      }
    }
    for (int t = lane_id; t < H4; t += 32) {
      f4 a = f4{0.f, 0.f, 0.f, 0.f};
      for (int n = 0; n < 32; ++n) {  // value channel = first half of pair 0
        const f4 u = ld4(A + n * s + chunk_off(0, 0, t)), v = ld4(A + n * s + chunk_off(0, 1, t));
        a.x += u.x; a.y += u.z; a.z += v.x; a.w += v.z;
      }
      float* gp = g + lay.bh(l) + t * 4;
      f4 v = ld4(gp);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      st4(gp, v);
    }
  }

  // ---- dW_l on the warp-level tensor path (mma.sync m16n8k8, 3xTF32 operand split) -------------------------------
  // dW_l[j][k] = sum over the 32 points n and C channels c of zbar_l[n][c][j] * h_{l-1}[n][c][k] as D (M = units j, N =
  // units k) += A (M x K) * B (K x N) with the contraction index K = (point, channel).  One k-step (K = 8) is four points
  // x the two channels of a pair: column t <-> (point P(t), channel 2cp), column t + 4 <-> (P(t), 2cp + 1).  A row buffer
  // keeps the units of a channel pair in 16-byte chunks (unit 2u: both channels, unit 2u + 1: both channels), so with tile
  // row g <-> even unit of chunk q = 8 mt + g and row g + 8 <-> its odd unit, ONE 128-bit load is the lane's whole A
  // fragment (a0, a2, a1, a3), and on the B side one 128-bit load is the fragments (b0, b1) of two n8 tiles (even units /
  // odd units of chunks 8 nb + g).  Every operand element is read once per layer (20 KB per tile and layer instead of
  // 100 KB for 4x4 register tiles: the FFMA2 GEMM was bound by the shared-memory return path).
  // Points of a k-step are P(t) = n0 + 2t: with an odd quad stride the quarter-warp (two g, four t) then lands on eight
  // distinct 16-byte bank groups because chunks of consecutive g are adjacent.
  // H = 20: 10 chunks -> two m16 tiles (the second with two valid row pairs), one chunk group (two n8 tiles) + the two
  // left-over chunks as one n8 tile loaded with 64-bit loads (columns 0..3), whose free column 4 carries ones for the
  // value channel: D[:, 4] of that tile is the bias gradient.  H = 32: two m16 tiles, two groups (four n8 tiles).
  static constexpr int NPAIR = H / 2;          // 16-byte chunks (unit pairs) per channel pair
  static constexpr int MT = (NPAIR + 7) / 8;   // m16 tiles
  static constexpr int NB2 = NPAIR / 8;        // full chunk groups on the N side
  static constexpr int NPL = NPAIR % 8;        // left-over chunks on the N side
  static constexpr int NTL = NPL == 0 ? 0 : (NPL <= 2 ? 1 : 2);
  static constexpr int NT = 2 * NB2 + NTL;     // n8 tiles
  static constexpr bool kBiasInMma = (NTL == 1);
  PINN_HD static int kstep_point(int ks, int t) { return (ks >> 1) * 8 + (ks & 1) + 2 * t; }
  PINN_HD static int chunk_unit(int q) { return 4 * (q % H4) + 2 * (q / H4); }  // even unit of chunk q

  PINN_HD static void load_a_frags(const float* A, int s, int cp, int ks, int lane_id, float (&a)[MT][4]) {
    const int gq = lane_id >> 2, t = lane_id & 3;
    const float* row = A + kstep_point(ks, t) * s + cp * 2 * H;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int q = 8 * mt + gq;
      f4 v = f4{0.f, 0.f, 0.f, 0.f};
      if (8 * mt + 7 < NPAIR || q < NPAIR) v = ld4(row + q * 4);
      a[mt][0] = v.x; a[mt][2] = v.y; a[mt][1] = v.z; a[mt][3] = v.w;
    }
  }
  PINN_HD static void load_b_frags(const float* B, int s, int cp, int ks, int lane_id, float (&b)[NT][2]) {
    const int gq = lane_id >> 2, t = lane_id & 3;
    const float* row = B + kstep_point(ks, t) * s + cp * 2 * H;
#pragma unroll
    for (int nb = 0; nb < NB2 + (NTL == 2 ? 1 : 0); ++nb) {
      const int q = 8 * nb + gq;
      f4 v = f4{0.f, 0.f, 0.f, 0.f};
      if (nb < NB2 || q < NPAIR) v = ld4(row + q * 4);
      b[2 * nb][0] = v.x; b[2 * nb][1] = v.y; b[2 * nb + 1][0] = v.z; b[2 * nb + 1][1] = v.w;
    }
    if (NTL == 1) {
      f2 v = f2{0.f, 0.f};
      if (gq < 2 * NPL) v = ld2(row + (8 * NB2 + (gq >> 1)) * 4 + (gq & 1) * 2);
      else if (gq == 4 && cp == 0) v.x = 1.0f;  // bias column: ones against the value channel
      b[NT - 1][0] = v.x; b[NT - 1][1] = v.y;
    }
  }
  // float offset in the gradient accumulators of the element pair (rows g, g + 8; column cn) of tile (mt, nt); -1: padding
  PINN_HD static int dw_target(const Layout& lay, int l, int mt, int nt, int gq, int cn) {
    const int qa = 8 * mt + gq;
    if (qa >= NPAIR) return -1;
    const int j = chunk_unit(qa);
    if (NTL == 1 && nt == NT - 1) {
      if (cn < 2 * NPL) return lay.wh(l) + (chunk_unit(8 * NB2 + (cn >> 1)) + (cn & 1)) * H + j;
      return cn == 4 ? lay.bh(l) + j : -1;
    }
    const int qb = 8 * (nt >> 1) + cn;
    if (qb >= NPAIR) return -1;
    return lay.wh(l) + (chunk_unit(qb) + (nt & 1)) * H + j;
  }

  PINN_HD static void phase_gemm_hidden(const NarrowParams& p, float* wm, float* g, int lane_id, int l) {
#if defined(PINN_NARROW_DW_FMA)
    phase_gemm_hidden_fma(p, wm, g, lane_id, l);
#else
    const int L = p.n_hidden;
    const Layout lay(H, L);
    const int s = S();
    const float* A = wm + off_a();  // zbar_l
    const float* B = wm + off_b();  // h_{l-1}
    const int gq = lane_id >> 2, t = lane_id & 3;
    float acc[MT][NT][4];
    // this lane's accumulator targets are requested up front: their L2 latency hides behind the MMA loop (and the loads
    // cannot be ordered behind the stores of other targets, which the compiler must assume may alias)
    int goff[MT][NT][2];
    f2 gv[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          goff[mt][nt][w] = dw_target(lay, l, mt, nt, gq, 2 * t + w);
          gv[mt][nt][w] = goff[mt][nt][w] >= 0 ? ld2(g + goff[mt][nt][w]) : f2{0.f, 0.f};
        }
      }
#if defined(__CUDA_ARCH__)
    constexpr int kUnroll = PINN_DW_UNROLL;
#pragma unroll 1
    for (int cp = 0; cp < CP; ++cp) {
#pragma unroll kUnroll
      for (int ks = 0; ks < 8; ++ks) {
        float a[MT][4], b[NT][2];
        load_a_frags(A, s, cp, ks, lane_id, a);
        load_b_frags(B, s, cp, ks, lane_id, b);
        uint32_t ah[MT][4], al[MT][4], bh[NT][2], bl[NT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int e = 0; e < 4; ++e) split_tf32(a[mt][e], ah[mt][e], al[mt][e]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) split_tf32(b[nt][e], bh[nt][e], bl[nt][e]);
        // term-major: consecutive MMAs go to different accumulators (small terms first)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_m16n8k8_tf32(acc[mt][nt], al[mt], bh[nt]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_m16n8k8_tf32(acc[mt][nt], ah[mt], bl[nt]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_m16n8k8_tf32(acc[mt][nt], ah[mt], bh[nt]);
      }
    }
#else
    // host emulation (tests/emu): this lane's D elements from the fragments the OTHER lanes would hold
    for (int cp = 0; cp < CP; ++cp)
      for (int ks = 0; ks < 8; ++ks) {
        float af[4][MT][4], bf[2][4][NT][2];
        for (int tt = 0; tt < 4; ++tt) {
          load_a_frags(A, s, cp, ks, gq * 4 + tt, af[tt]);
          for (int w = 0; w < 2; ++w) load_b_frags(B, s, cp, ks, (2 * t + w) * 4 + tt, bf[w][tt]);
        }
        for (int mt = 0; mt < MT; ++mt)
          for (int nt = 0; nt < NT; ++nt)
            for (int e = 0; e < 4; ++e) {
              const int row_hi = e >> 1, w = e & 1;
              float d = 0.f;
              for (int kk = 0; kk < 8; ++kk)
                d += mul_tf32x3(af[kk & 3][mt][row_hi + 2 * (kk >> 2)], bf[w][kk & 3][nt][kk >> 2]);
              acc[mt][nt][e] += d;
            }
      }
#endif
    // accumulate into this warp's gradient scratch: (rows g, g+8) = (even, odd unit of a chunk) are adjacent in Wt[k][j]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          if (goff[mt][nt][w] >= 0) {
            f2 v = gv[mt][nt][w];
            v.x += acc[mt][nt][w];
            v.y += acc[mt][nt][2 + w];
            st2(g + goff[mt][nt][w], v);
          }
        }
    if (!kBiasInMma) {
      for (int tq = lane_id; tq < H4; tq += 32) {
        f4 a = f4{0.f, 0.f, 0.f, 0.f};
        for (int n = 0; n < 32; ++n) {  // value channel = first half of pair 0
          const f4 u = ld4(A + n * s + chunk_off(0, 0, tq)), v = ld4(A + n * s + chunk_off(0, 1, tq));
          a.x += u.x; a.y += u.z; a.z += v.x; a.w += v.z;
        }
        float* gp = g + lay.bh(l) + tq * 4;
        f4 v = ld4(gp);
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        st4(gp, v);
      }
    }
#endif
  }

  PINN_HD static void phase_gemm_first(const NarrowParams& p, float* wm, float* g, int lane_id) {
    const int L = p.n_hidden;
    const Layout lay(H, L);
    const int s = S();
    const float* X = wm + off_a();  // zbar_1
    const float* xs = wm + off_xs();
    for (int j = lane_id; j < H; j += 32) {
      // targets first, first-order sums folded in registers: see NarrowWarp::phase_gemm_first
      float gd[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) gd[d] = g[lay.w1() + d * H + j];
      const float gb = g[lay.b1() + j];
      float ad[4] = {0.f, 0.f, 0.f, 0.f};
      float af[NF > 0 ? NF : 1];
#pragma unroll
      for (int i = 0; i < NF; ++i) af[i] = 0.f;
      float ab = 0.f;
      for (int n = 0; n < 32; ++n) {
        const float zv = X[n * s + elem_off(0, j)];
        const f4 xv = ld4(xs + n * 4);
        ad[0] += zv * xv.x; ad[1] += zv * xv.y; ad[2] += zv * xv.z; ad[3] += zv * xv.w;
        ab += zv;
#pragma unroll
        for (int i = 0; i < NF; ++i) af[i] += X[n * s + elem_off(1 + i, j)];
      }
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        float v = gd[d] + ad[d];
#pragma unroll
        for (int i = 0; i < NF; ++i)
          if (p.jets.first_in[i] == d) v += af[i];
        g[lay.w1() + d * H + j] = v;
      }
      g[lay.b1() + j] = gb + ab;
    }
  }
};

// even channel counts run the packed program, odd ones the scalar one
template <int H, int NF, int NS, int ACT, int X = 0, bool EVEN = (NarrowWarp<H, NF, NS, ACT, X>::C % 2 == 0)>
struct NarrowSelect {
  using type = NarrowWarp<H, NF, NS, ACT, X>;
};
template <int H, int NF, int NS, int ACT, int X>
struct NarrowSelect<H, NF, NS, ACT, X, true> {
  using type = NarrowWarpPair<H, NF, NS, ACT, X>;
};
template <int H, int NF, int NS, int ACT, int X = 0>
using NarrowWarpFor = typename NarrowSelect<H, NF, NS, ACT, X>::type;

// Weight staging: global (row-major [out,in]) -> smem padded/transposed Layout.
// idx is an index into the Layout; returns the value to store there.
PINN_HD float stage_weight(const NarrowParams& p, int Hreal, const Layout& lay, int idx) {
  const int H = lay.H, L = lay.L;
  if (idx < lay.b1()) {  // W1t[d][j]
    const int d = idx / H, j = idx % H;
    return (d < p.n_in && j < Hreal) ? p.W[0][j * p.n_in + d] : 0.f;
  }
  if (idx < lay.wh(2)) {  // b1
    const int j = idx - lay.b1();
    return j < Hreal ? p.b[0][j] : 0.f;
  }
  if (idx < lay.wo()) {
    const int r = idx - lay.wh(2);
    const int l = 2 + r / (H * H + H);
    const int q = r % (H * H + H);
    if (q < H * H) {  // Wt_l[k][j] = W_l[j][k]
      const int k = q / H, j = q % H;
      return (k < Hreal && j < Hreal) ? p.W[l - 1][j * Hreal + k] : 0.f;
    }
    const int j = q - H * H;
    return j < Hreal ? p.b[l - 1][j] : 0.f;
  }
  if (idx < lay.bo()) {  // Woutt[k][o]
    const int q = idx - lay.wo();
    const int k = q / 4, o = q % 4;
    return (k < Hreal && o < p.n_out) ? p.W[L][o * Hreal + k] : 0.f;
  }
  const int o = idx - lay.bo();
  return o < p.n_out ? p.b[L][o] : 0.f;
}

// Inverse map for the gradient: flat torch index (layer order, W row-major
// [out,in] then b) -> index into the padded Layout.
PINN_HD int flat_to_layout(int n_in, int n_out, int Hreal, const Layout& lay, int64_t flat) {
  const int H = lay.H, L = lay.L;
  int64_t r = flat;
  // layer 0
  if (r < (int64_t)Hreal * n_in) {
    const int j = (int)(r / n_in), d = (int)(r % n_in);
    return lay.w1() + d * H + j;
  }
  r -= (int64_t)Hreal * n_in;
  if (r < Hreal) return lay.b1() + (int)r;
  r -= Hreal;
  for (int l = 2; l <= L; ++l) {
    if (r < (int64_t)Hreal * Hreal) {
      const int j = (int)(r / Hreal), k = (int)(r % Hreal);
      return lay.wh(l) + k * H + j;
    }
    r -= (int64_t)Hreal * Hreal;
    if (r < Hreal) return lay.bh(l) + (int)r;
    r -= Hreal;
  }
  if (r < (int64_t)n_out * Hreal) {
    const int o = (int)(r / Hreal), k = (int)(r % Hreal);
    return lay.wo() + k * 4 + o;
  }
  r -= (int64_t)n_out * Hreal;
  return lay.bo() + (int)r;
}

}  // namespace pinn
