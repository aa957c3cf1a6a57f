// CPU emulation of the narrow warp programs (idrlnet_b200/csrc/narrow_warp.h).
// TEST INFRASTRUCTURE ONLY: lets the CPU-only test tier execute the very same
// per-lane phase functions the CUDA kernel runs (lane by lane, phase by phase)
// and compare them with the oracle.  The product library never links this.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../idrlnet_b200/csrc/narrow_warp.h"

using namespace pinn;

namespace {

struct Runner {
  virtual ~Runner() {}
  virtual int H() const = 0;
  virtual int warp_floats(int L) const = 0;
  virtual void run_cta(const NarrowParams& p, int Hreal, int cta, int n_cta, int n_warps, float* gpart_cta,
                       float* lpart_cta) const = 0;
};

template <class NW>
struct RunnerT : Runner {
  int H() const override { return NW::H; }
  int warp_floats(int L) const override { return NW::warp_floats(L); }
  void run_cta(const NarrowParams& p, int Hreal, int cta, int n_cta, int n_warps, float* gpart_cta,
               float* lpart_cta) const override {
    const int L = p.n_hidden;
    const Layout lay(NW::H, L);
    const int wf = NW::warp_floats(L);
    std::vector<float> smem(lay.size() + (size_t)n_warps * wf, 0.f);
    for (int i = 0; i < lay.size(); ++i) smem[i] = stage_weight(p, Hreal, lay, i);
    std::vector<typename NW::Lane> lanes((size_t)n_warps * 32);
    for (auto& l : lanes) l.loss = 0.f;
    std::vector<float> stash((size_t)n_warps * NW::stash_floats(L) + 4, 0.f);
    std::vector<float> gacc((size_t)n_warps * lay.size(), 0.f);
    for (int warp = 0; warp < n_warps; ++warp) {
      float* wm = smem.data() + lay.size() + (size_t)warp * wf;
      float* gs = stash.data() + (size_t)warp * NW::stash_floats(L);
      float* g = gacc.data() + (size_t)warp * lay.size();
      typename NW::Lane* ln = &lanes[(size_t)warp * 32];
      for (int64_t tile = (int64_t)warp * n_cta + cta; tile < p.n_tiles; tile += (int64_t)n_cta * n_warps) {  // narrow_kernel's map
        for (int lane = 0; lane < 32; ++lane) NW::phase_forward(p, smem.data(), wm, gs, lane, tile, ln[lane]);
        if (p.mode == MODE_FWD || p.mode == MODE_LOSS) continue;
        for (int l = L + 1; l >= 2; --l) {
          for (int lane = 0; lane < 32; ++lane) NW::phase_backward(p, smem.data(), wm, gs, lane, l, ln[lane]);
          for (int lane = 0; lane < 32; ++lane) {
            if (l == L + 1) NW::phase_gemm_out(p, wm, g, lane);
            else NW::phase_gemm_hidden(p, wm, g, lane, l);
          }
          for (int lane = 0; lane < 32; ++lane) NW::phase_store_zk(wm, lane, ln[lane]);
        }
        for (int lane = 0; lane < 32; ++lane) NW::phase_gemm_first(p, wm, g, lane);
      }
    }
    if (p.mode == MODE_FWD) return;
    for (int i = 0; i < lay.size() && p.mode != MODE_LOSS; ++i) {
      float s = 0.f;
      for (int w = 0; w < n_warps; ++w) s += gacc[(size_t)w * lay.size() + i];
      gpart_cta[i] += s;
    }
    if (p.mode == MODE_FUSED || p.mode == MODE_LOSS) {
      float s = 0.f;
      for (int w = 0; w < n_warps; ++w) {
        float sw = 0.f;
        for (int l2 = 0; l2 < 32; ++l2) sw += lanes[(size_t)w * 32 + l2].loss;
        s += sw;
      }
      *lpart_cta += s;
    }
  }
};

#define ROW(H, A)                                                                                         \
  {{new RunnerT<NarrowWarpFor<H, 0, 0, A>>, nullptr, nullptr, nullptr},                                      \
   {new RunnerT<NarrowWarpFor<H, 1, 0, A>>, new RunnerT<NarrowWarpFor<H, 1, 1, A>>, nullptr, nullptr},          \
   {new RunnerT<NarrowWarpFor<H, 2, 0, A>>, new RunnerT<NarrowWarpFor<H, 2, 1, A>>, new RunnerT<NarrowWarpFor<H, 2, 2, A>>, nullptr}, \
   {new RunnerT<NarrowWarpFor<H, 3, 0, A>>, new RunnerT<NarrowWarpFor<H, 3, 1, A>>, new RunnerT<NarrowWarpFor<H, 3, 2, A>>, new RunnerT<NarrowWarpFor<H, 3, 3, A>>}}
// extra-channel programs: [0..2] mixed (2, 0) / (2, 1) / (2, 2), [3] third order (1, 1), [4] third + fourth order (1, 1)
#define XROW(H, A)                                                                                        \
  {new RunnerT<NarrowWarpFor<H, 2, 0, A, 1>>, new RunnerT<NarrowWarpFor<H, 2, 1, A, 1>>,                    \
   new RunnerT<NarrowWarpFor<H, 2, 2, A, 1>>, new RunnerT<NarrowWarpFor<H, 1, 1, A, 2>>,                    \
   new RunnerT<NarrowWarpFor<H, 1, 1, A, 3>>}

const Runner* lookup(int width, int act, const pinn_jets* j) {
  static const Runner* t20[3][4][4] = {ROW(20, PINN_ACT_TANH), ROW(20, PINN_ACT_SILU), ROW(20, PINN_ACT_SIN)};
  static const Runner* t32[3][4][4] = {ROW(32, PINN_ACT_TANH), ROW(32, PINN_ACT_SILU), ROW(32, PINN_ACT_SIN)};
  static const Runner* x20[3][5] = {XROW(20, PINN_ACT_TANH), XROW(20, PINN_ACT_SILU), XROW(20, PINN_ACT_SIN)};
  static const Runner* x32[3][5] = {XROW(32, PINN_ACT_TANH), XROW(32, PINN_ACT_SILU), XROW(32, PINN_ACT_SIN)};
  const int nf = j->n_first, ns = j->n_second;
  if (act < 0 || act > 2 || nf < 0 || nf > 3 || ns < 0 || ns > nf || width > 32) return nullptr;
  if (j->n_mixed || j->n_high) {
    int k = -1;
    if (j->n_mixed == 1 && !j->n_high && nf == 2) k = ns;
    if (!j->n_mixed && (j->n_high == 1 || j->n_high == 2) && nf == 1 && ns == 1) k = 2 + j->n_high;
    if (k < 0) return nullptr;
    return width <= 20 ? x20[act][k] : x32[act][k];
  }
  if (width <= 20) return t20[act][nf][ns];
  return t32[act][nf][ns];
}

int64_t param_count(const pinn_mlp* m) {
  int64_t n = (int64_t)m->width * m->n_in + m->width;
  n += (int64_t)(m->n_hidden - 1) * ((int64_t)m->width * m->width + m->width);
  n += (int64_t)m->n_out * m->width + m->n_out;
  return n;
}

void fill(NarrowParams& p, const pinn_mlp* m, const pinn_jets* j) {
  memset(&p, 0, sizeof(p));
  p.n_in = m->n_in; p.n_out = m->n_out; p.n_hidden = m->n_hidden;
  for (int l = 0; l <= m->n_hidden; ++l) { p.W[l] = m->W[l]; p.b[l] = m->b[l]; }
  p.jets = *j;
}

}  // namespace

extern "C" {

// mode 0: fused (dom, flat_grad += , loss_out +=); mode 1: jets_out; mode 2: jets_bar -> flat_grad +=;
// mode 3: loss only (dom, loss_out +=)
__attribute__((visibility("default"))) int pinn_emu_run(const pinn_mlp* m, const pinn_jets* j, int mode,
                                                        const pinn_domain* dom, int64_t n_points,
                                                        const float* const* coords, float* jets_out,
                                                        const float* jets_bar, float* flat_grad, float* loss_out,
                                                        int n_cta, int n_warps) {
  const Runner* r = lookup(m->width, m->act, j);
  if (!r) return PINN_E_UNSUPPORTED;
  NarrowParams p;
  fill(p, m, j);
  p.mode = mode;
  if (mode == MODE_FUSED || mode == MODE_LOSS) {
    p.dom = *dom;
    n_points = dom->n_points;
    for (int d = 0; d < m->n_in; ++d) p.coords[d] = dom->coords[d];
  } else {
    for (int d = 0; d < m->n_in; ++d) p.coords[d] = coords[d];
  }
  p.n_points = n_points;
  p.jets_out = jets_out;
  p.jets_bar = jets_bar;
  p.n_tiles = (int32_t)((n_points + 31) / 32);
  p.n_warps = n_warps;
  const Layout lay(r->H(), m->n_hidden);
  std::vector<float> gpart((size_t)n_cta * lay.size(), 0.f), lpart(n_cta, 0.f);
  for (int c = 0; c < n_cta; ++c) r->run_cta(p, m->width, c, n_cta, n_warps, &gpart[(size_t)c * lay.size()], &lpart[c]);
  if (mode == MODE_FWD) return 0;
  const int64_t np = param_count(m);
  for (int64_t i = 0; i < np; ++i) {
    const int idx = flat_to_layout(m->n_in, m->n_out, m->width, lay, i);
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c) s += gpart[(size_t)c * lay.size() + idx];
    flat_grad[i] += s;
  }
  if ((mode == MODE_FUSED || mode == MODE_LOSS) && loss_out) {
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c) s += lpart[c];
    *loss_out += s;
  }
  return 0;
}
}

// Host evaluation of the sampler maps (idrlnet_b200/csrc/philox.h) so the device
// sampler can be checked bit for bit.
#include "../../idrlnet_b200/csrc/philox.h"
extern "C" __attribute__((visibility("default"))) void pinn_emu_sample_box(float* const* columns, int n_dims,
                                                                          const float* lo, const float* hi,
                                                                          int64_t n_points, uint64_t seed,
                                                                          uint64_t stream_id, uint64_t offset,
                                                                          float* sdf, int sdf_dims) {
  for (int64_t i = 0; i < n_points; ++i) {
    uint32_t r[4];
    pinn::philox4x32_10(seed, stream_id, offset + (uint64_t)i, r);
    float xs[PINN_MAX_IN];
    for (int d = 0; d < n_dims; ++d) {
      xs[d] = pinn::box_point(r[d], lo[d], hi[d]);
      columns[d][i] = xs[d];
    }
    if (sdf) sdf[i] = pinn::box_sdf(xs, lo, hi, sdf_dims);
  }
}
