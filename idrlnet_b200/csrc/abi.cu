// extern "C" entry points of libpinn_b200.so (see include/pinn_b200.h).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "narrow_kernel.cuh"
#include "select.cuh"
#include "geometry.cuh"
#include "wide.cuh"

namespace pinn {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}
void count_launch(int n) { g_launches += n; }

static int padded_width(int H) { return H <= 20 ? 20 : (H <= 32 ? 32 : 0); }

// extra-channel mode of the narrow kernels (NarrowWarp's X): 0 none, 1 mixed, 2 third order, 3 third + fourth; -1: no kernel
static int jets_extra(const pinn_jets* j) {
  if (j->n_mixed == 0 && j->n_high == 0) return 0;
  if (j->n_mixed == 1 && j->n_high == 0) return 1;
  if (j->n_mixed == 0 && (j->n_high == 1 || j->n_high == 2)) return 1 + j->n_high;
  return -1;
}

static const NarrowKernelInfo* narrow_lookup(const pinn_mlp* m, const pinn_jets* j) {
  const int Hp = padded_width(m->width);
  if (!Hp) return nullptr;
  if (Hp == 20) {
    if (m->act == PINN_ACT_TANH) return narrow_table_h20_tanh(j->n_first, j->n_second, jets_extra(j));
    if (m->act == PINN_ACT_SILU) return narrow_table_h20_silu(j->n_first, j->n_second, jets_extra(j));
    if (m->act == PINN_ACT_SIN) return narrow_table_h20_sin(j->n_first, j->n_second, jets_extra(j));
  } else {
    if (m->act == PINN_ACT_TANH) return narrow_table_h32_tanh(j->n_first, j->n_second, jets_extra(j));
    if (m->act == PINN_ACT_SILU) return narrow_table_h32_silu(j->n_first, j->n_second, jets_extra(j));
    if (m->act == PINN_ACT_SIN) return narrow_table_h32_sin(j->n_first, j->n_second, jets_extra(j));
  }
  return nullptr;
}

static int check_mlp(const pinn_mlp* m, const pinn_jets* j) {
  if (!m || !j) return fail(PINN_E_INVALID, "null descriptor");
  if (m->n_in < 1 || m->n_in > PINN_MAX_IN) return fail(PINN_E_INVALID, "n_in=%d out of range", m->n_in);
  if (m->n_out < 1 || m->n_out > PINN_MAX_OUT) return fail(PINN_E_INVALID, "n_out=%d out of range", m->n_out);
  if (m->n_hidden < 1 || m->n_hidden + 1 > PINN_MAX_LAYERS) return fail(PINN_E_INVALID, "n_hidden=%d out of range", m->n_hidden);
  if (m->width < 1) return fail(PINN_E_INVALID, "width=%d", m->width);
  if (m->act < 0 || m->act > PINN_ACT_SIN) return fail(PINN_E_INVALID, "act=%d", m->act);
  if (j->n_first < 0 || j->n_first > 3 || j->n_second < 0 || j->n_second > j->n_first)
    return fail(PINN_E_UNSUPPORTED, "jet plan (%d first, %d second) not supported", j->n_first, j->n_second);
  if (j->n_mixed < 0 || j->n_mixed > 1 || j->n_high < 0 || j->n_high > 2 || (j->n_mixed && j->n_first < 2) ||
      (j->n_high && j->n_second < 1) || (j->n_mixed && j->n_high))
    return fail(PINN_E_UNSUPPORTED, "jet plan (%d mixed, %d third/fourth-order channels) not supported", j->n_mixed, j->n_high);
  for (int i = 0; i < j->n_first; ++i)
    if (j->first_in[i] < 0 || j->first_in[i] >= m->n_in) return fail(PINN_E_INVALID, "first_in[%d]=%d", i, j->first_in[i]);
  for (int l = 0; l <= m->n_hidden; ++l)
    if (!m->W[l] || !m->b[l]) return fail(PINN_E_INVALID, "layer %d has null weights", l);
  return 0;
}

struct DeviceInfo {
  int sm_count = 0;
  int max_smem_optin = 0;
  int cc_major = 0;
  bool ok = false;
};
static DeviceInfo& device_info() {
  static thread_local int cached_dev = -1;
  static thread_local DeviceInfo info;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    info.ok = false;
    return info;
  }
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&info.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&info.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&info.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    info.ok = true;
    cached_dev = dev;
  }
  return info;
}

// ----------------------------------------------------------------------------
// workspace layout (floats): [gpart: MAX_CTA * PP][lpart: n_domains * MAX_CTA]
//                            [stash: MAX_CTA * 8 warps * stash_floats]
// ----------------------------------------------------------------------------
static const int kMaxCta = 160;  // >= SM count of a B200 (148); one persistent CTA per SM
static const int kMaxWarps = 10;

struct NarrowPlan {
  const NarrowKernelInfo* k;
  int Hp, L, PP, n_warps, smem_bytes, n_cta_cap;
  int64_t stash_floats_warp;
  int64_t head_bytes(int n_slots) const { return ((int64_t)kMaxCta * PP + (int64_t)n_slots * kMaxCta) * 4; }
  // L2-resident scratch at the END of the workspace (offsets independent of the slot count): per-warp jet stash,
  // preceded by the per-warp gradient accumulators
  int64_t stash_bytes() const { return (int64_t)kMaxCta * kMaxWarps * stash_floats_warp * 4; }
  int64_t gacc_bytes() const { return (int64_t)kMaxCta * kMaxWarps * PP * 4; }
  int64_t scratch_bytes() const { return stash_bytes() + gacc_bytes(); }
};

static int narrow_plan(const pinn_mlp* m, const pinn_jets* j, NarrowPlan* pl, bool need_device) {
  pl->k = narrow_lookup(m, j);
  if (!pl->k) return fail(PINN_E_UNSUPPORTED, "no narrow kernel for width %d", m->width);
  pl->Hp = pl->k->H;
  pl->L = m->n_hidden;
  const Layout lay(pl->Hp, pl->L);
  pl->PP = lay.size();
  const int wf = pl->k->warp_floats(pl->L);
  int max_smem = 232448;  // 227 KB
  int sms = 148;
  if (need_device) {
    DeviceInfo& di = device_info();
    if (!di.ok) return fail(PINN_E_NO_DEVICE, "no CUDA device");
    max_smem = di.max_smem_optin;
    sms = di.sm_count;
  }
  int nw = (max_smem - lay.size() * 4) / (wf * 4);
  if (nw > kMaxWarps) nw = kMaxWarps;
  if (nw > pl->k->max_warps) nw = pl->k->max_warps;
  {  // tuning knob for occupancy experiments (profiles/r01/narrow_warps_sweep.txt)
    static const int cap = getenv("PINN_NARROW_MAX_WARPS") ? atoi(getenv("PINN_NARROW_MAX_WARPS")) : 0;
    if (cap > 0 && nw > cap) nw = cap;
  }
  if (nw < 1) return fail(PINN_E_UNSUPPORTED, "net too deep for the narrow kernel (smem)");
  pl->n_warps = nw;
  pl->smem_bytes = (lay.size() + nw * wf) * 4;
  pl->n_cta_cap = sms < kMaxCta ? sms : kMaxCta;
  pl->stash_floats_warp = pl->k->stash_floats(pl->L);
  return 0;
}

// grid: ceil(n_params / 32) blocks of 256 threads = 32 outputs x 8 CTA groups; each thread sums
// its CTA group in fixed order, then the 8 partial sums are added in fixed order (deterministic).
__global__ void finalize_kernel(const float* gpart, const float* lpart, int max_cta, int PP, int n_in, int n_out,
                                int Hreal, int Hp, int L, int64_t n_params, float* flat_grad, int add, float* losses,
                                int n_dom) {
  __shared__ float red[8][33];
  const int il = threadIdx.x & 31, cg = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + il;
  float s = 0.f;
  if (i < n_params) {
    const Layout lay(Hp, L);
    const int idx = flat_to_layout(n_in, n_out, Hreal, lay, i);
    for (int c = cg; c < max_cta; c += 8) s += gpart[(size_t)c * PP + idx];
  }
  red[cg][il] = s;
  __syncthreads();
  if (cg == 0 && i < n_params) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][il];
    flat_grad[i] = add ? flat_grad[i] + t : t;
  }
  if (blockIdx.x == 0 && threadIdx.x < n_dom && losses) {
    float t = 0.f;
    for (int c = 0; c < max_cta; ++c) t += lpart[(size_t)threadIdx.x * max_cta + c];
    losses[threadIdx.x] = t;
  }
}

// finalize_kernel fused with a one-shot all-reduce over the peers' symmetric regions (include/pinn_b200.h).
// Region layout (floats): data[2 sets][world][n_total_pad], then flags (uint32) [2 sets][world][n_blocks].
struct PeerTable {
  float* base[PINN_MAX_RANKS];
  int rank, world;
  uint32_t epoch;
  int n_total_pad, n_blocks;
};
__host__ __device__ inline int64_t peer_flags_off(int world, int n_total_pad) { return (int64_t)2 * world * n_total_pad; }

__global__ void finalize_allreduce_kernel(const float* gpart, const float* lpart, int max_cta, int PP, int n_in, int n_out,
                                          int Hreal, int Hp, int L, int64_t n_params, int n_dom, PeerTable pt, float* out) {
  __shared__ float red[8][33];
  const int il = threadIdx.x & 31, cg = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + il;
  float s = 0.f;
  if (i < n_params) {
    const Layout lay(Hp, L);
    const int idx = flat_to_layout(n_in, n_out, Hreal, lay, i);
    for (int c = cg; c < max_cta; c += 8) s += gpart[(size_t)c * PP + idx];
  }
  red[cg][il] = s;
  __syncthreads();
  const int set = (int)(pt.epoch & 1u);
  const size_t slot_mine = ((size_t)set * pt.world + pt.rank) * pt.n_total_pad;
  if (cg == 0 && i < n_params) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][il];
    for (int r = 0; r < pt.world; ++r) pt.base[r][slot_mine + i] = t;  // 128-byte row segments into every peer (and self)
  }
  if (blockIdx.x == 0 && cg == 1 && il < n_dom) {  // (a second warp: the losses ride in the same buffer, after the gradient)
    float t = 0.f;
    for (int c = 0; c < max_cta; ++c) t += lpart[(size_t)il * max_cta + c];
    for (int r = 0; r < pt.world; ++r) pt.base[r][slot_mine + n_params + il] = t;
  }
  __threadfence_system();
  __syncthreads();
  const int64_t foff = peer_flags_off(pt.world, pt.n_total_pad);
  if (threadIdx.x < pt.world) {
    const int r = threadIdx.x;
    // publish: "block b of rank `rank` has written its values into your region"
    uint32_t* theirs = reinterpret_cast<uint32_t*>(pt.base[r] + foff) + ((size_t)set * pt.world + pt.rank) * pt.n_blocks + blockIdx.x;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(pt.epoch) : "memory");
    // wait for the same block of rank r
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(pt.base[pt.rank] + foff) + ((size_t)set * pt.world + r) * pt.n_blocks + blockIdx.x;
    const long long t0 = clock64();
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (v != pt.epoch) {
        if (clock64() - t0 > 120000000000ll) __trap();  // ~60 s: a rank is missing; fail instead of hanging
        if (clock64() - t0 > 200000ll) __nanosleep(200);  // a late peer (host-side skew): stop hammering the link
      }
    } while (v != pt.epoch);
  }
  __syncthreads();
  const float* local = pt.base[pt.rank] + (size_t)set * pt.world * pt.n_total_pad;
  if (cg == 0 && i < n_params) {
    float t = 0.f;
    for (int r = 0; r < pt.world; ++r) t += __ldcg(local + (size_t)r * pt.n_total_pad + i);  // rank order: the same sum everywhere
    out[i] = t;
  }
  if (blockIdx.x == 0 && cg == 1 && il < n_dom) {
    float t = 0.f;
    for (int r = 0; r < pt.world; ++r) t += __ldcg(local + (size_t)r * pt.n_total_pad + n_params + il);
    out[n_params + il] = t;
  }
}

static int launch_narrow(const pinn_mlp* m, const pinn_jets* j, NarrowParams& p, const NarrowPlan& pl, cudaStream_t st) {
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lk(mu);
    cudaError_t e = cudaFuncSetAttribute((const void*)pl.k->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(narrow)");
  }
  p.n_warps = pl.n_warps;
  const int64_t tiles = (p.n_points + 31) / 32;
  p.n_tiles = (int32_t)tiles;
  int64_t ctas = tiles;  // one tile per CTA before a second warp of any CTA gets one (narrow_kernel's tile map)
  if (ctas > pl.n_cta_cap) ctas = pl.n_cta_cap;
  if (ctas < 1) ctas = 1;
  narrow_kernel_t fn = pl.k->fn;
  // (A per-launch persisting-L2 access-policy window over the per-warp scratch was measured and NOT kept: DRAM writes of
  // the 2^20-point Burgers launch went from 210 MB to 1 364 MB -- profiles/r02/narrow_l2_persist_experiment.txt.)
  fn<<<(unsigned)ctas, pl.n_warps * 32, pl.smem_bytes, st>>>(p, m->width);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "narrow kernel launch");
  count_launch(1);
  return 0;
}

static void fill_net(NarrowParams& p, const pinn_mlp* m, const pinn_jets* j) {
  memset(&p, 0, sizeof(p));
  p.n_in = m->n_in;
  p.n_out = m->n_out;
  p.n_hidden = m->n_hidden;
  for (int l = 0; l <= m->n_hidden; ++l) {
    p.W[l] = m->W[l];
    p.b[l] = m->b[l];
  }
  p.jets = *j;
  p.max_cta = kMaxCta;
}

}  // namespace pinn

using namespace pinn;

extern "C" {

int pinn_abi_version(void) { return PINN_ABI_VERSION; }
const char* pinn_last_error(void) { return g_err; }
int64_t pinn_launch_count(void) { return g_launches.load(); }

int64_t pinn_param_count(const pinn_mlp* m) {
  if (!m) return 0;
  int64_t n = (int64_t)m->width * m->n_in + m->width;
  n += (int64_t)(m->n_hidden - 1) * ((int64_t)m->width * m->width + m->width);
  n += (int64_t)m->n_out * m->width + m->n_out;
  return n;
}

int pinn_supported(const pinn_mlp* m, const pinn_jets* j) {
  if (!m || !j) return 0;
  if (narrow_lookup(m, j)) {
    NarrowPlan pl;
    return narrow_plan(m, j, &pl, false) == 0 ? 1 : 0;
  }
  return wide_supported(m, j) ? 1 : 0;
}

int64_t pinn_workspace_bytes(const pinn_mlp* m, const pinn_jets* j, int32_t n_domains) {
  if (!m || !j || n_domains < 1) return 0;
  if (narrow_lookup(m, j)) {
    NarrowPlan pl;
    if (narrow_plan(m, j, &pl, false)) return 0;
    return pl.head_bytes(n_domains) + pl.scratch_bytes();
  }
  return wide_workspace_bytes(m, j, n_domains);
}

static int narrow_step_common(const pinn_mlp* m, const pinn_jets* j, NarrowParams& p, void* ws, int64_t ws_bytes,
                              int n_slots_needed, int accumulate, cudaStream_t st) {
  NarrowPlan pl;
  int rc = narrow_plan(m, j, &pl, true);
  if (rc) return rc;
  const int64_t gbytes = (int64_t)kMaxCta * pl.PP * 4;
  // the stash sits at the END of the workspace so that its offset does not depend on the slot count
  if (!ws || ws_bytes < pl.head_bytes(n_slots_needed) + pl.scratch_bytes())
    return fail(PINN_E_WORKSPACE, "workspace too small: %lld bytes", (long long)ws_bytes);
  p.gpart = (float*)ws;
  p.lpart = (float*)((char*)ws + gbytes);
  p.stash = (float*)((char*)ws + (ws_bytes - pl.stash_bytes()));
  p.gacc = (float*)((char*)ws + (ws_bytes - pl.scratch_bytes()));
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)(ws_bytes - pl.scratch_bytes()), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(workspace)");
  }
  return launch_narrow(m, j, p, pl, st);
}

static int step_or_loss(const pinn_mlp* m, const pinn_jets* j, const pinn_domain* dom, void* ws, int64_t ws_bytes,
                        int32_t domain_slot, int32_t accumulate, bool loss_only, void* stream) {
  int rc = check_mlp(m, j);
  if (rc) return rc;
  if (!dom) return fail(PINN_E_INVALID, "null domain");
  if (dom->n_points < 0 || dom->n_eq < 0 || dom->n_eq > PINN_MAX_EQ || dom->n_terms < 0 ||
      dom->n_terms > PINN_MAX_TERMS || dom->n_aux < 0 || dom->n_aux > PINN_MAX_AUX || domain_slot < 0)
    return fail(PINN_E_INVALID, "malformed domain descriptor");
  for (int d = 0; d < m->n_in; ++d)
    if (!dom->coords[d] && dom->n_points > 0) return fail(PINN_E_INVALID, "null coordinate column %d", d);
  for (int e = 0; e < dom->n_eq; ++e) {
    const pinn_equation& q = dom->eq[e];
    if (q.term_begin < 0 || q.term_end < q.term_begin || q.term_end > dom->n_terms)
      return fail(PINN_E_INVALID, "equation %d term range", e);
    for (int t = q.term_begin; t < q.term_end; ++t) {
      if (dom->terms[t].n_fac > PINN_MAX_FACTORS) return fail(PINN_E_INVALID, "term %d has too many factors", t);
      for (int f = 0; f < dom->terms[t].n_fac; ++f)
        if (dom->terms[t].fac[f] >= PINN_SYM_COUNT) return fail(PINN_E_INVALID, "term %d symbol id", t);
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (narrow_lookup(m, j)) {
    NarrowParams p;
    fill_net(p, m, j);
    p.mode = loss_only ? MODE_LOSS : MODE_FUSED;
    p.n_points = dom->n_points;
    for (int d = 0; d < m->n_in; ++d) p.coords[d] = dom->coords[d];
    p.dom = *dom;
    p.domain_slot = domain_slot;
    return narrow_step_common(m, j, p, ws, ws_bytes, domain_slot + 1, accumulate, st);
  }
  if (wide_supported(m, j)) return wide_step_fused(m, j, dom, ws, ws_bytes, domain_slot + 1, domain_slot, accumulate, loss_only, st);
  return fail(PINN_E_UNSUPPORTED, "no fused kernel for width %d / act %d", m->width, m->act);
}

int pinn_step_fused(const pinn_mlp* m, const pinn_jets* j, const pinn_domain* dom, void* ws, int64_t ws_bytes,
                    int32_t domain_slot, int32_t accumulate, void* stream) {
  return step_or_loss(m, j, dom, ws, ws_bytes, domain_slot, accumulate, false, stream);
}

int pinn_loss_fused(const pinn_mlp* m, const pinn_jets* j, const pinn_domain* dom, void* ws, int64_t ws_bytes,
                    int32_t domain_slot, int32_t accumulate, void* stream) {
  return step_or_loss(m, j, dom, ws, ws_bytes, domain_slot, accumulate, true, stream);
}

int pinn_step_finalize(const pinn_mlp* m, const pinn_jets* j, void* ws, int64_t ws_bytes, int32_t n_domains,
                       float* flat_grad, int32_t add_to_grad, float* losses, void* stream) {
  int rc = check_mlp(m, j);
  if (rc) return rc;
  if (!flat_grad) return fail(PINN_E_INVALID, "null flat_grad");
  cudaStream_t st = (cudaStream_t)stream;
  if (narrow_lookup(m, j)) {
    NarrowPlan pl;
    rc = narrow_plan(m, j, &pl, true);
    if (rc) return rc;
    const int64_t gbytes = (int64_t)kMaxCta * pl.PP * 4;
    if (!ws || ws_bytes < pl.head_bytes(n_domains)) return fail(PINN_E_WORKSPACE, "workspace too small");
    if (n_domains > 256) return fail(PINN_E_INVALID, "too many domains");
    const int64_t np = pinn_param_count(m);
    const int threads = 256;
    const int blocks = (int)((np + 31) / 32);
    finalize_kernel<<<blocks, threads, 0, st>>>((const float*)ws, (const float*)((char*)ws + gbytes), kMaxCta, pl.PP,
                                                m->n_in, m->n_out, m->width, pl.Hp, pl.L, np, flat_grad, add_to_grad,
                                                losses, n_domains);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "finalize launch");
    count_launch(1);
    return 0;
  }
  if (wide_supported(m, j)) return wide_step_finalize(m, j, ws, ws_bytes, n_domains, flat_grad, add_to_grad, losses, st);
  return fail(PINN_E_UNSUPPORTED, "no fused kernel for width %d", m->width);
}

static int peer_geometry(const pinn_mlp* m, int n_domains, int* n_total_pad, int* n_blocks) {
  const int64_t np = pinn_param_count(m);
  if (np + n_domains > (1 << 20)) return 0;  // small gradients only: the latency-bound case this exists for
  *n_total_pad = (int)((np + n_domains + 31) / 32 * 32);
  *n_blocks = (int)((np + 31) / 32);
  return 1;
}

int64_t pinn_allreduce_region_bytes(const pinn_mlp* m, const pinn_jets* j, int32_t n_domains, int32_t world) {
  if (!m || !j || world < 2 || world > PINN_MAX_RANKS || n_domains < 1 || n_domains > 32) return 0;
  if (!narrow_lookup(m, j)) return 0;  // wide nets: the all-reduce is 1e-3 of their step; they keep the library collective
  int ntp, nb;
  if (!peer_geometry(m, n_domains, &ntp, &nb)) return 0;
  return (peer_flags_off(world, ntp) + (int64_t)2 * world * nb) * 4;
}

int pinn_step_finalize_allreduce(const pinn_mlp* m, const pinn_jets* j, void* ws, int64_t ws_bytes, int32_t n_domains,
                                 float* out, const pinn_peers* peers, void* stream) {
  int rc = check_mlp(m, j);
  if (rc) return rc;
  if (!out || !peers) return fail(PINN_E_INVALID, "null out / peers");
  if (peers->world < 2 || peers->world > PINN_MAX_RANKS || peers->rank < 0 || peers->rank >= peers->world || peers->epoch == 0)
    return fail(PINN_E_INVALID, "peers: rank %d of %d, epoch %u", peers->rank, peers->world, peers->epoch);
  if (n_domains < 1 || n_domains > 32) return fail(PINN_E_INVALID, "n_domains=%d", n_domains);
  if (!narrow_lookup(m, j)) return fail(PINN_E_UNSUPPORTED, "finalize + all-reduce exists for narrow nets only");
  NarrowPlan pl;
  rc = narrow_plan(m, j, &pl, true);
  if (rc) return rc;
  PeerTable pt;
  memset(&pt, 0, sizeof(pt));
  if (!peer_geometry(m, n_domains, &pt.n_total_pad, &pt.n_blocks)) return fail(PINN_E_UNSUPPORTED, "gradient too large");
  for (int r = 0; r < peers->world; ++r) {
    if (!peers->base[r]) return fail(PINN_E_INVALID, "peers->base[%d] is null", r);
    pt.base[r] = (float*)peers->base[r];
  }
  pt.rank = peers->rank; pt.world = peers->world; pt.epoch = peers->epoch;
  const int64_t gbytes = (int64_t)kMaxCta * pl.PP * 4;
  if (!ws || ws_bytes < pl.head_bytes(n_domains)) return fail(PINN_E_WORKSPACE, "workspace too small");
  const int64_t np = pinn_param_count(m);
  finalize_allreduce_kernel<<<pt.n_blocks, 256, 0, (cudaStream_t)stream>>>(
      (const float*)ws, (const float*)((char*)ws + gbytes), kMaxCta, pl.PP, m->n_in, m->n_out, m->width, pl.Hp, pl.L, np,
      n_domains, pt, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "finalize + all-reduce launch");
  count_launch(1);
  return 0;
}

int pinn_jet_forward(const pinn_mlp* m, const pinn_jets* j, int64_t n_points, const float* const* coords, float* out,
                     void* ws, int64_t ws_bytes, void* stream) {
  int rc = check_mlp(m, j);
  if (rc) return rc;
  if (!coords || !out || n_points < 0) return fail(PINN_E_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_points == 0) return 0;
  if (narrow_lookup(m, j)) {
    NarrowPlan pl;
    rc = narrow_plan(m, j, &pl, true);
    if (rc) return rc;
    NarrowParams p;
    fill_net(p, m, j);
    p.mode = MODE_FWD;
    p.n_points = n_points;
    for (int d = 0; d < m->n_in; ++d) p.coords[d] = coords[d];
    p.jets_out = out;
    return launch_narrow(m, j, p, pl, st);
  }
  if (wide_supported(m, j)) return wide_jet_forward(m, j, n_points, coords, out, ws, ws_bytes, st);
  return fail(PINN_E_UNSUPPORTED, "no kernel for width %d", m->width);
}

int pinn_jet_backward(const pinn_mlp* m, const pinn_jets* j, int64_t n_points, const float* const* coords,
                      const float* out_bar, void* ws, int64_t ws_bytes, int32_t accumulate, void* stream) {
  int rc = check_mlp(m, j);
  if (rc) return rc;
  if (!coords || !out_bar || n_points < 0) return fail(PINN_E_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (narrow_lookup(m, j)) {
    NarrowParams p;
    fill_net(p, m, j);
    p.mode = MODE_BWD;
    p.n_points = n_points;
    for (int d = 0; d < m->n_in; ++d) p.coords[d] = coords[d];
    p.jets_bar = out_bar;
    return narrow_step_common(m, j, p, ws, ws_bytes, 1, accumulate, st);
  }
  if (wide_supported(m, j)) return wide_jet_backward(m, j, n_points, coords, out_bar, ws, ws_bytes, accumulate, st);
  return fail(PINN_E_UNSUPPORTED, "no kernel for width %d", m->width);
}

// ---- forward-only residuals and top-k (adaptive re-sampling) -------------------------------------
int64_t pinn_residual_workspace_bytes(const pinn_mlp* m, const pinn_jets* j, int64_t n_points) {
  if (!m || !j || n_points < 0) return PINN_E_INVALID;
  const int C = PINN_JET_CHANNELS(j);
  const int64_t jets = (((int64_t)C * m->n_out * n_points * (int64_t)sizeof(float)) + 255) & ~(int64_t)255;
  const int64_t fwd = pinn_workspace_bytes(m, j, 1);
  return fwd < 0 ? fwd : jets + fwd;
}

int pinn_residual_forward(const pinn_mlp* m, const pinn_jets* j, const pinn_domain* dom, float* out, void* ws,
                          int64_t ws_bytes, void* stream) {
  int rc = check_mlp(m, j);
  if (rc) return rc;
  if (!dom || !out || !ws) return fail(PINN_E_INVALID, "null argument");
  if (dom->n_eq < 1 || dom->n_eq > PINN_MAX_EQ || dom->n_terms > PINN_MAX_TERMS)
    return fail(PINN_E_INVALID, "bad equation table");
  if (dom->n_points == 0) return 0;
  const int C = PINN_JET_CHANNELS(j);
  const int64_t jets_bytes = (((int64_t)C * m->n_out * dom->n_points * (int64_t)sizeof(float)) + 255) & ~(int64_t)255;
  if (ws_bytes < pinn_residual_workspace_bytes(m, j, dom->n_points)) return fail(PINN_E_WORKSPACE, "workspace too small");
  float* jets = (float*)ws;
  rc = pinn_jet_forward(m, j, dom->n_points, dom->coords, jets, (char*)ws + jets_bytes, ws_bytes - jets_bytes, stream);
  if (rc) return rc;
  rc = residual_eval(dom, m->n_out, jets, out, (cudaStream_t)stream);
  return rc > 0 ? cuda_fail((cudaError_t)rc, "residual kernel launch") : rc;
}

int64_t pinn_topk_workspace_bytes(int32_t k) { return k < 1 || k > kTopkMax ? PINN_E_UNSUPPORTED : topk_workspace_bytes(k); }

int pinn_topk_abs(const float* values, int64_t n, int32_t k, int64_t* idx_out, float* val_out, void* ws, int64_t ws_bytes,
                  void* stream) {
  if (!values || !idx_out || !ws || n < 0) return fail(PINN_E_INVALID, "null argument");
  if (k < 1 || k > kTopkMax) return fail(PINN_E_UNSUPPORTED, "k = %d outside 1..%d", k, kTopkMax);
  if (k > n) return fail(PINN_E_INVALID, "k = %d exceeds the %lld values", k, (long long)n);
  if (n >= ((int64_t)1 << 32)) return fail(PINN_E_UNSUPPORTED, "more than 2^32 - 1 values");
  if (ws_bytes < topk_workspace_bytes(k)) return fail(PINN_E_WORKSPACE, "workspace too small");
  const int rc = topk_abs(values, n, k, idx_out, val_out, ws, (cudaStream_t)stream);
  return rc > 0 ? cuda_fail((cudaError_t)rc, "top-k launch") : rc;
}

// ---- CSG sampling --------------------------------------------------------------------------
int64_t pinn_sample_csg_workspace_bytes(int64_t n) { return n < 0 ? PINN_E_INVALID : csg_workspace_bytes(n); }

int pinn_sample_csg(float* const* columns, int32_t n_dims, const float* lo, const float* hi, int64_t n, int64_t capacity,
                    uint64_t seed, uint64_t stream_id, uint64_t offset, const pinn_shape_op* prog, int32_t n_ops, float* sdf,
                    int64_t* count, void* ws, int64_t ws_bytes, void* stream) {
  if (!columns || !lo || !hi || !prog || !count || !ws || n_dims < 1 || n_dims > PINN_MAX_IN || n < 0 || capacity < 0)
    return fail(PINN_E_INVALID, "null or out-of-range argument");
  if (n_ops < 1 || n_ops > PINN_MAX_SHAPE_OPS) return fail(PINN_E_UNSUPPORTED, "program of %d ops (max %d)", n_ops, PINN_MAX_SHAPE_OPS);
  if (n >= ((int64_t)1 << 31) * 256) return fail(PINN_E_UNSUPPORTED, "too many candidates");
  if (ws_bytes < csg_workspace_bytes(n)) return fail(PINN_E_WORKSPACE, "workspace too small");
  int depth = 0;  // validate the stack program on the host: never below 1 operand, ends with exactly one, depth <= 8
  for (int i = 0; i < n_ops; ++i) {
    const int op = prog[i].op;
    if (op >= PINN_SHAPE_INTERVAL && op <= PINN_SHAPE_SPHERE) {
      const int need = (op == PINN_SHAPE_INTERVAL) ? (int)prog[i].p[2] + 1 : (op == PINN_SHAPE_BOX || op == PINN_SHAPE_SPHERE) ? 3 : 2;
      if (need > n_dims || need < 1) return fail(PINN_E_INVALID, "op %d: primitive needs %d coordinates, have %d", i, need, n_dims);
      ++depth;
    } else if (op == PINN_CSG_UNION || op == PINN_CSG_DIFFERENCE || op == PINN_CSG_INTERSECTION) {
      if (depth < 2) return fail(PINN_E_INVALID, "op %d: binary operator on %d operands", i, depth);
      --depth;
    } else if (op == PINN_CSG_COMPLEMENT) {
      if (depth < 1) return fail(PINN_E_INVALID, "op %d: complement of nothing", i);
    } else {
      return fail(PINN_E_INVALID, "op %d: unknown opcode %d", i, op);
    }
    if (depth > 8) return fail(PINN_E_UNSUPPORTED, "CSG stack deeper than 8");
  }
  if (depth != 1) return fail(PINN_E_INVALID, "program leaves %d values on the stack", depth);
  for (int d = 0; d < n_dims; ++d)
    if (!columns[d]) return fail(PINN_E_INVALID, "null column");
  const int rc = csg_sample(columns, n_dims, lo, hi, n, capacity, seed, stream_id, offset, prog, n_ops, sdf, count, ws,
                            (cudaStream_t)stream);
  return rc > 0 ? cuda_fail((cudaError_t)rc, "csg sampling launch") : rc;
}

}  // extern "C"
