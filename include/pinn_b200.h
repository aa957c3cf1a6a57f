/*
 * pinn_b200.h -- C ABI of libpinn_b200.so: the B200 (sm_100a) implementation of the
 * IDRLnet PINN training step.
 *
 * The reference (idrl-lab/idrlnet v2.0.0) is pure Python and has no FFI; the seam
 * this library plugs into is its node protocol (idrlnet/node.py:11-66,
 * idrlnet/graph.py:95-121,170-180).  Each entry point below names the reference
 * code whose work it replaces.  All pointers are DEVICE pointers to FP32 unless
 * stated otherwise; every buffer is owned by the caller (PyTorch tensors in the
 * Python host layer); the library allocates nothing persistent, never
 * synchronises the host, and is stream-ordered on the `stream` argument
 * (a cudaStream_t passed as void*).
 *
 * Return value of every function: 0 = ok, <0 = PINN_E_* (invalid / unsupported
 * configuration -- the caller raises or takes the generic torch path),
 * >0 = a cudaError_t.  Nothing throws across the ABI, nothing calls exit().
 */
#ifndef PINN_B200_H
#define PINN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PINN_ABI_VERSION 2

#if defined(__GNUC__)
#define PINN_API __attribute__((visibility("default")))
#else
#define PINN_API
#endif

#define PINN_MAX_LAYERS 16 /* linear layers: hidden + output */
#define PINN_MAX_IN 4      /* network inputs (x, y, z, t / parameters) */
#define PINN_MAX_OUT 4     /* network outputs */
#define PINN_MAX_FIRST 4   /* first-order jet channels */
#define PINN_MAX_AUX 8     /* extra per-point symbols (normal_x, sdf, ...) */
#define PINN_MAX_EQ 8      /* equations (loss terms) per sampling domain */
#define PINN_MAX_TERMS 64  /* monomials over all equations of a domain */
#define PINN_MAX_FACTORS 7 /* factors per monomial */

/* symbol ids used by the residual monomials */
#define PINN_SYM_JET(c, o) ((c) * PINN_MAX_OUT + (o)) /* channel c of output o  */
#define PINN_SYM_COORD0 28                            /* + input index          */
#define PINN_SYM_AUX0 32                              /* + aux index            */
#define PINN_SYM_COUNT 40

enum { PINN_ACT_TANH = 0, PINN_ACT_SILU = 1, PINN_ACT_SIN = 2 };
enum { PINN_LOSS_SQUARE = 0, PINN_LOSS_L1 = 1, PINN_LOSS_IDENTITY = 2 };
/* PINN_PREC_FP32: FP32 FMA everywhere (1e-5 parity with the reference's FP32 autograd).
 * PINN_PREC_TF32: layer GEMMs of wide nets on tcgen05 tensor cores, multiplicands rounded
 * to TF32, FP32 accumulate; everything else FP32.  Narrow nets ignore the field. */
enum {
  PINN_PREC_FP32 = 0,
  PINN_PREC_TF32 = 1,          /* on-chip fused kernel where it applies (H <= 128, 2..4 hidden layers), else streamed */
  PINN_PREC_TF32_STREAMED = 2, /* always the layer-streamed tensor-core kernels (for comparison) */
  PINN_PREC_FP32X3 = 3         /* layer-streamed tensor-core kernels with 3xTF32 operand splitting: every
                                * multiplicand x is fed as tf32(x) + tf32(x - tf32(x)) and each product as
                                * lo*hi + hi*lo + hi*hi with FP32 accumulation -- FP32-grade products
                                * (~2^-22 relative): whole-gradient norm-wise error <= 1e-5 (observed <= 5e-6,
                                * PINN_PREC_FP32: <= 1e-6), single small layers up to 2e-5 */
};
enum {
  PINN_E_INVALID = -1,     /* malformed descriptor                        */
  PINN_E_UNSUPPORTED = -2, /* valid but no kernel for it (width, channels) */
  PINN_E_WORKSPACE = -3,   /* workspace too small                          */
  PINN_E_NO_DEVICE = -4    /* no sm_100 device / kernel image not loadable */
};

/* The MLP of a NetNode: idrlnet/architecture/mlp.py:18-76 (Linear + activation
 * stack, last layer linear), evaluated as idrlnet/net.py:14-36 does (inputs
 * concatenated in NetNode.inputs order).  W[l] is the EFFECTIVE weight of linear
 * layer l, row-major [out, in] as torch.nn.Linear stores it (for weight-norm
 * layers, idrlnet/architecture/layer.py:42-43, the caller passes g*v/||v||). */
typedef struct pinn_mlp {
  int32_t n_in;     /* d                                                   */
  int32_t n_out;    /* o                                                   */
  int32_t n_hidden; /* L >= 1 hidden layers, all `width` wide              */
  int32_t width;    /* H                                                   */
  int32_t act;      /* PINN_ACT_*                                          */
  int32_t precision; /* PINN_PREC_*: arithmetic of the wide (H > 32) layer GEMMs */
  const float* W[PINN_MAX_LAYERS]; /* L+1 pointers                         */
  const float* b[PINN_MAX_LAYERS];
} pinn_mlp;

/* Which input-derivative "jet" channels to propagate.  Replaces the nested
 * torch.autograd.grad calls of idrlnet/variable.py:161-208: channel 0 is the
 * value, channel 1+i is d/d(input first_in[i]), channel 1+n_first+s is
 * d2/d(input first_in[s])^2  (s < n_second <= n_first).  Extra channels follow in
 * this order (narrow nets only; at most one kind at a time):
 *   n_mixed (0/1): d2 / d(input first_in[0]) d(input first_in[1])   (needs n_first >= 2)
 *   n_high (0..2): d3 / d(input first_in[0])^3, then d4 / d(input first_in[0])^4
 *                  (needs n_second >= 1; the derivative chain of
 *                  examples/euler_beam/euler_beam.py:47-54). */
typedef struct pinn_jets {
  int32_t n_first;
  int32_t n_second;
  int32_t first_in[PINN_MAX_FIRST];
  int32_t n_mixed;
  int32_t n_high;
} pinn_jets;
/* number of channels of a jet plan */
#define PINN_JET_CHANNELS(j) (1 + (j)->n_first + (j)->n_second + (j)->n_mixed + (j)->n_high)

/* A residual is a polynomial in the symbols (what idrlnet/pde.py:65-79 +
 * idrlnet/torch_util.py:25-39 lambdify; every PDE of idrlnet/pde_op/equations.py
 * is of this form): sum over terms of coef * prod(sym[fac[i]]). */
typedef struct pinn_term {
  float coef;
  uint8_t n_fac;
  uint8_t fac[PINN_MAX_FACTORS];
} pinn_term;

typedef struct pinn_equation {
  int32_t term_begin, term_end; /* range in pinn_domain.terms               */
  const float* target;          /* [N] or NULL -> target_const              */
  const float* lambda;          /* [N] or NULL -> lambda_const              */
  float target_const;
  float lambda_const;           /* 1.0 when the domain has no lambda_<name> */
} pinn_equation;

/* One sampling domain (idrlnet/data.py:15-129) and its share of the loss
 * (idrlnet/solver.py:353-408, idrlnet/variable.py:40-80):
 *   loss_d = sum_eq sum_n f(r_eq[n] - target[n]) * lambda_eq[n] * area[n]
 * and the step differentiates sigma * loss_d. */
typedef struct pinn_domain {
  int64_t n_points;
  const float* coords[PINN_MAX_IN]; /* one [N] column per net input (SoA)   */
  const float* aux[PINN_MAX_AUX];
  const float* area; /* [N] or NULL -> area_const                           */
  float area_const;
  float sigma;
  int32_t n_aux;
  int32_t loss_kind; /* PINN_LOSS_*                                          */
  int32_t n_eq;
  int32_t n_terms;
  pinn_equation eq[PINN_MAX_EQ];
  pinn_term terms[PINN_MAX_TERMS];
} pinn_domain;

/* ---- queries ---------------------------------------------------------- */
PINN_API int pinn_abi_version(void);
/* Number of floats of the flat gradient: sum over layers of out*in + out,
 * in layer order, W (row-major [out,in]) then b -- the order of
 * torch.nn.Module.parameters() for a plain Linear stack. */
PINN_API int64_t pinn_param_count(const pinn_mlp* mlp);
/* Bytes of scratch the step functions need for this net (per-CTA gradient and
 * loss partials).  `n_domains` = how many pinn_step_* calls are accumulated
 * before pinn_step_finalize. */
PINN_API int64_t pinn_workspace_bytes(const pinn_mlp* mlp, const pinn_jets* jets, int32_t n_domains);
/* 1 if a fused kernel exists for (mlp, jets); 0 otherwise. */
PINN_API int pinn_supported(const pinn_mlp* mlp, const pinn_jets* jets);
/* Kernels launched by this library in this process (bench.py "gpu_launches"). */
PINN_API int64_t pinn_launch_count(void);
PINN_API const char* pinn_last_error(void);

/* ---- the fused training step ------------------------------------------ */
/* Forward jets + residual + weighted loss + reverse pass for ONE domain, no
 * autograd graph materialised.  Replaces, for that domain,
 * VertexTaskPipeline.forward_pipeline (idrlnet/graph.py:170-193: WrapEvaluate,
 * Variables.differentiate_, PdeEvaluate), Solver.compute_loss
 * (idrlnet/solver.py:353-408) and loss.backward() (idrlnet/solver.py:338).
 * Results go to per-CTA partials inside `workspace`; `domain_slot` selects the
 * loss slot, `accumulate` != 0 adds the gradient partials to those of earlier
 * calls (several domains sharing the net).  Call pinn_step_finalize afterwards. */
PINN_API int pinn_step_fused(const pinn_mlp* mlp, const pinn_jets* jets, const pinn_domain* dom,
                    void* workspace, int64_t workspace_bytes, int32_t domain_slot,
                    int32_t accumulate, void* stream);

/* The loss of one domain WITHOUT the reverse pass: forward jets, residual, weighted loss into the domain's
 * loss slot (same arguments and slot protocol as pinn_step_fused; finish with pinn_step_finalize for the losses --
 * its gradient output is unspecified after loss-only calls, so do not mix the two kinds of call in one
 * accumulate / finalize sequence).  Replaces the second,
 * gradient-free evaluation every reference step ends with (idrlnet/solver.py:341-344: re-sample, forward,
 * compute_loss -- the value train_pipe returns and logs): about a third of the cost of the full step. */
PINN_API int pinn_loss_fused(const pinn_mlp* mlp, const pinn_jets* jets, const pinn_domain* domain,
                    void* workspace, int64_t workspace_bytes, int32_t domain_slot, int32_t accumulate,
                    void* stream);

/* Deterministic reduction of the per-CTA partials:
 *   flat_grad[pinn_param_count] (+)= d(sum_d sigma_d*loss_d)/d(W,b)   (add_to_grad != 0 adds)
 *   losses[n_domains]            = loss_d (un-weighted by sigma, as
 *                                  solver.loss_component holds it)      */
PINN_API int pinn_step_finalize(const pinn_mlp* mlp, const pinn_jets* jets, void* workspace,
                       int64_t workspace_bytes, int32_t n_domains, float* flat_grad,
                       int32_t add_to_grad, float* losses, void* stream);

/* ---- finalize + all-reduce over peer memory (one kernel) --------------- */
/* Multi-GPU step of a NARROW net: the deterministic reduction of pinn_step_finalize fused with the SUM over the ranks
 * of one NVLink / NVSwitch box, so that no collective launch sits between the step kernels and the optimizer.  Replaces
 * pinn_step_finalize followed by torch.distributed.all_reduce(flat) (idrlnet has no multi-GPU path; the all-reduce is
 * the one SURVEY.md 8(e) prescribes).  Every rank passes the same `peers` table: base[r] is rank r's symmetric region
 * (pinn_allreduce_region_bytes bytes, zero-initialised once, mapped into this process -- e.g. the buffer_ptrs of a
 * torch.distributed._symmetric_memory rendezvous).  Each block of the kernel writes its 32 reduced values into slot
 * `rank` of every peer's region, publishes a per-block flag (st.release.sys), waits for the same block of every rank
 * (ld.acquire.sys, bounded: traps after ~60 s instead of hanging) and adds the `world` slots in rank order -- bitwise the
 * same sum on every rank.  `epoch` must increase by one per call on every rank, starting at 1 (slots and flags are
 * double-buffered on its parity).  out[n_params + n_domains] = all-reduced gradient, then the all-reduced losses. */
#define PINN_MAX_RANKS 16
typedef struct pinn_peers {
  int32_t rank, world;
  uint32_t epoch;
  int32_t reserved;
  void* base[PINN_MAX_RANKS];
} pinn_peers;
PINN_API int64_t pinn_allreduce_region_bytes(const pinn_mlp* mlp, const pinn_jets* jets, int32_t n_domains, int32_t world);
PINN_API int pinn_step_finalize_allreduce(const pinn_mlp* mlp, const pinn_jets* jets, void* workspace, int64_t workspace_bytes,
                                 int32_t n_domains, float* out, const pinn_peers* peers, void* stream);

/* ---- the generic path (residual evaluated by the caller) --------------- */
/* Jets only: out[(c*n_out + o)*N + n] = channel c of output o at point n.
 * Replaces WrapEvaluate + Variables.differentiate_ for names such as u__x__x
 * when the residual is not a polynomial (idrlnet/graph.py:172-175), and serves
 * Solver.infer_step (idrlnet/solver.py:410-420).  `coords` is a HOST array of
 * n_in device pointers.  `workspace` (pinn_workspace_bytes(mlp, jets, 1)) is
 * needed by the wide path only and may be NULL for narrow nets. */
PINN_API int pinn_jet_forward(const pinn_mlp* mlp, const pinn_jets* jets, int64_t n_points,
                     const float* const* coords, float* out, void* workspace,
                     int64_t workspace_bytes, void* stream);

/* Reverse pass for caller-supplied adjoints out_bar (same layout as `out`):
 * gradient partials into workspace as pinn_step_fused does. */
PINN_API int pinn_jet_backward(const pinn_mlp* mlp, const pinn_jets* jets, int64_t n_points,
                      const float* const* coords, const float* out_bar, void* workspace,
                      int64_t workspace_bytes, int32_t accumulate, void* stream);

/* ---- parameter-sized helpers ------------------------------------------- */
/* W[j,:] = g[j] * v[j,:] / ||v[j,:]||  (torch.nn.utils.weight_norm, dim=0;
 * idrlnet/architecture/layer.py:42-43). */
PINN_API int pinn_weight_norm_forward(const float* g, const float* v, float* W, int32_t out_dim,
                             int32_t in_dim, void* stream);
/* (dg, dv) from dW. */
PINN_API int pinn_weight_norm_backward(const float* g, const float* v, const float* dW, float* dg,
                              float* dv, int32_t out_dim, int32_t in_dim, void* stream);

/* The same two operations for every weight-norm layer of a net in ONE launch (a 5-layer MLP otherwise spends ten
 * 3-microsecond launches per training step on them); per-row arithmetic identical to the calls above. */
typedef struct pinn_wn_layer {
  const float* g;  /* [out_dim]                 */
  const float* v;  /* [out_dim, in_dim]         */
  const float* dW; /* backward: [out_dim, in_dim] */
  float* W;        /* forward:  [out_dim, in_dim] */
  float* dg;       /* backward: [out_dim]         */
  float* dv;       /* backward: [out_dim, in_dim] */
  int32_t out_dim, in_dim;
} pinn_wn_layer;
PINN_API int pinn_weight_norm_forward_batch(const pinn_wn_layer* layers, int32_t n_layers, void* stream);
PINN_API int pinn_weight_norm_backward_batch(const pinn_wn_layer* layers, int32_t n_layers, void* stream);

/* ---- adaptive re-sampling: forward-only residuals and top-k ------------- */
/* out[e*N + n] = value of equation e at point n (no target subtracted, no loss):
 * what Solver.infer_step (idrlnet/solver.py:410-420) returns for a PdeNode output
 * such as "AllenCahn_u" in examples/allen_cahn/allen_cahn.py:134-136.  Jets by
 * pinn_jet_forward into the workspace, then one elementwise kernel.  Only the
 * equation table, coords, aux and n_points of `domain` are read. */
PINN_API int64_t pinn_residual_workspace_bytes(const pinn_mlp* mlp, const pinn_jets* jets, int64_t n_points);
PINN_API int pinn_residual_forward(const pinn_mlp* mlp, const pinn_jets* jets, const pinn_domain* domain,
                          float* out, void* workspace, int64_t workspace_bytes, void* stream);
/* Indices of the k largest |values[i]| in descending order, ties by ascending index
 * (= np.argsort(-abs(values), kind="stable")[:k], the selection of
 * examples/allen_cahn/allen_cahn.py:137-141), 1 <= k <= 4096, k <= n < 2^32.
 * NaNs order above +inf.  val_out (may be NULL) receives values[idx_out[i]]. */
PINN_API int64_t pinn_topk_workspace_bytes(int32_t k);
PINN_API int pinn_topk_abs(const float* values, int64_t n, int32_t k, int64_t* idx_out, float* val_out,
                  void* workspace, int64_t workspace_bytes, void* stream);

/* ---- on-device sampling ------------------------------------------------ */
/* Uniform points in an axis-aligned box (idrlnet/geo_utils/geo.py:316-343
 * `_ranged_sample`; the docs' "future versions will sample on GPU"), Philox4x32-10
 * keyed by (seed, stream_id), counter = point index + offset: columns[d][n] in
 * [lo[d], hi[d]).  Also writes the rectangle / interval signed distance
 * (idrlnet/geo_utils/geo_obj.py:30-41,109-167) over the first `sdf_dims`
 * columns into `sdf` when it is non-NULL. */
PINN_API int pinn_sample_box(float* const* columns, int32_t n_dims, const float* lo, const float* hi,
                    int64_t n_points, uint64_t seed, uint64_t stream_id, uint64_t offset,
                    float* sdf, int32_t sdf_dims, void* stream);

/* ---- measurement helper ------------------------------------------------- */
/* Launches `blocks` x 256 threads each running 16 independent FP32 FMA chains of
 * `iters` steps; returns the number of flops issued (time it with CUDA events to
 * get the FP32-pipe roofline denominator), or a negative cudaError_t. */
PINN_API int64_t pinn_fp32_fma_probe(int32_t iters, int32_t blocks, float* scratch, void* stream);
/* The same chains as packed pairs (fma.rn.f32x2, SASS FFMA2). */
PINN_API int64_t pinn_fp32_fma2_probe(int32_t iters, int32_t blocks, float* scratch, void* stream);

/* Dense TF32 tensor-core rate: one CTA per SM, each issuing `iters` x 8 tcgen05.mma (M=128, N=256, K=8, kind::tf32)
 * on shared-memory-resident operands; returns the number of flops issued (time it with CUDA events: the roofline
 * denominator of the wide tensor-core kernels), or a negative error.  `scratch` needs one float per SM. */
PINN_API int64_t pinn_tf32_mma_probe(int32_t iters, float* scratch, void* stream);

/* Self-test of the tcgen05 plumbing used by the wide tensor-core path: one CTA computes
 * D[128][N] = A * B^T in TF32 (FP32 accumulate in TMEM).  variant 0: A[128][K], B[N][K]
 * (K-major operands); variant 1: A[K][128], B[K][N] (MN-major operands).  *status (device
 * int) is set to 1 if the MMA completion barrier timed out. */
PINN_API int pinn_tc_selftest(int variant, int N, int K, const float* A, const float* B, float* D,
                              int* status, void* stream);

/* ---- on-device sampling of CSG solids ----------------------------------- */
/* A solid is a postfix program over signed distances (positive inside): primitives push one value
 * (closed forms of idrlnet/geo_utils/geo_obj.py), operators combine the top of the stack as
 * idrlnet/geo_utils/geo.py:246-300 does (union = max, difference = min(a, -b), intersection = min). */
#define PINN_MAX_SHAPE_OPS 32
enum {
  PINN_SHAPE_INTERVAL = 0,  /* p = (x0, x1, axis)             Line1D            */
  PINN_SHAPE_RECT = 1,      /* p = (x0, y0, x1, y1)           Rectangle         */
  PINN_SHAPE_CIRCLE = 2,    /* p = (cx, cy, r)                Circle            */
  PINN_SHAPE_CHANNEL2D = 3, /* p = (y0, y1)                   Tube2D            */
  PINN_SHAPE_BOX = 4,       /* p = (x0, y0, z0, x1, y1, z1)   Box               */
  PINN_SHAPE_SPHERE = 5,    /* p = (cx, cy, cz, r)            Sphere            */
  PINN_CSG_UNION = 16,
  PINN_CSG_DIFFERENCE = 17,
  PINN_CSG_INTERSECTION = 18,
  PINN_CSG_COMPLEMENT = 19
};
typedef struct pinn_shape_op {
  int32_t op;
  float p[6];
} pinn_shape_op;

/* Geometry.sample_interior on the device (idrlnet/geo_utils/geo.py:204-244): candidate i (0 <= i <
 * n_candidates) is the point pinn_sample_box would draw for (seed, stream_id, offset + i) in the box
 * [lo, hi); it is kept when the program's signed distance is > 0.  Survivors are written in candidate
 * order to columns[d][0 .. *count) and sdf[0 .. *count) (sdf may be NULL); at most `capacity` are written
 * and *count (DEVICE int64) is clamped to it.  `prog` is a HOST array.  Three launches, no atomics:
 * the output is a deterministic function of the arguments. */
PINN_API int64_t pinn_sample_csg_workspace_bytes(int64_t n_candidates);
PINN_API int pinn_sample_csg(float* const* columns, int32_t n_dims, const float* lo, const float* hi,
                    int64_t n_candidates, int64_t capacity, uint64_t seed, uint64_t stream_id, uint64_t offset,
                    const pinn_shape_op* prog, int32_t n_ops, float* sdf, int64_t* count, void* workspace,
                    int64_t workspace_bytes, void* stream);

/* ---- on-device evaluation of parametric boundary pieces and symbolic expressions ------------------------------ */
/* Edge.sample on the device (idrlnet/geo_utils/geo.py:80-108): every point i draws `n_vars` parameters -- Philox4x32-10,
 * stream_id + v/4, counter offset + i, the generator of pinn_sample_box -- uniformly in [lo[v], hi[v]) and evaluates
 * `n_out` postfix programs over them (the edge's coordinate / normal / measure expressions, parameter columns, the
 * solid's signed distance composed with the coordinates), writing out[o][i].  `ops` / `begin` are HOST arrays:
 * program o is ops[begin[o] .. begin[o+1]).  Replaces the per-call NumPy lambdify + H2D copies of the reference. */
#define PINN_EXPR_MAX_OUT 12
#define PINN_EXPR_MAX_VARS 8
#define PINN_EXPR_MAX_OPS 448
#define PINN_EXPR_STACK 16
enum {
  PINN_EX_CONST = 0, /* push imm                     */
  PINN_EX_VAR = 1,   /* push parameter (int)imm      */
  PINN_EX_ADD = 2, PINN_EX_SUB = 3, PINN_EX_MUL = 4, PINN_EX_DIV = 5, PINN_EX_MIN = 6, PINN_EX_MAX = 7, PINN_EX_POW = 8,
  PINN_EX_ATAN2 = 9, /* binary: a op b with b on top of the stack */
  PINN_EX_NEG = 16, PINN_EX_ABS = 17, PINN_EX_SQRT = 18, PINN_EX_SIN = 19, PINN_EX_COS = 20, PINN_EX_TAN = 21, PINN_EX_EXP = 22,
  PINN_EX_LOG = 23, PINN_EX_TANH = 24, PINN_EX_SIGN = 25, PINN_EX_HEAVISIDE = 26, PINN_EX_SQUARE = 27, PINN_EX_RCP = 28,
  PINN_EX_SCALE = 29, /* top *= imm */
  PINN_EX_OFFSET = 30, /* top += imm */
  PINN_EX_ASIN = 31, PINN_EX_ACOS = 32, PINN_EX_ATAN = 33, PINN_EX_SINH = 34, PINN_EX_COSH = 35
};
typedef struct pinn_expr_op {
  int32_t op;
  float imm;
} pinn_expr_op;
PINN_API int pinn_sample_program(float* const* out, int32_t n_out, const pinn_expr_op* ops, const int32_t* begin,
                                 const float* lo, const float* hi, int32_t n_vars, int64_t n_points, uint64_t seed,
                                 uint64_t stream_id, uint64_t offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PINN_B200_H */
