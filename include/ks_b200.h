/* ks_b200.h -- C ABI of the B200-native Kuramoto-Sivashinsky orbit engine (libks_b200.so).
 *
 * Drop-in boundary for the spatiotemporal KS hot path of chaotic-systems/orbithunter.  The reference is pure
 * Python and has NO FFI; its boundary is the Orbit-subclass method protocol (template/class_template.py:277-461).
 * Each entry point below is the batched device implementation of one of those methods; the Python mirror in
 * orbithunter_b200/ (same class / method / kwarg names as the reference) binds them with ctypes.  INTEGRATION.md
 * shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - all array arguments are DEVICE pointers to C-contiguous float64 with a leading batch axis;
 *   - `modes[B][rows][cols]`: the reference's real-packed spatiotemporal Fourier modes; rows = n-1,
 *     cols = m-2 (OrbitKS, RelativeOrbitKS) or m/2-1 (ShiftReflectionOrbitKS, AntisymmetricOrbitKS)
 *     (orbithunter/ks/orbits.py:1403-1412, :3677-3686, :3254-3263);
 *   - `params[B][3]` = (T, L, S) = parameter labels ("t","x","s") (ks/orbits.py:207-221);
 *   - `cmask`: bit0/1/2 set = t/x/s held constant (the reference's `constraints` dict, core.py:1617-1664);
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls only enqueue work;
 *   - `ws`/`ws_bytes`: caller-owned scratch of at least ks_workspace_bytes(...) bytes (may be NULL when that is 0);
 *   - return value 0 = enqueued; negative = argument / launch error (message via ks_last_error()); never throws.
 * No entry point allocates device memory or has a CPU fallback.
 */
#ifndef KS_B200_H
#define KS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { KS_CLS_ORBIT = 0, KS_CLS_RELATIVE = 1, KS_CLS_SHIFT_REFLECTION = 2, KS_CLS_ANTISYMMETRIC = 3 };
enum { KS_BASIS_FIELD = 0, KS_BASIS_SPATIAL_MODES = 1, KS_BASIS_MODES = 2 };
enum { KS_ERR_ARG = -1, KS_ERR_SHAPE = -2, KS_ERR_WORKSPACE = -3, KS_ERR_LAUNCH = -4 };

int ks_b200_abi_version(void);
const char* ks_last_error(void);

/* test hook: route 32x32 through the generic-size kernel instead of the register-resident one (cross-check). */
void ks_debug_force_generic(int on);
/* tuning hook, 32x32 kernel build variant: 3 = two orbits per warp, 7 warps/SM, global reads staged with cp.async
 * (all ops); 4 (default) = one orbit per warp, 8 warps/SM, all state in registers (all ops);
 * 5 = one orbit per warp, 12 warps/SM, field rows stashed in shared memory (eqn and costgrad; other ops use 3);
 * 6 = variant 4's work split with the two warps of each scheduler taking turns on the fp64 pipe (eqn and costgrad);
 * 7 = variant 4 for eqn / costgrad, variant 3 for matvec / rmatvec (the default before the one-orbit-per-warp kernel
 * learned those two ops). */
void ks_debug_warp32_variant(int variant);
/* tuning hook, dense Newton step at 32x32: 2 (default) = matrix-free -- J J^T column by column as J (J^T e_j) by two launches
 * of the one-orbit-per-warp kernel, J^T y and J dx of the solve as single rmatvec / matvec launches, no Jacobian in memory;
 * 1 = Jacobian by that kernel's Jacobian mode (unit vectors made in the kernel), J J^T by batched library GEMMs;
 * 0 = Jacobian by a batched matvec over a stored identity and replicated orbits, J J^T by GEMMs (the path of other sizes). */
void ks_debug_dense_jacobian_mode(int mode);
/* tuning hook, dense Newton step: Cholesky blocks (of 64) per outer panel of the two-level factorisation; 4 (default) updates
 * the matrix to the right of a panel with rank 256 at once, 1 with rank 64 after every block. */
void ks_debug_dense_outer_blocks(int blocks);
/* tuning hook, 32x32 kernels: nanoseconds between the start of the warps that share one SM scheduler (default 0). */
void ks_debug_warp32_stagger(int ns);
/* tuning hook, 32x32 variant 6: 1 (default) = token passing between the two warp groups, 0 = free-running warps. */
void ks_debug_warp32_pingpong(int on);
/* tuning hook, large-grid (n, m in {64, 128, 256}) tile path: enabled = 0 routes those shapes through the generic
 * kernel (cross-check), 1 (default) = tile path with the CTA-resident fused kernel at 64x64, 2 = multi-pass tiles for
 * every shape; chunk_mb > 0 sets the workspace budget that bounds how many orbits one pass sequence covers. */
void ks_debug_tile(int enabled, int chunk_mb);
/* tuning hook, Newton-Krylov drivers at 64x64: 1 = N(u, u) is computed once per outer iteration and reused by every matvec /
 * rmatvec of the inner solve (6 instead of 8 line-transform sweeps per operator application), identical bits; 0 (default) =
 * every application recomputes it.  Measured at BASELINE configs[2]: -1 % (Relative) .. +7 % (Antisymmetric) on the
 * Newton-GMRES(50) wall time (profiles/r02oc_operator_cache.jsonl), so it stays off. */
void ks_debug_operator_cache(int on);
/* tuning / test hook, multi-pass tile path: 0 (default) = one CTA per tile, loads into registers; 1 = the row and column
 * pass kernels are persistent and stage their tiles with the Tensor Memory Accelerator (cp.async.bulk.tensor boxes in and out,
 * cp.async.bulk lines, two-stage shared-memory ring on mbarriers); 2 / 3 = TMA rows only / columns only.  Bit-identical
 * results; measured slower (two resident CTAs per SM instead of three), see csrc/ks_tile_tma.cuh. */
void ks_debug_tile_tma(int mode);
/* tuning / test hook, 32x32 adjoint descent: passes every orbit runs inside one launch of the multi-pass kernel with its
 * state on chip (default 128); 0 = the one-pass-per-launch kernel with the state in global memory (cross-check). */
void ks_debug_adjoint_inner(int passes);
/* test hook, GMRES step kernel: 1 (default) = thread-block cluster per orbit (1/2/4/8 CTAs by vector length), vector
 * slices in registers; 2/4/8 = at least that many CTAs per orbit; 0 = one CTA per orbit, vectors in HBM (cross-check;
 * also used automatically when N > 65536). */
void ks_debug_gmres_cluster(int mode);

/* rows*cols of the modes basis for (cls, n, m); 0 if the shape is not supported. */
int ks_mode_size(int cls, int n, int m);
/* scratch bytes required by any of the calls below for this shape and batch. */
size_t ks_workspace_bytes(int cls, int n, int m, int batch);

/* OrbitKS._populate_state (ks/orbits.py:1738-1849, default 'laplace' x 'gaussian' modulation) for a batch of integer
 * seeds: the reference's np.random.seed(seed); randn(n-1, cols) stream (legacy MT19937 + polar Box-Muller) reproduced on
 * the device, one warp per seed.  seeds [B] uint32, params [B,3] (T, L, S), modes_out [B, Nm]. */
int ks_populate_state(int cls, int n, int m, int batch, const uint32_t* seeds, const double* params, double* modes_out,
                      void* stream);
/* The same with the reference's other modulation profiles (the `spatial_modulation`, `temporal_modulation`, `xshift`,
 * `tscale`, `xscale`, `xvar`, `tvar` kwargs of _populate_state, ks/orbits.py:1772-1842).  `opt` is a HOST struct (NULL =
 * defaults); NaN in a scale / variance field = the reference's default derived from (T, L). */
enum { KS_SMOD_NONE = 0, KS_SMOD_GAUSSIAN, KS_SMOD_LAPLACE, KS_SMOD_LAPLACE_SQRT, KS_SMOD_PLATEAU_LINEAR,
       KS_SMOD_EXPONENTIAL_LINEAR, KS_SMOD_FLAT_TOP };
enum { KS_TMOD_NONE = 0, KS_TMOD_GAUSSIAN, KS_TMOD_LAPLACE, KS_TMOD_FLAT_TOP, KS_TMOD_LAPLACE_PLATEAU, KS_TMOD_TRUNCATE };
enum { KS_XSHIFT_NONE = 0, KS_XSHIFT_SQRT, KS_XSHIFT_LOG, KS_XSHIFT_LIN };
typedef struct KsPopulateOptions {
  int spatial_modulation, temporal_modulation, xshift, reserved;
  double tscale, xscale, xvar, tvar;
} KsPopulateOptions;
int ks_populate_state_ex(int cls, int n, int m, int batch, const uint32_t* seeds, const double* params,
                         const KsPopulateOptions* opt, double* modes_out, void* stream);

/* Orbit.resize for a batch (core.py:1167-1226 with OrbitKS._pad / _truncate, ks/orbits.py:1428-1564 and the SR / Anti
 * overrides :3325-3394, :3735-3804): modes of an (n, m) grid -> modes of an (n_new, m_new) grid by zero padding /
 * truncation of the time and space spectra, scaled by sqrt(n_new/n) then sqrt(m_new/m) as the reference does axis by
 * axis.  in [B][(n-1) cols(m)], out [B][(n_new-1) cols(m_new)]. */
int ks_resize_modes(int cls, int n, int m, int n_new, int m_new, int batch, const double* in, double* out, void* stream);

/* Orbit.rescale (core.py:1831-1853) on FIELD arrays [B][n][m], in place: method 0 'inf' (max |u| -> magnitude),
 * 1 'L1' (numpy.linalg.norm(field, ord=1) = largest column sum of |u|), 2 'L2' (Frobenius norm), 3 'LP'
 * (sign(u) |u|^magnitude).  norms_out [B] (may be NULL) receives the norm that was divided out (0 for 'LP'). */
int ks_rescale_field(int n, int m, int batch, double* field, double magnitude, int method, double* norms_out, void* stream);

/* The norms OrbitKS.preprocess (ks/orbits.py:1566-1596; Relative :2969-3008) thresholds, from the modes (the transforms are
 * orthogonal on the retained coefficients, so |u| = |modes| and |u_t| = |dt(modes)|): norms_out [B][2] = (|u|, |u_t|);
 * code_out [B] = 1 if |u| is under the class's threshold (zero solution), 2 if |u_t| is (equilibrium), else 0; T == 0
 * gives 2 (OrbitKS) / 1 (RelativeOrbitKS) as in the reference. */
int ks_preprocess(int cls, int n, int m, int batch, const double* modes, const double* params, int32_t* code_out,
                  double* norms_out, void* stream);

/* OrbitKS.transform (ks/orbits.py:231-303) and the class-specific overrides: any basis -> any basis. */
int ks_transform(int cls, int n, int m, int batch, int from_basis, int to_basis, const double* in, double* out,
                 void* ws, size_t ws_bytes, void* stream);

/* OrbitKS.eqn (ks/orbits.py:526-555) + Orbit.cost (core.py:986-1008): F(u) and 1/2 |F|^2 (cost may be NULL). */
int ks_eqn(int cls, int n, int m, int batch, const double* modes, const double* params, double* f_out,
           double* cost_out, void* ws, size_t ws_bytes, void* stream);

/* OrbitKS.jacobian (ks/orbits.py:557-657; Relative :2662-2760; ShiftReflection / Antisymmetric inherit): the dense Jacobian of
 * every orbit of the batch, transposed: jac_t[b][col][:] is column `col` (state slots in cdof order, then the free
 * parameters) -- J(u_b) applied to the unit vector of that slot.  32x32: one launch of the one-orbit-per-warp kernel, which
 * makes the unit vectors itself; other sizes: one batched matvec over a stored identity per orbit. */
size_t ks_jacobian_workspace_bytes(int cls, int n, int m, int batch, uint32_t cmask);
int ks_jacobian(int cls, int n, int m, int batch, const double* modes, const double* params, uint32_t cmask, double* jac_t,
                void* ws, size_t ws_bytes, void* stream);

/* OrbitKS.matvec (ks/orbits.py:659-709; Relative :2623-2660): J(u) v, with vparams = (dT, dL, dS). */
int ks_matvec(int cls, int n, int m, int batch, const double* modes, const double* params, const double* vmodes,
              const double* vparams, uint32_t cmask, double* out_state, void* ws, size_t ws_bytes, void* stream);

/* OrbitKS.rmatvec (ks/orbits.py:711-755, :1984-2045; Relative :3059-3115): J(u)^T v -> state + 3 parameters. */
int ks_rmatvec(int cls, int n, int m, int batch, const double* modes, const double* params, const double* vmodes,
               uint32_t cmask, double* out_state, double* out_params, void* ws, size_t ws_bytes, void* stream);

/* eqn followed by OrbitKS.costgrad (ks/orbits.py:757-791), fused: cost = 1/2|F|^2, grad = J^T F, optionally
 * rescaled by OrbitKS.precondition (:1034-1089).  f_out may be NULL. */
int ks_costgrad(int cls, int n, int m, int batch, const double* modes, const double* params, uint32_t cmask,
                int preconditioning, double* cost_out, double* grad_state, double* grad_params, double* f_out,
                void* ws, size_t ws_bytes, void* stream);

/* OrbitKS.precondition (ks/orbits.py:1034-1089): state / (|w_j| + q^2 + q^4); unconstrained t *= T^-pexp_t,
 * x *= L^-pexp_x (reference default pexp = (1, 4)); `params` supplies (T, L) (the reference's `pmult`).
 * gparams_in/out may both be NULL. */
int ks_precondition(int cls, int n, int m, int batch, const double* state_in, const double* gparams_in,
                    const double* params, uint32_t cmask, double pexp_t, double pexp_x, double* state_out,
                    double* gparams_out, void* stream);

/* OrbitKS.dt (ks/orbits.py:379-440; axis 0, `in` in the modes basis) and OrbitKS.dx (:442-524; axis 1, `in` with
 * [Re|Im] columns, i.e. spatial_modes or full-class modes, or the k columns of SR/Anti modes for even orders),
 * any non-negative order.  in/out: [batch][rows][cols], out must not alias in. */
int ks_deriv(int cls, int n, int m, int batch, int axis, int order, int rows, int cols, const double* in,
             const double* params, double* out, void* stream);

/* hunt(..., 'adj'): _adjoint_descent (optimize.py:367-509) with _process_correction (:2115-2239) for a whole batch.
 * modes/params are updated in place.  Per-orbit outputs: status (-1 has-not-failed, 0 min step, 2 maxiter,
 * 3 insufficient decrease -- hunt()'s final promotion -1 -> 1 is left to the caller, optimize.py:351-358), nit,
 * cost = 1/2|F|^2 at the returned orbit, final step size.  Unlike the evaluation calls this one SYNCHRONISES the
 * stream every `check_every` passes (<= 0: 64) to stop as soon as every orbit has left the loop; `passes_out`
 * (host int, may be NULL) receives the number of batched costgrad launches.  Workspace: see the _workspace_bytes. */
size_t ks_adjoint_descent_workspace_bytes(int cls, int n, int m, int batch);
int ks_adjoint_descent(int cls, int n, int m, int batch, double* modes, double* params, uint32_t cmask, double tol,
                       double ftol, int maxiter, double min_step, int preconditioning, double step_size,
                       int32_t* status_out, int32_t* nit_out, double* cost_out, double* step_out, int check_every,
                       int* passes_out, void* ws, size_t ws_bytes, void* stream);

/* hunt(..., 'gmres' | 'lsqr' | 'lsmr'): _scipy_sparse_linalg_solver_wrapper (optimize.py:1196-1404) for a whole batch.
 * method 0 = GMRES on the normal equations A^T A dx = -A^T F (optimize.py:1760-1797; restart / inner_maxiter = restart
 * cycles / rtol / atol as scipy.sparse.linalg.gmres; preconditioning = OrbitKS.preconditioner as M);
 * method 1 = LSQR on the rectangular system J dx = -F (optimize.py:1731-1759; inner_maxiter = iter_lim, atol, btol,
 * conlim as scipy.sparse.linalg.lsqr, damp = 0); method 2 = LSMR on the same system (optimize.py:1319-1321;
 * inner_maxiter = maxiter, atol, btol, conlim as scipy.sparse.linalg.lsmr, damp = 0); when every parameter is
 * constrained (cdof == Nm) method 0 solves the square system J dx = -F itself, as the reference does (optimize.py:1803-1850).
 * Each outer step takes the full step first and halves while
 * next_cost > cost and step > min_step, then _process_correction.  modes/params updated in place; status/nit/cost per
 * orbit as for ks_adjoint_descent.  Synchronises the stream while polling.  counters_out (host, 4 x int64, may be
 * NULL): operator applications, solver steps, line-search evaluations, outer iterations. */
size_t ks_newton_krylov_workspace_bytes(int cls, int n, int m, int batch, int method, int restart);
int ks_newton_krylov(int cls, int n, int m, int batch, double* modes, double* params, uint32_t cmask, int method,
                     double tol, double ftol, int maxiter, double min_step, int preconditioning, int restart,
                     int inner_maxiter, double rtol, double atol, double btol, double conlim, int32_t* status_out,
                     int32_t* nit_out, double* cost_out, long long* counters_out, int check_every, void* ws,
                     size_t ws_bytes, void* stream);

/* hunt(..., 'lstsq'): _lstsq (optimize.py:855-1022, no preconditioning) for a whole batch.  Each outer step takes, for the
 * orbits still in the loop, dx = the minimum-norm solution of J dx = -F (the reference: OrbitKS.jacobian, ks/orbits.py:557-657,
 * then scipy.linalg.lstsq) as J^T (J J^T)^-1 (-F).  J J^T: at 32x32 column by column as J (J^T e_j) -- an rmatvec launch over
 * unit vectors made in the kernel, then a matvec launch over the result, no Jacobian in memory; other sizes: the dense
 * Jacobian (one batched matvec over the identity) and batched GEMMs.  Then
 * a blocked Cholesky factorisation with `refine` rounds of iterative refinement (breakdown -> Tikhonov shift), then the
 * same line search / _process_correction as ks_newton_krylov.  `chunk` orbits are factored at a time (workspace about
 * 3 x 8 x (Nm + 3)^2 bytes per orbit of the chunk); the batched fp64 GEMMs go through cuBLAS (loaded at run time).
 * counters_out: operator applications per orbit spent on J J^T / the Jacobian, dense solves, line-search evaluations,
 * outer iterations. */
size_t ks_newton_dense_workspace_bytes(int cls, int n, int m, int batch, int chunk);
int ks_newton_dense(int cls, int n, int m, int batch, double* modes, double* params, uint32_t cmask, double tol, double ftol,
                    int maxiter, double min_step, int chunk, int refine, int32_t* status_out, int32_t* nit_out,
                    double* cost_out, long long* counters_out, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KS_B200_H */
