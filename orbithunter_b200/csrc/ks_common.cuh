// Shared definitions for the KS hot-path kernels (class ids, op ids, argument block, small helpers).
#pragma once
#include <utility>

#include "ks_platform.cuh"

// class ids: the four symmetry classes named by the north star (reference orbithunter/ks/orbits.py:24, :2554,
// :3582, :3189).  Numbering is shared with include/ks_b200.h and oracle/ks_oracle.py.
enum { KS_ORBIT = 0, KS_RELATIVE = 1, KS_SHIFT_REFLECTION = 2, KS_ANTISYMMETRIC = 3 };
// fused evaluation ops
enum { KS_OP_EQN = 0, KS_OP_COSTGRAD = 1, KS_OP_MATVEC = 2, KS_OP_RMATVEC = 3, KS_OP_ADJSTEP = 5, KS_OP_ADJMULTI = 6 };
// constraint mask bits (set = parameter held constant; reference core.py:1617-1664 `constraints`)
enum { KS_FIX_T = 1, KS_FIX_X = 2, KS_FIX_S = 4 };

// State of the fused adjoint-descent pass (one launch = decide on the previous trial + evaluate the next one, per orbit).
// u/g/p/gp are ping-pong buffers: index cur[b] holds the accepted orbit of orbit b and the gradient there, the other
// index receives the trial point and its gradient (an accepted step is then just cur[b] ^= 1, no copy).
struct KsAdjFused {
  double* u[2];    // [B, Nm]
  double* g[2];    // [B, Nm]
  double* p[2];    // [B, 3]
  double* gp[2];   // [B, 3]
  double* cost;    // [B]  cost at u[cur]
  double* tcost;   // [B]  cost at the last trial
  double* step;    // [B]
  int* status;     // [B]
  int* nit;        // [B]
  int* live;       // [B]
  int* cur;        // [B]
  int* n_live;     // [1]
  double tol, ftol, min_step;
  int maxiter;
  int first;       // first pass: evaluate u itself (initial cost and gradient), nothing to decide
  int inner;       // KS_OP_ADJMULTI: passes each orbit runs inside one launch (state stays in shared memory / registers)
};

struct KsEvalArgs {
  const double* modes;    // [B, Nm]   u in the modes basis
  const double* params;   // [B, 3]    (T, L, S)
  const double* vmodes;   // [B, Nm]   v (matvec / rmatvec), else null
  const double* vparams;  // [B, 3]    (dT, dL, dS) increments (matvec), else null
  double* out_state;      // [B, Nm]   F (eqn) | grad (costgrad) | J v | J^T v
  double* out_params;     // [B, 3]    parameter components of grad / J^T v (0 where constrained), or null
  double* out_cost;       // [B]       1/2 |F|^2 (eqn, costgrad), or null
  double* f_out;          // [B, Nm]   costgrad only: also emit F; null -> per-warp scratch is used instead
  double* scratch;        // workspace (ks_workspace_bytes)
  int batch;
  unsigned cmask;
  int precond;            // costgrad: apply the reference's diagonal preconditioner (ks/orbits.py:1034-1089)
  // cdof-layout vectors for the Krylov drivers (core.py:1617-1643): leading dimensions in doubles (0 = Nm) and
  // whether the free parameters ride at the tail of the vector ([state.ravel(), free params in label order])
  int v_ld, out_ld;
  int v_tail;             // matvec: read (dT, dL, dS) from vmodes[b*v_ld + Nm + slot] instead of vparams
  int out_tail;           // rmatvec/costgrad: write free parameter components to out_state[b*out_ld + Nm + slot]
  int pingpong;           // 32x32 variant 6: the two warp groups of a CTA take turns on the fp64 pipe (ks_debug_warp32_pingpong)
  int stagger_ns;         // 32x32 kernels: start delay between the warps of one scheduler (tuning, ks_debug_warp32_stagger)
  // 32x32 matvec / rmatvec on the one-orbit-per-warp kernel, column modes of the dense Newton step:
  int u_div;              // > 0: evaluation e works on orbit e / u_div (u from modes / params [e / u_div], read in place)
  int unit_v;             // with u_div: v is the unit vector of cdof slot e % u_div, made in the kernel (nothing read)
                          //   Jacobian:  matvec,  u_div = N,  unit_v = 1           -> column e % N of J at out_state + e * out_ld
                          //   J J^T:     rmatvec, u_div = Nm, unit_v = 1 (w = J^T e_j), then matvec, u_div = Nm, v = w from memory
  // Operator cache of the Krylov drivers (64x64 CTA-resident kernel): during one inner solve u is fixed, so N(u, u) =
  // 1/2 Dx(u u) is computed once (by the eqn evaluation that opens the outer iteration, n_out) and reused by every matvec /
  // rmatvec of the solve (n_cache), which then skip the two forward transforms of u u.
  const double* n_cache;   // [B, Nm] tuples of N(u, u), or null
  double* n_out;           // eqn: also emit N(u, u), or null
  KsAdjFused adj;         // KS_OP_ADJSTEP / KS_OP_ADJMULTI only
};

namespace ks {
// slot of parameter i (0 t, 1 x, 2 s) among the free ones, or -1 if it is constrained
KS_HD int free_slot(unsigned cmask, int i) {
  if ((cmask >> i) & 1u) return -1;
  int slot = 0;
  for (int k = 0; k < i; ++k) slot += ((cmask >> k) & 1u) ? 0 : 1;
  return slot;
}
}  // namespace ks

namespace ks {
template <int B, int E, class F>
KS_DEVI void static_for(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(f);
  }
}
constexpr double kTwoPi = 6.283185307179586476925286766559;
constexpr double kSqrt2 = 1.4142135623730950488016887242097;
}  // namespace ks
