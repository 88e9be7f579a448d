// Batched Krylov solvers for the Newton steps of hunt(): one CTA per orbit, state machines advanced in lock step.
//
// Reference path replaced: orbithunter/optimize.py _scipy_sparse_linalg_solver_wrapper :1196-1404, which hands
// the operators of _sparse_linalg_factory :1682-1850 to scipy.sparse.linalg.gmres / lsqr (third-party, SciPy 1.18.1:
// scipy/sparse/linalg/_isolve/iterative.py:593-853 and _isolve/lsqr.py).  The recurrences below restate those
// published algorithms step for step (GMRES: left preconditioning, modified Gram-Schmidt, LAPACK-dlartg Givens,
// breakdown test h1 <= eps*h0, adaptive inner tolerance ptol; LSQR: Golub-Kahan bidiagonalisation with SciPy's
// _sym_ortho and stopping rules 1-7) so that iteration counts are comparable.  Each solver "step" consumes the result
// `w` of one operator application on `vin` (issued by the host driver between steps: A^T A for GMRES on the normal
// equations :1764-1797, A or A^T alternately for LSQR :1731-1759) and writes the next operator input.
// All reductions are fixed-order (deterministic).  Vectors use the reference's cdof layout [state.ravel(), free
// parameters in label order] (core.py:1617-1643).
#pragma once
#include "ks_common.cuh"

namespace kskry {

constexpr int kThreads = 256;
constexpr double kEps = 2.220446049250313e-16;
enum { ST_ARNOLDI = 0, ST_RESID = 1, ST_DONE = 2 };
enum { SC_BNRM2 = 0, SC_ATOL, SC_PTOL, SC_PFAC, SC_PRESID, SC_RNORM, SC_COUNT = 8 };
enum { IC_STATE = 0, IC_COL, IC_ITER, IC_BREAK, IC_INFO, IC_INNER, IC_COUNT = 8 };

// fixed-order block sum; `sh` has kThreads/32 doubles.  All threads get the result.
KS_DEVI double block_sum(double v, double* sh) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) t += sh[i];
  return t;
}

struct Diag {  // OrbitKS.preconditioner (ks/orbits.py:1091-1141): diagonal of M in the cdof layout
  int n, m, cols, h, k, nm;
  unsigned cmask;
  double wbase, qbase, pt, px;
  KS_DEVI double at(int i) const {
    if (i < nm) {
      const int r = i / cols, c = i - r * cols;
      const int j = r <= h ? r : r - h, kp = c < k ? c : c - k;
      const double w = wbase * (j * (1.0 / n)), q = qbase * ((kp + 1) * (1.0 / m)), q2 = q * q;
      return 1.0 / ((w < 0 ? -w : w) + q2 + q2 * q2);
    }
    int slot = i - nm;  // free parameters in label order t, x, s
    if (!(cmask & KS_FIX_T)) { if (slot == 0) return pt; --slot; }
    if (!(cmask & KS_FIX_X)) { if (slot == 0) return px; --slot; }
    return 1.0;
  }
};
KS_DEVI Diag make_diag(int n, int m, int cols, unsigned cmask, double T, double L) {
  Diag d;
  d.n = n; d.m = m; d.cols = cols; d.h = n / 2 - 1; d.k = m / 2 - 1; d.nm = (n - 1) * cols; d.cmask = cmask;
  d.wbase = -1.0 * (ks::kTwoPi * n / T);
  d.qbase = ks::kTwoPi * m / L;
  d.pt = 1.0 / T;
  d.px = 1.0 / ((L * L) * (L * L));
  return d;
}

struct GmresState {
  double* V;        // [B][R+1][N]  Krylov basis
  double* H;        // [B][R][R+1]  h[col][k] as in SciPy (row = column index of the Hessenberg matrix)
  double* G;        // [B][R][2]    Givens (c, s)
  double* S;        // [B][R+1]
  double* x;        // [B][N]
  const double* b;  // [B][N]       right-hand side (A^T b for the normal equations)
  double* vin;      // [B][N]       next operator input
  const double* w;  // [B][N]       last operator output
  double* sc;       // [B][SC_COUNT]
  int* ic;          // [B][IC_COUNT]
  const double* params;  // [B][3]
  const int* live;  // [B] Newton-loop mask (orbits with 0 are skipped), may be null
  int* n_active;    // [1]
  int B, N, R, maxiter, n, m, cols, precond;
  unsigned cmask;
  double rtol, atol_in;
};

// LAPACK dlartg (3.10+): c >= 0, r carries the sign of f
KS_DEVI void lartg(double f, double g, double& c, double& s, double& r) {
  if (g == 0.0) {
    c = 1.0; s = 0.0; r = f;
  } else if (f == 0.0) {
    c = 0.0; s = g < 0 ? -1.0 : 1.0; r = g < 0 ? -g : g;
  } else {
    const double d = sqrt(f * f + g * g);
    c = (f < 0 ? -f : f) / d;
    r = f < 0 ? -d : d;
    s = g / r;
  }
}

// v0 = M r / |M r|, S = [|M r|, 0...], col = 0, vin = v0                      (iterative.py:741-748)
KS_DEVI void gmres_begin_cycle(const GmresState& g, int b, const double* r_unprec_minus, bool r_is_b, const Diag& dg, double* sh) {
  double* V0 = g.V + (size_t)b * (g.R + 1) * g.N;
  const double* bb = g.b + (size_t)b * g.N;
  const double* ww = g.w + (size_t)b * g.N;
  double acc = 0.0;
  for (int i = threadIdx.x; i < g.N; i += kThreads) {
    double r = r_is_b ? bb[i] : bb[i] - ww[i];
    if (g.precond) r *= dg.at(i);
    V0[i] = r;
    acc += r * r;
  }
  const double tmp = sqrt(block_sum(acc, sh));
  const double inv = 1.0 / tmp;
  double* vin = g.vin + (size_t)b * g.N;
  for (int i = threadIdx.x; i < g.N; i += kThreads) {
    const double v = V0[i] * inv;
    V0[i] = v;
    vin[i] = v;
  }
  if (threadIdx.x == 0) {
    double* S = g.S + (size_t)b * (g.R + 1);
    for (int i = 0; i <= g.R; ++i) S[i] = 0.0;
    S[0] = tmp;
    g.ic[b * IC_COUNT + IC_COL] = 0;
    g.ic[b * IC_COUNT + IC_BREAK] = 0;
    g.ic[b * IC_COUNT + IC_STATE] = ST_ARNOLDI;
  }
  (void)r_unprec_minus;
}

__global__ void __launch_bounds__(kThreads) gmres_start_kernel(const GmresState g) {
  __shared__ double sh[kThreads / 32];
  const int b = blockIdx.x;
  int* ic = g.ic + b * IC_COUNT;
  double* sc = g.sc + b * SC_COUNT;
  double* x = g.x + (size_t)b * g.N;
  for (int i = threadIdx.x; i < g.N; i += kThreads) x[i] = 0.0;
  if (g.live && !g.live[b]) {
    if (threadIdx.x == 0) ic[IC_STATE] = ST_DONE;
    return;
  }
  const double* bb = g.b + (size_t)b * g.N;
  const Diag dg = make_diag(g.n, g.m, g.cols, g.cmask, g.params[b * 3 + 0], g.params[b * 3 + 1]);
  double a0 = 0.0, a1 = 0.0;
  for (int i = threadIdx.x; i < g.N; i += kThreads) {
    const double v = bb[i], pv = g.precond ? v * dg.at(i) : v;
    a0 += v * v;
    a1 += pv * pv;
  }
  const double bnrm2 = sqrt(block_sum(a0, sh));
  const double mb = sqrt(block_sum(a1, sh));
  const double atol = fmax(g.atol_in, g.rtol * bnrm2);                                   // _get_atol_rtol
  if (threadIdx.x == 0) {
    sc[SC_BNRM2] = bnrm2;
    sc[SC_ATOL] = atol;
    sc[SC_PFAC] = 1.0;
    sc[SC_PTOL] = mb * fmin(1.0, atol / bnrm2);                                          // iterative.py:727-729
    sc[SC_PRESID] = 0.0;
    sc[SC_RNORM] = bnrm2;
    ic[IC_ITER] = 0;
    ic[IC_INFO] = 0;
    ic[IC_INNER] = 0;
  }
  // b == 0 -> x = b; |r| < atol at iteration 0 -> x = 0 is returned                     (:716, :746)
  if (bnrm2 == 0.0 || bnrm2 < atol || g.maxiter <= 0) {
    if (threadIdx.x == 0) ic[IC_STATE] = ST_DONE;
    return;
  }
  gmres_begin_cycle(g, b, nullptr, true, dg, sh);
  if (threadIdx.x == 0) atomicAdd(g.n_active, 1);
}

__global__ void __launch_bounds__(kThreads) gmres_step_kernel(const GmresState g) {
  __shared__ double sh[kThreads / 32];
  __shared__ int sh_i[2];
  const int b = blockIdx.x;
  int* ic = g.ic + b * IC_COUNT;
  const int state = ic[IC_STATE];
  if (state == ST_DONE) return;
  __syncthreads();  // every thread has read the state before thread 0 may change it
  double* sc = g.sc + b * SC_COUNT;
  const Diag dg = make_diag(g.n, g.m, g.cols, g.cmask, g.params[b * 3 + 0], g.params[b * 3 + 1]);
  double* V = g.V + (size_t)b * (g.R + 1) * g.N;
  double* H = g.H + (size_t)b * g.R * (g.R + 1);
  double* Gv = g.G + (size_t)b * g.R * 2;
  double* S = g.S + (size_t)b * (g.R + 1);
  double* x = g.x + (size_t)b * g.N;
  double* vin = g.vin + (size_t)b * g.N;
  const double* ww = g.w + (size_t)b * g.N;
  if (state == ST_ARNOLDI) {
    const int col = ic[IC_COL];
    double* wv = V + (size_t)(col + 1) * g.N;  // w lives in the next basis slot
    double acc = 0.0;
    for (int i = threadIdx.x; i < g.N; i += kThreads) {
      double v = ww[i];
      if (g.precond) v *= dg.at(i);
      wv[i] = v;
      acc += v * v;
    }
    const double h0 = sqrt(block_sum(acc, sh));
    for (int k = 0; k <= col; ++k) {                                                      // modified Gram-Schmidt (:767-772)
      const double* vk = V + (size_t)k * g.N;
      double d = 0.0;
      for (int i = threadIdx.x; i < g.N; i += kThreads) d += vk[i] * wv[i];
      d = block_sum(d, sh);
      if (threadIdx.x == 0) H[col * (g.R + 1) + k] = d;
      for (int i = threadIdx.x; i < g.N; i += kThreads) wv[i] -= d * vk[i];
    }
    acc = 0.0;
    for (int i = threadIdx.x; i < g.N; i += kThreads) acc += wv[i] * wv[i];
    const double h1 = sqrt(block_sum(acc, sh));
    const bool breakdown = h1 <= kEps * h0;                                               // (:777-781)
    if (!breakdown) {
      const double inv = 1.0 / h1;
      for (int i = threadIdx.x; i < g.N; i += kThreads) wv[i] *= inv;
    }
    if (threadIdx.x == 0) {
      double* hc = H + col * (g.R + 1);
      hc[col + 1] = breakdown ? 0.0 : h1;
      for (int k = 0; k < col; ++k) {                                                     // past rotations (:784-787)
        const double c = Gv[2 * k], s = Gv[2 * k + 1], n0 = hc[k], n1 = hc[k + 1];
        hc[k] = c * n0 + s * n1;
        hc[k + 1] = -s * n0 + c * n1;
      }
      double c, s, mag;
      lartg(hc[col], hc[col + 1], c, s, mag);                                             // (:790-792)
      Gv[2 * col] = c;
      Gv[2 * col + 1] = s;
      hc[col] = mag;
      hc[col + 1] = 0.0;
      const double tmp = -s * S[col];
      S[col] = c * S[col];
      S[col + 1] = tmp;
      const double presid = tmp < 0 ? -tmp : tmp;
      sc[SC_PRESID] = presid;
      ic[IC_INNER] += 1;
      ic[IC_BREAK] = breakdown ? 1 : 0;
      const bool end_cycle = (presid <= sc[SC_PTOL]) || breakdown || (col + 1 == g.R);    // (:805, loop bound)
      sh_i[0] = end_cycle ? 1 : 0;
      if (end_cycle) {
        // y = trsv(h[:col+1,:col+1].T, S[:col+1]) with pseudo-solve of singular pivots      (:812-824); y overwrites S
        if (hc[col] == 0.0) S[col] = 0.0;
        for (int k = col; k > 0; --k) {
          if (S[k] != 0.0) {
            S[k] /= H[k * (g.R + 1) + k];
            const double t = S[k];
            for (int i = 0; i < k; ++i) S[i] -= t * H[k * (g.R + 1) + i];
          }
        }
        if (S[0] != 0.0) S[0] /= H[0];
        ic[IC_STATE] = ST_RESID;
      } else {
        ic[IC_COL] = col + 1;
      }
    }
    __syncthreads();
    if (sh_i[0]) {
      for (int i = threadIdx.x; i < g.N; i += kThreads) {                                 // x += y @ v[:col+1]  (:826)
        double xi = x[i];
        for (int k = 0; k <= col; ++k) xi += S[k] * V[(size_t)k * g.N + i];
        x[i] = xi;
        vin[i] = xi;                                                                      // next: r = b - A x
      }
    } else {
      for (int i = threadIdx.x; i < g.N; i += kThreads) vin[i] = wv[i];
    }
    if (threadIdx.x == 0) atomicAdd(g.n_active, 1);
  } else {  // ST_RESID: w = A x                                                            (:828-851)
    const double* bb = g.b + (size_t)b * g.N;
    double acc = 0.0;
    for (int i = threadIdx.x; i < g.N; i += kThreads) {
      const double r = bb[i] - ww[i];
      acc += r * r;
    }
    const double rnorm = sqrt(block_sum(acc, sh));
    if (threadIdx.x == 0) {
      sc[SC_RNORM] = rnorm;
      const int iter = ic[IC_ITER] + 1;
      ic[IC_ITER] = iter;
      const double atol = sc[SC_ATOL], presid = sc[SC_PRESID];
      int done = 0;
      if (rnorm <= atol) {
        done = 1;
      } else if (ic[IC_BREAK]) {
        done = 1;
      } else {
        double f = sc[SC_PFAC];
        f = (presid <= sc[SC_PTOL]) ? fmax(kEps, 0.25 * f) : fmin(1.0, 1.5 * f);
        sc[SC_PFAC] = f;
        sc[SC_PTOL] = presid * fmin(f, atol / rnorm);
        if (iter >= g.maxiter) done = 1;
      }
      if (done) {
        ic[IC_INFO] = (rnorm <= atol) ? 0 : g.maxiter;
        ic[IC_STATE] = ST_DONE;
      }
      sh_i[1] = done;
    }
    __syncthreads();
    if (!sh_i[1]) {
      gmres_begin_cycle(g, b, nullptr, false, dg, sh);
      if (threadIdx.x == 0) atomicAdd(g.n_active, 1);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GMRES step, thread-block-cluster version: the C CTAs of one cluster share ONE orbit, each owning a slice of at most
// kSlice elements of every length-N vector.  The vector being orthogonalised stays in registers for the whole modified
// Gram-Schmidt sweep (the basis is the only HBM traffic, read exactly once, next vector prefetched during the
// reduction), dot products are block sums combined across the cluster through distributed shared memory in rank order
// (deterministic), and rank 0 runs the small Hessenberg / Givens update.  Same recurrences as gmres_step_kernel.
// All shared memory is dynamic (the CPU emulation runs the blocks of a cluster concurrently).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSliceRegs = 32;
constexpr int kSlice = kThreads * kSliceRegs;  // 8192 elements per CTA (clusters); short vectors use 8 or 16 registers per thread
constexpr int kMaxCluster = 8;
constexpr int kClusterSmemBytes = (kThreads / 32 + 2 * 2 + 2) * 8 + 16;

struct ClusterRed {
  double* sh;     // [kThreads/32]
  double* part;   // [2]  this block's partial, alternating slots
  double* bc;     // [2]  block-wide broadcast of the cluster total / of rank 0's flag
  int phase;
  unsigned C, rank;
};
// sum over the whole cluster; every thread of every block gets the same value
KS_DEVI double cluster_sum(double v, ClusterRed& cr) {
  const double t = block_sum(v, cr.sh);
  if (cr.C == 1) return t;
  if (threadIdx.x == 0) cr.part[cr.phase] = t;
  ks_cluster_sync();
  if (threadIdx.x < 32) {
    double r = threadIdx.x < cr.C ? *ks_cluster_map(cr.part + cr.phase, threadIdx.x) : 0.0;
    double tot = 0.0;
    for (unsigned k = 0; k < cr.C; ++k) tot += __shfl_sync(0xffffffffu, r, (int)k);  // rank order
    if (threadIdx.x == 0) cr.bc[0] = tot;
  }
  __syncthreads();
  const double tot = cr.bc[0];
  __syncthreads();
  cr.phase ^= 1;
  return tot;
}

template <int SR>  // vector elements per thread: 32 (clusters, N > 8192), 16 (N <= 4096) or 8 (N <= 2048) for one-CTA "clusters"
__global__ void __launch_bounds__(kThreads) gmres_step_cluster_kernel(const GmresState g) {
  constexpr int kSliceRegs = SR;
  KS_DYN_SMEM(double, dsm);
  ClusterRed cr;
  cr.sh = dsm;
  cr.part = dsm + kThreads / 32;
  cr.bc = cr.part + 2;
  double* flag = cr.bc + 2;  // [2] rank 0: end_cycle / done, read by the other ranks through DSMEM
  cr.phase = 0;
  cr.C = ks_cluster_size();
  cr.rank = ks_cluster_rank();
  const int b = blockIdx.x / cr.C;
  int* ic = g.ic + b * IC_COUNT;
  const int state = ic[IC_STATE];
  if (state == ST_DONE) return;  // uniform over the cluster
  ks_cluster_sync();             // every block has read the state before rank 0 may change it
  const bool lead = cr.rank == 0 && threadIdx.x == 0;
  double* sc = g.sc + b * SC_COUNT;
  const Diag dg = make_diag(g.n, g.m, g.cols, g.cmask, g.params[b * 3 + 0], g.params[b * 3 + 1]);
  double* V = g.V + (size_t)b * (g.R + 1) * g.N;
  double* H = g.H + (size_t)b * g.R * (g.R + 1);
  double* Gv = g.G + (size_t)b * g.R * 2;
  double* S = g.S + (size_t)b * (g.R + 1);
  double* x = g.x + (size_t)b * g.N;
  double* vin = g.vin + (size_t)b * g.N;
  const double* ww = g.w + (size_t)b * g.N;
  // this block's slice: elements lo + threadIdx.x + kThreads * r, r < kSliceRegs
  const int per = (g.N + (int)cr.C - 1) / (int)cr.C;
  const int lo = (int)cr.rank * per, hi = (lo + per < g.N) ? lo + per : g.N;
  const int i0 = lo + (int)threadIdx.x;
  auto rank0_flag = [&](int slot) -> bool {  // after a cluster sync: rank 0's flag[slot], same value in every thread
    if (threadIdx.x == 0) cr.bc[1] = *ks_cluster_map(flag + slot, 0);
    __syncthreads();
    const bool f = cr.bc[1] != 0.0;
    __syncthreads();
    return f;
  };
  auto begin_cycle = [&](bool r_is_b) {  // gmres_begin_cycle on slices                       (iterative.py:741-748)
    const double* bb = g.b + (size_t)b * g.N;
    double r[kSliceRegs];
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < kSliceRegs; ++q) {
      const int i = i0 + q * kThreads;
      r[q] = 0.0;
      if (i < hi) {
        r[q] = r_is_b ? bb[i] : bb[i] - ww[i];
        if (g.precond) r[q] *= dg.at(i);
      }
      acc += r[q] * r[q];
    }
    const double tmp = sqrt(cluster_sum(acc, cr));
    const double inv = 1.0 / tmp;
#pragma unroll
    for (int q = 0; q < kSliceRegs; ++q) {
      const int i = i0 + q * kThreads;
      if (i < hi) {
        const double v = r[q] * inv;
        V[i] = v;
        vin[i] = v;
      }
    }
    if (lead) {
      for (int i = 0; i <= g.R; ++i) S[i] = 0.0;
      S[0] = tmp;
      ic[IC_COL] = 0;
      ic[IC_BREAK] = 0;
      ic[IC_STATE] = ST_ARNOLDI;
    }
  };
  if (state == ST_ARNOLDI) {
    const int col = ic[IC_COL];
    double w[kSliceRegs], vc[kSliceRegs], vn[kSliceRegs];
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < kSliceRegs; ++q) {
      const int i = i0 + q * kThreads;
      w[q] = 0.0;
      vc[q] = 0.0;
      if (i < hi) {
        w[q] = g.precond ? ww[i] * dg.at(i) : ww[i];
        vc[q] = V[i];
      }
      acc += w[q] * w[q];
    }
    const double h0 = sqrt(cluster_sum(acc, cr));
    for (int k = 0; k <= col; ++k) {                                                      // modified Gram-Schmidt (:767-772)
      if (k < col) {
        const double* vk1 = V + (size_t)(k + 1) * g.N;
#pragma unroll
        for (int q = 0; q < kSliceRegs; ++q) {
          const int i = i0 + q * kThreads;
          vn[q] = i < hi ? vk1[i] : 0.0;
        }
      }
      double d = 0.0;
#pragma unroll
      for (int q = 0; q < kSliceRegs; ++q) d += vc[q] * w[q];
      d = cluster_sum(d, cr);
      if (lead) H[col * (g.R + 1) + k] = d;
#pragma unroll
      for (int q = 0; q < kSliceRegs; ++q) {
        w[q] -= d * vc[q];
        vc[q] = vn[q];
      }
    }
    acc = 0.0;
#pragma unroll
    for (int q = 0; q < kSliceRegs; ++q) acc += w[q] * w[q];
    const double h1 = sqrt(cluster_sum(acc, cr));
    const bool breakdown = h1 <= kEps * h0;                                               // (:777-781)
    double* wv = V + (size_t)(col + 1) * g.N;
    {
      const double inv = breakdown ? 1.0 : 1.0 / h1;
#pragma unroll
      for (int q = 0; q < kSliceRegs; ++q) {
        const int i = i0 + q * kThreads;
        w[q] *= inv;
        if (i < hi) wv[i] = w[q];
      }
    }
    if (lead) {
      double* hc = H + col * (g.R + 1);
      hc[col + 1] = breakdown ? 0.0 : h1;
      for (int k = 0; k < col; ++k) {                                                     // past rotations (:784-787)
        const double c = Gv[2 * k], s = Gv[2 * k + 1], n0 = hc[k], n1 = hc[k + 1];
        hc[k] = c * n0 + s * n1;
        hc[k + 1] = -s * n0 + c * n1;
      }
      double c, s, mag;
      lartg(hc[col], hc[col + 1], c, s, mag);                                             // (:790-792)
      Gv[2 * col] = c;
      Gv[2 * col + 1] = s;
      hc[col] = mag;
      hc[col + 1] = 0.0;
      const double tmp = -s * S[col];
      S[col] = c * S[col];
      S[col + 1] = tmp;
      const double presid = tmp < 0 ? -tmp : tmp;
      sc[SC_PRESID] = presid;
      ic[IC_INNER] += 1;
      ic[IC_BREAK] = breakdown ? 1 : 0;
      const bool end_cycle = (presid <= sc[SC_PTOL]) || breakdown || (col + 1 == g.R);    // (:805, loop bound)
      flag[0] = end_cycle ? 1.0 : 0.0;
      if (end_cycle) {
        if (hc[col] == 0.0) S[col] = 0.0;                                                 // (:812-824); y overwrites S
        for (int k = col; k > 0; --k) {
          if (S[k] != 0.0) {
            S[k] /= H[k * (g.R + 1) + k];
            const double t = S[k];
            for (int i = 0; i < k; ++i) S[i] -= t * H[k * (g.R + 1) + i];
          }
        }
        if (S[0] != 0.0) S[0] /= H[0];
        ic[IC_STATE] = ST_RESID;
      } else {
        ic[IC_COL] = col + 1;
      }
      __threadfence();
    }
    ks_cluster_sync();  // rank 0's flag (shared memory) and S (global) are visible to the whole cluster
    if (rank0_flag(0)) {
#pragma unroll 4
      for (int q = 0; q < kSliceRegs; ++q) {                                              // x += y @ v[:col+1]  (:826)
        const int i = i0 + q * kThreads;
        if (i < hi) {
          double xi = x[i];
          for (int k = 0; k <= col; ++k) xi += S[k] * V[(size_t)k * g.N + i];
          x[i] = xi;
          vin[i] = xi;                                                                    // next: r = b - A x
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < kSliceRegs; ++q) {
        const int i = i0 + q * kThreads;
        if (i < hi) vin[i] = w[q];
      }
    }
    if (lead) atomicAdd(g.n_active, 1);
  } else {  // ST_RESID: w = A x                                                            (:828-851)
    const double* bb = g.b + (size_t)b * g.N;
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < kSliceRegs; ++q) {
      const int i = i0 + q * kThreads;
      const double r = i < hi ? bb[i] - ww[i] : 0.0;
      acc += r * r;
    }
    const double rnorm = sqrt(cluster_sum(acc, cr));
    if (lead) {
      sc[SC_RNORM] = rnorm;
      const int iter = ic[IC_ITER] + 1;
      ic[IC_ITER] = iter;
      const double atol = sc[SC_ATOL], presid = sc[SC_PRESID];
      int done = 0;
      if (rnorm <= atol) {
        done = 1;
      } else if (ic[IC_BREAK]) {
        done = 1;
      } else {
        double f = sc[SC_PFAC];
        f = (presid <= sc[SC_PTOL]) ? fmax(kEps, 0.25 * f) : fmin(1.0, 1.5 * f);
        sc[SC_PFAC] = f;
        sc[SC_PTOL] = presid * fmin(f, atol / rnorm);
        if (iter >= g.maxiter) done = 1;
      }
      if (done) {
        ic[IC_INFO] = (rnorm <= atol) ? 0 : g.maxiter;
        ic[IC_STATE] = ST_DONE;
      }
      flag[1] = done ? 1.0 : 0.0;
    }
    ks_cluster_sync();
    if (!rank0_flag(1)) {
      begin_cycle(false);
      if (lead) atomicAdd(g.n_active, 1);
    }
  }
  ks_cluster_sync();  // no block may exit while a peer can still read its shared memory
}

// ---------------------------------------------------------------------------------------------------------------
// LSQR (SciPy _isolve/lsqr.py, damp = 0, x0 = None).  Phases: 0 = waiting for A v, 1 = waiting for A^T u.
// ---------------------------------------------------------------------------------------------------------------
enum { LQ_ALFA = 0, LQ_BETA, LQ_RHOBAR, LQ_PHIBAR, LQ_ANORM, LQ_DDNORM, LQ_XXNORM, LQ_Z, LQ_CS2, LQ_SN2, LQ_BNORM, LQ_RES2,
       // LSMR only (SciPy _isolve/lsmr.py): the names of its scalar recurrences
       LM_ZETABAR, LM_ALPHABAR, LM_RHO, LM_CBAR, LM_SBAR, LM_BETADD, LM_BETAD, LM_RHODOLD, LM_TAUTILDEOLD, LM_THETATILDE, LM_ZETA,
       LM_D, LM_NORMA2, LM_MAXRBAR, LM_MINRBAR, LQ_COUNT = 32 };
enum { LI_PHASE = 0, LI_ITN, LI_ISTOP, LI_DONE, LI_COUNT = 8 };

struct LsqrState {
  double* u;        // [B][M]   left vectors (length Nm)
  double* v;        // [B][N]   right vectors (cdof)
  double* wv;       // [B][N]   LSQR: w; LSMR: h
  double* hbar;     // [B][N]   LSMR only
  double* x;        // [B][N]
  const double* b;  // [B][M]   right-hand side (-F)
  double* vin;      // [B][max(M,N)]  next operator input
  const double* w;  // [B][max(M,N)]  last operator output
  double* sc;       // [B][LQ_COUNT]
  int* ic;          // [B][LI_COUNT]
  const int* live;
  int* n_active;
  int B, M, N, ld, iter_lim;
  int lsmr;         // 0: LSQR recurrences, 1: LSMR recurrences on the same Golub-Kahan bidiagonalisation
  double atol, btol, conlim;
};

KS_DEVI void sym_ortho(double a, double b, double& c, double& s, double& r) {
  const double aa = a < 0 ? -a : a, ab = b < 0 ? -b : b;
  if (b == 0.0) {
    c = (a > 0) - (a < 0); s = 0.0; r = aa;
  } else if (a == 0.0) {
    c = 0.0; s = (b > 0) - (b < 0); r = ab;
  } else if (ab > aa) {
    const double tau = a / b;
    s = ((b > 0) - (b < 0)) / sqrt(1.0 + tau * tau);
    c = s * tau;
    r = b / s;
  } else {
    const double tau = b / a;
    c = ((a > 0) - (a < 0)) / sqrt(1.0 + tau * tau);
    s = c * tau;
    r = a / c;
  }
}

// u = b / |b|; vin = u (operator: A^T).  The first A^T u result is consumed by lsqr_step in phase 1 with itn == 0.
__global__ void __launch_bounds__(kThreads) lsqr_start_kernel(const LsqrState g) {
  __shared__ double sh[kThreads / 32];
  const int b = blockIdx.x;
  int* ic = g.ic + b * LI_COUNT;
  double* sc = g.sc + b * LQ_COUNT;
  double* x = g.x + (size_t)b * g.N;
  for (int i = threadIdx.x; i < g.N; i += kThreads) x[i] = 0.0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < LQ_COUNT; ++i) sc[i] = 0.0;
    sc[LQ_CS2] = -1.0;
    ic[LI_ITN] = 0;
    ic[LI_ISTOP] = 0;
    ic[LI_PHASE] = 1;
    ic[LI_DONE] = 0;
  }
  if (g.live && !g.live[b]) {
    if (threadIdx.x == 0) ic[LI_DONE] = 1;
    return;
  }
  const double* bb = g.b + (size_t)b * g.M;
  double acc = 0.0;
  for (int i = threadIdx.x; i < g.M; i += kThreads) acc += bb[i] * bb[i];
  const double beta = sqrt(block_sum(acc, sh));
  if (threadIdx.x == 0) {
    sc[LQ_BNORM] = beta;
    sc[LQ_BETA] = beta;
  }
  if (!(beta > 0.0)) {  // arnorm == 0: x = 0 is returned                                   (lsqr.py "if arnorm == 0")
    if (threadIdx.x == 0) ic[LI_DONE] = 1;
    return;
  }
  double* u = g.u + (size_t)b * g.M;
  double* vin = g.vin + (size_t)b * g.ld;
  const double inv = 1.0 / beta;
  for (int i = threadIdx.x; i < g.M; i += kThreads) {
    const double t = bb[i] * inv;
    u[i] = t;
    vin[i] = t;
  }
  if (threadIdx.x == 0) atomicAdd(g.n_active, 1);
}

__global__ void __launch_bounds__(kThreads) lsqr_step_kernel(const LsqrState g) {
  __shared__ double sh[kThreads / 32];
  __shared__ double shd[6];
  const int b = blockIdx.x;
  int* ic = g.ic + b * LI_COUNT;
  if (ic[LI_DONE]) return;
  const int phase = ic[LI_PHASE], itn = ic[LI_ITN];
  __syncthreads();
  double* sc = g.sc + b * LQ_COUNT;
  double* u = g.u + (size_t)b * g.M;
  double* v = g.v + (size_t)b * g.N;
  double* wv = g.wv + (size_t)b * g.N;
  double* x = g.x + (size_t)b * g.N;
  double* vin = g.vin + (size_t)b * g.ld;
  const double* ww = g.w + (size_t)b * g.ld;
  if (phase == 0) {
    // u = A v - alfa u; beta = |u|; u /= beta                                                (main loop, first half)
    const double alfa = sc[LQ_ALFA];
    double acc = 0.0;
    for (int i = threadIdx.x; i < g.M; i += kThreads) {
      const double t = ww[i] - alfa * u[i];
      u[i] = t;
      acc += t * t;
    }
    const double beta = sqrt(block_sum(acc, sh));
    if (beta > 0.0) {
      const double inv = 1.0 / beta;
      for (int i = threadIdx.x; i < g.M; i += kThreads) {
        const double t = u[i] * inv;
        u[i] = t;
        vin[i] = t;
      }
    } else {
      for (int i = threadIdx.x; i < g.M; i += kThreads) vin[i] = 0.0;  // A^T 0 = 0: v, alfa keep their values below
    }
    if (threadIdx.x == 0) {
      sc[LQ_BETA] = beta;
      if (beta > 0.0 && !g.lsmr) {
        const double an = sc[LQ_ANORM];
        sc[LQ_ANORM] = sqrt(an * an + alfa * alfa + beta * beta);
      }
      ic[LI_ITN] = itn + 1;
      ic[LI_PHASE] = 1;
      atomicAdd(g.n_active, 1);
    }
    return;
  }
  // phase 1: v = A^T u - beta v; alfa = |v|; v /= alfa; then the scalar recurrences of this iteration
  const double beta = sc[LQ_BETA];
  double alfa = sc[LQ_ALFA];
  if (itn == 0 || beta > 0.0) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < g.N; i += kThreads) {
      const double t = (itn == 0) ? ww[i] : ww[i] - beta * v[i];
      v[i] = t;
      acc += t * t;
    }
    alfa = sqrt(block_sum(acc, sh));
    if (alfa > 0.0) {
      const double inv = 1.0 / alfa;
      for (int i = threadIdx.x; i < g.N; i += kThreads) v[i] *= inv;
    }
  }
  if (itn == 0) {
    // set-up: w = v (LSMR: h = v, hbar = 0), rhobar = alfa, phibar = beta; arnorm = alfa*beta == 0 -> finished with x = 0
    for (int i = threadIdx.x; i < g.N; i += kThreads) {
      wv[i] = v[i];
      vin[i] = v[i];
      if (g.lsmr) g.hbar[(size_t)b * g.N + i] = 0.0;
    }
    if (threadIdx.x == 0) {
      sc[LQ_ALFA] = alfa;
      sc[LQ_RHOBAR] = alfa;
      sc[LQ_PHIBAR] = beta;
      if (g.lsmr) {  // lsmr.py "Initialize variables for 1st iteration / for estimation of ||r|| / of ||A|| and cond(A)"
        sc[LM_ZETABAR] = alfa * beta; sc[LM_ALPHABAR] = alfa; sc[LM_RHO] = 1.0; sc[LQ_RHOBAR] = 1.0; sc[LM_CBAR] = 1.0;
        sc[LM_SBAR] = 0.0; sc[LM_BETADD] = beta; sc[LM_BETAD] = 0.0; sc[LM_RHODOLD] = 1.0; sc[LM_TAUTILDEOLD] = 0.0;
        sc[LM_THETATILDE] = 0.0; sc[LM_ZETA] = 0.0; sc[LM_D] = 0.0; sc[LM_NORMA2] = alfa * alfa; sc[LM_MAXRBAR] = 0.0;
        sc[LM_MINRBAR] = 1e100;
      }
      ic[LI_PHASE] = 0;
      if (alfa * beta == 0.0 || g.iter_lim <= 0) {
        ic[LI_DONE] = 1;
      } else {
        atomicAdd(g.n_active, 1);
      }
    }
    return;
  }
  if (g.lsmr) {
    // ---- LSMR (lsmr.py main loop after the bidiagonalisation step; damp = 0, so Qhat is the identity: chat = 1, shat = 0) ----
    if (threadIdx.x == 0) {
      const double alphabar0 = sc[LM_ALPHABAR];
      double chat, shat, alphahat;
      sym_ortho(alphabar0, 0.0, chat, shat, alphahat);
      const double rhoold = sc[LM_RHO];
      double c_, s_, rho;
      sym_ortho(alphahat, beta, c_, s_, rho);
      const double thetanew = s_ * alfa;
      sc[LM_ALPHABAR] = c_ * alfa;
      const double rhobarold = sc[LQ_RHOBAR], zetaold = sc[LM_ZETA];
      const double thetabar = sc[LM_SBAR] * rho, rhotemp = sc[LM_CBAR] * rho;
      double cbar, sbar, rhobar;
      sym_ortho(sc[LM_CBAR] * rho, thetanew, cbar, sbar, rhobar);
      const double zeta = cbar * sc[LM_ZETABAR];
      sc[LM_ZETABAR] = -sbar * sc[LM_ZETABAR];
      sc[LM_RHO] = rho; sc[LQ_RHOBAR] = rhobar; sc[LM_CBAR] = cbar; sc[LM_SBAR] = sbar; sc[LM_ZETA] = zeta;
      shd[0] = -(thetabar * rho / (rhoold * rhobarold));  // hbar scale
      shd[1] = zeta / (rho * rhobar);                      // x step along hbar
      shd[2] = -(thetanew / rho);                          // h scale
      // estimate of |r|
      const double betaacute = chat * sc[LM_BETADD], betacheck = -shat * sc[LM_BETADD];
      const double betahat = c_ * betaacute;
      sc[LM_BETADD] = -s_ * betaacute;
      const double thetatildeold = sc[LM_THETATILDE];
      double ctildeold, stildeold, rhotildeold;
      sym_ortho(sc[LM_RHODOLD], thetabar, ctildeold, stildeold, rhotildeold);
      sc[LM_THETATILDE] = stildeold * rhobar;
      sc[LM_RHODOLD] = ctildeold * rhobar;
      sc[LM_BETAD] = -stildeold * sc[LM_BETAD] + ctildeold * betahat;
      sc[LM_TAUTILDEOLD] = (zetaold - thetatildeold * sc[LM_TAUTILDEOLD]) / rhotildeold;
      const double taud = (zeta - sc[LM_THETATILDE] * sc[LM_TAUTILDEOLD]) / sc[LM_RHODOLD];
      sc[LM_D] += betacheck * betacheck;
      const double dd = sc[LM_BETAD] - taud;
      shd[3] = sqrt(sc[LM_D] + dd * dd + sc[LM_BETADD] * sc[LM_BETADD]);  // normr
      // estimates of |A| and cond(A)
      double na2 = sc[LM_NORMA2] + beta * beta;
      shd[4] = sqrt(na2);                                                   // normA
      sc[LM_NORMA2] = na2 + alfa * alfa;
      sc[LM_MAXRBAR] = fmax(sc[LM_MAXRBAR], rhobarold);
      if (itn > 1) sc[LM_MINRBAR] = fmin(sc[LM_MINRBAR], rhobarold);
      shd[5] = fmax(sc[LM_MAXRBAR], rhotemp) / fmin(sc[LM_MINRBAR], rhotemp);  // condA
      sc[LQ_ALFA] = alfa;
    }
    __syncthreads();
    const double sh_hbar = shd[0], sx = shd[1], sh_h = shd[2];
    double* hb = g.hbar + (size_t)b * g.N;
    double acc = 0.0;
    for (int i = threadIdx.x; i < g.N; i += kThreads) {
      const double hbn = fma(sh_hbar, hb[i], wv[i]);  // hbar = h + scale * hbar   (hbar *= scale; hbar += h)
      hb[i] = hbn;
      const double xn = x[i] + sx * hbn;
      x[i] = xn;
      acc += xn * xn;
      wv[i] = fma(sh_h, wv[i], v[i]);                  // h = v + scale * h
      vin[i] = v[i];
    }
    const double normx = sqrt(block_sum(acc, sh));
    if (threadIdx.x == 0) {
      const double normr = shd[3], normA = shd[4], condA = shd[5], normb = sc[LQ_BNORM];
      const double zb = sc[LM_ZETABAR];
      const double normar = zb < 0 ? -zb : zb;
      const double ctol = g.conlim > 0 ? 1.0 / g.conlim : 0.0;
      const double test1 = normr / normb;
      const double test2 = (normA * normr) != 0.0 ? normar / (normA * normr) : INFINITY;
      const double test3 = 1.0 / condA;
      const double t1 = test1 / (1.0 + normA * normx / normb);
      const double rtol = g.btol + g.atol * normA * normx / normb;
      int istop = 0;
      if (itn >= g.iter_lim) istop = 7;
      if (1.0 + test3 <= 1.0) istop = 6;
      if (1.0 + test2 <= 1.0) istop = 5;
      if (1.0 + t1 <= 1.0) istop = 4;
      if (test3 <= ctol) istop = 3;
      if (test2 <= g.atol) istop = 2;
      if (test1 <= rtol) istop = 1;
      ic[LI_ISTOP] = istop;
      ic[LI_PHASE] = 0;
      if (istop != 0)
        ic[LI_DONE] = 1;
      else
        atomicAdd(g.n_active, 1);
    }
    return;
  }
  if (threadIdx.x == 0) {
    double cs, sn, rho;
    sym_ortho(sc[LQ_RHOBAR], beta, cs, sn, rho);
    const double theta = sn * alfa;
    const double phi = cs * sc[LQ_PHIBAR];
    sc[LQ_RHOBAR] = -cs * alfa;
    sc[LQ_PHIBAR] = sn * sc[LQ_PHIBAR];
    shd[0] = phi / rho;     // t1
    shd[1] = -theta / rho;  // t2
    shd[2] = 1.0 / rho;
    shd[3] = theta;
    sc[LQ_ALFA] = alfa;
    sc[LQ_RES2] = sn * phi;  // tau (res2 itself stays 0 when damp = 0)
    // second rotation: estimate |x|
    const double delta = sc[LQ_SN2] * rho, gambar = -sc[LQ_CS2] * rho;
    const double rhs = phi - delta * sc[LQ_Z];
    const double zbar = rhs / gambar;
    const double xnorm = sqrt(sc[LQ_XXNORM] + zbar * zbar);
    const double gamma = sqrt(gambar * gambar + theta * theta);
    sc[LQ_CS2] = gambar / gamma;
    sc[LQ_SN2] = theta / gamma;
    sc[LQ_Z] = rhs / gamma;
    sc[LQ_XXNORM] += sc[LQ_Z] * sc[LQ_Z];
    shd[4] = xnorm;
  }
  __syncthreads();
  const double t1 = shd[0], t2 = shd[1], irho = shd[2];
  const double xnorm = shd[4];
  double acc = 0.0;
  for (int i = threadIdx.x; i < g.N; i += kThreads) {
    const double wi = wv[i];
    const double dk = irho * wi;
    acc += dk * dk;
    x[i] += t1 * wi;
    const double nw = v[i] + t2 * wi;
    wv[i] = nw;
    vin[i] = v[i];
  }
  const double dd = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    const double ddnorm = sc[LQ_DDNORM] + dd;
    sc[LQ_DDNORM] = ddnorm;
    const double anorm = sc[LQ_ANORM], bnorm = sc[LQ_BNORM];
    const double acond = anorm * sqrt(ddnorm);
    const double phibar = sc[LQ_PHIBAR];
    const double rnorm = sqrt(phibar * phibar);
    const double tau = sc[LQ_RES2];
    const double arnorm = alfa * (tau < 0 ? -tau : tau);
    const double ctol = g.conlim > 0 ? 1.0 / g.conlim : 0.0;
    const double test1 = rnorm / bnorm;
    const double test2 = arnorm / (anorm * rnorm + kEps);
    const double test3 = 1.0 / (acond + kEps);
    const double tt1 = test1 / (1.0 + anorm * xnorm / bnorm);
    const double rtol = g.btol + g.atol * anorm * xnorm / bnorm;
    int istop = 0;
    if (itn >= g.iter_lim) istop = 7;
    if (1.0 + test3 <= 1.0) istop = 6;
    if (1.0 + test2 <= 1.0) istop = 5;
    if (1.0 + tt1 <= 1.0) istop = 4;
    if (test3 <= ctol) istop = 3;
    if (test2 <= g.atol) istop = 2;
    if (test1 <= rtol) istop = 1;
    ic[LI_ISTOP] = istop;
    ic[LI_PHASE] = 0;
    if (istop != 0)
      ic[LI_DONE] = 1;
    else
      atomicAdd(g.n_active, 1);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Newton outer loop bookkeeping (optimize.py:1373-1399): trial = u + step*dx, backtrack while next_cost > cost.
// ---------------------------------------------------------------------------------------------------------------
struct NewtonState {
  double* u;         // [B][Nm] in/out
  double* up;        // [B][3]
  const double* dx;  // [B][N]  cdof correction from the linear solver
  double* tu;        // [B][Nm] trial
  double* tup;       // [B][3]
  const double* tcost;  // [B]  cost at the trial point
  double* cost;      // [B]
  double* step;      // [B]
  int* status;       // [B]
  int* nit;          // [B]
  int* live;         // [B]  still inside the outer while-loop
  int* searching;    // [B]  line search of the current outer iteration not yet decided
  int* n_searching;  // [1]
  int B, nm, N, maxiter;
  unsigned cmask;
  double tol, ftol, min_step;
  int mode;          // 0: start the line search (step = 1, first trial), 1: judge the trial
};

__global__ void __launch_bounds__(kThreads) newton_step_kernel(const NewtonState s) {
  const int b = blockIdx.x;
  __shared__ int sh_accept, sh_again;
  __shared__ double sh_step;
  const int live = s.live[b], searching = s.searching[b];
  __syncthreads();
  if (!live) return;
  if (s.mode == 1 && !searching) return;
  if (threadIdx.x == 0) {
    int accept = 0, again = 1;
    double step = 1.0;
    if (s.mode == 1) {
      step = s.step[b];
      double cost = s.cost[b];
      const double ct = s.tcost[b];
      int status = s.status[b];
      if (ct > cost && step > s.min_step) {                                               // (:1377-1379)
        step *= 0.5;
      } else {
        again = 0;
        const int nit = s.nit[b] + 1;                                                     // _process_correction
        s.nit[b] = nit;
        if (ct <= s.tol) {
          status = -1;
          accept = 1;
        } else if (step <= s.min_step) {
          status = 0;
        } else if (nit == s.maxiter) {
          status = 2;
          accept = (ct < cost) ? 1 : 0;
        } else {
          const double big = fmax(fmax(cost, ct), 1.0);
          if ((cost - ct) / big < s.ftol) status = 3; else accept = 1;
        }
        if (accept) cost = ct;
        int still = (cost > s.tol && status == -1) ? 1 : 0;
        if (!still && cost <= s.tol) status = -1;                                         // (:1401-1402)
        s.cost[b] = cost;
        s.status[b] = status;
        s.live[b] = still;
      }
    }
    s.step[b] = step;
    s.searching[b] = again;
    if (again) atomicAdd(s.n_searching, 1);
    sh_accept = accept;
    sh_again = again;
    sh_step = step;
  }
  __syncthreads();
  const int accept = sh_accept, again = sh_again;
  const double step = sh_step;
  double* u = s.u + (size_t)b * s.nm;
  double* tu = s.tu + (size_t)b * s.nm;
  const double* dx = s.dx + (size_t)b * s.N;
  for (int i = threadIdx.x; i < s.nm; i += kThreads) {
    double ui = u[i];
    if (accept) {
      ui = tu[i];
      u[i] = ui;
    }
    if (again) tu[i] = ui + step * dx[i];                                                 // increment(dx, step) (:1374, :1380)
  }
  if (threadIdx.x == 0) {
    int slot = 0;
    for (int i = 0; i < 3; ++i) {
      double pi = s.up[b * 3 + i];
      if (accept) {
        pi = s.tup[b * 3 + i];
        s.up[b * 3 + i] = pi;
      }
      const double dpi = (s.cmask >> i) & 1u ? 0.0 : dx[s.nm + slot++];
      if (again) s.tup[b * 3 + i] = pi + step * dpi;
    }
  }
}

// b = -F (LSQR right-hand side) : out = -in
__global__ void negate_kernel(const double* in, double* out, size_t count) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) out[i] = -in[i];
}

}  // namespace kskry
