// Batched adjoint descent: per-orbit backtracking line search and exit codes, driven by the fused costgrad kernels.
//
// Reference path replaced: orbithunter/optimize.py _adjoint_descent :367-509 and _process_correction :2115-2239,
// Orbit.increment core.py:1760-1796.  Every orbit carries its own (cost, step_size, nit, status); each outer pass is
// two launches for the whole batch: ks costgrad at the trial points, then `adj_commit_trial_kernel`, which applies
// the reference's decision logic per orbit (halve the step while next_cost >= cost and step > min_step; otherwise
// _process_correction in the order tol -> min_step -> maxiter -> ftol), commits accepted steps and forms the next
// trial point u - step * grad.  Orbits that have left the loop are frozen (their trial point stops changing).
#pragma once
#include "ks_common.cuh"

namespace ksadj {

struct AdjState {
  // persistent per-orbit state
  double* u;        // [B, Nm]  current orbit (in/out)
  double* up;       // [B, 3]
  double* g;        // [B, Nm]  gradient at u (possibly preconditioned)
  double* gp;       // [B, 3]
  double* cost;     // [B]
  double* step;     // [B]
  int* status;      // [B]   -1 running/has-not-failed, 0 min step, 2 maxiter, 3 insufficient decrease
  int* nit;         // [B]
  int* live;        // [B]   1 while the orbit is still inside the while-loop
  // trial point and its evaluation (written by costgrad)
  double* tu;       // [B, Nm]
  double* tup;      // [B, 3]
  double* tg;       // [B, Nm]
  double* tgp;      // [B, 3]
  double* tcost;    // [B]
  int* n_live;      // [1]  number of live orbits after this pass (atomic)
  int batch, nm;
  double tol, ftol, min_step;
  int maxiter;
  int first;        // 1: the evaluation was at u itself (initial F and gradient); no step is counted
};

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads) adj_commit_trial_kernel(const AdjState s) {
  const int b = blockIdx.x;
  if (!s.live[b]) return;
  __shared__ int sh_accept, sh_live;
  __shared__ double sh_step;
  if (threadIdx.x == 0) {
    double cost = s.cost[b], step = s.step[b];
    const double ct = s.tcost[b];
    int accept = 0, live = 1, status = s.status[b];
    if (s.first) {
      cost = ct;  // F = eqn(u), cost = 1/2 F.F                                       (optimize.py:462-463)
      accept = 1;
      live = (cost > s.tol) ? 1 : 0;                                               // while cost > tol ...  (:467)
    } else if (ct >= cost && step > s.min_step) {
      step *= 0.5;                                                                 // backtracking          (:478-480)
    } else {
      const int nit = s.nit[b] + 1;                                                // _process_correction   (:2174-2189)
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
        if ((cost - ct) / big < s.ftol)
          status = 3;
        else
          accept = 1;
      }
      if (accept) cost = ct;
      live = (cost > s.tol && status == -1) ? 1 : 0;
    }
    if (!live && cost <= s.tol) status = -1;                                       // loop exit fix-up      (:506-507)
    s.cost[b] = cost;
    s.step[b] = step;
    s.status[b] = status;
    if (live) atomicAdd(s.n_live, 1);
    sh_accept = accept;
    sh_live = live;
    sh_step = step;
  }
  __syncthreads();
  // s.live[b] is the entry test of every thread of this block: only publish it once all of them are past it
  if (threadIdx.x == 0) s.live[b] = sh_live;
  const int accept = sh_accept, live = sh_live;
  const double step = sh_step;
  double* u = s.u + (size_t)b * s.nm;
  double* g = s.g + (size_t)b * s.nm;
  double* tu = s.tu + (size_t)b * s.nm;
  const double* tg = s.tg + (size_t)b * s.nm;
  for (int i = threadIdx.x; i < s.nm; i += kThreads) {
    double ui = u[i], gi = g[i];
    if (accept) {
      ui = tu[i];
      gi = tg[i];
      u[i] = ui;
      g[i] = gi;
    }
    if (live) tu[i] = ui - step * gi;                                              // increment(grad, -step) (:471)
  }
  if (threadIdx.x < 3) {
    const int i = threadIdx.x;
    double pi = s.up[b * 3 + i], gi = s.gp[b * 3 + i];
    if (accept) {
      pi = s.tup[b * 3 + i];
      gi = s.tgp[b * 3 + i];
      s.up[b * 3 + i] = pi;
      s.gp[b * 3 + i] = gi;
    }
    if (live) s.tup[b * 3 + i] = pi - step * gi;
  }
}

}  // namespace ksadj
