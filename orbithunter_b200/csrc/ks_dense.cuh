// Batched dense Newton step of hunt(..., 'lstsq'): dx = the minimum-norm solution of J dx = -F for many orbits at once.
//
// Reference path replaced: orbithunter/optimize.py _lstsq :855-1022, which builds the dense Jacobian (ks/orbits.py
// jacobian :557-657) and calls scipy.linalg.lstsq (LAPACK gelsd) orbit by orbit.  Here, for a chunk of C orbits:
//   J   (Nm x N, column-major, one per orbit)  = one batched ks_matvec over the identity (N = Nm + free parameters),
//   G   = J J^T                                  (cuBLAS DgemmStridedBatched: the one plain library GEMM of the path),
//   G   = L L^T                                  blocked right-looking Cholesky: the diagonal blocks are factored and
//                                                inverted by chol_diag_kernel below, the panel and trailing updates are GEMMs,
//   y   = L^-T L^-1 (-F),  dx = J^T y            tri_solve_kernel / gemv_kernel (one CTA per orbit, L streamed once each way),
//   refinement: r = -F - J dx, dx += J^T G^-1 r  (twice).
// J has full row rank away from bifurcations (Nm equations, Nm + 2 or 3 unknowns), so J^T (J J^T)^-1 is its pseudo-inverse and
// dx is the same minimum-norm step the reference takes; a failed factorisation (non-positive pivot) is redone with a
// Tikhonov shift G + lambda I, which is what gelsd's singular-value cut-off amounts to for such a J.
#pragma once
#include "ks_common.cuh"

namespace ksd {

constexpr int kNb = 64;       // Cholesky block size
constexpr int kThreads = 256;
constexpr int kDiagSmemBytes = 2 * kNb * (kNb + 1) * 8;

// out[c][col][:] = in[c][:] for col < ncol (ld = stride of one copy): the orbit replicated once per Jacobian column
__global__ void replicate_kernel(const double* in, double* out, int len, int ncol, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / len;
    const int k = (int)(i - row * len);
    out[i] = in[(row / ncol) * len + k];
  }
}
// V[c][col][j] = (col == j): the identity in the cdof layout, one copy per orbit of the chunk
__global__ void identity_kernel(double* V, int ncol, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / ncol;
    V[i] = ((int)(row % ncol) == (int)(i - row * ncol)) ? 1.0 : 0.0;
  }
}
// gather the live orbits of a chunk: U[c] = modes[idx[c]], P[c] = params[idx[c]], rhs[c] = -F[idx[c]]
__global__ void gather_kernel(const int* idx, const double* modes, const double* params, const double* F, double* U, double* P, double* rhs,
                              int nm) {
  const int c = blockIdx.x, b = idx[c];
  for (int i = threadIdx.x; i < nm; i += blockDim.x) {
    U[(size_t)c * nm + i] = modes[(size_t)b * nm + i];
    rhs[(size_t)c * nm + i] = -F[(size_t)b * nm + i];
  }
  if (threadIdx.x < 3) P[c * 3 + threadIdx.x] = params[b * 3 + threadIdx.x];
}
__global__ void scatter_kernel(const int* idx, const double* dx, double* xsol, int N) {
  const int c = blockIdx.x, b = idx[c];
  for (int i = threadIdx.x; i < N; i += blockDim.x) xsol[(size_t)b * N + i] = dx[(size_t)c * N + i];
}
// out += in  /  out = rhs - out  over n doubles (matrix-free refinement: the products come from the matvec / rmatvec kernels)
__global__ void add_kernel(double* out, const double* in, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] += in[i];
}
__global__ void residual_kernel(double* out, const double* rhs, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = rhs[i] - out[i];
}
// G += lambda[c] * I
__global__ void shift_diag_kernel(double* G, const double* lambda, int nm) {
  const int c = blockIdx.x;
  const double l = lambda[c];
  if (l == 0.0) return;
  double* g = G + (size_t)c * nm * nm;
  for (int i = threadIdx.x; i < nm; i += blockDim.x) g[(size_t)i * nm + i] += l;
}
// lambda[c]: 0 on the first attempt; after a failed factorisation (fail[c] != 0) a shift relative to the mean diagonal of
// J J^T that grows by 1e3 per retry.  `redo` counts the orbits that need another attempt.
__global__ void lambda_kernel(const double* G0diag_mean, double* lambda, int* fail, int* redo, int attempt, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (attempt == 0) {
    lambda[c] = 0.0;
    fail[c] = 0;
  } else if (fail[c]) {
    double l = 1e-13 * G0diag_mean[c];
    for (int k = 1; k < attempt; ++k) l *= 1e3;
    lambda[c] = l;
    fail[c] = 0;
    atomicAdd(redo, 1);
  }
}
__global__ void diag_mean_kernel(const double* G, double* mean, int nm) {
  __shared__ double sh[kThreads / 32];
  const int c = blockIdx.x;
  const double* g = G + (size_t)c * nm * nm;
  double acc = 0.0;
  for (int i = threadIdx.x; i < nm; i += blockDim.x) acc += g[(size_t)i * nm + i];
  acc += __shfl_xor_sync(0xffffffffu, acc, 16);
  acc += __shfl_xor_sync(0xffffffffu, acc, 8);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) t += sh[i];
    mean[c] = t / nm;
  }
}

// Diagonal block [k0, k0 + b) of the blocked Cholesky, one CTA per orbit: L_kk = chol(G_kk) (lower) written to Lf, and
// Li = L_kk^-1 (lower, b x b, ld kNb).  A non-positive pivot sets fail[c]; the block is then completed with unit pivots so
// that the following GEMMs stay finite (the orbit is refactored with a shift).
__global__ void __launch_bounds__(kThreads) chol_diag_kernel(const double* G, double* Lf, double* Li, int* fail, int nm, int k0, int b, int blk) {
  KS_DYN_SMEM(double, dsm);  // 2 * kNb * (kNb + 1) doubles
  double (*A)[kNb + 1] = reinterpret_cast<double (*)[kNb + 1]>(dsm);                      // A[col][row]
  double (*X)[kNb + 1] = reinterpret_cast<double (*)[kNb + 1]>(dsm + kNb * (kNb + 1));
  __shared__ int bad;
  const int c = blockIdx.x;
  const double* g = G + (size_t)c * nm * nm;
  if (threadIdx.x == 0) bad = 0;
  for (int e = threadIdx.x; e < b * b; e += blockDim.x) {
    const int col = e / b, row = e - col * b;
    A[col][row] = (row >= col) ? g[(size_t)(k0 + col) * nm + k0 + row] : 0.0;
    X[col][row] = 0.0;
  }
  __syncthreads();
  for (int j = 0; j < b; ++j) {
    double d = A[j][j];
    __syncthreads();
    if (!(d > 0.0)) {
      d = 1.0;
      if (threadIdx.x == 0) bad = 1;
    }
    const double r = sqrt(d), ir = 1.0 / r;
    for (int i = j + threadIdx.x; i < b; i += blockDim.x) A[j][i] = (i == j) ? r : A[j][i] * ir;
    __syncthreads();
    // trailing update of the block: A[k][i] -= A[j][i] * A[j][k] for j < k <= i; thread = (row i, every fourth column k)
    {
      const int i = threadIdx.x & (kNb - 1);
      if (i > j && i < b) {
        const double aji = A[j][i];
        for (int kk = j + 1 + (threadIdx.x >> 6); kk <= i; kk += kThreads / kNb) A[kk][i] -= aji * A[j][kk];
      }
    }
    __syncthreads();
  }
  // inverse of the lower triangle, one column per thread (forward substitution with a unit vector)
  if (threadIdx.x < b) {
    const int j = threadIdx.x;
    X[j][j] = 1.0 / A[j][j];
    for (int i = j + 1; i < b; ++i) {
      double s = 0.0;
#pragma unroll 8
      for (int k = j; k < i; ++k) s = fma(A[k][i], X[j][k], s);
      X[j][i] = -s / A[i][i];
    }
  }
  __syncthreads();
  double* lf = Lf + (size_t)c * nm * nm;
  double* li = Li + ((size_t)c * ((nm + kNb - 1) / kNb) + blk) * kNb * kNb;
  for (int e = threadIdx.x; e < b * b; e += blockDim.x) {
    const int col = e / b, row = e - col * b;
    lf[(size_t)(k0 + col) * nm + k0 + row] = A[col][row];  // strictly upper part of the block: zeros
    li[col * kNb + row] = X[col][row];
  }
  if (threadIdx.x == 0 && bad) fail[c] = 1;
}

// y = L^-T L^-1 r for one orbit per CTA.  Lf holds L (column-major, ld nm), Li the inverted diagonal blocks.  r is read from
// `rhs`, y written to `y` (both [C][nm]); the vector lives in shared memory, L is streamed once per direction.
// The panel updates are the HBM traffic of the kernel.  Forward, a thread owns two adjacent rows (one 16-byte load per column,
// two columns in flight per step); backward, eight lanes share a column (the 64 rows of the block are 512 contiguous bytes:
// four 16-byte loads per lane, then a three-step shuffle reduction), four columns per warp.  Both need 16-byte alignment, i.e.
// an even nm; odd nm (ShiftReflection / Antisymmetric grids with odd mode counts) takes the scalar loops.
__global__ void __launch_bounds__(kThreads) tri_solve_kernel(const double* Lf, const double* Li, const double* rhs, double* y, int nm) {
  KS_DYN_SMEM(double, vec);  // nm doubles + kNb
  double* zk = vec + nm;
  const int c = blockIdx.x;
  const double* L = Lf + (size_t)c * nm * nm;
  const int nblk = (nm + kNb - 1) / kNb;
  const double* li0 = Li + (size_t)c * nblk * kNb * kNb;
  const bool vec2 = (nm & 1) == 0;
  for (int i = threadIdx.x; i < nm; i += blockDim.x) vec[i] = rhs[(size_t)c * nm + i];
  __syncthreads();
  // forward: z_k = Li_k v_k;  v[k1:] -= L[k1:, k0:k1] z_k
  for (int blk = 0; blk < nblk; ++blk) {
    const int k0 = blk * kNb, b = (nm - k0 < kNb) ? nm - k0 : kNb, k1 = k0 + b;
    const double* li = li0 + (size_t)blk * kNb * kNb;
    if (threadIdx.x < b) {
      const int i = threadIdx.x;
      double s = 0.0;
      for (int j = 0; j <= i; ++j) s = fma(li[j * kNb + i], vec[k0 + j], s);
      zk[i] = s;
    }
    __syncthreads();
    if (threadIdx.x < b) vec[k0 + threadIdx.x] = zk[threadIdx.x];
    if (vec2 && b == kNb) {  // (rows exist below the block only when it is a full one)
      for (int i = k1 + 2 * threadIdx.x; i < nm; i += 2 * blockDim.x) {
        const double* p = L + (size_t)k0 * nm + i;
        double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll 8
        for (int j = 0; j < kNb; j += 2) {
          const double2 a0 = *reinterpret_cast<const double2*>(p + (size_t)j * nm);
          const double2 a1 = *reinterpret_cast<const double2*>(p + (size_t)(j + 1) * nm);
          s0 = fma(a0.x, zk[j], s0);
          s1 = fma(a0.y, zk[j], s1);
          t0 = fma(a1.x, zk[j + 1], t0);
          t1 = fma(a1.y, zk[j + 1], t1);
        }
        vec[i] -= s0 + t0;
        vec[i + 1] -= s1 + t1;
      }
    } else {
      for (int i = k1 + threadIdx.x; i < nm; i += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < b; ++j) s = fma(L[(size_t)(k0 + j) * nm + i], zk[j], s);
        vec[i] -= s;
      }
    }
    __syncthreads();
  }
  // backward: y_k = Li_k^T v_k;  v[0:k0] -= L[k0:k1, 0:k0]^T y_k
  for (int blk = nblk - 1; blk >= 0; --blk) {
    const int k0 = blk * kNb, b = (nm - k0 < kNb) ? nm - k0 : kNb;
    const double* li = li0 + (size_t)blk * kNb * kNb;
    if (threadIdx.x < b) {
      const int i = threadIdx.x;
      double s = 0.0;
      for (int j = i; j < b; ++j) s = fma(li[i * kNb + j], vec[k0 + j], s);
      zk[i] = s;
    }
    __syncthreads();
    if (threadIdx.x < b) vec[k0 + threadIdx.x] = zk[threadIdx.x];
    if (vec2) {
      if (threadIdx.x >= b && threadIdx.x < kNb) zk[threadIdx.x] = 0.0;  // a short last block: pad the multiplier with zeros
      __syncthreads();
      const int l8 = threadIdx.x & 7, r0 = 8 * l8;
      // rows k0 + r0 .. + 7 of a column; past the end of a short block the multiplier is zero and the load is skipped
      double z[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) z[e] = zk[r0 + e];
      // warp-uniform trip count (the shuffles below name the whole warp); k0 is a multiple of 64, so j < k0 for every group
      for (int jw = (threadIdx.x >> 5) * 4; jw < k0; jw += (blockDim.x >> 5) * 4) {
        const int j = jw + ((threadIdx.x & 31) >> 3);
        const double* col = L + (size_t)j * nm + k0 + r0;
        double s = 0.0, t = 0.0;
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          if (j < k0 && r0 + e < b) {  // b is even when nm is: pairs never straddle the end
            const double2 a = *reinterpret_cast<const double2*>(col + e);
            s = fma(a.x, z[e], s);
            t = fma(a.y, z[e + 1], t);
          }
        }
        s += t;
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (l8 == 0 && j < k0) vec[j] -= s;
      }
    } else {
      for (int j = threadIdx.x; j < k0; j += blockDim.x) {
        const double* col = L + (size_t)j * nm + k0;
        double s = 0.0;
        for (int r = 0; r < b; ++r) s = fma(col[r], zk[r], s);
        vec[j] -= s;
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < nm; i += blockDim.x) y[(size_t)c * nm + i] = vec[i];
}

// mode 0: out[c][col] (+)= sum_i J[c][i + col * nm] y[c][i]   (J^T y: the cdof-layout correction; `accumulate` adds to out)
// mode 1: out[c][i] = rhs[c][i] - sum_col J[c][i + col * nm] x[c][col]   (residual of the linear system)
__global__ void __launch_bounds__(kThreads) gemv_kernel(const double* J, const double* x, const double* rhs, double* out, int nm, int ncol,
                                                        int mode, int accumulate) {
  KS_DYN_SMEM(double, xs);
  const int c = blockIdx.x;
  const double* Jc = J + (size_t)c * nm * ncol;
  const int nx = mode == 0 ? nm : ncol;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) xs[i] = x[(size_t)c * nx + i];
  __syncthreads();
  if (mode == 0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int col = warp; col < ncol; col += nwarp) {
      const double* jc = Jc + (size_t)col * nm;
      double s = 0.0;
      for (int i = lane; i < nm; i += 32) s = fma(jc[i], xs[i], s);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (lane == 0) out[(size_t)c * ncol + col] = accumulate ? out[(size_t)c * ncol + col] + s : s;
    }
  } else {
    for (int i = threadIdx.x; i < nm; i += blockDim.x) {
      double s = 0.0;
      for (int col = 0; col < ncol; ++col) s = fma(Jc[(size_t)col * nm + i], xs[col], s);
      out[(size_t)c * nm + i] = rhs[(size_t)c * nm + i] - s;
    }
  }
}

}  // namespace ksd
