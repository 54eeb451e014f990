// Stage D of the solve kernel: K = H^-1 by symmetric sweeps with the matrix held in register tiles.
//
// Thread (ty, tx) of a TY x TX grid owns rows i = ty + TY*a (a < TM) and columns j = tx + TX*b
// (b < TN) of the NPAD x NPAD padded Hessian.  Per pivot k the owners of column k publish it to
// shared memory (one double-buffered vector, one barrier), every thread applies the rank-1 update
// to its TM x TN tile from registers, and the pivot row / column are patched in place.  The pivot
// loop is rolled; the few accesses whose register index depends on k sit behind warp-uniform
// switches, so the whole stage is a few hundred instructions and stays in the instruction cache.
//
// Sweep of pivot k on a symmetric matrix (d = a_kk):
//   a_ij <- a_ij - a_ik a_kj / d   (i, j != k),   a_ik <- a_ik / d,   a_kk <- -1/d
// after all n pivots the matrix is -H^-1.  H = 2aI + 2B'SB is SPD, so no pivoting is needed.
#pragma once

namespace {

__device__ __forceinline__ double fast_rcp(double d) {
  // MUFU.RCP64H seed + two Newton steps: relative error ~1e-16 for the normal, positive pivots seen here
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

template <class S>
__device__ __forceinline__ void publish_column(const double (&A)[S::TM][S::TN], int b, int k, int ty, double* cb) {
  // column block b is warp-uniform: one case runs, with compile-time register indices
#pragma unroll
  for (int bb = 0; bb < S::TN; bb++) {
    if (bb == b) {
#pragma unroll
      for (int aa = 0; aa < S::TM; aa++) {
        const int i = ty + S::TY * aa;
        double v = A[aa][bb];
        if (i == k) { cb[S::NPAD] = v; v = -1.0; }  // pivot goes to its own slot; -1 makes the patches below yield -1/d
        cb[i] = v;
      }
    }
  }
}

template <class S>
__device__ __forceinline__ void build_invert_regtile2(const HessCtx& C, const int* rowinfo, int n, int tid, double* cbuf,
                                                      double* K) {
  constexpr int TY = S::TY, TX = S::TX, TM = S::TM, TN = S::TN, NPAD = S::NPAD, NT = S::NT;
  const int ty = tid / TX, tx = tid - ty * TX;
  // H into shared memory: lower triangle evaluated once and mirrored
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int I = warp; I < n; I += NT / 32) {
      const int ri = rowinfo[I];
      for (int J = lane; J <= I; J += 32) {
        const double v = hess_entry(C, ri, rowinfo[J], I == J);
        K[I * n + J] = v;
        K[J * n + I] = v;
      }
    }
  }
  __syncthreads();
  double A[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; a++)
#pragma unroll
    for (int b = 0; b < TN; b++) {
      const int i = ty + TY * a, j = tx + TX * b;
      A[a][b] = (i < n && j < n) ? K[i * n + j] : (i == j ? 1.0 : 0.0);  // identity padding beyond n
    }
  int par = 0;
#pragma unroll 1
  for (int k = 0; k < n; k++) {
    const int a = k / TY, kk = k - a * TY;  // row block / row-owner ty
    const int b = k / TX, ko = k - b * TX;  // column block / column-owner tx
    double* cb = cbuf + par * (NPAD + 2);
    if (tx == ko) publish_column<S>(A, b, k, ty, cb);
    __syncthreads();
    const double dinv = fast_rcp(cb[NPAD]);
    double ci[TM], cjd[TN];
#pragma unroll
    for (int aa = 0; aa < TM; aa++) ci[aa] = cb[ty + TY * aa];
#pragma unroll
    for (int bb = 0; bb < TN; bb++) cjd[bb] = cb[tx + TX * bb] * dinv;
#pragma unroll
    for (int aa = 0; aa < TM; aa++)
#pragma unroll
      for (int bb = 0; bb < TN; bb++) A[aa][bb] = fma(-ci[aa], cjd[bb], A[aa][bb]);
    if (ty == kk) {  // pivot row: a_kj / d  (and -1/d on the diagonal through the -1 published for i == k)
#pragma unroll
      for (int aa = 0; aa < TM; aa++)
        if (aa == a) {
#pragma unroll
          for (int bb = 0; bb < TN; bb++) A[aa][bb] = cjd[bb];
        }
    }
    if (tx == ko) {  // pivot column: a_ik / d
#pragma unroll
      for (int bb = 0; bb < TN; bb++)
        if (bb == b) {
#pragma unroll
          for (int aa = 0; aa < TM; aa++) A[aa][bb] = ci[aa] * dinv;
        }
    }
    par ^= 1;
  }
#pragma unroll
  for (int a = 0; a < TM; a++)
#pragma unroll
    for (int b = 0; b < TN; b++) {
      const int i = ty + TY * a, j = tx + TX * b;
      if (i < n && j < n) K[i * n + j] = -A[a][b];
    }
}

}  // namespace
