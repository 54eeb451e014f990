// Stage D of the solve kernel: K = H^-1 by symmetric sweeps with the matrix held in register tiles.
//
// Thread (ty, tx) of a TY x TX grid owns rows i = ty + TY*a (a < TM) and columns j = tx + TX*b
// (b < TN) of the NPAD x NPAD padded Hessian.  Per pivot k the threads that own ROW k (half a warp;
// row k equals column k by symmetry) publish it to shared memory (double-buffered vector, one
// barrier per pivot) and every thread applies one rank-1 update to its TM x TN tile from registers.
//
// Sweep of pivot k on a symmetric matrix (d = a_kk, c = column k):
//   a_ij <- a_ij - c_i c_j / d   (i, j != k),   a_ik <- c_i / d,   a_kk <- -1/d
// after all n pivots the matrix is -H^-1 (H = 2aI + 2B'SB is SPD: no pivoting).
//
// The pivot row and column are NOT patched separately.  Publishing  c_k := d - 1  in place of d makes
// the same rank-1 update produce them:
//   row k:     c_j - (d-1) c_j / d              =  c_j / d
//   column k:  c_i - c_i (d-1) / d              =  c_i / d
//   diagonal:  d   - (d-1)(d-1) / d             =  2 - 1/d      (a constant +2 off)
// A swept diagonal element is only ever ADDED to by later pivots, never used as a multiplier, so the
// +2 is removed once, after the last pivot, from the statically known diagonal registers.  The
// identities are free of cancellation for d <= 1; H is scaled by an exact power of two so that its
// largest diagonal entry, and hence every pivot (Schur complements of an SPD matrix only shrink the
// diagonal), is below 1.  The pivot loop is rolled and has no register index that depends on k other
// than the publishing row block, which is selected by a predicated chain in one warp only.
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
__device__ __forceinline__ void build_invert_regtile2(const HessCtx& C, const int* rowinfo, int n, int tid, double* cbuf,
                                                      double* K, double* red) {
  constexpr int TY = S::TY, TX = S::TX, TM = S::TM, TN = S::TN, NPAD = S::NPAD, NT = S::NT;
  const int ty = tid / TX, tx = tid - ty * TX;
  // H into shared memory: lower triangle evaluated once and mirrored; track the largest diagonal entry
  double dmax = 0.0;
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int I = warp; I < n; I += NT / 32) {
      const int ri = rowinfo[I];
      for (int J = lane; J <= I; J += 32) {
        const double v = hess_entry(C, ri, rowinfo[J], I == J);
        K[I * n + J] = v;
        K[J * n + I] = v;
        if (I == J) dmax = fmax(dmax, v);
      }
    }
  }
  {
    int dummy = 0;
    double neg = -dmax;
    block_argmin<NT>(neg, dummy, red, tid);  // also orders the K writes before the tile loads
    dmax = -neg;
  }
  if (NT <= 32) __syncthreads();
  int e2;
  frexp(dmax, &e2);                          // dmax = m * 2^e2, m in [0.5, 1)
  const double scale = ldexp(1.0, -e2);      // exact; scaled diagonal < 1
  double A[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; a++)
#pragma unroll
    for (int b = 0; b < TN; b++) {
      const int i = ty + TY * a, j = tx + TX * b;
      A[a][b] = (i < n && j < n) ? K[i * n + j] * scale : (i == j ? 0.5 : 0.0);  // harmless padding beyond n
    }
  int par = 0;
#pragma unroll 1
  for (int k = 0; k < n; k++) {
    const int a = k / TY, kk = k - a * TY;  // row block / row-owner ty
    double* cb = cbuf + par * (NPAD + 2);
    if (ty == kk) {
#pragma unroll
      for (int aa = 0; aa < TM; aa++) {
        if (aa == a) {
#pragma unroll
          for (int bb = 0; bb < TN; bb++) {
            const int j = tx + TX * bb;
            double v = A[aa][bb];
            if (j == k) { cb[NPAD] = v; v -= 1.0; }
            cb[j] = v;
          }
        }
      }
    }
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
    par ^= 1;
  }
  // -swept = (scaled H)^-1; undo the scaling and the +2 carried by every diagonal element
#pragma unroll
  for (int a = 0; a < TM; a++)
#pragma unroll
    for (int b = 0; b < TN; b++) {
      const int i = ty + TY * a, j = tx + TX * b;
      if (i < n && j < n) K[i * n + j] = -(A[a][b] - (i == j ? 2.0 : 0.0)) * scale;
    }
}

}  // namespace
