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

template <class S>
__device__ __forceinline__ void build_invert_regtile2(const HessCtx& C, const int* rowinfo, int n, int tid, double* cbuf,
                                                      double* K, double* red, PhaseClock& pc) {
  constexpr int TY = S::TY, TX = S::TX, TM = S::TM, TN = S::TN, NPAD = S::NPAD, NT = S::NT;
  const int ty = tid / TX, tx = tid - ty * TX;
  // H into shared memory, one 3x3 foot-step pair block (j1 >= j2) per thread and mirrored: the horizon
  // sums are fetched once per block and the component pattern of the x_drag terms is static.
  double dmax = 0.0;
  {
    const int nc = n / 3, nb = nc * (nc + 1) / 2, hh = C.h * C.h;
    for (int blk = tid; blk < nb; blk += NT) {
      int j1 = (int)((sqrt(8.0 * (double)blk + 1.0) - 1.0) * 0.5);
      while ((j1 + 1) * (j1 + 2) / 2 <= blk) j1++;
      while (j1 * (j1 + 1) / 2 > blk) j1--;
      const int j2 = blk - j1 * (j1 + 1) / 2;
      const int k1 = C.fs[j1], k2 = C.fs[j2];
      const int sa = k1 >> 2, fi = k1 & 3, sb = k2 >> 2, fj = k2 & 3;
      const int ab = sa * C.h + sb, ba = sb * C.h + sa;
      const double s11 = __ldg(C.sig + CMPC_SIG_11 * hh + ab), s22 = __ldg(C.sig + CMPC_SIG_22 * hh + ab);
      double x20 = 0.0, x02 = 0.0, x00 = 0.0;  // x_drag couplings (z,x), (x,z), (x,x)
      if (C.xd != 0.0) {
        x20 = C.xd * (C.wp[2] * __ldg(C.sig + CMPC_SIG_23 * hh + ab) + C.wp[5] * __ldg(C.sig + CMPC_SIG_12 * hh + ab));
        x02 = C.xd * (C.wp[2] * __ldg(C.sig + CMPC_SIG_23 * hh + ba) + C.wp[5] * __ldg(C.sig + CMPC_SIG_12 * hh + ba));
        x00 = C.xd * C.xd * (C.wp[2] * __ldg(C.sig + CMPC_SIG_33 * hh + ab) + C.wp[5] * s22);
      }
      const double* pt = C.sPT + (fi * 4 + fj) * 9;
      const double* po = C.sPO + (fi * 4 + fj) * 9;
#pragma unroll
      for (int c1 = 0; c1 < 3; c1++)
#pragma unroll
        for (int c2 = 0; c2 < 3; c2++) {
          double pv = 0.0;
          if (c1 == c2) pv = s22 * C.wp[c1] + s11 * C.wp[3 + c1];
          if (c1 == 2 && c2 == 0) pv += x20;
          if (c1 == 0 && c2 == 2) pv += x02;
          if (c1 == 0 && c2 == 0) pv += x00;
          double v = 2.0 * (s22 * pt[c1 * 3 + c2] + s11 * po[c1 * 3 + c2] + pv * C.m2);
          if (j1 == j2 && c1 == c2) { v += C.alpha2; dmax = fmax(dmax, v); }
          K[(3 * j1 + c1) * n + 3 * j2 + c2] = v;
          K[(3 * j2 + c2) * n + 3 * j1 + c1] = v;
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
  pc.tick(CMPC_PH_HESS);
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
  // Look-ahead pivots: a shared copy of the (scaled) diagonal is advanced alongside the tiles with the
  // very same fma, so every thread can form pivot k+1 and start its reciprocal while it is still
  // applying pivot k -- the MUFU + Newton chain leaves the per-pivot critical path.
  double* dg = cbuf + 2 * (NPAD + 2);  // two diagonal buffers of NPAD
  for (int i = tid; i < NPAD; i += NT) dg[i] = (i < n) ? K[i * n + i] * scale : 0.5;
  __syncthreads();
  pc.tick(CMPC_PH_LOAD);
  double dinv = fast_rcp(dg[0]);
  int par = 0;
#pragma unroll 1
  for (int k = 0; k < n; k++) {
    const int a = k / TY, kk = k - a * TY;  // row block / row-owner ty
    double* cb = cbuf + par * (NPAD + 2);
    const double* dold = dg + par * NPAD;
    double* dnew = dg + (par ^ 1) * NPAD;
    if (ty == kk) {
#pragma unroll
      for (int aa = 0; aa < TM; aa++) {
        if (aa == a) {
#pragma unroll
          for (int bb = 0; bb < TN; bb++) {
            const int j = tx + TX * bb;
            double v = A[aa][bb];
            if (j == k) v -= 1.0;
            cb[j] = v;
          }
        }
      }
    }
    __syncthreads();
    double ci[TM], cjd[TN];
#pragma unroll
    for (int aa = 0; aa < TM; aa++) ci[aa] = cb[ty + TY * aa];
#pragma unroll
    for (int bb = 0; bb < TN; bb++) cjd[bb] = cb[tx + TX * bb] * dinv;
    // next pivot and its reciprocal (same arithmetic as the tile update below)
    const int k1 = (k + 1 < n) ? k + 1 : k;
    const double c1 = cb[k1];
    const double dnext = fma(-c1, c1 * dinv, dold[k1]);
    const double dinv_next = fast_rcp(dnext);
    if (tid < NPAD) {
      const double c = cb[tid];
      dnew[tid] = fma(-c, c * dinv, dold[tid]);
    }
#pragma unroll
    for (int aa = 0; aa < TM; aa++)
#pragma unroll
      for (int bb = 0; bb < TN; bb++) A[aa][bb] = fma(-ci[aa], cjd[bb], A[aa][bb]);
    dinv = dinv_next;
    par ^= 1;
  }
  pc.tick(CMPC_PH_SWEEP);
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
