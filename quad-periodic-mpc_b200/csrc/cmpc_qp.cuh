// Stage E of the solve kernel: the friction-cone / force-bound QP
//     min 1/2 x'Hx + g'x   s.t.  0 <= C x <= ub          (fmat rows, SolverMPC.cpp:660, solved by
// qpOASES in the reference, SolverMPC.cpp:955-964) by the Goldfarb-Idnani dual active-set method in
// range-space form on K = H^-1 (stage D):
//
//   start at the unconstrained minimiser x = -K g; repeat: pick the most violated row p;
//   z = K (n_p - N r),  r = (N'KN)^-1 N'K n_p   (N = normals of the working set);
//   step t = min(primal step to make row p tight, largest dual step keeping multipliers >= 0);
//   a full step adds p to the working set, a partial step drops the blocking row and retries p.
//
// H is strictly positive definite, so the optimum is unique and the method reaches it in finitely
// many steps; rows are sparse (two entries), so every product with K touches two columns.  The
// inverse P of the working set's Schur complement N'KN is kept explicitly (packed symmetric) and
// bordered / deflated per step.
//
// The stage is run by the first GT threads of the CTA (GT = padded problem size: one variable per
// thread) on a named barrier: an iteration is a chain of short dependent vector operations, so
// wider groups only add redundant control instructions and barrier latency.
#pragma once

namespace {

template <int GT>
__device__ __forceinline__ void gsync() {
  if (GT <= 32) __syncwarp();
  else asm volatile("bar.sync 1, %0;" ::"n"(GT) : "memory");
}

template <int GT>
__device__ __forceinline__ void group_argmin(double& val, int& idx, double* red, int t) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, val, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov < val || (ov == val && oi < idx)) { val = ov; idx = oi; }
  }
  if (GT > 32) {
    int* redi = reinterpret_cast<int*>(red + 8);
    gsync<GT>();
    if ((t & 31) == 0) { red[t >> 5] = val; redi[t >> 5] = idx; }
    gsync<GT>();
    val = red[0]; idx = redi[0];
#pragma unroll
    for (int w = 1; w < GT / 32; w++) {
      double ov = red[w]; int oi = redi[w];
      if (ov < val || (ov == val && oi < idx)) { val = ov; idx = oi; }
    }
  }
}

template <int GT>
__device__ __forceinline__ double group_sum(double val, double* red, int t) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
  if (GT > 32) {
    gsync<GT>();
    if ((t & 31) == 0) red[16 + (t >> 5)] = val;
    gsync<GT>();
    val = red[16];
#pragma unroll
    for (int w = 1; w < GT / 32; w++) val += red[16 + w];
  }
  return val;
}

struct QpState {
  int status, iters, q;
};

// lane = 0..GT-1: index within the executing group.  All arrays are shared (or workspace) memory.
template <int GT>
__device__ __forceinline__ QpState qp_dual_active_set(int lane, int n, int m, double mu_inv, double tol_violation,
                                                      int max_iter, int qcap, const double* K, double* x, double* s,
                                                      double* rc, unsigned char* isact, short* act, double* u,
                                                      double* dvec, double* rvec, double* col, double* Pp, double* kn,
                                                      double* z, double* vv, double* red, double& flops_acc) {
  QpState st;
  st.status = CMPC_ST_SOLVED;
  st.iters = 0;
  int q = 0;
  bool done = false;
  while (!done) {
    // most violated row outside the working set
    double best = 1e300;
    int bidx = -1;
    for (int c = lane; c < m; c += GT)
      if (!isact[c]) { double sv = s[c]; if (sv < best) { best = sv; bidx = c; } }
    group_argmin<GT>(best, bidx, red, lane);
    if (!(best < -tol_violation)) break;
    const int p = bidx;
    int pia, piz;
    double pva, pvz;
    cons_of(p, mu_inv, pia, pva, piz, pvz);
    double up = 0.0;
    while (true) {
      st.iters++;
      if (st.iters > max_iter) { st.status = CMPC_ST_MAXITER; done = true; break; }
      for (int i = lane; i < n; i += GT) kn[i] = pva * K[pia * n + i] + pvz * K[piz * n + i];
      gsync<GT>();
      const double scale = pva * kn[pia] + pvz * kn[piz];
      for (int k = lane; k < q; k += GT) {
        int ia, iz;
        double va, vz;
        cons_of(act[k], mu_inv, ia, va, iz, vz);
        dvec[k] = va * kn[ia] + vz * kn[iz];
      }
      gsync<GT>();
      double dr = 0.0, ratio = 1e300;
      int kd = -1;
      for (int k = lane; k < q; k += GT) {
        double acc = 0.0;
        for (int l = 0; l < q; l++) acc = fma(psym(Pp, k, l), dvec[l], acc);
        rvec[k] = acc;
        rc[act[k]] = acc;
        dr = fma(dvec[k], acc, dr);
        if (acc > 0.0) { double t = u[k] / acc; if (t < ratio) { ratio = t; kd = k; } }
      }
      dr = group_sum<GT>(dr, red, lane);
      group_argmin<GT>(ratio, kd, red, lane);
      gsync<GT>();
      const double rho2 = scale - dr;
      const bool dependent = !(rho2 > 1e-12 * scale);
      if (!dependent) {
        // v = N r gathered per variable from the (at most five) rows of its foot-step
        for (int i = lane; i < n; i += GT) {
          int j = i / 3, comp = i - 3 * j;
          const double* rj = rc + 5 * j;
          double val;
          if (comp == 0) val = mu_inv * (rj[0] - rj[1]);
          else if (comp == 1) val = mu_inv * (rj[2] - rj[3]);
          else val = rj[0] + rj[1] + rj[2] + rj[3] - rj[4];
          vv[i] = val;
        }
        gsync<GT>();
        for (int i = lane; i < n; i += GT) {
          double acc = kn[i];
          for (int l = 0; l < n; l++) {
            double vl = vv[l];
            if (vl != 0.0) acc = fma(-K[l * n + i], vl, acc);
          }
          z[i] = acc;
        }
        gsync<GT>();
      }
      const double rho2_inv = dependent ? 0.0 : fast_rcp(rho2);
      const double t2 = dependent ? 1e300 : -s[p] * rho2_inv;
      const double t1 = ratio;
      const double t = fmin(t1, t2);
      if (t >= 1e299) { st.status = CMPC_ST_INFEASIBLE; done = true; break; }
      const bool full = (t2 <= t1);
      gsync<GT>();  // every lane has read s[p] before the slacks move
      if (!dependent) {
        for (int i = lane; i < n; i += GT) x[i] = fma(t, z[i], x[i]);
        for (int c = lane; c < m; c += GT) {
          int ia, iz;
          double va, vz;
          cons_of(c, mu_inv, ia, va, iz, vz);
          s[c] = fma(t, va * z[ia] + vz * z[iz], s[c]);
        }
      }
      for (int k = lane; k < q; k += GT) {
        u[k] = fma(-t, rvec[k], u[k]);
        rc[act[k]] = 0.0;
      }
      up += t;
      flops_acc += 2.0 * (4.0 * n + 4.0 * q + (double)q * q + 3.0 * n * (2.0 * q < n ? 2.0 * q : (double)n) + 4.0 * m + n);
      if (full) {
        if (q >= qcap) { st.status = CMPC_ST_WSOVERFLOW; done = true; break; }
        // border P with the new row: [P + r r'/rho2, -r/rho2; -r'/rho2, 1/rho2]
        for (int k = lane; k < q; k += GT) {
          double rk = rvec[k] * rho2_inv;
          for (int l = 0; l <= k; l++) Pp[k * (k + 1) / 2 + l] = fma(rk, rvec[l], Pp[k * (k + 1) / 2 + l]);
          Pp[q * (q + 1) / 2 + k] = -rk;
        }
        if (lane == 0) {
          Pp[q * (q + 1) / 2 + q] = rho2_inv;
          act[q] = (short)p;
          u[q] = up;
          isact[p] = 1;
        }
        q++;
        flops_acc += 2.0 * (double)q * q;
        gsync<GT>();
        break;
      }
      // partial step: row kd leaves the working set (P deflated by its row/column), p stays the candidate
      for (int k = lane; k < q; k += GT) col[k] = psym(Pp, k, kd);
      gsync<GT>();
      {
        const double inv = 1.0 / col[kd];
        for (int k = lane; k < q; k += GT) {
          if (k == kd) continue;
          double ck = col[k] * inv;
          for (int l = 0; l <= k; l++)
            if (l != kd) Pp[k * (k + 1) / 2 + l] = fma(-ck, col[l], Pp[k * (k + 1) / 2 + l]);
        }
      }
      gsync<GT>();
      const int last = q - 1;
      if (kd != last) {
        for (int l = lane; l < last; l += GT)
          if (l != kd) psym(Pp, kd, l) = psym(Pp, last, l);
        if (lane == 0) {
          Pp[kd * (kd + 1) / 2 + kd] = Pp[last * (last + 1) / 2 + last];
          isact[act[kd]] = 0;
          act[kd] = act[last];
          u[kd] = u[last];
        }
      } else if (lane == 0) {
        isact[act[kd]] = 0;
      }
      q--;
      flops_acc += 2.0 * (double)q * q;
      gsync<GT>();
    }
  }
  st.q = q;
  return st;
}

}  // namespace
