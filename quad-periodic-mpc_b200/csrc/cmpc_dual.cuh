// Kernel 2 of the two-kernel pipeline: the friction-pyramid / force-bound QP
//     min 1/2 x'Hx + g'x   s.t.  0 <= C x <= ub          (fmat rows, SolverMPC.cpp:660, solved by
// qpOASES in the reference, SolverMPC.cpp:955-964) by the Goldfarb-Idnani dual active-set method in
// range-space form on K = H^-1, ONE WARP PER INSTANCE.
//
// Kernel 1 left K, g, x0 = -K g and the contact list in the instance's workspace slot.  An iteration is
// a chain of short dependent vector operations, so a warp (no block barriers, reductions by shuffles)
// is the right grain, and many independent warps per SM hide each other's latency; warps are persistent
// and draw instances from a device-side counter, so a slow instance never holds a CTA-sized tail.
//
//   start at x = x0; repeat: pick the most violated row p outside the working set; fetch
//   kn = K n_p (two rows of K from the slot, L2); d = N' kn, r = P d with P = (N'KN)^-1 kept explicitly
//   (packed symmetric, bordered / deflated per step); z = kn - KN r where KN = K N is cached column by
//   column as rows enter the working set; step t = min(primal step that makes row p tight, largest dual
//   step keeping the multipliers >= 0); a full step adds p, a partial step drops the blocking row.
//
// H is strictly positive definite (alpha > 0), so the optimum is unique and the method terminates at it.
#pragma once

namespace {

struct DCarve {
  int x, kn, z, s, u, d, r, col, KN, Pp, act, isact, fs, gv, fsinv, total;
  CMPC_CANARY_FIELDS
};

__host__ __device__ inline DCarve make_dcarve(int nmax, int qcap) {
  DCarve c;
  int o = 0;
  CMPC_GUARD_INIT(c);
  const int m = 5 * (nmax / 3);
  c.x = o; o += align16(8 * nmax);
  CMPC_GUARD(o, c);
  c.kn = o; o += align16(8 * nmax);
  CMPC_GUARD(o, c);
  c.z = o; o += align16(8 * nmax);
  CMPC_GUARD(o, c);
  c.s = o; o += align16(8 * m);
  CMPC_GUARD(o, c);
  c.u = o; o += align16(8 * (qcap + 1));
  CMPC_GUARD(o, c);
  c.d = o; o += align16(8 * (qcap + 1));
  CMPC_GUARD(o, c);
  c.r = o; o += align16(8 * (qcap + 1));
  CMPC_GUARD(o, c);
  c.col = o; o += align16(8 * (qcap + 1));
  CMPC_GUARD(o, c);
  c.KN = o; o += align16(8 * qcap * nmax);
  CMPC_GUARD(o, c);
  c.Pp = o; o += align16(8 * ((qcap + 1) * (qcap + 2) / 2));
  CMPC_GUARD(o, c);
  c.act = o; o += align16(2 * (qcap + 1));
  CMPC_GUARD(o, c);
  c.isact = o; o += align16(m);
  CMPC_GUARD(o, c);
  c.fs = o; o += align16(CMPC_MAX_FS);
  CMPC_GUARD(o, c);
  c.gv = o; o += align16(CMPC_MAX_FS);
  CMPC_GUARD(o, c);
  c.fsinv = o; o += align16(CMPC_MAX_FS);
  CMPC_GUARD(o, c);
  c.total = o;
  return c;
}

__device__ __forceinline__ void warp_argmin(double& val, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, val, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov < val || (ov == val && oi < idx)) { val = ov; idx = oi; }
  }
}

// K[i][j] from a workspace slot: row-major n x n, or the lower-triangular 8x8 tiles of cmpc_condense_mma.cuh
__device__ __forceinline__ double k_entry(const double* slot, int n, bool tiled, int i, int j) {
  if (!tiled) return __ldg(slot + (size_t)i * n + j);
  const int It = i >> 3, ri = i & 7, Jt = j >> 3, cj = j & 7;
  const int off = (Jt <= It) ? (It * (It + 1) / 2 + Jt) * 64 + ri * 8 + cj : (Jt * (Jt + 1) / 2 + It) * 64 + cj * 8 + ri;
  return __ldg(slot + off);
}

__device__ __forceinline__ double warp_sum(double val) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
  return val;
}

}  // namespace

template <int WPC>
__global__ void __launch_bounds__(32 * WPC) cmpc_dual_kernel(const __grid_constant__ CmpcParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = P.horizon, nmax = P.nmax, qcap = P.qcap;
  const DCarve cv = make_dcarve(nmax, qcap);
  unsigned char* base = smem + (size_t)warp * cv.total;
  double* x = reinterpret_cast<double*>(base + cv.x);
  double* kn = reinterpret_cast<double*>(base + cv.kn);
  double* z = reinterpret_cast<double*>(base + cv.z);
  double* s = reinterpret_cast<double*>(base + cv.s);
  double* u = reinterpret_cast<double*>(base + cv.u);
  double* dvec = reinterpret_cast<double*>(base + cv.d);
  double* rvec = reinterpret_cast<double*>(base + cv.r);
  double* col = reinterpret_cast<double*>(base + cv.col);
  double* KN = reinterpret_cast<double*>(base + cv.KN);
  double* Pp = reinterpret_cast<double*>(base + cv.Pp);
  short* act = reinterpret_cast<short*>(base + cv.act);
  unsigned char* isact = base + cv.isact;
  unsigned char* fs = base + cv.fs;
  unsigned char* gv = base + cv.gv;
  signed char* fsinv = reinterpret_cast<signed char*>(base + cv.fsinv);

  const int count = P.count_ptr ? min(*P.count_ptr, P.count) : P.count;
  const double mu_inv = P.mu_inv;
  double flops_acc = 0.0;
#ifdef CMPC_CANARY
  canary_fill(base, cv.guard, cv.nguard, lane, 32);
  __syncwarp();
#endif

  while (true) {
    int slot_i = 0;
    if (lane == 0) slot_i = atomicAdd(P.sched, 1);
    slot_i = __shfl_sync(0xffffffffu, slot_i, 0);
    if (slot_i >= count) break;
    const int inst = P.worklist ? P.worklist[slot_i] : slot_i;
    const double* slot = P.qws + (size_t)inst * P.qws_stride;
    const double* gg = slot + P.qws_goff;
    const bool tiled = P.k_tiled != 0;
    const double* x0 = gg + nmax;
    const int* hdr = reinterpret_cast<const int*>(x0 + nmax + 2);
    const int nc = hdr[0];
    const int st0 = hdr[1];
    const unsigned char* hb = reinterpret_cast<const unsigned char*>(hdr + 2);
    const int n = 3 * nc, m = 5 * nc;
    const bool have = (st0 == CMPC_ST_SOLVED);

    for (int k = lane; k < 4 * h; k += 32) fsinv[k] = -1;
    __syncwarp();
    if (have) {
      for (int j = lane; j < nc; j += 32) {
        const unsigned char k = hb[j];
        fs[j] = k;
        gv[j] = hb[CMPC_MAX_FS + j];
        fsinv[k] = (signed char)j;
      }
      for (int i = lane; i < n; i += 32) x[i] = x0[i];
    }
    __syncwarp();
    int status = st0, iters = 0, q = 0;
    if (have) {
      for (int c = lane; c < m; c += 32) {
        int ia, iz;
        double va, vz;
        cons_of(c, mu_inv, ia, va, iz, vz);
        double b = 0.0;
        if (c % 5 == 4) b = -(double)gv[c / 5] * P.f_max;
        s[c] = va * x[ia] + vz * x[iz] - b;
        isact[c] = 0;
      }
      __syncwarp();
      // ---- resume: the first tier ran out of working-set capacity at a valid state of the method (x optimal on the
      //      face of its working set A, multipliers >= 0).  Rebuild that state from A alone: border P and cache K N
      //      row by row (no line searches), then u = -P s_A(x0), x = x0 + K N u ----
      if (P.resume_in && P.worklist) {
        const int* rs = P.resume_in + (size_t)slot_i * CMPC_RESUME_INTS;
        const int rq = min(rs[0] & 0xffff, qcap);
        const int pfmt = (rs[0] >> 16) & 3;  // 0: row ids only, 1: P as q x q rows, 2: P packed (lower triangle)
        const bool have_p = pfmt != 0 && P.rstate_in && slot_i < P.rstate_in_cap && rq == (rs[0] & 0xffff);
        if (have_p) {
          // the working set with its P: cache K N for every row (independent row fetches), copy P, done
          const double* stp = P.rstate_in + (size_t)slot_i * P.rstate_in_stride;
          for (int k = lane; k < rq; k += 32) {
            const int w2 = rs[2 + (k >> 1)];
            const int p = (k & 1) ? ((w2 >> 16) & 0xffff) : (w2 & 0xffff);
            act[k] = (short)p;
            isact[p] = 1;
          }
          __syncwarp();
          for (int k = 0; k < rq; k++) {
            int pia, piz;
            double pva, pvz;
            cons_of(act[k], mu_inv, pia, pva, piz, pvz);
            for (int i = lane; i < n; i += 32)
              KN[k * nmax + i] = pva * k_entry(slot, n, tiled, pia, i) + pvz * k_entry(slot, n, tiled, piz, i);
          }
          if (pfmt == 1) {
            for (int k = 0; k < rq; k++)
              for (int l = lane; l <= k; l += 32) Pp[k * (k + 1) / 2 + l] = stp[k * rq + l];
          } else {
            for (int e = lane; e < rq * (rq + 1) / 2; e += 32) Pp[e] = stp[e];
          }
          q = rq;
          __syncwarp();
        }
        for (int k = have_p ? rq : 0; k < rq; k++) {  // row ids only: border P and cache K N row by row
          const int w2 = rs[2 + (k >> 1)];
          const int p = (k & 1) ? ((w2 >> 16) & 0xffff) : (w2 & 0xffff);
          int pia, piz;
          double pva, pvz;
          cons_of(p, mu_inv, pia, pva, piz, pvz);
          for (int i = lane; i < n; i += 32)
            kn[i] = pva * k_entry(slot, n, tiled, pia, i) + pvz * k_entry(slot, n, tiled, piz, i);
          __syncwarp();
          const double scale = pva * kn[pia] + pvz * kn[piz];
          for (int l = lane; l < q; l += 32) {
            int ia, iz;
            double va, vz;
            cons_of(act[l], mu_inv, ia, va, iz, vz);
            dvec[l] = va * kn[ia] + vz * kn[iz];
          }
          __syncwarp();
          double dr = 0.0;
          for (int l = lane; l < q; l += 32) {
            double acc = 0.0;
            for (int j = 0; j < q; j++) acc = fma(psym(Pp, l, j), dvec[j], acc);
            rvec[l] = acc;
            dr = fma(dvec[l], acc, dr);
          }
          dr = warp_sum(dr);
          __syncwarp();
          const double rho2_inv = fast_rcp(scale - dr);
          for (int l = lane; l < q; l += 32) {
            const double rk = rvec[l] * rho2_inv;
            for (int j = 0; j <= l; j++) Pp[l * (l + 1) / 2 + j] = fma(rk, rvec[j], Pp[l * (l + 1) / 2 + j]);
            Pp[q * (q + 1) / 2 + l] = -rk;
          }
          for (int i = lane; i < n; i += 32) KN[q * nmax + i] = kn[i];
          if (lane == 0) {
            Pp[q * (q + 1) / 2 + q] = rho2_inv;
            act[q] = (short)p;
            isact[p] = 1;
          }
          q++;
          __syncwarp();
        }
        if (rq > 0) {
          for (int k = lane; k < q; k += 32) {
            double acc = 0.0;
            for (int l = 0; l < q; l++) acc = fma(psym(Pp, k, l), s[act[l]], acc);
            u[k] = fmax(-acc, 0.0);
          }
          __syncwarp();
          for (int i = lane; i < n; i += 32) {
            double acc = x[i];
            for (int k = 0; k < q; k++) acc = fma(u[k], KN[k * nmax + i], acc);
            x[i] = acc;
          }
          __syncwarp();
          for (int c = lane; c < m; c += 32) {
            int ia, iz;
            double va, vz;
            cons_of(c, mu_inv, ia, va, iz, vz);
            double b = 0.0;
            if (c % 5 == 4) b = -(double)gv[c / 5] * P.f_max;
            s[c] = va * x[ia] + vz * x[iz] - b;
          }
          iters = rs[1];
          flops_acc += 2.0 * (double)rq * ((double)q * q + 2.0 * n);
          __syncwarp();
        }
      }
      bool done = false;
      while (!done) {
        // most violated row outside the working set
        double best = 1e300;
        int bidx = -1;
        for (int c = lane; c < m; c += 32)
          if (!isact[c]) { const double sv = s[c]; if (sv < best) { best = sv; bidx = c; } }
        warp_argmin(best, bidx);
        if (!(best < -P.tol_violation)) break;
        const int p = bidx;
        int pia, piz;
        double pva, pvz;
        cons_of(p, mu_inv, pia, pva, piz, pvz);
        for (int i = lane; i < n; i += 32)
          kn[i] = pva * k_entry(slot, n, tiled, pia, i) + pvz * k_entry(slot, n, tiled, piz, i);
        __syncwarp();
        const double scale = pva * kn[pia] + pvz * kn[piz];
        double up = 0.0;
        while (true) {
          iters++;
          if (iters > P.max_iter) { status = CMPC_ST_MAXITER; done = true; break; }
          for (int k = lane; k < q; k += 32) {
            int ia, iz;
            double va, vz;
            cons_of(act[k], mu_inv, ia, va, iz, vz);
            dvec[k] = va * kn[ia] + vz * kn[iz];
          }
          __syncwarp();
          double dr = 0.0, ratio = 1e300;
          int kd = -1;
          for (int k = lane; k < q; k += 32) {
            double acc = 0.0;
            for (int l = 0; l < q; l++) acc = fma(psym(Pp, k, l), dvec[l], acc);
            rvec[k] = acc;
            dr = fma(dvec[k], acc, dr);
            if (acc > 0.0) { const double t = u[k] / acc; if (t < ratio) { ratio = t; kd = k; } }
          }
          dr = warp_sum(dr);
          warp_argmin(ratio, kd);
          __syncwarp();
          const double rho2 = scale - dr;
          const bool dependent = !(rho2 > 1e-12 * scale);
          if (!dependent) {
            for (int i = lane; i < n; i += 32) {
              double acc = kn[i];
              for (int k = 0; k < q; k++) acc = fma(-rvec[k], KN[k * nmax + i], acc);
              z[i] = acc;
            }
            __syncwarp();
          }
          const double rho2_inv = dependent ? 0.0 : fast_rcp(rho2);
          const double t2 = dependent ? 1e300 : -s[p] * rho2_inv;
          const double t1 = ratio;
          const double t = fmin(t1, t2);
          if (t >= 1e299) { status = CMPC_ST_INFEASIBLE; done = true; break; }
          const bool full = (t2 <= t1);
          __syncwarp();  // every lane has read s[p] before the slacks move
          if (!dependent) {
            for (int i = lane; i < n; i += 32) x[i] = fma(t, z[i], x[i]);
            for (int c = lane; c < m; c += 32) {
              int ia, iz;
              double va, vz;
              cons_of(c, mu_inv, ia, va, iz, vz);
              s[c] = fma(t, va * z[ia] + vz * z[iz], s[c]);
            }
          }
          for (int k = lane; k < q; k += 32) u[k] = fma(-t, rvec[k], u[k]);
          up += t;
          flops_acc += 2.0 * (4.0 * n + 4.0 * q + (double)q * q + (double)n * q + 4.0 * m + n);
          if (full) {
            if (q >= qcap) { status = CMPC_ST_WSOVERFLOW; done = true; break; }
            // border P with the new row: [P + r r'/rho2, -r/rho2; -r'/rho2, 1/rho2]; cache K n_p
            for (int k = lane; k < q; k += 32) {
              const double rk = rvec[k] * rho2_inv;
              for (int l = 0; l <= k; l++) Pp[k * (k + 1) / 2 + l] = fma(rk, rvec[l], Pp[k * (k + 1) / 2 + l]);
              Pp[q * (q + 1) / 2 + k] = -rk;
            }
            for (int i = lane; i < n; i += 32) KN[q * nmax + i] = kn[i];
            if (lane == 0) {
              Pp[q * (q + 1) / 2 + q] = rho2_inv;
              act[q] = (short)p;
              u[q] = up;
              isact[p] = 1;
            }
            q++;
            flops_acc += 2.0 * (double)q * q;
            __syncwarp();
            break;
          }
          // partial step: row kd leaves the working set (P deflated by its row/column), p stays the candidate
          for (int k = lane; k < q; k += 32) col[k] = psym(Pp, k, kd);
          __syncwarp();
          {
            const double inv = 1.0 / col[kd];
            for (int k = lane; k < q; k += 32) {
              if (k == kd) continue;
              const double ck = col[k] * inv;
              for (int l = 0; l <= k; l++)
                if (l != kd) Pp[k * (k + 1) / 2 + l] = fma(-ck, col[l], Pp[k * (k + 1) / 2 + l]);
            }
          }
          __syncwarp();
          const int last = q - 1;
          if (kd != last) {
            for (int l = lane; l < last; l += 32)
              if (l != kd) psym(Pp, kd, l) = psym(Pp, last, l);
            for (int i = lane; i < n; i += 32) KN[kd * nmax + i] = KN[last * nmax + i];
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
          __syncwarp();
        }
      }
    }
    if (status == CMPC_ST_WSOVERFLOW && P.overflow_list) {
      // left for the next capacity tier, with the working set reached here (up to 64 rows travel as 16-bit ids)
      int pos = 0;
      if (lane == 0) {
        pos = atomicAdd(P.overflow_count, 1);
        P.overflow_list[pos] = inst;
      }
      pos = __shfl_sync(0xffffffffu, pos, 0);
      if (P.resume_out) {
        int* rs = P.resume_out + (size_t)pos * CMPC_RESUME_INTS;
        const int qs = min(q, 2 * (CMPC_RESUME_INTS - 2));
        for (int k = lane; 2 * k < qs; k += 32) {
          const int lo = (unsigned short)act[2 * k], hi = (2 * k + 1 < qs) ? (unsigned short)act[2 * k + 1] : 0;
          rs[2 + k] = lo | (hi << 16);
        }
        bool with_p = false;
        if (P.rstate_out && pos < P.rstate_out_cap && qs == q && q * (q + 1) / 2 <= P.rstate_out_stride) {
          double* stp = P.rstate_out + (size_t)pos * P.rstate_out_stride;
          for (int e = lane; e < q * (q + 1) / 2; e += 32) stp[e] = Pp[e];
          with_p = true;
        }
        if (lane == 0) { rs[0] = qs | (with_p ? (2 << 16) : 0); rs[1] = iters - 1; }
      }
      __syncwarp();
      continue;
    }
    // ---- outputs: q_soln scatter (zeros for swing feet), objective, primal activity mask ----
    bool fin = true;  // a non-finite iterate (NaN / Inf upstream) reports CMPC_ST_NONFINITE and zero forces
    if (have)
      for (int i = lane; i < n; i += 32) fin = fin && isfinite(x[i]);
    fin = __all_sync(0xffffffffu, fin);
    if (have && !fin) status = CMPC_ST_NONFINITE;
    const bool have_x = have && fin;
    if (P.forces) {
      double* out = P.forces + (size_t)inst * 12 * h;
      for (int idx = lane; idx < 12 * h; idx += 32) {
        const int k = idx / 3, comp = idx - 3 * k;
        const int j = fsinv[k];
        out[idx] = (j >= 0 && have_x) ? x[3 * j + comp] : 0.0;
      }
    }
    if (P.active) {
      signed char* out = P.active + (size_t)inst * 20 * h;
      for (int idx = lane; idx < 20 * h; idx += 32) {
        const int k = idx / 5, t = idx - 5 * k;
        const int j = fsinv[k];
        signed char a = 0;
        if (j >= 0 && have_x) {
          const double fx = x[3 * j], fy = x[3 * j + 1], fz = x[3 * j + 2];
          const double row = (t == 0) ? fx * mu_inv + fz : (t == 1) ? -fx * mu_inv + fz : (t == 2) ? fy * mu_inv + fz
                           : (t == 3) ? -fy * mu_inv + fz : fz;
          if (row <= P.tol_active) a = -1;
          if (t == 4 && row >= (double)gv[j] * P.f_max - P.tol_active) a = 1;
        }
        out[idx] = a;
      }
    }
    {
      // objective 0.5 x'Hx + g'x = 0.5 g'x + 0.5 lambda'b at a KKT point
      double part = 0.0;
      if (have_x) {
        for (int i = lane; i < n; i += 32) part = fma(0.5 * __ldg(gg + i), x[i], part);
        for (int k = lane; k < q; k += 32) {
          const int c = act[k];
          if (c % 5 == 4) part -= 0.5 * u[k] * (double)gv[c / 5] * P.f_max;
        }
      }
      part = warp_sum(part);
      if (lane == 0) {
        if (P.objective) P.objective[inst] = have_x ? part : 0.0;
        if (P.status) P.status[inst] = status;
        if (P.iterations) P.iterations[inst] = iters;
      }
    }
    __syncwarp();
  }
#ifdef CMPC_CANARY
  __syncwarp();
  canary_check(base, cv.guard, cv.nguard, lane, 32, "cmpc_dual_kernel");
#endif
  if (lane == 0 && P.flops && flops_acc > 0.0) atomicAdd(P.flops + CMPC_K_DUAL, (unsigned long long)flops_acc);
}
