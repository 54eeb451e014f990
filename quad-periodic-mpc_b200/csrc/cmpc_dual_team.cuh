// Dual active-set kernel for LONG working sets: ONE CTA (four warps; eight for reduced problems beyond 64 variables)
// PER INSTANCE.
//
// Same method and the same arithmetic per step as cmpc_dual.cuh (Goldfarb-Idnani in range-space form on K = H^-1,
// P = (N'KN)^-1 kept explicitly, K N cached row by row); what changes is the grain.  The first tier
// (cmpc_dual_fast.cuh) gives every instance one warp and up to 32 working-set rows; an instance that outgrows it is
// handed over with its working set (and P), and from there an iteration is dominated by O(q^2) and O(n q) loops —
// P d, z = K n_p - K N r, the border / deflation of P — that one warp walks serially (2 - 5 us per iteration at 40 - 60
// rows) while K N and P of that single instance fill a third of an SM's shared memory, so hardly any other warp can
// be resident to hide it.  Here the 128 threads of the CTA split those loops:
//   P d            one working-set row per group of 1, 2 or 4 adjacent lanes (as many as 128 / q allows), partial dot
//                  products folded by shuffles; P is a full (not packed) square with an odd row stride;
//   z              one variable per thread, four accumulators over the rows of K N;
//   border/deflate every (row, column) pair of P, 2-D over the CTA;
//   argmin / ratio block reductions: warp partials by REDUX / shuffles, one barrier, every thread folds the partials;
//   K rows      the two rows of K of the next candidate are requested right after the slacks have moved and travel
//                  from L2 while P is bordered.
// Shared memory per instance is what the one-warp kernel needed, so the same number of instances is resident per SM
// with four times the threads working on them.
#pragma once

namespace {

struct TCarve {
  int x, kn, z, s, u, d, r, col, KN, Pm, act, isact, fs, gv, fsinv, red, bc, total, psq;
  CMPC_CANARY_FIELDS
};

__host__ __device__ inline TCarve make_tcarve(int nmax, int qcap) {
  TCarve c;
  int o = 0;
  CMPC_GUARD_INIT(c);
  const int m = 5 * (nmax / 3);
  c.psq = (qcap + 1) | 1;
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
  c.d = o; o += align16(8 * (qcap + 4));
  CMPC_GUARD(o, c);
  c.r = o; o += align16(8 * (qcap + 4));
  CMPC_GUARD(o, c);
  c.col = o; o += align16(8 * (qcap + 1));
  CMPC_GUARD(o, c);
  c.KN = o; o += align16(8 * qcap * nmax);
  CMPC_GUARD(o, c);
  c.Pm = o; o += align16(8 * (qcap + 1) * c.psq);
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
  c.red = o; o += 8 * 48;
  CMPC_GUARD(o, c);
  c.bc = o; o += 64;
  CMPC_GUARD(o, c);
  c.total = o;
  return c;
}

}  // namespace

template <int NT>
__global__ void __launch_bounds__(NT) cmpc_dual_team_kernel(const __grid_constant__ CmpcParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  const int h = P.horizon, nmax = P.nmax, qcap = P.qcap;
  const TCarve cv = make_tcarve(nmax, qcap);
  const int PSQ = cv.psq;
  double* x = reinterpret_cast<double*>(smem + cv.x);
  double* kn = reinterpret_cast<double*>(smem + cv.kn);
  double* z = reinterpret_cast<double*>(smem + cv.z);
  double* s = reinterpret_cast<double*>(smem + cv.s);
  double* u = reinterpret_cast<double*>(smem + cv.u);
  double* dvec = reinterpret_cast<double*>(smem + cv.d);
  double* rvec = reinterpret_cast<double*>(smem + cv.r);
  double* col = reinterpret_cast<double*>(smem + cv.col);
  double* KN = reinterpret_cast<double*>(smem + cv.KN);
  double* Pm = reinterpret_cast<double*>(smem + cv.Pm);
  short* act = reinterpret_cast<short*>(smem + cv.act);
  unsigned char* isact = smem + cv.isact;
  unsigned char* fs = smem + cv.fs;
  unsigned char* gv = smem + cv.gv;
  signed char* fsinv = reinterpret_cast<signed char*>(smem + cv.fsinv);
  double* red = reinterpret_cast<double*>(smem + cv.red);
  int* bc = reinterpret_cast<int*>(smem + cv.bc);

  const int count = P.count_ptr ? min(*P.count_ptr, P.count) : P.count;
  const double mu_inv = P.mu_inv;
  double flops_acc = 0.0;
#ifdef CMPC_CANARY
  canary_fill(smem, cv.guard, cv.nguard, tid, NT);
#endif

  while (true) {
    __syncthreads();  // the previous instance's arrays (and bc) are free
    if (tid == 0) bc[0] = atomicAdd(P.sched, 1);
    __syncthreads();
    const int slot_i = bc[0];
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

    for (int k = tid; k < 4 * h; k += NT) fsinv[k] = -1;
    __syncthreads();
    if (have) {
      for (int j = tid; j < nc; j += NT) {
        const unsigned char k = hb[j];
        fs[j] = k;
        gv[j] = hb[CMPC_MAX_FS + j];
        fsinv[k] = (signed char)j;
      }
      for (int i = tid; i < n; i += NT) x[i] = x0[i];
    }
    __syncthreads();
    int status = st0, iters = 0, q = 0;
    if (have) {
      for (int c = tid; c < m; c += NT) {
        int ia, iz;
        double va, vz;
        cons_of(c, mu_inv, ia, va, iz, vz);
        double b = 0.0;
        if (c % 5 == 4) b = -(double)gv[c / 5] * P.f_max;
        s[c] = va * x[ia] + vz * x[iz] - b;
        isact[c] = 0;
      }
      __syncthreads();
      // ---- resume: the tier before ran out of working-set capacity at a valid state of the method (x optimal on the
      //      face of its working set A, multipliers >= 0).  Rebuild that state from A: K N for every row, P as handed
      //      over (or bordered row by row when it did not travel), then u = -P s_A(x0), x = x0 + K N u ----
      if (P.resume_in && P.worklist) {
        const int* rs = P.resume_in + (size_t)slot_i * CMPC_RESUME_INTS;
        const int rq = min(rs[0] & 0xffff, qcap);
        const int pfmt = (rs[0] >> 16) & 3;  // 0: row ids only, 1: P as q x q rows, 2: P packed (lower triangle)
        const bool have_p = pfmt != 0 && P.rstate_in && slot_i < P.rstate_in_cap && rq == (rs[0] & 0xffff);
        for (int k = tid; k < rq; k += NT) {
          const int w2 = rs[2 + (k >> 1)];
          const int p = (k & 1) ? ((w2 >> 16) & 0xffff) : (w2 & 0xffff);
          act[k] = (short)p;
          if (have_p) isact[p] = 1;
        }
        __syncthreads();
        if (have_p) {
          const double* stp = P.rstate_in + (size_t)slot_i * P.rstate_in_stride;
          // K N: every (row, variable) pair is an independent pair of loads
          for (int e = tid; e < rq * n; e += NT) {
            const int k = e / n, i = e - k * n;
            int pia, piz;
            double pva, pvz;
            cons_of(act[k], mu_inv, pia, pva, piz, pvz);
            KN[k * nmax + i] = pva * k_entry(slot, n, tiled, pia, i) + pvz * k_entry(slot, n, tiled, piz, i);
          }
          for (int k = warp; k < rq; k += NW)
            for (int l = lane; l < rq; l += 32)
              Pm[k * PSQ + l] = (pfmt == 1) ? stp[k * rq + l] : (k >= l ? stp[k * (k + 1) / 2 + l] : stp[l * (l + 1) / 2 + k]);
          q = rq;
          __syncthreads();
        }
        for (int kk = have_p ? rq : 0; kk < rq; kk++) {  // row ids only: border P and cache K N row by row
          const int p = act[kk];
          int pia, piz;
          double pva, pvz;
          cons_of(p, mu_inv, pia, pva, piz, pvz);
          for (int i = tid; i < n; i += NT)
            kn[i] = pva * k_entry(slot, n, tiled, pia, i) + pvz * k_entry(slot, n, tiled, piz, i);
          __syncthreads();
          const double scale = pva * kn[pia] + pvz * kn[piz];
          for (int l = tid; l < q; l += NT) {
            int ia, iz;
            double va, vz;
            cons_of(act[l], mu_inv, ia, va, iz, vz);
            dvec[l] = va * kn[ia] + vz * kn[iz];
          }
          __syncthreads();
          double dr = 0.0;
          for (int l = tid; l < q; l += NT) {
            double acc = 0.0;
            for (int j = 0; j < q; j++) acc = fma(Pm[l * PSQ + j], dvec[j], acc);
            rvec[l] = acc;
            dr = fma(dvec[l], acc, dr);
          }
          dr = block_sum<NT>(dr, red, tid);
          __syncthreads();
          const double rho2_inv = fast_rcp(scale - dr);
          for (int k = warp; k < q; k += NW) {
            const double rk = rvec[k] * rho2_inv;
            for (int l = lane; l < q; l += 32) Pm[k * PSQ + l] = fma(rk, rvec[l], Pm[k * PSQ + l]);
          }
          for (int l = tid; l < q; l += NT) {
            const double rk = rvec[l] * rho2_inv;
            Pm[q * PSQ + l] = -rk;
            Pm[l * PSQ + q] = -rk;
          }
          for (int i = tid; i < n; i += NT) KN[q * nmax + i] = kn[i];
          if (tid == 0) {
            Pm[q * PSQ + q] = rho2_inv;
            isact[p] = 1;
          }
          q++;
          __syncthreads();
        }
        if (rq > 0) {
          for (int k = tid; k < q; k += NT) {
            double acc = 0.0;
            for (int l = 0; l < q; l++) acc = fma(Pm[k * PSQ + l], s[act[l]], acc);
            u[k] = fmax(-acc, 0.0);
          }
          __syncthreads();
          for (int i = tid; i < n; i += NT) {
            double acc = x[i];
            for (int k = 0; k < q; k++) acc = fma(u[k], KN[k * nmax + i], acc);
            x[i] = acc;
          }
          __syncthreads();
          for (int c = tid; c < m; c += NT) {
            int ia, iz;
            double va, vz;
            cons_of(c, mu_inv, ia, va, iz, vz);
            double b = 0.0;
            if (c % 5 == 4) b = -(double)gv[c / 5] * P.f_max;
            s[c] = va * x[ia] + vz * x[iz] - b;
          }
          iters = rs[1];
          flops_acc += 2.0 * (double)rq * ((double)q * q + 2.0 * n);
          __syncthreads();
        }
      }
      // ---- iterations.  Block-wide reductions cost ONE barrier each (warp partials by REDUX / shuffles, every thread
      //      folds the NW partials itself); the two rows of K of the NEXT candidate are requested from L2 as soon as the
      //      slacks have moved and travel while P is bordered ----
      int rpar = 0;  // reduction scratch alternates between two halves: a slow reader never meets the next writer
      auto team_argmin = [&](double& v, int& i) {
        warp_argmin_redux(v, i);
        double* rv = red + 24 * rpar;  // [0..7] values, [8..11] eight indices, [12..19] partial sums
        int* ri = reinterpret_cast<int*>(rv + 8);
        if (lane == 0) { rv[warp] = v; ri[warp] = i; }
        __syncthreads();
        v = rv[0];
        i = ri[0];
#pragma unroll
        for (int w = 1; w < NW; w++) {
          const double ov = rv[w];
          const int oi = ri[w];
          if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
        rpar ^= 1;
      };
      auto candidate = [&](double& best, int& bidx) {
        best = 1e300;
        bidx = 0x7fffffff;
        for (int c = tid; c < m; c += NT)
          if (!isact[c]) { const double sv = s[c]; if (sv < best) { best = sv; bidx = c; } }
        team_argmin(best, bidx);
      };
      double best;
      int p;
      candidate(best, p);
      bool done = !(best < -P.tol_violation);
      double ka = 0.0, kz = 0.0;  // this thread's entries of the candidate's two rows of K (n <= 128 <= NT)
      int pia = 0, piz = 0;
      double pva = 0.0, pvz = 0.0;
      if (!done) {
        cons_of(p, mu_inv, pia, pva, piz, pvz);
        if (tid < n) { ka = k_entry(slot, n, tiled, pia, tid); kz = k_entry(slot, n, tiled, piz, tid); }
      }
      while (!done) {
        if (tid < n) kn[tid] = pva * ka + pvz * kz;
        __syncthreads();
        const double scale = pva * kn[pia] + pvz * kn[piz];
        double up = 0.0;
        while (true) {
          iters++;
          if (iters > P.max_iter) { status = CMPC_ST_MAXITER; done = true; break; }
          const double sp = s[p];  // read before the slacks move
          for (int k = tid; k < q; k += NT) {
            int ia, iz;
            double va, vz;
            cons_of(act[k], mu_inv, ia, va, iz, vz);
            dvec[k] = va * kn[ia] + vz * kn[iz];
          }
          __syncthreads();
          // r = P d: a row per group of `parts` adjacent lanes (as many as NT / q allows); dr = d'r; ratio test
          const int parts = (4 * q <= NT) ? 4 : ((2 * q <= NT) ? 2 : 1);
          double dr = 0.0, ratio = 1e300;
          int kd = 0x7fffffff;
          for (int k0 = 0; k0 < q; k0 += NT / parts) {
            const int k = k0 + tid / parts, part = tid % parts;
            double a0 = 0.0, a1 = 0.0;
            if (k < q) {
              const double* prow = Pm + k * PSQ;
              int l = part;
              for (; l + parts < q; l += 2 * parts) {
                a0 = fma(prow[l], dvec[l], a0);
                a1 = fma(prow[l + parts], dvec[l + parts], a1);
              }
              if (l < q) a0 = fma(prow[l], dvec[l], a0);
            }
            double acc = a0 + a1;
            if (parts >= 2) acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            if (parts >= 4) acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (k < q && part == 0) {
              rvec[k] = acc;
              dr = fma(dvec[k], acc, dr);
              if (acc > 0.0) { const double t = u[k] * fast_rcp(acc); if (t < ratio) { ratio = t; kd = k; } }
            }
          }
          {  // both reductions behind one barrier
            dr = warp_sum(dr);
            warp_argmin_redux(ratio, kd);
            double* rv = red + 24 * rpar;
            int* ri = reinterpret_cast<int*>(rv + 8);
            if (lane == 0) { rv[warp] = ratio; ri[warp] = kd; rv[12 + warp] = dr; }
            __syncthreads();
            ratio = rv[0];
            kd = ri[0];
            dr = rv[12];
#pragma unroll
            for (int w = 1; w < NW; w++) {
              const double ov = rv[w];
              const int oi = ri[w];
              if (ov < ratio || (ov == ratio && oi < kd)) { ratio = ov; kd = oi; }
              dr += rv[12 + w];
            }
            rpar ^= 1;
          }
          const double rho2 = scale - dr;
          const bool dependent = !(rho2 > 1e-12 * scale);
          if (!dependent) {
            if (tid < n) {
              const int i = tid;
              double c0 = kn[i], c1 = 0.0, c2 = 0.0, c3 = 0.0;
              int k = 0;
              for (; k + 4 <= q; k += 4) {
                c0 = fma(-rvec[k], KN[k * nmax + i], c0);
                c1 = fma(-rvec[k + 1], KN[(k + 1) * nmax + i], c1);
                c2 = fma(-rvec[k + 2], KN[(k + 2) * nmax + i], c2);
                c3 = fma(-rvec[k + 3], KN[(k + 3) * nmax + i], c3);
              }
              for (; k < q; k++) c0 = fma(-rvec[k], KN[k * nmax + i], c0);
              z[i] = (c0 + c1) + (c2 + c3);
            }
            __syncthreads();
          }
          const double rho2_inv = dependent ? 0.0 : fast_rcp(rho2);
          const double t2 = dependent ? 1e300 : -sp * rho2_inv;
          const double t1 = ratio;
          const double t = fmin(t1, t2);
          if (t >= 1e299) { status = CMPC_ST_INFEASIBLE; done = true; break; }
          const bool full = (t2 <= t1);
          if (!dependent) {
            if (tid < n) x[tid] = fma(t, z[tid], x[tid]);
            for (int c = tid; c < m; c += NT) {
              int ia, iz;
              double va, vz;
              cons_of(c, mu_inv, ia, va, iz, vz);
              s[c] = fma(t, va * z[ia] + vz * z[iz], s[c]);
            }
          }
          for (int k = tid; k < q; k += NT) u[k] = fma(-t, rvec[k], u[k]);
          up += t;
          flops_acc += 2.0 * (4.0 * n + 4.0 * q + (double)q * q + (double)n * q + 4.0 * m + n);
          if (full) {
            if (q >= qcap) { status = CMPC_ST_WSOVERFLOW; done = true; break; }
            if (tid == 0) {
              act[q] = (short)p;
              u[q] = up;
              isact[p] = 1;
            }
            __syncthreads();
            // next candidate: request its rows of K now, border the working-set matrices while they travel
            double nbest;
            int pn;
            candidate(nbest, pn);
            const bool more = nbest < -P.tol_violation;
            int nia = 0, niz = 0;
            double nva = 0.0, nvz = 0.0, na = 0.0, nz = 0.0;
            if (more) {
              cons_of(pn, mu_inv, nia, nva, niz, nvz);
              if (tid < n) { na = k_entry(slot, n, tiled, nia, tid); nz = k_entry(slot, n, tiled, niz, tid); }
            }
            // border P with the new row: [P + r r'/rho2, -r/rho2; -r'/rho2, 1/rho2]; cache K n_p
            for (int k = warp; k < q; k += NW) {
              const double rk = rvec[k] * rho2_inv;
              for (int l = lane; l < q; l += 32) Pm[k * PSQ + l] = fma(rk, rvec[l], Pm[k * PSQ + l]);
            }
            for (int k = tid; k < q; k += NT) {
              const double rk = rvec[k] * rho2_inv;
              Pm[q * PSQ + k] = -rk;
              Pm[k * PSQ + q] = -rk;
            }
            if (tid < n) KN[q * nmax + tid] = kn[tid];
            if (tid == 0) Pm[q * PSQ + q] = rho2_inv;
            q++;
            flops_acc += 2.0 * (double)q * q;
            __syncthreads();
            if (!more) done = true;
            p = pn;
            pia = nia; piz = niz; pva = nva; pvz = nvz;
            ka = na; kz = nz;
            break;
          }
          // partial step: row kd leaves the working set (P deflated by its row / column), p stays the candidate
          __syncthreads();  // the multipliers and slacks have moved
          for (int k = tid; k < q; k += NT) col[k] = Pm[k * PSQ + kd];
          __syncthreads();
          {
            const double inv = 1.0 / col[kd];
            for (int k = warp; k < q; k += NW) {
              if (k == kd) continue;
              const double ck = col[k] * inv;
              for (int l = lane; l < q; l += 32)
                if (l != kd) Pm[k * PSQ + l] = fma(-ck, col[l], Pm[k * PSQ + l]);
            }
          }
          __syncthreads();
          const int last = q - 1;
          if (kd != last) {  // slot kd <- slot last
            for (int l = tid; l < last; l += NT)
              if (l != kd) {
                const double v = Pm[last * PSQ + l];
                Pm[kd * PSQ + l] = v;
                Pm[l * PSQ + kd] = v;
              }
            if (tid < n) KN[kd * nmax + tid] = KN[last * nmax + tid];
            if (tid == 0) {
              Pm[kd * PSQ + kd] = Pm[last * PSQ + last];
              isact[act[kd]] = 0;
              act[kd] = act[last];
              u[kd] = u[last];
            }
          } else if (tid == 0) {
            isact[act[kd]] = 0;
          }
          q--;
          flops_acc += 2.0 * (double)q * q;
          __syncthreads();
        }
      }
    }
    __syncthreads();
    if (status == CMPC_ST_WSOVERFLOW && P.overflow_list) {
      // left for the next capacity tier, with the working set reached here (up to 64 rows travel as 16-bit ids)
      if (tid == 0) {
        const int pos = atomicAdd(P.overflow_count, 1);
        P.overflow_list[pos] = inst;
        bc[1] = pos;
      }
      __syncthreads();
      const int pos = bc[1];
      if (P.resume_out) {
        int* rs = P.resume_out + (size_t)pos * CMPC_RESUME_INTS;
        const int qs = min(q, 2 * (CMPC_RESUME_INTS - 2));
        for (int k = tid; 2 * k < qs; k += NT) {
          const int lo = (unsigned short)act[2 * k], hi = (2 * k + 1 < qs) ? (unsigned short)act[2 * k + 1] : 0;
          rs[2 + k] = lo | (hi << 16);
        }
        bool with_p = false;
        if (P.rstate_out && pos < P.rstate_out_cap && qs == q && q * (q + 1) / 2 <= P.rstate_out_stride) {
          double* stp = P.rstate_out + (size_t)pos * P.rstate_out_stride;  // packed lower triangle
          for (int k = warp; k < q; k += NW)
            for (int l = lane; l <= k; l += 32) stp[k * (k + 1) / 2 + l] = Pm[k * PSQ + l];
          with_p = true;
        }
        if (tid == 0) { rs[0] = qs | (with_p ? (2 << 16) : 0); rs[1] = iters - 1; }
      }
      continue;
    }
    // ---- outputs: q_soln scatter (zeros for swing feet), objective, primal activity mask ----
    int fin = 1;  // a non-finite iterate (NaN / Inf upstream) reports CMPC_ST_NONFINITE and zero forces
    if (have)
      for (int i = tid; i < n; i += NT) fin = fin && isfinite(x[i]);
    fin = __syncthreads_and(fin);
    if (have && !fin) status = CMPC_ST_NONFINITE;
    const bool have_x = have && fin;
    if (P.forces) {
      double* out = P.forces + (size_t)inst * 12 * h;
      for (int idx = tid; idx < 12 * h; idx += NT) {
        const int k = idx / 3, comp = idx - 3 * k;
        const int j = fsinv[k];
        out[idx] = (j >= 0 && have_x) ? x[3 * j + comp] : 0.0;
      }
    }
    if (P.active) {
      signed char* out = P.active + (size_t)inst * 20 * h;
      for (int idx = tid; idx < 20 * h; idx += NT) {
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
        for (int i = tid; i < n; i += NT) part = fma(0.5 * __ldg(gg + i), x[i], part);
        for (int k = tid; k < q; k += NT) {
          const int c = act[k];
          if (c % 5 == 4) part -= 0.5 * u[k] * (double)gv[c / 5] * P.f_max;
        }
      }
      part = block_sum<NT>(part, red, tid);
      if (tid == 0) {
        if (P.objective) P.objective[inst] = have_x ? part : 0.0;
        if (P.status) P.status[inst] = status;
        if (P.iterations) P.iterations[inst] = iters;
      }
    }
  }
#ifdef CMPC_CANARY
  __syncthreads();
  canary_check(smem, cv.guard, cv.nguard, tid, NT, "cmpc_dual_team_kernel");
#endif
  if (tid == 0 && P.flops && flops_acc > 0.0) atomicAdd(P.flops + CMPC_K_DUAL, (unsigned long long)flops_acc);
}
