// Kernel 2 of the pipeline, fast tier: Goldfarb-Idnani dual active set on K = H^-1, one warp per instance,
// working sets of up to 32 rows (lane k owns working-set slot k).  Same method as cmpc_dual.cuh (which stays
// as the any-capacity tier for the few instances that outgrow 32 rows); this variant keeps the iterate in
// registers and trims the dependent chain of an iteration:
//   * x, the slacks s and the per-row constants (which two variables a row touches, with what sign) live in
//     registers: lane l owns variables l, l+32, ... and rows l, l+32, ...;
//   * argmin over the violated rows / the ratio test: two REDUX.MIN on an order-preserving 64-bit key plus a
//     ballot instead of five shuffle levels;
//   * the multipliers u and the slot's row constants live in the slot's lane; P = (N'KN)^-1 is a full
//     (not packed) matrix in shared memory, one row per lane; KN = K N is cached row by row;
//   * every O(q) loop is unrolled by four over zero-padded data so that its shared-memory loads overlap;
//   * the two rows of K of the NEXT candidate are requested from L2 as soon as the slacks have moved,
//     before the working-set matrices are bordered, so most of that latency is off the chain.
#pragma once

namespace {

struct FCarve {
  int kn, z, d, r, P, KN, misc, ostage, total;
  CMPC_CANARY_FIELDS
};

__host__ __device__ inline FCarve make_fcarve(int npl, int qcap) {
  FCarve c;
  int o = 0;
  CMPC_GUARD_INIT(c);
  c.kn = o; o += 8 * 32 * npl;
  CMPC_GUARD(o, c);
  c.z = o; o += 8 * 32 * npl;
  CMPC_GUARD(o, c);
  c.d = o; o += 8 * 40;
  CMPC_GUARD(o, c);
  c.r = o; o += 8 * 40;
  CMPC_GUARD(o, c);
  c.P = o; o += align16(8 * (qcap * ((qcap + 4) | 1) + 8));
  CMPC_GUARD(o, c);
  c.KN = o; o += 8 * (qcap + 4) * 32 * npl;
  CMPC_GUARD(o, c);
  c.misc = o; o += 3 * CMPC_MAX_FS + 16;  // fs, gv, fsinv bytes
  CMPC_GUARD(o, c);
  o = align16(o);
  c.ostage = o; o += 20 * CMPC_MAX_HORIZON + 4;  // activity bytes of one instance, staged for word-wide stores
  CMPC_GUARD(o, c);
  c.total = align16(o);
  return c;
}

// order-preserving map double -> uint64 (smaller double <=> smaller key)
__device__ __forceinline__ unsigned long long dkey(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// warp argmin of (val, idx): returns the smallest val and, among equal values, the smallest idx
__device__ __forceinline__ void warp_argmin_redux(double& val, int& idx) {
  const unsigned long long k = dkey(val);
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  const bool win = (hi == mhi) && (lo == mlo);
  const unsigned midx = __reduce_min_sync(0xffffffffu, win ? (unsigned)idx : 0xffffffffu);
  const unsigned src = __ffs(__ballot_sync(0xffffffffu, win && (unsigned)idx == midx)) - 1;
  val = __shfl_sync(0xffffffffu, val, src);
  idx = (int)midx;
}

// row c = 5j + t of the reduced problem as s(x) = va x[ia] + vz x[iz] - b >= 0, packed: ia | iz << 8 | t << 16
__device__ __forceinline__ int cons_pack(int c) {
  const int j = c / 5, t = c - 5 * j;
  const int iz = 3 * j + 2, ia = (t == 4) ? iz : 3 * j + (t >> 1);
  return ia | (iz << 8) | (t << 16);
}
__device__ __forceinline__ double cons_va(int pk, double mu_inv) {
  const int t = pk >> 16;
  return (t == 4) ? 0.0 : ((t & 1) ? -mu_inv : mu_inv);
}
__device__ __forceinline__ double cons_vz(int pk) { return ((pk >> 16) == 4) ? -1.0 : 1.0; }

}  // namespace

template <int WPC, int NPL, int MPL>
__global__ void __launch_bounds__(32 * WPC) cmpc_dual_fast_kernel(const __grid_constant__ CmpcParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = P.horizon, nmax = P.nmax, qcap = P.qcap;  // qcap <= 32
  const FCarve cv = make_fcarve(NPL, qcap);
  unsigned char* base = smem + (size_t)warp * cv.total;
  double* kns = reinterpret_cast<double*>(base + cv.kn);
  double* zs = reinterpret_cast<double*>(base + cv.z);
  double* dvec = reinterpret_cast<double*>(base + cv.d);
  double* rvec = reinterpret_cast<double*>(base + cv.r);
  double* Pm = reinterpret_cast<double*>(base + cv.P);
  double* KN = reinterpret_cast<double*>(base + cv.KN);
  unsigned char* fs = base + cv.misc;
  unsigned char* gv = fs + CMPC_MAX_FS;
  signed char* fsinv = reinterpret_cast<signed char*>(gv + CMPC_MAX_FS);
  signed char* astage = reinterpret_cast<signed char*>(base + cv.ostage);
  constexpr int NS = 32 * NPL;  // KN row stride
  const int PSQ = (qcap + 4) | 1;  // P row stride: room for the 4-wide padded loops, odd (conflict-free row-per-lane access)

  const int count = P.count_ptr ? min(*P.count_ptr, P.count) : P.count;
  const double mu_inv = P.mu_inv;
  unsigned flops_acc = 0u;  // integer: the flop count of an iteration costs no FP64 instruction
#ifdef CMPC_CANARY
  canary_fill(base, cv.guard, cv.nguard, lane, 32);
#endif
  for (int i = lane; i < qcap * PSQ + 8; i += 32) Pm[i] = 0.0;  // the unrolled loops read (finite, zero-weighted) padding
  for (int i = lane; i < (qcap + 4) * NS; i += 32) KN[i] = 0.0;
  for (int i = lane; i < 40; i += 32) { dvec[i] = 0.0; rvec[i] = 0.0; }
  __syncwarp();

  long long tclk = 0;
  if (P.phase_cycles && lane == 0) tclk = clock64();
  while (true) {
    int slot_i = 0;
    if (lane == 0) slot_i = atomicAdd(P.sched, 1);
    slot_i = __shfl_sync(0xffffffffu, slot_i, 0);
    if (slot_i >= count) break;
    const int inst = P.worklist ? P.worklist[slot_i] : slot_i;
    const double* slot = P.qws + (size_t)inst * P.qws_stride;
    const double* gg = slot + P.qws_goff;
    const double* x0 = gg + nmax;
    const int* hdr = reinterpret_cast<const int*>(x0 + nmax + 2);
    const bool tiled = P.k_tiled != 0;
    const int nc = hdr[0];
    const int st0 = hdr[1];
    const unsigned char* hb = reinterpret_cast<const unsigned char*>(hdr + 2);
    const int n = 3 * nc, m = 5 * nc;
    const bool have = (st0 == CMPC_ST_SOLVED) && n <= NS && m <= 32 * MPL;

    for (int k = lane; k < 4 * h; k += 32) fsinv[k] = -1;
    __syncwarp();
    double x[NPL], s[MPL];
    int cpk[MPL];
    unsigned amask = 0;  // bit j: row lane + 32 j is in the working set
#pragma unroll
    for (int e = 0; e < NPL; e++) x[e] = 0.0;
    if (have) {
      for (int j = lane; j < nc; j += 32) {
        const unsigned char k = hb[j];
        fs[j] = k;
        gv[j] = hb[CMPC_MAX_FS + j];
        fsinv[k] = (signed char)j;
      }
#pragma unroll
      for (int e = 0; e < NPL; e++) {
        const int i = lane + 32 * e;
        x[e] = (i < n) ? x0[i] : 0.0;
        kns[i] = x[e];  // staged for the slack gather below
      }
    }
    __syncwarp();
    int status = (st0 == CMPC_ST_SOLVED && !have) ? CMPC_ST_CAPACITY : st0, iters = 0, q = 0;
    // working-set slot state of this lane (slot = lane)
    int spk = 0, sact = -1;
    double u = 0.0;
    if (have) {
#pragma unroll
      for (int j = 0; j < MPL; j++) {
        const int c = lane + 32 * j;
        cpk[j] = 0;
        s[j] = 1e300;  // rows beyond m never violate
        if (c < m) {
          const int pk = cons_pack(c);
          cpk[j] = pk;
          const double b = ((pk >> 16) == 4) ? -(double)gv[c / 5] * P.f_max : 0.0;
          s[j] = cons_va(pk, mu_inv) * kns[pk & 0xff] + cons_vz(pk) * kns[(pk >> 8) & 0xff] - b;
        }
      }
      __syncwarp();
      if (P.phase_cycles && lane == 0) {
        const long long now = clock64();
        atomicAdd(P.phase_cycles + CMPC_PH_STORE, (unsigned long long)(now - tclk));  // instance setup
        tclk = now;
      }
      // most violated row outside the working set; its two rows of K
      double best = 1e300;
      int p = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < MPL; j++)
        if (s[j] < best) { best = s[j]; p = lane + 32 * j; }
      warp_argmin_redux(best, p);
      bool done = !(best < -P.tol_violation);
      double ka[NPL], kz[NPL];  // raw rows of K of the candidate
      int ppk = 0;
      if (!done) {
        ppk = cons_pack(p);
#pragma unroll
        for (int e = 0; e < NPL; e++) {
          const int i = lane + 32 * e;
          ka[e] = (i < n) ? k_entry(slot, n, tiled, ppk & 0xff, i) : 0.0;
          kz[e] = (i < n) ? k_entry(slot, n, tiled, (ppk >> 8) & 0xff, i) : 0.0;
        }
      }
      while (!done) {
        const int pia = ppk & 0xff, piz = (ppk >> 8) & 0xff;
        const double pva = cons_va(ppk, mu_inv), pvz = cons_vz(ppk);
        double kn[NPL];
#pragma unroll
        for (int e = 0; e < NPL; e++) {
          kn[e] = pva * ka[e] + pvz * kz[e];
          kns[lane + 32 * e] = kn[e];
        }
        __syncwarp();
        const double scale = pva * kns[pia] + pvz * kns[piz];
        double up = 0.0;
        while (true) {
          iters++;
          if (iters > P.max_iter) { status = CMPC_ST_MAXITER; done = true; break; }
          const int q4 = (q + 3) & ~3;
          // d = N' kn (slot per lane), r = P d
          const double d = (lane < q) ? cons_va(spk, mu_inv) * kns[spk & 0xff] + cons_vz(spk) * kns[(spk >> 8) & 0xff] : 0.0;
          dvec[lane] = d;
          __syncwarp();
          double rr = 0.0;
          if (lane < q) {
            const double* prow = Pm + lane * PSQ;
            double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
            for (int l = 0; l < q4; l += 4) {
              r0 = fma(prow[l], dvec[l], r0);
              r1 = fma(prow[l + 1], dvec[l + 1], r1);
              r2 = fma(prow[l + 2], dvec[l + 2], r2);
              r3 = fma(prow[l + 3], dvec[l + 3], r3);
            }
            rr = (r0 + r1) + (r2 + r3);
          }
          rvec[lane] = rr;
          const double dr = warp_sum(d * rr);
          double ratio = (lane < q && rr > 0.0) ? u * fast_rcp(rr) : 1e300;
          int kd = lane;
          warp_argmin_redux(ratio, kd);
          __syncwarp();
          const double rho2 = scale - dr;
          const bool dependent = !(rho2 > 1e-12 * scale);
          double z[NPL];
          if (!dependent) {
#pragma unroll
            for (int e = 0; e < NPL; e++) z[e] = kn[e];
            for (int k = 0; k < q4; k += 4) {  // rvec is zero beyond q, the KN padding rows are finite
#pragma unroll
              for (int kk = 0; kk < 4; kk++) {
                const double rk = rvec[k + kk];
#pragma unroll
                for (int e = 0; e < NPL; e++) z[e] = fma(-rk, KN[(k + kk) * NS + lane + 32 * e], z[e]);
              }
            }
#pragma unroll
            for (int e = 0; e < NPL; e++) zs[lane + 32 * e] = z[e];
            __syncwarp();
          }
          // slack of the candidate row (held by lane p & 31, register p >> 5)
          // (picked with integer masks: a chain of `if (j == p >> 5) sp = s[j]` is turned into an indexed load by the
          //  compiler, which moves s[] — read and written by every iteration — from registers to local memory)
          int sph = 0, spl = 0;
#pragma unroll
          for (int j = 0; j < MPL; j++) {
            const int pick = -(int)(j == (p >> 5));
            sph |= __double2hiint(s[j]) & pick;
            spl |= __double2loint(s[j]) & pick;
          }
          double sp = __hiloint2double(sph, spl);
          sp = __shfl_sync(0xffffffffu, sp, p & 31);
          const double rho2_inv = dependent ? 0.0 : fast_rcp(rho2);
          const double t2 = dependent ? 1e300 : -sp * rho2_inv;
          const double t1 = ratio;
          const double t = fmin(t1, t2);
          if (t >= 1e299) { status = CMPC_ST_INFEASIBLE; done = true; break; }
          const bool full = (t2 <= t1);
          if (!dependent) {
#pragma unroll
            for (int e = 0; e < NPL; e++) x[e] = fma(t, z[e], x[e]);
#pragma unroll
            for (int j = 0; j < MPL; j++) {
              const int pk = cpk[j];
              if (lane + 32 * j < m)
                s[j] = fma(t, cons_va(pk, mu_inv) * zs[pk & 0xff] + cons_vz(pk) * zs[(pk >> 8) & 0xff], s[j]);
            }
          }
          if (lane < q) u = fma(-t, rr, u);
          up += t;
          flops_acc += 2u * (unsigned)(5 * n + 4 * q + q * q + n * q + 4 * m);
          if (full) {
            if (q >= qcap) { status = CMPC_ST_WSOVERFLOW; done = true; break; }
            if (lane == (p & 31)) amask |= 1u << (p >> 5);
            // next candidate: request its rows of K now, border the working-set matrices while they travel
            double nbest = 1e300;
            int pn = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < MPL; j++)
              if (!((amask >> j) & 1u) && s[j] < nbest) { nbest = s[j]; pn = lane + 32 * j; }
            warp_argmin_redux(nbest, pn);
            const bool more = nbest < -P.tol_violation;
            double na[NPL], nz[NPL];
            int npk = 0;
            if (more) {
              npk = cons_pack(pn);
#pragma unroll
              for (int e = 0; e < NPL; e++) {
                const int i = lane + 32 * e;
                na[e] = (i < n) ? k_entry(slot, n, tiled, npk & 0xff, i) : 0.0;
                nz[e] = (i < n) ? k_entry(slot, n, tiled, (npk >> 8) & 0xff, i) : 0.0;
              }
            }
            // border P with the new row: [P + r r'/rho2, -r/rho2; -r'/rho2, 1/rho2]; cache K n_p; slot q <- row p
            if (lane < q) {
              const double rk = rr * rho2_inv;
              double* prow = Pm + lane * PSQ;
              for (int l = 0; l < q4; l += 4) {
                const double v0 = fma(rk, rvec[l], prow[l]), v1 = fma(rk, rvec[l + 1], prow[l + 1]);
                const double v2 = fma(rk, rvec[l + 2], prow[l + 2]), v3 = fma(rk, rvec[l + 3], prow[l + 3]);
                prow[l] = v0; prow[l + 1] = v1; prow[l + 2] = v2; prow[l + 3] = v3;
              }
            }
            __syncwarp();  // the padded writes above touch columns q..q4-1: order them before the new column
            if (lane < q) {
              const double rk = rr * rho2_inv;
              Pm[lane * PSQ + q] = -rk;
              Pm[q * PSQ + lane] = -rk;
            }
            if (lane == q) {
              Pm[q * PSQ + q] = rho2_inv;
              spk = ppk;
              sact = p;
              u = up;
            }
#pragma unroll
            for (int e = 0; e < NPL; e++) KN[q * NS + lane + 32 * e] = kn[e];
            q++;
            flops_acc += 2u * (unsigned)(q * q);
            __syncwarp();
            if (!more) done = true;
            p = pn;
            ppk = npk;
#pragma unroll
            for (int e = 0; e < NPL; e++) { ka[e] = na[e]; kz[e] = nz[e]; }
            break;
          }
          // partial step: slot kd leaves the working set (P deflated by its row/column), p stays the candidate
          {
            const double inv = fast_rcp(Pm[kd * PSQ + kd]);
            if (lane < q && lane != kd) {
              double* prow = Pm + lane * PSQ;
              const double* krow = Pm + kd * PSQ;
              const double ck = prow[kd] * inv;
              for (int l = 0; l < q4; l += 4) {
                const double v0 = fma(-ck, krow[l], prow[l]), v1 = fma(-ck, krow[l + 1], prow[l + 1]);
                const double v2 = fma(-ck, krow[l + 2], prow[l + 2]), v3 = fma(-ck, krow[l + 3], prow[l + 3]);
                prow[l] = v0; prow[l + 1] = v1; prow[l + 2] = v2; prow[l + 3] = v3;
              }
            }
          }
          __syncwarp();
          const int last = q - 1;
          const int dropped = __shfl_sync(0xffffffffu, sact, kd);
          if (lane == (dropped & 31)) amask &= ~(1u << (dropped >> 5));
          {
            // slot kd <- slot last (row, column, cached K n, multiplier, row constants); slot last is cleared
            const int lspk = __shfl_sync(0xffffffffu, spk, last), lsact = __shfl_sync(0xffffffffu, sact, last);
            const double lu = __shfl_sync(0xffffffffu, u, last);
            double vrow = 0.0;
            if (lane < last && lane != kd) vrow = Pm[last * PSQ + lane];
            const double vdiag = Pm[last * PSQ + last];
            __syncwarp();
            if (kd != last) {
              if (lane < last && lane != kd) {
                Pm[kd * PSQ + lane] = vrow;
                Pm[lane * PSQ + kd] = vrow;
              }
              if (lane == kd) {
                Pm[kd * PSQ + kd] = vdiag;
                spk = lspk;
                sact = lsact;
                u = lu;
              }
#pragma unroll
              for (int e = 0; e < NPL; e++) KN[kd * NS + lane + 32 * e] = KN[last * NS + lane + 32 * e];
            }
            // row and column `last` leave the matrix: zero them so that the padded loops keep reading zeros
#ifdef CMPC_CANARY_SELFTEST  /* the round-1 defect, to prove that the canaries see it */
            Pm[last * PSQ + lane] = 0.0;
#else
            if (lane < PSQ) Pm[last * PSQ + lane] = 0.0;
#endif  // (a row is PSQ wide: with a small capacity PSQ < 32, and row
                                                          //  qcap - 1 is the last one — lanes beyond it would write past P)
            if (lane < qcap) Pm[lane * PSQ + last] = 0.0;
            if (lane == last) { sact = -1; u = 0.0; }
          }
          q--;
          flops_acc += 2u * (unsigned)(q * q);
          __syncwarp();
        }
      }
    }
    if (P.phase_cycles && lane == 0) {
      const long long now = clock64();
      atomicAdd(P.phase_cycles + CMPC_PH_QP, (unsigned long long)(now - tclk));  // active-set iterations
      tclk = now;
    }
    if (status == CMPC_ST_WSOVERFLOW && P.overflow_list) {
      // left for the any-capacity launch, which resumes from the working set reached here (the candidate row that
      // did not fit is found again there): slot k's row id travels as 16 bits
      int pos = 0;
      if (lane == 0) {
        pos = atomicAdd(P.overflow_count, 1);
        P.overflow_list[pos] = inst;
      }
      pos = __shfl_sync(0xffffffffu, pos, 0);
      if (P.resume_out) {
        int* rs = P.resume_out + (size_t)pos * CMPC_RESUME_INTS;
        const int mine = (lane < q) ? sact : 0;
        const int nb = __shfl_down_sync(0xffffffffu, mine, 1);
        if (!(lane & 1)) rs[2 + (lane >> 1)] = (mine & 0xffff) | (nb << 16);
        // P = (N'KN)^-1 of the working set travels too when there is room for it (it depends on the rows only and has
        // not been bordered with the candidate): the next tier then derives u = -P s_A(x0) and x = x0 + K N u directly
        // instead of re-bordering row by row
        bool with_p = false;
        if (P.rstate_out && pos < P.rstate_out_cap && q * q <= P.rstate_out_stride) {
          double* st = P.rstate_out + (size_t)pos * P.rstate_out_stride;
          for (int k = 0; k < q; k++)
            if (lane < q) st[k * q + lane] = Pm[k * PSQ + lane];
          with_p = true;
        }
        if (lane == 0) { rs[0] = q | (with_p ? (1 << 16) : 0); rs[1] = iters - 1; }
      }
      __syncwarp();
      continue;
    }
    // ---- outputs: q_soln scatter (zeros for swing feet), objective, primal activity mask ----
    // NaN / Inf anywhere upstream (inputs, disturbance estimate) ends here as a non-finite iterate: the violation test
    // `!(best < -tol)` stops the iterations on NaN.  Such an instance reports CMPC_ST_NONFINITE and zero forces.
    bool fin = true;
#pragma unroll
    for (int e = 0; e < NPL; e++) fin = fin && isfinite(x[e]);
    fin = __all_sync(0xffffffffu, fin);
    if (have && !fin) status = CMPC_ST_NONFINITE;
    const bool have_x = have && fin;
#pragma unroll
    for (int e = 0; e < NPL; e++) kns[lane + 32 * e] = x[e];
    __syncwarp();
    // staged in shared memory (the first rows of KN are free now and only ever need to hold finite values), then
    // written as whole 16-byte / 4-byte words, contiguous across the warp: the output arrays may be pinned HOST
    // memory that the kernel writes over PCIe (cmpc_batch_solve_host), where full-line writes are what counts
    double* fstage = KN;
    for (int k = lane; k < 4 * h; k += 32) {  // one foot-step per lane
      const int j = fsinv[k];
      double fx = 0.0, fy = 0.0, fz = 0.0;
      if (j >= 0 && have_x) { fx = kns[3 * j]; fy = kns[3 * j + 1]; fz = kns[3 * j + 2]; }
      fstage[3 * k] = fx; fstage[3 * k + 1] = fy; fstage[3 * k + 2] = fz;
      if (P.active) {
        signed char a[5] = {0, 0, 0, 0, 0};
        if (j >= 0 && have_x) {
          const double ub = (double)gv[j] * P.f_max;
          a[0] = (fx * mu_inv + fz <= P.tol_active) ? -1 : 0;
          a[1] = (-fx * mu_inv + fz <= P.tol_active) ? -1 : 0;
          a[2] = (fy * mu_inv + fz <= P.tol_active) ? -1 : 0;
          a[3] = (-fy * mu_inv + fz <= P.tol_active) ? -1 : 0;
          a[4] = (fz <= P.tol_active) ? -1 : 0;
          if (fz >= ub - P.tol_active) a[4] = 1;
        }
#pragma unroll
        for (int t = 0; t < 5; t++) astage[5 * k + t] = a[t];
      }
    }
    __syncwarp();
    if (P.forces) {
      double2* out = reinterpret_cast<double2*>(P.forces + (size_t)inst * 12 * h);  // 96 h bytes per instance: 16-byte aligned
      const double2* src = reinterpret_cast<const double2*>(fstage);
      for (int i = lane; i < 6 * h; i += 32) out[i] = src[i];
    }
    if (P.active) {
      int* out = reinterpret_cast<int*>(P.active + (size_t)inst * 20 * h);  // 20 h bytes per instance: 4-byte aligned
      const int* src = reinterpret_cast<const int*>(astage);
      for (int i = lane; i < 5 * h; i += 32) out[i] = src[i];
    }
    {
      // objective 0.5 x'Hx + g'x = 0.5 g'x + 0.5 lambda'b at a KKT point
      double part = 0.0;
      if (have_x) {
#pragma unroll
        for (int e = 0; e < NPL; e++) {
          const int i = lane + 32 * e;
          if (i < n) part = fma(0.5 * __ldg(gg + i), x[e], part);
        }
        if (lane < q && (spk >> 16) == 4) part -= 0.5 * u * (double)gv[sact / 5] * P.f_max;
      }
      part = warp_sum(part);
      if (lane == 0) {
        if (P.objective) P.objective[inst] = have_x ? part : 0.0;
        if (P.status) P.status[inst] = status;
        if (P.iterations) P.iterations[inst] = iters;
      }
    }
    __syncwarp();
    if (P.phase_cycles && lane == 0) {
      const long long now = clock64();
      atomicAdd(P.phase_cycles + CMPC_PH_OUT, (unsigned long long)(now - tclk));  // outputs
      tclk = now;
    }
  }
#ifdef CMPC_CANARY
  __syncwarp();
  canary_check(base, cv.guard, cv.nguard, lane, 32, "cmpc_dual_fast_kernel");
#endif
  if (lane == 0 && P.flops && flops_acc) atomicAdd(P.flops + CMPC_K_DUAL, (unsigned long long)flops_acc);
}
