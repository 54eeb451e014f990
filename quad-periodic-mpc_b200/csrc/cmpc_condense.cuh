// Kernel 1 of the two-kernel pipeline: condensation + K = H^-1 for reduced problems of up to 128 variables.
//
// One CTA owns one MPC instance at a time; CTAs are persistent and draw instances from a device-side
// counter (the next record is prefetched by cp.async.bulk while the current one is processed).  Per
// instance: state build, closed-form c2qp, g and H of the contact-only QP (SolverMPC.cpp:566-950, see
// cmpc_kernels.cu for the stage map), then K = H^-1 by BLOCKED symmetric sweeps with H in register tiles:
//
//   the 8 rows of block s (all held in register block a = s of the 8 x TX thread grid) are published as a
//   panel C (8 x n); one warp inverts the 8 x 8 diagonal block D in registers (Gauss-Jordan over warp
//   shuffles); all threads form M = D^-1 C; then every thread applies the rank-8 update
//   A -= C' M to its TM x TN tile: TM*TN*8 independent DFMAs between barriers, operands from shared memory.
//
// As in the scalar sweep (cmpc_sweep.cuh) the panel is published with D - I in place of D, which makes the
// same rank-8 update produce the swept pivot rows (D^-1 C), pivot columns and the diagonal block
// (2I - D^-1): a swept diagonal entry carries a constant +2 that is removed once at the end.  H is scaled
// by an exact power of two so that every pivot block has eigenvalues below 1 (no cancellation).
//
// The kernel leaves, per instance, K (n x n), g, x0 = -K g and the contact list in a workspace slot
// (L2-resident for the chunk sizes the host picks); kernel 2 (cmpc_dual.cuh) runs the active-set
// iterations one warp per instance on top of it.
#pragma once

namespace {

template <int TM_, int TX_, int TN_, int MINB_>
struct CShape {
  static constexpr int TY = 8, TM = TM_, TX = TX_, TN = TN_, NT = 8 * TX_, NPAD = 8 * TM_, MINB = MINB_;
  static constexpr int PS = NPAD + 4;  // panel row stride in doubles: = 4 (mod 16), conflict-free for both access patterns
  static_assert(8 * TM_ == TX_ * TN_, "square padded matrix");
  static_assert(TX_ == 16 || TX_ == 32, "a warp covers one or two rows of the thread grid");
};
using CShape64 = CShape<8, 16, 4, 4>;    // n <= 64: 128 threads, 8x4 tiles
using CShape96 = CShape<12, 32, 3, 2>;   // n <= 96: 256 threads, 12x3 tiles
using CShape128 = CShape<16, 32, 4, 1>;  // n <= 128: 256 threads, 16x4 tiles

struct CCarve {
  int rec0, rec1, bars, sig, small, evec, agg, fs, fsinv, g, hs, pan, mm, dinv, red, total;
  CMPC_CANARY_FIELDS
};

__host__ __device__ inline CCarve make_ccarve(int h, int nmax, int rec_stride, int npad, bool adapt) {
  CCarve c;
  int o = 0;
  CMPC_GUARD_INIT(c);
  c.rec0 = o; o += align16(rec_stride);
  CMPC_GUARD(o, c);
  c.rec1 = o; o += align16(rec_stride);
  CMPC_GUARD(o, c);
  c.bars = o; o += 16;
  CMPC_GUARD(o, c);
  c.sig = o; o += align16(8 * CMPC_SIG_COUNT * h * h);
  CMPC_GUARD(o, c);
  c.small = o; o += align16(8 * (36 + 36 + 144 + 144 + 16));  // W, RW, PT, PO, scalars
  CMPC_GUARD(o, c);
  c.evec = o; o += align16(8 * 12 * h);
  CMPC_GUARD(o, c);
  c.agg = o; o += align16(8 * 10 * h);
  CMPC_GUARD(o, c);
  c.fs = o; o += align16(4 * CMPC_MAX_FS);
  CMPC_GUARD(o, c);
  c.fsinv = o; o += align16(4 * CMPC_MAX_FS);
  CMPC_GUARD(o, c);
  c.g = o; o += align16(8 * npad);
  CMPC_GUARD(o, c);
  {
    int hb = 8 * nmax * nmax;  // H staging; the estimator stage borrows it for 3 x 400 doubles
    if (adapt && hb < 8 * CMPC_ADAPT_SCRATCH) hb = 8 * CMPC_ADAPT_SCRATCH;
    c.hs = o; o += align16(hb);
    CMPC_GUARD(o, c);
  }
  c.pan = o; o += align16(8 * 8 * (npad + 4));
  CMPC_GUARD(o, c);
  c.mm = o; o += align16(8 * 8 * (npad + 4));
  CMPC_GUARD(o, c);
  c.dinv = o; o += 8 * (64 + 64 + 2);  // D^-1, D itself, the refinement flag of the block step
  CMPC_GUARD(o, c);
  c.red = o; o += 512;
  CMPC_GUARD(o, c);
  c.total = o;
  return c;
}

// 8 x 8 SPD inverse by one warp.  Lane (r = lane & 7, rep = lane >> 3) holds D[r][2 rep] and D[r][2 rep + 1];
// Gauss-Jordan without pivoting, operands exchanged by shuffles.  Returns the lane's two entries of D^-1.
__device__ __forceinline__ void warp_inv8(double& a0, double& a1, int lane) {
  const int r = lane & 7, rep = lane >> 3;
#pragma unroll
  for (int p = 0; p < 8; p++) {
    const int prep = p >> 1;
    const double mine = (p & 1) ? a1 : a0;
    const double arp = __shfl_sync(0xffffffffu, mine, r | (prep << 3));   // D[r][p]
    const double dpp = __shfl_sync(0xffffffffu, mine, p | (prep << 3));   // D[p][p]
    const double ap0 = __shfl_sync(0xffffffffu, a0, p | (rep << 3));      // D[p][c0]
    const double ap1 = __shfl_sync(0xffffffffu, a1, p | (rep << 3));      // D[p][c0 + 1]
    const double dinv = fast_rcp(dpp);
    const double t = arp * dinv;
    double n0 = fma(-t, ap0, a0), n1 = fma(-t, ap1, a1);
    if (r == p) { n0 = ap0 * dinv; n1 = ap1 * dinv; }
    if (rep == prep) {
      if (p & 1) n1 = (r == p) ? dinv : -t;
      else n0 = (r == p) ? dinv : -t;
    }
    a0 = n0;
    a1 = n1;
  }
}

template <class S>
__device__ __forceinline__ void sweep_blocked(double (&A)[S::TM][S::TN], int n, int tid, double* pan, double* mm,
                                              double* dinvs, double refine_above) {
  constexpr int TX = S::TX, TM = S::TM, TN = S::TN, PS = S::PS;
  const int ty = tid / TX, tx = tid - ty * TX;
  const int lane = tid & 31;
  const int nblk = (n + 7) >> 3;
  double* dsave = dinvs + 64;
  volatile int* flag = reinterpret_cast<volatile int*>(dinvs + 128);
#pragma unroll 1
  for (int s = 0; s < nblk; s++) {
    const int k0 = 8 * s;
    // publish the panel: row ty of block s, with D - I in the diagonal block
#pragma unroll
    for (int aa = 0; aa < TM; aa++) {
      if (aa == s) {
#pragma unroll
        for (int bb = 0; bb < TN; bb++) {
          const int j = tx + TX * bb;
          double v = A[aa][bb];
          // the pivot block itself goes to warp 0 as it is: (d - 1) + 1 would cost the small pivots their low bits
          if (j >= k0 && j < k0 + 8) dsave[ty * 8 + (j - k0)] = v;
          if (j == k0 + ty) v -= 1.0;
          pan[ty * PS + j] = v;
        }
      }
    }
    __syncthreads();
    // D^-1 by warp 0
    if (tid < 32) {
      const int r = lane & 7, c0 = 2 * (lane >> 3);
      double a0 = dsave[r * 8 + c0], a1 = dsave[r * 8 + c0 + 1];
      warp_inv8(a0, a1, lane);
      dinvs[r * 8 + c0] = a0;
      dinvs[r * 8 + c0 + 1] = a1;
      const bool big = refine_above >= 0.0 && __any_sync(0xffffffffu, fmax(fabs(a0), fabs(a1)) > refine_above);
      if (lane == 0) *flag = big ? 1 : 0;
    }
    __syncthreads();
    // M = D^-1 C  (thread (ty, tx): row ty, columns tx + TX b)
    double mreg[TN];
    {
      double di[8];
#pragma unroll
      for (int q = 0; q < 8; q++) di[q] = dinvs[ty * 8 + q];
#pragma unroll
      for (int bb = 0; bb < TN; bb++) {
        const int j = tx + TX * bb;
        // the eight panel entries up front (volatile loads): with the tile in registers the allocator otherwise recycles
        // one register pair and every DFMA of the chain waits a shared-memory round trip
        double pv[8];
#pragma unroll
        for (int q = 0; q < 8; q++) pv[q] = lds_f64v(pan + q * PS + j);
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < 8; q++) acc = fma(di[q], pv[q], acc);
        mm[ty * PS + j] = acc;
        mreg[bb] = acc;
      }
    }
    __syncthreads();
    // Block Gauss-Jordan with an explicitly inverted pivot block loses ~cond(D) digits more than a scalar sweep.  One
    // residual correction of the panel, M += D^-1 (C - D M), gives them back; it is applied only to block steps whose
    // pivot-block inverse is large (P.inv_refine; next to none on the A1 defaults).  The flag is uniform over the CTA.
    if (*flag) {
      double rr[TN];
      {
        double dr[8];
#pragma unroll
        for (int q = 0; q < 8; q++) dr[q] = dsave[ty * 8 + q];
#pragma unroll
        for (int bb = 0; bb < TN; bb++) {
          const int j = tx + TX * bb;
          double acc = pan[ty * PS + j];
#pragma unroll
          for (int q = 0; q < 8; q++) acc = fma(-dr[q], mm[q * PS + j], acc);
          rr[bb] = acc;
        }
      }
      __syncthreads();
#pragma unroll
      for (int bb = 0; bb < TN; bb++) mm[ty * PS + tx + TX * bb] = rr[bb];
      __syncthreads();
      {
        double di[8];
#pragma unroll
        for (int q = 0; q < 8; q++) di[q] = dinvs[ty * 8 + q];
#pragma unroll
        for (int bb = 0; bb < TN; bb++) {
          const int j = tx + TX * bb;
          double acc = mreg[bb];
#pragma unroll
          for (int q = 0; q < 8; q++) acc = fma(di[q], mm[q * PS + j], acc);
          mreg[bb] = acc;
        }
      }
      __syncthreads();
#pragma unroll
      for (int bb = 0; bb < TN; bb++) mm[ty * PS + tx + TX * bb] = mreg[bb];
      __syncthreads();
    }
    // rank-8 update of the tile
#pragma unroll 2
    for (int p = 0; p < 8; p++) {
      double ci[TM], cj[TN];
#pragma unroll
      for (int aa = 0; aa < TM; aa++) ci[aa] = pan[p * PS + ty + 8 * aa];
#pragma unroll
      for (int bb = 0; bb < TN; bb++) cj[bb] = mm[p * PS + tx + TX * bb];
#pragma unroll
      for (int aa = 0; aa < TM; aa++)
#pragma unroll
        for (int bb = 0; bb < TN; bb++) A[aa][bb] = fma(-ci[aa], cj[bb], A[aa][bb]);
    }
    // the next publish overwrites pan: every thread must be done reading it (mm is rewritten only after the
    // next two barriers)
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The same blocked sweep on the FP64 tensor cores (DMMA m8n8k4) for the 96 / 128 shapes: the padded symmetric matrix
// is held as its lower-triangular 8 x 8 tiles in accumulator layout (lane l: row l >> 2, columns 2 (l & 3), +1),
// dealt round-robin to the eight warps of the CTA (10 / 17 tiles per warp).  Block step s:
//   1. the owners publish the pivot rows as an 8 x NPAD panel C (tiles (s, J <= s) as they are, tiles (I > s, s)
//      transposed — the matrix is symmetric), D - I in the diagonal block;
//   2. warp 0 inverts D (Gauss-Jordan over shuffles) and leaves -D^-1;
//   3. M = -D^-1 C, the column tiles dealt to the warps (two DMMAs each);
//   4. every tile (I, J) += C_I' M_J (two DMMAs per tile).
// Same algebra as sweep_blocked / cmpc_invert_mma.cuh: a swept diagonal entry carries a constant +2.
// ---------------------------------------------------------------------------------------------------------------
template <class S>
struct DSweep {
  static constexpr int NBLK = S::NPAD / 8, NTILE = NBLK * (NBLK + 1) / 2, NW = S::NT / 32, TPW = (NTILE + NW - 1) / NW;
};

__device__ __forceinline__ void cdmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 8 x 8 inverse of a tile in accumulator layout (lane (r, q): D[r][2q], D[r][2q+1]), Gauss-Jordan without pivoting
__device__ __forceinline__ void cinv8_acc(double& a0, double& a1, int r, int q) {
#pragma unroll
  for (int p = 0; p < 8; p++) {
    const int pq = p >> 1;
    const double mine = (p & 1) ? a1 : a0;
    const double arp = __shfl_sync(0xffffffffu, mine, r * 4 + pq);  // D[r][p]
    const double dpp = __shfl_sync(0xffffffffu, mine, p * 4 + pq);  // D[p][p]
    const double ap0 = __shfl_sync(0xffffffffu, a0, p * 4 + q);     // D[p][2q]
    const double ap1 = __shfl_sync(0xffffffffu, a1, p * 4 + q);     // D[p][2q+1]
    const double dinv = fast_rcp(dpp);
    const double t = arp * dinv;
    double n0 = fma(-t, ap0, a0), n1 = fma(-t, ap1, a1);
    if (r == p) { n0 = ap0 * dinv; n1 = ap1 * dinv; }
    if (q == pq) {
      if (p & 1) n1 = (r == p) ? dinv : -t;
      else n0 = (r == p) ? dinv : -t;
    }
    a0 = n0;
    a1 = n1;
  }
}

template <class S>
__device__ __forceinline__ void sweep_dmma(double (&t)[DSweep<S>::TPW][2], const int (&tI)[DSweep<S>::TPW],
                                           const int (&tJ)[DSweep<S>::TPW], int n, int tid, double* pan, double* mm,
                                           double* dv, double refine_above) {
  constexpr int PS = S::PS, TPW = DSweep<S>::TPW, NW = DSweep<S>::NW;
  const int lane = tid & 31, w = tid >> 5, r = lane >> 2, q = lane & 3;
  const int fo = q * PS + r;  // fragment offset: element (k = q, row / column = r)
  const int nblk = (n + 7) >> 3;
#pragma unroll 1
  for (int s = 0; s < nblk; s++) {
    // 1. panel
#pragma unroll
    for (int i = 0; i < TPW; i++) {
      const int I = tI[i], J = tJ[i];
      if (I == s) {
        double v0 = t[i][0], v1 = t[i][1];
        if (J == s) {
          // the pivot block itself goes to warp 0 as it is: (d - 1) + 1 would cost the small pivots their low bits
          *reinterpret_cast<double2*>(dv + r * 8 + 2 * q) = make_double2(v0, v1);
          if (r == 2 * q) v0 -= 1.0;
          if (r == 2 * q + 1) v1 -= 1.0;
        }
        *reinterpret_cast<double2*>(pan + r * PS + 8 * J + 2 * q) = make_double2(v0, v1);
      } else if (J == s && I > s && I < nblk) {
        pan[(2 * q) * PS + 8 * I + r] = t[i][0];
        pan[(2 * q + 1) * PS + 8 * I + r] = t[i][1];
      }
    }
    __syncthreads();
    // 2. -D^-1 by warp 0
    if (w == 0) {
      const double2 d = *reinterpret_cast<const double2*>(dv + r * 8 + 2 * q);
      double d0 = d.x, d1 = d.y;
      *reinterpret_cast<double2*>(dv + 64 + r * 8 + 2 * q) = d;  // D itself stays for the refinement
      cinv8_acc(d0, d1, r, q);
      *reinterpret_cast<double2*>(dv + r * 8 + 2 * q) = make_double2(-d0, -d1);
      const bool big = refine_above >= 0.0 && __any_sync(0xffffffffu, fmax(fabs(d0), fabs(d1)) > refine_above);
      if (lane == 0) *reinterpret_cast<volatile int*>(dv + 128) = big ? 1 : 0;
    }
    __syncthreads();
    // 3. M = -D^-1 C; where the pivot-block inverse is large, one residual correction M += -D^-1 (C + D M) (see
    //    sweep_blocked) — a column tile of M belongs to one warp, so the correction needs no block barrier
    {
      const double a0 = dv[r * 8 + q], a1 = dv[r * 8 + 4 + q];
      const bool refine = *reinterpret_cast<volatile int*>(dv + 128) != 0;
      double da0 = 0.0, da1 = 0.0;
      if (refine) {
        da0 = dv[64 + r * 8 + q];
        da1 = dv[64 + r * 8 + 4 + q];
      }
      for (int J = w; J < nblk; J += NW) {
        double m0 = 0.0, m1 = 0.0;
        cdmma(m0, m1, a0, pan[fo + 8 * J]);
        cdmma(m0, m1, a1, pan[fo + 4 * PS + 8 * J]);
        double2* mt = reinterpret_cast<double2*>(mm + r * PS + 8 * J + 2 * q);
        *mt = make_double2(m0, m1);
        if (refine) {
          __syncwarp();
          const double2 c = *reinterpret_cast<const double2*>(pan + r * PS + 8 * J + 2 * q);
          double r0 = c.x, r1 = c.y;
          cdmma(r0, r1, da0, mm[fo + 8 * J]);
          cdmma(r0, r1, da1, mm[fo + 4 * PS + 8 * J]);
          __syncwarp();
          *mt = make_double2(r0, r1);
          __syncwarp();
          cdmma(m0, m1, a0, mm[fo + 8 * J]);
          cdmma(m0, m1, a1, mm[fo + 4 * PS + 8 * J]);
          __syncwarp();
          *mt = make_double2(m0, m1);
        }
      }
    }
    __syncthreads();
    // 4. every tile (I, J) += C_I' M_J
#pragma unroll
    for (int i = 0; i < TPW; i++) {
      const int I = tI[i], J = tJ[i];
      if (I >= 0 && I < nblk) {
        cdmma(t[i][0], t[i][1], pan[fo + 8 * I], mm[fo + 8 * J]);
        cdmma(t[i][0], t[i][1], pan[fo + 4 * PS + 8 * I], mm[fo + 4 * PS + 8 * J]);
      }
    }
    __syncthreads();  // the next publish overwrites pan
  }
}

}  // namespace

template <class S, bool ADAPT>
__global__ void __launch_bounds__(S::NT, S::MINB) cmpc_condense_kernel(const __grid_constant__ CmpcParams P) {
  constexpr int NT = S::NT, TX = S::TX, TM = S::TM, TN = S::TN, NPAD = S::NPAD;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int h = P.horizon, hh = h * h;
  const CCarve cv = make_ccarve(h, P.nmax, P.rec_stride, NPAD, ADAPT);
#ifdef CMPC_CANARY
  canary_fill(smem, cv.guard, cv.nguard, tid, NT);
  __syncthreads();
#endif
  unsigned char* recbuf[2] = {smem + cv.rec0, smem + cv.rec1};
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + cv.bars);
  double* sig = reinterpret_cast<double*>(smem + cv.sig);
  double* sW = reinterpret_cast<double*>(smem + cv.small);  // W[4][3][3]
  double* sRW = sW + 36;                                    // (R^T W)[4][3][3]
  double* sPT = sRW + 36;                                   // PT[4][4][3][3]
  double* sPO = sPT + 144;                                  // PO[4][4][3][3]
  double* sScal = sPO + 144;                                // [1..6] position / velocity weights, [8..10] roll pitch yaw
  double* ev = reinterpret_cast<double*>(smem + cv.evec);
  double* agg = reinterpret_cast<double*>(smem + cv.agg);
  int* fs = reinterpret_cast<int*>(smem + cv.fs);
  int* fsinv = reinterpret_cast<int*>(smem + cv.fsinv);
  double* g = reinterpret_cast<double*>(smem + cv.g);
  double* Hs = reinterpret_cast<double*>(smem + cv.hs);
  double* pan = reinterpret_cast<double*>(smem + cv.pan);
  double* mm = reinterpret_cast<double*>(smem + cv.mm);
  double* dinvs = reinterpret_cast<double*>(smem + cv.dinv);
  double* red = reinterpret_cast<double*>(smem + cv.red);
  int* redi = reinterpret_cast<int*>(red + 32);  // [0] = nc, [2], [3] = next instance (double-buffered)

  const int count = P.count;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < CMPC_SIG_COUNT * hh; i += NT) sig[i] = __ldg(P.sigma + i);
  if (tid == 0) {
    const int cur0 = atomicAdd(P.sched, 1);
    redi[2] = cur0;
    if (cur0 < count) {
      mbar_expect_tx(&bars[0], (uint32_t)P.rec_stride);
      bulk_g2s(recbuf[0], P.records + (size_t)cur0 * P.rec_stride, (uint32_t)P.rec_stride, &bars[0]);
    }
  }
  __syncthreads();
  int cur = redi[2];

  uint32_t phase[2] = {0u, 0u};
  const double dt = P.dt, minv = P.mass_inv;
  double flops_acc = 0.0;
  PhaseClock pc;
  pc.init(P.phase_cycles, reinterpret_cast<long long*>(red + 48), tid);
  const int ty = tid / TX, tx = tid - ty * TX;

  for (int buf = 0; cur < count; buf ^= 1) {
    const int inst = cur;
    if (tid == 0) {  // draw and prefetch the next instance
      const int nxt = atomicAdd(P.sched, 1);
      redi[2 + (buf ^ 1)] = nxt;
      if (nxt < count) {
        fence_proxy_async();
        mbar_expect_tx(&bars[buf ^ 1], (uint32_t)P.rec_stride);
        bulk_g2s(recbuf[buf ^ 1], P.records + (size_t)nxt * P.rec_stride, (uint32_t)P.rec_stride, &bars[buf ^ 1]);
      }
    }
    mbar_wait(&bars[buf], phase[buf]);
    phase[buf] ^= 1u;
    pc.tick(CMPC_PH_WAIT);
    const float* rec = reinterpret_cast<const float*>(recbuf[buf]);
    const unsigned char* gait = recbuf[buf] + 4 * (CMPC_REC_TRAJ + 12 * h);
    double* slot = P.qws + (size_t)inst * P.qws_stride;
    int* hdr = reinterpret_cast<int*>(slot + P.qws_goff + 2 * P.nmax + 2);

    // ---- 0. periodic-disturbance estimator (Adaptive MPC): xi for this instance, SolverMPC.cpp:688-798 ----
    if (ADAPT) {
      double* est_s = red + 40;
      if (P.adapt_mode == 0 || P.adapt_mode == 1) {
        estimate_disturbance<NT>(P, inst, tid, Hs, red, est_s);
      } else {
        if (tid < 4) est_s[tid] = P.est[(size_t)inst * 4 + tid];
        __syncthreads();
      }
      if (tid == 0) {
        const double simt = (double)P.sim_time[inst];
        const float comp = (float)(est_s[1] + sin(2.0 * M_PI * simt * est_s[2] + est_s[3]));
        float* fe = P.f_est + (size_t)inst * 6;
        fe[3] = comp;
        if (P.adapt_mode == 0 || P.adapt_mode == 1)
          for (int i = 0; i < 4; i++) P.est[(size_t)inst * 4 + i] = est_s[i];
        if (P.adapt_mode >= 1) {
          float* xi = reinterpret_cast<float*>(recbuf[buf]) + CMPC_REC_FDIST;
          for (int i = 0; i < 6; i++) xi[i] = (i == 3) ? comp : fe[i];
        }
      }
      __syncthreads();
      pc.tick(CMPC_PH_ADAPT);
    }

    // ---- A. contact foot-steps (warp 0), Euler angles (one lane of warps 1..3), W_f and R^T W_f ----
    if (tid < 32) {
      int cnt = 0;
      for (int base = 0; base < 4 * h; base += 32) {
        const int k = base + tid;
        bool keep = false;
        if (k < 4 * h) {
          const double ub = (double)gait[k] * P.f_max;
          keep = !(ub < 0.01 && ub > -0.01);  // the reference drops a foot-step whose fz bound is ~0
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        const int pos = cnt + __popc(mask & ((1u << tid) - 1u));
        if (k < 4 * h) fsinv[k] = keep ? pos : -1;
        if (keep) fs[pos] = k;
        cnt += __popc(mask);
      }
      if (tid == 0) redi[0] = cnt;
    } else if ((tid & 31) == 0 && tid < 128) {
      // quat_to_rpy, SolverMPC.cpp:352-361
      const double qw = rec[CMPC_REC_Q + 0], qx = rec[CMPC_REC_Q + 1], qy = rec[CMPC_REC_Q + 2], qz = rec[CMPC_REC_Q + 3];
      const int which = tid >> 5;
      double val;
      if (which == 1) val = atan2(2.0 * (qy * qz + qw * qx), qw * qw - qx * qx - qy * qy + qz * qz);
      else if (which == 2) val = asin(fmin(-2.0 * (qx * qz - qw * qy), 0.99999));
      else val = atan2(2.0 * (qx * qy + qw * qz), qw * qw + qx * qx - qy * qy - qz * qz);
      sScal[7 + which] = val;
    }
    double R[9];
    {
      const double qw = rec[CMPC_REC_Q + 0], qx = rec[CMPC_REC_Q + 1], qy = rec[CMPC_REC_Q + 2], qz = rec[CMPC_REC_Q + 3];
      const double tx2 = 2 * qx, ty2 = 2 * qy, tz2 = 2 * qz;
      const double twx = tx2 * qw, twy = ty2 * qw, twz = tz2 * qw, txx = tx2 * qx, txy = ty2 * qx, txz = tz2 * qx;
      const double tyy = ty2 * qy, tyz = tz2 * qy, tzz = tz2 * qz;
      R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
      R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
      R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
    }
    if (tid < 72) {
      double Ii[9];
      {
        double Iw[9];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++)
            Iw[i * 3 + j] = R[i * 3 + 0] * P.inertia[0] * R[j * 3 + 0] + R[i * 3 + 1] * P.inertia[1] * R[j * 3 + 1] +
                            R[i * 3 + 2] * P.inertia[2] * R[j * 3 + 2];
        const double c00 = Iw[4] * Iw[8] - Iw[5] * Iw[7], c01 = Iw[5] * Iw[6] - Iw[3] * Iw[8],
                     c02 = Iw[3] * Iw[7] - Iw[4] * Iw[6];
        const double id = 1.0 / (Iw[0] * c00 + Iw[1] * c01 + Iw[2] * c02);
        Ii[0] = c00 * id; Ii[1] = (Iw[2] * Iw[7] - Iw[1] * Iw[8]) * id; Ii[2] = (Iw[1] * Iw[5] - Iw[2] * Iw[4]) * id;
        Ii[3] = c01 * id; Ii[4] = (Iw[0] * Iw[8] - Iw[2] * Iw[6]) * id; Ii[5] = (Iw[2] * Iw[3] - Iw[0] * Iw[5]) * id;
        Ii[6] = c02 * id; Ii[7] = (Iw[1] * Iw[6] - Iw[0] * Iw[7]) * id; Ii[8] = (Iw[0] * Iw[4] - Iw[1] * Iw[3]) * id;
      }
      const int e = tid;
      const int which = e / 36, ee = e - 36 * which;
      const int f = ee / 9, i = (ee % 9) / 3, j = ee % 3;
      const double rx = rec[CMPC_REC_R + 0 * 4 + f], ry = rec[CMPC_REC_R + 1 * 4 + f], rz = rec[CMPC_REC_R + 2 * 4 + f];
      const double c0 = (j == 0) ? 0.0 : (j == 1 ? -rz : ry);  // column j of [r]x
      const double c1 = (j == 0) ? rz : (j == 1 ? 0.0 : -rx);
      const double c2 = (j == 0) ? -ry : (j == 1 ? rx : 0.0);
      const double w0 = Ii[0] * c0 + Ii[1] * c1 + Ii[2] * c2;
      const double w1 = Ii[3] * c0 + Ii[4] * c1 + Ii[5] * c2;
      const double w2 = Ii[6] * c0 + Ii[7] * c1 + Ii[8] * c2;
      const double r0 = (i == 0) ? R[0] : (i == 1 ? R[1] : R[2]);
      const double r1 = (i == 0) ? R[3] : (i == 1 ? R[4] : R[5]);
      const double r2 = (i == 0) ? R[6] : (i == 1 ? R[7] : R[8]);
      if (which == 0) sW[f * 9 + i * 3 + j] = (i == 0) ? w0 : (i == 1 ? w1 : w2);
      else sRW[f * 9 + i * 3 + j] = r0 * w0 + r1 * w1 + r2 * w2;
    }
    if (tid >= 96 && tid < 102) {
      const int t = tid - 96;
      sScal[1 + t] = (double)rec[CMPC_REC_WEIGHTS + (t < 3 ? 3 + t : 6 + t)];  // position, velocity weights
    }
    __syncthreads();
    const int nc = redi[0];
    const int n = 3 * nc;

    // ---- B. weighted tracking error of the free response, e_r = S (Adt^(r+1) x0 + sum_k Adt^k Qdt xi - Xd_r);
    //         foot-pair blocks PT = RW_i' S_theta RW_j, PO = W_i' S_omega W_j ----
    {
      const double xd = rec[CMPC_REC_XDRAG];
      const double om0 = rec[CMPC_REC_W + 0], om1 = rec[CMPC_REC_W + 1], om2 = rec[CMPC_REC_W + 2];
      const double ft0 = rec[CMPC_REC_FDIST + 0], ft1 = rec[CMPC_REC_FDIST + 1], ft2 = rec[CMPC_REC_FDIST + 2];
      const double ffx = rec[CMPC_REC_FDIST + 3];
      const double az = xd * (double)rec[CMPC_REC_V + 0] + P.gravity;  // row 11 of A x0
      for (int idx = tid; idx < 12 * h; idx += NT) {
        const int r = idx / 12, c = idx - 12 * r;
        const double T = (double)(r + 1) * dt, T2 = 0.5 * T * T;
        double val;
        if (c < 3) {
          const double ra = (c == 0) ? R[0] : (c == 1 ? R[1] : R[2]);
          const double rb = (c == 0) ? R[3] : (c == 1 ? R[4] : R[5]);
          const double rcc = (c == 0) ? R[6] : (c == 1 ? R[7] : R[8]);
          const double rto = ra * om0 + rb * om1 + rcc * om2;
          const double rtf = ra * ft0 + rb * ft1 + rcc * ft2;
          val = sScal[8 + c] + T * rto + T2 * rtf;
        } else if (c < 6) {
          const int a = c - 3;
          val = (double)rec[CMPC_REC_P + a] + T * (double)rec[CMPC_REC_V + a] + T2 * (double)rec[CMPC_REC_FDIST + 3 + a];
          if (a == 2) val += T2 * az + (T * T * T / 6.0) * xd * ffx;
        } else if (c < 9) {
          const int a = c - 6;
          val = (double)rec[CMPC_REC_W + a] + T * (double)rec[CMPC_REC_FDIST + a];
        } else {
          const int a = c - 9;
          val = (double)rec[CMPC_REC_V + a] + T * (double)rec[CMPC_REC_FDIST + 3 + a];
          if (a == 2) val += T * az + T2 * xd * ffx;
        }
        ev[idx] = (double)rec[CMPC_REC_WEIGHTS + c] * (val - (double)rec[CMPC_REC_TRAJ + idx]);
      }
    }
    for (int e = tid; e < 288; e += NT) {
      const int which = e / 144, ee = e - 144 * which;
      const int fi = ee / 36, fj = (ee / 9) & 3, a = (ee % 9) / 3, b = ee % 3;
      const double* Mi = (which == 0 ? sRW : sW) + fi * 9;
      const double* Mj = (which == 0 ? sRW : sW) + fj * 9;
      const int wo = which == 0 ? 0 : 6;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 3; k++) acc += Mi[k * 3 + a] * (double)rec[CMPC_REC_WEIGHTS + wo + k] * Mj[k * 3 + b];
      (which == 0 ? sPT : sPO)[ee] = acc;
    }
    __syncthreads();
    // horizon aggregates of e:  agg[c][0:3]=sum c2 e_theta, [3:6]=sum c1 e_omega,
    // [6:9]=(sum c2 e_p + c1 e_v)/m, [9]=xd/m (sum c3 e_pz + c2 e_vz)
    {
      const double xd = rec[CMPC_REC_XDRAG];
      for (int idx = tid; idx < 10 * h; idx += NT) {
        const int c = idx / 10, comp = idx - 10 * c;
        double acc = 0.0;
        for (int r = c; r < h; r++) {
          const double tau = (double)(r - c) * dt;
          const double c1 = dt, c2 = tau * dt + 0.5 * dt * dt,
                       c3 = 0.5 * tau * tau * dt + 0.5 * tau * dt * dt + dt * dt * dt / 6.0;
          const double* e = ev + 12 * r;
          if (comp < 3) acc += c2 * e[comp];
          else if (comp < 6) acc += c1 * e[6 + comp - 3];
          else if (comp < 9) acc += (c2 * e[3 + comp - 6] + c1 * e[9 + comp - 6]) * minv;
          else acc += (c3 * e[5] + c2 * e[11]) * xd * minv;
        }
        agg[idx] = acc;
      }
    }
    __syncthreads();
    pc.tick(CMPC_PH_PREP);

    int status = CMPC_ST_SOLVED;
    if (nc == 0) status = CMPC_ST_EMPTY;
    else if (n > P.nmax || n > NPAD) status = CMPC_ST_CAPACITY;
    if (status == CMPC_ST_SOLVED) {
      // ---- C. gradient; H assembled per 3x3 foot-step-pair block (lower triangle, mirrored) ----
      for (int I = tid; I < NPAD; I += NT) {
        double val = 0.0;
        if (I < n) {
          const int j = I / 3, comp = I - 3 * j;
          const int k = fs[j], step = k >> 2, f = k & 3;
          const double* a = agg + 10 * step;
          double acc = sRW[f * 9 + 0 + comp] * a[0] + sRW[f * 9 + 3 + comp] * a[1] + sRW[f * 9 + 6 + comp] * a[2] +
                       sW[f * 9 + 0 + comp] * a[3] + sW[f * 9 + 3 + comp] * a[4] + sW[f * 9 + 6 + comp] * a[5] + a[6 + comp];
          if (comp == 0) acc += a[9];
          val = 2.0 * acc;
        }
        g[I] = val;
      }
      double dmax = 0.0;
      {
        const double xd = rec[CMPC_REC_XDRAG];
        const double m2 = minv * minv, alpha2 = 2.0 * (double)rec[CMPC_REC_ALPHA];
        const double* wp = sScal + 1;
        const int nb = nc * (nc + 1) / 2;
        for (int blk = tid; blk < nb; blk += NT) {
          int j1 = (int)((sqrtf(8.0f * (float)blk + 1.0f) - 1.0f) * 0.5f);
          while ((j1 + 1) * (j1 + 2) / 2 <= blk) j1++;
          while (j1 * (j1 + 1) / 2 > blk) j1--;
          const int j2 = blk - j1 * (j1 + 1) / 2;
          const int k1 = fs[j1], k2 = fs[j2];
          const int sa = k1 >> 2, fi = k1 & 3, sb = k2 >> 2, fj = k2 & 3;
          const int ab = sa * h + sb, ba = sb * h + sa;
          const double s11 = sig[CMPC_SIG_11 * hh + ab], s22 = sig[CMPC_SIG_22 * hh + ab];
          double x20 = 0.0, x02 = 0.0, x00 = 0.0;  // x_drag couplings (z,x), (x,z), (x,x)
          if (xd != 0.0) {
            x20 = xd * (wp[2] * sig[CMPC_SIG_23 * hh + ab] + wp[5] * sig[CMPC_SIG_12 * hh + ab]);
            x02 = xd * (wp[2] * sig[CMPC_SIG_23 * hh + ba] + wp[5] * sig[CMPC_SIG_12 * hh + ba]);
            x00 = xd * xd * (wp[2] * sig[CMPC_SIG_33 * hh + ab] + wp[5] * s22);
          }
          const double* pt = sPT + (fi * 4 + fj) * 9;
          const double* po = sPO + (fi * 4 + fj) * 9;
#pragma unroll
          for (int c1 = 0; c1 < 3; c1++)
#pragma unroll
            for (int c2 = 0; c2 < 3; c2++) {
              double pv = 0.0;
              if (c1 == c2) pv = s22 * wp[c1] + s11 * wp[3 + c1];
              if (c1 == 2 && c2 == 0) pv += x20;
              if (c1 == 0 && c2 == 2) pv += x02;
              if (c1 == 0 && c2 == 0) pv += x00;
              double v = 2.0 * (s22 * pt[c1 * 3 + c2] + s11 * po[c1 * 3 + c2] + pv * m2);
              if (j1 == j2 && c1 == c2) { v += alpha2; dmax = fmax(dmax, v); }
              Hs[(3 * j1 + c1) * n + 3 * j2 + c2] = v;
              Hs[(3 * j2 + c2) * n + 3 * j1 + c1] = v;
            }
        }
      }
      {
        int dummy = 0;
        double neg = -dmax;
        block_argmin<NT>(neg, dummy, red, tid);  // max over the CTA; its barriers also order the Hs writes
        dmax = -neg;
      }
      pc.tick(CMPC_PH_HESS);
      // ---- D. K = H^-1 ----
      int e2;
      frexp(dmax, &e2);
      const double scale = ldexp(1.0, -e2);  // exact; scaled diagonal < 1
      if (P.sweep_dmma) {
        // ---- FP64 tensor-core sweep: tiles from the staged H, K back into the staging area, x0 = -K g, one
        //      coalesced copy to the workspace slot ----
        constexpr int TPW = DSweep<S>::TPW, NWARP = DSweep<S>::NW, NTILE = DSweep<S>::NTILE;
        const int lane = tid & 31, w = tid >> 5, r = lane >> 2, q = lane & 3;
        int tI[TPW], tJ[TPW];
        double t[TPW][2];
#pragma unroll
        for (int i = 0; i < TPW; i++) {
          const int k = w + NWARP * i;
          int I = -1, J = 0;
          if (k < NTILE) {
            I = 0;
            while ((I + 1) * (I + 2) / 2 <= k) I++;
            J = k - I * (I + 1) / 2;
          }
          tI[i] = I;
          tJ[i] = J;
          t[i][0] = 0.0;
          t[i][1] = 0.0;
          if (I >= 0) {
            const int ii = 8 * I + r, j0 = 8 * J + 2 * q;
            t[i][0] = (ii < n && j0 < n) ? Hs[ii * n + j0] * scale : (ii == j0 ? 0.5 : 0.0);
            t[i][1] = (ii < n && j0 + 1 < n) ? Hs[ii * n + j0 + 1] * scale : (ii == j0 + 1 ? 0.5 : 0.0);
          }
        }
        pc.tick(CMPC_PH_LOAD);
        __syncthreads();  // every tile is in registers before the staging area is reused
        sweep_dmma<S>(t, tI, tJ, n, tid, pan, mm, dinvs, P.inv_refine);
        pc.tick(CMPC_PH_SWEEP);
        // K_ij = -(A_ij - 2 d_ij) scale, both triangles, into the staging area
#pragma unroll
        for (int i = 0; i < TPW; i++) {
          const int I = tI[i], J = tJ[i];
          if (I < 0) continue;
          const int ii = 8 * I + r, j0 = 8 * J + 2 * q;
          const double k0 = -(t[i][0] - (ii == j0 ? 2.0 : 0.0)) * scale;
          const double k1 = -(t[i][1] - (ii == j0 + 1 ? 2.0 : 0.0)) * scale;
          if (ii < n && j0 < n) {
            Hs[ii * n + j0] = k0;
            if (I != J) Hs[j0 * n + ii] = k0;  // a diagonal tile holds both triangles itself
          }
          if (ii < n && j0 + 1 < n) {
            Hs[ii * n + j0 + 1] = k1;
            if (I != J) Hs[(j0 + 1) * n + ii] = k1;
          }
        }
        __syncthreads();
        for (int j = tid; j < n; j += NT) {
          // x0 = -K g as a compensated dot product (TwoProduct / TwoSum): the terms are ~1e8 times the result, a plain
          // FMA chain would leave its rounding error in the forces
          double hi = 0.0, lo = 0.0;
          for (int i = 0; i < n; i++) {
            const double a = Hs[i * n + j], b = g[i];
            const double p = a * b, pe = fma(a, b, -p);
            const double sum = hi + p, bp = sum - hi;
            lo += ((hi - (sum - bp)) + (p - bp)) + pe;
            hi = sum;
          }
          slot[P.qws_goff + j] = g[j];
          slot[P.qws_goff + P.nmax + j] = -(hi + lo);
        }
        for (int idx = tid; idx < n * n; idx += NT) slot[idx] = Hs[idx];
      } else {
      double A[TM][TN];
#pragma unroll
      for (int a = 0; a < TM; a++)
#pragma unroll
        for (int b = 0; b < TN; b++) {
          const int i = ty + 8 * a, j = tx + TX * b;
          A[a][b] = (i < n && j < n) ? Hs[i * n + j] * scale : (i == j ? 0.5 : 0.0);  // harmless padding beyond n
        }
      pc.tick(CMPC_PH_LOAD);
      sweep_blocked<S>(A, n, tid, pan, mm, dinvs, P.inv_refine);
      pc.tick(CMPC_PH_SWEEP);
      // -swept = (scaled H)^-1 with +2 on the diagonal:  K_ij = -(A_ij - 2 d_ij) scale
      // x0 = -K g: partial column sums over this thread's rows, reduced over ty through shared memory
      {
        double gi[TM];
#pragma unroll
        for (int a = 0; a < TM; a++) gi[a] = g[ty + 8 * a];
#pragma unroll
        for (int b = 0; b < TN; b++) {
          double acc = 0.0;
#pragma unroll
          for (int a = 0; a < TM; a++) acc = fma(A[a][b], gi[a], acc);
          pan[ty * S::PS + tx + TX * b] = acc;
        }
      }
      double* Kg = slot;
#pragma unroll
      for (int a = 0; a < TM; a++)
#pragma unroll
        for (int b = 0; b < TN; b++) {
          const int i = ty + 8 * a, j = tx + TX * b;
          if (i < n && j < n) Kg[(size_t)i * n + j] = -(A[a][b] - (i == j ? 2.0 : 0.0)) * scale;
        }
      __syncthreads();
      for (int j = tid; j < n; j += NT) {
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < 8; t++) acc += pan[t * S::PS + j];
        const double gj = g[j];
        slot[P.qws_goff + j] = gj;
        slot[P.qws_goff + P.nmax + j] = scale * (acc - 2.0 * gj);
      }
      }
      flops_acc += 2.0 * (double)n * n * n * 0.5 + 12.0 * (double)n * n + 2.0 * (double)n * n;
    }
    // contact list for kernel 2
    if (tid == 0) { hdr[0] = nc; hdr[1] = status; }
    {
      unsigned char* hb = reinterpret_cast<unsigned char*>(hdr + 2);
      for (int j = tid; j < nc; j += NT) {
        const int k = fs[j];
        hb[j] = (unsigned char)k;
        hb[CMPC_MAX_FS + j] = gait[k];
      }
    }
    __syncthreads();  // record buffer and work arrays are reused by the next instance
    pc.tick(CMPC_PH_STORE);
    cur = redi[2 + (buf ^ 1)];
  }
#ifdef CMPC_CANARY
  __syncthreads();
  canary_check(smem, cv.guard, cv.nguard, tid, NT, "cmpc_condense_kernel");
#endif
  if (tid == 0 && P.flops && flops_acc > 0.0) atomicAdd(P.flops + CMPC_K_ASSEMBLE, (unsigned long long)flops_acc);
}
