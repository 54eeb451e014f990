// Kernel 1 of the two-kernel pipeline for reduced problems of up to 63 variables (every A1 gait at
// horizon <= 10 except all-feet-down stances): condensation + K = H^-1 + x0 = -K g, one CTA of four
// warps per instance, the rank-8 updates of the blocked symmetric sweep on the FP64 tensor cores
// (DMMA m8n8k4).
//
// Why DMMA (profiles/r1_*): with DFMA register tiles (cmpc_condense.cuh) the sweep is bound by issue
// slots and shared-memory operand traffic (shared pipe 63 %, issue 45 %, FP64 pipe 39 %).  DMMA runs at
// the same FP64 rate on B200 (37 TFLOP/s measured, scripts/ubench) with 1/8 of the instructions and one
// operand fetch per 8 x 8 tile, and it makes symmetric (lower-triangular) storage natural: the matrix
// needs 18 accumulator registers per thread instead of 64, so six instances are resident per SM.
//
// The padded 64 x 64 symmetric matrix is held as its 36 lower-triangular 8 x 8 tiles in DMMA accumulator
// layout (lane l: row l>>2, columns 2(l&3), 2(l&3)+1); warp w owns tile rows w and 7-w (nine tiles).
// Block step s:
//   1. the owner of the diagonal tile (s, s) inverts it in registers (Gauss-Jordan over warp shuffles);
//      the pivot rows are published (tiles (s, J<=s) as they are, tiles (I>s, s) transposed: the matrix
//      is symmetric) as an 8 x 64 panel C in shared memory, with D - I in the diagonal block;
//   2. M = -D^-1 C: 16 DMMAs, two tile columns per warp;
//   3. every tile (I, J) += C_I' M_J: 72 DMMAs, operands fetched once per tile row / column.
// Publishing D - I makes the same update produce the swept pivot rows and columns; every swept diagonal
// entry carries a constant +2 removed at the end; H is scaled by an exact power of two (see cmpc_sweep.cuh).
//
// x0 = -K g rides along for free: n = 3 x contacts is at most 63, so row 63 of the padded matrix is
// never a real variable.  It holds g (the matrix is the symmetric bordered [H g; g' .]) and is never
// pivoted, so after the sweep it reads g' H^-1.
//
// K leaves the kernel tile-major (36 tiles of 64 doubles, lower triangle only); cmpc_dual.cuh reads
// row i of K from tiles (i/8, J) and, transposed, from tiles (I, i/8).
#pragma once

namespace {

constexpr int MMA_NPAD = 64;
constexpr int MMA_PS = 68;  // panel row stride (doubles): = 4 (mod 16), conflict-free fragment loads
constexpr int MMA_NT = 128;
constexpr int MMA_PQ = 292;  // 144 foot-pair / component-pair coefficients + a zero region addressed by padding rows / columns
__host__ __device__ constexpr int tix(int I, int J) { return I * (I + 1) / 2 + J; }

struct MCarve {
  int rec0, rec1, bars, sig, ss, small, evec, agg, fs, rowinfo, g, pq, xtab, pan, mm, dv, red, total;
  CMPC_CANARY_FIELDS
};

__host__ __device__ inline MCarve make_mcarve(int h, int rec_stride, bool adapt) {
  MCarve c;
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
  c.ss = o; o += 16 * h * h;  // (s22, s11) interleaved
  CMPC_GUARD(o, c);
  c.small = o; o += align16(8 * (36 + 36 + 144 + 144 + 24));  // W, RW, PT, PO, scalars
  CMPC_GUARD(o, c);
  c.evec = o; o += align16(8 * 12 * h);
  CMPC_GUARD(o, c);
  c.agg = o; o += align16(8 * 10 * h);
  CMPC_GUARD(o, c);
  c.fs = o; o += align16(CMPC_MAX_FS);
  CMPC_GUARD(o, c);
  c.rowinfo = o; o += 4 * MMA_NPAD;
  CMPC_GUARD(o, c);
  c.g = o; o += 8 * MMA_NPAD;
  CMPC_GUARD(o, c);
  c.pq = o; o += 16 * MMA_PQ;
  CMPC_GUARD(o, c);
  c.xtab = o; o += align16(32 * h * h);  // x_drag couplings: XA[ab], X0[ab], XA[ba], and a table of zeros
  CMPC_GUARD(o, c);
  // scratch of the estimator stage (adaptive launches only)
  c.pan = o;
  c.mm = o;
  c.dv = o;
  if (adapt) o = c.pan + 8 * CMPC_ADAPT_SCRATCH;
  if (adapt) CMPC_GUARD(o, c);
  c.red = o; o += 512;
  CMPC_GUARD(o, c);
  c.total = o;
  return c;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// H[(step a, foot fi, comp c1), (step b, foot fj, comp c2)] from shared-memory tables
struct HessS {
  const double* sig;  // shared, 5 tables of h*h
  const double* sPT;  // shared, [4][4][3][3]
  const double* sPO;
  const double* wp;   // shared, position weights [3] then velocity weights [3]
  int h, hh;
  double xd, m2, alpha2;
};

__device__ __forceinline__ double hess_entry_s(const HessS& C, int ri, int rj, bool diag) {
  const int a = ri & 0xff, fi = (ri >> 8) & 3, c1 = ri >> 16;
  const int b = rj & 0xff, fj = (rj >> 8) & 3, c2 = rj >> 16;
  const int ab = a * C.h + b;
  const double s11 = C.sig[CMPC_SIG_11 * C.hh + ab], s22 = C.sig[CMPC_SIG_22 * C.hh + ab];
  const int pidx = (fi * 4 + fj) * 9 + c1 * 3 + c2;
  double val = s22 * C.sPT[pidx] + s11 * C.sPO[pidx];
  double pv = 0.0;
  if (c1 == c2) pv = s22 * C.wp[c1] + s11 * C.wp[3 + c1];
  if (C.xd != 0.0) {
    const int ba = b * C.h + a;
    if (c1 == 2 && c2 == 0) pv += C.xd * (C.wp[2] * C.sig[CMPC_SIG_23 * C.hh + ab] + C.wp[5] * C.sig[CMPC_SIG_12 * C.hh + ab]);
    if (c1 == 0 && c2 == 2) pv += C.xd * (C.wp[2] * C.sig[CMPC_SIG_23 * C.hh + ba] + C.wp[5] * C.sig[CMPC_SIG_12 * C.hh + ba]);
    if (c1 == 0 && c2 == 0) pv += C.xd * C.xd * (C.wp[2] * C.sig[CMPC_SIG_33 * C.hh + ab] + C.wp[5] * s22);
  }
  val = 2.0 * (val + pv * C.m2);
  if (diag) val += C.alpha2;
  return val;
}

// 8 x 8 inverse of a tile held in accumulator layout (lane (r, q): D[r][2q], D[r][2q+1]) by Gauss-Jordan
// without pivoting (D is a Schur complement of an SPD matrix), operands exchanged by shuffles.
__device__ __forceinline__ void warp_inv8_acc(double& a0, double& a1, int r, int q) {
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

// the same chain in fp32 (a starting value for Newton steps in FP64, see cmpc_invert_ws_kernel)
__device__ __forceinline__ void warp_inv8_acc_f32(float& a0, float& a1, int r, int q) {
#pragma unroll
  for (int p = 0; p < 8; p++) {
    const int pq = p >> 1;
    const float mine = (p & 1) ? a1 : a0;
    const float arp = __shfl_sync(0xffffffffu, mine, r * 4 + pq);
    const float dpp = __shfl_sync(0xffffffffu, mine, p * 4 + pq);
    const float ap0 = __shfl_sync(0xffffffffu, a0, p * 4 + q);
    const float ap1 = __shfl_sync(0xffffffffu, a1, p * 4 + q);
    float dinv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(dinv) : "f"(dpp));
    const float t = arp * dinv;
    float n0 = fmaf(-t, ap0, a0), n1 = fmaf(-t, ap1, a1);
    if (r == p) { n0 = ap0 * dinv; n1 = ap1 * dinv; }
    if (q == pq) {
      if (p & 1) n1 = (r == p) ? dinv : -t;
      else n0 = (r == p) ? dinv : -t;
    }
    a0 = n0;
    a1 = n1;
  }
}

// free response of state component c at step rr (0-based), weighted tracking error against the reference trajectory
__device__ __forceinline__ double tracking_error(const float* rec, const double* sScal, int idx, double dt, double gravity) {
  const double xd = rec[CMPC_REC_XDRAG];
  const double ffx = rec[CMPC_REC_FDIST + 3];
  const double az = xd * (double)rec[CMPC_REC_V + 0] + gravity;  // row 11 of A x0
  const int rr = idx / 12, c = idx - 12 * rr;
  const double T = (double)(rr + 1) * dt, T2 = 0.5 * T * T;
  double val;
  if (c < 3) {
    val = sScal[8 + c] + T * sScal[16 + c] + T2 * sScal[19 + c];
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
  return (double)rec[CMPC_REC_WEIGHTS + c] * (val - (double)rec[CMPC_REC_TRAJ + idx]);
}

}  // namespace

template <bool ADAPT, int MINB>
__global__ void __launch_bounds__(MMA_NT, MINB) cmpc_assemble_mma_kernel(const __grid_constant__ CmpcParams P) {
  constexpr int NT = MMA_NT, PS = MMA_PS;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int r = lane >> 2, q = lane & 3;
  const int h = P.horizon, hh = h * h;
  const MCarve cv = make_mcarve(h, P.rec_stride, ADAPT);
#ifdef CMPC_CANARY
  canary_fill(smem, cv.guard, cv.nguard, threadIdx.x, MMA_NT);
  __syncthreads();
#endif
  unsigned char* recbuf[2] = {smem + cv.rec0, smem + cv.rec1};
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + cv.bars);
  double* sig = reinterpret_cast<double*>(smem + cv.sig);
  double2* SS = reinterpret_cast<double2*>(smem + cv.ss);
  double* sW = reinterpret_cast<double*>(smem + cv.small);
  double* sRW = sW + 36;
  double* sPT = sRW + 36;
  double* sPO = sPT + 144;
  double* sScal = sPO + 144;  // [1..6] position / velocity weights, [8..10] roll pitch yaw, [16..21] R'omega, R'tau_xi
  double* ev = reinterpret_cast<double*>(smem + cv.evec);
  double* agg = reinterpret_cast<double*>(smem + cv.agg);
  unsigned char* fs = smem + cv.fs;
  int* rowinfo = reinterpret_cast<int*>(smem + cv.rowinfo);
  double* g = reinterpret_cast<double*>(smem + cv.g);
  double2* PQ = reinterpret_cast<double2*>(smem + cv.pq);
  double* xtab = reinterpret_cast<double*>(smem + cv.xtab);
  double* pan = reinterpret_cast<double*>(smem + cv.pan);
  double* mm = reinterpret_cast<double*>(smem + cv.mm);
  double* dv = reinterpret_cast<double*>(smem + cv.dv);
  double* red = reinterpret_cast<double*>(smem + cv.red);
  int* redi = reinterpret_cast<int*>(red + 32);  // [0] = nc, [2], [3] = next instance (double-buffered)

  const int count = P.count;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < CMPC_SIG_COUNT * hh; i += NT) sig[i] = __ldg(P.sigma + i);
  for (int i = tid; i < hh; i += NT)
    SS[i] = make_double2(__ldg(P.sigma + CMPC_SIG_22 * hh + i), __ldg(P.sigma + CMPC_SIG_11 * hh + i));
  for (int i = 144 + tid; i < MMA_PQ; i += NT) PQ[i] = make_double2(0.0, 0.0);
  for (int i = tid; i < hh; i += NT) xtab[3 * hh + i] = 0.0;  // what a component pair without an x_drag coupling adds
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
  const int ILO = w, IHI = 7 - w;  // this warp's tile rows

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

    // ---- 0. periodic-disturbance estimator (Adaptive MPC), SolverMPC.cpp:688-798 ----
    if (ADAPT) {
      double* est_s = red + 40;
      if (P.adapt_mode == 0 || P.adapt_mode == 1) {
        estimate_disturbance<NT>(P, inst, tid, pan, red, est_s);
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

    // ---- A. contact foot-steps (warp 0); Euler angles (three lanes of warp 1, one code path);
    //         W_f = I^-1 [r_f]x and R^T W_f (threads 32..103); R'omega, R'tau (104..109); weights (110..115) ----
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
        if (keep) fs[cnt + __popc(mask & ((1u << tid) - 1u))] = (unsigned char)k;
        cnt += __popc(mask);
      }
      if (tid == 0) redi[0] = cnt;
    } else {
      const double qw = rec[CMPC_REC_Q + 0], qx = rec[CMPC_REC_Q + 1], qy = rec[CMPC_REC_Q + 2], qz = rec[CMPC_REC_Q + 3];
      if (tid < 35) {
        // quat_to_rpy, SolverMPC.cpp:352-361; asin(v) evaluated as atan2(v, sqrt(1 - v^2)) so the lanes share a path
        const int which = tid - 32;
        double y, x;
        if (which == 0) { y = 2.0 * (qy * qz + qw * qx); x = qw * qw - qx * qx - qy * qy + qz * qz; }
        else if (which == 1) { y = fmin(-2.0 * (qx * qz - qw * qy), 0.99999); x = sqrt(fma(-y, y, 1.0)); }
        else { y = 2.0 * (qx * qy + qw * qz); x = qw * qw + qx * qx - qy * qy - qz * qz; }
        sScal[8 + which] = atan2(y, x);
      }
      double R[9];
      {
        const double tx2 = 2 * qx, ty2 = 2 * qy, tz2 = 2 * qz;
        const double twx = tx2 * qw, twy = ty2 * qw, twz = tz2 * qw, txx = tx2 * qx, txy = ty2 * qx, txz = tz2 * qx;
        const double tyy = ty2 * qy, tyz = tz2 * qy, tzz = tz2 * qz;
        R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
        R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
        R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
      }
      if (tid < 104) {
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
        const int e = tid - 32;
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
      } else if (tid < 110) {
        // (R' omega)[c], (R' tau_xi)[c]
        const int t = tid - 104, c = t % 3;
        const int o = (t < 3) ? CMPC_REC_W : CMPC_REC_FDIST;
        // (column c of R picked with integer masks: R[c] with a run-time c would move all of R to local memory)
        sScal[16 + t] = pick3(R[0], R[1], R[2], c) * (double)rec[o + 0] + pick3(R[3], R[4], R[5], c) * (double)rec[o + 1] +
                        pick3(R[6], R[7], R[8], c) * (double)rec[o + 2];
      } else if (tid < 116) {
        const int t = tid - 110;
        sScal[1 + t] = (double)rec[CMPC_REC_WEIGHTS + (t < 3 ? 3 + t : 6 + t)];  // position, velocity weights
      }
    }
    __syncthreads();
    const int nc = redi[0];
    const int n = 3 * nc;
    int status = CMPC_ST_SOLVED;
    if (nc == 0) status = CMPC_ST_EMPTY;
    else if (n > P.nmax || n >= MMA_NPAD) status = CMPC_ST_CAPACITY;

    if (status == CMPC_ST_SOLVED) {
      // ---- B. weighted tracking error of the free response e_r = S (Adt^(r+1) x0 + sum_k Adt^k Qdt xi - Xd_r);
      //         foot-pair blocks PT = RW_i' S_theta RW_j, PO = W_i' S_omega W_j; variable -> (step, foot, comp) ----
      for (int idx = tid; idx < 12 * h; idx += NT) ev[idx] = tracking_error(rec, sScal, idx, dt, P.gravity);
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
      if (tid < MMA_NPAD) {
        int info = -1;
        if (tid < n) {
          const int j = tid / 3, comp = tid - 3 * j, k = fs[j];
          info = (k >> 2) | ((k & 3) << 8) | (comp << 16);
        }
        rowinfo[tid] = info;
      }
      __syncthreads();
      // ---- C. horizon aggregates of e:  agg[c][0:3]=sum c2 e_theta, [3:6]=sum c1 e_omega,
      //         [6:9]=(sum c2 e_p + c1 e_v)/m, [9]=xd/m (sum c3 e_pz + c2 e_vz);  largest diagonal entry of H ----
      HessS C;
      C.sig = sig; C.sPT = sPT; C.sPO = sPO; C.wp = sScal + 1; C.h = h; C.hh = hh;
      C.xd = rec[CMPC_REC_XDRAG];
      C.m2 = minv * minv;
      C.alpha2 = 2.0 * (double)rec[CMPC_REC_ALPHA];
      {
        const double xd = C.xd;
        for (int idx = tid; idx < 10 * h; idx += NT) {
          const int c = idx / 10, comp = idx - 10 * c;
          double acc = 0.0;
          for (int rr = c; rr < h; rr++) {
            const double tau = (double)(rr - c) * dt;
            const double c1 = dt, c2 = tau * dt + 0.5 * dt * dt,
                         c3 = 0.5 * tau * tau * dt + 0.5 * tau * dt * dt + dt * dt * dt / 6.0;
            const double* e = ev + 12 * rr;
            if (comp < 3) acc += c2 * e[comp];
            else if (comp < 6) acc += c1 * e[6 + comp - 3];
            else if (comp < 9) acc += (c2 * e[3 + comp - 6] + c1 * e[9 + comp - 6]) * minv;
            else acc += (c3 * e[5] + c2 * e[11]) * xd * minv;
          }
          agg[idx] = acc;
        }
      }
      double dmax = 0.0;
      if (tid >= 64 && tid - 64 < n) {
        const int ri = rowinfo[tid - 64];
        dmax = hess_entry_s(C, ri, ri, true);
      }
      {
        int dummy = 0;
        double neg = -dmax;
        block_argmin<NT>(neg, dummy, red, tid);
        dmax = -neg;
      }
      int e2;
      frexp(dmax, &e2);
      const double scale = ldexp(1.0, -e2);  // exact; scaled diagonal < 1
      // ---- gradient; scaled coefficient tables: H_ij scale = s22[ab] PQ[p].x + s11[ab] PQ[p].y (+ x_drag couplings)
      //      with p = (fi, c1, fj, c2); the block-diagonal position / velocity terms are folded into PQ ----
      const double sc2 = 2.0 * scale;
      const bool drag = (C.xd != 0.0);
      if (tid < MMA_NPAD) {
        const int I = tid;
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
      for (int e = tid; e < 144; e += NT) {
        const int c1 = (e % 9) / 3, c2 = e % 3;
        double pt = sPT[e], po = sPO[e];
        if (c1 == c2) { pt += C.m2 * C.wp[c1]; po += C.m2 * C.wp[3 + c1]; }
        PQ[e] = make_double2(sc2 * pt, sc2 * po);
      }
      if (drag) {
        for (int ab = tid; ab < hh; ab += NT) {
          const int ba = (ab % h) * h + ab / h;
          xtab[ab] = sc2 * C.m2 * C.xd * (C.wp[2] * sig[CMPC_SIG_23 * hh + ab] + C.wp[5] * sig[CMPC_SIG_12 * hh + ab]);
          xtab[hh + ab] = sc2 * C.m2 * C.xd * C.xd * (C.wp[2] * sig[CMPC_SIG_33 * hh + ab] + C.wp[5] * sig[CMPC_SIG_22 * hh + ab]);
          xtab[2 * hh + ab] = sc2 * C.m2 * C.xd * (C.wp[2] * sig[CMPC_SIG_23 * hh + ba] + C.wp[5] * sig[CMPC_SIG_12 * hh + ba]);
        }
      }
      __syncthreads();
      pc.tick(CMPC_PH_PREP);

      // ---- D. this warp's tiles of the bordered matrix [H g; g' .] (row 63 = g, never pivoted) ----
      const int nblk = (n + 7) >> 3;
      // tile rows ILO (J <= ILO, J < 4) and IHI (J <= IHI); a tile pair leaves for the workspace slot (accumulator
      // layout, scaled) as soon as it is formed — held back in arrays the 24 doubles went to local memory
      {
        int rah[2], rpo[2], xs0[2], xs2[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
          const int info = rowinfo[8 * (t ? IHI : ILO) + r];
          const bool ok = info >= 0;
          const int c1 = info >> 16;
          rah[t] = ok ? (info & 0xff) * h : 0;
          rpo[t] = ok ? ((info >> 8) & 3) * 36 + c1 * 3 : 144;
          // table of the x_drag coupling this row has with an x column / a z column
          xs0[t] = (ok && c1 == 0) ? hh : ((ok && c1 == 2) ? 0 : 3 * hh);
          xs2[t] = (ok && c1 == 0) ? 2 * hh : 3 * hh;
        }
        const double alpha2s = C.alpha2 * scale;
        const bool brow = (w == 0 && r == 7);  // row 63 lives in warp 0's tile row 7
#pragma unroll
        for (int J = 0; J < 8; J++) {
          if (J <= IHI) {
            const int2 info2 = *reinterpret_cast<const int2*>(rowinfo + 8 * J + 2 * q);
            double vv[2][2] = {{0.0, 0.0}, {0.0, 0.0}};  // [t][e]
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int info = e ? info2.y : info2.x;
              const bool ok = info >= 0;
              const int c2 = info >> 16;
              const int cb = ok ? (info & 0xff) : 0;
              const int cpo = ok ? ((info >> 8) & 3) * 9 + c2 : 144;
              const bool k0 = ok && c2 == 0, k2 = ok && c2 == 2;
#pragma unroll
              for (int t = 0; t < 2; t++) {
                if (t == 0 && (J >= 4 || J > ILO)) continue;
                const int I = t ? IHI : ILO;
                const double2 ss = SS[rah[t] + cb];
                const double2 pq = PQ[rpo[t] + cpo];
                double v = fma(ss.x, pq.x, ss.y * pq.y);
                if (drag) {
                  // (c1, c2) = (x, x): X0[ab]; (z, x): XA[ab]; (x, z): XA[ba]; every other pair: the zero table
                  const int tsel = k0 ? xs0[t] : (k2 ? xs2[t] : 3 * hh);
                  v += xtab[tsel + rah[t] + cb];
                }
                const int i = 8 * I + r, j = 8 * J + 2 * q + e;
                if (I == J && i == j) v += (i < n) ? alpha2s : 0.5;  // alpha; 1/2 on the padding diagonal
                if (t == 1 && brow) v = (j < n) ? g[j] : (j == 63 ? 0.5 : 0.0);                // border row
                if (t == 1 && I == 7 && J == 7 && j == 63 && i < 63) v = (i < n) ? g[i] : 0.0;  // border column in tile (7,7)
                vv[t][e] = v;
              }
            }
            if (J < 4 && J <= ILO) *reinterpret_cast<double2*>(slot + tix(ILO, J) * 64 + lane * 2) = make_double2(vv[0][0], vv[0][1]);
            *reinterpret_cast<double2*>(slot + tix(IHI, J) * 64 + lane * 2) = make_double2(vv[1][0], vv[1][1]);
          }
        }
      }
      {
        for (int j = tid; j < n; j += NT) slot[P.qws_goff + j] = g[j];
        if (tid == 0) slot[P.qws_goff + 2 * P.nmax] = scale;
      }
      pc.tick(CMPC_PH_HESS);
      flops_acc += 12.0 * (double)n * n;  // assembly; the inversion kernel accounts for n^3 + 2 n^2
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
  canary_check(smem, cv.guard, cv.nguard, tid, NT, "cmpc_assemble_mma_kernel");
#endif
  if (tid == 0 && P.flops && flops_acc > 0.0) atomicAdd(P.flops + CMPC_K_ASSEMBLE, (unsigned long long)flops_acc);
}
