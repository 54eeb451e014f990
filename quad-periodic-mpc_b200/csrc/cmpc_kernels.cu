// Fused condensation + dense QP kernel of the B200 batched convex-MPC engine.
//
// One CTA owns one MPC instance at a time (persistent grid-stride loop, next
// instance record prefetched into shared memory by cp.async.bulk while the
// current one is solved).  Per instance, entirely in shared memory / registers:
//
//   A. state build          RobotState::set + quat_to_rpy      RobotState.cpp:10, SolverMPC.cpp:352
//   B. ct_ss_mats + c2qp    closed form of the 31x31 exp()      SolverMPC.cpp:96, :260
//   C. H, g of the reduced  2(Bqp'SBqp + aI), 2Bqp'S(Aqp x0 +   SolverMPC.cpp:806-814 and the
//      (contact-only) QP    Qqp xi - Xd), swing feet eliminated swing elimination at :859-950
//   D. K = H^-1             symmetric sweep; the matrix lives in REGISTER TILES (TM x TN per
//                           thread, cyclic layout), one pivot column broadcast through shared
//                           memory per step
//   E. QP                   Goldfarb-Idnani dual active set,    replaces qpOASES, SolverMPC.cpp:955
//                           range-space form on K
//   F. q_soln scatter, objective, primal activity mask          SolverMPC.cpp:970-983
//
// All arithmetic is FP64 on FP32 inputs (the reference computes A..C in FP32
// and hands qpOASES doubles).  DESIGN.md §3 derives the closed forms.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "cmpc_device.h"

namespace {

// ---------------------------------------------------------------------------
// kernel shapes.  TY x TX threads, each holding a TM x TN register tile of the
// NPAD x NPAD (padded) Hessian; REG = false keeps the matrix in shared memory
// (any n that fits), used beyond the register tiers.
// ---------------------------------------------------------------------------
template <int TY_, int TX_, int TM_, int TN_, bool REG_, int MINB_, bool GMEM_ = false>
struct Shape {
  static constexpr int TY = TY_, TX = TX_, TM = TM_, TN = TN_, NT = TY_ * TX_, NPAD = TY_ * TM_, MINB = MINB_;
  static constexpr bool REG = REG_;
  static constexpr bool GMEM = GMEM_;  // K and the working-set inverse live in a per-CTA global workspace
  static_assert(TY_ * TM_ == TX_ * TN_, "square padded matrix");
  static_assert(TX_ % TY_ == 0, "column owner derivable from the row block");
};
using Shape64 = Shape<8, 8, 8, 8, true, 4>;       // n <= 64, 64 threads, 8x8 tiles
using Shape64w = Shape<8, 16, 8, 4, true, 4>;     // n <= 64, 128 threads, 8x4 tiles
using Shape128 = Shape<16, 16, 8, 8, true, 1>;    // n <= 128, 256 threads, 8x8 tiles
using ShapeMem = Shape<8, 16, 8, 4, false, 1>;    // any n that fits shared memory, 128 threads
using ShapeGmem = Shape<8, 16, 8, 4, false, 1, true>;  // beyond shared memory: L2-resident workspace

// ---------------------------------------------------------------------------
// shared memory carve-up (same arithmetic on host and device)
// ---------------------------------------------------------------------------
struct Carve {
  int rec0, rec1, bars, small, evec, agg, fs, fsinv, rowinfo, cbuf, K, g, x, kn, z, v, s, rc, isact, act, u, d, r,
      col, Pp, red, total;
  CMPC_CANARY_FIELDS
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline Carve make_carve(int h, int nmax, int qcap, int rec_stride, int npad, bool adapt, bool gmem) {
  Carve c;
  int o = 0;
  CMPC_GUARD_INIT(c);
  int nc = nmax / 3, m = 5 * nc;
  c.rec0 = o; o += align16(rec_stride);
  CMPC_GUARD(o, c);
  c.rec1 = o; o += align16(rec_stride);
  CMPC_GUARD(o, c);
  c.bars = o; o += 16;
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
  c.rowinfo = o; o += align16(4 * (npad > nmax ? npad : nmax));
  CMPC_GUARD(o, c);
  c.cbuf = o; o += align16(8 * (2 * (npad + 2) + 2 * npad));  // pivot rows (x2) + diagonal copies (x2)
  CMPC_GUARD(o, c);
  {
    int kb = gmem ? 0 : 8 * nmax * nmax;  // the estimator stage borrows this region for 3 x 400 doubles of work space
    if (adapt && kb < 8 * CMPC_ADAPT_SCRATCH) kb = 8 * CMPC_ADAPT_SCRATCH;
    c.K = o; o += align16(kb);
    CMPC_GUARD(o, c);
  }
  c.g = o; o += align16(8 * nmax);
  CMPC_GUARD(o, c);
  c.x = o; o += align16(8 * nmax);
  CMPC_GUARD(o, c);
  c.kn = o; o += align16(8 * nmax);
  CMPC_GUARD(o, c);
  c.z = o; o += align16(8 * nmax);
  CMPC_GUARD(o, c);
  c.v = o; o += align16(8 * nmax);
  CMPC_GUARD(o, c);
  c.s = o; o += align16(8 * m);
  CMPC_GUARD(o, c);
  c.rc = o; o += align16(8 * m);
  CMPC_GUARD(o, c);
  c.isact = o; o += align16(m);
  CMPC_GUARD(o, c);
  c.act = o; o += align16(2 * (qcap + 1));
  CMPC_GUARD(o, c);
  c.u = o; o += align16(8 * (qcap + 1));
  CMPC_GUARD(o, c);
  c.d = o; o += align16(8 * (qcap + 1));
  CMPC_GUARD(o, c);
  c.r = o; o += align16(8 * (qcap + 1));
  CMPC_GUARD(o, c);
  c.col = o; o += align16(8 * (qcap + 1));
  CMPC_GUARD(o, c);
  c.Pp = o; o += gmem ? 0 : align16(8 * ((qcap + 1) * (qcap + 2) / 2));
  CMPC_GUARD(o, c);
  c.red = o; o += 512;
  CMPC_GUARD(o, c);
  c.total = o;
  return c;
}

}  // namespace

#include "cmpc_common.cuh"

namespace {

// D (shared-memory tier): same sweep with the matrix in shared memory
template <int NT>
__device__ __forceinline__ void build_invert_smem(const HessCtx& C, const int* rowinfo, int n, int tid, double* kn,
                                                  double* K) {
  for (int i = 0; i < n; i++) {
    const int ri = rowinfo[i];
    for (int j = tid; j < n; j += NT) K[i * n + j] = hess_entry(C, ri, rowinfo[j], i == j);
  }
  __syncthreads();
  for (int k = 0; k < n; k++) {
    for (int i = tid; i < n; i += NT) kn[i] = K[k * n + i];
    __syncthreads();
    const double dinv = 1.0 / kn[k];
    for (int j = tid; j < n; j += NT) {
      const double cjd = kn[j] * dinv;
      if (j == k) {
        for (int i = 0; i < n; i++) K[i * n + j] = (i == k) ? -dinv : kn[i] * dinv;
      } else {
        for (int i = 0; i < n; i++) K[i * n + j] = (i == k) ? cjd : fma(-kn[i], cjd, K[i * n + j]);
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < n * n; idx += NT) K[idx] = -K[idx];
}

}  // namespace

#include "cmpc_sweep.cuh"
#include "cmpc_adapt.cuh"
#include "cmpc_qp.cuh"

template <class S, bool ADAPT>
__global__ void __launch_bounds__(S::NT, S::MINB) cmpc_solve_kernel(const __grid_constant__ CmpcParams P) {
  constexpr int NT = S::NT;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int h = P.horizon;
  constexpr bool gmem = S::GMEM;  // compile-time, so that K stays a shared-window pointer (LDS, not generic LD) otherwise
  const Carve cv = make_carve(h, P.nmax, P.qcap, P.rec_stride, S::NPAD, ADAPT, gmem);
#ifdef CMPC_CANARY
  canary_fill(smem, cv.guard, cv.nguard, tid, NT);
  __syncthreads();
#endif
  unsigned char* recbuf[2] = {smem + cv.rec0, smem + cv.rec1};
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + cv.bars);
  double* sW = reinterpret_cast<double*>(smem + cv.small);  // W[4][3][3]
  double* sRW = sW + 36;                                    // (R^T W)[4][3][3]
  double* sPT = sRW + 36;                                   // PT[4][4][3][3]
  double* sPO = sPT + 144;                                  // PO[4][4][3][3]
  double* sScal = sPO + 144;                                // broadcast scalars
  double* ev = reinterpret_cast<double*>(smem + cv.evec);   // e[h][12]
  double* agg = reinterpret_cast<double*>(smem + cv.agg);   // agg[h][10]
  int* fs = reinterpret_cast<int*>(smem + cv.fs);           // reduced foot-step -> global foot-step k
  int* fsinv = reinterpret_cast<int*>(smem + cv.fsinv);     // global foot-step -> reduced or -1
  int* rowinfo = reinterpret_cast<int*>(smem + cv.rowinfo); // reduced variable -> step | foot<<8 | comp<<16
  double* cbuf = reinterpret_cast<double*>(smem + cv.cbuf);
  double* K;
  if constexpr (gmem) K = P.gws + (size_t)blockIdx.x * P.gws_stride;
  else K = reinterpret_cast<double*>(smem + cv.K);
  double* g = reinterpret_cast<double*>(smem + cv.g);
  double* x = reinterpret_cast<double*>(smem + cv.x);
  double* kn = reinterpret_cast<double*>(smem + cv.kn);
  double* z = reinterpret_cast<double*>(smem + cv.z);
  double* vv = reinterpret_cast<double*>(smem + cv.v);
  double* s = reinterpret_cast<double*>(smem + cv.s);
  double* rc = reinterpret_cast<double*>(smem + cv.rc);
  unsigned char* isact = smem + cv.isact;
  short* act = reinterpret_cast<short*>(smem + cv.act);
  double* u = reinterpret_cast<double*>(smem + cv.u);
  double* dvec = reinterpret_cast<double*>(smem + cv.d);
  double* rvec = reinterpret_cast<double*>(smem + cv.r);
  double* col = reinterpret_cast<double*>(smem + cv.col);
  double* Pp;
  if constexpr (gmem) Pp = K + (size_t)P.nmax * P.nmax;
  else Pp = reinterpret_cast<double*>(smem + cv.Pp);
  double* red = reinterpret_cast<double*>(smem + cv.red);
  int* redi = reinterpret_cast<int*>(red + 32);  // shared ints: [0]=nc

  const int count = P.count_ptr ? min(*P.count_ptr, P.count) : P.count;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  uint32_t phase[2] = {0u, 0u};
  int slot = blockIdx.x;
  if (slot < count && tid == 0) {
    int inst = P.worklist ? P.worklist[slot] : slot;
    mbar_expect_tx(&bars[0], (uint32_t)P.rec_stride);
    bulk_g2s(recbuf[0], P.records + (size_t)inst * P.rec_stride, (uint32_t)P.rec_stride, &bars[0]);
  }

  const double dt = P.dt, mu_inv = P.mu_inv, minv = P.mass_inv;
  double flops_acc = 0.0;
  PhaseClock pc;
  pc.init(P.phase_cycles, reinterpret_cast<long long*>(red + 48), tid);

  for (int buf = 0; slot < count; slot += gridDim.x, buf ^= 1) {
    const int inst = P.worklist ? P.worklist[slot] : slot;
    // prefetch the next record while this one is solved
    {
      int nslot = slot + gridDim.x;
      if (tid == 0 && nslot < count) {
        int ninst = P.worklist ? P.worklist[nslot] : nslot;
        fence_proxy_async();
        mbar_expect_tx(&bars[buf ^ 1], (uint32_t)P.rec_stride);
        bulk_g2s(recbuf[buf ^ 1], P.records + (size_t)ninst * P.rec_stride, (uint32_t)P.rec_stride, &bars[buf ^ 1]);
      }
    }
    mbar_wait(&bars[buf], phase[buf]);
    phase[buf] ^= 1u;
    pc.tick(CMPC_PH_WAIT);
    const float* rec = reinterpret_cast<const float*>(recbuf[buf]);
    const unsigned char* gait = recbuf[buf] + 4 * (CMPC_REC_TRAJ + 12 * h);

    // ---- 0. periodic-disturbance estimator (Adaptive MPC): xi for this instance, SolverMPC.cpp:688-798 ----
    if (ADAPT) {
      double* est_s = red + 40;  // 4 doubles of the reduction scratch
      if (P.adapt_mode == 0 || P.adapt_mode == 1) {
        estimate_disturbance<NT>(P, inst, tid, reinterpret_cast<double*>(smem + cv.K), red, est_s);
      } else {
        if (tid < 4) est_s[tid] = P.est[(size_t)inst * 4 + tid];
        __syncthreads();
      }
      if (tid == 0) {
        // compensatory_force = amp + sin(2 pi t f + phase), written to f_est[3] (SolverMPC.cpp:766-772)
        const double simt = (double)P.sim_time[inst];
        const float comp = (float)(est_s[1] + sin(2.0 * M_PI * simt * est_s[2] + est_s[3]));
        float* fe = P.f_est + (size_t)inst * 6;
        fe[3] = comp;
        if (P.adapt_mode == 0 || P.adapt_mode == 1)
          for (int i = 0; i < 4; i++) P.est[(size_t)inst * 4 + i] = est_s[i];
        if (P.adapt_mode >= 1) {  // g sees Q_qp * f_est (SolverMPC.cpp:810): overwrite the record's xi slots
          float* xi = reinterpret_cast<float*>(recbuf[buf]) + CMPC_REC_FDIST;
          for (int i = 0; i < 6; i++) xi[i] = (i == 3) ? comp : fe[i];
        }
      }
      __syncthreads();
      pc.tick(CMPC_PH_ADAPT);
    }

    // ---- A1. contact foot-steps (the reference keeps a foot-step unless its fz bound is ~0) ----
    if (tid < 32) {
      int cnt = 0;
      for (int base = 0; base < 4 * h; base += 32) {
        int k = base + tid;
        bool keep = false;
        if (k < 4 * h) {
          double ub = (double)gait[k] * P.f_max;
          keep = !(ub < 0.01 && ub > -0.01);
        }
        unsigned mask = __ballot_sync(0xffffffffu, keep);
        int pos = cnt + __popc(mask & ((1u << tid) - 1u));
        if (k < 4 * h) fsinv[k] = keep ? pos : -1;
        if (keep) fs[pos] = k;
        cnt += __popc(mask);
      }
      if (tid == 0) redi[0] = cnt;
    }
    // ---- A2. rotation, inertia: per-thread copies (cheap, avoids a round of syncs) ----
    double R[9], Ii[9];
    {
      double qw = rec[CMPC_REC_Q + 0], qx = rec[CMPC_REC_Q + 1], qy = rec[CMPC_REC_Q + 2], qz = rec[CMPC_REC_Q + 3];
      double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
      double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx;
      double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
      R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
      R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
      R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
      double Iw[9];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
          Iw[i * 3 + j] = R[i * 3 + 0] * P.inertia[0] * R[j * 3 + 0] + R[i * 3 + 1] * P.inertia[1] * R[j * 3 + 1] +
                          R[i * 3 + 2] * P.inertia[2] * R[j * 3 + 2];
      double c00 = Iw[4] * Iw[8] - Iw[5] * Iw[7], c01 = Iw[5] * Iw[6] - Iw[3] * Iw[8], c02 = Iw[3] * Iw[7] - Iw[4] * Iw[6];
      double id = 1.0 / (Iw[0] * c00 + Iw[1] * c01 + Iw[2] * c02);
      Ii[0] = c00 * id; Ii[1] = (Iw[2] * Iw[7] - Iw[1] * Iw[8]) * id; Ii[2] = (Iw[1] * Iw[5] - Iw[2] * Iw[4]) * id;
      Ii[3] = c01 * id; Ii[4] = (Iw[0] * Iw[8] - Iw[2] * Iw[6]) * id; Ii[5] = (Iw[2] * Iw[3] - Iw[0] * Iw[5]) * id;
      Ii[6] = c02 * id; Ii[7] = (Iw[1] * Iw[6] - Iw[0] * Iw[7]) * id; Ii[8] = (Iw[0] * Iw[4] - Iw[1] * Iw[3]) * id;
    }
    // W_f = I^-1 [r_f]x  and  RW_f = R^T W_f
    for (int e = tid; e < 72; e += NT) {
      int which = e / 36, ee = e - 36 * which;
      int f = ee / 9, i = (ee % 9) / 3, j = ee % 3;
      double rx = rec[CMPC_REC_R + 0 * 4 + f], ry = rec[CMPC_REC_R + 1 * 4 + f], rz = rec[CMPC_REC_R + 2 * 4 + f];
      // column j of the cross-product matrix [r]x
      double c0 = (j == 0) ? 0.0 : (j == 1 ? -rz : ry);
      double c1 = (j == 0) ? rz : (j == 1 ? 0.0 : -rx);
      double c2 = (j == 0) ? -ry : (j == 1 ? rx : 0.0);
      double w0 = Ii[0] * c0 + Ii[1] * c1 + Ii[2] * c2;
      double w1 = Ii[3] * c0 + Ii[4] * c1 + Ii[5] * c2;
      double w2 = Ii[6] * c0 + Ii[7] * c1 + Ii[8] * c2;
      double r0 = (i == 0) ? R[0] : (i == 1 ? R[1] : R[2]);
      double r1 = (i == 0) ? R[3] : (i == 1 ? R[4] : R[5]);
      double r2 = (i == 0) ? R[6] : (i == 1 ? R[7] : R[8]);
      if (which == 0) sW[f * 9 + i * 3 + j] = (i == 0) ? w0 : (i == 1 ? w1 : w2);
      else sRW[f * 9 + i * 3 + j] = r0 * w0 + r1 * w1 + r2 * w2;
    }
    // ---- B/C. weighted tracking error of the free response, e_r = S (Adt^(r+1) x0 + sum_k Adt^k Qdt xi - Xd_r) ----
    {
      const double xd = rec[CMPC_REC_XDRAG];
      // quat_to_rpy, SolverMPC.cpp:352-361 (x0 = roll, pitch, yaw, p, omega, v, g)
      double qw = rec[CMPC_REC_Q + 0], qx = rec[CMPC_REC_Q + 1], qy = rec[CMPC_REC_Q + 2], qz = rec[CMPC_REC_Q + 3];
      double as = fmin(-2.0 * (qx * qz - qw * qy), 0.99999);
      double yaw = atan2(2.0 * (qx * qy + qw * qz), qw * qw + qx * qx - qy * qy - qz * qz);
      double pitch = asin(as);
      double roll = atan2(2.0 * (qy * qz + qw * qx), qw * qw - qx * qx - qy * qy + qz * qz);
      const double om0 = rec[CMPC_REC_W + 0], om1 = rec[CMPC_REC_W + 1], om2 = rec[CMPC_REC_W + 2];
      const double ft0 = rec[CMPC_REC_FDIST + 0], ft1 = rec[CMPC_REC_FDIST + 1], ft2 = rec[CMPC_REC_FDIST + 2];
      const double ffx = rec[CMPC_REC_FDIST + 3];
      const double az = xd * (double)rec[CMPC_REC_V + 0] + P.gravity;  // row 11 of A x0
      for (int idx = tid; idx < 12 * h; idx += NT) {
        int r = idx / 12, c = idx - 12 * r;
        double T = (double)(r + 1) * dt, T2 = 0.5 * T * T;
        double val;
        if (c < 3) {
          double ra = (c == 0) ? R[0] : (c == 1 ? R[1] : R[2]);
          double rb = (c == 0) ? R[3] : (c == 1 ? R[4] : R[5]);
          double rcc = (c == 0) ? R[6] : (c == 1 ? R[7] : R[8]);
          double rto = ra * om0 + rb * om1 + rcc * om2;
          double rtf = ra * ft0 + rb * ft1 + rcc * ft2;
          double th0 = (c == 0) ? roll : (c == 1 ? pitch : yaw);
          val = th0 + T * rto + T2 * rtf;
        } else if (c < 6) {
          int a = c - 3;
          val = (double)rec[CMPC_REC_P + a] + T * (double)rec[CMPC_REC_V + a] + T2 * (double)rec[CMPC_REC_FDIST + 3 + a];
          if (a == 2) val += T2 * az + (T * T * T / 6.0) * xd * ffx;
        } else if (c < 9) {
          int a = c - 6;
          val = (double)rec[CMPC_REC_W + a] + T * (double)rec[CMPC_REC_FDIST + a];
        } else {
          int a = c - 9;
          val = (double)rec[CMPC_REC_V + a] + T * (double)rec[CMPC_REC_FDIST + 3 + a];
          if (a == 2) val += T * az + T2 * xd * ffx;
        }
        ev[idx] = (double)rec[CMPC_REC_WEIGHTS + c] * (val - (double)rec[CMPC_REC_TRAJ + idx]);
      }
    }
    __syncthreads();
    const int nc = redi[0];
    const int n = 3 * nc, m = 5 * nc;
    // reduced variable -> (step, foot, component); -1 pads the register tiles
    for (int i = tid; i < max(S::NPAD, n); i += NT) {
      int info = -1;
      if (i < n) {
        int j = i / 3, comp = i - 3 * j, k = fs[j];
        info = (k >> 2) | ((k & 3) << 8) | (comp << 16);
      }
      if (i < max(S::NPAD, P.nmax)) rowinfo[i] = info;
    }
    if (tid < 6) sScal[1 + tid] = (double)rec[CMPC_REC_WEIGHTS + (tid < 3 ? 3 + tid : 6 + tid)];  // wp, wv
    // foot-pair blocks  PT = RW_i^T S_theta RW_j,  PO = W_i^T S_omega W_j
    for (int e = tid; e < 288; e += NT) {
      int which = e / 144, ee = e - 144 * which;
      int fi = ee / 36, fj = (ee / 9) & 3, a = (ee % 9) / 3, b = ee % 3;
      const double* Mi = (which == 0 ? sRW : sW) + fi * 9;
      const double* Mj = (which == 0 ? sRW : sW) + fj * 9;
      int wo = which == 0 ? 0 : 6;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 3; k++) acc += Mi[k * 3 + a] * (double)rec[CMPC_REC_WEIGHTS + wo + k] * Mj[k * 3 + b];
      (which == 0 ? sPT : sPO)[ee] = acc;
    }
    // horizon aggregates of e:  agg[c][0:3]=sum c2 e_theta, [3:6]=sum c1 e_omega,
    // [6:9]=(sum c2 e_p + c1 e_v)/m, [9]=xd/m (sum c3 e_pz + c2 e_vz)
    {
      const double xd = rec[CMPC_REC_XDRAG];
      for (int idx = tid; idx < 10 * h; idx += NT) {
        int c = idx / 10, comp = idx - 10 * c;
        double acc = 0.0;
        for (int r = c; r < h; r++) {
          double tau = (double)(r - c) * dt;
          double c1 = dt, c2 = tau * dt + 0.5 * dt * dt, c3 = 0.5 * tau * tau * dt + 0.5 * tau * dt * dt + dt * dt * dt / 6.0;
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

    int status = CMPC_ST_SOLVED;
    int iters = 0;
    if (nc == 0) {
      status = CMPC_ST_EMPTY;
    } else if (n > P.nmax || (S::REG && n > S::NPAD)) {
      status = CMPC_ST_CAPACITY;  // launch was sized for fewer contact foot-steps than this instance has
    } else {
      // ---- C. gradient and Hessian of the reduced QP ----
      for (int I = tid; I < n; I += NT) {
        int j = I / 3, comp = I - 3 * j;
        int k = fs[j], step = k >> 2, f = k & 3;
        const double* a = agg + 10 * step;
        double acc = sRW[f * 9 + 0 + comp] * a[0] + sRW[f * 9 + 3 + comp] * a[1] + sRW[f * 9 + 6 + comp] * a[2] +
                     sW[f * 9 + 0 + comp] * a[3] + sW[f * 9 + 3 + comp] * a[4] + sW[f * 9 + 6 + comp] * a[5] + a[6 + comp];
        if (comp == 0) acc += a[9];
        g[I] = 2.0 * acc;
      }
      HessCtx C;
      C.sig = P.sigma; C.sPT = sPT; C.sPO = sPO; C.h = h;
      C.wp = sScal + 1;
      C.fs = fs;
      C.xd = rec[CMPC_REC_XDRAG];
      C.m2 = minv * minv;
      C.alpha2 = 2.0 * (double)rec[CMPC_REC_ALPHA];
      pc.tick(CMPC_PH_PREP);
      // ---- D. K <- H^-1 (H is SPD: 2aI + 2B'SB, a > 0) ----
      if (S::REG) build_invert_regtile2<S>(C, rowinfo, n, tid, cbuf, K, red, pc);
      else build_invert_smem<NT>(C, rowinfo, n, tid, kn, K);
      __syncthreads();
      // x = -H^-1 g, slacks of every candidate row
      for (int i = tid; i < n; i += NT) {
        double acc = 0.0;
        for (int j = 0; j < n; j++) acc = fma(K[j * n + i], g[j], acc);
        x[i] = -acc;
      }
      __syncthreads();
      for (int c = tid; c < m; c += NT) {
        int ia, iz; double va, vz;
        cons_of(c, mu_inv, ia, va, iz, vz);
        double b = 0.0;
        if (c % 5 == 4) b = -(double)gait[fs[c / 5]] * P.f_max;
        s[c] = va * x[ia] + vz * x[iz] - b;
        rc[c] = 0.0;
        isact[c] = 0;
      }
      __syncthreads();
      flops_acc += 2.0 * (double)n * n * n * 0.5 + 12.0 * (double)n * n + 2.0 * (double)n * n;
      pc.tick(CMPC_PH_STORE);

      // ---- E. Goldfarb-Idnani dual active set on K (cmpc_qp.cuh), run by the first GT threads ----
#ifdef CMPC_QP_FULL
      constexpr int GT = NT;
#else
      constexpr int GT = (S::NPAD < NT) ? S::NPAD : NT;
#endif
      if (tid < GT) {
        QpState qs = qp_dual_active_set<GT>(tid, n, m, mu_inv, P.tol_violation, P.max_iter, P.qcap, K, x, s, rc, isact,
                                            act, u, dvec, rvec, col, Pp, kn, z, vv, red, flops_acc);
        if (tid == 0) { redi[1] = qs.status; redi[2] = qs.iters; redi[3] = qs.q; }
      }
      __syncthreads();
      pc.tick(CMPC_PH_QP);
      status = redi[1];
      iters = redi[2];
      const int q = redi[3];
      // ---- objective 0.5 x'Hx + g'x = 0.5 g'x + 0.5 lambda'b at a KKT point ----
      double part = 0.0;
      for (int i = tid; i < n; i += NT) part = fma(0.5 * g[i], x[i], part);
      for (int k = tid; k < q; k += NT) {
        int c = act[k];
        if (c % 5 == 4) part -= 0.5 * u[k] * (double)gait[fs[c / 5]] * P.f_max;
      }
      part = block_sum<NT>(part, red, tid);
      if (tid == 0) sScal[0] = part;
      if (status == CMPC_ST_WSOVERFLOW && tid == 0 && P.overflow_list) {
        int pos = atomicAdd(P.overflow_count, 1);
        P.overflow_list[pos] = inst;
      }
    }
    __syncthreads();
    // ---- F. outputs (an overflowed instance is left for the full-capacity launch) ----
    if (status != CMPC_ST_WSOVERFLOW || !P.overflow_list) {
      bool have_x = (nc > 0 && status != CMPC_ST_CAPACITY);
      {  // a non-finite iterate (NaN / Inf upstream) reports CMPC_ST_NONFINITE and zero forces
        int fin = 1;
        if (have_x)
          for (int i = tid; i < 3 * nc; i += NT) fin = fin && isfinite(x[i]);
        fin = __syncthreads_and(fin);
        if (have_x && !fin) { status = CMPC_ST_NONFINITE; have_x = false; }
      }
      if (P.forces) {
        double* out = P.forces + (size_t)inst * 12 * h;
        for (int idx = tid; idx < 12 * h; idx += NT) {
          int k = idx / 3, comp = idx - 3 * k;
          int j = fsinv[k];
          out[idx] = (j >= 0 && have_x) ? x[3 * j + comp] : 0.0;
        }
      }
      if (P.active) {
        signed char* out = P.active + (size_t)inst * 20 * h;
        for (int idx = tid; idx < 20 * h; idx += NT) {
          int k = idx / 5, t = idx - 5 * k;
          int j = fsinv[k];
          signed char a = 0;
          if (j >= 0 && have_x) {
            double fx = x[3 * j], fy = x[3 * j + 1], fz = x[3 * j + 2];
            double row = (t == 0) ? fx * mu_inv + fz : (t == 1) ? -fx * mu_inv + fz : (t == 2) ? fy * mu_inv + fz
                       : (t == 3) ? -fy * mu_inv + fz : fz;
            if (row <= P.tol_active) a = -1;
            if (t == 4 && row >= (double)gait[k] * P.f_max - P.tol_active) a = 1;
          }
          out[idx] = a;
        }
      }
      if (tid == 0) {
        if (P.objective) P.objective[inst] = have_x ? sScal[0] : 0.0;
        if (P.status) P.status[inst] = status;
        if (P.iterations) P.iterations[inst] = iters;
      }
    }
    __syncthreads();  // record buffer and work arrays are reused by the next instance
    pc.tick(CMPC_PH_OUT);
  }
#ifdef CMPC_CANARY
  __syncthreads();
  canary_check(smem, cv.guard, cv.nguard, tid, NT, "cmpc_solve_kernel");
#endif
  if (tid == 0 && P.flops && flops_acc > 0.0) atomicAdd(P.flops + CMPC_K_FUSED, (unsigned long long)flops_acc);
}

// ---------------------------------------------------------------------------
// host-side launch plumbing
// ---------------------------------------------------------------------------
namespace {
template <class S, bool ADAPT>
int launch_t(const CmpcParams& P, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e =
      cudaFuncSetAttribute(cmpc_solve_kernel<S, ADAPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  cmpc_solve_kernel<S, ADAPT><<<grid, S::NT, smem, st>>>(P);
  return (int)cudaGetLastError();
}
template <class S, bool ADAPT>
int occ_t(size_t smem) {
  int nb = 0;
  if (cudaFuncSetAttribute(cmpc_solve_kernel<S, ADAPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess)
    return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cmpc_solve_kernel<S, ADAPT>, S::NT, smem) != cudaSuccess)
    return -1;
  return nb;
}
int npad_of(int shape) {
  switch (shape) {
    case CMPC_SHAPE_64: return Shape64::NPAD;
    case CMPC_SHAPE_64W: return Shape64w::NPAD;
    case CMPC_SHAPE_128: return Shape128::NPAD;
    case CMPC_SHAPE_GMEM: return ShapeGmem::NPAD;
    default: return ShapeMem::NPAD;
  }
}
}  // namespace

size_t cmpc_smem_bytes(int horizon, int nmax, int qcap, int shape, bool adapt) {
  return (size_t)make_carve(horizon, nmax, qcap, cmpc_rec_stride(horizon), npad_of(shape), adapt,
                            shape == CMPC_SHAPE_GMEM).total;
}

int cmpc_shape_threads(int shape) {
  switch (shape) {
    case CMPC_SHAPE_64: return Shape64::NT;
    case CMPC_SHAPE_64W: return Shape64w::NT;
    case CMPC_SHAPE_128: return Shape128::NT;
    default: return ShapeMem::NT;
  }
}

#define CMPC_DISPATCH(FN, ...)                                                                   \
  switch (shape) {                                                                               \
    case CMPC_SHAPE_64: return adapt ? FN<Shape64, true>(__VA_ARGS__) : FN<Shape64, false>(__VA_ARGS__);     \
    case CMPC_SHAPE_64W: return adapt ? FN<Shape64w, true>(__VA_ARGS__) : FN<Shape64w, false>(__VA_ARGS__);  \
    case CMPC_SHAPE_128: return adapt ? FN<Shape128, true>(__VA_ARGS__) : FN<Shape128, false>(__VA_ARGS__);  \
    case CMPC_SHAPE_GMEM: return adapt ? FN<ShapeGmem, true>(__VA_ARGS__) : FN<ShapeGmem, false>(__VA_ARGS__); \
    default: return adapt ? FN<ShapeMem, true>(__VA_ARGS__) : FN<ShapeMem, false>(__VA_ARGS__);              \
  }

int cmpc_launch_solve(const CmpcParams& P, int shape, int grid, void* stream) {
  const bool adapt = P.adapt_mode >= 0;
  size_t smem = cmpc_smem_bytes(P.horizon, P.nmax, P.qcap, shape, adapt);
  cudaStream_t st = (cudaStream_t)stream;
  CMPC_DISPATCH(launch_t, P, grid, smem, st)
}

int cmpc_max_ctas_per_sm(int shape, size_t smem, bool adapt) {
  CMPC_DISPATCH(occ_t, smem)
}

// ---------------------------------------------------------------------------
// FP64 roofline denominator: independent DFMA chains on every SM
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cmpc_dfma_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) out[0] = s;
}

// the same for the FP64 tensor pipe: sixteen independent DMMA m8n8k4 accumulators per warp (512 flops each)
__global__ void __launch_bounds__(256) cmpc_dmma_peak_kernel(double* out, int iters, double seed) {
  double c[16][2];
#pragma unroll
  for (int k = 0; k < 16; k++) { c[k][0] = seed + k; c[k][1] = seed - k; }
  const double a = 0.999999 + 1e-9 * threadIdx.x, b = 1.0 / 4.0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; k++) s += c[k][0] + c[k][1];
  if (s == 123.456) out[0] = s;
}

int cmpc_run_dmma_peak(int sm_count, void* stream, double* out_dev, int iters) {
  cmpc_dmma_peak_kernel<<<sm_count * 4, 256, 0, (cudaStream_t)stream>>>(out_dev, iters, 1.0);
  return (int)cudaGetLastError();
}

int cmpc_run_dfma_peak(int sm_count, void* stream, double* out_dev, int iters) {
  cmpc_dfma_peak_kernel<<<sm_count * 8, 256, 0, (cudaStream_t)stream>>>(out_dev, iters, 1.0);
  return (int)cudaGetLastError();
}
