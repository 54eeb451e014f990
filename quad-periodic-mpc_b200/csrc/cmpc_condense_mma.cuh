// Kernel 1 of the two-kernel pipeline for reduced problems of up to 64 variables (every A1 gait at
// horizon <= 10 except all-feet-down stances): condensation + K = H^-1, ONE WARP PER INSTANCE, the
// rank-8 updates of the blocked symmetric sweep on the FP64 tensor cores (DMMA m8n8k4).
//
// Why a warp and why DMMA (profiles/r1_*): with DFMA register tiles (cmpc_condense.cuh) the sweep is bound by
// issue slots and shared-memory operand traffic (shared pipe 63 %, issue 45 %, FP64 pipe 39 %) and by CTA
// barriers around the serial 8 x 8 pivot-block inversion.  DMMA runs at the same FP64 rate on B200
// (37 TFLOP/s measured, scripts/ubench) with 1/8 of the instructions and one operand fetch per 8 x 8 tile,
// and a warp-private instance needs no block barrier at all: independent warps hide each other's latency.
//
// The padded 64 x 64 symmetric matrix is held as its 36 lower-triangular 8 x 8 tiles in DMMA accumulator
// layout (lane l: row l>>2, columns 2(l&3), 2(l&3)+1; 72 doubles per lane).  Block step s:
//   1. publish the pivot rows (tiles (s, J<=s) as they are, tiles (I>s, s) transposed: the matrix is
//      symmetric) as an 8 x 64 panel C in warp-private shared memory, with D - I in the diagonal block;
//   2. D^-1 by Gauss-Jordan over warp shuffles (8 x 8, serial chain of 8 pivots);
//   3. M = -D^-1 C: 16 DMMAs;
//   4. every tile (I, J) += C_I' M_J: 72 DMMAs, operands fetched once per tile row / column.
// Publishing D - I makes the same update produce the swept pivot rows and columns; every swept diagonal
// entry carries a constant +2 removed at the end; H is scaled by an exact power of two (see cmpc_sweep.cuh).
#pragma once

namespace {

constexpr int MMA_NPAD = 64;
constexpr int MMA_PS = 68;      // panel row stride (doubles): = 4 (mod 16), conflict-free fragment loads
constexpr int MMA_WPC = 4;      // independent warps (instances in flight) per CTA
constexpr int MMA_TILES = 36;
__host__ __device__ constexpr int tix(int I, int J) { return I * (I + 1) / 2 + J; }

constexpr int MMA_PQ = 292;    // 144 foot-pair/component-pair coefficients + a zero region addressed by padding rows / columns

struct MCarve {
  int sig, ss, warp0;                                                  // CTA-level
  int rec0, rec1, bars, fs, g, pq, un, warp_total;                     // per warp, relative to the warp base
  int small, evec, agg, rowinfo, xtab;                                 // pre-sweep scratch, inside the union
  int pan, mm, dv;                                                     // sweep buffers, inside the union
  int total;
};

__host__ __device__ inline MCarve make_mcarve(int h, int rec_stride, bool adapt) {
  MCarve c;
  int o = 0;
  c.sig = o; o += align16(8 * CMPC_SIG_COUNT * h * h);
  c.ss = o; o += 16 * h * h;  // (s22, s11) interleaved
  c.warp0 = o;
  int w = 0;
  c.rec0 = w; w += align16(rec_stride);
  c.rec1 = w; w += align16(rec_stride);
  c.bars = w; w += 16;
  c.fs = w; w += align16(CMPC_MAX_FS);
  c.g = w; w += 8 * MMA_NPAD;
  c.pq = w; w += 16 * MMA_PQ;
  c.un = w;
  // pre-sweep scratch ...
  int u = w;
  c.small = u; u += align16(8 * (36 + 36 + 144 + 144 + 16));
  c.evec = u; u += align16(8 * 12 * h);
  c.agg = u; u += align16(8 * 10 * h);
  c.rowinfo = u; u += 4 * MMA_NPAD;
  c.xtab = u; u += 24 * h * h;  // x_drag couplings: XA[ab], X0[ab], XA[ba]
  // ... shares its memory with the sweep buffers (pan, mm, dv contiguous: the x0 reduction, 32 x 25 doubles,
  // and the estimator, 3 x 400 doubles, borrow them)
  int v = w;
  c.pan = v; v += 8 * 8 * MMA_PS;
  c.mm = v; v += 8 * 8 * MMA_PS;
  c.dv = v; v += 8 * 64;
  if (adapt && v - c.pan < 8 * 3 * CMPC_ADAPT_WINDOW) v = c.pan + 8 * 3 * CMPC_ADAPT_WINDOW;
  w = u > v ? u : v;
  c.warp_total = align16(w);
  c.total = c.warp0 + MMA_WPC * c.warp_total;
  return c;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
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

// 8 x 8 inverse by one warp: lane (r = lane & 7, rep = lane >> 3) holds D[r][2 rep], D[r][2 rep + 1]
__device__ __forceinline__ void warp_inv8_mma(double& a0, double& a1, int lane) {
  const int r = lane & 7, rep = lane >> 3;
#pragma unroll
  for (int p = 0; p < 8; p++) {
    const int prep = p >> 1;
    const double mine = (p & 1) ? a1 : a0;
    const double arp = __shfl_sync(0xffffffffu, mine, r | (prep << 3));  // D[r][p]
    const double dpp = __shfl_sync(0xffffffffu, mine, p | (prep << 3));  // D[p][p]
    const double ap0 = __shfl_sync(0xffffffffu, a0, p | (rep << 3));     // D[p][c0]
    const double ap1 = __shfl_sync(0xffffffffu, a1, p | (rep << 3));     // D[p][c0 + 1]
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

}  // namespace

template <bool ADAPT, int MINB>
__global__ void __launch_bounds__(32 * MMA_WPC, MINB) cmpc_condense_mma_kernel(const __grid_constant__ CmpcParams P) {
  constexpr int PS = MMA_PS;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = P.horizon, hh = h * h;
  const MCarve cv = make_mcarve(h, P.rec_stride, ADAPT);
  double* sig = reinterpret_cast<double*>(smem + cv.sig);
  double2* SS = reinterpret_cast<double2*>(smem + cv.ss);
  unsigned char* wb = smem + cv.warp0 + (size_t)warp * cv.warp_total;
  unsigned char* recbuf[2] = {wb + cv.rec0, wb + cv.rec1};
  uint64_t* bars = reinterpret_cast<uint64_t*>(wb + cv.bars);
  double* sW = reinterpret_cast<double*>(wb + cv.small);
  double* sRW = sW + 36;
  double* sPT = sRW + 36;
  double* sPO = sPT + 144;
  double* sScal = sPO + 144;  // [1..6] position / velocity weights, [8..10] roll pitch yaw, [12..15] estimator
  double* ev = reinterpret_cast<double*>(wb + cv.evec);
  double* agg = reinterpret_cast<double*>(wb + cv.agg);
  unsigned char* fs = wb + cv.fs;
  int* rowinfo = reinterpret_cast<int*>(wb + cv.rowinfo);
  double* g = reinterpret_cast<double*>(wb + cv.g);
  double* pan = reinterpret_cast<double*>(wb + cv.pan);
  double* mm = reinterpret_cast<double*>(wb + cv.mm);
  double* dv = reinterpret_cast<double*>(wb + cv.dv);
  double2* PQ = reinterpret_cast<double2*>(wb + cv.pq);
  double* xtab = reinterpret_cast<double*>(wb + cv.xtab);

  for (int i = threadIdx.x; i < CMPC_SIG_COUNT * hh; i += blockDim.x) sig[i] = __ldg(P.sigma + i);
  for (int i = threadIdx.x; i < hh; i += blockDim.x)
    SS[i] = make_double2(__ldg(P.sigma + CMPC_SIG_22 * hh + i), __ldg(P.sigma + CMPC_SIG_11 * hh + i));
  for (int i = 144 + lane; i < MMA_PQ; i += 32) PQ[i] = make_double2(0.0, 0.0);
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const int count = P.count;
  int cur = 0;
  if (lane == 0) {
    cur = atomicAdd(P.sched, 1);
    if (cur < count) {
      mbar_expect_tx(&bars[0], (uint32_t)P.rec_stride);
      bulk_g2s(recbuf[0], P.records + (size_t)cur * P.rec_stride, (uint32_t)P.rec_stride, &bars[0]);
    }
  }
  cur = __shfl_sync(0xffffffffu, cur, 0);

  uint32_t phase[2] = {0u, 0u};
  const double dt = P.dt, minv = P.mass_inv;
  double flops_acc = 0.0;
  const int r = lane >> 2, q = lane & 3;

  for (int buf = 0; cur < count; buf ^= 1) {
    const int inst = cur;
    int nxt = 0;
    __syncwarp();
    if (lane == 0) {  // draw and prefetch the next instance
      nxt = atomicAdd(P.sched, 1);
      if (nxt < count) {
        fence_proxy_async();
        mbar_expect_tx(&bars[buf ^ 1], (uint32_t)P.rec_stride);
        bulk_g2s(recbuf[buf ^ 1], P.records + (size_t)nxt * P.rec_stride, (uint32_t)P.rec_stride, &bars[buf ^ 1]);
      }
    }
    nxt = __shfl_sync(0xffffffffu, nxt, 0);
    mbar_wait(&bars[buf], phase[buf]);
    phase[buf] ^= 1u;
    const float* rec = reinterpret_cast<const float*>(recbuf[buf]);
    const unsigned char* gait = recbuf[buf] + 4 * (CMPC_REC_TRAJ + 12 * h);
    double* slot = P.qws + (size_t)inst * P.qws_stride;
    int* hdr = reinterpret_cast<int*>(slot + (size_t)P.nmax * P.nmax + 2 * P.nmax);

    // ---- 0. periodic-disturbance estimator (Adaptive MPC), SolverMPC.cpp:688-798 ----
    if (ADAPT) {
      double* est_s = sScal + 12;
      if (P.adapt_mode == 0 || P.adapt_mode == 1) {
        estimate_disturbance_warp(P, inst, lane, pan, est_s);
      } else {
        if (lane < 4) est_s[lane] = P.est[(size_t)inst * 4 + lane];
        __syncwarp();
      }
      if (lane == 0) {
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
      __syncwarp();
    }

    // ---- A. contact foot-steps, Euler angles (lanes 0..2), W_f = I^-1 [r_f]x and R^T W_f ----
    int nc = 0;
    for (int base = 0; base < 4 * h; base += 32) {
      const int k = base + lane;
      bool keep = false;
      if (k < 4 * h) {
        const double ub = (double)gait[k] * P.f_max;
        keep = !(ub < 0.01 && ub > -0.01);  // the reference drops a foot-step whose fz bound is ~0
      }
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      if (keep) fs[nc + __popc(mask & ((1u << lane) - 1u))] = (unsigned char)k;
      nc += __popc(mask);
    }
    const int n = 3 * nc;
    if (lane < 3) {
      // quat_to_rpy, SolverMPC.cpp:352-361
      const double qw = rec[CMPC_REC_Q + 0], qx = rec[CMPC_REC_Q + 1], qy = rec[CMPC_REC_Q + 2], qz = rec[CMPC_REC_Q + 3];
      double val;
      if (lane == 0) val = atan2(2.0 * (qy * qz + qw * qx), qw * qw - qx * qx - qy * qy + qz * qz);
      else if (lane == 1) val = asin(fmin(-2.0 * (qx * qz - qw * qy), 0.99999));
      else val = atan2(2.0 * (qx * qy + qw * qz), qw * qw + qx * qx - qy * qy - qz * qz);
      sScal[8 + lane] = val;
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
    {
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
      for (int e = lane; e < 72; e += 32) {
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
    }
    if (lane >= 24 && lane < 30) {
      const int t = lane - 24;
      sScal[1 + t] = (double)rec[CMPC_REC_WEIGHTS + (t < 3 ? 3 + t : 6 + t)];  // position, velocity weights
    }
    __syncwarp();

    int status = CMPC_ST_SOLVED;
    if (nc == 0) status = CMPC_ST_EMPTY;
    else if (n > P.nmax || n > MMA_NPAD) status = CMPC_ST_CAPACITY;
    if (status == CMPC_ST_SOLVED) {
      // reduced variable -> (step, foot, component); -1 pads the tiles
      for (int i = lane; i < MMA_NPAD; i += 32) {
        int info = -1;
        if (i < n) {
          const int j = i / 3, comp = i - 3 * j, k = fs[j];
          info = (k >> 2) | ((k & 3) << 8) | (comp << 16);
        }
        rowinfo[i] = info;
      }
      // ---- B. weighted tracking error of the free response; foot-pair blocks PT, PO ----
      {
        const double xd = rec[CMPC_REC_XDRAG];
        const double om0 = rec[CMPC_REC_W + 0], om1 = rec[CMPC_REC_W + 1], om2 = rec[CMPC_REC_W + 2];
        const double ft0 = rec[CMPC_REC_FDIST + 0], ft1 = rec[CMPC_REC_FDIST + 1], ft2 = rec[CMPC_REC_FDIST + 2];
        const double ffx = rec[CMPC_REC_FDIST + 3];
        const double az = xd * (double)rec[CMPC_REC_V + 0] + P.gravity;  // row 11 of A x0
        for (int idx = lane; idx < 12 * h; idx += 32) {
          const int rr = idx / 12, c = idx - 12 * rr;
          const double T = (double)(rr + 1) * dt, T2 = 0.5 * T * T;
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
      for (int e = lane; e < 288; e += 32) {
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
      __syncwarp();
      // horizon aggregates of e:  agg[c][0:3]=sum c2 e_theta, [3:6]=sum c1 e_omega,
      // [6:9]=(sum c2 e_p + c1 e_v)/m, [9]=xd/m (sum c3 e_pz + c2 e_vz)
      {
        const double xd = rec[CMPC_REC_XDRAG];
        for (int idx = lane; idx < 10 * h; idx += 32) {
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
      __syncwarp();
      // ---- C. gradient ----
      for (int I = lane; I < MMA_NPAD; I += 32) {
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
      // ---- D. H straight into the accumulator tiles, scaled by a power of two ----
      HessS C;
      C.sig = sig; C.sPT = sPT; C.sPO = sPO; C.wp = sScal + 1; C.h = h; C.hh = hh;
      C.xd = rec[CMPC_REC_XDRAG];
      C.m2 = minv * minv;
      C.alpha2 = 2.0 * (double)rec[CMPC_REC_ALPHA];
      double dmax = 0.0;
      for (int i = lane; i < n; i += 32) {
        const int ri = rowinfo[i];
        dmax = fmax(dmax, hess_entry_s(C, ri, ri, true));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
      int e2;
      frexp(dmax, &e2);
      const double scale = ldexp(1.0, -e2);  // exact; scaled diagonal < 1
      const int nblk = (n + 7) >> 3;
      // Scaled coefficient tables: H_ij scale = s22[ab] PQ[p].x + s11[ab] PQ[p].y (+ x_drag couplings), with
      // p = (fi, c1, fj, c2); the block-diagonal (c1 == c2) position / velocity terms are folded into PQ.
      const double sc2 = 2.0 * scale;
      for (int e = lane; e < 144; e += 32) {
        const int c1 = (e % 9) / 3, c2 = e % 3;
        double pt = sPT[e], po = sPO[e];
        if (c1 == c2) { pt += C.m2 * C.wp[c1]; po += C.m2 * C.wp[3 + c1]; }
        PQ[e] = make_double2(sc2 * pt, sc2 * po);
      }
      const bool drag = (C.xd != 0.0);
      if (drag) {
        for (int ab = lane; ab < hh; ab += 32) {
          xtab[ab] = sc2 * C.m2 * C.xd * (C.wp[2] * sig[CMPC_SIG_23 * hh + ab] + C.wp[5] * sig[CMPC_SIG_12 * hh + ab]);
          xtab[hh + ab] = sc2 * C.m2 * C.xd * C.xd * (C.wp[2] * sig[CMPC_SIG_33 * hh + ab] + C.wp[5] * sig[CMPC_SIG_22 * hh + ab]);
          const int ba = (ab % h) * h + ab / h;
          xtab[2 * hh + ab] = sc2 * C.m2 * C.xd * (C.wp[2] * sig[CMPC_SIG_23 * hh + ba] + C.wp[5] * sig[CMPC_SIG_12 * hh + ba]);
        }
      }
      __syncwarp();
      double t[MMA_TILES][2];
      {
        // row descriptors: table offsets (padding rows and columns address the zero region of PQ)
        int rah[8], rpo[8];
        unsigned rfl = 0;  // bit I: row component is x, bit 8 + I: row component is z (x_drag couplings)
#pragma unroll
        for (int I = 0; I < 8; I++) {
          const int info = rowinfo[8 * I + r];
          const bool ok = info >= 0;
          const int c1 = info >> 16;
          rah[I] = ok ? (info & 0xff) * h : 0;
          rpo[I] = ok ? ((info >> 8) & 3) * 36 + c1 * 3 : 144;
          if (ok && c1 == 0) rfl |= 1u << I;
          if (ok && c1 == 2) rfl |= 256u << I;
        }
        const double alpha2s = C.alpha2 * scale;
#pragma unroll
        for (int J = 0; J < 8; J++) {
          const int2 info2 = *reinterpret_cast<const int2*>(rowinfo + 8 * J + 2 * q);
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int info = e ? info2.y : info2.x;
            const bool ok = info >= 0;
            const int c2 = info >> 16;
            const int cb = ok ? (info & 0xff) : 0;
            const int cpo = ok ? ((info >> 8) & 3) * 9 + c2 : 144;
            const bool k0 = ok && c2 == 0, k2 = ok && c2 == 2;
#pragma unroll
            for (int I = J; I < 8; I++) {
              const double2 ss = SS[rah[I] + cb];
              const double2 pq = PQ[rpo[I] + cpo];
              double v = fma(ss.x, pq.x, ss.y * pq.y);
              if (drag) {
                // (c1, c2) = (x, x): X0[ab]; (z, x): XA[ab]; (x, z): XA[ba]
                const bool r0 = (rfl >> I) & 1u, r2 = (rfl >> (8 + I)) & 1u;
                const int ab = rah[I] + cb;
                if (r0 && k0) v += xtab[hh + ab];
                if (r2 && k0) v += xtab[ab];
                if (r0 && k2) v += xtab[2 * hh + ab];
              }
              if (I == J && r == 2 * q + e) v += (8 * I + r < n) ? alpha2s : 0.5;  // alpha; 1/2 on the padding
              t[tix(I, J)][e] = v;
            }
          }
        }
      }
      __syncwarp();  // the sweep buffers overlay the tables read above
      // ---- blocked symmetric sweep ----
      const int fo = q * PS + r;  // fragment offset: element (k = q, row/col = r)
#pragma unroll 1
      for (int s = 0; s < nblk; s++) {
        // 1. publish the panel (rows 8s..8s+7 of the symmetric matrix), D - I in the diagonal block
#pragma unroll
        for (int I = 0; I < 8; I++)
#pragma unroll
          for (int J = 0; J <= I; J++) {
            if (I == s) {
              double v0 = t[tix(I, J)][0], v1 = t[tix(I, J)][1];
              if (J == I) {
                if (r == 2 * q) v0 -= 1.0;
                if (r == 2 * q + 1) v1 -= 1.0;
              }
              *reinterpret_cast<double2*>(pan + r * PS + 8 * J + 2 * q) = make_double2(v0, v1);
            } else if (J == s) {
              pan[(2 * q) * PS + 8 * I + r] = t[tix(I, J)][0];
              pan[(2 * q + 1) * PS + 8 * I + r] = t[tix(I, J)][1];
            }
          }
        __syncwarp();
        // 2. -D^-1
        {
          const int r8 = lane & 7, c0 = 2 * (lane >> 3);
          double2 d2 = *reinterpret_cast<const double2*>(pan + r8 * PS + 8 * s + c0);
          if (c0 == r8) d2.x += 1.0;
          if (c0 + 1 == r8) d2.y += 1.0;
          warp_inv8_mma(d2.x, d2.y, lane);
          *reinterpret_cast<double2*>(dv + r8 * 8 + c0) = make_double2(-d2.x, -d2.y);
        }
        __syncwarp();
        // 3. M = -D^-1 C
        {
          const double a0 = dv[r * 8 + q], a1 = dv[r * 8 + 4 + q];
#pragma unroll
          for (int J = 0; J < 8; J++) {
            double m0 = 0.0, m1 = 0.0;
            dmma884(m0, m1, a0, pan[fo + 8 * J]);
            dmma884(m0, m1, a1, pan[fo + 4 * PS + 8 * J]);
            *reinterpret_cast<double2*>(mm + r * PS + 8 * J + 2 * q) = make_double2(m0, m1);
          }
        }
        __syncwarp();
        // 4. every tile (I, J) += C_I' M_J
        {
          double mf[8][2];
#pragma unroll
          for (int J = 0; J < 8; J++) {
            mf[J][0] = mm[fo + 8 * J];
            mf[J][1] = mm[fo + 4 * PS + 8 * J];
          }
#pragma unroll
          for (int I = 0; I < 8; I++) {
            if (I < nblk) {
              const double p0 = pan[fo + 8 * I], p1 = pan[fo + 4 * PS + 8 * I];
#pragma unroll
              for (int J = 0; J <= I; J++) {
                dmma884(t[tix(I, J)][0], t[tix(I, J)][1], p0, mf[J][0]);
                dmma884(t[tix(I, J)][0], t[tix(I, J)][1], p1, mf[J][1]);
              }
            }
          }
        }
        __syncwarp();  // the next publish overwrites pan
      }
      // ---- K_ij = -(A_ij - 2 d_ij) scale to the workspace slot; x0 = -K g ----
      {
        double rowacc[8], colacc[8][2];
#pragma unroll
        for (int I = 0; I < 8; I++) rowacc[I] = 0.0;
#pragma unroll
        for (int J = 0; J < 8; J++) colacc[J][0] = colacc[J][1] = 0.0;
        double gr[8], gc[8][2];
#pragma unroll
        for (int I = 0; I < 8; I++) {
          gr[I] = g[8 * I + r];
          gc[I][0] = g[8 * I + 2 * q];
          gc[I][1] = g[8 * I + 2 * q + 1];
        }
#pragma unroll
        for (int I = 0; I < 8; I++)
#pragma unroll
          for (int J = 0; J <= I; J++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int i = 8 * I + r, j = 8 * J + 2 * q + e;
              const double a = t[tix(I, J)][e];
              rowacc[I] = fma(a, gc[J][e], rowacc[I]);
              if (J < I) colacc[J][e] = fma(a, gr[I], colacc[J][e]);
              if (i < n && j < n) {
                const double kv = -(a - (i == j ? 2.0 : 0.0)) * scale;
                slot[(size_t)i * n + j] = kv;
                if (J < I) slot[(size_t)j * n + i] = kv;
              }
            }
        double* part = pan;  // [32][25]
#pragma unroll
        for (int I = 0; I < 8; I++) part[lane * 25 + I] = rowacc[I];
#pragma unroll
        for (int J = 0; J < 8; J++) {
          part[lane * 25 + 8 + 2 * J] = colacc[J][0];
          part[lane * 25 + 8 + 2 * J + 1] = colacc[J][1];
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
          const int I = i >> 3, ri = i & 7, qi = (i & 7) >> 1, ei = i & 1;
          double acc = 0.0;
#pragma unroll
          for (int qq = 0; qq < 4; qq++) acc += part[(ri * 4 + qq) * 25 + I];
#pragma unroll
          for (int rr = 0; rr < 8; rr++) acc += part[(rr * 4 + qi) * 25 + 8 + 2 * I + ei];
          const double gi = g[i];
          slot[(size_t)P.nmax * P.nmax + i] = gi;
          slot[(size_t)P.nmax * P.nmax + P.nmax + i] = scale * (acc - 2.0 * gi);
        }
      }
      flops_acc += 2.0 * (double)n * n * n * 0.5 + 12.0 * (double)n * n + 2.0 * (double)n * n;
    }
    // contact list for kernel 2
    if (lane == 0) { hdr[0] = nc; hdr[1] = status; }
    {
      unsigned char* hb = reinterpret_cast<unsigned char*>(hdr + 2);
      for (int j = lane; j < nc; j += 32) {
        const int k = fs[j];
        hb[j] = (unsigned char)k;
        hb[CMPC_MAX_FS + j] = gait[k];
      }
    }
    cur = nxt;
  }
  if (lane == 0 && P.flops && flops_acc > 0.0) atomicAdd(P.flops, (unsigned long long)flops_acc);
}
