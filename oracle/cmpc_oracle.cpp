// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the reference's dense convex-MPC solve
// (be2r_cmpc_unitree/src/controllers/convexMPC/SolverMPC.cpp, solve_mpc()
// and helpers) used as the checker for the B200 kernels.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load the library built from this file.
//
// What is restated here and what is the real reference:
//   * The condensation (state build, continuous matrices, matrix exponential,
//     A_qp/B_qp/Q_qp, H and g, bounds, the swing-foot elimination) is RESTATED
//     below, function by function, each citing the reference file:line it
//     follows.  The reference's own translation unit needs Eigen, ROS, FFTW and
//     JCQP, none of which exist in this image, so it cannot be compiled here.
//   * The QP itself is solved by the REAL reference solver: qpOASES 3.2.0, the
//     vendored copy under be2r_cmpc_unitree/src/third_party/qpOASES, compiled
//     from the sources where they lie by oracle/Makefile into oracle/_ref/.
//     The call sequence follows SolverMPC.cpp:955-964.
//
// Parity pin: the QP stage is pinned by the real solver.  The condensation
// stage has no reference-run pin (the reference holds no tests or golden
// vectors for it); it is cross-checked by an independent numpy restatement
// (oracle/cmpc_numpy.py) and by analytic properties in tests/.
//
// The pipeline is templated on the scalar type: T=float reproduces the
// reference's `typedef float fpt` arithmetic (common_types.h:14) up to
// summation order, T=double is the parity anchor of the fp64 CUDA path.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include <qpOASES.hpp>

#include "cmpc_oracle.h"

namespace {

// ----------------------------------------------------------------------------
// tiny dense helpers (row-major, heap backed)
// ----------------------------------------------------------------------------
template <class T>
struct Mat {
  int r = 0, c = 0;
  std::vector<T> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, T(0)) {}
  T& operator()(int i, int j) { return a[(size_t)i * c + j]; }
  T operator()(int i, int j) const { return a[(size_t)i * c + j]; }
  static Mat eye(int n) {
    Mat m(n, n);
    for (int i = 0; i < n; i++) m(i, i) = T(1);
    return m;
  }
};

template <class T>
Mat<T> mul(const Mat<T>& x, const Mat<T>& y) {
  Mat<T> z(x.r, y.c);
  for (int i = 0; i < x.r; i++)
    for (int k = 0; k < x.c; k++) {
      T v = x(i, k);
      if (v == T(0)) continue;
      for (int j = 0; j < y.c; j++) z(i, j) += v * y(k, j);
    }
  return z;
}

template <class T>
Mat<T> axpby(T al, const Mat<T>& x, T be, const Mat<T>& y) {
  Mat<T> z(x.r, x.c);
  for (size_t i = 0; i < z.a.size(); i++) z.a[i] = al * x.a[i] + be * y.a[i];
  return z;
}

// solve P X = Q by LU with partial pivoting (P square)
template <class T>
Mat<T> lu_solve(Mat<T> P, Mat<T> Q) {
  int n = P.r;
  for (int k = 0; k < n; k++) {
    int piv = k;
    for (int i = k + 1; i < n; i++)
      if (std::fabs(P(i, k)) > std::fabs(P(piv, k))) piv = i;
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(P(k, j), P(piv, j));
      for (int j = 0; j < Q.c; j++) std::swap(Q(k, j), Q(piv, j));
    }
    T d = P(k, k);
    for (int i = k + 1; i < n; i++) {
      T f = P(i, k) / d;
      if (f == T(0)) continue;
      for (int j = k; j < n; j++) P(i, j) -= f * P(k, j);
      for (int j = 0; j < Q.c; j++) Q(i, j) -= f * Q(k, j);
    }
  }
  for (int j = 0; j < Q.c; j++)
    for (int i = n - 1; i >= 0; i--) {
      T s = Q(i, j);
      for (int k = i + 1; k < n; k++) s -= P(i, k) * Q(k, j);
      Q(i, j) = s / P(i, i);
    }
  return Q;
}

// Dense matrix exponential: Pade approximant with scaling and squaring
// (Higham 2005), the published algorithm behind Eigen's MatrixBase::exp()
// that SolverMPC.cpp:104 calls on the 31x31 augmented matrix.  Degree 13 for
// double, degree 7 for float (Eigen picks its maximum degree by scalar type).
// Deliberately generic: it knows nothing about the block structure, so it
// checks the closed form the CUDA path uses.
template <class T>
Mat<T> expm_pade(const Mat<T>& M) {
  const int n = M.r;
  double norm1 = 0;
  for (int j = 0; j < n; j++) {
    double s = 0;
    for (int i = 0; i < n; i++) s += std::fabs((double)M(i, j));
    norm1 = std::max(norm1, s);
  }
  const bool dbl = sizeof(T) == 8;
  const double theta = dbl ? 5.371920351148152 : 3.925724783138660;
  int squarings = 0;
  if (norm1 > theta) squarings = std::max(0, (int)std::ceil(std::log2(norm1 / theta)));
  Mat<T> A = M;
  T sc = (T)std::ldexp(1.0, -squarings);
  for (auto& v : A.a) v *= sc;
  Mat<T> I = Mat<T>::eye(n);
  Mat<T> A2 = mul(A, A), A4 = mul(A2, A2), A6 = mul(A4, A2);
  Mat<T> U, V;
  if (dbl) {
    const T b[] = {(T)64764752532480000., (T)32382376266240000., (T)7771770303897600.,
                   (T)1187353796428800.,  (T)129060195264000.,   (T)10559470521600.,
                   (T)670442572800.,      (T)33522128640.,       (T)1323241920.,
                   (T)40840800.,          (T)960960.,            (T)16380.,
                   (T)182.,               (T)1.};
    Mat<T> t1 = axpby(b[13], A6, b[11], A4);
    t1 = axpby(T(1), t1, b[9], A2);
    Mat<T> t2 = mul(A6, t1);
    Mat<T> t3 = axpby(b[7], A6, b[5], A4);
    t3 = axpby(T(1), t3, b[3], A2);
    t3 = axpby(T(1), t3, b[1], I);
    U = mul(A, axpby(T(1), t2, T(1), t3));
    Mat<T> s1 = axpby(b[12], A6, b[10], A4);
    s1 = axpby(T(1), s1, b[8], A2);
    Mat<T> s2 = mul(A6, s1);
    Mat<T> s3 = axpby(b[6], A6, b[4], A4);
    s3 = axpby(T(1), s3, b[2], A2);
    s3 = axpby(T(1), s3, b[0], I);
    V = axpby(T(1), s2, T(1), s3);
  } else {
    const T b[] = {(T)17297280., (T)8648640., (T)1995840., (T)277200.,
                   (T)25200.,    (T)1512.,    (T)56.,      (T)1.};
    Mat<T> t = axpby(b[7], A6, b[5], A4);
    t = axpby(T(1), t, b[3], A2);
    t = axpby(T(1), t, b[1], I);
    U = mul(A, t);
    V = axpby(b[6], A6, b[4], A4);
    V = axpby(T(1), V, b[2], A2);
    V = axpby(T(1), V, b[0], I);
  }
  Mat<T> num = axpby(T(1), V, T(1), U);
  Mat<T> den = axpby(T(1), V, T(-1), U);
  Mat<T> R = lu_solve(den, num);
  for (int s = 0; s < squarings; s++) R = mul(R, R);
  return R;
}

// ----------------------------------------------------------------------------
// reference state build
// ----------------------------------------------------------------------------
// Rotation matrix of a unit quaternion (w,x,y,z): what Eigen's
// Quaternionf::toRotationMatrix() returns in RobotState.cpp:38.
template <class T>
void quat_to_rot(const T q[4], T R[9]) {
  T w = q[0], x = q[1], y = q[2], z = q[3];
  T tx = 2 * x, ty = 2 * y, tz = 2 * z;
  T twx = tx * w, twy = ty * w, twz = tz * w;
  T txx = tx * x, txy = ty * x, txz = tz * x;
  T tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// SolverMPC.cpp:352-361 — returns (rpy0, rpy1, rpy2) = (yaw, pitch, roll)
template <class T>
void quat_to_rpy(const T q[4], T rpy[3]) {
  T w = q[0], x = q[1], y = q[2], z = q[3];
  double as_d = std::min(-2. * (double)(x * z - w * y), .99999);
  T as = (T)as_d;
  rpy[0] = std::atan2(T(2) * (x * y + w * z), w * w + x * x - y * y - z * z);
  rpy[1] = std::asin(as);
  rpy[2] = std::atan2(T(2) * (y * z + w * x), w * w - x * x - y * y + z * z);
}

template <class T>
void inv3(const T m[9], T o[9]) {
  T c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  T det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  T id = T(1) / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

const double kBig = 5e10;  // SolverMPC.cpp:19

bool nz(double a) { return a < 0.01 && a > -0.01; }   // SolverMPC.cpp:72
bool n1(double a) { return nz(a - 1); }               // SolverMPC.cpp:77

template <class T>
struct Condensed {
  int h = 0;
  Mat<T> Adt, Bdt, Qdt;  // 13x13, 13x12, 13x6
  Mat<T> Aqp, Bqp, Qqp;  // 13h x 13, 13h x 12h, 13h x 6
  Mat<T> H;              // 12h x 12h
  std::vector<T> g;      // 12h
  std::vector<T> x0;     // 13
  std::vector<T> ub;     // 20h
  Mat<T> C;              // 20h x 12h (fmat)
};

// solve_mpc() up to the QP hand-off, SolverMPC.cpp:566-816
template <class T>
void condense(const cmpc_oracle_setup* st, const cmpc_oracle_update* up, const double f_dist[6],
              Condensed<T>& o) {
  const int h = st->horizon;
  o.h = h;
  // RobotState::set, RobotState.cpp:10-54 (note R_yaw is overwritten by the
  // full rotation at :46, the body inertia is diag(.07,.26,.242), m = 12)
  T q[4], R[9];
  for (int i = 0; i < 4; i++) q[i] = (T)up->q[i];
  quat_to_rot(q, R);
  T rf[3][4];
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 4; b++) rf[a][b] = (T)up->r[a * 4 + b];
  const T Ib[3] = {(T)st->inertia[0], (T)st->inertia[1], (T)st->inertia[2]};
  const T mass = (T)st->mass;
  T rpy[3];
  quat_to_rpy(q, rpy);
  // SolverMPC.cpp:592
  o.x0.assign(13, T(0));
  o.x0[0] = rpy[2]; o.x0[1] = rpy[1]; o.x0[2] = rpy[0];
  for (int i = 0; i < 3; i++) {
    o.x0[3 + i] = (T)up->p[i];
    o.x0[6 + i] = (T)up->w[i];
    o.x0[9 + i] = (T)up->v[i];
  }
  o.x0[12] = (T)-9.8f;
  // SolverMPC.cpp:593  I_world = R I_body R^T
  T Iw[9], Iinv[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      T s = 0;
      for (int k = 0; k < 3; k++) s += R[i * 3 + k] * Ib[k] * R[j * 3 + k];
      Iw[i * 3 + j] = s;
    }
  inv3(Iw, Iinv);
  // ct_ss_mats, SolverMPC.cpp:260-279
  Mat<T> M(31, 31);
  M(3, 9) = 1; M(4, 10) = 1; M(5, 11) = 1;
  M(11, 9) = (T)up->x_drag;
  M(11, 12) = 1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) M(i, 6 + j) = R[j * 3 + i];
  for (int b = 0; b < 4; b++) {
    T rx = rf[0][b], ry = rf[1][b], rz = rf[2][b];
    T cm[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};  // SolverMPC.cpp:252-257
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        T s = 0;
        for (int k = 0; k < 3; k++) s += Iinv[i * 3 + k] * cm[k * 3 + j];
        M(6 + i, 13 + b * 3 + j) = s;
      }
    for (int i = 0; i < 3; i++) M(9 + i, 13 + b * 3 + i) = T(1) / mass;
  }
  // Q_ct, SolverMPC.cpp:607-614
  for (int i = 6; i < 12; i++) M(i, 25 + i - 6) = 1;
  // c2qp, SolverMPC.cpp:96-146
  const T dt = (T)st->dt;
  for (auto& v : M.a) v *= dt;
  Mat<T> E = expm_pade(M);
  o.Adt = Mat<T>(13, 13); o.Bdt = Mat<T>(13, 12); o.Qdt = Mat<T>(13, 6);
  for (int i = 0; i < 13; i++) {
    for (int j = 0; j < 13; j++) o.Adt(i, j) = E(i, j);
    for (int j = 0; j < 12; j++) o.Bdt(i, j) = E(i, 13 + j);
    for (int j = 0; j < 6; j++) o.Qdt(i, j) = E(i, 25 + j);
  }
  std::vector<Mat<T>> pw(h + 1);
  pw[0] = Mat<T>::eye(13);
  for (int i = 1; i <= h; i++) pw[i] = mul(o.Adt, pw[i - 1]);
  o.Aqp = Mat<T>(13 * h, 13); o.Bqp = Mat<T>(13 * h, 12 * h); o.Qqp = Mat<T>(13 * h, 6);
  for (int r = 0; r < h; r++) {
    for (int i = 0; i < 13; i++)
      for (int j = 0; j < 13; j++) o.Aqp(13 * r + i, j) = pw[r + 1](i, j);
    for (int c = 0; c <= r; c++) {
      Mat<T> pb = mul(pw[r - c], o.Bdt);
      Mat<T> pq = mul(pw[r - c], o.Qdt);
      for (int i = 0; i < 13; i++) {
        for (int j = 0; j < 12; j++) o.Bqp(13 * r + i, 12 * c + j) = pb(i, j);
        for (int j = 0; j < 6; j++) o.Qqp(13 * r + i, j) += pq(i, j);
      }
    }
  }
  // weights / trajectory / bounds, SolverMPC.cpp:624-665
  std::vector<T> S(13 * h, T(0)), Xd(13 * h, T(0));
  for (int i = 0; i < h; i++)
    for (int j = 0; j < 12; j++) {
      S[13 * i + j] = (T)up->weights[j];
      Xd[13 * i + j] = (T)up->traj[12 * i + j];
    }
  o.ub.assign(20 * h, T(0));
  for (int k = 0; k < 4 * h; k++) {
    for (int j = 0; j < 4; j++) o.ub[5 * k + j] = (T)kBig;
    o.ub[5 * k + 4] = (T)((T)up->gait[k] * (T)st->f_max);
  }
  const T mu_inv = T(1) / (T)st->mu;
  o.C = Mat<T>(20 * h, 12 * h);
  for (int k = 0; k < 4 * h; k++) {
    o.C(5 * k + 0, 3 * k + 0) = mu_inv;  o.C(5 * k + 0, 3 * k + 2) = 1;
    o.C(5 * k + 1, 3 * k + 0) = -mu_inv; o.C(5 * k + 1, 3 * k + 2) = 1;
    o.C(5 * k + 2, 3 * k + 1) = mu_inv;  o.C(5 * k + 2, 3 * k + 2) = 1;
    o.C(5 * k + 3, 3 * k + 1) = -mu_inv; o.C(5 * k + 3, 3 * k + 2) = 1;
    o.C(5 * k + 4, 3 * k + 2) = 1;
  }
  // H and g, SolverMPC.cpp:806-814 (f_dist is f_est or the zero vector, the
  // caller applies the 500-sample rule)
  std::vector<T> e(13 * h, T(0));
  for (int i = 0; i < 13 * h; i++) {
    T s = 0;
    for (int j = 0; j < 13; j++) s += o.Aqp(i, j) * o.x0[j];
    for (int j = 0; j < 6; j++) s += o.Qqp(i, j) * (T)f_dist[j];
    e[i] = S[i] * (s - Xd[i]);
  }
  const int n = 12 * h;
  o.H = Mat<T>(n, n);
  o.g.assign(n, T(0));
  // qH = 2 (Bqp' S Bqp + alpha I): row-scaled copy, then rank-1 accumulation over the 13h rows
  // (dense like the reference's Eigen product; contiguous inner loop so the compiler vectorises it)
  Mat<T> SB(13 * h, n);
  for (int k = 0; k < 13 * h; k++)
    for (int j = 0; j < n; j++) SB(k, j) = S[k] * o.Bqp(k, j);
  for (int k = 0; k < 13 * h; k++) {
    const T* bk = &o.Bqp.a[(size_t)k * n];
    const T* sk = &SB.a[(size_t)k * n];
    const T ek = e[k];
    for (int i = 0; i < n; i++) {
      const T bi = bk[i];
      if (bi == T(0)) continue;
      T* hi = &o.H.a[(size_t)i * n];
      for (int j = 0; j < n; j++) hi[j] += bi * sk[j];
      o.g[i] += bi * ek;
    }
  }
  for (int i = 0; i < n; i++) {
    o.H(i, i) += (T)up->alpha;
    for (int j = 0; j < n; j++) o.H(i, j) *= 2;
    o.g[i] *= 2;
  }
}

// Swing-foot elimination and the qpOASES hand-off, SolverMPC.cpp:841-983.
template <class T>
int reduce_and_solve(const Condensed<T>& c, int nWSR_in, cmpc_oracle_result* res) {
  const int h = c.h, nv = 12 * h, nc = 20 * h;
  std::vector<double> Hq((size_t)nv * nv), gq(nv), Aq((size_t)nc * nv), lb(nc, 0.0), ub(nc);
  for (int i = 0; i < nv; i++) {
    gq[i] = (double)c.g[i];
    for (int j = 0; j < nv; j++) Hq[(size_t)i * nv + j] = (double)c.H(i, j);
  }
  for (int i = 0; i < nc; i++) {
    ub[i] = (double)c.ub[i];
    for (int j = 0; j < nv; j++) Aq[(size_t)i * nv + j] = (double)c.C(i, j);
  }
  std::vector<char> ve(nv, 0), ce(nc, 0);
  int new_vars = nv, new_cons = nc;
  for (int i = 0; i < nc; i++) {
    if (!(nz(lb[i]) && nz(ub[i]))) continue;
    for (int j = 0; j < nv; j++)
      if (n1(Aq[(size_t)i * nv + j])) {
        new_vars -= 3; new_cons -= 5;
        int cs = (j * 5) / 3 - 3;
        ve[j - 2] = ve[j - 1] = ve[j] = 1;
        for (int k = 0; k < 5; k++) ce[cs + k] = 1;
      }
  }
  std::vector<int> vi, ci;
  for (int i = 0; i < nv; i++) if (!ve[i]) vi.push_back(i);
  for (int i = 0; i < nc; i++) if (!ce[i]) ci.push_back(i);
  res->n_var = new_vars;
  res->n_con = new_cons;
  std::fill(res->x, res->x + nv, 0.0);
  std::fill(res->con_status, res->con_status + nc, (int8_t)0);
  std::fill(res->y_con, res->y_con + nc, 0.0);
  for (int i = 0; i < nv; i++) res->var_elim[i] = ve[i];
  res->objective = 0;
  res->nwsr = 0;
  res->qp_return = 0;
  res->qp_status_ok = 1;
  if (new_vars <= 0) return 0;  // nothing in contact: the reference would hand qpOASES an empty problem
  std::vector<double> Hr((size_t)new_vars * new_vars), gr(new_vars), Ar((size_t)new_cons * new_vars),
      lbr(new_cons), ubr(new_cons), xr(new_vars, 0.0);
  for (int i = 0; i < new_vars; i++) {
    gr[i] = gq[vi[i]];
    for (int j = 0; j < new_vars; j++) Hr[(size_t)i * new_vars + j] = Hq[(size_t)vi[i] * nv + vi[j]];
  }
  for (int a = 0; a < new_cons; a++) {
    for (int b = 0; b < new_vars; b++) {
      float cval = (float)Aq[(size_t)ci[a] * nv + vi[b]];  // SolverMPC.cpp:941 narrows to float
      Ar[(size_t)a * new_vars + b] = cval;
    }
    lbr[a] = lb[ci[a]];
    ubr[a] = ub[ci[a]];
  }
  if (res->H_red) std::memcpy(res->H_red, Hr.data(), Hr.size() * sizeof(double));
  if (res->g_red) std::memcpy(res->g_red, gr.data(), gr.size() * sizeof(double));
  // SolverMPC.cpp:955-964
  qpOASES::QProblem problem(new_vars, new_cons);
  qpOASES::Options op;
  op.setToMPC();
  op.printLevel = qpOASES::PL_NONE;
  problem.setOptions(op);
  qpOASES::int_t nWSR = nWSR_in;
  int rv = problem.init(Hr.data(), gr.data(), Ar.data(), NULL, NULL, lbr.data(), ubr.data(), nWSR);
  int rv2 = problem.getPrimalSolution(xr.data());
  res->qp_return = rv;
  res->qp_status_ok = (rv2 == qpOASES::SUCCESSFUL_RETURN) ? 1 : 0;
  res->nwsr = (int)nWSR;
  for (int i = 0; i < new_vars; i++) res->x[vi[i]] = xr[i];
  if (res->qp_status_ok) {
    res->objective = problem.getObjVal();
    std::vector<double> y(new_vars + new_cons);
    problem.getDualSolution(y.data());
    qpOASES::Constraints cons;
    problem.getConstraints(cons);
    for (int a = 0; a < new_cons; a++) {
      res->con_status[ci[a]] = (int8_t)cons.getStatus(a);
      res->y_con[ci[a]] = y[new_vars + a];
    }
  }
  return 0;
}

template <class T>
int solve_one(const cmpc_oracle_setup* st, const cmpc_oracle_update* up, const double* f_dist,
              cmpc_oracle_result* res) {
  static const double zero6[6] = {0, 0, 0, 0, 0, 0};
  Condensed<T> c;
  condense<T>(st, up, f_dist ? f_dist : zero6, c);
  const int h = st->horizon;
  if (res->H_full)
    for (int i = 0; i < 12 * h; i++)
      for (int j = 0; j < 12 * h; j++) res->H_full[(size_t)i * 12 * h + j] = (double)c.H(i, j);
  if (res->g_full)
    for (int i = 0; i < 12 * h; i++) res->g_full[i] = (double)c.g[i];
  if (res->AdtBdtQdt)
    for (int i = 0; i < 13; i++) {
      for (int j = 0; j < 13; j++) res->AdtBdtQdt[i * 31 + j] = (double)c.Adt(i, j);
      for (int j = 0; j < 12; j++) res->AdtBdtQdt[i * 31 + 13 + j] = (double)c.Bdt(i, j);
      for (int j = 0; j < 6; j++) res->AdtBdtQdt[i * 31 + 25 + j] = (double)c.Qdt(i, j);
    }
  return reduce_and_solve<T>(c, st->nwsr > 0 ? st->nwsr : 100, res);
}

// ----------------------------------------------------------------------------
// periodic disturbance estimator, SolverMPC.cpp:404-553 and :688-798
// ----------------------------------------------------------------------------
// SolverMPC.cpp:404-437: float kernel, double data, edge samples repeated
std::vector<double> gauss_blur(const std::vector<double>& d, float sigma) {
  int radius = (int)std::ceil(3 * sigma);
  std::vector<float> k(2 * radius + 1);
  float sum = 0.0f;
  for (int i = -radius; i <= radius; i++) {
    float v = (float)std::exp(-0.5 * (i * i) / (sigma * sigma));
    k[i + radius] = v;
    sum += v;
  }
  for (auto& v : k) v /= sum;
  int n = (int)d.size();
  std::vector<double> out(n, 0.0);
  for (int i = 0; i < n; i++) {
    double acc = 0;
    for (int j = -radius; j <= radius; j++) {
      int idx = std::min(std::max(i + j, 0), n - 1);
      acc += d[idx] * k[j + radius];
    }
    out[i] = acc;
  }
  return out;
}

}  // namespace

extern "C" {

int cmpc_oracle_solve(const cmpc_oracle_setup* st, const cmpc_oracle_update* up, const double* f_dist,
                      int use_float, cmpc_oracle_result* res) {
  if (st->horizon < 1 || st->horizon > 19) return -1;  // SolverMPC.cpp:113
  return use_float ? solve_one<float>(st, up, f_dist, res) : solve_one<double>(st, up, f_dist, res);
}

// Sinusoid guess of SolverMPC.cpp:478-541 on one 400-sample window:
// band-pass by difference of two Gaussian blurs (:714-721), DFT magnitude peak
// over bins 1..n/2 (:497-511; the reference calls FFTW's r2c, a plain DFT here),
// amplitude sqrt(2)*std, offset mean, phase 0.
int cmpc_oracle_fit_window(const double* t, const double* d, int n, double out[4]) {
  if (n < 4) return -1;
  std::vector<double> w(d, d + n);
  std::vector<double> b1 = gauss_blur(w, 7.0f), b2 = gauss_blur(w, 27.0f), y(n);
  for (int i = 0; i < n; i++) y[i] = b1[i] - b2[i];
  double dt = t[1] - t[0];
  int best = 1;
  double best_mag = -1;
  for (int k = 1; k <= n / 2; k++) {
    double re = 0, im = 0;
    for (int i = 0; i < n; i++) {
      double ang = -2.0 * M_PI * (double)k * (double)i / (double)n;
      re += y[i] * std::cos(ang);
      im += y[i] * std::sin(ang);
    }
    double mag = std::sqrt(re * re + im * im);
    if (mag > best_mag) { best_mag = mag; best = k; }
  }
  double freq = std::fabs((double)best / ((double)n * dt));
  double mean = 0;
  for (int i = 0; i < n; i++) mean += y[i];
  mean /= n;
  double var = 0;
  for (int i = 0; i < n; i++) var += (y[i] - mean) * (y[i] - mean);
  double sd = std::sqrt(var / n);
  out[0] = mean;                 // stat (offset)
  out[1] = sd * std::sqrt(2.0);  // amplitude
  out[2] = freq;                 // Hz (omega / 2 pi)
  out[3] = 0.0;                  // phase
  return 0;
}

// One call of the adaptive bookkeeping in solve_mpc(), SolverMPC.cpp:688-798:
// push the sample, re-fit while 400 <= len <= 500, refresh f_est[3], decide
// whether g sees f_est (len > 500).  Returns 1 if f_dist_out is f_est.
int cmpc_oracle_adapt_step(cmpc_oracle_adapt* a, double sim_time, double f_ext3, double f_dist_out[6]) {
  const int window = 400;
  if (a->len < CMPC_ORACLE_HIST_MAX) {
    a->t_hist[a->len] = (float)sim_time;
    a->d_hist[a->len] = (float)f_ext3;
    a->len++;
  }
  if (a->len >= window) {
    if (a->len <= 500) {
      std::vector<double> t(window), d(window);
      for (int i = 0; i < window; i++) {
        t[i] = a->t_hist[a->len - window + i];
        d[i] = a->d_hist[a->len - window + i];
      }
      cmpc_oracle_fit_window(t.data(), d.data(), window, a->est);
    }
    float simt = (float)sim_time;
    double comp = a->est[1] + std::sin(2 * M_PI * simt * a->est[2] + a->est[3]);  // :766
    a->f_est[3] = (float)comp;                                                      // :772
  }
  for (int i = 0; i < 6; i++) a->f_est_smoothed[i] = 0.95f * a->f_est_smoothed[i] + 0.05f * a->f_est[i];
  a->f_est_static3 = 0.97f * a->f_est_static3 + 0.03f * (float)f_ext3;
  int use = a->len > 500;
  for (int i = 0; i < 6; i++) f_dist_out[i] = use ? (double)a->f_est[i] : 0.0;
  return use;
}

// Batch driver for the CPU baseline: one qpOASES solve per thread, static
// partition over `threads` std::threads.  Returns 0.
}  // extern "C"

#include <malloc.h>
#include <thread>
extern "C" int cmpc_oracle_solve_batch(const cmpc_oracle_setup* st, const cmpc_oracle_update* ups, int count,
                                       int use_float, int threads, double* forces_out /* count x 12h */,
                                       int* ok_out /* count */) {
  const int h = st->horizon;
  if (threads < 1) threads = 1;
  // keep the per-solve work matrices (100-200 KB each) on the heap arenas: the default mmap
  // threshold turns every solve into mmap/munmap + page faults, which serialises the threads
  mallopt(M_MMAP_THRESHOLD, 256 << 20);
  mallopt(M_TRIM_THRESHOLD, 512 << 20);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) {
    pool.emplace_back([=]() {
      std::vector<double> x(12 * h), y(20 * h);
      std::vector<int8_t> cs(20 * h), ve(12 * h);
      for (int i = t; i < count; i += threads) {
        cmpc_oracle_result r;
        std::memset(&r, 0, sizeof(r));
        r.x = x.data(); r.y_con = y.data(); r.con_status = cs.data(); r.var_elim = ve.data();
        cmpc_oracle_solve(st, &ups[i], nullptr, use_float, &r);
        if (forces_out) std::memcpy(forces_out + (size_t)i * 12 * h, x.data(), sizeof(double) * 12 * h);
        if (ok_out) ok_out[i] = r.qp_status_ok;
      }
    });
  }
  for (auto& th : pool) th.join();
  return 0;
}
