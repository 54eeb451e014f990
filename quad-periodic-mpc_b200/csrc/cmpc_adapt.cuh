// Periodic-disturbance estimator stage of the solve kernel (Adaptive MPC hook), fused in front of the
// condensation so that the same launch estimates xi and applies Q_qp*xi in the gradient.
//
// Reference (SolverMPC.cpp):
//   :704-721  last 400 samples of f_ext[3]; band-pass = Gaussian blur(sigma 7) - Gaussian blur(sigma 27)
//   :478-541  fit_sin(): frequency = |DFT| peak over bins 1..n/2, amplitude = sqrt(2)*std, offset = mean,
//             phase = 0  (the non-linear refinement is a stub in the reference)
//   :766-772  compensatory_force = amp + sin(2 pi t freq + phase);  f_est[3] = compensatory_force
//   :808-814  g uses Q_qp * f_est once the history passed 500 samples
//
// The 400-point DFT is evaluated as a 20 x 20 Cooley-Tukey decomposition (two passes of twenty
// 20-point DFTs with a twiddle in between) instead of FFTW's plan: ~35 k FMAs per instance instead
// of 160 k for the direct sum, same magnitudes to rounding.
#pragma once

namespace {

// NT threads cooperate on one instance: a whole CTA (WARP = false, barriers are __syncthreads) or one warp of
// a CTA whose warps work on different instances (WARP = true, NT = 32, barriers are __syncwarp).
template <int NT, bool WARP>
__device__ __forceinline__ void estimate_disturbance_impl(const CmpcParams& P, int inst, int tid, double* work,
                                                          double* red, double* est_out /* shared, 4 doubles */) {
  static_assert(!WARP || NT == 32, "the warp variant is warp-synchronous");
  auto sync = [] { if (WARP) __syncwarp(); else __syncthreads(); };
  constexpr int N = CMPC_ADAPT_WINDOW;
  double* y = work;            // band-passed window
  double* tre = work + N;      // first-pass DFT / staged raw window
  double* tim = work + 2 * N;
  const float* wd = P.win_d + (size_t)inst * N;
  const float* wt = P.win_t + (size_t)inst * N;
  for (int i = tid; i < N; i += NT) tre[i] = (double)wd[i];
  sync();
  // band-pass: difference of the two blurs, edge samples repeated (SolverMPC.cpp:425-434)
  const float* g1 = P.gk;
  const float* g2 = P.gk + 2 * CMPC_GK_R1 + 1;
  for (int i = tid; i < N; i += NT) {
    double a1 = 0.0, a2 = 0.0;
    for (int j = -CMPC_GK_R1; j <= CMPC_GK_R1; j++) {
      int idx = min(max(i + j, 0), N - 1);
      a1 += tre[idx] * (double)__ldg(g1 + j + CMPC_GK_R1);
    }
    for (int j = -CMPC_GK_R2; j <= CMPC_GK_R2; j++) {
      int idx = min(max(i + j, 0), N - 1);
      a2 += tre[idx] * (double)__ldg(g2 + j + CMPC_GK_R2);
    }
    y[i] = a1 - a2;
  }
  sync();
  // mean and (population) standard deviation
  double part = 0.0;
  for (int i = tid; i < N; i += NT) part += y[i];
  const double mean = block_sum<NT>(part, red, tid) / (double)N;
  part = 0.0;
  for (int i = tid; i < N; i += NT) { double dlt = y[i] - mean; part = fma(dlt, dlt, part); }
  const double var = block_sum<NT>(part, red, tid) / (double)N;
  sync();
  // pass 1: T[n2][k1] = W400^(n2 k1) * sum_n1 y[20 n1 + n2] W20^(n1 k1)
  for (int idx = tid; idx < N; idx += NT) {
    const int n2 = idx / 20, k1 = idx - 20 * n2;
    double re = 0.0, im = 0.0;
    int m = 0;  // (n1 * k1) mod 20
    for (int n1 = 0; n1 < 20; n1++) {
      const double yv = y[20 * n1 + n2];
      re = fma(yv, __ldg(P.twiddle + 2 * (20 * m)), re);
      im = fma(yv, __ldg(P.twiddle + 2 * (20 * m) + 1), im);
      m += k1;
      if (m >= 20) m -= 20;
    }
    const int tw = (n2 * k1) % N;
    const double c = __ldg(P.twiddle + 2 * tw), s = __ldg(P.twiddle + 2 * tw + 1);
    tre[idx] = re * c - im * s;
    tim[idx] = re * s + im * c;
  }
  sync();
  // pass 2: X[k1 + 20 k2] = sum_n2 T[n2][k1] W20^(n2 k2); only bins 1..200 are searched
  double best = 1e300;
  int bidx = 1 << 30;
  for (int k = 1 + tid; k <= N / 2; k += NT) {
    const int k2 = k / 20, k1 = k - 20 * k2;
    double re = 0.0, im = 0.0;
    int m = 0;  // (n2 * k2) mod 20
    for (int n2 = 0; n2 < 20; n2++) {
      const double c = __ldg(P.twiddle + 2 * (20 * m)), s = __ldg(P.twiddle + 2 * (20 * m) + 1);
      const double a = tre[20 * n2 + k1], b = tim[20 * n2 + k1];
      re += a * c - b * s;
      im += a * s + b * c;
      m += k2;
      if (m >= 20) m -= 20;
    }
    const double neg = -(re * re + im * im);
    if (neg < best) { best = neg; bidx = k; }
  }
  block_argmin<NT>(best, bidx, red, tid);
  if (tid == 0) {
    const double dts = (double)wt[1] - (double)wt[0];
    est_out[0] = mean;
    est_out[1] = sqrt(var) * sqrt(2.0);
    est_out[2] = fabs((double)bidx / ((double)N * dts));
    est_out[3] = 0.0;
  }
  sync();
}

template <int NT>
__device__ __forceinline__ void estimate_disturbance(const CmpcParams& P, int inst, int tid, double* work, double* red,
                                                     double* est_out) {
  estimate_disturbance_impl<NT, false>(P, inst, tid, work, red, est_out);
}

__device__ __forceinline__ void estimate_disturbance_warp(const CmpcParams& P, int inst, int lane, double* work,
                                                          double* est_out) {
  estimate_disturbance_impl<32, true>(P, inst, lane, work, nullptr, est_out);
}

}  // namespace
