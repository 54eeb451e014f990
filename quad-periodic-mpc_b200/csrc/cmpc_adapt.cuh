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
  // band-pass: difference of the two blurs, edge samples repeated (SolverMPC.cpp:425-434).  The window is staged with
  // its edges already repeated (no index clamping in the tap loops) next to the taps as doubles; a thread forms four
  // consecutive outputs, so a tap and a sample are fetched once per four FMAs.  Every output is still the same
  // left-to-right FMA chain over its taps.
  constexpr int R1 = CMPC_GK_R1, R2 = CMPC_GK_R2;
  double* xe = work + N;                  // [N + 2 R2], overlays tre / tim (dead until the DFT)
  double* tap1 = work + 2 * N + 2 * R2;   // [2 R1 + 1]
  double* tap2 = tap1 + 2 * R1 + 1;       // [2 R2 + 1]
  static_assert(3 * CMPC_ADAPT_WINDOW >= 2 * CMPC_ADAPT_WINDOW + 2 * CMPC_GK_R2 + CMPC_GK_TOTAL, "estimator scratch");
  for (int i = tid; i < N + 2 * R2; i += NT) xe[i] = (double)wd[min(max(i - R2, 0), N - 1)];
  for (int i = tid; i < CMPC_GK_TOTAL; i += NT) tap1[i] = (double)__ldg(P.gk + i);
  double* w20 = work + 3 * N;             // W20^m = W400^(20 m), m = 0..19: (cos, -sin) pairs for both DFT passes
  for (int i = tid; i < 40; i += NT) w20[i] = __ldg(P.twiddle + 2 * (20 * (i >> 1)) + (i & 1));
  sync();
  for (int o = tid; o < N / 4; o += NT) {
    const int i0 = 4 * o;
    double a2[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
    {
      const double* x = xe + i0;  // output i0 + c, tap j reads sample index (i0 + c) + j - R2, i.e. xe[i0 + c + j]
      double x0 = x[0], x1 = x[1], x2 = x[2];
#pragma unroll 4
      for (int j = 0; j <= 2 * R2; j++) {
        const double x3 = x[j + 3], g = tap2[j];
        a2[0] = fma(x0, g, a2[0]);
        a2[1] = fma(x1, g, a2[1]);
        a2[2] = fma(x2, g, a2[2]);
        a2[3] = fma(x3, g, a2[3]);
        x0 = x1; x1 = x2; x2 = x3;
      }
    }
    {
      const double* x = xe + i0 + (R2 - R1);
      double x0 = x[0], x1 = x[1], x2 = x[2];
#pragma unroll 4
      for (int j = 0; j <= 2 * R1; j++) {
        const double x3 = x[j + 3], g = tap1[j];
        a1[0] = fma(x0, g, a1[0]);
        a1[1] = fma(x1, g, a1[1]);
        a1[2] = fma(x2, g, a1[2]);
        a1[3] = fma(x3, g, a1[3]);
        x0 = x1; x1 = x2; x2 = x3;
      }
    }
#pragma unroll
    for (int c = 0; c < 4; c++) y[i0 + c] = a1[c] - a2[c];
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
      re = fma(yv, w20[2 * m], re);
      im = fma(yv, w20[2 * m + 1], im);
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
      const double c = w20[2 * m], s = w20[2 * m + 1];
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
