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
  // its edges already repeated by the larger radius (no index clamping in the tap loop; the clamp does not depend on
  // the radius, so the narrow blur reads the same staged samples) and the two normalised kernels are folded into ONE
  // set of 2 R2 + 1 taps, c[j] = g7[j] - g27[j] (g7 zero beyond its radius): 65 k FMAs per instance instead of 82 k.
  // A thread forms FOUR consecutive outputs (a tap and a sample are fetched once per four FMAs): 100 of the CTA's 128
  // threads work, on three of the four SM sub-partitions — eight outputs per thread halve the loads but leave the whole
  // stage to two warps, i.e. to the FP64 pipes of two sub-partitions (measured: no faster).
  constexpr int R1 = CMPC_GK_R1, R2 = CMPC_GK_R2;
  // thread o reads samples 4 o + t: a stride of four doubles would put a half-warp on four banks, so the staged
  // window is SKEWED by one slot per four samples (sample e at e + e / 4: lane stride five doubles, conflict-free)
  constexpr int NE = N + 2 * R2, NES = NE + NE / 4 + 1;
  double* xe = work + N;                  // [NES], overlays tre / tim (dead until the DFT)
  double* tapc = work + 3 * N + 40;       // [2 R2 + 1], behind the roots of unity
  static_assert(CMPC_ADAPT_WINDOW + NES <= 3 * CMPC_ADAPT_WINDOW && 3 * CMPC_ADAPT_WINDOW + 40 + 2 * CMPC_GK_R2 + 1 <= CMPC_ADAPT_SCRATCH, "estimator scratch");
  for (int i = tid; i < NE; i += NT) xe[i + (i >> 2)] = (double)wd[min(max(i - R2, 0), N - 1)];
  for (int j = tid; j <= 2 * R2; j += NT) {
    const int j1 = j - (R2 - R1);  // index into the narrow kernel
    const double g1 = (j1 >= 0 && j1 <= 2 * R1) ? (double)__ldg(P.gk + j1) : 0.0;
    tapc[j] = g1 - (double)__ldg(P.gk + (2 * R1 + 1) + j);
  }
  double* w20 = work + 3 * N;             // W20^m = W400^(20 m), m = 0..19: (cos, -sin) pairs for both DFT passes
  for (int i = tid; i < 40; i += NT) w20[i] = __ldg(P.twiddle + 2 * (20 * (i >> 1)) + (i & 1));
  sync();
  static_assert(CMPC_ADAPT_WINDOW % 4 == 0, "four outputs per thread");
  for (int o = tid; o < N / 4; o += NT) {
    // output 4 o + c, tap j reads sample (4 o + c) + j - R2 of the window, i.e. staged sample 4 o + c + j
    const double* x = xe + 5 * o;  // staged sample 4 o + t sits at 5 o + t + t / 4
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    double xr[4];
#pragma unroll
    for (int c = 0; c < 3; c++) xr[c] = x[c];
    constexpr int NTAP = 2 * R2 + 1;
#pragma unroll 2
    for (int jb = 0; jb < NTAP; jb += 4) {
      const double* xb = x + 5 * (jb >> 2);
#pragma unroll
      for (int uu = 0; uu < 4; uu++) {
        if (jb + uu < NTAP) {
          // sample t = jb + uu + 3: slot 5 (jb / 4) + (uu + 3) + (uu + 3) / 4
          xr[3] = xb[uu + 3 + ((uu + 3) >> 2)];
          const double g = tapc[jb + uu];
#pragma unroll
          for (int c = 0; c < 4; c++) a[c] = fma(xr[c], g, a[c]);
#pragma unroll
          for (int c = 0; c < 3; c++) xr[c] = xr[c + 1];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; c++) y[4 * o + c] = a[c];
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
