// Inversion kernel of the pipeline for reduced problems of up to 63 variables: K = H^-1 and x0 = -K g by a
// blocked symmetric sweep on the FP64 tensor cores (DMMA m8n8k4), ONE WARP PER INSTANCE, one warp per CTA.
//
// The assembly kernel (cmpc_condense_mma.cuh) left the scaled, bordered matrix [H g; g' .] of the instance as
// its 36 lower-triangular 8 x 8 tiles in DMMA accumulator layout.  A warp loads them into registers
// (72 doubles per lane), runs up to eight block steps and stores -(A - 2 diag) scale back in place:
//   1. invert the diagonal tile (s, s) in registers (Gauss-Jordan over warp shuffles, serial chain of 8
//      pivots) and publish the pivot rows as an 8 x 64 panel C in warp-private shared memory (tiles
//      (s, J<=s) as they are, tiles (I>s, s) transposed: the matrix is symmetric), D - I in the diagonal
//      block; tile indices are compile-time constants inside a switch over s;
//   2. M = -D^-1 C: 16 DMMAs, the two k-steps of a tile eight DMMAs apart;
//   3. every tile (I, J) += C_I' M_J: 72 independent DMMAs, the two k-steps of a tile 36 DMMAs apart.
// No block barrier anywhere: warps work on different instances and hide each other's pivot chains.  Measured
// (profiles/): the DMMA pipe is ~47 % busy at ten warps per SM; variants that form the next pivot-block
// inverse one step ahead between the update DMMAs, or that split an instance over two warps to double the
// warps per SM, measured the same or slower and were dropped.
// Row 63 (the border, g) is never pivoted and ends as g' H^-1 (see cmpc_condense_mma.cuh).
#pragma once

namespace {
// Panel of block step S from the register tiles: rows 8S..8S+7 of the symmetric matrix, tiles (S, J <= S) as
// they are and tiles (I > S, S) transposed, D - I in the diagonal block; row 63 (the border) is published as 0.
// -D^-1 of the diagonal tile of block step S (row / column 63, the border, excluded from the pivot block)
template <int S>
__device__ __forceinline__ void invert_diag(const double (&t)[36][2], double* dv, int r, int q) {
  double d0 = t[tix(S, S)][0], d1 = t[tix(S, S)][1];
  if (S == 7) {
    if (r == 7) { d0 = 0.0; d1 = (q == 3) ? 1.0 : 0.0; }
    else if (q == 3) d1 = 0.0;
  }
  warp_inv8_acc(d0, d1, r, q);
  *reinterpret_cast<double2*>(dv + r * 8 + 2 * q) = make_double2(-d0, -d1);
}

template <int S>
__device__ __forceinline__ void publish_panel(const double (&t)[36][2], double* pan, int r, int q) {
  constexpr int PS = MMA_PS;
#pragma unroll
  for (int J = 0; J <= S; J++) {
    double v0 = t[tix(S, J)][0], v1 = t[tix(S, J)][1];
    if (J == S) {
      if (r == 2 * q) v0 -= 1.0;
      if (r == 2 * q + 1) v1 -= 1.0;
    }
    if (S == 7 && r == 7) { v0 = 0.0; v1 = 0.0; }
    *reinterpret_cast<double2*>(pan + r * PS + 8 * J + 2 * q) = make_double2(v0, v1);
  }
#pragma unroll
  for (int I = S + 1; I < 8; I++) {
    pan[(2 * q) * PS + 8 * I + r] = t[tix(I, S)][0];
    pan[(2 * q + 1) * PS + 8 * I + r] = t[tix(I, S)][1];
  }
}

constexpr int INV_WPC = 1;  // warps per CTA: one, so that the register file holds ten instances per SM
constexpr int INV_WARP_SMEM = 8 * (2 * 8 * MMA_PS + 64);
}  // namespace

template <int MINB /* register cap */, bool CLK /* per-phase SM cycles (profiling launches only) */>
__global__ void __launch_bounds__(32 * INV_WPC) __maxnreg__(MINB) cmpc_invert_mma_kernel(const __grid_constant__ CmpcParams P) {
  constexpr int PS = MMA_PS;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, q = lane & 3;
  double* pan = reinterpret_cast<double*>(smem + (size_t)warp * INV_WARP_SMEM);
  double* mm = pan + 8 * PS;
  double* dv = mm + 8 * PS;
  const int count = P.count;
  unsigned flops_acc = 0u;  // algorithmic: n^3 for the symmetric inverse, 2 n^2 for x0 = -K g (n <= 63: fits 32 bits for 16 k instances per warp)
  const int fo = q * PS + r;  // fragment offset: element (k = q, row/col = r)

  long long tclk = 0;
  constexpr bool clk = CLK;
  if (clk) tclk = clock64();
#define INV_TICK(PH)                                                                    \
  if (clk) {                                                                            \
    const long long now_ = clock64();                                                   \
    if (lane == 0) atomicAdd(P.phase_cycles + (PH), (unsigned long long)(now_ - tclk)); \
    tclk = now_;                                                                        \
  }
  while (true) {
    int inst = 0;
    if (lane == 0) inst = atomicAdd(P.sched, 1);
    inst = __shfl_sync(0xffffffffu, inst, 0);
    if (inst >= count) break;
    double* slot = P.qws + (size_t)inst * P.qws_stride;
    const int* hdr = reinterpret_cast<const int*>(slot + P.qws_goff + 2 * P.nmax + 2);
    const int nc = hdr[0];
    if (hdr[1] != CMPC_ST_SOLVED) continue;
    const int n = 3 * nc, nblk = (n + 7) >> 3;
    flops_acc += (unsigned)(n * n * (n + 2));
    double t[36][2];
#pragma unroll
    for (int k = 0; k < 36; k++) {
      const double2 v = *reinterpret_cast<const double2*>(slot + k * 64 + lane * 2);
      t[k][0] = v.x;
      t[k][1] = v.y;
    }
    INV_TICK(CMPC_PH_LOAD)
#pragma unroll 1
    for (int s = 0; s < nblk; s++) {
      // 1. publish the panel (static tile indices per block step)
#define INV_HEAD(S) invert_diag<S>(t, dv, r, q); publish_panel<S>(t, pan, r, q);
      switch (s) {
        case 0: INV_HEAD(0) break;
        case 1: INV_HEAD(1) break;
        case 2: INV_HEAD(2) break;
        case 3: INV_HEAD(3) break;
        case 4: INV_HEAD(4) break;
        case 5: INV_HEAD(5) break;
        case 6: INV_HEAD(6) break;
        default: INV_HEAD(7) break;
      }
#undef INV_HEAD
      __syncwarp();
      INV_TICK(CMPC_PH_WAIT)
      // 2. M = -D^-1 C: the two k-steps of a tile are issued four DMMAs apart
      {
        const double a0 = dv[r * 8 + q], a1 = dv[r * 8 + 4 + q];
#pragma unroll
        for (int half = 0; half < 2; half++) {  // four tiles at a time: the accumulators share the register file with all 36 tiles
          double mt[4][2];
#pragma unroll
          for (int J = 0; J < 4; J++) {
            mt[J][0] = 0.0;
            mt[J][1] = 0.0;
            dmma884(mt[J][0], mt[J][1], a0, pan[fo + 8 * (4 * half + J)]);
          }
#pragma unroll
          for (int J = 0; J < 4; J++) {
            dmma884(mt[J][0], mt[J][1], a1, pan[fo + 4 * PS + 8 * (4 * half + J)]);
            *reinterpret_cast<double2*>(mm + r * PS + 8 * (4 * half + J) + 2 * q) = make_double2(mt[J][0], mt[J][1]);
          }
        }
      }
      __syncwarp();
      INV_TICK(CMPC_PH_ADAPT)
      // 3. every tile (I, J) += C_I' M_J: 72 independent DMMAs, the two k-steps of a tile 36 DMMAs apart
      {
        // J-major: the eight A fragments of a k-step stay in registers, the B fragments stream from shared memory
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
          double pf[8];
#pragma unroll
          for (int I = 0; I < 8; I++) pf[I] = pan[fo + 4 * PS * ks + 8 * I];
#pragma unroll
          for (int J = 0; J < 8; J++) {
            const double mj = mm[fo + 4 * PS * ks + 8 * J];
#pragma unroll
            for (int I = J; I < 8; I++) dmma884(t[tix(I, J)][0], t[tix(I, J)][1], pf[I], mj);
          }
        }
        __syncwarp();  // every lane is done with dv, pan and mm of this step
      }
      INV_TICK(CMPC_PH_SWEEP)
    }
    // K_ij = -(A_ij - 2 d_ij) scale in place; x0 = -scale A[63][:]
    const double scale = slot[P.qws_goff + 2 * P.nmax];
#pragma unroll
    for (int I = 0; I < 8; I++)
#pragma unroll
      for (int J = 0; J <= I; J++) {
        const double a0 = t[tix(I, J)][0], a1 = t[tix(I, J)][1];
        double2 kv;
        kv.x = -(a0 - ((I == J && r == 2 * q) ? 2.0 : 0.0)) * scale;
        kv.y = -(a1 - ((I == J && r == 2 * q + 1) ? 2.0 : 0.0)) * scale;
        *reinterpret_cast<double2*>(slot + tix(I, J) * 64 + lane * 2) = kv;
        if (I == 7 && r == 7) {
          double* xo = slot + P.qws_goff + P.nmax;
          const int j = 8 * J + 2 * q;
          if (j < n) xo[j] = -scale * a0;
          if (j + 1 < n) xo[j + 1] = -scale * a1;
        }
      }
    INV_TICK(CMPC_PH_LOAD)
  }
#undef INV_TICK
  if (lane == 0 && P.flops && flops_acc) atomicAdd(P.flops + CMPC_K_INVERT, (unsigned long long)flops_acc);
}
