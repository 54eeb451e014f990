// Inversion kernel of the pipeline for reduced problems of up to 63 variables: K = H^-1 and x0 = -K g by a
// blocked symmetric sweep on the FP64 tensor cores (DMMA m8n8k4), ONE WARP PER INSTANCE.
//
// The assembly kernel (cmpc_condense_mma.cuh) left the scaled, bordered matrix [H g; g' .] of the instance as
// its 36 lower-triangular 8 x 8 tiles in DMMA accumulator layout.  A warp loads them into registers
// (72 doubles per lane), runs up to eight block steps and stores -(A - 2 diag) scale back in place:
//   1. invert the diagonal tile (s, s) in registers (Gauss-Jordan over warp shuffles, serial chain of 8
//      pivots) and publish the pivot rows as an 8 x 64 panel C in warp-private shared memory (tiles
//      (s, J<=s) as they are, tiles (I>s, s) transposed: the matrix is symmetric), D - I in the diagonal block;
//   2. M = -D^-1 C: 16 DMMAs;
//   3. every tile (I, J) += C_I' M_J: 72 DMMAs, operands fetched once per tile row / column.
// No block barrier anywhere: warps of a CTA work on different instances and hide each other's pivot
// chains; the kernel is bound by the FP64 tensor pipe (88 DMMAs per step, 16 cycles each per SM quadrant).
// Row 63 (the border, g) is never pivoted and ends as g' H^-1 (see cmpc_condense_mma.cuh).
#pragma once

namespace {
constexpr int INV_WPC = 4;  // independent warps (instances in flight) per CTA
constexpr int INV_WARP_SMEM = 8 * (2 * 8 * MMA_PS + 64);
}  // namespace

template <int MINB>
__global__ void __launch_bounds__(32 * INV_WPC, MINB) cmpc_invert_mma_kernel(const __grid_constant__ CmpcParams P) {
  constexpr int PS = MMA_PS;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, q = lane & 3;
  double* pan = reinterpret_cast<double*>(smem + (size_t)warp * INV_WARP_SMEM);
  double* mm = pan + 8 * PS;
  double* dv = mm + 8 * PS;
  const int count = P.count;
  const int fo = q * PS + r;  // fragment offset: element (k = q, row/col = r)

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
    const double scale = slot[P.qws_goff + 2 * P.nmax];
    double t[36][2];
#pragma unroll
    for (int k = 0; k < 36; k++) {
      const double2 v = *reinterpret_cast<const double2*>(slot + k * 64 + lane * 2);
      t[k][0] = v.x;
      t[k][1] = v.y;
    }
#pragma unroll 1
    for (int s = 0; s < nblk; s++) {
      const bool excl = (s == 7);  // block 7 holds the border row 63: it is not a pivot
      // 1. -D^-1 from the diagonal tile; publish the panel with D - I in the diagonal block
      {
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int I = 0; I < 8; I++)
          if (I == s) { d0 = t[tix(I, I)][0]; d1 = t[tix(I, I)][1]; }
        if (excl) {
          if (r == 7) { d0 = 0.0; d1 = (q == 3) ? 1.0 : 0.0; }
          else if (q == 3) d1 = 0.0;
        }
        warp_inv8_acc(d0, d1, r, q);
        *reinterpret_cast<double2*>(dv + r * 8 + 2 * q) = make_double2(-d0, -d1);
      }
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
            if (I == 7 && r == 7) { v0 = 0.0; v1 = 0.0; }
            *reinterpret_cast<double2*>(pan + r * PS + 8 * J + 2 * q) = make_double2(v0, v1);
          } else if (J == s) {
            pan[(2 * q) * PS + 8 * I + r] = t[tix(I, J)][0];
            pan[(2 * q + 1) * PS + 8 * I + r] = t[tix(I, J)][1];
          }
        }
      __syncwarp();
      // 2. M = -D^-1 C
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
      // 3. every tile (I, J) += C_I' M_J
      {
        double mf[8][2];
#pragma unroll
        for (int J = 0; J < 8; J++) {
          mf[J][0] = mm[fo + 8 * J];
          mf[J][1] = mm[fo + 4 * PS + 8 * J];
        }
#pragma unroll
        for (int I = 0; I < 8; I++) {
          if (I < nblk || I == 7) {
            const double p0 = pan[fo + 8 * I], p1 = pan[fo + 4 * PS + 8 * I];
#pragma unroll
            for (int J = 0; J <= I; J++) {
              dmma884(t[tix(I, J)][0], t[tix(I, J)][1], p0, mf[J][0]);
              dmma884(t[tix(I, J)][0], t[tix(I, J)][1], p1, mf[J][1]);
            }
          }
        }
      }
      __syncwarp();  // the next publish overwrites pan and dv
    }
    // K_ij = -(A_ij - 2 d_ij) scale in place; x0 = -scale A[63][:]
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
  }
}
