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
// Panel of block step S from the register tiles: rows 8S..8S+7 of the symmetric matrix, tiles (S, J <= S) as
// they are and tiles (I > S, S) transposed, D - I in the diagonal block; row 63 (the border) is published as 0.
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
  double flops_acc = 0.0;  // algorithmic: n^3 for the symmetric inverse, 2 n^2 for x0 = -K g
  const int fo = q * PS + r;  // fragment offset: element (k = q, row/col = r)

  long long tclk = 0;
  const bool clk = P.phase_cycles != nullptr;
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
    const double scale = slot[P.qws_goff + 2 * P.nmax];
    flops_acc += (double)n * n * n + 2.0 * (double)n * n;
    double t[36][2];
#pragma unroll
    for (int k = 0; k < 36; k++) {
      const double2 v = *reinterpret_cast<const double2*>(slot + k * 64 + lane * 2);
      t[k][0] = v.x;
      t[k][1] = v.y;
    }
    INV_TICK(CMPC_PH_LOAD)
    // -D^-1 of the first diagonal tile; later ones are formed one step ahead, interleaved with the update DMMAs
    {
      double d0 = t[0][0], d1 = t[0][1];
      warp_inv8_acc(d0, d1, r, q);
      *reinterpret_cast<double2*>(dv + r * 8 + 2 * q) = make_double2(-d0, -d1);
    }
#pragma unroll 1
    for (int s = 0; s < nblk; s++) {
      // 1. publish the panel (static tile indices per block step)
      switch (s) {
        case 0: publish_panel<0>(t, pan, r, q); break;
        case 1: publish_panel<1>(t, pan, r, q); break;
        case 2: publish_panel<2>(t, pan, r, q); break;
        case 3: publish_panel<3>(t, pan, r, q); break;
        case 4: publish_panel<4>(t, pan, r, q); break;
        case 5: publish_panel<5>(t, pan, r, q); break;
        case 6: publish_panel<6>(t, pan, r, q); break;
        default: publish_panel<7>(t, pan, r, q); break;
      }
      __syncwarp();
      INV_TICK(CMPC_PH_WAIT)
      // 2. M = -D^-1 C: the two k-steps of a tile are issued eight DMMAs apart
      {
        const double a0 = dv[r * 8 + q], a1 = dv[r * 8 + 4 + q];
        double mt[8][2];
#pragma unroll
        for (int J = 0; J < 8; J++) {
          mt[J][0] = 0.0;
          mt[J][1] = 0.0;
          dmma884(mt[J][0], mt[J][1], a0, pan[fo + 8 * J]);
        }
#pragma unroll
        for (int J = 0; J < 8; J++) {
          dmma884(mt[J][0], mt[J][1], a1, pan[fo + 4 * PS + 8 * J]);
          *reinterpret_cast<double2*>(mm + r * PS + 8 * J + 2 * q) = make_double2(mt[J][0], mt[J][1]);
        }
      }
      __syncwarp();
      INV_TICK(CMPC_PH_ADAPT)
      // 3. every tile (I, J) += C_I' M_J.  First the next diagonal tile on its own, so that its inversion (a serial
      //    chain of shuffles and reciprocals) can be scheduled between the 72 independent DMMAs that follow.
      {
        const int sn = (s + 1 < 8) ? s + 1 : 7;
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int I = 0; I < 8; I++)
          if (I == sn) { d0 = t[tix(I, I)][0]; d1 = t[tix(I, I)][1]; }
        {
          const double pn0 = pan[fo + 8 * sn], pn1 = pan[fo + 4 * PS + 8 * sn];
          const double mn0 = mm[fo + 8 * sn], mn1 = mm[fo + 4 * PS + 8 * sn];
          dmma884(d0, d1, pn0, mn0);
          dmma884(d0, d1, pn1, mn1);
        }
        if (sn == 7) {  // the border row / column of block 7 is excluded from the pivot block
          if (r == 7) { d0 = 0.0; d1 = (q == 3) ? 1.0 : 0.0; }
          else if (q == 3) d1 = 0.0;
        }
        double mf[8][2], pf[8][2];
#pragma unroll
        for (int J = 0; J < 8; J++) {
          mf[J][0] = mm[fo + 8 * J];
          mf[J][1] = mm[fo + 4 * PS + 8 * J];
          pf[J][0] = pan[fo + 8 * J];
          pf[J][1] = pan[fo + 4 * PS + 8 * J];
        }
#pragma unroll
        for (int I = 0; I < 8; I++)
#pragma unroll
          for (int J = 0; J <= I; J++) dmma884(t[tix(I, J)][0], t[tix(I, J)][1], pf[I][0], mf[J][0]);
        warp_inv8_acc(d0, d1, r, q);
#pragma unroll
        for (int I = 0; I < 8; I++)
#pragma unroll
          for (int J = 0; J <= I; J++) dmma884(t[tix(I, J)][0], t[tix(I, J)][1], pf[I][1], mf[J][1]);
        __syncwarp();  // every lane is done with dv, pan and mm of this step
        *reinterpret_cast<double2*>(dv + r * 8 + 2 * q) = make_double2(-d0, -d1);
      }
      INV_TICK(CMPC_PH_SWEEP)
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
    INV_TICK(CMPC_PH_LOAD)
  }
#undef INV_TICK
  if (lane == 0 && P.flops && flops_acc > 0.0) atomicAdd(P.flops + CMPC_K_INVERT, (unsigned long long)flops_acc);
}
