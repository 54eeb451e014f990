// Inversion kernel of the pipeline for reduced problems of up to 63 variables: K = H^-1 and x0 = -K g by a
// blocked symmetric sweep on the FP64 tensor cores (DMMA m8n8k4), one MAIN WARP PER INSTANCE.
//
// The assembly kernel (cmpc_condense_mma.cuh) left the scaled, bordered matrix [H g; g' .] of the instance as
// its 36 lower-triangular 8 x 8 tiles in DMMA accumulator layout.  A warp loads them into registers
// (72 doubles per lane), runs up to eight block steps and stores -(A - 2 diag) scale back in place:
//   1. publish the pivot rows as an 8 x 64 panel C in warp-private shared memory (tiles (s, J<=s) as they are,
//      tiles (I>s, s) transposed: the matrix is symmetric), D - I in the diagonal block; tile indices are
//      compile-time constants inside a switch over s; the inverse of the diagonal tile (s, s) comes from a helper
//      warp (cmpc_invert_ws_kernel below);
//   2. M = -D^-1 C: 16 DMMAs, four tiles at a time, their B fragments loaded up front;
//   3. every tile (I, J) += C_I' M_J: 72 independent DMMAs, the two k-steps of a tile 36 DMMAs apart.
// No block barrier anywhere: warps work on different instances.
// Row 63 (the border, g) is never pivoted and ends as g' H^-1 (see cmpc_condense_mma.cuh).
// (The single-warp form of round 1 — the same warp inverts the diagonal tile — is kept under CMPC_EXPERIMENTS.)
#pragma once

namespace {
#ifdef CMPC_EXPERIMENTS
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
#endif

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

}  // namespace
#ifdef CMPC_EXPERIMENTS  // the single-warp kernel (superseded by cmpc_invert_ws_kernel below) is kept for A/B builds only
namespace {
// Hardest-first scheduling of the active-set kernel: count the constraint rows that x0 (staged in `xs`, shared
// memory) violates — five rows per contact foot-step, as cmpc_dual_fast.cuh lays them out — and file the instance.
// deferred form: the histogram atomic is issued here, its result (lane 0) is written by lpt_store one instance later,
// so the warp does not wait for the round trip
__device__ __forceinline__ int lpt_count(const CmpcParams& P, int nc, const int* hdr, const double* xs, int lane) {
  int viol = 0;
  if (nc > 0) {
    const unsigned char* gvb = reinterpret_cast<const unsigned char*>(hdr + 2) + CMPC_MAX_FS;
    for (int f = lane; f < nc; f += 32) {
      const double fx = xs[3 * f] * P.mu_inv, fy = xs[3 * f + 1] * P.mu_inv, fz = xs[3 * f + 2];
      const double tol = -P.tol_violation;
      viol += (fx + fz < tol) + (fz - fx < tol) + (fy + fz < tol) + (fz - fy < tol) + ((double)gvb[f] * P.f_max - fz < tol);
    }
    viol = __reduce_add_sync(0xffffffffu, viol);
  }
  return min(viol, 63);
}

__device__ __forceinline__ void lpt_file(const CmpcParams& P, int inst, int nc, const int* hdr, const double* xs, int lane) {
  if (!P.lpt_hist) return;
  int viol = 0;
  if (nc > 0) {
    const unsigned char* gvb = reinterpret_cast<const unsigned char*>(hdr + 2) + CMPC_MAX_FS;
    for (int f = lane; f < nc; f += 32) {
      const double fx = xs[3 * f] * P.mu_inv, fy = xs[3 * f + 1] * P.mu_inv, fz = xs[3 * f + 2];
      const double tol = -P.tol_violation;
      viol += (fx + fz < tol) + (fz - fx < tol) + (fy + fz < tol) + (fz - fy < tol) + ((double)gvb[f] * P.f_max - fz < tol);
    }
    viol = __reduce_add_sync(0xffffffffu, viol);
  }
  if (lane == 0) {
    const int key = min(viol, 63);
    const int pos = atomicAdd(P.lpt_hist + key, 1);
    P.lpt_key[inst] = (key << 24) | pos;
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
    if (hdr[1] != CMPC_ST_SOLVED) { lpt_file(P, inst, 0, hdr, pan, lane); continue; }
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
          pan[j] = -scale * a0;  // staged for lpt_file (the panel is free now)
          pan[j + 1] = -scale * a1;
        }
      }
    __syncwarp();
    lpt_file(P, inst, nc, hdr, pan, lane);
    __syncwarp();
    INV_TICK(CMPC_PH_LOAD)
  }
#undef INV_TICK
  if (lane == 0 && P.flops && flops_acc) atomicAdd(P.flops + CMPC_K_INVERT, (unsigned long long)flops_acc);
}
#endif  // CMPC_EXPERIMENTS

// ------------------------------------------------------------------------------------------------------------
// Warp-specialised kernel: the serial part of a block step — the 8-pivot chain that inverts the diagonal tile,
// 46 % of an instance's cycles when the same warp does it — belongs to a HELPER warp.  A CTA is two warpgroups: four
// main warps (one instance each, all 36 tiles in registers; setmaxnreg raises them to 216 registers) and four
// helper warps (setmaxnreg drops them to 40).  Main warp w and helper warp w + 4 sit on the same SM sub-partition
// and talk through two named barriers and 1.5 KB of shared memory:
//   main:    ... bar.sync(DV); column block s+1 of M and the update of tile (s+1, s+1) FIRST (four DMMAs), park the
//            tile in shared memory, bar.arrive(TILE); M = -D^-1 C; the 72 update DMMAs of step s; publish panel s+1 ...
//   helper:  bar.sync(TILE); invert the parked tile; write -D^-1; bar.arrive(DV)
// so the pivot chain of step s+1 runs while the main warp issues the update DMMAs of step s.  Two CTAs per SM
// (the register file: 2 x 128 x (216 + 40)), eight instances per SM, two main warps per FP64 tensor pipe.
// The helper's chain runs in fp32 and is finished by Newton steps X <- X + X (I - D X) in FP64 on the tensor cores
// (P.inv_f32; an FP64 instruction of the helper queues behind the main warps' DMMAs, fp32 and integer ones do not);
// block steps whose pivot-block inverse is large get one residual correction of M (P.inv_refine).  DESIGN.md §4.2.
// ------------------------------------------------------------------------------------------------------------
namespace {
constexpr int WS_MAIN = 4;                                   // main warps (= instances in flight) per CTA
constexpr int WS_TILE_BYTES = 8 * CMPC_KTILE_DOUBLES;                                 // the 36 tiles of one instance
constexpr int WS_PAIR_CTRL = 8 * (2 * 8 * MMA_PS + 64 + 64 + 64);                     // panel, M, -D^-1, parked tile, Newton scratch
constexpr int WS_PAIR_SMEM = WS_PAIR_CTRL + 16 + WS_TILE_BYTES;  // + control word, mbarrier, prefetched tiles of the NEXT instance
__device__ __forceinline__ void named_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
}  // namespace

template <bool CLK /* per-phase SM cycles of the main warps (profiling launches only) */>
__global__ void __launch_bounds__(64 * WS_MAIN, 2) cmpc_invert_ws_kernel(const __grid_constant__ CmpcParams P) {
  constexpr int PS = MMA_PS;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, q = lane & 3;
  const int pair = warp & (WS_MAIN - 1);
  double* pan = reinterpret_cast<double*>(smem + (size_t)pair * WS_PAIR_SMEM);
  double* mm = pan + 8 * PS;
  double* dv = mm + 8 * PS;
  double* dtile = dv + 64;
  double* nsc = dtile + 64;                                      // helper: residual of the Newton steps
  volatile int* ctrl = reinterpret_cast<volatile int*>(nsc + 64);
  uint64_t* tbar = reinterpret_cast<uint64_t*>(nsc + 65);        // mbarrier of the tile prefetch
  double* tbuf = nsc + 66;                                       // 36 tiles, filled by one cp.async.bulk
  const int BAR_TILE = 1 + 2 * pair, BAR_DV = 2 + 2 * pair;

  // The CTAs of an SM start together and every instance takes the same time, so the main warps that share a
  // sub-partition would run their DMMA phases in LOCKSTEP for the whole launch — pipe saturated, then idle
  // (scripts/ubench/dmma_phase_ubench.cu: 37 instead of 22 cycles per DMMA).  The second CTA to arrive on an SM
  // therefore starts half a block step late.
  __shared__ int sm_slot;
  if (P.sm_slots && P.inv_stagger > 0) {
    if (threadIdx.x == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      sm_slot = atomicAdd(P.sm_slots + (smid & (CMPC_SM_SLOTS - 1)), 1);
    }
    __syncthreads();
  }

  if (warp >= WS_MAIN) {
    // ---------------- helper: the pivot chains ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    while (true) {
      named_bar_sync(BAR_TILE);  // the first diagonal tile of an instance, or the end of the work
      const int nb = ctrl[0];
      if (nb < 0) break;
      for (int s = 0; s < nb; s++) {
        if (s > 0) named_bar_sync(BAR_TILE);
        const double2 d = *reinterpret_cast<const double2*>(dtile + r * 8 + 2 * q);
        double d0 = d.x, d1 = d.y;
        bool done = false;
        if (P.inv_f32) {
          // An FP64 instruction of this warp queues behind the DMMA streams of the two main warps of its sub-partition
          // (profiles/r2_ubench_helper_chain.txt), fp32 ones do not: the serial chain of eight pivots runs in fp32, and
          // X <- X + X (I - D X) on the tensor cores (four DMMAs a step) squares its error until it is below rounding.
          float f0 = (float)d0, f1 = (float)d1;
          warp_inv8_acc_f32(f0, f1, r, q);
          double x0 = (double)f0, x1 = (double)f1;
          const double nd0 = -dtile[r * 8 + q], nd1 = -dtile[r * 8 + 4 + q];
          for (int it = 0; it < 6 && !done; it++) {
            *reinterpret_cast<double2*>(dv + r * 8 + 2 * q) = make_double2(x0, x1);
            __syncwarp();
            double t0 = (r == 2 * q) ? 1.0 : 0.0, t1 = (r == 2 * q + 1) ? 1.0 : 0.0;
            dmma884(t0, t1, nd0, dv[q * 8 + r]);
            dmma884(t0, t1, nd1, dv[(4 + q) * 8 + r]);
            *reinterpret_cast<double2*>(nsc + r * 8 + 2 * q) = make_double2(t0, t1);
            // magnitude tests on the high words (integer pipe: an FP64 compare would queue behind the main warps' DMMAs);
            // a NaN compares as large
            const int tm = max(__double2hiint(t0) & 0x7fffffff, __double2hiint(t1) & 0x7fffffff);
            const bool small = !__any_sync(0xffffffffu, tm >= 0x3e45798e /* 1e-8 */);
            if (it == 0 && __any_sync(0xffffffffu, tm >= 0x3fd00000 /* 0.25 */)) break;  // fp32 was not enough: FP64 chain below
            __syncwarp();
            dmma884(x0, x1, dv[r * 8 + q], nsc[q * 8 + r]);
            dmma884(x0, x1, dv[r * 8 + 4 + q], nsc[(4 + q) * 8 + r]);
            __syncwarp();
            done = small;
          }
          if (done) { d0 = x0; d1 = x1; }
        }
        if (!done) warp_inv8_acc(d0, d1, r, q);
        *reinterpret_cast<double2*>(dv + r * 8 + 2 * q) = make_double2(-d0, -d1);
        named_bar_arrive(BAR_DV);
      }
    }
    return;
  }

  // ---------------- main: tiles in registers, DMMA ----------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 216;\n");
  if (P.sm_slots && P.inv_stagger > 0 && (sm_slot & 1)) {
    const long long s0 = clock64();
    while (clock64() - s0 < (long long)P.inv_stagger) { }
  }
  const int count = P.count;
  unsigned flops_acc = 0u;
  const int fo = q * PS + r;  // fragment offset: element (k = q, row/col = r)
  const int refine_hi = __double2hiint(P.inv_refine);  // the threshold of the panel refinement, compared on high words
  long long tclk = 0;
  if (CLK) tclk = clock64();
#define WS_TICK(PH)                                                                     \
  if (CLK) {                                                                            \
    const long long now_ = clock64();                                                   \
    if (lane == 0) atomicAdd(P.phase_cycles + (PH), (unsigned long long)(now_ - tclk)); \
    tclk = now_;                                                                        \
  }
  // diagonal tile S of the bordered matrix as the pivot block sees it: row / column 63 (the border) masked out
  auto park_tile = [&](double d0, double d1, int S) {
    if (S == 7) {
      if (r == 7) { d0 = 0.0; d1 = (q == 3) ? 1.0 : 0.0; }
      else if (q == 3) d1 = 0.0;
    }
    *reinterpret_cast<double2*>(dtile + r * 8 + 2 * q) = make_double2(d0, d1);
  };
  // Instance fetch and tile load run one instance ahead: while an instance is swept, the work counter, the header and
  // the 18 KB of tiles of the NEXT one are already on their way (one cp.async.bulk into shared memory).
  auto fetch = [&]() -> int {
    int i = 0;
    if (lane == 0) i = atomicAdd(P.sched, 1);
    return __shfl_sync(0xffffffffu, i, 0);
  };
  auto prefetch = [&](int i, int& nc_o, int& st_o, double& sc_o, int& gv_o) {
    const double* sl = P.qws + (size_t)i * P.qws_stride;
    const int* hd = reinterpret_cast<const int*>(sl + P.qws_goff + 2 * P.nmax + 2);
    nc_o = hd[0];
    st_o = hd[1];
    sc_o = sl[P.qws_goff + 2 * P.nmax];                                                      // the scale of H
    gv_o = reinterpret_cast<const unsigned char*>(hd + 2)[CMPC_MAX_FS + lane];               // fz bound of foot-step `lane`
    if (lane == 0) {
      fence_proxy_async();  // the buffer was read through the generic proxy a moment ago
      mbar_expect_tx(tbar, (uint32_t)WS_TILE_BYTES);
      bulk_g2s(tbuf, sl, (uint32_t)WS_TILE_BYTES, tbar);
    }
  };
  if (lane == 0) {
    mbar_init(tbar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  uint32_t tphase = 0u;
  int inst = fetch(), nc_cur = 0, st_cur = 0, gv_cur = 0;
  double sc_cur = 0.0;
  if (inst < count) prefetch(inst, nc_cur, st_cur, sc_cur, gv_cur);
  // the work counter is drawn TWO instances ahead (lane 0 holds the ticket, nobody waits for the atomic), and the
  // hardest-first filing of an instance is completed one instance later (same reason)
  int ticket = 0;
  if (lane == 0) ticket = atom_add_later(P.sched, 1);
  int file_inst = -1, file_key = 0, file_pos = 0;
  auto lpt_flush = [&]() {
    if (lane == 0 && file_inst >= 0) P.lpt_key[file_inst] = (file_key << 24) | file_pos;
    file_inst = -1;
  };
  auto lpt_defer = [&](int i, int key) {
    if (!P.lpt_hist) return;
    lpt_flush();
    if (lane == 0) {
      file_pos = atom_add_later(P.lpt_hist + key, 1);
      file_key = key;
      file_inst = i;
    }
  };
  while (inst < count) {
    double* slot = P.qws + (size_t)inst * P.qws_stride;
    const int* hdr = reinterpret_cast<const int*>(slot + P.qws_goff + 2 * P.nmax + 2);
    const int nc = nc_cur, gv_mine = gv_cur;
    const double scale = sc_cur;
    const bool solved = st_cur == CMPC_ST_SOLVED;
    WS_TICK(CMPC_PH_X3)
    mbar_wait(tbar, tphase);
    tphase ^= 1u;
    WS_TICK(CMPC_PH_X0)
    double t[36][2];
    if (solved) {
#pragma unroll
      for (int k = 0; k < 36; k++) {
        const double2 v = *reinterpret_cast<const double2*>(tbuf + k * 64 + lane * 2);
        t[k][0] = v.x;
        t[k][1] = v.y;
      }
    }
    __syncwarp();
    WS_TICK(CMPC_PH_X1)
    const int inst_next = __shfl_sync(0xffffffffu, ticket, 0);
    if (lane == 0) ticket = atom_add_later(P.sched, 1);
    int nc_next = 0, st_next = 0, gv_next = 0;
    double sc_next = 0.0;
    if (inst_next < count) prefetch(inst_next, nc_next, st_next, sc_next, gv_next);
    const int inst_this = inst;
    inst = inst_next;
    nc_cur = nc_next;
    st_cur = st_next;
    sc_cur = sc_next;
    gv_cur = gv_next;
    if (!solved) { lpt_defer(inst_this, 0); continue; }
    const int n = 3 * nc, nblk = (n + 7) >> 3;
    flops_acc += (unsigned)(n * n * (n + 2));
    if (lane == 0) ctrl[0] = nblk;
    park_tile(t[0][0], t[0][1], 0);
    named_bar_arrive(BAR_TILE);
    WS_TICK(CMPC_PH_LOAD)
#pragma unroll 1
    for (int s = 0; s < nblk; s++) {
      // 1. publish the panel of block step s (static tile indices per block step)
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
      WS_TICK(CMPC_PH_PUBLISH)
      named_bar_sync(BAR_DV);  // -D^-1 of this step from the helper
      WS_TICK(CMPC_PH_DVWAIT)
      // a last block of at most four real rows: its second k-step (padding rows, the border) contributes zeros
      const bool half_step = (8 * s + 4 >= n);
      const double a0 = dv[r * 8 + q], a1 = dv[r * 8 + 4 + q];
      // Block Gauss-Jordan with an explicitly inverted pivot block loses ~cond(D) digits more than the scalar sweep
      // (profiles/r2_block_gj_accuracy.txt).  One residual correction of the panel, M += -D^-1 (C + D M), gives them back;
      // it is applied only to block steps whose pivot-block inverse is large (P.inv_refine; next to none on the A1 defaults).
      const bool refine = P.inv_refine >= 0.0 &&
                          __any_sync(0xffffffffu, max(__double2hiint(a0) & 0x7fffffff, __double2hiint(a1) & 0x7fffffff) > refine_hi);
      double da0 = 0.0, da1 = 0.0;  // D itself as the A operand (the helper is done with the parked tile)
      if (refine) {
        da0 = dtile[r * 8 + q];
        da1 = dtile[r * 8 + 4 + q];
      }
      // 2a. the next diagonal tile FIRST: its column block of M (two DMMAs), its update (two DMMAs), then it is handed
      //     to the helper, whose pivot chain runs under the rest of M and the 72 update DMMAs of this step
      if (s + 1 < nblk) {
        const int o = 8 * (s + 1);
        double m0 = 0.0, m1 = 0.0;
        dmma884(m0, m1, a0, pan[fo + o]);
        dmma884(m0, m1, a1, pan[fo + 4 * PS + o]);
        *reinterpret_cast<double2*>(mm + r * PS + o + 2 * q) = make_double2(m0, m1);
        __syncwarp();
        if (refine) {
          const double2 c = *reinterpret_cast<const double2*>(pan + r * PS + o + 2 * q);
          double r0 = c.x, r1 = c.y;
          dmma884(r0, r1, da0, mm[fo + o]);
          dmma884(r0, r1, da1, mm[fo + 4 * PS + o]);
          __syncwarp();
          *reinterpret_cast<double2*>(mm + r * PS + o + 2 * q) = make_double2(r0, r1);
          __syncwarp();
          dmma884(m0, m1, a0, mm[fo + o]);
          dmma884(m0, m1, a1, mm[fo + 4 * PS + o]);
          __syncwarp();
          *reinterpret_cast<double2*>(mm + r * PS + o + 2 * q) = make_double2(m0, m1);
          __syncwarp();
        }
        double e0, e1;
        switch (s) {
          case 0: e0 = t[tix(1, 1)][0]; e1 = t[tix(1, 1)][1]; break;
          case 1: e0 = t[tix(2, 2)][0]; e1 = t[tix(2, 2)][1]; break;
          case 2: e0 = t[tix(3, 3)][0]; e1 = t[tix(3, 3)][1]; break;
          case 3: e0 = t[tix(4, 4)][0]; e1 = t[tix(4, 4)][1]; break;
          case 4: e0 = t[tix(5, 5)][0]; e1 = t[tix(5, 5)][1]; break;
          case 5: e0 = t[tix(6, 6)][0]; e1 = t[tix(6, 6)][1]; break;
          default: e0 = t[tix(7, 7)][0]; e1 = t[tix(7, 7)][1]; break;
        }
        dmma884(e0, e1, pan[fo + o], mm[fo + o]);
        dmma884(e0, e1, pan[fo + 4 * PS + o], mm[fo + 4 * PS + o]);
        park_tile(e0, e1, s + 1);
        named_bar_arrive(BAR_TILE);
      }
      // 2b. M = -D^-1 C (the column block of 2a is formed again, bit for bit the same): four tiles at a time, their eight
      //     B fragments loaded up front, the two k-steps of a tile four DMMAs apart
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        double bf[4][2], mt[4][2];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          bf[j][0] = lds_f64v(pan + fo + 8 * (4 * hf + j));
          bf[j][1] = lds_f64v(pan + fo + 4 * PS + 8 * (4 * hf + j));
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          mt[j][0] = 0.0;
          mt[j][1] = 0.0;
          dmma884v(mt[j][0], mt[j][1], a0, bf[j][0]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884v(mt[j][0], mt[j][1], a1, bf[j][1]);
#pragma unroll
        for (int j = 0; j < 4; j++)
          *reinterpret_cast<double2*>(mm + r * PS + 8 * (4 * hf + j) + 2 * q) = make_double2(mt[j][0], mt[j][1]);
      }
      if (refine) {  // four tiles at a time: R = C + D M, then M += -D^-1 R, the same operations as in 2a
        __syncwarp();
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          double rt[4][2], mt[4][2];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int J = 4 * hf + j;
            const double2 c = *reinterpret_cast<const double2*>(pan + r * PS + 8 * J + 2 * q);
            const double2 m = *reinterpret_cast<const double2*>(mm + r * PS + 8 * J + 2 * q);
            rt[j][0] = c.x;
            rt[j][1] = c.y;
            mt[j][0] = m.x;
            mt[j][1] = m.y;
            dmma884(rt[j][0], rt[j][1], da0, mm[fo + 8 * J]);
          }
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(rt[j][0], rt[j][1], da1, mm[fo + 4 * PS + 8 * (4 * hf + j)]);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; j++)
            *reinterpret_cast<double2*>(mm + r * PS + 8 * (4 * hf + j) + 2 * q) = make_double2(rt[j][0], rt[j][1]);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(mt[j][0], mt[j][1], a0, mm[fo + 8 * (4 * hf + j)]);
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(mt[j][0], mt[j][1], a1, mm[fo + 4 * PS + 8 * (4 * hf + j)]);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; j++)
            *reinterpret_cast<double2*>(mm + r * PS + 8 * (4 * hf + j) + 2 * q) = make_double2(mt[j][0], mt[j][1]);
        }
      }
      __syncwarp();
      WS_TICK(CMPC_PH_ADAPT)
      // 3b. every tile (I, J) += C_I' M_J: both operand sets of a k-step are loaded up front (the main warps own 216
      //     registers), then 36 independent DMMAs issue back to back
      {
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
          if (ks == 1 && half_step) break;
          double pf[8], mf[8];
#pragma unroll
          for (int I = 0; I < 8; I++) {
            pf[I] = pan[fo + 4 * PS * ks + 8 * I];
            mf[I] = mm[fo + 4 * PS * ks + 8 * I];
          }
#pragma unroll
          for (int I = 0; I < 8; I++)
#pragma unroll
            for (int J = 0; J <= I; J++) dmma884(t[tix(I, J)][0], t[tix(I, J)][1], pf[I], mf[J]);
        }
        __syncwarp();  // every lane is done with pan and mm of this step
      }
      WS_TICK(CMPC_PH_SWEEP)
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
          pan[j] = -scale * a0;  // staged for lpt_file (the panel is free now)
          pan[j + 1] = -scale * a1;
        }
      }
    __syncwarp();
    __syncwarp();
    WS_TICK(CMPC_PH_X2)
    {
      int key = 0;
      if (P.lpt_hist) {  // rows violated at x0 (five per contact foot-step, one foot-step per lane: nc <= 21 here)
        int viol = 0;
        if (lane < nc) {
          const double fx = pan[3 * lane] * P.mu_inv, fy = pan[3 * lane + 1] * P.mu_inv, fz = pan[3 * lane + 2];
          const double tol = -P.tol_violation;
          viol = (fx + fz < tol) + (fz - fx < tol) + (fy + fz < tol) + (fz - fy < tol) + ((double)gv_mine * P.f_max - fz < tol);
        }
        key = min(__reduce_add_sync(0xffffffffu, viol), 63);
      }
      lpt_defer(inst_this, key);
    }
    __syncwarp();
    WS_TICK(CMPC_PH_LOAD)
  }
#undef WS_TICK
  lpt_flush();
  // release the helper
  if (lane == 0) ctrl[0] = -1;
  __syncwarp();
  named_bar_arrive(BAR_TILE);
  if (lane == 0 && P.flops && flops_acc) atomicAdd(P.flops + CMPC_K_INVERT, (unsigned long long)flops_acc);
}
