// Shared host/device definitions of the batched convex-MPC engine.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/cmpc_b200.h"

#define CMPC_MAX_FS 80 /* 4 * CMPC_MAX_HORIZON rounded up */

/* shared-memory canaries of a -DCMPC_CANARY build (cmpc_common.cuh): guard words behind every region of a carve */
#ifdef CMPC_CANARY
#define CMPC_CANARY_FIELDS int guard[32]; int nguard;
#define CMPC_GUARD_INIT(c) (c).nguard = 0
#define CMPC_GUARD(o, c) do { (o) = ((o) + 15) & ~15; (c).guard[(c).nguard++] = (o); (o) += 64; } while (0)
#else
#define CMPC_CANARY_FIELDS
#define CMPC_GUARD_INIT(c)
#define CMPC_GUARD(o, c)
#endif
#define CMPC_SM_SLOTS 1024 /* >= the largest %smid + 1 */
#define CMPC_RESUME_INTS 34 /* q, iterations, up to 64 working-set rows as 16-bit ids */
#define CMPC_QCAP_MID 56    /* middle capacity tier of the one-warp active-set kernel (reduced problems beyond 64 variables; "dual_team" = 0) */
#define CMPC_QCAP_TEAM 64   /* capacity of the CTA-per-instance active-set tier (cmpc_dual_team.cuh) */

// ---------------------------------------------------------------------------
// Instance record (HBM, one per MPC instance, 16-byte aligned, fetched with a
// single cp.async.bulk per instance).  Offsets in floats.  Field meaning is
// update_data_t's (convexMPC_interface.h:23-42).
// ---------------------------------------------------------------------------
#define CMPC_REC_P 0         // p[3]
#define CMPC_REC_V 3         // v[3]
#define CMPC_REC_Q 6         // q[4]  w,x,y,z
#define CMPC_REC_W 10        // w[3]
#define CMPC_REC_R 13        // r[12] r[axis*4+leg]
#define CMPC_REC_WEIGHTS 25  // weights[12]
#define CMPC_REC_ALPHA 37
#define CMPC_REC_XDRAG 38
#define CMPC_REC_FDIST 39    // xi[6]: torque xyz, force xyz (SolverMPC.cpp:795)
#define CMPC_REC_SIMTIME 45
#define CMPC_REC_RSV 46      // 2 floats reserved
#define CMPC_REC_TRAJ 48     // traj[12*h], then gait bytes [4*h], padded to 16 B

static inline int cmpc_rec_stride(int h) {
  int bytes = 4 * (CMPC_REC_TRAJ + 12 * h) + 4 * h;
  return (bytes + 15) & ~15;
}

// sigma tables: five h*h double tables of horizon sums, see DESIGN.md §3
//   sig[t][a*h+b] = sum_{r=max(a,b)}^{h-1} cx(r-a) * cy(r-b)
#define CMPC_SIG_11 0
#define CMPC_SIG_22 1
#define CMPC_SIG_33 2
#define CMPC_SIG_23 3
#define CMPC_SIG_12 4
#define CMPC_SIG_COUNT 5

// Gaussian kernels of the disturbance band-pass (SolverMPC.cpp:714-715: sigma 7 and 27,
// radius ceil(3 sigma)), stored back to back
#define CMPC_GK_R1 21
#define CMPC_GK_R2 81
#define CMPC_GK_TOTAL (2 * CMPC_GK_R1 + 1 + 2 * CMPC_GK_R2 + 1)
/* shared-memory scratch of the estimator stage (doubles): three 400-sample arrays, the twenty 20th roots of unity,
 * the folded band-pass taps */
#define CMPC_ADAPT_SCRATCH (3 * CMPC_ADAPT_WINDOW + 40 + 2 * CMPC_GK_R2 + 2)

struct CmpcParams {
  int horizon;
  int count;        // instances (or worklist entries) this launch covers
  int rec_stride;   // bytes
  int nmax;         // max reduced variable count in this launch (3 * contact foot-steps)
  int qcap;         // working-set capacity of this launch
  int max_iter;
  int adapt_mode;   // -1 off, 0 estimate only, 1 estimate and apply, 2 apply the stored estimate
  int inv_stagger;  // inversion kernel: start offset (SM cycles) between the CTAs that share an SM; 0 = none
  int inv_f32;       // inversion kernel: pivot blocks by an fp32 Gauss-Jordan chain + FP64 Newton steps on the tensor cores
  double inv_refine; // inversion kernel: a block step whose pivot-block inverse has an entry above this (the matrix is scaled to a
                     // diagonal below 1) gets one residual correction of its panel; < 0 = never
  double dt;        // (double)(float)dt
  double mu_inv;    // (double)(1.f/(float)mu)   SolverMPC.cpp:657
  double f_max;     // (double)(float)f_max
  double mass_inv;  // 1/(double)(float)m
  double inertia[3];
  double gravity;   // (double)(-9.8f)          SolverMPC.cpp:592
  double tol_violation;
  double tol_active;
  const unsigned char* records;
  const double* sigma;
  const int* worklist;        // optional indirection (tier-2 launch); NULL = identity
  const int* count_ptr;       // optional device-side instance count (tier-2 launch), capped by `count`
  int* overflow_list;         // instances whose working set outgrew qcap
  int* overflow_count;
  int* resume_out;            // [overflow entries][CMPC_RESUME_INTS]: the working set an overflowed instance had reached,
                              // written next to overflow_list by the tier that ran out of capacity
  const int* resume_in;       // the same records, read by the launch that takes the worklist over (null: restart from x0)
  // the iterate that goes with a resume record, so that the next tier continues without re-bordering the working set:
  // per overflow entry [x (nmax)] [u (q)] [P = (N'KN)^-1: q x q full rows from the fast tier, packed lower triangle from
  // the any-capacity tier]; entries beyond rstate_*_cap carry the row ids only
  double* rstate_out;
  const double* rstate_in;
  int rstate_out_stride, rstate_in_stride;  // doubles per entry
  int rstate_out_cap, rstate_in_cap;        // entries
  double* forces;             // [count][12h]
  double* objective;          // [count]
  int* status;                // [count]
  int* iterations;            // [count]
  signed char* active;        // [count][20h]
  unsigned long long* flops;  // CMPC_K_COUNT counters of algorithmic flops, one per kernel class
  // periodic-disturbance estimator (adapt_mode >= 0)
  const double* twiddle;      // [400][2] cos, -sin of 2 pi m / 400
  const float* gk;            // CMPC_GK_TOTAL normalised Gaussian taps
  const float* win_t;         // [count][400]
  const float* win_d;         // [count][400]
  const float* sim_time;      // [count]
  double* est;                // [count][4] stat, amp, freq, phase
  float* f_est;               // [count][6]
  // global-memory workspace tier (reduced problems too large for shared memory): K and P per CTA
  double* gws;                // NULL unless shape == CMPC_SHAPE_GMEM
  size_t gws_stride;          // doubles per CTA
  // two-kernel pipeline (cmpc_pipeline.cu): per-instance workspace slots written by the condensation kernel
  // and read by the dual active-set kernel, and the device-side work counter of the launch
  double* qws;                // [count][qws_stride]: K (nmax x nmax, row stride n), g, x0, header
  size_t qws_stride;          // doubles per slot, >= cmpc_qws_slot_doubles(nmax)
  int* sched;                 // work counter of THIS launch (zeroed by the host)
  int sweep_dmma;             // condensation kernel of the 96 / 128 shapes: sweep on the FP64 tensor cores (else DFMA register tiles)
  int k_tiled;                // K is stored as 36 lower-triangular 8x8 tiles (cmpc_condense_mma.cuh), else row-major n x n
  int qws_goff;               // offset (doubles) of g in a slot; x0 follows at +nmax, the header at +2 nmax
  // hardest-first order of the active-set kernel: the inversion kernel files every instance under the number of
  // constraint rows its unconstrained optimum violates (a good predictor of the active-set iterations)
  int* lpt_hist;              // [64] instances per key, zeroed per launch; null = natural order
  int* lpt_key;               // [count] key << 24 | position within the key's bucket
  int* sm_slots;              // [CMPC_SM_SLOTS] arrival counters per SM (%smid), zeroed per launch; null = no stagger
  // optional phase clocks (profiling aid): CMPC_PH_COUNT counters of SM cycles summed over CTAs, thread 0 only
  unsigned long long* phase_cycles;
};

/* kernel classes (flop counters, per-kernel timing) */
#define CMPC_K_ASSEMBLE 0 /* condensation: state, c2qp closed form, g, H (plus the DFMA sweep on the 64 < n <= 128 shapes) */
#define CMPC_K_INVERT 1   /* K = H^-1, x0 = -K g on the FP64 tensor cores (n <= 63) */
#define CMPC_K_DUAL 2     /* dual active-set iterations, outputs */
#define CMPC_K_FUSED 3    /* single-kernel path (n > 128, or CMPC_PATH=fused) */
#define CMPC_K_COUNT 4

#define CMPC_PH_WAIT 0    /* record wait (mbarrier) */
#define CMPC_PH_ADAPT 1   /* disturbance estimator */
#define CMPC_PH_PREP 2    /* state, W, tracking error, aggregates, g */
#define CMPC_PH_HESS 3    /* H assembly */
#define CMPC_PH_LOAD 4    /* tile load, scaling, diagonal copy */
#define CMPC_PH_SWEEP 5   /* pivot loop */
#define CMPC_PH_STORE 6   /* K store, x0 = -K g, slacks */
#define CMPC_PH_QP 7      /* active-set iterations */
#define CMPC_PH_OUT 8     /* objective, outputs */
#define CMPC_PH_PUBLISH 9 /* warp-specialised inversion: panel publish */
#define CMPC_PH_DVWAIT 10 /* warp-specialised inversion: main warp waiting for the helper's pivot-block inverse */
#define CMPC_PH_X0 11 /* diagnostics of the inversion kernel's load / store bucket */
#define CMPC_PH_X1 12
#define CMPC_PH_X2 13
#define CMPC_PH_X3 14
#define CMPC_PH_COUNT 15

// kernel shapes (cmpc_kernels.cu): register-tile tiers by reduced problem size, shared-memory tier beyond
#define CMPC_SHAPE_64 0    /* n <= 64, 64 threads, 8x8 register tiles */
#define CMPC_SHAPE_64W 1   /* n <= 64, 128 threads, 8x4 register tiles */
#define CMPC_SHAPE_128 2   /* n <= 128, 256 threads, 8x8 register tiles */
#define CMPC_SHAPE_MEM 3   /* matrix in shared memory, 128 threads */
#define CMPC_SHAPE_GMEM 4  /* matrix and working-set inverse in an L2-resident global workspace, 128 threads */

// two-kernel pipeline shapes (cmpc_condense.cuh)
#define CMPC_CSHAPE_64 0   /* n <= 64, 128 threads */
#define CMPC_CSHAPE_96 1   /* n <= 96, 256 threads */
#define CMPC_CSHAPE_128 2  /* n <= 128, 256 threads */
#define CMPC_CSHAPE_MMA64 3 /* n <= 63: assembly kernel (CTA per instance) + DMMA inversion kernel (warp per instance) */
#define CMPC_PIPELINE_NMAX 128

#define CMPC_KTILE_DOUBLES (36 * 64)
static inline int cmpc_qws_goff(int nmax, int tiled) {
  int k = tiled ? CMPC_KTILE_DOUBLES : nmax * nmax;
  return (k + 1) & ~1;
}
static inline size_t cmpc_qws_slot_doubles(int nmax, int tiled) {
  // K, g, x0, scale (+ pad), header {int nc, int status, uchar fs[CMPC_MAX_FS], uchar gait[CMPC_MAX_FS]}
  size_t d = (size_t)cmpc_qws_goff(nmax, tiled) + 2 * (size_t)nmax + 2 + (8 + 2 * CMPC_MAX_FS + 7) / 8;
  return (d + 1) & ~(size_t)1;
}

size_t cmpc_condense_smem_bytes(int horizon, int nmax, int cshape, bool adapt);
int cmpc_condense_max_ctas_per_sm(int cshape, size_t smem, bool adapt);
int cmpc_launch_condense(const CmpcParams& P, int cshape, int grid, void* stream);
int cmpc_condense_instances_per_cta(int cshape);
int cmpc_invert_max_ctas_per_sm(void);
int cmpc_invert_instances_per_cta(void);
int cmpc_launch_invert(const CmpcParams& P, int grid, void* stream);
int cmpc_launch_lpt_order(const int* hist, const int* key, int* worklist, int count, void* stream);
size_t cmpc_dual_fast_smem_bytes(int nmax, int qcap);  /* per CTA of one warp; qcap <= 32, nmax <= 128 */
int cmpc_dual_fast_max_ctas_per_sm(int nmax, size_t smem);
int cmpc_launch_dual_fast(const CmpcParams& P, int grid, void* stream);
size_t cmpc_dual_smem_bytes_per_warp(int nmax, int qcap);
int cmpc_dual_max_ctas_per_sm(int warps_per_cta, size_t smem);
int cmpc_launch_dual(const CmpcParams& P, int warps_per_cta, int grid, void* stream);
size_t cmpc_dual_team_smem_bytes(int nmax, int qcap);
int cmpc_dual_team_max_ctas_per_sm(int nmax, size_t smem);
int cmpc_launch_dual_team(const CmpcParams& P, int grid, void* stream);

size_t cmpc_smem_bytes(int horizon, int nmax, int qcap, int shape, bool adapt);
int cmpc_shape_threads(int shape);
int cmpc_launch_solve(const CmpcParams& P, int shape, int grid, void* stream);
int cmpc_max_ctas_per_sm(int shape, size_t smem, bool adapt);
int cmpc_run_dfma_peak(int sm_count, void* stream, double* out_dev, int iters);
int cmpc_run_dmma_peak(int sm_count, void* stream, double* out_dev, int iters);
/* cmpc_frontend.cu: the caller of the path (updateMPCIfNeeded / solveDenseMPC / getMpcTable) on the device */
int cmpc_launch_frontend(const void* cmds, unsigned char* records, void* results, float* f_ext, float* sim_time, int count,
                         int horizon, int rec_stride, float dt, float alpha, const float weights[12], void* stream);
int cmpc_launch_history_push(const void* cmds, const float* f_ext, float* win_t, float* win_d, int count, int have,
                             void* stream);
int cmpc_launch_epilogue(const void* cmds, const double* forces, const int* status, const int* iterations, void* results,
                         int count, int horizon, void* stream);
/* cmpc_pack.cu: instance records from structure-of-arrays inputs (device-accessible pointers, e.g. pinned host memory) */
int cmpc_launch_pack(const void* p, const void* v, const void* q, const void* w, const void* r, const void* weights,
                     const void* traj, const void* alpha, const void* gait, const void* x_drag, const void* f_dist,
                     unsigned char* records, int rec_stride, int horizon, int count, int sm_count, void* stream,
                     int which /* CMPC_PACK_ALL, or the two halves of a split call: _REST (all but traj), _TRAJ (traj only) */);
enum { CMPC_PACK_ALL = 0, CMPC_PACK_REST = 1, CMPC_PACK_TRAJ = 2 };
