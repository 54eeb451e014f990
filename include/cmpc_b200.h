/* cmpc_b200 — C-ABI of the B200-native batched convex-MPC engine.
 *
 * Drop-in boundary for the reference's dense solve_mpc path
 * (be2r_cmpc_unitree/src/controllers/convexMPC).  Two groups of entry points:
 *
 *  1. The reference's own C interface, symbol for symbol
 *     (convexMPC_interface.h:44-52).  A controller that links this library
 *     instead of convexMPC_interface.cpp + SolverMPC.cpp keeps compiling and
 *     gets one MPC instance solved on the GPU per update_problem_data* call.
 *  2. cmpc_batch_*: the batched engine those wrappers sit on — many independent
 *     MPC instances (robot state, gait contact schedule, disturbance estimate)
 *     solved by one kernel launch.
 *
 * Plain pointers and sizes only; no CUDA or torch types.  Every function is
 * host-callable, returns 0 on success or a negative CMPC_E_* code, and fails
 * loudly (no CPU fallback) when no CUDA device is usable.
 */
#ifndef CMPC_B200_H
#define CMPC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMPC_MAX_HORIZON 19 /* SolverMPC.cpp:113 "horizon is too long" */
#define CMPC_ADAPT_WINDOW 400 /* SolverMPC.cpp:704 */

enum {
  CMPC_OK = 0,
  CMPC_E_ARG = -1,     /* bad argument (horizon out of range, null pointer, count > capacity) */
  CMPC_E_CUDA = -2,    /* CUDA runtime error; cmpc_last_error() has the text */
  CMPC_E_NODEVICE = -3,/* no CUDA device / kernel image not loadable: there is no CPU path */
  CMPC_E_STATE = -4    /* call order violated (solve before setup/upload) */
};

/* per-instance solver status written by the kernel */
enum {
  CMPC_ST_SOLVED = 0,     /* optimum found */
  CMPC_ST_EMPTY = 1,      /* no foot in contact over the horizon: forces are all zero */
  CMPC_ST_MAXITER = 2,    /* active-set iteration cap hit */
  CMPC_ST_INFEASIBLE = 3, /* cannot happen for this constraint set; kept for diagnostics */
  CMPC_ST_WSOVERFLOW = 4, /* working set outgrew the launch's capacity tier; the engine re-solves
                             these with the full-capacity kernel before returning */
  CMPC_ST_CAPACITY = 5,   /* instance has more contact foot-steps than the launch was sized for
                             (only reachable through cmpc_batch_set_count with a wrong bound) */
  CMPC_ST_NONFINITE = 6   /* NaN / Inf in the inputs, the disturbance estimate or the iterate: the forces of this
                             instance are all zero and must not be used (qpOASES would return garbage here) */
};

/* ------------------------------------------------------------------------
 * 1. reference interface (convexMPC_interface.h)
 * ---------------------------------------------------------------------- */
/* convexMPC_interface.h:44 / convexMPC_interface.cpp:44 */
void setup_problem(double dt, int horizon, double mu, double f_max);
/* convexMPC_interface.h:45 / convexMPC_interface.cpp:89 — solves on return */
void update_problem_data(double* p, double* v, double* q, double* w, double* r, double yaw, double* weights,
                         double* state_trajectory, double alpha, int* gait);
/* convexMPC_interface.h:48 / convexMPC_interface.cpp:132 — solves on return */
void update_problem_data_floats(float* p, float* v, float* q, float* w, float* r, float roll, float pitch,
                                float yaw, float* weights, float* state_trajectory, float alpha, int* gait);
/* convexMPC_interface.h:46 / convexMPC_interface.cpp:156 */
double get_solution(int index);
/* convexMPC_interface.h:47 / convexMPC_interface.cpp:109.  The JCQP (ADMM)
 * settings are accepted and ignored: the GPU path always returns the exact
 * active-set optimum that the use_jcqp == 0 (qpOASES) branch computes. */
void update_solver_settings(int max_iter, double rho, double sigma, double solver_alpha, double terminate,
                            double use_jcqp);
/* convexMPC_interface.h:52 / convexMPC_interface.cpp:151.  The reference declares this one WITHOUT extern "C", so an
 * unmodified reference translation unit looks for the C++ symbol _Z13update_x_dragf: the library exports both. */
void update_x_drag(float x_drag);
/* The reference hands the adaptive hook its inputs through two globals,
 * `Eigen::Matrix<float,6,1> f_ext` (convexMPC_interface.h:54, written at
 * ConvexMPCLocomotion.cpp:771) and `float simulation_time`
 * (be2r_cmpc_unitree.hpp:156).  A C-ABI cannot export an Eigen object, so the
 * controller calls these two setters where it used to assign the globals. */
void cmpc_set_external_force(const float f_ext[6]);
void cmpc_set_simulation_time(float t);
/* f_est after the last solve (SolverMPC.h:76 `extern f_est`) */
void cmpc_get_disturbance_estimate(float f_est[6]);
/* the two filtered estimates solve_mpc keeps next to it (SolverMPC.h:73-74):
 * f_est_smoothed = 0.95 f_est_smoothed + 0.05 f_est (SolverMPC.cpp:783), f_est_static[3] = 0.97 f_est_static[3] +
 * 0.03 f_ext[3] (SolverMPC.cpp:798), both updated once per solve */
void cmpc_get_disturbance_estimate_smoothed(float f_est_smoothed[6]);
void cmpc_get_disturbance_estimate_static(float f_est_static[6]);
/* What the reference interface does when a solve does not end in CMPC_ST_SOLVED / CMPC_ST_EMPTY or returns non-finite
 * forces (the reference prints "failed to solve!" and hands the controller whatever qpOASES left in memory):
 * CMPC_ON_ERROR_ABORT (default) prints the reason and aborts the process — a controller must not run on garbage;
 * CMPC_ON_ERROR_HOLD keeps the forces of the last good solve for get_solution() and records the status. */
enum { CMPC_ON_ERROR_ABORT = 0, CMPC_ON_ERROR_HOLD = 1 };
void cmpc_set_error_policy(int policy);
/* CMPC_ST_* of the last update_problem_data* call */
int cmpc_last_status(void);
/* Forget the accumulated f_ext / time history (the reference's file-scope vectors, SolverMPC.cpp:398-399,
 * live for the life of the process; a test or a controller restart needs a way to clear them). */
void cmpc_reset_history(void);

/* ------------------------------------------------------------------------
 * 2. batched engine
 * ---------------------------------------------------------------------- */
typedef struct cmpc_batch cmpc_batch;

/* Host-side view of one batch of inputs, structure-of-arrays, instance-major.
 * Field meaning is update_data_t's (convexMPC_interface.h:23-42). */
typedef struct {
  const float* p;        /* [count][3]  */
  const float* v;        /* [count][3]  */
  const float* q;        /* [count][4]  w,x,y,z */
  const float* w;        /* [count][3]  */
  const float* r;        /* [count][12] r[axis*4+leg] */
  const float* weights;  /* [count][12] */
  const float* traj;     /* [count][12*horizon] */
  const float* alpha;    /* [count] */
  const uint8_t* gait;   /* [count][4*horizon] gait[step*4+leg] */
  const float* x_drag;   /* [count] */
  const float* f_dist;   /* [count][6] disturbance estimate xi applied as Q_qp*xi in g
                            (SolverMPC.cpp:810); NULL = zeros (SolverMPC.cpp:813) */
} cmpc_inputs;

typedef struct {
  double* forces;     /* [count][12*horizon] q_soln (zeros for swing feet); may be NULL */
  double* objective;  /* [count] 0.5 x'Hx + g'x of the reduced QP; may be NULL */
  int32_t* status;    /* [count] CMPC_ST_*; may be NULL */
  int32_t* iterations;/* [count] working-set changes (qpOASES' nWSR analogue); may be NULL */
  int8_t* active;     /* [count][20*horizon] primal activity of fmat rows: -1 lower, 0, +1 upper; may be NULL */
} cmpc_outputs;

int cmpc_device_count(void);
const char* cmpc_last_error(void);

int cmpc_batch_create(cmpc_batch** out, int device, int capacity);
void cmpc_batch_destroy(cmpc_batch* b);
/* setup_problem for the whole batch.  mass/inertia NULL-able: defaults are the
 * reference's hard-coded 12 kg and diag(.07,.26,.242) (RobotState.h:24, RobotState.cpp:49). */
int cmpc_batch_setup(cmpc_batch* b, double dt, int horizon, double mu, double f_max);
int cmpc_batch_set_robot(cmpc_batch* b, double mass, const double inertia_diag[3]);
/* Diagnostic / test switches (which kernel path, capacity tiers, stream count ...; the keys are listed next to
 * cmpc_batch_set_option in csrc/cmpc_api.cu).  The library reads NO environment variable on any solve path: a
 * switch exists only through this call.  Drains the batch and drops its cached launch plans.
 * Two of them concern accuracy rather than speed: "inv_refine" (default 1024; -1 never, 0 always) is the size of a
 * pivot-block inverse above which the blocked sweeps correct their panel once — what keeps instances with a badly
 * conditioned Hessian (alpha ~1e-6, unweighted states) within 1e-6 N of qpOASES; "inv_f32" (default 1) lets the
 * n <= 63 inversion kernel seed its pivot-block inverses in fp32 and finish them with Newton steps in FP64 (0: FP64
 * chain).  "submit_copy" (default 0) hands the results of cmpc_batch_submit_bound to the copy engine: faster from four
 * batches in flight on, slower below. */
int cmpc_batch_set_option(cmpc_batch* b, const char* key, int value);
/* Pack `count` host instances into pinned records and copy them to the device (async on the batch stream). */
int cmpc_batch_upload(cmpc_batch* b, int count, const cmpc_inputs* in);
/* Launch the fused condensation + QP kernel over the uploaded instances (async). */
int cmpc_batch_solve(cmpc_batch* b);
/* Same, restricted to the uploaded instances [first, first+count): lets a caller keep several
 * batches resident in one buffer and solve them one after another (bench ring, multi-GPU shards). */
int cmpc_batch_solve_range(cmpc_batch* b, int first, int count);
/* Copy results back and wait. */
int cmpc_batch_download(cmpc_batch* b, const cmpc_outputs* out);
/* upload + solve + download: the end-to-end call with host buffers. */
int cmpc_batch_solve_host(cmpc_batch* b, int count, const cmpc_inputs* in, const cmpc_outputs* out);
/* The same call for a caller that reuses its arrays (a controller loop): bind them once — which arrays are pinned
 * is looked up here, not on every call — then solve `count` instances from / into the bound arrays.  With pinned
 * inputs the device reads the arrays itself and packs the records; outputs are written by the kernels straight into
 * host memory (the bound arrays if pinned, pinned staging otherwise).  The arrays must stay allocated and keep their
 * pinning while bound; bind (NULL, NULL) to drop the binding. */
int cmpc_batch_bind_host(cmpc_batch* b, const cmpc_inputs* in, const cmpc_outputs* out);
int cmpc_batch_solve_bound(cmpc_batch* b, int count);
/* The same in two halves: submit enqueues everything and returns, wait blocks until the results are in the bound
 * output arrays.  A caller with two batches (two cmpc_batch objects, two sets of arrays) submits batch k+1 before it
 * waits for batch k, so the PCIe traffic and kernel tails of one overlap the kernels of the other. */
int cmpc_batch_submit_bound(cmpc_batch* b, int count);
int cmpc_batch_wait_bound(cmpc_batch* b);
int cmpc_batch_sync(cmpc_batch* b);
/* Pin a caller-owned host array (cudaHostRegister): cmpc_batch_solve_host / cmpc_batch_download copy results
 * straight into pinned output arrays instead of staging them.  Unregister before freeing the array. */
int cmpc_host_register(void* ptr, size_t bytes);
int cmpc_host_unregister(void* ptr);

/* ------------------------------------------------------------------------
 * 2b. the caller of the path on the device: ConvexMPCLocomotion::updateMPCIfNeeded + solveDenseMPC
 *     (ConvexMPCLocomotion.cpp:511-870) and the gait's getMpcTable (Gait.cpp:158-215) for a whole batch.
 *     One cmpc_command per robot carries what those functions read from the state estimator, the command,
 *     the gait object and the /log_data message; the device builds the reference trajectory, the contact
 *     table, r = pFoot - p, the f_ext residual and the x_drag integral, pushes (simulation_time, f_ext[3])
 *     into the instance's disturbance history (SolverMPC.cpp:688-700), solves, and returns what
 *     solveDenseMPC leaves behind: Fr_des, f_ff = -rBody f, and the updated command state.
 * ---------------------------------------------------------------------- */
enum { CMPC_GAIT_OFFSET_DURATION = 0, CMPC_GAIT_MIXED_FREQUENCY = 1 };

typedef struct {
  /* StateEstimate<float> (stateEstimator->getResult()) */
  float position[3];
  float ground_z;          /* ground_truth_position[2], the z that solveDenseMPC hands the solver (:640) */
  float v_world[3];
  float omega_world[3];
  float orientation[4];    /* w,x,y,z */
  float rpy[3];
  float r_body[9];         /* rBody, row major */
  float p_foot[12];        /* pFoot[leg][axis], world frame */
  /* command state of ConvexMPCLocomotion */
  float x_vel_des, y_vel_des, yaw_turn_rate, yaw_des, body_height;
  float rpy_comp[2];
  float world_position_desired[2];
  float roll_des, pitch_des;
  float stand_traj[3];     /* stand_traj[0], [1], [5] (:526) */
  float x_comp_integral;
  float cmpc_x_drag;       /* _dyn_params->cmpc_x_drag */
  /* gait object (Gait.h): nIterations == horizon */
  int32_t gait_kind;       /* CMPC_GAIT_* */
  int32_t gait_iteration;  /* _iteration */
  int32_t gait_offsets[4]; /* OffsetDurationGait::_offsets; MixedFrequncyGait::_periods */
  int32_t gait_durations[4];
  float gait_duty;         /* MixedFrequncyGait::_duty_cycle */
  int32_t omni_mode;
  int32_t stand;           /* current_gait == 4 (:524) */
  int32_t have_log;        /* received_log_data_ (:647) */
  /* /log_data of the previous step (:650-765) */
  float log_x_prev[12];    /* euler_act xyz, pos_act xyz, vel_act.angular xyz, vel_act.linear xyz */
  float log_R[9];          /* R_00 .. R_22, row major */
  float log_r_feet[12];    /* r_x_1..4, r_y_1..4, r_z_1..4 */
  float log_foot_force[12];/* foot_force{0..3}_{x,y,z} */
  float log_x_drag;
  float sim_time;          /* simulation_time (be2r_cmpc_unitree.hpp:156) */
  float pad_;
} cmpc_command;            /* 464 bytes */

typedef struct {
  float fr_des[12];                 /* Fr_des[leg][axis] = first-step ground reaction forces (:836-845) */
  float f_ff[12];                   /* f_ff[leg] = -rBody * f (:841) */
  float world_position_desired[2];  /* after the max_pos_error clamp (:535-548) */
  float x_comp_integral;            /* after the update at :811-816 */
  float f_ext[6];                   /* the residual handed to the solver (:771); unchanged without /log_data */
  int32_t status;                   /* CMPC_ST_* */
  int32_t iterations;
  float pad_[1];
} cmpc_command_result;              /* 144 bytes */

/* Q and alpha of solveDenseMPC (:627, :634); defaults are the reference's hard-coded values. */
int cmpc_batch_set_weights(cmpc_batch* b, const float weights[12], float alpha);
/* One MPC update of `count` robots from host command structs (any host memory; pinned is faster).
 * forces_out may be NULL; when given it receives the full q_soln [count][12*horizon] as well.
 * Instance slot i is robot i for as long as the disturbance histories live: the sample count is kept per batch (the
 * reference keeps one history per process), so every call between two cmpc_batch_reset_history calls must pass the
 * same `count` (CMPC_E_STATE otherwise). */
int cmpc_batch_solve_commands(cmpc_batch* b, int count, const cmpc_command* commands, cmpc_command_result* results,
                              double* forces_out);
/* Forget the per-instance disturbance histories, estimates and f_ext (the reference's file-scope vectors). */
int cmpc_batch_reset_history(cmpc_batch* b);
/* samples pushed into the histories so far (the reference's time_history.size()) */
int cmpc_batch_history_length(cmpc_batch* b, int* samples);
/* Test / debug helper: copy the instance records [first, first+count) as the device holds them. */
int cmpc_batch_copy_records(cmpc_batch* b, int first, int count, void* dst);

/* Adaptive-MPC periodic disturbance estimation fused into the solve launch
 * (SolverMPC.cpp:688-798).  windows_t/windows_d hold, per instance, the last
 * CMPC_ADAPT_WINDOW samples of time and f_ext[3]; sim_time is the time the
 * compensating force is evaluated at.  mode 0: estimate only (history 400..500
 * samples, g sees no disturbance); mode 1: estimate and apply xi in g in the
 * same launch; mode 2: no new fit, refresh f_est[3] from the stored fit at sim_time and apply it
 * (history beyond 500 samples; windows may be NULL).  mode < 0 switches the stage off again. */
int cmpc_batch_upload_disturbance(cmpc_batch* b, int count, const float* windows_t, const float* windows_d,
                                  const float* sim_time, int mode);
/* est[count][4] = stat, amp, freq(Hz), phase; f_est[count][6] */
int cmpc_batch_download_disturbance(cmpc_batch* b, double* est, float* f_est);

/* Device-resident access for callers that generate or keep inputs in HBM
 * (bench `value` path, multi-GPU shards): record layout is documented in
 * DESIGN.md; pointers are CUDA device pointers carried as void*. */
int cmpc_batch_device_records(cmpc_batch* b, void** records, size_t* stride_bytes);
int cmpc_batch_device_forces(cmpc_batch* b, void** forces);
/* Re-use the last uploaded instances: mark `count` records as present without copying. */
int cmpc_batch_set_count(cmpc_batch* b, int count, int max_contact_feet);
/* Timing helpers (CUDA events on the batch stream): ms of the last solve launch(es). */
int cmpc_batch_last_solve_ms(cmpc_batch* b, float* ms);
int cmpc_batch_kernel_launches(cmpc_batch* b, long long* launches);
/* Region timing: record event 0 / 1 on the batch stream, then read the elapsed device time. */
int cmpc_batch_mark(cmpc_batch* b, int which);
int cmpc_batch_marked_ms(cmpc_batch* b, float* ms);
/* Zero the launch and flop counters. */
int cmpc_batch_reset_counters(cmpc_batch* b);
/* Algorithmic FP64 flop count since the last reset, accumulated by the kernel from its own loop counters. */
int cmpc_batch_last_flops(cmpc_batch* b, double* flops);

/* Per kernel class (0 assembly / condensation, 1 FP64-tensor inversion, 2 dual active set, 3 fused single-kernel
 * path): algorithmic flops since the last reset, and the device time of ONE solve of [first, first+count) on the
 * batch stream with a CUDA event between the classes (ms[i] stays 0 for a class the range does not use). */
int cmpc_batch_kernel_flops(cmpc_batch* b, double flops[4]);
int cmpc_batch_profile_range(cmpc_batch* b, int first, int count, float ms[4]);

/* Phase clocks (profiling aid): when enabled, thread 0 of every CTA charges SM cycles to the kernel's
 * phases (record wait, estimator, condensation, H assembly, tile load, sweep, K store, active set,
 * outputs); cycles[i] is the sum over CTAs since enabling.  Off by default. */
int cmpc_batch_enable_phase_clocks(cmpc_batch* b, int on);
int cmpc_batch_phase_cycles(cmpc_batch* b, unsigned long long* cycles, int n);

/* Measured FP64 FMA throughput of the device (dependent-free DFMA chains on every SM), the
 * denominator of the solve kernel's roofline; TFLOP/s. */
int cmpc_measure_fp64_peak(int device, double* tflops);
/* The same for the FP64 tensor pipe (independent DMMA m8n8k4 accumulators on every SM sub-partition): the
 * denominator of the inversion kernel's roofline; TFLOP/s. */
int cmpc_measure_dmma_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
