/* ORACLE — TEST INFRASTRUCTURE ONLY (see cmpc_oracle.cpp header). */
#ifndef CMPC_ORACLE_H
#define CMPC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMPC_ORACLE_MAX_SEG 36      /* K_MAX_GAIT_SEGMENTS, convexMPC_interface.h:3 */
#define CMPC_ORACLE_HIST_MAX 4096

/* problem_setup (convexMPC_interface.h:15-21) plus the robot constants the
 * reference hard-codes in RobotState.h:24 and RobotState.cpp:49 */
typedef struct {
  float dt, mu, f_max;
  int horizon;
  float mass;
  float inertia[3];
  int nwsr; /* 0 -> 100, SolverMPC.cpp:854 */
} cmpc_oracle_setup;

/* update_data_t (convexMPC_interface.h:23-42), solver-relevant fields */
typedef struct {
  float p[3], v[3], q[4], w[3], r[12];
  float roll, pitch, yaw;
  float weights[12];
  float traj[12 * CMPC_ORACLE_MAX_SEG];
  float alpha;
  unsigned char gait[4 * CMPC_ORACLE_MAX_SEG];
  float x_drag;
} cmpc_oracle_update;

typedef struct {
  /* required, caller-allocated */
  double* x;          /* 12h  : q_soln, zeros for swing feet */
  double* y_con;      /* 20h  : qpOASES constraint multipliers (0 where eliminated) */
  int8_t* con_status; /* 20h  : qpOASES working set: -1 lower, 0 inactive, +1 upper */
  int8_t* var_elim;   /* 12h  : 1 where the variable was eliminated */
  /* optional, may be NULL */
  double* H_full;     /* (12h)^2 */
  double* g_full;     /* 12h */
  double* H_red;      /* n_var^2 (allocate (12h)^2) */
  double* g_red;      /* n_var */
  double* AdtBdtQdt;  /* 13 x 31 */
  /* outputs */
  int n_var, n_con, nwsr, qp_return, qp_status_ok;
  double objective;
} cmpc_oracle_result;

typedef struct {
  int len;
  float t_hist[CMPC_ORACLE_HIST_MAX];
  float d_hist[CMPC_ORACLE_HIST_MAX];
  double est[4]; /* stat, amp, freq, phase */
  float f_est[6], f_est_smoothed[6], f_est_static3;
} cmpc_oracle_adapt;

int cmpc_oracle_solve(const cmpc_oracle_setup* st, const cmpc_oracle_update* up, const double* f_dist,
                      int use_float, cmpc_oracle_result* res);
int cmpc_oracle_fit_window(const double* t, const double* d, int n, double out[4]);
int cmpc_oracle_adapt_step(cmpc_oracle_adapt* a, double sim_time, double f_ext3, double f_dist_out[6]);
int cmpc_oracle_solve_batch(const cmpc_oracle_setup* st, const cmpc_oracle_update* ups, int count,
                            int use_float, int threads, double* forces_out, int* ok_out);

#ifdef __cplusplus
}
#endif
#endif
