// The caller of the solve path, on the device (row (f) of DESIGN.md §0): ConvexMPCLocomotion::updateMPCIfNeeded and
// solveDenseMPC (ConvexMPCLocomotion.cpp:511-870) plus the gait's getMpcTable (Gait.cpp:158-215) for a batch of
// robots.  Three small kernels around the solve pipeline:
//
//   cmpc_frontend_kernel      one thread per robot: reference trajectory, contact table, r = pFoot - p, the f_ext
//                             residual of the previous step's model, the x_drag integral -> the instance record
//   cmpc_history_push_kernel  one CTA per robot: (simulation_time, f_ext[3]) appended to the 400-sample window the
//                             disturbance estimator fits (SolverMPC.cpp:688-706)
//   cmpc_epilogue_kernel      one thread per robot: Fr_des, f_ff = -rBody f from the first horizon step
//
// The reference computes all of this in fp32; so do these kernels, with explicitly rounded operations
// (__fmul_rn / __fadd_rn, no FMA contraction) in a stated order — products left to right, the third term added
// last — so that the records they write can be checked bit for bit by the tests' CPU restatement.
#include <cuda_runtime.h>
#include <stdint.h>

#include "cmpc_device.h"

namespace {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
  return add(add(mul(a0, b0), mul(a1, b1)), mul(a2, b2));
}

struct FrontArgs {
  const cmpc_command* cmds;
  unsigned char* records;
  cmpc_command_result* results;
  float* f_ext;  // [count][6], persistent per instance (the reference's global f_ext)
  float* sim_time;  // [count] or null: simulation_time for the estimator stage
  int count, horizon, rec_stride;
  float dt;      // dtMPC
  float alpha;
  float weights[12];
};

constexpr int FRONT_NT = 64;  // robots per CTA: small CTAs, so that a few thousand robots still cover every SM

__global__ void __launch_bounds__(FRONT_NT) cmpc_frontend_kernel(const __grid_constant__ FrontArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = A.horizon;
  const float dt = A.dt;

  // ---- updateMPCIfNeeded (:511-594): reference trajectory.  The thread of a robot works out the twelve constant
  //      entries and the three walks (yaw, x, y); the CTA then fills the 12 h floats of its robots' records together,
  //      consecutive threads on consecutive words, every entry re-running its own float accumulation (:579-581) ----
  __shared__ float s_t0[FRONT_NT][12];
  __shared__ float s_walk[FRONT_NT][3];   // dt * yaw_turn_rate, dt * v_des_world[0], dt * v_des_world[1]; 0 for a stand
  __shared__ int s_gait[FRONT_NT][11];    // kind, iteration, offsets[4], durations[4], stand
  __shared__ float s_duty[FRONT_NT];
  const int tl = threadIdx.x;
  float wpd0 = 0.f, wpd1 = 0.f;
  if (i < A.count) {
    const cmpc_command& c = A.cmds[i];
    wpd0 = c.world_position_desired[0];
    wpd1 = c.world_position_desired[1];
    if (c.stand) {
      const float t0[12] = {c.roll_des, c.pitch_des, c.stand_traj[2], c.stand_traj[0], c.stand_traj[1], c.body_height,
                            0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < 12; j++) s_t0[tl][j] = t0[j];
      s_walk[tl][0] = s_walk[tl][1] = s_walk[tl][2] = 0.f;   // the stand trajectory is constant (:524-531)
    } else {
      float vw0 = c.x_vel_des, vw1 = c.y_vel_des;
      if (!c.omni_mode) {  // rBody^T * (x_vel_des, y_vel_des, 0)
        const float* R = c.r_body;
        vw0 = dot3(R[0], c.x_vel_des, R[3], c.y_vel_des, R[6], 0.f);
        vw1 = dot3(R[1], c.x_vel_des, R[4], c.y_vel_des, R[7], 0.f);
      }
      const float max_pos_error = .1f;
      const float p0 = c.position[0], p1 = c.position[1];
      float xs = wpd0, ys = wpd1;
      if (sub(xs, p0) > max_pos_error) xs = add(p0, max_pos_error);
      if (sub(p0, xs) > max_pos_error) xs = sub(p0, max_pos_error);
      if (sub(ys, p1) > max_pos_error) ys = add(p1, max_pos_error);
      if (sub(p1, ys) > max_pos_error) ys = sub(p1, max_pos_error);
      wpd0 = xs;
      wpd1 = ys;
      // i == 0: "start at current position": only the yaw does (:575)
      const float t0[12] = {c.rpy_comp[0], c.rpy_comp[1], c.rpy[2], xs, ys, c.body_height,
                            0.f, 0.f, c.yaw_turn_rate, vw0, vw1, 0.f};
      for (int j = 0; j < 12; j++) s_t0[tl][j] = t0[j];
      s_walk[tl][0] = mul(dt, c.yaw_turn_rate);
      s_walk[tl][1] = mul(dt, vw0);
      s_walk[tl][2] = mul(dt, vw1);
    }
    s_gait[tl][0] = c.gait_kind;
    s_gait[tl][1] = c.gait_iteration;
    for (int j = 0; j < 4; j++) { s_gait[tl][2 + j] = c.gait_offsets[j]; s_gait[tl][6 + j] = c.gait_durations[j]; }
    s_duty[tl] = c.gait_duty;
    s_gait[tl][10] = c.stand;
  }
  __syncthreads();
  {
    const int base = blockIdx.x * FRONT_NT;
    const int nrob = min(FRONT_NT, A.count - base);
    const int tw = 12 * h;                                         // trajectory words per record
    const int gw = (A.rec_stride - 4 * (CMPC_REC_TRAJ + 12 * h)) / 4;  // gait words per record (4 h bytes + padding)
    for (int e = threadIdx.x; e < nrob * tw; e += FRONT_NT) {
      const int rb = e / tw, w = e - rb * tw, k = w / 12, j = w - 12 * k;
      float v = s_t0[rb][j];
      if (j >= 2 && j <= 4 && !s_gait[rb][10]) {
        const float d = s_walk[rb][j - 2];
        for (int q = 0; q < k; q++) v = add(v, d);
      }
      reinterpret_cast<float*>(A.records + (size_t)(base + rb) * A.rec_stride)[CMPC_REC_TRAJ + w] = v;
    }
    // ---- getMpcTable (Gait.cpp:158-215), nIterations == horizon: four legs of a step per word ----
    for (int e = threadIdx.x; e < nrob * gw; e += FRONT_NT) {
      const int rb = e / gw, k = e - rb * gw;
      unsigned word = 0u;
      if (k < h) {
        const int kind = s_gait[rb][0], it = s_gait[rb][1];
        for (int j = 0; j < 4; j++) {
          int on;
          if (kind == CMPC_GAIT_MIXED_FREQUENCY) {
            const int period = s_gait[rb][2 + j] > 0 ? s_gait[rb][2 + j] : 1;
            const int progress = (k + it + 1) % period;
            on = (float)progress < mul((float)period, s_duty[rb]);
          } else {
            int progress = (k + it + 1) % h - s_gait[rb][2 + j];
            if (progress < 0) progress += h;
            on = progress < s_gait[rb][6 + j];
          }
          word |= (unsigned)on << (8 * j);
        }
      }
      reinterpret_cast<unsigned*>(A.records + (size_t)(base + rb) * A.rec_stride)[CMPC_REC_TRAJ + tw + k] = word;
    }
  }
  if (i >= A.count) return;
  const cmpc_command& c = A.cmds[i];
  float* rec = reinterpret_cast<float*>(A.records + (size_t)i * A.rec_stride);

  // ---- solveDenseMPC (:618-828) ----
  float fe[6];
  for (int k = 0; k < 6; k++) fe[k] = A.f_ext[(size_t)i * 6 + k];
  if (c.have_log) {
    // f_external = x_k - A_prev x_prev - B_prev u_prev, rows 6..11 (:650-771)
    const float* R = c.log_R;
    const float Ib[3] = {0.07f, 0.26f, 0.242f};
    float Iw[9];  // R I_body R^T
    for (int a = 0; a < 3; a++)
      for (int bb = 0; bb < 3; bb++)
        Iw[3 * a + bb] = dot3(mul(R[3 * a], Ib[0]), R[3 * bb], mul(R[3 * a + 1], Ib[1]), R[3 * bb + 1], mul(R[3 * a + 2], Ib[2]), R[3 * bb + 2]);
    // inverse by cofactors / determinant
    const float c00 = sub(mul(Iw[4], Iw[8]), mul(Iw[5], Iw[7]));
    const float c01 = sub(mul(Iw[5], Iw[6]), mul(Iw[3], Iw[8]));
    const float c02 = sub(mul(Iw[3], Iw[7]), mul(Iw[4], Iw[6]));
    const float det = add(add(mul(Iw[0], c00), mul(Iw[1], c01)), mul(Iw[2], c02));
    const float id = __fdiv_rn(1.f, det);
    float Ii[9];
    Ii[0] = mul(c00, id);
    Ii[1] = mul(sub(mul(Iw[2], Iw[7]), mul(Iw[1], Iw[8])), id);
    Ii[2] = mul(sub(mul(Iw[1], Iw[5]), mul(Iw[2], Iw[4])), id);
    Ii[3] = mul(c01, id);
    Ii[4] = mul(sub(mul(Iw[0], Iw[8]), mul(Iw[2], Iw[6])), id);
    Ii[5] = mul(sub(mul(Iw[2], Iw[3]), mul(Iw[0], Iw[5])), id);
    Ii[6] = mul(c02, id);
    Ii[7] = mul(sub(mul(Iw[1], Iw[6]), mul(Iw[0], Iw[7])), id);
    Ii[8] = mul(sub(mul(Iw[0], Iw[4]), mul(Iw[1], Iw[3])), id);
    float bu_w[3] = {0.f, 0.f, 0.f}, bu_v[3] = {0.f, 0.f, 0.f};  // (B_prev u_prev) rows 6..8 and 9..11
    const float minv = __fdiv_rn(1.f, 12.f);
    for (int leg = 0; leg < 4; leg++) {
      const float u0 = -c.log_foot_force[3 * leg], u1 = -c.log_foot_force[3 * leg + 1], u2 = -c.log_foot_force[3 * leg + 2];
      const float rx = c.log_r_feet[leg], ry = c.log_r_feet[4 + leg], rz = c.log_r_feet[8 + leg];
      // [r]x u
      const float t0 = sub(mul(ry, u2), mul(rz, u1));
      const float t1 = sub(mul(rz, u0), mul(rx, u2));
      const float t2 = sub(mul(rx, u1), mul(ry, u0));
      for (int a = 0; a < 3; a++) bu_w[a] = add(bu_w[a], dot3(Ii[3 * a], t0, Ii[3 * a + 1], t1, Ii[3 * a + 2], t2));
      bu_v[0] = add(bu_v[0], mul(minv, u0));
      bu_v[1] = add(bu_v[1], mul(minv, u1));
      bu_v[2] = add(bu_v[2], mul(minv, u2));
    }
    const float* xp = c.log_x_prev;
    // A_prev rows 6..10 are zero; row 11: x_drag * x_prev[9] + x_prev[12], x_prev[12] = -9.81
    const float a11 = add(mul(c.log_x_drag, xp[9]), -9.81f);
    float f6[6];
    for (int a = 0; a < 3; a++) f6[a] = sub(c.omega_world[a], bu_w[a]);
    f6[3] = sub(c.v_world[0], bu_v[0]);
    f6[4] = sub(c.v_world[1], bu_v[1]);
    f6[5] = sub(sub(c.v_world[2], a11), bu_v[2]);
    fe[0] = -f6[0]; fe[1] = -f6[1]; fe[2] = f6[2]; fe[3] = f6[3]; fe[4] = f6[4]; fe[5] = f6[5];
    for (int k = 0; k < 6; k++) A.f_ext[(size_t)i * 6 + k] = fe[k];
  }
  rec[CMPC_REC_P + 0] = c.position[0];
  rec[CMPC_REC_P + 1] = c.position[1];
  rec[CMPC_REC_P + 2] = c.ground_z;
  for (int k = 0; k < 3; k++) { rec[CMPC_REC_V + k] = c.v_world[k]; rec[CMPC_REC_W + k] = c.omega_world[k]; }
  for (int k = 0; k < 4; k++) rec[CMPC_REC_Q + k] = c.orientation[k];
  for (int k = 0; k < 12; k++) rec[CMPC_REC_R + k] = sub(c.p_foot[3 * (k % 4) + k / 4], c.position[k / 4]);  // :779
  for (int k = 0; k < 12; k++) rec[CMPC_REC_WEIGHTS + k] = A.weights[k];
  rec[CMPC_REC_ALPHA] = A.alpha;
  rec[CMPC_REC_XDRAG] = c.x_comp_integral;  // update_x_drag before the integral moves (:809)
  for (int k = 0; k < 6; k++) rec[CMPC_REC_FDIST + k] = 0.f;  // the estimator stage writes xi when it applies
  rec[CMPC_REC_SIMTIME] = c.sim_time;
  if (A.sim_time) A.sim_time[i] = c.sim_time;
  rec[CMPC_REC_RSV] = 0.f;
  rec[CMPC_REC_RSV + 1] = 0.f;
  float xci = c.x_comp_integral;
  const float vx = c.v_world[0];
  if (vx > 0.3f || vx < -0.3f) {
    const float pz_err = sub(c.ground_z, c.body_height);
    xci = add(xci, __fdiv_rn(mul(mul(c.cmpc_x_drag, pz_err), dt), vx));
  }
  cmpc_command_result& r = A.results[i];
  r.world_position_desired[0] = wpd0;
  r.world_position_desired[1] = wpd1;
  r.x_comp_integral = xci;
  for (int k = 0; k < 6; k++) r.f_ext[k] = fe[k];
}

// window[i] <- window[i+1], window[last] <- sample once the window is full; plain append before that
__global__ void __launch_bounds__(128) cmpc_history_push_kernel(const cmpc_command* cmds, const float* f_ext, float* win_t,
                                                                  float* win_d, int have) {
  const int inst = blockIdx.x, N = CMPC_ADAPT_WINDOW;
  float* wt = win_t + (size_t)inst * N;
  float* wd = win_d + (size_t)inst * N;
  const float t = cmds[inst].sim_time, d = f_ext[(size_t)inst * 6 + 3];
  if (have < N) {
    if (threadIdx.x == 0) { wt[have] = t; wd[have] = d; }
    return;
  }
  float vt[4], vd[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int j = threadIdx.x + 128 * k;
    vt[k] = (j + 1 < N) ? wt[j + 1] : t;
    vd[k] = (j + 1 < N) ? wd[j + 1] : d;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int j = threadIdx.x + 128 * k;
    if (j < N) { wt[j] = vt[k]; wd[j] = vd[k]; }
  }
}

struct EpiArgs {
  const cmpc_command* cmds;
  const double* forces;  // [count][12 h]
  const int* status;
  const int* iterations;
  cmpc_command_result* results;
  int count, horizon;
};

__global__ void __launch_bounds__(128) cmpc_epilogue_kernel(const __grid_constant__ EpiArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.count) return;
  const float* R = A.cmds[i].r_body;
  const double* q = A.forces + (size_t)i * 12 * A.horizon;
  cmpc_command_result& r = A.results[i];
  for (int leg = 0; leg < 4; leg++) {
    const float f0 = (float)q[3 * leg], f1 = (float)q[3 * leg + 1], f2 = (float)q[3 * leg + 2];  // get_solution -> float (:836)
    r.fr_des[3 * leg] = f0; r.fr_des[3 * leg + 1] = f1; r.fr_des[3 * leg + 2] = f2;
    for (int a = 0; a < 3; a++) r.f_ff[3 * leg + a] = -dot3(R[3 * a], f0, R[3 * a + 1], f1, R[3 * a + 2], f2);  // :841
  }
  r.status = A.status[i];
  r.iterations = A.iterations[i];
  r.pad_[0] = 0.f;
}

}  // namespace

int cmpc_launch_frontend(const void* cmds, unsigned char* records, void* results, float* f_ext, float* sim_time, int count,
                         int horizon, int rec_stride, float dt, float alpha, const float weights[12], void* stream) {
  if (count <= 0) return 0;
  FrontArgs A;
  A.cmds = static_cast<const cmpc_command*>(cmds);
  A.records = records;
  A.results = static_cast<cmpc_command_result*>(results);
  A.f_ext = f_ext;
  A.sim_time = sim_time;
  A.count = count;
  A.horizon = horizon;
  A.rec_stride = rec_stride;
  A.dt = dt;
  A.alpha = alpha;
  for (int i = 0; i < 12; i++) A.weights[i] = weights[i];
  cmpc_frontend_kernel<<<(count + FRONT_NT - 1) / FRONT_NT, FRONT_NT, 0, (cudaStream_t)stream>>>(A);
  return (int)cudaGetLastError();
}

int cmpc_launch_history_push(const void* cmds, const float* f_ext, float* win_t, float* win_d, int count, int have,
                             void* stream) {
  if (count <= 0) return 0;
  cmpc_history_push_kernel<<<count, 128, 0, (cudaStream_t)stream>>>(static_cast<const cmpc_command*>(cmds), f_ext, win_t,
                                                                     win_d, have);
  return (int)cudaGetLastError();
}

int cmpc_launch_epilogue(const void* cmds, const double* forces, const int* status, const int* iterations, void* results,
                         int count, int horizon, void* stream) {
  if (count <= 0) return 0;
  EpiArgs A;
  A.cmds = static_cast<const cmpc_command*>(cmds);
  A.forces = forces;
  A.status = status;
  A.iterations = iterations;
  A.results = static_cast<cmpc_command_result*>(results);
  A.count = count;
  A.horizon = horizon;
  cmpc_epilogue_kernel<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(A);
  return (int)cudaGetLastError();
}
