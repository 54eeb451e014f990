// The caller of the solve path, on the device (row (f) of DESIGN.md §0): ConvexMPCLocomotion::updateMPCIfNeeded and
// solveDenseMPC (ConvexMPCLocomotion.cpp:511-870) plus the gait's getMpcTable (Gait.cpp:158-215) for a batch of
// robots.  Three small kernels around the solve pipeline:
//
//   cmpc_frontend_kernel      one warp per robot: reference trajectory, contact table, r = pFoot - p, the f_ext
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

constexpr int FRONT_WPC = 4;  // robots (warps) per CTA

// the f_ext residual of the previous step's model: f_external = x_k - A_prev x_prev - B_prev u_prev, rows 6..11, with
// the signs of ConvexMPCLocomotion.cpp:771
__device__ void external_force(const cmpc_command& c, float fe[6]) {
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
  // A_prev rows 6..10 are zero; row 11: x_drag * x_prev[9] + x_prev[12], x_prev[12] = -9.81
  const float a11 = add(mul(c.log_x_drag, c.log_x_prev[9]), -9.81f);
  float f6[6];
  for (int a = 0; a < 3; a++) f6[a] = sub(c.omega_world[a], bu_w[a]);
  f6[3] = sub(c.v_world[0], bu_v[0]);
  f6[4] = sub(c.v_world[1], bu_v[1]);
  f6[5] = sub(sub(c.v_world[2], a11), bu_v[2]);
  fe[0] = -f6[0]; fe[1] = -f6[1]; fe[2] = f6[2]; fe[3] = f6[3]; fe[4] = f6[4]; fe[5] = f6[5];
}

// One warp per robot.  The command struct is staged in shared memory with coalesced loads; lane 0 works out the
// trajectory constants and the command state (updateMPCIfNeeded), lane 1 the f_ext residual (solveDenseMPC); then the
// warp fills the record together, consecutive lanes on consecutive words, every walking trajectory entry re-running
// its own float accumulation (:579-581).
__global__ void __launch_bounds__(32 * FRONT_WPC) cmpc_frontend_kernel(const __grid_constant__ FrontArgs A) {
  constexpr int CW = sizeof(cmpc_command) / 4;
  __shared__ __align__(16) unsigned s_cmd[FRONT_WPC][CW];
  __shared__ float s_t0[FRONT_WPC][12];
  __shared__ float s_walk[FRONT_WPC][3];  // dt * yaw_turn_rate, dt * v_des_world[0], dt * v_des_world[1]
  __shared__ float s_fe[FRONT_WPC][6];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * FRONT_WPC + w;
  if (i >= A.count) return;
  const int h = A.horizon;
  const float dt = A.dt;
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(A.cmds + i);
    for (int k = lane; k < CW; k += 32) s_cmd[w][k] = src[k];
  }
  __syncwarp();
  const cmpc_command& c = *reinterpret_cast<const cmpc_command*>(s_cmd[w]);
  cmpc_command_result& res = A.results[i];
  if (lane == 0) {
    // ---- updateMPCIfNeeded (:511-594) ----
    float wpd0 = c.world_position_desired[0], wpd1 = c.world_position_desired[1];
    if (c.stand) {
      const float t0[12] = {c.roll_des, c.pitch_des, c.stand_traj[2], c.stand_traj[0], c.stand_traj[1], c.body_height,
                            0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < 12; j++) s_t0[w][j] = t0[j];
      s_walk[w][0] = s_walk[w][1] = s_walk[w][2] = 0.f;  // the stand trajectory is constant (:524-531)
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
      for (int j = 0; j < 12; j++) s_t0[w][j] = t0[j];
      s_walk[w][0] = mul(dt, c.yaw_turn_rate);
      s_walk[w][1] = mul(dt, vw0);
      s_walk[w][2] = mul(dt, vw1);
    }
    // the x_drag integral moves after update_x_drag has taken the old value (:809-816)
    float xci = c.x_comp_integral;
    const float vx = c.v_world[0];
    if (vx > 0.3f || vx < -0.3f) {
      const float pz_err = sub(c.ground_z, c.body_height);
      xci = add(xci, __fdiv_rn(mul(mul(c.cmpc_x_drag, pz_err), dt), vx));
    }
    res.world_position_desired[0] = wpd0;
    res.world_position_desired[1] = wpd1;
    res.x_comp_integral = xci;
    if (A.sim_time) A.sim_time[i] = c.sim_time;
  } else if (lane == 1) {
    // ---- solveDenseMPC (:647-776): f_ext stays as it was without /log_data ----
    float fe[6];
    for (int k = 0; k < 6; k++) fe[k] = A.f_ext[(size_t)i * 6 + k];
    if (c.have_log) {
      external_force(c, fe);
      for (int k = 0; k < 6; k++) A.f_ext[(size_t)i * 6 + k] = fe[k];
    }
    for (int k = 0; k < 6; k++) { s_fe[w][k] = fe[k]; res.f_ext[k] = fe[k]; }
  }
  __syncwarp();
  float* rec = reinterpret_cast<float*>(A.records + (size_t)i * A.rec_stride);
  // ---- the fixed part of the record: what solveDenseMPC hands update_problem_data_floats (:779-828) ----
  for (int k = lane; k < CMPC_REC_TRAJ; k += 32) {
    float v = 0.f;  // xi (the estimator stage writes it when it applies), reserved words
    if (k < 2) v = c.position[k];
    else if (k == 2) v = c.ground_z;  // p_v (:640)
    else if (k < CMPC_REC_Q) v = c.v_world[k - CMPC_REC_V];
    else if (k < CMPC_REC_W) v = c.orientation[k - CMPC_REC_Q];
    else if (k < CMPC_REC_R) v = c.omega_world[k - CMPC_REC_W];
    else if (k < CMPC_REC_WEIGHTS) { const int j = k - CMPC_REC_R; v = sub(c.p_foot[3 * (j % 4) + j / 4], c.position[j / 4]); }  // :779
    else if (k < CMPC_REC_ALPHA) v = A.weights[k - CMPC_REC_WEIGHTS];
    else if (k == CMPC_REC_ALPHA) v = A.alpha;
    else if (k == CMPC_REC_XDRAG) v = c.x_comp_integral;  // update_x_drag before the integral moves (:809)
    else if (k == CMPC_REC_SIMTIME) v = c.sim_time;
    rec[k] = v;
  }
  // ---- trajAll (:565-583) ----
  const int tw = 12 * h;
  const bool stand = c.stand != 0;
  for (int e = lane; e < tw; e += 32) {
    const int k = e / 12, j = e - 12 * k;
    float v = s_t0[w][j];
    if (j >= 2 && j <= 4 && !stand) {
      const float d = s_walk[w][j - 2];
      for (int q = 0; q < k; q++) v = add(v, d);
    }
    rec[CMPC_REC_TRAJ + e] = v;
  }
  // ---- getMpcTable (Gait.cpp:158-215), nIterations == horizon: the four legs of a step are one word ----
  const int gw = (A.rec_stride - 4 * (CMPC_REC_TRAJ + tw)) / 4;  // 4 h bytes + padding
  for (int k = lane; k < gw; k += 32) {
    unsigned word = 0u;
    if (k < h) {
      for (int j = 0; j < 4; j++) {
        int on;
        if (c.gait_kind == CMPC_GAIT_MIXED_FREQUENCY) {
          const int period = c.gait_offsets[j] > 0 ? c.gait_offsets[j] : 1;
          const int progress = (k + c.gait_iteration + 1) % period;
          on = (float)progress < mul((float)period, c.gait_duty);
        } else {
          int progress = (k + c.gait_iteration + 1) % h - c.gait_offsets[j];
          if (progress < 0) progress += h;
          on = progress < c.gait_durations[j];
        }
        word |= (unsigned)on << (8 * j);
      }
    }
    reinterpret_cast<unsigned*>(rec)[CMPC_REC_TRAJ + tw + k] = word;
  }
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
  cmpc_frontend_kernel<<<(count + FRONT_WPC - 1) / FRONT_WPC, 32 * FRONT_WPC, 0, (cudaStream_t)stream>>>(A);
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
