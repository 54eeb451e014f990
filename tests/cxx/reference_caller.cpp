// A C++ caller that sees ONLY the reference's own declarations for the path (convexMPC_interface.h:9-52, restated
// below without the Eigen include the image does not have) and links libcmpc_b200.so in place of
// convexMPC_interface.cpp + SolverMPC.cpp.  update_x_drag is declared WITHOUT extern "C", exactly as the reference
// header does, so this translation unit references the C++-mangled symbol.
//
// usage: reference_caller <instance.bin> <forces.bin>
//   instance.bin: int32 horizon, float dt, mu, f_max, x_drag, alpha, then p[3] v[3] q[4] w[3] r[12] weights[12]
//                 traj[12 h] as float32 and gait[4 h] as int32
//   forces.bin:   12 h doubles from get_solution()
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifdef __cplusplus
#define EXTERNC extern "C"
#else
#define EXTERNC
#endif

EXTERNC void setup_problem(double dt, int horizon, double mu, double f_max);
EXTERNC void update_problem_data(double* p, double* v, double* q, double* w, double* r, double yaw, double* weights,
                                 double* state_trajectory, double alpha, int* gait);
EXTERNC double get_solution(int index);
EXTERNC void update_solver_settings(int max_iter, double rho, double sigma, double solver_alpha, double terminate,
                                    double use_jcqp);
EXTERNC void update_problem_data_floats(float* p, float* v, float* q, float* w, float* r, float roll, float pitch,
                                        float yaw, float* weights, float* state_trajectory, float alpha, int* gait);

void update_x_drag(float x_drag);

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  int32_t h = 0;
  float hdr[5];
  if (std::fread(&h, 4, 1, f) != 1 || std::fread(hdr, 4, 5, f) != 5) return 4;
  std::vector<float> st(3 + 3 + 4 + 3 + 12 + 12 + 12 * h);
  std::vector<int> gait(4 * h);
  if (std::fread(st.data(), 4, st.size(), f) != st.size()) return 4;
  if (std::fread(gait.data(), 4, gait.size(), f) != gait.size()) return 4;
  std::fclose(f);
  float* p = st.data();
  float *v = p + 3, *q = v + 3, *w = q + 4, *r = w + 3, *weights = r + 12, *traj = weights + 12;
  setup_problem(hdr[0], h, hdr[1], hdr[2]);
  update_x_drag(hdr[3]);
  update_solver_settings(100, 1e-7, 1e-8, 1.5, 0.1, 0.0);
  update_problem_data_floats(p, v, q, w, r, 0.f, 0.f, 0.f, weights, traj, hdr[4], gait.data());
  std::vector<double> out(12 * h);
  for (int i = 0; i < 12 * h; i++) out[i] = get_solution(i);
  FILE* g = std::fopen(argv[2], "wb");
  if (!g) return 5;
  std::fwrite(out.data(), 8, out.size(), g);
  std::fclose(g);
  return 0;
}
