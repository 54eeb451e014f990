// Rounding behaviour of DMMA m8n8k4 against an FMA chain: C = sum over many k-steps of A_k B_k with random data,
// compared with a long-double reference on the host.  Is the tensor-core accumulation round-to-nearest?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_precision dmma_precision.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
__global__ void k(const double* A, const double* B, double* Cmma, double* Cfma, int ksteps) {
  // one warp: C (8x8) += A_k (8x4) B_k (4x8); A stored [ks][8][4], B stored [ks][4][8]
  const int lane = threadIdx.x, r = lane >> 2, q = lane & 3;
  double c0 = 0.0, c1 = 0.0, f0 = 0.0, f1 = 0.0;
  for (int ks = 0; ks < ksteps; ks++) {
    const double a = A[ks * 32 + r * 4 + q];   // A[r][k=q]
    const double b = B[ks * 32 + q * 8 + r];   // B[k=q][n=r]
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    for (int kk = 0; kk < 4; kk++) {           // the same sum as an FMA chain in k order
      f0 = fma(A[ks * 32 + r * 4 + kk], B[ks * 32 + kk * 8 + 2 * q], f0);
      f1 = fma(A[ks * 32 + r * 4 + kk], B[ks * 32 + kk * 8 + 2 * q + 1], f1);
    }
  }
  Cmma[r * 8 + 2 * q] = c0; Cmma[r * 8 + 2 * q + 1] = c1;
  Cfma[r * 8 + 2 * q] = f0; Cfma[r * 8 + 2 * q + 1] = f1;
}
int main() {
  for (int ksteps : {1, 16, 256, 4096}) {
    for (int mode = 0; mode < 2; mode++) {  // 0: signed random (cancellation), 1: positive (growing sum)
      std::vector<double> A(ksteps * 32), B(ksteps * 32);
      srand(7);
      for (auto& x : A) x = (mode ? 0.0 : -0.5) + rand() / (double)RAND_MAX;
      for (auto& x : B) x = (mode ? 0.0 : -0.5) + rand() / (double)RAND_MAX;
      double *dA, *dB, *dC, *dF;
      cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, 512); cudaMalloc(&dF, 512);
      cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
      cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
      k<<<1, 32>>>(dA, dB, dC, dF, ksteps);
      double C[64], F[64];
      cudaMemcpy(C, dC, 512, cudaMemcpyDeviceToHost); cudaMemcpy(F, dF, 512, cudaMemcpyDeviceToHost);
      double em = 0, ef = 0, bm = 0, bf = 0, mag = 0;
      for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
          long double ref = 0;
          for (int ks = 0; ks < ksteps; ks++)
            for (int kk = 0; kk < 4; kk++) ref += (long double)A[ks * 32 + i * 4 + kk] * (long double)B[ks * 32 + kk * 8 + j];
          const double dm = (double)((long double)C[i * 8 + j] - ref), df = (double)((long double)F[i * 8 + j] - ref);
          em = fmax(em, fabs(dm)); ef = fmax(ef, fabs(df)); bm += dm; bf += df; mag = fmax(mag, fabs((double)ref));
        }
      printf("k-steps %5d  %s: |C| ~ %.3g   DMMA max err %.3e (mean %+.3e)   FMA chain max err %.3e (mean %+.3e)   ratio %.1f\n",
             ksteps, mode ? "positive data" : "signed data  ", mag, em, bm / 64, ef, bf / 64, em / fmax(ef, 1e-300));
      cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dF);
    }
  }
  return 0;
}
