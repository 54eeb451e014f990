// What does an instruction of a HELPER warp cost while the MAIN warps of its SM sub-partition stream DMMAs?
// CTA = 4 main warps + 4 helper warps (main w and helper w + 4 share a sub-partition), C CTAs per SM.  A main warp
// alternates a burst of 72 DMMA m8n8k4 with `gap` cycles of no FP64 work (as the inversion kernel's update phase and its
// other phases).  A helper warp runs a dependent chain of N instructions of one kind and reports cycles per instruction:
//   kind 0: DFMA   1: FFMA (fp32)   2: DMMA (dependent pairs)   3: SHFL.64 + DFMA   4: independent DFMAs (8 chains)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o helper_chain_ubench helper_chain_ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int KIND>
__global__ void __launch_bounds__(256) k(double* out, int iters, int gap, int mains_on, unsigned long long* cyc) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __shared__ volatile int stop;
  __shared__ int done_helpers;
  if (threadIdx.x == 0) { stop = 0; done_helpers = 0; }
  __syncthreads();
  if (w < 4) {  // main
    double t[36][2], pf[8], mf[8];
#pragma unroll
    for (int i = 0; i < 36; i++) { t[i][0] = 0.0; t[i][1] = 0.0; }
#pragma unroll
    for (int i = 0; i < 8; i++) { pf[i] = 1e-3 * (lane + i); mf[i] = 1e-3 * (lane - i); }
    for (int guard = 0; mains_on && !stop && guard < 200000; guard++) {  // bounded: never spin for ever
#pragma unroll
      for (int ks = 0; ks < 2; ks++)
#pragma unroll
        for (int I = 0; I < 8; I++)
#pragma unroll
          for (int J = 0; J <= I; J++) dmma(t[I * (I + 1) / 2 + J][0], t[I * (I + 1) / 2 + J][1], pf[I], mf[J]);
      const long long s0 = clock64();
      while (clock64() - s0 < gap) { }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 36; i++) s += t[i][0] + t[i][1];
    if (s == 123.456) out[0] = s;
    return;
  }
  // helper
  double a = 1.0 + lane, b = 0.5, c0 = 0.0, c1 = 0.0;
  float fa = 1.0f + lane;
  double e[8] = {1, 2, 3, 4, 5, 6, 7, 8};
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (KIND == 0) a = fma(a, 1.0000001, 1e-9);
      if (KIND == 1) fa = fmaf(fa, 1.0000001f, 1e-9f);
      if (KIND == 2) dmma(c0, c1, a, b);
      if (KIND == 3) { a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31); a = fma(a, 1.0000001, 1e-9); }
      if (KIND == 4) {
#pragma unroll
        for (int j = 0; j < 8; j++) e[j] = fma(e[j], 1.0000001, 1e-9);
      }
    }
  }
  const long long t1 = clock64();
  double s = a + fa + c0 + c1;
#pragma unroll
  for (int j = 0; j < 8; j++) s += e[j];
  if (s == 123.456) out[1] = s;
  if (lane == 0) atomicAdd(cyc, (unsigned long long)(t1 - t0));
  __syncwarp();
  if (w == 4 && lane == 0) { __threadfence_block(); }
  // the last helper of the CTA to finish releases the mains
  if (lane == 0) { if (atomicAdd((int*)&done_helpers, 1) == 3) stop = 1; }
}
template <int KIND>
void run(const char* tag, int ctas, int gap, int mains_on, int sms) {
  double* out; unsigned long long* cyc;
  cudaMalloc(&out, 16); cudaMalloc(&cyc, 8);
  const int iters = 300;
  for (int rep = 0; rep < 2; rep++) {
    cudaMemset(cyc, 0, 8);
    k<KIND><<<sms * ctas, 256>>>(out, iters, gap, mains_on, cyc);
    cudaDeviceSynchronize();
  }
  unsigned long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (sms * ctas * 4.0) / (iters * 16.0) / (KIND == 4 ? 8.0 : 1.0);
  printf("%-34s CTAs/SM %d  mains %-3s gap %5d: %6.1f cycles per helper instruction\n", tag, ctas, mains_on ? "on" : "off", gap, per);
  fflush(stdout);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  for (int ctas = 1; ctas <= 2; ctas++)
    for (int gap : {0, 1500, 3000}) {
      run<0>("dependent DFMA chain", ctas, gap, 1, sms);
      run<1>("dependent FFMA chain", ctas, gap, 1, sms);
      run<2>("dependent DMMA chain", ctas, gap, 1, sms);
      run<3>("SHFL.64 + DFMA chain", ctas, gap, 1, sms);
      run<4>("eight independent DFMA chains", ctas, gap, 1, sms);
    }
  run<0>("dependent DFMA chain", 2, 0, 0, sms);
  run<1>("dependent FFMA chain", 2, 0, 0, sms);
  run<2>("dependent DMMA chain", 2, 0, 0, sms);
  run<3>("SHFL.64 + DFMA chain", 2, 0, 0, sms);
  run<4>("eight independent DFMA chains", 2, 0, 0, sms);
  return 0;
}
