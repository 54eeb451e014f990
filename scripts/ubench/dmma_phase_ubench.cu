// Can two warps of an SM sub-partition overlap one's DMMA phase with the other's non-DMMA phase?
// Each warp alternates: spin for D cycles (no FP64), then U tile updates (72 DMMA m8n8k4 on 36 register tiles).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_phase_ubench dmma_phase_ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>  // 0: spin by clock, 1: spin phase does shared-memory traffic, 2: spin phase does FP64 FMA chain
__global__ void __launch_bounds__(128) k(double* out, int iters, int spin, long long* cyc, int* sm_slots, int stagger) {
  __shared__ double sm[4][1024];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double t[36][2];
#pragma unroll
  for (int i = 0; i < 36; i++) { t[i][0] = 0.0; t[i][1] = 0.0; }
  double pf[8], mf[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { pf[i] = 1e-3 * (lane + i); mf[i] = 1e-3 * (lane - i); }
  for (int i = lane; i < 1024; i += 32) sm[w][i] = i;
  __syncwarp();
  double acc = 1.0;
  long long dm = 0;
  if (stagger > 0) {  // the CTAs of an SM start `stagger` cycles apart, in the order they arrive on it
    __shared__ int slot_s;
    if (threadIdx.x == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      slot_s = atomicAdd(sm_slots + smid, 1);
    }
    __syncthreads();
    const long long s0 = clock64();
    while (clock64() - s0 < (long long)stagger * slot_s) { }
  }
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (spin > 0) {
      const long long s0 = clock64();
      if (MODE == 0) { while (clock64() - s0 < spin) { } }
      if (MODE == 1) { while (clock64() - s0 < spin) { acc += sm[w][(lane * 9 + (int)acc) & 1023]; sm[w][(lane + 64) & 1023] = acc; } }
      if (MODE == 2) { while (clock64() - s0 < spin) { acc = fma(acc, 1.0000001, 1e-9); } }
    }
    const long long d0 = clock64();
#pragma unroll
    for (int ks = 0; ks < 2; ks++)
#pragma unroll
      for (int I = 0; I < 8; I++)
#pragma unroll
        for (int J = 0; J <= I; J++) dmma(t[I * (I + 1) / 2 + J][0], t[I * (I + 1) / 2 + J][1], pf[I], mf[J]);
    // the phase ends when the results are there (as in the kernel: the next phase reads the tiles)
    double chk = 0.0;
#pragma unroll
    for (int i = 0; i < 36; i++) chk += t[i][0];
    if (chk == 123.456) acc += 1.0;
    dm += clock64() - d0;
  }
  const long long t1 = clock64();
  double s = acc;
#pragma unroll
  for (int i = 0; i < 36; i++) s += t[i][0] + t[i][1];
  if (s == 123.456) out[0] = s;
  if (lane == 0) { atomicAdd((unsigned long long*)cyc, (unsigned long long)(t1 - t0)); atomicAdd((unsigned long long*)cyc + 1, (unsigned long long)dm); }
}
template <int MODE>
void run(const char* tag, int ctas_per_sm, int spin, int sms, int stagger = 0) {
  double* out; long long* cyc; int* slots;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 16); cudaMalloc(&slots, 4 * 1024);
  const int iters = 400;
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<MODE>, 128, 0);
  for (int rep = 0; rep < 2; rep++) {
    cudaMemset(cyc, 0, 16);
    cudaMemset(slots, 0, 4 * 1024);
    k<MODE><<<sms * ctas_per_sm, 128>>>(out, iters, spin, cyc, slots, stagger);
    cudaDeviceSynchronize();
  }
  long long h[2];
  cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
  const double warps = (double)sms * ctas_per_sm * 4;
  const double per_iter = (double)h[0] / warps / iters, dmma_phase = (double)h[1] / warps / iters;
  printf("%-22s stagger %5d  warps/sub-partition %d  spin %5d cyc: %7.0f cyc per iteration, DMMA phase %6.0f cyc (%4.1f cyc per DMMA), pipe busy %4.1f%%  (occupancy limit %d CTAs)\n",
         tag, stagger, ctas_per_sm, spin, per_iter, dmma_phase, dmma_phase / 72.0, 100.0 * ctas_per_sm * 72 * 16 / per_iter, nb);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  for (int c = 1; c <= 2; c++) {
    run<0>("clock spin", c, 0, sms);
    run<0>("clock spin", c, 1200, sms);
    run<0>("clock spin", c, 2400, sms);
    run<1>("shared-memory traffic", c, 2400, sms);
    run<2>("FP64 FMA chain", c, 2400, sms);
  }
  // two warps per sub-partition, the second CTA of an SM started half a period late
  run<0>("clock spin", 2, 1200, sms, 1400);
  run<0>("clock spin", 2, 2400, sms, 2000);
  run<0>("clock spin", 2, 2400, sms, 2600);
  run<1>("shared-memory traffic", 2, 2400, sms, 2000);
  run<2>("FP64 FMA chain", 2, 2400, sms, 2000);
  return 0;
}
