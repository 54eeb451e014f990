// Microbenchmarks behind the sweep design: DFMA throughput / dependent latency, DMMA m8n8k4 throughput / latency,
// shuffle and LDS latency.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ubench fp64_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_tp(double* out, int iters) {
  double a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  const double m = 0.999999, c = 1e-9;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
  double s = 0; for (int i = 0; i < 8; i++) s += a[i];
  if (s == 1.2345) out[0] = s;
}
__global__ void k_dfma_lat(double* out, long long* cyc, int iters) {
  double a = threadIdx.x;
  const double m = 0.999999, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a = fma(a, m, c);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (a == 1.2345) out[0] = a;
}
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k_dmma_tp(double* out, int iters) {
  double d[NACC][2];
  for (int i = 0; i < NACC; i++) { d[i][0] = threadIdx.x; d[i][1] = i; }
  double a = 1e-3 * threadIdx.x, b = 1e-3;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma(d[i][0], d[i][1], a, b);
  double s = 0; for (int i = 0; i < NACC; i++) s += d[i][0] + d[i][1];
  if (s == 1.2345) out[0] = s;
}
// distinct A / B operand registers per DMMA (as in a real tile update), NACC accumulators
template <int NACC>
__global__ void k_dmma_tp_ops(double* out, int iters) {
  double d[NACC][2], a[NACC], b[NACC];
  for (int i = 0; i < NACC; i++) { d[i][0] = threadIdx.x; d[i][1] = i; a[i] = 1e-3 * (threadIdx.x + i); b[i] = 1e-3 * (i + 1); }
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma(d[i][0], d[i][1], a[i], b[(i + it) & (NACC - 1)]);
  double s = 0; for (int i = 0; i < NACC; i++) s += d[i][0] + d[i][1];
  if (s == 1.2345) out[0] = s;
}
// the tile update of the inversion kernel: 36 accumulator tiles, A fragment per tile row, B fragment per tile column
__global__ void k_dmma_update(double* out, int iters) {
  double t[36][2], pf[8][2], mf[8][2];
  for (int k = 0; k < 36; k++) { t[k][0] = threadIdx.x + k; t[k][1] = k; }
  for (int k = 0; k < 8; k++) { pf[k][0] = 1e-3 * (threadIdx.x + k); pf[k][1] = 2e-3 * k; mf[k][0] = 1e-3 * k; mf[k][1] = 3e-3 * (k + 1); }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int I = 0; I < 8; I++)
#pragma unroll
      for (int J = 0; J <= I; J++) dmma(t[I * (I + 1) / 2 + J][0], t[I * (I + 1) / 2 + J][1], pf[I][0], mf[J][0]);
#pragma unroll
    for (int I = 0; I < 8; I++)
#pragma unroll
      for (int J = 0; J <= I; J++) dmma(t[I * (I + 1) / 2 + J][0], t[I * (I + 1) / 2 + J][1], pf[I][1], mf[J][1]);
  }
  double s = 0; for (int k = 0; k < 36; k++) s += t[k][0] + t[k][1];
  if (s == 1.2345) out[0] = s;
}
__global__ void k_dmma_lat(double* out, long long* cyc, int iters) {
  double d0 = threadIdx.x, d1 = 1;
  double a = 1e-3 * threadIdx.x, b = 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) dmma(d0, d1, a, b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (d0 + d1 == 1.2345) out[0] = d0;
}
__global__ void k_shfl_lat(double* out, long long* cyc, int iters) {
  double a = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a = __shfl_xor_sync(0xffffffffu, a, 1 + (i & 3)) + 1.0;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (a == 1.2345) out[0] = a;
}
__global__ void k_rcp_lat(double* out, long long* cyc, int iters) {
  double a = 1.5 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
      double e = fma(-a, r, 1.0); r = fma(r, e, r); e = fma(-a, r, 1.0); r = fma(r, e, r);
      a = r + 1.25;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (a == 1.2345) out[0] = a;
}
__global__ void k_lds_lat(double* out, long long* cyc, int iters) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 7 + 1) & 1023;
  __syncthreads();
  int j = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) j = idx[j];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (j == 12345) out[0] = j;
}

template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMallocManaged(&cyc, 8);
  const int it = 20000;
  for (int wpsm : {4, 8, 16, 32, 64}) {
    int thr = 128, blocks = sms * wpsm * 32 / thr;
    float ms = timeit([&] { k_dfma_tp<<<blocks, thr>>>(out, it); });
    double fl = 2.0 * 8 * it * (double)blocks * thr;
    float ms2 = timeit([&] { k_dmma_tp<8><<<blocks, thr>>>(out, it); });
    double fl2 = 2.0 * 256 * 8 * it * (double)blocks * thr / 32;
    float ms3 = timeit([&] { k_dmma_tp<2><<<blocks, thr>>>(out, it); });
    double fl3 = 2.0 * 256 * 2 * it * (double)blocks * thr / 32;
    float ms4 = timeit([&] { k_dmma_tp_ops<8><<<blocks, thr>>>(out, it); });
    printf("warps/SM %2d: DFMA %.2f TF/s   DMMA(8 acc) %.2f TF/s   DMMA(2 acc) %.2f TF/s   DMMA(8 acc, distinct operands) %.2f TF/s\n", wpsm, fl / ms / 1e9, fl2 / ms2 / 1e9, fl3 / ms3 / 1e9, fl2 / ms4 / 1e9);
  }
  for (int wpsm : {4, 8, 12}) {
    int blocks = sms * wpsm;
    float ms = timeit([&] { k_dmma_update<<<blocks, 32>>>(out, 2000); });
    printf("tile update (72 DMMA, 36 tiles) at %d warps/SM: %.2f TF/s\n", wpsm, 2.0 * 256 * 72 * 2000 * (double)blocks / ms / 1e9);
  }
  k_dfma_lat<<<1, 32>>>(out, cyc, 1000); cudaDeviceSynchronize(); printf("DFMA dependent latency %.1f cyc\n", cyc[0] / 16000.0);
  k_dmma_lat<<<1, 32>>>(out, cyc, 1000); cudaDeviceSynchronize(); printf("DMMA dependent latency %.1f cyc\n", cyc[0] / 16000.0);
  k_shfl_lat<<<1, 32>>>(out, cyc, 1000); cudaDeviceSynchronize(); printf("SHFL.64 + DADD dependent latency %.1f cyc\n", cyc[0] / 16000.0);
  k_rcp_lat<<<1, 32>>>(out, cyc, 1000); cudaDeviceSynchronize(); printf("rcp (MUFU + 2 Newton) + DADD latency %.1f cyc\n", cyc[0] / 16000.0);
  k_lds_lat<<<1, 32>>>(out, cyc, 1000); cudaDeviceSynchronize(); printf("LDS dependent latency %.1f cyc\n", cyc[0] / 16000.0);
  printf("%s, %d SMs\n", p.name, sms);
  return 0;
}
