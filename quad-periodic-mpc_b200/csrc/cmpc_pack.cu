// Record packing on the device: the end-to-end call hands the kernel the caller's structure-of-arrays inputs
// (update_data_t's fields, convexMPC_interface.h:23-42, one array per field) as PINNED HOST pointers; the kernel
// reads them over PCIe with linear, fully coalesced loads (one grid-stride pass per source array) and scatters
// the words into the instance records in HBM (cmpc_device.h), so the host packs nothing and issues one launch
// per chunk instead of one copy per array.  The same kernel serves device-resident SoA inputs.
#include <cuda_runtime.h>
#include <stdint.h>

#include "cmpc_device.h"

namespace {

struct Seg {
  const uint32_t* src;  // source array, `width` words per instance (nullptr: write zeros)
  int width;            // words per instance in the source
  int dst;              // first word of the field in the record
};

constexpr int PACK_SEGS = 11;
struct PackArgs {
  Seg seg[PACK_SEGS];
  uint32_t* records;
  int rec_words;   // record stride in words
  int count;
  int gait_words;  // h: words of gait bytes per instance
  int tail_dst;    // first word after the gait bytes
  int tail_words;  // zero padding after the gait bytes
};

__global__ void __launch_bounds__(256) cmpc_pack_records_kernel(const __grid_constant__ PackArgs A) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 1
  for (int s = 0; s < PACK_SEGS; s++) {
    const Seg sg = A.seg[s];
    if (sg.width == 0) continue;
    const long long total = (long long)A.count * sg.width;
    for (long long j = t0; j < total; j += stride) {
      const int inst = (int)(j / sg.width), k = (int)(j - (long long)inst * sg.width);
      const uint32_t v = sg.src ? __ldcs(sg.src + j) : 0u;
      A.records[(size_t)inst * A.rec_words + sg.dst + k] = v;
    }
  }
  // reserved words and the padding behind the gait bytes
  const int zw = 3 + A.tail_words;
  for (long long j = t0; j < (long long)A.count * zw; j += stride) {
    const int inst = (int)(j / zw), k = (int)(j - (long long)inst * zw);
    const int dst = k < 3 ? CMPC_REC_SIMTIME + k : A.tail_dst + (k - 3);
    A.records[(size_t)inst * A.rec_words + dst] = 0u;
  }
}

}  // namespace

// in: the eleven arrays of cmpc_inputs as device-accessible pointers, already offset to the first instance
int cmpc_launch_pack(const void* p, const void* v, const void* q, const void* w, const void* r, const void* weights,
                     const void* traj, const void* alpha, const void* gait, const void* x_drag, const void* f_dist,
                     unsigned char* records, int rec_stride, int horizon, int count, int sm_count, void* stream) {
  if (count <= 0) return 0;
  PackArgs A;
  auto set = [&](int i, const void* src, int width, int dst) {
    A.seg[i].src = static_cast<const uint32_t*>(src);
    A.seg[i].width = width;
    A.seg[i].dst = dst;
  };
  const int h = horizon;
  set(0, traj, 12 * h, CMPC_REC_TRAJ);  // the largest array first: its loads are in flight while the rest is issued
  set(1, gait, h, CMPC_REC_TRAJ + 12 * h);
  set(2, r, 12, CMPC_REC_R);
  set(3, weights, 12, CMPC_REC_WEIGHTS);
  set(4, p, 3, CMPC_REC_P);
  set(5, v, 3, CMPC_REC_V);
  set(6, q, 4, CMPC_REC_Q);
  set(7, w, 3, CMPC_REC_W);
  set(8, alpha, 1, CMPC_REC_ALPHA);
  set(9, x_drag, 1, CMPC_REC_XDRAG);
  set(10, f_dist, 6, CMPC_REC_FDIST);  // nullptr: zeros (SolverMPC.cpp:813)
  A.records = reinterpret_cast<uint32_t*>(records);
  A.rec_words = rec_stride / 4;
  A.count = count;
  A.gait_words = h;
  A.tail_dst = CMPC_REC_TRAJ + 13 * h;
  A.tail_words = A.rec_words - A.tail_dst;
  const long long words = (long long)count * (12 * h);
  int grid = (int)((words + 255) / 256);
  if (grid > sm_count * 2) grid = sm_count * 2;  // few, fat CTAs: they also have to find room beside resident solve kernels
  if (grid < 1) grid = 1;
  cmpc_pack_records_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A);
  return (int)cudaGetLastError();
}
