// Record packing on the device: the end-to-end call hands the kernel the caller's structure-of-arrays inputs
// (update_data_t's fields, convexMPC_interface.h:23-42, one array per field) as PINNED HOST pointers; the kernel
// reads them over PCIe with linear, fully coalesced loads (one grid-stride pass per source array) and scatters
// the words into the instance records in HBM (cmpc_device.h), so the host packs nothing and issues one launch
// per chunk instead of one copy per array.  The same kernel serves device-resident SoA inputs.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

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
  int vec;         // 16-byte source loads where the array is aligned
  int zero_tail;   // write the reserved words and the padding (once per record)
};

// Each thread keeps PACK_UNROLL independent loads in flight per pass: the source is host memory behind PCIe
// (~2 us per request), so the bytes in flight decide the rate, not the thread count — few CTAs then read as fast
// as many and leave the SMs' register files to the solve kernels of other batches (scripts/e2e_depth.py).
constexpr int PACK_UNROLL = 8;
constexpr int PACK_THREADS = 512;
constexpr int PACK_VUNROLL = 4;

__global__ void __launch_bounds__(PACK_THREADS) cmpc_pack_records_kernel(const __grid_constant__ PackArgs A) {
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned t0 = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 1
  for (int s = 0; s < PACK_SEGS; s++) {
    const Seg sg = A.seg[s];
    if (sg.width == 0) continue;
    const unsigned width = (unsigned)sg.width;
    const unsigned total = (unsigned)A.count * width;  // < 2^31: the launcher cuts larger batches
    // 16-byte loads where the source allows: a warp then asks for 512 contiguous bytes per request
    unsigned done = 0;
    if (A.vec && sg.src && (reinterpret_cast<uintptr_t>(sg.src) & 15) == 0) {
      const unsigned nvec = total >> 2;
      const uint4* src4 = reinterpret_cast<const uint4*>(sg.src);
#pragma unroll 1
      for (unsigned base = t0; base < nvec; base += stride * PACK_VUNROLL) {
        uint4 v[PACK_VUNROLL];
#pragma unroll
        for (int u = 0; u < PACK_VUNROLL; u++) {
          const unsigned j = base + u * stride;
          v[u] = (j < nvec) ? __ldcs(src4 + j) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < PACK_VUNROLL; u++) {
          const unsigned j = base + u * stride;
          if (j < nvec) {
            unsigned inst = (4u * j) / width, k = 4u * j - inst * width;
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              A.records[(size_t)inst * A.rec_words + sg.dst + k] = w[e];
              if (++k == width) { k = 0; inst++; }
            }
          }
        }
      }
      done = nvec << 2;
    }
#pragma unroll 1
    for (unsigned base = done + t0; base < total; base += stride * PACK_UNROLL) {
      uint32_t v[PACK_UNROLL];
#pragma unroll
      for (int u = 0; u < PACK_UNROLL; u++) {
        const unsigned j = base + u * stride;
        v[u] = (sg.src && j < total) ? __ldcs(sg.src + j) : 0u;
      }
#pragma unroll
      for (int u = 0; u < PACK_UNROLL; u++) {
        const unsigned j = base + u * stride;
        if (j < total) {
          const unsigned inst = j / width, k = j - inst * width;
          A.records[(size_t)inst * A.rec_words + sg.dst + k] = v[u];
        }
      }
    }
  }
  // reserved words and the padding behind the gait bytes
  if (!A.zero_tail) return;
  const unsigned zw = 3 + A.tail_words;
  for (unsigned j = t0; j < (unsigned)A.count * zw; j += stride) {
    const unsigned inst = j / zw, k = j - inst * zw;
    const int dst = k < 3 ? CMPC_REC_SIMTIME + (int)k : A.tail_dst + ((int)k - 3);
    A.records[(size_t)inst * A.rec_words + dst] = 0u;
  }
}

}  // namespace

// in: the eleven arrays of cmpc_inputs as device-accessible pointers, already offset to the first instance
int cmpc_launch_pack(const void* p, const void* v, const void* q, const void* w, const void* r, const void* weights,
                     const void* traj, const void* alpha, const void* gait, const void* x_drag, const void* f_dist,
                     unsigned char* records, int rec_stride, int horizon, int count, int sm_count, void* stream, int which) {
  if (count <= 0) return 0;
  const int h = horizon;
  // 32-bit word indices inside the kernel: larger batches go out in pieces
  const int piece = (int)(0x7fffffffLL / (12LL * h + 64));
  if (count > piece) {
    auto adv = [&](const void* a, size_t bytes) -> const void* { return a ? static_cast<const char*>(a) + (size_t)piece * bytes : nullptr; };
    int rc = cmpc_launch_pack(p, v, q, w, r, weights, traj, alpha, gait, x_drag, f_dist, records, rec_stride, horizon, piece, sm_count, stream, which);
    if (rc) return rc;
    return cmpc_launch_pack(adv(p, 12), adv(v, 12), adv(q, 16), adv(w, 12), adv(r, 48), adv(weights, 48), adv(traj, 48 * (size_t)h),
                            adv(alpha, 4), adv(gait, 4 * (size_t)h), adv(x_drag, 4), adv(f_dist, 24),
                            records + (size_t)piece * rec_stride, rec_stride, horizon, count - piece, sm_count, stream, which);
  }
  PackArgs A;
  auto set = [&](int i, const void* src, int width, int dst) {
    A.seg[i].src = static_cast<const uint32_t*>(src);
    A.seg[i].width = width;
    A.seg[i].dst = dst;
  };
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
  if (which == CMPC_PACK_REST) A.seg[0].width = 0;
  if (which == CMPC_PACK_TRAJ)
    for (int i = 1; i < PACK_SEGS; i++) A.seg[i].width = 0;
  A.zero_tail = which != CMPC_PACK_TRAJ;
  A.records = reinterpret_cast<uint32_t*>(records);
  A.rec_words = rec_stride / 4;
  A.count = count;
  A.gait_words = h;
  A.tail_dst = CMPC_REC_TRAJ + 13 * h;
  A.tail_words = A.rec_words - A.tail_dst;
#ifdef CMPC_EXPERIMENTS
  static const int vec = [] { const char* e = std::getenv("CMPC_PACK_VEC"); return e ? std::atoi(e) : 1; }();
#else
  constexpr int vec = 1;
#endif
  A.vec = vec;
  const long long words = (long long)count * (12 * h);
  const long long per_cta = (long long)PACK_THREADS * PACK_UNROLL;
  int grid = (int)((words + per_cta - 1) / per_cta);
#ifdef CMPC_EXPERIMENTS
  static const int grid_cap = [] { const char* e = std::getenv("CMPC_PACK_GRID"); return e ? std::atoi(e) : 0; }();
#else
  constexpr int grid_cap = 0;
#endif
  const int cap = grid_cap > 0 ? grid_cap : (sm_count + 7) / 8;  // few, fat CTAs (measured: profiles/r1_s4_pack_grid.txt): they have to find room beside resident solve kernels
  if (which == CMPC_PACK_TRAJ) { if (grid > sm_count) grid = sm_count; }  // staged in HBM: latency is short, more CTAs finish sooner
  else if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  cmpc_pack_records_kernel<<<grid, PACK_THREADS, 0, (cudaStream_t)stream>>>(A);
  return (int)cudaGetLastError();
}
