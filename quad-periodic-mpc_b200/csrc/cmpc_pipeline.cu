// Kernel pipeline of the batched convex-MPC engine for reduced problems of up to 128 variables (every A1 gait at the
// reference's horizons except long all-feet-down stances).  Launchers and occupancy queries of:
//
//   n <= 63   cmpc_assemble_mma_kernel   one CTA per instance: condensation, [H g; g' .] as DMMA tiles   cmpc_condense_mma.cuh
//             cmpc_invert_ws_kernel      one main warp (+ helper warp) per instance: K = H^-1, x0 = -K g
//                                        on the FP64 tensor cores                                         cmpc_invert_mma.cuh
//             cmpc_lpt_order_kernel      hardest-first worklist for the active-set kernel                 (this file)
//   n <= 128  cmpc_condense_kernel       one CTA per instance: condensation + blocked DFMA sweep          cmpc_condense.cuh
//   all       cmpc_dual_fast_kernel      one warp per instance: Goldfarb-Idnani dual active set on K,
//                                        working sets of up to 32 rows, outputs                           cmpc_dual_fast.cuh
//             cmpc_dual_kernel           any working-set size; resumes what the fast tier handed over     cmpc_dual.cuh
//
// K travels between the kernels through a per-instance workspace slot (one workspace per stream).  Larger problems
// take the fused single-kernel path in cmpc_kernels.cu.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <cstdlib>

#include "cmpc_device.h"

namespace {
__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }
}  // namespace

#include "cmpc_common.cuh"
#include "cmpc_adapt.cuh"
#include "cmpc_condense.cuh"
#include "cmpc_condense_mma.cuh"
#include "cmpc_invert_mma.cuh"
#include "cmpc_dual.cuh"
#include "cmpc_dual_fast.cuh"
#include "cmpc_dual_team.cuh"

namespace {

// cudaFuncAttributeMaxDynamicSharedMemorySize of kernel K is raised when a launch (or an occupancy query) needs
// more than was set so far on the device, not on every launch
template <auto K>
struct SmemAttr {
  static size_t have[16];  // per device: bytes set + 1
};
template <auto K>
size_t SmemAttr<K>::have[16] = {};
template <auto K>
cudaError_t smem_attr(size_t smem) {
  int dev = 0;
  cudaGetDevice(&dev);
  size_t& have = SmemAttr<K>::have[dev & 15];
  if (have >= smem + 1) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) have = smem + 1;
  return e;
}

template <class S, bool ADAPT>
int launch_condense_t(const CmpcParams& P, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = smem_attr<cmpc_condense_kernel<S, ADAPT>>(smem);
  if (e != cudaSuccess) return (int)e;
  cmpc_condense_kernel<S, ADAPT><<<grid, S::NT, smem, st>>>(P);
  return (int)cudaGetLastError();
}
template <class S, bool ADAPT>
int occ_condense_t(size_t smem) {
  int nb = 0;
  if (smem_attr<cmpc_condense_kernel<S, ADAPT>>(smem) !=
      cudaSuccess)
    return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cmpc_condense_kernel<S, ADAPT>, S::NT, smem) != cudaSuccess)
    return -1;
  return nb;
}
template <bool ADAPT, int MINB>
int launch_mma_t(const CmpcParams& P, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = smem_attr<cmpc_assemble_mma_kernel<ADAPT, MINB>>(smem);
  if (e != cudaSuccess) return (int)e;
  cmpc_assemble_mma_kernel<ADAPT, MINB><<<grid, MMA_NT, smem, st>>>(P);
  return (int)cudaGetLastError();
}
template <bool ADAPT, int MINB>
int occ_mma_t(size_t smem) {
  int nb = 0;
  if (smem_attr<cmpc_assemble_mma_kernel<ADAPT, MINB>>(smem) != cudaSuccess)
    return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cmpc_assemble_mma_kernel<ADAPT, MINB>, MMA_NT, smem) !=
      cudaSuccess)
    return -1;
  return nb;
}
int cnpad_of(int cshape) {
  switch (cshape) {
    case CMPC_CSHAPE_64: return CShape64::NPAD;
    case CMPC_CSHAPE_96: return CShape96::NPAD;
    default: return CShape128::NPAD;
  }
}

template <int WPC>
int launch_dual_t(const CmpcParams& P, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = smem_attr<cmpc_dual_kernel<WPC>>(smem);
  if (e != cudaSuccess) return (int)e;
  cmpc_dual_kernel<WPC><<<grid, 32 * WPC, smem, st>>>(P);
  return (int)cudaGetLastError();
}
template <int WPC>
int occ_dual_t(size_t smem) {
  int nb = 0;
  if (smem_attr<cmpc_dual_kernel<WPC>>(smem) != cudaSuccess)
    return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cmpc_dual_kernel<WPC>, 32 * WPC, smem) != cudaSuccess) return -1;
  return nb;
}

}  // namespace

#define CMPC_CDISPATCH(FN, ...)                                                                              \
  switch (cshape) {                                                                                          \
    case CMPC_CSHAPE_64: return adapt ? FN<CShape64, true>(__VA_ARGS__) : FN<CShape64, false>(__VA_ARGS__);   \
    case CMPC_CSHAPE_96: return adapt ? FN<CShape96, true>(__VA_ARGS__) : FN<CShape96, false>(__VA_ARGS__);   \
    default: return adapt ? FN<CShape128, true>(__VA_ARGS__) : FN<CShape128, false>(__VA_ARGS__);             \
  }

size_t cmpc_condense_smem_bytes(int horizon, int nmax, int cshape, bool adapt) {
  if (cshape == CMPC_CSHAPE_MMA64) return (size_t)make_mcarve(horizon, cmpc_rec_stride(horizon), adapt).total;
  return (size_t)make_ccarve(horizon, nmax, cmpc_rec_stride(horizon), cnpad_of(cshape), adapt).total;
}

int cmpc_condense_max_ctas_per_sm(int cshape, size_t smem, bool adapt) {
  if (cshape == CMPC_CSHAPE_MMA64) return adapt ? occ_mma_t<true, 6>(smem) : occ_mma_t<false, 7>(smem);
  CMPC_CDISPATCH(occ_condense_t, smem)
}

int cmpc_condense_instances_per_cta(int cshape) { (void)cshape; return 1; }

int cmpc_launch_condense(const CmpcParams& P, int cshape, int grid, void* stream) {
  const bool adapt = P.adapt_mode >= 0;
  const size_t smem = cmpc_condense_smem_bytes(P.horizon, P.nmax, cshape, adapt);
  cudaStream_t st = (cudaStream_t)stream;
  if (cshape == CMPC_CSHAPE_MMA64)
    return adapt ? launch_mma_t<true, 6>(P, grid, smem, st) : launch_mma_t<false, 7>(P, grid, smem, st);
  CMPC_CDISPATCH(launch_condense_t, P, grid, smem, st)
}

size_t cmpc_dual_smem_bytes_per_warp(int nmax, int qcap) { return (size_t)make_dcarve(nmax, qcap).total; }

int cmpc_dual_max_ctas_per_sm(int wpc, size_t smem) {
  switch (wpc) {
    case 1: return occ_dual_t<1>(smem);
    case 2: return occ_dual_t<2>(smem);
    case 4: return occ_dual_t<4>(smem);
    default: return occ_dual_t<8>(smem);
  }
}

int cmpc_launch_dual(const CmpcParams& P, int wpc, int grid, void* stream) {
  const size_t smem = cmpc_dual_smem_bytes_per_warp(P.nmax, P.qcap) * (size_t)wpc;
  cudaStream_t st = (cudaStream_t)stream;
  switch (wpc) {
    case 1: return launch_dual_t<1>(P, grid, smem, st);
    case 2: return launch_dual_t<2>(P, grid, smem, st);
    case 4: return launch_dual_t<4>(P, grid, smem, st);
    default: return launch_dual_t<8>(P, grid, smem, st);
  }
}

// ---- inversion kernel of the n <= 63 path (cmpc_invert_mma.cuh): one warp per CTA; the register cap decides how
//      many instances an SM sub-partition holds (16384 registers each): 200 -> two, 168 -> three ----
#ifdef CMPC_EXPERIMENTS
namespace {
int inv_regs() {
  static int v = [] {
    const char* e = std::getenv("CMPC_INV_REGS");
    const int x = e ? std::atoi(e) : 168;
    return x >= 184 ? 200 : 168;
  }();
  return v;
}
template <int REGS, bool CLK>
int occ_invert_t() {
  int nb = 0;
  const size_t smem = (size_t)INV_WPC * INV_WARP_SMEM;
  if (smem_attr<cmpc_invert_mma_kernel<REGS, CLK>>(smem) != cudaSuccess) return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cmpc_invert_mma_kernel<REGS, CLK>, 32 * INV_WPC, smem) != cudaSuccess) return -1;
  return nb;
}
template <int REGS, bool CLK>
int launch_invert_t(const CmpcParams& P, int grid, cudaStream_t st) {
  const size_t smem = (size_t)INV_WPC * INV_WARP_SMEM;
  cudaError_t e = smem_attr<cmpc_invert_mma_kernel<REGS, CLK>>(smem);
  if (e != cudaSuccess) return (int)e;
  cmpc_invert_mma_kernel<REGS, CLK><<<grid, 32 * INV_WPC, smem, st>>>(P);
  return (int)cudaGetLastError();
}
}  // namespace
#endif
// hardest-first order of the active-set kernel: instance -> position from the key histogram the inversion kernel left
namespace {
__global__ void __launch_bounds__(256) cmpc_lpt_order_kernel(const int* __restrict__ hist, const int* __restrict__ key,
                                                             int* __restrict__ worklist, int count) {
  __shared__ int off[64];
  if (threadIdx.x < 64) {  // instances with a larger key come first
    int o = 0;
    for (int k = 63; k > (int)threadIdx.x; k--) o += hist[k];
    off[threadIdx.x] = o;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int kv = key[i];
  worklist[off[kv >> 24] + (kv & 0xffffff)] = i;
}
}  // namespace
int cmpc_launch_lpt_order(const int* hist, const int* key, int* worklist, int count, void* stream) {
  if (count <= 0) return 0;
  cmpc_lpt_order_kernel<<<(count + 255) / 256, 256, 0, (cudaStream_t)stream>>>(hist, key, worklist, count);
  return (int)cudaGetLastError();
}

// warp-specialised kernel: helper warps carry the pivot chains (an experiments build can select the single-warp
// kernel with CMPC_INV_WS=0)
namespace {
#ifdef CMPC_EXPERIMENTS
bool inv_ws() {
  static bool v = [] {
    const char* e = std::getenv("CMPC_INV_WS");
    return e ? std::atoi(e) != 0 : true;
  }();
  return v;
}
#else
constexpr bool inv_ws() { return true; }
#endif
}  // namespace
int cmpc_invert_max_ctas_per_sm(void) {
  if (inv_ws()) {
    int nb = 0;
    const size_t smem = (size_t)WS_MAIN * WS_PAIR_SMEM;
    if (smem_attr<cmpc_invert_ws_kernel<false>>(smem) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cmpc_invert_ws_kernel<false>, 64 * WS_MAIN, smem) != cudaSuccess) return -1;
#ifdef CMPC_EXPERIMENTS
    if (const char* e = std::getenv("CMPC_INV_CTAS_PER_SM")) nb = std::max(1, std::min(nb, std::atoi(e)));
#endif
    return nb;
  }
#ifdef CMPC_EXPERIMENTS
  return inv_regs() == 200 ? occ_invert_t<200, false>() : occ_invert_t<168, false>();
#else
  return -1;
#endif
}
int cmpc_invert_instances_per_cta(void) {
#ifdef CMPC_EXPERIMENTS
  if (!inv_ws()) return INV_WPC;
#endif
  return WS_MAIN;
}
int cmpc_launch_invert(const CmpcParams& P, int grid, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
#ifdef CMPC_EXPERIMENTS
  if (P.phase_cycles && !inv_ws()) return inv_regs() == 200 ? launch_invert_t<200, true>(P, grid, st) : launch_invert_t<168, true>(P, grid, st);
#endif
  if (P.phase_cycles) {
    const size_t smem = (size_t)WS_MAIN * WS_PAIR_SMEM;
    cudaError_t e = smem_attr<cmpc_invert_ws_kernel<true>>(smem);
    if (e != cudaSuccess) return (int)e;
    cmpc_invert_ws_kernel<true><<<grid, 64 * WS_MAIN, smem, st>>>(P);
    return (int)cudaGetLastError();
  }
  if (inv_ws()) {
    const size_t smem = (size_t)WS_MAIN * WS_PAIR_SMEM;
    cudaError_t e = smem_attr<cmpc_invert_ws_kernel<false>>(smem);
    if (e != cudaSuccess) return (int)e;
    cmpc_invert_ws_kernel<false><<<grid, 64 * WS_MAIN, smem, st>>>(P);
    return (int)cudaGetLastError();
  }
#ifdef CMPC_EXPERIMENTS
  return inv_regs() == 200 ? launch_invert_t<200, false>(P, grid, st) : launch_invert_t<168, false>(P, grid, st);
#else
  return (int)cudaErrorInvalidDeviceFunction;
#endif
}

// ---- CTA-per-instance tier of the dual active-set kernel (cmpc_dual_team.cuh): long working sets ----
size_t cmpc_dual_team_smem_bytes(int nmax, int qcap) { return (size_t)make_tcarve(nmax, qcap).total; }
namespace {
template <int NT>
int occ_team_t(size_t smem) {
  int nb = 0;
  if (smem_attr<cmpc_dual_team_kernel<NT>>(smem) != cudaSuccess) return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cmpc_dual_team_kernel<NT>, NT, smem) != cudaSuccess) return -1;
  return nb;
}
template <int NT>
int launch_team_t(const CmpcParams& P, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = smem_attr<cmpc_dual_team_kernel<NT>>(smem);
  if (e != cudaSuccess) return (int)e;
  cmpc_dual_team_kernel<NT><<<grid, NT, smem, st>>>(P);
  return (int)cudaGetLastError();
}
}  // namespace
// four warps per instance up to 64 variables, eight beyond (one variable per thread in the z update either way)
int cmpc_dual_team_max_ctas_per_sm(int nmax, size_t smem) { return nmax <= 64 ? occ_team_t<128>(smem) : occ_team_t<256>(smem); }
int cmpc_launch_dual_team(const CmpcParams& P, int grid, void* stream) {
  const size_t smem = cmpc_dual_team_smem_bytes(P.nmax, P.qcap);
  return P.nmax <= 64 ? launch_team_t<128>(P, grid, smem, (cudaStream_t)stream) : launch_team_t<256>(P, grid, smem, (cudaStream_t)stream);
}

// ---- fast tier of the dual active-set kernel (cmpc_dual_fast.cuh): working sets of up to 32 rows ----
namespace {
template <int NPL, int MPL>
int launch_fast_t(const CmpcParams& P, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = smem_attr<cmpc_dual_fast_kernel<1, NPL, MPL>>(smem);
  if (e != cudaSuccess) return (int)e;
  cmpc_dual_fast_kernel<1, NPL, MPL><<<grid, 32, smem, st>>>(P);
  return (int)cudaGetLastError();
}
template <int NPL, int MPL>
int occ_fast_t(size_t smem) {
  int nb = 0;
  if (smem_attr<cmpc_dual_fast_kernel<1, NPL, MPL>>(smem) != cudaSuccess)
    return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cmpc_dual_fast_kernel<1, NPL, MPL>, 32, smem) != cudaSuccess) return -1;
  return nb;
}
}  // namespace

size_t cmpc_dual_fast_smem_bytes(int nmax, int qcap) { return (size_t)make_fcarve(nmax <= 64 ? 2 : 4, qcap).total; }
int cmpc_dual_fast_max_ctas_per_sm(int nmax, size_t smem) { return nmax <= 64 ? occ_fast_t<2, 4>(smem) : occ_fast_t<4, 7>(smem); }
int cmpc_launch_dual_fast(const CmpcParams& P, int grid, void* stream) {
  const size_t smem = cmpc_dual_fast_smem_bytes(P.nmax, P.qcap);
  return P.nmax <= 64 ? launch_fast_t<2, 4>(P, grid, smem, (cudaStream_t)stream) : launch_fast_t<4, 7>(P, grid, smem, (cudaStream_t)stream);
}

