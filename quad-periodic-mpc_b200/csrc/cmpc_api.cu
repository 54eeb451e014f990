// Host side of the C-ABI in include/cmpc_b200.h: the batched engine
// (cmpc_batch_*) and, on top of it, the reference's single-instance interface
// (convexMPC_interface.cpp:44-162) symbol for symbol.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "cmpc_device.h"

namespace {

thread_local std::string g_err;

int fail_cuda(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  // the failure has been reported: do not leave it in the runtime's last-error slot, where the next kernel-launch check
  // (cudaGetLastError) would find it and blame the launch — e.g. cmpc_host_register on memory that is pinned already
  (void)cudaGetLastError();
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorNoKernelImageForDevice ||
          e == cudaErrorInvalidDeviceFunction)
             ? CMPC_E_NODEVICE
             : CMPC_E_CUDA;
}
#define CK(call)                                   \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
  } while (0)

int fail_arg(const char* what) {
  g_err = what;
  return CMPC_E_ARG;
}

// horizon sums of the discretisation polynomials, DESIGN.md §3:
//   c1(k) = dt, c2(k) = k dt^2 + dt^2/2, c3(k) = (k dt)^2 dt/2 + k dt^3/2 + dt^3/6
//   sig_xy[a][b] = sum_{r=max(a,b)}^{h-1} cx(r-a) cy(r-b)
void build_sigma(int h, double dt, std::vector<double>& sig) {
  std::vector<double> c1(h), c2(h), c3(h);
  for (int k = 0; k < h; k++) {
    double tau = k * dt;
    c1[k] = dt;
    c2[k] = tau * dt + 0.5 * dt * dt;
    c3[k] = 0.5 * tau * tau * dt + 0.5 * tau * dt * dt + dt * dt * dt / 6.0;
  }
  sig.assign((size_t)CMPC_SIG_COUNT * h * h, 0.0);
  auto fill = [&](int t, const std::vector<double>& x, const std::vector<double>& y) {
    for (int a = 0; a < h; a++)
      for (int b = 0; b < h; b++) {
        double s = 0;
        for (int r = std::max(a, b); r < h; r++) s += x[r - a] * y[r - b];
        sig[(size_t)t * h * h + a * h + b] = s;
      }
  };
  fill(CMPC_SIG_11, c1, c1);
  fill(CMPC_SIG_22, c2, c2);
  fill(CMPC_SIG_33, c3, c3);
  fill(CMPC_SIG_23, c2, c3);
  fill(CMPC_SIG_12, c1, c2);
}

// Estimator tables: DFT twiddles for N = 400 and the two normalised Gaussian
// kernels of gaussian_filter(), SolverMPC.cpp:404-419 (float kernel, float sum).
void build_adapt_tables(std::vector<double>& tw, std::vector<float>& gk) {
  const int N = CMPC_ADAPT_WINDOW;
  tw.resize(2 * N);
  for (int m = 0; m < N; m++) {
    tw[2 * m] = std::cos(2.0 * M_PI * m / N);
    tw[2 * m + 1] = -std::sin(2.0 * M_PI * m / N);  // e^{-2 pi i m / N}
  }
  gk.assign(CMPC_GK_TOTAL, 0.f);
  const float sigmas[2] = {7.0f, 27.0f};
  const int offs[2] = {0, 2 * CMPC_GK_R1 + 1};
  for (int s = 0; s < 2; s++) {
    float sigma = sigmas[s];
    int radius = (int)std::ceil(3 * sigma);
    float sum = 0.0f;
    for (int i = -radius; i <= radius; i++) {
      float v = (float)std::exp(-0.5 * (i * i) / (sigma * sigma));
      gk[offs[s] + i + radius] = v;
      sum += v;
    }
    for (int i = 0; i < 2 * radius + 1; i++) gk[offs[s] + i] /= sum;
  }
}

constexpr int kMaxChunks = 8;
constexpr int kMaxStreams = 8;
constexpr int kRstateStride = 32 * 32;                                   // P of a 32-row working set, full rows
constexpr int kRstate2Stride = CMPC_QCAP_TEAM * (CMPC_QCAP_TEAM + 1) / 2;  // P of a team- / middle-tier working set, packed
// control block of one chunk of a pipeline launch, zeroed by ONE memset per launch: work counters of the kernels,
// overflow counts of the two capacity hand-overs, the hardest-first histogram, the SM arrival counters of the stagger
constexpr int kCtlSched = 0, kCtlOvf = 4, kCtlOvf2 = 5, kCtlHist = 8, kCtlSlots = 8 + 64, kCtlInts = 8 + 64 + CMPC_SM_SLOTS;
constexpr int kPackStream = 2;  // end-to-end call: the record-packing kernels of all chunks, in order

// Small persistent worker pool for the host side of the end-to-end call (record packing, result unpacking):
// parallel_for(n, fn) runs fn(i) for i in [0, n) on the workers and the calling thread.
class HostPool {
 public:
  explicit HostPool(int workers) {
    for (int i = 0; i < workers; i++) threads_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  int size() const { return (int)threads_.size() + 1; }
  void parallel_for(int n, const std::function<void(int)>& fn) {
    if (n <= 0) return;
    if (threads_.empty() || n == 1) {
      for (int i = 0; i < n; i++) fn(i);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      n_ = n;
      next_.store(0);
      pending_ = (int)threads_.size();
      gen_++;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void work() {
    for (int i = next_.fetch_add(1); i < n_; i = next_.fetch_add(1)) (*fn_)(i);
  }
  void loop() {
    unsigned long long seen = 0;
    while (true) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      work();
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  std::atomic<int> next_{0};
  int n_ = 0, pending_ = 0;
  unsigned long long gen_ = 0;
  bool stop_ = false;
};

}  // namespace

namespace {
// Launch plan of the pipeline for one (reduced size bound, horizon, adaptive) combination: kernel shapes, shared
// memory, CTAs per SM.  The occupancy queries behind it cost tens of microseconds, so a batch keeps its plans.
struct PipePlan {
  int nmax = 0, h = 0, adapt = 0, qcap_pref = 0;
  int cshape = 0, tiled = 0;
  size_t slot = 0;
  int chunk_cap = 0;  // instances whose workspace stays L2-sized
  int per_sm1 = 0, per_sm_inv = 0;
  int qcap1 = 0, fast = 0, per_sm_fast = 0, wpc1 = 0, per_sm2 = 0, wpc2 = 0, per_sm3 = 0;
  int qcap_mid = 0, wpc_mid = 0, per_sm_mid = 0;  // middle capacity tier (0: none)
  int qcap_team = 0, per_sm_team = 0;             // CTA-per-instance tier (0: none)
};

}  // namespace

namespace {
struct SoaView {
  const char* p[11];
  bool ok;
};

// What the end-to-end call needs to know about the caller's arrays; resolving it costs a cudaPointerGetAttributes
// per array, so cmpc_batch_bind_host keeps it across calls.
struct HostBinding {
  cmpc_inputs in;
  cmpc_outputs out;
  SoaView soa;     // device views of the input arrays (ok: every one pinned)
  int direct = 0;  // output arrays that are pinned (kOut* bits)
  void* vo[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // device views the kernels write outputs through (zero-copy)
  bool zc_out = true;
  bool valid = false;
};

}  // namespace

// Diagnostic / test switches of a batch (cmpc_batch_set_option).  Nothing here is read from the environment.
struct Knobs {
  int cshape = -1;        // "cshape": force a condensation shape that still fits (CMPC_CSHAPE_*)
  int ws_mb = 160;        // "ws_mb": workspace budget of one pipeline chunk
  int qcap1 = 0;          // "qcap1": first-tier working-set capacity (0: 32 / 24 by mode)
  int dual_generic = 0;   // "dual_generic": first tier on the any-capacity kernel
  int wpc = 4;            // "wpc": warps per CTA of the any-capacity kernel's first tier
  int no_mid_tier = 0;    // "no_mid_tier"
  int path_fused = 0;     // "path_fused": single-kernel path for every size
  int shape = -1;         // "shape": kernel shape of the single-kernel path (CMPC_SHAPE_*)
  int host_pack = 0;      // "host_pack": pack records on the host even when the inputs are pinned
  int d2h_copy = 0;       // "d2h_copy": device-side outputs + copy engine instead of zero-copy stores
  int chunks = 0;         // "chunks": chunks of the end-to-end call (0: by size)
  int submit_copy = 0;    // "submit_copy": submit / wait hand the results to the copy engine
  int host_threads = 0;   // "host_threads": host workers of the end-to-end call (0: hardware)
  int resume_p = 1;       // "resume_p": a working set handed to the next tier travels with its P (0: row ids only, P is bordered again)
  int dual_team = 1;      // "dual_team": working sets beyond the first tier on the CTA-per-instance kernel (0: one warp per instance)
};

struct cmpc_batch {
  int device = 0;
  Knobs knobs;
  int capacity = 0;
  int sm_count = 0;
  cudaStream_t stream[kMaxStreams] = {};  // [0] is "the batch stream"; the others only carry pipelined chunks / solves
  int nstreams = kMaxStreams;             // streams that successive solve_range calls rotate through (CMPC_NSTREAMS); measured: profiles/
  int split = 1;                          // parts a solve_range call is cut into, one per stream in turn (CMPC_SPLIT)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, mark0 = nullptr, mark1 = nullptr;
  cudaEvent_t chunk_done[kMaxChunks] = {};
  cudaEvent_t packed[kMaxChunks] = {};    // end-to-end call: chunk c's records are in HBM
  cudaStream_t copy_stream = nullptr;     // end-to-end call: the copy engine brings the reference trajectories (71 % of the input bytes) ...
  cudaEvent_t traj_copied[kMaxChunks] = {};
  float* d_traj_stage = nullptr;          // ... here, while the packing kernel reads the ten small arrays over PCIe itself
  bool traj_copy = true;                  // CMPC_TRAJ_COPY=0: the packing kernel reads every array itself
#ifdef CMPC_EXPERIMENTS
  bool exp_skip_pack = false, exp_packed_once = false;  // "exp_skip_pack": the records of the first call are reused
#endif
  cudaEvent_t rec_copied = nullptr;       // the last copy out of the pinned record staging (h_rec) has finished
  bool rec_copy_pending = false;
  // successive solve_range calls rotate through the streams so that the latency-bound tails of a batch's kernels
  // overlap the kernels of the following batches; these events carry the cross-stream ordering
  cudaEvent_t join_ev = nullptr;      // scratch: "stream i has reached this point"
  cudaEvent_t fork_ev = nullptr;      // scratch: "stream 0 has reached this point"
  cudaEvent_t prof_ev[CMPC_K_COUNT + 1] = {};  // cmpc_batch_profile_range: events between the kernel classes
  float prof_ms[CMPC_K_COUNT] = {};
  bool profiling = false;
  unsigned rr = 0;                    // round-robin counter of solve_range
  bool serial = false;                // CMPC_SERIAL=1: keep every solve on the batch stream
  HostPool* pool = nullptr;           // host workers of the end-to-end call, created on first use
  unsigned dirty = 0;                 // bit i: stream i carries work that stream 0 has not waited for yet
  // problem setup
  bool is_setup = false;
  int h = 0;
  int rec_stride = 0;
  double dt = 0, mu = 0, f_max = 0;
  double mass = 12.0;
  double inertia[3] = {0.07, 0.26, 0.242};
  // buffers (sized for capacity x CMPC_MAX_HORIZON so setup never reallocates)
  unsigned char* h_rec = nullptr;  // pinned
  unsigned char* d_rec = nullptr;
  double* d_sigma = nullptr;
  double* d_forces = nullptr;
  double* d_obj = nullptr;
  int* d_status = nullptr;
  int* d_iters = nullptr;
  signed char* d_active = nullptr;
  int* d_overflow[kMaxStreams] = {};  // per stream: [capacity] list + [1] count at the end
  unsigned long long* d_flops = nullptr;
  unsigned long long* d_phase = nullptr;  // CMPC_PH_COUNT phase clocks, allocated by cmpc_batch_enable_phase_clocks
  double* d_gws = nullptr;  // global workspace of the large-problem tier, grown on demand
  size_t gws_bytes = 0;
  // two-kernel pipeline: per-stream workspace slots (K, g, x0 per instance of a chunk) and work counters
  double* d_qws[kMaxStreams] = {};
  size_t qws_bytes[kMaxStreams] = {};
  int* d_sched[kMaxStreams] = {};     // per stream: control blocks (kCtlInts ints per chunk of a pipeline launch)
  int sched_ints[kMaxStreams] = {};
  int* d_resume[kMaxStreams] = {};    // per stream: working sets of the instances in d_overflow, for the next capacity tier
  int* d_overflow2[kMaxStreams] = {}; // per stream: second overflow list (middle tier -> full capacity) + count
  int* d_resume2[kMaxStreams] = {};
  double* d_rstate[kMaxStreams] = {};   // per stream: P of the overflowed working sets, first tier -> next ([rstate_cap][1024])
  double* d_rstate2[kMaxStreams] = {};  // middle tier -> full capacity ([rstate2_cap][1600], packed)
  int rstate_cap = 0, rstate2_cap = 0;
  int* d_lpt[kMaxStreams] = {};       // per stream: [capacity] hardest-first keys, [capacity] worklist
  bool lpt = true;                    // CMPC_LPT=0: natural instance order in the active-set kernel
  bool throughput_mode = false;       // set by solve_range (batches pipelined over the streams), cleared by the end-to-end calls
  bool sweep_dmma = false;            // CMPC_SWEEP=dmma: tensor-core sweep in the 96 / 128 condensation shapes (measured:
                                      // +5 % on mixed gaits at h = 16, rounding error 50x the DFMA sweep's -> not the default)
  int inv_stagger = 2000;             // start offset (cycles) between the inversion CTAs of an SM (CMPC_INV_STAGGER, 0 = off)
  int inv_ctas = 0;                   // cap on the inversion kernel's CTAs per SM (0 = what fits; measurements)
  int inv_f32 = 1;                    // pivot blocks of the inversion kernel: fp32 chain + FP64 Newton steps on the tensor cores (0: FP64 chain)
  int inv_refine = 1024;              // blocked sweeps: refine the panel of block steps whose pivot-block inverse exceeds this; -1 = never
                                      // (measured, profiles/r2_illcond_accuracy.txt: 1024 costs nothing on the A1 trot batch and 1 % on
                                      // the mixed-gait h = 16 batch; 512 costs 8 % there)
  bool resume = true;                 // CMPC_RESUME=0: overflowed instances restart from scratch in the full-capacity launch
  std::vector<PipePlan> plans;        // launch plans by (reduced size bound, horizon, adaptive)
  HostBinding bound;                  // cmpc_batch_bind_host
  int pend_count = 0, pend_per = 0, pend_used = -1;  // cmpc_batch_submit_bound: chunks in flight (-1: none)
  // adaptive stage
  double* d_twiddle = nullptr;
  float* d_gk = nullptr;
  float* d_win_t = nullptr;
  float* d_win_d = nullptr;
  float* d_simtime = nullptr;
  double* d_est = nullptr;
  float* d_fest = nullptr;
  // command front end (cmpc_batch_solve_commands)
  unsigned char* d_cmds = nullptr;     // [capacity] cmpc_command
  unsigned char* d_results = nullptr;  // [capacity] cmpc_command_result
  float* d_fext = nullptr;             // [capacity][6] the reference's global f_ext, per instance
  int hist_len = 0;                    // samples pushed into the disturbance histories (time_history.size())
  int hist_count = 0;                  // instances those histories were built for
  float weights[12] = {0.25f, 0.25f, 10.f, 10.f, 2.f, 50.f, 0.f, 0.f, 0.3f, 0.2f, 0.2f, 0.1f};  // ConvexMPCLocomotion.cpp:627
  float alpha = 4e-5f;                 // :634
  // pinned result staging
  double* h_forces = nullptr;
  double* h_obj = nullptr;
  int* h_status = nullptr;
  int* h_iters = nullptr;
  signed char* h_active = nullptr;
  unsigned long long* h_flops = nullptr;
  // end-to-end call: device views of the (pinned) host arrays the kernels write their outputs to, instead of
  // the device arrays above; null = the device array
  double* o_forces = nullptr;
  double* o_obj = nullptr;
  int* o_status = nullptr;
  int* o_iters = nullptr;
  signed char* o_active = nullptr;
  // state
  int count = 0;
  int max_contact = 0;  // max contact foot-steps over the uploaded instances
  int adapt_mode = -1;
  long long launches = 0;
  bool timed = false;
};

namespace {

// pack instances [first, first+count) into the pinned records; returns the max contact foot-steps
int pack_records(cmpc_batch* b, const cmpc_inputs* in, int first, int count) {
  const int h = b->h, stride = b->rec_stride;
  const double fmax = (double)(float)b->f_max;
  int maxc = 0;
  for (int i = first; i < first + count; i++) {
    unsigned char* rec = b->h_rec + (size_t)i * stride;
    float* f = reinterpret_cast<float*>(rec);
    std::memcpy(f + CMPC_REC_P, in->p + 3 * (size_t)i, 12);
    std::memcpy(f + CMPC_REC_V, in->v + 3 * (size_t)i, 12);
    std::memcpy(f + CMPC_REC_Q, in->q + 4 * (size_t)i, 16);
    std::memcpy(f + CMPC_REC_W, in->w + 3 * (size_t)i, 12);
    std::memcpy(f + CMPC_REC_R, in->r + 12 * (size_t)i, 48);
    std::memcpy(f + CMPC_REC_WEIGHTS, in->weights + 12 * (size_t)i, 48);
    f[CMPC_REC_ALPHA] = in->alpha[i];
    f[CMPC_REC_XDRAG] = in->x_drag[i];
    if (in->f_dist) std::memcpy(f + CMPC_REC_FDIST, in->f_dist + 6 * (size_t)i, 24);
    else std::memset(f + CMPC_REC_FDIST, 0, 24);
    f[CMPC_REC_SIMTIME] = 0.f;
    f[CMPC_REC_RSV] = 0.f;
    f[CMPC_REC_RSV + 1] = 0.f;
    std::memcpy(f + CMPC_REC_TRAJ, in->traj + 12 * (size_t)h * i, 48 * (size_t)h);
    unsigned char* gz = rec + 4 * (CMPC_REC_TRAJ + 12 * h);
    const unsigned char* gsrc = in->gait + 4 * (size_t)h * i;
    int c = 0;
    for (int k = 0; k < 4 * h; k++) {
      gz[k] = gsrc[k];
      double ub = (double)gsrc[k] * fmax;
      c += !(ub < 0.01 && ub > -0.01);
    }
    for (int k = 4 * h; k < stride - 4 * (CMPC_REC_TRAJ + 12 * h); k++) gz[k] = 0;
    maxc = std::max(maxc, c);
  }
  return maxc;
}

HostPool* host_pool(cmpc_batch* b) {
  if (!b->pool) {
    int workers = (int)std::thread::hardware_concurrency() - 1;
    if (b->knobs.host_threads > 0) workers = b->knobs.host_threads - 1;
    b->pool = new HostPool(std::max(0, std::min(workers, 7)));
  }
  return b->pool;
}

// pack_records over the pool: blocks of 256 instances
int pack_records_parallel(cmpc_batch* b, const cmpc_inputs* in, int first, int count) {
  if (count < 8192) return pack_records(b, in, first, count);  // below this the hand-off costs more than it saves
  const int blk = 256, nb = (count + blk - 1) / blk;
  std::vector<int> maxc(nb, 0);
  host_pool(b)->parallel_for(nb, [&](int i) {
    const int f = first + i * blk, n = std::min(blk, first + count - f);
    maxc[i] = pack_records(b, in, f, n);
  });
  return *std::max_element(maxc.begin(), maxc.end());
}

bool is_pinned(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// Device-side view of a pinned (cudaHostAlloc / cudaHostRegister) host array, or nullptr if the array is pageable.
const void* device_view(const void* p) {
  if (!p) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (a.type != cudaMemoryTypeHost) return nullptr;
  return a.devicePointer;
}

// The eleven input arrays as device-accessible pointers when every one of them is pinned (and word aligned):
// the end-to-end call then packs the records on the device (cmpc_pack.cu) instead of on the host.
SoaView soa_view(const cmpc_inputs* in) {
  SoaView s;
  const void* src[11] = {in->p, in->v, in->q, in->w, in->r, in->weights, in->traj, in->alpha, in->gait, in->x_drag, in->f_dist};
  s.ok = true;
  for (int i = 0; i < 11; i++) {
    s.p[i] = static_cast<const char*>(device_view(src[i]));
    if (src[i] && (!s.p[i] || (reinterpret_cast<uintptr_t>(s.p[i]) & 3u))) s.ok = false;
  }
  return s;
}

// max contact foot-steps over instances [first, first+n) (what pack_records returns), from the gait bytes alone
int max_contact_scan(const cmpc_batch* b, const uint8_t* gait, int first, int n) {
  const int h4 = 4 * b->h;
  const double fmax = (double)(float)b->f_max;
  bool keep[256];
  bool plain = true;  // keep[v] == (v != 0): the usual case (f_max >= 0.01), counted eight bytes at a time
  for (int v = 0; v < 256; v++) {
    const double ub = (double)v * fmax;
    keep[v] = !(ub < 0.01 && ub > -0.01);
    plain = plain && (keep[v] == (v != 0));
  }
  int maxc = 0;
  for (int i = first; i < first + n; i++) {
    const uint8_t* g = gait + (size_t)i * h4;
    int c = 0, k = 0;
    if (plain) {
      for (; k + 8 <= h4; k += 8) {
        uint64_t x;
        std::memcpy(&x, g + k, 8);
        // bit 7 of every non-zero byte
        x = ((x & 0x7f7f7f7f7f7f7f7full) + 0x7f7f7f7f7f7f7f7full) | x;
        c += (int)((((x >> 7) & 0x0101010101010101ull) * 0x0101010101010101ull) >> 56);  // byte-wise sum of the flags
      }
    }
    for (; k < h4; k++) c += keep[g[k]];
    maxc = std::max(maxc, c);
  }
  return maxc;
}

int check_inputs(const cmpc_batch* b, int count, const cmpc_inputs* in, const char* who) {
  if (!b || !in) return fail_arg("null argument");
  if (!b->is_setup) { g_err = std::string(who) + ": call cmpc_batch_setup first"; return CMPC_E_STATE; }
  if (count < 0 || count > b->capacity) return fail_arg("count exceeds capacity");
  if (!in->p || !in->v || !in->q || !in->w || !in->r || !in->weights || !in->traj || !in->alpha || !in->gait ||
      !in->x_drag)
    return fail_arg("null input array");
  return CMPC_OK;
}

// the host waits until the last asynchronous copy out of the pinned record staging has run
int wait_rec_staging(cmpc_batch* b) {
  if (!b->rec_copy_pending) return CMPC_OK;
  CK(cudaEventSynchronize(b->rec_copied));
  b->rec_copy_pending = false;
  return CMPC_OK;
}
// stream 0 ("the batch stream") waits for everything enqueued on stream 1 so far
int join_streams(cmpc_batch* b) {
  for (int i = 1; i < kMaxStreams; i++) {
    if (!((b->dirty >> i) & 1u)) continue;
    CK(cudaEventRecord(b->join_ev, b->stream[i]));
    CK(cudaStreamWaitEvent(b->stream[0], b->join_ev, 0));
  }
  b->dirty = 0;
  return CMPC_OK;
}
// the host waits for every stream
int sync_all(cmpc_batch* b) {
  for (int i = 0; i < kMaxStreams; i++)
    if (b->stream[i]) CK(cudaStreamSynchronize(b->stream[i]));
  b->dirty = 0;
  return CMPC_OK;
}
// the host waits for the auxiliary streams (the batch stream keeps running)
int sync_aux(cmpc_batch* b) {
  for (int i = 1; i < kMaxStreams; i++)
    if ((b->dirty >> i) & 1u) CK(cudaStreamSynchronize(b->stream[i]));
  b->dirty = 0;
  return CMPC_OK;
}
// the auxiliary streams wait for everything enqueued on stream 0 so far (uploads, marks)
int fork_streams(cmpc_batch* b, int only = -1) {
  CK(cudaEventRecord(b->fork_ev, b->stream[0]));
  for (int i = 1; i < kMaxStreams; i++)
    if (only < 0 ? i < b->nstreams : i == only) CK(cudaStreamWaitEvent(b->stream[i], b->fork_ev, 0));
  return CMPC_OK;
}

// Two-kernel pipeline (cmpc_pipeline.cu) over the instances P describes, in chunks whose workspace stays
// L2-sized: condensation + K = H^-1 (one CTA per instance), then the dual active set (one warp per
// instance) with a small working-set capacity, then the few instances that outgrew it at full capacity.
int make_pipe_plan(const Knobs& kn, const CmpcParams& P, int qcap_pref, PipePlan& pl) {
  const bool adapt = P.adapt_mode >= 0;
  const int nmax = P.nmax;
  pl.nmax = nmax;
  pl.h = P.horizon;
  pl.adapt = adapt;
  int cshape = nmax < 64 ? CMPC_CSHAPE_MMA64 : (nmax <= 96 ? CMPC_CSHAPE_96 : CMPC_CSHAPE_128);
  if (kn.cshape >= 0) {  // forced shape, if it still fits
    const int sh = kn.cshape;
    if ((sh == CMPC_CSHAPE_64 && nmax <= 64) || (sh == CMPC_CSHAPE_96 && nmax <= 96) || sh == CMPC_CSHAPE_128) cshape = sh;
  }
  pl.cshape = cshape;
  pl.tiled = (cshape == CMPC_CSHAPE_MMA64) ? 1 : 0;
  pl.slot = cmpc_qws_slot_doubles(nmax, pl.tiled);
  size_t budget = (size_t)std::max(1, kn.ws_mb);
  budget <<= 20;
  pl.chunk_cap = (int)std::max<size_t>(1, budget / (pl.slot * sizeof(double)));
  // kernel 1
  const size_t smem1 = cmpc_condense_smem_bytes(P.horizon, nmax, cshape, adapt);
  pl.per_sm1 = cmpc_condense_max_ctas_per_sm(cshape, smem1, adapt);
  if (pl.per_sm1 < 1) {
    g_err = "cmpc_batch_solve: condensation kernel not launchable on this device (no sm_100a image?)";
    return CMPC_E_NODEVICE;
  }
  pl.per_sm_inv = pl.tiled ? cmpc_invert_max_ctas_per_sm() : 1;
  if (pl.per_sm_inv < 1) {
    g_err = "cmpc_batch_solve: inversion kernel not launchable on this device (no sm_100a image?)";
    return CMPC_E_NODEVICE;
  }
  // kernel 2, two working-set capacity tiers
  int qcap1 = kn.qcap1 > 0 ? kn.qcap1 : qcap_pref;
  if (qcap1 < 1 || qcap1 > nmax) qcap1 = nmax;
  bool fast = qcap1 <= 32 && !kn.dual_generic;  // tier 1 on the register-resident kernel (cmpc_dual_fast.cuh)
  if (fast) {
    pl.per_sm_fast = cmpc_dual_fast_max_ctas_per_sm(nmax, cmpc_dual_fast_smem_bytes(nmax, qcap1));
    if (pl.per_sm_fast < 1) fast = false;
  }
  int wpc1 = kn.wpc;
  if (wpc1 != 1 && wpc1 != 2 && wpc1 != 4 && wpc1 != 8) wpc1 = 4;
  const size_t smax = 227 * 1024;
  while (wpc1 > 1 && cmpc_dual_smem_bytes_per_warp(nmax, qcap1) * wpc1 > smax) wpc1 >>= 1;
  pl.per_sm2 = cmpc_dual_max_ctas_per_sm(wpc1, cmpc_dual_smem_bytes_per_warp(nmax, qcap1) * wpc1);
  int wpc2 = 4;
  while (wpc2 > 1 && cmpc_dual_smem_bytes_per_warp(nmax, nmax) * wpc2 > 100 * 1024) wpc2 >>= 1;
  pl.per_sm3 = cmpc_dual_max_ctas_per_sm(wpc2, cmpc_dual_smem_bytes_per_warp(nmax, nmax) * wpc2);
  if (pl.per_sm2 < 1 || pl.per_sm3 < 1) {
    g_err = "cmpc_batch_solve: active-set kernel not launchable on this device (no sm_100a image?)";
    return CMPC_E_NODEVICE;
  }
  // a middle tier for the larger reduced problems: at full capacity one warp's K N and P fill an SM's shared memory
  pl.qcap_mid = 0;
  if (nmax > 64 && qcap1 < CMPC_QCAP_MID && !kn.no_mid_tier) {
    int wm = 4;
    while (wm > 1 && cmpc_dual_smem_bytes_per_warp(nmax, CMPC_QCAP_MID) * wm > 100 * 1024) wm >>= 1;
    const int pm = cmpc_dual_max_ctas_per_sm(wm, cmpc_dual_smem_bytes_per_warp(nmax, CMPC_QCAP_MID) * wm);
    if (pm >= 1) {
      pl.qcap_mid = CMPC_QCAP_MID;
      pl.wpc_mid = wm;
      pl.per_sm_mid = pm;
    }
  }
  // the CTA-per-instance tier takes over what outgrows the first tier (it resumes from the working set and P handed over)
  pl.qcap_team = 0;
  if (kn.dual_team && fast && qcap1 < nmax) {
    const int qt = std::min(nmax, CMPC_QCAP_TEAM);
    const int pt = cmpc_dual_team_max_ctas_per_sm(nmax, cmpc_dual_team_smem_bytes(nmax, qt));
    if (pt >= 1) {
      pl.qcap_team = qt;
      pl.per_sm_team = pt;
      pl.qcap_mid = 0;
    }
  }
  pl.qcap1 = qcap1;
  pl.fast = fast ? 1 : 0;
  pl.wpc1 = wpc1;
  pl.wpc2 = wpc2;
  return CMPC_OK;
}

int launch_pipeline(cmpc_batch* b, CmpcParams P, int count, int si) {
  cudaStream_t st = b->stream[si];
  const bool adapt = P.adapt_mode >= 0;
  const int nmax = P.nmax;
  const PipePlan* plp = nullptr;
  // first-tier working-set capacity: 32 rows for one batch at a time (the shortest active-set kernel), 24 when
  // batches are pipelined over the streams (ten instead of seven warps per SM; the instances beyond 24 rows — ~1 % of
  // a trot batch, most of a mixed-gait h = 16 batch — are resumed by the CTA-per-instance tier, off the critical path
  // of the following batches) — measured, profiles/r2_qcap1_sweep.txt
  const int qcap_pref = b->throughput_mode ? 24 : 32;
  for (const PipePlan& c : b->plans)
    if (c.nmax == nmax && c.h == P.horizon && c.adapt == (int)adapt && c.qcap_pref == qcap_pref) plp = &c;
  if (!plp) {
    PipePlan pl;
    pl.qcap_pref = qcap_pref;
    if (int e = make_pipe_plan(b->knobs, P, qcap_pref, pl)) return e;
    b->plans.push_back(pl);
    plp = &b->plans.back();
  }
  const PipePlan& pl = *plp;
  const int cshape = pl.cshape, tiled = pl.tiled;
  const size_t slot = pl.slot;
  const int per_sm1 = pl.per_sm1, per_sm_inv = pl.per_sm_inv, qcap1 = pl.qcap1, per_sm_fast = pl.per_sm_fast;
  const int wpc1 = pl.wpc1, per_sm2 = pl.per_sm2, wpc2 = pl.wpc2, per_sm3 = pl.per_sm3;
  const bool fast = pl.fast != 0;
  const int chunk = std::min(count, pl.chunk_cap);
  const int nchunks = (count + chunk - 1) / chunk;
  const size_t need = (size_t)chunk * slot * sizeof(double);
  if (need > b->qws_bytes[si] || kCtlInts * nchunks > b->sched_ints[si]) {
    // grow the workspaces of every stream in use at once: the first solves of the other streams then find theirs
    { int rcs = sync_all(b); if (rcs) return rcs; }
    for (int k = 0; k < kMaxStreams; k++) {
      if (k != si && !(k < b->nstreams || k < 2)) continue;
      if (need > b->qws_bytes[k]) {
        if (b->d_qws[k]) CK(cudaFree(b->d_qws[k]));
        b->d_qws[k] = nullptr;
        b->qws_bytes[k] = 0;
        CK(cudaMalloc(&b->d_qws[k], need));
        b->qws_bytes[k] = need;
      }
      if (kCtlInts * nchunks > b->sched_ints[k]) {
        if (b->d_sched[k]) CK(cudaFree(b->d_sched[k]));
        b->d_sched[k] = nullptr;
        b->sched_ints[k] = 0;
        const int ints = kCtlInts * std::max(nchunks, 4);
        CK(cudaMalloc(&b->d_sched[k], sizeof(int) * ints));
        b->sched_ints[k] = ints;
      }
    }
  }
  CK(cudaMemsetAsync(b->d_sched[si], 0, sizeof(int) * kCtlInts * nchunks, st));
  const CmpcParams base = P;
  // cmpc_batch_profile_range: CUDA events between the kernel classes (the stream is drained per class)
  auto prof_begin = [&]() -> int {
    if (b->profiling) CK(cudaEventRecord(b->prof_ev[0], st));
    return CMPC_OK;
  };
  auto prof_end = [&](int cls) -> int {
    if (!b->profiling) return CMPC_OK;
    CK(cudaEventRecord(b->prof_ev[1], st));
    CK(cudaEventSynchronize(b->prof_ev[1]));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, b->prof_ev[0], b->prof_ev[1]));
    b->prof_ms[cls] += t;
    return CMPC_OK;
  };
  for (int c = 0; c < nchunks; c++) {
    const int off = c * chunk, cnt = std::min(chunk, count - off);
    CmpcParams Q = base;
    Q.count = cnt;
    Q.records = base.records + (size_t)off * base.rec_stride;
    Q.forces = base.forces + (size_t)off * 12 * base.horizon;
    Q.objective = base.objective + off;
    Q.status = base.status + off;
    Q.iterations = base.iterations + off;
    Q.active = base.active + (size_t)off * 20 * base.horizon;
    if (adapt) {
      Q.win_t = base.win_t + (size_t)off * CMPC_ADAPT_WINDOW;
      Q.win_d = base.win_d + (size_t)off * CMPC_ADAPT_WINDOW;
      Q.sim_time = base.sim_time + off;
      Q.est = base.est + (size_t)off * 4;
      Q.f_est = base.f_est + (size_t)off * 6;
    }
    Q.qws = b->d_qws[si];
    Q.qws_stride = slot;
    Q.k_tiled = tiled;
    Q.sweep_dmma = (!tiled && cshape != CMPC_CSHAPE_64 && b->sweep_dmma) ? 1 : 0;
    Q.inv_refine = (double)b->inv_refine;
    Q.inv_f32 = b->inv_f32;
    Q.qws_goff = cmpc_qws_goff(nmax, tiled);
    Q.worklist = nullptr;
    Q.count_ptr = nullptr;
    int* const ctl = b->d_sched[si] + kCtlInts * c;
    Q.sched = ctl + kCtlSched;
    const int ipc = cmpc_condense_instances_per_cta(cshape);
    if (int e = prof_begin()) return e;
    int rc = cmpc_launch_condense(Q, cshape, std::min((cnt + ipc - 1) / ipc, b->sm_count * per_sm1), st);
    if (rc != 0) return fail_cuda((cudaError_t)rc, "cmpc_condense_kernel launch");
    b->launches++;
    if (int e = prof_end(CMPC_K_ASSEMBLE)) return e;
    const bool lpt = tiled && fast && b->lpt;
    Q.lpt_hist = nullptr;
    Q.lpt_key = nullptr;
    if (tiled) {  // the assembly kernel left H tiles: invert them in place on the FP64 tensor cores
      Q.sm_slots = ctl + kCtlSlots;
      Q.inv_stagger = b->inv_stagger;
      if (lpt) {
        Q.lpt_hist = ctl + kCtlHist;
        Q.lpt_key = b->d_lpt[si];
      }
      Q.sched = ctl + kCtlSched + 3;
      const int ipc2 = cmpc_invert_instances_per_cta();
      if (int e = prof_begin()) return e;
      rc = cmpc_launch_invert(Q, std::min((cnt + ipc2 - 1) / ipc2, b->sm_count * (b->inv_ctas > 0 ? std::min(b->inv_ctas, per_sm_inv) : per_sm_inv)), st);
      if (rc != 0) return fail_cuda((cudaError_t)rc, "cmpc_invert_mma_kernel launch");
      b->launches++;
      if (int e = prof_end(CMPC_K_INVERT)) return e;
    }
    Q.sched = ctl + kCtlSched + 1;
    Q.qcap = qcap1;
    if (int e = prof_begin()) return e;
    Q.resume_in = nullptr;
    Q.resume_out = nullptr;
    if (qcap1 < nmax) {
      Q.overflow_list = b->d_overflow[si];
      Q.overflow_count = ctl + kCtlOvf;
      Q.rstate_out = nullptr;
      Q.rstate_in = nullptr;
      if (fast && b->resume) {
        Q.resume_out = b->d_resume[si];
        Q.rstate_out = b->d_rstate[si];
        Q.rstate_out_stride = kRstateStride;
        Q.rstate_out_cap = b->knobs.resume_p ? b->rstate_cap : 0;
      }
    } else {
      Q.overflow_list = nullptr;
    }
    if (lpt) {  // hardest instances first: the makespan of the kernel is its longest active-set run
      rc = cmpc_launch_lpt_order(Q.lpt_hist, Q.lpt_key, b->d_lpt[si] + b->capacity, cnt, st);
      if (rc != 0) return fail_cuda((cudaError_t)rc, "cmpc_lpt_order_kernel launch");
      b->launches++;
      Q.worklist = b->d_lpt[si] + b->capacity;
    }
    if (fast) rc = cmpc_launch_dual_fast(Q, std::min(cnt, b->sm_count * per_sm_fast), st);
    else rc = cmpc_launch_dual(Q, wpc1, std::min((cnt + wpc1 - 1) / wpc1, b->sm_count * per_sm2), st);
    if (rc != 0) return fail_cuda((cudaError_t)rc, "cmpc_dual_kernel launch");
    b->launches++;
    if (qcap1 < nmax) {
      // the instances that outgrew the first tier: resumed at a middle capacity (three warps per SM instead of one for
      // the larger problems), the few that outgrow that too at full capacity
      const bool mid = pl.qcap_mid > 0 && pl.qcap_mid < nmax && fast && b->resume;
      const bool team = pl.qcap_team > 0 && b->resume;
      Q.sched = ctl + kCtlSched + 2;
      Q.worklist = b->d_overflow[si];
      Q.count_ptr = ctl + kCtlOvf;
      Q.resume_in = Q.resume_out;
      Q.resume_out = nullptr;
      Q.rstate_in = Q.rstate_out;
      Q.rstate_in_stride = Q.rstate_out_stride;
      Q.rstate_in_cap = Q.rstate_out_cap;
      Q.rstate_out = nullptr;
      Q.overflow_list = nullptr;
      if (team || mid) {
        const bool last_tier = team && pl.qcap_team >= nmax;
        Q.qcap = team ? pl.qcap_team : pl.qcap_mid;
        if (!last_tier) {
          Q.overflow_list = b->d_overflow2[si];
          Q.overflow_count = ctl + kCtlOvf2;
          Q.resume_out = b->d_resume2[si];
          Q.rstate_out = b->d_rstate2[si];
          Q.rstate_out_stride = kRstate2Stride;
          Q.rstate_out_cap = b->knobs.resume_p ? b->rstate2_cap : 0;
        }
        if (team) rc = cmpc_launch_dual_team(Q, std::min(cnt, b->sm_count * pl.per_sm_team), st);
        else rc = cmpc_launch_dual(Q, pl.wpc_mid, std::min((cnt + pl.wpc_mid - 1) / pl.wpc_mid, b->sm_count * pl.per_sm_mid), st);
        if (rc != 0) return fail_cuda((cudaError_t)rc, "cmpc_dual_kernel (middle capacity) launch");
        b->launches++;
        if (last_tier) {
          if (int e = prof_end(CMPC_K_DUAL)) return e;
          continue;
        }
        Q.worklist = b->d_overflow2[si];
        Q.count_ptr = ctl + kCtlOvf2;
        Q.resume_in = b->d_resume2[si];
        Q.resume_out = nullptr;
        Q.rstate_in = b->d_rstate2[si];
        Q.rstate_in_stride = kRstate2Stride;
        Q.rstate_in_cap = b->rstate2_cap;
        Q.rstate_out = nullptr;
        Q.overflow_list = nullptr;
        Q.sched = ctl + kCtlSched + 3;  // the inversion kernel's counter, unused (and zero) on the shapes beyond 63 variables
      }
      Q.qcap = nmax;
      rc = cmpc_launch_dual(Q, wpc2, std::min((cnt + wpc2 - 1) / wpc2, b->sm_count * per_sm3), st);
      if (rc != 0) return fail_cuda((cudaError_t)rc, "cmpc_dual_kernel (full capacity) launch");
      b->launches++;
    }
    if (int e = prof_end(CMPC_K_DUAL)) return e;
  }
  return CMPC_OK;
}

// enqueue the solve of the uploaded instances [first, first+count) on stream si
int launch_range(cmpc_batch* b, int first, int count, int max_contact, int si) {
  if (count <= 0) return CMPC_OK;
  cudaStream_t st = b->stream[si];
  CmpcParams P;
  std::memset(&P, 0, sizeof(P));
  P.horizon = b->h;
  P.count = count;
  P.rec_stride = b->rec_stride;
  P.nmax = std::max(3, 3 * max_contact);
  P.max_iter = 20 * P.nmax + 100;
  P.adapt_mode = b->adapt_mode;
  P.dt = (double)(float)b->dt;
  P.mu_inv = (double)(1.f / (float)b->mu);
  P.f_max = (double)(float)b->f_max;
  P.mass_inv = 1.0 / (double)(float)b->mass;
  for (int i = 0; i < 3; i++) P.inertia[i] = (double)(float)b->inertia[i];
  P.gravity = (double)(-9.8f);
  P.tol_violation = 1e-9;
  P.tol_active = 1e-6;
  P.records = b->d_rec + (size_t)first * b->rec_stride;
  P.sigma = b->d_sigma;
  P.worklist = nullptr;
  P.overflow_list = b->d_overflow[si];
  P.overflow_count = b->d_overflow[si] + b->capacity;
  P.forces = (b->o_forces ? b->o_forces : b->d_forces) + (size_t)first * 12 * b->h;
  P.objective = (b->o_obj ? b->o_obj : b->d_obj) + first;
  P.status = (b->o_status ? b->o_status : b->d_status) + first;
  P.iterations = (b->o_iters ? b->o_iters : b->d_iters) + first;
  P.active = (b->o_active ? b->o_active : b->d_active) + (size_t)first * 20 * b->h;
  P.flops = b->d_flops;
  P.phase_cycles = b->d_phase;
  if (b->adapt_mode >= 0) {
    P.twiddle = b->d_twiddle;
    P.gk = b->d_gk;
    P.win_t = b->d_win_t + (size_t)first * CMPC_ADAPT_WINDOW;
    P.win_d = b->d_win_d + (size_t)first * CMPC_ADAPT_WINDOW;
    P.sim_time = b->d_simtime + first;
    P.est = b->d_est + (size_t)first * 4;
    P.f_est = b->d_fest + (size_t)first * 6;
  }
  // reduced problems of up to 128 variables take the two-kernel pipeline; CMPC_PATH=fused forces the
  // single-kernel path below (kept for larger problems and for A/B measurements)
  if (P.nmax <= CMPC_PIPELINE_NMAX && !b->knobs.path_fused) return launch_pipeline(b, P, count, si);
  // kernel shape by reduced problem size; the "shape" option overrides (0..3)
  int shape = P.nmax <= 64 ? CMPC_SHAPE_64W : (P.nmax <= 128 ? CMPC_SHAPE_128 : CMPC_SHAPE_MEM);
  if (b->knobs.shape >= 0) {
    int sh = b->knobs.shape;
    if (sh == CMPC_SHAPE_MEM || (sh == CMPC_SHAPE_128 && P.nmax <= 128) ||
        ((sh == CMPC_SHAPE_64 || sh == CMPC_SHAPE_64W) && P.nmax <= 64))
      shape = sh;
  }
  // two working-set capacity tiers: a small first tier keeps shared memory (and so occupancy) low;
  // the few instances that outgrow it are re-solved from scratch by a full-capacity launch
  int qcap1 = b->knobs.qcap1 > 0 ? b->knobs.qcap1 : 32;
  if (qcap1 < 1 || qcap1 > P.nmax) qcap1 = P.nmax;
  for (int tier = 0; tier < 2; tier++) {
    if (tier == 0) {
      P.qcap = qcap1;
      if (qcap1 < P.nmax) CK(cudaMemsetAsync(P.overflow_count, 0, sizeof(int), st));
      else P.overflow_list = nullptr;
    } else {
      if (qcap1 >= P.nmax) break;
      P.qcap = P.nmax;
      P.worklist = b->d_overflow[si];
      P.count_ptr = b->d_overflow[si] + b->capacity;
      P.overflow_list = nullptr;
      if (P.adapt_mode == 0 || P.adapt_mode == 1) P.adapt_mode = 2;  // the estimate already exists
    }
    size_t smem = cmpc_smem_bytes(P.horizon, P.nmax, P.qcap, shape, P.adapt_mode >= 0);
    if (smem > 200 * 1024 && shape == CMPC_SHAPE_MEM) {
      // the matrix no longer fits shared memory: keep K and P in a per-CTA global workspace (L2 resident)
      shape = CMPC_SHAPE_GMEM;
      smem = cmpc_smem_bytes(P.horizon, P.nmax, P.qcap, shape, P.adapt_mode >= 0);
    }
    int per_sm = cmpc_max_ctas_per_sm(shape, smem, P.adapt_mode >= 0);
    if (per_sm < 1) {
      g_err = "cmpc_batch_solve: kernel not launchable on this device (no sm_100a image?)";
      return CMPC_E_NODEVICE;
    }
    if (shape == CMPC_SHAPE_GMEM) per_sm = std::min(per_sm, 2);
    int grid = std::min(count, b->sm_count * per_sm);
    if (shape == CMPC_SHAPE_GMEM) {
      const size_t stride = (size_t)P.nmax * P.nmax + (size_t)(P.nmax + 1) * (P.nmax + 2) / 2;
      const size_t need = sizeof(double) * stride * (size_t)b->sm_count * 2 * kMaxStreams;  // every stream's worth
      if (need > b->gws_bytes) {
        { int rcs = sync_all(b); if (rcs) return rcs; }
        if (b->d_gws) CK(cudaFree(b->d_gws));
        b->d_gws = nullptr;
        CK(cudaMalloc(&b->d_gws, need));
        b->gws_bytes = need;
      }
      P.gws = b->d_gws + (size_t)si * stride * (size_t)b->sm_count * 2;
      P.gws_stride = stride;
    } else {
      P.gws = nullptr;
    }
    if (b->profiling) CK(cudaEventRecord(b->prof_ev[0], st));
    int rc = cmpc_launch_solve(P, shape, grid, st);
    if (rc != 0) return fail_cuda((cudaError_t)rc, "cmpc_solve_kernel launch");
    b->launches++;
    if (b->profiling) {
      CK(cudaEventRecord(b->prof_ev[1], st));
      CK(cudaEventSynchronize(b->prof_ev[1]));
      float t = 0.f;
      CK(cudaEventElapsedTime(&t, b->prof_ev[0], b->prof_ev[1]));
      b->prof_ms[CMPC_K_FUSED] += t;
    }
  }
  return CMPC_OK;
}

}  // namespace

extern "C" {

const char* cmpc_last_error(void) { return g_err.c_str(); }

int cmpc_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    fail_cuda(e, "cudaGetDeviceCount");
    return 0;
  }
  return n;
}

static int batch_create_impl(cmpc_batch* b, int device, int capacity);

int cmpc_batch_create(cmpc_batch** out, int device, int capacity) {
  if (!out || capacity < 1) return fail_arg("cmpc_batch_create: bad arguments");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_err = "cmpc_batch_create: no CUDA device (this engine has no CPU path)";
    return CMPC_E_NODEVICE;
  }
  if (device < 0 || device >= ndev) return fail_arg("cmpc_batch_create: device out of range");
  CK(cudaSetDevice(device));
  cmpc_batch* b = new cmpc_batch();
  const int rc = batch_create_impl(b, device, capacity);
  if (rc != CMPC_OK) {  // release whatever was allocated before the failure (destroy tolerates the nulls)
    const std::string why = g_err;
    cmpc_batch_destroy(b);
    (void)cudaGetLastError();
    g_err = why;
    return rc;
  }
  *out = b;
  return CMPC_OK;
}

static int batch_create_impl(cmpc_batch* b, int device, int capacity) {
  b->device = device;
  b->capacity = capacity;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  b->sm_count = prop.multiProcessorCount;
  for (int i = 0; i < kMaxStreams; i++) CK(cudaStreamCreateWithFlags(&b->stream[i], cudaStreamNonBlocking));
  CK(cudaEventCreate(&b->ev0));
  CK(cudaEventCreate(&b->ev1));
  CK(cudaEventCreate(&b->mark0));
  CK(cudaEventCreate(&b->mark1));
  for (int i = 0; i < kMaxChunks; i++) CK(cudaEventCreateWithFlags(&b->chunk_done[i], cudaEventDisableTiming));
  for (int i = 0; i < kMaxChunks; i++) CK(cudaEventCreateWithFlags(&b->packed[i], cudaEventDisableTiming));
  for (int i = 0; i < kMaxChunks; i++) CK(cudaEventCreateWithFlags(&b->traj_copied[i], cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&b->rec_copied, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&b->join_ev, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&b->fork_ev, cudaEventDisableTiming));
  const size_t cap = (size_t)capacity;
  const int hm = CMPC_MAX_HORIZON;
  const size_t rec_max = (size_t)cmpc_rec_stride(hm);
  CK(cudaMallocHost(&b->h_rec, cap * rec_max));
  CK(cudaMalloc(&b->d_rec, cap * rec_max));
  CK(cudaMalloc(&b->d_traj_stage, sizeof(float) * cap * 12 * hm));
  CK(cudaMalloc(&b->d_sigma, sizeof(double) * CMPC_SIG_COUNT * hm * hm));
  CK(cudaMalloc(&b->d_forces, sizeof(double) * cap * 12 * hm));
  CK(cudaMalloc(&b->d_obj, sizeof(double) * cap));
  CK(cudaMalloc(&b->d_status, sizeof(int) * cap));
  CK(cudaMalloc(&b->d_iters, sizeof(int) * cap));
  CK(cudaMalloc(&b->d_active, cap * 20 * hm));
  for (int i = 0; i < kMaxStreams; i++) CK(cudaMalloc(&b->d_overflow[i], sizeof(int) * (cap + 1)));
  for (int i = 0; i < kMaxStreams; i++) CK(cudaMalloc(&b->d_lpt[i], sizeof(int) * 2 * cap));
  for (int i = 0; i < kMaxStreams; i++) CK(cudaMalloc(&b->d_resume[i], sizeof(int) * CMPC_RESUME_INTS * cap));
  for (int i = 0; i < kMaxStreams; i++) CK(cudaMalloc(&b->d_overflow2[i], sizeof(int) * (cap + 1)));
  for (int i = 0; i < kMaxStreams; i++) CK(cudaMalloc(&b->d_resume2[i], sizeof(int) * CMPC_RESUME_INTS * cap));
  b->rstate_cap = (int)std::min<size_t>(cap, std::max<size_t>(1024, cap / 4));
  b->rstate2_cap = (int)std::min<size_t>(cap, std::max<size_t>(256, cap / 16));
  for (int i = 0; i < kMaxStreams; i++) CK(cudaMalloc(&b->d_rstate[i], sizeof(double) * kRstateStride * (size_t)b->rstate_cap));
  for (int i = 0; i < kMaxStreams; i++) CK(cudaMalloc(&b->d_rstate2[i], sizeof(double) * kRstate2Stride * (size_t)b->rstate2_cap));
  CK(cudaMalloc(&b->d_flops, sizeof(unsigned long long) * CMPC_K_COUNT));
  CK(cudaMemset(b->d_flops, 0, sizeof(unsigned long long) * CMPC_K_COUNT));
  CK(cudaMallocHost(&b->h_forces, sizeof(double) * cap * 12 * hm));
  CK(cudaMallocHost(&b->h_obj, sizeof(double) * cap));
  CK(cudaMallocHost(&b->h_status, sizeof(int) * cap));
  CK(cudaMallocHost(&b->h_iters, sizeof(int) * cap));
  CK(cudaMallocHost(&b->h_active, cap * 20 * hm));
  CK(cudaMallocHost(&b->h_flops, sizeof(unsigned long long) * CMPC_K_COUNT));
  for (int i = 0; i < CMPC_K_COUNT; i++) b->h_flops[i] = 0;
  return CMPC_OK;
}

void cmpc_batch_destroy(cmpc_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  for (int i = 0; i < kMaxStreams; i++) if (b->stream[i]) cudaStreamSynchronize(b->stream[i]);
  cudaFreeHost(b->h_rec); cudaFree(b->d_rec); cudaFree(b->d_sigma); cudaFree(b->d_forces); cudaFree(b->d_obj);
  cudaFree(b->d_status); cudaFree(b->d_iters); cudaFree(b->d_active); cudaFree(b->d_flops); cudaFree(b->d_phase);
  cudaFree(b->d_gws);
  for (int i = 0; i < kMaxStreams; i++) { cudaFree(b->d_overflow[i]); cudaFree(b->d_qws[i]); cudaFree(b->d_sched[i]); cudaFree(b->d_lpt[i]); cudaFree(b->d_resume[i]); cudaFree(b->d_overflow2[i]); cudaFree(b->d_resume2[i]); cudaFree(b->d_rstate[i]); cudaFree(b->d_rstate2[i]); }
  cudaFree(b->d_twiddle); cudaFree(b->d_gk); cudaFree(b->d_win_t); cudaFree(b->d_win_d); cudaFree(b->d_simtime);
  cudaFree(b->d_est); cudaFree(b->d_fest);
  cudaFree(b->d_cmds); cudaFree(b->d_results); cudaFree(b->d_fext);
  cudaFreeHost(b->h_forces); cudaFreeHost(b->h_obj); cudaFreeHost(b->h_status); cudaFreeHost(b->h_iters);
  cudaFreeHost(b->h_active); cudaFreeHost(b->h_flops);
  cudaEventDestroy(b->ev0); cudaEventDestroy(b->ev1); cudaEventDestroy(b->mark0); cudaEventDestroy(b->mark1);
  for (int i = 0; i < kMaxChunks; i++) cudaEventDestroy(b->chunk_done[i]);
  for (int i = 0; i < kMaxChunks; i++) cudaEventDestroy(b->packed[i]);
  for (int i = 0; i < kMaxChunks; i++) cudaEventDestroy(b->traj_copied[i]);
  if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
  cudaFree(b->d_traj_stage);
  for (int i = 0; i < CMPC_K_COUNT + 1; i++) if (b->prof_ev[i]) cudaEventDestroy(b->prof_ev[i]);
  cudaEventDestroy(b->join_ev); cudaEventDestroy(b->fork_ev); cudaEventDestroy(b->rec_copied);
  for (int i = 0; i < kMaxStreams; i++) cudaStreamDestroy(b->stream[i]);
  (void)cudaGetLastError();  // handles that were never created (a create that failed half way)
  delete b->pool;
  delete b;
}

int cmpc_batch_setup(cmpc_batch* b, double dt, int horizon, double mu, double f_max) {
  if (!b) return fail_arg("cmpc_batch_setup: null batch");
  if (horizon < 1 || horizon > CMPC_MAX_HORIZON) return fail_arg("cmpc_batch_setup: horizon must be 1..19 (SolverMPC.cpp:113)");
  if (!(mu > 0) || !(dt > 0)) return fail_arg("cmpc_batch_setup: dt and mu must be positive");
  CK(cudaSetDevice(b->device));
  const bool same = b->is_setup && b->h == horizon && b->dt == dt;
  b->h = horizon;
  b->dt = dt;
  b->mu = mu;
  b->f_max = f_max;
  b->rec_stride = cmpc_rec_stride(horizon);
  if (!same) {
    // the reference narrows dt to float (problem_setup.dt, convexMPC_interface.h:17)
    std::vector<double> sig;
    build_sigma(horizon, (double)(float)dt, sig);
    { int rcs = sync_all(b); if (rcs) return rcs; }
    CK(cudaMemcpyAsync(b->d_sigma, sig.data(), sizeof(double) * sig.size(), cudaMemcpyHostToDevice, b->stream[0]));
    CK(cudaStreamSynchronize(b->stream[0]));
    b->count = 0;
  }
  b->is_setup = true;
  return CMPC_OK;
}

int cmpc_batch_set_robot(cmpc_batch* b, double mass, const double inertia_diag[3]) {
  if (!b || !(mass > 0)) return fail_arg("cmpc_batch_set_robot: bad arguments");
  b->mass = mass;
  if (inertia_diag)
    for (int i = 0; i < 3; i++) b->inertia[i] = inertia_diag[i];
  return CMPC_OK;
}

int cmpc_batch_upload(cmpc_batch* b, int count, const cmpc_inputs* in) {
  int rc = check_inputs(b, count, in, "cmpc_batch_upload");
  if (rc) return rc;
  CK(cudaSetDevice(b->device));
  // the previous copy out of the pinned staging may still be queued behind solves: the host must not overwrite
  // the staging before it has run
  { int rcw = wait_rec_staging(b); if (rcw) return rcw; }
  b->max_contact = pack_records(b, in, 0, count);
  b->count = count;
  { int rcj = join_streams(b); if (rcj) return rcj; }  // solves in flight on stream 1 still read the records
  if (count > 0) {
    CK(cudaMemcpyAsync(b->d_rec, b->h_rec, (size_t)count * b->rec_stride, cudaMemcpyHostToDevice, b->stream[0]));
    CK(cudaEventRecord(b->rec_copied, b->stream[0]));
    b->rec_copy_pending = true;
  }
  return CMPC_OK;
}

int cmpc_batch_set_count(cmpc_batch* b, int count, int max_contact_feet) {
  if (!b || count < 0 || count > b->capacity) return fail_arg("cmpc_batch_set_count: bad arguments");
  if (!b->is_setup) { g_err = "cmpc_batch_set_count: call cmpc_batch_setup first"; return CMPC_E_STATE; }
  if (max_contact_feet < 0 || max_contact_feet > 4 * b->h) return fail_arg("cmpc_batch_set_count: bad contact bound");
  b->count = count;
  b->max_contact = max_contact_feet;
  return CMPC_OK;
}

int cmpc_batch_solve(cmpc_batch* b) {
  if (!b) return fail_arg("cmpc_batch_solve: null batch");
  return cmpc_batch_solve_range(b, 0, b->count);
}

int cmpc_batch_solve_range(cmpc_batch* b, int first, int count) {
  if (!b) return fail_arg("cmpc_batch_solve_range: null batch");
  if (!b->is_setup) { g_err = "cmpc_batch_solve: call cmpc_batch_setup first"; return CMPC_E_STATE; }
  if (first < 0 || count < 0 || first + count > b->count) return fail_arg("cmpc_batch_solve_range: range outside the uploaded instances");
  CK(cudaSetDevice(b->device));
  // a large range is cut into `split` parts on successive streams: kernels of different parts (assembly,
  // inversion, active set) then share the SMs instead of each waiting for the previous kernel's tail
  b->throughput_mode = !b->serial;
  const int parts = (b->serial || count < 512 * b->split) ? 1 : b->split;
  const int per = (count + parts - 1) / parts;
  for (int part = 0; part < parts; part++) {
    const int pf = first + part * per, pn = std::min(per, first + count - pf);
    if (pn <= 0) break;
    int si = 0;
    if (!b->serial) si = (int)(b->rr++ % (unsigned)b->nstreams);
    if (si != 0) {
      int rcf = fork_streams(b, si);  // after the uploads / marks already enqueued on the batch stream
      if (rcf) return rcf;
    }
    if (part == 0) CK(cudaEventRecord(b->ev0, b->stream[si]));
    int rc = launch_range(b, pf, pn, b->max_contact, si);
    if (si != 0) b->dirty |= 1u << si;  // after the launch: growing a workspace inside it drains every stream and clears the marks
    if (rc) return rc;
    if (part == parts - 1) CK(cudaEventRecord(b->ev1, b->stream[si]));
  }
  b->timed = true;
  return CMPC_OK;
}

int cmpc_batch_sync(cmpc_batch* b) {
  if (!b) return fail_arg("cmpc_batch_sync: null batch");
  CK(cudaSetDevice(b->device));
  return sync_all(b);
}

// Results travel device -> pinned staging -> caller's array; an output array that is itself pinned
// (cudaHostRegister / cmpc_host_register) receives the copy directly (bit i of `direct`).
enum { kOutForces = 1, kOutObj = 2, kOutStatus = 4, kOutIters = 8, kOutActive = 16 };

static int direct_mask(const cmpc_outputs* out) {
  int m = 0;
  if (is_pinned(out->forces)) m |= kOutForces;
  if (is_pinned(out->objective)) m |= kOutObj;
  if (is_pinned(out->status)) m |= kOutStatus;
  if (is_pinned(out->iterations)) m |= kOutIters;
  if (is_pinned(out->active)) m |= kOutActive;
  return m;
}

static int enqueue_d2h(cmpc_batch* b, const cmpc_outputs* out, int first, int n, cudaStream_t st, int direct = 0) {
  const int h = b->h;
  const size_t f = (size_t)first, c = (size_t)n;
  if (n <= 0) return CMPC_OK;
  if (out->forces) CK(cudaMemcpyAsync(((direct & kOutForces) ? out->forces : b->h_forces) + f * 12 * h, b->d_forces + f * 12 * h, sizeof(double) * c * 12 * h, cudaMemcpyDeviceToHost, st));
  if (out->objective) CK(cudaMemcpyAsync(((direct & kOutObj) ? out->objective : b->h_obj) + f, b->d_obj + f, sizeof(double) * c, cudaMemcpyDeviceToHost, st));
  if (out->status) CK(cudaMemcpyAsync(((direct & kOutStatus) ? out->status : b->h_status) + f, b->d_status + f, sizeof(int) * c, cudaMemcpyDeviceToHost, st));
  if (out->iterations) CK(cudaMemcpyAsync(((direct & kOutIters) ? out->iterations : b->h_iters) + f, b->d_iters + f, sizeof(int) * c, cudaMemcpyDeviceToHost, st));
  if (out->active) CK(cudaMemcpyAsync(((direct & kOutActive) ? out->active : b->h_active) + f * 20 * h, b->d_active + f * 20 * h, c * 20 * h, cudaMemcpyDeviceToHost, st));
  return CMPC_OK;
}

static void unpack_results(cmpc_batch* b, const cmpc_outputs* out, int first, int n, int direct = 0) {
  const int h = b->h;
  const size_t f = (size_t)first, c = (size_t)n;
  if (n <= 0) return;
  if (out->forces && !(direct & kOutForces)) {
    const size_t bytes = sizeof(double) * c * 12 * h;
    if (bytes >= (8u << 20)) {  // a large array: split over the host workers
      const int parts = 8;
      host_pool(b)->parallel_for(parts, [&](int i) {
        const size_t lo = bytes * i / parts, hi = bytes * (i + 1) / parts;
        std::memcpy(reinterpret_cast<char*>(out->forces + f * 12 * h) + lo, reinterpret_cast<const char*>(b->h_forces + f * 12 * h) + lo, hi - lo);
      });
    } else {
      std::memcpy(out->forces + f * 12 * h, b->h_forces + f * 12 * h, bytes);
    }
  }
  if (out->objective && !(direct & kOutObj)) std::memcpy(out->objective + f, b->h_obj + f, sizeof(double) * c);
  if (out->status && !(direct & kOutStatus)) std::memcpy(out->status + f, b->h_status + f, sizeof(int) * c);
  if (out->iterations && !(direct & kOutIters)) std::memcpy(out->iterations + f, b->h_iters + f, sizeof(int) * c);
  if (out->active && !(direct & kOutActive)) std::memcpy(out->active + f * 20 * h, b->h_active + f * 20 * h, c * 20 * h);
}

int cmpc_batch_download(cmpc_batch* b, const cmpc_outputs* out) {
  if (!b || !out) return fail_arg("cmpc_batch_download: null argument");
  CK(cudaSetDevice(b->device));
  { int rcs = sync_aux(b); if (rcs) return rcs; }
  const int direct = direct_mask(out);
  int rc = enqueue_d2h(b, out, 0, b->count, b->stream[0], direct);
  if (rc) return rc;
  CK(cudaStreamSynchronize(b->stream[0]));
  unpack_results(b, out, 0, b->count, direct);
  return CMPC_OK;
}

// End-to-end call with host buffers.  The batch is cut into chunks that alternate between two streams.  Inputs:
// pinned arrays are read by the device itself, which packs the records (cmpc_pack.cu); pageable arrays are packed
// into pinned records on the host, chunk c+1 while chunk c is copied in and solved.  Outputs: the kernels write
// them straight into host memory over PCIe as instances finish (the caller's arrays if pinned, else pinned staging
// that is copied out here), so no device-to-host copy waits behind the last kernel.  CMPC_D2H_COPY=1 restores
// device-side outputs + cudaMemcpyAsync.
static int resolve_binding(cmpc_batch* b, const cmpc_inputs* in, const cmpc_outputs* out, HostBinding& hb) {
  hb.in = *in;
  hb.out = *out;
  hb.direct = direct_mask(out);
  // pinned input arrays are read by the device itself (cmpc_pack.cu); pageable ones are packed into pinned records on the host
  hb.soa = soa_view(in);
  if (b->knobs.host_pack) hb.soa.ok = false;
  hb.zc_out = !b->knobs.d2h_copy;
  for (int i = 0; i < 5; i++) hb.vo[i] = nullptr;
  if (hb.zc_out) {
    auto view = [&](void* user, void* staging, int bit) -> void* {
      if (!user) return nullptr;
      return const_cast<void*>(device_view((hb.direct & bit) ? user : staging));
    };
    hb.vo[0] = view(out->forces, b->h_forces, kOutForces);
    hb.vo[1] = view(out->objective, b->h_obj, kOutObj);
    hb.vo[2] = view(out->status, b->h_status, kOutStatus);
    hb.vo[3] = view(out->iterations, b->h_iters, kOutIters);
    hb.vo[4] = view(out->active, b->h_active, kOutActive);
  }
  hb.valid = true;
  return CMPC_OK;
}

static int solve_host_core(cmpc_batch* b, int count, const HostBinding& hb, bool wait = true) {
#ifdef CMPC_EXPERIMENTS
  static const bool trace = std::getenv("CMPC_TRACE") != nullptr;
#else
  constexpr bool trace = false;
#endif
  static double tr_enq = 0, tr_wait = 0, tr_scan = 0, tr_pack = 0, tr_pipe = 0;
  static cudaEvent_t tr_ev[2 * kMaxChunks] = {};
  static int tr_n = 0;
  const auto tr0 = std::chrono::steady_clock::now();
  const cmpc_inputs* in = &hb.in;
  const cmpc_outputs* out = &hb.out;
  int rc = CMPC_OK;
  b->throughput_mode = false;
  CK(cudaSetDevice(b->device));
  int nchunks = 1;
  if (b->knobs.chunks > 0) nchunks = b->knobs.chunks;
  else nchunks = std::max(1, std::min(4, count / 4096));  // measured (scripts/e2e_chunks.py): one chunk up to 8191, four from 16384
  nchunks = std::max(1, std::min(nchunks, kMaxChunks));
  const int direct = hb.direct;
  const SoaView& soa = hb.soa;
  // submit / wait (batches in flight): results written over PCIe by the active-set kernel's own warps hold those warps
  // while other batches want the SMs (profiles/r1_s4_skip_pack.txt), so CMPC_SUBMIT_COPY=1 hands them to the copy engine
  const bool zc_out = hb.zc_out && !(b->knobs.submit_copy && !wait);
  void* const* vo = hb.vo;
  const int per = (count + nchunks - 1) / nchunks;
  { int rcs = sync_aux(b); if (rcs) return rcs; }
  CK(cudaEventRecord(b->ev0, b->stream[0]));
  b->count = count;
  int maxc_all = 0, used = 0;
  const int h = b->h;
  struct OutGuard {  // the overrides only live for this call
    cmpc_batch* b;
    ~OutGuard() { b->o_forces = nullptr; b->o_obj = nullptr; b->o_status = nullptr; b->o_iters = nullptr; b->o_active = nullptr; }
  } guard{b};
  // Pinned inputs: every chunk's records are packed by the device first, chunk after chunk on a stream of their own
  // (the reads are PCIe bound, and queued ahead of the solve kernels they find free SM slots); a chunk's solve
  // waits for its own records only.
  int chunk_maxc[kMaxChunks] = {};
  if (soa.ok) {
    cudaStream_t ps = b->stream[kPackStream];
    CK(cudaStreamWaitEvent(ps, b->ev0, 0));
    // The reference trajectories are 71 % of the input bytes: the copy engine moves them (52 GB/s against ~35 GB/s for
    // loads issued by SMs, scripts/pcie_copy_rate.py) while the packing kernel reads the ten small arrays over PCIe
    // itself; a second launch then scatters the staged trajectories from HBM.
#ifdef CMPC_EXPERIMENTS
    // experiments build only (scripts/e2e_depth.py): reuse the records of the previous call, to tell the cost of the input side
    const bool skip_pack = b->exp_skip_pack && b->exp_packed_once;
    b->exp_packed_once = true;
#else
    constexpr bool skip_pack = false;
#endif
    const bool tcopy = !skip_pack && b->traj_copy && b->d_traj_stage && soa.p[6];
    if (tcopy) {
      CK(cudaStreamWaitEvent(b->copy_stream, b->ev0, 0));
      for (int c = 0; c < nchunks; c++) {
        const int first = c * per, n = std::min(per, count - first);
        if (n <= 0) break;
        const size_t off = (size_t)first * 12 * h;
        CK(cudaMemcpyAsync(b->d_traj_stage + off, soa.p[6] + off * sizeof(float), sizeof(float) * (size_t)n * 12 * h, cudaMemcpyDefault, b->copy_stream));
        CK(cudaEventRecord(b->traj_copied[c], b->copy_stream));
      }
    }
    for (int c = 0; c < nchunks; c++) {
      const int first = c * per, n = std::min(per, count - first);
      if (n <= 0) break;
      const auto tq0 = std::chrono::steady_clock::now();
      const size_t f = (size_t)first;
      auto at = [&](int i, size_t bytes_per_instance) -> const void* { return soa.p[i] ? soa.p[i] + f * bytes_per_instance : nullptr; };
      const void* traj_src = tcopy ? static_cast<const void*>(b->d_traj_stage + f * 12 * h) : at(6, 48 * (size_t)h);
      if (skip_pack) { CK(cudaEventRecord(b->packed[c], ps)); continue; }
      int rcp = cmpc_launch_pack(at(0, 12), at(1, 12), at(2, 16), at(3, 12), at(4, 48), at(5, 48), traj_src,
                                 at(7, 4), at(8, 4 * (size_t)h), at(9, 4), at(10, 24),
                                 b->d_rec + f * b->rec_stride, b->rec_stride, h, n, b->sm_count, ps, tcopy ? CMPC_PACK_REST : CMPC_PACK_ALL);
      if (rcp == 0 && tcopy) {
        CK(cudaStreamWaitEvent(ps, b->traj_copied[c], 0));
        rcp = cmpc_launch_pack(at(0, 12), at(1, 12), at(2, 16), at(3, 12), at(4, 48), at(5, 48), traj_src,
                               at(7, 4), at(8, 4 * (size_t)h), at(9, 4), at(10, 24),
                               b->d_rec + f * b->rec_stride, b->rec_stride, h, n, b->sm_count, ps, CMPC_PACK_TRAJ);
        b->launches++;
      }
      if (rcp != 0) return fail_cuda((cudaError_t)rcp, "cmpc_pack_records_kernel launch");
      b->launches++;
      CK(cudaEventRecord(b->packed[c], ps));
      if (trace) {
        if (!tr_ev[0]) for (int i = 0; i < 2 * kMaxChunks; i++) cudaEventCreate(&tr_ev[i]);
        CK(cudaEventRecord(tr_ev[2 * c], ps));
        tr_pack += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tq0).count();
      }
    }
    // the host's only pass over the inputs (the contact bound that sizes the solve kernels) runs while the device is
    // already reading the arrays
    for (int c = 0; c < nchunks; c++) {
      const int first = c * per, n = std::min(per, count - first);
      if (n <= 0) break;
      const auto tq0 = std::chrono::steady_clock::now();
      chunk_maxc[c] = max_contact_scan(b, in->gait, first, n);
      if (trace) tr_scan += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tq0).count();
    }
  }
  for (int c = 0; c < nchunks; c++) {
    const int first = c * per, n = std::min(per, count - first);
    if (n <= 0) break;
    const int si = c & 1;
    cudaStream_t st = b->stream[si];
    int maxc;
    const auto tq0 = std::chrono::steady_clock::now();
    if (soa.ok) {
      maxc = chunk_maxc[c];
      CK(cudaStreamWaitEvent(st, b->packed[c], 0));
    } else {
      if (c == 0) { int rcw = wait_rec_staging(b); if (rcw) return rcw; }  // an earlier upload may still be reading the staging
      maxc = pack_records_parallel(b, in, first, n);
      CK(cudaMemcpyAsync(b->d_rec + (size_t)first * b->rec_stride, b->h_rec + (size_t)first * b->rec_stride,
                         (size_t)n * b->rec_stride, cudaMemcpyHostToDevice, st));
      CK(cudaEventRecord(b->rec_copied, st));
      b->rec_copy_pending = true;
    }
    maxc_all = std::max(maxc_all, maxc);
    // the staged, word-wide output stores live in the pipeline's active-set kernel (reduced problems of up to 128
    // variables); the single-kernel path of larger problems keeps device-side outputs and a copy
    const bool zc = zc_out && 3 * maxc <= CMPC_PIPELINE_NMAX;
    b->o_forces = zc ? static_cast<double*>(vo[0]) : nullptr;
    b->o_obj = zc ? static_cast<double*>(vo[1]) : nullptr;
    b->o_status = zc ? static_cast<int*>(vo[2]) : nullptr;
    b->o_iters = zc ? static_cast<int*>(vo[3]) : nullptr;
    b->o_active = zc ? static_cast<signed char*>(vo[4]) : nullptr;
    const auto tq1 = std::chrono::steady_clock::now();
    rc = launch_range(b, first, n, maxc, si);
    if (rc) return rc;
    if (trace) {
      if (!soa.ok) tr_pack += std::chrono::duration<double, std::micro>(tq1 - tq0).count();
      tr_pipe += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tq1).count();
    }
    if (!zc) {
      rc = enqueue_d2h(b, out, first, n, st, direct);
      if (rc) return rc;
    }
    CK(cudaEventRecord(b->chunk_done[c], st));
    if (trace && tr_ev[0]) CK(cudaEventRecord(tr_ev[2 * c + 1], st));
    used = c + 1;
  }
  b->max_contact = maxc_all;
  if (!wait) {  // cmpc_batch_submit_bound: the caller collects the results with cmpc_batch_wait_bound
    b->pend_count = count;
    b->pend_per = per;
    b->pend_used = used;
    return CMPC_OK;
  }
  const auto tr1 = std::chrono::steady_clock::now();
  for (int c = 0; c < used; c++) {
    const int first = c * per, n = std::min(per, count - first);
    CK(cudaEventSynchronize(b->chunk_done[c]));
    unpack_results(b, out, first, n, direct);
  }
  CK(cudaStreamWaitEvent(b->stream[0], b->chunk_done[used > 0 ? used - 1 : 0], 0));
  CK(cudaEventRecord(b->ev1, b->stream[0]));
  b->timed = true;
  if (trace) {
    const auto tr2 = std::chrono::steady_clock::now();
    tr_enq += std::chrono::duration<double, std::micro>(tr1 - tr0).count();
    tr_wait += std::chrono::duration<double, std::micro>(tr2 - tr1).count();
    if (tr_n == 49 && tr_ev[0] && soa.ok) {
      for (int c = 0; c < used; c++) {
        float a = 0.f, d = 0.f;
        cudaEventElapsedTime(&a, b->ev0, tr_ev[2 * c]);
        cudaEventElapsedTime(&d, b->ev0, tr_ev[2 * c + 1]);
        std::fprintf(stderr, "[cmpc]   chunk %d: records packed at %.1f us, solved at %.1f us after the call's first event\n", c, 1e3 * a, 1e3 * d);
      }
    }
    if (++tr_n == 50) {
      std::fprintf(stderr, "[cmpc] solve_host: enqueue %.1f us (gait scan %.1f, scan + pack launch %.1f, pipeline launches %.1f), wait %.1f us per call\n",
                   tr_enq / tr_n, tr_scan / tr_n, tr_pack / tr_n, tr_pipe / tr_n, tr_wait / tr_n);
      tr_enq = tr_wait = tr_scan = tr_pack = tr_pipe = 0;
      tr_n = 0;
    }
  }
  return CMPC_OK;
}

int cmpc_batch_solve_host(cmpc_batch* b, int count, const cmpc_inputs* in, const cmpc_outputs* out) {
  int rc = check_inputs(b, count, in, "cmpc_batch_solve_host");
  if (rc) return rc;
  if (!out) return fail_arg("cmpc_batch_solve_host: null outputs");
  CK(cudaSetDevice(b->device));
  HostBinding hb;
  rc = resolve_binding(b, in, out, hb);
  if (rc) return rc;
  return solve_host_core(b, count, hb);
}

int cmpc_batch_bind_host(cmpc_batch* b, const cmpc_inputs* in, const cmpc_outputs* out) {
  if (!b) return fail_arg("cmpc_batch_bind_host: null batch");
  if (!in && !out) {
    b->bound.valid = false;
    return CMPC_OK;
  }
  int rc = check_inputs(b, 0, in, "cmpc_batch_bind_host");
  if (rc) return rc;
  if (!out) return fail_arg("cmpc_batch_bind_host: null outputs");
  CK(cudaSetDevice(b->device));
  return resolve_binding(b, in, out, b->bound);
}

// Diagnostic / test switches.  Keys: nstreams, split, serial, lpt, resume, sweep_dmma, inv_stagger, inv_refine, inv_f32, inv_ctas, traj_copy, cshape,
// ws_mb, qcap1, dual_generic, wpc, no_mid_tier, path_fused, shape, host_pack, d2h_copy, chunks, submit_copy,
// host_threads, dual_team, resume_p (and exp_skip_pack in a -DCMPC_EXPERIMENTS build).
int cmpc_batch_set_option(cmpc_batch* b, const char* key, int value) {
  if (!b || !key) return fail_arg("cmpc_batch_set_option: null argument");
  CK(cudaSetDevice(b->device));
  { int rcs = sync_all(b); if (rcs) return rcs; }
  const std::string k(key);
  Knobs& kn = b->knobs;
  if (k == "nstreams") b->nstreams = std::max(1, std::min(kMaxStreams, value));
  else if (k == "split") b->split = std::max(1, std::min(kMaxStreams, value));
  else if (k == "serial") b->serial = value != 0;
  else if (k == "lpt") b->lpt = value != 0;
  else if (k == "resume") b->resume = value != 0;
  else if (k == "sweep_dmma") b->sweep_dmma = value != 0;
  else if (k == "inv_stagger") b->inv_stagger = std::max(0, value);
  else if (k == "inv_refine") b->inv_refine = std::max(-1, value);
  else if (k == "inv_f32") b->inv_f32 = value != 0;
  else if (k == "inv_ctas") b->inv_ctas = std::max(0, value);
  else if (k == "traj_copy") b->traj_copy = value != 0;
  else if (k == "cshape") kn.cshape = value;
  else if (k == "ws_mb") kn.ws_mb = std::max(1, value);
  else if (k == "qcap1") kn.qcap1 = std::max(0, value);
  else if (k == "dual_generic") kn.dual_generic = value != 0;
  else if (k == "wpc") kn.wpc = value;
  else if (k == "no_mid_tier") kn.no_mid_tier = value != 0;
  else if (k == "path_fused") kn.path_fused = value != 0;
  else if (k == "shape") kn.shape = value;
  else if (k == "host_pack") kn.host_pack = value != 0;
  else if (k == "d2h_copy") kn.d2h_copy = value != 0;
  else if (k == "chunks") kn.chunks = std::max(0, value);
  else if (k == "submit_copy") kn.submit_copy = value != 0;
  else if (k == "host_threads") kn.host_threads = std::max(0, value);
  else if (k == "dual_team") kn.dual_team = value != 0;
  else if (k == "resume_p") kn.resume_p = value != 0;
#ifdef CMPC_EXPERIMENTS
  else if (k == "exp_skip_pack") b->exp_skip_pack = value != 0;
#endif
  else {
    g_err = "cmpc_batch_set_option: unknown key '" + k + "'";
    return CMPC_E_ARG;
  }
  b->plans.clear();
  if (b->bound.valid) {  // the binding caches host_pack / d2h_copy
    HostBinding hb = b->bound;
    int rc = resolve_binding(b, &hb.in, &hb.out, b->bound);
    if (rc) return rc;
  }
  return CMPC_OK;
}

int cmpc_batch_submit_bound(cmpc_batch* b, int count) {
  if (!b) return fail_arg("cmpc_batch_submit_bound: null batch");
  if (!b->bound.valid) { g_err = "cmpc_batch_submit_bound: call cmpc_batch_bind_host first"; return CMPC_E_STATE; }
  if (!b->is_setup) { g_err = "cmpc_batch_submit_bound: call cmpc_batch_setup first"; return CMPC_E_STATE; }
  if (count < 0 || count > b->capacity) return fail_arg("cmpc_batch_submit_bound: count exceeds capacity");
  if (b->pend_used >= 0) { g_err = "cmpc_batch_submit_bound: the previous submission has not been waited for"; return CMPC_E_STATE; }
  b->pend_used = 0;
  int rc = solve_host_core(b, count, b->bound, false);
  if (rc) b->pend_used = -1;
  return rc;
}

int cmpc_batch_wait_bound(cmpc_batch* b) {
  if (!b) return fail_arg("cmpc_batch_wait_bound: null batch");
  if (b->pend_used < 0) { g_err = "cmpc_batch_wait_bound: nothing was submitted"; return CMPC_E_STATE; }
  CK(cudaSetDevice(b->device));
  const int used = b->pend_used, per = b->pend_per, count = b->pend_count;
  b->pend_used = -1;
  for (int c = 0; c < used; c++) {
    const int first = c * per, n = std::min(per, count - first);
    CK(cudaEventSynchronize(b->chunk_done[c]));
    unpack_results(b, &b->bound.out, first, n, b->bound.direct);
  }
  if (used > 0) {
    CK(cudaStreamWaitEvent(b->stream[0], b->chunk_done[used - 1], 0));
    CK(cudaEventRecord(b->ev1, b->stream[0]));
    b->timed = true;
  }
  return CMPC_OK;
}

int cmpc_batch_solve_bound(cmpc_batch* b, int count) {
  if (!b) return fail_arg("cmpc_batch_solve_bound: null batch");
  if (!b->bound.valid) { g_err = "cmpc_batch_solve_bound: call cmpc_batch_bind_host first"; return CMPC_E_STATE; }
  if (!b->is_setup) { g_err = "cmpc_batch_solve_bound: call cmpc_batch_setup first"; return CMPC_E_STATE; }
  if (count < 0 || count > b->capacity) return fail_arg("cmpc_batch_solve_bound: count exceeds capacity");
  return solve_host_core(b, count, b->bound);
}

// device tables and per-instance arrays of the disturbance estimator stage, allocated on first use
static int ensure_adapt_buffers(cmpc_batch* b) {
  if (b->d_twiddle) return CMPC_OK;
  const size_t cap = (size_t)b->capacity, N = CMPC_ADAPT_WINDOW;
  std::vector<double> tw;
  std::vector<float> gk;
  build_adapt_tables(tw, gk);
  CK(cudaMalloc(&b->d_twiddle, sizeof(double) * tw.size()));
  CK(cudaMalloc(&b->d_gk, sizeof(float) * gk.size()));
  CK(cudaMalloc(&b->d_win_t, sizeof(float) * cap * N));
  CK(cudaMalloc(&b->d_win_d, sizeof(float) * cap * N));
  CK(cudaMalloc(&b->d_simtime, sizeof(float) * cap));
  CK(cudaMalloc(&b->d_est, sizeof(double) * cap * 4));
  CK(cudaMalloc(&b->d_fest, sizeof(float) * cap * 6));
  CK(cudaMemset(b->d_est, 0, sizeof(double) * cap * 4));
  CK(cudaMemset(b->d_fest, 0, sizeof(float) * cap * 6));
  CK(cudaMemset(b->d_win_t, 0, sizeof(float) * cap * N));
  CK(cudaMemset(b->d_win_d, 0, sizeof(float) * cap * N));
  CK(cudaMemset(b->d_simtime, 0, sizeof(float) * cap));
  CK(cudaMemcpy(b->d_twiddle, tw.data(), sizeof(double) * tw.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b->d_gk, gk.data(), sizeof(float) * gk.size(), cudaMemcpyHostToDevice));
  return CMPC_OK;
}

int cmpc_batch_upload_disturbance(cmpc_batch* b, int count, const float* windows_t, const float* windows_d,
                                  const float* sim_time, int mode) {
  if (!b) return fail_arg("cmpc_batch_upload_disturbance: null batch");
  CK(cudaSetDevice(b->device));
  { int rcs = sync_all(b); if (rcs) return rcs; }  // solves in flight read the old windows
  if (mode < 0 || (!windows_t && mode != 2)) {
    b->adapt_mode = -1;
    return CMPC_OK;
  }
  if (mode > 2) return fail_arg("cmpc_batch_upload_disturbance: mode must be 0, 1 or 2");
  if (count < 0 || count > b->capacity) return fail_arg("cmpc_batch_upload_disturbance: count exceeds capacity");
  if (!sim_time) return fail_arg("cmpc_batch_upload_disturbance: null sim_time");
  const size_t N = CMPC_ADAPT_WINDOW;
  { int rca = ensure_adapt_buffers(b); if (rca) return rca; }
  cudaStream_t st = b->stream[0];
  if (windows_t && windows_d && mode != 2) {
    CK(cudaMemcpyAsync(b->d_win_t, windows_t, sizeof(float) * (size_t)count * N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b->d_win_d, windows_d, sizeof(float) * (size_t)count * N, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemcpyAsync(b->d_simtime, sim_time, sizeof(float) * (size_t)count, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));  // the host arrays may be pageable and reused by the caller
  b->adapt_mode = mode;
  return CMPC_OK;
}

int cmpc_batch_download_disturbance(cmpc_batch* b, double* est, float* f_est) {
  if (!b) return fail_arg("cmpc_batch_download_disturbance: null batch");
  if (!b->d_est) { g_err = "cmpc_batch_download_disturbance: no disturbance data was uploaded"; return CMPC_E_STATE; }
  CK(cudaSetDevice(b->device));
  { int rcs = sync_all(b); if (rcs) return rcs; }
  if (est) CK(cudaMemcpy(est, b->d_est, sizeof(double) * (size_t)b->count * 4, cudaMemcpyDeviceToHost));
  if (f_est) CK(cudaMemcpy(f_est, b->d_fest, sizeof(float) * (size_t)b->count * 6, cudaMemcpyDeviceToHost));
  return CMPC_OK;
}

int cmpc_batch_set_weights(cmpc_batch* b, const float weights[12], float alpha) {
  if (!b || !weights) return fail_arg("cmpc_batch_set_weights: null argument");
  for (int i = 0; i < 12; i++) b->weights[i] = weights[i];
  b->alpha = alpha;
  return CMPC_OK;
}

// contact foot-steps of the table getMpcTable will produce for this command (an upper bound is enough: it sizes the launch)
static int command_contacts(const cmpc_command& c, int h) {
  if (c.gait_kind == CMPC_GAIT_MIXED_FREQUENCY) {
    int n = 0;
    for (int k = 0; k < h; k++)
      for (int j = 0; j < 4; j++) {
        const int period = c.gait_offsets[j] > 0 ? c.gait_offsets[j] : 1;
        n += (float)((k + c.gait_iteration + 1) % period) < (float)period * c.gait_duty;
      }
    return n;
  }
  int n = 0;  // the table covers exactly one period: leg j is down for min(duration, h) of the h steps
  for (int j = 0; j < 4; j++) n += std::max(0, std::min(c.gait_durations[j], h));
  return n;
}

int cmpc_batch_solve_commands(cmpc_batch* b, int count, const cmpc_command* commands, cmpc_command_result* results,
                              double* forces_out) {
  if (!b || !commands || !results) return fail_arg("cmpc_batch_solve_commands: null argument");
  if (!b->is_setup) { g_err = "cmpc_batch_solve_commands: call cmpc_batch_setup first"; return CMPC_E_STATE; }
  if (count < 0 || count > b->capacity) return fail_arg("cmpc_batch_solve_commands: count exceeds capacity");
  static_assert(sizeof(cmpc_command) == 464 && sizeof(cmpc_command_result) == 144, "command layout");
  CK(cudaSetDevice(b->device));
  { int rcs = sync_aux(b); if (rcs) return rcs; }
  const size_t cap = (size_t)b->capacity;
  if (!b->d_cmds) {
    CK(cudaMalloc(&b->d_cmds, cap * sizeof(cmpc_command)));
    CK(cudaMalloc(&b->d_results, cap * sizeof(cmpc_command_result)));
    CK(cudaMalloc(&b->d_fext, cap * 6 * sizeof(float)));
    CK(cudaMemset(b->d_fext, 0, cap * 6 * sizeof(float)));
  }
  { int rca = ensure_adapt_buffers(b); if (rca) return rca; }
  if (count == 0) return CMPC_OK;
  if (b->hist_len > 0 && count != b->hist_count) {
    g_err = "cmpc_batch_solve_commands: count differs from the count the disturbance histories were built with "
            "(one sample counter per batch); call cmpc_batch_reset_history first";
    return CMPC_E_STATE;
  }
  b->hist_count = count;
  cudaStream_t st = b->stream[0];
  const int h = b->h;
  CK(cudaEventRecord(b->ev0, st));
  CK(cudaMemcpyAsync(b->d_cmds, commands, (size_t)count * sizeof(cmpc_command), cudaMemcpyHostToDevice, st));
  int maxc = 0;
  for (int i = 0; i < count; i++) maxc = std::max(maxc, command_contacts(commands[i], h));
  maxc = std::min(maxc, 4 * h);
  float alpha = b->alpha;
  if (alpha > 1e-4f) alpha = 1e-5f;  // "Alpha was set too high", ConvexMPCLocomotion.cpp:785-789
  int rc = cmpc_launch_frontend(b->d_cmds, b->d_rec, b->d_results, b->d_fext, b->d_simtime, count, h, b->rec_stride,
                                (float)b->dt, alpha, b->weights, st);
  if (rc != 0) return fail_cuda((cudaError_t)rc, "cmpc_frontend_kernel launch");
  b->launches++;
  // solve_mpc's bookkeeping (SolverMPC.cpp:688-706, :806-813): push the sample; fit while the history holds
  // 400..500 samples (the estimate is not applied yet); beyond 500 the stored fit is refreshed at
  // simulation_time and applied in g.  The windows are not read any more once the history is past 500.
  if (b->hist_len < 500) {
    rc = cmpc_launch_history_push(b->d_cmds, b->d_fext, b->d_win_t, b->d_win_d, count, b->hist_len, st);
    if (rc != 0) return fail_cuda((cudaError_t)rc, "cmpc_history_push_kernel launch");
    b->launches++;
  }
  if (b->hist_len < (1 << 30)) b->hist_len++;
  b->adapt_mode = b->hist_len < CMPC_ADAPT_WINDOW ? -1 : (b->hist_len <= 500 ? 0 : 2);
  b->throughput_mode = false;
  b->count = count;
  b->max_contact = maxc;
  rc = launch_range(b, 0, count, maxc, 0);
  if (rc) return rc;
  rc = cmpc_launch_epilogue(b->d_cmds, b->d_forces, b->d_status, b->d_iters, b->d_results, count, h, st);
  if (rc != 0) return fail_cuda((cudaError_t)rc, "cmpc_epilogue_kernel launch");
  b->launches++;
  CK(cudaMemcpyAsync(results, b->d_results, (size_t)count * sizeof(cmpc_command_result), cudaMemcpyDeviceToHost, st));
  if (forces_out)
    CK(cudaMemcpyAsync(forces_out, b->d_forces, sizeof(double) * (size_t)count * 12 * h, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(b->ev1, st));
  b->timed = true;
  CK(cudaStreamSynchronize(st));
  return CMPC_OK;
}

int cmpc_batch_reset_history(cmpc_batch* b) {
  if (!b) return fail_arg("cmpc_batch_reset_history: null batch");
  CK(cudaSetDevice(b->device));
  { int rcs = sync_all(b); if (rcs) return rcs; }
  b->hist_len = 0;
  b->hist_count = 0;
  b->adapt_mode = -1;
  const size_t cap = (size_t)b->capacity;
  if (b->d_fext) CK(cudaMemset(b->d_fext, 0, cap * 6 * sizeof(float)));
  if (b->d_est) {
    CK(cudaMemset(b->d_est, 0, sizeof(double) * cap * 4));
    CK(cudaMemset(b->d_fest, 0, sizeof(float) * cap * 6));
  }
  return CMPC_OK;
}

int cmpc_batch_history_length(cmpc_batch* b, int* samples) {
  if (!b || !samples) return fail_arg("cmpc_batch_history_length: null argument");
  *samples = b->hist_len;
  return CMPC_OK;
}

int cmpc_batch_copy_records(cmpc_batch* b, int first, int count, void* dst) {
  if (!b || !dst || first < 0 || count < 0 || first + count > b->capacity) return fail_arg("cmpc_batch_copy_records: bad arguments");
  if (!b->is_setup) { g_err = "cmpc_batch_copy_records: call cmpc_batch_setup first"; return CMPC_E_STATE; }
  CK(cudaSetDevice(b->device));
  { int rcs = sync_all(b); if (rcs) return rcs; }
  CK(cudaMemcpy(dst, b->d_rec + (size_t)first * b->rec_stride, (size_t)count * b->rec_stride, cudaMemcpyDeviceToHost));
  return CMPC_OK;
}

int cmpc_batch_device_records(cmpc_batch* b, void** records, size_t* stride_bytes) {
  if (!b || !records || !stride_bytes) return fail_arg("cmpc_batch_device_records: null argument");
  *records = b->d_rec;
  *stride_bytes = (size_t)b->rec_stride;
  return CMPC_OK;
}

int cmpc_batch_device_forces(cmpc_batch* b, void** forces) {
  if (!b || !forces) return fail_arg("cmpc_batch_device_forces: null argument");
  *forces = b->d_forces;
  return CMPC_OK;
}

int cmpc_batch_last_solve_ms(cmpc_batch* b, float* ms) {
  if (!b || !ms) return fail_arg("cmpc_batch_last_solve_ms: null argument");
  if (!b->timed) { g_err = "cmpc_batch_last_solve_ms: nothing solved yet"; return CMPC_E_STATE; }
  CK(cudaSetDevice(b->device));
  CK(cudaEventSynchronize(b->ev1));
  CK(cudaEventElapsedTime(ms, b->ev0, b->ev1));
  return CMPC_OK;
}

int cmpc_host_register(void* ptr, size_t bytes) {
  if (!ptr || bytes == 0) return fail_arg("cmpc_host_register: bad arguments");
  CK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return CMPC_OK;
}

int cmpc_host_unregister(void* ptr) {
  if (!ptr) return fail_arg("cmpc_host_unregister: null pointer");
  CK(cudaHostUnregister(ptr));
  return CMPC_OK;
}

int cmpc_batch_mark(cmpc_batch* b, int which) {
  if (!b || (which != 0 && which != 1)) return fail_arg("cmpc_batch_mark: bad arguments");
  CK(cudaSetDevice(b->device));
  { int rcj = join_streams(b); if (rcj) return rcj; }
  CK(cudaEventRecord(which ? b->mark1 : b->mark0, b->stream[0]));
  { int rcf = fork_streams(b); if (rcf) return rcf; }  // stream 1 starts nothing before the mark
  return CMPC_OK;
}

int cmpc_batch_marked_ms(cmpc_batch* b, float* ms) {
  if (!b || !ms) return fail_arg("cmpc_batch_marked_ms: null argument");
  CK(cudaSetDevice(b->device));
  CK(cudaEventSynchronize(b->mark1));
  CK(cudaEventElapsedTime(ms, b->mark0, b->mark1));
  return CMPC_OK;
}

int cmpc_batch_reset_counters(cmpc_batch* b) {
  if (!b) return fail_arg("cmpc_batch_reset_counters: null batch");
  CK(cudaSetDevice(b->device));
  { int rcj = join_streams(b); if (rcj) return rcj; }
  CK(cudaMemsetAsync(b->d_flops, 0, sizeof(unsigned long long) * CMPC_K_COUNT, b->stream[0]));
  b->launches = 0;
  return CMPC_OK;
}

int cmpc_batch_kernel_launches(cmpc_batch* b, long long* launches) {
  if (!b || !launches) return fail_arg("cmpc_batch_kernel_launches: null argument");
  *launches = b->launches;
  return CMPC_OK;
}

int cmpc_batch_last_flops(cmpc_batch* b, double* flops) {
  if (!b || !flops) return fail_arg("cmpc_batch_last_flops: null argument");
  CK(cudaSetDevice(b->device));
  { int rcs = sync_aux(b); if (rcs) return rcs; }
  CK(cudaMemcpyAsync(b->h_flops, b->d_flops, sizeof(unsigned long long) * CMPC_K_COUNT, cudaMemcpyDeviceToHost, b->stream[0]));
  CK(cudaStreamSynchronize(b->stream[0]));
  *flops = 0.0;
  for (int i = 0; i < CMPC_K_COUNT; i++) *flops += (double)b->h_flops[i];
  return CMPC_OK;
}

int cmpc_batch_kernel_flops(cmpc_batch* b, double flops[4]) {
  if (!b || !flops) return fail_arg("cmpc_batch_kernel_flops: null argument");
  double total = 0.0;
  int rc = cmpc_batch_last_flops(b, &total);
  if (rc) return rc;
  for (int i = 0; i < CMPC_K_COUNT; i++) flops[i] = (double)b->h_flops[i];
  return CMPC_OK;
}

// One solve of [first, first+count) on the batch stream with a CUDA event between the kernel classes.
int cmpc_batch_profile_range(cmpc_batch* b, int first, int count, float ms[4]) {
  if (!b || !ms) return fail_arg("cmpc_batch_profile_range: null argument");
  if (!b->is_setup) { g_err = "cmpc_batch_profile_range: call cmpc_batch_setup first"; return CMPC_E_STATE; }
  if (first < 0 || count < 0 || first + count > b->count) return fail_arg("cmpc_batch_profile_range: range outside the uploaded instances");
  CK(cudaSetDevice(b->device));
  { int rcs = sync_all(b); if (rcs) return rcs; }
  if (!b->prof_ev[0])
    for (int i = 0; i < CMPC_K_COUNT + 1; i++) CK(cudaEventCreate(&b->prof_ev[i]));
  for (int i = 0; i < CMPC_K_COUNT; i++) { ms[i] = 0.f; b->prof_ms[i] = 0.f; }
  b->profiling = true;
  b->throughput_mode = !b->serial;  // the kernels as solve_range configures them
  int rc = launch_range(b, first, count, b->max_contact, 0);
  b->profiling = false;
  if (rc) return rc;
  CK(cudaStreamSynchronize(b->stream[0]));
  for (int i = 0; i < CMPC_K_COUNT; i++) ms[i] = b->prof_ms[i];
  return CMPC_OK;
}

int cmpc_batch_enable_phase_clocks(cmpc_batch* b, int on) {
  if (!b) return fail_arg("cmpc_batch_enable_phase_clocks: null batch");
  CK(cudaSetDevice(b->device));
  { int rcs = sync_all(b); if (rcs) return rcs; }
  if (on && !b->d_phase) CK(cudaMalloc(&b->d_phase, sizeof(unsigned long long) * CMPC_PH_COUNT));
  if (!on && b->d_phase) {
    CK(cudaFree(b->d_phase));
    b->d_phase = nullptr;
  }
  if (b->d_phase) CK(cudaMemset(b->d_phase, 0, sizeof(unsigned long long) * CMPC_PH_COUNT));
  return CMPC_OK;
}

int cmpc_batch_phase_cycles(cmpc_batch* b, unsigned long long* cycles, int n) {
  if (!b || !cycles || n < 1) return fail_arg("cmpc_batch_phase_cycles: bad arguments");
  if (!b->d_phase) { g_err = "cmpc_batch_phase_cycles: phase clocks are not enabled"; return CMPC_E_STATE; }
  CK(cudaSetDevice(b->device));
  { int rcs = sync_all(b); if (rcs) return rcs; }
  unsigned long long tmp[CMPC_PH_COUNT];
  CK(cudaMemcpy(tmp, b->d_phase, sizeof(tmp), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) cycles[i] = i < CMPC_PH_COUNT ? tmp[i] : 0ull;
  return CMPC_OK;
}

static int measure_peak(int device, double* tflops, bool tensor);
int cmpc_measure_fp64_peak(int device, double* tflops) { return measure_peak(device, tflops, false); }
int cmpc_measure_dmma_peak(int device, double* tflops) { return measure_peak(device, tflops, true); }

static int measure_peak(int device, double* tflops, bool tensor) {
  if (!tflops) return fail_arg("cmpc_measure_fp64_peak: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    g_err = "cmpc_measure_fp64_peak: no such CUDA device";
    return CMPC_E_NODEVICE;
  }
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  double* d = nullptr;
  CK(cudaMalloc(&d, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int iters = 20000;
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0, 0));
    int rc = tensor ? cmpc_run_dmma_peak(prop.multiProcessorCount, nullptr, d, iters / 4)
                    : cmpc_run_dfma_peak(prop.multiProcessorCount, nullptr, d, iters);
    if (rc) return fail_cuda((cudaError_t)rc, "cmpc_dfma_peak_kernel");
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    // DFMA: 8 chains x 2 flops x 256 threads x 8 CTAs per SM; DMMA: 16 accumulators x 512 flops x 8 warps x 4 CTAs per SM
    double fl = tensor ? 16.0 * 512.0 * (iters / 4) * 8.0 * 4.0 * prop.multiProcessorCount
                       : 2.0 * 8.0 * iters * 256.0 * 8.0 * prop.multiProcessorCount;
    best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return CMPC_OK;
}

// ---------------------------------------------------------------------------
// reference single-instance interface (convexMPC_interface.cpp)
// ---------------------------------------------------------------------------
namespace {
struct Single {
  std::mutex mu;
  cmpc_batch* b = nullptr;
  int horizon = 0;
  bool has_solved = false;
  float x_drag = 0.f;
  float f_ext[6] = {0, 0, 0, 0, 0, 0};
  float sim_time = 0.f;
  float f_est[6] = {0, 0, 0, 0, 0, 0};
  float f_est_smoothed[6] = {0, 0, 0, 0, 0, 0};  // SolverMPC.cpp:783
  float f_est_static[6] = {0, 0, 0, 0, 0, 0};    // SolverMPC.cpp:798
  // adaptive bookkeeping of solve_mpc(), SolverMPC.cpp:688-798
  std::vector<float> time_history, diff_history;
  std::vector<double> q_soln, q_new;
  int policy = CMPC_ON_ERROR_ABORT;
  int last_status = CMPC_ST_SOLVED;
};
Single& single() {
  static Single s;
  return s;
}
void die(const char* where) {
  std::fprintf(stderr, "[cmpc_b200] %s failed: %s\n", where, cmpc_last_error());
  std::abort();  // no CPU fallback: a controller must not keep running on stale forces
}
}  // namespace

void setup_problem(double dt, int horizon, double mu, double f_max) {
  Single& s = single();
  std::lock_guard<std::mutex> lk(s.mu);
  if (!s.b) {
    const char* dev = std::getenv("CMPC_DEVICE");
    if (cmpc_batch_create(&s.b, dev ? std::atoi(dev) : 0, 1) != CMPC_OK) die("setup_problem/cmpc_batch_create");
  }
  if (cmpc_batch_setup(s.b, dt, horizon, mu, f_max) != CMPC_OK) die("setup_problem");
  s.horizon = horizon;
  s.q_soln.assign(12 * horizon, 0.0);
  s.q_new.assign(12 * horizon, 0.0);
}

static void solve_single(Single& s, const float* p, const float* v, const float* q, const float* w, const float* r,
                         const float* weights, const float* traj, float alpha, const int* gait) {
  if (!s.b) die("update_problem_data before setup_problem");
  const int h = s.horizon;
  std::vector<unsigned char> g8(4 * h);
  for (int i = 0; i < 4 * h; i++) g8[i] = (unsigned char)gait[i];  // mint_to_u8, convexMPC_interface.cpp:76
  // periodic-disturbance bookkeeping, SolverMPC.cpp:688-798: push the sample; fit the sinusoid while the
  // history holds 400..500 samples; refresh f_est[3] from the stored fit afterwards; g sees f_est beyond 500
  const int N = CMPC_ADAPT_WINDOW;
  s.diff_history.push_back(s.f_ext[3]);
  s.time_history.push_back(s.sim_time);
  const size_t len = s.time_history.size();
  int mode = -1;
  if (len >= (size_t)N) mode = (len <= 500) ? 0 : 2;
  if (mode == 0) {
    if (cmpc_batch_upload_disturbance(s.b, 1, s.time_history.data() + (len - N), s.diff_history.data() + (len - N),
                                      &s.sim_time, 0) != CMPC_OK)
      die("update_problem_data/disturbance");
  } else if (mode == 2) {
    if (cmpc_batch_upload_disturbance(s.b, 1, nullptr, nullptr, &s.sim_time, 2) != CMPC_OK)
      die("update_problem_data/disturbance");
  }
  if (len > 4096) {  // the reference lets the vectors grow without bound; only the last window is ever read
    s.time_history.erase(s.time_history.begin(), s.time_history.end() - 1024);
    s.diff_history.erase(s.diff_history.begin(), s.diff_history.end() - 1024);
  }
  float fd[6] = {0, 0, 0, 0, 0, 0};
  cmpc_inputs in;
  in.p = p; in.v = v; in.q = q; in.w = w; in.r = r; in.weights = weights; in.traj = traj;
  in.alpha = &alpha; in.gait = g8.data(); in.x_drag = &s.x_drag; in.f_dist = fd;
  cmpc_outputs out;
  std::memset(&out, 0, sizeof(out));
  int32_t status = CMPC_ST_SOLVED;
  out.forces = s.q_new.data();
  out.status = &status;
  if (cmpc_batch_solve_host(s.b, 1, &in, &out) != CMPC_OK) die("update_problem_data");
  if (mode >= 0 && cmpc_batch_download_disturbance(s.b, nullptr, s.f_est) != CMPC_OK) die("update_problem_data/f_est");
  // the two filtered estimates solve_mpc keeps up to date on every call (SolverMPC.cpp:783, :798), float arithmetic
  for (int i = 0; i < 6; i++) s.f_est_smoothed[i] = 0.95f * s.f_est_smoothed[i] + 0.05f * s.f_est[i];
  s.f_est_static[3] = 0.97f * s.f_est_static[3] + 0.03f * s.f_ext[3];
  // a solve that did not reach the optimum, or that met NaN / Inf on the way, must not reach the legs: the dual
  // iterate is primal infeasible before convergence (it can leave the friction cone), NaN is NaN
  bool finite = true;
  for (double f : s.q_new) finite = finite && std::isfinite(f);
  if (!finite && (status == CMPC_ST_SOLVED || status == CMPC_ST_EMPTY)) status = CMPC_ST_NONFINITE;
  s.last_status = status;
  if (status != CMPC_ST_SOLVED && status != CMPC_ST_EMPTY) {
    std::fprintf(stderr, "[cmpc_b200] update_problem_data: solve ended with status %d (%s)\n", (int)status,
                 status == CMPC_ST_NONFINITE ? "non-finite input, disturbance estimate or iterate"
                                             : status == CMPC_ST_MAXITER ? "iteration cap" : "see CMPC_ST_*");
    if (s.policy == CMPC_ON_ERROR_ABORT) {
      g_err = "solve did not reach the optimum";
      die("update_problem_data");
    }
    s.has_solved = true;  // CMPC_ON_ERROR_HOLD: get_solution() keeps serving the last good forces
    return;
  }
  s.q_soln.swap(s.q_new);
  s.has_solved = true;
}

void update_problem_data(double* p, double* v, double* q, double* w, double* r, double yaw, double* weights,
                         double* state_trajectory, double alpha, int* gait) {
  (void)yaw;
  Single& s = single();
  std::lock_guard<std::mutex> lk(s.mu);
  const int h = s.horizon;
  float pf[3], vf[3], qf[4], wf[3], rf[12], wt[12];
  std::vector<float> tr(12 * std::max(h, 1));
  for (int i = 0; i < 3; i++) { pf[i] = (float)p[i]; vf[i] = (float)v[i]; wf[i] = (float)w[i]; }
  for (int i = 0; i < 4; i++) qf[i] = (float)q[i];
  for (int i = 0; i < 12; i++) { rf[i] = (float)r[i]; wt[i] = (float)weights[i]; }
  for (int i = 0; i < 12 * h; i++) tr[i] = (float)state_trajectory[i];
  solve_single(s, pf, vf, qf, wf, rf, wt, tr.data(), (float)alpha, gait);
}

void update_problem_data_floats(float* p, float* v, float* q, float* w, float* r, float roll, float pitch, float yaw,
                                float* weights, float* state_trajectory, float alpha, int* gait) {
  (void)roll; (void)pitch; (void)yaw;  // carried by update_data_t but unused by solve_mpc (RobotState.cpp:46)
  Single& s = single();
  std::lock_guard<std::mutex> lk(s.mu);
  solve_single(s, p, v, q, w, r, weights, state_trajectory, alpha, gait);
}

double get_solution(int index) {
  Single& s = single();
  std::lock_guard<std::mutex> lk(s.mu);
  if (!s.has_solved) return 0.0;  // convexMPC_interface.cpp:158
  if (index < 0 || index >= (int)s.q_soln.size()) return 0.0;
  return s.q_soln[index];
}

void update_solver_settings(int max_iter, double rho, double sigma, double solver_alpha, double terminate,
                            double use_jcqp) {
  (void)max_iter; (void)rho; (void)sigma; (void)solver_alpha; (void)terminate; (void)use_jcqp;
}

void update_x_drag(float x_drag) {
  Single& s = single();
  std::lock_guard<std::mutex> lk(s.mu);
  s.x_drag = x_drag;
}

void cmpc_set_external_force(const float f_ext[6]) {
  Single& s = single();
  std::lock_guard<std::mutex> lk(s.mu);
  for (int i = 0; i < 6; i++) s.f_ext[i] = f_ext[i];
}

void cmpc_set_simulation_time(float t) {
  Single& s = single();
  std::lock_guard<std::mutex> lk(s.mu);
  s.sim_time = t;
}

void cmpc_get_disturbance_estimate(float f_est[6]) {
  Single& s = single();
  std::lock_guard<std::mutex> lk(s.mu);
  for (int i = 0; i < 6; i++) f_est[i] = s.f_est[i];
}

void cmpc_get_disturbance_estimate_smoothed(float f_est_smoothed[6]) {
  Single& s = single();
  std::lock_guard<std::mutex> lk(s.mu);
  for (int i = 0; i < 6; i++) f_est_smoothed[i] = s.f_est_smoothed[i];
}

void cmpc_get_disturbance_estimate_static(float f_est_static[6]) {
  Single& s = single();
  std::lock_guard<std::mutex> lk(s.mu);
  for (int i = 0; i < 6; i++) f_est_static[i] = s.f_est_static[i];
}

void cmpc_set_error_policy(int policy) {
  Single& s = single();
  std::lock_guard<std::mutex> lk(s.mu);
  s.policy = policy == CMPC_ON_ERROR_HOLD ? CMPC_ON_ERROR_HOLD : CMPC_ON_ERROR_ABORT;
}

int cmpc_last_status(void) {
  Single& s = single();
  std::lock_guard<std::mutex> lk(s.mu);
  return s.last_status;
}

void cmpc_reset_history(void) {
  Single& s = single();
  std::lock_guard<std::mutex> lk(s.mu);
  s.time_history.clear();
  s.diff_history.clear();
  for (int i = 0; i < 6; i++) s.f_est[i] = s.f_est_smoothed[i] = s.f_est_static[i] = 0.f;
  if (s.b && s.b->d_fest) {
    cudaSetDevice(s.b->device);
    cudaMemset(s.b->d_fest, 0, sizeof(float) * 6 * (size_t)s.b->capacity);
    cudaMemset(s.b->d_est, 0, sizeof(double) * 4 * (size_t)s.b->capacity);
  }
  s.b ? (void)(s.b->adapt_mode = -1) : (void)0;
}

}  // extern "C"

// convexMPC_interface.h:52 declares update_x_drag WITHOUT extern "C": an unmodified reference translation unit that
// includes that header references the C++-mangled symbol.  Same function, second name.
void cmpc_update_x_drag_cxx(float x_drag) __asm__("_Z13update_x_dragf");
void cmpc_update_x_drag_cxx(float x_drag) { update_x_drag(x_drag); }
