// fiss_abi.cu -- host side of libfissgpu.so: the C ABI declared in include/fiss_abi.h.
//
// Plain CUDA runtime; no torch, no Python.  A handle owns the uploaded scene tables, pinned host
// staging and device scratch for the *_host entry points; the *_dev entry points only launch on
// caller-owned buffers.  Every launch is a persistent grid sized from the SM count and the
// kernel's occupancy for the shared-memory footprint of the current scene.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nccl.h>  // types and prototypes only: libnccl is resolved at run time (dlopen), never linked

#include "fiss_grid_kernel.cuh"
#include "fiss_pick_exchange.cuh"
#include "fiss_record_kernel.cuh"
#include "fiss_spline_kernels.cuh"

namespace {

std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = std::max(bytes, (size_t)4096);
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = std::max(bytes, (size_t)4096);
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

constexpr size_t kArenaBytesMax = 256 * 1024;  // plan_common: results up to this size travel as one block
constexpr int kMaxParts = 4;                   // plan_common: pieces a big batch is pipelined in
constexpr size_t kSmemLimit = 227 * 1024;      // opt-in maximum per CTA on sm_100
constexpr size_t kSmemObsBudget = 100 * 1024;  // stage obstacle rows only while 2 CTAs/SM still fit

// One kernel launch as data: what the planning code decided (function, shape, arguments).  It is either issued on a
// stream right away or becomes a kernel node of a CUDA graph (StepGraph) whose parameters are patched from call to call.
struct KernelLaunch {
  const void* func = nullptr;
  dim3 grid{1, 1, 1}, block{1, 1, 1};
  size_t smem = 0;
  bool pdl = false;  // programmatic dependent launch behind the previous kernel (direct launches only)
  int n_params = 0;
  uint16_t off[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t arg_bytes = 0;
  alignas(16) unsigned char arg[1024];
  template <typename T>
  void push(const T& v) {
    static_assert(sizeof(T) <= sizeof(arg), "kernel argument too large");
    size_t o = (arg_bytes + alignof(T) - 1) & ~(alignof(T) - 1);
    std::memcpy(arg + o, &v, sizeof(T));
    off[n_params++] = (uint16_t)o;
    arg_bytes = o + sizeof(T);
  }
  void params(void** out) {
    for (int i = 0; i < n_params; ++i) out[i] = arg + off[i];
  }
  bool same_shape(const KernelLaunch& o) const {
    return grid.x == o.grid.x && block.x == o.block.x && smem == o.smem;
  }
  bool same_args(const KernelLaunch& o) const {
    return arg_bytes == o.arg_bytes && std::memcmp(arg, o.arg, arg_bytes) == 0;
  }
};

// A plan step as an instantiated CUDA graph: [H2D of the inputs] -> kernels in order -> [D2H of the results].  Built
// once per topology (same kernels, same copy endpoints); afterwards a call only patches the kernel nodes whose launch
// shape or arguments changed (cudaGraphExecKernelNodeSetParams: time_step_now moves every cycle of a closed loop) and
// launches the whole step with ONE cudaGraphLaunch.
struct StepGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  std::vector<cudaGraphNode_t> knodes;
  std::vector<KernelLaunch> kl;
  const void* h2d_src = nullptr;
  void* h2d_dst = nullptr;
  size_t h2d_bytes = 0;
  const void* d2h_src = nullptr;
  void* d2h_dst = nullptr;
  size_t d2h_bytes = 0;
  int64_t launches = 0, rebuilds = 0, patches = 0;
  void release() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    exec = nullptr;
    graph = nullptr;
    knodes.clear();
    kl.clear();
  }
};

// One in-flight call of the streaming entry points (fiss_plan_grid_submit / _wait): its own device buffers, pinned
// bounce buffers and events, so that the copy-back of one lane runs under the kernels of the other.
constexpr int kLanes = FISS_LANES;
struct Lane {
  DevBuf d_ego, d_cost, d_flags, d_win, d_records;
  PinBuf h_in, h_out;
  StepGraph graph;
  cudaEvent_t kernels_done = nullptr, copied = nullptr;
  bool pending = false;
  // where the results go on wait()
  int32_t B = 0, n_stride = 0;
  int32_t* best_idx = nullptr;
  double* best_cost = nullptr;
  int32_t* best_meta = nullptr;
  double* records = nullptr;
  bool rec_pinned = false;
  void release() {
    for (DevBuf* b : {&d_ego, &d_cost, &d_flags, &d_win, &d_records}) b->release();
    h_in.release();
    h_out.release();
    graph.release();
    if (kernels_done) cudaEventDestroy(kernels_done);
    if (copied) cudaEventDestroy(copied);
    kernels_done = copied = nullptr;
  }
};

}  // namespace

struct fiss_handle {
  int device = 0;
  int sm_count = 148;
  std::string err;
  int64_t launches = 0;
  // scene
  DevBuf spline;
  int K = 0, Kp = 0;
  DevBuf obs_tab, obs_const, obs_raw, obs_lw, obs_valid;
  int M = 0, Mp = 0, mp_shift = 0, T_obs = 0, final_time_step = 0;
  // *_host staging
  DevBuf d_ego, d_end, d_cost, d_flags, d_best_idx, d_best_cost, d_meta, d_records, d_es;  // d_es: in/out block of fiss_eval_end_states_host
  PinBuf h_in, h_out;
  // knot index of the installed spline (fiss_set_spline): uniform cells over the knot range, behind the table in h->spline
  int lut_cells = 0, lut_bytes = 0, lut_iters = 0;
  double lut_inv_h = 0.0;
  DevBuf d_arena;  // plan_common latency path: [winners | records | cost | flags]
  std::vector<double> end_cache;
  // product lattice (fiss_grid): device axes [4][kAxisMax] + the expanded [C][4] table in d_end
  DevBuf d_axes;
  DevBuf d_work;   // work counters of the lattice kernel (zero between launches), two pairs
  int work_parity = 0;
  uint32_t lattice_seq = 0;  // lattice launches issued on this handle (chained launches)
  int lattice_chained = 0;   // ... and whether the last one was chained IN A TRAIN (1: winner-only kernel, 2: materialising)
  PinBuf host_done;          // [1] u32, mapped: number of the last numbered lattice launch that is over (written by its last CTA)
  DevBuf d_shadow;           // chained launches: the first items' cost / flags, per CTA, two sets
  DevBuf d_fit_in, d_fit_out;  // fiss_fit_splines_host / fiss_frame_samples_host
  cudaStream_t capture_stream = nullptr;       // graph_run: launches are captured here when a graph is (re)built
  cudaStream_t copy_stream = nullptr;          // plan_common: D2H of one half of a big batch under the other half's kernels
  cudaEvent_t part_done[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<double> axes_cache;
  int grid_n_max = 0;
  // kernel launches are collected here instead of being issued while a plan step is being turned into a graph
  std::vector<KernelLaunch>* record = nullptr;
  StepGraph graph_dev, graph_host;  // fiss_plan_grid_dev / the small-batch path of the *_host plan calls
  Lane lanes[kLanes];               // fiss_plan_grid_submit / fiss_plan_grid_wait
  DevBuf d_pick;                    // fiss_allreduce_pick: slot table + record exchange buffer
  void* comm = nullptr;             // ncclComm_t created by fiss_comm_init
  int comm_rank = 0, comm_size = 1;
  size_t smem_attr[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  // resident CTAs per SM of kernel `which` at (threads, smem): asked from the runtime once per distinct launch shape
  // (the query costs microseconds on the latency path of every plan() call)
  struct OccKey { int threads = -1; size_t smem = 0; int occ = 1; } occ_cache[12];
  template <typename K>
  cudaError_t occupancy(int which, K kern, int threads, size_t smem, int* out) {
    OccKey& c = occ_cache[which];
    if (c.threads != threads || c.smem != smem) {
      int occ = 1;
      cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
      if (e != cudaSuccess) return e;
      c.threads = threads;
      c.smem = smem;
      c.occ = std::max(occ, 1);
    }
    *out = c.occ;
    return cudaSuccess;
  }
};

namespace {

int32_t fail(fiss_handle* h, int32_t code, const std::string& msg) {
  if (h) h->err = msg;
  else g_create_error = msg;
  return code;
}

#define FISS_CUDA(h, expr)                                                                        \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(h, FISS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));          \
  } while (0)

// Every entry point runs on the handle's device and leaves the caller's current device as it found it (a torch process
// that plans on cuda:1 must not find its later allocations redirected).
struct DeviceGuard {
  int prev = -1, dev;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
    if (prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define FISS_ON_DEVICE(h)             \
  DeviceGuard device_guard_(h->device); \
  FISS_CUDA(h, device_guard_.err)

// True when `p` is page-locked host memory known to the CUDA runtime (cudaHostAlloc / cudaHostRegister / a torch
// pinned tensor): the *_host entry points then DMA straight from / into it instead of bouncing through the
// handle's own pinned staging and a host memcpy.
bool host_is_pinned(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

// ---- NCCL, resolved at run time ----------------------------------------------------------------
// libfissgpu.so does not link libnccl: a single-GPU user needs none, and a torch process must end up with ONE copy of
// the library -- the one torch already loaded -- which dlopen by soname gives.
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
} g_nccl;

bool nccl_load(const char* path) {
  if (g_nccl.AllReduce) return true;
  void* lib = nullptr;
  if (path && *path) lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);  // the copy already in the process
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    const char* e = dlerror();
    g_nccl.err = std::string("cannot load libnccl: ") + (e ? e : "unknown error");
    return false;
  }
  g_nccl.lib = lib;
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(lib, "ncclAllReduce"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.GetErrorString) {
    g_nccl.err = "libnccl lacks an expected symbol";
    g_nccl.AllReduce = nullptr;
    return false;
  }
  return true;
}

#define FISS_NCCL(h, expr)                                                                          \
  do {                                                                                              \
    ncclResult_t r_ = (expr);                                                                       \
    if (r_ != ncclSuccess)                                                                          \
      return fail(h, FISS_ERR_NCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(r_));         \
  } while (0)

int32_t check_params(fiss_handle* h, const fiss_params* p) {
  if (!p) return fail(h, FISS_ERR_INVALID, "params is NULL");
  if (!(p->tick_t > 0.0)) return fail(h, FISS_ERR_INVALID, "tick_t must be > 0");
  if (p->check_res < 1) return fail(h, FISS_ERR_INVALID, "check_res must be >= 1");
  if (p->time_step_now < 0) return fail(h, FISS_ERR_INVALID, "time_step_now must be >= 0");
  return FISS_OK;
}

struct LaunchPlan {
  fiss::EvalArgs a;
  size_t smem = 0;
  int grid = 1;
  int threads = fiss::kThreads;
};

// Issue one kernel: on the stream, or -- while a plan step is being recorded for a graph -- into the record.
int32_t issue(fiss_handle* h, cudaStream_t st, KernelLaunch& kl) {
  if (h->record) {
    h->record->push_back(kl);
    return FISS_OK;
  }
  void* params[8];
  kl.params(params);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = kl.grid;
  cfg.blockDim = kl.block;
  cfg.dynamicSmemBytes = kl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr{};
  if (kl.pdl) {  // programmatic dependent launch behind the kernel that produces this one's inputs
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
  }
  FISS_CUDA(h, cudaLaunchKernelExC(&cfg, kl.func, params));
  h->launches++;
  return FISS_OK;
}

template <bool kMat, bool kRec>
int32_t launch_eval(fiss_handle* h, cudaStream_t st, LaunchPlan& lp, int which) {
  auto kern = fiss::fiss_eval_kernel<kMat, kRec>;
  if (lp.smem > h->smem_attr[which]) {
    FISS_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
    h->smem_attr[which] = kSmemLimit;
  }
  int occ = 1;
  FISS_CUDA(h, h->occupancy(which, kern, lp.threads, lp.smem, &occ));
  const int warps = lp.threads / 32;
  const int64_t need = (lp.a.total + warps - 1) / warps;
  lp.grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)h->sm_count * occ));
  KernelLaunch kl;
  kl.func = (const void*)kern;
  kl.grid = dim3((unsigned)lp.grid);
  kl.block = dim3((unsigned)lp.threads);
  kl.smem = lp.smem;
  kl.pdl = lp.a.after_producer != 0;
  kl.push(lp.a);
  return issue(h, st, kl);
}

// The winners' records with the pick fused in (fiss_record_kernel.cuh): W warps per problem.
template <int W>
int32_t launch_record(fiss_handle* h, cudaStream_t st, const fiss::EvalArgs& a, int which) {
  auto kern = fiss::fiss_record_kernel<W>;
  const size_t smem = fiss::record_smem_bytes(a.Kp, a.n_pad, W);
  if (smem > kSmemLimit) return fail(h, FISS_ERR_CAPACITY, "spline table too large for the record kernel's shared memory");
  if (smem > 48 * 1024 && smem > h->smem_attr[which]) {
    FISS_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
    h->smem_attr[which] = kSmemLimit;
  }
  int occ = 1;
  FISS_CUDA(h, h->occupancy(which, kern, fiss::kRecWarps * 32, smem, &occ));
  constexpr int per_cta = fiss::kRecWarps / W;
  const int64_t need = (a.total + per_cta - 1) / per_cta;
  KernelLaunch kl;
  kl.func = (const void*)kern;
  int64_t ctas = std::min<int64_t>(need, (int64_t)h->sm_count * occ);
  // Behind a CHAINED lattice launch (eval_grid) the record kernel's CTAs take SM slots as that launch's CTAs retire and sit
  // there until it is over, and the next step's lattice kernel cannot be scheduled before every record CTA is resident: a
  // small record grid (it loops over the problems) leaves the retiring slots to the next step sooner.  Nobody waits for this
  // kernel but the next step's second items and the next record kernel.  (FISS_REC_CTAS overrides.)
  // Measured on the B200 (cfg4 step trains): 32 CTAs behind the materialising kernel (0.125 -> 0.108 ms per step; 24: 0.112,
  // 48: 0.110, 64: 0.115, 128: 0.120); behind the winner-only kernel a record CTA fits the slot of one retiring lattice CTA and
  // the full grid is best (0.068 -> 0.056; 64: 0.058, 32: 0.066).
  static const int rec_cap = std::getenv("FISS_REC_CTAS") ? std::max(1, std::atoi(std::getenv("FISS_REC_CTAS"))) : 0;
  if (a.after_producer && h->record == nullptr && h->lattice_chained && (rec_cap > 0 || h->lattice_chained == 2))
    ctas = std::min<int64_t>(ctas, rec_cap > 0 ? rec_cap : 32);
  kl.grid = dim3((unsigned)std::max<int64_t>(1, ctas));
  kl.block = dim3((unsigned)(fiss::kRecWarps * 32));
  kl.smem = smem;
  kl.pdl = a.after_producer != 0;
  kl.push(a);
  return issue(h, st, kl);
}

// Fill everything of EvalArgs that depends on the scene and on the step bound `n_bound`.
int32_t plan_launch(fiss_handle* h, const fiss_params* p, int64_t total, int n_bound, LaunchPlan& lp) {
  if (h->K < 2) return fail(h, FISS_ERR_STATE, "fiss_set_spline has not been called");
  if (n_bound < 1) return fail(h, FISS_ERR_INVALID, "n_stride must be >= the largest step count n (>= 1)");
  fiss::EvalArgs& a = lp.a;
  a.p = *p;
  a.total = total;
  a.spline = h->spline.as<double>();
  a.K = h->K;
  a.Kp = h->Kp;
  int it = 0;
  while ((1 << it) < std::max(h->K - 1, 1)) ++it;
  a.search_iters = it;
  a.obs_tab = h->obs_tab.as<double>();
  a.obs_const = h->obs_const.as<double>();
  a.M = h->M;
  a.Mp = h->Mp;
  a.mp_shift = h->mp_shift;
  a.T_obs = h->T_obs;
  a.final_time_step = h->final_time_step;
  a.n_pad = (n_bound + 1) & ~1;
  a.e_cap = (a.n_pad + p->check_res - 1) / p->check_res;
  a.e_cap = (a.e_cap + 1) & ~1;
  const int horizon = std::max(0, std::min(n_bound, h->final_time_step - p->time_step_now));
  a.E_max = h->M > 0 ? (horizon + p->check_res - 1) / p->check_res : 0;
  // few candidates: fewer warps per CTA so that the work spreads over more SMs (latency case, B = 1)
  int warps = (int)std::min<int64_t>(fiss::kWarpsPerCta, std::max<int64_t>(1, (total + h->sm_count - 1) / h->sm_count));
  lp.threads = warps * 32;
  const size_t base = 16 + (size_t)9 * a.Kp * 8 + (size_t)4 * a.Mp * 8 +
                      (size_t)warps * fiss::scratch_doubles_per_warp(a.n_pad, a.e_cap) * 8;
  const size_t obs_bytes = (size_t)a.E_max * a.Mp * 32;
  a.obs_in_smem = (a.E_max > 0 && base + obs_bytes <= kSmemObsBudget) ? 1 : 0;
  if (!a.obs_in_smem) a.E_max = 0;
  lp.smem = base + (a.obs_in_smem ? obs_bytes : 0);
  if (lp.smem > kSmemLimit)
    return fail(h, FISS_ERR_CAPACITY, "spline table + per-warp scratch exceed 227 KB of shared memory");
  return FISS_OK;
}

// Run a recorded plan step as a graph (see StepGraph).  `kl` is what the planning code just recorded.
bool graphs_enabled() {
  static const bool off = std::getenv("FISS_NO_GRAPH") != nullptr;  // A/B switch: issue the launches one by one
  return !off;
}

int32_t graph_build_by_capture(fiss_handle* h, StepGraph& G, std::vector<KernelLaunch>& kl) {
  // The launches are captured on a private stream (the caller's may be the legacy default stream, which cannot be
  // captured): a launch with the programmatic-serialisation attribute is recorded as a programmatic edge by the runtime.
  if (!h->capture_stream) FISS_CUDA(h, cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking));
  FISS_CUDA(h, cudaStreamBeginCapture(h->capture_stream, cudaStreamCaptureModeThreadLocal));
  int32_t rc_cap = FISS_OK;
  const int64_t launches_before = h->launches;
  for (size_t i = 0; i < kl.size() && rc_cap == FISS_OK; ++i) rc_cap = issue(h, h->capture_stream, kl[i]);
  h->launches = launches_before;
  cudaError_t e_end = cudaStreamEndCapture(h->capture_stream, &G.graph);
  if (rc_cap != FISS_OK) return rc_cap;
  FISS_CUDA(h, e_end);
  size_t n_nodes = 0;
  FISS_CUDA(h, cudaGraphGetNodes(G.graph, nullptr, &n_nodes));
  std::vector<cudaGraphNode_t> nodes(n_nodes);
  FISS_CUDA(h, cudaGraphGetNodes(G.graph, nodes.data(), &n_nodes));
  G.knodes.assign(kl.size(), nullptr);
  for (cudaGraphNode_t n : nodes) {
    cudaGraphNodeType ty;
    FISS_CUDA(h, cudaGraphNodeGetType(n, &ty));
    if (ty != cudaGraphNodeTypeKernel) continue;
    cudaKernelNodeParams kp{};
    FISS_CUDA(h, cudaGraphKernelNodeGetParams(n, &kp));
    for (size_t i = 0; i < kl.size(); ++i)
      if (!G.knodes[i] && kp.func == kl[i].func) {
        G.knodes[i] = n;
        break;
      }
  }
  for (cudaGraphNode_t n : G.knodes)
    if (!n) return fail(h, FISS_ERR_CUDA, "graph capture: a kernel node was not found");
  return FISS_OK;
}

int32_t graph_build_explicit(fiss_handle* h, StepGraph& G, std::vector<KernelLaunch>& kl, const void* h2d_src, void* h2d_dst,
                             size_t h2d_bytes, const void* d2h_src, void* d2h_dst, size_t d2h_bytes) {
  FISS_CUDA(h, cudaGraphCreate(&G.graph, 0));
  cudaGraphNode_t prev = nullptr;
  if (h2d_bytes) {
    cudaGraphNode_t n;
    FISS_CUDA(h, cudaGraphAddMemcpyNode1D(&n, G.graph, nullptr, 0, h2d_dst, h2d_src, h2d_bytes, cudaMemcpyHostToDevice));
    prev = n;
  }
  for (size_t i = 0; i < kl.size(); ++i) {
    void* params[8];
    kl[i].params(params);
    cudaKernelNodeParams kp{};
    kp.func = const_cast<void*>(kl[i].func);
    kp.gridDim = kl[i].grid;
    kp.blockDim = kl[i].block;
    kp.sharedMemBytes = (unsigned)kl[i].smem;
    kp.kernelParams = params;
    cudaGraphNode_t n;
    static const bool no_pdl_edge = std::getenv("FISS_GRAPH_NO_PDL") != nullptr;
    if (kl[i].pdl && prev && i > 0 && !no_pdl_edge) {
      // programmatic edge from the producer kernel: this kernel may start (and stage its tables) once every CTA of
      // the producer has signalled griddepcontrol.launch_dependents; it waits for the producer's data itself
      FISS_CUDA(h, cudaGraphAddKernelNode(&n, G.graph, nullptr, 0, &kp));
      cudaGraphEdgeData ed{};
      ed.from_port = cudaGraphKernelNodePortProgrammatic;
      ed.to_port = 0;
      ed.type = cudaGraphDependencyTypeProgrammatic;
      FISS_CUDA(h, cudaGraphAddDependencies_v2(G.graph, &prev, &n, &ed, 1));
    } else {
      FISS_CUDA(h, cudaGraphAddKernelNode(&n, G.graph, prev ? &prev : nullptr, prev ? 1 : 0, &kp));
    }
    G.knodes.push_back(n);
    prev = n;
  }
  if (d2h_bytes) {
    cudaGraphNode_t n;
    FISS_CUDA(h, cudaGraphAddMemcpyNode1D(&n, G.graph, prev ? &prev : nullptr, prev ? 1 : 0, d2h_dst, d2h_src, d2h_bytes,
                                          cudaMemcpyDeviceToHost));
  }
  return FISS_OK;
}

int32_t graph_run(fiss_handle* h, cudaStream_t st, StepGraph& G, std::vector<KernelLaunch>& kl, const void* h2d_src,
                  void* h2d_dst, size_t h2d_bytes, const void* d2h_src, void* d2h_dst, size_t d2h_bytes) {
  bool same = G.exec != nullptr && G.kl.size() == kl.size() && G.h2d_src == h2d_src && G.h2d_dst == h2d_dst &&
              G.h2d_bytes == h2d_bytes && G.d2h_src == d2h_src && G.d2h_dst == d2h_dst && G.d2h_bytes == d2h_bytes;
  for (size_t i = 0; same && i < kl.size(); ++i) same = G.kl[i].func == kl[i].func && G.kl[i].pdl == kl[i].pdl;
  if (!same) {
    G.release();
    static const bool by_capture = std::getenv("FISS_GRAPH_CAPTURE") != nullptr;
    int32_t rc = (by_capture && !h2d_bytes && !d2h_bytes)
                     ? graph_build_by_capture(h, G, kl)
                     : graph_build_explicit(h, G, kl, h2d_src, h2d_dst, h2d_bytes, d2h_src, d2h_dst, d2h_bytes);
    if (rc != FISS_OK) {
      G.release();
      return rc;
    }
    FISS_CUDA(h, cudaGraphInstantiate(&G.exec, G.graph, 0));
    G.kl = kl;
    G.h2d_src = h2d_src; G.h2d_dst = h2d_dst; G.h2d_bytes = h2d_bytes;
    G.d2h_src = d2h_src; G.d2h_dst = d2h_dst; G.d2h_bytes = d2h_bytes;
    G.rebuilds++;
  } else {
    for (size_t i = 0; i < kl.size(); ++i) {
      if (G.kl[i].same_shape(kl[i]) && G.kl[i].same_args(kl[i])) continue;
      void* params[8];
      kl[i].params(params);
      cudaKernelNodeParams kp{};
      kp.func = const_cast<void*>(kl[i].func);
      kp.gridDim = kl[i].grid;
      kp.blockDim = kl[i].block;
      kp.sharedMemBytes = (unsigned)kl[i].smem;
      kp.kernelParams = params;
      FISS_CUDA(h, cudaGraphExecKernelNodeSetParams(G.exec, G.knodes[i], &kp));
      G.kl[i] = kl[i];
      G.patches++;
    }
  }
  FISS_CUDA(h, cudaGraphLaunch(G.exec, st));
  G.launches++;
  h->launches += (int64_t)kl.size();
  return FISS_OK;
}

int32_t check_end_states(fiss_handle* h, const double* end, int C, int n_stride) {
  for (int c = 0; c < C; ++c) {
    const double T = end[4 * (size_t)c + 2], n = end[4 * (size_t)c + 3];
    if (!(T > 0.0) || !(n >= 1.0) || n != std::floor(n))
      return fail(h, FISS_ERR_INVALID, "end state " + std::to_string(c) + ": need T > 0 and an integral n >= 1");
    if (n > n_stride) return fail(h, FISS_ERR_INVALID, "n_stride is smaller than a candidate's step count");
  }
  return FISS_OK;
}


// ---- product lattice ---------------------------------------------------------------------------
// Validate a fiss_grid, (re)upload its axes (and the expanded [C][4] end-state table the record kernel
// indexes) when they changed, and return the largest step count.
int32_t ensure_grid(fiss_handle* h, cudaStream_t st, const fiss_grid* g, const fiss_params* p, int* n_max_out) {
  if (!g || !g->d_end || !g->v_end || !g->T) return fail(h, FISS_ERR_INVALID, "grid or one of its axes is NULL");
  const int nd = g->nd, nv = g->nv, nt = g->nt;
  if (nd < 1 || nv < 1 || nt < 1 || nd > FISS_GRID_AXIS_MAX || nv > FISS_GRID_AXIS_MAX || nt > FISS_GRID_AXIS_MAX)
    return fail(h, FISS_ERR_CAPACITY, "grid axes must have 1.." + std::to_string(FISS_GRID_AXIS_MAX) + " points");
  // the strides must number [0, nd*nv*nt) densely
  struct Dim { int size, stride; } dims[3] = {{nd, g->stride_d}, {nv, g->stride_v}, {nt, g->stride_t}};
  std::sort(dims, dims + 3, [](const Dim& x, const Dim& y) { return x.stride < y.stride || (x.stride == y.stride && x.size < y.size); });
  int64_t expect = 1;
  for (const Dim& d : dims) {
    if (d.size > 1 && d.stride != expect) return fail(h, FISS_ERR_INVALID, "grid strides are not a dense numbering");
    if (d.size > 1) expect *= d.size;
  }
  constexpr int A = fiss::kAxisMax;
  std::vector<double> axes((size_t)4 * A + 6, 0.0);
  for (int i = 0; i < nd; ++i) axes[i] = g->d_end[i];
  for (int j = 0; j < nv; ++j) axes[A + j] = g->v_end[j];
  int n_max = 0;
  for (int k = 0; k < nt; ++k) {
    const double T = g->T[k];
    const int n = fiss_arange_len(T, p->tick_t);
    if (!(T > 0.0) || n < 1) return fail(h, FISS_ERR_INVALID, "grid horizon " + std::to_string(k) + ": need T > 0 and n >= 1");
    axes[2 * A + k] = T;
    axes[3 * A + k] = (double)n;
    n_max = std::max(n_max, n);
  }
  // the cache key also carries the shape and strides
  double* tail = axes.data() + 4 * A;
  tail[0] = nd; tail[1] = nv; tail[2] = nt; tail[3] = g->stride_d; tail[4] = g->stride_v; tail[5] = g->stride_t;
  *n_max_out = n_max;
  if (axes == h->axes_cache) return FISS_OK;
  const int C = nd * nv * nt;
  std::vector<double> end((size_t)C * 4);
  for (int i = 0; i < nd; ++i)
    for (int j = 0; j < nv; ++j)
      for (int k = 0; k < nt; ++k) {
        double* e = end.data() + 4 * ((size_t)i * g->stride_d + (size_t)j * g->stride_v + (size_t)k * g->stride_t);
        e[0] = axes[i]; e[1] = axes[A + j]; e[2] = axes[2 * A + k]; e[3] = axes[3 * A + k];
      }
  FISS_CUDA(h, h->d_axes.ensure((size_t)4 * A * 8));
  FISS_CUDA(h, h->d_end.ensure(end.size() * 8));
  // pageable sources: both copies are complete w.r.t. the host buffers on return, ordered on `st`
  FISS_CUDA(h, cudaMemcpyAsync(h->d_axes.p, axes.data(), (size_t)4 * A * 8, cudaMemcpyHostToDevice, st));
  FISS_CUDA(h, cudaMemcpyAsync(h->d_end.p, end.data(), end.size() * 8, cudaMemcpyHostToDevice, st));
  h->axes_cache = axes;
  h->end_cache = end;
  h->grid_n_max = n_max;
  return FISS_OK;
}

template <bool kYaw>
int32_t launch_grid(fiss_handle* h, cudaStream_t st, const fiss::GridArgs& a, size_t smem, int threads, int which) {
  auto kern = fiss::fiss_grid_kernel<kYaw>;
  if (smem > h->smem_attr[which]) {
    FISS_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
    h->smem_attr[which] = kSmemLimit;
  }
  int occ = 1;
  FISS_CUDA(h, h->occupancy(which, kern, threads, smem, &occ));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(a.items, (int64_t)h->sm_count * occ));
  KernelLaunch kl;
  kl.func = (const void*)kern;
  kl.grid = dim3((unsigned)grid);
  kl.block = dim3((unsigned)threads);
  kl.smem = smem;
  kl.pdl = a.chained != 0;
  fiss::GridArgs b = a;
  if (h->record == nullptr) {  // nothing can fail between here and the launch but the launch itself
    b.seq_prev = h->lattice_seq;
    if (++h->lattice_seq == 0) ++h->lattice_seq;  // (0 is "not numbered")
    b.seq = h->lattice_seq;
  }
  kl.push(b);
  const int32_t rc = issue(h, st, kl);
  // a launch that was never issued publishes nothing: its number must not become anybody's predecessor
  if (rc != FISS_OK && b.seq != 0) h->lattice_seq = b.seq_prev;
  return rc;
}

int32_t eval_grid(fiss_handle* h, cudaStream_t st, const double* d_ego, int B, const fiss_grid* g, int n_max,
                  const fiss_params* p, double* d_cost, uint32_t* d_flags, double* d_mat, int n_stride) {
  if (h->K < 2) return fail(h, FISS_ERR_STATE, "fiss_set_spline has not been called");
  if (n_stride < n_max) return fail(h, FISS_ERR_INVALID, "n_stride is smaller than a candidate's step count");
  fiss::GridArgs a{};
  a.ego = d_ego;
  a.axes = h->d_axes.as<double>();
  a.nd = g->nd; a.nv = g->nv; a.nt = g->nt;
  a.sd = g->stride_d; a.sv = g->stride_v; a.st = g->stride_t;
  a.B = B;
  a.C = g->nd * g->nv * g->nt;
  a.total = (int64_t)B * a.C;
  a.p = *p;
  a.spline = h->spline.as<double>();
  a.K = h->K;
  a.Kp = h->Kp;
  int it = 0;
  while ((1 << it) < std::max(h->K - 1, 1)) ++it;
  static const bool no_lut = std::getenv("FISS_NO_KNOT_INDEX") != nullptr;  // A/B switch
  a.lut_cells = no_lut ? 0 : h->lut_cells;
  a.lut_bytes = no_lut ? 0 : h->lut_bytes;
  a.lut_inv_h = h->lut_inv_h;
  a.search_iters = a.lut_cells > 0 ? h->lut_iters : it;
  a.obs_tab = h->obs_tab.as<double>();
  a.obs_const = h->obs_const.as<double>();
  a.M = h->M; a.Mp = h->Mp; a.mp_shift = h->mp_shift; a.T_obs = h->T_obs; a.final_time_step = h->final_time_step;
  a.words = std::max(1, h->Mp / 32);
  a.n_pad = (n_max + 2) | 1;  // >= n_max + 2 (two NaN pad elements behind a full-length row), odd (bank spread of the row tables)
  a.e_pad = (n_max + p->check_res - 1) / p->check_res;
  a.cost = d_cost; a.flags = d_flags; a.mat = d_mat; a.n_stride = n_stride;
  a.mat_pitch = a.total * n_stride;
  a.kap_limit = p->check_curvature ? p->max_curvature : HUGE_VAL;
  // work items: (ego state, horizon[, chunk of lateral rows]).  With few ego states the lateral axis is
  // split so that the launch still covers the SMs (the longitudinal rows are recomputed per chunk).
  const int64_t base_items = (int64_t)B * g->nt;
  const int64_t want = 2 * (int64_t)h->sm_count;
  int chunks = (int)std::min<int64_t>(g->nd, std::max<int64_t>(1, (want + base_items - 1) / base_items));
  a.d_chunk = std::max(1, (g->nd + chunks - 1) / chunks);  // even chunks (33 rows in 2 chunks: 17 + 16, not 16 + 16 + 1)
  a.n_chunks = (g->nd + a.d_chunk - 1) / a.d_chunk;
  // obstacle rows (centres) of the checked steps go to shared memory while the CTA stays under the budget, else
  // they are read from global memory / L2
  const int horizon = std::max(0, std::min(n_max, h->final_time_step - p->time_step_now));
  const int E_max = h->M > 0 ? (horizon + p->check_res - 1) / p->check_res : 0;
  const auto magic = [](int d) { return (1u << 20) / (uint32_t)std::max(d, 1) + 1u; };
  a.nv_magic = magic(a.nv);
  a.nt_magic = a.nt > 1 ? 0xffffffffu / (uint32_t)a.nt + 1u : 0u;
  a.ns_magic = magic(n_stride);
  const int n_groups = (a.d_chunk + fiss::kMatRows - 1) / fiss::kMatRows;
  a.ng_magic = magic(n_groups);
  // (ego, horizon) pairs per work item: as many as keep the variant's CTAs per SM resident (1 KB of shared memory per
  // CTA is reserved) and every CTA supplied with a few items (FISS_GRID_SLOTS overrides the upper bound, for A/B runs)
  const bool yaw = d_mat != nullptr || p->check_curvature;
  const int min_ctas = fiss::grid_min_ctas(yaw);
  const size_t kSmemCtaBudget = (228 * 1024) / min_ctas - 1024;
  static const bool slots_forced = std::getenv("FISS_GRID_SLOTS") != nullptr;  // also waives the supply rule (tests)
  static const int slots_env = slots_forced ? std::max(1, std::min(fiss::kMaxSlots, std::atoi(std::getenv("FISS_GRID_SLOTS")))) : 0;
  const int slots_cap = slots_forced ? slots_env : fiss::grid_slots(yaw);
  fiss::GridLayout L{};
  for (a.slots = a.n_chunks > 1 ? 1 : slots_cap;; --a.slots) {
    // the kernel's multiply-shift divisions are exact below 2^20
    const int64_t lon_elems = (int64_t)a.slots * a.nv * n_stride + 32, n_blocks = (lon_elems + 30) / 31;
    const bool exact = lon_elems * n_stride < (1 << 20) && (int64_t)a.slots * a.d_chunk * a.nv * a.nv < (1 << 20) &&
                       n_blocks * n_groups * n_groups < (1 << 20) && (int64_t)a.slots * a.C * n_stride < ((int64_t)1 << 31) &&
                       base_items * g->nt < ((int64_t)1 << 32);
    a.E_stage = E_max;
    L = fiss::grid_layout(a.Kp, a.Mp, a.E_stage, a.nv, a.d_chunk, a.n_pad, a.e_pad, a.words, a.slots, a.lut_bytes, yaw);
    // ... and the items must deal evenly over the resident CTAs: the launch lasts ceil(items / CTAs) item times
    // (640 four-slot items on 296 CTAs: 3 rounds, 72 % busy; 854 three-slot items: 3 rounds, 96 %)
    const int64_t items_s = (base_items + a.slots - 1) / a.slots, ctas = (int64_t)min_ctas * h->sm_count;
    const int64_t rounds = (items_s + ctas - 1) / ctas;
    const bool supplied = slots_forced || (items_s >= 2 * ctas && items_s * 10 >= rounds * ctas * 9);
    if (a.slots > 1) {
      if (exact && supplied && L.bytes <= kSmemCtaBudget) break;
      continue;
    }
    if (!exact) return fail(h, FISS_ERR_CAPACITY, "lattice or batch too large for the kernel's index arithmetic (nv * n_stride^2 < 2^20, B * nt^2 < 2^32)");
    if (L.bytes > kSmemCtaBudget) {
      a.E_stage = 0;
      L = fiss::grid_layout(a.Kp, a.Mp, a.E_stage, a.nv, a.d_chunk, a.n_pad, a.e_pad, a.words, a.slots, a.lut_bytes, yaw);
    }
    break;
  }
  // Chained launches (direct launches of batches that fill the GPU; FISS_CHAIN=0 switches it off): the lattice kernel is
  // launched as a programmatic dependent of the previous kernel of the stream, so that its CTAs fill the SMs the previous
  // step's last items leave idle (GridArgs::chained says what it waits for before it writes).  Two launches can then be
  // drawing work items at the same time: the counters alternate between two pairs.
  static const bool chain_on = std::getenv("FISS_CHAIN") ? std::atoi(std::getenv("FISS_CHAIN")) != 0 : true;
  a.chained = chain_on && h->record == nullptr && a.n_chunks == 1;
  // Work items: dealt with a grid stride, or handed out through a device counter in index order -- then FISS_BIG_FRAC percent
  // of the pairs travel in items of `slots` pairs and the rest one pair per item, so that the CTAs run out of work together
  // (tools/warp_trace.py: the time of an item follows its collision stage, which varies ~5x with the traffic around the ego
  // state).  Measured on the B200 (cfg4): the winner-only kernel gains 4-5 % (e2e +4-8 %) at 70 %.  The materialising
  // kernel on its own loses 2-5 % (single pairs cost it more than the even finish wins) and keeps the static deal -- but
  // in a CHAINED launch its CTAs start at different times (as the previous step's CTAs retire) and must draw their work:
  // with the static deal a late starter still has three items to do and the step loop falls back to one step after the
  // other.  A/B switches: FISS_DYN / FISS_DYN_MAT = 0 | 1 for the winner-only / materialising kernel, FISS_BIG_FRAC = percent.
  static const bool big_forced = std::getenv("FISS_BIG_FRAC") != nullptr;
  static const int big_env = big_forced ? std::max(0, std::min(100, std::atoi(std::getenv("FISS_BIG_FRAC")))) : 70;
  static const bool dyn_mat_forced = std::getenv("FISS_DYN_MAT") != nullptr;
  static const bool dyn_mat = dyn_mat_forced ? std::atoi(std::getenv("FISS_DYN_MAT")) != 0 : false;
  static const bool dyn_win = std::getenv("FISS_DYN") ? std::atoi(std::getenv("FISS_DYN")) != 0 : true;
  const bool can_deal = a.n_chunks == 1 && a.slots > 1;
  // Is this launch part of a TRAIN -- is an earlier numbered launch of this handle still running, so that this one will start
  // on the SMs it leaves behind?  Its last CTA writes its number to a page-locked host word when it is over.  Only then do the
  // throughput settings pay (work drawn dynamically, whole items, a small record grid behind it); a launch onto an idle GPU --
  // one step, then a synchronisation -- keeps the settings that make a single step fastest (0.128 vs 0.144 ms).
  const bool in_train = a.chained != 0 && (int32_t)(h->lattice_seq - *static_cast<volatile uint32_t*>(h->host_done.p)) > 0;
  a.dynamic = can_deal && (yaw ? (dyn_mat_forced ? dyn_mat : in_train) : dyn_win);
  // (in a train: whole items only -- the CTAs of the next launch fill in behind the last items, and single pairs cost more
  // than they even out: winner-only step train 0.0626 -> 0.0576 ms)
  const int big_pct = big_forced ? big_env : ((yaw || in_train) ? 100 : 70);
  const int64_t full_items = (base_items + a.slots - 1) / a.slots;
  if (a.n_chunks == 1) {
    const bool split = a.dynamic && big_pct < 100;
    a.n_big = split ? (int32_t)(base_items * big_pct / 100 / a.slots) : (int32_t)full_items;
    a.items = split ? a.n_big + (base_items - (int64_t)a.n_big * a.slots) : a.n_big;
  } else {
    a.n_big = 0;
    a.items = base_items * a.n_chunks;
  }
  // Directly issued launches are numbered (the number is taken when the launch is issued: launch_grid) and publish their
  // number when all their CTAs are done; a launch recorded for a graph carries no number, publishes nothing and draws from
  // a counter pair of its own -- its arguments then stay the same from call to call (no node to patch), and it never runs
  // beside another launch (a graph launch and the kernels around it serialise).
  a.seq = 0;
  h->lattice_chained = in_train ? (yaw ? 2 : 1) : 0;
  a.shadow = nullptr;
  a.shadow_stride = (a.slots * a.d_chunk * a.nv + 1) & ~1;
  if (a.chained) {  // shadow blocks for the first items' cost / flags: one per resident CTA, two sets
    const size_t set_bytes = (size_t)h->sm_count * 8 * a.shadow_stride * 20;  // (<= 8 resident CTAs per SM)
    FISS_CUDA(h, h->d_shadow.ensure(2 * set_bytes));
    a.shadow = h->d_shadow.as<unsigned char>() + (size_t)((h->lattice_seq + 1) & 1u) * set_bytes;
  }
  if (h->record == nullptr) h->work_parity ^= 1;
  a.work = h->d_work.as<uint32_t>() + (h->record == nullptr ? 2 * h->work_parity : 6);  // (allocated and zeroed by fiss_create)
  a.work_done = h->d_work.as<uint32_t>() + 4;
  a.work_done_host = h->host_done.as<uint32_t>();  // (page-locked memory is mapped into the device's address space: UVA)
  const int warps = std::max(1, std::min(fiss::grid_warps(yaw), a.slots * std::max(g->nv + a.d_chunk, a.d_chunk * g->nv)));
  if (L.bytes > kSmemLimit)
    return fail(h, FISS_ERR_CAPACITY,
                "lattice kernel needs " + std::to_string(L.bytes) + " B of shared memory (limit 232448): spline table " +
                    std::to_string((size_t)9 * a.Kp * 8 + a.lut_bytes) + " B for K = " + std::to_string(a.K) +
                    " knots (the whole table is staged per CTA: K <= ~3000), row tables " +
                    std::to_string((size_t)L.lon_cost - L.lon) + " B");
  a.lay = L;
  a.row_len = a.slots * a.nv * a.n_pad;
  return yaw ? launch_grid<true>(h, st, a, L.bytes, warps * 32, 4) : launch_grid<false>(h, st, a, L.bytes, warps * 32, 3);
}

}  // namespace

// ================================================================================================
extern "C" {

int32_t fiss_abi_version(void) { return 2; }

#ifdef FISS_PHASE_TIMING
// debug builds only (tools/phase_timing.py): read and reset the per-stage cycle counters of the lattice kernel
int32_t fiss_debug_phase_cycles(long long* out) {
  long long zero[16] = {0};
  if (cudaMemcpyFromSymbol(out, g_fiss_phase, 8 * sizeof(long long)) != cudaSuccess) return FISS_ERR_CUDA;
  if (cudaMemcpyToSymbol(g_fiss_phase, zero, sizeof(zero)) != cudaSuccess) return FISS_ERR_CUDA;
  return FISS_OK;
}
#endif

#ifdef FISS_TRACE
// debug builds only (tools/warp_trace.py): copy out and clear the per-warp stage stamps of the lattice kernel
int32_t fiss_debug_trace(long long* out, int64_t n) {
  const size_t bytes = sizeof(long long) * (size_t)std::min<int64_t>(n, (int64_t)(sizeof(g_fiss_trace) / sizeof(long long)));
  if (cudaMemcpyFromSymbol(out, g_fiss_trace, bytes) != cudaSuccess) return FISS_ERR_CUDA;
  void* sym = nullptr;
  if (cudaGetSymbolAddress(&sym, g_fiss_trace) != cudaSuccess) return FISS_ERR_CUDA;
  if (cudaMemset(sym, 0, sizeof(g_fiss_trace)) != cudaSuccess) return FISS_ERR_CUDA;
  return FISS_OK;
}
#endif

int32_t fiss_arange_len(double T, double tick) {
  if (!(tick > 0.0) || !(T > 0.0)) return 0;
  const double len = std::ceil((T - 0.0) / tick);  // NumPy: ceil((stop - start) / step) in double
  if (!(len < 2147483647.0)) return -1;
  return (int32_t)len;
}

int32_t fiss_create(int32_t device, fiss_handle** out) {
  if (!out) return fail(nullptr, FISS_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return fail(nullptr, FISS_ERR_CUDA,
                std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (device < 0 || device >= count) return fail(nullptr, FISS_ERR_INVALID, "device index out of range");
  DeviceGuard device_guard_(device);  // the caller's current device is restored on return
  e = device_guard_.err;
  if (e != cudaSuccess) return fail(nullptr, FISS_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, FISS_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
  if (prop.major < 10)
    return fail(nullptr, FISS_ERR_CUDA, "libfissgpu is built for sm_100a (Blackwell) only; found sm_" +
                                            std::to_string(prop.major) + std::to_string(prop.minor));
  fiss_handle* h = new (std::nothrow) fiss_handle();
  if (!h) return fail(nullptr, FISS_ERR_CUDA, "out of host memory");
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  // work counters of the lattice kernel: zero between launches (the last CTA of a launch clears them)
  e = h->d_work.ensure(256);
  if (e == cudaSuccess) e = cudaMemset(h->d_work.p, 0, 256);
  if (e == cudaSuccess) e = h->host_done.ensure(64);
  if (e == cudaSuccess) *h->host_done.as<uint32_t>() = 0u;
  if (e != cudaSuccess) {
    delete h;
    return fail(nullptr, FISS_ERR_CUDA, std::string("work counters: ") + cudaGetErrorString(e));
  }
  *out = h;
  return FISS_OK;
}

int32_t fiss_destroy(fiss_handle* h) {
  if (!h) return FISS_OK;
  DeviceGuard device_guard_(h->device);
  for (DevBuf* b : {&h->spline, &h->obs_tab, &h->obs_const, &h->obs_raw, &h->obs_lw, &h->obs_valid, &h->d_ego,
                    &h->d_end, &h->d_cost, &h->d_flags, &h->d_best_idx, &h->d_best_cost, &h->d_meta, &h->d_records,
                    &h->d_es, &h->d_axes, &h->d_work, &h->d_shadow, &h->d_fit_in, &h->d_fit_out, &h->d_arena})
    b->release();
  h->h_in.release();
  h->h_out.release();
  h->host_done.release();
  h->graph_dev.release();
  h->graph_host.release();
  h->d_pick.release();
  for (auto& ln : h->lanes) ln.release();
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(static_cast<ncclComm_t>(h->comm));
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
  for (auto& ev : h->part_done)
    if (ev) cudaEventDestroy(ev);
  delete h;
  return FISS_OK;
}

const char* fiss_last_error(const fiss_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int64_t fiss_launch_count(const fiss_handle* h) { return h ? h->launches : 0; }

// ------------------------------------------------------------------------------------------------
int32_t fiss_set_spline(fiss_handle* h, void* stream, const double* table, int32_t K) {
  if (!h) return FISS_ERR_INVALID;
  if (!table || K < 2) return fail(h, FISS_ERR_INVALID, "spline table needs K >= 2 knots");
  cudaStream_t st = (cudaStream_t)stream;
  FISS_ON_DEVICE(h);
  const int Kp = (K + 1) & ~1;  // even row length: every row is a multiple of 16 bytes for the bulk copy
  for (int k = 1; k < K; ++k)
    if (!(table[k] >= table[k - 1])) return fail(h, FISS_ERR_INVALID, "spline knots must be ascending");
  // knot index for the lattice kernel's segment search: kLutCells uniform cells over [knots[0], knots[K-1]),
  // lut[c] = the largest i with knots[i] <= start of cell c; the kernel brackets an abscissa by the cells around its own
  // (lut[c-1], lut[c+2] + 1) and bisects `lut_iters` times -- the bound over all cells computed here
  constexpr int kLutCells = 256;
  const double s0 = table[0], s1 = table[K - 1];
  int cells = 0, iters = 0;
  double inv_h = 0.0;
  std::vector<int32_t> lut;
  if (std::isfinite(s0) && std::isfinite(s1) && s1 > s0) {
    cells = kLutCells;
    inv_h = cells / (s1 - s0);
    lut.assign(cells + 1, 0);
    int i = 0;
    for (int c = 0; c <= cells; ++c) {
      const double start = s0 + c * ((s1 - s0) / cells);
      while (i + 1 < K && table[i + 1] <= start) ++i;
      lut[c] = i;
    }
    int span = 1;
    for (int c = 0; c < cells; ++c) {
      const int lo = lut[std::max(c - 1, 0)], hi = std::min(lut[std::min(c + 2, cells)] + 1, K - 1);
      span = std::max(span, hi - lo);
    }
    while ((1 << iters) < span) ++iters;
    if (!std::isfinite(inv_h)) cells = 0;
  }
  const int lut_bytes = cells > 0 ? (int)((((size_t)cells + 1) * 4 + 15) & ~(size_t)15) : 0;
  const size_t tab_bytes = (size_t)9 * Kp * 8;
  FISS_CUDA(h, h->h_in.ensure(tab_bytes + lut_bytes));
  double* stage = h->h_in.as<double>();
  for (int r = 0; r < 9; ++r) {
    std::memcpy(stage + (size_t)r * Kp, table + (size_t)r * K, (size_t)K * 8);
    for (int k = K; k < Kp; ++k) stage[(size_t)r * Kp + k] = r == 0 ? INFINITY : 0.0;
  }
  if (lut_bytes) {
    std::memset(reinterpret_cast<char*>(stage) + tab_bytes, 0, lut_bytes);
    std::memcpy(reinterpret_cast<char*>(stage) + tab_bytes, lut.data(), lut.size() * 4);
  }
  FISS_CUDA(h, h->spline.ensure(tab_bytes + lut_bytes));
  FISS_CUDA(h, cudaMemcpyAsync(h->spline.p, stage, tab_bytes + lut_bytes, cudaMemcpyHostToDevice, st));
  FISS_CUDA(h, cudaStreamSynchronize(st));
  h->lut_cells = cells;
  h->lut_bytes = lut_bytes;
  h->lut_iters = iters;
  h->lut_inv_h = inv_h;
  h->K = K;
  h->Kp = Kp;
  return FISS_OK;
}

int32_t fiss_fit_splines_host(fiss_handle* h, void* stream, const double* xy, int32_t L, int32_t K, double* tables,
                              int32_t install_lane) {
  if (!h) return FISS_ERR_INVALID;
  if (!xy || L < 1 || K < 2) return fail(h, FISS_ERR_INVALID, "fit_splines: need L >= 1 lanes of K >= 2 way points");
  if (install_lane >= L) return fail(h, FISS_ERR_INVALID, "fit_splines: install_lane out of range");
  cudaStream_t st = (cudaStream_t)stream;
  FISS_ON_DEVICE(h);
  const int Kp = (K + 1) & ~1;
  const size_t smem = (size_t)6 * Kp * 8;
  if (smem > kSmemLimit) return fail(h, FISS_ERR_CAPACITY, "fit_splines: more way points than 227 KB of shared memory hold");
  const size_t in_bytes = (size_t)L * K * 2 * 8, tab_doubles = (size_t)L * 9 * Kp;
  FISS_CUDA(h, h->h_in.ensure(in_bytes));
  std::memcpy(h->h_in.p, xy, in_bytes);
  FISS_CUDA(h, h->d_fit_in.ensure(in_bytes));
  FISS_CUDA(h, h->d_fit_out.ensure(tab_doubles * 8));
  FISS_CUDA(h, cudaMemcpyAsync(h->d_fit_in.p, h->h_in.p, in_bytes, cudaMemcpyHostToDevice, st));
  if (smem > 48 * 1024 && smem > h->smem_attr[6]) {
    FISS_CUDA(h, cudaFuncSetAttribute(fiss::fiss_spline_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
    h->smem_attr[6] = kSmemLimit;
  }
  fiss::fiss_spline_fit_kernel<<<L, fiss::kFitThreads, smem, st>>>(h->d_fit_in.as<double>(), K, Kp, h->d_fit_out.as<double>());
  h->launches++;
  FISS_CUDA(h, cudaGetLastError());
  if (install_lane >= 0) {
    FISS_CUDA(h, h->spline.ensure((size_t)9 * Kp * 8));
    FISS_CUDA(h, cudaMemcpyAsync(h->spline.p, h->d_fit_out.as<double>() + (size_t)install_lane * 9 * Kp, (size_t)9 * Kp * 8,
                                 cudaMemcpyDeviceToDevice, st));
    h->K = K;
    h->Kp = Kp;
    h->lut_cells = h->lut_bytes = h->lut_iters = 0;  // no knot index for a table fitted on the device: full bisection
    h->lut_inv_h = 0.0;
  }
  if (tables) {
    FISS_CUDA(h, h->h_out.ensure(tab_doubles * 8));
    FISS_CUDA(h, cudaMemcpyAsync(h->h_out.p, h->d_fit_out.p, tab_doubles * 8, cudaMemcpyDeviceToHost, st));
  }
  FISS_CUDA(h, cudaStreamSynchronize(st));
  if (tables) {
    const double* src = h->h_out.as<double>();
    for (int64_t r = 0; r < (int64_t)L * 9; ++r) std::memcpy(tables + r * K, src + r * Kp, (size_t)K * 8);
  }
  return FISS_OK;
}

int32_t fiss_frame_samples_host(fiss_handle* h, void* stream, double step, int32_t m, double* ref) {
  if (!h) return FISS_ERR_INVALID;
  if (!ref || m < 1 || !(step > 0.0)) return fail(h, FISS_ERR_INVALID, "frame_samples: bad arguments");
  if (h->K < 2) return fail(h, FISS_ERR_STATE, "no reference line: call fiss_set_spline / fiss_fit_splines_host first");
  cudaStream_t st = (cudaStream_t)stream;
  FISS_ON_DEVICE(h);
  const size_t bytes = (size_t)m * 4 * 8;
  FISS_CUDA(h, h->d_fit_out.ensure(bytes));
  FISS_CUDA(h, h->h_out.ensure(bytes));
  fiss::fiss_frame_samples_kernel<<<(m + 127) / 128, 128, 0, st>>>(h->spline.as<double>(), h->K, h->Kp, step, m,
                                                                  h->d_fit_out.as<double>());
  h->launches++;
  FISS_CUDA(h, cudaGetLastError());
  FISS_CUDA(h, cudaMemcpyAsync(h->h_out.p, h->d_fit_out.p, bytes, cudaMemcpyDeviceToHost, st));
  FISS_CUDA(h, cudaStreamSynchronize(st));
  std::memcpy(ref, h->h_out.p, bytes);
  return FISS_OK;
}

static int32_t obstacle_dims(fiss_handle* h, int32_t M, int32_t T_obs, int32_t final_time_step) {
  if (M < 0 || T_obs < 0) return fail(h, FISS_ERR_INVALID, "negative obstacle table size");
  h->M = M;
  h->T_obs = M > 0 ? T_obs : 0;
  h->final_time_step = final_time_step;
  if (M == 0) {
    h->Mp = 0;
    h->mp_shift = 0;
    return FISS_OK;
  }
  if (M <= 32) {
    int s = 0;
    while ((1 << s) < M) ++s;
    h->Mp = 1 << s;
    h->mp_shift = s;
  } else {
    h->Mp = (M + 31) & ~31;
    h->mp_shift = 5;
  }
  return FISS_OK;
}

int32_t fiss_set_obstacles(fiss_handle* h, void* stream, const double* xyth, const double* lw, const uint8_t* valid,
                           int32_t M, int32_t T_obs, int32_t final_time_step) {
  if (!h) return FISS_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  FISS_ON_DEVICE(h);
  if (M > 0 && (!xyth || !lw || !valid || T_obs < 1))
    return fail(h, FISS_ERR_INVALID, "obstacle arrays are NULL or T_obs < 1");
  int32_t rc = obstacle_dims(h, M, T_obs, final_time_step);
  if (rc != FISS_OK || M == 0) return rc;
  const size_t n_state = (size_t)M * T_obs;
  FISS_CUDA(h, h->obs_raw.ensure(n_state * 3 * 8));
  FISS_CUDA(h, h->obs_lw.ensure((size_t)M * 2 * 8));
  FISS_CUDA(h, h->obs_valid.ensure(n_state));
  FISS_CUDA(h, h->obs_tab.ensure((size_t)T_obs * h->Mp * 32));
  FISS_CUDA(h, h->obs_const.ensure((size_t)h->Mp * 32));
  FISS_CUDA(h, cudaMemcpyAsync(h->obs_raw.p, xyth, n_state * 3 * 8, cudaMemcpyHostToDevice, st));
  FISS_CUDA(h, cudaMemcpyAsync(h->obs_lw.p, lw, (size_t)M * 2 * 8, cudaMemcpyHostToDevice, st));
  FISS_CUDA(h, cudaMemcpyAsync(h->obs_valid.p, valid, n_state, cudaMemcpyHostToDevice, st));
  const int64_t work = std::max<int64_t>((int64_t)T_obs * h->Mp, h->Mp);
  fiss::fiss_obstacle_prep_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(
      h->obs_raw.as<double>(), h->obs_lw.as<double>(), h->obs_valid.as<uint8_t>(), M, h->Mp, T_obs,
      h->obs_tab.as<double>(), h->obs_const.as<double>());
  h->launches++;
  FISS_CUDA(h, cudaGetLastError());
  FISS_CUDA(h, cudaStreamSynchronize(st));  // the host arrays may be released by the caller on return
  return FISS_OK;
}

int32_t fiss_set_obstacles_waymo(fiss_handle* h, void* stream, const float* trajs, const uint8_t* mask, int32_t N,
                                 int32_t T, int32_t final_time_step, int32_t* n_kept) {
  if (!h) return FISS_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  FISS_ON_DEVICE(h);
  if (N < 0 || T < 0) return fail(h, FISS_ERR_INVALID, "negative waymo array size");
  if (N > 0 && (!trajs || !mask || T < 1)) return fail(h, FISS_ERR_INVALID, "waymo arrays are NULL or T < 1");
  // convert_waymo_obstacle_to_cr (waymo_interface.py:24-76): the trajectory of agent i is steps 1.. up to the first masked
  // step (:45-54); an agent whose trajectory comes out empty is NOT appended (:58) -- it has no state at any step, the
  // initial one included, and the agents behind it move up (obstacles[0], hence final_time_step, is the first KEPT agent)
  std::vector<int32_t> keep_end;  // [kept agent ids | last step with a state]
  std::vector<int32_t> t_end;
  for (int i = 0; i < N; ++i) {
    const uint8_t* mi = mask + (size_t)i * T;
    if (T < 2 || !mi[1]) continue;
    int te = 1;
    while (te + 1 < T && mi[te + 1]) ++te;
    keep_end.push_back(i);
    t_end.push_back(te);
  }
  const int32_t M = (int32_t)keep_end.size();
  if (n_kept) *n_kept = M;
  // obstacles[0].prediction.final_time_step = the last state of the first kept agent's trajectory
  // (frenet_optimal_planner.py:173); an explicit value >= 0 overrides it
  const int32_t fts = final_time_step >= 0 ? final_time_step : (M > 0 ? t_end[0] : 0);
  int32_t rc = obstacle_dims(h, M, T, fts);
  if (rc != FISS_OK || M == 0) return rc;
  keep_end.insert(keep_end.end(), t_end.begin(), t_end.end());
  const size_t n_state = (size_t)N * T;
  FISS_CUDA(h, h->obs_raw.ensure(n_state * 11 * 4));
  FISS_CUDA(h, h->obs_valid.ensure((size_t)2 * M * 4));
  FISS_CUDA(h, h->obs_tab.ensure((size_t)T * h->Mp * 32));
  FISS_CUDA(h, h->obs_const.ensure((size_t)h->Mp * 32));
  FISS_CUDA(h, cudaMemcpyAsync(h->obs_raw.p, trajs, n_state * 11 * 4, cudaMemcpyHostToDevice, st));
  FISS_CUDA(h, cudaMemcpyAsync(h->obs_valid.p, keep_end.data(), (size_t)2 * M * 4, cudaMemcpyHostToDevice, st));
  const int64_t work = std::max<int64_t>((int64_t)T * h->Mp, h->Mp);
  fiss::fiss_obstacle_prep_waymo_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(
      h->obs_raw.as<float>(), h->obs_valid.as<int32_t>(), h->obs_valid.as<int32_t>() + M, M, h->Mp, T,
      h->obs_tab.as<double>(), h->obs_const.as<double>());
  h->launches++;
  FISS_CUDA(h, cudaGetLastError());
  FISS_CUDA(h, cudaStreamSynchronize(st));  // keep_end (and the caller's arrays) may go away on return
  return FISS_OK;
}

// ------------------------------------------------------------------------------------------------
int32_t fiss_eval_candidates_dev(fiss_handle* h, void* stream, const double* d_ego, int32_t B, const double* d_end,
                                 int32_t C, const fiss_params* p, double* d_cost, uint32_t* d_flags, double* d_mat,
                                 int32_t n_stride) {
  if (!h) return FISS_ERR_INVALID;
  if (!d_ego || !d_end || B < 1 || C < 1) return fail(h, FISS_ERR_INVALID, "ego/end pointers or sizes invalid");
  int32_t rc = check_params(h, p);
  if (rc != FISS_OK) return rc;
  FISS_ON_DEVICE(h);
  LaunchPlan lp{};
  rc = plan_launch(h, p, (int64_t)B * C, n_stride, lp);
  if (rc != FISS_OK) return rc;
  lp.a.ego = d_ego;
  lp.a.end = d_end;
  lp.a.sel = nullptr;
  lp.a.B = B;
  lp.a.C = C;
  lp.a.per_problem = 0;
  lp.a.cost = d_cost;
  lp.a.flags = d_flags;
  lp.a.mat = d_mat;
  lp.a.records = nullptr;
  lp.a.n_stride = n_stride;
  if (d_mat || p->check_curvature) return launch_eval<true, false>(h, (cudaStream_t)stream, lp, 1);
  return launch_eval<false, false>(h, (cudaStream_t)stream, lp, 0);
}

int32_t fiss_full_records_dev(fiss_handle* h, void* stream, const double* d_ego6, const double* d_end,
                              const int32_t* d_sel, int32_t N, const fiss_params* p, double* d_records,
                              double* d_cost, uint32_t* d_flags, int32_t n_stride) {
  if (!h) return FISS_ERR_INVALID;
  if (!d_ego6 || !d_end || !d_records || N < 1) return fail(h, FISS_ERR_INVALID, "full_records: bad arguments");
  int32_t rc = check_params(h, p);
  if (rc != FISS_OK) return rc;
  FISS_ON_DEVICE(h);
  LaunchPlan lp{};
  rc = plan_launch(h, p, N, n_stride, lp);
  if (rc != FISS_OK) return rc;
  lp.a.ego = d_ego6;
  lp.a.end = d_end;
  lp.a.sel = d_sel;
  lp.a.B = 1;
  lp.a.C = N;
  lp.a.per_problem = 0;
  lp.a.cost = d_cost;
  lp.a.flags = d_flags;
  lp.a.mat = nullptr;
  lp.a.records = d_records;
  lp.a.n_stride = n_stride;
  return launch_eval<false, true>(h, (cudaStream_t)stream, lp, 2);
}

int32_t fiss_pick_winners_dev(fiss_handle* h, void* stream, const double* d_ego, int32_t B, const double* d_end,
                              int32_t C, const fiss_params* p, const double* d_cost, const uint32_t* d_flags,
                              int32_t* d_best_idx, double* d_best_cost, double* d_records, int32_t* d_best_meta,
                              int32_t n_stride) {
  if (!h) return FISS_ERR_INVALID;
  if (!d_ego || !d_end || !d_cost || !d_flags || !d_best_idx || !d_best_cost || B < 1 || C < 1)
    return fail(h, FISS_ERR_INVALID, "pick_winners: bad arguments");
  int32_t rc = check_params(h, p);
  if (rc != FISS_OK) return rc;
  FISS_ON_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_records) {
    KernelLaunch kl;
    kl.func = (const void*)fiss::fiss_pick_kernel;
    kl.grid = dim3((unsigned)B);
    kl.block = dim3((unsigned)fiss::kPickThreads);
    kl.push(d_cost);
    kl.push(d_flags);
    kl.push((int)C);
    kl.push(d_best_idx);
    kl.push(d_best_cost);
    kl.push(d_end);
    kl.push(d_best_meta);
    rc = issue(h, st, kl);
    if (rc != FISS_OK) return rc;
  } else {  // one launch: the record kernel's warp of problem b picks b's winner first
    LaunchPlan lp{};
    rc = plan_launch(h, p, B, n_stride, lp);
    if (rc != FISS_OK) return rc;
    lp.a.ego = d_ego;
    lp.a.end = d_end;
    // the winners' flags are already in d_flags: the record launch skips the collision stage and its obstacle rows
    lp.a.M = 0;
    if (lp.a.obs_in_smem) {
      lp.smem -= (size_t)lp.a.E_max * lp.a.Mp * 32;
      lp.a.obs_in_smem = 0;
    }
    lp.a.E_max = 0;
    lp.a.sel = nullptr;
    lp.a.pick_cost = d_cost;
    lp.a.pick_flags = d_flags;
    lp.a.pick_idx = d_best_idx;
    lp.a.pick_best = d_best_cost;
    lp.a.pick_meta = d_best_meta;
    lp.a.after_producer = 1;  // everything the kernel touches before its dependency wait (spline, obstacle constants)
                              // was uploaded synchronously by fiss_set_spline / fiss_set_obstacles
    lp.a.B = B;
    lp.a.C = C;
    lp.a.per_problem = 1;
    lp.a.cost = nullptr;
    lp.a.flags = nullptr;
    lp.a.mat = nullptr;
    lp.a.records = d_records;
    lp.a.n_stride = n_stride;
    lp.a.total = B;
    static const bool list_records = std::getenv("FISS_LIST_RECORDS") != nullptr;  // A/B: the generic one-warp-per-problem kernel
    if (list_records) return launch_eval<false, true>(h, st, lp, 2);
    // warps per problem: one pass over the steps (2 covers the 50-step horizons, 4 the 100-step ones)
    if (n_stride <= 32) return launch_record<1>(h, st, lp.a, 5);
    if (n_stride <= 64) return launch_record<2>(h, st, lp.a, 7);
    return launch_record<4>(h, st, lp.a, 8);
  }
  return FISS_OK;
}

// ------------------------------------------------------------------------------------------------
static int32_t upload_end_states(fiss_handle* h, cudaStream_t st, const double* end, int32_t C) {
  const size_t bytes = (size_t)C * 4 * 8;
  if (h->end_cache.size() == (size_t)C * 4 && std::memcmp(h->end_cache.data(), end, bytes) == 0) return FISS_OK;
  FISS_CUDA(h, h->d_end.ensure(bytes));
  h->end_cache.assign(end, end + (size_t)C * 4);
  h->axes_cache.clear();  // d_end is shared with the lattice path
  // the cache vector is pageable memory: the copy is synchronous w.r.t. the host buffer, ordered on `st`
  FISS_CUDA(h, cudaMemcpyAsync(h->d_end.p, h->end_cache.data(), bytes, cudaMemcpyHostToDevice, st));
  return FISS_OK;
}

// Shared tail of the two plan entry points: H2D of the ego states, the hot kernel (lattice kernel when
// `g` is given, generic list kernel otherwise; the [C][4] table is already in h->d_end), the pick, the
// winners' records and the D2H of everything the caller asked for.  Synchronous on return.
static int32_t plan_common(fiss_handle* h, cudaStream_t st, const double* ego, int32_t B, int32_t C,
                           const fiss_grid* g, int n_max, const fiss_params* p, int32_t* best_idx, double* best_cost,
                           int32_t* best_meta, double* records, int32_t n_stride, double* cost, uint32_t* flags) {
  int32_t rc;
  void* stream = (void*)st;
  const size_t total = (size_t)B * C;
  const size_t rec_doubles = records ? (size_t)B * FISS_REC_ROWS * n_stride : 0;
  FISS_CUDA(h, h->d_ego.ensure((size_t)B * 48));
  FISS_CUDA(h, h->d_cost.ensure(total * 8));
  FISS_CUDA(h, h->d_flags.ensure(total * 4));
  // winners: one device block [best_cost B x 8 | meta B x 8 | best_idx B x 4] so that they travel in ONE copy
  const size_t w_cost = 0, w_meta = (size_t)B * 8, w_idx = (size_t)B * 16, w_bytes = (size_t)B * 20;
  // Latency path (a few problems): everything the caller asked for is written into ONE device arena
  // [winners | records | cost | flags] and comes back in ONE copy -- a copy operation costs ~3.6 us of stream time
  // whatever its size, and a plan cycle of one ego state used to issue four of them.
  {
    const size_t a_rec = (w_bytes + 15) & ~(size_t)15, a_vol = a_rec + rec_doubles * 8, a_flags = a_vol + (cost ? total * 8 : 0),
                 a_end = a_flags + (flags ? total * 4 : 0);
    if (a_end <= kArenaBytesMax) {
      const size_t arena_cap = h->d_arena.cap;
      FISS_CUDA(h, h->d_arena.ensure(a_end));
      // alignment gaps between the blocks travel with the copy: defined bytes (compute-sanitizer initcheck)
      if (h->d_arena.cap != arena_cap) FISS_CUDA(h, cudaMemsetAsync(h->d_arena.p, 0, h->d_arena.cap, st));
      FISS_CUDA(h, h->h_out.ensure(a_end));
      FISS_CUDA(h, h->h_in.ensure((size_t)B * 48));
      char* da = h->d_arena.as<char>();
      char* ha = h->h_out.as<char>();
      const bool as_graph = graphs_enabled();
      // a few hundred bytes: staged through the handle's pinned block (asking the runtime whether the caller's pointer
      // is page-locked costs more than the copy)
      std::memcpy(h->h_in.p, ego, (size_t)B * 48);
      const double* ego_p = h->d_ego.as<double>();
      double* cost_p = cost ? reinterpret_cast<double*>(da + a_vol) : h->d_cost.as<double>();
      uint32_t* flags_p = flags ? reinterpret_cast<uint32_t*>(da + a_flags) : h->d_flags.as<uint32_t>();
      // the two kernels of the cycle -- lattice kernel, pick + record kernel behind a programmatic edge -- are ONE graph
      // launch; the two small copies stay ordinary stream copies (measured on the B200: as memcpy
      // NODES of the graph they cost 6 us more per cycle than cudaMemcpyAsync)
      std::vector<KernelLaunch> kl;
      FISS_CUDA(h, cudaMemcpyAsync(h->d_ego.p, h->h_in.p, (size_t)B * 48, cudaMemcpyHostToDevice, st));
      if (as_graph) h->record = &kl;
      if (g) {
        rc = eval_grid(h, st, ego_p, B, g, n_max, p, cost_p, flags_p, nullptr, n_stride);
      } else {
        rc = fiss_eval_candidates_dev(h, stream, ego_p, B, h->d_end.as<double>(), C, p, cost_p, flags_p, nullptr, n_stride);
      }
      if (rc == FISS_OK)
        rc = fiss_pick_winners_dev(h, stream, ego_p, B, h->d_end.as<double>(), C, p, cost_p, flags_p,
                                   reinterpret_cast<int32_t*>(da + w_idx), reinterpret_cast<double*>(da + w_cost),
                                   records ? reinterpret_cast<double*>(da + a_rec) : nullptr,
                                   reinterpret_cast<int32_t*>(da + w_meta), n_stride);
      h->record = nullptr;
      if (rc != FISS_OK) return rc;
      if (as_graph) {
        rc = graph_run(h, st, h->graph_host, kl, nullptr, nullptr, 0, nullptr, nullptr, 0);
        if (rc != FISS_OK) return rc;
      }
      FISS_CUDA(h, cudaMemcpyAsync(ha, da, a_end, cudaMemcpyDeviceToHost, st));
      FISS_CUDA(h, cudaStreamSynchronize(st));
      std::memcpy(best_cost, ha + w_cost, (size_t)B * 8);
      std::memcpy(best_idx, ha + w_idx, (size_t)B * 4);
      if (best_meta) std::memcpy(best_meta, ha + w_meta, (size_t)B * 8);
      if (records) std::memcpy(records, ha + a_rec, rec_doubles * 8);
      if (cost) std::memcpy(cost, ha + a_vol, total * 8);
      if (flags) std::memcpy(flags, ha + a_flags, total * 4);
      return FISS_OK;
    }
  }
  FISS_CUDA(h, h->d_best_cost.ensure(w_bytes));
  char* dw = h->d_best_cost.as<char>();
  double* d_bcost = reinterpret_cast<double*>(dw + w_cost);
  int32_t* d_bmeta = reinterpret_cast<int32_t*>(dw + w_meta);
  int32_t* d_bidx = reinterpret_cast<int32_t*>(dw + w_idx);
  if (records) FISS_CUDA(h, h->d_records.ensure(rec_doubles * 8));
  // pinned staging: [ego] in, [winners | records | cost | flags] out -- the big pieces only when the caller's own
  // buffer is not page-locked (a pinned caller buffer is the DMA source / target itself)
  const bool pin_ego = host_is_pinned(ego), pin_rec = host_is_pinned(records), pin_vol = host_is_pinned(cost),
             pin_flags = host_is_pinned(flags);
  const size_t o_win = 0, o_rec = o_win + ((w_bytes + 15) & ~(size_t)15), o_vol = o_rec + (pin_rec ? 0 : rec_doubles * 8),
               o_flags = o_vol + (cost && !pin_vol ? total * 8 : 0), o_end = o_flags + (flags && !pin_flags ? total * 4 : 0);
  FISS_CUDA(h, h->h_out.ensure(o_end));
  char* ho = h->h_out.as<char>();
  const void* ego_src = ego;
  if (!pin_ego) {
    FISS_CUDA(h, h->h_in.ensure((size_t)B * 48));
    std::memcpy(h->h_in.p, ego, (size_t)B * 48);
    ego_src = h->h_in.p;
  }
  FISS_CUDA(h, cudaMemcpyAsync(h->d_ego.p, ego_src, (size_t)B * 48, cudaMemcpyHostToDevice, st));
  void* t_rec = pin_rec ? (void*)records : (void*)(ho + o_rec);
  void* t_vol = pin_vol ? (void*)cost : (void*)(ho + o_vol);
  void* t_flags = pin_flags ? (void*)flags : (void*)(ho + o_flags);
  // A big batch is pipelined in up to kMaxParts pieces of >= 512 problems: the device->host copy of a piece's
  // records / volume runs on a second stream while the kernels of the next piece execute (a piece of 512 problems
  // still covers the GPU with ~6 items per resident CTA; measured on a B200: +5 % at B = 1024, +16 % at B = 2048,
  // -6 % if a batch of 512 is split, hence the floor).
  static const int part_size = std::getenv("FISS_SPLIT_PART") ? std::max(1, std::atoi(std::getenv("FISS_SPLIT_PART"))) : 512;
  const int n_parts = (records || cost || flags) ? std::max(1, std::min(kMaxParts, B / part_size)) : 1;
  if (n_parts > 1 && !h->copy_stream) {
    FISS_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (auto& ev : h->part_done) FISS_CUDA(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  // A failure after the first piece was issued must not return while copies into the caller's buffers are in flight
  // (the caller may free them, and the next call may reallocate the device buffers under the running copy).
  const auto drain = [&]() {
    cudaStreamSynchronize(st);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  };
#define FISS_CUDA_DRAIN(expr)                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      drain();                                                                                 \
      return fail(h, FISS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));       \
    }                                                                                          \
  } while (0)
  for (int part = 0; part < n_parts; ++part) {
    const int b0 = (int)((int64_t)B * part / n_parts), bn = (int)((int64_t)B * (part + 1) / n_parts) - b0;
    const double* ego_p = h->d_ego.as<double>() + (size_t)b0 * 6;
    double* cost_p = h->d_cost.as<double>() + (size_t)b0 * C;
    uint32_t* flags_p = h->d_flags.as<uint32_t>() + (size_t)b0 * C;
    if (g) {
      rc = eval_grid(h, st, ego_p, bn, g, n_max, p, cost_p, flags_p, nullptr, n_stride);
    } else {
      rc = fiss_eval_candidates_dev(h, stream, ego_p, bn, h->d_end.as<double>(), C, p, cost_p, flags_p, nullptr, n_stride);
    }
    if (rc != FISS_OK) {
      drain();
      return rc;
    }
    const size_t rec_off = (size_t)b0 * FISS_REC_ROWS * n_stride, rec_len = (size_t)bn * FISS_REC_ROWS * n_stride;
    rc = fiss_pick_winners_dev(h, stream, ego_p, bn, h->d_end.as<double>(), C, p, cost_p, flags_p, d_bidx + b0, d_bcost + b0,
                               records ? h->d_records.as<double>() + rec_off : nullptr, d_bmeta + 2 * (size_t)b0, n_stride);
    if (rc != FISS_OK) {
      drain();
      return rc;
    }
    cudaStream_t cs = st;
    if (n_parts > 1) {
      FISS_CUDA_DRAIN(cudaEventRecord(h->part_done[part], st));
      FISS_CUDA_DRAIN(cudaStreamWaitEvent(h->copy_stream, h->part_done[part], 0));
      cs = h->copy_stream;
    }
    if (records)
      FISS_CUDA_DRAIN(cudaMemcpyAsync((double*)t_rec + rec_off, h->d_records.as<double>() + rec_off, rec_len * 8,
                                      cudaMemcpyDeviceToHost, cs));
    if (cost)
      FISS_CUDA_DRAIN(cudaMemcpyAsync((double*)t_vol + (size_t)b0 * C, cost_p, (size_t)bn * C * 8, cudaMemcpyDeviceToHost, cs));
    if (flags)
      FISS_CUDA_DRAIN(cudaMemcpyAsync((uint32_t*)t_flags + (size_t)b0 * C, flags_p, (size_t)bn * C * 4, cudaMemcpyDeviceToHost, cs));
  }
  FISS_CUDA_DRAIN(cudaMemcpyAsync(ho + o_win, dw, w_bytes, cudaMemcpyDeviceToHost, st));
  FISS_CUDA_DRAIN(cudaStreamSynchronize(st));
  if (n_parts > 1) FISS_CUDA_DRAIN(cudaStreamSynchronize(h->copy_stream));
#undef FISS_CUDA_DRAIN
  std::memcpy(best_cost, ho + o_win + w_cost, (size_t)B * 8);
  std::memcpy(best_idx, ho + o_win + w_idx, (size_t)B * 4);
  if (best_meta) std::memcpy(best_meta, ho + o_win + w_meta, (size_t)B * 8);
  if (records && !pin_rec) std::memcpy(records, t_rec, rec_doubles * 8);
  if (cost && !pin_vol) std::memcpy(cost, t_vol, total * 8);
  if (flags && !pin_flags) std::memcpy(flags, t_flags, total * 4);
  return FISS_OK;
}

int32_t fiss_plan_lattice_host(fiss_handle* h, void* stream, const double* ego, int32_t B, const double* end,
                               int32_t C, const fiss_params* p, int32_t* best_idx, double* best_cost,
                               int32_t* best_meta, double* records, int32_t n_stride, double* cost, uint32_t* flags) {
  if (!h) return FISS_ERR_INVALID;
  if (!ego || !end || !best_idx || !best_cost || B < 1 || C < 1)
    return fail(h, FISS_ERR_INVALID, "plan_lattice: bad arguments");
  int32_t rc = check_params(h, p);
  if (rc != FISS_OK) return rc;
  rc = check_end_states(h, end, C, n_stride);
  if (rc != FISS_OK) return rc;
  FISS_ON_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  rc = upload_end_states(h, st, end, C);
  if (rc != FISS_OK) return rc;
  return plan_common(h, st, ego, B, C, nullptr, 0, p, best_idx, best_cost, best_meta, records, n_stride, cost, flags);
}

int32_t fiss_eval_grid_dev(fiss_handle* h, void* stream, const double* d_ego, int32_t B, const fiss_grid* g,
                           const fiss_params* p, double* d_cost, uint32_t* d_flags, double* d_mat, int32_t n_stride) {
  if (!h) return FISS_ERR_INVALID;
  if (!d_ego || !d_cost || !d_flags || B < 1) return fail(h, FISS_ERR_INVALID, "eval_grid: bad arguments");
  int32_t rc = check_params(h, p);
  if (rc != FISS_OK) return rc;
  FISS_ON_DEVICE(h);
  int n_max = 0;
  rc = ensure_grid(h, (cudaStream_t)stream, g, p, &n_max);
  if (rc != FISS_OK) return rc;
  return eval_grid(h, (cudaStream_t)stream, d_ego, B, g, n_max, p, d_cost, d_flags, d_mat, n_stride);
}

int32_t fiss_plan_grid_host(fiss_handle* h, void* stream, const double* ego, int32_t B, const fiss_grid* g,
                            const fiss_params* p, int32_t* best_idx, double* best_cost, int32_t* best_meta,
                            double* records, int32_t n_stride, double* cost, uint32_t* flags) {
  if (!h) return FISS_ERR_INVALID;
  if (!ego || !best_idx || !best_cost || B < 1) return fail(h, FISS_ERR_INVALID, "plan_grid: bad arguments");
  int32_t rc = check_params(h, p);
  if (rc != FISS_OK) return rc;
  FISS_ON_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  int n_max = 0;
  rc = ensure_grid(h, st, g, p, &n_max);
  if (rc != FISS_OK) return rc;
  if (n_stride < n_max) return fail(h, FISS_ERR_INVALID, "n_stride is smaller than a candidate's step count");
  const int C = g->nd * g->nv * g->nt;
  return plan_common(h, st, ego, B, C, g, n_max, p, best_idx, best_cost, best_meta, records, n_stride, cost, flags);
}

int32_t fiss_plan_grid_dev(fiss_handle* h, void* stream, const double* d_ego, int32_t B, const fiss_grid* g,
                           const fiss_params* p, double* d_cost, uint32_t* d_flags, double* d_mat, int32_t* d_best_idx,
                           double* d_best_cost, int32_t* d_best_meta, double* d_records, int32_t n_stride) {
  if (!h) return FISS_ERR_INVALID;
  if (!d_ego || !d_cost || !d_flags || !d_best_idx || !d_best_cost || B < 1)
    return fail(h, FISS_ERR_INVALID, "plan_grid_dev: bad arguments");
  int32_t rc = check_params(h, p);
  if (rc != FISS_OK) return rc;
  FISS_ON_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  int n_max = 0;
  rc = ensure_grid(h, st, g, p, &n_max);
  if (rc != FISS_OK) return rc;
  const int C = g->nd * g->nv * g->nt;
  // A batch that fills the GPU is issued launch by launch, so that consecutive steps chain (see eval_grid: the next step's
  // CTAs start on the SMs this step's last items leave idle); a small one -- launch-latency bound -- as one graph launch.
  static const bool chain_on = std::getenv("FISS_CHAIN") ? std::atoi(std::getenv("FISS_CHAIN")) != 0 : true;
  const bool fills_gpu = (int64_t)B * g->nt >= 2 * (int64_t)h->sm_count;
  const bool as_graph = graphs_enabled() && !(chain_on && fills_gpu);
  std::vector<KernelLaunch> kl;
  if (as_graph) h->record = &kl;
  rc = eval_grid(h, st, d_ego, B, g, n_max, p, d_cost, d_flags, d_mat, n_stride);
  if (rc == FISS_OK)
    rc = fiss_pick_winners_dev(h, stream, d_ego, B, h->d_end.as<double>(), C, p, d_cost, d_flags, d_best_idx, d_best_cost,
                               d_records, d_best_meta, n_stride);
  h->record = nullptr;
  if (rc != FISS_OK || !as_graph) return rc;
  return graph_run(h, st, h->graph_dev, kl, nullptr, nullptr, 0, nullptr, nullptr, 0);
}

int32_t fiss_plan_grid_submit(fiss_handle* h, void* stream, int32_t lane, const double* ego, int32_t B, const fiss_grid* g,
                              const fiss_params* p, int32_t* best_idx, double* best_cost, int32_t* best_meta,
                              double* records, int32_t n_stride) {
  if (!h) return FISS_ERR_INVALID;
  if (lane < 0 || lane >= kLanes) return fail(h, FISS_ERR_INVALID, "plan_grid_submit: lane must be 0 .. FISS_LANES - 1");
  if (!ego || !best_idx || !best_cost || B < 1) return fail(h, FISS_ERR_INVALID, "plan_grid_submit: bad arguments");
  Lane& ln = h->lanes[lane];
  if (ln.pending) return fail(h, FISS_ERR_STATE, "plan_grid_submit: the lane still has a call in flight (fiss_plan_grid_wait it first)");
  int32_t rc = check_params(h, p);
  if (rc != FISS_OK) return rc;
  FISS_ON_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  int n_max = 0;
  rc = ensure_grid(h, st, g, p, &n_max);
  if (rc != FISS_OK) return rc;
  if (n_stride < n_max) return fail(h, FISS_ERR_INVALID, "n_stride is smaller than a candidate's step count");
  if (!h->copy_stream) {
    FISS_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (auto& ev : h->part_done) FISS_CUDA(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  if (!ln.kernels_done) {
    FISS_CUDA(h, cudaEventCreateWithFlags(&ln.kernels_done, cudaEventDisableTiming));
    FISS_CUDA(h, cudaEventCreateWithFlags(&ln.copied, cudaEventDisableTiming));
  }
  const int C = g->nd * g->nv * g->nt;
  const size_t total = (size_t)B * C, rec_doubles = records ? (size_t)B * FISS_REC_ROWS * n_stride : 0;
  const size_t w_cost = 0, w_meta = (size_t)B * 8, w_idx = (size_t)B * 16, w_bytes = (size_t)B * 20;
  FISS_CUDA(h, ln.d_ego.ensure((size_t)B * 48));
  FISS_CUDA(h, ln.d_cost.ensure(total * 8));
  FISS_CUDA(h, ln.d_flags.ensure(total * 4));
  FISS_CUDA(h, ln.d_win.ensure(w_bytes));
  if (records) FISS_CUDA(h, ln.d_records.ensure(rec_doubles * 8));
  const bool pin_rec = host_is_pinned(records);
  const size_t o_rec = (w_bytes + 15) & ~(size_t)15;
  FISS_CUDA(h, ln.h_out.ensure(o_rec + (pin_rec ? 0 : rec_doubles * 8)));
  FISS_CUDA(h, ln.h_in.ensure((size_t)B * 48));
  const bool as_graph = graphs_enabled();
  const void* src = ego;
  if (!host_is_pinned(ego)) {
    std::memcpy(ln.h_in.p, ego, (size_t)B * 48);
    src = ln.h_in.p;
  }
  char* dw = ln.d_win.as<char>();
  std::vector<KernelLaunch> kl;
  FISS_CUDA(h, cudaMemcpyAsync(ln.d_ego.p, src, (size_t)B * 48, cudaMemcpyHostToDevice, st));
  if (as_graph) h->record = &kl;
  rc = eval_grid(h, st, ln.d_ego.as<double>(), B, g, n_max, p, ln.d_cost.as<double>(), ln.d_flags.as<uint32_t>(), nullptr, n_stride);
  if (rc == FISS_OK)
    rc = fiss_pick_winners_dev(h, stream, ln.d_ego.as<double>(), B, h->d_end.as<double>(), C, p, ln.d_cost.as<double>(),
                               ln.d_flags.as<uint32_t>(), reinterpret_cast<int32_t*>(dw + w_idx),
                               reinterpret_cast<double*>(dw + w_cost), records ? ln.d_records.as<double>() : nullptr,
                               reinterpret_cast<int32_t*>(dw + w_meta), n_stride);
  h->record = nullptr;
  if (rc != FISS_OK) return rc;
  if (as_graph) {
    rc = graph_run(h, st, ln.graph, kl, nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (rc != FISS_OK) return rc;
  }
  // the copy-back runs on the handle's second stream: the kernels of the next submit (other lane) execute under it
  FISS_CUDA(h, cudaEventRecord(ln.kernels_done, st));
  FISS_CUDA(h, cudaStreamWaitEvent(h->copy_stream, ln.kernels_done, 0));
  char* ho = ln.h_out.as<char>();
  FISS_CUDA(h, cudaMemcpyAsync(ho, dw, w_bytes, cudaMemcpyDeviceToHost, h->copy_stream));
  if (records)
    FISS_CUDA(h, cudaMemcpyAsync(pin_rec ? (void*)records : (void*)(ho + o_rec), ln.d_records.p, rec_doubles * 8,
                                 cudaMemcpyDeviceToHost, h->copy_stream));
  FISS_CUDA(h, cudaEventRecord(ln.copied, h->copy_stream));
  ln.pending = true;
  ln.B = B;
  ln.n_stride = n_stride;
  ln.best_idx = best_idx;
  ln.best_cost = best_cost;
  ln.best_meta = best_meta;
  ln.records = records;
  ln.rec_pinned = pin_rec;
  return FISS_OK;
}

int32_t fiss_plan_grid_wait(fiss_handle* h, int32_t lane) {
  if (!h) return FISS_ERR_INVALID;
  if (lane < 0 || lane >= kLanes) return fail(h, FISS_ERR_INVALID, "plan_grid_wait: lane must be 0 .. FISS_LANES - 1");
  Lane& ln = h->lanes[lane];
  if (!ln.pending) return fail(h, FISS_ERR_STATE, "plan_grid_wait: nothing in flight on this lane");
  FISS_ON_DEVICE(h);
  ln.pending = false;
  FISS_CUDA(h, cudaEventSynchronize(ln.copied));
  const size_t B = (size_t)ln.B, w_cost = 0, w_meta = B * 8, w_idx = B * 16, w_bytes = B * 20;
  const char* ho = ln.h_out.as<char>();
  std::memcpy(ln.best_cost, ho + w_cost, B * 8);
  std::memcpy(ln.best_idx, ho + w_idx, B * 4);
  if (ln.best_meta) std::memcpy(ln.best_meta, ho + w_meta, B * 8);
  if (ln.records && !ln.rec_pinned)
    std::memcpy(ln.records, ho + ((w_bytes + 15) & ~(size_t)15), B * FISS_REC_ROWS * ln.n_stride * 8);
  return FISS_OK;
}

// ---- cross-GPU pick ------------------------------------------------------------------------------
int32_t fiss_comm_unique_id(void* out128, const char* nccl_path) {
  if (!out128) return fail(nullptr, FISS_ERR_INVALID, "comm_unique_id: out is NULL");
  if (!nccl_load(nccl_path)) return fail(nullptr, FISS_ERR_NCCL, g_nccl.err);
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  FISS_NCCL(nullptr, g_nccl.GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
  return FISS_OK;
}

int32_t fiss_comm_init(fiss_handle* h, const void* id128, int32_t nranks, int32_t rank, const char* nccl_path) {
  if (!h) return FISS_ERR_INVALID;
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, FISS_ERR_INVALID, "comm_init: bad arguments");
  if (h->comm) return fail(h, FISS_ERR_STATE, "comm_init: the handle already owns a communicator");
  if (!nccl_load(nccl_path)) return fail(h, FISS_ERR_NCCL, g_nccl.err);
  FISS_ON_DEVICE(h);
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  FISS_NCCL(h, g_nccl.CommInitRank(&comm, nranks, id, rank));
  h->comm = comm;
  h->comm_rank = rank;
  h->comm_size = nranks;
  return FISS_OK;
}

int32_t fiss_comm_destroy(fiss_handle* h) {
  if (!h) return FISS_ERR_INVALID;
  if (h->comm) {
    FISS_ON_DEVICE(h);
    ncclComm_t comm = static_cast<ncclComm_t>(h->comm);
    h->comm = nullptr;
    FISS_NCCL(h, g_nccl.CommDestroy(comm));
  }
  h->comm_size = 1;
  h->comm_rank = 0;
  return FISS_OK;
}

int32_t fiss_allreduce_pick(fiss_handle* h, void* comm, int32_t nranks, int32_t rank, void* stream, int32_t B,
                            int64_t id_inner, int64_t id_outer, int64_t id_offset, int32_t* d_best_idx, double* d_best_cost, int32_t* d_best_meta,
                            double* d_records, int32_t n_stride) {
  if (!h) return FISS_ERR_INVALID;
  if (!d_best_idx || !d_best_cost || B < 1 || (d_records && n_stride < 1))
    return fail(h, FISS_ERR_INVALID, "allreduce_pick: bad arguments");
  if (!comm) {  // the handle's own communicator (fiss_comm_init)
    comm = h->comm;
    nranks = h->comm_size;
    rank = h->comm_rank;
  }
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(h, FISS_ERR_INVALID, "allreduce_pick: bad rank / size");
  if (id_inner < 1 || id_outer < 0 || id_offset < 0) return fail(h, FISS_ERR_INVALID, "allreduce_pick: bad id map");
  if (nranks > 1 && !comm) return fail(h, FISS_ERR_STATE, "allreduce_pick: no communicator (fiss_comm_init, or pass one)");
  if (nranks > 1 && !nccl_load(nullptr)) return fail(h, FISS_ERR_NCCL, g_nccl.err);
  FISS_ON_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  // [slot table B x nranks x 2 | exchange block B x (16 n_stride + 1)] of 64-bit words
  const size_t table_words = (size_t)B * nranks * 2;
  const size_t rec_words = d_records ? (size_t)FISS_REC_ROWS * n_stride : 0, xch_words = (size_t)B * (rec_words + 1);
  FISS_CUDA(h, h->d_pick.ensure((table_words + xch_words) * 8));
  unsigned long long* table = h->d_pick.as<unsigned long long>();
  unsigned long long* xch = table + table_words;
  const int threads = 128;
  fiss::fiss_pick_pack_kernel<<<(unsigned)((B * nranks + threads - 1) / threads), threads, 0, st>>>(
      B, nranks, rank, id_inner, id_outer, id_offset, d_best_idx, d_best_cost, table);
  h->launches++;
  FISS_CUDA(h, cudaGetLastError());
  // THE collective of the pick: one all-reduce (MIN over 64-bit keys), 16 B x nranks per problem
  if (nranks > 1)
    FISS_NCCL(h, g_nccl.AllReduce(table, table, table_words, ncclUint64, ncclMin, static_cast<ncclComm_t>(comm), st));
  fiss::fiss_pick_select_kernel<<<(unsigned)B, 256, 0, st>>>(B, nranks, rank, table, d_best_idx, d_best_cost, d_best_meta,
                                                            d_records, (int)rec_words, xch);
  h->launches++;
  FISS_CUDA(h, cudaGetLastError());
  // the winners' records travel from their owners: every other rank contributes zero words to an integer SUM, which
  // reproduces the owner's bit patterns exactly (NaN payloads included) -- no root, hence no host round trip
  if (nranks > 1)
    FISS_NCCL(h, g_nccl.AllReduce(xch, xch, xch_words, ncclUint64, ncclSum, static_cast<ncclComm_t>(comm), st));
  fiss::fiss_pick_unpack_kernel<<<(unsigned)B, 256, 0, st>>>(B, xch, d_best_idx, d_best_meta, d_records, (int)rec_words);
  h->launches++;
  FISS_CUDA(h, cudaGetLastError());
  return FISS_OK;
}

int32_t fiss_eval_end_states_host(fiss_handle* h, void* stream, const double* ego6, const double* end, int32_t N,
                                  const fiss_params* p, double* cost, uint32_t* flags, double* records,
                                  int32_t n_stride) {
  if (!h) return FISS_ERR_INVALID;
  if (!ego6 || !end || !cost || !flags || N < 1) return fail(h, FISS_ERR_INVALID, "eval_end_states: bad arguments");
  int32_t rc = check_params(h, p);
  if (rc != FISS_OK) return rc;
  rc = check_end_states(h, end, N, n_stride);
  if (rc != FISS_OK) return rc;
  FISS_ON_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rec_doubles = records ? (size_t)N * FISS_REC_ROWS * n_stride : 0;
  // one block in ([ego 48 B | end N x 32 B]) and one block out ([cost N x 8 | flags N x 4]): this call sits on the
  // latency path of the FISS / FISS+ searches (a handful of end states per call), so copies are what it costs
  const size_t in_bytes = 48 + (size_t)N * 32, out_bytes = (size_t)N * 12;
  const size_t in_pad = (in_bytes + 15) & ~(size_t)15;
  FISS_CUDA(h, h->d_es.ensure(in_pad + out_bytes));
  char* d_in = h->d_es.as<char>();
  char* d_out = d_in + in_pad;
  double* d_ego = reinterpret_cast<double*>(d_in);
  double* d_end = reinterpret_cast<double*>(d_in + 48);
  double* d_cost = reinterpret_cast<double*>(d_out);
  uint32_t* d_flags = reinterpret_cast<uint32_t*>(d_out + (size_t)N * 8);
  if (records) FISS_CUDA(h, h->d_records.ensure(rec_doubles * 8));
  FISS_CUDA(h, h->h_in.ensure(in_bytes));
  const bool pin_rec = host_is_pinned(records);
  const size_t o_rec = (out_bytes + 15) & ~(size_t)15;
  FISS_CUDA(h, h->h_out.ensure(o_rec + (pin_rec ? 0 : rec_doubles * 8)));
  char* ho = h->h_out.as<char>();
  std::memcpy(h->h_in.p, ego6, 48);
  std::memcpy(h->h_in.as<char>() + 48, end, (size_t)N * 32);
  FISS_CUDA(h, cudaMemcpyAsync(d_in, h->h_in.p, in_bytes, cudaMemcpyHostToDevice, st));
  if (records) {
    rc = fiss_full_records_dev(h, stream, d_ego, d_end, nullptr, N, p, h->d_records.as<double>(), d_cost, d_flags, n_stride);
  } else {
    rc = fiss_eval_candidates_dev(h, stream, d_ego, 1, d_end, N, p, d_cost, d_flags, nullptr, n_stride);
  }
  if (rc != FISS_OK) return rc;
  FISS_CUDA(h, cudaMemcpyAsync(ho, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
  void* t_rec = pin_rec ? (void*)records : (void*)(ho + o_rec);
  if (records) FISS_CUDA(h, cudaMemcpyAsync(t_rec, h->d_records.p, rec_doubles * 8, cudaMemcpyDeviceToHost, st));
  FISS_CUDA(h, cudaStreamSynchronize(st));
  std::memcpy(cost, ho, (size_t)N * 8);
  std::memcpy(flags, ho + (size_t)N * 8, (size_t)N * 4);
  if (records && !pin_rec) std::memcpy(records, t_rec, rec_doubles * 8);
  return FISS_OK;
}

}  // extern "C"
