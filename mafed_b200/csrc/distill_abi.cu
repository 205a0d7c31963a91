// C ABI (include/mafed_distill.h) over the sm_100a kernels.  No torch types, no allocation, no host
// synchronisation: every entry point validates its arguments, picks a launch geometry and enqueues
// kernels on the caller's stream.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "distill_common.cuh"
#include "distill_epilogue.cuh"
#include "distill_host.cuh"
#include "distill_ldg.cuh"
#include "distill_tma.cuh"

namespace mafed {
namespace {

enum Pass { kPassFwd = 0, kPassBwd = 1, kPassFused = 2 };

std::atomic<int> g_variant{0};   // 0 default, 1 ldg, 2 tma
std::atomic<int> g_tune[24] = {};  // experiment knobs, see mafed_distill_set_tuning

// per-pass keys: base + pass (fwd, bwd, fused)
enum TuneKey { kTuneTmaStages = 0, kTuneTmaRows = 3, kTuneVariant = 6, kTuneTmaWarps = 9, kTuneLdgBlocksPerSm = 10,
               kTuneBwdForward = 11, kTuneGridMul = 12, kTuneLoadPolicy = 13, kTuneStorePolicy = 14, kTuneNoPdl = 15,
               kTuneNoInlineScale = 16, kTuneNoTail = 17 };

// Launch with programmatic dependent launch enabled: the kernel may start its prologue while its
// predecessor in the stream is finishing; all kernels here call griddepcontrol.wait before touching
// global memory, so stream-order semantics are unchanged.
template <typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_tune[kTuneNoPdl].load() ? 0 : 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct DeviceInfo {
  int sm_count = 0;
  int smem_optin = 0;
  bool ok = false;
};

const DeviceInfo& device_info() {
  static DeviceInfo info[16];
  static DeviceInfo none;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return none;
  DeviceInfo& d = info[dev];
  if (!d.ok) {
    cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    d.ok = d.sm_count > 0;
  }
  return d;
}

int check_shape(const mafed_shape_t* sh) {
  if (sh == nullptr) return MAFED_E_ARG;
  if (sh->n_layers < 1 || sh->n_layers > kMaxLayers) return MAFED_E_ARG;
  if (sh->B < 1 || sh->T < 1 || sh->D < 1 || sh->n_vis < 0 || sh->n_vis > sh->T) return MAFED_E_ARG;
  if (sh->dtype < MAFED_F32 || sh->dtype > MAFED_F16) return MAFED_E_DTYPE;
  if (sh->loss_kind != MAFED_LOSS_MSE && sh->loss_kind != MAFED_LOSS_COSINE) return MAFED_E_DTYPE;
  return 0;
}

size_t elem_size(int dtype) { return dtype == MAFED_F32 ? 4 : 2; }

// Fill the geometry part of PathParams.  CLS mode (distillation.py:251-257) is the same path over
// B rows (position 0 of every sample), all of them "visual", T*D elements apart.
void fill_geometry(const mafed_shape_t& sh, PathParams& p) {
  p.n_layers = sh.n_layers;
  p.D = sh.D;
  if (sh.cls) {
    p.n_rows = sh.B;
    p.row_stride = (long long)sh.T * sh.D;
    p.T = 1;
    p.n_vis = 1;
    p.txt = 0;
  } else {
    p.n_rows = (long long)sh.B * sh.T;
    p.row_stride = sh.D;
    p.T = sh.T;
    p.n_vis = sh.n_vis;
    p.txt = sh.T - sh.n_vis;
  }
}

bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

bool needs_mask(const mafed_shape_t& sh) { return !sh.cls && sh.T > sh.n_vis; }
long long mask_entries(const mafed_shape_t& sh) { return needs_mask(sh) ? (long long)sh.B * (sh.T - sh.n_vis) : 0; }
double vis_rows(const mafed_shape_t& sh) { return sh.cls ? (double)sh.B : (double)sh.B * (double)sh.n_vis; }

// Arrival counters of the in-kernel tail (distill_tma.cuh).  A counter is zero whenever no kernel is using it
// (the last CTA resets it), so launches take them round-robin: two kernels can only share one if 128 tailed
// launches are in flight at once on one device, and each of them is a persistent whole-GPU kernel.  Launches
// recorded into a CUDA graph keep their counter for every replay, so they draw from a separate half of the
// pool: an eager launch on another stream can never land on the counter of a graph that is replaying.
constexpr unsigned kDoneSlots = 256;
__device__ unsigned int g_tail_done[kDoneSlots];
std::atomic<unsigned> g_tail_next{0};
std::atomic<unsigned> g_tail_next_captured{0};

unsigned int* next_tail_counter(cudaStream_t st) {
  static unsigned int* base[16] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  if (base[dev] == nullptr) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_tail_done) != cudaSuccess) return nullptr;
    base[dev] = reinterpret_cast<unsigned int*>(p);
  }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) return nullptr;
  constexpr unsigned half = kDoneSlots / 2;
  if (cap == cudaStreamCaptureStatusActive) return base[dev] + half + (g_tail_next_captured.fetch_add(1) % half);
  return base[dev] + (g_tail_next.fetch_add(1) % half);
}

// ---------------------------------------------------------------- launch helpers
template <typename K>
int blocks_per_sm(K kernel, int threads, size_t dyn_smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, dyn_smem) != cudaSuccess) n = 1;
  return n < 1 ? 1 : n;
}

long long clamp_grid(long long grid, long long total) {
  if (grid > total) grid = total;
  if (grid > kMaxPartials) grid = kMaxPartials;
  return grid < 1 ? 1 : grid;
}

template <typename T, int CPL, int RPI, int LOSS>
int launch_ldg(const PathParams& p, int pass, cudaStream_t st) {
  static int occ[3] = {0, 0, 0};
  const DeviceInfo& dv = device_info();
  if constexpr (LOSS == kLossL2Norm) {
    if (pass != kPassFwd) return MAFED_E_ARG;  // the token-norm reduction has no backward
    if (occ[0] == 0) occ[0] = blocks_per_sm(k_fwd_ldg<T, CPL, RPI, LOSS>, kLdgThreads, 0);
  } else if (occ[pass] == 0) {
    occ[pass] = pass == kPassFwd   ? blocks_per_sm(k_fwd_ldg<T, CPL, RPI, LOSS>, kLdgThreads, 0)
                : pass == kPassBwd ? blocks_per_sm(k_bwd_ldg<T, CPL, RPI, LOSS, kBackward>, kLdgThreads, 0)
                                   : blocks_per_sm(k_bwd_ldg<T, CPL, RPI, LOSS, kFused>, kLdgThreads, 0);
  }
  int per_sm = occ[pass];
  const int cap = g_tune[kTuneLdgBlocksPerSm].load();
  if (cap > 0 && cap < per_sm) per_sm = cap;
  const long long rows_per_iter = (long long)kLdgWarps * RPI;
  const long long total = ((p.n_rows + rows_per_iter - 1) / rows_per_iter) * p.n_layers;
  const unsigned grid = (unsigned)clamp_grid((long long)dv.sm_count * per_sm, total);
  if constexpr (LOSS == kLossL2Norm) {
    launch_pdl(k_fwd_ldg<T, CPL, RPI, LOSS>, grid, kLdgThreads, 0, st, p);
  } else {
    if (pass == kPassFwd) launch_pdl(k_fwd_ldg<T, CPL, RPI, LOSS>, grid, kLdgThreads, 0, st, p);
    else if (pass == kPassBwd) launch_pdl(k_bwd_ldg<T, CPL, RPI, LOSS, kBackward>, grid, kLdgThreads, 0, st, p);
    else launch_pdl(k_bwd_ldg<T, CPL, RPI, LOSS, kFused>, grid, kLdgThreads, 0, st, p);
  }
  return (int)cudaPeekAtLastError();
}

template <typename T, int LOSS>
int dispatch_ldg(const PathParams& p, int pass, cudaStream_t st) {
  const int cpl = (p.n_chunks + 31) / 32;
  if (cpl <= 1) return launch_ldg<T, 1, 4, LOSS>(p, pass, st);
  if (cpl <= 2) return launch_ldg<T, 2, 4, LOSS>(p, pass, st);
  if (cpl <= 3) return launch_ldg<T, 3, 2, LOSS>(p, pass, st);
  if (cpl <= 4) return launch_ldg<T, 4, 2, LOSS>(p, pass, st);
  if (cpl <= 6) return launch_ldg<T, 6, 1, LOSS>(p, pass, st);
  return launch_ldg<T, 8, 1, LOSS>(p, pass, st);  // multi-pass for rows longer than 4 KB
}

// Any D / any alignment: element-wise kernels.  The fused pass is simply forward then backward.
template <typename T, int LOSS>
int launch_generic(const PathParams& p, int pass, cudaStream_t st) {
  const DeviceInfo& dv = device_info();
  const long long total = ((p.n_rows + kLdgWarps - 1) / kLdgWarps) * p.n_layers;
  const unsigned grid = (unsigned)clamp_grid((long long)dv.sm_count * 4, total);
  if (pass == kPassFwd || pass == kPassFused) launch_pdl(k_fwd_generic<T, LOSS>, grid, kLdgThreads, 0, st, p);
  if constexpr (LOSS != kLossL2Norm) {
    if (pass == kPassBwd || pass == kPassFused)
      launch_pdl(k_bwd_generic<T, LOSS>, grid, kLdgThreads, 0, st, p, pass == kPassFused ? 1 : 0);
  }
  return (int)cudaPeekAtLastError();
}

bool tma_geometry(const PathParams& p, int pass, TmaGeom& geo, int loss = MAFED_LOSS_MSE) {
  const DeviceInfo& dv = device_info();
  const long long row_bytes = (long long)p.n_chunks * 16;
  if (row_bytes > 32768) return false;
  const long long budget = (long long)dv.smem_optin - 16 * 1024;  // static smem + slack
  int rows = g_tune[kTuneTmaRows + pass].load();
  bool cosine_fine = false;
  if (rows <= 0) {
    // Measured on B200 (profiles/r01_call3_sweep_step.json): a ring of 2 stages x 64-72 KB per SM is the
    // sweet spot for every pass; deeper rings (>= 192 KB in flight per SM) cost 4-8 % of HBM throughput.
    // The cosine gradient needs two sweeps over a row with a warp reduction in between, so a stage drains more
    // slowly: with rows >= 4 KB a finer ring (4 stages x 32 KB) keeps more rows in different phases at once --
    // 5-8 % faster on the 1B shape on two boxes (profiles/r01b_sweep_ring*.json), neutral or worse for short rows.
    cosine_fine = loss == MAFED_LOSS_COSINE && pass != kPassFwd && row_bytes >= 4096;
    const long long stage_target = cosine_fine ? 32 * 1024 : 72 * 1024;
    rows = (int)(stage_target / (2 * row_bytes));
    if (rows >= 8) rows &= ~7;
  }
  if (rows > kTmaMaxRows) rows = kTmaMaxRows;
  if (rows < 1) rows = 1;
  if (2 * rows * row_bytes > budget) rows = (int)(budget / (2 * row_bytes));
  if (rows < 1) return false;
  geo.rows = rows;
  geo.stage_bytes = (int)(2 * rows * row_bytes);
  int stages = g_tune[kTuneTmaStages + pass].load();
  if (stages <= 0) stages = cosine_fine ? 4 : ((2 * geo.stage_bytes >= 96 * 1024) ? 2 : 3);
  if (stages > kTmaMaxStages) stages = kTmaMaxStages;
  while (stages > 1 && (long long)stages * geo.stage_bytes > budget) --stages;
  geo.stages = stages;
  return true;
}

template <typename T, int LOSS, int NCW>
int launch_tma(const PathParams& p, const TmaGeom& geo, int pass, cudaStream_t st) {
  // the opt-in shared-memory size is a per-device function attribute: remember it per (pass, device)
  static unsigned attr_devices[3] = {0u, 0u, 0u};
  const DeviceInfo& dv = device_info();
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned dev_bit = 1u << (dev & 31);
  const size_t dyn = (size_t)geo.stages * geo.stage_bytes;
  const int max_dyn = dv.smem_optin - 16 * 1024;
  if constexpr (LOSS == kLossL2Norm) {
    if (pass != kPassFwd) return MAFED_E_ARG;
    if (!(attr_devices[0] & dev_bit)) {
      cudaError_t e = cudaFuncSetAttribute(k_fwd_tma<T, LOSS, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
      if (e != cudaSuccess) return (int)e;
      attr_devices[0] |= dev_bit;
    }
  } else if (!(attr_devices[pass] & dev_bit)) {
    cudaError_t e =
        pass == kPassFwd ? cudaFuncSetAttribute(k_fwd_tma<T, LOSS, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn)
        : pass == kPassBwd
            ? cudaFuncSetAttribute(k_bwd_tma<T, LOSS, NCW, kBackward>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn)
            : cudaFuncSetAttribute(k_bwd_tma<T, LOSS, NCW, kFused>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    if (e != cudaSuccess) return (int)e;
    attr_devices[pass] |= dev_bit;
  }
  const long long total = ((p.n_rows + geo.rows - 1) / geo.rows) * p.n_layers;
  int mul = g_tune[kTuneGridMul].load();
  if (mul <= 0) mul = 1;
  const unsigned grid = (unsigned)clamp_grid((long long)dv.sm_count * mul, total);
  constexpr int threads = (NCW + 1) * 32;
  if constexpr (LOSS == kLossL2Norm) {
    launch_pdl(k_fwd_tma<T, LOSS, NCW>, grid, threads, dyn, st, p, geo);
  } else {
    if (pass == kPassFwd) launch_pdl(k_fwd_tma<T, LOSS, NCW>, grid, threads, dyn, st, p, geo);
    else if (pass == kPassBwd) launch_pdl(k_bwd_tma<T, LOSS, NCW, kBackward>, grid, threads, dyn, st, p, geo);
    else launch_pdl(k_bwd_tma<T, LOSS, NCW, kFused>, grid, threads, dyn, st, p, geo);
  }
  return (int)cudaPeekAtLastError();
}

template <typename T, int LOSS>
int dispatch_typed(PathParams& p, bool vector_ok, int pass, cudaStream_t st) {
  if (!vector_ok) return launch_generic<T, LOSS>(p, pass, st);
  int variant = g_tune[kTuneVariant + pass].load();
  if (variant == 0) variant = g_variant.load();
  if (variant == 0) variant = 2;
  if (variant == 2) {
    TmaGeom geo;
    if (tma_geometry(p, pass, geo, LOSS)) {
      // 16 consumer warps: two stages are drained concurrently when a stage holds <= 8 rows (+2.6 % on the
      // 1B shape, neutral elsewhere; profiles/r01_call4_sweep_extra.json)
      const int ncw = g_tune[kTuneTmaWarps].load();
      if (ncw == 8) return launch_tma<T, LOSS, 8>(p, geo, pass, st);
      return launch_tma<T, LOSS, 16>(p, geo, pass, st);
    }
  }
  return dispatch_ldg<T, LOSS>(p, pass, st);
}

template <typename T>
int dispatch_loss(PathParams& p, int loss, bool vector_ok, int pass, cudaStream_t st) {
  if (loss == MAFED_LOSS_MSE) return dispatch_typed<T, MAFED_LOSS_MSE>(p, vector_ok, pass, st);
  if (loss == kLossL2Norm) return dispatch_typed<T, kLossL2Norm>(p, vector_ok, pass, st);
  return dispatch_typed<T, MAFED_LOSS_COSINE>(p, vector_ok, pass, st);
}

int dispatch(const mafed_shape_t& sh, PathParams& p, int pass, cudaStream_t st, int loss_override = -1) {
  // vector path: rows are whole, 16-byte aligned chunks
  const size_t es = elem_size(sh.dtype);
  bool vector_ok = ((size_t)sh.D * es) % 16 == 0 && ((size_t)p.row_stride * es) % 16 == 0;
  for (int l = 0; l < sh.n_layers && vector_ok; ++l)
    vector_ok = aligned_to(p.s[l], 16) && aligned_to(p.t[l], 16) && (pass == kPassFwd || aligned_to(p.g[l], 16));
  p.n_chunks = vector_ok ? (int)((size_t)sh.D * es / 16) : 0;
  p.load_policy = g_tune[kTuneLoadPolicy].load();
  p.store_policy = g_tune[kTuneStorePolicy].load();
  const int loss = loss_override >= 0 ? loss_override : sh.loss_kind;
  switch (sh.dtype) {
    case MAFED_F32: return dispatch_loss<float>(p, loss, vector_ok, pass, st);
    case MAFED_BF16: return dispatch_loss<__nv_bfloat16>(p, loss, vector_ok, pass, st);
    default: return dispatch_loss<__half>(p, loss, vector_ok, pass, st);
  }
}

// Will `dispatch` take the TMA-ring kernels for this call?  (Same decision, made ahead of the launch.)
bool uses_tma(const mafed_shape_t& sh, const PathParams& p, int pass) {
  const size_t es = elem_size(sh.dtype);
  bool vector_ok = ((size_t)sh.D * es) % 16 == 0 && ((size_t)p.row_stride * es) % 16 == 0;
  for (int l = 0; l < sh.n_layers && vector_ok; ++l)
    vector_ok = aligned_to(p.s[l], 16) && aligned_to(p.t[l], 16) && (pass == kPassFwd || aligned_to(p.g[l], 16));
  if (!vector_ok) return false;
  int variant = g_tune[kTuneVariant + pass].load();
  if (variant == 0) variant = g_variant.load();
  if (variant != 0 && variant != 2) return false;
  PathParams q = p;
  q.n_chunks = (int)((size_t)sh.D * es / 16);
  TmaGeom geo;
  return tma_geometry(q, pass, geo);   // eligibility does not depend on the loss kind
}

// Common argument checks + pointer-table copy for the three streaming passes.
int fill_params(const mafed_shape_t* shape, const void* const* student_ptrs, const void* const* teacher_ptrs,
                void* const* grad_ptrs, const int64_t* attn_mask, PathParams& p) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!student_ptrs || !teacher_ptrs || (needs_mask(*shape) && !attn_mask)) return MAFED_E_ARG;
  if (!device_info().ok) return MAFED_E_NODEVICE;
  memset(&p, 0, sizeof(p));
  fill_geometry(*shape, p);
  const size_t es = elem_size(shape->dtype);
  for (int l = 0; l < shape->n_layers; ++l) {
    if (!student_ptrs[l] || !teacher_ptrs[l]) return MAFED_E_ARG;
    if (!aligned_to(student_ptrs[l], es) || !aligned_to(teacher_ptrs[l], es)) return MAFED_E_ALIGN;
    p.s[l] = student_ptrs[l];
    p.t[l] = teacher_ptrs[l];
    if (grad_ptrs) {
      if (!aligned_to(grad_ptrs[l], es)) return MAFED_E_ALIGN;
      p.g[l] = grad_ptrs[l];
    }
  }
  p.mask = attn_mask;
  return 0;
}

}  // namespace
}  // namespace mafed

// Host handle of the peer-memory communicator (see distill_comm.cuh).
struct mafed_comm {
  int world = 0;
  int rank = 0;
  void* local = nullptr;
  void* peers[mafed::kCommMaxRanks] = {};
  long long timeout_cycles = 0;
  double cycles_per_second = 1.9e9;
  bool loopback = false;
};

namespace mafed {
namespace {

CommDev comm_dev(const mafed_comm* c) {
  CommDev d;
  memset(&d, 0, sizeof(d));
  if (c == nullptr || c->world <= 1) return d;
  d.world = c->world;
  d.rank = c->rank;
  d.timeout_cycles = c->timeout_cycles;
  for (int r = 0; r < c->world; ++r) d.ll[r] = reinterpret_cast<unsigned long long*>(c->peers[r]);
  char* tail = reinterpret_cast<char*>(c->local) + kCommDataBytes;
  d.epoch = reinterpret_cast<unsigned long long*>(tail);
  d.status = reinterpret_cast<int*>(tail + 16);
  d.trace = reinterpret_cast<unsigned long long*>(tail + 32);
  return d;
}

int launch_scalar_stage(const mafed_shape_t& sh, const mafed_weights_t* w, int flags, const int64_t* mask,
                        const void* ws, double* sums, float* out, float* bwd_scale, cudaStream_t st,
                        const mafed_comm* comm = nullptr, int comm_what = 0, double* counts_out = nullptr) {
  EpiParams e;
  memset(&e, 0, sizeof(e));
  e.comm = comm_dev(comm);
  if (e.comm.world > 1) {
    const int L2 = 2 * sh.n_layers;
    if (comm_what == (MAFED_COMM_SUMS | MAFED_COMM_COUNTS)) { e.a.comm_first = 0; e.a.comm_count = L2 + 2; }
    else if (comm_what == MAFED_COMM_SUMS) { e.a.comm_first = 0; e.a.comm_count = L2; }
    else if (comm_what == MAFED_COMM_COUNTS) { e.a.comm_first = L2; e.a.comm_count = 2; }
  }
  e.a.ws = reinterpret_cast<const float*>(ws);
  e.a.mask = mask;
  e.a.sums = sums;
  e.a.counts_out = counts_out;
  e.a.out = out;
  e.a.bwd_scale = bwd_scale;
  e.a.n_mask = mask_entries(sh);
  e.a.n_vis_rows = vis_rows(sh);
  e.a.n_layers = sh.n_layers;
  e.a.D = sh.D;
  e.a.loss_kind = sh.loss_kind;
  e.a.flags = flags;
  if (w != nullptr) e.w = *w;
  launch_pdl(k_epilogue, 1, kEpiThreads, 0, st, e);
  return (int)cudaPeekAtLastError();
}

// What mafed_distill_step wants produced besides the gradients; with it the fused kernel may run the loss
// algebra (and the modality masks) itself.
struct StepTail {
  float* out;
  double* sums;
  int64_t* lang_mask;
  int64_t* image_mask;
};

int fused_impl(const mafed_shape_t* shape, const void* const* student_ptrs, const void* const* teacher_ptrs,
               void* const* grad_ptrs, const int64_t* attn_mask, const mafed_weights_t* weights, float* bwd_scale,
               float assumed_grad_out, void* ws, mafed_comm* comm, void* stream, const StepTail* tail, bool* folded) {
  if (!grad_ptrs || !bwd_scale || !ws) return MAFED_E_ARG;
  if (comm != nullptr && comm->world > 1 && (!weights || comm->world > kCommMaxRanks)) return MAFED_E_ARG;
  PathParams p;
  int rc = fill_params(shape, student_ptrs, teacher_ptrs, grad_ptrs, attn_mask, p);
  if (rc) return rc;
  p.bwd_scale = bwd_scale;
  p.fixed_gout = assumed_grad_out;
  p.ws = reinterpret_cast<float*>(ws);
  const bool sharded = comm != nullptr && comm->world > 1;
  const bool tma = uses_tma(*shape, p, kPassFused);
  const bool want_tail = tail != nullptr && weights != nullptr && tma && !g_tune[kTuneNoTail].load();
  const bool want_inline = weights != nullptr && tma && mask_entries(*shape) <= 16384 && !g_tune[kTuneNoInlineScale].load();
  // The arrival counter: needed for the tail and, in a sharded step, for the in-kernel counts exchange (the last
  // CTA advances the device-side epoch).
  unsigned int* done = nullptr;
  if (want_tail || (want_inline && sharded)) done = next_tail_counter((cudaStream_t)stream);
  p.tail_done = done;
  if (done != nullptr && want_tail) {
    // the last CTA of the TMA kernel reduces the partials and forms the losses (and exchanges the sums with the
    // peers); all CTAs write the modality masks: no epilogue or mask launch
    p.tail_flags = kEpiReduce | kEpiLosses;
    p.tail_out = tail->out;
    p.tail_sums = tail->sums;
    p.lang_mask_out = tail->lang_mask;
    p.image_mask_out = tail->image_mask;
    p.tail_comm = sharded ? 1 : 0;
    if (folded != nullptr) *folded = true;
  }
  if (weights != nullptr) {
    p.loss_kind = shape->loss_kind;
    p.n_mask = mask_entries(*shape);
    p.n_vis_rows = vis_rows(*shape);
    p.w = *weights;
    if (sharded) p.comm = comm_dev(comm);
    // the scale table is this call's business.  Small masks: every CTA of the TMA kernel derives it itself while
    // its first tiles are in flight (the whole step is ONE launch; in a sharded step the counts exchange rides
    // inside the kernel); otherwise one prologue launch, which leaves the counts in the ws header for the tail.
    if (want_inline && (!sharded || done != nullptr)) {
      p.inline_scale = 1;
      p.bwd_scale_out = bwd_scale;
      p.comm_counts = sharded ? 1 : 0;
    } else {
      double* counts = reinterpret_cast<double*>(p.ws + kWsCountsAt);
      p.tail_counts_in = counts;
      rc = launch_scalar_stage(*shape, weights, MAFED_STAGE_COUNTS | MAFED_STAGE_SCALE, attn_mask, nullptr, nullptr,
                               nullptr, bwd_scale, (cudaStream_t)stream, comm, MAFED_COMM_COUNTS, counts);
      if (rc) return rc;
    }
  }
  return dispatch(*shape, p, kPassFused, (cudaStream_t)stream);
}

}  // namespace
}  // namespace mafed

using namespace mafed;

extern "C" {

int mafed_distill_abi_version(void) { return MAFED_ABI_VERSION; }

const char* mafed_distill_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case MAFED_E_ARG: return "invalid argument (null pointer, non-positive size or too many layers)";
    case MAFED_E_DTYPE: return "unsupported dtype or loss kind";
    case MAFED_E_ALIGN: return "pointer not aligned to its element size";
    case MAFED_E_NODEVICE: return "no usable CUDA device";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

size_t mafed_distill_ws_bytes(int n_layers) {
  if (n_layers < 1) n_layers = 1;
  if (n_layers > kMaxLayers) n_layers = kMaxLayers;
  return sizeof(float) * ((size_t)kWsHeaderFloats + (size_t)kMaxPartials * n_layers * 2);
}

int mafed_distill_sums_len(int n_layers) { return 2 * n_layers + 2; }
int mafed_distill_out_len(int n_layers) { return 1 + 3 * n_layers; }

int mafed_distill_fwd(const mafed_shape_t* shape, const void* const* student_ptrs, const void* const* teacher_ptrs,
                      const int64_t* attn_mask, void* ws, void* stream) {
  PathParams p;
  int rc = fill_params(shape, student_ptrs, teacher_ptrs, nullptr, attn_mask, p);
  if (rc) return rc;
  if (!ws) return MAFED_E_ARG;
  p.ws = reinterpret_cast<float*>(ws);
  return dispatch(*shape, p, kPassFwd, (cudaStream_t)stream);
}

int mafed_distill_token_norm_sums(const mafed_shape_t* shape, const void* const* tensor_ptrs, const int64_t* attn_mask,
                                  void* ws, void* stream) {
  PathParams p;
  int rc = fill_params(shape, tensor_ptrs, tensor_ptrs, nullptr, attn_mask, p);
  if (rc) return rc;
  if (!ws) return MAFED_E_ARG;
  p.ws = reinterpret_cast<float*>(ws);
  p.single_input = 1;
  return dispatch(*shape, p, kPassFwd, (cudaStream_t)stream, kLossL2Norm);
}

int mafed_distill_scalar_stage(const mafed_shape_t* shape, const mafed_weights_t* weights, int flags,
                               const int64_t* attn_mask, const void* ws, double* sums, float* out, float* bwd_scale,
                               void* stream) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if ((flags & MAFED_STAGE_REDUCE) && !ws) return MAFED_E_ARG;
  if ((flags & MAFED_STAGE_COUNTS) && needs_mask(*shape) && !attn_mask) return MAFED_E_ARG;
  if ((flags & (MAFED_STAGE_LOSSES | MAFED_STAGE_SCALE)) && !weights) return MAFED_E_ARG;
  if ((flags & MAFED_STAGE_LOSSES) && !out) return MAFED_E_ARG;
  if ((flags & MAFED_STAGE_SCALE) && !bwd_scale) return MAFED_E_ARG;
  const bool reads_sums = ((flags & MAFED_STAGE_LOSSES) && !(flags & MAFED_STAGE_REDUCE)) ||
                          ((flags & (MAFED_STAGE_LOSSES | MAFED_STAGE_SCALE)) && !(flags & MAFED_STAGE_COUNTS));
  const bool only_writes = (flags & (MAFED_STAGE_LOSSES | MAFED_STAGE_SCALE)) == 0;
  if ((reads_sums || only_writes) && !sums) return MAFED_E_ARG;
  return launch_scalar_stage(*shape, weights, flags, attn_mask, ws, sums, out, bwd_scale, (cudaStream_t)stream);
}

int mafed_distill_scalar_stage_comm(const mafed_shape_t* shape, const mafed_weights_t* weights, int flags,
                                    const int64_t* attn_mask, const void* ws, double* sums, float* out,
                                    float* bwd_scale, mafed_comm_t* comm, int comm_what, void* stream) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if ((flags & MAFED_STAGE_REDUCE) && !ws) return MAFED_E_ARG;
  if ((flags & MAFED_STAGE_COUNTS) && needs_mask(*shape) && !attn_mask) return MAFED_E_ARG;
  if ((flags & (MAFED_STAGE_LOSSES | MAFED_STAGE_SCALE)) && !weights) return MAFED_E_ARG;
  if ((flags & MAFED_STAGE_LOSSES) && !out) return MAFED_E_ARG;
  if ((flags & MAFED_STAGE_SCALE) && !bwd_scale) return MAFED_E_ARG;
  if (!sums) return MAFED_E_ARG;  // the distributed sequence always carries the sums vector between stages
  if (comm != nullptr && (comm->world > kCommMaxRanks || comm_what < 0 || comm_what > 3)) return MAFED_E_ARG;
  return launch_scalar_stage(*shape, weights, flags, attn_mask, ws, sums, out, bwd_scale, (cudaStream_t)stream, comm,
                             comm_what);
}

int mafed_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int mafed_comm_create(int world, int rank, void* ipc_handle_out, mafed_comm_t** out) {
  if (world < 1 || world > kCommMaxRanks || rank < 0 || rank >= world || !ipc_handle_out || !out) return MAFED_E_ARG;
  mafed_comm* c = new mafed_comm();
  c->world = world;
  c->rank = rank;
  cudaError_t e = cudaMalloc(&c->local, kCommMailboxBytes);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, kCommMailboxBytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->local);
  if (e != cudaSuccess) {
    if (c->local) cudaFree(c->local);
    delete c;
    return (int)e;
  }
  memcpy(ipc_handle_out, &h, sizeof(h));
  c->peers[rank] = c->local;
  int dev = 0, khz = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev) == cudaSuccess && khz > 0)
    c->cycles_per_second = 1e3 * (double)khz;
  double seconds = kCommDefaultTimeoutS;
  if (const char* env = getenv("MAFED_B200_COMM_TIMEOUT_S")) {
    const double v = atof(env);
    if (v > 0.0) seconds = v;
  }
  c->timeout_cycles = (long long)(seconds * c->cycles_per_second);
  *out = c;
  return 0;
}

int mafed_comm_set_timeout(mafed_comm_t* c, double seconds) {
  if (!c || !(seconds > 0.0)) return MAFED_E_ARG;
  c->timeout_cycles = (long long)(seconds * c->cycles_per_second);
  return 0;
}

int mafed_comm_connect(mafed_comm_t* c, const void* all_handles) {
  if (!c) return MAFED_E_ARG;
  if (!all_handles) {
    // diagnostic loop-back: every "peer" is this rank's own mailbox.  Rank r only ever fills slot r, so the other
    // ranks' slots stay empty and every exchange runs into the spin bound -- the way to test that bound on one GPU.
    for (int r = 0; r < c->world; ++r) c->peers[r] = c->local;
    c->loopback = true;
    return 0;
  }
  const char* hs = reinterpret_cast<const char*>(all_handles);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + (size_t)r * sizeof(h), sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(&c->peers[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

int mafed_comm_status(mafed_comm_t* c, int* status_out) {
  if (!c || !status_out) return MAFED_E_ARG;
  const char* tail = reinterpret_cast<const char*>(c->local) + kCommDataBytes;
  return (int)cudaMemcpy(status_out, tail + 16, sizeof(int), cudaMemcpyDeviceToHost);
}

int mafed_comm_trace(mafed_comm_t* c, unsigned long long* out4) {
  if (!c || !out4) return MAFED_E_ARG;
  const char* tail = reinterpret_cast<const char*>(c->local) + kCommDataBytes;
  return (int)cudaMemcpy(out4, tail + 32, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
}

int mafed_comm_destroy(mafed_comm_t* c) {
  if (!c) return 0;
  for (int r = 0; r < c->world; ++r)
    if (!c->loopback && r != c->rank && c->peers[r]) cudaIpcCloseMemHandle(c->peers[r]);
  if (c->local) cudaFree(c->local);
  delete c;
  return 0;
}

int mafed_distill_reduce(const mafed_shape_t* shape, const int64_t* attn_mask, const void* ws, double* sums,
                         void* stream) {
  return mafed_distill_scalar_stage(shape, nullptr, MAFED_STAGE_REDUCE | MAFED_STAGE_COUNTS, attn_mask, ws, sums,
                                    nullptr, nullptr, stream);
}

int mafed_distill_finalize(const mafed_shape_t* shape, const mafed_weights_t* weights, const double* sums, float* out,
                           float* bwd_scale, void* stream) {
  return mafed_distill_scalar_stage(shape, weights, MAFED_STAGE_LOSSES | (bwd_scale ? MAFED_STAGE_SCALE : 0), nullptr,
                                    nullptr, const_cast<double*>(sums), out, bwd_scale, stream);
}

int mafed_distill_epilogue(const mafed_shape_t* shape, const mafed_weights_t* weights, const int64_t* attn_mask,
                           const void* ws, double* sums, float* out, float* bwd_scale, void* stream) {
  const int flags = MAFED_STAGE_REDUCE | MAFED_STAGE_COUNTS | MAFED_STAGE_LOSSES | (bwd_scale ? MAFED_STAGE_SCALE : 0);
  return mafed_distill_scalar_stage(shape, weights, flags, attn_mask, ws, sums, out, bwd_scale, stream);
}

int mafed_distill_prologue(const mafed_shape_t* shape, const mafed_weights_t* weights, const int64_t* attn_mask,
                           double* global_counts, double* sums, float* bwd_scale, void* stream) {
  if (global_counts != nullptr)
    return mafed_distill_scalar_stage(shape, weights, MAFED_STAGE_SCALE, nullptr, nullptr, global_counts, nullptr,
                                      bwd_scale, stream);
  return mafed_distill_scalar_stage(shape, weights, MAFED_STAGE_COUNTS | MAFED_STAGE_SCALE, attn_mask, nullptr, sums,
                                    nullptr, bwd_scale, stream);
}

int mafed_distill_bwd(const mafed_shape_t* shape, const void* const* student_ptrs, const void* const* teacher_ptrs,
                      void* const* grad_ptrs, const int64_t* attn_mask, const float* bwd_scale, const float* grad_out,
                      const float* skip_if_equals, void* stream) {
  if (!grad_ptrs || !bwd_scale) return MAFED_E_ARG;
  PathParams p;
  int rc = fill_params(shape, student_ptrs, teacher_ptrs, grad_ptrs, attn_mask, p);
  if (rc) return rc;
  p.bwd_scale = bwd_scale;
  p.grad_out = grad_out;
  if (skip_if_equals != nullptr) {
    p.skip_if_gout_equals = 1;
    p.fixed_gout = *skip_if_equals;
  }
  p.reverse = g_tune[kTuneBwdForward].load() ? 0 : 1;
  return dispatch(*shape, p, kPassBwd, (cudaStream_t)stream);
}

int mafed_distill_fused(const mafed_shape_t* shape, const void* const* student_ptrs, const void* const* teacher_ptrs,
                        void* const* grad_ptrs, const int64_t* attn_mask, const mafed_weights_t* weights, float* bwd_scale,
                        float assumed_grad_out, void* ws, void* stream) {
  return fused_impl(shape, student_ptrs, teacher_ptrs, grad_ptrs, attn_mask, weights, bwd_scale, assumed_grad_out, ws,
                    nullptr, stream, nullptr, nullptr);
}

int mafed_distill_fused_comm(const mafed_shape_t* shape, const void* const* student_ptrs,
                             const void* const* teacher_ptrs, void* const* grad_ptrs, const int64_t* attn_mask,
                             const mafed_weights_t* weights, float* bwd_scale, float assumed_grad_out, void* ws,
                             mafed_comm_t* comm, void* stream) {
  return fused_impl(shape, student_ptrs, teacher_ptrs, grad_ptrs, attn_mask, weights, bwd_scale, assumed_grad_out, ws,
                    comm, stream, nullptr, nullptr);
}

int mafed_distill_step(const mafed_shape_t* shape, const void* const* student_ptrs, const void* const* teacher_ptrs,
                       void* const* grad_ptrs, const int64_t* attn_mask, const mafed_weights_t* weights,
                       float assumed_grad_out, void* ws, float* out, float* bwd_scale, double* sums,
                       int64_t* lang_mask, int64_t* image_mask, mafed_comm_t* comm, void* stream) {
  if (!weights || !out) return MAFED_E_ARG;
  if ((lang_mask == nullptr) != (image_mask == nullptr)) return MAFED_E_ARG;
  const bool sharded = comm != nullptr && comm->world > 1;
  if (sharded && !sums) return MAFED_E_ARG;
  int rc = check_shape(shape);
  if (rc) return rc;
  if (lang_mask != nullptr && shape->cls) return MAFED_E_ARG;
  StepTail tail = {out, sums, lang_mask, image_mask};
  bool folded = false;
  rc = fused_impl(shape, student_ptrs, teacher_ptrs, grad_ptrs, attn_mask, weights, bwd_scale, assumed_grad_out, ws,
                  comm, stream, &tail, &folded);
  if (rc || folded) return rc;
  // not a single-launch shape (rows > 32 KB, unaligned tensors, a mask of more than 16 Ki entries): same step as
  // separate launches
  if (lang_mask != nullptr) {
    rc = mafed_distill_modality_masks(shape, attn_mask, lang_mask, image_mask, stream);
    if (rc) return rc;
  }
  const int flags = MAFED_STAGE_REDUCE | MAFED_STAGE_COUNTS | MAFED_STAGE_LOSSES;
  return launch_scalar_stage(*shape, weights, flags, attn_mask, ws, sums, out, nullptr, (cudaStream_t)stream,
                             sharded ? comm : nullptr, MAFED_COMM_SUMS | MAFED_COMM_COUNTS);
}

int mafed_distill_fwd_step(const mafed_shape_t* shape, const void* const* student_ptrs,
                           const void* const* teacher_ptrs, const int64_t* attn_mask, const mafed_weights_t* weights,
                           void* ws, float* out, float* bwd_scale, double* sums, mafed_comm_t* comm, void* stream) {
  if (!weights || !out || !ws) return MAFED_E_ARG;
  const bool sharded = comm != nullptr && comm->world > 1;
  if (sharded && (!sums || comm->world > kCommMaxRanks)) return MAFED_E_ARG;
  PathParams p;
  int rc = fill_params(shape, student_ptrs, teacher_ptrs, nullptr, attn_mask, p);
  if (rc) return rc;
  p.ws = reinterpret_cast<float*>(ws);
  const int flags = MAFED_STAGE_REDUCE | MAFED_STAGE_COUNTS | MAFED_STAGE_LOSSES | (bwd_scale ? MAFED_STAGE_SCALE : 0);
  unsigned int* done = nullptr;
  if (!g_tune[kTuneNoTail].load() && uses_tma(*shape, p, kPassFwd)) done = next_tail_counter((cudaStream_t)stream);
  if (done != nullptr) {
    p.tail_flags = flags;
    p.tail_done = done;
    p.tail_out = out;
    p.tail_sums = sums;
    p.tail_bwd_scale = bwd_scale;
    p.loss_kind = shape->loss_kind;
    p.n_mask = mask_entries(*shape);
    p.n_vis_rows = vis_rows(*shape);
    p.w = *weights;
    if (sharded) {   // sums + counts exchange inside the tail
      p.comm = comm_dev(comm);
      p.tail_comm = 1;
    }
    return dispatch(*shape, p, kPassFwd, (cudaStream_t)stream);
  }
  rc = dispatch(*shape, p, kPassFwd, (cudaStream_t)stream);
  if (rc) return rc;
  return launch_scalar_stage(*shape, weights, flags, attn_mask, ws, sums, out, bwd_scale, (cudaStream_t)stream,
                             sharded ? comm : nullptr, MAFED_COMM_SUMS | MAFED_COMM_COUNTS);
}

int mafed_distill_modality_masks(const mafed_shape_t* shape, const int64_t* attn_mask, int64_t* lang_mask,
                                 int64_t* image_mask, void* stream) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!lang_mask || !image_mask || (needs_mask(*shape) && !attn_mask)) return MAFED_E_ARG;
  const long long n = (long long)shape->B * shape->T;
  const unsigned grid = (unsigned)clamp_grid((n + 255) / 256, 1 << 20);
  launch_pdl(k_modality_masks, grid > 1184 ? 1184u : grid, 256, 0, (cudaStream_t)stream, attn_mask, lang_mask,
             image_mask, n, (int)shape->T, (int)shape->n_vis);
  return (int)cudaPeekAtLastError();
}

// ---------------------------------------------------------------- host-buffer step
size_t mafed_host_step_device_bytes(const mafed_shape_t* shape) {
  if (check_shape(shape)) return 0;
  const size_t layer = (size_t)shape->B * shape->T * shape->D * elem_size(shape->dtype);
  const size_t mask = needs_mask(*shape) ? (size_t)shape->B * (shape->T - shape->n_vis) * sizeof(int64_t) : 0;
  return 3 * layer * shape->n_layers + mask + (size_t)shape->n_layers * (mafed_distill_ws_bytes(1) + 64) + 4096;
}

int mafed_host_step_create(const mafed_shape_t* shape, mafed_host_step_t** out) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!out || shape->cls) return MAFED_E_ARG;
  if (!device_info().ok) return MAFED_E_NODEVICE;
  mafed_host_step* h = new mafed_host_step();
  h->shape = *shape;
  const int L = shape->n_layers;
  h->layer_bytes = (((size_t)shape->B * shape->T * shape->D * elem_size(shape->dtype)) + 255) & ~(size_t)255;
  h->mask_bytes = ((needs_mask(*shape) ? (size_t)shape->B * (shape->T - shape->n_vis) * sizeof(int64_t) : 0) + 255) & ~(size_t)255;
  h->ws_bytes = (mafed_distill_ws_bytes(1) + 255) & ~(size_t)255;
  const size_t total = 3 * h->layer_bytes * L + h->mask_bytes + (size_t)L * h->ws_bytes + (size_t)L * 6 * sizeof(float) + 1024;
  cudaError_t e = cudaMalloc(&h->d_pool, total);
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_out, sizeof(float) * 4 * L);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_run, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    mafed_host_step_destroy(h);
    return (int)e;
  }
  char* p = h->d_pool;
  for (int l = 0; l < L; ++l) { h->d_s.push_back(p); p += h->layer_bytes; }
  for (int l = 0; l < L; ++l) { h->d_t.push_back(p); p += h->layer_bytes; }
  for (int l = 0; l < L; ++l) { h->d_g.push_back(p); p += h->layer_bytes; }
  h->d_mask = reinterpret_cast<int64_t*>(p); p += h->mask_bytes;
  h->d_ws = p; p += (size_t)L * h->ws_bytes;
  h->d_out = reinterpret_cast<float*>(p); p += sizeof(float) * 4 * L;
  h->d_scale = reinterpret_cast<float*>(p);
  h->ev_in.resize(L);
  h->ev_run.resize(L);
  for (int l = 0; l < L; ++l) {
    cudaEventCreateWithFlags(&h->ev_in[l], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_run[l], cudaEventDisableTiming);
  }
  *out = h;
  return 0;
}

int mafed_host_step_run(mafed_host_step_t* h, const mafed_weights_t* weights, const void* const* h_student,
                        const void* const* h_teacher, void* const* h_grad, const int64_t* h_mask, float grad_out,
                        float* h_out) {
  if (!h || !weights || !h_student || !h_teacher || !h_grad || !h_out) return MAFED_E_ARG;
  const mafed_shape_t& sh = h->shape;
  const int L = sh.n_layers;
  if (needs_mask(sh) && !h_mask) return MAFED_E_ARG;
  const size_t layer = (size_t)sh.B * sh.T * sh.D * elem_size(sh.dtype);
  const size_t mask = needs_mask(sh) ? (size_t)sh.B * (sh.T - sh.n_vis) * sizeof(int64_t) : 0;
  cudaError_t e = cudaSuccess;
  if (mask) e = cudaMemcpyAsync(h->d_mask, h_mask, mask, cudaMemcpyHostToDevice, h->s_in);
  for (int l = 0; l < L && e == cudaSuccess; ++l) {
    e = cudaMemcpyAsync(h->d_s[l], h_student[l], layer, cudaMemcpyHostToDevice, h->s_in);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_t[l], h_teacher[l], layer, cudaMemcpyHostToDevice, h->s_in);
    if (e == cudaSuccess) e = cudaEventRecord(h->ev_in[l], h->s_in);
  }
  if (e != cudaSuccess) return (int)e;
  mafed_shape_t one = sh;
  one.n_layers = 1;
  for (int l = 0; l < L; ++l) {
    mafed_weights_t w;
    memset(&w, 0, sizeof(w));
    w.modality_kind = weights->modality_kind;
    w.distill_coeff = weights->distill_coeff;
    w.layer_coeff[0] = weights->layer_coeff[l];
    w.lang_weight[0] = weights->lang_weight[l];
    cudaStreamWaitEvent(h->s_run, h->ev_in[l], 0);
    const void* sp[1] = {h->d_s[l]};
    const void* tp[1] = {h->d_t[l]};
    void* gp[1] = {h->d_g[l]};
    char* ws = h->d_ws + (size_t)l * h->ws_bytes;
    int rc = mafed_distill_step(&one, sp, tp, gp, h->d_mask, &w, grad_out, ws, h->d_out + 4 * l, h->d_scale + 2 * l,
                                nullptr, nullptr, nullptr, nullptr, h->s_run);
    if (rc) return rc;
    cudaEventRecord(h->ev_run[l], h->s_run);
    cudaStreamWaitEvent(h->s_out, h->ev_run[l], 0);
    e = cudaMemcpyAsync(h_grad[l], h->d_g[l], layer, cudaMemcpyDeviceToHost, h->s_out);
    if (e != cudaSuccess) return (int)e;
  }
  e = cudaMemcpyAsync(h->h_out, h->d_out, sizeof(float) * 4 * L, cudaMemcpyDeviceToHost, h->s_out);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_out);
  if (e != cudaSuccess) return (int)e;
  // out[1+3L] layout of the device path: total, layer losses, (text, vision) losses
  double total = 0.0;
  for (int l = 0; l < L; ++l) {
    total += (double)h->h_out[4 * l];
    h_out[1 + l] = h->h_out[4 * l + 1];
    h_out[1 + L + 2 * l] = h->h_out[4 * l + 2];
    h_out[1 + L + 2 * l + 1] = h->h_out[4 * l + 3];
  }
  h_out[0] = (float)total;
  return 0;
}

int mafed_host_step_destroy(mafed_host_step_t* h) {
  if (!h) return 0;
  for (cudaEvent_t ev : h->ev_in) cudaEventDestroy(ev);
  for (cudaEvent_t ev : h->ev_run) cudaEventDestroy(ev);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_run) cudaStreamDestroy(h->s_run);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  if (h->h_out) cudaFreeHost(h->h_out);
  if (h->d_pool) cudaFree(h->d_pool);
  delete h;
  return 0;
}

int mafed_host_register(void* ptr, size_t bytes) { return (int)cudaHostRegister(ptr, bytes, cudaHostRegisterDefault); }
int mafed_host_unregister(void* ptr) { return (int)cudaHostUnregister(ptr); }

int mafed_distill_set_variant(int variant) {
  if (variant < 0 || variant > 2) return MAFED_E_ARG;
  g_variant.store(variant);
  return 0;
}

int mafed_distill_set_tuning(int key, int value) {
  if (key < 0 || key >= 24) return MAFED_E_ARG;
  g_tune[key].store(value);
  return 0;
}

}  // extern "C"
