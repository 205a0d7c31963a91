// C ABI (include/mafed_distill.h) over the sm_100a kernels.  No torch types, no allocation, no host
// synchronisation: every entry point validates its arguments, picks a launch geometry and enqueues
// kernels on the caller's stream.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "distill_common.cuh"
#include "distill_epilogue.cuh"
#include "distill_ldg.cuh"
#include "distill_tma.cuh"

namespace mafed {
namespace {

std::atomic<int> g_variant{0};            // 0 default, 1 ldg, 2 tma
std::atomic<int> g_tune[8] = {};          // experiment knobs, see mafed_distill_set_tuning

enum TuneKey { kTuneTmaStages = 0, kTuneTmaRows = 1, kTuneTmaWarps = 2, kTuneLdgBlocksPerSm = 3, kTuneBwdReverse = 4,
               kTuneGridMul = 5 };

struct DeviceInfo {
  int sm_count = 0;
  int smem_optin = 0;
  int cc_major = 0;
  bool ok = false;
};

const DeviceInfo& device_info() {
  static DeviceInfo info[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) {
    static DeviceInfo none;
    return none;
  }
  DeviceInfo& d = info[dev];
  if (!d.ok) {
    cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
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
  p.n_chunks = 0;
  p.reverse = 0;
}

bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---------------------------------------------------------------- launch helpers
template <typename K>
int blocks_per_sm(K kernel, int threads, size_t dyn_smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, dyn_smem) != cudaSuccess) n = 1;
  return n < 1 ? 1 : n;
}

template <typename T, int CPL, int RPI, int LOSS>
int launch_ldg(const PathParams& p, bool backward, cudaStream_t st) {
  static int occ_f = 0, occ_b = 0;
  const DeviceInfo& dv = device_info();
  int& occ = backward ? occ_b : occ_f;
  if (occ == 0)
    occ = backward ? blocks_per_sm(k_bwd_ldg<T, CPL, RPI, LOSS>, kLdgThreads, 0)
                   : blocks_per_sm(k_fwd_ldg<T, CPL, RPI, LOSS>, kLdgThreads, 0);
  int per_sm = occ;
  const int cap = g_tune[kTuneLdgBlocksPerSm].load();
  if (cap > 0 && cap < per_sm) per_sm = cap;
  const long long rows_per_iter = (long long)kLdgWarps * RPI;
  const long long total = ((p.n_rows + rows_per_iter - 1) / rows_per_iter) * p.n_layers;
  long long grid = (long long)dv.sm_count * per_sm;
  if (grid > total) grid = total;
  if (grid > kMaxPartials) grid = kMaxPartials;
  if (grid < 1) grid = 1;
  if (backward) k_bwd_ldg<T, CPL, RPI, LOSS><<<(unsigned)grid, kLdgThreads, 0, st>>>(p);
  else k_fwd_ldg<T, CPL, RPI, LOSS><<<(unsigned)grid, kLdgThreads, 0, st>>>(p);
  return (int)cudaPeekAtLastError();
}

template <typename T, int LOSS>
int dispatch_ldg(const PathParams& p, bool backward, cudaStream_t st) {
  const int cpl = (p.n_chunks + 31) / 32;
  if (cpl <= 1) return launch_ldg<T, 1, 4, LOSS>(p, backward, st);
  if (cpl <= 2) return launch_ldg<T, 2, 4, LOSS>(p, backward, st);
  if (cpl <= 3) return launch_ldg<T, 3, 2, LOSS>(p, backward, st);
  if (cpl <= 4) return launch_ldg<T, 4, 2, LOSS>(p, backward, st);
  if (cpl <= 6) return launch_ldg<T, 6, 1, LOSS>(p, backward, st);
  return launch_ldg<T, 8, 1, LOSS>(p, backward, st);  // multi-pass for rows longer than 4 KB
}

template <typename T, int LOSS>
int launch_generic(const PathParams& p, bool backward, cudaStream_t st) {
  const DeviceInfo& dv = device_info();
  const long long total = ((p.n_rows + kLdgWarps - 1) / kLdgWarps) * p.n_layers;
  long long grid = (long long)dv.sm_count * 4;
  if (grid > total) grid = total;
  if (grid > kMaxPartials) grid = kMaxPartials;
  if (backward) k_bwd_generic<T, LOSS><<<(unsigned)grid, kLdgThreads, 0, st>>>(p);
  else k_fwd_generic<T, LOSS><<<(unsigned)grid, kLdgThreads, 0, st>>>(p);
  return (int)cudaPeekAtLastError();
}

bool tma_geometry(const PathParams& p, TmaGeom& geo) {
  const DeviceInfo& dv = device_info();
  const long long row_bytes = (long long)p.n_chunks * 16;
  if (row_bytes > 32768) return false;
  const long long budget = (long long)dv.smem_optin - 16 * 1024;  // static smem + slack
  int rows = g_tune[kTuneTmaRows].load();
  if (rows <= 0) {
    rows = (int)(65536 / (2 * row_bytes));
    if (rows >= 8) rows &= ~7;
  }
  if (rows > kTmaMaxRows) rows = kTmaMaxRows;
  if (rows < 1) rows = 1;
  geo.rows = rows;
  geo.stage_bytes = (int)(2 * rows * row_bytes);
  int stages = g_tune[kTuneTmaStages].load();
  if (stages <= 0) stages = 4;
  while (stages > 1 && (long long)stages * geo.stage_bytes > budget) --stages;
  if (stages > kTmaMaxStages) stages = kTmaMaxStages;
  if ((long long)stages * geo.stage_bytes > budget) return false;
  geo.stages = stages;
  return true;
}

template <typename T, int LOSS, int NCW>
int launch_tma(const PathParams& p, const TmaGeom& geo, bool backward, cudaStream_t st) {
  static bool attr_f = false, attr_b = false;
  const DeviceInfo& dv = device_info();
  const size_t dyn = (size_t)geo.stages * geo.stage_bytes;
  bool& attr = backward ? attr_b : attr_f;
  if (!attr) {
    cudaError_t e = backward
        ? cudaFuncSetAttribute(k_bwd_tma<T, LOSS, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, dv.smem_optin - 16 * 1024)
        : cudaFuncSetAttribute(k_fwd_tma<T, LOSS, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, dv.smem_optin - 16 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const long long total = ((p.n_rows + geo.rows - 1) / geo.rows) * p.n_layers;
  int mul = g_tune[kTuneGridMul].load();
  if (mul <= 0) mul = 1;
  long long grid = (long long)dv.sm_count * mul;
  if (grid > total) grid = total;
  if (grid > kMaxPartials) grid = kMaxPartials;
  constexpr int threads = (NCW + 1) * 32;
  if (backward) k_bwd_tma<T, LOSS, NCW><<<(unsigned)grid, threads, dyn, st>>>(p, geo);
  else k_fwd_tma<T, LOSS, NCW><<<(unsigned)grid, threads, dyn, st>>>(p, geo);
  return (int)cudaPeekAtLastError();
}

template <typename T, int LOSS>
int dispatch_typed(PathParams& p, bool vector_ok, bool backward, cudaStream_t st) {
  if (!vector_ok) return launch_generic<T, LOSS>(p, backward, st);
  int variant = g_variant.load();
  if (variant == 0) variant = 2;
  if (variant == 2) {
    TmaGeom geo;
    if (tma_geometry(p, geo)) {
      if (g_tune[kTuneTmaWarps].load() == 16) return launch_tma<T, LOSS, 16>(p, geo, backward, st);
      return launch_tma<T, LOSS, 8>(p, geo, backward, st);
    }
  }
  return dispatch_ldg<T, LOSS>(p, backward, st);
}

template <typename T>
int dispatch_loss(PathParams& p, int loss, bool vector_ok, bool backward, cudaStream_t st) {
  if (loss == MAFED_LOSS_MSE) return dispatch_typed<T, MAFED_LOSS_MSE>(p, vector_ok, backward, st);
  return dispatch_typed<T, MAFED_LOSS_COSINE>(p, vector_ok, backward, st);
}

int dispatch(const mafed_shape_t& sh, PathParams& p, bool backward, cudaStream_t st) {
  // vector path: rows are whole, 16-byte aligned chunks
  const size_t es = elem_size(sh.dtype);
  bool vector_ok = ((size_t)sh.D * es) % 16 == 0 && ((size_t)p.row_stride * es) % 16 == 0;
  for (int l = 0; l < sh.n_layers && vector_ok; ++l) {
    vector_ok = aligned_to(p.s[l], 16) && aligned_to(p.t[l], 16) && (!backward || aligned_to(p.g[l], 16));
  }
  p.n_chunks = vector_ok ? (int)((size_t)sh.D * es / 16) : 0;
  switch (sh.dtype) {
    case MAFED_F32: return dispatch_loss<float>(p, sh.loss_kind, vector_ok, backward, st);
    case MAFED_BF16: return dispatch_loss<__nv_bfloat16>(p, sh.loss_kind, vector_ok, backward, st);
    default: return dispatch_loss<__half>(p, sh.loss_kind, vector_ok, backward, st);
  }
}

int launch_epilogue(const mafed_shape_t& sh, const mafed_weights_t* w, const int64_t* mask, const void* ws,
                    double* sums, float* out, float* bwd_scale, bool do_reduce, bool do_finalize, cudaStream_t st) {
  EpiParams e;
  memset(&e, 0, sizeof(e));
  e.ws = reinterpret_cast<const float*>(ws);
  e.mask = mask;
  e.sums = sums;
  e.out = out;
  e.bwd_scale = bwd_scale;
  e.n_mask = sh.cls ? 0 : (long long)sh.B * (sh.T - sh.n_vis);
  e.n_vis_rows = sh.cls ? (double)sh.B : (double)sh.B * (double)sh.n_vis;
  e.n_layers = sh.n_layers;
  e.D = sh.D;
  e.loss_kind = sh.loss_kind;
  e.do_reduce = do_reduce;
  e.do_finalize = do_finalize;
  if (w != nullptr) e.w = *w;
  k_epilogue<<<1, kEpiThreads, 0, st>>>(e);
  return (int)cudaPeekAtLastError();
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
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!student_ptrs || !teacher_ptrs || !ws || (!shape->cls && shape->T > shape->n_vis && !attn_mask)) return MAFED_E_ARG;
  if (!device_info().ok) return MAFED_E_NODEVICE;
  PathParams p;
  memset(&p, 0, sizeof(p));
  fill_geometry(*shape, p);
  const size_t es = elem_size(shape->dtype);
  for (int l = 0; l < shape->n_layers; ++l) {
    if (!student_ptrs[l] || !teacher_ptrs[l]) return MAFED_E_ARG;
    if (!aligned_to(student_ptrs[l], es) || !aligned_to(teacher_ptrs[l], es)) return MAFED_E_ALIGN;
    p.s[l] = student_ptrs[l];
    p.t[l] = teacher_ptrs[l];
  }
  p.mask = attn_mask;
  p.ws = reinterpret_cast<float*>(ws);
  return dispatch(*shape, p, false, (cudaStream_t)stream);
}

int mafed_distill_reduce(const mafed_shape_t* shape, const int64_t* attn_mask, const void* ws, double* sums,
                         void* stream) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!ws || !sums) return MAFED_E_ARG;
  return launch_epilogue(*shape, nullptr, attn_mask, ws, sums, nullptr, nullptr, true, false, (cudaStream_t)stream);
}

int mafed_distill_finalize(const mafed_shape_t* shape, const mafed_weights_t* weights, const double* sums, float* out,
                           float* bwd_scale, void* stream) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!weights || !sums || !out || !bwd_scale) return MAFED_E_ARG;
  return launch_epilogue(*shape, weights, nullptr, nullptr, const_cast<double*>(sums), out, bwd_scale, false, true,
                         (cudaStream_t)stream);
}

int mafed_distill_epilogue(const mafed_shape_t* shape, const mafed_weights_t* weights, const int64_t* attn_mask,
                           const void* ws, double* sums, float* out, float* bwd_scale, void* stream) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!weights || !ws || !out || !bwd_scale) return MAFED_E_ARG;
  return launch_epilogue(*shape, weights, attn_mask, ws, sums, out, bwd_scale, true, true, (cudaStream_t)stream);
}

int mafed_distill_bwd(const mafed_shape_t* shape, const void* const* student_ptrs, const void* const* teacher_ptrs,
                      void* const* grad_ptrs, const int64_t* attn_mask, const float* bwd_scale, const float* grad_out,
                      void* stream) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!student_ptrs || !teacher_ptrs || !grad_ptrs || !bwd_scale) return MAFED_E_ARG;
  if (!shape->cls && shape->T > shape->n_vis && !attn_mask) return MAFED_E_ARG;
  if (!device_info().ok) return MAFED_E_NODEVICE;
  PathParams p;
  memset(&p, 0, sizeof(p));
  fill_geometry(*shape, p);
  const size_t es = elem_size(shape->dtype);
  for (int l = 0; l < shape->n_layers; ++l) {
    if (!student_ptrs[l] || !teacher_ptrs[l]) return MAFED_E_ARG;
    if (!aligned_to(student_ptrs[l], es) || !aligned_to(teacher_ptrs[l], es) || !aligned_to(grad_ptrs[l], es))
      return MAFED_E_ALIGN;
    p.s[l] = student_ptrs[l];
    p.t[l] = teacher_ptrs[l];
    p.g[l] = grad_ptrs[l];
  }
  p.mask = attn_mask;
  p.bwd_scale = bwd_scale;
  p.grad_out = grad_out;
  p.reverse = g_tune[kTuneBwdReverse].load() == 2 ? 0 : 1;
  return dispatch(*shape, p, true, (cudaStream_t)stream);
}

int mafed_distill_set_variant(int variant) {
  if (variant < 0 || variant > 2) return MAFED_E_ARG;
  g_variant.store(variant);
  return 0;
}

int mafed_distill_set_tuning(int key, int value) {
  if (key < 0 || key >= 8) return MAFED_E_ARG;
  g_tune[key].store(value);
  return 0;
}

}  // extern "C"
