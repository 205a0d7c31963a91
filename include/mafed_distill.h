/*
 * mafed_distill.h -- C ABI of the B200-native MAFED feature-distillation hot path.
 *
 * The reference (MalvinaNikandrou/mafed) is pure Python: it has no FFI / operator boundary of its
 * own.  Its plug-in boundary is the Python class registry mafed/methods/__init__.py:6-11
 * (CLMethod["featdistill"] -> FeatureDistillation), mirrored by mafed_b200/methods/.  This header
 * is the boundary UNDERNEATH that mirror: plain pointers and sizes, no torch types, one entry
 * point per step of the path.  Each function cites the reference code it replaces.
 *
 * Conventions
 *   - every `*_ptrs` argument is a HOST array of n_layers DEVICE pointers, one per selected layer;
 *     hidden states are separately allocated [B, T, D] tensors (vl_pythia.py:320-326), row-major,
 *     contiguous in D; consecutive token rows are `row_stride` elements apart (D when contiguous).
 *   - `attn_mask` is the DEVICE int64 [B, T - n_vis] left-padded text attention mask
 *     (data/vl_pythia_vqa_dataset.py:141-142).  Positions t <  n_vis are visual tokens with
 *     weight 1 (distillation.py:140-144); t >= n_vis are text tokens weighted by
 *     attn_mask[b, t - n_vis] (distillation.py:135-139).
 *   - `stream` is a cudaStream_t passed as void*.
 *   - return value: 0 = ok; < 0 = argument error (MAFED_E_*); > 0 = a cudaError_t.
 *   - no allocation, no host synchronisation, no global mutable state: CUDA-graph capturable.
 */
#ifndef MAFED_DISTILL_H_
#define MAFED_DISTILL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAFED_ABI_VERSION 1
#define MAFED_MAX_LAYERS 64

enum { MAFED_F32 = 0, MAFED_BF16 = 1, MAFED_F16 = 2 };
enum { MAFED_LOSS_MSE = 0, MAFED_LOSS_COSINE = 1 };
/* modality weighting (distillation_loss_weights.py:71-79):
 *   EQUAL: w_text = n_text/(n_text+n_vis) from the mask counts (:148-155);
 *   TABLE: w_text = lang_weight[l], w_vis = 1 - w_text (balanced :165-166 -> 0.5; adaptive :168-174);
 *   CLS  : only row 0 of each sample, mean over B (distillation.py:251-257);
 *   TEXT_ONLY: layer loss = masked text loss alone -- one call of _compute_*_distillation_loss
 *          with an arbitrary [B, T] mask (distillation.py:226-249; use n_vis = 0). */
enum { MAFED_MODW_EQUAL = 0, MAFED_MODW_TABLE = 1, MAFED_MODW_CLS = 2, MAFED_MODW_TEXT_ONLY = 3 };

enum {
  MAFED_E_ARG = -1,      /* null pointer / non-positive size / too many layers */
  MAFED_E_DTYPE = -2,    /* unknown dtype or loss kind */
  MAFED_E_ALIGN = -3,    /* a pointer is not aligned to its element size */
  MAFED_E_NODEVICE = -4  /* no CUDA device / not an sm_100 device */
};

/* Geometry of one step.  rows of a layer: N = B*T (or B in CLS mode). */
typedef struct mafed_shape {
  int32_t n_layers;   /* selected layers in this call, 1..MAFED_MAX_LAYERS */
  int32_t B;          /* samples in this rank's shard */
  int32_t T;          /* n_vis + text positions */
  int32_t n_vis;      /* 256 in the reference (distillation.py:73) */
  int32_t D;          /* hidden size */
  int32_t dtype;      /* MAFED_F32 | MAFED_BF16 | MAFED_F16 (student, teacher and gradient) */
  int32_t loss_kind;  /* MAFED_LOSS_MSE (distillation.py:237-249) | MAFED_LOSS_COSINE (:226-235) */
  int32_t cls;        /* 1: CLS distillation (distillation.py:126-132,251-257) */
} mafed_shape_t;

/* Host-side weight tables (distillation.py:110-113,163; distillation_loss_weights.py:49-60). */
typedef struct mafed_weights {
  int32_t modality_kind;                   /* MAFED_MODW_* */
  float distill_coeff;                     /* self.distillation_coeff */
  float layer_coeff[MAFED_MAX_LAYERS];     /* get_layer_loss_weight(layer) for each selected layer */
  float lang_weight[MAFED_MAX_LAYERS];     /* TABLE only */
} mafed_weights_t;

int mafed_distill_abi_version(void);
const char* mafed_distill_error_string(int code);

/* Bytes of device workspace mafed_distill_fwd needs (per-CTA partial sums). */
size_t mafed_distill_ws_bytes(int n_layers);

/* Length (in doubles) of the `sums` vector: 2*n_layers partial sums + [n_text, n_vis]. */
int mafed_distill_sums_len(int n_layers);
/* Length (in floats) of the `out` vector: total, n_layers layer losses, 2*n_layers modality losses. */
int mafed_distill_out_len(int n_layers);

/* Fused forward over all selected layers: one pass over student and teacher
 * (replaces the 2 x _compute_{mse,cosine}_distillation_loss passes per layer,
 * distillation.py:152-162,226-249).  Writes per-CTA partial [layer][text|vision] sums to `ws`. */
int mafed_distill_fwd(const mafed_shape_t* shape, const void* const* student_ptrs,
                      const void* const* teacher_ptrs, const int64_t* attn_mask, void* ws, void* stream);

/* Deterministic reduction of `ws` (fixed order, fp64) + mask counts -> sums[2L+2] on device.
 * This rank-local vector is what the single NCCL allreduce combines across batch shards. */
int mafed_distill_reduce(const mafed_shape_t* shape, const int64_t* attn_mask, const void* ws,
                         double* sums, void* stream);

/* sums (global) -> out[1+3L] = {total, layer losses, modality losses} and the backward scale table
 * bwd_scale[2L] = c_l * coeff * w_m * k / n_m  (k = 2/D for mse, 1 for cosine).
 * Replaces distillation.py:110-120,163 and distillation_loss_weights.py:148-174. */
int mafed_distill_finalize(const mafed_shape_t* shape, const mafed_weights_t* weights, const double* sums,
                           float* out, float* bwd_scale, void* stream);

/* reduce + finalize in one launch (single-GPU step); `sums` (may be NULL) also receives the sums. */
int mafed_distill_epilogue(const mafed_shape_t* shape, const mafed_weights_t* weights,
                           const int64_t* attn_mask, const void* ws, double* sums, float* out,
                           float* bwd_scale, void* stream);

/* Fused backward: grad[l][row] = grad_out * bwd_scale[l][m(row)] * w(row) * d f(h,p)/dh, zero for
 * padded text rows; one pass, 2 reads + 1 write (replaces autograd's per-layer, per-modality chains
 * of distillation.py:226-249).  `grad_out` is a DEVICE float scalar (NULL = 1.0).  A NULL entry in
 * grad_ptrs skips that layer. */
int mafed_distill_bwd(const mafed_shape_t* shape, const void* const* student_ptrs,
                      const void* const* teacher_ptrs, void* const* grad_ptrs, const int64_t* attn_mask,
                      const float* bwd_scale, const float* grad_out, void* stream);

/* Experiment knob (benchmarks only): select the kernel family for the next calls.
 * 0 = default, 1 = ldg (register-staged 128-bit loads), 2 = tma (cp.async.bulk + mbarrier ring). */
int mafed_distill_set_variant(int variant);
int mafed_distill_set_tuning(int key, int value);

#ifdef __cplusplus
}
#endif
#endif /* MAFED_DISTILL_H_ */
