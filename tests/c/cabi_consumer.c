/* A plain C99 consumer of include/mafed_distill.h: what a non-Python binding (cgo / JNI / FFI shim) compiles
 * against.  Run without a GPU it exercises only the entry points that do no device work. */
#include <stdio.h>
#include <string.h>

#include "mafed_distill.h"

int main(void) {
  mafed_shape_t shape;
  mafed_weights_t w;
  int failures = 0;
  memset(&shape, 0, sizeof(shape));
  memset(&w, 0, sizeof(w));
  if (mafed_distill_abi_version() != MAFED_ABI_VERSION) ++failures;
  if (mafed_distill_sums_len(15) != 32 || mafed_distill_out_len(15) != 46) ++failures;
  if (mafed_distill_ws_bytes(15) < 148u * 15u * 2u * sizeof(float)) ++failures;
  /* n_layers = 0 is rejected before any CUDA call */
  if (mafed_distill_fwd(&shape, NULL, NULL, NULL, NULL, NULL) != MAFED_E_ARG) ++failures;
  shape.n_layers = 1; shape.B = 1; shape.T = 4; shape.n_vis = 2; shape.D = 8; shape.dtype = 9;
  if (mafed_distill_fwd(&shape, NULL, NULL, NULL, NULL, NULL) != MAFED_E_DTYPE) ++failures;
  shape.dtype = MAFED_BF16;
  w.modality_kind = MAFED_MODW_EQUAL; w.distill_coeff = 1.0f; w.layer_coeff[0] = 1.0f;
  if (mafed_distill_step(&shape, NULL, NULL, NULL, NULL, &w, 1.0f, NULL, NULL, NULL, NULL, NULL, NULL, NULL,
                         NULL, NULL) != MAFED_E_ARG) ++failures;
  if (mafed_distill_bwd(&shape, NULL, NULL, NULL, NULL, NULL, NULL, 1.0f, NULL, NULL, NULL) != MAFED_E_ARG) ++failures;
  if (mafed_distill_prefetch_counts(&shape, NULL, NULL, NULL, NULL) != MAFED_E_ARG) ++failures;
  if (strstr(mafed_distill_error_string(MAFED_E_ALIGN), "aligned") == NULL) ++failures;
  printf("%s\n", failures ? "FAIL" : "OK");
  return failures;
}
