// Host-buffer step behind the C ABI: inputs and outputs live in host memory (pinned for full PCIe speed).
// Per layer: H2D(student, teacher) on a copy-in stream -> one-pass fused kernel (+ its two scalar stages)
// on the compute stream -> D2H(gradient) on a copy-out stream.  The three streams form a layer pipeline so
// that both PCIe directions and the kernels overlap.  The backward scale of a layer depends only on the
// mask counts and the host weight tables, which is what makes the per-layer pipeline legal.
#pragma once
#include <vector>

#include "distill_common.cuh"

struct mafed_host_step {
  mafed_shape_t shape;           // n_layers = layers per step
  size_t layer_bytes = 0;
  size_t mask_bytes = 0;
  char* d_pool = nullptr;        // device staging: L x (s, t, g) + mask + per-layer scalars
  std::vector<void*> d_s, d_t, d_g;
  int64_t* d_mask = nullptr;
  char* d_ws = nullptr;          // per-layer workspace
  size_t ws_bytes = 0;
  float* d_out = nullptr;        // [L][4]
  float* d_scale = nullptr;      // [L][2]
  double* d_sums = nullptr;      // [L][4] per-layer sums vector of a batch-sharded step
  int64_t* d_ticket = nullptr;   // [4] token counts sent ahead of the step (batch-sharded)
  cudaEvent_t ev_mask = nullptr;
  float* h_out = nullptr;        // pinned [L][4]
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_run;
};
