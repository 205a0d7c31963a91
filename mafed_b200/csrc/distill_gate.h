// Entry of the relocatable-device-code unit (distill_gate.cu) as the rest of the library sees it.
#pragma once
#include "mafed_distill.h"

namespace mafed_gate {

// mafed_distill_bwd(..., skip_if_equals): enqueue a 1-CTA gate that compares g = *grad_out * grad_out_scale with
// `assumed` on the device; equal -> nothing else happens; different -> the gate tail-launches the backward kernel
// the host would have chosen for this shape (same family, same geometry), stream-ordered behind itself.
int gated_backward(const mafed_shape_t* shape, const void* const* student_ptrs, const void* const* teacher_ptrs,
                   void* const* grad_ptrs, const int64_t* attn_mask, const float* bwd_scale, const float* grad_out,
                   float grad_out_scale, float assumed, float* grad_out_seen, void* stream);

// Ordering latch (see distill_gate.cu): every launch of the library asks `consume_gate_mark(stream)` and, if a gate
// was the last thing it sent to that stream, launches without the programmatic-dependent-launch attribute.
void note_gate_launch(void* stream);
bool consume_gate_mark(void* stream);

}  // namespace mafed_gate
