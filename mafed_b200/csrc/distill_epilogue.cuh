// Single-CTA scalar stage: deterministic fp64 reduction of the per-CTA partial sums, mask counts, the
// loss algebra of distillation.py:110-120,163 / distillation_loss_weights.py:148-174 and the
// backward scale table.  Which parts run is a set of flags so the same kernel serves as
//   epilogue  (reduce + counts + losses + scale)     single GPU, two-pass step
//   reduce    (reduce + counts)                      before the allreduce
//   finalize  (losses + scale from given sums)       after the allreduce
//   prologue  (counts + scale)                       before a fused single-pass step
#pragma once
#include "distill_comm.cuh"
#include "distill_common.cuh"

namespace MAFED_NS {

constexpr int kEpiThreads = 1024;

enum EpiFlags { kEpiReduce = 1, kEpiCounts = 2, kEpiLosses = 4, kEpiScale = 8 };

// What one scalar stage works on (the same record serves the stand-alone kernel and the in-kernel tail).
struct EpiArgs {
  const float* ws;        // partial sums (kEpiReduce): [kWsHeaderFloats] header, then [n_part][L][2]
  const int64_t* mask;    // [B, txt] (kEpiCounts)
  double* sums;           // [2L + 2]: read when a part is not computed here, written when it is (may be null)
  double* counts_out;     // optional second copy of the two (global) counts (the ws header, for a later tail)
  const double* counts_in;  // [2] token counts when kEpiCounts is not set (null: sums + 2L)
  float* out;             // [1 + 3L] (kEpiLosses)
  float* bwd_scale;       // [2L] (kEpiScale)
  long long n_mask;       // B * txt
  double n_vis_rows;      // B * n_vis (or B in cls mode)
  int n_part;             // number of per-CTA partial blocks in ws (<= 0: read it from the ws header)
  int n_layers;
  int D;
  int loss_kind;
  int flags;
  int comm_first;         // range of the sums vector to allreduce over the peer mailboxes
  int comm_count;         // (0 = no communication)
  unsigned long long comm_epoch;  // 0: next value of the device-side epoch counter; else the epoch to use (and store)
};

struct EpiParams {
  EpiArgs a;
  CommDev comm;
  mafed_weights_t w;
};

struct EpiSmem {
  double sums[2 * kMaxLayers + 2];
  double layer[kMaxLayers];
  long long cnt[kEpiThreads / 32];
};

// distillation.py:134-144: language mask = [0 x n_vis | attention_mask], image mask = [1 x n_vis | 0 x txt].
__global__ void __launch_bounds__(256) k_modality_masks(const int64_t* __restrict__ attn, int64_t* __restrict__ lang,
                                                        int64_t* __restrict__ image, long long n, int T, int n_vis) {
  pdl_wait();
  pdl_launch_dependents();
  const int txt = T - n_vis;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / T;
    const int t = (int)(i - b * T);
    const bool vis = t < n_vis;
    lang[i] = vis ? 0 : attn[b * txt + (t - n_vis)];
    image[i] = vis ? 1 : 0;
  }
}

// The scalar stage itself, run by threads tid = 0..NT-1 of one CTA that synchronise through `sync`.
template <int NT, typename Sync>
__device__ __forceinline__ void scalar_stage(const EpiArgs& p, const CommDev& comm, const mafed_weights_t& w,
                                             EpiSmem& sm, int tid, Sync sync) {
  const int warp = tid >> 5, lane = tid & 31;
  const int L = p.n_layers;
  double* s_sums = sm.sums;

  if (p.flags & kEpiReduce) {
    const int n_part = p.n_part > 0 ? p.n_part : reinterpret_cast<const int*>(p.ws)[0];
    const float* part = p.ws + kWsHeaderFloats;
    // one warp per (layer, modality) pair; lanes stride over the CTA partials in a fixed order (L2 loads: other
    // CTAs of the same grid may have written them)
    for (int pair = warp; pair < 2 * L; pair += NT / 32) {
      double acc = 0.0;
      for (int b = lane; b < n_part; b += 32) acc += (double)__ldcg(part + (size_t)b * 2 * L + pair);
      acc = warp_sum(acc);
      if (lane == 0) s_sums[pair] = acc;
    }
  } else if (p.flags & kEpiLosses) {
    for (int i = tid; i < 2 * L; i += NT) s_sums[i] = p.sums[i];
  }
  if (p.flags & kEpiCounts) {
    // text-token count = sum of the attention mask (distillation.py:248 `mask.sum()`)
    long long c = 0;
    for (long long i = tid; i < p.n_mask; i += NT) c += p.mask[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) sm.cnt[warp] = c;
    sync();
    if (tid == 0) {
      long long tot = 0;
      for (int i = 0; i < NT / 32; ++i) tot += sm.cnt[i];
      s_sums[2 * L] = (double)tot;
      s_sums[2 * L + 1] = p.n_vis_rows;
    }
  } else if (tid < 2) {
    const double* src = p.counts_in != nullptr ? p.counts_in : p.sums + 2 * L;
    s_sums[2 * L + tid] = src[tid];
  }
  sync();
  // batch-sharded step: combine this rank's sums / counts with its peers' over NVLink, in this kernel
  if (comm.world > 1 && p.comm_count > 0)
    peer_allreduce<NT>(comm, s_sums + p.comm_first, p.comm_count, tid, sync, p.comm_epoch);
  if (p.sums != nullptr) {
    if (p.flags & kEpiReduce)
      for (int i = tid; i < 2 * L; i += NT) p.sums[i] = s_sums[i];
    if (((p.flags & kEpiCounts) || p.counts_in != nullptr) && tid < 2) p.sums[2 * L + tid] = s_sums[2 * L + tid];
  }
  if (p.counts_out != nullptr && tid < 2) p.counts_out[tid] = s_sums[2 * L + tid];
  if (!(p.flags & (kEpiLosses | kEpiScale))) return;

  const bool want_loss = p.flags & kEpiLosses, want_scale = p.flags & kEpiScale;
  const double n_text = s_sums[2 * L], n_vis = s_sums[2 * L + 1];
  const double k = (p.loss_kind == MAFED_LOSS_MSE) ? 1.0 / (double)p.D : 1.0;
  const bool cls = w.modality_kind == MAFED_MODW_CLS;
  const bool text_only = w.modality_kind == MAFED_MODW_TEXT_ONLY;
  for (int l = tid; l < L; l += NT) {
    double wt, wv;
    modality_weights(w, l, n_text, n_vis, wt, wv);
    const double c = (double)w.layer_coeff[l] * (double)w.distill_coeff;
    if (want_loss) {
      const double text_loss = cls ? 0.0 : s_sums[2 * l] * k / n_text;  // 0/0 -> NaN, as in the reference
      const double vis_loss = text_only ? 0.0 : s_sums[2 * l + 1] * k / n_vis;
      const double layer_loss = cls ? vis_loss : (text_only ? text_loss : wt * text_loss + wv * vis_loss);
      sm.layer[l] = c * layer_loss;
      p.out[1 + l] = (float)layer_loss;
      p.out[1 + L + 2 * l] = (float)text_loss;
      p.out[1 + L + 2 * l + 1] = (float)vis_loss;
    }
    if (want_scale) backward_scales(w, l, n_text, n_vis, p.loss_kind, p.D, p.bwd_scale[2 * l], p.bwd_scale[2 * l + 1]);
  }
  if (!want_loss) return;
  sync();
  if (tid == 0) {
    double tot = 0.0;
    for (int l = 0; l < L; ++l) tot += sm.layer[l];
    p.out[0] = (float)tot;
  }
}

// Token counts ahead of the step (mafed_distill_prefetch_counts): one CTA sums the attention mask, takes the next
// value of the prefetch epoch counter, fires {n_text, n_vis rows} of this rank into every rank's mailbox (no wait)
// and leaves the ticket {epoch, bits(n_text), bits(n_vis rows)} for the step that will consume the counts.
struct PrefetchParams {
  const int64_t* mask;
  long long n_mask;
  double n_vis_rows;
  long long* ticket;   // [4]
  CommDev comm;
};

__global__ void __launch_bounds__(kEpiThreads) k_prefetch_counts(const __grid_constant__ PrefetchParams p) {
  __shared__ long long s_cnt[kEpiThreads / 32];
  __shared__ double s_vals[2];
  __shared__ unsigned long long s_epoch;
  pdl_wait();
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long c = 0;
  for (long long i = tid; i < p.n_mask; i += kEpiThreads) c += p.mask[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) s_cnt[warp] = c;
  __syncthreads();
  if (tid == 0) {
    long long tot = 0;
    for (int i = 0; i < kEpiThreads / 32; ++i) tot += s_cnt[i];
    s_vals[0] = (double)tot;
    s_vals[1] = p.n_vis_rows;
    unsigned long long e = 0ull;
    if (p.comm.world > 1) e = ++(*p.comm.epoch_counts);
    s_epoch = e;
    p.ticket[0] = (long long)e;
    p.ticket[1] = __double_as_longlong(s_vals[0]);
    p.ticket[2] = __double_as_longlong(s_vals[1]);
    p.ticket[3] = 0;
  }
  __syncthreads();
  if (p.comm.world > 1 && tid < 2 * p.comm.world) {
    const int peer = tid >> 1, k = tid & 1;
    const unsigned long long e = s_epoch;
    ll_store(ll_count_slot(p.comm.ll[peer], (int)(e % kCommCountSlots), p.comm.rank, k), s_vals[k], ll_tag(e));
  }
}

__global__ void __launch_bounds__(kEpiThreads) k_epilogue(const __grid_constant__ EpiParams p) {
  __shared__ EpiSmem sm;
  pdl_wait();
  pdl_launch_dependents();
  scalar_stage<kEpiThreads>(p.a, p.comm, p.w, sm, (int)threadIdx.x, SyncCta());
}

}  // namespace MAFED_NS
