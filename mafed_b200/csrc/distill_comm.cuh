// One-shot, latency-bound allreduce of <= 130 doubles over NVLink peer memory, executed INSIDE the path's own
// kernels (so the distributed step has the same launch count as the single-GPU step and no NCCL call on its
// critical path).
//
// Every rank owns a small cudaMalloc'ed mailbox that its peers map with CUDA IPC:
//   ll [2 parities][world][kCommSlots][2] uint64   -- slot [parity][r] is written only by rank r
// Each double travels as two self-validating 8-byte words {32 data bits | 32-bit epoch tag} (the "LL" scheme of
// NCCL's low-latency protocol): an aligned 8-byte store is never torn, so a word whose tag equals the current
// epoch carries valid data.  No fence and no separate flag: a sender just fires its stores at every peer's
// mailbox (one NVLink one-way trip, ~1-2 us) and goes on; a receiver polls the words of its OWN mailbox
// (local L2) until the tags match and sums the values in rank order -- a fixed order, so all ranks obtain
// bit-identical results.  (The first version stored data, fenced system-wide and then published a flag: the fence
// alone cost ~4 us per exchange, `mafed_comm_trace`.)  Two parities suffice: a rank can be at most one epoch ahead
// of the slowest rank, because finishing epoch e requires every rank to have entered epoch e; a slot therefore
// still holds tag e-2 (or 0) until its owner writes epoch e.
// The epoch counter lives in device memory and is bumped by the kernel, so the sequence is CUDA-graph safe.
// Spins are bounded (60 s by default, mafed_comm_set_timeout): on timeout the kernel records an error code, returns
// NaN for the missing values and goes on -- it never hangs the GPU, and the step's loss and gradients come out NaN.
#pragma once
#include "distill_common.cuh"

namespace MAFED_NS {

constexpr size_t kCommDataBytes = sizeof(unsigned long long) * 2 * 2 * kCommMaxRanks * kCommSlots;
// Second region: token counts sent AHEAD of the step (mafed_distill_prefetch_counts).  They have their own epoch
// counter and kCommCountSlots generations, so a rank may prefetch the counts of up to kCommCountSlots - 1 batches
// before the step that consumes the first of them (the sums exchange of every step keeps the ranks within one
// step of each other):
//   cnt [kCommCountSlots][world][2 values][2 words] uint64 -- slot [e % kCommCountSlots][r] is written only by rank r
constexpr int kCommCountSlots = 4;
constexpr size_t kCommCountsBytes = sizeof(unsigned long long) * kCommCountSlots * kCommMaxRanks * 2 * 2;
// tail of the mailbox: +0 epoch, +8 counts epoch, +16 (unused), +32 trace[4]
constexpr size_t kCommTailAt = kCommDataBytes + kCommCountsBytes;
constexpr size_t kCommMailboxBytes = kCommTailAt + 64;

__device__ __forceinline__ uint32_t ll_tag(unsigned long long epoch) { return (uint32_t)epoch | 0x80000000u; }

// Word pair of slot k of rank r's vector in a mailbox.
__device__ __forceinline__ unsigned long long* ll_slot(unsigned long long* mailbox, int parity, int r, int k) {
  return mailbox + (((size_t)parity * kCommMaxRanks + r) * kCommSlots + k) * 2;
}

// Word pair of count k (0 text, 1 vision rows) of rank r, generation `gen`, in a mailbox.
__device__ __forceinline__ unsigned long long* ll_count_slot(unsigned long long* mailbox, int gen, int r, int k) {
  return mailbox + kCommDataBytes / sizeof(unsigned long long) + (((size_t)gen * kCommMaxRanks + r) * 2 + k) * 2;
}

__device__ __forceinline__ void ll_store(unsigned long long* slot, double v, uint32_t tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = (b & 0xffffffffull) | ((unsigned long long)tag << 32);
  const unsigned long long w1 = (b >> 32) | ((unsigned long long)tag << 32);
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" :: "l"(slot), "l"(w0), "l"(w1) : "memory");
}

// Spin until both words of `slot` carry `tag`; returns the value.  After the spin bound: *status = 1 and NaN, so
// that a peer that never arrived poisons the loss and the gradients visibly instead of leaving plausible numbers.
__device__ __forceinline__ double ll_wait(const unsigned long long* slot, uint32_t tag, int* status,
                                          long long timeout_cycles) {
  const long long t0 = clock64();
  for (;;) {
    unsigned long long w0, w1;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
    if ((uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag)
      return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    if (clock64() - t0 > timeout_cycles) {
      *status = 1;
      return __longlong_as_double(0x7ff8000000000000LL);
    }
  }
}

// How the threads taking part in a scalar stage synchronise: the whole CTA (stand-alone k_epilogue) or the
// consumer warps of a streaming kernel through a named barrier (the in-kernel tail, whose producer warp has
// already exited).
struct SyncCta {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
template <int NT>
struct SyncNamed {
  __device__ __forceinline__ void operator()() const { asm volatile("bar.sync 1, %0;" :: "n"(NT) : "memory"); }
};

// In-place SUM-allreduce of vals[0..n) (shared memory) across the communicator; threads tid = 0..NT-1 take
// part.  `epoch` = 0: take the next epoch from the device-side counter; otherwise use the given epoch (the one
// after an earlier exchange of the same kernel) and set the counter to it.  Either way the epochs live on the
// device only, so every sequence of exchanges is CUDA-graph replayable.
template <int NT, typename Sync>
__device__ __forceinline__ void peer_allreduce(const CommDev& c, double* vals, int n, int tid, Sync sync,
                                               unsigned long long epoch = 0ull) {
  __shared__ unsigned long long s_epoch;
  const long long t_in = clock64();
  sync();
  if (tid == 0) {
    if (epoch == 0ull) s_epoch = ++(*c.epoch);
    else { s_epoch = epoch; *c.epoch = epoch; }
  }
  sync();
  const unsigned long long e = s_epoch;
  const int par = (int)(e & 1ull);
  const uint32_t tag = ll_tag(e);
  // 1. fire my vector into my slot of every mailbox (peer stores over NVLink; my own included)
  for (int i = tid; i < n * c.world; i += NT) {
    const int peer = i / n, k = i - peer * n;
    ll_store(ll_slot(c.ll[peer], par, c.rank, k), vals[k], tag);
  }
  const long long t_pub = clock64();
  sync();  // every thread has read its vals[] before they are overwritten below
  // 2. element k: wait for every rank's word pair in my own mailbox, sum in rank order (identical on every rank)
  for (int k = tid; k < n; k += NT) {
    double acc = 0.0;
    for (int r = 0; r < c.world; ++r) acc += ll_wait(ll_slot(c.ll[c.rank], par, r, k), tag, c.status, c.timeout_cycles);
    vals[k] = acc;
  }
  sync();
  if (tid == 0) {   // where the time of an exchange goes (mafed_comm_trace): own stores vs waiting for the peers
    const long long t_seen = clock64();
    c.trace[1] += (unsigned long long)(t_pub - t_in);
    c.trace[2] += (unsigned long long)(t_seen - t_pub);
    c.trace[3] += 1ull;
  }
}

}  // namespace MAFED_NS
