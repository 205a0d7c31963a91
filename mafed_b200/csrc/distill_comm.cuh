// One-shot, latency-bound allreduce of <= 130 doubles over NVLink peer memory, executed INSIDE the
// single-CTA scalar stage (so the distributed step has the same launch count as the single-GPU step and
// no NCCL call on its critical path).
//
// Every rank owns a small cudaMalloc'ed mailbox that its peers map with CUDA IPC:
//   data  [2 parities][world][kCommSlots] doubles   -- slot [parity][r] is written only by rank r
//   flags [2 parities][world] uint64                 -- epoch number, written by rank r after its data
// A collective with epoch e: rank r stores its vector into slot [e&1][r] of EVERY rank's mailbox (NVLink
// peer stores), fences system-wide, then publishes e into flag [e&1][r] of every mailbox; it then spins on
// its OWN flags until all ranks have published e and sums the slots in rank order -- a fixed order, so all
// ranks obtain bit-identical results.  Two parities suffice: a rank can be at most one epoch ahead of the
// slowest rank, because finishing epoch e requires every rank to have entered epoch e.
// The epoch counter lives in device memory and is bumped by the kernel, so the sequence is CUDA-graph safe.
// Spins are bounded (~2 s): on timeout the kernel records an error code and goes on, it never hangs the GPU.
#pragma once
#include "distill_common.cuh"

namespace mafed {

constexpr size_t kCommDataBytes = sizeof(double) * 2 * kCommMaxRanks * kCommSlots;
constexpr size_t kCommFlagBytes = sizeof(unsigned long long) * 2 * kCommMaxRanks;
constexpr size_t kCommMailboxBytes = kCommDataBytes + kCommFlagBytes + 64;  // + epoch + status

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
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
// part.  `epoch` = 0: take the next epoch from the device-side counter (CUDA-graph safe); otherwise the
// host-tracked epoch of this exchange (the counter is set to it).
template <int NT, typename Sync>
__device__ __forceinline__ void peer_allreduce(const CommDev& c, double* vals, int n, int tid, Sync sync,
                                               unsigned long long epoch = 0ull) {
  __shared__ unsigned long long s_epoch;
  sync();
  if (tid == 0) {
    if (epoch == 0ull) s_epoch = ++(*c.epoch);
    else { s_epoch = epoch; *c.epoch = epoch; }
  }
  sync();
  const unsigned long long e = s_epoch;
  const int par = (int)(e & 1ull);
  // 1. scatter my vector into my slot of every mailbox
  for (int i = tid; i < n * c.world; i += NT) {
    const int peer = i / n, k = i - peer * n;
    c.data[peer][((size_t)par * kCommMaxRanks + c.rank) * kCommSlots + k] = vals[k];
  }
  __threadfence_system();
  sync();
  // 2. publish
  if (tid < c.world) st_release_sys(c.flags[tid] + par * kCommMaxRanks + c.rank, e);
  // 3. wait for everybody's vector to land in my mailbox
  if (tid < c.world) {
    const unsigned long long* f = c.flags[c.rank] + par * kCommMaxRanks + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < e) {
      if (clock64() - t0 > kCommTimeoutCycles) { *c.status = 1; break; }
    }
  }
  sync();
  // 4. reduce in rank order (identical on every rank)
  const double* mine = c.data[c.rank] + (size_t)par * kCommMaxRanks * kCommSlots;
  for (int k = tid; k < n; k += NT) {
    double acc = 0.0;
    for (int r = 0; r < c.world; ++r) acc += __ldcg(mine + (size_t)r * kCommSlots + k);  // L2: peers wrote it
    vals[k] = acc;
  }
  sync();
}

}  // namespace mafed
