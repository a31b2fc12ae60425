// Device side of the fused all-gather of a column-parallel linear: peer buffers and the flag protocol that
// replaces host-issued barriers.
//
// Every rank owns a symmetric-memory output buffer [M, ld] and a flag array of `world + 2` 32-bit words, all
// peer-mapped over NVLink.  The kernels of a gathered linear store their output tiles into the column slab
// [rank * n_local, (rank + 1) * n_local) of EVERY rank's buffer.  Synchronisation is carried by the kernels themselves:
//
//   flags[r]        (r < world)  written by rank r: "rank r has finished (stores included) its c-th gathered kernel"
//   flags[world]    local: number of gathered kernels this rank has completed (= the epoch the next one waits for)
//   flags[world+1]  local: CTAs of the running kernel that have finished (elects the last CTA)
//
//   start   a kernel whose INPUT is a gathered buffer reads c = flags[world] and waits until flags[r] >= c for all r
//           (every rank has finished the previous gathered kernel, so its slab of my input has landed).  The weight
//           stream does not wait: only the role that reads the activations does.
//   end     the last CTA to finish publishes c + 1 to flags[rank] of every rank (release.sys after a system fence) and
//           to its own flags[world]; with `wait_end` it also waits until all ranks have published c + 1, so that the
//           local buffer is complete when the kernel completes (what a non-kernel consumer -- a D2H copy, an ATen op --
//           needs).
//
// All ranks run the same sequence of gathered kernels (SPMD), which keeps the epochs in lockstep without any host
// traffic and makes the scheme CUDA-graph safe (nothing call-specific is baked into the launch).  A rank can never be
// more than one gathered kernel ahead of its slowest peer, so a buffer may be rewritten as soon as two other gathered
// kernels have run since it was last read (parallel.py gives every linear of a layer its own buffer).
// Every spin is watchdog-bounded: a missing peer traps the launch instead of hanging the GPU.
#pragma once

#include "common.cuh"

namespace qb {

constexpr int kMaxGatherWorld = 8;

struct GatherInfo {
  void* out_peer[kMaxGatherWorld];        // [M, ld] output buffer of every rank; entry 0 = this rank's own buffer
  uint32_t* flag_peer[kMaxGatherWorld];   // flag array of every rank, same order as out_peer
  uint32_t* flags;                        // this rank's flag array (== flag_peer[0])
  int n_out;                              // number of buffers (world); 1 = ordinary call, no protocol
  int world, rank;
  int wait_start, wait_end;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// one thread: wait until every rank's flag has reached `epoch`.  The polls are relaxed loads and ONE system-scope fence
// follows them (an acquire load per poll and rank costs a fence each: measured ~1 us per rank on 8 GPUs).
__device__ __forceinline__ void gather_wait_epoch(const uint32_t* flags, int world, uint32_t epoch) {
  for (int r = 0; r < world; ++r) {
    uint32_t probes = 0;
    uint64_t t0 = 0;
    while (static_cast<int32_t>(ld_relaxed_sys(flags + r) - epoch) < 0) {
      if (++probes == 1024u) {
        const uint64_t now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > QB_WATCHDOG_NS) __trap();
        probes = 0;
      }
    }
  }
  __threadfence_system();  // acquire side: the peers' data (written before their flags) is visible to what follows
}

// Called by ONE thread of the role that reads a gathered activation, before its first read.
__device__ __forceinline__ void gather_wait_start(const GatherInfo& g) {
  if (g.n_out <= 1 || !g.wait_start) return;
  const uint32_t c = ld_relaxed_gpu(g.flags + g.world);  // written by the previous kernel on this stream
  gather_wait_epoch(g.flags, g.world, c);
}

// Called by ONE thread per CTA after a CTA-wide barrier that follows the CTA's last output store (bulk stores waited
// for).  The last CTA of the grid publishes the new epoch.
// `callers`: how many threads of the grid call this (one per CTA, or one per CTA pair).
__device__ __forceinline__ void gather_signal_end(const GatherInfo& g, uint32_t callers) {
  if (g.n_out <= 1) return;
  __threadfence_system();  // this CTA's peer stores are performed before the ticket below is taken
  uint32_t* done = g.flags + g.world + 1;
  const uint32_t ticket = atomicAdd(done, 1u);
  if (ticket != callers - 1) return;
  __threadfence_system();  // acquire side of the tickets: every CTA's stores are ordered before the flags written below
  *done = 0u;
  const uint32_t c = ld_relaxed_gpu(g.flags + g.world) + 1u;
  // release side: the fence above orders every CTA's data before these flag stores; relaxed stores, not one release
  // (= one more fence) per peer
  for (int q = 0; q < g.n_out; ++q) st_relaxed_sys(g.flag_peer[q] + g.rank, c);
  if (g.wait_end) gather_wait_epoch(g.flags, g.world, c);
  g.flags[g.world] = c;
}
__device__ __forceinline__ void gather_signal_end(const GatherInfo& g) { gather_signal_end(g, gridDim.x); }

}  // namespace qb
