// peer.cuh -- device side of the sharded system's exchange over peer memory (see PeerComm in internal.h).
//
// One leapfrog sub-step `e` on rank q, two launches on q's stream, no host involvement:
//   step kernel e    waits for cnt_flag[e-1] of every peer (local memory): the particle counts of the lower
//                    ranks offset the cumulative mass; a particle whose new key leaves q's range is stored
//                    straight into the owner's inbox over NVLink; the last CTA to finish publishes, to every
//                    peer, how many records it wrote there (in_flag[e]).
//   inject kernel e  waits for in_flag[e] of every peer, appends the inbox records to the local buckets, and
//                    the last CTA publishes the new local count to every peer (cnt_flag[e]).
// Inboxes and flag words are double-buffered by the parity of e: a peer can only be one epoch ahead (its step
// e+1 needs this rank's cnt_flag[e]), and its writes of epoch e+2 need this rank's in_flag[e+1], which is
// issued after this rank's inject e has finished reading the parity-e buffers.
//
// A failure anywhere (bucket or inbox overflow, or a failure learnt from a peer) sets the fail bit of every later
// flag this rank sends, and every kernel launched behind it only forwards flags: within one flag exchange all
// ranks have stopped, the host of each rank sees fail_seq at call end, and all ranks roll back to the input of
// the failing sub-step (api.cu).  Waits give up after PeerComm::timeout_ns (a peer's host died): the launch
// fails instead of hanging the GPU.
#pragma once
#include "common.cuh"
#include "internal.h"

namespace wendy {

__device__ __forceinline__ unsigned long long peer_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned long long peer_ld_flag(const unsigned long long *p) {
  unsigned long long w;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}
__device__ __forceinline__ void peer_st_flag(unsigned long long *p, unsigned long long w) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(w) : "memory");
}

// Wait until the word written by a peer carries `epoch`; false after a timeout.
__device__ __forceinline__ bool peer_wait(const unsigned long long *flag, unsigned epoch, unsigned long long &w,
                                          unsigned long long timeout_ns) {
  w = peer_ld_flag(flag);
  if ((unsigned)(w >> 32) == epoch) return true;
  const unsigned long long t0 = peer_now_ns();
  unsigned spins = 0;
  while (true) {
    w = peer_ld_flag(flag);
    if ((unsigned)(w >> 32) == epoch) return true;
    if ((++spins & 1023u) == 0u && peer_now_ns() - t0 > timeout_ns) return false;
    __nanosleep(64);
  }
}

// Called by ONE thread: sum the peers' counts of `epoch` (cnt_flag) below my rank; bad = some peer failed / timed out.
__device__ __forceinline__ long long peer_wait_counts(const PeerComm *pc, unsigned epoch, bool &bad) {
  long long off = 0;
  bad = false;
  for (int r = 0; r < pc->nranks; r++) {
    if (r == pc->my_rank) continue;
    unsigned long long w;
    if (!peer_wait(pc->cnt_flag + (epoch & 1u) * PEER_MAX + r, epoch, w, pc->timeout_ns)) {
      bad = true;
      atomicOr(pc->peer_stat, 1u);
      continue;
    }
    if ((w >> 31) & 1ull) bad = true;
    if (r < pc->my_rank) off += (long long)(w & 0x7fffffffull);
  }
  return off;
}

// Called by ONE thread after every record of this launch is visible system-wide: tell every peer how many
// records of sub-step `epoch` were written into its inbox.
__device__ __forceinline__ void peer_signal_step(const PeerComm *pc, unsigned epoch, bool fail) {
  for (int r = 0; r < pc->nranks; r++) {
    if (r == pc->my_rank) continue;
    const unsigned c = fail ? 0u : __ldcg(pc->out_cnt + r);
    peer_st_flag(pc->peer_in_flag[r] + (epoch & 1u) * PEER_MAX + pc->my_rank, peer_pack(epoch, fail, c));
  }
}
__device__ __forceinline__ void peer_signal_count(const PeerComm *pc, unsigned epoch, bool fail, long long n) {
  for (int r = 0; r < pc->nranks; r++) {
    if (r == pc->my_rank) continue;
    peer_st_flag(pc->peer_cnt_flag[r] + (epoch & 1u) * PEER_MAX + pc->my_rank, peer_pack(epoch, fail, (unsigned)n));
  }
}

}  // namespace wendy
