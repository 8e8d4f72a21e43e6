// lookback.cuh -- decoupled look-back over the buckets of one segment (device side).
// Two channels:
//   * counts: one packed 64-bit word per bucket, resolved at kernel entry (counts are known
//     before any sorting happens);
//   * masses (general-mass path only): exact 128-bit fixed-point sums + a status word.
// All sums are exact integers, so results never depend on which predecessors happened to
// have published an inclusive prefix already (bit-reproducible run to run).
#pragma once
#include "common.cuh"
#include "internal.h"

namespace wendy {

__device__ __forceinline__ i128 make_i128(unsigned long long lo, unsigned long long hi) {
  return (i128)(((u128)hi << 64) | (u128)lo);
}

// Decoupled look-back over the buckets of one segment; called by all lanes of warp 0.
// Returns the exclusive prefix (mass, count) of bucket b and publishes its inclusive one.
// Sums are exact integers, so the result does not depend on which predecessors happened
// to have published an inclusive prefix already.
__device__ __forceinline__ void lookback(Desc *descs, unsigned *status, unsigned epoch, int b, int seg_lo,
                                         i128 agg, long long n, int lane, i128 &P, long long &Pc) {
  P = 0;
  Pc = 0;
  Desc *me = descs + b;
  if (b > seg_lo) {
    if (lane == 0) {
      me->agg_lo = (unsigned long long)agg;
      me->agg_hi = (unsigned long long)((u128)agg >> 64);
      me->agg_cnt = n;
      __threadfence();
      *(volatile unsigned *)(status + b) = (epoch << 2) | 1u;
    }
    int j = b - 1;
    while (true) {
      int idx = j - lane;
      unsigned st = 2u;  // lanes before the segment start act as "inclusive prefix 0"
      i128 val = 0;
      long long c = 0;
      if (idx >= seg_lo) {
        do {
          st = ld_volatile_u32(status + idx);
        } while ((st >> 2) != epoch);
        st &= 3u;
        __threadfence();
        const Desc *d = descs + idx;
        if (st == 2u) {
          val = make_i128(__ldcg(&d->inc_lo), __ldcg(&d->inc_hi));
          c = __ldcg(&d->inc_cnt);
        } else {
          val = make_i128(__ldcg(&d->agg_lo), __ldcg(&d->agg_hi));
          c = __ldcg(&d->agg_cnt);
        }
      }
      unsigned incl = __ballot_sync(WENDY_FULL_MASK, st == 2u);
      int first = incl ? (__ffs(incl) - 1) : 32;
      if (lane > first) {
        val = 0;
        c = 0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        val += shfl_xor_i128(val, o);
        c += __shfl_xor_sync(WENDY_FULL_MASK, c, o);
      }
      P += val;
      Pc += c;
      if (incl) break;
      j -= 32;
    }
  }
  if (lane == 0) {
    i128 inc = P + agg;
    me->inc_lo = (unsigned long long)inc;
    me->inc_hi = (unsigned long long)((u128)inc >> 64);
    me->inc_cnt = Pc + n;
    __threadfence();
    *(volatile unsigned *)(status + b) = (epoch << 2) | 2u;
  }
}

// Count look-back through one packed 64-bit word per bucket: epoch(30) | state(2) | value(32).
// Counts are known when a CTA starts, so this runs at kernel entry and never waits for
// a predecessor's sort.  Called by all lanes of warp 0; returns the exclusive prefix.
__device__ __forceinline__ unsigned long long pack_cnt(unsigned epoch, unsigned state, unsigned v) {
  return ((unsigned long long)epoch << 34) | ((unsigned long long)state << 32) | v;
}
__device__ __forceinline__ unsigned count_lookback(volatile unsigned long long *cd, unsigned epoch, int b,
                                                   int seg_lo, unsigned n, int lane) {
  unsigned Pc = 0;
  if (b > seg_lo) {
    if (lane == 0) cd[b] = pack_cnt(epoch, 1u, n);
    int j = b - 1;
    while (true) {
      int idx = j - lane;
      unsigned st = 2u, val = 0u;
      if (idx >= seg_lo) {
        unsigned long long w;
        do {
          w = cd[idx];
        } while ((unsigned)(w >> 34) != epoch);
        st = (unsigned)(w >> 32) & 3u;
        val = (unsigned)w;
      }
      unsigned incl = __ballot_sync(WENDY_FULL_MASK, st == 2u);
      int first = incl ? (__ffs(incl) - 1) : 32;
      if (lane > first) val = 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(WENDY_FULL_MASK, val, o);
      Pc += val;
      if (incl) break;
      j -= 32;
    }
  }
  if (lane == 0) cd[b] = pack_cnt(epoch, 2u, Pc + n);
  return Pc;
}


}  // namespace wendy
