// radix.cu -- stable LSD radix sort of (u64 key, u32 value) pairs, 8-bit digits.
//
// Replaces the reference's argsort switch (wendy/wendy.c:341-357: quick/merge/tim from
// wendy/sort.h, libc qsort, and the OpenMP task mergesort wendy/parallel_sort.c:95-137).
// Keys are the order-preserving u64 image of the fp64 positions (common.cuh); every
// pass is stable, so starting from values in particle-index order the result is the
// (key, index) order -- "ties broken by particle index".
//
// In this framework the full radix sort is the general entry point (initial sort,
// re-balancing of the bucket layout, forced 'gpu-radix' stepping, energy); the per-step
// fast path is the bucket kernel in tile.cu.
//
// Per pass: (1) per-tile digit histogram, (2) exclusive scan of the digit-major
// [256][ntiles] table, (3) stable scatter.  Optional extra passes sort on the bits of a
// segment id derived from the value (ensembles of independent realisations).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "internal.h"

namespace wendy {

constexpr int RT = 256;             // threads per block
constexpr int RI = 16;              // items per thread
constexpr int RTILE = RT * RI;      // 4096 pairs per tile
constexpr int RWARPS = RT / 32;

struct DigitSrc {
  int from_value;        // 0: digit from key, 1: digit from (value / seg_div)
  int shift;
  unsigned seg_div;
};

__device__ __forceinline__ unsigned digit_of(uint64_t key, uint32_t val, DigitSrc s) {
  if (s.from_value) return ((val / s.seg_div) >> s.shift) & 255u;
  return (unsigned)(key >> s.shift) & 255u;
}

__global__ void __launch_bounds__(RT)
radix_hist_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                  size_t n, DigitSrc src, uint32_t *__restrict__ table, unsigned ntiles) {
  __shared__ unsigned h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  size_t base = (size_t)blockIdx.x * RTILE;
#pragma unroll 4
  for (int j = 0; j < RI; j++) {
    size_t i = base + (size_t)j * RT + threadIdx.x;
    if (i < n) {
      uint32_t v = src.from_value ? vals[i] : 0u;
      uint64_t k = src.from_value ? 0ull : keys[i];
      atomicAdd(&h[digit_of(k, v, src)], 1u);
    }
  }
  __syncthreads();
  table[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// ---- exclusive scan of a u32 array (three small kernels) ---------------------------
constexpr int ST = 512, SI = 8, SCHUNK = ST * SI;

__device__ __forceinline__ unsigned block_exclusive_scan_u32(unsigned v, unsigned *warp_tot,
                                                             unsigned *total) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned inc = warp_inclusive_scan_u32(v, lane);
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned t = lane < nw ? warp_tot[lane] : 0u;
    unsigned ti = warp_inclusive_scan_u32(t, lane);
    if (lane < nw) warp_tot[lane] = ti - t;
    if (lane == 31) *total = ti;
  }
  __syncthreads();
  return inc - v + warp_tot[w];
}

__global__ void __launch_bounds__(ST)
scan_chunks_kernel(uint32_t *__restrict__ a, size_t n, uint32_t *__restrict__ sums) {
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned total;
  size_t base = (size_t)blockIdx.x * SCHUNK + (size_t)threadIdx.x * SI;
  unsigned v[SI], run = 0;
#pragma unroll
  for (int q = 0; q < SI; q++) {
    v[q] = (base + q < n) ? a[base + q] : 0u;
    run += v[q];
  }
  unsigned ex = block_exclusive_scan_u32(run, warp_tot, &total);
#pragma unroll
  for (int q = 0; q < SI; q++) {
    if (base + q < n) a[base + q] = ex;
    ex += v[q];
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
scan_sums_kernel(uint32_t *__restrict__ sums, unsigned m) {
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned total;
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (unsigned base = 0; base < m; base += 1024) {
    unsigned i = base + threadIdx.x;
    unsigned v = i < m ? sums[i] : 0u;
    unsigned ex = block_exclusive_scan_u32(v, warp_tot, &total);
    if (i < m) sums[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(ST)
scan_add_kernel(uint32_t *__restrict__ a, size_t n, const uint32_t *__restrict__ sums) {
  unsigned add = sums[blockIdx.x];
  size_t base = (size_t)blockIdx.x * SCHUNK + (size_t)threadIdx.x * SI;
#pragma unroll
  for (int q = 0; q < SI; q++)
    if (base + q < n) a[base + q] += add;
}

// ---- stable scatter ----------------------------------------------------------------
// Warp w of a block owns the contiguous run [tile + w*32*RI, tile + (w+1)*32*RI); item j
// of lane l is element j*32+l of that run, so (warp, j, lane) order is input order.
__global__ void __launch_bounds__(RT)
radix_scatter_kernel(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin,
                     uint64_t *__restrict__ kout, uint32_t *__restrict__ vout, size_t n,
                     DigitSrc src, const uint32_t *__restrict__ table, unsigned ntiles) {
  __shared__ unsigned wc[RWARPS][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RWARPS * 256; i += RT) (&wc[0][0])[i] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * RTILE + (size_t)w * (32 * RI);
  const unsigned lt = (1u << lane) - 1u;
  uint64_t key[RI];
  uint32_t val[RI];
  unsigned short dg[RI];
  unsigned rank[RI];
#pragma unroll
  for (int j = 0; j < RI; j++) {
    size_t i = base + (size_t)j * 32 + lane;
    bool ok = i < n;
    key[j] = ok ? kin[i] : 0ull;
    val[j] = ok ? vin[i] : 0u;
    unsigned d = ok ? digit_of(key[j], val[j], src) : 0xffffu;
    dg[j] = (unsigned short)d;
    unsigned mask = __match_any_sync(WENDY_FULL_MASK, d);
    int leader = __ffs(mask) - 1;
    unsigned old = 0;
    if (lane == leader && ok) {
      old = wc[w][d];
      wc[w][d] = old + __popc(mask);
    }
    __syncwarp();
    old = __shfl_sync(WENDY_FULL_MASK, old, leader);
    rank[j] = old + __popc(mask & lt);
  }
  __syncthreads();
  {  // thread d turns the per-warp counts of digit d into global exclusive bases
    unsigned d = threadIdx.x;
    unsigned run = table[(size_t)d * ntiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < RWARPS; ww++) {
      unsigned t = wc[ww][d];
      wc[ww][d] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RI; j++) {
    if (dg[j] != 0xffffu) {
      size_t pos = (size_t)wc[w][dg[j]] + rank[j];
      kout[pos] = key[j];
      vout[pos] = val[j];
    }
  }
}

// =====================================================================================================
// Onesweep: ONE read-and-write of the pairs per digit (Adinets & Merrill's scheme, written for this
// library).  All digit histograms are taken in a single up-front pass over the keys; each sorting pass is
// then one kernel in which a tile (4096 pairs) ranks its items, publishes its per-digit counts, resolves
// the counts of all preceding tiles by decoupled look-back (tiles are taken in launch order through a ticket, so
// a predecessor is always running or done), reorders the tile in shared memory so that every digit's items
// leave as one contiguous run, and stores them.  Passes whose digit is the same for every key are skipped
// (the high exponent bytes of clustered positions, unused segment-id bits).
// Traffic: 8 B/pair once (histograms) + 24 B/pair per pass, against 32 B/pair per pass + a table scan before.
constexpr int OT = 256;            // threads per tile
#ifndef WENDY_OI_BIG
#define WENDY_OI_BIG 16
#endif
constexpr int OI_BIG = WENDY_OI_BIG;  // pairs per thread: 4096 pairs per tile (large inputs: bandwidth-bound passes)
constexpr int OI_SMALL = 4;        // ... 1024 pairs per tile for inputs that live in L2 (a pass is then bound by the
                                   // serial work of one thread and by launch latency: four times the tiles, a
                                   // quarter of the work each; N=1e6: 20 -> 6 us per pass)
constexpr size_t SWEEP_SMALL_N = (size_t)1 << 22;
constexpr int OWARPS = OT / 32;
constexpr int MAXPASS = 12;        // 8 key bytes + up to 4 segment-id bytes
constexpr unsigned DESC_AGG = 1u << 30, DESC_INC = 2u << 30, DESC_VAL = (1u << 30) - 1u;

struct SweepPlan {
  int npass;
  DigitSrc src[MAXPASS];
  int any_value;                   // some pass takes its digit from the value
};

__global__ void __launch_bounds__(512)
radix_hist_all_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, size_t n,
                      const SweepPlan plan, unsigned *__restrict__ ghist) {
  __shared__ unsigned h[MAXPASS * 256];
  for (int i = threadIdx.x; i < plan.npass * 256; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t k = keys[i];
    const uint32_t v = plan.any_value ? vals[i] : 0u;
    for (int p = 0; p < plan.npass; p++)  // (spread shared atomics cost about as much as a shared load on sm_100a)
      atomicAdd(&h[p * 256 + digit_of(k, v, plan.src[p])], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < plan.npass * 256; i += blockDim.x)
    if (h[i]) atomicAdd(&ghist[i], h[i]);
}

// per pass: exclusive scan of the 256 global digit counts; skip[p] = 1 when one digit holds every key
__global__ void __launch_bounds__(256)
radix_prefix_kernel(const unsigned *__restrict__ ghist, unsigned *__restrict__ gbase, int *__restrict__ skip,
                    int npass, unsigned n) {
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned total;
  for (int p = 0; p < npass; p++) {
    const unsigned c = ghist[p * 256 + threadIdx.x];
    const unsigned ex = block_exclusive_scan_u32(c, warp_tot, &total);
    gbase[p * 256 + threadIdx.x] = ex;
    if (c == n) skip[p] = 1;
    __syncthreads();
  }
}

template <int OI>
struct SweepSmem {
  static constexpr int OTILE = OT * OI;
  uint64_t k[OTILE];
  uint32_t v[OTILE];
  unsigned wc[OWARPS][256];
  unsigned dig_off[256];
  unsigned glob_off[256];
  unsigned warp_tot[32];
  unsigned total;
  unsigned tile;
};

template <int OI>
__global__ void __launch_bounds__(OT)
onesweep_kernel(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint64_t *__restrict__ kout,
                uint32_t *__restrict__ vout, size_t n, const DigitSrc src, const unsigned *__restrict__ gbase,
                unsigned *desc, unsigned *ticket) {
  constexpr int OTILE = OT * OI;
  extern __shared__ __align__(16) unsigned char sweep_raw[];
  SweepSmem<OI> &S = *reinterpret_cast<SweepSmem<OI> *>(sweep_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  if (tid == 0) S.tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < OWARPS * 256; i += OT) (&S.wc[0][0])[i] = 0;
  __syncthreads();
  const unsigned tile = S.tile;
  const size_t tbase = (size_t)tile * OTILE;
  const unsigned tile_n = (unsigned)((n - tbase < (size_t)OTILE) ? (n - tbase) : (size_t)OTILE);
  // warp w owns the contiguous run [tbase + w*32*OI, ...): item j of lane l is element j*32 + l of it, so
  // (warp, j, lane) order is input order and the ranking below is stable
  const size_t wbase = tbase + (size_t)w * (32 * OI);
  uint64_t key[OI];
  uint32_t val[OI];
  unsigned short dg[OI], rk[OI];
#pragma unroll
  for (int j = 0; j < OI; j++) {
    const size_t i = wbase + (size_t)j * 32 + lane;
    const bool ok = i < n;
    key[j] = ok ? kin[i] : 0ull;
    val[j] = ok ? vin[i] : 0u;
  }
#pragma unroll
  for (int j = 0; j < OI; j++) {
    const bool ok = wbase + (size_t)j * 32 + lane < n;
    const unsigned d = ok ? digit_of(key[j], val[j], src) : 0xffffu;
    dg[j] = (unsigned short)d;
    // lanes holding the same digit: eight ballots (MATCH.ANY costs ~33 SM cycles per warp instruction on sm_100a
    // when the lanes' values differ, eight VOTEs ~5: scripts/ubench/atoms.cu)
    unsigned mask = __ballot_sync(WENDY_FULL_MASK, ok);
#pragma unroll
    for (int bit = 0; bit < 8; bit++) {
      const bool one = (d >> bit) & 1u;
      const unsigned bb = __ballot_sync(WENDY_FULL_MASK, one);
      mask &= one ? bb : ~bb;
    }
    if (!ok) mask = 1u << lane;
    const int leader = __ffs(mask) - 1;
    unsigned old = 0;
    if (lane == leader && ok) {
      old = S.wc[w][d];
      S.wc[w][d] = old + __popc(mask);
    }
    __syncwarp();
    old = __shfl_sync(WENDY_FULL_MASK, old, leader);
    rk[j] = (unsigned short)(old + __popc(mask & lt));
  }
  __syncthreads();
  // thread d owns digit d: per-warp counts -> offsets inside the digit's run; tile count published at once
  unsigned cnt_d = 0;
  {
    const unsigned d = tid;
#pragma unroll
    for (int ww = 0; ww < OWARPS; ww++) {
      const unsigned t = S.wc[ww][d];
      S.wc[ww][d] = cnt_d;
      cnt_d += t;
    }
    *(volatile unsigned *)(desc + (size_t)tile * 256 + d) = (tile == 0 ? DESC_INC : DESC_AGG) | cnt_d;
    const unsigned ex = block_exclusive_scan_u32(cnt_d, S.warp_tot, &S.total);
    S.dig_off[d] = ex;
  }
  __syncthreads();
  // reorder the tile by digit in shared memory (stable)
#pragma unroll
  for (int j = 0; j < OI; j++) {
    if (dg[j] != 0xffffu) {
      const unsigned pos = S.dig_off[dg[j]] + S.wc[w][dg[j]] + rk[j];
      S.k[pos] = key[j];
      S.v[pos] = val[j];
    }
  }
  // decoupled look-back, one digit per thread: items of digit d in all preceding tiles
  {
    const unsigned d = tid;
    unsigned excl = 0;
    if (tile > 0) {
      long long t = (long long)tile - 1;
      while (true) {
        unsigned vv;
        do {
          vv = *(volatile unsigned *)(desc + (size_t)t * 256 + d);
        } while ((vv >> 30) == 0u);
        excl += vv & DESC_VAL;
        if ((vv >> 30) == 2u) break;
        t--;
      }
      *(volatile unsigned *)(desc + (size_t)tile * 256 + d) = DESC_INC | (excl + cnt_d);
    }
    S.glob_off[d] = gbase[d] + excl - S.dig_off[d];
  }
  __syncthreads();
#pragma unroll 4
  for (unsigned i = tid; i < tile_n; i += OT) {
    const uint64_t kk = S.k[i];
    const uint32_t vv = S.v[i];
    const size_t pos = (size_t)S.glob_off[digit_of(kk, vv, src)] + i;
    kout[pos] = kk;
    vout[pos] = vv;
  }
}

// ---- host driver ---------------------------------------------------------------------
static inline unsigned ntiles_for(size_t n) { return (unsigned)((n + RTILE - 1) / RTILE); }

size_t radix_table_entries(size_t n) {
  // small inputs: look-back descriptors of ALL passes of the onesweep sort at 1024 pairs per tile (one clear per sort)
  // (scratch allocated for n also serves every smaller sort: the bound is monotone in n)
  const size_t ns = n < SWEEP_SMALL_N ? n : SWEEP_SMALL_N;
  const size_t small = (size_t)256 * MAXPASS * ((ns + OT * OI_SMALL - 1) / (OT * OI_SMALL) + 1);
  const size_t big = (size_t)256 * ntiles_for(n);
  return small > big ? small : big;
}
// (the onesweep path keeps its global histograms, their prefixes, the skip flags and the tile tickets here too)
constexpr size_t SWEEP_WORDS = (size_t)2 * MAXPASS * 256 + 2 * MAXPASS + 8;
size_t radix_sums_entries(size_t n) {
  const size_t a = (radix_table_entries(n) + SCHUNK - 1) / SCHUNK;
  return a > SWEEP_WORDS ? a : SWEEP_WORDS;
}

static void one_pass(cudaStream_t st, const uint64_t *kin, const uint32_t *vin, uint64_t *kout,
                     uint32_t *vout, size_t n, DigitSrc src, uint32_t *table, uint32_t *sums) {
  unsigned nt = ntiles_for(n);
  size_t m = (size_t)256 * nt;
  unsigned nchunks = (unsigned)((m + SCHUNK - 1) / SCHUNK);
  radix_hist_kernel<<<nt, RT, 0, st>>>(kin, vin, n, src, table, nt);
  scan_chunks_kernel<<<nchunks, ST, 0, st>>>(table, m, sums);
  if (nchunks > 1) {
    scan_sums_kernel<<<1, 1024, 0, st>>>(sums, nchunks);
    scan_add_kernel<<<nchunks, ST, 0, st>>>(table, m, sums);
  }
  radix_scatter_kernel<<<nt, RT, 0, st>>>(kin, vin, kout, vout, n, src, table, nt);
}

// WENDY_B200_RADIX=lsd selects the round-1 five-launch-per-pass sort (A/B runs)
static bool sweep_allowed() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("WENDY_B200_RADIX");
    v = !(e && e[0] == 'l');
  }
  return v != 0;
}

static int onesweep_sort_pairs(cudaStream_t st, RadixScratch &s, size_t n, int seg_bits, unsigned seg_div) {
  SweepPlan plan;
  memset(&plan, 0, sizeof(plan));
  for (int shift = 0; shift < 64; shift += 8) plan.src[plan.npass++] = DigitSrc{0, shift, 1u};
  for (int shift = 0; shift < seg_bits && plan.npass < MAXPASS; shift += 8) {
    plan.src[plan.npass++] = DigitSrc{1, shift, seg_div};
    plan.any_value = 1;
  }
  unsigned *ghist = s.sums, *gbase = s.sums + MAXPASS * 256, *ticket = s.sums + 2 * MAXPASS * 256;
  int *skip = (int *)(ticket + MAXPASS);
  cudaMemsetAsync(s.sums, 0, SWEEP_WORDS * sizeof(unsigned), st);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t hb = (n + 511) / 512;
  if (hb > (size_t)sms * 4) hb = (size_t)sms * 4;
  radix_hist_all_kernel<<<(unsigned)hb, 512, 0, st>>>(s.key[0], s.val[0], n, plan, ghist);
  radix_prefix_kernel<<<1, 256, 0, st>>>(ghist, gbase, skip, plan.npass, (unsigned)n);
  static OnceFlags attr_set;
  WENDY_ONCE_PER_DEVICE(attr_set) {
    cudaFuncSetAttribute(onesweep_kernel<OI_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SweepSmem<OI_BIG>));
    cudaFuncSetAttribute(onesweep_kernel<OI_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SweepSmem<OI_SMALL>));
  }
  int cur = 0;
  if (n < SWEEP_SMALL_N) {
    // Latency-bound regime: no host round trip (single-digit passes are not skipped -- a pass costs microseconds
    // here -- so the result buffer is known in advance) and one clear of the descriptors of all passes.
    const unsigned nt = (unsigned)((n + OT * OI_SMALL - 1) / (OT * OI_SMALL));
    cudaMemsetAsync(s.table, 0, (size_t)nt * 256 * plan.npass * sizeof(uint32_t), st);
    for (int p = 0; p < plan.npass; p++) {
      onesweep_kernel<OI_SMALL><<<nt, OT, sizeof(SweepSmem<OI_SMALL>), st>>>(
          s.key[cur], s.val[cur], s.key[cur ^ 1], s.val[cur ^ 1], n, plan.src[p], gbase + p * 256,
          s.table + (size_t)p * nt * 256, ticket + p);
      cur ^= 1;
    }
    return cur;
  }
  int h_skip[MAXPASS];
  cudaMemcpyAsync(h_skip, skip, sizeof(int) * MAXPASS, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);  // (which passes can be skipped decides the buffer the result ends in)
  const unsigned nt = (unsigned)((n + OT * OI_BIG - 1) / (OT * OI_BIG));
  for (int p = 0; p < plan.npass; p++) {
    if (h_skip[p]) continue;
    cudaMemsetAsync(s.table, 0, (size_t)nt * 256 * sizeof(uint32_t), st);
    onesweep_kernel<OI_BIG><<<nt, OT, sizeof(SweepSmem<OI_BIG>), st>>>(
        s.key[cur], s.val[cur], s.key[cur ^ 1], s.val[cur ^ 1], n, plan.src[p], gbase + p * 256, s.table, ticket + p);
    cur ^= 1;
  }
  return cur;
}

// Sorts the n pairs in (s.key[0], s.val[0]); returns the index (0/1) of the buffer pair
// that holds the result.  seg_bits > 0 adds most-significant passes on (val / seg_div).
int radix_sort_pairs(cudaStream_t st, RadixScratch &s, size_t n, int seg_bits, unsigned seg_div) {
  int cur = 0;
  if (n == 0) return cur;
  if (sweep_allowed() && n < ((size_t)1 << 30) && seg_bits <= 32) return onesweep_sort_pairs(st, s, n, seg_bits, seg_div);
  for (int shift = 0; shift < 64; shift += 8) {
    DigitSrc src = {0, shift, 1u};
    one_pass(st, s.key[cur], s.val[cur], s.key[cur ^ 1], s.val[cur ^ 1], n, src, s.table, s.sums);
    cur ^= 1;
  }
  for (int shift = 0; shift < seg_bits; shift += 8) {
    DigitSrc src = {1, shift, seg_div};
    one_pass(st, s.key[cur], s.val[cur], s.key[cur ^ 1], s.val[cur ^ 1], n, src, s.table, s.sums);
    cur ^= 1;
  }
  return cur;
}

}  // namespace wendy
