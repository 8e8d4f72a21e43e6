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

// ---- host driver ---------------------------------------------------------------------
static inline unsigned ntiles_for(size_t n) { return (unsigned)((n + RTILE - 1) / RTILE); }

size_t radix_table_entries(size_t n) { return (size_t)256 * ntiles_for(n); }
size_t radix_sums_entries(size_t n) { return (radix_table_entries(n) + SCHUNK - 1) / SCHUNK; }

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

// Sorts the n pairs in (s.key[0], s.val[0]); returns the index (0/1) of the buffer pair
// that holds the result.  seg_bits > 0 adds most-significant passes on (val / seg_div).
int radix_sort_pairs(cudaStream_t st, RadixScratch &s, size_t n, int seg_bits, unsigned seg_div) {
  int cur = 0;
  if (n == 0) return cur;
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
