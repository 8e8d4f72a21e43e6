// hostmem.cuh -- host-side memory plumbing of the handle API (included by api.cu only):
//   * a cache of large device blocks released by destroyed handles,
//   * piecewise page-locking of host ranges and host<->device copies split at the same boundaries,
//   * the ring of page-locked bounce buffers behind the read-out of pageable destinations,
//   * set-up helpers (page pre-faulting) and the WENDY_B200_TRACE phase timer.
// Nothing here touches particle data semantics; the reference has no counterpart (its state lives in numpy
// arrays, wendy/wendy.py:369-387).
#pragma once

#include <omp.h>
#include <stdint.h>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/wendy_b200.h"

// Host threads the library's own copy / validation loops may use.  Default: what OpenMP offers, at most 32.  A
// multi-process job (one rank per GPU) must divide the cores between its ranks -- oversubscribed OpenMP teams spin
// on each other -- so multi.py calls wendy_host_set_threads(cores / ranks on this node); WENDY_B200_HOST_THREADS
// overrides both.
static int g_host_threads = 0;
static int host_threads() {
  static int env = -1;
  if (env < 0) { const char *e = getenv("WENDY_B200_HOST_THREADS"); env = e ? std::max(1, atoi(e)) : 0; }
  if (env > 0) return env;
  if (g_host_threads > 0) return g_host_threads;
  return std::max(1, std::min(32, omp_get_max_threads()));
}
extern "C" void wendy_host_set_threads(int n) { g_host_threads = n > 0 ? n : 0; }

// ---- device block cache -----------------------------------------------------------------------------------
// cudaMalloc / cudaFree of the multi-GB state arrays cost 30-100 ms per GB on this platform (measured:
// profiles/r01/e2e_phases_N1e8.txt), which dominates the set-up of a generator.  Blocks of >= 32 MB released
// by wendy_cuda_destroy are therefore kept and handed to the next handle that asks for exactly the same size
// on the same device (a new generator over the same N: the usual notebook pattern).  The cache is bounded
// (WENDY_B200_ALLOC_CACHE_GB, default 48; 0 disables it), emptied when a cudaMalloc fails, and can be
// emptied by the caller with wendy_cuda_trim().
namespace {
struct BlockCache {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void *> idle;
  std::map<void *, std::pair<int, size_t>> live;
  size_t idle_bytes = 0;
};
BlockCache g_blocks;
constexpr size_t kCacheMinBlock = (size_t)32 << 20;
size_t cache_limit_bytes() {
  static long long v = -1;
  if (v < 0) {
    const char *e = getenv("WENDY_B200_ALLOC_CACHE_GB");
    v = (long long)((e ? atof(e) : 48.) * 1073741824.);
    if (v < 0) v = 0;
  }
  return (size_t)v;
}
void cache_trim() {
  std::vector<void *> drop;
  {
    std::lock_guard<std::mutex> lk(g_blocks.mu);
    for (auto &kv : g_blocks.idle) drop.push_back(kv.second);
    g_blocks.idle.clear();
    g_blocks.idle_bytes = 0;
  }
  for (void *q : drop) cudaFree(q);
}
cudaError_t dev_alloc_bytes(void **out, size_t bytes) {
  *out = nullptr;
  const bool big = bytes >= kCacheMinBlock && cache_limit_bytes() > 0;
  int dev = 0;
  if (big) {
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_blocks.mu);
    auto it = g_blocks.idle.find(std::make_pair(dev, bytes));
    if (it != g_blocks.idle.end()) {
      *out = it->second;
      g_blocks.idle.erase(it);
      g_blocks.idle_bytes -= bytes;
      g_blocks.live[*out] = std::make_pair(dev, bytes);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e != cudaSuccess) {  // give the idle blocks back to the driver and try once more
    cudaGetLastError();
    cache_trim();
    e = cudaMalloc(out, bytes);
  }
  if (e == cudaSuccess && big) {
    std::lock_guard<std::mutex> lk(g_blocks.mu);
    g_blocks.live[*out] = std::make_pair(dev, bytes);
  }
  return e;
}
}  // namespace

template <class T>
static cudaError_t dev_alloc(T **out, size_t bytes) {
  return dev_alloc_bytes(reinterpret_cast<void **>(out), bytes);
}
// the caller guarantees that no work using the block is still in flight (wendy_cuda_destroy synchronises)
static void dev_free(void *q) {
  if (!q) return;
  {
    std::lock_guard<std::mutex> lk(g_blocks.mu);
    auto it = g_blocks.live.find(q);
    if (it != g_blocks.live.end()) {
      const std::pair<int, size_t> key = it->second;
      g_blocks.live.erase(it);
      if (g_blocks.idle_bytes + key.second <= cache_limit_bytes()) {
        g_blocks.idle.insert(std::make_pair(key, q));
        g_blocks.idle_bytes += key.second;
        return;
      }
    }
  }
  cudaFree(q);
}

void wendy_cuda_trim(void) { cache_trim(); }

// ---- page-locked host ranges ------------------------------------------------------------------------------
// cudaHostRegister of a GB-sized range holds a driver lock for hundreds of milliseconds, during which every
// other CUDA call of the process waits (measured: the allocations of a handle being created on another thread
// took 495 ms instead of 90).  wendy_cuda_pin therefore registers a range piecewise, in chunks that end on
// absolute 16 MB address boundaries, and every large host<->device copy of this library is split at the same
// boundaries, so that each piece lies inside one registration and runs as a page-locked copy.
namespace {
uintptr_t pin_chunk() {  // WENDY_B200_PIN_CHUNK_MB overrides the piece size (A/B runs)
  static uintptr_t v = 0;
  if (!v) {
    const char *e = getenv("WENDY_B200_PIN_CHUNK_MB");
    const long mb = e ? atol(e) : 16;
    v = (uintptr_t)(mb > 0 ? mb : 16) << 20;
  }
  return v;
}
std::mutex g_pin_mu;
std::map<void *, std::vector<std::pair<void *, size_t>>> g_pins;
inline size_t pin_piece(const void *host, size_t left) {
  const uintptr_t a = (uintptr_t)host;
  const uintptr_t to_edge = pin_chunk() - (a % pin_chunk());
  return left < to_edge ? left : (size_t)to_edge;
}
cudaError_t copy_split(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
  char *d = (char *)dst;
  const char *s = (const char *)src;
  while (bytes) {
    const size_t piece = pin_piece(kind == cudaMemcpyDeviceToHost ? (const void *)d : (const void *)s, bytes);
    cudaError_t e = cudaMemcpyAsync(d, s, piece, kind, st);
    if (e != cudaSuccess) return e;
    d += piece; s += piece; bytes -= piece;
  }
  return cudaSuccess;
}
}  // namespace

// ---- bounce-buffered device -> host copies -------------------------------------------------------------------
// Page-locking the arrays a generator yields costs 350-450 ms per 1.6 GB on this platform (cudaHostRegister,
// profiles/r01/pcie_probe.txt) -- more than everything else in the set-up together.  A destination that is NOT
// page-locked is therefore filled through a small ring of page-locked bounce buffers: the copy engine writes
// piece i+1.. while all host threads copy piece i into the caller's array (the mirror image of
// upload_host_arrays).  Rings are pooled per process; a handle borrows one for its lifetime.
#ifndef WENDY_BOUNCE_NB
#define WENDY_BOUNCE_NB 4
#endif
#ifndef WENDY_BOUNCE_MB
#define WENDY_BOUNCE_MB 16
#endif
struct BounceRing {
  static constexpr int NB = WENDY_BOUNCE_NB;
  static constexpr size_t BYTES = (size_t)WENDY_BOUNCE_MB << 20;
  void *buf[NB] = {};
  cudaEvent_t ev[NB] = {};
  int device = 0;
};
namespace {
std::mutex g_ring_mu;
std::vector<BounceRing *> g_rings;
BounceRing *ring_acquire(int device) {
  {
    std::lock_guard<std::mutex> lk(g_ring_mu);
    for (size_t i = 0; i < g_rings.size(); i++)
      if (g_rings[i]->device == device) {
        BounceRing *r = g_rings[i];
        g_rings.erase(g_rings.begin() + i);
        return r;
      }
  }
  BounceRing *r = new BounceRing;
  r->device = device;
  for (int i = 0; i < BounceRing::NB; i++) {
    if (cudaMallocHost(&r->buf[i], BounceRing::BYTES) != cudaSuccess ||
        cudaEventCreateWithFlags(&r->ev[i], cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      for (int j = 0; j <= i; j++) {
        if (r->buf[j]) cudaFreeHost(r->buf[j]);
        if (r->ev[j]) cudaEventDestroy(r->ev[j]);
      }
      delete r;
      return nullptr;
    }
  }
  return r;
}
void ring_release(BounceRing *r) {
  if (!r) return;
  std::lock_guard<std::mutex> lk(g_ring_mu);
  g_rings.push_back(r);
}
bool host_byte_is_pinned(const void *q) {
  cudaPointerAttributes at;
  const bool pinned = cudaPointerGetAttributes(&at, q) == cudaSuccess && at.type == cudaMemoryTypeHost;
  cudaGetLastError();
  return pinned;
}
// first and last byte: wendy_cuda_pin registers a range piecewise in address order, possibly on another thread
// while read-outs go on, so a range counts as page-locked only once its last piece is
bool host_range_is_pinned(const void *q, size_t bytes) {
  return host_byte_is_pinned(q) && (bytes < 2 || host_byte_is_pinned((const char *)q + bytes - 1));
}
// Copy out of a bounce buffer with non-temporal stores: the destination is written once and not read back
// soon, so ordinary stores would first read every destination line into the cache (one third more memory
// traffic on the path that bounds the read-out).
inline void stream_copy(char *d, const char *s, size_t n) {
#if defined(__x86_64__) || defined(_M_X64)
  size_t head = (size_t)((16 - ((uintptr_t)d & 15)) & 15);
  if (head > n) head = n;
  memcpy(d, s, head);
  d += head; s += head; n -= head;
  const size_t nv = n / 64;
  for (size_t i = 0; i < nv; i++) {
    const __m128i a = _mm_loadu_si128((const __m128i *)s), b = _mm_loadu_si128((const __m128i *)(s + 16));
    const __m128i c = _mm_loadu_si128((const __m128i *)(s + 32)), e = _mm_loadu_si128((const __m128i *)(s + 48));
    _mm_stream_si128((__m128i *)d, a);
    _mm_stream_si128((__m128i *)(d + 16), b);
    _mm_stream_si128((__m128i *)(d + 32), c);
    _mm_stream_si128((__m128i *)(d + 48), e);
    s += 64; d += 64;
  }
  memcpy(d, s, n - nv * 64);
  _mm_sfence();
#else
  memcpy(d, s, n);
#endif
}
// dst (pageable host) <- src (device), through the ring, on stream st; returns when the data is in dst
int bounce_d2h(BounceRing *r, cudaStream_t st, void *const *dst, const void *const *src, const size_t *bytes, int narr) {
  struct Piece { char *d; const char *s; size_t n; };
  std::vector<Piece> pieces;
  for (int a = 0; a < narr; a++) {
    if (!dst[a]) continue;
    for (size_t off = 0; off < bytes[a]; off += BounceRing::BYTES)
      pieces.push_back({(char *)dst[a] + off, (const char *)src[a] + off, std::min(BounceRing::BYTES, bytes[a] - off)});
  }
  const int np = (int)pieces.size();
  int issued = 0;
  auto issue = [&](int i) -> cudaError_t {
    const int sl = i % BounceRing::NB;
    cudaError_t e = cudaMemcpyAsync(r->buf[sl], pieces[i].s, pieces[i].n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaEventRecord(r->ev[sl], st);
    return e;
  };
  for (; issued < np && issued < BounceRing::NB - 1; issued++)
    if (issue(issued) != cudaSuccess) return -1;
  for (int i = 0; i < np; i++) {
    // keep the copy engine NB-1 pieces ahead; the slot of piece i-1 was drained in the previous iteration
    if (issued < np) { if (issue(issued) != cudaSuccess) return -1; issued++; }
    const int sl = i % BounceRing::NB;
    if (cudaEventSynchronize(r->ev[sl]) != cudaSuccess) return -1;
    const char *bp = (const char *)r->buf[sl];
    char *dp = pieces[i].d;
    const long long nblk = (long long)((pieces[i].n + 262143) / 262144);
    // (dynamic: with one rank per GPU the cores are shared with the other ranks' threads; a static split would wait
    // for whichever thread was descheduled)
#pragma omp parallel for schedule(dynamic, 2) num_threads(host_threads())
    for (long long blk = 0; blk < nblk; blk++) {
      const size_t b0 = (size_t)blk * 262144, bl = std::min((size_t)262144, pieces[i].n - b0);
      stream_copy(dp + b0, bp + b0, bl);
    }
  }
  return 0;
}
}  // namespace

// Touch every page of a freshly allocated host array with a few threads (the contents are kept): the first
// read-out into untouched numpy memory otherwise pays the page faults of 16 bytes/particle inside its copy
// threads (159 ms instead of 42 ms at N=1e8).  The generator calls this on a helper thread during set-up.
extern "C" void wendy_host_prefault(void *host_ptr, unsigned long long bytes) {
  if (!host_ptr || !bytes) return;
  const size_t page = 4096;
  const long long np = (long long)((bytes + page - 1) / page);
  int nt = host_threads();
  if (nt > 8) nt = 8;
#pragma omp parallel for schedule(static) num_threads(nt)
  for (long long i = 0; i < np; i++) {
    volatile char *q = (volatile char *)host_ptr + (size_t)i * page;
    *q = *q;
  }
}

// test hook (tests/test_abi.py): the host-side copy used by the bounce-buffered read-out, multi-threaded
extern "C" void wendy_host_stream_copy(void *dst, const void *src, unsigned long long bytes) {
  const long long nblk = (long long)((bytes + 262143) / 262144);
#pragma omp parallel for schedule(static) num_threads(host_threads())
  for (long long blk = 0; blk < nblk; blk++) {
    const size_t b0 = (size_t)blk * 262144, bl = std::min((size_t)262144, (size_t)bytes - b0);
    stream_copy((char *)dst + b0, (const char *)src + b0, bl);
  }
}

// WENDY_B200_D2H=pinned keeps pageable destinations on the driver's own staging path (A/B runs)
static bool bounce_allowed() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("WENDY_B200_D2H"); v = !(e && e[0] == 'p'); }
  return v != 0;
}

// WENDY_B200_TRACE=1: host-side phase timings of set-up, layout builds and read-outs on stderr (each mark
// synchronises the stream, so the numbers are only meaningful for finding where the time goes)
static bool trace_on() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("WENDY_B200_TRACE"); v = (e && e[0] && e[0] != '0') ? 1 : 0; }
  return v != 0;
}
static void trace_mark(cudaStream_t st, const char *label) {
  if (!trace_on()) return;
  static std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
  cudaStreamSynchronize(st);
  const auto now = std::chrono::steady_clock::now();
  fprintf(stderr, "[wendy_b200 trace] %-44s %9.2f ms\n", label, 1e3 * std::chrono::duration<double>(now - last).count());
  last = now;
}
