// common.cuh -- device helpers shared by the wendy_b200 kernels (sm_100a only).
//
//  * order-preserving fp64 -> u64 key transform (radix sort keys)
//  * exact 128-bit fixed-point mass arithmetic: the cumulative mass of the force
//    (reference wendy/wendy.c:359-360 is a serial fp64 running sum) is computed here as
//    the CORRECTLY ROUNDED exact prefix sum, which makes it independent of tiling,
//    of the order in which tiles finish, of the sort path taken and of the GPU count.
//  * warp / block scan primitives on 128-bit integers
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>

typedef __int128 i128;
typedef unsigned __int128 u128;

#define WENDY_FULL_MASK 0xffffffffu

// Per-function attributes (dynamic shared-memory size, carve-out) are per DEVICE: a process that steps systems on
// two GPUs must set them on both.  flags: a static object owned by the call site (one cudaGetDevice per launch).
// Once-per-device set-up of a kernel (cudaFuncSetAttribute ...), safe when several host threads drive handles of one
// device (the ranks of a sharded system living in one process, tests/multi_helpers.py): a thread that finds the set-up
// under way waits for it instead of launching a kernel whose attributes are not set yet ("invalid argument").
//   static OnceFlags flags;  WENDY_ONCE_PER_DEVICE(flags) { cudaFuncSetAttribute(...); }
struct OnceFlags {
  std::atomic<int> done[64];
};
class OnceGuard {
 public:
  explicit OnceGuard(OnceFlags &f) : f_(f) {
    cudaGetDevice(&dev_);
    dev_ &= 63;
    if (f_.done[dev_].load(std::memory_order_acquire)) return;
    mutex().lock();
    locked_ = true;
    todo_ = f_.done[dev_].load(std::memory_order_relaxed) == 0;
  }
  ~OnceGuard() {
    if (locked_) mutex().unlock();
  }
  OnceGuard(const OnceGuard &) = delete;
  OnceGuard &operator=(const OnceGuard &) = delete;
  bool run() const { return todo_; }
  void done() {
    f_.done[dev_].store(1, std::memory_order_release);
    todo_ = false;
  }

 private:
  static std::mutex &mutex() {
    static std::mutex m;
    return m;
  }
  OnceFlags &f_;
  int dev_ = 0;
  bool locked_ = false, todo_ = false;
};
#define WENDY_ONCE_PER_DEVICE(flags) for (::OnceGuard once_guard_(flags); once_guard_.run(); once_guard_.done())

// ---------------------------------------------------------------------------------
// key transform: ascending u64 order == ascending fp64 order; -0.0 is canonicalised to
// +0.0 first because the reference comparator (wendy/wendy.c:21) treats them as equal.
__host__ __device__ __forceinline__ uint64_t key_from_double(double x) {
  x = x + 0.0;  // -0.0 + 0.0 == +0.0 under round-to-nearest
#ifdef __CUDA_ARCH__
  uint64_t b = (uint64_t)__double_as_longlong(x);
#else
  uint64_t b;
  memcpy(&b, &x, 8);
#endif
  uint64_t mask = (uint64_t)(-(int64_t)(b >> 63)) | 0x8000000000000000ull;
  return b ^ mask;
}

__host__ __device__ __forceinline__ double double_from_key(uint64_t k) {
  uint64_t mask = ((k >> 63) - 1ull) | 0x8000000000000000ull;
  uint64_t b = k ^ mask;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double x;
  memcpy(&x, &b, 8);
  return x;
#endif
}

// ---------------------------------------------------------------------------------
// fixed point: value * 2^E as a signed 128-bit integer.  E is chosen on the host so that
// sum(|m|) * 2^E < 2^124 (api.cu: choose_fx_exponent); bits below 2^-E (masses more than
// ~2^-70 times smaller than the total) are truncated towards zero.
__device__ __forceinline__ i128 fx_from_double(double m, int E) {
  long long bits = __double_as_longlong(m);
  int ex = (int)((bits >> 52) & 0x7ff);
  unsigned long long mant = (unsigned long long)bits & 0xFFFFFFFFFFFFFull;
  if (ex == 0) {
    if (mant == 0) return (i128)0;
    ex = 1;  // subnormal
  } else {
    mant |= (1ull << 52);
  }
  int sh = ex - 1075 + E;  // value = mant * 2^(ex-1075)
  u128 mag;
  if (sh >= 0) {
    if (sh > 73) sh = 73;  // unreachable for a valid E; keeps garbage input in range
    mag = (u128)mant << sh;
  } else {
    mag = (sh > -64) ? (u128)(mant >> (-sh)) : (u128)0;
  }
  return bits < 0 ? -(i128)mag : (i128)mag;
}

// round-to-nearest-even conversion back to fp64 (one rounding in total).
__device__ __forceinline__ double fx_to_double(i128 val, int E) {
  if (val == 0) return 0.0;
  bool neg = val < 0;
  u128 mag = neg ? (u128)(-val) : (u128)val;
  unsigned long long hi = (unsigned long long)(mag >> 64), lo = (unsigned long long)mag;
  unsigned long long top;
  int e2 = 0;
  if (hi) {
    int s = 64 - __clzll((long long)hi);  // 1..64 bits to drop so that the rest fits 64 bits
    top = (unsigned long long)(mag >> s);
    u128 dropped = mag & ((((u128)1) << s) - 1);
    if (dropped) top |= 1ull;  // sticky bit (bit 0 is far below the fp64 rounding position)
    e2 = s;
  } else {
    top = lo;
  }
  // scale by 2^(e2-E): one exact multiply when that power of two is a normal double
  const int k = e2 - E;
  double d = __ull2double_rn(top);
  if (k > -1000 && k < 1000) d *= __hiloint2double((1023 + k) << 20, 0);
  else d = scalbn(d, k);
  return neg ? -d : d;
}

// ---------------------------------------------------------------------------------
// shuffles / scans on 128-bit integers
__device__ __forceinline__ i128 shfl_up_i128(i128 v, int delta) {
  unsigned long long lo = (unsigned long long)v, hi = (unsigned long long)((u128)v >> 64);
  lo = __shfl_up_sync(WENDY_FULL_MASK, lo, delta);
  hi = __shfl_up_sync(WENDY_FULL_MASK, hi, delta);
  return (i128)(((u128)hi << 64) | lo);
}
__device__ __forceinline__ i128 shfl_i128(i128 v, int src) {
  unsigned long long lo = (unsigned long long)v, hi = (unsigned long long)((u128)v >> 64);
  lo = __shfl_sync(WENDY_FULL_MASK, lo, src);
  hi = __shfl_sync(WENDY_FULL_MASK, hi, src);
  return (i128)(((u128)hi << 64) | lo);
}
__device__ __forceinline__ i128 shfl_xor_i128(i128 v, int m) {
  unsigned long long lo = (unsigned long long)v, hi = (unsigned long long)((u128)v >> 64);
  lo = __shfl_xor_sync(WENDY_FULL_MASK, lo, m);
  hi = __shfl_xor_sync(WENDY_FULL_MASK, hi, m);
  return (i128)(((u128)hi << 64) | lo);
}
__device__ __forceinline__ i128 warp_inclusive_scan_i128(i128 v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    i128 u = shfl_up_i128(v, o);
    if (lane >= o) v += u;
  }
  return v;
}
__device__ __forceinline__ unsigned warp_inclusive_scan_u32(unsigned v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned u = __shfl_up_sync(WENDY_FULL_MASK, v, o);
    if (lane >= o) v += u;
  }
  return v;
}
// The same scan with the shuffle's own "source lane in range" predicate guarding the addition (no compare and
// select per step: three instructions per step instead of four)
__device__ __forceinline__ unsigned warp_inclusive_scan_u32_p(unsigned v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 t;\n\t"
        "shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t"
        "@p add.u32 %0, %0, t;\n\t"
        "}"
        : "+r"(v)
        : "r"(o));
  }
  return v;
}

// global loads that must not hit a stale L1 line (cross-CTA communication)
// ---- TMA bulk copies (global -> shared) completing on an mbarrier ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}

// per-thread asynchronous 8-byte copy global -> shared (LDGSTS): no register is held while it is in flight
__device__ __forceinline__ void cp_async_8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ~20-bit reciprocal (one MUFU + fix-up) for quantities that only steer a monotone binning or a search
// guess -- never for anything that reaches the particle state.
__device__ __forceinline__ double rcp_approx(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}

__device__ __forceinline__ float rcp_approx_f(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) {
  return *(const volatile unsigned *)p;
}
