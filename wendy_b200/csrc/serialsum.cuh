// serialsum.cuh -- the reference's SERIAL cumulative-mass sum (wendy/wendy.c:359-360) in closed form
// for equal masses.
//
//   cum[0] = 0;  cum[i+1] = RN(cum[i] + m0)          (what the reference's loop computes when every m is m0)
//
// depends only on the sorted position i, not on the permutation, and although it is a serial recurrence it
// is piecewise LINEAR in i: while the running sum stays inside one binade [2^e, 2^(e+1)) and the exact sum
// c + m0 stays below 2^(e+1), every addition rounds on the same grid (spacing u = 2^(e-52)), so
// RN(c + m0) - c is the same multiple of u at every step (with round-half-even ties the increment settles
// after the first step inside the binade: the first result is even, and even + q*u + u/2 always rounds the
// same way).  The whole table for N < 2^31 is therefore ~3 linear pieces per binade of the running sum,
// about a hundred (i0, c0, inc) triples, and
//
//   cum[i] = c0 + (i - i0) * inc                      (both operations exact in fp64: every value is a
//                                                      multiple of u below 2^53 u)
//
// reproduces the reference BIT FOR BIT at any N -- including its accumulated rounding bias of ~N*2^-54
// (2.3e-9 relative at N=1e8), which the correctly rounded RN(i*m0) does not have.
// Validated against numpy.cumsum at N=1e8 (tests/test_serialsum.py through wendy_serial_cum).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace wendy {

constexpr int SS_MAX = 320;  // pieces: 3 per binade of the running sum plus the single steps of the first binades

struct SerialTab {
  int nseg;
  int first[64];          // first[b] = piece containing index 2^b (b = 0: index 0), where the search for k in [2^b, 2^(b+1)) starts
  long long i0[SS_MAX + 1];  // piece s covers sorted positions [i0[s], i0[s+1]); i0[nseg] = LLONG_MAX
  double c0[SS_MAX];
  double inc[SS_MAX];
};

// Host: build the table for positions [0, n_max).  Returns 0, or -1 if SS_MAX pieces do not suffice.
inline int serial_tab_build(SerialTab &T, double m0, long long n_max) {
  T.nseg = 0;
  auto push = [&](long long i0, double c0, double inc) -> bool {
    if (T.nseg >= SS_MAX) return false;
    T.i0[T.nseg] = i0; T.c0[T.nseg] = c0; T.inc[T.nseg] = inc;
    T.nseg++;
    return true;
  };
  long long i = 0;
  volatile double c = 0.0;  // volatile: plain fp64 additions, whatever the host compiler would like to contract
  const double am = fabs(m0);
  while (i < n_max) {
    volatile double c1 = c + m0, c2 = c1 + m0, c3 = c2 + m0;
    int e1, e2, e3;
    frexp(c1, &e1); frexp(c2, &e2); frexp(c3, &e3);
    if (c1 != 0.0 && e1 == e2 && e2 == e3 && isfinite(c3)) {
      // c1, c2, c3 share a binade: from c2 on the increment is steady while the exact sum stays below its top
      const double inc = c3 - c2;  // exact (same binade)
      // in units of u = 2^(e1-53):  |c2| = C2 u, |inc| = I u, top = 2^53 u; valid steps j = 0 .. n-1 need
      // C2 + j I + |m0|/u < 2^53  <=>  j I <= 2^53 - C2 - floor(|m0|/u) - 1
      const long long C2 = (long long)ldexp(fabs(c2), 53 - e1);
      const long long I = (long long)ldexp(fabs(inc), 53 - e1);
      const long long F = (long long)floor(ldexp(am, 53 - e1));
      const long long room = (1ll << 53) - C2 - F - 1;
      long long n;  // number of steady steps taken from c2
      if (I == 0) n = n_max;              // the sum has saturated (m0 below half a unit in the last place)
      else if (room < 0) n = 0;
      else n = room / I + 1;
      if (!push(i, c, 0.0) || !push(i + 1, c1, 0.0) || !push(i + 2, c2, inc)) return -1;
      if (n > n_max) n = n_max;
      // positions i+2 .. i+2+n hold c2 + j*inc; the step out of the last one crosses the binade: done serially
      const double clast = c2 + (double)n * inc;  // exact
      i = i + 2 + n + 1;
      c = clast + m0;
    } else {
      if (!push(i, c, 0.0)) return -1;
      i += 1;
      c = c1;
    }
  }
  T.i0[T.nseg] = 0x7fffffffffffffffll;
  int s = 0;
  for (int b = 0; b < 64; b++) {
    const long long k = (b == 0) ? 0ll : (b < 62) ? (1ll << b) : 0x7fffffffffffffffll - 1;  // (b = 0 serves k = 0 and 1)
    while (s + 1 < T.nseg && T.i0[s + 1] <= k) s++;
    T.first[b] = s;
  }
  return 0;
}

// Host evaluation (tests, small systems): cum below sorted position k.
inline double serial_tab_eval_host(const SerialTab &T, long long k) {
  int lo = 0, hi = T.nseg;
  while (hi - lo > 1) {
    const int mid = (lo + hi) / 2;
    if (T.i0[mid] <= k) lo = mid; else hi = mid;
  }
  volatile double prod = (double)(k - T.i0[lo]) * T.inc[lo];
  return T.c0[lo] + prod;
}

#ifdef __CUDACC__
// piece containing position k >= 0
__device__ __forceinline__ int serial_find(const SerialTab *__restrict__ T, long long k) {
  int s = __ldg(&T->first[63 - __clzll(k | 1ll)]);
  while (k >= __ldg(&T->i0[s + 1])) s++;
  return s;
}
__device__ __forceinline__ double serial_cum_at(const SerialTab *__restrict__ T, long long k) {
  const int s = serial_find(T, k);
  return __dadd_rn(__ldg(&T->c0[s]), __dmul_rn((double)(k - __ldg(&T->i0[s])), __ldg(&T->inc[s])));
}
// Per-bucket form: positions [k0, k0 + n) nearly always lie in ONE piece; then cum(k0 + r) = c0 + (j0 + r) * inc.
struct SerialRun {
  double c0, inc;
  unsigned j0;   // k0 - i0 of the piece (below 2^31)
  bool uniform;  // all n positions in the same piece
};
__device__ __forceinline__ SerialRun serial_run(const SerialTab *__restrict__ T, long long k0, unsigned n) {
  SerialRun R;
  const int s = serial_find(T, k0);
  const long long i0 = __ldg(&T->i0[s]);
  R.c0 = __ldg(&T->c0[s]);
  R.inc = __ldg(&T->inc[s]);
  R.j0 = (unsigned)(k0 - i0);
  R.uniform = (k0 + (long long)n <= __ldg(&T->i0[s + 1])) && (k0 - i0 + (long long)n < (1ll << 31));
  return R;
}
__device__ __forceinline__ double serial_cum_run(const SerialRun &R, unsigned r) {
  return __dadd_rn(R.c0, __dmul_rn((double)(R.j0 + r), R.inc));
}
#endif

}  // namespace wendy
