// small.cu -- resident kernel for small systems: ONE CTA per system (segment), particle state in
// shared memory, all nleap leapfrog sub-steps of a reference call (wendy/wendy.c:385-418) in ONE launch.
//
// For N <= 1024 per system the bucket machinery is pointless: the whole system is one "bucket".  Per
// sub-step the CTA sorts by (x, id) with the same interpolation counting sort as the big kernels,
// computes the exact cumulative mass (equal masses: rank*m0; general: 128-bit fixed-point block scan),
// applies force / kick / drift in the reference's operation order, and loops -- no HBM traffic and no
// kernel launch between sub-steps (the reference's tests take up to 1e5 sub-steps per output on three
// particles).  State stays in particle-index order in HBM ("dense"), so slot == particle id.
#include <math_constants.h>

#include "common.cuh"
#include "internal.h"

namespace wendy {

template <int SCAP, int THREADS, int EQM>
struct SmallSmem {
  static constexpr int E = SCAP / THREADS;
  static constexpr int PADN = SCAP + SCAP / E + 4;
  double sx[SCAP];           // positions by particle slot
  double sv[SCAP];
  double sm[EQM ? 2 : SCAP];
  double skey[SCAP];         // keys in sub-bucket order (ranking reads them sequentially)
  unsigned cnt[PADN];        // sub-bucket counters -> offsets
  unsigned short slot[SCAP]; // particle slots grouped by sub-bucket
  double mcum[EQM ? 2 : PADN];  // masses in sorted order -> cumulative mass
  double dred[2][32];
  unsigned long long wlo[32], whi[32];
  unsigned uw[32];
};

template <int SCAP, int THREADS, int EQM>
__global__ void __launch_bounds__(THREADS)
small_kernel(double *__restrict__ x, double *__restrict__ v, const double *__restrict__ m, long long seg_len,
             const double *__restrict__ tot_seg, double m0, const SerialTab *__restrict__ stab, double omega2,
             int fxE, double dt, int nleap) {
  using SM = SmallSmem<SCAP, THREADS, EQM>;
  constexpr int E = SM::E;
  constexpr int NW = THREADS / 32;
  constexpr int BK = SCAP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM &S = *reinterpret_cast<SM *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned n = (unsigned)seg_len;
  const size_t base = (size_t)blockIdx.x * (size_t)seg_len;
  const double tot = tot_seg[blockIdx.x];

  for (unsigned i = tid; i < n; i += THREADS) {
    S.sx[i] = x[base + i];
    S.sv[i] = v[base + i];
    if (!EQM) S.sm[i] = m[base + i];
  }
  __syncthreads();

  for (int step = 0; step < nleap; step++) {
    // ---- keys: positions at force time (leading half drift in the first sub-step) -------------------
    double xk[E];
    double lmin = CUDART_INF, lmax = -CUDART_INF;
#pragma unroll
    for (int k = 0; k < E; k++) {
      const unsigned i = tid + k * THREADS;
      xk[k] = 0.0;
      if (i < n) {
        double xx = S.sx[i];
        if (step == 0) xx = __dadd_rn(xx, __dmul_rn(dt / 2., S.sv[i]));
        xk[k] = xx;
        lmin = fmin(lmin, xx);
        lmax = fmax(lmax, xx);
      }
    }
    for (int i = tid; i < SM::PADN; i += THREADS) S.cnt[i] = 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lmin = fmin(lmin, __shfl_xor_sync(WENDY_FULL_MASK, lmin, o));
      lmax = fmax(lmax, __shfl_xor_sync(WENDY_FULL_MASK, lmax, o));
    }
    if (lane == 0) {
      S.dred[0][wid] = lmin;
      S.dred[1][wid] = lmax;
    }
    __syncthreads();
    double xmin = S.dred[0][0], xmax = S.dred[1][0];
#pragma unroll
    for (int w = 1; w < NW; w++) {
      xmin = fmin(xmin, S.dred[0][w]);
      xmax = fmax(xmax, S.dred[1][w]);
    }
    const double range = xmax - xmin;
    const double scale = (range > 0.0 && range < CUDART_INF) ? (double)(BK - 1) / range : 0.0;
    // ---- interpolation counting sort --------------------------------------------------------------------
    unsigned pk[E];
#pragma unroll
    for (int k = 0; k < E; k++) {
      pk[k] = 0;
      if (tid + k * THREADS < n) {
        int sub = (int)((xk[k] - xmin) * scale);
        sub = max(0, min(BK - 1, sub));
        const unsigned o = atomicAdd(&S.cnt[sub + sub / E], 1u);
        pk[k] = (unsigned)sub | (o << 16);
      }
    }
    __syncthreads();
    {
      unsigned c[E], run = 0;
      unsigned *cp = &S.cnt[tid * (E + 1)];
#pragma unroll
      for (int q = 0; q < E; q++) {
        c[q] = cp[q];
        run += c[q];
      }
      const unsigned inc = warp_inclusive_scan_u32(run, lane);
      if (lane == 31) S.uw[wid] = inc;
      __syncthreads();
      if (wid == 0) {
        const unsigned t = lane < NW ? S.uw[lane] : 0u;
        const unsigned ti = warp_inclusive_scan_u32(t, lane);
        if (lane < NW) S.uw[lane] = ti - t;
      }
      __syncthreads();
      unsigned ex = inc - run + S.uw[wid];
#pragma unroll
      for (int q = 0; q < E; q++) {
        cp[q] = ex;
        ex += c[q];
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < E; k++) {
      const unsigned i = tid + k * THREADS;
      if (i < n) {
        const unsigned sub = pk[k] & 0xffffu;
        const unsigned pos = S.cnt[sub + sub / E] + (pk[k] >> 16);
        S.skey[pos] = xk[k];
        S.slot[pos] = (unsigned short)i;
      }
    }
    __syncthreads();
    // ---- exact rank under (x, id); slot == particle id in this layout --------------------------------
    unsigned r[E];
#pragma unroll
    for (int k = 0; k < E; k++) {
      const unsigned i = tid + k * THREADS;
      r[k] = 0;
      if (i < n) {
        const unsigned sub = pk[k] & 0xffffu;
        const unsigned s0 = S.cnt[sub + sub / E];
        const unsigned s1 = (sub + 1 < (unsigned)BK) ? S.cnt[(sub + 1) + (sub + 1) / E] : n;
        unsigned rr = s0;
        const double xi = xk[k];
#pragma unroll 1
        for (unsigned q = s0; q < s1; q++) {
          const double xj = S.skey[q];
          rr += (xj < xi || (xj == xi && (unsigned)S.slot[q] < i)) ? 1u : 0u;
        }
        r[k] = rr;
      }
    }
    // ---- cumulative mass below every particle ---------------------------------------------------------------
    double cum[E];
    if (EQM) {
#pragma unroll
      for (int k = 0; k < E; k++)  // serial table: the reference's own running sum, bit for bit (serialsum.cuh)
        cum[k] = stab ? serial_cum_at(stab, (long long)r[k]) : __dmul_rn((double)r[k], m0);
    } else {
#pragma unroll
      for (int k = 0; k < E; k++) {
        const unsigned i = tid + k * THREADS;
        if (i < n) S.mcum[r[k] + r[k] / E] = S.sm[i];
      }
      __syncthreads();
      i128 loc[E], tsum = 0;
      double *mp = &S.mcum[tid * (E + 1)];
#pragma unroll
      for (int q = 0; q < E; q++) {
        loc[q] = tsum;
        if ((unsigned)(tid * E + q) < n) tsum += fx_from_double(mp[q], fxE);
      }
      const i128 winc = warp_inclusive_scan_i128(tsum, lane);
      if (lane == 31) {
        S.wlo[wid] = (unsigned long long)winc;
        S.whi[wid] = (unsigned long long)((u128)winc >> 64);
      }
      __syncthreads();
      if (wid == 0) {
        const i128 t = lane < NW ? (i128)(((u128)S.whi[lane] << 64) | (u128)S.wlo[lane]) : (i128)0;
        const i128 tex = warp_inclusive_scan_i128(t, lane) - t;
        if (lane < NW) {
          S.wlo[lane] = (unsigned long long)tex;
          S.whi[lane] = (unsigned long long)((u128)tex >> 64);
        }
      }
      __syncthreads();
      const i128 bs = (i128)(((u128)S.whi[wid] << 64) | (u128)S.wlo[wid]) + (winc - tsum);
#pragma unroll
      for (int q = 0; q < E; q++)
        if ((unsigned)(tid * E + q) < n) mp[q] = fx_to_double(bs + loc[q], fxE);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < E; k++) {
        cum[k] = 0.0;
        if (tid + k * THREADS < n) cum[k] = S.mcum[r[k] + r[k] / E];
      }
    }
    // ---- force, kick, drift (wendy/wendy.c:375-383, 324-333) ------------------------------------------------
    const double dt_drift = (step == nleap - 1) ? dt / 2. : dt;
#pragma unroll
    for (int k = 0; k < E; k++) {
      const unsigned i = tid + k * THREADS;
      if (i < n) {
        const double mk = EQM ? m0 : S.sm[i];
        double acc = __dsub_rn(__dsub_rn(tot, __dmul_rn(2.0, cum[k])), mk);
        if (omega2 >= 0.0) acc = __dsub_rn(acc, __dmul_rn(omega2, xk[k]));
        const double v2 = __dadd_rn(S.sv[i], __dmul_rn(dt, acc));
        S.sv[i] = v2;
        S.sx[i] = __dadd_rn(xk[k], __dmul_rn(dt_drift, v2));
      }
    }
    __syncthreads();
  }
  for (unsigned i = tid; i < n; i += THREADS) {
    x[base + i] = S.sx[i];
    v[base + i] = S.sv[i];
  }
}

int small_max_particles() { return 1024; }

void launch_small(cudaStream_t st, double *x, double *v, const double *m, long long seg_len, int nseg,
                  const double *tot_seg, int eqm, double m0, const SerialTab *stab, double omega2, int fxE,
                  double dt, int nleap) {
  if (nseg <= 0 || seg_len <= 0) return;
  if (eqm) {
    const size_t sm = sizeof(SmallSmem<1024, 256, 1>);
    static OnceFlags set1;
    WENDY_ONCE_PER_DEVICE(set1) {
      cudaFuncSetAttribute(small_kernel<1024, 256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    }
    small_kernel<1024, 256, 1><<<nseg, 256, sm, st>>>(x, v, m, seg_len, tot_seg, m0, stab, omega2, fxE, dt, nleap);
  } else {
    const size_t sm = sizeof(SmallSmem<1024, 256, 0>);
    static OnceFlags set0;
    WENDY_ONCE_PER_DEVICE(set0) {
      cudaFuncSetAttribute(small_kernel<1024, 256, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    }
    small_kernel<1024, 256, 0><<<nseg, 256, sm, st>>>(x, v, m, seg_len, tot_seg, m0, nullptr, omega2, fxE, dt, nleap);
  }
}

}  // namespace wendy
