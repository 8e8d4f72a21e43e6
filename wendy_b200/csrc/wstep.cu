// wstep.cu -- the steady-state hot kernel: one leapfrog sub-step, ONE WARP PER BUCKET.
//
// Same algorithm as tile_kernel<LOAD_BUCKET, EMIT_SPLITTER> (tile.cu), re-shaped for the
// SM: a bucket holds at most WCAP = 256 particles and is owned by a single warp, so the
// whole sub-step -- stage, sort, scan, force, kick, drift, re-bucket -- needs no block barrier.
// 32 independent warps per SM sit in different phases and hide each other's latencies.
//
//   stage   lane 0 issues TMA bulk copies (cp.async.bulk, mbarrier complete_tx) of the bucket's
//           x, v, id (and m) segments into the warp's private shared-memory slab
//   prefix  while the copies fly: count look-back (packed 64-bit words) and, for general
//           masses, the order-independent bucket mass is published early (lookback.cuh)
//   sort    interpolation counting sort on the key range [split[b], split[b+1]) + exact rank
//           under (x, id) by comparison inside a sub-bucket                    (wendy.c:341-357)
//   scan    equal masses: cum = RN(rank * m0); general: exact 128-bit warp scan (wendy.c:359-360)
//   step    a = a_ext + (((M - 2 cum) - m) - omega^2 x); v += dt a; x += dt v   (wendy.c:375-383,324-333)
//   emit    destination bucket by splitter search (registers + shuffles for the 31 nearest
//           buckets, galloping search beyond), slots by warp-aggregated atomics, stores
#include <math_constants.h>

#include "common.cuh"
#include "internal.h"
#include "lookback.cuh"

#ifndef WS_WARPS
#define WS_WARPS 8
#endif

namespace wendy {

template <int WCAP, int EQM>
struct __align__(16) WarpSlab {
  static constexpr int E = WCAP / 32;
  static constexpr int PADN = WCAP + WCAP / E + 8;
  double sx[WCAP];
  double sv[WCAP];
  double sm[EQM ? 2 : WCAP];   // masses by load slot
  double so[EQM ? 2 : PADN];   // masses in sorted order -> cumulative mass (padded)
  int sid[WCAP];
  union {
    unsigned cnt[PADN];      // sort phase: sub-bucket counters -> start offsets
    struct {                 // emit phase (the counters are dead by then)
      double wsp[33];        // lower splitters of the 32 buckets around the home bucket (+ upper end)
      unsigned dcnt[32], dbase[32];
    } w;
  };
  unsigned short slot[WCAP];
  unsigned long long mbar;
};

// largest d in [lo0, hi0) with split[d] <= key, starting from a guess (split[lo0] is -inf)
__device__ __noinline__ int gallop_search(const double *__restrict__ split, double key, int guess,
                                             int lo0, int hi0) {
  int lo = min(max(guess, lo0), hi0 - 1), hi;
  int step = 1;
  if (__ldg(split + lo) <= key) {
    hi = lo + 1;
    while (hi < hi0 && __ldg(split + hi) <= key) {
      lo = hi;
      step <<= 1;
      hi = lo + step;
    }
    if (hi > hi0) hi = hi0;
  } else {
    hi = lo;
    lo = hi - 1;
    while (lo > lo0 && __ldg(split + lo) > key) {
      hi = lo;
      step <<= 1;
      lo = hi - step;
    }
    if (lo < lo0) lo = lo0;
  }
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(split + mid) <= key) lo = mid; else hi = mid;
  }
  return lo;
}

// destination bucket of a key beyond the 32-bucket window: guess from the home bucket's width, then gallop
__device__ __noinline__ int far_destination(const double *__restrict__ split, double key, double home_lo,
                                            double inv_w, int b, int seg_lo, int seg_hi) {
  double gq = (key - home_lo) * inv_w;
  gq = fmax(-2.0e9, fmin(2.0e9, gq));
  const int g = (int)max((long long)seg_lo, min((long long)seg_hi - 1, (long long)b + (long long)floor(gq)));
  return gallop_search(split, key, g, seg_lo, seg_hi);
}

// rank correction for exact coincidences inside a sub-bucket: members with the same key and a smaller id
template <typename SL>
__device__ __noinline__ unsigned tie_rank(const SL &S, unsigned s0, unsigned s1, double xi, int ii) {
  unsigned r = 0;
  for (unsigned q = s0; q < s1; q++)
    if (S.sx[q] == xi) r += (S.sid[S.slot[q]] < ii) ? 1u : 0u;
  return r;
}

template <int WCAP, int WARPS, int EQM, int SHARD>
__global__ void __launch_bounds__(WARPS * 32, 1024 / (WARPS * 32))
wstep_kernel(const TileParams p) {
  using SL = WarpSlab<WCAP, EQM>;
  constexpr int E = SL::E;
  constexpr int BK = WCAP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned s_ticket;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  SL &S = reinterpret_cast<SL *>(smem_raw)[wid];

  if (ld_volatile_u32(p.fail_seq) < p.seq) return;
  if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1u);
  __syncthreads();  // the only block-wide barrier
  const int b = (int)s_ticket * WARPS + wid;
  if (b >= p.nb) return;
  const int seg = (p.nbps == p.nb) ? 0 : b / p.nbps;
  const int seg_lo = seg * p.nbps, seg_hi = seg_lo + p.nbps;

  unsigned n = p.cnt_in[b];
  if (n > (unsigned)WCAP) {
    n = WCAP;
    if (lane == 0) atomicMin(p.fail_seq, p.seq);
  }
  const size_t base = (size_t)b * WCAP;
  const uint32_t bar = smem_u32(&S.mbar);
  // ---- stage: TMA bulk copies of the live part of the bucket --------------------------------
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (b == 0 && p.ticket_zero) *p.ticket_zero = 0;
    if (p.cnt_zero) p.cnt_zero[b] = 0;
    if (n > (unsigned)(WCAP - WCAP / 16)) atomicMax(p.stats, n);
    if (n) {
      const uint32_t b8 = (n * 8u + 15u) & ~15u, b4 = (n * 4u + 15u) & ~15u;
      const uint32_t total = b8 * (EQM ? 2u : 3u) + b4;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(total) : "memory");
      tma_load_1d(S.sx, p.xin + base, b8, bar);
      tma_load_1d(S.sv, p.vin + base, b8, bar);
      if (!EQM) tma_load_1d(S.sm, p.min + base, b8, bar);
      tma_load_1d(S.sid, p.idin + base, b4, bar);
    }
  }
#pragma unroll
  for (int i = lane; i < SL::PADN; i += 32) S.cnt[i] = 0;
  // particles in the preceding buckets of the segment (count_prefix kernel ran just before)
  const long long Pc = (long long)p.cpre[b] - (long long)seg * p.seg_len + (SHARD ? p.pc_offset : 0ll);
  SerialRun SR;
  SR.c0 = 0.0; SR.inc = 0.0; SR.j0 = 0u; SR.uniform = true;
  if (EQM && p.stab) SR = serial_run(p.stab, Pc, n);
  // lower splitters of the 32 buckets around b, in shared memory (only leaving lanes search)
  int wlo = b - 15;
  if (wlo > seg_hi - 32) wlo = seg_hi - 32;
  if (wlo < seg_lo) wlo = seg_lo;
  const double wsp_reg = (wlo + lane < seg_hi) ? __ldg(p.split + wlo + lane) : CUDART_INF;
  const double wsp_end = (wlo + 32 < seg_hi) ? __ldg(p.split + wlo + 32) : CUDART_INF;
  const double home_lo = __ldg(p.split + b);
  const double home_hi = (b + 1 < seg_hi) ? __ldg(p.split + b + 1) : CUDART_INF;
  const double tot = p.tot[seg];
  __syncwarp();
  if (n == 0) {
    if (!EQM && !p.mpre) {
      i128 P; long long pc2;
      lookback(p.desc, p.status, p.epoch, b, seg_lo, (i128)0, 0ll, lane, P, pc2);
    }
    return;
  }
  mbar_wait(bar, 0);

  // ---- keys: position at force time; key range -------------------------------------------------
  // key range of THIS bucket in the input layout (its own edges may since have been advected)
  double xmin = __ldg(p.split_in + b), xmax = (b + 1 < seg_hi) ? __ldg(p.split_in + b + 1) : CUDART_INF;
  if (p.h_pre != 0.0) {
#pragma unroll
    for (int k = 0; k < E; k++) {
      const unsigned i = lane + 32 * k;
      if (i < n) S.sx[i] = __dadd_rn(S.sx[i], __dmul_rn(p.h_pre, S.sv[i]));
    }
  }
  if (!(xmin > -CUDART_INF && xmax < CUDART_INF)) {  // edge bucket: reduce min / max in the warp
    double lmin = CUDART_INF, lmax = -CUDART_INF;
#pragma unroll
    for (int k = 0; k < E; k++) {
      const unsigned i = lane + 32 * k;
      if (i < n) {
        lmin = fmin(lmin, S.sx[i]);
        lmax = fmax(lmax, S.sx[i]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lmin = fmin(lmin, __shfl_xor_sync(WENDY_FULL_MASK, lmin, o));
      lmax = fmax(lmax, __shfl_xor_sync(WENDY_FULL_MASK, lmax, o));
    }
    xmin = lmin;
    xmax = lmax;
  }
  const double range = xmax - xmin;
  const double scale = (range > 0.0 && range < CUDART_INF) ? (double)(BK - 1) * rcp_approx(range) : 0.0;

  // general masses: the bucket's mass does not depend on the order -> publish it now and
  // resolve the prefix while the sort runs on the other warps
  i128 P = 0;
  if (!EQM) {
    if (p.mpre) {  // exact bucket-mass prefix precomputed by bucket_mass + mass_prefix kernels
      const ulonglong2 a = p.mpre[b], z = p.mpre[seg_lo];
      P = make_i128(a.x, a.y) - make_i128(z.x, z.y);
    } else {
      i128 agg = 0;
#pragma unroll
      for (int k = 0; k < E; k++) {
        const unsigned i = lane + 32 * k;
        if (i < n) agg += fx_from_double(S.sm[i], p.fxE);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) agg += shfl_xor_i128(agg, o);
      long long pc2;
      lookback(p.desc, p.status, p.epoch, b, seg_lo, agg, (long long)n, lane, P, pc2);
    }
  }

  // ---- sort: interpolation sub-bucket, arrival slot ------------------------------------------------
  // pk[k] packs, per particle: sub-bucket (bits 0-7), position in sub-bucket order (8-15),
  // exact rank (16-23); all < 256 = WCAP.
  static_assert(WCAP <= 256, "packed sort state assumes 8-bit indices");
  unsigned pk[E];
  double xr[E];
#pragma unroll
  for (int k = 0; k < E; k++) {
    const unsigned i = lane + 32 * k;
    pk[k] = 0;
    xr[k] = 0.0;
    if (32u * k < n && i < n) {
      xr[k] = S.sx[i];
      int sub = (int)((xr[k] - xmin) * scale);
      sub = max(0, min(BK - 1, sub));
      unsigned o = atomicAdd(&S.cnt[sub + sub / E], 1u);
      pk[k] = (unsigned)sub | (o << 8);
    }
  }
  __syncwarp();
  {  // exclusive scan of the sub-bucket counters: E consecutive (padded) entries per lane
    unsigned c[E], run = 0;
    unsigned *cp = &S.cnt[lane * (E + 1)];
#pragma unroll
    for (int q = 0; q < E; q++) {
      c[q] = cp[q];
      run += c[q];
    }
    unsigned ex = warp_inclusive_scan_u32(run, lane) - run;
#pragma unroll
    for (int q = 0; q < E; q++) {
      cp[q] = ex;
      ex += c[q];
    }
  }
  __syncwarp();
  // keys move (through registers) into sub-bucket order IN PLACE: sx[pos] = key, slot[pos] = load slot
#pragma unroll
  for (int k = 0; k < E; k++) {
    const unsigned i = lane + 32 * k;
    if (32u * k < n && i < n) {
      const unsigned sub = pk[k] & 0xffu;
      const unsigned pos = S.cnt[sub + sub / E] + (pk[k] >> 8);
      S.sx[pos] = xr[k];
      S.slot[pos] = (unsigned short)i;
      pk[k] = sub | (pos << 8);
    }
  }
  __syncwarp();
  // exact rank under (x, id): sub-bucket start + members that compare smaller
#pragma unroll
  for (int k = 0; k < E; k++) {
    const unsigned i = lane + 32 * k;
    if (32u * k < n && i < n) {
      const unsigned sub = pk[k] & 0xffu;
      const unsigned s0 = S.cnt[sub + sub / E];
      const unsigned s1 = (sub + 1 < (unsigned)BK) ? S.cnt[(sub + 1) + (sub + 1) / E] : n;
      unsigned rr = s0;
      if (s1 - s0 > 1u) {
        const double xi = xr[k];
        unsigned eq = 0;
        const unsigned c = s1 - s0;
        // sub-buckets hold ~1.5 members on average: the first four are handled by predicated
        // straight-line code, a loop only runs for crowded sub-buckets
#pragma unroll
        for (unsigned j = 0; j < 4; j++) {
          if (j < c) {
            const double xj = S.sx[s0 + j];
            rr += (xj < xi) ? 1u : 0u;
            eq += (xj == xi) ? 1u : 0u;
          }
        }
        if (c > 4u) {
#pragma unroll 1
          for (unsigned q = s0 + 4; q < s1; q++) {
            const double xj = S.sx[q];
            rr += (xj < xi) ? 1u : 0u;
            eq += (xj == xi) ? 1u : 0u;
          }
        }
        // every particle ties with itself; real coincidences are ordered by particle index (rare path)
        if (eq > 1u) rr += tie_rank(S, s0, s1, xi, S.sid[i]);
      }
      pk[k] |= rr << 16;
    }
  }
#define WS_POS(k) ((pk[k] >> 8) & 0xffu)
#define WS_RANK(k) (pk[k] >> 16)
  // ---- scan (general masses): exact 128-bit prefix in sorted order ------------------------------------
  if (!EQM) {
#pragma unroll
    for (int k = 0; k < E; k++) {
      const unsigned i = lane + 32 * k;
      if (i < n) S.so[WS_RANK(k) + WS_RANK(k) / E] = S.sm[i];
    }
    __syncwarp();
    i128 loc[E], tsum = 0;
    double *mp = &S.so[lane * (E + 1)];
#pragma unroll
    for (int q = 0; q < E; q++) {
      loc[q] = tsum;
      if ((unsigned)(lane * E + q) < n) tsum += fx_from_double(mp[q], p.fxE);
    }
    const i128 bs = P + (warp_inclusive_scan_i128(tsum, lane) - tsum);
#pragma unroll
    for (int q = 0; q < E; q++)
      if ((unsigned)(lane * E + q) < n) mp[q] = fx_to_double(bs + loc[q], p.fxE);
    __syncwarp();
  }
  // ---- step + destination -------------------------------------------------------------------------------
  __syncwarp();  // the sort counters are dead: their storage becomes the emit window
  S.w.wsp[lane] = wsp_reg;
  if (lane == 0) S.w.wsp[32] = wsp_end;
  S.w.dcnt[lane] = 0;
  __syncwarp();
  int dest[E];
  unsigned hoff[E];
  unsigned hc = 0, outside = 0;
  double dsum = 0.0;  // advection statistic: total key displacement of this bucket's particles
  unsigned dcount = 0;
  bool sh_overflow = false;
  const double sh_lo = SHARD ? __ldg(p.bounds + p.my_rank) : 0.0;
  const double sh_hi = SHARD ? __ldg(p.bounds + p.my_rank + 1) : 0.0;
  const double wdt = home_hi - home_lo;
  const double inv_w = (wdt > 0.0 && wdt < CUDART_INF) ? rcp_approx(wdt) : 0.0;
#pragma unroll
  for (int k = 0; k < E; k++) {
    dest[k] = -1;
    hoff[k] = 0;
    if (32u * k < n) {  // warp-uniform: rounds past the live part of the bucket are skipped
      const unsigned i = lane + 32 * k;
      const bool ok = i < n;
      int d = -1;
      if (ok) {
        const unsigned ps = WS_POS(k);
        const double xk = S.sx[ps], v = S.sv[i];
        double c, mk;
        if (EQM) {
          mk = p.m0;
          // Pc includes the lower ranks' particles; serial table: the reference's own running sum (serialsum.cuh)
          if (p.stab) c = SR.uniform ? serial_cum_run(SR, WS_RANK(k)) : serial_cum_at(p.stab, Pc + (long long)WS_RANK(k));
          else c = __dmul_rn((double)(Pc + (long long)WS_RANK(k)), p.m0);
        } else {
          mk = S.sm[i];
          c = S.so[WS_RANK(k) + WS_RANK(k) / E];
        }
        double acc = __dsub_rn(__dsub_rn(tot, __dmul_rn(2.0, c)), mk);
        if (p.omega2 >= 0.0) acc = __dsub_rn(acc, __dmul_rn(p.omega2, xk));
        if (p.aext) acc = __dadd_rn(p.aext[base + i], acc);
        const double v2 = __dadd_rn(v, __dmul_rn(p.dt_kick, acc));
        const double x2 = __dadd_rn(xk, __dmul_rn(p.dt_drift, v2));
        const double key = (p.h_next != 0.0) ? __dadd_rn(x2, __dmul_rn(p.h_next, v2)) : x2;
        S.sx[ps] = x2;
        S.sv[i] = v2;
        if (p.rank_out) p.rank_out[S.sid[i]] = (int)(Pc + (long long)WS_RANK(k));
        dsum += key - xk;
        dcount++;
        if (SHARD && (key < sh_lo || key >= sh_hi)) {
          // leaves this GPU's key range: append to the outbox of the rank that owns the key
          int peer = 0;
          while (peer + 1 < p.nranks && key >= __ldg(p.bounds + peer + 1)) peer++;
          const unsigned slot = atomicAdd(p.out_cnt + peer, 1u);
          if (slot < p.ocap) {
            double *rec = p.out_rec + ((size_t)peer * p.ocap + slot) * 3;  // packed (x, v, id) record
            rec[0] = x2;
            rec[1] = v2;
            rec[2] = (double)S.sid[i];
          } else {
            sh_overflow = true;
          }
          d = -3;
        } else if (key >= home_lo && key < home_hi) {
          d = b;
        } else if (key >= S.w.wsp[0] && key < S.w.wsp[32]) {  // one of the 32 nearby buckets
          // interpolated guess from the home bucket's width, then a short walk to the exact bucket
          int lo = (b - wlo) + (int)floor(fmax(-64.0, fmin(64.0, (key - home_lo) * inv_w)));
          lo = max(0, min(31, lo));
#pragma unroll 1
          while (lo > 0 && S.w.wsp[lo] > key) lo--;
#pragma unroll 1
          while (lo < 31 && S.w.wsp[lo + 1] <= key) lo++;
          d = wlo + lo;
          hoff[k] = atomicAdd(&S.w.dcnt[lo], 1u);
        } else {  // beyond the window: interpolated guess, then galloping search (rare path, not inlined)
          d = far_destination(p.split, key, home_lo, inv_w, b, seg_lo, seg_hi);
          hoff[k] = atomicAdd(&p.cnt_out[d], 1u);  // final slot
          outside++;
        }
      }
      dest[k] = d;
      const unsigned home = __ballot_sync(WENDY_FULL_MASK, d == b);
      if (d == b) hoff[k] = hc + __popc(home & lt);
      hc += __popc(home);
    }
  }
  if (p.knot_sum) {  // mean flow per cell of buckets, used to advect the splitters before the next sub-step
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dsum += __shfl_xor_sync(WENDY_FULL_MASK, dsum, o);
      dcount += __shfl_xor_sync(WENDY_FULL_MASK, dcount, o);
    }
    if (lane == 0 && dcount) {
      atomicAdd(p.knot_sum + b / p.knot_g, dsum);
      atomicAdd(p.knot_n + b / p.knot_g, dcount);
    }
  }
  // ---- slots: one global atomic per destination bucket of the window, all in one round trip ----------
  __syncwarp();
  {
    unsigned c = S.w.dcnt[lane];
    if (lane == b - wlo) c += hc;  // in-window away particles were counted first: home goes after
    if (c) S.w.dbase[lane] = atomicAdd(&p.cnt_out[wlo + lane], c);
  }
  {  // one atomic per warp, spread over 64 counters (a single hot address serialises in L2)
    const unsigned wsum = __reduce_add_sync(WENDY_FULL_MASK, outside);
    if (lane == 0 && wsum) atomicAdd(p.outside + (b & 63), (unsigned long long)wsum);
  }
  __syncwarp();
  const unsigned home_shift = S.w.dcnt[b - wlo];
  // ---- stores ----------------------------------------------------------------------------------------------
  bool overflow = false;
#pragma unroll
  for (int k = 0; k < E; k++) {
    if (32u * k < n) {
      const unsigned i = lane + 32 * k;
      const int d = dest[k];
      if (d >= 0) {
        unsigned pos = hoff[k];
        if (d == b) pos += S.w.dbase[b - wlo] + home_shift;
        else if (d >= wlo && d < wlo + 32) pos += S.w.dbase[d - wlo];
        if (pos < (unsigned)WCAP) {
          const size_t o = (size_t)d * WCAP + pos;
          p.xout[o] = S.sx[WS_POS(k)];
          p.vout[o] = S.sv[i];
          if (!EQM) p.mout[o] = S.sm[i];
          p.idout[o] = S.sid[i];
        } else {
          overflow = true;
        }
      }
    }
  }
  if (overflow || sh_overflow) atomicMin(p.fail_seq, p.seq);
#undef WS_POS
#undef WS_RANK
}

// ---- exclusive prefix of the bucket counts: one pass, look-back over CTA tiles ---------------------------
constexpr int CP_T = 1024, CP_I = 4, CP_TILE = CP_T * CP_I;
__global__ void __launch_bounds__(CP_T)
count_prefix_kernel(const unsigned *__restrict__ cnt, int nb, unsigned *__restrict__ cpre,
                    unsigned long long *desc, unsigned *ticket, unsigned epoch) {
  __shared__ unsigned s_t, s_w[32], s_pre;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_t = atomicAdd(ticket, 1u);
  __syncthreads();
  const int t = (int)s_t;
  if (t == (int)gridDim.x - 1 && threadIdx.x == 0) *ticket = 0;  // every ticket is taken: re-arm
  const int i0 = t * CP_TILE + threadIdx.x * CP_I;
  unsigned c[CP_I], run = 0;
#pragma unroll
  for (int q = 0; q < CP_I; q++) {
    c[q] = (i0 + q < nb) ? cnt[i0 + q] : 0u;
    run += c[q];
  }
  const unsigned inc = warp_inclusive_scan_u32(run, lane);
  if (lane == 31) s_w[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    const unsigned w = s_w[lane];
    const unsigned wi = warp_inclusive_scan_u32(w, lane);
    s_w[lane] = wi - w;
    const unsigned total = __shfl_sync(WENDY_FULL_MASK, wi, 31);
    const unsigned pre = count_lookback(desc, epoch, t, 0, total, lane);
    if (lane == 0) s_pre = pre;
  }
  __syncthreads();
  unsigned ex = s_pre + s_w[wid] + inc - run;
#pragma unroll
  for (int q = 0; q < CP_I; q++) {
    if (i0 + q < nb) cpre[i0 + q] = ex;
    ex += c[q];
  }
}

int count_prefix_tiles(int nb) { return (nb + CP_TILE - 1) / CP_TILE; }

// ---- Lagrangian splitters ---------------------------------------------------------------------------------------
// Coherent flows (cold collapse, bulk translation) carry whole regions across fixed bucket edges.  The step
// kernels therefore measure the mean key displacement per cell of G buckets; before the next sub-step the
// edges are moved by the piecewise-linear map through the cell-centre knots (X_j -> Y_j = cummax(X_j + D_j)),
// which is monotone by construction.  The layout only affects speed, never results.
__global__ void __launch_bounds__(1024)
advect_knots_kernel(const double *__restrict__ split_old, int nb, int G, double *__restrict__ knot_sum,
                    unsigned *__restrict__ knot_n, double *__restrict__ knot_x, double *__restrict__ knot_y,
                    int ncell) {
  __shared__ double wmax[32];
  __shared__ double carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = -CUDART_INF;
  __syncthreads();
  for (int base = 0; base < ncell; base += 1024) {
    const int j = base + threadIdx.x;
    double y = -CUDART_INF;
    if (j < ncell) {
      const int c = min(j * G + G / 2, nb - 1);
      const double X = split_old[c];
      const unsigned n = knot_n[j];
      const double D = n ? knot_sum[j] / (double)n : 0.0;
      knot_sum[j] = 0.0;
      knot_n[j] = 0u;
      knot_x[j] = X;
      y = (X > -CUDART_INF && X < CUDART_INF) ? X + D : X;
    }
    // inclusive running maximum over the cells
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_up_sync(WENDY_FULL_MASK, y, o);
      if (lane >= o) y = fmax(y, u);
    }
    if (lane == 31) wmax[wid] = y;
    __syncthreads();
    if (wid == 0) {
      double t = wmax[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(WENDY_FULL_MASK, t, o);
        if (lane >= o) t = fmax(t, u);
      }
      wmax[lane] = t;
    }
    __syncthreads();
    double pre = carry;
    if (wid > 0) pre = fmax(pre, wmax[wid - 1]);
    y = fmax(y, pre);
    if (j < ncell) knot_y[j] = y;
    __syncthreads();
    if (threadIdx.x == 1023) carry = y;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
advect_apply_kernel(const double *__restrict__ split_old, double *__restrict__ split_new, int nb, int G,
                    const double *__restrict__ knot_x, const double *__restrict__ knot_y, int ncell) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const double s = split_old[b];
  double r = s;
  if (s > -CUDART_INF && s < CUDART_INF) {
    const int j = (b >= G / 2) ? (b - G / 2) / G : -1;  // knots sit at bucket jG + G/2
    if (j < 0) {
      const double X0 = knot_x[0], Y0 = knot_y[0];
      if (X0 > -CUDART_INF && X0 < CUDART_INF) r = fmin(s + (Y0 - X0), Y0);
    } else {
      const double Xj = knot_x[j], Yj = knot_y[j];
      const bool right_ok = (j + 1 < ncell) && knot_x[j + 1] < CUDART_INF;
      if (!(Xj > -CUDART_INF && Xj < CUDART_INF)) {
        r = s;
      } else if (!right_ok) {
        r = fmax(s + (Yj - Xj), Yj);
      } else {
        const double Xn = knot_x[j + 1], Yn = knot_y[j + 1];
        double t = (Xn > Xj) ? (s - Xj) / (Xn - Xj) : 0.0;
        t = fmin(1.0, fmax(0.0, t));
        r = fmin(Yn, fmax(Yj, Yj + t * (Yn - Yj)));
      }
    }
  }
  split_new[b] = r;
}

void launch_advect_splitters(cudaStream_t st, const double *split_old, double *split_new, int nb, int G,
                             double *knot_sum, unsigned *knot_n, double *knot_x, double *knot_y) {
  if (nb <= 0) return;
  const int ncell = (nb + G - 1) / G;
  advect_knots_kernel<<<1, 1024, 0, st>>>(split_old, nb, G, knot_sum, knot_n, knot_x, knot_y, ncell);
  advect_apply_kernel<<<(nb + 255) / 256, 256, 0, st>>>(split_old, split_new, nb, G, knot_x, knot_y, ncell);
}

// ---- general masses: exact 128-bit mass of every bucket, then its exclusive prefix ---------------------------
// (one extra coalesced read of m per sub-step; replaces a decoupled look-back whose chains are long at
// one-warp-per-bucket granularity)
__global__ void __launch_bounds__(256)
bucket_mass_kernel(const double *__restrict__ m, const unsigned *__restrict__ cnt, int cap, int nb, int fxE,
                   ulonglong2 *__restrict__ magg) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= nb) return;
  const unsigned n = min(cnt[b], (unsigned)cap);
  const double *mp = m + (size_t)b * cap;
  i128 agg = 0;
  for (unsigned i = lane; i < n; i += 32) agg += fx_from_double(mp[i], fxE);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) agg += shfl_xor_i128(agg, o);
  if (lane == 0) magg[b] = make_ulonglong2((unsigned long long)agg, (unsigned long long)((u128)agg >> 64));
}

constexpr int MP_T = 256, MP_I = 4, MP_TILE = MP_T * MP_I;
__global__ void __launch_bounds__(MP_T)
mass_prefix_kernel(const ulonglong2 *__restrict__ magg, int nb, ulonglong2 *__restrict__ mpre, Desc *desc,
                   unsigned *status, unsigned *ticket, unsigned epoch) {
  __shared__ unsigned s_t;
  __shared__ unsigned long long s_lo[8], s_hi[8], s_plo, s_phi;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_t = atomicAdd(ticket, 1u);
  __syncthreads();
  const int t = (int)s_t;
  if (t == (int)gridDim.x - 1 && threadIdx.x == 0) *ticket = 0;
  const int i0 = t * MP_TILE + threadIdx.x * MP_I;
  i128 loc[MP_I], run = 0;
#pragma unroll
  for (int q = 0; q < MP_I; q++) {
    loc[q] = run;
    if (i0 + q < nb) {
      const ulonglong2 a = magg[i0 + q];
      run += make_i128(a.x, a.y);
    }
  }
  const i128 inc = warp_inclusive_scan_i128(run, lane);
  if (lane == 31) {
    s_lo[wid] = (unsigned long long)inc;
    s_hi[wid] = (unsigned long long)((u128)inc >> 64);
  }
  __syncthreads();
  if (wid == 0) {
    i128 w = lane < MP_T / 32 ? make_i128(s_lo[lane], s_hi[lane]) : (i128)0;
    const i128 wi = warp_inclusive_scan_i128(w, lane);
    const i128 total = shfl_i128(wi, 31);
    const i128 wex = wi - w;
    if (lane < MP_T / 32) {
      s_lo[lane] = (unsigned long long)wex;
      s_hi[lane] = (unsigned long long)((u128)wex >> 64);
    }
    i128 P;
    long long pc;
    lookback(desc, status, epoch, t, 0, total, 0ll, lane, P, pc);
    if (lane == 0) {
      s_plo = (unsigned long long)P;
      s_phi = (unsigned long long)((u128)P >> 64);
    }
  }
  __syncthreads();
  const i128 base = make_i128(s_plo, s_phi) + make_i128(s_lo[wid], s_hi[wid]) + (inc - run);
#pragma unroll
  for (int q = 0; q < MP_I; q++) {
    if (i0 + q < nb) {
      const i128 e = base + loc[q];
      mpre[i0 + q] = make_ulonglong2((unsigned long long)e, (unsigned long long)((u128)e >> 64));
    }
  }
}

int mass_prefix_tiles(int nb) { return (nb + MP_TILE - 1) / MP_TILE; }

void launch_mass_prefix(cudaStream_t st, const double *m, const unsigned *cnt, int cap, int nb, int fxE,
                        ulonglong2 *magg, ulonglong2 *mpre, Desc *desc, unsigned *status, unsigned *ticket,
                        unsigned epoch) {
  if (nb <= 0) return;
  bucket_mass_kernel<<<(nb + 7) / 8, 256, 0, st>>>(m, cnt, cap, nb, fxE, magg);
  mass_prefix_kernel<<<mass_prefix_tiles(nb), MP_T, 0, st>>>(magg, nb, mpre, desc, status, ticket, epoch);
}

void launch_count_prefix(cudaStream_t st, const unsigned *cnt, int nb, unsigned *cpre,
                         unsigned long long *tile_desc, unsigned *ticket, unsigned epoch) {
  if (nb > 0) count_prefix_kernel<<<count_prefix_tiles(nb), CP_T, 0, st>>>(cnt, nb, cpre, tile_desc, ticket, epoch);
}

template <int WCAP, int WARPS, int EQM, int SHARD>
static void launch_wstep_t(cudaStream_t st, const TileParams &p) {
  static OnceFlags attr_set;
  const size_t sm = sizeof(WarpSlab<WCAP, EQM>) * WARPS;
  WENDY_ONCE_PER_DEVICE(attr_set) {
    cudaFuncSetAttribute(wstep_kernel<WCAP, WARPS, EQM, SHARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  }
  const int grid = (p.nb + WARPS - 1) / WARPS;
  wstep_kernel<WCAP, WARPS, EQM, SHARD><<<grid, WARPS * 32, sm, st>>>(p);
}

bool wstep_cap_supported(int cap) { return cap == 256; }

void launch_wstep(cudaStream_t st, int cap, const TileParams &p) {
  if (p.nb <= 0) return;
  if (p.bounds) launch_wstep_t<256, WS_WARPS, 1, 1>(st, p);  // sharded mode: equal masses only
  else if (p.eqm) launch_wstep_t<256, WS_WARPS, 1, 0>(st, p);
  else launch_wstep_t<256, WS_WARPS, 0, 0>(st, p);
}

}  // namespace wendy
