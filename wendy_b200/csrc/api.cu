// api.cu -- handle, step orchestration and the C ABI of libwendy_b200.so
// (include/wendy_b200.h).  Host logic only; the kernels are in tile.cu and radix.cu.
//
// Step structure follows the reference driver wendy/wendy.c:385-418 exactly:
//   drift dt/2 ; (nleap-1) x [force, kick dt, drift dt] ; force, kick dt, drift dt/2 ; de-sort
// with the leading half drift folded into the first sub-step's key computation (h_pre) and
// the de-sort deferred to wendy_cuda_read().
#include <math.h>
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

#include "../../include/wendy_b200.h"
#include "common.cuh"
#include "internal.h"

using namespace wendy;

static thread_local std::string g_err;
static int set_err(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return set_err(WENDY_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));           \
  } while (0)

namespace wendy {
void launch_iota(cudaStream_t st, int *id, long long n);
void launch_gather_by_id(cudaStream_t st, const double *a_by_id, const int *id, const unsigned *cnt,
                         int cap, int nb, double *a_slots);
void launch_validate(cudaStream_t st, const double *x, const double *v, const double *m, long long n, double *out3);
void launch_make_keys_packed(cudaStream_t st, const double *packed, int prec, double h, long long n, uint64_t *keys,
                             uint32_t *vals);
void launch_fix_ties(cudaStream_t st, const uint64_t *keys, uint32_t *vals, const int *id_by_slot, long long n,
                     unsigned seg_div);
void launch_make_keys_by_id(cudaStream_t st, const double *x, const double *v, const int *id, const unsigned *cnt,
                            int cap, int nb, long long n, double h, uint32_t *inv_scratch, uint64_t *keys,
                            uint32_t *vals);
}

struct BounceRing;
struct wendy_cuda_handle {
  long long N = 0, seg_len = 0;
  int nseg = 1, mode = 0, fxE = 0;
  int nb_alloc = 0;
  bool eqm = false;        // all masses identical: no mass arrays
  double m0 = 0.;
  SerialTab *stab = nullptr;  // equal masses: closed form of the reference's serial running sum (serialsum.cuh)
  double omega2 = -1.;
  int cap = 0, fill = 0, nb = 0, nbps = 0;
  size_t slots = 0;
  cudaStream_t st = nullptr;
  int sm_count = 148;
  // state
  double *x[2] = {nullptr, nullptr}, *v[2] = {nullptr, nullptr}, *m[2] = {nullptr, nullptr};
  int *id[2] = {nullptr, nullptr};
  int cur = 0;
  unsigned *cnt[3] = {nullptr, nullptr, nullptr};
  int ccur = 0;
  bool small_ok = false;   // systems of <= 1024 particles: resident kernel while the state is dense
  bool adaptive = false;   // cap chosen by the library: 256 (warp kernel) <-> 2048 (CTA kernel)
  int dw = 256;            // destination window of the persistent CTA kernel, in buckets: doubled (up to tile_window_max())
  bool dw_fixed = false;   // while more than 0.2 % of the particles leave it per sub-step (WENDY_B200_DW=<n> pins it)
  long long dw_out_mark = 0, dw_sub_mark = 0;
  int want_cap = 0;        // geometry to switch to at the next layout rebuild (0: keep)
  bool coarse_default = false;  // large equal-mass systems start (and stay) on 2048-slot buckets
  bool fill_backoff = false;    // a library-chosen coarse layout overflowed at the optimistic fill: use 3/4 from now on
  bool rebuild_pending = false;  // shard: rebuild at the start of the next sub-step (state complete)
  int orec = 3;                          // doubles per migrant record: (x, v, id) or, general masses, (x, v, id, m)
  unsigned long long pm_lo = 0, pm_hi = 0;  // general masses, sharded: exact 128-bit mass owned by the lower ranks
  bool mpre_ready = false;               // magg / mpre already hold the bucket masses of the CURRENT state
  double fail_score = 0.;                // recent bucket overflows (decays with every clean call): sparser buckets at 3
  int shard_retry_k = -1, shard_retry_n = 0;  // shard: sub-step of the last rollback, consecutive rollbacks to it
  // WENDY_B200_SHARD_TRACE=1: CUDA events around the three launches of every sharded sub-step (peer exchange)
  std::vector<cudaEvent_t> tr_ev;
  double tr_ms[3] = {0., 0., 0.};
  unsigned tr_wait[2] = {0u, 0u};
  long long tr_n = 0;
  bool ext_async = false;        // ext-force stepping: between wendy_cuda_ext_begin and wendy_cuda_ext_end
  bool ext_half_done = false;    // ext-force stepping: the leading half drift of the call is already in x
  int ext_fail_streak = 0;       // ... consecutive overflows of the same sub-step (two: take it on the radix path)
  int nb_last = 0;               // last bucket of the layout that has a finite lower edge (the tail may be unused)
  int user_fill = 0, user_cap = 0;
  double last_dt = 0.;
  bool dense = true;       // state is the dense upload in buffer `cur` (no layout yet)
  bool has_split = false;  // splitters are valid for key = x + bucket_h * v
  double bucket_h = 0.;
  double *split = nullptr, *tot = nullptr;
  // Lagrangian splitters (advect_on): second edge buffer + per-cell flow statistics
  double *split_alt = nullptr, *knot_sum = nullptr, *knot_x = nullptr, *knot_y = nullptr;
  unsigned *knot_n = nullptr;
  bool advect_on = false;
  // sticky radix fallback: sub-steps still to run on the radix path, and the current back-off length
  int radix_left = 0, radix_streak = 0;
  long long fail_mark = 0;
  // cross-CTA machinery
  unsigned *ticket = nullptr;  // [3]
  int tcur = 0;
  unsigned *status = nullptr;
  Desc *desc = nullptr;
  unsigned *cpre = nullptr;            // exclusive prefix of the current bucket counts
  unsigned long long *cp_desc = nullptr;  // look-back words of the count_prefix kernel
  unsigned *cp_ticket = nullptr;
  ulonglong2 *magg = nullptr, *mpre = nullptr;  // general masses: bucket masses and their prefix
  Desc *mp_desc = nullptr;
  unsigned *mp_status = nullptr, *mp_ticket = nullptr;
  unsigned *flags = nullptr;    // [0] fail_seq, [1] max count; [8..135] = 64 x u64 outside-window counters
  unsigned *h_flags = nullptr;  // pinned mirror
  unsigned seq = 1;
  // scratch
  unsigned long long *offs = nullptr;
  RadixScratch rs;
  double *xo = nullptr, *vo = nullptr;
  double *xo2 = nullptr, *vo2 = nullptr;  // second de-sort staging set (wendy_cuda_stage_ahead)
  int stage_sel = 0;                      // staging set the last / current read-out copies from
  bool stage_ok = false;                  // the OTHER set holds the de-sorted result of the call in flight
  double *epart = nullptr, *eout = nullptr, *h_eout = nullptr;
  int *rank = nullptr;
  // sharded single system (one key range per GPU)
  int nranks = 1, my_rank = 0;
  long long n_cap = 0;          // particle capacity of this shard
  double *bounds = nullptr;     // device, nranks+1
  double *out_rec = nullptr;
  int *cid = nullptr;
  unsigned *out_cnt = nullptr, *h_out_cnt = nullptr;
  long long ocap = 0, pc_offset = 0;
  // ... device-driven exchange over peer memory (peer.cuh): comm buffer (peers write into it), mapped peers
  void *comm = nullptr;
  size_t comm_bytes = 0;
  void *peer_map[PEER_MAX] = {nullptr};   // cudaIpcOpenMemHandle results (closed on destroy)
  PeerComm *peer_dev = nullptr;           // device copy of the pointer table
  PeerComm peer_host;
  unsigned *peer_scratch = nullptr;       // out_cnt[PEER_MAX] | cta_done[2] | peer_stat[1]
  long long *peer_n = nullptr;            // n_local | n_hist[PEER_NHIST]
  long long *h_peer_n = nullptr;          // pinned mirror
  unsigned pepoch = 1, peer_mig_seen = 0;
  bool peer_on = false;
  std::vector<unsigned> p_seq_inj;
  std::vector<long long> p_nstart;
  // asynchronous call in flight (wendy_cuda_step_begin / _end) and overlapped read-out
  bool pending = false;
  double p_dt = 0.;
  int p_nleap = 0, p_k0 = 0;
  std::vector<unsigned> p_seq;
  std::vector<int> p_cur, p_ccur;
  cudaStream_t st_copy = nullptr;
  cudaEvent_t ev_unsort = nullptr;
  cudaEvent_t ev_call[2] = {nullptr, nullptr};  // asynchronous call: device time from the first to the last launch
  double call_host_s = 0.;                      // ... plus the host time of layout rebuilds inside begin / end
  int device = 0;               // CUDA device of the handle (worker threads select it)
  std::thread reader;           // bounce-buffered read-out in flight (wendy_cuda_read_begin / _end)
  int reader_rc = 0;
  struct BounceRing *ring = nullptr;
  // counters
  long long n_sub = 0, n_rebuild = 0, n_fail = 0, max_cnt = 0, n_outside = 0, n_launch = 0;
  long long n_radix_fallback = 0;
};
typedef wendy_cuda_handle H;

#include "hostmem.cuh"

static int choose_fx_exponent(double sum_abs) {
  if (!(sum_abs > 0.) || !std::isfinite(sum_abs)) return 0;
  int e;
  frexp(sum_abs, &e);  // sum_abs < 2^e
  return 124 - e;
}

static int alloc_radix(H *h, size_t n) {
  if (h->rs.n_alloc >= n) return 0;
  for (int i = 0; i < 2; i++) {
    if (h->rs.key[i]) dev_free(h->rs.key[i]);
    if (h->rs.val[i]) dev_free(h->rs.val[i]);
  }
  if (h->rs.table) dev_free(h->rs.table);
  if (h->rs.sums) dev_free(h->rs.sums);
  for (int i = 0; i < 2; i++) {
    CK(dev_alloc(&h->rs.key[i], n * sizeof(uint64_t)));
    CK(dev_alloc(&h->rs.val[i], n * sizeof(uint32_t)));
  }
  CK(dev_alloc(&h->rs.table, radix_table_entries(n) * sizeof(uint32_t)));
  CK(dev_alloc(&h->rs.sums, (radix_sums_entries(n) + 1) * sizeof(uint32_t)));
  h->rs.n_alloc = n;
  return 0;
}

static int seg_bits(const H *h) {
  int bits = 0;
  while ((1ll << bits) < h->nseg) bits++;
  return h->nseg > 1 ? bits : 0;
}

static void adapt_window(H *h);
// Sync and fetch {fail_seq, max count, outside count}.
static int fetch_flags(H *h) {
  CK(cudaMemcpyAsync(h->h_flags, h->flags, 136 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  CK(cudaGetLastError());
  if (h->h_flags[1] > h->max_cnt) h->max_cnt = h->h_flags[1];
  adapt_window(h);
  return 0;
}

static long long outside_total(const H *h) {
  long long t = 0;
  const unsigned long long *c = (const unsigned long long *)(h->h_flags + 8);
  for (int i = 0; i < 64; i++) t += (long long)c[i];
  return t;
}

// Destination window of the persistent CTA kernel (tile.cu): particles whose new key falls outside the window of
// splitters held in shared memory take a galloping search in the global table, which a whole warp executes for a few
// lanes.  Large N dt (long sub-steps, or one system over many GPUs) sends a noticeable share there; a wider window
// costs a little at small displacements (more splitters staged per bucket), so it is only widened on evidence.
// Called wherever the flags have just been fetched; results never depend on the window.
static void adapt_window(H *h) {
  if (h->dw_fixed || h->cap == 256) return;
  const long long out = h->n_outside + outside_total(h), sub = h->n_sub;
  const long long dsub = sub - h->dw_sub_mark, dout = out - h->dw_out_mark;
  if (dsub <= 0) {
    if (dsub < 0) { h->dw_sub_mark = sub; h->dw_out_mark = out; }
    return;
  }
  if (h->dw < tile_window_max() && (double)dout > 0.002 * (double)h->N * (double)dsub) h->dw = std::min(2 * h->dw, tile_window_max());
  h->dw_sub_mark = sub;
  h->dw_out_mark = out;
}

static int reset_flags(H *h) {
  static const unsigned init[2] = {0xffffffffu, 0u};
  h->n_outside += outside_total(h);
  memset(h->h_flags + 8, 0, 128 * sizeof(unsigned));
  CK(cudaMemcpyAsync(h->flags, init, sizeof(init), cudaMemcpyHostToDevice, h->st));
  CK(cudaMemsetAsync(h->flags + 8, 0, 128 * sizeof(unsigned), h->st));
  CK(cudaMemsetAsync(h->ticket, 0, 3 * sizeof(unsigned), h->st));
  CK(cudaStreamSynchronize(h->st));  // `init` is a stack buffer
  h->tcur = 0;
  return 0;
}

// The launch sequence number doubles as the epoch of the look-back protocols, which keep 30 bits of it
// (lookback.cuh: status = epoch << 2 | state; count prefix: epoch << 34).  Long before it wraps -- after about 2^30
// launches, hours of stepping a small system -- start again from 1 with clean look-back words.  Called at the
// start of a stepping call, when nothing of this handle is in flight.
static int renew_epochs(H *h) {
  static const unsigned limit = [] {  // (WENDY_B200_EPOCH_RENEW_AT: tests exercise the renewal after a few launches)
    const char *e = getenv("WENDY_B200_EPOCH_RENEW_AT");
    return e ? (unsigned)strtoul(e, nullptr, 10) : 0x3f000000u;
  }();
  if (h->seq < limit) return 0;
  CK(cudaStreamSynchronize(h->st));
  CK(cudaMemsetAsync(h->status, 0, (size_t)h->nb_alloc * sizeof(unsigned), h->st));
  CK(cudaMemsetAsync(h->cp_desc, 0, (size_t)(count_prefix_tiles(h->nb_alloc) + 1) * sizeof(unsigned long long), h->st));
  if (h->mp_status)
    CK(cudaMemsetAsync(h->mp_status, 0, ((size_t)mass_prefix_tiles(h->nb_alloc) + 1) * sizeof(unsigned), h->st));
  h->seq = 1;
  return reset_flags(h);
}

// keys (x + hkey*v) of the current state -> radix scratch buffer 0, compact segment-major order
static int make_keys(H *h, double hkey, int val_mode) {
  if (alloc_radix(h, (size_t)h->N)) return WENDY_E_CUDA;
  if (h->dense) {
    launch_make_keys(h->st, h->x[h->cur], h->v[h->cur], hkey, nullptr, nullptr, 0, 0, h->N,
                     h->rs.key[0], h->rs.val[0], val_mode, h->seg_len, 0);
  } else {
    launch_scan_counts(h->st, h->cnt[h->ccur], h->nb, h->offs);
    launch_make_keys(h->st, h->x[h->cur], h->v[h->cur], hkey, h->cnt[h->ccur], h->offs, h->cap,
                     h->nb, 0, h->rs.key[0], h->rs.val[0], val_mode, h->seg_len, h->nbps);
    h->n_launch += 1;
  }
  h->n_launch += 1;
  return 0;
}

// WENDY_B200_ADVECT=0 disables the Lagrangian splitters (A/B experiments)
static bool advect_allowed() {
  const char *e = getenv("WENDY_B200_ADVECT");
  return !(e && e[0] == '0');
}

// Target particles per bucket.  Coarse buckets chosen by the library are filled to 13/16 (1664 of 2048 slots:
// per-bucket overheads amortise over more particles, +2 % at N=1e8); the head-room is then 6.7 sigma of the
// steady-state count noise (sqrt(2*fill)) instead of 9, so the first overflow or
// nearly-full bucket of a handle -- a system whose density evolves -- moves it back to 3/4 for good.
#ifndef WENDY_COARSE_FILL_16THS
#define WENDY_COARSE_FILL_16THS 13
#endif
static int default_fill(const H *h, int cap) {
  if (h->user_fill > 0 && cap == h->user_cap) return h->user_fill;
  if (cap == 256) return 128;
  // (shards too: migrants land in the buckets their keys fall in, spread over the same +-15 buckets as everybody
  // else's movers, and an overflow is rolled back like any other -- +5 % per sub-step at 1e8 particles per rank)
  return (h->adaptive && !h->fill_backoff) ? cap * WENDY_COARSE_FILL_16THS / 16 : cap * 3 / 4;
}

// called when a bucket of the current layout overflowed or came close: give up the optimistic fill
static void fill_back_off(H *h) {
  if (h->adaptive && !h->fill_backoff && h->cap != 256 && h->fill > h->cap * 3 / 4 &&
      !(h->user_fill > 0 && h->cap == h->user_cap)) {
    h->fill_backoff = true;
    h->fill = h->cap * 3 / 4;
  }
}

// A sharded sub-step that overflows again on a FRESH layout (the density changes by more than the head-room within one
// sub-step: violent relaxation at a coarse dt) has no radix path to fall back on: buy head-room instead.  Every
// further retry halves the fill, down to what the shard's storage can hold (2 N slots for small shards = 2.5 x
// head-room; 1.7 x for shards created coarse).
static void fill_escalate(H *h) {
  const int cap = h->want_cap ? h->want_cap : h->cap;
  if (cap == 256) return;
  const long long nb_max = (long long)(h->slots / (size_t)cap) / h->nseg;  // buckets one segment can have
  const long long need = h->bounds ? (long long)((double)h->N * 1.02) + 1 : h->n_cap / h->nseg;  // (as rebucket sizes it)
  int f_min = nb_max > 0 ? (int)((need + nb_max - 1) / nb_max) + 1 : cap;
  f_min = std::max(f_min, cap / 16);
  const int cur = (cap == h->cap) ? h->fill : cap * 3 / 4;
  const int nf = std::max(f_min, cur / 2);
  if (nf < cur) {
    h->fill_backoff = true;
    h->user_fill = nf; h->user_cap = cap;  // (default_fill honours it when the geometry is switched as well)
    if (cap == h->cap) h->fill = nf;
  }
}

// (Re)build the bucket layout for key = x + hkey*v: exact quantile splitters from a radix
// sort of the keys, then one streaming scatter of the state into the other buffer (optionally
// together with `n_extra` packed migrant records: shard inject).
static int rebucket(H *h, double hkey, const double *extra = nullptr, long long n_extra = 0) {
  // target geometry (may differ from the geometry the state is currently stored in)
  const int ncap = h->want_cap ? h->want_cap : h->cap;
  const int nfill = (ncap == h->cap) ? h->fill : default_fill(h, ncap);
  // buckets for the particles actually present: a shard is allocated for n_cap > N particles (migrants may
  // accumulate), but a layout sized for n_cap would run every bucket under-filled (more buckets, same
  // per-bucket overheads).  2 % head-room; if the shard outgrows it the overflow protocol rebuilds.
  long long n_target = h->n_cap / h->nseg;
  if (h->bounds) n_target = std::min(n_target, (long long)((double)(h->N + n_extra) * 1.02) + 1);
  const int nnbps = (int)((n_target + nfill - 1) / nfill);
  const int nnb = nnbps * h->nseg;
  trace_mark(h->st, "(rebucket: start)");
  if (alloc_radix(h, (size_t)(h->N + n_extra))) return WENDY_E_CUDA;
  trace_mark(h->st, "rebucket: radix scratch");
  if (make_keys(h, hkey, VAL_SEGMENT)) return WENDY_E_CUDA;
  if (n_extra > 0)  // shard inject: the layout is built from the union of local state and inbox
    launch_make_keys_packed(h->st, extra, h->orec, hkey, n_extra, h->rs.key[0] + h->N, h->rs.val[0] + h->N);
  const long long n_all = h->N + n_extra;
  int res = radix_sort_pairs(h->st, h->rs, (size_t)n_all, seg_bits(h), 1u);
  h->n_launch += 5 * (8 + (seg_bits(h) + 7) / 8);
  launch_pick_splitters(h->st, h->rs.key[res], n_extra > 0 ? n_all : h->seg_len, nfill, nnbps, nnb, h->split);
  trace_mark(h->st, "rebucket: keys + radix sort + splitters");
  int c1 = (h->ccur + 1) % 3, c2 = (h->ccur + 2) % 3;
  CK(cudaMemsetAsync(h->cnt[c1], 0, (size_t)h->nb_alloc * sizeof(unsigned), h->st));
  CK(cudaMemsetAsync(h->cnt[c2], 0, (size_t)h->nb_alloc * sizeof(unsigned), h->st));
  ScatterParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.xin = h->x[h->cur]; sp.vin = h->v[h->cur]; sp.min = h->m[h->cur]; sp.idin = h->id[h->cur];
  sp.cnt_in = h->dense ? nullptr : h->cnt[h->ccur];
  sp.n_dense = h->N; sp.cap_in = h->cap; sp.nb_in = h->nb; sp.nbps_in = h->nbps;
  sp.h = hkey;
  int o = h->cur ^ 1;
  sp.xout = h->x[o]; sp.vout = h->v[o]; sp.mout = h->m[o]; sp.idout = h->id[o];
  sp.cnt_out = h->cnt[c1]; sp.split = h->split; sp.cap_out = ncap; sp.nbps_out = nnbps;
  sp.seg_len = h->seg_len; sp.fail_seq = h->flags; sp.seq = h->seq++;
  launch_scatter(h->st, sp, h->sm_count);
  if (n_extra > 0) {
    sp.packed_in = extra; sp.prec = h->orec; sp.cnt_in = nullptr; sp.n_dense = n_extra; sp.min = nullptr;
    sp.seg_len = h->n_cap + n_extra + 1; sp.seq = h->seq++;
    launch_scatter(h->st, sp, h->sm_count);
  }
  h->n_launch += 2;
  if (fetch_flags(h)) return WENDY_E_CUDA;
  if (h->h_flags[0] != 0xffffffffu) {
    reset_flags(h);
    return set_err(WENDY_E_OVERFLOW, "bucket overflow while building the layout: too many exactly "
                                     "coincident particles for one bucket");
  }
  h->cur = o; h->ccur = c1; h->dense = false; h->has_split = true; h->bucket_h = hkey; h->mpre_ready = false;
  trace_mark(h->st, "rebucket: scatter");
  h->cap = ncap; h->fill = nfill; h->nbps = nnbps; h->nb = nnb; h->want_cap = 0;
  h->nb_last = (int)(((n_extra > 0 ? n_all : h->seg_len) - 1) / nfill);
  {
    const size_t nc = (size_t)h->nb_alloc / 8 + 8;  // flow statistics belong to the old layout
    CK(cudaMemsetAsync(h->knot_sum, 0, nc * sizeof(double), h->st));
    CK(cudaMemsetAsync(h->knot_n, 0, nc * sizeof(unsigned), h->st));
  }
  h->n_rebuild++;
  CK(cudaMemsetAsync(h->flags + 1, 0, sizeof(unsigned), h->st));  // max-count is per layout
  h->h_flags[1] = 0;
  return 0;
}

static void fill_tile_params(H *h, TileParams &p) {
  memset(&p, 0, sizeof(p));
  int c = h->cur, o = c ^ 1;
  p.xin = h->x[c]; p.vin = h->v[c]; p.min = h->m[c]; p.idin = h->id[c];
  p.xout = h->x[o]; p.vout = h->v[o]; p.mout = h->m[o]; p.idout = h->id[o];
  p.cnt_in = h->cnt[h->ccur];
  p.cnt_out = h->cnt[(h->ccur + 1) % 3];
  p.cnt_zero = h->cnt[(h->ccur + 2) % 3];
  p.split = h->split; p.split_in = h->split;
  p.nb = h->nb; p.nbps = h->nbps; p.seg_len = h->seg_len; p.dw = h->dw;
  p.omega2 = h->omega2; p.tot = h->tot; p.fxE = h->fxE;
  p.status = h->status; p.desc = h->desc;
  p.eqm = h->eqm ? 1 : 0; p.m0 = h->m0; p.stab = h->stab;
  p.nranks = h->nranks; p.my_rank = h->my_rank; p.bounds = h->bounds;
  p.out_rec = h->out_rec; p.out_cnt = h->out_cnt; p.orec = h->orec; p.pm_lo = h->pm_lo; p.pm_hi = h->pm_hi;
  p.ocap = (unsigned)h->ocap; p.pc_offset = h->pc_offset;
  p.peer = h->peer_on ? h->peer_dev : nullptr; p.pepoch = h->pepoch; p.nb_last = h->nb_last;
  p.ticket = h->ticket + h->tcur; p.ticket_zero = h->ticket + (h->tcur + 2) % 3;
  p.fail_seq = h->flags; p.stats = h->flags + 1; p.outside = (unsigned long long *)(h->flags + 8);
  p.seq = h->seq; p.epoch = h->seq;
}

static void advance_after_tile(H *h) {
  h->seq++;
  h->tcur = (h->tcur + 1) % 3;
  h->n_launch++;
}

// One sub-step on the bucket fast path (asynchronous).
static void launch_bucket_substep(H *h, double h_pre, double dt_kick, double dt_drift, double h_next,
                                  const double *aext, int *rank_out) {
  TileParams p;
  fill_tile_params(h, p);
  p.h_pre = h_pre; p.dt_kick = dt_kick; p.dt_drift = dt_drift; p.h_next = h_next;
  p.aext = aext; p.rank_out = rank_out;
  if (h->advect_on && h->nseg == 1) {
    // move the bucket edges with the flow measured by the previous sub-step: input layout = h->split,
    // output layout = the advected copy
    const int G = (h->cap == 256) ? 64 : 8;
    launch_advect_splitters(h->st, h->split, h->split_alt, h->nb, G, h->knot_sum, h->knot_n, h->knot_x, h->knot_y);
    h->n_launch += 2;
    p.split_in = h->split; p.split = h->split_alt;
    p.knot_sum = h->knot_sum; p.knot_n = h->knot_n; p.knot_g = G;
    std::swap(h->split, h->split_alt);
  }
  launch_count_prefix(h->st, p.cnt_in, h->nb, h->cpre, h->cp_desc, h->cp_ticket, p.epoch);
  h->n_launch++;
  p.cpre = h->cpre;
  if (!h->eqm) {
    if (!h->mpre_ready) {  // (a shard has just computed them for wendy_cuda_shard_mass_total)
      launch_mass_prefix(h->st, p.min, p.cnt_in, h->cap, h->nb, h->fxE, h->magg, h->mpre, h->mp_desc, h->mp_status,
                         h->mp_ticket, p.epoch);
      h->n_launch += 2;
    }
    h->mpre_ready = false;
    p.mpre = h->mpre;
  }
  if (wstep_cap_supported(h->cap)) launch_wstep(h->st, h->cap, p);
  else launch_tile(h->st, h->cap, LOAD_BUCKET, EMIT_SPLITTER, 1, p);
  advance_after_tile(h);
  h->cur ^= 1; h->ccur = (h->ccur + 1) % 3; h->bucket_h = h_next;
  h->n_sub++;
}

// One sub-step on the radix path: full sort of (key, slot), then the same tile kernel fed
// through the sorted permutation; the output is a compact sorted layout.
static int launch_radix_substep(H *h, double h_pre, double dt_kick, double dt_drift, const double *aext,
                                int *rank_out) {
  // keys in storage order, value = storage slot.  A dense upload stores particle i in slot i, so the stable sort
  // already breaks ties by index; a bucket layout does not -- its (rare) ties are put right after the sort.
  // WENDY_B200_RADIX_BYID=1 restores the earlier scheme (keys generated in particle-id order) for A/B runs.
  static const bool by_id = getenv("WENDY_B200_RADIX_BYID") != nullptr;
  const bool fix = !(h->dense || h->bounds) && !by_id;
  if (h->dense || h->bounds || fix) {
    if (make_keys(h, h_pre, VAL_SLOT)) return WENDY_E_CUDA;
  } else {
    // keys generated in particle-id order (segment-major, ids are contiguous per segment)
    if (alloc_radix(h, (size_t)h->N)) return WENDY_E_CUDA;
    launch_make_keys_by_id(h->st, h->x[h->cur], h->v[h->cur], h->id[h->cur], h->cnt[h->ccur], h->cap, h->nb,
                           h->N, h_pre, h->rs.val[1], h->rs.key[0], h->rs.val[0]);
    h->n_launch += 2;
  }
  unsigned seg_div = h->dense ? (unsigned)h->seg_len : (unsigned)((long long)h->nbps * h->cap);
  int res = radix_sort_pairs(h->st, h->rs, (size_t)h->N, seg_bits(h), seg_div);
  h->n_launch += 5 * (8 + (seg_bits(h) + 7) / 8);
  if (fix) {
    launch_fix_ties(h->st, h->rs.key[res], h->rs.val[res], h->id[h->cur], h->N, seg_bits(h) ? seg_div : 0u);
    h->n_launch++;
  }
  TileParams p;
  fill_tile_params(h, p);
  p.perm = h->rs.val[res];
  p.h_pre = h_pre; p.dt_kick = dt_kick; p.dt_drift = dt_drift; p.h_next = 0.;
  p.aext = aext; p.rank_out = rank_out;
  launch_tile(h->st, h->cap, LOAD_GATHER, EMIT_RANK, 1, p);
  advance_after_tile(h);
  h->cur ^= 1; h->ccur = (h->ccur + 1) % 3; h->dense = false; h->has_split = false;
  h->n_sub++;
  return 0;
}

extern "C" {

const char *wendy_cuda_last_error(void) { return g_err.c_str(); }

void wendy_cuda_destroy(wendy_cuda_handle *h) {
  if (!h) return;
  // released blocks may be handed to another handle at once: nothing of this one may still be running
  if (h->reader.joinable()) h->reader.join();
  if (h->st_copy) cudaStreamSynchronize(h->st_copy);
  cudaStreamSynchronize(h->st);
  if (h->tr_n > 0) {
    char line[512];  // one write(2): the ranks' lines must not interleave
    int len = snprintf(line, sizeof(line), "wendy_b200 shard trace rank %d: %lld sub-steps, ms each: prefix %.4f step %.4f "
                       "inject %.4f; waiting (CTA 0): step %.4f inject %.4f\n", h->my_rank, h->tr_n, h->tr_ms[0] / h->tr_n,
                       h->tr_ms[1] / h->tr_n, h->tr_ms[2] / h->tr_n,
                       h->tr_wait[0] * 1.024e-3 / std::max(1ll, h->n_sub), h->tr_wait[1] * 1.024e-3 / std::max(1ll, h->n_sub));
    if (len > 0 && fwrite(line, 1, (size_t)std::min(len, (int)sizeof(line) - 1), stderr) > 0) fflush(stderr);
  }
  for (cudaEvent_t e : h->tr_ev) cudaEventDestroy(e);
  h->tr_ev.clear();
  ring_release(h->ring);
  h->ring = nullptr;
  for (int i = 0; i < 2; i++) {
    dev_free(h->x[i]); dev_free(h->v[i]); dev_free(h->m[i]); dev_free(h->id[i]);
    dev_free(h->rs.key[i]); dev_free(h->rs.val[i]);
  }
  for (int i = 0; i < 3; i++) dev_free(h->cnt[i]);
  dev_free(h->rs.table); dev_free(h->rs.sums);
  dev_free(h->split_alt); dev_free(h->knot_sum); dev_free(h->knot_x); dev_free(h->knot_y); dev_free(h->knot_n);
  dev_free(h->split); dev_free(h->tot); dev_free(h->ticket); dev_free(h->status); dev_free(h->desc);
  dev_free(h->cpre); dev_free(h->cp_desc); dev_free(h->cp_ticket);
  dev_free(h->magg); dev_free(h->mpre); dev_free(h->mp_desc); dev_free(h->mp_status); dev_free(h->mp_ticket);
  dev_free(h->flags); dev_free(h->offs); dev_free(h->xo); dev_free(h->vo); dev_free(h->xo2); dev_free(h->vo2); dev_free(h->epart);
  dev_free(h->eout); dev_free(h->rank);
  dev_free(h->bounds); dev_free(h->out_rec); dev_free(h->out_cnt);
  dev_free(h->cid); dev_free(h->stab);
  for (int r = 0; r < PEER_MAX; r++)
    if (h->peer_map[r]) cudaIpcCloseMemHandle(h->peer_map[r]);
  if (h->comm) cudaFree(h->comm);
  dev_free(h->peer_dev); dev_free(h->peer_scratch); dev_free(h->peer_n);
  if (h->h_peer_n) cudaFreeHost(h->h_peer_n);
  if (h->h_out_cnt) cudaFreeHost(h->h_out_cnt);
  if (h->st_copy) cudaStreamDestroy(h->st_copy);
  if (h->ev_unsort) cudaEventDestroy(h->ev_unsort);
  for (int i = 0; i < 2; i++) if (h->ev_call[i]) cudaEventDestroy(h->ev_call[i]);
  if (h->h_flags) cudaFreeHost(h->h_flags);
  if (h->h_eout) cudaFreeHost(h->h_eout);
  delete h;
}

static inline bool dev_inputs_shard(const int *ids) { return ids != nullptr; }  // shard handles carry global ids

// Host -> device upload of large arrays.  Page-locked sources go straight to cudaMemcpyAsync; pageable ones
// (what a numpy caller normally has) are staged: all host threads copy 32 MB chunks into two page-locked
// bounce buffers while the previous chunk is in flight, which beats the driver's single-threaded staging
// several times over on many-core hosts.
// Pageable host arrays -> device through page-locked staging: all host threads fill piece i+1 while the copy engine
// moves piece i.  `ring`: the pooled bounce ring of the read-out (four 16 MB buffers: nothing to allocate -- obtaining
// and releasing 64 MB of page-locked memory per upload cost as much as the copies); without one, two private buffers.
static int upload_host_arrays(cudaStream_t st, const double *const *src, double *const *dst, int narr, size_t n,
                              std::string &err, BounceRing *ring = nullptr) {
  constexpr int MAXB = BounceRing::NB > 2 ? BounceRing::NB : 2;
  const int nbuf = ring ? BounceRing::NB : 2;
  const size_t CH = ring ? BounceRing::BYTES / sizeof(double) : ((size_t)4 << 20);  // doubles per chunk
  double *stage[MAXB] = {};
  cudaEvent_t ev[MAXB] = {};
  bool used[MAXB] = {};
  if (ring)
    for (int i = 0; i < nbuf; i++) { stage[i] = (double *)ring->buf[i]; ev[i] = ring->ev[i]; }
  int rc = 0, turn = 0;
  auto fail = [&](cudaError_t e, const char *what) { err = std::string(what) + ": " + cudaGetErrorString(e); rc = -1; };
  for (int a = 0; a < narr && !rc; a++) {
    if (!src[a] || !dst[a]) continue;
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, src[a]) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned || n < CH) {
      cudaError_t e = copy_split(dst[a], src[a], n * sizeof(double), cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) fail(e, "cudaMemcpyAsync");
      continue;
    }
    for (size_t off = 0; off < n && !rc; off += CH) {
      const size_t len = std::min(CH, n - off);
      const int bsel = turn % nbuf;
      turn++;
      if (!stage[bsel]) {
        cudaError_t e = cudaMallocHost(&stage[bsel], CH * sizeof(double));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[bsel], cudaEventDisableTiming);
        if (e != cudaSuccess) { fail(e, "cudaMallocHost"); break; }
      }
      if (used[bsel]) cudaEventSynchronize(ev[bsel]);  // the copy that last read this bounce buffer
      const double *sp = src[a] + off;
      double *bp = stage[bsel];
#pragma omp parallel for schedule(static) num_threads(host_threads())
      for (long long blk = 0; blk < (long long)((len + 65535) / 65536); blk++) {
        const size_t b0 = (size_t)blk * 65536, bl = std::min((size_t)65536, len - b0);
        memcpy(bp + b0, sp + b0, bl * sizeof(double));
      }
      cudaError_t e = cudaMemcpyAsync(dst[a] + off, bp, len * sizeof(double), cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaEventRecord(ev[bsel], st);
      if (e != cudaSuccess) { fail(e, "cudaMemcpyAsync"); break; }
      used[bsel] = true;
    }
  }
  for (int i = 0; i < nbuf; i++) {
    if (used[i]) cudaEventSynchronize(ev[i]);
    if (!ring) {
      if (ev[i]) cudaEventDestroy(ev[i]);
      if (stage[i]) cudaFreeHost(stage[i]);
    }
  }
  return rc;
}

// dev_inputs: x, v (and m unless null: then all masses equal m0_dev) are DEVICE arrays
static int create_impl(wendy_cuda_handle **out, long long N, long long n_cap, const double *x, const double *v,
                       const double *m, const int *ids, const double *totmass, double omega2, int n_segments,
                       int flags, int cap, int fill, void *cuda_stream, bool dev_inputs = false,
                       double m0_dev = 0.) {
  if (!out || !x || !v || (!m && !dev_inputs) || !totmass) return set_err(WENDY_E_ARG, "null argument");
  if (N <= 0 || N >= (1ll << 31) || n_cap < N || n_cap >= (1ll << 31))
    return set_err(WENDY_E_ARG, "N must be in [1, 2^31) and not exceed the capacity");
  if (n_segments < 1 || N % n_segments) return set_err(WENDY_E_ARG, "N must be a multiple of n_segments");
  const bool adaptive = (cap == 0);
  if (cap == 0) cap = 256;
  trace_mark((cudaStream_t)cuda_stream, "(create: start)");
  if (!tile_cap_supported(cap)) return set_err(WENDY_E_ARG, "cap must be 2048 or 256");  // (2048: tile_coarse_cap())
  // Splitters are exact quantiles of ONE random sample, so bucket widths carry their own
  // 1/sqrt(fill) noise and the steady-state count variance is 2*fill (measured: DESIGN.md).
  // Defaults leave >= 8 sigma of head-room: 128 + 8*sqrt(256) = 256, 1536 + 9*sqrt(3072) < 2048.
  const int fill_arg = fill;
  if (fill == 0) fill = (cap == 256) ? 128 : cap * 3 / 4;
  if (fill < 1 || fill > cap) return set_err(WENDY_E_ARG, "fill must be in [1, cap]");
  H *h = new H;
  *out = nullptr;
  h->user_fill = fill_arg; h->user_cap = cap;
  h->N = N; h->n_cap = n_cap; h->nseg = n_segments; h->seg_len = N / n_segments; h->omega2 = omega2;
  h->mode = flags & 0xf; h->cap = cap; h->fill = fill;
  h->adaptive = adaptive && h->mode != WENDY_SORT_RADIX;
  if (const char *edw = getenv("WENDY_B200_DW")) {  // A/B runs: a fixed destination window
    h->dw = std::max(8, std::min(tile_window_max(), atoi(edw)));
    h->dw_fixed = true;
  }
  h->small_ok = h->adaptive && !dev_inputs_shard(ids) && (N / n_segments) <= small_max_particles();
  h->nbps = (int)(((n_cap / n_segments) + fill - 1) / fill);
  // mid-size systems (the whole state is a few tens of MB): half as many slots again, so that a system whose
  // density keeps changing faster than the layout's head-room (violent relaxation) can be given sparser buckets
  // (fill_escalate) instead of a rebuild every other sub-step
  if (adaptive && N >= (1ll << 16) && N < (1ll << 20)) h->nbps += h->nbps / 2;
  long long nb = (long long)h->nbps * n_segments;
  if (nb * cap >= (1ll << 32)) { delete h; return set_err(WENDY_E_ARG, "too many storage slots for u32 indices"); }
  h->nb = (int)nb; h->nb_alloc = (int)nb; h->slots = (size_t)nb * cap;
  h->st = (cudaStream_t)cuda_stream;
  int dev = 0;
#define CKD(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { std::string s__ = std::string(#call) + ": " + cudaGetErrorString(e__); wendy_cuda_destroy(h); return set_err(WENDY_E_CUDA, s__); } } while (0)
  CKD(cudaGetDevice(&dev));
  h->device = dev;
  CKD(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev));
  // Geometry chosen by the library: the persistent CTA kernel (2048-slot buckets, next bucket prefetched by
  // TMA) is the fastest step for large equal-mass systems at every dt measured; small
  // systems start on 256-slot buckets (warp kernel) and switch when the window statistic says so.  Large systems
  // of unequal masses start coarse as well: their CTA kernel beats their warp kernel (128-bit scan across one warp,
  // 11 KB slab) at every dt measured (N=2e7, dt_leap=1e-4: 0.68 against 1.08 ms, profiles/r02/kernels_tour_N1e8.md).
  if (h->adaptive && N >= (1ll << 20)) {
    // ... and stay there, so the storage is sized for that geometry: n/(3/4) slots per particle array instead
    // of the 2n of the fine layout (less to allocate -- cudaMalloc is a visible part of the set-up time --
    // and N=1e9 needs 75 GB instead of 106 GB).  Bucket arrays are sized for the conservative fill, which
    // the handle falls back to after an overflow (fill_back_off).
    h->coarse_default = true;
    h->cap = tile_coarse_cap();
    h->fill = default_fill(h, h->cap);
    h->nbps = (int)(((n_cap / n_segments) + (h->cap * 3 / 4) - 1) / (h->cap * 3 / 4));
    if (N < (1ll << 23)) h->nbps *= 2;  // (mid-size: room for sparser buckets, see above; 53 -> 107 B/particle)
    h->nb = h->nbps * n_segments;
    h->nb_alloc = h->nb;
    h->slots = (size_t)h->nb * h->cap;
  }
  // The bulk of the device memory (both state buffers) is allocated by a helper thread while this one validates
  // the host arrays: 40-55 ms of cudaMalloc at N=1e8 that used to sit between validation and upload.
  struct AllocJob {
    std::thread th;
    cudaError_t err = cudaSuccess;
    ~AllocJob() { if (th.joinable()) th.join(); }
  } alloc_job;
  if (!dev_inputs && (size_t)N >= ((size_t)1 << 22)) {
    H *hh = h;
    AllocJob *job = &alloc_job;
    alloc_job.th = std::thread([hh, job]() {
      cudaSetDevice(hh->device);
      for (int i = 0; i < 2 && job->err == cudaSuccess; i++) {
        job->err = dev_alloc(&hh->x[i], hh->slots * sizeof(double));
        if (job->err == cudaSuccess) job->err = dev_alloc(&hh->v[i], hh->slots * sizeof(double));
        if (job->err == cudaSuccess) job->err = dev_alloc(&hh->id[i], hh->slots * sizeof(int));
      }
    });
  }
  double sum_abs = 0.;
  if (dev_inputs) {
    // device inputs: one validation kernel (finiteness, sum |m|, equal-mass test)
    double *d_out = nullptr;
    double h_out[3] = {0., 0., 0.};
    CKD(dev_alloc(&d_out, 3 * sizeof(double)));
    CKD(cudaMemsetAsync(d_out, 0, 3 * sizeof(double), h->st));
    launch_validate(h->st, x, v, m, N, d_out);
    CKD(cudaMemcpyAsync(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost, h->st));
    CKD(cudaStreamSynchronize(h->st));
    dev_free(d_out);
    if (!(h_out[0] == 0.)) {
      wendy_cuda_destroy(h);
      return set_err(WENDY_E_ARG, "x, v, m must be finite (NaN keys are undefined in the reference sort too)");
    }
    sum_abs = m ? h_out[1] : fabs(m0_dev) * (double)h->seg_len;  // an upper bound per segment suffices
    h->eqm = !m || (!(flags & WENDY_FLAG_GENERAL_MASSES) && h_out[2] == 0.);
  } else {
    // one multi-threaded pass over the inputs: finiteness (x*0 is 0 for finite x, NaN otherwise),
    // per-segment sum of |m| for the fixed-point exponent, and the equal-mass test
    double probe = 0.;
    long long n_diff = 0;
    const double m_first = m[0];
    const long long L = h->seg_len;
#pragma omp parallel for reduction(+ : probe, n_diff) reduction(max : sum_abs) schedule(static) if (n_segments > 1) num_threads(host_threads())
    for (int s = 0; s < (n_segments > 1 ? n_segments : 0); s++) {
      double a = 0.;
      for (long long i = s * L; i < (s + 1) * L; i++) {
        probe += x[i] * 0. + v[i] * 0. + m[i] * 0.;
        a += fabs(m[i]);
        n_diff += (m[i] != m_first);
      }
      if (a > sum_abs) sum_abs = a;
    }
    if (n_segments == 1) {  // a single segment: parallelise over chunks instead
      probe = 0.; n_diff = 0; sum_abs = 0.;
#pragma omp parallel for reduction(+ : probe, n_diff, sum_abs) schedule(static) num_threads(host_threads())
      for (long long i = 0; i < N; i++) {
        probe += x[i] * 0. + v[i] * 0. + m[i] * 0.;
        sum_abs += fabs(m[i]);
        n_diff += (m[i] != m_first);
      }
    }
    if (!(probe == 0.)) {
      if (alloc_job.th.joinable()) alloc_job.th.join();
      wendy_cuda_destroy(h);
      return set_err(WENDY_E_ARG, "x, v, m must be finite (NaN keys are undefined in the reference sort too)");
    }
    h->eqm = !(flags & WENDY_FLAG_GENERAL_MASSES) && n_diff == 0;
  }
  h->fxE = choose_fx_exponent(sum_abs);
  trace_mark(h->st, "create: validation pass");
  // (joined here: every error path below releases the handle)
  if (alloc_job.th.joinable()) alloc_job.th.join();
  if (dev_inputs) {
    h->m0 = m0_dev;
    if (m) CKD(cudaMemcpy(&h->m0, m, sizeof(double), cudaMemcpyDeviceToHost));
  } else {
    h->m0 = m[0];
  }
#ifdef WENDY_FORCE_EXACT_SCAN  // A/B builds: what the serial table costs
  flags |= WENDY_FLAG_EXACT_SCAN;
#endif
  if (h->eqm && !(flags & WENDY_FLAG_EXACT_SCAN) && h->m0 != 0. && std::isfinite(h->m0)) {
    // Equal masses: reproduce the reference's serial fp64 running sum (wendy/wendy.c:359-360) bit for bit
    // through its closed form, for every sorted position a system (or a sharded system's global rank) can
    // have.  WENDY_FLAG_EXACT_SCAN keeps the correctly rounded exact sum RN(rank*m0) of the general path.
    SerialTab *T = new SerialTab;
    if (serial_tab_build(*T, h->m0, 1ll << 31) == 0) {
      cudaError_t e = dev_alloc(&h->stab, sizeof(SerialTab));
      if (e == cudaSuccess) e = cudaMemcpyAsync(h->stab, T, sizeof(SerialTab), cudaMemcpyHostToDevice, h->st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
      if (e != cudaSuccess) { delete T; std::string s__ = cudaGetErrorString(e); wendy_cuda_destroy(h); return set_err(WENDY_E_CUDA, s__); }
    }
    delete T;
  }
  if (alloc_job.th.joinable()) alloc_job.th.join();
  CKD(alloc_job.err);
  for (int i = 0; i < 2; i++) {
    if (!h->x[i]) CKD(dev_alloc(&h->x[i], h->slots * sizeof(double)));
    if (!h->v[i]) CKD(dev_alloc(&h->v[i], h->slots * sizeof(double)));
    if (!h->eqm) CKD(dev_alloc(&h->m[i], h->slots * sizeof(double)));
    if (!h->id[i]) CKD(dev_alloc(&h->id[i], h->slots * sizeof(int)));
    CKD(cudaMemsetAsync(h->x[i], 0, h->slots * sizeof(double), h->st));
    CKD(cudaMemsetAsync(h->v[i], 0, h->slots * sizeof(double), h->st));
  }
  for (int i = 0; i < 3; i++) {
    CKD(dev_alloc(&h->cnt[i], (size_t)h->nb * sizeof(unsigned)));
    CKD(cudaMemsetAsync(h->cnt[i], 0, (size_t)h->nb * sizeof(unsigned), h->st));
  }
  CKD(dev_alloc(&h->split, (size_t)h->nb * sizeof(double)));
  CKD(dev_alloc(&h->split_alt, (size_t)h->nb * sizeof(double)));
  {
    const size_t nc = (size_t)h->nb / 8 + 8;  // cells of >= 8 buckets
    CKD(dev_alloc(&h->knot_sum, nc * sizeof(double)));
    CKD(dev_alloc(&h->knot_x, nc * sizeof(double)));
    CKD(dev_alloc(&h->knot_y, nc * sizeof(double)));
    CKD(dev_alloc(&h->knot_n, nc * sizeof(unsigned)));
    CKD(cudaMemsetAsync(h->knot_sum, 0, nc * sizeof(double), h->st));
    CKD(cudaMemsetAsync(h->knot_n, 0, nc * sizeof(unsigned), h->st));
  }
  CKD(dev_alloc(&h->tot, (size_t)n_segments * sizeof(double)));
  CKD(dev_alloc(&h->ticket, 3 * sizeof(unsigned)));
  CKD(dev_alloc(&h->status, (size_t)h->nb * sizeof(unsigned)));
  CKD(dev_alloc(&h->desc, (size_t)h->nb * sizeof(Desc)));
  CKD(dev_alloc(&h->cpre, (size_t)h->nb * sizeof(unsigned)));
  CKD(dev_alloc(&h->cp_desc, (size_t)(count_prefix_tiles(h->nb) + 1) * sizeof(unsigned long long)));
  CKD(cudaMemsetAsync(h->cp_desc, 0, (size_t)(count_prefix_tiles(h->nb) + 1) * sizeof(unsigned long long), h->st));
  CKD(dev_alloc(&h->cp_ticket, sizeof(unsigned)));
  CKD(cudaMemsetAsync(h->cp_ticket, 0, sizeof(unsigned), h->st));
  if (!h->eqm) {
    const size_t mt = (size_t)mass_prefix_tiles(h->nb) + 1;
    CKD(dev_alloc(&h->magg, (size_t)h->nb * sizeof(ulonglong2)));
    CKD(dev_alloc(&h->mpre, (size_t)h->nb * sizeof(ulonglong2)));
    CKD(dev_alloc(&h->mp_desc, mt * sizeof(Desc)));
    CKD(dev_alloc(&h->mp_status, mt * sizeof(unsigned)));
    CKD(cudaMemsetAsync(h->mp_status, 0, mt * sizeof(unsigned), h->st));
    CKD(dev_alloc(&h->mp_ticket, sizeof(unsigned)));
    CKD(cudaMemsetAsync(h->mp_ticket, 0, sizeof(unsigned), h->st));
  }
  CKD(dev_alloc(&h->flags, 136 * sizeof(unsigned)));
  CKD(cudaMemsetAsync(h->flags, 0, 136 * sizeof(unsigned), h->st));
  CKD(dev_alloc(&h->offs, (size_t)h->nb * sizeof(unsigned long long)));
  CKD(dev_alloc(&h->epart, (size_t)h->nb * 4 * sizeof(double)));
  CKD(dev_alloc(&h->eout, 4 * sizeof(double)));
  CKD(cudaMallocHost(&h->h_flags, 136 * sizeof(unsigned)));
  CKD(cudaMallocHost(&h->h_eout, 4 * sizeof(double)));
  trace_mark(h->st, "create: device allocations");
  memset(h->h_flags, 0, 136 * sizeof(unsigned));
  CKD(cudaMemsetAsync(h->status, 0, (size_t)h->nb * sizeof(unsigned), h->st));
  if (dev_inputs) {
    CKD(cudaMemcpyAsync(h->x[0], x, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    CKD(cudaMemcpyAsync(h->v[0], v, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    if (!h->eqm) CKD(cudaMemcpyAsync(h->m[0], m, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  } else {
    const double *src[3] = {x, v, h->eqm ? nullptr : m};
    double *dst[3] = {h->x[0], h->v[0], h->eqm ? nullptr : h->m[0]};
    std::string uerr;
    // (large systems: through the read-out's bounce ring, which this handle keeps for its read-outs)
    if ((size_t)N * sizeof(double) >= ((size_t)64 << 20) && bounce_allowed() && !h->ring) h->ring = ring_acquire(h->device);
    if (upload_host_arrays(h->st, src, dst, 3, (size_t)N, uerr, h->ring)) { wendy_cuda_destroy(h); return set_err(WENDY_E_CUDA, uerr); }
  }
  CKD(cudaMemcpyAsync(h->tot, totmass, (size_t)n_segments * sizeof(double), cudaMemcpyHostToDevice, h->st));
  if (ids) CKD(cudaMemcpyAsync(h->id[0], ids, (size_t)N * sizeof(int), cudaMemcpyDefault, h->st));
  else launch_iota(h->st, h->id[0], N);
  trace_mark(h->st, "create: upload");
  if (reset_flags(h)) { std::string s = g_err; wendy_cuda_destroy(h); return set_err(WENDY_E_CUDA, s); }
  CKD(cudaGetLastError());
#undef CKD
  h->dense = true; h->cur = 0; h->ccur = 0;
  *out = h;
  return 0;
}

// Replace the total mass per segment given at creation.  Lets a host language compute it (the reference's
// numpy.sum over all masses, wendy/wendy.py:383: tens of milliseconds at N=1e8) while wendy_cuda_create
// validates and uploads; must be called before the first step.
int wendy_cuda_set_totmass(wendy_cuda_handle *h, const double *totmass) {
  if (!h || !totmass) return set_err(WENDY_E_ARG, "null argument");
  if (h->pending) return set_err(WENDY_E_ARG, "a call is in flight: wendy_cuda_step_end first");
  CK(cudaMemcpyAsync(h->tot, totmass, (size_t)h->nseg * sizeof(double), cudaMemcpyHostToDevice, h->st));
  CK(cudaStreamSynchronize(h->st));
  return 0;
}

int wendy_cuda_create(wendy_cuda_handle **out, long long N, const double *x, const double *v,
                      const double *m, const double *totmass, double omega2, int n_segments, int flags,
                      int cap, int fill, void *cuda_stream) {
  return create_impl(out, N, N, x, v, m, nullptr, totmass, omega2, n_segments, flags, cap, fill, cuda_stream);
}

// Same as wendy_cuda_create, from DEVICE arrays (no host staging).  m_dev may be
// NULL: every particle then has mass m0.
int wendy_cuda_create_dev(wendy_cuda_handle **out, long long N, const double *x_dev, const double *v_dev,
                          const double *m_dev, double m0, const double *totmass, double omega2, int n_segments,
                          int flags, int cap, int fill, void *cuda_stream) {
  return create_impl(out, N, N, x_dev, v_dev, m_dev, nullptr, totmass, omega2, n_segments, flags, cap, fill,
                     cuda_stream, true, m0);
}

// ---- sharded single system: this GPU owns the key range [bounds[rank], bounds[rank+1]) ------------------
static int create_shard_impl(wendy_cuda_handle **out, long long n_local, long long n_capacity, const double *x,
                             const double *v, const int *ids, double m0, double totmass, double omega2,
                             int nranks, int rank, const double *bounds, long long outbox_capacity,
                             void *cuda_stream, bool dev, const double *m_general, double sum_abs_global);
int wendy_cuda_create_shard(wendy_cuda_handle **out, long long n_local, long long n_capacity, const double *x,
                            const double *v, const int *ids, double m0, double totmass, double omega2,
                            int nranks, int rank, const double *bounds, long long outbox_capacity,
                            void *cuda_stream) {
  if (!ids || !bounds || nranks < 1 || rank < 0 || rank >= nranks || outbox_capacity < 1)
    return set_err(WENDY_E_ARG, "bad shard argument");
  return create_shard_impl(out, n_local, n_capacity, x, v, ids, m0, totmass, omega2, nranks, rank, bounds,
                           outbox_capacity, cuda_stream, false, nullptr, 0.);
}

// Unequal masses (host-orchestrated exchange only): m[n_local] already times twopiG; sum_abs_m_global = sum of |m|
// over ALL ranks (any common upper bound will do: it fixes the shared 128-bit fixed-point scale).  Migrant records
// are (x, v, id, m); the cumulative mass of a rank is offset by the exact total of the lower ranks
// (wendy_cuda_shard_mass_total / wendy_cuda_shard_set_mass_offset).
int wendy_cuda_create_shard_m(wendy_cuda_handle **out, long long n_local, long long n_capacity, const double *x,
                              const double *v, const double *m, const int *ids, double sum_abs_m_global,
                              double totmass, double omega2, int nranks, int rank, const double *bounds,
                              long long outbox_capacity, void *cuda_stream) {
  if (!m || !(sum_abs_m_global > 0.)) return set_err(WENDY_E_ARG, "bad shard argument");
  return create_shard_impl(out, n_local, n_capacity, x, v, ids, 0., totmass, omega2, nranks, rank, bounds,
                           outbox_capacity, cuda_stream, false, m, sum_abs_m_global);
}

// Same, from DEVICE arrays (the partition of multi.py runs on the GPU and hands its result over in place).
int wendy_cuda_create_shard_dev(wendy_cuda_handle **out, long long n_local, long long n_capacity,
                                const double *x_dev, const double *v_dev, const int *ids_dev, double m0,
                                double totmass, double omega2, int nranks, int rank, const double *bounds,
                                long long outbox_capacity, void *cuda_stream) {
  return create_shard_impl(out, n_local, n_capacity, x_dev, v_dev, ids_dev, m0, totmass, omega2, nranks, rank,
                           bounds, outbox_capacity, cuda_stream, true, nullptr, 0.);
}

static int create_shard_impl(wendy_cuda_handle **out, long long n_local, long long n_capacity, const double *x,
                             const double *v, const int *ids, double m0, double totmass, double omega2,
                             int nranks, int rank, const double *bounds, long long outbox_capacity,
                             void *cuda_stream, bool dev, const double *m_general, double sum_abs_global) {
  if (!ids || !bounds || nranks < 1 || rank < 0 || rank >= nranks || outbox_capacity < 1)
    return set_err(WENDY_E_ARG, "bad shard argument");
  int rc;
  if (m_general) {
    // unequal masses: the general path even if THIS range happens to hold equal ones; coarse buckets (the warp
    // kernel's sharded instance is equal-mass only)
    rc = create_impl(out, n_local, n_capacity, x, v, m_general, ids, &totmass, omega2, 1, WENDY_FLAG_GENERAL_MASSES,
                     tile_coarse_cap(), 0, cuda_stream);
  } else if (dev) {
    rc = create_impl(out, n_local, n_capacity, x, v, nullptr, ids, &totmass, omega2, 1, 0, 0, 0, cuda_stream,
                     true, m0);
  } else {
    std::vector<double> m((size_t)n_local, m0);
    rc = create_impl(out, n_local, n_capacity, x, v, m.data(), ids, &totmass, omega2, 1, 0, 0, 0, cuda_stream);
  }
  if (rc) return rc;
  H *h = *out;
  h->nranks = nranks; h->my_rank = rank; h->ocap = outbox_capacity;
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = dev_alloc(&h->bounds, (size_t)(nranks + 1) * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(h->bounds, bounds, (size_t)(nranks + 1) * sizeof(double), cudaMemcpyHostToDevice);
  size_t ob = (size_t)nranks * (size_t)outbox_capacity;
  if (m_general) {
    h->orec = 4;
    // one fixed-point scale for the whole system: the ranks' 128-bit mass totals must add exactly
    h->fxE = choose_fx_exponent(sum_abs_global);
  }
  if (e == cudaSuccess) e = dev_alloc(&h->out_rec, ob * (size_t)h->orec * sizeof(double));
  if (e == cudaSuccess) e = dev_alloc(&h->out_cnt, (size_t)nranks * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_out_cnt, (size_t)nranks * sizeof(unsigned));
  if (e == cudaSuccess) e = dev_alloc(&h->cid, (size_t)n_capacity * sizeof(int));
  if (e != cudaSuccess) { std::string msg = cudaGetErrorString(e); wendy_cuda_destroy(h); *out = nullptr; return set_err(WENDY_E_CUDA, msg); }
  if (h->cap != 256) h->fill = default_fill(h, h->cap);
  return 0;
}

// One leapfrog sub-step on the local range.  out_counts[p] = particles placed in the outbox of peer p.
int wendy_cuda_shard_substep(wendy_cuda_handle *h, double h_pre, double dt_kick, double dt_drift, double h_next,
                             long long pc_offset, unsigned *out_counts) {
  if (!h || !out_counts || !h->bounds) return set_err(WENDY_E_ARG, "not a shard handle");
  if (renew_epochs(h)) return WENDY_E_CUDA;
  h->pc_offset = pc_offset;
  for (int attempt = 0; attempt < 3; attempt++) {
    if (h->dense || !h->has_split || h->bucket_h != h_pre || h->rebuild_pending) {
      int rc = rebucket(h, h_pre);
      if (rc) return rc;
      h->rebuild_pending = false;
    }
    CK(cudaMemsetAsync(h->out_cnt, 0, (size_t)h->nranks * sizeof(unsigned), h->st));
    int cur0 = h->cur, ccur0 = h->ccur;
    launch_bucket_substep(h, h_pre, dt_kick, dt_drift, h_next, nullptr, nullptr);
    CK(cudaMemcpyAsync(h->h_out_cnt, h->out_cnt, (size_t)h->nranks * sizeof(unsigned), cudaMemcpyDeviceToHost, h->st));
    if (fetch_flags(h)) return WENDY_E_CUDA;
    if (h->h_flags[0] == 0xffffffffu) {
      long long gone = 0;
      for (int r = 0; r < h->nranks; r++) { out_counts[r] = h->h_out_cnt[r]; gone += h->h_out_cnt[r]; }
      h->N -= gone; h->seg_len = h->N;
      // adaptive layout (see finish_substeps): judged on this single sub-step
      if (h->adaptive && h->cap == 256 && (double)outside_total(h) > 0.5 * (double)h->N) {
        // switch to coarse buckets, but only once the incoming migrants have been injected: a
        // layout built now would see a boundary region depleted of the particles in flight
        h->want_cap = tile_coarse_cap();
        h->rebuild_pending = true;
      }
      h->n_outside += outside_total(h);
      memset(h->h_flags + 8, 0, 128 * sizeof(unsigned));
      CK(cudaMemsetAsync(h->flags + 8, 0, 128 * sizeof(unsigned), h->st));
      return 0;
    }
    h->n_fail++; h->n_sub--;
    fill_back_off(h);
    h->advect_on = advect_allowed();
    h->cur = cur0; h->ccur = ccur0; h->has_split = false;
    if (reset_flags(h)) return WENDY_E_CUDA;
  }
  return set_err(WENDY_E_OVERFLOW, "shard: bucket or outbox overflow persists after re-balancing");
}

// General masses: the exact 128-bit fixed-point total of the masses this rank holds NOW (low, high word), for the
// layout keyed on x + h_pre*v that the next wendy_cuda_shard_substep(h_pre, ...) will use (rebuilt here if needed, so
// that the bucket masses computed for the total are the ones that sub-step consumes).  The host adds the totals of
// the lower ranks (128-bit integer addition) and hands the sum back through wendy_cuda_shard_set_mass_offset.
int wendy_cuda_shard_mass_total(wendy_cuda_handle *h, double h_pre, unsigned long long *total2) {
  if (!h || !h->bounds || !total2) return set_err(WENDY_E_ARG, "not a shard handle");
  total2[0] = total2[1] = 0ull;
  if (h->eqm) return set_err(WENDY_E_ARG, "equal masses: the particle count is the offset");
  if (h->dense || !h->has_split || h->bucket_h != h_pre || h->rebuild_pending) {
    int rc = rebucket(h, h_pre);
    if (rc) return rc;
    h->rebuild_pending = false;
  }
  launch_mass_prefix(h->st, h->m[h->cur], h->cnt[h->ccur], h->cap, h->nb, h->fxE, h->magg, h->mpre, h->mp_desc,
                     h->mp_status, h->mp_ticket, h->seq++);
  h->n_launch += 2;
  ulonglong2 last[2];
  CK(cudaMemcpyAsync(&last[0], h->mpre + (h->nb - 1), sizeof(ulonglong2), cudaMemcpyDeviceToHost, h->st));
  CK(cudaMemcpyAsync(&last[1], h->magg + (h->nb - 1), sizeof(ulonglong2), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  const unsigned __int128 t = (((unsigned __int128)last[0].y << 64) | last[0].x) + (((unsigned __int128)last[1].y << 64) | last[1].x);
  total2[0] = (unsigned long long)t;
  total2[1] = (unsigned long long)(t >> 64);
  h->mpre_ready = true;
  return 0;
}

int wendy_cuda_shard_set_mass_offset(wendy_cuda_handle *h, unsigned long long lo, unsigned long long hi) {
  if (!h || !h->bounds) return set_err(WENDY_E_ARG, "not a shard handle");
  h->pm_lo = lo; h->pm_hi = hi;
  return 0;
}

// General masses: the masses of the local particles in the order of the LAST wendy_cuda_shard_read (compacted by
// the same kernel), for a global re-partition through the host.
int wendy_cuda_shard_read_masses(wendy_cuda_handle *h, double *m_host) {
  if (!h || !h->bounds || !m_host) return set_err(WENDY_E_ARG, "bad argument");
  if (h->eqm) return set_err(WENDY_E_ARG, "equal masses");
  if (h->pending || h->reader.joinable()) return set_err(WENDY_E_ARG, "finish the call / read-out in flight first");
  if (!h->xo) {
    CK(dev_alloc(&h->xo, (size_t)h->n_cap * sizeof(double)));
    CK(dev_alloc(&h->vo, (size_t)h->n_cap * sizeof(double)));
  }
  if (h->dense) {
    CK(copy_split(m_host, h->m[h->cur], (size_t)h->N * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  } else {
    launch_count_prefix(h->st, h->cnt[h->ccur], h->nb, h->cpre, h->cp_desc, h->cp_ticket, h->seq++);
    launch_compact(h->st, h->m[h->cur], h->m[h->cur], h->id[h->cur], h->cnt[h->ccur], h->cpre, h->cap, h->nb,
                   h->xo, h->vo, h->cid);
    h->n_launch += 2;
    CK(copy_split(m_host, h->xo, (size_t)h->N * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  }
  CK(cudaStreamSynchronize(h->st));
  return 0;
}

int wendy_cuda_shard_outbox(wendy_cuda_handle *h, double **records, long long *ocap) {
  if (!h || !h->bounds || !records || !ocap) return set_err(WENDY_E_ARG, "not a shard handle");
  *records = h->out_rec; *ocap = h->ocap;
  return 0;
}

// Append n particles (DEVICE arrays) whose keys lie in this shard's range to the current layout.
int wendy_cuda_shard_inject(wendy_cuda_handle *h, const double *records_dev, long long n) {
  if (!h || !h->bounds) return set_err(WENDY_E_ARG, "not a shard handle");
  if (n <= 0) return 0;
  if (h->N + n > h->n_cap) return set_err(WENDY_E_OVERFLOW, "shard capacity exceeded (global re-partition needed)");
  h->mpre_ready = false;
  if (h->dense || !h->has_split) return set_err(WENDY_E_ARG, "inject needs a layout");
  ScatterParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.packed_in = records_dev; sp.prec = h->orec; sp.min = nullptr;
  sp.cnt_in = nullptr; sp.n_dense = n; sp.h = h->bucket_h;
  int c = h->cur;
  sp.xout = h->x[c]; sp.vout = h->v[c]; sp.mout = h->eqm ? nullptr : h->m[c]; sp.idout = h->id[c];
  sp.cnt_out = h->cnt[h->ccur]; sp.split = h->split; sp.cap_out = h->cap; sp.nbps_out = h->nbps;
  sp.seg_len = h->n_cap + 1; sp.fail_seq = h->flags; sp.seq = h->seq++;
  // snapshot of the counts: an overflowing append is rolled back (slots beyond the counts are dead)
  CK(cudaMemcpyAsync(h->cnt[(h->ccur + 2) % 3], h->cnt[h->ccur], (size_t)h->nb * sizeof(unsigned),
                     cudaMemcpyDeviceToDevice, h->st));
  launch_scatter(h->st, sp, h->sm_count);
  h->n_launch++;
  if (fetch_flags(h)) return WENDY_E_CUDA;
  if (h->h_flags[0] != 0xffffffffu) {
    // some bucket near the range edge cannot take its share of the migrants: rebuild the layout from
    // the union of the local particles and the inbox
    if (reset_flags(h)) return WENDY_E_CUDA;
    CK(cudaMemcpyAsync(h->cnt[h->ccur], h->cnt[(h->ccur + 2) % 3], (size_t)h->nb * sizeof(unsigned),
                       cudaMemcpyDeviceToDevice, h->st));
    h->n_fail++;
    int rc = rebucket(h, h->bucket_h, records_dev, n);
    if (rc) return rc;
  } else {
    CK(cudaMemsetAsync(h->cnt[(h->ccur + 2) % 3], 0, (size_t)h->nb_alloc * sizeof(unsigned), h->st));
  }
  h->N += n; h->seg_len = h->N;
  return 0;
}

int wendy_cuda_shard_count(wendy_cuda_handle *h, long long *n_local) {
  if (!h || !n_local) return set_err(WENDY_E_ARG, "null argument");
  *n_local = h->N;
  return 0;
}

static int start_read_n(H *h, void *const *host, const void *const *src, const size_t *nby, int na, cudaStream_t st);
static int finish_read(H *h, cudaStream_t st);

// Compact (x, v, id) of the local particles to HOST arrays (capacity entries); *n = local count.
// _begin: compaction on the compute stream, device -> host copies on the private copy stream (a worker thread
// drives the bounce ring for pageable destinations); the caller may step the shard meanwhile.  _end waits.
int wendy_cuda_shard_read_begin(wendy_cuda_handle *h, double *x_host, double *v_host, int *id_host, long long *n) {
  if (!h || !h->bounds || !x_host || !v_host || !id_host || !n) return set_err(WENDY_E_ARG, "bad argument");
  if (h->pending) return set_err(WENDY_E_ARG, "finish the call in flight first");
  if (h->reader.joinable()) return set_err(WENDY_E_ARG, "a read-out is in flight: wendy_cuda_shard_read_end first");
  if (!h->st_copy) {
    CK(cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_unsort, cudaEventDisableTiming));
  }
  if (!h->xo) {
    trace_mark(h->st, "(read: start)");
    CK(dev_alloc(&h->xo, (size_t)h->n_cap * sizeof(double)));
    CK(dev_alloc(&h->vo, (size_t)h->n_cap * sizeof(double)));
    trace_mark(h->st, "read: staging allocation");
  }
  const void *src[3] = {h->xo, h->vo, h->cid};
  if (h->dense) {  // (staged as well: the state buffers are rewritten by the next call while the copy runs)
    CK(cudaMemcpyAsync(h->xo, h->x[h->cur], (size_t)h->N * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    CK(cudaMemcpyAsync(h->vo, h->v[h->cur], (size_t)h->N * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    CK(cudaMemcpyAsync(h->cid, h->id[h->cur], (size_t)h->N * sizeof(int), cudaMemcpyDeviceToDevice, h->st));
  } else {
    launch_count_prefix(h->st, h->cnt[h->ccur], h->nb, h->cpre, h->cp_desc, h->cp_ticket, h->seq++);
    launch_compact(h->st, h->x[h->cur], h->v[h->cur], h->id[h->cur], h->cnt[h->ccur], h->cpre, h->cap, h->nb,
                   h->xo, h->vo, h->cid);
    h->n_launch += 2;
  }
  CK(cudaEventRecord(h->ev_unsort, h->st));
  CK(cudaStreamWaitEvent(h->st_copy, h->ev_unsort, 0));
  void *const dst[3] = {x_host, v_host, id_host};
  const size_t nby[3] = {(size_t)h->N * sizeof(double), (size_t)h->N * sizeof(double), (size_t)h->N * sizeof(int)};
  *n = h->N;
  return start_read_n(h, dst, src, nby, 3, h->st_copy);
}

int wendy_cuda_shard_read_end(wendy_cuda_handle *h) {
  if (!h || !h->bounds) return set_err(WENDY_E_ARG, "not a shard handle");
  if (h->st_copy) {
    int rc = finish_read(h, h->st_copy);
    if (rc) return rc;
  }
  CK(cudaGetLastError());
  return 0;
}

int wendy_cuda_shard_read(wendy_cuda_handle *h, double *x_host, double *v_host, int *id_host, long long *n) {
  int rc = wendy_cuda_shard_read_begin(h, x_host, v_host, id_host, n);
  if (rc) return rc;
  rc = wendy_cuda_shard_read_end(h);
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->st));
  return 0;
}

// ---- sharded system: device-driven exchange over peer memory (peer.cuh) -------------------------------------
constexpr int PEER_NHIST = 4096;  // sub-steps per call whose particle counts are recorded (nleap limit in this mode)
static size_t comm_flag_bytes() { return (size_t)4 * PEER_MAX * sizeof(unsigned long long); }
static size_t comm_inbox_bytes(const H *h) { return (size_t)2 * h->nranks * (size_t)h->ocap * 3 * sizeof(double); }

// Allocate this rank's comm buffer (flags + inboxes; plain cudaMalloc so that it can be exported) and describe
// it: raw device pointer, size, and the 64-byte CUDA IPC handle another PROCESS on this node opens it with.
int wendy_cuda_shard_comm_export(wendy_cuda_handle *h, unsigned long long *ptr, unsigned long long *bytes,
                                 unsigned char *ipc_handle64) {
  if (!h || !h->bounds || !ptr || !bytes || !ipc_handle64) return set_err(WENDY_E_ARG, "not a shard handle");
  if (h->nranks > PEER_MAX) return set_err(WENDY_E_ARG, "peer exchange supports at most 16 ranks");
  if (!h->eqm) return set_err(WENDY_E_ARG, "peer exchange: equal masses only");
  if (!h->comm) {
    h->comm_bytes = 256 + comm_flag_bytes() + comm_inbox_bytes(h);
    CK(cudaMalloc(&h->comm, h->comm_bytes));
    CK(cudaMemset(h->comm, 0, h->comm_bytes));
  }
  *ptr = (unsigned long long)(uintptr_t)h->comm;
  *bytes = (unsigned long long)h->comm_bytes;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t ih;
  memset(&ih, 0, sizeof(ih));
  cudaError_t e = cudaIpcGetMemHandle(&ih, h->comm);
  if (e != cudaSuccess) cudaGetLastError();  // no IPC on this platform: ranks in one process can still use raw pointers
  memcpy(ipc_handle64, &ih, 64);
  return e == cudaSuccess ? 0 : 1;
}

// Map the peers' comm buffers and switch the handle to the device-driven exchange.  For rank r != my rank:
// raw_ptrs[r] != 0 -> that rank lives in THIS process (or the pointer is otherwise valid here); else its IPC
// handle (64 bytes at ipc_handles + 64 r) is opened.
int wendy_cuda_shard_comm_open(wendy_cuda_handle *h, const unsigned char *ipc_handles, const unsigned long long *raw_ptrs) {
  if (!h || !h->bounds || !h->comm || !raw_ptrs) return set_err(WENDY_E_ARG, "export the comm buffer first");
  PeerComm &pc = h->peer_host;
  memset(&pc, 0, sizeof(pc));
  pc.nranks = h->nranks; pc.my_rank = h->my_rank; pc.ocap = (unsigned)h->ocap;
  auto carve = [&](void *base, unsigned long long *&inf, unsigned long long *&cntf, double *&inbox) {
    char *b = (char *)base;
    inf = (unsigned long long *)b;
    cntf = inf + 2 * PEER_MAX;
    inbox = (double *)(b + 256 + comm_flag_bytes());
  };
  {
    unsigned long long *a, *b; double *c;
    carve(h->comm, a, b, c);
    pc.in_flag = a; pc.cnt_flag = b; pc.inbox = c;
  }
  for (int r = 0; r < h->nranks; r++) {
    void *base = nullptr;
    if (r == h->my_rank) base = h->comm;
    else if (raw_ptrs[r]) base = (void *)(uintptr_t)raw_ptrs[r];
    else {
      if (!ipc_handles) return set_err(WENDY_E_ARG, "no IPC handle for a peer in another process");
      cudaIpcMemHandle_t ih;
      memcpy(&ih, ipc_handles + (size_t)64 * r, 64);
      CK(cudaIpcOpenMemHandle(&base, ih, cudaIpcMemLazyEnablePeerAccess));
      h->peer_map[r] = base;
    }
    carve(base, pc.peer_in_flag[r], pc.peer_cnt_flag[r], pc.peer_inbox[r]);
  }
  if (!h->peer_scratch) {
    CK(dev_alloc(&h->peer_scratch, (PEER_MAX + 8) * sizeof(unsigned)));
    CK(dev_alloc(&h->peer_n, (size_t)(PEER_NHIST + 2) * sizeof(long long)));
    CK(dev_alloc(&h->peer_dev, sizeof(PeerComm)));
    CK(cudaMallocHost(&h->h_peer_n, (size_t)(PEER_NHIST + 2) * sizeof(long long)));
  }
  CK(cudaMemsetAsync(h->peer_scratch, 0, (PEER_MAX + 8) * sizeof(unsigned), h->st));
  CK(cudaMemsetAsync(h->peer_n, 0, (size_t)(PEER_NHIST + 2) * sizeof(long long), h->st));
  pc.out_cnt = h->peer_scratch; pc.cta_done = h->peer_scratch + PEER_MAX; pc.peer_stat = h->peer_scratch + PEER_MAX + 2;
  pc.n_local = h->peer_n; pc.n_hist = h->peer_n + 1;
  {
    const char *te = getenv("WENDY_B200_PEER_TIMEOUT_MS");
    const long long ms = te ? atoll(te) : 20000;
    pc.timeout_ns = (unsigned long long)(ms > 0 ? ms : 20000) * 1000000ull;
  }
  CK(cudaMemcpyAsync(h->peer_dev, &pc, sizeof(pc), cudaMemcpyHostToDevice, h->st));
  CK(cudaStreamSynchronize(h->st));
  tile_prepare_persistent();  // (no first-use set-up call behind a kernel that is waiting for a peer)
  h->peer_on = true;
  // the exchange lives in the persistent CTA kernel: coarse buckets from the start
  if (h->cap != tile_coarse_cap()) {
    h->want_cap = tile_coarse_cap();
    h->has_split = false;
  }
  return 0;
}

// (Re)seed the count flags: counts[r] = particles rank r owns now (from a host collective).  Needed once after
// the partition and after every rollback; between calls the flags of the last sub-step are already in place.
int wendy_cuda_shard_seed_counts(wendy_cuda_handle *h, const long long *counts) {
  if (!h || !h->peer_on || !counts) return set_err(WENDY_E_ARG, "peer exchange is not set up");
  unsigned long long w[PEER_MAX];
  const unsigned e = h->pepoch - 1u;
  for (int r = 0; r < PEER_MAX; r++) w[r] = r < h->nranks ? peer_pack(e, false, (unsigned)counts[r]) : 0ull;
  CK(cudaStreamSynchronize(h->st));
  CK(cudaMemcpy((void *)(h->peer_host.cnt_flag + (e & 1u) * PEER_MAX), w, sizeof(w), cudaMemcpyHostToDevice));
  const long long mine = counts[h->my_rank];
  CK(cudaMemcpy(h->peer_n, &mine, sizeof(mine), cudaMemcpyHostToDevice));
  CK(cudaMemset(h->peer_scratch, 0, (PEER_MAX + 8) * sizeof(unsigned)));
  h->peer_mig_seen = 0;
  h->N = mine; h->seg_len = mine;
  return 0;
}

// Make sure the layout is the one sub-step k0 of a call needs (first call, changed dt, after a rollback): local and
// synchronous.  step_begin does the same on demand; calling this first keeps every potentially device-synchronising
// CUDA call (allocations, first-time kernel loads) out of the window in which kernels wait for peers -- which matters
// when several ranks share ONE device (ranks as threads of a process: tests), not with one process per GPU.
int wendy_cuda_shard_prepare(wendy_cuda_handle *h, double dt, int k0) {
  if (!h || !h->peer_on) return set_err(WENDY_E_ARG, "peer exchange is not set up");
  if (h->pending) return set_err(WENDY_E_ARG, "a call is already in flight");
  const double h_pre = (k0 == 0) ? dt / 2. : 0.;
  if (h->dense || !h->has_split || h->bucket_h != h_pre || h->rebuild_pending) {
    int rc = rebucket(h, h_pre);
    if (rc) return rc;
    h->rebuild_pending = false;
  }
  return 0;
}

// Enqueue sub-steps [k0, nleap) of one call on every rank's own stream; nothing here waits for a peer's HOST.
// (A layout rebuild -- first call, changed dt, after a rollback -- is local and synchronous.)
int wendy_cuda_shard_step_begin(wendy_cuda_handle *h, double dt, int nleap, int k0) {
  if (!h || !h->peer_on) return set_err(WENDY_E_ARG, "peer exchange is not set up");
  if (nleap < 1 || nleap > PEER_NHIST || k0 < 0 || k0 >= nleap) return set_err(WENDY_E_ARG, "bad nleap");
  if (h->pending) return set_err(WENDY_E_ARG, "a call is already in flight");
  if (k0 == 0) {
    if (renew_epochs(h)) return WENDY_E_CUDA;
    h->p_seq.clear(); h->p_seq_inj.clear(); h->p_cur.clear(); h->p_ccur.clear();
  }
  h->p_seq.resize(k0); h->p_seq_inj.resize(k0); h->p_cur.resize(k0); h->p_ccur.resize(k0);
  h->p_dt = dt; h->p_nleap = nleap; h->p_k0 = k0;
  static const bool trace_on = getenv("WENDY_B200_SHARD_TRACE") != nullptr;
  if (trace_on) {
    while (h->tr_ev.size() < (size_t)4 * PEER_NHIST) {
      cudaEvent_t e; CK(cudaEventCreate(&e)); h->tr_ev.push_back(e);
    }
  }
  for (int k = k0; k < nleap; k++) {
    const double h_pre = (k == 0) ? dt / 2. : 0.;
    const bool last = (k == nleap - 1);
    if (h->dense || !h->has_split || h->bucket_h != h_pre || h->rebuild_pending) {
      if (k != k0) return set_err(WENDY_E_CUDA, "internal: layout lost inside a call");
      int rc = rebucket(h, h_pre);
      if (rc) return rc;
      h->rebuild_pending = false;
    }
    h->p_cur.push_back(h->cur);
    h->p_ccur.push_back(h->ccur);
    h->p_seq.push_back(h->seq);
    // step kernel (count prefix first)
    {
      TileParams p;
      fill_tile_params(h, p);
      p.h_pre = h_pre; p.dt_kick = dt; p.dt_drift = last ? dt / 2. : dt; p.h_next = last ? dt / 2. : 0.;
      p.kcall = k;
      if (trace_on) cudaEventRecord(h->tr_ev[4 * k], h->st);
      launch_count_prefix(h->st, p.cnt_in, h->nb, h->cpre, h->cp_desc, h->cp_ticket, p.epoch);
      if (trace_on) cudaEventRecord(h->tr_ev[4 * k + 1], h->st);
      p.cpre = h->cpre;
      launch_tile(h->st, h->cap, LOAD_BUCKET, EMIT_SPLITTER, 1, p);
      if (trace_on) cudaEventRecord(h->tr_ev[4 * k + 2], h->st);
      h->n_launch += 2;
      advance_after_tile(h);
      h->n_launch--;
      h->cur ^= 1; h->ccur = (h->ccur + 1) % 3; h->bucket_h = p.h_next;
      h->n_sub++;
    }
    // inject kernel
    {
      InjectParams q;
      memset(&q, 0, sizeof(q));
      q.peer = h->peer_dev; q.pepoch = h->pepoch; q.kcall = k; q.h = h->bucket_h;
      q.xout = h->x[h->cur]; q.vout = h->v[h->cur]; q.idout = h->id[h->cur];
      q.cnt_out = h->cnt[h->ccur]; q.split = h->split; q.cap = h->cap; q.nb = h->nb;
      q.fail_seq = h->flags; q.seq = h->seq; q.stats = h->flags + 1;
      h->p_seq_inj.push_back(h->seq);
      h->seq++;
      int grid = h->sm_count;
      if (const char *ge = getenv("WENDY_B200_PERSIST_GRID")) grid = std::max(1, std::min(grid, atoi(ge)));
      launch_peer_inject(h->st, q, grid);
      if (trace_on) cudaEventRecord(h->tr_ev[4 * k + 3], h->st);
      h->n_launch++;
    }
    h->pepoch++;
  }
  h->pending = true;
  return 0;
}

// Wait for the call; *k_fail = index of the first sub-step that did not complete on THIS rank (nleap: all did).
// The ranks must agree on min(k_fail) (one host collective per call) and, if it is < nleap, all roll back.
int wendy_cuda_shard_step_end(wendy_cuda_handle *h, int *k_fail, long long *n_local, long long *migrated_in) {
  if (!h || !h->peer_on || !k_fail) return set_err(WENDY_E_ARG, "peer exchange is not set up");
  if (!h->pending) return set_err(WENDY_E_ARG, "no call in flight");
  h->pending = false;
  const int nleap = h->p_nleap;
  CK(cudaMemcpyAsync(h->h_peer_n, h->peer_n, (size_t)(nleap + 2) * sizeof(long long), cudaMemcpyDeviceToHost, h->st));
  unsigned pstat[4] = {0u, 0u, 0u, 0u};  // [0] a wait timed out, [1] records received so far (wraps), [2] / [3] ~us CTA 0
                                         // of the step / inject kernels waited for the peers' flags (wrap)
  CK(cudaMemcpyAsync(pstat, h->peer_scratch + PEER_MAX + 2, sizeof(pstat), cudaMemcpyDeviceToHost, h->st));
  if (fetch_flags(h)) return WENDY_E_CUDA;
  if (!h->tr_ev.empty()) {
    for (int k = h->p_k0; k < nleap; k++) {
      for (int j = 0; j < 3; j++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->tr_ev[4 * k + j], h->tr_ev[4 * k + j + 1]) == cudaSuccess) h->tr_ms[j] += ms;
      }
      h->tr_n++;
    }
    h->tr_wait[0] = pstat[2]; h->tr_wait[1] = pstat[3];
  }
  if (pstat[0]) return set_err(WENDY_E_CUDA, "shard: timed out waiting for a peer GPU");
  if (migrated_in) *migrated_in = (long long)(unsigned)(pstat[1] - h->peer_mig_seen);
  h->peer_mig_seen = pstat[1];
  const unsigned f = h->h_flags[0];
  int kf = nleap;
  if (f != 0xffffffffu) {
    kf = -1;
    for (int k = 0; k < nleap; k++)
      if (h->p_seq[k] == f || h->p_seq_inj[k] == f) kf = k;
    if (kf < 0) {
      // flagged by the first step kernel of this call on behalf of a peer's failure in the previous call's last
      // inject: cannot happen (that call's step_end saw it on every rank), so treat it as sub-step k0
      kf = h->p_k0;
    }
    h->n_fail++;
  } else {
    h->shard_retry_k = -1;  // (a retried sub-step may be k = 0 of a call: only a completed call resets the escalation)
    h->N = h->h_peer_n[1 + nleap]; h->seg_len = h->N;
    if (h->h_flags[1] > (unsigned)(h->cap - (h->cap - h->fill) / 16)) {  // nearly full bucket: re-balance (local)
      fill_back_off(h);
      h->rebuild_pending = true;
    }
    h->n_outside += outside_total(h);
    memset(h->h_flags + 8, 0, 128 * sizeof(unsigned));
    CK(cudaMemsetAsync(h->flags + 8, 0, 128 * sizeof(unsigned), h->st));
  }
  *k_fail = kf;
  if (n_local) *n_local = (kf == nleap) ? h->N : h->h_peer_n[1 + (kf < 0 ? 0 : kf)];
  return 0;
}

// Every rank calls this with the agreed failing sub-step: the state goes back to the input of sub-step k (still
// intact: double buffering), the layout is rebuilt at the next step_begin(k0 = k).  Returns the particle count
// to seed the count flags with (after a host all-gather).
int wendy_cuda_shard_rollback(wendy_cuda_handle *h, int k, long long *n_local) {
  if (!h || !h->peer_on || k < 0 || k >= (int)h->p_cur.size()) return set_err(WENDY_E_ARG, "bad rollback");
  h->n_sub -= (h->p_nleap - k);
  h->cur = h->p_cur[k]; h->ccur = h->p_ccur[k];
  h->has_split = false;
  h->N = h->h_peer_n[1 + k]; h->seg_len = h->N;
  fill_back_off(h);
  if (h->shard_retry_k == k) h->shard_retry_n++; else { h->shard_retry_k = k; h->shard_retry_n = 0; }
  if (h->shard_retry_n >= 1) fill_escalate(h);  // the fresh layout of the first retry overflowed as well
  if (reset_flags(h)) return WENDY_E_CUDA;
  if (n_local) *n_local = h->N;
  return 0;
}

// Enqueue sub-steps [k0, nleap) of one reference call (asynchronous apart from layout rebuilds).
static int enqueue_substeps(H *h, double dt, int nleap, int k0) {
  h->p_seq.clear(); h->p_cur.clear(); h->p_ccur.clear();
  h->p_dt = dt; h->p_nleap = nleap; h->p_k0 = k0;
  if (h->small_ok && h->dense && k0 == 0) {
    // small systems: the whole call is ONE launch of the resident kernel (state stays dense, in id order)
    launch_small(h->st, h->x[h->cur], h->v[h->cur], h->m[h->cur], h->seg_len, h->nseg, h->tot, h->eqm ? 1 : 0,
                 h->m0, h->stab, h->omega2, h->fxE, dt, nleap);
    h->n_launch++;
    h->n_sub += nleap;
    h->last_dt = dt;
    return 0;
  }
  if (k0 == 0) {
    if (renew_epochs(h)) return WENDY_E_CUDA;
    h->n_outside += outside_total(h);
    memset(h->h_flags + 8, 0, 128 * sizeof(unsigned));
    CK(cudaMemsetAsync(h->flags + 8, 0, 128 * sizeof(unsigned), h->st));  // per-call window statistic
    if (h->adaptive && !h->coarse_default && h->last_dt != 0. && dt != h->last_dt && h->cap != 256) {
      h->want_cap = 256;  // a new time step: start again from the fine layout and re-measure
      h->has_split = false;
    }
    if (h->last_dt != 0. && dt != h->last_dt && !h->dw_fixed) h->dw = 256;  // ... and from the narrow destination window
    h->last_dt = dt;
  }
  for (int kk = k0; kk < nleap; kk++) {
    const double h_pre = (kk == 0) ? dt / 2. : 0.;
    const double dt_drift = (kk == nleap - 1) ? dt / 2. : dt;
    const bool radix = h->mode == WENDY_SORT_RADIX || h->radix_left > 0;
    if (!radix && (h->dense || !h->has_split || h->bucket_h != h_pre)) {
      int rc = rebucket(h, h_pre);  // synchronous; only at start-up, after radix sub-steps or a dt change
      if (rc) return rc;
    }
    h->p_seq.push_back(h->seq);
    h->p_cur.push_back(h->cur);
    h->p_ccur.push_back(h->ccur);
    if (radix) {
      int rc = launch_radix_substep(h, h_pre, dt, dt_drift, nullptr, nullptr);
      if (rc) return rc;
      if (h->radix_left > 0) h->radix_left--;
    } else {
      launch_bucket_substep(h, h_pre, dt, dt_drift, (kk == nleap - 1) ? dt / 2. : 0., nullptr, nullptr);
    }
  }
  return 0;
}

// Wait for the enqueued sub-steps; on a bucket overflow re-balance and resume from the failed one.
static int finish_substeps(H *h) {
  const double dt = h->p_dt;
  const int nleap = h->p_nleap;
  int last_fail = -1, attempts = 0;
  while (true) {
    if (fetch_flags(h)) return WENDY_E_CUDA;
    unsigned f = h->h_flags[0];
    if (f == 0xffffffffu) break;
    // launch f overflowed: the state it read is intact; everything after it did nothing
    int kf = -1;
    for (size_t i = 0; i < h->p_seq.size(); i++)
      if (h->p_seq[i] == f) kf = h->p_k0 + (int)i;
    if (kf < 0) return set_err(WENDY_E_CUDA, "internal: unknown failing launch");
    h->n_fail++;
    h->stage_ok = false;  // (a de-sort queued behind the failed launch saw an intermediate state)
    fill_back_off(h);
    h->fail_score += 1.;
    if (h->adaptive && h->fail_score >= 3.) {  // overflows keep coming: buy head-room with sparser buckets, if the storage allows
      fill_escalate(h);
      h->fail_score = 0.;
    }
    if (h->nseg == 1 && advect_allowed()) h->advect_on = true;  // bucket edges could not keep up with the flow: let them move with it
    h->n_sub -= (nleap - kf);
    h->cur = h->p_cur[kf - h->p_k0];
    h->ccur = h->p_ccur[kf - h->p_k0];
    h->has_split = false;  // force a rebuild for the key of sub-step kf
    if (reset_flags(h)) return WENDY_E_CUDA;
    attempts = (kf == last_fail) ? attempts + 1 : 1;
    last_fail = kf;
    int k = kf;
    if (attempts >= 2) {
      // even a freshly balanced layout overflows within this one sub-step (the density changes
      // by more than the bucket head-room): take it on the radix path, which cannot overflow
      double h_pre = (k == 0) ? dt / 2. : 0.;
      int rc = launch_radix_substep(h, h_pre, dt, (k == nleap - 1) ? dt / 2. : dt, nullptr, nullptr);
      if (rc) return rc;
      if (fetch_flags(h)) return WENDY_E_CUDA;
      h->n_radix_fallback++;
      // violent phase (the density changes faster than the bucket head-room allows): stay on the
      // radix path for a while before probing the bucket path again (exponential back-off)
      h->radix_streak = h->radix_streak ? std::min(16, 2 * h->radix_streak) : 1;
      h->radix_left = h->radix_streak;
      k++;
      attempts = 0;
      last_fail = -1;
    }
    int rc = enqueue_substeps(h, dt, nleap, k);
    if (rc) return rc;
  }
  if (last_fail < 0 && h->radix_left == 0 && h->n_fail == h->fail_mark) h->radix_streak = 0;  // a clean call
  if (h->n_fail == h->fail_mark) h->fail_score *= 0.97;
  h->fail_mark = h->n_fail;
  // adaptive layout: when most particles leave the 32-bucket window of the warp kernel every
  // sub-step (large N*dt), 2048-slot buckets with CTA-aggregated emission are faster
  if (h->adaptive && h->cap == 256 && !h->dense) {
    const double moved = (double)outside_total(h);
    if (moved > 0.5 * (double)h->N * (double)nleap) {
      h->want_cap = tile_coarse_cap();
      int rc = rebucket(h, h->bucket_h);
      if (rc) return rc;
      return 0;
    }
  }
  // cheap insurance: re-balance between calls when some bucket is nearly full
  if (h->mode != WENDY_SORT_RADIX && h->h_flags[1] > (unsigned)(h->cap - (h->cap - h->fill) / 16)) {
    fill_back_off(h);
    int rc = rebucket(h, h->bucket_h);
    if (rc) return rc;
  }
  return 0;
}

static int run_substeps(H *h, double dt, int nleap) {
  int rc = enqueue_substeps(h, dt, nleap, 0);
  if (rc) return rc;
  return finish_substeps(h);
}

int wendy_cuda_step_begin(wendy_cuda_handle *h, double dt_leap, int nleap) {
  if (!h) return set_err(WENDY_E_ARG, "null handle");
  if (nleap < 1) return set_err(WENDY_E_ARG, "nleap must be >= 1");
  if (h->pending) return set_err(WENDY_E_ARG, "a call is already in flight");
  // time of the call as the reference reports it (wendy/wendy.c time_begin / time_end around the integration):
  // device time between two events around the enqueued sub-steps; a caller that overlaps its read-out and its own
  // work with the call (as the generator does) must not see those in time_elapsed (round-1 advisor finding)
  if (!h->ev_call[0]) { CK(cudaEventCreate(&h->ev_call[0])); CK(cudaEventCreate(&h->ev_call[1])); }
  CK(cudaEventRecord(h->ev_call[0], h->st));
  int rc = enqueue_substeps(h, dt_leap, nleap, 0);
  if (rc) return rc;
  CK(cudaEventRecord(h->ev_call[1], h->st));
  h->pending = true;
  return 0;
}

int wendy_cuda_step_end(wendy_cuda_handle *h) {
  if (!h) return set_err(WENDY_E_ARG, "null handle");
  if (!h->pending) return 0;
  h->pending = false;
  return finish_substeps(h);
}

// Seconds the last wendy_cuda_step_begin / _step_end call spent on the device (between its first and last launch as
// enqueued; sub-steps re-run after an overflow are not included).  Call after wendy_cuda_step_end.
int wendy_cuda_last_call_seconds(wendy_cuda_handle *h, double *seconds) {
  if (!h || !seconds) return set_err(WENDY_E_ARG, "null argument");
  *seconds = 0.;
  if (!h->ev_call[0]) return 0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev_call[0], h->ev_call[1]) != cudaSuccess) { cudaGetLastError(); return 0; }
  *seconds = 1e-3 * (double)ms;
  return 0;
}

int wendy_cuda_step(wendy_cuda_handle *h, double dt_leap, int nleap, double *time_elapsed) {
  if (!h) return set_err(WENDY_E_ARG, "null handle");
  if (nleap < 1) return set_err(WENDY_E_ARG, "nleap must be >= 1");
  if (h->pending) return set_err(WENDY_E_ARG, "a call is already in flight (wendy_cuda_step_end first)");
  auto t0 = std::chrono::steady_clock::now();
  int rc = run_substeps(h, dt_leap, nleap);
  if (time_elapsed)
    *time_elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

int wendy_cuda_force_positions(wendy_cuda_handle *h, double dt_leap, int first_substep, double **x_dev,
                               long long *n_slots) {
  if (!h || !x_dev || !n_slots) return set_err(WENDY_E_ARG, "null argument");
  // The leading half drift is materialised in the stored positions below; a caller that comes back for the same
  // sub-step after WENDY_RETRY (the layout overflowed and was rebuilt) must not have it applied a second time.
  if (first_substep && !h->ext_async && renew_epochs(h)) return WENDY_E_CUDA;
  double need_h = (first_substep && !h->ext_half_done) ? dt_leap / 2. : 0.;
  if (h->mode == WENDY_SORT_RADIX) {
    // radix mode keeps no splitters; give it a (compact or bucketed) layout to expose
    if (h->dense) { int rc = rebucket(h, need_h); if (rc) return rc; }
  } else if (h->dense || !h->has_split || h->bucket_h != need_h) {
    int rc = rebucket(h, need_h);
    if (rc) return rc;
  }
  if (need_h != 0.) {  // materialise the half drift: keys become the stored positions
    launch_apply_drift(h->st, h->x[h->cur], h->v[h->cur], need_h, h->cnt[h->ccur], h->cap, h->nb);
    h->n_launch++;
    h->bucket_h = 0.;
    h->ext_half_done = true;
  }
  *x_dev = h->x[h->cur];
  *n_slots = (long long)h->slots;
  return 0;
}

int wendy_cuda_substep(wendy_cuda_handle *h, double dt_kick, double dt_drift, double h_next,
                       const double *a_ext_dev) {
  if (!h) return set_err(WENDY_E_ARG, "null handle");
  if (h->dense) return set_err(WENDY_E_ARG, "call wendy_cuda_force_positions first");
  int cur0 = h->cur, ccur0 = h->ccur;
  if (h->mode == WENDY_SORT_RADIX) {
    int rc = launch_radix_substep(h, 0., dt_kick, dt_drift, a_ext_dev, nullptr);
    if (rc) return rc;
    // radix mode applies no trailing re-bucket key: the next call's half drift goes via h_pre
    if (fetch_flags(h)) return WENDY_E_CUDA;
    h->ext_half_done = false;
    return 0;
  }
  if (!h->has_split || h->bucket_h != 0.) return set_err(WENDY_E_ARG, "layout is not keyed on the stored positions");
  if (h->ext_fail_streak >= 2) {
    // even a freshly balanced layout overflowed within this one sub-step: take it on the radix path, which
    // cannot overflow (a_ext is in the storage order of the current layout, which is what that path gathers from)
    int rc = launch_radix_substep(h, 0., dt_kick, dt_drift, a_ext_dev, nullptr);
    if (rc) return rc;
    if (fetch_flags(h)) return WENDY_E_CUDA;
    h->n_radix_fallback++;
    h->ext_fail_streak = 0;
    h->ext_half_done = false;
    return 0;
  }
  launch_bucket_substep(h, 0., dt_kick, dt_drift, h_next, a_ext_dev, nullptr);
  if (fetch_flags(h)) return WENDY_E_CUDA;
  if (h->h_flags[0] != 0xffffffffu) {
    h->n_fail++; h->n_sub--;
    h->ext_fail_streak++;
    fill_back_off(h);
    h->cur = cur0; h->ccur = ccur0;
    if (reset_flags(h)) return WENDY_E_CUDA;
    int rc = rebucket(h, 0.);
    if (rc) return rc;
    return WENDY_RETRY;
  }
  h->ext_half_done = false;
  h->ext_fail_streak = 0;
  if (h->h_flags[1] > (unsigned)(h->cap - (h->cap - h->fill) / 16)) {
    // nearly full bucket: re-balance now; the caller's next force_positions sees the new slots
    fill_back_off(h);
    int rc = rebucket(h, h_next);
    if (rc) return rc;
  }
  return 0;
}

// ---- external force, asynchronous: the sub-steps of one call are enqueued without a host round trip each ----------
// wendy_cuda_ext_begin; per sub-step wendy_cuda_force_positions -> (caller evaluates F on the stream) ->
// wendy_cuda_substep_async; then wendy_cuda_ext_end waits once.  A bucket overflow in sub-step k makes every launch
// queued behind it a no-op (fail_seq); ext_end then restores the input of sub-step k, rebuilds the layout and
// reports k: the caller re-runs sub-steps k.. through the synchronous wendy_cuda_substep (which may retry or take the
// radix path).  The a_ext arrays must stay alive until ext_end.
int wendy_cuda_ext_begin(wendy_cuda_handle *h) {
  if (!h) return set_err(WENDY_E_ARG, "null handle");
  if (h->pending) return set_err(WENDY_E_ARG, "a call is already in flight");
  if (renew_epochs(h)) return WENDY_E_CUDA;
  h->p_seq.clear(); h->p_cur.clear(); h->p_ccur.clear();
  h->ext_async = true;
  return 0;
}

int wendy_cuda_substep_async(wendy_cuda_handle *h, double dt_kick, double dt_drift, double h_next,
                             const double *a_ext_dev) {
  if (!h || !h->ext_async) return set_err(WENDY_E_ARG, "wendy_cuda_ext_begin first");
  if (h->dense || h->mode == WENDY_SORT_RADIX || !h->has_split || h->bucket_h != 0.)
    return set_err(WENDY_E_ARG, "layout is not keyed on the stored positions");
  h->p_seq.push_back(h->seq); h->p_cur.push_back(h->cur); h->p_ccur.push_back(h->ccur);
  launch_bucket_substep(h, 0., dt_kick, dt_drift, h_next, a_ext_dev, nullptr);
  return 0;
}

// *k_done = number of sub-steps (since ext_begin) that completed; == the number enqueued unless one overflowed.
int wendy_cuda_ext_end(wendy_cuda_handle *h, int *k_done) {
  if (!h || !k_done || !h->ext_async) return set_err(WENDY_E_ARG, "no asynchronous call in flight");
  h->ext_async = false;
  const int nq = (int)h->p_seq.size();
  if (fetch_flags(h)) return WENDY_E_CUDA;
  *k_done = nq;
  if (h->h_flags[0] != 0xffffffffu) {
    int kf = 0;
    for (int k = 0; k < nq; k++)
      if (h->p_seq[k] == h->h_flags[0]) kf = k;
    h->n_fail++;
    h->n_sub -= (nq - kf);
    h->ext_fail_streak = 1;
    fill_back_off(h);
    h->cur = h->p_cur[kf]; h->ccur = h->p_ccur[kf];
    if (reset_flags(h)) return WENDY_E_CUDA;
    int rc = rebucket(h, 0.);
    if (rc) return rc;
    *k_done = kf;
    return 0;
  }
  h->ext_half_done = false;
  h->ext_fail_streak = 0;
  if (nq > 0 && h->h_flags[1] > (unsigned)(h->cap - (h->cap - h->fill) / 16)) {
    fill_back_off(h);  // nearly full bucket: re-balance now (the layout is keyed on bucket_h = the last h_next)
    int rc = rebucket(h, h->bucket_h);
    if (rc) return rc;
  }
  return 0;
}

int wendy_cuda_read_dev(wendy_cuda_handle *h, double *x_dev, double *v_dev) {
  if (!h) return set_err(WENDY_E_ARG, "null handle");
  if (h->pending) return set_err(WENDY_E_ARG, "a call is in flight: wendy_cuda_step_end first");
  if (!h->xo) {
    trace_mark(h->st, "(read: start)");
    CK(dev_alloc(&h->xo, (size_t)h->n_cap * sizeof(double)));
    CK(dev_alloc(&h->vo, (size_t)h->n_cap * sizeof(double)));
    trace_mark(h->st, "read: staging allocation");
  }
  double *xd = x_dev ? x_dev : h->xo, *vd = v_dev ? v_dev : h->vo;
  if (h->dense) {
    CK(cudaMemcpyAsync(xd, h->x[h->cur], (size_t)h->N * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    CK(cudaMemcpyAsync(vd, h->v[h->cur], (size_t)h->N * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  } else {
    launch_unsort(h->st, h->x[h->cur], h->v[h->cur], h->id[h->cur], h->cnt[h->ccur], h->cap, h->nb, xd, vd);
    h->n_launch++;
  }
  CK(cudaGetLastError());
  return 0;
}

// Start the device -> host copy of the de-sorted staging arrays on stream st: page-locked destinations get
// plain (split) copies; large pageable ones are filled through the bounce ring by a worker thread, so that
// the caller can go on enqueueing work.  finish_read() waits for both.
static int start_read_n(H *h, void *const *host, const void *const *src, const size_t *nby, int na, cudaStream_t st) {
  if (h->reader.joinable()) return set_err(WENDY_E_ARG, "a read-out is in flight: wendy_cuda_read_end first");
  std::vector<void *> d;
  std::vector<const void *> sr;
  std::vector<size_t> nb;
  for (int a = 0; a < na; a++) {
    if (!host[a] || !nby[a]) continue;
    if (nby[a] >= ((size_t)4 << 20) && bounce_allowed() && !host_range_is_pinned(host[a], nby[a])) {
      d.push_back(host[a]); sr.push_back(src[a]); nb.push_back(nby[a]);
    } else {
      CK(copy_split(host[a], src[a], nby[a], cudaMemcpyDeviceToHost, st));
    }
  }
  if (d.empty()) return 0;
  if (!h->ring) h->ring = ring_acquire(h->device);
  if (!h->ring) {  // no page-locked memory to be had: let the driver stage the copies
    for (size_t a = 0; a < d.size(); a++) CK(copy_split(d[a], sr[a], nb[a], cudaMemcpyDeviceToHost, st));
    return 0;
  }
  h->reader_rc = 0;
  H *hh = h;
  h->reader = std::thread([hh, d, sr, nb, st]() {
    cudaSetDevice(hh->device);
    hh->reader_rc = bounce_d2h(hh->ring, st, d.data(), sr.data(), nb.data(), (int)d.size());
  });
  return 0;
}

static int start_read(H *h, double *x_host, double *v_host, cudaStream_t st) {
  const size_t bytes = (size_t)h->N * sizeof(double);
  void *const host[2] = {x_host, v_host};
  const void *const src[2] = {h->stage_sel ? h->xo2 : h->xo, h->stage_sel ? h->vo2 : h->vo};
  const size_t nby[2] = {bytes, bytes};
  return start_read_n(h, host, src, nby, 2, st);
}

static int finish_read(H *h, cudaStream_t st) {
  if (h->reader.joinable()) {
    h->reader.join();
    if (h->reader_rc) return set_err(WENDY_E_CUDA, "device -> host copy through the bounce buffers failed");
  }
  CK(cudaStreamSynchronize(st));
  return 0;
}

int wendy_cuda_read(wendy_cuda_handle *h, double *x_host, double *v_host) {
  if (h) { h->stage_sel = 0; h->stage_ok = false; }
  int rc = wendy_cuda_read_dev(h, nullptr, nullptr);
  if (rc) return rc;
  rc = start_read(h, x_host, v_host, h->st);
  if (rc) return rc;
  return finish_read(h, h->st);
}

// Overlapped read-out: de-sort into staging on the compute stream, D2H on a private copy stream.
// Between _begin and _end the caller may enqueue the next call (wendy_cuda_step_begin).
// De-sort AHEAD of the read-out: enqueue, behind the sub-steps of the call in flight, the de-sort of the state that
// call will leave, into the staging set the running device -> host copy is NOT reading.  The next
// wendy_cuda_read_begin then starts copying at once instead of de-sorting first (8.9 ms at N=1e8, which otherwise
// sits between the end of a call and the start of its 33 ms PCIe copy).  If a sub-step of the call has to be re-run
// the staged copy is dropped and read_begin de-sorts as before.  Costs a second staging set (16 B per particle).
int wendy_cuda_stage_ahead(wendy_cuda_handle *h) {
  if (!h) return set_err(WENDY_E_ARG, "null handle");
  if (h->bounds) return set_err(WENDY_E_ARG, "not for shards");
  if (!h->xo) {
    CK(dev_alloc(&h->xo, (size_t)h->n_cap * sizeof(double)));
    CK(dev_alloc(&h->vo, (size_t)h->n_cap * sizeof(double)));
  }
  if (!h->xo2) {
    if (dev_alloc(&h->xo2, (size_t)h->n_cap * sizeof(double)) != cudaSuccess ||
        dev_alloc(&h->vo2, (size_t)h->n_cap * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();  // no room for the second set: read_begin de-sorts as before
      dev_free(h->xo2); h->xo2 = nullptr; h->vo2 = nullptr;
      return 0;
    }
  }
  const int t = h->stage_sel ^ 1;
  double *xd = t ? h->xo2 : h->xo, *vd = t ? h->vo2 : h->vo;
  if (h->dense) {
    CK(cudaMemcpyAsync(xd, h->x[h->cur], (size_t)h->N * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    CK(cudaMemcpyAsync(vd, h->v[h->cur], (size_t)h->N * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  } else {
    launch_unsort(h->st, h->x[h->cur], h->v[h->cur], h->id[h->cur], h->cnt[h->ccur], h->cap, h->nb, xd, vd);
    h->n_launch++;
  }
  h->stage_ok = true;
  return 0;
}

int wendy_cuda_read_begin(wendy_cuda_handle *h, double *x_host, double *v_host) {
  if (!h) return set_err(WENDY_E_ARG, "null handle");
  if (h->pending) return set_err(WENDY_E_ARG, "finish the call in flight first");
  if (h->reader.joinable()) return set_err(WENDY_E_ARG, "a read-out is in flight: wendy_cuda_read_end first");
  if (!h->st_copy) {
    CK(cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_unsort, cudaEventDisableTiming));
  }
  if (h->stage_ok && h->xo2) {
    // the de-sorted state is already in the other staging set (wendy_cuda_stage_ahead; wendy_cuda_step_end has
    // waited for the stream): copy from there
    h->stage_sel ^= 1;
    h->stage_ok = false;
  } else {
    h->stage_ok = false;
    h->stage_sel = 0;
    int rc = wendy_cuda_read_dev(h, nullptr, nullptr);
    if (rc) return rc;
  }
  CK(cudaEventRecord(h->ev_unsort, h->st));
  CK(cudaStreamWaitEvent(h->st_copy, h->ev_unsort, 0));
  return start_read(h, x_host, v_host, h->st_copy);
}

int wendy_cuda_read_end(wendy_cuda_handle *h) {
  if (!h) return set_err(WENDY_E_ARG, "null handle");
  if (h->st_copy) {
    int rc = finish_read(h, h->st_copy);
    if (rc) return rc;
  }
  trace_mark(h->st, "read_end: D2H done");
  return 0;
}

int wendy_cuda_energy(wendy_cuda_handle *h, double out[4]) {
  if (!h || !out) return set_err(WENDY_E_ARG, "null argument");
  if (h->pending) return set_err(WENDY_E_ARG, "a call is in flight: wendy_cuda_step_end first");
  // sort the synchronised positions (the layout may be keyed on x + h*v) and run the tile
  // kernel in diagnostic mode through the sorted permutation
  if (make_keys(h, 0., VAL_SLOT)) return WENDY_E_CUDA;
  unsigned seg_div = h->dense ? (unsigned)h->seg_len : (unsigned)((long long)h->nbps * h->cap);
  int res = radix_sort_pairs(h->st, h->rs, (size_t)h->N, seg_bits(h), seg_div);
  h->n_launch += 5 * (8 + (seg_bits(h) + 7) / 8);
  TileParams p;
  fill_tile_params(h, p);
  p.perm = h->rs.val[res];
  p.cnt_out = nullptr; p.cnt_zero = nullptr; p.ticket_zero = h->ticket + (h->tcur + 2) % 3;
  p.energy_part = h->epart;
  launch_tile(h->st, h->cap, LOAD_GATHER, EMIT_NONE, 0, p);
  advance_after_tile(h);
  launch_reduce_energy(h->st, h->epart, h->nb, h->eout);
  h->n_launch++;
  CK(cudaMemcpyAsync(h->h_eout, h->eout, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  CK(cudaGetLastError());
  for (int i = 0; i < 4; i++) out[i] = h->h_eout[i];
  return 0;
}

// Host evaluation of the closed-form serial sum the equal-mass kernels use (serialsum.cuh): out[i] = cum below
// sorted position k0 + i, i.e. the reference's cumulmass[k0 + i] (wendy/wendy.c:359-360) for N equal masses m0.
// No GPU involved; lets a caller (and the CPU test-suite) check the table against a plain serial loop.
int wendy_serial_cum(double m0, long long k0, long long n, double *out) {
  if (k0 < 0 || n < 0 || k0 + n > (1ll << 31) || (n > 0 && !out)) return set_err(WENDY_E_ARG, "bad argument");
  SerialTab *T = new SerialTab;
  if (serial_tab_build(*T, m0, 1ll << 31)) { delete T; return set_err(WENDY_E_ARG, "no closed form for this m0"); }
  int s = 0;
  for (long long i = 0; i < n; i++) {
    const long long k = k0 + i;
    while (k >= T->i0[s + 1]) s++;
    volatile double prod = (double)(k - T->i0[s]) * T->inc[s];
    out[i] = T->c0[s] + prod;
  }
  const int pieces = T->nseg;
  delete T;
  return pieces;
}

int wendy_cuda_stats(wendy_cuda_handle *h, long long *out, int n) {
  if (!h || !out) return set_err(WENDY_E_ARG, "null argument");
  long long s[9] = {h->n_sub, h->n_rebuild, h->n_fail, h->max_cnt, h->n_outside + outside_total(h), h->n_launch,
                    (long long)h->cap, (long long)h->nb, h->n_radix_fallback};
  for (int i = 0; i < n && i < 9; i++) out[i] = s[i];
  return 0;
}

int wendy_cuda_pin(void *host_ptr, unsigned long long bytes) {
  if (!host_ptr || !bytes) return set_err(WENDY_E_ARG, "null argument");
  std::vector<std::pair<void *, size_t>> done;
  char *q = (char *)host_ptr;
  size_t left = (size_t)bytes;
  while (left) {
    const size_t piece = pin_piece(q, left);
    cudaError_t e = cudaHostRegister(q, piece, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
      for (auto &d : done) cudaHostUnregister(d.first);
      cudaGetLastError();
      return set_err(WENDY_E_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
    }
    done.push_back(std::make_pair((void *)q, piece));
    q += piece; left -= piece;
  }
  std::lock_guard<std::mutex> lk(g_pin_mu);
  g_pins[host_ptr] = done;
  return 0;
}

int wendy_cuda_unpin(void *host_ptr) {
  if (!host_ptr) return set_err(WENDY_E_ARG, "null argument");
  std::vector<std::pair<void *, size_t>> pieces;
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    auto it = g_pins.find(host_ptr);
    if (it == g_pins.end()) return set_err(WENDY_E_ARG, "not a range registered by wendy_cuda_pin");
    pieces = it->second;
    g_pins.erase(it);
  }
  for (auto &d : pieces) CK(cudaHostUnregister(d.first));
  return 0;
}

int wendy_cuda_debug_layout(wendy_cuda_handle *h, unsigned *counts, double *splitters, int nb_max) {
  if (!h) return set_err(WENDY_E_ARG, "null handle");
  if (h->dense) return 0;
  int nb = h->nb < nb_max ? h->nb : nb_max;
  CK(cudaStreamSynchronize(h->st));
  if (counts) CK(cudaMemcpy(counts, h->cnt[h->ccur], (size_t)nb * sizeof(unsigned), cudaMemcpyDeviceToHost));
  if (splitters) CK(cudaMemcpy(splitters, h->split, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost));
  return nb;
}

int wendy_cuda_argsort(const double *x_host, long long N, int *perm_out) {
  if (!x_host || !perm_out || N < 0 || N >= (1ll << 31)) return set_err(WENDY_E_ARG, "bad argument");
  if (N == 0) return 0;
  H tmp;  // only the radix scratch is used
  double *dx = nullptr;
  if (alloc_radix(&tmp, (size_t)N)) return WENDY_E_CUDA;
  int rc = 0;
  do {
    if (dev_alloc(&dx, (size_t)N * sizeof(double)) != cudaSuccess) { rc = set_err(WENDY_E_CUDA, "cudaMalloc"); break; }
    cudaMemcpy(dx, x_host, (size_t)N * sizeof(double), cudaMemcpyHostToDevice);
    launch_make_keys(nullptr, dx, nullptr, 0., nullptr, nullptr, 0, 0, N, tmp.rs.key[0], tmp.rs.val[0],
                     VAL_INDEX, N, 0);
    int res = radix_sort_pairs(nullptr, tmp.rs, (size_t)N, 0, 1u);
    cudaError_t e = cudaMemcpy(perm_out, tmp.rs.val[res], (size_t)N * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) rc = set_err(WENDY_E_CUDA, cudaGetErrorString(e));
  } while (0);
  dev_free(dx);
  for (int i = 0; i < 2; i++) { dev_free(tmp.rs.key[i]); dev_free(tmp.rs.val[i]); }
  dev_free(tmp.rs.table); dev_free(tmp.rs.sums);
  return rc;
}

// ---- diagnostics on arbitrary (x, v, m): potential(y) and per-particle energies ------------------------------
// Reference wendy/wendy.py:494-517 (potential) and :466-470 (energy(individual=True)); every array may be a
// host or a device pointer, m is the UNSCALED mass and twopiG multiplies the sum as in the reference.
int wendy_cuda_potential(const double *y, long long Y, const double *x, const double *m, long long N,
                         double twopiG, double omega2, double *out, void *cuda_stream) {
  if (Y < 0 || N < 1 || N >= (1ll << 31) || !x || !m || (Y > 0 && (!y || !out)))
    return set_err(WENDY_E_ARG, "bad argument");
  std::string err;
  if (wendy::potential_eval((cudaStream_t)cuda_stream, x, nullptr, m, N, y, Y, twopiG, omega2, out, 0, err))
    return set_err(WENDY_E_CUDA, err);
  return 0;
}

int wendy_cuda_energy_individual(const double *x, const double *v, const double *m, long long N, double twopiG,
                                 double omega2, double *out, void *cuda_stream) {
  if (N < 1 || N >= (1ll << 31) || !x || !v || !m || !out) return set_err(WENDY_E_ARG, "bad argument");
  std::string err;
  if (wendy::potential_eval((cudaStream_t)cuda_stream, x, v, m, N, nullptr, 0, twopiG, omega2, out, 1, err))
    return set_err(WENDY_E_CUDA, err);
  return 0;
}

// ---- compat export -------------------------------------------------------------------------------------
// The reference's own entry point (wendy/wendy.c:385-393) on host pointers.  State does not
// survive the call (the reference C side is stateless as well); use the resident API for
// throughput.  `a` and `cumulmass` are scratch in the reference (never read by its Python
// side, wendy/wendy.py:424-437) and are left untouched; `err` is never written on this path
// in the reference either.  On a CUDA failure *err is set to -1 and the message is available
// from wendy_cuda_last_error().
void _wendy_nbody_approx_onestep(int N, struct wendy_array_w_index *xi, double *x, double *v, double *m,
                                 double *a, double totmass, double dt, int nleap, double *t0,
                                 double omega2, double (*ext_force)(int, double *, double, double *),
                                 int sort_type, int *err, double *time_elapsed, double *cumulmass) {
  (void)a; (void)cumulmass; (void)sort_type;
  auto tb = std::chrono::steady_clock::now();
  if (N <= 0) return;
  std::vector<double> xs((size_t)N);
  for (int i = 0; i < N; i++) xs[xi[i].idx] = xi[i].val;  // xi is the authoritative position state
  const char *env = getenv("WENDY_B200_SORT");
  int flags = (env && !strcmp(env, "radix")) ? WENDY_SORT_RADIX : WENDY_SORT_AUTO;
  const char *ecap = getenv("WENDY_B200_CAP");
  H *h = nullptr;
  int rc = wendy_cuda_create(&h, N, xs.data(), v, m, &totmass, omega2, 1, flags, ecap ? atoi(ecap) : 0, 0, nullptr);
  std::vector<int> rank((size_t)N);
  if (!rc) rc = (dev_alloc(&h->rank, (size_t)N * sizeof(int)) == cudaSuccess) ? 0 : WENDY_E_CUDA;
  if (!rc && !ext_force) {
    // sub-step by sub-step so that the sort order of the LAST force evaluation is recorded
    for (int k = 0; k < nleap && !rc; k++) {
      double *xd; long long ns;
      rc = wendy_cuda_force_positions(h, dt, k == 0, &xd, &ns);
      if (rc) break;
      int tries = 0;
      do {
        int cur0 = h->cur, ccur0 = h->ccur;
        if (h->mode == WENDY_SORT_RADIX || tries >= 2) {  // (two overflows of fresh layouts: the radix path cannot overflow)
          rc = launch_radix_substep(h, 0., dt, k == nleap - 1 ? dt / 2. : dt, nullptr, h->rank);
          if (!rc) rc = fetch_flags(h);
        } else {
          launch_bucket_substep(h, 0., dt, k == nleap - 1 ? dt / 2. : dt, 0., nullptr, h->rank);
          rc = fetch_flags(h);
          if (!rc && h->h_flags[0] != 0xffffffffu) {
            h->cur = cur0; h->ccur = ccur0;
            rc = reset_flags(h);
            if (!rc) rc = rebucket(h, 0.);
            if (!rc) rc = WENDY_RETRY;
          }
        }
      } while (rc == WENDY_RETRY && ++tries < 4);
      h->ext_half_done = false;
    }
  } else if (!rc) {
    std::vector<double> xh((size_t)N), ah((size_t)N);
    double *a_id = nullptr, *a_slot = nullptr;
    if (dev_alloc(&a_id, (size_t)N * sizeof(double)) != cudaSuccess) rc = WENDY_E_CUDA;
    for (int k = 0; k < nleap && !rc; k++) {
      int tries = 0;
      do {
        double *xd; long long ns;
        rc = wendy_cuda_force_positions(h, dt, k == 0, &xd, &ns);
        if (rc) break;
        if (!a_slot && dev_alloc(&a_slot, (size_t)ns * sizeof(double)) != cudaSuccess) { rc = WENDY_E_CUDA; break; }
        rc = wendy_cuda_read(h, xh.data(), nullptr);  // positions at force time, particle order
        if (rc) break;
        if (N > 10) {  // EXTERNAL_SWITCH, wendy/wendy.h:9-11 and wendy/wendy.c:362-370
          ext_force(N, xh.data(), *t0, ah.data());
        } else {
          for (int i = 0; i < N; i++) ah[i] = ext_force(1, &xh[i], *t0, nullptr);
        }
        cudaMemcpyAsync(a_id, ah.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice, h->st);
        launch_gather_by_id(h->st, a_id, h->id[h->cur], h->cnt[h->ccur], h->cap, h->nb, a_slot);
        if (h->mode == WENDY_SORT_RADIX || tries >= 2) {
          rc = launch_radix_substep(h, 0., dt, k == nleap - 1 ? dt / 2. : dt, a_slot, h->rank);
          if (!rc) rc = fetch_flags(h);
        } else {
          int cur0 = h->cur, ccur0 = h->ccur;
          launch_bucket_substep(h, 0., dt, k == nleap - 1 ? dt / 2. : dt, 0., a_slot, h->rank);
          rc = fetch_flags(h);
          if (!rc && h->h_flags[0] != 0xffffffffu) {
            h->cur = cur0; h->ccur = ccur0;
            rc = reset_flags(h);
            if (!rc) rc = rebucket(h, 0.);
            if (!rc) rc = WENDY_RETRY;
          }
        }
      } while (rc == WENDY_RETRY && ++tries < 4);
      h->ext_half_done = false;
      if (!rc) *t0 += dt;  // wendy/wendy.c:403-404,409-410
    }
    dev_free(a_id); dev_free(a_slot);
  }
  if (!rc) rc = wendy_cuda_read(h, x, v);
  if (!rc) {
    cudaError_t e = cudaMemcpy(rank.data(), h->rank, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = set_err(WENDY_E_CUDA, cudaGetErrorString(e));
  }
  if (!rc) {
    // xi: order of the last force evaluation, values after the final half drift
    for (int i = 0; i < N; i++) {
      xi[rank[i]].idx = i;
      xi[rank[i]].val = x[i];
    }
  } else if (err) {
    *err = -1;
  }
  wendy_cuda_destroy(h);
  if (time_elapsed)
    *time_elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - tb).count();
}

}  // extern "C"
