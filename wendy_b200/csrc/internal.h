// internal.h -- host-side declarations shared by the translation units of
// libwendy_b200.so.  Not part of the public C ABI (that is include/wendy_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>

#include "serialsum.cuh"

namespace wendy {

// ---- radix.cu -----------------------------------------------------------------------
struct RadixScratch {
  uint64_t *key[2] = {nullptr, nullptr};
  uint32_t *val[2] = {nullptr, nullptr};
  uint32_t *table = nullptr;  // [256][ntiles] digit-major histogram / offsets
  uint32_t *sums = nullptr;
  size_t n_alloc = 0;
};
size_t radix_table_entries(size_t n);
size_t radix_sums_entries(size_t n);
int radix_sort_pairs(cudaStream_t st, RadixScratch &s, size_t n, int seg_bits, unsigned seg_div);

// ---- tile.cu ------------------------------------------------------------------------
// One look-back descriptor per bucket: exact 128-bit mass sums and element counts.
struct __align__(16) Desc {
  unsigned long long agg_lo, agg_hi;  // this bucket's mass (fixed point)
  unsigned long long inc_lo, inc_hi;  // inclusive prefix over the segment
  long long agg_cnt, inc_cnt;
};

// ---- sharded single system, device-driven exchange over peer memory (NVLink) -------------------------------
// Every rank owns one COMM BUFFER (plain cudaMalloc, exported to the other ranks by CUDA IPC or, for ranks
// that share a process, by its raw pointer).  Peers write into it, the owner only reads it:
//   in_flag [2][PEER_MAX]   word (epoch, fail, count) from source r: "my step kernel of sub-step `epoch` is
//                           complete; `count` of my particles moved into your range, their (x, v, id)
//                           records are in inbox[epoch & 1][r][0 .. count)"
//   cnt_flag[2][PEER_MAX]   word (epoch, fail, n) from rank r: "after the inject of sub-step `epoch` I own n
//                           particles" -- the prefix over lower ranks offsets the cumulative mass
//   inbox   [2][nranks][ocap][3] doubles
// One 64-bit store carries flag and payload, so no ordering between two words is ever needed; records are made
// visible before the flag by __threadfence_system() in every writing CTA and in the CTA that signals.
constexpr int PEER_MAX = 16;
struct PeerComm {
  int nranks, my_rank;
  unsigned ocap;
  // local comm buffer
  const double *inbox;
  const unsigned long long *in_flag;
  const unsigned long long *cnt_flag;
  // the same three, in every peer's comm buffer
  double *peer_inbox[PEER_MAX];
  unsigned long long *peer_in_flag[PEER_MAX];
  unsigned long long *peer_cnt_flag[PEER_MAX];
  // local scratch
  unsigned *out_cnt;        // [nranks] records written to each peer by the step kernel in flight
  unsigned *cta_done;       // [2] finished-CTA counters: step kernel, inject kernel
  long long *n_local;       // particles owned after the last inject
  long long *n_hist;        // [k] particles owned at the start of sub-step k of the call in flight (k <= nleap)
  unsigned *peer_stat;      // [0] set when a wait for a peer timed out, [1] records received so far
  unsigned long long timeout_ns;  // a wait for a peer gives up after this long (WENDY_B200_PEER_TIMEOUT_MS, default 20 s)
};
__host__ __device__ __forceinline__ unsigned long long peer_pack(unsigned epoch, bool fail, unsigned v) {
  return ((unsigned long long)epoch << 32) | ((unsigned long long)(fail ? 1u : 0u) << 31) | (v & 0x7fffffffu);
}

enum { LOAD_BUCKET = 0, LOAD_GATHER = 1 };
enum { EMIT_SPLITTER = 0, EMIT_RANK = 1, EMIT_NONE = 2 };

struct TileParams {
  // input state (bucketed storage: bucket b owns slots [b*cap, b*cap + cnt_in[b]))
  const double *xin, *vin, *min;
  const int *idin;
  const double *aext;       // external acceleration per storage slot, or null
  const unsigned *cnt_in;   // LOAD_BUCKET
  const uint32_t *perm;     // LOAD_GATHER: storage slots in (segment, key) order
  // output state
  double *xout, *vout, *mout;
  int *idout;
  unsigned *cnt_out;        // EMIT_SPLITTER: must be zero on entry; EMIT_RANK: written
  unsigned *cnt_zero;       // array this launch clears for the launch after next (or null)
  const double *split;      // EMIT_SPLITTER: split[b] = lower edge of bucket b in the OUTPUT layout
  const double *split_in;   // lower edges of the INPUT layout (differs from split when splitters are advected)
  double *knot_sum;         // advection: per cell of knot_g buckets, sum of key displacements ... (or null)
  unsigned *knot_n;         // ... and number of particles
  int knot_g;
  // geometry
  int nb, nbps;             // buckets in total / per segment
  int dw;                   // persistent instances: destination window in buckets (<= tile_window_max(); 0 = smallest)
  long long seg_len;        // particles per segment
  // physics (reference wendy/wendy.c:375-383 and :324-333)
  double h_pre, dt_kick, dt_drift, h_next, omega2;
  const double *tot;        // total mass per segment
  int fxE;                  // fixed-point exponent
  int eqm;                  // all masses equal m0 (no mass arrays)
  double m0;
  const SerialTab *stab;    // equal masses: the reference's serial sum in closed form (serialsum.cuh);
                            // null: correctly rounded exact sum RN(rank * m0)
  // cross-CTA machinery
  unsigned *status;         // per bucket: (epoch << 2) | state   (mass look-back)
  Desc *desc;
  const unsigned *cpre;       // exclusive prefix of cnt_in over ALL buckets (count_prefix kernel)
  const ulonglong2 *mpre;     // general masses: exclusive 128-bit mass prefix over ALL buckets, or null
  unsigned epoch;
  unsigned *ticket, *ticket_zero;
  unsigned *fail_seq;       // smallest launch sequence number that failed
  unsigned seq;
  unsigned *stats;          // [0] max bucket count seen
  unsigned long long *outside;  // particles emitted outside the 32-bucket window (per call)
  // sharded single system (one key range per GPU): particles whose new key leaves
  // [bounds[my_rank], bounds[my_rank+1]) are appended to the outbox of the owning peer
  int nranks, my_rank;
  const double *bounds;     // nranks+1 ascending range edges (first -inf, last +inf), device
  double *out_rec;          // [nranks][ocap][orec] outboxes of packed (x, v, id[, m]) records
  int orec;                 // doubles per record: 3, or 4 with general masses
  unsigned long long pm_lo, pm_hi;  // general masses, sharded: exact 128-bit mass of the lower ranks
  unsigned *out_cnt;        // [nranks]
  unsigned ocap;
  long long pc_offset;      // particles owned by lower ranks (added to every rank)
  const PeerComm *peer;     // device-driven exchange (PERSIST = 3 instance): pc_offset comes from the peers' flags
  unsigned pepoch;          // ... of this sub-step (same number on every rank)
  int kcall;                // ... index of the sub-step within the call (n_hist slot)
  int nb_last;              // last bucket with a finite lower edge: together with bucket 0 the only ones whose
                            // range reaches beyond this GPU's key range
  // optional outputs
  int *rank_out;            // rank_out[id] = rank within the segment at this force evaluation
  double *energy_part;      // [nb][4]: kinetic, harmonic, potential, momentum partial sums
};

void launch_tile(cudaStream_t st, int cap, int load, int emit, int physics, const TileParams &p);
int tile_window_max();  // widest destination window of the persistent instances (TK_DWP)
// small.cu: resident kernel, one CTA per small system, all nleap sub-steps of a call in one launch
void launch_small(cudaStream_t st, double *x, double *v, const double *m, long long seg_len, int nseg,
                  const double *tot_seg, int eqm, double m0, const SerialTab *stab, double omega2, int fxE,
                  double dt, int nleap);
int small_max_particles();
// wstep.cu: warp-per-bucket sub-step (cap 256)
void launch_wstep(cudaStream_t st, int cap, const TileParams &p);
// exclusive prefix sum of the bucket counts (single pass, decoupled look-back over CTA tiles)
void launch_count_prefix(cudaStream_t st, const unsigned *cnt, int nb, unsigned *cpre,
                         unsigned long long *tile_desc, unsigned *ticket, unsigned epoch);
int count_prefix_tiles(int nb);
// Lagrangian splitters: move the bucket edges with the measured mean flow (monotone piecewise-linear map)
void launch_advect_splitters(cudaStream_t st, const double *split_old, double *split_new, int nb, int G,
                             double *knot_sum, unsigned *knot_n, double *knot_x, double *knot_y);
// general masses: exact bucket masses (one read of m) and their exclusive 128-bit prefix
void launch_mass_prefix(cudaStream_t st, const double *m, const unsigned *cnt, int cap, int nb, int fxE,
                        ulonglong2 *magg, ulonglong2 *mpre, Desc *desc, unsigned *status, unsigned *ticket,
                        unsigned epoch);
int mass_prefix_tiles(int nb);
bool wstep_cap_supported(int cap);
bool tile_cap_supported(int cap);
void tile_prepare_persistent();  // attributes + grid of the persistent instances, ahead of their first launch
int tile_coarse_cap();  // slots per bucket of the CTA kernel (2048)

struct ScatterParams {
  const double *xin, *vin, *min;
  const int *idin;
  const unsigned *cnt_in;   // null: source is dense (n_dense elements, segment = i / seg_len)
  const double *packed_in;  // non-null: dense source of packed (x, v, id[, m]) records (migrants)
  int prec;                 // doubles per packed record (3; 4 with general masses)
  long long n_dense;
  int cap_in, nb_in, nbps_in;
  double h;                 // bucket key = x + h*v
  double *xout, *vout, *mout;
  int *idout;
  unsigned *cnt_out;
  const double *split;
  int cap_out, nbps_out;
  long long seg_len;
  unsigned *fail_seq;
  unsigned seq;
};
void launch_scatter(cudaStream_t st, const ScatterParams &p, int sm_count);

// shard inject over peer memory: wait for every peer's migrants of sub-step `pepoch`, append them to the local
// buckets, publish the new local particle count to every peer
struct InjectParams {
  const PeerComm *peer;
  unsigned pepoch;
  int kcall;
  double h;                 // bucket key = x + h*v
  double *xout, *vout;
  int *idout;
  unsigned *cnt_out;
  const double *split;
  int cap, nb;
  unsigned *fail_seq;
  unsigned seq;
  unsigned *stats;
};
void launch_peer_inject(cudaStream_t st, const InjectParams &p, int grid);

// keys for the radix sort, written in compact (segment-major) order
void launch_make_keys(cudaStream_t st, const double *x, const double *v, double h,
                      const unsigned *cnt_in, const unsigned long long *offs, int cap, int nb,
                      long long n_dense, uint64_t *keys, uint32_t *vals, int val_mode,
                      long long seg_len, int nbps);
enum { VAL_SLOT = 0, VAL_SEGMENT = 1, VAL_INDEX = 2 };
void launch_scan_counts(cudaStream_t st, const unsigned *cnt, int nb, unsigned long long *offs);
void launch_pick_splitters(cudaStream_t st, const uint64_t *sorted_keys, long long seg_len,
                           int fill, int nbps, int nb, double *split);
void launch_apply_drift(cudaStream_t st, double *x, const double *v, double h,
                        const unsigned *cnt, int cap, int nb);
void launch_unsort(cudaStream_t st, const double *x, const double *v, const int *id,
                   const unsigned *cnt, int cap, int nb, double *xo, double *vo);
void launch_reduce_energy(cudaStream_t st, const double *part, int nb, double *out4);
// compact (x, v, id) of the live slots into dense arrays (order: bucket-major, arbitrary inside)
void launch_compact(cudaStream_t st, const double *x, const double *v, const int *id, const unsigned *cnt,
                    const unsigned *cpre, int cap, int nb, double *xo, double *vo, int *ido);

// ---- potential.cu ----------------------------------------------------------------------------------
int potential_eval(cudaStream_t st, const double *x, const double *v, const double *m, long long N,
                   const double *y, long long Y, double twopiG, double omega2, double *out, int individual,
                   std::string &err);

}  // namespace wendy
