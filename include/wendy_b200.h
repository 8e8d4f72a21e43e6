/*
 * wendy_b200.h -- C ABI of libwendy_b200.so, the B200-native (sm_100a) replacement for the
 * approximate-integration hot path of jobovy/wendy.
 *
 * The boundary this library replaces is the reference's single FFI call
 *     void _wendy_nbody_approx_onestep(...)          reference wendy/wendy.h:29-34,
 *                                                    wendy/wendy.c:385-418,
 *                                                    bound by ctypes at wendy/wendy.py:76-93
 * which the Python generator _nbody_approx calls once per output step
 * (wendy/wendy.py:424-433).  Two entry styles are exported:
 *
 *  (1) COMPAT:   _wendy_nbody_approx_onestep with the reference's exact 16-argument
 *                signature on HOST pointers (H2D, nleap sub-steps on the GPU, D2H), so the
 *                reference's own ctypes binding can load this library unchanged
 *                (INTEGRATION.md shows the two-line change).
 *  (2) RESIDENT: a handle API that keeps (x, v, m, id) in HBM across calls; this is what
 *                wendy_b200.nbody() uses.
 *
 * Plain C types only; no CUDA or torch types appear in any signature (streams and device
 * pointers cross as void* / double*).  All functions return 0 on success and a negative
 * WENDY_E_* code on failure; wendy_cuda_last_error() gives the message.  A handle is not
 * thread-safe; distinct handles are independent (the reference C side is stateless too,
 * SURVEY.md section 8b).
 */
#ifndef WENDY_B200_H
#define WENDY_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wendy_cuda_handle wendy_cuda_handle;

#define WENDY_OK 0
#define WENDY_E_CUDA (-1)       /* a CUDA runtime call failed */
#define WENDY_E_ARG (-2)        /* invalid argument */
#define WENDY_E_OVERFLOW (-3)   /* a bucket cannot hold its particles even after re-balancing
                                   (more than ~cap/4 exactly coincident particles) */
#define WENDY_RETRY 1           /* wendy_cuda_substep only: layout was re-balanced, storage
                                   slots moved; re-evaluate a_ext and call again */

/* sort modes (flags & 0xf) */
#define WENDY_SORT_AUTO 0    /* bucket fast path, radix sort to (re)build the layout */
#define WENDY_SORT_RADIX 1   /* full LSD radix sort every sub-step (A/B and fallback) */
/* Cumulative mass (reference wendy/wendy.c:359-360, a SERIAL fp64 running sum over the sorted order):
 *  - all masses bit-identical (the usual case): the library evaluates the reference's own serial sum in
 *    closed form (it depends only on the sorted position), so x and v are bit-identical to the reference
 *    C path at any N.  WENDY_FLAG_EXACT_SCAN selects the correctly rounded exact sum RN(rank * m0) instead
 *    (what the general path below computes; differs from the reference by the reference's own accumulated
 *    rounding, ~N * 2^-54 relative).
 *  - general masses: correctly rounded exact prefix sum (128-bit fixed point); WENDY_FLAG_GENERAL_MASSES
 *    forces this path (and its mass arrays) even when the masses are equal. */
#define WENDY_FLAG_GENERAL_MASSES 0x10
#define WENDY_FLAG_EXACT_SCAN 0x20

/* Mirror of the reference record, wendy/wendy.h:12-16 (int idx; 4 bytes pad; double val). */
struct wendy_array_w_index {
  int idx;
  double val;
};

/* ---- (2) resident API ------------------------------------------------------------------ */

/* Create device state from HOST arrays in particle-index order.
 *   N          total particles (all segments), N < 2^31 (ids are int, wendy/wendy.h:14)
 *   m          masses ALREADY multiplied by twopiG (wendy/wendy.py:371)
 *   totmass    n_segments values; the reference passes numpy.sum(m) (wendy/wendy.py:383)
 *   omega2     < 0: no harmonic term (wendy/wendy.py:363-366, wendy/wendy.c:375)
 *   n_segments independent realisations of N/n_segments particles each, laid out
 *              contiguously (1 for a single system)
 *   flags      WENDY_SORT_*
 *   cap, fill  bucket capacity (0: default 256 = one warp per bucket; 2048 = one CTA per bucket)
 *              and target fill (0: 128 of 256, 1536 of 2048)
 *   cuda_stream cudaStream_t to run on (NULL: the legacy default stream)                   */
int wendy_cuda_create(wendy_cuda_handle **h, long long N, const double *x, const double *v,
                      const double *m, const double *totmass, double omega2, int n_segments,
                      int flags, int cap, int fill, void *cuda_stream);

/* Replace the total mass per segment given at creation (n_segments values).  Lets the caller compute
 * numpy.sum(m) (wendy/wendy.py:383) concurrently with wendy_cuda_create; call before the first step. */
int wendy_cuda_set_totmass(wendy_cuda_handle *h, const double *totmass);

/* Same, from DEVICE arrays (no host staging).  m_dev may be NULL: all particles have mass m0. */
int wendy_cuda_create_dev(wendy_cuda_handle **h, long long N, const double *x_dev, const double *v_dev,
                          const double *m_dev, double m0, const double *totmass, double omega2,
                          int n_segments, int flags, int cap, int fill, void *cuda_stream);

/* One call of the reference entry point: drift dt/2, nleap x [force, kick, drift], with the
 * last drift dt/2 (wendy/wendy.c:398-411).  Synchronous; *time_elapsed (may be NULL) gets the
 * wall seconds of the call (wendy/wendy.c:396,416-417).  No external force. */
int wendy_cuda_step(wendy_cuda_handle *h, double dt_leap, int nleap, double *time_elapsed);

/* Asynchronous form of wendy_cuda_step: _begin enqueues the sub-steps and returns, _end waits,
 * checks for bucket overflow and (rarely) re-runs.  wendy_cuda_read_begin / _end is the
 * overlapped read-out: the de-sort runs on the compute stream, the D2H copies on a private copy
 * stream, so   step_end(k); read_begin; step_begin(k+1); read_end   hides the copy behind the
 * next call's kernels.  Host buffers may be page-locked (wendy_cuda_pin: direct copies) or ordinary
 * pageable memory (filled through the library's page-locked bounce buffers by a worker thread that
 * wendy_cuda_read_end joins); the same holds for wendy_cuda_read. */
int wendy_cuda_step_begin(wendy_cuda_handle *h, double dt_leap, int nleap);
int wendy_cuda_step_end(wendy_cuda_handle *h);
int wendy_cuda_last_call_seconds(wendy_cuda_handle *h, double *seconds);  /* device time of the last begin/end call */
int wendy_cuda_stage_ahead(wendy_cuda_handle *h);  /* after wendy_cuda_step_begin: de-sort the state that call will leave
                                   into a spare staging set, so that the next read_begin starts its copy at once */
int wendy_cuda_read_begin(wendy_cuda_handle *h, double *x_host, double *v_host);
int wendy_cuda_read_end(wendy_cuda_handle *h);

/* External-force stepping, one sub-step at a time (the caller evaluates F on device memory):
 *   wendy_cuda_force_positions: positions at the next force evaluation, as a DEVICE array of
 *       *n_slots doubles (storage order; slots not holding a particle contain finite junk).
 *       first_substep != 0 applies the leading half drift of a call (wendy/wendy.c:398).
 *   wendy_cuda_substep: force + kick(dt_kick) + drift(dt_drift); a_ext_dev is a DEVICE array
 *       in the same storage order (or NULL); h_next is the half drift the NEXT call will
 *       start with (dt_leap/2 after the last sub-step of a call, else 0).
 *       Returns WENDY_RETRY if the layout had to be re-balanced (see above). */
int wendy_cuda_force_positions(wendy_cuda_handle *h, double dt_leap, int first_substep,
                               double **x_dev, long long *n_slots);
int wendy_cuda_substep(wendy_cuda_handle *h, double dt_kick, double dt_drift, double h_next,
                       const double *a_ext_dev);
/* The same without a host round trip per sub-step: wendy_cuda_ext_begin; per sub-step wendy_cuda_force_positions,
 * the caller's F evaluation on the handle's stream, wendy_cuda_substep_async; then wendy_cuda_ext_end waits once.
 * *k_done = sub-steps completed; if it is less than the number enqueued, sub-step k_done overflowed a bucket: its
 * input has been restored and the layout rebuilt, and the caller runs the remaining sub-steps through
 * wendy_cuda_force_positions / wendy_cuda_substep.  The a_ext arrays must stay alive until wendy_cuda_ext_end. */
int wendy_cuda_ext_begin(wendy_cuda_handle *h);
int wendy_cuda_substep_async(wendy_cuda_handle *h, double dt_kick, double dt_drift, double h_next,
                             const double *a_ext_dev);
int wendy_cuda_ext_end(wendy_cuda_handle *h, int *k_done);

/* De-sort (wendy/wendy.c:413-415) and copy to HOST arrays of N doubles (either may be NULL). */
int wendy_cuda_read(wendy_cuda_handle *h, double *x_host, double *v_host);
/* Same, into DEVICE arrays of N doubles (no host copy). */
int wendy_cuda_read_dev(wendy_cuda_handle *h, double *x_dev, double *v_dev);

/* Diagnostics of the synchronised state, per the formulas of wendy/wendy.py:458-475,491 with
 * the stored (twopiG-scaled) masses: out[0] kinetic, out[1] harmonic, out[2] potential,
 * out[3] momentum, each summed over all segments.  E_reference = (out0+out1+out2)/twopiG. */
int wendy_cuda_energy(wendy_cuda_handle *h, double out[4]);

/* Closed form of the reference's serial cumulative-mass sum for N equal masses (wendy/wendy.c:359-360:
 * cumulmass[0] = 0, cumulmass[i+1] = cumulmass[i] + m0 in fp64): out[i] = cumulmass[k0 + i], i < n.  Host only
 * (no GPU): this is the table the equal-mass kernels evaluate.  Returns the number of linear pieces (> 0). */
int wendy_serial_cum(double m0, long long k0, long long n, double *out);

/* Counters: out[0] sub-steps, [1] layout rebuilds, [2] failed (re-run) sub-steps,
 * [3] max bucket count seen, [4] particles that left the 32-bucket window,
 * [5] kernels launched, [6] bucket capacity, [7] buckets, [8] sub-steps that fell back to the
 * radix path because a freshly balanced layout still overflowed. */
int wendy_cuda_stats(wendy_cuda_handle *h, long long *out, int n);

/* ---- sharded single system: one contiguous key range per GPU -------------------
 * The caller (wendy_b200/multi.py) owns the collectives; the library owns the local work.
 * Equal masses only (m0).  ids are GLOBAL particle indices.  bounds has nranks+1 ascending
 * edges (first -inf, last +inf); this GPU owns keys in [bounds[rank], bounds[rank+1]).
 *   create_shard   upload n_local particles whose positions lie in the range
 *   shard_substep  one leapfrog sub-step; particles whose new key leaves the range are written
 *                  to per-peer outboxes (out_counts[p] of them for peer p); pc_offset = number of
 *                  particles owned by lower ranks (offsets the cumulative mass)
 *   shard_outbox   DEVICE pointer of the outboxes: packed (x, v, id-as-double) records of three
 *                  doubles, peer p starts at record p * ocap
 *   shard_inject   append n received particles (DEVICE array of packed records) to the layout
 *   shard_read     compact local (x, v, id) to HOST arrays of capacity entries                  */
int wendy_cuda_create_shard(wendy_cuda_handle **h, long long n_local, long long n_capacity,
                            const double *x, const double *v, const int *ids, double m0,
                            double totmass, double omega2, int nranks, int rank,
                            const double *bounds, long long outbox_capacity, void *cuda_stream);
/* create_shard from DEVICE arrays (x_dev, v_dev, ids_dev), for a partition done on the GPU. */
int wendy_cuda_create_shard_dev(wendy_cuda_handle **h, long long n_local, long long n_capacity,
                                const double *x_dev, const double *v_dev, const int *ids_dev,
                                double m0, double totmass, double omega2, int nranks, int rank,
                                const double *bounds, long long outbox_capacity, void *cuda_stream);
/* Unequal masses (host-orchestrated exchange; the reference's force takes arbitrary m, wendy/wendy.c:375-383):
 * m[n_local] already times twopiG; sum_abs_m_global = sum |m| over ALL ranks (fixes the shared 128-bit fixed-point
 * scale, so that the ranks' exact mass totals add exactly and the result does not depend on the number of ranks).
 * Migrant records are then (x, v, id, m), four doubles.  Before every sub-step: wendy_cuda_shard_mass_total on
 * every rank (exact total of the masses it holds, two 64-bit words), an all-gather of those by the host language,
 * wendy_cuda_shard_set_mass_offset(sum over the lower ranks); then wendy_cuda_shard_substep as for equal masses. */
int wendy_cuda_create_shard_m(wendy_cuda_handle **h, long long n_local, long long n_capacity,
                              const double *x, const double *v, const double *m, const int *ids,
                              double sum_abs_m_global, double totmass, double omega2, int nranks, int rank,
                              const double *bounds, long long outbox_capacity, void *cuda_stream);
int wendy_cuda_shard_mass_total(wendy_cuda_handle *h, double h_pre, unsigned long long *total2);
int wendy_cuda_shard_set_mass_offset(wendy_cuda_handle *h, unsigned long long lo, unsigned long long hi);
int wendy_cuda_shard_read_masses(wendy_cuda_handle *h, double *m_host);  /* order of the last shard_read */
int wendy_cuda_shard_substep(wendy_cuda_handle *h, double h_pre, double dt_kick, double dt_drift,
                             double h_next, long long pc_offset, unsigned *out_counts);
int wendy_cuda_shard_outbox(wendy_cuda_handle *h, double **records, long long *ocap);
int wendy_cuda_shard_inject(wendy_cuda_handle *h, const double *records_dev, long long n);
int wendy_cuda_shard_count(wendy_cuda_handle *h, long long *n_local);
int wendy_cuda_shard_read(wendy_cuda_handle *h, double *x_host, double *v_host, int *id_host,
                          long long *n);
/* the same, overlapped: _begin compacts on the compute stream and starts the device -> host copies on a private
 * copy stream; the caller may step the shard before _end (which waits for the copies) */
int wendy_cuda_shard_read_begin(wendy_cuda_handle *h, double *x_host, double *v_host, int *id_host,
                                long long *n);
int wendy_cuda_shard_read_end(wendy_cuda_handle *h);

/* ---- sharded single system, device-driven exchange over peer memory (NVLink) ---------------------
 * With these entry points a sub-step needs NO host round trip and no host-side collective: the step kernel
 * stores migrants straight into the owner's inbox (peer memory), an inject kernel appends them, and the ranks
 * hand each other counts through flag words in each other's comm buffers (wendy_b200/csrc/peer.cuh).  One
 * process per GPU (CUDA IPC) or several ranks in one process (raw pointers).  Equal masses, <= 16 ranks.
 *   comm_export   allocate this rank's comm buffer; *ptr / *bytes describe it, ipc_handle64 receives the 64-byte
 *                 CUDA IPC handle (returns 1 instead of 0 when IPC is unavailable: raw pointers only)
 *   comm_open     map every peer's buffer: raw_ptrs[r] != 0 (same process) or ipc_handles + 64 r
 *   seed_counts   counts[r] = particles rank r owns; once after the partition and after every rollback
 *   step_begin    enqueue sub-steps [k0, nleap) of one call (wendy/wendy.c:398-411) on this rank's stream
 *   step_end      wait; *k_fail = first sub-step that did not complete here (nleap: none), *n_local = particles
 *                 owned now, *migrated_in = records received during the call.  The ranks agree on min(k_fail)
 *                 with ONE host collective per call; if it is < nleap every rank calls
 *   rollback      back to the input of sub-step k (then: all-gather the counts, seed_counts, step_begin(k0 = k)) */
int wendy_cuda_shard_comm_export(wendy_cuda_handle *h, unsigned long long *ptr, unsigned long long *bytes,
                                 unsigned char *ipc_handle64);
int wendy_cuda_shard_comm_open(wendy_cuda_handle *h, const unsigned char *ipc_handles,
                               const unsigned long long *raw_ptrs);
int wendy_cuda_shard_seed_counts(wendy_cuda_handle *h, const long long *counts);
int wendy_cuda_shard_prepare(wendy_cuda_handle *h, double dt_leap, int k0);  /* optional: the (local, synchronous)
                                   layout rebuild step_begin would do on demand, ahead of it */
int wendy_cuda_shard_step_begin(wendy_cuda_handle *h, double dt_leap, int nleap, int k0);
int wendy_cuda_shard_step_end(wendy_cuda_handle *h, int *k_fail, long long *n_local, long long *migrated_in);
int wendy_cuda_shard_rollback(wendy_cuda_handle *h, int k, long long *n_local);

/* Diagnostics on arbitrary particle arrays (each pointer may be HOST or DEVICE memory); m is the
 * unscaled mass, twopiG multiplies the sums as in the reference; omega2 < 0: no harmonic term.
 *   potential          out[j] = omega2 y_j^2/2 + twopiG sum_i m_i |x_i - y_j|      (wendy/wendy.py:494-517)
 *   energy_individual  out[i] = m_i (omega2 x_i^2/2 + v_i^2/2 + twopiG sum_k m_k |x_k - x_i|)
 *                                                                                  (wendy/wendy.py:466-470)
 * O((N + Y) log N): one radix sort + prefix sums + a binary search per point, instead of the
 * reference's O(N Y) broadcast. */
int wendy_cuda_potential(const double *y, long long Y, const double *x, const double *m, long long N,
                         double twopiG, double omega2, double *out, void *cuda_stream);
int wendy_cuda_energy_individual(const double *x, const double *v, const double *m, long long N,
                                 double twopiG, double omega2, double *out, void *cuda_stream);

/* Page-lock (cudaHostRegister) / release a HOST buffer the caller passes repeatedly to wendy_cuda_read, so
 * that the per-yield D2H copy is a direct DMA (37 instead of 42 ms per output at N=1e8; registering costs
 * about 0.2 s per GB, which is why wendy_b200.nbody does not do it by default).  The range is registered
 * piecewise (16 MB pieces ending on absolute address boundaries) so that other CUDA calls of the process are
 * not held up behind one long registration; wendy_cuda_unpin takes the pointer given to wendy_cuda_pin. */
int wendy_cuda_pin(void *host_ptr, unsigned long long bytes);
int wendy_cuda_unpin(void *host_ptr);

/* Device blocks of >= 32 MB released by wendy_cuda_destroy are cached for the next handle of the same size
 * (bounded by WENDY_B200_ALLOC_CACHE_GB, default 48, 0 = off; emptied automatically when an allocation
 * fails).  wendy_cuda_trim returns them to the driver.  No reference counterpart: the reference keeps its
 * state in numpy arrays (wendy/wendy.py:369-387). */
void wendy_cuda_trim(void);

/* Touch every page of a freshly allocated host array (contents kept) with a few threads, so that the first
 * read-out into it does not pay the page faults.  Host-only; no CUDA call. */
void wendy_host_set_threads(int n);  /* host threads the library's own copy / validation loops may use (0: default =
                                        OpenMP's maximum, at most 32).  One rank per GPU: cores / ranks on the node
                                        (wendy_b200/multi.py sets it); WENDY_B200_HOST_THREADS overrides */
void wendy_host_prefault(void *host_ptr, unsigned long long bytes);

/* Test/diagnostic hook: copies the per-bucket particle counts and lower splitters of the current
 * layout to HOST arrays of nb_max entries; returns the number of buckets (0: no layout yet). */
int wendy_cuda_debug_layout(wendy_cuda_handle *h, unsigned *counts, double *splitters, int nb_max);

void wendy_cuda_destroy(wendy_cuda_handle *h);
const char *wendy_cuda_last_error(void);

/* Stand-alone argsort of HOST keys by (value, index): the radix sort of this library, exposed
 * for parity tests against the reference's argsort (wendy/wendy.c:341-357). */
int wendy_cuda_argsort(const double *x_host, long long N, int *perm_out);

/* ---- (1) compat export: the reference's own symbol and signature ----------------------- */
void _wendy_nbody_approx_onestep(int N, struct wendy_array_w_index *xi, double *x, double *v,
                                 double *m, double *a, double totmass, double dt, int nleap,
                                 double *t0, double omega2,
                                 double (*ext_force)(int N, double *x, double t, double *a),
                                 int sort_type, int *err, double *time_elapsed,
                                 double *cumulmass);

#ifdef __cplusplus
}
#endif
#endif /* WENDY_B200_H */
