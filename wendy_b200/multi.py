"""Multi-GPU modes of the approximate integrator (no reference
counterpart -- the reference is single-node OpenMP only).

Two modes, one process per GPU (``torch.distributed``, NCCL on GPUs, gloo in CPU tests):

* ``shard_ensemble``: independent realisations are dealt out to ranks; no data-path collective.
* ``ShardedSystem``: ONE large system, range-partitioned over ranks by position ("sample
  sort"): splitters from an all-gathered key sample, an initial all-to-all of particles, then
  per leapfrog sub-step
      local step on every rank (particles whose new key leaves the rank's range land in
      per-peer outboxes)  ->  all-to-all of the migrants (x, v, id)  ->  append on the receiver
      ->  all-gather of the per-rank particle counts, whose prefix offsets the cumulative mass.
  Equal masses (cumulative mass = the reference's serial sum at the GLOBAL rank, exactly what the
  single-GPU path computes, so results are independent of the number of ranks bit for bit).

  On GPUs of one node the per-sub-step exchange is DEVICE-DRIVEN (``CudaShardEngine.enable_peer``):
  the step kernel stores migrants straight into the owner's inbox over NVLink peer memory, a second
  kernel appends them, counts travel through flag words -- no host round trip and no collective per
  sub-step; the host runs one small collective per CALL to agree that no sub-step failed
  (wendy_b200/csrc/peer.cuh).  The host-orchestrated exchange below (all-gather + all-to-all per
  sub-step) remains as the portable path (gloo / numpy engine, several nodes).

The host logic here is engine-agnostic: the local work is done by an *engine* object
(``CudaShardEngine`` in the product; the tests plug in a numpy engine to run the same logic
under gloo on CPUs).
"""
import ctypes

import numpy

from . import _lib


# ---- ensembles --------------------------------------------------------------------------------
def shard_ensemble(n_realisations, rank, world_size):
    """Contiguous block of realisation indices owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(n_realisations, world_size)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


# ---- communication layer ----------------------------------------------------------------------
class TorchComm(object):
    """torch.distributed plumbing: small all-gathers and a variable-size all-to-all built from
    batched point-to-point ops (works on both NCCL and gloo)."""

    def __init__(self, group=None, device='cpu'):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group, self.device = torch, dist, group, device
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)

    def allgather_vec(self, vec):
        """vec: 1-D sequence of float64 -> (size, len) numpy array."""
        t = self.torch.as_tensor(numpy.asarray(vec, dtype=numpy.float64), device=self.device)
        out = [self.torch.empty_like(t) for _ in range(self.size)]
        self.dist.all_gather(out, t, group=self.group)
        return numpy.stack([o.cpu().numpy() for o in out])

    def exchange(self, send, counts=None):
        """send[p]: (n_p, w) float64 tensor for peer p (send[rank] is delivered locally); w = 3 (x, v, id) or
        4 (x, v, id, m).  counts[src][dst] may be passed when it is already known (saves one all-gather).
        Returns ONE contiguous (m, w) tensor with everything received from the other ranks."""
        torch, dist = self.torch, self.dist
        w = int(send[0].shape[1])
        if counts is None:
            counts = self.allgather_vec([s.shape[0] for s in send]).astype(numpy.int64)  # [src][dst]
        n_in = [int(counts[p][self.rank]) if p != self.rank else 0 for p in range(self.size)]
        n_out = [int(send[p].shape[0]) if p != self.rank else 0 for p in range(self.size)]
        total = sum(n_in)
        if total > (1 << 22):  # the one-off initial partition: do not keep gigabytes around
            inbox = torch.empty((total, w), dtype=torch.float64, device=self.device)
        else:
            if (getattr(self, '_inbox', None) is None or self._inbox.shape[0] < total
                    or self._inbox.shape[1] != w):
                self._inbox = torch.empty((max(total, 1024) * 2, w), dtype=torch.float64, device=self.device)
            inbox = self._inbox[:total]
        if total == 0 and sum(n_out) == 0:
            return inbox
        if dist.get_backend(self.group) == 'nccl':
            sendbuf = torch.cat([send[p] for p in range(self.size) if n_out[p]], dim=0) if sum(n_out) \
                else torch.empty((0, w), dtype=torch.float64, device=self.device)
            dist.all_to_all_single(inbox, sendbuf, output_split_sizes=n_in, input_split_sizes=n_out,
                                   group=self.group)
            return inbox
        ops, off = [], 0
        for p in range(self.size):
            if n_in[p]:
                ops.append(dist.P2POp(dist.irecv, inbox[off:off + n_in[p]], p, group=self.group))
                off += n_in[p]
            if n_out[p]:
                ops.append(dist.P2POp(dist.isend, send[p].contiguous(), p, group=self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        return inbox


# ---- local engine on a B200 ----------------------------------------------------------------------
class _DevView(object):
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': typestr,
                                         'data': (int(ptr), False), 'version': 2, 'strides': None}


class CudaShardEngine(object):
    """One key range on one GPU, over the shard entry points of libwendy_b200.so."""

    def __init__(self, x, v, ids, m0, totmass, omega2, nranks, rank, bounds, capacity,
                 outbox_capacity, device=None, m=None, sum_abs_m=0.):
        import torch
        self.torch = torch
        self._lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        self.nranks, self.rank = nranks, rank
        self._h = ctypes.c_void_p()
        bounds = numpy.ascontiguousarray(bounds, dtype=numpy.float64)
        self.capacity = int(capacity)
        # a stream of its own: kernels of the device-driven exchange wait for peers, which must never block
        # (or be blocked by) unrelated work -- in particular other ranks living in the same process
        self.stream = torch.cuda.Stream(device=self.device)
        st = ctypes.c_void_p(self.stream.cuda_stream)
        self.peer = False
        self.width = 3 if m is None else 4  # doubles per migrant record
        if m is not None:
            # unequal masses: host arrays, host-orchestrated exchange (wendy_cuda_create_shard_m)
            if torch.is_tensor(x):
                x, v, ids, m = (t.cpu().numpy() for t in (x, v, ids, m))
            _lib.check(self._lib.wendy_cuda_create_shard_m(
                ctypes.byref(self._h), len(x), self.capacity, numpy.ascontiguousarray(x, dtype=numpy.float64),
                numpy.ascontiguousarray(v, dtype=numpy.float64), numpy.ascontiguousarray(m, dtype=numpy.float64),
                numpy.ascontiguousarray(ids, dtype=numpy.int32), float(sum_abs_m), float(totmass), float(omega2),
                nranks, rank, bounds, int(outbox_capacity), st))
        elif torch.is_tensor(x):
            # the partition ran on this GPU: hand the device arrays over in place
            x = x.to(device=self.device, dtype=torch.float64).contiguous()
            v = v.to(device=self.device, dtype=torch.float64).contiguous()
            ids = ids.to(device=self.device, dtype=torch.int32).contiguous()
            torch.cuda.current_stream().synchronize()  # the arrays were produced on torch's stream
            _lib.check(self._lib.wendy_cuda_create_shard_dev(
                ctypes.byref(self._h), x.shape[0], self.capacity, x.data_ptr(), v.data_ptr(), ids.data_ptr(),
                float(m0), float(totmass), float(omega2), nranks, rank, bounds, int(outbox_capacity), st))
            self.stream.synchronize()  # ... and are copied on the engine's
        else:
            x = numpy.ascontiguousarray(x, dtype=numpy.float64)
            v = numpy.ascontiguousarray(v, dtype=numpy.float64)
            ids = numpy.ascontiguousarray(ids, dtype=numpy.int32)
            _lib.check(self._lib.wendy_cuda_create_shard(
                ctypes.byref(self._h), len(x), self.capacity, x, v, ids, float(m0), float(totmass),
                float(omega2), nranks, rank, bounds, int(outbox_capacity), st))
        self._pinned = None
        pr, oc = ctypes.c_void_p(), ctypes.c_longlong()
        _lib.check(self._lib.wendy_cuda_shard_outbox(self._h, ctypes.byref(pr), ctypes.byref(oc)))
        self._ocap = oc.value
        # zero-copy view of the packed (x, v, id[, m]) outbox records: [peer][slot][width]
        self._orec = torch.as_tensor(_DevView(pr.value, nranks * oc.value * self.width, '<f8'),
                                     device=self.device).view(nranks, oc.value, self.width)

    def close(self):
        if getattr(self, '_h', None):
            self._lib.wendy_cuda_destroy(self._h)
            self._h = None

    __del__ = close

    # -- device-driven exchange over peer memory ---------------------------------------------------------
    def enable_peer(self, comm):
        """Exchange comm-buffer descriptors through ``comm`` and map the peers.  Ranks in the same process hand
        over raw device pointers, other processes of the node CUDA IPC handles.  Returns False (and leaves the
        host-orchestrated path in place) if some rank cannot export its buffer."""
        import os
        ptr, nbytes = ctypes.c_ulonglong(), ctypes.c_ulonglong()
        handle = numpy.zeros(64, dtype=numpy.uint8)
        # (unequal masses: the export refuses, every rank sees it in the all-gather below, all stay host-orchestrated)
        rc = self._lib.wendy_cuda_shard_comm_export(self._h, ctypes.byref(ptr), ctypes.byref(nbytes), handle)
        ok = rc in (0, 1)
        vec = numpy.concatenate(([float(os.getpid()), float(ptr.value >> 32), float(ptr.value & 0xffffffff),
                                  1. if ok else 0., 1. if rc == 0 else 0.], handle.astype(numpy.float64)))
        allv = comm.allgather_vec(vec)
        if not numpy.all(allv[:, 3] == 1.):
            return False
        same = allv[:, 0] == float(os.getpid())
        if not numpy.all(same | (allv[:, 4] == 1.)):
            return False  # a peer in another process without IPC
        raw = numpy.zeros(self.nranks, dtype=numpy.uint64)
        for r in range(self.nranks):
            if same[r]:
                raw[r] = (int(allv[r, 1]) << 32) | int(allv[r, 2])
        handles = numpy.ascontiguousarray(allv[:, 5:69].astype(numpy.uint8).ravel())
        _lib.check(self._lib.wendy_cuda_shard_comm_open(self._h, handles, raw))
        self.peer = True
        return True

    def seed_counts(self, counts):
        _lib.check(self._lib.wendy_cuda_shard_seed_counts(
            self._h, numpy.ascontiguousarray(counts, dtype=numpy.int64)))

    def prepare(self, dt_leap, k0=0):
        _lib.check(self._lib.wendy_cuda_shard_prepare(self._h, float(dt_leap), int(k0)))

    def step_begin(self, dt_leap, nleap, k0=0):
        _lib.check(self._lib.wendy_cuda_shard_step_begin(self._h, float(dt_leap), int(nleap), int(k0)))

    def step_end(self):
        """(first sub-step that did not complete here, or nleap; particles owned; records received)."""
        kf, n, mig = ctypes.c_int(), ctypes.c_longlong(), ctypes.c_longlong()
        _lib.check(self._lib.wendy_cuda_shard_step_end(self._h, ctypes.byref(kf), ctypes.byref(n), ctypes.byref(mig)))
        return kf.value, n.value, mig.value

    def rollback(self, k):
        n = ctypes.c_longlong()
        _lib.check(self._lib.wendy_cuda_shard_rollback(self._h, int(k), ctypes.byref(n)))
        return n.value

    def stats(self):
        out = numpy.zeros(9, dtype=numpy.int64)
        _lib.check(self._lib.wendy_cuda_stats(self._h, out, 9))
        keys = ['substeps', 'rebuilds', 'failed_substeps', 'max_bucket_count', 'left_window',
                'kernel_launches', 'cap', 'buckets', 'radix_fallbacks']
        return dict(zip(keys, (int(o) for o in out)))

    def substep(self, h_pre, dt_kick, dt_drift, h_next, pc_offset):
        """Returns, per peer, an (n, 3) float64 CUDA tensor of migrants (x, v, id)."""
        torch = self.torch
        cnt = numpy.zeros(self.nranks, dtype=numpy.uint32)
        _lib.check(self._lib.wendy_cuda_shard_substep(self._h, h_pre, dt_kick, dt_drift, h_next,
                                                      int(pc_offset), cnt))
        return [self._orec[p, :int(cnt[p])] for p in range(self.nranks)]

    def inject(self, packed):
        """packed: contiguous (n, 3) float64 CUDA tensor of particles that now belong to this range."""
        n = packed.shape[0]
        if n == 0:
            return
        packed = packed.contiguous()
        self.torch.cuda.current_stream().synchronize()
        _lib.check(self._lib.wendy_cuda_shard_inject(self._h, packed.data_ptr(), n))

    def count(self):
        n = ctypes.c_longlong()
        _lib.check(self._lib.wendy_cuda_shard_count(self._h, ctypes.byref(n)))
        return n.value

    def mass_total(self, h_pre):
        """Exact 128-bit fixed-point total of the masses held now, as a Python int (unequal masses)."""
        t = numpy.zeros(2, dtype=numpy.uint64)
        _lib.check(self._lib.wendy_cuda_shard_mass_total(self._h, float(h_pre), t))
        return int(t[0]) | (int(t[1]) << 64)

    def set_mass_offset(self, total):
        total &= (1 << 128) - 1
        _lib.check(self._lib.wendy_cuda_shard_set_mass_offset(self._h, total & 0xffffffffffffffff, total >> 64))

    def read_masses(self):
        """Masses in the order of the last read()."""
        m = numpy.empty(self.capacity)
        _lib.check(self._lib.wendy_cuda_shard_read_masses(self._h, m))
        return m[:self.count()]

    def _host_buffers(self):
        """Ordinary numpy memory: the library fills pageable destinations through its page-locked bounce buffers
        at PCIe speed (page-locking capacity * 20 bytes per rank would cost about a second).  Two sets, the second
        made on first use: one is being filled (read_begin) while the caller still looks at the other.  Only the
        part a read-out will touch is pre-faulted."""
        if self._pinned is None:
            self._pinned, self._pin_sel = [None, None], 0
        if self._pinned[self._pin_sel] is None:
            bufs = (numpy.empty(self.capacity), numpy.empty(self.capacity),
                    numpy.empty(self.capacity, dtype=numpy.int32))
            touch = min(self.capacity, int(self.count() * 1.05) + 4096)
            for arr in bufs:
                self._lib.wendy_host_prefault(arr.ctypes.data, touch * arr.itemsize)
            self._pinned[self._pin_sel] = bufs
        return self._pinned[self._pin_sel]

    def read_begin(self):
        """Start the read-out of the local particles (compaction on the compute stream, device -> host copies on
        a copy stream); the shard may be stepped before ``read_end``."""
        x, v, ids = self._host_buffers()
        n = ctypes.c_longlong()
        _lib.check(self._lib.wendy_cuda_shard_read_begin(self._h, x, v, ids, ctypes.byref(n)))
        self._reading = (ids[:n.value], x[:n.value], v[:n.value])
        self._pin_sel ^= 1

    def read_end(self):
        """(ids, x, v) started by ``read_begin``: views of host buffers that are re-used two read-outs later."""
        _lib.check(self._lib.wendy_cuda_shard_read_end(self._h))
        out, self._reading = self._reading, None
        return out

    def read(self):
        """(ids, x, v) of the local particles: views of host buffers this engine re-uses (copy them to keep a
        snapshot across more than one further read)."""
        self.read_begin()
        out = self.read_end()
        self.stream.synchronize()
        return out

    def to_device(self, arr):
        return self.torch.as_tensor(numpy.ascontiguousarray(arr), device=self.device)


# ---- the sharded system ------------------------------------------------------------------------------
def choose_bounds(all_samples, nranks, n_total=0, disp=0., per_bucket=1664., slope=0.008):
    """Range edges from the pooled key sample, -inf / +inf at the ends; returns (bounds, share) with share[r] the
    fraction of the particles rank r is expected to own.

    Default: equal-count quantiles.  With ``n_total`` and ``disp`` (= sigma_v * dt_leap, the typical displacement per
    sub-step) the quantiles are weighted by the estimated COST of a particle instead: the step kernel's time per
    particle grows with the displacement measured in buckets, D_b = (number density) * disp / (particles per bucket)
    -- about 1 + 0.008 D_b, fitted to the single-GPU dt series and to the per-rank kernel times of the 8-GPU run
     -- so the dense central ranges of a large system get fewer particles and every rank's
    sub-step takes about the same time."""
    s = numpy.sort(numpy.asarray(all_samples, dtype=numpy.float64).ravel())
    s = s[numpy.isfinite(s)]
    b = numpy.empty(nranks + 1)
    b[0], b[-1] = -numpy.inf, numpy.inf
    share = numpy.full(nranks, 1. / nranks)
    if len(s) == 0:
        b[1:-1] = 0.
        return b, share
    w = numpy.ones(len(s))
    if n_total > 0 and disp > 0. and len(s) > 256:
        k = 32
        idx = numpy.arange(len(s))
        lo, hi = numpy.maximum(idx - k, 0), numpy.minimum(idx + k, len(s) - 1)
        span = numpy.maximum(s[hi] - s[lo], 1e-300)
        dens = float(n_total) / len(s) * (hi - lo) / span  # particles per unit length around every sample point
        w = 1. + slope * numpy.minimum(dens * disp / per_bucket, 1024.)
    cw = numpy.cumsum(w)
    cut = numpy.searchsorted(cw, cw[-1] * numpy.arange(1, nranks) / nranks)
    cut = numpy.minimum(cut, len(s) - 1)
    b[1:-1] = s[cut]
    b = numpy.maximum.accumulate(b)
    edges = numpy.concatenate(([0], cut, [len(s)])).astype(float)
    share = numpy.maximum(numpy.diff(edges), 0.) / len(s)
    return b, share


def route(keys, bounds):
    """Owner rank of every key: largest p with bounds[p] <= key."""
    return numpy.clip(numpy.searchsorted(bounds, keys, side='right') - 1, 0, len(bounds) - 2)


class ShardedSystem(object):
    """One self-gravitating system spread over ``comm.size`` ranks (see module docstring).

    Every rank passes the particles it happens to hold (any subset, with their GLOBAL ids);
    ``m0`` is the common particle mass ALREADY times twopiG, ``totmass`` the global total as the
    reference computes it (numpy.sum of the scaled masses, wendy/wendy.py:383).

    Unequal masses: pass ``m=`` (the masses of the particles given, times twopiG; ``m0`` is then ignored).  The
    cumulative mass becomes the correctly rounded exact prefix sum of the general single-GPU path -- every rank
    adds the exact 128-bit total of the lower ranks to its own scan, so the result is again independent of the
    number of ranks bit for bit.  The exchange is host-orchestrated in this mode (records carry the mass)."""

    def __init__(self, x, v, ids, m0, totmass, comm, omega=None, engine_factory=None,
                 capacity_factor=1.3, outbox_fraction=0.05, n_sample=65536, m=None, balance='cost'):
        self.comm = comm
        self.balance = balance  # 'cost' (default; only acts on systems of >= 2^24 particles) or 'count'
        self.m0, self.totmass = float(m0), float(totmass)
        self._m = None if m is None else numpy.ascontiguousarray(m, dtype=numpy.float64)
        self.general = m is not None
        self.sum_abs_m = 0.
        self.omega2 = -1. if omega is None else float(omega) ** 2.
        # (only read, at the first step: no copies -- 2 GB per rank at 1e8 particles)
        self._raw = (numpy.ascontiguousarray(x, dtype=numpy.float64), numpy.ascontiguousarray(v, dtype=numpy.float64),
                     numpy.ascontiguousarray(ids, dtype=numpy.int32))
        self.engine_factory = engine_factory or CudaShardEngine
        self.capacity_factor, self.outbox_fraction, self.n_sample = capacity_factor, outbox_fraction, n_sample
        self.engine = None
        self.bounds = None
        self.dt_leap = None
        self.pc_offset = 0
        self.migrated = 0
        self.peer, self.peer_tried = False, False
        self.timing = {'substep': 0., 'allgather': 0., 'exchange+inject': 0.}
        if self.engine_factory is CudaShardEngine:
            # one rank per GPU: the library's host-side copy threads must share the node's cores between the ranks
            import os
            local = int(os.environ.get('LOCAL_WORLD_SIZE', comm.size) or comm.size)
            _lib.load().wendy_host_set_threads(max(1, min(32, (os.cpu_count() or 1) // max(1, local) - 1)))

    # -- set-up: global sample sort on the keys of the FIRST force evaluation -------------------------
    def _partition(self, dt_leap):
        x, v, ids = self._raw
        comm = self.comm
        key = x + (dt_leap / 2.) * v  # position at the first force evaluation (wendy/wendy.c:398)
        n_tot = int(comm.allgather_vec([len(x)]).sum())
        take = numpy.linspace(0, max(len(key) - 1, 0), num=min(self.n_sample, len(key))).astype(int)
        sample = numpy.full(self.n_sample, numpy.nan)
        sample[:len(take)] = key[take] if len(key) else []  # a strided subset is an unbiased key sample
        disp = 0.
        if self.balance == 'cost' and n_tot >= (1 << 24):
            # typical displacement per sub-step: global velocity dispersion * dt_leap
            mom = comm.allgather_vec([float(len(v)), float(numpy.sum(v)), float(numpy.sum(v * v))]).sum(axis=0)
            disp = float(numpy.sqrt(max(mom[2] / mom[0] - (mom[1] / mom[0]) ** 2., 0.))) * abs(dt_leap)
        self.bounds, share = choose_bounds(comm.allgather_vec(sample), comm.size, n_tot, disp)
        # (head-room over the share of the particles this rank is expected to own, not over the mean)
        cap = int(self.capacity_factor * n_tot * max(float(share[comm.rank]), 0.5 / comm.size)) + 1024
        # (the same on every rank: a sender addresses the receiver's inbox with its own outbox capacity)
        self._ocap = max(1024, int(self.outbox_fraction * self.capacity_factor * n_tot / comm.size))
        if str(comm.device).startswith('cuda') and self.engine_factory is CudaShardEngine and not self.general:
            return self._partition_device(dt_leap, cap)
        owner = route(key, self.bounds)
        # the engine's tensors decide where the exchange buffers live (cuda for NCCL, cpu for gloo)
        import torch
        cols = [x, v, ids.astype(numpy.float64)] + ([self._m] if self.general else [])
        send = [torch.as_tensor(numpy.stack([c[owner == p] for c in cols], axis=1), device=comm.device)
                for p in range(comm.size)]
        keep = send[comm.rank]
        mine = torch.cat((keep, comm.exchange(send)), dim=0).cpu().numpy()
        extra = {}
        if self.general:
            # one fixed-point scale for all ranks: the global sum of |m| (same additions in the same order everywhere)
            self.sum_abs_m = float(numpy.sum(comm.allgather_vec([float(numpy.sum(numpy.abs(self._m)))])[:, 0]))
            extra = {'m': mine[:, 3].copy(), 'sum_abs_m': self.sum_abs_m}
        self.engine = self.engine_factory(mine[:, 0].copy(), mine[:, 1].copy(), mine[:, 2].astype(numpy.int32),
                                          self.m0, self.totmass, self.omega2, comm.size, comm.rank,
                                          self.bounds, cap, self._ocap, **extra)
        self._raw = None
        self._m = None
        self.dt_leap = dt_leap
        self._update_offset()

    def _partition_device(self, dt_leap, cap):
        """The same partition with the bulk work on the GPU: one H2D of the raw arrays, owner by
        binary search in the bounds, grouping by owner with one sort, NCCL all-to-all, and the engine
        takes the device arrays as they are."""
        import torch
        x, v, ids = self._raw
        comm = self.comm
        dev = comm.device
        xd, vd = torch.as_tensor(x, device=dev), torch.as_tensor(v, device=dev)
        key = xd + (dt_leap / 2.) * vd
        inner = torch.as_tensor(self.bounds[1:-1], device=dev)
        owner = torch.bucketize(key, inner, right=True)  # == route(): largest p with bounds[p] <= key
        del key
        n_to = torch.bincount(owner, minlength=comm.size).cpu().numpy()
        order = torch.argsort(owner)
        del owner
        packed = torch.stack((xd, vd, torch.as_tensor(ids, device=dev).to(torch.float64)), dim=1)[order]
        del order, xd, vd
        send = list(torch.split(packed, [int(c) for c in n_to], dim=0))
        mine = torch.cat((send[comm.rank], comm.exchange(send)), dim=0)
        del send, packed
        self.engine = CudaShardEngine(mine[:, 0], mine[:, 1], mine[:, 2].to(torch.int32),
                                      self.m0, self.totmass, self.omega2, comm.size, comm.rank,
                                      self.bounds, cap, self._ocap)
        del mine
        self._raw = None
        self.dt_leap = dt_leap
        self._update_offset()

    def _update_offset(self):
        counts = self.comm.allgather_vec([self.engine.count()])[:, 0]
        self.counts = counts.astype(numpy.int64)
        self.pc_offset = int(self.counts[:self.comm.rank].sum())
        # device-driven exchange where the engine offers it (CUDA engine, all ranks reachable by peer memory)
        import os
        mode = os.environ.get('WENDY_B200_SHARD_PEER', '1')  # 'force': also with one rank (kernel diagnostics)
        if (not self.peer_tried and hasattr(self.engine, 'enable_peer') and mode != '0'
                and (self.comm.size > 1 or mode == 'force')):
            self.peer_tried = True
            self.peer = bool(self.engine.enable_peer(self.comm))
        if self.peer:
            self.engine.seed_counts(self.counts)

    def _repartition(self, dt_leap):
        """A new time step changes the key of the first force evaluation (x + dt/2 v), hence which rank owns the
        particles near the range edges: partition again from the current (synchronised) state.  Rare and not
        optimised: the state goes through the host."""
        ids, x, v = self.engine.read()
        self._raw = (numpy.array(x), numpy.array(v), numpy.array(ids, dtype=numpy.int32))
        if self.general:
            self._m = numpy.array(self.engine.read_masses())
        self.engine.close()
        self.engine = None
        self.peer, self.peer_tried = False, False
        self._partition(dt_leap)

    def _step_peer(self, dt_leap, nleap):
        """One call with the device-driven exchange: enqueue everything, wait once, agree once."""
        import time
        tm = self.timing
        eng = self.engine
        k0, tries = 0, 0
        while True:
            t0 = time.perf_counter()
            if getattr(self.comm, 'in_process', False):
                # ranks that are threads of ONE process share a device: a device-synchronising CUDA call of one
                # rank (an allocation, a first-time kernel load inside a layout rebuild) would wait for kernels of
                # another rank that are waiting for this one.  Rebuild first, then start together.
                eng.prepare(dt_leap, k0)
                self.comm.barrier()
            eng.step_begin(dt_leap, nleap, k0)
            t1 = time.perf_counter()
            kf, n_local, mig = eng.step_end()
            t2 = time.perf_counter()
            info = self.comm.allgather_vec([kf, n_local]).astype(numpy.int64)  # the one collective of the call
            tm['substep'] += t1 - t0
            tm['exchange+inject'] += t2 - t1
            tm['allgather'] += time.perf_counter() - t2
            self.migrated += int(mig)
            kf_all = int(info[:, 0].min())
            if kf_all >= nleap:
                self.counts = info[:, 1]
                self.pc_offset = int(self.counts[:self.comm.rank].sum())
                return
            # some rank could not complete sub-step kf_all (bucket or inbox overflow): everyone goes back to its
            # input, rebuilds its layout and runs the rest of the call again
            tries += 1
            if tries > 8:
                raise RuntimeError('wendy_b200: sharded sub-step keeps overflowing after re-balancing (the density changes '
                                   'by more than the head-room of a fresh layout, or more particles cross a range '
                                   'edge than the inbox holds, within ONE sub-step: use a smaller dt or a larger '
                                   'outbox_fraction)')
            n_back = eng.rollback(kf_all)
            self.counts = self.comm.allgather_vec([n_back])[:, 0].astype(numpy.int64)
            eng.seed_counts(self.counts)
            k0 = kf_all

    # -- one reference call: drift dt/2, nleap x [force, kick, drift] (wendy/wendy.c:385-418) ------------
    def step(self, dt_leap, nleap):
        if self.engine is None:
            self._partition(dt_leap)
        elif dt_leap != self.dt_leap:
            self._repartition(dt_leap)
        import time
        tm = self.timing
        if self.peer:
            self._step_peer(dt_leap, nleap)
            return self
        for k in range(nleap):
            last = k == nleap - 1
            t0 = time.perf_counter()
            if self.general:
                # exact mass of the lower ranks: every rank's 128-bit total, all-gathered as four 32-bit pieces
                # (exact in float64), summed as Python integers
                tot = self.engine.mass_total(dt_leap / 2. if k == 0 else 0.)
                parts = self.comm.allgather_vec([float((tot >> (32 * j)) & 0xffffffff) for j in range(4)])
                below = sum(sum(int(parts[r, j]) << (32 * j) for j in range(4)) for r in range(self.comm.rank))
                self.engine.set_mass_offset(below)
            out = self.engine.substep(dt_leap / 2. if k == 0 else 0., dt_leap,
                                      dt_leap / 2. if last else dt_leap,
                                      dt_leap / 2. if last else 0., self.pc_offset)
            t1 = time.perf_counter()
            self.migrated += sum(int(o.shape[0]) for p, o in enumerate(out) if p != self.comm.rank)
            # ONE small all-gather per sub-step: every rank's outgoing counts and its particle count
            # after the export; incoming counts and the new count prefix follow from it on the host
            info = self.comm.allgather_vec([o.shape[0] for o in out] + [self.engine.count()]).astype(numpy.int64)
            sent = info[:, :-1]
            t2 = time.perf_counter()
            self.engine.inject(self.comm.exchange(out, counts=sent))
            tm['substep'] += t1 - t0
            tm['allgather'] += t2 - t1
            tm['exchange+inject'] += time.perf_counter() - t2
            self.counts = info[:, -1] + sent.sum(axis=0) - numpy.diag(sent)
            self.pc_offset = int(self.counts[:self.comm.rank].sum())
        return self

    def read_local(self):
        """(ids, x, v) of the particles this rank currently owns (synchronised state)."""
        return self.engine.read()

    def read_local_begin(self):
        """Overlapped read-out: start copying the particles this rank owns NOW to the host; ``step`` may be called
        before ``read_local_end`` returns them (the copy runs beside the next call's kernels)."""
        self.engine.read_begin()

    def read_local_end(self):
        return self.engine.read_end()

    def gather(self, n_total):
        """Full (x, v) in particle-index order on every rank (diagnostics / tests)."""
        ids, x, v = self.read_local()
        pad = int(self.comm.allgather_vec([len(ids)]).max())
        buf = numpy.full((3, pad), numpy.nan)
        buf[0, :len(ids)], buf[1, :len(ids)], buf[2, :len(ids)] = ids, x, v
        allb = self.comm.allgather_vec(buf.ravel()).reshape(self.comm.size, 3, pad)
        X, V = numpy.empty(n_total), numpy.empty(n_total)
        for b in allb:
            ok = numpy.isfinite(b[0])
            X[b[0][ok].astype(int)] = b[1][ok]
            V[b[0][ok].astype(int)] = b[2][ok]
        return X, V

    def close(self):
        if self.engine is not None:
            self.engine.close()
