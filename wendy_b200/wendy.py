"""Host-side mirror of the reference's Python API for the approximate integrator.

``nbody(x, v, m, dt, approx=True, nleap=...)`` has the signature and generator semantics of
reference wendy/wendy.py:112-206 (approx branch :151-157 -> ``_nbody_approx`` :336-437); the
per-output-step FFI call into wendy.c (wendy/wendy.py:424-433) is replaced by calls into
libwendy_b200.so with all particle state resident in HBM between ``next()`` calls.

Differences a user can observe:
  * ``approx=False`` (the exact event-driven solver, wendy/wendy.c:33-315) is out of scope.
  * ``ext_force(x, t)`` is called ONCE per leapfrog sub-step with ``x`` a 1-D CUDA
    ``torch.Tensor`` (float64) and must return a tensor of the same shape -- no per-particle
    Python/numba callback (reference wendy/wendy.py:389-420).  ``numpy.tanh`` becomes
    ``torch.tanh``.  The tensor covers the storage slots, a superset of the particles;
    F must be element-wise.
  * ``sort=`` accepts the reference's names (all map to the GPU default) plus
    ``'gpu'``/``'gpu-bucket'`` and ``'gpu-radix'``.
  * cumulative mass (reference wendy/wendy.c:359-360, a serial fp64 running sum): for equal
    masses the library evaluates that serial sum in closed form, so x and v are BIT-IDENTICAL
    to the reference C path at any N (tests: 1e4 ... 1e8 against the compiled reference);
    for unequal masses it is the correctly rounded exact prefix sum, which differs from the
    reference by the reference's own accumulated rounding (about N * 1e-16 relative in the
    cumulative mass; measured differences in x, v are in DESIGN.md section 4).
"""
import ctypes
import os

import numpy

from . import _lib

_REFERENCE_SORTS = ('quick', 'merge', 'tim', 'qsort', 'parallel')  # wendy/wendy.py:102


class _CudaArrayView(object):
    """Zero-copy view of a device buffer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': '<f8',
                                         'data': (int(ptr), False), 'version': 2,
                                         'strides': None}


class ApproxState(object):
    """Device-resident particle state + the step driver (one per generator)."""

    def __init__(self, x, v, m, omega2=-1., n_segments=1, sort='gpu', cap=0, fill=0, stream=None,
                 general_masses=False, exact_scan=False):
        self._lib = _lib.load()
        x = numpy.require(x, dtype=numpy.float64, requirements=['C'])
        v = numpy.require(v, dtype=numpy.float64, requirements=['C'])
        m = numpy.require(m, dtype=numpy.float64, requirements=['C'])
        if not (x.ndim == v.ndim == m.ndim == 1 and len(x) == len(v) == len(m)):
            raise ValueError('x, v, m must be 1-D arrays of the same length')
        self.N = len(x)
        self.n_segments = int(n_segments)
        if sort in _REFERENCE_SORTS:
            sort = 'gpu'
        if sort not in _lib.SORT_FLAGS:
            raise KeyError(sort)

        # totmass exactly as the reference computes it: numpy's pairwise sum
        # (wendy/wendy.py:383), one value per independent segment
        def total_mass():
            return numpy.ascontiguousarray(numpy.sum(m.reshape(self.n_segments, -1), axis=1))

        self._h = ctypes.c_void_p()
        self._sort, self._stream = sort, int(stream or 0)
        flags = _lib.SORT_FLAGS[sort] | (0x10 if general_masses else 0) | (0x20 if exact_scan else 0)
        st = ctypes.c_void_p(stream) if stream else None
        if self.N < (1 << 22) or self.N % self.n_segments:
            _lib.check(self._lib.wendy_cuda_create(ctypes.byref(self._h), self.N, x, v, m, total_mass(),
                                                   float(omega2), self.n_segments, flags, int(cap), int(fill), st))
        else:
            # the sum takes tens of milliseconds at N=1e8: a helper thread forms it while the library validates
            # and uploads (numpy and ctypes both release the GIL), then it replaces the placeholder
            import threading
            box = {}
            th = threading.Thread(target=lambda: box.__setitem__('tot', total_mass()))
            th.start()
            try:
                _lib.check(self._lib.wendy_cuda_create(ctypes.byref(self._h), self.N, x, v, m,
                                                       numpy.zeros(self.n_segments), float(omega2),
                                                       self.n_segments, flags, int(cap), int(fill), st))
            finally:
                th.join()
            _lib.check(self._lib.wendy_cuda_set_totmass(self._h, box['tot']))
        self.time_elapsed = 0.

    @classmethod
    def from_device(cls, x, v, m, omega2=-1., n_segments=1, sort='gpu', cap=0, fill=0, exact_scan=False):
        """Build the state from CUDA torch tensors without host staging.

        ``m`` is a float (equal masses, already times twopiG) or a CUDA tensor.  ``totmass`` per
        segment is ``m * seg_len`` for equal masses and a device sum otherwise (there is no
        host-side numpy.sum to mirror for device-generated initial conditions)."""
        import torch
        self = cls.__new__(cls)
        self._lib = _lib.load()
        x = x.detach().to(dtype=torch.float64).contiguous()
        v = v.detach().to(dtype=torch.float64).contiguous()
        self.N = x.numel()
        self.n_segments = int(n_segments)
        if torch.is_tensor(m):
            m = m.detach().to(dtype=torch.float64).contiguous()
            tot = m.reshape(self.n_segments, -1).sum(dim=1).cpu().numpy()
            mptr, m0 = ctypes.c_void_p(m.data_ptr()), 0.
        else:
            tot = numpy.full(self.n_segments, float(m) * (self.N // self.n_segments))
            mptr, m0 = None, float(m)
        if sort in _REFERENCE_SORTS:
            sort = 'gpu'
        self._h = ctypes.c_void_p()
        self._sort, self._stream = sort, int(torch.cuda.current_stream().cuda_stream)
        torch.cuda.current_stream().synchronize()
        _lib.check(self._lib.wendy_cuda_create_dev(
            ctypes.byref(self._h), self.N, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(v.data_ptr()), mptr, m0,
            numpy.ascontiguousarray(tot, dtype=numpy.float64), float(omega2), self.n_segments,
            _lib.SORT_FLAGS[sort] | (0x20 if exact_scan else 0), int(cap), int(fill),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.time_elapsed = 0.
        return self

    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            self._lib.wendy_cuda_destroy(self._h)
            self._h = None

    __del__ = close

    def step(self, dt_leap, nleap):
        """One reference call (wendy/wendy.c:385-418) without external force."""
        t = ctypes.c_double(0.)
        _lib.check(self._lib.wendy_cuda_step(self._h, dt_leap, nleap, ctypes.byref(t)))
        self.time_elapsed = t.value
        return self

    def step_begin(self, dt_leap, nleap):
        """Enqueue one reference call without waiting for it (pair with step_end)."""
        import time
        self._t_begin = time.perf_counter()
        _lib.check(self._lib.wendy_cuda_step_begin(self._h, dt_leap, nleap))

    def step_end(self):
        _lib.check(self._lib.wendy_cuda_step_end(self._h))
        # the reference reports the time of the integration only (wendy/wendy.c: time_begin / time_end around it):
        # device time between the first and the last launch of the call, not the wall time since step_begin
        # (which would include the overlapped read-out and whatever the consumer did between two next() calls)
        t = ctypes.c_double(0.)
        _lib.check(self._lib.wendy_cuda_last_call_seconds(self._h, ctypes.byref(t)))
        self.time_elapsed = t.value

    def stage_ahead(self):
        """After step_begin: de-sort the state that call will leave ahead of its read-out (wendy_cuda_stage_ahead)."""
        _lib.check(self._lib.wendy_cuda_stage_ahead(self._h))

    def read_begin(self, x_out, v_out):
        _lib.check(self._lib.wendy_cuda_read_begin(self._h, x_out.ctypes.data, v_out.ctypes.data))

    def read_end(self):
        _lib.check(self._lib.wendy_cuda_read_end(self._h))

    def step_ext(self, dt_leap, nleap, ext_force, t0):
        """Same, with a torch-vectorised external force; returns the advanced t0.

        Time convention of the reference: force k of a call sees t0 + k*dt_leap
        (wendy/wendy.c:402-404,408-410)."""
        import time
        import torch
        tb = time.perf_counter()
        xptr, ns = ctypes.c_void_p(), ctypes.c_longlong()

        # F runs on torch's current stream, the library on the handle's: the same stream unless the caller entered a
        # torch.cuda.stream() context (then every hand-over is fenced by a device synchronisation, and the
        # asynchronous path is not used)
        same_stream = int(torch.cuda.current_stream().cuda_stream) == self._stream

        def force(k, t):
            _lib.check(self._lib.wendy_cuda_force_positions(self._h, dt_leap, int(k == 0),
                                                            ctypes.byref(xptr), ctypes.byref(ns)))
            if not same_stream:
                torch.cuda.synchronize()
            xs = torch.as_tensor(_CudaArrayView(xptr.value, ns.value), device='cuda')
            a = ext_force(xs, t)
            if not torch.is_tensor(a):
                a = torch.as_tensor(a, dtype=torch.float64, device=xs.device)
            a = a.to(dtype=torch.float64).expand_as(xs).contiguous()
            if not same_stream:
                torch.cuda.synchronize()
            return a

        k = 0
        if same_stream and self._sort != 'gpu-radix' and os.environ.get('WENDY_B200_EXT_ASYNC', '1') != '0':
            # all sub-steps of the call are enqueued without a host round trip each (the F evaluations are stream
            # work like the kernels); one wait at the end.  An overflowing sub-step voids the launches behind it:
            # the library restores its input and the loop below finishes the call synchronously.
            _lib.check(self._lib.wendy_cuda_ext_begin(self._h))
            keep = []
            for kk in range(nleap):
                a = force(kk, t0 + kk * dt_leap)
                keep.append(a)  # alive until the kernels that read it have run
                last = kk == nleap - 1
                _lib.check(self._lib.wendy_cuda_substep_async(
                    self._h, dt_leap, dt_leap / 2. if last else dt_leap, dt_leap / 2. if last else 0.,
                    ctypes.c_void_p(a.data_ptr())))
            done = ctypes.c_int(0)
            _lib.check(self._lib.wendy_cuda_ext_end(self._h, ctypes.byref(done)))
            del keep
            k = done.value
            t0 += k * dt_leap
        while k < nleap:
            a = force(k, t0)
            last = k == nleap - 1
            rc = _lib.check(self._lib.wendy_cuda_substep(
                self._h, dt_leap, dt_leap / 2. if last else dt_leap, dt_leap / 2. if last else 0.,
                ctypes.c_void_p(a.data_ptr())))
            if rc == _lib.WENDY_RETRY:
                continue  # layout re-balanced: slots moved, evaluate F again
            t0 += dt_leap
            k += 1
        self.time_elapsed = time.perf_counter() - tb
        return t0

    def read(self, x_out=None, v_out=None):
        """De-sort + D2H (wendy/wendy.c:413-415) into caller-owned ndarrays."""
        if x_out is None:
            x_out = numpy.empty(self.N)
        if v_out is None:
            v_out = numpy.empty(self.N)
        _lib.check(self._lib.wendy_cuda_read(self._h, x_out.ctypes.data, v_out.ctypes.data))
        return x_out, v_out

    def energy_terms(self):
        """[kinetic, harmonic, potential, momentum] with the stored (twopiG-scaled) masses."""
        out = numpy.zeros(4)
        _lib.check(self._lib.wendy_cuda_energy(self._h, out))
        return out

    def layout(self):
        """(counts, splitters) of the current bucket layout (diagnostic)."""
        nb = self.stats()['buckets']
        cnt = numpy.zeros(nb, dtype=numpy.uint32)
        spl = numpy.zeros(nb)
        nb = _lib.check(self._lib.wendy_cuda_debug_layout(self._h, cnt, spl, nb))
        return cnt[:nb], spl[:nb]

    def stats(self):
        out = numpy.zeros(9, dtype=numpy.int64)
        _lib.check(self._lib.wendy_cuda_stats(self._h, out, 9))
        keys = ['substeps', 'rebuilds', 'failed_substeps', 'max_bucket_count', 'left_window',
                'kernel_launches', 'cap', 'buckets', 'radix_fallbacks']
        return dict(zip(keys, (int(o) for o in out)))


def nbody(x, v, m, dt, t0=0., twopiG=1., omega=None, ext_force=None,
          approx=False, nleap=None, sort='gpu',
          maxcoll=100000, warn_maxcoll=False,
          full_output=False, n_segments=1, output='host', _cap=0, _fill=0, _general_masses=False,
          _exact_scan=False):
    """
    NAME:
       nbody
    PURPOSE:
       run an N-body simulation in 1D on a B200 (drop-in for reference wendy.nbody with approx=True)
    INPUT:
       x, v, m - positions, velocities, masses [N]
       dt - output time step
       t0= (0.) initial time (only matters for a time-dependent ext_force)
       twopiG= (1.) value of 2 pi G
       omega= (None) if set, frequency of an external harmonic oscillator Phi = omega^2 x^2/2
       ext_force= (None) F(x,t) on CUDA torch tensors (see module docstring)
       approx= must be True (the exact solver is out of scope)
       nleap= leapfrog sub-steps per dt
       sort= reference names accepted; 'gpu' (default), 'gpu-radix'
       full_output= (False) also yield the wall time of the step (reference: time_elapsed)
       n_segments= (1) treat the input as that many independent, equal-size realisations
       output= ('host') 'device': yield CUDA torch tensors (de-sorted on the GPU, no copy to the host) for
               consumers that stay on the device; as with the host arrays the SAME two tensors are yielded
               every time
    OUTPUT:
       Generator: each iteration returns (x,v) [+ time_elapsed]; as in the reference the SAME
       two ndarrays are yielded every time, updated in place
    """
    if not approx:
        raise NotImplementedError('wendy_b200 implements the approximate integrator only '
                                  '(approx=True); the exact event-driven solver is out of scope')
    if nleap is None:  # message pinned by reference tests/test_approx.py:187-196
        raise ValueError('When approx is True, the number of leapfrog steps nleap= per output time step needs to be set')
    for item in _nbody_approx(x, v, m, dt, nleap, t0=t0, sort=sort, omega=omega,
                              ext_force=ext_force, twopiG=twopiG, full_output=full_output,
                              n_segments=n_segments, _cap=_cap, _fill=_fill,
                              _general_masses=_general_masses, _exact_scan=_exact_scan, output=output):
        yield item


def _nbody_approx(x, v, m, dt, nleap, t0=0., omega=None, ext_force=None, sort='gpu',
                  twopiG=1., full_output=False, n_segments=1, _cap=0, _fill=0, _general_masses=False,
                  _exact_scan=False, output='host'):
    """Setup follows reference wendy/wendy.py:363-387,422; loop follows :424-437."""
    omega2 = -1. if omega is None else omega ** 2.
    # The inputs are only read (the reference copies them, wendy/wendy.py:369-370; here the "copy" is the
    # upload); the arrays that are yielded -- the same two every time, as in the reference -- are separate
    # page-locked buffers, so that the per-output D2H copy runs at full PCIe speed.
    xin = numpy.ascontiguousarray(x, dtype=numpy.float64)
    vin = numpy.ascontiguousarray(v, dtype=numpy.float64)
    # masses are only read by the library: scale (reference wendy/wendy.py:371) but do not copy needlessly
    ms = numpy.ascontiguousarray(m, dtype=numpy.float64)
    if twopiG != 1.:
        ms = twopiG * ms
    n = xin.shape[0]
    out = {}
    lib = _lib.load()

    out['pinned'] = []
    # The yielded arrays are ordinary numpy memory: the library fills pageable destinations through its own
    # page-locked bounce buffers at PCIe speed (all host threads copy piece i while the copy engine delivers
    # piece i+1), which avoids page-locking 16 bytes/particle per generator -- 350-450 ms at N=1e8, more than
    # the rest of the set-up together.  WENDY_B200_D2H=pinned restores the page-locked yield buffers.
    d2h_mode = os.environ.get('WENDY_B200_D2H', 'bounce')
    pin_outputs = d2h_mode.startswith('p')
    # 'hybrid' (opt-in, not measured yet): start as above, page-lock the arrays in the background after the
    # second output; the library switches to direct copies (37 instead of 42 ms per output at N=1e8) by
    # itself once a whole array is registered
    pin_later = d2h_mode.startswith('h')

    def alloc_outputs():
        # (cudaHostRegister of untouched numpy memory: faster than cudaHostAlloc, and the arrays stay valid
        # ordinary memory after the generator has been closed and the registration dropped; done on a helper
        # thread while the main thread validates, uploads and builds the first layout)
        for name in ('x', 'v'):
            a = numpy.empty(n)
            out[name] = a
            if a.nbytes < (1 << 20):
                continue
            if pin_outputs:
                if lib.wendy_cuda_pin(a.ctypes.data, a.nbytes) == 0:
                    out['pinned'].append(a)
            else:
                lib.wendy_host_prefault(a.ctypes.data, a.nbytes)  # page faults now, not inside the first read-out

    if output == 'device':
        # The consumer stays on the GPU: de-sort into two CUDA tensors (wendy/wendy.c:413-415), nothing crosses
        # PCIe.  The library works on torch's current stream, so the tensors are ready for torch ops in stream order.
        import torch
        state = None
        try:
            cs = torch.cuda.current_stream().cuda_stream
            state = ApproxState(xin, vin, ms, omega2=omega2, n_segments=n_segments, sort=sort, cap=_cap, fill=_fill,
                                general_masses=_general_masses, exact_scan=_exact_scan, stream=cs)
            xt = torch.empty(n, dtype=torch.float64, device='cuda')
            vt = torch.empty(n, dtype=torch.float64, device='cuda')
            dt_leap = dt / nleap
            while True:
                if ext_force is None:
                    state.step(dt_leap, nleap)
                else:
                    t0 = state.step_ext(dt_leap, nleap, ext_force, t0)
                _lib.check(lib.wendy_cuda_read_dev(state._h, ctypes.c_void_p(xt.data_ptr()), ctypes.c_void_p(vt.data_ptr())))
                yield (xt, vt, state.time_elapsed) if full_output else (xt, vt)
        finally:
            if state is not None:
                state.close()
        return
    if output != 'host':
        raise ValueError("output must be 'host' or 'device'")

    # de-sort ahead of the read-out (a second staging set, 16 B per particle) where the read-out is long enough to
    # matter; WENDY_B200_STAGE_AHEAD=0 / 1 forces it off / on
    stage = os.environ.get('WENDY_B200_STAGE_AHEAD', '1' if n >= (1 << 20) else '0') != '0'
    import threading
    helpers = [threading.Thread(target=alloc_outputs)]
    for t in helpers:
        t.start()
    state = None
    dt_leap = dt / nleap
    try:
        try:
            state = ApproxState(xin, vin, ms, omega2=omega2, n_segments=n_segments, sort=sort, cap=_cap,
                                fill=_fill, general_masses=_general_masses, exact_scan=_exact_scan)
            if ext_force is None:
                state.step_begin(dt_leap, nleap)
                if stage:
                    state.stage_ahead()
        finally:
            for t in helpers:
                t.join()
        x, v = out['x'], out['v']
        del xin, vin
        if ext_force is None:
            # The generator is infinite (reference wendy/wendy.py:424), so the call after this
            # one is always needed: enqueue it before the D2H copy of this one has finished.
            n_out = 0
            while True:
                state.step_end()
                te = state.time_elapsed
                state.read_begin(x, v)
                state.step_begin(dt_leap, nleap)
                if stage:
                    # ... and its de-sort, behind its sub-steps: output k+1's copy will start the moment the call ends
                    state.stage_ahead()
                state.read_end()
                n_out += 1
                if pin_later and n_out == 2 and x.nbytes >= (1 << 20):
                    def pin_in_background():
                        for a in (x, v):
                            if lib.wendy_cuda_pin(a.ctypes.data, a.nbytes) == 0:
                                out['pinned'].append(a)
                    out['pin_thread'] = threading.Thread(target=pin_in_background, daemon=True)
                    out['pin_thread'].start()
                if full_output:
                    yield (x, v, te)
                else:
                    yield (x, v)
        # External force: the same speculation as above -- the de-sort + D2H of output k runs on the copy stream
        # (and the library's reader thread) while the sub-steps of call k+1, F evaluations included, are issued.
        t0 = state.step_ext(dt_leap, nleap, ext_force, t0)
        while True:
            te = state.time_elapsed
            state.read_begin(x, v)
            t0 = state.step_ext(dt_leap, nleap, ext_force, t0)
            state.read_end()
            if full_output:
                yield (x, v, te)
            else:
                yield (x, v)
    finally:
        if state is not None:
            state.close()
        if out.get('pin_thread') is not None:
            out['pin_thread'].join()
        for a in out.get('pinned', []):
            lib.wendy_cuda_unpin(a.ctypes.data)


def energy(x, v, m, twopiG=1., individual=False, omega=None, n_segments=1):
    """System energy by the formula of reference wendy/wendy.py:458-475 (individual=False),
    evaluated on the GPU: one radix sort + the exact mass scan + a fixed-order reduction.
    With n_segments > 1 the sum over all segments is returned.  individual=True gives the
    reference's per-particle energies (wendy/wendy.py:466-470) in O(N log N)."""
    if individual:
        return _diagnostic(None, x, v, m, twopiG, omega)
    x = numpy.asarray(x, dtype=numpy.float64)
    ms = twopiG * numpy.asarray(m, dtype=numpy.float64)
    st = ApproxState(x, v, ms, omega2=-1. if omega is None else omega ** 2., n_segments=n_segments)
    try:
        ke, he, pe, _ = st.energy_terms()
    finally:
        st.close()
    if twopiG == 0.:
        return numpy.sum(numpy.asarray(m) * numpy.asarray(v) ** 2. / 2.)
    return (he + pe + ke) / twopiG


def _as_f8(a):
    """(pointer, keep-alive object, is_cuda_tensor) of a float64 array: numpy-like or a torch CUDA tensor."""
    if hasattr(a, 'is_cuda') and a.is_cuda:
        import torch
        a = a.to(torch.float64).contiguous()
        return a.data_ptr(), a, True
    a = numpy.require(numpy.atleast_1d(a), dtype=numpy.float64, requirements=['C'])
    return a.ctypes.data, a, False


def _diagnostic(y, x, v, m, twopiG, omega):
    lib = _lib.load()
    om2 = -1. if omega is None else float(omega) ** 2.
    px, kx, cx = _as_f8(x)
    n = kx.shape[0]
    if not hasattr(m, 'is_cuda') and numpy.ndim(m) == 0:  # one common mass, as numpy broadcasting allows
        if cx:
            import torch
            m = torch.full((n,), float(m), dtype=torch.float64, device=kx.device)
        else:
            m = numpy.full(n, float(m))
    pm, km, _ = _as_f8(m)
    if km.shape[0] != n:
        raise ValueError('x and m must have the same length')
    stream = None
    if cx:
        import torch
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    if y is None:
        pv, kv, _ = _as_f8(v)
        if cx:
            out = torch.empty(n, dtype=torch.float64, device=kx.device)
            po = out.data_ptr()
        else:
            out = numpy.empty(n)
            po = out.ctypes.data
        _lib.check(lib.wendy_cuda_energy_individual(px, pv, pm, n, float(twopiG), om2, po, stream))
        return out
    py, ky, cy = _as_f8(y)
    ny = ky.shape[0]
    if cy:
        import torch
        out = torch.empty(ny, dtype=torch.float64, device=ky.device)
        po = out.data_ptr()
    else:
        out = numpy.empty(ny)
        po = out.ctypes.data
    _lib.check(lib.wendy_cuda_potential(py, ny, px, pm, n, float(twopiG), om2, po, stream))
    return out


def potential(y, x, v, m, twopiG=1., omega=None):
    """Gravitational (+ harmonic) potential at the points y -- reference wendy/wendy.py:494-517, same
    arguments (v is unused there too).  O((N + Y) log N) on the GPU instead of the reference's O(N Y)
    broadcast: radix sort, prefix sums of m and m x, one binary search per point.  Arrays may be
    numpy-like (result: numpy) or torch CUDA tensors (result: CUDA tensor, nothing crosses PCIe)."""
    return _diagnostic(y, x, v, m, twopiG, omega)


def trim():
    """Return the device blocks cached from closed generators to the CUDA driver (wendy_cuda_trim)."""
    _lib.load().wendy_cuda_trim()


def momentum(v, m):
    """reference wendy/wendy.py:477-491 (a single dot product; stays on the host)."""
    return numpy.sum(numpy.asarray(m) * numpy.asarray(v))


def argsort(x):
    """(value, index) argsort on the GPU radix sort -- parity hook for wendy/wendy.c:341-357."""
    x = numpy.require(x, dtype=numpy.float64, requirements=['C'])
    out = numpy.empty(len(x), dtype=numpy.int32)
    _lib.check(_lib.load().wendy_cuda_argsort(x, len(x), out))
    return out
