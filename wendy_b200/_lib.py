"""ctypes binding of libwendy_b200.so (C ABI: include/wendy_b200.h).

Plays the role of the reference's loader + argtypes block, wendy/wendy.py:13-101.  There is
deliberately NO fallback: if the CUDA library is missing or cannot be loaded, importing the
product path raises (the reference also hard-fails without wendy_c, wendy/wendy.py:40).
"""
import ctypes
import os

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
# WENDY_B200_LIB selects an alternative build of the same library (kernel A/B experiments)
LIB_PATH = os.environ.get('WENDY_B200_LIB') or os.path.join(_HERE, 'libwendy_b200.so')
c_double_p = ctypes.POINTER(ctypes.c_double)
c_ll_p = ctypes.POINTER(ctypes.c_longlong)

WENDY_RETRY = 1
SORT_FLAGS = {'gpu': 0, 'gpu-bucket': 0, 'gpu-radix': 1}
#: numpy mirror of struct wendy_array_w_index / reference wendy/wendy.h:12-16
XI_DTYPE = numpy.dtype([('idx', 'i4'), ('val', 'f8')], align=True)

_lib = None


def _nd(dtype):
    return numpy.ctypeslib.ndpointer(dtype=dtype, flags=('C_CONTIGUOUS',))


def load():
    """Load the shared library (once) and declare every prototype of include/wendy_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("wendy_b200: CUDA library %s not found -- build it with "
                          "`python -c 'import __graft_entry__ as g; g.build()'` or "
                          "`make -C wendy_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    lib.wendy_cuda_last_error.restype = ctypes.c_char_p
    lib.wendy_cuda_last_error.argtypes = []
    lib.wendy_cuda_create.restype = ctypes.c_int
    lib.wendy_cuda_create.argtypes = [ctypes.POINTER(vp), ctypes.c_longlong, _nd('f8'), _nd('f8'),
                                      _nd('f8'), _nd('f8'), ctypes.c_double, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
    lib.wendy_cuda_set_totmass.restype = ctypes.c_int
    lib.wendy_cuda_set_totmass.argtypes = [vp, _nd('f8')]
    lib.wendy_cuda_create_dev.restype = ctypes.c_int
    lib.wendy_cuda_create_dev.argtypes = [ctypes.POINTER(vp), ctypes.c_longlong, vp, vp, vp, ctypes.c_double, _nd('f8'),
                                          ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
    lib.wendy_cuda_step.restype = ctypes.c_int
    lib.wendy_cuda_step.argtypes = [vp, ctypes.c_double, ctypes.c_int, c_double_p]
    lib.wendy_cuda_step_begin.restype = ctypes.c_int
    lib.wendy_cuda_step_begin.argtypes = [vp, ctypes.c_double, ctypes.c_int]
    lib.wendy_cuda_stage_ahead.restype = ctypes.c_int
    lib.wendy_cuda_stage_ahead.argtypes = [vp]
    lib.wendy_cuda_last_call_seconds.restype = ctypes.c_int
    lib.wendy_cuda_last_call_seconds.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    lib.wendy_cuda_step_end.restype = ctypes.c_int
    lib.wendy_cuda_step_end.argtypes = [vp]
    lib.wendy_cuda_read_begin.restype = ctypes.c_int
    lib.wendy_cuda_read_begin.argtypes = [vp, vp, vp]
    lib.wendy_cuda_read_end.restype = ctypes.c_int
    lib.wendy_cuda_read_end.argtypes = [vp]
    lib.wendy_cuda_force_positions.restype = ctypes.c_int
    lib.wendy_cuda_force_positions.argtypes = [vp, ctypes.c_double, ctypes.c_int,
                                               ctypes.POINTER(vp), c_ll_p]
    lib.wendy_cuda_substep.restype = ctypes.c_int
    lib.wendy_cuda_substep.argtypes = [vp, ctypes.c_double, ctypes.c_double, ctypes.c_double, vp]
    lib.wendy_cuda_ext_begin.restype = ctypes.c_int
    lib.wendy_cuda_ext_begin.argtypes = [vp]
    lib.wendy_cuda_substep_async.restype = ctypes.c_int
    lib.wendy_cuda_substep_async.argtypes = [vp, ctypes.c_double, ctypes.c_double, ctypes.c_double, vp]
    lib.wendy_cuda_ext_end.restype = ctypes.c_int
    lib.wendy_cuda_ext_end.argtypes = [vp, ctypes.POINTER(ctypes.c_int)]
    lib.wendy_cuda_read.restype = ctypes.c_int
    lib.wendy_cuda_read.argtypes = [vp, vp, vp]
    lib.wendy_cuda_read_dev.restype = ctypes.c_int
    lib.wendy_cuda_read_dev.argtypes = [vp, vp, vp]
    lib.wendy_cuda_energy.restype = ctypes.c_int
    lib.wendy_cuda_energy.argtypes = [vp, _nd('f8')]
    lib.wendy_serial_cum.restype = ctypes.c_int
    lib.wendy_serial_cum.argtypes = [ctypes.c_double, ctypes.c_longlong, ctypes.c_longlong, _nd('f8')]
    lib.wendy_cuda_stats.restype = ctypes.c_int
    lib.wendy_cuda_stats.argtypes = [vp, _nd('i8'), ctypes.c_int]
    lib.wendy_cuda_create_shard.restype = ctypes.c_int
    lib.wendy_cuda_create_shard.argtypes = [ctypes.POINTER(vp), ctypes.c_longlong, ctypes.c_longlong,
                                            _nd('f8'), _nd('f8'), _nd('i4'), ctypes.c_double,
                                            ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                            _nd('f8'), ctypes.c_longlong, vp]
    lib.wendy_cuda_create_shard_dev.restype = ctypes.c_int
    lib.wendy_cuda_create_shard_dev.argtypes = [ctypes.POINTER(vp), ctypes.c_longlong, ctypes.c_longlong,
                                                vp, vp, vp, ctypes.c_double,
                                                ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                                _nd('f8'), ctypes.c_longlong, vp]
    lib.wendy_cuda_create_shard_m.restype = ctypes.c_int
    lib.wendy_cuda_create_shard_m.argtypes = [ctypes.POINTER(vp), ctypes.c_longlong, ctypes.c_longlong,
                                              _nd('f8'), _nd('f8'), _nd('f8'), _nd('i4'), ctypes.c_double,
                                              ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                              _nd('f8'), ctypes.c_longlong, vp]
    lib.wendy_cuda_shard_mass_total.restype = ctypes.c_int
    lib.wendy_cuda_shard_mass_total.argtypes = [vp, ctypes.c_double, _nd('u8')]
    lib.wendy_cuda_shard_set_mass_offset.restype = ctypes.c_int
    lib.wendy_cuda_shard_set_mass_offset.argtypes = [vp, ctypes.c_ulonglong, ctypes.c_ulonglong]
    lib.wendy_cuda_shard_read_masses.restype = ctypes.c_int
    lib.wendy_cuda_shard_read_masses.argtypes = [vp, _nd('f8')]
    lib.wendy_cuda_potential.restype = ctypes.c_int
    lib.wendy_cuda_potential.argtypes = [vp, ctypes.c_longlong, vp, vp, ctypes.c_longlong, ctypes.c_double,
                                         ctypes.c_double, vp, vp]
    lib.wendy_cuda_energy_individual.restype = ctypes.c_int
    lib.wendy_cuda_energy_individual.argtypes = [vp, vp, vp, ctypes.c_longlong, ctypes.c_double,
                                                 ctypes.c_double, vp, vp]
    lib.wendy_cuda_shard_substep.restype = ctypes.c_int
    lib.wendy_cuda_shard_substep.argtypes = [vp, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                             ctypes.c_double, ctypes.c_longlong, _nd('u4')]
    lib.wendy_cuda_shard_outbox.restype = ctypes.c_int
    lib.wendy_cuda_shard_outbox.argtypes = [vp, ctypes.POINTER(vp), c_ll_p]
    lib.wendy_cuda_shard_inject.restype = ctypes.c_int
    lib.wendy_cuda_shard_inject.argtypes = [vp, vp, ctypes.c_longlong]
    c_ull_p = ctypes.POINTER(ctypes.c_ulonglong)
    lib.wendy_cuda_shard_comm_export.restype = ctypes.c_int
    lib.wendy_cuda_shard_comm_export.argtypes = [vp, c_ull_p, c_ull_p, _nd('u1')]
    lib.wendy_cuda_shard_comm_open.restype = ctypes.c_int
    lib.wendy_cuda_shard_comm_open.argtypes = [vp, _nd('u1'), _nd('u8')]
    lib.wendy_cuda_shard_seed_counts.restype = ctypes.c_int
    lib.wendy_cuda_shard_seed_counts.argtypes = [vp, _nd('i8')]
    lib.wendy_host_set_threads.restype = None
    lib.wendy_host_set_threads.argtypes = [ctypes.c_int]
    lib.wendy_cuda_shard_prepare.restype = ctypes.c_int
    lib.wendy_cuda_shard_prepare.argtypes = [vp, ctypes.c_double, ctypes.c_int]
    lib.wendy_cuda_shard_step_begin.restype = ctypes.c_int
    lib.wendy_cuda_shard_step_begin.argtypes = [vp, ctypes.c_double, ctypes.c_int, ctypes.c_int]
    lib.wendy_cuda_shard_step_end.restype = ctypes.c_int
    lib.wendy_cuda_shard_step_end.argtypes = [vp, ctypes.POINTER(ctypes.c_int), c_ll_p, c_ll_p]
    lib.wendy_cuda_shard_rollback.restype = ctypes.c_int
    lib.wendy_cuda_shard_rollback.argtypes = [vp, ctypes.c_int, c_ll_p]
    lib.wendy_cuda_shard_count.restype = ctypes.c_int
    lib.wendy_cuda_shard_count.argtypes = [vp, c_ll_p]
    lib.wendy_cuda_shard_read.restype = ctypes.c_int
    lib.wendy_cuda_shard_read.argtypes = [vp, _nd('f8'), _nd('f8'), _nd('i4'), c_ll_p]
    lib.wendy_cuda_shard_read_begin.restype = ctypes.c_int
    lib.wendy_cuda_shard_read_begin.argtypes = [vp, _nd('f8'), _nd('f8'), _nd('i4'), c_ll_p]
    lib.wendy_cuda_shard_read_end.restype = ctypes.c_int
    lib.wendy_cuda_shard_read_end.argtypes = [vp]
    lib.wendy_host_prefault.restype = None
    lib.wendy_host_prefault.argtypes = [vp, ctypes.c_ulonglong]
    lib.wendy_cuda_trim.restype = None
    lib.wendy_cuda_trim.argtypes = []
    lib.wendy_cuda_pin.restype = ctypes.c_int
    lib.wendy_cuda_pin.argtypes = [vp, ctypes.c_ulonglong]
    lib.wendy_cuda_unpin.restype = ctypes.c_int
    lib.wendy_cuda_unpin.argtypes = [vp]
    lib.wendy_cuda_debug_layout.restype = ctypes.c_int
    lib.wendy_cuda_debug_layout.argtypes = [vp, _nd('u4'), _nd('f8'), ctypes.c_int]
    lib.wendy_cuda_destroy.restype = None
    lib.wendy_cuda_destroy.argtypes = [vp]
    lib.wendy_cuda_argsort.restype = ctypes.c_int
    lib.wendy_cuda_argsort.argtypes = [_nd('f8'), ctypes.c_longlong, _nd('i4')]
    lib._wendy_nbody_approx_onestep.restype = None
    lib._wendy_nbody_approx_onestep.argtypes = [
        ctypes.c_int, vp, _nd('f8'), _nd('f8'), _nd('f8'), _nd('f8'), ctypes.c_double,
        ctypes.c_double, ctypes.c_int, c_double_p, ctypes.c_double, vp, ctypes.c_int,
        ctypes.POINTER(ctypes.c_int), c_double_p, _nd('f8')]
    _lib = lib
    return lib


#: every symbol include/wendy_b200.h declares (checked by tests/test_abi.py)
EXPORTED = ['wendy_cuda_create', 'wendy_cuda_create_dev', 'wendy_cuda_step', 'wendy_cuda_step_begin', 'wendy_cuda_step_end', 'wendy_cuda_last_call_seconds',
            'wendy_cuda_stage_ahead', 'wendy_cuda_read_begin', 'wendy_cuda_read_end', 'wendy_cuda_force_positions',
            'wendy_cuda_substep', 'wendy_cuda_ext_begin', 'wendy_cuda_substep_async', 'wendy_cuda_ext_end', 'wendy_cuda_read', 'wendy_cuda_read_dev', 'wendy_cuda_energy',
            'wendy_cuda_stats', 'wendy_serial_cum', 'wendy_cuda_create_shard', 'wendy_cuda_create_shard_dev', 'wendy_cuda_create_shard_m', 'wendy_cuda_shard_mass_total', 'wendy_cuda_shard_set_mass_offset', 'wendy_cuda_shard_read_masses', 'wendy_cuda_shard_substep', 'wendy_cuda_shard_outbox',
            'wendy_cuda_shard_inject', 'wendy_cuda_shard_comm_export', 'wendy_cuda_shard_comm_open', 'wendy_cuda_shard_seed_counts', 'wendy_cuda_shard_prepare', 'wendy_cuda_shard_step_begin', 'wendy_cuda_shard_step_end', 'wendy_cuda_shard_rollback', 'wendy_cuda_shard_count', 'wendy_cuda_shard_read', 'wendy_cuda_shard_read_begin', 'wendy_cuda_shard_read_end', 'wendy_cuda_potential', 'wendy_cuda_energy_individual', 'wendy_cuda_set_totmass', 'wendy_cuda_trim', 'wendy_host_prefault', 'wendy_host_set_threads', 'wendy_cuda_pin', 'wendy_cuda_unpin', 'wendy_cuda_debug_layout', 'wendy_cuda_destroy', 'wendy_cuda_last_error',
            'wendy_cuda_argsort', '_wendy_nbody_approx_onestep']


def check(rc):
    if rc < 0:
        raise RuntimeError('wendy_b200: %s (code %d)' % (load().wendy_cuda_last_error().decode(), rc))
    return rc
