"""CPU: the C-ABI library loads and exports every symbol include/wendy_b200.h declares
(no compute calls -- there is no GPU here), and argument validation works without a device."""
import ctypes
import os
import re

import numpy
import pytest

from conftest import ROOT


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__
    if not os.path.exists(os.path.join(ROOT, 'wendy_b200', 'libwendy_b200.so')):
        __graft_entry__.build()
    from wendy_b200 import _lib
    return _lib.load()


def test_header_symbols_all_exported(lib):
    from wendy_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'wendy_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(_?wendy_[a-z_]+)\s*\(', hdr))
    assert declared == set(_lib.EXPORTED)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_no_torch_or_cuda_types_in_signatures():
    hdr = open(os.path.join(ROOT, 'include', 'wendy_b200.h')).read()
    code = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    assert 'cudaStream_t' not in code and 'torch' not in code and 'at::' not in code


def test_argument_validation_needs_no_device(lib):
    h = ctypes.c_void_p()
    x = numpy.zeros(4)
    tot = numpy.ones(1)
    rc = lib.wendy_cuda_create(ctypes.byref(h), 4, x, x, x, tot, -1., 3, 0, 0, 0, None)
    assert rc == -2 and b'multiple' in lib.wendy_cuda_last_error()
    rc = lib.wendy_cuda_create(ctypes.byref(h), 4, x, x, x, tot, -1., 1, 0, 100, 0, None)
    assert rc == -2 and b'cap' in lib.wendy_cuda_last_error()
    assert lib.wendy_cuda_set_totmass(None, tot) == -2


def test_record_layout_matches_reference_struct():
    """struct array_w_index is {int idx; double val;} = 16 bytes (reference wendy/wendy.h:12-16)."""
    from wendy_b200 import _lib
    assert _lib.XI_DTYPE.itemsize == 16 and _lib.XI_DTYPE.fields['val'][1] == 8


def test_product_path_has_no_oracle_or_cpu_fallback():
    for fn in os.listdir(os.path.join(ROOT, 'wendy_b200')):
        if fn.endswith('.py'):
            src = open(os.path.join(ROOT, 'wendy_b200', fn)).read()
            assert 'import oracle' not in src and 'from oracle' not in src, fn


def test_missing_library_fails_loudly(monkeypatch):
    from wendy_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libwendy_b200.so')
    with pytest.raises(ImportError):
        _lib.load()


def test_nbody_argument_errors():
    import wendy_b200
    with pytest.raises(ValueError) as e:  # message pinned by reference tests/test_approx.py:187-196
        next(wendy_b200.nbody([0.], [0.], [1.], 0.1, approx=True))
    assert 'nleap' in str(e.value)
    with pytest.raises(NotImplementedError):
        next(wendy_b200.nbody([0.], [0.], [1.], 0.1))


def test_host_stream_copy_is_exact_for_every_alignment_and_size():
    """The non-temporal host copy behind the bounce-buffered read-out (csrc/api.cu stream_copy): heads and
    tails that are not multiples of 16 / 64 bytes, sizes across the 256 KB block size.  Pure host code."""
    import ctypes
    import numpy
    from wendy_b200 import _lib
    lib = _lib.load()
    lib.wendy_host_stream_copy.restype = None
    lib.wendy_host_stream_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_ulonglong]
    rs = numpy.random.RandomState(1)
    src = rs.randint(0, 256, size=3 * 262144 + 977, dtype=numpy.uint8)
    for n in (0, 1, 15, 16, 17, 63, 64, 65, 1000, 262143, 262144, 262145, 3 * 262144 + 900):
        for so in (0, 1, 8, 13):
            for do in (0, 3, 8, 16, 29):
                dst = numpy.full(n + do + 64, 7, dtype=numpy.uint8)
                lib.wendy_host_stream_copy(dst.ctypes.data + do, src.ctypes.data + so, n)
                assert numpy.array_equal(dst[do:do + n], src[so:so + n]), (n, so, do)
                assert numpy.all(dst[:do] == 7) and numpy.all(dst[do + n:] == 7), (n, so, do)


def test_host_prefault_keeps_contents():
    import numpy
    from wendy_b200 import _lib
    lib = _lib.load()
    a = numpy.arange(300001, dtype=numpy.float64)
    b = a.copy()
    lib.wendy_host_prefault(a.ctypes.data + 8, a.nbytes - 8)
    lib.wendy_host_prefault(None, 0)
    assert numpy.array_equal(a, b)
    c = numpy.empty(5000000)
    lib.wendy_host_prefault(c.ctypes.data, c.nbytes)
    c[:] = 1.
    assert c.sum() == 5000000.


def test_library_is_sm100a_with_tma_in_the_step_kernels_and_no_spills_in_the_plain_one(lib):
    """What DESIGN.md section 3.0 says about the built library, read from its SASS: sm_100a cubins only, bulk
    copies (UBLKCP) and mbarrier waits (SYNCS) in the persistent step kernel, whose plain instance -- the
    dominant kernel of the bench -- does not touch local memory (no register spills)."""
    import shutil
    import subprocess
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump is not installed')
    so = os.path.join(ROOT, 'wendy_b200', 'libwendy_b200.so')
    archs = set(re.findall(r'arch = (sm_\w+)', subprocess.run(['cuobjdump', '-lelf', so], capture_output=True, text=True).stdout
                           + subprocess.run(['cuobjdump', so], capture_output=True, text=True).stdout))
    assert archs <= {'sm_100a'}, archs
    sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
    fn, body = None, {}
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = m.group(1)
            body[fn] = []
        elif fn is not None:
            body[fn].append(line)
    plain = [k for k in body if re.search(r'tile_kernelILi2048ELi512ELi0ELi0ELi1ELi1ELi2E', k)]
    assert len(plain) == 1, plain
    text = '\n'.join(body[plain[0]])
    assert 'UBLKCP' in text and 'SYNCS' in text
    assert not re.search(r'\b(STL|LDL)\b', text)
    assert any('UBLKCP' in '\n'.join(v) for k, v in body.items() if 'wstep_kernel' in k)
