import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# several "ranks" of the sharded mode run as threads on ONE GPU in the tests, with kernels that wait for each other:
# a lazily loaded kernel of one rank must not need a context-wide synchronisation while another rank's kernel spins
os.environ.setdefault('CUDA_MODULE_LOADING', 'EAGER')
# ... and their streams must not share a hardware queue (default: 8 connections; a kernel queued behind a spinning one of
# another rank would never start)
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
os.environ.setdefault('OMP_NUM_THREADS', '1')  # reference OpenMP regions are slow in VMs
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session', autouse=True)
def _build_oracle():
    """The C restatement is test infrastructure: build it on demand (gcc only)."""
    subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle'), 'oracle'])


def load_golden(name):
    import numpy
    d = numpy.load(os.path.join(GOLDEN, name + '.npz'))
    return {k: d[k] for k in d.files}
