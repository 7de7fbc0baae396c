import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, 'tests') not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, 'tests'))

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + '.npz')))

    return load


def pytest_sessionfinish(session, exitstatus):
    """Writes the parity margins the GPU tests achieved (tests/parity_tolerances.py::record) next to the other run artefacts:
    $CGP_PARITY_REPORT, else gpurun_out/parity_report.txt when that directory exists."""
    try:
        from parity_tolerances import RECORDS
    except Exception:  # noqa: BLE001
        return
    if not RECORDS:
        return
    path = os.environ.get('CGP_PARITY_REPORT')
    if not path:
        out_dir = os.path.join(ROOT, 'gpurun_out')
        if not os.path.isdir(out_dir):
            return
        path = os.path.join(out_dir, 'parity_report.txt')
    with open(path, 'w') as fh:
        fh.write('# achieved parity margins (CUDA path vs oracle / golden fixtures): max |a-b|, max |a-b|/|b|, worst error as a '
                 'fraction of the allowance atol + rtol |b|\n')
        fh.write('%-62s %-28s %-18s %10s %10s %8s %8s %8s\n' % ('test', 'output', 'shape', 'max_abs', 'max_rel', 'used', 'rtol', 'atol'))
        for t, w, e, r, u, rt, at, shp in RECORDS:
            fh.write('%-62s %-28s %-18s %10.2e %10.2e %8.3f %8.0e %8.0e\n' % (t[:62], w[:28], str(tuple(shp)), e, r, u, rt, at))
