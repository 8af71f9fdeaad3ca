import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNVERIFIED = 'added after the GPU budget of the round was spent: not yet run on hardware'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run with -m gpu on the B200 box)')
    config.addinivalue_line('markers', 'unverified_gpu: GPU test that has not run on hardware yet: expected-failure '
                            '(non-strict) and scheduled after every verified test')


def pytest_collection_modifyitems(config, items):
    """GPU tests written when no GPU time was left (`unverified_gpu`) run LAST and as non-strict xfail: a pass shows as
    XPASS, a failure -- or a CUDA context they might leave in an error state -- cannot touch the verified tests.  Remove the
    marker from a test after its first green run on a B200."""
    late = [it for it in items if it.get_closest_marker('unverified_gpu')]
    if not late:
        return
    for it in late:
        it.add_marker(pytest.mark.gpu)
        it.add_marker(pytest.mark.xfail(strict=False, reason=UNVERIFIED))
    items[:] = [it for it in items if not it.get_closest_marker('unverified_gpu')] + late
