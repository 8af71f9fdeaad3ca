"""`FEModel3D.analyze_modal` through the GPU: frequencies and mode shapes against the reference's (tests/golden/modal.npz).

The host side (mass matrix, block Lanczos, result storing) is checked against the same golden values in
tests/test_modal_cpu.py with an injected solver; here the `K11^-1` applications are `pn_solve_linear` batches on the device.
First green run on a B200: profiles/r2_final_pytest_gpu_modal.log."""
import os

import numpy as np
import pytest

from tests import models
from tests.util import rel_err

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'modal.npz'))
SOLVED = [n for n, (_, kw) in models.MODAL_MODELS.items() if kw['num_modes']]


@pytest.mark.parametrize('name', SOLVED)
def test_modal_analysis(name):
    from pynite_b200 import FEModel3D
    build, kw = models.MODAL_MODELS[name]
    m = build(FEModel3D)
    m.analyze_modal(log=False, **kw)
    f_ref, modes_ref = GOLD[name + '/frequencies'], GOLD[name + '/modes']
    assert m.solution == 'Modal' and m.modal_info['converged']
    assert np.max(np.abs(m.frequencies/f_ref - 1)) < 1e-9
    for i in range(kw['num_modes']):
        ours, ref = m._D[f'Mode {i + 1}'], modes_ref[:, i:i + 1]
        free = np.abs(ref) != 0.25                               # (the enforced support displacement of `modal_frame`)
        sign = 1.0 if float((ours.T @ ref)[0, 0]) > 0 else -1.0
        assert rel_err(np.where(free, sign*ours, ours), ref) < 1e-7, i
    # the linear drivers still work on the same model afterwards (the mode combinations are dropped again)
    m.analyze_linear()
    assert not [c for c in m.load_combos if c.startswith('Mode')]
