"""Shear wall results (`pynite_b200/ShearWall.py`): pier / coupling-beam section forces and story stiffness against the
reference's (tests/golden/shear_walls.npz, written by tests/golden/make_shear_wall_golden.py running the reference).  The
generated mesh, supports, loads, piers and beams themselves are compared in tests/test_mesh_cpu.py (cases 'shear_wall',
'shear_wall_regenerated').

`-m "not gpu"`: the model is given the reference's displacements and quad end forces, so only the wall's own arithmetic is
under test.  `-m gpu`: the same numbers after `analyze_linear` on the B200 (marked `unverified_gpu`: written after the
round's GPU minutes were spent, tests/conftest.py)."""
import os

import numpy as np
import pytest

from pynite_b200 import FEModel3D, analysis
from tests import models
from tests.util import SHEAR_WALL_COMBOS, shear_wall_record

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'shear_walls.npz'))


def _compare(rec, tol):
    for key, ours in rec.items():
        ref = GOLD[key]
        assert ours.shape == ref.shape, key
        if ours.ndim == 3:                                   # [combo, region, (P, M, V, M/(V L))]
            for k in range(4):
                scale = np.abs(ref[:, :, k]).max()
                assert np.abs(ours[:, :, k] - ref[:, :, k]).max() <= tol*scale, (key, k)
        else:
            assert abs(ours/ref - 1) <= tol, key


def test_section_forces_and_stiffness_from_reference_results(capsys):
    m = models.shear_wall_case(FEModel3D)
    analysis.prepare_model(m)
    quads = list(m.quads.values())
    forces = {}
    for c, combo in enumerate(SHEAR_WALL_COMBOS):
        m._D[combo] = GOLD['D'][c].reshape(-1, 1)
        for q in quads:
            forces[(combo, q.ID)] = GOLD['quad_f'][c, q.ID].reshape(24, 1)
    m._element_force = lambda kind, elem_id, combo_name, local: forces[(combo_name, elem_id)]
    _compare(shear_wall_record(m), 1e-12)
    m.shear_walls['W1'].print_piers('1.0E')
    m.shear_walls['W2'].print_coupling_beams('1.0E')
    out = capsys.readouterr().out
    assert '| Wall Pier Results |' in out and '| Wall Coupling Beam Results |' in out and ' P7 ' in out and ' B4 ' in out
    with pytest.raises(Exception, match='does not exist'):
        m.shear_walls['W1'].stiffness('Attic')
    with pytest.raises(Exception, match='falls outside'):
        m.shear_walls['W1'].add_support(elevation=17)
    with pytest.raises(NameError):
        m.add_shear_wall('W1', 1, 10, 10, 1, 'CMU')


def test_prepare_model_generates_and_regenerates_walls():
    m = FEModel3D()
    m.add_material('CMU', 259200.0, 103680.0, 0.17, 0.140)
    m.add_shear_wall(None, 2, 12, 8, 0.667, 'CMU', ky_mod=1)
    w = m.shear_walls['SW1']
    w.add_support()
    w.add_story('Roof', 8)
    analysis.prepare_model(m)
    assert w.is_generated and len(m.quads) == 24 and list(w.piers) == ['P1'] and not w.coupling_beams
    assert 'Stiffness: Roof' in m.load_combos and m.load_combos['Stiffness: Roof'].combo_tags == 'stiffness'
    w.add_opening('Door', 4, 0, 4, 6)
    assert w.needs_update
    analysis.prepare_model(m)
    assert not w.needs_update and len(m.quads) == 24 - 6 and len(w.piers) == 3 and list(w.coupling_beams) == ['B1']


@pytest.mark.gpu
@pytest.mark.unverified_gpu
def test_shear_wall_through_the_gpu():
    m = models.shear_wall_case(FEModel3D)
    m.analyze_linear()
    _compare(shear_wall_record(m), 1e-7)
    m.shear_walls['W1'].print_piers('1.0E')
    m.shear_walls['W1'].print_coupling_beams('1.0E')
