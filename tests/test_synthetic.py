"""The vectorised array builders used by bench.py produce exactly what `flatten_model` produces for the same
model built through the public `add_node/add_quad/add_member/...` API."""
import numpy as np
import pytest

from pynite_b200 import FEModel3D, analysis, synthetic
from pynite_b200.flatten import flatten_model
from tests import models

FIELDS = ['xyz', 'support', 'enf_mask', 'enf_val', 'nspring_k', 'nspring_flag', 'spring_ij', 'spring_ks',
          'spring_active', 'mem_ij', 'mem_props', 'mem_rot', 'mem_rel', 'mem_active', 'quad_ijmn', 'quad_props',
          'plate_ijmn', 'plate_props', 'nl_dof', 'nl_case', 'nl_val', 'ml_member', 'ml_case', 'ml_kind', 'ml_params',
          'qp_elem', 'qp_case', 'qp_val', 'pp_elem', 'pp_case', 'pp_val', 'factors']


def _same(A, B):
    assert A.n_nodes == B.n_nodes and A.cases == B.cases and A.combo_names == B.combo_names
    for f in FIELDS:
        a, b = np.asarray(getattr(A, f)), np.asarray(getattr(B, f))
        assert a.shape == b.shape, f
        assert np.array_equal(a, b), f


@pytest.mark.parametrize('n,nc', [(3, 2), (9, 5)])
def test_quad_plate_arrays_match_flatten(n, nc):
    m = models.big_quad_plate(FEModel3D, n=n, n_combos=nc)
    analysis.prepare_model(m)
    _same(synthetic.quad_plate_arrays(n=n, n_combos=nc), flatten_model(m, list(m.load_combos.values())))


def test_moment_frame_arrays_match_flatten():
    m = models.moment_frame(FEModel3D, bays_x=3, bays_z=2, stories=4, n_combos=18)
    analysis.prepare_model(m)
    _same(synthetic.moment_frame_arrays(bays_x=3, bays_z=2, stories=4, n_combos=18),
          flatten_model(m, list(m.load_combos.values())))


def test_mat_arrays_shape():
    A = synthetic.quad_plate_arrays(n=4, n_combos=2, mat_springs=0.1)
    assert A.nspring_flag[:, 1].tolist() == [3]*25 and A.nspring_k[0, 1] == 0.1*144*0.25
    assert np.all(A.xyz[:, 1] == 0)
