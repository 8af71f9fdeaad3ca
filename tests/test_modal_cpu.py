"""Modal analysis, host side (`pynite_b200/modal.py`): the mass matrix against the reference's (tests/golden/modal.npz,
written by tests/golden/make_modal_golden.py running the reference), and the block Lanczos iteration + result storing of
`analyze_modal` against the reference's frequencies and mode shapes.  The GPU is not available to these tests, so the
`K11^-1` operator is injected: the oracle's stiffness matrix factorised with SuperLU stands in for `modal.EngineInverse`
(test infrastructure only -- the product has no such path).  The same comparison through the GPU is
`tests/test_modal_gpu.py::test_modal_analysis`."""
import os

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

from oracle import model as om
from pynite_b200 import FEModel3D, modal
from tests import models
from tests.util import rel_err

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'modal.npz'))
SOLVED = [n for n, (_, kw) in models.MODAL_MODELS.items() if kw['num_modes']]


def oracle_inverse(model, A, d1):
    K11 = om.ke_csr(A)[d1, :][:, d1]
    lu = spla.splu(sps.csc_matrix(K11))
    return lambda B: lu.solve(np.ascontiguousarray(B))


@pytest.mark.parametrize('name', list(models.MODAL_MODELS))
def test_mass_matrices_match_reference(name):
    build, kw = models.MODAL_MODELS[name]
    m = build(FEModel3D)
    from pynite_b200 import analysis
    analysis.prepare_model(m)
    args = (kw['mass_combo_name'], kw['mass_direction'], kw['gravity'])
    M = m.M(*args)
    assert sps.issparse(M)
    ref = GOLD[name + '/M']
    assert rel_err(np.asarray(M.todense()), ref) < 1e-13
    ours = np.asarray(M.todense())
    assert np.array_equal(np.diag(ours) != 0, np.diag(ref) != 0)           # the stabilisation terms sit on the same DOFs
    differs = (ours != 0) != (ref != 0)      # inclined / rolled members: rounding decides which zeros are exact
    assert np.all(np.abs(ours[differs]) < 1e-14*np.abs(ref).max()) and np.all(np.abs(ref[differs]) < 1e-14*np.abs(ref).max())
    if 'cantilever' in name:
        assert not differs.any()
    subs = [s for pm in m.members.values() for s in pm.sub_members.values()]
    assert len(subs) == len(GOLD[name + '/member_m'])
    for k, s in enumerate(subs):
        assert rel_err(s.m(*args), GOLD[name + '/member_m'][k]) < 1e-13, s.name
        assert rel_err(s.M(*args), GOLD[name + '/member_M'][k]) < 1e-13, s.name
    for k, node in enumerate(m.nodes.values()):
        assert np.array_equal(node.M(*args, characteristic_length=2.5), GOLD[name + '/node_M'][k])
    assert np.isclose(m._calculate_characteristic_length(), float(GOLD[name + '/char_length']), rtol=1e-15)


@pytest.mark.parametrize('name', SOLVED)
def test_analyze_modal_host_logic(name):
    build, kw = models.MODAL_MODELS[name]
    m = build(FEModel3D)
    modal.run_modal(m, kw['num_modes'], kw['mass_combo_name'], kw['mass_direction'], kw['gravity'], False, True,
                    inverse_factory=oracle_inverse)
    f_ref, modes_ref = GOLD[name + '/frequencies'], GOLD[name + '/modes']
    assert m.solution == 'Modal'
    assert list(m.load_combos.keys()) == list(GOLD[name + '/combos'])
    assert np.max(np.abs(m.frequencies/f_ref - 1)) < 1e-9
    assert m.modal_info['converged']
    for i in range(kw['num_modes']):
        ours, ref = m._D[f'Mode {i + 1}'], modes_ref[:, i:i + 1]
        sign = 1.0 if float((ours.T @ ref)[0, 0]) > 0 else -1.0
        assert rel_err(sign*ours + (1 - sign)*np.where(_known(m), ours, 0), ref) < 1e-7, i      # eigsh's sign is arbitrary
    node = list(m.nodes.values())[-1]
    assert node.DY['Mode 1'] == m._D['Mode 1'][node.ID*6 + 1, 0]
    # a second call replaces the mode combinations instead of stacking them
    modal.run_modal(m, 2, kw['mass_combo_name'], kw['mass_direction'], kw['gravity'], False, True,
                    inverse_factory=oracle_inverse)
    assert [c for c in m.load_combos if c.startswith('Mode')] == ['Mode 1', 'Mode 2']
    assert np.max(np.abs(m.frequencies/f_ref[:2] - 1)) < 1e-9


def _known(model):
    """(6N, 1) mask of the supported / enforced DOFs: their mode-shape entries are the enforced values, not eigenvector
    components, so they do not take part in the sign flip."""
    out = np.zeros((6*len(model.nodes), 1), dtype=bool)
    for node in model.nodes.values():
        for k, dof in enumerate(('DX', 'DY', 'DZ', 'RX', 'RY', 'RZ')):
            out[node.ID*6 + k, 0] = bool(getattr(node, 'support_' + dof)) or getattr(node, 'Enforced' + dof) is not None
    return out


def test_lanczos_on_a_known_spectrum():
    """diag(K) = 1..n against M = I plus a rank-one coupling: eigenvalues checked against LAPACK, residuals, and
    M-orthonormality; block smaller than k, exhausted Krylov space (n = 12, k = 11)."""
    rng = np.random.default_rng(3)
    for n, k, block in ((400, 10, 4), (12, 11, None), (60, 5, 8)):
        K = np.diag(np.arange(1.0, n + 1)) + 0.3*np.ones((n, n))/n
        u = rng.standard_normal(n)
        M = np.eye(n)*rng.uniform(0.5, 2.0, n) + 0.05*np.outer(u, u)/n
        lam, phi, info = modal.lanczos_modes(lambda B: np.linalg.solve(K, B), sps.csr_matrix(M), k, block=block)
        import scipy.linalg as sla
        ref = sla.eigh(K, M, eigvals_only=True)[:k]
        assert np.max(np.abs(lam/ref - 1)) < 1e-10, (n, k)
        assert np.max(np.abs(phi.T @ M @ phi - np.eye(k))) < 1e-9
        assert np.max(np.abs(K @ phi - (M @ phi)*lam)) < 1e-7*np.abs(K).max()
    with pytest.raises(TypeError):
        modal.lanczos_modes(lambda B: B, sps.identity(4, format='csr'), 4)


def test_modal_errors():
    m = models.modal_cantilever(FEModel3D)
    with pytest.raises(ValueError):
        m.M('no such combo')
    m.add_load_combo('Empty', {'nothing': 1.0})
    with pytest.raises(Exception, match='massless'):
        m.M('Empty')
    # a member load in a local direction: the reference stops in `T_local.T()` (`Member3D.py:506`)
    m = models.modal_cantilever(FEModel3D)
    m.add_member_pt_load('M3', 'Fy', -1.0, 0.2, case='D1')
    with pytest.raises(TypeError):
        m.M('Consistent Mass Combo')
