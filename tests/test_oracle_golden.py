"""Pins the CPU oracle (`oracle/`) against the reference: fixtures made by running JWock82/Pynite itself
(tests/golden/make_golden.py) and the textbook values the reference's own tests / examples hold."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import model as om
from tests.util import (FIXTURES, GENERAL_ORIENTATION, PDELTA_FIXTURES, TC_FIXTURES, compare_general_pattern, load_fixture,
                        rel_err)

LINEARISH = [f for f in FIXTURES if f not in TC_FIXTURES]


@pytest.mark.parametrize('name', FIXTURES)
def test_element_matrices(name):
    A, ref, _ = load_fixture(name)
    for kind, fn, count in (('spring', om.spring_Ke, len(A.spring_ks)), ('member', om.member_Ke, len(A.mem_rot)),
                            ('quad', om.quad_Ke, len(A.quad_ijmn)), ('plate', om.plate_Ke, len(A.plate_ijmn))):
        for e in range(count):
            K = fn(A, e)[0]
            R = ref[kind + '_Ke'][e]
            assert rel_err(K, R) < 2e-13, (kind, e)


@pytest.mark.parametrize('name', FIXTURES)
def test_coo_and_csr(name):
    A, ref, _ = load_fixture(name)
    r, c, d = om.assemble_coo(A)
    if name == 'skew_shell':
        # general orientation: which entries are exactly 0.0 depends on rounding; compare
        # after duplicate summation instead of entry by entry
        K = sps.coo_matrix((d, (r, c)), shape=(A.n_dof, A.n_dof)).tocsr()
        Kr = sps.csr_matrix((ref['data'], ref['indices'], ref['indptr']), shape=(A.n_dof, A.n_dof))
        compare_general_pattern(K, Kr)
        return
    assert np.array_equal(r, ref['coo_row'])
    assert np.array_equal(c, ref['coo_col'])
    assert rel_err(d, ref['coo_data']) < 2e-13
    K = om.ke_csr(A)
    assert np.array_equal(K.indptr, ref['indptr'])
    assert np.array_equal(K.indices, ref['indices'])
    assert rel_err(K.data, ref['data']) < 2e-13


@pytest.mark.parametrize('name', FIXTURES)
def test_partition(name):
    A, ref, _ = load_fixture(name)
    d1, d2, D2 = om.partition_D(A)
    assert np.array_equal(np.asarray(d1, dtype=np.int32), ref['d1'])
    assert np.array_equal(np.asarray(d2, dtype=np.int32), ref['d2'])
    assert np.array_equal(np.asarray(D2).reshape(-1), ref['D2'])
    if name == 'skew_shell':
        return
    blocks = om.partition_matrix(om.ke_csr(A), d1, d2)
    for nm, B in zip(('11', '12', '21', '22'), blocks):
        B = B.tocsr()
        assert np.array_equal(B.indptr, ref['K' + nm + '_indptr'])
        assert np.array_equal(B.indices, ref['K' + nm + '_indices'])


@pytest.mark.parametrize('name', FIXTURES)
def test_load_vectors(name):
    A, ref, _ = load_fixture(name)
    for c in range(A.factors.shape[0]):
        P, FER = om.load_vectors(A, c)
        assert rel_err(P.reshape(-1), ref['P'][c]) < 1e-13
        assert rel_err(FER.reshape(-1), ref['FER'][c]) < 1e-12


@pytest.mark.parametrize('name', [f for f in LINEARISH if not f.startswith('pdelta')])
def test_solve_linear(name):
    A, ref, _ = load_fixture(name)
    D, R = om.solve_linear(A)
    for c in range(A.factors.shape[0]):
        assert rel_err(D[c].reshape(-1), ref['D'][c]) < 1e-9
        assert rel_err(R[c].reshape(-1), ref['Rxn'][c]) < 1e-9


@pytest.mark.parametrize('name', PDELTA_FIXTURES)
def test_solve_pdelta(name):
    A, ref, _ = load_fixture(name)
    D, R = om.solve_pdelta(A)
    for c in range(A.factors.shape[0]):
        assert rel_err(D[c].reshape(-1), ref['D'][c]) < 1e-9
        assert rel_err(R[c].reshape(-1), ref['Rxn'][c]) < 1e-8
    P = om.member_axial_P(A, ref['D'][0].reshape(-1, 1))
    assert rel_err(P, ref['member_P']) < 1e-9
    Kg = om.assemble_kg_coo(A, ref['member_P'])
    assert np.array_equal(Kg.row, ref['kg_row'])
    assert np.array_equal(Kg.col, ref['kg_col'])
    assert rel_err(Kg.data, ref['kg_data']) < 1e-12


def test_textbook_space_frame():
    """Logan Example 5.8 (`Examples/Space Frame - Nodal Loads 1.py:59-60`), printed to 4 significant figures."""
    A, ref, _ = load_fixture('space_frame')
    D, R = om.solve_linear(A)
    got = D[0].reshape(-1)[:6]
    expected = np.array([7.098e-5, -0.014, -2.352e-3, -3.996e-3, 1.78e-5, -1.033e-4])
    assert np.allclose(got, expected, rtol=2e-3)
    assert ref['indptr'].shape == (25,) and ref['data'].shape == (108,) and ref['coo_data'].shape == (120,)


@pytest.mark.parametrize('name', ['quad_mesh_xy', 'quad_mesh_yz_ortho', 'plate_mesh_xz', 'skew_shell', 'mat_foundation',
                                  'mixed_building'])
def test_shell_stresses(name):
    """Internal shears / moments / membrane stresses of quads and plates (`Quad3D.py:1017-1239`, `Plate3D.py:590-692`)
    from the reference's own displacement vectors, against what the reference's methods returned."""
    A, ref, _ = load_fixture(name)
    nc = A.factors.shape[0]
    # (compared over all combos at once: a combo of pure uniform settlement has shears that are rounding noise)
    if len(A.quad_ijmn):
        for local, tag in ((True, 'local'), (False, 'global')):
            got = np.array([om.shell_stresses(A, ref['D'][c], 'quad', om.STRESS_POINTS, local) for c in range(nc)])
            for k0, k1 in ((0, 3), (3, 6), (6, 9)):
                assert rel_err(got[..., k0:k1], ref['quad_stress_' + tag][..., k0:k1]) < 1e-10, (tag, k0)
    if len(A.plate_ijmn):
        got = np.array([om.shell_stresses(A, ref['D'][c], 'plate', om.PLATE_FRACTIONS) for c in range(nc)])
        for k0, k1 in ((0, 3), (3, 6), (6, 9)):
            assert rel_err(got[..., k0:k1], ref['plate_stress'][..., k0:k1]) < 1e-9, k0


def test_memoised_oracle_is_bit_identical():
    """The oracle caches per-element results under the bits of the difference vectors the reference's routines read
    (oracle/model.py header): with and without the cache every output must be identical, on a generated mesh (all
    hits), on a jittered one (no hits) and on a frame with releases / rolls / interior nodes."""
    from pynite_b200 import synthetic
    cases = [synthetic.quad_plate_arrays(n=9, n_combos=2), synthetic.quad_plate_arrays(n=7, n_combos=2, mat_springs=0.1),
             synthetic.moment_frame_arrays(bays_x=2, bays_z=2, stories=3, n_combos=2), load_fixture('portal_frame')[0],
             load_fixture('mixed_building')[0]]
    J = synthetic.quad_plate_arrays(n=8, n_combos=2)
    J.xyz = J.xyz + np.random.default_rng(3).uniform(-0.1, 0.1, size=J.xyz.shape)*np.array([1.0, 1.0, 0.0])
    cases.append(J)
    for A in cases:
        res = []
        for memo in (False, True):
            om.MEMOISE = memo
            om._MEMO.clear()
            try:
                r, c, d = om.assemble_coo(A)
                D, R = om.solve_linear(A)
                P, F = om.load_vectors(A, 1)
            finally:
                om.MEMOISE = True
            res.append((r, c, d, np.array(D), np.array(R), P, F))
        for a, b in zip(*res):
            assert np.array_equal(a, b)
