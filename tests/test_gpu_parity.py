"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the golden fixtures the
reference produced and against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): CSR pattern, COO index map and DOF partition bit-exact; element and
assembled values <= 1e-12 relative; displacements and reactions <= 1e-9 relative (1e-6 for P-Delta)."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import model as om
from tests import models
from tests.util import (FIXTURES, GENERAL_ORIENTATION, PDELTA_FIXTURES, TC_FIXTURES, compare_general_pattern, load_fixture,
                        rel_err)

pytestmark = pytest.mark.gpu

TOL_VALUES = 1e-12
TOL_D = 1e-9
TOL_PDELTA = 1e-6


@pytest.fixture(scope='module')
def eng():
    from pynite_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _upload(eng, A):
    eng.set_model(A)
    eng.set_loads(A)
    eng.set_combos(A.factors)


@pytest.mark.parametrize('name', FIXTURES)
def test_element_matrices(eng, name):
    A, ref, _ = load_fixture(name)
    _upload(eng, A)
    for kind in ('spring', 'member', 'quad', 'plate'):
        K = eng.element_ke(kind)
        assert K.shape == ref[kind + '_Ke'].shape
        if K.size:
            for e in range(K.shape[0]):
                assert rel_err(K[e], ref[kind + '_Ke'][e]) < TOL_VALUES, (kind, e)


@pytest.mark.parametrize('name', FIXTURES)
def test_pattern_and_values(eng, name):
    A, ref, _ = load_fixture(name)
    _upload(eng, A)
    s = eng.symbolic(0)
    assert s.n_dof == A.n_dof
    d1, d2, v2 = eng.partition()
    assert np.array_equal(d1, ref['d1']) and np.array_equal(d2, ref['d2']) and np.array_equal(v2, ref['D2'])
    eng.assemble_ke()
    indptr, indices = eng.pattern()
    vals = eng.csr_values()
    n = A.n_dof
    if name in GENERAL_ORIENTATION:
        K = sps.csr_matrix((vals, indices, indptr), shape=(n, n))
        Kr = sps.csr_matrix((ref['data'], ref['indices'], ref['indptr']), shape=(n, n))
        n_only = compare_general_pattern(K, Kr, TOL_VALUES)
        print(f'{name}: {n_only} stored entries differ between the patterns, all rounding noise (SURVEY H1)')
        return
    row, col, data = eng.coo()
    assert s.nnz_coo == len(ref['coo_data'])
    assert np.array_equal(row, ref['coo_row']) and np.array_equal(col, ref['coo_col'])
    assert rel_err(data, ref['coo_data']) < TOL_VALUES
    assert np.array_equal(indptr, ref['indptr'])
    assert np.array_equal(indices, ref['indices'])
    assert rel_err(vals, ref['data']) < TOL_VALUES
    for b, nm in enumerate(('11', '12', '21', '22')):
        ip, ix, dv = eng.block(b)
        assert np.array_equal(ip, ref['K' + nm + '_indptr']), nm
        assert np.array_equal(ix, ref['K' + nm + '_indices']), nm
        assert rel_err(dv, ref['K' + nm + '_data']) < TOL_VALUES, nm


@pytest.mark.parametrize('name', FIXTURES)
def test_assembly_is_deterministic(eng, name):
    A, _, _ = load_fixture(name)
    _upload(eng, A)
    eng.symbolic(0)
    eng.assemble_ke()
    v1 = eng.csr_values().copy()
    eng.assemble_ke()
    assert np.array_equal(v1, eng.csr_values())


@pytest.mark.parametrize('name', FIXTURES + ['synthetic_plate', 'synthetic_frame', 'synthetic_mat'])
def test_class_gather_assembly_equals_general_path(eng, name):
    """The class-based block-gather assembly (pn_assemble.cu) must reproduce the general path (dense element
    values + k_scatter_add) bit for bit: same addends in the same order."""
    import os
    from pynite_b200 import synthetic
    if name == 'synthetic_plate':
        A = synthetic.quad_plate_arrays(n=37, n_combos=2)
    elif name == 'synthetic_frame':
        A = synthetic.moment_frame_arrays(bays_x=4, bays_z=3, stories=5, n_combos=2)
    elif name == 'synthetic_mat':
        A = synthetic.quad_plate_arrays(n=23, n_combos=2, mat_springs=0.1)
    else:
        A, _, _ = load_fixture(name)
    out = []
    for disable in (True, False):
        if disable:
            os.environ['PN_DISABLE_FAST_ASSEMBLY'] = '1'
        else:
            os.environ.pop('PN_DISABLE_FAST_ASSEMBLY', None)
        try:
            _upload(eng, A)
            eng.symbolic(0)
            eng.assemble_ke()
            v1 = eng.csr_values().copy()
            eng.assemble_ke()                 # numeric re-assembly on the fixed pattern
            assert np.array_equal(v1, eng.csr_values())
            out.append(v1)
        finally:
            os.environ.pop('PN_DISABLE_FAST_ASSEMBLY', None)
    assert np.array_equal(out[0], out[1])


@pytest.mark.parametrize('name', FIXTURES)
def test_load_vectors(eng, name):
    A, ref, _ = load_fixture(name)
    _upload(eng, A)
    eng.symbolic(0)
    for c in range(A.factors.shape[0]):
        assert rel_err(eng.load_vector(c, 'P').reshape(-1), ref['P'][c]) < 1e-12
        assert rel_err(eng.load_vector(c, 'FER').reshape(-1), ref['FER'][c]) < 1e-11


@pytest.mark.parametrize('solver', ['direct', 'pcg'])
@pytest.mark.parametrize('name', [f for f in FIXTURES if f not in PDELTA_FIXTURES and f not in TC_FIXTURES])
def test_solve_linear(eng, name, solver):
    A, ref, _ = load_fixture(name)
    if solver == 'pcg' and name in ('quad_mesh_yz_ortho', 'skew_shell'):
        pytest.skip('unsymmetric membrane matrix (kx_mod != ky_mod): PCG refuses, direct solver covers it')
    _upload(eng, A)
    eng.symbolic(0)
    eng.assemble_ke()
    opts = eng.make_opts(solver=solver, tol=1e-13, max_iter=20000)
    D, R, st = eng.solve_linear(opts)
    for c in range(A.factors.shape[0]):
        assert rel_err(D[c], ref['D'][c]) < TOL_D, (c, rel_err(D[c], ref['D'][c]))
        assert rel_err(R[c], ref['Rxn'][c]) < TOL_D, (c, rel_err(R[c], ref['Rxn'][c]))
    assert st.kernel_launches > 0
    if solver == 'direct':
        # the DOF-type classes predicted from the element list are never contradicted by the pattern the reference
        # produced: flat shell fixtures in an axis plane take three classes, everything else one
        ctr = eng.counters()
        assert ctr['dof_class_rejects'] == 0, ctr
        flat = name in ('quad_mesh_xy', 'quad_mesh_yz_ortho', 'plate_mesh_xz', 'mat_foundation', 'rect_mesh_slab')
        assert ctr['dof_classes'] == (3 if flat else 1), (name, ctr)
    # element end forces (Member3D.F/f, Quad3D.F, Plate3D.F, Spring3D.F)
    for c in range(A.factors.shape[0]):
        for kind in ('spring', 'member', 'quad', 'plate'):
            F = eng.element_forces(kind, c)
            if F.size:
                assert rel_err(F, ref[kind + '_F'][c]) < TOL_D, (kind, c)
        f = eng.element_forces('member', c, local=True)
        if f.size:
            assert rel_err(f, ref['member_f'][c]) < TOL_D


@pytest.mark.parametrize('name', PDELTA_FIXTURES)
def test_solve_pdelta(eng, name):
    A, ref, _ = load_fixture(name)
    _upload(eng, A)
    D, R, st = eng.solve_pdelta(eng.make_opts())
    for c in range(A.factors.shape[0]):
        assert rel_err(D[c], ref['D'][c]) < TOL_PDELTA, (c, rel_err(D[c], ref['D'][c]))
        assert rel_err(R[c], ref['Rxn'][c]) < TOL_PDELTA, (c, rel_err(R[c], ref['Rxn'][c]))
    row, col, data = eng.kg_coo(0, first_step=False)
    assert np.array_equal(row, ref['kg_row']) and np.array_equal(col, ref['kg_col'])
    assert rel_err(data, ref['kg_data']) < TOL_PDELTA
    Kg = eng.member_kg(ref['member_P'])
    assert rel_err(Kg, ref['member_Kg']) < 1e-11


@pytest.mark.parametrize('name', ['quad_mesh_xy', 'moment_frame', 'mixed_building'])
def test_against_oracle_on_seeded_inputs(eng, name):
    """Same inputs through the oracle and through the CUDA path (the oracle itself is pinned in
    tests/test_oracle_golden.py)."""
    A, _, _ = load_fixture(name)
    rng = np.random.default_rng(11)
    A.factors = np.round(rng.uniform(-1.5, 1.5, size=(5, A.factors.shape[1])), 3)
    _upload(eng, A)
    D, R, _ = eng.solve_linear(eng.make_opts())
    Do, Ro = om.solve_linear(A)
    for c in range(5):
        assert rel_err(D[c], Do[c].reshape(-1)) < TOL_D
        assert rel_err(R[c], Ro[c].reshape(-1)) < TOL_D


def test_spmm_matches_scipy(eng):
    A, ref, _ = load_fixture('mat_foundation')
    _upload(eng, A)
    eng.symbolic(0)
    eng.assemble_ke()
    n1 = len(ref['d1'])
    K11 = sps.csr_matrix((ref['K11_data'], ref['K11_indices'], ref['K11_indptr']), shape=(n1, n1))
    rng = np.random.default_rng(5)
    for s in (1, 2, 3, 8, 17, 64):
        X = rng.standard_normal((n1, s))
        Y, ms = eng.spmm_k11(X, repeat=2)
        assert rel_err(Y, K11 @ X) < 1e-12
        assert ms > 0


def test_compact_reaction_download_matches_oracle(eng):
    """`reactions_out` moves only the rows of DOFs that can carry a reaction across PCIe when they are fewer than one in
    eight and scatters them into a zero-filled host array: same numbers as the oracle, exact zeros elsewhere."""
    from pynite_b200 import synthetic
    A = synthetic.quad_plate_arrays(n=30, n_combos=3)
    sup = np.unpackbits(np.asarray(A.support, dtype=np.uint8).reshape(-1, 1), axis=1, bitorder='little')[:, :6].reshape(-1)
    assert 8*int(sup.sum()) < A.n_dof                   # the compact path is the one exercised here
    _upload(eng, A)
    R_host = np.full((3, A.n_dof), np.nan)              # must be overwritten completely
    D, R, _ = eng.solve_linear(eng.make_opts(), R_out=R_host)
    Do, Ro = om.solve_linear(A)
    for c in range(3):
        assert rel_err(D[c], Do[c].reshape(-1)) < TOL_D
        assert rel_err(R[c], Ro[c].reshape(-1)) < TOL_D
        assert np.all(R[c][sup == 0] == 0.0)


def test_sweeps_with_odd_right_hand_side_counts(eng):
    """The shared-memory and left-looking sweep kernels work on 8-, 16- and 32-column chunks: column counts that are
    not multiples of the chunk (1, 3, 17, 33) on a model with several elimination levels, against the oracle."""
    from pynite_b200 import synthetic
    import copy
    A = synthetic.quad_plate_arrays(n=40, n_combos=33)
    _upload(eng, A)
    picks = [0, 2, 16, 32]                      # the oracle solves these four combos only (it is a per-combo Python loop)
    A4 = copy.copy(A)
    A4.factors = A.factors[picks]
    Do, _ = om.solve_linear(A4)
    import os
    # three kernel families: shared-memory-resident sweeps (default for these front sizes), the left-looking persistent
    # kernel on every level (what the large levels of a big model use), and the per-step GEMM fallback
    for env in ({}, {'PN_NO_SMEM_SWEEPS': '1'}, {'PN_NO_SMEM_SWEEPS': '1', 'PN_NO_LL_SWEEPS': '1'}):
        os.environ.update(env)
        try:
            for k in (1, 3, 17, 33):
                eng.set_combos(A.factors[:k])
                D, _, st = eng.solve_linear(eng.make_opts(), reactions=False)
                assert D.shape[0] == k
                for q, c in enumerate(picks):
                    if c < k:
                        assert rel_err(D[c], Do[q].reshape(-1)) < TOL_D, (env, k, c)
        finally:
            for name in env:
                os.environ.pop(name, None)


def test_assembly_graph_equals_launch_by_launch(eng):
    """The numeric class / recipe assembly replayed as one CUDA graph writes the same bits as the individual launches
    and as the general path (first assembly of a pattern)."""
    from pynite_b200 import synthetic
    A = synthetic.quad_plate_arrays(n=41, n_combos=2, mat_springs=0.1)
    _upload(eng, A)
    eng.symbolic(0)
    eng.assemble_ke()
    v_general = eng.csr_values().copy()
    eng.assemble_ke()                                   # second assembly: class / recipe maps, graph captured
    v_graph = eng.csr_values().copy()
    eng.assemble_ke()                                   # graph replay
    assert np.array_equal(v_graph, eng.csr_values())
    eng.set_option('assembly_graph', 0)
    try:
        eng.assemble_ke()
        v_plain = eng.csr_values().copy()
    finally:
        eng.set_option('assembly_graph', 1)
    assert np.array_equal(v_general, v_graph) and np.array_equal(v_graph, v_plain)


def test_unstable_and_singular_detection(eng):
    """`Analysis._check_stability` (:111-168) and the residual check of `_solve_unknown_disp` (:228-239)."""
    from pynite_b200 import FEModel3D
    from pynite_b200.engine import UnstableNodes
    A, _, _ = load_fixture('space_frame')
    # release both ends' torsion of every member: zero rotational diagonal somewhere -> nodal instability
    A.mem_rel = np.full_like(A.mem_rel, (1 << 3) | (1 << 9) | (1 << 4) | (1 << 5) | (1 << 10) | (1 << 11))
    _upload(eng, A)
    eng.symbolic(0)
    eng.assemble_ke()
    with pytest.raises(UnstableNodes) as ei:
        eng.check_nodal_stability()
    assert list(ei.value.dofs) == [3, 4, 5]
    # a free-floating beam: non-zero diagonals but rigid-body motion -> singular
    m = FEModel3D()
    m.add_material('S', 29000, 11200, 0.3, 1e-4)
    m.add_section('W', 10, 100, 100, 5)
    m.add_node('A', 0, 0, 0)
    m.add_node('B', 100, 0, 0)
    m.add_member('M', 'A', 'B', 'S', 'W')
    m.def_support('A', True, True, True, False, False, False)
    m.add_node_load('B', 'FY', -1.0)
    with pytest.raises(Exception) as ei:
        m.analyze_linear()
    assert 'singular' in str(ei.value)


def test_public_api_config1():
    """BASELINE config 1 through the drop-in API (`Examples/Space Frame - Nodal Loads 1.py`)."""
    from pynite_b200 import FEModel3D
    _, ref, _ = load_fixture('space_frame')
    m = models.space_frame(FEModel3D)
    m.analyze(check_statics=True)
    assert m.solution == 'Nonlinear TC'
    n1 = m.nodes['N1']
    got = np.array([n1.DX['Combo 1'], n1.DY['Combo 1'], n1.DZ['Combo 1'], n1.RX['Combo 1'], n1.RY['Combo 1'], n1.RZ['Combo 1']])
    assert rel_err(got, ref['D'][0][:6]) < TOL_D
    assert np.allclose(got, [7.098e-5, -0.014, -2.352e-3, -3.996e-3, 1.78e-5, -1.033e-4], rtol=2e-3)
    assert abs(m.nodes['N4'].RxnFY['Combo 1'] - 41.98540472056211) < 1e-8
    K = m.Ke('Combo 1')
    assert K.shape == (24, 24) and K.nnz == 120 and K.tocsr().nnz == 108
    m.analyze_linear()
    assert m.solution == 'Linear'
    assert rel_err(m._D['Combo 1'].reshape(-1), ref['D'][0]) < TOL_D


@pytest.mark.parametrize('name', FIXTURES)
def test_public_api_all_models(name):
    """Every fixture model built through `pynite_b200.FEModel3D` and run with the driver the reference used."""
    from pynite_b200 import FEModel3D
    _, ref, mode = load_fixture(name)
    build, _ = models.SMALL_MODELS[name]
    m = build(FEModel3D)
    if mode == 'linear':
        m.analyze_linear()
    elif mode == 'tc':
        m.analyze()
    elif mode == 'tc_steps':
        m.analyze(**models.TC_STEPS_KW)
    else:
        m.analyze_PDelta()
    assert m.solution == str(ref['solution'])
    tol = TOL_PDELTA if mode == 'pdelta' else TOL_D
    for c, combo in enumerate(m.load_combos):
        assert rel_err(m._D[combo].reshape(-1), ref['D'][c]) < tol, (combo, rel_err(m._D[combo].reshape(-1), ref['D'][c]))
        assert rel_err(m._Rxn[combo].reshape(-1), ref['Rxn'][c]) < tol, combo
    node = list(m.nodes.values())[-1]
    combo = list(m.load_combos)[0]
    assert node.DY[combo] == m._D[combo][node.ID*6 + 1, 0]
    assert node.RxnFY[combo] == m._Rxn[combo][node.ID*6 + 1, 0]
    subs = [s for pm in m.members.values() for s in pm.sub_members.values()]
    assert [s.name for s in subs] == [str(x) for x in ref['sub_names']]
    if subs and mode not in ('tc', 'tc_steps'):
        assert rel_err(subs[0].F(combo).reshape(-1), ref['member_F'][0][0]) < 1e-6


def test_tc_iteration_braced_frame():
    """Tension-only braces / springs and a compression-only support spring: `_check_TC_convergence`."""
    from pynite_b200 import FEModel3D
    _, ref, _ = load_fixture('braced_tc')
    m = models.braced_tc(FEModel3D)
    m.analyze()
    for c, combo in enumerate(m.load_combos):
        assert rel_err(m._D[combo].reshape(-1), ref['D'][c]) < TOL_D, combo
        assert rel_err(m._Rxn[combo].reshape(-1), ref['Rxn'][c]) < TOL_D, combo


def test_medium_quad_plate_properties():
    """A 60x60 quad plate (21.6k DOF): size-independent properties -- equilibrium of reactions with applied
    loads, linearity in the combo factors, residual of the solve, determinism."""
    from pynite_b200 import FEModel3D
    m = models.big_quad_plate(FEModel3D, n=60, n_combos=4)
    # combo 4 = combo 1 + 2 * combo 2 (linearity check)
    f1, f2 = m.load_combos['C1'].factors, m.load_combos['C2'].factors
    m.load_combos['C4'].factors = {k: f1[k] + 2*f2[k] for k in f1}
    m.analyze_linear(check_stability=True)
    D = {c: m._D[c].reshape(-1) for c in m.load_combos}
    assert rel_err(D['C4'], D['C1'] + 2*D['C2']) < 1e-9
    assert m.last_stats['max_rel_residual'] < 1e-10
    eng = m._engine
    for j, c in enumerate(m.load_combos):
        F = (eng.load_vector(j, 'P') - eng.load_vector(j, 'FER')).reshape(-1, 6)
        R = m._Rxn[c].reshape(-1, 6)
        assert abs(F[:, 2].sum() + R[:, 2].sum()) < 1e-8*abs(F[:, 2]).sum()
    m2 = models.big_quad_plate(FEModel3D, n=60, n_combos=4)
    m2.load_combos['C4'].factors = dict(m.load_combos['C4'].factors)
    m2.analyze_linear()
    assert np.array_equal(m2._D['C3'], m._D['C3'])


def test_full_size_config3_properties():
    """BASELINE config 3 at full size (1 506 006 DOF) with 8 combos through the C ABI: size-independent properties.
    The reference cannot run this size (O(N^2) helpers), so the checks are: bit-exact closed-form pattern
    counts, linearity in the combo factors, global equilibrium of the reactions, the residual of the solve and
    run-to-run determinism."""
    from pynite_b200 import synthetic
    from pynite_b200.engine import Engine
    A = synthetic.quad_plate_arrays(n=500, n_combos=8)
    A.factors[7] = A.factors[0] + 2.0*A.factors[1]
    e = Engine(0)
    try:
        e.set_model(A); e.set_loads(A); e.set_combos(A.factors)
        s = e.symbolic(0)
        assert (s.n_dof, s.nnz_coo, s.nnz_csr) == (1506006, 53000000, 29540014)      # SURVEY 8 closed forms
        e.assemble_ke()
        D, R, st = e.solve_linear(e.make_opts())
        assert st.max_rel_residual < 1e-6
        assert rel_err(D[7], D[0] + 2.0*D[1]) < 1e-9
        for c in (0, 3):
            F = (e.load_vector(c, 'P') - e.load_vector(c, 'FER')).reshape(-1, 6)
            Rc = R[c].reshape(-1, 6)
            assert abs(F[:, 2].sum() + Rc[:, 2].sum()) < 1e-7*abs(F[:, 2]).sum()
        e.assemble_ke()
        D2, _, _ = e.solve_linear(e.make_opts(), reactions=False)
        assert np.array_equal(D, D2)
    finally:
        e.close()


def test_public_inspection_calls():
    """`FEModel3D.Ke/Kg/FER/P/D` as stand-alone calls and after an analysis (`FEModel3D.py:1551, 1802, 2136, 2184, 2237`)."""
    from pynite_b200 import FEModel3D
    _, ref, _ = load_fixture('portal_frame')
    m = models.portal_frame(FEModel3D)
    K = m.Ke('1.4D')                       # before any analysis: the call prepares and uploads the model itself
    assert K.shape == (42, 42)
    Kr = sps.csr_matrix((ref['data'], ref['indices'], ref['indptr']), shape=(42, 42))
    assert abs(K.tocsr() - Kr).max() < 1e-12*abs(Kr).max()
    m.analyze_linear(check_statics=True)
    for c, combo in enumerate(m.load_combos):
        assert rel_err(m.FER(combo).reshape(-1), ref['FER'][c]) < 1e-11
        assert rel_err(m.P(combo).reshape(-1), ref['P'][c]) < 1e-12
        assert rel_err(m.D(combo).reshape(-1), ref['D'][c]) < TOL_D
    _, refp, _ = load_fixture('pdelta_column')
    p = models.pdelta_column(FEModel3D)
    p.analyze_PDelta()
    Kg = p.Kg('C1', first_step=False)
    assert np.array_equal(Kg.row, refp['kg_row']) and np.array_equal(Kg.col, refp['kg_col'])
    assert rel_err(Kg.data, refp['kg_data']) < TOL_PDELTA
    # linear re-analysis of a model whose last solution was P-Delta keeps the P-Delta end-force branch
    p.analyze_linear()
    assert p.solution == 'Linear'


def test_load_stepping_with_a_spring_that_flips_mid_stepping():
    """`analyze(num_steps=4, spring_tolerance=0.2)` on a propped cantilever whose compression-only prop engages only
    after the tip has moved by more than the tolerance: steps are undone and redone (`Analysis._first_order` :262-305,
    `_sum_displacements` :882-914) and the result is path dependent -- it differs from the one-step analysis by ~5 %.
    Compared with the fixture the reference produced with the same arguments, including element end forces of every
    combo (saved per combo: the engine only keeps the last one resident)."""
    from pynite_b200 import FEModel3D
    _, ref, _ = load_fixture('tc_load_steps')
    m = models.tc_load_steps(FEModel3D)
    m.analyze(**models.TC_STEPS_KW)
    one = models.tc_load_steps(FEModel3D)
    one.analyze()                 # one step, zero tolerance (with the 0.2 tolerance a single step flip-flops, also in the reference)
    differs = 0.0
    for c, combo in enumerate(m.load_combos):
        assert rel_err(m._D[combo].reshape(-1), ref['D'][c]) < TOL_D, combo
        assert rel_err(m._Rxn[combo].reshape(-1), ref['Rxn'][c]) < TOL_D, combo
        differs = max(differs, rel_err(one._D[combo].reshape(-1), ref['D'][c]))
    assert differs > 1e-3                     # the stepping really changed the answer
    subs = [s for pm in m.members.values() for s in pm.sub_members.values()]
    for c, combo in enumerate(m.load_combos):      # first combos are no longer resident on the engine
        for k, sub in enumerate(subs):
            assert rel_err(sub.F(combo).reshape(-1), ref['member_F'][c][k]) < TOL_D, (combo, k)
            assert rel_err(sub.f(combo).reshape(-1), ref['member_f'][c][k]) < TOL_D, (combo, k)
        sp = m.springs['TIE']
        assert rel_err(sp.F(combo).reshape(-1), ref['spring_F'][c][0]) < TOL_D       # also while inactive (Spring3D.D :222-225)


def test_inspection_calls_do_not_disturb_results():
    """`FER/P/Ke` for a combo that is not resident (every combo but the last after a per-combo analysis)
    must not invalidate stored results: `member.f()` of another combo still works afterwards."""
    from pynite_b200 import FEModel3D
    _, ref, _ = load_fixture('braced_tc')
    m = models.braced_tc(FEModel3D)
    m.analyze()
    combos = list(m.load_combos)
    fer = m.FER(combos[0])
    assert fer.shape == (36, 1)
    K = m.Ke(combos[0])
    assert K.shape == (36, 36)
    sub = [s for pm in m.members.values() for s in pm.sub_members.values()][0]
    for c, combo in enumerate(combos):
        assert rel_err(sub.F(combo).reshape(-1), ref['member_F'][c][0]) < TOL_D, combo
        assert rel_err(m._D[combo].reshape(-1), ref['D'][c]) < TOL_D


@pytest.mark.parametrize('name', ['quad_mesh_xy', 'quad_mesh_yz_ortho', 'plate_mesh_xz', 'skew_shell', 'mat_foundation',
                                  'mixed_building'])
def test_shell_stress_recovery(eng, name):
    """batched `Quad3D.shear/moment/membrane` (`Quad3D.py:1017-1239`) and `Plate3D.shear/moment/membrane`
    (`Plate3D.py:590-692`) over elements x combos on the device, against the values the reference's methods returned.
    Bar 1e-9 of the largest value of each group (shear / moment / membrane) over the model's combos."""
    A, ref, _ = load_fixture(name)
    _upload(eng, A)
    eng.solve_linear(eng.make_opts(), reactions=False)
    nc = A.factors.shape[0]
    if len(A.quad_ijmn):
        for local, tag in ((True, 'local'), (False, 'global')):
            got = eng.shell_stresses('quad', 0, nc, om.STRESS_POINTS, local)
            for k0, k1 in ((0, 3), (3, 6), (6, 9)):
                assert rel_err(got[..., k0:k1], ref['quad_stress_' + tag][..., k0:k1]) < TOL_D, (tag, k0)
    if len(A.plate_ijmn):
        want = ref['plate_stress']
        got = np.empty_like(want)
        # plate points are local (x, y): the fixture's fractions of (width, height) differ per element only through w, h,
        # which are uniform in these meshes
        p4 = A.xyz[A.plate_ijmn[0]]
        w, h = np.linalg.norm(p4[0] - p4[1]), np.linalg.norm(p4[0] - p4[3])
        got = eng.shell_stresses('plate', 0, nc, np.array(om.PLATE_FRACTIONS)*np.array([w, h]), True)
        for k0, k1 in ((0, 3), (3, 6), (6, 9)):
            assert rel_err(got[..., k0:k1], want[..., k0:k1]) < TOL_D, k0


def test_shell_stress_public_methods():
    """`Quad3D.shear/moment/membrane` and `Plate3D.*` on the drop-in objects, local and global."""
    from pynite_b200 import FEModel3D
    _, ref, _ = load_fixture('mixed_building')
    m = models.mixed_building(FEModel3D)
    m.analyze_linear()
    combos = list(m.load_combos)
    for c, combo in enumerate(combos):
        for e, q in enumerate(m.quads.values()):
            for k, (xi, eta) in enumerate(om.STRESS_POINTS[:3]):
                want = ref['quad_stress_local'][c][e][k]
                scale = np.abs(ref['quad_stress_local'][..., 3:6]).max()
                assert np.abs(q.moment(xi, eta, True, combo).reshape(-1) - want[3:6]).max() < TOL_D*scale
                assert q.shear(xi, eta, True, combo).shape == (2, 1)
                wantg = ref['quad_stress_global'][c][e][k]
                assert np.abs(q.membrane(xi, eta, False, combo).reshape(-1) - wantg[6:9]).max() < \
                    TOL_D*np.abs(ref['quad_stress_global'][..., 6:9]).max()
                assert q.shear(xi, eta, False, combo).shape == (3, 1)
        for e, p in enumerate(m.plates.values()):
            fx, fy = om.PLATE_FRACTIONS[4]
            want = ref['plate_stress'][c][e][4]
            got = p.moment(fx*p.width(), fy*p.height(), True, combo).reshape(-1)
            assert np.abs(got - want[3:6]).max() < TOL_D*np.abs(ref['plate_stress'][..., 3:6]).max()


def test_pdelta_near_buckling_uses_preconditioned_gmres(eng):
    """at 0.6 ... 0.95 P_cr the stationary iteration on the elastic factor contracts too slowly; the columns
    are finished by GMRES preconditioned with that factor instead of an unpivoted LU of the near-singular / indefinite
    Ke + Kg.  Parity with the reference (SuperLU with partial pivoting) at the P-Delta bar."""
    A, ref, _ = load_fixture('pdelta_near_buckling')
    _upload(eng, A)
    eng.reset_stats()
    D, R, st = eng.solve_pdelta(eng.make_opts())
    for c in range(A.factors.shape[0]):
        assert rel_err(D[c], ref['D'][c]) < TOL_PDELTA, (c, rel_err(D[c], ref['D'][c]))
        assert rel_err(R[c], ref['Rxn'][c]) < TOL_PDELTA, c
    ctr = eng.counters()
    assert ctr['pdelta_fallback_columns'] == A.factors.shape[0]
    assert st.max_rel_residual < 1e-9


# ---------------------------------------------------------------------------------------------------------------------
# Mid-size parity: the CUDA path against the oracle at sizes where the production kernels of the big BASELINE
# configurations run (pivot groups beyond the first, rank-64 strips, front groups on side streams, shared-memory and
# left-looking sweeps), asserted from the library's path counters rather than assumed.
# ---------------------------------------------------------------------------------------------------------------------
def _refined_splu(K11, rhs):
    """Independent high-accuracy solution: SuperLU + iterative refinement with a long-double residual."""
    from scipy.sparse.linalg import splu
    lu = splu(K11.tocsc())          # COLAMD, like spsolve (MMD_AT_PLUS_A takes minutes at this size)
    b = np.asarray(rhs, dtype=float).reshape(-1)
    x = lu.solve(b)
    Kc = K11.tocoo()
    vals, rows, cols = Kc.data.astype(np.longdouble), Kc.row, Kc.col
    for _ in range(4):
        ax = np.zeros(len(b), dtype=np.longdouble)
        np.add.at(ax, rows, vals*x.astype(np.longdouble)[cols])
        r = (b.astype(np.longdouble) - ax).astype(float)
        dx = lu.solve(r)
        x = x + dx
        if np.abs(dx).max() < 1e-15*np.abs(x).max():
            break
    return x


def _production_paths_taken(ctr, want_later_groups, want_ll=True):
    assert ctr['lu_update'] > 0 and ctr['lu_strip'] > 0, ctr
    assert ctr['split_levels'] > 0, ctr
    assert ctr['sweep_smem'] > 0 and (ctr['sweep_ll'] > 0 or not want_ll), ctr
    if want_later_groups:
        assert ctr['lu_update_later_groups'] > 0 and ctr['max_front_pivots'] > 256, ctr


@pytest.mark.parametrize('classes', [1, 0])
def test_midsize_plate_150_against_oracle(eng, classes):
    """BASELINE config 3 shape at 150 x 150 quads (136 806 DOF), 2 combos: pattern bit-exact, values 1e-12,
    displacements and reactions against the oracle's raw `spsolve` (what the reference calls, `Analysis.py:217`).
    With the class-split ordering (default: membrane / bending / drilling fronts; every level then fits the
    shared-memory sweeps at this size) and with the node-based ordering (left-looking sweeps on the top levels)."""
    from pynite_b200 import synthetic
    A = synthetic.quad_plate_arrays(n=150, n_combos=2)
    eng.set_option('dof_classes', classes)
    try:
        _upload(eng, A)
        eng.reset_stats()
        D, R, st = eng.solve_linear(eng.make_opts())
    finally:
        eng.set_option('dof_classes', 1)
    ctr = eng.counters()
    _production_paths_taken(ctr, want_later_groups=True, want_ll=not classes)
    assert ctr['dof_classes'] == (3 if classes else 1) and ctr['dof_class_rejects'] == 0, ctr
    K = om.ke_csr(A)
    indptr, indices = eng.pattern()
    assert np.array_equal(indptr, K.indptr) and np.array_equal(indices, K.indices)
    assert rel_err(eng.csr_values(), K.data) < TOL_VALUES
    Do, Ro = om.solve_linear(A)
    errs = [rel_err(D[c], Do[c].reshape(-1)) for c in range(2)]
    print('150x150 plate: |D - spsolve| / |D| per combo =', errs, 'residual', st.max_rel_residual, ctr)
    for c in range(2):
        assert errs[c] < TOL_D, errs
        assert rel_err(R[c], Ro[c].reshape(-1)) < TOL_D


@pytest.mark.parametrize('case', ['plate_xy', 'plate_yz', 'mat', 'frame', 'plate_jitter', 'plate_ortho'])
def test_dof_class_ordering_equals_node_ordering(eng, case):
    """`pn_set_option("dof_classes")`: the class-split ordering (1, default), the node-based ordering (0) and a prediction
    the device verification has to reject (2) give the same displacements and reactions; flat shell models use three
    classes, a frame keeps one."""
    from pynite_b200 import synthetic
    if case == 'plate_xy':
        A = synthetic.quad_plate_arrays(n=72, n_combos=3)
    elif case == 'plate_jitter':                # general in-plane geometry: still planar, still three classes
        A = synthetic.quad_plate_arrays(n=56, n_combos=2, jitter=0.1)
    elif case == 'plate_ortho':                 # kx_mod != ky_mod: unsymmetric membrane block, general LU mode
        A = synthetic.quad_plate_arrays(n=64, n_combos=2)
        A.quad_props = np.ascontiguousarray(A.quad_props)
        A.quad_props[:, 3] = 0.6
    elif case == 'mat':
        A = synthetic.quad_plate_arrays(n=60, n_combos=2, mat_springs=0.1)
    elif case == 'frame':
        A = synthetic.moment_frame_arrays(bays_x=4, bays_z=3, stories=6, n_combos=2)
    else:
        A, _, _ = load_fixture('quad_mesh_yz_ortho')
    out = {}
    try:
        for mode in (1, 0, 2):
            eng.set_option('dof_classes', mode)
            _upload(eng, A)
            eng.reset_stats()
            D, R, st = eng.solve_linear(eng.make_opts())
            out[mode] = (D, R, eng.counters(), st.max_rel_residual)
    finally:
        eng.set_option('dof_classes', 1)
    shells = case != 'frame'
    assert out[1][2]['dof_classes'] == (3 if shells else 1), out[1][2]
    assert out[0][2]['dof_classes'] == 1 and out[0][2]['dof_class_rejects'] == 0
    assert out[2][2]['dof_classes'] == 1 and out[2][2]['dof_class_rejects'] == 1, out[2][2]
    for mode in (0, 2):
        assert rel_err(out[mode][0], out[1][0]) < 1e-10 and rel_err(out[mode][1], out[1][1]) < 1e-9, (mode, case)
    assert np.array_equal(out[0][0], out[2][0])          # the rejected prediction falls back to exactly the node ordering
    if shells:
        Do, Ro = om.solve_linear(A)
        for c in range(len(Do)):
            assert rel_err(out[1][0][c], Do[c].reshape(-1)) < TOL_D


def test_rectangle_mesh_model_and_mesh_queries():
    """A slab with an opening built by `FEModel3D.add_rectangle_mesh` (pynite_b200/Mesh.py; generated names, order and
    coordinates are checked against the reference's generator in tests/test_mesh_cpu.py) through `analyze_linear`, and the
    mesh result queries `Mesh.max_* / min_*` (`Mesh.py:172-716`: corners and centre of every element, one batched
    `pn_shell_stresses` table per combination here) against what the reference's element-by-element loops returned."""
    from pynite_b200 import FEModel3D
    A, ref, _ = load_fixture('rect_mesh_slab')
    m = models.rect_mesh_slab(FEModel3D)
    m.analyze_linear()
    for c, combo in enumerate(m.load_combos):
        assert rel_err(m._D[combo].reshape(-1), ref['D'][c]) < TOL_D, combo
    queries = [('shear', d) for d in ('Qx', 'Qy', 'QX', 'QY')] + [('moment', d) for d in ('Mx', 'My', 'Mxy', 'MX', 'MY', 'MZ')] + \
              [('membrane', d) for d in ('Sx', 'Sy', 'Sxy', 'SX', 'SY')]
    gold = ref['mesh_extremes']
    mesh = m.meshes['SLAB']
    scale = np.abs(gold).max(axis=(0, 1, 3))
    for b, combo in enumerate(m.load_combos):
        for q, (what, direction) in enumerate(queries):
            hi = getattr(mesh, 'max_' + what)(direction, combo)
            lo = getattr(mesh, 'min_' + what)(direction, combo)
            assert abs(hi - gold[0, b, q, 0]) <= 1e-9*scale[q] + 1e-300, (combo, what, direction, hi, gold[0, b, q, 0])
            assert abs(lo - gold[0, b, q, 1]) <= 1e-9*scale[q] + 1e-300, (combo, what, direction, lo, gold[0, b, q, 1])
    with pytest.raises(Exception):
        mesh.max_moment('Mq', 'C1')


def test_pcg_reports_non_convergence(eng):
    """`PN_ERR_NOT_CONVERGED`: a PCG run that hits its iteration cap raises instead of returning a partial answer (the
    reference has no iterative solver; its analogue is the `singular` exception of `Analysis.py:239`)."""
    from pynite_b200 import synthetic
    from pynite_b200.engine import NotConverged
    A = synthetic.quad_plate_arrays(n=40, n_combos=2)
    _upload(eng, A)
    with pytest.raises(NotConverged):
        eng.solve_linear(eng.make_opts(solver='pcg', tol=1e-13, max_iter=5))
    D, R, st = eng.solve_linear(eng.make_opts(solver='pcg', tol=1e-12, max_iter=200000))       # the context stays usable
    Dd, Rd, _ = eng.solve_linear(eng.make_opts())
    assert rel_err(D, Dd) < 1e-8 and st.iterations > 5


def test_output_combo_range(eng):
    """`pn_set_output_combos`: the host arrays of a solve hold a range of the combinations only (a rank of a multi-GPU run
    delivers its share); same rows as the full download, compact and full reaction paths."""
    from pynite_b200 import synthetic
    for A in (synthetic.quad_plate_arrays(n=40, n_combos=5), synthetic.moment_frame_arrays(bays_x=3, bays_z=2, stories=4, n_combos=5)):
        _upload(eng, A)
        D, R, _ = eng.solve_linear(eng.make_opts())
        try:
            eng.set_output_combos(1, 3)
            D1, R1, _ = eng.solve_linear(eng.make_opts())
            assert D1.shape[0] == 3 and np.array_equal(D1, D[1:4]) and np.array_equal(R1, R[1:4])
            eng.set_output_combos(4, 9)                     # clipped to the last combination
            D2, R2, _ = eng.solve_linear(eng.make_opts())
            assert D2.shape[0] == 1 and np.array_equal(D2[0], D[4]) and np.array_equal(R2[0], R[4])
        finally:
            eng.set_output_combos(0, -1)
        D3, R3, _ = eng.solve_linear(eng.make_opts())
        assert np.array_equal(D3, D) and np.array_equal(R3, R)


def test_subtree_distributed_solve_on_eight_ranks():
    """The class-split tree (two roots on the top level, cost-balanced cut) shared out over 8 ranks, against one rank."""
    from pynite_b200 import distsolve, synthetic
    from pynite_b200.engine import Engine
    A = synthetic.quad_plate_arrays(n=96, n_combos=3)
    e0 = Engine(0)
    try:
        _upload(e0, A)
        D0, R0, _ = e0.solve_linear(e0.make_opts())
    finally:
        e0.close()
    world = 8
    group = distsolve.LocalGroup(world)
    engines = [Engine(0) for _ in range(world)]

    def rank_fn(r):
        e = engines[r]
        group.attach(e, r)
        _upload(e, A)
        e.reset_stats()
        D, R, st = e.solve_linear(e.make_opts())
        return D, R, e.counters()
    try:
        out = group.run(rank_fn)
    finally:
        for e in engines:
            e.close()
    assert not group.errors
    for r, (D, R, ctr) in enumerate(out):
        assert ctr['dist_cut_level'] >= 0 and ctr['dof_classes'] == 3, ctr
        assert np.array_equal(D, D0) and np.array_equal(R, R0), r


def test_midsize_plate_256_against_spsolve_and_refined_splu(eng):
    """256 x 256 quads (396 294 DOF; the top separator has 1 542 pivots, so fronts run several 256-pivot groups).
    SURVEY H3: cond(K11) grows like n^4, and raw `spsolve` -- the reference's answer -- carries a forward error of
    cond x 1e-16 of its own.  The CUDA solution (refined with a double-double residual) is compared with BOTH the raw
    `spsolve` and a refined `splu` solution; the 1e-9 bar is asserted against the refined one, and the distance of the
    raw solve from both is printed (and bounded) so that the record shows which solution 1e-9 holds against."""
    from pynite_b200 import synthetic
    from scipy.sparse.linalg import spsolve
    A = synthetic.quad_plate_arrays(n=256, n_combos=1)
    _upload(eng, A)
    eng.reset_stats()
    D, R, st = eng.solve_linear(eng.make_opts())
    ctr = eng.counters()
    # (class-split ordering: the largest front has 768 pivots and every level still fits the shared-memory sweeps; the
    # left-looking sweeps are covered by the node-ordering run at 150 x 150 and by the full-size config 3 test)
    _production_paths_taken(ctr, want_later_groups=True, want_ll=False)
    d1, d2, D2 = om.partition_D(A)
    K11, K12, _, _ = om.partition_matrix(om.ke_csr(A), d1, d2)
    P, FER = om.load_vectors(A, 0)
    rhs = np.subtract(np.subtract(P[d1, :], FER[d1, :]), K12.tocsr() @ D2)
    x_raw = np.asarray(spsolve(K11.tocsr(), rhs)).reshape(-1)
    x_ref = _refined_splu(K11, rhs)
    x_gpu = D[0][np.asarray(d1)]
    e_ref, e_raw, raw_vs_ref = rel_err(x_gpu, x_ref), rel_err(x_gpu, x_raw), rel_err(x_raw, x_ref)
    print(f'256x256 plate: |gpu - refined splu| = {e_ref:.2e}, |gpu - raw spsolve| = {e_raw:.2e}, '
          f'|raw spsolve - refined splu| = {raw_vs_ref:.2e}, gpu residual {st.max_rel_residual:.2e}', ctr)
    assert e_ref < TOL_D
    assert e_raw < max(TOL_D, 3*raw_vs_ref)          # the GPU result is as close to the raw solve as the refined one is


def test_midsize_frame_against_oracle():
    """BASELINE config 2 shape: 8 x 8 bays x 10 stories with one interior node per column (`PhysMember.descritize`)
    and end releases on a fifth of the beams, 4 combos, through `FEModel3D.analyze_linear`."""
    from pynite_b200 import FEModel3D
    m = models.moment_frame(FEModel3D, bays_x=8, bays_z=8, stories=10, interior=1, n_combos=4)
    beams = [name for name, pm in m.members.items() if pm.section.name == 'Beam']
    for k, name in enumerate(beams):
        if k % 5 == 0:
            m.def_releases(name, Rzi=True) if k % 10 == 0 else m.def_releases(name, Ryj=True, Rzj=True)
    m.analyze_linear()
    A = m._run.arrays
    assert A.n_dof > 10000 and len(A.mem_rot) > 2500
    Do, Ro = om.solve_linear(A)
    for c, combo in enumerate(m.load_combos):
        assert rel_err(m._D[combo].reshape(-1), Do[c].reshape(-1)) < TOL_D, combo
        assert rel_err(m._Rxn[combo].reshape(-1), Ro[c].reshape(-1)) < TOL_D, combo


@pytest.mark.parametrize('solver', ['direct', 'pcg'])
def test_midsize_mat_on_springs_against_oracle(eng, solver):
    """BASELINE config 4 shape: 150 x 150 quads in the XZ plane on Winkler springs (136 806 DOF), direct and PCG."""
    from pynite_b200 import synthetic
    A = synthetic.quad_plate_arrays(n=150, n_combos=2, mat_springs=0.1)
    _upload(eng, A)
    eng.reset_stats()
    D, R, st = eng.solve_linear(eng.make_opts(solver=solver, tol=1e-12, max_iter=200000))
    if solver == 'direct':
        _production_paths_taken(eng.counters(), want_later_groups=True, want_ll=False)
        assert eng.counters()['dof_classes'] == 3
    Do, Ro = _mat_oracle(A)
    for c in range(2):
        assert rel_err(D[c], Do[c].reshape(-1)) < TOL_D, (solver, c, rel_err(D[c], Do[c].reshape(-1)), st.iterations)
        assert rel_err(R[c], Ro[c].reshape(-1)) < TOL_D, (solver, c)


_MAT_ORACLE = {}


def _mat_oracle(A):
    key = (A.n_nodes, A.factors.tobytes())
    if key not in _MAT_ORACLE:
        _MAT_ORACLE.clear()
        _MAT_ORACLE[key] = om.solve_linear(A)
    return _MAT_ORACLE[key]


def test_midsize_pdelta_frame_against_oracle():
    """BASELINE config 5 shape: 6 x 6 bays x 12 stories, one interior node per column, 8 combos, `analyze_PDelta`
    against the oracle's two-step procedure with one SuperLU factorisation of Ke + Kg(P) per combo."""
    from pynite_b200 import FEModel3D
    m = models.moment_frame(FEModel3D, bays_x=6, bays_z=6, stories=12, interior=1, n_combos=8)
    m.analyze_PDelta()
    A = m._run.arrays
    Do, Ro = om.solve_pdelta(A)
    ctr = m._engine.counters()
    assert ctr['pdelta_stationary_batches'] + ctr['pdelta_fallback_columns'] > 0
    for c, combo in enumerate(m.load_combos):
        assert rel_err(m._D[combo].reshape(-1), Do[c].reshape(-1)) < TOL_PDELTA, combo
        assert rel_err(m._Rxn[combo].reshape(-1), Ro[c].reshape(-1)) < TOL_PDELTA, combo


def test_row_range_kernels_for_ddpcg(eng):
    """The device-pointer building blocks of the domain-decomposed PCG (strip SpMM, node-range block-Jacobi, rhs /
    solution hand-over) on one GPU: two strips processed one after the other must equal the whole-matrix result,
    and a single-rank ddpcg.solve must reproduce the direct solver."""
    import torch
    from pynite_b200 import ddpcg
    A, ref, _ = load_fixture('mat_foundation')
    _upload(eng, A)
    eng.symbolic(0)
    eng.assemble_ke()
    D_direct, _, _ = eng.solve_linear(eng.make_opts(), reactions=False)
    ops = ddpcg.EngineOps(eng, A.n_nodes)
    try:
        n1 = ops.n1
        K11 = sps.csr_matrix((ref['K11_data'], ref['K11_indices'], ref['K11_indptr']), shape=(n1, n1))
        rng = np.random.default_rng(9)
        Xh = rng.standard_normal((n1, 5))
        X = torch.from_numpy(Xh).cuda()
        Y = torch.zeros_like(X)
        cut_node = A.n_nodes//2
        cut = int(ops.node_row_start[cut_node])
        ops.spmm_rows(0, cut, X, Y)
        ops.spmm_rows(cut, n1, X, Y)
        torch.cuda.synchronize()
        assert rel_err(Y.cpu().numpy(), K11 @ Xh) < 1e-12
        Z = torch.zeros_like(X)
        ops.precond_nodes(0, cut_node, X, Z)
        ops.precond_nodes(cut_node, A.n_nodes, X, Z)
        torch.cuda.synchronize()
        Zh = Z.cpu().numpy()
        Ks = ((K11 + K11.T)*0.5).tocsr()
        for nd in (0, cut_node, A.n_nodes - 1):
            a, b = int(ops.node_row_start[nd]), int(ops.node_row_start[nd + 1])
            if b > a:
                assert rel_err(Zh[a:b], np.linalg.solve(Ks[a:b, a:b].toarray(), Xh[a:b])) < 1e-10
        Xs, it, info = ddpcg.solve(ops, 0, A.factors.shape[0], tol=1e-12, max_iter=20000, check_every=10)
        ops.set_solution(0, Xs)
        for c in range(A.factors.shape[0]):
            assert rel_err(eng.displacements(c), D_direct[c]) < 1e-8
    finally:
        eng.set_stream(None)


# ---------------------------------------------------------------------------------------------------------------------
# Subtree-to-GPU direct solve (pn_set_distributed).  `distsolve.LocalGroup` runs `world` ranks as contexts of this
# process on the one GPU of the test box: the real kernels, plans and pack / unpack paths of every rank, with a
# device-to-device copy standing in for the NCCL all-gather.
# ---------------------------------------------------------------------------------------------------------------------
def _dist_case(name):
    from pynite_b200 import synthetic
    if name == 'plate':
        return synthetic.quad_plate_arrays(n=64, n_combos=5)
    if name == 'mat':
        return synthetic.quad_plate_arrays(n=48, n_combos=3, mat_springs=0.1)
    return synthetic.moment_frame_arrays(bays_x=6, bays_z=5, stories=9, n_combos=4)


@pytest.mark.parametrize('world', [2, 3, 4])
@pytest.mark.parametrize('case', ['plate', 'mat', 'frame'])
def test_subtree_distributed_solve_equals_single_gpu(case, world):
    """Every rank ends with all displacements and reactions, bit-identical to the single-context solve."""
    from pynite_b200 import distsolve
    from pynite_b200.engine import Engine
    A = _dist_case(case)
    e0 = Engine(0)
    try:
        _upload(e0, A)
        D0, R0, _ = e0.solve_linear(e0.make_opts())
    finally:
        e0.close()
    group = distsolve.LocalGroup(world)
    engines = [Engine(0) for _ in range(world)]

    def rank_fn(r):
        e = engines[r]
        group.attach(e, r)
        _upload(e, A)
        e.reset_stats()
        D, R, st = e.solve_linear(e.make_opts())
        D2, _, _ = e.solve_linear(e.make_opts(), reactions=False)       # second solve on the cached plan
        return D, R, D2, e.counters(), st.max_rel_residual
    try:
        out = group.run(rank_fn)
    finally:
        for e in engines:
            e.close()
    assert not group.errors
    for r, (D, R, D2, ctr, res) in enumerate(out):
        assert ctr['dist_cut_level'] >= 0, ctr
        assert ctr['dist_exchanges'] >= 6, ctr          # (matrices + 2 x (vectors, rows)) per solve with one refinement
        assert np.array_equal(D, D0), (r, rel_err(D, D0))
        assert np.array_equal(R, R0), r
        assert np.array_equal(D2, D0), r
        assert res < 1e-9


def test_subtree_distributed_pdelta_and_sweep_families():
    """P-Delta (stationary iteration on the distributed elastic factor) on two ranks, and the left-looking / per-step
    sweep families on the levels around the cut."""
    import os
    from pynite_b200 import distsolve, synthetic
    from pynite_b200.engine import Engine
    A = synthetic.moment_frame_arrays(bays_x=4, bays_z=4, stories=8, n_combos=4)
    e0 = Engine(0)
    try:
        _upload(e0, A)
        D0, R0, _ = e0.solve_pdelta(e0.make_opts())
    finally:
        e0.close()
    for env in ({}, {'PN_NO_SMEM_SWEEPS': '1'}, {'PN_NO_SMEM_SWEEPS': '1', 'PN_NO_LL_SWEEPS': '1'}):
        os.environ.update(env)
        group = distsolve.LocalGroup(2)
        engines = [Engine(0) for _ in range(2)]

        def rank_fn(r):
            e = engines[r]
            group.attach(e, r)
            _upload(e, A)
            return e.solve_pdelta(e.make_opts())[:2]
        try:
            out = group.run(rank_fn)
        finally:
            for e in engines:
                e.close()
            for k in env:
                os.environ.pop(k, None)
        for D, R in out:
            assert rel_err(D, D0) < 1e-12 and rel_err(R, R0) < 1e-10, env


def test_tc_iteration_compression_only_mat():
    """a mat on compression-only Winkler springs (`MatFoundation.py:136`, direction '-') that lifts off along
    one edge.  `pn_tc_update` switches the support springs in kernels (no Python loop over nodes, no re-upload);
    displacements, reactions and the final spring states against the fixture the reference produced."""
    from pynite_b200 import FEModel3D
    A, ref, _ = load_fixture('mat_tc')
    m = models.mat_tc(FEModel3D)
    m.analyze()
    for c, combo in enumerate(m.load_combos):
        assert rel_err(m._D[combo].reshape(-1), ref['D'][c]) < TOL_D, combo
        assert rel_err(m._Rxn[combo].reshape(-1), ref['Rxn'][c]) < TOL_D, combo
    # spring states after the last combination, as the reference left them (fixture inputs = reference's final flags)
    got = np.array([[1 if nd.spring_DY[2] else 0] for nd in m.nodes.values()]).reshape(-1)
    assert np.array_equal(got, (A.nspring_flag[:, 1] >> 1) & 1)


def test_mat_foundation_builder_through_analyze():
    """BASELINE config 4's caller: `FEModel3D.add_mat_foundation` + `add_mat_pt_load` (pynite_b200/MatFoundation.py; the
    generated nodes, spring stiffnesses and loads are checked against the reference's generator in tests/test_mesh_cpu.py),
    generated by `analyze()` itself, compression-only springs switched on the device: displacements, reactions and
    `MatFoundation.soil_pressure` (`MatFoundation.py:138-223`) against the reference."""
    from pynite_b200 import FEModel3D
    A, ref, _ = load_fixture('mat_mesh_tc')
    m = models.mat_mesh_tc(FEModel3D)
    m.analyze()
    for c, combo in enumerate(m.load_combos):
        assert rel_err(m._D[combo].reshape(-1), ref['D'][c]) < TOL_D, combo
        assert rel_err(m._Rxn[combo].reshape(-1), ref['Rxn'][c]) < TOL_D, combo
    gold = ref['soil_pressure']
    mat = m.mats['MAT']
    for b, combo in enumerate(m.load_combos):
        for k, (x, z) in enumerate(models.MAT_PRESSURE_POINTS):
            p = mat.soil_pressure(x, z, combo)
            assert abs(p - gold[0, b, k]) <= 1e-9*np.abs(gold).max(), (combo, x, z, p, gold[0, b, k])
    assert (gold[0] == 0).any() and (gold[0] > 0).any()          # part of the mat lifted off, part bears


def test_tc_iteration_mat_config4_shape_properties():
    """BASELINE config 4's stretch: 60 x 60 quads on compression-only springs through `analyze()`: the converged state
    satisfies the reference's own convergence rule (a '-' spring is active iff its node moved down), vertical
    equilibrium holds with the active springs only."""
    from pynite_b200 import FEModel3D
    n = 60
    m = models.mat_tc(FEModel3D, nx=n, ny=n)
    m.analyze(max_iter=60)
    nodes = list(m.nodes.values())
    combo = 'D+C+U'
    dy = np.array([nd.DY[combo] for nd in nodes])
    # (the flags kept on the node objects are those of the LAST combination; the run keeps every combo's end state)
    nsf = m._run.end_flags[combo][0]
    active = (nsf[:, 1] & 2) != 0
    assert np.all(dy[active] <= 0) and np.all(dy[~active] >= 0)
    assert 0 < (~active).sum() < len(nodes)
    # vertical equilibrium with the spring state of this combination: applied load + active spring forces = 0
    eng = m._engine
    from pynite_b200 import analysis as an
    jj = an._make_resident(m, combo)
    F = (eng.load_vector(jj, 'P') - eng.load_vector(jj, 'FER')).reshape(-1, 6)
    ks = m._run.arrays.nspring_k[:, 1]
    spring_force = -(ks*dy)[active].sum()
    assert abs(F[:, 1].sum() + spring_force) < 1e-8*np.abs(F[:, 1]).sum()
    j = m._run.combo_index[combo]
    assert j == 1


def test_config3_through_the_public_api():
    """BASELINE config 3 at full size (251 001 nodes, 250 000 quads) built with `add_node / add_quad /
    add_quad_surface_pressure / add_node_load / add_load_combo` and solved with `analyze_linear` -- the drop-in path, not
    the vectorised array builder: same displacements as the array path bit for bit, host pre-pass time printed."""
    import time
    from pynite_b200 import FEModel3D, synthetic
    from pynite_b200.engine import Engine
    t0 = time.perf_counter()
    m = models.big_quad_plate(FEModel3D, n=500, n_combos=4)
    t1 = time.perf_counter()
    m.analyze_linear(check_stability=True)
    t2 = time.perf_counter()
    st = m.last_stats
    print(f'config 3 through FEModel3D: build {t1 - t0:.1f} s, analyze_linear {t2 - t1:.1f} s '
          f'(host prepare + flatten {st["host_prepare_s"]:.1f} s, upload + device {st["device_total_s"]:.2f} s)')
    assert m._run.arrays.n_dof == 1506006
    A = synthetic.quad_plate_arrays(n=500, n_combos=4)
    e = Engine(0)
    try:
        _upload(e, A)
        D, R, _ = e.solve_linear(e.make_opts())
    finally:
        e.close()
    for c, combo in enumerate(m.load_combos):
        assert np.array_equal(m._D[combo].reshape(-1), D[c]), combo
        assert np.array_equal(m._Rxn[combo].reshape(-1), R[c]), combo
    nd = m.nodes['N' + str(250*501 + 250 + 1)]    # centre node
    assert nd.DZ['C1'] == m._D['C1'][nd.ID*6 + 2, 0] and nd.DZ['C1'] != 0
    assert t2 - t0 < 120
