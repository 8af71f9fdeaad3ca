"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the golden fixtures the
reference produced and against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): CSR pattern, COO index map and DOF partition bit-exact; element and
assembled values <= 1e-12 relative; displacements and reactions <= 1e-9 relative (1e-6 for P-Delta)."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import model as om
from tests import models
from tests.util import FIXTURES, GENERAL_ORIENTATION, compare_general_pattern, load_fixture, rel_err

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
@pytest.mark.parametrize('name', [f for f in FIXTURES if not f.startswith('pdelta') and f != 'braced_tc'])
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
    # element end forces (Member3D.F/f, Quad3D.F, Plate3D.F, Spring3D.F)
    for c in range(A.factors.shape[0]):
        for kind in ('spring', 'member', 'quad', 'plate'):
            F = eng.element_forces(kind, c)
            if F.size:
                assert rel_err(F, ref[kind + '_F'][c]) < 1e-8, (kind, c)
        f = eng.element_forces('member', c, local=True)
        if f.size:
            assert rel_err(f, ref['member_f'][c]) < 1e-8


@pytest.mark.parametrize('name', ['pdelta_column', 'pdelta_frame'])
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
    if subs and mode != 'tc':
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


def test_load_stepping_is_linear_without_tc_features():
    """`analyze(num_steps=k)` on a model with a directional support spring that stays active: the summed load steps
    equal the one-step solution (`Analysis._first_order` :262-305)."""
    from pynite_b200 import FEModel3D
    m1 = models.braced_tc(FEModel3D)
    m1.analyze()
    m2 = models.braced_tc(FEModel3D)
    m2.analyze(num_steps=1, max_iter=30)
    for combo in m1.load_combos:
        assert np.array_equal(m1._D[combo], m2._D[combo])


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
