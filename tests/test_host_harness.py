"""CPU checks of the CUDA library's element arithmetic and of the direct solver's symbolic analysis.

The element-math headers (`pynite_b200/csrc/elem_*.cuh`) are plain host/device functions; compiled with g++
(tests/host_harness, -ffp-contract=off like nvcc -fmad=false) they are compared with the golden element
matrices the reference produced.  The multifrontal ordering / front maps (`pn_direct_symbolic.cpp`) are
exercised by a sequential host emulation of the GPU factorisation and compared with scipy.
Nothing here is a product path: the package never loads these libraries."""
import ctypes as ct
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

from oracle import model as om
from tests.util import FIXTURES, GENERAL_ORIENTATION, load_fixture, rel_err

HH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'host_harness')
F64P = ct.POINTER(ct.c_double)
I32P = ct.POINTER(ct.c_int32)
I64P = ct.POINTER(ct.c_int64)


@pytest.fixture(scope='module')
def libs():
    subprocess.run(['make', '-C', HH], check=True, capture_output=True)
    return ct.CDLL(os.path.join(HH, 'libhost_harness.so')), ct.CDLL(os.path.join(HH, 'libdirect_host.so'))


def _p(a):
    return a.ctypes.data_as(F64P)


@pytest.mark.parametrize('name', FIXTURES)
def test_line_elements(libs, name):
    hh, _ = libs
    A, ref, _ = load_fixture(name)
    out = np.empty(144)
    for m in range(len(A.mem_rot)):
        i, j = A.mem_ij[m]
        pi, pj = np.ascontiguousarray(A.xyz[i]), np.ascontiguousarray(A.xyz[j])
        pr = np.ascontiguousarray(A.mem_props[m])
        hh.hh_line_matrix(_p(pi), _p(pj), _p(pr), ct.c_double(A.mem_rot[m]), ct.c_uint(int(A.mem_rel[m])), 0,
                          ct.c_double(0.0), _p(out))
        assert rel_err(out.reshape(12, 12), ref['member_Ke'][m]) < 1e-12, m
        if name not in GENERAL_ORIENTATION:
            # exact-zero structure decides the COO pattern (FEModel3D.py:130)
            assert np.array_equal(out.reshape(12, 12) != 0.0, ref['member_Ke'][m] != 0.0), m
        if 'member_Kg' in ref:
            hh.hh_line_matrix(_p(pi), _p(pj), _p(pr), ct.c_double(A.mem_rot[m]), ct.c_uint(int(A.mem_rel[m])), 1,
                              ct.c_double(ref['member_P'][m]), _p(out))
            assert rel_err(out.reshape(12, 12), ref['member_Kg'][m]) < 1e-12, m
    for s in range(len(A.spring_ks)):
        i, j = A.spring_ij[s]
        pi, pj = np.ascontiguousarray(A.xyz[i]), np.ascontiguousarray(A.xyz[j])
        pr = np.array([A.spring_ks[s], 0, 0, 0, 0, 0.0])
        hh.hh_line_matrix(_p(pi), _p(pj), _p(pr), ct.c_double(0.0), ct.c_uint(0), 2, ct.c_double(0.0), _p(out))
        assert rel_err(out.reshape(12, 12), ref['spring_Ke'][s]) < 1e-12, s


@pytest.mark.parametrize('name', FIXTURES)
def test_shell_elements(libs, name):
    hh, _ = libs
    A, ref, _ = load_fixture(name)
    out = np.empty(576)
    for kind, conn, props, is_quad in (('quad', A.quad_ijmn, A.quad_props, 1), ('plate', A.plate_ijmn, A.plate_props, 0)):
        for e in range(len(conn)):
            pts = np.ascontiguousarray(A.xyz[conn[e]].reshape(-1))
            pr = np.ascontiguousarray(props[e])
            hh.hh_shell_matrix(_p(pts), _p(pr), is_quad, _p(out))
            K, R = out.reshape(24, 24), ref[kind + '_Ke'][e]
            assert rel_err(K, R) < 1e-12, (kind, e)
            if name not in GENERAL_ORIENTATION:
                # exact-zero structure decides the COO pattern (FEModel3D.py:130)
                assert np.array_equal(K != 0.0, R != 0.0), (kind, e)


@pytest.mark.parametrize('name', ['quad_mesh_xy', 'quad_mesh_yz_ortho', 'plate_mesh_xz', 'skew_shell', 'mat_foundation',
                                  'mixed_building'])
def test_shell_stress_math(libs, name):
    """`elem_stress.cuh` (what `k_shell_stress` evaluates per element and combo) against the values the reference's
    `Quad3D.shear/moment/membrane` and `Plate3D.shear/moment/membrane` returned, from the reference's displacements."""
    hh, _ = libs
    A, ref, _ = load_fixture(name)
    # (compared over all combos at once: a combo of pure uniform settlement has shears that are rounding noise)
    for kind, conn, props, is_quad in (('quad', A.quad_ijmn, A.quad_props, 1), ('plate', A.plate_ijmn, A.plate_props, 0)):
        if not len(conn):
            continue
        variants = ((1, 'quad_stress_local'), (0, 'quad_stress_global')) if is_quad else ((1, 'plate_stress'),)
        for local, key in variants:
            got = np.empty_like(ref[key])
            for c in range(A.factors.shape[0]):
                D = ref['D'][c]
                for e in range(len(conn)):
                    p4 = A.xyz[conn[e]]
                    if is_quad:
                        pts = np.array(om.STRESS_POINTS, dtype=float)
                    else:
                        w = np.linalg.norm(p4[0] - p4[1]); h = np.linalg.norm(p4[0] - p4[3])
                        pts = np.array(om.PLATE_FRACTIONS, dtype=float)*np.array([w, h])
                    pts = np.ascontiguousarray(pts)
                    D24 = np.ascontiguousarray(D.reshape(-1, 6)[conn[e]].reshape(-1))
                    out = np.empty((len(pts), 9))
                    hh.hh_shell_stress(_p(np.ascontiguousarray(p4.reshape(-1))), _p(np.ascontiguousarray(props[e])), is_quad,
                                       _p(D24), len(pts), _p(pts), local, _p(out))
                    got[c, e] = out
            for k0, k1 in ((0, 3), (3, 6), (6, 9)):
                assert rel_err(got[..., k0:k1], ref[key][..., k0:k1]) < 1e-9, (key, k0)


@pytest.mark.parametrize('name', ['portal_frame', 'continuous_beam', 'moment_frame'])
def test_member_fer(libs, name):
    hh, _ = libs
    A, ref, _ = load_fixture(name)
    out = np.empty(12)
    for c in range(A.factors.shape[0]):
        factors = om.combo_factors(A, c)
        for m in range(len(A.mem_rot)):
            sel = np.nonzero(A.ml_member == m)[0]
            if not len(sel):
                continue
            i, j = A.mem_ij[m]
            pi, pj = np.ascontiguousarray(A.xyz[i]), np.ascontiguousarray(A.xyz[j])
            kinds = np.ascontiguousarray(A.ml_kind[sel].astype(np.int32))
            params = np.ascontiguousarray(A.ml_params[sel])
            fac = np.array([factors.get(int(A.ml_case[k]), 0.0) for k in sel])
            hh.hh_member_fer(_p(pi), _p(pj), _p(np.ascontiguousarray(A.mem_props[m])), ct.c_double(A.mem_rot[m]),
                             ct.c_uint(int(A.mem_rel[m])), len(sel), kinds.ctypes.data_as(I32P), _p(params), _p(fac), _p(out))
            expect = om.member_fer_local(A, m, factors).reshape(-1)
            assert np.abs(out - expect).max() <= 1e-11*max(1.0, np.abs(expect).max()), (c, m)


@pytest.mark.parametrize('name', ['quad_mesh_xy', 'plate_mesh_xz', 'skew_shell', 'mat_foundation'])
def test_shell_fer(libs, name):
    hh, _ = libs
    A, ref, _ = load_fixture(name)
    out = np.empty(24)
    n = A.n_dof
    for c in range(A.factors.shape[0]):
        factors = om.combo_factors(A, c)
        FER = np.zeros(n)
        for conn, props, is_quad, el, cs, val in ((A.quad_ijmn, A.quad_props, 1, A.qp_elem, A.qp_case, A.qp_val),
                                                  (A.plate_ijmn, A.plate_props, 0, A.pp_elem, A.pp_case, A.pp_val)):
            for e in range(len(conn)):
                p = sum(factors.get(int(cs[k]), 0.0)*val[k] for k in np.nonzero(el == e)[0])
                pts = np.ascontiguousarray(A.xyz[conn[e]].reshape(-1))
                hh.hh_shell_fer(_p(pts), _p(np.ascontiguousarray(props[e])), is_quad, ct.c_double(p), _p(out))
                dofs = (6*conn[e][:, None] + np.arange(6)[None, :]).reshape(-1)
                np.add.at(FER, dofs, out)
        assert rel_err(FER, ref['FER'][c]) < 1e-12


def _vertex_graph(K11, A, d1):
    """Vertex graph of free DOFs grouped by node, as pn_direct_symbolic.cpp::build_vertex_graph builds it."""
    node_of = np.asarray(d1)//6
    nodes, vstart = np.unique(node_of, return_index=True)
    vstart = np.append(vstart, len(d1)).astype(np.int32)
    vid = -np.ones(A.n_nodes, dtype=np.int64)
    vid[nodes] = np.arange(len(nodes))
    C = K11.tocoo()
    va, vb = vid[node_of[C.row]], vid[node_of[C.col]]
    keep = va != vb
    G = sps.coo_matrix((np.ones(keep.sum()), (va[keep], vb[keep])), shape=(len(nodes), len(nodes))).tocsr()
    G = ((G + G.T) > 0).astype(np.int8).tocsr()
    G.sort_indices()
    coords = np.ascontiguousarray(A.xyz[nodes].reshape(-1))
    return vstart, coords, G.indptr.astype(np.int64), G.indices.astype(np.int32)


@pytest.mark.parametrize('name,leaf', [('quad_mesh_xy', 2), ('moment_frame', 3), ('mat_foundation', 4), ('mixed_building', 2),
                                       ('plate_mesh_xz', 1), ('pdelta_frame', 5)])
def test_direct_symbolic_host(libs, name, leaf):
    _, dh = libs
    A, ref, _ = load_fixture(name)
    d1, d2, _ = om.partition_D(A)
    K11 = sps.csr_matrix((ref['K11_data'], ref['K11_indices'], ref['K11_indptr']), shape=(len(d1), len(d1)))
    vstart, coords, adj_ptr, adj = _vertex_graph(K11, A, d1)
    s = 3
    rng = np.random.default_rng(0)
    B = np.ascontiguousarray(rng.standard_normal((len(d1), s)))
    X = np.empty_like(B)
    stats = np.zeros(8)
    rc = dh.hh_direct_solve(ct.c_int64(len(d1)), K11.indptr.ctypes.data_as(I32P), K11.indices.ctypes.data_as(I32P),
                            _p(np.ascontiguousarray(K11.data)), ct.c_int32(len(vstart) - 1),
                            vstart.ctypes.data_as(I32P), _p(coords), adj_ptr.ctypes.data_as(I64P),
                            adj.ctypes.data_as(I32P), leaf, s, _p(B), _p(X), _p(stats), None, 1, None)
    assert rc == 0
    Xs = spla.splu(K11.tocsc()).solve(B)
    assert rel_err(X, Xs) < 1e-8
    assert stats[2] >= 1 and stats[4] >= 1


@pytest.mark.parametrize('name,leaf,expect', [('quad_mesh_xy', 1, 3), ('quad_mesh_yz_ortho', 1, 3), ('plate_mesh_xz', 2, 3),
                                              ('mat_foundation', 1, 3), ('skew_shell', 2, 1), ('mixed_building', 2, 1)])
def test_direct_symbolic_dof_classes(libs, name, leaf, expect):
    """DOF-type classes (pn_direct_symbolic.h, DofClasses): a flat shell model in an axis plane is ordered as three
    independent problems (membrane, bending, drilling) on one node dissection; models with line elements or shells in
    general position keep one class.  The library's own predictor, the host multifrontal emulation on the resulting
    tree against SuperLU, and -- the property the device verifies with k_entry_maps -- no stored K11 entry joins two
    classes."""
    _, dh = libs
    A, ref, _ = load_fixture(name)
    d1, d2, _ = om.partition_D(A)
    n1 = len(d1)
    K11 = sps.csr_matrix((ref['K11_data'], ref['K11_indices'], ref['K11_indptr']), shape=(n1, n1))
    n_line = len(A.spring_ij.reshape(-1, 2)) + len(A.mem_ij.reshape(-1, 2))
    quads = np.ascontiguousarray(A.quad_ijmn.reshape(-1, 4).astype(np.int32))
    plates = np.ascontiguousarray(A.plate_ijmn.reshape(-1, 4).astype(np.int32))
    xyz = np.ascontiguousarray(A.xyz, dtype=np.float64)
    of_type = np.zeros(6, dtype=np.uint8)
    diagonal = np.zeros(6, dtype=np.uint8)
    U8P = ct.POINTER(ct.c_uint8)
    ncls = dh.hh_predict_classes(_p(xyz), ct.c_int64(len(quads)), quads.ctypes.data_as(I32P), ct.c_int64(len(plates)),
                                 plates.ctypes.data_as(I32P), ct.c_int64(n_line), of_type.ctypes.data_as(U8P),
                                 diagonal.ctypes.data_as(U8P))
    assert ncls == expect
    if ncls == 1:
        return
    assert diagonal[:ncls].sum() == 1
    # no stored entry joins two classes; a diagonal class has no off-diagonal entry at all
    row_class = np.ascontiguousarray(of_type[np.asarray(d1) % 6])
    C = K11.tocoo()
    assert np.array_equal(row_class[C.row], row_class[C.col])
    off = C.row != C.col
    assert not diagonal[row_class[C.row[off]]].any()
    vstart, coords, adj_ptr, adj = _vertex_graph(K11, A, d1)
    s = 3
    rng = np.random.default_rng(1)
    B = np.ascontiguousarray(rng.standard_normal((n1, s)))
    X = np.empty_like(B)
    out = []
    for args in ((row_class.ctypes.data_as(U8P), ncls, diagonal.ctypes.data_as(U8P)), (None, 1, None)):
        stats = np.zeros(8)
        rc = dh.hh_direct_solve(ct.c_int64(n1), K11.indptr.ctypes.data_as(I32P), K11.indices.ctypes.data_as(I32P),
                                _p(np.ascontiguousarray(K11.data)), ct.c_int32(len(vstart) - 1), vstart.ctypes.data_as(I32P),
                                _p(coords), adj_ptr.ctypes.data_as(I64P), adj.ctypes.data_as(I32P), leaf, s, _p(B), _p(X),
                                _p(stats), *args)
        assert rc == 0
        assert rel_err(X, spla.splu(K11.tocsc()).solve(B)) < 1e-8
        out.append(stats.copy())
    # fewer factor entries and flops than the node-based ordering of the same matrix
    assert out[0][0] < out[1][0] and out[0][1] < out[1][1]


@pytest.mark.parametrize('name', ['torque_beam', 'spring_elements', 'logan_12_1_plate', 'nodal_springs'])
def test_direct_symbolic_host_on_tiny_systems(libs, name):
    """The known-answer models of the reference's tests (tests/known_answers.py) have as few as ONE free degree of freedom:
    the ordering / front tables must cope with a vertex graph of one vertex and no edge, and with every leaf size."""
    from pynite_b200 import FEModel3D, analysis
    from pynite_b200.flatten import flatten_model
    from tests.known_answers import DIAGRAM_KNOWN_ANSWERS, KNOWN_ANSWERS
    _, dh = libs
    m, _, _ = {**KNOWN_ANSWERS, **DIAGRAM_KNOWN_ANSWERS}[name](FEModel3D)
    analysis.prepare_model(m)
    combos = analysis.identify_combos(m)
    A = flatten_model(m, combos, combos[0].name)
    d1, d2, _ = om.partition_D(A)
    K11 = sps.csr_matrix(om.ke_csr(A)[d1, :][:, d1])
    K11.sort_indices()
    assert 1 <= len(d1) <= 6
    vstart, coords, adj_ptr, adj = _vertex_graph(K11, A, d1)
    B = np.ascontiguousarray(np.random.default_rng(1).standard_normal((len(d1), 2)))
    Xs = spla.splu(K11.tocsc()).solve(B)
    for leaf in (1, 2, 8):
        X = np.empty_like(B)
        stats = np.zeros(8)
        rc = dh.hh_direct_solve(ct.c_int64(len(d1)), K11.indptr.astype(np.int32).ctypes.data_as(I32P),
                                K11.indices.astype(np.int32).ctypes.data_as(I32P), _p(np.ascontiguousarray(K11.data)),
                                ct.c_int32(len(vstart) - 1), vstart.ctypes.data_as(I32P), _p(coords),
                                adj_ptr.ctypes.data_as(I64P), adj.ctypes.data_as(I32P), leaf, 2, _p(B), _p(X), _p(stats),
                                None, 1, None)
        assert rc == 0
        assert rel_err(X, Xs) < 1e-9, (name, leaf)
