"""Vectorised builders of `flatten.ModelArrays` for the large synthetic BASELINE configurations.

Building config 3 (251 001 nodes, 250 000 quads) through `add_node/add_quad` and `flatten_model` costs
minutes of pure-Python object handling; these functions produce, with numpy only, exactly the arrays
`flatten_model(tests.models.big_quad_plate(FEModel3D, ...))` would produce (checked at small sizes in
tests/test_synthetic.py), so bench.py and the large-size tests can feed the C ABI directly.
Node numbering is row-major and quad connectivity CCW i-j-m-n, as `RectangleMesh` emits them
(`Pynite/Mesh.py:867-943`).
"""
from __future__ import annotations

import numpy as np

from pynite_b200.flatten import ModelArrays


def _empty_common(A, N):
    A.n_nodes = N
    A.enf_mask = np.zeros(N, dtype=np.uint8)
    A.enf_val = np.zeros((N, 6))
    A.nspring_k = np.zeros((N, 6))
    A.nspring_flag = np.zeros((N, 6), dtype=np.uint8)
    A.spring_ij = np.zeros((0, 2), dtype=np.int32)
    A.spring_ks = np.zeros(0)
    A.spring_active = np.zeros(0, dtype=np.uint8)
    A.mem_ij = np.zeros((0, 2), dtype=np.int32)
    A.mem_props = np.zeros((0, 6))
    A.mem_rot = np.zeros(0)
    A.mem_rel = np.zeros(0, dtype=np.uint16)
    A.mem_active = np.zeros(0, dtype=np.uint8)
    A.quad_ijmn = np.zeros((0, 4), dtype=np.int32)
    A.quad_props = np.zeros((0, 5))
    A.plate_ijmn = np.zeros((0, 4), dtype=np.int32)
    A.plate_props = np.zeros((0, 5))
    for f in ('nl_dof', 'nl_case', 'ml_member', 'ml_case', 'ml_kind', 'qp_elem', 'qp_case', 'pp_elem', 'pp_case'):
        setattr(A, f, np.zeros(0, dtype=np.int32))
    for f in ('nl_val', 'qp_val', 'pp_val'):
        setattr(A, f, np.zeros(0))
    A.ml_params = np.zeros((0, 4))
    A.sub_owner = np.zeros(0, dtype=np.int32)


def quad_plate_arrays(n=500, n_combos=64, a=12.0, t=8.0, E=3605.0, nu=0.17, seed=20261019, mat_springs=None, jitter=0.0):
    """BASELINE config 3: n x n DKMQ quads in the XY plane, perimeter pinned.

    `mat_springs` = ks (kci) switches to the config-4 shape instead: the mesh lies in the XZ plane, every
    node carries a DY Winkler spring ks*tributary area (`MatFoundation.py:121-136`, linear springs), and
    only two corner nodes are restrained in-plane.

    `jitter` > 0 moves every node in the plane of the mesh by U(-jitter, jitter) in both directions (own random
    stream): no two quads are bit-identical any more, which is what the general-geometry assembly path is measured
    on (the generated meshes of the BASELINE configs collapse to a handful of element classes)."""
    A = ModelArrays()
    N = (n + 1)*(n + 1)
    _empty_common(A, N)
    c, r = np.meshgrid(np.arange(n + 1), np.arange(n + 1))
    c, r = c.reshape(-1), r.reshape(-1)
    xyz = np.zeros((N, 3))
    xyz[:, 0] = c*a
    xyz[:, 2 if mat_springs is not None else 1] = r*a
    if jitter:
        jr = np.random.default_rng(seed + 7919)
        xyz[:, 0] += jr.uniform(-jitter, jitter, N)
        xyz[:, 2 if mat_springs is not None else 1] += jr.uniform(-jitter, jitter, N)
    A.xyz = xyz
    support = np.zeros(N, dtype=np.uint8)
    if mat_springs is None:
        support[(c == 0) | (c == n) | (r == 0) | (r == n)] = 0b000111
    else:
        support[0] = 0b000101                  # DX, DZ
        support[n] = 0b000100                  # DZ
        trib = a*a*np.where((r == 0) | (r == n), 0.5, 1.0)*np.where((c == 0) | (c == n), 0.5, 1.0)
        A.nspring_k[:, 1] = mat_springs*trib
        A.nspring_flag[:, 1] = 3
    A.support = support
    qc, qr = np.meshgrid(np.arange(n), np.arange(n))
    qc, qr = qc.reshape(-1), qr.reshape(-1)
    i = qr*(n + 1) + qc
    A.quad_ijmn = np.ascontiguousarray(np.stack([i, i + 1, i + n + 2, i + n + 1], axis=1).astype(np.int32))
    Q = n*n
    A.quad_props = np.ascontiguousarray(np.tile(np.array([t, E, nu, 1.0, 1.0]), (Q, 1)))
    # pressures: D on every quad, L on the checkerboard, listed quad by quad like flatten_model does
    even = (qr + qc) % 2 == 0
    cnt = 1 + even.astype(np.int64)
    start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    total = int(cnt.sum())
    qp_elem = np.repeat(np.arange(Q), cnt).astype(np.int32)
    qp_case = np.zeros(total, dtype=np.int32)
    qp_val = np.full(total, 0.001)
    lpos = start[even] + 1
    qp_case[lpos] = 1
    qp_val[lpos] = 0.0005
    A.qp_elem, A.qp_case, A.qp_val = qp_elem, qp_case, qp_val
    # nodal loads: same RNG call sequence as tests.models.big_quad_plate, then model (node) order
    rng = np.random.default_rng(seed)
    n_pt = min(100, max(1, (n - 1)*(n - 1)))
    normal = 1 if mat_springs is not None else 2
    recs = []
    for p in range(6):
        for _ in range(n_pt):
            cc, rr = int(rng.integers(1, n)), int(rng.integers(1, n))
            recs.append((rr*(n + 1) + cc, 2 + p, float(rng.uniform(-10, 0))))
    node = np.array([x[0] for x in recs])
    order = np.argsort(node, kind='stable')
    A.nl_dof = (node[order]*6 + normal).astype(np.int32)
    A.nl_case = np.array([recs[k][1] for k in order], dtype=np.int32)
    A.nl_val = np.array([recs[k][2] for k in order])
    A.cases = ['D', 'L'] + ['P' + str(p + 1) for p in range(6)]
    A.factors = np.ascontiguousarray(np.round(rng.uniform(0.5, 1.6, size=(n_combos, 8)), 2))
    A.combo_names = ['C' + str(k + 1) for k in range(n_combos)]
    return A


def moment_frame_arrays(bays_x=20, bays_z=20, stories=30, bay=360.0, story=156.0, n_combos=16, seed=7):
    """BASELINE config 2: 3-D moment frame, fixed bases, D/L line loads on beams, WX/WZ nodal
    loads; equals `flatten_model(tests.models.moment_frame(FEModel3D, ...))` for `interior=0`."""
    A = ModelArrays()
    nx, nz = bays_x + 1, bays_z + 1
    N = nx*nz*(stories + 1)
    _empty_common(A, N)
    s, iz, ix = np.meshgrid(np.arange(stories + 1), np.arange(nz), np.arange(nx), indexing='ij')
    s, iz, ix = s.reshape(-1), iz.reshape(-1), ix.reshape(-1)
    A.xyz = np.ascontiguousarray(np.stack([ix*bay, s*story, iz*bay], axis=1).astype(float))
    support = np.zeros(N, dtype=np.uint8)
    support[s == 0] = 0b111111
    A.support = support

    def nid(ix_, s_, iz_):
        return (s_*nz + iz_)*nx + ix_
    ij, props, loaded = [], [], []
    col = (29000.0, 11200.0, 50.0, 800.0, 2000.0, 30.0)
    beam = (29000.0, 11200.0, 20.0, 100.0, 1200.0, 3.0)
    for s_ in range(stories):
        for iz_ in range(nz):
            for ix_ in range(nx):
                ij.append((nid(ix_, s_, iz_), nid(ix_, s_ + 1, iz_)))
                props.append(col)
                loaded.append(False)
    for s_ in range(1, stories + 1):
        for iz_ in range(nz):
            for ix_ in range(bays_x):
                ij.append((nid(ix_, s_, iz_), nid(ix_ + 1, s_, iz_)))
                props.append(beam)
                loaded.append(True)
        for iz_ in range(bays_z):
            for ix_ in range(nx):
                ij.append((nid(ix_, s_, iz_), nid(ix_, s_, iz_ + 1)))
                props.append(beam)
                loaded.append(True)
    M = len(ij)
    A.mem_ij = np.ascontiguousarray(np.array(ij, dtype=np.int32).reshape(M, 2))
    A.mem_props = np.ascontiguousarray(np.array(props, dtype=float).reshape(M, 6))
    A.mem_rot = np.zeros(M)
    A.mem_rel = np.zeros(M, dtype=np.uint16)
    A.mem_active = np.ones(M, dtype=np.uint8)
    A.sub_owner = np.arange(M, dtype=np.int32)
    lm = np.nonzero(np.array(loaded))[0]
    A.ml_member = np.repeat(lm, 2).astype(np.int32)
    A.ml_case = np.tile(np.array([0, 1], dtype=np.int32), len(lm))
    A.ml_kind = np.full(2*len(lm), 13, dtype=np.int32)          # distributed local Fy
    par = np.zeros((2*len(lm), 4))
    par[0::2, 0] = par[0::2, 1] = -0.1
    par[1::2, 0] = par[1::2, 1] = -0.06
    par[:, 3] = bay
    A.ml_params = par
    # nodal loads in model (node) order; per node the builder appends FX (if ix == 0) before FZ (if iz == 0)
    dof, case, val = [], [], []
    for s_ in range(1, stories + 1):
        for iz_ in range(nz):
            for ix_ in range(nx):
                if ix_ == 0:
                    dof.append(nid(ix_, s_, iz_)*6 + 0); case.append(2); val.append(2.0*s_)
                if iz_ == 0:
                    dof.append(nid(ix_, s_, iz_)*6 + 2); case.append(3); val.append(2.0*s_)
    A.nl_dof = np.array(dof, dtype=np.int32)
    A.nl_case = np.array(case, dtype=np.int32)
    A.nl_val = np.array(val, dtype=float)
    A.cases = ['D', 'L', 'WX', 'WZ']
    table = [(1.4, 0, 0, 0), (1.2, 1.6, 0, 0), (1.2, 1.0, 1.0, 0), (1.2, 1.0, -1.0, 0),
             (1.2, 1.0, 0, 1.0), (1.2, 1.0, 0, -1.0), (0.9, 0, 1.0, 0), (0.9, 0, -1.0, 0),
             (0.9, 0, 0, 1.0), (0.9, 0, 0, -1.0), (1.2, 0.5, 0.5, 0.5), (1.2, 0.5, -0.5, 0.5),
             (1.2, 0.5, 0.5, -0.5), (1.2, 0.5, -0.5, -0.5), (1.0, 1.0, 0.7, 0.3), (1.0, 1.0, 0.3, 0.7)]
    rng = np.random.default_rng(seed)
    rows = []
    for c_ in range(n_combos):
        if c_ < len(table):
            f = table[c_]
        else:
            f = tuple(np.round(rng.uniform(-1.0, 1.6, size=4), 2))
            f = (abs(f[0]) + 0.5, abs(f[1]), f[2], f[3])
        rows.append([float(v) for v in f])
    A.factors = np.ascontiguousarray(np.array(rows).reshape(n_combos, 4))
    A.combo_names = ['C' + str(k + 1) for k in range(n_combos)]
    return A
