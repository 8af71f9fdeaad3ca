"""Model-level restatement of the reference path on flattened arrays.  TEST INFRASTRUCTURE ONLY.

Input is any object with the attributes of `pynite_b200.flatten.ModelArrays` (plain numpy arrays).
The functions mirror, phase by phase:

  assemble_coo      FEModel3D.Ke + _append_sparse_block      FEModel3D.py:1551-1791, 99-151
  ke_csr            .tocsr()                                  FEModel3D.py:2279
  partition_D       Analysis._partition_D                     Analysis.py:1412-1505
  load_vectors      FEModel3D.P / FEModel3D.FER               FEModel3D.py:2136-2235
  solve_linear      analyze_linear combo loop + spsolve       FEModel3D.py:2287-2317, Analysis.py:176-241
  reactions         Analysis._calc_reactions                  Analysis.py:1059-1313
  solve_pdelta      Analysis._PDelta (two-step)               Analysis.py:329-471, FEModel3D.py:1802-1895
  shell_stresses    Quad3D / Plate3D shear, moment, membrane  Quad3D.py:1017-1239, Plate3D.py:590-692

Like the reference, everything is a Python loop over elements calling small numpy routines, and the
sparse algebra is scipy's (coo_matrix, tocsr, fancy-index partition, SuperLU spsolve).

Memoisation (`MEMOISE`, on by default): the reference's element routines read node coordinates only through
the difference vectors they form first (`Quad3D.py:107-141, 884-936`, `Plate3D.py:451-487`, `Member3D.py:933-1027`
plus the three `isclose` branch tests on absolute coordinates).  The per-element routines below are therefore pure
functions of (difference vectors, branch flags, properties); their results are cached under exactly those bits, so a
generated mesh with bit-identical differences costs one evaluation per distinct element instead of one per element.
The cached value IS the output of the restated routine on the first element with that key; nothing is approximated
(`tests/test_oracle_golden.py::test_memoised_oracle_is_bit_identical`).  This is what lets the parity tests compare
the CUDA path with the oracle at 150x150 ... 256x256 quads within seconds.
"""
import numpy as np
import scipy.sparse as sps
from scipy.sparse.linalg import spsolve

from oracle import elements as el

_PT_NAMES = ['Fx', 'Fy', 'Fz', 'Mx', 'My', 'Mz', 'FX', 'FY', 'FZ', 'MX', 'MY', 'MZ']
_DIST_NAMES = ['Fx', 'Fy', 'Fz', 'FX', 'FY', 'FZ']


MEMOISE = True
_MEMO = {}


def _cached(key, fn):
    if not MEMOISE:
        return fn()
    hit = _MEMO.get(key)
    if hit is None:
        if len(_MEMO) > 200000:
            _MEMO.clear()
        hit = _MEMO[key] = fn()
    return hit


def _shell_key(A, ijmn, props):
    p = A.xyz[ijmn]
    return ((p[1] - p[0]).tobytes(), (p[2] - p[0]).tobytes(), (p[3] - p[0]).tobytes(), props.tobytes())


def _line_key(A, i, j, rot=0.0):
    from math import isclose
    pi, pj = A.xyz[i], A.xyz[j]
    return ((pj - pi).tobytes(), (pi - pj).tobytes(), isclose(pi[0], pj[0]), isclose(pi[1], pj[1]), isclose(pi[2], pj[2]),
            float(rot))


def _dofs(*node_ids):
    """FEModel3D._build_dof_vector (`FEModel3D.py:66-94`)."""
    return np.concatenate([np.arange(6, dtype=np.int64) + 6*int(n) for n in node_ids])


# ---- per-element global matrices ---------------------------------------------------------------
def spring_Ke(A, s):
    i, j = A.spring_ij[s]
    return _cached(('sK', _line_key(A, i, j), float(A.spring_ks[s])), lambda: _spring_Ke(A, s))


def _spring_Ke(A, s):
    i, j = A.spring_ij[s]
    R = el.line_dircos(A.xyz[i], A.xyz[j])
    k = np.zeros((12, 12))
    ks = A.spring_ks[s]
    k[0, 0] = k[6, 6] = ks
    k[0, 6] = k[6, 0] = -ks
    T = el.block_diag_T(R, 4)
    return el.to_global(T, k), T, k


def _member_key(A, m):
    i, j = A.mem_ij[m]
    return (_line_key(A, i, j, A.mem_rot[m]), A.mem_props[m].tobytes(), int(A.mem_rel[m]))


def member_geometry(A, m):
    return _cached(('mG', _member_key(A, m)), lambda: _member_geometry(A, m))


def _member_geometry(A, m):
    i, j = A.mem_ij[m]
    L = float(((A.xyz[i] - A.xyz[j])**2).sum()**0.5)
    R = el.line_dircos(A.xyz[i], A.xyz[j], float(A.mem_rot[m]))
    rel = [bool((int(A.mem_rel[m]) >> k) & 1) for k in range(12)]
    return L, R, rel


def member_Ke(A, m):
    return _cached(('mK', _member_key(A, m)), lambda: _member_Ke(A, m))


def _member_Ke(A, m):
    L, R, rel = member_geometry(A, m)
    E, G, Ar, Iy, Iz, J = A.mem_props[m]
    k = el.member_ke(E, G, Ar, Iy, Iz, J, L, rel)
    T = el.block_diag_T(R, 4)
    return el.to_global(T, k), T, k


def member_Kg(A, m, P):
    L, R, rel = member_geometry(A, m)
    E, G, Ar, Iy, Iz, J = A.mem_props[m]
    k = el.member_kg(P, Ar, Iy, Iz, L, rel)
    T = el.block_diag_T(R, 4)
    return el.to_global(T, k), T, k


def quad_Ke(A, q):
    return _cached(('qK', _shell_key(A, A.quad_ijmn[q], A.quad_props[q])), lambda: _quad_Ke(A, q))


def _quad_Ke(A, q):
    n = A.quad_ijmn[q]
    p = [A.xyz[k] for k in n]
    t, E, nu, kx, ky = A.quad_props[q]
    k = el.quad_ke(p[0], p[1], p[2], p[3], t, E, nu, kx, ky)
    T = el.block_diag_T(el.shell_dircos(p[0], p[1], p[3]), 8)
    return el.to_global(T, k), T, k


def plate_Ke(A, e):
    return _cached(('pK', _shell_key(A, A.plate_ijmn[e], A.plate_props[e])), lambda: _plate_Ke(A, e))


def _plate_Ke(A, e):
    n = A.plate_ijmn[e]
    p = [A.xyz[k] for k in n]
    t, E, nu, kx, ky = A.plate_props[e]
    w = float(((p[0] - p[1])**2).sum()**0.5)
    h = float(((p[0] - p[3])**2).sum()**0.5)
    k = el.plate_ke(w, h, t, E, nu, kx, ky)
    T = el.block_diag_T(el.shell_dircos(p[0], p[1], p[3]), 8)
    return el.to_global(T, k), T, k


# ---- assembly ----------------------------------------------------------------------------------
def _append_block(dofs, block, rows, cols, data, keep_zeros=False):
    """FEModel3D._append_sparse_block (`FEModel3D.py:99-151`): drop exact zeros."""
    flat = np.asarray(block, dtype=float).reshape(-1)
    mask = np.ones(flat.size, dtype=bool) if keep_zeros else (flat != 0.0)
    if not mask.any():
        return
    rows.append(np.repeat(dofs, dofs.size)[mask])
    cols.append(np.tile(dofs, dofs.size)[mask])
    data.append(flat[mask])


def assemble_coo(A):
    """COO triplets of Ke in the reference's emission order (`FEModel3D.py:1587-1791`)."""
    rows, cols, data = [], [], []
    for n in range(A.n_nodes):
        for k in range(6):
            f = int(A.nspring_flag[n, k])
            if (f & 1) and (f & 2):
                rows.append(np.array([6*n + k], dtype=np.int64))
                cols.append(np.array([6*n + k], dtype=np.int64))
                data.append(np.array([A.nspring_k[n, k]], dtype=float))
    for s in range(len(A.spring_ks)):
        if A.spring_active[s]:
            _append_block(_dofs(*A.spring_ij[s]), spring_Ke(A, s)[0], rows, cols, data)
    for m in range(len(A.mem_rot)):
        if A.mem_active[m]:
            _append_block(_dofs(*A.mem_ij[m]), member_Ke(A, m)[0], rows, cols, data)
    for q in range(len(A.quad_ijmn)):
        _append_block(_dofs(*A.quad_ijmn[q]), quad_Ke(A, q)[0], rows, cols, data)
    for p in range(len(A.plate_ijmn)):
        _append_block(_dofs(*A.plate_ijmn[p]), plate_Ke(A, p)[0], rows, cols, data)
    if rows:
        return np.concatenate(rows), np.concatenate(cols), np.concatenate(data)
    return np.array([], dtype=np.int64), np.array([], dtype=np.int64), np.array([], dtype=float)


def ke_coo(A):
    r, c, d = assemble_coo(A)
    return sps.coo_matrix((d, (r, c)), shape=(6*A.n_nodes, 6*A.n_nodes))


def ke_csr(A):
    return ke_coo(A).tocsr()


def assemble_kg_coo(A, P):
    """Kg triplets: all 144 entries of every active member, no zero filter (`FEModel3D.py:1826-1892`)."""
    rows, cols, data = [], [], []
    for m in range(len(A.mem_rot)):
        if A.mem_active[m]:
            _append_block(_dofs(*A.mem_ij[m]), member_Kg(A, m, P[m])[0], rows, cols, data, keep_zeros=True)
    n = 6*A.n_nodes
    if rows:
        return sps.coo_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    return sps.coo_matrix((n, n))


def partition_D(A):
    """Analysis._partition_D (`Analysis.py:1412-1505`), one node and DOF at a time."""
    d1, d2, v2 = [], [], []
    for n in range(A.n_nodes):
        for k in range(6):
            sup = (int(A.support[n]) >> k) & 1
            enf = (int(A.enf_mask[n]) >> k) & 1
            if not sup and not enf:
                d1.append(6*n + k)
            elif enf:
                d2.append(6*n + k)
                v2.append(A.enf_val[n, k])
            else:
                d2.append(6*n + k)
                v2.append(0.0)
    return d1, d2, np.array(v2, ndmin=2).T.reshape(len(v2), 1)


def partition_matrix(K, d1, d2):
    """Analysis._partition (`Analysis.py:1522-1540`)."""
    K = K.tocsr()
    return K[d1, :][:, d1], K[d1, :][:, d2], K[d2, :][:, d1], K[d2, :][:, d2]


# ---- loads -------------------------------------------------------------------------------------
def _index_by(A, field):
    """{element: [positions]} for a load table, built once per array object (keeps the element loops linear)."""
    cache = A.__dict__.setdefault('_oracle_index', {})
    arr = getattr(A, field)
    hit = cache.get(field)
    if hit is None or hit[0] is not arr:
        table = {}
        for k, e in enumerate(arr.tolist()):
            table.setdefault(e, []).append(k)
        cache[field] = hit = (arr, table)
    return hit[1]


def _member_loads(A, m):
    pts, dists = [], []
    for k in _index_by(A, 'ml_member').get(m, ()):
        kind = int(A.ml_kind[k])
        p = A.ml_params[k]
        if kind < 12:
            pts.append((_PT_NAMES[kind], p[0], p[1], int(A.ml_case[k])))
        else:
            dists.append((_DIST_NAMES[kind - 12], p[0], p[1], p[2], p[3], int(A.ml_case[k])))
    return pts, dists


def member_fer_local(A, m, factors):
    """Condensed local FER of one member for a {case index: factor} dict (`Member3D.py:742-772`)."""
    L, R, rel = member_geometry(A, m)
    pts, dists = _member_loads(A, m)
    fu = el.member_fer_unc(pts, dists, factors, R, L)
    E, G, Ar, Iy, Iz, J = A.mem_props[m]
    return el.condense_vector(el.member_ke_unc(E, G, Ar, Iy, Iz, J, L), fu, rel)


def _pressure(A, field, case_ids, vals, e, factors):
    p = 0
    rows = _index_by(A, field).get(e, ())
    for case, factor in factors.items():
        for k in rows:
            if int(case_ids[k]) == case:
                p += factor*vals[k]
    return p


def quad_fer_local(A, q, factors):
    p = _pressure(A, 'qp_elem', A.qp_case, A.qp_val, q, factors)

    def work():
        pts = [A.xyz[k] for k in A.quad_ijmn[q]]
        return el.quad_fer(pts[0], pts[1], pts[2], pts[3], p)
    return _cached(('qF', _shell_key(A, A.quad_ijmn[q], A.quad_props[q])[:3], float(p)), work)


def shell_T_inv(A, ijmn):
    """inv(T) of a shell as `Quad3D.FER` / `Plate3D.FER` use it (`Quad3D.py:870-882`, `Plate3D.py:510-522`)."""
    def work():
        p = A.xyz[ijmn]
        return np.linalg.inv(el.block_diag_T(el.shell_dircos(p[0], p[1], p[3]), 8))
    return _cached(('sTi', _shell_key(A, ijmn, np.zeros(0))[:3]), work)


def plate_fer_local(A, e, factors):
    p = _pressure(A, 'pp_elem', A.pp_case, A.pp_val, e, factors)

    def work():
        pts = [A.xyz[k] for k in A.plate_ijmn[e]]
        w = float(((pts[0] - pts[1])**2).sum()**0.5)
        h = float(((pts[0] - pts[3])**2).sum()**0.5)
        return el.plate_fer(w, h, p)
    return _cached(('pF', _shell_key(A, A.plate_ijmn[e], A.plate_props[e])[:3], float(p)), work)


def combo_factors(A, c):
    return {k: A.factors[c, k] for k in range(A.factors.shape[1]) if A.factors[c, k] != 0.0}


def load_vectors(A, c):
    """Global P and FER vectors of combo index c (`FEModel3D.py:2136-2235`)."""
    n = 6*A.n_nodes
    factors = combo_factors(A, c)
    P = np.zeros((n, 1))
    for k in range(len(A.nl_val)):
        f = factors.get(int(A.nl_case[k]))
        if f is not None:
            P[A.nl_dof[k], 0] += f*A.nl_val[k]
    FER = np.zeros((n, 1))
    for m in range(len(A.mem_rot)):
        if m in _index_by(A, 'ml_member'):
            T = el.block_diag_T(member_geometry(A, m)[1], 4)
            FER[_dofs(*A.mem_ij[m]), 0] += (np.linalg.inv(T) @ member_fer_local(A, m, factors)).reshape(-1)
    for e in range(len(A.plate_ijmn)):
        if e in _index_by(A, 'pp_elem'):
            FER[_dofs(*A.plate_ijmn[e]), 0] += (shell_T_inv(A, A.plate_ijmn[e]) @ plate_fer_local(A, e, factors)).reshape(-1)
    for q in range(len(A.quad_ijmn)):
        if q in _index_by(A, 'qp_elem'):
            FER[_dofs(*A.quad_ijmn[q]), 0] += (shell_T_inv(A, A.quad_ijmn[q]) @ quad_fer_local(A, q, factors)).reshape(-1)
    return P, FER


# ---- solve + reactions -------------------------------------------------------------------------
def _merge(A, D1, D2, d1, d2):
    D = np.zeros((6*A.n_nodes, 1))
    D[d1, 0] = np.asarray(D1).reshape(-1)
    D[d2, 0] = np.asarray(D2).reshape(-1)
    return D


def reactions(A, c, D, pdelta=False):
    """Analysis._calc_reactions for combo index c (`Analysis.py:1059-1313`)."""
    n = 6*A.n_nodes
    factors = combo_factors(A, c)
    R = np.zeros(n)
    has_support = A.support != 0

    def add(node_ids, F):
        for a, nd in enumerate(node_ids):
            if has_support[nd]:
                for k in range(6):
                    if (int(A.support[nd]) >> k) & 1:
                        R[6*nd + k] += F[6*a + k, 0]

    def member_D(ids, active=True):
        d = D[_dofs(*ids), :].copy()
        if not active:                       # Member3D.D / Spring3D.D quirk (:1110-1112)
            d[0, 0] = 0.0
            d[6, 0] = 0.0
        return d

    for s in range(len(A.spring_ks)):
        if A.spring_active[s] and has_support[A.spring_ij[s]].any():
            _, T, k = spring_Ke(A, s)
            add(A.spring_ij[s], np.linalg.inv(T) @ (k @ (T @ member_D(A.spring_ij[s]))))
    for m in range(len(A.mem_rot)):
        if A.mem_active[m] and has_support[A.mem_ij[m]].any():
            _, T, k = member_Ke(A, m)
            d = T @ member_D(A.mem_ij[m])
            fer = member_fer_local(A, m, factors)
            if pdelta:                       # Member3D.f P-Delta branch (:886-891)
                L = member_geometry(A, m)[0]
                E, G, Ar, Iy, Iz, J = A.mem_props[m]
                P = (d[6, 0] - d[0, 0])*Ar*E/L
                k = k + member_Kg(A, m, P)[2]
            add(A.mem_ij[m], np.linalg.inv(T) @ (k @ d + fer))
    for e in range(len(A.plate_ijmn)):
        if has_support[A.plate_ijmn[e]].any():
            _, T, k = plate_Ke(A, e)
            f = k @ (T @ D[_dofs(*A.plate_ijmn[e]), :]) + plate_fer_local(A, e, factors)
            add(A.plate_ijmn[e], np.linalg.inv(T) @ f)
    for q in range(len(A.quad_ijmn)):
        if has_support[A.quad_ijmn[q]].any():
            _, T, k = quad_Ke(A, q)
            f = k @ (T @ D[_dofs(*A.quad_ijmn[q]), :]) + quad_fer_local(A, q, factors)
            add(A.quad_ijmn[q], np.linalg.inv(T) @ f)
    # supported nodal loads (:1270-1287)
    for k in range(len(A.nl_val)):
        f = factors.get(int(A.nl_case[k]))
        dof = int(A.nl_dof[k])
        if f is not None and has_support[dof//6] and (int(A.support[dof//6]) >> (dof % 6)) & 1:
            R[dof] -= A.nl_val[k]*f
    # active support springs, at every node (:1290-1313)
    for nd in range(A.n_nodes):
        for k in range(6):
            fl = int(A.nspring_flag[nd, k])
            if (fl & 1) and (fl & 2):
                R[6*nd + k] -= A.nspring_k[nd, k]*D[6*nd + k, 0]
    return R.reshape(n, 1)


def solve_linear(A, with_reactions=True):
    """analyze_linear on arrays: one factorisation per combo with SuperLU, as the reference does."""
    d1, d2, D2 = partition_D(A)
    K = ke_csr(A)
    K11, K12, K21, K22 = partition_matrix(K, d1, d2)
    out_D, out_R = [], []
    for c in range(A.factors.shape[0]):
        P, FER = load_vectors(A, c)
        if K11.shape == (0, 0):
            D1 = []
        else:
            rhs = np.subtract(np.subtract(P[d1, :], FER[d1, :]), K12.tocsr() @ D2)
            D1 = spsolve(K11.tocsr(), rhs)
        D = _merge(A, D1, D2, d1, d2)
        out_D.append(D)
        if with_reactions:
            out_R.append(reactions(A, c, D))
    return out_D, out_R


def member_axial_P(A, D):
    """P = E A / L (d6 - d0) per member from global displacements (`FEModel3D.py:1849-1850`)."""
    P = np.zeros(len(A.mem_rot))
    for m in range(len(A.mem_rot)):
        L, R, _ = member_geometry(A, m)
        d = el.block_diag_T(R, 4) @ D[_dofs(*A.mem_ij[m]), :]
        E, G, Ar, Iy, Iz, J = A.mem_props[m]
        P[m] = E*Ar/L*(d[6, 0] - d[0, 0])
    return P


def solve_pdelta(A, with_reactions=True):
    """Analysis._PDelta without tension/compression-only features: solve, Kg(P), re-solve."""
    d1, d2, D2 = partition_D(A)
    K = ke_csr(A)
    K11, K12, _, _ = partition_matrix(K, d1, d2)
    out_D, out_R = [], []
    for c in range(A.factors.shape[0]):
        P, FER = load_vectors(A, c)
        rhs = lambda k12: np.subtract(np.subtract(P[d1, :], FER[d1, :]), k12 @ D2)   # noqa: E731
        D = _merge(A, spsolve(K11.tocsr(), rhs(K12.tocsr())), D2, d1, d2)
        Kg = assemble_kg_coo(A, member_axial_P(A, D))
        Kg11, Kg12, _, _ = partition_matrix(Kg, d1, d2)
        Kt11 = K11 + Kg11.tocsr()
        Kt12 = K12 + Kg12.tocsr()
        D = _merge(A, spsolve(Kt11.tocsr(), rhs(Kt12.tocsr())), D2, d1, d2)
        out_D.append(D)
        if with_reactions:
            out_R.append(reactions(A, c, D, pdelta=True))
    return out_D, out_R


# ---- internal forces / stresses of shells --------------------------------------------------------
# evaluation points of the stress fixtures: natural coordinates for quads, fractions of (width, height) for plates
STRESS_POINTS = ((0.0, 0.0), (-1.0, -1.0), (1.0, -1.0), (1.0, 1.0), (-1.0, 1.0), (0.3, -0.6))
PLATE_FRACTIONS = ((0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (0.5, 0.5), (0.3, 0.8))


def shell_stresses(A, D, kind, points, local=True):
    """[count][n_pts][9] = Qx Qy (Qz | 0) Mx My Mxy Sx Sy Txy of every quad ('quad') or plate ('plate') for one
    displacement vector D (6N), at `points` (quads: (xi, eta); plates: local (x, y) given as fractions of width and
    height)."""
    D = np.asarray(D, dtype=float).reshape(-1, 1)
    conn, props = (A.quad_ijmn, A.quad_props) if kind == 'quad' else (A.plate_ijmn, A.plate_props)
    out = np.zeros((len(conn), len(points), 9))
    for e in range(len(conn)):
        p = [A.xyz[k] for k in conn[e]]
        T = el.block_diag_T(el.shell_dircos(p[0], p[1], p[3]), 8)
        d = T @ D[_dofs(*conn[e]), :]
        t, E, nu, kx, ky = props[e]
        for q, (u, v) in enumerate(points):
            if kind == 'quad':
                Q, M, S = el.quad_stress(p[0], p[1], p[2], p[3], t, E, nu, kx, ky, d, u, v, local)
            else:
                w = float(((p[0] - p[1])**2).sum()**0.5)
                h = float(((p[0] - p[3])**2).sum()**0.5)
                Q, M, S = el.plate_stress(w, h, t, E, nu, kx, ky, d, u*w, v*h)
            out[e, q, 0:len(Q)] = Q
            out[e, q, 3:6] = M
            out[e, q, 6:9] = S
    return out
