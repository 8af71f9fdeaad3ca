"""Modal analysis on top of the assembly + solve path: `FEModel3D.M` (`FEModel3D.py:2004-2134`) and
`FEModel3D.analyze_modal` (`:2469-2568`) -- the follow-on consumer SURVEY §8 (f4) names.

The reference hands `K11`, `M11` to ARPACK (`eigsh(K11, k, M=M11, sigma=0.0, which='LM')`, shift-invert about zero:
every Lanczos step is one solve with the SuperLU factor of K11).  Here the solves are the GPU's: a *block* Lanczos
iteration (`lanczos_modes`) asks for `K11^-1 (M11 X)` eight or more columns at a time, which is exactly the multi-RHS
shape the multifrontal factorisation of `pn_solve_linear` is built for -- the block is uploaded as nodal load cases
(`pn_set_loads`), solved, and read back.  The projected (block tridiagonal in exact arithmetic, kept full here because
the basis is fully re-orthogonalised) matrix is a few dozen rows and is diagonalised with LAPACK on the host.

The mass matrix (members and nodal loads only, as in the reference: plates and quads carry no mass there) is built on the
host with numpy over all sub-members at once: it is a linear-time pre-pass whose output feeds `M11 @ X` products of the
same sparsity as the members' stiffness blocks.  There is no CPU solver in this module: without the CUDA library
`analyze_modal` fails in `engine.Engine`, like every other driver.
"""
from __future__ import annotations

import copy

import numpy as np
import scipy.sparse as sps

from pynite_b200 import engine as _eng

_AXIS = {'X': 0, 'Y': 1, 'Z': 2}


# ---------------------------------------------------------------------------------------------
# mass from loads
# ---------------------------------------------------------------------------------------------
def node_mass(node, mass_combo, mass_direction, gravity):
    """`Node3D._calc_mass` (`Node3D.py:153-199`): |sum of factored nodal forces along the mass direction| / g."""
    want = 'F' + mass_direction
    total = 0.0
    for direction, value, case in node.NodeLoads:
        if case in mass_combo.factors and direction == want:
            total += mass_combo.factors[case]*value
    return abs(total/gravity if gravity != 0.0 else 0.0)


def member_load_mass(sub, mass_combo, mass_direction, gravity):
    """`Member3D._calc_load_mass` (`Member3D.py:455-603`): mass of the member's non-self-weight loads in the mass
    combination and the station it is lumped about.  Global-direction loads only: for a local direction the reference
    evaluates `T_local.T()` (calling an array) and stops with a TypeError, which is kept."""
    axis = _AXIS.get(mass_direction)
    force = force_x = 0.0

    def component(direction, *values):
        if direction in ('Fx', 'Fy', 'Fz'):
            raise TypeError("'numpy.ndarray' object is not callable")    # the reference's `T_local.T()` (:506, :549)
        if direction in ('FX', 'FY', 'FZ') and axis is not None and 'XYZ'.index(direction[1]) == axis:
            return values
        return tuple(0.0 for _ in values)

    for direction, P, x, case in sub.PtLoads:
        if case in mass_combo.factors:
            (p,) = component(direction, P)
            force += mass_combo.factors[case]*p
            force_x += mass_combo.factors[case]*p*x
    for direction, w1, w2, x1, x2, case, self_weight in sub.DistLoads:
        if self_weight or case not in mass_combo.factors:
            continue
        factor = mass_combo.factors[case]
        c1, c2 = component(direction, w1, w2)
        avg = (c1 + c2)/2
        force += factor*avg*(x2 - x1)
        force_x += factor*avg*(w2 - w1)*(c1*(x2 - x1)**2/2 + 0.5*(c2 - c1)*(x2 - x1)**2*(2/3))    # (:591-592) as written
    station = force_x/force if force_x != 0.0 else sub.L()/2
    return abs(force/gravity), station


def member_material_mass(sub, mass_combo, gravity):
    """Self-weight mass of `Member3D.consistent_m` (`Member3D.py:362-421`): factor * rho * L * A / g per self-weight
    load whose case is in the mass combination."""
    total = 0.0
    for load in sub.DistLoads:
        if load[6] and load[5] in mass_combo.factors:
            total += mass_combo.factors[load[5]]*sub.material.rho*sub.L()*sub.section.A/gravity
    return total


def member_local_mass(L, A, J, load_mass, station, material_mass):
    """Uncondensed local mass matrices (n, 12, 12) of n members at once: the consistent matrix of the self-weight mass
    (`Member3D.py:388-419`) plus the load mass lumped to the two ends (`lumped_m`, `:423-453`, with its `(1 - x)/L` and
    `x/L` end shares and the `m x^2`, `m (L - x)^2` rotary terms exactly as the reference writes them)."""
    n = len(L)
    m = np.zeros((n, 12, 12))
    c = np.abs(material_mass)/420
    tor = J/A
    for a, b, v in ((0, 0, 140), (0, 6, 70), (6, 6, 140), (1, 1, 156), (1, 7, 54), (7, 7, 156), (2, 2, 156), (2, 8, 54),
                    (8, 8, 156)):
        m[:, a, b] = m[:, b, a] = v*c
    for a, b, v in ((3, 3, 140), (3, 9, 70), (9, 9, 140)):
        m[:, a, b] = m[:, b, a] = v*tor*c
    for a, b, v in ((1, 5, 22), (1, 11, -13), (2, 4, -22), (2, 10, 13), (5, 7, 13), (4, 8, -13), (7, 11, -22), (8, 10, 22)):
        m[:, a, b] = m[:, b, a] = v*L*c
    for a, b, v in ((4, 4, 4), (5, 5, 4), (10, 10, 4), (11, 11, 4), (4, 10, -3), (5, 11, -3)):
        m[:, a, b] = m[:, b, a] = v*L*L*c
    mi = load_mass*(1 - station)/L
    mj = load_mass*station/L
    for k in (0, 1, 2):
        m[:, k, k] += mi
        m[:, 6 + k, 6 + k] += mj
    for k in (4, 5):
        m[:, k, k] += mi*station**2
        m[:, 6 + k, 6 + k] += mj*(L - station)**2
    return m


def condense_released(m, releases):
    """`Member3D.m` (`:605-646`): static condensation of the released DOFs, `m11 - m12 pinv(m22) m21`, expanded back with
    zero rows / columns at the released DOFs."""
    rel = np.asarray(releases, dtype=bool)
    if not rel.any():
        return m
    r1, r2 = np.nonzero(~rel)[0], np.nonzero(rel)[0]
    cond = m[np.ix_(r1, r1)] - m[np.ix_(r1, r2)] @ np.linalg.pinv(m[np.ix_(r2, r2)]) @ m[np.ix_(r2, r1)]
    out = np.zeros((12, 12))
    out[np.ix_(r1, r1)] = cond
    return out


def member_global_mass(model, subs, mass_combo, mass_direction, gravity):
    """`Member3D.M` (`:648-672`) for a list of sub-members: (n, 12, 12) global matrices `T^T m T`."""
    from pynite_b200.analysis import member_dircos_host
    n = len(subs)
    L = np.array([s.L() for s in subs], dtype=float)
    A = np.array([s.section.A for s in subs], dtype=float)
    J = np.array([s.section.J for s in subs], dtype=float)
    lm = np.array([member_load_mass(s, mass_combo, mass_direction, gravity) for s in subs], dtype=float).reshape(n, 2)
    mm = np.array([member_material_mass(s, mass_combo, gravity) for s in subs], dtype=float)
    m = member_local_mass(L, A, J, lm[:, 0], lm[:, 1], mm)
    for k, s in enumerate(subs):
        if True in s.Releases:
            m[k] = condense_released(m[k], s.Releases)
    R = np.array([member_dircos_host(s) for s in subs], dtype=float).reshape(n, 3, 3)
    blocks = m.reshape(n, 4, 3, 4, 3)                       # [member, block row, i, block column, j]
    return np.einsum('nia,nbicj,njd->nbacd', R, blocks, R, optimize=True).reshape(n, 12, 12)


def global_M(model, mass_combo_name=None, mass_direction='Y', gravity=1.0, log=False, sparse=True):
    """`FEModel3D.M` (`FEModel3D.py:2004-2134`): member blocks (active physical members, sub-member by sub-member), then
    the translational nodal masses, exact zeros dropped from every block (`_append_sparse_block`, `:99-151`), and a mass of
    1e-6 x the smallest positive diagonal entry on every DOF whose diagonal is zero."""
    if mass_combo_name not in model.load_combos:
        raise ValueError(f'Load combination {mass_combo_name} not found in the model.')
    combo = model.load_combos[mass_combo_name]
    if any(nd.ID is None for nd in model.nodes.values()) or any(
            not pm.sub_members for pm in model.members.values()):
        from pynite_b200.analysis import renumber
        renumber(model)
    n_dof = 6*len(model.nodes)
    rows, cols, vals = [], [], []
    subs = [s for pm in model.members.values() if pm.active.get(mass_combo_name, True) == True   # noqa: E712
            for s in pm.sub_members.values()]
    if subs:
        Mg = member_global_mass(model, subs, combo, mass_direction, gravity).reshape(len(subs), 144)
        ij = np.array([(s.i_node.ID, s.j_node.ID) for s in subs], dtype=np.int64)
        dofs = (ij[:, :, None]*6 + np.arange(6)).reshape(len(subs), 12)
        keep = Mg != 0.0
        rows.append(np.repeat(dofs, 12, axis=1)[keep])
        cols.append(np.tile(dofs, (1, 12))[keep])
        vals.append(Mg[keep])
    nodes = list(model.nodes.values())
    mass = np.array([node_mass(nd, combo, mass_direction, gravity) for nd in nodes], dtype=float)
    heavy = np.nonzero(mass > 0)[0]
    if len(heavy):
        d = (heavy[:, None]*6 + np.arange(3)).reshape(-1)
        rows.append(d)
        cols.append(d)
        vals.append(np.repeat(mass[heavy], 3))
    cat = (lambda parts, dt: np.concatenate(parts) if parts else np.array([], dtype=dt))
    M = sps.coo_matrix((cat(vals, float), (cat(rows, np.int64), cat(cols, np.int64))), shape=(n_dof, n_dof))
    diag = M.diagonal()
    positive = diag[diag > 0]
    if positive.size == 0:
        raise Exception('Unable to perform modal analysis. Model is massless.')
    M = M + sps.diags(positive.min()*1e-6*(diag == 0).astype(float), 0, shape=M.shape)
    return M if (sparse is True or sparse == True) else M.toarray()   # noqa: E712


# ---------------------------------------------------------------------------------------------
# eigenvalue iteration
# ---------------------------------------------------------------------------------------------
def _m_orthonormalise(Y, MY, floor, room=None):
    """Columns of Y made M-orthonormal: returns (X, MX, R) with Y = X R; directions whose M-norm is below `floor` are
    dropped (the Krylov space is exhausted, or the block has converged onto an invariant subspace), and at most `room`
    directions -- the strongest -- are kept."""
    G = Y.T @ MY
    w, Q = np.linalg.eigh((G + G.T)/2)
    keep = w > floor*floor
    if room is not None:
        keep[:max(0, len(w) - room)] = False             # eigh sorts ascending
    if not keep.any():
        return Y[:, :0], MY[:, :0], np.zeros((0, Y.shape[1]))
    s = np.sqrt(w[keep])
    Qk = Q[:, keep]
    return (Y @ Qk)/s, (MY @ Qk)/s, (Qk*s).T


def lanczos_modes(solve, M, k, tol=1e-12, block=None, max_basis=None, seed=20261019):
    """The k smallest eigenpairs of `K phi = lambda M phi` by block shift-invert Lanczos about zero with full
    M-re-orthogonalisation.  `solve(B)` returns `K^-1 B` for an (n, b) block; M is the (n, n) sparse mass matrix.

    Returns (lambda ascending (k,), phi (n, k) with phi^T M phi = I, info).  The Ritz pairs of `K^-1 M` on the basis
    V = [X0, X1, ...] come from T = V^T M K^-1 M V, whose block column j is read off the orthogonalisation coefficients of
    `K^-1 M X_j`; a pair has converged when the coupling of its Ritz vector to the next block, relative to its Ritz
    value, is below `tol`."""
    n = M.shape[0]
    if not 0 < k < n:
        raise TypeError(f'k={k} must be between 1 and the number of free degrees of freedom - 1 = {n - 1}')
    b = min(n, block or max(8, k))
    max_basis = min(n, max_basis or max(20*b, 12*k))
    rng = np.random.default_rng(seed)
    Y = rng.standard_normal((n, b))
    X, MX, _ = _m_orthonormalise(Y, M @ Y, 0.0)
    V, MV = [X], [MX]
    T = np.zeros((0, 0))
    theta_max = 0.0
    info = {'solves': 0, 'columns': 0, 'converged': False}
    while True:
        Y = solve(MV[-1])
        info['solves'] += 1
        info['columns'] += MV[-1].shape[1]
        Vc, MVc = np.hstack(V), np.hstack(MV)
        H = MVc.T @ Y
        Y = Y - Vc @ H
        H2 = MVc.T @ Y                                   # second pass: the basis stays M-orthonormal to rounding
        Y = Y - Vc @ H2
        H += H2
        m_old, m_new = T.shape[0], Vc.shape[1]
        T2 = np.zeros((m_new, m_new))
        T2[:m_old, :m_old] = T
        T2[:, m_old:] = H                                # block column of the newest block, rows of every block so far
        T2[m_old:, :m_old] = H[:m_old].T
        T = (T2 + T2.T)/2
        theta, S = np.linalg.eigh(T)
        theta_max = max(theta_max, float(np.abs(theta).max()))
        # what is left of Y after the projection is the next block; below 1e-10 of the largest Ritz value it is rounding
        # noise of the projection (directions that weak only matter to eigenvalues 1e10 times the first one)
        Xn, MXn, R = _m_orthonormalise(Y, M @ Y, 1e-10*theta_max, max_basis - m_new)
        top = np.argsort(theta)[::-1][:k]
        last = S[m_new - V[-1].shape[1]:, top]           # rows of the newest block
        resid = np.linalg.norm(R @ last, axis=0)/np.abs(theta[top])
        info['residual'] = float(resid.max())
        if resid.max() < tol or (Xn.shape[1] == 0 and (m_new == n or R.shape[0] == 0)):
            info['converged'] = True                     # (a basis of n vectors spans everything: the pairs are exact)
            break
        if Xn.shape[1] == 0:
            break
        V.append(Xn)
        MV.append(MXn)
    lam = 1.0/theta[top]
    phi = np.hstack(V) @ S[:, top]
    order = np.argsort(lam)
    info['basis'] = T.shape[0]
    return lam[order], phi[:, order], info


class EngineInverse:
    """`B -> K11^-1 B` through the C ABI: the columns of B become nodal load cases on the free DOFs
    (`pn_set_loads` / `pn_set_combos` with a unit factor table) and `pn_solve_linear` factorises and solves them as one
    multi-RHS batch.  The model on the engine must carry zero enforced displacements (so that the right-hand side is B
    alone, `Analysis.py:197-207`)."""

    def __init__(self, eng, arrays, d1, check_stability=True):
        self.eng, self.d1 = eng, np.asarray(d1, dtype=np.int32)
        self.opts = _eng.Engine.make_opts(solver='direct', check_stability=check_stability)
        if getattr(eng, '_out_range', None) is not None:
            eng.set_output_combos(0, -1)                 # a multi-GPU solve may have narrowed the download range
        L = copy.copy(arrays)
        for name, dt, shape in (('ml_member', np.int32, (0,)), ('ml_case', np.int32, (0,)), ('ml_kind', np.int32, (0,)),
                                ('ml_params', np.float64, (0, 4)), ('qp_elem', np.int32, (0,)), ('qp_case', np.int32, (0,)),
                                ('qp_val', np.float64, (0,)), ('pp_elem', np.int32, (0,)), ('pp_case', np.int32, (0,)),
                                ('pp_val', np.float64, (0,))):
            setattr(L, name, np.zeros(shape, dtype=dt))
        self.loads = L

    def __call__(self, B):
        n1, s = B.shape
        L = self.loads
        L.cases = list(range(s))
        L.nl_dof = np.ascontiguousarray(np.tile(self.d1, s))
        L.nl_case = np.repeat(np.arange(s, dtype=np.int32), n1)
        L.nl_val = np.ascontiguousarray(B.T).reshape(-1)
        self.eng.set_loads(L)
        self.eng.set_combos(np.eye(s))
        D, _, _ = self.eng.solve_linear(self.opts, reactions=False)
        return np.ascontiguousarray(D[:, self.d1].T)


# ---------------------------------------------------------------------------------------------
# driver
# ---------------------------------------------------------------------------------------------
def run_modal(model, num_modes, mass_combo_name, mass_direction, gravity, log, check_stability, inverse_factory=None):
    """`FEModel3D.analyze_modal` (`FEModel3D.py:2469-2568`).  `inverse_factory(model, arrays, d1)` replaces the GPU
    operator in the host-logic tests only."""
    from pynite_b200 import analysis as _an
    from pynite_b200.flatten import flatten_model, partition_dofs
    if log:
        print('+------------------+')
        print('| Analyzing: Modal |')
        print('+------------------+')
    _an.prepare_model(model, num_modes)
    combo = model.load_combos[mass_combo_name]
    A = flatten_model(model, [combo], mass_combo_name)
    d1, d2, v2 = partition_dofs(A)
    if log:
        print('- Assembling global stiffness matrix')
    if inverse_factory is None:
        B = copy.copy(A)
        B.enf_val = np.zeros_like(A.enf_val)             # mode shapes: K11 phi = lambda M11 phi, no K12 D2 term
        eng = _an._engine(model)
        eng.set_model(B)
        eng.set_loads(B)
        eng.set_combos(B.factors)
        _an._assemble(model, eng, check_stability, log)
        inverse = EngineInverse(eng, B, d1, check_stability)
    else:
        inverse = inverse_factory(model, A, d1)
    model._run = None
    if log:
        print('- Assembling global mass matrix')
    M = sps.csr_matrix(global_M(model, mass_combo_name, mass_direction, gravity, log))
    M11 = M[d1, :][:, d1]
    if M11.nnz == 0:
        raise Exception('No mass terms found. Ensure materials have density or provide mass_combo_name.')
    if log:
        print('- Solving eigenvalue problem')
    try:
        lam, phi, info = lanczos_modes(inverse, M11, num_modes)
    except _eng.SingularMatrix as err:
        raise Exception(f'Eigenvalue solution failed: {err}. Check matrix conditioning.')
    frequencies = np.sqrt(lam)/(2*np.pi)
    if log:
        print('- Processing mode shapes')
    for i in range(len(frequencies)):
        D = np.zeros((A.n_dof, 1))
        D[d1, 0] = phi[:, i]
        D[d2, 0] = v2
        model._D[f'Mode {i + 1}'] = D
    model.frequencies = frequencies
    model.modal_info = info
    model.solution = 'Modal'
    if log:
        print(f'- Found {len(frequencies)} modes')
        for i, freq in enumerate(frequencies):
            print(f'  Mode {i + 1}: {freq:.3f} Hz')
        print('- Modal analysis complete')
