"""An `engine.Engine` look-alike backed by the CPU oracle -- TEST INFRASTRUCTURE ONLY.

The drivers of `pynite_b200/analysis.py` and `modal.py` talk to the GPU through `engine.Engine`.  The container the tests
run in has no GPU, so the Python glue between the reference-facing API and the engine (upload, result storing, diagrams after
an analysis, the modal operator, the shear-wall results, the per-object matrices) would only ever run on the B200 box.
`OracleEngine` implements the slice of the Engine interface those drivers use for models WITHOUT tension/compression-only
features, with `oracle.model` doing the arithmetic, so that the glue can be driven end to end on the CPU
(tests/test_drivers_cpu.py).  Nothing in the product package imports this module, and the parity of the CUDA kernels is
established by the `-m gpu` tests, not here."""
import copy

import numpy as np

from oracle import model as om
from pynite_b200 import engine as real


class _Stats:
    def as_dict(self):
        return {}


class OracleEngine:

    def __init__(self, device=0):
        self.device = device
        self.A = None
        self.D = None
        self.pdelta = False
        self._check = True

    # ---- upload ----------------------------------------------------------------------------------------------
    def set_model(self, A):
        self.A = copy.copy(A)
        self.n_dof = 6*A.n_nodes
        self.counts = {'spring': len(A.spring_ks), 'member': len(A.mem_rot), 'quad': len(A.quad_ijmn),
                       'plate': len(A.plate_ijmn)}

    def set_loads(self, L):
        for name in ('cases', 'nl_dof', 'nl_case', 'nl_val', 'ml_member', 'ml_case', 'ml_kind', 'ml_params', 'qp_elem', 'qp_case',
                     'qp_val', 'pp_elem', 'pp_case', 'pp_val'):
            setattr(self.A, name, copy.copy(getattr(L, name)))

    def set_combos(self, factors):
        n_case = max(len(self.A.cases), 1)
        self.A.factors = np.asarray(factors, dtype=float).reshape(-1, n_case) if len(self.A.cases) else np.zeros((len(factors), 0))
        self.n_combo = self.A.factors.shape[0]

    def set_output_combos(self, combo0, n_combo):
        pass

    def reset_stats(self):
        pass

    def stats(self):
        return _Stats()

    # ---- assembly / solve ------------------------------------------------------------------------------------
    def symbolic(self, flags=0):
        return self.sizes()

    def sizes(self):
        from types import SimpleNamespace
        d1, d2, _ = om.partition_D(self.A)
        return SimpleNamespace(n_dof=self.n_dof, n1=len(d1), n2=len(d2))

    def coo(self):
        return om.assemble_coo(self.A)

    def assemble_ke(self):
        pass

    def check_nodal_stability(self, cap=4096):
        """`Analysis._check_stability` (`Analysis.py:111-168`): a free DOF with a zero diagonal stiffness."""
        from math import isclose
        d1, _, _ = om.partition_D(self.A)
        diag = om.ke_csr(self.A).diagonal()
        bad = [int(i) for i in d1 if isclose(diag[i], 0.0)]
        if bad:
            raise real.UnstableNodes('unstable nodes', bad)

    def partition(self):
        return om.partition_D(self.A)

    def _solve(self, fn, reactions):
        try:
            D, R = fn(self.A, with_reactions=reactions)
        except Exception as err:                              # noqa: BLE001
            raise real.SingularMatrix(str(err))
        self.D = np.array([np.asarray(d).reshape(-1) for d in D]).reshape(len(D), self.n_dof)
        if self._check and len(D):
            # the reference's solution check (`Analysis.py:228-239`): finite, and the residual on the free DOFs small
            d1, _, _ = om.partition_D(self.A)
            K = om.ke_csr(self.A)
            for c in range(len(D)):
                P, FER = om.load_vectors(self.A, c)
                rhs = (P - FER)[d1, 0]
                res = (K @ self.D[c])[d1] - rhs
                scale = np.linalg.norm(rhs)
                if not np.isfinite(self.D[c]).all() or (fn is om.solve_linear and np.linalg.norm(res) > 1e-6*max(scale, 1e-300)
                                                        and scale > 0):
                    raise real.SingularMatrix('singular stiffness matrix')
        Rm = np.array([np.asarray(r).reshape(-1) for r in R]).reshape(len(R), self.n_dof) if reactions else None
        return self.D.copy(), Rm, _Stats()

    def solve_linear(self, opts=None, reactions=True, D_out=None, R_out=None, download=True):
        self.pdelta = False
        self._check = bool(opts is None or opts.check_stability)
        return self._solve(om.solve_linear, reactions)

    def solve_pdelta(self, opts=None, reactions=True, D_out=None, R_out=None, download=True):
        self.pdelta = True
        self._check = bool(opts is None or opts.check_stability)
        return self._solve(om.solve_pdelta, reactions)

    def set_displacements(self, combo, D):
        self.D[combo] = np.asarray(D, dtype=float).reshape(-1)

    def load_vector(self, combo, which):
        P, FER = om.load_vectors(self.A, combo)
        return np.asarray(FER if which in (1, 'FER', 'fer') else P).reshape(-1, 1)

    def reactions(self, combo, pdelta=False, out=None):
        return om.reactions(self.A, combo, self.D[combo].reshape(-1, 1), pdelta=pdelta)

    # ---- tension / compression-only iteration (`pn_tc.cu`, `Analysis._check_TC_convergence` `Analysis.py:917-1056`) -------
    def set_tc_modes(self, spring_mode, member_mode, group_start):
        self.smode, self.mmode, self.starts = np.asarray(spring_mode), np.asarray(member_mode), np.asarray(group_start)

    def tc_reset(self):
        self.A.spring_active = np.ones_like(self.A.spring_active)
        self.A.mem_active = np.ones_like(self.A.mem_active)

    def activity(self, n_nodes):
        return self.A.nspring_flag.copy(), self.A.spring_active.copy(), self.A.mem_active.copy()

    def set_activity(self, nspring_flag=None, spring_active=None, member_active=None):
        for name, value in (('nspring_flag', nspring_flag), ('spring_active', spring_active), ('mem_active', member_active)):
            if value is not None:
                setattr(self.A, name, np.array(value, copy=True))

    def _member_axial_extremes(self, m, f, D):
        """(max, min) of the axial force diagram of sub-member m, through the package's own segment table."""
        from pynite_b200.diagrams import SegmentTable
        A = self.A
        L, R, _ = om.member_geometry(A, m)
        factors = om.combo_factors(A, 0)
        pts, dists, stations = [], [], []
        raw_pts, raw_dists = om._member_loads(A, m)
        for direction, P, x, case in raw_pts:
            stations.append(x)
            if case in factors:
                v = np.zeros(6)
                k = 'xyz'.find(direction[1].lower())
                if direction[1].isupper():
                    v[3*(direction[0] == 'M'):3*(direction[0] == 'M') + 3] = R[:, k]*P*factors[case]
                else:
                    v[3*(direction[0] == 'M') + k] = P*factors[case]
                pts.append((x, v))
        for direction, w1, w2, x1, x2, case in raw_dists:
            stations += [x1, x2]
            if case in factors:
                k = 'xyz'.find(direction[1].lower())
                if direction[1].isupper():
                    a, b = R[:, k]*w1*factors[case], R[:, k]*w2*factors[case]
                else:
                    a, b = np.zeros(3), np.zeros(3)
                    a[k], b[k] = w1*factors[case], w2*factors[case]
                dists.append((x1, x2, a, b))
        E, G, Ar, Iy, Iz, J = A.mem_props[m]
        dofs = [6*int(n) + j for n in A.mem_ij[m] for j in range(6)]
        d = np.kron(np.eye(4), R) @ D[dofs]
        table = SegmentTable(L, E, Ar, Iy, Iz, f, d, pts, dists, stations, False)
        return table.extreme('axial', 'z', True), table.extreme('axial', 'z', False)

    def tc_update(self, combo, pdelta, spring_tol, member_tol):
        A, D = self.A, self.D[combo]
        changes = 0
        flags = A.nspring_flag.copy()
        for n, k in zip(*np.nonzero(flags & 12)):
            disp = D[6*n + k]
            should = bool(((flags[n, k] & 8) and disp <= -spring_tol) or ((flags[n, k] & 4) and disp >= spring_tol))
            if bool(flags[n, k] & 2) != should:
                flags[n, k] = (flags[n, k] & ~np.uint8(2)) | (2 if should else 0)
                changes += 1
        sf = self.element_forces('spring', combo, local=True)
        spring_active = A.spring_active.copy()
        for s in range(self.counts['spring']):
            if spring_active[s] and (((self.smode[s] & 1) and sf[s, 0] > spring_tol) or ((self.smode[s] & 2) and sf[s, 0] < -spring_tol)):
                spring_active[s] = 0
                changes += 1
        mf = self.element_forces('member', combo, local=True, pdelta=pdelta)
        mem_active = A.mem_active.copy()
        for g in range(len(self.starts) - 1):
            subs = range(int(self.starts[g]), int(self.starts[g + 1]))
            if not len(subs) or not mem_active[subs[0]] or not self.mmode[subs[0]]:
                continue
            ext = [self._member_axial_extremes(m, mf[m], D) for m in subs]
            if ((self.mmode[subs[0]] & 1) and max(e[0] for e in ext) > member_tol) or \
                    ((self.mmode[subs[0]] & 2) and min(e[1] for e in ext) < -member_tol):
                mem_active[list(subs)] = 0
                changes += 1
        A.nspring_flag, A.spring_active, A.mem_active = flags, spring_active, mem_active
        return changes

    # ---- element level ---------------------------------------------------------------------------------------
    def element_ke(self, kind):
        fn = {'spring': om.spring_Ke, 'member': om.member_Ke, 'quad': om.quad_Ke, 'plate': om.plate_Ke}[kind]
        nd = 12 if kind in ('spring', 'member') else 24
        return np.array([fn(self.A, k)[0] for k in range(self.counts[kind])]).reshape(self.counts[kind], nd, nd)

    def member_kg(self, P):
        return np.array([om.member_Kg(self.A, k, float(P[k]))[0] for k in range(self.counts['member'])]).reshape(-1, 12, 12)

    def element_forces(self, kind, combo, local=False, pdelta=False):
        A, D = self.A, self.D[combo].reshape(-1, 1)
        factors = om.combo_factors(A, combo)
        conn = {'spring': A.spring_ij, 'member': A.mem_ij, 'quad': A.quad_ijmn, 'plate': A.plate_ijmn}[kind]
        out = []
        for k in range(self.counts[kind]):
            dofs = [6*int(n) + j for n in conn[k] for j in range(6)]
            if kind == 'spring':
                _, T, ke = om.spring_Ke(A, k)
                f = ke @ (T @ D[dofs, :])*(1.0 if A.spring_active[k] else 0.0)
            elif kind == 'member' and not A.mem_active[k]:
                T, f = om.member_Ke(A, k)[1], np.zeros((12, 1))
            elif kind == 'member':
                _, T, ke = om.member_Ke(A, k)
                d = T @ D[dofs, :]
                if pdelta:
                    L = om.member_geometry(A, k)[0]
                    E, G, Ar, Iy, Iz, J = A.mem_props[k]
                    ke = ke + om.member_Kg(A, k, (d[6, 0] - d[0, 0])*Ar*E/L)[2]
                f = ke @ d + om.member_fer_local(A, k, factors)
            elif kind == 'quad':
                _, T, ke = om.quad_Ke(A, k)
                f = ke @ (T @ D[dofs, :]) + om.quad_fer_local(A, k, factors)
            else:
                _, T, ke = om.plate_Ke(A, k)
                f = ke @ (T @ D[dofs, :]) + om.plate_fer_local(A, k, factors)
            out.append((f if local else np.linalg.inv(T) @ f).reshape(-1))
        nd = 12 if kind in ('spring', 'member') else 24
        return np.array(out).reshape(self.counts[kind], nd)

    def shell_stresses(self, kind, combo0, n_combo, pts, local=True):
        """Quads: natural coordinates.  Plates: LOCAL (x, y) as in `pn_shell_stresses` -- the oracle takes fractions of the
        plate's width and height, so the points are converted element by element."""
        pts = np.asarray(pts, dtype=float).reshape(-1, 2)
        A = self.A
        out = np.zeros((n_combo, self.counts[kind], len(pts), 9))
        for c in range(n_combo):
            D = self.D[combo0 + c]
            if kind == 'quad':
                out[c] = om.shell_stresses(A, D, 'quad', [tuple(p) for p in pts], local)
                continue
            for e in range(self.counts['plate']):
                p = [A.xyz[k] for k in A.plate_ijmn[e]]
                w = float(((p[0] - p[1])**2).sum()**0.5)
                h = float(((p[0] - p[3])**2).sum()**0.5)
                one = copy.copy(A)
                one.plate_ijmn, one.plate_props = A.plate_ijmn[e:e + 1], A.plate_props[e:e + 1]
                out[c, e] = om.shell_stresses(one, D, 'plate', [(x/w, y/h) for x, y in pts], True)[0]
        return out

def install(monkeypatch):
    """Makes `analysis._engine(model)` hand out an OracleEngine (one per model) for the duration of a test."""
    from pynite_b200 import analysis

    def engine_of(model):
        if not isinstance(getattr(model, '_engine', None), OracleEngine):
            model._engine = OracleEngine()
        return model._engine
    monkeypatch.setattr(analysis, '_engine', engine_of)
    monkeypatch.setattr(real, 'Engine', type('Engine', (OracleEngine,), {'make_opts': staticmethod(real.Engine.make_opts)}))
