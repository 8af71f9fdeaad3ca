"""Member result diagrams: shear, moment, axial force, torque and deflection along a member, their extremes and sampled
arrays -- the consumers of the element end forces the GPU solve produces (`pn_element_forces`), SURVEY §8 (f1).

The reference cuts every sub-member into "mathematically continuous" segments at its load stations and carries, per
segment and bending plane, the start values of the internal forces and of the elastic curve
(`Member3D._segment_member`, `Member3D.py:2836-3147`; `BeamSegZ.py`, `BeamSegY.py`), then answers every query by
walking the segment objects.  Here one `SegmentTable` per (sub-member, combination) holds those start values as numpy
columns -- one row per segment, the two bending planes side by side -- so that a query at a point is a row lookup and a
sampled diagram is one vectorised evaluation.  The two planes share every formula up to one sign (`_SIGMA`): 'z' is bending
about local z (shear Fy, moment Mz, deflection dy), 'y' is bending about local y (Fz, My, dz).

Reference behaviour kept on purpose (each is pinned by tests/golden/member_diagrams.npz):
  * segment lookup compares stations rounded to 10 decimals (`Member3D.py:1186-1201`); arrays use plain comparisons
    (`_extract_vector_results`, `:3149-3247`);
  * axial point loads enter the axial force of the 'z' plane only, so the P-delta term of the 'y' plane does not see them
    (`:2988-2989, 3012`);
  * the 'y' plane's `max_moment` evaluates its candidates without the P-delta term, its `min_moment` with it
    (`BeamSegY.py:95-134`);
  * deflection extremes are the extremes over 100 equally spaced stations (`Member3D.py:2428-2433`);
  * an inactive member has zero internal forces and deflects as the straight chord between its nodes (`:2358-2370`).
"""
from __future__ import annotations

from math import isclose

import numpy as np

_SIGMA = {'z': 1.0, 'y': -1.0}
_LOCAL = {'Fx': 0, 'Fy': 1, 'Fz': 2, 'Mx': 3, 'My': 4, 'Mz': 5}
_GLOBAL_F = {'FX': 0, 'FY': 1, 'FZ': 2}
_GLOBAL_M = {'MX': 0, 'MY': 1, 'MZ': 2}


def _r10(v):
    return round(v, 10)


def local_loads(sub, combo, R):
    """The sub-member's loads whose case is in the combination, factored and resolved on the local axes:
    point loads as (x, [Fx, Fy, Fz, Mx, My, Mz]) and distributed loads as (x1, x2, w1[3], w2[3])."""
    pts, dists = [], []
    for direction, P, x, case in sub.PtLoads:
        factor = combo.factors.get(case)
        if factor is None:
            continue
        v = np.zeros(6)
        if direction in _LOCAL:
            v[_LOCAL[direction]] = factor*P
        elif direction in _GLOBAL_F:
            v[:3] = factor*(R[:, _GLOBAL_F[direction]]*P)
        elif direction in _GLOBAL_M:
            v[3:] = factor*(R[:, _GLOBAL_M[direction]]*P)
        else:
            raise Exception('Invalid member point load direction specified.')
        pts.append((x, v))
    for load in sub.DistLoads:
        direction, w1, w2, x1, x2, case = load[:6]
        factor = combo.factors.get(case)
        if factor is None:
            continue
        if direction in ('Fx', 'Fy', 'Fz'):
            a, b = np.zeros(3), np.zeros(3)
            a[_LOCAL[direction]], b[_LOCAL[direction]] = factor*w1, factor*w2
        elif direction in _GLOBAL_F:
            a, b = R[:, _GLOBAL_F[direction]]*(factor*w1), R[:, _GLOBAL_F[direction]]*(factor*w2)
        else:
            continue
        dists.append((x1, x2, a, b))
    return pts, dists


def fixed_end_moments(pts, dists, L):
    """Entries (4, 5, 10, 11) of the uncondensed local fixed-end reaction vector (`Member3D._fer_unc`, `:774-850`): the end
    moments [My_i, Mz_i, My_j, Mz_j] of the clamped member (`fixed_end.member_fer_unc`)."""
    from pynite_b200.fixed_end import member_fer_unc
    return member_fer_unc(pts, dists, L)[[4, 5, 10, 11]]


class SegmentTable:
    """Start values of every segment of one sub-member for one combination (see the module docstring)."""

    def __init__(self, L, E, A, Iy, Iz, f, d, pts, dists, stations, p_delta):
        f, d = np.asarray(f, dtype=float).reshape(12), np.asarray(d, dtype=float).reshape(12)
        self.L, self.p_delta = L, bool(p_delta)
        self.EI = {'z': E*Iz, 'y': E*Iy}
        self.EA = E*A
        xs = sorted(set([0, L] + list(stations)))
        n = len(xs) - 1
        self.x1, self.x2 = np.array(xs[:-1], dtype=float), np.array(xs[1:], dtype=float)
        z = lambda: np.zeros(n)                            # noqa: E731
        self.p1, self.p2, self.T1 = z(), z(), z()
        self.P1 = {'z': z(), 'y': z()}
        self.w1, self.w2 = {'z': z(), 'y': z()}, {'z': z(), 'y': z()}
        self.V1, self.M1 = {'z': z(), 'y': z()}, {'z': z(), 'y': z()}
        self.theta1, self.delta1 = {'z': z(), 'y': z()}, {'z': z(), 'y': z()}
        self.dx1 = {'z': z(), 'y': z()}
        fem = fixed_end_moments(pts, dists, L)
        # slope-deflection at the i end (`Member3D.py:2914-2915`): end moments less the clamped ones, plus the chord rotation
        self.theta1['z'][0] = 1/3*((f[5] - fem[1])*L/self.EI['z'] - (f[11] - fem[3])*L/(2*self.EI['z']) + 3*(d[7] - d[1])/L)
        self.theta1['y'][0] = -1/3*((-f[4] + fem[0])*L/self.EI['y'] - (-f[10] + fem[2])*L/(2*self.EI['y']) + 3*(d[8] - d[2])/L)
        self.delta1['z'][0], self.delta1['y'][0] = d[1], d[2]
        self.dx1['z'][0] = self.dx1['y'][0] = d[0]
        for i in range(n):
            x, xe = self.x1[i], self.x2[i]
            if i > 0:
                h = self.x2[i - 1] - self.x1[i - 1]
                for pl in ('z', 'y'):
                    self.theta1[pl][i] = self.slope(pl, i - 1, h)
                    self.delta1[pl][i] = self.deflection(pl, i - 1, h)
                    self.dx1[pl][i] = self.axial_deflection(pl, i - 1, h)
            Pz = Py = f[0]
            Vz, Mz, Vy, My, T = f[1], f[5] - f[1]*x, f[2], f[4] + f[2]*x, f[3]
            for xp, v in pts:
                if _r10(xp) <= _r10(x):
                    Pz += v[0]
                    Vz += v[1]
                    Mz += v[5] - v[1]*(x - xp)
                    Vy += v[2]
                    My += v[4] + v[2]*(x - xp)
                    T += v[3]
            for xa, xb, a, b in dists:
                if _r10(xa) > _r10(x):
                    continue
                if _r10(xb) > _r10(x):                       # the load runs on past the segment start: its intensity here
                    slope = (b - a)/(xb - xa)
                    here, there = a + slope*(x - xa), a + slope*(xe - xa)
                    self.p1[i] += here[0]
                    self.p2[i] += there[0]
                    self.w1['z'][i] += here[1]
                    self.w2['z'][i] += there[1]
                    self.w1['y'][i] += here[2]
                    self.w2['y'][i] += there[2]
                    b, xb = here, x
                h = xb - xa
                total = (a + b)/2*h                          # resultant of the part of the load left of x ...
                arm = total*(x - xa) - h*h*(a + 2*b)/6       # ... and its moment about x
                Pz += total[0]
                Py += total[0]
                Vz += total[1]
                Mz -= arm[1]
                Vy += total[2]
                My += arm[2]
            self.P1['z'][i], self.P1['y'][i], self.T1[i] = Pz, Py, T
            self.V1['z'][i], self.M1['z'][i], self.V1['y'][i], self.M1['y'][i] = Vz, Mz, Vy, My

    # ---- one segment, local station x (scalar or array) ---------------------------------------------------------
    def _h(self, i):
        return self.x2[i] - self.x1[i]

    def shear(self, pl, i, x):
        return self.V1[pl][i] + self.w1[pl][i]*x + x**2*(self.w2[pl][i] - self.w1[pl][i])/(2*self._h(i))

    def deflection(self, pl, i, x, p_delta=None):
        pd = self.p_delta if p_delta is None else p_delta
        s, EI, w1, w2 = _SIGMA[pl], self.EI[pl], self.w1[pl][i], self.w2[pl][i]
        v = (self.delta1[pl][i] + s*self.theta1[pl][i]*x + self.V1[pl][i]*x**3/(6*EI) + w1*x**4/(24*EI)
             - s*self.M1[pl][i]*x**2/(2*EI) + x**5*(w2 - w1)/(120*EI*self._h(i)))
        if pd:
            P = self.P1[pl][i]
            v = (v + P*self.delta1[pl][i]*x**2/(2*EI))/(1 + P*x**2/(2*EI))
        return v

    def moment(self, pl, i, x, p_delta=None):
        pd = self.p_delta if p_delta is None else p_delta
        w1, w2 = self.w1[pl][i], self.w2[pl][i]
        M = _SIGMA[pl]*self.M1[pl][i] - self.V1[pl][i]*x - w1*x**2/2 - x**3*(w2 - w1)/(6*self._h(i))
        if pd:
            M = M + self.P1[pl][i]*(self.deflection(pl, i, x, True) - self.delta1[pl][i])
        return M

    def slope(self, pl, i, x, p_delta=None):
        pd = self.p_delta if p_delta is None else p_delta
        s, EI, w1, w2 = _SIGMA[pl], self.EI[pl], self.w1[pl][i], self.w2[pl][i]
        th = (self.theta1[pl][i] - s*(-self.V1[pl][i]*x**2/2 - w1*x**3/6 + x**4*(w1 - w2)/(24*self._h(i)))/EI
              - x*self.M1[pl][i]/EI)
        if pd:
            th = th - s*x*self.P1[pl][i]*(self.deflection(pl, i, x, True) - self.delta1[pl][i])/EI
        return th

    def axial(self, i, x):
        return self.P1['z'][i] + (self.p2[i] - self.p1[i])/(2*self._h(i))*x**2 + self.p1[i]*x

    def axial_deflection(self, pl, i, x):
        return self.dx1[pl][i] - 1/self.EA*(self.P1[pl][i]*x + self.p1[i]*x**2/2 + (self.p2[i] - self.p1[i])*x**3/(6*self._h(i)))

    def torque(self, i, x):
        return self.T1[i] + 0*x

    # ---- whole member -----------------------------------------------------------------------------------------
    def find(self, x):
        """Row of the segment holding station x (stations compared at 10 decimals; x = L belongs to the last one)."""
        for i in range(len(self.x1)):
            if _r10(x) >= _r10(self.x1[i]) and _r10(x) < _r10(self.x2[i]):
                return i
        if isclose(x, self.L):
            return len(self.x1) - 1
        return None

    def value(self, what, pl, x):
        i = self.find(x)
        if i is None:
            return None
        return float(self._eval(what, pl, i, x - self.x1[i]))

    def _eval(self, what, pl, i, x):
        if what == 'shear':
            return self.shear(pl, i, x)
        if what == 'moment':
            return self.moment(pl, i, x)
        if what == 'deflection':
            return self.deflection(pl, i, x)
        if what == 'axial':
            return self.axial(i, x)
        if what == 'axial_deflection':
            return self.axial_deflection('z', i, x)
        if what == 'torque':
            return self.torque(i, x)
        raise ValueError(f'Unsupported result_name: {what}')

    def sample(self, what, pl, xs):
        """(2, N) array of stations and values (`_extract_vector_results`): plain comparisons, last segment closed."""
        xs = np.asarray(xs, dtype=float)
        n = len(self.x1)
        row = np.clip(np.searchsorted(self.x2, xs, side='right'), 0, n - 1)
        ok = (xs >= self.x1[0]) & (xs <= self.x2[-1])
        xs, row = xs[ok], row[ok]
        out = np.empty(len(xs))
        for i in np.unique(row):
            sel = row == i
            out[sel] = self._eval(what, pl, i, xs[sel] - self.x1[i])
        return np.vstack((xs, out))

    def extreme(self, what, pl, largest):
        pick = max if largest else min
        best = None
        for i in range(len(self.x1)):
            h = self._h(i)
            if what == 'torque':
                cands = [self.T1[i]]
            elif what == 'shear':
                w1, w2 = self.w1[pl][i], self.w2[pl][i]
                x0 = w1*h/(w1 - w2) if w1 - w2 != 0 else 0
                cands = [self.shear(pl, i, x) for x in (self._inside(x0, h), 0, h)]
            elif what == 'axial':
                p1, p2 = self.p1[i], self.p2[i]
                x0 = h*p1/(p1 - p2) if p1 - p2 != 0 else 0
                cands = [self.axial(i, x) for x in (self._inside(x0, h), 0, h)]
            else:                                            # moment: zeros of the shear polynomial
                s = _SIGMA[pl]
                a, b, c = -s*(self.w2[pl][i] - self.w1[pl][i])/(2*h), -s*self.w1[pl][i], -s*self.V1[pl][i]
                if a == 0:
                    xa, xb = (-c/b if b != 0 else 0), 0
                elif b*b - 4*a*c < 0:
                    xa = xb = 0
                else:
                    root = (b*b - 4*a*c)**0.5
                    xa, xb = (-b + root)/(2*a), (-b - root)/(2*a)
                pd = self.p_delta and not (pl == 'y' and largest)
                cands = [self.moment(pl, i, x, pd) for x in (self._inside(xa, h), self._inside(xb, h), 0, h)]
            m = pick(cands)
            best = m if best is None else pick(best, m)
        return float(best)

    @staticmethod
    def _inside(x, h):
        return 0 if _r10(x) < 0 or _r10(x) > _r10(h) else x


# ---------------------------------------------------------------------------------------------
# queries on the member objects (mixed into `Member3D` / `PhysMember`)
# ---------------------------------------------------------------------------------------------
_PLANE = {'Fy': 'z', 'Fz': 'y', 'Mz': 'z', 'My': 'y', 'dy': 'z', 'dz': 'y'}


def _plane(direction, allowed, label):
    if direction not in allowed:
        raise ValueError(f"Direction must be {label}. {direction} was given.")
    return _PLANE.get(direction, 'z')


def _combo_names(member, combo_tags):
    if isinstance(combo_tags, str):
        return [combo_tags]
    return [name for name, combo in member.model.load_combos.items()
            if combo.combo_tags is not None and any(tag in combo.combo_tags for tag in combo_tags)]


def plot_diagram(member, array_method, direction, label, combo_name, n_points, figsize):
    """The reference's `plot_*` methods (`Member3D.py:1354-1412` and siblings): one combination as a line, a list of
    combination tags as the max / min envelope of the matching combinations.  matplotlib is imported here, on use."""
    from matplotlib import pyplot as plt
    names = _combo_names(member, combo_name)
    args = () if direction is None else (direction,)
    fig, ax = plt.subplots(figsize=figsize)
    ax.axhline(0, color='black', lw=1)
    ax.grid()
    if len(names) == 1:
        x, y = getattr(member, array_method)(*args, n_points, names[0])
        ax.plot(x, y)
        ax.set_title('Member ' + member.name + '\n' + names[0])
    else:
        high = low = None
        for name in names:
            x, y = getattr(member, array_method)(*args, n_points, name)
            high = y if high is None else np.maximum(high, y)
            low = y if low is None else np.minimum(low, y)
        ax.plot(x, high, color='green', lw=2, label='Max Envelope')
        ax.plot(x, low, color='red', lw=2, label='Min Envelope')
        ax.legend(fontsize='small')
        ax.set_title('Member ' + member.name + '\nEnvelope')
    ax.set_ylabel(label)
    ax.set_xlabel('Location')
    plt.show()


class _Plots:
    """`plot_shear / plot_moment / plot_torque / plot_axial / plot_deflection / plot_rel_deflection`, shared by the
    sub-member and the physical member (each plots its own `*_array`)."""

    def plot_shear(self, Direction, combo_name='Combo 1', n_points=20, figsize=(7, 3)):
        plot_diagram(self, 'shear_array', Direction, 'Shear', combo_name, n_points, figsize)

    def plot_moment(self, Direction, combo_name='Combo 1', n_points=20, figsize=(7, 3)):
        plot_diagram(self, 'moment_array', Direction, 'Moment', combo_name, n_points, figsize)

    def plot_torque(self, combo_name='Combo 1', n_points=20, figsize=(7, 3)):
        plot_diagram(self, 'torque_array', None, 'Torsional Moment (Warping Torsion Not Included)', combo_name, n_points, figsize)

    def plot_axial(self, combo_name='Combo 1', n_points=20, figsize=(7, 3)):
        plot_diagram(self, 'axial_array', None, 'Axial Force', combo_name, n_points, figsize)

    def plot_deflection(self, Direction, combo_name='Combo 1', n_points=20, figsize=(7, 3)):
        plot_diagram(self, 'deflection_array', Direction, 'Deflection', combo_name, n_points, figsize)

    def plot_rel_deflection(self, Direction, combo_name='Combo 1', n_points=20, figsize=(7, 3)):
        plot_diagram(self, 'rel_deflection_array', Direction, 'Relative Deflection', combo_name, n_points, figsize)


class SubMemberDiagrams(_Plots):
    """Result queries of an analysis sub-member (`Member3D.py:1093-2834`)."""

    def T(self):
        """12 x 12 transformation matrix (`Member3D.T`, `:933-1036`): the direction cosines four times on the diagonal."""
        from pynite_b200.analysis import block_T
        return block_T(self.model._member_dircos(self), 4)

    def D(self, combo_name='Combo 1'):
        """Global end displacements (12, 1); an inactive member reports no global X translation (`:1093-1127`)."""
        Di = self.model._D[combo_name][self.i_node.ID*6:self.i_node.ID*6 + 6, 0]
        Dj = self.model._D[combo_name][self.j_node.ID*6:self.j_node.ID*6 + 6, 0]
        D = np.concatenate((Di, Dj)).reshape(12, 1).copy()
        if not self.active[combo_name] == True:             # noqa: E712
            D[0, 0] = D[6, 0] = 0.0
        return D

    def d(self, combo_name='Combo 1'):
        """Local end displacements `T D` (`:920-931`)."""
        return self.T() @ self.D(combo_name)

    def _inactive_local_disp(self, combo_name='Combo 1'):
        Di = self.model._D[combo_name][self.i_node.ID*6:self.i_node.ID*6 + 6, 0]
        Dj = self.model._D[combo_name][self.j_node.ID*6:self.j_node.ID*6 + 6, 0]
        return self.T() @ np.concatenate((Di, Dj)).reshape(12, 1)

    def _p_delta(self):
        return self.model.solution == 'P-Delta'

    def _table(self, combo_name):
        """The segment table of a combination, rebuilt when the stored displacements or the solution type change."""
        key = (combo_name, id(self.model._D.get(combo_name)), self.model.solution)
        cache = self.__dict__.setdefault('_tables', {})
        if key not in cache:
            cache.clear()
            combo = self.model.load_combos[combo_name]
            R = np.asarray(self.model._member_dircos(self), dtype=float)
            pts, dists = local_loads(self, combo, R)
            stations = [load[2] for load in self.PtLoads] + [s for load in self.DistLoads for s in (load[3], load[4])]
            cache[key] = SegmentTable(self.L(), self.material.E, self.section.A, self.section.Iy, self.section.Iz,
                                      self.f(combo_name), self.d(combo_name), pts, dists, stations, self._p_delta())
        return cache[key]

    def _point(self, what, direction, x, combo_name):
        if not self.active[combo_name]:
            return 0
        return self._table(combo_name).value(what, _PLANE.get(direction, 'z'), x)

    def _extreme(self, what, direction, combo_tags, largest):
        best, governing = None, None
        for name in _combo_names(self, combo_tags):
            if not self.active.get(name, False):
                continue
            if what == 'deflection':
                v = self.deflection(direction, 0, name)
                for i in range(100):
                    d = self.deflection(direction, self.L()*i/99, name)
                    if (d > v) if largest else (d < v):
                        v = d
            else:
                v = self._table(name).extreme(what, _PLANE.get(direction, 'z'), largest)
            if best is None or ((v > best) if largest else (v < best)):
                best, governing = v, name
        if best is None:
            best = 0
        return (best, governing) if isinstance(combo_tags, list) else best

    def _array(self, what, direction, n_points, combo_name, x_array):
        L = self.L()
        if x_array is None:
            x_array = np.linspace(0, L, n_points)
        elif any(x_array < 0) or any(x_array > L):
            raise ValueError(f"All x values must be in the range 0 to {L}")
        x_array = np.asarray(x_array, dtype=float)
        pl = _PLANE.get(direction, 'z')
        if what == 'deflection' and not self.active[combo_name]:
            d = self._inactive_local_disp(combo_name)
            k = {'dx': 0, 'dy': 1, 'dz': 2}[direction]
            return np.array([x_array, d[k, 0] + (d[6 + k, 0] - d[k, 0])*x_array/L])
        table = self._table(combo_name)
        if table.p_delta and what in ('moment', 'deflection'):
            fn = self.moment if what == 'moment' else self.deflection
            return np.array([x_array, np.array([fn(direction, x, combo_name) for x in x_array])])
        return table.sample('axial_deflection' if direction == 'dx' else what, pl, x_array)

    # -- the reference's method names ---------------------------------------------------------------------------
    def shear(self, Direction, x, combo_name='Combo 1'):
        return self._point('shear', Direction, x, combo_name)

    def max_shear(self, Direction, combo_tags='Combo 1'):
        return self._extreme('shear', Direction, combo_tags, True)

    def min_shear(self, Direction, combo_tags='Combo 1'):
        return self._extreme('shear', Direction, combo_tags, False)

    def shear_array(self, Direction, n_points, combo_name='Combo 1', x_array=None):
        _plane(Direction, ('Fy', 'Fz'), "'Fy' or 'Fz'")
        return self._array('shear', Direction, n_points, combo_name, x_array)

    def moment(self, Direction, x, combo_name='Combo 1'):
        if self.active[combo_name]:
            _plane(Direction, ('My', 'Mz'), "'My' or 'Mz'")
        return self._point('moment', Direction, x, combo_name)

    def max_moment(self, Direction, combo_tags='Combo 1'):
        return self._extreme('moment', Direction, combo_tags, True)

    def min_moment(self, Direction, combo_tags='Combo 1'):
        return self._extreme('moment', Direction, combo_tags, False)

    def moment_array(self, Direction, n_points, combo_name='Combo 1', x_array=None):
        _plane(Direction, ('My', 'Mz'), "'My' or 'Mz'")
        return self._array('moment', Direction, n_points, combo_name, x_array)

    def torque(self, x, combo_name='Combo 1'):
        return self._point('torque', None, x, combo_name)

    def max_torque(self, combo_tags='Combo 1'):
        return self._extreme('torque', None, combo_tags, True)

    def min_torque(self, combo_tags='Combo 1'):
        return self._extreme('torque', None, combo_tags, False)

    def torque_array(self, n_points, combo_name='Combo 1', x_array=None):
        return self._array('torque', None, n_points, combo_name, x_array)

    def axial(self, x, combo_name='Combo 1'):
        return self._point('axial', None, x, combo_name)

    def max_axial(self, combo_tags='Combo 1'):
        return self._extreme('axial', None, combo_tags, True)

    def min_axial(self, combo_tags='Combo 1'):
        return self._extreme('axial', None, combo_tags, False)

    def axial_array(self, n_points, combo_name='Combo 1', x_array=None):
        return self._array('axial', None, n_points, combo_name, x_array)

    def deflection(self, Direction, x, combo_name='Combo 1'):
        if not self.active[combo_name]:
            d = self._inactive_local_disp(combo_name)
            k = {'dx': 0, 'dy': 1, 'dz': 2}.get(Direction)
            if k is None:
                raise ValueError(f"Direction must be 'dx', 'dy' or 'dz'. {Direction} was given.")
            return d[k, 0] + (d[6 + k, 0] - d[k, 0])*x/self.L()
        what = 'axial_deflection' if Direction == 'dx' else 'deflection'
        return self._table(combo_name).value(what, _PLANE.get(Direction, 'z'), x)

    def max_deflection(self, Direction, combo_tags='Combo 1'):
        return self._extreme('deflection', Direction, combo_tags, True)

    def min_deflection(self, Direction, combo_tags='Combo 1'):
        return self._extreme('deflection', Direction, combo_tags, False)

    def deflection_array(self, Direction, n_points, combo_name='Combo 1', x_array=None):
        _plane(Direction, ('dx', 'dy', 'dz'), "'dx', 'dy' or 'dz'")
        return self._array('deflection', Direction, n_points, combo_name, x_array)

    def rel_deflection(self, Direction, x, combo_name='Combo 1'):
        """Deflection relative to the chord between the two ends (`:2668-2738`; 'dy' and 'dz' only, as there)."""
        if not self.active[combo_name]:
            return 0
        if Direction not in ('dy', 'dz'):
            return None
        d = self.d(combo_name)
        k = 1 if Direction == 'dy' else 2
        table = self._table(combo_name)
        i = table.find(x)
        if i is None:
            return None
        v = float(table.deflection(_PLANE[Direction], i, x - table.x1[i]))
        if _r10(x) < _r10(table.x2[i]):
            return v - (d[k, 0] + (d[6 + k, 0] - d[k, 0])/self.L()*x)
        return v - d[6 + k, 0]

    def rel_deflection_array(self, Direction, n_points, combo_name='Combo 1', x_array=None):
        L = self.L()
        xs = np.linspace(0, L, n_points) if x_array is None else np.asarray(x_array, dtype=float)
        return np.array([xs, np.array([self.rel_deflection(Direction, x, combo_name) for x in xs])])


class PhysMemberDiagrams(_Plots):
    """The same queries on a physical member: each is answered by the sub-member the station lies on, extremes run over
    all sub-members, arrays are stitched in member coordinates (`PhysMember.py:202-1404`)."""

    def find_member(self, x):
        """(sub-member, station on it) for a station on the physical member (`PhysMember.py:1380-1404`)."""
        end = 0
        subs = list(self.sub_members.values())
        for i, member in enumerate(subs):
            end += member.L()
            if x < end or (isclose(x, end) and i == len(subs) - 1):
                return member, x - (end - member.L())
        raise ValueError(f"Location x={x} does not lie on this member")

    def _over_subs(self, method, args, combo_tags, largest):
        best, governing = None, None
        for name in _combo_names(self, combo_tags):
            for member in self.sub_members.values():
                v = getattr(member, method)(*args, name)
                if best is None or ((v > best) if largest else (v < best)):
                    best, governing = v, name
        if best is None:
            best = 0
        return (best, governing) if isinstance(combo_tags, list) else best

    def _stitch(self, method, args, n_points, combo_name, x_array):
        L = self.L()
        if x_array is None:
            x_array = np.linspace(0, L, n_points)
        elif any(x_array < 0) or any(x_array > L):
            raise ValueError(f"All x values must be in the range 0 to {L}")
        x_array = np.asarray(x_array, dtype=float)
        parts, start = [], 0
        subs = list(self.sub_members.values())
        for i, sub in enumerate(subs):
            end = start + sub.L()
            inside = (x_array >= start) & ((x_array <= end) if i == len(subs) - 1 else (x_array < end))
            if inside.any():
                part = getattr(sub, method)(*args, None, combo_name, x_array[inside] - start)
                part[0] = part[0] + start
                parts.append(part)
            start = end
        return np.hstack(parts) if parts else np.empty((2, 0))

    def shear(self, Direction, x, combo_name='Combo 1'):
        member, xm = self.find_member(x)
        return member.shear(Direction, xm, combo_name)

    def max_shear(self, Direction, combo_tags='Combo 1'):
        return self._over_subs('max_shear', (Direction,), combo_tags, True)

    def min_shear(self, Direction, combo_tags='Combo 1'):
        return self._over_subs('min_shear', (Direction,), combo_tags, False)

    def shear_array(self, Direction, n_points, combo_name='Combo 1', x_array=None):
        return self._stitch('shear_array', (Direction,), n_points, combo_name, x_array)

    def moment(self, Direction, x, combo_name='Combo 1'):
        member, xm = self.find_member(x)
        return member.moment(Direction, xm, combo_name)

    def max_moment(self, Direction, combo_tags='Combo 1'):
        return self._over_subs('max_moment', (Direction,), combo_tags, True)

    def min_moment(self, Direction, combo_tags='Combo 1'):
        return self._over_subs('min_moment', (Direction,), combo_tags, False)

    def moment_array(self, Direction, n_points, combo_name='Combo 1', x_array=None):
        return self._stitch('moment_array', (Direction,), n_points, combo_name, x_array)

    def torque(self, x, combo_name='Combo 1'):
        member, xm = self.find_member(x)
        return member.torque(xm, combo_name)

    def max_torque(self, combo_tags='Combo 1'):
        return self._over_subs('max_torque', (), combo_tags, True)

    def min_torque(self, combo_tags='Combo 1'):
        return self._over_subs('min_torque', (), combo_tags, False)

    def torque_array(self, n_points, combo_name='Combo 1', x_array=None):
        return self._stitch('torque_array', (), n_points, combo_name, x_array)

    def axial(self, x, combo_name='Combo 1'):
        member, xm = self.find_member(x)
        return member.axial(xm, combo_name)

    def max_axial(self, combo_tags='Combo 1'):
        return self._over_subs('max_axial', (), combo_tags, True)

    def min_axial(self, combo_tags='Combo 1'):
        return self._over_subs('min_axial', (), combo_tags, False)

    def axial_array(self, n_points, combo_name='Combo 1', x_array=None):
        return self._stitch('axial_array', (), n_points, combo_name, x_array)

    def deflection(self, Direction, x, combo_name='Combo 1'):
        member, xm = self.find_member(x)
        return member.deflection(Direction, xm, combo_name)

    def max_deflection(self, Direction, combo_tags='Combo 1'):
        return self._over_subs('max_deflection', (Direction,), combo_tags, True)

    def min_deflection(self, Direction, combo_tags='Combo 1'):
        return self._over_subs('min_deflection', (Direction,), combo_tags, False)

    def rel_deflection(self, Direction, x, combo_name='Combo 1'):
        member, xm = self.find_member(x)
        return member.rel_deflection(Direction, xm, combo_name)

    def deflection_array(self, Direction, n_points, combo_name='Combo 1', x_array=None):
        return self._stitch('deflection_array', (Direction,), n_points, combo_name, x_array)
