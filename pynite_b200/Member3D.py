"""Frame member records: analysis sub-members and the user-facing physical member.

`Member3D` carries what `Pynite/Member3D.py:37-105` stores for the stiffness path (nodes, material,
section, roll angle, loads, releases, tension/compression-only flags, per-combo activity).  All
stiffness math (`_ke_unc` :176, `ke` :147, `kg` :210, `T` :933, `Ke` :1038, `fer` :742, `f` :872)
runs in the sm_100a member kernels (`csrc/elem_line.cuh`, `csrc/pn_elements.cu`); `f()/F()` here read the end forces the
GPU solve produced, and the result diagrams (`shear/moment/axial/torque/deflection`, their extremes and arrays,
`Member3D.py:1163-2834`, `PhysMember.py:202-1404`) are evaluated from them in `diagrams.py`.

`PhysMember.descritize` restates `Pynite/PhysMember.py:37-200`: a physical member is split at every
model node lying strictly inside it (perpendicular distance <= 1e-12 (1+L)), sub-members are named
name+'a', 'b', ... by distance along the member, end releases go to the end sub-members, and
distributed / point loads are clipped onto the sub-members.
"""
from __future__ import annotations

from math import isclose

import numpy as np

from pynite_b200.diagrams import PhysMemberDiagrams, SubMemberDiagrams


class Member3D(SubMemberDiagrams):

    def __init__(self, model, name, i_node, j_node, material_name, section_name,
                 rotation=0.0, tension_only=False, comp_only=False):
        self.name = name
        self.ID = None
        self.i_node = i_node
        self.j_node = j_node
        try:
            self.material = model.materials[material_name]
        except KeyError:
            raise NameError(f"No material named '{material_name}'")
        try:
            self.section = model.sections[section_name]
        except KeyError:
            raise NameError(f"No section names '{section_name}'")
        self.rotation = rotation
        self.PtLoads = []        # (direction, P, x, case)
        self.DistLoads = []      # (direction, w1, w2, x1, x2, case, self_weight)
        self.Releases = [False]*12
        self.tension_only = tension_only
        self.comp_only = comp_only
        self.active = {}         # {combo name: bool}
        self._solved_combo = None
        self.model = model

    def __repr__(self):
        return f"Member3D(name={self.name!r}, i_node={self.i_node.name!r}, j_node={self.j_node.name!r})"

    def L(self):
        return self.i_node.distance(self.j_node)

    def f(self, combo_name='Combo 1'):
        """Local end forces (12, 1) for a combo, from the last GPU solve (`Member3D.py:872-918`)."""
        return self.model._element_force('member', self.ID, combo_name, local=True)

    def F(self, combo_name='Combo 1'):
        """Global end forces (12, 1) for a combo (`Member3D.py:1072-1078`)."""
        return self.model._element_force('member', self.ID, combo_name, local=False)

    def _engine_id(self):
        """Row of this member in the engine's member arrays.  A physical member is there through its sub-members: its own
        matrices exist only when it was not split."""
        subs = getattr(self, 'sub_members', None)
        if subs is None:
            return self.ID
        if not subs:
            from pynite_b200.analysis import renumber
            renumber(self.model)
        if len(self.sub_members) != 1:
            raise ValueError(f"member '{self.name}' is split into {len(self.sub_members)} sub-members: query "
                             "`member.sub_members[...]` instead")
        return next(iter(self.sub_members.values())).ID

    def _local_loads(self, combo_name):
        from pynite_b200.diagrams import local_loads
        return local_loads(self, self.model.load_combos[combo_name], np.asarray(self.model._member_dircos(self), dtype=float))

    def _fer_unc(self, combo_name='Combo 1'):
        """Local fixed-end reaction vector (12, 1) ignoring the end releases (`Member3D.py:774-850`)."""
        from pynite_b200.fixed_end import member_fer_unc
        return member_fer_unc(*self._local_loads(combo_name), self.L()).reshape(12, 1)

    def fer(self, combo_name='Combo 1'):
        """Condensed (and expanded) local fixed-end reaction vector (`Member3D.py:742-772`)."""
        from pynite_b200.fixed_end import condense_vector, member_ke_unc
        k = member_ke_unc(self.material.E, self.material.G, self.section.A, self.section.Iy, self.section.Iz, self.section.J,
                          self.L())
        return condense_vector(k, self._fer_unc(combo_name).reshape(12), self.Releases).reshape(12, 1)

    def FER(self, combo_name='Combo 1'):
        """Global fixed-end reaction vector `T^T fer` (`Member3D.py:1080-1091`)."""
        return self.T().T @ self.fer(combo_name)

    def Ke(self):
        """Global stiffness matrix (12, 12), computed by the member kernel (`Member3D.py:1038-1046`)."""
        from pynite_b200.analysis import element_matrix
        return element_matrix(self.model, 'member', self._engine_id())

    def ke(self):
        """Local (condensed) stiffness matrix `T Ke T^T` (`Member3D.py:147-174`)."""
        T = self.T()
        return T @ self.Ke() @ T.T

    def Kg(self, P=0.0):
        """Global geometric stiffness matrix for an axial force P (`Member3D.py:1048-1058`)."""
        from pynite_b200.analysis import element_matrix
        return element_matrix(self.model, 'member', self._engine_id(), 'Kg', P)

    def kg(self, P=0):
        """Local geometric stiffness matrix (`Member3D.py:210-266`)."""
        T = self.T()
        return T @ self.Kg(P) @ T.T

    def m(self, mass_combo_name, mass_direction='Y', gravity=1.0):
        """Condensed local mass matrix (`Member3D.py:605-646`)."""
        from pynite_b200 import modal
        combo = self.model.load_combos[mass_combo_name]
        lm, x = modal.member_load_mass(self, combo, mass_direction, gravity)
        m = modal.member_local_mass(np.array([self.L()]), np.array([self.section.A]), np.array([self.section.J]),
                                    np.array([lm]), np.array([x]),
                                    np.array([modal.member_material_mass(self, combo, gravity)]))[0]
        return modal.condense_released(m, self.Releases)

    def M(self, mass_combo_name=None, mass_direction='Y', gravity=1.0):
        """Global mass matrix `T^T m T` (`Member3D.py:648-672`)."""
        from pynite_b200 import modal
        return modal.member_global_mass(self.model, [self], self.model.load_combos[mass_combo_name], mass_direction,
                                        gravity)[0]


class PhysMember(PhysMemberDiagrams, Member3D):

    def __init__(self, model, name, i_node, j_node, material_name, section_name,
                 rotation=0.0, tension_only=False, comp_only=False):
        super().__init__(model, name, i_node, j_node, material_name, section_name,
                         rotation, tension_only, comp_only)
        self.sub_members = {}

    def __repr__(self):
        return f"PhysMember(name={self.name!r}, i_node={self.i_node.name!r}, j_node={self.j_node.name!r})"

    def descritize(self, node_index=None):
        """Splits the member at interior collinear nodes (see module docstring).

        `node_index` is an optional `flatten.NodeIndex` (coordinate arrays sorted by X) that limits
        the collinearity scan to the member's bounding box; without it every node is scanned.
        """
        self.sub_members = {}
        ni, nj = self.i_node, self.j_node
        Xi, Yi, Zi = ni.X, ni.Y, ni.Z
        Xj, Yj, Zj = nj.X, nj.Y, nj.Z
        L = np.linalg.norm(np.array([Xj - Xi, Yj - Yi, Zj - Zi]))
        if L == 0:
            return
        u = np.array([Xj - Xi, Yj - Yi, Zj - Zi])/L

        stations = [(ni, 0.0), (nj, L)]
        tol = 1e-12*(1.0 + L)
        bb = 1e-9*(1.0 + L)
        lo = (min(Xi, Xj) - bb, min(Yi, Yj) - bb, min(Zi, Zj) - bb)
        hi = (max(Xi, Xj) + bb, max(Yi, Yj) + bb, max(Zi, Zj) + bb)

        if node_index is None:
            from pynite_b200.flatten import NodeIndex
            node_index = NodeIndex(self.model)
        cand, X, Y, Z = node_index.in_box(lo, hi)
        if cand.size:
            dx, dy, dz = X - Xi, Y - Yi, Z - Zi
            t = dx*u[0] + dy*u[1] + dz*u[2]
            px, py, pz = dx - t*u[0], dy - t*u[1], dz - t*u[2]
            d_perp = (px**2 + py**2 + pz**2)**0.5
            keep = (t > 0.0) & (t < L) & (d_perp <= tol)
            nodes = node_index.nodes
            # Model (insertion) order, as the reference scans `model.nodes.values()`.
            for k in np.sort(cand[keep]):
                node = nodes[k]
                if node is ni or node is nj:
                    continue
                sel = np.nonzero(cand == k)[0][0]
                stations.append((node, t[sel]))
        stations.sort(key=lambda s: s[1])

        n_sub = len(stations) - 1
        combos = list(self.model.load_combos.keys())
        for i in range(n_sub):
            sub_name = self.name + chr(i + 97)
            a_node, xi = stations[i]
            b_node, xj = stations[i + 1]
            sub = Member3D(self.model, sub_name, a_node, b_node, self.material.name,
                           self.section.name, self.rotation, self.tension_only, self.comp_only)
            for combo_name in combos:
                sub.active[combo_name] = True
            if i == 0:
                sub.Releases[0:6] = self.Releases[0:6]
            if i == n_sub - 1:
                sub.Releases[6:12] = self.Releases[6:12]

            for direction, w1, w2, x1_load, x2_load, case, self_weight in self.DistLoads:
                if x1_load <= xj and x2_load > xi:
                    span = x2_load - x1_load
                    if x1_load > xi:
                        x1 = x1_load - xi
                    else:
                        x1 = 0
                        w1 = (w2 - w1)/span*(xi - x1_load) + w1
                    if x2_load < xj:
                        x2 = x2_load - xi
                    else:
                        x2 = xj - xi
                        # NOTE: evaluated with the *already clipped* w1, exactly as the reference's
                        # late-binding closure does (`PhysMember.py:165-178`).
                        w2 = (w2 - w1)/span*(xj - x1_load) + w1
                    sub.DistLoads.append([direction, w1, w2, x1, x2, case, self_weight])

            for direction, P, x, case in self.PtLoads:
                if x >= xi and x < xj or (isclose(x, xj) and isclose(xj, self.L())):
                    sub.PtLoads.append([direction, P, x - xi, case])

            self.sub_members[sub_name] = sub
