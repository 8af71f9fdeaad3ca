"""Four-node shell element records: DKMQ quadrilateral and rectangular Kirchhoff plate.

Data fields follow `Pynite/Quad3D.py:30-57` and `Pynite/Plate3D.py:16-72` (nodes i-j-m-n, thickness,
E and nu copied from the material at construction, in-plane stiffness modifiers, pressure list).
Stiffness, fixed-end forces and end forces (`Quad3D.py:107-950`, `Plate3D.py:77-522`) are computed by
the sm_100a shell kernels (`csrc/elem_shell.cuh`); internal shears, moments and membrane stresses
(`Quad3D.py:1017-1239`, `Plate3D.py:590-692`) by `csrc/elem_stress.cuh` / `pn_stress.cu`, batched over all
elements of the kind for the requested combination and point.
"""
from __future__ import annotations

import numpy as np


class _Shell4:
    _kind = None

    def __init__(self, name, i_node, j_node, m_node, n_node, t, material_name, model,
                 kx_mod=1.0, ky_mod=1.0):
        self.name = name
        self.ID = None
        self.i_node = i_node
        self.j_node = j_node
        self.m_node = m_node
        self.n_node = n_node
        self.t = t
        self.kx_mod = kx_mod
        self.ky_mod = ky_mod
        self.pressures = []      # [pressure, case]
        self.model = model
        try:
            self.E = model.materials[material_name].E
            self.nu = model.materials[material_name].nu
        except Exception:
            raise KeyError('Please define the material ' + str(material_name) +
                           ' before assigning it to plates.')

    def _nodes(self):
        return (self.i_node, self.j_node, self.m_node, self.n_node)

    def T(self):
        """24 x 24 transformation matrix (`Quad3D.T` :884-950, `Plate3D.T` :451-500)."""
        from pynite_b200.analysis import block_T, shell_dircos_host
        return block_T(shell_dircos_host(self), 8)

    def Ke(self):
        """Global stiffness matrix (24, 24), computed by the shell kernel (`Quad3D.Ke` :858-868, `Plate3D.Ke` :502-508)."""
        from pynite_b200.analysis import element_matrix
        if self.ID is None:
            from pynite_b200.analysis import renumber
            renumber(self.model)
        return element_matrix(self.model, self._kind, self.ID)

    def ke(self):
        """Local stiffness matrix `T Ke T^T` (`Quad3D.ke` :107-141, `Plate3D.ke` :77-96)."""
        T = self.T()
        return T @ self.Ke() @ T.T

    def fer(self, combo_name='Combo 1'):
        """Local fixed-end reaction vector (24, 1) of the element's pressures (`Quad3D.fer` :721-788, `Plate3D.fer` :324-393)."""
        from pynite_b200 import fixed_end
        fn = fixed_end.quad_fer if self._kind == 'quad' else fixed_end.plate_fer
        return fn(self, self.model.load_combos[combo_name])

    def FER(self, combo_name='Combo 1'):
        """Global fixed-end reaction vector `T^T fer` (`Quad3D.FER` :870-882, `Plate3D.FER` :510-522)."""
        return self.T().T @ self.fer(combo_name)

    def D(self, combo_name='Combo 1'):
        """Global nodal displacements (24, 1) (`Quad3D.D` :790-829)."""
        from pynite_b200.analysis import element_D
        return element_D(self.model, self._nodes(), combo_name)

    def d(self, combo_name='Combo 1'):
        """Local nodal displacements `T D` (`Quad3D.d` :779-788)."""
        return self.T() @ self.D(combo_name)

    def F(self, combo_name='Combo 1'):
        """Global end force vector (24, 1) from the last GPU solve."""
        return self.model._element_force(self._kind, self.ID, combo_name, local=False)

    def f(self, combo_name='Combo 1'):
        """Local end force vector (24, 1) from the last GPU solve."""
        return self.model._element_force(self._kind, self.ID, combo_name, local=True)


class Quad3D(_Shell4):
    _kind = 'quad'
    type = 'Quad'

    # Results at natural coordinates (xi, eta), extrapolated from the Gauss points like the reference does.
    def shear(self, xi=0.0, eta=0.0, local=True, combo_name='Combo 1'):
        """[[Qx], [Qy]] per unit length (`Quad3D.py:1017-1094`); global: [[Qx], [Qy], [Qz]]."""
        r = self.model._shell_stress('quad', self.ID, combo_name, xi, eta, local)
        return r[0:2].reshape(2, 1).copy() if local else r[0:3].reshape(3, 1).copy()

    def moment(self, xi=0.0, eta=0.0, local=True, combo_name='Combo 1'):
        """[[Mx], [My], [Mxy]] per unit length (`Quad3D.py:1096-1171`)."""
        return self.model._shell_stress('quad', self.ID, combo_name, xi, eta, local)[3:6].reshape(3, 1).copy()

    def membrane(self, xi=0.0, eta=0.0, local=True, combo_name='Combo 1'):
        """[[Sx], [Sy], [Txy]] (`Quad3D.py:1173-1239`)."""
        return self.model._shell_stress('quad', self.ID, combo_name, xi, eta, local)[6:9].reshape(3, 1).copy()

    def __repr__(self):
        return (f"Quad3D(name={self.name!r}, i_node={self.i_node.name!r}, j_node={self.j_node.name!r}, "
                f"m_node={self.m_node.name!r}, n_node={self.n_node.name!r})")


class Plate3D(_Shell4):
    _kind = 'plate'
    type = 'Rect'

    def __repr__(self):
        return (f"Plate3D(name={self.name!r}, i_node={self.i_node.name!r}, j_node={self.j_node.name!r}, "
                f"m_node={self.m_node.name!r}, n_node={self.n_node.name!r})")

    def width(self):
        return self.i_node.distance(self.j_node)

    def height(self):
        return self.i_node.distance(self.n_node)

    # Results at local coordinates (x, y).  Like the reference's, these ignore `local` (`Plate3D.py:590-692`).
    def moment(self, x, y, local=True, combo_name='Combo 1'):
        return self.model._shell_stress('plate', self.ID, combo_name, x, y, True)[3:6].reshape(3, 1).copy()

    def shear(self, x, y, local=True, combo_name='Combo 1'):
        return self.model._shell_stress('plate', self.ID, combo_name, x, y, True)[0:2].reshape(2, 1).copy()

    def membrane(self, x, y, local=True, combo_name='Combo 1'):
        return self.model._shell_stress('plate', self.ID, combo_name, x, y, True)[6:9].reshape(3, 1).copy()
