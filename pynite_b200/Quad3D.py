"""Four-node shell element records: DKMQ quadrilateral and rectangular Kirchhoff plate.

Data fields follow `Pynite/Quad3D.py:30-57` and `Pynite/Plate3D.py:16-72` (nodes i-j-m-n, thickness,
E and nu copied from the material at construction, in-plane stiffness modifiers, pressure list).
Stiffness, fixed-end forces and end forces (`Quad3D.py:107-950`, `Plate3D.py:77-522`) are computed by
the sm_100a shell kernels (`csrc/elem_quad.cuh`, `csrc/elem_plate.cuh`).
"""
from __future__ import annotations


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

    def F(self, combo_name='Combo 1'):
        """Global end force vector (24, 1) from the last GPU solve."""
        return self.model._element_force(self._kind, self.ID, combo_name, local=False)

    def f(self, combo_name='Combo 1'):
        """Local end force vector (24, 1) from the last GPU solve."""
        return self.model._element_force(self._kind, self.ID, combo_name, local=True)


class Quad3D(_Shell4):
    _kind = 'quad'
    type = 'Quad'

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
