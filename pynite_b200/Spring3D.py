"""Two-node axial spring record (`Pynite/Spring3D.py:25-43`).

The stiffness (`ke` :59, `T` :114, `Ke` :191) and end forces (`F` :200) are evaluated on the GPU with
the member kernels; the Python object only carries data and reads results back.
"""
from __future__ import annotations


class Spring3D:

    def __init__(self, name, i_node, j_node, ks, LoadCombos=None, tension_only=False, comp_only=False):
        self.name = name
        self.ID = None
        self.i_node = i_node
        self.j_node = j_node
        self.ks = ks
        self.load_combos = {} if LoadCombos is None else LoadCombos
        self.tension_only = tension_only
        self.comp_only = comp_only
        self.active = {}               # {combo name: bool}
        self.model = i_node.model

    def __repr__(self):
        return (f"Spring3D(name={self.name!r}, i_node={self.i_node.name!r}, "
                f"j_node={self.j_node.name!r}, ks={self.ks})")

    def L(self):
        return self.i_node.distance(self.j_node)

    def f(self, combo_name='Combo 1'):
        """Local end force vector (12, 1) from the last GPU solve (`Spring3D.py:86-97`)."""
        return self.model._element_force('spring', self.ID, combo_name, local=True)

    def F(self, combo_name='Combo 1'):
        """Global end force vector (12, 1) from the last GPU solve (`Spring3D.py:200-206`)."""
        return self.model._element_force('spring', self.ID, combo_name, local=False)

    def axial(self, combo_name='Combo 1'):
        """Axial force = first local end force (`Spring3D.py:245-256`)."""
        return self.f(combo_name)[0, 0]
