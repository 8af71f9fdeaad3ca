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

    def T(self):
        """12 x 12 transformation matrix (`Spring3D.py:114-188`): a member's, without a roll angle."""
        from pynite_b200.analysis import block_T, member_dircos_host
        return block_T(member_dircos_host(self), 4)

    def Ke(self):
        """Global stiffness matrix (12, 12), computed by the member kernel (`Spring3D.py:191-197`)."""
        from pynite_b200.analysis import element_matrix, renumber
        if self.ID is None:
            renumber(self.model)
        return element_matrix(self.model, 'spring', self.ID)

    def ke(self):
        """Local stiffness matrix `T Ke T^T` (`Spring3D.py:59-83`)."""
        T = self.T()
        return T @ self.Ke() @ T.T

    def D(self, combo_name='Combo 1'):
        """Global end displacements (12, 1); an inactive spring reports no global X translation (`Spring3D.py:209-243`)."""
        from pynite_b200.analysis import element_D
        D = element_D(self.model, (self.i_node, self.j_node), combo_name)
        if not self.active[combo_name] == True:             # noqa: E712
            D[0, 0] = D[6, 0] = 0.0
        return D

    def d(self, combo_name='Combo 1'):
        """Local end displacements `T D` (`Spring3D.py:100-111`)."""
        return self.T() @ self.D(combo_name)

    def f(self, combo_name='Combo 1'):
        """Local end force vector (12, 1) from the last GPU solve (`Spring3D.py:86-97`)."""
        return self.model._element_force('spring', self.ID, combo_name, local=True)

    def F(self, combo_name='Combo 1'):
        """Global end force vector (12, 1) from the last GPU solve (`Spring3D.py:200-206`)."""
        return self.model._element_force('spring', self.ID, combo_name, local=False)

    def axial(self, combo_name='Combo 1'):
        """Axial force = first local end force (`Spring3D.py:245-256`)."""
        return self.f(combo_name)[0, 0]
