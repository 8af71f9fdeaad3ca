"""Material and section property records (data only on the assembly/solve path).

Mirrors the fields the reference path reads: `Pynite/Material.py:13-38` (E, G, nu, rho, fy) and
`Pynite/Section.py:17-37` (A, Iy, Iz, J).  `SteelSection` keeps the elastic record and the stress ratio `Phi`
(`Section.py:84-152`); the yield-surface gradient (`Section.py:154-207`) belongs to pushover analysis and is out of scope.
"""
from __future__ import annotations


class Material:
    __slots__ = ('model', 'name', 'E', 'G', 'nu', 'rho', 'fy')

    def __init__(self, model, name, E, G, nu, rho, fy=None):
        self.model = model
        self.name = name
        self.E = E
        self.G = G
        self.nu = nu
        self.rho = rho
        self.fy = fy

    def __repr__(self):
        return f"Material(name={self.name!r}, E={self.E}, G={self.G}, nu={self.nu})"


class Section:
    __slots__ = ('model', 'name', 'A', 'Iy', 'Iz', 'J')

    def __init__(self, model, name, A, Iy, Iz, J):
        self.model = model
        self.name = name
        self.A = A
        self.Iy = Iy
        self.Iz = Iz
        self.J = J

    def __repr__(self):
        return f"Section(name={self.name!r}, A={self.A}, Iy={self.Iy}, Iz={self.Iz}, J={self.J})"


class SteelSection(Section):
    """`Pynite.Section.SteelSection` (`Section.py:84-152`): a section with plastic moduli and its own material.  The
    assembly / solve path reads only (A, Iy, Iz, J)."""
    __slots__ = ('ry', 'rz', 'Zy', 'Zz', 'material')

    def __init__(self, model, name, A, Iy, Iz, J, Zy, Zz, material_name):
        super().__init__(model, name, A, Iy, Iz, J)
        self.ry = (Iy/A)**0.5
        self.rz = (Iz/A)**0.5
        self.Zy = Zy
        self.Zz = Zz
        self.material = model.materials[material_name]

    def __repr__(self):
        return f"SteelSection(name={self.name!r}, A={self.A}, Iy={self.Iy}, Iz={self.Iz}, J={self.J})"

    def Phi(self, fx=0, my=0, mz=0):
        """Stress ratio of the cross-section, < 1 while elastic (`Section.py:125-152`, McGuire et al. Eq. 10.18)."""
        fy = self.material.fy
        p, m_y, m_z = abs(fx/(fy*self.A)), abs(my/(fy*self.Zy)), abs(mz/(fy*self.Zz))
        return p**2 + m_z**2 + m_y**4 + 3.5*p**2*m_z**2 + 3*p**6*m_y**2 + 4.5*m_z**4*m_y**2
