"""Material and section property records (data only on the assembly/solve path).

Mirrors the fields the reference path reads: `Pynite/Material.py:13-38` (E, G, nu, rho, fy) and
`Pynite/Section.py:17-37` (A, Iy, Iz, J).  `SteelSection` adds the plastic moduli, the stress ratio `Phi` and its gradient `G`
(`Section.py:84-207`); their consumer, the pushover driver, is out of scope.
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

    def Phi(self, fx=0, my=0, mz=0):
        """Stress ratio of the cross-section (`Section.py:42-49`): defined by the subclasses."""
        raise NotImplementedError("Phi method must be implemented in subclasses.")

    def G(self, fx, my, mz):
        """Gradient of `Phi` by central differences, (3, 1) (`Section.py:51-82`)."""
        import numpy as np
        h = 1e-6
        return np.array([[(self.Phi(fx + h, my, mz) - self.Phi(fx - h, my, mz))/(2*h)],
                         [(self.Phi(fx, my + h, mz) - self.Phi(fx, my - h, mz))/(2*h)],
                         [(self.Phi(fx, my, mz + h) - self.Phi(fx, my, mz - h))/(2*h)]])


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

    def G(self, fx, my, mz):
        """Gradient of the yield surface on the member's six end actions, (6, 1): zero while elastic, else the analytic
        partial derivatives of `Phi` with respect to |fx|, |my|, |mz| (`Section.py:154-205`)."""
        import numpy as np
        if self.Phi(fx, my, mz) < 1.0:
            return np.zeros((6, 1), dtype=int)
        fy = self.material.fy
        Py, Mpy, Mpz = fy*self.A, fy*self.Zy, fy*self.Zz
        p, m_y, m_z = abs(fx)/Py, abs(my)/Mpy, abs(mz)/Mpz
        d_p = (2*p + 7.0*p*m_z**2 + 18*p**5*m_y**2)/Py
        d_my = (4*m_y**3 + 6*p**6*m_y + 9.0*m_z**4*m_y)/Mpy
        d_mz = (2*m_z + 7.0*p**2*m_z + 18.0*m_z**3*m_y**2)/Mpz
        return np.array([[d_p], [0], [0], [0], [d_my], [d_mz]])
