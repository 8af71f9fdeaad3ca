"""Material and section property records (data only on the assembly/solve path).

Mirrors the fields the reference path reads: `Pynite/Material.py:13-38` (E, G, nu, rho, fy) and
`Pynite/Section.py:17-37` (A, Iy, Iz, J).  The steel plastic-surface methods of the reference
(`Section.py:42-207`) belong to pushover analysis and are out of scope.
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
