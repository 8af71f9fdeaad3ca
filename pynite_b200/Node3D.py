"""Node record: coordinates, supports, support springs, enforced displacements, result views.

Field names follow `Pynite/Node3D.py:28-85`.  The reference keeps twelve `{combo: float}` dicts per
node; at the sizes this build targets (2.5e5 nodes x 64 combos) that is 2e8 Python objects, so the
twelve attributes here are *views* onto the model-level result arrays written by the GPU solve
(`model._D[combo]`, `model._Rxn[combo]`).  `node.DX['Combo 1']` reads exactly what the reference
would have stored (`Analysis.py:875-880`, `:1087-1313`).
"""
from __future__ import annotations

from collections.abc import MutableMapping

import numpy as np

_DISP = ('DX', 'DY', 'DZ', 'RX', 'RY', 'RZ')
_RXN = ('RxnFX', 'RxnFY', 'RxnFZ', 'RxnMX', 'RxnMY', 'RxnMZ')


class _NodeResultView(MutableMapping):
    """`{combo_name: value}` mapping backed by one component of a model-level (6N, 1) array."""

    __slots__ = ('_node', '_store_name', '_comp')

    def __init__(self, node, store_name, comp):
        self._node = node
        self._store_name = store_name
        self._comp = comp

    def _store(self):
        return getattr(self._node.model, self._store_name)

    def _row(self):
        node_id = self._node.ID
        if node_id is None:
            raise KeyError('node has not been numbered yet (no analysis has been run)')
        return node_id*6 + self._comp

    def __getitem__(self, combo_name):
        return self._store()[combo_name][self._row(), 0]

    def __setitem__(self, combo_name, value):
        store = self._store()
        if combo_name not in store:
            store[combo_name] = np.zeros((len(self._node.model.nodes)*6, 1))
        store[combo_name][self._row(), 0] = value

    def __delitem__(self, combo_name):
        raise TypeError('per-node result entries cannot be deleted individually')

    def __iter__(self):
        return iter(self._store())

    def __len__(self):
        return len(self._store())

    def __repr__(self):
        return repr({k: self[k] for k in self})


class Node3D:

    def __init__(self, model, name, X, Y, Z):
        self.name = name
        self.ID = None
        self.X = X
        self.Y = Y
        self.Z = Z
        self.NodeLoads = []            # (direction, magnitude, case)
        self.support_DX = False
        self.support_DY = False
        self.support_DZ = False
        self.support_RX = False
        self.support_RY = False
        self.support_RZ = False
        # [stiffness, direction ('+', '-' or None), active]
        self.spring_DX = [None, None, None]
        self.spring_DY = [None, None, None]
        self.spring_DZ = [None, None, None]
        self.spring_RX = [None, None, None]
        self.spring_RY = [None, None, None]
        self.spring_RZ = [None, None, None]
        self.EnforcedDX = None
        self.EnforcedDY = None
        self.EnforcedDZ = None
        self.EnforcedRX = None
        self.EnforcedRY = None
        self.EnforcedRZ = None
        self.contour = []
        self.model = model

    def __repr__(self):
        return f"Node3D(name={self.name!r}, X={self.X}, Y={self.Y}, Z={self.Z})"

    def distance(self, other):
        """Euclidean distance to another node (`Pynite/Node3D.py:90-99`)."""
        return ((self.X - other.X)**2 + (self.Y - other.Y)**2 + (self.Z - other.Z)**2)**0.5

    def M(self, mass_combo_name=None, mass_direction='Y', gravity=1.0, characteristic_length=None):
        """6 x 6 nodal mass matrix (`Pynite/Node3D.py:101-151`): the mass of the nodal loads of the mass combination on
        the three translations, and 1 % of m L^2 on the unsupported rotations when a characteristic length is given."""
        from pynite_b200.modal import node_mass
        if mass_combo_name not in self.model.load_combos:
            raise ValueError(f'Load combination {mass_combo_name} not found in the model.')
        m = np.zeros((6, 6))
        total = node_mass(self, self.model.load_combos[mass_combo_name], mass_direction, gravity)
        if total > 0:
            m[0, 0] = m[1, 1] = m[2, 2] = total
            if characteristic_length is not None:
                for k, held in enumerate((self.support_RX, self.support_RY, self.support_RZ)):
                    if not held:
                        m[3 + k, 3 + k] = total*characteristic_length**2*0.01
        return m


def _make_view_property(store_name, comp):
    def getter(self):
        return _NodeResultView(self, store_name, comp)

    def setter(self, value):
        # The reference resets results with `node.DX = {}`; here the arrays live on the model.
        if value:
            view = _NodeResultView(self, store_name, comp)
            for k, v in value.items():
                view[k] = v
    return property(getter, setter)


for _i, _n in enumerate(_DISP):
    setattr(Node3D, _n, _make_view_property('_D', _i))
for _i, _n in enumerate(_RXN):
    setattr(Node3D, _n, _make_view_property('_Rxn', _i))
