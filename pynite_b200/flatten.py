"""Flattens the FEModel3D object graph into structure-of-arrays buffers for the C ABI.

This is the single host->device crossing of the path: after `_renumber`/`descritize`
have fixed node IDs and the sub-member list, every quantity the reference reads inside
`FEModel3D.Ke/FER/P` (`FEModel3D.py:1551-1800, 2136-2235`) and `Analysis._partition_D`
(`Analysis.py:1412-1505`) is copied once into contiguous numpy arrays.  Element order is the
reference's COO emission order: node support springs, springs, sub-members (physical-member order,
then 'a', 'b', ...), quads, plates (`FEModel3D.py:1587-1771`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from operator import attrgetter

import numpy as np

PT_KINDS = {'Fx': 0, 'Fy': 1, 'Fz': 2, 'Mx': 3, 'My': 4, 'Mz': 5,
            'FX': 6, 'FY': 7, 'FZ': 8, 'MX': 9, 'MY': 10, 'MZ': 11}
DIST_KINDS = {'Fx': 12, 'Fy': 13, 'Fz': 14, 'FX': 15, 'FY': 16, 'FZ': 17}
_NODE_LOAD_DOF = {'FX': 0, 'FY': 1, 'FZ': 2, 'MX': 3, 'MY': 4, 'MZ': 5}
_DOFS = ('DX', 'DY', 'DZ', 'RX', 'RY', 'RZ')
# one C-level attribute sweep per node / shell instead of one `getattr` per field (2.5e5 nodes x 21 fields otherwise)
_NODE_XYZ = attrgetter('X', 'Y', 'Z')
_NODE_SUPPORTS = attrgetter(*('support_' + d for d in _DOFS))
_NODE_ENFORCED = attrgetter(*('Enforced' + d for d in _DOFS))
_NODE_SPRINGS = attrgetter(*('spring_' + d for d in _DOFS))
_NOTHING_ENFORCED = (None,)*6
_NO_SPRINGS = tuple([None, None, None] for _ in range(6))
_SHELL_NODES = attrgetter('i_node.ID', 'j_node.ID', 'm_node.ID', 'n_node.ID')
_SHELL_PROPS = attrgetter('t', 'E', 'nu', 'kx_mod', 'ky_mod')


class NodeIndex:
    """Node coordinates sorted by X, for bounding-box candidate queries in `descritize`."""

    def __init__(self, model):
        self.nodes = list(model.nodes.values())
        n = len(self.nodes)
        self.X = np.fromiter((nd.X for nd in self.nodes), dtype=float, count=n)
        self.Y = np.fromiter((nd.Y for nd in self.nodes), dtype=float, count=n)
        self.Z = np.fromiter((nd.Z for nd in self.nodes), dtype=float, count=n)
        self.order = np.argsort(self.X, kind='stable')
        self.Xs = self.X[self.order]

    def in_box(self, lo, hi):
        """Indices (model order positions) and coordinates of nodes inside the closed box."""
        a = np.searchsorted(self.Xs, lo[0], side='left')
        b = np.searchsorted(self.Xs, hi[0], side='right')
        cand = self.order[a:b]
        Y = self.Y[cand]
        Z = self.Z[cand]
        keep = (Y >= lo[1]) & (Y <= hi[1]) & (Z >= lo[2]) & (Z <= hi[2])
        cand = cand[keep]
        return cand, self.X[cand], self.Y[cand], self.Z[cand]


@dataclass
class ModelArrays:
    """SoA image of a model; every array is C-contiguous with the dtype the C ABI expects."""
    n_nodes: int = 0
    xyz: np.ndarray = None            # (N, 3) f64
    support: np.ndarray = None        # (N,)  u8, bit k = DOF k supported
    enf_mask: np.ndarray = None       # (N,)  u8, bit k = DOF k has an enforced displacement
    enf_val: np.ndarray = None        # (N, 6) f64
    nspring_k: np.ndarray = None      # (N, 6) f64
    nspring_flag: np.ndarray = None   # (N, 6) u8: bit0 defined, bit1 active, bit2 dir '+', bit3 dir '-'
    spring_ij: np.ndarray = None      # (S, 2) i32
    spring_ks: np.ndarray = None      # (S,) f64
    spring_active: np.ndarray = None  # (S,) u8
    mem_ij: np.ndarray = None         # (M, 2) i32
    mem_props: np.ndarray = None      # (M, 6) f64: E, G, A, Iy, Iz, J
    mem_rot: np.ndarray = None        # (M,) f64 degrees
    mem_rel: np.ndarray = None        # (M,) u16, bit k = local DOF k released
    mem_active: np.ndarray = None     # (M,) u8
    quad_ijmn: np.ndarray = None      # (Q, 4) i32
    quad_props: np.ndarray = None     # (Q, 5) f64: t, E, nu, kx_mod, ky_mod
    plate_ijmn: np.ndarray = None     # (P, 4) i32
    plate_props: np.ndarray = None    # (P, 5) f64
    # loads
    cases: list = field(default_factory=list)
    nl_dof: np.ndarray = None         # nodal loads: global dof (i32), case (i32), value (f64)
    nl_case: np.ndarray = None
    nl_val: np.ndarray = None
    ml_member: np.ndarray = None      # member loads: sub-member index, case, kind, params (K, 4)
    ml_case: np.ndarray = None
    ml_kind: np.ndarray = None
    ml_params: np.ndarray = None
    qp_elem: np.ndarray = None        # quad pressures: element, case, value
    qp_case: np.ndarray = None
    qp_val: np.ndarray = None
    pp_elem: np.ndarray = None        # plate pressures
    pp_case: np.ndarray = None
    pp_val: np.ndarray = None
    combo_names: list = field(default_factory=list)
    factors: np.ndarray = None        # (n_combo, n_case) f64
    # bookkeeping back to the objects
    sub_members: list = field(default_factory=list)
    sub_owner: np.ndarray = None      # (M,) i32 index of the physical member

    @property
    def n_dof(self):
        return 6*self.n_nodes


def _i32(seq, shape=None):
    a = np.asarray(seq, dtype=np.int32)
    if shape is not None:
        a = a.reshape(shape)
    return np.ascontiguousarray(a)


def _f64(seq, shape=None):
    a = np.asarray(seq, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return np.ascontiguousarray(a)


def flatten_model(model, combo_list, active_combo=None):
    """Builds a `ModelArrays` for `model` and the given list of LoadCombo objects.

    `active_combo`: name of the combo whose `active[...]` flags gate springs / members (the
    reference's `Ke(combo_name)` argument, `FEModel3D.py:1696,1717`).  None = first combo.
    """
    A = ModelArrays()
    nodes = list(model.nodes.values())
    N = len(nodes)
    A.n_nodes = N
    if active_combo is None:
        active_combo = combo_list[0].name if combo_list else None

    A.xyz = _f64(list(map(_NODE_XYZ, nodes)), (N, 3))
    enf_mask = np.zeros(N, dtype=np.uint8)
    enf_val = np.zeros((N, 6))
    nsk = np.zeros((N, 6))
    nsf = np.zeros((N, 6), dtype=np.uint8)
    sup = np.array(list(map(_NODE_SUPPORTS, nodes)), dtype=bool).reshape(N, 6)         # (None counts as unsupported)
    support = (sup.astype(np.uint8) << np.arange(6, dtype=np.uint8)).sum(axis=1).astype(np.uint8)
    # enforced displacements and support springs are rare: only the nodes that differ from a fresh node are walked
    for i, enforced in enumerate(map(_NODE_ENFORCED, nodes)):
        if enforced != _NOTHING_ENFORCED:
            for k, e in enumerate(enforced):
                if e is not None:
                    enf_mask[i] |= (1 << k)
                    enf_val[i, k] = e
    for i, node_springs in enumerate(map(_NODE_SPRINGS, nodes)):
        if node_springs != _NO_SPRINGS:
            for k, sp in enumerate(node_springs):
                if sp[0] is not None:
                    nsk[i, k] = float(sp[0])
                    flag = 1
                    if sp[2] == True:   # noqa: E712  (the reference tests `== True`)
                        flag |= 2
                    if sp[1] == '+':
                        flag |= 4
                    elif sp[1] == '-':
                        flag |= 8
                    nsf[i, k] = flag
    A.support, A.enf_mask, A.enf_val = support, enf_mask, np.ascontiguousarray(enf_val)
    A.nspring_k, A.nspring_flag = np.ascontiguousarray(nsk), np.ascontiguousarray(nsf)

    springs = list(model.springs.values())
    A.spring_ij = _i32([(s.i_node.ID, s.j_node.ID) for s in springs], (len(springs), 2))
    A.spring_ks = _f64([s.ks for s in springs], (len(springs),))
    A.spring_active = np.asarray([1 if s.active.get(active_combo, True) == True else 0 for s in springs],  # noqa: E712
                                 dtype=np.uint8)

    subs, owner, active = [], [], []
    for pi, pm in enumerate(model.members.values()):
        is_active = 1 if pm.active.get(active_combo, True) == True else 0   # noqa: E712
        for sub in pm.sub_members.values():
            subs.append(sub)
            owner.append(pi)
            active.append(is_active)
    M = len(subs)
    A.sub_members = subs
    A.sub_owner = _i32(owner, (M,))
    A.mem_ij = _i32([(m.i_node.ID, m.j_node.ID) for m in subs], (M, 2))
    A.mem_props = _f64([(m.material.E, m.material.G, m.section.A, m.section.Iy, m.section.Iz, m.section.J)
                        for m in subs], (M, 6))
    A.mem_rot = _f64([m.rotation for m in subs], (M,))
    rel = np.zeros(M, dtype=np.uint16)
    for i, m in enumerate(subs):
        r = 0
        for k, flag in enumerate(m.Releases):
            if flag:
                r |= (1 << k)
        rel[i] = r
    A.mem_rel = rel
    A.mem_active = np.asarray(active, dtype=np.uint8).reshape(M)

    quads = list(model.quads.values())
    plates = list(model.plates.values())
    A.quad_ijmn = _i32(list(map(_SHELL_NODES, quads)), (len(quads), 4))
    A.quad_props = _f64(list(map(_SHELL_PROPS, quads)), (len(quads), 5))
    A.plate_ijmn = _i32(list(map(_SHELL_NODES, plates)), (len(plates), 4))
    A.plate_props = _f64(list(map(_SHELL_PROPS, plates)), (len(plates), 5))

    # ---- load cases, combos ------------------------------------------------------------------
    case_set = {}
    for combo in combo_list:
        for case in combo.factors:
            case_set.setdefault(case, None)
    nl, ml, qp, pp = [], [], [], []
    for nd in nodes:
        for direction, mag, case in nd.NodeLoads:
            idx = _NODE_LOAD_DOF.get(direction.upper())
            if idx is None:
                continue
            case_set.setdefault(case, None)
            nl.append((nd.ID*6 + idx, case, mag))
    for mi, m in enumerate(subs):
        for load in m.PtLoads:
            direction, P, x, case = load[0], load[1], load[2], load[3]
            if direction not in PT_KINDS:
                raise Exception('Invalid member point load direction specified.')
            case_set.setdefault(case, None)
            ml.append((mi, case, PT_KINDS[direction], (P, x, 0.0, 0.0)))
        for load in m.DistLoads:
            direction, w1, w2, x1, x2, case = load[:6]
            if direction not in DIST_KINDS:
                continue
            case_set.setdefault(case, None)
            ml.append((mi, case, DIST_KINDS[direction], (w1, w2, x1, x2)))
    for qi, q in enumerate(quads):
        for pressure, case in q.pressures:
            case_set.setdefault(case, None)
            qp.append((qi, case, pressure))
    for pi, p in enumerate(plates):
        for pressure, case in p.pressures:
            case_set.setdefault(case, None)
            pp.append((pi, case, pressure))

    A.cases = list(case_set.keys())
    cidx = {c: i for i, c in enumerate(A.cases)}
    A.nl_dof = _i32([t[0] for t in nl], (len(nl),))
    A.nl_case = _i32([cidx[t[1]] for t in nl], (len(nl),))
    A.nl_val = _f64([t[2] for t in nl], (len(nl),))
    A.ml_member = _i32([t[0] for t in ml], (len(ml),))
    A.ml_case = _i32([cidx[t[1]] for t in ml], (len(ml),))
    A.ml_kind = _i32([t[2] for t in ml], (len(ml),))
    A.ml_params = _f64([t[3] for t in ml], (len(ml), 4))
    A.qp_elem = _i32([t[0] for t in qp], (len(qp),))
    A.qp_case = _i32([cidx[t[1]] for t in qp], (len(qp),))
    A.qp_val = _f64([t[2] for t in qp], (len(qp),))
    A.pp_elem = _i32([t[0] for t in pp], (len(pp),))
    A.pp_case = _i32([cidx[t[1]] for t in pp], (len(pp),))
    A.pp_val = _f64([t[2] for t in pp], (len(pp),))

    A.combo_names = [c.name for c in combo_list]
    fac = np.zeros((len(combo_list), len(A.cases)))
    for ci, combo in enumerate(combo_list):
        for case, factor in combo.factors.items():
            fac[ci, cidx[case]] = factor
    A.factors = np.ascontiguousarray(fac)
    return A


def partition_dofs(A):
    """Free / known DOF index lists and known values (`Analysis._partition_D`, `Analysis.py:1412-1505`).

    A DOF is unknown (D1) when it is neither supported nor has an enforced displacement; otherwise
    it is known (D2) with value = the enforced displacement, or 0.0 for a plain support.
    """
    bits = (1 << np.arange(6, dtype=np.uint8))
    sup = (A.support[:, None] & bits) != 0
    enf = (A.enf_mask[:, None] & bits) != 0
    known = (sup | enf).reshape(-1)
    d1 = np.nonzero(~known)[0].astype(np.int32)
    d2 = np.nonzero(known)[0].astype(np.int32)
    vals = np.where(enf, A.enf_val, 0.0).reshape(-1)[d2]
    return d1, d2, np.ascontiguousarray(vals)
