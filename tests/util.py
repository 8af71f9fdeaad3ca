"""Helpers shared by the CPU and GPU tests: fixture loading and comparison metrics."""
import os
from types import SimpleNamespace

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

FIXTURES = ['space_frame', 'portal_frame', 'continuous_beam', 'quad_mesh_xy', 'quad_mesh_yz_ortho', 'plate_mesh_xz',
            'skew_shell', 'moment_frame', 'mat_foundation', 'mixed_building', 'pdelta_column', 'pdelta_frame',
            'braced_tc', 'pdelta_near_buckling', 'tc_load_steps', 'mat_tc', 'rect_mesh_slab', 'mat_mesh_tc', 'cylinder_tank']


# fixtures whose stored displacements come from the tension/compression-only iteration (input arrays carry the
# reference's final element activity; a plain linear solve of them is not what the reference stored)
TC_FIXTURES = ('braced_tc', 'tc_load_steps', 'mat_tc', 'mat_mesh_tc')
PDELTA_FIXTURES = ('pdelta_column', 'pdelta_frame', 'pdelta_near_buckling')


def load_fixture(name):
    """Returns (A, ref, mode): flattened input arrays (attribute access like `flatten.ModelArrays`),
    the dict of reference outputs, and the analysis mode the reference ran."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    A = SimpleNamespace()
    ref = {}
    for k in z.files:
        if k.startswith('in_'):
            setattr(A, k[3:], z[k])
        elif k.startswith('ref_'):
            ref[k[4:]] = z[k]
    A.n_nodes = int(A.n_nodes)
    A.n_dof = 6*A.n_nodes
    A.cases = [str(c) for c in A.cases]
    A.combo_names = [str(c) for c in A.combo_names]
    return A, ref, str(z['mode'])


def rel_err(a, b):
    """max |a - b| / max |b| (vector-relative), 0 when both are empty or identical."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    if a.size == 0 and b.size == 0:
        return 0.0
    scale = np.abs(b).max()
    d = np.abs(a - b).max()
    return float(d if scale == 0 else d/scale)


# Fixtures whose elements are all axis-aligned: the reference's exact-zero filter (FEModel3D.py:130) is then
# structural and the COO / CSR pattern must match bit for bit.  The others contain inclined or rolled
# members / skewed shells, where `inv(T) @ k @ T` through LAPACK leaves rounding noise (|v| ~ 1e-17 |K|) in
# entries that are mathematically zero, so the reference's own pattern is rounding-defined.
GENERAL_ORIENTATION = ('portal_frame', 'skew_shell', 'braced_tc', 'cylinder_tank')


def compare_general_pattern(K, Kr, tol_values=1e-12, tol_noise=1e-14):
    """For general-orientation models: values agree on the union pattern, and every entry stored by only one
    side is rounding noise.  Returns the number of such entries (reported by the caller)."""
    import scipy.sparse as sps
    K, Kr = sps.csr_matrix(K), sps.csr_matrix(Kr)
    scale = abs(Kr).max()
    assert abs(K - Kr).max() < tol_values*scale
    Pa = K.copy(); Pa.data[:] = 1.0
    Pb = Kr.copy(); Pb.data[:] = 1.0
    only = (Pa - Pb).tocoo()
    n_only = int(np.count_nonzero(only.data))
    for r, c in zip(only.row[only.data != 0], only.col[only.data != 0]):
        assert abs(K[r, c]) <= tol_noise*scale and abs(Kr[r, c]) <= tol_noise*scale
    return n_only


def mesh_record(model):
    """Names, coordinates and connectivity of a model built with `add_rectangle_mesh`, and of each of its meshes, in
    dictionary order (works on the reference's objects and on ours)."""
    rec = {}
    nodes = list(model.nodes.values())
    rec['node_names'] = np.array([n.name for n in nodes])
    rec['node_keys'] = np.array(list(model.nodes.keys()))
    rec['node_xyz'] = np.array([[n.X, n.Y, n.Z] for n in nodes], dtype=float).reshape(-1, 3)
    for kind, table in (('quad', model.quads), ('plate', model.plates)):
        elems = list(table.values())
        rec[kind + '_names'] = np.array([e.name for e in elems], dtype=str)
        rec[kind + '_keys'] = np.array(list(table.keys()), dtype=str)
        rec[kind + '_conn'] = np.array([[e.i_node.name, e.j_node.name, e.m_node.name, e.n_node.name] for e in elems],
                                       dtype=str).reshape(-1, 4)
        rec[kind + '_props'] = np.array([[e.t, e.kx_mod, e.ky_mod, e.E, e.nu] for e in elems], dtype=float).reshape(-1, 5)
    for kind, table in (('member', model.members), ('spring', model.springs)):
        rec[kind + '_keys'] = np.array(list(table.keys()), dtype=str)
        rec[kind + '_names'] = np.array([e.name for e in table.values()], dtype=str)
        rec[kind + '_ij'] = np.array([[e.i_node.name, e.j_node.name] for e in table.values()], dtype=str).reshape(-1, 2)
    rec['combos'] = np.array([f'{c.name}|{c.combo_tags}|{sorted(c.factors.items())}' for c in model.load_combos.values()], dtype=str)
    rec['section_reprs'] = np.array([repr(sec) for sec in model.sections.values()], dtype=str)
    springs, loads = [], []
    for n in nodes:
        for dof in ('DX', 'DY', 'DZ', 'RX', 'RY', 'RZ'):
            k, direction, active = getattr(n, 'spring_' + dof)
            if k is not None:
                springs.append((n.name, dof, repr(float(k)), str(direction), str(active)))
        for load in n.NodeLoads:
            loads.append((n.name, str(load[0]), repr(float(load[1])), str(load[2])))
    rec['removed'] = np.array(list(getattr(model, '_test_removed', [])), dtype=str)
    rec['node_supports'] = np.array([[n.support_DX, n.support_DY, n.support_DZ, n.support_RX, n.support_RY, n.support_RZ]
                                     for n in nodes], dtype=bool).reshape(-1, 6)
    rec['node_springs'] = np.array(springs, dtype=str).reshape(-1, 5)
    rec['node_loads'] = np.array(loads, dtype=str).reshape(-1, 4)
    for mname, mesh in list(model.meshes.items()) + list(getattr(model, 'mats', {}).items()):
        rec['mesh_' + mname + '_nodes'] = np.array(list(mesh.nodes.keys()), dtype=str)
        rec['mesh_' + mname + '_elements'] = np.array(list(mesh.elements.keys()), dtype=str)
        rec['mesh_' + mname + '_last'] = np.array([getattr(mesh.last_node, 'name', 'None'), getattr(mesh.last_element, 'name', 'None')])
    for wname, wall in getattr(model, 'shear_walls', {}).items():
        for label, table, ext in (('piers', wall.piers, 'width'), ('beams', wall.coupling_beams, 'length')):
            regions = list(table.values())
            rec[f'wall_{wname}_{label}_names'] = np.array([r.name for r in regions] + list(table.keys()), dtype=str)
            rec[f'wall_{wname}_{label}_box'] = np.array([[r.x, r.y, getattr(r, ext), r.height] for r in regions],
                                                        dtype=float).reshape(-1, 4)
            rec[f'wall_{wname}_{label}_quads'] = np.array([','.join(q.name for q in r.plates) for r in regions], dtype=str)
    return rec


# ---------------------------------------------------------------------------------------------
# member result diagrams (tests/golden/member_diagrams.npz, tests/test_diagrams_cpu.py)
# ---------------------------------------------------------------------------------------------
DIAGRAM_FIXTURES = ['portal_frame', 'continuous_beam', 'moment_frame', 'pdelta_frame', 'pdelta_column',
                    'pdelta_near_buckling', 'braced_tc', 'tc_load_steps']
DIAGRAM_POINTS = 9
DIAGRAM_ARRAYS = [('shear_array', 'Fy'), ('shear_array', 'Fz'), ('moment_array', 'My'), ('moment_array', 'Mz'),
                  ('axial_array', None), ('torque_array', None), ('deflection_array', 'dx'), ('deflection_array', 'dy'),
                  ('deflection_array', 'dz')]
DIAGRAM_POINT_QUERIES = [('shear', 'Fy'), ('shear', 'Fz'), ('moment', 'My'), ('moment', 'Mz'), ('axial', None),
                         ('torque', None), ('deflection', 'dx'), ('deflection', 'dy'), ('deflection', 'dz'),
                         ('rel_deflection', 'dy'), ('rel_deflection', 'dz')]
DIAGRAM_EXTREMES = [('shear', 'Fy'), ('shear', 'Fz'), ('moment', 'My'), ('moment', 'Mz'), ('axial', None), ('torque', None),
                    ('deflection', 'dx'), ('deflection', 'dy'), ('deflection', 'dz')]
DIAGRAM_STATIONS = (0.0, 0.37, 0.5, 0.81, 1.0)             # fractions of the physical member's length


def _call(obj, method, direction, *rest):
    fn = getattr(obj, method)
    return fn(direction, *rest) if direction is not None else fn(*rest)


def diagram_record(model):
    """Every physical member's diagrams for every combination (works on the reference's objects and on ours):
    arrays[combo, member, quantity, 2, DIAGRAM_POINTS], points[combo, member, query, station], extremes[combo, member,
    quantity, (max, min)], and the members' activity."""
    combos = list(model.load_combos.keys())
    members = list(model.members.values())
    arrays = np.zeros((len(combos), len(members), len(DIAGRAM_ARRAYS), 2, DIAGRAM_POINTS))
    points = np.zeros((len(combos), len(members), len(DIAGRAM_POINT_QUERIES), len(DIAGRAM_STATIONS)))
    extremes = np.zeros((len(combos), len(members), len(DIAGRAM_EXTREMES), 2))
    active = np.zeros((len(combos), len(members)), dtype=bool)
    for c, combo in enumerate(combos):
        for k, member in enumerate(members):
            active[c, k] = bool(member.active[combo])
            L = member.L()
            for q, (method, direction) in enumerate(DIAGRAM_ARRAYS):
                arrays[c, k, q] = _call(member, method, direction, DIAGRAM_POINTS, combo)
            for q, (method, direction) in enumerate(DIAGRAM_POINT_QUERIES):
                for s, frac in enumerate(DIAGRAM_STATIONS):
                    points[c, k, q, s] = _call(member, method, direction, frac*L, combo)
            for q, (what, direction) in enumerate(DIAGRAM_EXTREMES):
                extremes[c, k, q, 0] = _call(member, 'max_' + what, direction, combo)
                extremes[c, k, q, 1] = _call(member, 'min_' + what, direction, combo)
    return {'arrays': arrays, 'points': points, 'extremes': extremes, 'active': active,
            'member_names': np.array([m.name for m in members], dtype=str)}


# ---------------------------------------------------------------------------------------------
# shear wall results (tests/golden/shear_walls.npz)
# ---------------------------------------------------------------------------------------------
SHEAR_WALL_COMBOS = ['1.0E', 'Stiffness: Roof']


def shear_wall_record(model):
    """`sum_forces` of every pier and coupling beam of every wall for SHEAR_WALL_COMBOS, and the walls' roof stiffness."""
    rec = {}
    for wname, wall in model.shear_walls.items():
        for label, table in (('piers', wall.piers), ('beams', wall.coupling_beams)):
            rec[f'{wname}_{label}'] = np.array([[r.sum_forces(c) for r in table.values()] for c in SHEAR_WALL_COMBOS], dtype=float)
        rec[f'{wname}_stiffness'] = np.array(wall.stiffness('Roof'))
    return rec


# ---------------------------------------------------------------------------------------------
# per-element fixed-end reactions (tests/golden/element_fer.npz)
# ---------------------------------------------------------------------------------------------
FER_FIXTURES = ['portal_frame', 'continuous_beam', 'moment_frame', 'quad_mesh_xy', 'skew_shell', 'plate_mesh_xz', 'mixed_building']


def fer_record(model):
    """[combo][element][...] fixed-end reaction vectors of every sub-member, quad and plate (reference objects or ours)."""
    combos = list(model.load_combos.keys())
    subs = [s for pm in model.members.values() for s in pm.sub_members.values()]
    rec = {}
    for label, elems, methods in (('member', subs, ('_fer_unc', 'fer', 'FER')), ('quad', list(model.quads.values()), ('fer', 'FER')),
                                  ('plate', list(model.plates.values()), ('fer', 'FER'))):
        for method in methods:
            width = 12 if label == 'member' else 24
            rec[f'{label}_{method}'] = np.array([[np.asarray(getattr(e, method)(c), dtype=float).reshape(-1) for e in elems]
                                                 for c in combos], dtype=float).reshape(len(combos), len(elems), width)
    return rec
