"""`FEModel3D`: the reference's model facade, with the assembly + solve path running on a B200.

Builders (`add_node` ... `add_quad_surface_pressure`) keep the reference signatures, naming rules
and error behaviour (`Pynite/FEModel3D.py:224-1549`).  The three analysis drivers
(`analyze_linear` :2250, `analyze` :2334, `analyze_PDelta` :2413) keep their keyword arguments, but
everything between "objects are numbered" and "results are stored" -- element stiffness, COO->CSR
assembly, partition, right-hand sides, the solve for every load combination, reactions -- is done by
the sm_100a library behind `pynite_b200.engine` (C ABI in `include/pynite_b200.h`).

There is no CPU solver in this package: a missing CUDA library raises at the first analysis call, and
`sparse=False` (the reference's dense NumPy path) runs the same device path -- only `Ke / Kg / M` then return
dense arrays, as in the reference.
"""
from __future__ import annotations

import numpy as np

from pynite_b200 import analysis as _analysis
from pynite_b200.LoadCombo import LoadCombo
from pynite_b200.Material import Material, Section
from pynite_b200.Member3D import PhysMember
from pynite_b200.Node3D import Node3D
from pynite_b200.Quad3D import Plate3D, Quad3D
from pynite_b200.Spring3D import Spring3D

_NODE_DOFS = ('DX', 'DY', 'DZ', 'RX', 'RY', 'RZ')
_NODE_LOADS = ('FX', 'FY', 'FZ', 'MX', 'MY', 'MZ')
_MEMBER_PT = ('Fx', 'Fy', 'Fz', 'FX', 'FY', 'FZ', 'Mx', 'My', 'Mz', 'MX', 'MY', 'MZ')
_MEMBER_DIST = ('Fx', 'Fy', 'Fz', 'FX', 'FY', 'FZ')


class FEModel3D:

    def __init__(self):
        self.nodes = {}
        self.materials = {}
        self.sections = {}
        self.springs = {}
        self.members = {}
        self.quads = {}
        self.plates = {}
        self.meshes = {}
        self.mats = {}
        self.shear_walls = {}
        self.load_combos = {}
        self._D = {}        # {combo: (6N, 1) displacement vector}
        self._Rxn = {}      # {combo: (6N, 1) reaction vector}; read through node.RxnFX[...] etc.
        self.solution = None
        self._engine = None     # device context (lazy)
        self._run = None        # bookkeeping of the last GPU analysis (arrays, combo index)
        self.last_stats = None  # per-phase timings / solver statistics of the last analysis

    def __repr__(self):
        return f"FEModel3D(solution={self.solution!r})"

    # ------------------------------------------------------------------------------------------
    # naming helpers
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _auto_name(name, table, prefix, kind):
        """User name must be new; empty/None gets prefix + count, bumped until unused."""
        if name:
            if name in table:
                raise NameError(f"{kind} name '{name}' already exists")
            return name
        name = prefix + str(len(table))
        bump = 1
        while name in table:
            name = prefix + str(len(table) + bump)
            bump += 1
        return name

    def unique_name(self, dictionary, prefix):
        """Next available name for a dictionary of objects (`FEModel3D.py:2882-2901`): prefix + (count + 1), bumped
        until unused -- the names the mesh builders hand out are API-visible, so the rule is the reference's."""
        name = prefix + str(len(dictionary) + 1)
        i = 2
        while name in dictionary:
            name = prefix + str(len(dictionary) + i)
            i += 1
        return name

    def _nodes_by_name(self, *names):
        try:
            return [self.nodes[n] for n in names]
        except KeyError as e:
            raise NameError(f"Node '{e.args[0]}' does not exist in the model")

    @property
    def load_cases(self):
        """Sorted list of every load case named by a load (`FEModel3D.py:169-222`)."""
        cases = []
        for node in self.nodes.values():
            cases.extend(load[2] for load in node.NodeLoads)
        for member in self.members.values():
            cases.extend(load[3] for load in member.PtLoads)
            cases.extend(load[5] for load in member.DistLoads)
        for shell in list(self.plates.values()) + list(self.quads.values()):
            cases.extend(load[1] for load in shell.pressures)
        for wall in self.shear_walls.values():
            cases.extend(load[2] for load in wall._shears)
            cases.extend(load[2] for load in wall._axials)
        for mat in self.mats.values():
            cases.extend(load[3] for load in mat.pt_loads)
        return sorted(dict.fromkeys(cases))

    # ------------------------------------------------------------------------------------------
    # builders
    # ------------------------------------------------------------------------------------------
    def add_node(self, name, X, Y, Z):
        name = self._auto_name(name, self.nodes, 'N', 'Node')
        self.nodes[name] = Node3D(self, name, X, Y, Z)
        self.solution = None
        return name

    def add_material(self, name, E, G, nu, rho, fy=None):
        name = self._auto_name(name, self.materials, 'M', 'Material')
        self.materials[name] = Material(self, name, E, G, nu, rho, fy)
        self.solution = None
        return name

    def add_section(self, name, A, Iy, Iz, J):
        name = self._auto_name(name, self.sections, 'SC', 'Section')
        self.sections[name] = Section(self, name, A, Iy, Iz, J)
        return name

    def add_steel_section(self, name, A, Iy, Iz, J, Zy, Zz, material_name):
        """`FEModel3D.add_steel_section` (:340-377)."""
        from pynite_b200.Material import SteelSection
        name = self._auto_name(name, self.sections, 'SC', 'Section')
        self.sections[name] = SteelSection(self, name, A, Iy, Iz, J, Zy, Zz, material_name)
        return name

    def add_spring(self, name, i_node, j_node, ks, tension_only=False, comp_only=False):
        name = self._auto_name(name, self.springs, 'S', 'Spring')
        ni, nj = self._nodes_by_name(i_node, j_node)
        self.springs[name] = Spring3D(name, ni, nj, ks, self.load_combos,
                                      tension_only=tension_only, comp_only=comp_only)
        self.solution = None
        return name

    def add_member(self, name, i_node, j_node, material_name, section_name, rotation=0.0,
                   tension_only=False, comp_only=False):
        name = self._auto_name(name, self.members, 'M', 'Member')
        ni, nj = self._nodes_by_name(i_node, j_node)
        self.members[name] = PhysMember(self, name, ni, nj, material_name, section_name,
                                        rotation=rotation, tension_only=tension_only, comp_only=comp_only)
        self.solution = None
        return name

    def add_plate(self, name, i_node, j_node, m_node, n_node, t, material_name, kx_mod=1.0, ky_mod=1.0):
        name = self._auto_name(name, self.plates, 'P', 'Plate')
        n = self._nodes_by_name(i_node, j_node, m_node, n_node)
        self.plates[name] = Plate3D(name, n[0], n[1], n[2], n[3], t, material_name, self, kx_mod, ky_mod)
        self.solution = None
        return name

    def add_quad(self, name, i_node, j_node, m_node, n_node, t, material_name, kx_mod=1.0, ky_mod=1.0):
        name = self._auto_name(name, self.quads, 'Q', 'Quad')
        n = self._nodes_by_name(i_node, j_node, m_node, n_node)
        self.quads[name] = Quad3D(name, n[0], n[1], n[2], n[3], t, material_name, self, kx_mod, ky_mod)
        self.solution = None
        return name

    def add_rectangle_mesh(self, name, mesh_size, width, height, thickness, material_name, kx_mod=1.0, ky_mod=1.0,
                           origin=(0, 0, 0), plane='XY', x_control=None, y_control=None, start_node=None,
                           start_element=None, element_type='Quad'):
        """`FEModel3D.add_rectangle_mesh` (:617-683).  The mesh is generated by `model.meshes[name].generate()` or by the
        next analysis (`Analysis._prepare_model` :53-56), as in the reference."""
        from pynite_b200.Mesh import RectangleMesh
        if name:
            if name in self.meshes:
                raise NameError(f"Mesh name '{name}' already exists")
        else:
            name = self.unique_name(self.meshes, 'MSH')
        if start_node is None:
            start_node = self.unique_name(self.nodes, 'N')
        if element_type == 'Rect' and start_element is None:
            start_element = self.unique_name(self.plates, 'R')
        elif element_type == 'Quad' and start_element is None:
            start_element = self.unique_name(self.quads, 'Q')
        self.meshes[name] = RectangleMesh(mesh_size, width, height, thickness, material_name, self, kx_mod, ky_mod, origin,
                                          plane, x_control, y_control, start_node, start_element, element_type=element_type)
        self.solution = None
        return name

    def _new_mesh_name(self, name):
        if name:
            if name in self.meshes:
                raise NameError(f"Mesh name '{name}' already exists")
            return name
        return self.unique_name(self.meshes, 'MSH')

    def add_annulus_mesh(self, name, mesh_size, outer_radius, inner_radius, thickness, material_name, kx_mod=1.0, ky_mod=1.0,
                         origin=(0, 0, 0), axis='Y', start_node=None, start_element=None):
        """`FEModel3D.add_annulus_mesh` (:685-746)."""
        from pynite_b200.Mesh import AnnulusMesh
        name = self._new_mesh_name(name)
        start_node = self.unique_name(self.nodes, 'N') if start_node is None else start_node
        start_element = self.unique_name(self.quads, 'Q') if start_element is None else start_element
        self.meshes[name] = AnnulusMesh(mesh_size, outer_radius, inner_radius, thickness, material_name, self, kx_mod, ky_mod,
                                        origin, axis, start_node, start_element)
        self.solution = None
        return name

    def add_frustrum_mesh(self, name, mesh_size, large_radius, small_radius, height, thickness, material_name, kx_mod=1.0,
                          ky_mod=1.0, origin=(0, 0, 0), axis='Y', start_node=None, start_element=None):
        """`FEModel3D.add_frustrum_mesh` (:748-808)."""
        from pynite_b200.Mesh import FrustrumMesh
        name = self._new_mesh_name(name)
        start_node = self.unique_name(self.nodes, 'N') if start_node is None else start_node
        start_element = self.unique_name(self.quads, 'Q') if start_element is None else start_element
        self.meshes[name] = FrustrumMesh(mesh_size, large_radius, small_radius, height, thickness, material_name, self, kx_mod,
                                         ky_mod, origin, axis, start_node, start_element)
        self.solution = None
        return name

    def add_cylinder_mesh(self, name, mesh_size, radius, height, thickness, material_name, kx_mod=1, ky_mod=1, origin=(0, 0, 0),
                          axis='Y', num_elements=None, start_node=None, start_element=None, element_type='Quad'):
        """`FEModel3D.add_cylinder_mesh` (:810-890).  The cylinder is generated on construction, as in the reference."""
        from pynite_b200.Mesh import CylinderMesh
        name = self._new_mesh_name(name)
        if start_node is None:
            start_node = self.unique_name(self.nodes, 'N')
        if element_type == 'Rect' and start_element is None:
            start_element = self.unique_name(self.plates, 'R')
        elif element_type == 'Quad' and start_element is None:
            start_element = self.unique_name(self.quads, 'Q')
        self.meshes[name] = CylinderMesh(mesh_size, radius, height, thickness, material_name, self, kx_mod, ky_mod, origin, axis,
                                         start_node, start_element, num_elements, element_type)
        self.solution = None
        return name

    def add_shear_wall(self, name, mesh_size, length, height, thickness, material_name, ky_mod=0.35, plane='XY',
                       origin=[0, 0, 0]):
        """`FEModel3D.add_shear_wall` (:890-933).  Returns nothing, like the reference's."""
        from pynite_b200.ShearWall import ShearWall
        if name:
            if name in self.shear_walls:
                raise NameError(f"Shear wall name '{name}' already exists")
        else:
            name = self.unique_name(self.shear_walls, 'SW')
        self.shear_walls[name] = ShearWall(self, name, mesh_size, length, height, thickness, material_name, ky_mod, origin,
                                           plane)
        self.solution = None

    def add_mat_foundation(self, name, mesh_size, length_X, length_Z, thickness, material_name, ks, origin=[0, 0, 0],
                           x_control=[], y_control=[]):
        """`FEModel3D.add_mat_foundation` (:935-951).  (The default control lists are shared between calls, as the
        reference's mutable defaults are.)"""
        from pynite_b200.MatFoundation import MatFoundation
        if name:
            if name in self.mats:
                raise NameError(f"Mat foundation name '{name}' already exists")
        else:
            name = self.unique_name(self.mats, 'MAT')
        self.mats[name] = MatFoundation(name, mesh_size, length_X, length_Z, thickness, material_name, self, ks, origin,
                                        x_control, y_control)
        self.solution = None

    def merge_duplicate_nodes(self, tolerance=0.001):
        """`FEModel3D.merge_duplicate_nodes` (:953-1058): a node within `tolerance` of an EARLIER surviving node is merged
        into the first such node (references in elements and meshes, supports, support springs) and removed; returns the
        removed names in the reference's order.  The reference compares all pairs; here the survivors are kept in a
        spatial hash with cells of one tolerance, so two adjacent 500 x 500 meshes merge in seconds."""
        node_lookup = {name: [] for name in self.nodes}
        for table in (self.springs, self.members, self.plates, self.quads):
            for element in table.values():
                for node_type in ('i_node', 'j_node', 'm_node', 'n_node'):
                    node = getattr(element, node_type, None)
                    if node is not None:
                        node_lookup[node.name].append((element, node_type))
        names = list(self.nodes.keys())
        cell = tolerance if tolerance > 0 else 1.0
        grid = {}            # cell -> [(order, name)] of surviving nodes
        merged_into = {}     # removed name -> surviving name
        absorbed = {}        # surviving name -> [removed names] in dictionary order
        for order, name in enumerate(names):
            node = self.nodes[name]
            cx, cy, cz = int(node.X//cell), int(node.Y//cell), int(node.Z//cell)
            best = None
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        for o, other in grid.get((cx + dx, cy + dy, cz + dz), ()):
                            if (best is None or o < best[0]) and not self.nodes[other].distance(node) > tolerance:
                                best = (o, other)
            if best is None:
                grid.setdefault((cx, cy, cz), []).append((order, name))
            else:
                merged_into[name] = best[1]
                absorbed.setdefault(best[1], []).append(name)
        # the reference removes in the order (survivor by dictionary position, then its duplicates by position)
        remove_list = []
        for name_1 in names:
            if name_1 in merged_into:
                continue
            node_1 = self.nodes[name_1]
            for name_2 in absorbed.get(name_1, ()):
                node_2 = self.nodes[name_2]
                for element, node_type in node_lookup[name_2]:
                    setattr(element, node_type, node_1)
                for dof in ('support_DX', 'support_DY', 'support_DZ', 'support_RX', 'support_RY', 'support_RZ'):
                    if getattr(node_2, dof) == True:                 # noqa: E712 -- the reference's comparison
                        setattr(node_1, dof, True)
                for dof in ('spring_DX', 'spring_DY', 'spring_DZ', 'spring_RX', 'spring_RY', 'spring_RZ'):
                    value = getattr(node_2, dof)
                    if value != [None, None, None]:
                        setattr(node_1, dof, value)
                for mesh in self.meshes.values():
                    if name_2 in mesh.nodes:
                        mesh.nodes[name_2] = node_1
                        mesh.nodes[name_1] = mesh.nodes.pop(name_2)
                        # (the elements of the mesh are the model's element objects: their references were replaced above)
                remove_list.append(name_2)
        for name in remove_list:
            self.nodes.pop(name)
        self.solution = None
        return remove_list

    def delete_node(self, node_name):
        """Removes a node and every member / plate / quad attached to it (`FEModel3D.py:1060-1077`)."""
        self.nodes.pop(node_name)

        def touches(elem, attrs):
            return any(getattr(elem, a).name == node_name for a in attrs)
        self.members = {k: v for k, v in self.members.items() if not touches(v, ('i_node', 'j_node'))}
        four = ('i_node', 'j_node', 'm_node', 'n_node')
        self.plates = {k: v for k, v in self.plates.items() if not touches(v, four)}
        self.quads = {k: v for k, v in self.quads.items() if not touches(v, four)}
        self.solution = None

    def delete_spring(self, spring_name):
        self.springs.pop(spring_name)
        self.solution = None

    def delete_member(self, member_name):
        self.members.pop(member_name)
        self.solution = None

    def delete_mesh(self, mesh_name):
        """Removes a mesh and its elements; nodes shared with elements outside the mesh stay (`FEModel3D.py:1107-1130`)."""
        if mesh_name not in self.meshes:
            raise KeyError(f"Mesh '{mesh_name}' does not exist in the model.")
        self.meshes[mesh_name]._remove_from_model()
        self.meshes.pop(mesh_name)
        self.solution = None

    def rename(self):
        """Renames nodes N1.., springs S1.., members M1.., plates P1.., quads Q1.. in dictionary order
        (`FEModel3D.py:2904-2952`).  The entries are re-keyed one at a time over a snapshot of the old keys, as in the
        reference: where an old name equals a new name handed out earlier the reference's outcome (an entry replaced)
        is reproduced rather than repaired."""
        for table, prefix in ((self.nodes, 'N'), (self.springs, 'S'), (self.members, 'M'), (self.plates, 'P'),
                              (self.quads, 'Q')):
            for k, old_key in enumerate(list(table.keys()), start=1):
                new_key = prefix + str(k)
                table[new_key] = table.pop(old_key)
                table[new_key].name = new_key

    def orphaned_nodes(self):
        """Names of the nodes no spring, member, plate or quad is attached to (`FEModel3D.py:2954-2981`), in node order;
        one pass over the elements instead of one pass per node."""
        attached = set()
        for elem in list(self.quads.values()) + list(self.plates.values()):
            attached.update((id(elem.i_node), id(elem.j_node), id(elem.m_node), id(elem.n_node)))
        for elem in list(self.members.values()) + list(self.springs.values()):
            attached.update((id(elem.i_node), id(elem.j_node)))
        return [node.name for node in self.nodes.values() if id(node) not in attached]

    def def_support(self, node_name, support_DX=False, support_DY=False, support_DZ=False,
                    support_RX=False, support_RY=False, support_RZ=False):
        (node,) = self._nodes_by_name(node_name)
        node.support_DX = support_DX
        node.support_DY = support_DY
        node.support_DZ = support_DZ
        node.support_RX = support_RX
        node.support_RY = support_RY
        node.support_RZ = support_RZ
        self.solution = None

    def def_support_spring(self, node_name, dof, stiffness, direction=None):
        if dof not in _NODE_DOFS:
            raise ValueError("Invalid support spring degree of freedom. Specify 'DX', 'DY', 'DZ', 'RX', 'RY', or 'RZ'")
        if direction not in ('+', '-', None):
            raise ValueError("Invalid support spring direction. Specify '+', '-', or None.")
        (node,) = self._nodes_by_name(node_name)
        setattr(node, 'spring_' + dof, [stiffness, direction, True])
        self.solution = None

    def def_node_disp(self, node_name, direction, magnitude):
        if direction not in _NODE_DOFS:
            raise ValueError(f"direction must be 'DX', 'DY', 'DZ', 'RX', 'RY', or 'RZ'. {direction} was given.")
        (node,) = self._nodes_by_name(node_name)
        setattr(node, 'Enforced' + direction, magnitude)
        self.solution = None

    def def_releases(self, member_name, Dxi=False, Dyi=False, Dzi=False, Rxi=False, Ryi=False, Rzi=False,
                     Dxj=False, Dyj=False, Dzj=False, Rxj=False, Ryj=False, Rzj=False):
        try:
            self.members[member_name].Releases = [Dxi, Dyi, Dzi, Rxi, Ryi, Rzi, Dxj, Dyj, Dzj, Rxj, Ryj, Rzj]
        except KeyError:
            raise NameError(f"Member '{member_name}' does not exist in the model")
        self.solution = None

    def add_load_combo(self, name, factors, combo_tags=None):
        self.load_combos[name] = LoadCombo(name, combo_tags, factors)
        self.solution = None

    def add_node_load(self, node_name, direction, P, case='Case 1'):
        if direction not in _NODE_LOADS:
            raise ValueError(f"direction must be 'FX', 'FY', 'FZ', 'MX', 'MY', or 'MZ'. {direction} was given.")
        (node,) = self._nodes_by_name(node_name)
        node.NodeLoads.append((direction, P, case))
        self.solution = None

    def add_member_pt_load(self, member_name, direction, P, x, case='Case 1'):
        if direction not in _MEMBER_PT:
            raise ValueError("direction must be 'Fx', 'Fy', 'Fz', 'FX', 'FY', FZ', 'Mx', 'My', 'Mz', 'MX', 'MY', or 'MZ'. "
                             f"{direction} was given.")
        try:
            self.members[member_name].PtLoads.append((direction, P, x, case))
        except KeyError:
            raise NameError(f"Member '{member_name}' does not exist in the model")
        self.solution = None

    def add_member_dist_load(self, member_name, direction, w1, w2, x1=None, x2=None, case='Case 1',
                             self_weight=False):
        if direction not in _MEMBER_DIST:
            raise ValueError(f"direction must be 'Fx', 'Fy', 'Fz', 'FX', 'FY', or 'FZ'. {direction} was given.")
        try:
            member = self.members[member_name]
        except KeyError:
            raise NameError(f"Member '{member_name}' does not exist in the model")
        start = 0 if x1 is None else x1
        end = member.L() if x2 is None else x2
        member.DistLoads.append((direction, w1, w2, start, end, case, self_weight))
        self.solution = None

    def add_member_self_weight(self, global_direction, factor, case='Case 1'):
        if global_direction in ('Fx', 'Fy', 'Fz'):
            raise ValueError(f"Local direction '{global_direction}' is not allowed for self-weight.  "
                             "Use global directions 'FX', 'FY', or 'FZ' instead.")
        if global_direction not in ('FX', 'FY', 'FZ'):
            raise ValueError(f"Direction must be 'FX', 'FY', or 'FZ'. {global_direction} was given.")
        for member in self.members.values():
            w = factor*member.material.rho*member.section.A
            self.add_member_dist_load(member.name, global_direction, w, w, case=case, self_weight=True)

    def add_plate_surface_pressure(self, plate_name, pressure, case='Case 1'):
        try:
            self.plates[plate_name].pressures.append([pressure, case])
        except KeyError:
            raise NameError(f"Plate '{plate_name}' does not exist in the model")
        self.solution = None

    def add_quad_surface_pressure(self, quad_name, pressure, case='Case 1'):
        try:
            self.quads[quad_name].pressures.append([pressure, case])
        except KeyError:
            raise NameError(f"Quad '{quad_name}' does not exist in the model")
        self.solution = None

    def delete_loads(self):
        """Removes every load and clears stored results (`FEModel3D.py:1509-1549`)."""
        for member in self.members.values():
            member.DistLoads = []
            member.PtLoads = []
        for shell in list(self.plates.values()) + list(self.quads.values()):
            shell.pressures = []
        for node in self.nodes.values():
            node.NodeLoads = []
        self._D = {}
        self._Rxn = {}
        self.solution = None

    # ------------------------------------------------------------------------------------------
    # matrices / vectors for inspection (computed on the GPU, returned as the reference returns them)
    # ------------------------------------------------------------------------------------------
    def Ke(self, combo_name='Combo 1', log=False, check_stability=True, sparse=True):
        """Global elastic stiffness as `scipy.sparse.coo_matrix`, entries in the reference's
        emission order with exact zeros dropped (`FEModel3D.py:1551-1800`)."""
        return _analysis.global_Ke(self, combo_name, log, check_stability, sparse)

    def Kg(self, combo_name='Combo 1', log=False, sparse=True, first_step=True):
        """Global geometric stiffness as `coo_matrix` (`FEModel3D.py:1802-1895`)."""
        return _analysis.global_Kg(self, combo_name, log, sparse, first_step)

    def FER(self, combo_name='Combo 1'):
        """Global fixed-end reaction vector (6N, 1) (`FEModel3D.py:2136-2182`)."""
        return _analysis.global_load_vector(self, combo_name, 'FER')

    def P(self, combo_name='Combo 1'):
        """Global nodal force vector (6N, 1) (`FEModel3D.py:2184-2235`)."""
        return _analysis.global_load_vector(self, combo_name, 'P')

    def D(self, combo_name='Combo 1'):
        """Global displacement vector of a solved combo (`FEModel3D.py:2237-2248`)."""
        return self._D[combo_name]

    # ------------------------------------------------------------------------------------------
    # analysis drivers
    # ------------------------------------------------------------------------------------------
    def analyze_linear(self, log=False, check_stability=True, check_statics=False, sparse=True,
                       combo_tags=None):
        """First-order linear analysis: K assembled once, all combos solved as one multi-RHS job."""
        _analysis.run_linear(self, log, check_stability, check_statics, sparse, combo_tags)

    def analyze(self, log=False, check_stability=True, check_statics=False, max_iter=30, sparse=True,
                combo_tags=None, spring_tolerance=0, member_tolerance=0, num_steps=1):
        """First-order analysis with tension/compression-only iteration (`FEModel3D.py:2334-2411`)."""
        _analysis.run_first_order(self, log, check_stability, check_statics, max_iter, sparse, combo_tags,
                                  spring_tolerance, member_tolerance, num_steps)

    def analyze_PDelta(self, log=False, check_stability=True, max_iter=30, sparse=True, combo_tags=None):
        """Two-step P-Delta analysis (`FEModel3D.py:2413-2467`, `Analysis.py:329-471`)."""
        _analysis.run_pdelta(self, log, check_stability, max_iter, sparse, combo_tags)

    def _calculate_characteristic_length(self):
        """Average member length, else the bounding-box diagonal of the nodes, else 1.0 (`FEModel3D.py:1984-2002`)."""
        if self.members:
            return sum(member.L() for member in self.members.values())/len(self.members)
        if self.nodes:
            xyz = np.array([(node.X, node.Y, node.Z) for node in self.nodes.values()], dtype=float)
            return float(sum((xyz.max(axis=0) - xyz.min(axis=0))**2)**0.5)
        return 1.0

    def M(self, mass_combo_name=None, mass_direction='Y', gravity=1.0, log=False, sparse=True):
        """Global mass matrix (`FEModel3D.py:2004-2134`): consistent member mass from self-weight loads, member and
        nodal loads of the mass combination along `mass_direction` converted with m = F/g."""
        from pynite_b200 import modal as _modal
        return _modal.global_M(self, mass_combo_name, mass_direction, gravity, log, sparse)

    def analyze_modal(self, num_modes=12, mass_combo_name='Combo 1', mass_direction='Y', gravity=1.0, log=False,
                      check_stability=True):
        """Natural frequencies (`model.frequencies`, Hz) and mode shapes (`model._D['Mode i']`, `node.DX['Mode i']`)
        (`FEModel3D.py:2469-2568`); block shift-invert Lanczos whose solves run on the GPU (`modal.py`)."""
        from pynite_b200 import modal as _modal
        _modal.run_modal(self, num_modes, mass_combo_name, mass_direction, gravity, log, check_stability)

    # ------------------------------------------------------------------------------------------
    # result plumbing used by the element objects
    # ------------------------------------------------------------------------------------------
    def _element_force(self, kind, elem_id, combo_name, local):
        return _analysis.element_force(self, kind, elem_id, combo_name, local)

    def _shell_stress(self, kind, elem_id, combo_name, u, v, local):
        return _analysis.shell_stress(self, kind, elem_id, combo_name, u, v, local)

    def _member_dircos(self, sub):
        return _analysis.member_dircos_host(sub)
