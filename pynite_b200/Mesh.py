"""Rectangular shell meshes: the caller-side generator of BASELINE configs 3 and 4 (`Pynite/Mesh.py:14-170, 719-1124`).

`FEModel3D.add_rectangle_mesh` + `RectangleMesh.generate` define node and element names, their order and their
coordinates -- all API-visible -- so the generator reproduces the reference's stepping exactly: the same
control-point clean-up, the same repeated floating-point additions (`x += b`), the same numbering formula, the same
opening rules and the same renaming of names the model already uses.  What changes is the cost: the coordinate lists
are computed once per direction instead of once per row, the orphan-node search is a set lookup instead of four list
scans per node (`Mesh.py:1013-1021`, quadratic in the reference), and the result queries (`max_shear`, `min_moment`, ...)
read one batched stress table per load combination from the device (`pn_shell_stresses`) instead of five kernel-free
Python evaluations per element.

Annulus, cylinder and frustrum meshes are not rebuilt (SURVEY 2 marks the meshers out of the hot path; the rectangle is
the one the benchmark configurations use).
"""
from __future__ import annotations

from math import ceil, isclose

import numpy as np

from pynite_b200.Node3D import Node3D
from pynite_b200.Quad3D import Plate3D, Quad3D


class Mesh:
    """Common state of a generated mesh (`Mesh.py:14-51`)."""

    def __init__(self, thickness, material_name, model, kx_mod=1, ky_mod=1, start_node='N1', start_element='Q1'):
        self.thickness = thickness
        self.material_name = material_name
        self.model = model
        self.kx_mod = kx_mod
        self.ky_mod = ky_mod
        self.start_node = start_node
        self.last_node = None
        self.start_element = start_element
        self.last_element = None
        self.nodes = {}
        self.elements = {}
        self.element_type = 'Quad'
        self.is_generated = False
        self.needs_update = False

    def __repr__(self):
        return f"Mesh(material_name={self.material_name!r}, thickness={self.thickness})"

    def generate(self):
        pass

    # ---- `Mesh._remove_from_model` (:56-108): nodes shared with elements outside the mesh stay in the model ----------
    def _remove_from_model(self):
        model = self.model
        for name, element in list(self.elements.items()):
            table = model.plates if element.type == 'Rect' else model.quads
            if name in table:
                del table[name]
        shared = set()
        for element in list(model.quads.values()) + list(model.plates.values()):
            if element.name not in self.elements:
                for node in (element.i_node, element.j_node, element.m_node, element.n_node):
                    if node.name in self.nodes:
                        shared.add(node.name)
        for element in list(model.members.values()) + list(model.springs.values()):
            for node in (element.i_node, element.j_node):
                if node.name in self.nodes:
                    shared.add(node.name)
        for name in list(self.nodes.keys()):
            if name in model.nodes and name not in shared:
                del model.nodes[name]
        self.nodes.clear()
        self.elements.clear()

    # ---- `Mesh._rename_duplicates` (:110-164) ----------------------------------------------------------------------
    def _rename_duplicates(self):
        model = self.model
        revised_nodes, revised_elements = {}, {}
        for node in self.nodes.values():
            if node.name in model.nodes:
                node.name = model.unique_name(model.nodes, 'N')
            model.nodes[node.name] = node
            revised_nodes[node.name] = node
        for element in self.elements.values():
            table, prefix = (model.plates, 'R') if element.type == 'Rect' else (model.quads, 'Q')
            if element.name in table:
                element.name = model.unique_name(table, prefix)
            table[element.name] = element
            revised_elements[element.name] = element
        self.nodes = revised_nodes
        self.elements = revised_elements

    # ---- result queries (`Mesh.py:172-716`): corners + centre of every element, all tagged combinations ----------------
    def _combos(self, combo_tags):
        out = []
        for combo in self.model.load_combos.values():
            if isinstance(combo_tags, str):
                include = combo.name == combo_tags
            elif isinstance(combo_tags, list):
                include = any(tag in (combo.combo_tags or []) for tag in combo_tags)
            else:
                include = False
            if include:
                out.append(combo.name)
        return out

    def _extreme(self, column, local, combo_tags, largest):
        from pynite_b200 import analysis
        best = None
        quads = [e for e in self.elements.values() if e.type == 'Quad']
        rects = [e for e in self.elements.values() if e.type == 'Rect']
        for combo in self._combos(combo_tags):
            vals = []
            if quads:
                pts = [[-1, -1], [1, -1], [1, 1], [-1, 1], [0.0, 0.0]]
                table = analysis.shell_stress_points(self.model, 'quad', combo, pts, local)
                vals.append(table[[e.ID for e in quads]][:, :, column].reshape(-1))
            # rectangles are evaluated in their own local lengths: one table per distinct (width, height)
            groups = {}
            for e in rects:
                groups.setdefault((e.width(), e.height()), []).append(e.ID)
            for (w, h), ids in groups.items():
                pts = [[0, 0], [w, 0], [w, h], [0, h], [(0 + w + w + 0)/4, (0 + 0 + h + h)/4]]
                table = analysis.shell_stress_points(self.model, 'plate', combo, pts, True)
                vals.append(table[ids][:, :, column].reshape(-1))
            for v in vals:
                if v.size:
                    cand = float(v.max() if largest else v.min())
                    if best is None or (cand > best if largest else cand < best):
                        best = cand
        return 0.0 if best is None else best

    # component index and local / global flag, by the reference's rules (`Mesh.py:190-204, 370-387, 552-568`)
    @staticmethod
    def _shear_component(direction):
        local = direction not in ['QX', 'QY']
        if direction.upper() == 'QX':
            return 0, local
        if direction.upper() == 'QY':
            return 1, local
        raise Exception("Invalid direction specified for mesh shear results. Valid values are 'Qx', 'Qy', 'QX', or 'QY'.")

    @staticmethod
    def _moment_component(direction):
        local = direction not in ['MX', 'MY', 'MZ']
        if direction.upper() == 'MX':
            return 0, local
        if direction.upper() == 'MY':
            return 1, local
        if direction == 'Mxy' or direction.upper() == 'MZ':
            return 2, local
        raise Exception("Invalid direction specified for mesh moment results. "
                        "Valid values are 'Mx', 'My', 'Mxy', 'MX', 'MY', or 'MZ'.")

    @staticmethod
    def _membrane_component(direction):
        local = direction not in ['SX', 'SY']
        if direction.upper() == 'SX':
            return 0, local
        if direction.upper() == 'SY':
            return 1, local
        if direction == 'Sxy':
            return 2, local
        raise Exception("Invalid direction specified for mesh membrane stress results. Valid values are 'Sx', 'Sy', or 'Sxy'.")

    def max_shear(self, direction='Qx', combo_tags='Combo 1'):
        i, local = self._shear_component(direction)
        return self._extreme(0 + i, local, combo_tags, True)

    def min_shear(self, direction='Qx', combo_tags='Combo 1'):
        i, local = self._shear_component(direction)
        return self._extreme(0 + i, local, combo_tags, False)

    def max_moment(self, direction='Mx', combo_tags='Combo 1'):
        i, local = self._moment_component(direction)
        return self._extreme(3 + i, local, combo_tags, True)

    def min_moment(self, direction='Mx', combo_tags='Combo 1'):
        i, local = self._moment_component(direction)
        return self._extreme(3 + i, local, combo_tags, False)

    def max_membrane(self, direction='Sx', combo_tags='Combo 1'):
        i, local = self._membrane_component(direction)
        return self._extreme(6 + i, local, combo_tags, True)

    def min_membrane(self, direction='Sx', combo_tags='Combo 1'):
        i, local = self._membrane_component(direction)
        return self._extreme(6 + i, local, combo_tags, False)


class RectOpening:
    """`Mesh.py:1099-1123`."""

    def __init__(self, x_left, y_bott, width, height):
        self.x_left = x_left
        self.y_bott = y_bott
        self.width = width
        self.height = height

    def __repr__(self):
        return f"RectOpening(x_left={self.x_left}, y_bott={self.y_bott}, width={self.width}, height={self.height})"


class RectangleMesh(Mesh):
    """`Mesh.py:719-1096`."""

    def __init__(self, mesh_size, width, height, thickness, material_name, model, kx_mod=1.0, ky_mod=1.0, origin=[0, 0, 0],
                 plane='XY', x_control=None, y_control=None, start_node='N1', start_element='Q1', element_type='Quad'):
        super().__init__(thickness, material_name, model, kx_mod, ky_mod, start_node, start_element)
        self.mesh_size = mesh_size
        self.width = width
        self.height = height
        self.origin = origin
        self.plane = plane
        self.x_control = [] if x_control is None else x_control
        self.y_control = [] if y_control is None else y_control
        self.element_type = element_type
        self.openings = {}

    def __repr__(self):
        return (f"RectangleMesh(material_name={self.material_name!r}, width={self.width}, height={self.height}, "
                f"thickness={self.thickness})")

    @staticmethod
    def _clean(control, limit):
        """Sorted control points without (near) duplicates and without points outside [0, limit] (`:802-829`)."""
        control = sorted(control)
        unique = [control[i] for i in range(len(control) - 1) if not isclose(control[i], control[i + 1])]
        unique.append(control[-1])
        return [v for v in unique if v >= -1e-10 and v <= limit + 1e-10]

    @staticmethod
    def _stations(control, mesh_size):
        """Coordinates of the node lines along one direction: the reference's stepping loop (`:852-927`), verbatim in its
        floating-point operations (the value is carried from one control interval to the next by `-= old step`,
        `+= new step`, and advanced by repeated addition)."""
        out = []
        v, step = 0, None
        for j in range(1, len(control)):
            if j != 1:
                v -= step
            n = max(1, (control[j] - control[j - 1])/mesh_size)
            step = (control[j] - control[j - 1])/ceil(n)
            if j != 1:
                v += step
            while round(v, 10) <= round(control[j], 10):
                out.append(v)
                v += step
        return out

    def generate(self):
        if self.is_generated:
            self._remove_from_model()
        plane = self.plane
        if plane not in ('XY', 'YZ', 'XZ'):
            raise Exception('Invalid plane selected for RectangleMesh.')
        if self.element_type == 'Quad':
            prefix, cls = 'Q', Quad3D
        elif self.element_type == 'Rect':
            prefix, cls = 'R', Plate3D
        else:
            raise Exception('Invalid element type specified for RectangleMesh. Select \'Quad\' or \'Rect\'.')
        Xo, Yo, Zo = self.origin[0], self.origin[1], self.origin[2]
        # the mesh's boundaries join the control points (appended to the mesh's own lists, as the reference does)
        self.x_control.append(0)
        self.x_control.append(self.width)
        self.y_control.append(0)
        self.y_control.append(self.height)
        xs = self._stations(self._clean(self.x_control, self.width), self.mesh_size)
        ys = self._stations(self._clean(self.y_control, self.height), self.mesh_size)
        node_offset = int(self.start_node[1:]) - 1
        element_offset = int(self.start_element[1:]) - 1
        ncol_nodes, nrow_nodes = len(xs), len(ys)
        num_cols, num_rows = ncol_nodes - 1, nrow_nodes - 1

        model = self.model
        node_list = []
        num = 1
        for y in ys:
            for x in xs:
                if plane == 'XY':
                    X, Y, Z = Xo + x, Yo + y, Zo + 0
                elif plane == 'YZ':
                    X, Y, Z = Xo + 0, Yo + y, Zo + x
                else:
                    X, Y, Z = Xo + x, Yo + 0, Zo + y
                node_list.append(Node3D(model, 'N' + str(num + node_offset), X, Y, Z))
                num += 1
        elem_list = []
        for i in range(1, num_cols*num_rows + 1):
            i_node = i + (i - 1)//num_cols            # (1-based index into node_list)
            j_node = i_node + 1
            m_node = j_node + (num_cols + 1)
            n_node = m_node - 1
            elem_list.append(cls(prefix + str(i + element_offset), node_list[i_node - 1], node_list[j_node - 1],
                                 node_list[m_node - 1], node_list[n_node - 1], self.thickness, self.material_name, model,
                                 self.kx_mod, self.ky_mod))

        # openings (`:964-1011`): nodes strictly inside one, elements that lie within one
        if self.openings:
            def inside(node):
                x, y = self.node_local_coords(node)
                return any(round(x, 10) > round(o.x_left, 10) and round(x, 10) < round(o.x_left + o.width, 10)
                           and round(y, 10) > round(o.y_bott, 10) and round(y, 10) < round(o.y_bott + o.height, 10)
                           for o in self.openings.values())

            def within(element):
                left, top = self.node_local_coords(element.n_node)
                right, bott = self.node_local_coords(element.j_node)
                return any(round(o.y_bott + o.height, 10) >= round(top, 10) and round(o.y_bott, 10) <= round(bott, 10)
                           and round(o.x_left, 10) <= round(left, 10) and round(o.x_left + o.width, 10) >= round(right, 10)
                           for o in self.openings.values())
            gone = {id(n) for n in node_list if inside(n)}
            elem_list = [e for e in elem_list if not within(e)]
            node_list = [n for n in node_list if id(n) not in gone]
        # orphaned nodes: not referenced by any remaining element (`:1013-1025`)
        used = set()
        for e in elem_list:
            used.update((id(e.i_node), id(e.j_node), id(e.m_node), id(e.n_node)))
        node_list = [n for n in node_list if id(n) in used]

        self.nodes = {n.name: n for n in node_list}
        self.elements = {e.name: e for e in elem_list}
        self.last_node = node_list[-1]
        self.last_element = elem_list[-1]
        self._rename_duplicates()          # also files the nodes and elements in the model (`:1031-1044`)
        self.is_generated = True
        self.needs_update = False

    def node_local_coords(self, node):
        if self.plane == 'XY':
            return node.X - self.origin[0], node.Y - self.origin[1]
        if self.plane == 'YZ':
            return node.Z - self.origin[2], node.Y - self.origin[1]
        return node.X - self.origin[0], node.Z - self.origin[2]

    def add_rect_opening(self, name, x_left, y_bott, width, height):
        self.openings[name] = RectOpening(x_left, y_bott, width, height)
        self.x_control.append(x_left)
        self.y_control.append(y_bott)
        self.x_control.append(x_left + width)
        self.y_control.append(y_bott + height)
        if self.is_generated:
            self.needs_update = True


def mesh_arrays(mesh):
    """(names, xyz) of the mesh's nodes and (names, i-j-m-n node names) of its elements, for tests and tools."""
    nodes = list(mesh.nodes.values())
    elems = list(mesh.elements.values())
    return ([n.name for n in nodes], np.array([[n.X, n.Y, n.Z] for n in nodes], dtype=float).reshape(-1, 3),
            [e.name for e in elems], [[e.i_node.name, e.j_node.name, e.m_node.name, e.n_node.name] for e in elems])
