"""Rectangular shell meshes: the caller-side generator of BASELINE configs 3 and 4 (`Pynite/Mesh.py:14-170, 719-1124`).

`FEModel3D.add_rectangle_mesh` + `RectangleMesh.generate` define node and element names, their order and their
coordinates -- all API-visible -- so the generator reproduces the reference's stepping exactly: the same
control-point clean-up, the same repeated floating-point additions (`x += b`), the same numbering formula, the same
opening rules and the same renaming of names the model already uses.  What changes is the cost: the coordinate lists
are computed once per direction instead of once per row, the orphan-node search is a set lookup instead of four list
scans per node (`Mesh.py:1013-1021`, quadratic in the reference), and the result queries (`max_shear`, `min_moment`, ...)
read one batched stress table per load combination from the device (`pn_shell_stresses`) instead of five kernel-free
Python evaluations per element.

The annulus, frustrum and cylinder meshes (`Mesh.py:1126-2036`) follow at the end of the file, ring by ring like the
reference's.
"""
from __future__ import annotations

from math import ceil, cos, isclose, pi, sin

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


# ---------------------------------------------------------------------------------------------------------------------
# Curved meshes (`Mesh.py:1126-2036`): annulus, frustrum, cylinder, built ring by ring like the reference's.  A ring is
# a list of node circles and a connectivity rule; the coordinates use the reference's expressions (`Xo + r*cos(angle)`
# with the angle formed exactly as there), because names, order and coordinates are API-visible.  The reference's
# quirks are kept: rings file themselves in the model when they are constructed, a cylinder uses only the axis
# component of its origin and stacks rings up to the ABSOLUTE coordinate `height`, and its rings are built with
# kx_mod = ky_mod = 1.
# ---------------------------------------------------------------------------------------------------------------------
def _on_circle(axis, Xo, Yo, Zo, r, angle, lift, who):
    """A point of a circle about a global axis; `lift` (or None) is added along the axis."""
    if axis == 'Y':
        return Xo + r*cos(angle), (Yo if lift is None else Yo + lift), Zo + r*sin(angle)
    if axis == 'X':
        return (Xo if lift is None else Xo + lift), Yo + r*sin(angle), Zo + r*cos(angle)
    if axis == 'Z':
        return Xo + r*sin(angle), Yo + r*cos(angle), (Zo if lift is None else Zo + lift)
    raise Exception('Invalid axis specified for ' + who + '.')


class _RingBase(Mesh):
    """Nodes on consecutive circles (numbered circle by circle from `start_node`), quads from 1-based node indices."""

    def _build(self, circles, conn, prefix, cls, who):
        if self.is_generated:
            self._remove_from_model()
        node_offset = self._offset(self.start_node, 'Invalid node name. Enter a letter followed by a number (e.g. \'N25\')')
        element_offset = self._offset(self.start_element, 'Invalid element ame. Enter a letter followed by a number (e.g. \'Q83\')')
        made = []
        for r, angles, lift in circles:
            for angle in angles:
                name = 'N' + str(len(made) + 1 + node_offset)
                node = Node3D(self.model, name, *_on_circle(self.axis, self.Xo, self.Yo, self.Zo, r, angle, lift, who))
                self.nodes[name] = node
                made.append(node)
        for k, (i_node, j_node, m_node, n_node) in enumerate(conn, start=1):
            name = prefix + str(k + element_offset)
            self.elements[name] = cls(name, made[i_node - 1], made[j_node - 1], made[m_node - 1], made[n_node - 1],
                                      self.thickness, self.material_name, self.model, self.kx_mod, self.ky_mod)
        for node in self.nodes.values():
            self.model.nodes[node.name] = node
        for element in self.elements.values():
            (self.model.quads if element.type == 'Quad' else self.model.plates)[element.name] = element
        self.is_generated = True
        self.needs_update = False

    @staticmethod
    def _offset(name, message):
        try:
            return int(name[1:]) - 1
        except Exception:
            raise ValueError(message)

    @staticmethod
    def _band(n):
        """Quads between a circle of n nodes (1 .. n) and the next one (n + 1 .. 2 n), closing on the first column."""
        return [(i + n, (i + 1 + n) if i != n else (1 + n), (i + 1) if i != n else 1, i) for i in range(1, n + 1)]


class AnnulusRingMesh(_RingBase):
    """`Mesh.py:1259-1418`: one ring of `num_quads` quads between two radii."""

    def __init__(self, outer_radius, inner_radius, num_quads, thickness, material_name, model, kx_mod=1, ky_mod=1,
                 origin=[0, 0, 0], axis='Y', start_node='N1', start_element='Q1'):
        super().__init__(thickness, material_name, model, kx_mod, ky_mod, start_node=start_node, start_element=start_element)
        self.inner_radius = inner_radius
        self.outer_radius = outer_radius
        self.n = num_quads
        self.Xo, self.Yo, self.Zo = origin[0], origin[1], origin[2]
        self.axis = axis
        self.generate()

    def __repr__(self):
        return (f"AnnulusRingMesh(material_name={self.material_name!r}, outer_radius={self.outer_radius}, "
                f"inner_radius={self.inner_radius}, thickness={self.thickness})")

    def generate(self):
        n = self.n
        theta = 2*pi/self.n
        angles = [theta*(i - 1) for i in range(1, n + 1)]
        self._build([(self.inner_radius, angles, None), (self.outer_radius, angles, None)], self._band(n), 'Q', Quad3D,
                    'AnnulusRingMesh')


class AnnulusTransRingMesh(_RingBase):
    """`Mesh.py:1420-1621`: a ring whose outer edge has three times the quads of its inner edge (n coarse quads, then
    3 n fine ones), through a middle circle of 2 n nodes."""

    def __init__(self, outer_radius, inner_radius, num_inner_quads, thickness, material_name, model, kx_mod=1, ky_mod=1,
                 origin=[0, 0, 0], axis='Y', start_node='N1', start_element='Q1'):
        super().__init__(thickness, material_name, model, kx_mod, ky_mod, start_node=start_node, start_element=start_element)
        self.inner_radius = inner_radius
        self.outer_radius = (inner_radius + outer_radius)/2
        self.r3 = outer_radius
        self.n = num_inner_quads
        self.Xo, self.Yo, self.Zo = origin[0], origin[1], origin[2]
        self.axis = axis
        self.generate()

    def __repr__(self):
        return (f"AnnulusTransRingMesh(material_name={self.material_name!r}, outer_radius={self.r3}, "
                f"inner_radius={self.inner_radius}, thickness={self.thickness})")

    def generate(self):
        n = self.n
        theta1 = 2*pi/self.n
        theta2 = 2*pi/(self.n*3)
        theta3 = 2*pi/(self.n*3)
        inner = [theta1*(i - 1) for i in range(1, n + 1)]
        middle, angle = [], 0
        for k in range(1, 2*n + 1):          # the angle is accumulated: thirds of a coarse quad, skipping its corners
            if k == 1:
                angle = theta2
            elif k % 2 == 0:
                angle += theta2
            else:
                angle += 2*theta2
            middle.append(angle)
        outer = [0 if k == 1 else theta3*(k - 1) for k in range(1, 3*n + 1)]
        conn = []
        for i in range(1, 4*n + 1):
            third = (i - (n + 1))//3
            if i <= n:                                    # coarse quads on the inner circle
                conn.append((2*i + n - 1, 2*i + n, (i + 1) if i != n else 1, i))
            elif (i - n) % 3 == 1:
                conn.append((i + 2*n, i + 2*n + 1, i - third, 1 + third))
            elif (i - n) % 3 == 2:
                conn.append((i + 2*n, i + 2*n + 1, i - third, i - 1 - third))
            elif i != 4*n:
                conn.append((i + 2*n, i + 2*n + 1, 2 + third, i - 1 - third))
            else:
                conn.append((i + 2*n, 1 + 3*n, 1, i - 1 - third))
        self._build([(self.inner_radius, inner, None), (self.outer_radius, middle, None), (self.r3, outer, None)], conn, 'Q',
                    Quad3D, 'AnnulusTransRingMesh')


class CylinderRingMesh(_RingBase):
    """`Mesh.py:1847-2036`: one course of a cylinder."""

    def __init__(self, radius, height, num_elements, thickness, material_name, model, kx_mod=1, ky_mod=1, origin=[0, 0, 0],
                 axis='Y', start_node='N1', start_element='Q1', element_type='Quad'):
        super().__init__(thickness, material_name, model, kx_mod, ky_mod, start_node=start_node, start_element=start_element)
        self.radius = radius
        self.height = height
        self.num_elements = num_elements
        self.Xo, self.Yo, self.Zo = origin[0], origin[1], origin[2]
        self.axis = axis
        self.element_type = element_type
        self.generate()

    def __repr__(self):
        return (f"CylinderRingMesh(material_name={self.material_name!r}, radius={self.radius}, height={self.height}, "
                f"thickness={self.thickness})")

    def generate(self):
        n = self.num_elements
        theta = 2*pi/n
        angles = [theta*(i - 1) for i in range(1, n + 1)]
        if self.element_type == 'Quad':
            prefix, cls = 'Q', Quad3D
        elif self.element_type == 'Rect':
            prefix, cls = 'R', Plate3D
        else:
            raise Exception('Invalid element type specified for cylinder ring mesh.')
        self._build([(self.radius, angles, None), (self.radius, angles, self.height)], self._band(n), prefix, cls,
                    'CylinderRingMesh')


class _Stacked(Mesh):
    """A mesh made of rings that share their boundary circles: later rings' nodes replace the equally named nodes of
    earlier ones (`dict.update`), the elements are re-pointed at the surviving objects, everything is filed in the model."""

    def _adopt(self, ring):
        self.nodes.update(ring.nodes)
        self.elements.update(ring.elements)

    def _file(self):
        for element in self.elements.values():
            element.i_node = self.nodes[element.i_node.name]
            element.j_node = self.nodes[element.j_node.name]
            element.m_node = self.nodes[element.m_node.name]
            element.n_node = self.nodes[element.n_node.name]
        for node in self.nodes.values():
            self.model.nodes[node.name] = node
        for element in self.elements.values():
            (self.model.quads if element.type.upper() == 'QUAD' else self.model.plates)[element.name] = element
        self.is_generated = True
        self.needs_update = False


class AnnulusMesh(_Stacked):
    """`Mesh.py:1126-1257`: rings from the inner to the outer radius; a ring whose quads have become wider than three mesh
    sizes is a transition ring that triples the count."""

    def __init__(self, mesh_size, outer_radius, inner_radius, thickness, material_name, model, kx_mod=1, ky_mod=1,
                 origin=[0, 0, 0], axis='Y', start_node='N1', start_element='Q1'):
        super().__init__(thickness, material_name, model, kx_mod, ky_mod, start_node, start_element)
        self.inner_radius = inner_radius
        self.outer_radius = outer_radius
        self.mesh_size = mesh_size
        self.origin = origin
        self.axis = axis
        self.num_quads_inner = None
        self.num_quads_outer = None

    def __repr__(self):
        return (f"AnnulusMesh(material_name={self.material_name!r}, outer_radius={self.outer_radius}, "
                f"inner_radius={self.inner_radius}, thickness={self.thickness})")

    def generate(self):
        if self.is_generated:
            self._remove_from_model()
        mesh_size, r_outer, r_inner = self.mesh_size, self.outer_radius, self.inner_radius
        n = int(self.start_node[1:])
        q = int(self.start_element[1:])
        n_circ = int(2*pi*r_inner/mesh_size)
        self.num_quads_outer = n_circ
        while round(r_inner, 10) < round(r_outer, 10):
            radial = r_outer - r_inner
            b_circ = 2*pi*r_inner/n_circ
            n_rad = int(radial/min(mesh_size, 3*b_circ))
            h_rad = radial/n_rad
            args = (self.thickness, self.material_name, self.model, self.kx_mod, self.ky_mod, self.origin, self.axis,
                    'N' + str(n), 'Q' + str(q))
            if b_circ > 3*mesh_size:
                ring = AnnulusTransRingMesh(r_inner + h_rad, r_inner, n_circ, *args)
                n += 3*n_circ
                q += 4*n_circ
                n_circ *= 3
                self.num_quads_outer = n_circ
            else:
                ring = AnnulusRingMesh(r_inner + h_rad, r_inner, n_circ, *args)
                n += n_circ
                q += n_circ
            self._adopt(ring)
            r_inner += h_rad
        self._file()


class FrustrumMesh(AnnulusMesh):
    """`Mesh.py:1623-1701`: an annulus whose nodes are lifted along the axis in proportion to their radius."""

    def __init__(self, mesh_size, large_radius, small_radius, height, thickness, material_name, model, kx_mod=1, ky_mod=1,
                 origin=[0, 0, 0], axis='Y', start_node='N1', start_element='Q1'):
        super().__init__(mesh_size, large_radius, small_radius, thickness, material_name, model, kx_mod, ky_mod, origin, axis,
                         start_node, start_element)
        self.height = height

    def __repr__(self):
        return (f"FrustrumMesh(material_name={self.material_name!r}, large_radius={self.outer_radius}, "
                f"small_radius={self.inner_radius}, height={self.height}, thickness={self.thickness})")

    def generate(self):
        super().generate()
        Xo, Yo, Zo = self.origin[0], self.origin[1], self.origin[2]
        span = self.outer_radius - self.inner_radius
        for node in self.nodes.values():
            X, Y, Z = node.X, node.Y, node.Z
            if self.axis == 'Y':
                node.Y += (((X - Xo)**2 + (Z - Zo)**2)**0.5 - self.outer_radius)/span*self.height
            elif self.axis == 'X':
                node.X += (((Y - Yo)**2 + (Z - Zo)**2)**0.5 - self.outer_radius)/span*self.height
            elif self.axis == 'Z':
                node.Z += (((X - Xo)**2 + (Y - Yo)**2)**0.5 - self.outer_radius)/span*self.height
            else:
                raise Exception('Invalid axis specified for frustrum mesh.')


class CylinderMesh(_Stacked):
    """`Mesh.py:1703-1845`: courses of `CylinderRingMesh` stacked along the axis (generated on construction)."""

    def __init__(self, mesh_size, radius, height, thickness, material_name, model, kx_mod=1, ky_mod=1, origin=[0, 0, 0],
                 axis='Y', start_node='N1', start_element='Q1', num_elements=None, element_type='Quad'):
        super().__init__(thickness, material_name, model, kx_mod, ky_mod, start_node, start_element)
        self.radius = radius
        self.h = height
        self.mesh_size = mesh_size
        self.origin = origin
        self.axis = axis
        self.num_elements = int(round(2*pi*radius/mesh_size, 0)) if num_elements is None else num_elements
        self.element_type = element_type
        self.generate()

    def __repr__(self):
        return f"CylinderMesh(material_name={self.material_name!r}, radius={self.radius}, height={self.h}, thickness={self.thickness})"

    def generate(self):
        if self.is_generated:
            self._remove_from_model()
        along = {'Y': 1, 'X': 0, 'Z': 2}[self.axis]
        y = self.origin[along]
        n = int(self.start_node[1:])
        q = int(self.start_element[1:])
        while round(y, 10) < round(self.h, 10):
            height = self.h - y
            h_y = height/max(int(abs(height)/self.mesh_size), 1)
            base = [0, 0, 0]
            base[along] = y
            self._adopt(CylinderRingMesh(self.radius, h_y, self.num_elements, self.thickness, self.material_name, self.model, 1, 1,
                                         base, self.axis, 'N' + str(n), 'Q' + str(q), self.element_type))
            n += self.num_elements
            q += self.num_elements
            y += h_y
        self._file()
