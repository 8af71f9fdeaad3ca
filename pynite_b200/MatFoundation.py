"""Mat foundation on Winkler springs: the caller-side generator of BASELINE config 4 (`Pynite/MatFoundation.py:26-223`).

A `RectangleMesh` in the XZ plane whose `generate()` also applies the registered point loads and gives every node a
compression-only soil spring `ks * tributary area` in DY (`MatFoundation.py:96-136`, direction '-': the springs the
tension/compression-only iteration switches on the device, `csrc/pn_tc.cu`).  Same names, arguments and arithmetic as the
reference (the tributary areas are accumulated plate by plate in element order, so the spring stiffnesses are
bit-identical); the node-by-plate scans of the reference (quadratic) are replaced by one pass over the elements.
"""
from __future__ import annotations

import numpy as np

from pynite_b200.Mesh import RectangleMesh


class MatFoundation(RectangleMesh):

    def __init__(self, name, mesh_size, length_X, length_Z, thickness, material_name, model, ks, origin=[0, 0, 0],
                 x_control=None, y_control=None):
        super().__init__(mesh_size, length_X, length_Z, thickness, material_name, model, 1, 1, origin, 'XZ', x_control, y_control)
        self.name = name
        self.ks = ks
        self.pt_loads = []          # [XZ_coord, direction, magnitude, case]
        self.needs_update = False
        self._incidence = None      # node name -> indices of the attached plates, in element order (soil_pressure)

    def __repr__(self):
        return f"MatFoundation(name={self.name!r}, self.width={self.width}, self.height={self.height}, self.ks={self.ks})"

    def add_rect_opening(self, name, X_min, Z_min, X_max, Z_max):
        """Opening by corner coordinates in the global X-Z plane (`MatFoundation.py:59-75`)."""
        super().add_rect_opening(name, X_min, Z_min, X_max - X_min, Z_max - Z_min)
        if self.is_generated:
            self.needs_update = True

    def add_mat_pt_load(self, XZ_coord, direction, magnitude, case='Case 1'):
        """A concentrated load at an X-Z coordinate; the coordinate becomes a control point of the mesh (`:77-94`)."""
        self.x_control.append(XZ_coord[0])
        self.y_control.append(XZ_coord[1])
        self.pt_loads.append([XZ_coord, direction, magnitude, case])
        if self.is_generated:
            self.needs_update = True

    def generate(self):
        super().generate()
        self._incidence = None
        model = self.model
        nodes = list(self.nodes.values())
        if self.pt_loads and nodes:
            X = np.array([n.X for n in nodes])
            Z = np.array([n.Z for n in nodes])
            hit = [np.isclose(X, pt[0][0]) & np.isclose(Z, pt[0][1]) for pt in self.pt_loads]
            for k, node in enumerate(nodes):                 # node-major, load-minor: the reference's order of node loads
                for pt, h in zip(self.pt_loads, hit):
                    if h[k]:
                        model.add_node_load(node.name, pt[1], pt[2], pt[3])
        # tributary area = a quarter of every attached plate, accumulated in element order (`:121-133`)
        trib = {id(n): 0 for n in nodes}
        for plate in self.elements.values():
            quarter = abs(plate.j_node.X - plate.i_node.X)*abs(plate.m_node.Z - plate.j_node.Z)/4
            for node in (plate.i_node, plate.j_node, plate.m_node, plate.n_node):
                trib[id(node)] += quarter
        for node in nodes:
            model.def_support_spring(node.name, 'DY', self.ks*trib[id(node)], '-')

    def soil_pressure(self, x, y, combo_name='Combo 1'):
        """Contact pressure at a point: nodal reactions over tributary areas, bilinearly interpolated inside the first plate
        (element order) whose X-Z bounds contain the point (`MatFoundation.py:138-223`)."""
        plates = list(self.elements.values())
        if not plates:
            return 0
        if self._incidence is None:
            inc = {}
            for k, p in enumerate(plates):
                for node in (p.i_node, p.j_node, p.m_node, p.n_node):
                    inc.setdefault(node.name, []).append(k)
            bx = np.array([[min(p.i_node.X, p.j_node.X, p.m_node.X, p.n_node.X), max(p.i_node.X, p.j_node.X, p.m_node.X, p.n_node.X),
                            min(p.i_node.Z, p.j_node.Z, p.m_node.Z, p.n_node.Z), max(p.i_node.Z, p.j_node.Z, p.m_node.Z, p.n_node.Z)]
                           for p in plates])
            self._incidence = (inc, bx)
        inc, bx = self._incidence
        inside = np.nonzero((x >= bx[:, 0] - 1e-9) & (x <= bx[:, 1] + 1e-9) & (y >= bx[:, 2] - 1e-9) & (y <= bx[:, 3] + 1e-9))[0]
        if inside.size == 0:
            return 0
        plate = plates[int(inside[0])]
        if self.model.solution is None:
            raise RuntimeError(f'No nodal reactions found for combination {combo_name!r}. Run analysis first.')

        def plate_area(p):
            return abs(p.j_node.X - p.i_node.X)*abs(p.n_node.Z - p.i_node.Z)

        def tributary_area(node):
            area = 0.0
            for k in inc[node.name]:
                area += plate_area(plates[k])/4.0
            return area
        p_i = plate.i_node.RxnFY[combo_name]/tributary_area(plate.i_node)
        p_j = plate.j_node.RxnFY[combo_name]/tributary_area(plate.j_node)
        p_m = plate.m_node.RxnFY[combo_name]/tributary_area(plate.m_node)
        p_n = plate.n_node.RxnFY[combo_name]/tributary_area(plate.n_node)
        xi = (x - plate.i_node.X)/(plate.j_node.X - plate.i_node.X)
        eta = (y - plate.i_node.Z)/(plate.n_node.Z - plate.i_node.Z)
        xi = min(max(xi, 0.0), 1.0)
        eta = min(max(eta, 0.0), 1.0)
        N_i = (1 - xi)*(1 - eta)
        N_j = xi*(1 - eta)
        N_m = xi*eta
        N_n = (1 - xi)*eta
        return N_i*p_i + N_j*p_j + N_m*p_m + N_n*p_n
