"""Shear wall builder (`Pynite/ShearWall.py`): a rectangle mesh with openings, flanges, supports and story loads, the piers
and coupling beams the openings cut it into, and their section forces from the quads' end forces (`pn_element_forces`).

Host-side model building only; the analysis is the GPU path's.  Names, node and element order, supports and loads are those
of the reference (golden cases in tests/golden/rect_meshes.npz and shear_walls.npz), including three things a reader might
take for slips here: the wall mesh is created with thickness 1 and `kx_mod = thickness` (`ShearWall.py:201`; the material
regions then set every wall quad's real thickness, `kx_mod` stays), the material / pier / beam passes look at every quad of
the *model*, and the stiffness combination's tag is the string 'stiffness' (`:105`).

Piers and coupling beams are one construction run along different axes: cut the wall into the cells of the openings' grid,
drop the cells inside openings, merge neighbours across one axis and then along the other -- in dictionary order, restarting
after every merge, which is what fixes the outcome when several merges are possible -- renumber, collect the quads whose
centre lies inside.  The reference spells the two out separately (`_identify_piers` :344-505, `_identify_coupling_beams`
:507-677); here they are `_regions(...)` with the axes swapped.  Plotting (`draw_*`, `screenshots`) is not rebuilt.
"""
from __future__ import annotations

from math import isclose


def _global2local(X, Y, Z, origin=(0, 0, 0), plane='XY'):
    """Wall coordinates (x along the wall, y up, z out of plane) of a model point (`ShearWall.py:904-922`)."""
    dX, dY, dZ = X - origin[0], Y - origin[1], Z - origin[2]
    if plane == 'XY':
        return [dX, dY, dZ]
    if plane == 'YZ':
        return [dZ, dY, -dX]
    if plane == 'XZ':
        return [dX, dZ, -dY]
    raise Exception('Invalid plane specified for coordinate transformation')


def _r(v):
    return round(v, 10)


def _distinct(values):
    """Sorted values with near-duplicates removed: of a run of close values the last one stays."""
    values = sorted(values)
    return [v for v, nxt in zip(values, values[1:]) if not isclose(v, nxt)] + [values[-1]]


class _Region:
    """A rectangular part of the wall: `a` is its extent along the wall, `b` its extent in elevation."""

    def __init__(self, name, x, y, a, b, shear_wall):
        self.name = name
        self.x = x
        self.y = y
        self._a = a
        self._b = b
        self.shear_wall = shear_wall
        self.plane = shear_wall.plane
        self.origin = shear_wall.origin
        self.plates = []

    def _local(self, node):
        return _global2local(node.X, node.Y, node.Z, self.origin, self.plane)


class Pier(_Region):
    """Wall pier (`ShearWall.py:924-985`)."""
    width = property(lambda self: self._a, lambda self, v: setattr(self, '_a', v))
    height = property(lambda self: self._b, lambda self, v: setattr(self, '_b', v))

    def __repr__(self):
        return f"Pier(name={self.name!r}, width={self.width}, height={self.height})"

    def sum_forces(self, combo_name='Combo 1'):
        """(P, M, V, M/(V L)) on the pier's bottom section: the local end forces of the quads standing on it, nodes i and j
        (`:944-985`).  Flange quads (i and j at the same x) contribute their out-of-plane nodal force as shear."""
        P = M = V = 0
        mid = self.x + self.width/2
        for plate in self.plates:
            xi, yi, _ = self._local(plate.i_node)
            xj, _, _ = self._local(plate.j_node)
            if not isclose(yi, self.y):
                continue
            f = plate.f(combo_name)
            P += -f[1][0] - f[7][0]
            M += -f[1][0]*(xi - mid) - f[7][0]*(xj - mid)
            k = 2 if isclose(xi, xj) else 0
            V += f[k][0] + f[6 + k][0]
        return P, M, V, M/(V*self.width)


class CouplingBeam(_Region):
    """Coupling beam (`ShearWall.py:988-1052`)."""
    length = property(lambda self: self._a, lambda self, v: setattr(self, '_a', v))
    height = property(lambda self: self._b, lambda self, v: setattr(self, '_b', v))

    def __repr__(self):
        return f"CouplingBeam(name={self.name!r}, length={self.length}, height={self.height})"

    def sum_forces(self, combo_name='Combo 1'):
        """(P, M, V, M/(V H)) on the beam's left section: nodes i and n of the wall quads starting there (`:1007-1052`)."""
        P = M = V = 0
        mid = self.y + self.height/2
        for plate in self.plates:
            xi, yi, _ = self._local(plate.i_node)
            xj, _, _ = self._local(plate.j_node)
            _, yn, _ = self._local(plate.n_node)
            if not isclose(xi, self.x) or isclose(xi, xj):
                continue
            f = plate.f(combo_name)
            P += -f[0][0] - f[18][0]
            M += -f[0][0]*(yi - mid) - f[18][0]*(yn - mid)
            V += f[1][0] + f[19][0]
        return P, M, V, M/(V*self.height)


class ShearWall:

    def __init__(self, model, name, mesh_size, length, height, thickness, material_name, ky_mod=0.35, origin=[0, 0, 0],
                 plane='XY'):
        self.model = model
        self.name = name
        self.mesh_size = mesh_size
        self.L = length
        self.H = height
        self.ky_mod = ky_mod
        self.origin = origin
        self.plane = plane
        self._openings = []      # [name, x_start, y_start, width, height, tie]
        self._flanges = []       # [thickness, width, x, y_start, y_end, material, side]
        self._supports = []      # [elevation, x_start, x_end]
        self._stories = []       # [name, elevation, x_start, x_end]
        self._shears = []        # [story, force, case]
        self._axials = []        # [story, force, case]
        self._materials = []     # [material, thickness, x_start, x_end, y_start, y_end]
        self.piers = {}
        self.coupling_beams = {}
        self.is_generated = False
        self.needs_update = False
        self.asign_material(material_name, thickness)

    def __repr__(self):
        return f"ShearWall(name={self.name!r}, length={self.L}, height={self.H})"

    def _touch(self):
        if self.is_generated:
            self.needs_update = True

    # ---- definition (`ShearWall.py:49-116`) -------------------------------------------------------------------------
    def asign_material(self, name, t, x_start=None, x_end=None, y_start=None, y_end=None):
        self._materials.append([name, t, 0 if x_start is None else x_start, self.L if x_end is None else x_end,
                                0 if y_start is None else y_start, self.H if y_end is None else y_end])
        self._touch()

    def add_opening(self, name, x_start, y_start, width, height, tie=None):
        self._openings.append([name, x_start, y_start, width, height, None])      # (the tie is not used by the reference either)
        self._touch()

    def add_flange(self, thickness, width, x, y_start, y_end, material, side):
        self._flanges.append([thickness, width, x, y_start, y_end, material, side])
        self._touch()

    def add_support(self, elevation=None, x_start=None, x_end=None):
        elevation = 0 if elevation is None else elevation
        x_start = 0 if x_start is None else x_start
        x_end = self.L if x_end is None else x_end
        if elevation < 0 or elevation > self.H or x_start < 0 or x_start > self.L or x_end < 0 or x_end > self.L:
            raise Exception(f'Support for shear wall `{self.name}` falls outside the extents of the wall and is invalid.')
        self._supports.append([elevation, x_start, x_end])
        self._touch()

    def add_story(self, story_name, elevation, x_start=None, x_end=None):
        """A diaphragm level.  Every story also gets a 100-unit shear in a case and a combination of its own
        ('Stiffness: <story>'), which `stiffness()` reads."""
        self._stories.append([story_name, self.H if elevation is None else elevation, 0 if x_start is None else x_start,
                              self.L if x_end is None else x_end])
        self._touch()
        self.model.add_load_combo('Stiffness: ' + story_name, {story_name: 1.0}, 'stiffness')
        self.add_shear(story_name, 100, case=story_name)

    def add_shear(self, story_name, force, case='Case 1'):
        self._shears.append([story_name, force, case])
        self._touch()

    def add_axial(self, story_name, force, case='Case 1'):
        self._axials.append([story_name, force, case])
        self._touch()

    # ---- generation (`ShearWall.py:118-342`) ------------------------------------------------------------------------
    def _flange_name(self, i):
        return self.name + ' Flg ' + str(i + 1)

    def _remove_from_model(self):
        for name in [self.name] + [self._flange_name(i) for i in range(len(self._flanges))]:
            if name in self.model.meshes:
                self.model.meshes[name]._remove_from_model()
                del self.model.meshes[name]

    def generate(self):
        model = self.model
        if self.is_generated:
            self._remove_from_model()
        # every edge of a material region, flange, support, story and opening becomes a mesh control line
        x_control, y_control, z_control = [0, self.L], [0, self.H], [0]
        for _, _, x0, x1, y0, y1 in self._materials:
            x_control += [x0, x1]
            y_control += [y0, y1]
        for _, b, x, y0, y1, _, side in self._flanges:
            z_control.append(b if side == '+z' else -b)
            x_control.append(x)
            y_control += [y0, y1]
        for elevation, x0, x1 in self._supports:
            x_control += [x0, x1]
            y_control.append(elevation)
        for _, elevation, x0, x1 in self._stories:
            x_control += [x0, x1]
            y_control.append(elevation)
        for _, _, y0, _, h, _ in self._openings:
            y_control += [y0, y0 + h]
        material_name, thickness = self._materials[0][:2]
        model.add_rectangle_mesh(self.name, self.mesh_size, self.L, self.H, 1, material_name, thickness, self.ky_mod,
                                 self.origin, self.plane, x_control=x_control, y_control=y_control)
        if 'Tie' not in model.materials:
            model.add_material('Tie', 1, 1, 0, 0)
        for name, x0, y0, w, h, _ in self._openings:
            model.meshes[self.name].add_rect_opening(name, x0, y0, w, h)
        Xo, Yo, Zo = self.origin
        for i, (t, b, x, y0, y1, material, side) in enumerate(self._flanges):
            if side == '+z':
                z = 0
                fx = [v for v in z_control if _r(v) >= 0 and _r(v) <= b]
            else:
                z = -b
                fx = [b + v for v in z_control if _r(v) <= 0 and _r(v) >= -b]
            fy = [v - y0 for v in y_control if _r(v) >= _r(y0) and _r(v) <= _r(y1)]
            if self.plane == 'XY':
                corner, flange_plane = [Xo + x, Yo + y0, Zo + z], 'YZ'
            elif self.plane == 'XZ':
                corner, flange_plane = [Xo + x, Yo + z, Zo + y0], 'YZ'
            else:
                corner, flange_plane = [Xo + z, Yo + y0, Zo + x], 'XY'
            model.add_rectangle_mesh(self._flange_name(i), self.mesh_size, b, y1 - y0, t, material, 1, self.ky_mod, corner,
                                     flange_plane, fx, fy)
        model.meshes[self.name].generate()
        for i in range(len(self._flanges)):
            model.meshes[self._flange_name(i)].generate()
        model.merge_duplicate_nodes()

        local = {id(nd): _global2local(nd.X, nd.Y, nd.Z, self.origin, self.plane) for nd in model.nodes.values()}
        for quad in model.quads.values():                    # material regions: in-plane quads lying inside
            (xi, yi, _), (xj, _, _), (xm, ym, _) = local[id(quad.i_node)], local[id(quad.j_node)], local[id(quad.m_node)]
            if isclose(xi, xj):
                continue
            for name, t, x0, x1, y0, y1 in self._materials:
                if _r(xi) >= _r(x0) and _r(xm) <= _r(x1) and _r(yi) >= _r(y0) and _r(ym) <= _r(y1):
                    quad.E = model.materials[name].E
                    quad.nu = model.materials[name].nu
                    quad.t = t
        for elevation, x0, x1 in self._supports:
            for node in model.nodes.values():
                x, y, _ = local[id(node)]
                if isclose(y, elevation) and _r(x) >= _r(x0) and _r(x) <= _r(x1):
                    model.def_support(node.name, True, True, True, True, True, True)
        for story_name, elevation, x0, x1 in self._stories:
            on_line = [nd for nd in model.nodes.values()
                       if isclose(local[id(nd)][1], elevation) and x0 <= local[id(nd)][0] <= x1 and isclose(local[id(nd)][2], 0)]
            for node in on_line:
                for story, force, case in self._shears:
                    if self.plane == 'XY':
                        direction = 'FX'
                    elif self.plane == 'YZ':
                        direction = 'FZ'
                    if story == story_name:
                        model.add_node_load(node.name, direction, force/len(on_line), case)
                for story, force, case in self._axials:
                    if story == story_name:
                        model.add_node_load(node.name, 'FY', -force/len(on_line), case)
        self.piers = self._regions(Pier, 'P', along_first=True)
        self.coupling_beams = self._regions(CouplingBeam, 'B', along_first=False)
        self.is_generated = True
        self.needs_update = False

    def _identify_piers(self):
        self.piers = self._regions(Pier, 'P', along_first=True)

    def _identify_coupling_beams(self):
        self.coupling_beams = self._regions(CouplingBeam, 'B', along_first=False)

    def _regions(self, cls, prefix, along_first):
        """Piers (`along_first`: strips along the wall, cut by elevation; merged sideways, then upwards) or coupling beams
        (strips by elevation, cut along the wall; merged upwards, then sideways; none at the base)."""
        xs = _distinct([0, self.L] + [v for o in self._openings for v in (o[1], o[1] + o[3])])
        ys = _distinct([0, self.H] + [v for o in self._openings for v in (o[2], o[2] + o[4])])
        spans_x, spans_y = list(zip(xs, xs[1:])), list(zip(ys, ys[1:]))
        cells = ([(x0, y0, x1 - x0, y1 - y0) for x0, x1 in spans_x for y0, y1 in spans_y] if along_first else
                 [(x0, y0, x1 - x0, y1 - y0) for y0, y1 in spans_y for x0, x1 in spans_x])
        table = {}
        for k, (x, y, a, b) in enumerate(cells):
            inside = any(_r(x) >= _r(o[1]) and _r(x + a) <= _r(o[1] + o[3]) and _r(y) >= _r(o[2]) and _r(y + b) <= _r(o[2] + o[4])
                         for o in self._openings)
            if not inside:
                table[k] = cls(prefix + str(k + 1), x, y, a, b, self)

        def sideways(p, q):
            return isclose(p.y, q.y) and isclose(p.x + p._a, q.x) and isclose(p._b, q._b)

        def upwards(p, q):
            return isclose(p.x, q.x) and isclose(p.y + p._b, q.y) and isclose(p._a, q._a)

        for touching in ((sideways, upwards) if along_first else (upwards, sideways)):
            merged = True
            while merged:
                merged = False
                for k1, p in list(table.items()):
                    for k2, q in list(table.items()):
                        if k1 != k2 and touching(p, q):
                            if touching is sideways:
                                p._a = p._a + q._a
                            else:
                                p._b = p._b + q._b
                            del table[k2]
                            merged = True
                            break
                    if merged:
                        break
        regions = [p for p in table.values() if along_first or p.y != 0]
        for k, region in enumerate(regions):
            region.name = prefix + str(k + 1)
        for quad in self.model.quads.values():
            xi, yi, _ = _global2local(quad.i_node.X, quad.i_node.Y, quad.i_node.Z, self.origin, self.plane)
            xm, ym, _ = _global2local(quad.m_node.X, quad.m_node.Y, quad.m_node.Z, self.origin, self.plane)
            xc, yc = _r((xi + xm)/2), _r((yi + ym)/2)
            for region in regions:
                if _r(region.x) <= xc <= _r(region.x + region._a) and _r(region.y) <= yc <= _r(region.y + region._b):
                    region.plates.append(quad)
        return {region.name: region for region in regions}

    def _sort_openings(self):
        """Openings ordered by their left edge, ties by their bottom edge (two stable passes, `:739-753`)."""
        self._openings.sort(key=lambda o: o[2])
        self._openings.sort(key=lambda o: o[1])

    # ---- results ----------------------------------------------------------------------------------------------------
    def stiffness(self, story_name):
        """100 / the largest global DX of the story's nodes under the story's own 100-unit shear (`:755-794`)."""
        story = next((s for s in self._stories if s[0] == story_name), None)
        if story is None:
            raise Exception(f'Unable to obtain stiffness. Story {story_name} does not exist.')
        d_max = 0
        for node in self.model.nodes.values():
            x, y, z = _global2local(node.X, node.Y, node.Z, self.origin, self.plane)
            if _r(x) >= _r(story[2]) and _r(x) <= _r(story[3]) and isclose(story[1], y) and isclose(z, 0):
                d_max = max(d_max, node.DX['Stiffness: ' + story_name])
        return 100/d_max

    def _print(self, title, header, rows):
        widths = [max(len(str(v)) for v in col) for col in zip(header, *rows)]
        rule = '+' + '+'.join('-'*(w + 2) for w in widths) + '+'
        line = lambda row: '|' + '|'.join(' ' + str(v).center(w) + ' ' for v, w in zip(row, widths)) + '|'   # noqa: E731
        print('+' + '-'*(len(title) + 2) + '+')
        print('| ' + title + ' |')
        print('+' + '-'*(len(title) + 2) + '+')
        print('\n'.join([rule, line(header), rule] + [line(r) for r in rows] + [rule]))

    def print_piers(self, combo_name='Combo 1'):
        rows = []
        for pier in self.piers.values():
            P, M, V, M_VL = pier.sum_forces(combo_name)
            rows.append([pier.name, pier.width, pier.height, M_VL, V, M, P])
        self._print('Wall Pier Results', ['ID', 'Length', 'Height', 'M/(VL)', 'V', 'M', 'P'], rows)

    def print_coupling_beams(self, combo_name='Combo 1'):
        rows = []
        for beam in self.coupling_beams.values():
            P, M, V, M_VH = beam.sum_forces(combo_name)
            rows.append([beam.name, beam.length, beam.height, M_VH, V, M, P])
        self._print('Wall Coupling Beam Results', ['ID', 'Length', 'Height', 'M/(VH)', 'V', 'M', 'P'], rows)
