"""Model builders shared by the golden-fixture generator, the parity tests, smoke() and bench.py.

Every builder takes the `FEModel3D` class to instantiate, so the same script builds the model with
the reference (`/root/reference/Pynite`, in the fixture generator only) and with `pynite_b200`.
Only the reference's public builder API is used (`FEModel3D.py:224-1507`).
"""
import numpy as np


def space_frame(FE):
    """BASELINE config 1: `Examples/Space Frame - Nodal Loads 1.py` (Logan, Example 5.8)."""
    m = FE()
    m.add_node('N1', 0, 0, 0)
    m.add_node('N2', -100, 0, 0)
    m.add_node('N3', 0, 0, -100)
    m.add_node('N4', 0, -100, 0)
    for n in ('N2', 'N3', 'N4'):
        m.def_support(n, True, True, True, True, True, True)
    m.add_section('MySection', 10, 100, 100, 50)
    m.add_material('Steel', 30000, 10000, 0.3, 2.836e-4)
    m.add_member('M1', 'N2', 'N1', 'Steel', 'MySection')
    m.add_member('M2', 'N3', 'N1', 'Steel', 'MySection')
    m.add_member('M3', 'N4', 'N1', 'Steel', 'MySection')
    m.add_node_load('N1', 'FY', -50)
    m.add_node_load('N1', 'MX', -1000)
    return m


def portal_frame(FE):
    """Gable frame with inclined rafters, roll angle, end releases, local and global member loads,
    three load cases and four combinations."""
    m = FE()
    m.add_material('Steel', 29000, 11200, 0.3, 2.836e-4)
    m.add_section('Col', 14.4, 53.4, 272.0, 1.0)
    m.add_section('Raf', 10.3, 15.3, 510.0, 0.5)
    m.add_section('Brace', 3.0, 2.0, 2.0, 0.1)
    m.add_node('A', 0, 0, 0)
    m.add_node('B', 0, 180, 0)
    m.add_node('C', 144, 252, 0)
    m.add_node('D', 288, 180, 0)
    m.add_node('E', 288, 0, 0)
    m.add_node('F', 144, 180, 96)      # out-of-plane point tied by a skew brace
    m.add_node('G', 144, 0, 96)
    m.def_support('A', True, True, True, True, True, False)
    m.def_support('E', True, True, True, True, True, True)
    m.def_support('G', True, True, True, False, False, False)
    m.add_member('AB', 'A', 'B', 'Steel', 'Col')
    m.add_member('BC', 'B', 'C', 'Steel', 'Raf', rotation=0.0)
    m.add_member('CD', 'C', 'D', 'Steel', 'Raf')
    m.add_member('ED', 'E', 'D', 'Steel', 'Col', rotation=90.0)
    m.add_member('CF', 'C', 'F', 'Steel', 'Brace', rotation=30.0)
    m.add_member('GF', 'G', 'F', 'Steel', 'Col')
    m.add_member('BF', 'B', 'F', 'Steel', 'Brace')
    m.def_releases('CF', Ryi=True, Rzi=True)
    m.def_releases('BF', Rzi=True, Ryj=True, Rzj=True)
    m.add_member_dist_load('BC', 'Fy', -0.05, -0.08, case='D')
    m.add_member_dist_load('CD', 'FY', -0.06, -0.06, 10.0, 120.0, case='D')
    m.add_member_dist_load('BC', 'Fx', 0.01, 0.02, case='L')
    m.add_member_dist_load('AB', 'FX', 0.02, 0.03, case='W')
    m.add_member_pt_load('CD', 'Fy', -3.0, 60.0, case='L')
    m.add_member_pt_load('AB', 'FZ', 1.5, 90.0, case='W')
    m.add_member_pt_load('ED', 'Mz', 40.0, 100.0, case='L')
    m.add_member_pt_load('BC', 'MX', 25.0, 50.0, case='W')
    m.add_member_pt_load('CF', 'Fx', 2.0, 40.0, case='W')
    m.add_member_pt_load('BF', 'Fz', -1.0, 70.0, case='L')
    m.add_node_load('C', 'FX', 4.0, case='W')
    m.add_node_load('F', 'FY', -6.0, case='D')
    m.add_node_load('F', 'MZ', 30.0, case='L')
    m.add_load_combo('1.4D', {'D': 1.4})
    m.add_load_combo('1.2D+1.6L', {'D': 1.2, 'L': 1.6})
    m.add_load_combo('1.2D+L+W', {'D': 1.2, 'L': 1.0, 'W': 1.0})
    m.add_load_combo('0.9D-W', {'D': 0.9, 'W': -1.0})
    return m


def continuous_beam(FE):
    """Physical member through interior nodes (`PhysMember.descritize`), support settlement, node support
    springs, a two-node spring element, clipped distributed and point loads."""
    m = FE()
    m.add_material('Steel', 29000, 11200, 0.3, 2.836e-4)
    m.add_section('W', 9.12, 37.1, 375.0, 0.5)
    xs = [0, 120, 300, 420, 600]
    for i, x in enumerate(xs):
        m.add_node('N' + str(i + 1), x, 0, 0)
    m.add_node('S1', 300, -20, 0)
    m.add_member('B', 'N1', 'N5', 'Steel', 'W')
    m.def_releases('B', Rzj=True)
    m.def_support('N1', True, True, True, True, True, True)
    m.def_support('N2', False, True, True, False, False, False)
    m.def_support('N4', False, True, True, False, False, False)
    m.def_support('N5', True, True, True, True, False, True)
    m.def_support('S1', True, True, True, True, True, True)
    m.add_spring('SP', 'N3', 'S1', 55.0)
    m.def_support_spring('N3', 'DZ', 12.0)
    m.def_support_spring('N3', 'RX', 500.0)
    m.def_node_disp('N2', 'DY', -0.25)
    m.def_node_disp('N5', 'RY', 0.002)
    m.add_member_dist_load('B', 'Fy', -0.02, -0.05, 60.0, 480.0, case='D')
    m.add_member_dist_load('B', 'Fz', 0.01, 0.01, case='D')
    m.add_member_pt_load('B', 'Fy', -8.0, 210.0, case='L')
    m.add_member_pt_load('B', 'Fy', -2.0, 600.0, case='L')
    m.add_member_pt_load('B', 'Mz', 120.0, 350.0, case='L')
    m.add_node_load('N3', 'FZ', 1.0, case='L')
    m.add_node_load('N3', 'FZ', 0.5, case='L')
    m.add_load_combo('D', {'D': 1.0})
    m.add_load_combo('D+L', {'D': 1.0, 'L': 1.0})
    return m


def _grid_nodes(m, nx, ny, a, b, plane, origin=(0.0, 0.0, 0.0)):
    names = {}
    k = 1
    for r in range(ny + 1):
        for c in range(nx + 1):
            u, v = c*a, r*b
            if plane == 'XY':
                p = (origin[0] + u, origin[1] + v, origin[2])
            elif plane == 'XZ':
                p = (origin[0] + u, origin[1], origin[2] + v)
            else:
                p = (origin[0], origin[1] + u, origin[2] + v)
            names[(c, r)] = m.add_node('N' + str(k), *p)
            k += 1
    return names


def quad_mesh(FE, nx=6, ny=5, a=12.0, b=10.0, t=8.0, plane='XY', kx_mod=1.0, ky_mod=1.0, seed=1):
    """Rectangular DKMQ quad mesh, perimeter pinned (DX, DY, DZ), pressures + nodal loads, 3 combos.
    Node order row-major and CCW i-j-m-n connectivity like `RectangleMesh` (`Mesh.py:867-943`)."""
    m = FE()
    m.add_material('Conc', 3605.0, 1540.6, 0.17, 8.68e-5)
    N = _grid_nodes(m, nx, ny, a, b, plane)
    rng = np.random.default_rng(seed)
    k = 1
    for r in range(ny):
        for c in range(nx):
            name = m.add_quad('Q' + str(k), N[(c, r)], N[(c + 1, r)], N[(c + 1, r + 1)], N[(c, r + 1)], t, 'Conc',
                              kx_mod, ky_mod)
            m.add_quad_surface_pressure(name, 0.001, case='D')
            if (r + c) % 2 == 0:
                m.add_quad_surface_pressure(name, 0.0005, case='L')
            k += 1
    for r in range(ny + 1):
        for c in range(nx + 1):
            if r in (0, ny) or c in (0, nx):
                m.def_support(N[(c, r)], True, True, True, False, False, False)
    normal = {'XY': 'FZ', 'XZ': 'FY', 'YZ': 'FX'}[plane]
    for _ in range(4):
        c, r = int(rng.integers(1, nx)), int(rng.integers(1, ny))
        m.add_node_load(N[(c, r)], normal, float(-rng.uniform(0, 10)), case='P')
    m.add_load_combo('C1', {'D': 1.4})
    m.add_load_combo('C2', {'D': 1.2, 'L': 1.6})
    m.add_load_combo('C3', {'D': 1.2, 'L': 0.5, 'P': 1.0})
    return m


def plate_mesh(FE, nx=5, ny=4, a=15.0, b=9.0, t=6.0, plane='XZ'):
    """Rectangular Plate3D mesh, two opposite edges fixed, pressure + nodal loads."""
    m = FE()
    m.add_material('Conc', 3605.0, 1540.6, 0.17, 8.68e-5)
    N = _grid_nodes(m, nx, ny, a, b, plane, origin=(5.0, -3.0, 2.0))
    k = 1
    for r in range(ny):
        for c in range(nx):
            name = m.add_plate('P' + str(k), N[(c, r)], N[(c + 1, r)], N[(c + 1, r + 1)], N[(c, r + 1)], t, 'Conc')
            m.add_plate_surface_pressure(name, 0.002, case='D')
            if c % 2 == 0:
                m.add_plate_surface_pressure(name, -0.001, case='L')
            k += 1
    for r in range(ny + 1):
        m.def_support(N[(0, r)], True, True, True, True, True, True)
        m.def_support(N[(nx, r)], True, True, True, False, False, False)
    normal = {'XY': 'FZ', 'XZ': 'FY', 'YZ': 'FX'}[plane]
    m.add_node_load(N[(2, 2)], normal, -7.5, case='L')
    m.add_load_combo('C1', {'D': 1.0})
    m.add_load_combo('C2', {'D': 1.2, 'L': 1.6})
    return m


def skew_shell(FE):
    """Quads in general position: a folded, skewed strip (non-rectangular elements, arbitrary orientation)
    with orthotropic membrane modifiers on one element."""
    m = FE()
    m.add_material('Conc', 4000.0, 1700.0, 0.2, 8.68e-5)
    pts = [(0, 0, 0), (10, 1, 2), (21, -1, 3), (30, 2, 7),
           (1, 9, 1), (11, 11, 4), (20, 10, 5), (31, 12, 10),
           (-1, 20, 5), (9, 19, 8), (22, 21, 9), (29, 20, 15)]
    for i, p in enumerate(pts):
        m.add_node('N' + str(i + 1), *p)
    quads = [(1, 2, 6, 5), (2, 3, 7, 6), (3, 4, 8, 7), (5, 6, 10, 9), (6, 7, 11, 10), (7, 8, 12, 11)]
    for k, q in enumerate(quads):
        kx, ky = (0.35, 0.7) if k == 4 else (1.0, 1.0)
        name = m.add_quad('Q' + str(k + 1), *['N' + str(i) for i in q], 0.75, 'Conc', kx, ky)
        m.add_quad_surface_pressure(name, 0.01*(k + 1), case='D')
    for n in (1, 2, 3, 4):
        m.def_support('N' + str(n), True, True, True, True, True, True)
    m.add_node_load('N10', 'FZ', -3.0, case='D')
    m.add_node_load('N12', 'FX', 1.0, case='W')
    m.add_load_combo('D', {'D': 1.0})
    m.add_load_combo('D+W', {'D': 1.2, 'W': 1.0})
    return m


def moment_frame(FE, bays_x=2, bays_z=1, stories=3, bay=360.0, story=156.0, interior=0, n_combos=4, seed=7):
    """3-D moment frame like BASELINE config 2 / config 5: columns along Y, beams along X and Z,
    fixed bases, gravity line loads on beams (D, L), lateral nodal loads (WX, WZ).  `interior` extra nodes
    per column (P-little-delta, `docs/source/PDelta.rst:39`)."""
    m = FE()
    m.add_material('Steel', 29000.0, 11200.0, 0.3, 2.836e-4)
    m.add_section('Col', 50.0, 800.0, 2000.0, 30.0)
    m.add_section('Beam', 20.0, 100.0, 1200.0, 3.0)
    name = {}
    k = 1
    for s in range(stories + 1):
        for iz in range(bays_z + 1):
            for ix in range(bays_x + 1):
                name[(ix, s, iz)] = m.add_node('N' + str(k), ix*bay, s*story, iz*bay)
                k += 1
    for s in range(stories):
        for iz in range(bays_z + 1):
            for ix in range(bays_x + 1):
                for q in range(interior):
                    m.add_node('N' + str(k), ix*bay, (s + (q + 1)/(interior + 1))*story, iz*bay)
                    k += 1
    for iz in range(bays_z + 1):
        for ix in range(bays_x + 1):
            m.def_support(name[(ix, 0, iz)], True, True, True, True, True, True)
    e = 1
    for s in range(stories):
        for iz in range(bays_z + 1):
            for ix in range(bays_x + 1):
                m.add_member('M' + str(e), name[(ix, s, iz)], name[(ix, s + 1, iz)], 'Steel', 'Col')
                e += 1
    for s in range(1, stories + 1):
        for iz in range(bays_z + 1):
            for ix in range(bays_x):
                nm = m.add_member('M' + str(e), name[(ix, s, iz)], name[(ix + 1, s, iz)], 'Steel', 'Beam')
                m.add_member_dist_load(nm, 'Fy', -0.1, -0.1, case='D')
                m.add_member_dist_load(nm, 'Fy', -0.06, -0.06, case='L')
                e += 1
        for iz in range(bays_z):
            for ix in range(bays_x + 1):
                nm = m.add_member('M' + str(e), name[(ix, s, iz)], name[(ix, s, iz + 1)], 'Steel', 'Beam')
                m.add_member_dist_load(nm, 'Fy', -0.1, -0.1, case='D')
                m.add_member_dist_load(nm, 'Fy', -0.06, -0.06, case='L')
                e += 1
    for s in range(1, stories + 1):
        for iz in range(bays_z + 1):
            m.add_node_load(name[(0, s, iz)], 'FX', 2.0*s, case='WX')
        for ix in range(bays_x + 1):
            m.add_node_load(name[(ix, s, 0)], 'FZ', 2.0*s, case='WZ')
    table = [(1.4, 0, 0, 0), (1.2, 1.6, 0, 0), (1.2, 1.0, 1.0, 0), (1.2, 1.0, -1.0, 0),
             (1.2, 1.0, 0, 1.0), (1.2, 1.0, 0, -1.0), (0.9, 0, 1.0, 0), (0.9, 0, -1.0, 0),
             (0.9, 0, 0, 1.0), (0.9, 0, 0, -1.0), (1.2, 0.5, 0.5, 0.5), (1.2, 0.5, -0.5, 0.5),
             (1.2, 0.5, 0.5, -0.5), (1.2, 0.5, -0.5, -0.5), (1.0, 1.0, 0.7, 0.3), (1.0, 1.0, 0.3, 0.7)]
    rng = np.random.default_rng(seed)
    for c in range(n_combos):
        if c < len(table):
            f = table[c]
        else:
            f = tuple(np.round(rng.uniform(-1.0, 1.6, size=4), 2))
            f = (abs(f[0]) + 0.5, abs(f[1]), f[2], f[3])
        m.add_load_combo('C' + str(c + 1), {'D': float(f[0]), 'L': float(f[1]), 'WX': float(f[2]), 'WZ': float(f[3])})
    return m


def pdelta_column(FE):
    """AISC benchmark case 2 flavour (`Testing/test_AISC_PDelta_benchmarks.py:138-270`): cantilever split by
    interior nodes, axial + lateral load, four axial load levels."""
    m = FE()
    m.add_material('Steel', 29000.0, 11200.0, 0.3, 2.836e-4)
    m.add_section('W14', 26.5, 362.0, 999.0, 5.45)
    n = 5
    for i in range(n):
        m.add_node('N' + str(i + 1), 0, i*28.0*12/(n - 1), 0)
    m.add_member('M1', 'N1', 'N' + str(n), 'Steel', 'W14')
    m.def_support('N1', True, True, True, True, True, True)
    m.add_node_load('N' + str(n), 'FX', 1.0, case='H')
    m.add_node_load('N' + str(n), 'FY', -100.0, case='P')
    for i, p in enumerate((0.0, 1.0, 1.5, 2.0)):
        m.add_load_combo('C' + str(i + 1), {'H': 1.0, 'P': p})
    return m


def pdelta_near_buckling(FE):
    """The cantilever of `pdelta_column` loaded to 0.6 ... 0.95 of its in-plane Euler load (P_cr = pi^2 E Iz / (4 L^2)
    = 633 kip for bending about z): Ke + Kg is close to singular in plane and already indefinite out of plane
    (weak-axis P_cr = 229 kip), which is where an unpivoted factorisation of Ke + Kg must not be trusted."""
    m = FE()
    m.add_material('Steel', 29000.0, 11200.0, 0.3, 2.836e-4)
    m.add_section('W14', 26.5, 362.0, 999.0, 5.45)
    n = 7
    for i in range(n):
        m.add_node('N' + str(i + 1), 0, i*28.0*12/(n - 1), 0)
    m.add_member('M1', 'N1', 'N' + str(n), 'Steel', 'W14')
    m.def_support('N1', True, True, True, True, True, True)
    m.add_node_load('N' + str(n), 'FX', 1.0, case='H')
    m.add_node_load('N' + str(n), 'FZ', 0.05, case='H')
    m.add_node_load('N' + str(n), 'FY', -100.0, case='P')
    for i, p in enumerate((1.2, 3.8, 5.0, 5.7, 6.0)):
        m.add_load_combo('C' + str(i + 1), {'H': 1.0, 'P': p})
    return m


def braced_tc(FE):
    """Braced bay with tension-only diagonals, a compression-only node support spring and a tension-only
    spring element: exercises `_check_TC_convergence` (`Analysis.py:917-1056`)."""
    m = FE()
    m.add_material('Steel', 29000.0, 11200.0, 0.3, 2.836e-4)
    m.add_section('Col', 20.0, 300.0, 300.0, 10.0)
    m.add_section('Rod', 1.0, 0.1, 0.1, 0.05)
    m.add_node('A', 0, 0, 0)
    m.add_node('B', 0, 120, 0)
    m.add_node('C', 144, 120, 0)
    m.add_node('D', 144, 0, 0)
    m.add_node('E', 288, 0, 0)
    m.add_node('F', 288, 20, 0)
    m.def_support('A', True, True, True, True, True, False)
    m.def_support('D', True, False, True, True, True, False)
    m.def_support('E', True, True, True, True, True, True)
    m.def_support('F', True, True, True, True, True, True)
    m.def_support_spring('D', 'DY', 400.0, '-')
    m.add_member('AB', 'A', 'B', 'Steel', 'Col')
    m.add_member('BC', 'B', 'C', 'Steel', 'Col')
    m.add_member('DC', 'D', 'C', 'Steel', 'Col')
    m.add_member('AC', 'A', 'C', 'Steel', 'Rod', tension_only=True)
    m.add_member('DB', 'D', 'B', 'Steel', 'Rod', tension_only=True)
    for d in ('AC', 'DB'):
        m.def_releases(d, Ryi=True, Rzi=True, Ryj=True, Rzj=True)
    m.add_spring('T', 'D', 'E', 30.0, tension_only=True)
    m.add_node_load('B', 'FX', 10.0, case='W')
    m.add_node_load('C', 'FY', -25.0, case='D')
    m.add_node_load('B', 'FY', -25.0, case='D')
    m.add_load_combo('D+W', {'D': 1.0, 'W': 1.0})
    m.add_load_combo('D-W', {'D': 1.0, 'W': -1.0})
    m.add_load_combo('0.6D+W', {'D': 0.6, 'W': 1.5})
    return m


def mat_foundation(FE, nx=6, ny=5, a=12.0, t=24.0, ks=0.1):
    """Quad mesh in the XZ plane on Winkler springs (BASELINE config 4 shape, `MatFoundation.py:121-136`
    with linear springs): every node gets `def_support_spring('DY', ks*tributary)`."""
    m = FE()
    m.add_material('Conc', 3605.0, 1540.6, 0.17, 8.68e-5)
    N = _grid_nodes(m, nx, ny, a, a, 'XZ')
    k = 1
    for r in range(ny):
        for c in range(nx):
            # CCW seen from +Y would give a -Y normal in XZ; keep i-j-m-n as the mesher emits them
            name = m.add_quad('Q' + str(k), N[(c, r)], N[(c + 1, r)], N[(c + 1, r + 1)], N[(c, r + 1)], t, 'Conc')
            m.add_quad_surface_pressure(name, 0.002, case='D')
            k += 1
    for r in range(ny + 1):
        for c in range(nx + 1):
            trib = a*a*(0.5 if r in (0, ny) else 1.0)*(0.5 if c in (0, nx) else 1.0)
            m.def_support_spring(N[(c, r)], 'DY', ks*trib)
    m.def_support(N[(0, 0)], True, False, True, False, False, False)
    m.def_support(N[(nx, 0)], False, False, True, False, False, False)
    m.add_node_load(N[(nx//2, ny//2)], 'FY', -150.0, case='C')
    m.add_node_load(N[(1, 1)], 'FY', -80.0, case='C')
    m.add_load_combo('D', {'D': 1.0})
    m.add_load_combo('D+C', {'D': 1.2, 'C': 1.6})
    return m


def mat_tc(FE, nx=6, ny=5, a=12.0, t=10.0, ks=0.1):
    """Mat foundation on COMPRESSION-ONLY Winkler springs (`MatFoundation.py:136` uses direction '-'), loaded so that
    one side lifts off: the tension/compression-only iteration has to switch a band of support springs off, and a second
    combination switches some of them back on."""
    m = FE()
    m.add_material('Conc', 3605.0, 1540.6, 0.17, 8.68e-5)
    N = _grid_nodes(m, nx, ny, a, a, 'XZ')
    k = 1
    for r in range(ny):
        for c in range(nx):
            name = m.add_quad('Q' + str(k), N[(c, r)], N[(c + 1, r)], N[(c + 1, r + 1)], N[(c, r + 1)], t, 'Conc')
            m.add_quad_surface_pressure(name, 0.0004, case='D')
            k += 1
    for r in range(ny + 1):
        for c in range(nx + 1):
            trib = a*a*(0.5 if r in (0, ny) else 1.0)*(0.5 if c in (0, nx) else 1.0)
            m.def_support_spring(N[(c, r)], 'DY', ks*trib, '-')
    m.def_support(N[(0, 0)], True, False, True, False, False, False)
    m.def_support(N[(nx, 0)], False, False, True, False, False, False)
    for r in range(ny + 1):
        m.add_node_load(N[(0, r)], 'FY', -6.0, case='C')         # heavy wall along one edge
        m.add_node_load(N[(nx, r)], 'FY', 1.0, case='U')         # uplift along the opposite edge
    m.add_node_load(N[(nx//2, ny//2)], 'FY', -30.0, case='C')
    m.add_load_combo('D+C', {'D': 1.0, 'C': 1.0})
    m.add_load_combo('D+C+U', {'D': 1.0, 'C': 1.0, 'U': 1.0})
    m.add_load_combo('D', {'D': 1.4})
    return m


def mixed_building(FE):
    """One-storey frame carrying a quad slab and a plate wall: all four element kinds in one matrix."""
    m = FE()
    m.add_material('Steel', 29000.0, 11200.0, 0.3, 2.836e-4)
    m.add_material('Conc', 3605.0, 1540.6, 0.17, 8.68e-5)
    m.add_section('Col', 20.0, 300.0, 300.0, 10.0)
    nx = ny = 2
    a = 60.0
    N = {}
    k = 1
    for r in range(ny + 1):
        for c in range(nx + 1):
            N[(c, r)] = m.add_node('T' + str(k), c*a, 120.0, r*a)
            k += 1
    B = {}
    for i, (c, r) in enumerate([(0, 0), (nx, 0), (nx, ny), (0, ny)]):
        B[(c, r)] = m.add_node('B' + str(i + 1), c*a, 0.0, r*a)
        m.def_support(B[(c, r)], True, True, True, True, True, True)
        m.add_member('C' + str(i + 1), B[(c, r)], N[(c, r)], 'Steel', 'Col')
    q = 1
    for r in range(ny):
        for c in range(nx):
            nm = m.add_quad('Q' + str(q), N[(c, r)], N[(c + 1, r)], N[(c + 1, r + 1)], N[(c, r + 1)], 6.0, 'Conc')
            m.add_quad_surface_pressure(nm, 0.0008, case='D')
            q += 1
    # wall along the X edge at Z = 0, two rectangular plates between the base line and the slab edge
    W = {}
    for c in range(nx + 1):
        W[c] = B[(c, 0)] if (c, 0) in B else m.add_node('W' + str(c), c*a, 0.0, 0.0)
        if (c, 0) not in B:
            m.def_support(W[c], True, True, True, True, True, True)
    for c in range(nx):
        nm = m.add_plate('P' + str(c + 1), W[c], W[c + 1], N[(c + 1, 0)], N[(c, 0)], 8.0, 'Conc')
        m.add_plate_surface_pressure(nm, 0.0003, case='W')
    m.add_spring('SP1', N[(1, 1)], N[(0, 0)], 5.0)
    m.add_node_load(N[(1, 1)], 'FY', -12.0, case='D')
    m.add_node_load(N[(0, ny)], 'FX', 3.0, case='W')
    m.add_load_combo('D', {'D': 1.4})
    m.add_load_combo('D+W', {'D': 1.2, 'W': 1.0})
    return m


def big_quad_plate(FE, n=500, n_combos=64, a=12.0, t=8.0, seed=20261019):
    """BASELINE config 3: n x n DKMQ quads in the XY plane, perimeter pinned, cases D (uniform
    pressure), L (checkerboard), P1..P6 (100 nodal FZ loads each), `n_combos` random factor rows."""
    m = FE()
    m.add_material('Conc', 3605.0, 1540.6, 0.17, 8.68e-5)
    N = _grid_nodes(m, n, n, a, a, 'XY')
    k = 1
    for r in range(n):
        for c in range(n):
            name = m.add_quad('Q' + str(k), N[(c, r)], N[(c + 1, r)], N[(c + 1, r + 1)], N[(c, r + 1)], t, 'Conc')
            m.add_quad_surface_pressure(name, 0.001, case='D')
            if (r + c) % 2 == 0:
                m.add_quad_surface_pressure(name, 0.0005, case='L')
            k += 1
    for i in range(n + 1):
        for key in ((i, 0), (i, n), (0, i), (n, i)):
            m.def_support(N[key], True, True, True, False, False, False)
    rng = np.random.default_rng(seed)
    n_pt = min(100, max(1, (n - 1)*(n - 1)))
    for p in range(6):
        for _ in range(n_pt):
            c, r = int(rng.integers(1, n)), int(rng.integers(1, n))
            m.add_node_load(N[(c, r)], 'FZ', float(rng.uniform(-10, 0)), case='P' + str(p + 1))
    fac = np.round(rng.uniform(0.5, 1.6, size=(n_combos, 8)), 2)
    cases = ['D', 'L'] + ['P' + str(p + 1) for p in range(6)]
    for c in range(n_combos):
        m.add_load_combo('C' + str(c + 1), {cs: float(fac[c, j]) for j, cs in enumerate(cases)})
    return m


def tc_load_steps(FE):
    """Propped cantilever whose prop is a compression-only support spring that engages only once the tip has moved
    by more than the spring tolerance, plus a tension-only tie: with `analyze(num_steps=4, spring_tolerance=0.05)`
    the spring flips in the middle of the load stepping (`Analysis._first_order` :262-305 undo / redo branches)."""
    m = FE()
    m.add_material('Steel', 29000.0, 11200.0, 0.3, 2.836e-4)
    m.add_section('Beam', 10.0, 100.0, 150.0, 5.0)
    m.add_node('A', 0, 0, 0)
    m.add_node('B', 120, 0, 0)
    m.add_node('C', 240, 0, 0)
    m.add_node('T', 240, 60, 0)
    m.def_support('A', True, True, True, True, True, True)
    m.def_support('T', True, True, True, True, True, True)
    m.add_member('AB', 'A', 'B', 'Steel', 'Beam')
    m.add_member('BC', 'B', 'C', 'Steel', 'Beam')
    m.def_support_spring('C', 'DY', 1.0, '-')
    m.def_support_spring('B', 'DY', 8.0, '+')
    m.add_spring('TIE', 'C', 'T', 4.0, tension_only=True)
    m.add_node_load('C', 'FY', -6.0, case='G')
    m.add_node_load('B', 'FY', -3.0, case='G')
    m.add_member_dist_load('AB', 'Fy', -0.01, -0.01, case='G')
    m.add_node_load('C', 'FY', 5.0, case='U')
    m.add_load_combo('G', {'G': 1.0})
    m.add_load_combo('G+U', {'G': 1.0, 'U': 1.6})
    m.add_load_combo('0.5G', {'G': 0.5})
    return m


# ---------------------------------------------------------------------------------------------------------------------
# `FEModel3D.add_rectangle_mesh` (`Mesh.py:719-1096`): every case returns a model whose meshes have been generated.
# tests/golden/make_mesh_golden.py runs them with the reference and stores names, coordinates and connectivity.
# ---------------------------------------------------------------------------------------------------------------------
def _mesh_model(FE):
    m = FE()
    m.add_material('Conc', 3605.0, 1540.6, 0.17, 8.68e-5)
    return m


def rect_mesh_basic(FE):
    m = _mesh_model(FE)
    m.add_rectangle_mesh('M1', 1.0, 10.0, 6.0, 8.0, 'Conc')
    m.meshes['M1'].generate()
    return m


def rect_mesh_uneven_yz_rect(FE):
    """mesh size that does not divide the sides, YZ plane, shifted origin, rectangular plate elements"""
    m = _mesh_model(FE)
    m.add_rectangle_mesh('W', 1.0, 7.3, 4.1, 6.0, 'Conc', origin=[2.5, -1.25, 40.0], plane='YZ', element_type='Rect')
    m.meshes['W'].generate()
    return m


def rect_mesh_controls_xz(FE):
    """control points (one of them a near duplicate of an edge, one outside the mesh) in the XZ plane"""
    m = _mesh_model(FE)
    m.add_rectangle_mesh('S', 0.75, 6.0, 5.0, 7.0, 'Conc', kx_mod=0.8, ky_mod=0.9, origin=[1.0, 2.0, 3.0], plane='XZ',
                         x_control=[2.5, 6.0 + 1e-12, 7.5], y_control=[1.7, 3.3, 1.7])
    m.meshes['S'].generate()
    return m


def rect_mesh_opening(FE):
    m = _mesh_model(FE)
    m.add_rectangle_mesh('M', 1.0, 12.0, 9.0, 8.0, 'Conc')
    m.meshes['M'].add_rect_opening('door', 3.0, 0.0, 3.0, 4.5)
    m.meshes['M'].add_rect_opening('window', 7.5, 3.2, 2.2, 2.6)
    m.meshes['M'].generate()
    return m


def rect_mesh_renamed(FE):
    """two meshes added before either is generated (both start at N1 / Q1), user nodes already in the model, and a
    regeneration after an opening was added: `_rename_duplicates`, `unique_name`, `_remove_from_model`"""
    m = _mesh_model(FE)
    m.add_node('N2', -5.0, 0.0, 0.0)
    m.add_node('A', -6.0, 0.0, 0.0)
    m.add_rectangle_mesh('M1', 2.0, 6.0, 4.0, 8.0, 'Conc')
    m.add_rectangle_mesh('M2', 2.0, 4.0, 4.0, 8.0, 'Conc', origin=[6.0, 0.0, 0.0])
    m.meshes['M1'].generate()
    m.meshes['M2'].generate()
    m.meshes['M1'].add_rect_opening('o', 2.0, 0.0, 2.0, 2.0)
    m.meshes['M1'].generate()
    return m


def mat_mesh_generated(FE):
    """`add_mat_foundation` + point loads (control points) + an opening, generated explicitly"""
    m = _mesh_model(FE)
    m.add_mat_foundation('MAT1', 2.0, 14.0, 9.0, 18.0, 'Conc', 0.15, origin=[1.0, 0.0, -2.0], x_control=[], y_control=[])
    mat = m.mats['MAT1']
    mat.add_mat_pt_load([4.5, 1.5], 'FY', -120.0, case='D')
    mat.add_mat_pt_load([11.0, 5.25], 'FY', -80.0, case='L')
    mat.add_mat_pt_load([11.0, 5.25], 'MX', 30.0, case='L')
    mat.add_rect_opening('pit', 7.0, 2.0, 9.5, 4.0)
    mat.generate()
    return m


def rect_mesh_merged(FE):
    """two meshes sharing an edge (and a wall standing on the slab), supports on one of the duplicates, merged"""
    m = _mesh_model(FE)
    m.add_rectangle_mesh('A', 2.0, 6.0, 4.0, 8.0, 'Conc')
    m.meshes['A'].generate()
    m.add_rectangle_mesh('B', 2.0, 4.0, 4.0, 8.0, 'Conc', origin=[6.0, 0.0, 0.0])
    m.meshes['B'].generate()
    m.add_rectangle_mesh('W', 2.0, 4.0, 6.0, 6.0, 'Conc', origin=[6.0, 0.0, 0.0], plane='YZ')
    m.meshes['W'].generate()
    first_b = list(m.meshes['B'].nodes.values())[0]
    m.def_support(first_b.name, True, True, True, False, False, False)
    m.def_support_spring(list(m.meshes['W'].nodes.values())[1].name, 'DY', 55.0, '-')
    m._test_removed = m.merge_duplicate_nodes()
    return m


def annulus_mesh_case(FE):
    """annulus about Y with two transition rings, then a second annulus about Z that continues the numbering"""
    m = _mesh_model(FE)
    m.add_annulus_mesh('A1', 1.0, 10.0, 1.0, 8.0, 'Conc', origin=[3.0, 1.0, -2.0])
    m.meshes['A1'].generate()
    m.add_annulus_mesh('A2', 2.0, 9.0, 4.0, 6.0, 'Conc', kx_mod=0.7, ky_mod=0.8, axis='Z')
    m.meshes['A2'].generate()
    return m


def frustrum_mesh_case(FE):
    m = _mesh_model(FE)
    m.add_frustrum_mesh('F', 1.5, 8.0, 3.0, 5.0, 7.0, 'Conc', origin=[0.0, 2.0, 0.0], axis='X')
    m.meshes['F'].generate()
    return m


def cylinder_mesh_case(FE):
    """cylinders are generated on construction; the second one uses rectangles, a given element count and the Z axis"""
    m = _mesh_model(FE)
    m.add_cylinder_mesh('C1', 2.0, 5.0, 9.0, 8.0, 'Conc', origin=[4.0, 1.0, 7.0])
    m.add_cylinder_mesh('C2', 1.5, 3.0, 4.0, 6.0, 'Conc', axis='Z', num_elements=10, element_type='Rect')
    return m


def mesh_edited_case(FE):
    """two meshes sharing an edge, one deleted again (shared nodes stay), a free-standing node, `orphaned_nodes`, then
    `rename`: the second mesh's names start at N100 / Q100, so every entry is re-keyed, and the loose node is named like
    a name `rename` hands out earlier (the reference then replaces an entry -- reproduced, see `FEModel3D.rename`)"""
    m = _mesh_model(FE)
    m.add_rectangle_mesh('A', 1.0, 3.0, 2.0, 8.0, 'Conc')
    m.meshes['A'].generate()
    m.add_rectangle_mesh('B', 1.0, 2.0, 2.0, 8.0, 'Conc', origin=[3.0, 0.0, 0.0], start_node='N100', start_element='Q100')
    m.meshes['B'].generate()
    m.merge_duplicate_nodes()
    m.add_node('N3000', 50.0, 0.0, 0.0)
    m.add_steel_section('W', 10.0, 20.0, 100.0, 1.5, 12.0, 30.0, 'Conc')
    m.add_node('LOOSE', 70.0, 1.0, 0.0)
    m.add_member('BM', 'N3000', 'LOOSE', 'Conc', 'W')
    m.add_spring('SP', 'N1', 'N3000', 5.0)
    m.delete_mesh('A')
    m.add_node('ALONE', 80.0, 0.0, 0.0)
    m.add_node('N2', 90.0, 0.0, 0.0)      # mesh A's N2 is gone; `rename` hands 'N2' out before it reaches this entry
    m._test_removed = m.orphaned_nodes()
    m.rename()
    return m


def shear_wall_case(FE):
    """the two walls of the reference's `Testing/test_shear_wall.py:6-60` (openings, three flanges each, a support line, a
    story with a shear), one in the XY plane and one in YZ, on a 2 ft mesh; plus a second material region, an axial load and
    a partial-length story on the first wall"""
    m = FE()
    Em = 900*2000/1000*144
    m.add_material('CMU', Em, 0.4*Em, 0.17, 0.140)
    m.add_material('Grouted', 1.3*Em, 0.52*Em, 0.2, 0.150)
    for wall_name, plane, origin in (('W1', 'XY', [0, 0, 0]), ('W2', 'YZ', [40, 0, 0])):
        m.add_shear_wall(wall_name, mesh_size=2, length=26, height=16, thickness=0.667, material_name='CMU', ky_mod=1,
                         plane=plane, origin=origin)
        w = m.shear_walls[wall_name]
        w.add_opening('Door 1', x_start=2, y_start=0, width=4, height=12, tie=None)
        w.add_opening('Window 1', x_start=8, y_start=8, width=4, height=4, tie=None)
        w.add_opening('Window 2', x_start=14, y_start=8, width=4, height=4, tie=None)
        w.add_opening('Door 2', x_start=20, y_start=0, width=4, height=12, tie=None)
        w.add_support(elevation=0, x_start=0, x_end=26)
        w.add_story('Roof', elevation=16, x_start=0, x_end=26)
        w.add_shear(story_name='Roof', force=100, case='E')
        w.add_flange(thickness=0.667, width=0.75*16, x=0, y_start=0, y_end=16, material='CMU', side='-z')
        w.add_flange(thickness=0.667, width=0.75*16, x=7, y_start=0, y_end=16, material='CMU', side='+z')
        w.add_flange(thickness=0.667, width=0.75*16, x=26, y_start=0, y_end=16, material='CMU', side='-z')
    w1 = m.shear_walls['W1']
    w1.asign_material('Grouted', 1.0, x_start=6, x_end=20, y_start=0, y_end=8)
    w1.add_story('Mezzanine', elevation=8, x_start=6, x_end=20)
    w1.add_axial('Roof', 30.0, case='D')
    m.add_load_combo('1.0E', {'E': 1.0}, combo_tags='strength')
    for wall in m.shear_walls.values():
        wall.generate()
    return m


def shear_wall_regenerated(FE):
    """`Testing/test_shear_wall.py:364-440`: a plain wall, generated, then two openings added and generated again"""
    m = FE()
    Em = 900*2000/1000*144
    m.add_material('CMU', Em, 0.4*Em, 0.17, 0.140)
    m.add_shear_wall('Wall', mesh_size=2, length=24, height=16, thickness=0.667, material_name='CMU', ky_mod=1, plane='XY',
                     origin=[0, 0, 0])
    w = m.shear_walls['Wall']
    w.add_support(elevation=0, x_start=0, x_end=24)
    w.add_story('Roof', elevation=16, x_start=0, x_end=24)
    w.add_shear(story_name='Roof', force=100, case='E')
    w.generate()
    m.add_load_combo('1.0E', {'E': 1.0}, combo_tags='strength')
    w.add_opening('Door', x_start=6, y_start=0, width=6, height=12, tie=None)
    w.add_opening('Window', x_start=16, y_start=8, width=6, height=4, tie=None)
    assert w.needs_update
    w.generate()
    return m


RECT_MESH_CASES = {
    'basic': rect_mesh_basic,
    'uneven_yz_rect': rect_mesh_uneven_yz_rect,
    'controls_xz': rect_mesh_controls_xz,
    'opening': rect_mesh_opening,
    'renamed': rect_mesh_renamed,
    'mat': mat_mesh_generated,
    'merged': rect_mesh_merged,
    'annulus': annulus_mesh_case,
    'frustrum': frustrum_mesh_case,
    'cylinder': cylinder_mesh_case,
    'edited': mesh_edited_case,
    'shear_wall': shear_wall_case,
    'shear_wall_regenerated': shear_wall_regenerated,
}


def rect_mesh_slab(FE):
    """A slab with an opening built by `add_rectangle_mesh` (generated by the analysis itself), edge nodes pinned,
    pressures on every element, three combos: the mesh generator inside the full path."""
    m = _mesh_model(FE)
    m.add_rectangle_mesh('SLAB', 12.0, 96.0, 72.0, 8.0, 'Conc', x_control=[30.0], y_control=[20.0, 50.0])
    mesh = m.meshes['SLAB']
    mesh.add_rect_opening('shaft', 30.0, 20.0, 24.0, 30.0)
    mesh.generate()
    for node in mesh.nodes.values():
        x, y = mesh.node_local_coords(node)
        if x in (0.0, 96.0) or y in (0.0, 72.0):
            m.def_support(node.name, True, True, True, False, False, False)
    for k, name in enumerate(mesh.elements):
        m.add_quad_surface_pressure(name, 0.002, case='D')
        if k % 3 == 0:
            m.add_quad_surface_pressure(name, 0.001, case='L')
    m.add_load_combo('C1', {'D': 1.4})
    m.add_load_combo('C2', {'D': 1.2, 'L': 1.6})
    m.add_load_combo('C3', {'D': 0.9})
    return m


def mat_mesh_tc(FE):
    """A mat built by `add_mat_foundation` (generated by the analysis), one heavy eccentric column load: part of the mat
    lifts off its compression-only springs (`MatFoundation.py:136`), so `analyze` iterates."""
    m = _mesh_model(FE)
    m.add_mat_foundation('MAT', 30.0, 240.0, 180.0, 24.0, 'Conc', 0.1, x_control=[], y_control=[])
    mat = m.mats['MAT']
    mat.add_mat_pt_load([40.0, 45.0], 'FY', -400.0, case='D')
    mat.add_mat_pt_load([40.0, 45.0], 'MZ', 9000.0, case='W')
    mat.add_mat_pt_load([200.0, 135.0], 'FY', -150.0, case='D')
    mat.generate()
    # in-plane restraint (the springs act in DY only)
    first, last = list(mat.nodes.values())[0], mat.last_node
    m.def_support(first.name, True, False, True, False, True, False)
    m.def_support(last.name, False, False, True, False, False, False)
    m.add_load_combo('G', {'D': 1.0})
    m.add_load_combo('GW', {'D': 0.9, 'W': 1.0})
    return m


# points (global X, Z) at which `MatFoundation.soil_pressure` is compared with the reference
MAT_PRESSURE_POINTS = [(40.0, 45.0), (0.0, 0.0), (123.4, 56.7), (240.0, 180.0), (200.0, 135.0), (15.0, 170.0), (300.0, 10.0)]


def cylinder_tank(FE):
    """A tank wall (`add_cylinder_mesh`) with a frustrum roof (`add_frustrum_mesh`, merged onto the wall's top ring): quads
    in general position, base fixed, pressure on the wall and on the roof, two combos."""
    m = _mesh_model(FE)
    m.add_cylinder_mesh('WALL', 30.0, 60.0, 72.0, 10.0, 'Conc', num_elements=12)
    m.add_frustrum_mesh('ROOF', 30.0, 60.0, 15.0, -18.0, 8.0, 'Conc', origin=[0.0, 72.0, 0.0])
    m.meshes['ROOF'].generate()
    m.merge_duplicate_nodes(tolerance=1e-6)
    for node in m.nodes.values():
        if node.Y == 0.0:
            m.def_support(node.name, True, True, True, True, True, True)
    for name in m.meshes['WALL'].elements:
        m.add_quad_surface_pressure(name, 0.004, case='H')
    for k, name in enumerate(m.meshes['ROOF'].elements):
        m.add_quad_surface_pressure(name, -0.001 if k % 2 else -0.0015, case='S')
    m.add_load_combo('C1', {'H': 1.0})
    m.add_load_combo('C2', {'H': 1.2, 'S': 1.6})
    return m


SMALL_MODELS = {
    'space_frame': (space_frame, 'tc'),
    'portal_frame': (portal_frame, 'linear'),
    'continuous_beam': (continuous_beam, 'linear'),
    'quad_mesh_xy': (quad_mesh, 'linear'),
    'quad_mesh_yz_ortho': (lambda FE: quad_mesh(FE, nx=4, ny=3, a=9.0, b=11.0, t=5.0, plane='YZ', kx_mod=0.5, ky_mod=0.9, seed=3), 'linear'),
    'plate_mesh_xz': (plate_mesh, 'linear'),
    'skew_shell': (skew_shell, 'linear'),
    'moment_frame': (moment_frame, 'linear'),
    'mat_foundation': (mat_foundation, 'linear'),
    'mixed_building': (mixed_building, 'linear'),
    'pdelta_column': (pdelta_column, 'pdelta'),
    'pdelta_frame': (lambda FE: moment_frame(FE, bays_x=2, bays_z=1, stories=4, interior=1, n_combos=4), 'pdelta'),
    'braced_tc': (braced_tc, 'tc'),
    'pdelta_near_buckling': (pdelta_near_buckling, 'pdelta'),
    'tc_load_steps': (tc_load_steps, 'tc_steps'),
    'mat_tc': (mat_tc, 'tc'),
    'rect_mesh_slab': (rect_mesh_slab, 'linear'),
    'mat_mesh_tc': (mat_mesh_tc, 'tc'),
    'cylinder_tank': (cylinder_tank, 'linear'),
}

# keyword arguments of `analyze` for the load-stepping fixture (mode 'tc_steps')
TC_STEPS_KW = dict(num_steps=4, spring_tolerance=0.2, member_tolerance=0.0)



# ---------------------------------------------------------------------------------------------
# modal analysis (`FEModel3D.analyze_modal`, `FEModel3D.py:2469-2568`): builder -> keyword arguments of the call
# ---------------------------------------------------------------------------------------------
def modal_cantilever(FE):
    """ten members along X (the shape of the reference's `test_modal_analysis.py` cantilever, with Iy != Iz so that the
    two bending families are not degenerate): consistent mass from self weight plus nodal masses in one combination"""
    m = FE()
    L, A, rho = 5.0, 0.01, 7800.0
    for n in range(11):
        m.add_node(f'N{n + 1}', n*L/10, 0.0, 0.0)
    m.def_support('N1', True, True, True, True, True, True)
    m.add_material('Steel', 200e9, 80e9, 0.3, rho)
    m.add_section('BeamSection', A, 8.33e-6, 5.1e-6, 1.67e-6)
    for n in range(10):
        m.add_member(f'M{n + 1}', f'N{n + 1}', f'N{n + 2}', 'Steel', 'BeamSection')
    m.add_member_self_weight('FY', -1.0, 'D1')
    for n in range(11):
        m.add_node_load(f'N{n + 1}', 'FY', -A*rho*L/(20 if n in (0, 10) else 10), 'D2')
    m.add_load_combo('Consistent Mass Combo', {'D1': 1.0})
    m.add_load_combo('Both', {'D1': 1.0, 'D2': 0.5})
    return m


def modal_frame(FE):
    """two-bay space frame in general position: a physical member split at an interior node, an inclined and a rolled
    member, an end release, non-self-weight distributed and point loads in the mass direction, nodal masses, a spring, an
    enforced support displacement (must come back in the mode shapes' known DOFs) and a support spring"""
    m = FE()
    m.add_material('Steel', 29000.0, 11200.0, 0.3, 0.000283)
    m.add_section('Col', 20.0, 150.0, 400.0, 3.0)
    m.add_section('Bm', 12.0, 60.0, 300.0, 1.5)
    pts = {'A0': (0, 0, 0), 'B0': (240, 0, 0), 'C0': (240, 0, 200), 'D0': (0, 0, 200),
           'A1': (0, 144, 0), 'B1': (240, 144, 0), 'C1': (250, 150, 210), 'D1': (0, 144, 200), 'AM': (0, 72, 0), 'E1': (120, 190, 100)}
    for name, (x, y, z) in pts.items():
        m.add_node(name, float(x), float(y), float(z))
    for base in ('A0', 'B0', 'C0'):
        m.def_support(base, True, True, True, True, True, True)
    m.def_support('D0', True, True, True, False, False, False)
    m.def_node_disp('D0', 'DY', -0.25)
    m.def_support_spring('E1', 'DY', 40.0)
    m.add_member('CA', 'A0', 'A1', 'Steel', 'Col')           # split at AM
    m.add_member('CB', 'B0', 'B1', 'Steel', 'Col', rotation=30.0)
    m.add_member('CC', 'C0', 'C1', 'Steel', 'Col')           # inclined
    m.add_member('CD', 'D0', 'D1', 'Steel', 'Col')
    m.add_member('B1', 'A1', 'B1', 'Steel', 'Bm')
    m.add_member('B2', 'B1', 'C1', 'Steel', 'Bm')
    m.add_member('B3', 'C1', 'D1', 'Steel', 'Bm')
    m.add_member('B4', 'D1', 'A1', 'Steel', 'Bm')
    m.add_member('R1', 'A1', 'E1', 'Steel', 'Bm')
    m.add_member('R2', 'C1', 'E1', 'Steel', 'Bm', rotation=-15.0)
    m.def_releases('B3', Rzj=True, Ryj=True)
    m.add_spring('BR', 'AM', 'B1', 35.0)
    m.add_member_self_weight('FY', -1.0, 'SW')
    m.add_member_dist_load('B1', 'FY', -0.05, -0.08, 20.0, 200.0, case='SD')
    m.add_member_dist_load('B2', 'FY', -0.04, -0.04, case='SD')
    m.add_member_dist_load('B2', 'FX', 0.03, 0.03, case='SD')        # not in the mass direction
    m.add_member_pt_load('B4', 'FY', -3.0, 80.0, case='SD')
    m.add_member_pt_load('CA', 'FY', -2.0, 100.0, case='SD')
    m.add_node_load('E1', 'FY', -6.0, 'SD')
    m.add_node_load('B1', 'FY', -1.5, 'L')
    m.add_node_load('B1', 'FX', 4.0, 'L')
    m.add_load_combo('Mass', {'SW': 1.0, 'SD': 1.0, 'L': 0.25})
    m.add_load_combo('Mass nodal', {'SW': 1.0, 'L': 0.25})
    return m


MODAL_MODELS = {
    'modal_cantilever_consistent': (modal_cantilever, dict(num_modes=3, mass_combo_name='Consistent Mass Combo',
                                                           mass_direction='Y', gravity=9.81)),
    'modal_cantilever_both': (modal_cantilever, dict(num_modes=6, mass_combo_name='Both', mass_direction='Y', gravity=9.81)),
    'modal_frame': (modal_frame, dict(num_modes=9, mass_combo_name='Mass nodal', mass_direction='Y', gravity=386.4)),
    # member loads as masses: the reference's `lumped_m` end shares `(1 - x)/L`, `x/L` make this mass matrix indefinite (its
    # own `analyze_modal` returns NaN frequencies), so only the matrices are compared
    'modal_frame_member_loads': (modal_frame, dict(num_modes=0, mass_combo_name='Mass', mass_direction='Y', gravity=386.4)),
}
