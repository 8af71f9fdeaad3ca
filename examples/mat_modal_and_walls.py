"""A tour of the builders that sit on top of the GPU path: a mat foundation with compression-only soil springs
(`analyze`, lift-off iteration on the device), a shear wall with openings (`analyze_linear`, pier forces, stiffness) and
the natural frequencies of a cantilever (`analyze_modal`, block Lanczos with the solves on the GPU).  Needs a B200.

    python examples/mat_modal_and_walls.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pynite_b200 import FEModel3D          # noqa: E402

# ---- mat foundation on Winkler springs --------------------------------------------------------------------------------
mat_model = FEModel3D()
mat_model.add_material('Conc', 3605.0, 1540.6, 0.17, 8.68e-5)
mat_model.add_mat_foundation('MAT', 12.0, 240.0, 180.0, 24.0, 'Conc', ks=0.1)
mat_model.mats['MAT'].add_mat_pt_load([60.0, 45.0], 'FY', -150.0, case='D')
mat_model.mats['MAT'].add_mat_pt_load([180.0, 135.0], 'FY', -90.0, case='D')
mat_model.add_load_combo('1.0D', {'D': 1.0})
mat_model.analyze()
print('largest soil pressure under the mat:', max(mat_model.mats['MAT'].soil_pressure(x, z, '1.0D')
                                                   for x in (0.0, 60.0, 120.0, 240.0) for z in (0.0, 45.0, 180.0)))
print('largest mat moment Mx:', mat_model.mats['MAT'].max_moment('Mx', '1.0D'))

# ---- shear wall with a door and a window ------------------------------------------------------------------------------
wall_model = FEModel3D()
wall_model.add_material('CMU', 259200.0, 103680.0, 0.17, 0.140)
wall_model.add_shear_wall('W1', mesh_size=1, length=24, height=16, thickness=0.667, material_name='CMU', ky_mod=1)
wall = wall_model.shear_walls['W1']
wall.add_opening('Door', 6, 0, 6, 12)
wall.add_opening('Window', 16, 8, 6, 4)
wall.add_support(elevation=0)
wall.add_story('Roof', elevation=16)
wall.add_shear('Roof', 100, case='E')
wall_model.add_load_combo('1.0E', {'E': 1.0}, combo_tags='strength')
wall_model.analyze_linear()
print('wall stiffness at the roof:', wall.stiffness('Roof'))
wall.print_piers('1.0E')

# ---- modal analysis of a cantilever ---------------------------------------------------------------------------------
beam = FEModel3D()
for n in range(11):
    beam.add_node(f'N{n + 1}', 0.5*n, 0.0, 0.0)
beam.def_support('N1', True, True, True, True, True, True)
beam.add_material('Steel', 200e9, 80e9, 0.3, 7800.0)
beam.add_section('Box', 0.01, 8.33e-6, 5.1e-6, 1.67e-6)
for n in range(10):
    beam.add_member(f'M{n + 1}', f'N{n + 1}', f'N{n + 2}', 'Steel', 'Box')
beam.add_member_self_weight('FY', -1.0, 'D')
beam.add_load_combo('Mass', {'D': 1.0})
beam.analyze_modal(num_modes=4, mass_combo_name='Mass', mass_direction='Y', gravity=9.81)
print('natural frequencies (Hz):', beam.frequencies)
