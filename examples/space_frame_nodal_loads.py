"""BASELINE config 1: the reference's `Examples/Space Frame - Nodal Loads 1.py` (Logan, example 5.8) through pynite_b200.
Only the import changes; the rendering lines of the original are dropped.  Needs a B200 and the built library
(`python -c "import __graft_entry__ as g; g.build()"`).

    python examples/space_frame_nodal_loads.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pynite_b200 import FEModel3D          # noqa: E402   (the reference: `from Pynite import FEModel3D`)

frame = FEModel3D()
frame.add_node('N1', 0, 0, 0)
frame.add_node('N2', -100, 0, 0)
frame.add_node('N3', 0, 0, -100)
frame.add_node('N4', 0, -100, 0)
for name in ('N2', 'N3', 'N4'):
    frame.def_support(name, True, True, True, True, True, True)
frame.add_material('Steel', 30000, 10000, 0.3, 2.836e-4)
frame.add_section('FrameSection', 10, 100, 100, 50)
frame.add_member('M1', 'N2', 'N1', 'Steel', 'FrameSection')
frame.add_member('M2', 'N3', 'N1', 'Steel', 'FrameSection')
frame.add_member('M3', 'N4', 'N1', 'Steel', 'FrameSection')
frame.add_node_load('N1', 'FY', -50)
frame.add_node_load('N1', 'MX', -1000)
frame.analyze(check_statics=True)

node = frame.nodes['N1']
print('Calculated:', [node.DX['Combo 1'], node.DY['Combo 1'], node.DZ['Combo 1'], node.RX['Combo 1'], node.RY['Combo 1'],
                      node.RZ['Combo 1']])
print('Textbook:  ', [7.098e-5, -0.014, -2.352e-3, -3.996e-3, 1.78e-5, -1.033e-4])
print('M3 axial force along the member:', frame.members['M3'].axial_array(5)[1])
print('M1 extreme moments about z:', frame.members['M1'].max_moment('Mz'), frame.members['M1'].min_moment('Mz'))
