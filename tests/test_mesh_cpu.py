"""`FEModel3D.add_rectangle_mesh` / `RectangleMesh.generate` (pynite_b200/Mesh.py) against what the reference's generator
produced for the same calls (tests/golden/rect_meshes.npz, written by tests/golden/make_mesh_golden.py running
`/root/reference/Pynite/Mesh.py`): names, dictionary order, bit-identical coordinates, connectivity, mesh membership."""
import os

import numpy as np
import pytest

from pynite_b200 import FEModel3D
from tests import models
from tests.util import mesh_record

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'rect_meshes.npz'))


@pytest.mark.parametrize('case', list(models.RECT_MESH_CASES))
def test_rectangle_mesh_matches_reference(case):
    rec = mesh_record(models.RECT_MESH_CASES[case](FEModel3D))
    keys = [k[len(case) + 1:] for k in GOLD.files if k.startswith(case + '/')]
    assert sorted(keys) == sorted(rec.keys())
    for key in keys:
        ref, ours = GOLD[case + '/' + key], rec[key]
        assert ref.shape == ours.shape, (key, ref.shape, ours.shape)
        if ref.dtype.kind == 'f':
            assert np.array_equal(ref, ours), key              # bit-identical coordinates and properties
        else:
            assert list(ref.reshape(-1)) == list(ours.reshape(-1)), key


def test_unique_name_rule():
    """`FEModel3D.unique_name` (:2882-2901): count + 1, bumped until unused."""
    m = FEModel3D()
    m.add_node('A', 0, 0, 0)
    m.add_node('N2', 1, 0, 0)
    assert m.unique_name(m.nodes, 'N') == 'N3'
    m.add_node('N3', 2, 0, 0)
    assert m.unique_name(m.nodes, 'N') == 'N4'
    assert m.unique_name(m.quads, 'Q') == 'Q1'


def test_large_mesh_is_linear_time():
    """BASELINE config 3's generator call: 500 x 500 quads.  The reference's orphan-node search makes this quadratic;
    here it has to finish in seconds."""
    import time
    m = FEModel3D()
    m.add_material('Conc', 3605.0, 1540.6, 0.17, 8.68e-5)
    m.add_rectangle_mesh('P', 12.0, 6000.0, 6000.0, 8.0, 'Conc')
    t0 = time.perf_counter()
    m.meshes['P'].generate()
    dt = time.perf_counter() - t0
    assert len(m.nodes) == 501*501 and len(m.quads) == 500*500
    q = m.quads['Q250000']
    assert (q.i_node.name, q.j_node.name, q.m_node.name, q.n_node.name) == ('N250499', 'N250500', 'N251001', 'N251000')
    assert m.nodes['N251001'].X == 6000.0 and m.nodes['N251001'].Y == 6000.0
    assert dt < 60.0, dt


def test_merge_duplicate_nodes_is_linear_time():
    """Two 150 x 150 meshes sharing an edge: the reference compares all 1e9 node pairs; the spatial hash here merges the
    151 duplicates in seconds, and the quads of the second mesh end up on the first mesh's edge nodes."""
    import time
    m = FEModel3D()
    m.add_material('Conc', 3605.0, 1540.6, 0.17, 8.68e-5)
    m.add_rectangle_mesh('A', 1.0, 150.0, 150.0, 8.0, 'Conc')
    m.meshes['A'].generate()
    m.add_rectangle_mesh('B', 1.0, 150.0, 150.0, 8.0, 'Conc', origin=[150.0, 0.0, 0.0])
    m.meshes['B'].generate()
    t0 = time.perf_counter()
    removed = m.merge_duplicate_nodes()
    dt = time.perf_counter() - t0
    assert len(removed) == 151 and len(m.nodes) == 2*151*151 - 151
    first_b = list(m.meshes['B'].elements.values())[0]
    assert first_b.i_node is m.meshes['A'].nodes['N151'] and first_b.i_node.X == 150.0
    assert dt < 30.0, dt
