"""Generates tests/golden/rect_meshes.npz by RUNNING THE REFERENCE'S `add_rectangle_mesh` / `RectangleMesh.generate`
(`/root/reference/Pynite/Mesh.py:719-1096`) on the cases of tests/models.py::RECT_MESH_CASES: node names and coordinates in
dictionary order, element names and connectivity, for the model and for every mesh.  Development container only.

    python tests/golden/make_mesh_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, '_stubs'))
sys.path.insert(0, '/root/reference')

from Pynite import FEModel3D as RefModel          # noqa: E402

from tests import models                          # noqa: E402
from tests.util import mesh_record                # noqa: E402


def main():
    data = {}
    for case, build in models.RECT_MESH_CASES.items():
        rec = mesh_record(build(RefModel))
        for key, val in rec.items():
            data[case + '/' + key] = val
        print(f'{case:18s} nodes {len(rec["node_names"]):5d}  quads {len(rec["quad_names"]):5d}  plates {len(rec["plate_names"]):5d}')
    path = os.path.join(HERE, 'rect_meshes.npz')
    np.savez_compressed(path, **data)
    print('->', path, f'{os.path.getsize(path)/1024:.1f} KiB')


if __name__ == '__main__':
    main()
