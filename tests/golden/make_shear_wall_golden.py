"""Generates tests/golden/shear_walls.npz by RUNNING THE REFERENCE on tests/models.py::shear_wall_case: `analyze_linear`, then
the displacements, the quads' local end forces, every pier's and coupling beam's `sum_forces` and the walls' story
stiffness (`/root/reference/Pynite/ShearWall.py:755-794, 944-1052`).  Development container only.

    python tests/golden/make_shear_wall_golden.py
"""
import contextlib
import io
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
from tests.util import SHEAR_WALL_COMBOS, shear_wall_record   # noqa: E402


def main():
    ref = models.shear_wall_case(RefModel)
    with contextlib.redirect_stdout(io.StringIO()):
        ref.analyze_linear()
    data = shear_wall_record(ref)
    data['D'] = np.array([ref._D[c].reshape(-1) for c in SHEAR_WALL_COMBOS])
    data['quad_f'] = np.array([[q.f(c).reshape(-1) for q in ref.quads.values()] for c in SHEAR_WALL_COMBOS])
    path = os.path.join(HERE, 'shear_walls.npz')
    np.savez_compressed(path, **data)
    print({k: v.shape for k, v in data.items()})
    print('->', path, f'{os.path.getsize(path)/1024:.1f} KiB')


if __name__ == '__main__':
    main()
