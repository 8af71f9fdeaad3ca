"""Generates tests/golden/member_diagrams.npz by RUNNING THE REFERENCE on the member fixtures of tests/models.py and
querying every physical member's result diagrams (`PhysMember.py:202-1404` -> `Member3D.py:1163-2834`, `BeamSegZ.py`,
`BeamSegY.py`) for every combination: sampled arrays, point values, extremes.  Development container only.

    python tests/golden/make_diagram_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden                                # noqa: E402  (puts the stubs and /root/reference on the path)

from tests import models                          # noqa: E402
from tests.util import DIAGRAM_FIXTURES, DIAGRAM_POINTS, diagram_record   # noqa: E402


def main():
    data = {}
    for name in DIAGRAM_FIXTURES:
        build, mode = models.SMALL_MODELS[name]
        ref = make_golden.reference_run(build, mode)
        rec = diagram_record(ref)
        for key, val in rec.items():
            data[name + '/' + key] = val
        print(f'{name:22s} {mode:8s} members {len(ref.members):3d} combos {len(ref.load_combos):2d} '
              f'values {sum(v.size for v in rec.values() if v.dtype.kind == "f")}')
    path = os.path.join(HERE, 'member_diagrams.npz')
    np.savez_compressed(path, **data)
    print('->', path, f'{os.path.getsize(path)/1024:.1f} KiB', 'points per array:', DIAGRAM_POINTS)


if __name__ == '__main__':
    main()
