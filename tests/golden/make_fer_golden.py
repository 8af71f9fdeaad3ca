"""Generates tests/golden/element_fer.npz by RUNNING THE REFERENCE on member / shell fixtures of tests/models.py: every
sub-member's `_fer_unc`, `fer`, `FER` and every quad's / plate's `fer`, `FER` for every combination
(`/root/reference/Pynite/Member3D.py:742-850, 1080-1091`, `Quad3D.py:721-788, 870-882`, `Plate3D.py:324-393, 510-522`).
Development container only.

    python tests/golden/make_fer_golden.py
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
from tests.util import FER_FIXTURES, fer_record   # noqa: E402


def main():
    data = {}
    for name in FER_FIXTURES:
        build, mode = models.SMALL_MODELS[name]
        ref = make_golden.reference_run(build, mode)
        for key, val in fer_record(ref).items():
            data[name + '/' + key] = val
        print(name, {k: v.shape for k, v in fer_record(ref).items()})
    path = os.path.join(HERE, 'element_fer.npz')
    np.savez_compressed(path, **data)
    print('->', path, f'{os.path.getsize(path)/1024:.1f} KiB')


if __name__ == '__main__':
    main()
