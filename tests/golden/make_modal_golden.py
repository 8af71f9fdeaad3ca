"""Generates tests/golden/modal.npz by RUNNING THE REFERENCE'S `FEModel3D.M` / `analyze_modal`
(`/root/reference/Pynite/FEModel3D.py:2004-2134, 2469-2568`) on tests/models.py::MODAL_MODELS: the global mass matrix
(dense), every sub-member's local and global mass matrix, the frequencies and the stored mode shapes.  Development
container only.

    python tests/golden/make_modal_golden.py
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
from Pynite import Analysis                       # noqa: E402

from tests import models                          # noqa: E402


def main():
    data = {}
    for name, (build, kw) in models.MODAL_MODELS.items():
        ref = build(RefModel)
        if kw['num_modes']:
            ref.analyze_modal(log=False, **kw)
        else:
            Analysis._prepare_model(ref)
        args = (kw['mass_combo_name'], kw['mass_direction'], kw['gravity'])
        subs = [s for pm in ref.members.values() for s in pm.sub_members.values()]
        if kw['num_modes']:
            data[name + '/frequencies'] = np.asarray(ref.frequencies, dtype=float)
            data[name + '/modes'] = np.hstack([ref._D[f'Mode {i + 1}'] for i in range(kw['num_modes'])])
        data[name + '/M'] = np.asarray(ref.M(*args, log=False, sparse=True).todense())
        data[name + '/member_m'] = np.array([s.m(*args) for s in subs])
        data[name + '/member_M'] = np.array([s.M(*args) for s in subs])
        data[name + '/node_M'] = np.array([n.M(*args, characteristic_length=2.5) for n in ref.nodes.values()])
        data[name + '/char_length'] = np.array(ref._calculate_characteristic_length())
        data[name + '/combos'] = np.array(list(ref.load_combos.keys()))
        print(f'{name:30s} dof {6*len(ref.nodes):4d}  f = {np.array2string(np.asarray(getattr(ref, "frequencies", [])), precision=4)}')
    path = os.path.join(HERE, 'modal.npz')
    np.savez_compressed(path, **data)
    print('->', path, f'{os.path.getsize(path)/1024:.1f} KiB')


if __name__ == '__main__':
    main()
