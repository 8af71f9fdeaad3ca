"""Helpers shared by the CPU and GPU tests: fixture loading and comparison metrics."""
import os
from types import SimpleNamespace

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

FIXTURES = ['space_frame', 'portal_frame', 'continuous_beam', 'quad_mesh_xy', 'quad_mesh_yz_ortho', 'plate_mesh_xz',
            'skew_shell', 'moment_frame', 'mat_foundation', 'mixed_building', 'pdelta_column', 'pdelta_frame',
            'braced_tc']


def load_fixture(name):
    """Returns (A, ref, mode): flattened input arrays (attribute access like `flatten.ModelArrays`),
    the dict of reference outputs, and the analysis mode the reference ran."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    A = SimpleNamespace()
    ref = {}
    for k in z.files:
        if k.startswith('in_'):
            setattr(A, k[3:], z[k])
        elif k.startswith('ref_'):
            ref[k[4:]] = z[k]
    A.n_nodes = int(A.n_nodes)
    A.n_dof = 6*A.n_nodes
    A.cases = [str(c) for c in A.cases]
    A.combo_names = [str(c) for c in A.combo_names]
    return A, ref, str(z['mode'])


def rel_err(a, b):
    """max |a - b| / max |b| (vector-relative), 0 when both are empty or identical."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    if a.size == 0 and b.size == 0:
        return 0.0
    scale = np.abs(b).max()
    d = np.abs(a - b).max()
    return float(d if scale == 0 else d/scale)
