"""Member result diagrams (`pynite_b200/diagrams.py`) against the reference's (tests/golden/member_diagrams.npz, written by
tests/golden/make_diagram_golden.py running the reference): sampled arrays, point values and extremes of shear, moment,
axial force, torque and deflection for every physical member and combination of eight fixtures (linear, P-Delta,
tension/compression-only with slack members, load stepping).

The diagrams are host arithmetic on the solve's results, so these tests feed the model the reference's own results -- the
stored displacements and the sub-members' local end forces of the fixture files, which the GPU parity tests hold the CUDA
path to -- instead of running a solve."""
import os

import numpy as np
import pytest

from pynite_b200 import FEModel3D, analysis
from tests import models
from tests.util import DIAGRAM_FIXTURES, diagram_record, load_fixture

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'member_diagrams.npz'))


def solved_like_the_reference(name):
    build, mode = models.SMALL_MODELS[name]
    _, ref, _ = load_fixture(name)
    m = build(FEModel3D)
    analysis.prepare_model(m)
    combos = list(m.load_combos.keys())
    subs = [s for pm in m.members.values() for s in pm.sub_members.values()]
    assert [s.name for s in subs] == [str(n) for n in ref['sub_names']]
    for c, combo in enumerate(combos):
        m._D[combo] = ref['D'][c].reshape(-1, 1).copy()
        for k, pm in enumerate(m.members.values()):
            on = bool(GOLD[name + '/active'][c, k])
            pm.active[combo] = on
            for s in pm.sub_members.values():
                s.active[combo] = on
    forces = {(combo, s.ID): ref['member_f'][c, i].reshape(12, 1) for c, combo in enumerate(combos)
              for i, s in enumerate(subs)}
    m._element_force = lambda kind, elem_id, combo_name, local: forces[(combo_name, elem_id)]
    m.solution = str(ref['solution'])
    return m


def close(ours, ref, tol):
    scale = max(np.abs(ref).max(), 1e-300)
    return np.abs(ours - ref).max() <= tol*scale


@pytest.mark.parametrize('name', DIAGRAM_FIXTURES)
def test_diagrams_match_reference(name):
    m = solved_like_the_reference(name)
    rec = diagram_record(m)
    assert list(rec['member_names']) == [str(n) for n in GOLD[name + '/member_names']]
    assert np.array_equal(rec['arrays'][:, :, :, 0], GOLD[name + '/arrays'][:, :, :, 0])          # stations
    nq = rec['arrays'].shape[2]
    for key in ('arrays', 'points', 'extremes'):
        ours, ref = rec[key], GOLD[name + '/' + key]
        assert ours.shape == ref.shape
        for q in range(ours.shape[2]):                       # per quantity: forces and deflections differ by orders
            a, b = (ours[:, :, q, 1], ref[:, :, q, 1]) if key == 'arrays' else (ours[:, :, q], ref[:, :, q])
            assert close(a, b, 1e-9), (key, q, float(np.abs(a - b).max()), float(np.abs(b).max()))
    assert nq == 9


def test_tag_lists_and_errors():
    m = solved_like_the_reference('portal_frame')
    names = list(m.load_combos.keys())
    m.load_combos[names[0]].combo_tags = ['service']
    m.load_combos[names[1]].combo_tags = ['service', 'strength']
    pm = list(m.members.values())[0]
    best, governing = pm.max_moment('Mz', ['service'])
    each = {n: pm.max_moment('Mz', n) for n in names[:2]}
    assert governing == max(each, key=each.get) and best == each[governing]
    worst, governing = pm.min_shear('Fy', ['strength'])
    assert governing == names[1] and worst == pm.min_shear('Fy', names[1])
    assert pm.max_axial(['no such tag']) == (0, None)
    with pytest.raises(ValueError):
        pm.find_member(pm.L()*1.5)
    with pytest.raises(ValueError):
        pm.moment('Mq', 0.0, names[0])
    with pytest.raises(ValueError):
        pm.shear_array('Fy', 5, names[0], x_array=np.array([-1.0, 0.5]))
    sub = list(pm.sub_members.values())[0]
    xs = np.array([0.0, 0.25, 0.5])*sub.L()
    arr = sub.moment_array('Mz', 0, names[0], x_array=xs)
    assert np.allclose(arr[1], [sub.moment('Mz', x, names[0]) for x in xs], rtol=1e-12, atol=0)
    assert sub.T().shape == (12, 12) and np.allclose(sub.T() @ sub.T().T, np.eye(12))


def test_plot_methods_draw_the_arrays(monkeypatch):
    """`plot_*` with a recording stand-in for matplotlib (absent from this image): one line for a single combination, the
    max / min envelope for a list of tags."""
    import sys
    import types
    calls = []

    class Axes:
        def __getattr__(self, name):
            return lambda *a, **k: calls.append((name, a, k))

    pyplot = types.ModuleType('matplotlib.pyplot')
    pyplot.subplots = lambda **k: (object(), Axes())
    pyplot.show = lambda: calls.append(('show', (), {}))
    pkg = types.ModuleType('matplotlib')
    pkg.pyplot = pyplot
    monkeypatch.setitem(sys.modules, 'matplotlib', pkg)
    monkeypatch.setitem(sys.modules, 'matplotlib.pyplot', pyplot)
    m = solved_like_the_reference('portal_frame')
    names = list(m.load_combos.keys())
    pm = list(m.members.values())[0]
    pm.plot_moment('Mz', names[0], n_points=7)
    line = [c for c in calls if c[0] == 'plot'][0]
    assert np.array_equal(line[1][1], pm.moment_array('Mz', 7, names[0])[1])
    assert ('set_ylabel', ('Moment',), {}) in calls and calls[-1][0] == 'show'
    del calls[:]
    for name in names[:2]:
        m.load_combos[name].combo_tags = ['env']
    pm.plot_shear('Fy', ['env'], n_points=5)
    high, low = [c for c in calls if c[0] == 'plot']
    both = np.array([pm.shear_array('Fy', 5, n)[1] for n in names[:2]])
    assert np.array_equal(high[1][1], both.max(axis=0)) and np.array_equal(low[1][1], both.min(axis=0))
    for method, args in (('plot_torque', ()), ('plot_axial', ()), ('plot_deflection', ('dy',)), ('plot_rel_deflection', ('dy',))):
        getattr(list(pm.sub_members.values())[0], method)(*args, names[0], 4)
    assert [c[1][0] for c in calls if c[0] == 'set_ylabel'][-4:] == ['Torsional Moment (Warping Torsion Not Included)',
                                                                    'Axial Force', 'Deflection', 'Relative Deflection']
