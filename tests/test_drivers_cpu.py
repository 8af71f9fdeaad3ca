"""The Python drivers end to end on the CPU, with `tests/oracle_engine.OracleEngine` standing in for `engine.Engine`
(test infrastructure: the oracle does the arithmetic the CUDA library does on the B200).  What is under test is the glue the
GPU tests of the same names exercise on hardware -- upload, result storing, the per-object queries after an analysis, the
modal operator, the shear-wall results -- for the features added after the round's GPU minutes were spent
(`unverified_gpu` tests): with this file green, what those GPU tests still have to show is only that the kernels accept the
models, not that the Python around them is right."""
import os

import numpy as np
import pytest

from pynite_b200 import FEModel3D
from tests import models, oracle_engine
from tests.known_answers import DIAGRAM_KNOWN_ANSWERS, KNOWN_ANSWERS
from tests.test_known_answers import _check, _diagram_value
from tests.util import SHEAR_WALL_COMBOS, diagram_record, load_fixture, rel_err, shear_wall_record

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(autouse=True)
def oracle_backed_engine(monkeypatch):
    oracle_engine.install(monkeypatch)


def _analyze(m, mode):
    if mode == 'pdelta':
        m.analyze_PDelta()
    elif mode == 'linear':
        m.analyze_linear()
    else:
        m.analyze()


@pytest.mark.parametrize('name', list(KNOWN_ANSWERS))
def test_known_answers_through_the_public_api(name):
    m, mode, checks = KNOWN_ANSWERS[name](FEModel3D)
    _analyze(m, mode)
    for kind, node, comp, expected, rel_tol, abs_tol in checks:
        value = getattr(m.nodes[node], comp if kind == 'D' else 'Rxn' + comp)[next(iter(m.load_combos))]
        _check(value, expected, rel_tol, abs_tol, (name, kind, node, comp))
    if name == 'torque_beam':
        assert abs(m.members['M1'].max_torque() - 8.93) < 5e-3 and abs(m.members['M1'].min_torque() + 6.07) < 5e-3


@pytest.mark.parametrize('name', list(DIAGRAM_KNOWN_ANSWERS))
def test_diagram_known_answers_through_the_public_api(name):
    m, mode, checks = DIAGRAM_KNOWN_ANSWERS[name](FEModel3D)
    _analyze(m, mode)
    for query, member, direction, station, combo, expected, rel_tol, *abs_tol in checks:
        value = _diagram_value(m.members[member], query, direction, station, combo)
        _check(value, expected, rel_tol, abs_tol[0] if abs_tol else 0, (name, query, combo))


@pytest.mark.parametrize('name', ['portal_frame', 'continuous_beam', 'pdelta_column'])
def test_diagrams_after_an_analysis(name):
    gold = np.load(os.path.join(GOLDEN, 'member_diagrams.npz'))
    build, mode = models.SMALL_MODELS[name]
    m = build(FEModel3D)
    _analyze(m, mode)
    rec = diagram_record(m)
    tol = 1e-5 if mode == 'pdelta' else 1e-8
    for key in ('arrays', 'points', 'extremes'):
        ours, ref = rec[key], gold[name + '/' + key]
        for q in range(ours.shape[2]):
            a, b = (ours[:, :, q, 1], ref[:, :, q, 1]) if key == 'arrays' else (ours[:, :, q], ref[:, :, q])
            assert np.abs(a - b).max() <= tol*max(np.abs(b).max(), 1e-300), (key, q)


@pytest.mark.parametrize('name', ['portal_frame', 'mixed_building', 'plate_mesh_xz'])
def test_element_matrices_through_the_objects(name):
    build, mode = models.SMALL_MODELS[name]
    _, ref, _ = load_fixture(name)
    m = build(FEModel3D)
    m.analyze_linear()
    subs = [s for pm in m.members.values() for s in pm.sub_members.values()]
    for kind, elems in (('member', subs), ('spring', list(m.springs.values())), ('quad', list(m.quads.values())),
                        ('plate', list(m.plates.values()))):
        for k, e in enumerate(elems):
            assert rel_err(e.Ke(), ref[kind + '_Ke'][k]) < 1e-12, (kind, k)
            T = e.T()
            assert rel_err(e.ke(), T @ ref[kind + '_Ke'][k] @ T.T) < 1e-12, (kind, k)
    if subs:
        assert np.allclose(subs[0].kg(12.0), subs[0].T() @ subs[0].Kg(12.0) @ subs[0].T().T)
    first = next(iter(m.load_combos))
    assert m.solution == 'Linear' and first in m._D
    if subs:
        assert subs[0].f(first).shape == (12, 1)           # results still readable after the inspection calls


def test_inspection_of_an_unsolved_model():
    m = models.SMALL_MODELS['portal_frame'][0](FEModel3D)
    pm = next(iter(m.members.values()))
    K = pm.Ke() if len(pm.sub_members) <= 1 else next(iter(pm.sub_members.values())).Ke()
    assert K.shape == (12, 12) and np.allclose(K, K.T)


def test_shear_wall_through_analyze_linear():
    gold = np.load(os.path.join(GOLDEN, 'shear_walls.npz'))
    m = models.shear_wall_case(FEModel3D)
    m.analyze_linear()
    assert set(SHEAR_WALL_COMBOS) <= set(m._D)
    for key, ours in shear_wall_record(m).items():
        ref = gold[key]
        if ours.ndim == 3:
            for k in range(4):
                assert np.abs(ours[:, :, k] - ref[:, :, k]).max() <= 1e-7*np.abs(ref[:, :, k]).max(), (key, k)
        else:
            assert abs(ours/ref - 1) <= 1e-7, key


@pytest.mark.parametrize('name', [n for n, (_, kw) in models.MODAL_MODELS.items() if kw['num_modes']])
def test_modal_through_the_engine_operator(name):
    """`modal.EngineInverse`: the Lanczos block goes in as nodal load cases with a unit factor table and comes back as the
    free rows of the displacement table -- through `set_loads` / `set_combos` / `solve_linear`, as on the GPU."""
    gold = np.load(os.path.join(GOLDEN, 'modal.npz'))
    build, kw = models.MODAL_MODELS[name]
    m = build(FEModel3D)
    m.analyze_modal(log=False, **kw)
    assert np.max(np.abs(m.frequencies/gold[name + '/frequencies'] - 1)) < 1e-9
    assert m.solution == 'Modal' and m.modal_info['converged']
    m.analyze_linear()                                       # the mode combinations are dropped again
    assert not [c for c in m.load_combos if c.startswith('Mode')] and m.solution == 'Linear'
