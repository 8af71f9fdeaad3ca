"""Generates the golden fixtures under tests/golden/*.npz by RUNNING THE REFERENCE ITSELF.

Run in the development container only (the reference is not present on the GPU box):

    python tests/golden/make_golden.py            # all models in tests/models.py::SMALL_MODELS
    python tests/golden/make_golden.py quad_mesh_xy

For every model the script
  1. builds it with the reference `Pynite.FEModel3D` (imported from /root/reference with the two stub
     packages in tests/golden/_stubs that replace the absent `prettytable` / `matplotlib`),
  2. runs the reference's own driver (`analyze_linear`, `analyze` or `analyze_PDelta`, sparse=True),
  3. records what the reference produced on this path: COO triplets of `Ke()` in emission order, the CSR
     after `.tocsr()`, `_partition_D` lists, the four partition blocks' patterns, `FER`/`P`/`_D` and the
     per-node reactions of every combo, every element's global `Ke()`, element end forces `F()`, and for
     P-Delta models the `Kg()` triplets of the solved state,
  4. builds the same model with `pynite_b200.FEModel3D`, flattens BOTH object graphs with
     `pynite_b200.flatten.flatten_model`, checks they are identical, and stores the flattened input arrays.
The fixtures are the pin for `oracle/` (tests/test_oracle_golden.py) and the target of the GPU parity
tests (tests/test_gpu_parity.py).
"""
import io
import os
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, '_stubs'))
sys.path.insert(0, '/root/reference')

from Pynite import FEModel3D as RefModel          # noqa: E402
from Pynite import Analysis as RefAnalysis        # noqa: E402

from pynite_b200 import FEModel3D as OurModel     # noqa: E402
from pynite_b200 import analysis as our_analysis  # noqa: E402
from pynite_b200.flatten import flatten_model     # noqa: E402
from tests import models                          # noqa: E402

ARRAY_FIELDS = ['xyz', 'support', 'enf_mask', 'enf_val', 'nspring_k', 'nspring_flag', 'spring_ij', 'spring_ks',
                'spring_active', 'mem_ij', 'mem_props', 'mem_rot', 'mem_rel', 'mem_active', 'quad_ijmn',
                'quad_props', 'plate_ijmn', 'plate_props', 'nl_dof', 'nl_case', 'nl_val', 'ml_member', 'ml_case',
                'ml_kind', 'ml_params', 'qp_elem', 'qp_case', 'qp_val', 'pp_elem', 'pp_case', 'pp_val', 'factors']
RXN = ('RxnFX', 'RxnFY', 'RxnFZ', 'RxnMX', 'RxnMY', 'RxnMZ')


def reference_run(build, mode):
    ref = build(RefModel)
    with contextlib.redirect_stdout(io.StringIO()):
        if mode == 'linear':
            ref.analyze_linear(sparse=True)
        elif mode == 'tc':
            ref.analyze(sparse=True)
        elif mode == 'tc_steps':
            ref.analyze(sparse=True, **models.TC_STEPS_KW)
        else:
            ref.analyze_PDelta(sparse=True)
    return ref


def collect(ref, mode):
    out = {}
    combos = list(ref.load_combos.keys())
    first = combos[0]
    n = 6*len(ref.nodes)
    # For tension/compression-only models the matrix depends on the combo's final element activity, so the
    # "Ke" snapshot is taken for the first combo after the analysis finished, as `_calc_reactions` sees it.
    K = ref.Ke(first, False, False, True)
    out['ref_coo_row'] = np.asarray(K.row, dtype=np.int64)
    out['ref_coo_col'] = np.asarray(K.col, dtype=np.int64)
    out['ref_coo_data'] = np.asarray(K.data, dtype=np.float64)
    csr = K.tocsr()
    out['ref_indptr'] = csr.indptr.astype(np.int32)
    out['ref_indices'] = csr.indices.astype(np.int32)
    out['ref_data'] = csr.data.astype(np.float64)
    d1, d2, D2 = RefAnalysis._partition_D(ref)
    out['ref_d1'] = np.asarray(d1, dtype=np.int32)
    out['ref_d2'] = np.asarray(d2, dtype=np.int32)
    out['ref_D2'] = np.asarray(D2, dtype=np.float64).reshape(-1)
    blocks = RefAnalysis._partition(ref, csr, d1, d2)
    for name, B in zip(('11', '12', '21', '22'), blocks):
        B = B.tocsr()
        out['ref_K' + name + '_indptr'] = B.indptr.astype(np.int32)
        out['ref_K' + name + '_indices'] = B.indices.astype(np.int32)
        out['ref_K' + name + '_data'] = B.data.astype(np.float64)
    out['ref_D'] = np.stack([np.asarray(ref._D[c]).reshape(-1) for c in combos])
    R = np.zeros((len(combos), n))
    for ci, c in enumerate(combos):
        for node in ref.nodes.values():
            for k, r in enumerate(RXN):
                R[ci, node.ID*6 + k] = getattr(node, r).get(c, 0.0)
    out['ref_Rxn'] = R
    out['ref_FER'] = np.stack([np.asarray(ref.FER(c)).reshape(-1) for c in combos])
    out['ref_P'] = np.stack([np.asarray(ref.P(c)).reshape(-1) for c in combos])
    subs = [s for pm in ref.members.values() for s in pm.sub_members.values()]
    out['ref_spring_Ke'] = np.array([np.asarray(s.Ke()) for s in ref.springs.values()]).reshape(-1, 12, 12)
    out['ref_member_Ke'] = np.array([np.asarray(s.Ke()) for s in subs]).reshape(-1, 12, 12)
    out['ref_quad_Ke'] = np.array([np.asarray(q.Ke()) for q in ref.quads.values()]).reshape(-1, 24, 24)
    out['ref_plate_Ke'] = np.array([np.asarray(p.Ke()) for p in ref.plates.values()]).reshape(-1, 24, 24)
    out['ref_member_F'] = np.array([[np.asarray(s.F(c)).reshape(-1) for s in subs] for c in combos]).reshape(len(combos), -1, 12)
    out['ref_member_f'] = np.array([[np.asarray(s.f(c)).reshape(-1) for s in subs] for c in combos]).reshape(len(combos), -1, 12)
    out['ref_spring_F'] = np.array([[np.asarray(s.F(c)).reshape(-1) for s in ref.springs.values()]
                                    for c in combos]).reshape(len(combos), -1, 12)
    out['ref_quad_F'] = np.array([[np.asarray(q.F(c)).reshape(-1) for q in ref.quads.values()]
                                  for c in combos]).reshape(len(combos), -1, 24)
    out['ref_plate_F'] = np.array([[np.asarray(p.F(c)).reshape(-1) for p in ref.plates.values()]
                                   for c in combos]).reshape(len(combos), -1, 24)
    out['ref_sub_names'] = np.array([s.name for s in subs])
    # internal shears / moments / membrane stresses (Quad3D.py:1017-1239, Plate3D.py:590-692) at the points of
    # oracle.model.STRESS_POINTS (quads, natural coordinates) / PLATE_FRACTIONS (plates, fractions of width and height)
    from oracle.model import STRESS_POINTS, PLATE_FRACTIONS

    def nine(Q, M, S):
        v = np.zeros(9)
        Q, M, S = (np.asarray(x, dtype=float).reshape(-1) for x in (Q, M, S))
        v[:len(Q)] = Q
        v[3:6] = M
        v[6:9] = S
        return v
    for loc, tag in ((True, 'local'), (False, 'global')):
        out['ref_quad_stress_' + tag] = np.array(
            [[[nine(q.shear(xi, eta, loc, c), q.moment(xi, eta, loc, c), q.membrane(xi, eta, loc, c))
               for (xi, eta) in STRESS_POINTS] for q in ref.quads.values()] for c in combos]).reshape(
            len(combos), len(ref.quads), len(STRESS_POINTS), 9)
    out['ref_plate_stress'] = np.array(
        [[[nine(p.shear(fx*p.width(), fy*p.height(), True, c), p.moment(fx*p.width(), fy*p.height(), True, c),
                p.membrane(fx*p.width(), fy*p.height(), True, c))
           for (fx, fy) in PLATE_FRACTIONS] for p in ref.plates.values()] for c in combos]).reshape(
        len(combos), len(ref.plates), len(PLATE_FRACTIONS), 9)
    if mode == 'pdelta':
        Kg = ref.Kg(first, False, True, False)
        out['ref_kg_row'] = np.asarray(Kg.row, dtype=np.int64)
        out['ref_kg_col'] = np.asarray(Kg.col, dtype=np.int64)
        out['ref_kg_data'] = np.asarray(Kg.data, dtype=np.float64)
        P = []
        for s in subs:
            d = np.asarray(s.d(first)).reshape(-1)
            P.append(s.material.E*s.section.A/s.L()*(d[6] - d[0]))
        out['ref_member_P'] = np.array(P)
        out['ref_member_Kg'] = np.array([np.asarray(s.Kg(p)) for s, p in zip(subs, P)]).reshape(-1, 12, 12)
    out['ref_solution'] = np.array(str(ref.solution))
    return out


def flatten_pair(build, ref, mode):
    ours = build(OurModel)
    our_analysis.prepare_model(ours)
    combos = list(ours.load_combos.values())
    A = flatten_model(ours, combos)
    # For T/C models the reference's final activity flags are what the Ke snapshot was taken with.
    B = flatten_model(ref, list(ref.load_combos.values()))
    for f in ARRAY_FIELDS:
        a, b = getattr(A, f), getattr(B, f)
        if mode in ('tc', 'tc_steps') and f in ('spring_active', 'mem_active', 'nspring_flag'):
            continue
        if a.shape != b.shape or not np.array_equal(a, b):
            raise AssertionError(f'flattened field {f} differs between pynite_b200 and the reference object graph')
    assert A.cases == B.cases and A.combo_names == B.combo_names
    assert [s.name for s in A.sub_members] == [s.name for s in B.sub_members]
    src = B if mode in ('tc', 'tc_steps') else A
    out = {'in_' + f: getattr(src, f) for f in ARRAY_FIELDS}
    out['in_n_nodes'] = np.array(A.n_nodes)
    out['in_cases'] = np.array(A.cases)
    out['in_combo_names'] = np.array(A.combo_names)
    return out


MESH_QUERIES = [('shear', d) for d in ('Qx', 'Qy', 'QX', 'QY')] + [('moment', d) for d in ('Mx', 'My', 'Mxy', 'MX', 'MY', 'MZ')] + \
               [('membrane', d) for d in ('Sx', 'Sy', 'Sxy', 'SX', 'SY')]


def mesh_extremes(ref):
    """`Mesh.max_* / min_*` (`Mesh.py:172-716`) of every mesh for every combination: [mesh][combo][query][max, min]."""
    out = np.zeros((len(ref.meshes), len(ref.load_combos), len(MESH_QUERIES), 2))
    for a, mesh in enumerate(ref.meshes.values()):
        for b, combo in enumerate(ref.load_combos):
            for q, (what, direction) in enumerate(MESH_QUERIES):
                out[a, b, q, 0] = getattr(mesh, 'max_' + what)(direction, combo)
                out[a, b, q, 1] = getattr(mesh, 'min_' + what)(direction, combo)
    return {'ref_mesh_extremes': out}


def main(argv):
    names = argv[1:] or list(models.SMALL_MODELS.keys())
    for name in names:
        build, mode = models.SMALL_MODELS[name]
        ref = reference_run(build, mode)
        data = collect(ref, mode)
        data.update(flatten_pair(build, ref, mode))
        if getattr(ref, 'meshes', None):
            data.update(mesh_extremes(ref))
        if getattr(ref, 'mats', None):
            data['ref_soil_pressure'] = np.array([[[mat.soil_pressure(x, z, combo) for (x, z) in models.MAT_PRESSURE_POINTS]
                                                   for combo in ref.load_combos] for mat in ref.mats.values()], dtype=float)
        data['mode'] = np.array(mode)
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **data)
        print(f'{name:22s} mode={mode:7s} dof={6*len(ref.nodes):6d} coo={len(data["ref_coo_data"]):7d} '
              f'csr={len(data["ref_data"]):7d} -> {os.path.getsize(path)/1024:.0f} KiB')


if __name__ == '__main__':
    main(sys.argv)
