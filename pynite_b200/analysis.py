"""Analysis drivers: what `Pynite/Analysis.py` and the three `FEModel3D.analyze*` methods do, with the
numerical work handed to the B200 library.

Host side (this file) keeps the cheap, order-defining bookkeeping of the reference:
`_prepare_model` (`Analysis.py:20-79`), `_identify_combos` (:82-108), `_renumber` (:1543-1571), the
tension/compression-only iteration and load stepping of `_first_order` (:244-326) / `_PDelta`
(:329-471) / `_check_TC_convergence` (:917-1056), and the console messages of `_check_stability`
(:111-168).  Everything numerical -- element stiffness, COO->CSR assembly, partition, load vectors,
the multi-right-hand-side solve, displacements, reactions -- runs in `libpynite_b200.so`.
"""
from __future__ import annotations

import time

import numpy as np
import scipy.sparse as sps

from pynite_b200 import engine as _eng
from pynite_b200 import sharding as _sh
from pynite_b200.LoadCombo import LoadCombo
from pynite_b200.flatten import NodeIndex, flatten_model

_DIRECTION_TEXT = ('for translation in the global X direction.', 'for translation in the global Y direction.',
                   'for translation in the global Z direction.', 'for rotation about the global X axis.',
                   'for rotation about the global Y axis.', 'for rotation about the global Z axis.')
_DOFS = ('DX', 'DY', 'DZ', 'RX', 'RY', 'RZ')

# solver used by the drivers: 'direct' (multifrontal LU + iterative refinement) or 'pcg'
DEFAULT_SOLVER = 'direct'


class RunState:
    """Bookkeeping of the last GPU analysis of a model."""

    def __init__(self, arrays, combo_names):
        self.arrays = arrays
        self.combo_index = {name: i for i, name in enumerate(combo_names)}
        self.engine_combo = dict(self.combo_index)   # combo name -> index inside the engine's combo set
        self.pdelta = False
        self.force_cache = {}


# ---------------------------------------------------------------------------------------------
# model preparation
# ---------------------------------------------------------------------------------------------
def renumber(model):
    """`Analysis._renumber` (:1543-1571): IDs by dictionary order; members are discretised first."""
    for i, node in enumerate(model.nodes.values()):
        node.ID = i
    for i, spring in enumerate(model.springs.values()):
        spring.ID = i
    index = NodeIndex(model) if model.members else None
    k = 0
    for phys_member in model.members.values():
        phys_member.descritize(index)
        for member in phys_member.sub_members.values():
            member.ID = k
            k += 1
    for i, plate in enumerate(model.plates.values()):
        plate.ID = i
    for i, quad in enumerate(model.quads.values()):
        quad.ID = i


def prepare_model(model, n_modes=0):
    """`Analysis._prepare_model` (:20-79): meshes and mat foundations are generated here when they have not been; the
    combinations tagged 'modal' are dropped and `n_modes` fresh ones ('Mode 1', ...) added (:44-51)."""
    model._D = {}
    model._Rxn = {}
    if model.load_combos == {}:
        model.load_combos['Combo 1'] = LoadCombo('Combo 1', factors={'Case 1': 1.0})
    for name in [n for n, c in model.load_combos.items() if c.combo_tags is not None and 'modal' in c.combo_tags]:
        model.load_combos.pop(name)
    for i in range(n_modes):
        model.add_load_combo(f'Mode {i + 1}', {}, ['modal'])
    # (:53-66) meshes, then shear walls -- which create and generate meshes of their own --, then mat foundations
    for table in ('meshes', 'shear_walls', 'mats'):
        for mesh in list(getattr(model, table, {}).values()):
            if not mesh.is_generated or mesh.needs_update:
                mesh.generate()
    combos = list(model.load_combos.keys())
    for spring in model.springs.values():
        for name in combos:
            spring.active[name] = True
    for phys_member in model.members.values():
        for name in combos:
            phys_member.active[name] = True
    renumber(model)


def identify_combos(model, combo_tags=None):
    """`Analysis._identify_combos` (:82-108)."""
    if combo_tags is None:
        return list(model.load_combos.values())
    return [c for c in model.load_combos.values()
            if c.combo_tags is not None and any(tag in c.combo_tags for tag in combo_tags)]


def _engine(model):
    if model._engine is None:
        import os
        model._engine = _eng.Engine(getattr(model, 'device', int(os.environ.get('LOCAL_RANK', '0'))))
    return model._engine


def _report_unstable(model, err):
    """Console lines of `Analysis._check_stability` (:153-166), then the reference's exception."""
    nodes = list(model.nodes.values())
    for dof in err.dofs:
        print('* Nodal instability detected: node ' + nodes[int(dof)//6].name + ' is unstable ' + _DIRECTION_TEXT[int(dof) % 6])
    raise Exception('Unstable node(s). See console output for details.')


def _has_tc_features(model):
    """True when the tension/compression-only machinery of `_check_TC_convergence` can change anything."""
    for node in model.nodes.values():
        for d in _DOFS:
            if getattr(node, 'spring_' + d)[1] is not None:
                return True
    if any(s.tension_only or s.comp_only for s in model.springs.values()):
        return True
    return any(m.tension_only or m.comp_only for m in model.members.values())


def _upload(model, combo_list, active_combo=None):
    A = flatten_model(model, combo_list, active_combo)
    eng = _engine(model)
    eng.set_model(A)
    eng.set_loads(A)
    eng.set_combos(A.factors)
    return A, eng


def _assemble(model, eng, check_stability, log, flags=0):
    if log:
        print('- Assembling the global stiffness matrix on the GPU')
    eng.symbolic(flags)
    eng.assemble_ke()
    if check_stability:
        if log:
            print('- Checking nodal stability')
        try:
            eng.check_nodal_stability()
        except _eng.UnstableNodes as err:
            _report_unstable(model, err)


def _opts(check_stability):
    return _eng.Engine.make_opts(solver=DEFAULT_SOLVER, check_stability=check_stability)


def _store(model, combo_list, D, R):
    for i, combo in enumerate(combo_list):
        model._D[combo.name] = D[i].reshape(-1, 1)
        if R is not None:
            model._Rxn[combo.name] = R[i].reshape(-1, 1)


def _wants_dense(sparse):
    """True when the caller asked for the reference's dense storage (`sparse=False`).  The drivers run the same device
    path either way -- the flag picks a host storage format and solver in the reference, not a different result, and there
    is no CPU solver here -- and the inspection calls (`Ke`, `Kg`, `M`) return dense arrays instead of COO matrices."""
    return not (sparse is True or sparse == True)   # noqa: E712


# ---------------------------------------------------------------------------------------------
# drivers
# ---------------------------------------------------------------------------------------------
def run_linear(model, log, check_stability, check_statics, sparse, combo_tags):
    """`FEModel3D.analyze_linear` (:2250-2332): one stiffness matrix, every combo as one multi-RHS solve."""
    _wants_dense(sparse)
    if log:
        print('+-------------------+')
        print('| Analyzing: Linear |')
        print('+-------------------+')
    t0 = time.perf_counter()
    prepare_model(model)
    combo_list = identify_combos(model, combo_tags)
    first = list(model.load_combos.keys())[0]
    # One process per GPU (torch.distributed initialised by the caller): every rank assembles and factorises K
    # and solves its own contiguous block of combos; results are all-gathered so each rank ends up with all.
    rank, ws = _sh.world()
    distribute = ws > 1 and getattr(model, 'distribute_combos', True)
    local_list = [combo_list[i] for i in _sh.shard_combos(len(combo_list), rank, ws)] if distribute else combo_list
    A, eng = _upload(model, local_list, first)
    t1 = time.perf_counter()
    eng.reset_stats()
    model._run = RunState(A, [c.name for c in local_list])
    # the reference keeps the P-Delta end-force branch alive until `solution` is reset
    model._run.pdelta = model.solution == 'P-Delta'
    D = np.zeros((0, A.n_dof))
    R = np.zeros((0, A.n_dof))
    failure = None
    try:
        _assemble(model, eng, check_stability, log)
        if local_list:
            if log:
                print('- Solving ' + str(len(local_list)) + ' load combination(s)')
            try:
                D, R, st = eng.solve_linear(_opts(check_stability), reactions=not model._run.pdelta)
            except _eng.SingularMatrix:
                raise Exception(_eng.SINGULAR_MSG)
            if model._run.pdelta:
                R = np.stack([eng.reactions(i, True).reshape(-1) for i in range(len(local_list))])
    except Exception as err:              # noqa: BLE001 -- re-raised below, on every rank when distributed
        if not distribute:
            raise
        failure = err
    model.last_stats = eng.stats().as_dict()
    if distribute:
        import torch.distributed as dist
        dev = ('cuda:%d' % eng.device) if dist.get_backend() == 'nccl' else None
        # failure is collective: a rank that raised (singular matrix, unstable node) must not leave the others
        # waiting in the all-gather, and a rank without combos of its own must still learn of it
        if _sh.any_rank_failed(failure is not None, dev):
            raise failure if failure is not None else Exception(_eng.SINGULAR_MSG)
        counts = [len(_sh.shard_combos(len(combo_list), r, ws)) for r in range(ws)]
        D = _sh.all_gather_rows(D, counts, dev)
        R = _sh.all_gather_rows(R, counts, dev)
        # every rank holds K: make all combinations resident so that element forces, stresses and the statics check
        # work for the ones solved elsewhere exactly as for the local ones
        A_all = flatten_factors(model, combo_list, A)
        eng.set_combos(A_all.factors)
        for i in range(len(combo_list)):
            eng.set_displacements(i, D[i])
        model._run = RunState(A_all, [c.name for c in combo_list])
        model._run.pdelta = model.solution == 'P-Delta'
    _store(model, combo_list, D, R)
    t2 = time.perf_counter()
    if model.last_stats is not None:
        model.last_stats['host_prepare_s'] = t1 - t0
        model.last_stats['device_total_s'] = t2 - t1
    if log:
        print('')
        print('- Analysis complete')
        print('')
    if check_statics:
        check_statics_table(model, combo_tags)
    model.solution = 'Linear'


def _set_tc_modes(model, eng, A):
    """Element modes for `pn_tc_update`: bit0 tension-only, bit1 compression-only; physical members as sub-member ranges."""
    smode = np.array([(1 if s.tension_only else 0) | (2 if s.comp_only else 0) for s in model.springs.values()], dtype=np.uint8)
    mmode, starts = [], [0]
    for pm in model.members.values():
        md = (1 if pm.tension_only else 0) | (2 if pm.comp_only else 0)
        mmode.extend([md]*len(pm.sub_members))
        starts.append(len(mmode))
    eng.set_tc_modes(smode, np.array(mmode, dtype=np.uint8), np.array(starts, dtype=np.int64))


def _mirror_activity(model, combo_name, flags, log, before=None):
    """Writes the device's activity flags back into the objects the reference keeps them in (`Node3D.spring_D*[2]`,
    `Spring3D.active[combo]`, `PhysMember.active[combo]` and its sub-members), touching only the elements that can
    change state.  With `log` the deactivations since `before` are printed like `_check_TC_convergence` does."""
    nsf, sa, ma = flags
    nodes = list(model.nodes.values())
    for i, k in zip(*np.nonzero(nsf & 12)):
        getattr(nodes[int(i)], 'spring_' + _DOFS[int(k)])[2] = bool(nsf[i, k] & 2)
    for i, spring in enumerate(model.springs.values()):
        if spring.tension_only or spring.comp_only:
            on = bool(sa[i])
            if log and before is not None and before[1][i] and not on:
                print(f'- Deactivating spring {spring.name}')
            spring.active[combo_name] = on
    k = 0
    for pm in model.members.values():
        nsub = len(pm.sub_members)
        if (pm.tension_only or pm.comp_only) and nsub:
            on = bool(ma[k])
            if log and before is not None and before[2][k] and not on:
                print(f'- Deactivating member {pm.name}')
            pm.active[combo_name] = on
            for sub in pm.sub_members.values():
                sub.active[combo_name] = on
        k += nsub


def _per_combo_loop(model, combo_list, log, check_stability, max_iter, spring_tolerance, member_tolerance,
                    num_steps, pdelta):
    """Shared tension/compression-only iteration: `_first_order` (:244-326) and `_PDelta` (:329-471)."""
    A, eng = _upload(model, combo_list, combo_list[0].name if combo_list else None)
    model._run = RunState(A, [c.name for c in combo_list])
    model._run.pdelta = pdelta
    _set_tc_modes(model, eng, A)
    n = A.n_dof
    end_flags = {}
    flags = _eng.PN_SYM_DENSE_MEMBER_BLOCKS if pdelta else 0
    for ci, combo in enumerate(combo_list):
        if log:
            print('')
            print('- Analyzing load combination ' + combo.name)
        eng.set_combos(A.factors[ci:ci + 1])
        if ci > 0:
            eng.tc_reset()          # springs / members start every combination active; node support springs keep their state
        model._run.engine_combo = {combo.name: 0}
        model._run.force_cache = {}
        total = np.zeros((n, 1))
        load_step = 1
        while load_step <= num_steps:
            iter_count = 1
            converged = False
            while not converged:
                if iter_count > max_iter:
                    if pdelta:
                        raise Exception('- Model diverged during tension/compression-only analysis')
                    raise Exception('Model diverged during tension/compression-only analysis')
                if log:
                    print(f'- Analyzing load step #{load_step}' if not pdelta else
                          '- Beginning tension/compression-only iteration #' + str(iter_count))
                _assemble(model, eng, check_stability, log, flags)
                try:
                    if pdelta:
                        D, _, _ = eng.solve_pdelta(_opts(check_stability), reactions=False)
                    else:
                        D, _, _ = eng.solve_linear(_opts(check_stability), reactions=False)
                except _eng.SingularMatrix:
                    if pdelta:
                        raise ValueError(_eng.SINGULAR_MSG_PDELTA)
                    raise Exception(_eng.SINGULAR_MSG)
                delta = D[0].reshape(-1, 1)/num_steps
                total = total + delta if load_step > 1 else delta
                model._D[combo.name] = total
                eng.set_displacements(0, total)
                model._run.force_cache = {}
                # `_check_TC_convergence` (:917-1056) in kernels: flags, displacements and end forces stay on the device
                if log:
                    print('- Checking for tension/compression-only support spring, spring and member convergence')
                before = eng.activity(A.n_nodes) if log else None
                converged = eng.tc_update(0, pdelta, spring_tolerance, member_tolerance) == 0
                if log and not converged:
                    _mirror_activity(model, combo.name, eng.activity(A.n_nodes), True, before)
                if not converged:
                    if log:
                        print(f'- Undoing load step #{load_step} due to failed convergence.')
                    total = total - delta
                    model._D[combo.name] = total
                else:
                    load_step += 1
                iter_count += 1
        end_flags[combo.name] = eng.activity(A.n_nodes)
        _mirror_activity(model, combo.name, end_flags[combo.name], False)
        # reactions with this combo's final element activity (`_calc_reactions`, :1059-1313)
        model._Rxn[combo.name] = eng.reactions(0, pdelta)
        # element end forces must stay readable after the engine moves on to the next combo
        model._run.saved = getattr(model._run, 'saved', {})
        model._run.saved[combo.name] = {kind: eng.element_forces(kind, 0, local=False, pdelta=pdelta)
                                        for kind in ('spring', 'member', 'quad', 'plate')}
    # The reference computes reactions after ALL combinations (`_calc_reactions` at the end of `analyze`): every
    # element contributes with its own per-combination activity, but the node support springs with the state the LAST
    # combination left them in.  Combinations whose support springs ended differently are recomputed with that state.
    model._run.end_flags = end_flags
    if combo_list:
        final_nsf = end_flags[combo_list[-1].name][0]
        for ci, combo in enumerate(combo_list[:-1]):
            nsf, sa, ma = end_flags[combo.name]
            if np.array_equal(nsf, final_nsf):
                continue
            eng.set_activity(final_nsf, sa, ma)
            eng.set_combos(A.factors[ci:ci + 1])
            eng.set_displacements(0, model._D[combo.name])
            model._Rxn[combo.name] = eng.reactions(0, pdelta)
            model._run.engine_combo = {combo.name: 0}
    model.last_stats = eng.stats().as_dict()


def run_first_order(model, log, check_stability, check_statics, max_iter, sparse, combo_tags,
                    spring_tolerance, member_tolerance, num_steps):
    """`FEModel3D.analyze` (:2334-2411)."""
    _wants_dense(sparse)
    if log:
        print('+-----------+')
        print('| Analyzing |')
        print('+-----------+')
    prepare_model(model)
    combo_list = identify_combos(model, combo_tags)
    if not _has_tc_features(model):
        # No element can change state, so every combo sees the same stiffness matrix and the iteration
        # of `_first_order` ends after one pass: solve all combos together.
        first = list(model.load_combos.keys())[0]
        A, eng = _upload(model, combo_list, first)
        eng.reset_stats()
        model._run = RunState(A, [c.name for c in combo_list])
        _assemble(model, eng, check_stability, log)
        if combo_list:
            try:
                D, R, st = eng.solve_linear(_opts(check_stability), reactions=True)
            except _eng.SingularMatrix:
                raise Exception(_eng.SINGULAR_MSG)
            _store(model, combo_list, D, R)
            model.last_stats = eng.stats().as_dict()
    else:
        _per_combo_loop(model, combo_list, log, check_stability, max_iter, spring_tolerance, member_tolerance,
                        num_steps, pdelta=False)
    if log:
        print('')
        print('- Analysis complete')
        print('')
    if check_statics:
        check_statics_table(model, combo_tags)
    model.solution = 'Nonlinear TC'


def run_pdelta(model, log, check_stability, max_iter, sparse, combo_tags):
    """`FEModel3D.analyze_PDelta` (:2413-2467) -> `Analysis._PDelta` (:329-471)."""
    _wants_dense(sparse)
    if log:
        print('+--------------------+')
        print('| Analyzing: P-Delta |')
        print('+--------------------+')
    prepare_model(model)
    combo_list = identify_combos(model, combo_tags)
    if not _has_tc_features(model):
        first = list(model.load_combos.keys())[0]
        A, eng = _upload(model, combo_list, first)
        eng.reset_stats()
        model._run = RunState(A, [c.name for c in combo_list])
        model._run.pdelta = True
        _assemble(model, eng, check_stability, log, _eng.PN_SYM_DENSE_MEMBER_BLOCKS)
        if combo_list:
            try:
                D, R, st = eng.solve_pdelta(_opts(check_stability), reactions=True)
            except _eng.SingularMatrix:
                raise ValueError(_eng.SINGULAR_MSG_PDELTA)
            _store(model, combo_list, D, R)
            model.last_stats = eng.stats().as_dict()
    else:
        _per_combo_loop(model, combo_list, log, check_stability, max_iter, 0, 0, 1, pdelta=True)
    if log:
        print('')
        print('- Analysis complete')
        print('')
    model.solution = 'P-Delta'


# ---------------------------------------------------------------------------------------------
# inspection helpers behind FEModel3D.Ke / Kg / FER / P and the element objects
# ---------------------------------------------------------------------------------------------
def flatten_factors(model, combo_list, A):
    """Copy of the flattened arrays `A` with the factor table of `combo_list` (same model, same load cases)."""
    import copy
    B = copy.copy(A)
    B.factors = np.array([[c.factors.get(case, 0.0) for case in A.cases] for c in combo_list], dtype=float).reshape(
        len(combo_list), len(A.cases))
    B.combo_names = [c.name for c in combo_list]
    return B


def _fresh_engine_state(model, combo_name):
    """Engine state for a stand-alone `Ke/Kg/FER/P` call.

    Reuses the analysis state when it holds the combo.  Otherwise the model is uploaded to a *scratch* context when an
    analysis exists (its displacements, saved end forces and P-Delta flag must stay readable: in the reference these
    inspection calls never touch stored results), and to the main context when nothing has been analysed yet."""
    run = model._run
    if run is not None and model.solution is not None and combo_name in run.engine_combo:
        return run, _engine(model)
    if any(n.ID is None for n in model.nodes.values()) or model._run is None:
        prepare_model(model)
    combo = model.load_combos[combo_name]
    if model._run is not None and model.solution is not None:
        import os
        if getattr(model, '_scratch_engine', None) is None:
            model._scratch_engine = _eng.Engine(getattr(model, 'device', int(os.environ.get('LOCAL_RANK', '0'))))
        eng = model._scratch_engine
        A = flatten_model(model, [combo], combo_name)
        eng.set_model(A)
        eng.set_loads(A)
        eng.set_combos(A.factors)
        return RunState(A, [combo_name]), eng
    A, eng = _upload(model, [combo], combo_name)
    model._run = RunState(A, [combo_name])
    return model._run, eng


def global_Ke(model, combo_name, log, check_stability, sparse):
    dense = _wants_dense(sparse)
    run, eng = _fresh_engine_state(model, combo_name)
    s = eng.symbolic(0)
    row, col, data = eng.coo()
    if check_stability:
        eng.assemble_ke()
        try:
            eng.check_nodal_stability()
        except _eng.UnstableNodes as err:
            _report_unstable(model, err)
    K = sps.coo_matrix((data, (row, col)), shape=(s.n_dof, s.n_dof))
    return K.toarray() if dense else K


def global_Kg(model, combo_name, log, sparse, first_step):
    dense = _wants_dense(sparse)
    run = model._run
    eng = _engine(model)
    if first_step or run is None or combo_name not in run.engine_combo:
        run, eng = _fresh_engine_state(model, combo_name)
        first_step = True if (model.solution is None or eng is not model._engine) else first_step
    row, col, data = eng.kg_coo(run.engine_combo[combo_name], first_step)
    n = run.arrays.n_dof
    K = sps.coo_matrix((data, (row, col)), shape=(n, n))
    return K.toarray() if dense else K


def global_load_vector(model, combo_name, which):
    run, eng = _fresh_engine_state(model, combo_name)
    return eng.load_vector(run.engine_combo[combo_name], which)


def element_matrix(model, kind, elem_id, which='Ke', P=0.0):
    """Global stiffness matrix `Ke()` of one element (`Member3D.Ke` :1038, `Spring3D.Ke` :191, `Quad3D.Ke` :858,
    `Plate3D.Ke` :502) or a member's geometric stiffness `Kg(P)` (:1048).  One batched launch computes the matrices of
    every element of the kind (`pn_element_ke` / `pn_member_kg`); they are kept with the engine state they belong to, so a
    loop over the elements of a solved model costs one launch."""
    name = next(iter(model.load_combos), 'Combo 1')
    run, eng = _fresh_engine_state(model, name)
    cache = run.__dict__.setdefault('matrix_cache', {})
    key = (kind, which, float(P))
    if key not in cache:
        if which == 'Ke':
            cache[key] = eng.element_ke(kind)
        else:
            cache[key] = eng.member_kg(np.full(len(run.arrays.mem_rot), float(P)))
    return cache[key][elem_id].copy()


def block_T(R, blocks):
    """Transformation matrix with the 3 x 3 direction cosines `R` repeated on the diagonal (`Member3D.T` :1029-1036,
    `Quad3D.T` :941-950)."""
    T = np.zeros((3*blocks, 3*blocks))
    for k in range(blocks):
        T[3*k:3*k + 3, 3*k:3*k + 3] = R
    return T


def element_D(model, nodes, combo_name):
    """Global displacements of an element's nodes as a column (`Quad3D.D` :790-829, `Plate3D.D`, `Spring3D.D`)."""
    D = model._D[combo_name]
    return np.concatenate([D[6*n.ID:6*n.ID + 6, 0] for n in nodes]).reshape(-1, 1)


def element_force(model, kind, elem_id, combo_name, local):
    """End forces of one element for a solved combo (`Member3D.f/F`, `Quad3D.f/F`, `Plate3D.f/F`, `Spring3D.f/F`)."""
    run = model._run
    if run is None or combo_name not in model._D:
        raise KeyError(f"load combination '{combo_name}' has not been analysed")
    key = (kind, combo_name, bool(local))
    if key not in run.force_cache:
        eng = _engine(model)
        saved = getattr(run, 'saved', {})
        if combo_name in saved:
            # per-combination analyses keep every combo's end forces as computed with that combo's own element activity
            F = saved[combo_name][kind]
            run.force_cache[key] = _rotate_saved(model, kind, F) if local else F
        else:
            j = _make_resident(model, combo_name)
            run.force_cache[key] = eng.element_forces(kind, j, local=local, pdelta=run.pdelta)
    return run.force_cache[key][elem_id].reshape(-1, 1)


def _make_resident(model, combo_name):
    """Engine combo index whose stored displacements are those of `combo_name`.  After a per-combination analysis
    (tension/compression-only, load stepping) only the last combination is resident: the requested one is brought
    back with its factor row and its stored displacement vector (`pn_set_displacements`)."""
    run = model._run
    eng = _engine(model)
    if combo_name in run.engine_combo:
        return run.engine_combo[combo_name]
    ci = run.combo_index[combo_name]
    eng.set_combos(run.arrays.factors[ci:ci + 1])
    eng.set_displacements(0, model._D[combo_name])
    run.engine_combo = {combo_name: 0}
    return 0


def shell_stress(model, kind, elem_id, combo_name, u, v, local):
    """(9,) results of one quad / plate at (u, v) for a solved combo: shears, moments, membrane stresses
    (`Quad3D.shear/moment/membrane` :1017-1239, `Plate3D` :590-692).  One batched device call evaluates every element
    of the kind at that point; the table is cached, so walking over the elements costs one launch."""
    run = model._run
    if run is None or combo_name not in model._D:
        raise KeyError(f"load combination '{combo_name}' has not been analysed")
    key = ('stress', kind, combo_name, float(u), float(v), bool(local))
    if key not in run.force_cache:
        j = _make_resident(model, combo_name)
        if len(run.force_cache) > 64:
            run.force_cache = {k: val for k, val in run.force_cache.items() if k[0] != 'stress'}
        run.force_cache[key] = _engine(model).shell_stresses(kind, j, 1, [[u, v]], local)[0, :, 0, :]
    return run.force_cache[key][elem_id]


def shell_stress_points(model, kind, combo_name, pts, local):
    """[count][n_pts][9] results of every quad / plate of the model at a list of points for a solved combo: one
    `pn_shell_stresses` launch (the mesh queries `Mesh.max_shear` ... read corners + centre of all elements at once)."""
    run = model._run
    if run is None or combo_name not in model._D:
        raise KeyError(f"load combination '{combo_name}' has not been analysed")
    j = _make_resident(model, combo_name)
    return _engine(model).shell_stresses(kind, j, 1, [[float(u), float(v)] for u, v in pts], local)[0]


def _rotate_saved(model, kind, F):
    """Global -> local end forces for results saved from an earlier combo of a per-combo analysis."""
    out = np.empty_like(F)
    if kind in ('spring', 'member'):
        elems = list(model.springs.values()) if kind == 'spring' else \
            [s for pm in model.members.values() for s in pm.sub_members.values()]
        for i, e in enumerate(elems):
            R = member_dircos_host(e)
            out[i] = (F[i].reshape(4, 3) @ R.T).reshape(-1)
    else:
        elems = list((model.quads if kind == 'quad' else model.plates).values())
        for i, e in enumerate(elems):
            R = shell_dircos_host(e)
            out[i] = (F[i].reshape(8, 3) @ R.T).reshape(-1)
    return out


def member_dircos_host(sub):
    """3x3 direction cosines of a member/spring (rows = local x, y, z), as `Member3D.T` (:933-1027) builds
    them.  Host copy used only for result post-processing of single elements (axial-force extremes)."""
    from math import isclose
    ni, nj = sub.i_node, sub.j_node
    Xi, Yi, Zi, Xj, Yj, Zj = ni.X, ni.Y, ni.Z, nj.X, nj.Y, nj.Z
    L = ni.distance(nj)
    x = np.array([(Xj - Xi)/L, (Yj - Yi)/L, (Zj - Zi)/L])
    if isclose(Xi, Xj) and isclose(Zi, Zj):
        y = np.array([-1.0, 0, 0]) if Yj > Yi else np.array([1.0, 0, 0])
        z = np.array([0.0, 0, 1.0])
    elif isclose(Yi, Yj):
        y = np.array([0.0, 1.0, 0])
        z = np.cross(x, y)
        z = z/np.linalg.norm(z)
    else:
        proj = np.array([Xj - Xi, 0, Zj - Zi])
        z = np.cross(proj, x) if Yj > Yi else np.cross(x, proj)
        z = z/np.linalg.norm(z)
        y = np.cross(z, x)
        y = y/np.linalg.norm(y)
    rot = getattr(sub, 'rotation', 0.0)
    if rot != 0.0:
        th = np.radians(rot)
        c, s = np.cos(th), np.sin(th)
        y2 = y*c + np.cross(x, y)*s + x*np.dot(x, y)*(1 - c)
        z2 = z*c + np.cross(x, z)*s + x*np.dot(x, z)*(1 - c)
        y, z = y2/np.linalg.norm(y2), z2/np.linalg.norm(z2)
    return np.array([x, y, z])


def shell_dircos_host(e):
    """Shell triad as `Quad3D.T` (:884-950) / `Plate3D.T` (:451-500): x along i->j, z = x cross (i->n)."""
    pi = np.array([e.i_node.X, e.i_node.Y, e.i_node.Z])
    x = np.array([e.j_node.X, e.j_node.Y, e.j_node.Z]) - pi
    xy = np.array([e.n_node.X, e.n_node.Y, e.n_node.Z]) - pi
    x = x/np.linalg.norm(x)
    z = np.cross(x, xy)
    z = z/np.linalg.norm(z)
    y = np.cross(z, x)
    y = y/np.linalg.norm(y)
    return np.array([x, y, z])


def check_statics_table(model, combo_tags=None):
    """Sum of applied loads vs sum of reactions per combo (`Analysis._check_statics`, :1316-1409),
    printed as a plain table.  Uses the global P and FER vectors and the stored reactions."""
    combo_list = identify_combos(model, combo_tags)
    run = model._run
    eng = _engine(model)
    xyz = run.arrays.xyz
    print('+----------------+')
    print('| Statics Check: |')
    print('+----------------+')
    print('')
    header = ['Load Combination', 'Sum FX', 'Sum RX', 'Sum FY', 'Sum RY', 'Sum FZ', 'Sum RZ',
              'Sum MX', 'Sum RMX', 'Sum MY', 'Sum RMY', 'Sum MZ', 'Sum RMZ']
    print(' | '.join(header))
    for combo in combo_list:
        if combo.name not in run.engine_combo:
            continue
        j = run.engine_combo[combo.name]
        F = (eng.load_vector(j, 'P') - eng.load_vector(j, 'FER')).reshape(-1, 6)
        R = model._Rxn[combo.name].reshape(-1, 6)

        def totals(V):
            f = V[:, :3].sum(axis=0)
            m = V[:, 3:].sum(axis=0) + np.cross(xyz, V[:, :3]).sum(axis=0)
            return f, m
        (f, m), (rf, rm) = totals(F), totals(R)
        row = [combo.name] + ['{:.3g}'.format(v) for v in
                              (f[0], rf[0], f[1], rf[1], f[2], rf[2], m[0], rm[0], m[1], rm[1], m[2], rm[2])]
        print(' | '.join(row))
    print('')
