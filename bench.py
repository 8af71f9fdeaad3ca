#!/usr/bin/env python
"""Benchmark of the stiffness-assembly + linear-solve path (BASELINE.json metric: DOF*combos/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE config 3 -- a 500x500 DKMQ quad plate (1 506 006 DOF, 250 000 quads,
CSR nnz 29 540 014) with 64 load combinations per GPU, synthetic data.

One *step* = one pass of the hot path over that model: batched element stiffness -> COO->CSR scatter-add ->
K11/K12/K21/K22 partition -> right-hand sides for all combos -> numeric factorisation -> multi-RHS solve with
iterative refinement -> displacement merge.  `value` times steps with the model resident in HBM (topology
dependent symbolic data -- pattern, scatter map, elimination tree -- built once before the timed region, as it
is reused across re-analyses of one topology).  `e2e` times the whole call sequence a host caller makes
through the C ABI with HOST buffers every step: upload of the SoA arrays, symbolic phase, assembly, solve,
reactions, and the device->host copies of D and the reactions (pinned memory).

Multi-GPU (`--gpus N`, one process per GPU under torchrun): BASELINE config 3 is "64 load combos sharded across
1/2/4/8 GPUs", i.e. STRONG scaling -- the default: the 64 combinations are split across the ranks
(`sharding.shard_combos`), every rank assembles and factorises K and solves its share; `value` = n_dof x 64 x steps
/ max-over-ranks time.  `--scaling weak` gives every rank its own 64 combinations instead (N x the work).

`--impl reference` and the in-arm `cpu_baseline` time the CPU restatement of the reference (oracle/, memoisation
off: the reference recomputes every element) on a BOUNDED SAMPLE of the workload -- a smaller plate with 2 combos --
and say so: the sample's real size is what `config` reports, `same_config` is false, and an extrapolation to
config 3 (element phases linear in the element count, SuperLU with the exponent fitted from two sizes) is given
separately and labelled as such.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'assembly+solve throughput (DOF*combos/s), 1.5M-DOF quad plate x 64 load combos, 1/2/4/8 B200'
UNIT = 'DOF*combos/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--n', type=int, default=500, help='quads per side (500 = BASELINE config 3)')
    ap.add_argument('--combos', type=int, default=64, help='load combinations: total (strong scaling) or per GPU (weak)')
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'])
    ap.add_argument('--dist', default='subtree', choices=['subtree', 'combos'],
                    help='strong scaling on N > 1 GPUs: subtree = elimination tree split across the GPUs, all combos on '
                         'every GPU (pn_set_distributed); combos = combos split, K factorised on every GPU')
    ap.add_argument('--solver', default='direct', choices=['direct', 'pcg'])
    ap.add_argument('--dof-classes', type=int, default=1, choices=[0, 1],
                    help='pn_set_option("dof_classes"): 1 = membrane / bending / drilling fronts (default), 0 = node-based ordering (A/B)')
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--cpu-n', type=int, default=96, help='quads per side of the CPU-baseline sample')
    ap.add_argument('--cpu-combos', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-general-assembly', action='store_true', help='skip the jittered-mesh general assembly roofline entry')
    return ap.parse_args()


def workload_config(args, world):
    n = args.n
    total = args.combos if args.scaling == 'strong' else args.combos*world
    per = [len(x) for x in combo_shards(args, world)]
    subtree = subtree_mode(args, world)
    if subtree:
        sharding = ('strong scaling, subtree-to-GPU multifrontal: every GPU assembles K, factorises and sweeps the subtrees '
                    'below its share of the cut level for all combos, the levels above the cut are replicated; three NCCL '
                    'all-gathers per factor + solve (update matrices, update vectors, solved rows)')
    elif args.scaling == 'strong':
        sharding = 'strong scaling: the combos are split across the GPUs; K is assembled and factorised on every GPU, no data-path collective'
    else:
        sharding = 'weak scaling: every GPU owns its own set of combos; no data-path collective'
    return {'workload': f'BASELINE config 3: {n}x{n} DKMQ quad plate, perimeter pinned, {total} load combos in total',
            'n_dof': 6*(n + 1)*(n + 1), 'n_quads': n*n, 'combos_total': total,
            'combos_per_gpu': [total]*world if subtree else per, 'solver': args.solver, 'sharding': sharding,
            'ordering': ('nested dissection of the node graph, one front per separator and DOF-type class (membrane / bending / '
                         'drilling never meet in a stored entry of a flat plate; verified on the device)' if args.dof_classes
                         else 'nested dissection of the node graph, one front per separator (--dof-classes 0)'),
            'l2': 'inputs larger than L2 (CSR values 236 MB, factor 3.6 GB per step); no explicit flush',
            'symbolic': 'pattern / scatter map / elimination tree built once outside the timed region for `value`, inside for `e2e`'}


def subtree_mode(args, world):
    return world > 1 and args.scaling == 'strong' and args.dist == 'subtree' and args.solver == 'direct'


def combo_shards(args, world):
    """Combo indices owned by every rank.  strong: one table of `--combos` rows split with sharding.shard_combos;
    weak: every rank has its own `--combos` rows."""
    from pynite_b200 import sharding
    if args.scaling == 'strong':
        return [sharding.shard_combos(args.combos, r, world) for r in range(world)]
    return [list(range(args.combos)) for _ in range(world)]


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (numpy element loops + scipy tocsr/SuperLU)
# ---------------------------------------------------------------------------------------------
def cpu_port_step(n, n_combos):
    """One pass of the reference algorithm (restated in oracle/) on an n x n quad plate with n_combos combos.
    Returns (units, seconds, phase dict).  Phases are the ones BASELINE.md section 3 lists.  The oracle's
    memoisation is switched off: the reference evaluates every element and every element load separately."""
    from oracle import model as om
    from pynite_b200 import synthetic
    from scipy.sparse.linalg import spsolve
    A = synthetic.quad_plate_arrays(n=n, n_combos=n_combos)
    ph = {}
    memo = om.MEMOISE
    om.MEMOISE = False
    try:
        t0 = time.perf_counter()
        K = om.ke_csr(A)
        ph['Ke+tocsr'] = time.perf_counter() - t0
        t = time.perf_counter()
        d1, d2, D2 = om.partition_D(A)
        K11, K12, _, _ = om.partition_matrix(K, d1, d2)
        ph['partition'] = time.perf_counter() - t
        ph['FER+P'] = ph['spsolve'] = ph['reactions'] = 0.0
        for c in range(n_combos):
            t = time.perf_counter()
            P, FER = om.load_vectors(A, c)
            ph['FER+P'] += time.perf_counter() - t
            t = time.perf_counter()
            rhs = np.subtract(np.subtract(P[d1, :], FER[d1, :]), K12.tocsr() @ D2)
            D1 = spsolve(K11.tocsr(), rhs)          # one full SuperLU factorisation per combo, as the reference does
            ph['spsolve'] += time.perf_counter() - t
            t = time.perf_counter()
            D = om._merge(A, D1, D2, d1, d2)
            om.reactions(A, c, D)
            ph['reactions'] += time.perf_counter() - t
        total = time.perf_counter() - t0
    finally:
        om.MEMOISE = memo
    return A.n_dof*n_combos, total, {k: round(v, 3) for k, v in ph.items()}


def cpu_sample_n(args):
    """Quads per side of the CPU sample, the same for `--impl reference` and for the in-arm `cpu_baseline`: one step
    costs about 2.1e-3 n^2 seconds (2 combos, one host core); the `--steps/--warmup` run of the reference arm is kept
    near two minutes."""
    budget = 120.0/max(1, args.steps + 1)
    return int(max(16, min(args.cpu_n, (budget/2.1e-3)**0.5)))


def cpu_sample_record(n, combos, phases, seconds):
    return (f'{n}x{n} quad plate ({6*(n + 1)**2} DOF, {n*n} quads) x {combos} combos per step, {seconds:.1f} s: oracle port of the '
            f'reference (numpy element loops, scipy coo->csr, SuperLU spsolve per combo, reactions), one host core; '
            f'phases of the last step (s): {phases}; host cores visible: {os.cpu_count()}')


def extrapolate_to_config3(n, combos, phases, n_small, phases_small, target_n=500, target_combos=64):
    """Config-3 time of the CPU path from the sample: element phases (Ke, FER, reactions, partition) scale with the
    element count (and FER/reactions/spsolve with the combo count); SuperLU with the exponent fitted between the two
    sample sizes.  An EXTRAPOLATION, labelled as such in the record."""
    q, qs, qt = n*n, n_small*n_small, target_n*target_n
    dof, dofs, doft = 6*(n + 1)**2, 6*(n_small + 1)**2, 6*(target_n + 1)**2
    ps = max(phases_small['spsolve'], 1e-9)
    expo = float(np.log(max(phases['spsolve'], 1e-9)/ps)/np.log(dof/dofs))
    cs = target_combos/combos
    t = (phases['Ke+tocsr'] + phases['partition'])*qt/q + (phases['FER+P'] + phases['reactions'])*qt/q*cs + \
        phases['spsolve']*(doft/dof)**expo*cs
    return {'value': doft*target_combos/t, 'unit': UNIT, 'seconds_per_step': t, 'spsolve_exponent_in_dof': expo,
            'fitted_between': [f'{n_small}x{n_small}', f'{n}x{n}'], 'extrapolated': True,
            'note': 'the reference itself cannot run config 3: _store_displacements / _calc_reactions are quadratic (SURVEY H4)'}


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path.  /root/reference (pure Python) does
    not exist on the GPU box, so this times the oracle port of it; every step is a bounded sample, and the record
    says which: `config` describes the SAMPLE that was timed, not config 3."""
    if rank != 0:
        return
    n = cpu_sample_n(args)
    n_small = max(12, n//2)
    _, t_small, phases_small = cpu_port_step(n_small, args.cpu_combos)       # also the warm-up
    t_all, units = 0.0, 0
    phases = None
    for _ in range(args.steps):
        u, t, phases = cpu_port_step(n, args.cpu_combos)
        units += u
        t_all += t
    value = units/t_all
    sample = cpu_sample_record(n, args.cpu_combos, phases, t_all/args.steps)
    cfg = {'workload': f'SAMPLE of BASELINE config 3: {n}x{n} DKMQ quad plate, perimeter pinned, {args.cpu_combos} load combos '
                       f'(config 3 itself is 500x500 x 64 and does not finish on the CPU path)',
           'n_dof': 6*(n + 1)**2, 'n_quads': n*n, 'combos_total': args.cpu_combos, 'sample': True,
           'sample_of': workload_config(args, world)['workload']}
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': 1, 'ms_per_step': 1e3*t_all/args.steps, 'higher_is_better': True, 'scaling': args.scaling,
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': cfg, 'same_config': False, 'sample': True,
            'extrapolated_config3': extrapolate_to_config3(n, args.cpu_combos, phases, n_small, phases_small),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for row in self.rows:
            f = [x.strip() for x in row.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def measure_fp64_gemm_peak(torch, dev):
    """cuBLAS DGEMM burst on this GPU: the denominator for the FP64 dense kernels (MEASURED_PEAKS.json only holds
    the HBM copy and the bf16 GEMM).  4096^3, best of 5, CUDA events on torch's stream."""
    a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    b = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2*4096**3/(best*1e-3)/1e12


def load_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'MEASURED_PEAKS.json'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def general_assembly_roofline(args, local_rank, hbm_peak):
    """The GENERAL numeric assembly path (every element evaluated: k_shell_ke -> element-value buffer -> k_scatter_add)
    on the same mesh with every node moved in plane by U(-0.1, 0.1): no two quads are bit-identical, the element-class
    path of the headline workload declines, and each re-assembly recomputes all 250 000 stiffness matrices.  Reported
    against the same algorithmic bytes (152 B per quad in, 8 B per stored CSR entry out) and, because this path is
    FP64-bound before it is HBM-bound, with its FP64 instruction floor beside it."""
    from pynite_b200 import synthetic
    from pynite_b200.engine import Engine
    Aj = synthetic.quad_plate_arrays(n=args.n, n_combos=1, jitter=0.1)
    e = Engine(local_rank)
    try:
        e.set_model(Aj)
        e.set_loads(Aj)
        e.set_combos(Aj.factors)
        sz = e.symbolic(0)
        for _ in range(3):
            e.assemble_ke()             # the second one tries the class maps (and declines: no compression)
        e.profile_enable(True)
        reps = 5
        for _ in range(reps):
            e.assemble_ke()
        prof = e.profile()
        e.profile_enable(False)
    finally:
        e.close()
    ke_ms = prof.get('element_ke', (0.0, 1))[0]/reps
    sc_ms = prof.get('scatter_add', (0.0, 1))[0]/reps
    nbytes = Aj.quad_ijmn.shape[0]*152 + sz.nnz_csr*8
    ms = ke_ms + sc_ms
    a = nbytes/(ms*1e-3)/1e9 if ms > 0 else 0.0
    return {'bound': 'hbm', 'achieved': a, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': a/hbm_peak, 'algorithmic_bytes': nbytes,
            'ms': ms, 'ms_element_ke': ke_ms, 'ms_scatter_add': sc_ms, 'class_path_used': 'assembly_total' in prof,
            'mesh': f'{args.n}x{args.n} quads, nodes jittered in plane by U(-0.1, 0.1)', 'nnz_csr': int(sz.nnz_csr),
            'note': 'FP64-bound: ~17 kflop per DKMQ quad without contraction of the geometry code = 4.4 GFLOP per assembly'}


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from pynite_b200 import synthetic
    from pynite_b200.engine import Engine

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    shards = combo_shards(args, world)
    subtree = subtree_mode(args, world)
    if subtree:
        # one model, all combinations on every rank; the elimination tree is what is split (pn_set_distributed)
        A = synthetic.quad_plate_arrays(n=args.n, n_combos=args.combos, seed=20261019)
    elif args.scaling == 'strong':
        # one model, one table of --combos combinations; this rank solves rows shards[rank]
        A = synthetic.quad_plate_arrays(n=args.n, n_combos=args.combos, seed=20261019)
        A.factors = np.ascontiguousarray(A.factors[shards[rank]])
        A.combo_names = [A.combo_names[i] for i in shards[rank]]
    else:
        # every rank owns its own --combos combinations of the same mesh (rank-specific factor table and point loads)
        A = synthetic.quad_plate_arrays(n=args.n, n_combos=args.combos, seed=20261019 + rank)
    my_combos = args.combos if subtree else len(shards[rank])
    eng = Engine(local_rank)
    eng.set_option('dof_classes', args.dof_classes)
    if subtree:
        from pynite_b200 import distsolve
        distsolve.attach(eng)
    opts = eng.make_opts(solver=args.solver, check_stability=True, refine_steps=2, tol=1e-10, max_iter=500000)
    units_per_step_all = A.n_dof*sum(len(x) for x in shards)      # whole job, all ranks

    # ---- device-resident arm -----------------------------------------------------------------
    eng.set_model(A)
    eng.set_loads(A)
    eng.set_combos(A.factors)
    sizes = eng.symbolic(0)
    eng.assemble_ke()
    eng.solve_linear(opts, reactions=False, download=False)          # builds the elimination tree (once per topology)
    first_stats = eng.stats().as_dict()

    def step():
        eng.assemble_ke()
        eng.solve_linear(opts, reactions=False, download=False)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    eng.reset_stats()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    eng.timer_start()
    for _ in range(args.steps):
        step()
    ms = eng.timer_stop()
    barrier()
    clocks = sampler.stop() if sampler else None
    st = eng.stats().as_dict()
    # second pass of the same K steps with CUDA events around every launch (per-kernel durations for the rooflines).
    # It is not the timed region of `value`: per-launch events cost ~5 % and the factorisation's side-by-side front
    # groups are serialised while they are on, so that no kernel is charged a concurrent kernel's time.
    eng.profile_enable(True)
    eng.timer_start()
    for _ in range(args.steps):
        step()
    ms_profiled = eng.timer_stop()
    prof = eng.profile()
    eng.profile_enable(False)
    # the numeric assembly runs as one CUDA graph (its total is in `prof`); its kernels are timed one by one in a few
    # extra assemblies with the graph switched off
    eng.set_option('assembly_graph', 0)
    eng.profile_enable(True)
    for _ in range(5):
        eng.assemble_ke()
    prof_asm = eng.profile()
    eng.profile_enable(False)
    eng.set_option('assembly_graph', 1)
    for name in ('element_ke', 'scatter_add'):
        if name in prof_asm and name not in prof:
            prof[name] = (prof_asm[name][0]*args.steps/5.0, prof_asm[name][1]*args.steps/5.0)
    ms = max_over_ranks(ms)
    value = units_per_step_all*args.steps/(ms*1e-3)

    # ---- end-to-end arm: host buffers in, host buffers out, everything inside the timed region ---
    n_dof = A.n_dof
    n_out = len(shards[rank]) if subtree else my_combos          # combos this rank delivers to its host
    D_host = torch.empty((n_out, n_dof), dtype=torch.float64, pin_memory=True).numpy()
    R_host = torch.empty((n_out, n_dof), dtype=torch.float64, pin_memory=True).numpy()
    in_fields = ['xyz', 'support', 'enf_mask', 'enf_val', 'nspring_k', 'nspring_flag', 'quad_ijmn', 'quad_props', 'nl_dof',
                 'nl_case', 'nl_val', 'qp_elem', 'qp_case', 'qp_val', 'factors']
    # the step's inputs live in pinned host memory (same values, page-locked copies of the arrays the builder produced)
    for f in in_fields:
        a = np.ascontiguousarray(getattr(A, f))
        if a.size:
            pinned = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True).numpy()
            pinned[...] = a
            setattr(A, f, pinned)
    h2d = int(sum(np.asarray(getattr(A, f)).nbytes for f in in_fields))
    # reactions cross PCIe as the rows of the DOFs that can carry one (supported, or with an active support spring) when
    # those are few (pn_api.cu reactions_out); the host array is zero-filled and the rows scattered into it
    sup_bits = np.unpackbits(np.asarray(A.support, dtype=np.uint8).reshape(-1, 1), axis=1, bitorder='little')[:, :6].reshape(-1)
    nsf = np.asarray(A.nspring_flag, dtype=np.uint8).reshape(-1)
    n_rxn = int(np.count_nonzero((sup_bits != 0) | (((nsf & 1) != 0) & ((nsf & 2) != 0))))
    rxn_bytes = n_rxn*n_out*8 if n_rxn*8 < n_dof else R_host.nbytes
    d2h = int(D_host.nbytes + rxn_bytes)

    e2e_phases = {}
    if subtree:
        eng.set_output_combos(shards[rank][0] if len(shards[rank]) else 0, len(shards[rank]))

    def e2e_step():
        t = [time.perf_counter()]
        eng.set_model(A)
        eng.set_loads(A)
        eng.set_combos(A.factors)
        t.append(time.perf_counter())
        eng.symbolic(0)
        t.append(time.perf_counter())
        eng.assemble_ke()
        t.append(time.perf_counter())
        if subtree:
            # all ranks hold all of X after the solve; each delivers its share of the combinations to its host
            # (pn_set_output_combos: one copy of the result reaches host memory, 1/N of it per rank)
            eng.solve_linear(opts, reactions=True, D_out=D_host, R_out=R_host)
        else:
            eng.solve_linear(opts, reactions=True, D_out=D_host, R_out=R_host)
        t.append(time.perf_counter())
        for name, a, b in (('upload', 0, 1), ('symbolic', 1, 2), ('assemble', 2, 3), ('solve+reactions+download', 3, 4)):
            e2e_phases[name] = round(1e3*(t[b] - t[a]), 2)

    e2e_step()
    barrier()
    e2e_each = []
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        ts = time.perf_counter()
        e2e_step()
        e2e_each.append(round(1e3*(time.perf_counter() - ts), 2))
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = units_per_step_all*args.e2e_steps/t_e2e
    e2e_stats = eng.stats().as_dict()
    checksum = float(np.abs(D_host).max())

    # ---- per-kernel breakdown and rooflines ---------------------------------------------------------
    hbm_peak, hbm_src = load_peaks()
    f64_peak = measure_fp64_gemm_peak(torch, dev)
    K = args.steps
    total_kernel_ms = sum(v[0] for v in prof.values())
    kern = {}
    for name, (kms, n) in prof.items():
        kern[name] = {'ms_per_step': kms/K, 'launches_per_step': n/K, 'share_of_kernel_time': kms/total_kernel_ms}
    asm_bytes = A.quad_ijmn.shape[0]*152 + sizes.nnz_csr*8                     # 152 B per quad + 8 B per stored entry
    rooflines = {}
    if 'assembly_total' in prof:
        # class-based path: device time from the first assembly launch to the join of the verification stream
        asm_ms = prof['assembly_total'][0]/prof['assembly_total'][1]
        a = asm_bytes/(asm_ms*1e-3)/1e9
        rooflines['assembly(verify + class Ke + recipe gather)'] = {'bound': 'hbm', 'achieved': a, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': a/hbm_peak,
                                                                    'algorithmic_bytes': asm_bytes, 'ms': asm_ms}
        gms = prof['scatter_add'][0]/prof['scatter_add'][1]
        gbytes = sizes.nnz_csr*8 + A.n_nodes*8
        a = gbytes/(gms*1e-3)/1e9
        rooflines['k_recipe_gather (CSR value stream)'] = {'bound': 'hbm', 'achieved': a, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': a/hbm_peak,
                                                           'algorithmic_bytes': gbytes, 'ms': gms}
    else:
        asm_ms = (prof.get('element_ke', (0, 0))[0] + prof.get('scatter_add', (0, 0))[0])/K
        if asm_ms > 0:
            a = asm_bytes/(asm_ms*1e-3)/1e9
            rooflines['assembly(element_ke+scatter_add)'] = {'bound': 'hbm', 'achieved': a, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': a/hbm_peak,
                                                             'algorithmic_bytes': asm_bytes, 'ms': asm_ms}
    for srhs in (1, 8, 64):
        sms = eng.spmm_k11_time(srhs, 20)
        sb = sizes.nnz11*12 + (sizes.n1 + 1)*4 + 2*sizes.n1*srhs*8
        a = sb/(sms*1e-3)/1e9
        rooflines['spmm(K11, %d rhs)' % srhs] = {'bound': 'hbm', 'achieved': a, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': a/hbm_peak,
                                                 'algorithmic_bytes': sb, 'ms': sms}
    if world == 1 and not args.no_general_assembly:
        rooflines['assembly, general geometry (k_shell_ke + k_scatter_add, jittered mesh)'] = general_assembly_roofline(args, local_rank, hbm_peak)
    if 'spmm' in prof:
        s = my_combos
        spmm_bytes = sizes.nnz11*12 + (sizes.n1 + 1)*4 + 2*sizes.n1*s*8
        kms, n = prof['spmm']
        a = spmm_bytes/(kms/n*1e-3)/1e9
        # 10 FP64 instructions per stored entry and column (TwoProd by fma, TwoSum, two accumulations; k_residual_dd in
        # pn_sparse.cu), none of them contractable: against the FP64 issue rate (half the DFMA flop rate measured by
        # tools/microbench/fp64_pipes.cu, 36.7 TFLOP/s) the kernel is FP64-bound before it is HBM-bound
        dd_instr = 10.0*sizes.nnz11*s
        dd_floor_ms = dd_instr/(36.7e12/2)*1e3
        rooflines['residual_dd(K11, %d rhs, double-double)' % s] = {'bound': 'hbm', 'achieved': a, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': a/hbm_peak,
                                              'algorithmic_bytes': spmm_bytes, 'ms': kms/n, 'fp64_instructions': dd_instr,
                                              'fp64_issue_floor_ms': dd_floor_ms, 'frac_of_fp64_issue_rate': dd_floor_ms/(kms/n),
                                              'note': 'FP64-issue-bound: the HBM fraction is reported for comparison with the plain SpMM'}
    if 'lu_update' in prof and first_stats.get('update_flops', 0) > 0:
        kms, n = prof['lu_update']
        a = first_stats['update_flops']*K/(kms*1e-3)/1e12
        rooflines['lu_update(FP64 trailing GEMM)'] = {'bound': 'tensor', 'achieved': a, 'peak': f64_peak, 'unit': 'TFLOP/s', 'frac': a/f64_peak,
                                                      'flops_per_step': first_stats['update_flops'], 'ms': kms/K,
                                                      'peak_source': 'cuBLAS DGEMM 4096^3 burst measured in this run (FP64 is not in MEASURED_PEAKS.json)'}
    if 'solve_update' in prof and first_stats.get('solve_flops', 0) > 0:
        # all kernels of the triangular sweeps: per-step GEMM launches on the narrow top levels (solve_update) and the
        # one-CTA-per-front fused kernels on the wide levels (solve_front_fused: gather + children + steps + store)
        kms = prof['solve_update'][0] + prof.get('solve_front_fused', (0.0, 0))[0]
        n = prof['solve_update'][1] + prof.get('solve_front_fused', (0.0, 0))[1]
        sweeps = 1 + st['iterations']/K          # the first solve plus the refinement corrections actually taken
        fl = first_stats['solve_flops']*my_combos*sweeps
        a = fl*K/(kms*1e-3)/1e12
        rooflines['solve sweeps (solve_update + solve_front_fused, %d rhs)' % my_combos] = {
            'bound': 'tensor', 'achieved': a, 'peak': f64_peak, 'unit': 'TFLOP/s', 'frac': a/f64_peak, 'flops_per_step': fl, 'ms': kms/K,
            'peak_source': 'cuBLAS DGEMM 4096^3 burst measured in this run'}
    dominant = max(prof.items(), key=lambda kv: kv[1][0])[0] if prof else None
    dom_key = next((k for k in rooflines if k.startswith(dominant or '~') or (dominant or '~') in k), None)
    if dom_key is None and rooflines:
        dom_key = max(rooflines, key=lambda k: rooflines[k]['ms'])
    roof = dict(rooflines[dom_key]) if dom_key else None
    if roof is not None:
        roof['kernel'] = dom_key
        roof['traffic'] = None
        # DRAM bytes per launch of the dominant kernel from the committed ncu capture of the same workload (a stopgap
        # the line says so: the traffic is not measured inside this run)
        root = os.path.dirname(os.path.abspath(__file__))
        tpath = os.path.join(root, 'profiles', 'r2_final_dram.json')
        if args.n == 500 and args.combos == 64 and world == 1 and args.dof_classes and os.path.exists(tpath):
            with open(tpath) as f:
                tr = json.load(f)
            if dom_key.startswith('solve sweeps'):
                sw = tr['sweeps']
                roof['traffic'] = sw['dram_bytes_per_launch']
                roof['traffic_unit'] = ('bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu, %d launches of the shared-memory '
                                        'sweep kernels per step; the left-looking sweeps of the top levels add %d launches)'
                                        % (sw['launches'], int(prof['solve_update'][1]/K)))
                roof['traffic_source'] = 'profiles/r2_final_dram.json (committed ncu capture of tools/one_step.py, not measured in this run)'
                roof['flops_per_launch'] = roof['flops_per_step']/(sw['launches'] + prof['solve_update'][1]/K)
        roof['peak_is'] = ('measured (' + hbm_src + ')') if roof['bound'] == 'hbm' else roof.get('peak_source')

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms/args.steps, 'ms_per_step_profiled_pass': ms_profiled/args.steps, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(args, world),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': 1e3*t_e2e/args.e2e_steps, 'steps': args.e2e_steps, 'ms_each_step': e2e_each, 'host_phases_ms_last_step': e2e_phases,
                    'device_phases_ms_last_step': {k: round(v, 2) for k, v in e2e_stats.items() if k.startswith('ms_')},
                    'includes': 'pn_set_* uploads, symbolic phase, elimination-tree analysis, assembly, solve, reactions, D and Rxn copies to pinned host memory'},
            'gpu_launches': int(st['kernel_launches']), 'clocks': clocks,
            'roofline': roof, 'rooflines': rooflines, 'kernels': kern,
            'phases_ms_per_step': {k: st[k]/K for k in ('ms_elements', 'ms_scatter', 'ms_partition', 'ms_rhs', 'ms_factor', 'ms_solve')},
            'solver_stats': {'factor_nnz': first_stats['factor_nnz'], 'factor_flops': first_stats['factor_flops'],
                             'counters': eng.counters(),
                             'max_rel_residual': e2e_stats['max_rel_residual'], 'refine_steps': opts.refine_steps,
                             'fp64_gemm_peak_tflops_measured': f64_peak},
            'sizes': {'n_dof': int(sizes.n_dof), 'n1': int(sizes.n1), 'nnz_coo': int(sizes.nnz_coo), 'nnz_csr': int(sizes.nnz_csr),
                      'nnz11': int(sizes.nnz11)},
            'checksum_max_abs_D': checksum}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = cpu_sample_n(args)
        u, t, phases = cpu_port_step(n_cpu, args.cpu_combos)
        line['cpu_baseline'] = {'value': u/t, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'sample_is_config3': False,
                                'sample': cpu_sample_record(n_cpu, args.cpu_combos, phases, t)}
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        # torchrun exports OMP_NUM_THREADS=1; the host-side symbolic phases (ordering, class / recipe maps) are OpenMP
        # loops, so give every rank its share of the host cores before the library (libgomp) is loaded
        os.environ['OMP_NUM_THREADS'] = str(max(1, (os.cpu_count() or 1)//world))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
