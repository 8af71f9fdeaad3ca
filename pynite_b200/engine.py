"""ctypes binding of `libpynite_b200.so` (C ABI in `include/pynite_b200.h`).

This is the only place the package touches native code.  There is no CPU implementation behind any
call: if the shared library is missing, or no CUDA device is visible, the first analysis call
raises.  Error codes are mapped to the exception types and messages the reference raises at the
corresponding places (`Analysis.py:163-166, 172, 239, 442`).
"""
from __future__ import annotations

import ctypes as ct
import os

import numpy as np

_LIB_NAME = 'libpynite_b200.so'
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)

PN_SPRING, PN_MEMBER, PN_QUAD, PN_PLATE = 1, 2, 3, 4
PN_SYM_DENSE_MEMBER_BLOCKS = 1
PN_SOLVER_DIRECT, PN_SOLVER_PCG = 0, 1
PN_OK, PN_ERR_CUDA, PN_ERR_ARG, PN_ERR_SINGULAR, PN_ERR_UNSTABLE_NODE, PN_ERR_NOT_CONVERGED, PN_ERR_NO_DEVICE = \
    0, -1, -2, -3, -4, -5, -6

SINGULAR_MSG = ('The stiffness matrix is singular, which implies rigid body motion. The structure '
                'is unstable. Aborting analysis.')                       # Analysis.py:171-172
SINGULAR_MSG_PDELTA = 'The stiffness matrix is singular, which indicates that the structure is unstable.'   # :442

KIND = {'spring': PN_SPRING, 'member': PN_MEMBER, 'quad': PN_QUAD, 'plate': PN_PLATE}
_ND = {PN_SPRING: 12, PN_MEMBER: 12, PN_QUAD: 24, PN_PLATE: 24}


class EngineUnavailable(RuntimeError):
    """The CUDA library could not be loaded or no B200 is visible."""


class SingularMatrix(Exception):
    pass


class NotConverged(Exception):
    pass


class UnstableNodes(Exception):
    def __init__(self, msg, dofs):
        super().__init__(msg)
        self.dofs = dofs


class Sizes(ct.Structure):
    _fields_ = [(n, ct.c_int64) for n in
                ('n_dof', 'nnz_coo', 'nnz_csr', 'n1', 'n2', 'nnz11', 'nnz12', 'nnz21', 'nnz22')]


class SolveOpts(ct.Structure):
    _fields_ = [('solver', ct.c_int32), ('check_stability', ct.c_int32), ('max_iter', ct.c_int32),
                ('refine_steps', ct.c_int32), ('tol', ct.c_double), ('singular_tol', ct.c_double)]


class Stats(ct.Structure):
    _fields_ = [('ms_elements', ct.c_double), ('ms_symbolic', ct.c_double), ('ms_scatter', ct.c_double),
                ('ms_partition', ct.c_double), ('ms_rhs', ct.c_double), ('ms_factor', ct.c_double),
                ('ms_solve', ct.c_double), ('ms_reactions', ct.c_double), ('ms_spmm', ct.c_double),
                ('spmm_launches', ct.c_int64), ('kernel_launches', ct.c_int64), ('iterations', ct.c_int64),
                ('max_rel_residual', ct.c_double), ('factor_nnz', ct.c_int64), ('factor_flops', ct.c_double),
                ('update_flops', ct.c_double), ('solve_flops', ct.c_double)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


_lib = None

_P = ct.c_void_p
_I32P = ct.POINTER(ct.c_int32)
_F64P = ct.POINTER(ct.c_double)
_U8P = ct.POINTER(ct.c_uint8)
_U16P = ct.POINTER(ct.c_uint16)

# name -> (restype, argtypes); every symbol include/pynite_b200.h declares
SIGNATURES = {
    'pn_create': (ct.c_int, [ct.POINTER(_P), ct.c_int]),
    'pn_destroy': (None, [_P]),
    'pn_last_error': (ct.c_char_p, [_P]),
    'pn_version': (ct.c_char_p, []),
    'pn_set_nodes': (ct.c_int, [_P, ct.c_int64, _F64P, _U8P, _U8P, _F64P, _F64P, _U8P]),
    'pn_set_springs': (ct.c_int, [_P, ct.c_int64, _I32P, _F64P, _U8P]),
    'pn_set_members': (ct.c_int, [_P, ct.c_int64, _I32P, _F64P, _F64P, _U16P, _U8P]),
    'pn_set_quads': (ct.c_int, [_P, ct.c_int64, _I32P, _F64P]),
    'pn_set_plates': (ct.c_int, [_P, ct.c_int64, _I32P, _F64P]),
    'pn_element_ke': (ct.c_int, [_P, ct.c_int, _F64P]),
    'pn_member_kg': (ct.c_int, [_P, _F64P, _F64P]),
    'pn_symbolic': (ct.c_int, [_P, ct.c_uint32]),
    'pn_get_sizes': (ct.c_int, [_P, ct.POINTER(Sizes)]),
    'pn_get_coo': (ct.c_int, [_P, _I32P, _I32P, _F64P]),
    'pn_get_pattern': (ct.c_int, [_P, _I32P, _I32P]),
    'pn_get_partition': (ct.c_int, [_P, _I32P, _I32P, _F64P]),
    'pn_get_block': (ct.c_int, [_P, ct.c_int, _I32P, _I32P, _F64P]),
    'pn_assemble_ke': (ct.c_int, [_P]),
    'pn_get_csr_values': (ct.c_int, [_P, _F64P]),
    'pn_check_nodal_stability': (ct.c_int, [_P, _I32P, ct.c_int64, ct.POINTER(ct.c_int64)]),
    'pn_set_loads': (ct.c_int, [_P, ct.c_int32,
                                ct.c_int64, _I32P, _I32P, _F64P,
                                ct.c_int64, _I32P, _I32P, _I32P, _F64P,
                                ct.c_int64, _I32P, _I32P, _F64P,
                                ct.c_int64, _I32P, _I32P, _F64P]),
    'pn_set_combos': (ct.c_int, [_P, ct.c_int32, _F64P]),
    'pn_get_load_vector': (ct.c_int, [_P, ct.c_int32, ct.c_int32, _F64P]),
    'pn_solve_linear': (ct.c_int, [_P, ct.POINTER(SolveOpts), _F64P, _F64P, ct.POINTER(Stats)]),
    'pn_solve_pdelta': (ct.c_int, [_P, ct.POINTER(SolveOpts), _F64P, _F64P, ct.POINTER(Stats)]),
    'pn_get_kg_coo': (ct.c_int, [_P, ct.c_int32, ct.c_int32, _I32P, _I32P, _F64P, ct.POINTER(ct.c_int64)]),
    'pn_set_displacements': (ct.c_int, [_P, ct.c_int32, _F64P]),
    'pn_reactions': (ct.c_int, [_P, ct.c_int32, ct.c_int32, _F64P]),
    'pn_element_forces': (ct.c_int, [_P, ct.c_int, ct.c_int32, ct.c_int32, ct.c_int32, _F64P]),
    'pn_set_tc_modes': (ct.c_int, [_P, _U8P, _U8P, ct.c_int64, ct.POINTER(ct.c_int64)]),
    'pn_tc_update': (ct.c_int, [_P, ct.c_int32, ct.c_int32, ct.c_double, ct.c_double, ct.POINTER(ct.c_int64)]),
    'pn_tc_reset': (ct.c_int, [_P]),
    'pn_get_activity': (ct.c_int, [_P, _U8P, _U8P, _U8P]),
    'pn_set_activity': (ct.c_int, [_P, _U8P, _U8P, _U8P]),
    'pn_shell_stresses': (ct.c_int, [_P, ct.c_int, ct.c_int32, ct.c_int32, ct.c_int32, _F64P, ct.c_int32, _F64P]),
    'pn_set_distributed': (ct.c_int, [_P, ct.c_int32, ct.c_int32, ct.c_void_p, ct.c_void_p]),
    'pn_set_stream': (ct.c_int, [_P, ct.c_void_p, ct.c_int]),
    'pn_node_row_starts': (ct.c_int, [_P, ct.POINTER(ct.c_int64)]),
    'pn_prefetch_analysis': (ct.c_int, [_P]),
    'pn_k11_col_range': (ct.c_int, [_P, ct.c_int64, ct.c_int64, ct.POINTER(ct.c_int64), ct.POINTER(ct.c_int64)]),
    'pn_set_option': (ct.c_int, [_P, ct.c_char_p, ct.c_int64]),
    'pn_set_output_combos': (ct.c_int, [_P, ct.c_int32, ct.c_int32]),
    'pn_dev_spmm_k11_rows': (ct.c_int, [_P, ct.c_int32, ct.c_int64, ct.c_int64, ct.c_void_p, ct.c_void_p]),
    'pn_dev_block_jacobi_setup': (ct.c_int, [_P]),
    'pn_dev_block_jacobi_nodes': (ct.c_int, [_P, ct.c_int32, ct.c_int64, ct.c_int64, ct.c_void_p, ct.c_void_p]),
    'pn_dev_cg_dots': (ct.c_int, [_P, ct.c_int32, ct.c_int64, ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_void_p]),
    'pn_dev_cg_update': (ct.c_int, [_P, ct.c_int32, ct.c_int64, ct.c_int64, ct.c_int32] + [ct.c_void_p]*8),
    'pn_dev_cg_direction': (ct.c_int, [_P, ct.c_int32, ct.c_int64, ct.c_int64, ct.c_int32] + [ct.c_void_p]*4),
    'pn_dev_get_rhs': (ct.c_int, [_P, ct.c_int32, ct.c_int32, ct.c_void_p]),
    'pn_dev_set_solution': (ct.c_int, [_P, ct.c_int32, ct.c_int32, ct.c_void_p]),
    'pn_get_displacements': (ct.c_int, [_P, ct.c_int32, _F64P]),
    'pn_spmm_k11': (ct.c_int, [_P, ct.c_int32, _F64P, _F64P, ct.c_int32, _F64P]),
    'pn_get_stats': (ct.c_int, [_P, ct.POINTER(Stats)]),
    'pn_reset_stats': (ct.c_int, [_P]),
    'pn_get_counter': (ct.c_int, [_P, ct.c_char_p, ct.POINTER(ct.c_int64)]),
    'pn_counter_name': (ct.c_char_p, [ct.c_int]),
    'pn_timer_start': (ct.c_int, [_P]),
    'pn_timer_stop': (ct.c_int, [_P, _F64P]),
    'pn_profile_enable': (ct.c_int, [_P, ct.c_int]),
    'pn_profile_get': (ct.c_int, [_P, ct.c_int, _F64P, ct.POINTER(ct.c_int64)]),
    'pn_profile_slot_name': (ct.c_char_p, [ct.c_int]),
}


def library_path():
    return _LIB_PATH


def load_library():
    """Loads the shared library and declares every prototype.  Raises `EngineUnavailable` if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise EngineUnavailable(
            f'{_LIB_NAME} has not been built ({_LIB_PATH} is missing). Run `python -c "import __graft_entry__ as g; '
            'g.build()"` or `make -C pynite_b200/csrc`. pynite_b200 has no CPU solver.')
    lib = ct.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a, typ):
    if a is None:
        return None
    return a.ctypes.data_as(typ)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Engine:
    """One device context (`pn_ctx`) bound to a GPU."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = _P()
        rc = self._lib.pn_create(ct.byref(h), int(device))
        if rc == PN_ERR_NO_DEVICE:
            raise EngineUnavailable('no CUDA device is visible; pynite_b200 has no CPU solver')
        if rc != PN_OK:
            raise EngineUnavailable(f'pn_create failed with code {rc}')
        self._h = h
        self.device = device
        self._keep = []          # arrays whose memory must outlive a call
        # one process per GPU (torchrun): the ranks of a box share its cores, and torchrun pins OMP_NUM_THREADS to 1, which
        # would leave the ordering of the direct solver single-threaded.  Give every rank its share, less the caller's thread.
        local_world = int(os.environ.get('LOCAL_WORLD_SIZE', '1') or 1)
        if local_world > 1:
            self.set_option('host_threads', max(1, (os.cpu_count() or 1)//local_world - 1))

    def close(self):
        if getattr(self, '_h', None):
            self._lib.pn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- error mapping ------------------------------------------------------------------------
    def _check(self, rc, bad_dofs=None):
        if rc == PN_OK:
            return
        msg = (self._lib.pn_last_error(self._h) or b'').decode()
        if rc == PN_ERR_SINGULAR:
            raise SingularMatrix(SINGULAR_MSG)
        if rc == PN_ERR_UNSTABLE_NODE:
            raise UnstableNodes('Unstable node(s). See console output for details.', bad_dofs)
        if rc == PN_ERR_NOT_CONVERGED:
            raise NotConverged(msg)
        if rc == PN_ERR_ARG:
            raise ValueError(msg)
        raise RuntimeError(f'pynite_b200 native error {rc}: {msg}')

    # ---- model upload -------------------------------------------------------------------------
    def set_model(self, A):
        """Uploads a `flatten.ModelArrays`."""
        L = self._lib
        self._check(L.pn_set_nodes(self._h, A.n_nodes, _ptr(_c(A.xyz, np.float64), _F64P),
                                   _ptr(_c(A.support, np.uint8), _U8P), _ptr(_c(A.enf_mask, np.uint8), _U8P),
                                   _ptr(_c(A.enf_val, np.float64), _F64P), _ptr(_c(A.nspring_k, np.float64), _F64P),
                                   _ptr(_c(A.nspring_flag, np.uint8), _U8P)))
        self.set_springs(A)
        self.set_members(A)
        self._check(L.pn_set_quads(self._h, len(A.quad_ijmn), _ptr(_c(A.quad_ijmn, np.int32), _I32P),
                                   _ptr(_c(A.quad_props, np.float64), _F64P)))
        self._check(L.pn_set_plates(self._h, len(A.plate_ijmn), _ptr(_c(A.plate_ijmn, np.int32), _I32P),
                                    _ptr(_c(A.plate_props, np.float64), _F64P)))
        self.n_dof = 6*A.n_nodes
        self.counts = {PN_SPRING: len(A.spring_ks), PN_MEMBER: len(A.mem_rot),
                       PN_QUAD: len(A.quad_ijmn), PN_PLATE: len(A.plate_ijmn)}
        # the model is complete: the ordering of the direct solver can start on a host worker thread
        self._check(L.pn_prefetch_analysis(self._h))

    def k11_col_range(self, row0, row1):
        lo, hi = ct.c_int64(0), ct.c_int64(0)
        self._check(self._lib.pn_k11_col_range(self._h, row0, row1, ct.byref(lo), ct.byref(hi)))
        return lo.value, hi.value

    def set_output_combos(self, combo0, n_combo):
        """`solve_linear` / `solve_pdelta` deliver combinations [combo0, combo0 + n_combo) to the host arrays (n_combo < 0:
        all).  A rank of a multi-GPU solve downloads its share only."""
        self._check(self._lib.pn_set_output_combos(self._h, int(combo0), int(n_combo)))
        self._out_range = None if n_combo < 0 else (int(combo0), int(n_combo))

    def set_option(self, name, value):
        self._check(self._lib.pn_set_option(self._h, name.encode(), int(value)))

    def set_nodes(self, A):
        self._check(self._lib.pn_set_nodes(self._h, A.n_nodes, _ptr(_c(A.xyz, np.float64), _F64P),
                                           _ptr(_c(A.support, np.uint8), _U8P), _ptr(_c(A.enf_mask, np.uint8), _U8P),
                                           _ptr(_c(A.enf_val, np.float64), _F64P),
                                           _ptr(_c(A.nspring_k, np.float64), _F64P),
                                           _ptr(_c(A.nspring_flag, np.uint8), _U8P)))

    def set_springs(self, A):
        self._check(self._lib.pn_set_springs(self._h, len(A.spring_ks), _ptr(_c(A.spring_ij, np.int32), _I32P),
                                             _ptr(_c(A.spring_ks, np.float64), _F64P),
                                             _ptr(_c(A.spring_active, np.uint8), _U8P)))

    def set_members(self, A):
        self._check(self._lib.pn_set_members(self._h, len(A.mem_rot), _ptr(_c(A.mem_ij, np.int32), _I32P),
                                             _ptr(_c(A.mem_props, np.float64), _F64P),
                                             _ptr(_c(A.mem_rot, np.float64), _F64P),
                                             _ptr(_c(A.mem_rel, np.uint16), _U16P),
                                             _ptr(_c(A.mem_active, np.uint8), _U8P)))

    def set_loads(self, A):
        n_case = len(A.cases)
        a = [_c(A.nl_dof, np.int32), _c(A.nl_case, np.int32), _c(A.nl_val, np.float64),
             _c(A.ml_member, np.int32), _c(A.ml_case, np.int32), _c(A.ml_kind, np.int32), _c(A.ml_params, np.float64),
             _c(A.qp_elem, np.int32), _c(A.qp_case, np.int32), _c(A.qp_val, np.float64),
             _c(A.pp_elem, np.int32), _c(A.pp_case, np.int32), _c(A.pp_val, np.float64)]
        self._check(self._lib.pn_set_loads(
            self._h, n_case,
            len(a[2]), _ptr(a[0], _I32P), _ptr(a[1], _I32P), _ptr(a[2], _F64P),
            len(a[5]), _ptr(a[3], _I32P), _ptr(a[4], _I32P), _ptr(a[5], _I32P), _ptr(a[6], _F64P),
            len(a[9]), _ptr(a[7], _I32P), _ptr(a[8], _I32P), _ptr(a[9], _F64P),
            len(a[12]), _ptr(a[10], _I32P), _ptr(a[11], _I32P), _ptr(a[12], _F64P)))
        self.n_case = n_case

    def set_combos(self, factors):
        f = _c(factors, np.float64).reshape(-1, max(self.n_case, 1)) if self.n_case else np.zeros((len(factors), 0))
        self.n_combo = f.shape[0]
        self._check(self._lib.pn_set_combos(self._h, self.n_combo, _ptr(_c(f, np.float64), _F64P)))

    # ---- symbolic / assembly ------------------------------------------------------------------
    def symbolic(self, flags=0):
        self._check(self._lib.pn_symbolic(self._h, flags))
        return self.sizes()

    def sizes(self):
        s = Sizes()
        self._check(self._lib.pn_get_sizes(self._h, ct.byref(s)))
        return s

    def element_ke(self, kind):
        kind = KIND.get(kind, kind)
        nd = _ND[kind]
        out = np.empty((self.counts[kind], nd, nd))
        self._check(self._lib.pn_element_ke(self._h, kind, _ptr(out, _F64P)))
        return out

    def member_kg(self, P):
        P = _c(P, np.float64)
        out = np.empty((self.counts[PN_MEMBER], 12, 12))
        self._check(self._lib.pn_member_kg(self._h, _ptr(P, _F64P), _ptr(out, _F64P)))
        return out

    def coo(self):
        s = self.sizes()
        row = np.empty(s.nnz_coo, dtype=np.int32)
        col = np.empty(s.nnz_coo, dtype=np.int32)
        data = np.empty(s.nnz_coo)
        self._check(self._lib.pn_get_coo(self._h, _ptr(row, _I32P), _ptr(col, _I32P), _ptr(data, _F64P)))
        return row, col, data

    def pattern(self):
        s = self.sizes()
        indptr = np.empty(s.n_dof + 1, dtype=np.int32)
        indices = np.empty(s.nnz_csr, dtype=np.int32)
        self._check(self._lib.pn_get_pattern(self._h, _ptr(indptr, _I32P), _ptr(indices, _I32P)))
        return indptr, indices

    def partition(self):
        s = self.sizes()
        d1 = np.empty(s.n1, dtype=np.int32)
        d2 = np.empty(s.n2, dtype=np.int32)
        v2 = np.empty(s.n2)
        self._check(self._lib.pn_get_partition(self._h, _ptr(d1, _I32P), _ptr(d2, _I32P), _ptr(v2, _F64P)))
        return d1, d2, v2

    def block(self, b, values=True):
        s = self.sizes()
        rows = (s.n1, s.n1, s.n2, s.n2)[b]
        nnz = (s.nnz11, s.nnz12, s.nnz21, s.nnz22)[b]
        indptr = np.empty(rows + 1, dtype=np.int32)
        indices = np.empty(nnz, dtype=np.int32)
        data = np.empty(nnz) if values else None
        self._check(self._lib.pn_get_block(self._h, b, _ptr(indptr, _I32P), _ptr(indices, _I32P), _ptr(data, _F64P)))
        return indptr, indices, data

    def assemble_ke(self):
        self._check(self._lib.pn_assemble_ke(self._h))

    def csr_values(self):
        out = np.empty(self.sizes().nnz_csr)
        self._check(self._lib.pn_get_csr_values(self._h, _ptr(out, _F64P)))
        return out

    def check_nodal_stability(self, cap=4096):
        bad = np.empty(cap, dtype=np.int32)
        n = ct.c_int64(0)
        rc = self._lib.pn_check_nodal_stability(self._h, _ptr(bad, _I32P), cap, ct.byref(n))
        self._check(rc, bad[:min(n.value, cap)].copy())

    def load_vector(self, combo, which):
        out = np.empty((self.n_dof, 1))
        self._check(self._lib.pn_get_load_vector(self._h, combo, {'P': 0, 'FER': 1}[which], _ptr(out, _F64P)))
        return out

    # ---- solves -------------------------------------------------------------------------------
    @staticmethod
    def make_opts(solver='direct', check_stability=True, max_iter=0, refine_steps=2, tol=1e-12, singular_tol=1e-6):
        o = SolveOpts()
        o.solver = PN_SOLVER_PCG if solver == 'pcg' else PN_SOLVER_DIRECT
        o.check_stability = 1 if check_stability else 0
        o.max_iter = max_iter
        o.refine_steps = refine_steps
        o.tol = tol
        o.singular_tol = singular_tol
        return o

    def _solve(self, fn, opts, want_reactions, D_out=None, R_out=None, download=True):
        n = self.n_dof
        rng = getattr(self, '_out_range', None)
        rows = self.n_combo if rng is None else max(0, min(rng[1], self.n_combo - rng[0]))
        D = (D_out if D_out is not None else np.empty((rows, n))) if download else None
        R = (R_out if R_out is not None else np.empty((rows, n))) if want_reactions else None
        for a in (D, R):
            if a is not None and a.size < rows*n:
                raise ValueError('solve: host array smaller than the output range (set_output_combos)')
        st = Stats()
        rc = fn(self._h, ct.byref(opts) if opts is not None else None, _ptr(D, _F64P), _ptr(R, _F64P), ct.byref(st))
        self._check(rc)
        return D, R, st

    def solve_linear(self, opts=None, reactions=True, D_out=None, R_out=None, download=True):
        """download=False keeps the displacements on the device (no D2H copy); reactions imply a copy."""
        return self._solve(self._lib.pn_solve_linear, opts, reactions, D_out, R_out, download)

    def solve_pdelta(self, opts=None, reactions=True, D_out=None, R_out=None, download=True):
        return self._solve(self._lib.pn_solve_pdelta, opts, reactions, D_out, R_out, download)

    def kg_coo(self, combo, first_step):
        cap = 144*self.counts[PN_MEMBER]
        row = np.empty(cap, dtype=np.int32)
        col = np.empty(cap, dtype=np.int32)
        data = np.empty(cap)
        n = ct.c_int64(0)
        self._check(self._lib.pn_get_kg_coo(self._h, combo, 1 if first_step else 0, _ptr(row, _I32P), _ptr(col, _I32P),
                                            _ptr(data, _F64P), ct.byref(n)))
        return row[:n.value], col[:n.value], data[:n.value]

    def set_displacements(self, combo, D):
        D = _c(D, np.float64).reshape(-1)
        self._check(self._lib.pn_set_displacements(self._h, combo, _ptr(D, _F64P)))

    def reactions(self, combo, pdelta=False, out=None):
        out = np.empty((self.n_dof, 1)) if out is None else out
        self._check(self._lib.pn_reactions(self._h, combo, 1 if pdelta else 0, _ptr(out, _F64P)))
        return out

    def element_forces(self, kind, combo, local=False, pdelta=False):
        kind = KIND.get(kind, kind)
        out = np.empty((self.counts[kind], _ND[kind]))
        self._check(self._lib.pn_element_forces(self._h, kind, combo, 1 if local else 0, 1 if pdelta else 0,
                                                _ptr(out, _F64P)))
        return out

    # ---- tension/compression-only iteration on the device ----------------------------------------------
    def set_tc_modes(self, spring_mode, member_mode, group_start):
        sm, mm = _c(spring_mode, np.uint8), _c(member_mode, np.uint8)
        gs = _c(group_start, np.int64)
        self._check(self._lib.pn_set_tc_modes(self._h, _ptr(sm, _U8P), _ptr(mm, _U8P), max(len(gs) - 1, 0),
                                              gs.ctypes.data_as(ct.POINTER(ct.c_int64))))

    def tc_update(self, combo, pdelta, spring_tol, member_tol):
        """Number of activity changes `Analysis._check_TC_convergence` would make for the resident combo (0 = converged)."""
        n = ct.c_int64(0)
        self._check(self._lib.pn_tc_update(self._h, combo, 1 if pdelta else 0, float(spring_tol), float(member_tol), ct.byref(n)))
        return n.value

    def tc_reset(self):
        self._check(self._lib.pn_tc_reset(self._h))

    def activity(self, n_nodes):
        nsf = np.empty((n_nodes, 6), dtype=np.uint8)
        sa = np.empty(self.counts[PN_SPRING], dtype=np.uint8)
        ma = np.empty(self.counts[PN_MEMBER], dtype=np.uint8)
        self._check(self._lib.pn_get_activity(self._h, _ptr(nsf, _U8P), _ptr(sa, _U8P), _ptr(ma, _U8P)))
        return nsf, sa, ma

    def set_activity(self, nspring_flag=None, spring_active=None, member_active=None):
        a = [None if x is None else _c(x, np.uint8) for x in (nspring_flag, spring_active, member_active)]
        self._check(self._lib.pn_set_activity(self._h, _ptr(a[0], _U8P), _ptr(a[1], _U8P), _ptr(a[2], _U8P)))

    def shell_stresses(self, kind, combo0, n_combo, pts, local=True):
        """[n_combo][count][n_pts][9] internal shears / moments / membrane stresses of every quad or plate."""
        kind = KIND.get(kind, kind)
        pts = _c(pts, np.float64).reshape(-1, 2)
        out = np.empty((n_combo, self.counts[kind], len(pts), 9))
        self._check(self._lib.pn_shell_stresses(self._h, kind, combo0, n_combo, len(pts), _ptr(pts, _F64P), 1 if local else 0,
                                                _ptr(out, _F64P)))
        return out

    def spmm_k11(self, X, repeat=1):
        X = _c(X, np.float64)
        Y = np.empty_like(X)
        ms = ct.c_double(0)
        self._check(self._lib.pn_spmm_k11(self._h, X.shape[1], _ptr(X, _F64P), _ptr(Y, _F64P), repeat, ct.byref(ms)))
        return Y, ms.value

    def spmm_k11_time(self, s, repeat=10):
        ms = ct.c_double(0)
        self._check(self._lib.pn_spmm_k11(self._h, s, None, None, repeat, ct.byref(ms)))
        return ms.value

    # ---- device-pointer building blocks (domain-decomposed PCG, pynite_b200/ddpcg.py) -----------------------
    def set_stream(self, cuda_stream):
        """Launch on the caller's CUDA stream (integer handle, 0 = legacy default stream); None = the library's own."""
        if cuda_stream is None:
            self._check(self._lib.pn_set_stream(self._h, None, 1))
        else:
            self._check(self._lib.pn_set_stream(self._h, ct.c_void_p(cuda_stream), 0))

    def node_row_starts(self, n_nodes):
        out = np.empty(n_nodes + 1, dtype=np.int64)
        self._check(self._lib.pn_node_row_starts(self._h, out.ctypes.data_as(ct.POINTER(ct.c_int64))))
        return out

    def dev_spmm_rows(self, s, row0, row1, x_ptr, y_ptr):
        self._check(self._lib.pn_dev_spmm_k11_rows(self._h, s, row0, row1, ct.c_void_p(x_ptr), ct.c_void_p(y_ptr)))

    def dev_block_jacobi_setup(self):
        self._check(self._lib.pn_dev_block_jacobi_setup(self._h))

    def dev_block_jacobi_nodes(self, s, node0, node1, r_ptr, z_ptr):
        self._check(self._lib.pn_dev_block_jacobi_nodes(self._h, s, node0, node1, ct.c_void_p(r_ptr), ct.c_void_p(z_ptr)))

    def dev_cg_dots(self, s, row0, row1, a_ptr, b_ptr, out_ptr):
        self._check(self._lib.pn_dev_cg_dots(self._h, s, row0, row1, ct.c_void_p(a_ptr), ct.c_void_p(b_ptr), ct.c_void_p(out_ptr)))

    def dev_cg_update(self, s, node0, node1, first, rz, pAp, P, AP, X, R, Z, out2):
        self._check(self._lib.pn_dev_cg_update(self._h, s, node0, node1, 1 if first else 0, *[ct.c_void_p(p) for p in (rz, pAp, P, AP, X, R, Z, out2)]))

    def dev_cg_direction(self, s, row0, row1, first, rz, rz_new, Z, P):
        self._check(self._lib.pn_dev_cg_direction(self._h, s, row0, row1, 1 if first else 0, *[ct.c_void_p(p) for p in (rz, rz_new, Z, P)]))

    def dev_get_rhs(self, combo0, s, b_ptr):
        self._check(self._lib.pn_dev_get_rhs(self._h, combo0, s, ct.c_void_p(b_ptr)))

    def dev_set_solution(self, combo0, s, x_ptr):
        self._check(self._lib.pn_dev_set_solution(self._h, combo0, s, ct.c_void_p(x_ptr)))

    def displacements(self, combo, out=None):
        out = np.empty(self.n_dof) if out is None else out
        self._check(self._lib.pn_get_displacements(self._h, combo, _ptr(out, _F64P)))
        return out

    def stats(self):
        st = Stats()
        self._check(self._lib.pn_get_stats(self._h, ct.byref(st)))
        return st

    def timer_start(self):
        self._check(self._lib.pn_timer_start(self._h))

    def timer_stop(self):
        ms = ct.c_double(0)
        self._check(self._lib.pn_timer_stop(self._h, ct.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        self._check(self._lib.pn_profile_enable(self._h, 1 if on else 0))

    def profile(self):
        """{kernel slot: (summed device ms, launches)} since profile_enable(True)."""
        out = {}
        slot = 0
        while True:
            name = self._lib.pn_profile_slot_name(slot)
            if not name:
                break
            ms, n = ct.c_double(0), ct.c_int64(0)
            self._check(self._lib.pn_profile_get(self._h, slot, ct.byref(ms), ct.byref(n)))
            if n.value:
                out[name.decode()] = (ms.value, n.value)
            slot += 1
        return out

    def reset_stats(self):
        self._check(self._lib.pn_reset_stats(self._h))

    def counters(self):
        """{name: count} of the direct solver's kernel families since `reset_stats` (`pn_get_counter`)."""
        out = {}
        i = 0
        while True:
            name = self._lib.pn_counter_name(i)
            if not name:
                break
            v = ct.c_int64(0)
            self._check(self._lib.pn_get_counter(self._h, name, ct.byref(v)))
            out[name.decode()] = v.value
            i += 1
        return out
