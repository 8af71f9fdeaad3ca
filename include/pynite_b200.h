/*
 * pynite_b200.h -- C ABI of the B200-native stiffness-assembly + linear-solve library.
 *
 * The reference (JWock82/Pynite) is pure Python and has no FFI; its "operator interface" for this
 * path is the set of Python functions listed beside each entry point below (file:line relative to
 * /root/reference/Pynite).  A maintainer who wanted the reference itself to use this library would
 * bind these symbols with ctypes exactly as pynite_b200/engine.py does.
 *
 * Conventions
 *   - every pointer is a HOST pointer to caller-owned memory; the library copies in/out;
 *   - integers are int32 unless stated, reals are IEEE binary64;
 *   - a global DOF is node_id*6 + k, k = 0..5 for DX DY DZ RX RY RZ (FEModel3D.py:66-94);
 *   - return value 0 = success, negative = pn_status error; pn_last_error() gives the text;
 *   - one pn_ctx per model and per GPU; calls on one ctx are not re-entrant.
 *   - there is no CPU implementation behind any entry point: without a CUDA device pn_create fails.
 */
#ifndef PYNITE_B200_H
#define PYNITE_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define PN_EXPORT __attribute__((visibility("default")))
#else
#define PN_EXPORT
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pn_ctx pn_ctx;

enum pn_status {
    PN_OK = 0,
    PN_ERR_CUDA = -1,          /* CUDA runtime error (text in pn_last_error)                       */
    PN_ERR_ARG = -2,           /* bad argument / call order                                        */
    PN_ERR_SINGULAR = -3,      /* Analysis.py:172,239  "stiffness matrix is singular"              */
    PN_ERR_UNSTABLE_NODE = -4, /* Analysis.py:111-166  zero diagonal on an unsupported DOF         */
    PN_ERR_NOT_CONVERGED = -5, /* iterative solver hit max_iter before reaching tol                */
    PN_ERR_NO_DEVICE = -6
};

enum pn_elem_kind { PN_SPRING = 1, PN_MEMBER = 2, PN_QUAD = 3, PN_PLATE = 4 };

/* flags for pn_symbolic */
#define PN_SYM_DENSE_MEMBER_BLOCKS 1u /* keep all 144 entries of every member block (Kg has no zero
                                          filter: FEModel3D.py:1859-1885); used by the P-Delta solve */

/* solver selection for pn_solve_* */
#define PN_SOLVER_DIRECT 0 /* sparse multifrontal LU (no pivoting) + double-double refinement      */
#define PN_SOLVER_PCG 1    /* node-block-Jacobi preconditioned CG, all right-hand sides at once    */

typedef struct pn_sizes {
    int64_t n_dof;    /* 6 * nodes                                                                 */
    int64_t nnz_coo;  /* COO entries the reference would emit (FEModel3D.py:1773-1791)             */
    int64_t nnz_csr;  /* stored entries after duplicate summation (.tocsr(), FEModel3D.py:2279)    */
    int64_t n1, n2;   /* unknown / known DOF counts (Analysis.py:1412-1505)                        */
    int64_t nnz11, nnz12, nnz21, nnz22;
} pn_sizes;

typedef struct pn_solve_opts {
    int32_t solver;          /* PN_SOLVER_*                                                        */
    int32_t check_stability; /* residual / finiteness check of Analysis.py:228-239                 */
    int32_t max_iter;        /* PCG iteration cap (0 = library default)                            */
    int32_t refine_steps;    /* iterative-refinement sweeps after the direct solve (default 2)     */
    double tol;              /* PCG: relative residual target per right-hand side (default 1e-12)  */
    double singular_tol;     /* Analysis.py:176  tol = 1e-6 on ||K x - b|| / ||b||                 */
} pn_solve_opts;

typedef struct pn_stats {
    double ms_elements;   /* element stiffness kernels                                             */
    double ms_symbolic;   /* pattern + scatter map + partition maps                                */
    double ms_scatter;    /* COO -> CSR scatter-add                                                */
    double ms_partition;  /* K11 / K12 / K21 / K22 value gather                                    */
    double ms_rhs;        /* load vectors and K12 * D2                                             */
    double ms_factor;     /* direct: numeric factorisation; PCG: preconditioner setup              */
    double ms_solve;      /* triangular solves + refinement, or PCG iterations                     */
    double ms_reactions;  /* element end forces and reactions                                      */
    double ms_spmm;       /* time inside the SpMM kernel only (summed over launches)               */
    int64_t spmm_launches;
    int64_t kernel_launches; /* kernels of this library launched since pn_create                   */
    int64_t iterations;      /* PCG iterations (max over right-hand sides) / refinement sweeps     */
    double max_rel_residual; /* max over combos of ||K11 x - b|| / ||b||                            */
    int64_t factor_nnz;      /* stored entries of the direct factor                                */
    double factor_flops;
    double update_flops;     /* flops of the trailing-matrix updates alone (2 bs rest^2 per block step)    */
    double solve_flops;      /* flops of one forward+backward sweep for ONE right-hand side               */
} pn_stats;

/* ---- life cycle ------------------------------------------------------------------------------ */
PN_EXPORT int pn_create(pn_ctx **out, int device_id);
PN_EXPORT void pn_destroy(pn_ctx *ctx);
PN_EXPORT const char *pn_last_error(pn_ctx *ctx);
PN_EXPORT const char *pn_version(void);

/* ---- model upload: replaces the object reads inside FEModel3D.Ke (FEModel3D.py:1587-1771) ------
 * nodes:   Node3D.py:28-85.  support_mask bit k = DOF k supported; enforced_mask likewise;
 *          spring_flag[n*6+k]: bit0 = spring defined, bit1 = active (FEModel3D.py:1590-1593).
 * springs: Spring3D.py:25-43.  members: Member3D.py:37-105 (sub-members after descritize,
 *          PhysMember.py:37-200); props = E, G, A, Iy, Iz, J; releases bit k = local DOF k released.
 * quads / plates: Quad3D.py:30-57, Plate3D.py:16-72; props = t, E, nu, kx_mod, ky_mod.            */
PN_EXPORT int pn_set_nodes(pn_ctx *ctx, int64_t n_nodes, const double *xyz, const uint8_t *support_mask,
                 const uint8_t *enforced_mask, const double *enforced_val, const double *spring_k,
                 const uint8_t *spring_flag);
PN_EXPORT int pn_set_springs(pn_ctx *ctx, int64_t n, const int32_t *ij, const double *ks, const uint8_t *active);
PN_EXPORT int pn_set_members(pn_ctx *ctx, int64_t n, const int32_t *ij, const double *props, const double *rotation_deg,
                   const uint16_t *releases, const uint8_t *active);
PN_EXPORT int pn_set_quads(pn_ctx *ctx, int64_t n, const int32_t *ijmn, const double *props);
PN_EXPORT int pn_set_plates(pn_ctx *ctx, int64_t n, const int32_t *ijmn, const double *props);

/* ---- per-element global matrices, for inspection and parity tests -----------------------------
 * Member3D.Ke :1038, Spring3D.Ke :191, Quad3D.Ke :858, Plate3D.Ke :502 -> out[count][nd*nd], nd = 12 or 24.
 * Member3D.Kg :1048 with axial force P[m] -> out[count][144].                                     */
PN_EXPORT int pn_element_ke(pn_ctx *ctx, int kind, double *out);
PN_EXPORT int pn_member_kg(pn_ctx *ctx, const double *P, double *out);

/* ---- symbolic phase: FEModel3D._build_dof_vector :66, _append_sparse_block :99 (zero filter),
 *      coo_matrix + tocsr pattern (:1791, :2279), Analysis._partition_D :1412, _partition :1508 -- */
/* Tuning switches (no effect on results).  "assembly_graph" (default 1): the numeric class/recipe assembly is replayed
 * as one CUDA graph; 0 issues its launches one by one so that pn_profile_enable can time them individually.
 * "spmm_grouped" (default 0): K11 SpMM (pn_spmm_k11, PCG) with the rows of one node sharing every X-row load; 0 = one
 * lane group per row (same bits; the grouped kernel measured slower and is kept for A/B runs).
 * "dof_classes" (default 1): ordering of the direct solver.  The reference stores no exact zeros (FEModel3D.py:130),
 * so K11 of a flat shell model in an axis plane is reducible -- membrane translations, the bending triple and the drilling
 * rotation never meet in a stored entry -- and SuperLU's column ordering inside spsolve (Analysis.py:217) works on
 * that scalar graph.  1 = the classes are predicted from the element list, every separator of the node dissection
 * yields one front per class, and the prediction is verified against the assembled pattern (rejected -> node-based
 * ordering, counter "dof_class_rejects"); 0 = node-based ordering only; 2 = self-test (a prediction that must be
 * rejected).  Same solution up to rounding; counter "dof_classes" reports the number of classes in use.
 * "host_threads": OpenMP threads of the host worker (ordering, result zero fill). */
PN_EXPORT int pn_set_option(pn_ctx *ctx, const char *name, int64_t value);
/* Multi-GPU runs: every rank of the subtree-to-GPU solve (pn_set_distributed) ends with all displacements on its device,
 * but only one copy of the result has to reach host memory.  After this call D_out / Rxn_out of pn_solve_linear and
 * pn_solve_pdelta hold rows for combinations [combo0, combo0 + n_combo) only (arrays of n_combo x 6 N doubles);
 * n_combo < 0 restores "all combinations".  FEModel3D._D / node.Rxn* (Analysis.py:847, 1059) are per-combination
 * dictionaries, so a rank fills the entries of its share and the caller gathers them if it wants them everywhere. */
PN_EXPORT int pn_set_output_combos(pn_ctx *ctx, int32_t combo0, int32_t n_combo);
/* Optional hint, callable once every pn_set_nodes / pn_set_springs / pn_set_members / pn_set_quads / pn_set_plates call
 * of a model has been made: starts the host half of the direct solver's analysis (node graph, nested dissection,
 * fronts; the reference has no counterpart, SuperLU orders inside spsolve, Analysis.py:217) on a worker thread so
 * that it overlaps the load upload and the symbolic phase.  pn_symbolic starts it itself if this was not called; a
 * later pn_set_* call discards it. */
PN_EXPORT int pn_prefetch_analysis(pn_ctx *ctx);
PN_EXPORT int pn_symbolic(pn_ctx *ctx, uint32_t flags);
PN_EXPORT int pn_get_sizes(pn_ctx *ctx, pn_sizes *out);
PN_EXPORT int pn_get_coo(pn_ctx *ctx, int32_t *row, int32_t *col, double *data); /* emission order           */
PN_EXPORT int pn_get_pattern(pn_ctx *ctx, int32_t *indptr, int32_t *indices);    /* CSR of Ke, sorted        */
PN_EXPORT int pn_get_partition(pn_ctx *ctx, int32_t *d1, int32_t *d2, double *d2_values);
/* block: 0 = K11, 1 = K12, 2 = K21, 3 = K22; any output pointer may be NULL */
PN_EXPORT int pn_get_block(pn_ctx *ctx, int block, int32_t *indptr, int32_t *indices, double *data);

/* ---- numeric assembly: duplicate summation of .tocsr() (FEModel3D.py:2279) -------------------- */
PN_EXPORT int pn_assemble_ke(pn_ctx *ctx);
PN_EXPORT int pn_get_csr_values(pn_ctx *ctx, double *data);
/* Analysis._check_stability :111-168; on PN_ERR_UNSTABLE_NODE the first *n_bad (<= cap) DOFs are
 * written to bad_dofs                                                                             */
PN_EXPORT int pn_check_nodal_stability(pn_ctx *ctx, int32_t *bad_dofs, int64_t cap, int64_t *n_bad);

/* ---- loads: FEModel3D.P :2184, FEModel3D.FER :2136, Member3D.fer :742, Quad3D.fer :721,
 *      Plate3D.fer :324, FixedEndReactions.py ----------------------------------------------------
 * nodal loads: (dof, case, value) triplets in model order.
 * member loads: kind 0..5 = point Fx Fy Fz Mx My Mz (local), 6..11 = point FX FY FZ MX MY MZ
 *   (global), params = (P, x, -, -); 12..14 = distributed Fx Fy Fz, 15..17 = FX FY FZ,
 *   params = (w1, w2, x1, x2).  Entries must be grouped by member (ascending member index).
 * pressures: (element, case, value) triplets.                                                     */
PN_EXPORT int pn_set_loads(pn_ctx *ctx, int32_t n_case,
                 int64_t n_nodal, const int32_t *nl_dof, const int32_t *nl_case, const double *nl_val,
                 int64_t n_mload, const int32_t *ml_member, const int32_t *ml_case, const int32_t *ml_kind,
                 const double *ml_params,
                 int64_t n_qp, const int32_t *qp_elem, const int32_t *qp_case, const double *qp_val,
                 int64_t n_pp, const int32_t *pp_elem, const int32_t *pp_case, const double *pp_val);
/* LoadCombo.py:8-21: factors[n_combo][n_case] */
PN_EXPORT int pn_set_combos(pn_ctx *ctx, int32_t n_combo, const double *factors);
/* which: 0 = P, 1 = FER; out[6N] for one combo */
PN_EXPORT int pn_get_load_vector(pn_ctx *ctx, int32_t combo, int32_t which, double *out);

/* ---- solve: FEModel3D.analyze_linear :2287-2320 (all combos), Analysis._solve_unknown_disp :176,
 *      _store_displacements :847, _calc_reactions :1059 ------------------------------------------
 * D_out[n_combo][6N], Rxn_out[n_combo][6N] (either may be NULL).                                  */
PN_EXPORT int pn_solve_linear(pn_ctx *ctx, const pn_solve_opts *opts, double *D_out, double *Rxn_out, pn_stats *stats);
/* Analysis._PDelta :329-471 + FEModel3D.Kg :1802: solve, P = EA/L (d6 - d0), Kg, K = Ke + Kg, re-solve */
PN_EXPORT int pn_solve_pdelta(pn_ctx *ctx, const pn_solve_opts *opts, double *D_out, double *Rxn_out, pn_stats *stats);
/* global geometric stiffness for one solved combo as COO, 144 entries per active member, unfiltered
 * (FEModel3D.py:1826-1892); first_step != 0 uses P = 0.  row/col/data sized 144 * active members  */
PN_EXPORT int pn_get_kg_coo(pn_ctx *ctx, int32_t combo, int32_t first_step, int32_t *row, int32_t *col, double *data,
                  int64_t *n_out);

/* Analysis._sum_displacements :882-914 (load stepping in _first_order :262-305): overwrites the stored
 * displacement vector of a combo, so that element forces and reactions refer to the summed state.  Also how
 * displacements obtained elsewhere (another rank's share of the combinations) become resident: valid for any
 * combo index of the current pn_set_combos table once the matrix is assembled.                         */
PN_EXPORT int pn_set_displacements(pn_ctx *ctx, int32_t combo, const double *D);
/* Analysis._calc_reactions :1059-1313 for one solved combo with the current element activity; out[6N] */
PN_EXPORT int pn_reactions(pn_ctx *ctx, int32_t combo, int32_t pdelta, double *out);

/* ---- element end forces of a solved combo: Member3D.f :872 / F :1072, Spring3D.f :86 / F :200,
 *      Quad3D.f :713 / F :798, Plate3D.f :316 / F :403.  out[count][nd]; local != 0 -> local axes;
 *      pdelta != 0 uses the (ke + kg(P)) branch of Member3D.f :886-891                             */
PN_EXPORT int pn_element_forces(pn_ctx *ctx, int kind, int32_t combo, int32_t local, int32_t pdelta, double *out);

/* ---- tension/compression-only iteration on the device: Analysis._check_TC_convergence :917-1056 -----------------
 * pn_set_tc_modes: per two-node spring and per sub-member, bit0 = tension-only, bit1 = compression-only
 *   (Spring3D.py:25-43, Member3D.py:37-105); group_start[n_groups + 1] = sub-member range of every physical member
 *   (a physical member is deactivated as a whole, Analysis.py:1017-1047).  Node support springs carry their
 *   direction in spring_flag bits 2 ('+') and 3 ('-') of pn_set_nodes.
 * pn_tc_update: for the solved combo `combo` (displacements as stored / set by pn_set_displacements) switches node
 *   support springs with a direction on or off, deactivates springs and members whose axial force has the wrong sign
 *   (member extremes: Member3D.max_axial / min_axial), all in kernels; *n_changed = number of state changes, 0 =
 *   converged.  When anything changed the model's activity flags are updated in place (no re-upload) and the next
 *   pn_symbolic / pn_assemble_ke / pn_solve_* re-assembles with them.
 * pn_tc_reset: every spring and member active again (start of a new combination, Analysis.py:68-79); node support
 *   springs keep their state, as in the reference.  pn_get_activity: current flags (any pointer may be NULL).    */
PN_EXPORT int pn_set_tc_modes(pn_ctx *ctx, const uint8_t *spring_mode, const uint8_t *member_mode, int64_t n_groups,
                              const int64_t *group_start);
PN_EXPORT int pn_tc_update(pn_ctx *ctx, int32_t combo, int32_t pdelta, double spring_tol, double member_tol, int64_t *n_changed);
PN_EXPORT int pn_tc_reset(pn_ctx *ctx);
PN_EXPORT int pn_get_activity(pn_ctx *ctx, uint8_t *nspring_flag, uint8_t *spring_active, uint8_t *member_active);
/* Overwrites the activity flags (any pointer may be NULL = unchanged) without touching stored displacements or load
 * vectors: Analysis._calc_reactions :1059-1313 runs after ALL combinations and sees every element's per-combination
 * `active[combo]` but the node support springs as the LAST combination left them; this call reproduces that state
 * before pn_reactions / pn_element_forces.  The matrix has to be re-assembled before the next solve.                */
PN_EXPORT int pn_set_activity(pn_ctx *ctx, const uint8_t *nspring_flag, const uint8_t *spring_active, const uint8_t *member_active);

/* ---- internal forces / stresses of shells for solved combos: Quad3D.shear :1017, moment :1096, membrane :1173
 *      (points = natural coordinates (xi, eta), values at the Gauss points extrapolated), Plate3D.shear :608,
 *      moment :590, membrane :650 (points = local (x, y); the reference's plate methods ignore `local`).
 *      Every element of `kind` (PN_QUAD / PN_PLATE), combos [combo0, combo0 + n_combo), n_pts points pts[n_pts][2]:
 *      out[n_combo][count][n_pts][9] = Qx Qy 0 | Mx My Mxy | Sx Sy Txy  (local != 0)
 *                                      Qx Qy Qz | Mx My Mxy | Sx Sy Sxy  (local == 0, quads: Quad3D.py:1069-1092,
 *                                      1146-1171, 1214-1239)                                                        */
PN_EXPORT int pn_shell_stresses(pn_ctx *ctx, int kind, int32_t combo0, int32_t n_combo, int32_t n_pts, const double *pts,
                                int32_t local, double *out);

/* ---- multi-GPU direct solve: subtree-to-GPU multifrontal (one pn_ctx per GPU / rank) -------------------
 * Every rank holds the whole model and K.  The elimination tree is cut at the highest level that has at least `world`
 * fronts: rank r factorises and sweeps the subtrees below its share of that level for ALL right-hand sides; the
 * levels above the cut are factorised and swept by every rank (replicated).  Three exchanges per factor + solve, each
 * an all-gather of equal-size segments:
 *   (1) the update matrices of the subtree roots (after the local factorisation),
 *   (2) the update vectors of the subtree roots (after the local forward sweep),
 *   (3) the solved rows of every subtree (after the local backward sweep), so that every rank ends with all of X.
 * The library packs its segment into a device buffer it owns and calls `allgather(user, send, recv, seg_doubles)`:
 * the callee must all-gather send[0 : seg] of every rank into recv[r * seg : (r + 1) * seg] (r = 0 .. world-1) on every
 * rank and return 0 once the data is visible to work submitted afterwards on the library's stream (pn_set_stream) --
 * e.g. ncclAllGather on that stream, or torch.distributed.all_gather_into_tensor as pynite_b200/distsolve.py does.
 * Both pointers are DEVICE pointers.  world <= 1 (or a NULL callback) switches the mode off.  A model whose tree has
 * fewer than `world` fronts on every level is solved replicated, without exchanges.
 * Results are bit-identical to the single-GPU solve (same kernels, same summation order).                        */
typedef int (*pn_allgather_fn)(void *user, const double *send_dev, double *recv_dev, int64_t seg_doubles);
PN_EXPORT int pn_set_distributed(pn_ctx *ctx, int32_t rank, int32_t world, pn_allgather_fn allgather, void *user);

/* ---- device-pointer building blocks for the domain-decomposed PCG --------------------------------
 * pynite_b200/ddpcg.py runs one process per GPU; every rank holds the whole K11 (assembly is 0.1 ms), owns a
 * node-aligned strip of rows, and drives the CG recurrence with these kernels, exchanging the halo rows of the search
 * direction and all-reducing the dot products with torch.distributed (NCCL).  Every pointer below is a DEVICE pointer
 * on the context's GPU; vectors are [n1][s] row-major.  pn_set_stream makes the library launch on the caller's CUDA
 * stream (e.g. torch's current stream) so that its kernels are ordered with the caller's own work.               */
/* reset != 0: back to the stream pn_create made; otherwise cuda_stream is used as is (0 = the legacy default stream) */
PN_EXPORT int pn_set_stream(pn_ctx *ctx, void *cuda_stream, int reset);
/* first K11 row of every node's free DOFs, [n_nodes + 1] (host array): strips are cut at node boundaries          */
PN_EXPORT int pn_node_row_starts(pn_ctx *ctx, int64_t *row_start);
/* Y[row0:row1] = K11[row0:row1, :] X                                                                            */
PN_EXPORT int pn_dev_spmm_k11_rows(pn_ctx *ctx, int32_t s, int64_t row0, int64_t row1, const double *X_dev, double *Y_dev);
/* node-block (<= 6x6) Jacobi preconditioner of K11: setup once per assembly, then Z[rows of nodes node0..node1) = Minv R */
/* Columns [lo, hi) referenced by rows [row0, row1) of K11: the halo a domain-decomposed strip needs (computed on the
 * device; the pattern of a 6 M-DOF model is 0.5 GB and does not have to visit the host for this). */
PN_EXPORT int pn_k11_col_range(pn_ctx *ctx, int64_t row0, int64_t row1, int64_t *lo, int64_t *hi);
PN_EXPORT int pn_dev_block_jacobi_setup(pn_ctx *ctx);
PN_EXPORT int pn_dev_block_jacobi_nodes(pn_ctx *ctx, int32_t s, int64_t node0, int64_t node1, const double *R_dev, double *Z_dev);
/* fused CG vector kernels on a strip; the per-column scalars (rz, pAp, ...) are DEVICE arrays of s doubles so the caller
 * all-reduces them in place between the kernels:
 *   pn_dev_cg_dots       out[j] = sum over rows [row0,row1) of A[i][j] B[i][j]
 *   pn_dev_cg_update     x += alpha p, r -= alpha Ap with alpha = rz/pAp (skipped when first != 0), z = Minv r on nodes
 *                        [node0,node1); out2[0:s] = local r.z, out2[s:2s] = local r.r
 *   pn_dev_cg_direction  p = z + (rz_new/rz) p on rows [row0,row1)   (p = z when first != 0)                        */
PN_EXPORT int pn_dev_cg_dots(pn_ctx *ctx, int32_t s, int64_t row0, int64_t row1, const double *A_dev, const double *B_dev, double *out_dev);
PN_EXPORT int pn_dev_cg_update(pn_ctx *ctx, int32_t s, int64_t node0, int64_t node1, int32_t first, const double *rz_dev,
                               const double *pAp_dev, const double *P_dev, const double *AP_dev, double *X_dev, double *R_dev,
                               double *Z_dev, double *out2_dev);
PN_EXPORT int pn_dev_cg_direction(pn_ctx *ctx, int32_t s, int64_t row0, int64_t row1, int32_t first, const double *rz_dev,
                                  const double *rz_new_dev, const double *Z_dev, double *P_dev);
/* right-hand sides P1 - FER1 - K12 D2 (FEModel3D.py:2311) of combos [combo0, combo0 + s) into B_dev [n1][s]        */
PN_EXPORT int pn_dev_get_rhs(pn_ctx *ctx, int32_t combo0, int32_t s, double *B_dev);
/* stores solved free-DOF displacements X_dev [n1][s] of combos [combo0, combo0 + s) (Analysis._store_displacements
 * :847): afterwards pn_reactions / pn_element_forces / pn_get_displacements work for those combos                 */
PN_EXPORT int pn_dev_set_solution(pn_ctx *ctx, int32_t combo0, int32_t s, const double *X_dev);
PN_EXPORT int pn_get_displacements(pn_ctx *ctx, int32_t combo, double *D_out);

/* ---- building blocks exposed for measurement and tests ----------------------------------------
 * Y[n1][s] = K11 * X[n1][s] (row-major, s right-hand sides), repeated `repeat` times; returns the
 * average kernel time in ms through *ms_out (CUDA events on the launching stream).                */
PN_EXPORT int pn_spmm_k11(pn_ctx *ctx, int32_t s, const double *X, double *Y, int32_t repeat, double *ms_out);
PN_EXPORT int pn_get_stats(pn_ctx *ctx, pn_stats *out);
PN_EXPORT int pn_reset_stats(pn_ctx *ctx);
/* Which kernel families the direct solver took since pn_reset_stats (no reference counterpart: SuperLU is a black box
 * behind spsolve, Analysis.py:217).  Names: pn_counter_name(0..) until NULL -- "lu_update", "lu_update_later_groups"
 * (fronts with more than 256 pivots), "lu_strip", "split_levels" (front groups on side streams), "sweep_smem",
 * "sweep_ll", "sweep_step", "max_front_pivots", "levels", "factorisations", "pdelta_stationary_batches",
 * "pdelta_fallback_columns", "factor_fused_levels", "dist_exchanges", "dist_cut_level", "sweep_overlapped", "lookahead_groups", "sweep_ll_gemm".  The parity tests assert from these that a mid-size model exercises the
 * kernels the full-size configurations run. */
PN_EXPORT int pn_get_counter(pn_ctx *ctx, const char *name, int64_t *out);
PN_EXPORT const char *pn_counter_name(int index);
/* Device time of a host-chosen region on the library's stream (CUDA events): start synchronises the stream
 * and records, stop records, waits and returns the elapsed milliseconds.                              */
PN_EXPORT int pn_timer_start(pn_ctx *ctx);
PN_EXPORT int pn_timer_stop(pn_ctx *ctx, double *ms_out);
/* Per-kernel timing: when enabled, every launch of the main kernels is bracketed by CUDA events on the
 * library's stream; pn_profile_get returns the summed device time and launch count of one kernel slot.
 * pn_profile_slot_name(slot) returns NULL past the last slot.                                          */
PN_EXPORT int pn_profile_enable(pn_ctx *ctx, int on);
PN_EXPORT int pn_profile_get(pn_ctx *ctx, int slot, double *ms, int64_t *launches);
PN_EXPORT const char *pn_profile_slot_name(int slot);

#ifdef __cplusplus
}
#endif
#endif /* PYNITE_B200_H */
