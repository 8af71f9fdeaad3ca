// Internal declarations of the pynite_b200 CUDA library: context, device buffers, launcher
// prototypes.  Not part of the public ABI (that is include/pynite_b200.h).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <functional>
#include <future>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pynite_b200.h"

namespace pn {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct ArgError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct StatusError : std::runtime_error {
    int code;
    StatusError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define PN_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t err__ = (call);                                                                \
        if (err__ != cudaSuccess)                                                                  \
            throw pn::CudaError(std::string(#call) + " failed: " + cudaGetErrorString(err__) +     \
                                " (" __FILE__ ":" + std::to_string(__LINE__) + ")");               \
    } while (0)

// Stream every allocation / free of the current API call is ordered on (set by PN_API_BEGIN).  Device memory comes
// from the device's default stream-ordered pool with an unlimited release threshold, so temporaries of the symbolic
// phase are recycled without cudaMalloc / cudaFree round trips to the driver (those cost milliseconds each and, on a
// busy host, occasionally hundreds).
extern thread_local cudaStream_t tls_stream;

template <typename T>
struct DevBuf {
    T *ptr = nullptr;
    int64_t n = 0;        // elements in use
    int64_t cap = 0;      // elements allocated
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (ptr) {
            if (tls_stream) cudaFreeAsync(ptr, tls_stream);
            else cudaFree(ptr);
        }
        ptr = nullptr;
        n = cap = 0;
    }
    // grow-only allocation; contents are undefined afterwards
    void resize(int64_t count) {
        if (count > cap) {
            release();
            size_t bytes = (size_t)(count > 0 ? count : 1) * sizeof(T);
            if (tls_stream) PN_CUDA(cudaMallocAsync((void **)&ptr, bytes, tls_stream));
            else PN_CUDA(cudaMalloc((void **)&ptr, bytes));
            cap = count > 0 ? count : 1;
        }
        n = count;
    }
    void zero(cudaStream_t st) {
        if (n) PN_CUDA(cudaMemsetAsync(ptr, 0, (size_t)n * sizeof(T), st));
    }
    void upload(const T *host, int64_t count, cudaStream_t st) {
        resize(count);
        if (count) PN_CUDA(cudaMemcpyAsync(ptr, host, (size_t)count * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    void download(T *host, cudaStream_t st) const {
        if (n) PN_CUDA(cudaMemcpyAsync(host, ptr, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, st));
    }
};

struct Model {
    int64_t n_nodes = 0, n_springs = 0, n_members = 0, n_quads = 0, n_plates = 0;
    DevBuf<double> xyz;
    DevBuf<uint8_t> support, enf_mask, nspring_flag;
    DevBuf<double> enf_val, nspring_k;
    DevBuf<int32_t> spring_ij;
    DevBuf<double> spring_ks;
    DevBuf<uint8_t> spring_active;
    DevBuf<int32_t> mem_ij;
    DevBuf<double> mem_props, mem_rot;
    DevBuf<uint16_t> mem_rel;
    DevBuf<uint8_t> mem_active;
    DevBuf<int32_t> quad_ijmn, plate_ijmn;
    DevBuf<double> quad_props, plate_props;
    // host copies needed by host-side logic
    std::vector<uint8_t> h_support, h_enf_mask, h_nspring_flag, h_spring_active, h_mem_active;
    std::vector<double> h_nspring_k, h_xyz;        // (enforced values are only needed on the device)
    std::vector<int32_t> h_spring_ij, h_mem_ij, h_quad_ijmn, h_plate_ijmn;
    std::vector<double> h_spring_ks, h_mem_props, h_mem_rot, h_quad_props, h_plate_props;
    std::vector<uint16_t> h_mem_rel;
    bool symmetric = true;      // false when a shell has kx_mod != ky_mod (membrane matrix unsymmetric)
};

struct Loads {
    int32_t n_case = 0;
    // nodal loads (unique (dof, case), pre-summed by the host binding)
    DevBuf<int32_t> nl_dof, nl_case;
    DevBuf<double> nl_val;
    // member loads grouped by (member, case)
    int64_t n_mloads = 0, n_mgroups = 0;
    DevBuf<int32_t> ml_kind;
    DevBuf<double> ml_params;
    DevBuf<int32_t> mg_member, mg_case;
    DevBuf<int64_t> mg_start;
    DevBuf<double> mg_fer_local, mg_fer_global;     // [n_mgroups][12]
    // shell pressures (unique (element, case), pre-summed by the host binding)
    DevBuf<int32_t> qp_elem, qp_case, pp_elem, pp_case;
    DevBuf<double> qp_val, pp_val;
    std::vector<uint8_t> case_has_quad, case_has_plate;
    // dense per-case vectors
    DevBuf<double> Pcase;      // [n_case][n_dof]  nodal loads
    DevBuf<double> fer_case;   // [n_case][ef_count] element fixed-end forces, global axes
    DevBuf<double> Gcase;      // [n_case][n_dof]  P - FER
    int32_t n_combo = 0;
    DevBuf<double> factors;    // [n_combo][n_case]
    std::vector<double> h_factors;
    bool vectors_ready = false;
};

// A block of the partitioned matrix (K11, K12, K21, K22) in CSR with a gather map into K's values.
struct CsrBlock {
    int64_t rows = 0, cols = 0, nnz = 0;
    DevBuf<int32_t> indptr, indices;
    DevBuf<uint32_t> src;      // slot in the unpartitioned CSR
    DevBuf<double> vals;
};

// Per-kernel device timing (CUDA events on the context stream around individual launches), switched on by
// pn_profile_enable: bench.py reads it to report each kernel's share of a step and its roofline fraction.
enum ProfSlot {
    PS_ELEM_KE = 0, PS_SCATTER, PS_GATHER_BLOCKS, PS_RHS, PS_SPMM, PS_FRONT_ASSEMBLE, PS_EXTEND_ADD, PS_LU_DIAG,
    PS_LU_PANELS, PS_LU_UPDATE, PS_COPY_FACTORS, PS_SOLVE_GATHER, PS_SOLVE_DIAG, PS_SOLVE_UPDATE, PS_SOLVE_MISC,
    PS_REACTIONS, PS_PCG_VECTOR, PS_MERGE, PS_ASSEMBLY_TOTAL, PS_SOLVE_FUSED, PS_FACTOR_FUSED, PS_COUNT
};
struct Profiler {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Rec { int slot; size_t e0, e1; };
    std::vector<Rec> recs;
    double ms[PS_COUNT] = {0};
    int64_t n[PS_COUNT] = {0};
};

enum PathCounter {
    PC_LU_UPDATE = 0,      // k_lu_update launches (rank <= 256 trailing update, one per pivot group)
    PC_LU_UPDATE_LATER,    // ... of a second or later pivot group: fronts with more than 256 pivots
    PC_LU_STRIP,           // k_lu_strip launches (rank-64 strips inside a pivot group)
    PC_SPLIT_LEVELS,       // levels whose fronts ran as groups on the side streams
    PC_SWEEP_SMEM,         // k_front_forward_sm / k_front_backward_sm launches
    PC_SWEEP_LL,           // k_ll_sweep launches
    PC_SWEEP_STEP,         // per-step GEMM sweep launches (fallback family)
    PC_MAX_FRONT_NS,       // largest separator (pivots of one front)
    PC_LEVELS,             // elimination-tree levels
    PC_FACTORISATIONS,     // numeric factorisations
    PC_PDELTA_STATIONARY,  // P-Delta batches solved by the stationary iteration on the elastic factor
    PC_PDELTA_PERCOLUMN,   // P-Delta columns finished by preconditioned GMRES (or their own factorisation)
    PC_FACTOR_FUSED,       // levels factorised by the one-CTA-per-front kernel
    PC_DIST_EXCHANGES,     // all-gathers of the subtree-to-GPU solve (update matrices, update vectors, solved rows)
    PC_DIST_CUT_LEVEL,     // level of the subtree roots (-1: not distributed)
    PC_SWEEP_OVERLAPPED,   // factorisations whose first forward sweep ran behind them on the side stream
    PC_LOOKAHEAD,          // front groups factorised with the look-ahead schedule (trailing update split in two streams)
    PC_SWEEP_LL_GEMM,      // boundary-row GEMM launches beside k_ll_sweep (k_solve_forward_boundary / k_solve_boundary)
    PC_DOF_CLASSES,        // DOF-type classes the ordering dissected separately (1: node-based dissection)
    PC_DOF_CLASS_REJECTS,  // analyses redone with one class because an assembled entry joined two predicted classes
    PC_COUNT
};

struct DirectSolver;   // pn_direct.cu
struct FastAssembly;   // pn_assemble.cu

struct ElemInputs {
    int64_t n_springs, n_members, n_quads, n_plates;
    const int32_t *spring_ij; const double *spring_ks;
    const int32_t *mem_ij; const double *mem_props, *mem_rot; const uint16_t *mem_rel;
    const int32_t *quad_ijmn; const double *quad_props;
    const int32_t *plate_ijmn; const double *plate_props;
    const double *xyz;
};
struct PcgWork;        // pn_pcg.cu

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;  // the stream created by pn_create (pn_set_stream may redirect `stream`)
    cudaStream_t stream2 = nullptr;     // side stream: work that may overlap the main stream (class verification)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    pn_stats stats{};
    int64_t launches = 0;
    void count_launch(int64_t k = 1) { launches += k; }

    Model model;
    Loads loads;
    bool quad_symmetric = true, plate_symmetric = true;

    // element-value buffer: [node springs | springs*144 | members*144 | quads*576 | plates*576]
    DevBuf<double> ev;
    int64_t ev_count = 0, n_nsp = 0;
    int64_t ev_off_spring = 0, ev_off_member = 0, ev_off_quad = 0, ev_off_plate = 0;
    DevBuf<int32_t> nsp_dof;       // global DOF of each active node support spring, model order
    DevBuf<double> nsp_val;
    bool ev_ready = false;
    int ev_assemblies = 0;          // assemblies that consumed the current element values (see assemble())

    // symbolic results
    bool symbolic_ready = false;
    uint32_t sym_flags = 0;
    int64_t n_dof = 0, nnz_coo = 0, nnz_csr = 0, n1 = 0, n2 = 0;
    DevBuf<uint64_t> coo_key;      // sorted (row << 32 | col)
    DevBuf<uint32_t> coo_src;      // ev position of each sorted COO entry
    DevBuf<uint32_t> coo_emit;     // ev position of each COO entry in emission order
    DevBuf<uint32_t> seg_start;    // [nnz_csr + 1] first sorted COO entry of each CSR slot
    DevBuf<int32_t> indptr, indices;
    DevBuf<double> vals;
    bool vals_ready = false;
    DevBuf<int32_t> d1, d2, newidx; // newidx[dof] >= 0: position in d1; < 0: -1 - position in d2
    DevBuf<double> d2_val;
    CsrBlock blk[4];               // K11, K12, K21, K22
    // K11 rows grouped by node (the free DOFs of a node are consecutive rows): row_group[g] .. row_group[g + 1];
    // built lazily from the host masks, used by the node-grouped SpMM (one X-row load per node pair)
    DevBuf<int32_t> row_group;
    int64_t n_row_groups = 0;
    bool row_groups_ready = false;

    // element-force buffer layout [springs*12 | members*12 | quads*24 | plates*24] and DOF incidence
    int64_t ef_count = 0, ef_off_spring = 0, ef_off_member = 0, ef_off_quad = 0, ef_off_plate = 0;
    DevBuf<uint32_t> inc_src;      // ef positions sorted by global DOF (stable = emission order)
    DevBuf<uint32_t> inc_start;    // [n_dof + 1]
    bool incidence_ready = false;
    DevBuf<int32_t> sup_list[4];   // per element kind: elements touching a supported node (reactions only need these)

    // solve state
    DevBuf<double> B, X;           // [n1][s] row-major
    DevBuf<double> kd2;            // K12 * D2, [n1]
    DevBuf<double> D_all;          // [n_combo][n_dof]
    DevBuf<double> ef;             // element end forces of the combo being post-processed
    DevBuf<double> ef_batch;       // ... of a batch of combos (compute_reactions_batch)
    DevBuf<double> fer_combo;      // combined FER of that combo, ef layout
    DevBuf<double> memP;           // member axial forces
    bool solved = false;
    bool solved_pdelta = false;

    DirectSolver *direct = nullptr;
    PcgWork *pcg = nullptr;
    FastAssembly *fast = nullptr;
    int opt_host_threads = 0;       // pn_set_option("host_threads"): OpenMP threads of the host-worker jobs (0 = default)
    // pn_set_option("dof_classes"): 1 = the ordering of the direct solver treats DOF-type classes that never meet in a
    // stored entry as separate problems (predicted from the element list, verified against the assembled pattern);
    // 0 = node-based dissection only; 2 = self-test: predicts one class per DOF type, which the verification must reject
    int opt_dof_classes = 1;
    bool opt_spmm_grouped = false;  // pn_set_option("spmm_grouped"): K11 SpMM with the rows of a node sharing their X-row loads
    bool opt_assembly_graph = true; // pn_set_option("assembly_graph"): replay the numeric assembly as one CUDA graph
    int64_t n_assemblies = 0;       // numeric assemblies on the current pattern
    int64_t assembly_epoch = 0;     // incremented by every numeric assembly: the direct solver reuses a factor of the same epoch
    bool fast_build_tried = false;  // the class / recipe maps are built lazily, at the second assembly

    DevBuf<uint8_t> cub_tmp;       // scratch for cub device-wide primitives
    // scratch of the symbolic phase, kept (grow-only) so that re-analysing a model of the same size allocates nothing
    DevBuf<uint8_t> sym_flag, sym_known;
    DevBuf<uint32_t> sym_pos, sym_head, sym_incl, sym_kpos, sym_k32a, sym_k32b, sym_k32c;
    DevBuf<uint64_t> sym_key;
    DevBuf<int32_t> sym_slot_row, sym_cnt, sym_ent_row, sym_bad;
    DevBuf<int> sym_nb;
    DevBuf<double> sym_press;
    // persistent work arrays: a steady-state step performs no cudaMalloc / cudaFree
    DevBuf<double> work_r, work_ax, work_rx, dots_partial, dots_out;
    // DOFs whose reaction can be nonzero (supported, or with an active node support spring) and their rows
    bool rxn_list_ready = false;
    std::vector<int32_t> h_rxn_dofs;
    DevBuf<int32_t> rxn_dofs;
    DevBuf<double> rxn_rows;
    std::future<void> rxn_zero;    // zero fill of the caller's reaction array running on the host worker
    std::vector<double> h_rxn_rows;
    Profiler prof;
    cudaEvent_t tm0 = nullptr, tm1 = nullptr;   // pn_timer_start / pn_timer_stop
    // which kernel families the direct solver took since pn_reset_stats (pn_get_counter): the parity tests assert from
    // these that a model is large enough to exercise the production paths of the big configurations
    int64_t ctr[PC_COUNT] = {0};
    // tension/compression-only element modes (pn_set_tc_modes): bit0 tension-only, bit1 compression-only
    DevBuf<uint8_t> tc_spring_mode, tc_member_mode;
    DevBuf<int64_t> tc_group_start;      // sub-member range of every physical member
    int64_t tc_n_groups = 0;
    bool tc_any_member_mode = false;
    // pn_set_output_combos: the host arrays of pn_solve_linear / pn_solve_pdelta hold combinations [out_combo0,
    // out_combo0 + out_ncombo) only (out_ncombo < 0: all).  A rank of a multi-GPU solve delivers its share.
    int32_t out_combo0 = 0, out_ncombo = -1;
    // multi-GPU direct solve (pn_set_distributed): this context is rank `dist_rank` of `dist_world`
    int dist_rank = 0, dist_world = 1;
    pn_allgather_fn dist_allgather = nullptr;
    void *dist_user = nullptr;
};

void prof_flush(Ctx *c);
struct ProfScope {
    Ctx *c;
    int slot;
    size_t e0 = 0;
    bool live;
    cudaStream_t stream;
    static size_t grab(Ctx *c) {
        Profiler &p = c->prof;
        if (p.used == p.pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            p.pool.push_back(e);
        }
        return p.used++;
    }
    // `on`: the stream the timed launches go to (default: the context stream)
    ProfScope(Ctx *ctx, int s, cudaStream_t on = nullptr) : c(ctx), slot(s), live(ctx->prof.on), stream(on ? on : ctx->stream) {
        if (!live) return;
        if (c->prof.used > 60000) prof_flush(c);
        e0 = grab(c);
        cudaEventRecord(c->prof.pool[e0], stream);
    }
    ~ProfScope() {
        if (!live) return;
        size_t e1 = grab(c);
        cudaEventRecord(c->prof.pool[e1], stream);
        c->prof.recs.push_back({slot, e0, e1});
    }
};

// Host wall-clock checkpoints of the symbolic phases, printed when PN_DEBUG_TIMING is set
struct HostTick {
    const char *tag;
    bool on;
    bool sync = true;         // false on worker threads: wall clock only, no device synchronisation
    double t0;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    explicit HostTick(const char *t) : tag(t), on(getenv("PN_DEBUG_TIMING") != nullptr), t0(now()) {}
    void lap(const char *what) {
        if (!on) return;
        if (sync) cudaDeviceSynchronize();
        double t = now();
        fprintf(stderr, "[pn-time] %-18s %-28s %8.2f ms\n", tag, what, t - t0);
        t0 = t;
    }
};

// RAII timer on the context stream
struct PhaseTimer {
    Ctx *c;
    double *acc;
    PhaseTimer(Ctx *ctx, double *accum) : c(ctx), acc(accum) { cudaEventRecord(c->ev0, c->stream); }
    ~PhaseTimer() {
        cudaEventRecord(c->ev1, c->stream);
        cudaEventSynchronize(c->ev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        *acc += ms;
    }
};

// One persistent host worker thread per process for the jobs that run beside the device work (ordering of the direct
// solver, zero fill of result arrays).  A persistent thread keeps its OpenMP team between jobs: a fresh std::async thread
// per job paid the creation of a 16-thread team every time (ordering 23 ms in situ against 5.6 ms warm).  Jobs run FIFO.
std::future<void> host_worker_submit(std::function<void()> job);
int host_worker_threads(const Ctx *c);

// ---- pn_elements.cu -------------------------------------------------------------------------------
void launch_element_ke(Ctx *c, cudaStream_t st);
void launch_element_ke_arrays(Ctx *c, const ElemInputs &in, double *out_spring, double *out_member, double *out_quad,
                              double *out_plate, cudaStream_t st);
void launch_member_kg(Ctx *c, const double *P_dev, double *out_dev, cudaStream_t st);
void launch_member_axial(Ctx *c, const double *D_dev, double *P_dev, cudaStream_t st);
void launch_member_fer_groups(Ctx *c, cudaStream_t st);
void launch_shell_fer(Ctx *c, bool quad, const double *pressure_dev, double *out_dev, cudaStream_t st);
void launch_element_forces(Ctx *c, int kind, const double *D_dev, int32_t combo, bool supported_only, double *out_dev,
                           cudaStream_t st, int n_batch = 1, int64_t out_stride = 0);
void launch_member_pdelta_forces(Ctx *c, const double *D_dev, bool supported_only, double *F_dev, cudaStream_t st);
void launch_forces_to_local(Ctx *c, int kind, double *F_dev, cudaStream_t st);

// ---- pn_tc.cu -------------------------------------------------------------------------------------
int64_t tc_update(Ctx *c, int32_t combo, bool pdelta, double spring_tol, double member_tol);
void tc_reset_elements(Ctx *c);

// ---- pn_stress.cu ---------------------------------------------------------------------------------
void launch_shell_stress(Ctx *c, bool quad, const double *D_dev, int n_combo, int n_pts, const double *pts_dev, int local,
                         double *out_dev, cudaStream_t st);

// ---- pn_sparse.cu ---------------------------------------------------------------------------------
void build_symbolic(Ctx *c, uint32_t flags);
void scatter_add_values(Ctx *c, const double *ev_dev, double *vals_dev);
void gather_block_values(Ctx *c, const double *vals_dev);
void build_incidence(Ctx *c);
void build_load_vectors(Ctx *c);
void build_rhs(Ctx *c, int32_t combo0, int s, const double *kd2_percol);   // B = factors * G restricted to D1 - K12 D2
void merge_displacements(Ctx *c, int32_t combo0, int s);   // X, D2 -> D_all[combo0 .. combo0+s)
void export_coo(Ctx *c, int32_t *row_dev, int32_t *col_dev, double *data_dev);
void compute_reactions(Ctx *c, int32_t combo, bool pdelta, double *rxn_dev /* [n_dof] */);
// Reactions of combinations combo0 .. combo0 + n_batch - 1 in one pass.  rows = null: all DOFs, out[j][n_dof];
// otherwise only the listed DOFs, out[j][n_rows] (the rows that can be nonzero: pn_api.cu reactions_out).
void compute_reactions_batch(Ctx *c, int32_t combo0, int n_batch, bool pdelta, const int32_t *rows_dev, int64_t n_rows,
                             double *out_dev);
int64_t check_nodal_stability(Ctx *c, int32_t *bad_host, int64_t cap);
// Y = A X for a CSR block, X/Y row-major with s columns
void spmm(Ctx *c, const CsrBlock &A, const double *vals, const double *X, double *Y, int s, cudaStream_t st);
void residual_dd(Ctx *c, const CsrBlock &A, const double *vals, int64_t vals_col_stride, const double *X, const double *B,
                 double *R, int s, cudaStream_t st, const int32_t *rowlist = nullptr, int64_t row0 = 0, int64_t n_rows = -1);
void spmm_rows(Ctx *c, const CsrBlock &A, const double *vals, const double *X, double *Y, int s, int64_t row0, int64_t row1,
               cudaStream_t st);
void spmm_k11_grouped(Ctx *c, const double *vals, const double *X, double *Y, int s, cudaStream_t st);
void spmv_plain(Ctx *c, const CsrBlock &A, const double *vals, const double *x, double *y, cudaStream_t st);

// ---- pn_assemble.cu -------------------------------------------------------------------------------
void fast_assembly_build(Ctx *c);             // symbolic: element classes + block-gather maps (after build_symbolic)
bool fast_assembly_run(Ctx *c, double *vals); // numeric: false when unavailable / a class check failed (caller falls back)
void fast_assembly_release(Ctx *c);
void fast_assembly_invalidate(Ctx *c);

// ---- pn_pcg.cu / pn_direct.cu ---------------------------------------------------------------------
// Solve K11 X = B for all s columns; `vals11` may differ per column when per_column_vals != 0
// (P-Delta: one matrix per combo, same pattern): vals11 is then [s][nnz11].
void pcg_solve(Ctx *c, const double *vals11, int per_column_vals, int s, const pn_solve_opts &o);
void pcg_release(Ctx *c);
void pcg_setup_precond(Ctx *c, const double *vals11);
void pcg_apply_precond_nodes(Ctx *c, int s, int64_t node0, int64_t node1, const double *R, double *Z);
void pcg_node_row_starts(Ctx *c, int64_t *out_host);
void pcg_strip_dots(Ctx *c, int s, int64_t row0, int64_t row1, const double *A, const double *B, double *out_dev);
void pcg_strip_update(Ctx *c, int s, int64_t node0, int64_t node1, int first, const double *rz_dev, const double *pAp_dev,
                      const double *P, const double *AP, double *X, double *R, double *Z, double *out2_dev);
void pcg_strip_direction(Ctx *c, int s, int64_t row0, int64_t row1, int first, const double *rz_dev, const double *rz_new_dev,
                         const double *Z, double *P);
void column_dots(Ctx *c, int64_t n, int s, const double *A, const double *B, double *out_host);
// up to three products in one pass; out_host is [n_which][s]
void column_dots_multi(Ctx *c, int64_t n, int s, int n_which, const double *const *A, const double *const *B, double *out_host);
void direct_solve(Ctx *c, const double *vals11, int per_column_vals, int s, const pn_solve_opts &o);
void direct_release(Ctx *c);
void direct_prefetch(Ctx *c);     // start the host half of the analysis on a worker thread (after the DOF partition is known)
void direct_invalidate(Ctx *c);
void direct_invalidate_pattern(Ctx *c);
void direct_set_distributed(Ctx *c);
bool direct_dist_sum(Ctx *c, double *vals, int n);                            // sum of host values over the ranks of a distributed solve
void direct_dist_row_share(Ctx *c, int64_t n1, int64_t *r0, int64_t *r1);     // this rank's share of the rows for residual norms
   // the rank / world / callback in the context changed

}  // namespace pn
