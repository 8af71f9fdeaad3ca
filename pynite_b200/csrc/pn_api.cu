// C ABI of the pynite_b200 library (include/pynite_b200.h): argument checking, host<->device copies,
// phase sequencing.  All arithmetic happens in the kernels of pn_elements.cu / pn_sparse.cu /
// pn_pcg.cu / pn_direct.cu; nothing here computes on the host.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <future>
#include <mutex>
#include <numeric>
#include <omp.h>
#include <thread>
#include <unistd.h>

#include "pn_internal.h"

namespace pn {
struct HostWorker {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::packaged_task<void()>> q;
    pid_t pid;
    HostWorker() : pid(getpid()) { std::thread([this] { run(); }).detach(); }
    void run() {
        for (;;) {
            std::packaged_task<void()> t;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !q.empty(); });
                t = std::move(q.front());
                q.pop_front();
            }
            t();                              // exceptions travel through the task's future
        }
    }
};

// OpenMP threads of the host-worker jobs: pn_set_option("host_threads") when set (several processes sharing the cores
// of one box: one rank per GPU), else one less than the OpenMP default -- the caller's thread is busy with the uploads
// and the symbolic phase while the ordering runs, and a team as wide as the machine then waits at every barrier for the
// member that shares its core (measured on the 16-core box: ordering 42 ms with 16 threads beside the caller, 23 ms alone).
int host_worker_threads(const Ctx *c) {
    if (c->opt_host_threads > 0) return c->opt_host_threads;
    return std::max(1, omp_get_max_threads() - 1);
}

std::future<void> host_worker_submit(std::function<void()> job) {
    // never destroyed: the thread sleeps on its condition variable until the process exits.  A forked child gets its own
    // worker (threads do not survive fork).
    static std::mutex guard;
    static HostWorker *w = nullptr;
    std::packaged_task<void()> t(std::move(job));
    std::future<void> f = t.get_future();
    std::lock_guard<std::mutex> g(guard);
    if (!w || w->pid != getpid()) w = new HostWorker();
    {
        std::lock_guard<std::mutex> lk(w->mu);
        w->q.push_back(std::move(t));
    }
    w->cv.notify_one();
    return f;
}
}  // namespace pn

using namespace pn;

namespace pn {
thread_local cudaStream_t tls_stream = nullptr;
}

namespace {

inline Ctx *C(pn_ctx *p) { return reinterpret_cast<Ctx *>(p); }

#define PN_API_BEGIN(ctx)                                                                          \
    if (!(ctx)) return PN_ERR_ARG;                                                                 \
    Ctx *c = C(ctx);                                                                               \
    try {                                                                                          \
        PN_CUDA(cudaSetDevice(c->device));                                                         \
        pn::tls_stream = c->stream;

#define PN_API_END                                                                                 \
    }                                                                                              \
    catch (const StatusError &e) { c->err = e.what(); return e.code; }                             \
    catch (const ArgError &e) { c->err = e.what(); return PN_ERR_ARG; }                            \
    catch (const CudaError &e) { c->err = e.what(); cudaGetLastError(); return PN_ERR_CUDA; }      \
    catch (const std::exception &e) { c->err = e.what(); return PN_ERR_ARG; }                      \
    return PN_OK;

inline unsigned grid_for(int64_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

void invalidate(Ctx *c) {
    c->rxn_list_ready = false;
    c->ev_ready = false;
    c->symbolic_ready = false;
    c->vals_ready = false;
    c->incidence_ready = false;
    c->solved = c->solved_pdelta = false;
    c->loads.vectors_ready = false;
    direct_invalidate(c);
}

// host copy of a model array (the analysis reads these copies: ordering, element classes, symmetric-mode detection).
// No OpenMP here: a parallel region on this thread would leave its team spinning on the cores the ordering worker
// (pn_prefetch_analysis) is about to use (measured: e2e 157 -> 187 ms).  Large arrays are cut into four pieces copied by
// plain threads that exit when they are done (20 MB of node arrays and 14 MB of quad arrays for config 3: the setters sit
// before the ordering can start, so their time counts twice).
template <typename T>
void host_copy(std::vector<T> &v, const T *src, int64_t n) {
    v.resize((size_t)n);
    if (!n) return;
    const size_t bytes = (size_t)n * sizeof(T);
    char *dst = reinterpret_cast<char *>(v.data());
    const char *s8 = reinterpret_cast<const char *>(src);
    constexpr int NT = 4;
    if (bytes < ((size_t)2 << 20)) { memcpy(dst, s8, bytes); return; }
    const size_t piece = ((bytes + NT - 1) / NT + 63) / 64 * 64;
    std::thread helpers[NT - 1];
    for (int q = 1; q < NT; ++q) {
        const size_t o = std::min(bytes, piece * q), len = std::min(bytes, piece * (q + 1)) - o;
        helpers[q - 1] = std::thread([dst, s8, o, len]() { if (len) memcpy(dst + o, s8 + o, len); });
    }
    memcpy(dst, s8, std::min(bytes, piece));
    for (auto &h : helpers) h.join();
}

template <typename T>
void download(const DevBuf<T> &b, T *host, cudaStream_t st) {
    if (host && b.n) PN_CUDA(cudaMemcpyAsync(host, b.ptr, (size_t)b.n * sizeof(T), cudaMemcpyDeviceToHost, st));
}

// element-value buffer layout and the list of active node support springs (FEModel3D.py:1587-1690)
void layout_ev(Ctx *c) {
    Model &m = c->model;
    std::vector<int32_t> dof;
    std::vector<double> val;
    for (int64_t nd = 0; nd < m.n_nodes; ++nd)
        for (int k = 0; k < 6; ++k) {
            uint8_t f = m.h_nspring_flag[6 * nd + k];
            if ((f & 1u) && (f & 2u)) {
                dof.push_back((int32_t)(6 * nd + k));
                val.push_back(m.h_nspring_k[6 * nd + k]);
            }
        }
    c->n_nsp = (int64_t)dof.size();
    c->nsp_dof.upload(dof.data(), c->n_nsp, c->stream);
    c->nsp_val.upload(val.data(), c->n_nsp, c->stream);
    c->ev_off_spring = c->n_nsp;
    c->ev_off_member = c->ev_off_spring + 144 * m.n_springs;
    c->ev_off_quad = c->ev_off_member + 144 * m.n_members;
    c->ev_off_plate = c->ev_off_quad + 576 * m.n_quads;
    c->ev_count = c->ev_off_plate + 576 * m.n_plates;
    c->ev.resize(c->ev_count);
    PN_CUDA(cudaStreamSynchronize(c->stream));     // dof/val are stack-local
}

void compute_ev(Ctx *c) {
    layout_ev(c);
    PhaseTimer tm(c, &c->stats.ms_elements);
    if (c->n_nsp)
        PN_CUDA(cudaMemcpyAsync(c->ev.ptr, c->nsp_val.ptr, sizeof(double) * c->n_nsp, cudaMemcpyDeviceToDevice, c->stream));
    launch_element_ke(c, c->stream);
    c->ev_ready = true;
    c->ev_assemblies = 0;
}

void ensure_ev(Ctx *c) { if (!c->ev_ready) compute_ev(c); }

void ensure_symbolic(Ctx *c, uint32_t flags) {
    ensure_ev(c);
    if (c->symbolic_ready && c->sym_flags == flags) return;
    direct_invalidate_pattern(c);
    PhaseTimer tm(c, &c->stats.ms_symbolic);
    build_symbolic(c, flags);
    direct_prefetch(c);
    // The element-class / node-recipe maps of the streaming assembly cost ~60 ms of host work at 250 000 quads, the
    // general path 3 ms of device time: they pay off from the second assembly of a pattern on (P-Delta, T/C
    // iterations, repeated analyses), so they are built lazily by assemble() unless PN_EAGER_FAST_ASSEMBLY is set.
    fast_assembly_invalidate(c);
    c->n_assemblies = 0;
    c->fast_build_tried = false;
    static const bool eager = getenv("PN_EAGER_FAST_ASSEMBLY") != nullptr;
    if (eager) { fast_assembly_build(c); c->fast_build_tried = true; }
    c->loads.vectors_ready = c->loads.vectors_ready && c->incidence_ready;
}

// refresh: recompute the element values from the device-resident model first (a numeric re-assembly)
void assemble(Ctx *c, bool refresh) {
    bool done = false;
    if (++c->n_assemblies >= 2 && !c->fast_build_tried) {
        PhaseTimer tm(c, &c->stats.ms_symbolic);
        fast_assembly_build(c);
        c->fast_build_tried = true;
    }
    {
        PhaseTimer tm(c, &c->stats.ms_scatter);
        done = fast_assembly_run(c, c->vals.ptr);      // element classes + block gather: no COO value reaches HBM
    }
    if (!done) {
        // a re-assembly recomputes the element values, except when they were computed for this very model a moment
        // ago (by the symbolic phase of the same analysis) and no assembly has consumed them yet
        if (refresh && c->ev_assemblies > 0) compute_ev(c);
        ++c->ev_assemblies;
        PhaseTimer tm(c, &c->stats.ms_scatter);
        scatter_add_values(c, c->ev.ptr, c->vals.ptr);
    }
    {
        PhaseTimer tm(c, &c->stats.ms_partition);
        gather_block_values(c, c->vals.ptr);
    }
    c->vals_ready = true;
    ++c->assembly_epoch;
}

void ensure_assembled(Ctx *c, uint32_t flags) {
    ensure_symbolic(c, flags);
    if (!c->vals_ready) assemble(c, false);
}

void ensure_loads(Ctx *c) {
    if (c->loads.vectors_ready) return;
    PhaseTimer tm(c, &c->stats.ms_rhs);
    build_load_vectors(c);
}

pn_solve_opts default_opts(const pn_solve_opts *o) {
    pn_solve_opts r;
    if (o) r = *o;
    else { r.solver = PN_SOLVER_DIRECT; r.check_stability = 1; r.max_iter = 0; r.refine_steps = 2; r.tol = 1e-12; r.singular_tol = 1e-6; }
    if (r.singular_tol <= 0) r.singular_tol = 1e-6;
    if (r.tol <= 0) r.tol = 1e-12;
    return r;
}

__global__ void k_add(int64_t n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) out[t] = a[t] + b[t];
}
__global__ void k_put_col(int64_t n, int s, int j, const double *__restrict__ v, double *__restrict__ M) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) M[i * s + j] = v[i];
}
__global__ void k_gather_vals(int64_t n, const uint32_t *__restrict__ src, const double *__restrict__ in, double *__restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[src[i]];
}
// Analysis._solve_unknown_disp :228-239: non-finite values or ||K x - b|| > tol ||b|| (||x|| > tol when
// b = 0) mean a singular matrix.  Returns the largest relative residual.
double residual_check(Ctx *c, const double *vals11, int percol, int s, const pn_solve_opts &o, bool raise) {
    int64_t n1 = c->n1;
    if (!n1 || !s) return 0.0;
    cudaStream_t st = c->stream;
    DevBuf<double> &AX = c->work_ax;
    AX.resize(n1 * s);
    // subtree-to-GPU solve: X is replicated, so every rank takes 1/world of the rows and the squared norms are summed
    int64_t r0 = 0, r1 = n1;
    if (o.solver != PN_SOLVER_PCG) direct_dist_row_share(c, n1, &r0, &r1);
    residual_dd(c, c->blk[0], vals11, percol ? c->blk[0].nnz : 0, c->X.ptr, c->B.ptr, AX.ptr, s, st, nullptr, r0, r1 - r0);
    std::vector<double> rr(s), bb(s), xx(s);
    {
        const double *v[3] = {AX.ptr + r0 * s, c->B.ptr + r0 * s, c->X.ptr + r0 * s};
        std::vector<double> all((size_t)3 * s);
        column_dots_multi(c, r1 - r0, s, 3, v, v, all.data());
        for (int j = 0; j < s; ++j) { rr[j] = all[j]; bb[j] = all[s + j]; xx[j] = all[2 * s + j]; }
    }
    if (r1 - r0 != n1) {
        std::vector<double> all((size_t)3 * s);
        for (int j = 0; j < s; ++j) { all[j] = rr[j]; all[s + j] = bb[j]; all[2 * s + j] = xx[j]; }
        direct_dist_sum(c, all.data(), 3 * s);
        for (int j = 0; j < s; ++j) { rr[j] = all[j]; bb[j] = all[s + j]; xx[j] = all[2 * s + j]; }
    }
    double worst = 0.0;
    for (int j = 0; j < s; ++j) {
        bool finite = std::isfinite(rr[j]) && std::isfinite(xx[j]);
        double rn = std::sqrt(rr[j]), bn = std::sqrt(bb[j]), xn = std::sqrt(xx[j]);
        bool bad = !finite || (bn > 0 ? rn > o.singular_tol * bn : xn > o.singular_tol);
        double rel = bn > 0 ? rn / bn : rn;
        if (!finite) rel = INFINITY;
        worst = std::max(worst, rel);
        if (bad && raise)
            throw StatusError(PN_ERR_SINGULAR, "The stiffness matrix is singular, which indicates that the structure is unstable.");
    }
    return worst;
}

void run_solver(Ctx *c, const double *vals11, int percol, int s, const pn_solve_opts &o) {
    if (!c->n1 || !s) return;
    if (o.solver == PN_SOLVER_PCG) pcg_solve(c, vals11, percol, s, o);
    else direct_solve(c, vals11, percol, s, o);
    double worst = residual_check(c, vals11, percol, s, o, o.check_stability != 0);
    c->stats.max_rel_residual = std::max(c->stats.max_rel_residual, worst);
}

// combos are processed in batches so that the [n1][s] work arrays stay bounded
int batch_size(Ctx *c, int n_combo) {
    int64_t per = std::max<int64_t>(c->n1, 1) * 8;          // bytes per column of one work array
    int64_t cap = (int64_t)(8.0e9 / (double)per);            // <= 8 GB per work array
    int s = (int)std::min<int64_t>(std::min<int64_t>(n_combo, 64), std::max<int64_t>(cap, 1));
    return std::max(s, 1);
}

// Reactions are exactly zero except at supported DOFs and at DOFs with an active node support spring (k_reactions).
// When those are few (a plate pinned on its perimeter: 6 000 of 1.5 M DOFs) only their rows are computed and cross PCIe;
// the host array is zero-filled by the host worker (OpenMP) while the device factorises (reactions_prepare, called at the
// start of the solve), then the rows are scattered into it.
static bool reactions_compact(Ctx *c) {
    const int64_t n_dof = c->n_dof;
    const Model &m = c->model;
    if (!c->rxn_list_ready) {
        c->h_rxn_dofs.clear();
        const uint8_t *flag = m.h_nspring_flag.data();
        const uint8_t *sup = m.h_support.data();
        for (int64_t nd = 0; nd < m.n_nodes; ++nd) {
            const uint8_t su = sup[nd];
            for (int k = 0; k < 6; ++k) {
                const uint8_t f = flag[6 * nd + k];
                if (((su >> k) & 1u) || ((f & 1u) && (f & 2u))) c->h_rxn_dofs.push_back((int32_t)(6 * nd + k));
            }
        }
        c->rxn_dofs.upload(c->h_rxn_dofs.data(), (int64_t)c->h_rxn_dofs.size(), c->stream);
        c->rxn_list_ready = true;
    }
    static const bool full = getenv("PN_FULL_REACTION_DOWNLOAD") != nullptr;
    return !full && (int64_t)c->h_rxn_dofs.size() * 8 < n_dof;
}

// Starts the zero fill of the caller's reaction array on the host worker.  The job queues behind the ordering of the
// direct solver (same worker, FIFO), i.e. it runs while the device factorises and its threads do not compete with the
// ordering, which is on the critical path.
// combinations [*o0, *o0 + *on) go to the caller's host arrays (pn_set_output_combos; default: all)
static void output_range(const Ctx *c, int *o0, int *on) {
    const int nc = c->loads.n_combo;
    *o0 = 0; *on = nc;
    if (c->out_ncombo >= 0) {
        *o0 = std::min<int>(std::max<int>(c->out_combo0, 0), nc);
        *on = std::min<int>(c->out_ncombo, nc - *o0);
    }
}

void reactions_prepare(Ctx *c, double *Rxn_out) {
    c->rxn_zero = std::future<void>();
    if (!Rxn_out || !reactions_compact(c)) return;
    int o0 = 0, nc = 0;
    output_range(c, &o0, &nc);
    const int64_t n_dof = c->n_dof;
    const int wt = host_worker_threads(c);
    c->rxn_zero = host_worker_submit([Rxn_out, nc, n_dof, wt]() {
        const int64_t total = (int64_t)nc * n_dof, chunk = 1 << 20;
#pragma omp parallel for schedule(static) num_threads(wt)
        for (int64_t b = 0; b < (total + chunk - 1) / chunk; ++b)
            memset(Rxn_out + b * chunk, 0, sizeof(double) * (size_t)std::min(chunk, total - b * chunk));
    });
}

void reactions_out(Ctx *c, bool pdelta, double *Rxn_out) {
    if (!Rxn_out) return;
    PhaseTimer tm(c, &c->stats.ms_reactions);
    HostTick tick("reactions_out");
    tick.sync = false;
    build_incidence(c);
    tick.lap("incidence");
    int o0 = 0, nc = 0;
    output_range(c, &o0, &nc);
    if (!nc) return;
    const int64_t n_dof = c->n_dof;
    const bool compact = reactions_compact(c);
    const int64_t nr = (int64_t)c->h_rxn_dofs.size();
    if (!compact) {
        DevBuf<double> &rx = c->work_rx;
        rx.resize((int64_t)nc * n_dof);
        compute_reactions_batch(c, o0, nc, pdelta, nullptr, 0, rx.ptr);
        PN_CUDA(cudaMemcpyAsync(Rxn_out, rx.ptr, sizeof(double) * (size_t)nc * n_dof, cudaMemcpyDeviceToHost, c->stream));
        PN_CUDA(cudaStreamSynchronize(c->stream));
        return;
    }
    if (!c->rxn_zero.valid()) reactions_prepare(c, Rxn_out);        // (callers that did not announce the array)
    c->rxn_rows.resize(nr * nc);
    c->h_rxn_rows.resize((size_t)(nr * nc));
    if (nr) {
        compute_reactions_batch(c, o0, nc, pdelta, c->rxn_dofs.ptr, nr, c->rxn_rows.ptr);
        download(c->rxn_rows, c->h_rxn_rows.data(), c->stream);
    }
    tick.lap("launches enqueued");
    PN_CUDA(cudaStreamSynchronize(c->stream));
    tick.lap("stream sync");
    c->rxn_zero.get();
    c->rxn_zero = std::future<void>();
    tick.lap("wait for zero fill");
    const int32_t *rd = c->h_rxn_dofs.data();
    for (int j = 0; j < nc; ++j) {
        double *dst = Rxn_out + (int64_t)j * n_dof;
        const double *src = c->h_rxn_rows.data() + (int64_t)j * nr;
        for (int64_t i = 0; i < nr; ++i) dst[rd[i]] = src[i];
    }
    tick.lap("scatter rows");
}

// D_out != null: the displacements of a finished batch are copied to the host on the second stream while the next batch
// is solved (the factorisation is done once, see DirectSolver::factor_epoch).  PN_DOWNLOAD_OVERLAP=1 additionally cuts a
// single batch of >= 32 combinations in two so that half of the copy hides behind the second solve.  Measured on config 3
// (774 MB = 13.5 ms at 57 GB/s): two solves of 32 columns cost 38 ms against 31 ms for one of 64, which is more than the
// 7 ms of copy they hide (e2e 146.4 ms split, 143.0 ms unsplit), so it is off by default.
void solve_linear_all(Ctx *c, const pn_solve_opts &o, double *D_out = nullptr) {
    int nc = c->loads.n_combo;
    c->D_all.resize((int64_t)nc * c->n_dof);
    int bs = batch_size(c, nc);
    static const bool split = getenv("PN_DOWNLOAD_OVERLAP") != nullptr;
    if (D_out && split && nc >= 32 && bs >= nc && o.solver == PN_SOLVER_DIRECT) bs = (nc + 1) / 2;
    for (int j0 = 0; j0 < nc; j0 += bs) {
        int s = std::min(bs, nc - j0);
        {
            PhaseTimer tm(c, &c->stats.ms_rhs);
            build_rhs(c, j0, s, nullptr);
        }
        // solve, merge and start the download before the residual check (Analysis.py:228-239): the check is one more
        // pass over K11 plus host-synchronised norms (~3 ms for config 3) that the 771 MB copy does not have to wait
        // for; if it raises, the caller's buffer holds the rejected solution, as the reference's `D1` would
        if (c->n1 && s) {
            if (o.solver == PN_SOLVER_PCG) pcg_solve(c, c->blk[0].vals.ptr, 0, s, o);
            else direct_solve(c, c->blk[0].vals.ptr, 0, s, o);
        }
        merge_displacements(c, j0, s);
        if (D_out) {
            int o0 = 0, on = 0;
            output_range(c, &o0, &on);
            const int a = std::max(j0, o0), b = std::min(j0 + s, o0 + on);      // this batch's part of the output range
            if (b > a) {
                PN_CUDA(cudaEventRecord(c->ev_fork, c->stream));
                PN_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
                PN_CUDA(cudaMemcpyAsync(D_out + (int64_t)(a - o0) * c->n_dof, c->D_all.ptr + (int64_t)a * c->n_dof,
                                        sizeof(double) * (size_t)(b - a) * c->n_dof, cudaMemcpyDeviceToHost, c->stream2));
            }
        }
        if (c->n1 && s) {
            const double worst = residual_check(c, c->blk[0].vals.ptr, 0, s, o, o.check_stability != 0);
            c->stats.max_rel_residual = std::max(c->stats.max_rel_residual, worst);
        }
    }
}

}  // namespace

namespace pn {
void prof_flush(Ctx *c) {
    Profiler &p = c->prof;
    if (p.recs.empty()) { p.used = 0; return; }
    cudaDeviceSynchronize();            // scopes may have been recorded on the factorisation's side streams
    for (const Profiler::Rec &r : p.recs) {
        float ms = 0;
        cudaEventElapsedTime(&ms, p.pool[r.e0], p.pool[r.e1]);
        p.ms[r.slot] += ms;
        p.n[r.slot] += 1;
    }
    p.recs.clear();
    p.used = 0;
}
}  // namespace pn

extern "C" {

const char *pn_version(void) { return "pynite_b200 0.1 (sm_100a)"; }

int pn_create(pn_ctx **out, int device_id) {
    if (!out) return PN_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return PN_ERR_NO_DEVICE; }
    if (device_id < 0 || device_id >= n) return PN_ERR_ARG;
    if (cudaSetDevice(device_id) != cudaSuccess) { cudaGetLastError(); return PN_ERR_CUDA; }
    Ctx *c = new Ctx();
    c->device = device_id;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
        cudaGetLastError();
        delete c;
        return PN_ERR_CUDA;
    }
    c->own_stream = c->stream;
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    *out = reinterpret_cast<pn_ctx *>(c);
    return PN_OK;
}

void pn_destroy(pn_ctx *ctx) {
    if (!ctx) return;
    Ctx *c = C(ctx);
    cudaSetDevice(c->device);
    pn::tls_stream = c->stream;
    cudaStreamSynchronize(c->stream);
    direct_release(c);
    pcg_release(c);
    fast_assembly_release(c);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->tm0) cudaEventDestroy(c->tm0);
    if (c->tm1) cudaEventDestroy(c->tm1);
    for (cudaEvent_t e : c->prof.pool) cudaEventDestroy(e);
    cudaStream_t st = c->own_stream, st2 = c->stream2;
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    delete c;
    if (st) cudaStreamSynchronize(st);
    pn::tls_stream = nullptr;
    if (st) cudaStreamDestroy(st);
    if (st2) cudaStreamDestroy(st2);
}

const char *pn_last_error(pn_ctx *ctx) { return ctx ? C(ctx)->err.c_str() : "null context"; }

int pn_set_nodes(pn_ctx *ctx, int64_t n_nodes, const double *xyz, const uint8_t *support_mask,
                 const uint8_t *enforced_mask, const double *enforced_val, const double *spring_k,
                 const uint8_t *spring_flag) {
    PN_API_BEGIN(ctx)
    if (n_nodes < 0 || (n_nodes && (!xyz || !support_mask || !enforced_mask || !enforced_val || !spring_k || !spring_flag)))
        throw ArgError("pn_set_nodes: null array");
    if (6 * n_nodes >= (int64_t)0x7fffffff) throw ArgError("pn_set_nodes: more than 2^31 DOFs");
    Model &m = c->model;
    cudaStream_t st = c->stream;
    invalidate(c);
    m.n_nodes = n_nodes;
    // uploads are issued from the caller's buffers first so that the DMA (full speed when they are pinned) runs while
    // the host copies the analysis needs later (ordering, element classes) are made
    m.xyz.upload(xyz, 3 * n_nodes, st);
    m.support.upload(support_mask, n_nodes, st);
    m.enf_mask.upload(enforced_mask, n_nodes, st);
    m.enf_val.upload(enforced_val, 6 * n_nodes, st);
    m.nspring_k.upload(spring_k, 6 * n_nodes, st);
    m.nspring_flag.upload(spring_flag, 6 * n_nodes, st);
    host_copy(m.h_xyz, xyz, (int64_t)(3 * n_nodes));
    host_copy(m.h_support, support_mask, (int64_t)(n_nodes));
    host_copy(m.h_enf_mask, enforced_mask, (int64_t)(n_nodes));
    host_copy(m.h_nspring_k, spring_k, (int64_t)(6 * n_nodes));
    host_copy(m.h_nspring_flag, spring_flag, (int64_t)(6 * n_nodes));
    PN_CUDA(cudaStreamSynchronize(st));
    PN_API_END
}

static void check_conn(const Ctx *c, int64_t n, int nn, const int32_t *conn, const char *what) {
    for (int64_t k = 0; k < n * nn; ++k)
        if (conn[k] < 0 || conn[k] >= c->model.n_nodes) throw ArgError(std::string(what) + ": node index out of range");
}

int pn_set_springs(pn_ctx *ctx, int64_t n, const int32_t *ij, const double *ks, const uint8_t *active) {
    PN_API_BEGIN(ctx)
    if (n < 0 || (n && (!ij || !ks || !active))) throw ArgError("pn_set_springs: null array");
    check_conn(c, n, 2, ij, "pn_set_springs");
    Model &m = c->model;
    invalidate(c);
    m.n_springs = n;
    host_copy(m.h_spring_ij, ij, (int64_t)(2 * n));
    host_copy(m.h_spring_active, active, (int64_t)(n));
    host_copy(m.h_spring_ks, ks, (int64_t)(n));
    m.spring_ij.upload(ij, 2 * n, c->stream);
    m.spring_ks.upload(ks, n, c->stream);
    m.spring_active.upload(active, n, c->stream);
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

int pn_set_members(pn_ctx *ctx, int64_t n, const int32_t *ij, const double *props, const double *rotation_deg,
                   const uint16_t *releases, const uint8_t *active) {
    PN_API_BEGIN(ctx)
    if (n < 0 || (n && (!ij || !props || !rotation_deg || !releases || !active))) throw ArgError("pn_set_members: null array");
    check_conn(c, n, 2, ij, "pn_set_members");
    Model &m = c->model;
    invalidate(c);
    m.n_members = n;
    host_copy(m.h_mem_ij, ij, (int64_t)(2 * n));
    host_copy(m.h_mem_active, active, (int64_t)(n));
    host_copy(m.h_mem_props, props, (int64_t)(6 * n));
    host_copy(m.h_mem_rot, rotation_deg, (int64_t)(n));
    host_copy(m.h_mem_rel, releases, (int64_t)(n));
    m.mem_ij.upload(ij, 2 * n, c->stream);
    m.mem_props.upload(props, 6 * n, c->stream);
    m.mem_rot.upload(rotation_deg, n, c->stream);
    m.mem_rel.upload(releases, n, c->stream);
    m.mem_active.upload(active, n, c->stream);
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

static void set_shells(Ctx *c, bool quad, int64_t n, const int32_t *ijmn, const double *props) {
    if (n < 0 || (n && (!ijmn || !props))) throw ArgError("pn_set_quads/plates: null array");
    check_conn(c, n, 4, ijmn, "pn_set_quads/plates");
    Model &m = c->model;
    invalidate(c);
    (quad ? m.n_quads : m.n_plates) = n;
    (quad ? m.quad_ijmn : m.plate_ijmn).upload(ijmn, 4 * n, c->stream);
    (quad ? m.quad_props : m.plate_props).upload(props, 5 * n, c->stream);
    host_copy((quad ? m.h_quad_ijmn : m.h_plate_ijmn), ijmn, (int64_t)(4 * n));
    host_copy((quad ? m.h_quad_props : m.h_plate_props), props, (int64_t)(5 * n));
    // the membrane matrix is unsymmetric when kx_mod != ky_mod (Quad3D.py:517-519)
    bool sym = true;
    for (int64_t e = 0; e < n; ++e) if (props[5 * e + 3] != props[5 * e + 4]) sym = false;
    if (quad) c->quad_symmetric = sym; else c->plate_symmetric = sym;
    m.symmetric = c->quad_symmetric && c->plate_symmetric;
    PN_CUDA(cudaStreamSynchronize(c->stream));
}

int pn_set_quads(pn_ctx *ctx, int64_t n, const int32_t *ijmn, const double *props) {
    PN_API_BEGIN(ctx)
    set_shells(c, true, n, ijmn, props);
    PN_API_END
}
int pn_set_plates(pn_ctx *ctx, int64_t n, const int32_t *ijmn, const double *props) {
    PN_API_BEGIN(ctx)
    set_shells(c, false, n, ijmn, props);
    PN_API_END
}

int pn_element_ke(pn_ctx *ctx, int kind, double *out) {
    PN_API_BEGIN(ctx)
    if (!out) throw ArgError("pn_element_ke: null output");
    ensure_ev(c);
    const Model &m = c->model;
    int64_t off, cnt;
    switch (kind) {
    case PN_SPRING: off = c->ev_off_spring; cnt = 144 * m.n_springs; break;
    case PN_MEMBER: off = c->ev_off_member; cnt = 144 * m.n_members; break;
    case PN_QUAD: off = c->ev_off_quad; cnt = 576 * m.n_quads; break;
    case PN_PLATE: off = c->ev_off_plate; cnt = 576 * m.n_plates; break;
    default: throw ArgError("pn_element_ke: unknown element kind");
    }
    if (cnt) PN_CUDA(cudaMemcpyAsync(out, c->ev.ptr + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost, c->stream));
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

int pn_member_kg(pn_ctx *ctx, const double *P, double *out) {
    PN_API_BEGIN(ctx)
    int64_t n = c->model.n_members;
    if (n && (!P || !out)) throw ArgError("pn_member_kg: null array");
    DevBuf<double> dP, dK;
    dP.upload(P, n, c->stream);
    dK.resize(144 * n);
    launch_member_kg(c, dP.ptr, dK.ptr, c->stream);
    download(dK, out, c->stream);
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

int pn_set_option(pn_ctx *ctx, const char *name, int64_t value) {
    PN_API_BEGIN(ctx)
    if (!name) throw ArgError("pn_set_option: null name");
    if (!strcmp(name, "assembly_graph")) c->opt_assembly_graph = value != 0;
    else if (!strcmp(name, "spmm_grouped")) c->opt_spmm_grouped = value != 0;
    else if (!strcmp(name, "host_threads")) c->opt_host_threads = (int)std::max<int64_t>(0, value);
    else if (!strcmp(name, "dof_classes")) {
        if (value < 0 || value > 2) throw ArgError("pn_set_option: dof_classes takes 0 (off), 1 (predict and verify) or 2 (self-test)");
        if ((int)value != c->opt_dof_classes) { c->opt_dof_classes = (int)value; direct_invalidate(c); }
    }
    else throw ArgError(std::string("pn_set_option: unknown option ") + name);
    PN_API_END
}

int pn_set_output_combos(pn_ctx *ctx, int32_t combo0, int32_t n_combo) {
    PN_API_BEGIN(ctx)
    if (n_combo >= 0 && combo0 < 0) throw ArgError("pn_set_output_combos: negative first combination");
    c->out_combo0 = n_combo < 0 ? 0 : combo0;
    c->out_ncombo = n_combo < 0 ? -1 : n_combo;
    PN_API_END
}

int pn_prefetch_analysis(pn_ctx *ctx) {
    PN_API_BEGIN(ctx)
    direct_prefetch(c);
    PN_API_END
}

int pn_symbolic(pn_ctx *ctx, uint32_t flags) {
    PN_API_BEGIN(ctx)
    ensure_symbolic(c, flags);
    PN_API_END
}

int pn_get_sizes(pn_ctx *ctx, pn_sizes *out) {
    PN_API_BEGIN(ctx)
    if (!out) throw ArgError("pn_get_sizes: null output");
    if (!c->symbolic_ready) throw ArgError("pn_get_sizes: call pn_symbolic first");
    out->n_dof = c->n_dof; out->nnz_coo = c->nnz_coo; out->nnz_csr = c->nnz_csr; out->n1 = c->n1; out->n2 = c->n2;
    out->nnz11 = c->blk[0].nnz; out->nnz12 = c->blk[1].nnz; out->nnz21 = c->blk[2].nnz; out->nnz22 = c->blk[3].nnz;
    PN_API_END
}

int pn_get_coo(pn_ctx *ctx, int32_t *row, int32_t *col, double *data) {
    PN_API_BEGIN(ctx)
    if (!c->symbolic_ready) throw ArgError("pn_get_coo: call pn_symbolic first");
    DevBuf<int32_t> r, cc;
    DevBuf<double> d;
    r.resize(c->nnz_coo); cc.resize(c->nnz_coo); d.resize(c->nnz_coo);
    export_coo(c, r.ptr, cc.ptr, d.ptr);
    download(r, row, c->stream); download(cc, col, c->stream); download(d, data, c->stream);
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

int pn_get_pattern(pn_ctx *ctx, int32_t *indptr, int32_t *indices) {
    PN_API_BEGIN(ctx)
    if (!c->symbolic_ready) throw ArgError("pn_get_pattern: call pn_symbolic first");
    download(c->indptr, indptr, c->stream);
    download(c->indices, indices, c->stream);
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

int pn_get_partition(pn_ctx *ctx, int32_t *d1, int32_t *d2, double *d2_values) {
    PN_API_BEGIN(ctx)
    if (!c->symbolic_ready) throw ArgError("pn_get_partition: call pn_symbolic first");
    download(c->d1, d1, c->stream);
    download(c->d2, d2, c->stream);
    download(c->d2_val, d2_values, c->stream);
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

int pn_get_block(pn_ctx *ctx, int block, int32_t *indptr, int32_t *indices, double *data) {
    PN_API_BEGIN(ctx)
    if (block < 0 || block > 3) throw ArgError("pn_get_block: block must be 0..3");
    if (!c->symbolic_ready) throw ArgError("pn_get_block: call pn_symbolic first");
    if (data && !c->vals_ready) throw ArgError("pn_get_block: call pn_assemble_ke first");
    const CsrBlock &B = c->blk[block];
    download(B.indptr, indptr, c->stream);
    download(B.indices, indices, c->stream);
    download(B.vals, data, c->stream);
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

int pn_assemble_ke(pn_ctx *ctx) {
    PN_API_BEGIN(ctx)
    bool refresh = c->symbolic_ready;          // numeric re-assembly on a fixed pattern
    ensure_symbolic(c, c->symbolic_ready ? c->sym_flags : 0u);
    assemble(c, refresh);
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

int pn_get_csr_values(pn_ctx *ctx, double *data) {
    PN_API_BEGIN(ctx)
    if (!c->vals_ready) throw ArgError("pn_get_csr_values: call pn_assemble_ke first");
    download(c->vals, data, c->stream);
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

int pn_check_nodal_stability(pn_ctx *ctx, int32_t *bad_dofs, int64_t cap, int64_t *n_bad) {
    PN_API_BEGIN(ctx)
    if (!c->vals_ready) throw ArgError("pn_check_nodal_stability: call pn_assemble_ke first");
    std::vector<int32_t> tmp((size_t)std::max<int64_t>(cap, 1));
    int64_t n = check_nodal_stability(c, tmp.data(), std::max<int64_t>(cap, 1));
    if (n_bad) *n_bad = n;
    if (bad_dofs) for (int64_t i = 0; i < std::min(n, cap); ++i) bad_dofs[i] = tmp[i];
    if (n > 0) throw StatusError(PN_ERR_UNSTABLE_NODE, "Unstable node(s). See console output for details.");
    PN_API_END
}

int pn_set_loads(pn_ctx *ctx, int32_t n_case,
                 int64_t n_nodal, const int32_t *nl_dof, const int32_t *nl_case, const double *nl_val,
                 int64_t n_mload, const int32_t *ml_member, const int32_t *ml_case, const int32_t *ml_kind,
                 const double *ml_params,
                 int64_t n_qp, const int32_t *qp_elem, const int32_t *qp_case, const double *qp_val,
                 int64_t n_pp, const int32_t *pp_elem, const int32_t *pp_case, const double *pp_val) {
    PN_API_BEGIN(ctx)
    if (n_case < 0 || n_nodal < 0 || n_mload < 0 || n_qp < 0 || n_pp < 0) throw ArgError("pn_set_loads: negative count");
    Loads &ld = c->loads;
    const Model &m = c->model;
    cudaStream_t st = c->stream;
    ld.vectors_ready = false;
    c->solved = c->solved_pdelta = false;
    ld.n_case = n_case;
    int64_t n_dof = 6 * m.n_nodes;

    // duplicates of one (index, case) pair are summed in input (= model) order
    auto unique_sum = [&](int64_t n, const int32_t *idx, const int32_t *cs, const double *val, int64_t idx_limit,
                          const char *what, std::vector<int32_t> &oi, std::vector<int32_t> &oc, std::vector<double> &ov) {
        for (int64_t k = 0; k < n; ++k)
            if (idx[k] < 0 || idx[k] >= idx_limit || cs[k] < 0 || cs[k] >= n_case)
                throw ArgError(std::string(what) + ": index or case out of range");
        // (case, index) order by a stable counting sort on the case, then a comparison sort only inside the cases whose
        // indices do not already ascend (generated input arrives element-major: linear time)
        std::vector<int64_t> order(n), start((size_t)n_case + 1, 0);
        for (int64_t k = 0; k < n; ++k) ++start[cs[k] + 1];
        for (int32_t q = 0; q < n_case; ++q) start[q + 1] += start[q];
        {
            std::vector<int64_t> fill(start.begin(), start.end() - 1);
            for (int64_t k = 0; k < n; ++k) order[fill[cs[k]]++] = k;
        }
        for (int32_t q = 0; q < n_case; ++q) {
            auto b = order.begin() + start[q], e = order.begin() + start[q + 1];
            auto less = [&](int64_t a, int64_t bb) { return idx[a] < idx[bb]; };
            if (!std::is_sorted(b, e, less)) std::stable_sort(b, e, less);
        }
        oi.clear(); oc.clear(); ov.clear();
        oi.reserve(n); oc.reserve(n); ov.reserve(n);
        for (int64_t q = 0; q < n; ++q) {
            int64_t k = order[q];
            if (!oi.empty() && oi.back() == idx[k] && oc.back() == cs[k]) ov.back() += val[k];
            else { oi.push_back(idx[k]); oc.push_back(cs[k]); ov.push_back(val[k]); }
        }
    };
    std::vector<int32_t> oi, oc;
    std::vector<double> ov;
    unique_sum(n_nodal, nl_dof, nl_case, nl_val, n_dof, "pn_set_loads(nodal)", oi, oc, ov);
    ld.nl_dof.upload(oi.data(), (int64_t)oi.size(), st);
    ld.nl_case.upload(oc.data(), (int64_t)oc.size(), st);
    ld.nl_val.upload(ov.data(), (int64_t)ov.size(), st);
    PN_CUDA(cudaStreamSynchronize(st));
    unique_sum(n_qp, qp_elem, qp_case, qp_val, m.n_quads, "pn_set_loads(quad pressure)", oi, oc, ov);
    ld.case_has_quad.assign(n_case, 0);
    for (int32_t cs : oc) ld.case_has_quad[cs] = 1;
    ld.qp_elem.upload(oi.data(), (int64_t)oi.size(), st);
    ld.qp_case.upload(oc.data(), (int64_t)oc.size(), st);
    ld.qp_val.upload(ov.data(), (int64_t)ov.size(), st);
    PN_CUDA(cudaStreamSynchronize(st));
    unique_sum(n_pp, pp_elem, pp_case, pp_val, m.n_plates, "pn_set_loads(plate pressure)", oi, oc, ov);
    ld.case_has_plate.assign(n_case, 0);
    for (int32_t cs : oc) ld.case_has_plate[cs] = 1;
    ld.pp_elem.upload(oi.data(), (int64_t)oi.size(), st);
    ld.pp_case.upload(oc.data(), (int64_t)oc.size(), st);
    ld.pp_val.upload(ov.data(), (int64_t)ov.size(), st);
    PN_CUDA(cudaStreamSynchronize(st));

    // member loads grouped by (member, case), input order kept inside a group
    {
        std::vector<int64_t> order(n_mload);
        std::iota(order.begin(), order.end(), 0);
        for (int64_t k = 0; k < n_mload; ++k)
            if (ml_member[k] < 0 || ml_member[k] >= m.n_members || ml_case[k] < 0 || ml_case[k] >= n_case ||
                ml_kind[k] < 0 || ml_kind[k] > 17)
                throw ArgError("pn_set_loads(member): member, case or kind out of range");
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
            return ml_member[a] != ml_member[b] ? ml_member[a] < ml_member[b] : ml_case[a] < ml_case[b];
        });
        std::vector<int32_t> kind(n_mload), gm, gc;
        std::vector<double> par(4 * n_mload);
        std::vector<int64_t> gs;
        for (int64_t q = 0; q < n_mload; ++q) {
            int64_t k = order[q];
            kind[q] = ml_kind[k];
            for (int t = 0; t < 4; ++t) par[4 * q + t] = ml_params[4 * k + t];
            if (gm.empty() || gm.back() != ml_member[k] || gc.back() != ml_case[k]) {
                gm.push_back(ml_member[k]); gc.push_back(ml_case[k]); gs.push_back(q);
            }
        }
        gs.push_back(n_mload);
        ld.n_mloads = n_mload;
        ld.n_mgroups = (int64_t)gm.size();
        ld.ml_kind.upload(kind.data(), n_mload, st);
        ld.ml_params.upload(par.data(), 4 * n_mload, st);
        ld.mg_member.upload(gm.data(), ld.n_mgroups, st);
        ld.mg_case.upload(gc.data(), ld.n_mgroups, st);
        ld.mg_start.upload(gs.data(), ld.n_mgroups + 1, st);
        ld.mg_fer_local.resize(12 * ld.n_mgroups);
        ld.mg_fer_global.resize(12 * ld.n_mgroups);
        PN_CUDA(cudaStreamSynchronize(st));
    }
    PN_API_END
}

int pn_set_combos(pn_ctx *ctx, int32_t n_combo, const double *factors) {
    PN_API_BEGIN(ctx)
    Loads &ld = c->loads;
    if (n_combo < 0 || (n_combo && ld.n_case && !factors)) throw ArgError("pn_set_combos: null factors");
    ld.n_combo = n_combo;
    ld.h_factors.assign(factors, factors + (int64_t)n_combo * ld.n_case);
    ld.factors.upload(ld.h_factors.data(), (int64_t)ld.h_factors.size(), c->stream);
    PN_CUDA(cudaStreamSynchronize(c->stream));
    c->solved = c->solved_pdelta = false;
    PN_API_END
}

int pn_get_load_vector(pn_ctx *ctx, int32_t combo, int32_t which, double *out) {
    PN_API_BEGIN(ctx)
    Loads &ld = c->loads;
    if (combo < 0 || combo >= ld.n_combo || (which != 0 && which != 1) || !out) throw ArgError("pn_get_load_vector: bad argument");
    ensure_symbolic(c, c->symbolic_ready ? c->sym_flags : 0u);
    ensure_loads(c);
    int64_t n = c->n_dof;
    std::vector<double> P((size_t)ld.n_case * n), G;
    PN_CUDA(cudaMemcpyAsync(P.data(), ld.Pcase.ptr, sizeof(double) * P.size(), cudaMemcpyDeviceToHost, c->stream));
    if (which == 1) {
        G.resize(P.size());
        PN_CUDA(cudaMemcpyAsync(G.data(), ld.Gcase.ptr, sizeof(double) * G.size(), cudaMemcpyDeviceToHost, c->stream));
    }
    PN_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t d = 0; d < n; ++d) {
        double acc = 0.0;
        for (int cs = 0; cs < ld.n_case; ++cs) {
            double v = which == 0 ? P[(size_t)cs * n + d] : P[(size_t)cs * n + d] - G[(size_t)cs * n + d];
            acc = std::fma(ld.h_factors[(size_t)combo * ld.n_case + cs], v, acc);
        }
        out[d] = acc;
    }
    PN_API_END
}

int pn_solve_linear(pn_ctx *ctx, const pn_solve_opts *opts, double *D_out, double *Rxn_out, pn_stats *stats) {
    PN_API_BEGIN(ctx)
    pn_solve_opts o = default_opts(opts);
    Loads &ld = c->loads;
    c->solved = c->solved_pdelta = false;
    HostTick tick("pn_solve_linear");
    tick.sync = false;
    ensure_assembled(c, c->symbolic_ready ? c->sym_flags : 0u);
    ensure_loads(c);
    tick.lap("assembled + load vectors");
    reactions_prepare(c, Rxn_out);
    solve_linear_all(c, o, D_out);      // (the displacement download runs on the second stream, batch by batch)
    tick.lap("solve_linear_all");
    c->solved = true;
    reactions_out(c, false, Rxn_out);
    tick.lap("reactions_out");
    PN_CUDA(cudaStreamSynchronize(c->stream2));
    PN_CUDA(cudaStreamSynchronize(c->stream));
    tick.lap("final syncs (D download)");
    (void)ld;
    c->stats.kernel_launches = c->launches;
    if (stats) *stats = c->stats;
    PN_API_END
}

int pn_solve_pdelta(pn_ctx *ctx, const pn_solve_opts *opts, double *D_out, double *Rxn_out, pn_stats *stats) {
    PN_API_BEGIN(ctx)
    pn_solve_opts o = default_opts(opts);
    Loads &ld = c->loads;
    const Model &m = c->model;
    cudaStream_t st = c->stream;
    c->solved = c->solved_pdelta = false;
    // Kg has no zero filter (FEModel3D.py:1859-1885): the pattern must hold every member block densely
    ensure_assembled(c, PN_SYM_DENSE_MEMBER_BLOCKS);
    ensure_loads(c);
    solve_linear_all(c, o);                         // step 1 (Analysis.py:382-445)
    int nc = ld.n_combo;
    if (m.n_members && c->n1) {
        const CsrBlock &K11 = c->blk[0], &K12 = c->blk[1];
        DevBuf<double> ev2, memP, kg, valsj, vals11, vals12, kd2col, kd2all;
        ev2.resize(c->ev_count);
        PN_CUDA(cudaMemcpyAsync(ev2.ptr, c->ev.ptr, sizeof(double) * c->ev_count, cudaMemcpyDeviceToDevice, st));
        memP.resize(m.n_members); kg.resize(144 * m.n_members); valsj.resize(c->nnz_csr);
        vals12.resize(K12.nnz); kd2col.resize(c->n1);
        int bs = batch_size(c, nc);
        // also bound the per-combo K11 copies to ~8 GB
        bs = (int)std::max<int64_t>(1, std::min<int64_t>(bs, (int64_t)(8.0e9 / (8.0 * std::max<int64_t>(K11.nnz, 1)))));
        for (int j0 = 0; j0 < nc; j0 += bs) {
            int s = std::min(bs, nc - j0);
            vals11.resize((int64_t)s * K11.nnz);
            kd2all.resize(c->n1 * s);
            for (int j = 0; j < s; ++j) {          // step 2 matrices (FEModel3D.Kg :1826-1892, Analysis.py:395-402)
                PhaseTimer tm(c, &c->stats.ms_scatter);
                const double *D = c->D_all.ptr + (int64_t)(j0 + j) * c->n_dof;
                launch_member_axial(c, D, memP.ptr, st);
                launch_member_kg(c, memP.ptr, kg.ptr, st);
                k_add<<<grid_for(144 * m.n_members, 256), 256, 0, st>>>(144 * m.n_members, c->ev.ptr + c->ev_off_member,
                                                                         kg.ptr, ev2.ptr + c->ev_off_member);
                scatter_add_values(c, ev2.ptr, valsj.ptr);
                if (K11.nnz)
                    k_gather_vals<<<grid_for(K11.nnz, 256), 256, 0, st>>>(K11.nnz, K11.src.ptr, valsj.ptr,
                                                                          vals11.ptr + (int64_t)j * K11.nnz);
                kd2col.zero(st);
                if (K12.nnz) {
                    k_gather_vals<<<grid_for(K12.nnz, 256), 256, 0, st>>>(K12.nnz, K12.src.ptr, valsj.ptr, vals12.ptr);
                    spmv_plain(c, K12, vals12.ptr, c->d2_val.ptr, kd2col.ptr, st);
                }
                k_put_col<<<grid_for(c->n1, 256), 256, 0, st>>>(c->n1, s, j, kd2col.ptr, kd2all.ptr);
                c->count_launch(4);
            }
            {
                PhaseTimer tm(c, &c->stats.ms_rhs);
                build_rhs(c, j0, s, kd2all.ptr);
            }
            run_solver(c, vals11.ptr, 1, s, o);
            merge_displacements(c, j0, s);
        }
    }
    c->solved = c->solved_pdelta = true;
    if (D_out) {
        int o0 = 0, on = 0;
        output_range(c, &o0, &on);
        if (on) PN_CUDA(cudaMemcpyAsync(D_out, c->D_all.ptr + (int64_t)o0 * c->n_dof, sizeof(double) * (size_t)on * c->n_dof, cudaMemcpyDeviceToHost, st));
    }
    PN_CUDA(cudaStreamSynchronize(st));
    reactions_out(c, true, Rxn_out);
    c->stats.kernel_launches = c->launches;
    if (stats) *stats = c->stats;
    PN_API_END
}

int pn_get_kg_coo(pn_ctx *ctx, int32_t combo, int32_t first_step, int32_t *row, int32_t *col, double *data,
                  int64_t *n_out) {
    PN_API_BEGIN(ctx)
    const Model &m = c->model;
    if (!first_step && (!c->solved || combo < 0 || combo >= c->loads.n_combo))
        throw ArgError("pn_get_kg_coo: combo has not been solved");
    cudaStream_t st = c->stream;
    DevBuf<double> memP, kg;
    memP.resize(m.n_members); kg.resize(144 * m.n_members);
    if (first_step) memP.zero(st);
    else launch_member_axial(c, c->D_all.ptr + (int64_t)combo * c->n_dof, memP.ptr, st);
    launch_member_kg(c, memP.ptr, kg.ptr, st);
    std::vector<double> h(144 * (size_t)m.n_members);
    download(kg, h.data(), st);
    PN_CUDA(cudaStreamSynchronize(st));
    int64_t o = 0;
    for (int64_t e = 0; e < m.n_members; ++e) {
        if (!m.h_mem_active[e]) continue;
        for (int r = 0; r < 12; ++r)
            for (int q = 0; q < 12; ++q) {
                if (row) row[o] = m.h_mem_ij[2 * e + r / 6] * 6 + r % 6;
                if (col) col[o] = m.h_mem_ij[2 * e + q / 6] * 6 + q % 6;
                if (data) data[o] = h[(size_t)e * 144 + r * 12 + q];
                ++o;
            }
    }
    if (n_out) *n_out = o;
    PN_API_END
}

int pn_set_displacements(pn_ctx *ctx, int32_t combo, const double *D) {
    PN_API_BEGIN(ctx)
    if (combo < 0 || combo >= c->loads.n_combo || !D) throw ArgError("pn_set_displacements: bad combo index");
    if (!c->n_dof || c->n_dof != 6 * c->model.n_nodes) throw ArgError("pn_set_displacements: call pn_symbolic first");
    // Also the way results computed elsewhere (another rank's share of the combinations, an earlier per-combination
    // analysis) are made resident again for pn_reactions / pn_element_forces / pn_shell_stresses.
    const int64_t need = (int64_t)c->loads.n_combo * c->n_dof;
    if (c->D_all.n < need) {
        DevBuf<double> grown;
        grown.resize(need);
        grown.zero(c->stream);
        if (c->D_all.n && c->solved)
            PN_CUDA(cudaMemcpyAsync(grown.ptr, c->D_all.ptr, sizeof(double) * c->D_all.n, cudaMemcpyDeviceToDevice, c->stream));
        std::swap(grown.ptr, c->D_all.ptr); std::swap(grown.n, c->D_all.n); std::swap(grown.cap, c->D_all.cap);
    }
    PN_CUDA(cudaMemcpyAsync(c->D_all.ptr + (int64_t)combo * c->n_dof, D, sizeof(double) * c->n_dof, cudaMemcpyHostToDevice, c->stream));
    PN_CUDA(cudaStreamSynchronize(c->stream));
    c->solved = true;
    PN_API_END
}

int pn_reactions(pn_ctx *ctx, int32_t combo, int32_t pdelta, double *out) {
    PN_API_BEGIN(ctx)
    if (!c->solved || combo < 0 || combo >= c->loads.n_combo || !out) throw ArgError("pn_reactions: combo has not been solved");
    PhaseTimer tm(c, &c->stats.ms_reactions);
    ensure_ev(c);
    build_incidence(c);
    ensure_loads(c);
    DevBuf<double> rx;
    rx.resize(c->n_dof);
    compute_reactions(c, combo, pdelta != 0, rx.ptr);
    download(rx, out, c->stream);
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

int pn_element_forces(pn_ctx *ctx, int kind, int32_t combo, int32_t local, int32_t pdelta, double *out) {
    PN_API_BEGIN(ctx)
    if (!c->solved || combo < 0 || combo >= c->loads.n_combo) throw ArgError("pn_element_forces: combo has not been solved");
    if (!out) throw ArgError("pn_element_forces: null output");
    const Model &m = c->model;
    cudaStream_t st = c->stream;
    ensure_ev(c);
    build_incidence(c);
    ensure_loads(c);
    const double *D = c->D_all.ptr + (int64_t)combo * c->n_dof;
    int64_t off, cnt;
    switch (kind) {
    case PN_SPRING: off = c->ef_off_spring; cnt = 12 * m.n_springs; break;
    case PN_MEMBER: off = c->ef_off_member; cnt = 12 * m.n_members; break;
    case PN_QUAD: off = c->ef_off_quad; cnt = 24 * m.n_quads; break;
    case PN_PLATE: off = c->ef_off_plate; cnt = 24 * m.n_plates; break;
    default: throw ArgError("pn_element_forces: unknown element kind");
    }
    double *F = c->ef.ptr + off;
    launch_element_forces(c, kind, D, combo, false, F, st);
    if (kind == PN_MEMBER && pdelta) launch_member_pdelta_forces(c, D, false, F, st);
    if (local) launch_forces_to_local(c, kind, F, st);
    if (cnt) PN_CUDA(cudaMemcpyAsync(out, F, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    PN_API_END
}

int pn_shell_stresses(pn_ctx *ctx, int kind, int32_t combo0, int32_t n_combo, int32_t n_pts, const double *pts, int32_t local,
                      double *out) {
    PN_API_BEGIN(ctx)
    if (kind != PN_QUAD && kind != PN_PLATE) throw ArgError("pn_shell_stresses: kind must be PN_QUAD or PN_PLATE");
    if (!c->solved || combo0 < 0 || n_combo < 0 || combo0 + n_combo > c->loads.n_combo)
        throw ArgError("pn_shell_stresses: combos have not been solved");
    if (n_pts < 0 || (n_pts && !pts) || !out) throw ArgError("pn_shell_stresses: null array");
    const int64_t n = kind == PN_QUAD ? c->model.n_quads : c->model.n_plates;
    if (n && n_combo && n_pts) {
        cudaStream_t st = c->stream;
        DevBuf<double> dp, dout;
        dp.upload(pts, 2 * (int64_t)n_pts, st);
        // bounded device buffer: combos in batches of <= 2 GB of results
        const int64_t per_combo = n * n_pts * 9;
        const int step = (int)std::max<int64_t>(1, std::min<int64_t>(n_combo, (int64_t)(2.5e8 / (double)per_combo)));
        dout.resize(per_combo * step);
        for (int j0 = 0; j0 < n_combo; j0 += step) {
            const int nj = std::min(step, n_combo - j0);
            launch_shell_stress(c, kind == PN_QUAD, c->D_all.ptr + (int64_t)(combo0 + j0) * c->n_dof, nj, n_pts, dp.ptr, local, dout.ptr, st);
            PN_CUDA(cudaMemcpyAsync(out + (int64_t)j0 * per_combo, dout.ptr, sizeof(double) * per_combo * nj, cudaMemcpyDeviceToHost, st));
            PN_CUDA(cudaStreamSynchronize(st));
        }
    }
    PN_API_END
}

int pn_set_tc_modes(pn_ctx *ctx, const uint8_t *spring_mode, const uint8_t *member_mode, int64_t n_groups,
                    const int64_t *group_start) {
    PN_API_BEGIN(ctx)
    const Model &m = c->model;
    if ((m.n_springs && !spring_mode) || (m.n_members && (!member_mode || !group_start))) throw ArgError("pn_set_tc_modes: null array");
    if (m.n_members && (n_groups < 1 || group_start[0] != 0 || group_start[n_groups] != m.n_members))
        throw ArgError("pn_set_tc_modes: group_start must cover all sub-members");
    c->tc_spring_mode.upload(spring_mode, m.n_springs, c->stream);
    c->tc_member_mode.upload(member_mode, m.n_members, c->stream);
    c->tc_n_groups = m.n_members ? n_groups : 0;
    c->tc_group_start.upload(group_start, m.n_members ? n_groups + 1 : 0, c->stream);
    c->tc_any_member_mode = false;
    for (int64_t e = 0; e < m.n_members; ++e) if (member_mode[e]) c->tc_any_member_mode = true;
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

// the activity flags changed on the device: refresh the host copies and drop what depends on them
static void tc_sync_activity(Ctx *c, bool elements_changed) {
    Model &m = c->model;
    cudaStream_t st = c->stream;
    if (m.nspring_flag.n) PN_CUDA(cudaMemcpyAsync(m.h_nspring_flag.data(), m.nspring_flag.ptr, m.nspring_flag.n, cudaMemcpyDeviceToHost, st));
    if (m.n_springs) PN_CUDA(cudaMemcpyAsync(m.h_spring_active.data(), m.spring_active.ptr, m.n_springs, cudaMemcpyDeviceToHost, st));
    if (m.n_members) PN_CUDA(cudaMemcpyAsync(m.h_mem_active.data(), m.mem_active.ptr, m.n_members, cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    if (elements_changed) { invalidate(c); return; }
    // only node support springs switched: the element graph (and with it the ordering of the direct solver) is unchanged
    c->rxn_list_ready = false;
    c->ev_ready = false;
    c->symbolic_ready = false;
    c->vals_ready = false;
    c->solved = c->solved_pdelta = false;
}

int pn_tc_update(pn_ctx *ctx, int32_t combo, int32_t pdelta, double spring_tol, double member_tol, int64_t *n_changed) {
    PN_API_BEGIN(ctx)
    if (!c->solved || combo < 0 || combo >= c->loads.n_combo || !n_changed) throw ArgError("pn_tc_update: combo has not been solved");
    ensure_ev(c);
    build_incidence(c);
    ensure_loads(c);
    Model &m = c->model;
    // element flags before, to tell element deactivations from node-spring switches
    std::vector<uint8_t> s0 = m.h_spring_active, m0 = m.h_mem_active;
    const int64_t n = tc_update(c, combo, pdelta != 0, spring_tol, member_tol);
    *n_changed = n;
    if (n > 0) {
        tc_sync_activity(c, false);
        if (s0 != m.h_spring_active || m0 != m.h_mem_active) invalidate(c);
    }
    PN_API_END
}

int pn_tc_reset(pn_ctx *ctx) {
    PN_API_BEGIN(ctx)
    Model &m = c->model;
    bool any = false;
    for (uint8_t a : m.h_spring_active) if (!a) any = true;
    for (uint8_t a : m.h_mem_active) if (!a) any = true;
    if (any) {
        tc_reset_elements(c);
        tc_sync_activity(c, true);
    }
    PN_API_END
}

int pn_set_activity(pn_ctx *ctx, const uint8_t *nspring_flag, const uint8_t *spring_active, const uint8_t *member_active) {
    PN_API_BEGIN(ctx)
    Model &m = c->model;
    cudaStream_t st = c->stream;
    bool elements = false;
    if (nspring_flag && m.n_nodes) {
        memcpy(m.h_nspring_flag.data(), nspring_flag, (size_t)(6 * m.n_nodes));
        m.nspring_flag.upload(nspring_flag, 6 * m.n_nodes, st);
    }
    if (spring_active && m.n_springs) {
        elements = elements || memcmp(m.h_spring_active.data(), spring_active, (size_t)m.n_springs) != 0;
        memcpy(m.h_spring_active.data(), spring_active, (size_t)m.n_springs);
        m.spring_active.upload(spring_active, m.n_springs, st);
    }
    if (member_active && m.n_members) {
        elements = elements || memcmp(m.h_mem_active.data(), member_active, (size_t)m.n_members) != 0;
        memcpy(m.h_mem_active.data(), member_active, (size_t)m.n_members);
        m.mem_active.upload(member_active, m.n_members, st);
    }
    PN_CUDA(cudaStreamSynchronize(st));
    // stored displacements and load vectors stay valid (pn_reactions / pn_element_forces with this activity); the matrix does not
    c->rxn_list_ready = false;
    c->ev_ready = false;
    c->symbolic_ready = false;
    c->vals_ready = false;
    if (elements) direct_invalidate(c);
    PN_API_END
}

int pn_get_activity(pn_ctx *ctx, uint8_t *nspring_flag, uint8_t *spring_active, uint8_t *member_active) {
    PN_API_BEGIN(ctx)
    const Model &m = c->model;
    if (nspring_flag && !m.h_nspring_flag.empty()) memcpy(nspring_flag, m.h_nspring_flag.data(), m.h_nspring_flag.size());
    if (spring_active && m.n_springs) memcpy(spring_active, m.h_spring_active.data(), (size_t)m.n_springs);
    if (member_active && m.n_members) memcpy(member_active, m.h_mem_active.data(), (size_t)m.n_members);
    PN_API_END
}

int pn_set_distributed(pn_ctx *ctx, int32_t rank, int32_t world, pn_allgather_fn allgather, void *user) {
    PN_API_BEGIN(ctx)
    if (world > 1 && (rank < 0 || rank >= world)) throw ArgError("pn_set_distributed: rank out of range");
    if (world > 1 && !allgather) throw ArgError("pn_set_distributed: null all-gather callback");
    c->dist_rank = world > 1 ? rank : 0;
    c->dist_world = world > 1 ? world : 1;
    c->dist_allgather = world > 1 ? allgather : nullptr;
    c->dist_user = user;
    direct_set_distributed(c);
    PN_API_END
}

int pn_spmm_k11(pn_ctx *ctx, int32_t s, const double *X, double *Y, int32_t repeat, double *ms_out) {
    PN_API_BEGIN(ctx)
    if (!c->vals_ready) throw ArgError("pn_spmm_k11: call pn_assemble_ke first");
    if (s < 1 || s > 256 || repeat < 1) throw ArgError("pn_spmm_k11: bad s / repeat");
    cudaStream_t st = c->stream;
    DevBuf<double> dX, dY;
    int64_t n = c->n1 * s;
    dX.resize(n); dY.resize(n);
    if (X) PN_CUDA(cudaMemcpyAsync(dX.ptr, X, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    else {
        std::vector<double> ones((size_t)n, 1.0);
        PN_CUDA(cudaMemcpyAsync(dX.ptr, ones.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
        PN_CUDA(cudaStreamSynchronize(st));
    }
    auto apply = [&]() {
        if (c->opt_spmm_grouped) spmm_k11_grouped(c, c->blk[0].vals.ptr, dX.ptr, dY.ptr, s, st);
        else spmm(c, c->blk[0], c->blk[0].vals.ptr, dX.ptr, dY.ptr, s, st);
    };
    apply();                                                             // warm-up
    PN_CUDA(cudaEventRecord(c->ev0, st));
    for (int r = 0; r < repeat; ++r) apply();
    PN_CUDA(cudaEventRecord(c->ev1, st));
    PN_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0;
    PN_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (ms_out) *ms_out = (double)ms / repeat;
    c->stats.ms_spmm += ms;
    c->stats.spmm_launches += repeat;
    if (Y) PN_CUDA(cudaMemcpyAsync(Y, dY.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    PN_API_END
}

int pn_set_stream(pn_ctx *ctx, void *cuda_stream, int reset) {
    if (!ctx) return PN_ERR_ARG;
    Ctx *c = C(ctx);
    try {
        PN_CUDA(cudaSetDevice(c->device));
        PN_CUDA(cudaStreamSynchronize(c->stream));
        c->stream = reset ? c->own_stream : reinterpret_cast<cudaStream_t>(cuda_stream);
        pn::tls_stream = c->stream;
    } catch (const std::exception &e) { c->err = e.what(); return PN_ERR_CUDA; }
    return PN_OK;
}

int pn_node_row_starts(pn_ctx *ctx, int64_t *row_start) {
    PN_API_BEGIN(ctx)
    if (!c->symbolic_ready || !row_start) throw ArgError("pn_node_row_starts: call pn_symbolic first");
    pcg_node_row_starts(c, row_start);
    PN_API_END
}

int pn_dev_spmm_k11_rows(pn_ctx *ctx, int32_t s, int64_t row0, int64_t row1, const double *X_dev, double *Y_dev) {
    PN_API_BEGIN(ctx)
    if (!c->vals_ready) throw ArgError("pn_dev_spmm_k11_rows: call pn_assemble_ke first");
    if (s < 1 || s > 256 || row0 < 0 || row1 > c->n1 || row0 > row1 || !X_dev || !Y_dev) throw ArgError("pn_dev_spmm_k11_rows: bad argument");
    spmm_rows(c, c->blk[0], c->blk[0].vals.ptr, X_dev, Y_dev, s, row0, row1, c->stream);
    c->stats.spmm_launches += 1;
    PN_API_END
}

// smallest / largest column referenced by rows [row0, row1) of K11 (rows are sorted: first and last entry of each row)
__global__ void k_row_col_range(int64_t row0, int64_t row1, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                int32_t *__restrict__ lohi) {
    int64_t r = row0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= row1) return;
    const int32_t a = indptr[r], b = indptr[r + 1];
    if (b > a) {
        atomicMin(lohi, indices[a]);
        atomicMax(lohi + 1, indices[b - 1]);
    }
}

int pn_k11_col_range(pn_ctx *ctx, int64_t row0, int64_t row1, int64_t *lo, int64_t *hi) {
    PN_API_BEGIN(ctx)
    if (!c->symbolic_ready || row0 < 0 || row1 > c->n1 || row0 > row1 || !lo || !hi) throw ArgError("pn_k11_col_range: bad argument");
    const CsrBlock &A = c->blk[0];
    DevBuf<int32_t> lohi;
    int32_t h[2] = {INT32_MAX, -1};
    lohi.upload(h, 2, c->stream);
    if (row1 > row0) {
        k_row_col_range<<<grid_for(row1 - row0, 256), 256, 0, c->stream>>>(row0, row1, A.indptr.ptr, A.indices.ptr, lohi.ptr);
        c->count_launch();
    }
    PN_CUDA(cudaMemcpyAsync(h, lohi.ptr, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    PN_CUDA(cudaStreamSynchronize(c->stream));
    if (h[1] < 0) { *lo = row0; *hi = row1; }       // empty strip
    else { *lo = h[0]; *hi = (int64_t)h[1] + 1; }
    PN_API_END
}

int pn_dev_block_jacobi_setup(pn_ctx *ctx) {
    PN_API_BEGIN(ctx)
    if (!c->vals_ready) throw ArgError("pn_dev_block_jacobi_setup: call pn_assemble_ke first");
    if (!c->model.symmetric) throw ArgError("pcg: the stiffness matrix is unsymmetric (a shell has kx_mod != ky_mod); use the direct solver");
    pcg_setup_precond(c, c->blk[0].vals.ptr);
    PN_API_END
}

int pn_dev_block_jacobi_nodes(pn_ctx *ctx, int32_t s, int64_t node0, int64_t node1, const double *R_dev, double *Z_dev) {
    PN_API_BEGIN(ctx)
    if (s < 1 || node0 < 0 || node1 > c->model.n_nodes || node0 > node1 || !R_dev || !Z_dev) throw ArgError("pn_dev_block_jacobi_nodes: bad argument");
    pcg_apply_precond_nodes(c, s, node0, node1, R_dev, Z_dev);
    PN_API_END
}

int pn_dev_cg_dots(pn_ctx *ctx, int32_t s, int64_t row0, int64_t row1, const double *A_dev, const double *B_dev, double *out_dev) {
    PN_API_BEGIN(ctx)
    if (!c->pcg || s < 1 || row0 < 0 || row1 > c->n1 || row0 > row1 || !A_dev || !B_dev || !out_dev) throw ArgError("pn_dev_cg_dots: bad argument");
    pcg_strip_dots(c, s, row0, row1, A_dev, B_dev, out_dev);
    PN_API_END
}

int pn_dev_cg_update(pn_ctx *ctx, int32_t s, int64_t node0, int64_t node1, int32_t first, const double *rz_dev,
                     const double *pAp_dev, const double *P_dev, const double *AP_dev, double *X_dev, double *R_dev,
                     double *Z_dev, double *out2_dev) {
    PN_API_BEGIN(ctx)
    if (!c->pcg) throw ArgError("pn_dev_cg_update: call pn_dev_block_jacobi_setup first");
    if (s < 1 || node0 < 0 || node1 > c->model.n_nodes || node0 > node1 || !X_dev || !R_dev || !Z_dev || !out2_dev) throw ArgError("pn_dev_cg_update: bad argument");
    pcg_strip_update(c, s, node0, node1, first, rz_dev, pAp_dev, P_dev, AP_dev, X_dev, R_dev, Z_dev, out2_dev);
    PN_API_END
}

int pn_dev_cg_direction(pn_ctx *ctx, int32_t s, int64_t row0, int64_t row1, int32_t first, const double *rz_dev,
                        const double *rz_new_dev, const double *Z_dev, double *P_dev) {
    PN_API_BEGIN(ctx)
    if (s < 1 || row0 < 0 || row1 > c->n1 || row0 > row1 || !Z_dev || !P_dev) throw ArgError("pn_dev_cg_direction: bad argument");
    pcg_strip_direction(c, s, row0, row1, first, rz_dev, rz_new_dev, Z_dev, P_dev);
    PN_API_END
}

int pn_dev_get_rhs(pn_ctx *ctx, int32_t combo0, int32_t s, double *B_dev) {
    PN_API_BEGIN(ctx)
    if (combo0 < 0 || s < 1 || combo0 + s > c->loads.n_combo || !B_dev) throw ArgError("pn_dev_get_rhs: bad combo range");
    ensure_assembled(c, c->symbolic_ready ? c->sym_flags : 0u);
    ensure_loads(c);
    build_rhs(c, combo0, s, nullptr);
    PN_CUDA(cudaMemcpyAsync(B_dev, c->B.ptr, sizeof(double) * c->n1 * s, cudaMemcpyDeviceToDevice, c->stream));
    PN_API_END
}

int pn_dev_set_solution(pn_ctx *ctx, int32_t combo0, int32_t s, const double *X_dev) {
    PN_API_BEGIN(ctx)
    if (combo0 < 0 || s < 1 || combo0 + s > c->loads.n_combo || !X_dev) throw ArgError("pn_dev_set_solution: bad combo range");
    if (!c->vals_ready) throw ArgError("pn_dev_set_solution: call pn_assemble_ke first");
    c->X.resize(c->n1 * s);
    PN_CUDA(cudaMemcpyAsync(c->X.ptr, X_dev, sizeof(double) * c->n1 * s, cudaMemcpyDeviceToDevice, c->stream));
    c->D_all.resize((int64_t)c->loads.n_combo * c->n_dof);
    merge_displacements(c, combo0, s);
    c->solved = true;
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

int pn_get_displacements(pn_ctx *ctx, int32_t combo, double *D_out) {
    PN_API_BEGIN(ctx)
    if (!c->solved || combo < 0 || combo >= c->loads.n_combo || !D_out) throw ArgError("pn_get_displacements: combo has not been solved");
    PN_CUDA(cudaMemcpyAsync(D_out, c->D_all.ptr + (int64_t)combo * c->n_dof, sizeof(double) * c->n_dof, cudaMemcpyDeviceToHost, c->stream));
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_API_END
}

int pn_get_stats(pn_ctx *ctx, pn_stats *out) {
    PN_API_BEGIN(ctx)
    if (!out) throw ArgError("pn_get_stats: null output");
    c->stats.kernel_launches = c->launches;
    *out = c->stats;
    PN_API_END
}

static const char *kSlotNames[PS_COUNT] = {
    "element_ke", "scatter_add", "gather_blocks", "build_rhs", "spmm", "front_assemble", "extend_add", "lu_diag",
    "lu_panels", "lu_update", "copy_factors", "solve_gather_store", "solve_diag", "solve_update", "solve_misc",
    "reactions", "pcg_vector", "merge_displacements", "assembly_total", "solve_front_fused", "factor_front_fused"};

const char *pn_profile_slot_name(int slot) { return (slot >= 0 && slot < PS_COUNT) ? kSlotNames[slot] : nullptr; }

int pn_profile_enable(pn_ctx *ctx, int on) {
    PN_API_BEGIN(ctx)
    prof_flush(c);
    for (int i = 0; i < PS_COUNT; ++i) { c->prof.ms[i] = 0; c->prof.n[i] = 0; }
    c->prof.on = on != 0;
    PN_API_END
}

int pn_profile_get(pn_ctx *ctx, int slot, double *ms, int64_t *launches) {
    PN_API_BEGIN(ctx)
    if (slot < 0 || slot >= PS_COUNT) throw ArgError("pn_profile_get: bad slot");
    prof_flush(c);
    if (ms) *ms = c->prof.ms[slot];
    if (launches) *launches = c->prof.n[slot];
    PN_API_END
}

int pn_timer_start(pn_ctx *ctx) {
    PN_API_BEGIN(ctx)
    if (!c->tm0) { PN_CUDA(cudaEventCreate(&c->tm0)); PN_CUDA(cudaEventCreate(&c->tm1)); }
    PN_CUDA(cudaStreamSynchronize(c->stream));
    PN_CUDA(cudaEventRecord(c->tm0, c->stream));
    PN_API_END
}

int pn_timer_stop(pn_ctx *ctx, double *ms_out) {
    PN_API_BEGIN(ctx)
    if (!c->tm0) throw ArgError("pn_timer_stop: pn_timer_start was not called");
    PN_CUDA(cudaEventRecord(c->tm1, c->stream));
    PN_CUDA(cudaEventSynchronize(c->tm1));
    float ms = 0;
    PN_CUDA(cudaEventElapsedTime(&ms, c->tm0, c->tm1));
    if (ms_out) *ms_out = ms;
    PN_API_END
}

int pn_reset_stats(pn_ctx *ctx) {
    PN_API_BEGIN(ctx)
    c->stats = pn_stats{};
    c->launches = 0;
    for (int i = 0; i < PC_COUNT; ++i) c->ctr[i] = 0;
    PN_API_END
}

static const char *kCounterNames[PC_COUNT] = {
    "lu_update", "lu_update_later_groups", "lu_strip", "split_levels", "sweep_smem", "sweep_ll", "sweep_step",
    "max_front_pivots", "levels", "factorisations", "pdelta_stationary_batches", "pdelta_fallback_columns",
    "factor_fused_levels", "dist_exchanges", "dist_cut_level", "sweep_overlapped", "lookahead_groups", "sweep_ll_gemm",
    "dof_classes", "dof_class_rejects"};

const char *pn_counter_name(int index) { return (index >= 0 && index < PC_COUNT) ? kCounterNames[index] : nullptr; }

int pn_get_counter(pn_ctx *ctx, const char *name, int64_t *out) {
    PN_API_BEGIN(ctx)
    if (!name || !out) throw ArgError("pn_get_counter: null argument");
    for (int i = 0; i < PC_COUNT; ++i)
        if (!strcmp(name, kCounterNames[i])) { *out = c->ctr[i]; return PN_OK; }
    throw ArgError(std::string("pn_get_counter: unknown counter ") + name);
    PN_API_END
}

}  // extern "C"
