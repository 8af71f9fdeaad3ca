// Sparse direct solver for K11 X = B: multifrontal LU without pivoting on a symmetric pattern.
//
//   ordering / fronts : pn_direct_symbolic.cpp (nested dissection of the node graph by coordinates), on a worker
//                       thread from the moment the model arrays are set (direct_prefetch)
//   numeric           : level by level up the separator tree; every front is a dense row-major
//                       (ns+nb)^2 matrix in a per-level arena.  Right-looking partial LU in block
//                       steps of NB = 64 pivots, grouped GB = 256 pivots at a time:
//                         diag    k_lu_diag: LU of the NB x NB diagonal block + inverses of its triangles (Linv, Uinv)
//                         panels  k_lu_panels: L[rest, k] = F[rest, k] Uinv (GEMM); in symmetric mode the same
//                                 kernel writes the row panel U = D L^T, U12 and the pre-multiplied permanent panel
//                                 from its accumulators; general mode: U[k, rest] = Linv F[k, rest] as a second GEMM
//                         strip   k_lu_strip: rank-NB update of what the rest of the group needs
//                         update  k_lu_update: F[rest, rest] -= L[rest, group] U[group, rest], once per group
//                       All GEMM shapes run through gemm_tile (128 x 64 tile per CTA, DMMA.8x8x4, cp.async ring).
//                       The trailing nb x nb block is the update matrix, extend-added into the
//                       parent in a fixed child order (deterministic, no atomics).  Levels with few fronts are cut
//                       into groups of fronts on side streams.
//   stored factor     : the permanent panels are pre-multiplied by the diagonal-block inverses:
//                         Lp[r, k] = L[r, k] * Linv_kk  (r below block k),  Lp[r, k] = U[r, k] * Uinv_kk (r above)
//                       With y_k = Linv_kk b_k the forward update b_r -= L[r,k] y_k equals
//                       b_r -= Lp[r,k] b_k, i.e. it reads the *old* b_k and can run concurrently with
//                       the diagonal solve; likewise backwards with Uinv.
//   solve             : all right-hand sides at once; forward sweep passes update vectors up the
//                       tree like the update matrices, backward sweep reads ancestors' results.  Per level one of
//                         k_front_forward_sm / k_front_backward_sm  (front's right-hand-side block resident in shared memory),
//                         k_ll_sweep                                 (left-looking persistent kernel with per-block flags),
//                         k_solve_step + gather / extend-add / store (per-step GEMM launches: fallback, > 256 columns).
//
// LU (not Cholesky) because the reference's K is not exactly symmetric: B^T H B is summed in a fixed
// order per element and the orthotropic membrane matrix is unsymmetric when kx_mod != ky_mod
// (Quad3D.py:517-519).  K11 is positive definite for a stable structure, so no pivoting is needed;
// a zero / non-finite pivot is reported as PN_ERR_SINGULAR (Analysis.py:172, 238-239).
#include <cub/cub.cuh>

#include "pn_internal.h"
#include "pn_direct_symbolic.h"

#include <algorithm>
#include <cmath>
#include <future>
#include <omp.h>
#include <memory>
#include <new>
#include <cstdio>
#include <cstdlib>

// page-locked host memory for the index lists of the analysis (pn_direct_symbolic.h, PinnedAllocator); falls back to
// pageable memory if the driver refuses (the uploads are then staged, nothing else changes)
extern "C" void *pn_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (bytes == 0) bytes = 1;
    if (cudaMallocHost(&p, bytes + 16) != cudaSuccess) {
        cudaGetLastError();
        p = malloc(bytes + 16);
        if (!p) throw std::bad_alloc();
        *static_cast<uint64_t *>(p) = 0;
    } else {
        *static_cast<uint64_t *>(p) = 1;
    }
    return static_cast<char *>(p) + 16;
}
extern "C" void pn_host_free(void *q) {
    if (!q) return;
    char *p = static_cast<char *>(q) - 16;
    if (*reinterpret_cast<uint64_t *>(p) == 1) cudaFreeHost(p); else free(p);
}

namespace pn {

static inline unsigned grid_for(int64_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

constexpr int NB = 64;        // pivots per block step
constexpr int BM = 128;       // GEMM tile rows
constexpr int BN = 64;        // GEMM tile columns
constexpr int GEMM_SLICES = 3;             // ring of 16-wide k slices in shared memory
constexpr int BK = 16 * GEMM_SLICES;
constexpr int SA_LD = BK + 8; // = 8 (mod 16) doubles, SB_LD = 2 (mod 8): conflict-free 16-byte fragment loads (see gemm_tile)
constexpr int SB_LD = BN + 2;
constexpr int GEMM_SMEM = (BM * SA_LD + BK * SB_LD) * (int)sizeof(double);   // 82 688 B, two CTAs per SM
constexpr int DIAG_LD = NB + 1;
constexpr int DIAG_SMEM = (3 * NB * DIAG_LD + 5 * NB + 2) * (int)sizeof(double);      // 102 416 B, two CTAs per SM
// leading dimension of a front matrix in its arena: n rounded up to even, so that every row starts 16-byte aligned
// (16-byte cp.async / vector loads in gemm_tile)
__host__ __device__ __forceinline__ int front_ld(int n) { return (n + 1) & ~1; }
constexpr int GB = 256;      // pivots per group: the trailing matrix is streamed once per group (rank-GB update)

struct FrontDev {
    int64_t f_off;      // front matrix offset in its level arena
    int64_t w_off;      // solve workspace offset in its level arena ((ns+nb) x s)
    int64_t l_off;      // permanent pre-multiplied panels: (ns+nb) x ns, ld = ns
    int64_t u_off;      // permanent U12: ns x nb, ld = nb
    int64_t inv_off;    // diagonal-block inverses: per block step [Linv | Uinv], NB x NB each
    int64_t idx_off;    // into fidx / fpos
    int64_t rel_off;    // into rel
    int32_t ns, nb, parent, child_rank;
};

// Subtree-to-GPU distribution of the elimination tree (pn_set_distributed).  Fronts are level-major and, inside a level,
// in post-order, so the fronts of one subtree form a contiguous run on every level: rank r owns fronts [lo[l], hi[l]) of
// every level l <= cut and everything above the cut is replicated.
struct DistPlan {
    bool active = false;                      // world > 1 and the tree is wide enough
    int cut = -1;                             // level of the subtree roots
    std::vector<int32_t> lo, hi;              // [n_levels] this rank's front range per level (all fronts above the cut)
    std::vector<int32_t> cut_lo, cut_hi;      // [world] ranges of the cut level per rank
    std::vector<int64_t> um_off;              // per cut-level front (index inside the level): offset of its update matrix in its owner's segment
    std::vector<int64_t> uv_off;              // ... of its update-vector rows (to be scaled by s)
    int64_t um_seg = 0, uv_seg = 0;           // segment sizes (doubles; uv in rows)
    bool um_tri = false;                      // symmetric mode: update matrices travel as packed lower triangles
    std::vector<int64_t> pos_lo, pos_hi;      // [world] elimination-position range whose solved rows the rank owns
    int64_t x_seg = 0;                        // max rows of a rank's range
    DevBuf<int64_t> d_um_off, d_uv_off;
    DevBuf<int32_t> d_cut_owner;
    DevBuf<double> send, recv;
    // K11 rows this rank's sweeps read from a right-hand side: the pivots of its own subtrees and of every front above
    // the cut.  The refinement residual is computed on these rows only (all other rows of the array are overwritten by
    // the solve: they arrive with the other ranks' solved rows).
    DevBuf<int32_t> res_rows;
    int64_t n_res_rows = 0;
};

struct DirectSolver {
    DirectSym sym;
    DistPlan dist;
    DevBuf<int32_t> pos_dof;                  // K11 row at every elimination position
    bool analysed = false;
    int n_levels = 0;
    std::vector<int32_t> level_ptr;           // fronts are stored level-major on the device
    std::vector<int32_t> level_max_ns, level_max_n, level_max_nb, level_max_children;
    std::vector<int64_t> level_f_total, level_n_total;   // sum n^2, sum n per level
    std::vector<FrontDev> h_fronts;           // level-major
    DevBuf<FrontDev> fronts;
    DevBuf<int32_t> fidx, fpos, rel, perm_pos, dof_front_lm;   // dof_front in level-major numbering
    // K11 entries grouped by owning front (level-major id, so also by level): front f is assembled from entries
    // [front_eptr[f], front_eptr[f+1]) -- a level (or a rank's range of it) touches its own entries only
    DevBuf<int32_t> ent_front;                // per grouped entry: owning front
    DevBuf<uint32_t> ent_dst;                 // offset inside the front matrix
    DevBuf<uint32_t> ent_src;                 // position in the K11 value array
    std::vector<int64_t> front_eptr;          // [n_fronts + 1] (host)
    DevBuf<int32_t> ent_front_raw;            // scratch of the grouping, kept so that a re-analysis allocates nothing
    DevBuf<uint32_t> ent_dst_raw, ent_iota;
    DevBuf<int64_t> d_front_eptr;
    DevBuf<double> arena[2];
    DevBuf<double> warena[2];
    DevBuf<double> yarena;                    // per-front solved own rows of the level being processed
    DevBuf<int32_t> child_ptr, child_idx;     // children of every front (level-major ids, child-rank order)
    DevBuf<FrontDev> fronts_scaled;           // fronts with w_off scaled by the current number of right-hand sides
    int fronts_s = 0;
    DevBuf<double> L, U, inv;
    bool attrs_set = false;
    DevBuf<int> status;                       // 0 ok, 1 zero / non-finite pivot, 2 entry outside front
    bool shared_factor = false;               // the stored factor belongs to the matrix shared by all columns (Ke)
    int64_t factor_epoch = -1;                // Ctx::assembly_epoch the stored factor was computed from
    const double *factor_vals = nullptr;
    DevBuf<double> col_in, col_out;           // per-column solves (P-Delta)
    cudaStream_t fstream[4] = {nullptr, nullptr, nullptr, nullptr};   // factorisation: chains of front groups side by side
    cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
    // look-ahead: the part of a group's trailing update that the next group does not need runs on its own stream
    cudaStream_t ustream[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_prio[4] = {nullptr, nullptr, nullptr, nullptr}, ev_rem[4] = {nullptr, nullptr, nullptr, nullptr};
    int n_fstreams = 0, split_max_fronts = 1 << 30;
    // forward sweep of the first solve running one level behind the factorisation on its own (low-priority) stream
    cudaStream_t sstream = nullptr;
    cudaEvent_t ev_level[64] = {}, ev_sweep = nullptr;
    ~DirectSolver() {
        if (sstream) cudaStreamDestroy(sstream);
        for (int q = 0; q < 64; ++q) if (ev_level[q]) cudaEventDestroy(ev_level[q]);
        if (ev_sweep) cudaEventDestroy(ev_sweep);
        for (int q = 0; q < 4; ++q) {
            if (fstream[q]) cudaStreamDestroy(fstream[q]);
            if (ustream[q]) cudaStreamDestroy(ustream[q]);
            if (ev_join[q]) cudaEventDestroy(ev_join[q]);
            if (ev_prio[q]) cudaEventDestroy(ev_prio[q]);
            if (ev_rem[q]) cudaEventDestroy(ev_rem[q]);
        }
        if (ev_fork) cudaEventDestroy(ev_fork);
    }
    SymVec<int32_t> h_dof_front_lm;           // page-locked staging of dof_front in level-major numbering
    std::vector<int32_t> g_vstart, g_adj;     // vertex graph of the last analysis (host)
    std::vector<uint8_t> g_row_class;         // DOF-type class of every K11 row when the DOF types fall into classes
    int n_classes = 1;                        // DOF-type classes of the current ordering (pn_direct_symbolic.h, DofClasses)
    bool classes_rejected = false;            // the assembled pattern joined two predicted classes: one class from now on
    std::vector<double> g_coords;
    std::vector<int64_t> g_adj_ptr;
    VertexGraphScratch g_scratch;
    std::future<void> host_future;            // host half of the analysis running ahead on a worker thread
    int host_state = 0;                       // 0 none, 1 worker running, 2 d.sym valid for the current model
    int64_t prefetch_n1 = -1;
    DevBuf<uint32_t> ll_flags;                // left-looking sweeps: one flag per (pivot block, column chunk)
    uint32_t ll_epoch = 0;
    int ll_grid = 0;                          // resident CTAs of k_ll_sweep (all units' owners must be co-resident)
    int64_t arena_cap = 0;
};

// ------------------------------------------------------------------------------------------------
// symbolic maps on the device
// ------------------------------------------------------------------------------------------------
__global__ void k_entry_rows(int64_t rows, const int32_t *__restrict__ indptr, int32_t *__restrict__ ent_row) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) ent_row[k] = (int32_t)r;
}

__device__ __forceinline__ int front_find(const int32_t *__restrict__ list, int n, int32_t pos) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (list[mid] < pos) lo = mid + 1; else hi = mid;
    }
    return (lo < n && list[lo] == pos) ? lo : -1;
}

__global__ void k_entry_maps(int64_t nnz, const int32_t *__restrict__ ent_row, const int32_t *__restrict__ indices,
                             const int32_t *__restrict__ perm_pos, const int32_t *__restrict__ dof_front,
                             const FrontDev *__restrict__ fronts, const int32_t *__restrict__ fpos,
                             int32_t *__restrict__ ent_front, uint32_t *__restrict__ ent_dst, int *__restrict__ status) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    int32_t i = ent_row[k], j = indices[k];
    int32_t pi = perm_pos[i], pj = perm_pos[j];
    int32_t f = pi < pj ? dof_front[i] : dof_front[j];
    FrontDev F = fronts[f];
    int n = F.ns + F.nb;
    const int32_t *list = fpos + F.idx_off;
    int li = front_find(list, n, pi), lj = front_find(list, n, pj);
    if (li < 0 || lj < 0) { atomicMax(status, 2); li = lj = 0; }
    ent_front[k] = f;
    ent_dst[k] = (uint32_t)li * (uint32_t)front_ld(n) + (uint32_t)lj;
}

// ------------------------------------------------------------------------------------------------
// numeric factorisation kernels
// ------------------------------------------------------------------------------------------------
// entries [k0, k1) of the front-grouped entry list: one store each (every K11 entry has its own destination)
__global__ void k_assemble_entries(int64_t k0, int64_t k1, const int32_t *__restrict__ ent_front, const uint32_t *__restrict__ ent_dst,
                                   const uint32_t *__restrict__ ent_src, const FrontDev *__restrict__ fronts,
                                   const double *__restrict__ vals, double *__restrict__ arena) {
    int64_t k = k0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= k1) return;
    arena[fronts[ent_front[k]].f_off + ent_dst[k]] = vals[ent_src[k]];
}
__global__ void k_iota_u32(int64_t n, uint32_t *__restrict__ out) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) out[k] = (uint32_t)k;
}
__global__ void k_gather_by_u32(int64_t n, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ in, uint32_t *__restrict__ out) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) out[k] = in[perm[k]];
}
// eptr[f] = first position of the sorted front list with value >= f, f = 0 .. nf
__global__ void k_front_entry_ptr(int nf, int64_t n, const int32_t *__restrict__ sorted_front, int64_t *__restrict__ eptr) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f > nf) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (sorted_front[mid] < f) lo = mid + 1; else hi = mid;
    }
    eptr[f] = lo;
}

// ---- subtree-to-GPU exchanges (DistPlan) ---------------------------------------------------------------------------
// update matrix (trailing nb x nb block, row-major with ld = n) of cut-level fronts <-> contiguous nb x nb segments.
// grid = (front of the range, row chunk).  pack: arena -> buf;  unpack: buf -> arena.
__global__ void k_dist_um(int first, const FrontDev *__restrict__ fronts, int cut_first, const int64_t *__restrict__ um_off,
                          double *__restrict__ arena, double *__restrict__ buf, int unpack, int sym) {
    const int q = first + blockIdx.x;
    const FrontDev F = fronts[q];
    const int n = F.ns + F.nb;
    double *seg = buf + um_off[q - cut_first];
    for (int a = blockIdx.y; a < F.nb; a += gridDim.y) {
        double *row = arena + F.f_off + (int64_t)(F.ns + a) * front_ld(n) + F.ns;
        double *srow = seg + (sym ? (int64_t)a * (a + 1) / 2 : (int64_t)a * F.nb);
        const int bend = sym ? a + 1 : F.nb;       // symmetric mode keeps the lower triangle only
        if (unpack) for (int b = threadIdx.x; b < bend; b += blockDim.x) row[b] = srow[b];
        else        for (int b = threadIdx.x; b < bend; b += blockDim.x) srow[b] = row[b];
    }
}
// update vectors: rows [ns, n) x s of the fronts' solve workspaces <-> contiguous segments
__global__ void k_dist_uv(int first, int s, const FrontDev *__restrict__ fronts /* w_off scaled by s */, int cut_first,
                          const int64_t *__restrict__ uv_off, double *__restrict__ W, double *__restrict__ buf, int unpack) {
    const int q = first + blockIdx.x;
    const FrontDev F = fronts[q];
    double *w = W + F.w_off + (int64_t)F.ns * s;
    double *seg = buf + uv_off[q - cut_first] * s;
    const int64_t total = (int64_t)F.nb * s;
    for (int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.y * blockDim.x) {
        if (unpack) w[t] = seg[t]; else seg[t] = w[t];
    }
}
// solved rows by elimination position: X[pos_dof[p]] <-> buf[p - p0], p in [p0, p1)
__global__ void k_dist_x(int64_t p0, int64_t p1, int s, const int32_t *__restrict__ pos_dof, double *__restrict__ X,
                         double *__restrict__ buf, int unpack) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (p1 - p0) * s) return;
    const int64_t p = p0 + t / s;
    const int j = (int)(t % s);
    double *x = X + (int64_t)pos_dof[p] * s + j;
    if (unpack) *x = buf[t]; else buf[t] = *x;
}

// Symmetric mode reads the strict upper triangle of a front only where the transposed panels wrote it, so only the lower
// triangle (with the diagonal) is zeroed before assembly: half the bytes of a memset of the whole arena.
// grid = (front, row chunk); a warp per row, 16-byte stores.
__global__ void k_zero_lower(int first, const FrontDev *__restrict__ fronts, double *__restrict__ arena) {
    const FrontDev F = fronts[first + blockIdx.x];
    const int n = F.ns + F.nb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = blockIdx.y * nw + warp; r < n; r += gridDim.y * nw) {
        double *row = arena + F.f_off + (int64_t)r * front_ld(n);
        // (kept general: peels if a row start is not 16-byte aligned)
        // f_off + r ld may be odd: peel to 16-byte alignment, then double2 stores
        int c = 0;
        if ((reinterpret_cast<uintptr_t>(row) & 15) != 0) { if (lane == 0) row[0] = 0.0; c = 1; }
        const int end = r + 1;
        for (int j = c + 2 * lane; j + 1 < end; j += 64) *reinterpret_cast<double2 *>(row + j) = make_double2(0.0, 0.0);
        if (((end - c) & 1) && lane == 0) row[end - 1] = 0.0;
    }
}

// Adds the update matrix of every child with the given rank into its parent (one level up).
// grid.x = child index within the child level, grid.y = row chunk of the nb x nb update matrix; a thread block walks
// rows, threads walk columns (coalesced reads of the child row; the parent columns rel[b] are ascending runs).
__global__ void k_extend_add(int first_child, int n_children, int rank, const FrontDev *__restrict__ fronts,
                             const int32_t *__restrict__ rel, const double *__restrict__ child_arena,
                             double *__restrict__ parent_arena, int sym) {
    int ci = blockIdx.x;
    if (ci >= n_children) return;
    FrontDev C = fronts[first_child + ci];
    if (C.parent < 0 || C.child_rank != rank || C.nb == 0) return;
    FrontDev P = fronts[C.parent];
    const int nc = front_ld(C.ns + C.nb), np = front_ld(P.ns + P.nb);     // leading dimensions
    const int32_t *r = rel + C.rel_off;
    for (int a = blockIdx.y; a < C.nb; a += gridDim.y) {
        const double *src = child_arena + C.f_off + (int64_t)(C.ns + a) * nc + C.ns;
        double *dst = parent_arena + P.f_off + (int64_t)r[a] * np;
        // symmetric mode maintains the lower triangle only (rel is ascending, so it maps into the parent's lower triangle)
        const int bend = sym ? a + 1 : C.nb;
        for (int b = threadIdx.x; b < bend; b += blockDim.x) dst[r[b]] += src[b];
    }
}

extern __shared__ double pn_dyn_smem[];

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc, bool pred) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    int bytes = pred ? 8 : 0;                 // src-size 0: nothing is read, the 8 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(gsrc), "r"(bytes));
}
// 16-byte copy (SASS LDGSTS.E.BYPASS.128); the first `bytes` (0, 8 or 16) come from global memory, the rest is zero-filled
__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc, int bytes) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// D(8x8) = A(8x4) B(4x8) + C on the FP64 tensor path (DMMA.8x8x4).  Fragment layout (PTX ISA, m8n8k4 .f64):
// a: row = lane/4, k = lane%4;  b: k = lane%4, col = lane/4;  c/d: row = lane/4, cols 2 (lane%4) + {0, 1}.
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// C[m x n] (op)= A[m x k] B[k x n], row-major with leading dimensions; one BM x BN tile per call,
// 256 threads = 8 warps in a 4 x 2 grid, each warp owning a 32 x 32 block as 4 x 4 DMMA tiles
// (tools/microbench/fp64_pipes.cu: DMMA 37.0 vs DFMA 36.7 TFLOP/s peak on B200 -- the tensor path is used
// because a fragment needs 1/6 of the shared-memory bytes the FMA register tile needs, not for its peak).
// mode 0: C -= A B;  mode 1: C = A B.
// Round-2 layout (tools/microbench/gemm_variants.cu, profiles/r2_gemm_variants.md: 22 -> 27 TFLOP/s on the rank-256
// update of the 4 500-row fronts; the ablation there puts the remaining distance to the 33 TFLOP/s of the same loop
// without copies on the global -> shared copies):
//   * operands stream through a ring of three 16-wide k slices filled by 16-byte cp.async (SASS LDGSTS.E.BYPASS.128)
//     when the leading dimensions are even and the operand origins 16-byte aligned, by 8-byte copies otherwise;
//   * fragments are read with 16-byte shared loads: the DMMA contracts over the k index held by lane%4, and WHICH four
//     k of a slice those are is free as long as A and B agree, so lane tg takes k = 8 kp + 2 tg + e for two consecutive
//     DMMAs (e = 0, 1): one LDS.128 of A[row][8 kp + 2 tg .. +1] feeds both.  Likewise the 8 columns of a DMMA tile need
//     not be adjacent: tile ni = 2 np + e2 holds columns 16 np + 2 g + e2, so one LDS.128 of B[k][16 np + 2 g .. +1] feeds
//     tiles 2 np and 2 np + 1, and a lane's accumulators of a tile pair are the four consecutive columns 16 np + 4 tg + {0..3}
//     (32-byte global loads / stores of C).  Shared leading dimensions: A = 8 (mod 16) doubles, B = 2 (mod 8): the eight
//     16-byte accesses of a quarter-warp fall into eight different bank groups;
//   * mode 0 keeps -C in the accumulators (negated on load and on store: exact), so the loop has no sign multiply.
// In-place use (C aliasing A or B) is safe when the tile covers the aliased operand's full extent in the other
// dimension: every global read is complete (wait_group 0 + barrier) before the first store.
// Warps whose 32 x 32 block lies outside the matrix, or (lower_only: C is a symmetric trailing matrix of which only the
// lower triangle is kept, rows and columns counted from the same origin) strictly above the diagonal, take part in the
// copies and barriers but skip their fragment loads and DMMAs: on the small fronts of the bottom levels (trailing
// blocks of 144 ... 1140 rows in 128 x 64 tiles) that removes more than half of the tensor work.
__device__ __forceinline__ void gemm_tile(const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc,
                                          int m, int n, int k, int tile_m, int tile_n, int mode, bool lower_only = false,
                                          double *Ct = nullptr, int64_t ldct = 0, const double *ct_scale = nullptr, int64_t ct_scale_ld = 0,
                                          double *Ct2 = nullptr, int64_t ldct2 = 0, int ct2_row0 = 0) {
    double *sA = pn_dyn_smem;                      // [BM][SA_LD]
    double *sB = pn_dyn_smem + BM * SA_LD;         // [BK][SB_LD]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
    const int row0 = tile_m * BM, col0 = tile_n * BN;
    const int wrow = (warp >> 1) * 32, wcol = (warp & 1) * 32;
    // acc[mi][2 np + e2][e1] = (mode 0: minus) C[row0 + wrow + 8 mi + g][col0 + wcol + 16 np + 4 tg + 2 e1 + e2]
    double acc[4][4][2];
    const bool warp_on = row0 + wrow < m && col0 + wcol < n && !(lower_only && row0 + wrow + 31 < col0 + wcol);
    const bool full_tile = row0 + BM <= m && col0 + BN <= n;
    const bool c16 = full_tile && (ldc & 1) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
    if (mode != 0 || !warp_on) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    } else if (c16) {
        const double *c0p = C + (int64_t)(row0 + wrow + g) * ldc + col0 + wcol + 4 * tg;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                const double2 lo = *reinterpret_cast<const double2 *>(c0p + (int64_t)(8 * mi) * ldc + 16 * np);
                const double2 hi = *reinterpret_cast<const double2 *>(c0p + (int64_t)(8 * mi) * ldc + 16 * np + 2);
                acc[mi][2 * np][0] = -lo.x; acc[mi][2 * np + 1][0] = -lo.y;
                acc[mi][2 * np][1] = -hi.x; acc[mi][2 * np + 1][1] = -hi.y;
            }
    } else {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const int gr = row0 + wrow + 8 * mi + g;
#pragma unroll
            for (int np = 0; np < 2; ++np)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int gc = col0 + wcol + 16 * np + 4 * tg + q;
                    acc[mi][2 * np + (q & 1)][q >> 1] = (gr < m && gc < n) ? -C[(int64_t)gr * ldc + gc] : 0.0;
                }
        }
    }
    // k runs through a ring of GEMM_SLICES 16-wide slices of the shared tiles: slices sl + 1, sl + 2 are in flight while
    // slice sl is multiplied, one barrier per slice.
    // The DMMA holds the issue port of its scheduler for its whole duration (tools/microbench/dmma_occupancy.cu:
    // every other instruction costs ~0.85 cycles of tensor time, at any occupancy), so the copy loop issues as few
    // instructions as it can: per copy one 32-bit offset from a per-slice base (IMAD), one IMAD.WIDE, one LDGSTS; the
    // row / column predicates are decided once per tile.
    // Rows >= m of the A tile and columns >= n of the B tile only feed accumulators that are never stored, so in
    // full slices their copies are clamped to a valid row / column instead of being predicated.
    const int nsl = (k + 15) >> 4;
    const int a_last = m - 1 - row0;                        // last valid tile row
    const int ncol = n - col0;                              // valid columns of the B tile
    const uint32_t lda32 = (uint32_t)lda, ldb32 = (uint32_t)ldb;
    const bool a16 = (lda & 1) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
    const bool b16 = (ldb & 1) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0;
    // 16-byte copies.  A slice: 128 rows x 8 chunks, rows ra + 32 i (i < 4), k offset ca;  B slice: 16 k rows x 32 chunks,
    // k rows kr + 8 i (i < 2), column cb (clamped to the last valid chunk, of which `bbytes` are inside the matrix)
    const int ra = tid >> 3, ca = (tid & 7) * 2;
    const int kr = tid >> 5, cb = (tid & 31) * 2;
    const int cbc = min(cb, (ncol - 1) & ~1);
    const int bbytes = min(16, (ncol - cbc) * 8);
    const double *Abase = A + (int64_t)row0 * lda + ca;
    const double *Bbase = B + col0 + cbc;
    auto issue_a16 = [&](int slot, int kbase) {
        const double *sa = Abase + kbase;
        double *ta = sA + ra * SA_LD + 16 * slot + ca;
        if (kbase + 16 <= k) {
#pragma unroll
            for (int i = 0; i < 4; ++i) cp_async16(ta + 32 * i * SA_LD, sa + (uint32_t)min(ra + 32 * i, a_last) * lda32, 16);
        } else {                                            // last, partial slice: zero-fill k >= k_valid
            const int ab = max(0, min(16, (k - kbase - ca) * 8));
#pragma unroll
            for (int i = 0; i < 4; ++i) cp_async16(ta + 32 * i * SA_LD, ab ? sa + (uint32_t)min(ra + 32 * i, a_last) * lda32 : A, ab);
        }
    };
    auto issue_b16 = [&](int slot, int kbase) {
        const double *sb = Bbase + (int64_t)kbase * ldb;
        double *tb = sB + (16 * slot + kr) * SB_LD + cb;
        if (kbase + 16 <= k) {
#pragma unroll
            for (int i = 0; i < 2; ++i) cp_async16(tb + 8 * i * SB_LD, sb + (uint32_t)(kr + 8 * i) * ldb32, bbytes);
        } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool ok = kbase + kr + 8 * i < k;
                cp_async16(tb + 8 * i * SB_LD, ok ? sb + (uint32_t)(kr + 8 * i) * ldb32 : B, ok ? bbytes : 0);
            }
        }
    };
    // 8-byte copies into the same layout (odd leading dimension or origin): A rows ra8 + 16 i (i < 8), k column ka8;
    // B k rows kr8 + 4 i (i < 4), column cb8
    const int ra8 = tid >> 4, ka8 = tid & 15, kr8 = tid >> 6, cb8 = tid & 63;
    const int cbc8 = min(cb8, ncol - 1);
    auto issue_a8 = [&](int slot, int kbase) {
        const bool oka = kbase + ka8 < k;
        const double *sa = A + (int64_t)row0 * lda + kbase + ka8;
        double *ta = sA + ra8 * SA_LD + 16 * slot + ka8;
#pragma unroll
        for (int i = 0; i < 8; ++i) cp_async8(ta + 16 * i * SA_LD, oka ? sa + (uint32_t)min(ra8 + 16 * i, a_last) * lda32 : A, oka);
    };
    auto issue_b8 = [&](int slot, int kbase) {
        const double *sb = B + col0 + cbc8 + (int64_t)kbase * ldb;
        double *tb = sB + (16 * slot + kr8) * SB_LD + cb8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool ok = kbase + kr8 + 4 * i < k;
            cp_async8(tb + 4 * i * SB_LD, ok ? sb + (uint32_t)(kr8 + 4 * i) * ldb32 : B, ok);
        }
    };
    auto issue = [&](int sl) {
        if (sl < nsl) {
            const int slot = sl % GEMM_SLICES, kbase = sl << 4;
            if (a16) issue_a16(slot, kbase); else issue_a8(slot, kbase);
            if (b16) issue_b16(slot, kbase); else issue_b8(slot, kbase);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < GEMM_SLICES - 1; ++i) issue(i);
    for (int sl = 0; sl < nsl; ++sl) {
        cp_async_wait<GEMM_SLICES - 2>();
        __syncthreads();
        issue(sl + GEMM_SLICES - 1);
        const int base = (sl % GEMM_SLICES) << 4;
        if (warp_on)
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
            double2 a[4], b[2][2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = *reinterpret_cast<const double2 *>(sA + (wrow + 8 * mi + g) * SA_LD + base + 8 * kp + 2 * tg);
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int np = 0; np < 2; ++np)
                    b[e][np] = *reinterpret_cast<const double2 *>(sB + (base + 8 * kp + 2 * tg + e) * SB_LD + wcol + 16 * np + 2 * g);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    dmma884(acc[mi][2 * np][0], acc[mi][2 * np][1], a[mi].x, b[0][np].x);
                    dmma884(acc[mi][2 * np + 1][0], acc[mi][2 * np + 1][1], a[mi].x, b[0][np].y);
                }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    dmma884(acc[mi][2 * np][0], acc[mi][2 * np][1], a[mi].y, b[1][np].x);
                    dmma884(acc[mi][2 * np + 1][0], acc[mi][2 * np + 1][1], a[mi].y, b[1][np].y);
                }
        }
    }
    cp_async_wait<0>();
    __syncthreads();          // the tiles may be refilled by the caller's next gemm_tile call
    if (!warp_on) return;
    if (mode == 0) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = -acc[mi][ni][0]; acc[mi][ni][1] = -acc[mi][ni][1]; }
    }
    if (Ct) {
        // second, transposed and column-scaled copy of the result: Ct[col][row] = scale[col] * C[row][col] (the row
        // panel U = D L^T of the symmetric factorisation, written from the accumulators instead of by a transpose
        // kernel); the eight lanes that share a column write eight consecutive rows = one 64-byte segment
#pragma unroll
        for (int np = 0; np < 2; ++np)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gc = col0 + wcol + 16 * np + 4 * tg + q;
                if (gc >= n) continue;
                const double sc = ct_scale[(int64_t)gc * ct_scale_ld];
                double *dst = Ct + (int64_t)gc * ldct;
                double *dst2 = Ct2 ? Ct2 + (int64_t)gc * ldct2 - ct2_row0 : nullptr;   // rows >= ct2_row0 also go to the permanent U12
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    const int gr = row0 + wrow + 8 * mi + g;
                    if (gr < m) {
                        const double v = sc * acc[mi][2 * np + (q & 1)][q >> 1];
                        dst[gr] = v;
                        if (dst2 && gr >= ct2_row0) dst2[gr] = v;
                    }
                }
            }
    }
    if (c16) {
        double *c0p = C + (int64_t)(row0 + wrow + g) * ldc + col0 + wcol + 4 * tg;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                *reinterpret_cast<double2 *>(c0p + (int64_t)(8 * mi) * ldc + 16 * np) = make_double2(acc[mi][2 * np][0], acc[mi][2 * np + 1][0]);
                *reinterpret_cast<double2 *>(c0p + (int64_t)(8 * mi) * ldc + 16 * np + 2) = make_double2(acc[mi][2 * np][1], acc[mi][2 * np + 1][1]);
            }
        return;
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int gr = row0 + wrow + 8 * mi + g;
        if (gr >= m) continue;
        double *crow = C + (int64_t)gr * ldc;
#pragma unroll
        for (int np = 0; np < 2; ++np)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gc = col0 + wcol + 16 * np + 4 * tg + q;
                if (gc < n) crow[gc] = acc[mi][2 * np + (q & 1)][q >> 1];
            }
    }
}

// LU (no pivoting) of one NB x NB diagonal block, then the inverses of its unit-lower and upper triangles.
// The kernel sits on the critical path of every block step (one launch per 64 pivots, a single CTA per front on the
// top levels), so it is organised for latency, not throughput:
//   LU      : 256 threads as a 16 x 16 grid, each owning a 4 x 4 register tile.  Step p: the owners of row p and of
//             column p publish them (and the pivot's reciprocal: MUFU seed + two Newton steps, not an IEEE division)
//             through double-buffered shared memory, ONE barrier, then every thread updates its tile.
//   inverse : the two 32 x 32 diagonal blocks of each triangle by substitution in axpy form (one warp per block, one
//             lane per column, 32 accumulators in registers: the dependent chain is 32 steps long, all other FMAs
//             are independent), then the off-diagonal blocks as two small GEMMs by all 256 threads:
//               inv([[A,0],[C,B]]) = [[Ai,0],[-Bi C Ai, Bi]],   inv([[A,C],[0,B]]) = [[Ai,-Ai C Bi],[0,Bi]].
__device__ __forceinline__ double rcp_newton(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}

// Diagonal block [kb, kb+bs) of one front: load, factor + invert, write back.  All 256 threads of the CTA call it.
#define PN_CLK(i) do { if (clk && threadIdx.x == 0) clk[i] = clock64(); } while (0)
__device__ __forceinline__ void diag_step(const FrontDev &F, int kb, double *arena, double *inv, int *status, int sym,
                                          long long *clk = nullptr) {
    PN_CLK(0);
    double(*S)[DIAG_LD] = reinterpret_cast<double(*)[DIAG_LD]>(pn_dyn_smem);
    double(*Li)[DIAG_LD] = reinterpret_cast<double(*)[DIAG_LD]>(pn_dyn_smem + NB * DIAG_LD);
    double(*Ui)[DIAG_LD] = reinterpret_cast<double(*)[DIAG_LD]>(pn_dyn_smem + 2 * NB * DIAG_LD);
    double *rowb = pn_dyn_smem + 3 * NB * DIAG_LD;      // [2][NB]
    double *colb = rowb + 2 * NB;                       // [2][NB]
    double *rdiag = colb + 2 * NB;                      // [NB] reciprocal pivots
    double *rcpb = rdiag + NB;                          // [2]
    const int bs = min(NB, F.ns - kb);
    const int n = front_ld(F.ns + F.nb);       // leading dimension of the front
    double *A = arena + F.f_off + (int64_t)kb * n + kb;
    // thread (ti, tj) owns rows ti + 16 r and columns tj + 16 c (cyclic: the published row / column and the smem and
    // global accesses are contiguous across the lanes of a half-warp, and every thread stays busy until the end)
    const int t = threadIdx.x, ti = t >> 4, tj = t & 15;
    double a[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = ti + 16 * r, j = tj + 16 * c;
            // symmetric mode: only the lower triangle of the front is maintained; mirror it
            const int si = (sym && j > i) ? j : i, sj = (sym && j > i) ? i : j;
            a[r][c] = (i < bs && j < bs) ? A[(int64_t)si * n + sj] : (i == j ? 1.0 : 0.0);
        }
    PN_CLK(1);
    // rows / columns >= bs are identity padding: their pivot steps are exact no-ops and are skipped (a 16-pivot tail
    // block of a 144-pivot leaf front used to cost a full 64-step chain); their reciprocal pivots are 1
    if (t >= bs && t < NB) rdiag[t] = 1.0;
#pragma unroll
    for (int qb = 0; qb < 4; ++qb) {
        for (int pp = 0; pp < 16; ++pp) {
            const int p = 16 * qb + pp;
            if (p >= bs) break;                  // (uniform; the trip counts stay compile-time constants)
            double *R = rowb + (pp & 1) * NB, *Cc = colb + (pp & 1) * NB;
            if (ti == pp) {
#pragma unroll
                for (int c = 0; c < 4; ++c) R[tj + 16 * c] = a[qb][c];
            }
            if (tj == pp) {
#pragma unroll
                for (int r = 0; r < 4; ++r) Cc[ti + 16 * r] = a[r][qb];
                if (ti == pp) {
                    const double piv = a[qb][qb];
                    if (piv == 0.0 || !(fabs(piv) <= 1.7976931348623157e308)) atomicMax(status, 1);
                    const double rc = rcp_newton(piv);
                    rcpb[pp & 1] = rc;
                    rdiag[p] = rc;
                }
            }
            __syncthreads();
            // branch-free update: rows / columns that are not below / right of the pivot get a zero multiplier
            const double rc = rcpb[pp & 1];
            double l[4], u[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const bool act = (r > qb) || (r == qb && ti > pp);
                l[r] = act ? Cc[ti + 16 * r] * rc : 0.0;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const bool act = (c > qb) || (c == qb && tj > pp);
                u[c] = act ? R[tj + 16 * c] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) a[r][c] = fma(-l[r], u[c], a[r][c]);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const bool act = (tj == pp) && ((r > qb) || (r == qb && ti > pp));
                a[r][qb] = act ? l[r] : a[r][qb];
            }
        }
    }
    PN_CLK(2);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = ti + 16 * r, j = tj + 16 * c;
            S[i][j] = a[r][c];
            if (i < bs && j < bs) A[(int64_t)i * n + j] = a[r][c];
        }
    __syncthreads();
    PN_CLK(3);
    // ---- 32 x 32 diagonal blocks of the two inverses: warp 0/1 = Linv blocks, warp 2/3 = Uinv blocks ----
    const int warp = t >> 5, lane = t & 31;
    if (warp < 2) {
        const int o = 32 * warp;
        double acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            const double xp = acc[p];
#pragma unroll
            for (int i = p + 1; i < 32; ++i) acc[i] = fma(-S[o + i][o + p], xp, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) Li[o + i][o + lane] = acc[i];
    } else if (warp < 4) {
        const int o = 32 * (warp - 2);
        double acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int p = 31; p >= 0; --p) {
            const double xp = acc[p] * rdiag[o + p];
            acc[p] = xp;
#pragma unroll
            for (int i = 0; i < p; ++i) acc[i] = fma(-S[o + i][o + p], xp, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) Ui[o + i][o + lane] = acc[i];
    }
    __syncthreads();
    PN_CLK(4);
    // ---- off-diagonal blocks: T = C Ai (lower) / C Bi (upper), then X = -Bi T / -Ai T ----
    // thread = (matrix m, row group w of 8 rows, column lane): the B operand row is read contiguously across the
    // lanes, the A operand as warp-wide broadcasts
    const int m = t >> 7, r0 = 8 * ((t >> 5) & 3);
    if (bs <= 32) {
        // the off-diagonal blocks of a block with at most 32 real pivots are zero (identity padding): no GEMMs
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (m == 0) Li[32 + r0 + r][lane] = 0.0; else Ui[r0 + r][32 + lane] = 0.0;
        }
        __syncthreads();
    } else {
    {
        double acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = 0.0;
        if (m == 0) {
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                const double bv = Li[k][lane];
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] = fma(S[32 + r0 + r][k], bv, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) Li[r0 + r][32 + lane] = acc[r];          // scratch: the (zero) upper-right block
        } else {
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                const double bv = Ui[32 + k][32 + lane];
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] = fma(S[r0 + r][32 + k], bv, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) Ui[32 + r0 + r][lane] = acc[r];          // scratch: the (zero) lower-left block
        }
    }
    __syncthreads();
    {
        double acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = 0.0;
        if (m == 0) {
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                const double bv = Li[k][32 + lane];
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] = fma(-Li[32 + r0 + r][32 + k], bv, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) Li[32 + r0 + r][lane] = acc[r];
        } else {
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                const double bv = Ui[32 + k][lane];
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] = fma(-Ui[r0 + r][k], bv, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) Ui[r0 + r][32 + lane] = acc[r];
        }
    }
    __syncthreads();
    }
    PN_CLK(5);
    double *out = inv + F.inv_off + (int64_t)(kb / NB) * (2 * NB * NB);
    for (int e = t; e < NB * NB; e += blockDim.x) {
        const int i = e / NB, j = e % NB;
        out[e] = (j > i) ? 0.0 : Li[i][j];               // the scratch blocks are not part of the result
        out[NB * NB + e] = (j < i) ? 0.0 : Ui[i][j];
    }
    PN_CLK(6);
}

__global__ void __launch_bounds__(256, 2)
k_lu_diag(int first, int kb, const FrontDev *__restrict__ fronts, double *__restrict__ arena, double *__restrict__ inv,
          int *__restrict__ status, int sym, long long *clk) {
    FrontDev F = fronts[first + blockIdx.x];
    if (kb >= F.ns) return;
    diag_step(F, kb, arena, inv, status, sym, blockIdx.x == 0 ? clk : nullptr);
}

// Panels of block step kb as GEMMs with the diagonal inverses.  grid.x = front; grid.y < tiles_u: column
// tile of the row panel U[k, rest] = Linv F[k, rest]; otherwise row tile of the column panel
// L[rest, k] = F[rest, k] Uinv.  Both are in place (see gemm_tile).
__global__ void __launch_bounds__(256, 2)
k_lu_panels(int first, int kb, int tiles_u, const FrontDev *__restrict__ fronts, double *__restrict__ arena,
            const double *__restrict__ inv, int sym, double *__restrict__ U12, double *__restrict__ Lst) {
    FrontDev F = fronts[first + blockIdx.x];
    if (kb >= F.ns) return;
    const int bs = min(NB, F.ns - kb);
    const int nn = F.ns + F.nb, n = front_ld(nn);       // n = leading dimension from here on
    const int r0 = kb + bs, rest = nn - r0;
    if (rest <= 0) return;
    double *A = arena + F.f_off;
    const double *Linv = inv + F.inv_off + (int64_t)(kb / NB) * (2 * NB * NB);
    const double *Uinv = Linv + NB * NB;
    if ((int)blockIdx.y < tiles_u) {
        if (sym) return;                       // the row panel is the scaled transpose of the column panel
        int t = blockIdx.y;
        if (t * BN >= rest) return;
        double *P = A + (int64_t)kb * n + r0;
        gemm_tile(Linv, NB, P, n, P, n, bs, rest, bs, 0, t, 1);
    } else {
        int t = blockIdx.y - tiles_u;
        if (t * BM >= rest) return;
        double *P = A + (int64_t)r0 * n + kb;
        if (sym)      // U[kb + j, r0 + i] = pivot_j * L[r0 + i, kb + j]
            gemm_tile(P, n, Uinv, NB, P, n, rest, bs, bs, t, 0, 1, false, A + (int64_t)kb * n + r0, n, A + (int64_t)kb * n + kb, n + 1,
                      F.nb ? U12 + F.u_off + (int64_t)kb * F.nb : nullptr, F.nb, F.ns - r0);
        else
            gemm_tile(P, n, Uinv, NB, P, n, rest, bs, bs, t, 0, 1);
        // permanent, pre-multiplied copy of the tile for the sweeps: Lp[r, k] = L[r, k] Linv_kk, from the tile this CTA
        // has just written (L1 / L2 hits) instead of a second pass over the arena at the end of the level
        __syncthreads();
        gemm_tile(P, n, Linv, NB, Lst + F.l_off + (int64_t)r0 * F.ns + kb, F.ns, rest, bs, bs, t, 0, 1);
    }
}

// Symmetric mode: U[k, c] = D_k L[c, k]^T (D_k = pivots of block k), so the row panel of step kb is the scaled
// transpose of the column panel just computed.  grid = (front, 32-row tile of the column panel); block 32 x 8.
__global__ void k_lu_transpose_panel(int first, int kb, const FrontDev *__restrict__ fronts, double *__restrict__ arena) {
    __shared__ double tile[32][33];
    FrontDev F = fronts[first + blockIdx.x];
    if (kb >= F.ns) return;
    const int bs = min(NB, F.ns - kb);
    const int nn = F.ns + F.nb, n = front_ld(nn);
    const int r0 = kb + bs, rest = nn - r0;
    const int c0 = blockIdx.y * 32;
    if (c0 >= rest) return;
    double *A = arena + F.f_off;
    for (int h = 0; h < bs; h += 32) {         // the panel is bs wide: two 32-column halves
        for (int y = threadIdx.y; y < 32; y += 8) {
            int row = c0 + y, col = h + threadIdx.x;
            tile[y][threadIdx.x] = (row < rest && col < bs) ? A[(int64_t)(r0 + row) * n + kb + col] : 0.0;
        }
        __syncthreads();
        for (int y = threadIdx.y; y < 32; y += 8) {
            int pi = h + y, col = c0 + threadIdx.x;
            if (pi < bs && col < rest) A[(int64_t)(kb + pi) * n + r0 + col] = A[(int64_t)(kb + pi) * n + kb + pi] * tile[threadIdx.x][y];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Whole factorisation of one front inside one CTA, for the levels with many small fronts (the bottom of the tree:
// 8192 ... 512 fronts of 270 ... 666 rows in config 3).  The level-synchronous path launches diag / panels / strip /
// update / store kernels per 64 pivots over ALL fronts of the level, and every launch streams the level's arena
// (gigabytes) through HBM for GEMMs that are all prologue and epilogue.  Here a CTA walks the block steps of its own
// front back to back with the same device functions (diag_step, gemm_tile): the front (<= 0.6 ... 3.5 MB, lower
// triangle only in symmetric mode) stays in L2 between steps, there are no launch gaps or inter-kernel dependencies, and
// the grid of thousands of independent CTAs keeps every SM busy.  Same arithmetic in the same order as the per-step
// kernels, so the factor is bit-identical to theirs.
// ------------------------------------------------------------------------------------------------
constexpr int FUSED_FACTOR_SMEM = GEMM_SMEM > DIAG_SMEM ? GEMM_SMEM : DIAG_SMEM;
// a level takes the one-CTA-per-front kernel when it has at least this many fronts.  Measured on config 3 (B200,
// profiles/r2_fused_factor_ab.md): 8192 fronts 10.6 -> 9.9 ms, 4096 fronts 3.0 -> 2.9 ms, but 512 fronts 2.5 -> 3.0 ms and
// 256 fronts 3.2 -> 4.5 ms (a single CTA per 1000-row front is a long serial chain), hence the threshold.
static int fused_factor_min() {
    static const int v = getenv("PN_FUSED_FACTOR_MIN") ? atoi(getenv("PN_FUSED_FACTOR_MIN")) : 4096;
    return v;
}

__global__ void __launch_bounds__(256, 2)
k_front_factor_fused(int first, const FrontDev *__restrict__ fronts, double *__restrict__ arena, double *__restrict__ inv,
                     int *__restrict__ status, int sym, double *__restrict__ U12, double *__restrict__ Lst) {
    const FrontDev F = fronts[first + blockIdx.x];
    const int nn = F.ns + F.nb, n = front_ld(nn);       // n = leading dimension, nn = rows
    double *A = arena + F.f_off;
    for (int g0 = 0; g0 < F.ns; g0 += GB) {
        const int ge = min(g0 + GB, F.ns);
        for (int kb = g0; kb < ge; kb += NB) {
            diag_step(F, kb, arena, inv, status, sym);
            __syncthreads();
            const int bs = min(NB, F.ns - kb);
            const int r0 = kb + bs, rest = nn - r0;
            if (rest <= 0) continue;
            const double *Linv = inv + F.inv_off + (int64_t)(kb / NB) * (2 * NB * NB);
            const double *Uinv = Linv + NB * NB;
            // panels (k_lu_panels)
            if (!sym) {
                double *P = A + (int64_t)kb * n + r0;
                for (int t = 0; t * BN < rest; ++t) gemm_tile(Linv, NB, P, n, P, n, bs, rest, bs, 0, t, 1);
            }
            {
                double *P = A + (int64_t)r0 * n + kb;
                for (int t = 0; t * BM < rest; ++t) {
                    if (sym)
                        gemm_tile(P, n, Uinv, NB, P, n, rest, bs, bs, t, 0, 1, false, A + (int64_t)kb * n + r0, n, A + (int64_t)kb * n + kb, n + 1,
                                  F.nb ? U12 + F.u_off + (int64_t)kb * F.nb : nullptr, F.nb, F.ns - r0);
                    else
                        gemm_tile(P, n, Uinv, NB, P, n, rest, bs, bs, t, 0, 1);
                    __syncthreads();
                    gemm_tile(P, n, Linv, NB, Lst + F.l_off + (int64_t)r0 * F.ns + kb, F.ns, rest, bs, bs, t, 0, 1);
                }
            }
            __syncthreads();
            // strips the rest of the group needs (k_lu_strip)
            if (kb + NB < ge) {
                const int q0 = kb + NB;
                const int wa = ge - q0, ha = nn - q0;
                const int ta_n = (wa + BN - 1) / BN, ta_m = (ha + BM - 1) / BM;
                const double *Lp = A + (int64_t)q0 * n + kb;
                for (int t = 0; t < ta_m * ta_n; ++t) {
                    if (sym && (t / ta_n + 1) * BM <= (t % ta_n) * BN) continue;
                    gemm_tile(Lp, n, A + (int64_t)kb * n + q0, n, A + (int64_t)q0 * n + q0, n, ha, wa, NB, t / ta_n, t % ta_n, 0, sym != 0);
                }
                if (!sym) {
                    const int hb = ge - q0, wb = nn - ge;
                    const int tb_n = (wb + BN - 1) / BN, tb_m = (hb + BM - 1) / BM;
                    for (int t = 0; t < tb_m * tb_n; ++t)
                        gemm_tile(Lp, n, A + (int64_t)kb * n + ge, n, A + (int64_t)q0 * n + ge, n, hb, wb, NB, t / tb_n, t % tb_n, 0);
                }
                __syncthreads();
            }
        }
        // trailing update of the group (k_lu_update)
        const int rest = nn - ge;
        if (rest > 0) {
            const int tn_ = (rest + BN - 1) / BN, tm_ = (rest + BM - 1) / BM;
            for (int tm = 0; tm < tm_; ++tm)
                for (int tn = 0; tn < tn_; ++tn) {
                    if (sym && (tm + 1) * BM <= tn * BN) continue;
                    gemm_tile(A + (int64_t)ge * n + g0, n, A + (int64_t)g0 * n + ge, n, A + (int64_t)ge * n + ge, n, rest, rest, ge - g0, tm, tn, 0,
                              sym != 0);
                }
            __syncthreads();
        }
    }
    // permanent panels above the diagonal blocks (k_store_panels) and, in general mode, U12 (k_copy_u12)
    for (int kb = NB; kb < F.ns; kb += NB) {
        const int bs = min(NB, F.ns - kb);
        const double *Uinv = inv + F.inv_off + (int64_t)(kb / NB) * (2 * NB * NB) + NB * NB;
        for (int t = 0; t * BM < kb; ++t) gemm_tile(A + kb, n, Uinv, NB, Lst + F.l_off + kb, F.ns, kb, bs, bs, t, 0, 1);
    }
    if (!sym && F.nb) {
        const int64_t total = (int64_t)F.ns * F.nb;
        for (int64_t t = threadIdx.x; t < total; t += blockDim.x) {
            const int64_t r = t / F.nb;
            U12[F.u_off + t] = A[r * n + F.ns + (t - r * F.nb)];
        }
    }
}

// Block steps are grouped GB pivots at a time so that the (HBM-resident) trailing matrix is read and written once
// per group instead of once per step.  Inside group [g0, ge), after the panels of step kb, only the strips the
// remaining steps of the group will factor are updated (rank NB):
//   (a) F[r0:n, r0:ge]  -= L[r0:n, kb:r0] U[kb:r0, r0:ge]      (r0 = kb + NB)
//   (b) F[r0:ge, ge:n]  -= L[r0:ge, kb:r0] U[kb:r0, ge:n]
// grid = (front, tile); tiles enumerate region (a) row-major, then region (b).
__global__ void __launch_bounds__(256, 2)
k_lu_strip(int first, int kb, int g0, const FrontDev *__restrict__ fronts, double *__restrict__ arena, int sym, int front_on_x) {
    FrontDev F = fronts[first + (front_on_x ? blockIdx.x : blockIdx.y)];
    const int ge = min(g0 + GB, F.ns);
    const int r0 = kb + NB;
    if (r0 >= ge) return;
    const int nn = F.ns + F.nb, n = front_ld(nn);       // n = leading dimension, nn = rows
    double *A = arena + F.f_off;
    const int wa = ge - r0, ha = nn - r0;
    const int ta_n = (wa + BN - 1) / BN, ta_m = (ha + BM - 1) / BM;
    const int hb = ge - r0, wb = nn - ge;
    const int tb_n = (wb + BN - 1) / BN, tb_m = (hb + BM - 1) / BM;
    int t = front_on_x ? blockIdx.y : blockIdx.x;
    const double *Lp = A + (int64_t)r0 * n + kb;        // L[r0:, kb:r0]
    if (t < ta_m * ta_n) {
        if (sym && (t / ta_n + 1) * BM <= (t % ta_n) * BN) return;      // tile strictly above the diagonal
        gemm_tile(Lp, n, A + (int64_t)kb * n + r0, n, A + (int64_t)r0 * n + r0, n, ha, wa, NB, t / ta_n, t % ta_n, 0, sym != 0);
        return;
    }
    if (sym) return;                           // region (b) is the upper side
    t -= ta_m * ta_n;
    if (t < tb_m * tb_n)
        gemm_tile(Lp, n, A + (int64_t)kb * n + ge, n, A + (int64_t)r0 * n + ge, n, hb, wb, NB, t / tb_n, t % tb_n, 0);
}

// trailing update after group [g0, ge): F[ge:, ge:] -= L[ge:, g0:ge] U[g0:ge, ge:]   (rank ge - g0 <= GB)
__global__ void __launch_bounds__(256, 2)
k_lu_update(int first, int g0, const FrontDev *__restrict__ fronts, double *__restrict__ arena, int sym, int front_on_x, int tn0) {
    // front slowest (grid z) when the level has at most 65535 fronts: the CTAs of one front run together and its
    // panels stay in L2; otherwise the front index needs the x dimension.  tn0 = first column tile of this launch
    // (look-ahead: the next group's columns and the remainder are separate launches).
    const int tcol = tn0 + (front_on_x ? blockIdx.z : blockIdx.x);
    FrontDev F = fronts[first + (front_on_x ? blockIdx.x : blockIdx.z)];
    if (g0 >= F.ns) return;
    const int ge = min(g0 + GB, F.ns);
    const int nn = F.ns + F.nb, n = front_ld(nn);       // n = leading dimension, nn = rows
    const int rest = nn - ge;
    if ((int)blockIdx.y * BM >= rest || tcol * BN >= rest) return;
    if (sym && ((int)blockIdx.y + 1) * BM <= tcol * BN) return;      // tile strictly above the diagonal
    double *A = arena + F.f_off;
    gemm_tile(A + (int64_t)ge * n + g0, n, A + (int64_t)g0 * n + ge, n, A + (int64_t)ge * n + ge, n, rest, rest, ge - g0, blockIdx.y,
              tcol, 0, sym != 0);
}

// Permanent factor of a level: for block column k of every front, rows above the diagonal block hold
// U[., k] Uinv_kk and rows below hold L[., k] Linv_kk (see the header comment).  grid = (front, block column,
// row tile); a row tile index first enumerates the tiles above the diagonal block, then those below.
__global__ void __launch_bounds__(256, 2)
k_store_panels(int first, const FrontDev *__restrict__ fronts, const double *__restrict__ arena,
               const double *__restrict__ inv, double *__restrict__ L) {
    FrontDev F = fronts[first + blockIdx.x];
    const int kb = blockIdx.y * NB;
    if (kb >= F.ns) return;
    const int bs = min(NB, F.ns - kb);
    const int nn = F.ns + F.nb, n = front_ld(nn);
    const int up_rows = kb, lo_rows = nn - kb - bs;
    const int up_tiles = (up_rows + BM - 1) / BM, lo_tiles = (lo_rows + BM - 1) / BM;
    int t = blockIdx.z;
    const double *Linv = inv + F.inv_off + (int64_t)blockIdx.y * (2 * NB * NB);
    const double *A = arena + F.f_off;
    double *out = L + F.l_off;
    (void)lo_tiles;      // the rows below the diagonal block are stored by k_lu_panels
    if (t < up_tiles) gemm_tile(A + kb, n, Linv + NB * NB, NB, out + kb, F.ns, up_rows, bs, bs, t, 0, 1);
}

// U12 of every front of a level to permanent storage
__global__ void k_copy_u12(int first, const FrontDev *__restrict__ fronts, const double *__restrict__ arena,
                           double *__restrict__ U) {
    FrontDev F = fronts[first + blockIdx.x];
    int n = front_ld(F.ns + F.nb);
    int64_t total = (int64_t)F.ns * F.nb;
    const double *A = arena + F.f_off;
    for (int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.y * blockDim.x) {
        int64_t r = t / F.nb; int cidx = (int)(t - r * F.nb);
        U[F.u_off + t] = A[r * n + F.ns + cidx];
    }
}

// ------------------------------------------------------------------------------------------------
// solve kernels (W = per-front workspace (ns+nb) x s, row-major; Y = solved own rows, same offsets)
// ------------------------------------------------------------------------------------------------
// mode 0 (forward): own rows <- X[fidx], boundary rows <- 0
// mode 1 (backward): own rows <- X[fidx] (holds y), boundary rows <- X[fidx] (final x of ancestors)
__global__ void k_solve_gather(int first, int s, int mode, const FrontDev *__restrict__ fronts,
                               const int32_t *__restrict__ fidx, const double *__restrict__ X, double *__restrict__ W) {
    FrontDev F = fronts[first + blockIdx.x];
    int n = F.ns + F.nb;
    int64_t total = (int64_t)n * s;
    for (int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.y * blockDim.x) {
        int r = (int)(t / s), j = (int)(t - (int64_t)r * s);
        double v = 0.0;
        if (r < F.ns || mode == 1) v = X[(int64_t)fidx[F.idx_off + r] * s + j];
        W[F.w_off + t] = v;
    }
}
__global__ void k_solve_store(int first, int s, const FrontDev *__restrict__ fronts, const int32_t *__restrict__ fidx,
                              const double *__restrict__ Y, double *__restrict__ X) {
    FrontDev F = fronts[first + blockIdx.x];
    int64_t total = (int64_t)F.ns * s;
    for (int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.y * blockDim.x) {
        int r = (int)(t / s), j = (int)(t - (int64_t)r * s);
        X[(int64_t)fidx[F.idx_off + r] * s + j] = Y[F.w_off + t];
    }
}
__global__ void k_solve_extend_add(int first_child, int n_children, int rank, int s, const FrontDev *__restrict__ fronts,
                                   const int32_t *__restrict__ rel, const double *__restrict__ Wc, double *__restrict__ Wp) {
    int ci = blockIdx.x;
    if (ci >= n_children) return;
    FrontDev C = fronts[first_child + ci];
    if (C.parent < 0 || C.child_rank != rank || C.nb == 0) return;
    FrontDev P = fronts[C.parent];
    const int32_t *r = rel + C.rel_off;
    int64_t total = (int64_t)C.nb * s;
    for (int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.y * blockDim.x) {
        int a = (int)(t / s), j = (int)(t - (int64_t)a * s);
        Wp[P.w_off + (int64_t)r[a] * s + j] += Wc[C.w_off + (int64_t)(C.ns + a) * s + j];
    }
}
// One block step of a sweep for every front of a level; grid = (front, row tile, column tile of s).
//   forward : tile 0: Y[k] = Linv_kk W[k];  tile t >= 1: W[r] -= Lp[r, k] W[k] for the rows below block k
//   backward: tile 0: Y[k] = Uinv_kk W[k];  tile t >= 1: W[r] -= Lp[r, k] W[k] for the rows above block k
// Every tile reads the same, unmodified W[k] (the solved block goes to Y), so one launch does both.
__global__ void __launch_bounds__(256, 2)
k_solve_step(int first, int kb, int s, int backward, const FrontDev *__restrict__ fronts, const double *__restrict__ L,
             const double *__restrict__ inv, double *__restrict__ W, double *__restrict__ Y) {
    FrontDev F = fronts[first + blockIdx.x];
    if (kb >= F.ns) return;
    const int bs = min(NB, F.ns - kb);
    const int n = F.ns + F.nb;
    if ((int)blockIdx.z * BN >= s) return;
    double *w = W + F.w_off;
    const double *wk = w + (int64_t)kb * s;
    const double *Linv = inv + F.inv_off + (int64_t)(kb / NB) * (2 * NB * NB);
    int t = blockIdx.y;
    if (t == 0) {
        gemm_tile(backward ? Linv + NB * NB : Linv, NB, wk, s, Y + F.w_off + (int64_t)kb * s, s, bs, s, bs, 0, blockIdx.z, 1);
        return;
    }
    --t;
    if (!backward) {
        const int r0 = kb + bs, rest = n - r0;
        if (t * BM >= rest) return;
        gemm_tile(L + F.l_off + (int64_t)r0 * F.ns + kb, F.ns, wk, s, w + (int64_t)r0 * s, s, rest, s, bs, t, blockIdx.z, 0);
    } else {
        if (t * BM >= kb) return;
        gemm_tile(L + F.l_off + kb, F.ns, wk, s, w, s, kb, s, bs, t, blockIdx.z, 0);
    }
}
// Whole forward / backward sweep of one front inside one CTA (levels with many fronts): gather from X, pull the
// children's update vectors (forward), all block steps back to back, store.  The front's workspace is touched by
// this CTA only, so between steps it stays in L2 and the level costs one launch instead of 2 + children + steps.
constexpr int FUSED_SOLVE_MIN_FRONTS = 128;

__global__ void __launch_bounds__(256, 2)
k_front_forward(int first, int s, const FrontDev *__restrict__ fronts, const int32_t *__restrict__ fidx,
                const int32_t *__restrict__ rel, const int32_t *__restrict__ child_ptr, const int32_t *__restrict__ child_idx,
                const double *__restrict__ L, const double *__restrict__ inv, double *__restrict__ X, double *__restrict__ W,
                const double *__restrict__ Wchild, double *__restrict__ Y) {
    const int f = first + blockIdx.x;
    FrontDev F = fronts[f];
    const int n = F.ns + F.nb;
    double *w = W + F.w_off;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < n; r += 8) {
        const double *src = r < F.ns ? X + (int64_t)fidx[F.idx_off + r] * s : nullptr;
        for (int j = lane; j < s; j += 32) w[(int64_t)r * s + j] = src ? src[j] : 0.0;
    }
    __syncthreads();
    for (int ci = child_ptr[f]; ci < child_ptr[f + 1]; ++ci) {
        FrontDev C = fronts[child_idx[ci]];
        const int32_t *rc = rel + C.rel_off;
        const double *wc = Wchild + C.w_off + (int64_t)C.ns * s;
        for (int a = warp; a < C.nb; a += 8) {
            double *dst = w + (int64_t)rc[a] * s;
            for (int j = lane; j < s; j += 32) dst[j] += wc[(int64_t)a * s + j];
        }
        __syncthreads();
    }
    const int gz = (s + BN - 1) / BN;
    double *y = Y + F.w_off;
    for (int kb = 0; kb < F.ns; kb += NB) {
        const int bs = min(NB, F.ns - kb);
        const double *Linv = inv + F.inv_off + (int64_t)(kb / NB) * (2 * NB * NB);
        const double *wk = w + (int64_t)kb * s;
        const int r0 = kb + bs, rest = n - r0;
        for (int tn = 0; tn < gz; ++tn) {
            gemm_tile(Linv, NB, wk, s, y + (int64_t)kb * s, s, bs, s, bs, 0, tn, 1);
            for (int tm = 0; tm * BM < rest; ++tm)
                gemm_tile(L + F.l_off + (int64_t)r0 * F.ns + kb, F.ns, wk, s, w + (int64_t)r0 * s, s, rest, s, bs, tm, tn, 0);
        }
        __syncthreads();
    }
    for (int r = warp; r < F.ns; r += 8) {
        double *dst = X + (int64_t)fidx[F.idx_off + r] * s;
        for (int j = lane; j < s; j += 32) dst[j] = y[(int64_t)r * s + j];
    }
}

__global__ void __launch_bounds__(256, 2)
k_front_backward(int first, int s, const FrontDev *__restrict__ fronts, const int32_t *__restrict__ fidx,
                 const double *__restrict__ L, const double *__restrict__ U, const double *__restrict__ inv,
                 double *__restrict__ X, double *__restrict__ W, double *__restrict__ Y) {
    FrontDev F = fronts[first + blockIdx.x];
    const int n = F.ns + F.nb;
    double *w = W + F.w_off;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < n; r += 8) {
        const double *src = X + (int64_t)fidx[F.idx_off + r] * s;
        for (int j = lane; j < s; j += 32) w[(int64_t)r * s + j] = src[j];
    }
    __syncthreads();
    const int gz = (s + BN - 1) / BN;
    if (F.nb) {
        for (int tn = 0; tn < gz; ++tn)
            for (int tm = 0; tm * BM < F.ns; ++tm)
                gemm_tile(U + F.u_off, F.nb, w + (int64_t)F.ns * s, s, w, s, F.ns, s, F.nb, tm, tn, 0);
        __syncthreads();
    }
    double *y = Y + F.w_off;
    for (int kb = ((F.ns - 1) / NB) * NB; kb >= 0; kb -= NB) {
        const int bs = min(NB, F.ns - kb);
        const double *Uinv = inv + F.inv_off + (int64_t)(kb / NB) * (2 * NB * NB) + NB * NB;
        const double *wk = w + (int64_t)kb * s;
        for (int tn = 0; tn < gz; ++tn) {
            gemm_tile(Uinv, NB, wk, s, y + (int64_t)kb * s, s, bs, s, bs, 0, tn, 1);
            for (int tm = 0; tm * BM < kb; ++tm)
                gemm_tile(L + F.l_off + kb, F.ns, wk, s, w, s, kb, s, bs, tm, tn, 0);
        }
        __syncthreads();
    }
    for (int r = warp; r < F.ns; r += 8) {
        double *dst = X + (int64_t)fidx[F.idx_off + r] * s;
        for (int j = lane; j < s; j += 32) dst[j] = y[(int64_t)r * s + j];
    }
}

// ------------------------------------------------------------------------------------------------
// Sweeps with the front's right-hand-side block resident in shared memory.
// A CTA owns (front, chunk of CW right-hand-side columns): the (ns+nb) x CW block w lives in shared memory for the
// whole sweep of the front, so the only global traffic is the gather from X / the children, ONE streaming pass over
// the front's stored panels, and the result rows.  Warps own 16-row slabs of the panel; the DMMA A fragments are read
// straight from global memory (every quad of lanes reads one full 32-byte sector) into registers, one 32-wide k
// chunk ahead of the math; the B fragments (rows of w for the pivot block) and the accumulators come from shared
// memory with leading dimension CW + 4 (conflict-free fragment loads, see gemm_tile).
// Used for every level whose largest front fits (n + 32) * (CW + 4) * 8 bytes <= 227 KB for some CW in {32, 16, 8}.
// ------------------------------------------------------------------------------------------------
constexpr int SW_PAD_ROWS = 32;       // zeroed rows after the block: k chunks may run past the last valid row
constexpr int SW_MAX_SMEM = 232448;   // 227 KB
constexpr int SW_TWO_CTA_SMEM = 113 * 1024;

// acc += (NEG ? -1 : 1) * A[0:16, 0:kc] * Bs[0:kc, 0:CW]   for one 16-row slab.  A: global, row-major, leading
// dimension lda; rows >= rows_valid and columns >= kc read as zero.  Bs: shared, leading dimension CW + 4.
template <int CW, bool NEG>
__device__ __forceinline__ void slab_mma(double (&acc)[2][CW / 8][2], const double *__restrict__ A, int64_t lda, int rows_valid,
                                         int kc, const double *Bs, int lane) {
    constexpr int LDW = CW + 4, NT = CW / 8;
    const int g = lane >> 2, tg = lane & 3;
    const bool v0 = g < rows_valid, v1 = g + 8 < rows_valid;
    const double *a0p = A + (int64_t)g * lda + tg;
    const double *a1p = A + (int64_t)(g + 8) * lda + tg;
    constexpr int KS = (CW >= 32) ? 4 : 8;          // k steps (of 4) per prefetched chunk: registers vs distance
    constexpr int KCH = 4 * KS;
    const int nch = (kc + KCH - 1) / KCH;
    double cur[2][KS], nxt[2][KS];
    auto load = [&](double (&d)[2][KS], int ch) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int k = KCH * ch + 4 * ks;
            const bool kv = k + tg < kc;
            d[0][ks] = (v0 && kv) ? __ldg(a0p + k) : 0.0;
            d[1][ks] = (v1 && kv) ? __ldg(a1p + k) : 0.0;
        }
    };
    load(cur, 0);
    for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) load(nxt, ch + 1);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const double *brow = Bs + (KCH * ch + 4 * ks + tg) * LDW + g;
            double b[NT];
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) b[ni] = brow[8 * ni];
            const double a0 = NEG ? -cur[0][ks] : cur[0][ks], a1 = NEG ? -cur[1][ks] : cur[1][ks];
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) {
                dmma884(acc[0][ni][0], acc[0][ni][1], a0, b[ni]);
                dmma884(acc[1][ni][0], acc[1][ni][1], a1, b[ni]);
            }
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) { cur[0][ks] = nxt[0][ks]; cur[1][ks] = nxt[1][ks]; }
    }
}

// accumulators <-> rows [r, r + 16) of the shared block (c fragment: row g / g + 8, columns 2 tg, 2 tg + 1 of every n tile)
template <int CW>
__device__ __forceinline__ void slab_load_c(double (&acc)[2][CW / 8][2], const double *w, int r, int n, int lane) {
    constexpr int LDW = CW + 4, NT = CW / 8;
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
        const int row = r + 8 * mi + g;
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {
            double2 v = row < n ? *reinterpret_cast<const double2 *>(w + row * LDW + 8 * ni + 2 * tg) : make_double2(0.0, 0.0);
            acc[mi][ni][0] = v.x; acc[mi][ni][1] = v.y;
        }
    }
}
template <int CW>
__device__ __forceinline__ void slab_store_c(const double (&acc)[2][CW / 8][2], double *w, int r, int n, int lane) {
    constexpr int LDW = CW + 4, NT = CW / 8;
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
        const int row = r + 8 * mi + g;
        if (row >= n) continue;
#pragma unroll
        for (int ni = 0; ni < NT; ++ni)
            *reinterpret_cast<double2 *>(w + row * LDW + 8 * ni + 2 * tg) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
    }
}
// solved rows [i0, i0 + 16) of pivot block kb straight to X
template <int CW>
__device__ __forceinline__ void slab_store_x(const double (&acc)[2][CW / 8][2], double *__restrict__ X, const int32_t *__restrict__ fi,
                                             int kb, int i0, int bs, int s, int c0, int lane) {
    constexpr int NT = CW / 8;
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
        const int i = i0 + 8 * mi + g;
        if (i >= bs) continue;
        double *dst = X + (int64_t)fi[kb + i] * s + c0;
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {
            const int col = 8 * ni + 2 * tg;
            if (c0 + col < s) dst[col] = acc[mi][ni][0];
            if (c0 + col + 1 < s) dst[col + 1] = acc[mi][ni][1];
        }
    }
}

// w[r][0:CW] = X[fi[r]][c0 : c0 + CW] for r < rows_x, zero for rows_x <= r < rows_all; four rows per thread in flight
template <int CW>
__device__ __noinline__ void sweep_gather(double *w, const double *__restrict__ X, const int32_t *__restrict__ fi, int rows_x,
                                             int rows_all, int s, int c0, int tid) {
    constexpr int LDW = CW + 4, RPP = 256 / CW;
    const int j = tid % CW, a0 = tid / CW;
    const bool cv = c0 + j < s;
    for (int base = 0; base < rows_all; base += 4 * RPP) {
        int32_t src[4];
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = base + q * RPP + a0;
            src[q] = (cv && r < rows_x) ? fi[r] : -1;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = src[q] >= 0 ? X[(int64_t)src[q] * s + c0 + j] : 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = base + q * RPP + a0;
            if (r < rows_all) w[r * LDW + j] = v[q];
        }
    }
}

// w[rc[a]][0:CW] += wc[a][0:CW] for the nb boundary rows of one child; four rows per thread in flight
template <int CW>
__device__ __noinline__ void sweep_add_child(double *w, const double *__restrict__ wc, const int32_t *__restrict__ rc, int nb, int s,
                                             int c0, int tid) {
    constexpr int LDW = CW + 4, RPP = 256 / CW;
    const int j = tid % CW, a0 = tid / CW;
    const bool cv = c0 + j < s;
    for (int base = 0; base < nb; base += 4 * RPP) {
        double v[4];
        int dst[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int a = base + q * RPP + a0;
            const bool ok = cv && a < nb;
            dst[q] = ok ? rc[a] : -1;
            v[q] = ok ? wc[(int64_t)a * s + j] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (dst[q] >= 0) w[dst[q] * LDW + j] += v[q];
    }
}

template <int CW, int MINB = 2, int PAD = SW_PAD_ROWS>
__global__ void __launch_bounds__(256, MINB)
k_front_forward_sm(int first, int s, int H, const FrontDev *__restrict__ fronts, const int32_t *__restrict__ fidx,
                   const int32_t *__restrict__ rel, const int32_t *__restrict__ child_ptr, const int32_t *__restrict__ child_idx,
                   const double *__restrict__ L, const double *__restrict__ inv, double *__restrict__ X,
                   double *__restrict__ Wout, const double *__restrict__ Wchild) {
    constexpr int LDW = CW + 4, NT = CW / 8;
    double *w = pn_dyn_smem;
    // 1-D grid, column chunk fastest: the CTAs that stream the same front's panels are scheduled together (second read from L2)
    const int f = first + blockIdx.x / H;
    const FrontDev F = fronts[f];
    const int n = F.ns + F.nb, c0 = (blockIdx.x % H) * CW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t *fi = fidx + F.idx_off;
    // own rows from X, boundary rows and the padding rows zero
    sweep_gather<CW>(w, X, fi, F.ns, n + PAD, s, c0, tid);
    __syncthreads();
    // update vectors of the children, in child-rank order (fixed summation order)
    for (int ci = child_ptr[f]; ci < child_ptr[f + 1]; ++ci) {
        const FrontDev C = fronts[child_idx[ci]];
        const int32_t *rc = rel + C.rel_off;
        const double *wc = Wchild + C.w_off + (int64_t)C.ns * s + c0;
        sweep_add_child<CW>(w, wc, rc, C.nb, s, c0, tid);
        __syncthreads();
    }
    for (int kb = 0; kb < F.ns; kb += NB) {
        const int bs = min(NB, F.ns - kb);
        const double *Linv = inv + F.inv_off + (int64_t)(kb / NB) * (2 * NB * NB);
        const int r0 = kb + bs, rest = n - r0;
        const int dslabs = (bs + 15) >> 4, nslabs = dslabs + ((rest + 15) >> 4);
        const double *Bs = w + kb * LDW;
        for (int sl = warp; sl < nslabs; sl += 8) {
            double acc[2][NT][2];
            if (sl < dslabs) {                 // y_k = Linv_kk w_k  -> X
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NT; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
                slab_mma<CW, false>(acc, Linv + (int64_t)(16 * sl) * NB, NB, NB - 16 * sl, bs, Bs, lane);   // columns >= bs are zero
                slab_store_x<CW>(acc, X, fi, kb, 16 * sl, bs, s, c0, lane);
            } else {                           // w[r] -= Lp[r, k] w_k  (old w_k: the panel is pre-multiplied by Linv_kk)
                const int r = r0 + 16 * (sl - dslabs);
                slab_load_c<CW>(acc, w, r, n, lane);
                slab_mma<CW, true>(acc, L + F.l_off + (int64_t)r * F.ns + kb, F.ns, n - r, bs, Bs, lane);
                slab_store_c<CW>(acc, w, r, n, lane);
            }
        }
        __syncthreads();
    }
    // update vector for the parent
    double *wo = Wout + F.w_off + (int64_t)F.ns * s + c0;
    for (int e = tid; e < F.nb * CW; e += 256) {
        const int a = e / CW, j = e % CW;
        if (c0 + j < s) wo[(int64_t)a * s + j] = w[(F.ns + a) * LDW + j];
    }
}

template <int CW, int MINB = 2, int PAD = SW_PAD_ROWS>
__global__ void __launch_bounds__(256, MINB)
k_front_backward_sm(int first, int s, int H, const FrontDev *__restrict__ fronts, const int32_t *__restrict__ fidx,
                    const double *__restrict__ L, const double *__restrict__ U, const double *__restrict__ inv,
                    double *__restrict__ X) {
    constexpr int LDW = CW + 4, NT = CW / 8;
    double *w = pn_dyn_smem;
    const FrontDev F = fronts[first + blockIdx.x / H];
    const int n = F.ns + F.nb, c0 = (blockIdx.x % H) * CW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t *fi = fidx + F.idx_off;
    // own rows hold y, boundary rows the final x of the ancestors
    sweep_gather<CW>(w, X, fi, n, n + PAD, s, c0, tid);
    __syncthreads();
    if (F.nb) {                                // w[0:ns] -= U12 w[ns:n]
        const int nslabs = (F.ns + 15) >> 4;
        for (int sl = warp; sl < nslabs; sl += 8) {
            double acc[2][NT][2];
            const int r = 16 * sl;
            slab_load_c<CW>(acc, w, r, F.ns, lane);
            slab_mma<CW, true>(acc, U + F.u_off + (int64_t)r * F.nb, F.nb, F.ns - r, F.nb, w + F.ns * LDW, lane);
            slab_store_c<CW>(acc, w, r, F.ns, lane);
        }
        __syncthreads();
    }
    for (int kb = ((F.ns - 1) / NB) * NB; kb >= 0; kb -= NB) {
        const int bs = min(NB, F.ns - kb);
        const double *Uinv = inv + F.inv_off + (int64_t)(kb / NB) * (2 * NB * NB) + NB * NB;
        const int dslabs = (bs + 15) >> 4, nslabs = dslabs + ((kb + 15) >> 4);
        const double *Bs = w + kb * LDW;
        for (int sl = warp; sl < nslabs; sl += 8) {
            double acc[2][NT][2];
            if (sl < dslabs) {                 // x_k = Uinv_kk w_k  -> X
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NT; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
                slab_mma<CW, false>(acc, Uinv + (int64_t)(16 * sl) * NB, NB, NB - 16 * sl, bs, Bs, lane);   // columns >= bs are zero
                slab_store_x<CW>(acc, X, fi, kb, 16 * sl, bs, s, c0, lane);
            } else {                           // w[r] -= Lp[r, k] w_k for the rows above block k
                const int r = 16 * (sl - dslabs);
                slab_load_c<CW>(acc, w, r, kb, lane);
                slab_mma<CW, true>(acc, L + F.l_off + (int64_t)r * F.ns + kb, F.ns, kb - r, bs, Bs, lane);
                slab_store_c<CW>(acc, w, r, kb, lane);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Left-looking sweeps for the levels with few, large fronts (the per-step launches of the general path cost one
// launch latency per 64 pivots: 47 dependent launches for a 3000-pivot root separator).
// One persistent kernel per level and direction.  A unit = (64-row tile of a front, 32-column chunk of the right-hand
// sides); the CTA that owns a unit keeps the tile in its accumulators and subtracts panel x w_k for EVERY earlier
// pivot block k in one long k loop (the tile's rows of the stored panels are streamed exactly once, nothing is
// read-modify-written), waiting for block k's rows through a per-block flag in global memory.  When its own rows are
// final it publishes them (release store of the sweep's epoch number) and only then multiplies by the diagonal
// inverse.  Units are numbered tile-major and every CTA takes units in increasing order with all CTAs resident, so
// the lowest unfinished unit never waits on an unscheduled one.  Summation order is fixed (k ascending / descending).
// ------------------------------------------------------------------------------------------------
constexpr int LL_MAX_CHUNKS = 8;          // 32-column chunks: up to 256 right-hand sides
__device__ __forceinline__ void flag_wait(const uint32_t *flag, uint32_t epoch, int lane) {
    if (lane == 0) {
        uint32_t v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v < epoch) __nanosleep(40);
        } while (v < epoch);
    }
    __syncwarp();
}

// acc -= A[0:16, 0:kc] * Bg[0:kc, 0:16]  (A: global row-major, ld lda;  Bg: global row-major, ld ldb, read through
// L2 only: the rows were written by another CTA of this kernel)
__device__ __forceinline__ void slab_mma_g(double (&acc)[2][2][2], const double *__restrict__ A, int64_t lda, int rows_valid, int kc,
                                           const double *Bg, int64_t ldb, int cols_valid, int lane, bool neg) {
    const int g = lane >> 2, tg = lane & 3;
    const bool v0 = g < rows_valid, v1 = g + 8 < rows_valid;
    const bool c0v = g < cols_valid, c1v = g + 8 < cols_valid;
    const double *a0p = A + (int64_t)g * lda + tg;
    const double *a1p = A + (int64_t)(g + 8) * lda + tg;
    const double *bp = Bg + (int64_t)tg * ldb + g;
    for (int kb = 0; kb < kc; kb += 16) {
        double a0[4], a1[4], b0[4], b1[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int k = kb + 4 * ks;
            const bool kv = k + tg < kc;
            a0[ks] = (v0 && kv) ? __ldg(a0p + k) : 0.0;
            a1[ks] = (v1 && kv) ? __ldg(a1p + k) : 0.0;
            b0[ks] = (c0v && kv) ? __ldcg(bp + (int64_t)k * ldb) : 0.0;
            b1[ks] = (c1v && kv) ? __ldcg(bp + (int64_t)k * ldb + 8) : 0.0;
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const double x0 = neg ? -a0[ks] : a0[ks], x1 = neg ? -a1[ks] : a1[ks];
            dmma884(acc[0][0][0], acc[0][0][1], x0, b0[ks]);
            dmma884(acc[0][1][0], acc[0][1][1], x0, b1[ks]);
            dmma884(acc[1][0][0], acc[1][0][1], x1, b0[ks]);
            dmma884(acc[1][1][0], acc[1][1][1], x1, b1[ks]);
        }
    }
}

// c fragments <-> 16 rows x 16 columns of a global row-major block (ld = s); rows / columns past the valid range skipped
__device__ __forceinline__ void frag_load_g(double (&acc)[2][2][2], const double *P, int64_t ld, int rows_valid, int cols_valid, int lane) {
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = 8 * mi + g, cidx = 8 * ni + 2 * tg + e;
                acc[mi][ni][e] = (r < rows_valid && cidx < cols_valid) ? __ldcg(P + (int64_t)r * ld + cidx) : 0.0;
            }
}
__device__ __forceinline__ void frag_store_g(const double (&acc)[2][2][2], double *P, int64_t ld, int rows_valid, int cols_valid, int lane) {
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = 8 * mi + g, cidx = 8 * ni + 2 * tg + e;
                if (r < rows_valid && cidx < cols_valid) __stcg(P + (int64_t)r * ld + cidx, acc[mi][ni][e]);
            }
}

__global__ void __launch_bounds__(256, 2)
k_ll_sweep(int first, int nfl, int s, int H, int T, int backward, const FrontDev *__restrict__ fronts, const double *__restrict__ L,
           const double *__restrict__ U, const double *__restrict__ inv, double *W, double *__restrict__ Y, uint32_t *flags,
           uint32_t epoch, int pivots_only) {
    // pivots_only: the boundary rows are handled by full-size GEMM tiles in separate launches -- forward: after this
    // kernel (k_solve_forward_boundary), so only the pivot tiles are swept here; backward: before it (k_solve_boundary),
    // so the U12 term is already in the workspace
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = warp >> 1, chh = warp & 1;             // 16-row group of the tile, 16-column half of the chunk
    const int per_t = nfl * H;
    for (int u = blockIdx.x; u < T * per_t; u += gridDim.x) {
        const int tt = u / per_t, rem = u - tt * per_t, f = rem / H, h = rem - f * H;
        const FrontDev F = fronts[first + f];
        const int n = F.ns + F.nb, ntp = (F.ns + NB - 1) / NB, ntb = (F.nb + NB - 1) / NB;
        int t;
        if (!backward) { t = tt; if (t >= ntp + ntb || (pivots_only && t >= ntp)) continue; }
        else { t = ntp - 1 - tt; if (t < 0) continue; }
        const bool piv = t < ntp;
        const int r0 = piv ? NB * t : F.ns + NB * (t - ntp);
        const int rows = min(NB, (piv ? F.ns : n) - r0);
        const int c0 = 32 * h + 16 * chh, cv = s - c0;     // this warp's columns
        double *w = W + F.w_off;
        uint32_t *fl = flags + (F.inv_off / (2 * NB * NB)) * LL_MAX_CHUNKS + h;   // one flag per (pivot block, column chunk)
        double acc[2][2][2];
        const int wr = r0 + 16 * rg, rv = rows - 16 * rg;  // this warp's rows
        frag_load_g(acc, w + (int64_t)wr * s + c0, s, rv, cv, lane);
        const double *Lrow = L + F.l_off + (int64_t)wr * F.ns;
        if (!backward) {
            const int kend = piv ? t : ntp;
            for (int k = 0; k < kend; ++k) {
                flag_wait(fl + LL_MAX_CHUNKS * k, epoch, lane);
                slab_mma_g(acc, Lrow + NB * k, F.ns, rv, min(NB, F.ns - NB * k), w + (int64_t)(NB * k) * s + c0, s, cv, lane, true);
            }
        } else {
            if (F.nb && !pivots_only) slab_mma_g(acc, U + F.u_off + (int64_t)wr * F.nb, F.nb, rv, F.nb, w + (int64_t)F.ns * s + c0, s, cv, lane, true);
            for (int k = ntp - 1; k > t; --k) {
                flag_wait(fl + LL_MAX_CHUNKS * k, epoch, lane);
                slab_mma_g(acc, Lrow + NB * k, F.ns, rv, min(NB, F.ns - NB * k), w + (int64_t)(NB * k) * s + c0, s, cv, lane, true);
            }
        }
        frag_store_g(acc, w + (int64_t)wr * s + c0, s, rv, cv, lane);
        if (piv) {
            __threadfence();
            __syncthreads();
            if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(fl + LL_MAX_CHUNKS * t), "r"(epoch) : "memory");
            // solved rows: y_t = Linv_tt w_t (forward) / x_t = Uinv_tt w_t (backward); only this CTA's columns are needed
            const double *Dinv = inv + F.inv_off + (int64_t)t * (2 * NB * NB) + (backward ? NB * NB : 0);
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            slab_mma_g(acc, Dinv + (int64_t)(16 * rg) * NB, NB, NB - 16 * rg, rows, w + (int64_t)r0 * s + c0, s, cv, lane, false);
            frag_store_g(acc, Y + F.w_off + (int64_t)wr * s + c0, s, rv, cv, lane);
        }
    }
}

// Largest column chunk in {32, 16} whose block leaves room for two CTAs per SM (measured: with one CTA per SM, or
// 8-column chunks, the per-step GEMM launches of the general path are faster); 8 only when there are <= 8
// right-hand sides.  0 = the level uses the general path.
// Three CTAs per SM (80 registers per thread, 16 padding rows: the 32-column kernels read B rows at most 15 past the
// block) for the levels of the smallest fronts: their sweeps are bound by the latency of the panel loads (ncu:
// long_scoreboard first, tensor pipe 23 % active at 16 warps per SM).  Measured on config 3 (B200, 64 right-hand sides,
// forward / backward): fronts up to 162 rows 0.904 / 1.101 -> 0.874 / 0.986 ms, up to 174 rows 0.346 / 0.390 -> 0.287 /
// 0.338 ms, but up to 240 rows 0.840 / 0.866 -> 0.879 / 0.909 ms (the spills cost more than the third CTA hides), hence
// the 64 KB limit (206 rows).
constexpr int SW_THREE_CTA_SMEM = 64 * 1024, SW_PAD_ROWS_3 = 16;
static int sweep_chunk(int maxn, int s, size_t *smem, bool *three = nullptr) {
    static const bool no3 = getenv("PN_SWEEP_TWO_CTAS") != nullptr;
    const int cand[3] = {32, 16, 8};
    const int scap = std::max(8, ((s + 7) / 8) * 8);
    if (three) {
        *three = false;
        const size_t need3 = (size_t)(maxn + SW_PAD_ROWS_3) * (32 + 4) * sizeof(double);
        if (!no3 && scap >= 32 && need3 <= (size_t)SW_THREE_CTA_SMEM) { *smem = need3; *three = true; return 32; }
    }
    for (int cw : cand) {
        if (cw > scap || (cw == 8 && s > 8)) continue;
        size_t need = (size_t)(maxn + SW_PAD_ROWS) * (cw + 4) * sizeof(double);
        if (need <= (size_t)SW_TWO_CTA_SMEM) { *smem = need; return cw; }
    }
    return 0;
}

// backward boundary update  W[0:ns] -= U12[ns x nb] W[ns:n]
__global__ void __launch_bounds__(256, 2)
k_solve_boundary(int first, int s, const FrontDev *__restrict__ fronts, const double *__restrict__ U, double *__restrict__ W) {
    FrontDev F = fronts[first + blockIdx.x];
    if (F.nb == 0 || (int)blockIdx.y * BM >= F.ns || (int)blockIdx.z * BN >= s) return;
    double *w = W + F.w_off;
    gemm_tile(U + F.u_off, F.nb, w + (int64_t)F.ns * s, s, w, s, F.ns, s, F.nb, blockIdx.y, blockIdx.z, 0);
}

// forward boundary update  W[ns:n] -= Lp[ns:n, 0:ns] W[0:ns]  (all pivot blocks of the front are final; the stored panels
// are pre-multiplied, so the right-hand side is the accumulated, unsolved W[0:ns])
__global__ void __launch_bounds__(256, 2)
k_solve_forward_boundary(int first, int s, const FrontDev *__restrict__ fronts, const double *__restrict__ L, double *__restrict__ W) {
    FrontDev F = fronts[first + blockIdx.x];
    if (F.nb == 0 || (int)blockIdx.y * BM >= F.nb || (int)blockIdx.z * BN >= s) return;
    double *w = W + F.w_off;
    gemm_tile(L + F.l_off + (int64_t)F.ns * F.ns, F.ns, w, s, w + (int64_t)F.ns * s, s, F.nb, s, F.ns, blockIdx.y, blockIdx.z, 0);
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
// Host half of the analysis (vertex graph, nested dissection, fronts): needs only the model's connectivity and the
// DOF partition, so direct_prefetch() starts it on a second host thread while the symbolic phase of the assembly
// (element classes, node recipes: also host work) is still running.
static void analyse_host(Ctx *c, const std::vector<int32_t> &newidx, bool timing, int64_t n1) {
    DirectSolver &d = *c->direct;
    const Model &m = c->model;
    HostTick tick("direct analyse");
    tick.sync = timing;                                   // worker thread: wall clock only
    // vertex graph from element connectivity (a superset of the K11 pattern)
    std::vector<ConnView> conn;
    conn.push_back({m.n_springs, 2, m.h_spring_ij.data(), m.h_spring_active.data()});
    conn.push_back({m.n_members, 2, m.h_mem_ij.data(), m.h_mem_active.data()});
    conn.push_back({m.n_quads, 4, m.h_quad_ijmn.data(), nullptr});
    conn.push_back({m.n_plates, 4, m.h_plate_ijmn.data(), nullptr});
    // the graph vectors and the builder's scratch live in the solver object: a re-analysis touches no fresh pages
    std::vector<int32_t> &vstart = d.g_vstart, &adj = d.g_adj;
    std::vector<double> &coords = d.g_coords;
    std::vector<int64_t> &adj_ptr = d.g_adj_ptr;
    // DOF-type classes predicted from the element list (flat shell models in an axis plane: membrane / bending /
    // drilling never meet in a stored entry); k_entry_maps verifies the prediction against the assembled pattern
    static const bool no_classes = getenv("PN_NO_DOF_CLASSES") != nullptr;
    DofClasses cls;
    if (!no_classes && !d.classes_rejected && c->opt_dof_classes == 1) {
        std::vector<ConnView> shells(conn.begin() + 2, conn.end());
        cls = predict_dof_classes(m.h_xyz.data(), shells, m.n_springs + m.n_members);
    } else if (!d.classes_rejected && c->opt_dof_classes == 2) {
        cls.n = 6;                                        // self-test of the verification: no model decouples like this
        for (int k = 0; k < 6; ++k) cls.of_type[k] = (uint8_t)k;
    }
    d.n_classes = cls.n;
    build_vertex_graph(m.n_nodes, newidx.data(), m.h_xyz.data(), conn, vstart, coords, adj_ptr, adj, &d.g_scratch);
    tick.lap("vertex graph");
    int32_t nv = (int32_t)vstart.size() - 1;
    int leaf = 24;
    if (const char *e = getenv("PN_LEAF_SIZE")) leaf = std::max(1, atoi(e));
    if (cls.n > 1) {
        d.g_row_class.resize((size_t)n1);
        const int64_t n_dof = 6 * m.n_nodes;
#pragma omp parallel for schedule(static)
        for (int64_t dd = 0; dd < n_dof; ++dd)
            if (newidx[dd] >= 0) d.g_row_class[newidx[dd]] = cls.of_type[dd % 6];
    }
    direct_symbolic(n1, nv, vstart.data(), coords.data(), adj_ptr.data(), adj.data(), leaf, d.sym,
                    cls.n > 1 ? d.g_row_class.data() : nullptr, cls.n, cls.diagonal);
    {
        // front of every K11 row in level-major numbering (page-locked: uploaded as is).  Done here, on the thread
        // that already owns an OpenMP team: a parallel region on the caller's thread would leave a second team
        // spinning on the cores.
        const DirectSym &S = d.sym;
        const int nf = (int)S.fronts.size();
        std::vector<int32_t> lm_of(nf);
        for (int q = 0; q < nf; ++q) lm_of[S.level_fronts[q]] = q;
        d.h_dof_front_lm.resize(S.dof_front.size());
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)S.dof_front.size(); ++i) d.h_dof_front_lm[i] = lm_of[S.dof_front[i]];
    }
    tick.lap("nested dissection + fronts");
}

static void join_host_analysis(DirectSolver &d) {
    if (d.host_future.valid()) d.host_future.get();      // rethrows what the worker threw
}

// Starts the host half of the analysis on a worker thread.  The DOF partition is recomputed on the host from the
// model's support / enforced-displacement masks (same rule as k_known_flags + k_partition_lists), so the worker needs
// nothing from the device and can start as soon as the model arrays are set (pn_prefetch_analysis), well before the
// symbolic phase.  A model change joins and discards it (direct_invalidate).
static bool shared_ordering(const Ctx *c);
void direct_prefetch(Ctx *c) {
    static const bool off = getenv("PN_NO_ANALYSIS_PREFETCH") != nullptr;
    const Model &m = c->model;
    if (off || !m.n_nodes) return;
    if (shared_ordering(c) && c->dist_rank != 0) return;          // rank 0 orders for all ranks (share_ordering)
    if (!c->direct) c->direct = new DirectSolver();
    DirectSolver &d = *c->direct;
    if (d.host_state != 0) return;                       // running, or done and still valid for this model
    try { join_host_analysis(d); } catch (...) {}
    d.analysed = false;
    d.host_state = 1;
    const int wt = host_worker_threads(c);
    d.host_future = host_worker_submit([c, wt]() {
        cudaSetDevice(c->device);          // the page-locked index lists are allocated from this thread
        omp_set_num_threads(wt);
        const Model &mm = c->model;
        const int64_t n_dof = 6 * mm.n_nodes;
        std::vector<int32_t> newidx((size_t)n_dof);
        int32_t fp = 0, kp = 0;
        for (int64_t dd = 0; dd < n_dof; ++dd) {
            const int64_t nd = dd / 6;
            const bool known = ((mm.h_support[nd] | mm.h_enf_mask[nd]) >> (dd % 6)) & 1u;
            newidx[dd] = known ? -1 - kp++ : fp++;
        }
        c->direct->prefetch_n1 = fp;
        analyse_host(c, newidx, false, fp);
    });
}

static void build_dist_plan(Ctx *c);

// ---- one ordering for all ranks of a subtree-to-GPU solve -----------------------------------------------------------
// The ranks of one box share its cores (16 cores / 8 ranks leaves one OpenMP thread per rank) and would all compute the
// same elimination tree; rank 0 computes it with the threads of the whole box and the others receive it: a header and
// one blob (index lists, front table, level lists) through the caller's all-gather, i.e. over NVLink, segment 0 being
// rank 0's.  ~90 MB for config 3; the other ranks' segments carry nothing.
static bool shared_ordering(const Ctx *c) {
    static const bool off = getenv("PN_DIST_PRIVATE_ORDERING") != nullptr;
    return !off && c->dist_world > 1 && c->dist_allgather != nullptr;
}
static void dist_allgather(Ctx *c, int64_t seg);
static void share_ordering(Ctx *c) {
    DirectSolver &d = *c->direct;
    DirectSym &S = d.sym;
    DistPlan &P = d.dist;
    cudaStream_t st = c->stream;
    const int world = c->dist_world;
    const bool root = c->dist_rank == 0;
    constexpr int NH = 16;
    double hdr[NH] = {0};
    if (root) {
        hdr[0] = (double)S.n; hdr[1] = (double)S.fronts.size(); hdr[2] = (double)S.fidx.size(); hdr[3] = (double)S.rel.size();
        hdr[4] = (double)S.level_ptr.size(); hdr[5] = (double)S.max_children; hdr[6] = (double)S.factor_entries;
        hdr[7] = S.factor_flops; hdr[8] = (double)S.max_front; hdr[9] = (double)d.n_classes; hdr[10] = 1.0;
    }
    P.send.resize(std::max<int64_t>(P.send.n, NH));
    P.recv.resize(std::max<int64_t>(P.recv.n, (int64_t)NH * world));
    PN_CUDA(cudaMemcpyAsync(P.send.ptr, hdr, sizeof(hdr), cudaMemcpyHostToDevice, st));
    dist_allgather(c, NH);
    PN_CUDA(cudaMemcpyAsync(hdr, P.recv.ptr, sizeof(hdr), cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    if (hdr[10] != 1.0) throw StatusError(PN_ERR_CUDA, "pn_set_distributed: rank 0 did not deliver the ordering");
    const int64_t n = (int64_t)hdr[0], nf = (int64_t)hdr[1], nidx = (int64_t)hdr[2], nrel = (int64_t)hdr[3], nlp = (int64_t)hdr[4];
    if (!root) {
        S.n = n; S.max_children = (int32_t)hdr[5]; S.factor_entries = (int64_t)hdr[6]; S.factor_flops = hdr[7];
        S.max_front = (int64_t)hdr[8]; d.n_classes = (int)hdr[9];
        S.perm_pos.resize(n); S.dof_front.resize(n); S.s_pos_dof.resize(n); d.h_dof_front_lm.resize(n);
        S.fronts.resize(nf); S.fidx.resize(nidx); S.fpos.resize(nidx); S.rel.resize(nrel);
        S.level_ptr.resize(nlp); S.level_fronts.resize(nf);
    }
    struct Part { void *host; int64_t bytes; };
    const Part parts[] = {
        {S.perm_pos.data(), n * 4}, {S.dof_front.data(), n * 4}, {S.s_pos_dof.data(), n * 4}, {d.h_dof_front_lm.data(), n * 4},
        {S.fidx.data(), nidx * 4}, {S.fpos.data(), nidx * 4}, {S.rel.data(), nrel * 4},
        {S.level_ptr.data(), nlp * 4}, {S.level_fronts.data(), nf * 4}, {S.fronts.data(), nf * (int64_t)sizeof(FrontSym)}};
    int64_t total = 0;
    for (const Part &pt : parts) total += (pt.bytes + 15) / 16 * 16;
    const int64_t seg = total / 8;
    P.send.resize(std::max<int64_t>(P.send.n, seg));
    P.recv.resize(std::max<int64_t>(P.recv.n, seg * world));
    if (root) {
        int64_t off = 0;
        for (const Part &pt : parts) {
            if (pt.bytes) PN_CUDA(cudaMemcpyAsync((char *)P.send.ptr + off, pt.host, (size_t)pt.bytes, cudaMemcpyHostToDevice, st));
            off += (pt.bytes + 15) / 16 * 16;
        }
    }
    dist_allgather(c, seg);
    if (!root) {
        int64_t off = 0;
        for (const Part &pt : parts) {
            if (pt.bytes) PN_CUDA(cudaMemcpyAsync(pt.host, (const char *)P.recv.ptr + off, (size_t)pt.bytes, cudaMemcpyDeviceToHost, st));
            off += (pt.bytes + 15) / 16 * 16;
        }
        PN_CUDA(cudaStreamSynchronize(st));
    }
}

static void analyse(Ctx *c) {
    if (!c->direct) c->direct = new DirectSolver();
    DirectSolver &d = *c->direct;
    cudaStream_t st = c->stream;
    const Model &m = c->model;
    (void)m;
    HostTick tick("direct analyse");
    const bool shared = shared_ordering(c);
    if (shared && c->dist_rank != 0) {
        // rank 0 computes the ordering for everybody (share_ordering)
        if (d.host_state == 1) { try { join_host_analysis(d); } catch (...) {} }
        share_ordering(c);
        d.host_state = 2;
        tick.lap("ordering received from rank 0");
    }
    if (d.host_state == 1) {
        d.host_state = 0;
        join_host_analysis(d);
        if (d.prefetch_n1 == c->n1) d.host_state = 2;    // (always, unless the partition rule and its host copy diverge)
        tick.lap("join prefetched host analysis");
    }
    const bool computed_here = !(shared && c->dist_rank != 0);
    if (d.host_state != 2) {
        std::vector<int32_t> newidx(c->n_dof);
        PN_CUDA(cudaMemcpyAsync(newidx.data(), c->newidx.ptr, sizeof(int32_t) * c->n_dof, cudaMemcpyDeviceToHost, st));
        PN_CUDA(cudaStreamSynchronize(st));
        analyse_host(c, newidx, true, c->n1);
        d.host_state = 2;
    }
    if (shared && computed_here) {
        share_ordering(c);
        tick.lap("ordering sent to the other ranks");
    }
    const DirectSym &S = d.sym;
    // the index lists are page-locked: their DMA runs while the host builds the front table and the children lists
    d.fidx.upload(S.fidx.data(), (int64_t)S.fidx.size(), st);
    d.fpos.upload(S.fpos.data(), (int64_t)S.fpos.size(), st);
    d.rel.upload(S.rel.data(), (int64_t)S.rel.size(), st);
    d.perm_pos.upload(S.perm_pos.data(), (int64_t)S.perm_pos.size(), st);
    d.dof_front_lm.upload(d.h_dof_front_lm.data(), (int64_t)d.h_dof_front_lm.size(), st);     // filled by analyse_host

    // level-major renumbering of fronts
    int nf = (int)S.fronts.size();
    d.n_levels = (int)S.level_ptr.size() - 1;
    d.level_ptr = S.level_ptr;
    std::vector<int32_t> lm_of(nf);
    for (int q = 0; q < nf; ++q) lm_of[S.level_fronts[q]] = q;
    d.h_fronts.assign(nf, FrontDev());
    d.level_max_ns.assign(d.n_levels, 0); d.level_max_n.assign(d.n_levels, 0); d.level_max_nb.assign(d.n_levels, 0);
    d.level_max_children.assign(d.n_levels, 0);
    d.level_f_total.assign(d.n_levels, 0); d.level_n_total.assign(d.n_levels, 0);
    std::vector<uint8_t> front_level(nf);
    int64_t l_total = 0, u_total = 0, inv_total = 0;
    for (int l = 0; l < d.n_levels; ++l) {
        int64_t foff = 0, noff = 0;
        for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
            const FrontSym &fs = S.fronts[S.level_fronts[q]];
            FrontDev &F = d.h_fronts[q];
            F.ns = fs.ns; F.nb = fs.nb;
            F.parent = fs.parent >= 0 ? lm_of[fs.parent] : -1;
            F.child_rank = fs.child_rank;
            F.idx_off = fs.idx_off; F.rel_off = fs.rel_off;
            int64_t n = fs.ns + fs.nb;
            F.f_off = foff; foff += n * front_ld((int)n);
            F.w_off = noff; noff += n;             // scaled by s at solve time
            F.l_off = F.u_off = F.inv_off = 0;     // assigned below, for the fronts this rank factorises
            front_level[q] = (uint8_t)l;
            d.level_max_ns[l] = std::max(d.level_max_ns[l], fs.ns);
            d.level_max_nb[l] = std::max(d.level_max_nb[l], fs.nb);
            d.level_max_n[l] = std::max<int32_t>(d.level_max_n[l], (int32_t)n);
            d.level_max_children[l] = std::max(d.level_max_children[l], fs.child_rank + 1);
            if (n * n >= (int64_t)0xffffffffu) throw ArgError("direct solver: front too large");
        }
        d.level_f_total[l] = foff; d.level_n_total[l] = noff;
    }
    if (d.n_levels > 255) throw ArgError("direct solver: elimination tree too deep");
    tick.lap("maps: front table");
    // Subtree-to-GPU plan first: the permanent factor (panels, U12, diagonal inverses) is stored only for the fronts
    // this rank factorises -- its subtrees and the replicated top -- so N GPUs hold an N times larger bottom of the tree.
    build_dist_plan(c);
    for (int l = 0; l < d.n_levels; ++l) {
        const int q0 = d.dist.active ? d.dist.lo[l] : d.level_ptr[l], q1 = d.dist.active ? d.dist.hi[l] : d.level_ptr[l + 1];
        for (int q = q0; q < q1; ++q) {
            FrontDev &F = d.h_fronts[q];
            const int64_t n = F.ns + F.nb;
            F.l_off = l_total; l_total += n * F.ns;
            F.u_off = u_total; u_total += (int64_t)F.ns * F.nb;
            F.inv_off = inv_total; inv_total += (int64_t)((F.ns + NB - 1) / NB) * (2 * NB * NB);
        }
    }
    {
        // children of every front in child-rank order (level-major ids): counting sort, ranks are 0 .. #children - 1
        std::vector<int32_t> cptr(nf + 1, 0), cidx;
        for (int q = 0; q < nf; ++q)
            if (d.h_fronts[q].parent >= 0) ++cptr[d.h_fronts[q].parent + 1];
        for (int q = 0; q < nf; ++q) cptr[q + 1] += cptr[q];
        cidx.assign((size_t)cptr[nf], -1);
        for (int q = 0; q < nf; ++q) {
            const FrontDev &F = d.h_fronts[q];
            if (F.parent < 0) continue;
            const int32_t slot = cptr[F.parent] + F.child_rank;
            if (slot >= cptr[F.parent + 1] || cidx[slot] >= 0) throw ArgError("direct solver: inconsistent child ranks");
            cidx[slot] = q;
        }
        d.child_ptr.upload(cptr.data(), nf + 1, st);
        d.child_idx.upload(cidx.data(), (int64_t)cidx.size(), st);
        PN_CUDA(cudaStreamSynchronize(st));
    }
    tick.lap("maps: dist plan + children");
    d.fronts.upload(d.h_fronts.data(), nf, st);
    d.L.resize(l_total); d.U.resize(u_total); d.inv.resize(inv_total);
    d.ll_flags.resize(inv_total / (2 * NB * NB) * LL_MAX_CHUNKS + 1);
    PN_CUDA(cudaMemsetAsync(d.ll_flags.ptr, 0, sizeof(uint32_t) * (inv_total / (2 * NB * NB) * LL_MAX_CHUNKS + 1), st));
    d.ll_epoch = 0;
    if (!d.ll_grid) {
        int per_sm = 0, dev = 0, sms = 0;
        PN_CUDA(cudaGetDevice(&dev));
        PN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        PN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ll_sweep, 256, 0));
        d.ll_grid = std::max(1, per_sm) * sms;
    }
    // (L needs no zero fill: the sweeps read, per block column, only the rows below / above the diagonal block, and
    // those are all written by k_lu_panels / k_store_panels; columns past a partial block are masked by the k count)
    d.fronts_s = 0;
    if (!d.attrs_set) {
        PN_CUDA(cudaFuncSetAttribute(k_lu_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_lu_strip, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_front_factor_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_FACTOR_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_lu_panels, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_lu_update, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_store_panels, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_solve_step, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_front_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_front_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_solve_boundary, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_solve_forward_boundary, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_front_forward_sm<32, 3, SW_PAD_ROWS_3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_MAX_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_front_backward_sm<32, 3, SW_PAD_ROWS_3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_MAX_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_front_forward_sm<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_MAX_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_front_forward_sm<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_MAX_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_front_forward_sm<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_MAX_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_front_backward_sm<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_MAX_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_front_backward_sm<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_MAX_SMEM));
        PN_CUDA(cudaFuncSetAttribute(k_front_backward_sm<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_MAX_SMEM));
        d.attrs_set = true;
    }
    if (!d.n_fstreams) {
        int ns_ = 4;
        if (const char *e = getenv("PN_FACTOR_STREAMS")) ns_ = std::min(4, std::max(1, atoi(e)));
        if (const char *e = getenv("PN_FACTOR_SPLIT_MAX")) d.split_max_fronts = std::max(0, atoi(e));
        // the chains of block steps are latency-bound: their CTAs go first whenever SM slots free up; the remainder updates
        // of the look-ahead schedule are throughput work and take what is left
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);          // hi = numerically smallest = most urgent
        static const bool flat_prio = getenv("PN_FLAT_STREAM_PRIORITY") != nullptr;
        if (flat_prio) prio_lo = prio_hi = 0;
        for (int q = 0; q < ns_; ++q) {
            PN_CUDA(cudaStreamCreateWithPriority(&d.fstream[q], cudaStreamNonBlocking, prio_hi));
            PN_CUDA(cudaEventCreateWithFlags(&d.ev_join[q], cudaEventDisableTiming));
        }
        for (int q = 0; q < 4; ++q) {
            PN_CUDA(cudaStreamCreateWithPriority(&d.ustream[q], cudaStreamNonBlocking, prio_lo));
            PN_CUDA(cudaEventCreateWithFlags(&d.ev_prio[q], cudaEventDisableTiming));
            PN_CUDA(cudaEventCreateWithFlags(&d.ev_rem[q], cudaEventDisableTiming));
        }
        PN_CUDA(cudaEventCreateWithFlags(&d.ev_fork, cudaEventDisableTiming));
        d.n_fstreams = ns_;
    }
    d.status.resize(1);
    d.status.zero(st);
    tick.lap("maps: uploads + factor storage");
    // per-entry destination maps
    const CsrBlock &A = c->blk[0];
    DevBuf<int32_t> &ent_row = c->sym_ent_row;
    ent_row.resize(A.nnz);
    d.ent_front.resize(A.nnz); d.ent_dst.resize(A.nnz); d.ent_src.resize(A.nnz);
    d.front_eptr.assign((size_t)nf + 1, 0);
    if (A.nnz) {
        // destination of every entry in K11 order, then a stable sort by owning front (fronts are level-major)
        DevBuf<int32_t> &front_raw = d.ent_front_raw;
        DevBuf<uint32_t> &dst_raw = d.ent_dst_raw, &iota = d.ent_iota;
        DevBuf<int64_t> &eptr = d.d_front_eptr;
        front_raw.resize(A.nnz); dst_raw.resize(A.nnz); iota.resize(A.nnz); eptr.resize(nf + 1);
        k_entry_rows<<<grid_for(A.rows, 128), 128, 0, st>>>(A.rows, A.indptr.ptr, ent_row.ptr);
        k_entry_maps<<<grid_for(A.nnz, 256), 256, 0, st>>>(A.nnz, ent_row.ptr, A.indices.ptr, d.perm_pos.ptr,
                                                            d.dof_front_lm.ptr, d.fronts.ptr, d.fpos.ptr,
                                                            front_raw.ptr, dst_raw.ptr, d.status.ptr);
        k_iota_u32<<<grid_for(A.nnz, 256), 256, 0, st>>>(A.nnz, iota.ptr);
        int end_bit = 1;
        while ((1 << end_bit) <= nf && end_bit < 31) ++end_bit;
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, front_raw.ptr, d.ent_front.ptr, iota.ptr, d.ent_src.ptr, (int)A.nnz, 0, end_bit, st);
        c->cub_tmp.resize((int64_t)bytes);
        PN_CUDA(cub::DeviceRadixSort::SortPairs(c->cub_tmp.ptr, bytes, front_raw.ptr, d.ent_front.ptr, iota.ptr, d.ent_src.ptr, (int)A.nnz, 0, end_bit, st));
        k_gather_by_u32<<<grid_for(A.nnz, 256), 256, 0, st>>>(A.nnz, d.ent_src.ptr, dst_raw.ptr, d.ent_dst.ptr);
        k_front_entry_ptr<<<grid_for(nf + 1, 128), 128, 0, st>>>(nf, A.nnz, d.ent_front.ptr, eptr.ptr);
        PN_CUDA(cudaMemcpyAsync(d.front_eptr.data(), eptr.ptr, sizeof(int64_t) * (nf + 1), cudaMemcpyDeviceToHost, st));
        c->count_launch(9);
        PN_CUDA(cudaStreamSynchronize(st));
    }
    int status = 0;
    PN_CUDA(cudaMemcpyAsync(&status, d.status.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    tick.lap("maps: entry destinations");
    if (status == 2 && d.n_classes > 1) {
        // an assembled entry joins two of the predicted DOF-type classes: order again with one class (node-based)
        d.classes_rejected = true;
        d.host_state = 0;
        c->ctr[PC_DOF_CLASS_REJECTS]++;
        analyse(c);
        return;
    }
    if (status == 2) throw ArgError("direct solver: a K11 entry falls outside its front (symbolic analysis bug)");
    c->ctr[PC_DOF_CLASSES] = d.n_classes;
    int64_t amax = 0;
    for (int l = 0; l < d.n_levels; ++l) amax = std::max(amax, d.level_f_total[l]);
    d.arena[0].resize(amax); d.arena[1].resize(amax);
    d.analysed = true;
    tick.lap("device maps + allocation");
    c->stats.factor_nnz = S.factor_entries;
    c->stats.factor_flops = S.factor_flops;
    // flops the trailing-update kernel (k_lu_update) really performs on THIS rank (tile-granular, clipped to the matrix): general
    // mode = every tile of the trailing square per pivot group, symmetric mode = the tiles that touch or lie below the
    // diagonal; with a subtree-to-GPU plan only the fronts this rank factorises
    {
        static const bool force_lu = getenv("PN_FORCE_UNSYMMETRIC") != nullptr;
        const bool sym = c->model.symmetric && !force_lu;
        double upd = 0.0;
        for (int l = 0; l < d.n_levels; ++l) {
            const int q0 = d.dist.active ? d.dist.lo[l] : d.level_ptr[l], q1 = d.dist.active ? d.dist.hi[l] : d.level_ptr[l + 1];
            if (q1 - q0 >= fused_factor_min()) continue;       // factorised by k_front_factor_fused, not by k_lu_update
            for (int q = q0; q < q1; ++q) {
                const FrontDev &fs = d.h_fronts[q];
                int64_t n = fs.ns + fs.nb;
                for (int g0 = 0; g0 < fs.ns; g0 += GB) {
                    int ge = std::min(g0 + GB, fs.ns);
                    int64_t rest = n - ge;
                    double K = ge - g0;
                    // executed at the granularity of a warp's 32 x 32 block: blocks inside the matrix and, in symmetric
                    // mode, not strictly above the diagonal (gemm_tile's warp_on)
                    for (int64_t r = 0; r < rest; r += 32)
                        for (int64_t cc = 0; cc < rest; cc += 32)
                            if (!(sym && r + 31 < cc)) upd += 2.0 * K * 32.0 * 32.0;
                }
            }
        }
        c->stats.update_flops = upd;
    }
    c->stats.solve_flops = 2.0 * (double)S.factor_entries;
}

// PN_DEBUG_LEVELS: device time of every launch class inside one level (events around each launch)
struct LevelProf {
    bool on;
    cudaStream_t st;
    std::vector<cudaEvent_t> ev;
    std::vector<int> cls;
    size_t used = 0;
    static constexpr int NC = 8;
    LevelProf(bool enable, cudaStream_t s) : on(enable), st(s) {}
    void begin(int c) {
        if (!on) return;
        if (used + 2 > ev.size()) { ev.resize(used + 2); cudaEventCreate(&ev[used]); cudaEventCreate(&ev[used + 1]); }
        cls.push_back(c);
        cudaEventRecord(ev[used], st);
    }
    void end() {
        if (!on) return;
        cudaEventRecord(ev[used + 1], st);
        used += 2;
    }
    void report(int level) {
        if (!on) return;
        cudaStreamSynchronize(st);
        double ms[NC] = {0}; int n[NC] = {0};
        for (size_t i = 0; i < cls.size(); ++i) {
            float t = 0;
            cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]);
            ms[cls[i]] += t; n[cls[i]]++;
        }
        static const char *names[NC] = {"zero+assemble", "extend_add", "diag", "panels", "transpose", "strip", "update", "store"};
        fprintf(stderr, "[pn]   level %2d kernels:", level);
        for (int c = 0; c < NC; ++c) if (n[c]) fprintf(stderr, " %s %.3f ms/%d", names[c], ms[c], n[c]);
        fprintf(stderr, "\n");
        cls.clear(); used = 0;
    }
    ~LevelProf() { for (auto e : ev) cudaEventDestroy(e); }
};

// Subtree-to-GPU plan for the current tree (see DistPlan).  Level-major ids; parent ids are level-major as well.
static void build_dist_plan(Ctx *c) {
    DirectSolver &d = *c->direct;
    DistPlan &P = d.dist;
    const int world = c->dist_world, rank = c->dist_rank;
    const int nl = d.n_levels;
    P.active = false;
    P.cut = -1;
    auto whole_levels = [&]() {
        P.lo.assign(nl, 0); P.hi.assign(nl, 0);
        for (int l = 0; l < nl; ++l) { P.lo[l] = d.level_ptr[l]; P.hi[l] = d.level_ptr[l + 1]; }
    };
    whole_levels();
    c->ctr[PC_DIST_CUT_LEVEL] = -1;
    if (world <= 1 || !c->dist_allgather || nl < 2) return;
    // Cut level.  Everything above it is replicated on every rank, everything on and below it is shared out as whole
    // subtrees in contiguous, cost-balanced blocks of the cut level's fronts.  Candidates run from the highest level
    // with at least `world` fronts down to the first one with 8 x world fronts; the one with the smallest
    // (largest rank share + replicated part) in factorisation flops is taken: a higher cut replicates less but may
    // balance badly (fronts of different DOF-type classes on one level differ by a factor of three in cost).
    const int nf = (int)d.h_fronts.size();
    auto front_cost = [&](int q) {
        const double ns = d.h_fronts[q].ns, nb = d.h_fronts[q].nb;
        return 2.0 / 3.0 * ns * ns * ns + 2.0 * ns * ns * nb + 2.0 * ns * nb * nb + 1e4;
    };
    std::vector<double> level_cost(nl, 0.0);
    for (int l = 0; l < nl; ++l)
        for (int q = d.level_ptr[l]; q < d.level_ptr[l + 1]; ++q) level_cost[l] += front_cost(q);
    std::vector<int32_t> anc(nf, -1);
    std::vector<double> cost;
    auto subtree_costs = [&](int cut_) {            // fills anc (ancestor on the cut level) and cost (per cut front)
        const int cf_ = d.level_ptr[cut_], nc_ = d.level_ptr[cut_ + 1] - cf_;
        std::fill(anc.begin(), anc.end(), -1);
        cost.assign(nc_, 0.0);
        for (int q = cf_; q < cf_ + nc_; ++q) anc[q] = q;
        for (int l = cut_ - 1; l >= 0; --l)
            for (int q = d.level_ptr[l]; q < d.level_ptr[l + 1]; ++q) {
                const int p = d.h_fronts[q].parent;
                anc[q] = p >= 0 ? anc[p] : -1;
            }
        for (int l = 0; l <= cut_; ++l)
            for (int q = d.level_ptr[l]; q < d.level_ptr[l + 1]; ++q)
                if (anc[q] >= 0) cost[anc[q] - cf_] += front_cost(q);
    };
    // contiguous blocks of the cut level, balanced by cost, at least one group per rank (a group = the fronts of one
    // separator of the node dissection, one per DOF-type class: they sit side by side on the level and their
    // descendants interleave below, so they go to one rank); returns the largest share, or -1 with too few groups
    const DirectSym &S = d.sym;
    auto share_out = [&](int cut_, std::vector<int32_t> &lo_, std::vector<int32_t> &hi_) -> double {
        const int cf_ = d.level_ptr[cut_], nc_ = d.level_ptr[cut_ + 1] - cf_;
        std::vector<int32_t> gbeg;                      // first front (index into the level) of every group
        std::vector<double> gcost;
        for (int q = 0; q < nc_; ++q) {
            if (q == 0 || S.fronts[S.level_fronts[cf_ + q]].group != S.fronts[S.level_fronts[cf_ + q - 1]].group) {
                gbeg.push_back(q);
                gcost.push_back(0.0);
            }
            gcost.back() += cost[q];
        }
        const int ng = (int)gbeg.size();
        gbeg.push_back(nc_);
        if (ng < world) return -1.0;
        lo_.assign(world, 0); hi_.assign(world, 0);
        double total = 0.0, worst = 0.0;
        for (double v : gcost) total += v;
        int g = 0;
        double acc = 0.0;
        for (int r = 0; r < world; ++r) {
            lo_[r] = cf_ + gbeg[g];
            const int must_leave = world - 1 - r;          // one group at least for every later rank
            const double before = acc;
            acc += gcost[g]; ++g;
            while (g < ng - must_leave && acc + 0.5 * gcost[g] <= total * (r + 1) / world) { acc += gcost[g]; ++g; }
            if (r == world - 1) { while (g < ng) { acc += gcost[g]; ++g; } }
            hi_[r] = cf_ + gbeg[g];
            worst = std::max(worst, acc - before);
        }
        return worst;
    };
    int cut = -1;
    {
        double best = 0.0, above = level_cost[nl - 1];
        std::vector<int32_t> lo_, hi_;
        for (int l = nl - 2; l >= 0; --l) {
            if (d.level_ptr[l + 1] - d.level_ptr[l] >= world) {
                subtree_costs(l);
                const double worst = share_out(l, lo_, hi_);
                if (worst >= 0.0) {
                    const double obj = worst + above;
                    if (cut < 0 || obj < best) { cut = l; best = obj; }
                    if (d.level_ptr[l + 1] - d.level_ptr[l] >= 8 * world) break;
                }
            }
            above += level_cost[l];
        }
    }
    if (const char *e = getenv("PN_DIST_CUT_BELOW")) cut = std::max(0, cut - atoi(e));     // experiments: deeper cut
    if (cut < 0 || cut >= nl - 1) return;              // too narrow: replicated solve
    const int cf = d.level_ptr[cut], nc = d.level_ptr[cut + 1] - cf;
    subtree_costs(cut);
    if (share_out(cut, P.cut_lo, P.cut_hi) < 0.0) return;
    // per-level ranges of this rank: fronts whose cut ancestor is in [cut_lo[rank], cut_hi[rank]); contiguous per level.
    // A tree this does not hold for (a front without ancestor on the cut level, a rank's fronts interleaved with
    // another rank's) is solved replicated on every rank instead.
    for (int l = 0; l <= cut; ++l) {
        int lo = -1, hi = -1;
        for (int q = d.level_ptr[l]; q < d.level_ptr[l + 1]; ++q) {
            const bool mine = anc[q] >= P.cut_lo[rank] && anc[q] < P.cut_hi[rank];
            if (mine) { if (lo < 0) lo = q; hi = q + 1; }
            else if (anc[q] < 0) { whole_levels(); return; }
        }
        if (lo < 0) { lo = hi = d.level_ptr[l]; }
        for (int q = lo; q < hi; ++q)
            if (!(anc[q] >= P.cut_lo[rank] && anc[q] < P.cut_hi[rank])) { whole_levels(); return; }
        P.lo[l] = lo; P.hi[l] = hi;
    }
    // exchange layouts
    P.um_off.assign(nc, 0); P.uv_off.assign(nc, 0);
    P.um_seg = 0; P.uv_seg = 0;
    P.um_tri = c->model.symmetric && getenv("PN_FORCE_UNSYMMETRIC") == nullptr;       // as factorise() decides `sym`
    std::vector<int32_t> owner(nc, 0);
    for (int r = 0; r < world; ++r) {
        int64_t um = 0, uv = 0;
        for (int q = P.cut_lo[r]; q < P.cut_hi[r]; ++q) {
            const int64_t nb = d.h_fronts[q].nb;
            P.um_off[q - cf] = um; um += P.um_tri ? nb * (nb + 1) / 2 : nb * nb;
            P.uv_off[q - cf] = uv; uv += nb;
            owner[q - cf] = r;
        }
        P.um_seg = std::max(P.um_seg, um);
        P.uv_seg = std::max(P.uv_seg, uv);
    }
    // elimination-position range of every rank: from the first pivot of the first front of its first subtree to the last
    // pivot of its last subtree root (top-level fronts that fall inside are replicated anyway)
    P.pos_lo.assign(world, 0); P.pos_hi.assign(world, 0);
    P.x_seg = 0;
    {
        std::vector<int64_t> first_pos(nc, INT64_MAX), last_pos(nc, 0);
        for (int l = 0; l <= cut; ++l)
            for (int q = d.level_ptr[l]; q < d.level_ptr[l + 1]; ++q) {
                const FrontSym &fs = S.fronts[S.level_fronts[q]];
                const int a = anc[q] - cf;
                first_pos[a] = std::min<int64_t>(first_pos[a], fs.pos_start);
                last_pos[a] = std::max<int64_t>(last_pos[a], (int64_t)fs.pos_start + fs.ns);
            }
        for (int r = 0; r < world; ++r) {
            int64_t lo = INT64_MAX, hi = 0;
            for (int q = P.cut_lo[r]; q < P.cut_hi[r]; ++q) { lo = std::min(lo, first_pos[q - cf]); hi = std::max(hi, last_pos[q - cf]); }
            if (lo > hi) lo = hi = 0;
            P.pos_lo[r] = lo; P.pos_hi[r] = hi;
            P.x_seg = std::max(P.x_seg, hi - lo);
        }
        for (int r = 0; r + 1 < world; ++r)
            if (P.pos_hi[r] > P.pos_lo[r + 1]) { whole_levels(); return; }      // replicated solve
    }
    cudaStream_t st = c->stream;
    {
        // (ascending row numbers through a mark array: sorting 800 000 indices cost 20 ms of every analysis)
        std::vector<uint8_t> mark((size_t)S.n, 0);
        for (int l = 0; l < nl; ++l)
            for (int q = (l <= cut ? P.lo[l] : d.level_ptr[l]); q < (l <= cut ? P.hi[l] : d.level_ptr[l + 1]); ++q) {
                const FrontSym &fs = S.fronts[S.level_fronts[q]];
                for (int64_t pp = fs.pos_start; pp < (int64_t)fs.pos_start + fs.ns; ++pp) mark[(size_t)S.s_pos_dof[(size_t)pp]] = 1;
            }
        std::vector<int32_t> rows;
        rows.reserve((size_t)(P.pos_hi[rank] - P.pos_lo[rank]) + 16384);
        for (int64_t r = 0; r < S.n; ++r) if (mark[(size_t)r]) rows.push_back((int32_t)r);
        P.n_res_rows = (int64_t)rows.size();
        P.res_rows.upload(rows.data(), P.n_res_rows, st);
    }
    P.d_um_off.upload(P.um_off.data(), nc, st);
    P.d_uv_off.upload(P.uv_off.data(), nc, st);
    P.d_cut_owner.upload(owner.data(), nc, st);
    d.pos_dof.upload(S.s_pos_dof.data(), (int64_t)S.s_pos_dof.size(), st);
    PN_CUDA(cudaStreamSynchronize(st));
    P.cut = cut;
    P.active = true;
    c->ctr[PC_DIST_CUT_LEVEL] = cut;
}

// all-gather of `seg` doubles per rank through the caller's collective; the library's stream is idle on both sides
static void dist_allgather(Ctx *c, int64_t seg) {
    DistPlan &P = c->direct->dist;
    PN_CUDA(cudaStreamSynchronize(c->stream));
    const int rc = c->dist_allgather(c->dist_user, P.send.ptr, P.recv.ptr, seg);
    if (rc != 0) throw StatusError(PN_ERR_CUDA, "pn_set_distributed: the all-gather callback failed");
    c->ctr[PC_DIST_EXCHANGES]++;
}

// Sum over the ranks of n host doubles (every rank ends with the same values: the all-gather order is fixed).
// Returns false when this context is not part of a distributed solve.
bool direct_dist_sum(Ctx *c, double *vals, int n) {
    if (!c->direct || !c->direct->dist.active || c->dist_world <= 1) return false;
    DistPlan &P = c->direct->dist;
    const int world = c->dist_world;
    P.send.resize(std::max<int64_t>(P.send.n, n));
    P.recv.resize(std::max<int64_t>(P.recv.n, (int64_t)n * world));
    PN_CUDA(cudaMemcpyAsync(P.send.ptr, vals, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    dist_allgather(c, n);
    std::vector<double> all((size_t)n * world);
    PN_CUDA(cudaMemcpyAsync(all.data(), P.recv.ptr, sizeof(double) * n * world, cudaMemcpyDeviceToHost, c->stream));
    PN_CUDA(cudaStreamSynchronize(c->stream));
    for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int r = 0; r < world; ++r) acc += all[(size_t)r * n + j];
        vals[j] = acc;
    }
    return true;
}

// rows [*r0, *r1) of K11 whose part of a residual norm this rank computes (the whole range when not distributed)
void direct_dist_row_share(Ctx *c, int64_t n1, int64_t *r0, int64_t *r1) {
    *r0 = 0; *r1 = n1;
    if (!c->direct || !c->direct->dist.active || c->dist_world <= 1) return;
    const int64_t per = (n1 + c->dist_world - 1) / c->dist_world;
    *r0 = std::min<int64_t>(n1, per * c->dist_rank);
    *r1 = std::min<int64_t>(n1, *r0 + per);
}

void direct_set_distributed(Ctx *c) {
    if (!c->direct) return;
    c->direct->analysed = false;          // the plan is rebuilt with the device maps at the next solve
    c->direct->shared_factor = false;
    c->direct->fronts_s = 0;
}

// After the local factorisation of the cut level: every rank receives the update matrices of all subtree roots into the
// cut level's arena, where the extend-add of the first replicated level expects them.
static void exchange_update_matrices(Ctx *c, int sym) {
    DirectSolver &d = *c->direct;
    DistPlan &P = d.dist;
    cudaStream_t st = c->stream;
    const int world = c->dist_world, rank = c->dist_rank;
    const int cf = d.level_ptr[P.cut];
    double *arena = d.arena[P.cut & 1].ptr;
    if ((sym != 0) != P.um_tri) throw ArgError("direct solver: the exchange layout was planned for the other symmetry mode");
    P.send.resize(std::max<int64_t>(P.um_seg, 1));
    P.recv.resize(std::max<int64_t>(P.um_seg, 1) * world);
    const int mine = P.cut_hi[rank] - P.cut_lo[rank];
    const unsigned gy = (unsigned)std::min<int64_t>(std::max<int64_t>(d.level_max_nb[P.cut], 1), 4096);
    if (mine > 0) {
        k_dist_um<<<dim3(mine, gy), 256, 0, st>>>(P.cut_lo[rank], d.fronts.ptr, cf, P.d_um_off.ptr, arena, P.send.ptr, 0, sym);
        c->count_launch();
    }
    dist_allgather(c, P.um_seg);
    for (int r = 0; r < world; ++r) {
        const int cnt = P.cut_hi[r] - P.cut_lo[r];
        if (r == rank || cnt <= 0) continue;
        k_dist_um<<<dim3(cnt, gy), 256, 0, st>>>(P.cut_lo[r], d.fronts.ptr, cf, P.d_um_off.ptr, arena, P.recv.ptr + (int64_t)r * P.um_seg, 1, sym);
        c->count_launch();
    }
}

static void factorise(Ctx *c, const double *vals11, const std::function<void(int)> *level_done = nullptr) {
    DirectSolver &d = *c->direct;
    cudaStream_t st = c->stream;
    // K is symmetric up to rounding unless a shell has kx_mod != ky_mod: then only the lower triangle of every front
    // is updated (half the trailing-update flops) and the row panels are scaled transposes of the column panels.
    // The ~1e-16 relative asymmetry of the assembled values this ignores is removed by the refinement against K itself.
    static const bool force_lu = getenv("PN_FORCE_UNSYMMETRIC") != nullptr;
    const int sym = (c->model.symmetric && !force_lu) ? 1 : 0;
    static const bool debug_levels = getenv("PN_DEBUG_LEVELS") != nullptr;
    cudaEvent_t dbg0 = nullptr, dbg1 = nullptr;
    if (debug_levels) { cudaEventCreate(&dbg0); cudaEventCreate(&dbg1); }
    static const bool debug_kernels = getenv("PN_DEBUG_LEVEL_KERNELS") != nullptr;
    // symmetric mode: the row panels are written by the panel GEMM's epilogue; the stand-alone transpose kernel is kept for A/B runs
    static const bool transpose_kernel = getenv("PN_TRANSPOSE_KERNEL") != nullptr;
    LevelProf lp(debug_levels && debug_kernels, st);
    long long *dbg_clk = nullptr;
    if (debug_levels && debug_kernels) cudaMalloc(&dbg_clk, 8 * sizeof(long long));
    d.status.zero(st);
    c->ctr[PC_FACTORISATIONS]++;
    c->ctr[PC_LEVELS] = d.n_levels;
    c->ctr[PC_DOF_CLASSES] = d.n_classes;
    for (int l = 0; l < d.n_levels; ++l) c->ctr[PC_MAX_FRONT_NS] = std::max<int64_t>(c->ctr[PC_MAX_FRONT_NS], d.level_max_ns[l]);
    const DistPlan &DP = d.dist;
    for (int l = 0; l < d.n_levels; ++l) {
        // subtree-to-GPU: below and on the cut level only this rank's subtrees, above it every front (replicated)
        int first = DP.active ? DP.lo[l] : d.level_ptr[l], nfl = (DP.active ? DP.hi[l] : d.level_ptr[l + 1]) - first;
        if (DP.active && l == DP.cut + 1) exchange_update_matrices(c, sym);
        if (!nfl) { if (level_done) (*level_done)(l); continue; }
        double *arena = d.arena[l & 1].ptr;
        if (debug_levels) cudaEventRecord(dbg0, st);
        lp.begin(0);
        if (sym) {
            unsigned gy = (unsigned)std::min<int64_t>(std::max<int64_t>((d.level_max_n[l] + 7) / 8, 1), std::max<int64_t>(1, 32768 / std::max(nfl, 1)));
            k_zero_lower<<<dim3(nfl, gy), 256, 0, st>>>(first, d.fronts.ptr, arena);
            c->count_launch();
        } else {
            PN_CUDA(cudaMemsetAsync(arena, 0, sizeof(double) * d.level_f_total[l], st));
        }
        {
            const int64_t k0 = d.front_eptr[first], k1 = d.front_eptr[first + nfl];
            if (k1 > k0) {
                ProfScope ps_(c, PS_FRONT_ASSEMBLE);
                k_assemble_entries<<<grid_for(k1 - k0, 256), 256, 0, st>>>(k0, k1, d.ent_front.ptr, d.ent_dst.ptr, d.ent_src.ptr, d.fronts.ptr, vals11, arena);
            }
        }
        lp.end();
        c->count_launch();
        if (l > 0) {
            // children: this rank's range of the level below (all of it above the cut, where the update matrices of the
            // other ranks' subtree roots have just been received)
            int cfirst = (DP.active && l - 1 <= DP.cut && l <= DP.cut) ? DP.lo[l - 1] : d.level_ptr[l - 1];
            int ncl = ((DP.active && l - 1 <= DP.cut && l <= DP.cut) ? DP.hi[l - 1] : d.level_ptr[l]) - cfirst;
            int64_t nbmax = d.level_max_nb[l - 1];
            // enough row chunks to fill the GPU when the level has few children, few when it has thousands
            unsigned gy = (unsigned)std::min<int64_t>(std::max<int64_t>(nbmax, 1), std::max<int64_t>(1, 16384 / std::max(ncl, 1)));
            for (int r = 0; r < d.level_max_children[l - 1] && ncl > 0; ++r) {
                lp.begin(1);
                { ProfScope ps_(c, PS_EXTEND_ADD); k_extend_add<<<dim3(ncl, gy), 256, 0, st>>>(cfirst, ncl, r, d.fronts.ptr, d.rel.ptr, d.arena[(l - 1) & 1].ptr,
                                                             arena, sym); }
                lp.end();
                c->count_launch();
            }
        }
        int maxns = d.level_max_ns[l], maxn = d.level_max_n[l];
        // The block-step chain (diag -> panels -> strip per 64 pivots) of a level with few fronts keeps only a handful of
        // CTAs busy, while the rank-256 updates are throughput work: such levels are cut into up to four groups of
        // fronts whose chains run on separate streams, so that one group's chain overlaps the others' updates.
        // Look-ahead (symmetric mode, few fronts, more than one pivot group): only the next group's 256 columns of the
        // trailing update sit on the chain of block steps; the remainder (columns further right) runs on a second stream
        // while the next group's diag -> panels -> strip chain proceeds.  Every entry still receives its updates in
        // group order (the remainder stream is in order, and the next priority update waits for the previous remainder),
        // so the factor is bit-identical to the serial schedule.
        static const bool no_lookahead = getenv("PN_NO_LOOKAHEAD") != nullptr;
        static const int la_max_fronts = getenv("PN_LOOKAHEAD_MAX_FRONTS") ? atoi(getenv("PN_LOOKAHEAD_MAX_FRONTS")) : 16;
        auto run_steps = [&](int f0_, int nf_, cudaStream_t sq, int lane) {
            const bool la = sym && !no_lookahead && !lp.on && !c->prof.on && maxns > GB && nf_ <= la_max_fronts;
            cudaStream_t su = d.ustream[lane];
            bool rem_pending = false;
            if (la) c->ctr[PC_LOOKAHEAD]++;
            for (int g0 = 0; g0 < maxns; g0 += GB) {
                int gend = std::min(g0 + GB, maxns);
                for (int kb = g0; kb < gend; kb += NB) {
                    lp.begin(2);
                    { ProfScope ps_(c, PS_LU_DIAG, sq); k_lu_diag<<<nf_, 256, DIAG_SMEM, sq>>>(f0_, kb, d.fronts.ptr, arena, d.inv.ptr, d.status.ptr, sym, dbg_clk); }
                    if (dbg_clk && kb == 0) {
                        long long h[8];
                        cudaMemcpyAsync(h, dbg_clk, sizeof(h), cudaMemcpyDeviceToHost, sq);
                        cudaStreamSynchronize(sq);
                        fprintf(stderr, "[pn]   level %2d diag clocks: load %lld  lu %lld  store %lld  inv32 %lld  gemm %lld  out %lld\n", l, h[1] - h[0],
                                h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5]);
                    }
                    lp.end();
                    c->count_launch();
                    int rest = maxn - kb - 1;          // upper bound of the trailing size over the level's fronts
                    if (rest > 0) {
                        int tiles_u = (rest + BN - 1) / BN, tiles_l = (rest + BM - 1) / BM;
                        lp.begin(3);
                        { ProfScope ps_(c, PS_LU_PANELS, sq); k_lu_panels<<<dim3(nf_, tiles_u + tiles_l), 256, GEMM_SMEM, sq>>>(f0_, kb, tiles_u, d.fronts.ptr, arena, d.inv.ptr, sym, d.U.ptr, d.L.ptr);
                          lp.end(); lp.begin(4);
                          if (sym && transpose_kernel) k_lu_transpose_panel<<<dim3(nf_, (rest + 31) / 32), dim3(32, 8), 0, sq>>>(f0_, kb, d.fronts.ptr, arena); }
                        lp.end();
                        c->count_launch(1 + (sym && transpose_kernel));
                        if (kb + NB < gend) {
                            int wa = gend - kb - NB;       // strips inside the group
                            int tiles = ((rest + BM - 1) / BM) * ((wa + BN - 1) / BN) + ((wa + BM - 1) / BM) * tiles_u;
                            lp.begin(5);
                            { ProfScope ps_(c, PS_LU_PANELS, sq); k_lu_strip<<<nf_ <= 65535 ? dim3(tiles, nf_) : dim3(nf_, tiles), 256, GEMM_SMEM, sq>>>(f0_, kb, g0, d.fronts.ptr, arena, sym, nf_ <= 65535 ? 0 : 1); }
                            lp.end();
                            c->count_launch();
                            c->ctr[PC_LU_STRIP]++;
                        }
                    }
                }
                int rest = maxn - g0 - 1;
                if (rest > 0) {
                    int tiles_u = (rest + BN - 1) / BN, tiles_l = (rest + BM - 1) / BM;
                    const int prio = GB / BN;          // column tiles of the next group
                    if (la && gend < maxns && tiles_u > prio) {
                        // (the previous group's remainder wrote the columns this priority update continues on)
                        if (rem_pending) PN_CUDA(cudaStreamWaitEvent(sq, d.ev_rem[lane], 0));
                        k_lu_update<<<dim3(prio, tiles_l, nf_), 256, GEMM_SMEM, sq>>>(f0_, g0, d.fronts.ptr, arena, sym, 0, 0);
                        PN_CUDA(cudaEventRecord(d.ev_prio[lane], sq));
                        PN_CUDA(cudaStreamWaitEvent(su, d.ev_prio[lane], 0));
                        k_lu_update<<<dim3(tiles_u - prio, tiles_l, nf_), 256, GEMM_SMEM, su>>>(f0_, g0, d.fronts.ptr, arena, sym, 0, prio);
                        PN_CUDA(cudaEventRecord(d.ev_rem[lane], su));
                        rem_pending = true;
                        c->count_launch(2);
                    } else {
                        if (rem_pending) { PN_CUDA(cudaStreamWaitEvent(sq, d.ev_rem[lane], 0)); rem_pending = false; }
                        lp.begin(6);
                        { ProfScope ps_(c, PS_LU_UPDATE, sq); k_lu_update<<<nf_ <= 65535 ? dim3(tiles_u, tiles_l, nf_) : dim3(nf_, tiles_l, tiles_u), 256, GEMM_SMEM, sq>>>(f0_, g0, d.fronts.ptr, arena, sym, nf_ <= 65535 ? 0 : 1, 0); }
                        lp.end();
                        c->count_launch();
                    }
                    c->ctr[PC_LU_UPDATE]++;
                    if (g0 > 0) c->ctr[PC_LU_UPDATE_LATER]++;
                }
            }
            if (rem_pending) PN_CUDA(cudaStreamWaitEvent(sq, d.ev_rem[lane], 0));
        };
        // levels with many fronts: one CTA per front does all block steps (k_front_factor_fused)
        if (nfl >= fused_factor_min() && !lp.on) {
            { ProfScope ps_(c, PS_FACTOR_FUSED); k_front_factor_fused<<<nfl, 256, FUSED_FACTOR_SMEM, st>>>(first, d.fronts.ptr, arena, d.inv.ptr, d.status.ptr, sym, d.U.ptr, d.L.ptr); }
            c->count_launch();
            c->ctr[PC_FACTOR_FUSED]++;
            if (debug_levels) {
                cudaEventRecord(dbg1, st);
                cudaEventSynchronize(dbg1);
                float ms = 0;
                cudaEventElapsedTime(&ms, dbg0, dbg1);
                fprintf(stderr, "[pn] level %2d: %6d fronts, max ns %5d, max n %5d, arena %8.1f MB, %8.3f ms (fused per-front kernel)\n", l, nfl,
                        d.level_max_ns[l], d.level_max_n[l], d.level_f_total[l] * 8e-6, ms);
            }
            if (level_done) (*level_done)(l);
            continue;
        }
        int nsplit = 1;
        // (per-kernel profiling serialises the groups: overlapping kernels would be charged each other's time)
        if (!lp.on && !c->prof.on && nfl >= 2 && nfl <= d.split_max_fronts) nsplit = std::min(nfl, d.n_fstreams);
        if (nsplit == 1) {
            run_steps(first, nfl, st, 0);
        } else {
            c->ctr[PC_SPLIT_LEVELS]++;
            PN_CUDA(cudaEventRecord(d.ev_fork, st));
            for (int q = 0; q < nsplit; ++q) {
                int a = first + (int)((int64_t)nfl * q / nsplit), b = first + (int)((int64_t)nfl * (q + 1) / nsplit);
                PN_CUDA(cudaStreamWaitEvent(d.fstream[q], d.ev_fork, 0));
                run_steps(a, b - a, d.fstream[q], q);
                PN_CUDA(cudaEventRecord(d.ev_join[q], d.fstream[q]));
                PN_CUDA(cudaStreamWaitEvent(st, d.ev_join[q], 0));
            }
        }
        {
            lp.begin(7);
            ProfScope ps_(c, PS_COPY_FACTORS);
            int steps = (maxns + NB - 1) / NB;
            int row_tiles = (maxns + BM - 1) / BM;          // rows above the diagonal blocks only
            if (steps > 1) k_store_panels<<<dim3(nfl, steps, row_tiles), 256, GEMM_SMEM, st>>>(first, d.fronts.ptr, arena, d.inv.ptr, d.L.ptr);
            if (d.level_max_nb[l] > 0 && !(sym && !transpose_kernel)) {     // symmetric mode: written by the panel GEMMs
                int64_t per = (int64_t)maxns * d.level_max_nb[l];
                unsigned gy = (unsigned)std::min<int64_t>(std::max<int64_t>((per + 255) / 256, 1), 2048);
                k_copy_u12<<<dim3(nfl, gy), 256, 0, st>>>(first, d.fronts.ptr, arena, d.U.ptr);
            }
            lp.end();
            c->count_launch(2);
        }
        if (debug_levels) {
            cudaEventRecord(dbg1, st);
            cudaEventSynchronize(dbg1);
            float ms = 0;
            cudaEventElapsedTime(&ms, dbg0, dbg1);
            double fl = 0;
            for (int q = first; q < first + nfl; ++q) {
                double ns = d.h_fronts[q].ns, nb = d.h_fronts[q].nb;
                fl += 2.0 / 3.0 * ns * ns * ns + 2.0 * ns * ns * nb + 2.0 * ns * nb * nb;
            }
            fprintf(stderr, "[pn] level %2d: %6d fronts, max ns %5d, max n %5d, arena %8.1f MB, %8.3f ms, %7.2f GFLOP -> %6.2f TFLOP/s\n",
                    l, nfl, maxns, maxn, d.level_f_total[l] * 8e-6, ms, fl * 1e-9, fl / (ms * 1e-3) * 1e-12);
            lp.report(l);
        }
        if (level_done) (*level_done)(l);
    }
    if (debug_levels) { cudaEventDestroy(dbg0); cudaEventDestroy(dbg1); }
    if (dbg_clk) cudaFree(dbg_clk);
    int status = 0;
    PN_CUDA(cudaMemcpyAsync(&status, d.status.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    if (status == 1) throw StatusError(PN_ERR_SINGULAR, "zero or non-finite pivot in the stiffness factorisation");
}

// After the local forward sweep of the cut level: the update vectors (boundary rows of the subtree roots' workspaces)
// of all ranks, into the cut level's workspace arena where the first replicated level pulls them from.
static void exchange_update_vectors(Ctx *c, int s, const FrontDev *fr) {
    DirectSolver &d = *c->direct;
    DistPlan &P = d.dist;
    cudaStream_t st = c->stream;
    const int world = c->dist_world, rank = c->dist_rank;
    const int cf = d.level_ptr[P.cut];
    double *W = d.warena[P.cut & 1].ptr;
    const int64_t seg = std::max<int64_t>(P.uv_seg, 1) * s;
    P.send.resize(seg);
    P.recv.resize(seg * world);
    const unsigned gy = (unsigned)std::min<int64_t>(std::max<int64_t>(((int64_t)d.level_max_nb[P.cut] * s + 255) / 256, 1), 1024);
    const int mine = P.cut_hi[rank] - P.cut_lo[rank];
    if (mine > 0) {
        k_dist_uv<<<dim3(mine, gy), 256, 0, st>>>(P.cut_lo[rank], s, fr, cf, P.d_uv_off.ptr, W, P.send.ptr, 0);
        c->count_launch();
    }
    dist_allgather(c, seg);
    for (int r = 0; r < world; ++r) {
        const int cnt = P.cut_hi[r] - P.cut_lo[r];
        if (r == rank || cnt <= 0) continue;
        k_dist_uv<<<dim3(cnt, gy), 256, 0, st>>>(P.cut_lo[r], s, fr, cf, P.d_uv_off.ptr, W, P.recv.ptr + (int64_t)r * seg, 1);
        c->count_launch();
    }
}

// After the backward sweep: every rank's solved rows (its range of elimination positions) to every rank.
static void exchange_solved_rows(Ctx *c, double *X, int s) {
    DirectSolver &d = *c->direct;
    DistPlan &P = d.dist;
    cudaStream_t st = c->stream;
    const int world = c->dist_world, rank = c->dist_rank;
    const int64_t seg = std::max<int64_t>(P.x_seg, 1) * s;
    P.send.resize(seg);
    P.recv.resize(seg * world);
    const int64_t mine = (P.pos_hi[rank] - P.pos_lo[rank]) * s;
    if (mine > 0) {
        k_dist_x<<<grid_for(mine, 256), 256, 0, st>>>(P.pos_lo[rank], P.pos_hi[rank], s, d.pos_dof.ptr, X, P.send.ptr, 0);
        c->count_launch();
    }
    dist_allgather(c, seg);
    for (int r = 0; r < world; ++r) {
        const int64_t cnt = (P.pos_hi[r] - P.pos_lo[r]) * s;
        if (r == rank || cnt <= 0) continue;
        k_dist_x<<<grid_for(cnt, 256), 256, 0, st>>>(P.pos_lo[r], P.pos_hi[r], s, d.pos_dof.ptr, X, P.recv.ptr + (int64_t)r * seg, 1);
        c->count_launch();
    }
}

// k_ll_sweep levels: boundary rows by gemm_tile launches (PN_LL_GEMM_BOUNDARY=0 restores the all-in-one sweep kernel)
static bool ll_gemm_boundary() {
    static const bool v = !(getenv("PN_LL_GEMM_BOUNDARY") && atoi(getenv("PN_LL_GEMM_BOUNDARY")) == 0);
    return v;
}
static int ll_gemm_min_tiles() {
    static const int v = getenv("PN_LL_GEMM_MIN_TILES") ? atoi(getenv("PN_LL_GEMM_MIN_TILES")) : 128;   // measured on config 3: below ~128 tiles of 128 x 64 the long-K tiles leave most SMs idle
    return v;
}

// workspaces and the per-s copy of the front table (workspace offsets were stored per unit column)
static void solve_prepare(Ctx *c, int s) {
    DirectSolver &d = *c->direct;
    cudaStream_t st = c->stream;
    int64_t wmax = 0;
    for (int l = 0; l < d.n_levels; ++l) wmax = std::max(wmax, d.level_n_total[l]);
    d.warena[0].resize(wmax * s); d.warena[1].resize(wmax * s); d.yarena.resize(wmax * s);
    if (d.fronts_s != s) {
        std::vector<FrontDev> hf = d.h_fronts;
        for (auto &F : hf) F.w_off *= s;
        d.fronts_scaled.upload(hf.data(), (int64_t)hf.size(), st);
        PN_CUDA(cudaStreamSynchronize(st));
        d.fronts_s = s;
    }
}

// forward sweep of level l on stream st (the caller has run solve_prepare and, subtree-to-GPU, the exchange before the
// first replicated level)
static void solve_forward_level(Ctx *c, double *X, int s, int l, cudaStream_t st) {
    DirectSolver &d = *c->direct;
    const FrontDev *fr = d.fronts_scaled.ptr;
    double *Y = d.yarena.ptr;
    const int gz = (s + BN - 1) / BN;
    // (read on every call: the parity tests switch kernel families inside one process)
    const bool no_sm_sweeps = getenv("PN_NO_SMEM_SWEEPS") != nullptr;
    const bool no_ll_sweeps = getenv("PN_NO_LL_SWEEPS") != nullptr;
    const bool use_ll = !no_ll_sweeps && s <= 32 * LL_MAX_CHUNKS;
    const DistPlan &DP = d.dist;
    int first = DP.active ? DP.lo[l] : d.level_ptr[l], nfl = (DP.active ? DP.hi[l] : d.level_ptr[l + 1]) - first;
    if (!nfl) return;
    double *W = d.warena[l & 1].ptr;
    int maxn = d.level_max_n[l], maxns = d.level_max_ns[l];
    size_t sw_smem = 0;
    bool three = false;
    const int cw = no_sm_sweeps ? 0 : sweep_chunk(maxn, s, &sw_smem, &three);
    if (cw) {
        ProfScope ps_(c, PS_SOLVE_FUSED);
        const int H = (s + cw - 1) / cw;
        const unsigned grid = (unsigned)nfl * (unsigned)H;
#define PN_FWD(...) k_front_forward_sm<__VA_ARGS__><<<grid, 256, sw_smem, st>>>(first, s, H, fr, d.fidx.ptr, d.rel.ptr, d.child_ptr.ptr, d.child_idx.ptr, \
                                                                d.L.ptr, d.inv.ptr, X, W, d.warena[(l + 1) & 1].ptr)
        if (three) PN_FWD(32, 3, SW_PAD_ROWS_3); else if (cw == 32) PN_FWD(32); else if (cw == 16) PN_FWD(16); else PN_FWD(8);
#undef PN_FWD
        c->count_launch();
        c->ctr[PC_SWEEP_SMEM]++;
        return;
    }
    if (!use_ll && nfl >= FUSED_SOLVE_MIN_FRONTS) {     // fallback for more than 256 right-hand sides
        ProfScope ps_(c, PS_SOLVE_FUSED);
        k_front_forward<<<nfl, 256, GEMM_SMEM, st>>>(first, s, fr, d.fidx.ptr, d.rel.ptr, d.child_ptr.ptr, d.child_idx.ptr, d.L.ptr,
                                                       d.inv.ptr, X, W, d.warena[(l + 1) & 1].ptr, Y);
        c->count_launch();
        return;
    }
    unsigned gy = (unsigned)std::min<int64_t>(std::max<int64_t>(((int64_t)maxn * s + 255) / 256, 1), 1024);
    { ProfScope ps_(c, PS_SOLVE_GATHER); k_solve_gather<<<dim3(nfl, gy), 256, 0, st>>>(first, s, 0, fr, d.fidx.ptr, X, W); }
    c->count_launch();
    if (l > 0) {
        const bool own_children = DP.active && l <= DP.cut;
        int cfirst = own_children ? DP.lo[l - 1] : d.level_ptr[l - 1], ncl = (own_children ? DP.hi[l - 1] : d.level_ptr[l]) - cfirst;
        unsigned gyc = (unsigned)std::min<int64_t>(std::max<int64_t>(((int64_t)d.level_max_nb[l - 1] * s + 255) / 256, 1), 1024);
        for (int r = 0; r < d.level_max_children[l - 1] && ncl > 0; ++r) {
            { ProfScope ps_(c, PS_SOLVE_MISC); k_solve_extend_add<<<dim3(ncl, gyc), 256, 0, st>>>(cfirst, ncl, r, s, fr, d.rel.ptr,
                                                                d.warena[(l - 1) & 1].ptr, W); }
            c->count_launch();
        }
    }
    if (use_ll) {
        // boundary rows as full 128 x 64 GEMM tiles (K = all pivots of the front) after the flag-synchronised sweep of the
        // pivot tiles; the 16 x 16-per-warp tiles of k_ll_sweep are kept for the boundary only when the level has too few
        // boundary tiles to fill the GPU with 128-row tiles
        const int64_t gemm_tiles = (int64_t)nfl * ((d.level_max_nb[l] + BM - 1) / BM) * gz;
        const bool split = ll_gemm_boundary() && d.level_max_nb[l] > 0 && gemm_tiles >= ll_gemm_min_tiles();
        const int H = (s + 31) / 32, T = (maxns + NB - 1) / NB + (split ? 0 : (d.level_max_nb[l] + NB - 1) / NB);
        const int units = T * nfl * H;
        ProfScope ps_(c, PS_SOLVE_UPDATE);
        k_ll_sweep<<<std::min(units, d.ll_grid), 256, 0, st>>>(first, nfl, s, H, T, 0, fr, d.L.ptr, d.U.ptr, d.inv.ptr, W, Y,
                                                              d.ll_flags.ptr, ++d.ll_epoch, split ? 1 : 0);
        c->count_launch();
        c->ctr[PC_SWEEP_LL]++;
        if (split) {
            k_solve_forward_boundary<<<dim3(nfl, (d.level_max_nb[l] + BM - 1) / BM, gz), 256, GEMM_SMEM, st>>>(first, s, fr, d.L.ptr, W);
            c->count_launch();
            c->ctr[PC_SWEEP_LL_GEMM]++;
        }
    } else
    for (int kb = 0; kb < maxns; kb += NB) {
        int rest = std::max(maxn - kb - 1, 0);
        int tiles = 1 + (rest + BM - 1) / BM;
        { ProfScope ps_(c, PS_SOLVE_UPDATE); k_solve_step<<<dim3(nfl, tiles, gz), 256, GEMM_SMEM, st>>>(first, kb, s, 0, fr, d.L.ptr, d.inv.ptr, W, Y); }
        c->count_launch();
        c->ctr[PC_SWEEP_STEP]++;
    }
    unsigned gys = (unsigned)std::min<int64_t>(std::max<int64_t>(((int64_t)maxns * s + 255) / 256, 1), 1024);
    { ProfScope ps_(c, PS_SOLVE_GATHER); k_solve_store<<<dim3(nfl, gys), 256, 0, st>>>(first, s, fr, d.fidx.ptr, Y, X); }
    c->count_launch();
}

// solves in place: X holds the right-hand sides on entry ([n1][s] row-major).  forward_done: the forward sweep has
// already been issued (direct_solve runs it behind the factorisation).
static void solve_inplace(Ctx *c, double *X, int s, bool forward_done = false) {
    DirectSolver &d = *c->direct;
    cudaStream_t st = c->stream;
    solve_prepare(c, s);
    const FrontDev *fr = d.fronts_scaled.ptr;
    double *Y = d.yarena.ptr;
    int gz = (s + BN - 1) / BN;
    static const bool debug_levels = getenv("PN_DEBUG_LEVELS") != nullptr;
    // (read on every call: the parity tests switch kernel families inside one process)
    const bool no_sm_sweeps = getenv("PN_NO_SMEM_SWEEPS") != nullptr;
    const bool no_ll_sweeps = getenv("PN_NO_LL_SWEEPS") != nullptr;
    const bool use_ll = !no_ll_sweeps && s <= 32 * LL_MAX_CHUNKS;
    cudaEvent_t dbg[64];
    if (debug_levels) for (int i = 0; i < 64; ++i) cudaEventCreate(&dbg[i]);
    const DistPlan &DP = d.dist;
    // forward sweep (subtree-to-GPU: own subtrees up to the cut level, then every front)
    for (int l = 0; l < d.n_levels; ++l) {
        if (debug_levels && l < 31) cudaEventRecord(dbg[l], st);
        if (forward_done) continue;
        if (DP.active && l == DP.cut + 1) exchange_update_vectors(c, s, fr);
        solve_forward_level(c, X, s, l, st);
    }
    if (debug_levels) cudaEventRecord(dbg[31], st);
    // backward sweep
    for (int l = d.n_levels - 1; l >= 0; --l) {
        int first = DP.active ? DP.lo[l] : d.level_ptr[l], nfl = (DP.active ? DP.hi[l] : d.level_ptr[l + 1]) - first;
        if (debug_levels && l < 31) cudaEventRecord(dbg[32 + l + 1], st);
        if (!nfl) continue;
        double *W = d.warena[l & 1].ptr;
        int maxn = d.level_max_n[l], maxns = d.level_max_ns[l];
        size_t sw_smem = 0;
        bool three = false;
        const int cw = no_sm_sweeps ? 0 : sweep_chunk(maxn, s, &sw_smem, &three);
        if (cw) {
            ProfScope ps_(c, PS_SOLVE_FUSED);
            const int H = (s + cw - 1) / cw;
            const unsigned grid = (unsigned)nfl * (unsigned)H;
#define PN_BWD(...) k_front_backward_sm<__VA_ARGS__><<<grid, 256, sw_smem, st>>>(first, s, H, fr, d.fidx.ptr, d.L.ptr, d.U.ptr, d.inv.ptr, X)
            if (three) PN_BWD(32, 3, SW_PAD_ROWS_3); else if (cw == 32) PN_BWD(32); else if (cw == 16) PN_BWD(16); else PN_BWD(8);
#undef PN_BWD
            c->count_launch();
            c->ctr[PC_SWEEP_SMEM]++;
            continue;
        }
        if (!use_ll && nfl >= FUSED_SOLVE_MIN_FRONTS) {
            ProfScope ps_(c, PS_SOLVE_FUSED);
            k_front_backward<<<nfl, 256, GEMM_SMEM, st>>>(first, s, fr, d.fidx.ptr, d.L.ptr, d.U.ptr, d.inv.ptr, X, W, Y);
            c->count_launch();
            continue;
        }
        unsigned gy = (unsigned)std::min<int64_t>(std::max<int64_t>(((int64_t)maxn * s + 255) / 256, 1), 1024);
        { ProfScope ps_(c, PS_SOLVE_GATHER); k_solve_gather<<<dim3(nfl, gy), 256, 0, st>>>(first, s, 1, fr, d.fidx.ptr, X, W); }
        c->count_launch();
        if (use_ll) {
            const int H = (s + 31) / 32, T = (maxns + NB - 1) / NB;
            const int units = T * nfl * H;
            const int64_t gemm_tiles = (int64_t)nfl * ((maxns + BM - 1) / BM) * gz;
            const bool split = ll_gemm_boundary() && d.level_max_nb[l] > 0 && gemm_tiles >= ll_gemm_min_tiles();
            ProfScope ps_(c, PS_SOLVE_UPDATE);
            if (split) {         // W[0:ns] -= U12 W[ns:n] as full GEMM tiles (K = nb), then the sweep of the pivot tiles
                k_solve_boundary<<<dim3(nfl, (maxns + BM - 1) / BM, gz), 256, GEMM_SMEM, st>>>(first, s, fr, d.U.ptr, W);
                c->count_launch();
                c->ctr[PC_SWEEP_LL_GEMM]++;
            }
            k_ll_sweep<<<std::min(units, d.ll_grid), 256, 0, st>>>(first, nfl, s, H, T, 1, fr, d.L.ptr, d.U.ptr, d.inv.ptr, W, Y,
                                                                  d.ll_flags.ptr, ++d.ll_epoch, split ? 1 : 0);
            c->count_launch();
            c->ctr[PC_SWEEP_LL]++;
        } else {
        if (d.level_max_nb[l] > 0) {
            { ProfScope ps_(c, PS_SOLVE_UPDATE); k_solve_boundary<<<dim3(nfl, (maxns + BM - 1) / BM, gz), 256, GEMM_SMEM, st>>>(first, s, fr, d.U.ptr, W); }
            c->count_launch();
        }
        int last_kb = ((maxns - 1) / NB) * NB;
        for (int kb = last_kb; kb >= 0; kb -= NB) {
            int tiles = 1 + (kb + BM - 1) / BM;
            { ProfScope ps_(c, PS_SOLVE_UPDATE); k_solve_step<<<dim3(nfl, tiles, gz), 256, GEMM_SMEM, st>>>(first, kb, s, 1, fr, d.L.ptr, d.inv.ptr, W, Y); }
            c->count_launch();
            c->ctr[PC_SWEEP_STEP]++;
        }
        }
        unsigned gys = (unsigned)std::min<int64_t>(std::max<int64_t>(((int64_t)maxns * s + 255) / 256, 1), 1024);
        { ProfScope ps_(c, PS_SOLVE_GATHER); k_solve_store<<<dim3(nfl, gys), 256, 0, st>>>(first, s, fr, d.fidx.ptr, Y, X); }
        c->count_launch();
    }
    if (DP.active) exchange_solved_rows(c, X, s);
    if (debug_levels) {
        cudaEventRecord(dbg[32], st);
        cudaEventSynchronize(dbg[32]);
        for (int l = 0; l < d.n_levels && l < 31; ++l) {
            float f = 0, b = 0;
            cudaEventElapsedTime(&f, dbg[l], l + 1 < d.n_levels ? dbg[l + 1] : dbg[31]);
            cudaEventElapsedTime(&b, dbg[32 + l + 1], dbg[32 + l]);
            fprintf(stderr, "[pn] solve level %2d (s=%d): %6d fronts, forward %7.3f ms, backward %7.3f ms\n", l, s,
                    d.level_ptr[l + 1] - d.level_ptr[l], f, b);
        }
        for (int i = 0; i < 64; ++i) cudaEventDestroy(dbg[i]);
    }
}

// ---- residual / refinement helpers ------------------------------------------------------------------
__global__ void k_residual(int64_t n, const double *__restrict__ B, const double *__restrict__ AX, double *__restrict__ R) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) R[t] = B[t] - AX[t];
}
__global__ void k_axpy1(int64_t n, const double *__restrict__ dx, double *__restrict__ x) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) x[t] += dx[t];
}
__global__ void k_get_col(int64_t n, int s, int j, const double *__restrict__ M, double *__restrict__ v) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = M[i * s + j];
}
__global__ void k_set_col(int64_t n, int s, int j, const double *__restrict__ v, double *__restrict__ M) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) M[i * s + j] = v[i];
}

// ---- GMRES on K_j x_j = b_j for all columns at once, right-preconditioned with the stored factor (of Ke) -----------
// Y[i][j] += alpha[j] X[i][j]  /  Y[i][j] = alpha[j] X[i][j]
__global__ void k_col_axpy(int64_t n, int s, const double *__restrict__ alpha, const double *__restrict__ X, double *__restrict__ Y) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n * s) Y[t] = fma(alpha[t % s], X[t], Y[t]);
}
__global__ void k_col_scale(int64_t n, int s, const double *__restrict__ alpha, const double *__restrict__ X, double *__restrict__ Y) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n * s) Y[t] = alpha[t % s] * X[t];
}

static void gmres_columns(Ctx *c, const double *vals11, int s) {
    cudaStream_t st = c->stream;
    const CsrBlock &A = c->blk[0];
    const int64_t n1 = c->n1, ns = n1 * s;
    int m = (int)std::min<int64_t>(40, std::max<int64_t>(5, (int64_t)(4.0e9 / (8.0 * (double)std::max<int64_t>(ns, 1))) - 3));
    DevBuf<double> V, W, Z, zero, alpha;
    V.resize((int64_t)(m + 1) * ns); W.resize(ns); Z.resize(ns); zero.resize(ns); alpha.resize(s);
    zero.zero(st);
    std::vector<double> a(s), beta(s), bnorm(s), dots(s);
    std::vector<std::vector<double>> H(s, std::vector<double>((size_t)(m + 1) * m, 0.0)), cs(s, std::vector<double>(m)),
        sn(s, std::vector<double>(m)), g(s, std::vector<double>(m + 1));
    auto put_alpha = [&]() {
        PN_CUDA(cudaMemcpyAsync(alpha.ptr, a.data(), sizeof(double) * s, cudaMemcpyHostToDevice, st));
        PN_CUDA(cudaStreamSynchronize(st));
    };
    const unsigned grid = grid_for(ns, 256);
    // start from x = Ke^-1 b (the stationary iterate may have drifted along a divergent mode)
    PN_CUDA(cudaMemcpyAsync(c->X.ptr, c->B.ptr, sizeof(double) * ns, cudaMemcpyDeviceToDevice, st));
    solve_inplace(c, c->X.ptr, s);
    column_dots(c, n1, s, c->B.ptr, c->B.ptr, bnorm.data());
    for (int j = 0; j < s; ++j) bnorm[j] = std::sqrt(bnorm[j]);
    const double tol = 1e-13;
    double worst = 0.0;
    for (int cycle = 0; cycle < 12; ++cycle) {
        residual_dd(c, A, vals11, A.nnz, c->X.ptr, c->B.ptr, W.ptr, s, st);
        column_dots(c, n1, s, W.ptr, W.ptr, beta.data());
        worst = 0.0;
        for (int j = 0; j < s; ++j) {
            beta[j] = std::sqrt(beta[j]);
            if (!(beta[j] == beta[j])) beta[j] = INFINITY;
            worst = std::max(worst, bnorm[j] > 0 ? beta[j] / bnorm[j] : beta[j]);
        }
        if (worst <= 1e-12) break;
        if (!std::isfinite(worst)) break;
        for (int j = 0; j < s; ++j) {
            a[j] = beta[j] > 0 ? 1.0 / beta[j] : 0.0;
            std::fill(g[j].begin(), g[j].end(), 0.0);
            g[j][0] = beta[j];
            std::fill(H[j].begin(), H[j].end(), 0.0);
        }
        put_alpha();
        k_col_scale<<<grid, 256, 0, st>>>(n1, s, alpha.ptr, W.ptr, V.ptr);
        c->count_launch();
        int kdone = 0;
        for (int k = 0; k < m; ++k) {
            // w = K_j Ke^-1 v_k
            PN_CUDA(cudaMemcpyAsync(Z.ptr, V.ptr + (int64_t)k * ns, sizeof(double) * ns, cudaMemcpyDeviceToDevice, st));
            solve_inplace(c, Z.ptr, s);
            residual_dd(c, A, vals11, A.nnz, Z.ptr, zero.ptr, W.ptr, s, st);       // W = -K z
            for (int j = 0; j < s; ++j) a[j] = -1.0;
            put_alpha();
            k_col_scale<<<grid, 256, 0, st>>>(n1, s, alpha.ptr, W.ptr, W.ptr);
            c->count_launch();
            for (int i = 0; i <= k; ++i) {                                        // modified Gram-Schmidt
                column_dots(c, n1, s, W.ptr, V.ptr + (int64_t)i * ns, dots.data());
                for (int j = 0; j < s; ++j) { H[j][(size_t)i * m + k] = dots[j]; a[j] = -dots[j]; }
                put_alpha();
                k_col_axpy<<<grid, 256, 0, st>>>(n1, s, alpha.ptr, V.ptr + (int64_t)i * ns, W.ptr);
                c->count_launch();
            }
            column_dots(c, n1, s, W.ptr, W.ptr, dots.data());
            double est = 0.0;
            for (int j = 0; j < s; ++j) {
                const double hk1 = std::sqrt(dots[j]);
                H[j][(size_t)(k + 1) * m + k] = hk1;
                a[j] = hk1 > 1e-300 ? 1.0 / hk1 : 0.0;
                // Givens rotations of column k
                std::vector<double> &Hj = H[j];
                for (int i = 0; i < k; ++i) {
                    const double t0 = cs[j][i] * Hj[(size_t)i * m + k] + sn[j][i] * Hj[(size_t)(i + 1) * m + k];
                    Hj[(size_t)(i + 1) * m + k] = -sn[j][i] * Hj[(size_t)i * m + k] + cs[j][i] * Hj[(size_t)(i + 1) * m + k];
                    Hj[(size_t)i * m + k] = t0;
                }
                const double hkk = Hj[(size_t)k * m + k], r = std::hypot(hkk, hk1);
                cs[j][k] = r > 0 ? hkk / r : 1.0;
                sn[j][k] = r > 0 ? hk1 / r : 0.0;
                Hj[(size_t)k * m + k] = r;
                Hj[(size_t)(k + 1) * m + k] = 0.0;
                g[j][k + 1] = -sn[j][k] * g[j][k];
                g[j][k] = cs[j][k] * g[j][k];
                est = std::max(est, bnorm[j] > 0 ? std::fabs(g[j][k + 1]) / bnorm[j] : std::fabs(g[j][k + 1]));
            }
            put_alpha();
            k_col_scale<<<grid, 256, 0, st>>>(n1, s, alpha.ptr, W.ptr, V.ptr + (int64_t)(k + 1) * ns);
            c->count_launch();
            c->stats.iterations += 1;
            kdone = k + 1;
            if (est <= tol) break;
        }
        // x += Ke^-1 (V y),  H y = g
        std::vector<std::vector<double>> y(s, std::vector<double>(kdone, 0.0));
        for (int j = 0; j < s; ++j)
            for (int i = kdone - 1; i >= 0; --i) {
                double acc = g[j][i];
                for (int q = i + 1; q < kdone; ++q) acc -= H[j][(size_t)i * m + q] * y[j][q];
                const double dgn = H[j][(size_t)i * m + i];
                y[j][i] = dgn != 0.0 ? acc / dgn : 0.0;
            }
        Z.zero(st);
        for (int i = 0; i < kdone; ++i) {
            for (int j = 0; j < s; ++j) a[j] = y[j][i];
            put_alpha();
            k_col_axpy<<<grid, 256, 0, st>>>(n1, s, alpha.ptr, V.ptr + (int64_t)i * ns, Z.ptr);
            c->count_launch();
        }
        solve_inplace(c, Z.ptr, s);
        k_axpy1<<<grid, 256, 0, st>>>(ns, Z.ptr, c->X.ptr);
        c->count_launch();
    }
    c->ctr[PC_PDELTA_PERCOLUMN] += s;
    if (!(worst <= 1e-8))
        throw StatusError(PN_ERR_SINGULAR, "The stiffness matrix is singular, which indicates that the structure is unstable.");
}

void direct_release(Ctx *c) {
    if (c->direct) { try { join_host_analysis(*c->direct); } catch (...) {} }
    delete c->direct;
    c->direct = nullptr;
}

// The model changed: the elimination tree must be rebuilt, but the (grow-only) device buffers are kept so
// that re-analysing a model of the same size performs no cudaMalloc / cudaFree.
void direct_invalidate(Ctx *c) {
    if (!c->direct) return;
    try { join_host_analysis(*c->direct); } catch (...) {}     // the worker reads the host copy of the model
    c->direct->host_state = 0;
    c->direct->classes_rejected = false;
    c->direct->analysed = false;
    c->direct->shared_factor = false;
    c->direct->fronts_s = 0;
}

// The pattern was rebuilt for the same model (other symbolic flags): the device maps must be redone, the ordering stays.
void direct_invalidate_pattern(Ctx *c) {
    if (!c->direct) return;
    c->direct->analysed = false;
    c->direct->shared_factor = false;
    c->direct->fronts_s = 0;
}

// |d|^2 and |x|^2 per column for the refinement's contraction estimate.  Distributed solve: both arrays are replicated,
// every rank reduces its share of the rows and the partial sums are added across the ranks.
static void refinement_norms(Ctx *c, int64_t n1, int s, const double *Dv, const double *Xv, double *dd, double *xx) {
    int64_t r0 = 0, r1 = n1;
    direct_dist_row_share(c, n1, &r0, &r1);
    {
        const double *v[2] = {Dv + r0 * s, Xv + r0 * s};
        std::vector<double> all((size_t)2 * s);
        column_dots_multi(c, r1 - r0, s, 2, v, v, all.data());
        for (int j = 0; j < s; ++j) { dd[j] = all[j]; xx[j] = all[s + j]; }
    }
    if (r1 - r0 != n1) {
        std::vector<double> both((size_t)2 * s);
        for (int j = 0; j < s; ++j) { both[j] = dd[j]; both[s + j] = xx[j]; }
        direct_dist_sum(c, both.data(), 2 * s);
        for (int j = 0; j < s; ++j) { dd[j] = both[j]; xx[j] = both[s + j]; }
    }
}

void direct_solve(Ctx *c, const double *vals11, int percol, int s, const pn_solve_opts &o) {
    cudaStream_t st = c->stream;
    int64_t n1 = c->n1;
    if (!c->direct || !c->direct->analysed) {
        PhaseTimer tm(c, &c->stats.ms_symbolic);
        analyse(c);
    }
    DirectSolver &d = *c->direct;
    int refine = o.refine_steps >= 0 ? o.refine_steps : 1;
    DevBuf<double> &R = c->work_r;
    if (!percol) {
        // one factorisation per assembled matrix: later batches of right-hand sides reuse it
        bool forward_done = false;
        if (!(d.shared_factor && d.factor_epoch == c->assembly_epoch && d.factor_vals == vals11)) {
            d.shared_factor = false;
            // The forward sweep of level l needs the factor of levels <= l only, and the top of the tree is factorised
            // by latency-bound chains of block steps (root separator: one front, 47 dependent steps).  PN_SWEEP_OVERLAP=1
            // runs the first forward sweep one level behind the factorisation on its own low-priority stream.
            // Measured on config 3 (profiles/r2_dropped_experiments.md): factor 54.9 -> 61.0 ms, solve 31.3 -> 24.3 ms,
            // step 90.0 -> 89.6 ms -- the persistent sweep kernels hold the SM slots the block-step chain needs, so the
            // chain is delayed by as much as the sweep saves.  Off by default.
            static const bool want_overlap = getenv("PN_SWEEP_OVERLAP") != nullptr, debug_levels = getenv("PN_DEBUG_LEVELS") != nullptr;
            const bool overlap = want_overlap && !debug_levels && !d.dist.active && !c->prof.on && d.n_levels <= 64 && d.n_levels >= 3;
            PhaseTimer tm(c, &c->stats.ms_factor);
            if (overlap) {
                if (!d.sstream) {
                    int lo = 0, hi = 0;
                    cudaDeviceGetStreamPriorityRange(&lo, &hi);          // lo = numerically largest = lowest priority
                    PN_CUDA(cudaStreamCreateWithPriority(&d.sstream, cudaStreamNonBlocking, lo));
                    PN_CUDA(cudaEventCreateWithFlags(&d.ev_sweep, cudaEventDisableTiming));
                }
                solve_prepare(c, s);
                PN_CUDA(cudaMemcpyAsync(c->X.ptr, c->B.ptr, sizeof(double) * n1 * s, cudaMemcpyDeviceToDevice, st));
                double *Xp = c->X.ptr;
                std::function<void(int)> hook = [&, Xp](int l) {
                    if (!d.ev_level[l]) PN_CUDA(cudaEventCreateWithFlags(&d.ev_level[l], cudaEventDisableTiming));
                    PN_CUDA(cudaEventRecord(d.ev_level[l], st));
                    PN_CUDA(cudaStreamWaitEvent(d.sstream, d.ev_level[l], 0));
                    solve_forward_level(c, Xp, s, l, d.sstream);
                };
                // (a singular matrix throws out of factorise after the sweeps were queued: they run on garbage and
                // are harmless; the stream is drained before anything else uses the workspaces)
                try { factorise(c, vals11, &hook); }
                catch (...) { cudaStreamSynchronize(d.sstream); throw; }
                PN_CUDA(cudaEventRecord(d.ev_sweep, d.sstream));
                PN_CUDA(cudaStreamWaitEvent(st, d.ev_sweep, 0));
                forward_done = true;
                c->ctr[PC_SWEEP_OVERLAPPED]++;
            } else {
                factorise(c, vals11);
            }
            d.factor_epoch = c->assembly_epoch;
            d.factor_vals = vals11;
        }
        d.shared_factor = true;
        PhaseTimer tm(c, &c->stats.ms_solve);
        if (!forward_done) PN_CUDA(cudaMemcpyAsync(c->X.ptr, c->B.ptr, sizeof(double) * n1 * s, cudaMemcpyDeviceToDevice, st));
        solve_inplace(c, c->X.ptr, s, forward_done);
        if (refine > 0) R.resize(n1 * s);
        // Iterative refinement with a double-double residual: each correction d shrinks the forward error by
        // rho ~ cond(K11) * 2^-53.  rho is estimated per column as |d| / |x| of the first correction; refinement
        // stops as soon as the error predicted after applying d (rho * |d| / |x| = rho^2) is below 1e-11.
        std::vector<double> dd(s), xx(s);
        for (int it = 0; it < refine; ++it) {
            if (d.dist.active) residual_dd(c, c->blk[0], vals11, 0, c->X.ptr, c->B.ptr, R.ptr, s, st, d.dist.res_rows.ptr, 0, d.dist.n_res_rows);
            else residual_dd(c, c->blk[0], vals11, 0, c->X.ptr, c->B.ptr, R.ptr, s, st);
            c->stats.spmm_launches += 1;
            solve_inplace(c, R.ptr, s);
            refinement_norms(c, n1, s, R.ptr, c->X.ptr, dd.data(), xx.data());
            { ProfScope ps_(c, PS_SOLVE_MISC); k_axpy1<<<grid_for(n1 * s, 256), 256, 0, st>>>(n1 * s, R.ptr, c->X.ptr); }
            c->count_launch();
            c->stats.iterations += 1;
            double worst = 0.0;
            for (int j = 0; j < s; ++j) {
                double rho = xx[j] > 0 ? std::sqrt(dd[j] / xx[j]) : (dd[j] > 0 ? 1.0 : 0.0);
                if (!(rho == rho)) rho = 1.0;
                worst = std::max(worst, rho);
            }
            if (worst * worst < 1e-11) break;
        }
    } else {
        // One matrix per column (P-Delta: K_j = Ke + Kg(P_j), same pattern).  Kg is a small perturbation of Ke for
        // loads below buckling, and the factor of Ke is still in place from step 1 of the procedure, so the columns
        // are solved together by the stationary iteration  x <- x + Ke^-1 (b_j - K_j x)  with double-double
        // residuals: one multi-right-hand-side sweep per iteration instead of one factorisation per combination.
        // Contraction rate = spectral radius of Ke^-1 Kg (~ P / P_cr); if it is too slow the columns fall back to
        // their own factorisations below.
        if (d.shared_factor) {
            PhaseTimer tm(c, &c->stats.ms_solve);
            const CsrBlock &A = c->blk[0];
            R.resize(n1 * s);
            PN_CUDA(cudaMemcpyAsync(c->X.ptr, c->B.ptr, sizeof(double) * n1 * s, cudaMemcpyDeviceToDevice, st));
            solve_inplace(c, c->X.ptr, s);
            std::vector<double> dd(s), xx(s);
            double prev = 1e300;
            bool converged = false;
            for (int it = 0; it < 80; ++it) {
                if (d.dist.active) residual_dd(c, A, vals11, A.nnz, c->X.ptr, c->B.ptr, R.ptr, s, st, d.dist.res_rows.ptr, 0, d.dist.n_res_rows);
                else residual_dd(c, A, vals11, A.nnz, c->X.ptr, c->B.ptr, R.ptr, s, st);
                c->stats.spmm_launches += 1;
                solve_inplace(c, R.ptr, s);
                refinement_norms(c, n1, s, R.ptr, c->X.ptr, dd.data(), xx.data());
                { ProfScope ps_(c, PS_SOLVE_MISC); k_axpy1<<<grid_for(n1 * s, 256), 256, 0, st>>>(n1 * s, R.ptr, c->X.ptr); }
                c->count_launch();
                c->stats.iterations += 1;
                double worst = 0.0;
                for (int j = 0; j < s; ++j) {
                    double rho = xx[j] > 0 ? std::sqrt(dd[j] / xx[j]) : (dd[j] > 0 ? 1.0 : 0.0);
                    if (!(rho == rho)) rho = 1e300;
                    worst = std::max(worst, rho);
                }
                if (worst < 1e-12) { converged = true; break; }
                if (it >= 6 && worst > 0.7 * prev) break;          // contracting too slowly (or diverging)
                prev = worst;
            }
            if (converged) { c->ctr[PC_PDELTA_STATIONARY]++; return; }
            static const bool percol_lu = getenv("PN_PDELTA_PERCOLUMN_LU") != nullptr;
            if (!percol_lu) {
                // Contraction too slow: the load is near or above a buckling load of some mode (P / P_cr -> 1), where
                // Ke + Kg is near-singular or indefinite.  An LU without pivoting must not be trusted there (the reference
                // uses SuperLU with partial pivoting, Analysis.py:434), so the columns are finished by GMRES on K_j with the
                // factor of Ke as right preconditioner: Ke^-1 K_j = I + Ke^-1 Kg has one outlying eigenvalue per
                // buckling mode the load approaches, everything else clusters at 1.
                gmres_columns(c, vals11, s);
                return;
            }
        }
        d.col_in.resize(n1); d.col_out.resize(n1);
        DevBuf<double> ax;
        ax.resize(n1);
        const CsrBlock &A = c->blk[0];
        for (int j = 0; j < s; ++j) {
            const double *vj = vals11 + (int64_t)j * A.nnz;
            d.shared_factor = false;
            c->ctr[PC_PDELTA_PERCOLUMN]++;
            { PhaseTimer tm(c, &c->stats.ms_factor); factorise(c, vj); }
            PhaseTimer tm(c, &c->stats.ms_solve);
            k_get_col<<<grid_for(n1, 256), 256, 0, st>>>(n1, s, j, c->B.ptr, d.col_in.ptr);
            PN_CUDA(cudaMemcpyAsync(d.col_out.ptr, d.col_in.ptr, sizeof(double) * n1, cudaMemcpyDeviceToDevice, st));
            solve_inplace(c, d.col_out.ptr, 1);
            for (int it = 0; it < refine; ++it) {
                residual_dd(c, A, vj, 0, d.col_out.ptr, d.col_in.ptr, ax.ptr, 1, st);
                solve_inplace(c, ax.ptr, 1);
                { ProfScope ps_(c, PS_SOLVE_MISC); k_axpy1<<<grid_for(n1, 256), 256, 0, st>>>(n1, ax.ptr, d.col_out.ptr); }
                c->count_launch();
            }
            k_set_col<<<grid_for(n1, 256), 256, 0, st>>>(n1, s, j, d.col_out.ptr, c->X.ptr);
            c->count_launch(2);
        }
    }
}

}  // namespace pn
