// Sparse direct solver for K11 X = B: multifrontal LU without pivoting on a symmetric pattern.
//
//   ordering / fronts : pn_direct_symbolic.cpp (nested dissection of the node graph by coordinates)
//   numeric           : level by level up the separator tree; every front is a dense row-major
//                       (ns+nb)^2 matrix in a per-level arena.  Right-looking partial LU in block
//                       steps of NB pivots: diagonal block (one CTA), row/column panels, rank-NB update
//                       of the trailing matrix (tiled FP64 GEMM).  The trailing nb x nb block is the
//                       update matrix, extend-added into the parent in a fixed child order
//                       (deterministic, no atomics).
//   solve             : all right-hand sides at once; forward sweep passes update vectors up the
//                       tree like the update matrices, backward sweep reads ancestors' results.
//
// LU (not Cholesky) because the reference's K is not exactly symmetric: B^T H B is summed in a fixed
// order per element and the orthotropic membrane matrix is unsymmetric when kx_mod != ky_mod
// (Quad3D.py:517-519).  K11 is positive definite for a stable structure, so no pivoting is needed;
// a zero / non-finite pivot is reported as PN_ERR_SINGULAR (Analysis.py:172, 238-239).
#include "pn_internal.h"
#include "pn_direct_symbolic.h"

#include <algorithm>

namespace pn {

static inline unsigned grid_for(int64_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

constexpr int NB = 32;        // pivots per block step
constexpr int TILE = 64;      // GEMM tile edge
constexpr int KC = 16;        // GEMM k chunk

struct FrontDev {
    int64_t f_off;      // front matrix offset in its level arena
    int64_t w_off;      // solve workspace offset in its level arena ((ns+nb) x s)
    int64_t l_off;      // permanent L panel: (ns+nb) x ns, ld = ns (holds L11\U11 on top)
    int64_t u_off;      // permanent U12: ns x nb, ld = nb
    int64_t idx_off;    // into fidx / fpos
    int64_t rel_off;    // into rel
    int32_t ns, nb, parent, child_rank;
};

struct DirectSolver {
    DirectSym sym;
    bool analysed = false;
    int n_levels = 0;
    std::vector<int32_t> level_ptr;           // fronts are stored level-major on the device
    std::vector<int32_t> level_max_ns, level_max_n, level_max_nb, level_max_children;
    std::vector<int64_t> level_f_total, level_n_total;   // sum n^2, sum n per level
    std::vector<FrontDev> h_fronts;           // level-major
    DevBuf<FrontDev> fronts;
    DevBuf<int32_t> fidx, fpos, rel, perm_pos, dof_front_lm;   // dof_front in level-major numbering
    DevBuf<int32_t> ent_front;                // per K11 entry: owning front (level-major id)
    DevBuf<uint32_t> ent_dst;                 // offset inside the front matrix
    DevBuf<uint8_t> ent_level;
    DevBuf<double> arena[2];
    DevBuf<double> warena[2];
    DevBuf<double> L, U;
    DevBuf<int> status;                       // 0 ok, 1 zero / non-finite pivot, 2 entry outside front
    DevBuf<double> col_in, col_out;           // per-column solves (P-Delta)
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
                             const uint8_t *__restrict__ front_level, int32_t *__restrict__ ent_front,
                             uint32_t *__restrict__ ent_dst, uint8_t *__restrict__ ent_level, int *__restrict__ status) {
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
    ent_dst[k] = (uint32_t)li * (uint32_t)n + (uint32_t)lj;
    ent_level[k] = front_level[f];
}

// ------------------------------------------------------------------------------------------------
// numeric factorisation kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_assemble_entries(int64_t nnz, int level, const uint8_t *__restrict__ ent_level,
                                   const int32_t *__restrict__ ent_front, const uint32_t *__restrict__ ent_dst,
                                   const FrontDev *__restrict__ fronts, const double *__restrict__ vals,
                                   double *__restrict__ arena) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nnz || ent_level[k] != level) return;
    arena[fronts[ent_front[k]].f_off + ent_dst[k]] = vals[k];
}

// Adds the update matrix of every child with the given rank into its parent (one level up).
// grid.x = child index within the child level, grid.y = tile over the nb x nb update matrix.
__global__ void k_extend_add(int first_child, int n_children, int rank, const FrontDev *__restrict__ fronts,
                             const int32_t *__restrict__ rel, const double *__restrict__ child_arena,
                             double *__restrict__ parent_arena) {
    int ci = blockIdx.x;
    if (ci >= n_children) return;
    FrontDev C = fronts[first_child + ci];
    if (C.parent < 0 || C.child_rank != rank || C.nb == 0) return;
    FrontDev P = fronts[C.parent];
    int nc = C.ns + C.nb, np = P.ns + P.nb;
    const int32_t *r = rel + C.rel_off;
    int64_t total = (int64_t)C.nb * C.nb;
    for (int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.y * blockDim.x) {
        int a = (int)(t / C.nb), b = (int)(t - (int64_t)a * C.nb);
        parent_arena[P.f_off + (int64_t)r[a] * np + r[b]] += child_arena[C.f_off + (int64_t)(C.ns + a) * nc + C.ns + b];
    }
}

// LU of the diagonal block [kb, kb+bs) of every front in the level (no pivoting), in shared memory.
__global__ void __launch_bounds__(256)
k_lu_diag(int first, int kb, const FrontDev *__restrict__ fronts, double *__restrict__ arena, int *__restrict__ status) {
    __shared__ double D[NB][NB + 1];
    FrontDev F = fronts[first + blockIdx.x];
    if (kb >= F.ns) return;
    int bs = min(NB, F.ns - kb);
    int n = F.ns + F.nb;
    double *A = arena + F.f_off + (int64_t)kb * n + kb;
    int t = threadIdx.x;
    for (int e = t; e < bs * bs; e += blockDim.x) D[e / bs][e % bs] = A[(int64_t)(e / bs) * n + e % bs];
    __syncthreads();
    for (int p = 0; p < bs; ++p) {
        double piv = D[p][p];
        if (t == 0 && (piv == 0.0 || !(fabs(piv) <= 1.7976931348623157e308))) atomicMax(status, 1);
        __syncthreads();
        for (int i = p + 1 + t; i < bs; i += blockDim.x) D[i][p] = D[i][p] / piv;
        __syncthreads();
        int rem = bs - p - 1;
        for (int e = t; e < rem * rem; e += blockDim.x) {
            int i = p + 1 + e / rem, j = p + 1 + e % rem;
            D[i][j] = fma(-D[i][p], D[p][j], D[i][j]);
        }
        __syncthreads();
    }
    for (int e = t; e < bs * bs; e += blockDim.x) A[(int64_t)(e / bs) * n + e % bs] = D[e / bs][e % bs];
}

// Row panel U[kb:kb+bs, r0:n] = L_kk^-1 F[...] (thread per column) and column panel
// L[r0:n, kb:kb+bs] = F[...] U_kk^-1 (thread per row).  grid.x = front, grid.y = tile of 128 columns
// / rows; the first half of grid.y does the row panel, the second half the column panel.
__global__ void __launch_bounds__(128)
k_lu_panels(int first, int kb, int tiles, const FrontDev *__restrict__ fronts, double *__restrict__ arena) {
    __shared__ double D[NB][NB + 1];
    FrontDev F = fronts[first + blockIdx.x];
    if (kb >= F.ns) return;
    int bs = min(NB, F.ns - kb);
    int n = F.ns + F.nb;
    int r0 = kb + bs;
    int rest = n - r0;
    int tile = blockIdx.y % tiles;
    bool col_panel = blockIdx.y >= tiles;
    if (tile * 128 >= rest) return;
    double *A = arena + F.f_off;
    for (int e = threadIdx.x; e < bs * bs; e += blockDim.x) D[e / bs][e % bs] = A[(int64_t)(kb + e / bs) * n + kb + e % bs];
    __syncthreads();
    int q = tile * 128 + threadIdx.x;
    if (q >= rest) return;
    double v[NB];
    if (!col_panel) {
        int c = r0 + q;
        for (int i = 0; i < bs; ++i) {
            double acc = A[(int64_t)(kb + i) * n + c];
            for (int j = 0; j < i; ++j) acc = fma(-D[i][j], v[j], acc);
            v[i] = acc;
        }
        for (int i = 0; i < bs; ++i) A[(int64_t)(kb + i) * n + c] = v[i];
    } else {
        double *row = A + (int64_t)(r0 + q) * n + kb;
        for (int j = 0; j < bs; ++j) {
            double acc = row[j];
            for (int i = 0; i < j; ++i) acc = fma(-v[i], D[i][j], acc);
            v[j] = acc / D[j][j];
        }
        for (int j = 0; j < bs; ++j) row[j] = v[j];
    }
}

// C[m x n] -= A[m x k] B[k x n], all row-major with leading dimensions; 64x64 tile per CTA,
// 256 threads, 4x4 outputs per thread, k staged through shared memory in chunks of KC.
__device__ __forceinline__ void gemm_tile_sub(const double *__restrict__ A, int lda, const double *__restrict__ B, int ldb,
                                              double *__restrict__ C, int ldc, int m, int n, int k, int tile_m, int tile_n) {
    __shared__ double sA[KC][TILE + 1];      // transposed: sA[kk][row]
    __shared__ double sB[KC][TILE];
    int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    int row0 = tile_m * TILE, col0 = tile_n * TILE;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < k; k0 += KC) {
        for (int e = threadIdx.x; e < TILE * KC; e += 256) {
            int r = e / KC, kk = e % KC;
            int gr = row0 + r, gk = k0 + kk;
            sA[kk][r] = (gr < m && gk < k) ? A[(int64_t)gr * lda + gk] : 0.0;
        }
        for (int e = threadIdx.x; e < KC * TILE; e += 256) {
            int kk = e / TILE, cc = e % TILE;
            int gk = k0 + kk, gc = col0 + cc;
            sB[kk][cc] = (gk < k && gc < n) ? B[(int64_t)gk * ldb + gc] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gr = row0 + ty * 4 + i;
        if (gr >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gc = col0 + tx + 16 * j;
            if (gc < n) C[(int64_t)gr * ldc + gc] -= acc[i][j];
        }
    }
}

// trailing update of the partial LU: F[r0:, r0:] -= L[r0:, kb:kb+bs] U[kb:kb+bs, r0:]
__global__ void __launch_bounds__(256)
k_lu_update(int first, int kb, const FrontDev *__restrict__ fronts, double *__restrict__ arena) {
    FrontDev F = fronts[first + blockIdx.x];
    if (kb >= F.ns) return;
    int bs = min(NB, F.ns - kb);
    int n = F.ns + F.nb;
    int r0 = kb + bs, rest = n - r0;
    if ((int)blockIdx.y * TILE >= rest || (int)blockIdx.z * TILE >= rest) return;
    double *A = arena + F.f_off;
    gemm_tile_sub(A + (int64_t)r0 * n + kb, n, A + (int64_t)kb * n + r0, n, A + (int64_t)r0 * n + r0, n, rest, rest, bs,
                  blockIdx.y, blockIdx.z);
}

// copies the factor parts of every front of a level to permanent storage
__global__ void k_copy_factors(int first, const FrontDev *__restrict__ fronts, const double *__restrict__ arena,
                               double *__restrict__ L, double *__restrict__ U) {
    FrontDev F = fronts[first + blockIdx.x];
    int n = F.ns + F.nb;
    int64_t totalL = (int64_t)n * F.ns, totalU = (int64_t)F.ns * F.nb;
    const double *A = arena + F.f_off;
    for (int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; t < totalL + totalU; t += (int64_t)gridDim.y * blockDim.x) {
        if (t < totalL) {
            int64_t r = t / F.ns; int cidx = (int)(t - r * F.ns);
            L[F.l_off + t] = A[r * n + cidx];
        } else {
            int64_t u = t - totalL;
            int64_t r = u / F.nb; int cidx = (int)(u - r * F.nb);
            U[F.u_off + u] = A[r * n + F.ns + cidx];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// solve kernels (W = per-front workspace (ns+nb) x s, row-major)
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
                              const double *__restrict__ W, double *__restrict__ X) {
    FrontDev F = fronts[first + blockIdx.x];
    int64_t total = (int64_t)F.ns * s;
    for (int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.y * blockDim.x) {
        int r = (int)(t / s), j = (int)(t - (int64_t)r * s);
        X[(int64_t)fidx[F.idx_off + r] * s + j] = W[F.w_off + t];
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
// forward: w[kb:kb+bs] <- L_kk^-1 w[kb:kb+bs] (unit lower); backward: w[kb:kb+bs] <- U_kk^-1 w[kb:kb+bs]
__global__ void __launch_bounds__(128)
k_solve_diag(int first, int kb, int s, int backward, const FrontDev *__restrict__ fronts, const double *__restrict__ L,
             double *__restrict__ W) {
    __shared__ double D[NB][NB + 1];
    FrontDev F = fronts[first + blockIdx.x];
    if (kb >= F.ns) return;
    int bs = min(NB, F.ns - kb);
    const double *Lp = L + F.l_off + (int64_t)kb * F.ns + kb;
    for (int e = threadIdx.x; e < bs * bs; e += blockDim.x) D[e / bs][e % bs] = Lp[(int64_t)(e / bs) * F.ns + e % bs];
    __syncthreads();
    int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= s) return;
    double *w = W + F.w_off + (int64_t)kb * s + j;
    double v[NB];
    if (!backward) {
        for (int i = 0; i < bs; ++i) {
            double acc = w[(int64_t)i * s];
            for (int p = 0; p < i; ++p) acc = fma(-D[i][p], v[p], acc);
            v[i] = acc;
        }
    } else {
        for (int i = bs - 1; i >= 0; --i) {
            double acc = w[(int64_t)i * s];
            for (int p = i + 1; p < bs; ++p) acc = fma(-D[i][p], v[p], acc);
            v[i] = acc / D[i][i];
        }
    }
    for (int i = 0; i < bs; ++i) w[(int64_t)i * s] = v[i];
}
// mode 0: forward update   W[r0:n]  -= Lpanel[r0:n, kb:kb+bs] W[kb:kb+bs]
// mode 1: backward update  W[0:kb]  -= U11[0:kb, kb:kb+bs]    W[kb:kb+bs]
// mode 2: boundary update  W[0:ns]  -= U12[ns x nb]           W[ns:n]
__global__ void __launch_bounds__(256)
k_solve_update(int first, int kb, int s, int mode, const FrontDev *__restrict__ fronts, const double *__restrict__ L,
               const double *__restrict__ U, double *__restrict__ W) {
    FrontDev F = fronts[first + blockIdx.x];
    int n = F.ns + F.nb;
    double *w = W + F.w_off;
    if (mode == 2) {
        if (F.nb == 0 || (int)blockIdx.y * TILE >= F.ns || (int)blockIdx.z * TILE >= s) return;
        gemm_tile_sub(U + F.u_off, F.nb, w + (int64_t)F.ns * s, s, w, s, F.ns, s, F.nb, blockIdx.y, blockIdx.z);
        return;
    }
    if (kb >= F.ns) return;
    int bs = min(NB, F.ns - kb);
    if (mode == 0) {
        int r0 = kb + bs, rest = n - r0;
        if ((int)blockIdx.y * TILE >= rest || (int)blockIdx.z * TILE >= s) return;
        gemm_tile_sub(L + F.l_off + (int64_t)r0 * F.ns + kb, F.ns, w + (int64_t)kb * s, s, w + (int64_t)r0 * s, s, rest, s, bs,
                      blockIdx.y, blockIdx.z);
    } else {
        if ((int)blockIdx.y * TILE >= kb || (int)blockIdx.z * TILE >= s) return;
        gemm_tile_sub(L + F.l_off + kb, F.ns, w + (int64_t)kb * s, s, w, s, kb, s, bs, blockIdx.y, blockIdx.z);
    }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
static void analyse(Ctx *c) {
    if (!c->direct) c->direct = new DirectSolver();
    DirectSolver &d = *c->direct;
    cudaStream_t st = c->stream;
    const Model &m = c->model;
    // vertex graph from element connectivity (a superset of the K11 pattern)
    std::vector<int32_t> newidx(c->n_dof);
    PN_CUDA(cudaMemcpyAsync(newidx.data(), c->newidx.ptr, sizeof(int32_t) * c->n_dof, cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    std::vector<ConnView> conn;
    conn.push_back({m.n_springs, 2, m.h_spring_ij.data(), m.h_spring_active.data()});
    conn.push_back({m.n_members, 2, m.h_mem_ij.data(), m.h_mem_active.data()});
    conn.push_back({m.n_quads, 4, m.h_quad_ijmn.data(), nullptr});
    conn.push_back({m.n_plates, 4, m.h_plate_ijmn.data(), nullptr});
    std::vector<int32_t> vstart, adj;
    std::vector<double> coords;
    std::vector<int64_t> adj_ptr;
    build_vertex_graph(m.n_nodes, newidx.data(), m.h_xyz.data(), conn, vstart, coords, adj_ptr, adj);
    int32_t nv = (int32_t)vstart.size() - 1;
    direct_symbolic(c->n1, nv, vstart.data(), coords.data(), adj_ptr.data(), adj.data(), 24, d.sym);
    const DirectSym &S = d.sym;

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
    int64_t l_total = 0, u_total = 0;
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
            F.f_off = foff; foff += n * n;
            F.w_off = noff; noff += n;             // scaled by s at solve time
            F.l_off = l_total; l_total += n * fs.ns;
            F.u_off = u_total; u_total += (int64_t)fs.ns * fs.nb;
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
    d.fronts.upload(d.h_fronts.data(), nf, st);
    d.fidx.upload(S.fidx.data(), (int64_t)S.fidx.size(), st);
    d.fpos.upload(S.fpos.data(), (int64_t)S.fpos.size(), st);
    d.rel.upload(S.rel.data(), (int64_t)S.rel.size(), st);
    d.perm_pos.upload(S.perm_pos.data(), (int64_t)S.perm_pos.size(), st);
    std::vector<int32_t> dof_front_lm(S.dof_front.size());
    for (size_t i = 0; i < dof_front_lm.size(); ++i) dof_front_lm[i] = lm_of[S.dof_front[i]];
    d.dof_front_lm.upload(dof_front_lm.data(), (int64_t)dof_front_lm.size(), st);
    DevBuf<uint8_t> d_front_level;
    d_front_level.upload(front_level.data(), nf, st);
    d.L.resize(l_total); d.U.resize(u_total);
    d.status.resize(1);
    d.status.zero(st);
    // per-entry destination maps
    const CsrBlock &A = c->blk[0];
    DevBuf<int32_t> ent_row;
    ent_row.resize(A.nnz);
    d.ent_front.resize(A.nnz); d.ent_dst.resize(A.nnz); d.ent_level.resize(A.nnz);
    if (A.nnz) {
        k_entry_rows<<<grid_for(A.rows, 128), 128, 0, st>>>(A.rows, A.indptr.ptr, ent_row.ptr);
        k_entry_maps<<<grid_for(A.nnz, 256), 256, 0, st>>>(A.nnz, ent_row.ptr, A.indices.ptr, d.perm_pos.ptr,
                                                            d.dof_front_lm.ptr, d.fronts.ptr, d.fpos.ptr, d_front_level.ptr,
                                                            d.ent_front.ptr, d.ent_dst.ptr, d.ent_level.ptr, d.status.ptr);
        c->count_launch(2);
    }
    int status = 0;
    PN_CUDA(cudaMemcpyAsync(&status, d.status.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    if (status == 2) throw ArgError("direct solver: a K11 entry falls outside its front (symbolic analysis bug)");
    int64_t amax = 0;
    for (int l = 0; l < d.n_levels; ++l) amax = std::max(amax, d.level_f_total[l]);
    d.arena[0].resize(amax); d.arena[1].resize(amax);
    d.analysed = true;
    c->stats.factor_nnz = S.factor_entries;
    c->stats.factor_flops = S.factor_flops;
    double upd = 0.0;
    for (const FrontSym &fs : S.fronts) {
        int64_t n = fs.ns + fs.nb;
        for (int kb = 0; kb < fs.ns; kb += NB) {
            double bs = std::min(NB, fs.ns - kb), rest = (double)(n - kb) - bs;
            upd += 2.0 * bs * rest * rest;
        }
    }
    c->stats.update_flops = upd;
    c->stats.solve_flops = 2.0 * (double)S.factor_entries;
}

static void factorise(Ctx *c, const double *vals11) {
    DirectSolver &d = *c->direct;
    cudaStream_t st = c->stream;
    const CsrBlock &A = c->blk[0];
    d.status.zero(st);
    for (int l = 0; l < d.n_levels; ++l) {
        int first = d.level_ptr[l], nfl = d.level_ptr[l + 1] - first;
        if (!nfl) continue;
        double *arena = d.arena[l & 1].ptr;
        PN_CUDA(cudaMemsetAsync(arena, 0, sizeof(double) * d.level_f_total[l], st));
        { ProfScope ps_(c, PS_FRONT_ASSEMBLE); k_assemble_entries<<<grid_for(A.nnz, 256), 256, 0, st>>>(A.nnz, l, d.ent_level.ptr, d.ent_front.ptr, d.ent_dst.ptr,
                                                                  d.fronts.ptr, vals11, arena); }
        c->count_launch();
        if (l > 0) {
            int cfirst = d.level_ptr[l - 1], ncl = d.level_ptr[l] - cfirst;
            int64_t nbmax = d.level_max_nb[l - 1];
            unsigned gy = (unsigned)std::min<int64_t>(std::max<int64_t>((nbmax * nbmax + 255) / 256, 1), 1024);
            for (int r = 0; r < d.level_max_children[l - 1]; ++r) {
                { ProfScope ps_(c, PS_EXTEND_ADD); k_extend_add<<<dim3(ncl, gy), 256, 0, st>>>(cfirst, ncl, r, d.fronts.ptr, d.rel.ptr, d.arena[(l - 1) & 1].ptr,
                                                             arena); }
                c->count_launch();
            }
        }
        int maxns = d.level_max_ns[l], maxn = d.level_max_n[l];
        for (int kb = 0; kb < maxns; kb += NB) {
            { ProfScope ps_(c, PS_LU_DIAG); k_lu_diag<<<nfl, 256, 0, st>>>(first, kb, d.fronts.ptr, arena, d.status.ptr); }
            int rest = maxn - kb - 1;
            if (rest > 0) {
                int tiles = (rest + 127) / 128;
                { ProfScope ps_(c, PS_LU_PANELS); k_lu_panels<<<dim3(nfl, 2 * tiles), 128, 0, st>>>(first, kb, tiles, d.fronts.ptr, arena); }
                int gt = (rest + TILE - 1) / TILE;
                { ProfScope ps_(c, PS_LU_UPDATE); k_lu_update<<<dim3(nfl, gt, gt), 256, 0, st>>>(first, kb, d.fronts.ptr, arena); }
                c->count_launch(2);
            }
            c->count_launch();
        }
        int64_t per = (int64_t)maxn * std::max(maxns, 1) * 2;
        unsigned gy = (unsigned)std::min<int64_t>(std::max<int64_t>((per + 255) / 256, 1), 2048);
        { ProfScope ps_(c, PS_COPY_FACTORS); k_copy_factors<<<dim3(nfl, gy), 256, 0, st>>>(first, d.fronts.ptr, arena, d.L.ptr, d.U.ptr); }
        c->count_launch();
    }
    int status = 0;
    PN_CUDA(cudaMemcpyAsync(&status, d.status.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    if (status == 1) throw StatusError(PN_ERR_SINGULAR, "zero or non-finite pivot in the stiffness factorisation");
}

// solves in place: X holds the right-hand sides on entry ([n1][s] row-major)
static void solve_inplace(Ctx *c, double *X, int s) {
    DirectSolver &d = *c->direct;
    cudaStream_t st = c->stream;
    int64_t wmax = 0;
    for (int l = 0; l < d.n_levels; ++l) wmax = std::max(wmax, d.level_n_total[l]);
    d.warena[0].resize(wmax * s); d.warena[1].resize(wmax * s);
    // workspace offsets were stored per unit column; scale once per s
    std::vector<FrontDev> hf = d.h_fronts;
    for (auto &F : hf) F.w_off *= s;
    DevBuf<FrontDev> fr;
    fr.upload(hf.data(), (int64_t)hf.size(), st);
    int gz = (s + TILE - 1) / TILE;
    // forward sweep
    for (int l = 0; l < d.n_levels; ++l) {
        int first = d.level_ptr[l], nfl = d.level_ptr[l + 1] - first;
        if (!nfl) continue;
        double *W = d.warena[l & 1].ptr;
        int maxn = d.level_max_n[l], maxns = d.level_max_ns[l];
        unsigned gy = (unsigned)std::min<int64_t>(std::max<int64_t>(((int64_t)maxn * s + 255) / 256, 1), 1024);
        { ProfScope ps_(c, PS_SOLVE_GATHER); k_solve_gather<<<dim3(nfl, gy), 256, 0, st>>>(first, s, 0, fr.ptr, d.fidx.ptr, X, W); }
        c->count_launch();
        if (l > 0) {
            int cfirst = d.level_ptr[l - 1], ncl = d.level_ptr[l] - cfirst;
            unsigned gyc = (unsigned)std::min<int64_t>(std::max<int64_t>(((int64_t)d.level_max_nb[l - 1] * s + 255) / 256, 1), 1024);
            for (int r = 0; r < d.level_max_children[l - 1]; ++r) {
                { ProfScope ps_(c, PS_SOLVE_MISC); k_solve_extend_add<<<dim3(ncl, gyc), 256, 0, st>>>(cfirst, ncl, r, s, fr.ptr, d.rel.ptr,
                                                                    d.warena[(l - 1) & 1].ptr, W); }
                c->count_launch();
            }
        }
        for (int kb = 0; kb < maxns; kb += NB) {
            { ProfScope ps_(c, PS_SOLVE_DIAG); k_solve_diag<<<dim3(nfl, (s + 127) / 128), 128, 0, st>>>(first, kb, s, 0, fr.ptr, d.L.ptr, W); }
            int rest = maxn - kb - 1;
            if (rest > 0) {
                { ProfScope ps_(c, PS_SOLVE_UPDATE); k_solve_update<<<dim3(nfl, (rest + TILE - 1) / TILE, gz), 256, 0, st>>>(first, kb, s, 0, fr.ptr, d.L.ptr,
                                                                                        d.U.ptr, W); }
                c->count_launch();
            }
            c->count_launch();
        }
        unsigned gys = (unsigned)std::min<int64_t>(std::max<int64_t>(((int64_t)maxns * s + 255) / 256, 1), 1024);
        { ProfScope ps_(c, PS_SOLVE_GATHER); k_solve_store<<<dim3(nfl, gys), 256, 0, st>>>(first, s, fr.ptr, d.fidx.ptr, W, X); }
        c->count_launch();
    }
    // backward sweep
    for (int l = d.n_levels - 1; l >= 0; --l) {
        int first = d.level_ptr[l], nfl = d.level_ptr[l + 1] - first;
        if (!nfl) continue;
        double *W = d.warena[l & 1].ptr;
        int maxn = d.level_max_n[l], maxns = d.level_max_ns[l];
        unsigned gy = (unsigned)std::min<int64_t>(std::max<int64_t>(((int64_t)maxn * s + 255) / 256, 1), 1024);
        { ProfScope ps_(c, PS_SOLVE_GATHER); k_solve_gather<<<dim3(nfl, gy), 256, 0, st>>>(first, s, 1, fr.ptr, d.fidx.ptr, X, W); }
        c->count_launch();
        if (d.level_max_nb[l] > 0) {
            { ProfScope ps_(c, PS_SOLVE_UPDATE); k_solve_update<<<dim3(nfl, (maxns + TILE - 1) / TILE, gz), 256, 0, st>>>(first, 0, s, 2, fr.ptr, d.L.ptr, d.U.ptr, W); }
            c->count_launch();
        }
        int last_kb = ((maxns - 1) / NB) * NB;
        for (int kb = last_kb; kb >= 0; kb -= NB) {
            { ProfScope ps_(c, PS_SOLVE_DIAG); k_solve_diag<<<dim3(nfl, (s + 127) / 128), 128, 0, st>>>(first, kb, s, 1, fr.ptr, d.L.ptr, W); }
            c->count_launch();
            if (kb > 0) {
                { ProfScope ps_(c, PS_SOLVE_UPDATE); k_solve_update<<<dim3(nfl, (kb + TILE - 1) / TILE, gz), 256, 0, st>>>(first, kb, s, 1, fr.ptr, d.L.ptr, d.U.ptr, W); }
                c->count_launch();
            }
        }
        unsigned gys = (unsigned)std::min<int64_t>(std::max<int64_t>(((int64_t)maxns * s + 255) / 256, 1), 1024);
        { ProfScope ps_(c, PS_SOLVE_GATHER); k_solve_store<<<dim3(nfl, gys), 256, 0, st>>>(first, s, fr.ptr, d.fidx.ptr, W, X); }
        c->count_launch();
    }
    PN_CUDA(cudaStreamSynchronize(st));
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

void direct_release(Ctx *c) {
    delete c->direct;
    c->direct = nullptr;
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
    DevBuf<double> R, AX;
    if (!percol) {
        { PhaseTimer tm(c, &c->stats.ms_factor); factorise(c, vals11); }
        PhaseTimer tm(c, &c->stats.ms_solve);
        PN_CUDA(cudaMemcpyAsync(c->X.ptr, c->B.ptr, sizeof(double) * n1 * s, cudaMemcpyDeviceToDevice, st));
        solve_inplace(c, c->X.ptr, s);
        if (refine > 0) { R.resize(n1 * s); AX.resize(n1 * s); }
        for (int it = 0; it < refine; ++it) {
            spmm(c, c->blk[0], vals11, c->X.ptr, AX.ptr, s, st);
            c->stats.spmm_launches += 1;
            { ProfScope ps_(c, PS_SOLVE_MISC); k_residual<<<grid_for(n1 * s, 256), 256, 0, st>>>(n1 * s, c->B.ptr, AX.ptr, R.ptr); }
            solve_inplace(c, R.ptr, s);
            { ProfScope ps_(c, PS_SOLVE_MISC); k_axpy1<<<grid_for(n1 * s, 256), 256, 0, st>>>(n1 * s, R.ptr, c->X.ptr); }
            c->count_launch(2);
            c->stats.iterations += 1;
        }
    } else {
        // one matrix per column (P-Delta): factorise and solve column by column
        d.col_in.resize(n1); d.col_out.resize(n1);
        DevBuf<double> ax;
        ax.resize(n1);
        const CsrBlock &A = c->blk[0];
        for (int j = 0; j < s; ++j) {
            const double *vj = vals11 + (int64_t)j * A.nnz;
            { PhaseTimer tm(c, &c->stats.ms_factor); factorise(c, vj); }
            PhaseTimer tm(c, &c->stats.ms_solve);
            k_get_col<<<grid_for(n1, 256), 256, 0, st>>>(n1, s, j, c->B.ptr, d.col_in.ptr);
            PN_CUDA(cudaMemcpyAsync(d.col_out.ptr, d.col_in.ptr, sizeof(double) * n1, cudaMemcpyDeviceToDevice, st));
            solve_inplace(c, d.col_out.ptr, 1);
            for (int it = 0; it < refine; ++it) {
                spmm(c, A, vj, d.col_out.ptr, ax.ptr, 1, st);
                { ProfScope ps_(c, PS_SOLVE_MISC); k_residual<<<grid_for(n1, 256), 256, 0, st>>>(n1, d.col_in.ptr, ax.ptr, ax.ptr); }
                solve_inplace(c, ax.ptr, 1);
                { ProfScope ps_(c, PS_SOLVE_MISC); k_axpy1<<<grid_for(n1, 256), 256, 0, st>>>(n1, ax.ptr, d.col_out.ptr); }
                c->count_launch(2);
            }
            k_set_col<<<grid_for(n1, 256), 256, 0, st>>>(n1, s, j, d.col_out.ptr, c->X.ptr);
            c->count_launch(2);
        }
    }
}

}  // namespace pn
