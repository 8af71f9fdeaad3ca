// Node-block-Jacobi preconditioned conjugate gradients for K11 X = B, all right-hand sides at once.
// Vectors are row-major [n1][s]; every column runs its own CG recurrence (own alpha, beta), the
// matrix is streamed once per iteration for all columns (SpMM).  Reductions use a fixed grid and a
// fixed summation tree, so results are bit-reproducible run to run.
#include "pn_internal.h"

namespace pn {

static inline unsigned grid_for(int64_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

struct PcgWork {
    DevBuf<double> R, Z, P, AP;
    DevBuf<double> minv;          // [n_nodes][36] inverse diagonal blocks (row-major nb x nb, ld 6)
    DevBuf<int32_t> nb_start;     // [n_nodes] first K11 row of the node's free DOFs
    DevBuf<uint8_t> nb_size;      // [n_nodes] number of free DOFs (0..6)
    DevBuf<double> partial;       // [grid][3][s]
    DevBuf<double> scal;          // rz, rz_new, pAp, rr, bb : [5][s]
    int red_grid = 0;
};

constexpr int RED_TPB = 256;

// ---- preconditioner setup ---------------------------------------------------------------------
__global__ void k_node_blocks(int64_t n_nodes, const int32_t *__restrict__ newidx, int32_t *__restrict__ start,
                              uint8_t *__restrict__ size) {
    int64_t nd = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (nd >= n_nodes) return;
    int32_t first = -1; int cnt = 0;
    for (int k = 0; k < 6; ++k) {
        int32_t ni = newidx[6 * nd + k];
        if (ni >= 0) { if (first < 0) first = ni; ++cnt; }
    }
    start[nd] = first; size[nd] = (uint8_t)cnt;
}

__global__ void k_block_inverse(int64_t n_nodes, const int32_t *__restrict__ start, const uint8_t *__restrict__ size,
                                const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                const double *__restrict__ vals, double *__restrict__ minv) {
    int64_t nd = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (nd >= n_nodes) return;
    int nb = size[nd];
    if (!nb) return;
    int32_t i0 = start[nd];
    double a[36], inv[36];
    for (int i = 0; i < 36; ++i) { a[i] = 0.0; inv[i] = 0.0; }
    for (int i = 0; i < nb; ++i) {
        inv[i * 6 + i] = 1.0;
        for (int32_t k = indptr[i0 + i]; k < indptr[i0 + i + 1]; ++k) {
            int32_t cidx = indices[k] - i0;
            if (cidx >= 0 && cidx < nb) a[i * 6 + cidx] = vals[k];
        }
    }
    // symmetrise (K is symmetric up to rounding) so the preconditioner is exactly symmetric
    for (int i = 0; i < nb; ++i)
        for (int j = i + 1; j < nb; ++j) { double m = 0.5 * (a[i * 6 + j] + a[j * 6 + i]); a[i * 6 + j] = m; a[j * 6 + i] = m; }
    // Gauss-Jordan with partial pivoting
    for (int col = 0; col < nb; ++col) {
        int piv = col; double best = fabs(a[col * 6 + col]);
        for (int r = col + 1; r < nb; ++r) if (fabs(a[r * 6 + col]) > best) { best = fabs(a[r * 6 + col]); piv = r; }
        if (best == 0.0) { a[col * 6 + col] = 1.0; piv = col; }    // singular block: fall back to identity on this DOF
        if (piv != col)
            for (int j = 0; j < nb; ++j) {
                double t = a[col * 6 + j]; a[col * 6 + j] = a[piv * 6 + j]; a[piv * 6 + j] = t;
                t = inv[col * 6 + j]; inv[col * 6 + j] = inv[piv * 6 + j]; inv[piv * 6 + j] = t;
            }
        double d = 1.0 / a[col * 6 + col];
        for (int j = 0; j < nb; ++j) { a[col * 6 + j] *= d; inv[col * 6 + j] *= d; }
        for (int r = 0; r < nb; ++r) {
            if (r == col) continue;
            double f = a[r * 6 + col];
            if (f == 0.0) continue;
            for (int j = 0; j < nb; ++j) { a[r * 6 + j] -= f * a[col * 6 + j]; inv[r * 6 + j] -= f * inv[col * 6 + j]; }
        }
    }
    for (int i = 0; i < 36; ++i) minv[nd * 36 + i] = inv[i];
}

// ---- SpMM with one value array per column (P-Delta: same pattern, one matrix per combo) ----------
__global__ void k_spmm_percol(int64_t rows, int s, int64_t nnz, const int32_t *__restrict__ indptr,
                              const int32_t *__restrict__ indices, const double *__restrict__ vals,
                              const double *__restrict__ X, double *__restrict__ Y) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= rows * s) return;
    int64_t r = t / s; int j = (int)(t - r * s);
    const double *v = vals + (int64_t)j * nnz;
    double acc = 0.0;
    for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) acc = fma(v[k], X[(int64_t)indices[k] * s + j], acc);
    Y[t] = acc;
}

static void apply_matrix(Ctx *c, const double *vals11, int percol, const double *X, double *Y, int s) {
    const CsrBlock &A = c->blk[0];
    if (percol) {
        { ProfScope ps_(c, PS_SPMM); k_spmm_percol<<<grid_for(A.rows * s, 256), 256, 0, c->stream>>>(A.rows, s, A.nnz, A.indptr.ptr, A.indices.ptr,
                                                                         vals11, X, Y); }
        c->count_launch();
    } else if (c->opt_spmm_grouped) {
        spmm_k11_grouped(c, vals11, X, Y, s, c->stream);
    } else {
        spmm(c, A, vals11, X, Y, s, c->stream);
    }
    c->stats.spmm_launches += 1;
}

// ---- column-wise reductions ---------------------------------------------------------------------
// partial[block][which][j] = sum over this block's rows of A[i][j] * B[i][j]
__global__ void __launch_bounds__(RED_TPB)
k_dot_partial(int64_t n, int s, const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ partial,
              int which, int n_which) {
    __shared__ double sm[RED_TPB];
    int lanes = RED_TPB / s;             // row lanes per block (s <= 256)
    int t = threadIdx.x;
    int j = t % s, rl = t / s;
    double acc = 0.0;
    if (rl < lanes)
        for (int64_t i = (int64_t)blockIdx.x * lanes + rl; i < n; i += (int64_t)gridDim.x * lanes)
            acc = fma(A[i * s + j], B[i * s + j], acc);
    sm[t] = acc;
    __syncthreads();
    if (t < s) {
        double tot = 0.0;
        for (int l = 0; l < lanes; ++l) tot += sm[l * s + t];
        partial[((int64_t)blockIdx.x * n_which + which) * s + t] = tot;
    }
}
__global__ void k_dot_final(int grid, int s, int n_which, const double *__restrict__ partial, double *__restrict__ out0,
                            double *__restrict__ out1, double *__restrict__ out2) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= s * n_which) return;
    int which = t / s, j = t % s;
    double tot = 0.0;
    for (int b = 0; b < grid; ++b) tot += partial[((int64_t)b * n_which + which) * s + j];
    double *o = which == 0 ? out0 : (which == 1 ? out1 : out2);
    o[j] = tot;
}

// ---- fused: x += alpha p, r -= alpha Ap, z = Minv r, partial dots r.z and r.r ----------------------
__global__ void __launch_bounds__(RED_TPB)
k_update_xrz(int64_t node0, int64_t n_nodes, int s, const int32_t *__restrict__ start, const uint8_t *__restrict__ size,
             const double *__restrict__ minv, const double *__restrict__ rz, const double *__restrict__ pAp,
             const double *__restrict__ P, const double *__restrict__ AP, double *__restrict__ X, double *__restrict__ R,
             double *__restrict__ Z, double *__restrict__ partial, int first) {
    __shared__ double sm0[RED_TPB], sm1[RED_TPB];
    int lanes = RED_TPB / s;
    int t = threadIdx.x;
    int j = t % s, nl = t / s;
    double a_rz = 0.0, a_rr = 0.0;
    if (nl < lanes) {
        double alpha = 0.0;
        if (!first) { double d = pAp[j]; alpha = d != 0.0 ? rz[j] / d : 0.0; }
        for (int64_t nd = node0 + (int64_t)blockIdx.x * lanes + nl; nd < n_nodes; nd += (int64_t)gridDim.x * lanes) {
            int nb = size[nd];
            if (!nb) continue;
            int64_t i0 = start[nd];
            double r[6];
            for (int k = 0; k < nb; ++k) {
                int64_t idx = (i0 + k) * s + j;
                double rv = R[idx];
                if (!first) {
                    X[idx] = fma(alpha, P[idx], X[idx]);
                    rv = fma(-alpha, AP[idx], rv);
                    R[idx] = rv;
                }
                r[k] = rv;
            }
            const double *m = minv + nd * 36;
            for (int k = 0; k < nb; ++k) {
                double z = 0.0;
                for (int l = 0; l < nb; ++l) z = fma(m[k * 6 + l], r[l], z);
                Z[(i0 + k) * s + j] = z;
                a_rz = fma(r[k], z, a_rz);
                a_rr = fma(r[k], r[k], a_rr);
            }
        }
    }
    sm0[t] = a_rz; sm1[t] = a_rr;
    __syncthreads();
    if (t < s) {
        double t0 = 0.0, t1 = 0.0;
        for (int l = 0; l < lanes; ++l) { t0 += sm0[l * s + t]; t1 += sm1[l * s + t]; }
        partial[((int64_t)blockIdx.x * 2 + 0) * s + t] = t0;
        partial[((int64_t)blockIdx.x * 2 + 1) * s + t] = t1;
    }
}

// p = z + beta p with beta = rz_new / rz; then rz <- rz_new (done by thread row 0)
__global__ void k_update_p(int64_t n, int s, const double *__restrict__ rz, const double *__restrict__ rz_new,
                           const double *__restrict__ Z, double *__restrict__ P, int first) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * s) return;
    int j = (int)(t % s);
    double beta = 0.0;
    if (!first) { double d = rz[j]; beta = d != 0.0 ? rz_new[j] / d : 0.0; }
    P[t] = fma(beta, P[t], Z[t]);
}
__global__ void k_copy_s(int s, const double *__restrict__ a, double *__restrict__ b) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < s) b[j] = a[j];
}

// out_host[j] = sum_i A[i][j] * B[i][j] for row-major [n][s] arrays (deterministic reduction tree)
// Up to three column-wise dot products in ONE pass and one host synchronisation: out_host[w][j] = sum_i A_w[i][j] B_w[i][j].
// (The refinement needs |d|^2 and |x|^2, the residual check |r|^2, |b|^2 and |x|^2: five launches of k_dot_partial, five
// serial 592-term sums in k_dot_final and five stream synchronisations became two of each.)  Block partials as in
// k_dot_partial; the final sum of a (product, column) pair is done by one warp in a fixed order.
struct DotPairs { const double *A[3], *B[3]; };
__global__ void __launch_bounds__(RED_TPB)
k_dot_partial_multi(int64_t n, int s, DotPairs p, int n_which, double *__restrict__ partial) {
    __shared__ double sm[3][RED_TPB];
    const int lanes = RED_TPB / s;
    const int t = threadIdx.x, j = t % s, rl = t / s;
    double acc[3] = {0.0, 0.0, 0.0};
    if (rl < lanes)
        for (int64_t i = (int64_t)blockIdx.x * lanes + rl; i < n; i += (int64_t)gridDim.x * lanes)
#pragma unroll
            for (int w = 0; w < 3; ++w)
                if (w < n_which) acc[w] = fma(p.A[w][i * s + j], p.B[w][i * s + j], acc[w]);
#pragma unroll
    for (int w = 0; w < 3; ++w) sm[w][t] = acc[w];
    __syncthreads();
    if (t < s)
        for (int w = 0; w < n_which; ++w) {
            double tot = 0.0;
            for (int l = 0; l < lanes; ++l) tot += sm[w][l * s + t];
            partial[((int64_t)blockIdx.x * n_which + w) * s + t] = tot;
        }
}
// one warp per (product, column): lane l adds the partials of blocks l, l + 32, ... in order, then a shuffle tree
__global__ void k_dot_final_warp(int grid, int s, int n_which, const double *__restrict__ partial, double *__restrict__ out) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= s * n_which) return;
    const int w = wid / s, j = wid % s;
    double tot = 0.0;
    for (int b = lane; b < grid; b += 32) tot += partial[((int64_t)b * n_which + w) * s + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xffffffffu, tot, o);
    if (lane == 0) out[w * s + j] = tot;
}
void column_dots_multi(Ctx *c, int64_t n, int s, int n_which, const double *const *A, const double *const *B, double *out_host) {
    if (s > 256) throw ArgError("column_dots: at most 256 columns");
    if (n_which < 1 || n_which > 3) throw ArgError("column_dots_multi: one to three products");
    const int grid = 148 * 4;
    DevBuf<double> &partial = c->dots_partial, &out = c->dots_out;
    partial.resize((int64_t)grid * s * n_which);
    out.resize((int64_t)s * n_which);
    DotPairs p{};
    for (int w = 0; w < 3; ++w) { p.A[w] = A[w < n_which ? w : 0]; p.B[w] = B[w < n_which ? w : 0]; }
    { ProfScope ps_(c, PS_PCG_VECTOR); k_dot_partial_multi<<<grid, RED_TPB, 0, c->stream>>>(n, s, p, n_which, partial.ptr); }
    { ProfScope ps_(c, PS_PCG_VECTOR); k_dot_final_warp<<<grid_for((int64_t)s * n_which * 32, 128), 128, 0, c->stream>>>(grid, s, n_which, partial.ptr, out.ptr); }
    c->count_launch(2);
    PN_CUDA(cudaMemcpyAsync(out_host, out.ptr, sizeof(double) * s * n_which, cudaMemcpyDeviceToHost, c->stream));
    PN_CUDA(cudaStreamSynchronize(c->stream));
}

void column_dots(Ctx *c, int64_t n, int s, const double *A, const double *B, double *out_host) {
    if (s > 256) throw ArgError("column_dots: at most 256 columns");
    const int grid = 148 * 4;
    DevBuf<double> &partial = c->dots_partial, &out = c->dots_out;
    partial.resize((int64_t)grid * s);
    out.resize(s);
    { ProfScope ps_(c, PS_PCG_VECTOR); k_dot_partial<<<grid, RED_TPB, 0, c->stream>>>(n, s, A, B, partial.ptr, 0, 1); }
    { ProfScope ps_(c, PS_PCG_VECTOR); k_dot_final<<<grid_for(s, 128), 128, 0, c->stream>>>(grid, s, 1, partial.ptr, out.ptr, nullptr, nullptr); }
    c->count_launch(2);
    PN_CUDA(cudaMemcpyAsync(out_host, out.ptr, sizeof(double) * s, cudaMemcpyDeviceToHost, c->stream));
    PN_CUDA(cudaStreamSynchronize(c->stream));
}

// Z[rows of node] = Minv_node R[rows of node] for nodes [node0, node1); thread = (node, column)
__global__ void k_apply_minv(int64_t node0, int64_t n_range, int s, const int32_t *__restrict__ start, const uint8_t *__restrict__ size,
                             const double *__restrict__ minv, const double *__restrict__ R, double *__restrict__ Z) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_range * s) return;
    int64_t nd = node0 + t / s;
    int j = (int)(t % s);
    int nb = size[nd];
    if (!nb) return;
    int64_t i0 = start[nd];
    double r[6];
    for (int k = 0; k < nb; ++k) r[k] = R[(i0 + k) * s + j];
    const double *m = minv + nd * 36;
    for (int k = 0; k < nb; ++k) {
        double z = 0.0;
        for (int l = 0; l < nb; ++l) z = fma(m[k * 6 + l], r[l], z);
        Z[(i0 + k) * s + j] = z;
    }
}

void pcg_setup_precond(Ctx *c, const double *vals11) {
    if (!c->pcg) c->pcg = new PcgWork();
    PcgWork &w = *c->pcg;
    cudaStream_t st = c->stream;
    const CsrBlock &A = c->blk[0];
    int64_t nn = c->model.n_nodes;
    w.nb_start.resize(nn); w.nb_size.resize(nn); w.minv.resize(nn * 36);
    k_node_blocks<<<grid_for(nn, 256), 256, 0, st>>>(nn, c->newidx.ptr, w.nb_start.ptr, w.nb_size.ptr);
    k_block_inverse<<<grid_for(nn, 128), 128, 0, st>>>(nn, w.nb_start.ptr, w.nb_size.ptr, A.indptr.ptr, A.indices.ptr, vals11, w.minv.ptr);
    c->count_launch(2);
}

void pcg_apply_precond_nodes(Ctx *c, int s, int64_t node0, int64_t node1, const double *R, double *Z) {
    if (!c->pcg || !c->pcg->minv.n) throw ArgError("block-Jacobi preconditioner has not been set up");
    if (node1 <= node0) return;
    PcgWork &w = *c->pcg;
    ProfScope ps_(c, PS_PCG_VECTOR);
    k_apply_minv<<<grid_for((node1 - node0) * s, 256), 256, 0, c->stream>>>(node0, node1 - node0, s, w.nb_start.ptr, w.nb_size.ptr,
                                                                              w.minv.ptr, R, Z);
    c->count_launch();
}

// ---- fused CG vector kernels on a strip (domain-decomposed PCG): the scalars live in device memory so that the caller
// can all-reduce them between the kernels without a host round trip
void pcg_strip_dots(Ctx *c, int s, int64_t row0, int64_t row1, const double *A, const double *B, double *out_dev) {
    PcgWork &w = *c->pcg;
    if (s > 256) throw ArgError("pcg: at most 256 right-hand sides per call");
    w.red_grid = 148 * 4;
    w.partial.resize((int64_t)w.red_grid * 3 * s);
    ProfScope ps_(c, PS_PCG_VECTOR);
    k_dot_partial<<<w.red_grid, RED_TPB, 0, c->stream>>>(row1 - row0, s, A + row0 * s, B + row0 * s, w.partial.ptr, 0, 1);
    k_dot_final<<<grid_for(s, 128), 128, 0, c->stream>>>(w.red_grid, s, 1, w.partial.ptr, out_dev, nullptr, nullptr);
    c->count_launch(2);
}
// x += alpha p, r -= alpha Ap (alpha = rz / pAp per column; skipped when first), z = Minv r on nodes [node0, node1);
// out2_dev[0:s] = local r.z, out2_dev[s:2s] = local r.r
void pcg_strip_update(Ctx *c, int s, int64_t node0, int64_t node1, int first, const double *rz_dev, const double *pAp_dev,
                      const double *P, const double *AP, double *X, double *R, double *Z, double *out2_dev) {
    PcgWork &w = *c->pcg;
    if (s > 256) throw ArgError("pcg: at most 256 right-hand sides per call");
    w.red_grid = 148 * 4;
    w.partial.resize((int64_t)w.red_grid * 3 * s);
    ProfScope ps_(c, PS_PCG_VECTOR);
    k_update_xrz<<<w.red_grid, RED_TPB, 0, c->stream>>>(node0, node1, s, w.nb_start.ptr, w.nb_size.ptr, w.minv.ptr, rz_dev, pAp_dev, P,
                                                         AP, X, R, Z, w.partial.ptr, first);
    k_dot_final<<<grid_for(2 * s, 128), 128, 0, c->stream>>>(w.red_grid, s, 2, w.partial.ptr, out2_dev, out2_dev + s, nullptr);
    c->count_launch(2);
}
// p = z + (rz_new / rz) p on rows [row0, row1)   (p = z when first)
void pcg_strip_direction(Ctx *c, int s, int64_t row0, int64_t row1, int first, const double *rz_dev, const double *rz_new_dev,
                         const double *Z, double *P) {
    if (row1 <= row0) return;
    ProfScope ps_(c, PS_PCG_VECTOR);
    k_update_p<<<grid_for((row1 - row0) * s, 256), 256, 0, c->stream>>>(row1 - row0, s, rz_dev, rz_new_dev, Z + row0 * s, P + row0 * s, first);
    c->count_launch();
}

// host array [n_nodes + 1]: first K11 row of every node (nodes without free DOFs repeat the next start)
void pcg_node_row_starts(Ctx *c, int64_t *out_host) {
    int64_t nn = c->model.n_nodes;
    std::vector<int32_t> newidx(c->n_dof);
    PN_CUDA(cudaMemcpyAsync(newidx.data(), c->newidx.ptr, sizeof(int32_t) * c->n_dof, cudaMemcpyDeviceToHost, c->stream));
    PN_CUDA(cudaStreamSynchronize(c->stream));
    int64_t next = c->n1;
    out_host[nn] = c->n1;
    for (int64_t nd = nn - 1; nd >= 0; --nd) {
        for (int k = 5; k >= 0; --k)
            if (newidx[6 * nd + k] >= 0) next = newidx[6 * nd + k];
        out_host[nd] = next;
    }
}

void pcg_release(Ctx *c) {
    delete c->pcg;
    c->pcg = nullptr;
}

void pcg_solve(Ctx *c, const double *vals11, int percol, int s, const pn_solve_opts &o) {
    if (s > 256) throw ArgError("pcg: at most 256 right-hand sides per call");
    if (!c->model.symmetric)
        throw ArgError("pcg: the stiffness matrix is unsymmetric (a shell has kx_mod != ky_mod); use the direct solver");
    cudaStream_t st = c->stream;
    const CsrBlock &A = c->blk[0];
    int64_t n1 = c->n1;
    if (!c->pcg) c->pcg = new PcgWork();
    PcgWork &w = *c->pcg;
    int64_t nn = c->model.n_nodes;
    {
        PhaseTimer tm(c, &c->stats.ms_factor);
        w.nb_start.resize(nn); w.nb_size.resize(nn); w.minv.resize(nn * 36);
        k_node_blocks<<<grid_for(nn, 256), 256, 0, st>>>(nn, c->newidx.ptr, w.nb_start.ptr, w.nb_size.ptr);
        // P-Delta: precondition every column with the elastic blocks of column 0's matrix
        k_block_inverse<<<grid_for(nn, 128), 128, 0, st>>>(nn, w.nb_start.ptr, w.nb_size.ptr, A.indptr.ptr, A.indices.ptr,
                                                            vals11, w.minv.ptr);
        c->count_launch(2);
    }
    PhaseTimer tm(c, &c->stats.ms_solve);
    w.R.resize(n1 * s); w.Z.resize(n1 * s); w.P.resize(n1 * s); w.AP.resize(n1 * s);
    w.red_grid = 148 * 4;
    w.partial.resize((int64_t)w.red_grid * 3 * s);
    w.scal.resize(5 * (int64_t)s);
    double *rz = w.scal.ptr, *rz_new = rz + s, *pAp = rz + 2 * s, *rr = rz + 3 * s, *bb = rz + 4 * s;
    double tol = o.tol > 0 ? o.tol : 1e-12;
    int max_iter = o.max_iter > 0 ? o.max_iter : 200000;

    c->X.zero(st);
    PN_CUDA(cudaMemcpyAsync(w.R.ptr, c->B.ptr, (size_t)(n1 * s) * sizeof(double), cudaMemcpyDeviceToDevice, st));
    // z = Minv r, rz = r.z, bb = r.r (x0 = 0)
    { ProfScope ps_(c, PS_PCG_VECTOR); k_update_xrz<<<w.red_grid, RED_TPB, 0, st>>>(0, nn, s, w.nb_start.ptr, w.nb_size.ptr, w.minv.ptr, rz, pAp, w.P.ptr, w.AP.ptr,
                                                  c->X.ptr, w.R.ptr, w.Z.ptr, w.partial.ptr, 1); }
    { ProfScope ps_(c, PS_PCG_VECTOR); k_dot_final<<<grid_for(2 * s, 128), 128, 0, st>>>(w.red_grid, s, 2, w.partial.ptr, rz, bb, nullptr); }
    { ProfScope ps_(c, PS_PCG_VECTOR); k_update_p<<<grid_for(n1 * s, 256), 256, 0, st>>>(n1, s, rz, rz_new, w.Z.ptr, w.P.ptr, 1); }
    c->count_launch(3);

    std::vector<double> h_bb(s), h_rr(s);
    PN_CUDA(cudaMemcpyAsync(h_bb.data(), bb, s * sizeof(double), cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    bool all_zero = true;
    for (int j = 0; j < s; ++j) if (h_bb[j] != 0.0) all_zero = false;

    int64_t it = 0;
    bool converged = all_zero;
    const int check_every = 20;
    while (!converged && it < max_iter) {
        apply_matrix(c, vals11, percol, w.P.ptr, w.AP.ptr, s);
        { ProfScope ps_(c, PS_PCG_VECTOR); k_dot_partial<<<w.red_grid, RED_TPB, 0, st>>>(n1, s, w.P.ptr, w.AP.ptr, w.partial.ptr, 0, 1); }
        { ProfScope ps_(c, PS_PCG_VECTOR); k_dot_final<<<grid_for(s, 128), 128, 0, st>>>(w.red_grid, s, 1, w.partial.ptr, pAp, nullptr, nullptr); }
        { ProfScope ps_(c, PS_PCG_VECTOR); k_update_xrz<<<w.red_grid, RED_TPB, 0, st>>>(0, nn, s, w.nb_start.ptr, w.nb_size.ptr, w.minv.ptr, rz, pAp, w.P.ptr,
                                                      w.AP.ptr, c->X.ptr, w.R.ptr, w.Z.ptr, w.partial.ptr, 0); }
        { ProfScope ps_(c, PS_PCG_VECTOR); k_dot_final<<<grid_for(2 * s, 128), 128, 0, st>>>(w.red_grid, s, 2, w.partial.ptr, rz_new, rr, nullptr); }
        { ProfScope ps_(c, PS_PCG_VECTOR); k_update_p<<<grid_for(n1 * s, 256), 256, 0, st>>>(n1, s, rz, rz_new, w.Z.ptr, w.P.ptr, 0); }
        { ProfScope ps_(c, PS_PCG_VECTOR); k_copy_s<<<1, 256, 0, st>>>(s, rz_new, rz); }
        c->count_launch(6);
        ++it;
        if (it % check_every == 0 || it == max_iter) {
            PN_CUDA(cudaMemcpyAsync(h_rr.data(), rr, s * sizeof(double), cudaMemcpyDeviceToHost, st));
            PN_CUDA(cudaStreamSynchronize(st));
            converged = true;
            for (int j = 0; j < s; ++j) {
                if (!(h_rr[j] <= tol * tol * h_bb[j])) converged = false;      // also false for NaN
                if (h_rr[j] != h_rr[j]) { it = max_iter; break; }
            }
        }
    }
    c->stats.iterations = it;
    if (!converged) throw StatusError(PN_ERR_NOT_CONVERGED, "PCG did not reach the requested tolerance in " +
                                                                std::to_string(it) + " iterations");
}

}  // namespace pn
