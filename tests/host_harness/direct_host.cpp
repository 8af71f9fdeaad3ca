// TEST INFRASTRUCTURE ONLY.  A sequential host emulation of the multifrontal LU that
// pynite_b200/csrc/pn_direct.cu runs on the GPU, driven by the same symbolic analysis
// (pn_direct_symbolic.cpp).  It validates the ordering, front index lists and extend-add maps
// against scipy on a machine without a GPU.  Nothing in the product package loads this.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <cstring>
#include <vector>

#include "../../pynite_b200/csrc/pn_direct_symbolic.h"

using namespace pn;

extern "C" {

// returns 0 on success; stats[0] = factor entries, stats[1] = flops, stats[2] = #fronts, stats[3] = max front,
// stats[4] = #levels
int hh_direct_solve(int64_t n, const int32_t *indptr, const int32_t *indices, const double *vals, int32_t nv,
                    const int32_t *vstart, const double *coords, const int64_t *adj_ptr, const int32_t *adj, int leaf,
                    int s, const double *B, double *X, double *stats, const uint8_t *row_class, int n_classes,
                    const uint8_t *class_diagonal) {
    DirectSym S;
    try {
        direct_symbolic(n, nv, vstart, coords, adj_ptr, adj, leaf, S, row_class, n_classes, class_diagonal);
    } catch (std::exception &e) {
        fprintf(stderr, "symbolic failed: %s\n", e.what());
        return -1;
    }
    int nf = (int)S.fronts.size();
    stats[0] = (double)S.factor_entries; stats[1] = S.factor_flops; stats[2] = nf; stats[3] = (double)S.max_front;
    stats[4] = (double)S.level_ptr.size() - 1;
    std::vector<std::vector<double>> F(nf);
    for (int f = 0; f < nf; ++f) { int m = S.fronts[f].ns + S.fronts[f].nb; F[f].assign((size_t)m * m, 0.0); }
    // assemble: entry (i, j) belongs to the front of the earlier-eliminated index
    for (int64_t i = 0; i < n; ++i)
        for (int32_t k = indptr[i]; k < indptr[i + 1]; ++k) {
            int32_t j = indices[k];
            int32_t pi = S.perm_pos[i], pj = S.perm_pos[j];
            int f = pi < pj ? S.dof_front[i] : S.dof_front[j];
            const FrontSym &fs = S.fronts[f];
            int m = fs.ns + fs.nb;
            const int32_t *pl = S.fpos.data() + fs.idx_off;
            int li = (int)(std::lower_bound(pl, pl + m, pi) - pl), lj = (int)(std::lower_bound(pl, pl + m, pj) - pl);
            if (li >= m || pl[li] != pi || lj >= m || pl[lj] != pj) { fprintf(stderr, "entry outside front\n"); return -2; }
            F[f][(size_t)li * m + lj] += vals[k];
        }
    // numeric: level by level (children strictly one level below the parent)
    int nl = (int)S.level_ptr.size() - 1;
    for (int l = 0; l < nl; ++l) {
        // extend-add children of the fronts in this level (children live in level l-1), rank by rank
        if (l > 0)
            for (int r = 0; r < S.max_children; ++r)
                for (int q = S.level_ptr[l - 1]; q < S.level_ptr[l]; ++q) {
                    int cidx = S.level_fronts[q];
                    const FrontSym &c = S.fronts[cidx];
                    if (c.parent < 0 || c.child_rank != r) continue;
                    const FrontSym &p = S.fronts[c.parent];
                    int mc = c.ns + c.nb, mp = p.ns + p.nb;
                    const int32_t *rel = S.rel.data() + c.rel_off;
                    for (int a = 0; a < c.nb; ++a)
                        for (int b = 0; b < c.nb; ++b)
                            F[c.parent][(size_t)rel[a] * mp + rel[b]] += F[cidx][(size_t)(c.ns + a) * mc + c.ns + b];
                }
        for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
            int f = S.level_fronts[q];
            const FrontSym &fs = S.fronts[f];
            int m = fs.ns + fs.nb;
            double *A = F[f].data();
            for (int p = 0; p < fs.ns; ++p) {
                double piv = A[(size_t)p * m + p];
                if (piv == 0.0 || piv != piv) return -3;
                for (int i = p + 1; i < m; ++i) {
                    double lf = A[(size_t)i * m + p] / piv;
                    A[(size_t)i * m + p] = lf;
                    if (lf != 0.0) for (int j = p + 1; j < m; ++j) A[(size_t)i * m + j] -= lf * A[(size_t)p * m + j];
                }
            }
        }
    }
    // forward: bottom-up with update vectors
    std::vector<std::vector<double>> W(nf);
    std::vector<double> Y((size_t)n * s);
    memcpy(Y.data(), B, sizeof(double) * n * s);
    for (int l = 0; l < nl; ++l) {
        for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
            int f = S.level_fronts[q];
            const FrontSym &fs = S.fronts[f];
            int m = fs.ns + fs.nb;
            W[f].assign((size_t)m * s, 0.0);
            for (int a = 0; a < fs.ns; ++a)
                for (int j = 0; j < s; ++j) W[f][(size_t)a * s + j] = Y[(size_t)S.fidx[fs.idx_off + a] * s + j];
        }
        if (l > 0)
            for (int r = 0; r < S.max_children; ++r)
                for (int q = S.level_ptr[l - 1]; q < S.level_ptr[l]; ++q) {
                    int cidx = S.level_fronts[q];
                    const FrontSym &c = S.fronts[cidx];
                    if (c.parent < 0 || c.child_rank != r) continue;
                    const int32_t *rel = S.rel.data() + c.rel_off;
                    for (int a = 0; a < c.nb; ++a)
                        for (int j = 0; j < s; ++j) W[c.parent][(size_t)rel[a] * s + j] += W[cidx][(size_t)(c.ns + a) * s + j];
                }
        for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
            int f = S.level_fronts[q];
            const FrontSym &fs = S.fronts[f];
            int m = fs.ns + fs.nb;
            const double *A = F[f].data();
            double *w = W[f].data();
            for (int p = 0; p < fs.ns; ++p)
                for (int i = p + 1; i < m; ++i) {
                    double lf = A[(size_t)i * m + p];
                    if (lf != 0.0) for (int j = 0; j < s; ++j) w[(size_t)i * s + j] -= lf * w[(size_t)p * s + j];
                }
            for (int a = 0; a < fs.ns; ++a)
                for (int j = 0; j < s; ++j) Y[(size_t)S.fidx[fs.idx_off + a] * s + j] = w[(size_t)a * s + j];
        }
    }
    // backward: top-down
    for (int l = nl - 1; l >= 0; --l)
        for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
            int f = S.level_fronts[q];
            const FrontSym &fs = S.fronts[f];
            int m = fs.ns + fs.nb;
            const double *A = F[f].data();
            std::vector<double> w((size_t)m * s);
            for (int a = 0; a < m; ++a)
                for (int j = 0; j < s; ++j) w[(size_t)a * s + j] = Y[(size_t)S.fidx[fs.idx_off + a] * s + j];
            for (int p = fs.ns - 1; p >= 0; --p)
                for (int j = 0; j < s; ++j) {
                    double acc = w[(size_t)p * s + j];
                    for (int c = p + 1; c < m; ++c) acc -= A[(size_t)p * m + c] * w[(size_t)c * s + j];
                    w[(size_t)p * s + j] = acc / A[(size_t)p * m + p];
                }
            for (int a = 0; a < fs.ns; ++a)
                for (int j = 0; j < s; ++j) Y[(size_t)S.fidx[fs.idx_off + a] * s + j] = w[(size_t)a * s + j];
        }
    memcpy(X, Y.data(), sizeof(double) * n * s);
    // every child exactly one level below its parent, one root, subtrees contiguous in front order
    {
        std::vector<int> level_of(nf, -1);
        for (int l = 0; l < nl; ++l)
            for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) level_of[S.level_fronts[q]] = l;
        int roots = 0;
        for (int f = 0; f < nf; ++f) {
            const FrontSym &fs = S.fronts[f];
            if (fs.parent < 0) { ++roots; continue; }
            if (fs.parent <= f || level_of[fs.parent] != level_of[f] + 1) return -4;
        }
        stats[5] = roots;
        stats[6] = S.max_children;
    }
    return 0;
}

// The DOF-type classes predicted from the shell lists (the library's own predict_dof_classes).
int hh_predict_classes(const double *xyz, int64_t n_quads, const int32_t *quads, int64_t n_plates, const int32_t *plates,
                       int64_t n_line, uint8_t *of_type, uint8_t *diagonal) {
    std::vector<ConnView> shells = {{n_quads, 4, quads, nullptr}, {n_plates, 4, plates, nullptr}};
    DofClasses cls = predict_dof_classes(xyz, shells, n_line);
    for (int k = 0; k < 6; ++k) { of_type[k] = cls.of_type[k]; diagonal[k] = cls.diagonal[k]; }
    return cls.n;
}
}
