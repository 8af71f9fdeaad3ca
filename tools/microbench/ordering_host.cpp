// Host timing driver for the ordering phase (vertex graph + nested dissection + front index lists) on the
// BASELINE config 3 mesh (501 x 501 nodes, perimeter pinned).  No GPU needed.
// build: g++ -O2 -fopenmp -I../../pynite_b200/csrc ordering_host.cpp ../../pynite_b200/csrc/pn_direct_symbolic.cpp -o ordering_host
#include "pn_direct_symbolic.h"
#include <chrono>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
using namespace pn;
int main(int argc, char **argv) {
    int n = argc > 1 ? atoi(argv[1]) : 500;
    int64_t nn = (int64_t)(n + 1) * (n + 1);
    std::vector<double> xyz(3 * nn);
    std::vector<int32_t> newidx(6 * nn), quads(4 * (int64_t)n * n);
    int32_t next = 0;
    for (int r = 0; r <= n; ++r)
        for (int c = 0; c <= n; ++c) {
            int64_t nd = (int64_t)r * (n + 1) + c;
            xyz[3 * nd] = 12.0 * c; xyz[3 * nd + 1] = 12.0 * r; xyz[3 * nd + 2] = 0;
            bool edge = r == 0 || c == 0 || r == n || c == n;
            for (int k = 0; k < 6; ++k) newidx[6 * nd + k] = (edge && k < 3) ? -1 : next++;
        }
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
            int64_t e = (int64_t)r * n + c, i = (int64_t)r * (n + 1) + c;
            quads[4 * e] = i; quads[4 * e + 1] = i + 1; quads[4 * e + 2] = i + n + 2; quads[4 * e + 3] = i + n + 1;
        }
    std::vector<int32_t> vstart, adj;
    std::vector<double> coords;
    std::vector<int64_t> adj_ptr;
    VertexGraphScratch scratch;
    DirectSym S;
    std::vector<uint8_t> row_class(next);
    const bool use_classes = argc > 2 && atoi(argv[2]) != 0;      // second argument 1: DOF-type classes (DofClasses)
    const int leaf = argc > 3 ? atoi(argv[3]) : 24;
    for (int rep = 0; rep < 5; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<ConnView> conn = {{(int64_t)n * n, 4, quads.data(), nullptr}};
        DofClasses cls;
        if (use_classes) cls = predict_dof_classes(xyz.data(), conn, 0);
        build_vertex_graph(nn, newidx.data(), xyz.data(), conn, vstart, coords, adj_ptr, adj, &scratch);
        auto t1 = std::chrono::steady_clock::now();
        if (cls.n > 1)
            for (int64_t dd = 0; dd < 6 * nn; ++dd) if (newidx[dd] >= 0) row_class[newidx[dd]] = cls.of_type[dd % 6];
        direct_symbolic(next, (int32_t)vstart.size() - 1, vstart.data(), coords.data(), adj_ptr.data(), adj.data(), leaf, S,
                        cls.n > 1 ? row_class.data() : nullptr, cls.n, cls.diagonal);
        auto t2 = std::chrono::steady_clock::now();
        printf("classes %d, vertex graph %.1f ms, direct_symbolic %.1f ms, fronts %zu, levels %zu, factor entries %lld, flops %.3e, max children %d\n",
               cls.n, std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count(),
               S.fronts.size(), S.level_ptr.size() - 1, (long long)S.factor_entries, S.factor_flops, S.max_children);
    }
    for (size_t l = 0; l + 1 < S.level_ptr.size(); ++l) {
        int mx_ns = 0, mx_n = 0; double fl = 0;
        for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
            const FrontSym &f = S.fronts[S.level_fronts[q]];
            mx_ns = std::max(mx_ns, f.ns); mx_n = std::max(mx_n, f.ns + f.nb);
            double ns = f.ns, nb = f.nb;
            fl += 2.0 / 3.0 * ns * ns * ns + 2.0 * ns * ns * nb + 2.0 * ns * nb * nb;
        }
        printf("level %2zu: %6d fronts, max ns %5d, max n %5d, %8.2f GFLOP (LU)\n", l, S.level_ptr[l + 1] - S.level_ptr[l], mx_ns, mx_n, fl * 1e-9);
    }
    return 0;
}
