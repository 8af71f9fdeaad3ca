// Host-side symbolic analysis for the class-based block-gather assembly.  See pn_assemble_symbolic.h.
#include "pn_assemble_symbolic.h"

#include <algorithm>
#include <cstring>
#include <numeric>

namespace pn {

void classify_elements(int64_t count, int words, const uint64_t *sig, std::vector<int32_t> &cls, std::vector<int32_t> &reps) {
    cls.assign(count, -1);
    reps.clear();
    if (!count) return;
    std::vector<uint64_t> hash(count);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < count; ++e) {
        uint64_t h = 1469598103934665603ull;
        for (int w = 0; w < words; ++w) {
            uint64_t x = sig[e * words + w];
            x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33;
            h = (h ^ x) * 1099511628211ull;
        }
        hash[e] = h;
    }
    std::vector<int64_t> order(count);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return hash[a] != hash[b] ? hash[a] < hash[b] : a < b; });
    // runs of equal hash; inside a run compare exactly against the run's distinct representatives
    std::vector<int64_t> rep_first;      // representative elements in discovery order (by hash run)
    int64_t q = 0;
    while (q < count) {
        int64_t r = q;
        while (r < count && hash[order[r]] == hash[order[q]]) ++r;
        std::vector<int64_t> run_reps;
        for (int64_t t = q; t < r; ++t) {
            int64_t e = order[t];
            int64_t hit = -1;
            for (int64_t cand : run_reps)
                if (!memcmp(sig + cand * words, sig + e * words, sizeof(uint64_t) * words)) { hit = cand; break; }
            if (hit < 0) { run_reps.push_back(e); rep_first.push_back(e); hit = e; }
            cls[e] = (int32_t)hit;             // temporarily: representative element id
        }
        q = r;
    }
    // number classes by their representative's element index so the numbering does not depend on hash order
    std::sort(rep_first.begin(), rep_first.end());
    std::vector<int32_t> id_of(count, -1);
    reps.resize(rep_first.size());
    for (size_t k = 0; k < rep_first.size(); ++k) { id_of[rep_first[k]] = (int32_t)k; reps[k] = (int32_t)rep_first[k]; }
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < count; ++e) cls[e] = id_of[cls[e]];
}

namespace {
struct Inc {                 // one (element, local node) incidence of a node, emission order = (kind, elem)
    uint8_t kind, la;
    int32_t elem;
};
}  // namespace

void build_assembly_sym(int64_t n_nodes, int64_t n_nsp, const int32_t *nsp_dof, const ElemView views[4], const int32_t *indptr,
                        const int32_t *indices, AssemblySym &out) {
    out = AssemblySym();
    // node -> incident elements, emission order
    std::vector<uint32_t> inc_ptr(n_nodes + 1, 0);
    for (int k = 0; k < 4; ++k) {
        const ElemView &v = views[k];
        for (int64_t e = 0; e < v.count; ++e) {
            if (v.active && !v.active[e]) continue;
            for (int a = 0; a < v.nn; ++a) ++inc_ptr[v.conn[v.nn * e + a] + 1];
        }
    }
    for (int64_t a = 0; a < n_nodes; ++a) inc_ptr[a + 1] += inc_ptr[a];
    std::vector<Inc> inc(inc_ptr[n_nodes]);
    {
        std::vector<uint32_t> fill(inc_ptr.begin(), inc_ptr.end() - 1);
        for (int k = 0; k < 4; ++k) {
            const ElemView &v = views[k];
            for (int64_t e = 0; e < v.count; ++e) {
                if (v.active && !v.active[e]) continue;
                for (int a = 0; a < v.nn; ++a) inc[fill[v.conn[v.nn * e + a]]++] = Inc{(uint8_t)k, (uint8_t)a, (int32_t)e};
            }
        }
    }
    // node support springs per node
    std::vector<uint32_t> nsp_ptr(n_nodes + 1, 0);
    for (int64_t q = 0; q < n_nsp; ++q) ++nsp_ptr[nsp_dof[q] / 6 + 1];
    for (int64_t a = 0; a < n_nodes; ++a) nsp_ptr[a + 1] += nsp_ptr[a];      // nsp list is already in node order

    // pass 1: pairs per node
    std::vector<uint32_t> npairs(n_nodes + 1, 0);
#pragma omp parallel for schedule(dynamic, 2048)
    for (int64_t a = 0; a < n_nodes; ++a) {
        int32_t nb[256];
        int cnt = 0;
        bool self = nsp_ptr[a + 1] > nsp_ptr[a];
        for (uint32_t t = inc_ptr[a]; t < inc_ptr[a + 1]; ++t) {
            const ElemView &v = views[inc[t].kind];
            for (int b = 0; b < v.nn; ++b) {
                int32_t nbn = v.conn[v.nn * inc[t].elem + b];
                bool seen = false;
                for (int u = 0; u < cnt; ++u) if (nb[u] == nbn) { seen = true; break; }
                if (!seen && cnt < 256) nb[cnt++] = nbn;
            }
        }
        if (self) {
            bool seen = false;
            for (int u = 0; u < cnt; ++u) if (nb[u] == (int32_t)a) { seen = true; break; }
            if (!seen && cnt < 256) nb[cnt++] = (int32_t)a;
        }
        npairs[a + 1] = (uint32_t)cnt;
    }
    int maxp = 0;
    for (int64_t a = 0; a < n_nodes; ++a) { maxp = std::max<int>(maxp, (int)npairs[a + 1]); }
    out.max_pairs_per_node = maxp;
    if (maxp >= 256 || maxp * 8 > 65535) return;          // out.ok stays false: caller uses the general path
    out.node_pair_ptr.assign(n_nodes + 1, 0);
    for (int64_t a = 0; a < n_nodes; ++a) out.node_pair_ptr[a + 1] = out.node_pair_ptr[a] + npairs[a + 1];
    const int64_t P = out.node_pair_ptr[n_nodes];
    std::vector<int32_t> pair_node(P);
    std::vector<uint32_t> pair_ccount(P + 1, 0);
    // pass 2: sorted neighbour lists and contribution counts
#pragma omp parallel for schedule(dynamic, 2048)
    for (int64_t a = 0; a < n_nodes; ++a) {
        int32_t nb[256];
        int cnt = 0;
        for (uint32_t t = inc_ptr[a]; t < inc_ptr[a + 1]; ++t) {
            const ElemView &v = views[inc[t].kind];
            for (int b = 0; b < v.nn; ++b) {
                int32_t nbn = v.conn[v.nn * inc[t].elem + b];
                bool seen = false;
                for (int u = 0; u < cnt; ++u) if (nb[u] == nbn) { seen = true; break; }
                if (!seen) nb[cnt++] = nbn;
            }
        }
        if (nsp_ptr[a + 1] > nsp_ptr[a]) {
            bool seen = false;
            for (int u = 0; u < cnt; ++u) if (nb[u] == (int32_t)a) { seen = true; break; }
            if (!seen) nb[cnt++] = (int32_t)a;
        }
        std::sort(nb, nb + cnt);
        uint32_t p0 = out.node_pair_ptr[a];
        for (int u = 0; u < cnt; ++u) {
            pair_node[p0 + u] = nb[u];
            uint32_t cc = 0;
            if (nb[u] == (int32_t)a) cc += nsp_ptr[a + 1] - nsp_ptr[a];
            for (uint32_t t = inc_ptr[a]; t < inc_ptr[a + 1]; ++t) {
                const ElemView &v = views[inc[t].kind];
                for (int b = 0; b < v.nn; ++b)
                    if (v.conn[v.nn * inc[t].elem + b] == nb[u]) ++cc;
            }
            pair_ccount[p0 + u + 1] = cc;
        }
    }
    out.pair_cptr.assign(P + 1, 0);
    for (int64_t p = 0; p < P; ++p) out.pair_cptr[p + 1] = out.pair_cptr[p] + pair_ccount[p + 1];
    out.contrib.resize(out.pair_cptr[P]);
    // pass 3: contributions in emission order (node springs first, then springs, members, quads, plates)
#pragma omp parallel for schedule(dynamic, 2048)
    for (int64_t a = 0; a < n_nodes; ++a) {
        for (uint32_t p = out.node_pair_ptr[a]; p < out.node_pair_ptr[a + 1]; ++p) {
            int32_t b = pair_node[p];
            uint32_t w = out.pair_cptr[p];
            if (b == (int32_t)a)
                for (uint32_t q = nsp_ptr[a]; q < nsp_ptr[a + 1]; ++q)
                    out.contrib[w++] = Contrib{(uint32_t)q, (uint8_t)(nsp_dof[q] % 6), (uint8_t)(nsp_dof[q] % 6), 0, 0};
            for (uint32_t t = inc_ptr[a]; t < inc_ptr[a + 1]; ++t) {
                const ElemView &v = views[inc[t].kind];
                const int nd = 6 * v.nn;
                for (int lb = 0; lb < v.nn; ++lb)
                    if (v.conn[v.nn * inc[t].elem + lb] == b) {
                        uint64_t base = (uint64_t)v.table_off + (uint64_t)v.cls[inc[t].elem] * nd * nd;
                        out.contrib[w++] = Contrib{(uint32_t)base, inc[t].la, (uint8_t)lb, (uint8_t)nd, 0};
                    }
            }
        }
    }
    // pass 4: slot codes from the CSR pattern
    const int64_t nnz = indptr[6 * n_nodes];
    out.slot_code.assign(nnz, 0);
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 2048) reduction(| : bad)
    for (int64_t a = 0; a < n_nodes; ++a) {
        uint32_t p0 = out.node_pair_ptr[a], p1 = out.node_pair_ptr[a + 1];
        for (int i = 0; i < 6; ++i) {
            int64_t r = 6 * a + i;
            uint32_t p = p0;
            for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) {
                int32_t nbn = indices[k] / 6;
                while (p < p1 && pair_node[p] < nbn) ++p;
                if (p >= p1 || pair_node[p] != nbn) { bad = 1; break; }
                out.slot_code[k] = (uint16_t)(((p - p0) << 3) | (uint32_t)(indices[k] % 6));
            }
        }
    }
    out.ok = !bad;
}

}  // namespace pn
