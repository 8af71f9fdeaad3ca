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
                        const int32_t *indices, uint32_t zero_index, AssemblySym &out) {
    out.ok = false;
    out.recipes_ok = false;
    out.width = 0;
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
    out.rowpair_code.assign((size_t)6 * P, 0);
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
                uint16_t &rp = out.rowpair_code[(size_t)6 * p0 + (size_t)i * (p1 - p0) + (p - p0)];
                if (!(rp & 63u)) {
                    uint32_t off = (uint32_t)(k - indptr[r]);
                    if (off > 1023u) { bad = 1; break; }
                    rp = (uint16_t)(off << 6);
                }
                rp |= (uint16_t)(1u << (indices[k] % 6));
            }
        }
    }
    out.ok = !bad;
    if (!out.ok) return;

    // ---- node recipes ---------------------------------------------------------------------------------
    int width = 0;
    for (int64_t p = 0; p < P; ++p) {
        int cnt = 0;
        for (uint32_t q = out.pair_cptr[p]; q < out.pair_cptr[p + 1]; ++q) cnt += out.contrib[q].nd != 0;
        width = std::max(width, cnt);
    }
    width = std::max(4, (width + 3) & ~3);        // tuples are read four indices at a time
    if (width > 12) return;                       // recipes_ok stays false: the per-slot gather kernel is used
    out.width = width;
    out.nsp_first.assign(nsp_ptr.begin(), nsp_ptr.end() - 1);
    // descriptor hash per node: pair count, per pair the contributions (support springs by local index), the
    // row/pair codes and the six row lengths -- everything the slot-by-slot addend lists are a function of
    std::vector<uint64_t> nh(n_nodes);
    auto mix = [](uint64_t h, uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; return (h ^ x) * 1099511628211ull; };
    auto contrib_word = [&](int64_t a, const Contrib &ct) -> uint64_t {
        uint64_t base = ct.nd ? ct.base : (uint64_t)(ct.base - nsp_ptr[a]);
        return (base << 24) | ((uint64_t)ct.la << 16) | ((uint64_t)ct.lb << 8) | ct.nd;
    };
#pragma omp parallel for schedule(dynamic, 2048)
    for (int64_t a = 0; a < n_nodes; ++a) {
        uint32_t p0 = out.node_pair_ptr[a], p1 = out.node_pair_ptr[a + 1];
        uint64_t h = 1469598103934665603ull;
        h = mix(h, p1 - p0);
        for (uint32_t p = p0; p < p1; ++p) {
            h = mix(h, out.pair_cptr[p + 1] - out.pair_cptr[p]);
            for (uint32_t q = out.pair_cptr[p]; q < out.pair_cptr[p + 1]; ++q) h = mix(h, contrib_word(a, out.contrib[q]));
        }
        for (uint32_t t = 6 * p0; t < 6 * p1; ++t) h = mix(h, out.rowpair_code[t]);
        for (int i = 0; i < 6; ++i) h = mix(h, (uint64_t)(indptr[6 * a + i + 1] - indptr[6 * a + i]));
        nh[a] = h;
    }
    auto same_desc = [&](int64_t a, int64_t b) {
        uint32_t pa = out.node_pair_ptr[a], pb = out.node_pair_ptr[b];
        uint32_t na = out.node_pair_ptr[a + 1] - pa;
        if (na != out.node_pair_ptr[b + 1] - pb) return false;
        for (int i = 0; i < 6; ++i)
            if (indptr[6 * a + i + 1] - indptr[6 * a + i] != indptr[6 * b + i + 1] - indptr[6 * b + i]) return false;
        for (uint32_t u = 0; u < na; ++u) {
            uint32_t ca = out.pair_cptr[pa + u], cb = out.pair_cptr[pb + u];
            uint32_t nca = out.pair_cptr[pa + u + 1] - ca;
            if (nca != out.pair_cptr[pb + u + 1] - cb) return false;
            for (uint32_t q = 0; q < nca; ++q)
                if (contrib_word(a, out.contrib[ca + q]) != contrib_word(b, out.contrib[cb + q])) return false;
        }
        for (uint32_t t = 0; t < 6 * na; ++t)
            if (out.rowpair_code[6 * pa + t] != out.rowpair_code[6 * pb + t]) return false;
        return true;
    };
    std::vector<int64_t> order(n_nodes);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return nh[x] != nh[y] ? nh[x] < nh[y] : x < y; });
    out.node_recipe.assign(n_nodes, 0);
    std::vector<int64_t> rec_rep;
    {
        // runs of equal hash: every member is compared (in parallel) with the run's first node; the rare mismatches
        // (hash collisions) are matched serially against the run's further representatives
        std::vector<uint8_t> match(n_nodes, 0);
        int64_t q = 0;
        while (q < n_nodes) {
            int64_t r = q;
            while (r < n_nodes && nh[order[r]] == nh[order[q]]) ++r;
            const int64_t rep0 = order[q];
#pragma omp parallel for schedule(static) if (r - q > 4096)
            for (int64_t t = q; t < r; ++t) match[t] = same_desc(rep0, order[t]) ? 1 : 0;
            std::vector<std::pair<int64_t, uint32_t>> run;          // (representative node, recipe id)
            uint32_t id0 = (uint32_t)rec_rep.size();
            rec_rep.push_back(rep0);
            for (int64_t t = q; t < r; ++t) {
                int64_t a = order[t];
                if (match[t]) { out.node_recipe[a] = id0; continue; }
                int64_t hit = -1;
                for (auto &rr : run)
                    if (same_desc(rr.first, a)) { hit = rr.second; break; }
                if (hit < 0) { hit = (int64_t)rec_rep.size(); rec_rep.push_back(a); run.push_back({a, (uint32_t)hit}); }
                out.node_recipe[a] = (uint32_t)hit;
            }
            q = r;
        }
    }
    const int64_t R = (int64_t)rec_rep.size();
    out.recipe_ptr.assign(R + 1, 0);
    for (int64_t r = 0; r < R; ++r) {
        int64_t a = rec_rep[r];
        out.recipe_ptr[r + 1] = out.recipe_ptr[r] + (uint32_t)(indptr[6 * a + 6] - indptr[6 * a]);
    }
    const int64_t S = out.recipe_ptr[R];
    out.recipe_tuples.assign((size_t)S * width, zero_index);
    out.recipe_nsp.assign(S, 0xff);
    int bad2 = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(| : bad2)
    for (int64_t r = 0; r < R; ++r) {
        int64_t a = rec_rep[r];
        uint32_t p0 = out.node_pair_ptr[a];
        for (int i = 0; i < 6; ++i)
            for (int32_t k = indptr[6 * a + i]; k < indptr[6 * a + i + 1]; ++k) {
                int64_t slot = out.recipe_ptr[r] + (k - indptr[6 * a]);
                uint32_t code = out.slot_code[k];
                uint32_t p = p0 + (code >> 3);
                int j = (int)(code & 7u), w = 0;
                for (uint32_t q = out.pair_cptr[p]; q < out.pair_cptr[p + 1]; ++q) {
                    const Contrib &ct = out.contrib[q];
                    if (ct.nd == 0) {
                        if ((int)ct.la == i && j == i) {
                            uint32_t local = ct.base - nsp_ptr[a];
                            if (local >= 0xff || out.recipe_nsp[slot] != 0xff) bad2 = 1;
                            out.recipe_nsp[slot] = (uint8_t)local;
                        }
                    } else {
                        out.recipe_tuples[(size_t)slot * width + w++] = ct.base + (uint32_t)((6 * ct.la + i) * ct.nd + 6 * ct.lb + j);
                    }
                }
            }
    }
    out.recipes_ok = !bad2;
}

}  // namespace pn
