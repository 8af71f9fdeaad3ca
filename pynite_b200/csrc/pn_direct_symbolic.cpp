// Host-side symbolic analysis for the multifrontal solver.  See pn_direct_symbolic.h.
#include "pn_direct_symbolic.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <numeric>
#include <stdexcept>

namespace pn {

void build_vertex_graph(int64_t n_nodes, const int32_t *newidx, const double *xyz, const std::vector<ConnView> &conn,
                        std::vector<int32_t> &vstart, std::vector<double> &coords, std::vector<int64_t> &adj_ptr,
                        std::vector<int32_t> &adj, VertexGraphScratch *scratch) {
    VertexGraphScratch local;
    VertexGraphScratch &sc = scratch ? *scratch : local;
    // vertices = nodes with at least one free DOF, in node order, so K11 rows stay contiguous per vertex
    std::vector<int32_t> &vid = sc.vid;
    vid.assign(n_nodes, -1);
    vstart.clear();
    coords.clear();
    int32_t nv = 0;
    int64_t n1 = 0;
    for (int64_t nd = 0; nd < n_nodes; ++nd) {
        int32_t first = -1, cnt = 0;
        for (int k = 0; k < 6; ++k) {
            int32_t ni = newidx[6 * nd + k];
            if (ni >= 0) { if (first < 0) first = ni; ++cnt; }
        }
        n1 += cnt;
        if (cnt) {
            vid[nd] = nv++;
            vstart.push_back(first);
            coords.push_back(xyz[3 * nd]); coords.push_back(xyz[3 * nd + 1]); coords.push_back(xyz[3 * nd + 2]);
        }
    }
    // vstart needs an end sentinel: rows are contiguous and ascending over vertices
    vstart.push_back((int32_t)n1);
    // count, fill (with duplicates, in any order: atomics), then sort + unique per vertex
    std::vector<int64_t> &deg = sc.deg, &ptr = sc.ptr, &fill = sc.fill, &cnt = sc.cnt;
    std::vector<int32_t> &raw = sc.raw;
    deg.assign(nv + 1, 0);
    for (const ConnView &cv : conn) {
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < cv.count; ++e) {
            if (cv.active && !cv.active[e]) continue;
            for (int a = 0; a < cv.nn; ++a) {
                int32_t va = vid[cv.nodes[cv.nn * e + a]];
                if (va < 0) continue;
                int64_t add = 0;
                for (int b = 0; b < cv.nn; ++b) {
                    int32_t vb = vid[cv.nodes[cv.nn * e + b]];
                    if (vb >= 0 && vb != va) ++add;
                }
                if (add) {
#pragma omp atomic
                    deg[va] += add;
                }
            }
        }
    }
    ptr.assign(nv + 1, 0);
    for (int32_t v = 0; v < nv; ++v) ptr[v + 1] = ptr[v] + deg[v];
    raw.resize(ptr[nv]);
    fill.assign(ptr.begin(), ptr.end() - 1);
    for (const ConnView &cv : conn) {
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < cv.count; ++e) {
            if (cv.active && !cv.active[e]) continue;
            for (int a = 0; a < cv.nn; ++a) {
                int32_t va = vid[cv.nodes[cv.nn * e + a]];
                if (va < 0) continue;
                for (int b = 0; b < cv.nn; ++b) {
                    int32_t vb = vid[cv.nodes[cv.nn * e + b]];
                    if (vb >= 0 && vb != va) {
                        int64_t at;
#pragma omp atomic capture
                        at = fill[va]++;
                        raw[at] = vb;
                    }
                }
            }
        }
    }
    cnt.assign(nv, 0);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int32_t v = 0; v < nv; ++v) {
        auto b = raw.begin() + ptr[v], e = raw.begin() + ptr[v + 1];
        std::sort(b, e);
        cnt[v] = std::unique(b, e) - b;
    }
    adj_ptr.assign(nv + 1, 0);
    for (int32_t v = 0; v < nv; ++v) adj_ptr[v + 1] = adj_ptr[v] + cnt[v];
    adj.resize(adj_ptr[nv]);
#pragma omp parallel for schedule(static)
    for (int32_t v = 0; v < nv; ++v) std::copy(raw.begin() + ptr[v], raw.begin() + ptr[v] + cnt[v], adj.begin() + adj_ptr[v]);
}

namespace {

// Fronts of one subproblem in post-order (children before parents), parents as local indices (-1 = root of the
// subproblem's forest).
struct SubTree {
    std::vector<std::vector<int32_t>> verts;
    std::vector<int32_t> parent;
    std::vector<int32_t> roots;
    int32_t append(SubTree &&o) {            // returns the offset the other tree's fronts got
        int32_t off = (int32_t)verts.size();
        for (auto &v : o.verts) verts.push_back(std::move(v));
        for (int32_t p : o.parent) parent.push_back(p < 0 ? -1 : p + off);
        for (int32_t r : o.roots) roots.push_back(r + off);
        return off;
    }
};

struct Builder {
    int32_t nv;
    const int32_t *vstart;
    const double *coords;
    const int64_t *adj_ptr;
    const int32_t *adj;
    int leaf_size;
    std::vector<int32_t> region;          // marker per vertex; a vertex belongs to exactly one live subproblem
    std::atomic<int32_t> next_region{1};

    // Orders the vertex set `verts` and everything it induces.  Large subproblems recurse as OpenMP tasks; the
    // result does not depend on the schedule (sub-results are concatenated in a fixed order).
    SubTree dissect(std::vector<int32_t> &verts) {
        SubTree out;
        if ((int)verts.size() <= leaf_size) {
            if (verts.empty()) return out;
            std::sort(verts.begin(), verts.end());
            out.verts.push_back(std::move(verts));
            out.parent.push_back(-1);
            out.roots.push_back(0);
            return out;
        }
        // widest coordinate axis
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int32_t v : verts)
            for (int k = 0; k < 3; ++k) {
                lo[k] = std::min(lo[k], coords[3 * v + k]);
                hi[k] = std::max(hi[k], coords[3 * v + k]);
            }
        int axis = 0;
        for (int k = 1; k < 3; ++k) if (hi[k] - lo[k] > hi[axis] - lo[axis]) axis = k;
        size_t mid = verts.size() / 2;
        auto cmp = [&](int32_t a, int32_t b) {
            double ca = coords[3 * a + axis], cb = coords[3 * b + axis];
            return ca < cb || (ca == cb && a < b);
        };
        // Large sets: the cut coordinate is the median of a fixed-stride sample (the exact median needs an nth_element
        // over indirect keys, which was a third of the serial time at the top of the recursion); the set is then
        // split by coordinate, so vertices sharing the cut coordinate stay on one side and the separators of
        // structured meshes are straight lines / planes.  Small or degenerate sets take the exact path.
        bool split_done = false;
        if (verts.size() > 8192) {
            const size_t ns = 2047, stride = verts.size() / ns;
            double sample[2047];
            for (size_t q = 0; q < ns; ++q) sample[q] = coords[3 * verts[q * stride] + axis];
            std::nth_element(sample, sample + ns / 2, sample + ns);
            const double cm = sample[ns / 2];
            auto it = std::partition(verts.begin(), verts.end(), [&](int32_t v) { return coords[3 * v + axis] < cm; });
            const size_t cut = it - verts.begin();
            if (cut > verts.size() / 4 && cut < verts.size() - verts.size() / 4) { mid = cut; split_done = true; }
        }
        if (!split_done) {
            std::nth_element(verts.begin(), verts.begin() + mid, verts.end(), cmp);
            // keep vertices with the same coordinate as the median on one side when possible
            double cm = coords[3 * verts[mid] + axis];
            size_t lo_cnt = 0;
            for (int32_t v : verts) if (coords[3 * v + axis] < cm) ++lo_cnt;
            if (lo_cnt > verts.size() / 4) {
                auto it = std::partition(verts.begin(), verts.end(), [&](int32_t v) { return coords[3 * v + axis] < cm; });
                mid = it - verts.begin();
            }
        }
        int32_t ra = next_region.fetch_add(2), rb = ra + 1;
        const bool par = verts.size() > 16384;
        const int64_t nvs = (int64_t)verts.size();
        // the large subproblems near the root are few: their vertex loops are split into tasks as well
#pragma omp taskloop grainsize(8192) if (par) default(shared)
        for (int64_t i = 0; i < nvs; ++i) region[verts[i]] = i < (int64_t)mid ? ra : rb;
        std::vector<int32_t> A, B, S;
        // separator = vertices of B touching A (B holds the median plane, which is the natural cut)
        std::vector<uint8_t> touch(par ? nvs : 0);
        if (par) {
#pragma omp taskloop grainsize(4096) default(shared)
            for (int64_t i = (int64_t)mid; i < nvs; ++i) {
                int32_t v = verts[i];
                uint8_t t = 0;
                for (int64_t k = adj_ptr[v]; k < adj_ptr[v + 1]; ++k)
                    if (region[adj[k]] == ra) { t = 1; break; }
                touch[i] = t;
            }
        }
        A.assign(verts.begin(), verts.begin() + mid);
        if (par) { B.reserve(verts.size() - mid); S.reserve(4096); }
        for (size_t i = mid; i < verts.size(); ++i) {
            int32_t v = verts[i];
            bool touches = false;
            if (par) touches = touch[i] != 0;
            else
                for (int64_t k = adj_ptr[v]; k < adj_ptr[v + 1]; ++k)
                    if (region[adj[k]] == ra) { touches = true; break; }
            (touches ? S : B).push_back(v);
        }
        std::vector<int32_t>().swap(verts);
        SubTree ta, tb;
        if (par) {
#pragma omp task shared(ta, A)
            ta = dissect(A);
#pragma omp task shared(tb, B)
            tb = dissect(B);
#pragma omp taskwait
        } else {
            ta = dissect(A);
            tb = dissect(B);
        }
        out = std::move(ta);
        out.append(std::move(tb));
        if (S.empty()) return out;          // A and B are disconnected: no separator, pass the forests up
        int32_t id = (int32_t)out.verts.size();
        std::sort(S.begin(), S.end());
        out.verts.push_back(std::move(S));
        out.parent.push_back(-1);
        for (int32_t r : out.roots) out.parent[r] = id;
        out.roots.assign(1, id);
        return out;
    }
};

}  // namespace

void direct_symbolic(int64_t n, int32_t nv, const int32_t *vstart, const double *coords, const int64_t *adj_ptr,
                     const int32_t *adj, int leaf_size, DirectSym &out) {
    out.n = n;
    out.max_children = 0; out.factor_entries = 0; out.factor_flops = 0; out.max_front = 0;
    if (n == 0 || nv == 0) {
        out.perm_pos.clear(); out.dof_front.clear(); out.fronts.clear(); out.fidx.clear(); out.fpos.clear(); out.rel.clear();
        out.level_fronts.clear();
        out.level_ptr.assign(1, 0);
        return;
    }
    Builder b;
    b.nv = nv; b.vstart = vstart; b.coords = coords; b.adj_ptr = adj_ptr; b.adj = adj; b.leaf_size = leaf_size;
    b.region.swap(out.s_region);
    b.region.assign(nv, 0);
    std::vector<int32_t> all(nv);
    std::iota(all.begin(), all.end(), 0);
    SubTree tree;
    const bool timing = getenv("PN_DEBUG_TIMING") != nullptr;
    auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    double t0 = now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        double t = now();
        fprintf(stderr, "[pn-time] direct_symbolic    %-28s %8.2f ms\n", what, t - t0);
        t0 = t;
    };
#pragma omp parallel
#pragma omp single
    tree = b.dissect(all);
    lap("dissect");
    std::vector<std::vector<int32_t>> &front_verts = tree.verts;
    std::vector<int32_t> &parent_of = tree.parent;
    const int32_t nf = (int32_t)front_verts.size();

    // vertex elimination positions (post-order front ids = elimination order) and DOF positions
    b.region.swap(out.s_region);                       // keep the capacity for the next call
    std::vector<int32_t> &vfront = out.s_vfront, &vorder = out.s_vorder;      // vorder: rank of the vertex in elimination order
    vfront.resize(nv); vorder.resize(nv);
    std::vector<int32_t> &vert_seq = out.s_vert_seq;   // vertices in elimination order
    vert_seq.clear();
    vert_seq.reserve(nv);
    std::vector<int32_t> &f_vbeg = out.s_f_vbeg;
    f_vbeg.assign(nf + 1, 0);
    for (int32_t f = 0; f < nf; ++f) {
        f_vbeg[f] = (int32_t)vert_seq.size();
        for (int32_t v : front_verts[f]) { vfront[v] = f; vorder[v] = (int32_t)vert_seq.size(); vert_seq.push_back(v); }
    }
    f_vbeg[nf] = (int32_t)vert_seq.size();
    if ((int32_t)vert_seq.size() != nv) throw std::runtime_error("direct_symbolic: vertices lost in dissection");
    std::vector<int32_t> &vdofpos = out.s_vdofpos;     // DOF elimination position of vertex with order r
    vdofpos.assign(nv + 1, 0);
    for (int32_t r = 0; r < nv; ++r) {
        int32_t v = vert_seq[r];
        vdofpos[r + 1] = vdofpos[r] + (vstart[v + 1] - vstart[v]);
    }
    out.perm_pos.resize(n);
    out.dof_front.resize(n);
#pragma omp parallel for schedule(static)
    for (int32_t r = 0; r < nv; ++r) {            // every K11 row belongs to exactly one vertex
        int32_t v = vert_seq[r];
        for (int32_t k = 0; k < vstart[v + 1] - vstart[v]; ++k) {
            out.perm_pos[vstart[v] + k] = vdofpos[r] + k;
            out.dof_front[vstart[v] + k] = vfront[v];
        }
    }

    lap("orders");
    // boundary vertex sets, bottom-up (as vertex orders, ascending); fronts of one depth are independent
    std::vector<std::vector<int32_t>> bnd(nf);
    std::vector<std::vector<int32_t>> children(nf);
    for (int32_t f = 0; f < nf; ++f) if (parent_of[f] >= 0) children[parent_of[f]].push_back(f);
    std::vector<int32_t> depth(nf, 0);
    int32_t maxdepth = 0;
    for (int32_t f = nf - 1; f >= 0; --f) {
        depth[f] = parent_of[f] < 0 ? 0 : depth[parent_of[f]] + 1;
        maxdepth = std::max(maxdepth, depth[f]);
    }
    std::vector<std::vector<int32_t>> by_depth(maxdepth + 1);
    for (int32_t f = 0; f < nf; ++f) by_depth[depth[f]].push_back(f);
    int bad_root = 0;
    for (int32_t dd = maxdepth; dd >= 0; --dd) {
        const std::vector<int32_t> &grp = by_depth[dd];
#pragma omp parallel for schedule(dynamic, 16) if (grp.size() > 32)
        for (int64_t q = 0; q < (int64_t)grp.size(); ++q) {
            const int32_t f = grp[q];
            int32_t vend = f_vbeg[f + 1];
            std::vector<int32_t> cur;
            for (int32_t v : front_verts[f])
                for (int64_t k = adj_ptr[v]; k < adj_ptr[v + 1]; ++k) {
                    int32_t o = vorder[adj[k]];
                    if (o >= vend) cur.push_back(o);
                }
            std::sort(cur.begin(), cur.end());
            cur.erase(std::unique(cur.begin(), cur.end()), cur.end());
            for (int32_t c : children[f]) {
                std::vector<int32_t> merged;
                merged.reserve(cur.size() + bnd[c].size());
                auto first = std::lower_bound(bnd[c].begin(), bnd[c].end(), vend);
                std::set_union(cur.begin(), cur.end(), first, bnd[c].end(), std::back_inserter(merged));
                cur.swap(merged);
            }
            bnd[f].swap(cur);
            // every boundary vertex must live in an ancestor; a root must have an empty boundary
            if (parent_of[f] < 0 && !bnd[f].empty()) {
#pragma omp atomic write
                bad_root = 1;
            }
        }
    }
    if (bad_root) throw std::runtime_error("direct_symbolic: root front with a non-empty boundary");

    lap("boundary sets");
    // fronts, index lists
    out.fronts.assign(nf, FrontSym());
    int64_t total_idx = 0, total_rel = 0;
    for (int32_t f = 0; f < nf; ++f) {
        FrontSym &F = out.fronts[f];
        F.parent = parent_of[f];
        F.depth = depth[f];
        F.pos_start = vdofpos[f_vbeg[f]];
        F.ns = vdofpos[f_vbeg[f + 1]] - vdofpos[f_vbeg[f]];
        int32_t nbd = 0;
        for (int32_t o : bnd[f]) nbd += vdofpos[o + 1] - vdofpos[o];
        F.nb = nbd;
        F.idx_off = total_idx;
        F.rel_off = total_rel;
        total_idx += F.ns + F.nb;
        total_rel += F.nb;
        out.max_front = std::max<int64_t>(out.max_front, F.ns + F.nb);
        double ns = F.ns, nbv = F.nb;
        out.factor_entries += (int64_t)(F.ns + F.nb) * F.ns + (int64_t)F.ns * F.nb;
        out.factor_flops += 2.0 / 3.0 * ns * ns * ns + 2.0 * ns * ns * nbv + 2.0 * ns * nbv * nbv;
    }
    for (int32_t f = 0; f < nf; ++f) {
        int32_t rank = 0;
        for (int32_t c : children[f]) out.fronts[c].child_rank = rank++;
        out.max_children = std::max(out.max_children, rank);
    }
    lap("front sizes");
    // inverse of perm_pos for index lists in K11 numbering
    std::vector<int32_t> &pos_dof = out.s_pos_dof;
    pos_dof.resize(n);
    for (int64_t d = 0; d < n; ++d) pos_dof[out.perm_pos[d]] = (int32_t)d;
    out.fpos.resize(total_idx);
    out.fidx.resize(total_idx);
    out.rel.resize(total_rel);
    lap("alloc lists");
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t f = 0; f < nf; ++f) {
        const FrontSym &F = out.fronts[f];
        int64_t o = F.idx_off;
        for (int32_t k = 0; k < F.ns; ++k) out.fpos[o++] = F.pos_start + k;
        for (int32_t vo : bnd[f])
            for (int32_t p = vdofpos[vo]; p < vdofpos[vo + 1]; ++p) out.fpos[o++] = p;
        for (int64_t k = F.idx_off; k < F.idx_off + F.ns + F.nb; ++k) out.fidx[k] = pos_dof[out.fpos[k]];
    }
    lap("index lists");
    int bad_rel = 0;
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t f = 0; f < nf; ++f) {
        const FrontSym &F = out.fronts[f];
        if (F.parent < 0) continue;
        const FrontSym &P = out.fronts[F.parent];
        const int32_t *pl = out.fpos.data() + P.idx_off;
        int32_t plen = P.ns + P.nb;
        int32_t q = 0;
        for (int32_t k = 0; k < F.nb; ++k) {
            int32_t p = out.fpos[F.idx_off + F.ns + k];
            while (q < plen && pl[q] < p) ++q;
            if (q >= plen || pl[q] != p) {
#pragma omp atomic write
                bad_rel = 1;
                break;
            }
            out.rel[F.rel_off + k] = q;
        }
    }
    if (bad_rel) throw std::runtime_error("direct_symbolic: child boundary not contained in parent front");
    lap("rel maps");
    // levels: level = maxdepth - depth, so children sit exactly one level below their parent
    int32_t nl = maxdepth + 1;
    out.level_ptr.assign(nl + 1, 0);
    for (int32_t f = 0; f < nf; ++f) ++out.level_ptr[maxdepth - depth[f] + 1];
    for (int32_t l = 0; l < nl; ++l) out.level_ptr[l + 1] += out.level_ptr[l];
    out.level_fronts.resize(nf);
    std::vector<int32_t> fillp(out.level_ptr.begin(), out.level_ptr.end() - 1);
    for (int32_t f = 0; f < nf; ++f) out.level_fronts[fillp[maxdepth - depth[f]]++] = f;
}

}  // namespace pn
