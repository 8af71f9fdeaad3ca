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

DofClasses predict_dof_classes(const double *xyz, const std::vector<ConnView> &shells, int64_t n_line_elements) {
    DofClasses out;
    if (n_line_elements > 0) return out;
    // planes met: bit g = some shell has all four nodes at one coordinate g (bitwise equal: the element triad is then
    // built from differences that are exact zeros along g); bit 3 = a shell in general position
    unsigned planes = 0;
    int64_t total = 0;
    for (const ConnView &cv : shells) {
        total += cv.count;
        unsigned local = 0;
#pragma omp parallel for schedule(static) reduction(| : local)
        for (int64_t e = 0; e < cv.count; ++e) {
            if (cv.active && !cv.active[e]) continue;
            const int32_t *nd = cv.nodes + cv.nn * e;
            int g = -1, hits = 0;
            for (int k = 0; k < 3; ++k) {
                bool same = true;
                for (int a = 1; a < cv.nn; ++a) same = same && xyz[3 * (int64_t)nd[a] + k] == xyz[3 * (int64_t)nd[0] + k];
                if (same) { g = k; ++hits; }
            }
            local |= hits == 1 ? 1u << g : 8u;
        }
        planes |= local;
    }
    if (!total || (planes & 8u)) return out;
    int parent[6] = {0, 1, 2, 3, 4, 5};
    auto find = [&](int a) { while (parent[a] != a) a = parent[a]; return a; };
    auto join = [&](int a, int b) { a = find(a); b = find(b); if (a != b) parent[std::max(a, b)] = std::min(a, b); };
    for (int g = 0; g < 3; ++g) {
        if (!(planes >> g & 1u)) continue;
        const int h1 = (g + 1) % 3, h2 = (g + 2) % 3;
        join(h1, h2);                          // membrane
        join(g, 3 + h1); join(g, 3 + h2);      // bending: normal translation + in-plane rotations
    }                                          // (drilling rotation 3 + g: diagonal only)
    int id[6], n = 0;
    for (int k = 0; k < 6; ++k) id[k] = -1;
    for (int k = 0; k < 6; ++k) {
        const int r = find(k);
        if (id[r] < 0) id[r] = n++;
        out.of_type[k] = (uint8_t)id[r];
    }
    out.n = n;
    // a class made of drilling rotations only: every member is the normal rotation of a plane that was met and was
    // joined to nothing
    for (int c = 0; c < n; ++c) {
        bool only_drilling = true;
        for (int k = 0; k < 6; ++k)
            if (out.of_type[k] == c && !(k >= 3 && (planes >> (k - 3) & 1u))) only_drilling = false;
        out.diagonal[c] = only_drilling ? 1 : 0;
    }
    return out;
}

void build_vertex_graph(int64_t n_nodes, const int32_t *newidx, const double *xyz, const std::vector<ConnView> &conn,
                        std::vector<int32_t> &vstart, std::vector<double> &coords, std::vector<int64_t> &adj_ptr,
                        std::vector<int32_t> &adj, VertexGraphScratch *scratch) {
    VertexGraphScratch local;
    VertexGraphScratch &sc = scratch ? *scratch : local;
    // vertices = nodes with at least one free DOF, in node order, so K11 rows stay contiguous per vertex
    std::vector<int32_t> &vid = sc.vid;
    vid.resize(n_nodes);
    // free DOFs per node (chunked count, prefix over the chunks, fill)
    constexpr int64_t CH = 16384;
    const int64_t nch = (n_nodes + CH - 1) / CH;
    std::vector<int64_t> cv_(nch + 1, 0), cr_(nch + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t ch = 0; ch < nch; ++ch) {
        int64_t v = 0, r = 0;
        for (int64_t nd = ch * CH; nd < std::min(n_nodes, (ch + 1) * CH); ++nd) {
            int cnt = 0;
            for (int k = 0; k < 6; ++k) cnt += newidx[6 * nd + k] >= 0;
            v += cnt > 0; r += cnt;
        }
        cv_[ch + 1] = v; cr_[ch + 1] = r;
    }
    for (int64_t ch = 0; ch < nch; ++ch) { cv_[ch + 1] += cv_[ch]; cr_[ch + 1] += cr_[ch]; }
    const int32_t nv = (int32_t)cv_[nch];
    const int64_t n1 = cr_[nch];
    vstart.resize((size_t)nv);
    coords.resize(3 * (size_t)nv);
#pragma omp parallel for schedule(static)
    for (int64_t ch = 0; ch < nch; ++ch) {
        int32_t v = (int32_t)cv_[ch];
        for (int64_t nd = ch * CH; nd < std::min(n_nodes, (ch + 1) * CH); ++nd) {
            int32_t first = -1;
            for (int k = 0; k < 6; ++k) {
                const int32_t ni = newidx[6 * nd + k];
                if (ni >= 0 && first < 0) first = ni;
            }
            if (first < 0) { vid[nd] = -1; continue; }
            vid[nd] = v;
            vstart[v] = first;
            coords[3 * (size_t)v] = xyz[3 * nd]; coords[3 * (size_t)v + 1] = xyz[3 * nd + 1]; coords[3 * (size_t)v + 2] = xyz[3 * nd + 2];
            ++v;
        }
    }
    // vstart needs an end sentinel: rows are contiguous and ascending over vertices
    vstart.push_back((int32_t)n1);
    // count, fill (with duplicates, in any order: atomics), then sort + unique per vertex
    std::vector<int64_t> &deg = sc.deg, &ptr = sc.ptr, &fill = sc.fill;
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
                int32_t nb_[8];
                int add = 0;
                for (int b = 0; b < cv.nn; ++b) {
                    int32_t vb = vid[cv.nodes[cv.nn * e + b]];
                    if (vb >= 0 && vb != va && add < 8) nb_[add++] = vb;
                }
                if (!add) continue;
                int64_t at;
#pragma omp atomic capture
                { at = fill[va]; fill[va] += add; }
                for (int q = 0; q < add; ++q) raw[at + q] = nb_[q];
            }
        }
    }
    // The lists keep their duplicates (a neighbour shared by two elements appears twice) and the order the threads
    // produced: the dissection only asks "does v touch a vertex of the other side" and the boundary sets are sorted and
    // made unique where they are built, so neither shows in the result; sorting 250 000 lists cost more than it saved.
    adj_ptr.assign(ptr.begin(), ptr.end());
    adj.swap(raw);
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
        const bool par = verts.size() > 16384;
        const int64_t nvs = (int64_t)verts.size();
        const int64_t verts_n = nvs;              // (verts is released before the recursion)
        // the large subproblems near the root are few, so their passes over the vertex set are cut into chunks that run
        // as tasks (bounding box, side of the cut, contact with the other side, ordered compaction): the serial passes
        // of the top three levels were half of the ordering time.  Chunk results are combined in chunk order, so the
        // tree does not depend on the schedule.
        constexpr int64_t CH = 8192;
        const int64_t nch = (nvs + CH - 1) / CH;
        // widest coordinate axis
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        if (par) {
            std::vector<double> box((size_t)nch * 6);
#pragma omp taskloop grainsize(1) default(shared)
            for (int64_t ch = 0; ch < nch; ++ch) {
                double l[3] = {1e300, 1e300, 1e300}, h[3] = {-1e300, -1e300, -1e300};
                const int64_t e = std::min(nvs, (ch + 1) * CH);
                for (int64_t i = ch * CH; i < e; ++i)
                    for (int k = 0; k < 3; ++k) {
                        const double x = coords[3 * verts[i] + k];
                        l[k] = std::min(l[k], x); h[k] = std::max(h[k], x);
                    }
                for (int k = 0; k < 3; ++k) { box[ch * 6 + k] = l[k]; box[ch * 6 + 3 + k] = h[k]; }
            }
            for (int64_t ch = 0; ch < nch; ++ch)
                for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], box[ch * 6 + k]); hi[k] = std::max(hi[k], box[ch * 6 + 3 + k]); }
        } else {
            for (int32_t v : verts)
                for (int k = 0; k < 3; ++k) {
                    lo[k] = std::min(lo[k], coords[3 * v + k]);
                    hi[k] = std::max(hi[k], coords[3 * v + k]);
                }
        }
        int axis = 0;
        for (int k = 1; k < 3; ++k) if (hi[k] - lo[k] > hi[axis] - lo[axis]) axis = k;
        size_t mid = verts.size() / 2;
        auto cmp = [&](int32_t a, int32_t b) {
            double ca = coords[3 * a + axis], cb = coords[3 * b + axis];
            return ca < cb || (ca == cb && a < b);
        };
        std::vector<int32_t> A, B, S;
        // Large sets: the cut coordinate is the median of a fixed-stride sample (the exact median needs an nth_element
        // over indirect keys, which was a third of the serial time at the top of the recursion); the set is then
        // split by coordinate, so vertices sharing the cut coordinate stay on one side and the separators of
        // structured meshes are straight lines / planes.  Small or degenerate sets take the exact path.
        bool split_done = false;
        int32_t ra = 0, rb = 0;
        if (verts.size() > 8192) {
            const size_t ns = 2047, stride = verts.size() / ns;
            double sample[2047];
            for (size_t q = 0; q < ns; ++q) sample[q] = coords[3 * verts[q * stride] + axis];
            std::nth_element(sample, sample + ns / 2, sample + ns);
            const double cm = sample[ns / 2];
            if (par) {
                // side of every vertex and the region marks in one pass; sizes per chunk
                std::vector<uint8_t> side(nvs);
                std::vector<int64_t> cntA(nch + 1, 0), cntB(nch + 1, 0), cntS(nch + 1, 0);
                ra = next_region.fetch_add(2); rb = ra + 1;
#pragma omp taskloop grainsize(1) default(shared)
                for (int64_t ch = 0; ch < nch; ++ch) {
                    const int64_t e = std::min(nvs, (ch + 1) * CH);
                    int64_t a = 0;
                    for (int64_t i = ch * CH; i < e; ++i) {
                        const bool low = coords[3 * verts[i] + axis] < cm;
                        side[i] = low ? 0 : 1;
                        region[verts[i]] = low ? ra : rb;
                        a += low;
                    }
                    cntA[ch + 1] = a;
                }
                int64_t cut = 0;
                for (int64_t ch = 0; ch < nch; ++ch) cut += cntA[ch + 1];
                if (cut > nvs / 4 && cut < nvs - nvs / 4) {
                    // separator = vertices of the upper side touching the lower side (the upper side holds the median plane)
#pragma omp taskloop grainsize(1) default(shared)
                    for (int64_t ch = 0; ch < nch; ++ch) {
                        const int64_t e = std::min(nvs, (ch + 1) * CH);
                        int64_t nb_ = 0, ns_ = 0;
                        for (int64_t i = ch * CH; i < e; ++i) {
                            if (!side[i]) continue;
                            const int32_t v = verts[i];
                            bool t = false;
                            for (int64_t k = adj_ptr[v]; k < adj_ptr[v + 1]; ++k)
                                if (region[adj[k]] == ra) { t = true; break; }
                            if (t) { side[i] = 2; ++ns_; } else ++nb_;
                        }
                        cntB[ch + 1] = nb_; cntS[ch + 1] = ns_;
                    }
                    for (int64_t ch = 0; ch < nch; ++ch) { cntA[ch + 1] += cntA[ch]; cntB[ch + 1] += cntB[ch]; cntS[ch + 1] += cntS[ch]; }
                    A.resize(cntA[nch]); B.resize(cntB[nch]); S.resize(cntS[nch]);
#pragma omp taskloop grainsize(1) default(shared)
                    for (int64_t ch = 0; ch < nch; ++ch) {
                        const int64_t e = std::min(nvs, (ch + 1) * CH);
                        int64_t pa = cntA[ch], pb = cntB[ch], ps = cntS[ch];
                        for (int64_t i = ch * CH; i < e; ++i) {
                            if (side[i] == 0) A[pa++] = verts[i];
                            else if (side[i] == 1) B[pb++] = verts[i];
                            else S[ps++] = verts[i];
                        }
                    }
                    split_done = true;
                }
            } else {
                auto it = std::partition(verts.begin(), verts.end(), [&](int32_t v) { return coords[3 * v + axis] < cm; });
                const size_t cut = it - verts.begin();
                if (cut > verts.size() / 4 && cut < verts.size() - verts.size() / 4) { mid = cut; split_done = true; }
            }
        }
        if (!(par && split_done)) {
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
            ra = next_region.fetch_add(2); rb = ra + 1;
            for (int64_t i = 0; i < nvs; ++i) region[verts[i]] = i < (int64_t)mid ? ra : rb;
            // separator = vertices of B touching A (B holds the median plane, which is the natural cut)
            A.assign(verts.begin(), verts.begin() + mid);
            for (size_t i = mid; i < verts.size(); ++i) {
                int32_t v = verts[i];
                bool touches = false;
                for (int64_t k = adj_ptr[v]; k < adj_ptr[v + 1]; ++k)
                    if (region[adj[k]] == ra) { touches = true; break; }
                (touches ? S : B).push_back(v);
            }
        }
        std::vector<int32_t>().swap(verts);
        SubTree ta, tb;
        // the two halves as tasks down to 2 048 vertices: with the 16 384 of the chunked passes there were 16 serial
        // subproblems for 15 threads, i.e. two rounds
        if (verts_n > 2048) {
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

// Joins a forest into a set of trees of one height.  Roots of the tallest trees stay roots (they share the top level);
// every shorter tree has an empty boundary at its root, so giving that root a parent adds nothing to the parent's
// matrix -- the link only says on which level, and (subtree-to-GPU solve) on which rank, the tree is processed.  A tree
// of height h hangs from a front of a tallest tree at depth H - 1 - h, so that its leaves land on the bottom level;
// candidates of that depth are taken round-robin, those with the fewest children first.  Returns the post-order
// (new id -> old id) and rewrites `parent` in new ids.
static std::vector<int32_t> join_forest(std::vector<int32_t> &parent) {
    const int32_t nf = (int32_t)parent.size();
    std::vector<int32_t> depth(nf, 0), root_of(nf, 0), height(nf, 0), nchild(nf, 0);
    for (int32_t f = nf - 1; f >= 0; --f) {           // parents have larger ids
        const int32_t p = parent[f];
        depth[f] = p < 0 ? 0 : depth[p] + 1;
        root_of[f] = p < 0 ? f : root_of[p];
        if (p >= 0) ++nchild[p];
    }
    int32_t H = 0;
    for (int32_t f = 0; f < nf; ++f) {
        height[root_of[f]] = std::max(height[root_of[f]], depth[f] + 1);
        H = std::max(H, depth[f] + 1);
    }
    std::vector<std::vector<int32_t>> cand(H);
    for (int pass = 0; pass <= 2; ++pass)
        for (int32_t f = 0; f < nf; ++f)
            if (height[root_of[f]] == H && std::min(nchild[f], 2) == pass) cand[depth[f]].push_back(f);
    std::vector<size_t> next(H, 0);
    for (int32_t f = 0; f < nf; ++f) {
        if (parent[f] >= 0 || height[f] == H) continue;
        const int32_t dp = std::max(0, H - 1 - height[f]);
        parent[f] = cand[dp][next[dp]++ % cand[dp].size()];
    }
    // Renumbering.  The fronts of the tallest trees keep their relative order -- the caller lists them by (node front,
    // class), i.e. node-tree post-order with the classes of one separator side by side, so that the fronts of one
    // geometric region form one id range on every level (the subtree-to-GPU plan hands such ranges to the ranks) --
    // and every attached tree is emitted, in post-order, right before the front it hangs from.
    std::vector<int32_t> cptr(nf + 1, 0), cidx(nf);
    for (int32_t f = 0; f < nf; ++f) if (parent[f] >= 0) ++cptr[parent[f] + 1];
    for (int32_t f = 0; f < nf; ++f) cptr[f + 1] += cptr[f];
    {
        std::vector<int32_t> fillp(cptr.begin(), cptr.end() - 1);
        for (int32_t f = 0; f < nf; ++f) if (parent[f] >= 0) cidx[fillp[parent[f]]++] = f;
    }
    std::vector<int32_t> new_id(nf, -1), order;
    order.reserve(nf);
    std::vector<std::pair<int32_t, int32_t>> stack;     // (front, next child slot)
    auto emit_tree = [&](int32_t r) {                   // post-order of the (short) tree rooted at r
        stack.push_back({r, cptr[r]});
        while (!stack.empty()) {
            auto &top = stack.back();
            if (top.second < cptr[top.first + 1]) {
                const int32_t ch = cidx[top.second++];
                stack.push_back({ch, cptr[ch]});
            } else {
                new_id[top.first] = (int32_t)order.size();
                order.push_back(top.first);
                stack.pop_back();
            }
        }
    };
    for (int32_t f = 0; f < nf; ++f) {
        if (height[root_of[f]] != H) continue;          // (root_of / height: of the forest before the join)
        for (int32_t k = cptr[f]; k < cptr[f + 1]; ++k)
            if (height[root_of[cidx[k]]] != H) emit_tree(cidx[k]);
        new_id[f] = (int32_t)order.size();
        order.push_back(f);
    }
    if ((int32_t)order.size() != nf) throw std::runtime_error("direct_symbolic: forest join lost fronts");
    std::vector<int32_t> np(nf);
    for (int32_t q = 0; q < nf; ++q) np[q] = parent[order[q]] < 0 ? -1 : new_id[parent[order[q]]];
    parent.swap(np);
    return order;
}

void direct_symbolic(int64_t n, int32_t nv, const int32_t *vstart, const double *coords, const int64_t *adj_ptr,
                     const int32_t *adj, int leaf_size, DirectSym &out, const uint8_t *row_class, int n_classes,
                     const uint8_t *class_diagonal) {
    out.n = n;
    out.max_children = 0; out.factor_entries = 0; out.factor_flops = 0; out.max_front = 0;
    if (n == 0 || nv == 0) {
        out.perm_pos.clear(); out.dof_front.clear(); out.fronts.clear(); out.fidx.clear(); out.fpos.clear(); out.rel.clear();
        out.level_fronts.clear();
        out.level_ptr.assign(1, 0);
        return;
    }
    const int C = (row_class && n_classes > 1) ? n_classes : 1;
    const bool timing = getenv("PN_DEBUG_TIMING") != nullptr;
    auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    double t0 = now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        double t = now();
        fprintf(stderr, "[pn-time] direct_symbolic    %-28s %8.2f ms\n", what, t - t0);
        t0 = t;
    };
    // rows of every (vertex, class); with classes the leaves are scaled so that the largest class keeps about
    // 6 * leaf_size pivots per leaf
    std::vector<uint8_t> vc_cnt;
    int leaf_nodes = leaf_size;
    if (C > 1) {
        vc_cnt.assign((size_t)nv * C, 0);
        std::vector<int64_t> rows_c(C, 0);
#pragma omp parallel for schedule(static)
        for (int32_t v = 0; v < nv; ++v)
            for (int32_t r = vstart[v]; r < vstart[v + 1]; ++r) ++vc_cnt[(size_t)v * C + row_class[r]];
        for (int64_t r = 0; r < n; ++r) ++rows_c[row_class[r]];
        int64_t big = 1;
        for (int c = 0; c < C; ++c) big = std::max(big, rows_c[c]);
        leaf_nodes = (int)std::max<int64_t>(leaf_size, (6 * (int64_t)leaf_size * nv + big / 2) / big);
    }
    auto cnt_of = [&](int32_t v, int c) -> int32_t { return C > 1 ? vc_cnt[(size_t)v * C + c] : vstart[v + 1] - vstart[v]; };

    Builder b;
    b.nv = nv; b.vstart = vstart; b.coords = coords; b.adj_ptr = adj_ptr; b.adj = adj; b.leaf_size = leaf_nodes;
    b.region.swap(out.s_region);
    b.region.assign(nv, 0);
    std::vector<int32_t> all(nv);
    std::iota(all.begin(), all.end(), 0);
    SubTree tree;
#pragma omp parallel
#pragma omp single
    tree = b.dissect(all);
    lap("dissect");
    std::vector<std::vector<int32_t>> &front_verts = tree.verts;
    std::vector<int32_t> &nparent = tree.parent;
    const int32_t nf0 = (int32_t)front_verts.size();       // fronts of the node tree

    // node level: elimination order of the vertices (post-order front ids = elimination order)
    b.region.swap(out.s_region);                       // keep the capacity for the next call
    std::vector<int32_t> &vfront = out.s_vfront, &vorder = out.s_vorder;      // vorder: rank of the vertex in elimination order
    vfront.resize(nv); vorder.resize(nv);
    std::vector<int32_t> &vert_seq = out.s_vert_seq;   // vertices in elimination order
    std::vector<int32_t> &f_vbeg = out.s_f_vbeg;
    f_vbeg.assign(nf0 + 1, 0);
    for (int32_t f = 0; f < nf0; ++f) f_vbeg[f + 1] = f_vbeg[f] + (int32_t)front_verts[f].size();
    if (f_vbeg[nf0] != nv) throw std::runtime_error("direct_symbolic: vertices lost in dissection");
    vert_seq.resize(nv);
#pragma omp parallel for schedule(dynamic, 256)
    for (int32_t f = 0; f < nf0; ++f) {
        int32_t o = f_vbeg[f];
        for (int32_t v : front_verts[f]) { vfront[v] = f; vorder[v] = o; vert_seq[o++] = v; }
    }
    lap("orders");

    // boundary vertex sets of the node tree, bottom-up (as vertex orders, ascending); fronts of one depth are independent
    std::vector<std::vector<int32_t>> bnd(nf0);
    {
        std::vector<std::vector<int32_t>> children(nf0);
        for (int32_t f = 0; f < nf0; ++f) if (nparent[f] >= 0) children[nparent[f]].push_back(f);
        std::vector<int32_t> depth(nf0, 0);
        int32_t maxdepth = 0;
        for (int32_t f = nf0 - 1; f >= 0; --f) {
            depth[f] = nparent[f] < 0 ? 0 : depth[nparent[f]] + 1;
            maxdepth = std::max(maxdepth, depth[f]);
        }
        std::vector<std::vector<int32_t>> by_depth(maxdepth + 1);
        for (int32_t f = 0; f < nf0; ++f) by_depth[depth[f]].push_back(f);
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
                if (nparent[f] < 0 && !bnd[f].empty()) {
#pragma omp atomic write
                    bad_root = 1;
                }
            }
        }
        if (bad_root) throw std::runtime_error("direct_symbolic: root front with a non-empty boundary");
    }
    lap("boundary sets");

    // ---- fronts = (node front, class) pairs that own at least one pivot, in node-front order (children before parents
    // within every class); the rows of a diagonal class become runs of 6 * leaf_size pivots without boundary -------------
    struct ClassFront { int32_t node_front, cls, run; };      // run >= 0: index into diag_runs instead of a node front
    std::vector<ClassFront> cf;
    std::vector<int32_t> parent_of;
    std::vector<std::vector<int32_t>> diag_runs;              // rows (K11 numbering) of each run
    std::vector<int32_t> order;                               // final front id -> index into cf
    if (C == 1) {
        cf.resize(nf0);
        for (int32_t f = 0; f < nf0; ++f) cf[f] = {f, 0, -1};
        parent_of = nparent;
        order.resize(nf0);
        std::iota(order.begin(), order.end(), 0);
    } else {
        std::vector<uint8_t> diag(C, 0);
        for (int c = 0; c < C; ++c) diag[c] = class_diagonal ? class_diagonal[c] : 0;
        std::vector<int32_t> cf_id((size_t)nf0 * C, -1);
        std::vector<uint8_t> owns((size_t)nf0 * C, 0);
#pragma omp parallel for schedule(static)
        for (int32_t f = 0; f < nf0; ++f)
            for (int32_t v : front_verts[f])
                for (int c = 0; c < C; ++c)
                    if (!diag[c] && vc_cnt[(size_t)v * C + c]) owns[(size_t)f * C + c] = 1;
        for (int32_t f = 0; f < nf0; ++f)
            for (int c = 0; c < C; ++c)
                if (owns[(size_t)f * C + c]) { cf_id[(size_t)f * C + c] = (int32_t)cf.size(); cf.push_back({f, c, -1}); }
        parent_of.assign(cf.size(), -1);
        for (size_t q = 0; q < cf.size(); ++q) {
            int32_t p = nparent[cf[q].node_front];
            while (p >= 0 && cf_id[(size_t)p * C + cf[q].cls] < 0) p = nparent[p];
            parent_of[q] = p < 0 ? -1 : cf_id[(size_t)p * C + cf[q].cls];
        }
        int64_t run_len = 2 * (int64_t)leaf_size;        // (B200, config 3: 16 / 48 / 144 pivots per run -> step 47.2 / 45.7 / 48.0 ms)
        if (const char *e = getenv("PN_DIAG_RUN")) run_len = std::max(1, atoi(e));
        for (int c = 0; c < C; ++c) {
            if (!diag[c]) continue;
            // rows of the class in K11 order (chunked count + ordered fill), cut into runs
            constexpr int64_t CH = 65536;
            const int64_t nch = (n + CH - 1) / CH;
            std::vector<int64_t> cnt(nch + 1, 0);
#pragma omp parallel for schedule(static)
            for (int64_t ch = 0; ch < nch; ++ch) {
                int64_t k = 0;
                for (int64_t r = ch * CH; r < std::min(n, (ch + 1) * CH); ++r) k += row_class[r] == c;
                cnt[ch + 1] = k;
            }
            for (int64_t ch = 0; ch < nch; ++ch) cnt[ch + 1] += cnt[ch];
            std::vector<int32_t> rows_c((size_t)cnt[nch]);
#pragma omp parallel for schedule(static)
            for (int64_t ch = 0; ch < nch; ++ch) {
                int64_t k = cnt[ch];
                for (int64_t r = ch * CH; r < std::min(n, (ch + 1) * CH); ++r)
                    if (row_class[r] == c) rows_c[k++] = (int32_t)r;
            }
            for (int64_t i = 0; i < (int64_t)rows_c.size(); i += run_len) {
                cf.push_back({-1, c, (int32_t)diag_runs.size()});
                parent_of.push_back(-1);
                diag_runs.emplace_back(rows_c.begin() + i, rows_c.begin() + std::min<int64_t>((int64_t)rows_c.size(), i + run_len));
            }
        }
        order = join_forest(parent_of);
    }
    const int32_t nf = (int32_t)cf.size();
    lap("class fronts");

    // pivots per front, elimination positions
    std::vector<int32_t> ns_of(nf), pos_start(nf + 1, 0);
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t q = 0; q < nf; ++q) {
        const ClassFront &F = cf[order[q]];
        int32_t ns = 0;
        if (F.run >= 0) ns = (int32_t)diag_runs[F.run].size();
        else for (int32_t v : front_verts[F.node_front]) ns += cnt_of(v, F.cls);
        ns_of[q] = ns;
    }
    for (int32_t q = 0; q < nf; ++q) pos_start[q + 1] = pos_start[q] + ns_of[q];
    if (pos_start[nf] != n) throw std::runtime_error("direct_symbolic: rows lost in the class expansion");
    out.perm_pos.resize(n);
    out.dof_front.resize(n);
    std::vector<int32_t> &vcpos = out.s_vdofpos;       // position of the first row of (vertex, class)
    vcpos.assign((size_t)nv * C, -1);
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t q = 0; q < nf; ++q) {
        const ClassFront &F = cf[order[q]];
        int32_t pos = pos_start[q];
        if (F.run >= 0) {
            for (int32_t r : diag_runs[F.run]) { out.perm_pos[r] = pos++; out.dof_front[r] = q; }
            continue;
        }
        for (int32_t v : front_verts[F.node_front]) {
            if (!cnt_of(v, F.cls)) continue;
            vcpos[(size_t)v * C + F.cls] = pos;
            for (int32_t r = vstart[v]; r < vstart[v + 1]; ++r)
                if (C == 1 || row_class[r] == F.cls) { out.perm_pos[r] = pos++; out.dof_front[r] = q; }
        }
    }
    lap("positions");

    // fronts, index lists
    std::vector<int32_t> depth(nf, 0);
    int32_t maxdepth = 0;
    for (int32_t f = nf - 1; f >= 0; --f) {
        depth[f] = parent_of[f] < 0 ? 0 : depth[parent_of[f]] + 1;
        maxdepth = std::max(maxdepth, depth[f]);
    }
    out.fronts.assign(nf, FrontSym());
    std::vector<int32_t> nb_of(nf, 0);
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t q = 0; q < nf; ++q) {
        const ClassFront &F = cf[order[q]];
        if (F.run >= 0) continue;
        int32_t nbd = 0;
        for (int32_t o : bnd[F.node_front]) nbd += cnt_of(vert_seq[o], F.cls);
        nb_of[q] = nbd;
    }
    int64_t total_idx = 0, total_rel = 0;
    for (int32_t f = 0; f < nf; ++f) {
        FrontSym &F = out.fronts[f];
        F.parent = parent_of[f];
        F.depth = depth[f];
        F.pos_start = pos_start[f];
        F.ns = ns_of[f];
        F.nb = nb_of[f];
        F.group = cf[order[f]].run < 0 ? cf[order[f]].node_front : -1 - cf[order[f]].run;
        F.idx_off = total_idx;
        F.rel_off = total_rel;
        total_idx += F.ns + F.nb;
        total_rel += F.nb;
        out.max_front = std::max<int64_t>(out.max_front, F.ns + F.nb);
        double ns = F.ns, nbv = F.nb;
        out.factor_entries += (int64_t)(F.ns + F.nb) * F.ns + (int64_t)F.ns * F.nb;
        out.factor_flops += 2.0 / 3.0 * ns * ns * ns + 2.0 * ns * ns * nbv + 2.0 * ns * nbv * nbv;
    }
    {
        // child ranks in ascending child id (children precede their parent, so one counter per parent suffices)
        std::vector<int32_t> seen(nf, 0);
        for (int32_t f = 0; f < nf; ++f) {
            const int32_t p = parent_of[f];
            if (p < 0) continue;
            out.fronts[f].child_rank = seen[p]++;
            out.max_children = std::max(out.max_children, seen[p]);
        }
    }
    lap("front sizes");
    // inverse of perm_pos for index lists in K11 numbering
    std::vector<int32_t> &pos_dof = out.s_pos_dof;
    pos_dof.resize(n);
#pragma omp parallel for schedule(static)
    for (int64_t d = 0; d < n; ++d) pos_dof[out.perm_pos[d]] = (int32_t)d;
    out.fpos.resize(total_idx);
    out.fidx.resize(total_idx);
    out.rel.resize(total_rel);
    lap("alloc lists");
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t f = 0; f < nf; ++f) {
        const FrontSym &F = out.fronts[f];
        const ClassFront &K = cf[order[f]];
        int64_t o = F.idx_off;
        for (int32_t k = 0; k < F.ns; ++k) out.fpos[o++] = F.pos_start + k;
        if (K.run < 0)
            for (int32_t vo : bnd[K.node_front]) {
                const int32_t v = vert_seq[vo], cnt = cnt_of(v, K.cls);
                const int32_t p0 = vcpos[(size_t)v * C + K.cls];
                for (int32_t k = 0; k < cnt; ++k) out.fpos[o++] = p0 + k;
            }
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
