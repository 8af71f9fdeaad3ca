// Class-based block-gather assembly: the numeric COO->CSR step without any COO value in HBM.
//
// The reference emits 576 (quad/plate) or 144 (member/spring) triplets per element and lets scipy sum the
// duplicates (FEModel3D.py:99-151, 1773-1791, 2279).  Two observations make the numeric phase pure streaming:
//   1. elements whose stiffness inputs are bit-identical (same node-to-node difference vectors, same section /
//      material / thickness, same releases, same orientation branch) have bit-identical global matrices, and
//      generated meshes and regular frames consist of a handful of such classes.  The element kernels therefore
//      run on one representative per class; the class table (a few KB .. MB) lives in L1/L2.
//   2. every stored CSR entry (row 6a+i, col 6b+j) is the sum, in emission order, of the (i, j) entries of the
//      blocks of the <= 4 elements that contain both nodes a and b.  A per-node-pair contribution list plus one
//      16-bit code per CSR slot (pair index within the row node, column DOF) is all the map that is needed.
// Numeric phase per assembly:  k_verify_classes (reads every element's inputs: 152 B per quad, SURVEY 8d, and
// proves the class table is valid for the current coordinates / properties)  ->  element kernels on the
// representatives  ->  k_block_gather (one warp per node, rows written as contiguous runs).
// The result is bit-identical to k_scatter_add on the dense element-value buffer: same addends, same order
// (adding the exact zeros the reference filters out does not change a sum).
// When the classes do not compress (general unstructured geometry) fast_assembly_build declines and the
// caller keeps the general path (dense element values + k_scatter_add).
#include "pn_internal.h"
#include "pn_assemble_symbolic.h"

#include "elem_line.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace pn {

static inline unsigned grid_for(int64_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

constexpr int SIG_SPRING = 5, SIG_MEMBER = 12, SIG_SHELL = 14;

struct FastAssembly {
    bool ready = false;
    int64_t n_class[4] = {0, 0, 0, 0};
    int64_t table_off[4] = {0, 0, 0, 0};
    int64_t table_count = 0;
    DevBuf<double> table;
    DevBuf<int32_t> cls[4];          // class of every element
    DevBuf<uint64_t> rep_sig[4];     // signature of every class representative
    DevBuf<int32_t> rep_conn[4];     // connectivity of the representatives (node ids into the global xyz)
    DevBuf<double> rep_props[4], rep_rot;
    DevBuf<uint16_t> rep_rel;
    DevBuf<uint32_t> node_pair_ptr, pair_cptr;
    DevBuf<Contrib> contrib;
    DevBuf<uint16_t> slot_code, rowpair_code;
    DevBuf<uint32_t> node_recipe, recipe_ptr, recipe_tuples, nsp_first;
    DevBuf<uint16_t> recipe_tuples16;
    DevBuf<double> recipe_vals;
    int64_t recipe_slots = 0;
    bool idx16 = false;
    DevBuf<uint8_t> recipe_nsp;
    int width = 0;
    bool recipes = false;
    int max_pairs = 0;
    DevBuf<int> flag;
    // host-side staging kept across symbolic phases: pinned copies of the CSR pattern, and the symbolic maps (their
    // vectors keep their capacity, so re-analysing a model of the same size does not page-fault 200 MB of fresh memory)
    int32_t *h_indptr = nullptr, *h_indices = nullptr;
    int64_t h_indptr_cap = 0, h_indices_cap = 0;
    AssemblySym sym;
    // the numeric assembly (verification on the side stream, class table, recipe values, gather) as one CUDA graph:
    // four dependent launches of 5-50 us each lose ~10 us to launch gaps when issued one by one
    cudaGraphExec_t graph_exec = nullptr;
    double *graph_vals = nullptr;
    cudaStream_t graph_stream = nullptr;
    int graph_launches = 0;
    void drop_graph() {
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        graph_exec = nullptr;
    }
    ~FastAssembly() {
        drop_graph();
        if (h_indptr) cudaFreeHost(h_indptr);
        if (h_indices) cudaFreeHost(h_indices);
    }
};

// ---- signatures: every word the element kernels read, as raw bits ---------------------------------------
PN_HD uint64_t pn_bits(double v) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(v);
#else
    uint64_t u; memcpy(&u, &v, sizeof(u)); return u;
#endif
}
// orientation branch of pn_line_dircos: it tests the ABSOLUTE coordinates with a relative tolerance
PN_HD uint64_t pn_line_branch(const double *pi, const double *pj) {
    uint64_t f = 0;
    if (pn_isclose(pi[0], pj[0]) && pn_isclose(pi[2], pj[2])) f |= 1;
    if (pn_isclose(pi[1], pj[1])) f |= 2;
    if (pj[1] > pi[1]) f |= 4;
    return f;
}
PN_HD void pn_sig_spring(const double *pi, const double *pj, double ks, uint64_t *s) {
    s[0] = pn_bits(pj[0] - pi[0]); s[1] = pn_bits(pj[1] - pi[1]); s[2] = pn_bits(pj[2] - pi[2]);
    s[3] = pn_bits(ks); s[4] = pn_line_branch(pi, pj);
}
PN_HD void pn_sig_member(const double *pi, const double *pj, const double *props, double rot, uint16_t rel, uint64_t *s) {
    s[0] = pn_bits(pj[0] - pi[0]); s[1] = pn_bits(pj[1] - pi[1]); s[2] = pn_bits(pj[2] - pi[2]);
    s[3] = pn_bits(rot);
    for (int k = 0; k < 6; ++k) s[4 + k] = pn_bits(props[k]);
    s[10] = rel; s[11] = pn_line_branch(pi, pj);
}
PN_HD void pn_sig_shell(const double *p0, const double *p1, const double *p2, const double *p3, const double *props, uint64_t *s) {
    for (int k = 0; k < 3; ++k) { s[k] = pn_bits(p1[k] - p0[k]); s[3 + k] = pn_bits(p2[k] - p0[k]); s[6 + k] = pn_bits(p3[k] - p0[k]); }
    for (int k = 0; k < 5; ++k) s[9 + k] = pn_bits(props[k]);
}

// kind: 0 spring, 1 member, 2 quad, 3 plate.  Sets *flag when an element no longer matches its class.
__global__ void k_verify_classes(int kind, int64_t n, const int32_t *__restrict__ conn, const double *__restrict__ xyz,
                                 const double *__restrict__ props, const double *__restrict__ rot,
                                 const uint16_t *__restrict__ rel, const int32_t *__restrict__ cls,
                                 const uint64_t *__restrict__ rep_sig, int *__restrict__ flag) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n) return;
    uint64_t s[SIG_SHELL];
    int words;
    if (kind < 2) {
        int64_t i = conn[2 * e], j = conn[2 * e + 1];
        double pi[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, pj[3] = {xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]};
        if (kind == 0) { pn_sig_spring(pi, pj, props[e], s); words = SIG_SPRING; }
        else {
            double pr[6];
            for (int k = 0; k < 6; ++k) pr[k] = props[6 * e + k];
            pn_sig_member(pi, pj, pr, rot[e], rel[e], s);
            words = SIG_MEMBER;
        }
    } else {
        double p[4][3], pr[5];
        for (int a = 0; a < 4; ++a) {
            int64_t nd = conn[4 * e + a];
            p[a][0] = xyz[3 * nd]; p[a][1] = xyz[3 * nd + 1]; p[a][2] = xyz[3 * nd + 2];
        }
        for (int k = 0; k < 5; ++k) pr[k] = props[5 * e + k];
        pn_sig_shell(p[0], p[1], p[2], p[3], pr, s);
        words = SIG_SHELL;
    }
    const uint64_t *r = rep_sig + (int64_t)cls[e] * words;
    bool same = true;
    for (int w = 0; w < words; ++w) same = same && (s[w] == r[w]);
    if (!same) atomicOr(flag, 1);
}

// Generic (fallback) gather: one warp per node, one CSR slot per lane and iteration; used when a node has more
// than GATHER_MAX_PAIRS neighbour nodes or more than GATHER_MAX_CONTRIB contributing (element, node) incidences.
__global__ void __launch_bounds__(256)
k_block_gather_slots(int64_t n_nodes, const int32_t *__restrict__ indptr, const uint16_t *__restrict__ slot_code,
                     const uint32_t *__restrict__ node_pair_ptr, const uint32_t *__restrict__ pair_cptr,
                     const Contrib *__restrict__ contrib, const double *__restrict__ table, const double *__restrict__ nsp_val,
                     double *__restrict__ vals) {
    const int64_t a = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (a >= n_nodes) return;
    const uint32_t p0 = node_pair_ptr[a];
    int32_t rb[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) rb[i] = indptr[6 * a + i];
    for (int32_t k = rb[0] + lane; k < rb[6]; k += 32) {
        int i = 0;
#pragma unroll
        for (int q = 1; q < 6; ++q) i += (k >= rb[q]) ? 1 : 0;
        const uint32_t code = slot_code[k];
        const uint32_t p = p0 + (code >> 3);
        const int j = (int)(code & 7u);
        double acc = 0.0;
        const uint32_t c1 = pair_cptr[p + 1];
        for (uint32_t cidx = pair_cptr[p]; cidx < c1; ++cidx) {
            const Contrib ct = contrib[cidx];
            if (ct.nd == 0) {
                if ((int)ct.la == i && j == i) acc += nsp_val[ct.base];
            } else {
                acc += table[(uint64_t)ct.base + (uint64_t)((6 * ct.la + i) * ct.nd + 6 * ct.lb + j)];
            }
        }
        vals[k] = acc;
    }
}

// Main gather: node recipes.  Nodes whose slot-by-slot addend lists are identical (all interior nodes of a generated
// mesh, all like joints of a regular frame) share one recipe: for every stored entry of the node's six CSR rows --
// one contiguous run of the CSR arrays -- `width` class-table indices in emission order (padded with the index of a
// 0.0) and an optional node-support-spring addend.  One warp per node, one slot per lane and iteration: coalesced
// tuple reads (L1-resident for shared recipes), bank-parallel shared-memory table reads, coalesced stores.  Per node
// only its recipe id (4 B) and CSR row start are read from HBM; the 8 B per stored entry are the output itself.
constexpr int GATHER_WARPS = 8;
constexpr int GATHER_NODES_PER_WARP = 4;

// Step 1: the value of every recipe slot that has no support-spring addend, summed once per recipe (not once per
// node): thread per recipe slot, addends in emission order.
template <bool IDX16>
__global__ void k_recipe_values(int64_t n_slots, const void *__restrict__ tuples, int width, const double *__restrict__ table,
                                double *__restrict__ rvals) {
    int64_t slot = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    double acc = 0.0;
    for (int t = 0; t < width; ++t) {
        uint32_t idx = IDX16 ? (uint32_t)reinterpret_cast<const uint16_t *>(tuples)[slot * width + t]
                             : reinterpret_cast<const uint32_t *>(tuples)[slot * width + t];
        acc += table[idx];
    }
    rvals[slot] = acc;
}

// Step 2: one warp per node streams its recipe's slot values into the node's CSR run (coalesced 8-byte loads from the
// L1-resident recipe values, coalesced stores).  A slot with a support-spring addend is summed in full here, spring
// first, because floating-point addition is not associative and the reference's order is spring, then elements
// (FEModel3D.py:1587-1771).
template <bool IDX16, bool HAS_NSP>
__global__ void __launch_bounds__(32 * GATHER_WARPS)
k_recipe_gather(int64_t n_nodes, const int32_t *__restrict__ indptr, const uint32_t *__restrict__ node_recipe,
                const uint32_t *__restrict__ recipe_ptr, const double *__restrict__ rvals, const void *__restrict__ tuples,
                const uint8_t *__restrict__ rnsp, int width, const uint32_t *__restrict__ nsp_first,
                const double *__restrict__ nsp_val, const double *__restrict__ table, double *__restrict__ vals) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int rep = 0; rep < GATHER_NODES_PER_WARP; ++rep) {
        const int64_t a = (blockIdx.x * (int64_t)GATHER_NODES_PER_WARP + rep) * GATHER_WARPS + w;
        if (a >= n_nodes) break;
        const int32_t k0 = indptr[6 * a], k1 = indptr[6 * a + 6];
        const uint32_t r0 = recipe_ptr[node_recipe[a]];
        const uint32_t nf = HAS_NSP ? nsp_first[a] : 0u;
        for (int32_t k = k0 + lane; k < k1; k += 32) {
            const uint32_t slot = r0 + (uint32_t)(k - k0);
            double v = rvals[slot];
            if (HAS_NSP) {
                const uint32_t ns = rnsp[slot];
                if (ns != 0xffu) {
                    double acc = 0.0;
                    acc += nsp_val[nf + ns];
                    for (int t = 0; t < width; ++t) {
                        uint32_t idx = IDX16 ? (uint32_t)reinterpret_cast<const uint16_t *>(tuples)[(size_t)slot * width + t]
                                             : reinterpret_cast<const uint32_t *>(tuples)[(size_t)slot * width + t];
                        acc += table[idx];
                    }
                    v = acc;
                }
            }
            vals[k] = v;
        }
    }
}

void fast_assembly_release(Ctx *c) {
    delete c->fast;
    c->fast = nullptr;
}

void fast_assembly_invalidate(Ctx *c) {
    if (c->fast) { c->fast->ready = false; c->fast->drop_graph(); }
}

void fast_assembly_build(Ctx *c) {
    if (!c->fast) c->fast = new FastAssembly();
    FastAssembly &f = *c->fast;
    f.ready = false;
    f.drop_graph();
    const Model &m = c->model;
    cudaStream_t st = c->stream;
    const int64_t counts[4] = {m.n_springs, m.n_members, m.n_quads, m.n_plates};
    const int words[4] = {SIG_SPRING, SIG_MEMBER, SIG_SHELL, SIG_SHELL};
    const int nn[4] = {2, 2, 4, 4};
    const std::vector<int32_t> *conn[4] = {&m.h_spring_ij, &m.h_mem_ij, &m.h_quad_ijmn, &m.h_plate_ijmn};
    if (counts[0] + counts[1] + counts[2] + counts[3] == 0 && c->n_nsp == 0) return;
    if (getenv("PN_DISABLE_FAST_ASSEMBLY")) return;       // tests compare both paths bit for bit

    HostTick tick("fast_assembly");
    std::vector<int32_t> cls[4], reps[4];
    int64_t total_elems = 0, total_classes = 0;
    for (int k = 0; k < 4; ++k) {
        const int64_t n = counts[k];
        std::vector<uint64_t> sig((size_t)n * words[k]);
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < n; ++e) {
            const int32_t *cn = conn[k]->data() + nn[k] * e;
            uint64_t *s = sig.data() + e * words[k];
            const double *x = m.h_xyz.data();
            if (k == 0) pn_sig_spring(x + 3 * cn[0], x + 3 * cn[1], m.h_spring_ks[e], s);
            else if (k == 1) pn_sig_member(x + 3 * cn[0], x + 3 * cn[1], m.h_mem_props.data() + 6 * e, m.h_mem_rot[e], m.h_mem_rel[e], s);
            else pn_sig_shell(x + 3 * cn[0], x + 3 * cn[1], x + 3 * cn[2], x + 3 * cn[3],
                              (k == 2 ? m.h_quad_props : m.h_plate_props).data() + 5 * e, s);
        }
        classify_elements(n, words[k], sig.data(), cls[k], reps[k]);
        f.n_class[k] = (int64_t)reps[k].size();
        total_elems += n;
        total_classes += f.n_class[k];
        // device copies: classes, representative signatures and representative element inputs
        f.cls[k].upload(cls[k].data(), n, st);
        std::vector<uint64_t> rsig((size_t)f.n_class[k] * words[k]);
        std::vector<int32_t> rconn((size_t)f.n_class[k] * nn[k]);
        for (int64_t q = 0; q < f.n_class[k]; ++q) {
            int64_t e = reps[k][q];
            memcpy(rsig.data() + q * words[k], sig.data() + e * words[k], sizeof(uint64_t) * words[k]);
            for (int a = 0; a < nn[k]; ++a) rconn[q * nn[k] + a] = (*conn[k])[nn[k] * e + a];
        }
        f.rep_sig[k].upload(rsig.data(), (int64_t)rsig.size(), st);
        f.rep_conn[k].upload(rconn.data(), (int64_t)rconn.size(), st);
        const int pw = k == 0 ? 1 : (k == 1 ? 6 : 5);
        const std::vector<double> &hp = k == 0 ? m.h_spring_ks : (k == 1 ? m.h_mem_props : (k == 2 ? m.h_quad_props : m.h_plate_props));
        std::vector<double> rprops((size_t)f.n_class[k] * pw);
        for (int64_t q = 0; q < f.n_class[k]; ++q)
            for (int t = 0; t < pw; ++t) rprops[q * pw + t] = hp[(size_t)reps[k][q] * pw + t];
        f.rep_props[k].upload(rprops.data(), (int64_t)rprops.size(), st);
        if (k == 1) {
            std::vector<double> rrot(f.n_class[k]);
            std::vector<uint16_t> rrel(f.n_class[k]);
            for (int64_t q = 0; q < f.n_class[k]; ++q) { rrot[q] = m.h_mem_rot[reps[k][q]]; rrel[q] = m.h_mem_rel[reps[k][q]]; }
            f.rep_rot.upload(rrot.data(), f.n_class[k], st);
            f.rep_rel.upload(rrel.data(), f.n_class[k], st);
        }
        PN_CUDA(cudaStreamSynchronize(st));      // host staging vectors go out of scope
    }
    tick.lap("signatures + classes");
    // decline when the classes do not compress: the table would be as large as the dense element values
    if (total_elems > 4096 && total_classes * 4 > total_elems) return;
    const int64_t blk[4] = {144, 144, 576, 576};
    int64_t off = 0;
    for (int k = 0; k < 4; ++k) { f.table_off[k] = off; off += f.n_class[k] * blk[k]; }
    f.table_count = off;
    if (off >= (int64_t)0xffffffffu) return;
    f.table.resize(off + 1);          // + one 0.0 that pads the recipe tuples
    PN_CUDA(cudaMemsetAsync(f.table.ptr + off, 0, sizeof(double), st));

    if (f.h_indptr_cap < c->n_dof + 1) {
        if (f.h_indptr) cudaFreeHost(f.h_indptr);
        f.h_indptr_cap = c->n_dof + 1;
        PN_CUDA(cudaMallocHost((void **)&f.h_indptr, sizeof(int32_t) * f.h_indptr_cap));
    }
    if (f.h_indices_cap < std::max<int64_t>(c->nnz_csr, 1)) {
        if (f.h_indices) cudaFreeHost(f.h_indices);
        f.h_indices_cap = std::max<int64_t>(c->nnz_csr, 1);
        PN_CUDA(cudaMallocHost((void **)&f.h_indices, sizeof(int32_t) * f.h_indices_cap));
    }
    int32_t *h_indptr = f.h_indptr, *h_indices = f.h_indices;
    std::vector<int32_t> h_nsp(c->n_nsp);
    PN_CUDA(cudaMemcpyAsync(h_indptr, c->indptr.ptr, sizeof(int32_t) * (c->n_dof + 1), cudaMemcpyDeviceToHost, st));
    if (c->nnz_csr) PN_CUDA(cudaMemcpyAsync(h_indices, c->indices.ptr, sizeof(int32_t) * c->nnz_csr, cudaMemcpyDeviceToHost, st));
    if (c->n_nsp) PN_CUDA(cudaMemcpyAsync(h_nsp.data(), c->nsp_dof.ptr, sizeof(int32_t) * c->n_nsp, cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    tick.lap("pattern D2H");
    ElemView views[4];
    for (int k = 0; k < 4; ++k) {
        views[k].count = counts[k]; views[k].nn = nn[k]; views[k].conn = conn[k]->data();
        views[k].active = k == 0 ? m.h_spring_active.data() : (k == 1 ? m.h_mem_active.data() : nullptr);
        views[k].cls = cls[k].data(); views[k].table_off = f.table_off[k];
    }
    AssemblySym &sym = f.sym;
    build_assembly_sym(m.n_nodes, c->n_nsp, h_nsp.data(), views, h_indptr, h_indices, (uint32_t)off, sym);
    tick.lap("build_assembly_sym (host)");
    if (!sym.ok) return;
    f.recipes = sym.recipes_ok;
    if (!f.recipes) {        // the per-slot fallback kernel needs the pair / contribution tables on the device
        f.node_pair_ptr.upload(sym.node_pair_ptr.data(), (int64_t)sym.node_pair_ptr.size(), st);
        f.pair_cptr.upload(sym.pair_cptr.data(), (int64_t)sym.pair_cptr.size(), st);
        f.contrib.upload(sym.contrib.data(), (int64_t)sym.contrib.size(), st);
        f.slot_code.upload(sym.slot_code.data(), (int64_t)sym.slot_code.size(), st);
    }
    f.width = sym.width;
    if (f.recipes) {
        f.node_recipe.upload(sym.node_recipe.data(), (int64_t)sym.node_recipe.size(), st);
        f.recipe_ptr.upload(sym.recipe_ptr.data(), (int64_t)sym.recipe_ptr.size(), st);
        f.idx16 = off + 1 < 65536;
        if (f.idx16) {
            std::vector<uint16_t> t16(sym.recipe_tuples.size());
            for (size_t q = 0; q < t16.size(); ++q) t16[q] = (uint16_t)sym.recipe_tuples[q];
            f.recipe_tuples16.upload(t16.data(), (int64_t)t16.size(), st);
            PN_CUDA(cudaStreamSynchronize(st));
        } else {
            f.recipe_tuples.upload(sym.recipe_tuples.data(), (int64_t)sym.recipe_tuples.size(), st);
        }
        f.recipe_nsp.upload(sym.recipe_nsp.data(), (int64_t)sym.recipe_nsp.size(), st);
        f.recipe_slots = (int64_t)sym.recipe_nsp.size();
        f.recipe_vals.resize(f.recipe_slots);
        f.nsp_first.upload(sym.nsp_first.data(), (int64_t)sym.nsp_first.size(), st);
    }
    f.flag.resize(1);
    PN_CUDA(cudaStreamSynchronize(st));
    tick.lap("uploads");
    f.ready = true;
}

static void fast_assembly_enqueue(Ctx *c, double *vals) {
    FastAssembly &f = *c->fast;
    const Model &m = c->model;
    cudaStream_t st = c->stream;
    {
        // class verification runs on the side stream, concurrently with the class table + gather on the main stream
        cudaStream_t s2 = c->stream2;
        PN_CUDA(cudaEventRecord(c->ev_fork, st));
        PN_CUDA(cudaStreamWaitEvent(s2, c->ev_fork, 0));
        f.flag.zero(s2);
        if (m.n_springs) k_verify_classes<<<grid_for(m.n_springs, 128), 128, 0, s2>>>(0, m.n_springs, m.spring_ij.ptr, m.xyz.ptr, m.spring_ks.ptr, nullptr, nullptr, f.cls[0].ptr, f.rep_sig[0].ptr, f.flag.ptr);
        if (m.n_members) k_verify_classes<<<grid_for(m.n_members, 128), 128, 0, s2>>>(1, m.n_members, m.mem_ij.ptr, m.xyz.ptr, m.mem_props.ptr, m.mem_rot.ptr, m.mem_rel.ptr, f.cls[1].ptr, f.rep_sig[1].ptr, f.flag.ptr);
        if (m.n_quads) k_verify_classes<<<grid_for(m.n_quads, 128), 128, 0, s2>>>(2, m.n_quads, m.quad_ijmn.ptr, m.xyz.ptr, m.quad_props.ptr, nullptr, nullptr, f.cls[2].ptr, f.rep_sig[2].ptr, f.flag.ptr);
        if (m.n_plates) k_verify_classes<<<grid_for(m.n_plates, 128), 128, 0, s2>>>(3, m.n_plates, m.plate_ijmn.ptr, m.xyz.ptr, m.plate_props.ptr, nullptr, nullptr, f.cls[3].ptr, f.rep_sig[3].ptr, f.flag.ptr);
        c->count_launch(4);
        PN_CUDA(cudaEventRecord(c->ev_join, s2));
    }
    ElemInputs in{f.n_class[0], f.n_class[1], f.n_class[2], f.n_class[3], f.rep_conn[0].ptr, f.rep_props[0].ptr,
                  f.rep_conn[1].ptr, f.rep_props[1].ptr, f.rep_rot.ptr, f.rep_rel.ptr, f.rep_conn[2].ptr, f.rep_props[2].ptr,
                  f.rep_conn[3].ptr, f.rep_props[3].ptr, m.xyz.ptr};
    double *t = f.table.ptr;
    launch_element_ke_arrays(c, in, t + f.table_off[0], t + f.table_off[1], t + f.table_off[2], t + f.table_off[3], st);
    if (m.n_nodes) {
        ProfScope ps_(c, PS_SCATTER);
        if (f.recipes) {
            unsigned grid = grid_for(m.n_nodes, GATHER_WARPS * GATHER_NODES_PER_WARP);
            const void *tp = f.idx16 ? (const void *)f.recipe_tuples16.ptr : (const void *)f.recipe_tuples.ptr;
            if (f.idx16) k_recipe_values<true><<<grid_for(f.recipe_slots, 256), 256, 0, st>>>(f.recipe_slots, tp, f.width, f.table.ptr, f.recipe_vals.ptr);
            else k_recipe_values<false><<<grid_for(f.recipe_slots, 256), 256, 0, st>>>(f.recipe_slots, tp, f.width, f.table.ptr, f.recipe_vals.ptr);
            c->count_launch();
#define PN_RG(I16, NSP) k_recipe_gather<I16, NSP><<<grid, 32 * GATHER_WARPS, 0, st>>>(                                      \
    m.n_nodes, c->indptr.ptr, f.node_recipe.ptr, f.recipe_ptr.ptr, f.recipe_vals.ptr, tp, f.recipe_nsp.ptr, f.width, f.nsp_first.ptr, \
    c->nsp_val.ptr, f.table.ptr, vals)
            if (f.idx16 && c->n_nsp) PN_RG(true, true);
            else if (f.idx16) PN_RG(true, false);
            else if (c->n_nsp) PN_RG(false, true);
            else PN_RG(false, false);
#undef PN_RG
        }
        else
            k_block_gather_slots<<<grid_for(m.n_nodes * 32, 256), 256, 0, st>>>(m.n_nodes, c->indptr.ptr, f.slot_code.ptr,
                                                                                 f.node_pair_ptr.ptr, f.pair_cptr.ptr, f.contrib.ptr,
                                                                                 f.table.ptr, c->nsp_val.ptr, vals);
        c->count_launch();
    }
    PN_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0));
}

bool fast_assembly_run(Ctx *c, double *vals) {
    if (!c->fast || !c->fast->ready) return false;
    FastAssembly &f = *c->fast;
    cudaStream_t st = c->stream;
    if (c->opt_assembly_graph) {
        if (!f.graph_exec || f.graph_vals != vals || f.graph_stream != st) {
            f.drop_graph();
            const bool prof_was = c->prof.on;
            const int64_t launches_before = c->launches;
            c->prof.on = false;                            // events recorded inside a capture cannot be timed
            cudaGraph_t g = nullptr;
            PN_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            try {
                fast_assembly_enqueue(c, vals);
            } catch (...) {
                cudaStreamEndCapture(st, &g);
                if (g) cudaGraphDestroy(g);
                c->prof.on = prof_was;
                throw;
            }
            PN_CUDA(cudaStreamEndCapture(st, &g));
            c->prof.on = prof_was;
            f.graph_launches = (int)(c->launches - launches_before);
            c->launches = launches_before;
            PN_CUDA(cudaGraphInstantiate(&f.graph_exec, g, 0));
            cudaGraphDestroy(g);
            f.graph_vals = vals;
            f.graph_stream = st;
        }
        {
            ProfScope total_(c, PS_ASSEMBLY_TOTAL);
            PN_CUDA(cudaGraphLaunch(f.graph_exec, st));
        }
        c->count_launch(f.graph_launches);
    } else {
        ProfScope total_(c, PS_ASSEMBLY_TOTAL);
        fast_assembly_enqueue(c, vals);
    }
    int flag = 0;
    PN_CUDA(cudaMemcpyAsync(&flag, f.flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    if (flag) { f.ready = false; f.drop_graph(); return false; }      // coordinates / properties changed under a fixed class map
    return true;
}

}  // namespace pn
