// Symbolic analysis for the sparse direct solver (host side, no CUDA): nested-dissection ordering
// of the node graph from node coordinates, separator tree = supernodal elimination tree, front
// index lists and extend-add maps.  See pn_direct.cu for the numeric multifrontal factorisation.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

// The large index lists go to the device right after the analysis.  In the library build (PN_SYM_PINNED) their vectors
// allocate page-locked host memory (pn_host_alloc / pn_host_free, defined in pn_direct.cu with cudaMallocHost), so the
// uploads run at full PCIe speed without a staging copy; the vectors live in the solver object and keep their
// capacity, so the allocation cost is paid once.  The host-only test harness uses plain std::allocator.
#ifdef PN_SYM_PINNED
extern "C" void *pn_host_alloc(size_t bytes);
extern "C" void pn_host_free(void *p);
namespace pn {
template <typename T>
struct PinnedAllocator {
    using value_type = T;
    PinnedAllocator() = default;
    template <typename U> PinnedAllocator(const PinnedAllocator<U> &) {}
    T *allocate(size_t n) { return static_cast<T *>(pn_host_alloc(n * sizeof(T))); }
    void deallocate(T *p, size_t) { pn_host_free(p); }
    template <typename U> bool operator==(const PinnedAllocator<U> &) const { return true; }
    template <typename U> bool operator!=(const PinnedAllocator<U> &) const { return false; }
};
template <typename T> using SymVec = std::vector<T, PinnedAllocator<T>>;
}  // namespace pn
#else
namespace pn {
template <typename T> using SymVec = std::vector<T>;
}
#endif

namespace pn {

struct FrontSym {
    int32_t ns = 0;          // pivot DOFs eliminated in this front
    int32_t nb = 0;          // boundary DOFs (rows/cols of the update matrix)
    int32_t parent = -1;     // parent front or -1 for a root
    int32_t depth = 0;       // root = 0
    int32_t child_rank = 0;  // 0, 1, ... among the children of `parent`
    int32_t pos_start = 0;   // elimination position of the first pivot DOF
    int64_t idx_off = 0;     // offset into fidx / fpos (length ns + nb)
    int64_t rel_off = 0;     // offset into rel (length nb)
    int32_t group = 0;       // fronts of one separator / leaf of the node dissection (one per DOF-type class) share a group
};

struct DirectSym {
    int64_t n = 0;                    // order of K11
    SymVec<int32_t> perm_pos;         // [n] elimination position of each K11 row
    std::vector<int32_t> dof_front;   // [n] front that eliminates each K11 row
    std::vector<FrontSym> fronts;     // post-order: children before parents
    SymVec<int32_t> fidx;             // front index lists as K11 row numbers
    SymVec<int32_t> fpos;             // the same lists as elimination positions (ascending per front)
    SymVec<int32_t> rel;              // for each boundary DOF: its position in the parent's index list
    std::vector<int32_t> level_ptr;   // level l holds fronts level_fronts[level_ptr[l] .. level_ptr[l+1])
    std::vector<int32_t> level_fronts;   // level 0 = deepest; every child is exactly one level below its parent
    int32_t max_children = 0;
    int64_t factor_entries = 0;       // stored entries of L and U
    double factor_flops = 0;
    int64_t max_front = 0;            // largest ns + nb
    // scratch of direct_symbolic, kept between calls: re-analysing a model of the same size then touches no fresh
    // pages (the page faults of ~100 MB of new vectors were a third of the ordering time)
    std::vector<int32_t> s_pos_dof, s_vfront, s_vorder, s_vert_seq, s_f_vbeg, s_vdofpos, s_region, s_all;
};

// scratch of build_vertex_graph, kept by the caller between calls for the same reason
struct VertexGraphScratch {
    std::vector<int32_t> vid, raw;
    std::vector<int64_t> deg, ptr, fill, cnt;
};

// DOF-type classes.  The reference stores no exact zeros (`FEModel3D.py:130`), so K11 of a flat shell model in an
// axis plane is reducible: membrane translations, the bending triple (normal translation + the two in-plane
// rotations) and the drilling rotation never meet in a stored entry.  A graph-based ordering of the scalar matrix
// (COLAMD inside the reference's `spsolve`, `Analysis.py:217`) sees that by itself; a node-based dissection does not,
// and factorises fronts that are block diagonal as dense ones (6^3 = 216 against 3^3 + 2^3 = 35 per node triple).  `of_type[k]` is the class of global
// DOF type k (DX DY DZ RX RY RZ); every separator / leaf of the node dissection then yields one front per class.
struct DofClasses {
    int n = 1;
    uint8_t of_type[6] = {0, 0, 0, 0, 0, 0};
    // 1: the class holds only drilling rotations (`Quad3D.py:617`, `Plate3D.py:307`: a weak spring on the diagonal),
    // which couple to nothing -- its rows are cut into leaves without boundary
    uint8_t diagonal[6] = {0, 0, 0, 0, 0, 0};
};

struct ConnView;
// Classes predicted from the element list alone (so the ordering can start before anything is assembled): shells that
// lie in a plane normal to global axis g couple {the two in-plane translations}, {translation g + the two in-plane
// rotations} and {rotation g} separately, whatever their in-plane orientation; a shell in general position, or any
// line element, couples everything (one class).  The prediction is verified against the assembled pattern on the
// device (k_entry_maps flags an entry that joins two classes) and the analysis is redone with one class if it fails.
DofClasses predict_dof_classes(const double *xyz, const std::vector<ConnView> &shells, int64_t n_line_elements);

// Vertex graph: vertex v (a node with free DOFs) owns K11 rows vstart[v] .. vstart[v+1]; coords[3 v ..]; adjacency in
// CSR (no self loops needed; symmetric).  leaf_size = vertices per leaf region.
// With `row_class` (class of every K11 row, n_classes > 1) the node graph is still dissected once, and every separator
// / leaf then yields one front per class that has pivots there: pivots and boundary of (front, class) are the rows of
// that class at the front's vertices / boundary vertices, its parent is the nearest ancestor that owns pivots of the
// class.  The rows of a diagonal class become runs of 2 * leaf_size pivots without boundary.  The resulting forest is
// joined (join_forest in the .cpp): roots of the tallest trees share the top level, a shorter tree hangs, as an extra
// child without update matrix, from a front at the depth that puts its leaves on the bottom level.  Leaves are scaled
// so that the largest class keeps about 6 * leaf_size pivots per leaf.
void direct_symbolic(int64_t n, int32_t nv, const int32_t *vstart, const double *coords, const int64_t *adj_ptr,
                     const int32_t *adj, int leaf_size, DirectSym &out, const uint8_t *row_class = nullptr,
                     int n_classes = 1, const uint8_t *class_diagonal = nullptr);

// Builds the vertex graph of the free DOFs from element connectivity.
// newidx[dof] >= 0 gives the K11 row of a free DOF.  `conn` lists for each element kind are
// (count, nodes per element, node ids, optional active flags).
struct ConnView {
    int64_t count;
    int nn;
    const int32_t *nodes;
    const uint8_t *active;   // may be null
};
void build_vertex_graph(int64_t n_nodes, const int32_t *newidx, const double *xyz, const std::vector<ConnView> &conn,
                        std::vector<int32_t> &vstart, std::vector<double> &coords, std::vector<int64_t> &adj_ptr,
                        std::vector<int32_t> &adj, VertexGraphScratch *scratch = nullptr);

}  // namespace pn
