// Host-side symbolic data for the class-based block-gather assembly (see pn_assemble.cu).
#pragma once
#include <cstdint>
#include <vector>

namespace pn {

struct Contrib {
    uint32_t base;   // nd != 0: offset (doubles) of the element class' dense nd x nd block in the class table
                     // nd == 0: index into the node-support-spring value list
    uint8_t la, lb;  // local node of the row / column node inside the element (nd == 0: la = DOF 0..5)
    uint8_t nd;      // 12, 24, or 0 for a node support spring
    uint8_t pad;
};

struct ElemView {
    int64_t count = 0;
    int nn = 0;                       // nodes per element
    const int32_t *conn = nullptr;
    const uint8_t *active = nullptr;  // may be null
    const int32_t *cls = nullptr;     // class of each element
    int64_t table_off = 0;            // offset (doubles) of this kind's first class block in the class table
};

struct AssemblySym {
    std::vector<uint32_t> node_pair_ptr;   // [N + 1]
    std::vector<uint32_t> pair_cptr;       // [P + 1]
    std::vector<Contrib> contrib;
    std::vector<uint16_t> slot_code;       // [nnz_csr]  (pair index within the node << 3) | column DOF
    std::vector<uint16_t> rowpair_code;    // [6 P] for node a, row i, pair p (index 6 ptr[a] + i np + p):
                                           // (offset of the pair's first stored entry within the CSR row << 6) | 6-bit column mask
    // Node recipes: nodes whose slot-by-slot addend lists are identical share one recipe.  A recipe lists, for every
    // stored entry of the node's six CSR rows (one contiguous run), the class-table indices of its addends in
    // emission order, padded to `width` with `zero_index` (a 0.0 kept at the end of the class table), plus the local
    // index of a node-support-spring addend (0xff = none).
    std::vector<uint32_t> node_recipe;     // [N]
    std::vector<uint32_t> recipe_ptr;      // [R + 1] first slot of each recipe in the tuple arrays
    std::vector<uint32_t> recipe_tuples;   // [slots][width]
    std::vector<uint8_t> recipe_nsp;       // [slots]
    std::vector<uint32_t> nsp_first;       // [N] index of the node's first support spring in the nsp list
    int width = 0;
    bool recipes_ok = false;
    int max_pairs_per_node = 0;
    bool ok = false;                       // false: some stored entry has no contributing block (never expected)
};

// Equivalence classes of elements with bit-identical stiffness inputs.  `sig` is [count][words] of raw 64-bit
// words; returns class ids and the representative (first element) of each class.
void classify_elements(int64_t count, int words, const uint64_t *sig, std::vector<int32_t> &cls, std::vector<int32_t> &reps);

// nsp_dof: global DOF of every active node support spring in emission order.  views: springs, members, quads,
// plates (emission order).  indptr/indices: CSR pattern of Ke.
void build_assembly_sym(int64_t n_nodes, int64_t n_nsp, const int32_t *nsp_dof, const ElemView views[4], const int32_t *indptr,
                        const int32_t *indices, uint32_t zero_index, AssemblySym &out);

}  // namespace pn
