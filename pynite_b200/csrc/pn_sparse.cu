// Sparse machinery: symbolic phase (COO keys in emission order -> sorted unique CSR pattern and the
// scatter map), deterministic COO->CSR scatter-add, partition into K11/K12/K21/K22, DOF incidence,
// load vectors, right-hand sides, reactions, SpMM.
//
// Device-wide sort / scan use CUB (shipped with the CUDA toolkit) in the *symbolic* phase only, which
// runs once per topology; the numeric kernels that run per assembly / per iteration are written here.
#include <cub/cub.cuh>

#include <algorithm>
#include <climits>

#include "pn_internal.h"

namespace pn {

static inline unsigned grid_for(int64_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

// ------------------------------------------------------------------------------------------------
// layout helpers: (element kind, element, local row, local col) of an element-value position
// ------------------------------------------------------------------------------------------------
struct EvLayout {
    int64_t n_nsp, off_spring, off_member, off_quad, off_plate, count;
    const int32_t *nsp_dof, *spring_ij, *mem_ij, *quad_ijmn, *plate_ijmn;
    const uint8_t *spring_active, *mem_active;
};

__device__ __forceinline__ void ev_row_col(const EvLayout &L, int64_t p, int32_t &row, int32_t &col, int &kind,
                                           int64_t &elem) {
    if (p < L.n_nsp) {
        row = col = L.nsp_dof[p];
        kind = 0; elem = p;
        return;
    }
    const int32_t *conn;
    int nd, nn;
    int64_t q;
    if (p < L.off_member) { q = p - L.off_spring; conn = L.spring_ij; nd = 12; nn = 2; kind = PN_SPRING; }
    else if (p < L.off_quad) { q = p - L.off_member; conn = L.mem_ij; nd = 12; nn = 2; kind = PN_MEMBER; }
    else if (p < L.off_plate) { q = p - L.off_quad; conn = L.quad_ijmn; nd = 24; nn = 4; kind = PN_QUAD; }
    else { q = p - L.off_plate; conn = L.plate_ijmn; nd = 24; nn = 4; kind = PN_PLATE; }
    int64_t e = q / (nd * nd);
    int rem = (int)(q - e * (nd * nd));
    int r = rem / nd, cc = rem - r * nd;
    row = conn[nn * e + r / 6] * 6 + r % 6;
    col = conn[nn * e + cc / 6] * 6 + cc % 6;
    elem = e;
}

static EvLayout make_layout(Ctx *c) {
    EvLayout L;
    L.n_nsp = c->n_nsp; L.off_spring = c->ev_off_spring; L.off_member = c->ev_off_member;
    L.off_quad = c->ev_off_quad; L.off_plate = c->ev_off_plate; L.count = c->ev_count;
    L.nsp_dof = c->nsp_dof.ptr; L.spring_ij = c->model.spring_ij.ptr; L.mem_ij = c->model.mem_ij.ptr;
    L.quad_ijmn = c->model.quad_ijmn.ptr; L.plate_ijmn = c->model.plate_ijmn.ptr;
    L.spring_active = c->model.spring_active.ptr; L.mem_active = c->model.mem_active.ptr;
    return L;
}

// FEModel3D._append_sparse_block :130 keeps an entry iff value != 0.0; node support springs are
// appended unconditionally (:1587-1690); inactive springs / members are skipped (:1696, :1717).
__global__ void k_entry_flags(EvLayout L, const double *__restrict__ ev, int dense_members, uint8_t *__restrict__ flag) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= L.count) return;
    uint8_t f;
    if (p < L.n_nsp) f = 1;
    else if (p < L.off_member) f = L.spring_active[(p - L.off_spring) / 144] && ev[p] != 0.0;
    else if (p < L.off_quad) f = L.mem_active[(p - L.off_member) / 144] && (dense_members || ev[p] != 0.0);
    else f = ev[p] != 0.0;
    flag[p] = f;
}

__global__ void k_emit_coo(EvLayout L, const uint8_t *__restrict__ flag, const uint32_t *__restrict__ pos,
                           uint64_t *__restrict__ key, uint32_t *__restrict__ src) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= L.count || !flag[p]) return;
    int32_t row, col; int kind; int64_t e;
    ev_row_col(L, p, row, col, kind, e);
    uint32_t cpos = pos[p];
    key[cpos] = ((uint64_t)(uint32_t)row << 32) | (uint32_t)col;
    src[cpos] = (uint32_t)p;
}

__global__ void k_heads(int64_t n, const uint64_t *__restrict__ key, uint32_t *__restrict__ head) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    head[k] = (k == 0 || key[k] != key[k - 1]) ? 1u : 0u;
}

__global__ void k_fill_pattern(int64_t n, const uint64_t *__restrict__ key, const uint32_t *__restrict__ head,
                               const uint32_t *__restrict__ incl, int32_t *__restrict__ indices,
                               int32_t *__restrict__ slot_row, uint32_t *__restrict__ seg_start) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n || !head[k]) return;
    uint32_t s = incl[k] - 1;
    indices[s] = (int32_t)(key[k] & 0xffffffffu);
    slot_row[s] = (int32_t)(key[k] >> 32);
    seg_start[s] = (uint32_t)k;
}

// indptr[r] = first position in sorted `vals` (length n) with vals[pos] >= r, for r = 0..rows
__global__ void k_lower_bounds(int64_t rows, int64_t n, const int32_t *__restrict__ vals, int32_t *__restrict__ indptr) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r > rows) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (vals[mid] < r) lo = mid + 1; else hi = mid;
    }
    indptr[r] = (int32_t)lo;
}
__global__ void k_lower_bounds_u32(int64_t rows, int64_t n, const uint32_t *__restrict__ vals, uint32_t *__restrict__ out) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r > rows) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)vals[mid] < r) lo = mid + 1; else hi = mid;
    }
    out[r] = (uint32_t)lo;
}

template <typename T>
static T read_back(const T *dev, cudaStream_t st) {
    T h;
    PN_CUDA(cudaMemcpyAsync(&h, dev, sizeof(T), cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    return h;
}

static void exclusive_scan_u8(Ctx *c, const uint8_t *in, uint32_t *out, int64_t n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, c->stream);
    c->cub_tmp.resize((int64_t)bytes);
    PN_CUDA(cub::DeviceScan::ExclusiveSum(c->cub_tmp.ptr, bytes, in, out, (int)n, c->stream));
    c->count_launch(2);
}
static void inclusive_scan_u32(Ctx *c, const uint32_t *in, uint32_t *out, int64_t n) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, (int)n, c->stream);
    c->cub_tmp.resize((int64_t)bytes);
    PN_CUDA(cub::DeviceScan::InclusiveSum(c->cub_tmp.ptr, bytes, in, out, (int)n, c->stream));
    c->count_launch(2);
}
static void exclusive_scan_i32(Ctx *c, const int32_t *in, int32_t *out, int64_t n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, c->stream);
    c->cub_tmp.resize((int64_t)bytes);
    PN_CUDA(cub::DeviceScan::ExclusiveSum(c->cub_tmp.ptr, bytes, in, out, (int)n, c->stream));
    c->count_launch(2);
}

static int bits_for(int64_t v) {
    int b = 1;
    while ((1ll << b) <= v && b < 62) ++b;
    return b;
}

// ------------------------------------------------------------------------------------------------
// partition (Analysis._partition_D :1412-1505, _partition :1508-1540)
// ------------------------------------------------------------------------------------------------
__global__ void k_known_flags(int64_t n_dof, const uint8_t *__restrict__ support, const uint8_t *__restrict__ enf,
                              uint8_t *__restrict__ known) {
    int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (d >= n_dof) return;
    int64_t nd = d / 6; int k = (int)(d % 6);
    known[d] = (((support[nd] | enf[nd]) >> k) & 1u);
}
__global__ void k_partition_lists(int64_t n_dof, const uint8_t *__restrict__ known, const uint32_t *__restrict__ kpos,
                                  const uint8_t *__restrict__ enf, const double *__restrict__ enf_val,
                                  int32_t *__restrict__ d1, int32_t *__restrict__ d2, double *__restrict__ d2_val,
                                  int32_t *__restrict__ newidx) {
    int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (d >= n_dof) return;
    uint32_t kp = kpos[d];
    if (known[d]) {
        d2[kp] = (int32_t)d;
        int64_t nd = d / 6; int k = (int)(d % 6);
        d2_val[kp] = ((enf[nd] >> k) & 1u) ? enf_val[d] : 0.0;
        newidx[d] = -1 - (int32_t)kp;
    } else {
        int32_t fp = (int32_t)(d - kp);
        d1[fp] = (int32_t)d;
        newidx[d] = fp;
    }
}
// want_known: 0 -> keep columns in D1, 1 -> keep columns in D2
__global__ void k_block_count(int64_t rows, const int32_t *__restrict__ rowlist, const int32_t *__restrict__ indptr,
                              const int32_t *__restrict__ indices, const int32_t *__restrict__ newidx, int want_known,
                              int32_t *__restrict__ cnt) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= rows) return;
    int32_t r = rowlist[i], n = 0;
    for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) n += ((newidx[indices[k]] < 0) == (want_known != 0));
    cnt[i] = n;
}
__global__ void k_block_fill(int64_t rows, const int32_t *__restrict__ rowlist, const int32_t *__restrict__ indptr,
                             const int32_t *__restrict__ indices, const int32_t *__restrict__ newidx, int want_known,
                             const int32_t *__restrict__ bptr, int32_t *__restrict__ bind, uint32_t *__restrict__ bsrc) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= rows) return;
    int32_t r = rowlist[i], o = bptr[i];
    for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) {
        int32_t ni = newidx[indices[k]];
        if ((ni < 0) == (want_known != 0)) {
            bind[o] = ni < 0 ? -1 - ni : ni;
            bsrc[o] = (uint32_t)k;
            ++o;
        }
    }
}
__global__ void k_gather_u32(int64_t n, const uint32_t *__restrict__ src, const double *__restrict__ in, double *__restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[src[i]];
}

static void build_partition(Ctx *c) {
    c->row_groups_ready = false;
    cudaStream_t st = c->stream;
    int64_t n = c->n_dof;
    DevBuf<uint8_t> &known = c->sym_known;
    DevBuf<uint32_t> &kpos = c->sym_kpos;
    known.resize(n); kpos.resize(n);
    k_known_flags<<<grid_for(n, 256), 256, 0, st>>>(n, c->model.support.ptr, c->model.enf_mask.ptr, known.ptr);
    c->count_launch();
    exclusive_scan_u8(c, known.ptr, kpos.ptr, n);
    int64_t n2 = 0;
    if (n > 0) n2 = (int64_t)read_back(kpos.ptr + n - 1, st) + read_back(known.ptr + n - 1, st);
    c->n2 = n2; c->n1 = n - n2;
    c->d1.resize(c->n1); c->d2.resize(c->n2); c->d2_val.resize(c->n2); c->newidx.resize(n);
    k_partition_lists<<<grid_for(n, 256), 256, 0, st>>>(n, known.ptr, kpos.ptr, c->model.enf_mask.ptr,
                                                         c->model.enf_val.ptr, c->d1.ptr, c->d2.ptr, c->d2_val.ptr,
                                                         c->newidx.ptr);
    c->count_launch();
    for (int b = 0; b < 4; ++b) {
        CsrBlock &B = c->blk[b];
        const DevBuf<int32_t> &rows = (b < 2) ? c->d1 : c->d2;
        int want_known = b & 1;
        B.rows = rows.n; B.cols = want_known ? c->n2 : c->n1;
        DevBuf<int32_t> &cnt = c->sym_cnt;
        cnt.resize(B.rows + 1);
        cnt.zero(st);
        B.indptr.resize(B.rows + 1);
        if (B.rows) {
            k_block_count<<<grid_for(B.rows, 128), 128, 0, st>>>(B.rows, rows.ptr, c->indptr.ptr, c->indices.ptr,
                                                                  c->newidx.ptr, want_known, cnt.ptr);
            c->count_launch();
        }
        exclusive_scan_i32(c, cnt.ptr, B.indptr.ptr, B.rows + 1);
        B.nnz = read_back(B.indptr.ptr + B.rows, st);
        B.indices.resize(B.nnz); B.src.resize(B.nnz); B.vals.resize(B.nnz);
        if (B.rows && B.nnz) {
            k_block_fill<<<grid_for(B.rows, 128), 128, 0, st>>>(B.rows, rows.ptr, c->indptr.ptr, c->indices.ptr,
                                                                 c->newidx.ptr, want_known, B.indptr.ptr, B.indices.ptr,
                                                                 B.src.ptr);
            c->count_launch();
        }
    }
    PN_CUDA(cudaStreamSynchronize(st));
}

void gather_block_values(Ctx *c, const double *vals_dev) {
    for (int b = 0; b < 4; ++b) {
        CsrBlock &B = c->blk[b];
        if (B.nnz) {
            { ProfScope ps_(c, PS_GATHER_BLOCKS); k_gather_u32<<<grid_for(B.nnz, 256), 256, 0, c->stream>>>(B.nnz, B.src.ptr, vals_dev, B.vals.ptr); }
            c->count_launch();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// symbolic phase
// ------------------------------------------------------------------------------------------------
void build_symbolic(Ctx *c, uint32_t flags) {
    cudaStream_t st = c->stream;
    HostTick tick("build_symbolic");
    c->sym_flags = flags;
    c->n_dof = 6 * c->model.n_nodes;
    EvLayout L = make_layout(c);
    int64_t N = c->ev_count;
    if (N >= (int64_t)0xffffffffu) throw ArgError("model too large for 32-bit element-value positions");

    DevBuf<uint8_t> &flag = c->sym_flag;
    DevBuf<uint32_t> &pos = c->sym_pos;
    flag.resize(N); pos.resize(N);
    int64_t nnz_coo = 0;
    if (N) {
        k_entry_flags<<<grid_for(N, 256), 256, 0, st>>>(L, c->ev.ptr, (flags & PN_SYM_DENSE_MEMBER_BLOCKS) ? 1 : 0, flag.ptr);
        c->count_launch();
        exclusive_scan_u8(c, flag.ptr, pos.ptr, N);
        nnz_coo = (int64_t)read_back(pos.ptr + N - 1, st) + read_back(flag.ptr + N - 1, st);
    }
    tick.lap("entry flags + scan");
    c->nnz_coo = nnz_coo;

    DevBuf<uint64_t> &key_in = c->sym_key;
    key_in.resize(nnz_coo);
    c->coo_emit.resize(nnz_coo);
    c->coo_key.resize(nnz_coo);
    c->coo_src.resize(nnz_coo);
    if (nnz_coo) {
        k_emit_coo<<<grid_for(N, 256), 256, 0, st>>>(L, flag.ptr, pos.ptr, key_in.ptr, c->coo_emit.ptr);
        c->count_launch();
        tick.lap("emit coo");
        // stable LSD radix sort: equal keys keep emission order, which fixes the summation order
        int end_bit = 32 + bits_for(c->n_dof);
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_in.ptr, c->coo_key.ptr, c->coo_emit.ptr, c->coo_src.ptr,
                                        (int)nnz_coo, 0, end_bit, st);
        c->cub_tmp.resize((int64_t)bytes);
        PN_CUDA(cub::DeviceRadixSort::SortPairs(c->cub_tmp.ptr, bytes, key_in.ptr, c->coo_key.ptr, c->coo_emit.ptr,
                                                c->coo_src.ptr, (int)nnz_coo, 0, end_bit, st));
        c->count_launch(8);
    }
    tick.lap("flags + emit + sort");

    DevBuf<uint32_t> &head = c->sym_head, &incl = c->sym_incl;
    DevBuf<int32_t> &slot_row = c->sym_slot_row;
    head.resize(nnz_coo); incl.resize(nnz_coo);
    int64_t nnz_csr = 0;
    if (nnz_coo) {
        k_heads<<<grid_for(nnz_coo, 256), 256, 0, st>>>(nnz_coo, c->coo_key.ptr, head.ptr);
        c->count_launch();
        inclusive_scan_u32(c, head.ptr, incl.ptr, nnz_coo);
        nnz_csr = read_back(incl.ptr + nnz_coo - 1, st);
    }
    c->nnz_csr = nnz_csr;
    c->indices.resize(nnz_csr); c->seg_start.resize(nnz_csr + 1); c->vals.resize(nnz_csr);
    slot_row.resize(nnz_csr);
    c->indptr.resize(c->n_dof + 1);
    if (nnz_coo) {
        k_fill_pattern<<<grid_for(nnz_coo, 256), 256, 0, st>>>(nnz_coo, c->coo_key.ptr, head.ptr, incl.ptr,
                                                               c->indices.ptr, slot_row.ptr, c->seg_start.ptr);
        c->count_launch();
    }
    uint32_t last = (uint32_t)nnz_coo;
    PN_CUDA(cudaMemcpyAsync(c->seg_start.ptr + nnz_csr, &last, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    k_lower_bounds<<<grid_for(c->n_dof + 1, 256), 256, 0, st>>>(c->n_dof, nnz_csr, slot_row.ptr, c->indptr.ptr);
    c->count_launch();
    PN_CUDA(cudaStreamSynchronize(st));

    tick.lap("pattern");
    build_partition(c);
    tick.lap("partition");
    c->symbolic_ready = true;
    c->vals_ready = false;
    c->incidence_ready = false;
}

// COO triplets in emission order (FEModel3D.py:1773-1791), for inspection / parity tests.
__global__ void k_export_coo(EvLayout L, int64_t n, const uint32_t *__restrict__ emit, const double *__restrict__ ev,
                             int32_t *__restrict__ row, int32_t *__restrict__ col, double *__restrict__ data) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    int32_t r, cc; int kind; int64_t e;
    ev_row_col(L, emit[k], r, cc, kind, e);
    row[k] = r; col[k] = cc; data[k] = ev[emit[k]];
}
void export_coo(Ctx *c, int32_t *row_dev, int32_t *col_dev, double *data_dev) {
    if (!c->nnz_coo) return;
    k_export_coo<<<grid_for(c->nnz_coo, 256), 256, 0, c->stream>>>(make_layout(c), c->nnz_coo, c->coo_emit.ptr, c->ev.ptr,
                                                                    row_dev, col_dev, data_dev);
    c->count_launch();
}

// ------------------------------------------------------------------------------------------------
// numeric COO -> CSR: each stored entry is the sum of its COO duplicates taken in emission order
// (fixed at symbolic time by the stable sort), so the result is bit-reproducible run to run.
// A warp walks 32 consecutive CSR slots per step so the 8-byte stores are fully coalesced.
// ------------------------------------------------------------------------------------------------
__global__ void k_scatter_add(int64_t nnz_csr, const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ src,
                              const double *__restrict__ ev, double *__restrict__ vals) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nnz_csr) return;
    uint32_t a = seg_start[s], b = seg_start[s + 1];
    double acc = ev[src[a]];
    for (uint32_t k = a + 1; k < b; ++k) acc += ev[src[k]];
    vals[s] = acc;
}

void scatter_add_values(Ctx *c, const double *ev_dev, double *vals_dev) {
    if (!c->nnz_csr) return;
    { ProfScope ps_(c, PS_SCATTER); k_scatter_add<<<grid_for(c->nnz_csr, 256), 256, 0, c->stream>>>(c->nnz_csr, c->seg_start.ptr, c->coo_src.ptr, ev_dev,
                                                                     vals_dev); }
    c->count_launch();
}

// Analysis._check_stability :111-168: a zero (or absent) diagonal on an unsupported DOF.
__global__ void k_zero_diag(int64_t n_dof, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                            const double *__restrict__ vals, const uint8_t *__restrict__ support,
                            int32_t *__restrict__ bad, int cap, int *__restrict__ n_bad) {
    int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (d >= n_dof) return;
    if ((support[d / 6] >> (d % 6)) & 1u) return;
    double diag = 0.0;
    for (int32_t k = indptr[d]; k < indptr[d + 1]; ++k)
        if (indices[k] == (int32_t)d) diag = vals[k];
    if (diag == 0.0) {       // math.isclose(x, 0) with abs_tol = 0 is an exact-zero test
        int slot = atomicAdd(n_bad, 1);
        if (slot < cap) bad[slot] = (int32_t)d;
    }
}

int64_t check_nodal_stability(Ctx *c, int32_t *bad_host, int64_t cap) {
    DevBuf<int32_t> &bad = c->sym_bad;
    DevBuf<int> &nb = c->sym_nb;
    bad.resize(cap); nb.resize(1);
    nb.zero(c->stream);
    k_zero_diag<<<grid_for(c->n_dof, 256), 256, 0, c->stream>>>(c->n_dof, c->indptr.ptr, c->indices.ptr, c->vals.ptr,
                                                                 c->model.support.ptr, bad.ptr, (int)cap, nb.ptr);
    c->count_launch();
    int n = read_back(nb.ptr, c->stream);
    int64_t m = n < cap ? n : cap;
    if (m) {
        PN_CUDA(cudaMemcpy(bad_host, bad.ptr, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost));
        // the reference reports in DOF order
        std::vector<int32_t> tmp(bad_host, bad_host + m);
        std::sort(tmp.begin(), tmp.end());
        for (int64_t i = 0; i < m; ++i) bad_host[i] = tmp[i];
    }
    return n;
}

// ------------------------------------------------------------------------------------------------
// DOF incidence of element end-force entries (used to gather FER vectors and reactions without
// atomics: entries of one DOF are summed in element emission order)
// ------------------------------------------------------------------------------------------------
struct EfLayout {
    int64_t off_spring, off_member, off_quad, off_plate, count;
    const int32_t *spring_ij, *mem_ij, *quad_ijmn, *plate_ijmn;
};
__device__ __forceinline__ int32_t ef_dof(const EfLayout &L, int64_t q) {
    const int32_t *conn; int nd, nn; int64_t r;
    if (q < L.off_member) { r = q - L.off_spring; conn = L.spring_ij; nd = 12; nn = 2; }
    else if (q < L.off_quad) { r = q - L.off_member; conn = L.mem_ij; nd = 12; nn = 2; }
    else if (q < L.off_plate) { r = q - L.off_quad; conn = L.quad_ijmn; nd = 24; nn = 4; }
    else { r = q - L.off_plate; conn = L.plate_ijmn; nd = 24; nn = 4; }
    int64_t e = r / nd; int k = (int)(r - e * nd);
    return conn[nn * e + k / 6] * 6 + k % 6;
}
__global__ void k_ef_keys(EfLayout L, uint32_t *__restrict__ key, uint32_t *__restrict__ val) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= L.count) return;
    key[q] = (uint32_t)ef_dof(L, q);
    val[q] = (uint32_t)q;
}

void build_incidence(Ctx *c) {
    if (c->incidence_ready) return;
    cudaStream_t st = c->stream;
    const Model &m = c->model;
    c->ef_off_spring = 0;
    c->ef_off_member = 12 * m.n_springs;
    c->ef_off_quad = c->ef_off_member + 12 * m.n_members;
    c->ef_off_plate = c->ef_off_quad + 24 * m.n_quads;
    c->ef_count = c->ef_off_plate + 24 * m.n_plates;
    if (c->ef_count >= (int64_t)0xffffffffu) throw ArgError("model too large for 32-bit element-force positions");
    EfLayout L{c->ef_off_spring, c->ef_off_member, c->ef_off_quad, c->ef_off_plate, c->ef_count,
               m.spring_ij.ptr, m.mem_ij.ptr, m.quad_ijmn.ptr, m.plate_ijmn.ptr};
    int64_t n = c->ef_count;
    DevBuf<uint32_t> &key_in = c->sym_k32a, &key_out = c->sym_k32b, &val_in = c->sym_k32c;
    key_in.resize(n); key_out.resize(n); val_in.resize(n);
    c->inc_src.resize(n);
    c->inc_start.resize(c->n_dof + 1);
    if (n) {
        k_ef_keys<<<grid_for(n, 256), 256, 0, st>>>(L, key_in.ptr, val_in.ptr);
        c->count_launch();
        size_t bytes = 0;
        int end_bit = bits_for(c->n_dof);
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_in.ptr, key_out.ptr, val_in.ptr, c->inc_src.ptr, (int)n, 0,
                                        end_bit, st);
        c->cub_tmp.resize((int64_t)bytes);
        PN_CUDA(cub::DeviceRadixSort::SortPairs(c->cub_tmp.ptr, bytes, key_in.ptr, key_out.ptr, val_in.ptr,
                                                c->inc_src.ptr, (int)n, 0, end_bit, st));
        c->count_launch(6);
    }
    k_lower_bounds_u32<<<grid_for(c->n_dof + 1, 256), 256, 0, st>>>(c->n_dof, n, key_out.ptr, c->inc_start.ptr);
    c->count_launch();
    PN_CUDA(cudaStreamSynchronize(st));
    c->ef.resize(c->ef_count);
    c->ef.zero(st);
    // elements that touch a supported node, per kind (host scan of the connectivity, once per topology)
    {
        struct View { int64_t n; int nn; const std::vector<int32_t> *conn; };
        View views[4] = {{m.n_springs, 2, &m.h_spring_ij}, {m.n_members, 2, &m.h_mem_ij},
                         {m.n_quads, 4, &m.h_quad_ijmn}, {m.n_plates, 4, &m.h_plate_ijmn}};
        for (int k = 0; k < 4; ++k) {
            std::vector<int32_t> list;
            const View &v = views[k];
            for (int64_t e = 0; e < v.n; ++e) {
                bool sup = false;
                for (int a = 0; a < v.nn; ++a) sup = sup || m.h_support[(*v.conn)[v.nn * e + a]] != 0;
                if (sup) list.push_back((int32_t)e);
            }
            c->sup_list[k].upload(list.data(), (int64_t)list.size(), st);
            PN_CUDA(cudaStreamSynchronize(st));
        }
    }
    c->incidence_ready = true;
}

// ------------------------------------------------------------------------------------------------
// load vectors (FEModel3D.P :2184-2235, FEModel3D.FER :2136-2182), per load case with unit factor
// ------------------------------------------------------------------------------------------------
__global__ void k_scatter_triplets(int64_t n, const int32_t *__restrict__ idx, const int32_t *__restrict__ cs,
                                   const double *__restrict__ val, int64_t stride, double *__restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[(int64_t)cs[i] * stride + idx[i]] = val[i];
}
__global__ void k_scatter_member_fer(int64_t n_groups, const int32_t *__restrict__ g_member, const int32_t *__restrict__ g_case,
                                     const double *__restrict__ gfer, int64_t ef_count, int64_t ef_off_member,
                                     double *__restrict__ fer_case) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_groups * 12) return;
    int64_t g = t / 12; int k = (int)(t % 12);
    fer_case[(int64_t)g_case[g] * ef_count + ef_off_member + (int64_t)g_member[g] * 12 + k] = gfer[t];
}
// G[case][dof] = P[case][dof] - sum over incident element entries of fer_case[case][.]
__global__ void k_case_vectors(int64_t n_dof, int n_case, const uint32_t *__restrict__ inc_start,
                               const uint32_t *__restrict__ inc_src, const double *__restrict__ Pcase,
                               const double *__restrict__ fer_case, int64_t ef_count, double *__restrict__ G) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_dof * n_case) return;
    int cs = (int)(t / n_dof);
    int64_t d = t - (int64_t)cs * n_dof;
    const double *f = fer_case + (int64_t)cs * ef_count;
    double acc = 0.0;
    for (uint32_t k = inc_start[d]; k < inc_start[d + 1]; ++k) acc += f[inc_src[k]];
    G[t] = Pcase[t] - acc;
}

void build_load_vectors(Ctx *c) {
    cudaStream_t st = c->stream;
    Loads &ld = c->loads;
    const Model &m = c->model;
    HostTick tick("load vectors");
    build_incidence(c);
    tick.lap("incidence");
    int64_t n_dof = c->n_dof;
    int nc = ld.n_case;
    ld.Pcase.resize((int64_t)nc * n_dof);
    ld.Gcase.resize((int64_t)nc * n_dof);
    ld.fer_case.resize((int64_t)nc * c->ef_count);
    ld.Pcase.zero(st);
    ld.fer_case.zero(st);
    if (ld.nl_val.n) {
        k_scatter_triplets<<<grid_for(ld.nl_val.n, 256), 256, 0, st>>>(ld.nl_val.n, ld.nl_dof.ptr, ld.nl_case.ptr,
                                                                        ld.nl_val.ptr, n_dof, ld.Pcase.ptr);
        c->count_launch();
    }
    if (ld.n_mgroups) {
        launch_member_fer_groups(c, st);
        k_scatter_member_fer<<<grid_for(ld.n_mgroups * 12, 256), 256, 0, st>>>(
            ld.n_mgroups, ld.mg_member.ptr, ld.mg_case.ptr, ld.mg_fer_global.ptr, c->ef_count, c->ef_off_member,
            ld.fer_case.ptr);
        c->count_launch();
    }
    for (int shell = 0; shell < 2; ++shell) {
        bool quad = shell == 0;
        int64_t ne = quad ? m.n_quads : m.n_plates;
        DevBuf<double> &val = quad ? ld.qp_val : ld.pp_val;
        if (!ne || !val.n) continue;
        DevBuf<double> &press = c->sym_press;
        press.resize((int64_t)nc * ne);
        press.zero(st);
        k_scatter_triplets<<<grid_for(val.n, 256), 256, 0, st>>>(val.n, quad ? ld.qp_elem.ptr : ld.pp_elem.ptr,
                                                                  quad ? ld.qp_case.ptr : ld.pp_case.ptr, val.ptr, ne,
                                                                  press.ptr);
        c->count_launch();
        const std::vector<uint8_t> &has = quad ? ld.case_has_quad : ld.case_has_plate;
        for (int cs = 0; cs < nc; ++cs)
            if (has[cs])
                launch_shell_fer(c, quad, press.ptr + (int64_t)cs * ne,
                                 ld.fer_case.ptr + (int64_t)cs * c->ef_count + (quad ? c->ef_off_quad : c->ef_off_plate), st);
        PN_CUDA(cudaStreamSynchronize(st));
    }
    tick.lap("P / FER per case");
    if (nc && n_dof) {
        k_case_vectors<<<grid_for(n_dof * nc, 256), 256, 0, st>>>(n_dof, nc, c->inc_start.ptr, c->inc_src.ptr,
                                                                   ld.Pcase.ptr, ld.fer_case.ptr, c->ef_count, ld.Gcase.ptr);
        c->count_launch();
    }
    tick.lap("case vectors");
    ld.vectors_ready = true;
}

// ------------------------------------------------------------------------------------------------
// right-hand sides: B[i][j] = sum_case f[j][case] G[case][d1[i]] - (K12 D2)[i]   (FEModel3D.py:2311)
// ------------------------------------------------------------------------------------------------
__global__ void k_spmv_rows(int64_t rows, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                            const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double acc = 0.0;
    for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) acc = fma(vals[k], x[indices[k]], acc);
    y[r] = acc;
}
void spmv_plain(Ctx *c, const CsrBlock &A, const double *vals, const double *x, double *y, cudaStream_t st) {
    if (!A.rows) return;
    k_spmv_rows<<<grid_for(A.rows, 128), 128, 0, st>>>(A.rows, A.indptr.ptr, A.indices.ptr, vals, x, y);
    c->count_launch();
}

__global__ void k_build_rhs(int64_t n1, int s, int n_case, const int32_t *__restrict__ d1, const double *__restrict__ G,
                            int64_t n_dof, const double *__restrict__ fac, const double *__restrict__ kd2, int kd2_percol,
                            double *__restrict__ B) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n1 * s) return;
    int64_t i = t / s; int j = (int)(t - i * s);
    int64_t d = d1[i];
    double acc = 0.0;
    for (int cs = 0; cs < n_case; ++cs) acc = fma(fac[(int64_t)j * n_case + cs], G[(int64_t)cs * n_dof + d], acc);
    B[t] = acc - (kd2_percol ? kd2[t] : kd2[i]);
}

// Same arithmetic (fma over the cases in ascending order: bit-identical), eight combinations per thread: the case
// vector entries G[cs][d] are loaded once per eight outputs and the factor table sits in shared memory as [case][combo]
// (the one-output-per-thread kernel above issued 16 loads per output and ran at 4 % of the HBM rate of its 0.9 GB).
constexpr int RHS_JB = 8;
__global__ void __launch_bounds__(256)
k_build_rhs8(int64_t n1, int s, int n_case, const int32_t *__restrict__ d1, const double *__restrict__ G, int64_t n_dof,
             const double *__restrict__ fac, const double *__restrict__ kd2, int kd2_percol, double *__restrict__ B) {
    extern __shared__ double s_fac[];                      // [n_case][s]
    for (int e = threadIdx.x; e < s * n_case; e += blockDim.x) {
        const int j = e / n_case, cs = e - j * n_case;
        s_fac[cs * s + j] = fac[e];
    }
    __syncthreads();
    const int groups = (s + RHS_JB - 1) / RHS_JB;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n1 * groups) return;
    const int64_t i = t / groups;
    const int jj = (int)(t - i * groups);                  // this thread's combinations: jj, jj + groups, jj + 2 groups, ...
    const int64_t d = d1[i];                               // (consecutive threads write consecutive columns of B)
    double acc[RHS_JB];
#pragma unroll
    for (int q = 0; q < RHS_JB; ++q) acc[q] = 0.0;
    for (int cs = 0; cs < n_case; ++cs) {
        const double g = G[(int64_t)cs * n_dof + d];
        const double *f = s_fac + cs * s + jj;
#pragma unroll
        for (int q = 0; q < RHS_JB; ++q)
            if (jj + q * groups < s) acc[q] = fma(f[q * groups], g, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < RHS_JB; ++q) {
        const int j = jj + q * groups;
        if (j < s) B[i * s + j] = acc[q] - (kd2_percol ? kd2[i * s + j] : kd2[i]);
    }
}

// kd2_percol: optional [n1][s] array of per-combo K12 D2 products (P-Delta second step); null = use
// the elastic K12 for every combo.
void build_rhs(Ctx *c, int32_t combo0, int s, const double *kd2_percol) {
    cudaStream_t st = c->stream;
    Loads &ld = c->loads;
    c->B.resize(c->n1 * s);
    c->X.resize(c->n1 * s);
    c->kd2.resize(c->n1);
    c->kd2.zero(st);
    if (c->n1 == 0 || s == 0) return;
    if (!kd2_percol && c->n2) spmv_plain(c, c->blk[1], c->blk[1].vals.ptr, c->d2_val.ptr, c->kd2.ptr, st);
    {
        ProfScope ps_(c, PS_RHS);
        const size_t fac_smem = (size_t)s * std::max(ld.n_case, 1) * sizeof(double);
        if (fac_smem <= 40 * 1024) {
            const int64_t work = c->n1 * ((s + RHS_JB - 1) / RHS_JB);
            k_build_rhs8<<<grid_for(work, 256), 256, fac_smem, st>>>(c->n1, s, ld.n_case, c->d1.ptr, ld.Gcase.ptr, c->n_dof,
                                                                     ld.factors.ptr + (int64_t)combo0 * ld.n_case,
                                                                     kd2_percol ? kd2_percol : c->kd2.ptr, kd2_percol ? 1 : 0, c->B.ptr);
        } else {
            k_build_rhs<<<grid_for(c->n1 * s, 256), 256, 0, st>>>(c->n1, s, ld.n_case, c->d1.ptr, ld.Gcase.ptr, c->n_dof,
                                                                  ld.factors.ptr + (int64_t)combo0 * ld.n_case,
                                                                  kd2_percol ? kd2_percol : c->kd2.ptr, kd2_percol ? 1 : 0, c->B.ptr);
        }
    }
    c->count_launch();
}

// Analysis._store_displacements / _unpartition :810-880: D = merge(D1, D2)
__global__ void k_merge(int64_t n_dof, int s, const int32_t *__restrict__ newidx, const double *__restrict__ X,
                        const double *__restrict__ d2_val, double *__restrict__ D) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_dof * s) return;
    int j = (int)(t / n_dof);
    int64_t d = t - (int64_t)j * n_dof;
    int32_t ni = newidx[d];
    D[t] = ni >= 0 ? X[(int64_t)ni * s + j] : d2_val[-1 - ni];
}
// The same copy as a tiled transpose: X is [n1][s] (combination fastest), D is [s][n_dof] (DOF fastest); a block moves
// 32 DOFs x 32 combinations through shared memory so that both sides are accessed in 256-byte runs (the direct kernel
// above read X with a stride of s doubles: 1.9 ms for 1.5 GB at config 3).
__global__ void __launch_bounds__(256)
k_merge_tiled(int64_t n_dof, int s, const int32_t *__restrict__ newidx, const double *__restrict__ X,
              const double *__restrict__ d2_val, double *__restrict__ D) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
    const int64_t d0 = (int64_t)blockIdx.x * 32;
    for (int c0 = 0; c0 < s; c0 += 32) {
        for (int dd = ty; dd < 32; dd += 8) {
            const int64_t d = d0 + dd;
            double v = 0.0;
            if (d < n_dof && c0 + tx < s) {
                const int32_t ni = newidx[d];
                v = ni >= 0 ? X[(int64_t)ni * s + c0 + tx] : d2_val[-1 - ni];
            }
            tile[dd][tx] = v;
        }
        __syncthreads();
        for (int jj = ty; jj < 32; jj += 8) {
            const int j = c0 + jj;
            if (j < s && d0 + tx < n_dof) D[(int64_t)j * n_dof + d0 + tx] = tile[tx][jj];
        }
        __syncthreads();
    }
}

void merge_displacements(Ctx *c, int32_t combo0, int s) {
    if (!c->n_dof || !s) return;
    { ProfScope ps_(c, PS_MERGE);
      if (s >= 8)
          k_merge_tiled<<<grid_for(c->n_dof, 32), 256, 0, c->stream>>>(c->n_dof, s, c->newidx.ptr, c->X.ptr, c->d2_val.ptr,
                                                                        c->D_all.ptr + (int64_t)combo0 * c->n_dof);
      else
          k_merge<<<grid_for(c->n_dof * s, 256), 256, 0, c->stream>>>(c->n_dof, s, c->newidx.ptr, c->X.ptr, c->d2_val.ptr,
                                                                     c->D_all.ptr + (int64_t)combo0 * c->n_dof); }
    c->count_launch();
}

// ------------------------------------------------------------------------------------------------
// reactions (Analysis._calc_reactions :1059-1313)
// ------------------------------------------------------------------------------------------------
// thread = (row i, combination blockIdx.y of the batch); rows = null: every DOF (i = d), else the listed DOFs
__global__ void k_reactions(int64_t n_rows, const int32_t *__restrict__ rows, int64_t n_dof, int n_case,
                            const uint8_t *__restrict__ support,
                            const uint32_t *__restrict__ inc_start, const uint32_t *__restrict__ inc_src,
                            const double *__restrict__ ef, int64_t ef_stride, const double *__restrict__ fac_row,
                            const double *__restrict__ Pcase, const uint8_t *__restrict__ nsp_flag,
                            const double *__restrict__ nsp_k, const double *__restrict__ D, double *__restrict__ R) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const int64_t d = rows ? rows[i] : i;
    ef += blockIdx.y * ef_stride;
    fac_row += blockIdx.y * (int64_t)n_case;
    D += blockIdx.y * n_dof;
    R += blockIdx.y * n_rows;
    uint8_t sup = support[d / 6];
    double r = 0.0;
    if ((sup >> (d % 6)) & 1u) {
        for (uint32_t k = inc_start[d]; k < inc_start[d + 1]; ++k) r += ef[inc_src[k]];
        double pl = 0.0;
        for (int cs = 0; cs < n_case; ++cs) pl = fma(fac_row[cs], Pcase[(int64_t)cs * n_dof + d], pl);
        r -= pl;
    }
    uint8_t f = nsp_flag[d];
    if ((f & 1u) && (f & 2u)) r -= nsp_k[d] * D[d];
    R[i] = r;
}

void compute_reactions(Ctx *c, int32_t combo, bool pdelta, double *rxn_dev) {
    cudaStream_t st = c->stream;
    Loads &ld = c->loads;
    const Model &m = c->model;
    const double *D = c->D_all.ptr + (int64_t)combo * c->n_dof;
    launch_element_forces(c, PN_SPRING, D, combo, true, c->ef.ptr + c->ef_off_spring, st);
    launch_element_forces(c, PN_MEMBER, D, combo, true, c->ef.ptr + c->ef_off_member, st);
    if (pdelta) launch_member_pdelta_forces(c, D, true, c->ef.ptr + c->ef_off_member, st);
    launch_element_forces(c, PN_QUAD, D, combo, true, c->ef.ptr + c->ef_off_quad, st);
    launch_element_forces(c, PN_PLATE, D, combo, true, c->ef.ptr + c->ef_off_plate, st);
    if (c->n_dof) {
        ProfScope ps_(c, PS_REACTIONS);
        k_reactions<<<grid_for(c->n_dof, 256), 256, 0, st>>>(c->n_dof, nullptr, c->n_dof, ld.n_case, m.support.ptr, c->inc_start.ptr,
                                                              c->inc_src.ptr, c->ef.ptr, 0,
                                                              ld.factors.ptr + (int64_t)combo * ld.n_case, ld.Pcase.ptr,
                                                              m.nspring_flag.ptr, m.nspring_k.ptr, D, rxn_dev);
        c->count_launch();
    }
}

void compute_reactions_batch(Ctx *c, int32_t combo0, int n_batch, bool pdelta, const int32_t *rows_dev, int64_t n_rows,
                             double *out_dev) {
    if (n_batch <= 0 || !c->n_dof) return;
    cudaStream_t st = c->stream;
    Loads &ld = c->loads;
    const Model &m = c->model;
    const int64_t nr = rows_dev ? n_rows : c->n_dof;
    // element end forces of up to `chunk` combinations at a time (<= 2 GB of scratch)
    const int chunk = (int)std::max<int64_t>(1, std::min<int64_t>(n_batch, (int64_t)(2.0e9 / (8.0 * (double)std::max<int64_t>(c->ef_count, 1)))));
    c->ef_batch.resize((int64_t)chunk * c->ef_count);
    for (int j0 = 0; j0 < n_batch; j0 += chunk) {
        const int nb = std::min(chunk, n_batch - j0);
        const int32_t combo = combo0 + j0;
        const double *D = c->D_all.ptr + (int64_t)combo * c->n_dof;
        double *ef = c->ef_batch.ptr;
        launch_element_forces(c, PN_SPRING, D, combo, true, ef + c->ef_off_spring, st, nb, c->ef_count);
        launch_element_forces(c, PN_MEMBER, D, combo, true, ef + c->ef_off_member, st, nb, c->ef_count);
        if (pdelta)
            for (int j = 0; j < nb; ++j)
                launch_member_pdelta_forces(c, D + (int64_t)j * c->n_dof, true, ef + (int64_t)j * c->ef_count + c->ef_off_member, st);
        launch_element_forces(c, PN_QUAD, D, combo, true, ef + c->ef_off_quad, st, nb, c->ef_count);
        launch_element_forces(c, PN_PLATE, D, combo, true, ef + c->ef_off_plate, st, nb, c->ef_count);
        if (nr) {
            ProfScope ps_(c, PS_REACTIONS);
            k_reactions<<<dim3(grid_for(nr, 256), nb), 256, 0, st>>>(nr, rows_dev, c->n_dof, ld.n_case, m.support.ptr, c->inc_start.ptr,
                                                                     c->inc_src.ptr, ef, c->ef_count,
                                                                     ld.factors.ptr + (int64_t)combo * ld.n_case, ld.Pcase.ptr,
                                                                     m.nspring_flag.ptr, m.nspring_k.ptr, D, out_dev + (int64_t)j0 * nr);
            c->count_launch();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// SpMM: Y[rows][s] = A X, X and Y row-major with s columns.  G lanes cooperate on one row, each lane
// owning CPL columns spaced G apart, so every X access is a contiguous G*8-byte segment.
// ------------------------------------------------------------------------------------------------
template <int G, int CPL>
__global__ void __launch_bounds__(256)
k_spmm(int64_t rows, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
       const double *__restrict__ vals, const double *__restrict__ X, double *__restrict__ Y, int s) {
    int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t r = gid / G;
    int lane = (int)(gid % G);
    if (r >= rows) return;
    double acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = 0.0;
    int32_t k0 = indptr[r], k1 = indptr[r + 1];
    for (int32_t k = k0; k < k1; ++k) {
        double v = vals[k];
        const double *x = X + (int64_t)indices[k] * s + lane;
#pragma unroll
        for (int c = 0; c < CPL; ++c)
            if (lane + c * G < s) acc[c] = fma(v, x[c * G], acc[c]);
    }
    double *y = Y + r * s + lane;
#pragma unroll
    for (int c = 0; c < CPL; ++c)
        if (lane + c * G < s) y[c * G] = acc[c];
}

// ------------------------------------------------------------------------------------------------
// Node-grouped SpMM.  The rows of one node (<= 6 free DOFs) reference almost the same columns -- the DOFs of the
// node's neighbours -- so the row-per-lane-group kernel above loads every X row once per ROW that references it
// (measured: 15 GB of L2 -> SM traffic at 64 right-hand sides for 1.9 GB of algorithmic bytes).  Here LANES lanes
// own the <= 8 rows of a group together: lane r < R walks row r's entries (one entry ahead), the group takes the
// smallest pending column, loads that X row ONCE (coalesced, CPL columns per lane) and every row whose cursor sits on
// the column accumulates it; the values travel by shuffle.  Summation order per row = column order, as before.
// ------------------------------------------------------------------------------------------------
template <int LANES, int CPL>
__global__ void __launch_bounds__(256)
k_spmm_grouped(int64_t n_groups, const int32_t *__restrict__ group, const int32_t *__restrict__ indptr,
               const int32_t *__restrict__ indices, const double *__restrict__ vals, const double *__restrict__ X,
               double *__restrict__ Y, int s) {
    constexpr int MAXR = 8;
    const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t g = gid / LANES;
    const int lane = (int)(gid % LANES);
    const unsigned seg = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
    if (g >= n_groups) return;                      // (whole lane groups leave together: n_groups * LANES is the bound)
    const int r0 = group[g], R = group[g + 1] - r0;
    const int base = (threadIdx.x & 31) / LANES * LANES;      // first lane of this group inside the warp
    // cursor of row `lane` (lanes >= R idle in the merge): current and next entry preloaded
    int k = 0, kend = 0, col = INT32_MAX, col_n = INT32_MAX;
    double val = 0.0, val_n = 0.0;
    if (lane < R) {
        k = indptr[r0 + lane]; kend = indptr[r0 + lane + 1];
        if (k < kend) { col = indices[k]; val = vals[k]; }
        if (k + 1 < kend) { col_n = indices[k + 1]; val_n = vals[k + 1]; }
    }
    double acc[MAXR][CPL];
#pragma unroll
    for (int r = 0; r < MAXR; ++r)
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[r][c] = 0.0;
    while (true) {
        const int cmin = __reduce_min_sync(seg, col);
        if (cmin == INT32_MAX) break;
        double x[CPL];
        const double *xp = X + (int64_t)cmin * s + lane;
#pragma unroll
        for (int c = 0; c < CPL; ++c) x[c] = (lane + c * LANES < s) ? xp[c * LANES] : 0.0;
        const unsigned hit = __ballot_sync(seg, col == cmin) >> base;
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
            if (hit & (1u << r)) {
                const double v = __shfl_sync(seg, val, base + r);
#pragma unroll
                for (int c = 0; c < CPL; ++c) acc[r][c] = fma(v, x[c], acc[r][c]);
            }
        }
        if (col == cmin) {                         // advance this row's cursor; keep one entry prefetched
            ++k;
            col = col_n; val = val_n;
            if (k + 1 < kend) { col_n = indices[k + 1]; val_n = vals[k + 1]; } else col_n = INT32_MAX;
        }
    }
#pragma unroll
    for (int r = 0; r < MAXR; ++r)
        if (r < R) {
            double *y = Y + (int64_t)(r0 + r) * s + lane;
#pragma unroll
            for (int c = 0; c < CPL; ++c)
                if (lane + c * LANES < s) y[c * LANES] = acc[r][c];
        }
}

// groups of K11 rows by node (host masks -> device list), once per partition
static void ensure_row_groups(Ctx *c) {
    if (c->row_groups_ready) return;
    const Model &m = c->model;
    std::vector<int32_t> g;
    g.reserve((size_t)m.n_nodes + 1);
    int32_t row = 0;
    for (int64_t nd = 0; nd < m.n_nodes; ++nd) {
        const unsigned known = (unsigned)(m.h_support[nd] | m.h_enf_mask[nd]) & 63u;
        const int free_dofs = 6 - __builtin_popcount(known);
        if (free_dofs) { g.push_back(row); row += free_dofs; }
    }
    g.push_back(row);
    if (row != c->n1) throw ArgError("spmm: node row groups do not cover K11");
    c->n_row_groups = (int64_t)g.size() - 1;
    c->row_group.upload(g.data(), (int64_t)g.size(), c->stream);
    PN_CUDA(cudaStreamSynchronize(c->stream));
    c->row_groups_ready = true;
}

// K11 only (the grouping is by node of the free DOFs); s <= 256
void spmm_k11_grouped(Ctx *c, const double *vals, const double *X, double *Y, int s, cudaStream_t st) {
    const CsrBlock &A = c->blk[0];
    if (!A.rows || !s) return;
    ensure_row_groups(c);
    const int64_t ng = c->n_row_groups;
    ProfScope ps_(c, PS_SPMM);
#define PN_SPG(L_, C_) k_spmm_grouped<L_, C_><<<grid_for(ng * L_, 256), 256, 0, st>>>(ng, c->row_group.ptr, A.indptr.ptr, A.indices.ptr, vals, X, Y, s)
    if (s <= 8) PN_SPG(8, 1);
    else if (s <= 16) PN_SPG(16, 1);
    else if (s <= 32) PN_SPG(32, 1);
    else if (s <= 64) PN_SPG(32, 2);
    else if (s <= 128) PN_SPG(32, 4);
    else if (s <= 256) PN_SPG(32, 8);
    else throw ArgError("spmm: more than 256 right-hand sides per call");
#undef PN_SPG
    c->count_launch();
}

// R[rows][s] = B - A X with the row sums carried in double-double (TwoProd by fma, TwoSum), so that the residual
// that drives iterative refinement and the singularity check is accurate to ~1e-32 |A||X| instead of 1e-16 |A||X|.
// K11 of a large plate has cond ~ 1e9..1e10: a plain FP64 residual is then pure rounding noise at 1e-7 |B|.
template <int G, int CPL>
__global__ void __launch_bounds__(256)
k_residual_dd(int64_t rows, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
              const double *__restrict__ vals, const double *__restrict__ X, const double *__restrict__ B,
              double *__restrict__ R, int s, int64_t vals_col_stride, const int32_t *__restrict__ rowlist, int64_t row0) {
    int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t r = gid / G;
    int lane = (int)(gid % G);
    if (r >= rows) return;
    r = rowlist ? (int64_t)rowlist[r] : r + row0;          // a subset of the rows (subtree-to-GPU solve), or a row range
    double hi[CPL], lo[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) { hi[c] = (lane + c * G < s) ? B[r * s + lane + c * G] : 0.0; lo[c] = 0.0; }
    int32_t k0 = indptr[r], k1 = indptr[r + 1];
    for (int32_t k = k0; k < k1; ++k) {
        const double *x = X + (int64_t)indices[k] * s + lane;
#pragma unroll
        for (int c = 0; c < CPL; ++c)
            if (lane + c * G < s) {
                double v = vals[k + (int64_t)(lane + c * G) * vals_col_stride];
                double xv = x[c * G];
                double t = v * xv;
                double te = fma(v, xv, -t);            // v * xv = t + te exactly
                double sum = hi[c] - t;
                double bb = sum - hi[c];
                double err = (hi[c] - (sum - bb)) + (-t - bb);
                hi[c] = sum;
                lo[c] += err - te;
            }
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c)
        if (lane + c * G < s) R[r * s + lane + c * G] = hi[c] + lo[c];
}

template <int G>
static void residual_dispatch_cpl(Ctx *c, const CsrBlock &A, const double *vals, int64_t stride, const double *X,
                                  const double *B, double *R, int s, cudaStream_t st, const int32_t *rowlist, int64_t row0,
                                  int64_t n_rows) {
    int cpl = (s + G - 1) / G;
    unsigned grid = grid_for(n_rows * G, 256);
#define PN_RES(CPL_) do { ProfScope ps_(c, PS_SPMM); k_residual_dd<G, CPL_><<<grid, 256, 0, st>>>(n_rows, A.indptr.ptr, A.indices.ptr, vals, X, B, R, s, stride, rowlist, row0); } while (0)
    if (cpl <= 1) PN_RES(1);
    else if (cpl <= 2) PN_RES(2);
    else if (cpl <= 4) PN_RES(4);
    else if (cpl <= 8) PN_RES(8);
    else throw ArgError("residual: more than 256 right-hand sides per call");
#undef PN_RES
    c->count_launch();
}

// vals_col_stride = 0: one matrix for all columns; = nnz: vals is [s][nnz] (one matrix per column, P-Delta).
// rowlist (n_rows entries) restricts the residual to those rows, otherwise rows [row0, row0 + n_rows) (n_rows < 0: all);
// R keeps the full [rows][s] layout either way.
void residual_dd(Ctx *c, const CsrBlock &A, const double *vals, int64_t vals_col_stride, const double *X, const double *B,
                 double *R, int s, cudaStream_t st, const int32_t *rowlist, int64_t row0, int64_t n_rows) {
    if (n_rows < 0) { rowlist = nullptr; row0 = 0; n_rows = A.rows; }
    if (!A.rows || !s || !n_rows) return;
#define PN_RESG(G_) residual_dispatch_cpl<G_>(c, A, vals, vals_col_stride, X, B, R, s, st, rowlist, row0, n_rows)
    if (s == 1) PN_RESG(1);
    else if (s == 2) PN_RESG(2);
    else if (s <= 4) PN_RESG(4);
    else if (s <= 8) PN_RESG(8);
    else if (s <= 16) PN_RESG(16);
    else PN_RESG(32);
#undef PN_RESG
}

template <int G>
static void spmm_dispatch_cpl(Ctx *c, const CsrBlock &A, const double *vals, const double *X, double *Y, int s,
                              cudaStream_t st) {
    int cpl = (s + G - 1) / G;
    unsigned grid = grid_for(A.rows * G, 256);
#define PN_SPMM(CPL_) do { ProfScope ps_(c, PS_SPMM); k_spmm<G, CPL_><<<grid, 256, 0, st>>>(A.rows, A.indptr.ptr, A.indices.ptr, vals, X, Y, s); } while (0)
    if (cpl <= 1) PN_SPMM(1);
    else if (cpl <= 2) PN_SPMM(2);
    else if (cpl <= 4) PN_SPMM(4);
    else if (cpl <= 8) PN_SPMM(8);
    else throw ArgError("spmm: more than 256 right-hand sides per call");
#undef PN_SPMM
    c->count_launch();
}

// rows [row0, row1) only: indptr values are absolute, so offsetting the row pointer and the output is enough
void spmm_rows(Ctx *c, const CsrBlock &A, const double *vals, const double *X, double *Y, int s, int64_t row0, int64_t row1,
               cudaStream_t st) {
    if (row1 <= row0 || !s) return;
    CsrBlock view;
    view.rows = row1 - row0;
    view.cols = A.cols;
    view.nnz = A.nnz;
    view.indptr.ptr = A.indptr.ptr + row0;      // non-owning view: detached again before it goes out of scope
    view.indices.ptr = A.indices.ptr;
    try {
        spmm(c, view, vals, X, Y + row0 * s, s, st);
    } catch (...) {
        view.indptr.ptr = nullptr; view.indices.ptr = nullptr;
        throw;
    }
    view.indptr.ptr = nullptr; view.indices.ptr = nullptr;
}

void spmm(Ctx *c, const CsrBlock &A, const double *vals, const double *X, double *Y, int s, cudaStream_t st) {
    if (!A.rows || !s) return;
    if (s == 1) spmm_dispatch_cpl<1>(c, A, vals, X, Y, s, st);
    else if (s == 2) spmm_dispatch_cpl<2>(c, A, vals, X, Y, s, st);
    else if (s <= 4) spmm_dispatch_cpl<4>(c, A, vals, X, Y, s, st);
    else if (s <= 8) spmm_dispatch_cpl<8>(c, A, vals, X, Y, s, st);
    else if (s <= 16) spmm_dispatch_cpl<16>(c, A, vals, X, Y, s, st);
    else spmm_dispatch_cpl<32>(c, A, vals, X, Y, s, st);
}

}  // namespace pn
