// Tension/compression-only activity update on the device: what Analysis._check_TC_convergence
// (Analysis.py:917-1056) decides after every solve of `analyze` / `analyze_PDelta`, without the flags, displacements
// or end forces leaving the GPU:
//   node support springs with a direction ('+' / '-'): active iff the displacement has crossed the tolerance in that
//       direction (:938-960) -- they can switch both ways;
//   two-node springs, tension-only / compression-only: deactivated when their axial force has the wrong sign (:974-996);
//   physical members, tension-only / compression-only: deactivated (all their sub-members) when the largest / smallest
//       axial force along the member has the wrong sign (:1017-1047; Member3D.max_axial / min_axial walk the axial
//       diagram: end force plus the axial loads up to the section).
// Element deactivations are never undone inside a combination (the reference only ever sets active = False).
#include "pn_internal.h"
#include "elem_line.cuh"

namespace pn {

static inline unsigned grid_for(int64_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

// spring_flag bits: 1 defined, 2 active, 4 direction '+', 8 direction '-'
__global__ void k_tc_node_springs(int64_t n_dof, const double *__restrict__ D, double tol, uint8_t *__restrict__ flag,
                                  unsigned long long *__restrict__ changed) {
    int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (d >= n_dof) return;
    const uint8_t f = flag[d];
    if (!(f & 1u) || !(f & 12u)) return;
    const double u = D[d];
    const bool want = ((f & 8u) && u <= -tol) || ((f & 4u) && u >= tol);
    const bool is = (f & 2u) != 0;
    if (want != is) {
        flag[d] = want ? (uint8_t)(f | 2u) : (uint8_t)(f & ~2u);
        atomicAdd(changed, 1ull);
    }
}

// mode bits: 1 tension-only, 2 compression-only.  Spring3D.axial = f[0] = ks (d0 - d6) in local axes (Spring3D.py:86-112)
__global__ void k_tc_springs(int64_t n, const int32_t *__restrict__ ij, const double *__restrict__ ks, const double *__restrict__ xyz,
                             const double *__restrict__ D, const uint8_t *__restrict__ mode, double tol, uint8_t *__restrict__ active,
                             unsigned long long *__restrict__ changed) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n || !mode[e] || !active[e]) return;
    const int64_t i = ij[2 * e], j = ij[2 * e + 1];
    double pi[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    double pj[3] = {xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]};
    double R[9];
    pn_line_dircos(pi, pj, 0.0, R);
    const double d0 = R[0] * D[6 * i] + R[1] * D[6 * i + 1] + R[2] * D[6 * i + 2];
    const double d6 = R[0] * D[6 * j] + R[1] * D[6 * j + 1] + R[2] * D[6 * j + 2];
    const double axial = ks[e] * d0 + (-ks[e]) * d6;
    if (((mode[e] & 1u) && axial > tol) || ((mode[e] & 2u) && axial < -tol)) {
        active[e] = 0;
        atomicAdd(changed, 1ull);
    }
}

// (max, min) of the axial-force diagram of every sub-member: f_local[0] plus the axial loads applied before the section.
// One thread per sub-member; its loads are the (member, case) groups of pn_set_loads (sorted by member).
__global__ void k_tc_member_extremes(int64_t n, const int32_t *__restrict__ ij, const double *__restrict__ rot,
                                     const double *__restrict__ xyz, const double *__restrict__ f_local /* [n][12] */,
                                     int64_t n_groups, const int32_t *__restrict__ g_member, const int32_t *__restrict__ g_case,
                                     const int64_t *__restrict__ g_start, const int32_t *__restrict__ ml_kind,
                                     const double *__restrict__ ml_params, const double *__restrict__ fac_row,
                                     double *__restrict__ out /* [n][2] */) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double f0 = f_local[e * 12];
    // first group of this member (binary search in the sorted group list)
    int64_t lo = 0, hi = n_groups;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (g_member[mid] < e) lo = mid + 1; else hi = mid;
    }
    int64_t g0 = lo, g1 = lo;
    while (g1 < n_groups && g_member[g1] == e) ++g1;
    double best_max = f0, best_min = f0;
    if (g1 == g0) { out[2 * e] = best_max; out[2 * e + 1] = best_min; return; }
    const int64_t i = ij[2 * e], j = ij[2 * e + 1];
    double pi[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    double pj[3] = {xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]};
    double R[9];
    const double L = pn_line_dircos(pi, pj, rot[e], R);
    // axial load of load record l: point (x, p) or distributed (x1, x2, p1, p2); returns 0 none, 1 point, 2 distributed
    auto axial_load = [&](int64_t g, int64_t l, double &a, double &b, double &p1, double &p2) -> int {
        const double factor = fac_row[g_case[g]];
        if (factor == 0.0) return 0;                    // case absent from the combination (Member3D.py: factors.get)
        const int kind = ml_kind[l];
        const double *p = ml_params + 4 * l;
        if (kind == 0) { a = p[1]; p1 = factor * p[0]; return 1; }
        if (kind >= 6 && kind <= 8) { a = p[1]; p1 = factor * (R[kind - 6] * p[0]); return 1; }
        if (kind == 12) { a = p[2]; b = p[3]; p1 = factor * p[0]; p2 = factor * p[1]; return 2; }
        if (kind >= 15 && kind <= 17) { a = p[2]; b = p[3]; p1 = factor * (R[kind - 15] * p[0]); p2 = factor * (R[kind - 15] * p[1]); return 2; }
        return 0;
    };
    auto axial_at = [&](double x, bool after) -> double {
        double nforce = f0;
        for (int64_t g = g0; g < g1; ++g)
            for (int64_t l = g_start[g]; l < g_start[g + 1]; ++l) {
                double a = 0, b = 0, p1 = 0, p2 = 0;
                const int t = axial_load(g, l, a, b, p1, p2);
                if (t == 1) { if (a < x || (after && a == x)) nforce += p1; }
                else if (t == 2 && x > a && b > a) {
                    const double xe = fmin(x, b);
                    const double pe = p1 + (p2 - p1) * (xe - a) / (b - a);
                    nforce += 0.5 * (p1 + pe) * (xe - a);
                }
            }
        return nforce;
    };
    auto visit = [&](double x) {
        for (int after = 0; after < 2; ++after) {
            const double v = axial_at(x, after != 0);
            best_max = fmax(best_max, v);
            best_min = fmin(best_min, v);
        }
    };
    best_max = -1.7976931348623157e308; best_min = 1.7976931348623157e308;
    visit(0.0);
    visit(L);
    for (int64_t g = g0; g < g1; ++g)
        for (int64_t l = g_start[g]; l < g_start[g + 1]; ++l) {
            double a = 0, b = 0, p1 = 0, p2 = 0;
            const int t = axial_load(g, l, a, b, p1, p2);
            if (t == 1) visit(a);
            else if (t == 2) {
                visit(a);
                visit(b);
                if (p1 * p2 < 0) visit(a + (b - a) * p1 / (p1 - p2));
            }
        }
    out[2 * e] = best_max;
    out[2 * e + 1] = best_min;
}

// one thread per physical member (its sub-members are contiguous: [start[g], start[g + 1]))
__global__ void k_tc_members(int64_t n_groups, const int64_t *__restrict__ start, const uint8_t *__restrict__ mode /* per sub-member */,
                             const double *__restrict__ ext, double tol, uint8_t *__restrict__ active,
                             unsigned long long *__restrict__ changed) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int64_t a = start[g], b = start[g + 1];
    if (b <= a || !mode[a] || !active[a]) return;
    double mx = -1.7976931348623157e308, mn = 1.7976931348623157e308;
    for (int64_t e = a; e < b; ++e) { mx = fmax(mx, ext[2 * e]); mn = fmin(mn, ext[2 * e + 1]); }
    if (((mode[a] & 1u) && mx > tol) || ((mode[a] & 2u) && mn < -tol)) {
        for (int64_t e = a; e < b; ++e) active[e] = 0;
        atomicAdd(changed, 1ull);
    }
}

__global__ void k_fill_u8(int64_t n, uint8_t v, uint8_t *__restrict__ p) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// Returns the number of state changes (0 = converged).  D_dev: the combination's displacement vector.
int64_t tc_update(Ctx *c, int32_t combo, bool pdelta, double spring_tol, double member_tol) {
    Model &m = c->model;
    Loads &ld = c->loads;
    cudaStream_t st = c->stream;
    const double *D = c->D_all.ptr + (int64_t)combo * c->n_dof;
    DevBuf<unsigned long long> changed;
    changed.resize(1);
    changed.zero(st);
    if (c->n_dof) {
        k_tc_node_springs<<<grid_for(c->n_dof, 256), 256, 0, st>>>(c->n_dof, D, spring_tol, m.nspring_flag.ptr, changed.ptr);
        c->count_launch();
    }
    if (m.n_springs && c->tc_spring_mode.n == m.n_springs) {
        k_tc_springs<<<grid_for(m.n_springs, 128), 128, 0, st>>>(m.n_springs, m.spring_ij.ptr, m.spring_ks.ptr, m.xyz.ptr, D,
                                                                  c->tc_spring_mode.ptr, spring_tol, m.spring_active.ptr, changed.ptr);
        c->count_launch();
    }
    if (m.n_members && c->tc_member_mode.n == m.n_members && c->tc_any_member_mode) {
        // local end forces of every sub-member (Member3D.f :872-918, with the kg(P) d term after a P-Delta step)
        double *F = c->ef.ptr + c->ef_off_member;
        launch_element_forces(c, PN_MEMBER, D, combo, false, F, st);
        if (pdelta) launch_member_pdelta_forces(c, D, false, F, st);
        launch_forces_to_local(c, PN_MEMBER, F, st);
        DevBuf<double> ext;
        ext.resize(2 * m.n_members);
        k_tc_member_extremes<<<grid_for(m.n_members, 64), 64, 0, st>>>(
            m.n_members, m.mem_ij.ptr, m.mem_rot.ptr, m.xyz.ptr, F, ld.n_mgroups, ld.mg_member.ptr, ld.mg_case.ptr, ld.mg_start.ptr,
            ld.ml_kind.ptr, ld.ml_params.ptr, ld.factors.ptr + (int64_t)combo * ld.n_case, ext.ptr);
        k_tc_members<<<grid_for(c->tc_n_groups, 128), 128, 0, st>>>(c->tc_n_groups, c->tc_group_start.ptr, c->tc_member_mode.ptr, ext.ptr,
                                                                    member_tol, m.mem_active.ptr, changed.ptr);
        c->count_launch(2);
        PN_CUDA(cudaStreamSynchronize(st));          // ext is released on return
    }
    unsigned long long h = 0;
    PN_CUDA(cudaMemcpyAsync(&h, changed.ptr, sizeof(h), cudaMemcpyDeviceToHost, st));
    PN_CUDA(cudaStreamSynchronize(st));
    return (int64_t)h;
}

void tc_reset_elements(Ctx *c) {
    Model &m = c->model;
    cudaStream_t st = c->stream;
    if (m.n_springs) k_fill_u8<<<grid_for(m.n_springs, 256), 256, 0, st>>>(m.n_springs, 1, m.spring_active.ptr);
    if (m.n_members) k_fill_u8<<<grid_for(m.n_members, 256), 256, 0, st>>>(m.n_members, 1, m.mem_active.ptr);
    c->count_launch(2);
}

}  // namespace pn
