// Element kernels: batched FP64 element stiffness in global axes, written densely (all nd*nd
// entries, zeros included) into the element-value buffer in the reference's COO emission order.
//   springs, members : one thread per element (12x12; condensation needs the whole matrix)
//   quads, plates    : 16 threads per element -- stage B thread = (gauss point, node),
//                      stage C thread = (node a, node b) producing one 6x6 block in registers
// Also: member geometric stiffness, element fixed-end force vectors per load group, element end
// forces for reactions.
#include "pn_internal.h"
#include "elem_line.cuh"
#include "elem_shell.cuh"

namespace pn {

// ------------------------------------------------------------------------------------------------
// two-node elements
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_rotated_12(const double *R, const double *k, double *out) {
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            double B[9], G[9];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) B[3 * i + j] = k[(3 * a + i) * 12 + 3 * b + j];
            pn_congruence3(R, B, G);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) out[(3 * a + i) * 12 + 3 * b + j] = G[3 * i + j];
        }
}

__global__ void k_spring_ke(int64_t n, const int32_t *__restrict__ ij, const double *__restrict__ ks,
                            const double *__restrict__ xyz, double *__restrict__ out) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n) return;
    int i = ij[2 * e], j = ij[2 * e + 1];
    double pi[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    double pj[3] = {xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]};
    double R[9], k[144];
    pn_line_dircos(pi, pj, 0.0, R);
    pn_zero144(k);
    double s = ks[e];
    k[0] = s; k[6 * 12 + 6] = s; k[6] = -s; k[6 * 12] = -s;       // Spring3D.ke :59-83
    store_rotated_12(R, k, out + e * 144);
}

// mode 0: elastic stiffness; mode 1: geometric stiffness with axial force P[e]
__global__ void k_member_matrix(int64_t n, const int32_t *__restrict__ ij, const double *__restrict__ props,
                                const double *__restrict__ rot, const uint16_t *__restrict__ rel,
                                const double *__restrict__ xyz, int mode, const double *__restrict__ P,
                                double *__restrict__ out) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n) return;
    int i = ij[2 * e], j = ij[2 * e + 1];
    double pi[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    double pj[3] = {xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]};
    double pr[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) pr[k] = props[6 * e + k];
    double R[9], k[144];
    double L = pn_line_dircos(pi, pj, rot[e], R);
    if (mode == 0) pn_member_ke(pr, L, rel[e], k);
    else pn_member_kg(pr, L, rel[e], P[e], k);
    store_rotated_12(R, k, out + e * 144);
}

// Axial force per member from a global displacement vector: P = E A / L (d6 - d0), d = T D
// (FEModel3D.py:1849-1850).  D is one combo's 6N vector.
__global__ void k_member_axial(int64_t n, const int32_t *__restrict__ ij, const double *__restrict__ props,
                               const double *__restrict__ rot, const double *__restrict__ xyz,
                               const double *__restrict__ D, double *__restrict__ P) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n) return;
    int i = ij[2 * e], j = ij[2 * e + 1];
    double pi[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    double pj[3] = {xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]};
    double R[9];
    double L = pn_line_dircos(pi, pj, rot[e], R);
    double d0 = R[0] * D[6 * (int64_t)i] + R[1] * D[6 * (int64_t)i + 1] + R[2] * D[6 * (int64_t)i + 2];
    double d6 = R[0] * D[6 * (int64_t)j] + R[1] * D[6 * (int64_t)j + 1] + R[2] * D[6 * (int64_t)j + 2];
    P[e] = props[6 * e] * props[6 * e + 2] / L * (d6 - d0);
}

// Condensed local fixed-end vector of one (member, case) load group with unit load factor
// (Member3D.fer :742-772 is linear in the loads, so per-case vectors are combined by factors later).
__global__ void k_member_fer_groups(int64_t n_groups, const int32_t *__restrict__ g_member,
                                    const int64_t *__restrict__ g_start, const int32_t *__restrict__ ml_kind,
                                    const double *__restrict__ ml_params, const int32_t *__restrict__ ij,
                                    const double *__restrict__ props, const double *__restrict__ rot,
                                    const uint16_t *__restrict__ rel, const double *__restrict__ xyz,
                                    double *__restrict__ out /* [n_groups][12] local */) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    int e = g_member[g];
    int i = ij[2 * e], j = ij[2 * e + 1];
    double pi[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    double pj[3] = {xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]};
    double R[9], f[12];
    double L = pn_line_dircos(pi, pj, rot[e], R);
    for (int k = 0; k < 12; ++k) f[k] = 0.0;
    for (int64_t l = g_start[g]; l < g_start[g + 1]; ++l) {
        double p[4] = {ml_params[4 * l], ml_params[4 * l + 1], ml_params[4 * l + 2], ml_params[4 * l + 3]};
        pn_member_load_fer(ml_kind[l], p, R, L, f);
    }
    unsigned r = rel[e];
    if (r) {
        double k[144];
        pn_member_ke_unc(props[6 * e], props[6 * e + 1], props[6 * e + 2], props[6 * e + 3], props[6 * e + 4],
                         props[6 * e + 5], L, k);
        pn_condense(k, f, r);
    }
    for (int k = 0; k < 12; ++k) out[g * 12 + k] = f[k];
}

// ------------------------------------------------------------------------------------------------
// shells: 16 threads per element
// ------------------------------------------------------------------------------------------------
constexpr int SHELL_TPB = 128;               // 8 elements per block
constexpr int SHELL_EPB = SHELL_TPB / 16;
constexpr int SHELL_OUT_LD = 580;            // 576 + 4: the two elements of a warp start 8 banks apart

// Per-element data every one of the 16 threads needs: computed ONCE per element by three of them (local
// coordinates, triad, material) and four more (one DKMQ edge each) instead of redundantly by all 16 -- the square
// roots and divisions in here were half of the kernel's FP64 instructions.
struct ShellShared {
    PnShellGeom G;
    PnShellMat mat;
    double R[9];
    double width, height;
    double diag[4];
};

// The 16 threads of an element sit in one half warp, so every hand-over through shared memory below needs a
// __syncwarp() only.  The finished 24x24 block is staged in shared memory (in the space of the stage-B records,
// which are dead by then) and leaves as 128-byte contiguous stores.
template <bool QUAD>
__global__ void __launch_bounds__(SHELL_TPB, 4)
k_shell_ke(int64_t n, const int32_t *__restrict__ ijmn, const double *__restrict__ props,
           const double *__restrict__ xyz, double *__restrict__ out) {
    __shared__ double s_buf[SHELL_EPB][SHELL_OUT_LD];      // stage-B records (16 x 18 doubles), later the output block
    __shared__ ShellShared s_el[SHELL_EPB];
    static_assert(sizeof(PnGpNode) * 16 <= sizeof(double) * SHELL_OUT_LD, "stage-B records must fit the staging buffer");
    const int le = threadIdx.x >> 4, t = threadIdx.x & 15;
    int64_t e = blockIdx.x * (int64_t)SHELL_EPB + le;
    const bool valid = e < n;
    if (!valid) e = n - 1;                    // keep every thread alive for the warp-level hand-overs
    ShellShared &S = s_el[le];
    PnGpNode *s_rec = reinterpret_cast<PnGpNode *>(s_buf[le]);      // [gauss point][node]

    // stage A: one thread each for the local corner coordinates, the triad and the material
    if (t < 3) {
        double p[4][3];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            int64_t nd = ijmn[4 * e + a];
            p[a][0] = xyz[3 * nd]; p[a][1] = xyz[3 * nd + 1]; p[a][2] = xyz[3 * nd + 2];
        }
        if (t == 0) {
            if (QUAD) {
                pn_quad_local_xy(p, S.G.xl, S.G.yl);
            } else {
                // Plate3D.width/height :77-87 -- the plate is treated as a width x height rectangle
                double width = sqrt((p[0][0] - p[1][0]) * (p[0][0] - p[1][0]) + (p[0][1] - p[1][1]) * (p[0][1] - p[1][1]) +
                                    (p[0][2] - p[1][2]) * (p[0][2] - p[1][2]));
                double height = sqrt((p[0][0] - p[3][0]) * (p[0][0] - p[3][0]) + (p[0][1] - p[3][1]) * (p[0][1] - p[3][1]) +
                                     (p[0][2] - p[3][2]) * (p[0][2] - p[3][2]));
                S.width = width; S.height = height;
                S.G.xl[0] = 0; S.G.yl[0] = 0; S.G.xl[1] = width; S.G.yl[1] = 0;
                S.G.xl[2] = width; S.G.yl[2] = height; S.G.xl[3] = 0; S.G.yl[3] = height;
            }
        } else if (t == 1) {
            pn_shell_dircos(p[0], p[1], p[3], S.R);
        } else {
            double pr[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) pr[k] = props[5 * e + k];
            pn_shell_material(pr, S.mat);
        }
    }
    __syncwarp();
    if (QUAD && t < 4) pn_quad_edge(S.G, t, props[5 * e], props[5 * e + 2]);
    __syncwarp();

    // stage B: (gauss point, node)
    {
        PnGpNode rec;
        pn_shell_gp_node(S.G, t >> 2, t & 3, QUAD, rec);      // edge data indexed at run time: straight from shared memory
        s_rec[t] = rec;
    }
    __syncwarp();

    // stage C: (node a, node b)
    const int a = t >> 2, b = t & 3;
    const PnShellMat mat = S.mat;
    double kb[9], km[4];
    {
        PnGpNode A[4], B[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) { A[g] = s_rec[4 * g + a]; B[g] = s_rec[4 * g + b]; }
        pn_quad_pair(A, B, mat, kb, km);
    }
    if (!QUAD) {
        double pr[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) pr[k] = props[5 * e + k];
        pn_plate_pair_bending(a, b, S.width, S.height, pr, kb);
    }
    if (a == b) S.diag[a] = fmin(fabs(kb[4]), fabs(kb[8]));
    __syncwarp();                                // also: every thread has read its stage-B records
    // drilling stiffness: smallest rotational diagonal / 1000 (Quad3D.py:577-579, Plate3D.py:262-264)
    double k_rz = fmin(fmin(S.diag[0], S.diag[1]), fmin(S.diag[2], S.diag[3])) / 1000;

    double K[36], Kg[36], R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = S.R[k];
    pn_shell_local_block(kb, km, QUAD, a == b, k_rz, K);
    pn_shell_rotate_block(R, K, Kg);
    double *so = s_buf[le] + (6 * a) * 24 + 6 * b;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) so[i * 24 + j] = Kg[6 * i + j];
    __syncwarp();
    if (valid) {
        double *o = out + e * 576 + t;
        const double *si = s_buf[le] + t;
#pragma unroll 6
        for (int k = 0; k < 36; ++k) o[16 * k] = si[16 * k];
    }
}

// Global fixed-end force vectors of shells for unit-factor pressures of one case:
// out[e][24] = inv(T) fer(pressure[e]) (Quad3D.FER :870-882, Plate3D.FER :510-522).
template <bool QUAD>
__global__ void k_shell_fer(int64_t n, const int32_t *__restrict__ ijmn, const double *__restrict__ xyz,
                            const double *__restrict__ pressure, double *__restrict__ out) {
    int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t e = tid >> 2;
    int a = (int)(tid & 3);
    if (e >= n) return;
    double pq = pressure[e];
    double *o = out + e * 24 + 6 * a;
    if (pq == 0.0) {
        for (int k = 0; k < 6; ++k) o[k] = 0.0;
        return;
    }
    double p[4][3];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        int64_t nd = ijmn[4 * e + c];
        p[c][0] = xyz[3 * nd]; p[c][1] = xyz[3 * nd + 1]; p[c][2] = xyz[3 * nd + 2];
    }
    double R[9], tl[3] = {0, 0, 0}, rl[3] = {0, 0, 0};
    pn_shell_dircos(p[0], p[1], p[3], R);
    if (QUAD) {
        PnShellGeom G;
        pn_quad_local_xy(p, G.xl, G.yl);
        tl[2] = pn_quad_fer_w(G, a, pq);
    } else {
        double width = sqrt((p[0][0] - p[1][0]) * (p[0][0] - p[1][0]) + (p[0][1] - p[1][1]) * (p[0][1] - p[1][1]) +
                            (p[0][2] - p[1][2]) * (p[0][2] - p[1][2]));
        double height = sqrt((p[0][0] - p[3][0]) * (p[0][0] - p[3][0]) + (p[0][1] - p[3][1]) * (p[0][1] - p[3][1]) +
                             (p[0][2] - p[3][2]) * (p[0][2] - p[3][2]));
        double f3[3];
        pn_plate_fer_node(a, width, height, pq, f3);
        tl[2] = f3[0]; rl[0] = f3[1]; rl[1] = f3[2];
    }
    pn_rot_to_global(R, tl, o);
    pn_rot_to_global(R, rl, o + 3);
}

// Rotate per-group local member FER vectors to global axes: out[g][12] = inv(T) fer (Member3D.FER :1080).
__global__ void k_member_fer_to_global(int64_t n_groups, const int32_t *__restrict__ g_member,
                                       const int32_t *__restrict__ ij, const double *__restrict__ rot,
                                       const double *__restrict__ xyz, const double *__restrict__ loc,
                                       double *__restrict__ glob) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    int e = g_member[g];
    int i = ij[2 * e], j = ij[2 * e + 1];
    double pi[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    double pj[3] = {xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]};
    double R[9];
    pn_line_dircos(pi, pj, rot[e], R);
    for (int b = 0; b < 4; ++b) {
        double v[3] = {loc[g * 12 + 3 * b], loc[g * 12 + 3 * b + 1], loc[g * 12 + 3 * b + 2]};
        pn_rot_to_global(R, v, glob + g * 12 + 3 * b);
    }
}

// ------------------------------------------------------------------------------------------------
// element end forces for one combo (Member3D.f/F :872-918/:1072, Spring3D.F :200, Quad3D.F :798,
// Plate3D.F :403): F = Ke_global D_e + FER_global, where Ke_global is the stored element matrix.
// In local axes f = T F.  `fer` may be null (no loads on this element type).
// ------------------------------------------------------------------------------------------------
template <int ND>
__global__ void k_element_forces(int64_t n, const int32_t *__restrict__ list /* element ids or null = all */,
                                 const int32_t *__restrict__ conn, const double *__restrict__ ke,
                                 const double *__restrict__ D, const double *__restrict__ fer_case /* or null */,
                                 int n_case, const double *__restrict__ fac_row, int64_t case_stride,
                                 const uint8_t *__restrict__ active, int inactive_quirk, double *__restrict__ out,
                                 int64_t d_stride, int64_t out_stride) {
    // thread = (element, row); the fixed-end term is combined from the per-case vectors on the fly.
    // grid.y = combination of a batch: D, the factor row and the output advance by their strides.
    D += blockIdx.y * d_stride;
    fac_row += blockIdx.y * (int64_t)n_case;
    out += blockIdx.y * out_stride;
    int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t q = tid / ND;
    int r = (int)(tid % ND);
    if (q >= n) return;
    int64_t e = list ? list[q] : q;
    constexpr int NN = ND / 6;
    // Inactive springs / members: excluded from the reactions (Analysis.py:1100, 1139).  Their own F() / f() in the
    // reference still evaluate ke (T D) + fer with only the GLOBAL DX of both nodes dropped from D (Spring3D.D :222-225,
    // Member3D.D :1110-1112) -- reproduced when `inactive_quirk` is set (the user-facing end forces).
    const bool off = active && !active[e];
    if (off && !inactive_quirk) { out[e * ND + r] = 0.0; return; }
    const double *row = ke + e * (int64_t)(ND * ND) + (int64_t)r * ND;
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < NN; ++a) {
        int64_t nd = conn[NN * e + a];
#pragma unroll
        for (int k = 0; k < 6; ++k) acc = fma(row[6 * a + k], (off && k == 0) ? 0.0 : D[6 * nd + k], acc);
    }
    if (fer_case) {
        double f = 0.0;
        for (int cs = 0; cs < n_case; ++cs) f = fma(fac_row[cs], fer_case[(int64_t)cs * case_stride + e * ND + r], f);
        acc += f;
    }
    out[e * ND + r] = acc;
}

// Adds kg(P) d in global axes to member end forces (P-Delta branch of Member3D.f :886-891).
__global__ void k_member_pdelta_forces(int64_t n, const int32_t *__restrict__ ij, const double *__restrict__ props,
                                       const double *__restrict__ rot, const uint16_t *__restrict__ rel,
                                       const double *__restrict__ xyz, const double *__restrict__ D,
                                       const uint8_t *__restrict__ active, const int32_t *__restrict__ list,
                                       double *__restrict__ F) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    int64_t e = list ? list[q] : q;
    if (active && !active[e]) return;
    int i = ij[2 * e], j = ij[2 * e + 1];
    double pi[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    double pj[3] = {xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]};
    double R[9], d[12], k[144], pr[6];
    for (int c = 0; c < 6; ++c) pr[c] = props[6 * e + c];
    double L = pn_line_dircos(pi, pj, rot[e], R);
    for (int b = 0; b < 4; ++b) {
        int64_t nd = b < 2 ? i : j;
        const double *g = D + 6 * nd + 3 * (b & 1);
        double gl[3] = {g[0], g[1], g[2]};
        pn_rot_to_local(R, gl, d + 3 * b);
    }
    double P = (d[6] - d[0]) * pr[2] * pr[0] / L;
    pn_member_kg(pr, L, rel[e], P, k);
    double f[12];
    for (int r = 0; r < 12; ++r) {
        double acc = 0.0;
        for (int c = 0; c < 12; ++c) acc = fma(k[r * 12 + c], d[c], acc);
        f[r] = acc;
    }
    for (int b = 0; b < 4; ++b) {
        double g[3];
        pn_rot_to_global(R, f + 3 * b, g);
        F[e * 12 + 3 * b] += g[0]; F[e * 12 + 3 * b + 1] += g[1]; F[e * 12 + 3 * b + 2] += g[2];
    }
}

// F_local = T F_global, per element (ND = 12 or 24), in place.
template <int ND>
__global__ void k_forces_to_local(int64_t n, const int32_t *__restrict__ conn, const double *__restrict__ rot,
                                  const double *__restrict__ xyz, int shell, double *__restrict__ F) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n) return;
    constexpr int NN = ND / 6;
    double R[9];
    if (shell) {
        double p[3][3];
        int idx[3] = {0, 1, 3};
        for (int a = 0; a < 3; ++a) {
            int64_t nd = conn[NN * e + idx[a]];
            p[a][0] = xyz[3 * nd]; p[a][1] = xyz[3 * nd + 1]; p[a][2] = xyz[3 * nd + 2];
        }
        pn_shell_dircos(p[0], p[1], p[2], R);
    } else {
        int64_t i = conn[2 * e], j = conn[2 * e + 1];
        double pi[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        double pj[3] = {xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]};
        pn_line_dircos(pi, pj, rot ? rot[e] : 0.0, R);
    }
    for (int b = 0; b < ND / 3; ++b) {
        double g[3] = {F[e * ND + 3 * b], F[e * ND + 3 * b + 1], F[e * ND + 3 * b + 2]}, l[3];
        pn_rot_to_local(R, g, l);
        F[e * ND + 3 * b] = l[0]; F[e * ND + 3 * b + 1] = l[1]; F[e * ND + 3 * b + 2] = l[2];
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
static inline unsigned grid_for(int64_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

void launch_element_ke_arrays(Ctx *c, const ElemInputs &in, double *out_spring, double *out_member, double *out_quad,
                              double *out_plate, cudaStream_t st) {
    if (in.n_springs) {
        { ProfScope ps_(c, PS_ELEM_KE); k_spring_ke<<<grid_for(in.n_springs, 64), 64, 0, st>>>(in.n_springs, in.spring_ij, in.spring_ks, in.xyz, out_spring); }
        c->count_launch();
    }
    if (in.n_members) {
        { ProfScope ps_(c, PS_ELEM_KE); k_member_matrix<<<grid_for(in.n_members, 64), 64, 0, st>>>(in.n_members, in.mem_ij, in.mem_props, in.mem_rot, in.mem_rel, in.xyz, 0, nullptr, out_member); }
        c->count_launch();
    }
    if (in.n_quads) {
        { ProfScope ps_(c, PS_ELEM_KE); k_shell_ke<true><<<grid_for(in.n_quads, SHELL_EPB), SHELL_TPB, 0, st>>>(in.n_quads, in.quad_ijmn, in.quad_props, in.xyz, out_quad); }
        c->count_launch();
    }
    if (in.n_plates) {
        { ProfScope ps_(c, PS_ELEM_KE); k_shell_ke<false><<<grid_for(in.n_plates, SHELL_EPB), SHELL_TPB, 0, st>>>(in.n_plates, in.plate_ijmn, in.plate_props, in.xyz, out_plate); }
        c->count_launch();
    }
}

void launch_element_ke(Ctx *c, cudaStream_t st) {
    const Model &m = c->model;
    ElemInputs in{m.n_springs, m.n_members, m.n_quads, m.n_plates, m.spring_ij.ptr, m.spring_ks.ptr, m.mem_ij.ptr,
                  m.mem_props.ptr, m.mem_rot.ptr, m.mem_rel.ptr, m.quad_ijmn.ptr, m.quad_props.ptr, m.plate_ijmn.ptr,
                  m.plate_props.ptr, m.xyz.ptr};
    double *ev = c->ev.ptr;
    launch_element_ke_arrays(c, in, ev + c->ev_off_spring, ev + c->ev_off_member, ev + c->ev_off_quad, ev + c->ev_off_plate, st);
}

void launch_member_kg(Ctx *c, const double *P_dev, double *out_dev, cudaStream_t st) {
    const Model &m = c->model;
    if (!m.n_members) return;
    k_member_matrix<<<grid_for(m.n_members, 64), 64, 0, st>>>(m.n_members, m.mem_ij.ptr, m.mem_props.ptr, m.mem_rot.ptr,
                                                               m.mem_rel.ptr, m.xyz.ptr, 1, P_dev, out_dev);
    c->count_launch();
}

void launch_member_axial(Ctx *c, const double *D_dev, double *P_dev, cudaStream_t st) {
    const Model &m = c->model;
    if (!m.n_members) return;
    k_member_axial<<<grid_for(m.n_members, 128), 128, 0, st>>>(m.n_members, m.mem_ij.ptr, m.mem_props.ptr,
                                                                m.mem_rot.ptr, m.xyz.ptr, D_dev, P_dev);
    c->count_launch();
}

void launch_member_fer_groups(Ctx *c, cudaStream_t st) {
    const Model &m = c->model;
    Loads &ld = c->loads;
    if (!ld.n_mgroups) return;
    k_member_fer_groups<<<grid_for(ld.n_mgroups, 64), 64, 0, st>>>(
        ld.n_mgroups, ld.mg_member.ptr, ld.mg_start.ptr, ld.ml_kind.ptr, ld.ml_params.ptr, m.mem_ij.ptr,
        m.mem_props.ptr, m.mem_rot.ptr, m.mem_rel.ptr, m.xyz.ptr, ld.mg_fer_local.ptr);
    c->count_launch();
    k_member_fer_to_global<<<grid_for(ld.n_mgroups, 64), 64, 0, st>>>(ld.n_mgroups, ld.mg_member.ptr, m.mem_ij.ptr,
                                                                       m.mem_rot.ptr, m.xyz.ptr, ld.mg_fer_local.ptr,
                                                                       ld.mg_fer_global.ptr);
    c->count_launch();
}

void launch_shell_fer(Ctx *c, bool quad, const double *pressure_dev, double *out_dev, cudaStream_t st) {
    const Model &m = c->model;
    int64_t n = quad ? m.n_quads : m.n_plates;
    if (!n) return;
    if (quad)
        k_shell_fer<true><<<grid_for(4 * n, 128), 128, 0, st>>>(n, m.quad_ijmn.ptr, m.xyz.ptr, pressure_dev, out_dev);
    else
        k_shell_fer<false><<<grid_for(4 * n, 128), 128, 0, st>>>(n, m.plate_ijmn.ptr, m.xyz.ptr, pressure_dev, out_dev);
    c->count_launch();
}

// combo: fixed-end vectors are combined with that combo's factors; supported_only restricts the work to the
// elements that touch a supported node (all that Analysis._calc_reactions :1081-1267 needs).
void launch_element_forces(Ctx *c, int kind, const double *D_dev, int32_t combo, bool supported_only, double *out_dev,
                           cudaStream_t st, int n_batch, int64_t out_stride) {
    // n_batch > 1: combinations combo .. combo + n_batch - 1 in one launch; D_dev is combo's vector inside D_all
    // (stride n_dof), out_dev its output block (stride out_stride)
    const int64_t d_stride = c->n_dof;
    const unsigned gy = (unsigned)std::max(n_batch, 1);
    const Model &m = c->model;
    const Loads &ld = c->loads;
    const double *ev = c->ev.ptr;
    const double *fac = ld.factors.ptr + (int64_t)combo * ld.n_case;
    const double *ferc = ld.fer_case.n ? ld.fer_case.ptr : nullptr;
    int k = kind - 1;
    const int32_t *list = supported_only ? c->sup_list[k].ptr : nullptr;
    ProfScope ps_(c, PS_REACTIONS);
    switch (kind) {
    case PN_SPRING: {
        int64_t n = supported_only ? c->sup_list[k].n : m.n_springs;
        if (n) {
            k_element_forces<12><<<dim3(grid_for(n * 12, 96), gy), 96, 0, st>>>(n, list, m.spring_ij.ptr, ev + c->ev_off_spring, D_dev,
                                                                       nullptr, ld.n_case, fac, 0, m.spring_active.ptr, supported_only ? 0 : 1, out_dev,
                                                                       d_stride, out_stride);
            c->count_launch();
        }
        break;
    }
    case PN_MEMBER: {
        int64_t n = supported_only ? c->sup_list[k].n : m.n_members;
        if (n) {
            k_element_forces<12><<<dim3(grid_for(n * 12, 96), gy), 96, 0, st>>>(
                n, list, m.mem_ij.ptr, ev + c->ev_off_member, D_dev, ferc ? ferc + c->ef_off_member : nullptr, ld.n_case, fac,
                c->ef_count, m.mem_active.ptr, supported_only ? 0 : 1, out_dev, d_stride, out_stride);
            c->count_launch();
        }
        break;
    }
    case PN_QUAD: {
        int64_t n = supported_only ? c->sup_list[k].n : m.n_quads;
        if (n) {
            k_element_forces<24><<<dim3(grid_for(n * 24, 96), gy), 96, 0, st>>>(
                n, list, m.quad_ijmn.ptr, ev + c->ev_off_quad, D_dev, ferc ? ferc + c->ef_off_quad : nullptr, ld.n_case, fac,
                c->ef_count, nullptr, 0, out_dev, d_stride, out_stride);
            c->count_launch();
        }
        break;
    }
    case PN_PLATE: {
        int64_t n = supported_only ? c->sup_list[k].n : m.n_plates;
        if (n) {
            k_element_forces<24><<<dim3(grid_for(n * 24, 96), gy), 96, 0, st>>>(
                n, list, m.plate_ijmn.ptr, ev + c->ev_off_plate, D_dev, ferc ? ferc + c->ef_off_plate : nullptr, ld.n_case, fac,
                c->ef_count, nullptr, 0, out_dev, d_stride, out_stride);
            c->count_launch();
        }
        break;
    }
    }
}

void launch_member_pdelta_forces(Ctx *c, const double *D_dev, bool supported_only, double *F_dev, cudaStream_t st) {
    const Model &m = c->model;
    int64_t n = supported_only ? c->sup_list[PN_MEMBER - 1].n : m.n_members;
    if (!n) return;
    ProfScope ps_(c, PS_REACTIONS);
    k_member_pdelta_forces<<<grid_for(n, 64), 64, 0, st>>>(n, m.mem_ij.ptr, m.mem_props.ptr, m.mem_rot.ptr, m.mem_rel.ptr,
                                                            m.xyz.ptr, D_dev, m.mem_active.ptr,
                                                            supported_only ? c->sup_list[PN_MEMBER - 1].ptr : nullptr, F_dev);
    c->count_launch();
}

void launch_forces_to_local(Ctx *c, int kind, double *F_dev, cudaStream_t st) {
    const Model &m = c->model;
    switch (kind) {
    case PN_SPRING:
        if (m.n_springs) {
            k_forces_to_local<12><<<grid_for(m.n_springs, 128), 128, 0, st>>>(m.n_springs, m.spring_ij.ptr, nullptr,
                                                                               m.xyz.ptr, 0, F_dev);
            c->count_launch();
        }
        break;
    case PN_MEMBER:
        if (m.n_members) {
            k_forces_to_local<12><<<grid_for(m.n_members, 128), 128, 0, st>>>(m.n_members, m.mem_ij.ptr, m.mem_rot.ptr,
                                                                               m.xyz.ptr, 0, F_dev);
            c->count_launch();
        }
        break;
    case PN_QUAD:
        if (m.n_quads) {
            k_forces_to_local<24><<<grid_for(m.n_quads, 128), 128, 0, st>>>(m.n_quads, m.quad_ijmn.ptr, nullptr,
                                                                             m.xyz.ptr, 1, F_dev);
            c->count_launch();
        }
        break;
    case PN_PLATE:
        if (m.n_plates) {
            k_forces_to_local<24><<<grid_for(m.n_plates, 128), 128, 0, st>>>(m.n_plates, m.plate_ijmn.ptr, nullptr,
                                                                              m.xyz.ptr, 1, F_dev);
            c->count_launch();
        }
        break;
    }
}

}  // namespace pn
