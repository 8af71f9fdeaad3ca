// Batched shell force / stress recovery: Quad3D.shear / moment / membrane (Quad3D.py:1017-1239) and
// Plate3D.moment / shear / membrane (Plate3D.py:590-692) for every element of a kind, a range of solved load
// combinations and a list of evaluation points, in one launch.
//   16 threads per element: stage B (Gauss point, node) records go to shared memory exactly as in k_shell_ke; plates
//   additionally build the closed-form inv(C) of their displacement polynomial there (9 entries per thread).  The 16
//   threads then take the combinations round-robin: local displacement vector d = T D_e, the values at the four Gauss
//   points (or the polynomial constants a = inv(C) d), and one 9-component result per evaluation point:
//     local  : Qx Qy 0  | Mx My Mxy | Sx Sy Txy
//     global : Qx Qy Qz | Mx My Mxy | Sx Sy Sxy      (quads only; the reference's plate methods ignore `local`)
#include "pn_internal.h"
#include "elem_stress.cuh"

namespace pn {

constexpr int ST_TPB = 128;
constexpr int ST_EPB = ST_TPB / 16;

template <bool QUAD>
__global__ void __launch_bounds__(ST_TPB)
k_shell_stress(int64_t n, const int32_t *__restrict__ ijmn, const double *__restrict__ props, const double *__restrict__ xyz,
               const double *__restrict__ D /* [n_combo][n_dof] */, int64_t n_dof, int n_combo, int n_pts,
               const double *__restrict__ pts /* [n_pts][2] */, int local, double *__restrict__ out /* [n_combo][n][n_pts][9] */) {
    __shared__ PnGpNode s_rec[ST_EPB][16];
    __shared__ double s_cinv[QUAD ? 1 : ST_EPB][QUAD ? 1 : 144];
    const int le = threadIdx.x >> 4, t = threadIdx.x & 15;
    int64_t e = blockIdx.x * (int64_t)ST_EPB + le;
    const bool valid = e < n;
    if (!valid) e = n - 1;
    double p[4][3];
    int64_t nd[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        nd[a] = ijmn[4 * e + a];
        p[a][0] = xyz[3 * nd[a]]; p[a][1] = xyz[3 * nd[a] + 1]; p[a][2] = xyz[3 * nd[a] + 2];
    }
    double pr[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) pr[k] = props[5 * e + k];
    PnShellGeom G;
    double width = 0.0, height = 0.0;
    if (QUAD) {
        pn_quad_local_xy(p, G.xl, G.yl);
        pn_quad_edges(G, pr[0], pr[2]);
    } else {
        width = sqrt((p[0][0] - p[1][0]) * (p[0][0] - p[1][0]) + (p[0][1] - p[1][1]) * (p[0][1] - p[1][1]) +
                     (p[0][2] - p[1][2]) * (p[0][2] - p[1][2]));
        height = sqrt((p[0][0] - p[3][0]) * (p[0][0] - p[3][0]) + (p[0][1] - p[3][1]) * (p[0][1] - p[3][1]) +
                      (p[0][2] - p[3][2]) * (p[0][2] - p[3][2]));
        G.xl[0] = 0; G.yl[0] = 0; G.xl[1] = width; G.yl[1] = 0;
        G.xl[2] = width; G.yl[2] = height; G.xl[3] = 0; G.yl[3] = height;
        double pw[7], ph[7];
        pn_plate_powers(width, pw);
        pn_plate_powers(height, ph);
        for (int k = t; k < 144; k += 16) s_cinv[le][k] = pn_plate_cinv_entry(k, pw, ph);
    }
    pn_shell_gp_node(G, t >> 2, t & 3, QUAD, s_rec[le][t]);
    __syncthreads();
    if (!valid) return;
    PnShellMat mat;
    pn_shell_material(pr, mat);
    double R[9];
    pn_shell_dircos(p[0], p[1], p[3], R);
    for (int j = t; j < n_combo; j += 16) {
        const double *Dj = D + (int64_t)j * n_dof;
        double d[24];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double *g = Dj + 6 * nd[a];
            double tr[3] = {g[0], g[1], g[2]}, ro[3] = {g[3], g[4], g[5]};
            pn_rot_to_local(R, tr, d + 6 * a);
            pn_rot_to_local(R, ro, d + 6 * a + 3);
        }
        double V[4][8];
        pn_shell_gp_values(s_rec[le], mat, d, QUAD, V);
        double a12[12];
        if (!QUAD) {
            // a = inv(C) d_b with d_b = (dz, rx, ry) per node (Plate3D._a :573-588)
            for (int i = 0; i < 12; ++i) {
                double acc = 0.0;
                for (int q = 0; q < 12; ++q) acc = fma(s_cinv[le][12 * i + q], d[6 * (q / 3) + 2 + q % 3], acc);
                a12[i] = acc;
            }
        }
        double *o = out + (((int64_t)j * n + e) * n_pts) * 9;
        for (int q = 0; q < n_pts; ++q) {
            const double u = pts[2 * q], v = pts[2 * q + 1];
            double loc[8], res[9];
            if (QUAD) {
                pn_shell_extrapolate(V, u, v, 0, 8, loc);
                if (local) {
                    res[0] = loc[0]; res[1] = loc[1]; res[2] = 0.0;
                    for (int k = 2; k < 8; ++k) res[k + 1] = loc[k];
                } else {
                    pn_quad_stress_global(R, loc, res);
                }
            } else {
                // points are local (x, y); membrane in natural coordinates r = -1 + 2x / width, s = -1 + 2y / height (:653-654)
                pn_shell_extrapolate(V, -1 + 2 * u / width, -1 + 2 * v / height, 5, 8, loc);
                double b5[5];
                pn_plate_bending_results(pr, a12, u, v, b5);
                res[0] = b5[0]; res[1] = b5[1]; res[2] = 0.0; res[3] = b5[2]; res[4] = b5[3]; res[5] = b5[4];
                res[6] = loc[5]; res[7] = loc[6]; res[8] = loc[7];
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) o[9 * q + k] = res[k];
        }
    }
}

void launch_shell_stress(Ctx *c, bool quad, const double *D_dev, int n_combo, int n_pts, const double *pts_dev, int local,
                         double *out_dev, cudaStream_t st) {
    const Model &m = c->model;
    const int64_t n = quad ? m.n_quads : m.n_plates;
    if (!n || !n_combo || !n_pts) return;
    const unsigned grid = (unsigned)((n + ST_EPB - 1) / ST_EPB);
    ProfScope ps_(c, PS_REACTIONS);
    if (quad)
        k_shell_stress<true><<<grid, ST_TPB, 0, st>>>(n, m.quad_ijmn.ptr, m.quad_props.ptr, m.xyz.ptr, D_dev, c->n_dof, n_combo, n_pts,
                                                       pts_dev, local, out_dev);
    else
        k_shell_stress<false><<<grid, ST_TPB, 0, st>>>(n, m.plate_ijmn.ptr, m.plate_props.ptr, m.xyz.ptr, D_dev, c->n_dof, n_combo,
                                                        n_pts, pts_dev, local, out_dev);
    c->count_launch();
}

}  // namespace pn
