// TEST INFRASTRUCTURE ONLY.  Compiles the element math headers of the CUDA library with g++ so the
// arithmetic can be checked against the oracle on a machine without a GPU.  The glue below repeats,
// sequentially, what the 16-threads-per-element kernels of pynite_b200/csrc/pn_elements.cu do.
// Nothing in the product package loads this library.
#include "../../pynite_b200/csrc/elem_line.cuh"
#include "../../pynite_b200/csrc/elem_shell.cuh"
#include "../../pynite_b200/csrc/elem_stress.cuh"

extern "C" {

// global 12x12 Ke (or Kg when P_or_nan is finite... see `mode`): mode 0 = Ke, 1 = Kg(P), 2 = spring
void hh_line_matrix(const double *pi, const double *pj, const double *props, double rot, unsigned rel,
                    int mode, double P, double *out) {
    double R[9], k[144];
    double L = pn_line_dircos(pi, pj, mode == 2 ? 0.0 : rot, R);
    if (mode == 0) pn_member_ke(props, L, rel, k);
    else if (mode == 1) pn_member_kg(props, L, rel, P, k);
    else { pn_zero144(k); k[0] = props[0]; k[78] = props[0]; k[6] = -props[0]; k[72] = -props[0]; }
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            double B[9], G[9];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) B[3*i+j] = k[(3*a+i)*12 + 3*b+j];
            pn_congruence3(R, B, G);
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out[(3*a+i)*12 + 3*b+j] = G[3*i+j];
        }
}

// condensed local FER of a member for n loads with per-load factors
void hh_member_fer(const double *pi, const double *pj, const double *props, double rot, unsigned rel,
                   int n, const int *kind, const double *params, const double *factor, double *out12) {
    double R[9], k[144], f[12];
    double L = pn_line_dircos(pi, pj, rot, R);
    for (int i = 0; i < 12; ++i) f[i] = 0.0;
    for (int l = 0; l < n; ++l) {
        double g[12];
        for (int i = 0; i < 12; ++i) g[i] = 0.0;
        pn_member_load_fer(kind[l], params + 4*l, R, L, g);
        for (int i = 0; i < 12; ++i) f[i] += factor[l]*g[i];
    }
    pn_member_ke_unc(props[0], props[1], props[2], props[3], props[4], props[5], L, k);
    pn_condense(k, f, rel);
    for (int i = 0; i < 12; ++i) out12[i] = f[i];
}

// global 24x24 Ke of a quad (is_quad = 1) or plate (0); p = 4 corner points
void hh_shell_matrix(const double *pts, const double *props, int is_quad, double *out) {
    double p[4][3];
    for (int a = 0; a < 4; ++a) for (int k = 0; k < 3; ++k) p[a][k] = pts[3*a+k];
    double R[9];
    pn_shell_dircos(p[0], p[1], p[3], R);
    PnShellMat mat;
    pn_shell_material(props, mat);
    PnShellGeom G;
    double width = 0, height = 0;
    if (is_quad) {
        pn_quad_local_xy(p, G.xl, G.yl);
        pn_quad_edges(G, props[0], props[2]);
    } else {
        double d;
        d = 0; for (int k = 0; k < 3; ++k) d += (p[0][k]-p[1][k])*(p[0][k]-p[1][k]); width = sqrt(d);
        d = 0; for (int k = 0; k < 3; ++k) d += (p[0][k]-p[3][k])*(p[0][k]-p[3][k]); height = sqrt(d);
        G.xl[0] = 0; G.yl[0] = 0; G.xl[1] = width; G.yl[1] = 0; G.xl[2] = width; G.yl[2] = height; G.xl[3] = 0; G.yl[3] = height;
    }
    PnGpNode rec[4][4];
    for (int g = 0; g < 4; ++g) for (int a = 0; a < 4; ++a) pn_shell_gp_node(G, g, a, is_quad != 0, rec[g][a]);
    double kb[4][4][9], km[4][4][4], dmin = 1e300;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            PnGpNode A[4], B[4];
            for (int g = 0; g < 4; ++g) { A[g] = rec[g][a]; B[g] = rec[g][b]; }
            pn_quad_pair(A, B, mat, kb[a][b], km[a][b]);
            if (!is_quad) pn_plate_pair_bending(a, b, width, height, props, kb[a][b]);
            if (a == b) { dmin = fmin(dmin, fmin(fabs(kb[a][b][4]), fabs(kb[a][b][8]))); }
        }
    double k_rz = dmin/1000;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            double K[36], Kg[36];
            pn_shell_local_block(kb[a][b], km[a][b], is_quad != 0, a == b, k_rz, K);
            pn_shell_rotate_block(R, K, Kg);
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) out[(6*a+i)*24 + 6*b+j] = Kg[6*i+j];
        }
}

// global 24-vector FER of a shell for a (factored) pressure
void hh_shell_fer(const double *pts, const double *props, int is_quad, double pressure, double *out24) {
    double p[4][3];
    for (int a = 0; a < 4; ++a) for (int k = 0; k < 3; ++k) p[a][k] = pts[3*a+k];
    double R[9];
    pn_shell_dircos(p[0], p[1], p[3], R);
    for (int a = 0; a < 4; ++a) {
        double tl[3] = {0, 0, 0}, rl[3] = {0, 0, 0};
        if (is_quad) {
            PnShellGeom G;
            pn_quad_local_xy(p, G.xl, G.yl);
            tl[2] = pn_quad_fer_w(G, a, pressure);
        } else {
            double d, width, height, f3[3];
            d = 0; for (int k = 0; k < 3; ++k) d += (p[0][k]-p[1][k])*(p[0][k]-p[1][k]); width = sqrt(d);
            d = 0; for (int k = 0; k < 3; ++k) d += (p[0][k]-p[3][k])*(p[0][k]-p[3][k]); height = sqrt(d);
            pn_plate_fer_node(a, width, height, pressure, f3);
            tl[2] = f3[0]; rl[0] = f3[1]; rl[1] = f3[2];
        }
        pn_rot_to_global(R, tl, out24 + 6*a);
        pn_rot_to_global(R, rl, out24 + 6*a + 3);
    }
}

// Internal shears / moments / membrane stresses of one shell (sequential restatement of k_shell_stress in
// pynite_b200/csrc/pn_stress.cu): D24 = global displacements of the four nodes, pts[n_pts][2], out[n_pts][9].
void hh_shell_stress(const double *pts12, const double *props, int is_quad, const double *D24, int n_pts, const double *pts,
                     int local, double *out) {
    double p[4][3];
    for (int a = 0; a < 4; ++a) for (int k = 0; k < 3; ++k) p[a][k] = pts12[3*a+k];
    double R[9];
    pn_shell_dircos(p[0], p[1], p[3], R);
    PnShellMat mat;
    pn_shell_material(props, mat);
    PnShellGeom G;
    double width = 0, height = 0, cinv[144];
    if (is_quad) {
        pn_quad_local_xy(p, G.xl, G.yl);
        pn_quad_edges(G, props[0], props[2]);
    } else {
        double d;
        d = 0; for (int k = 0; k < 3; ++k) d += (p[0][k]-p[1][k])*(p[0][k]-p[1][k]); width = sqrt(d);
        d = 0; for (int k = 0; k < 3; ++k) d += (p[0][k]-p[3][k])*(p[0][k]-p[3][k]); height = sqrt(d);
        G.xl[0] = 0; G.yl[0] = 0; G.xl[1] = width; G.yl[1] = 0; G.xl[2] = width; G.yl[2] = height; G.xl[3] = 0; G.yl[3] = height;
        double pw[7], ph[7];
        pn_plate_powers(width, pw);
        pn_plate_powers(height, ph);
        for (int k = 0; k < 144; ++k) cinv[k] = pn_plate_cinv_entry(k, pw, ph);
    }
    PnGpNode rec[16];
    for (int t = 0; t < 16; ++t) pn_shell_gp_node(G, t >> 2, t & 3, is_quad != 0, rec[t]);
    double d[24];
    for (int a = 0; a < 4; ++a) {
        double tr[3] = {D24[6*a], D24[6*a+1], D24[6*a+2]}, ro[3] = {D24[6*a+3], D24[6*a+4], D24[6*a+5]};
        pn_rot_to_local(R, tr, d + 6*a);
        pn_rot_to_local(R, ro, d + 6*a + 3);
    }
    double V[4][8], a12[12];
    pn_shell_gp_values(rec, mat, d, is_quad != 0, V);
    if (!is_quad)
        for (int i = 0; i < 12; ++i) {
            double acc = 0.0;
            for (int q = 0; q < 12; ++q) acc = fma(cinv[12*i + q], d[6*(q/3) + 2 + q % 3], acc);
            a12[i] = acc;
        }
    for (int q = 0; q < n_pts; ++q) {
        const double u = pts[2*q], v = pts[2*q+1];
        double loc[8], res[9];
        if (is_quad) {
            pn_shell_extrapolate(V, u, v, 0, 8, loc);
            if (local) { res[0] = loc[0]; res[1] = loc[1]; res[2] = 0.0; for (int k = 2; k < 8; ++k) res[k+1] = loc[k]; }
            else pn_quad_stress_global(R, loc, res);
        } else {
            pn_shell_extrapolate(V, -1 + 2*u/width, -1 + 2*v/height, 5, 8, loc);
            double b5[5];
            pn_plate_bending_results(props, a12, u, v, b5);
            res[0] = b5[0]; res[1] = b5[1]; res[2] = 0.0; res[3] = b5[2]; res[4] = b5[3]; res[5] = b5[4];
            res[6] = loc[5]; res[7] = loc[6]; res[8] = loc[7];
        }
        for (int k = 0; k < 9; ++k) out[9*q + k] = res[k];
    }
}
}
