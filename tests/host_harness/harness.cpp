// TEST INFRASTRUCTURE ONLY.  Compiles the element math headers of the CUDA library with g++ so the
// arithmetic can be checked against the oracle on a machine without a GPU.  The glue below repeats,
// sequentially, what the 16-threads-per-element kernels of pynite_b200/csrc/pn_elements.cu do.
// Nothing in the product package loads this library.
#include "../../pynite_b200/csrc/elem_line.cuh"
#include "../../pynite_b200/csrc/elem_shell.cuh"

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
}
