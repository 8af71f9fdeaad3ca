// Four-node shells: DKMQ quadrilateral (Quad3D) and 12-term rectangular plate (Plate3D).
// The math is organised for the 16-threads-per-element kernels in pn_elements.cu:
//   stage B: thread (gauss point g, node a) builds that node's columns of B_b, B_s, B_m at g
//   stage C: thread (node a, node b) contracts them into the local 6x6 block K_ab, applies the
//            reference's slot mapping / sign flips, and rotates it to global axes.
// Reference behaviour restated from Quad3D.py:107-184 (local coords, edges, phi_k), :276-465
// (J, B_b, B_s, B_m), :467-521 (Hb, Hs, Cm), :523-711 (ke_b, ke_m, ke), :721-788 (fer), :884-950 (T)
// and Plate3D.py:89-227 (Dm, J, B_m, ke_m), :229-314 (ke_b), :324-393 (fer), :451-500 (T).
#pragma once
#include "pn_common.cuh"

#define PN_GP 0.57735026918962576451   // 1/sqrt(3)

// Rows x, y, z of the shell triad: x along i->j, z = x cross (i->n), y = z cross x.
PN_HD void pn_shell_dircos(const double *pi, const double *pj, const double *pn, double *R) {
    double x[3] = {pj[0] - pi[0], pj[1] - pi[1], pj[2] - pi[2]};
    double n = pn_norm3(x);
    x[0] /= n; x[1] /= n; x[2] /= n;
    double xy[3] = {pn[0] - pi[0], pn[1] - pi[1], pn[2] - pi[2]};
    double z[3], y[3];
    pn_cross(x, xy, z);
    n = pn_norm3(z);
    z[0] /= n; z[1] /= n; z[2] /= n;
    pn_cross(z, x, y);
    n = pn_norm3(y);
    y[0] /= n; y[1] /= n; y[2] /= n;
    for (int k = 0; k < 3; ++k) { R[k] = x[k]; R[3 + k] = y[k]; R[6 + k] = z[k]; }
}

struct PnShellGeom {
    double xl[4], yl[4];    // in-plane corner coordinates
    double Lk[4], Ck[4], Sk[4];   // edge k: node k -> node (k+1)%4
    double aD[4];           // -3/2 / (1 + phi_k)        (A_Delta_inv_DKMQ, Quad3D.py:327-337)
    double gS[4];           // A_gamma_k * phi_k/(1+phi_k) (A_gamma :294-305, A_phi_Delta :339-349)
};

struct PnShellMat {
    double Db, nu;          // bending rigidity E t^3 / (12 (1 - nu^2)), Poisson ratio
    double Hs;              // shear rigidity  E t kappa / (2 (1 + nu)), kappa = 5/6
    double c11, c12, c21, c22, c33;   // membrane matrix (orthotropic through kx_mod, ky_mod)
    double t;
};

PN_HD void pn_shell_material(const double *props, PnShellMat &m) {
    double t = props[0], E = props[1], nu = props[2], kx = props[3], ky = props[4];
    m.t = t; m.nu = nu;
    m.Db = E * (t * t * t) / (12 * (1 - nu * nu));
    m.Hs = E * t * (5.0 / 6.0) / (2 * (1 + nu));
    double Ex = E * kx, Ey = E * ky, G = E / (2 * (1 + nu));
    double inv = 1 / (1 - nu * nu);
    m.c11 = inv * Ex; m.c12 = inv * (nu * Ex); m.c21 = inv * (nu * Ey); m.c22 = inv * Ey;
    m.c33 = inv * ((1 - nu * nu) * G);
}

// Local corner coordinates of a quad (Quad3D._local_coords :107-141): x along 1->2,
// z = x cross (1->3), y = z cross x; node 1 at the origin.
PN_HD void pn_quad_local_xy(const double p[4][3], double *xl, double *yl) {
    double v12[3], v13[3], v14[3];
    for (int k = 0; k < 3; ++k) { v12[k] = p[1][k] - p[0][k]; v13[k] = p[2][k] - p[0][k]; v14[k] = p[3][k] - p[0][k]; }
    double xa[3] = {v12[0], v12[1], v12[2]}, za[3], ya[3];
    pn_cross(xa, v13, za);
    pn_cross(za, xa, ya);
    double nx = pn_norm3(xa), ny = pn_norm3(ya);
    for (int k = 0; k < 3; ++k) { xa[k] /= nx; ya[k] /= ny; }
    xl[0] = 0.0; yl[0] = 0.0;
    xl[1] = v12[0] * xa[0] + v12[1] * xa[1] + v12[2] * xa[2];
    xl[2] = v13[0] * xa[0] + v13[1] * xa[1] + v13[2] * xa[2];
    xl[3] = v14[0] * xa[0] + v14[1] * xa[1] + v14[2] * xa[2];
    yl[1] = v12[0] * ya[0] + v12[1] * ya[1] + v12[2] * ya[2];
    yl[2] = v13[0] * ya[0] + v13[1] * ya[1] + v13[2] * ya[2];
    yl[3] = v14[0] * ya[0] + v14[1] * ya[1] + v14[2] * ya[2];
}

// Edge data for the DKMQ formulation (L_k :143, dir_cos :157, phi_k :179), one edge / all four.
PN_HD void pn_quad_edge(PnShellGeom &g, int k, double t, double nu) {
    const double kappa = 5.0 / 6.0;
    int b = (k + 1) & 3;
    double dx = g.xl[b] - g.xl[k], dy = g.yl[b] - g.yl[k];
    double L = sqrt(dx * dx + dy * dy);
    g.Lk[k] = L; g.Ck[k] = dx / L; g.Sk[k] = dy / L;
    double r = t / L;
    double phi = 2 / (kappa * (1 - nu)) * (r * r);
    g.aD[k] = -1.5 * (1 / (1 + phi));
    double ag = (k < 2) ? L / 2 : -L / 2;
    g.gS[k] = ag * (phi / (1 + phi));
}
PN_HD void pn_quad_edges(PnShellGeom &g, double t, double nu) {
    for (int k = 0; k < 4; ++k) pn_quad_edge(g, k, t, nu);
}

// Jacobian of the bilinear map at (r, s): J = [[x_r, y_r], [x_s, y_s]] (Quad3D.J :276-286).
PN_HD void pn_jacobian(const double *xl, const double *yl, double r, double s, double *J) {
    J[0] = 0.25 * (xl[0] * (s - 1) - xl[1] * (s - 1) + xl[2] * (s + 1) - xl[3] * (s + 1));
    J[1] = 0.25 * (yl[0] * (s - 1) - yl[1] * (s - 1) + yl[2] * (s + 1) - yl[3] * (s + 1));
    J[2] = 0.25 * (xl[0] * (r - 1) - xl[1] * (r + 1) + xl[2] * (r + 1) - xl[3] * (r - 1));
    J[3] = 0.25 * (yl[0] * (r - 1) - yl[1] * (r + 1) + yl[2] * (r + 1) - yl[3] * (r - 1));
}

PN_HD void pn_gauss_point(int g, double &r, double &s) {
    // order (-,-), (+,-), (+,+), (-,+) as in Quad3D.ke_b :535-544
    r = (g == 1 || g == 2) ? PN_GP : -PN_GP;
    s = (g >= 2) ? PN_GP : -PN_GP;
}

struct PnGpNode {
    double Bb[9];   // 3 strains x (w, t1, t2) columns of B_b for this node
    double Bs[6];   // 2 shear strains x (w, t1, t2)
    double dN[2];   // dN/dx, dN/dy of the bilinear shape function (membrane B_m, :450-465)
    double detJ;
};

// Stage B for one (gauss point, node).  `dkmq` = false skips the bending/shear part (plates).
PN_HD void pn_shell_gp_node(const PnShellGeom &G, int g, int a, bool dkmq, PnGpNode &o) {
    double r, s;
    pn_gauss_point(g, r, s);
    double J[4];
    pn_jacobian(G.xl, G.yl, r, s, J);
    double det = J[0] * J[3] - J[1] * J[2];
    double i11 = J[3] / det, i12 = -J[1] / det, i21 = -J[2] / det, i22 = J[0] / det;
    o.detJ = det;
    // bilinear shape function derivatives in natural coordinates for node a
    double sr = (a == 0 || a == 3) ? -1.0 : 1.0, ss = (a < 2) ? -1.0 : 1.0;
    double Nr = 0.25 * sr * (1 + ss * s), Ns = 0.25 * ss * (1 + sr * r);
    double Nx = i11 * Nr + i12 * Ns, Ny = i21 * Nr + i22 * Ns;
    o.dN[0] = Nx; o.dN[1] = Ny;
    if (!dkmq) {
        for (int k = 0; k < 9; ++k) o.Bb[k] = 0.0;
        for (int k = 0; k < 6; ++k) o.Bs[k] = 0.0;
        return;
    }

    // B_b_beta columns (Quad3D.py:381-383): strains (kx, ky, kxy) from (t1, t2)
    o.Bb[0] = 0; o.Bb[1] = Nx; o.Bb[2] = 0;
    o.Bb[3] = 0; o.Bb[4] = 0;  o.Bb[5] = Ny;
    o.Bb[6] = 0; o.Bb[7] = Ny; o.Bb[8] = Nx;
    o.Bs[0] = o.Bs[1] = o.Bs[2] = o.Bs[3] = o.Bs[4] = o.Bs[5] = 0;

    // the two edges meeting at node a: e0 = a (node a is its first end), e1 = a-1 (second end)
    for (int side = 0; side < 2; ++side) {
        int e = side == 0 ? a : ((a + 3) & 3);
        // quadratic edge functions P_e and their natural derivatives (Quad3D.py:397-404)
        double Pr, Ps;
        if (e == 0)      { Pr = r * (s - 1);               Ps = 0.5 * (r - 1) * (r + 1); }
        else if (e == 1) { Pr = -0.5 * (s - 1) * (s + 1);  Ps = -s * (r + 1); }
        else if (e == 2) { Pr = -r * (s + 1);              Ps = -0.5 * (r - 1) * (r + 1); }
        else             { Pr = 0.5 * (s - 1) * (s + 1);   Ps = s * (r - 1); }
        double Px = i11 * Pr + i12 * Ps, Py = i21 * Pr + i22 * Ps;
        double C = G.Ck[e], S = G.Sk[e], L = G.Lk[e];
        // A_u row e restricted to node a (Quad3D.py:322-325): 1/2 (-+2/L, C, S)
        double au0 = 0.5 * ((side == 0 ? -2.0 : 2.0) / L), au1 = 0.5 * C, au2 = 0.5 * S;
        // bending: B_b_Delta_beta[:, e] * aD[e] * A_u[e, node a]  (:420-430)
        double bd0 = Px * C * G.aD[e], bd1 = Py * S * G.aD[e], bd2 = (Py * C + Px * S) * G.aD[e];
        o.Bb[0] += bd0 * au0; o.Bb[1] += bd0 * au1; o.Bb[2] += bd0 * au2;
        o.Bb[3] += bd1 * au0; o.Bb[4] += bd1 * au1; o.Bb[5] += bd1 * au2;
        o.Bb[6] += bd2 * au0; o.Bb[7] += bd2 * au1; o.Bb[8] += bd2 * au2;
        // shear: J^-1 N_gamma[:, e] * gS[e] * A_u[e, node a]  (:288-292, :432-438)
        double n0, n1;   // N_gamma column e: edges 0, 2 feed gamma_xi, edges 1, 3 feed gamma_eta
        if (e == 0)      { n0 = 0.5 * (1 - s); n1 = 0; }
        else if (e == 1) { n0 = 0;             n1 = 0.5 * (1 + r); }
        else if (e == 2) { n0 = 0.5 * (1 + s); n1 = 0; }
        else             { n0 = 0;             n1 = 0.5 * (1 - r); }
        double m0 = (i11 * n0 + i12 * n1) * G.gS[e], m1 = (i21 * n0 + i22 * n1) * G.gS[e];
        o.Bs[0] += m0 * au0; o.Bs[1] += m0 * au1; o.Bs[2] += m0 * au2;
        o.Bs[3] += m1 * au0; o.Bs[4] += m1 * au1; o.Bs[5] += m1 * au2;
    }
}

// Stage C, DKMQ: 3x3 bending+shear block kb (freedoms w, t1, t2) and 2x2 membrane block km of the
// node pair (a, b), summed over the four Gauss points.  A[g], B[g] are the stage-B records of
// node a and node b at gauss point g.
PN_HD void pn_quad_pair(const PnGpNode *A, const PnGpNode *B, const PnShellMat &m, double *kb, double *km) {
    for (int i = 0; i < 9; ++i) kb[i] = 0.0;
    km[0] = km[1] = km[2] = km[3] = 0.0;
    double ks[9];
    for (int i = 0; i < 9; ++i) ks[i] = 0.0;
    const double d11 = m.Db, d12 = m.nu * m.Db, d33 = m.Db * ((1 - m.nu) / 2);
    for (int g = 0; g < 4; ++g) {
        const PnGpNode &a = A[g];
        const PnGpNode &b = B[g];
        double dj = a.detJ;
        // H_b B_b(b): rows (kx, ky, kxy)
        double hb[9];
        for (int j = 0; j < 3; ++j) {
            hb[j] = d11 * b.Bb[j] + d12 * b.Bb[3 + j];
            hb[3 + j] = d12 * b.Bb[j] + d11 * b.Bb[3 + j];
            hb[6 + j] = d33 * b.Bb[6 + j];
        }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double v = pn_fma(a.Bb[6 + i], hb[6 + j], pn_fma(a.Bb[3 + i], hb[3 + j], a.Bb[i] * hb[j]));
                kb[3 * i + j] = pn_fma(v, dj, kb[3 * i + j]);
                double w = pn_fma(a.Bs[3 + i], b.Bs[3 + j], a.Bs[i] * b.Bs[j]) * m.Hs;
                ks[3 * i + j] = pn_fma(w, dj, ks[3 * i + j]);
            }
        // membrane: B_m(a)^T C B_m(b), B_m(n) = [[Nx, 0], [0, Ny], [Ny, Nx]]
        double ax = a.dN[0], ay = a.dN[1], bx = b.dN[0], by = b.dN[1];
        double cb00 = m.c11 * bx, cb01 = m.c12 * by;     // row 0 of C B_m(b)
        double cb10 = m.c21 * bx, cb11 = m.c22 * by;     // row 1
        double cb20 = m.c33 * by, cb21 = m.c33 * bx;     // row 2
        km[0] = pn_fma(pn_fma(ay, cb20, ax * cb00), dj, km[0]);
        km[1] = pn_fma(pn_fma(ay, cb21, ax * cb01), dj, km[1]);
        km[2] = pn_fma(pn_fma(ax, cb20, ay * cb10), dj, km[2]);
        km[3] = pn_fma(pn_fma(ax, cb21, ay * cb11), dj, km[3]);
    }
    for (int i = 0; i < 9; ++i) kb[i] += ks[i];      // ke_b = bending sum, then += shear sum (:548-563)
    for (int i = 0; i < 4; ++i) km[i] *= m.t;        // t * sum (:665-668)
}

// Rectangular plate bending block (a, b) from the rational tables (tools/derive_plate_tables.py):
// k = S_i S_j [Dx c/b^3 A1 + Dx b/c^3 A2 + nu Dx/(b c) A3 + Dxy/(b c) A4] / DEN, S = (1, c, b).
#if defined(__CUDACC__)
#define PN_PLATE_TABLE(name) static __constant__ int pn_plate_##name[144]
#else
#define PN_PLATE_TABLE(name) static const int pn_plate_##name[144]
#endif
#include "plate_tables.inc"

PN_D void pn_plate_pair_bending(int a, int b, double width, double height, const double *props, double *kb) {
    double t = props[0], E = props[1], nu = props[2];
    double hb = width / 2, hc = height / 2;
    double G = E / (2 * (1 + nu));
    double Dx = E * (t * t * t) / (12 * (1 - nu * nu));
    double D1 = nu * Dx, Dxy = G * (t * t * t) / 12;
    double f1 = Dx * hc / (hb * hb * hb), f2 = Dx * hb / (hc * hc * hc), f3 = D1 / (hb * hc), f4 = Dxy / (hb * hc);
    double S[3] = {1.0, hc, hb};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            int idx = (3 * a + i) * 12 + 3 * b + j;
            double core = (f1 * pn_plate_A1[idx] + f2 * pn_plate_A2[idx] + f3 * pn_plate_A3[idx] +
                           f4 * pn_plate_A4[idx]) / PN_PLATE_TABLE_DEN;
            kb[3 * i + j] = core * (S[i] * S[j]);
        }
}

// Places the 3x3 bending block and 2x2 membrane block into the local 6x6 node-pair block
// (slots dx, dy, dz, rx, ry, rz) and adds the drilling spring on the diagonal pair.
//   quad : w -> dz, t1 -> ry (+), t2 -> rx (-)   (expansion :586-615, flip :624-625, swap :630-631)
//   plate: w -> dz, theta_x -> rx, theta_y -> ry (Plate3D.py:276-305)
PN_HD void pn_shell_local_block(const double *kb, const double *km, bool quad, bool diag, double k_rz, double *K) {
    for (int i = 0; i < 36; ++i) K[i] = 0.0;
    K[0] = km[0]; K[1] = km[1]; K[6] = km[2]; K[7] = km[3];
    int slot[3];
    double sg[3];
    slot[0] = 2; sg[0] = 1.0;
    if (quad) { slot[1] = 4; sg[1] = 1.0; slot[2] = 3; sg[2] = -1.0; }
    else      { slot[1] = 3; sg[1] = 1.0; slot[2] = 4; sg[2] = 1.0; }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) K[slot[i] * 6 + slot[j]] = (sg[i] * sg[j]) * kb[3 * i + j];
    if (diag) K[35] = k_rz;
}

// Global 6x6 block = blockdiag(R, R)^T K blockdiag(R, R).
PN_HD void pn_shell_rotate_block(const double *R, const double *K, double *Kg) {
    double B[9], Gb[9];
    for (int bi = 0; bi < 2; ++bi)
        for (int bj = 0; bj < 2; ++bj) {
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) B[3 * i + j] = K[(3 * bi + i) * 6 + 3 * bj + j];
            pn_congruence3(R, B, Gb);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) Kg[(3 * bi + i) * 6 + 3 * bj + j] = Gb[3 * i + j];
        }
}

// Uniform-pressure fixed-end force on the local w freedom of node a.
//   quad  (Quad3D.fer :721-788): sum_gp N_a p det J with p = -pressure
//   plate (Plate3D.fer :324-393): closed form -4 p b c (1/4, +-c/12, -+b/12)
PN_HD double pn_quad_fer_w(const PnShellGeom &G, int a, double pressure) {
    double p = -pressure, f = 0.0;
    for (int g = 0; g < 4; ++g) {
        double r, s, J[4];
        pn_gauss_point(g, r, s);
        pn_jacobian(G.xl, G.yl, r, s, J);
        double det = J[0] * J[3] - J[1] * J[2];
        double sr = (a == 0 || a == 3) ? -1.0 : 1.0, ss = (a < 2) ? -1.0 : 1.0;
        double N = 0.25 * (1 + sr * r) * (1 + ss * s);
        f += N * p * det;
    }
    return f;
}
PN_HD void pn_plate_fer_node(int a, double width, double height, double pressure, double *f3) {
    double b = width / 2, c = height / 2;
    double q = -4 * pressure * c * b;
    double sx = (a < 2) ? 1.0 : -1.0;                 // theta_x term: +c/12 at i, j; -c/12 at m, n
    double sy = (a == 0 || a == 3) ? -1.0 : 1.0;      // theta_y term: -b/12 at i, n; +b/12 at j, m
    f3[0] = q * 0.25;
    f3[1] = q * (sx * c / 12);
    f3[2] = q * (sy * b / 12);
}
