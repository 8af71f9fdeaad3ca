// Internal force / stress recovery for the four-node shells, restated from
//   Quad3D.shear :1017-1094, Quad3D.moment :1096-1171, Quad3D.membrane :1173-1239 (DKMQ: values at the four Gauss
//   points, extrapolated to (xi, eta) with the bilinear functions of (xi, eta) / gp), and
//   Plate3D._C / _Q / _a :524-588, Plate3D.moment :590-606, shear :608-648, membrane :650-692.
// Everything here is per element and per load combination; pn_stress.cu batches it over elements x combos.
#pragma once
#include "elem_shell.cuh"

// Values at the four Gauss points from the local displacement vector d (24: dx dy dz rx ry rz per node).
//   V[g] = { Qx, Qy, Mx, My, Mxy, Sx, Sy, Txy }
// rec[4 * g + a] is the stage-B record of node a at Gauss point g (pn_shell_gp_node).  `bending` = false (plates)
// leaves the first five entries zero: the plate's bending results come from its displacement polynomial instead.
PN_HD void pn_shell_gp_values(const PnGpNode *rec, const PnShellMat &m, const double *d, bool bending, double V[4][8]) {
    const double d11 = m.Db, d12 = m.nu * m.Db, d33 = m.Db * ((1 - m.nu) / 2);
    for (int g = 0; g < 4; ++g) {
        double kx = 0, ky = 0, kxy = 0, g0 = 0, g1 = 0, ex = 0, ey = 0, exy = 0;
        for (int a = 0; a < 4; ++a) {
            const PnGpNode &r = rec[4 * g + a];
            if (bending) {
                // Quad3D.shear :1045-1048: d[[3, 9, 15, 21]] *= -1, then (dz, ry, rx) per node
                const double w = d[6 * a + 2], t1 = d[6 * a + 4], t2 = -d[6 * a + 3];
                kx = pn_fma(r.Bb[2], t2, pn_fma(r.Bb[1], t1, pn_fma(r.Bb[0], w, kx)));
                ky = pn_fma(r.Bb[5], t2, pn_fma(r.Bb[4], t1, pn_fma(r.Bb[3], w, ky)));
                kxy = pn_fma(r.Bb[8], t2, pn_fma(r.Bb[7], t1, pn_fma(r.Bb[6], w, kxy)));
                g0 = pn_fma(r.Bs[2], t2, pn_fma(r.Bs[1], t1, pn_fma(r.Bs[0], w, g0)));
                g1 = pn_fma(r.Bs[5], t2, pn_fma(r.Bs[4], t1, pn_fma(r.Bs[3], w, g1)));
            }
            const double u = d[6 * a], v = d[6 * a + 1];
            ex = pn_fma(r.dN[0], u, ex);
            ey = pn_fma(r.dN[1], v, ey);
            exy = pn_fma(r.dN[0], v, pn_fma(r.dN[1], u, exy));
        }
        V[g][0] = m.Hs * g0;
        V[g][1] = m.Hs * g1;
        V[g][2] = d11 * kx + d12 * ky;
        V[g][3] = d12 * kx + d11 * ky;
        V[g][4] = d33 * kxy;
        V[g][5] = m.c11 * ex + m.c12 * ey;
        V[g][6] = m.c21 * ex + m.c22 * ey;
        V[g][7] = m.c33 * exy;
    }
}

// Extrapolation from the Gauss points to natural coordinates (r, s): H evaluated at (r / gp, s / gp)
// (Quad3D.py:1053-1059, Plate3D.py:666-672).  Components [k0, k1) of V.
PN_HD void pn_shell_extrapolate(const double V[4][8], double r, double s, int k0, int k1, double *out) {
    const double rx = r / PN_GP, sx = s / PN_GP;
    const double H[4] = {0.25 * ((1 - rx) * (1 - sx)), 0.25 * ((1 + rx) * (1 - sx)), 0.25 * ((1 + rx) * (1 + sx)),
                         0.25 * ((1 - rx) * (1 + sx))};
    for (int k = k0; k < k1; ++k) out[k] = H[0] * V[0][k] + H[1] * V[1][k] + H[2] * V[2][k] + H[3] * V[3][k];
}

// local (Qx Qy | Mx My Mxy | Sx Sy Txy) -> the reference's "global" results (Quad3D.py:1069-1092, 1146-1171, 1214-1239):
//   shear   : dir_cos^T diag(Qx, Qy, 0) -> entries [0,0], [1,1], [2,1]
//   moment  : dir_cos [[Mx, Mxy, 0], [Mxy, -My, 0], 0] dir_cos^T -> entries [0,0], [1,1], [0,1]
//   membrane: dir_cos [[Sx, Txy, 0], [Txy, Sy, 0], 0] dir_cos^T  -> entries [0,0], [1,1], [0,1]
PN_HD void pn_quad_stress_global(const double *R, const double *loc, double *g9) {
    g9[0] = R[0] * loc[0];
    g9[1] = R[4] * loc[1];
    g9[2] = R[5] * loc[1];
    for (int q = 0; q < 2; ++q) {
        const double a = loc[q == 0 ? 2 : 5], b = q == 0 ? -loc[3] : loc[6], cxy = loc[q == 0 ? 4 : 7];
        // T = R M: rows i, columns l < 2 (third column of M is zero); G = T R^T
        double T[3][2];
        for (int i = 0; i < 3; ++i) {
            T[i][0] = R[3 * i] * a + R[3 * i + 1] * cxy;
            T[i][1] = R[3 * i] * cxy + R[3 * i + 1] * b;
        }
        g9[3 + 3 * q] = T[0][0] * R[0] + T[0][1] * R[1];
        g9[4 + 3 * q] = T[1][0] * R[3] + T[1][1] * R[4];
        g9[5 + 3 * q] = T[0][0] * R[3] + T[0][1] * R[4];
    }
}

// ---- rectangular plate: closed form of inv(C) (tools/derive_plate_cinv.py) ------------------------------------------
#if defined(__CUDACC__)
#define PN_PLATE_CINV_TABLE(name) static __constant__ signed char pn_cinv_##name[144]
#else
#define PN_PLATE_CINV_TABLE(name) static const signed char pn_cinv_##name[144]
#endif
#include "plate_cinv.inc"

// one entry of inv(C) for a W x H plate; pw[k + 3] = W^k, ph[k + 3] = H^k for k = -3 .. 3
PN_D double pn_plate_cinv_entry(int idx, const double *pw, const double *ph) {
    const int num = pn_cinv_NUM[idx];
    if (num == 0) return 0.0;
    return ((double)num / PN_PLATE_CINV_DEN) * (pw[pn_cinv_PW[idx] + 3] * ph[pn_cinv_PH[idx] + 3]);
}

PN_HD void pn_plate_powers(double W, double *pw) {
    pw[3] = 1.0; pw[4] = W; pw[5] = W * W; pw[6] = W * W * W;
    pw[2] = 1.0 / W; pw[1] = 1.0 / (W * W); pw[0] = 1.0 / (W * W * W);
}

// Plate3D.moment / shear at local (x, y) from the polynomial constants a (12): out = { Qx, Qy, Mx, My, Mxy }.
// Db uses the kx_mod / ky_mod factors (Plate3D.Db :115-134), unlike the plate's bending stiffness itself.
PN_HD void pn_plate_bending_results(const double *props, const double *a, double x, double y, double *out) {
    const double t = props[0], E = props[1], nu = props[2], kx = props[3], ky = props[4];
    const double f = (t * t * t) / (12 * (1 - nu * nu));
    const double Ex = E * kx, Ey = E * ky, G = E / (2 * (1 + nu));
    const double b00 = f * Ex, b01 = f * (nu * Ex), b10 = f * (nu * Ey), b11 = f * Ey, b22 = f * G;
    // curvatures Q(x, y) a  (Plate3D._Q :559-571)
    const double q0 = -2 * a[3] - 6 * x * a[6] - 2 * y * a[7] - 6 * x * y * a[10];
    const double q1 = -2 * a[5] - 2 * x * a[8] - 6 * y * a[9] - 6 * x * y * a[11];
    const double q2 = -2 * a[4] - 4 * x * a[7] - 4 * y * a[8] - 6 * (x * x) * a[10] - 6 * (y * y) * a[11];
    out[2] = -(b00 * q0 + b01 * q1);
    out[3] = -(b10 * q0 + b11 * q1);
    out[4] = -(b22 * q2);
    // derivative matrices of Plate3D.shear :627-641
    const double r0 = -6 * a[6] - 6 * y * a[10];                 // d/dx of the curvature rows
    const double r1 = -2 * a[8] - 6 * y * a[11];
    const double r2 = -4 * a[7] - 12 * x * a[10];
    const double s0 = -2 * a[7] - 6 * x * a[10];                 // d/dy
    const double s1 = -6 * a[9] - 6 * x * a[11];
    const double s2 = -4 * a[8] - 12 * y * a[11];
    const double dMx_dx = b00 * r0 + b01 * r1, dMxy_dx = b22 * r2;
    const double dMy_dy = b10 * s0 + b11 * s1, dMxy_dy = b22 * s2;
    out[0] = dMx_dx + dMxy_dy;
    out[1] = dMy_dy + dMxy_dx;
}
