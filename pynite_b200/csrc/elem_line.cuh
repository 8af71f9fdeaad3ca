// Two-node elements: frame member (12x12 with end releases), axial spring, geometric stiffness,
// fixed-end reactions.  Plain host/device functions; the kernels that call them are in
// pn_elements.cu.  Reference behaviour restated (not translated) from:
//   Member3D._ke_unc :176-208, ke :147-173, kg :210-266, T :933-1036, fer :742-772, _fer_unc :774-850
//   Spring3D.ke :59-83, T :114-188;  FixedEndReactions.py :20-212
#pragma once
#include "pn_common.cuh"

// Direction cosines R (rows = local x, y, z) of the line i -> j with roll angle `rot_deg`.
// Returns the length.  Vertical / horizontal / inclined branches follow Member3D.T.
PN_HD double pn_line_dircos(const double *pi, const double *pj, double rot_deg, double *R) {
    double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
    // Node3D.distance squares (self - other); the sign does not matter for the squares.
    double L = sqrt((pi[0] - pj[0]) * (pi[0] - pj[0]) + (pi[1] - pj[1]) * (pi[1] - pj[1]) +
                    (pi[2] - pj[2]) * (pi[2] - pj[2]));
    double x[3] = {dx / L, dy / L, dz / L};
    double y[3], z[3];
    if (pn_isclose(pi[0], pj[0]) && pn_isclose(pi[2], pj[2])) {
        y[0] = (pj[1] > pi[1]) ? -1.0 : 1.0; y[1] = 0.0; y[2] = 0.0;
        z[0] = 0.0; z[1] = 0.0; z[2] = 1.0;
    } else if (pn_isclose(pi[1], pj[1])) {
        y[0] = 0.0; y[1] = 1.0; y[2] = 0.0;
        pn_cross(x, y, z);
        double n = pn_norm3(z);
        z[0] /= n; z[1] /= n; z[2] /= n;
    } else {
        double proj[3] = {dx, 0.0, dz};
        if (pj[1] > pi[1]) pn_cross(proj, x, z); else pn_cross(x, proj, z);
        double n = pn_norm3(z);
        z[0] /= n; z[1] /= n; z[2] /= n;
        pn_cross(z, x, y);
        n = pn_norm3(y);
        y[0] /= n; y[1] /= n; y[2] /= n;
    }
    if (rot_deg != 0.0) {
        // Rodrigues rotation of y and z about x (Member3D.py:1004-1024).
        double theta = rot_deg * (3.14159265358979323846 / 180.0);
        double c = cos(theta), s = sin(theta);
        double v[3], w[3], cr[3];
        for (int pass = 0; pass < 2; ++pass) {
            double *t = pass == 0 ? y : z;
            v[0] = t[0]; v[1] = t[1]; v[2] = t[2];
            pn_cross(x, v, cr);
            double d = x[0] * v[0] + x[1] * v[1] + x[2] * v[2];
            for (int k = 0; k < 3; ++k) w[k] = v[k] * c + cr[k] * s + x[k] * d * (1.0 - c);
            double n = pn_norm3(w);
            t[0] = w[0] / n; t[1] = w[1] / n; t[2] = w[2] / n;
        }
    }
    for (int k = 0; k < 3; ++k) { R[k] = x[k]; R[3 + k] = y[k]; R[6 + k] = z[k]; }
    return L;
}

PN_HD void pn_zero144(double *k) {
    for (int i = 0; i < 144; ++i) k[i] = 0.0;
}

// Uncondensed local frame stiffness, k[12*12] row-major.
PN_HD void pn_member_ke_unc(double E, double G, double A, double Iy, double Iz, double J, double L, double *k) {
    pn_zero144(k);
    double L2 = L * L, L3 = L2 * L;
    double ax = A * E / L, tq = G * J / L;
    double z12 = 12 * E * Iz / L3, z6 = 6 * E * Iz / L2, z4 = 4 * E * Iz / L, z2 = 2 * E * Iz / L;
    double y12 = 12 * E * Iy / L3, y6 = 6 * E * Iy / L2, y4 = 4 * E * Iy / L, y2 = 2 * E * Iy / L;
#define K_(r, c) k[(r) * 12 + (c)]
    K_(0, 0) = ax; K_(6, 6) = ax; K_(0, 6) = -ax; K_(6, 0) = -ax;
    K_(3, 3) = tq; K_(9, 9) = tq; K_(3, 9) = -tq; K_(9, 3) = -tq;
    // bending about local z: dofs 1 (v_i), 5 (rz_i), 7 (v_j), 11 (rz_j)
    K_(1, 1) = z12;  K_(1, 5) = z6;   K_(1, 7) = -z12;  K_(1, 11) = z6;
    K_(5, 1) = z6;   K_(5, 5) = z4;   K_(5, 7) = -z6;   K_(5, 11) = z2;
    K_(7, 1) = -z12; K_(7, 5) = -z6;  K_(7, 7) = z12;   K_(7, 11) = -z6;
    K_(11, 1) = z6;  K_(11, 5) = z2;  K_(11, 7) = -z6;  K_(11, 11) = z4;
    // bending about local y: dofs 2 (w_i), 4 (ry_i), 8 (w_j), 10 (ry_j)
    K_(2, 2) = y12;  K_(2, 4) = -y6;  K_(2, 8) = -y12;  K_(2, 10) = -y6;
    K_(4, 2) = -y6;  K_(4, 4) = y4;   K_(4, 8) = y6;    K_(4, 10) = y2;
    K_(8, 2) = -y12; K_(8, 4) = y6;   K_(8, 8) = y12;   K_(8, 10) = y6;
    K_(10, 2) = -y6; K_(10, 4) = y2;  K_(10, 8) = y6;   K_(10, 10) = y4;
#undef K_
}

// Uncondensed local geometric stiffness for axial force P.
PN_HD void pn_member_kg_unc(double P, double A, double Iy, double Iz, double L, double *k) {
    pn_zero144(k);
    double s = P / L;
    double a = 1.0 * s, t = (Iy + Iz) / A * s;
    double c65 = 6.0 / 5.0 * s, l10 = L / 10.0 * s, l15 = 2.0 * L * L / 15.0 * s, l30 = L * L / 30.0 * s;
#define K_(r, c) k[(r) * 12 + (c)]
    K_(0, 0) = a; K_(6, 6) = a; K_(0, 6) = -a; K_(6, 0) = -a;
    K_(3, 3) = t; K_(9, 9) = t; K_(3, 9) = -t; K_(9, 3) = -t;
    K_(1, 1) = c65;  K_(1, 5) = l10;   K_(1, 7) = -c65;  K_(1, 11) = l10;
    K_(5, 1) = l10;  K_(5, 5) = l15;   K_(5, 7) = -l10;  K_(5, 11) = -l30;
    K_(7, 1) = -c65; K_(7, 5) = -l10;  K_(7, 7) = c65;   K_(7, 11) = -l10;
    K_(11, 1) = l10; K_(11, 5) = -l30; K_(11, 7) = -l10; K_(11, 11) = l15;
    K_(2, 2) = c65;  K_(2, 4) = -l10;  K_(2, 8) = -c65;  K_(2, 10) = -l10;
    K_(4, 2) = -l10; K_(4, 4) = l15;   K_(4, 8) = l10;   K_(4, 10) = -l30;
    K_(8, 2) = -c65; K_(8, 4) = l10;   K_(8, 8) = c65;   K_(8, 10) = l10;
    K_(10, 2) = -l10; K_(10, 4) = -l30; K_(10, 8) = l10;  K_(10, 10) = l15;
#undef K_
}

// Static condensation of the released DOFs (bit r of `rel`): k <- k11 - k12 k22^-1 k21 on the kept
// DOFs, zero rows/columns on the released ones (Member3D.py:157-173).  Done as successive
// elimination of the released DOFs, which is the same Schur complement.  An optional vector f
// (12 entries) is condensed alongside: f1 - k12 k22^-1 f2 (Member3D.py:756-772).
PN_HD void pn_condense(double *k, double *f, unsigned rel) {
    if (rel == 0) return;
    for (int r = 0; r < 12; ++r) {
        if (!((rel >> r) & 1u)) continue;
        double piv = k[r * 12 + r];
        double fr = f ? f[r] : 0.0;
        for (int i = 0; i < 12; ++i) {
            if (i == r) continue;
            double kir = k[i * 12 + r];
            if (kir == 0.0) continue;
            double m = kir / piv;
            for (int j = 0; j < 12; ++j) k[i * 12 + j] -= m * k[r * 12 + j];
            if (f) f[i] -= m * fr;
        }
        for (int j = 0; j < 12; ++j) { k[r * 12 + j] = 0.0; k[j * 12 + r] = 0.0; }
        if (f) f[r] = 0.0;
    }
}

// Condensed local elastic stiffness of a member.
PN_HD void pn_member_ke(const double *props, double L, unsigned rel, double *k) {
    pn_member_ke_unc(props[0], props[1], props[2], props[3], props[4], props[5], L, k);
    pn_condense(k, nullptr, rel);
}

// Condensed local geometric stiffness; identically zero when P == 0 (math.isclose(P, 0.0) with
// abs_tol = 0 is an exact-zero test, Member3D.py:250).
PN_HD void pn_member_kg(const double *props, double L, unsigned rel, double P, double *k) {
    if (P == 0.0) { pn_zero144(k); return; }
    pn_member_kg_unc(P, props[2], props[3], props[4], L, k);
    pn_condense(k, nullptr, rel);
}

// ---- fixed-end reactions (local, uncondensed), accumulated into f[12] --------------------------
// transverse point load, axis 1 = local y, 2 = local z (FixedEndReactions.py:20-55)
PN_HD void pn_fer_pt(double P, double x, double L, int axis, double *f) {
    double b = L - x;
    double v1 = -P * b * b * (L + 2 * x) / (L * L * L);
    double v2 = -P * x * x * (L + 2 * b) / (L * L * L);
    double m1 = P * x * b * b / (L * L);
    double m2 = P * x * x * b / (L * L);
    if (axis == 1) { f[1] += v1; f[5] += -m1; f[7] += v2; f[11] += m2; }
    else           { f[2] += v1; f[4] += m1;  f[8] += v2; f[10] += -m2; }
}
// concentrated moment about local y (axis 1) or z (axis 2) (FixedEndReactions.py:58-93)
PN_HD void pn_fer_moment(double M, double x, double L, int axis, double *f) {
    double b = L - x;
    double v = 6 * M * x * b / (L * L * L);
    double m1 = M * b * (2 * x - b) / (L * L);
    double m2 = M * x * (2 * b - x) / (L * L);
    if (axis == 2) { f[1] += v;  f[5] += m1; f[7] += -v; f[11] += m2; }
    else           { f[2] += -v; f[4] += m1; f[8] += v;  f[10] += m2; }
}
// linearly varying transverse load on [x1, x2] (FixedEndReactions.py:97-134): minus the integrals
// of the clamped-beam Hermite shape functions against w(x), by exact 4-point Gauss quadrature.
PN_HD void pn_fer_lin(double w1, double w2, double x1, double x2, double L, int axis, double *f) {
    const double gx[4] = {-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526};
    const double gw[4] = {0.3478548451374538, 0.6521451548625461, 0.6521451548625461, 0.3478548451374538};
    double h = 0.5 * (x2 - x1), xm = 0.5 * (x1 + x2);
    double i1 = 0, i2 = 0, i3 = 0, i4 = 0;
    for (int g = 0; g < 4; ++g) {
        double xs = xm + h * gx[g];
        double ws = (x2 != x1) ? w1 + (w2 - w1) * (xs - x1) / (x2 - x1) : 0.0;
        double s = xs / L, s2 = s * s, s3 = s2 * s;
        double c = gw[g] * ws;
        i1 += c * (1 - 3 * s2 + 2 * s3);
        i2 += c * (L * (s - 2 * s2 + s3));
        i3 += c * (3 * s2 - 2 * s3);
        i4 += c * (L * (-s2 + s3));
    }
    i1 *= h; i2 *= h; i3 *= h; i4 *= h;
    if (axis == 1) { f[1] += -i1; f[5] += -i2; f[7] += -i3; f[11] += -i4; }
    else           { f[2] += -i1; f[4] += i2;  f[8] += -i3; f[10] += i4; }
}
PN_HD void pn_fer_axial_pt(double P, double x, double L, double *f) {      // :138-159
    f[0] += -P * (L - x) / L;
    f[6] += -P * x / L;
}
PN_HD void pn_fer_axial_lin(double p1, double p2, double x1, double x2, double L, double *f) {   // :163-188
    f[0] += 1.0 / (6 * L) * (x1 - x2) * (3 * L * p1 + 3 * L * p2 - 2 * p1 * x1 - p1 * x2 - p2 * x1 - 2 * p2 * x2);
    f[6] += 1.0 / (6 * L) * (x1 - x2) * (2 * p1 * x1 + p1 * x2 + p2 * x1 + 2 * p2 * x2);
}
PN_HD void pn_fer_torque(double T, double x, double L, double *f) {       // :191-212
    f[3] += -T * (L - x) / L;
    f[9] += -T * x / L;
}

// One member load record (kind / params as documented for pn_set_loads) with unit factor, added to
// the uncondensed local vector f[12] (Member3D._fer_unc :774-850).
PN_HD void pn_member_load_fer(int kind, const double *p, const double *R, double L, double *f) {
    if (kind < 12) {
        double P = p[0], x = p[1];
        switch (kind) {
        case 0: pn_fer_axial_pt(P, x, L, f); break;
        case 1: pn_fer_pt(P, x, L, 1, f); break;
        case 2: pn_fer_pt(P, x, L, 2, f); break;
        case 3: pn_fer_torque(P, x, L, f); break;
        case 4: pn_fer_moment(P, x, L, 1, f); break;
        case 5: pn_fer_moment(P, x, L, 2, f); break;
        default: {
            int k = (kind - 6) % 3;                 // global axis of the load
            double l0 = R[k] * P, l1 = R[3 + k] * P, l2 = R[6 + k] * P;
            if (kind < 9) {
                pn_fer_axial_pt(l0, x, L, f); pn_fer_pt(l1, x, L, 1, f); pn_fer_pt(l2, x, L, 2, f);
            } else {
                pn_fer_torque(l0, x, L, f); pn_fer_moment(l1, x, L, 1, f); pn_fer_moment(l2, x, L, 2, f);
            }
        }
        }
    } else {
        double w1 = p[0], w2 = p[1], x1 = p[2], x2 = p[3];
        if (kind == 12) pn_fer_axial_lin(w1, w2, x1, x2, L, f);
        else if (kind == 13) pn_fer_lin(w1, w2, x1, x2, L, 1, f);
        else if (kind == 14) pn_fer_lin(w1, w2, x1, x2, L, 2, f);
        else {
            int k = kind - 15;
            pn_fer_axial_lin(R[k] * w1, R[k] * w2, x1, x2, L, f);
            pn_fer_lin(R[3 + k] * w1, R[3 + k] * w2, x1, x2, L, 1, f);
            pn_fer_lin(R[6 + k] * w1, R[6 + k] * w2, x1, x2, L, 2, f);
        }
    }
}
