"""Per-element fixed-end reactions on the host: the inspection methods `Member3D._fer_unc / fer / FER`
(`Member3D.py:742-850, 1080-1091`), `Quad3D.fer / FER` (`Quad3D.py:721-788, 870-882`) and `Plate3D.fer / FER`
(`Plate3D.py:324-393, 510-522`) for ONE element.  The analysis itself never calls these: there the load vectors of all
elements and all load cases come from the kernels `k_member_fer_groups` / `k_shell_fer` (`csrc/pn_elements.cu`), which
tests/test_gpu_parity.py holds to the reference's global `FER`.

A member's clamped-end reactions are written once, as the influence of a unit force or moment at station s; point loads
evaluate it, linearly varying loads integrate it with three Gauss points (exact: the integrand has degree four), instead
of the reference's expanded polynomials (`FixedEndReactions.py`)."""
from __future__ import annotations

import numpy as np

_G3_X = np.array([-(3/5)**0.5, 0.0, (3/5)**0.5])
_G3_W = np.array([5/9, 8/9, 5/9])


def clamped_force(v, s, L):
    """Fixed-end reaction vector (12,) of a force v = (Fx, Fy, Fz) at station s of a member clamped at both ends."""
    b = L - s
    fx, fy, fz = v
    near, far = b*b*(L + 2*s)/L**3, s*s*(L + 2*b)/L**3          # shares of a transverse force
    mi, mj = s*b*b/L**2, s*s*b/L**2                             # its end moments
    return np.array([-fx*b/L, -fy*near, -fz*near, 0.0, fz*mi, -fy*mi, -fx*s/L, -fy*far, -fz*far, 0.0, -fz*mj, fy*mj])


def clamped_moment(v, s, L):
    """Fixed-end reaction vector (12,) of a moment v = (Mx, My, Mz) at station s."""
    b = L - s
    mx, my, mz = v
    q = 6*s*b/L**3
    ci, cj = b*(2*s - b)/L**2, s*(2*b - s)/L**2
    return np.array([0.0, mz*q, -my*q, -mx*b/L, my*ci, mz*ci, 0.0, -mz*q, my*q, -mx*s/L, my*cj, mz*cj])


def member_fer_unc(pts, dists, L):
    """`Member3D._fer_unc` from the local-axis loads of `diagrams.local_loads`: point loads (x, [F(3), M(3)]) and
    distributed loads (x1, x2, w1(3), w2(3))."""
    fer = np.zeros(12)
    for x, v in pts:
        fer += clamped_force(v[:3], x, L) + clamped_moment(v[3:], x, L)
    for x1, x2, a, b in dists:
        half, mid = (x2 - x1)/2, (x2 + x1)/2
        for gx, gw in zip(_G3_X, _G3_W):
            s = mid + half*gx
            w = a + (b - a)*((s - x1)/(x2 - x1)) if x2 != x1 else a
            fer += gw*half*clamped_force(w, s, L)
    return fer


def member_ke_unc(E, G, A, Iy, Iz, J, L):
    """Uncondensed local stiffness of a frame member (`Member3D._ke_unc`, `:176-208`), built block by block: axial and
    torsional springs, and the two Hermite bending blocks (about z on (v, rz), about y on (w, ry) with its sign flip)."""
    k = np.zeros((12, 12))
    for i, j, s in ((0, 6, A*E/L), (3, 9, G*J/L)):
        k[np.ix_([i, j], [i, j])] += s*np.array([[1.0, -1.0], [-1.0, 1.0]])
    hermite = np.array([[12/L**3, 6/L**2, -12/L**3, 6/L**2], [6/L**2, 4/L, -6/L**2, 2/L], [-12/L**3, -6/L**2, 12/L**3, -6/L**2],
                        [6/L**2, 2/L, -6/L**2, 4/L]])
    k[np.ix_([1, 5, 7, 11], [1, 5, 7, 11])] += E*Iz*hermite
    flip = np.diag([1.0, -1.0, 1.0, -1.0])
    k[np.ix_([2, 4, 8, 10], [2, 4, 8, 10])] += E*Iy*(flip @ hermite @ flip)
    return k


def condense_vector(k, f, releases):
    """`Member3D.fer` (`:742-772`): `f1 - k12 k22^-1 f2` on the unreleased DOFs, zeros at the released ones."""
    rel = np.asarray(releases, dtype=bool)
    if not rel.any():
        return f.copy()
    r1, r2 = np.nonzero(~rel)[0], np.nonzero(rel)[0]
    out = np.zeros(12)
    out[r1] = f[r1] - k[np.ix_(r1, r2)] @ np.linalg.inv(k[np.ix_(r2, r2)]) @ f[r2]
    return out


def quad_fer(quad, combo):
    """`Quad3D.fer`: the pressure's consistent nodal forces on the four local w DOFs, 2 x 2 Gauss points (`:721-788`).
    The in-plane coordinates are those of `Quad3D._local_coords` (`:107-141`): x along i -> j, z = x cross (i -> m)."""
    p = [np.array([n.X, n.Y, n.Z], dtype=float) for n in (quad.i_node, quad.j_node, quad.m_node, quad.n_node)]
    x_axis = p[1] - p[0]
    y_axis = np.cross(np.cross(x_axis, p[2] - p[0]), x_axis)
    x_axis, y_axis = x_axis/np.linalg.norm(x_axis), y_axis/np.linalg.norm(y_axis)
    xy = np.array([[np.dot(q - p[0], x_axis), np.dot(q - p[0], y_axis)] for q in p])
    pressure = -sum(combo.factors[case]*value for value, case in quad.pressures if case in combo.factors)
    corner = np.array([[-1.0, -1.0], [1.0, -1.0], [1.0, 1.0], [-1.0, 1.0]])
    g = 1/3**0.5
    w = np.zeros(4)
    for xi, eta in corner*g:
        N = (1 + corner[:, 0]*xi)*(1 + corner[:, 1]*eta)/4
        dN = np.array([corner[:, 0]*(1 + corner[:, 1]*eta), corner[:, 1]*(1 + corner[:, 0]*xi)])/4
        w += N*pressure*np.linalg.det(dN @ xy)
    fer = np.zeros((24, 1))
    fer[[2, 8, 14, 20], 0] = w
    return fer


def plate_fer(plate, combo):
    """`Plate3D.fer`: closed-form consistent forces and moments of a uniform pressure on a rectangle (`:324-393`)."""
    pressure = sum(combo.factors[case]*value for value, case in plate.pressures if case in combo.factors)
    b, c = plate.width()/2, plate.height()/2
    fer = np.zeros((24, 1))
    for k, (sx, sy) in enumerate(((1, -1), (1, 1), (-1, 1), (-1, -1))):
        fer[6*k + 2:6*k + 5, 0] = -4*pressure*c*b*np.array([1/4, sx*c/12, sy*b/12])
    return fer
