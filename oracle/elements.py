"""Per-element restatement of the reference's element routines (numpy, one element at a time).

Every function cites the reference lines it follows.  Operation order is kept close to the
reference (same matrix products, `inv`/`det` from LAPACK) so that results agree to a few ulp.
"""
from math import isclose

import numpy as np
from numpy.linalg import inv, det, norm

from oracle import plate_tables as _pt

GP = 1/3**0.5
_GAUSS = ((-GP, -GP), (GP, -GP), (GP, GP), (-GP, GP))      # Quad3D.py:532-544 ordering


# ------------------------------------------------------------------------------------------------
# frame members and springs
# ------------------------------------------------------------------------------------------------
def line_dircos(pi, pj, rotation=0.0):
    """3x3 direction cosines of a 2-node element (rows = local x, y, z).

    Member3D.T (`Member3D.py:933-1027`) / Spring3D.T (`Spring3D.py:114-179`): vertical members take
    y = (-+1, 0, 0), z = (0, 0, 1); horizontal members y = (0, 1, 0), z = x cross y; otherwise z is
    normal to the vertical plane containing the member.  A non-zero `rotation` (degrees) rolls y
    and z about x with Rodrigues' formula.
    """
    Xi, Yi, Zi = pi
    Xj, Yj, Zj = pj
    L = ((Xi - Xj)**2 + (Yi - Yj)**2 + (Zi - Zj)**2)**0.5
    x = [(Xj - Xi)/L, (Yj - Yi)/L, (Zj - Zi)/L]
    if isclose(Xi, Xj) and isclose(Zi, Zj):
        y = [-1, 0, 0] if Yj > Yi else [1, 0, 0]
        z = [0, 0, 1]
    elif isclose(Yi, Yj):
        y = [0, 1, 0]
        z = np.cross(x, y)
        z = np.divide(z, (z[0]**2 + z[1]**2 + z[2]**2)**0.5)
    else:
        proj = [Xj - Xi, 0, Zj - Zi]
        z = np.cross(proj, x) if Yj > Yi else np.cross(x, proj)
        z = np.divide(z, (z[0]**2 + z[1]**2 + z[2]**2)**0.5)
        y = np.cross(z, x)
        y = np.divide(y, (y[0]**2 + y[1]**2 + y[2]**2)**0.5)
    if rotation != 0.0:
        theta = np.radians(rotation)
        c, s = np.cos(theta), np.sin(theta)
        u = np.array(x)
        v = np.array(y)
        y = v*c + np.cross(u, v)*s + u*np.dot(u, v)*(1 - c)
        y /= norm(y)
        v = np.array(z)
        z = v*c + np.cross(u, v)*s + u*np.dot(u, v)*(1 - c)
        z /= norm(z)
    return np.array([x, y, z], dtype=float)


def block_diag_T(R, nblocks):
    """Transformation matrix with `nblocks` copies of R on the diagonal (`Member3D.py:1030-1034`)."""
    T = np.zeros((3*nblocks, 3*nblocks))
    for b in range(nblocks):
        T[3*b:3*b+3, 3*b:3*b+3] = R
    return T


def member_ke_unc(E, G, A, Iy, Iz, J, L):
    """Uncondensed local 12x12 frame stiffness (`Member3D.py:176-208`)."""
    k = np.zeros((12, 12))
    ax = A*E/L
    tq = G*J/L
    for a, b, v in ((0, 6, ax), (3, 9, tq)):
        k[a, a] = v
        k[b, b] = v
        k[a, b] = -v
        k[b, a] = -v
    # bending about local z (dofs 1, 5, 7, 11) and about local y (dofs 2, 4, 8, 10)
    z12, z6, z4, z2 = 12*E*Iz/L**3, 6*E*Iz/L**2, 4*E*Iz/L, 2*E*Iz/L
    y12, y6, y4, y2 = 12*E*Iy/L**3, 6*E*Iy/L**2, 4*E*Iy/L, 2*E*Iy/L
    k[np.ix_([1, 5, 7, 11], [1, 5, 7, 11])] = [[z12, z6, -z12, z6],
                                               [z6, z4, -z6, z2],
                                               [-z12, -z6, z12, -z6],
                                               [z6, z2, -z6, z4]]
    k[np.ix_([2, 4, 8, 10], [2, 4, 8, 10])] = [[y12, -y6, -y12, -y6],
                                               [-y6, y4, y6, y2],
                                               [-y12, y6, y12, y6],
                                               [-y6, y2, y6, y4]]
    return k


def member_kg_unc(P, A, Iy, Iz, L):
    """Uncondensed local geometric stiffness (`Member3D.py:224-242`)."""
    Ip = Iy + Iz
    k = np.zeros((12, 12))
    for a, b, v in ((0, 6, 1.0), (3, 9, Ip/A)):
        k[a, a] = v
        k[b, b] = v
        k[a, b] = -v
        k[b, a] = -v
    k[np.ix_([1, 5, 7, 11], [1, 5, 7, 11])] = [[6/5, L/10, -6/5, L/10],
                                               [L/10, 2*L**2/15, -L/10, -L**2/30],
                                               [-6/5, -L/10, 6/5, -L/10],
                                               [L/10, -L**2/30, -L/10, 2*L**2/15]]
    k[np.ix_([2, 4, 8, 10], [2, 4, 8, 10])] = [[6/5, -L/10, -6/5, -L/10],
                                               [-L/10, 2*L**2/15, L/10, -L**2/30],
                                               [-6/5, L/10, 6/5, L/10],
                                               [-L/10, -L**2/30, L/10, 2*L**2/15]]
    return k*P/L


def _split(releases):
    r1 = [i for i in range(12) if not releases[i]]
    r2 = [i for i in range(12) if releases[i]]
    return r1, r2


def condense_matrix(k, releases):
    """k11 - k12 inv(k22) k21 over released dofs, re-expanded with zero rows/cols (`Member3D.py:157-173`)."""
    r1, r2 = _split(releases)
    k11 = k[r1, :][:, r1]
    k12 = k[r1, :][:, r2]
    k21 = k[r2, :][:, r1]
    k22 = k[r2, :][:, r2]
    kc = np.subtract(k11, np.matmul(np.matmul(k12, inv(k22)), k21))
    out = np.zeros((12, 12))
    out[np.ix_(r1, r1)] = kc
    return out


def member_ke(E, G, A, Iy, Iz, J, L, releases):
    """Condensed local stiffness (`Member3D.py:147-173`)."""
    return condense_matrix(member_ke_unc(E, G, A, Iy, Iz, J, L), releases)


def member_kg(P, A, Iy, Iz, L, releases):
    """Condensed local geometric stiffness; all-zero when P == 0 (`Member3D.py:210-266`)."""
    kg = member_kg_unc(P, A, Iy, Iz, L)
    if isclose(P, 0.0):
        r1, _ = _split(releases)
        return np.zeros((12, 12))
    return condense_matrix(kg, releases)


def condense_vector(k_unc, f_unc, releases):
    """fer1 - k12 inv(k22) fer2, re-expanded with zeros (`Member3D.py:753-772`)."""
    r1, r2 = _split(releases)
    k12 = k_unc[r1, :][:, r2]
    k22 = k_unc[r2, :][:, r2]
    f1 = f_unc[r1, :]
    f2 = f_unc[r2, :]
    fc = np.subtract(f1, np.matmul(np.matmul(k12, inv(k22)), f2))
    out = np.zeros((12, 1))
    out[r1, :] = fc
    return out


def to_global(T, k):
    """inv(T) k T as the reference writes it (`Member3D.py:1046`, `Quad3D.py:867`, `Plate3D.py:508`)."""
    return np.matmul(np.matmul(inv(T), k), T)


# -- fixed-end reactions (FixedEndReactions.py) --------------------------------------------------
def fer_pt(P, x, L, axis):
    """Transverse point load (`FixedEndReactions.py:20-55`); axis 'y' or 'z'."""
    b = L - x
    f = np.zeros((12, 1))
    v1 = -P*b**2*(L + 2*x)/L**3
    v2 = -P*x**2*(L + 2*b)/L**3
    m1 = P*x*b**2/L**2
    m2 = P*x**2*b/L**2
    if axis == 'y':
        f[1, 0], f[5, 0], f[7, 0], f[11, 0] = v1, -m1, v2, m2
    else:
        f[2, 0], f[4, 0], f[8, 0], f[10, 0] = v1, m1, v2, -m2
    return f


def fer_moment(M, x, L, axis):
    """Concentrated moment (`FixedEndReactions.py:58-93`); axis 'y' or 'z'."""
    b = L - x
    f = np.zeros((12, 1))
    v = 6*M*x*b/L**3
    m1 = M*b*(2*x - b)/L**2
    m2 = M*x*(2*b - x)/L**2
    if axis == 'z':
        f[1, 0], f[5, 0], f[7, 0], f[11, 0] = v, m1, -v, m2
    else:
        f[2, 0], f[4, 0], f[8, 0], f[10, 0] = -v, m1, v, m2
    return f


def _hermite_load_integrals(w1, w2, x1, x2, L):
    """Integrals of the four cubic Hermite shape functions against a linearly varying load.

    The reference lists expanded polynomials (`FixedEndReactions.py:124-132`); those are
    -Int N_k(x) w(x) dx over [x1, x2] for the clamped-clamped beam shape functions.  Evaluated here
    with exact 4-point Gauss-Legendre quadrature (integrand degree 4).
    """
    gx = np.array([-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526])
    gw = np.array([0.3478548451374538, 0.6521451548625461, 0.6521451548625461, 0.3478548451374538])
    h = 0.5*(x2 - x1)
    xs = 0.5*(x1 + x2) + h*gx
    ws = w1 + (w2 - w1)*(xs - x1)/(x2 - x1) if x2 != x1 else np.zeros(4)
    s = xs/L
    N1 = 1 - 3*s**2 + 2*s**3
    N2 = L*(s - 2*s**2 + s**3)
    N3 = 3*s**2 - 2*s**3
    N4 = L*(-s**2 + s**3)
    return [h*np.sum(gw*ws*N) for N in (N1, N2, N3, N4)]


def fer_lin(w1, w2, x1, x2, L, axis):
    """Linearly varying transverse load between x1 and x2 (`FixedEndReactions.py:97-134`)."""
    i1, i2, i3, i4 = _hermite_load_integrals(w1, w2, x1, x2, L)
    f = np.zeros((12, 1))
    if axis == 'y':
        f[1, 0], f[5, 0], f[7, 0], f[11, 0] = -i1, -i2, -i3, -i4
    else:
        f[2, 0], f[4, 0], f[8, 0], f[10, 0] = -i1, i2, -i3, i4
    return f


def fer_axial_pt(P, x, L):
    """Axial point load (`FixedEndReactions.py:138-159`)."""
    f = np.zeros((12, 1))
    f[0, 0] = -P*(L - x)/L
    f[6, 0] = -P*x/L
    return f


def fer_axial_lin(p1, p2, x1, x2, L):
    """Linearly varying axial load (`FixedEndReactions.py:163-188`)."""
    f = np.zeros((12, 1))
    f[0, 0] = 1/(6*L)*(x1 - x2)*(3*L*p1 + 3*L*p2 - 2*p1*x1 - p1*x2 - p2*x1 - 2*p2*x2)
    f[6, 0] = 1/(6*L)*(x1 - x2)*(2*p1*x1 + p1*x2 + p2*x1 + 2*p2*x2)
    return f


def fer_torque(T, x, L):
    """Concentrated torque (`FixedEndReactions.py:191-212`)."""
    f = np.zeros((12, 1))
    f[3, 0] = -T*(L - x)/L
    f[9, 0] = -T*x/L
    return f


def member_fer_unc(pt_loads, dist_loads, factors, R, L):
    """Uncondensed local fixed-end vector of one member for one combo (`Member3D.py:774-850`).

    pt_loads: (direction, P, x, case); dist_loads: (direction, w1, w2, x1, x2, case, ...);
    factors: {case: factor}; R: 3x3 direction cosines.
    """
    f = np.zeros((12, 1))
    for case, factor in factors.items():
        for d, P, x, c in [pl[:4] for pl in pt_loads]:
            if c != case:
                continue
            if d == 'Fx':
                f = f + fer_axial_pt(factor*P, x, L)
            elif d == 'Fy':
                f = f + fer_pt(factor*P, x, L, 'y')
            elif d == 'Fz':
                f = f + fer_pt(factor*P, x, L, 'z')
            elif d == 'Mx':
                f = f + fer_torque(factor*P, x, L)
            elif d == 'My':
                f = f + fer_moment(factor*P, x, L, 'y')
            elif d == 'Mz':
                f = f + fer_moment(factor*P, x, L, 'z')
            elif d in ('FX', 'FY', 'FZ'):
                g = np.zeros(3)
                g['XYZ'.index(d[1])] = P
                loc = R @ g
                f = f + fer_axial_pt(factor*loc[0], x, L)
                f = f + fer_pt(factor*loc[1], x, L, 'y')
                f = f + fer_pt(factor*loc[2], x, L, 'z')
            elif d in ('MX', 'MY', 'MZ'):
                g = np.zeros(3)
                g['XYZ'.index(d[1])] = P
                loc = R @ g
                f = f + fer_torque(factor*loc[0], x, L)
                f = f + fer_moment(factor*loc[1], x, L, 'y')
                f = f + fer_moment(factor*loc[2], x, L, 'z')
            else:
                raise Exception('Invalid member point load direction specified.')
        for dl in dist_loads:
            d, w1, w2, x1, x2, c = dl[:6]
            if c != case:
                continue
            if d == 'Fx':
                f = f + fer_axial_lin(factor*w1, factor*w2, x1, x2, L)
            elif d in ('Fy', 'Fz'):
                f = f + fer_lin(factor*w1, factor*w2, x1, x2, L, d[1])
            elif d in ('FX', 'FY', 'FZ'):
                g1 = np.zeros(3)
                g2 = np.zeros(3)
                k = 'XYZ'.index(d[1])
                g1[k], g2[k] = w1, w2
                l1, l2 = R @ g1, R @ g2
                f = f + fer_axial_lin(factor*l1[0], factor*l2[0], x1, x2, L)
                f = f + fer_lin(factor*l1[1], factor*l2[1], x1, x2, L, 'y')
                f = f + fer_lin(factor*l1[2], factor*l2[2], x1, x2, L, 'z')
    return f


# ------------------------------------------------------------------------------------------------
# shells: common pieces
# ------------------------------------------------------------------------------------------------
def shell_dircos(pi, pj, pn):
    """Rows x, y, z of a 4-node shell: x along i->j, z = x cross (i->n), y = z cross x.

    `Quad3D.py:884-936`, `Plate3D.py:451-487`.
    """
    x = np.array([pj[0] - pi[0], pj[1] - pi[1], pj[2] - pi[2]], dtype=float)
    x = x/norm(x)
    xy = [pn[0] - pi[0], pn[1] - pi[1], pn[2] - pi[2]]
    z = np.cross(x, xy)
    z = z/norm(z)
    y = np.cross(z, x)
    y = y/norm(y)
    return np.array([x, y, z])


def membrane_C(E, nu, kx_mod, ky_mod):
    """Orthotropic plane-stress matrix (`Quad3D.py:498-521`, `Plate3D.py:89-113`)."""
    Ex, Ey = E*kx_mod, E*ky_mod
    G = E/(2*(1 + nu))
    return 1/(1 - nu*nu)*np.array([[Ex, nu*Ex, 0],
                                   [nu*Ey, Ey, 0],
                                   [0, 0, (1 - nu*nu)*G]])


def jacobian(xy, r, s):
    """2x2 Jacobian of the bilinear map at (r, s) (`Quad3D.py:276-286`, `Plate3D.py:136-146`)."""
    (x1, y1), (x2, y2), (x3, y3), (x4, y4) = xy
    return 1/4*np.array([[x1*(s - 1) - x2*(s - 1) + x3*(s + 1) - x4*(s + 1),
                          y1*(s - 1) - y2*(s - 1) + y3*(s + 1) - y4*(s + 1)],
                         [x1*(r - 1) - x2*(r + 1) + x3*(r + 1) - x4*(r - 1),
                          y1*(r - 1) - y2*(r + 1) + y3*(r + 1) - y4*(r - 1)]])


def membrane_B(xy, r, s):
    """3x8 membrane strain-displacement matrix (`Quad3D.py:450-465`, `Plate3D.py:148-163`)."""
    dH = np.matmul(inv(jacobian(xy, r, s)), 1/4*np.array([[s - 1, -s + 1, s + 1, -s - 1],
                                                          [r - 1, -r - 1, r + 1, -r + 1]]))
    B = np.zeros((3, 8))
    for a in range(4):
        B[0, 2*a] = dH[0, a]
        B[1, 2*a + 1] = dH[1, a]
        B[2, 2*a] = dH[1, a]
        B[2, 2*a + 1] = dH[0, a]
    return B


def membrane_ke(xy, t, C):
    """8x8 membrane stiffness by 2x2 Gauss quadrature (`Quad3D.py:635-668`, `Plate3D.py:171-195`)."""
    k = np.zeros((8, 8))
    for r, s in _GAUSS:
        B = membrane_B(xy, r, s)
        k = k + np.matmul(B.T, np.matmul(C, B))*det(jacobian(xy, r, s))
    return t*k


# local dof slots of a shell node block (dx, dy, dz, rx, ry, rz)
def _expand_membrane(k8):
    """8x8 (u, v per node) -> 24x24: u -> 6a, v -> 6a+1 (`Quad3D.py:670-698`)."""
    idx = [6*(i//2) + (i % 2) for i in range(8)]
    out = np.zeros((24, 24))
    out[np.ix_(idx, idx)] = k8
    return out


def _bending_slots():
    """(w, tx, ty) per node -> slots 6a+2, 6a+3, 6a+4 (`Quad3D.py:586-615`, `Plate3D.py:276-305`)."""
    return [6*(i//3) + 2 + (i % 3) for i in range(12)]


# ------------------------------------------------------------------------------------------------
# DKMQ quadrilateral (Quad3D.py)
# ------------------------------------------------------------------------------------------------
def quad_local_xy(p1, p2, p3, p4):
    """In-plane coordinates of the four corners (`Quad3D.py:107-141`)."""
    p1, p2, p3, p4 = (np.asarray(p, dtype=float) for p in (p1, p2, p3, p4))
    v12, v13, v14 = p2 - p1, p3 - p1, p4 - p1
    x_axis = v12
    z_axis = np.cross(x_axis, v13)
    y_axis = np.cross(z_axis, x_axis)
    x_axis = x_axis/norm(x_axis)
    y_axis = y_axis/norm(y_axis)
    return ((0.0, 0.0),
            (np.dot(v12, x_axis), np.dot(v12, y_axis)),
            (np.dot(v13, x_axis), np.dot(v13, y_axis)),
            (np.dot(v14, x_axis), np.dot(v14, y_axis)))


def _quad_edges(xy):
    """Edge lengths and direction cosines for edges 5..8 (`Quad3D.py:143-177`)."""
    Lk, Ck, Sk = [], [], []
    for a, b in ((0, 1), (1, 2), (2, 3), (3, 0)):
        dx, dy = xy[b][0] - xy[a][0], xy[b][1] - xy[a][1]
        L = (dx**2 + dy**2)**0.5
        Lk.append(L)
        Ck.append(dx/L)
        Sk.append(dy/L)
    return Lk, Ck, Sk


def _quad_Au(Lk, Ck, Sk):
    """4x12 edge-freedom matrix (`Quad3D.py:307-325`)."""
    A = np.zeros((4, 12))
    for e, (a, b) in enumerate(((0, 1), (1, 2), (2, 3), (3, 0))):
        A[e, 3*a:3*a+3] = [-2/Lk[e], Ck[e], Sk[e]]
        A[e, 3*b:3*b+3] = [2/Lk[e], Ck[e], Sk[e]]
    return 1/2*A


def quad_ke_bending(xy, t, E, nu):
    """12x12 DKMQ bending + shear stiffness, then 24x24 with drilling springs (`Quad3D.py:523-633`)."""
    Lk, Ck, Sk = _quad_edges(xy)
    kappa = 5/6
    phi = [2/(kappa*(1 - nu))*(t/L)**2 for L in Lk]                       # Quad3D.py:179-184
    A_Dinv = -3/2*np.diag([1/(1 + p) for p in phi])                      # :327-337
    A_phiD = np.diag([p/(1 + p) for p in phi])                           # :339-349
    A_gam = np.diag([Lk[0]/2, Lk[1]/2, -Lk[2]/2, -Lk[3]/2])               # :294-305
    A_u = _quad_Au(Lk, Ck, Sk)
    Hb = E*t**3/(12*(1 - nu**2))*np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu)/2]])   # :467-481
    Hs = E*t*kappa/(2*(1 + nu))*np.eye(2)                                # :483-496

    def B_b(r, s):
        Ji = inv(jacobian(xy, r, s))
        dN = 0.25*np.array([[s - 1, -(s - 1), s + 1, -(s + 1)],
                            [r - 1, -(r + 1), r + 1, -(r - 1)]])
        Nxy = Ji @ dN                                                    # rows d/dx, d/dy
        Bb = np.zeros((3, 12))
        for a in range(4):
            Bb[0, 3*a + 1] = Nxy[0, a]
            Bb[1, 3*a + 2] = Nxy[1, a]
            Bb[2, 3*a + 1] = Nxy[1, a]
            Bb[2, 3*a + 2] = Nxy[0, a]
        dP = np.array([[r*(s - 1), -0.5*(s - 1)*(s + 1), -r*(s + 1), 0.5*(s - 1)*(s + 1)],
                       [0.5*(r - 1)*(r + 1), -s*(r + 1), -0.5*(r - 1)*(r + 1), s*(r - 1)]])
        Pxy = Ji @ dP
        BD = np.array([[Pxy[0, e]*Ck[e] for e in range(4)],
                       [Pxy[1, e]*Sk[e] for e in range(4)],
                       [Pxy[1, e]*Ck[e] + Pxy[0, e]*Sk[e] for e in range(4)]])
        return np.add(Bb, BD @ A_Dinv @ A_u)                              # :424-430

    def B_s(r, s):
        Ng = np.array([[1/2*(1 - s), 0, 1/2*(1 + s), 0],
                       [0, 1/2*(1 + r), 0, 1/2*(1 - r)]])
        return inv(jacobian(xy, r, s)) @ Ng @ A_gam @ A_phiD @ A_u       # :432-438

    dets = [det(jacobian(xy, r, s)) for r, s in _GAUSS]
    k = sum((B_b(r, s).T @ Hb @ B_b(r, s))*d for (r, s), d in zip(_GAUSS, dets))
    k = k + sum((B_s(r, s).T @ Hs @ B_s(r, s))*d for (r, s), d in zip(_GAUSS, dets))

    k_rz = min(abs(k[i, i]) for i in (1, 2, 4, 5, 7, 8, 10, 11))/1000    # :577-579
    out = np.zeros((24, 24))
    slots = _bending_slots()
    out[np.ix_(slots, slots)] = k
    for a in range(4):
        out[6*a + 5, 6*a + 5] = k_rz
    # sign flip of the ty rows/cols, then swap rx <-> ry slots (:624-631)
    flip = [4, 10, 16, 22]
    out[flip, :] *= -1
    out[:, flip] *= -1
    perm = np.arange(24)
    for a in range(4):
        perm[6*a + 3], perm[6*a + 4] = 6*a + 4, 6*a + 3
    out = out[perm, :][:, perm]
    return out


def quad_ke(p1, p2, p3, p4, t, E, nu, kx_mod, ky_mod):
    """Local 24x24 quad stiffness (`Quad3D.py:702-711`)."""
    xy = quad_local_xy(p1, p2, p3, p4)
    km = _expand_membrane(membrane_ke(xy, t, membrane_C(E, nu, kx_mod, ky_mod)))
    return np.add(quad_ke_bending(xy, t, E, nu), km)


def quad_fer(p1, p2, p3, p4, pressure):
    """Local 24x1 fixed-end vector for a uniform pressure (already summed over cases with factors).

    `Quad3D.py:721-788`: p = -pressure; f_w,i = sum_gp N_i p det J; w slots only.
    """
    xy = quad_local_xy(p1, p2, p3, p4)
    p = -pressure
    f = np.zeros((24, 1))
    for r, s in _GAUSS:
        Ni = 1/4*np.array([(1 - r)*(1 - s), (1 + r)*(1 - s), (1 + r)*(1 + s), (1 - r)*(1 + s)])
        dJ = det(jacobian(xy, r, s))
        for a in range(4):
            f[6*a + 2, 0] += Ni[a]*p*dJ
    return f


# ------------------------------------------------------------------------------------------------
# rectangular plate (Plate3D.py)
# ------------------------------------------------------------------------------------------------
def plate_ke(width, height, t, E, nu, kx_mod, ky_mod):
    """Local 24x24 plate stiffness (`Plate3D.py:165-314`).

    Bending: 12-term polynomial rectangle, isotropic (E unmodified, `Plate3D.py:237-238`), built
    from the rational tables derived in `tools/derive_plate_tables.py`.
    """
    b, c = width/2, height/2
    G = E/(2*(1 + nu))
    Dx = E*t**3/(12*(1 - nu*nu))
    D1 = nu*Dx
    Dxy = G*t**3/12
    core = (Dx*c/b**3*_pt.A1 + Dx*b/c**3*_pt.A2 + D1/(b*c)*_pt.A3 + Dxy/(b*c)*_pt.A4)/_pt.DEN
    S = np.array([1.0, c, b]*4)
    kb = core*np.outer(S, S)
    k_rz = min(abs(kb[i, i]) for i in (1, 2, 4, 5, 7, 8, 10, 11))/1000    # Plate3D.py:262-264
    out = np.zeros((24, 24))
    slots = _bending_slots()
    out[np.ix_(slots, slots)] = kb
    for a in range(4):
        out[6*a + 5, 6*a + 5] = k_rz
    xy = ((0, 0), (width, 0), (width, height), (0, height))               # Plate3D.py:142
    km = _expand_membrane(membrane_ke(xy, t, membrane_C(E, nu, kx_mod, ky_mod)))
    return np.add(out, km)


def plate_fer(width, height, pressure):
    """Local 24x1 fixed-end vector for a uniform pressure (`Plate3D.py:324-393`)."""
    b, c = width/2, height/2
    v = -4*pressure*c*b*np.array([1/4, c/12, -b/12, 1/4, c/12, b/12, 1/4, -c/12, b/12, 1/4, -c/12, -b/12])
    f = np.zeros((24, 1))
    f[_bending_slots(), 0] = v
    return f


# ------------------------------------------------------------------------------------------------
# internal forces / stresses of shells (Quad3D.py:1017-1239, Plate3D.py:524-692)
# ------------------------------------------------------------------------------------------------
def _quad_bending_ops(xy, t, nu):
    """B_b(r, s) and B_s(r, s) of the DKMQ element as closures over the element geometry (`Quad3D.py:288-438`)."""
    Lk, Ck, Sk = _quad_edges(xy)
    kappa = 5/6
    phi = [2/(kappa*(1 - nu))*(t/L)**2 for L in Lk]
    A_Dinv = -3/2*np.diag([1/(1 + p) for p in phi])
    A_phiD = np.diag([p/(1 + p) for p in phi])
    A_gam = np.diag([Lk[0]/2, Lk[1]/2, -Lk[2]/2, -Lk[3]/2])
    A_u = _quad_Au(Lk, Ck, Sk)

    def B_b(r, s):
        Ji = inv(jacobian(xy, r, s))
        dN = 0.25*np.array([[s - 1, -(s - 1), s + 1, -(s + 1)],
                            [r - 1, -(r + 1), r + 1, -(r - 1)]])
        Nxy = Ji @ dN
        Bb = np.zeros((3, 12))
        for a in range(4):
            Bb[0, 3*a + 1] = Nxy[0, a]
            Bb[1, 3*a + 2] = Nxy[1, a]
            Bb[2, 3*a + 1] = Nxy[1, a]
            Bb[2, 3*a + 2] = Nxy[0, a]
        dP = np.array([[r*(s - 1), -0.5*(s - 1)*(s + 1), -r*(s + 1), 0.5*(s - 1)*(s + 1)],
                       [0.5*(r - 1)*(r + 1), -s*(r + 1), -0.5*(r - 1)*(r + 1), s*(r - 1)]])
        Pxy = Ji @ dP
        BD = np.array([[Pxy[0, e]*Ck[e] for e in range(4)],
                       [Pxy[1, e]*Sk[e] for e in range(4)],
                       [Pxy[1, e]*Ck[e] + Pxy[0, e]*Sk[e] for e in range(4)]])
        return np.add(Bb, BD @ A_Dinv @ A_u)

    def B_s(r, s):
        Ng = np.array([[1/2*(1 - s), 0, 1/2*(1 + s), 0],
                       [0, 1/2*(1 + r), 0, 1/2*(1 - r)]])
        return inv(jacobian(xy, r, s)) @ Ng @ A_gam @ A_phiD @ A_u
    return B_b, B_s


def _extrapolation_H(xi, eta):
    """Bilinear functions at (xi, eta)/gp: Gauss-point values -> any point (`Quad3D.py:1050-1059`)."""
    x, e = xi/GP, eta/GP
    return 1/4*np.array([(1 - x)*(1 - e), (1 + x)*(1 - e), (1 + x)*(1 + e), (1 - x)*(1 + e)])


def quad_stress(p1, p2, p3, p4, t, E, nu, kx_mod, ky_mod, d, xi, eta, local=True):
    """(shear, moment, membrane) of a quad at natural coordinates (xi, eta) from its LOCAL displacement vector d
    (24, 1) (`Quad3D.shear` :1017-1094, `moment` :1096-1171, `membrane` :1173-1239)."""
    xy = quad_local_xy(p1, p2, p3, p4)
    B_b, B_s = _quad_bending_ops(xy, t, nu)
    db = np.array(d, dtype=float).reshape(24, 1).copy()
    db[[3, 9, 15, 21], :] *= -1
    db = db[[2, 4, 3, 8, 10, 9, 14, 16, 15, 20, 22, 21], :]
    H = _extrapolation_H(xi, eta)
    Hs = E*t*(5/6)/(2*(1 + nu))*np.eye(2)
    Hb = E*t**3/(12*(1 - nu**2))*np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu)/2]])
    q = [np.matmul(Hs, np.matmul(B_s(r, s), db)) for r, s in _GAUSS]
    m = [np.matmul(Hb, np.matmul(B_b(r, s), db)) for r, s in _GAUSS]
    dm = np.asarray(d, dtype=float).reshape(24, 1)[[0, 1, 6, 7, 12, 13, 18, 19], :]
    C = membrane_C(E, nu, kx_mod, ky_mod)
    sg = [np.matmul(C, np.matmul(membrane_B(xy, r, s), dm)) for r, s in _GAUSS]
    Q = np.array([sum(H[g]*q[g][k] for g in range(4)) for k in range(2)]).reshape(2)
    M = np.array([sum(H[g]*m[g][k] for g in range(4)) for k in range(3)]).reshape(3)
    S = np.array([sum(H[g]*sg[g][k] for g in range(4)) for k in range(3)]).reshape(3)
    if local:
        return Q, M, S
    R = shell_dircos(p1, p2, p4)
    Qg = np.matmul(R.T, np.array([[Q[0], 0, 0], [0, Q[1], 0], [0, 0, 0]]))
    Qo = np.array([Qg[0, 0], Qg[1, 1], Qg[2, 1]])
    Mg = R @ np.array([[M[0], M[2], 0.0], [M[2], -M[1], 0.0], [0.0, 0.0, 0.0]]) @ R.T
    Sg = R @ np.array([[S[0], S[2], 0.0], [S[2], S[1], 0.0], [0.0, 0.0, 0.0]]) @ R.T
    return Qo, np.array([Mg[0, 0], Mg[1, 1], Mg[0, 1]]), np.array([Sg[0, 0], Sg[1, 1], Sg[0, 1]])


def plate_C(width, height):
    """Displacement coefficient matrix [C] of the rectangular plate (`Plate3D._C` :524-557)."""
    rows = []
    for x, y in ((0, 0), (width, 0), (width, height), (0, height)):
        rows.append([1, x, y, x**2, x*y, y**2, x**3, x**2*y, x*y**2, y**3, x**3*y, x*y**3])
        rows.append([0, 0, 1, 0, x, 2*y, 0, x**2, 2*x*y, 3*y**2, x**3, 3*x*y**2])
        rows.append([0, -1, 0, -2*x, -y, 0, -3*x**2, -2*x*y, -y**2, 0, -3*x**2*y, -y**3])
    return np.array(rows, dtype=float)


def plate_stress(width, height, t, E, nu, kx_mod, ky_mod, d, x, y):
    """(shear, moment, membrane) of a rectangular plate at local (x, y) from its LOCAL displacement vector
    (`Plate3D._a` :573-588, `moment` :590-606, `shear` :608-648, `membrane` :650-692)."""
    d = np.asarray(d, dtype=float).reshape(24, 1)
    a = inv(plate_C(width, height)) @ d[[2, 3, 4, 8, 9, 10, 14, 15, 16, 20, 21, 22], :]
    Ex, Ey, G = E*kx_mod, E*ky_mod, E/(2*(1 + nu))
    Db = t**3/(12*(1 - nu*nu))*np.array([[Ex, nu*Ex, 0], [nu*Ey, Ey, 0], [0, 0, G]])
    Qm = np.array([[0, 0, 0, -2, 0, 0, -6*x, -2*y, 0, 0, -6*x*y, 0],
                   [0, 0, 0, 0, 0, -2, 0, 0, -2*x, -6*y, 0, -6*x*y],
                   [0, 0, 0, 0, -2, 0, 0, -4*x, -4*y, 0, -6*x**2, -6*y**2]])
    M = (-Db @ Qm @ a).reshape(3)
    R1 = np.array([[0, 0, 0, 0, 0, 0, -6, 0, 0, 0, -6*y, 0],
                   [0, 0, 0, 0, 0, 0, 0, 0, -2, 0, 0, -6*y],
                   [0, 0, 0, 0, 0, 0, 0, -4, 0, 0, -12*x, 0]])
    R2 = np.array([[0, 0, 0, 0, 0, 0, 0, -2, 0, 0, -6*x, 0],
                   [0, 0, 0, 0, 0, 0, 0, 0, 0, -6, 0, -6*x],
                   [0, 0, 0, 0, 0, 0, 0, 0, -4, 0, 0, -12*y]])
    v1, v2 = (Db @ R1 @ a).reshape(3), (Db @ R2 @ a).reshape(3)
    Q = np.array([v1[0] + v2[2], v2[1] + v1[2]])
    r, s = -1 + 2*x/width, -1 + 2*y/height
    H = _extrapolation_H(r, s)
    xy = ((0, 0), (width, 0), (width, height), (0, height))
    dm = d[[0, 1, 6, 7, 12, 13, 18, 19], :]
    C = membrane_C(E, nu, kx_mod, ky_mod)
    sg = [np.matmul(C, np.matmul(membrane_B(xy, rr, ss), dm)) for rr, ss in _GAUSS]
    S = np.array([sum(H[g]*sg[g][k] for g in range(4)) for k in range(3)]).reshape(3)
    return Q, M, S
