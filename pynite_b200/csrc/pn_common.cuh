// Shared definitions for the pynite_b200 CUDA library (sm_100a).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cmath>

#if defined(__CUDACC__)
#define PN_HD __host__ __device__ __forceinline__
#define PN_D __device__ __forceinline__
#else
// The element math headers are also compiled by g++ into tests/host_harness (test infrastructure
// only: it lets the arithmetic be checked against the oracle on a box without a GPU).
#define PN_HD inline
#define PN_D inline
#endif

// Fused multiply-add used in the hot accumulation loops.  The library is compiled with
// -fmad=false so that geometry code (cross products, a*b - c*d) rounds like the reference's numpy
// arithmetic and produces the same exact zeros; inner products ask for FMA explicitly.
PN_HD double pn_fma(double a, double b, double c) { return fma(a, b, c); }

// math.isclose(a, b) with Python's defaults rel_tol = 1e-9, abs_tol = 0 (Member3D.py:957,968).
PN_HD bool pn_isclose(double a, double b) {
    if (a == b) return true;
    if (fabs(a) > 1.7976931348623157e308 || fabs(b) > 1.7976931348623157e308) return false;
    double diff = fabs(b - a);
    return (diff <= fabs(1e-9 * b)) || (diff <= fabs(1e-9 * a));
}

PN_HD void pn_cross(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

PN_HD double pn_norm3(const double *a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// G = R^T B R for 3x3 blocks (row-major).  This is one diagonal-block pair of inv(T) k T
// (Member3D.py:1046, Quad3D.py:867, Plate3D.py:508): T is block-diagonal with the same rotation R.
PN_HD void pn_congruence3(const double *R, const double *B, double *G) {
    double W[9];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            W[3 * p + j] = pn_fma(B[3 * p + 2], R[6 + j], pn_fma(B[3 * p + 1], R[3 + j], B[3 * p] * R[j]));
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            G[3 * i + j] = pn_fma(R[6 + i], W[6 + j], pn_fma(R[3 + i], W[3 + j], R[i] * W[j]));
}

// v_global = R^T v_local, v_local = R v_global for 3-vectors.
PN_HD void pn_rot_to_global(const double *R, const double *v, double *g) {
#pragma unroll
    for (int i = 0; i < 3; ++i) g[i] = pn_fma(R[6 + i], v[2], pn_fma(R[3 + i], v[1], R[i] * v[0]));
}
PN_HD void pn_rot_to_local(const double *R, const double *g, double *v) {
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = pn_fma(R[3 * i + 2], g[2], pn_fma(R[3 * i + 1], g[1], R[3 * i] * g[0]));
}
