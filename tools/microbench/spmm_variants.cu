// SpMM Y = K11 X on the K11 pattern of BASELINE config 3 (500 x 500 DKMQ quad plate: per interior node the rows DX, DY
// reference the DX, DY columns of the 9 neighbour nodes, DZ, RX, RY the DZ, RX, RY columns, RZ only itself) for 1, 8 and
// 64 right-hand sides: the shipped row-per-lane-group kernel against a pattern-group kernel that loads an X row once for
// all rows of a node that share a column list.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo spmm_variants.cu -o spmm_variants
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int G, int CPL>
__global__ void __launch_bounds__(256)
k_spmm(int64_t rows, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const double *__restrict__ vals,
       const double *__restrict__ X, double *__restrict__ Y, int s) {
    int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t r = gid / G;
    int lane = (int)(gid % G);
    if (r >= rows) return;
    double acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = 0.0;
    int32_t k0 = indptr[r], k1 = indptr[r + 1];
    for (int32_t k = k0; k < k1; ++k) {
        double v = vals[k];
        const double *x = X + (int64_t)indices[k] * s + lane;
#pragma unroll
        for (int c = 0; c < CPL; ++c)
            if (lane + c * G < s) acc[c] = fma(v, x[c * G], acc[c]);
    }
    double *y = Y + r * s + lane;
#pragma unroll
    for (int c = 0; c < CPL; ++c)
        if (lane + c * G < s) y[c * G] = acc[c];
}

// lead[r] = first row of r's node with the same column list (r itself if none)
__global__ void k_pattern_leaders(int64_t n_groups, const int32_t *__restrict__ group, const int32_t *__restrict__ indptr,
                                  const int32_t *__restrict__ indices, int32_t *__restrict__ lead) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int r0 = group[g], R = group[g + 1] - r0;
    for (int i = 0; i < R; ++i) {
        const int ki = indptr[r0 + i], li = indptr[r0 + i + 1] - ki;
        int ld = r0 + i;
        for (int j = 0; j < i; ++j) {
            if (lead[r0 + j] != r0 + j) continue;
            const int kj = indptr[r0 + j];
            if (indptr[r0 + j + 1] - kj != li) continue;
            bool same = true;
            for (int k = 0; k < li && same; ++k) same = indices[ki + k] == indices[kj + k];
            if (same) { ld = r0 + j; break; }
        }
        lead[r0 + i] = ld;
    }
}

// LANES lanes per node; every pattern group of the node walks its column list once.
template <int LANES, int CPL, int MAXR>
__global__ void __launch_bounds__(256)
k_spmm_pattern(int64_t n_groups, const int32_t *__restrict__ group, const int32_t *__restrict__ lead, const int32_t *__restrict__ indptr,
               const int32_t *__restrict__ indices, const double *__restrict__ vals, const double *__restrict__ X, double *__restrict__ Y, int s) {
    const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t g = gid / LANES;
    const int lane = (int)(gid % LANES);
    if (g >= n_groups) return;
    const int r0 = group[g], R = min(group[g + 1] - r0, MAXR);
    int start[MAXR], ld[MAXR];
#pragma unroll
    for (int i = 0; i < MAXR; ++i) { start[i] = i < R ? indptr[r0 + i] : 0; ld[i] = i < R ? lead[r0 + i] - r0 : -1; }
    const int end_last = indptr[r0 + R];
#pragma unroll
    for (int q = 0; q < MAXR; ++q) {
        if (q >= R || ld[q] != q) continue;
        const int k0 = start[q], L = ((q + 1 < R) ? start[q + 1] : end_last) - k0;
        unsigned mask = 0;
#pragma unroll
        for (int i = q; i < MAXR; ++i) mask |= (ld[i] == q) ? (1u << i) : 0u;
        double acc[MAXR][CPL];
#pragma unroll
        for (int i = 0; i < MAXR; ++i)
#pragma unroll
            for (int c = 0; c < CPL; ++c) acc[i][c] = 0.0;
        for (int k = 0; k < L; ++k) {
            const double *xp = X + (int64_t)indices[k0 + k] * s + lane;
            double x[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) x[c] = (lane + c * LANES < s) ? xp[c * LANES] : 0.0;
#pragma unroll
            for (int i = q; i < MAXR; ++i)
                if (mask & (1u << i)) {
                    const double v = vals[start[i] + k];
#pragma unroll
                    for (int c = 0; c < CPL; ++c) acc[i][c] = fma(v, x[c], acc[i][c]);
                }
        }
#pragma unroll
        for (int i = q; i < MAXR; ++i)
            if (mask & (1u << i)) {
                double *y = Y + (int64_t)(r0 + i) * s + lane;
#pragma unroll
                for (int c = 0; c < CPL; ++c)
                    if (lane + c * LANES < s) y[c * LANES] = acc[i][c];
            }
    }
}

// the same two kernels with U entries of a row in flight: indices, values and X rows of U consecutive entries are loaded
// before the first multiply-add (same summation order)
template <int G, int CPL, int U>
__global__ void __launch_bounds__(256)
k_spmm_u(int64_t rows, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const double *__restrict__ vals,
         const double *__restrict__ X, double *__restrict__ Y, int s) {
    int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t r = gid / G;
    int lane = (int)(gid % G);
    if (r >= rows) return;
    double acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = 0.0;
    const int32_t k0 = indptr[r], k1 = indptr[r + 1];
    const bool colok[1] = {true}; (void)colok;
    int32_t k = k0;
    for (; k + U <= k1; k += U) {
        int32_t ci[U]; double v[U], x[U][CPL];
#pragma unroll
        for (int u = 0; u < U; ++u) { ci[u] = indices[k + u]; v[u] = vals[k + u]; }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int c = 0; c < CPL; ++c) x[u][c] = (lane + c * G < s) ? X[(int64_t)ci[u] * s + lane + c * G] : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int c = 0; c < CPL; ++c) acc[c] = fma(v[u], x[u][c], acc[c]);
    }
    for (; k < k1; ++k) {
        const double v = vals[k];
        const double *x = X + (int64_t)indices[k] * s + lane;
#pragma unroll
        for (int c = 0; c < CPL; ++c)
            if (lane + c * G < s) acc[c] = fma(v, x[c * G], acc[c]);
    }
    double *y = Y + r * s + lane;
#pragma unroll
    for (int c = 0; c < CPL; ++c)
        if (lane + c * G < s) y[c * G] = acc[c];
}

template <int LANES, int CPL, int MAXR, int U>
__global__ void __launch_bounds__(256)
k_spmm_pattern_u(int64_t n_groups, const int32_t *__restrict__ group, const int32_t *__restrict__ lead, const int32_t *__restrict__ indptr,
                 const int32_t *__restrict__ indices, const double *__restrict__ vals, const double *__restrict__ X, double *__restrict__ Y, int s) {
    const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t g = gid / LANES;
    const int lane = (int)(gid % LANES);
    if (g >= n_groups) return;
    const int r0 = group[g], R = min(group[g + 1] - r0, MAXR);
    int start[MAXR], ld[MAXR];
#pragma unroll
    for (int i = 0; i < MAXR; ++i) { start[i] = i < R ? indptr[r0 + i] : 0; ld[i] = i < R ? lead[r0 + i] - r0 : -1; }
    const int end_last = indptr[r0 + R];
#pragma unroll
    for (int q = 0; q < MAXR; ++q) {
        if (q >= R || ld[q] != q) continue;
        const int k0 = start[q], L = ((q + 1 < R) ? start[q + 1] : end_last) - k0;
        unsigned mask = 0;
#pragma unroll
        for (int i = q; i < MAXR; ++i) mask |= (ld[i] == q) ? (1u << i) : 0u;
        double acc[MAXR][CPL];
#pragma unroll
        for (int i = 0; i < MAXR; ++i)
#pragma unroll
            for (int c = 0; c < CPL; ++c) acc[i][c] = 0.0;
        int k = 0;
        for (; k + U <= L; k += U) {
            double x[U][CPL];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double *xp = X + (int64_t)indices[k0 + k + u] * s + lane;
#pragma unroll
                for (int c = 0; c < CPL; ++c) x[u][c] = (lane + c * LANES < s) ? xp[c * LANES] : 0.0;
            }
#pragma unroll
            for (int i = q; i < MAXR; ++i)
                if (mask & (1u << i)) {
                    double v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) v[u] = vals[start[i] + k + u];
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int c = 0; c < CPL; ++c) acc[i][c] = fma(v[u], x[u][c], acc[i][c]);
                }
        }
        for (; k < L; ++k) {
            const double *xp = X + (int64_t)indices[k0 + k] * s + lane;
            double x[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) x[c] = (lane + c * LANES < s) ? xp[c * LANES] : 0.0;
#pragma unroll
            for (int i = q; i < MAXR; ++i)
                if (mask & (1u << i)) {
                    const double v = vals[start[i] + k];
#pragma unroll
                    for (int c = 0; c < CPL; ++c) acc[i][c] = fma(v, x[c], acc[i][c]);
                }
        }
#pragma unroll
        for (int i = q; i < MAXR; ++i)
            if (mask & (1u << i)) {
                double *y = Y + (int64_t)(r0 + i) * s + lane;
#pragma unroll
                for (int c = 0; c < CPL; ++c)
                    if (lane + c * LANES < s) y[c * LANES] = acc[i][c];
            }
    }
}

int main() {
    const int n = 500, nn = n + 1;
    // free DOFs: perimeter nodes keep RX, RY, RZ (pinned), interior nodes all six
    std::vector<int32_t> newidx((size_t)nn * nn * 6, -1), group;
    int32_t n1 = 0;
    for (int r = 0; r < nn; ++r)
        for (int c = 0; c < nn; ++c) {
            const bool edge = r == 0 || c == 0 || r == n || c == n;
            group.push_back(n1);
            for (int k = 0; k < 6; ++k)
                if (!(edge && k < 3)) newidx[((size_t)r * nn + c) * 6 + k] = n1++;
        }
    group.push_back(n1);
    std::vector<int32_t> indptr(1, 0), indices;
    std::vector<double> vals;
    uint64_t sd = 88172645463325252ull;
    auto rnd = [&]() { sd ^= sd << 13; sd ^= sd >> 7; sd ^= sd << 17; return (double)(sd >> 11) / 9007199254740992.0 - 0.5; };
    for (int r = 0; r < nn; ++r)
        for (int c = 0; c < nn; ++c)
            for (int k = 0; k < 6; ++k) {
                if (newidx[((size_t)r * nn + c) * 6 + k] < 0) continue;
                for (int dr = -1; dr <= 1; ++dr)
                    for (int dc = -1; dc <= 1; ++dc) {
                        const int rr = r + dr, cc = c + dc;
                        if (rr < 0 || cc < 0 || rr > n || cc > n) continue;
                        for (int kk = 0; kk < 6; ++kk) {
                            const bool couple = (k < 2 && kk < 2) || (k >= 2 && k < 5 && kk >= 2 && kk < 5) || (k == 5 && kk == 5 && dr == 0 && dc == 0);
                            const int32_t col = newidx[((size_t)rr * nn + cc) * 6 + kk];
                            if (couple && col >= 0) { indices.push_back(col); vals.push_back(rnd()); }
                        }
                    }
                indptr.push_back((int32_t)indices.size());
            }
    const int64_t nnz = (int64_t)indices.size(), ng = (int64_t)group.size() - 1;
    printf("n1 %d nnz %lld node groups %lld\n", n1, (long long)nnz, (long long)ng);
    int32_t *d_indptr, *d_indices, *d_group, *d_lead;
    double *d_vals, *X, *Y, *Y2;
    const int smax = 64;
    CK(cudaMalloc(&d_indptr, (n1 + 1) * 4)); CK(cudaMalloc(&d_indices, nnz * 4)); CK(cudaMalloc(&d_group, (ng + 1) * 4)); CK(cudaMalloc(&d_lead, (size_t)n1 * 4));
    CK(cudaMalloc(&d_vals, nnz * 8)); CK(cudaMalloc(&X, (size_t)n1 * smax * 8)); CK(cudaMalloc(&Y, (size_t)n1 * smax * 8)); CK(cudaMalloc(&Y2, (size_t)n1 * smax * 8));
    CK(cudaMemcpy(d_indptr, indptr.data(), (n1 + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_indices, indices.data(), nnz * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_group, group.data(), (ng + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_vals, vals.data(), nnz * 8, cudaMemcpyHostToDevice));
    {
        std::vector<double> hx((size_t)n1 * smax);
        for (auto &v : hx) v = rnd();
        CK(cudaMemcpy(X, hx.data(), hx.size() * 8, cudaMemcpyHostToDevice));
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    {
        CK(cudaEventRecord(e0));
        k_pattern_leaders<<<(unsigned)((ng + 127) / 128), 128>>>(ng, d_group, d_indptr, d_indices, d_lead);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("pattern leaders: %.3f ms\n", ms);
    }
    auto grid_for = [](int64_t t, int b) { return (unsigned)((t + b - 1) / b); };
    for (int s : {1, 8, 16, 32, 64}) {
        const double bytes = nnz * 12.0 + (n1 + 1) * 4.0 + 2.0 * n1 * s * 8.0;
        auto time_it = [&](auto launch, const char *name, double *out) {
            float best = 1e30f;
            for (int rep = 0; rep < 5; ++rep) {
                CK(cudaMemset(out, 0, (size_t)n1 * s * 8));
                CK(cudaEventRecord(e0));
                launch(out);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                best = std::min(best, ms);
            }
            printf("  s=%2d %-34s %7.3f ms  %7.1f GB/s algorithmic\n", s, name, best, bytes / (best * 1e-3) * 1e-9);
        };
        time_it([&](double *out) {
            if (s == 1) k_spmm<1, 1><<<grid_for((int64_t)n1, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 8) k_spmm<8, 1><<<grid_for((int64_t)n1 * 8, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 16) k_spmm<16, 1><<<grid_for((int64_t)n1 * 16, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 32) k_spmm<32, 1><<<grid_for((int64_t)n1 * 32, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
            else k_spmm<32, 2><<<grid_for((int64_t)n1 * 32, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
        }, "row per lane group (shipped)", Y);
        time_it([&](double *out) {
            if (s == 1) k_spmm_pattern<1, 1, 6><<<grid_for(ng, 256), 256>>>(ng, d_group, d_lead, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 8) k_spmm_pattern<8, 1, 6><<<grid_for(ng * 8, 256), 256>>>(ng, d_group, d_lead, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 16) k_spmm_pattern<16, 1, 6><<<grid_for(ng * 16, 256), 256>>>(ng, d_group, d_lead, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 32) k_spmm_pattern<32, 1, 6><<<grid_for(ng * 32, 256), 256>>>(ng, d_group, d_lead, d_indptr, d_indices, d_vals, X, out, s);
            else k_spmm_pattern<32, 2, 6><<<grid_for(ng * 32, 256), 256>>>(ng, d_group, d_lead, d_indptr, d_indices, d_vals, X, out, s);
        }, "pattern groups, LANES = min(s, 32)", Y2);
        time_it([&](double *out) {
            if (s == 1) k_spmm_u<1, 1, 4><<<grid_for((int64_t)n1, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 8) k_spmm_u<8, 1, 4><<<grid_for((int64_t)n1 * 8, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 16) k_spmm_u<16, 1, 4><<<grid_for((int64_t)n1 * 16, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 32) k_spmm_u<32, 1, 4><<<grid_for((int64_t)n1 * 32, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
            else k_spmm_u<32, 2, 4><<<grid_for((int64_t)n1 * 32, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
        }, "row per lane group, 4 in flight", Y2);
        time_it([&](double *out) {
            if (s == 1) k_spmm_u<1, 1, 8><<<grid_for((int64_t)n1, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 8) k_spmm_u<8, 1, 8><<<grid_for((int64_t)n1 * 8, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 16) k_spmm_u<16, 1, 8><<<grid_for((int64_t)n1 * 16, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 32) k_spmm_u<32, 1, 8><<<grid_for((int64_t)n1 * 32, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
            else k_spmm_u<32, 2, 8><<<grid_for((int64_t)n1 * 32, 256), 256>>>(n1, d_indptr, d_indices, d_vals, X, out, s);
        }, "row per lane group, 8 in flight", Y2);
        time_it([&](double *out) {
            if (s == 1) k_spmm_pattern_u<1, 1, 6, 3><<<grid_for(ng, 256), 256>>>(ng, d_group, d_lead, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 8) k_spmm_pattern_u<8, 1, 6, 3><<<grid_for(ng * 8, 256), 256>>>(ng, d_group, d_lead, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 16) k_spmm_pattern_u<16, 1, 6, 3><<<grid_for(ng * 16, 256), 256>>>(ng, d_group, d_lead, d_indptr, d_indices, d_vals, X, out, s);
            else if (s <= 32) k_spmm_pattern_u<32, 1, 6, 3><<<grid_for(ng * 32, 256), 256>>>(ng, d_group, d_lead, d_indptr, d_indices, d_vals, X, out, s);
            else k_spmm_pattern_u<32, 2, 6, 3><<<grid_for(ng * 32, 256), 256>>>(ng, d_group, d_lead, d_indptr, d_indices, d_vals, X, out, s);
        }, "pattern groups, 3 in flight", Y2);
        if (s == 64)
            time_it([&](double *out) { k_spmm_pattern<16, 4, 6><<<grid_for(ng * 16, 256), 256>>>(ng, d_group, d_lead, d_indptr, d_indices, d_vals, X, out, s); },
                    "pattern groups, 16 lanes x 4 cols", Y2);
        if (s == 8)
            time_it([&](double *out) { k_spmm_pattern<4, 2, 6><<<grid_for(ng * 4, 256), 256>>>(ng, d_group, d_lead, d_indptr, d_indices, d_vals, X, out, s); },
                    "pattern groups, 4 lanes x 2 cols", Y2);
        std::vector<double> a((size_t)n1 * s), b((size_t)n1 * s);
        CK(cudaMemcpy(a.data(), Y, a.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), Y2, b.size() * 8, cudaMemcpyDeviceToHost));
        size_t diff = 0;
        for (size_t i = 0; i < a.size(); ++i) diff += a[i] != b[i];
        printf("  s=%2d entries that differ bitwise: %zu of %zu\n", s, diff, a.size());
    }
    return 0;
}
