// FP64 trailing-update GEMM (the multifrontal solver's k_lu_update: C[lower] -= A B, rank <= 256) in several tilings, on
// the front shapes of BASELINE config 3's upper levels.  Question it answers: what separates the shipped 128 x 64 tile
// (8-byte cp.async, 8-byte fragment loads) from the DMMA pipe's 37 TFLOP/s -- copy granularity, fragment-load count,
// tile shape or occupancy?
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo gemm_variants.cu -o gemm_variants
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

extern __shared__ double dyn_smem[];

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8z(double *smem_dst, const double *gsrc, bool pred) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    int bytes = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc, int bytes) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---------------------------------------------------------------------------------------------------------------------
// V0: the shipped tile (128 x 64, 8 warps of 32 x 32, ring of four 16-wide k slices, 8-byte copies and fragment loads)
// ---------------------------------------------------------------------------------------------------------------------
namespace v0 {
constexpr int BM = 128, BN = 64, BK = 64, SA_LD = BK + 4, SB_LD = BN + 4;
constexpr int SMEM = (BM * SA_LD + BK * SB_LD) * 8;
__device__ __forceinline__ void gemm_tile(const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc, int m, int n, int k,
                                          int tile_m, int tile_n, bool lower_only) {
    double(*sA)[SA_LD] = reinterpret_cast<double(*)[SA_LD]>(dyn_smem);
    double(*sB)[SB_LD] = reinterpret_cast<double(*)[SB_LD]>(dyn_smem + BM * SA_LD);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
    const int row0 = tile_m * BM, col0 = tile_n * BN;
    const int wrow = (warp >> 1) * 32, wcol = (warp & 1) * 32;
    double acc[4][4][2];
    const bool warp_on = row0 + wrow < m && col0 + wcol < n && !(lower_only && row0 + wrow + 31 < col0 + wcol);
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        int gr = row0 + wrow + 8 * mi + g;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            int gc = col0 + wcol + 8 * ni + 2 * tg;
            bool ok = gr < m && warp_on;
            acc[mi][ni][0] = (ok && gc < n) ? C[(int64_t)gr * ldc + gc] : 0.0;
            acc[mi][ni][1] = (ok && gc + 1 < n) ? C[(int64_t)gr * ldc + gc + 1] : 0.0;
        }
    }
    const int nsl = (k + 15) >> 4;
    const int ra = tid >> 4, ka = tid & 15, kr = tid >> 6, cb = tid & 63;
    const int a_last = m - 1 - row0, cbc = min(cb, n - 1 - col0);
    const uint32_t lda32 = (uint32_t)lda, ldb32 = (uint32_t)ldb;
    const double *Abase = A + (int64_t)row0 * lda + ka;
    const double *Bbase = B + col0 + cbc;
    auto issue = [&](int sl) {
        if (sl < nsl) {
            const int slot = sl & 3, kbase = sl << 4;
            const double *sa = Abase + kbase;
            const double *sb = Bbase + (int64_t)kbase * ldb;
            double *ta = &sA[ra][16 * slot + ka], *tb = &sB[16 * slot + kr][cb];
#pragma unroll
            for (int i = 0; i < 8; ++i) cp_async8(ta + 16 * i * SA_LD, sa + (uint32_t)min(ra + 16 * i, a_last) * lda32);
#pragma unroll
            for (int i = 0; i < 4; ++i) cp_async8(tb + 4 * i * SB_LD, sb + (uint32_t)(kr + 4 * i) * ldb32);
        }
        cp_async_commit();
    };
    issue(0); issue(1); issue(2);
    for (int sl = 0; sl < nsl; ++sl) {
        cp_async_wait<2>();
        __syncthreads();
        issue(sl + 3);
        const int base = (sl & 3) << 4;
        if (warp_on)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int kk = base + 4 * ks + tg;
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = -1.0 * sA[wrow + 8 * mi + g][kk];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = sB[kk][wcol + 8 * ni + g];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    if (!warp_on) return;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        int gr = row0 + wrow + 8 * mi + g;
        if (gr >= m) continue;
        double *crow = C + (int64_t)gr * ldc;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            int gc = col0 + wcol + 8 * ni + 2 * tg;
            if (gc < n) crow[gc] = acc[mi][ni][0];
            if (gc + 1 < n) crow[gc + 1] = acc[mi][ni][1];
        }
    }
}
}  // namespace v0

// ---------------------------------------------------------------------------------------------------------------------
// V1: 16-byte cp.async, 16-byte fragment loads (k pairs for A, column pairs for B), accumulators hold -C so the loop has
// no sign multiply.  Tile = (WARPS_M x 8 MI) x (WARPS_N x 32); a warp owns (8 MI) x 32 as MI x 4 DMMA tiles.
// SLICES = ring depth in 16-wide k slices.
// ---------------------------------------------------------------------------------------------------------------------
template <int MI, int WARPS_M, int WARPS_N, int SLICES, int ABL = 0>   // ABL (ablation): 1 = no global->shared copies, 2 = also no fragment loads
struct V1 {
    static constexpr int NT = 32 * WARPS_M * WARPS_N;
    static constexpr int BM = WARPS_M * 8 * MI, BN = WARPS_N * 32;
    static constexpr int BK = 16 * SLICES;
    static constexpr int SA_LD = BK + 8;          // = 8 mod 16 doubles: the two rows of a quarter-warp's LDS.128 fall into different 64-byte halves
    static constexpr int SB_LD = BN + 2;          // = 2 mod 8 doubles: four k rows (stride 2) x two 16-byte column chunks cover all 8 bank groups
    static constexpr int SMEM = (BM * SA_LD + BK * SB_LD) * 8;

    __device__ static __forceinline__ void gemm_tile(const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc, int m, int n,
                                                     int k, int tile_m, int tile_n, bool lower_only) {
        double *sA = dyn_smem;
        double *sB = dyn_smem + BM * SA_LD;
        const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
        const int row0 = tile_m * BM, col0 = tile_n * BN;
        const int wrow = (warp / WARPS_N) * (8 * MI), wcol = (warp % WARPS_N) * 32;
        double acc[MI][4][2];
        const bool warp_on = row0 + wrow < m && col0 + wcol < n && !(lower_only && row0 + wrow + 8 * MI - 1 < col0 + wcol);
        const bool full_tile = row0 + BM <= m && col0 + BN <= n && (ldc & 1) == 0 && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
        // lane's columns of pair np: c0 = col0 + wcol + 16 np + 4 tg + {0,1,2,3} = acc[mi][2np][0], acc[mi][2np+1][0], acc[mi][2np][1], acc[mi][2np+1][1]
        if (!warp_on) {
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        } else if (full_tile) {
            const double *cp = C + (int64_t)(row0 + wrow + g) * ldc + col0 + wcol + 4 * tg;
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    const double2 lo = *reinterpret_cast<const double2 *>(cp + (int64_t)(8 * mi) * ldc + 16 * np);
                    const double2 hi = *reinterpret_cast<const double2 *>(cp + (int64_t)(8 * mi) * ldc + 16 * np + 2);
                    acc[mi][2 * np][0] = -lo.x; acc[mi][2 * np + 1][0] = -lo.y; acc[mi][2 * np][1] = -hi.x; acc[mi][2 * np + 1][1] = -hi.y;
                }
        } else {
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                const int gr = row0 + wrow + 8 * mi + g;
#pragma unroll
                for (int np = 0; np < 2; ++np)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int gc = col0 + wcol + 16 * np + 4 * tg + q;
                        acc[mi][2 * np + (q & 1)][q >> 1] = (gr < m && gc < n) ? -C[(int64_t)gr * ldc + gc] : 0.0;
                    }
            }
        }
        const int nsl = (k + 15) >> 4;
        // A slice: BM rows x 8 chunks of 16 B; B slice: 16 k rows x BN/2 chunks
        constexpr int A_PER = BM * 8 / NT, B_PER = 16 * (BN / 2) / NT;
        static_assert(BM * 8 % NT == 0 && 16 * (BN / 2) % NT == 0, "copy split");
        const int a_last = m - 1 - row0;
        const int ra = tid >> 3, ca = (tid & 7) * 2;                   // A: rows ra + (NT/8) i, k offset ca
        const int kr = tid / (BN / 2), cb = (tid % (BN / 2)) * 2;      // B: k rows kr + (NT/(BN/2)) i, column cb
        const int ncol = n - col0;                                      // valid columns of the B tile
        const int cbc = min(cb, (ncol - 1) & ~1);
        const int bbytes = min(16, (ncol - cbc) * 8);
        const uint32_t lda32 = (uint32_t)lda, ldb32 = (uint32_t)ldb;
        const double *Abase = A + (int64_t)row0 * lda + ca;
        const double *Bbase = B + col0 + cbc;
        // 16-byte copies need even leading dimensions and 16-byte aligned operand origins; otherwise 8-byte copies into the same layout
        const bool al16 = ((lda | ldb) & 1) == 0 && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0;
        const int ra8 = tid >> 4, ka8 = tid & 15, kr8 = tid / BN, cb8 = tid % BN;
        const int cbc8 = min(cb8, ncol - 1);
        auto issue8 = [&](int sl) {
            if (sl < nsl) {
                const int slot = sl % SLICES, kbase = sl << 4;
                const bool oka = kbase + ka8 < k;
                const double *sa = A + (int64_t)row0 * lda + kbase + ka8;
                double *ta = sA + ra8 * SA_LD + 16 * slot + ka8;
#pragma unroll
                for (int i = 0; i < BM * 16 / NT; ++i)
                    cp_async8z(ta + (NT / 16) * i * SA_LD, oka ? sa + (uint32_t)min(ra8 + (NT / 16) * i, a_last) * lda32 : A, oka);
                const double *sb = B + col0 + cbc8 + (int64_t)kbase * ldb;
                double *tb = sB + (16 * slot + kr8) * SB_LD + cb8;
#pragma unroll
                for (int i = 0; i < 16 * BN / NT; ++i) {
                    const bool ok = kbase + kr8 + (NT / BN) * i < k;
                    cp_async8z(tb + (NT / BN) * i * SB_LD, ok ? sb + (uint32_t)(kr8 + (NT / BN) * i) * ldb32 : B, ok);
                }
            }
            cp_async_commit();
        };
        auto issue16 = [&](int sl) {
            if (sl < nsl) {
                const int slot = sl % SLICES, kbase = sl << 4;
                const double *sa = Abase + kbase;
                const double *sb = Bbase + (int64_t)kbase * ldb;
                double *ta = sA + ra * SA_LD + 16 * slot + ca, *tb = sB + (16 * slot + kr) * SB_LD + cb;
                if (kbase + 16 <= k) {
#pragma unroll
                    for (int i = 0; i < A_PER; ++i) cp_async16(ta + (NT / 8) * i * SA_LD, sa + (uint32_t)min(ra + (NT / 8) * i, a_last) * lda32, 16);
#pragma unroll
                    for (int i = 0; i < B_PER; ++i) cp_async16(tb + (NT / (BN / 2)) * i * SB_LD, sb + (uint32_t)(kr + (NT / (BN / 2)) * i) * ldb32, bbytes);
                } else {
                    const int ab = max(0, min(16, (k - kbase - ca) * 8));
#pragma unroll
                    for (int i = 0; i < A_PER; ++i) cp_async16(ta + (NT / 8) * i * SA_LD, ab ? sa + (uint32_t)min(ra + (NT / 8) * i, a_last) * lda32 : A, ab);
#pragma unroll
                    for (int i = 0; i < B_PER; ++i) {
                        const bool ok = kbase + kr + (NT / (BN / 2)) * i < k;
                        cp_async16(tb + (NT / (BN / 2)) * i * SB_LD, ok ? sb + (uint32_t)(kr + (NT / (BN / 2)) * i) * ldb32 : B, ok ? bbytes : 0);
                    }
                }
            }
            cp_async_commit();
        };
        auto issue = [&](int sl) { if (ABL) return; if (al16) issue16(sl); else issue8(sl); };
#pragma unroll
        for (int i = 0; i < SLICES - 1; ++i) issue(i);
        for (int sl = 0; sl < nsl; ++sl) {
            cp_async_wait<SLICES - 2>();
            __syncthreads();
            issue(sl + SLICES - 1);
            const int base = (sl % SLICES) << 4;
            if (warp_on) {
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    double2 a[MI], b[2][2];
                    if (ABL == 2) {
#pragma unroll
                        for (int mi = 0; mi < MI; ++mi) a[mi] = make_double2(1e-3 * (mi + lane + sl), 1e-3 * (mi - lane));
#pragma unroll
                        for (int e = 0; e < 2; ++e)
#pragma unroll
                            for (int np = 0; np < 2; ++np) b[e][np] = make_double2(1e-3 * (e + lane), 1e-3 * (np - lane + sl));
                    } else {
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi) a[mi] = *reinterpret_cast<const double2 *>(sA + (wrow + 8 * mi + g) * SA_LD + base + 8 * kp + 2 * tg);
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int np = 0; np < 2; ++np)
                            b[e][np] = *reinterpret_cast<const double2 *>(sB + (base + 8 * kp + 2 * tg + e) * SB_LD + wcol + 16 * np + 2 * g);
                    }
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                        for (int np = 0; np < 2; ++np) {
                            dmma884(acc[mi][2 * np][0], acc[mi][2 * np][1], a[mi].x, b[0][np].x);
                            dmma884(acc[mi][2 * np + 1][0], acc[mi][2 * np + 1][1], a[mi].x, b[0][np].y);
                        }
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                        for (int np = 0; np < 2; ++np) {
                            dmma884(acc[mi][2 * np][0], acc[mi][2 * np][1], a[mi].y, b[1][np].x);
                            dmma884(acc[mi][2 * np + 1][0], acc[mi][2 * np + 1][1], a[mi].y, b[1][np].y);
                        }
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        if (!warp_on) return;
        if (full_tile) {
            double *cp = C + (int64_t)(row0 + wrow + g) * ldc + col0 + wcol + 4 * tg;
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    *reinterpret_cast<double2 *>(cp + (int64_t)(8 * mi) * ldc + 16 * np) = make_double2(-acc[mi][2 * np][0], -acc[mi][2 * np + 1][0]);
                    *reinterpret_cast<double2 *>(cp + (int64_t)(8 * mi) * ldc + 16 * np + 2) = make_double2(-acc[mi][2 * np][1], -acc[mi][2 * np + 1][1]);
                }
            return;
        }
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            const int gr = row0 + wrow + 8 * mi + g;
            if (gr >= m) continue;
#pragma unroll
            for (int np = 0; np < 2; ++np)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int gc = col0 + wcol + 16 * np + 4 * tg + q;
                    if (gc < n) C[(int64_t)gr * ldc + gc] = -acc[mi][2 * np + (q & 1)][q >> 1];
                }
        }
    }
};

struct Front { int64_t off; int n, ns; };

// grid = (column tile, row tile, front): F[ge:, ge:] -= F[ge:, g0:ge] F[g0:ge, ge:], lower triangle only
template <int BM, int BN, class Tile>
__global__ void __launch_bounds__(Tile::NT, Tile::MINB) k_update(const Front *fr, double *arena, int g0, int gb) {
    const Front F = fr[blockIdx.z];
    const int ge = min(g0 + gb, F.ns), n = F.n, rest = n - ge;
    if ((int)blockIdx.y * BM >= rest || (int)blockIdx.x * BN >= rest) return;
    if (((int)blockIdx.y + 1) * BM <= (int)blockIdx.x * BN) return;
    double *A = arena + F.off;
    Tile::run(A + (int64_t)ge * n + g0, n, A + (int64_t)g0 * n + ge, n, A + (int64_t)ge * n + ge, n, rest, rest, ge - g0, blockIdx.y, blockIdx.x);
}

struct T0 { static constexpr int NT = 256, MINB = 2, SMEM = v0::SMEM;
    __device__ static __forceinline__ void run(const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc, int m, int n, int k, int tm, int tn) {
        v0::gemm_tile(A, lda, B, ldb, C, ldc, m, n, k, tm, tn, true); } };
template <int MI, int WM, int WN, int SL, int MB, int ABL = 0>
struct T1 { using G = V1<MI, WM, WN, SL, ABL>; static constexpr int NT = G::NT, MINB = MB, SMEM = G::SMEM;
    __device__ static __forceinline__ void run(const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc, int m, int n, int k, int tm, int tn) {
        G::gemm_tile(A, lda, B, ldb, C, ldc, m, n, k, tm, tn, true); } };

static std::vector<double> g_ref;

template <int BM, int BN, class Tile>
void bench(const char *name, const std::vector<Front> &fronts, const Front *dfr, double *arena, const double *arena0, size_t total, int gb, bool check) {
    CK(cudaFuncSetAttribute(k_update<BM, BN, Tile>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile::SMEM));
    int maxrest = 0; double useful = 0;
    for (auto &f : fronts) { int rest = f.n - std::min(gb, f.ns); maxrest = std::max(maxrest, rest); useful += (double)rest * rest * std::min(gb, f.ns); }
    dim3 grid((maxrest + BN - 1) / BN, (maxrest + BM - 1) / BM, (unsigned)fronts.size());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemcpy(arena, arena0, total * 8, cudaMemcpyDeviceToDevice));
        CK(cudaEventRecord(e0));
        k_update<BM, BN, Tile><<<grid, Tile::NT, Tile::SMEM>>>(dfr, arena, 0, gb);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    double err = -1;
    if (check) {
        std::vector<double> h(total);
        CK(cudaMemcpy(h.data(), arena, total * 8, cudaMemcpyDeviceToHost));
        if (g_ref.empty()) g_ref = h;
        else {
            err = 0; double mx = 0;
            for (size_t i = 0; i < total; ++i) { err = std::max(err, std::fabs(h[i] - g_ref[i])); mx = std::max(mx, std::fabs(g_ref[i])); }
            err /= mx;
        }
    }
    printf("  %-34s %8.3f ms  %6.2f TFLOP/s useful (lower triangle)   max rel diff vs v0 %.2e\n", name, best, useful / (best * 1e-3) * 1e-12, err);
}

int main(int argc, char **argv) {
    struct Case { const char *name; int nf, n, ns, gb; } cases[] = {
        {"4 fronts n=4503 ns=1497 rank 256", 4, 4503, 1497, 256},
        {"1 front  n=4503 ns=1497 rank 256", 1, 4503, 1497, 256},
        {"16 fronts n=3822 ns=750 rank 256", 16, 3822, 750, 256},
        {"128 fronts n=1344 ns=192 rank 192", 128, 1344, 192, 256},
        {"512 fronts n=684 ns=96 rank 96", 512, 684, 96, 256},
        {"4 fronts n=4504 ns=1498 rank 256 (even)", 4, 4504, 1498, 256},
    };
    for (auto &cs : cases) {
        std::vector<Front> fronts;
        size_t total = 0;
        for (int i = 0; i < cs.nf; ++i) { fronts.push_back({(int64_t)total, cs.n, cs.ns}); total += (size_t)cs.n * cs.n; total += total & 1; }
        std::vector<double> h(total);
        uint64_t s = 88172645463325252ull;
        for (auto &v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = ((double)(s >> 11) / 9007199254740992.0 - 0.5) * 0.1; }
        double *arena, *arena0; Front *dfr;
        CK(cudaMalloc(&arena, total * 8)); CK(cudaMalloc(&arena0, total * 8)); CK(cudaMalloc(&dfr, fronts.size() * sizeof(Front)));
        CK(cudaMemcpy(arena0, h.data(), total * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dfr, fronts.data(), fronts.size() * sizeof(Front), cudaMemcpyHostToDevice));
        printf("%s\n", cs.name);
        g_ref.clear();
        const bool check = total < (size_t)200e6;
        bench<128, 64, T0>("v0 128x64 8B copies (shipped)", fronts, dfr, arena, arena0, total, cs.gb, check);
        bench<128, 64, T1<4, 4, 2, 4, 2>>("v1 128x64  8w  ring4  2 CTA/SM", fronts, dfr, arena, arena0, total, cs.gb, check);
        bench<128, 64, T1<4, 4, 2, 3, 2>>("v1 128x64  8w  ring3  2 CTA/SM", fronts, dfr, arena, arena0, total, cs.gb, check);
        bench<128, 64, T1<4, 4, 2, 2, 2>>("v1 128x64  8w  ring2  2 CTA/SM", fronts, dfr, arena, arena0, total, cs.gb, check);
        bench<128, 64, T1<4, 4, 2, 3, 2, 1>>("v1 ring3, no global copies", fronts, dfr, arena, arena0, total, cs.gb, false);
        bench<128, 64, T1<4, 4, 2, 3, 2, 2>>("v1 ring3, no copies, no LDS", fronts, dfr, arena, arena0, total, cs.gb, false);
        bench<128, 128, T1<4, 4, 4, 3, 1>>("v1 128x128 16w ring3  1 CTA/SM", fronts, dfr, arena, arena0, total, cs.gb, check);
        bench<64, 64, T1<4, 2, 2, 3, 4>>("v1 64x64 4w ring3  4 CTA/SM", fronts, dfr, arena, arena0, total, cs.gb, check);
        cudaFree(arena); cudaFree(arena0); cudaFree(dfr);
    }
    return 0;
}
