// Block-level FP64 GEMM on the tensor cores (mma.sync.m8n8k4.f64 = SASS DMMA.8x8x4) with operands in shared
// memory. Shared by the dense per-geometry kernels (ic_dense.cu: every product; ic_sparse.cu: the two dense
// intdim^3 products of Hessian_cart2int).
#pragma once
#include <stdint.h>

namespace tchem_b200 {
namespace {

// ---------------------------------------------------------------------------------------
// DMMA: D(8x8) = A(8x4, row) * B(4x8, col) + C.  Fragments (PTX ISA, mma.m8n8k4 .f64):
//   a : row = lane >> 2, col = lane & 3          b : row = lane & 3, col = lane >> 2
//   c0, c1 : row = lane >> 2, col = 2 * (lane & 3) + {0, 1}
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double & d0, double & d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// C[M x N] = (ACC ? C : 0) + s * opA(A)[M x K] * opB(B)[K x N], all operands in shared memory.
//   TA: A is stored K x M (opA(A)[m][k] = A[k * lda + m]);  TB: B is stored N x K.
//   TRI: 0 = all tiles; 1 = only the 16 x 16 tiles on or below the diagonal (symmetric result of
//        which the lower triangle is wanted); 2 = those tiles, mirrored into the upper triangle.
//   KS:  operands with structural zeros - 1: terms with k < m0 vanish (W^T W, W lower triangular),
//        2: terms with k < n0 vanish (L W).
// Each warp owns 16 x 16 tiles of C (2 x 2 DMMA tiles). Leading dimensions = 4 or 12 (mod 16)
// keep both fragment access patterns free of bank conflicts. Edge tiles read a clamped (in-bounds)
// row / column instead of predicating every load: D[m][n] only depends on row m of A and column n
// of B, and the duplicated rows are simply not stored. K is walked in unpredicated steps of 4 with
// one masked tail step.
// kmask (optional): J's structural zeros skip whole k-steps. side 1: masks[] indexed by the 8-column block of
// C's columns (B = J, k runs over J's rows); side 2: by the 8-row block of C's rows (A = J^T); side 3: masks[]
// indexed by 8-row blocks of J for both C's rows and columns, intersected (J J^T, k runs over J's columns).
// Products with structural zeros are exact zeros, so the result is the same sum.
struct KMask { const uint32_t * masks; int side; };

template <bool TA, bool TB, bool ACC, int TRI, int KS, int WN>
__device__ __forceinline__ void block_dgemm_tiles(int M, int N, int K, const double * A, int lda, const double * B, int ldb,
                                                  double * C, int ldc, double s, KMask km) {
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    constexpr int TW = 8 * WN;                                // tile width (columns of C)
    const int tm = (M + 15) >> 4, tn = (N + TW - 1) / TW;
    const int K4 = K & ~3;
    const int sa = TA ? lda : 1, sb = TB ? 1 : ldb;          // stride between consecutive k
    // tile coordinates, advanced without division. Triangular products (TRI != 0; square, tm == tn) enumerate only the
    // tiles on or below the diagonal, row by row, so that the warps share THOSE evenly (walking the full grid and
    // skipping the upper tiles left some warps with three tiles and others with none)
    int mt = 0, nt = warp;
    if (TRI == 0) { while (nt >= tn) { nt -= tn; mt++; } }
    else          { while (nt > mt) { nt -= mt + 1; mt++; } }
    for (; mt < tm; ) {
        {
            const int m0 = mt * 16, n0 = nt * TW;
            const int kbeg = KS == 1 ? m0 : KS == 2 ? n0 : 0;
            double acc[2][WN][2];
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < WN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
            const double * pa[2], * pb[WN];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int m = min(m0 + 8 * i + gid, M - 1);
                pa[i] = A + (TA ? m : m * lda) + tig * sa;
            }
#pragma unroll
            for (int j = 0; j < WN; j++) {
                const int n = min(n0 + 8 * j + gid, N - 1);
                pb[j] = B + (TB ? n * ldb : n) + tig * sb;
            }
            uint32_t live = 0xffffffffu;                     // k-steps that can contribute to this tile
            if (km.masks) {
                uint32_t mrow = km.side == 1 ? 0xffffffffu : 0u, mcol = km.side == 2 ? 0xffffffffu : 0u;
                if (km.side != 1) {
#pragma unroll
                    for (int i = 0; i < 2; i++)
                        if (m0 + 8 * i < M) mrow |= km.masks[(m0 >> 3) + i];
                }
                if (km.side != 2) {
#pragma unroll
                    for (int j = 0; j < WN; j++)
                        if (n0 + 8 * j < N) mcol |= km.masks[(n0 >> 3) + j];
                }
                live = mrow & mcol;
            }
            if (km.masks) {
                uint32_t steps = live & (uint32_t)((1ull << (K4 >> 2)) - 1ull);
                if (KS != 0) steps &= ~(uint32_t)((1ull << (kbeg >> 2)) - 1ull);
                // software pipelined: the fragments of the next live k-step are loaded before this step's DMMAs
                if (steps) {
                    double a[2], b[WN];
                    int k0 = 4 * (__ffs(steps) - 1);
                    steps &= steps - 1;
#pragma unroll
                    for (int i = 0; i < 2; i++) a[i] = pa[i][k0 * sa];
#pragma unroll
                    for (int j = 0; j < WN; j++) b[j] = pb[j][k0 * sb];
                    while (steps) {
                        k0 = 4 * (__ffs(steps) - 1);
                        steps &= steps - 1;
                        double a2[2], b2[WN];
#pragma unroll
                        for (int i = 0; i < 2; i++) a2[i] = pa[i][k0 * sa];
#pragma unroll
                        for (int j = 0; j < WN; j++) b2[j] = pb[j][k0 * sb];
#pragma unroll
                        for (int i = 0; i < 2; i++)
#pragma unroll
                            for (int j = 0; j < WN; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
#pragma unroll
                        for (int i = 0; i < 2; i++) a[i] = a2[i];
#pragma unroll
                        for (int j = 0; j < WN; j++) b[j] = b2[j];
                    }
#pragma unroll
                    for (int i = 0; i < 2; i++)
#pragma unroll
                        for (int j = 0; j < WN; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
            } else {
#pragma unroll 2
                for (int k0 = kbeg; k0 < K4; k0 += 4) {
                    double a[2], b[WN];
#pragma unroll
                    for (int i = 0; i < 2; i++) a[i] = pa[i][k0 * sa];
#pragma unroll
                    for (int j = 0; j < WN; j++) b[j] = pb[j][k0 * sb];
#pragma unroll
                    for (int i = 0; i < 2; i++)
#pragma unroll
                        for (int j = 0; j < WN; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
            }
            // (unmasked operands can be wider than 128: the tail bit then lies beyond the 32-bit mask, so it is not consulted)
            if (K4 < K && (!km.masks || ((live >> (K4 >> 2)) & 1u))) {
                const bool in = K4 + tig < K;
                const int k0 = in ? K4 : K - 1 - tig;        // clamped address, value masked
                double a[2], b[WN];
#pragma unroll
                for (int i = 0; i < 2; i++) a[i] = in ? pa[i][k0 * sa] : 0.0;
#pragma unroll
                for (int j = 0; j < WN; j++) b[j] = in ? pb[j][k0 * sb] : 0.0;
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < WN; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < WN; j++) {
                    const int m = m0 + 8 * i + gid, n = n0 + 8 * j + 2 * tig;
                    if (m < M) {
                        if (n < N)     C[m * ldc + n]     = (ACC ? C[m * ldc + n] : 0.0) + s * acc[i][j][0];
                        if (n + 1 < N) C[m * ldc + n + 1] = (ACC ? C[m * ldc + n + 1] : 0.0) + s * acc[i][j][1];
                        if (TRI == 2 && nt < mt) {           // square result: M == N
                            if (n < N) C[n * ldc + m] = (ACC ? C[n * ldc + m] : 0.0) + s * acc[i][j][0];
                            if (n + 1 < N) C[(n + 1) * ldc + m] = (ACC ? C[(n + 1) * ldc + m] : 0.0) + s * acc[i][j][1];
                        }
                    }
                }
        }
        nt += nwarps;
        if (TRI == 0) { while (nt >= tn) { nt -= tn; mt++; } }
        else          { while (nt > mt) { nt -= mt + 1; mt++; } }
    }
}

// Warp tiles are 16 x 16 (2 x 2 DMMA tiles) or, for the full (non-triangular) products, 16 x 24 when
// that shape hands the block's warps fewer DMMA steps: C4's 84 x 90 and 90 x 90 results are 36 tiles
// of 16 x 16 (4.5 rounds of 8 warps -> 5) but 24 tiles of 16 x 24 (3 rounds exactly, 10 % fewer DMMAs
// on the critical path and 5 instead of 6 fragment loads per 6 DMMAs).
template <bool TA, bool TB, bool ACC, int TRI = 0, int KS = 0>
__device__ __forceinline__ void block_dgemm(int M, int N, int K, const double * A, int lda, const double * B, int ldb,
                                            double * C, int ldc, double s = 1.0, KMask km = KMask{nullptr, 0}) {
    if (TRI == 0 && KS == 0) {
        const int nwarps = blockDim.x >> 5, tm = (M + 15) >> 4;
        const int steps2 = 2 * ((tm * ((N + 15) / 16) + nwarps - 1) / nwarps);
        const int steps3 = 3 * ((tm * ((N + 23) / 24) + nwarps - 1) / nwarps);
        if (steps3 < steps2) { block_dgemm_tiles<TA, TB, ACC, 0, 0, 3>(M, N, K, A, lda, B, ldb, C, ldc, s, km); return; }
    }
    block_dgemm_tiles<TA, TB, ACC, TRI, KS, 2>(M, N, K, A, lda, B, ldb, C, ldc, s, km);
}

} // namespace
} // namespace tchem_b200
