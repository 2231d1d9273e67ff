// Dense per-geometry transforms: one thread block owns one geometry at a time, J (and the
// other operands) live in shared memory, the contractions run on the FP64 tensor cores
// (mma.sync.m8n8k4.f64 = DMMA).
//
// Replaces, batched (reference source/IC/IntCoordSet.cpp):
//   gradient_cart2int_matrix :178-188   M   = (J J^T)^-1 J
//   gradient_cart2int        :190-205   g_q = M g_r
//   Hessian_int2cart         :248-266   H_r = J^T H_q J + sum_k g_q[k] K[k]
//   Hessian_cart2int         :221-246   H_q = A^T H_r A - sum_k (A^T g_r)[k] A^T K[k] A,  A^T = M
// K (intdim x cartdim x cartdim per geometry) is never materialised: the second-derivative
// blocks of every term are formed in registers (Dual passes through the analytic J expression,
// what the reference asks autograd for at source/IC/InvDisp.cpp:859-868) and added, weighted by
// g_q[k] * coeff, straight into the cartdim x cartdim accumulator. Hessian_cart2int uses
//   sum_k g_k A^T K_k A = A^T (sum_k g_k K_k) A
// instead of the reference's intdim separate triple products.
// (J J^T)^-1 follows the reference's algorithm - Cholesky factor, explicit inverse of the
// factor, L^-T L^-1 (LAPACK potrf + potri behind at::cholesky / at::cholesky_inverse) - because a
// mathematically equivalent solve already differs from it by ~cond * eps.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "dmma_gemm.cuh"
#include "ic_eval.cuh"
#include "ic_table.h"

namespace tchem_b200 {

// ic_sparse.cu: the same four transforms with J kept sparse; -1 = not taken (this file's kernels serve the call)
int launch_sparse(int op, const tchem_ic_table * t, const double * r, const double * in1, const double * in2, int64_t B,
                  double * out, cudaStream_t stream);
int sparse_status(const tchem_ic_table * t);

struct DenseTerm {
    PrimDesc pd;       // pd.out = internal coordinate (row) of the term
    double coef;
};
struct DenseTask { int32_t term, dir; };
struct DenseProgram {                 // device pointers
    const DenseTerm * terms;
    const int32_t *   row_ptr;        // CSR over rows
    const DenseTask * tasks;          // (term, Cartesian direction) pairs for the K passes
    int32_t nterms, ntasks, intdim, cartdim;
    int32_t ldj, ldh;                 // leading dimensions: width >= cartdim / intdim, = 4 or 12 mod 16
    // structural sparsity of J at DMMA k-step granularity (null when intdim or cartdim > 128):
    //   colmask[cb], cb = 8-column block of J: bit s set when rows 4s..4s+3 of J have a structural nonzero there
    //   rowmask[rb], rb = 8-row block of J:    bit s set when columns 4s..4s+3 have one
    const uint32_t *  colmask;
    const uint32_t *  rowmask;
    int32_t blob_doubles;             // size of the table blob (terms | row_ptr | tasks | masks, starts at `terms`)
    int32_t tab_doubles;              // per launch: blob_doubles when the kernel keeps a copy in shared memory, else 0
};

namespace {

constexpr int kDenseThreads = 256;

// -DTCHEM_DENSE_PHASES: block 0 accumulates the clock cycles of each phase of its geometries in
// g_dense_phase (read back through tchem_debug_dense_phases; tools/dense_phases.py). Off in the product.
#ifdef TCHEM_DENSE_PHASES
__device__ unsigned long long g_dense_phase[32];
// (a clock read right after __syncthreads captures the barrier's issue, not its release: the read is made
// to depend on a shared-memory load, which is where the deferred block fires)
__device__ __forceinline__ long long phase_clock(const int * probe) {
    const int x = *reinterpret_cast<const volatile int *>(probe);
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(x) : "memory");
    return t;
}
#define PHASE_BEGIN() __shared__ int ph_probe; long long ph_t0 = phase_clock(&ph_probe)
#define PHASE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long ph_t = phase_clock(&ph_probe); g_dense_phase[i] += (unsigned long long)(ph_t - ph_t0); ph_t0 = ph_t; } } while (0)
#else
#define PHASE_BEGIN() do {} while (0)
#define PHASE(i) do {} while (0)
#endif

// ---------------------------------------------------------------------------------------
// sinks of the per-term evaluation
// ---------------------------------------------------------------------------------------
struct RowSink {                       // J[row][.] += coef * J_term   (row owned by one thread)
    double * row;
    double coef;
    __device__ __forceinline__ void value(double) const {}
    __device__ __forceinline__ void jac_block(const int * atoms, int i, const V3<double> & J) const {
        double * p = row + 3 * atoms[i];
        p[0] = fma(coef, J.x, p[0]);
        p[1] = fma(coef, J.y, p[1]);
        p[2] = fma(coef, J.z, p[2]);
    }
    __device__ __forceinline__ void jac_entry(const int * atoms, int m, double j) const {
        const int a = m / 3;
        double * p = row + 3 * atoms[a] + (m - 3 * a);
        p[0] = fma(coef, j, p[0]);
    }
    template <int N> __device__ __forceinline__ void hess_block(const int *, int, int, const V3<Dual> &) const {}
    __device__ __forceinline__ void hess_entry5(const int *, int, int, double) const {}
};

struct HessSink {                      // H[.][.] += w * K_term   (shared accumulator, atomics)
    double * H;
    int ld;
    double w;
    __device__ __forceinline__ void value(double) const {}
    __device__ __forceinline__ void jac_block(const int *, int, const V3<double> &) const {}
    __device__ __forceinline__ void jac_entry(const int *, int, double) const {}
    __device__ __forceinline__ void add(int r, int c, double v, bool mirror) const {
#ifdef TCHEM_GK_PLAIN        // timing experiment only (racy): what the phase costs without the CAS loops
        H[r * ld + c] += v;
        if (mirror) H[c * ld + r] += v;
#else
        atomicAdd(H + r * ld + c, v);
        if (mirror) atomicAdd(H + c * ld + r, v);
#endif
    }
    // tangent direction (atom j, component l) of block i: entry (k, l) of block (i, j), i <= j,
    // mirrored into block (j, i) like source/IC/InvDisp.cpp:869-876
    template <int N> __device__ __forceinline__ void hess_block(const int * atoms, int dir, int i, const V3<Dual> & J) const {
        const int j = dir / 3, l = dir - 3 * j;
        if (i > j) return;
        const int r = 3 * atoms[i], c = 3 * atoms[j] + l;
        add(r, c, w * J.x.d, i < j);
        add(r + 1, c, w * J.y.d, i < j);
        add(r + 2, c, w * J.z.d, i < j);
    }
    __device__ __forceinline__ void hess_entry5(const int * atoms, int m1, int m2, double k) const {
        const int a1 = m1 / 3, a2 = m2 / 3;
        add(3 * atoms[a1] + (m1 - 3 * a1), 3 * atoms[a2] + (m2 - 3 * a2), w * k, m1 != m2);
    }
};

// J[intdim][ldj] of the block's geometry: zero, then one thread per internal coordinate adds its
// terms (source/IC/IntCoord.cpp:62-76)
template <bool HAS5, class Idle>
__device__ void block_compute_J(const DenseProgram & dp, const double * rg, double * J, Idle idle_warps) {
    for (int i = threadIdx.x; i < dp.intdim * dp.ldj; i += blockDim.x) J[i] = 0.0;
    __syncthreads();
    // rows occupy the first ceil(intdim / 32) warps; the others run idle_warps(first idle warp) meanwhile
    // (the H_q copy: issued from here it does not queue ahead of the row threads' table reads)
    const int row_warps = (dp.intdim + kLanes - 1) / kLanes;
    if ((int)(threadIdx.x >> 5) >= row_warps) idle_warps(row_warps);
    for (int row = threadIdx.x; row < dp.intdim; row += blockDim.x) {
        RowSink sink;
        sink.row = J + row * dp.ldj;
        for (int t = dp.row_ptr[row]; t < dp.row_ptr[row + 1]; t++) {
            sink.coef = dp.terms[t].coef;
            eval_primitive<MODE_QJ, HAS5>(dp.terms[t].pd, GeomLoader{rg}, sink);
        }
    }
    __syncthreads();
}
template <bool HAS5>
__device__ void block_compute_J(const DenseProgram & dp, const double * rg, double * J) {
    block_compute_J<HAS5>(dp, rg, J, [](int) {});
}

// H[cartdim][ld] += sum over terms of weight[row] * coef * K_term, one (term, direction) per thread
template <bool HAS5>
__device__ void block_add_gK(const DenseProgram & dp, const double * rg, const double * weight, double sign,
                             double * H, int ld) {
    for (int k = threadIdx.x; k < dp.ntasks; k += blockDim.x) {
        const DenseTask task = dp.tasks[k];
        const DenseTerm & term = dp.terms[task.term];
        HessSink sink;
        sink.H = H; sink.ld = ld;
        sink.w = sign * weight[term.pd.out] * term.coef;
        eval_primitive<MODE_KDIR, HAS5>(term.pd, GeomLoader{rg}, sink, task.dir, task.dir + 1);
    }
    __syncthreads();
}

// global [rows][cols] -> shared [rows][ld], asynchronously (cp.async; 16 bytes per copy when the row
// pitch allows, else 8): the caller overlaps the transfer with arithmetic and calls
// block_load_wait() + __syncthreads() before touching the tile. One warp per row at a time.
// (workspace mode - operands in global memory because the molecule is too large for shared memory -
// copies synchronously: cp.async only writes shared memory)
// Warps [warp0, nwarps_total) take part (default: all of them).
__device__ __forceinline__ void block_load_async(const double * __restrict__ g, double * s, int rows, int cols, int ld, bool in_smem = true,
                                                 int warp0 = 0) {
    const int warp = (threadIdx.x >> 5) - warp0, nwarps = (blockDim.x >> 5) - warp0, lane = threadIdx.x & 31;
    if (warp < 0) return;
    if (!in_smem) {
        for (int r = warp; r < rows; r += nwarps)
            for (int c = lane; c < cols; c += kLanes) s[r * ld + c] = g[(int64_t)r * cols + c];
        return;
    }
    const bool vec = ((cols | ld) & 1) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s)) & 15) == 0;
    for (int r = warp; r < rows; r += nwarps) {
        const double * src = g + (int64_t)r * cols;
        double * dst = s + r * ld;
        if (vec) {
            for (int c = 2 * lane; c < cols; c += 2 * kLanes) {
                const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + c);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src + c) : "memory");
            }
        } else {
            for (int c = lane; c < cols; c += kLanes) {
                const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + c);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src + c) : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void block_load_wait() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// shared [rows][ld] -> global [rows][cols], one warp per row at a time (no index division)
__device__ __forceinline__ void block_store(const double * s, double * __restrict__ g, int rows, int cols, int ld) {
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const bool vec = ((cols | ld) & 1) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s)) & 15) == 0;
    if (vec) {
        // two rows per pass: four independent 16-byte loads before the first store
        for (int r = warp; r < rows; r += 2 * nwarps) {
            const int r2 = r + nwarps;
            const double * src = s + r * ld, * src2 = s + r2 * ld;
            double * dst = g + (int64_t)r * cols, * dst2 = g + (int64_t)r2 * cols;
            for (int c = 2 * lane; c < cols; c += 4 * kLanes) {
                const int c2 = c + 2 * kLanes;
                const double2 a = *reinterpret_cast<const double2 *>(src + c);
                double2 b2 = a, a2 = a, b3 = a;
                if (c2 < cols) b2 = *reinterpret_cast<const double2 *>(src + c2);
                if (r2 < rows) {
                    a2 = *reinterpret_cast<const double2 *>(src2 + c);
                    if (c2 < cols) b3 = *reinterpret_cast<const double2 *>(src2 + c2);
                }
                __stcs(reinterpret_cast<double2 *>(dst + c), a);
                if (c2 < cols) __stcs(reinterpret_cast<double2 *>(dst + c2), b2);
                if (r2 < rows) {
                    __stcs(reinterpret_cast<double2 *>(dst2 + c), a2);
                    if (c2 < cols) __stcs(reinterpret_cast<double2 *>(dst2 + c2), b3);
                }
            }
        }
    } else {
        for (int r = warp; r < rows; r += nwarps) {
            const double * src = s + r * ld;
            double * dst = g + (int64_t)r * cols;
            for (int c = lane; c < cols; c += kLanes) __stcs(dst + c, src[c]);
        }
    }
}

// (J J^T)^-1 the way LAPACK does it behind at::cholesky / at::cholesky_inverse: blocked potrf, then
// potri = trtri (explicit inverse of the factor) + lauum (L^-T L^-1), block size 16.
//   A (n x n, lda): in  = the SPD matrix (lower triangle read),  out = its inverse (full, symmetric)
//   W (n x n, ldw): scratch, ends as L^-1 (lower, zeros above)
// Work split: the 16 x 16 diagonal blocks are handled by one warp in registers (the sequential part:
// 16 dependent sqrt / scale steps), the panel below by one thread per row, everything else - the
// trailing updates, the block rows of L^-1, the final product - by DMMA.
// Returns false (all threads) when a pivot is not positive.
constexpr int kNB = 16;

// In-place lower Cholesky factor of the diagonal block at D (bs <= 16 rows); dinv[j] = 1 / L[j][j].
// One warp runs this while the block waits, and what it costs is that warp's instruction count and the
// chain of 16 dependent pivots, so both are kept short:
//  * lane = (row i, column half h): 8 columns of one row in registers, all 32 lanes busy; the pivot column
//    reaches the other lanes through 2 x 16 doubles of scratch (one store, four 16-byte broadcast loads)
//    instead of one shuffle per column;
//  * the columns stay UNSCALED during the elimination (u = L sqrt(d), the L D L^T recurrence), which leaves
//    on the chain, per pivot, one reciprocal, one multiply and one fma - local to the lane that owns the
//    next pivot - and one broadcast:   d[j+1] = a[j+1][j+1] - u[j+1][j] * (u[j+1][j] / d[j]);
//    the 16 values 1 / sqrt(d[j]) that turn u into L are taken once, in parallel, at the end.
// Rows / columns >= bs carry an identity so every lane runs the same code. scr: 80 doubles of SHARED memory,
// 16-byte aligned.
__device__ __forceinline__ bool warp_potrf16(double * D, int ld, int bs, double * dinv, double * scr, int lane) {
    const int i = lane & 15, h = lane >> 4;
    const bool live = i < bs;
    double a[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int k = 8 * h + c;
        a[c] = (live && k <= i) ? D[i * ld + k] : (k == i ? 1.0 : 0.0);
    }
    double d = __shfl_sync(0xffffffffu, a[0], 0), mydiag = 1.0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < kNB; j++) {
        const int hj = j >> 3, jj = j & 7;                          // static after unrolling
        double * colu = scr + 32 * (j & 1), * colt = colu + 16;     // double buffered: one __syncwarp per pivot
        ok = ok && d > 0.0;
        if (i == j) mydiag = d;
        const double r = __drcp_rn(d);
        const double t = a[jj] * r;                                 // u[i][j] / d[j] where h == hj
        if (jj < 7) {                                               // the next pivot lives in the same lane half
            const double dn = fma(-a[jj], t, a[jj + 1]);
            d = __shfl_sync(0xffffffffu, dn, j + 1 + 16 * hj);
        }
        if (h == hj) { colu[i] = a[jj]; colt[i] = t; }
        __syncwarp();
        const double u = colu[i];
        double tk[8];
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(colt + 8 * h + c);
            tk[c] = v.x; tk[c + 1] = v.y;
        }
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int k = 8 * h + c;
            if (k > j && i >= k) a[c] = fma(-u, tk[c], a[c]);
        }
        if (jj == 7 && j + 1 < kNB) d = __shfl_sync(0xffffffffu, a[0], j + 1 + 16);
    }
    // L = u diag(1 / sqrt(d)),  L[i][i] = sqrt(d[i]) (= d / sqrt(d), one Newton correction)
    const double isd = rsqrt(mydiag);
    double sd = mydiag * isd;
    sd = fma(fma(-sd, sd, mydiag), 0.5 * isd, sd);
    double * cis = scr + 64;
    if (h == 0) cis[i] = isd;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int k = 8 * h + c;
        if (live && k <= i) D[i * ld + k] = k == i ? sd : a[c] * cis[k];
    }
    if (live && h == 0) dinv[i] = isd;
    __syncwarp();
    return ok;
}

// lanes 0..15 = columns of the inverse of the lower-triangular diagonal block D (bs rows), written
// to X (same shape, zeros above the diagonal); dinv = reciprocals of D's diagonal
__device__ __forceinline__ void warp_trtri16(const double * D, int ld, int bs, const double * dinv, double * X, int ldx, int lane) {
    // column c of the inverse by forward substitution, written right-looking: as soon as x[k] is known it
    // is added into the sums of all later rows, so the dependent chain is two operations per row instead
    // of one per term (every sum still receives its terms in the order k = 0, 1, ...)
    double x[kNB], sum[kNB];
    const int c = lane;
#pragma unroll
    for (int i = 0; i < kNB; i++) sum[i] = 0.0;
#pragma unroll
    for (int k = 0; k < kNB; k++) {
        x[k] = k == c ? (k < bs ? dinv[k] : 0.0) : (k > c && k < bs ? -sum[k] * dinv[k] : 0.0);
#pragma unroll
        for (int i = 0; i < kNB; i++)
            if (i > k) sum[i] = fma(i < bs ? D[i * ld + k] : 0.0, x[k], sum[i]);
    }
    if (c < bs) {
#pragma unroll
        for (int i = 0; i < kNB; i++)
            if (i < bs) X[i * ldx + c] = x[i];
    }
}

__device__ bool block_spd_inverse(int n, double * A, int lda, double * W, int ldw, double * dinv /* n doubles */) {
    __shared__ int bad;
    __shared__ __align__(16) double potrf_scr[80];   // always shared memory, also in the global-workspace mode
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    PHASE_BEGIN();
    // ---- potrf, right-looking over 16-column panels
    for (int jb = 0; jb < n; jb += kNB) {
        const int bs = min(kNB, n - jb), below = n - jb - bs;
        double * D = A + jb * lda + jb;
        if (warp == 0) {
            const bool ok = warp_potrf16(D, lda, bs, dinv + jb, potrf_scr, lane);
            if (!ok && lane == 0) bad = 1;
        }
        __syncthreads();
        PHASE(16);
        if (bad) break;
        if (below > 0) {
            // panel rows: x L11^T = a, one thread per row (L11 reads are broadcasts)
            for (int i = threadIdx.x; i < below; i += blockDim.x) {
                double * row = A + (jb + bs + i) * lda + jb;
                double x[kNB];
#pragma unroll
                for (int j = 0; j < kNB; j++) x[j] = row[j];
#pragma unroll
                for (int k = 0; k < kNB; k++) {          // right-looking: two dependent operations per column
                    x[k] *= dinv[jb + k];
#pragma unroll
                    for (int j = 0; j < kNB; j++)
                        if (j > k) x[j] = fma(-x[k], D[j * lda + k], x[j]);
                }
#pragma unroll
                for (int j = 0; j < kNB; j++) row[j] = x[j];
            }
            __syncthreads();
            PHASE(17);
            // trailing block: A22 -= P P^T (lower tiles)
            const double * P = A + (jb + bs) * lda + jb;
            block_dgemm<false, true, true, 1>(below, below, kNB, P, lda, P, lda, A + (jb + bs) * lda + jb + bs, lda, -1.0);
            __syncthreads();
            PHASE(18);
        }
    }
    __syncthreads();
    if (bad) return false;
    // ---- trtri: W = L^-1. Diagonal blocks first (one warp each), zeros above the diagonal
    const int nblk = (n + kNB - 1) / kNB;
    for (int i = threadIdx.x; i < n * ldw; i += blockDim.x) W[i] = 0.0;
    __syncthreads();
    for (int b = warp; b < nblk; b += nwarps) {
        const int jb = b * kNB;
        if (lane < kNB) warp_trtri16(A + jb * lda + jb, lda, min(kNB, n - jb), dinv + jb, W + jb * ldw + jb, ldw, lane);
    }
    __syncthreads();
    PHASE(19);
    // block row I: W[I][0:I] = -W[I][I] * (L[I][0:I] * W[0:I][0:I])
    for (int ib = kNB; ib < n; ib += kNB) {
        const int bs = min(kNB, n - ib);
        block_dgemm<false, false, false, 0, 2>(bs, ib, ib, A + ib * lda, lda, W, ldw, W + ib * ldw, ldw);
        __syncthreads();
        PHASE(20);
        // in place: every warp reads the columns of its own tile completely before it stores them
        block_dgemm<false, false, false>(bs, ib, bs, W + ib * ldw + ib, ldw, W + ib * ldw, ldw, W + ib * ldw, ldw, -1.0);
        __syncthreads();
        PHASE(21);
    }
    // ---- lauum: A <- W^T W
    block_dgemm<true, false, false, 2, 1>(n, n, n, W, ldw, W, ldw, A, lda);
    __syncthreads();
    PHASE(22);
    return true;
}

// ---------------------------------------------------------------------------------------
// shared-memory plan (in doubles). Regions:  rg[2] | v1[2] | v2 | dinv | M | ARENA
// ---------------------------------------------------------------------------------------
struct DensePlan {
    int rg, v1, v2, dinv, M, arena, tab, total;
};
__host__ __device__ inline int even(int x) { return (x + 1) & ~1; }
__host__ __device__ inline DensePlan dense_plan(int op, int n, int cd, int ldj, int ldh, int tab_doubles) {
    DensePlan p;
    p.rg = 0;                                     // rg and v1: two buffers each (next geometry prefetched)
    p.v1 = p.rg + 2 * even(cd);
    p.v2 = p.v1 + 2 * even(cd > n ? cd : n);
    p.dinv = p.v2 + even(cd > n ? cd : n);
    p.M = p.dinv + even(n);
    const int szJ = n * ldj, szA = n * ldh, szH = cd * ldj, szT2 = cd * ldh;
    int arena = 0;
    switch (op) {
    case 0: case 1: p.arena = p.M + szJ; arena = szJ + szA; break;                         // M | J, A (W = M region)
    case 2: p.arena = p.M + szJ; arena = szJ + (szA > szH ? szA : szH); break;            // J | T, Hq/Hacc
    default: p.arena = p.M + szJ; arena = (szJ + szA > szH + szT2) ? szJ + szA : szH + szT2; break;
    }
    p.tab = p.arena + arena;                      // copy of the definition tables (when it fits)
    p.total = p.tab + even(tab_doubles);
    return p;
}

// OP 0: gradient_cart2int_matrix   out = M[n][cd]
// OP 1: gradient_cart2int          in1 = g_r[cd]                       out = g_q[n]
// OP 2: Hessian_int2cart           in1 = g_q[n],  in2 = H_q[n][n]      out = H_r[cd][cd]
// OP 3: Hessian_cart2int           in1 = g_r[cd], in2 = H_r[cd][cd]    out = H_q[n][n]
template <int OP, bool HAS5, bool WS>
__global__ void __launch_bounds__(kDenseThreads, 1)
dense_kernel(const DenseProgram dp_global, const double * __restrict__ r, const double * __restrict__ in1,
             const double * __restrict__ in2, const int64_t B, double * __restrict__ out, int * __restrict__ status,
             double * __restrict__ workspace) {
    // operands in shared memory, or - molecules beyond the shared-memory plan - in a per-block slice of
    // a global workspace (same code, served from L2; functional rather than fast)
    // (WS is a template parameter so that the shared-memory flavour keeps shared-state-space loads)
    extern __shared__ __align__(16) double sm_dyn[];
    constexpr bool in_smem = !WS;
    const int n = dp_global.intdim, cd = dp_global.cartdim, ldj = dp_global.ldj, ldh = dp_global.ldh;
    const DensePlan pl = dense_plan(OP, n, cd, ldj, ldh, dp_global.tab_doubles);
    double * sm = sm_dyn;
    if (WS) sm = workspace + (size_t)blockIdx.x * pl.total;
    // the definition tables (terms, row pointers, K tasks; a few KB) are read by every geometry: kept in
    // shared memory when they fit, so those reads never queue behind the operand copies in the L1 pipeline
    DenseProgram dp = dp_global;
    if (in_smem && dp_global.tab_doubles > 0) {
        double * tab = sm + pl.tab;
        const double * src = reinterpret_cast<const double *>(dp_global.terms);
        for (int i = threadIdx.x; i < dp_global.tab_doubles; i += blockDim.x) tab[i] = src[i];
        const char * base = reinterpret_cast<const char *>(dp_global.terms), * tb = reinterpret_cast<const char *>(tab);
        dp.terms = reinterpret_cast<const DenseTerm *>(tb);
        dp.row_ptr = reinterpret_cast<const int32_t *>(tb + (reinterpret_cast<const char *>(dp_global.row_ptr) - base));
        dp.tasks = reinterpret_cast<const DenseTask *>(tb + (reinterpret_cast<const char *>(dp_global.tasks) - base));
        if (dp_global.colmask) {
            dp.colmask = reinterpret_cast<const uint32_t *>(tb + (reinterpret_cast<const char *>(dp_global.colmask) - base));
            dp.rowmask = reinterpret_cast<const uint32_t *>(tb + (reinterpret_cast<const char *>(dp_global.rowmask) - base));
        }
    }                                                // (visible after the first barrier of the loop)
    double * v2 = sm + pl.v2, * Mreg = sm + pl.M, * arena = sm + pl.arena;
    // the small per-geometry inputs (coordinates; the gradient in1) are double buffered: the next
    // geometry's copies are issued (cp.async) as soon as this one's have been picked up, so their HBM
    // latency never shows. The global-workspace mode loads them synchronously into one buffer.
    const int rg_len = cd, v1_len = OP == 0 ? 0 : OP == 2 ? n : cd;
    const int rg_buf = even(cd), v1_buf = even(cd > n ? cd : n);
    auto fetch_small = [&](int64_t bb, int buf) {
        double * rgd = sm + pl.rg + buf * rg_buf, * v1d = sm + pl.v1 + buf * v1_buf;
        const double * rs = r + bb * rg_len, * vs = in1 + bb * v1_len;
        if (in_smem) {
            for (int i = threadIdx.x; i < rg_len; i += blockDim.x)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((uint32_t)__cvta_generic_to_shared(rgd + i)), "l"(rs + i) : "memory");
            for (int i = threadIdx.x; i < v1_len; i += blockDim.x)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((uint32_t)__cvta_generic_to_shared(v1d + i)), "l"(vs + i) : "memory");
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        } else {
            for (int i = threadIdx.x; i < rg_len; i += blockDim.x) rgd[i] = rs[i];
            for (int i = threadIdx.x; i < v1_len; i += blockDim.x) v1d[i] = vs[i];
        }
    };
    int cur = 0;
    if (in_smem && (int64_t)blockIdx.x < B) fetch_small(blockIdx.x, 0);

    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        if (!in_smem) fetch_small(b, 0);             // (every iteration ends with a barrier)
        double * rg = sm + pl.rg + cur * rg_buf, * v1 = sm + pl.v1 + cur * v1_buf;
        block_load_wait();
        __syncthreads();
        if (in_smem) {                               // the other buffers were last read before this barrier
            if (b + gridDim.x < B) fetch_small(b + gridDim.x, cur ^ 1);
            cur ^= 1;
        }

        if (OP == 2) {
            double * J = Mreg, * T = arena, * X = arena + n * ldj;       // X: H_q, later the H_r accumulator
            PHASE_BEGIN();
            // H_q travels while J is formed, issued by the warps that own no row of J (all warps, up front,
            // when every warp owns rows)
            const double * Hq_g = in2 + b * (int64_t)n * n;
            const bool idle_issue = (n + kLanes - 1) / kLanes < (int)(blockDim.x >> 5);
            if (!idle_issue) block_load_async(Hq_g, X, n, n, ldh, in_smem);
            block_compute_J<HAS5>(dp, rg, J, [&](int warp0) { if (idle_issue) block_load_async(Hq_g, X, n, n, ldh, in_smem, warp0); });   // syncs
            PHASE(0);
            block_load_wait();
            __syncthreads();
            PHASE(1);
            block_dgemm<false, false, false>(n, cd, n, X, ldh, J, ldj, T, ldj, 1.0, KMask{dp.colmask, 1});      // T = H_q J
            __syncthreads();
            PHASE(2);
            block_dgemm<true, false, false>(cd, cd, n, J, ldj, T, ldj, X, ldj, 1.0, KMask{dp.colmask, 2});      // X = J^T T (H_q is dead)
            __syncthreads();
            PHASE(3);
            block_add_gK<HAS5>(dp, rg, v1, 1.0, X, ldj);                  // X += sum_k g_q[k] K[k]
            PHASE(4);
            block_store(X, out + b * (int64_t)cd * cd, cd, cd, ldj);
            __syncthreads();
            PHASE(6);
            continue;
        }

        // OP 0, 1, 3 start with M = (J J^T)^-1 J
        double * J = arena, * A = arena + n * ldj, * W = Mreg;           // W (n x n, ldh) borrows the M region
        PHASE_BEGIN();
        block_compute_J<HAS5>(dp, rg, J);
        PHASE(8);
        block_dgemm<false, true, false, 1>(n, n, cd, J, ldj, J, ldj, A, ldh, 1.0, KMask{dp.rowmask, 3});        // A = J J^T (lower)
        __syncthreads();
        PHASE(9);
        const bool ok = block_spd_inverse(n, A, ldh, W, ldh, sm + pl.dinv);                        // A = (J J^T)^-1
        PHASE(10);
        if (!ok) {
            // the reference's at::cholesky throws here; the batched call flags the table (tchem_ic_factor_status) and
            // returns NaN for this geometry instead of numbers computed on a half-factored matrix
            if (threadIdx.x == 0) atomicExch(status, 1);
            for (int i = threadIdx.x; i < n * ldh; i += blockDim.x) A[i] = __longlong_as_double(0x7ff8000000000000ll);
            __syncthreads();
        }
        if (OP == 1) {
            // g_q = inv (J g_r)   (= (inv J) g_r, source/IC/IntCoordSet.cpp:202-203, without forming inv J)
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                double s = 0.0;
                for (int k = 0; k < cd; k++) s = fma(J[i * ldj + k], v1[k], s);
                v2[i] = s;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                double s = 0.0;
                for (int k = 0; k < n; k++) s = fma(A[i * ldh + k], v2[k], s);
                out[b * n + i] = s;
            }
            __syncthreads();
            continue;
        }
        block_dgemm<false, false, false>(n, cd, n, A, ldh, J, ldj, Mreg, ldj, 1.0, KMask{dp.colmask, 1});       // M = inv J
        __syncthreads();
        PHASE(11);
        if (OP == 0) {
            block_store(Mreg, out + b * (int64_t)n * cd, n, cd, ldj);
            __syncthreads();
            continue;
        }
        // OP 3: H_q = M (H_r - sum_k g_q[k] K[k]) M^T,  g_q = M g_r
        double * Hc = arena, * T2 = arena + cd * ldj;                    // J and A are dead now
        block_load_async(in2 + b * (int64_t)cd * cd, Hc, cd, cd, ldj, in_smem);       // in flight while g_q = M g_r is formed
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double s = 0.0;
            for (int k = 0; k < cd; k++) s = fma(Mreg[i * ldj + k], v1[k], s);
            v2[i] = s;
        }
        block_load_wait();
        __syncthreads();
        PHASE(12);
        block_add_gK<HAS5>(dp, rg, v2, -1.0, Hc, ldj);                   // Hc = H_r - sum g_q K
        PHASE(13);
        block_dgemm<false, true, false>(cd, n, cd, Hc, ldj, Mreg, ldj, T2, ldh);     // T2 = Hc M^T
        __syncthreads();
        block_dgemm<false, false, false>(n, n, cd, Mreg, ldj, T2, ldh, Hc, ldh);     // H_q = M T2 (into Hc, ld = ldh)
        __syncthreads();
        PHASE(14);
        block_store(Hc, out + b * (int64_t)n * n, n, n, ldh);
        __syncthreads();
        PHASE(15);
    }
}

struct DenseTable {
    DenseProgram prog;
    void * blob = nullptr;
    int * status = nullptr;
};
std::map<const tchem_ic_table *, DenseTable> & dense_tables() {
    static std::map<const tchem_ic_table *, DenseTable> m;
    return m;
}
std::mutex & dense_mutex() { static std::mutex m; return m; }

int pick_ld(int width) {      // smallest ld >= width with ld = 4 or 12 (mod 16)
    int ld = width;
    while ((ld & 15) != 4 && (ld & 15) != 12) ld++;
    return ld;
}

int get_dense(const tchem_ic_table * t, DenseTable & out) {
    std::lock_guard<std::mutex> lock(dense_mutex());
    auto it = dense_tables().find(t);
    if (it != dense_tables().end()) { out = it->second; return TCHEM_OK; }
    std::vector<std::vector<int>> rows(t->intdim);
    for (size_t k = 0; k < t->terms.size(); k++) rows[t->terms[k].ic].push_back((int)k);
    std::vector<DenseTerm> terms;
    std::vector<int32_t> row_ptr(1, 0);
    std::vector<DenseTask> tasks;
    for (int i = 0; i < t->intdim; i++) {
        for (int k : rows[i]) {
            const auto & d = t->terms[k];
            DenseTerm dt;
            std::memset(&dt, 0, sizeof(dt));
            dt.pd.type = d.type; dt.pd.natoms = d.natoms;
            for (int a = 0; a < 5; a++) dt.pd.atoms[a] = a < d.natoms ? d.atoms[a] : 0;
            dt.pd.out = i;
            dt.pd.min = d.min;
            dt.coef = d.coeff;
            for (int dir = 0; dir < 3 * d.natoms; dir++) tasks.push_back(DenseTask{(int32_t)terms.size(), dir});
            terms.push_back(dt);
        }
        row_ptr.push_back((int32_t)terms.size());
    }
    // heavier tasks first (torsions before bends before stretches) so the tail of the task loop is short;
    // within one type direction-major: the 32 lanes of a warp then hold the SAME direction of 32 different
    // terms - uniform control flow in the scatter (which blocks a direction touches depends on it), and
    // lanes that add into the shared accumulator in the same instruction hit different elements (same
    // local block of different terms), so the CAS loops of the double-precision adds rarely retry
    std::stable_sort(tasks.begin(), tasks.end(), [&](const DenseTask & a, const DenseTask & b) {
        const auto & ta = terms[a.term].pd; const auto & tb = terms[b.term].pd;
        if (ta.natoms != tb.natoms) return ta.natoms > tb.natoms;
        if (ta.type != tb.type) return ta.type < tb.type;
        return a.dir < b.dir;
    });
    auto a16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    // structural nonzeros of J at k-step granularity (see DenseProgram)
    const bool masked = t->intdim <= 128 && t->cartdim <= 128;
    std::vector<uint32_t> colmask((t->cartdim + 7) / 8 + 4, 0), rowmask((t->intdim + 7) / 8 + 4, 0);
    if (masked) for (const DenseTerm & dt : terms)
        for (int a = 0; a < dt.pd.natoms; a++)
            for (int c = 0; c < 3; c++) {
                const int row = dt.pd.out, col = 3 * dt.pd.atoms[a] + c;
                colmask[col / 8] |= 1u << (row / 4);
                rowmask[row / 8] |= 1u << (col / 4);
            }
    const size_t o_t = 0, o_r = a16(terms.size() * sizeof(DenseTerm)), o_k = o_r + a16(row_ptr.size() * 4);
    const size_t o_cm = o_k + a16(tasks.size() * sizeof(DenseTask)), o_rm = o_cm + a16(colmask.size() * 4);
    const size_t total = o_rm + a16(rowmask.size() * 4) + 16;
    std::vector<unsigned char> blob(total, 0);
    std::memcpy(blob.data() + o_t, terms.data(), terms.size() * sizeof(DenseTerm));
    std::memcpy(blob.data() + o_r, row_ptr.data(), row_ptr.size() * 4);
    std::memcpy(blob.data() + o_k, tasks.data(), tasks.size() * sizeof(DenseTask));
    std::memcpy(blob.data() + o_cm, colmask.data(), colmask.size() * 4);
    std::memcpy(blob.data() + o_rm, rowmask.data(), rowmask.size() * 4);
    void * dev = nullptr;
    TC_CUDA(cudaMalloc(&dev, total));
    cudaError_t e = cudaMemcpy(dev, blob.data(), total, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dev); return set_error(TCHEM_ECUDA, cudaGetErrorString(e)); }
    DenseTable dt;
    dt.blob = dev;
    TC_CUDA(cudaMalloc(reinterpret_cast<void **>(&dt.status), sizeof(int)));
    TC_CUDA(cudaMemset(dt.status, 0, sizeof(int)));
    unsigned char * d = static_cast<unsigned char *>(dev);
    dt.prog.terms = reinterpret_cast<const DenseTerm *>(d + o_t);
    dt.prog.row_ptr = reinterpret_cast<const int32_t *>(d + o_r);
    dt.prog.tasks = reinterpret_cast<const DenseTask *>(d + o_k);
    dt.prog.colmask = masked ? reinterpret_cast<const uint32_t *>(d + o_cm) : nullptr;
    dt.prog.rowmask = masked ? reinterpret_cast<const uint32_t *>(d + o_rm) : nullptr;
    dt.prog.nterms = (int)terms.size();
    dt.prog.ntasks = (int)tasks.size();
    dt.prog.intdim = t->intdim;
    dt.prog.cartdim = t->cartdim;
    dt.prog.ldj = pick_ld(t->cartdim);
    dt.prog.ldh = pick_ld(t->intdim);
    dt.prog.blob_doubles = (int32_t)(total / 8);
    dt.prog.tab_doubles = 0;
    dense_tables()[t] = dt;
    out = dt;
    return TCHEM_OK;
}

template <int OP> const void * dense_fn(bool has5, bool ws) {
    return ws ? (has5 ? reinterpret_cast<const void *>(&dense_kernel<OP, true, true>) : reinterpret_cast<const void *>(&dense_kernel<OP, false, true>))
              : (has5 ? reinterpret_cast<const void *>(&dense_kernel<OP, true, false>) : reinterpret_cast<const void *>(&dense_kernel<OP, false, false>));
}

int launch_dense(int op, const char * who, const tchem_ic_table * t, const double * r, const double * in1,
                 const double * in2, int64_t B, double * out, cudaStream_t stream) {
    const std::string name = std::string("tchem::IC::IntCoordSet::") + who;
    if (!t) return set_error(TCHEM_EINVAL, name + ": null table");
    if (B < 0) return set_error(TCHEM_EINVAL, name + ": negative batch");
    if (B == 0) return TCHEM_OK;
    if (!r || !out || (op >= 1 && !in1) || (op >= 2 && !in2)) return set_error(TCHEM_EINVAL, name + ": null argument");
    if (op != 2 && t->intdim > t->cartdim)
        return set_error(TCHEM_EINVAL, name + ": more internal coordinates than Cartesian coordinates (J J^T is singular)");
    DeviceGuard guard(t->device);
    if (!guard.ok) return set_error(TCHEM_ECUDA, "cudaSetDevice failed");
    {
        // definition tables whose J is sparse (the usual case: <= 15 nonzeros per term) take the compressed-J path
        const int rs = launch_sparse(op, t, r, in1, in2, B, out, stream);
        if (rs != -1) return rs;
    }
    DenseTable dt;
    int rc = get_dense(t, dt);
    if (rc != TCHEM_OK) return rc;
    // with a shared-memory copy of the definition tables when that still fits, without it otherwise
    DensePlan pl = dense_plan(op, t->intdim, t->cartdim, dt.prog.ldj, dt.prog.ldh, dt.prog.blob_doubles);
    dt.prog.tab_doubles = dt.prog.blob_doubles;
    if ((size_t)pl.total * sizeof(double) + 4096 > (size_t)t->max_smem_optin) {
        pl = dense_plan(op, t->intdim, t->cartdim, dt.prog.ldj, dt.prog.ldh, 0);
        dt.prog.tab_doubles = 0;
    }
    size_t smem = (size_t)pl.total * sizeof(double);
    // Molecules whose operands do not fit the shared memory of one SM (more than ~32 atoms for a full
    // 3N-6 set) run the same kernel on a per-block slice of a stream-ordered global workspace: correct
    // for any size, served from L2 instead of shared memory (TCHEM_DENSE_WORKSPACE=1 forces this mode)
    static const bool force_ws = [] { const char * e = getenv("TCHEM_DENSE_WORKSPACE"); return e && atoi(e) != 0; }();
    const bool use_ws = force_ws || smem + 4096 > (size_t)t->max_smem_optin;
    const void * fn = op == 0 ? dense_fn<0>(t->has5, use_ws) : op == 1 ? dense_fn<1>(t->has5, use_ws)
                    : op == 2 ? dense_fn<2>(t->has5, use_ws) : dense_fn<3>(t->has5, use_ws);
    double * workspace = nullptr;
    int64_t grid;
    if (use_ws) {
        smem = 0;
        int per_sm = 0;
        TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kDenseThreads, 0));
        grid = std::min<int64_t>(B, (int64_t)t->sm_count * std::max(per_sm, 1));
        const size_t bytes = (size_t)grid * pl.total * sizeof(double);
        keep_async_pool(t->device);
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&workspace), bytes, stream);
        if (e != cudaSuccess) { cudaGetLastError(); return set_error(TCHEM_ECUDA, name + ": cannot allocate the workspace for a molecule of this size"); }
    } else {
        { const int rc_ = ensure_max_dynamic_smem(fn, t->device, t->max_smem_optin); if (rc_ != TCHEM_OK) return rc_; }
        int per_sm = 0;
        TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kDenseThreads, smem));
        if (per_sm < 1) per_sm = 1;
        grid = std::min<int64_t>(B, (int64_t)t->sm_count * per_sm);
    }
    void * args[] = {(void *)&dt.prog, (void *)&r, (void *)&in1, (void *)&in2, (void *)&B, (void *)&out, (void *)&dt.status, (void *)&workspace};
    cudaError_t le = cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(kDenseThreads), args, smem, stream);
    if (workspace) cudaFreeAsync(workspace, stream);
    if (le != cudaSuccess) return set_error(TCHEM_ECUDA, cudaGetErrorString(le));
    count_launch(op == 0 ? "dense_kernel<0:gradient_cart2int_matrix>" : op == 1 ? "dense_kernel<1:gradient_cart2int>" : op == 2 ? "dense_kernel<2:Hessian_int2cart>" : "dense_kernel<3:Hessian_cart2int>");
    return TCHEM_OK;
}

} // namespace

void release_dense_program(const tchem_ic_table * t) {
    std::lock_guard<std::mutex> lock(dense_mutex());
    auto it = dense_tables().find(t);
    if (it != dense_tables().end()) { cudaFree(it->second.blob); cudaFree(it->second.status); dense_tables().erase(it); }
}

// 1 when a previous cart2int launch on this table met a non-positive Cholesky pivot (redundant or
// singular internal coordinates; the reference's at::cholesky throws there). Synchronises.
int dense_status(const tchem_ic_table * t) {
    std::lock_guard<std::mutex> lock(dense_mutex());
    auto it = dense_tables().find(t);
    if (it == dense_tables().end()) return 0;
    int s = 0;
    cudaMemcpy(&s, it->second.status, sizeof(int), cudaMemcpyDeviceToHost);
    if (s) cudaMemset(it->second.status, 0, sizeof(int));
    return s;
}

} // namespace tchem_b200

using namespace tchem_b200;

extern "C" {

int tchem_ic_grad_cart2int_matrix(const tchem_ic_table * t, const double * r, int64_t B, double * M, tchem_stream_t stream) {
    return launch_dense(0, "gradient_cart2int", t, r, nullptr, nullptr, B, M, static_cast<cudaStream_t>(stream));
}
int tchem_ic_grad_cart2int(const tchem_ic_table * t, const double * r, const double * cartgrad, int64_t B,
                           double * intgrad, tchem_stream_t stream) {
    return launch_dense(1, "gradient_cart2int", t, r, cartgrad, nullptr, B, intgrad, static_cast<cudaStream_t>(stream));
}
int tchem_ic_hess_int2cart(const tchem_ic_table * t, const double * r, const double * intgrad, const double * intHess,
                           int64_t B, double * cartHess, tchem_stream_t stream) {
    return launch_dense(2, "Hessian_int2cart", t, r, intgrad, intHess, B, cartHess, static_cast<cudaStream_t>(stream));
}
int tchem_ic_hess_cart2int(const tchem_ic_table * t, const double * r, const double * cartgrad, const double * cartHess,
                           int64_t B, double * intHess, tchem_stream_t stream) {
    return launch_dense(3, "Hessian_cart2int", t, r, cartgrad, cartHess, B, intHess, static_cast<cudaStream_t>(stream));
}
#ifdef TCHEM_DENSE_PHASES
__attribute__((visibility("default"))) int tchem_debug_dense_phases(unsigned long long * out32, int reset) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out32, g_dense_phase, sizeof(g_dense_phase)) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_dense_phase, z, sizeof(z)); }
    return 0;
}
#endif
// 0 = every Cholesky factorisation since the last call succeeded
int tchem_ic_factor_status(const tchem_ic_table * t) {
    if (!t) return 0;
    DeviceGuard guard(t->device);
    const int a = dense_status(t), b = sparse_status(t);
    return (a || b) ? 1 : 0;
}

} // extern "C"
