// The gradient / Hessian transforms with J kept SPARSE (round 2; ic_dense.cu keeps the all-DMMA flavour for
// definition tables whose J is not sparse).
//
// Replaces, batched (reference source/IC/IntCoordSet.cpp):
//   gradient_cart2int_matrix :178-188   M   = (J J^T)^-1 J
//   gradient_cart2int        :190-205   g_q = (J J^T)^-1 (J g_r)
//   Hessian_int2cart         :248-266   H_r = J^T H_q J + sum_k g_q[k] K[k]
//   Hessian_cart2int         :221-246   H_q = A^T H_r A - sum_k (A^T g_r)[k] A^T K[k] A,  A^T = (J J^T)^-1 J
//
// A row of J has at most 15 structural nonzeros per term (3 per atom of the invariant displacement), known from
// the definition table alone. Every product with J therefore runs over a compressed row / column list with plain
// FP64 FMAs (C4: 750 nonzeros instead of 7 560 elements: ~10x fewer flops than the dense-masked DMMA products),
// and only the two genuinely dense intdim^3 products of Hessian_cart2int stay on the tensor cores.
//
// Two kernels:
//  * factor_kernel - ONE WARP PER GEOMETRY, six resident per SM, no block barrier anywhere. J J^T is assembled
//    from the compressed rows into ROW-MAJOR PACKED lower storage (28 KB for 84 coordinates), factored in place
//    (right-looking Cholesky, panels of 8 columns held in registers, the diagonal block replicated in every lane), with the
//    right-hand side J g_r riding along as a bordered row (= the forward substitution), then back substitution.
//    For the matrix / Hessian entry points the same warp continues with the reference's potri (explicit inverse of
//    the factor, row-oriented; then W^T W), still in place, and writes the inverse - expanded to the full symmetric
//    matrix - and g_q to a workspace.
//    The serial pivot chain of one geometry overlaps with the other warps' geometries instead of idling a block;
//    geometries are handed out through a per-launch counter (the warps do not run at the same speed).
//  * hess_kernel - one block per geometry: compressed-column products U = J^T H_q, X = U J (int -> cart), or
//    M^T = J^T (J J^T)^-1 by the same sparse product followed by the two dense DMMA products T2 = H_c M^T,
//    H_q = M T2 (cart -> int; the reference's order of operations - see the kernel). The second-derivative term
//    sum_k g_k K[k] is evaluated as (term, direction) Dual passes that park their 3 x 3 blocks in slots, and a fixed-
//    order gather adds the slots into the accumulator: no atomics, bit-reproducible. The next geometry's matrices
//    are in flight (TMA bulk copies completing on an mbarrier; cp.async where alignment forbids) while the current
//    one is being worked on, and its J rows are evaluated by the warps left over by the (term, direction) tasks.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include "dmma_gemm.cuh"
#include "ic_eval.cuh"
#include "ic_table.h"

namespace tchem_b200 {

struct SpTerm {                       // 72 bytes = 18 words: the lanes of a warp step read the same field of CONSECUTIVE terms, and a
    PrimDesc pd;                      // 16-word stride would put all 32 of them on two banks (measured: a third of the block kernel's
    double coef;                      // shared-memory wavefronts were bank conflicts); pd.out = internal coordinate (row) of the term
    int16_t pos[5];                   // Jv position of the 3 entries of atom slot a (row-compressed J)
    int16_t pad_;
    int32_t kslot;                    // first slot of the term's block-upper second-derivative storage
    int32_t pad2_[2];
};
static_assert(sizeof(SpTerm) == 72, "SpTerm layout");

struct SparseProgram {                // device pointers into one blob (starts at `terms`)
    const SpTerm *   terms;           // row order
    const int32_t *  row_ptr;         // [intdim + 1] terms of a row
    const int32_t *  jrow_order;      // [njrow] rows in evaluation order: grouped by type, groups padded to 32 with -1
    const uint32_t * ktasks;          // [32 * nsteps] warp steps of 32 same-kind tasks: (term << 8) | direction, 0x80000000 | row (J row of the
                                      // next geometry), 0xffffffff = none
    const uint16_t * sched_ptr;       // [hess warps + 1] static schedule: the steps of each warp (longest-processing-time assignment)
    const uint16_t * sched_step;      // [nsteps]
    const uint16_t * nz_ptr;          // [intdim + 1] CSR over Jv
    const uint16_t * nz_col;          // [nnz] column of every Jv position
    const uint16_t * col_ptr;         // [natoms + 1] compressed COLUMN BLOCKS: the three columns of an atom share their row list
    const uint32_t * col_ent;         // [nnz / 3] (row << 16) | Jv position of the x component (y, z follow)
    const uint16_t * aw_ptr;          // [hess warps + 1] the atoms (column blocks) each warp of the block kernel owns in the products with J
    const uint16_t * aw_atom;         // [natoms] longest-processing-time assignment by the length of the row lists
    const uint32_t * pair_dest;       // [npairs_pad] packed index (i (i + 1) / 2 + j) of a structurally nonzero A[i][j], j <= i
    const uint32_t * pair_ent;        // [maxc][npairs_pad] (Jv position in row i << 16) | Jv position in row j of a common atom
    const uint32_t * gk_blk;          // [ngblk] (atom A << 16) | atom B: 3 x 3 block of the g.K accumulator that receives slots
    const uint32_t * gk_ptr;          // [ngblk + 1]
    const uint32_t * gk_src;          // [ngsrc] (slot index << 2) | (negated << 1) | transposed
    int32_t nterms, njrow, nktasks, nnz, nnz_pad, npairs_pad, maxc, ngblk, ngsrc, nslots;
    int32_t natoms;
    int32_t intdim, cartdim, ldh;     // ldh: leading dimension of the intdim x intdim DMMA operands (4 or 12 mod 16)
    int32_t jv_alias, npairs_lo;      // jv_alias: the factor kernel keeps Jv in the tail of the packed matrix's storage (see factor_kernel);
                                      // npairs_lo: pairs whose destination lies below that tail
    int32_t factor_tab_bytes;         // leading part of the blob the factor kernel copies to shared memory: pair lists | common tables
    int32_t hess_tab_offset, blob_bytes;   // the block kernel copies [hess_tab_offset, blob_bytes): common tables | its own
    int32_t tab_in_smem;              // per launch: the block kernel works on a shared-memory copy of its part of the blob
};

namespace {

constexpr int kNB = 8;                // panel width of the warp-level factorisation
constexpr int kMaxSlots = 4;          // lanes x 4 = 128: upper limit of intdim + 1 and cartdim on this path
constexpr int kHessThreads = 384;     // 12 warps: the (term, direction) tasks + next J rows of C4 (834) run as torsion, bend and stretch rounds

// -DTCHEM_SPARSE_PHASES: thread 0 of block 0 accumulates the clock cycles of each phase of its geometries
// (tchem_debug_sparse_phases; tools/sparse_phases.py). Off in the product.
#ifdef TCHEM_SPARSE_PHASES
__device__ unsigned long long g_sparse_phase[32];
__device__ unsigned long long g_sparse_step[96];       // [0, 64): cycles of each warp step of the task phase; [64, 80): cycles of each warp's whole list
__device__ __forceinline__ long long sphase_clock(const double * probe) {
    const double x = *reinterpret_cast<const volatile double *>(probe);      // the read waits for the barrier's release
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "d"(x) : "memory");
    return t;
}
#define SPHASE_BEGIN(probe) const double * sph_probe = (probe); long long sph_t0 = sphase_clock(sph_probe)
#define SPHASE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long sph_t = sphase_clock(sph_probe); g_sparse_phase[i] += (unsigned long long)(sph_t - sph_t0); sph_t0 = sph_t; } } while (0)
#else
#define SPHASE_BEGIN(probe) do {} while (0)
#define SPHASE(i) do {} while (0)
#endif

__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }
__device__ __forceinline__ double nan64() { return __longlong_as_double(0x7ff8000000000000ll); }

// ---- sinks -------------------------------------------------------------------------------------------
struct JvSink {                        // Jv[pos[atom slot] + component] += coef * J_term   (row owned by one thread)
    double * Jv;
    const int16_t * pos;
    double coef;
    __device__ __forceinline__ void value(double) const {}
    __device__ __forceinline__ void jac_block(const int *, int i, const V3<double> & J) const {
        double * p = Jv + pos[i];
        p[0] = fma(coef, J.x, p[0]);
        p[1] = fma(coef, J.y, p[1]);
        p[2] = fma(coef, J.z, p[2]);
    }
    __device__ __forceinline__ void jac_entry(const int *, int m, double j) const {
        const int a = m / 3;
        double * p = Jv + pos[a] + (m - 3 * a);
        p[0] = fma(coef, j, p[0]);
    }
    template <int N> __device__ __forceinline__ void hess_block(const int *, int, int, const V3<Dual> &) const {}
    __device__ __forceinline__ void hess_entry5(const int *, int, int, double) const {}
};

struct SlotSink {                      // S[9 kblock(i, j) + 3 k + l] = w * d2 q / d r_(i,k) d r_(j,l),  i <= j
    double * S;
    double w;
    bool ends_free;                    // plain torsion: only the passes of atoms 1 and 2 run (see get_sparse), so block (j, 3) comes
                                       // from the pass of atom j as the transpose of what this pass emits for atom 3
    __device__ __forceinline__ void value(double) const {}
    __device__ __forceinline__ void jac_block(const int *, int, const V3<double> &) const {}
    __device__ __forceinline__ void jac_entry(const int *, int, double) const {}
    template <int N> __device__ __forceinline__ void hess_block(const int *, int dir, int i, const V3<Dual> & J) const {
        const int j = dir / 3, l = dir - 3 * j;
        if (i > j) {
            if (N == 4 && ends_free && i == 3) {
                double * p = S + 9 * kblock(j, i, N) + 3 * l;
                p[0] = w * J.x.d;
                p[1] = w * J.y.d;
                p[2] = w * J.z.d;
            }
            return;
        }
        double * p = S + 9 * kblock(i, j, N) + l;
        p[0] = w * J.x.d;
        p[3] = w * J.y.d;
        p[6] = w * J.z.d;
    }
    __device__ __forceinline__ void hess_entry5(const int *, int m1, int m2, double k) const {
        const int i = m1 / 3, j = m2 / 3, kk = m1 - 3 * i, ll = m2 - 3 * j;
        double * p = S + 9 * kblock(i, j, 5);
        p[3 * kk + ll] = w * k;
        if (i == j) p[3 * ll + kk] = w * k;
    }
};

// J rows of one geometry into the compressed value array Jv (zeroed here); `tid` of `nthreads` take the rows in
// the table's evaluation order (same-type rows share a warp)
template <bool HAS5>
__device__ __forceinline__ void eval_J_rows(const SparseProgram & sp, const double * rg, double * Jv, int tid, int nthreads) {
    for (int k = tid; k < sp.njrow; k += nthreads) {
        const int row = sp.jrow_order[k];
        if (row < 0) continue;
        for (int p = sp.nz_ptr[row]; p < sp.nz_ptr[row + 1]; p++) Jv[p] = 0.0;
        for (int t = sp.row_ptr[row]; t < sp.row_ptr[row + 1]; t++) {
            const SpTerm & term = sp.terms[t];
            JvSink sink;
            sink.Jv = Jv; sink.pos = term.pos; sink.coef = term.coef;
            eval_primitive<MODE_QJ, HAS5>(term.pd, GeomLoader{rg}, sink);
        }
    }
}

// ======================================================================================================
// factor_kernel: one warp = one geometry
// ======================================================================================================
// The O(n^3) parts of the factorisation run on the FP64 tensor cores, one 8 x 8 tile per DMMA pair
// (mma.m8n8k4.f64; fragments: a = A[gid][tig], b = B[tig][gid], c0 / c1 = C[gid][2 tig + {0, 1}], gid = lane >> 2,
// tig = lane & 3), with the fragments addressed straight in the row-major PACKED lower storage A[i (i + 1) / 2 + j]:
// a plain-FMA version of the same loops spent 80 % of its instructions on addressing and predicates (53 000 warp
// instructions per 84 x 84 factorisation at 0.27 IPC; tools/proto/warp_chol.py holds the lane-level emulation both
// versions were checked with).

// In-place lower Cholesky factor of the packed matrix A (rows 0 .. n-1; with `border` row n holds a right-hand
// side b and ends as L^-1 b). Right-looking over panels of 8 columns:
//  * panel: lane = row (rows j0 + lane + 32 m), the row's 8 panel entries in registers; pivot c belongs to lane c
//    (slot 0); its reciprocal square root is taken by every lane, the scaled pivot column's diagonal-block entries
//    travel by shuffle;
//  * trailing update A[i][j] -= sum_c L[i][c] L[j][c] by 8 x 8 tiles (I >= J), two DMMAs each.
// rdiag[j] = 1 / L[j][j]. Returns false when a pivot is not positive (NaN included).
// panel of 8 columns starting at j0, NS = row slots in use (rows j0 + lane + 32 m): registers only
template <int NS>
__device__ __forceinline__ bool warp_potrf_panel(double * A, int n, int last, int j0, int bs, double * rdiag, int lane) {
    double a[NS][kNB];
    bool ok = true;
#pragma unroll
    for (int m = 0; m < NS; m++) {
        const int i = j0 + lane + 32 * m;
        const double * Ai = A + tri(min(i, last)) + j0;
        const int cmax = i > last ? -1 : i == n ? bs - 1 : min(i - j0, bs - 1);       // last stored panel column of the row
#pragma unroll
        for (int c = 0; c < kNB; c++) a[m][c] = c <= cmax ? Ai[c] : 0.0;
    }
    // The 8 x 8 diagonal block is ALSO held, whole, by every lane (broadcast loads) and factored redundantly alongside:
    // pivots and multipliers then come from the lane's own registers, and the chain per column is rsqrt -> scale -> fma
    // (about 115 cycles) instead of shuffle -> rsqrt -> scale -> shuffle -> fma (about 175). Same operations in the same
    // order as the lanes that own those rows perform on a[0][.], so both copies hold the same bits.
    double dd[kNB][kNB];
#pragma unroll
    for (int r = 0; r < kNB; r++) {
        const int i = min(j0 + r, last);
        const double * Ad = A + tri(i) + j0;
#pragma unroll
        for (int c = 0; c < kNB; c++) dd[r][c] = (c <= r && r < bs) ? Ad[c] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < kNB; c++) {
        if (c < bs) {
            const double d = dd[c][c];
            ok = ok && d > 0.0;
            double rs = rsqrt(d);
            rs = fma(0.5 * rs, fma(-d * rs, rs, 1.0), rs);              // one Newton step: correctly rounded in all but rare cases
            if (lane == 0) rdiag[j0 + c] = rs;
#pragma unroll
            for (int r = c; r < kNB; r++) dd[r][c] *= rs;
#pragma unroll
            for (int m = 0; m < NS; m++) a[m][c] *= rs;                 // the pivot's own lane: d / sqrt(d) = sqrt(d)
#pragma unroll
            for (int c2 = c + 1; c2 < kNB; c2++) {
                const double t = dd[c2][c];                             // L[j0 + c2][j0 + c]  (columns >= bs: zeros, harmless)
#pragma unroll
                for (int r = c2; r < kNB; r++) dd[r][c2] = fma(-dd[r][c], t, dd[r][c2]);
#pragma unroll
                for (int m = 0; m < NS; m++) a[m][c2] = fma(-a[m][c], t, a[m][c2]);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < NS; m++) {
        const int i = j0 + lane + 32 * m;
        double * Ai = A + tri(min(i, last)) + j0;
        const int cmax = i > last ? -1 : i == n ? bs - 1 : min(i - j0, bs - 1);
#pragma unroll
        for (int c = 0; c < kNB; c++)
            if (c <= cmax) Ai[c] = a[m][c];
    }
    return ok;
}

__device__ __forceinline__ bool warp_potrf_packed(double * A, int n, bool border, double * rdiag, int lane) {
    const int last = border ? n : n - 1;
    const int gid = lane >> 2, tig = lane & 3;
    bool ok = true;
    for (int j0 = 0; j0 < n; j0 += kNB) {
#ifdef TCHEM_SPARSE_PHASES
        const long long sub_t0 = clock64();
#endif
        const int bs = min(kNB, n - j0);
        const int nrows = last - j0 + 1;
        bool pok;
        if (nrows <= 32)      pok = warp_potrf_panel<1>(A, n, last, j0, bs, rdiag, lane);
        else if (nrows <= 64) pok = warp_potrf_panel<2>(A, n, last, j0, bs, rdiag, lane);
        else if (nrows <= 96) pok = warp_potrf_panel<3>(A, n, last, j0, bs, rdiag, lane);
        else                  pok = warp_potrf_panel<4>(A, n, last, j0, bs, rdiag, lane);
        ok = ok && pok;
        __syncwarp();
#ifdef TCHEM_SPARSE_PHASES
        long long sub_t1 = 0;
        if (blockIdx.x == 0 && threadIdx.x == 0) { sub_t1 = clock64(); g_sparse_phase[7] += (unsigned long long)(sub_t1 - sub_t0); }
#endif
        const int t0 = j0 + bs;
        if (t0 < n) {                                   // (bs == 8 here: only the last panel can be narrower)
            // two row tiles (I, I + 8) x two column tiles per pass: the B fragments are loaded once for both row tiles (-6 % on the
            // phase; applying the updates one pass later, after the next pass's DMMAs, was measured SLOWER: on the schedulers that
            // host two of the six warps the phase is bound by the DMMA pipe itself, 16 cycles per instruction, not by the chain)
            for (int I = t0; I <= last; I += 16) {
                const int ia = I + gid, ib = I + 8 + gid, ica = min(ia, last), icb = min(ib, last);
                const double * La = A + tri(ica) + j0 + tig, * Lb = A + tri(icb) + j0 + tig;
                const double a0 = La[0], a1 = La[4], f0 = Lb[0], f1 = Lb[4];       // A fragments: L[i][j0 + tig (+ 4)]
                double * Ca = A + tri(ica) + 2 * tig, * Cb = A + tri(icb) + 2 * tig;
                const int jmax_a = ia <= last ? min(ia, n - 1) : -1, jmax_b = ib <= last ? min(ib, n - 1) : -1;
                const int jend = min(I + 16, n);
                // B fragments: L[J + gid][j0 + tig (+ 4)]; the row pointer advances by tri(j + 8) - tri(j) = 8 j + 36
                int jb = t0 + gid;
                const double * Lj = A + tri(min(jb, n - 1)) + j0 + tig;
                for (int J = t0; J < jend; J += 16) {
                    const double b00 = Lj[0], b01 = Lj[4];
                    const double * Lj2 = Lj + (jb + 8 < n ? 8 * jb + 36 : 0);
                    const double b10 = Lj2[0], b11 = Lj2[4];
                    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0, g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
                    dmma(c0, c1, a0, b00);
                    dmma(e0, e1, a0, b10);
                    dmma(g0, g1, f0, b00);
                    dmma(h0, h1, f0, b10);
                    dmma(c0, c1, a1, b01);
                    dmma(e0, e1, a1, b11);
                    dmma(g0, g1, f1, b01);
                    dmma(h0, h1, f1, b11);
                    const int j = J + 2 * tig;
                    if (j <= jmax_a) Ca[J] -= c0;                  // (tiles right of a row's diagonal: nothing stored)
                    if (j + 1 <= jmax_a) Ca[J + 1] -= c1;
                    if (j + 8 <= jmax_a) Ca[J + 8] -= e0;
                    if (j + 9 <= jmax_a) Ca[J + 9] -= e1;
                    if (j <= jmax_b) Cb[J] -= g0;
                    if (j + 1 <= jmax_b) Cb[J + 1] -= g1;
                    if (j + 8 <= jmax_b) Cb[J + 8] -= h0;
                    if (j + 9 <= jmax_b) Cb[J + 9] -= h1;
                    Lj = Lj2 + (jb + 16 < n ? 8 * (jb + 8) + 36 : 0);
                    jb += 16;
                }
            }
            __syncwarp();
        }
#ifdef TCHEM_SPARSE_PHASES
        if (blockIdx.x == 0 && threadIdx.x == 0) g_sparse_phase[8] += (unsigned long long)(clock64() - sub_t1);
#endif
    }
    return ok;
}

// L^T x = y, y = the bordered row, blocks of 8 unknowns from the bottom: the block's own 8 x 8 triangle by shuffles
// (one shuffle, one multiply and one fma on the chain per unknown), then one rank-8 update of everything above.
// x[m] = element lane + 32 m.
__device__ __forceinline__ void warp_backsolve_packed(const double * A, int n, const double * rdiag, double * x, int lane) {
#pragma unroll
    for (int m = 0; m < kMaxSlots; m++) {
        const int i = lane + 32 * m;
        x[m] = i < n ? A[tri(n) + i] : 0.0;
    }
    for (int k0 = ((n - 1) >> 3) << 3; k0 >= 0; k0 -= 8) {
        const int bs = min(8, n - k0);
        const int km = k0 >> 5, kl0 = k0 & 31;
        double xb = x[0];                               // the slot that holds the block (a block never straddles slots)
#pragma unroll
        for (int m = 1; m < kMaxSlots; m++) xb = km == m ? x[m] : xb;
        const int ib = lane + 32 * km;                  // this lane's element of that slot
        // everything the chain needs is fetched up front (clamped addresses, no divergent branches on the chain):
        // the block's reciprocal diagonal and, for the lanes inside the block, their column of the 8 x 8 triangle
        double rd[8], lcol[8];
        const bool inblk = ib >= k0 && ib < k0 + bs;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int k = min(k0 + c, n - 1);
            rd[c] = rdiag[k];
            lcol[c] = (inblk && ib < k0 + c && c < bs) ? A[tri(k) + ib] : 0.0;
        }
        double xk[8];
#pragma unroll
        for (int c = 7; c >= 0; c--) {
            const double v = __shfl_sync(kFull, xb, (kl0 + c) & 31) * rd[c];
            xk[c] = c < bs ? v : 0.0;
            xb = (lane == kl0 + c && c < bs) ? v : fma(-lcol[c], xk[c], xb);
        }
#pragma unroll
        for (int m = 0; m < kMaxSlots; m++) {
            if (m == km) x[m] = xb;
            if (32 * m < k0) {
                const int i = lane + 32 * m, ic = min(i, k0 - 1);
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    s0 = fma(A[tri(min(k0 + c, n - 1)) + ic], xk[c], s0);            // (xk = 0 beyond the block)
                    s1 = fma(A[tri(min(k0 + c + 1, n - 1)) + ic], xk[c + 1], s1);
                }
                if (i < k0) x[m] -= s0 + s1;
            }
        }
    }
}

// in place: L -> W = L^-1 (lower), row panels I of 8 from the top:
//   D = L[I][I]^-1 (registers, forward substitution, lane & 7 = column),
//   W[I][J] = -D (L[I][J:i0] W[J:i0][J]) for the column tiles J = 0, 8, .. in ASCENDING order: tile J reads L[I][k >= J]
//   only, so each finished tile can overwrite its own columns of rows I at once.
// Ds: 192 doubles of scratch (D, and the two tiles in transit between the DMMA products)
__device__ __forceinline__ void warp_trtri_packed(double * A, int n, const double * rdiag, double * Ds, int lane) {
    const int gid = lane >> 2, tig = lane & 3;
    double * Ts = Ds + 64;
    for (int i0 = 0; i0 < n; i0 += kNB) {
        const int bs = min(kNB, n - i0);
        const int c = lane & 7;
        double Dc[kNB];
#pragma unroll
        for (int r = 0; r < kNB; r++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < r; k++) {
                const double l = r < bs ? A[tri(i0 + r) + i0 + k] : 0.0;
                if (k >= c) s = fma(l, Dc[k], s);
            }
            const double rd = r < bs ? rdiag[i0 + r] : 0.0;
            Dc[r] = (r < c || c >= bs) ? 0.0 : r == c ? rd : -s * rd;
        }
        if (lane < kNB) {
#pragma unroll
            for (int r = 0; r < kNB; r++) Ds[r * kNB + lane] = Dc[r];
        }
        __syncwarp();
        const double d0 = -Ds[gid * kNB + tig], d1 = -Ds[gid * kNB + 4 + tig];      // A fragments of -D
        const int ri = min(i0 + gid, n - 1);
        const double * Lr = A + tri(ri) + tig;
        // two column tiles (J, J + 8) per pass share the A fragments: two independent DMMA chains. Both are computed
        // before either is written (tile J reads the columns tile J + 8 overwrites).
        for (int J = 0; J < i0; J += 16) {
            const bool two = J + 8 < i0;
            double t0 = 0.0, t1 = 0.0, u0 = 0.0, u1 = 0.0;
            // k-steps J and J + 4 meet the diagonal of W (zero above it); from J + 8 on every element of tile J is stored,
            // from J + 16 on every element of tile J + 8. Row pointer of W: tri(k + 4) - tri(k) = 4 k + 10.
            int kk = J + tig;
            const double * Wk = A + tri(kk) + J + gid;
            const int jj = J + gid;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                dmma(t0, t1, Lr[kk - tig], jj <= kk ? Wk[0] : 0.0);
                Wk += 4 * kk + 10;
                kk += 4;
            }
            if (two) {
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const double a = Lr[kk - tig];
                    dmma(t0, t1, a, Wk[0]);
                    dmma(u0, u1, a, jj + 8 <= kk ? Wk[8] : 0.0);
                    Wk += 4 * kk + 10;
                    kk += 4;
                }
                for (int k0 = J + 16; k0 < i0; k0 += 4) {
                    const double a = Lr[k0];
                    dmma(t0, t1, a, Wk[0]);
                    dmma(u0, u1, a, Wk[8]);
                    Wk += 4 * kk + 10;
                    kk += 4;
                }
            }
            __syncwarp();                                 // rows I, columns >= J have been read by every lane
            Ts[gid * kNB + 2 * tig] = t0;
            Ts[gid * kNB + 2 * tig + 1] = t1;
            Ts[64 + gid * kNB + 2 * tig] = u0;
            Ts[64 + gid * kNB + 2 * tig + 1] = u1;
            __syncwarp();
            double r0 = 0.0, r1 = 0.0, q0 = 0.0, q1 = 0.0;
            dmma(r0, r1, d0, Ts[tig * kNB + gid]);
            dmma(q0, q1, d0, Ts[64 + tig * kNB + gid]);
            dmma(r0, r1, d1, Ts[(4 + tig) * kNB + gid]);
            dmma(q0, q1, d1, Ts[64 + (4 + tig) * kNB + gid]);
            if (gid < bs) {
                double * o = A + tri(i0 + gid) + J + 2 * tig;
                o[0] = r0; o[1] = r1;
                if (two) { o[8] = q0; o[9] = q1; }
            }
        }
        __syncwarp();
        if (lane < bs) {
#pragma unroll
            for (int r = 0; r < kNB; r++)
                if (r < bs && r >= lane) A[tri(i0 + r) + i0 + lane] = Dc[r];
        }
        __syncwarp();
    }
}

// in place: W (lower) -> lower triangle of W^T W, row panels I of 8 ascending (rows below a panel are still W); tile
// (I, J) = sum over k >= i0 of W[k][I]^T W[k][J], column tiles ascending so that the diagonal tile - the only one
// whose columns the others read in rows I - is written last
__device__ __forceinline__ void warp_lauum_packed(double * A, int n, int lane) {
    const int gid = lane >> 2, tig = lane & 3;
    for (int i0 = 0; i0 < n; i0 += kNB) {
        const int ra = i0 + gid;
        // two column tiles (J, J + 8) per pass share the A fragments W[k][I]; the pair that holds the diagonal tile comes last
        for (int J = 0; J <= i0; J += 16) {
            const bool two = J + 8 <= i0;
            const int cb = J + gid;
            double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
            // k-steps i0 and i0 + 4 meet the diagonal block; from i0 + 8 on every element of the fragments is stored
            // (only rows past the end are masked). Row pointer: tri(k + 4) - tri(k) = 4 k + 10.
            int kk = i0 + tig;
            const double * Wk = A + tri(min(kk, n - 1));
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const double av = (kk < n && ra <= kk) ? Wk[ra] : 0.0;
                dmma(c0, c1, av, (kk < n && cb <= kk) ? Wk[cb] : 0.0);
                dmma(e0, e1, av, (two && kk < n && cb + 8 <= kk) ? Wk[cb + 8] : 0.0);
                if (kk + 4 < n) Wk += 4 * kk + 10;
                kk += 4;
            }
            const int ra_c = min(ra, n - 1);               // rows of the result past the end are never stored: any valid column will do
            const int cb2 = two ? cb + 8 : cb;
            for (int k0 = i0 + 8; k0 + 3 < n; k0 += 4) {
                const double av = Wk[ra_c];
                dmma(c0, c1, av, Wk[cb]);
                dmma(e0, e1, av, Wk[cb2]);
                Wk += 4 * kk + 10;
                kk += 4;
            }
            if (kk - tig < n && kk - tig >= i0 + 8) {      // tail k-step (n not a multiple of 4)
                const int kc = min(kk, n - 1);
                const double * Wt = A + tri(kc);
                const double av = kk < n ? Wt[ra_c] : 0.0;
                dmma(c0, c1, av, kk < n ? Wt[cb] : 0.0);
                dmma(e0, e1, av, kk < n ? Wt[cb2] : 0.0);
            }
            __syncwarp();
            const int j = J + 2 * tig;
            if (ra < n) {
                double * o = A + tri(ra) + j;
                if (j <= ra) o[0] = c0;
                if (j + 1 <= ra) o[1] = c1;
                if (two && j + 8 <= ra) o[8] = e0;
                if (two && j + 9 <= ra) o[9] = e1;
            }
            __syncwarp();
        }
    }
}

__host__ __device__ inline int even_up(int x) { return (x + 1) & ~1; }
// shared memory of one warp (doubles): A[tri(n) + n + 2] | rdiag[n] | Ds[192] | rg[cd] | g[cd] | Jv[nnz_pad]; with jv_alias Jv
// lives in the TAIL of A instead (it is dead once J J^T and the bordered row exist): 32 KB instead of 38 KB per geometry of 84
// coordinates, i.e. six geometries in flight per SM instead of five
__host__ __device__ inline int factor_smem_doubles(int n, int cd, int nnz_pad, int jv_alias) {
    return (jv_alias ? 0 : nnz_pad) + ((n * (n + 1)) / 2 + n + 2) + ((n + 1) & ~1) + 192 + 2 * ((cd + 1) & ~1);
}
constexpr int kMaxHiPairs = 8;        // pairs per lane that can wait in registers while Jv still occupies their destination

// OPF 0: g_q = (J J^T)^-1 (J g_r)                     -> gq_out[B][n]
// OPF 1: the same (g_r may be null) plus the inverse as a full symmetric matrix -> ainv_out[B][even_up(n n)] (what the block
// kernel's single bulk copy per geometry wants; expanding the packed form there cost 7 600 cycles of 8-byte cp.async issue)
// The block's warps work on different geometries and never meet after the definition tables (the leading
// `factor_tab_bytes` of the blob: pair lists, terms, compressed-row structure) have been copied to shared memory:
// every table read is a shared-memory read (from global memory each of them is an L2 round trip on the critical path
// of a single warp: measured 10x slower).
template <int OPF, bool HAS5>
__global__ void __launch_bounds__(192)
factor_kernel(const SparseProgram sp_global, const double * __restrict__ r, const double * __restrict__ gr, const int64_t B,
              double * __restrict__ gq_out, double * __restrict__ ainv_out, int * __restrict__ status,
              unsigned long long * __restrict__ next_geometry) {
    extern __shared__ __align__(16) double sm_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int n = sp_global.intdim, cd = sp_global.cartdim;
    const int ntri = (n * (n + 1)) >> 1;
    SparseProgram sp = sp_global;
    const int tab_doubles = (sp_global.factor_tab_bytes + 7) / 8;
    {
        const double * src = reinterpret_cast<const double *>(sp_global.pair_dest);          // the blob starts with the pair lists
        for (int i = threadIdx.x; i < tab_doubles; i += blockDim.x) sm_dyn[i] = src[i];
        const char * base = reinterpret_cast<const char *>(sp_global.pair_dest), * tb = reinterpret_cast<const char *>(sm_dyn);
#define TC_REBASE(field, type) sp.field = reinterpret_cast<const type *>(tb + (reinterpret_cast<const char *>(sp_global.field) - base))
        TC_REBASE(pair_dest, uint32_t); TC_REBASE(pair_ent, uint32_t); TC_REBASE(terms, SpTerm); TC_REBASE(row_ptr, int32_t);
        TC_REBASE(jrow_order, int32_t); TC_REBASE(nz_ptr, uint16_t); TC_REBASE(nz_col, uint16_t);
#undef TC_REBASE
    }
    __syncthreads();
    const int atot = ntri + n + 2;
    double * A = sm_dyn + even_up(tab_doubles) + (size_t)warp * even_up(factor_smem_doubles(n, cd, sp.nnz_pad, sp.jv_alias));
    double * rdiag = A + atot;
    double * Ds = rdiag + ((n + 1) & ~1);
    double * rg = Ds + 192, * gs = rg + ((cd + 1) & ~1);
    // Jv: behind everything else, or (jv_alias) in the tail of A: the pairs / bordered-row entries whose destination it
    // covers are formed in registers and written once Jv is dead
    double * Jv = sp.jv_alias ? A + (atot - sp.nnz_pad) : gs + ((cd + 1) & ~1);
    const int lo_end = sp.jv_alias ? atot - sp.nnz_pad : atot;             // A[0 .. lo_end) does not overlap Jv
    SPHASE_BEGIN(Ds);
    // Geometries are handed out one at a time (a global counter owned by the launch, zeroed before it; every warp starts on its own index
    // and asks for the next one while it works on the current): six warps on four schedulers means two schedulers host two
    // warps, whose DMMA and FP64 issue they share - with a fixed stride the two lone warps would idle a quarter of the time.
    int64_t b = (int64_t)blockIdx.x * nwarps + warp;
    while (b < B) {
        unsigned long long taken = 0;
        if (lane == 0) taken = atomicAdd(next_geometry, 1ull);
        // the geometry (and the gradient) are read once, coalesced: every later use is a shared-memory read
        for (int i = lane; i < cd; i += kLanes) { rg[i] = r[b * cd + i]; if (gr != nullptr) gs[i] = gr[b * cd + i]; }
        for (int p = sp.nnz + lane; p < sp.nnz_pad; p += kLanes) Jv[p] = 0.0;      // the zero entries the padded pair lists point to
        __syncwarp();
        eval_J_rows<HAS5>(sp, rg, Jv, lane, kLanes);
        for (int p = lane; p < lo_end; p += kLanes) A[p] = 0.0;
        __syncwarp();
        SPHASE(0);
        // A = J J^T over the structurally nonzero pairs of rows: 3 products per common atom
        auto pair_value = [&](int p) {
            double s = 0.0;
            for (int e = 0; e < sp.maxc; e += 4) {                                  // (lists padded to a multiple of 4 with zero entries)
                uint32_t ent[4];
#pragma unroll
                for (int u = 0; u < 4; u++) ent[u] = sp.pair_ent[(e + u) * sp.npairs_pad + p];
                double t[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const double * ji = Jv + (ent[u] >> 16), * jj = Jv + (ent[u] & 0xffffu);
                    t[u] = fma(ji[2], jj[2], fma(ji[1], jj[1], ji[0] * jj[0]));
                }
                s += (t[0] + t[1]) + (t[2] + t[3]);
            }
            return s;
        };
        for (int p = lane; p < sp.npairs_lo; p += kLanes) A[sp.pair_dest[p]] = pair_value(p);
        // the pairs whose destination Jv still occupies (none without jv_alias; the padding pairs write the spare element)
        // and the bordered row b = J g_r wait in registers
        double hv[kMaxHiPairs], bv[kMaxSlots];
#pragma unroll
        for (int k = 0; k < kMaxHiPairs; k++) {
            const int p = sp.npairs_lo + lane + 32 * k;
            hv[k] = p < sp.npairs_pad ? pair_value(p) : 0.0;
        }
#pragma unroll
        for (int m = 0; m < kMaxSlots; m++) {
            const int i = lane + 32 * m;
            double s = 0.0;
            if (gr != nullptr && i < n)
                for (int p = sp.nz_ptr[i]; p < sp.nz_ptr[i + 1]; p++) s = fma(Jv[p], gs[sp.nz_col[p]], s);
            bv[m] = s;
        }
        __syncwarp();                                                               // Jv is dead
        for (int p = lo_end + lane; p < atot; p += kLanes) A[p] = 0.0;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < kMaxHiPairs; k++) {
            const int p = sp.npairs_lo + lane + 32 * k;
            if (p < sp.npairs_pad) A[sp.pair_dest[p]] = hv[k];
        }
#pragma unroll
        for (int m = 0; m < kMaxSlots; m++) {
            const int i = lane + 32 * m;
            if (i < n) A[ntri + i] = bv[m];
        }
        __syncwarp();
        SPHASE(1);
        const bool ok = warp_potrf_packed(A, n, true, rdiag, lane);
        SPHASE(2);
        if (!ok) {
            // the reference's at::cholesky throws (IntCoordSet.cpp:182): flag the table, NaN for this geometry
            if (lane == 0) atomicExch(status, 1);
            if (gq_out) for (int i = lane; i < n; i += kLanes) gq_out[b * n + i] = nan64();
            if (OPF == 1) for (int p = lane; p < n * n; p += kLanes) ainv_out[b * even_up(n * n) + p] = nan64();
            __syncwarp();
            b = (int64_t)gridDim.x * nwarps + (int64_t)__shfl_sync(kFull, taken, 0);
            continue;
        }
        if (gq_out) {
            double x[kMaxSlots];
            warp_backsolve_packed(A, n, rdiag, x, lane);
#pragma unroll
            for (int m = 0; m < kMaxSlots; m++) {
                const int i = lane + 32 * m;
                if (i < n) gq_out[b * n + i] = x[m];
            }
        }
        SPHASE(3);
        if (OPF == 1) {
            warp_trtri_packed(A, n, rdiag, Ds, lane);
            SPHASE(4);
            warp_lauum_packed(A, n, lane);
            SPHASE(5);
            // full symmetric matrix, four rows per trip (their loads are independent: one latency per trip, not per element)
            double * dst = ainv_out + b * even_up(n * n);
            int jc[kMaxSlots], tj[kMaxSlots];
#pragma unroll
            for (int m = 0; m < kMaxSlots; m++) { jc[m] = min(lane + 32 * m, n - 1); tj[m] = tri(jc[m]); }
            for (int i0 = 0; i0 < n; i0 += 4) {
                double v[4][kMaxSlots];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = min(i0 + u, n - 1), ti = tri(i);
#pragma unroll
                    for (int m = 0; m < kMaxSlots; m++)
                        if (32 * m < n) v[u][m] = A[jc[m] <= i ? ti + jc[m] : tj[m] + i];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
#pragma unroll
                    for (int m = 0; m < kMaxSlots; m++)
                        if (i0 + u < n && lane + 32 * m < n) __stcg(dst + (i0 + u) * n + lane + 32 * m, v[u][m]);
                }
            }
        }
        __syncwarp();
        SPHASE(6);
        b = (int64_t)gridDim.x * nwarps + (int64_t)__shfl_sync(kFull, taken, 0);
    }
}

// ======================================================================================================
// hess_kernel: one block = one geometry at a time
// ======================================================================================================
// shared-memory plan (doubles)
struct HessPlan { int small, jv, big, mid, third, tab, total; };
__host__ __device__ inline int ld_dmma(int width) { int ld = width; while ((ld & 15) != 4 && (ld & 15) != 12) ld++; return ld; }
__host__ __device__ inline HessPlan hess_plan(int op, int n, int cd, int ldh, int nnz_pad, int nslots, int tab_doubles) {
    HessPlan p;
    const int ldo = n | 1, ldj = ld_dmma(cd);              // odd leading dimensions: lanes that run down a column hit different banks
    p.small = 0;                                      // 2 x (coordinates | vector)
    p.jv = p.small + 2 * (even_up(cd) + even_up(cd > n ? cd : n));
    p.big = p.jv + 2 * nnz_pad;
    int big = 0, mid = 0, third = 0;
    switch (op) {
    case 2:  big = n * n; mid = max(cd * ldo, nslots); third = cd * cd; break;            // H_q | U, slots | X
    case 3:  big = max(cd * ldj, n * ldh); mid = max(nslots, cd * ldh); third = cd * ldh; break;   // H_c, H_q | slots, M^T | inverse, T2
    default: big = n * ldh; mid = n * cd; third = 0; break;                                // inverse | M
    }
    p.mid = p.big + even_up(big);
    p.third = p.mid + even_up(mid);
    p.tab = p.third + even_up(third);
    p.total = p.tab + even_up(tab_doubles);
    return p;
}

// global [count] -> shared, 16 bytes per cp.async when both sides allow, else 8; all threads, one commit group
__device__ __forceinline__ void block_copy_async(const double * __restrict__ g, double * s, int count) {
    const bool vec = (count & 1) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s)) & 15) == 0;
    if (vec) {
        for (int i = 2 * threadIdx.x; i < count; i += 2 * blockDim.x)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(s + i)), "l"(g + i) : "memory");
    } else {
        for (int i = threadIdx.x; i < count; i += blockDim.x)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((uint32_t)__cvta_generic_to_shared(s + i)), "l"(g + i) : "memory");
    }
}
// global [rows][cols] -> shared [rows][ld], one warp per row at a time (no commit)
__device__ __forceinline__ void block_rows_async(const double * __restrict__ g, double * s, int rows, int cols, int ld) {
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const bool vec = ((cols | ld) & 1) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s)) & 15) == 0;
    for (int r = warp; r < rows; r += nwarps) {
        const double * src = g + (int64_t)r * cols;
        double * dst = s + r * ld;
        if (vec) {
            for (int c = 2 * lane; c < cols; c += 2 * kLanes)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(dst + c)), "l"(src + c) : "memory");
        } else {
            for (int c = lane; c < cols; c += kLanes)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((uint32_t)__cvta_generic_to_shared(dst + c)), "l"(src + c) : "memory");
        }
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// shared [rows][ld] -> global [rows][cols] (streaming stores, 16 bytes when everything is aligned)
__device__ __forceinline__ void block_store_rows(const double * s, double * __restrict__ g, int rows, int cols, int ld) {
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
    if (ld == cols && ((rows * cols) & 1) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s)) & 15) == 0) {
        const int total = rows * cols;
        for (int i = 2 * threadIdx.x; i < total; i += 2 * blockDim.x)
            __stcs(reinterpret_cast<double2 *>(g + i), *reinterpret_cast<const double2 *>(s + i));
        return;
    }
    // padded rows of even length: 16 bytes per lane (two thirds of the store instructions of the 8-byte loop for 84 columns)
    const bool vec = ((cols | ld) & 1) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s)) & 15) == 0;
    for (int r = warp; r < rows; r += nwarps) {
        const double * src = s + r * ld;
        double * dst = g + (int64_t)r * cols;
        if (vec) for (int c = 2 * lane; c < cols; c += 2 * kLanes) __stcs(reinterpret_cast<double2 *>(dst + c), *reinterpret_cast<const double2 *>(src + c));
        else     for (int c = lane; c < cols; c += kLanes) __stcs(dst + c, src[c]);
    }
}

__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t smem_addr(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }

// The products with J run over compressed column BLOCKS: the x, y, z columns of an atom have the same row list, so one
// operand load feeds three FMAs (one shared-memory load per FMA would make the phase shared-memory bound), and a warp
// owns one atom at a time with up to 3 x 4 accumulators.
//
// OUT[3 A + c][l] = sum over the rows k of atom A of  J[k][3 A + c] * IN[k][l]     (OUT: cartdim x width): OUT = J^T IN
// lanes along l (contiguous in IN)
template <int NS>
__device__ __forceinline__ void spmm_JT_slots(const SparseProgram & sp, const double * Jv, const double * IN, int ldin, int width,
                                              double * OUT, int ldout) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t in_m[NS];                 // (lanes past the end read a clamped address; their sums are never stored)
#pragma unroll
    for (int m = 0; m < NS; m++) in_m[m] = smem_addr(IN) + 8u * (uint32_t)(lane + 32 * m < width ? lane + 32 * m : lane);
    const uint32_t jv0 = smem_addr(Jv), rowb = 8u * (uint32_t)ldin;
    for (int ai = sp.aw_ptr[warp]; ai < sp.aw_ptr[warp + 1]; ai++) {
        const int A = sp.aw_atom[ai];
        double acc[3][NS];
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int m = 0; m < NS; m++) acc[c][m] = 0.0;
        int e = sp.col_ptr[A];
        const int e1 = sp.col_ptr[A + 1];
        // the entry word and the three J values of the NEXT entry are fetched while this one's FMAs run (the chain
        // entry -> position -> values is two load latencies; col_ent is padded and position 0 is always valid)
        uint32_t ent = sp.col_ent[e];
        uint32_t jp = jv0 + 8u * (ent & 0xffffu);
        double v0 = lds64(jp), v1 = lds64(jp + 8), v2 = lds64(jp + 16);
        for (; e < e1; e++) {
            const uint32_t row = (ent >> 16) * rowb;
            double x[NS];
#pragma unroll
            for (int m = 0; m < NS; m++) x[m] = lds64(in_m[m] + row);
            ent = sp.col_ent[e + 1];
            jp = jv0 + 8u * (ent & 0xffffu);
            const double n0 = lds64(jp), n1 = lds64(jp + 8), n2 = lds64(jp + 16);
#pragma unroll
            for (int m = 0; m < NS; m++) {
                acc[0][m] = fma(v0, x[m], acc[0][m]);
                acc[1][m] = fma(v1, x[m], acc[1][m]);
                acc[2][m] = fma(v2, x[m], acc[2][m]);
            }
            v0 = n0; v1 = n1; v2 = n2;
        }
#pragma unroll
        for (int m = 0; m < NS; m++) {
            const int l = lane + 32 * m;
            if (l < width) {
#pragma unroll
                for (int c = 0; c < 3; c++) OUT[(3 * A + c) * ldout + l] = acc[c][m];
            }
        }
    }
}
__device__ __forceinline__ void spmm_JT(const SparseProgram & sp, const double * Jv, const double * IN, int ldin, int width,
                                        double * OUT, int ldout) {
    switch ((width + 31) >> 5) {       // the slot count as a compile-time constant: no predicates in the entry loop
    case 1:  spmm_JT_slots<1>(sp, Jv, IN, ldin, width, OUT, ldout); break;
    case 2:  spmm_JT_slots<2>(sp, Jv, IN, ldin, width, OUT, ldout); break;
    case 3:  spmm_JT_slots<3>(sp, Jv, IN, ldin, width, OUT, ldout); break;
    default: spmm_JT_slots<4>(sp, Jv, IN, ldin, width, OUT, ldout); break;
    }
}
// OUT[a][3 B + c] = sum over the rows l of atom B of  IN[a][l] * J[l][3 B + c]     (OUT: height x cartdim): OUT = IN J
// lanes along a (IN's leading dimension is odd: conflict-free)
template <int NS>
__device__ __forceinline__ void spmm_J_slots(const SparseProgram & sp, const double * Jv, const double * IN, int ldin, int height,
                                             double * OUT, int ldout) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t in_m[NS];
#pragma unroll
    for (int m = 0; m < NS; m++) in_m[m] = smem_addr(IN) + 8u * (uint32_t)(min(lane + 32 * m, height - 1) * ldin);
    const uint32_t jv0 = smem_addr(Jv);
    for (int ai = sp.aw_ptr[warp]; ai < sp.aw_ptr[warp + 1]; ai++) {
        const int Bc = sp.aw_atom[ai];
        double acc[3][NS];
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int m = 0; m < NS; m++) acc[c][m] = 0.0;
        int e = sp.col_ptr[Bc];
        const int e1 = sp.col_ptr[Bc + 1];
        uint32_t ent = sp.col_ent[e];                                       // next entry's words and values in flight: see spmm_JT
        uint32_t jp = jv0 + 8u * (ent & 0xffffu);
        double v0 = lds64(jp), v1 = lds64(jp + 8), v2 = lds64(jp + 16);
        for (; e < e1; e++) {
            const uint32_t off = 8u * (ent >> 16);
            double x[NS];
#pragma unroll
            for (int m = 0; m < NS; m++) x[m] = lds64(in_m[m] + off);
            ent = sp.col_ent[e + 1];
            jp = jv0 + 8u * (ent & 0xffffu);
            const double n0 = lds64(jp), n1 = lds64(jp + 8), n2 = lds64(jp + 16);
#pragma unroll
            for (int m = 0; m < NS; m++) {
                acc[0][m] = fma(v0, x[m], acc[0][m]);
                acc[1][m] = fma(v1, x[m], acc[1][m]);
                acc[2][m] = fma(v2, x[m], acc[2][m]);
            }
            v0 = n0; v1 = n1; v2 = n2;
        }
#pragma unroll
        for (int m = 0; m < NS; m++) {
            const int a = lane + 32 * m;
            if (a < height) {
                double * o = OUT + a * ldout + 3 * Bc;
                o[0] = acc[0][m]; o[1] = acc[1][m]; o[2] = acc[2][m];
            }
        }
    }
}
__device__ __forceinline__ void spmm_J(const SparseProgram & sp, const double * Jv, const double * IN, int ldin, int height,
                                       double * OUT, int ldout) {
    switch ((height + 31) >> 5) {
    case 1:  spmm_J_slots<1>(sp, Jv, IN, ldin, height, OUT, ldout); break;
    case 2:  spmm_J_slots<2>(sp, Jv, IN, ldin, height, OUT, ldout); break;
    case 3:  spmm_J_slots<3>(sp, Jv, IN, ldin, height, OUT, ldout); break;
    default: spmm_J_slots<4>(sp, Jv, IN, ldin, height, OUT, ldout); break;
    }
}

// (term, direction) second-derivative tasks of geometry `rg` into the slots, weights w[row] * sign * coef, and the J
// rows of the NEXT geometry (rg_next == nullptr: none). The tasks come as warp steps of 32 same-kind tasks (uniform
// control flow) that a static schedule built with the table deals out to the block's warps by estimated cost, so the
// phase lasts about (total cost / warps) instead of (rounds x the most expensive kind).
template <bool HAS5>
__device__ __forceinline__ void run_tasks(const SparseProgram & sp, const double * rg, const double * weight, double sign, double * slots,
                                          const double * rg_next, double * Jv_next) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef TCHEM_SPARSE_PHASES
    const long long steps_t0 = clock64();
    long long step_t0 = steps_t0;
    int step_prev = -1;
#endif
    for (int s = sp.sched_ptr[warp]; s < sp.sched_ptr[warp + 1]; s++) {
#ifdef TCHEM_SPARSE_PHASES
        __syncwarp();
        if (blockIdx.x == 0 && lane == 0) {
            const long long t = clock64();
            if (step_prev >= 0 && step_prev < 64) g_sparse_step[step_prev] += (unsigned long long)(t - step_t0);
            step_t0 = t; step_prev = sp.sched_step[s];
        }
#endif
        const uint32_t task = sp.ktasks[32 * sp.sched_step[s] + lane];
        if (task == 0xffffffffu) continue;
        if (!(task & 0x80000000u)) {
            const SpTerm & term = sp.terms[task >> 8];
            const int dir = (int)(task & 0xffu);
            SlotSink sink;
            sink.S = slots + term.kslot;
            sink.w = sign * weight[term.pd.out] * term.coef;
            sink.ends_free = term.pd.type == TCHEM_TORSION;
            eval_primitive<MODE_KDIR, HAS5>(term.pd, GeomLoader{rg}, sink, dir, dir + 1);
        } else if (rg_next != nullptr) {
            const int row = (int)(task & 0x7fffffffu);
            for (int p = sp.nz_ptr[row]; p < sp.nz_ptr[row + 1]; p++) Jv_next[p] = 0.0;
            for (int t = sp.row_ptr[row]; t < sp.row_ptr[row + 1]; t++) {
                const SpTerm & term = sp.terms[t];
                JvSink sink;
                sink.Jv = Jv_next; sink.pos = term.pos; sink.coef = term.coef;
                eval_primitive<MODE_QJ, HAS5>(term.pd, GeomLoader{rg_next}, sink);
            }
        }
    }
#ifdef TCHEM_SPARSE_PHASES
    __syncwarp();
    if (blockIdx.x == 0 && lane == 0) {
        const long long t = clock64();
        if (step_prev >= 0 && step_prev < 64) g_sparse_step[step_prev] += (unsigned long long)(t - step_t0);
        if (warp < 16) g_sparse_step[64 + warp] += (unsigned long long)(t - steps_t0);
    }
#endif
}

// H[3 A + k][3 B + l] += sum of the block's slots, in table order (no atomics: bit-reproducible); one thread per
// (block, row k), the three columns l together
__device__ __forceinline__ void gather_slots(const SparseProgram & sp, const double * slots, double * H, int ld) {
    for (int e = threadIdx.x; e < 3 * sp.ngblk; e += blockDim.x) {
        const int blk = e / 3, k = e - 3 * blk;
        const uint32_t ab = sp.gk_blk[blk];
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        const uint32_t p0 = sp.gk_ptr[blk], p1 = sp.gk_ptr[blk + 1];
        // four sources per trip: their table words, then their slots, are loaded together, so a block with many sources
        // (the diagonal block of a ring atom) pays one load latency per four instead of per source; the additions stay
        // in table order
        for (uint32_t p = p0; p < p1; p += 4) {
            uint32_t src[4];
            double v[4][3], sg[4];
#pragma unroll
            for (int u = 0; u < 4; u++) src[u] = sp.gk_src[min(p + u, p1 - 1)];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const bool tr = src[u] & 1u;
                const double * q = slots + (src[u] >> 2) + (tr ? k : 3 * k);
                const int st = tr ? 3 : 1;
                v[u][0] = q[0]; v[u][1] = q[st]; v[u][2] = q[2 * st];
                // negated: block (0, 0) of a term from translational invariance; past the end: no contribution
                sg[u] = (p + u < p1) ? ((src[u] & 2u) ? -1.0 : 1.0) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) { s0 = fma(sg[u], v[u][0], s0); s1 = fma(sg[u], v[u][1], s1); s2 = fma(sg[u], v[u][2], s2); }
        }
        double * h = H + (3 * (ab >> 16) + k) * ld + 3 * (ab & 0xffffu);
        h[0] += s0; h[1] += s1; h[2] += s2;
    }
}

// shared [count] (contiguous) -> global through the TMA unit (1-D bulk copy): issued by one thread AFTER every thread
// has fenced its generic-proxy writes and the block has met; the region may be rewritten once bulk_store_read_done()
// has returned. False: shape / alignment does not allow it (the caller stores with ordinary instructions).
__device__ __forceinline__ bool bulk_store_ok(const double * s, const double * g, int count) {
    return (count & 1) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s)) & 15) == 0;
}
__device__ __forceinline__ void bulk_store_issue(const double * s, double * g, int count) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(g), "r"(smem_addr(s)), "r"(count * 8) : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void bulk_store_read_done() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_store_all_done() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// global -> shared through the TMA unit (1-D bulk copy) with completion on an mbarrier: one instruction per matrix (or per
// row when the shared-memory rows are padded) instead of thousands of per-thread cp.async - the issue of those was a phase
// of its own. Protocol per geometry: thread 0 arms the barrier with the byte count, the copies complete_tx on it, every
// thread waits for the phase (parity flips per use).
__device__ __forceinline__ void mbar_init(uint64_t * mb, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(mb)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * mb, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr(mb)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(double * dst, const double * src, uint32_t bytes, uint64_t * mb) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(mb)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * mb, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@!p bra WAIT_%=;\n}\n" ::"r"(smem_addr(mb)), "r"(parity) : "memory");
}

// OP 0: gradient_cart2int_matrix   ainv (workspace)                                  out = M[n][cd]
// OP 2: Hessian_int2cart           in1 = g_q[n],                in2 = H_q[n][n]      out = H_r[cd][cd]
// OP 3: Hessian_cart2int           in1 = g_q[n] (workspace), ainv, in2 = H_r[cd][cd] out = H_q[n][n]
// TABS: the definition tables are copied to shared memory (the usual case; as a template parameter so that every table
// read is a shared-state-space load). cp.async groups per thread, oldest first: [small inputs of the next geometry]
// [operand matrix (+ inverse) of the next geometry].
template <int OP, bool HAS5, bool TABS>
__global__ void __launch_bounds__(kHessThreads, 1)
hess_kernel(const SparseProgram sp_global, const double * __restrict__ r, const double * __restrict__ in1,
            const double * __restrict__ in2, const double * __restrict__ ainv, const int64_t B, double * __restrict__ out) {
    extern __shared__ __align__(16) double sm[];
    __shared__ __align__(8) uint64_t big_mbar;          // completion of the bulk copies of a geometry's matrices
    if (threadIdx.x == 0) { mbar_init(&big_mbar, 1); fence_async_proxy(); }
    const int n = sp_global.intdim, cd = sp_global.cartdim, ldh = sp_global.ldh;
    const int tab_doubles = TABS ? (sp_global.blob_bytes - sp_global.hess_tab_offset + 7) / 8 : 0;
    const HessPlan pl = hess_plan(OP, n, cd, ldh, sp_global.nnz_pad, sp_global.nslots, tab_doubles);
    SparseProgram sp = sp_global;
    if (TABS) {
        double * tab = sm + pl.tab;
        const double * src = reinterpret_cast<const double *>(sp_global.terms);              // first of the common tables
        for (int i = threadIdx.x; i < tab_doubles; i += blockDim.x) tab[i] = src[i];
        const char * base = reinterpret_cast<const char *>(sp_global.terms), * tb = reinterpret_cast<const char *>(tab);
#define TC_REBASE(field, type) sp.field = reinterpret_cast<const type *>(tb + (reinterpret_cast<const char *>(sp_global.field) - base))
        TC_REBASE(terms, SpTerm); TC_REBASE(row_ptr, int32_t); TC_REBASE(jrow_order, int32_t); TC_REBASE(ktasks, uint32_t);
        TC_REBASE(nz_ptr, uint16_t); TC_REBASE(nz_col, uint16_t); TC_REBASE(col_ptr, uint16_t); TC_REBASE(col_ent, uint32_t);
        TC_REBASE(aw_ptr, uint16_t); TC_REBASE(aw_atom, uint16_t);
        TC_REBASE(gk_blk, uint32_t); TC_REBASE(gk_ptr, uint32_t); TC_REBASE(gk_src, uint32_t);
        TC_REBASE(sched_ptr, uint16_t); TC_REBASE(sched_step, uint16_t);
#undef TC_REBASE
    }
    const int ldo = n | 1, ldj = ld_dmma(cd);
    const int rg_buf = even_up(cd), v_buf = even_up(cd > n ? cd : n);
    const int v_len = OP == 0 ? 0 : n;                // g_q in both Hessian transforms
    double * BIG = sm + pl.big, * MID = sm + pl.mid, * THIRD = sm + pl.third;
    auto rg_of = [&](int buf) { return sm + pl.small + buf * (rg_buf + v_buf); };
    auto v_of = [&](int buf) { return sm + pl.small + buf * (rg_buf + v_buf) + rg_buf; };
    auto jv_of = [&](int buf) { return sm + pl.jv + buf * sp.nnz_pad; };
    auto fetch_small = [&](int64_t bb, int buf) {
        double * rgd = rg_of(buf), * vd = v_of(buf);
        const double * rs = r + bb * cd, * vs = in1 + bb * v_len;
        for (int i = threadIdx.x; i < cd; i += blockDim.x)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_addr(rgd + i)), "l"(rs + i) : "memory");
        for (int i = threadIdx.x; i < v_len; i += blockDim.x)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_addr(vd + i)), "l"(vs + i) : "memory");
    };
    const int big_count = OP == 2 ? n * n : cd * cd;                       // H_q / H_r of one geometry
    // The matrices of a geometry come through the TMA unit whenever shape and alignment allow 16-byte granules (else per-thread
    // cp.async, same group discipline): H_q as one bulk copy, the inverse (full symmetric, from the factor kernel's workspace) as
    // one, H_r row by row into its padded DMMA layout (one instruction each from cd different threads).
    const int astride = even_up(n * n);
    const bool bulk_inv = OP != 2;
    const bool bulk_hq = OP == 2 && ((n * n) & 1) == 0 && (reinterpret_cast<uintptr_t>(in2) & 15) == 0;
    const bool bulk_hr = OP == 3 && (cd & 1) == 0 && (reinterpret_cast<uintptr_t>(in2) & 15) == 0;
    const bool bulk_any = bulk_inv || bulk_hq;
    uint32_t big_parity = 0;
    auto fetch_big = [&](int64_t bb) {
        if (OP == 2) {
            if (bulk_hq) {
                if (threadIdx.x == 0) {
                    fence_async_proxy();
                    mbar_expect_tx(&big_mbar, (uint32_t)big_count * 8u);
                    bulk_load(BIG, in2 + bb * (int64_t)big_count, (uint32_t)big_count * 8u, &big_mbar);
                }
            } else block_copy_async(in2 + bb * (int64_t)big_count, BIG, big_count);
        }
        if (OP == 3) {
            double * Ai = THIRD;
            if (threadIdx.x == 0) {
                fence_async_proxy();
                mbar_expect_tx(&big_mbar, (uint32_t)astride * 8u + (bulk_hr ? (uint32_t)big_count * 8u : 0u));
                bulk_load(Ai, ainv + bb * (int64_t)astride, (uint32_t)astride * 8u, &big_mbar);
            }
            if (bulk_hr) {                                                            // DMMA operand: padded rows
                // (the bulk-copy instruction takes warp-uniform operands, so a warp issues its lanes' copies one after the
                // other: the rows are dealt to the first lanes of ALL warps, not to the first cd threads)
                const int nwarps = blockDim.x >> 5, per_warp = (cd + nwarps - 1) / nwarps;
                const int lane = threadIdx.x & 31, row = (threadIdx.x >> 5) * per_warp + lane;
                if (lane < per_warp && row < cd) {
                    fence_async_proxy();
                    bulk_load(BIG + row * ldj, in2 + bb * (int64_t)big_count + row * cd, (uint32_t)cd * 8u, &big_mbar);
                }
            } else block_rows_async(in2 + bb * (int64_t)big_count, BIG, cd, cd, ldj);
        }
        if (OP == 0 && threadIdx.x == 0) {
            fence_async_proxy();
            mbar_expect_tx(&big_mbar, (uint32_t)astride * 8u);
            bulk_load(BIG, ainv + bb * (int64_t)astride, (uint32_t)astride * 8u, &big_mbar);
        }
    };
    auto wait_big = [&]() {
        if (bulk_any) { mbar_wait(&big_mbar, big_parity); big_parity ^= 1u; }
    };

    // prologue: the first geometry's small inputs and J rows; then its successor's small inputs and its own matrices fly
    int cur = 0;
    const int64_t b0 = blockIdx.x;
    if (b0 >= B) return;
    fetch_small(b0, 0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();                                                       // (also publishes the table copy)
    for (int p = sp.nnz + threadIdx.x; p < sp.nnz_pad; p += blockDim.x) { jv_of(0)[p] = 0.0; jv_of(1)[p] = 0.0; }
    eval_J_rows<HAS5>(sp, rg_of(0), jv_of(0), threadIdx.x, blockDim.x);
    // group order (oldest first): Hessian_int2cart waits for its matrix first, the other two for the small inputs first
    if (OP == 2) { fetch_big(b0); cp_async_commit(); }
    if (b0 + gridDim.x < B) fetch_small(b0 + gridDim.x, 1);
    cp_async_commit();
    if (OP != 2) { fetch_big(b0); cp_async_commit(); }

    for (int64_t b = b0; b < B; b += gridDim.x) {
        const int64_t bn = b + gridDim.x, bnn = bn + gridDim.x;
        const bool has_next = bn < B;
        const double * rg = rg_of(cur), * gvec = v_of(cur), * Jv = jv_of(cur);
        SPHASE_BEGIN(sm);
        if (OP == 2) {
            double * Hq = BIG, * U = MID, * X = THIRD;
            cp_async_wait<1>();                                            // H_q of b (the small inputs of bn, fetched last, may still fly)
            wait_big();
            __syncthreads();                                               // + J rows of b (prologue or previous task phase)
            SPHASE(10);
            spmm_JT(sp, Jv, Hq, n, n, U, ldo);                             // U = J^T H_q        (cd x n)
            if (threadIdx.x == 0) bulk_store_read_done();                  // the previous geometry's X has left shared memory
            __syncthreads();
            SPHASE(11);
            if (has_next) fetch_big(bn);                                   // H_q is dead: the next one travels during the rest
            cp_async_commit();
            spmm_J(sp, Jv, U, ldo, cd, X, cd);                             // X = U J            (cd x cd)
            cp_async_wait<1>();                                            // the small inputs of bn (the next H_q may still fly)
            __syncthreads();
            SPHASE(12);
            run_tasks<HAS5>(sp, rg, gvec, 1.0, U /* slots */, has_next ? rg_of(cur ^ 1) : nullptr, jv_of(cur ^ 1));
            __syncthreads();
            SPHASE(13);
            gather_slots(sp, U, X, cd);                                    // X += sum_k g_q[k] K[k]
            if (bnn < B) fetch_small(bnn, cur);                            // (rg / g_q of b are dead after the tasks)
            cp_async_commit();
            double * Xg = out + b * (int64_t)cd * cd;
            const bool bulk = bulk_store_ok(X, Xg, cd * cd);
            if (bulk) fence_async_proxy();
            __syncthreads();
            SPHASE(14);
            // the result leaves through the TMA unit while the next geometry's U phase runs (X is rewritten after it)
            if (bulk) { if (threadIdx.x == 0) bulk_store_issue(X, Xg, cd * cd); }
            else block_store_rows(X, Xg, cd, cd, cd);
            SPHASE(15);
            cur ^= 1;
            continue;                                                      // (the next iteration's first barrier orders the reuse of X)
        }
        if (OP == 0) {
            double * Ai = BIG, * M = MID;
            cp_async_wait<0>();
            wait_big();
            __syncthreads();
            spmm_JT(sp, Jv, Ai, n, n, M, n);            // (J^T inv)[c][i] = M[i][c]: stored transposed below
            __syncthreads();
            if (has_next) fetch_big(bn);
            cp_async_commit();
            // M^T lives as [cd][n]; write out[i][c] with lanes along c
            double * dst = out + b * (int64_t)n * cd;
            for (int i = threadIdx.x >> 5; i < n; i += blockDim.x >> 5)
                for (int c = threadIdx.x & 31; c < cd; c += kLanes) __stcs(dst + (int64_t)i * cd + c, M[c * n + i]);
            if (has_next) eval_J_rows<HAS5>(sp, rg_of(cur ^ 1), jv_of(cur ^ 1), threadIdx.x, blockDim.x);
            __syncthreads();
            if (bnn < B) fetch_small(bnn, cur);
            cp_async_commit();
            cur ^= 1;
            continue;
        }
        // OP 3: H_q = M (H_r - sum_k g_q[k] K[k]) M^T,  M = inv J  (the reference's order of operations: forming
        // inv (J H_c J^T) inv instead loses cond(J J^T) in accuracy, measured 5e-10 on the 30-atom chain)
        {
            double * Hc = BIG, * SL = MID, * Ai = THIRD;
            cp_async_wait<1>();                                            // the small inputs of bn (the matrices may still fly)
            __syncthreads();
            SPHASE(20);
            run_tasks<HAS5>(sp, rg, gvec, -1.0, SL, has_next ? rg_of(cur ^ 1) : nullptr, jv_of(cur ^ 1));
            cp_async_wait<0>();                                            // H_r and the inverse of b
            wait_big();
            __syncthreads();
            SPHASE(21);
            gather_slots(sp, SL, Hc, ldj);                                 // H_c = H_r - sum_k g_q[k] K[k]
            __syncthreads();
            SPHASE(22);
            double * MT = MID;
            spmm_JT(sp, Jv, Ai, n, n, MT, ldh);                            // M^T = J^T inv       (cd x n; inv is symmetric, contiguous rows)
            __syncthreads();
            SPHASE(23);
            double * T2 = THIRD;
            block_dgemm<false, false, false>(cd, n, cd, Hc, ldj, MT, ldh, T2, ldh);      // T2 = H_c M^T   (cd x n)
            __syncthreads();
            SPHASE(24);
            double * Q = BIG;
            // (a symmetric second product - lower tiles mirrored - would need an exactly symmetric H_c: the diagonal 3 x 3
            // blocks of sum_k g_k K[k] come from different Dual passes and are symmetric only to rounding, and checking
            // costs more than half of what the triangular product saves: measured, not kept)
            block_dgemm<true, false, false>(n, n, cd, MT, ldh, T2, ldh, Q, ldh);         // H_q = M T2     (n x n)
            __syncthreads();
            SPHASE(25);
            double * Qg = out + b * (int64_t)n * n;
            if (ldh == n && bulk_store_ok(Q, Qg, n * n)) {
                fence_async_proxy();
                __syncthreads();
                if (threadIdx.x == 0) { bulk_store_issue(Q, Qg, n * n); bulk_store_read_done(); }
            } else block_store_rows(Q, Qg, n, n, ldh);
            __syncthreads();                                               // Q has been read, T2 is dead: the next matrices may land
            SPHASE(26);
            if (bnn < B) fetch_small(bnn, cur);
            cp_async_commit();
            if (has_next) fetch_big(bn);
            cp_async_commit();
            SPHASE(27);
            cur ^= 1;
        }
    }
    cp_async_wait<0>();
    if (threadIdx.x == 0) bulk_store_all_done();
}

// ======================================================================================================
// host side: the compressed tables
// ======================================================================================================
struct SparseTable {
    SparseProgram prog;
    void * blob = nullptr;
    int * status = nullptr;
    double density = 1.0;             // structural nonzeros of J / (intdim * cartdim)
};
std::map<const tchem_ic_table *, SparseTable> & sparse_tables() {
    static std::map<const tchem_ic_table *, SparseTable> m;
    return m;
}
std::mutex & sparse_mutex() { static std::mutex m; return m; }

int pick_ld_dmma(int width) {      // smallest ld >= width with ld = 4 or 12 (mod 16)
    int ld = width;
    while ((ld & 15) != 4 && (ld & 15) != 12) ld++;
    return ld;
}

int type_cost(int type, int natoms) {          // relative cost of one evaluation, for the task order
    if (natoms == 5) return 40;
    if (natoms == 4) return type == TCHEM_OUTOFPLANE || type == TCHEM_SINOOP ? 11 : 12;
    if (natoms == 3) return 6;
    return natoms == 2 ? 2 : 0;
}

// Cost of one warp step of the block kernel's task phase, in units of ~100 cycles as measured on C4 with three warps per
// scheduler (tools/sparse_phases.py prints the cycles of every step): a Dual pass of a primitive's J, and the plain J
// of the terms of a row (the J-row steps also zero their row and add into it: a fixed part per step).
int kdir_step_cost(int type, int natoms) {
    if (natoms == 5) return 100;
    if (natoms == 4) return 30;
    if (natoms == 3) return 21;
    return natoms == 2 ? 12 : 1;
}
int jrow_step_cost(int type, int natoms, int nterms) {
    const int per_term = natoms == 5 ? 60 : natoms == 4 ? 23 : natoms == 3 ? 10 : 8;
    return 12 + per_term * nterms;
}

int get_sparse(const tchem_ic_table * t, SparseTable & out) {
    std::lock_guard<std::mutex> lock(sparse_mutex());
    auto it = sparse_tables().find(t);
    if (it != sparse_tables().end()) { out = it->second; return TCHEM_OK; }
    const int n = t->intdim, cd = t->cartdim;
    std::vector<std::vector<int>> rows(n);
    for (size_t k = 0; k < t->terms.size(); k++) rows[t->terms[k].ic].push_back((int)k);
    // compressed rows: the atoms a row touches, ascending; 3 consecutive Jv positions per atom
    std::vector<uint16_t> nz_ptr(1, 0), nz_col;
    std::vector<std::vector<int>> row_atoms(n);
    std::vector<std::map<int, int>> atom_pos(n);              // atom -> first Jv position in that row
    int nnz_true = 0;
    for (int i = 0; i < n; i++) {
        std::set<int> atoms;
        for (int k : rows[i]) for (int a = 0; a < t->terms[k].natoms; a++) atoms.insert(t->terms[k].atoms[a]);
        for (int a : atoms) {
            atom_pos[i][a] = (int)nz_col.size();
            row_atoms[i].push_back(a);
            for (int c = 0; c < 3; c++) nz_col.push_back((uint16_t)(3 * a + c));
        }
        nnz_true += (int)(nz_col.size() - nz_ptr.back());
        // rows of even length get one padding position (value always zero, column = the row's last): the lanes of a warp take
        // consecutive rows of one type, and an odd distance between their positions keeps their 8-byte accesses off each other's banks
        if (!atoms.empty() && ((nz_col.size() - nz_ptr.back()) & 1) == 0) nz_col.push_back(nz_col.back());
        nz_ptr.push_back((uint16_t)nz_col.size());
    }
    const int nnz = (int)nz_col.size();
    const int zero_pos = nnz;                                  // 3 zeros after the values: target of the padded pair lists
    const int nnz_pad = (nnz + 3 + 1) & ~1;
    std::vector<SpTerm> terms;
    std::vector<int32_t> row_ptr(1, 0);
    std::vector<uint32_t> ktasks;
    int nslots = 0;
    for (int i = 0; i < n; i++) {
        for (int k : rows[i]) {
            const auto & d = t->terms[k];
            SpTerm st;
            std::memset(&st, 0, sizeof(st));
            st.pd.type = d.type; st.pd.natoms = d.natoms;
            for (int a = 0; a < 5; a++) st.pd.atoms[a] = a < d.natoms ? d.atoms[a] : 0;
            st.pd.out = i;
            st.pd.min = d.min;
            st.coef = d.coeff;
            for (int a = 0; a < 5; a++) st.pos[a] = (int16_t)(a < d.natoms ? atom_pos[i][d.atoms[a]] : zero_pos);
            st.kslot = nslots;
            // odd distance between the slot blocks of consecutive terms: the lanes' 8-byte stores spread over all banks
            nslots += (9 * (d.natoms * (d.natoms + 1) / 2)) | 1;
            // Translational invariance: the second-derivative blocks of a row of atoms sum to zero, K(i, 0) = - sum_{j >= 1} K(i, j).
            // The direction passes of atom 0 only produce block (0, 0), so they are not run at all for the 2 - 4 atom types
            // (torsion 9 passes instead of 12, bend 6 of 9, stretch 3 of 6): the gather below builds (0, 0) from the others.
            // Plain torsion: the gradient block of an end atom does not depend on the other end (J_0 is a function of atoms 0, 1, 2
            // only), so K(0, 3) = 0 exactly and K(3, 3) = - K(3, 1) - K(3, 2): the passes of atom 3 are not needed either - 6 of 12.
            // (Not the sine / cosine variants: their J_0 carries a factor cos / sin of the angle, which depends on atom 3.)
            const bool tors0 = d.type == TCHEM_TORSION && d.natoms == 4;
            for (int dir = (d.natoms >= 2 && d.natoms <= 4) ? 3 : 0; dir < (tors0 ? 9 : 3 * d.natoms); dir++)
                ktasks.push_back(((uint32_t)terms.size() << 8) | (uint32_t)dir);
            terms.push_back(st);
        }
        row_ptr.push_back((int32_t)terms.size());
    }
    // heaviest tasks first; within a type direction-major, so a warp runs one formula in lock step
    std::stable_sort(ktasks.begin(), ktasks.end(), [&](uint32_t a, uint32_t b) {
        const PrimDesc & ta = terms[a >> 8].pd, & tb = terms[b >> 8].pd;
        const int ca = type_cost(ta.type, ta.natoms), cb = type_cost(tb.type, tb.natoms);
        if (ca != cb) return ca > cb;
        if (ta.type != tb.type) return ta.type < tb.type;
        return (a & 0xffu) < (b & 0xffu);
    });
    // J rows in evaluation order: grouped by the type of the first term (heaviest first), groups padded to whole warps
    std::vector<int32_t> jrow_order;
    {
        std::vector<int> order(n);
        for (int i = 0; i < n; i++) order[i] = i;
        auto key = [&](int i) {
            if (rows[i].empty()) return std::make_pair(0, 0);
            const auto & d = t->terms[rows[i][0]];
            return std::make_pair(type_cost(d.type, d.natoms) * (int)rows[i].size(), (int)d.type);
        };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) > key(b); });
        for (size_t k = 0; k < order.size(); k++) {
            if (k > 0 && key(order[k]) != key(order[k - 1]))
                while (jrow_order.size() % 32) jrow_order.push_back(-1);
            jrow_order.push_back(order[k]);
        }
        // (padding only pays when the groups are large; many small groups would only lengthen the list)
        if ((int)jrow_order.size() > 2 * n + 64) { jrow_order.assign(order.begin(), order.end()); }
    }
    // warp steps: 32 tasks of one kind each (K tasks of one primitive type; J rows whose first term has one type), and their
    // longest-processing-time assignment to the block kernel's warps
    std::vector<uint32_t> step_tasks;
    std::vector<int> step_cost;
    {
        auto flush_group = [&](const std::vector<uint32_t> & grp, int cost) {
            for (size_t k = 0; k < grp.size(); k += 32) {
                for (size_t u = 0; u < 32; u++) step_tasks.push_back(k + u < grp.size() ? grp[k + u] : 0xffffffffu);
                step_cost.push_back(cost);
            }
        };
        std::vector<uint32_t> grp;
        int gtype = -1, gcost = 0;
        for (uint32_t task : ktasks) {                          // already sorted by cost, type, direction
            const PrimDesc & pd = terms[task >> 8].pd;
            if (pd.type != gtype && !grp.empty()) { flush_group(grp, gcost); grp.clear(); }
            gtype = pd.type; gcost = kdir_step_cost(pd.type, pd.natoms);
            grp.push_back(task);
        }
        if (!grp.empty()) flush_group(grp, gcost);
        grp.clear();
        int prev_cost = -1, prev_type = -1;
        for (int32_t row : jrow_order) {
            if (row < 0) continue;
            const auto & d = t->terms[rows[row][0]];
            const int cost = jrow_step_cost(d.type, d.natoms, (int)rows[row].size());
            if ((cost != prev_cost || (int)d.type != prev_type) && !grp.empty()) { flush_group(grp, std::max(prev_cost, 1)); grp.clear(); }
            prev_cost = cost; prev_type = d.type;
            grp.push_back(0x80000000u | (uint32_t)row);
        }
        if (!grp.empty()) flush_group(grp, std::max(prev_cost, 1));
    }
    const int hess_warps = kHessThreads / 32;
    std::vector<uint16_t> sched_ptr(hess_warps + 1, 0), sched_step;
    {
        std::vector<int> order(step_cost.size());
        for (size_t k = 0; k < order.size(); k++) order[k] = (int)k;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return step_cost[a] > step_cost[b]; });
        std::vector<std::vector<uint16_t>> per_warp(hess_warps);
        std::vector<long> load(hess_warps, 0);
        // two levels: warps w, w + 4, w + 8 share a scheduler (and its FP64 pipe), and a step lasts longer the more its
        // scheduler has to do - so the heaviest remaining step goes to the least loaded SCHEDULER, there to its least loaded warp
        // (warps as independent machines left one scheduler with 15 % more work than the others, and the barrier waits for it)
        long smsp_load[4] = {0, 0, 0, 0};
        for (int k : order) {
            int q = 0;
            for (int c = 1; c < 4; c++) if (smsp_load[c] < smsp_load[q]) q = c;
            int w = q;
            for (int c = q + 4; c < hess_warps; c += 4) if (load[c] < load[w]) w = c;
            per_warp[w].push_back((uint16_t)k);
            load[w] += step_cost[k];
            smsp_load[q] += step_cost[k];
        }
        for (int w = 0; w < hess_warps; w++) {
            sched_ptr[w] = (uint16_t)sched_step.size();
            for (uint16_t k : per_warp[w]) sched_step.push_back(k);
        }
        sched_ptr[hess_warps] = (uint16_t)sched_step.size();
    }
    // compressed column blocks: per atom, the rows that touch it (the atom's three columns share the list)
    const int natoms = t->natoms;
    std::vector<uint16_t> col_ptr(natoms + 1, 0);
    std::vector<uint32_t> col_ent;
    for (int a = 0; a < natoms; a++) {
        col_ptr[a] = (uint16_t)col_ent.size();
        for (int i = 0; i < n; i++) {
            auto f = atom_pos[i].find(a);
            if (f != atom_pos[i].end()) col_ent.push_back(((uint32_t)i << 16) | (uint32_t)f->second);
        }
    }
    col_ptr[natoms] = (uint16_t)col_ent.size();
    // atoms -> warps of the block kernel: a ring atom's row list is three times a hydrogen's, so round robin leaves
    // some warps with twice the work of others
    std::vector<uint16_t> aw_ptr(hess_warps + 1, 0), aw_atom;
    {
        std::vector<int> order(natoms);
        for (int a = 0; a < natoms; a++) order[a] = a;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return col_ptr[a + 1] - col_ptr[a] > col_ptr[b + 1] - col_ptr[b]; });
        std::vector<std::vector<uint16_t>> per_warp(hess_warps);
        std::vector<int> load(hess_warps, 0);
        for (int a : order) {
            const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            per_warp[w].push_back((uint16_t)a);
            load[w] += 2 + col_ptr[a + 1] - col_ptr[a];           // (2: the fixed cost of a column block)
        }
        for (int w = 0; w < hess_warps; w++) {
            aw_ptr[w] = (uint16_t)aw_atom.size();
            for (uint16_t a : per_warp[w]) aw_atom.push_back(a);
        }
        aw_ptr[hess_warps] = (uint16_t)aw_atom.size();
    }
    // J J^T: pairs of rows that share atoms
    std::vector<uint32_t> pair_dest;
    std::vector<std::vector<uint32_t>> pair_lists;
    int maxc = 1;
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            std::vector<uint32_t> common;
            for (int a : row_atoms[i]) {
                auto f = atom_pos[j].find(a);
                if (f != atom_pos[j].end()) common.push_back(((uint32_t)atom_pos[i][a] << 16) | (uint32_t)f->second);
            }
            if (common.empty()) continue;
            pair_dest.push_back((uint32_t)(i * (i + 1) / 2 + j));
            maxc = std::max(maxc, (int)common.size());
            pair_lists.push_back(common);
        }
    maxc = (maxc + 3) & ~3;                                  // the factor kernel reads the lists four entries at a time
    const int npairs = (int)pair_dest.size();
    const int npairs_pad = (npairs + 31) & ~31;
    const uint32_t spare = (uint32_t)(n * (n + 1) / 2 + n);             // the spare element behind the bordered row
    pair_dest.resize(npairs_pad, spare);
    std::vector<uint32_t> pair_ent((size_t)maxc * npairs_pad, ((uint32_t)zero_pos << 16) | (uint32_t)zero_pos);
    for (int p = 0; p < npairs; p++)
        for (size_t e = 0; e < pair_lists[p].size(); e++) pair_ent[e * npairs_pad + p] = pair_lists[p][e];
    // g.K gather: ordered atom pairs (A, B) <- slots of every term that contains both
    std::map<std::pair<int, int>, std::vector<uint32_t>> gk;
    for (const SpTerm & st : terms) {
        const int N = st.pd.natoms;
        for (int i = 0; i < N; i++)
            for (int j = i; j < N; j++) {
                const uint32_t base = (uint32_t)(st.kslot + 9 * (i * N - (i * (i - 1)) / 2 + (j - i)));
                const int A = st.pd.atoms[i], Bt = st.pd.atoms[j];
                const bool tors0 = st.pd.type == TCHEM_TORSION && N == 4;
                if (i == 0 && j == 0 && N >= 2 && N <= 4) {
                    // block (0, 0) = - sum_{j >= 1} K(0, j): negated sources, no slot of its own is ever written
                    for (int jj = 1; jj < (tors0 ? 3 : N); jj++) gk[{A, A}].push_back(((uint32_t)(st.kslot + 9 * jj) << 2) | 2u);
                    continue;
                }
                if (tors0 && i == 0 && j == 3) continue;                       // K(0, 3) = 0
                if (tors0 && i == 3 && j == 3) {
                    // K(3, 3) = - K(3, 1) - K(3, 2) = - K(1, 3)^T - K(2, 3)^T: negated and transposed sources
                    for (int ii = 1; ii < 3; ii++)
                        gk[{A, A}].push_back(((uint32_t)(st.kslot + 9 * (ii * N - (ii * (ii - 1)) / 2 + (3 - ii))) << 2) | 3u);
                    continue;
                }
                gk[{A, Bt}].push_back(base << 2);
                if (i != j) gk[{Bt, A}].push_back((base << 2) | 1u);
            }
    }
    std::vector<uint32_t> gk_blk, gk_ptr(1, 0), gk_src;
    {
        // heaviest blocks first: the gather's first round of threads takes them, the light ones fill the later rounds
        std::vector<const std::pair<const std::pair<int, int>, std::vector<uint32_t>> *> order;
        for (const auto & kv : gk) order.push_back(&kv);
        std::stable_sort(order.begin(), order.end(), [](auto a, auto b) { return a->second.size() > b->second.size(); });
        for (auto kv : order) {
            gk_blk.push_back(((uint32_t)kv->first.first << 16) | (uint32_t)kv->first.second);
            for (uint32_t s : kv->second) gk_src.push_back(s);
            gk_ptr.push_back((uint32_t)gk_src.size());
        }
    }
    // blob
    auto a16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    size_t off = 0;
    auto place = [&](size_t bytes) { const size_t o = off; off += a16(bytes); return o; };
    // order: [pair lists] [common: terms, row structure] [block-kernel tables]; the factor kernel stages the first two
    // parts in shared memory, the block kernel the last two
    const size_t o_pdest = place(pair_dest.size() * 4), o_pent = place(pair_ent.size() * 4);
    const size_t common_begin = off;
    const size_t o_terms = place(terms.size() * sizeof(SpTerm)), o_rowptr = place(row_ptr.size() * 4), o_jrow = place(jrow_order.size() * 4),
                 o_nzptr = place(nz_ptr.size() * 2), o_nzcol = place(nz_col.size() * 2 + 2);
    const size_t hess_begin = off;
    const size_t o_ktasks = place(step_tasks.size() * 4), o_schedptr = place(sched_ptr.size() * 2), o_schedstep = place(sched_step.size() * 2 + 2), o_colptr = place(col_ptr.size() * 2), o_colent = place(col_ent.size() * 4 + 4),
                 o_awptr = place(aw_ptr.size() * 2), o_awatom = place(aw_atom.size() * 2 + 2),
                 o_gkblk = place(gk_blk.size() * 4 + 4), o_gkptr = place(gk_ptr.size() * 4), o_gksrc = place(gk_src.size() * 4 + 4);
    const size_t total = off + 16;
    std::vector<unsigned char> blob(total, 0);
    std::memcpy(blob.data() + o_terms, terms.data(), terms.size() * sizeof(SpTerm));
    std::memcpy(blob.data() + o_rowptr, row_ptr.data(), row_ptr.size() * 4);
    std::memcpy(blob.data() + o_jrow, jrow_order.data(), jrow_order.size() * 4);
    std::memcpy(blob.data() + o_ktasks, step_tasks.data(), step_tasks.size() * 4);
    std::memcpy(blob.data() + o_schedptr, sched_ptr.data(), sched_ptr.size() * 2);
    std::memcpy(blob.data() + o_schedstep, sched_step.data(), sched_step.size() * 2);
    std::memcpy(blob.data() + o_nzptr, nz_ptr.data(), nz_ptr.size() * 2);
    std::memcpy(blob.data() + o_nzcol, nz_col.data(), nz_col.size() * 2);
    std::memcpy(blob.data() + o_colptr, col_ptr.data(), col_ptr.size() * 2);
    std::memcpy(blob.data() + o_colent, col_ent.data(), col_ent.size() * 4);
    std::memcpy(blob.data() + o_awptr, aw_ptr.data(), aw_ptr.size() * 2);
    std::memcpy(blob.data() + o_awatom, aw_atom.data(), aw_atom.size() * 2);
    std::memcpy(blob.data() + o_gkblk, gk_blk.data(), gk_blk.size() * 4);
    std::memcpy(blob.data() + o_gkptr, gk_ptr.data(), gk_ptr.size() * 4);
    std::memcpy(blob.data() + o_gksrc, gk_src.data(), gk_src.size() * 4);
    std::memcpy(blob.data() + o_pdest, pair_dest.data(), pair_dest.size() * 4);
    std::memcpy(blob.data() + o_pent, pair_ent.data(), pair_ent.size() * 4);
    void * dev = nullptr;
    TC_CUDA(cudaMalloc(&dev, total));
    cudaError_t e = cudaMemcpy(dev, blob.data(), total, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dev); return set_error(TCHEM_ECUDA, cudaGetErrorString(e)); }
    SparseTable st;
    st.blob = dev;
    TC_CUDA(cudaMalloc(reinterpret_cast<void **>(&st.status), sizeof(int)));
    TC_CUDA(cudaMemset(st.status, 0, sizeof(int)));
    unsigned char * d = static_cast<unsigned char *>(dev);
    SparseProgram & p = st.prog;
    p.terms = reinterpret_cast<const SpTerm *>(d + o_terms);
    p.row_ptr = reinterpret_cast<const int32_t *>(d + o_rowptr);
    p.jrow_order = reinterpret_cast<const int32_t *>(d + o_jrow);
    p.ktasks = reinterpret_cast<const uint32_t *>(d + o_ktasks);
    p.sched_ptr = reinterpret_cast<const uint16_t *>(d + o_schedptr);
    p.sched_step = reinterpret_cast<const uint16_t *>(d + o_schedstep);
    p.nz_ptr = reinterpret_cast<const uint16_t *>(d + o_nzptr);
    p.nz_col = reinterpret_cast<const uint16_t *>(d + o_nzcol);
    p.col_ptr = reinterpret_cast<const uint16_t *>(d + o_colptr);
    p.col_ent = reinterpret_cast<const uint32_t *>(d + o_colent);
    p.aw_ptr = reinterpret_cast<const uint16_t *>(d + o_awptr);
    p.aw_atom = reinterpret_cast<const uint16_t *>(d + o_awatom);
    p.pair_dest = reinterpret_cast<const uint32_t *>(d + o_pdest);
    p.pair_ent = reinterpret_cast<const uint32_t *>(d + o_pent);
    p.gk_blk = reinterpret_cast<const uint32_t *>(d + o_gkblk);
    p.gk_ptr = reinterpret_cast<const uint32_t *>(d + o_gkptr);
    p.gk_src = reinterpret_cast<const uint32_t *>(d + o_gksrc);
    p.nterms = (int)terms.size(); p.njrow = (int)jrow_order.size(); p.nktasks = (int)ktasks.size();
    p.nnz = nnz; p.nnz_pad = nnz_pad; p.npairs_pad = npairs_pad; p.maxc = maxc;
    p.ngblk = (int)gk_blk.size(); p.ngsrc = (int)gk_src.size(); p.nslots = nslots;
    p.natoms = natoms;
    p.intdim = n; p.cartdim = cd; p.ldh = pick_ld_dmma(n);
    p.factor_tab_bytes = (int32_t)hess_begin;
    p.hess_tab_offset = (int32_t)common_begin;
    p.blob_bytes = (int32_t)off;
    {
        // Jv in the tail of the packed matrix (factor_kernel): possible when the pairs that land in that tail fit the registers
        const int atot = n * (n + 1) / 2 + n + 2, lo_end = atot - nnz_pad;
        int lo = 0;
        while (lo < npairs && (int)pair_dest[lo] < lo_end) lo++;
        p.jv_alias = (lo_end >= 0 && npairs_pad - lo <= 32 * kMaxHiPairs) ? 1 : 0;
        p.npairs_lo = p.jv_alias ? lo : npairs_pad;
    }
    p.tab_in_smem = 0;
    st.density = (double)nnz_true / ((double)n * cd);
    sparse_tables()[t] = st;
    out = st;
    return TCHEM_OK;
}

template <int OP> const void * hess_fn(bool has5, bool tabs) {
    return tabs ? (has5 ? reinterpret_cast<const void *>(&hess_kernel<OP, true, true>) : reinterpret_cast<const void *>(&hess_kernel<OP, false, true>))
                : (has5 ? reinterpret_cast<const void *>(&hess_kernel<OP, true, false>) : reinterpret_cast<const void *>(&hess_kernel<OP, false, false>));
}
template <int OPF> const void * factor_fn(bool has5) {
    return has5 ? reinterpret_cast<const void *>(&factor_kernel<OPF, true>) : reinterpret_cast<const void *>(&factor_kernel<OPF, false>);
}

} // namespace

#ifdef TCHEM_SPARSE_PHASES
extern "C" __attribute__((visibility("default"))) int tchem_debug_sparse_phases(unsigned long long * out32, int reset) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out32, g_sparse_phase, sizeof(g_sparse_phase)) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_sparse_phase, z, sizeof(z)); }
    return 0;
}
extern "C" __attribute__((visibility("default"))) int tchem_debug_sparse_steps(unsigned long long * out96, int reset) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out96, g_sparse_step, sizeof(g_sparse_step)) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[96] = {0}; cudaMemcpyToSymbol(g_sparse_step, z, sizeof(z)); }
    return 0;
}
#endif

void release_sparse_program(const tchem_ic_table * t) {
    std::lock_guard<std::mutex> lock(sparse_mutex());
    auto it = sparse_tables().find(t);
    if (it != sparse_tables().end()) { cudaFree(it->second.blob); cudaFree(it->second.status); sparse_tables().erase(it); }
}

int sparse_status(const tchem_ic_table * t) {
    std::lock_guard<std::mutex> lock(sparse_mutex());
    auto it = sparse_tables().find(t);
    if (it == sparse_tables().end()) return 0;
    int s = 0;
    cudaMemcpy(&s, it->second.status, sizeof(int), cudaMemcpyDeviceToHost);
    if (s) cudaMemset(it->second.status, 0, sizeof(int));
    return s;
}

// op: 0 gradient_cart2int_matrix, 1 gradient_cart2int, 2 Hessian_int2cart, 3 Hessian_cart2int (as ic_dense.cu).
// Returns -1 when this table / operation stays on the all-DMMA kernels of ic_dense.cu: J not sparse enough, more than
// 127 internal or 128 Cartesian coordinates, or an operand plan beyond the shared memory of one SM.
int launch_sparse(int op, const tchem_ic_table * t, const double * r, const double * in1, const double * in2, int64_t B,
                  double * out, cudaStream_t stream) {
    static const int mode = [] { const char * e = getenv("TCHEM_DENSE_DMMA"); return e ? atoi(e) : 0; }();   // 1: always the DMMA kernels
    if (mode == 1) return -1;
    const int n = t->intdim, cd = t->cartdim;
    if (n + 1 > 32 * kMaxSlots || cd > 32 * kMaxSlots || n > cd && op != 2) return -1;
    SparseTable st;
    int rc = get_sparse(t, st);
    if (rc != TCHEM_OK) return rc;
    // (any density: on the small molecules whose J is dense the compressed lists simply hold every element, and the results
    // stay bit-reproducible; TCHEM_SPARSE_MAX_DENSITY sends denser tables to the DMMA kernels)
    static const double max_density = [] { const char * e = getenv("TCHEM_SPARSE_MAX_DENSITY"); return e ? atof(e) : 1.01; }();
    if (st.density > max_density || st.prog.nnz_pad > 65000 || st.prog.nslots > (1 << 30)) return -1;
    // ---- kernel 1 (cart -> int): factor / solve / inverse, one warp per geometry ---------------------------------
    static const int alias_env = [] { const char * e = getenv("TCHEM_FACTOR_ALIAS"); return e ? atoi(e) : 1; }();
    if (!alias_env) { st.prog.jv_alias = 0; st.prog.npairs_lo = st.prog.npairs_pad; }
    const size_t fwarp = (size_t)even_up(factor_smem_doubles(n, cd, st.prog.nnz_pad, st.prog.jv_alias)) * sizeof(double);
    const size_t ftab = (size_t)even_up((st.prog.factor_tab_bytes + 7) / 8) * sizeof(double);
    int fwarps = 0;
    if (op != 2) {
        static const int w_env = [] { const char * e = getenv("TCHEM_FACTOR_WARPS"); return e ? atoi(e) : 6; }();
        fwarps = std::min(std::max(w_env, 1), 6);
        while (fwarps > 0 && ftab + fwarps * fwarp + 1024 > (size_t)t->max_smem_optin) fwarps--;
        if (fwarps < 1) return -1;
    }
    const size_t fsmem = ftab + fwarps * fwarp;
    // ---- kernel 2: one block per geometry ---------------------------------------------------------------------------
    size_t hsmem = 0;
    if (op != 1) {
        HessPlan pl = hess_plan(op, n, cd, st.prog.ldh, st.prog.nnz_pad, st.prog.nslots, (st.prog.blob_bytes - st.prog.hess_tab_offset + 7) / 8);
        st.prog.tab_in_smem = 1;
        static const int tabs_env = [] { const char * e = getenv("TCHEM_SPARSE_TABS"); return e ? atoi(e) : 1; }();   // 0: tables stay in global memory
        if (!tabs_env || (size_t)pl.total * sizeof(double) + 1024 > (size_t)t->max_smem_optin) {
            pl = hess_plan(op, n, cd, st.prog.ldh, st.prog.nnz_pad, st.prog.nslots, 0);
            st.prog.tab_in_smem = 0;
        }
        hsmem = (size_t)pl.total * sizeof(double);
        if (hsmem + 1024 > (size_t)t->max_smem_optin) return -1;
    }
    auto launch_factor = [&](int opf, const double * rr, const double * gr, int64_t nb, double * gq, double * ainv) -> int {
        const void * fn = opf == 0 ? factor_fn<0>(t->has5) : factor_fn<1>(t->has5);
        { const int rc_ = ensure_max_dynamic_smem(fn, t->device, t->max_smem_optin); if (rc_ != TCHEM_OK) return rc_; }
        const int64_t grid = std::min<int64_t>((nb + fwarps - 1) / fwarps, (int64_t)t->sm_count);
        // the geometry counter belongs to THIS launch (stream-ordered allocation from the retained pool: microseconds), so that
        // launches on the same table from different streams cannot take each other's geometries
        unsigned long long * counter = nullptr;
        keep_async_pool(t->device);
        TC_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&counter), sizeof(unsigned long long), stream));
        cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
        void * args[] = {(void *)&st.prog, (void *)&rr, (void *)&gr, (void *)&nb, (void *)&gq, (void *)&ainv, (void *)&st.status, (void *)&counter};
        if (e == cudaSuccess) e = cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(32 * fwarps), args, fsmem, stream);
        cudaFreeAsync(counter, stream);
        if (e != cudaSuccess) return set_error(TCHEM_ECUDA, std::string("factor_kernel launch: ") + cudaGetErrorString(e));
        count_launch(opf == 0 ? "factor_kernel<solve>" : "factor_kernel<inverse>");
        return TCHEM_OK;
    };
    auto launch_hess = [&](int hop, const double * rr, const double * v1, const double * m2, const double * ainv, int64_t nb, double * o) -> int {
        const bool tabs = st.prog.tab_in_smem != 0;
        const void * fn = hop == 0 ? hess_fn<0>(t->has5, tabs) : hop == 2 ? hess_fn<2>(t->has5, tabs) : hess_fn<3>(t->has5, tabs);
        { const int rc_ = ensure_max_dynamic_smem(fn, t->device, t->max_smem_optin); if (rc_ != TCHEM_OK) return rc_; }
        const int64_t grid = std::min<int64_t>(nb, (int64_t)t->sm_count);
        void * args[] = {(void *)&st.prog, (void *)&rr, (void *)&v1, (void *)&m2, (void *)&ainv, (void *)&nb, (void *)&o};
        TC_CUDA(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(kHessThreads), args, hsmem, stream));
        count_launch(hop == 0 ? "hess_kernel<0:gradient_cart2int_matrix>" : hop == 2 ? "hess_kernel<2:Hessian_int2cart>" : "hess_kernel<3:Hessian_cart2int>");
        return TCHEM_OK;
    };
    if (op == 1) return launch_factor(0, r, in1, B, out, nullptr);
    if (op == 2) return launch_hess(2, r, in1, in2, nullptr, B, out);
    // matrix / Hessian cart -> int: the inverses (and g_q) of a slice of the batch go through a stream-ordered workspace
    // slices of at most 32 768 geometries, equal in size (100 000 -> 4 x 25 000: no short last slice whose factor kernel would run two
    // nearly empty rounds; every slice boundary costs the factor kernel about one geometry time of tail)
    static const int64_t max_slice = [] { const char * e = getenv("TCHEM_CART2INT_SLICE"); return (int64_t)std::max(1, e ? atoi(e) : 32768); }();
    const int64_t nslices = (B + max_slice - 1) / max_slice;
    const int64_t slice = (B + nslices - 1) / nslices;
    double * ws = nullptr;
    const size_t astride = (size_t)even_up(n * n);              // full symmetric inverse per geometry (+ g_q)
    const size_t ws_doubles = (size_t)slice * (astride + n);
    keep_async_pool(t->device);
    if (cudaMallocAsync(reinterpret_cast<void **>(&ws), ws_doubles * sizeof(double), stream) != cudaSuccess) {
        cudaGetLastError();
        return set_error(TCHEM_ENOMEM, "tchem::IC::IntCoordSet: cannot allocate the workspace of the cart2int transforms");
    }
    double * ws_ainv = ws, * ws_gq = ws + (size_t)slice * astride;
    rc = TCHEM_OK;
    for (int64_t b0 = 0; b0 < B && rc == TCHEM_OK; b0 += slice) {
        const int64_t nb = std::min<int64_t>(slice, B - b0);
        const double * rr = r + b0 * cd;
        if (op == 0) {
            rc = launch_factor(1, rr, nullptr, nb, nullptr, ws_ainv);
            if (rc == TCHEM_OK) rc = launch_hess(0, rr, nullptr, nullptr, ws_ainv, nb, out + b0 * (int64_t)n * cd);
        } else {
            rc = launch_factor(1, rr, in1 + b0 * cd, nb, ws_gq, ws_ainv);
            if (rc == TCHEM_OK) rc = launch_hess(3, rr, ws_gq, in2 + b0 * (int64_t)cd * cd, ws_ainv, nb, out + b0 * (int64_t)n * n);
        }
    }
    cudaFreeAsync(ws, stream);
    return rc;
}

} // namespace tchem_b200
