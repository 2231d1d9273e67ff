// SAPSet value + Jacobian, stand-alone and chained behind the internal coordinates.
// Replaces, batched: SAPSet::operator() and SAPSet::Jacobian_(xs, J)
// (reference source/polynomial/SAPSet.cpp:134-144, :178-201; per-term arithmetic
// source/polynomial/SAP.cpp:93-99 and :147-175).
//
// Same tile scheme as the q/J/K kernel: one lane per geometry evaluates every monomial and
// its <= order distinct partial derivatives in registers, parks them in the per-warp compact
// buffer and the dense y[B][nterms] / J[B][nterms][sumdim] come out through gather_store.
#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

#include "ic_eval.cuh"
#include "ic_table.h"

#include "sap_table.h"

namespace tchem_b200 {

int launch_jit_chain(const tchem_sap_table * st, const tchem_ic_table * t, const double * r, int64_t B,
                     double * q, double * y, double * J, cudaStream_t stream, bool cart = false);
void release_jit_sap(const tchem_sap_table * t);
int launch_jit_sap_hess(const tchem_sap_table * st, const double * x, int64_t B, double * K, cudaStream_t stream);
extern thread_local int64_t g_batch_hint;

namespace {

// x[col]^p for every power a monomial of the set needs, per lane: XP[(pow_off[col] + p) * kPad + lane]
__device__ __forceinline__ void build_powers(const int32_t * pow_off, int sumdim, const double * X, double * XP, int lane) {
    for (int col = 0; col < sumdim; col++) {
        const int e0 = pow_off[col], e1 = pow_off[col + 1];
        const double x = X[col * kPad + lane];
        double v = 1.0;
        for (int e = e0; e < e1; e++) { XP[e * kPad + lane] = v; v *= x; }
    }
}

// one monomial with U distinct coordinates, fully unrolled: every factor is read once from the
// power table, value = prod x^p (SAP.cpp:93-99), d/dx_u = p x_u^(p-1) prod_{v != u} x_v^(p_v)
// (SAP.cpp:147-175)
template <int U>
__device__ __forceinline__ void eval_sap_term(const SapUniq * uq, const double * XP, double * Pv, double * Pg, bool want_J) {
    double xp[U > 0 ? U : 1], xm[U > 0 ? U : 1];
#pragma unroll
    for (int u = 0; u < U; u++) {
        xp[u] = XP[uq[u].e_pow * kPad];
        xm[u] = (double)uq[u].power * XP[uq[u].e_powm1 * kPad];
    }
    double v = 1.0;
#pragma unroll
    for (int u = 0; u < U; u++) v *= xp[u];
    Pv[0] = v;
    if (!want_J) return;
#pragma unroll
    for (int u = 0; u < U; u++) {
        double g = xm[u];
#pragma unroll
        for (int w = 0; w < U; w++)
            if (w != u) g *= xp[w];
        Pg[u * kPad] = g;
    }
}

// monomials [t0, t1) for this lane's geometry; XP and P have the lane folded in
__device__ __forceinline__ void eval_sap_terms(const SapTerm * terms, const SapUniq * uniqs, int t0, int t1,
                                               const double * XP, double * P, int lane, bool want_J) {
    const int first = terms[t0].ubegin;
    XP += lane;
    P += lane;
    for (int t = t0; t < t1; t++) {
        const SapTerm term = terms[t];
        const SapUniq * uq = uniqs + term.ubegin;
        double * Pv = P + (t - t0) * kPad, * Pg = P + ((t1 - t0) + (term.ubegin - first)) * kPad;
        switch (term.uend - term.ubegin) {
        case 0: eval_sap_term<0>(uq, XP, Pv, Pg, want_J); break;
        case 1: eval_sap_term<1>(uq, XP, Pv, Pg, want_J); break;
        case 2: eval_sap_term<2>(uq, XP, Pv, Pg, want_J); break;
        case 3: eval_sap_term<3>(uq, XP, Pv, Pg, want_J); break;
        case 4: eval_sap_term<4>(uq, XP, Pv, Pg, want_J); break;
        default: {
            double v = 1.0;
            for (int u = term.ubegin; u < term.uend; u++) v *= XP[uniqs[u].e_pow * kPad];
            Pv[0] = v;
            if (!want_J) break;
            for (int u = term.ubegin; u < term.uend; u++) {
                double g = (double)uniqs[u].power * XP[uniqs[u].e_powm1 * kPad];
                for (int w = term.ubegin; w < term.uend; w++)
                    if (w != u) g *= XP[uniqs[w].e_pow * kPad];
                Pg[(u - term.ubegin) * kPad] = g;
            }
        }
        }
    }
}

// CHAINED: x = q(r) evaluated with the compute_IC_J formulas from the IC program (value-mode
// chunking, one compact entry per primitive); otherwise x is read from global memory.
// dynamic shared memory: [SapTerm][SapUniq][SAP ChunkDesc][pow_off] and, chained, [PrimDesc][IC ChunkDesc]
// [q map][q sources]; then per warp: R | PI | X | XP | P
__device__ __forceinline__ size_t sap_align16(size_t x) { return (x + 15) & ~size_t(15); }

template <class T> __device__ __forceinline__ T * stage_table(unsigned char *& cursor, const T * src, int n) {
    T * dst = reinterpret_cast<T *>(cursor);
    const int words = n * (int)(sizeof(T) / 4);
    const int32_t * s = reinterpret_cast<const int32_t *>(src);
    int32_t * d = reinterpret_cast<int32_t *>(dst);
    for (int i = threadIdx.x; i < words; i += blockDim.x) d[i] = s[i];
    cursor += sap_align16(sizeof(T) * n);
    return dst;
}

template <bool CHAINED, bool HAS5>
__global__ void __launch_bounds__(128)
sap_eval_kernel(const SapProgram sp, const Program ic, const double * __restrict__ in, const int64_t B,
                double * __restrict__ q_out, double * __restrict__ y, double * __restrict__ J) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char * cursor = smem_raw;
    const SapTerm * terms = stage_table(cursor, sp.terms, sp.nterms);
    const SapUniq * uniqs = stage_table(cursor, sp.uniqs, sp.nuniqs);
    const ChunkDesc * schunks = stage_table(cursor, sp.chunks, sp.nchunks);
    const int32_t * pow_off = stage_table(cursor, sp.pow_off, sp.sumdim + 1);
    const PrimDesc * prims = nullptr;
    const ChunkDesc * ichunks = nullptr;
    const uint32_t * qmap = nullptr;
    const Src * qsrcs = nullptr;
    if (CHAINED) {
        prims = stage_table(cursor, ic.prims, ic.nprims);
        ichunks = stage_table(cursor, ic.chunks, ic.nchunks);
        qmap = stage_table(cursor, ic.map, ic.nmap);
        qsrcs = stage_table(cursor, ic.qsrcs, ic.nqsrcs);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int sumdim = sp.sumdim;
    const int cartdim = CHAINED ? ic.cartdim : 0, iccap = CHAINED ? ic.cap : 0;
    const size_t per_warp = (size_t)(cartdim + iccap + sumdim + sp.npow + sp.cap) * kPad;
    double * R = reinterpret_cast<double *>(cursor) + warp * per_warp;
    double * PI = R + (size_t)cartdim * kPad;
    double * X = PI + (size_t)iccap * kPad;
    double * XP = X + (size_t)sumdim * kPad;
    double * P = XP + (size_t)sp.npow * kPad;
    const int64_t ntiles = (B + kLanes - 1) / kLanes;
    const int64_t tile_step = (int64_t)gridDim.x * nwarps;
    double * stage_dst = CHAINED ? R : X;
    const int stage_dim = CHAINED ? cartdim : sumdim;

    int64_t tile = (int64_t)blockIdx.x * nwarps + warp;
    if (tile < ntiles) stage_issue(in, tile * kLanes, (int)min((int64_t)kLanes, B - tile * kLanes), stage_dim, stage_dst, lane);
    for (; tile < ntiles; tile += tile_step) {
        const int64_t b0 = tile * kLanes;
        const int ntile = (int)min((int64_t)kLanes, B - b0);
        stage_finish(ntile, stage_dim, stage_dst, lane);
        if (CHAINED) {
            for (int c = 0; c < ic.nchunks; c++) {
                const ChunkDesc & ch = ichunks[c];
                for (int p = ch.prim_begin; p < ch.prim_end; p++)
                    eval_primitive<MODE_QPATH, HAS5>(prims[p], TileLoader{R, lane}, CompactSink(PI, prims[p].out, lane));
                // every lane combines its own column: x[row] = sum coef * q_prim (IntCoord.cpp:66-74)
                for (int row = 0; row < ch.nrows; row++) {
                    const uint32_t w = qmap[ch.qmap + row];
                    const int cnt = (int)(w >> kMapCountShift);
                    const uint32_t off = w & kMapOffsetMask;
                    double x = 0.0;
                    for (int j = 0; j < cnt; j++) x = fma(qsrcs[off + j].coef, PI[qsrcs[off + j].entry * kPad + lane], x);
                    double * xp = X + (ch.row0 + row) * kPad + lane;
                    *xp = ch.accumulate ? *xp + x : x;
                }
            }
            __syncwarp();
            // R is dead: the next tile's coordinates can start moving now
            if (tile + tile_step < ntiles) {
                const int64_t nb0 = (tile + tile_step) * kLanes;
                stage_issue(in, nb0, (int)min((int64_t)kLanes, B - nb0), stage_dim, stage_dst, lane);
            }
            if (q_out != nullptr) store_rows(X, 0, sumdim, q_out + b0 * sumdim, sumdim, ntile, false, lane);
        }
        build_powers(pow_off, sumdim, X, XP, lane);
        if (!CHAINED) {
            __syncwarp();
            if (tile + tile_step < ntiles) {
                const int64_t nb0 = (tile + tile_step) * kLanes;
                stage_issue(in, nb0, (int)min((int64_t)kLanes, B - nb0), stage_dim, stage_dst, lane);
            }
        }
        for (int c = 0; c < sp.nchunks; c++) {
            const ChunkDesc & ch = schunks[c];
            eval_sap_terms(terms, uniqs, ch.prim_begin, ch.prim_end, XP, P, lane, J != nullptr);
            __syncwarp();
            // monomial values sit in compact entries [0, nrows): written (geometry, term)-flattened
            if (y != nullptr) store_rows(P, 0, ch.nrows, y + b0 * sp.nterms + ch.row0, sp.nterms, ntile, false, lane);
            if (J != nullptr)
                gather_store(ch.jspan, sp.emap, sp.psrcs, sp.coefs, sp.runs, P, J + (b0 * sp.nterms + ch.row0) * sumdim,
                             (int64_t)sp.nterms * sumdim, ntile, false, lane);
            __syncwarp();
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Second-order Jacobian of the set, concatenated and symmetric: K[B][nterms][sumdim][sumdim]
// = SAPSet::Jacobian2nd_(xs, K) (source/polynomial/SAPSet.cpp:243-276), per monomial
// SAP::Hessian_ (source/polynomial/SAP.cpp:226-263): diagonal p (p-1) x^(p-2) prod_{others} x_v^p_v
// (0 when p < 2), off-diagonal p_u p_v x_u^(p_u-1) x_v^(p_v-1) prod_{others}, the other factors
// multiplied in ascending (irreducible, index) order like the reference; upper triangle mirrored.
//
// Same warp tile: lane = geometry. The output is cut into chunks of consecutive rows of the
// [nterms * sumdim][sumdim] matrix; a chunk is staged dense [element][lane] (structural zeros
// included - the result is API-mandated dense) and written out geometry after geometry.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
sap_hess_kernel(const SapProgram sp, const double * __restrict__ in, const int64_t B, double * __restrict__ K, const int rows_per) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char * cursor = smem_raw;
    const SapTerm * terms = stage_table(cursor, sp.terms, sp.nterms);
    const SapUniq * uniqs = stage_table(cursor, sp.uniqs, sp.nuniqs);
    const int32_t * pow_off = stage_table(cursor, sp.pow_off, sp.sumdim + 1);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int sd = sp.sumdim, nrows = sp.nterms * sd;
    const size_t per_warp = (size_t)(sd + sp.npow + rows_per * sd) * kPad;
    double * X = reinterpret_cast<double *>(cursor) + warp * per_warp;
    double * XP = X + (size_t)sd * kPad;
    double * D = XP + (size_t)sp.npow * kPad;
    const int64_t ntiles = (B + kLanes - 1) / kLanes, tile_step = (int64_t)gridDim.x * nwarps;
    const int64_t gstride = (int64_t)nrows * sd;
    const bool aligned = gstride % 2 == 0 && (reinterpret_cast<uint64_t>(K) & 15) == 0;

    int64_t tile = (int64_t)blockIdx.x * nwarps + warp;
    if (tile < ntiles) stage_issue(in, tile * kLanes, (int)min((int64_t)kLanes, B - tile * kLanes), sd, X, lane);
    for (; tile < ntiles; tile += tile_step) {
        const int64_t b0 = tile * kLanes;
        const int ntile = (int)min((int64_t)kLanes, B - b0);
        stage_finish(ntile, sd, X, lane);
        build_powers(pow_off, sd, X, XP, lane);
        __syncwarp();
        if (tile + tile_step < ntiles) {
            const int64_t nb0 = (tile + tile_step) * kLanes;
            stage_issue(in, nb0, (int)min((int64_t)kLanes, B - nb0), sd, X, lane);
        }
        const double * xp = XP + lane;
        for (int r0 = 0; r0 < nrows; r0 += rows_per) {
            const int r1 = min(nrows, r0 + rows_per), span = (r1 - r0) * sd;
            double * d = D + lane;
            for (int e = 0; e < span; e++) d[e * kPad] = 0.0;
            const int t0 = r0 / sd, t1 = (r1 - 1) / sd;
            for (int t = t0; t <= t1; t++) {
                const SapTerm term = terms[t];
                const int rbase = t * sd - r0;                      // row of coordinate c in the chunk: rbase + c
                for (int a = term.ubegin; a < term.uend; a++) {
                    const SapUniq ua = uniqs[a];
                    const int ra = rbase + ua.col;
                    const bool a_in = ra >= 0 && ra < r1 - r0;
                    if (ua.power >= 2 && a_in) {
                        double v = (double)(ua.power * (ua.power - 1)) * xp[(ua.e_pow - 2) * kPad];
                        for (int w = term.ubegin; w < term.uend; w++)
                            if (w != a) v *= xp[uniqs[w].e_pow * kPad];
                        d[(ra * sd + ua.col) * kPad] = v;
                    }
                    for (int b = a + 1; b < term.uend; b++) {
                        const SapUniq ub = uniqs[b];
                        const int rb = rbase + ub.col;
                        const bool b_in = rb >= 0 && rb < r1 - r0;
                        if (!a_in && !b_in) continue;
                        double v = (double)(ua.power * ub.power) * xp[ua.e_powm1 * kPad] * xp[ub.e_powm1 * kPad];
                        for (int w = term.ubegin; w < term.uend; w++)
                            if (w != a && w != b) v *= xp[uniqs[w].e_pow * kPad];
                        if (a_in) d[(ra * sd + ub.col) * kPad] = v;
                        if (b_in) d[(rb * sd + ua.col) * kPad] = v;         // SAPSet.cpp:271-274
                    }
                }
            }
            __syncwarp();
            // geometry after geometry, lanes along the chunk's elements (pairs when everything is 16-byte aligned)
            double * out = K + b0 * gstride + (int64_t)r0 * sd;
            if (aligned && span % 2 == 0 && (r0 * sd) % 2 == 0) {
                const int pairs = span >> 1, total = ntile * pairs;
                int g = lane / pairs, p = lane - g * pairs;
                const int dg = kLanes / pairs, dp = kLanes - dg * pairs;
                for (int f = lane; f < total; f += kLanes) {
                    const double * s0 = D + (2 * p) * kPad + g;
                    *reinterpret_cast<double2 *>(out + (int64_t)g * gstride + 2 * p) = make_double2(s0[0], s0[kPad]);
                    g += dg; p += dp;
                    if (p >= pairs) { p -= pairs; g++; }
                }
            } else {
                const int total = ntile * span;
                int g = lane / span, k = lane - g * span;
                const int dg = kLanes / span, dk = kLanes - dg * span;
                for (int f = lane; f < total; f += kLanes) {
                    out[(int64_t)g * gstride + k] = D[k * kPad + g];
                    g += dg; k += dk;
                    if (k >= span) { k -= span; g++; }
                }
            }
            __syncwarp();
        }
    }
}

int launch_sap_hess(const tchem_sap_table * st, const double * x, int64_t B, double * K, cudaStream_t stream) {
    {
        // large batches: the kernel specialised to this SAPSet (jit.cu); -1 = not eligible, interpret
        const int rc = launch_jit_sap_hess(st, x, B, K, stream);
        if (rc != -1) return rc;
    }
    const int sd = st->sumdim, nrows = st->nterms * sd;
    // rows per chunk: ~96 staged elements, whole sectors (4-row multiples when sumdim is odd) when possible
    int rows_per = std::max(1, 96 / sd);
    if (rows_per >= 4) rows_per &= ~3;
    rows_per = std::min(rows_per, nrows);
    auto a16 = [](size_t v) { return (v + 15) & ~size_t(15); };
    const size_t tb = a16(sizeof(SapTerm) * st->prog.nterms) + a16(sizeof(SapUniq) * st->prog.nuniqs) + a16(4 * (sd + 1));
    const size_t wb = (size_t)(sd + st->prog.npow + rows_per * sd) * kPad * sizeof(double);
    int warps = 4;
    while (warps > 1 && tb + warps * wb > (size_t)st->max_smem_optin) warps--;
    const size_t smem = tb + warps * wb;
    if (smem > (size_t)st->max_smem_optin) return set_error(TCHEM_EINVAL, "tchem_sap_jacobian2nd: set too large for the shared-memory tile");
    const void * fn = reinterpret_cast<const void *>(&sap_hess_kernel);
    { const int rc_ = ensure_max_dynamic_smem(fn, st->device, st->max_smem_optin); if (rc_ != TCHEM_OK) return rc_; }
    int per_sm = 0;
    TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, warps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t ntiles = (B + kLanes - 1) / kLanes;
    const int64_t grid = std::min<int64_t>((ntiles + warps - 1) / warps, (int64_t)st->sm_count * per_sm);
    if (grid < 1) return TCHEM_OK;
    void * args[] = {(void *)&st->prog, (void *)&x, (void *)&B, (void *)&K, (void *)&rows_per};
    TC_CUDA(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(warps * 32), args, smem, stream));
    count_launch("sap_hess_kernel");
    return TCHEM_OK;
}

int launch_sap(const tchem_sap_table * st, const tchem_ic_table * ict, const double * in, int64_t B,
               double * q_out, double * y, double * J, cudaStream_t stream) {
    const bool chained = ict != nullptr;
    {
        // large batches: the kernel specialised to this SAPSet (and, chained, to the IntCoordSet)
        const int rc = launch_jit_chain(st, ict, in, B, chained ? q_out : nullptr, y, J, stream);
        if (rc != -1) return rc;
    }
    const void * fn = !chained ? reinterpret_cast<const void *>(&sap_eval_kernel<false, false>)
                    : ict->has5 ? reinterpret_cast<const void *>(&sap_eval_kernel<true, true>)
                                : reinterpret_cast<const void *>(&sap_eval_kernel<true, false>);
    Program icp;
    std::memset(&icp, 0, sizeof(icp));
    if (chained) icp = ict->prog[MODE_VALUE];
    auto a16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    size_t tb = a16(sizeof(SapTerm) * st->prog.nterms) + a16(sizeof(SapUniq) * st->prog.nuniqs)
              + a16(sizeof(ChunkDesc) * st->prog.nchunks) + a16(4 * (st->sumdim + 1));
    if (chained) tb += a16(sizeof(PrimDesc) * icp.nprims) + a16(sizeof(ChunkDesc) * icp.nchunks) + a16(4 * icp.nmap)
                     + a16(sizeof(Src) * icp.nqsrcs);
    const size_t wb = (size_t)((chained ? icp.cartdim + icp.cap : 0) + st->sumdim + st->prog.npow + st->prog.cap) * kPad * sizeof(double);
    int warps = 4;
    while (warps > 1 && tb + warps * wb > (size_t)st->max_smem_optin) warps--;
    const size_t smem = tb + warps * wb;
    if (smem > (size_t)st->max_smem_optin) return set_error(TCHEM_EINVAL, "tchem_sap_eval: set too large for the shared-memory tile");
    { const int rc_ = ensure_max_dynamic_smem(fn, st->device, st->max_smem_optin); if (rc_ != TCHEM_OK) return rc_; }
    int per_sm = 0;
    TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, warps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t ntiles = (B + kLanes - 1) / kLanes;
    int64_t grid = std::min<int64_t>((ntiles + warps - 1) / warps, (int64_t)st->sm_count * per_sm);
    if (grid < 1) return TCHEM_OK;
    void * args[] = {(void *)&st->prog, (void *)&icp, (void *)&in, (void *)&B, (void *)&q_out, (void *)&y, (void *)&J};
    TC_CUDA(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(warps * 32), args, smem, stream));
    count_launch(chained ? "sap_eval_kernel<chained>" : "sap_eval_kernel");
    return TCHEM_OK;
}

} // namespace
} // namespace tchem_b200

using namespace tchem_b200;

extern "C" {

int tchem_sap_table_create(const int32_t * term_ptr, const int32_t * coords, int n_terms, int irreducible,
                           const int32_t * dims, int n_irreducibles, int device, tchem_sap_table ** out) {
    if (!term_ptr || !dims || !out || n_terms <= 0 || n_irreducibles <= 0)
        return set_error(TCHEM_EINVAL, "tchem_sap_table_create: empty definition");
    if (term_ptr[n_terms] > 0 && !coords) return set_error(TCHEM_EINVAL, "tchem_sap_table_create: coords is null");
    std::vector<int32_t> offset(n_irreducibles + 1, 0);
    for (int i = 0; i < n_irreducibles; i++) {
        if (dims[i] < 0) return set_error(TCHEM_EINVAL, "tchem_sap_table_create: negative dimension");
        offset[i + 1] = offset[i] + dims[i];
    }
    const int sumdim = offset[n_irreducibles];
    if (sumdim <= 0) return set_error(TCHEM_EINVAL, "tchem_sap_table_create: no coordinates");
    std::vector<SapTerm> terms(n_terms);
    std::vector<SapUniq> uniqs;
    for (int t = 0; t < n_terms; t++) {
        // SAPSet::construct_orders_, source/polynomial/SAPSet.cpp:47-48
        if (term_ptr[t + 1] == term_ptr[t] && irreducible != 0)
            return set_error(TCHEM_EINVAL, "tchem::SAPSet::construct_orders_: Only the totally symmetric irreducible can have 0th order term");
        std::map<int32_t, int32_t> powers;              // ascending column
        for (int c = term_ptr[t]; c < term_ptr[t + 1]; c++) {
            const int irr = coords[2 * c], idx = coords[2 * c + 1];
            if (irr < 0 || irr >= n_irreducibles || idx < 0 || idx >= dims[irr])
                return set_error(TCHEM_EINVAL, "tchem_sap_table_create: coordinate outside the declared dimensions");
            powers[offset[irr] + idx]++;
        }
        terms[t].ubegin = (int32_t)uniqs.size();
        for (auto & kv : powers) uniqs.push_back(SapUniq{kv.first, kv.second, 0, 0});
        terms[t].uend = (int32_t)uniqs.size();
    }
    // power table layout: column c owns entries pow_off[c] .. pow_off[c+1]-1 = x^0 .. x^maxpow(c)
    std::vector<int32_t> maxpow(sumdim, 0), pow_off(sumdim + 1, 0);
    for (auto & u : uniqs) maxpow[u.col] = std::max(maxpow[u.col], u.power);
    for (int c = 0; c < sumdim; c++) pow_off[c + 1] = pow_off[c] + maxpow[c] + 1;
    for (auto & u : uniqs) { u.e_pow = pow_off[u.col] + u.power; u.e_powm1 = u.e_pow - 1; }
    // chunk consecutive terms: entries = terms + their distinct coordinates
    int max_term = 1;
    for (auto & t : terms) max_term = std::max(max_term, 1 + t.uend - t.ubegin);
    const int total_entries = n_terms + (int)uniqs.size();
    const int cap = std::max(max_term, std::min(total_entries, 32));
    std::vector<ChunkDesc> chunks;
    std::vector<ElemMap> map;
    std::vector<uint32_t> srcs;
    std::vector<double> coefs;
    std::vector<ZeroRun> runs;
    for (int t0 = 0; t0 < n_terms;) {
        int t1 = t0, entries = 0;
        while (t1 < n_terms && entries + 1 + terms[t1].uend - terms[t1].ubegin <= cap) {
            entries += 1 + terms[t1].uend - terms[t1].ubegin;
            t1++;
        }
        // end on a 4-term boundary when possible: y (8 B per term) and J then fill whole 32-byte
        // sectors instead of leaving them half written between chunks
        if (t1 < n_terms && t1 % 4 != 0 && (t1 & ~3) > t0) t1 &= ~3;
        ChunkDesc ch;
        std::memset(&ch, 0, sizeof(ch));
        ch.prim_begin = t0; ch.prim_end = t1; ch.row0 = t0; ch.nrows = t1 - t0;
        std::vector<std::vector<Src>> jl((size_t)(t1 - t0) * sumdim);
        const int first = terms[t0].ubegin;
        for (int t = t0; t < t1; t++)
            for (int u = terms[t].ubegin; u < terms[t].uend; u++)
                jl[(size_t)(t - t0) * sumdim + uniqs[u].col].push_back(Src{(t1 - t0) + (u - first), 0, 1.0});
        if (!emit_span(jl, (int64_t)t0 * sumdim, (int64_t)n_terms * sumdim, map, srcs, coefs, runs, ch.jspan))
            return set_error(TCHEM_EINVAL, "tchem_sap_table_create: set too large");
        chunks.push_back(ch);
        t0 = t1;
    }
    int ndev = tchem_device_count();
    if (device < 0 || device >= ndev)
        return set_error(TCHEM_ECUDA, "tchem_sap_table_create: no such CUDA device (this library has no CPU path)");
    auto a16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    const size_t o_terms = 0, o_pow = o_terms + a16(terms.size() * sizeof(SapTerm));
    const size_t o_uniqs = o_pow + a16(pow_off.size() * 4);
    const size_t o_chunks = o_uniqs + a16(uniqs.size() * sizeof(SapUniq) + 8);
    const size_t o_map = o_chunks + a16(chunks.size() * sizeof(ChunkDesc));
    const size_t o_srcs = o_map + a16(map.size() * sizeof(ElemMap));
    const size_t o_coefs = o_srcs + a16(srcs.size() * sizeof(uint32_t));
    const size_t o_runs = o_coefs + a16(coefs.size() * sizeof(double) + 8);
    const size_t total = o_runs + a16(runs.size() * sizeof(ZeroRun)) + 16;
    std::vector<unsigned char> blob(total, 0);
    std::memcpy(blob.data() + o_terms, terms.data(), terms.size() * sizeof(SapTerm));
    std::memcpy(blob.data() + o_uniqs, uniqs.data(), uniqs.size() * sizeof(SapUniq));
    std::memcpy(blob.data() + o_pow, pow_off.data(), pow_off.size() * 4);
    std::memcpy(blob.data() + o_chunks, chunks.data(), chunks.size() * sizeof(ChunkDesc));
    std::memcpy(blob.data() + o_map, map.data(), map.size() * sizeof(ElemMap));
    std::memcpy(blob.data() + o_srcs, srcs.data(), srcs.size() * sizeof(uint32_t));
    std::memcpy(blob.data() + o_coefs, coefs.data(), coefs.size() * sizeof(double));
    std::memcpy(blob.data() + o_runs, runs.data(), runs.size() * sizeof(ZeroRun));
    DeviceGuard guard(device);
    if (!guard.ok) return set_error(TCHEM_ECUDA, "cudaSetDevice failed");
    void * dev = nullptr;
    TC_CUDA(cudaMalloc(&dev, total));
    cudaError_t e = cudaMemcpy(dev, blob.data(), total, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dev); return set_error(TCHEM_ECUDA, cudaGetErrorString(e)); }
    tchem_sap_table * t = new tchem_sap_table();
    t->device = device; t->nterms = n_terms; t->sumdim = sumdim; t->irreducible = irreducible;
    t->dims.assign(dims, dims + n_irreducibles);
    for (const SapTerm & tm : terms) {
        t->h_terms.emplace_back();
        for (int u = tm.ubegin; u < tm.uend; u++) t->h_terms.back().push_back({uniqs[u].col, uniqs[u].power});
    }
    cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&t->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    unsigned char * d = static_cast<unsigned char *>(dev);
    t->prog.terms = reinterpret_cast<const SapTerm *>(d + o_terms);
    t->prog.uniqs = reinterpret_cast<const SapUniq *>(d + o_uniqs);
    t->prog.pow_off = reinterpret_cast<const int32_t *>(d + o_pow);
    t->prog.npow = pow_off[sumdim];
    t->prog.chunks = reinterpret_cast<const ChunkDesc *>(d + o_chunks);
    t->prog.emap = reinterpret_cast<const ElemMap *>(d + o_map);
    t->prog.psrcs = reinterpret_cast<const uint4 *>(d + o_srcs);
    t->prog.coefs = reinterpret_cast<const double *>(d + o_coefs);
    t->prog.runs = reinterpret_cast<const ZeroRun *>(d + o_runs);
    t->prog.nterms = n_terms; t->prog.nuniqs = (int)uniqs.size(); t->prog.nchunks = (int)chunks.size();
    t->prog.cap = cap; t->prog.sumdim = sumdim;
    t->dev_blob = dev;
    *out = t;
    return TCHEM_OK;
}

void tchem_sap_table_destroy(tchem_sap_table * t) {
    if (!t) return;
    {
        DeviceGuard guard(t->device);
        release_jit_sap(t);
        cudaFree(t->dev_blob);
        for (int s = 0; s < SapHostPipe::kSlots; s++) {
            cudaFree(t->pipe.buf[s]);
            if (t->pipe.ok) cudaStreamDestroy(t->pipe.stream[s]);
        }
    }
    delete t;
}
int tchem_sap_table_nterms(const tchem_sap_table * t) { return t ? t->nterms : 0; }
int tchem_sap_table_sumdim(const tchem_sap_table * t) { return t ? t->sumdim : 0; }

int tchem_sap_eval(const tchem_sap_table * t, const double * x, int64_t B, double * y, double * J,
                   tchem_stream_t stream) {
    if (!t) return set_error(TCHEM_EINVAL, "tchem_sap_eval: null table");
    if (B < 0) return set_error(TCHEM_EINVAL, "tchem_sap_eval: negative batch");
    if (B == 0) return TCHEM_OK;
    if (!x || (!y && !J)) return set_error(TCHEM_EINVAL, "tchem_sap_eval: null argument");
    DeviceGuard guard(t->device);
    if (!guard.ok) return set_error(TCHEM_ECUDA, "cudaSetDevice failed");
    return launch_sap(t, nullptr, x, B, nullptr, y, J, static_cast<cudaStream_t>(stream));
}

int tchem_sap_jacobian2nd(const tchem_sap_table * t, const double * x, int64_t B, double * K, tchem_stream_t stream) {
    if (!t) return set_error(TCHEM_EINVAL, "tchem_sap_jacobian2nd: null table");
    if (B < 0) return set_error(TCHEM_EINVAL, "tchem_sap_jacobian2nd: negative batch");
    if (B == 0) return TCHEM_OK;
    if (!x || !K) return set_error(TCHEM_EINVAL, "tchem_sap_jacobian2nd: null argument");
    DeviceGuard guard(t->device);
    if (!guard.ok) return set_error(TCHEM_ECUDA, "cudaSetDevice failed");
    return launch_sap_hess(t, x, B, K, static_cast<cudaStream_t>(stream));
}

int tchem_ic_sap_eval(const tchem_ic_table * ic, const tchem_sap_table * sap, const double * r, int64_t B,
                      double * q_out, double * y, double * J, tchem_stream_t stream) {
    if (!ic || !sap) return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval: null table");
    if (ic->intdim != sap->sumdim)
        return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval: x must have a same dimension as the coordinates (intdim != sum of dimensions)");
    if (ic->device != sap->device) return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval: tables live on different devices");
    if (B < 0) return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval: negative batch");
    if (B == 0) return TCHEM_OK;
    if (!r || (!y && !J && !q_out)) return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval: null argument");
    int rc = ensure_program(const_cast<tchem_ic_table *>(ic), MODE_VALUE);
    if (rc != TCHEM_OK) return rc;
    DeviceGuard guard(ic->device);
    if (!guard.ok) return set_error(TCHEM_ECUDA, "cudaSetDevice failed");
    return launch_sap(sap, ic, r, B, q_out, y, J, static_cast<cudaStream_t>(stream));
}

// Host buffers in, host buffers out: pieces of the batch rotate over 3 stream/buffer slots so that
// the upload of piece i+1 and the download of piece i-1 overlap the kernel of piece i.
int tchem_ic_sap_eval_host(const tchem_ic_table * ic, const tchem_sap_table * sapc, const double * r, int64_t B,
                           double * q_out, double * y, double * J) {
    tchem_sap_table * sap = const_cast<tchem_sap_table *>(sapc);
    if (!ic || !sap) return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval_host: null table");
    if (ic->intdim != sap->sumdim)
        return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval_host: x must have a same dimension as the coordinates (intdim != sum of dimensions)");
    if (ic->device != sap->device) return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval_host: tables live on different devices");
    if (B < 0) return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval_host: negative batch");
    if (B == 0) return TCHEM_OK;
    if (!r || (!y && !J && !q_out)) return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval_host: null argument");
    int rc = ensure_program(const_cast<tchem_ic_table *>(ic), MODE_VALUE);
    if (rc != TCHEM_OK) return rc;
    DeviceGuard guard(ic->device);
    if (!guard.ok) return set_error(TCHEM_ECUDA, "cudaSetDevice failed");
    const size_t cd = ic->cartdim, n = sap->sumdim, nt = sap->nterms;
    const size_t per_geom = cd + (q_out ? n : 0) + (y ? nt : 0) + (J ? nt * n : 0);          // doubles
    int64_t piece = (int64_t)std::max<size_t>(32, ((size_t(48) << 20) / (8 * per_geom)) / 32 * 32);
    piece = std::min<int64_t>(piece, (B + 31) / 32 * 32);
    SapHostPipe & hp = sap->pipe;
    std::lock_guard<std::mutex> lock(hp.mutex);
    if (!hp.ok) {
        for (int s = 0; s < SapHostPipe::kSlots; s++) TC_CUDA(cudaStreamCreateWithFlags(&hp.stream[s], cudaStreamNonBlocking));
        hp.ok = true;
    }
    const size_t need = (size_t)piece * per_geom * 8;
    if (need > hp.cap) {
        for (int s = 0; s < SapHostPipe::kSlots; s++) {
            cudaFree(hp.buf[s]);
            hp.buf[s] = nullptr;
            TC_CUDA(cudaMalloc(reinterpret_cast<void **>(&hp.buf[s]), need));
        }
        hp.cap = need;
    }
    int slot = 0;
    struct HintScope { HintScope(int64_t b) { g_batch_hint = b; } ~HintScope() { g_batch_hint = 0; } } hint(B);
    for (int64_t b0 = 0; b0 < B; b0 += piece, slot = (slot + 1) % SapHostPipe::kSlots) {
        const int64_t nb = std::min<int64_t>(piece, B - b0);
        cudaStream_t s = hp.stream[slot];
        double * d_r = hp.buf[slot];
        double * d_q = d_r + piece * cd;
        double * d_y = d_q + (q_out ? piece * n : 0);
        double * d_J = d_y + (y ? piece * nt : 0);
        TC_CUDA(cudaMemcpyAsync(d_r, r + b0 * cd, nb * cd * 8, cudaMemcpyHostToDevice, s));
        rc = launch_sap(sap, ic, d_r, nb, q_out ? d_q : nullptr, y ? d_y : nullptr, J ? d_J : nullptr, s);
        if (rc != TCHEM_OK) {                      // no copy into the caller's buffers may stay pending
            for (int k = 0; k < SapHostPipe::kSlots; k++) cudaStreamSynchronize(hp.stream[k]);
            return rc;
        }
        if (q_out) TC_CUDA(cudaMemcpyAsync(q_out + b0 * n, d_q, nb * n * 8, cudaMemcpyDeviceToHost, s));
        if (y) TC_CUDA(cudaMemcpyAsync(y + b0 * nt, d_y, nb * nt * 8, cudaMemcpyDeviceToHost, s));
        if (J) TC_CUDA(cudaMemcpyAsync(J + b0 * nt * n, d_J, nb * nt * n * 8, cudaMemcpyDeviceToHost, s));
    }
    for (int s = 0; s < SapHostPipe::kSlots; s++) TC_CUDA(cudaStreamSynchronize(hp.stream[s]));
    return TCHEM_OK;
}

} // extern "C"
