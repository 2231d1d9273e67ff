// q / J / K kernel: one warp = one tile of 32 geometries (see ic_program.h for the scheme).
// Replaces, batched: IntCoordSet::operator() / compute_IC_J / compute_IC_J_K
// (reference source/IC/IntCoordSet.cpp:139-175).
#include "ic_eval.cuh"
#include <cstdio>
#include <cstdlib>

#include "ic_table.h"

namespace tchem_b200 {

namespace {

// dynamic shared memory: [PrimDesc x nprims][ChunkDesc x nchunks][per warp: R | P]
__device__ __forceinline__ size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

// launch bounds: <= 128 threads per block; 3 resident blocks (12 warps / SM) cap the q+J
// flavours at 168 registers, the K flavour keeps the full 255
template <int MODE, bool HAS5, int NW, int NB>
__global__ void __launch_bounds__(NW * 32, NB)
ic_eval_kernel(const Program prog, const double * __restrict__ r, const int64_t B,
               double * __restrict__ q, double * __restrict__ J, double * __restrict__ K) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PrimDesc * prims = reinterpret_cast<PrimDesc *>(smem_raw);
    size_t off = align16(sizeof(PrimDesc) * prog.nprims);
    ChunkDesc * chunks = reinterpret_cast<ChunkDesc *>(smem_raw + off);
    off += align16(sizeof(ChunkDesc) * prog.nchunks);
    Src * qsrcs = reinterpret_cast<Src *>(smem_raw + off);
    off += align16(sizeof(Src) * prog.nqsrcs);
    uint32_t * qmap = reinterpret_cast<uint32_t *>(smem_raw + off);
    off += align16(sizeof(uint32_t) * prog.nmap);
    const int span_bytes = prog.span_bytes;
    const int ncoef_s = (span_bytes == 0 && prog.ncoefs <= kMaxStagedCoefs) ? prog.ncoefs : 0;
    double * coef_s = reinterpret_cast<double *>(smem_raw + off);
    off += align16(sizeof(double) * ncoef_s);
    unsigned char * span_s = smem_raw + off;
    off += align16((size_t)span_bytes);
    const unsigned char * span_g = reinterpret_cast<const unsigned char *>(prog.emap);
    auto in_smem = [&](const void * g) { return span_s + (reinterpret_cast<const unsigned char *>(g) - span_g); };
    const ElemMap * emap = span_bytes ? reinterpret_cast<const ElemMap *>(span_s) : prog.emap;
    const uint4 * psrcs = span_bytes ? reinterpret_cast<const uint4 *>(in_smem(prog.psrcs)) : prog.psrcs;
    const ZeroRun * runs = span_bytes ? reinterpret_cast<const ZeroRun *>(in_smem(prog.runs)) : prog.runs;
    const double * coefs = span_bytes ? reinterpret_cast<const double *>(in_smem(prog.coefs)) : ncoef_s ? coef_s : prog.coefs;
    double * warp_base = reinterpret_cast<double *>(smem_raw + off);

    // stage the definition tables once per block (coalesced 4-byte copies)
    {
        const int32_t * src = reinterpret_cast<const int32_t *>(prog.prims);
        int32_t * dst = reinterpret_cast<int32_t *>(prims);
        const int n = prog.nprims * (int)(sizeof(PrimDesc) / 4);
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
        const int32_t * src2 = reinterpret_cast<const int32_t *>(prog.chunks);
        int32_t * dst2 = reinterpret_cast<int32_t *>(chunks);
        const int n2 = prog.nchunks * (int)(sizeof(ChunkDesc) / 4);
        for (int i = threadIdx.x; i < n2; i += blockDim.x) dst2[i] = src2[i];
        const int32_t * src3 = reinterpret_cast<const int32_t *>(prog.qsrcs);
        int32_t * dst3 = reinterpret_cast<int32_t *>(qsrcs);
        const int n3 = prog.nqsrcs * (int)(sizeof(Src) / 4);
        for (int i = threadIdx.x; i < n3; i += blockDim.x) dst3[i] = src3[i];
        const int n4 = prog.nmap;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) qmap[i] = prog.map[i];
        for (int i = threadIdx.x; i < ncoef_s; i += blockDim.x) coef_s[i] = prog.coefs[i];
        for (int i = threadIdx.x; i < span_bytes / 16; i += blockDim.x)
            reinterpret_cast<uint4 *>(span_s)[i] = reinterpret_cast<const uint4 *>(span_g)[i];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cartdim = prog.cartdim, intdim = prog.intdim;
    const size_t tile_doubles = (((size_t)(cartdim + prog.cap) * kPad) + 1) & ~size_t(1);
    double * R = warp_base + (size_t)warp * (tile_doubles + prog.line_elems);
    double * P = R + (size_t)cartdim * kPad;
    const int64_t ntiles = (B + kLanes - 1) / kLanes;
    const int64_t jstride = (int64_t)intdim * cartdim, kstride = jstride * cartdim;

    // Work items are (tile, chunk) pairs, handed out as one contiguous range per warp: consecutive
    // items of a warp share the tile (its coordinates stay staged), and a batch with few tiles
    // but wide outputs (q+J+K) still spreads over the whole grid.
    // (a set with rows split into accumulating passes keeps whole tiles together: the passes of a
    // row must run in order on one warp)
    const int nchunks = prog.split_rows ? 1 : prog.nchunks, cper = prog.split_rows ? prog.nchunks : 1;
    const int64_t items = ntiles * nchunks, nw = (int64_t)gridDim.x * nwarps, gw = (int64_t)blockIdx.x * nwarps + warp;
    const int64_t i0 = (items / nw) * gw + min(gw, items % nw), i1 = i0 + items / nw + (gw < items % nw ? 1 : 0);
    int64_t staged = -1;                                   // tile whose coordinates are in R (or in flight)
    for (int64_t it = i0; it < i1; it++) {
        const int64_t tile = it / nchunks;
        const int c = (int)(it - tile * nchunks);
        const int64_t b0 = tile * kLanes;
        const int ntile = (int)min((int64_t)kLanes, B - b0);
        if (staged != tile) stage_issue(r, b0, ntile, cartdim, R, lane);
        if (staged != tile || c == 0 || it == i0) stage_finish(ntile, cartdim, R, lane);
        staged = tile;
        for (int cc = c * cper; cc < (c + 1) * cper; cc++) {
        const ChunkDesc & ch = chunks[cc];
        SpanPrefetch pfJ;
        if (MODE != MODE_VALUE && J != nullptr) pfJ = span_prefetch(ch.jspan, emap, lane);
        eval_chunk<MODE, HAS5>(prims, ch.prim_begin, ch.prim_end, R, P, lane);
        if (q != nullptr) combine_rows(qmap + ch.qmap, qsrcs, ch.nrows, P, ch.qslot, lane);
        __syncwarp();
        // R is dead after a tile's last chunk: start fetching the next tile now, its DRAM latency
        // hides behind this chunk's stores
        if (c == nchunks - 1 && cc == (c + 1) * cper - 1 && it + 1 < i1) {
            const int64_t nb0 = (tile + 1) * kLanes;
            stage_issue(r, nb0, (int)min((int64_t)kLanes, B - nb0), cartdim, R, lane);
            staged = tile + 1;
        }
        if (q != nullptr)
            store_rows(P, ch.qslot, ch.nrows, q + b0 * intdim + ch.row0, intdim, ntile, ch.accumulate != 0, lane);
        // (no line-buffer write-out here: it only pays for the multi-row chunks of the warp-group
        // kernel, and even an untaken out-of-line call costs this kernel registers)
        if (MODE != MODE_VALUE && J != nullptr)
            gather_store(ch.jspan, emap, psrcs, coefs, runs, P, J + b0 * jstride + (int64_t)ch.row0 * cartdim,
                         jstride, ntile, ch.accumulate != 0, lane, pfJ);
        if (MODE == MODE_QJK && K != nullptr)
            gather_store<true>(ch.kspan, emap, psrcs, coefs, runs, P,
                         K + b0 * kstride + (int64_t)ch.row0 * cartdim * cartdim, kstride, ntile, ch.accumulate != 0, lane);
        __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------
// Warp-group flavour for mid-size molecules (q + J): G warps share one tile of 32 geometries - its
// coordinates are staged once per group - and take the tile's chunks round-robin, so in every round
// the group writes G adjacent row groups of the same 32 geometries at about the same time. Measured
// on the 20-atom workload: what limits the one-warp-per-tile kernel there is not the SM but the
// write pattern - 480-byte pieces 26 KB apart; the longer the contiguous piece per geometry, the
// faster the same bytes go out (and the shared coordinates leave room for 12 warps per SM).
// ---------------------------------------------------------------------------------------
template <bool HAS5, int G>
__global__ void __launch_bounds__(384, 1)
ic_eval_group_kernel(const Program prog, const double * __restrict__ r, const int64_t B,
                     double * __restrict__ q, double * __restrict__ J) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PrimDesc * prims = reinterpret_cast<PrimDesc *>(smem_raw);
    size_t off = align16(sizeof(PrimDesc) * prog.nprims);
    ChunkDesc * chunks = reinterpret_cast<ChunkDesc *>(smem_raw + off);
    off += align16(sizeof(ChunkDesc) * prog.nchunks);
    Src * qsrcs = reinterpret_cast<Src *>(smem_raw + off);
    off += align16(sizeof(Src) * prog.nqsrcs);
    uint32_t * qmap = reinterpret_cast<uint32_t *>(smem_raw + off);
    off += align16(sizeof(uint32_t) * prog.nmap);
    const int span_bytes = prog.span_bytes;
    const int ncoef_s = (span_bytes == 0 && prog.ncoefs <= kMaxStagedCoefs) ? prog.ncoefs : 0;
    double * coef_s = reinterpret_cast<double *>(smem_raw + off);
    off += align16(sizeof(double) * ncoef_s);
    unsigned char * span_s = smem_raw + off;
    off += align16((size_t)span_bytes);
    const unsigned char * span_g = reinterpret_cast<const unsigned char *>(prog.emap);
    auto in_smem = [&](const void * g) { return span_s + (reinterpret_cast<const unsigned char *>(g) - span_g); };
    const ElemMap * emap = span_bytes ? reinterpret_cast<const ElemMap *>(span_s) : prog.emap;
    const uint4 * psrcs = span_bytes ? reinterpret_cast<const uint4 *>(in_smem(prog.psrcs)) : prog.psrcs;
    const ZeroRun * runs = span_bytes ? reinterpret_cast<const ZeroRun *>(in_smem(prog.runs)) : prog.runs;
    const double * coefs = span_bytes ? reinterpret_cast<const double *>(in_smem(prog.coefs)) : ncoef_s ? coef_s : prog.coefs;
    double * warp_base = reinterpret_cast<double *>(smem_raw + off);
    {
        auto copy4 = [&](const void * src, void * dst, int n4) {
            for (int i = threadIdx.x; i < n4; i += blockDim.x) reinterpret_cast<int32_t *>(dst)[i] = reinterpret_cast<const int32_t *>(src)[i];
        };
        copy4(prog.prims, prims, prog.nprims * (int)(sizeof(PrimDesc) / 4));
        copy4(prog.chunks, chunks, prog.nchunks * (int)(sizeof(ChunkDesc) / 4));
        copy4(prog.qsrcs, qsrcs, prog.nqsrcs * (int)(sizeof(Src) / 4));
        copy4(prog.map, qmap, prog.nmap);
        copy4(prog.coefs, coef_s, 2 * ncoef_s);
        copy4(span_g, span_s, span_bytes / 4);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, ngroups = (blockDim.x >> 5) / G;
    const int grp = warp / G, wg = warp - grp * G, gt = wg * kLanes + lane;
    if (grp >= ngroups) return;                              // (block sizes are multiples of G warps)
    const int cartdim = prog.cartdim, intdim = prog.intdim, nchunks = prog.nchunks;
    const size_t tile_doubles = (((size_t)(cartdim + G * prog.cap) * kPad) + 1) & ~size_t(1);
    double * R = warp_base + (size_t)grp * (tile_doubles + (size_t)G * prog.line_elems);
    double * P = R + (size_t)(cartdim + wg * prog.cap) * kPad;
    double * L = R + tile_doubles + (size_t)wg * prog.line_elems;
    const int64_t ntiles = (B + kLanes - 1) / kLanes, jstride = (int64_t)intdim * cartdim;
    const int64_t ng = (int64_t)gridDim.x * ngroups, gg = (int64_t)blockIdx.x * ngroups + grp;
    const int64_t t0 = (ntiles / ng) * gg + min(gg, ntiles % ng), t1 = t0 + ntiles / ng + (gg < ntiles % ng ? 1 : 0);
    const int bar = 1 + grp;
    for (int64_t tile = t0; tile < t1; tile++) {
        const int64_t b0 = tile * kLanes;
        const int ntile = (int)min((int64_t)kLanes, B - b0);
        group_barrier(bar, G * kLanes);                      // every warp is done reading the previous tile
        stage_issue_group(r, b0, ntile, cartdim, R, gt, G * kLanes);
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        group_barrier(bar, G * kLanes);
        if (ntile < kLanes) {                                // ragged tile: idle lanes compute on a copy of the last geometry
            for (int k = wg; k < cartdim; k += G)
                if (lane >= ntile) R[k * kPad + lane] = R[k * kPad + ntile - 1];
            group_barrier(bar, G * kLanes);
        }
        for (int c = wg; c < nchunks; c += G) {
            const ChunkDesc & ch = chunks[c];
            SpanPrefetch pfJ;
            if (J != nullptr) pfJ = span_prefetch(ch.jspan, emap, lane);
            eval_chunk<MODE_QJ, HAS5>(prims, ch.prim_begin, ch.prim_end, R, P, lane);
            if (q != nullptr) combine_rows(qmap + ch.qmap, qsrcs, ch.nrows, P, ch.qslot, lane);
            __syncwarp();
            if (q != nullptr)
                store_rows(P, ch.qslot, ch.nrows, q + b0 * intdim + ch.row0, intdim, ntile, false, lane);
            if (J != nullptr) {
                double * Jc = J + b0 * jstride + (int64_t)ch.row0 * cartdim;
                const int E = ch.nrows * cartdim;
                if (prog.line_elems >= 2 * ((E + 1) & ~1) && line_eligible(ch.jspan)) {
                    if (ntile == kLanes) gather_store_line<true>(ch.jspan, psrcs, coefs, P, L, E, Jc, jstride, ntile, lane, pfJ);
                    else                 gather_store_line<false>(ch.jspan, psrcs, coefs, P, L, E, Jc, jstride, ntile, lane, pfJ);
                } else {
                    gather_store(ch.jspan, emap, psrcs, coefs, runs, P, Jc, jstride, ntile, false, lane, pfJ);
                }
            }
            __syncwarp();
        }
    }
}

template <int MODE, bool HAS5, int NW, int NB> const void * kernel_ptr() { return reinterpret_cast<const void *>(&ic_eval_kernel<MODE, HAS5, NW, NB>); }

} // namespace

size_t ic_eval_table_bytes(const Program & p) {
    auto a16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    return a16(sizeof(PrimDesc) * p.nprims) + a16(sizeof(ChunkDesc) * p.nchunks) + a16(sizeof(Src) * p.nqsrcs)
         + a16(sizeof(uint32_t) * p.nmap) + a16(sizeof(double) * (p.span_bytes == 0 && p.ncoefs <= kMaxStagedCoefs ? p.ncoefs : 0))
         + a16((size_t)p.span_bytes);
}
size_t ic_eval_warp_bytes(const Program & p) { return (((((size_t)(p.cartdim + p.cap) * kPad) + 1) & ~size_t(1)) + p.line_elems) * sizeof(double); }

int launch_jit_qJ(const tchem_ic_table * t, const double * r, int64_t B, double * q, double * J, cudaStream_t stream);

// picks warps/block and grid, launches on `stream`
int launch_ic_eval(const tchem_ic_table * t, int mode, const double * r, int64_t B,
                   double * q, double * J, double * K, cudaStream_t stream) {
    if (mode == MODE_QJ) {
        // large batches: the kernel specialised to this table (jit.cu); -1 = not eligible, interpret
        const int rc = launch_jit_qJ(t, r, B, q, J, stream);
        if (rc != -1) return rc;
    }
    const Program & p = t->prog[mode];
    // launch shapes (warps per block x resident blocks): the register budget of the q / q+J flavours
    // allows 12 warps per SM; whichever split of those 12 warps fits the most warps into shared
    // memory wins (the per-block tables are paid once per block). q+J+K: 2 warps per block.
    tchem_ic_table::LaunchCfg & cfg = t->cfg[mode];
    if (mode == MODE_QJ && t->group_warps > 0) {
        const int G = t->group_warps;
        if (!cfg.fn) {
            const bool has5 = t->has5;
            const void * fn = G == 2 ? (has5 ? (const void *)&ic_eval_group_kernel<true, 2> : (const void *)&ic_eval_group_kernel<false, 2>)
                            : G == 4 ? (has5 ? (const void *)&ic_eval_group_kernel<true, 4> : (const void *)&ic_eval_group_kernel<false, 4>)
                                     : (has5 ? (const void *)&ic_eval_group_kernel<true, 6> : (const void *)&ic_eval_group_kernel<false, 6>);
            const size_t tb = ic_eval_table_bytes(p), gb = (((((size_t)(p.cartdim + G * p.cap) * kPad) + 1) & ~size_t(1)) + (size_t)G * p.line_elems) * sizeof(double);
            int ngroups = 12 / G;
            while (ngroups > 1 && tb + ngroups * gb > (size_t)t->max_smem_optin) ngroups--;
            const size_t smem = tb + ngroups * gb;
            if (smem > (size_t)t->max_smem_optin)
                return set_error(TCHEM_EINVAL, "tchem_ic_eval: molecule too large for the shared-memory tile (cartdim + row capacity)");
            if (8.0 * p.intdim * p.cartdim * kLanes >= 2147483648.0)
                return set_error(TCHEM_EINVAL, "tchem_ic_eval: per-geometry output block too large");
            { const int rc = ensure_max_dynamic_smem(fn, t->device, t->max_smem_optin); if (rc != TCHEM_OK) return rc; }
            cfg.warps = ngroups * G; cfg.per_sm = 1; cfg.smem = smem;
            cfg.fn = fn;
        }
        const int ngroups = cfg.warps / G;
        const int64_t ntiles = (B + kLanes - 1) / kLanes;
        const int64_t grid = std::min<int64_t>((ntiles + ngroups - 1) / ngroups, (int64_t)t->sm_count);
        if (grid < 1) return TCHEM_OK;
        static const bool debug = getenv("TCHEM_DEBUG") != nullptr;
        if (debug) fprintf(stderr, "[tchem] ic_eval group kernel: G %d, grid %lld x %d threads, %zu B smem, cap %d, %d chunks\n",
                           G, (long long)grid, cfg.warps * 32, cfg.smem, p.cap, p.nchunks);
        void * args[] = {(void *)&p, (void *)&r, (void *)&B, (void *)&q, (void *)&J};
        TC_CUDA(cudaLaunchKernel(cfg.fn, dim3((unsigned)grid), dim3(cfg.warps * 32), args, cfg.smem, stream));
        count_launch("ic_eval_group_kernel");
        return TCHEM_OK;
    }
    if (!cfg.fn) {
        const size_t tb = ic_eval_table_bytes(p), wb = ic_eval_warp_bytes(p);
        const bool has5 = t->has5;
        const void * fn = nullptr;
        int warps = 0;
        if (mode == MODE_QJK) {
            // q+J+K: 255 registers; 3 warps per block measured marginally best on the 20-atom workload
            // (2 x 2 warps: 0.78, 3: 0.80, 4: 0.79 of the HBM roofline; TCHEM_K_WARPS overrides)
            static const int k_env = [] { const char * e = getenv("TCHEM_K_WARPS"); return e ? atoi(e) : 3; }();
            warps = std::min(std::max(k_env, 1), 4);
            while (warps > 1 && tb + warps * wb > (size_t)t->max_smem_optin) warps--;
            fn = warps > 2 ? (has5 ? kernel_ptr<MODE_QJK, true, 4, 1>() : kernel_ptr<MODE_QJK, false, 4, 1>())
                           : (has5 ? kernel_ptr<MODE_QJK, true, 2, 1>() : kernel_ptr<MODE_QJK, false, 2, 1>());
        } else {
            struct Shape { int nw, nb; const void * fn; };
            const Shape shapes[3] = {
                {4, 3, mode == MODE_VALUE ? kernel_ptr<MODE_VALUE, false, 4, 3>() : has5 ? kernel_ptr<MODE_QJ, true, 4, 3>() : kernel_ptr<MODE_QJ, false, 4, 3>()},
                {6, 2, mode == MODE_VALUE ? kernel_ptr<MODE_VALUE, false, 6, 2>() : has5 ? kernel_ptr<MODE_QJ, true, 6, 2>() : kernel_ptr<MODE_QJ, false, 6, 2>()},
                {12, 1, mode == MODE_VALUE ? kernel_ptr<MODE_VALUE, false, 12, 1>() : has5 ? kernel_ptr<MODE_QJ, true, 12, 1>() : kernel_ptr<MODE_QJ, false, 12, 1>()}};
            int best = 0;
            for (const Shape & sh : shapes) {
                // largest warp count <= nw that fits, times the blocks that then fit (1 KB reserved per block)
                int nw = sh.nw;
                while (nw > 1 && tb + nw * wb > (size_t)t->max_smem_optin) nw--;
                const size_t per_block = tb + nw * wb + 1024;
                const int nb = std::min<int>(sh.nb, (int)((size_t)t->max_smem_sm / per_block));
                if (nw * nb > best) { best = nw * nb; warps = nw; fn = sh.fn; }
            }
            if (!fn) { fn = shapes[0].fn; warps = 1; }
        }
        const size_t smem = tb + warps * wb;
        // the gather addresses geometry g of a tile as base + g * stride_bytes in 32 bits
        const double widest = 8.0 * p.intdim * p.cartdim * (mode == MODE_QJK ? (double)p.cartdim : 1.0);
        if (widest * kLanes >= 2147483648.0)
            return set_error(TCHEM_EINVAL, "tchem_ic_eval: per-geometry output block too large");
        if (smem > (size_t)t->max_smem_optin)
            return set_error(TCHEM_EINVAL, "tchem_ic_eval: molecule too large for the shared-memory tile (cartdim + row capacity)");
        { const int rc = ensure_max_dynamic_smem(fn, t->device, t->max_smem_optin); if (rc != TCHEM_OK) return rc; }
        int per_sm = 0;
        TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, warps * 32, smem));
        if (per_sm < 1) per_sm = 1;
        cfg.warps = warps; cfg.per_sm = per_sm; cfg.smem = smem;
        cfg.fn = fn;                  // last: a concurrent caller either sees the whole entry or resolves it again
    }
    const void * fn = cfg.fn;
    const int warps = cfg.warps, per_sm = cfg.per_sm;
    const size_t smem = cfg.smem;
    const int64_t ntiles = (B + kLanes - 1) / kLanes, items = ntiles * (p.split_rows ? 1 : p.nchunks);
    int64_t grid = (items + warps - 1) / warps;
    grid = std::min<int64_t>(grid, (int64_t)t->sm_count * per_sm);
    if (grid < 1) return TCHEM_OK;
    static const bool debug = getenv("TCHEM_DEBUG") != nullptr;
    if (debug) fprintf(stderr, "[tchem] ic_eval mode %d: grid %lld x %d threads, %zu B smem, %d blocks/SM, cap %d, %d chunks\n",
                       mode, (long long)grid, warps * 32, smem, per_sm, p.cap, p.nchunks);
    void * args[] = {(void *)&p, (void *)&r, (void *)&B, (void *)&q, (void *)&J, (void *)&K};
    TC_CUDA(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(warps * 32), args, smem, stream));
    count_launch(mode == MODE_VALUE ? "ic_eval_kernel<VALUE>" : mode == MODE_QJ ? "ic_eval_kernel<QJ>" : "ic_eval_kernel<QJK>");
    return TCHEM_OK;
}

} // namespace tchem_b200
