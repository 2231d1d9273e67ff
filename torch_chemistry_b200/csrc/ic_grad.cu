// Gradient int -> cart:  g_r = J^T g_q, J never materialised.
// Replaces, batched: IntCoordSet::gradient_int2cart (reference source/IC/IntCoordSet.cpp:207-218).
//
// Warp tile of 32 geometries, lane = geometry. For every distinct invariant displacement p the
// lane forms the weight  W_p = sum over the terms that use p of coeff * g_q[ic]  and adds
// W_p * J_p (<= 15 entries, in registers) into its own column of a shared-memory g_r tile,
// which is then written out transposed (coalesced).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "ic_eval.cuh"
#include "ic_table.h"

namespace tchem_b200 {

struct GradUse { int32_t ic; int32_t pad_; double coef; };
struct GradProgram {
    const PrimDesc * prims;       // out = first use, natoms etc. as usual
    const int32_t *  use_ptr;     // CSR over prims
    const GradUse *  uses;
    int32_t nprims, intdim, cartdim;
    // warp-per-geometry flavour: primitives sorted by type, every type padded to a multiple of 32 with
    // dummies (natoms = 0), pd.out = first slot of the primitive's weighted J in the per-warp buffer;
    // comp_ptr / comp_idx: for every Cartesian component the slots that add up to it
    const PrimDesc * wprims;
    const int32_t *  wuse_ptr;
    const GradUse *  wuses;
    const int32_t *  comp_ptr;
    const int32_t *  comp_idx;
    int32_t nwprims, nslots, ncomp_idx, nwuses;
};

namespace {

struct ScatterSink {
    double * G;      // g_r tile, lane folded in: G[col * kPad]
    double W;
    __device__ __forceinline__ void value(double) const {}
    __device__ __forceinline__ void jac_block(const int * atoms, int i, const V3<double> & J) const {
        double * g = G + 3 * atoms[i] * kPad;
        g[0] = fma(W, J.x, g[0]);
        g[kPad] = fma(W, J.y, g[kPad]);
        g[2 * kPad] = fma(W, J.z, g[2 * kPad]);
    }
    __device__ __forceinline__ void jac_entry(const int * atoms, int m, double j) const {
        const int a = m / 3;
        double * g = G + (3 * atoms[a] + (m - 3 * a)) * kPad;
        g[0] = fma(W, j, g[0]);
    }
    template <int N> __device__ __forceinline__ void hess_block(const int *, int, int, const V3<Dual> &) const {}
    __device__ __forceinline__ void hess_entry5(const int *, int, int, double) const {}
};

template <bool HAS5>
__global__ void __launch_bounds__(128)
grad_int2cart_kernel(const GradProgram gp, const double * __restrict__ r, const double * __restrict__ gq,
                     const int64_t B, double * __restrict__ gr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cartdim = gp.cartdim, intdim = gp.intdim;
    double * R = reinterpret_cast<double *>(smem_raw) + (size_t)warp * (2 * cartdim + intdim) * kPad;
    double * Q = R + (size_t)cartdim * kPad;      // g_q tile
    double * G = Q + (size_t)intdim * kPad;       // g_r tile
    const int64_t ntiles = (B + kLanes - 1) / kLanes;
    for (int64_t tile = (int64_t)blockIdx.x * nwarps + warp; tile < ntiles; tile += (int64_t)gridDim.x * nwarps) {
        const int64_t b0 = tile * kLanes;
        const int ntile = (int)min((int64_t)kLanes, B - b0);
        stage_coordinates(r, b0, ntile, cartdim, R, lane);
        stage_coordinates(gq, b0, ntile, intdim, Q, lane);
        for (int k = 0; k < cartdim; k++) G[k * kPad + lane] = 0.0;
        for (int p = 0; p < gp.nprims; p++) {
            const PrimDesc pd = gp.prims[p];
            double W = 0.0;
            for (int u = gp.use_ptr[p]; u < gp.use_ptr[p + 1]; u++) {
                const GradUse gu = gp.uses[u];
                W = fma(gu.coef, Q[gu.ic * kPad + lane], W);
            }
            ScatterSink sink;
            sink.G = G + lane;
            sink.W = W;
            eval_primitive<MODE_QJ, HAS5>(pd, TileLoader{R, lane}, sink);
        }
        __syncwarp();
        double * out = gr + b0 * cartdim;
        for (int k = lane; k < cartdim; k += kLanes)
            for (int g = 0; g < ntile; g++) out[(int64_t)g * cartdim + k] = G[k * kPad + g];
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------
// Warp-per-geometry flavour: lane = primitive. The lane-per-geometry kernel above needs
// (2 cartdim + intdim) * 264 bytes of shared memory per warp - 70 KB for a 30-atom molecule, two or
// three warps per SM - while the arithmetic per geometry is the same either way. Here a warp owns one
// geometry at a time: its lanes evaluate 32 primitives of the same type at once (tables sorted by
// type and padded), park W_p * J_p in a per-warp slot buffer (a few KB), and the Cartesian components
// are then gathered from the slots that contribute to them - no atomics, fixed summation order.
// ---------------------------------------------------------------------------------------
struct SlotSink {
    double * S;      // first slot of this primitive
    double W;
    __device__ __forceinline__ void value(double) const {}
    __device__ __forceinline__ void jac_block(const int *, int i, const V3<double> & J) const {
        S[3 * i] = W * J.x; S[3 * i + 1] = W * J.y; S[3 * i + 2] = W * J.z;
    }
    __device__ __forceinline__ void jac_entry(const int *, int m, double j) const { S[m] = W * j; }
    template <int N> __device__ __forceinline__ void hess_block(const int *, int, int, const V3<Dual> &) const {}
    __device__ __forceinline__ void hess_entry5(const int *, int, int, double) const {}
};

template <bool HAS5>
__global__ void __launch_bounds__(128, 3)
grad_int2cart_warp_kernel(const GradProgram gp, const double * __restrict__ r, const double * __restrict__ gq,
                          const int64_t B, double * __restrict__ gr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    auto a16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    // per block: the tables
    PrimDesc * prims = reinterpret_cast<PrimDesc *>(smem_raw);
    size_t off = a16(sizeof(PrimDesc) * gp.nwprims);
    int32_t * use_ptr = reinterpret_cast<int32_t *>(smem_raw + off);  off += a16(4 * (gp.nwprims + 1));
    GradUse * uses = reinterpret_cast<GradUse *>(smem_raw + off);     off += a16(sizeof(GradUse) * gp.nwuses);
    int32_t * comp_ptr = reinterpret_cast<int32_t *>(smem_raw + off); off += a16(4 * (gp.cartdim + 1));
    int32_t * comp_idx = reinterpret_cast<int32_t *>(smem_raw + off); off += a16(4 * gp.ncomp_idx);
    {
        auto copy4 = [&](const void * src, void * dst, int n4) {
            for (int i = threadIdx.x; i < n4; i += blockDim.x) reinterpret_cast<int32_t *>(dst)[i] = reinterpret_cast<const int32_t *>(src)[i];
        };
        copy4(gp.wprims, prims, gp.nwprims * (int)(sizeof(PrimDesc) / 4));
        copy4(gp.wuse_ptr, use_ptr, gp.nwprims + 1);
        copy4(gp.wuses, uses, gp.nwuses * (int)(sizeof(GradUse) / 4));
        copy4(gp.comp_ptr, comp_ptr, gp.cartdim + 1);
        copy4(gp.comp_idx, comp_idx, gp.ncomp_idx);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cartdim = gp.cartdim, intdim = gp.intdim;
    const size_t per_warp = (size_t)((cartdim + intdim + gp.nslots + 1) & ~1);
    double * rs = reinterpret_cast<double *>(smem_raw + off) + (size_t)warp * per_warp;
    double * qs = rs + cartdim;
    double * S = qs + intdim;
    // the next geometry's coordinates and weights are fetched into registers while the current one is
    // processed (up to kPre values per lane each; longer rows fall back to a plain load)
    constexpr int kPre = 4;
    const int64_t bstep = (int64_t)gridDim.x * nwarps;
    const bool pre = cartdim <= kPre * kLanes && intdim <= kPre * kLanes;
    double pr[kPre], pq[kPre];
    int64_t b = (int64_t)blockIdx.x * nwarps + warp;
    if (pre && b < B) {
#pragma unroll
        for (int u = 0; u < kPre; u++) {
            const int k = lane + u * kLanes;
            pr[u] = k < cartdim ? r[b * cartdim + k] : 0.0;
            pq[u] = k < intdim ? gq[b * intdim + k] : 0.0;
        }
    }
    for (; b < B; b += bstep) {
        if (pre) {
#pragma unroll
            for (int u = 0; u < kPre; u++) {
                const int k = lane + u * kLanes;
                if (k < cartdim) rs[k] = pr[u];
                if (k < intdim) qs[k] = pq[u];
            }
            const int64_t nb = b + bstep;
            if (nb < B) {
#pragma unroll
                for (int u = 0; u < kPre; u++) {
                    const int k = lane + u * kLanes;
                    pr[u] = k < cartdim ? r[nb * cartdim + k] : 0.0;
                    pq[u] = k < intdim ? gq[nb * intdim + k] : 0.0;
                }
            }
        } else {
            for (int k = lane; k < cartdim; k += kLanes) rs[k] = r[b * cartdim + k];
            for (int k = lane; k < intdim; k += kLanes) qs[k] = gq[b * intdim + k];
        }
        __syncwarp();
        for (int p = lane; p < gp.nwprims; p += kLanes) {
            const PrimDesc pd = prims[p];
            if (pd.natoms == 0) continue;                  // padding / dummy
            double W = 0.0;
            for (int u = use_ptr[p]; u < use_ptr[p + 1]; u++) W = fma(uses[u].coef, qs[uses[u].ic], W);
            SlotSink sink;
            sink.S = S + pd.out;
            sink.W = W;
            eval_primitive<MODE_QJ, HAS5>(pd, GeomLoader{rs}, sink);
        }
        __syncwarp();
        for (int c = lane; c < cartdim; c += kLanes) {
            // two partial sums: the slot loads of consecutive contributions do not wait for each other
            const int k0 = comp_ptr[c], k1 = comp_ptr[c + 1];
            double s0 = 0.0, s1 = 0.0;
            int k = k0;
            for (; k + 1 < k1; k += 2) { s0 += S[comp_idx[k]]; s1 += S[comp_idx[k + 1]]; }
            if (k < k1) s0 += S[comp_idx[k]];
            gr[b * cartdim + c] = s0 + s1;
        }
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------
// Two geometries per warp pass. The warp-per-geometry kernel above is bound by the latency of its
// dependent FP64 chains (rsqrt -> products -> blocks; ~25 % of the FP64 pipe at 12-16 warps per SM).
// Pair<double> is a scalar type that carries the same quantity of TWO geometries: the J formulas of
// ic_math.cuh are templates over the scalar type, so instantiating them on Pair evaluates geometry
// b and b + 1 in one pass with two independent instruction streams (ILP 2) and one table walk.
// ---------------------------------------------------------------------------------------
struct Pair { double a, b; };
__device__ __forceinline__ Pair mkp(double a, double b) { Pair r; r.a = a; r.b = b; return r; }
__device__ __forceinline__ Pair operator+(Pair x, Pair y) { return mkp(x.a + y.a, x.b + y.b); }
__device__ __forceinline__ Pair operator-(Pair x, Pair y) { return mkp(x.a - y.a, x.b - y.b); }
__device__ __forceinline__ Pair operator-(Pair x) { return mkp(-x.a, -x.b); }
__device__ __forceinline__ Pair operator*(Pair x, Pair y) { return mkp(x.a * y.a, x.b * y.b); }
__device__ __forceinline__ Pair operator-(double c, Pair x) { return mkp(c - x.a, c - x.b); }
__device__ __forceinline__ Pair rsqrt_(Pair x) { return mkp(rsqrt(x.a), rsqrt(x.b)); }

template <int DUMMY>
__global__ void __launch_bounds__(128, 3)
grad_int2cart_pair_kernel(const GradProgram gp, const double * __restrict__ r, const double * __restrict__ gq,
                          const int64_t B, double * __restrict__ gr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    auto a16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    PrimDesc * prims = reinterpret_cast<PrimDesc *>(smem_raw);
    size_t off = a16(sizeof(PrimDesc) * gp.nwprims);
    int32_t * use_ptr = reinterpret_cast<int32_t *>(smem_raw + off);  off += a16(4 * (gp.nwprims + 1));
    GradUse * uses = reinterpret_cast<GradUse *>(smem_raw + off);     off += a16(sizeof(GradUse) * gp.nwuses);
    int32_t * comp_ptr = reinterpret_cast<int32_t *>(smem_raw + off); off += a16(4 * (gp.cartdim + 1));
    int32_t * comp_idx = reinterpret_cast<int32_t *>(smem_raw + off); off += a16(4 * gp.ncomp_idx);
    {
        auto copy4 = [&](const void * src, void * dst, int n4) {
            for (int i = threadIdx.x; i < n4; i += blockDim.x) reinterpret_cast<int32_t *>(dst)[i] = reinterpret_cast<const int32_t *>(src)[i];
        };
        copy4(gp.wprims, prims, gp.nwprims * (int)(sizeof(PrimDesc) / 4));
        copy4(gp.wuse_ptr, use_ptr, gp.nwprims + 1);
        copy4(gp.wuses, uses, gp.nwuses * (int)(sizeof(GradUse) / 4));
        copy4(gp.comp_ptr, comp_ptr, gp.cartdim + 1);
        copy4(gp.comp_idx, comp_idx, gp.ncomp_idx);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cartdim = gp.cartdim, intdim = gp.intdim, nslots = gp.nslots;
    // per warp: coordinates and weights of both geometries (rows b, b + 1 are adjacent in global memory),
    // then the two slot buffers
    const size_t per_warp = (size_t)((2 * (cartdim + intdim + nslots) + 1) & ~1);
    double * rs = reinterpret_cast<double *>(smem_raw + off) + (size_t)warp * per_warp;   // [2][cartdim]
    double * qs = rs + 2 * cartdim;                                                       // [2][intdim]
    double * SA = qs + 2 * intdim, * SB = SA + nslots;
    const int64_t npairs = (B + 1) / 2, pstep = (int64_t)gridDim.x * nwarps;
    for (int64_t pi = (int64_t)blockIdx.x * nwarps + warp; pi < npairs; pi += pstep) {
        const int64_t b = 2 * pi;
        const bool two = b + 1 < B;                           // warp-uniform
        // a lone last geometry is evaluated twice (the second copy is never stored)
        const int nr = two ? 2 * cartdim : cartdim, nq = two ? 2 * intdim : intdim;
        for (int k = lane; k < 2 * cartdim; k += kLanes) rs[k] = r[b * cartdim + (k < nr ? k : k - cartdim)];
        for (int k = lane; k < 2 * intdim; k += kLanes) qs[k] = gq[b * intdim + (k < nq ? k : k - intdim)];
        __syncwarp();
        for (int p = lane; p < gp.nwprims; p += kLanes) {
            const PrimDesc pd = prims[p];
            if (pd.natoms == 0) continue;                  // padding / dummy
            double WA = 0.0, WB = 0.0;
            for (int u = use_ptr[p]; u < use_ptr[p + 1]; u++) {
                const GradUse gu = uses[u];
                WA = fma(gu.coef, qs[gu.ic], WA);
                WB = fma(gu.coef, qs[intdim + gu.ic], WB);
            }
            V3<Pair> a[4];
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < pd.natoms) {
                    const double * ra = rs + 3 * pd.atoms[i], * rb = ra + cartdim;
                    a[i] = v3<Pair>(mkp(ra[0], rb[0]), mkp(ra[1], rb[1]), mkp(ra[2], rb[2]));
                }
            double * sa = SA + pd.out, * sb = SB + pd.out;
            auto emit = [&](int i, const V3<Pair> & J) {
                sa[3 * i] = WA * J.x.a; sa[3 * i + 1] = WA * J.y.a; sa[3 * i + 2] = WA * J.z.a;
                sb[3 * i] = WB * J.x.b; sb[3 * i + 1] = WB * J.y.b; sb[3 * i + 2] = WB * J.z.b;
            };
            Pair q, sp = mkp(0.0, 0.0), cp = sp, st = sp;
            switch (pd.type) {
            case TCHEM_STRETCHING: stretching_J<Pair>(a, q, emit); break;
            case TCHEM_BENDING:    bending_J<Pair, false>(a, q, emit); break;
            case TCHEM_COSBENDING: bending_J<Pair, true>(a, q, emit); break;
            case TCHEM_TORSION:    torsion_J<Pair, 0>(a, sp, cp, st, emit); break;
            case TCHEM_SINTORSION: torsion_J<Pair, 1>(a, sp, cp, st, emit); break;
            case TCHEM_COSTORSION: torsion_J<Pair, 2>(a, sp, cp, st, emit); break;
            case TCHEM_OUTOFPLANE: oop_J<Pair, false>(a, q, emit); break;
            case TCHEM_SINOOP:     oop_J<Pair, true>(a, q, emit); break;
            default: break;
            }
        }
        __syncwarp();
        for (int c = lane; c < cartdim; c += kLanes) {
            const int k0 = comp_ptr[c], k1 = comp_ptr[c + 1];
            double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
            int k = k0;
            for (; k + 1 < k1; k += 2) {
                const int i0 = comp_idx[k], i1 = comp_idx[k + 1];
                s0 += SA[i0]; s1 += SA[i1];
                t0 += SB[i0]; t1 += SB[i1];
            }
            if (k < k1) { s0 += SA[comp_idx[k]]; t0 += SB[comp_idx[k]]; }
            gr[b * cartdim + c] = s0 + s1;
            if (two) gr[(b + 1) * cartdim + c] = t0 + t1;
        }
        __syncwarp();
    }
}

} // namespace

struct GradTable {
    GradProgram prog;
    void * blob = nullptr;
};

} // namespace tchem_b200

using namespace tchem_b200;

namespace tchem_b200 {

namespace {
std::map<const tchem_ic_table *, GradTable> & grad_tables() {
    static std::map<const tchem_ic_table *, GradTable> m;
    return m;
}
std::mutex & grad_mutex() { static std::mutex m; return m; }
}

void release_grad_program(const tchem_ic_table * t) {
    std::lock_guard<std::mutex> lock(grad_mutex());
    auto it = grad_tables().find(t);
    if (it != grad_tables().end()) { cudaFree(it->second.blob); grad_tables().erase(it); }
}

static int get_grad_program(const tchem_ic_table * t, GradProgram & out) {
    std::lock_guard<std::mutex> lock(grad_mutex());
    auto it = grad_tables().find(t);
    if (it != grad_tables().end()) { out = it->second.prog; return TCHEM_OK; }
    // distinct primitives + their uses
    typedef std::tuple<int, int, int, int, int, int, double> Key;
    std::map<Key, int> index;
    std::vector<PrimDesc> prims;
    std::vector<std::vector<GradUse>> uses;
    for (const auto & d : t->terms) {
        int a[5];
        for (int i = 0; i < 5; i++) a[i] = i < d.natoms ? d.atoms[i] : -1;
        Key k(d.type, a[0], a[1], a[2], a[3], a[4], 0.0);   // `min` does not influence J
        auto f = index.find(k);
        int id;
        if (f == index.end()) {
            id = (int)prims.size();
            index[k] = id;
            PrimDesc pd;
            std::memset(&pd, 0, sizeof(pd));
            pd.type = d.type; pd.natoms = d.natoms;
            for (int i = 0; i < 5; i++) pd.atoms[i] = i < d.natoms ? d.atoms[i] : 0;
            pd.min = d.min;
            prims.push_back(pd);
            uses.emplace_back();
        } else id = f->second;
        uses[id].push_back(GradUse{d.ic, 0, d.coeff});
    }
    std::vector<int32_t> ptr(1, 0);
    std::vector<GradUse> flat;
    for (auto & u : uses) { flat.insert(flat.end(), u.begin(), u.end()); ptr.push_back((int32_t)flat.size()); }
    // warp-per-geometry tables: by type, every type padded to whole warps; slot strides odd (bank spread)
    std::vector<int> order(prims.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return prims[a].type < prims[b].type; });
    std::vector<PrimDesc> wprims;
    std::vector<int32_t> wptr(1, 0);
    std::vector<GradUse> wflat;
    std::vector<std::vector<int32_t>> comp(t->cartdim);
    int nslots = 0;
    for (size_t k = 0; k < order.size(); k++) {
        const int id = order[k];
        if (k > 0 && prims[order[k - 1]].type != prims[id].type)
            while (wprims.size() % 32) { PrimDesc pad; std::memset(&pad, 0, sizeof(pad)); wprims.push_back(pad); wptr.push_back((int32_t)wflat.size()); }
        PrimDesc pd = prims[id];
        pd.out = nslots;
        for (int i = 0; i < pd.natoms; i++)
            for (int c = 0; c < 3; c++) comp[3 * pd.atoms[i] + c].push_back(nslots + 3 * i + c);
        nslots += (3 * pd.natoms) | 1;
        wprims.push_back(pd);
        wflat.insert(wflat.end(), uses[id].begin(), uses[id].end());
        wptr.push_back((int32_t)wflat.size());
    }
    std::vector<int32_t> comp_ptr(1, 0), comp_idx;
    for (auto & c : comp) { comp_idx.insert(comp_idx.end(), c.begin(), c.end()); comp_ptr.push_back((int32_t)comp_idx.size()); }
    auto a16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    const size_t o_p = 0, o_ptr = a16(prims.size() * sizeof(PrimDesc)), o_u = o_ptr + a16(ptr.size() * 4);
    const size_t o_wp = o_u + a16(flat.size() * sizeof(GradUse));
    const size_t o_wptr = o_wp + a16(wprims.size() * sizeof(PrimDesc));
    const size_t o_wu = o_wptr + a16(wptr.size() * 4);
    const size_t o_cp = o_wu + a16(wflat.size() * sizeof(GradUse));
    const size_t o_ci = o_cp + a16(comp_ptr.size() * 4);
    const size_t total = o_ci + a16(comp_idx.size() * 4) + 16;
    std::vector<unsigned char> blob(total, 0);
    std::memcpy(blob.data() + o_p, prims.data(), prims.size() * sizeof(PrimDesc));
    std::memcpy(blob.data() + o_ptr, ptr.data(), ptr.size() * 4);
    std::memcpy(blob.data() + o_u, flat.data(), flat.size() * sizeof(GradUse));
    std::memcpy(blob.data() + o_wp, wprims.data(), wprims.size() * sizeof(PrimDesc));
    std::memcpy(blob.data() + o_wptr, wptr.data(), wptr.size() * 4);
    std::memcpy(blob.data() + o_wu, wflat.data(), wflat.size() * sizeof(GradUse));
    std::memcpy(blob.data() + o_cp, comp_ptr.data(), comp_ptr.size() * 4);
    std::memcpy(blob.data() + o_ci, comp_idx.data(), comp_idx.size() * 4);
    void * dev = nullptr;
    TC_CUDA(cudaMalloc(&dev, total));
    cudaError_t e = cudaMemcpy(dev, blob.data(), total, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dev); return set_error(TCHEM_ECUDA, cudaGetErrorString(e)); }
    GradTable gt;
    unsigned char * d = static_cast<unsigned char *>(dev);
    gt.blob = dev;
    gt.prog.prims = reinterpret_cast<const PrimDesc *>(d + o_p);
    gt.prog.use_ptr = reinterpret_cast<const int32_t *>(d + o_ptr);
    gt.prog.uses = reinterpret_cast<const GradUse *>(d + o_u);
    gt.prog.nprims = (int)prims.size();
    gt.prog.intdim = t->intdim;
    gt.prog.cartdim = t->cartdim;
    gt.prog.wprims = reinterpret_cast<const PrimDesc *>(d + o_wp);
    gt.prog.wuse_ptr = reinterpret_cast<const int32_t *>(d + o_wptr);
    gt.prog.wuses = reinterpret_cast<const GradUse *>(d + o_wu);
    gt.prog.comp_ptr = reinterpret_cast<const int32_t *>(d + o_cp);
    gt.prog.comp_idx = reinterpret_cast<const int32_t *>(d + o_ci);
    gt.prog.nwprims = (int)wprims.size();
    gt.prog.nslots = nslots;
    gt.prog.ncomp_idx = (int)comp_idx.size();
    gt.prog.nwuses = (int)wflat.size();
    grad_tables()[t] = gt;
    out = gt.prog;
    return TCHEM_OK;
}

} // namespace tchem_b200

extern "C" {

int tchem_ic_grad_int2cart(const tchem_ic_table * t, const double * r, const double * intgrad, int64_t B,
                           double * cartgrad, tchem_stream_t stream) {
    if (!t) return set_error(TCHEM_EINVAL, "tchem::IC::IntCoordSet::gradient_int2cart: null table");
    if (B < 0) return set_error(TCHEM_EINVAL, "tchem::IC::IntCoordSet::gradient_int2cart: negative batch");
    if (B == 0) return TCHEM_OK;
    if (!r || !intgrad || !cartgrad) return set_error(TCHEM_EINVAL, "tchem::IC::IntCoordSet::gradient_int2cart: null argument");
    DeviceGuard guard(t->device);
    if (!guard.ok) return set_error(TCHEM_ECUDA, "cudaSetDevice failed");
    GradProgram gp;
    int rc = get_grad_program(t, gp);
    if (rc != TCHEM_OK) return rc;
    {
        // warp-per-geometry kernel whenever its tables and slot buffers fit (TCHEM_GRAD_TILE=1 forces
        // the lane-per-geometry kernel)
        static const bool force_tile = [] { const char * e = getenv("TCHEM_GRAD_TILE"); return e && atoi(e) != 0; }();
        auto a16 = [](size_t x) { return (x + 15) & ~size_t(15); };
        const size_t tb = a16(sizeof(PrimDesc) * gp.nwprims) + a16(4 * (gp.nwprims + 1)) + a16(sizeof(GradUse) * gp.nwuses)
                        + a16(4 * (gp.cartdim + 1)) + a16(4 * gp.ncomp_idx);
        const size_t pw = (size_t)((gp.cartdim + gp.intdim + gp.nslots + 1) & ~1) * sizeof(double);
        const int wwarps = 4;
        // two geometries per warp pass (ILP 2) for tables without 5-atom terms; TCHEM_GRAD_PAIR=0 keeps one
        static const bool use_pair = [] { const char * e = getenv("TCHEM_GRAD_PAIR"); return e && atoi(e) != 0; }();   // opt-in: measured 15 % slower than one geometry per pass on C4
        const size_t pw2 = (size_t)((2 * (gp.cartdim + gp.intdim + gp.nslots) + 1) & ~1) * sizeof(double);
        if (!force_tile && use_pair && !t->has5 && tb + wwarps * pw2 <= (size_t)t->max_smem_optin / 2) {
            const void * pfn = reinterpret_cast<const void *>(&grad_int2cart_pair_kernel<0>);
            const size_t psmem = tb + wwarps * pw2;
            { const int rc_ = ensure_max_dynamic_smem(pfn, t->device, t->max_smem_optin); if (rc_ != TCHEM_OK) return rc_; }
            int per_sm = 0;
            TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pfn, wwarps * 32, psmem));
            if (per_sm < 1) per_sm = 1;
            const int64_t npairs = (B + 1) / 2;
            const int64_t grid = std::min<int64_t>((npairs + wwarps - 1) / wwarps, (int64_t)t->sm_count * per_sm);
            void * args[] = {(void *)&gp, (void *)&r, (void *)&intgrad, (void *)&B, (void *)&cartgrad};
            TC_CUDA(cudaLaunchKernel(pfn, dim3((unsigned)grid), dim3(wwarps * 32), args, psmem, static_cast<cudaStream_t>(stream)));
            count_launch("grad_int2cart_pair_kernel");
            return TCHEM_OK;
        }
        const size_t wsmem = tb + wwarps * pw;
        if (!force_tile && wsmem <= (size_t)t->max_smem_optin / 2) {
            const void * wfn = t->has5 ? reinterpret_cast<const void *>(&grad_int2cart_warp_kernel<true>)
                                       : reinterpret_cast<const void *>(&grad_int2cart_warp_kernel<false>);
            { const int rc_ = ensure_max_dynamic_smem(wfn, t->device, t->max_smem_optin); if (rc_ != TCHEM_OK) return rc_; }
            int per_sm = 0;
            TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wfn, wwarps * 32, wsmem));
            if (per_sm < 1) per_sm = 1;
            const int64_t grid = std::min<int64_t>((B + wwarps - 1) / wwarps, (int64_t)t->sm_count * per_sm);
            void * args[] = {(void *)&gp, (void *)&r, (void *)&intgrad, (void *)&B, (void *)&cartgrad};
            TC_CUDA(cudaLaunchKernel(wfn, dim3((unsigned)grid), dim3(wwarps * 32), args, wsmem, static_cast<cudaStream_t>(stream)));
            count_launch("grad_int2cart_warp_kernel");
            return TCHEM_OK;
        }
    }
    const void * fn = t->has5 ? reinterpret_cast<const void *>(&grad_int2cart_kernel<true>)
                              : reinterpret_cast<const void *>(&grad_int2cart_kernel<false>);
    const size_t wb = (size_t)(2 * t->cartdim + t->intdim) * kPad * sizeof(double);
    int warps = 4;
    while (warps > 1 && warps * wb > (size_t)t->max_smem_optin) warps--;
    const size_t smem = warps * wb;
    if (smem > (size_t)t->max_smem_optin)
        return set_error(TCHEM_EINVAL, "tchem::IC::IntCoordSet::gradient_int2cart: molecule too large for the shared-memory tile");
    { const int rc_ = ensure_max_dynamic_smem(fn, t->device, t->max_smem_optin); if (rc_ != TCHEM_OK) return rc_; }
    int per_sm = 0;
    TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, warps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t ntiles = (B + kLanes - 1) / kLanes;
    int64_t grid = std::min<int64_t>((ntiles + warps - 1) / warps, (int64_t)t->sm_count * per_sm);
    void * args[] = {(void *)&gp, (void *)&r, (void *)&intgrad, (void *)&B, (void *)&cartgrad};
    TC_CUDA(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(warps * 32), args, smem, static_cast<cudaStream_t>(stream)));
    count_launch("grad_int2cart_kernel");
    return TCHEM_OK;
}

} // extern "C"
