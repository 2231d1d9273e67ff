// Gradient int -> cart:  g_r = J^T g_q, J never materialised.
// Replaces, batched: IntCoordSet::gradient_int2cart (reference source/IC/IntCoordSet.cpp:207-218).
//
// Warp tile of 32 geometries, lane = geometry. For every distinct invariant displacement p the
// lane forms the weight  W_p = sum over the terms that use p of coeff * g_q[ic]  and adds
// W_p * J_p (<= 15 entries, in registers) into its own column of a shared-memory g_r tile,
// which is then written out transposed (coalesced).
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
    auto a16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    const size_t o_p = 0, o_ptr = a16(prims.size() * sizeof(PrimDesc)), o_u = o_ptr + a16(ptr.size() * 4);
    const size_t total = o_u + a16(flat.size() * sizeof(GradUse)) + 16;
    std::vector<unsigned char> blob(total, 0);
    std::memcpy(blob.data() + o_p, prims.data(), prims.size() * sizeof(PrimDesc));
    std::memcpy(blob.data() + o_ptr, ptr.data(), ptr.size() * 4);
    std::memcpy(blob.data() + o_u, flat.data(), flat.size() * sizeof(GradUse));
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
    const void * fn = t->has5 ? reinterpret_cast<const void *>(&grad_int2cart_kernel<true>)
                              : reinterpret_cast<const void *>(&grad_int2cart_kernel<false>);
    const size_t wb = (size_t)(2 * t->cartdim + t->intdim) * kPad * sizeof(double);
    int warps = 4;
    while (warps > 1 && warps * wb > (size_t)t->max_smem_optin) warps >>= 1;
    const size_t smem = warps * wb;
    if (smem > (size_t)t->max_smem_optin)
        return set_error(TCHEM_EINVAL, "tchem::IC::IntCoordSet::gradient_int2cart: molecule too large for the shared-memory tile");
    TC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, warps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t ntiles = (B + kLanes - 1) / kLanes;
    int64_t grid = std::min<int64_t>((ntiles + warps - 1) / warps, (int64_t)t->sm_count * per_sm);
    void * args[] = {(void *)&gp, (void *)&r, (void *)&intgrad, (void *)&B, (void *)&cartgrad};
    TC_CUDA(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(warps * 32), args, smem, static_cast<cudaStream_t>(stream)));
    count_launch();
    return TCHEM_OK;
}

} // extern "C"
