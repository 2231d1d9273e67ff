// Chained CARTESIAN derivatives of a SAPSet evaluated on internal coordinates:
//   dy/dr   [B][NSAP][cartdim]          = (dSAP/dq) J
//   d2y/dr2 [B][NSAP][cartdim][cartdim] = J^T (d2SAP/dq2) J + sum_k (dSAP/dq_k) K[k]
// The reference has no entry point for this: its callers compose SAPSet::Jacobian_ / Jacobian2nd_
// (source/polynomial/SAPSet.cpp:178-201, :243-276) with IntCoordSet::compute_IC_J / compute_IC_J_K
// (source/IC/IntCoordSet.cpp:147-175) by hand (test/normal_mode/main.cpp:56-61).
//
// First order, large batches of small molecules: ONE kernel specialised to the (IntCoordSet, SAPSet) pair
// (jit.cu, tchem_jit_chain_cart) - q, J and dSAP/dq never leave the chip. Everything else (small batches, large
// molecules, the second order) is composed from the library's own kernels through a stream-ordered workspace,
// slice by slice, plus the two contraction kernels below.
#include <algorithm>
#include <cstdlib>

#include "ic_table.h"
#include "sap_table.h"

namespace tchem_b200 {

int launch_jit_chain(const tchem_sap_table * st, const tchem_ic_table * t, const double * r, int64_t B,
                     double * q, double * y, double * J, cudaStream_t stream, bool cart);

namespace {

// dydr[b][t][c] = sum_k g[b][t][k] J[b][k][c]; one thread per output element, consecutive threads along c
__global__ void __launch_bounds__(256)
chain_cart_first_kernel(const double * __restrict__ g, const double * __restrict__ J, int64_t B, int nt, int n, int cd,
                        double * __restrict__ out) {
    const int64_t total = B * nt * cd;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % cd);
        const int64_t bt = e / cd;
        const int64_t b = bt / nt;
        const double * gp = g + bt * n, * Jp = J + b * n * cd + c;
        double s = 0.0;
        for (int k = 0; k < n; k++) s = fma(gp[k], Jp[(int64_t)k * cd], s);
        out[e] = s;
    }
}

// out[b][t][a][c] = sum_kl J[b][k][a] H[b][t][k][l] J[b][l][c] + sum_k g[b][t][k] K[b][k][a][c]
__global__ void __launch_bounds__(256)
chain_cart_second_kernel(const double * __restrict__ g, const double * __restrict__ H, const double * __restrict__ J,
                         const double * __restrict__ K, int64_t B, int nt, int n, int cd, double * __restrict__ out) {
    const int64_t total = B * nt * cd * cd;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % cd);
        const int a = (int)((e / cd) % cd);
        const int64_t bt = e / ((int64_t)cd * cd);
        const int64_t b = bt / nt;
        const double * gp = g + bt * n, * Hp = H + bt * n * n, * Jb = J + b * n * cd, * Kb = K + b * (int64_t)n * cd * cd + (int64_t)a * cd + c;
        double s = 0.0;
        for (int k = 0; k < n; k++) {
            double u = 0.0;                                    // (H J)[k][c]
            for (int l = 0; l < n; l++) u = fma(Hp[k * n + l], Jb[l * cd + c], u);
            s = fma(Jb[k * cd + a], u, s);
            s = fma(gp[k], Kb[(int64_t)k * cd * cd], s);
        }
        out[e] = s;
    }
}

// The same contraction with the geometry's J and K staged in shared memory: one block per geometry at a time, its warps
// take the monomials in turn - V = H_t J first (lanes over the n x cd elements), then the cartdim^2 outputs of the term
// (lanes over (a, c): consecutive lanes write consecutive doubles, K[k][a][c] and the output are read / written contiguously).
// Per element the same fma order as the kernel above (bit-identical), which serves the molecules whose K does not fit.
// The element-per-thread version re-read H, J and K from global memory for every output: 95 % of the composed path's time.
constexpr int kCart2Warps = 4;
__host__ __device__ inline size_t cart2_smem_doubles(int n, int cd, int nt) {
    return (size_t)n * cd + (size_t)n * cd * cd + (size_t)nt * ((size_t)n + (size_t)n * n) + (size_t)kCart2Warps * (size_t)n * cd;
}
__global__ void __launch_bounds__(32 * kCart2Warps)
chain_cart_second_block_kernel(const double * __restrict__ g, const double * __restrict__ H, const double * __restrict__ J,
                               const double * __restrict__ K, int64_t B, int nt, int n, int cd, double * __restrict__ out) {
    extern __shared__ __align__(16) double sm2[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncd = n * cd, nk = n * cd * cd, ncc = cd * cd, nn = n * n;
    // one coalesced load phase per geometry (J, K and every monomial's gradient and Hessian): loads from global memory inside
    // the per-term loops cost one L2 latency per loop trip; several resident blocks per SM overlap this phase with the others' arithmetic
    double * Js = sm2, * Ks = Js + ncd, * Gs = Ks + nk, * Hs = Gs + (size_t)nt * n, * Vs = Hs + (size_t)nt * nn + (size_t)warp * ncd;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();                                           // the previous geometry's operands are no longer read
        for (int i = threadIdx.x; i < ncd; i += blockDim.x) Js[i] = J[b * ncd + i];
        for (int i = threadIdx.x; i < nk; i += blockDim.x) Ks[i] = K[b * (int64_t)nk + i];
        for (int i = threadIdx.x; i < nt * n; i += blockDim.x) Gs[i] = g[b * nt * n + i];
        for (int i = threadIdx.x; i < nt * nn; i += blockDim.x) Hs[i] = H[b * (int64_t)nt * nn + i];
        __syncthreads();
        for (int t = warp; t < nt; t += kCart2Warps) {
            const double * gs = Gs + t * n, * Hp = Hs + t * nn;
            for (int e = lane; e < ncd; e += 32) {
                const int k = e / cd, c = e - k * cd;
                double u = 0.0;                                    // (H J)[k][c]
                for (int l = 0; l < n; l++) u = fma(Hp[k * n + l], Js[l * cd + c], u);
                Vs[e] = u;
            }
            __syncwarp();
            double * op = out + (b * nt + t) * (int64_t)ncc;
            for (int e = lane; e < ncc; e += 32) {
                const int a = e / cd, c = e - a * cd;
                double s = 0.0;
                for (int k = 0; k < n; k++) {
                    s = fma(Js[k * cd + a], Vs[k * cd + c], s);
                    s = fma(gs[k], Ks[(k * cd + a) * cd + c], s);
                }
                __stcs(op + e, s);
            }
            __syncwarp();
        }
    }
}

// The same kernel for small molecules with the sizes as template parameters (N internal coordinates, IT = ceil(cartdim^2 / 32)
// outputs per lane): the lane's (a, c) pairs, J[k][a] and K[k][a][c] do not depend on the monomial and live in registers across the
// whole term loop, the k loops are unrolled. The generic version spends 66 instructions per output on 6 useful FMAs (index
// division, runtime loop bounds, three shared-memory loads per FMA pair) and is issue bound.
template <int N, int IT>
__global__ void __launch_bounds__(32 * kCart2Warps)
chain_cart_second_small_kernel(const double * __restrict__ g, const double * __restrict__ H, const double * __restrict__ J,
                               const double * __restrict__ K, int64_t B, int nt, int cd, double * __restrict__ out) {
    extern __shared__ __align__(16) double sm2[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncd = N * cd, nk = N * cd * cd, ncc = cd * cd;
    double * Js = sm2, * Ks = Js + ncd, * Gs = Ks + nk, * Hs = Gs + (size_t)nt * N, * Vs = Hs + (size_t)nt * N * N + (size_t)warp * ncd;
    int ea[IT], ec[IT];
#pragma unroll
    for (int i = 0; i < IT; i++) {
        const int e = min(lane + 32 * i, ncc - 1);
        ea[i] = e / cd; ec[i] = e - ea[i] * cd;
    }
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < ncd; i += blockDim.x) Js[i] = J[b * ncd + i];
        for (int i = threadIdx.x; i < nk; i += blockDim.x) Ks[i] = K[b * (int64_t)nk + i];
        for (int i = threadIdx.x; i < nt * N; i += blockDim.x) Gs[i] = g[b * nt * N + i];
        for (int i = threadIdx.x; i < nt * N * N; i += blockDim.x) Hs[i] = H[b * (int64_t)nt * N * N + i];
        __syncthreads();
        double Jr[IT][N], Kr[IT][N];
#pragma unroll
        for (int i = 0; i < IT; i++)
#pragma unroll
            for (int k = 0; k < N; k++) { Jr[i][k] = Js[k * cd + ea[i]]; Kr[i][k] = Ks[(k * cd + ea[i]) * cd + ec[i]]; }
        for (int t = warp; t < nt; t += kCart2Warps) {
            const double * gs = Gs + t * N, * Hp = Hs + t * N * N;
            for (int e = lane; e < ncd; e += 32) {
                const int k = e / cd, c = e - k * cd;
                double u = 0.0;                                    // (H J)[k][c]
#pragma unroll
                for (int l = 0; l < N; l++) u = fma(Hp[k * N + l], Js[l * cd + c], u);
                Vs[e] = u;
            }
            double gr[N];
#pragma unroll
            for (int k = 0; k < N; k++) gr[k] = gs[k];
            __syncwarp();
            double * op = out + (b * nt + t) * (int64_t)ncc;
#pragma unroll
            for (int i = 0; i < IT; i++) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < N; k++) {
                    s = fma(Jr[i][k], Vs[k * cd + ec[i]], s);
                    s = fma(gr[k], Kr[i][k], s);
                }
                if (lane + 32 * i < ncc) __stcs(op + lane + 32 * i, s);
            }
            __syncwarp();
        }
    }
}

} // namespace
} // namespace tchem_b200

using namespace tchem_b200;

extern "C" int tchem_ic_sap_eval_cart(const tchem_ic_table * ic, const tchem_sap_table * sap, const double * r, int64_t B,
                                      double * y, double * dydr, double * d2ydr2, tchem_stream_t stream_) {
    if (!ic || !sap) return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval_cart: null table");
    if (ic->intdim != sap->sumdim)
        return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval_cart: x must have a same dimension as the coordinates (intdim != sum of dimensions)");
    if (ic->device != sap->device) return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval_cart: tables live on different devices");
    if (B < 0) return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval_cart: negative batch");
    if (B == 0) return TCHEM_OK;
    if (!r || (!y && !dydr && !d2ydr2)) return set_error(TCHEM_EINVAL, "tchem_ic_sap_eval_cart: null argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DeviceGuard guard(ic->device);
    if (!guard.ok) return set_error(TCHEM_ECUDA, "cudaSetDevice failed");
    const int n = ic->intdim, cd = ic->cartdim, nt = sap->nterms;
    bool first_done = false;
    if (dydr) {
        const int rc = launch_jit_chain(sap, ic, r, B, nullptr, y, dydr, stream, true);
        if (rc != -1 && rc != TCHEM_OK) return rc;
        first_done = rc == TCHEM_OK;
    }
    if (first_done && !d2ydr2) return TCHEM_OK;
    // composition through a workspace: q | dSAP/dq | J [| K | d2SAP/dq2] per geometry, slices of at most ~256 MB
    const bool second = d2ydr2 != nullptr;
    const size_t per_geom = (size_t)n + (size_t)nt * n + (size_t)n * cd + (second ? (size_t)n * cd * cd + (size_t)nt * n * n : 0);
    static const size_t budget = [] { const char * e = getenv("TCHEM_CART_WORKSPACE_MB"); return (size_t)(e ? atoi(e) : 256) << 20; }();
    int64_t slice = (int64_t)std::max<size_t>(1, budget / (per_geom * sizeof(double)));
    slice = std::min<int64_t>(slice, B);
    double * ws = nullptr;
    keep_async_pool(ic->device);
    if (cudaMallocAsync(reinterpret_cast<void **>(&ws), (size_t)slice * per_geom * sizeof(double), stream) != cudaSuccess) {
        cudaGetLastError();
        return set_error(TCHEM_ENOMEM, "tchem_ic_sap_eval_cart: cannot allocate the workspace");
    }
    double * q_ws = ws, * g_ws = q_ws + (size_t)slice * n, * J_ws = g_ws + (size_t)slice * nt * n;
    double * K_ws = second ? J_ws + (size_t)slice * n * cd : nullptr;
    double * H_ws = second ? K_ws + (size_t)slice * n * cd * cd : nullptr;
    int rc = TCHEM_OK;
    for (int64_t b0 = 0; b0 < B && rc == TCHEM_OK; b0 += slice) {
        const int64_t nb = std::min<int64_t>(slice, B - b0);
        const double * rr = r + b0 * cd;
        rc = tchem_ic_sap_eval(ic, sap, rr, nb, q_ws, (y && !first_done) ? y + b0 * nt : nullptr, g_ws, stream_);
        if (rc == TCHEM_OK) rc = tchem_ic_eval(ic, rr, nb, nullptr, J_ws, K_ws, 0, stream_);
        if (rc == TCHEM_OK && dydr && !first_done) {
            const int64_t total = nb * nt * cd;
            const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)ic->sm_count * 8);
            chain_cart_first_kernel<<<grid, 256, 0, stream>>>(g_ws, J_ws, nb, nt, n, cd, dydr + b0 * nt * cd);
            count_launch("chain_cart_first_kernel");
        }
        if (rc == TCHEM_OK && second) {
            rc = tchem_sap_jacobian2nd(sap, q_ws, nb, H_ws, stream_);
            if (rc == TCHEM_OK) {
                const size_t smem2 = cart2_smem_doubles(n, cd, nt) * sizeof(double);
                if (smem2 + 1024 <= (size_t)ic->max_smem_optin) {
                    const void * fn = reinterpret_cast<const void *>(&chain_cart_second_block_kernel);
                    rc = ensure_max_dynamic_smem(fn, ic->device, ic->max_smem_optin);
                    if (rc == TCHEM_OK) {
                        // small molecules: several blocks per SM (their shared memory allows it), one geometry per block at a time
                        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (size_t)ic->max_smem_optin / (smem2 + 1024)));
                        const unsigned grid = (unsigned)std::min<int64_t>(nb, (int64_t)ic->sm_count * per_sm);
                        double * o2 = d2ydr2 + b0 * (int64_t)nt * cd * cd;
                        if (n == 3 && cd * cd <= 96) {             // three atoms (H2O: the BASELINE SAPSet workload)
                            rc = ensure_max_dynamic_smem(reinterpret_cast<const void *>(&chain_cart_second_small_kernel<3, 3>), ic->device, ic->max_smem_optin);
                            if (rc != TCHEM_OK) break;
                            chain_cart_second_small_kernel<3, 3><<<grid, 32 * kCart2Warps, smem2, stream>>>(g_ws, H_ws, J_ws, K_ws, nb, nt, cd, o2);
                            count_launch("chain_cart_second_small_kernel");
                        } else if (n == 6 && cd * cd <= 160) {     // four atoms
                            rc = ensure_max_dynamic_smem(reinterpret_cast<const void *>(&chain_cart_second_small_kernel<6, 5>), ic->device, ic->max_smem_optin);
                            if (rc != TCHEM_OK) break;
                            chain_cart_second_small_kernel<6, 5><<<grid, 32 * kCart2Warps, smem2, stream>>>(g_ws, H_ws, J_ws, K_ws, nb, nt, cd, o2);
                            count_launch("chain_cart_second_small_kernel");
                        } else {
                            chain_cart_second_block_kernel<<<grid, 32 * kCart2Warps, smem2, stream>>>(g_ws, H_ws, J_ws, K_ws, nb, nt, n, cd, o2);
                            count_launch("chain_cart_second_block_kernel");
                        }
                    }
                } else {
                    const int64_t total = nb * nt * cd * cd;
                    const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)ic->sm_count * 8);
                    chain_cart_second_kernel<<<grid, 256, 0, stream>>>(g_ws, H_ws, J_ws, K_ws, nb, nt, n, cd, d2ydr2 + b0 * (int64_t)nt * cd * cd);
                    count_launch("chain_cart_second_kernel");
                }
            }
        }
        if (rc == TCHEM_OK && cudaGetLastError() != cudaSuccess) rc = set_error(TCHEM_ECUDA, "tchem_ic_sap_eval_cart: kernel launch failed");
    }
    cudaFreeAsync(ws, stream);
    return rc;
}

