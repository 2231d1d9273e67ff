// Ceiling of the HBM write stream for the tile write patterns of the q+J kernels (no arithmetic):
//   linear : every warp streams through its tile's contiguous 32 * gdoubles region
//   rows   : for each group of `rows` rows (piece = rows * cartdim doubles) the warp walks the 32
//            geometries of its tile, one 256-byte warp store at a time (what the interpreter does)
// usage: write_pattern cartdim intdim B warps_per_sm
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void k_linear(double * out, long ntiles, long tile_doubles) {
    const int lane = threadIdx.x & 31;
    const long w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long)gridDim.x * (blockDim.x >> 5);
    for (long t = w; t < ntiles; t += nw) {
        double2 * p = reinterpret_cast<double2 *>(out + t * tile_doubles);
        for (long e = lane; e < tile_doubles / 2; e += 32) p[e] = make_double2(0.0, 1.0);
    }
}
__global__ void k_rows(double * out, long ntiles, int gdoubles, int piece) {
    const int lane = threadIdx.x & 31;
    const long w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long)gridDim.x * (blockDim.x >> 5);
    for (long t = w; t < ntiles; t += nw) {
        double * base = out + t * 32L * gdoubles;
        for (int r0 = 0; r0 < gdoubles; r0 += piece) {
            const int len = min(piece, gdoubles - r0);
            for (int e0 = 0; e0 < len; e0 += 32) {
                if (e0 + lane < len) {
                    double * o = base + r0 + e0 + lane;
#pragma unroll 8
                    for (int g = 0; g < 32; g++) o[(long)g * gdoubles] = 1.0;
                }
            }
        }
    }
}
// linear per tile with 32-byte stores (st.global.v4.f64, sm_100) and with streaming / write-through hints
template <int KIND>
__global__ void k_linear_x(double * out, long ntiles, long tile_doubles) {
    const int lane = threadIdx.x & 31;
    const long w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long)gridDim.x * (blockDim.x >> 5);
    for (long t = w; t < ntiles; t += nw) {
        double * p = out + t * tile_doubles;
        if (KIND == 0) {
            for (long e = 4 * lane; e + 3 < tile_doubles; e += 128)
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p + e), "d"(0.0), "d"(1.0), "d"(2.0), "d"(3.0) : "memory");
        } else if (KIND == 1) {
            for (long e = 2 * lane; e + 1 < tile_doubles; e += 64) __stcs(reinterpret_cast<double2 *>(p + e), make_double2(0.0, 1.0));
        } else if (KIND == 2) {
            for (long e = 2 * lane; e + 1 < tile_doubles; e += 64) __stwt(reinterpret_cast<double2 *>(p + e), make_double2(0.0, 1.0));
        } else {
            for (long e = 2 * lane; e + 1 < tile_doubles; e += 64) __stcg(reinterpret_cast<double2 *>(p + e), make_double2(0.0, 1.0));
        }
    }
}
// rows pattern with 16-byte stores (lane = 2 adjacent elements)
__global__ void k_rows16(double * out, long ntiles, int gdoubles, int piece) {
    const int lane = threadIdx.x & 31;
    const long w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long)gridDim.x * (blockDim.x >> 5);
    for (long t = w; t < ntiles; t += nw) {
        double * base = out + t * 32L * gdoubles;
        for (int r0 = 0; r0 < gdoubles; r0 += piece) {
            const int len = min(piece, gdoubles - r0);
            for (int e0 = 0; e0 < len; e0 += 64) {
                if (e0 + 2 * lane < len) {
                    double * o = base + r0 + e0 + 2 * lane;
#pragma unroll 8
                    for (int g = 0; g < 32; g++) *reinterpret_cast<double2 *>(o + (long)g * gdoubles) = make_double2(1.0, 2.0);
                }
            }
        }
    }
}
// geometry-major inside a piece: all elements of the piece for geometry g, then g + 1
template <int VEC>
__global__ void k_gmajor(double * out, long ntiles, int gdoubles, int piece) {
    const int lane = threadIdx.x & 31;
    const long w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long)gridDim.x * (blockDim.x >> 5);
    for (long t = w; t < ntiles; t += nw) {
        double * base = out + t * 32L * gdoubles;
        for (int r0 = 0; r0 < gdoubles; r0 += piece) {
            const int len = min(piece, gdoubles - r0);
            for (int g = 0; g < 32; g++) {
                double * o = base + (long)g * gdoubles + r0;
                if (VEC == 2) { for (int e = 2 * lane; e < len; e += 64) *reinterpret_cast<double2 *>(o + e) = make_double2(1.0, 2.0); }
                else          { for (int e = lane; e < len; e += 32) o[e] = 1.0; }
            }
        }
    }
}

// bulk (TMA) stores: every lane owns one geometry of the tile and issues one
// cp.async.bulk.global.shared::cta of the piece (rows * cartdim doubles) from the warp's shared
// staging rows; REFILL = the warp rewrites the staging rows (16-byte shared stores) before every piece
template <bool REFILL>
__global__ void k_bulk(double * out, long ntiles, int gdoubles, int piece, int pitch_doubles) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double * S = sm + (size_t)warp * 32 * pitch_doubles;
    for (int e = 0; e < piece; e++) S[lane * pitch_doubles + e] = 1.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const long w = (long)blockIdx.x * (blockDim.x >> 5) + warp, nw = (long)gridDim.x * (blockDim.x >> 5);
    const unsigned src = (unsigned)__cvta_generic_to_shared(S + lane * pitch_doubles);
    for (long t = w; t < ntiles; t += nw) {
        double * base = out + t * 32L * gdoubles + (long)lane * gdoubles;
        for (int r0 = 0; r0 < gdoubles; r0 += piece) {
            const int len = min(piece, gdoubles - r0);
            if (REFILL) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                for (int e = 0; e < len; e += 2) *reinterpret_cast<double2 *>(S + lane * pitch_doubles + e) = make_double2(1.0, (double)r0);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
            }
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + r0), "r"(src), "r"(len * 8) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main(int argc, char ** argv) {
    const int cartdim = atoi(argv[1]), intdim = atoi(argv[2]);
    const long B = atol(argv[3]);
    const int wps = atoi(argv[4]);
    const int gdoubles = cartdim * intdim;
    const long ntiles = B / 32;
    double * out;
    cudaMalloc(&out, sizeof(double) * B * gdoubles);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double gb = 8.0 * B * gdoubles / 1e9;
    auto run = [&](const char * name, auto launch) {
        for (int i = 0; i < 3; i++) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 10; i++) launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s %8.4f ms  %7.1f GB/s\n", name, ms / 10, gb / (ms / 10) * 1e3);
    };
    run("cudaMemset", [&] { cudaMemsetAsync(out, 0, sizeof(double) * B * gdoubles); });
    run("linear per tile", [&] { k_linear<<<148, wps * 32>>>(out, ntiles, 32L * gdoubles); });
    run("linear v4.f64 (32 B)", [&] { k_linear_x<0><<<148, wps * 32>>>(out, ntiles, 32L * gdoubles); });
    run("linear .cs", [&] { k_linear_x<1><<<148, wps * 32>>>(out, ntiles, 32L * gdoubles); });
    run("linear .wt", [&] { k_linear_x<2><<<148, wps * 32>>>(out, ntiles, 32L * gdoubles); });
    run("linear .cg", [&] { k_linear_x<3><<<148, wps * 32>>>(out, ntiles, 32L * gdoubles); });
    run("linear 2x148 blocks", [&] { k_linear<<<296, wps * 16>>>(out, ntiles, 32L * gdoubles); });
    run("linear 1184 blocks x 64 thr", [&] { k_linear<<<1184, 64>>>(out, ntiles, 32L * gdoubles); });
    for (int rows : {4, intdim}) {
        char name[64]; snprintf(name, 64, "rows=%d (piece %d B)", rows, rows * cartdim * 8);
        run(name, [&] { k_rows<<<148, wps * 32>>>(out, ntiles, gdoubles, rows * cartdim); });
    }
    for (int rows : {1, 4, 16}) {
        char name[64]; snprintf(name, 64, "rows16=%d (piece %d B)", rows, rows * cartdim * 8);
        run(name, [&] { k_rows16<<<148, wps * 32>>>(out, ntiles, gdoubles, rows * cartdim); });
    }
    for (int rows : {1, 4, 16}) {
        char name[64]; snprintf(name, 64, "gmajor8 rows=%d", rows);
        run(name, [&] { k_gmajor<1><<<148, wps * 32>>>(out, ntiles, gdoubles, rows * cartdim); });
        snprintf(name, 64, "gmajor16 rows=%d", rows);
        run(name, [&] { k_gmajor<2><<<148, wps * 32>>>(out, ntiles, gdoubles, rows * cartdim); });
    }

    for (int rows : {1, 2, 4, 16}) {
        if (rows > intdim) continue;
        const int piece = rows * cartdim;
        if (piece % 2) continue;
        int pitch = piece + 2;
        if ((pitch / 2) % 2 == 0) pitch += 2;       // pitch / 16 B odd: conflict-free 16-byte staging stores
        const size_t smem = (size_t)wps * 32 * pitch * 8;
        if (smem > 220 * 1024) continue;
        cudaFuncSetAttribute(k_bulk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_bulk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        char name[64]; snprintf(name, 64, "bulk rows=%d (%d B)", rows, piece * 8);
        run(name, [&] { k_bulk<false><<<148, wps * 32, smem>>>(out, ntiles, gdoubles, piece, pitch); });
        snprintf(name, 64, "bulk+refill rows=%d", rows);
        run(name, [&] { k_bulk<true><<<148, wps * 32, smem>>>(out, ntiles, gdoubles, piece, pitch); });
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
