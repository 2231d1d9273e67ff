// B200 FP64 pipe: dependent-chain latency and per-SM throughput of DFMA, plus rsqrt / shuffle / LDS chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_latency tools/micro/fp64_latency.cu && /tmp/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double * out, int iters, long long * cycles) {
    double a[ILP];
    for (int i = 0; i < ILP; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], b, c);
    }
    const long long t1 = clock64();
    double s = 0.0;
    for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int ILP>
__global__ void dmma_chain(double * out, int iters, long long * cycles) {
    double c0[ILP], c1[ILP];
    for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = 1.0; }
    const double a = 1.0000001, b = 0.999999;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    const long long t1 = clock64();
    double s = 0.0;
    for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
__global__ void rsqrt_chain(double * out, int iters, long long * cycles) {
    double a = 2.0 + threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) a = rsqrt(a) + 2.0;
    const long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}
__global__ void shfl_chain(double * out, int iters, long long * cycles) {
    double a = 2.0 + threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
    const long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}
__global__ void lds_chain(double * out, int iters, long long * cycles) {
    __shared__ int next[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) next[i] = (i * 33 + 7) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) p = next[p];
    const long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}
int main() {
    double * out; long long * cyc, h;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    const int iters = 4096;
#define RUN(label, kern, grid, block, ops) kern<<<grid, block>>>(out, iters, cyc); cudaDeviceSynchronize(); \
    kern<<<grid, block>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-58s %8.2f cycles per iteration (%d ops)\n", label, (double)h / iters, ops);
    RUN("DFMA dependent chain, 1 warp", dfma_chain<1>, 1, 32, 1)
    RUN("DFMA 2 independent chains, 1 warp", dfma_chain<2>, 1, 32, 2)
    RUN("DFMA 4 independent chains, 1 warp", dfma_chain<4>, 1, 32, 4)
    RUN("DFMA 8 independent chains, 1 warp", dfma_chain<8>, 1, 32, 8)
    RUN("DFMA 16 independent chains, 1 warp", dfma_chain<16>, 1, 32, 16)
    RUN("DFMA 8 chains, 4 warps (1 per SMSP)", dfma_chain<8>, 1, 128, 8)
    RUN("DFMA 8 chains, 8 warps", dfma_chain<8>, 1, 256, 8)
    RUN("DFMA 8 chains, 16 warps", dfma_chain<8>, 1, 512, 8)
    RUN("DFMA 8 chains, 32 warps", dfma_chain<8>, 1, 1024, 8)
    RUN("DFMA 8 chains, 32 warps x 148 blocks", dfma_chain<8>, 148, 1024, 8)
    RUN("DMMA m8n8k4 dependent chain, 1 warp", dmma_chain<1>, 1, 32, 1)
    RUN("DMMA 2 independent accumulators, 1 warp", dmma_chain<2>, 1, 32, 2)
    RUN("DMMA 4 independent accumulators, 1 warp", dmma_chain<4>, 1, 32, 4)
    RUN("DMMA 8 independent accumulators, 1 warp", dmma_chain<8>, 1, 32, 8)
    RUN("DMMA 4 accumulators, 4 warps (1 per SMSP)", dmma_chain<4>, 1, 128, 4)
    RUN("DMMA 4 accumulators, 8 warps", dmma_chain<4>, 1, 256, 4)
    RUN("DMMA 4 accumulators, 16 warps", dmma_chain<4>, 1, 512, 4)
    RUN("rsqrt(double) + add dependent chain", rsqrt_chain, 1, 32, 1)
    RUN("shfl (double) dependent chain", shfl_chain, 1, 32, 1)
    RUN("LDS dependent chain (pointer chase)", lds_chain, 1, 32, 1)
    return 0;
}
