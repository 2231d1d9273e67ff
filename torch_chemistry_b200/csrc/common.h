// error plumbing shared by the C-ABI translation units
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/tchem_b200.h"

namespace tchem_b200 {

int set_error(int code, const std::string & msg);          // stores msg thread-locally, returns code
void count_launch(const char * kernel = nullptr, int n = 1);   // counts a launch; `kernel` (a string literal) feeds tchem_kernel_log
// Raises the dynamic shared-memory limit of kernel `fn` to the device's opt-in maximum, once per
// (function, device). The attribute belongs to the FUNCTION, which every table shares: setting it to
// one table's need would lower it under another table's cached launch shape.
int ensure_max_dynamic_smem(const void * fn, int device, int max_smem_optin);
// The slice workspaces of the transforms come from the device's default stream-ordered pool (cudaMallocAsync). Its release
// threshold is 0 by default: every synchronisation hands the memory back to the driver and the next call maps it again
// (milliseconds for a 0.9 GB slice - seen as occasional 30 % slower steps). Raised once per device so that the pool keeps what the
// largest call needed; TCHEM_KEEP_WORKSPACE=0 leaves the pool alone.
void keep_async_pool(int device);

#define TC_CUDA(expr)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return ::tchem_b200::set_error(TCHEM_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

struct DeviceGuard {                                         // cudaSetDevice for the scope
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

} // namespace tchem_b200
