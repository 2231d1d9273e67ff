// error plumbing shared by the C-ABI translation units
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/tchem_b200.h"

namespace tchem_b200 {

int set_error(int code, const std::string & msg);          // stores msg thread-locally, returns code
void count_launch(int n = 1);

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
