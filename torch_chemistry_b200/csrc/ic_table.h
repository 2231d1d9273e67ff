// host-side compiled IntCoordSet
#pragma once
#include <vector>

#include "common.h"
#include "ic_program.h"

struct tchem_ic_table {
    int device = 0;
    int intdim = 0, cartdim = 0, natoms = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    std::vector<tchem_invdisp_t> terms;
    tchem_b200::Program prog[tchem_b200::MODE_COUNT];       // device pointers
    void * dev_blob[tchem_b200::MODE_COUNT] = {nullptr, nullptr, nullptr};
    // host pipeline state for the *_host entry points (lazily created)
    struct HostPipe * pipe = nullptr;
};

namespace tchem_b200 {

struct HostProgram {
    std::vector<PrimDesc> prims;
    std::vector<ChunkDesc> chunks;
    std::vector<uint32_t> map;
    std::vector<Src> srcs;
    int cap = 0;
};

// pure host: flatten terms into a program for one mode
int build_program(const std::vector<tchem_invdisp_t> & terms, int n_ic, int cartdim, int mode,
                  HostProgram & out);

} // namespace tchem_b200
