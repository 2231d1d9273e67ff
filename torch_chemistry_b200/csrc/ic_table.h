// host-side compiled IntCoordSet
#pragma once
#include <mutex>
#include <vector>

#include "common.h"
#include "ic_program.h"

struct tchem_ic_table {
    int device = 0;
    int intdim = 0, cartdim = 0, natoms = 0;
    int sm_count = 0;
    bool has5 = false;            // any 5-atom (torsion2 family) term
    int max_smem_optin = 0;
    int max_smem_sm = 0;          // shared memory per SM
    std::vector<tchem_invdisp_t> terms;
    tchem_b200::Program prog[tchem_b200::MODE_COUNT];       // device pointers
    void * dev_blob[tchem_b200::MODE_COUNT] = {nullptr, nullptr, nullptr};
    // launch shape of ic_eval_kernel per mode, resolved on first use (function attribute and
    // occupancy queries cost microseconds each: too much for 10^4-geometry calls)
    struct LaunchCfg { const void * fn = nullptr; int warps = 0, per_sm = 0; size_t smem = 0; };
    int group_warps = 0;          // > 0: the q+J program was built for the warp-group kernel (G warps share a tile)
    mutable LaunchCfg cfg[tchem_b200::MODE_COUNT];
    // host pipeline state for the *_host entry points (lazily created)
    struct HostPipe * pipe = nullptr;
    std::mutex pipe_mutex;        // serialises host-buffer calls on THIS table only (they share the pipe's slots)
};

namespace tchem_b200 {

struct HostProgram {
    std::vector<PrimDesc> prims;
    std::vector<ChunkDesc> chunks;
    std::vector<uint32_t> map;
    std::vector<ElemMap> emap;
    std::vector<Src> qsrcs;
    std::vector<uint32_t> psrcs;          // packed J / K sources, quads
    std::vector<double> coefs;            // distinct coefficients
    std::vector<ZeroRun> runs;
    int cap = 0;
};

// Appends the element map of one dense span (lists[k] = sources of element k) to emap/srcs/runs,
// fills `span`; false when a list is too long for the encoding. Elements are grouped in
// sectors of 4 (aligned to `abs_off`, the span's element offset inside the geometry block, when
// the block stride keeps sectors aligned) and the sectors are ordered by source count.
// zero_sectors: whole sectors of structural zeros may become single entries (kZeroSectorWord; K spans,
// whose write-out is compiled with ZSEC).
bool emit_span(const std::vector<std::vector<Src>> & lists, int64_t abs_off, int64_t gstride,
               std::vector<ElemMap> & emap, std::vector<uint32_t> & psrcs, std::vector<double> & coefs,
               std::vector<ZeroRun> & runs, SpanDesc & span, bool zero_sectors = false);

// pure host: flatten terms into a program for one mode
// cap_hint > 0: compact capacity to aim for (at least the largest primitive), 0: default policy;
// max_rows > 0: at most that many rows per chunk
int build_program(const std::vector<tchem_invdisp_t> & terms, int n_ic, int cartdim, int mode,
                  HostProgram & out, int cap_hint = 0, int max_rows = 0);

} // namespace tchem_b200
