// host-side compiled SAPSet (shared by sap.cu and jit.cu)
#pragma once
#include <mutex>
#include <utility>
#include <vector>

#include "common.h"
#include "ic_program.h"

namespace tchem_b200 {

// distinct coordinate of a monomial: x[col]^power; e_pow / e_powm1 = entries of the per-lane power
// table holding x[col]^power and x[col]^(power-1)
struct SapUniq { int32_t col, power, e_pow, e_powm1; };
struct SapTerm { int32_t ubegin, uend; };         // distinct coordinates of one monomial

struct SapProgram {                               // device pointers
    const SapTerm *   terms;
    const SapUniq *   uniqs;
    const ChunkDesc * chunks;                     // prim_begin/end = term range, row0/nrows likewise
    const ElemMap *   emap;
    const uint4 *     psrcs;
    const double *    coefs;
    const ZeroRun *   runs;
    const int32_t *   pow_off;                    // power table: entries pow_off[col] + p hold x[col]^p, p = 0..maxpow
    int32_t nterms, nuniqs, nchunks, cap, sumdim, npow;
};

int ensure_program(tchem_ic_table * t, int mode);

} // namespace tchem_b200

struct SapHostPipe {                     // device staging of the chained host-buffer entry point
    static constexpr int kSlots = 3;
    cudaStream_t stream[kSlots] = {};
    double * buf[kSlots] = {};
    size_t cap = 0;
    bool ok = false;
    std::mutex mutex;                    // host-buffer calls on one SAPSet share the slots: one at a time
};

struct tchem_sap_table {
    SapHostPipe pipe;
    int device = 0;
    int nterms = 0, sumdim = 0, irreducible = 0;
    int sm_count = 0, max_smem_optin = 0;
    std::vector<int32_t> dims;
    tchem_b200::SapProgram prog;
    void * dev_blob = nullptr;
    // host copy of the monomials as (column in the concatenated x, power) lists, for the generator
    std::vector<std::vector<std::pair<int, int>>> h_terms;
};

