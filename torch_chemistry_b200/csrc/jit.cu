// Table-specialised q + J kernels.
//
// The generic kernel (ic_eval.cu) INTERPRETS the compiled definition table: every warp walks the
// primitive list, switches on the type, looks atoms and gather sources up at run time. With 12
// resident warps per SM that loop is latency bound. For large batches of a small-to-medium set the
// table is instead turned into straight-line CUDA C++ - atom indices, coefficients and the sparsity
// of every J row become compile-time constants, the hand-written device functions of ic_math.cuh
// are called back to back so their independent arithmetic interleaves (ILP), common bond vectors
// are shared by the compiler - and compiled once per table for sm_100a with NVRTC.
//
// Same warp-tile scheme as the interpreter: lane = geometry, cp.async staging of the coordinates
// with prefetch of the next tile, dense row chunks staged [element][lane] in shared memory and
// written out transposed (coalesced 256-byte segments). Replaces IntCoordSet::compute_IC_J
// (reference source/IC/IntCoordSet.cpp:147-159) for tables without the 5-atom torsion2 family.
#include <cuda.h>          // CUtensorMap types only: the encoder is fetched with cudaGetDriverEntryPoint
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include "ic_table.h"
#include "sap_table.h"

namespace tchem_b200 {

namespace {

const char kMathHeader[] =
#include "ic_math_embedded.inc"
    ;

// ---- NVRTC through dlopen: the library itself does not link against it ----------------------
typedef struct _nvrtcProgram * nvrtcProgram;
struct Nvrtc {
    void * lib = nullptr;
    int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char * const *, const char * const *) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char * const *) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    int (*DestroyProgram)(nvrtcProgram *) = nullptr;
    bool ok = false;
    Nvrtc() {
        const char * names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
        for (const char * n : names) { lib = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (lib) break; }
        if (!lib) return;
#define TC_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(lib, name)); if (!field) return;
        TC_SYM(CreateProgram, "nvrtcCreateProgram")
        TC_SYM(CompileProgram, "nvrtcCompileProgram")
        TC_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
        TC_SYM(GetCUBIN, "nvrtcGetCUBIN")
        TC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
        TC_SYM(GetProgramLog, "nvrtcGetProgramLog")
        TC_SYM(DestroyProgram, "nvrtcDestroyProgram")
#undef TC_SYM
        ok = true;
    }
};
Nvrtc & nvrtc() { static Nvrtc n; return n; }

// monomials per chunk of the chained kernel: as many as fit `dcap` staged entries, and - when the set
// needs several chunks - a multiple of 4, so that every chunk's y (8 bytes per term) and J pieces end
// on 32-byte sector boundaries (a half-written sector costs a DRAM read-modify-write)
int chain_terms_per_chunk(int nt, int per_term, int dcap) {
    const int nchunks = (nt * per_term + dcap - 1) / dcap;
    int tper = (nt + nchunks - 1) / nchunks;
    if (nchunks > 1 && tper >= 4) tper &= ~3;
    return tper;
}

// staging row of element e of a span (mirrors staged_row<SPAN> in the prelude)
int stage_row(int e, int span) { return span % 2 == 0 ? (e >> 1) + (e & 1) * (span / 2) : e; }

std::string hexd(double v) {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%a", v);
    return buf;
}

const char kPrelude[] = R"TCHEM(
typedef long long int64_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
#include "ic_math.cuh"
using namespace tchem_b200;
#define KPAD 33
#define LANES 32
// global r[b0 .. b0+ntile)[CD] -> R[c * KPAD + lane] with cp.async (see ic_eval.cuh::stage_issue)
__device__ __forceinline__ void stage_issue(const double * __restrict__ r, int64_t b0, int ntile, double * R, int lane) {
    const double * src = r + b0 * CD;
    const int total = ntile * CD;
    int g = lane / CD, c = lane - g * CD;
    const int dg = LANES / CD, dc = LANES - dg * CD;
    for (int f = lane; f < total; f += LANES) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(R + c * KPAD + g);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src + f) : "memory");
        g += dg; c += dc;
        if (c >= CD) { c -= CD; g++; }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
// the same copy issued by all `nthreads` threads of a block (thread t), one commit per thread
__device__ __forceinline__ void stage_issue_block(const double * __restrict__ r, int64_t b0, int ntile, double * R, int t, int nthreads) {
    const double * src = r + b0 * CD;
    const int total = ntile * CD;
    int g = t / CD, c = t - g * CD;
    const int dg = nthreads / CD, dc = nthreads - dg * CD;
    for (int f = t; f < total; f += nthreads) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(R + c * KPAD + g);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src + f) : "memory");
        g += dg; c += dc;
        if (c >= CD) { c -= CD; g++; }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void stage_finish(int ntile, double * R, int lane) {
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
    if (ntile < LANES && lane >= ntile)
        for (int k = 0; k < CD; k++) R[k * KPAD + lane] = R[k * KPAD + ntile - 1];
    __syncwarp();
}
// staged [span][lane] -> out[(b0 + g) * gstride + k]: lanes along k, coalesced segments; two
// adjacent elements per lane and 16-byte stores when the span, the stride and the pointer are even
// Row of element e in the staging buffer: when the span is even, elements (2p, 2p+1) - one 16-byte
// store - sit in rows p and p + SPAN/2, so that the lanes of a store read rows one apart (stride 33
// doubles, conflict-free) instead of two apart (2-way bank conflicts). The generator stages with
// the same permutation (stage_row in jit.cu).
template <int SPAN>
__device__ __forceinline__ constexpr int staged_row(int e) { return SPAN % 2 == 0 ? (e >> 1) + (e & 1) * (SPAN / 2) : e; }
template <int SPAN, int GSTRIDE>
__device__ __forceinline__ void copy_out(const double * D, double * __restrict__ out, int ntile, int lane) {
    // geometry after geometry, lanes running over the (geometry, element pair) pairs: consecutive
    // stores continue where the previous one ended. (Walking the 32 geometries element step by
    // element step instead caps the HBM write stream: tools/micro/write_pattern.cu.)
    if (SPAN % 2 == 0 && GSTRIDE % 2 == 0 && (reinterpret_cast<uint64_t>(out) & 15) == 0) {
        constexpr int PAIRS = SPAN / 2, DG = LANES / PAIRS, DP = LANES - DG * PAIRS;
        const int total = ntile * PAIRS;
        int g = lane / PAIRS, p = lane - g * PAIRS;
#pragma unroll 4
        for (int f = lane; f < total; f += LANES) {
            const double * s0 = D + p * KPAD + g;
            *reinterpret_cast<double2 *>(out + (int64_t)g * GSTRIDE + 2 * p) = make_double2(s0[0], s0[PAIRS * KPAD]);
            g += DG; p += DP;
            if (p >= PAIRS) { p -= PAIRS; g++; }
        }
        return;
    }
    constexpr int DG = LANES / SPAN, DK = LANES - DG * SPAN;
    const int total = ntile * SPAN;
    int g = lane / SPAN, k = lane - g * SPAN;
    for (int f = lane; f < total; f += LANES) {
        out[(int64_t)g * GSTRIDE + k] = D[staged_row<SPAN>(k) * KPAD + g];
        g += DG; k += DK;
        if (k >= SPAN) { k -= SPAN; g++; }
    }
}
// tensor map (CUtensorMap, 128 opaque bytes) of the bulk-store variant
struct __align__(64) TMap { unsigned long long opaque[16]; };
// 16-byte store at byte offset `off` of a box laid out with the TMA unit's 64B swizzle
// (16-byte unit index bits 0-1 ^= address bits 7-8)
__device__ __forceinline__ void stage_pair(char * box, int off, double a, double b) {
    *reinterpret_cast<double2 *>(box + (off ^ (((off >> 7) & 3) << 4))) = make_double2(a, b);
}
// q rows of a chunk, (geometry, row)-flattened
template <int NROWS, int GSTRIDE>
__device__ __forceinline__ void copy_rows(const double * Q, double * __restrict__ out, int ntile, int lane) {
    const int total = ntile * NROWS;
    int g = lane / NROWS, row = lane - g * NROWS;
    const int dg = LANES / NROWS, dr = LANES - dg * NROWS;
    for (int f = lane; f < total; f += LANES) {
        out[(int64_t)g * GSTRIDE + row] = Q[row * KPAD + g];
        g += dg; row += dr;
        if (row >= NROWS) { row -= NROWS; g++; }
    }
}
)TCHEM";

struct JitPrim { int type, natoms, atoms[5]; double min; };

} // namespace

namespace {

// distinct primitives + per-row (coefficient, primitive) lists of a definition; false: not eligible
bool collect(const std::vector<tchem_invdisp_t> & terms, int n_ic, std::vector<JitPrim> & prims,
             std::vector<std::vector<std::pair<double, int>>> & rows) {
    typedef std::tuple<int, int, int, int, int, int, double> Key;
    std::map<Key, int> index;
    rows.assign(n_ic, {});
    for (const auto & d : terms) {
        if (d.natoms == 5) return false;
        int a[5];
        for (int i = 0; i < 5; i++) a[i] = i < d.natoms ? d.atoms[i] : -1;
        const double mn = d.type == TCHEM_TORSION ? d.min : 0.0;
        Key k(d.type, a[0], a[1], a[2], a[3], a[4], mn);
        auto f = index.find(k);
        int id;
        if (f == index.end()) {
            id = (int)prims.size();
            index[k] = id;
            JitPrim p;
            p.type = d.type; p.natoms = d.natoms; p.min = d.min;
            for (int i = 0; i < 5; i++) p.atoms[i] = a[i];
            prims.push_back(p);
        } else id = f->second;
        rows[d.ic].push_back({d.coeff, id});
    }
    return true;
}

// `double q<p>; V3<double> j<p>[n];` + the call that fills them (want_J false: the gradient
// blocks are dropped on the floor and the compiler removes their arithmetic)
bool emit_prim(std::ostringstream & o, const JitPrim & pr, size_t p, bool want_J, const char * indent) {
    o << indent << "double q" << p << " = 1.0;";
    if (pr.natoms == 0) { o << "\n"; return true; }
    if (want_J) o << " V3<double> j" << p << "[" << pr.natoms << "];";
    o << "\n" << indent << "{ const V3<double> A[" << pr.natoms << "] = {";
    for (int i = 0; i < pr.natoms; i++) o << (i ? ", " : "") << "a[" << pr.atoms[i] << "]";
    if (want_J) o << "}; auto emit = [&](int i, const V3<double> & v) { j" << p << "[i] = v; };";
    else        o << "}; auto emit = [&](int, const V3<double> &) {};";
    switch (pr.type) {
    case TCHEM_STRETCHING: o << " stretching_J<double>(A, q" << p << ", emit);"; break;
    case TCHEM_BENDING:    o << " double c; bending_J<double, false>(A, c, emit); q" << p << " = acos(c);"; break;
    case TCHEM_COSBENDING: o << " bending_J<double, true>(A, q" << p << ", emit);"; break;
    case TCHEM_TORSION:    o << " double sp = 0.0, cp = 0.0, st = 0.0; torsion_J<double, 0>(A, sp, cp, st, emit); q" << p
                             << " = signed_wrapped_angle(cp, st, " << hexd(pr.min) << ");"; break;
    case TCHEM_SINTORSION: o << " double sp = 0.0, cp = 0.0, st = 0.0; torsion_J<double, 1>(A, sp, cp, st, emit); q" << p << " = sp;"; break;
    case TCHEM_COSTORSION: o << " double sp = 0.0, cp = 0.0, st = 0.0; torsion_J<double, 2>(A, sp, cp, st, emit); q" << p << " = cp;"; break;
    case TCHEM_OUTOFPLANE: o << " double s; oop_J<double, false>(A, s, emit); q" << p << " = asin(s);"; break;
    case TCHEM_SINOOP:     o << " oop_J<double, true>(A, q" << p << ", emit);"; break;
    default: return false;
    }
    o << " }\n";
    return true;
}

// kernel signature up to and including the per-tile prologue (coordinates staged, atoms in
// registers, next tile's copy in flight); `params` = the kernel's pointer parameters after r, B
// plain_cols > 0: the input rows are used as they are (x0p1 .. x{plain_cols-1}p1), no atoms / primitives
// box_doubles > 0: the per-warp region starts with a 16-byte aligned box of that many doubles (TMA bulk-store
// staging), then R, then D, each padded to an even number of doubles (chain_warp_doubles)
size_t chain_warp_doubles(int cartdim, int dcap_entries, int box_doubles) {
    return (size_t)box_doubles + ((cartdim * 33 + 1) & ~1) + ((dcap_entries * 33 + 1) & ~1);
}
void emit_prologue(std::ostringstream & o, const char * name, const char * params, int cartdim, int dcap_entries,
                   int warps, int blocks, const std::vector<JitPrim> & prims, int plain_cols = 0, int box_doubles = 0) {
    const int natoms = cartdim / 3;
    o << "#define CD " << cartdim << "\n#define DCAP " << dcap_entries << "\n";
    if (box_doubles > 0)
        o << "#define BOXD " << box_doubles << "\n#define RSIZE " << ((cartdim * 33 + 1) & ~1) << "\n#define WARPD "
          << chain_warp_doubles(cartdim, dcap_entries, box_doubles) << "\n";
    o << kPrelude;
    o << "extern \"C\" __global__ void __launch_bounds__(" << warps * 32 << ", " << blocks << ")\n"
      << name << "(const double * __restrict__ r, const long long B, " << params << ") {\n"
      << "    extern __shared__ __align__(16) double smem[];\n"
      << "    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;\n";
    if (box_doubles > 0)
        // block tile = 32 * nwarps consecutive geometries: the warps stage adjacent 32-row slices of ONE box per
        // result array, so the block's results are contiguous in global memory and leave in one bulk store each
        o << "    double * Box = smem;                                       // [nwarps * BOXD], shared by the block\n"
          << "    double * R = smem + (size_t)nwarps * BOXD + (size_t)warp * (WARPD - BOXD);\n"
          << "    double * D = R + RSIZE;\n"
          << "    // structural zeros never change: staged once, the box keeps them from tile to tile\n"
          << "    for (int e = 2 * threadIdx.x; e < nwarps * BOXD; e += 2 * blockDim.x) *reinterpret_cast<double2 *>(Box + e) = make_double2(0.0, 0.0);\n"
          << "    __syncthreads();\n"
          << "    const int64_t ntiles = (B + LANES * nwarps - 1) / (LANES * nwarps), step = gridDim.x;     // block tiles\n"
          << "    int64_t tile = blockIdx.x;\n"
          << "    if (tile < ntiles) { const int64_t b0 = (tile * nwarps + warp) * LANES; if (b0 < B) stage_issue(r, b0, (int)min((int64_t)LANES, B - b0), R, lane); }\n"
          << "    for (; tile < ntiles; tile += step) {\n"
          << "        const int64_t b0 = (tile * nwarps + warp) * LANES;\n"
          << "        // a warp past the end of the batch (ntile <= 0) runs on stale coordinates and only keeps the barriers\n"
          << "        const int ntile = (int)max((int64_t)0, min((int64_t)LANES, B - b0));\n"
          << "        stage_finish(max(ntile, 1), R, lane);\n";
    else
        o << "    double * R = smem + (size_t)warp * (CD + DCAP) * KPAD;\n"
          << "    double * D = R + CD * KPAD;\n"
          << "    const int64_t ntiles = (B + LANES - 1) / LANES, step = (int64_t)gridDim.x * nwarps;\n"
          << "    int64_t tile = (int64_t)blockIdx.x * nwarps + warp;\n"
          << "    if (tile < ntiles) { const int64_t b0 = tile * LANES; stage_issue(r, b0, (int)min((int64_t)LANES, B - b0), R, lane); }\n"
          << "    for (; tile < ntiles; tile += step) {\n"
          << "        const int64_t b0 = tile * LANES;\n"
          << "        const int ntile = (int)min((int64_t)LANES, B - b0);\n"
          << "        stage_finish(ntile, R, lane);\n";
    if (plain_cols > 0) {
        for (int i = 0; i < plain_cols; i++) o << "        const double x" << i << "p1 = R[" << i << " * KPAD + lane];\n";
    } else {
        o << "        V3<double> a[" << natoms << "];\n";
    }
    std::vector<char> atom_used(natoms, 0);
    for (const auto & p : prims) for (int i = 0; i < p.natoms; i++) atom_used[p.atoms[i]] = 1;
    for (int at = 0; at < natoms; at++)
        if (atom_used[at])
            o << "        a[" << at << "] = v3<double>(R[" << 3 * at << " * KPAD + lane], R[" << 3 * at + 1 << " * KPAD + lane], R["
              << 3 * at + 2 << " * KPAD + lane]);\n";
    o << "        __syncwarp();\n"
      << "        // the coordinates now live in registers: the next tile's copy can start\n";
    if (box_doubles > 0)
        o << "        if (tile + step < ntiles) { const int64_t nb0 = ((tile + step) * nwarps + warp) * LANES; if (nb0 < B) stage_issue(r, nb0, (int)min((int64_t)LANES, B - nb0), R, lane); }\n";
    else
        o << "        if (tile + step < ntiles) { const int64_t nb0 = (tile + step) * LANES; stage_issue(r, nb0, (int)min((int64_t)LANES, B - nb0), R, lane); }\n";
}

std::string row_expr(const std::vector<std::pair<double, int>> & row, const char * var) {
    std::string e;
    for (auto & t : row) {
        const std::string term = hexd(t.first) + " * " + var + std::to_string(t.second);
        e = e.empty() ? term : "fma(" + hexd(t.first) + ", " + var + std::to_string(t.second) + ", " + e + ")";
    }
    return e;
}

} // namespace

// Chained r -> q -> SAPSet: straight-line q rows (compute_IC_J value formulas, gradients elided),
// then every monomial and its partial derivatives as explicit products of shared powers. Staged
// per chunk of monomials: [values][Jacobian rows], written out like the q+J kernel.
// `bulk` (single-chunk sets with even NT and NT * SD): y and the Jacobian are staged per lane in the
// layout they have in global memory, so the tile's 32 geometries are one contiguous block of each
// array and one lane hands them to the TMA unit as two 1-D bulk copies (cp.async.bulk.global.shared::cta)
// - no copy-out instructions, a write front that is linear per warp.
bool chain_bulk_eligible(int nt, int sd, int dcap) {
    return nt % 2 == 0 && (nt * sd) % 2 == 0 && nt * (1 + sd) <= dcap;
}
// `cart` (chained only): the Jacobian goes out in CARTESIAN coordinates, d SAP / d r = (d SAP / d q) J [NT][CD] - the product
// callers of the reference form by hand (test/normal_mode/main.cpp:56-61) - with J's rows emitted as literal combinations of
// the primitives' analytic gradient blocks and every structural zero skipped.
std::string jit_chain_source(const std::vector<tchem_invdisp_t> & terms, int n_ic, int cartdim,
                             const std::vector<std::vector<std::pair<int, int>>> & saps, int dcap, int warps, int blocks, bool bulk = false,
                             bool cart = false) {
    // terms empty: the SAPSet alone, evaluated on given coordinates x[B][n_ic] (cartdim = n_ic)
    const bool plain = terms.empty();
    std::vector<JitPrim> prims;
    std::vector<std::vector<std::pair<double, int>>> rows;
    if (!plain && !collect(terms, n_ic, prims, rows)) return std::string();
    const int nt = (int)saps.size(), sd = n_ic;
    if (cart && plain) return std::string();
    const int sdo = cart ? cartdim : sd;              // row length of the Jacobian that leaves the kernel
    const int per_term = 1 + sdo;
    if (per_term > dcap) return std::string();
    const int tper = chain_terms_per_chunk(nt, per_term, dcap);
    // (cart + bulk: the box holds [32][NT] values and [32][NT * CD] Cartesian rows per warp; dcap only sizes the q staging)
    if (bulk && !(cart ? nt % 2 == 0 && (nt * sdo) % 2 == 0 : chain_bulk_eligible(nt, sd, dcap))) return std::string();
    std::ostringstream o;
    o << "#define NT " << nt << "\n#define SD " << sd << "\n#define SDO " << sdo << "\n";
    if (bulk)
        emit_prologue(o, cart ? "tchem_jit_chain_cart_bulk" : "tchem_jit_chain_bulk", "double * __restrict__ q, double * __restrict__ y, double * __restrict__ J",
                      cartdim, sd, warps, blocks, prims, plain ? sd : 0, 32 * nt * per_term);
    else
    emit_prologue(o, cart ? "tchem_jit_chain_cart" : "tchem_jit_chain", "double * __restrict__ q, double * __restrict__ y, double * __restrict__ J",
                  cartdim, std::max(tper * per_term, sd), warps, blocks, prims, plain ? sd : 0);
    for (size_t p = 0; p < prims.size(); p++)
        if (!emit_prim(o, prims[p], p, cart, "        ")) return std::string();
    // cart: the structurally nonzero elements of J as named values J<row>_<col>
    std::vector<std::vector<char>> jnz(n_ic, std::vector<char>(cartdim, 0));
    if (cart) {
        static const char * comp[3] = {".x", ".y", ".z"};
        for (int i = 0; i < n_ic; i++) {
            std::vector<std::string> expr(cartdim);
            for (auto & t : rows[i]) {
                const JitPrim & pr = prims[t.second];
                for (int a = 0; a < pr.natoms; a++)
                    for (int c = 0; c < 3; c++) {
                        std::string & e = expr[3 * pr.atoms[a] + c];
                        const std::string v = "j" + std::to_string(t.second) + "[" + std::to_string(a) + "]" + comp[c];
                        e = e.empty() ? hexd(t.first) + " * " + v : "fma(" + hexd(t.first) + ", " + v + ", " + e + ")";
                    }
            }
            for (int c = 0; c < cartdim; c++)
                if (!expr[c].empty()) {
                    jnz[i][c] = 1;
                    o << "        const double J" << i << "_" << c << " = " << expr[c] << ";\n";
                }
        }
    }
    std::vector<int> maxpow(sd, 0);
    for (auto & t : saps) for (auto & u : t) maxpow[u.first] = std::max(maxpow[u.first], u.second);
    for (int i = 0; i < n_ic; i++) {
        if (!plain) o << "        const double x" << i << "p1 = " << row_expr(rows[i], "q") << ";\n";
        for (int p = 2; p <= maxpow[i]; p++) o << "        const double x" << i << "p" << p << " = x" << i << "p" << p - 1 << " * x" << i << "p1;\n";
    }
    o << "        if (q) {\n";
    for (int i = 0; i < n_ic; i++) o << "            D[" << i << " * KPAD + lane] = x" << i << "p1;\n";
    o << "            __syncwarp();\n            copy_rows<SD, SD>(D, q + b0 * SD, ntile, lane);\n            __syncwarp();\n        }\n";
    auto power = [](int col, int p) { return "x" + std::to_string(col) + "p" + std::to_string(p); };
    if (bulk) {
        // boxes: [32][NT] values, then [32][NT * SD] Jacobian rows; a lane owns row `lane` of both
        std::vector<std::string> yv(nt), jv((size_t)nt * sdo, "0.0");
        std::ostringstream gdefs;                          // cart: the nonzero d SAP / d q_k as named values
        for (int t = 0; t < nt; t++) {
            std::string v;
            for (auto & u : saps[t]) v = v.empty() ? power(u.first, u.second) : v + " * " + power(u.first, u.second);
            yv[t] = v.empty() ? "1.0" : v;
            std::vector<std::string> g(sd, "0.0");
            for (size_t ui = 0; ui < saps[t].size(); ui++) {
                const auto & u = saps[t][ui];
                std::string e = hexd((double)u.second);
                if (u.second > 1) e += " * " + power(u.first, u.second - 1);
                for (size_t wi = 0; wi < saps[t].size(); wi++)
                    if (wi != ui) e += " * " + power(saps[t][wi].first, saps[t][wi].second);
                g[u.first] = e;
            }
            if (!cart) {
                for (int k = 0; k < sd; k++) jv[(size_t)t * sd + k] = g[k];
                continue;
            }
            // d/dr_c = sum_k (d/dq_k) J[k][c] over the rows k where both factors are structurally nonzero
            for (int k = 0; k < sd; k++)
                if (g[k] != "0.0") gdefs << "            const double g" << t << "_" << k << " = " << g[k] << ";\n";
            for (int c = 0; c < sdo; c++) {
                std::string e;
                for (int k = 0; k < sd; k++) {
                    if (g[k] == "0.0" || !jnz[k][c]) continue;
                    const std::string gk = "g" + std::to_string(t) + "_" + std::to_string(k), jk = "J" + std::to_string(k) + "_" + std::to_string(c);
                    e = e.empty() ? gk + " * " + jk : "fma(" + gk + ", " + jk + ", " + e + ")";
                }
                if (!e.empty()) jv[(size_t)t * sdo + c] = e;
            }
        }
        o << "        {\n" << gdefs.str()
          << "            const int row = warp * LANES + lane;                   // this geometry's row in the block's boxes\n"
          << "            double * Yb = Box + row * NT, * Jb = Box + nwarps * LANES * NT + row * (NT * SDO);\n"
          << "            // the previous block tile's stores have read the boxes\n"
          << "            if (threadIdx.x == 0) asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\");\n"
          << "            __syncthreads();\n";
        for (int t = 0; t < nt; t += 2)
            o << "            *reinterpret_cast<double2 *>(Yb + " << t << ") = make_double2(" << yv[t] << ", " << yv[t + 1] << ");\n";
        for (int e = 0; e < nt * sdo; e += 2)
            if (jv[e] != "0.0" || jv[e + 1] != "0.0")
                o << "            *reinterpret_cast<double2 *>(Jb + " << e << ") = make_double2(" << jv[e] << ", " << jv[e + 1] << ");\n";
        o << "            asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n"
          << "            __syncthreads();\n"
          << "            if (threadIdx.x == 0) {\n"
          << "                const int64_t g0 = tile * nwarps * LANES;\n"
          << "                const int nblk = (int)min((int64_t)nwarps * LANES, B - g0);\n"
          << "                const uint32_t box_sh = (uint32_t)__cvta_generic_to_shared(Box);\n"
          << "                if (y) asm volatile(\"cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\" ::\"l\"(y + g0 * NT), \"r\"(box_sh), \"r\"(nblk * (NT * 8)) : \"memory\");\n"
          << "                if (J) asm volatile(\"cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\" ::\"l\"(J + g0 * (NT * SDO)), \"r\"(box_sh + nwarps * LANES * NT * 8), \"r\"(nblk * (NT * SDO * 8)) : \"memory\");\n"
          << "                asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\");\n"
          << "            }\n        }\n    }\n"
          << "    if (threadIdx.x == 0) asm volatile(\"cp.async.bulk.wait_group 0;\" ::: \"memory\");\n}\n";
        return o.str();
    }
    for (int t0 = 0; t0 < nt; t0 += tper) {
        const int t1 = std::min(nt, t0 + tper), n = t1 - t0;
        o << "        {   // monomials " << t0 << " .. " << t1 - 1 << "\n";
        for (int t = t0; t < t1; t++) {
            // value: product of the powers (source/polynomial/SAP.cpp:93-99)
            std::string v;
            for (auto & u : saps[t]) v = v.empty() ? power(u.first, u.second) : v + " * " + power(u.first, u.second);
            o << "            D[" << (t - t0) << " * KPAD + lane] = " << (v.empty() ? "1.0" : v) << ";\n";
            // d/dx_u = p x_u^(p-1) prod_{w != u} x_w^(p_w)   (SAP.cpp:147-175); zero elsewhere
            std::vector<std::string> g(sd, "0.0");
            for (size_t ui = 0; ui < saps[t].size(); ui++) {
                const auto & u = saps[t][ui];
                std::string e = hexd((double)u.second);
                if (u.second > 1) e += " * " + power(u.first, u.second - 1);
                for (size_t wi = 0; wi < saps[t].size(); wi++)
                    if (wi != ui) e += " * " + power(saps[t][wi].first, saps[t][wi].second);
                g[u.first] = e;
            }
            if (cart) {
                // d/dr_c = sum_k (d/dq_k) J[k][c] over the rows k where both factors are structurally nonzero
                for (int k = 0; k < sd; k++)
                    if (g[k] != "0.0") o << "            const double g" << t << "_" << k << " = " << g[k] << ";\n";
                std::vector<std::string> gc(sdo, "0.0");
                for (int c = 0; c < sdo; c++) {
                    std::string e;
                    for (int k = 0; k < sd; k++) {
                        if (g[k] == "0.0" || !jnz[k][c]) continue;
                        const std::string gk = "g" + std::to_string(t) + "_" + std::to_string(k), jk = "J" + std::to_string(k) + "_" + std::to_string(c);
                        e = e.empty() ? gk + " * " + jk : "fma(" + gk + ", " + jk + ", " + e + ")";
                    }
                    if (!e.empty()) gc[c] = e;
                }
                g = gc;
            }
            for (int c = 0; c < sdo; c++)
                o << "            D[" << n + stage_row((t - t0) * sdo + c, n * sdo) << " * KPAD + lane] = " << g[c] << ";\n";
        }
        o << "            __syncwarp();\n"
          << "            if (y) copy_rows<" << n << ", NT>(D, y + b0 * NT + " << t0 << ", ntile, lane);\n"
          << "            if (J) copy_out<" << n * sdo << ", NT * SDO>(D + " << n << " * KPAD, J + b0 * (NT * SDO) + " << t0 * sdo << ", ntile, lane);\n"
          << "            __syncwarp();\n        }\n";
    }
    o << "    }\n}\n";
    return o.str();
}

// Second-order Jacobian of a SAPSet (SAPSet::Jacobian2nd_(xs, K), source/polynomial/SAPSet.cpp:243-276)
// specialised to the set: every element of the symmetric [NT][SD][SD] block is a literal product of
// shared powers (SAP::Hessian_, source/polynomial/SAP.cpp:226-263; structural zeros are literal
// zeros). Each lane stages its WHOLE geometry (G = NT * SD * SD doubles) in the layout the result
// has in global memory, so a tile's 32 geometries are one contiguous block and ONE lane hands it
// to the TMA unit as a single 1-D bulk copy (cp.async.bulk.global.shared::cta): no copy-out
// instructions, and a write front that is linear per warp (the best pattern of
// tools/micro/write_pattern.cu). With G = 2 mod 4 the row pitch is an odd number of 16-byte
// units and the staging stores are bank-conflict free.
std::string jit_hess_source(const std::vector<std::vector<std::pair<int, int>>> & saps, int sd, int warps, int blocks) {
    const int nt = (int)saps.size(), G = nt * sd * sd;
    if (G % 2) return std::string();
    const int rsize = (sd * 33 + 1) & ~1;
    std::ostringstream o;
    o << "#define CD " << sd << "\n#define DCAP 0\n#define G " << G << "\n#define RSIZE " << rsize << "\n";
    o << kPrelude;
    o << "extern \"C\" __global__ void __launch_bounds__(" << warps * 32 << ", " << blocks << ")\n"
      << "tchem_jit_sap_hess(const double * __restrict__ r, const long long B, double * __restrict__ K) {\n"
      << "    extern __shared__ __align__(16) double smem[];\n"
      << "    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;\n"
      << "    double * Sbox = smem + (size_t)warp * (LANES * G + RSIZE);\n"
      << "    double * R = Sbox + LANES * G;\n"
      << "    double * S = Sbox + lane * G;               // this lane's geometry, laid out like the result\n"
      << "    const uint32_t S_sh = (uint32_t)__cvta_generic_to_shared(Sbox);\n"
      << "    const int64_t ntiles = (B + LANES - 1) / LANES, step = (int64_t)gridDim.x * nwarps;\n"
      << "    int64_t tile = (int64_t)blockIdx.x * nwarps + warp;\n"
      << "    if (tile < ntiles) { const int64_t b0 = tile * LANES; stage_issue(r, b0, (int)min((int64_t)LANES, B - b0), R, lane); }\n"
      << "    // structural zeros never change: staged once, the box keeps them from tile to tile\n"
      << "    for (int e = 0; e < G; e += 2) *reinterpret_cast<double2 *>(S + e) = make_double2(0.0, 0.0);\n"
      << "    for (; tile < ntiles; tile += step) {\n"
      << "        const int64_t b0 = tile * LANES;\n"
      << "        const int ntile = (int)min((int64_t)LANES, B - b0);\n"
      << "        stage_finish(ntile, R, lane);\n";
    for (int i = 0; i < sd; i++) o << "        const double x" << i << "p1 = R[" << i << " * KPAD + lane];\n";
    o << "        __syncwarp();\n"
      << "        if (tile + step < ntiles) { const int64_t nb0 = (tile + step) * LANES; stage_issue(r, nb0, (int)min((int64_t)LANES, B - nb0), R, lane); }\n";
    std::vector<int> maxpow(sd, 0);
    for (auto & t : saps) for (auto & u : t) maxpow[u.first] = std::max(maxpow[u.first], u.second);
    for (int i = 0; i < sd; i++)
        for (int p = 2; p <= maxpow[i]; p++) o << "        const double x" << i << "p" << p << " = x" << i << "p" << p - 1 << " * x" << i << "p1;\n";
    auto power = [](int col, int p) { return "x" + std::to_string(col) + "p" + std::to_string(p); };
    std::vector<std::string> expr((size_t)G, "0.0");
    for (int t = 0; t < nt; t++) {
        const auto & us = saps[t];                       // (column, power), ascending column = uniques_orders_
        for (size_t a = 0; a < us.size(); a++) {
            const int ca = us[a].first, pa = us[a].second;
            if (pa >= 2) {
                std::string e = hexd((double)(pa * (pa - 1)));
                if (pa > 2) e += " * " + power(ca, pa - 2);
                for (size_t w = 0; w < us.size(); w++) if (w != a) e += " * " + power(us[w].first, us[w].second);
                expr[(size_t)(t * sd + ca) * sd + ca] = e;
            }
            for (size_t b = a + 1; b < us.size(); b++) {
                const int cb = us[b].first, pb = us[b].second;
                std::string e = hexd((double)(pa * pb));
                if (pa > 1) e += " * " + power(ca, pa - 1);
                if (pb > 1) e += " * " + power(cb, pb - 1);
                for (size_t w = 0; w < us.size(); w++) if (w != a && w != b) e += " * " + power(us[w].first, us[w].second);
                expr[(size_t)(t * sd + ca) * sd + cb] = e;
                expr[(size_t)(t * sd + cb) * sd + ca] = e;                // SAPSet.cpp:271-274
            }
        }
    }
    o << "        // the previous tile's store has read the box\n"
      << "        if (lane == 0) asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\");\n"
      << "        __syncwarp();\n";
    for (int e = 0; e < G; e += 2)
        if (expr[e] != "0.0" || expr[e + 1] != "0.0")
            o << "        *reinterpret_cast<double2 *>(S + " << e << ") = make_double2(" << expr[e] << ", " << expr[e + 1] << ");\n";
    o << "        asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n"
      << "        __syncwarp();\n"
      << "        if (lane == 0) {\n"
      << "            asm volatile(\"cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\" ::\"l\"(K + b0 * G), \"r\"(S_sh), \"r\"(ntile * (G * 8)) : \"memory\");\n"
      << "            asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\");\n"
      << "        }\n"
      << "    }\n"
      << "    if (lane == 0) asm volatile(\"cp.async.bulk.wait_group 0;\" ::: \"memory\");\n"
      << "}\n";
    return o.str();
}

// The CUDA C++ source of the specialised kernel for this table; `dcap` = dense elements staged
// per chunk. Empty string when the table is not eligible.
// `bulk`: the write-out goes through the TMA unit. The chunk is staged geometry-major in shared
// memory as the dense box [32 geometries][piece / 8][8 doubles] a tiled tensor map describes
// (J seen as a 3-D tensor: 64-byte groups x groups per geometry x B), each lane storing its own
// geometry with 16-byte shared stores at the 64B-swizzled address (conflict-free: the row pitch is a
// multiple of 64 bytes, the swizzle spreads the 8 lanes of a quarter-warp over the 8 bank groups),
// and ONE lane issues one cp.async.bulk.tensor store per chunk (q rows: a second, 2-D map). The
// warp copy-out it replaces was 36 % of the kernel's instructions. (Per-lane 1-D bulk copies were
// tried first: UBLKCP takes uniform-register operands, so the compiler serialises the 32 lanes -
// ~330 instructions per copy - and nothing is gained.) Rows of a ragged tail tile lie outside the
// tensor and are clipped by the TMA unit. Shapes: the piece and the geometry must be whole 64-byte
// groups, chunks of equal even row count.
bool jit_bulk_eligible(int n_ic, int cartdim, int dcap) {
    if (cartdim > dcap) return false;
    const int rows_per = std::min(n_ic, std::max(1, dcap / (cartdim + 1)));
    return rows_per % 2 == 0 && n_ic % rows_per == 0 && (rows_per * cartdim) % 8 == 0 && rows_per * cartdim / 8 <= 256;
}
// per-warp shared memory of the bulk variant: [J box][q box][R], every part a multiple of 512 bytes
size_t jit_bulk_warp_bytes(int n_ic, int cartdim, int dcap) {
    const int rows_per = std::min(n_ic, std::max(1, dcap / (cartdim + 1)));
    auto up = [](size_t v) { return (v + 511) & ~(size_t)511; };
    return up((size_t)32 * rows_per * cartdim * 8) + up((size_t)32 * rows_per * 8) + up((size_t)cartdim * 33 * 8);
}

std::string jit_source(const std::vector<tchem_invdisp_t> & terms, int n_ic, int cartdim, int dcap, int warps, int blocks, bool bulk = false) {
    if (bulk && !jit_bulk_eligible(n_ic, cartdim, dcap)) return std::string();
    typedef std::tuple<int, int, int, int, int, int, double> Key;
    std::map<Key, int> index;
    std::vector<JitPrim> prims;
    std::vector<std::vector<std::pair<double, int>>> rows(n_ic);       // (coef, prim) in file order
    for (const auto & d : terms) {
        if (d.natoms == 5) return std::string();
        int a[5];
        for (int i = 0; i < 5; i++) a[i] = i < d.natoms ? d.atoms[i] : -1;
        const double mn = d.type == TCHEM_TORSION ? d.min : 0.0;
        Key k(d.type, a[0], a[1], a[2], a[3], a[4], mn);
        auto f = index.find(k);
        int id;
        if (f == index.end()) {
            id = (int)prims.size();
            index[k] = id;
            JitPrim p;
            p.type = d.type; p.natoms = d.natoms; p.min = d.min;
            for (int i = 0; i < 5; i++) p.atoms[i] = a[i];
            prims.push_back(p);
        } else id = f->second;
        rows[d.ic].push_back({d.coeff, id});
    }
    if (cartdim > dcap) return std::string();
    // chunks of consecutive rows, <= dcap dense elements (+ their q rows)
    std::vector<std::pair<int, int>> chunks;
    const int rows_per = std::max(1, dcap / (cartdim + 1));
    for (int r0 = 0; r0 < n_ic; r0 += rows_per) chunks.push_back({r0, std::min(n_ic, r0 + rows_per)});
    const int natoms = cartdim / 3;

    std::ostringstream o;
    o << "#define CD " << cartdim << "\n#define NIC " << n_ic << "\n#define DCAP " << rows_per * (cartdim + 1) << "\n";
    if (bulk) {
        auto up = [](size_t v) { return (v + 511) & ~(size_t)511; };
        o << "#define PIECEB " << rows_per * cartdim * 8 << "\n#define QB " << rows_per * 8
          << "\n#define JBOX " << up((size_t)32 * rows_per * cartdim * 8) << "\n#define QBOX " << up((size_t)32 * rows_per * 8)
          << "\n#define WARPB " << jit_bulk_warp_bytes(n_ic, cartdim, dcap) << "\n";
    }
    o << kPrelude;
    o << "extern \"C\" __global__ void __launch_bounds__(" << warps * 32 << ", " << blocks << ")\n"
      << (bulk ? "tchem_jit_qJ_bulk" : "tchem_jit_qJ") << "(const double * __restrict__ r, const long long B, double * __restrict__ q, double * __restrict__ J"
      << (bulk ? ", const __grid_constant__ TMap tmJ, const __grid_constant__ TMap tmq" : "") << ") {\n"
      << "    extern __shared__ __align__(16) double smem[];\n"
      << "    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;\n";
    if (bulk)
        o << "    // swizzle patterns follow shared-memory address bits: the boxes start on 1024-byte boundaries\n"
          << "    char * Sbox = reinterpret_cast<char *>(smem) + ((1024 - ((uint32_t)__cvta_generic_to_shared(smem) & 1023)) & 1023) + (size_t)warp * WARPB;\n"
          << "    char * Qbox = Sbox + JBOX;\n"
          << "    double * R = reinterpret_cast<double *>(Qbox + QBOX);\n"
          << "    const uint32_t S_sh = (uint32_t)__cvta_generic_to_shared(Sbox), Q_sh = (uint32_t)__cvta_generic_to_shared(Qbox);\n"
          << "    double * Qrow = reinterpret_cast<double *>(Qbox + lane * QB);\n";
    else
        o << "    double * R = smem + (size_t)warp * (CD + DCAP) * KPAD;\n"
          << "    double * D = R + CD * KPAD;\n";
    o
      << "    const int64_t ntiles = (B + LANES - 1) / LANES, step = (int64_t)gridDim.x * nwarps;\n"
      << "    int64_t tile = (int64_t)blockIdx.x * nwarps + warp;\n"
      << "    if (tile < ntiles) { const int64_t b0 = tile * LANES; stage_issue(r, b0, (int)min((int64_t)LANES, B - b0), R, lane); }\n"
      << "    for (; tile < ntiles; tile += step) {\n"
      << "        const int64_t b0 = tile * LANES;\n"
      << "        const int ntile = (int)min((int64_t)LANES, B - b0);\n"
      << "        stage_finish(ntile, R, lane);\n"
      << "        V3<double> a[" << natoms << "];\n";
    std::vector<char> atom_used(natoms, 0);
    for (const auto & p : prims) for (int i = 0; i < p.natoms; i++) atom_used[p.atoms[i]] = 1;
    for (int at = 0; at < natoms; at++)
        if (atom_used[at])
            o << "        a[" << at << "] = v3<double>(R[" << 3 * at << " * KPAD + lane], R[" << 3 * at + 1 << " * KPAD + lane], R["
              << 3 * at + 2 << " * KPAD + lane]);\n";
    o << "        __syncwarp();\n"
      << "        // the coordinates now live in registers: the next tile's copy can start\n"
      << "        if (tile + step < ntiles) { const int64_t nb0 = (tile + step) * LANES; stage_issue(r, nb0, (int)min((int64_t)LANES, B - nb0), R, lane); }\n";
    for (const auto & ch : chunks) {
        const int r0 = ch.first, r1 = ch.second, nrows = r1 - r0;
        o << "        {   // rows " << r0 << " .. " << r1 - 1 << "\n";
        std::vector<char> need(prims.size(), 0);
        for (int i = r0; i < r1; i++) for (auto & t : rows[i]) need[t.second] = 1;
        for (size_t p = 0; p < prims.size(); p++) {
            if (!need[p]) continue;
            const JitPrim & pr = prims[p];
            o << "            double q" << p << " = 1.0;";
            if (pr.natoms == 0) { o << "\n"; continue; }
            o << " V3<double> j" << p << "[" << pr.natoms << "];\n            { const V3<double> A[" << pr.natoms << "] = {";
            for (int i = 0; i < pr.natoms; i++) o << (i ? ", " : "") << "a[" << pr.atoms[i] << "]";
            o << "}; auto emit = [&](int i, const V3<double> & v) { j" << p << "[i] = v; };";
            switch (pr.type) {
            case TCHEM_STRETCHING: o << " stretching_J<double>(A, q" << p << ", emit);"; break;
            case TCHEM_BENDING:    o << " double c; bending_J<double, false>(A, c, emit); q" << p << " = acos(c);"; break;
            case TCHEM_COSBENDING: o << " bending_J<double, true>(A, q" << p << ", emit);"; break;
            case TCHEM_TORSION:    o << " double sp = 0.0, cp = 0.0, st = 0.0; torsion_J<double, 0>(A, sp, cp, st, emit); q" << p
                                     << " = signed_wrapped_angle(cp, st, " << hexd(pr.min) << ");"; break;
            case TCHEM_SINTORSION: o << " double sp = 0.0, cp = 0.0, st = 0.0; torsion_J<double, 1>(A, sp, cp, st, emit); q" << p << " = sp;"; break;
            case TCHEM_COSTORSION: o << " double sp = 0.0, cp = 0.0, st = 0.0; torsion_J<double, 2>(A, sp, cp, st, emit); q" << p << " = cp;"; break;
            case TCHEM_OUTOFPLANE: o << " double s; oop_J<double, false>(A, s, emit); q" << p << " = asin(s);"; break;
            case TCHEM_SINOOP:     o << " oop_J<double, true>(A, q" << p << ", emit);"; break;
            default: return std::string();
            }
            o << " }\n";
        }
        // dense elements of the chunk: [row - r0][col] then the q rows
        std::vector<std::string> bulk_J, bulk_qv;      // bulk: the chunk's expressions in output order
        for (int i = r0; i < r1; i++) {
            std::vector<std::string> expr(cartdim);
            for (auto & t : rows[i]) {
                const JitPrim & pr = prims[t.second];
                for (int s = 0; s < pr.natoms; s++)
                    for (int c = 0; c < 3; c++) {
                        std::string & e = expr[3 * pr.atoms[s] + c];
                        std::ostringstream term;
                        term << hexd(t.first) << " * j" << t.second << "[" << s << "]." << "xyz"[c];
                        e = e.empty() ? term.str() : "fma(" + hexd(t.first) + ", j" + std::to_string(t.second) + "[" + std::to_string(s)
                                                      + "]." + std::string(1, "xyz"[c]) + ", " + e + ")";
                    }
            }
            std::string qe;
            for (auto & t : rows[i]) {
                const std::string term = hexd(t.first) + " * q" + std::to_string(t.second);
                qe = qe.empty() ? term : "fma(" + hexd(t.first) + ", q" + std::to_string(t.second) + ", " + qe + ")";
            }
            if (bulk) {
                for (int c = 0; c < cartdim; c++) bulk_J.push_back(expr[c].empty() ? "0.0" : expr[c]);
                bulk_qv.push_back(qe);
                continue;
            }
            for (int c = 0; c < cartdim; c++)
                o << "            D[" << stage_row((i - r0) * cartdim + c, nrows * cartdim) << " * KPAD + lane] = " << (expr[c].empty() ? "0.0" : expr[c]) << ";\n";
            o << "            D[" << nrows * cartdim + (i - r0) << " * KPAD + lane] = " << qe << ";\n";
        }
        if (bulk) {
            // the previous chunk's stores have read the boxes
            o << "            if (lane == 0) asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\");\n"
              << "            __syncwarp();\n";
            for (size_t e = 0; e + 1 < bulk_J.size(); e += 2)
                o << "            stage_pair(Sbox, lane * PIECEB + " << e * 8 << ", " << bulk_J[e] << ", " << bulk_J[e + 1] << ");\n";
            for (size_t e = 0; e + 1 < bulk_qv.size(); e += 2)
                o << "            *reinterpret_cast<double2 *>(Qrow + " << e << ") = make_double2(" << bulk_qv[e] << ", " << bulk_qv[e + 1] << ");\n";
            o << "            asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n"
              << "            __syncwarp();\n"
              << "            if (lane == 0) {\n"
              << "                if (J) asm volatile(\"cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];\" ::\"l\"(&tmJ), \"r\"(0), \"r\"("
              << r0 * cartdim / 8 << "), \"r\"((int)b0), \"r\"(S_sh) : \"memory\");\n"
              << "                if (q) asm volatile(\"cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\" ::\"l\"(&tmq), \"r\"("
              << r0 << "), \"r\"((int)b0), \"r\"(Q_sh) : \"memory\");\n"
              << "                asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\");\n"
              << "            }\n        }\n";
            continue;
        }
        o << "            __syncwarp();\n"
          << "            if (J) copy_out<" << nrows * cartdim << ", NIC * CD>(D, J + b0 * (NIC * CD) + " << r0 * cartdim << ", ntile, lane);\n"
          << "            if (q) copy_rows<" << nrows << ", NIC>(D + " << nrows * cartdim << " * KPAD, q + b0 * NIC + " << r0 << ", ntile, lane);\n"
          << "            __syncwarp();\n        }\n";
    }
    o << "    }\n";
    if (bulk) o << "    if (lane == 0) asm volatile(\"cp.async.bulk.wait_group 0;\" ::: \"memory\");\n";
    o << "}\n";
    return o.str();
}

// ---- "group" variant of the bulk q + J kernel (opt-in experiment, TCHEM_JIT_GROUP = G) ---------------------------
// The bulk variant above writes a tile in row chunks: 32 pieces of (rows per chunk) x cartdim doubles, one geometry
// stride apart (C2H4: 576-byte pieces 1 728 bytes apart) - a write front that tops out at 80-85 % of the HBM peak
// even with no arithmetic (tools/micro/write_pattern.cu), while whole-geometry stores of a tile are one contiguous
// block (C2H4: 55 KB). One warp cannot stage whole geometries (55 KB per warp leaves 3 warps per SM), so here the G
// warps of a BLOCK share one tile: the rows are dealt out to the warps by estimated cost - rows that share primitives
// stay together, so no primitive is evaluated twice; lanes still run identical code - every warp stages its rows into
// the ONE whole-geometry box of the block, and after a block barrier one thread hands the box to the TMA unit as a
// single tensor store (plus one for q). Structural zeros are staged once per kernel.
// MEASURED (B200, C2H4, 10^6 geometries): bit-identical results, but 0.46-0.49 ms for G = 2 .. 4 against 0.40 ms of the
// per-warp chunked kernel: three 55 KB boxes per SM mean three tiles in flight instead of eight, each with three block
// barriers, and the per-warp code sections thrash the instruction cache (ncu: barrier and no_instructions are the top
// stalls). The better write front does not pay for the lost overlap, so the variant stays off by default.
// Shapes: cartdim even (16-byte pairs never straddle rows), whole 64-byte groups per geometry, intdim even.
bool jit_group_eligible(int n_ic, int cartdim) {
    const int G8 = n_ic * cartdim;
    return cartdim % 2 == 0 && n_ic % 2 == 0 && G8 % 8 == 0 && G8 / 8 <= 256 && n_ic <= 256;
}
size_t jit_group_block_bytes(int n_ic, int cartdim) {
    auto up = [](size_t v) { return (v + 511) & ~(size_t)511; };
    return 1024 + up((size_t)32 * n_ic * cartdim * 8) + up((size_t)32 * n_ic * 8) + up((size_t)cartdim * 33 * 8);
}
std::string jit_source_group(const std::vector<tchem_invdisp_t> & terms, int n_ic, int cartdim, int G, int blocks) {
    if (!jit_group_eligible(n_ic, cartdim) || G < 1 || G > 8) return std::string();
    std::vector<JitPrim> prims;
    std::vector<std::vector<std::pair<double, int>>> rows;
    if (!collect(terms, n_ic, prims, rows)) return std::string();
    const int natoms = cartdim / 3;
    auto prim_cost = [&](int p) {
        const JitPrim & pr = prims[p];
        return pr.natoms == 4 ? (pr.type == TCHEM_OUTOFPLANE || pr.type == TCHEM_SINOOP ? 11 : 12) : pr.natoms == 3 ? 6 : pr.natoms == 2 ? 2 : 0;
    };
    // rows -> warps: rows that share primitives stay together (connected components: a split would make two warps
    // evaluate the same primitives), the components are dealt out longest processing time first
    std::vector<int> parent(n_ic);
    for (int i = 0; i < n_ic; i++) parent[i] = i;
    std::function<int(int)> find = [&](int x) { return parent[x] == x ? x : parent[x] = find(parent[x]); };
    {
        std::map<int, int> owner;                          // primitive -> first row that uses it
        for (int i = 0; i < n_ic; i++)
            for (auto & t : rows[i]) {
                auto f = owner.find(t.second);
                if (f == owner.end()) owner[t.second] = i;
                else parent[find(i)] = find(f->second);
            }
    }
    std::map<int, std::vector<int>> members;
    for (int i = 0; i < n_ic; i++) members[find(i)].push_back(i);
    std::vector<std::pair<int, std::vector<int>>> comps;   // (cost, rows)
    for (auto & kv : members) {
        std::set<int> ps;
        for (int i : kv.second) for (auto & t : rows[i]) ps.insert(t.second);
        int c = (int)kv.second.size();
        for (int p : ps) c += prim_cost(p);
        comps.push_back({c, kv.second});
    }
    std::stable_sort(comps.begin(), comps.end(), [](const std::pair<int, std::vector<int>> & a, const std::pair<int, std::vector<int>> & b) { return a.first > b.first; });
    std::vector<std::vector<int>> bins(G);
    std::vector<std::set<int>> bin_prims(G);
    std::vector<int> load(G, 0);
    for (auto & c : comps) {
        const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        load[w] += c.first;
        for (int i : c.second) { bins[w].push_back(i); for (auto & t : rows[i]) bin_prims[w].insert(t.second); }
    }
    auto up = [](size_t v) { return (v + 511) & ~(size_t)511; };
    const size_t GB = (size_t)n_ic * cartdim * 8;          // bytes of J per geometry
    std::ostringstream o;
    o << "#define CD " << cartdim << "\n#define NIC " << n_ic << "\n#define DCAP 1\n#define GEOMB " << GB
      << "\n#define JBOX " << up(32 * GB) << "\n#define QBOX " << up((size_t)32 * n_ic * 8) << "\n";
    o << kPrelude;
    o << "// rows per warp (estimated cost):";
    for (int w = 0; w < G; w++) { o << "  w" << w << " {"; for (int i : bins[w]) o << " " << i; o << " } = " << load[w]; }
    o << "\nextern \"C\" __global__ void __launch_bounds__(" << G * 32 << ", " << blocks << ")\n"
      << "tchem_jit_qJ_group(const double * __restrict__ r, const long long B, double * __restrict__ q, double * __restrict__ J"
      << ", const __grid_constant__ TMap tmJ, const __grid_constant__ TMap tmq) {\n"
      << "    extern __shared__ __align__(16) double smem[];\n"
      << "    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;\n"
      << "    // swizzle patterns follow shared-memory address bits: the boxes start on a 1024-byte boundary\n"
      << "    char * Sbox = reinterpret_cast<char *>(smem) + ((1024 - ((uint32_t)__cvta_generic_to_shared(smem) & 1023)) & 1023);\n"
      << "    char * Qbox = Sbox + JBOX;\n"
      << "    double * R = reinterpret_cast<double *>(Qbox + QBOX);\n"
      << "    const uint32_t S_sh = (uint32_t)__cvta_generic_to_shared(Sbox), Q_sh = (uint32_t)__cvta_generic_to_shared(Qbox);\n"
      << "    double * Qrow = reinterpret_cast<double *>(Qbox) + lane * NIC;\n"
      << "    // structural zeros never change: staged once, the box keeps them from tile to tile\n"
      << "    for (int e = threadIdx.x; e < JBOX / 16; e += blockDim.x) reinterpret_cast<double2 *>(Sbox)[e] = make_double2(0.0, 0.0);\n"
      << "    const int64_t ntiles = (B + LANES - 1) / LANES, step = gridDim.x;\n"
      << "    int64_t tile = blockIdx.x;\n"
      << "    if (tile < ntiles) { const int64_t b0 = tile * LANES; stage_issue_block(r, b0, (int)min((int64_t)LANES, B - b0), R, threadIdx.x, blockDim.x); }\n"
      << "    for (; tile < ntiles; tile += step) {\n"
      << "        const int64_t b0 = tile * LANES;\n"
      << "        const int ntile = (int)min((int64_t)LANES, B - b0);\n"
      << "        asm volatile(\"cp.async.wait_group 0;\" ::: \"memory\");\n"
      << "        __syncthreads();                                   // the tile's coordinates have landed (and, first tile, the zeros)\n"
      << "        const int gl = min(lane, ntile - 1);               // ragged tile: idle lanes compute on a copy of the last geometry\n"
      << "        V3<double> a[" << natoms << "];\n";
    std::vector<char> atom_used(natoms, 0);
    for (const auto & p : prims) for (int i = 0; i < p.natoms; i++) atom_used[p.atoms[i]] = 1;
    for (int at = 0; at < natoms; at++)
        if (atom_used[at])
            o << "        a[" << at << "] = v3<double>(R[" << 3 * at << " * KPAD + gl], R[" << 3 * at + 1 << " * KPAD + gl], R["
              << 3 * at + 2 << " * KPAD + gl]);\n";
    o << "        // the previous tile's stores have read the boxes; every warp holds its coordinates in registers\n"
      << "        if (threadIdx.x == 0) asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\");\n"
      << "        __syncthreads();\n"
      << "        if (tile + step < ntiles) { const int64_t nb0 = (tile + step) * LANES; stage_issue_block(r, nb0, (int)min((int64_t)LANES, B - nb0), R, threadIdx.x, blockDim.x); }\n"
      << "        switch (warp) {\n";
    static const char * xyz[3] = {".x", ".y", ".z"};
    for (int w = 0; w < G; w++) {
        o << "        case " << w << ": {\n";
        for (int p : bin_prims[w])
            if (!emit_prim(o, prims[p], (size_t)p, true, "            ")) return std::string();
        std::vector<int> mine = bins[w];
        std::sort(mine.begin(), mine.end());
        for (int i : mine) {
            std::vector<std::string> expr(cartdim);
            for (auto & t : rows[i]) {
                const JitPrim & pr = prims[t.second];
                for (int a = 0; a < pr.natoms; a++)
                    for (int c = 0; c < 3; c++) {
                        std::string & e = expr[3 * pr.atoms[a] + c];
                        const std::string v = "j" + std::to_string(t.second) + "[" + std::to_string(a) + "]" + xyz[c];
                        e = e.empty() ? hexd(t.first) + " * " + v : "fma(" + hexd(t.first) + ", " + v + ", " + e + ")";
                    }
            }
            // (the swizzle follows the FULL box offset, lane part included)
            for (int c = 0; c < cartdim; c += 2)
                if (!expr[c].empty() || !expr[c + 1].empty())
                    o << "            stage_pair(Sbox, lane * GEOMB + " << (size_t)(i * cartdim + c) * 8 << ", " << (expr[c].empty() ? "0.0" : expr[c]) << ", "
                      << (expr[c + 1].empty() ? "0.0" : expr[c + 1]) << ");\n";
            o << "            Qrow[" << i << "] = " << row_expr(rows[i], "q") << ";\n";
        }
        o << "        } break;\n";
    }
    o << "        default: break;\n        }\n"
      << "        asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n"
      << "        __syncthreads();                                   // the whole tile is staged\n"
      << "        if (threadIdx.x == 0) {\n"
      << "            if (J) asm volatile(\"cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];\" ::\"l\"(&tmJ), \"r\"(0), \"r\"(0), \"r\"((int)b0), \"r\"(S_sh) : \"memory\");\n"
      << "            if (q) asm volatile(\"cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\" ::\"l\"(&tmq), \"r\"(0), \"r\"((int)b0), \"r\"(Q_sh) : \"memory\");\n"
      << "            asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\");\n"
      << "        }\n    }\n"
      << "    if (threadIdx.x == 0) asm volatile(\"cp.async.bulk.wait_group 0;\" ::: \"memory\");\n}\n";
    return o.str();
}

// compiles `src` for sm_100a; returns the cubin (empty on failure, reason in `log`)
std::vector<char> jit_compile(const std::string & src, std::string & log, bool verbose) {
    std::vector<char> cubin;
    Nvrtc & n = nvrtc();
    if (!n.ok) { log = "NVRTC (libnvrtc.so.12) not available"; return cubin; }
    nvrtcProgram prog = nullptr;
    const char * hdr_src[] = {kMathHeader};
    const char * hdr_name[] = {"ic_math.cuh"};
    if (n.CreateProgram(&prog, src.c_str(), "tchem_jit.cu", 1, hdr_src, hdr_name) != 0) { log = "nvrtcCreateProgram failed"; return cubin; }
    std::vector<const char *> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
    if (verbose) opts.push_back("--ptxas-options=-v");
    const int rc = n.CompileProgram(prog, (int)opts.size(), opts.data());
    size_t ls = 0;
    n.GetProgramLogSize(prog, &ls);
    if (ls > 1) { log.resize(ls); n.GetProgramLog(prog, &log[0]); }
    if (rc == 0) {
        size_t cs = 0;
        if (n.GetCUBINSize(prog, &cs) == 0 && cs > 0) { cubin.resize(cs); n.GetCUBIN(prog, cubin.data()); }
    }
    n.DestroyProgram(&prog);
    return cubin;
}

// the host-buffer entry point cuts one batch into pieces: the kernel choice follows the whole batch
thread_local int64_t g_batch_hint = 0;

struct JitKernel {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t fn = nullptr;
    int warps = 4, blocks = 3, per_sm = 1;
    size_t smem = 0;
    bool failed = false;
};

namespace {
// (table, variant): 0 = warp copy-out, 1 = bulk (TMA) stores
typedef std::pair<const tchem_ic_table *, int> JitKey;
std::map<JitKey, JitKernel> & jit_kernels() { static std::map<JitKey, JitKernel> m; return m; }
std::mutex & jit_mutex() { static std::mutex m; return m; }
int env_int(const char * name, int dflt) { const char * v = getenv(name); return v ? atoi(v) : dflt; }
}

namespace {
// cuTensorMapEncodeTiled through the runtime (no link against libcuda); nullptr when the driver has none
typedef CUresult (*TensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                         const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeTiled tensor_map_encoder() {
    static const TensorMapEncodeTiled fn = [] {
        void * p = nullptr;
        cudaDriverEntryPointQueryResult res;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &res) != cudaSuccess || res != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<TensorMapEncodeTiled>(p);
    }();
    return fn;
}
}

namespace {
// (SAPSet, IntCoordSet or nullptr, variant): 0 = warp copy-out, 1 = TMA bulk stores
typedef std::tuple<const tchem_sap_table *, const tchem_ic_table *, int> ChainKey;
std::map<ChainKey, JitKernel> & chain_kernels() {
    static std::map<ChainKey, JitKernel> m;
    return m;
}
}

void release_jit(const tchem_ic_table * t) {
    std::lock_guard<std::mutex> lock(jit_mutex());
    for (int variant = 0; variant < 3; variant++) {
        auto it = jit_kernels().find({t, variant});
        if (it != jit_kernels().end()) { if (it->second.lib) cudaLibraryUnload(it->second.lib); jit_kernels().erase(it); }
    }
    for (auto c = chain_kernels().begin(); c != chain_kernels().end();) {
        if (std::get<1>(c->first) == t) { if (c->second.lib) cudaLibraryUnload(c->second.lib); c = chain_kernels().erase(c); }
        else ++c;
    }
}
namespace {
std::map<const tchem_sap_table *, JitKernel> & hess_kernels() { static std::map<const tchem_sap_table *, JitKernel> m; return m; }
}
void release_jit_sap(const tchem_sap_table * t) {
    std::lock_guard<std::mutex> lock(jit_mutex());
    {
        auto h = hess_kernels().find(t);
        if (h != hess_kernels().end()) { if (h->second.lib) cudaLibraryUnload(h->second.lib); hess_kernels().erase(h); }
    }
    for (auto c = chain_kernels().begin(); c != chain_kernels().end();) {
        if (std::get<0>(c->first) == t) { if (c->second.lib) cudaLibraryUnload(c->second.lib); c = chain_kernels().erase(c); }
        else ++c;
    }
}

namespace {
// ---- on-disk cubin cache ------------------------------------------------------------------------
// key = FNV-1a of (source, embedded math header, compile options, CUDA runtime version); directory
// TCHEM_JIT_CACHE (empty or "0" disables), else $XDG_CACHE_HOME/tchem_b200, $HOME/.cache/tchem_b200,
// /tmp/tchem_b200_cache. Files are written to a temporary name and renamed, so concurrent ranks never
// read a partial cubin.
std::string jit_cache_dir() {
    const char * e = getenv("TCHEM_JIT_CACHE");
    if (e) return (e[0] == 0 || std::strcmp(e, "0") == 0) ? std::string() : std::string(e);
    if (const char * x = getenv("XDG_CACHE_HOME")) if (x[0]) return std::string(x) + "/tchem_b200";
    if (const char * h = getenv("HOME")) if (h[0]) return std::string(h) + "/.cache/tchem_b200";
    return "/tmp/tchem_b200_cache";
}
void mkdir_p(const std::string & dir) {
    for (size_t i = 1; i <= dir.size(); i++)
        if (i == dir.size() || dir[i] == '/') ::mkdir(dir.substr(0, i).c_str(), 0755);
}
std::string jit_cache_path(const std::string & src) {
    const std::string dir = jit_cache_dir();
    if (dir.empty()) return std::string();
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const char * p, size_t n) { for (size_t i = 0; i < n; i++) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; } };
    mix(src.data(), src.size());
    mix(kMathHeader, std::strlen(kMathHeader));
    int rt = 0;
    cudaRuntimeGetVersion(&rt);
    const std::string tag = "sm_100a c++17 lineinfo rt" + std::to_string(rt);
    mix(tag.data(), tag.size());
    char name[40];
    snprintf(name, sizeof(name), "/%016llx.cubin", (unsigned long long)h);
    return dir + name;
}
std::vector<char> jit_cache_load(const std::string & path) {
    std::vector<char> v;
    if (path.empty()) return v;
    FILE * f = fopen(path.c_str(), "rb");
    if (!f) return v;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n > 0) { v.resize((size_t)n); if (fread(v.data(), 1, (size_t)n, f) != (size_t)n) v.clear(); }
    fclose(f);
    return v;
}
void jit_cache_store(const std::string & path, const std::vector<char> & cubin) {
    if (path.empty() || cubin.empty()) return;
    mkdir_p(path.substr(0, path.rfind('/')));
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    FILE * f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    fclose(f);
    if (ok) ::rename(tmp.c_str(), path.c_str()); else ::remove(tmp.c_str());
}

bool jit_strict() { static const bool v = env_int("TCHEM_JIT_STRICT", 0) != 0; return v; }
struct JitStats { std::atomic<int64_t> compiled{0}, cache_hits{0}, failed{0}, compile_us{0}; };
JitStats & jit_stats() { static JitStats s; return s; }

// NVRTC-compiles `src` (or takes the cubin from the on-disk cache), loads `name` and sizes the launch; marks
// `jk.failed` on any problem. A failure is never silent: one line on stderr per kernel (the interpreter then
// serves the table at roughly half the throughput), and with TCHEM_JIT_STRICT=1 the launch fails instead.
void jit_finish(JitKernel & jk, const std::string & src, const char * name, size_t smem, int max_smem_optin, bool debug) {
    std::string log, why;
    std::vector<char> cubin;
    bool from_cache = false;
    const std::string cpath = src.empty() ? std::string() : jit_cache_path(src);
    if (!src.empty()) {
        cubin = jit_cache_load(cpath);
        from_cache = !cubin.empty();
        if (!from_cache) {
            const auto t0 = std::chrono::steady_clock::now();
            cubin = jit_compile(src, log, debug);
            jit_stats().compile_us += std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        }
    }
    if (debug) fprintf(stderr, "[tchem] jit %s: %zu bytes of source, %zu bytes of cubin%s\n%s\n", name, src.size(), cubin.size(),
                       from_cache ? " (disk cache)" : "", log.c_str());
    jk.smem = smem;
    if (cubin.empty()) { jk.failed = true; why = src.empty() ? "table not expressible" : log.empty() ? "NVRTC produced no cubin" : log.substr(0, 300); }
    else if (smem > (size_t)max_smem_optin) { jk.failed = true; why = "shared-memory plan exceeds the device limit"; }
    else if (cudaLibraryLoadData(&jk.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess
             || cudaLibraryGetKernel(&jk.fn, jk.lib, name) != cudaSuccess
             || cudaFuncSetAttribute((const void *)jk.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess
             || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&jk.per_sm, (const void *)jk.fn, jk.warps * 32, smem) != cudaSuccess
             || jk.per_sm < 1) {
        why = std::string("loading the cubin failed: ") + cudaGetErrorString(cudaGetLastError());
        jk.failed = true;
        if (from_cache) ::remove(cpath.c_str());          // a stale / corrupt cache entry must not fail forever
    }
    if (jk.failed) {
        jit_stats().failed++;
        if (!src.empty())
            fprintf(stderr, "[tchem] WARNING: the specialised kernel %s could not be built (%s); the interpreter kernel serves this table "
                            "(about half the throughput). TCHEM_JIT_STRICT=1 turns this into an error.\n", name, why.c_str());
    } else if (from_cache) jit_stats().cache_hits++;
    else { jit_stats().compiled++; jit_cache_store(cpath, cubin); }
    if (debug) fprintf(stderr, "[tchem] jit %s: %s, %d warps/block, %d blocks/SM, %zu B smem\n", name,
                       jk.failed ? "FAILED (interpreter serves this table)" : "ok", jk.warps, jk.per_sm, jk.smem);
}
}

// chained r -> q -> SAPSet through a kernel specialised to (IntCoordSet, SAPSet); -1: not eligible
int launch_jit_chain(const tchem_sap_table * st, const tchem_ic_table * t, const double * r, int64_t B,
                     double * q, double * y, double * J, cudaStream_t stream, bool cart) {
    static const int min_batch = env_int("TCHEM_JIT_MIN_BATCH", 16384), enabled = env_int("TCHEM_JIT", 1);
    static const bool debug = getenv("TCHEM_DEBUG") != nullptr;
    // t == nullptr: the SAPSet alone on given coordinates (r = x[B][sumdim])
    if (!enabled || std::max(B, g_batch_hint) < min_batch || (int64_t)st->nterms * (1 + st->sumdim) > 4096) return -1;
    if (t ? (t->has5 || t->cartdim > 36 || t->terms.size() > 128) : st->sumdim > 36) return -1;
    if (cart && (!t || (int64_t)st->nterms * (1 + t->cartdim) > 4096)) return -1;
    const int in_dim = t ? t->cartdim : st->sumdim;
    // one chunk when the whole set fits ~128 staged entries: a chunk that ends inside a geometry's
    // y / J block leaves 32-byte sectors half written, and partial-sector writes cost DRAM
    // read-modify-write traffic (measured on C5: 2 chunks 3.55 ms, 1 chunk 1.79 ms)
    const int per_term = 1 + (cart ? t->cartdim : st->sumdim), nt = st->nterms;
    const int dcap = env_int("TCHEM_JIT_DCAP", nt * per_term <= 128 ? nt * per_term : std::max(per_term, cart ? 120 : 76));
    static const int want_bulk = env_int("TCHEM_JIT_BULK", 1);
    // single-chunk sets: whole-tile TMA bulk stores (16-byte aligned results, even row lengths)
    int variant = want_bulk && chain_bulk_eligible(nt, st->sumdim, dcap)
                  && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(J)) & 15) == 0 ? 1 : 0;
    // Cartesian Jacobian: variant 3 (whole-tile TMA bulk stores: [32][NT] values + [32][NT * cartdim] rows per warp must leave
    // room for at least two boxes per SM) or variant 2 (warp copy-out)
    const size_t cart_box = chain_warp_doubles(in_dim, st->sumdim, 32 * nt * per_term) * sizeof(double);
    if (cart) variant = want_bulk && nt % 2 == 0 && (nt * t->cartdim) % 2 == 0 && 2 * cart_box + 2048 <= (size_t)st->max_smem_optin
                        && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(J)) & 15) == 0 ? 3 : 2;
    static const char * const names[4] = {"tchem_jit_chain", "tchem_jit_chain_bulk", "tchem_jit_chain_cart", "tchem_jit_chain_cart_bulk"};
    JitKernel * k = nullptr;
    for (; variant >= 0; variant--) {
        std::lock_guard<std::mutex> lock(jit_mutex());
        JitKernel & jk = chain_kernels()[ChainKey(st, t, variant)];
        if (!jk.fn && !jk.failed) {
            // bulk variant: the 4 warps of a block share the boxes (128 consecutive geometries per store); measured on C5
            // 2 warps x 4 blocks 86.5 %, 4 x 2 89.0 %, 8 x 1 83 % (too few independent blocks per SM)
            jk.warps = env_int("TCHEM_JIT_WARPS", 4);
            jk.blocks = env_int("TCHEM_JIT_BLOCKS", 2);
            if (variant == 3) {
                // one warp per block, as many independent blocks as fit: a block's store overlaps the other blocks' arithmetic
                jk.warps = env_int("TCHEM_JIT_CART_WARPS", 1);
                jk.blocks = std::max(1, std::min(env_int("TCHEM_JIT_CART_BLOCKS", 8), (int)((size_t)st->max_smem_optin / (jk.warps * cart_box + 1024))));
            }
            const int tper = chain_terms_per_chunk(nt, per_term, dcap);
            static const std::vector<tchem_invdisp_t> no_terms;
            const std::string src = jit_chain_source(t ? t->terms : no_terms, st->sumdim, in_dim, st->h_terms, dcap, jk.warps, jk.blocks, variant == 1 || variant == 3, variant >= 2);
            const size_t smem = (variant == 1 || variant == 3) ? (size_t)jk.warps * chain_warp_doubles(in_dim, st->sumdim, 32 * nt * per_term) * sizeof(double)
                                                               : (size_t)jk.warps * (in_dim + std::max(tper * per_term, st->sumdim)) * 33 * sizeof(double);
            jit_finish(jk, src, names[variant], smem, st->max_smem_optin, debug);
        }
        if (!jk.failed) { k = &jk; break; }
        if (cart && variant == 2) break;           // no other specialised flavour computes the Cartesian product
    }
    if (!k) return jit_strict() ? set_error(TCHEM_ECUDA, "tchem: the specialised chained kernel could not be built (TCHEM_JIT_STRICT=1)") : -1;
    const int64_t ntiles = (B + 31) / 32;
    const int64_t grid = std::min<int64_t>((ntiles + k->warps - 1) / k->warps, (int64_t)st->sm_count * k->per_sm);
    long long Bll = B;
    void * args[] = {(void *)&r, (void *)&Bll, (void *)&q, (void *)&y, (void *)&J};
    TC_CUDA(cudaLaunchKernel((const void *)k->fn, dim3((unsigned)grid), dim3(k->warps * 32), args, k->smem, stream));
    count_launch(names[variant]);
    return TCHEM_OK;
}

// SAPSet::Jacobian2nd_ through the kernel specialised to the set; -1: not eligible (interpreter)
int launch_jit_sap_hess(const tchem_sap_table * st, const double * x, int64_t B, double * K, cudaStream_t stream) {
    static const int min_batch = env_int("TCHEM_JIT_MIN_BATCH", 16384), enabled = env_int("TCHEM_JIT", 1);
    static const bool debug = getenv("TCHEM_DEBUG") != nullptr;
    const int64_t G = (int64_t)st->nterms * st->sumdim * st->sumdim;
    // one geometry = one staged row: even (16-byte stores), at least 3 warps' boxes per SM, 16-byte aligned result
    const size_t per_warp = (size_t)(32 * G + ((st->sumdim * 33 + 1) & ~1)) * sizeof(double);
    if (!enabled || std::max(B, g_batch_hint) < min_batch || G % 2 || st->sumdim > 36 || G > 4096
        || 3 * per_warp > (size_t)st->max_smem_optin || (reinterpret_cast<uintptr_t>(K) & 15)) return -1;
    JitKernel * k;
    {
        std::lock_guard<std::mutex> lock(jit_mutex());
        JitKernel & jk = hess_kernels()[st];
        k = &jk;
        if (!jk.fn && !jk.failed) {
            jk.warps = env_int("TCHEM_JIT_HESS_WARPS", 1);
            const int fit = (int)((size_t)st->max_smem_optin / (per_warp * jk.warps + 1024));
            jk.blocks = std::max(1, std::min(env_int("TCHEM_JIT_HESS_BLOCKS", 16), fit));
            const std::string src = jit_hess_source(st->h_terms, st->sumdim, jk.warps, jk.blocks);
            jit_finish(jk, src, "tchem_jit_sap_hess", per_warp * jk.warps, st->max_smem_optin, debug);
        }
        if (jk.failed) return jit_strict() ? set_error(TCHEM_ECUDA, "tchem: the specialised Jacobian2nd kernel could not be built (TCHEM_JIT_STRICT=1)") : -1;
    }
    const int64_t ntiles = (B + 31) / 32;
    const int64_t grid = std::min<int64_t>((ntiles + k->warps - 1) / k->warps, (int64_t)st->sm_count * k->per_sm);
    long long Bll = B;
    void * args[] = {(void *)&x, (void *)&Bll, (void *)&K};
    TC_CUDA(cudaLaunchKernel((const void *)k->fn, dim3((unsigned)grid), dim3(k->warps * 32), args, k->smem, stream));
    count_launch("tchem_jit_sap_hess");
    return TCHEM_OK;
}

// q + J through the specialised kernel. Returns TCHEM_OK when launched, -1 when this table / batch
// is served by the interpreter instead (never an error by itself).
int launch_jit_qJ(const tchem_ic_table * t, const double * r, int64_t B, double * q, double * J, cudaStream_t stream) {
    static const int min_batch = env_int("TCHEM_JIT_MIN_BATCH", 16384), enabled = env_int("TCHEM_JIT", 1);
    static const bool debug = getenv("TCHEM_DEBUG") != nullptr;
    // (measured: wins for small molecules - C2H4: 0.42 vs 0.64 ms - and loses beyond ~12 atoms, where dense
    // row staging is mostly zeros and the interpreter's compact staging is the better scheme)
    if (!enabled || std::max(B, g_batch_hint) < min_batch || t->has5 || t->cartdim > 36 || t->terms.size() > 128) return -1;
    static const int want_bulk = env_int("TCHEM_JIT_BULK", 1);
    const int dcap = env_int("TCHEM_JIT_DCAP", std::max(t->cartdim + 1, 76));
    // the TMA write-out needs 16-byte aligned results, whole 64-byte groups and 32-bit tile coordinates
    const bool tma_ok = want_bulk && B < ((int64_t)1 << 31) && tensor_map_encoder()
                        && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(J)) & 15) == 0;
    // variant 2 (opt-in): G warps of a block share one tile and one whole-geometry TMA box (TCHEM_JIT_GROUP = G)
    static const int group_g = env_int("TCHEM_JIT_GROUP", 0);
    int variant = tma_ok && group_g > 0 && jit_group_eligible(t->intdim, t->cartdim)
                  && jit_group_block_bytes(t->intdim, t->cartdim) <= (size_t)t->max_smem_optin ? 2
                  : tma_ok && jit_bulk_eligible(t->intdim, t->cartdim, dcap) ? 1 : 0;
    JitKernel * k = nullptr;
    for (; variant >= 0 && !k; variant--) {
        if (variant == 1 && !(tma_ok && jit_bulk_eligible(t->intdim, t->cartdim, dcap))) continue;
        std::lock_guard<std::mutex> lock(jit_mutex());
        JitKernel & jk = jit_kernels()[{t, variant}];
        if (!jk.fn && !jk.failed && variant == 2) {
            jk.warps = std::min(group_g, 8);
            jk.blocks = env_int("TCHEM_JIT_GROUP_BLOCKS", 3);
            const std::string src = jit_source_group(t->terms, t->intdim, t->cartdim, jk.warps, jk.blocks);
            jit_finish(jk, src, "tchem_jit_qJ_group", jit_group_block_bytes(t->intdim, t->cartdim), t->max_smem_optin, debug);
        }
        if (!jk.fn && !jk.failed) {
            // 8 warps per SM is what the ~250-register kernel allows (at 227 registers - 9 warps - it
            // spills and loses 25 %); as 4 blocks of 2 warps they finish 2 % sooner than as 2 blocks of 4
            jk.warps = env_int("TCHEM_JIT_WARPS", 2);
            jk.blocks = env_int("TCHEM_JIT_BLOCKS", 4);
            const std::string src = jit_source(t->terms, t->intdim, t->cartdim, dcap, jk.warps, jk.blocks, variant == 1);
            const int rows_per = std::max(1, dcap / (t->cartdim + 1));
            size_t smem = (size_t)jk.warps * (t->cartdim + rows_per * (t->cartdim + 1)) * 33 * sizeof(double);
            if (variant == 1) smem = 1024 + (size_t)jk.warps * jit_bulk_warp_bytes(t->intdim, t->cartdim, dcap);
            jit_finish(jk, src, variant == 1 ? "tchem_jit_qJ_bulk" : "tchem_jit_qJ", smem, t->max_smem_optin, debug);
        }
        if (!jk.failed) { k = &jk; break; }
    }
    if (!k) return jit_strict() ? set_error(TCHEM_ECUDA, "tchem_ic_eval: the specialised q+J kernel could not be built (TCHEM_JIT_STRICT=1)") : -1;
    const int64_t ntiles = (B + 31) / 32;
    const int64_t grid = variant == 2 ? std::min<int64_t>(ntiles, (int64_t)t->sm_count * k->per_sm)
                                      : std::min<int64_t>((ntiles + k->warps - 1) / k->warps, (int64_t)t->sm_count * k->per_sm);
    long long Bll = B;
    if (variant >= 1) {
        // J as [B][intdim * cartdim / 8][8 doubles] with the 64B swizzle, q as [B][intdim]; boxes = one chunk of one tile
        // (group variant: the whole geometry)
        const int rows_per = variant == 2 ? t->intdim : std::min(t->intdim, std::max(1, dcap / (t->cartdim + 1)));
        const uint64_t G = (uint64_t)t->intdim * t->cartdim;
        CUtensorMap tmJ, tmq;
        std::memset(&tmJ, 0, sizeof(tmJ));
        std::memset(&tmq, 0, sizeof(tmq));
        const cuuint32_t ones[3] = {1, 1, 1};
        if (J) {
            const cuuint64_t dims[3] = {8, G / 8, (cuuint64_t)B}, strides[2] = {64, G * 8};
            const cuuint32_t box[3] = {8, (cuuint32_t)(rows_per * t->cartdim / 8), 32};
            if (tensor_map_encoder()(&tmJ, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, J, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return set_error(TCHEM_ECUDA, "tchem_ic_eval: cuTensorMapEncodeTiled(J) failed");
        }
        if (q) {
            const cuuint64_t dims[2] = {(cuuint64_t)t->intdim, (cuuint64_t)B}, strides[1] = {(cuuint64_t)t->intdim * 8};
            const cuuint32_t box[2] = {(cuuint32_t)rows_per, 32};
            if (tensor_map_encoder()(&tmq, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, q, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return set_error(TCHEM_ECUDA, "tchem_ic_eval: cuTensorMapEncodeTiled(q) failed");
        }
        void * args[] = {(void *)&r, (void *)&Bll, (void *)&q, (void *)&J, (void *)&tmJ, (void *)&tmq};
        TC_CUDA(cudaLaunchKernel((const void *)k->fn, dim3((unsigned)grid), dim3(k->warps * 32), args, k->smem, stream));
    } else {
        void * args[] = {(void *)&r, (void *)&Bll, (void *)&q, (void *)&J};
        TC_CUDA(cudaLaunchKernel((const void *)k->fn, dim3((unsigned)grid), dim3(k->warps * 32), args, k->smem, stream));
    }
    count_launch(variant == 2 ? "tchem_jit_qJ_group" : variant == 1 ? "tchem_jit_qJ_bulk" : "tchem_jit_qJ");
    return TCHEM_OK;
}

} // namespace tchem_b200

using namespace tchem_b200;

extern "C" {

// the generated source of the specialised q + J kernel (testing / inspection); returns the length
// needed (0: table not eligible), copies at most `capacity` bytes including the terminator
int64_t tchem_ic_jit_source(const tchem_invdisp_t * terms, int n_terms, int n_ic, int cartdim, int dcap, char * buf, int64_t capacity) {
    if (!terms || n_terms <= 0) return 0;
    std::vector<tchem_invdisp_t> v(terms, terms + n_terms);
    const std::string s = jit_source(v, n_ic, cartdim, dcap > 0 ? dcap : std::max(cartdim + 1, 76), 4, 3);
    if (buf && capacity > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)capacity - 1);
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int64_t)s.size() + (s.empty() ? 0 : 1);
}

// the group variant (G warps of a block share one tile and one whole-geometry TMA box); 0 when not eligible
int64_t tchem_ic_jit_source_group(const tchem_invdisp_t * terms, int n_terms, int n_ic, int cartdim, int G, char * buf, int64_t capacity) {
    if (!terms || n_terms <= 0) return 0;
    std::vector<tchem_invdisp_t> v(terms, terms + n_terms);
    const std::string s = jit_source_group(v, n_ic, cartdim, G, 3);
    if (buf && capacity > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)capacity - 1);
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int64_t)s.size() + (s.empty() ? 0 : 1);
}

// same for the bulk-store (TMA write-out) variant of the kernel; 0 when the shape is not eligible
int64_t tchem_ic_jit_source_bulk(const tchem_invdisp_t * terms, int n_terms, int n_ic, int cartdim, int dcap, char * buf, int64_t capacity) {
    if (!terms || n_terms <= 0) return 0;
    std::vector<tchem_invdisp_t> v(terms, terms + n_terms);
    const std::string s = jit_source(v, n_ic, cartdim, dcap > 0 ? dcap : std::max(cartdim + 1, 76), 2, 4, true);
    if (buf && capacity > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)capacity - 1);
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int64_t)s.size() + (s.empty() ? 0 : 1);
}

// the generated source of the set-specialised SAPSet::Jacobian2nd_ kernel (testing / inspection); same
// definition arrays as tchem_sap_table_create; returns the length needed, 0 when the set is not eligible
int64_t tchem_sap_jit_hess_source(const int32_t * term_ptr, const int32_t * coords, int n_terms, const int32_t * dims, int n_irreducibles,
                                  char * buf, int64_t capacity) {
    if (!term_ptr || !dims || n_terms <= 0 || n_irreducibles <= 0) return 0;
    std::vector<int> offset(n_irreducibles + 1, 0);
    for (int i = 0; i < n_irreducibles; i++) offset[i + 1] = offset[i] + dims[i];
    std::vector<std::vector<std::pair<int, int>>> saps(n_terms);
    for (int t = 0; t < n_terms; t++) {
        std::map<int, int> powers;
        for (int c = term_ptr[t]; c < term_ptr[t + 1]; c++) powers[offset[coords[2 * c]] + coords[2 * c + 1]]++;
        for (auto & kv : powers) saps[t].push_back({kv.first, kv.second});
    }
    const std::string s = jit_hess_source(saps, offset[n_irreducibles], 1, 4);
    if (buf && capacity > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)capacity - 1);
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int64_t)s.size() + (s.empty() ? 0 : 1);
}

// the generated source of the set-specialised SAPSet value + Jacobian kernel on given coordinates
// (tchem_sap_eval, large batches); bulk != 0: the TMA bulk-store variant. 0 when not eligible.
int64_t tchem_sap_jit_source(const int32_t * term_ptr, const int32_t * coords, int n_terms, const int32_t * dims, int n_irreducibles,
                             int bulk, char * buf, int64_t capacity) {
    if (!term_ptr || !dims || n_terms <= 0 || n_irreducibles <= 0) return 0;
    std::vector<int> offset(n_irreducibles + 1, 0);
    for (int i = 0; i < n_irreducibles; i++) offset[i + 1] = offset[i] + dims[i];
    std::vector<std::vector<std::pair<int, int>>> saps(n_terms);
    for (int t = 0; t < n_terms; t++) {
        std::map<int, int> powers;
        for (int c = term_ptr[t]; c < term_ptr[t + 1]; c++) powers[offset[coords[2 * c]] + coords[2 * c + 1]]++;
        for (auto & kv : powers) saps[t].push_back({kv.first, kv.second});
    }
    const int sd = offset[n_irreducibles], per_term = 1 + sd;
    const int dcap = n_terms * per_term <= 128 ? n_terms * per_term : std::max(per_term, 76);
    static const std::vector<tchem_invdisp_t> no_terms;
    const std::string s = jit_chain_source(no_terms, sd, sd, saps, dcap, bulk ? 2 : 4, bulk ? 4 : 2, bulk != 0);
    if (buf && capacity > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)capacity - 1);
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int64_t)s.size() + (s.empty() ? 0 : 1);
}

int64_t tchem_ic_sap_jit_cart_source(const tchem_invdisp_t * terms, int n_terms, int n_ic, int cartdim, const int32_t * term_ptr,
                                     const int32_t * coords, int n_sap, const int32_t * dims, int n_irreducibles, char * buf, int64_t capacity) {
    if (!terms || !term_ptr || !dims || n_terms <= 0 || n_sap <= 0 || n_irreducibles <= 0) return 0;
    std::vector<int> offset(n_irreducibles + 1, 0);
    for (int i = 0; i < n_irreducibles; i++) offset[i + 1] = offset[i] + dims[i];
    if (offset[n_irreducibles] != n_ic) return 0;
    std::vector<std::vector<std::pair<int, int>>> saps(n_sap);
    for (int t = 0; t < n_sap; t++) {
        std::map<int, int> powers;
        for (int c = term_ptr[t]; c < term_ptr[t + 1]; c++) powers[offset[coords[2 * c]] + coords[2 * c + 1]]++;
        for (auto & kv : powers) saps[t].push_back({kv.first, kv.second});
    }
    const int per_term = 1 + cartdim;
    const int dcap = n_sap * per_term <= 128 ? n_sap * per_term : std::max(per_term, 120);
    std::vector<tchem_invdisp_t> v(terms, terms + n_terms);
    const std::string s = jit_chain_source(v, n_ic, cartdim, saps, dcap, 4, 2, false, true);
    if (buf && capacity > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)capacity - 1);
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int64_t)s.size() + (s.empty() ? 0 : 1);
}

int64_t tchem_ic_sap_jit_cart_bulk_source(const tchem_invdisp_t * terms, int n_terms, int n_ic, int cartdim, const int32_t * term_ptr,
                                          const int32_t * coords, int n_sap, const int32_t * dims, int n_irreducibles, char * buf, int64_t capacity) {
    if (!terms || !term_ptr || !dims || n_terms <= 0 || n_sap <= 0 || n_irreducibles <= 0) return 0;
    std::vector<int> offset(n_irreducibles + 1, 0);
    for (int i = 0; i < n_irreducibles; i++) offset[i + 1] = offset[i] + dims[i];
    if (offset[n_irreducibles] != n_ic) return 0;
    std::vector<std::vector<std::pair<int, int>>> saps(n_sap);
    for (int t = 0; t < n_sap; t++) {
        std::map<int, int> powers;
        for (int c = term_ptr[t]; c < term_ptr[t + 1]; c++) powers[offset[coords[2 * c]] + coords[2 * c + 1]]++;
        for (auto & kv : powers) saps[t].push_back({kv.first, kv.second});
    }
    std::vector<tchem_invdisp_t> v(terms, terms + n_terms);
    const std::string s = jit_chain_source(v, n_ic, cartdim, saps, std::max(n_ic, 1 + cartdim), 1, 3, true, true);
    if (buf && capacity > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)capacity - 1);
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int64_t)s.size() + (s.empty() ? 0 : 1);
}

int tchem_jit_stats(int64_t * out4) {
    if (!out4) return set_error(TCHEM_EINVAL, "tchem_jit_stats: null argument");
    out4[0] = jit_stats().compiled.load(); out4[1] = jit_stats().cache_hits.load();
    out4[2] = jit_stats().failed.load();   out4[3] = jit_stats().compile_us.load();
    return TCHEM_OK;
}

// 1 when tchem_ic_eval(J != NULL, K == NULL) on a batch of B geometries runs the specialised kernel
int tchem_ic_uses_specialised(const tchem_ic_table * t, int64_t B) {
    static const int min_batch = env_int("TCHEM_JIT_MIN_BATCH", 16384), enabled = env_int("TCHEM_JIT", 1);
    if (!t || !enabled || B < min_batch || t->has5 || t->cartdim > 36 || t->terms.size() > 128) return 0;
    std::lock_guard<std::mutex> lock(jit_mutex());
    auto it = jit_kernels().find({t, 0});
    return it == jit_kernels().end() || !it->second.failed ? 1 : 0;
}

// 1 when tchem_ic_sap_eval on B geometries runs the pair-specialised chained kernel
int tchem_ic_sap_uses_specialised(const tchem_sap_table * st, const tchem_ic_table * t, int64_t B) {
    static const int min_batch = env_int("TCHEM_JIT_MIN_BATCH", 16384), enabled = env_int("TCHEM_JIT", 1);
    if (!st || !enabled || B < min_batch || (int64_t)st->nterms * (1 + st->sumdim) > 4096) return 0;
    if (t ? (t->has5 || t->cartdim > 36 || t->terms.size() > 128) : st->sumdim > 36) return 0;
    std::lock_guard<std::mutex> lock(jit_mutex());
    auto it = chain_kernels().find(ChainKey(st, t, 0));
    return it == chain_kernels().end() || !it->second.failed ? 1 : 0;
}

// compiles that source for sm_100a (no device needed); 0 on success, the NVRTC / ptxas log in `log`
int tchem_ic_jit_compile_check(const char * source, char * log, int64_t capacity) {
    std::string l;
    const std::vector<char> cubin = jit_compile(source ? source : "", l, true);
    if (log && capacity > 0) {
        const size_t n = std::min<size_t>(l.size(), (size_t)capacity - 1);
        std::memcpy(log, l.data(), n);
        log[n] = 0;
    }
    return cubin.empty() ? 1 : 0;
}

} // extern "C"
