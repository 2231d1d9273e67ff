// Device-resident "program" an IntCoordSet is compiled into (host side: ic_table.cpp,
// device side: ic_eval.cu).
//
// Layout idea. A warp owns a tile of 32 geometries, one geometry per lane. The IC set is cut
// into chunks of consecutive rows (internal coordinates). For a chunk, every DISTINCT
// invariant displacement ("primitive") used by its rows is evaluated once per geometry in
// registers and its sparse results (value, <=15 gradient entries, and in K mode the <=135
// block-upper second-derivative entries) are parked in a per-warp compact shared-memory
// buffer P[entry][lane]. The dense, API-mandated outputs q / J / K are then produced by a
// gather: every output element owns a short list of (compact entry, coefficient) sources -
// empty for the structural zeros - so that linear combinations, repeated atoms and the
// zero fill all happen on the way out, in fully coalesced stores.
#pragma once
#include <stdint.h>

namespace tchem_b200 {

constexpr int kLanes = 32;
constexpr int kPad = 33;          // P[entry * kPad + lane]: conflict-free for both phases
constexpr int kMaxAtoms = 5;

// MODE_VALUE/QJ/QJK are table modes (each has its own compiled program); MODE_QPATH is an
// evaluation flavour only: values by the compute_IC_J formulas, no derivatives.
// MODE_KDIR: only the second-derivative passes of a range of Cartesian directions.
enum Mode { MODE_VALUE = 0, MODE_QJ = 1, MODE_QJK = 2, MODE_COUNT = 3, MODE_QPATH = 4, MODE_KDIR = 5 };

struct PrimDesc {                  // one distinct InvDisp of a chunk
    int32_t type;
    int32_t natoms;
    int32_t atoms[kMaxAtoms];
    int32_t out;                   // first compact entry: [q][J 3n][K block-upper 9 n(n+1)/2]
    double  min;
};

struct Src {                       // one source of one output element
    int32_t entry;                 // compact entry index
    int32_t pad_;
    double  coef;
};

// One dense output span (the J rows or the K rows of a chunk): `n` element-map entries starting
// at emap[map] cover every element that has sources (plus the zeros inside their sectors and short
// zero stretches); long stretches of structural zeros are listed as runs instead and are stored
// without any table lookup.
struct SpanDesc {
    int64_t map;                   // first ElemMap entry
    int32_t n;                     // number of ElemMap entries
    int32_t run0, nruns;           // zero runs in runs[]
    int32_t nnz;                   // low 16 bits: the first nnz entries cover every sector that holds a non-zero
                                   // element; bits 16+: largest source count among them; -1: see emit_span
};

struct ChunkDesc {
    int32_t prim_begin, prim_end;  // range in prims[]
    int32_t row0, nrows;           // rows (internal coordinates) this chunk produces
    int32_t accumulate;            // != 0: add to what an earlier chunk stored (split row)
    int32_t qslot;                 // first of nrows compact entries that hold the combined q rows
    int64_t qmap;                  // offset into map[]:  nrows words, in row order
    SpanDesc jspan, kspan;         // nrows*cartdim and nrows*cartdim^2 elements
};

struct ZeroRun { uint32_t k, len; };   // elements [k, k + len) of a span are structural zeros
constexpr int kMinZeroRun = 64;

// map word: (count << 24) | first source index (q rows: index into qsrcs; J / K elements: index of
// the element's first 16-byte quad of packed sources)
constexpr uint32_t kMapCountShift = 24;
constexpr uint32_t kMapOffsetMask = 0x00ffffffu;
// J / K elements: this word marks a whole 32-byte sector of structural zeros (4 elements from k on,
// 32-byte aligned in the output): one table entry and one 32-byte store instead of four of each
constexpr uint32_t kZeroSectorWord = 0x00ffffffu;

// Packed source of a J / K element: (coefficient index << 16) | compact entry. An element's
// sources start on a quad boundary, so one 16-byte load fetches up to four of them.
constexpr uint32_t kSrcEntryMask = 0xffffu;
constexpr int kMaxStagedCoefs = 128;     // coefficient tables up to this size are staged in shared memory

// One dense output element of a span: its source word and its position in the span. The
// entries of a span are ordered so that groups of 4 consecutive elements (one 32-byte sector)
// with similar source counts sit next to each other: a warp step then works on 32 elements
// that need the same number of gather rounds, and all-zero steps take the store-only path.
struct ElemMap {
    uint32_t word;                 // (count << 24) | first source index
    uint32_t k;                    // element offset inside the span
};

struct Program {                   // all pointers are device pointers
    const PrimDesc *  prims;
    const ChunkDesc * chunks;
    const uint32_t *  map;         // q rows: (count << 24) | first index into qsrcs
    const Src *       qsrcs;
    const ElemMap *   emap;        // J / K elements, sources in psrcs
    const uint4 *     psrcs;       // packed sources, 4 per 16-byte quad (see PackedSrc)
    const double *    coefs;       // distinct coefficient values
    const ZeroRun *   runs;
    int32_t ncoefs;
    int32_t nqsrcs, nmap;          // sizes of qsrcs[] and map[]
    int32_t split_rows;            // some row is split into accumulating chunks
    int32_t nprims, nchunks;
    int32_t cap;                   // compact entries per lane
    int32_t intdim, cartdim;
    int32_t line_elems;            // > 0: J spans are written through a per-warp line buffer of this many doubles
    int32_t span_bytes;            // > 0: emap | psrcs | coefs | runs (contiguous from emap) are staged in shared memory
};

// compact entries of one primitive
__host__ __device__ inline int prim_entries(int natoms, int mode) {
    if (mode == MODE_VALUE) return 1;
    int n = 1 + 3 * natoms;
    if (mode == MODE_QJK) n += 9 * (natoms * (natoms + 1) / 2);
    return n;
}
// index of 3x3 block (i <= j) in the block-upper K storage of an n-atom primitive
__host__ __device__ inline int kblock(int i, int j, int n) { return i * n - (i * (i - 1)) / 2 + (j - i); }

} // namespace tchem_b200
