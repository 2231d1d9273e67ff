// Host-side "compiler": std::vector<IntCoord> (flattened as tchem_invdisp_t terms) -> Program.
// Reference structures being flattened: include/tchem/IC/IntCoordSet.hpp:11-12,
// include/tchem/IC/IntCoord.hpp:11-12, include/tchem/IC/InvDisp.hpp:54-63.
#include "ic_table.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>

namespace tchem_b200 {

namespace {

const int kTypeAtoms[TCHEM_NTYPES] = {0, 2, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 4, 4};

struct PrimKey {
    int type, natoms, atoms[kMaxAtoms];
    double min;
    bool operator<(const PrimKey & o) const {
        return std::tie(type, natoms, atoms[0], atoms[1], atoms[2], atoms[3], atoms[4], min)
             < std::tie(o.type, o.natoms, o.atoms[0], o.atoms[1], o.atoms[2], o.atoms[3], o.atoms[4], o.min);
    }
};

PrimKey key_of(const tchem_invdisp_t & t) {
    PrimKey k;
    k.type = t.type; k.natoms = t.natoms;
    for (int i = 0; i < kMaxAtoms; i++) k.atoms[i] = i < t.natoms ? t.atoms[i] : -1;
    // `min` only matters for the two angle-valued dihedrals
    k.min = (t.type == TCHEM_TORSION || t.type == TCHEM_TORSION2) ? t.min : 0.0;
    return k;
}

} // namespace

namespace {
// index of `value` in the table of distinct coefficients (bitwise equality), appended when new
int coef_index(std::vector<double> & coefs, double value) {
    for (size_t i = 0; i < coefs.size(); i++)
        if (std::memcmp(&coefs[i], &value, sizeof(double)) == 0) return (int)i;
    coefs.push_back(value);
    return (int)coefs.size() - 1;
}
}

// shortest stretch of structural zeros written as a run (16-byte stores, no table lookups) instead of
// element by element; TCHEM_MIN_ZERO_RUN overrides for experiments
static int min_zero_run() {
    static const int v = [] { const char * e = getenv("TCHEM_MIN_ZERO_RUN"); return e ? std::max(4, atoi(e)) : kMinZeroRun; }();
    return v;
}

bool emit_span(const std::vector<std::vector<Src>> & lists, int64_t abs_off, int64_t gstride,
               std::vector<ElemMap> & emap, std::vector<uint32_t> & psrcs, std::vector<double> & coefs,
               std::vector<ZeroRun> & runs, SpanDesc & span, bool zero_sectors) {
    const int64_t S = (int64_t)lists.size();
    span.map = (int64_t)emap.size();
    span.run0 = (int32_t)runs.size();
    span.nnz = 0;
    const int64_t shift = (gstride % 4 == 0) ? (abs_off % 4) : 0;     // sector phase of element 0
    struct Unit { int64_t first, last; int maxcnt; };
    std::vector<Unit> units;
    for (int64_t k = 0; k < S;) {
        int64_t end = std::min<int64_t>(S, ((k + shift) / 4 + 1) * 4 - shift);
        int m = 0;
        for (int64_t e = k; e < end; e++) m = std::max<int>(m, (int)lists[e].size());
        units.push_back(Unit{k, end, m});
        k = end;
    }
    // long stretches of all-zero sectors become runs
    std::vector<char> in_run(units.size(), 0);
    for (size_t u = 0; u < units.size();) {
        if (units[u].maxcnt != 0) { u++; continue; }
        size_t v = u;
        while (v < units.size() && units[v].maxcnt == 0) v++;
        const int64_t len = units[v - 1].last - units[u].first;
        if (len >= min_zero_run()) {
            runs.push_back(ZeroRun{(uint32_t)units[u].first, (uint32_t)len});
            for (size_t w = u; w < v; w++) in_run[w] = 1;
        }
        u = v;
    }
    std::vector<Unit> rest;
    for (size_t u = 0; u < units.size(); u++) if (!in_run[u]) rest.push_back(units[u]);
    // rounds of 4 sources: order by the number of rounds, then by 1 / 2 / 3-4 sources inside the
    // last round, keeping the address order otherwise
    auto klass = [](int m) { return m == 0 ? 0 : m == 1 ? 1 : m == 2 ? 2 : m <= 4 ? 3 : 4 + (m - 1) / 4; };
    std::stable_sort(rest.begin(), rest.end(), [&](const Unit & a, const Unit & b) { return klass(a.maxcnt) > klass(b.maxcnt); });
    // Zero sectors that do not belong to a run: one entry (and one 32-byte store) per sector instead of
    // four - but only in warp steps of their own (a step that mixes them with gathered elements pays
    // for both loops), i.e. from the next multiple of 32 entries on and when at least one full step
    // of them remains. The sectors before that boundary stay element-wise.
    const bool sectors_aligned = zero_sectors && gstride % 4 == 0;
    size_t sector_from = rest.size();
    if (sectors_aligned) {
        size_t entries = 0, first_zero = rest.size();
        for (size_t i = 0; i < rest.size(); i++) {
            if (rest[i].maxcnt == 0) { first_zero = i; break; }
            entries += (size_t)(rest[i].last - rest[i].first);
        }
        size_t i = first_zero;
        while (i < rest.size() && entries % kLanes != 0) { entries += (size_t)(rest[i].last - rest[i].first); i++; }
        if (i < rest.size() && rest.size() - i >= (size_t)kLanes) sector_from = i;
    }
    for (size_t ui = 0; ui < rest.size(); ui++) {
        const Unit & u = rest[ui];
        if (u.maxcnt > 0) span.nnz += (int32_t)(u.last - u.first);
        if (ui >= sector_from && u.maxcnt == 0 && u.last - u.first == 4 && (abs_off + u.first) % 4 == 0) {
            ElemMap m;
            m.word = kZeroSectorWord;
            m.k = (uint32_t)u.first;
            emap.push_back(m);
            continue;
        }
        for (int64_t e = u.first; e < u.last; e++) {
            const auto & l = lists[e];
            if (l.size() > 255 || psrcs.size() / 4 + l.size() >= kMapOffsetMask) return false;
            ElemMap m;
            m.word = l.empty() ? 0u : (((uint32_t)l.size() << kMapCountShift) | (uint32_t)(psrcs.size() / 4));
            m.k = (uint32_t)e;
            emap.push_back(m);
            for (const Src & s : l) {
                const int ci = coef_index(coefs, s.coef);
                if (ci > 0xffff || s.entry > (int)kSrcEntryMask) return false;
                psrcs.push_back(((uint32_t)ci << 16) | (uint32_t)s.entry);
            }
            while (psrcs.size() % 4) psrcs.push_back(0u);       // the next element starts a new quad
        }
    }
    span.n = (int32_t)(emap.size() - span.map);
    span.nruns = (int32_t)runs.size() - span.run0;
    // line-buffer write-out: <= 64 such elements with <= 4 sources each; the largest source count
    // rides in bits 16+ (picks the unrolled gather), -1 = not eligible
    int maxcnt = 0;
    for (const Unit & u : rest) maxcnt = std::max(maxcnt, u.maxcnt);
    span.nnz = (maxcnt > 4 || span.nnz > 0xffff) ? -1 : (span.nnz | (maxcnt << 16));
    return true;
}

int build_program(const std::vector<tchem_invdisp_t> & terms, int n_ic, int cartdim, int mode,
                  HostProgram & hp, int cap_hint, int max_rows) {
    hp = HostProgram();
    // terms grouped by row, file order kept (source/IC/IntCoord.cpp:62-76 accumulates in order)
    std::vector<std::vector<int>> rows(n_ic);
    for (size_t t = 0; t < terms.size(); t++) {
        const auto & d = terms[t];
        if (d.ic < 0 || d.ic >= n_ic) return set_error(TCHEM_EINVAL, "tchem_ic_table_create: term with ic index out of range");
        if (d.type < 0 || d.type >= TCHEM_NTYPES) return set_error(TCHEM_EDOMAIN, "tchem_ic_table_create: unimplemented internal coordinate type");
        if (d.natoms != kTypeAtoms[d.type]) return set_error(TCHEM_EINVAL, "tchem_ic_table_create: wrong number of atoms for type");
        for (int i = 0; i < d.natoms; i++)
            if (d.atoms[i] < 0 || 3 * d.atoms[i] + 2 >= cartdim) return set_error(TCHEM_EINVAL, "tchem_ic_table_create: atom index out of range");
        rows[d.ic].push_back((int)t);
    }
    for (int i = 0; i < n_ic; i++)
        if (rows[i].empty()) return set_error(TCHEM_EINVAL, "tchem_ic_table_create: internal coordinate without terms");

    // capacity policy: at least the largest primitive; otherwise the largest whole row up to a
    // target that keeps the per-warp shared-memory footprint moderate
    int max_prim = 1, max_row = 1;
    for (int i = 0; i < n_ic; i++) {
        std::map<PrimKey, int> seen;
        int row_entries = 1;           // the row's combined q value
        for (int t : rows[i]) {
            int e = prim_entries(terms[t].natoms, mode);
            max_prim = std::max(max_prim, e + 1);
            if (seen.emplace(key_of(terms[t]), 0).second) row_entries += e;
        }
        max_row = std::max(max_row, row_entries);
    }
    const int target = mode == MODE_QJK ? 160 : 64;
    hp.cap = std::max(max_prim, std::min(max_row, target));
    if (cap_hint > 0) hp.cap = std::max(max_prim, cap_hint);
    if (const char * e = getenv("TCHEM_CAP_MIN")) hp.cap = std::max(hp.cap, atoi(e));   // experiments

    // greedy chunking over rows; a row that does not fit by itself is split into several
    // single-row chunks, the later ones accumulating
    struct Pending { int row0, nrows; bool accumulate; std::vector<int> term_ids; };
    std::vector<Pending> pend;
    {
        Pending cur{0, 0, false, {}};
        auto flush = [&]() {
            if (cur.nrows > 0) pend.push_back(cur);
            cur = Pending{0, 0, false, {}};
        };
        // entries of rows [s, e) staged together (shared primitives counted once)
        auto entries_of = [&](int s, int e) {
            std::map<PrimKey, int> seen;
            int n = e - s;                               // the rows' q slots
            for (int i = s; i < e; i++)
                for (int t : rows[i])
                    if (seen.emplace(key_of(terms[t]), 0).second) n += prim_entries(terms[t].natoms, mode);
            return n;
        };
        // a chunk that ends inside a 32-byte sector of q (4 rows) or of J leaves that sector half
        // written; the partial write costs a DRAM read-modify-write.  Prefer ends on 4-row boundaries.
        // (programs with a fixed row count per chunk - the warp-group kernel's - keep that count: an
        // even spread of the chunks over the group's warps matters more than the q sectors)
        auto aligned = [&](int e) { return max_rows > 0 || e == n_ic || e % 4 == 0; };
        for (int i = 0; i < n_ic; ) {
            int e_max = i;
            while (e_max < n_ic && (max_rows <= 0 || e_max - i < max_rows) && entries_of(i, e_max + 1) <= hp.cap) e_max++;
            if (e_max > i) {
                int e = e_max;
                while (e > i && !aligned(e)) e--;
                // ... unless the rows pushed out share primitives with the ones kept (recomputing
                // them costs more than the sector)
                if (e == i || entries_of(i, e) + entries_of(e, e_max) != entries_of(i, e_max)) e = e_max;
                cur.row0 = i; cur.nrows = e - i;
                for (int r = i; r < e; r++) for (int t : rows[r]) cur.term_ids.push_back(t);
                flush();
                i = e;
                continue;
            }
            // split row
            bool first = true;
            Pending part{i, 1, false, {}};
            std::map<PrimKey, int> part_prims;
            int part_entries = 1;
            for (int t : rows[i]) {
                PrimKey k = key_of(terms[t]);
                int e = part_prims.count(k) ? 0 : prim_entries(terms[t].natoms, mode);
                if (part_entries + e > hp.cap) {
                    part.accumulate = !first;
                    pend.push_back(part);
                    first = false;
                    part = Pending{i, 1, true, {}};
                    part_prims.clear();
                    part_entries = 1;
                    e = prim_entries(terms[t].natoms, mode);
                }
                part_prims.emplace(k, 0);
                part_entries += e;
                part.term_ids.push_back(t);
            }
            part.accumulate = !first;
            pend.push_back(part);
            i++;
        }
        flush();
    }

    // emit chunks
    const int64_t cd = cartdim;
    for (const Pending & pc : pend) {
        ChunkDesc ch;
        std::memset(&ch, 0, sizeof(ch));
        ch.row0 = pc.row0; ch.nrows = pc.nrows; ch.accumulate = pc.accumulate ? 1 : 0;
        ch.prim_begin = (int)hp.prims.size();
        std::map<PrimKey, int> index;   // key -> prim index (global)
        int next_entry = 0;
        for (int t : pc.term_ids) {
            PrimKey k = key_of(terms[t]);
            if (index.count(k)) continue;
            PrimDesc pd;
            std::memset(&pd, 0, sizeof(pd));
            pd.type = terms[t].type; pd.natoms = terms[t].natoms;
            for (int a = 0; a < kMaxAtoms; a++) pd.atoms[a] = a < pd.natoms ? terms[t].atoms[a] : 0;
            pd.out = next_entry;
            pd.min = terms[t].min;
            next_entry += prim_entries(pd.natoms, mode);
            index[k] = (int)hp.prims.size();
            hp.prims.push_back(pd);
        }
        ch.prim_end = (int)hp.prims.size();
        ch.qslot = next_entry;

        // per-row source lists
        auto emit = [&](std::vector<std::vector<Src>> & lists, int64_t & map_off) {
            map_off = (int64_t)hp.map.size();
            for (auto & l : lists) {
                if (l.size() > 255 || hp.qsrcs.size() + l.size() > kMapOffsetMask) return false;
                uint32_t w = l.empty() ? 0u : (((uint32_t)l.size() << kMapCountShift) | (uint32_t)hp.qsrcs.size());
                hp.map.push_back(w);
                for (auto & s : l) hp.qsrcs.push_back(s);
            }
            return true;
        };
        auto emit_sorted = [&](std::vector<std::vector<Src>> & lists, int64_t abs_off, int64_t gstride, SpanDesc & span, bool zsec = false) {
            return emit_span(lists, abs_off, gstride, hp.emap, hp.psrcs, hp.coefs, hp.runs, span, zsec);
        };
        std::vector<std::vector<Src>> ql(pc.nrows), jl, kl;
        if (mode != MODE_VALUE) jl.resize((size_t)pc.nrows * cd);
        if (mode == MODE_QJK) kl.resize((size_t)pc.nrows * cd * cd);
        for (int t : pc.term_ids) {
            const auto & d = terms[t];
            const PrimDesc & pd = hp.prims[index[key_of(d)]];
            const int lr = d.ic - pc.row0;
            Src s; s.pad_ = 0; s.coef = d.coeff;
            s.entry = pd.out;
            ql[lr].push_back(s);
            if (mode == MODE_VALUE) continue;
            const int n = d.natoms;
            for (int a = 0; a < n; a++)
                for (int c = 0; c < 3; c++) {
                    s.entry = pd.out + 1 + 3 * a + c;
                    jl[(size_t)lr * cd + 3 * d.atoms[a] + c].push_back(s);
                }
            if (mode != MODE_QJK) continue;
            const int kbase = pd.out + 1 + 3 * n;
            for (int a = 0; a < n; a++)
                for (int b = 0; b < n; b++)
                    for (int c = 0; c < 3; c++)
                        for (int e = 0; e < 3; e++) {
                            // block (a<=b) stored as is, the mirror block reads it transposed
                            // (source/IC/InvDisp.cpp:869-876)
                            if (a <= b) s.entry = kbase + 9 * kblock(a, b, n) + 3 * c + e;
                            else        s.entry = kbase + 9 * kblock(b, a, n) + 3 * e + c;
                            size_t row = 3 * d.atoms[a] + c, col = 3 * d.atoms[b] + e;
                            kl[((size_t)lr * cd + row) * cd + col].push_back(s);
                        }
        }
        bool ok = emit(ql, ch.qmap);
        if (ok && mode != MODE_VALUE) ok = emit_sorted(jl, (int64_t)pc.row0 * cd, (int64_t)n_ic * cd, ch.jspan);
        if (ok && mode == MODE_QJK) ok = emit_sorted(kl, (int64_t)pc.row0 * cd * cd, (int64_t)n_ic * cd * cd, ch.kspan, true);
        if (!ok) return set_error(TCHEM_EINVAL, "tchem_ic_table_create: internal coordinate definition too large for the gather map");
        hp.chunks.push_back(ch);
    }
    return TCHEM_OK;
}

} // namespace tchem_b200
