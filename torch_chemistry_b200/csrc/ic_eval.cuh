// Device side of the q / J / K path: per-warp tile staging, the primitive interpreter and the
// gather-store. Shared by the q/J/K kernel (ic_eval.cu), the gradient / Hessian transforms
// (ic_transform.cu) and the chained SAP kernel (sap.cu).
#pragma once
#include "ic_math.cuh"
#include "ic_program.h"
#include "../../include/tchem_b200.h"

namespace tchem_b200 {

constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------
// stage the tile's Cartesian coordinates: global r[b0 .. b0+ntile)[cartdim] (contiguous,
// coalesced reads) -> R[c * kPad + lane]. Lanes >= ntile get a copy of the last valid
// geometry so that the arithmetic below never sees garbage.
// The copy is asynchronous (cp.async, 8 bytes per element straight into the transposed shared
// layout): stage_issue() can be called for the NEXT tile as soon as the current tile no longer
// reads R, so that the DRAM latency hides behind the current tile's gather phase.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_issue(const double * __restrict__ r, int64_t b0, int ntile,
                                            int cartdim, double * R, int lane) {
    const double * src = r + b0 * cartdim;
    const int total = ntile * cartdim;
    // running (geometry, coordinate) of flat element f = lane + 32 j, no division in the loop
    int g = lane / cartdim, c = lane - g * cartdim;
    const int dg = kLanes / cartdim, dc = kLanes - dg * cartdim;
    for (int f = lane; f < total; f += kLanes) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(R + c * kPad + g);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src + f) : "memory");
        g += dg; c += dc;
        if (c >= cartdim) { c -= cartdim; g++; }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void stage_finish(int ntile, int cartdim, double * R, int lane) {
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
    if (ntile < kLanes && lane >= ntile)
        for (int k = 0; k < cartdim; k++) R[k * kPad + lane] = R[k * kPad + ntile - 1];
    __syncwarp();
}
// the same copy issued by a group of `nthreads` threads (thread t of the group), no commit / wait
__device__ __forceinline__ void stage_issue_group(const double * __restrict__ r, int64_t b0, int ntile, int cartdim,
                                                  double * R, int t, int nthreads) {
    const double * src = r + b0 * cartdim;
    const int total = ntile * cartdim;
    int g = t / cartdim, c = t - g * cartdim;
    const int dg = nthreads / cartdim, dc = nthreads - dg * cartdim;
    for (int f = t; f < total; f += nthreads) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(R + c * kPad + g);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src + f) : "memory");
        g += dg; c += dc;
        if (c >= cartdim) { c -= cartdim; g++; }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
// named barrier over `nthreads` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void stage_coordinates(const double * __restrict__ r, int64_t b0, int ntile,
                                                  int cartdim, double * R, int lane) {
    stage_issue(r, b0, ntile, cartdim, R, lane);
    stage_finish(ntile, cartdim, R, lane);
}

__device__ __forceinline__ V3<double> load_atom(const double * R, int atom, int lane) {
    const double * p = R + 3 * atom * kPad + lane;
    return v3<double>(p[0], p[kPad], p[2 * kPad]);
}
// where a primitive reads its atoms from: the warp tile (one geometry per lane) ...
struct TileLoader {
    const double * R;
    int lane;
    __device__ __forceinline__ V3<double> operator()(int atom) const { return load_atom(R, atom, lane); }
};
// ... or one geometry r[cartdim] shared by the whole block (dense per-geometry kernels)
struct GeomLoader {
    const double * r;
    __device__ __forceinline__ V3<double> operator()(int atom) const {
        return v3<double>(r[3 * atom], r[3 * atom + 1], r[3 * atom + 2]);
    }
};

template <class S> __device__ __forceinline__ S seed(double v, int coord, int d1, int d2);
template <> __device__ __forceinline__ Dual seed<Dual>(double v, int coord, int d1, int) {
    return mk(v, coord == d1 ? 1.0 : 0.0);
}
template <> __device__ __forceinline__ Dual2 seed<Dual2>(double v, int coord, int d1, int d2) {
    return mk2(v, coord == d1 ? 1.0 : 0.0, coord == d2 ? 1.0 : 0.0, 0.0);
}
template <class S, int N>
__device__ __forceinline__ void seed_atoms(const V3<double> * a, V3<S> * s, int d1, int d2) {
#pragma unroll
    for (int i = 0; i < N; i++)
        s[i] = v3<S>(seed<S>(a[i].x, 3 * i, d1, d2), seed<S>(a[i].y, 3 * i + 1, d1, d2),
                     seed<S>(a[i].z, 3 * i + 2, d1, d2));
}

// ---------------------------------------------------------------------------------------
// Sinks: where a primitive's results go. CompactSink parks them in the per-warp compact
// buffer P[entry * kPad + lane] (layout [q][J 3n][K block-upper], see ic_program.h).
// ---------------------------------------------------------------------------------------
struct CompactSink {
    double * Pq;    // value slot, lane folded in
    double * PJ;    // first J slot, lane folded in
    __device__ __forceinline__ CompactSink(double * P, int out, int lane)
        : Pq(P + out * kPad + lane), PJ(P + (out + 1) * kPad + lane) {}
    __device__ __forceinline__ void value(double v) const { Pq[0] = v; }
    __device__ __forceinline__ void jac_block(const int *, int i, const V3<double> & J) const {
        PJ[(3 * i + 0) * kPad] = J.x;
        PJ[(3 * i + 1) * kPad] = J.y;
        PJ[(3 * i + 2) * kPad] = J.z;
    }
    __device__ __forceinline__ void jac_entry(const int *, int m, double j) const { PJ[m * kPad] = j; }
    // tangent direction (atom j, component l) of block i of the J expression: kept when i <= j,
    // entry (k, l) of block (i, j)
    template <int N> __device__ __forceinline__ void hess_block(const int *, int dir, int i, const V3<Dual> & J) const {
        const int j = dir / 3, l = dir - 3 * j;
        if (i <= j) {
            double * blk = PJ + 3 * N * kPad + (9 * kblock(i, j, N) + l) * kPad;
            blk[0 * 3 * kPad] = J.x.d;
            blk[1 * 3 * kPad] = J.y.d;
            blk[2 * 3 * kPad] = J.z.d;
        }
    }
    // second derivative (m1 <= m2) of a 5-atom primitive
    __device__ __forceinline__ void hess_entry5(const int *, int m1, int m2, double k) const {
        const int i = m1 / 3, j = m2 / 3, kk = m1 - 3 * i, ll = m2 - 3 * j;
        double * blk = PJ + 15 * kPad + 9 * kblock(i, j, 5) * kPad;
        blk[(3 * kk + ll) * kPad] = k;
        if (i == j) blk[(3 * ll + kk) * kPad] = k;
    }
};

// 5-atom torsion2 family, kept out of line: it is rare and register hungry (Dual / Dual2 passes
// over 15 coordinates) and must not inflate the register allocation of the common types.
template <int MODE, class Sink, class Loader>
__device__ __noinline__ void eval_primitive5(const PrimDesc & pd, const Loader load, const Sink sink, int dir0, int dir1) {
    const int type = pd.type;
    V3<double> a[5];
#pragma unroll
    for (int i = 0; i < 5; i++) a[i] = load(pd.atoms[i]);
        // 5-atom torsion2 family: the reference differentiates the value expression with
    // autograd (source/IC/InvDisp.cpp:415-570 first order, :893-1065 second order).
    // px/py = sin(theta) * cos/sin(phi): the reference applies the product rule with an
    // analytic J(sin theta) (:530-536, :1022); differentiating the product directly is
    // the same function.
    double sv, cv, tv;
    torsion2_sincos<double>(a, sv, cv, tv);
    const bool use_cos = fabs(sv) > fabs(cv);       // :432, :904
    double sintheta = 0.0;
    if (type == TCHEM_PXTORSION2 || type == TCHEM_PYTORSION2) {
        V3<double> u12 = a[1] - a[0], u23 = a[2] - a[1];
        u12 = u12 / norm(u12); u23 = u23 / norm(u23);
        double ct = -dot(u12, u23);
        sintheta = sqrt(1.0 - ct * ct);
    }
    if (MODE != MODE_KDIR) {
        switch (type) {
        case TCHEM_TORSION2:    sink.value(signed_wrapped_angle(cv, tv, pd.min)); break;
        case TCHEM_SINTORSION2: sink.value(sv); break;
        case TCHEM_COSTORSION2: sink.value(cv); break;
        case TCHEM_PXTORSION2:  sink.value(sintheta * cv); break;
        default:                sink.value(sintheta * sv); break;
        }
    }
    if (MODE == MODE_QJ) {
        for (int dir = 0; dir < 15; dir++) {
            V3<Dual> s[5]; Dual sd, cd, td;
            seed_atoms<Dual, 5>(a, s, dir, -1);
            torsion2_sincos<Dual>(s, sd, cd, td);
            double j;
            if (type == TCHEM_TORSION2) j = use_cos ? cd.d / (-sv) : sd.d / cv;
            else if (type == TCHEM_SINTORSION2) j = sd.d;
            else if (type == TCHEM_COSTORSION2) j = cd.d;
            else {
                V3<Dual> u12 = s[1] - s[0], u23 = s[2] - s[1];
                u12 = u12 / norm(u12); u23 = u23 / norm(u23);
                Dual ct = -dot(u12, u23);
                Dual sth = sqrt_(1.0 - ct * ct);
                j = (sth * (type == TCHEM_PXTORSION2 ? cd : sd)).d;
            }
            sink.jac_entry(pd.atoms, dir, j);
        }
    } else if (MODE == MODE_QJK || MODE == MODE_KDIR) {
        for (int d2 = max(dir0, 0); d2 < min(dir1, 15); d2++) {
            for (int d1 = 0; d1 <= d2; d1++) {
                V3<Dual2> s[5]; Dual2 sd, cd, td;
                seed_atoms<Dual2, 5>(a, s, d1, d2);
                torsion2_sincos<Dual2>(s, sd, cd, td);
                double j2, k;      // d/d(d2) and d2/d(d1)d(d2)
                if (type == TCHEM_TORSION2) {
                    if (use_cos) {  // J = Jcos / -sin ; K = (Kcos + Jsin (x) J) / -sin   (:917-918)
                        j2 = cd.b / (-sv);
                        k = (cd.c + sd.a * j2) / (-sv);
                    } else {        // J = Jsin / cos ; K = (Ksin - Jcos (x) J) / cos     (:935-936)
                        j2 = sd.b / cv;
                        k = (sd.c - cd.a * j2) / cv;
                    }
                } else if (type == TCHEM_SINTORSION2) { j2 = sd.b; k = sd.c; }
                else if (type == TCHEM_COSTORSION2) { j2 = cd.b; k = cd.c; }
                else {
                    V3<Dual2> u12 = s[1] - s[0], u23 = s[2] - s[1];
                    u12 = u12 / norm(u12); u23 = u23 / norm(u23);
                    Dual2 ct = -dot(u12, u23);
                    Dual2 sth = sqrt_(1.0 - ct * ct);
                    Dual2 f = sth * (type == TCHEM_PXTORSION2 ? cd : sd);
                    j2 = f.b; k = f.c;
                }
                if (d1 == d2 && MODE != MODE_KDIR) sink.jac_entry(pd.atoms, d2, j2);
                sink.hess_entry5(pd.atoms, d1, d2, k);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// One invariant displacement for this lane's geometry.
//   MODE_VALUE : clamped value formulas of InvDisp::operator()
//   MODE_QPATH : value by the compute_IC_J formulas, no derivatives
//   MODE_QJ    : value + analytic J
//   MODE_QJK   : + one Dual pass per Cartesian direction through the same J expression
//                (what the reference asks autograd for, source/IC/InvDisp.cpp:859-868)
// ---------------------------------------------------------------------------------------
// HAS5: the table contains 5-atom (torsion2 family) primitives. Those are evaluated by an
// out-of-line function; a kernel instantiated with HAS5 = false carries no call at all, which
// keeps its long-lived state in registers (values live across a call site go to the stack).
template <int MODE, bool HAS5, class Sink, class Loader>
__device__ __forceinline__ void eval_primitive(const PrimDesc & pd, const Loader & load, const Sink & sink,
                                               int dir0 = 0, int dir1 = 15) {
    constexpr bool WANT_J = MODE == MODE_QJ || MODE == MODE_QJK;
    constexpr bool WANT_K = MODE == MODE_QJK || MODE == MODE_KDIR;
    constexpr bool WANT_Q = MODE != MODE_KDIR;
    const int type = pd.type;
    if (type == TCHEM_DUMMY) { if (WANT_Q) sink.value(1.0); return; }
    if (MODE != MODE_VALUE && pd.natoms == 5) {
        if (HAS5) eval_primitive5<MODE, Sink, Loader>(pd, load, sink, dir0, dir1);
        return;
    }
    V3<double> a[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (i < pd.natoms) a[i] = load(pd.atoms[i]);
    if (MODE == MODE_VALUE && pd.natoms == 5) {
        V3<double> a5[5] = {a[0], a[1], a[2], a[3], load(pd.atoms[4])};
        sink.value(value_path(type, a5, pd.min));
        return;
    }
    if (MODE == MODE_VALUE) {
        V3<double> a5[5] = {a[0], a[1], a[2], a[3], a[3]};
        sink.value(value_path(type, a5, pd.min));
        return;
    }

    // J blocks leave the registers as soon as they are formed
    auto emitJ = [&](int i, const V3<double> & v) { if (WANT_J) sink.jac_block(pd.atoms, i, v); };
    switch (type) {
    case TCHEM_STRETCHING: {
        if (WANT_Q) {
            double q;
            stretching_J<double>(a, q, emitJ);
            sink.value(q);
        }
        if (WANT_K) {
            for (int dir = dir0; dir < min(dir1, 6); dir++) {
                V3<Dual> s[2]; Dual qd;
                seed_atoms<Dual, 2>(a, s, dir, -1);
                stretching_J<Dual>(s, qd, [&](int i, const V3<Dual> & v) { sink.template hess_block<2>(pd.atoms, dir, i, v); });
            }
        }
        break;
    }
    case TCHEM_BENDING: case TCHEM_COSBENDING: {
        if (WANT_Q) {
            double c;
            if (type == TCHEM_BENDING) { bending_J<double, false>(a, c, emitJ); sink.value(acos(c)); }
            else                       { bending_J<double, true>(a, c, emitJ);  sink.value(c); }
        }
        if (WANT_K) {
            for (int dir = dir0; dir < min(dir1, 9); dir++) {
                V3<Dual> s[3]; Dual cd;
                seed_atoms<Dual, 3>(a, s, dir, -1);
                auto emitK = [&](int i, const V3<Dual> & v) { sink.template hess_block<3>(pd.atoms, dir, i, v); };
                if (type == TCHEM_BENDING) bending_J<Dual, false>(s, cd, emitK);
                else                       bending_J<Dual, true>(s, cd, emitK);
            }
        }
        break;
    }
    case TCHEM_TORSION: case TCHEM_SINTORSION: case TCHEM_COSTORSION: {
        if (WANT_Q) {
            double sp = 0.0, cp = 0.0, st = 0.0;
            if (type == TCHEM_TORSION) {
                torsion_J<double, 0>(a, sp, cp, st, emitJ);
                sink.value(signed_wrapped_angle(cp, st, pd.min));
            } else if (type == TCHEM_SINTORSION) { torsion_J<double, 1>(a, sp, cp, st, emitJ); sink.value(sp); }
            else                                 { torsion_J<double, 2>(a, sp, cp, st, emitJ); sink.value(cp); }
        }
        if (WANT_K) {
            for (int dir = dir0; dir < min(dir1, 12); dir++) {
                V3<Dual> s[4]; Dual spd, cpd, std_;
                seed_atoms<Dual, 4>(a, s, dir, -1);
                auto emitK = [&](int i, const V3<Dual> & v) { sink.template hess_block<4>(pd.atoms, dir, i, v); };
                if (type == TCHEM_TORSION)         torsion_J<Dual, 0>(s, spd, cpd, std_, emitK);
                else if (type == TCHEM_SINTORSION) torsion_J<Dual, 1>(s, spd, cpd, std_, emitK);
                else                               torsion_J<Dual, 2>(s, spd, cpd, std_, emitK);
            }
        }
        break;
    }
    case TCHEM_OUTOFPLANE: case TCHEM_SINOOP: {
        if (WANT_Q) {
            double s_;
            if (type == TCHEM_OUTOFPLANE) { oop_J<double, false>(a, s_, emitJ); sink.value(asin(s_)); }
            else                          { oop_J<double, true>(a, s_, emitJ);  sink.value(s_); }
        }
        if (WANT_K) {
            for (int dir = dir0; dir < min(dir1, 12); dir++) {
                V3<Dual> s[4]; Dual sd;
                seed_atoms<Dual, 4>(a, s, dir, -1);
                auto emitK = [&](int i, const V3<Dual> & v) { sink.template hess_block<4>(pd.atoms, dir, i, v); };
                if (type == TCHEM_OUTOFPLANE) oop_J<Dual, false>(s, sd, emitK);
                else                          oop_J<Dual, true>(s, sd, emitK);
            }
        }
        break;
    }
    default:
        break;
    }
}

// All distinct primitives of one chunk for this lane's geometry -> compact buffer.
template <int MODE, bool HAS5>
__device__ __forceinline__ void eval_chunk(const PrimDesc * prims, int p0, int p1, const double * R, double * P, int lane) {
    for (int p = p0; p < p1; p++) eval_primitive<MODE, HAS5>(prims[p], TileLoader{R, lane}, CompactSink(P, prims[p].out, lane));
}

// ---------------------------------------------------------------------------------------
// gather-store: dense output span of S elements per geometry, element k of geometry g is
//   sum_j coef_j * P[entry_j][g]     (no sources: structural zero)
// Lanes run along the (sector-sorted, see ElemMap) elements, the g loop walks the tile. Up to 4
// sources per element are held in registers; elements with more take further rounds that
// accumulate.
// ---------------------------------------------------------------------------------------
// read-only table loads through the non-coherent path
// (generic loads: the tables live in shared memory when they are small, in global memory otherwise)
__device__ __forceinline__ uint4 load_quad(const uint4 * p) { return *p; }
__device__ __forceinline__ uint32_t quad_get(const uint4 & q, int j) { return j == 0 ? q.x : j == 1 ? q.y : j == 2 ? q.z : q.w; }
__device__ __forceinline__ ElemMap load_emap(const ElemMap * p) {
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    ElemMap m;
    m.word = v.x; m.k = v.y;
    return m;
}

// one round of <= NS (<= 4) sources per element, all of them in the quad `q`
template <int NS, bool FULL, bool ACC>
__device__ __forceinline__ void gather_round(const uint4 & q, int cnt, const double * __restrict__ coefs,
                                             const double * P, char * o, int gbytes, int ntile) {
    double c[NS];
    const double * p[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) {
        c[j] = 0.0;
        p[j] = P;
        if (j < cnt) {
            const uint32_t s = quad_get(q, j);
            c[j] = coefs[s >> 16];
            p[j] = P + (s & kSrcEntryMask) * kPad;
        }
    }
    const int n = FULL ? kLanes : ntile;
#pragma unroll 8
    for (int g = 0; g < n; g++) {
        // (predicated on purpose: unconditional reads of a zero row measured slower - they add
        // shared-memory wavefronts to steps where most lanes have fewer than NS sources)
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < NS; j++)
            if (j < cnt) v = fma(c[j], p[j][g], v);
        double * og = reinterpret_cast<double *>(o + (int64_t)g * gbytes);
        if (ACC) v += *og;
        *og = v;
    }
}

// a stretch of structural zeros [0, len) of every geometry of the tile: 16-byte stores when the
// alignment allows, no table lookups
template <bool FULL>
__device__ __forceinline__ void zero_run(double * base, uint32_t len, int gbytes, int ntile, int lane) {
    const int n = FULL ? kLanes : ntile;
    char * row = reinterpret_cast<char *>(base);
    if (((reinterpret_cast<uintptr_t>(base) | (uint32_t)gbytes) & 15) == 0) {
        const uint32_t pairs = len >> 1;
        for (int g = 0; g < n; g++, row += gbytes) {
            double2 * p = reinterpret_cast<double2 *>(row);
#pragma unroll 4
            for (uint32_t e = lane; e < pairs; e += kLanes) p[e] = make_double2(0.0, 0.0);
            if ((len & 1) && lane == 0) reinterpret_cast<double *>(row)[len - 1] = 0.0;
        }
    } else {
        for (int g = 0; g < n; g++, row += gbytes) {
            double * p = reinterpret_cast<double *>(row);
#pragma unroll 4
            for (uint32_t e = lane; e < len; e += kLanes) p[e] = 0.0;
        }
    }
}

// the byte distance between consecutive geometries of a tile fits 31 bits (checked on the host).
// Table reads are software-pipelined: while step s runs, the element words of steps s+1 and s+2
// and the source quad of step s+1 are already in flight.
// the first two element words of a span, fetched before the chunk's arithmetic so that their (L2)
// latency is not paid at the start of every write-out
struct SpanPrefetch { ElemMap m1, m2; };
__device__ __forceinline__ SpanPrefetch span_prefetch(const SpanDesc & sp, const ElemMap * __restrict__ emap_all, int lane) {
    SpanPrefetch pf;
    pf.m1.word = 0u; pf.m1.k = 0u; pf.m2 = pf.m1;
    const ElemMap * emap = emap_all + sp.map;
    if (lane < sp.n) pf.m1 = load_emap(emap + lane);
    if (lane + kLanes < sp.n) pf.m2 = load_emap(emap + lane + kLanes);
    return pf;
}

template <bool FULL, bool ACC, bool ZSEC>
__device__ __forceinline__ void gather_store_impl(const SpanDesc & sp, const ElemMap * __restrict__ emap_all,
                                                  const uint4 * __restrict__ psrcs, const double * __restrict__ coefs,
                                                  const ZeroRun * __restrict__ runs, const double * P, double * out,
                                                  int gbytes, int ntile, int lane, const SpanPrefetch & pf) {
    if (!ACC)
        for (int i = 0; i < sp.nruns; i++) {
            const ZeroRun run = runs[sp.run0 + i];
            zero_run<FULL>(out + run.k, run.len, gbytes, ntile, lane);
        }
    const ElemMap * emap = emap_all + sp.map;
    const int S = sp.n;
    const bool al32 = (reinterpret_cast<uintptr_t>(out) & 31) == 0;     // (the geometry stride is a multiple of 32 bytes whenever sector entries exist)
    ElemMap m1 = pf.m1, m2 = pf.m2;
    uint4 q1 = make_uint4(0u, 0u, 0u, 0u);
    if (m1.word >> kMapCountShift) q1 = load_quad(psrcs + (m1.word & kMapOffsetMask));
    for (int k0 = 0; k0 < S; k0 += kLanes) {
        const bool valid = k0 + lane < S;
        const ElemMap m = m1;
        const uint4 q = q1;
        m1 = m2;
        m2.word = 0u; m2.k = 0u;
        if (k0 + 2 * kLanes + lane < S) m2 = load_emap(emap + k0 + 2 * kLanes + lane);
        if (m1.word >> kMapCountShift) q1 = load_quad(psrcs + (m1.word & kMapOffsetMask));
        const int cnt = (int)(m.word >> kMapCountShift);
        const uint32_t off = m.word & kMapOffsetMask;
        char * o = reinterpret_cast<char *>(out + m.k);
        const int maxcnt = __reduce_max_sync(kFull, valid ? cnt : 0);
        if (ZSEC && valid && !ACC && m.word == kZeroSectorWord) {
            // a whole sector of zeros: one 32-byte store per geometry (its first element is written
            // once more, as 0.0, if this step also holds non-zero elements: harmless)
            const int n = FULL ? kLanes : ntile;
            // the 32-byte store needs the STORE address aligned, not just the chunk base: sectors are aligned to the
            // geometry block, whose offset from `out` need not be a multiple of 4 elements
            if (al32 && (reinterpret_cast<uintptr_t>(o) & 31) == 0 && (gbytes & 31) == 0) {
#pragma unroll 8
                for (int g = 0; g < n; g++)
                    asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};\n" ::"l"(o + (int64_t)g * gbytes), "d"(0.0) : "memory");
            } else {                               // caller's buffer is not 32-byte aligned
                for (int g = 0; g < n; g++) {
                    double * og = reinterpret_cast<double *>(o + (int64_t)g * gbytes);
                    og[0] = 0.0; og[1] = 0.0; og[2] = 0.0; og[3] = 0.0;
                }
            }
        }
        if (maxcnt == 0) {
            if (valid && !ACC && !(ZSEC && m.word == kZeroSectorWord)) {
                const int n = FULL ? kLanes : ntile;
#pragma unroll 8
                for (int g = 0; g < n; g++) *reinterpret_cast<double *>(o + (int64_t)g * gbytes) = 0.0;
            }
            continue;
        }
        if (!valid) continue;
        {   // first round
            if (maxcnt == 1)      gather_round<1, FULL, ACC>(q, cnt, coefs, P, o, gbytes, ntile);
            else if (maxcnt == 2) gather_round<2, FULL, ACC>(q, cnt, coefs, P, o, gbytes, ntile);
            else                  gather_round<4, FULL, ACC>(q, cnt, coefs, P, o, gbytes, ntile);
        }
        for (int base = 4; base < maxcnt; base += 4) {     // rare: more than 4 sources
            uint4 q2 = make_uint4(0u, 0u, 0u, 0u);
            if (cnt > base) q2 = load_quad(psrcs + off + base / 4);
            gather_round<4, FULL, true>(q2, cnt - base, coefs, P, o, gbytes, ntile);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Line-buffer write-out (J spans of moderate width, <= 64 elements in non-zero sectors, <= 4 sources
// each): instead of walking the tile's 32 geometries element step by element step - 8-byte stores
// one geometry stride apart, the pattern that caps the HBM write stream at ~4.7 TB/s on B200 - the
// dense piece of ONE geometry is assembled in a per-warp shared-memory line (zero background
// written once, the non-zero positions are the same for every geometry) and streamed out with
// consecutive 16-byte stores, geometry after geometry (measured ceiling of that pattern: 5.4-5.8 TB/s).
// ---------------------------------------------------------------------------------------
constexpr int kLineMaxNnz = 2 * kLanes;
__device__ __forceinline__ bool line_eligible(const SpanDesc & sp) { return sp.nnz >= 0 && (sp.nnz & 0xffff) <= kLineMaxNnz; }
template <int NS>
struct LineSrc {
    double c[NS];
    uint32_t p[NS];                // compact entry * kPad
    uint32_t k;
    bool live;
};
template <int NS>
__device__ __forceinline__ void line_sources(const ElemMap & m, bool live, const uint4 * __restrict__ psrcs,
                                             const double * __restrict__ coefs, LineSrc<NS> & s) {
    const int cnt = live ? (int)(m.word >> kMapCountShift) : 0;
    s.k = m.k;
    s.live = live;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (cnt) q = load_quad(psrcs + (m.word & kMapOffsetMask));
#pragma unroll
    for (int j = 0; j < NS; j++) {
        // an absent source reads entry 0 with coefficient 0: no predicate inside the geometry loop
        const uint32_t w = quad_get(q, j);
        s.c[j] = j < cnt ? coefs[w >> 16] : 0.0;
        s.p[j] = j < cnt ? (w & kSrcEntryMask) * kPad : 0u;
    }
}
template <int NS>
__device__ __forceinline__ double line_value(const LineSrc<NS> & s, const double * P, int g) {
    double v = s.c[0] * P[s.p[0] + g];
#pragma unroll
    for (int j = 1; j < NS; j++) v = fma(s.c[j], P[s.p[j] + g], v);
    return v;
}
template <bool FULL, int NS>
__device__ __forceinline__ void gather_store_line_ns(const SpanDesc & sp, const uint4 * __restrict__ psrcs,
                                                     const double * __restrict__ coefs, const double * P, double * L, int E,
                                                     double * out, int64_t gstride, int ntile, int lane, const SpanPrefetch & pf) {
    const int Epad = (E + 1) & ~1, nnz = sp.nnz & 0xffff;
    // two line buffers: while geometry g streams out of one, the values of g + 1 are gathered into
    // the other (one warp barrier per geometry, the loads of the next step already in flight)
    double * L1 = L + Epad;
    for (int e = 2 * lane; e < 2 * Epad; e += 2 * kLanes) *reinterpret_cast<double2 *>(L + e) = make_double2(0.0, 0.0);
    LineSrc<NS> s0, s1;
    line_sources<NS>(pf.m1, lane < nnz, psrcs, coefs, s0);
    line_sources<NS>(pf.m2, lane + kLanes < nnz, psrcs, coefs, s1);
    const bool two = nnz > kLanes;                      // warp-uniform
    const bool vec = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && ((gstride & 1) == 0);
    const int n = FULL ? kLanes : ntile;
    const int e0 = 2 * lane;
    double v0 = s0.live ? line_value<NS>(s0, P, 0) : 0.0;
    double v1 = (two && s1.live) ? line_value<NS>(s1, P, 0) : 0.0;
    __syncwarp();
    for (int g = 0; g < n; g++) {
        double * Lg = (g & 1) ? L1 : L;
        if (s0.live) Lg[s0.k] = v0;
        if (two && s1.live) Lg[s1.k] = v1;
        if (g + 1 < n) {
            if (s0.live) v0 = line_value<NS>(s0, P, g + 1);
            if (two && s1.live) v1 = line_value<NS>(s1, P, g + 1);
        }
        __syncwarp();
        double * og = out + (int64_t)g * gstride;
        if (vec) {
            // pieces up to 256 doubles: four independent 16-byte copies per lane, then the rest
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (e0 + 64 * u + 1 < E) *reinterpret_cast<double2 *>(og + e0 + 64 * u) = *reinterpret_cast<const double2 *>(Lg + e0 + 64 * u);
            for (int e = e0 + 256; e + 1 < E; e += 2 * kLanes)
                *reinterpret_cast<double2 *>(og + e) = *reinterpret_cast<const double2 *>(Lg + e);
            if ((E & 1) && lane == 0) og[E - 1] = Lg[E - 1];
        } else {
            for (int e = lane; e < E; e += kLanes) og[e] = Lg[e];
        }
    }
    __syncwarp();
}
// (out of line on purpose: at the call site only a few pointers are live, and inside the write-out gets
// a register allocation of its own instead of competing with the primitive evaluation's)
template <bool FULL>
__device__ __noinline__ void gather_store_line(const SpanDesc & sp, const uint4 * __restrict__ psrcs,
                                                  const double * __restrict__ coefs, const double * P, double * L, int E,
                                                  double * out, int64_t gstride, int ntile, int lane, const SpanPrefetch & pf) {
    const int maxcnt = sp.nnz >> 16;
    if (maxcnt <= 1)      gather_store_line_ns<FULL, 1>(sp, psrcs, coefs, P, L, E, out, gstride, ntile, lane, pf);
    else if (maxcnt == 2) gather_store_line_ns<FULL, 2>(sp, psrcs, coefs, P, L, E, out, gstride, ntile, lane, pf);
    else                  gather_store_line_ns<FULL, 4>(sp, psrcs, coefs, P, L, E, out, gstride, ntile, lane, pf);
}

// `coefs` = the table of distinct coefficients (shared-memory copy when it is small).
// ZSEC: the span may hold zero-sector entries (K spans; see emit_span)
template <bool ZSEC = false>
__device__ __forceinline__ void gather_store(const SpanDesc & sp, const ElemMap * __restrict__ emap, const uint4 * __restrict__ psrcs,
                                             const double * __restrict__ coefs, const ZeroRun * __restrict__ runs,
                                             const double * P, double * out, int64_t gstride, int ntile, bool accumulate, int lane,
                                             const SpanPrefetch & pf) {
    const int gbytes = (int)(gstride * sizeof(double));
    if (accumulate) {
        if (ntile == kLanes) gather_store_impl<true, true, ZSEC>(sp, emap, psrcs, coefs, runs, P, out, gbytes, ntile, lane, pf);
        else                 gather_store_impl<false, true, ZSEC>(sp, emap, psrcs, coefs, runs, P, out, gbytes, ntile, lane, pf);
    } else {
        if (ntile == kLanes) gather_store_impl<true, false, ZSEC>(sp, emap, psrcs, coefs, runs, P, out, gbytes, ntile, lane, pf);
        else                 gather_store_impl<false, false, ZSEC>(sp, emap, psrcs, coefs, runs, P, out, gbytes, ntile, lane, pf);
    }
}
template <bool ZSEC = false>
__device__ __forceinline__ void gather_store(const SpanDesc & sp, const ElemMap * __restrict__ emap, const uint4 * __restrict__ psrcs,
                                             const double * __restrict__ coefs, const ZeroRun * __restrict__ runs,
                                             const double * P, double * out, int64_t gstride, int ntile, bool accumulate, int lane) {
    gather_store<ZSEC>(sp, emap, psrcs, coefs, runs, P, out, gstride, ntile, accumulate, lane, span_prefetch(sp, emap, lane));
}

// ---------------------------------------------------------------------------------------
// q rows of a chunk: every lane combines its own geometry, q[row] = sum coef * q_prim
// (source/IC/IntCoord.cpp:53-60 / :66-74), parks the rows in the chunk's q slots and the warp
// writes the [ntile][nrows] block with lanes running over (geometry, row) pairs.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void combine_rows(const uint32_t * __restrict__ qmap, const Src * __restrict__ srcs,
                                             int nrows, double * P, int qslot, int lane) {
    for (int row = 0; row < nrows; row++) {
        const uint32_t w = qmap[row];
        const int cnt = (int)(w >> kMapCountShift);
        const uint32_t off = w & kMapOffsetMask;
        double x = 0.0;
        for (int j = 0; j < cnt; j++) {
            const Src s = srcs[off + j];
            x = fma(s.coef, P[s.entry * kPad + lane], x);
        }
        P[(qslot + row) * kPad + lane] = x;
    }
}

__device__ __forceinline__ void store_rows(const double * P, int qslot, int nrows, double * out /* &q[b0][row0] */,
                                           int64_t gstride, int ntile, bool accumulate, int lane) {
    const int total = ntile * nrows;
    int g = lane / nrows, row = lane - g * nrows;
    const int dg = kLanes / nrows, dr = kLanes - dg * nrows;
    for (int f = lane; f < total; f += kLanes) {
        double v = P[(qslot + row) * kPad + g];
        double * o = out + (int64_t)g * gstride + row;
        if (accumulate) v += *o;
        *o = v;
        g += dg; row += dr;
        if (row >= nrows) { row -= nrows; g++; }
    }
}

} // namespace tchem_b200
