"""CPU check of the host-side "compiler" (torch_chemistry_b200/csrc/ic_table.cpp): the program it
emits - distinct primitives per chunk, compact entry layout, q maps, sector-sorted element maps -
is exported through the C ABI and interpreted here in numpy with the ORACLE supplying each
primitive's value / J / K. The dense q, J, K this reconstructs must equal the oracle's own. This
pins the chunking, de-duplication and gather logic without a GPU."""
import ctypes as C

import numpy as np
import pytest

from conftest import assert_parity, golden
from oracle import tchem_oracle as O

capi = pytest.importorskip("torch_chemistry_b200._capi")
lib = capi.lib


class PrimDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("natoms", C.c_int32), ("atoms", C.c_int32 * 5), ("out", C.c_int32), ("min", C.c_double)]


class SpanDesc(C.Structure):
    _fields_ = [("map", C.c_int64), ("n", C.c_int32), ("run0", C.c_int32), ("nruns", C.c_int32), ("nnz", C.c_int32)]


class ChunkDesc(C.Structure):
    _fields_ = [("prim_begin", C.c_int32), ("prim_end", C.c_int32), ("row0", C.c_int32), ("nrows", C.c_int32),
                ("accumulate", C.c_int32), ("qslot", C.c_int32), ("qmap", C.c_int64), ("jspan", SpanDesc), ("kspan", SpanDesc)]


class ZeroRun(C.Structure):
    _fields_ = [("k", C.c_uint32), ("len", C.c_uint32)]


class Src(C.Structure):
    _fields_ = [("entry", C.c_int32), ("pad", C.c_int32), ("coef", C.c_double)]


class ElemMap(C.Structure):
    _fields_ = [("word", C.c_uint32), ("k", C.c_uint32)]


def export(text, fmt, cartdim, mode):
    nt, nic = C.c_int(0), C.c_int(0)
    capi.check(lib.tchem_ic_parse_text(fmt.encode(), text.encode(), None, 0, C.byref(nt), C.byref(nic)))
    terms = (capi.InvDisp * nt.value)()
    capi.check(lib.tchem_ic_parse_text(fmt.encode(), text.encode(), terms, nt.value, C.byref(nt), C.byref(nic)))
    counts = (C.c_int64 * 16)()
    capi.check(lib.tchem_ic_program_export(terms, nt.value, nic.value, cartdim, mode, counts, None, None, None, None, None, None, None, None))
    prims, chunks = (PrimDesc * counts[0])(), (ChunkDesc * counts[1])()
    qmap, qsrcs = (C.c_uint32 * max(counts[2], 1))(), (Src * max(counts[3], 1))()
    emap, psrcs = (ElemMap * max(counts[4], 1))(), (C.c_uint32 * max(counts[5], 1))()
    runs, coefs = (ZeroRun * max(counts[7], 1))(), (C.c_double * max(counts[8], 1))()
    capi.check(lib.tchem_ic_program_export(terms, nt.value, nic.value, cartdim, mode, counts, prims, chunks, qmap, qsrcs, emap, psrcs, runs, coefs))
    return dict(prims=prims, chunks=chunks, qmap=qmap, qsrcs=qsrcs, emap=emap, psrcs=psrcs, coefs=coefs, runs=runs, cap=counts[6],
                nic=nic.value, nruns=counts[7])


def interpret(prog, r, mode):
    """numpy twin of the device interpreter: compact buffer per chunk, then the gathers"""
    B, cd = r.shape
    n = prog["nic"]
    q = np.full((B, n), np.nan)
    J = np.full((B, n, cd), np.nan) if mode >= 1 else None
    K = np.full((B, n, cd, cd), np.nan) if mode == 2 else None
    for ch in prog["chunks"]:
        P = np.zeros((prog["cap"], B))
        for p in range(ch.prim_begin, ch.prim_end):
            pd = prog["prims"][p]
            d = O.InvDisp(O.TYPES[pd.type], [pd.atoms[i] for i in range(pd.natoms)], pd.min)
            na = pd.natoms
            if mode == 0:
                P[pd.out] = O.invdisp_value(d, r)
                continue
            qv, Jl, Kl = O.invdisp_qJK(d, r, mode == 2)
            P[pd.out] = qv
            for m in range(3 * na):
                P[pd.out + 1 + m] = Jl[:, m]
            if mode == 2:
                kbase = pd.out + 1 + 3 * na
                for i in range(na):
                    for j in range(i, na):
                        blk = i * na - (i * (i - 1)) // 2 + (j - i)
                        for c in range(3):
                            for e in range(3):
                                P[kbase + 9 * blk + 3 * c + e] = Kl[:, 3 * i + c, 3 * j + e]
        assert ch.qslot + ch.nrows <= prog["cap"]

        def gather(word, table):                        # q rows: 16-byte (entry, coefficient) records
            cnt, off = word >> 24, word & 0xFFFFFF
            v = np.zeros(B)
            for s in range(off, off + cnt):
                v = v + table[s].coef * P[table[s].entry]
            return v

        def gather_packed(word):                        # J / K elements: packed sources, quad aligned
            cnt, quad = word >> 24, word & 0xFFFFFF
            v = np.zeros(B)
            for s in range(4 * quad, 4 * quad + cnt):
                w = prog["psrcs"][s]
                v = v + prog["coefs"][w >> 16] * P[w & 0xFFFF]
            return v

        for row in range(ch.nrows):
            v = gather(prog["qmap"][ch.qmap + row], prog["qsrcs"])
            q[:, ch.row0 + row] = v + (q[:, ch.row0 + row] if ch.accumulate else 0.0)
        for name, out, span, sd in (("J", J, ch.nrows * cd, ch.jspan), ("K", K, ch.nrows * cd * cd, ch.kspan)):
            if out is None:
                continue
            flat = out.reshape(B, -1)
            base = ch.row0 * (cd if name == "J" else cd * cd)
            seen = np.zeros(span, dtype=bool)
            for i in range(sd.nruns):                       # long stretches of structural zeros
                run = prog["runs"][sd.run0 + i]
                assert run.len >= 64 and not seen[run.k:run.k + run.len].any()
                seen[run.k:run.k + run.len] = True
                if not ch.accumulate:
                    flat[:, base + run.k:base + run.k + run.len] = 0.0
            for e in range(sd.n):
                m = prog["emap"][sd.map + e]
                if m.word == 0x00FFFFFF:                    # a whole 32-byte sector of structural zeros
                    assert (base + m.k) % 4 == 0 and not seen[m.k:m.k + 4].any()
                    seen[m.k:m.k + 4] = True
                    if not ch.accumulate:
                        flat[:, base + m.k:base + m.k + 4] = 0.0
                    continue
                assert not seen[m.k], "element emitted twice"
                seen[m.k] = True
                v = gather_packed(m.word)
                flat[:, base + m.k] = v + (flat[:, base + m.k] if ch.accumulate else 0.0)
            assert seen.all(), "element map and zero runs do not cover the span"
    return q, J, K


CASES = [("ref_aniline_default", 2), ("ref_aniline_advanced", 2), ("ref_normal_mode", 1), ("workload_C2", 2), ("workload_C3", 1)]


@pytest.mark.parametrize("name,mode", CASES)
def test_compiled_program_reproduces_the_oracle(name, mode):
    g = golden(name)
    fmt = str(g["format"]) if "format" in g else "default"
    text = str(g["definition"])
    r = g["r"][:2]
    prog = export(text, fmt, r.shape[1], mode)
    orc = O.IntCoordSet(fmt, text)
    q, J, K = interpret(prog, r, mode)
    if mode == 2:
        qo, Jo, Ko = orc.qJK(r)
        assert_parity(K, Ko, name + " K through the program")
    else:
        qo, Jo = orc.qJ(r)
    assert_parity(q, qo, name + " q through the program")
    assert_parity(J, Jo, name + " J through the program")
    # value-path program
    q0, _, _ = interpret(export(text, fmt, r.shape[1], 0), r, 0)
    assert_parity(q0, orc.q(r), name + " q (value path) through the program")


def test_primitives_are_shared_inside_a_chunk():
    """C2: 41 terms but 15 distinct invariant displacements; the four C-H stretch combinations
    share their four stretches."""
    g = golden("workload_C2")
    prog = export(str(g["definition"]), "default", 18, 1)
    assert len(prog["prims"]) == 15
    assert sum(ch.nrows for ch in prog["chunks"]) == 12


def test_row_wider_than_capacity_is_split_and_accumulated():
    # one coordinate = 8 torsions: 8 * 13 + 1 entries exceed the q+J capacity target of 64
    lines = []
    for k in range(8):
        head = "%6d" % 1 if k == 0 else " " * 6
        lines.append(head + "%12.6f%14s%6d%6d%6d%6d" % (1.0, "torsion", k + 1, k + 2, k + 3, k + 4))
    text = "\n".join(lines) + "\n"
    prog = export(text, "default", 33, 1)
    assert len(prog["chunks"]) >= 2 and [ch.accumulate for ch in prog["chunks"]][1:] == [1] * (len(prog["chunks"]) - 1)
    rng = np.random.default_rng(3)
    r = rng.standard_normal((2, 33)) * 2.0
    q, J, _ = interpret(prog, r, 1)
    qo, Jo = O.IntCoordSet("default", text).qJ(r)
    assert_parity(J, Jo, "split row J")
    assert np.abs(q - qo).max() < 1e-9


def test_specialised_kernel_source_compiles_for_sm100a():
    """jit.cu: the generated CUDA C++ for a definition is valid and NVRTC compiles it for sm_100a
    without a device (skipped when libnvrtc is not installed)."""
    for name, cd in (("workload_C2", 18), ("workload_C1", 9), ("ref_aniline_default", 42)):
        g = golden(name)
        text = str(g["definition"])
        nt, nic = C.c_int(0), C.c_int(0)
        capi.check(lib.tchem_ic_parse_text(b"default", text.encode(), None, 0, C.byref(nt), C.byref(nic)))
        terms = (capi.InvDisp * nt.value)()
        capi.check(lib.tchem_ic_parse_text(b"default", text.encode(), terms, nt.value, C.byref(nt), C.byref(nic)))
        n = lib.tchem_ic_jit_source(terms, nt.value, nic.value, cd, 0, None, 0)
        assert n > 0
        buf = C.create_string_buffer(n)
        lib.tchem_ic_jit_source(terms, nt.value, nic.value, cd, 0, buf, n)
        src = buf.value.decode()
        assert "tchem_jit_qJ" in src and "stage_issue" in src
        assert src.count("stretching_J<double>") >= 1
        log = C.create_string_buffer(1 << 16)
        rc = lib.tchem_ic_jit_compile_check(src.encode(), log, 1 << 16)
        if rc != 0 and "not available" in log.value.decode():
            pytest.skip("NVRTC not available")
        assert rc == 0, log.value.decode()[-2000:]
        assert "Used" in log.value.decode()
    # tables with 5-atom (torsion2 family) terms are not eligible: the interpreter serves them
    a = golden("ref_aniline_advanced")
    text = str(a["definition"])
    capi.check(lib.tchem_ic_parse_text(b"whatever", text.encode(), None, 0, C.byref(nt), C.byref(nic)))
    terms = (capi.InvDisp * nt.value)()
    capi.check(lib.tchem_ic_parse_text(b"whatever", text.encode(), terms, nt.value, C.byref(nt), C.byref(nic)))
    assert lib.tchem_ic_jit_source(terms, nt.value, nic.value, 42, 0, None, 0) == 0


def test_specialised_sap_second_order_source_compiles_for_sm100a():
    """jit.cu: the generated kernel for SAPSet::Jacobian2nd_ of the C5 set (literal products, whole-geometry
    staging, one cp.async.bulk store per tile) compiles with NVRTC for sm_100a without a device."""
    g = golden("workload_C5")
    text, dims = str(g["sap_definition"]), [int(d) for d in g["sap_dims"]]
    nt, nc = C.c_int(0), C.c_int(0)
    capi.check(lib.tchem_sap_parse_text(text.encode(), None, 0, None, 0, C.byref(nt), C.byref(nc)))
    tp, cs = (C.c_int32 * (nt.value + 1))(), (C.c_int32 * max(2 * nc.value, 2))()
    capi.check(lib.tchem_sap_parse_text(text.encode(), tp, nt.value, cs, nc.value, C.byref(nt), C.byref(nc)))
    d = (C.c_int32 * len(dims))(*dims)
    sources = []
    n = lib.tchem_sap_jit_hess_source(tp, cs, nt.value, d, len(dims), None, 0)
    assert n > 0
    buf = C.create_string_buffer(n)
    lib.tchem_sap_jit_hess_source(tp, cs, nt.value, d, len(dims), buf, n)
    sources.append(("tchem_jit_sap_hess", buf.value.decode()))
    # value + Jacobian on given coordinates: the warp copy-out and the TMA bulk-store variants
    for bulk, name in ((0, "tchem_jit_chain"), (1, "tchem_jit_chain_bulk")):
        n = lib.tchem_sap_jit_source(tp, cs, nt.value, d, len(dims), bulk, None, 0)
        assert n > 0
        buf = C.create_string_buffer(n)
        lib.tchem_sap_jit_source(tp, cs, nt.value, d, len(dims), bulk, buf, n)
        sources.append((name, buf.value.decode()))
    for name, src in sources:
        assert name in src and ("cp.async.bulk.global.shared::cta" in src) == (name != "tchem_jit_chain")
        log = C.create_string_buffer(1 << 16)
        rc = lib.tchem_ic_jit_compile_check(src.encode(), log, 1 << 16)
        if rc != 0 and "not available" in log.value.decode():
            pytest.skip("NVRTC not available")
        assert rc == 0, name + ": " + log.value.decode()[-2000:]


def test_chained_cartesian_jacobian_source_compiles_for_sm100a():
    """jit.cu: the (IntCoordSet, SAPSet) pair kernel with the Jacobian in Cartesian coordinates - J's structurally
    nonzero elements as named values, every output element a literal fma chain - for the C5 pair and for C2H4 chained
    into a small set; compiled with NVRTC for sm_100a without a device."""
    g = golden("workload_C5")
    cases = [(str(g["definition"]), 9, str(g["sap_definition"]), [int(d) for d in g["sap_dims"]])]
    c2 = golden("workload_C2")
    cases.append((str(c2["definition"]), 18, "1,1\n1,2    1,1\n1,12    1,12\n1,3    1,2    1,1\n", [12]))
    for ic_text, cd, sap_text, dims in cases:
        nt, nic = C.c_int(0), C.c_int(0)
        capi.check(lib.tchem_ic_parse_text(b"default", ic_text.encode(), None, 0, C.byref(nt), C.byref(nic)))
        terms = (capi.InvDisp * nt.value)()
        capi.check(lib.tchem_ic_parse_text(b"default", ic_text.encode(), terms, nt.value, C.byref(nt), C.byref(nic)))
        ns, nc = C.c_int(0), C.c_int(0)
        capi.check(lib.tchem_sap_parse_text(sap_text.encode(), None, 0, None, 0, C.byref(ns), C.byref(nc)))
        tp, cs = (C.c_int32 * (ns.value + 1))(), (C.c_int32 * max(2 * nc.value, 2))()
        capi.check(lib.tchem_sap_parse_text(sap_text.encode(), tp, ns.value, cs, nc.value, C.byref(ns), C.byref(nc)))
        d = (C.c_int32 * len(dims))(*dims)
        n = lib.tchem_ic_sap_jit_cart_source(terms, nt.value, nic.value, cd, tp, cs, ns.value, d, len(dims), None, 0)
        assert n > 0
        buf = C.create_string_buffer(n)
        lib.tchem_ic_sap_jit_cart_source(terms, nt.value, nic.value, cd, tp, cs, ns.value, d, len(dims), buf, n)
        src = buf.value.decode()
        assert "tchem_jit_chain_cart" in src and "J0_0" in src and "#define SDO %d" % cd in src
        log = C.create_string_buffer(1 << 16)
        rc = lib.tchem_ic_jit_compile_check(src.encode(), log, 1 << 16)
        if rc != 0 and "not available" in log.value.decode():
            pytest.skip("NVRTC not available")
        assert rc == 0, log.value.decode()[-2000:]
        # the same pair with whole-tile TMA bulk stores (needs an even number of monomials)
        n = lib.tchem_ic_sap_jit_cart_bulk_source(terms, nt.value, nic.value, cd, tp, cs, ns.value, d, len(dims), None, 0)
        if ns.value % 2:
            assert n == 0
            continue
        assert n > 0
        buf = C.create_string_buffer(n)
        lib.tchem_ic_sap_jit_cart_bulk_source(terms, nt.value, nic.value, cd, tp, cs, ns.value, d, len(dims), buf, n)
        src = buf.value.decode()
        assert "tchem_jit_chain_cart_bulk" in src and "cp.async.bulk.global.shared::cta" in src and "#define SDO %d" % cd in src
        rc = lib.tchem_ic_jit_compile_check(src.encode(), log, 1 << 16)
        assert rc == 0, log.value.decode()[-2000:]
