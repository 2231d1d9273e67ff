"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/tchem_b200.h
declares, parses the reference's definition formats exactly like the reference, maps errors to
the reference's exception types, and refuses to compute without a CUDA device (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import EPS, ROOT, golden, terms_from_golden

capi = pytest.importorskip("torch_chemistry_b200._capi")
lib = capi.lib


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tchem_b200.h")).read()
    return sorted(set(re.findall(r"TCHEM_API[^;(]*?\b(tchem_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = header_symbols()
    assert len(names) >= 27
    assert sorted(capi.DECLARED) == names
    for n in names:
        assert hasattr(lib, n), n


def parse_text(fmt, text):
    nt, nic = C.c_int(0), C.c_int(0)
    capi.check(lib.tchem_ic_parse_text(fmt.encode(), text.encode(), None, 0, C.byref(nt), C.byref(nic)))
    buf = (capi.InvDisp * max(nt.value, 1))()
    capi.check(lib.tchem_ic_parse_text(fmt.encode(), text.encode(), buf, nt.value, C.byref(nt), C.byref(nic)))
    return [buf[i] for i in range(nt.value)], nic.value


@pytest.mark.parametrize("name", ["ref_aniline_default", "ref_aniline_columbus7", "ref_aniline_advanced", "ref_normal_mode"])
def test_parser_matches_reference_tables(name):
    g = golden(name)
    terms, nic = parse_text(str(g["format"]), str(g["definition"]))
    ref = terms_from_golden(g)
    assert len(terms) == len(ref)
    assert nic == g["q"].shape[1]
    for t, (ic, coeff, typ, atoms, mn) in zip(terms, ref):
        assert t.ic == ic and capi.TYPE_NAMES[t.type] == typ
        assert [t.atoms[i] for i in range(t.natoms)] == atoms
        assert abs(t.coeff - coeff) <= 4 * EPS * abs(coeff)
        assert t.min == mn


def test_parser_drops_unterminated_last_line_like_getline():
    # source/IC/IntCoordSet.cpp:64-65: `std::getline; if (!ifs.good()) break;`
    terms, nic = parse_text("default", "     1    1.0    stretching     1     2\n     2    1.0    stretching     1     3")
    assert nic == 1 and len(terms) == 1


def test_parser_errors_map_to_reference_exceptions(tmp_path):
    with pytest.raises(ValueError, match="unsupported internal coordinate type"):
        parse_text("default", "     1    1.0    wagging     1     2\n")
    nt, nic = C.c_int(0), C.c_int(0)
    rc = lib.tchem_ic_parse_file(b"default", str(tmp_path / "missing").encode(), None, 0, C.byref(nt), C.byref(nic))
    assert rc == capi.EFILE
    with pytest.raises(capi.FileError):
        capi.check(rc)


def test_sap_parser_known_fixture():
    g = golden("ref_sap_1")
    text = str(g["definition"])
    nt, nc = C.c_int(0), C.c_int(0)
    capi.check(lib.tchem_sap_parse_text(text.encode(), None, 0, None, 0, C.byref(nt), C.byref(nc)))
    assert nt.value == 9 and nc.value == 2 * 1 + 6 * 2
    tp = (C.c_int32 * (nt.value + 1))()
    cs = (C.c_int32 * (2 * nc.value))()
    capi.check(lib.tchem_sap_parse_text(text.encode(), tp, nt.value, cs, nc.value, C.byref(nt), C.byref(nc)))
    assert list(tp) == [0, 0, 1, 2, 4, 6, 8, 10, 12, 14]
    # "1,2  1,1" -> sorted descending (SAP.cpp:51): (0,1),(0,0)
    assert [cs[i] for i in range(2 * tp[4], 2 * tp[5])] == [0, 1, 0, 0]


def test_no_cpu_fallback():
    """without a CUDA device the table cannot even be created: there is no CPU path to fall into"""
    if lib.tchem_device_count() > 0:
        pytest.skip("a CUDA device is present")
    terms, nic = parse_text("default", "     1    1.0    stretching     1     2\n")
    arr = (capi.InvDisp * len(terms))(*terms)
    h = C.c_void_p()
    rc = lib.tchem_ic_table_create(arr, len(terms), nic, 0, 0, C.byref(h))
    assert rc == capi.ECUDA
    with pytest.raises(RuntimeError, match="no CPU path"):
        capi.check(rc)
