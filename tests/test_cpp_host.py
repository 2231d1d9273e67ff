"""The C++ drop-in classes (include/tchem/, host/libtchem.so): built in-tree by __graft_entry__.build(),
exercised here by host/test_intcoord.exe, which mirrors the reference's own test programs."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden

EXE = os.path.join(ROOT, "host", "test_intcoord.exe")


def test_host_library_is_built_and_links_the_c_abi():
    lib = os.path.join(ROOT, "host", "libtchem.so")
    if not os.path.exists(lib):
        pytest.skip("host/libtchem.so not built (run __graft_entry__.build())")
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    for sym in ("IntCoordSet", "compute_IC_J_K", "Hessian_int2cart", "Hessian_cart2int", "gradient_cart2int_matrix",
                "SAPSet", "Jacobian_"):
        assert sym in out, sym
    undefined = subprocess.run(["nm", "-D", "--undefined-only", lib], capture_output=True, text=True).stdout
    assert "tchem_ic_eval" in undefined and "tchem_sap_eval" in undefined     # computes only through the C ABI


@pytest.mark.gpu
def test_cpp_drop_in_classes(tmp_path):
    assert os.path.exists(EXE), "host/test_intcoord.exe not built"
    d, c, a = golden("ref_aniline_default"), golden("ref_aniline_columbus7"), golden("ref_aniline_advanced")
    (tmp_path / "IntCoordDef").write_text(str(d["definition"]))
    (tmp_path / "intcfl").write_text(str(c["definition"]))
    (tmp_path / "advanced_IntCoordDef").write_text(str(a["definition"]))
    np.savetxt(tmp_path / "r_default.txt", d["r"][0], fmt="%.17e")     # slow-1.5.xyz in bohr
    np.savetxt(tmp_path / "r_columbus.txt", c["r"][0], fmt="%.17e")    # Columbus geom
    (tmp_path / "sap1.in").write_text(str(golden("ref_sap_1")["definition"]))
    (tmp_path / "sap2.in").write_text(str(golden("ref_sap_2")["definition"]))
    p = subprocess.run([EXE, str(tmp_path)], capture_output=True, text=True, timeout=300)
    print(p.stdout)
    print(p.stderr[-2000:])
    assert p.returncode == 0, p.stdout[-3000:]
    assert "FAIL" not in p.stdout
