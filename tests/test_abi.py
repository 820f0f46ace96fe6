"""CPU tests: the C-ABI library builds, loads and exports every symbol include/zquatev_b200.h
declares; without a GPU every compute entry fails loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from zquatev_b200.build import build
    return build()


def header_functions():
    txt = open(os.path.join(ROOT, "include", "zquatev_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b(zquatev_b200\w*|zq_test_\w+)\s*\(", txt)
    return sorted(set(names))


def test_header_symbols_exported(built):
    L = ctypes.CDLL(built)
    names = header_functions()
    assert "zquatev_b200" in names and len(names) >= 10
    for nme in names:
        assert hasattr(L, nme), f"{nme} declared in include/zquatev_b200.h but not exported"


def test_python_binding_covers_header(built):
    from zquatev_b200 import api
    assert sorted(api.SYMBOLS) == header_functions()
    assert "sm_100a" in api.version()


def test_cxx_symbol_of_reference_exported(built):
    """ts::zquatev(int, std::complex<double>*, int, double*) -- the mangled name the reference's
    test.cc links against (reference zquatev.h:54)."""
    out = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True).stdout
    assert "_ZN2ts7zquatevEiPSt7complexIdEiPd" in out


def test_argument_checks(built):
    from zquatev_b200 import api
    L = api.lib()
    D = np.zeros((4, 4), dtype=np.complex128, order="F")
    e = np.zeros(2)
    assert L.zquatev_b200(3, D.ctypes.data, 4, e.ctypes.data) == -1
    assert L.zquatev_b200(4, None, 4, e.ctypes.data) == -2
    assert L.zquatev_b200(4, D.ctypes.data, 2, e.ctypes.data) == -3
    assert L.zquatev_b200(4, D.ctypes.data, 4, None) == -4
    assert L.zquatev_b200(0, D.ctypes.data, 4, e.ctypes.data) == 0


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import zquatev_b200 as z
    D = np.zeros((4, 4), dtype=np.complex128, order="F")
    e = np.zeros(2)
    with pytest.raises(RuntimeError):
        z.zquatev(4, D, 4, e)


def test_product_never_imports_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "zquatev_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace(
                    "oracle/.", ""), f


def test_headers_compile_standalone(tmp_path):
    """include/zquatev_b200.h is plain C (no torch / CUDA types in the signatures); include/zquatev.h is the C++
    declaration of the reference symbol -- both must compile on their own with the host compilers."""
    inc = os.path.join(ROOT, "include")
    c = tmp_path / "t.c"
    c.write_text('#include "zquatev_b200.h"\nint main(void) { zq_options o = {0}; (void)o; return zquatev_b200_version() == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(c)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cc = tmp_path / "t.cc"
    cc.write_text('#include "zquatev.h"\n#include "zquatev_b200.h"\n'
                  'int f(int n2, std::complex<double>* D, double* e) { return ts::zquatev(n2, D, n2, e); }\n')
    r = subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(cc)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
