"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol include/lte_b200.h declares,
and refuses to compute without a GPU (no silent CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "lte_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lteb200_\w+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib.dll, s), f"{s} declared in include/lte_b200.h but not exported by {lib.path}"
    from openlte_b200.lib import SYMBOLS
    assert sorted(SYMBOLS) == syms, "openlte_b200/lib.py and include/lte_b200.h disagree about the ABI"


def test_no_torch_or_cxx_types_in_the_abi():
    txt = open(os.path.join(ROOT, "include", "lte_b200.h")).read()
    assert 'extern "C"' in txt
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    for banned in ("torch", "at::", "std::", "cudaStream_t", "template"):
        assert banned not in txt


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.create(C.byref(h), 0) == -2          # LTEB200_ERR_CUDA
    with pytest.raises(RuntimeError):
        lib.new_ctx()


def test_missing_library_fails_loudly(tmp_path):
    from openlte_b200.lib import Lib, LibraryMissing
    with pytest.raises(LibraryMissing):
        Lib(str(tmp_path / "nope.so"))


def test_product_does_not_reference_the_oracle():
    """Nothing under openlte_b200/ or include/ may import, link or execute oracle/ (the oracle is test infrastructure)."""
    for base in ("openlte_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")) or fn == "Makefile":
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    code = re.sub(r"//.*|/\*.*?\*/|#.*|\"\"\".*?\"\"\"", "", txt, flags=re.S)
                    assert "liboracle" not in code and "libref_oracle" not in code and "from oracle" not in code and "import oracle" not in code, os.path.join(dp, fn)
    out = subprocess.run(["ldd", os.path.join(ROOT, "openlte_b200", "libopenlte_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_standalone_drop_in_header_has_the_reference_layout(tmp_path):
    """include/liblte_phy_b200.h in its self-contained mode (no reference headers on the include path) against the reference's own liblte_phy.h:
    error / sample-rate / modulation / channel-type / pre-coder enum values, sizes of the subframe, allocation, coarse-timing, PDCCH, PCFICH, PHICH
    and bit-message structs and the offsets of every field the path reads or writes -- a caller compiled against either header passes the same bytes."""
    import subprocess
    import pytest
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "liblte", "hdr")):
        pytest.skip("needs the reference headers (/root/reference)")
    src = os.path.join(ROOT, "tests", "cpp", "layout_probe.cc")
    a, b = str(tmp_path / "ref"), str(tmp_path / "b200")
    subprocess.run(["g++", "-fsigned-char", "-DUSE_REF", "-I", os.path.join(ref, "liblte", "hdr"), "-I", os.path.join(ref, "cmn_hdr"), "-I", os.path.join(ROOT, "oracle", "fftw3_shim"),
                    src, "-o", a], check=True, capture_output=True)
    subprocess.run(["g++", "-DLIBLTE_PHY_B200_STANDALONE", "-I", os.path.join(ROOT, "include"), src, "-o", b], check=True, capture_output=True)
    ra = subprocess.run([a], capture_output=True, text=True, check=True).stdout
    rb = subprocess.run([b], capture_output=True, text=True, check=True).stdout
    assert ra == rb and "SUBFRAME 1382404" in ra, (ra, rb)
