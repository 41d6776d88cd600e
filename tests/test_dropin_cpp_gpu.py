"""The C++ drop-in surface (include/liblte_phy_b200.h): a scan harness written like the reference's caller is compiled
with g++ against the header, linked to libopenlte_b200.so and run on the reference's capture (BASELINE config 1)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "scan_harness.cc")


def build(tmp_path):
    exe = str(tmp_path / "scan_harness")
    cmd = ["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", os.path.join(ROOT, "openlte_b200"),
           "-lopenlte_b200", "-Wl,-rpath," + os.path.join(ROOT, "openlte_b200")]
    subprocess.check_call(cmd)
    return exe


def test_harness_compiles_and_links_against_the_dropin_header(tmp_path):
    """CPU part of the boundary check: the header is self-contained C++ and every liblte_phy_* symbol it declares resolves."""
    build(tmp_path)


@pytest.mark.gpu
def test_scan_on_reference_capture(tmp_path):
    exe = build(tmp_path)
    z = np.load(os.path.join(ROOT, "tests", "golden", "capture_cfg1.npz"))
    iq = z["iq"]
    binf = tmp_path / "cap.bin"
    iq.astype(np.int8).tofile(binf)
    out = subprocess.run([exe, str(binf), str(iq.shape[0]), "20", "5", hex(int(z["rnti"])), str(int(z["tbs"])), str(len(z["prb"])), str(int(z["mod"])),
                          str(int(z["rv"])), str(int(z["cfi"]))], capture_output=True, text=True, timeout=300).stdout
    assert "coarse n_corr_peaks=5 symb_start0=2192" in out, out
    assert "pss N_id_2=0 pss_symb=5 freq_offset=0" in out, out
    assert "sss N_id_1=0 frame_start_idx=307185 N_id_cell=0" in out, out
    mib = "".join(map(str, z["mib_bits"]))
    assert f"bch err=0 N_ant={int(z['mib_N_ant'])} offset={int(z['mib_offset'])} bits={mib}" in out, out      # MIB 1010100..., BASELINE.md P2
    bits = "".join(map(str, z["sib1_bits"]))
    assert f"pdsch err=0 N_bits=144 bits={bits}" in out, out
    # the allocation found by the PDCCH decode is the one the reference finds, and decoding with it gives the same SIB1
    assert f"pdcch err=0 cfi={int(z['cfi'])} N_symbs={int(z['cfi'])}" in out and f"rnti=0xFFFF tbs=144 N_prb=4 prb0=0 mod=1 rv={int(z['rv'])}" in out, out
    assert f"pdsch_from_dci err=0 N_bits=144 bits={bits}" in out, out
    assert "null_check=1" in out            # LIBLTE_ERROR_INVALID_INPUTS
