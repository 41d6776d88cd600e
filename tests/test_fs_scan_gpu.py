"""The reference's OWN file-scan block on the drop-in surface.

oracle/Makefile (target `fs`) compiles /root/reference/LTE_fdd_dl_file_scan/src/LTE_fdd_dl_fs_samp_buf.cc + liblte_rrc.cc + liblte_common.cc where they
lie -- GNU Radio replaced by the inert stand-ins of oracle/gr_stub/ -- twice: against the reference's liblte_phy.cc (fs_scan_ref, CPU) and against
include/dropin/liblte_phy.h + libopenlte_b200.so (fs_scan_b200).  The second build is itself the proof that the drop-in header shares a translation
unit with liblte_rrc.h / liblte_common.h / typedefs.h (LTE_fdd_dl_fs_samp_buf.h:55-56).  Here both run the block's work() over the reference's shipped
capture (Test captures/LTE_test_file_int8.bin, 30 frames at 30.72 MHz) exactly as the flow graph would, and must print the same cell: MIB, SIB1, SIB2
(BASELINE.md P2).  Because the caller is the reference's code, this also pins freq_shift (fs.cc:696-713), the int8 de-interleave (fs.cc:657-694) and
the window carry-over (fs.cc:601-611) to the reference itself: they run unmodified on both sides, and every liblte_phy_* call sees their output.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "fs_scan_ref")
B200_BIN = os.path.join(ROOT, "oracle", "_ref", "fs_scan_b200")
CAPTURE = os.path.join(ROOT, "oracle", "_ref", "LTE_test_file_int8.bin")
GOLDEN = os.path.join(ROOT, "tests", "golden", "fs_scan_capture_ref_output.txt")

have = all(os.path.exists(p) for p in (REF_BIN, B200_BIN, CAPTURE))


def run(binary, chunk=32768):
    r = subprocess.run([binary, CAPTURE, str(chunk)], input="\n", capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def normalise(text):
    # the float "Frequency Offset" line is printed with two decimals; everything else is integers, enums and strings
    return [ln.rstrip() for ln in text.splitlines()]


@pytest.mark.skipif(not have, reason="oracle/_ref file-scan harness not built (needs /root/reference at build time)")
def test_reference_scan_block_reproduces_its_golden_output_on_cpu():
    """fs_scan_ref (the reference end to end, CPU only) against the output recorded when the fixture was made: the harness itself is sound."""
    assert normalise(run(REF_BIN)) == normalise(open(GOLDEN).read())


@pytest.mark.gpu
@pytest.mark.skipif(not have, reason="oracle/_ref file-scan harness not built (needs /root/reference at build time)")
def test_reference_scan_block_on_the_dropin_surface():
    want = normalise(open(GOLDEN).read())
    got = normalise(run(B200_BIN))
    assert any("MIB Decoded" in ln for ln in got) and any("SIB1 Decoded" in ln for ln in got), "\n".join(got[:40])
    assert got == want, "the reference's scan block prints a different cell on the drop-in surface:\n" + "\n".join(
        f"{a!r} != {b!r}" for a, b in zip(got, want) if a != b)[:3000]


@pytest.mark.gpu
@pytest.mark.skipif(not have, reason="oracle/_ref file-scan harness not built (needs /root/reference at build time)")
def test_scan_is_independent_of_the_scheduler_chunking():
    """GNU Radio hands work() whatever chunk sizes it likes (odd sizes exercise the I/Q carry of fs.cc:672-690)."""
    assert normalise(run(B200_BIN, chunk=4095)) == normalise(open(GOLDEN).read())
