#!/usr/bin/env python3
"""Static report of the shipped cubins: per-kernel registers / stack / static shared memory (cuobjdump -res-usage) and, for the hot
kernels, counts of the SASS mnemonics that show which hardware paths are used (TMA bulk copy UBLKCP, cp.async LDGSTS, packed 16-bit
DPX VIADDMNMX / VIADD.16x2, shared-memory atomics ATOMS, byte permutes PRMT).  Needs no GPU.

usage: python tools/static_report.py [path/to/libopenlte_b200.so] > profiles/rN_static_report.txt
"""
import collections, re, subprocess, sys

so = sys.argv[1] if len(sys.argv) > 1 else "openlte_b200/libopenlte_b200.so"


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(n):
    n = n.replace("(anonymous namespace)::", "")
    n = re.sub(r"^void ", "", n)
    return re.sub(r"\(.*$", "", n)


res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout.splitlines()
rows, fn = [], None
for ln in res:
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        fn = m.group(1)
        continue
    if fn and "REG:" in ln:
        kv = dict(re.findall(r"(\w+(?:\[\d\])?):(\d+)", ln))
        rows.append((fn, int(kv["REG"]), int(kv["STACK"]), int(kv["SHARED"]), int(kv["LOCAL"])))
        fn = None
dm = demangle([r[0] for r in rows])
print(f"# static report of {so} (sm_100a cubins)")
print("# kernel | registers/thread | stack bytes (per-thread arrays / spills) | static shared bytes | local bytes")
for fn, reg, stack, sh, loc in sorted(rows, key=lambda r: short(dm[r[0]])):
    print(f"{short(dm[fn]):58s} {reg:4d} {stack:5d} {sh:6d} {loc:4d}")

sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines()
WATCH = ["UBLKCP", "LDGSTS", "VIADDMNMX", "VIADD", "VIMNMX", "PRMT", "ATOMS", "SHFL", "BAR", "MUFU", "FFMA", "LDS", "STS", "LDG", "STG", "SYNCS"]
cur, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for ln in sass:
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Za-z0-9_.]+)", ln)
    if cur and m:
        op = m.group(1)
        total[cur] += 1
        base = op.split(".")[0]
        if base in WATCH:
            counts[cur][base + (".16x2" if "16x2" in op else "")] += 1
dm2 = demangle(list(total))
print("\n# SASS mnemonic counts (static) of the hot kernels")
HOT = ("turbo_mlm2_kernel<384", "pdsch_llr_q8_kernel", "ofdm_demod_r16_kernel<2048, char2, false, false>", "ofdm_demod_r16_kernel<2048, char2, false, true>", "chan_est_kernel", "turbo_ref_siso", "turbo_mlm_kernel<16, 384, true>")
for fn in sorted(total, key=lambda f: short(dm2[f])):
    name = dm2[fn].replace("(anonymous namespace)::", "").replace("void ", "")
    if any(name.startswith(h) for h in HOT):
        c = counts[fn]
        print(f"{re.sub(r'[(].*$', '', name)}: {total[fn]} instructions; " + ", ".join(f"{k} {v}" for k, v in sorted(c.items())))
