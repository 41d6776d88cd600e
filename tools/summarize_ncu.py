"""Turns ncu output into the small text/JSON summaries kept under profiles/.

  python tools/summarize_ncu.py launches <launches.csv> <out.txt> "<command that was profiled>"
      launch list from `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <cmd>`
  python tools/summarize_ncu.py full <raw.csv> <out.json> "<what was captured>"
      raw page of a `--set full` report: `ncu -i x.ncu-rep --page raw --csv > raw.csv`
"""
import csv, json, sys, collections, re

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else v)      # -> ms
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values()); n = sum(a[0] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list of `%s` (gpu__time_duration.sum, --clock-control none)\n" % cmd)
        f.write("# per-launch times are cold-cache and serialised: compare SHARES. total %.1f ms over %d launches\n\n" % (tot, n))
        for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-80s launches=%4d total_ms=%10.3f share=%5.1f%% avg_ms=%8.4f\n" % (name[:80], c, t, 100 * t / tot, t / c))


def full(src, dst, what):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    out = {"capture": what, "kernels": []}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        k = {"name": d["Kernel Name"]}
        for m in KEEP:
            if m in d:
                k[m] = "%s %s" % (d[m], units[hdr.index(m)])
        st = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(d[h]) for h in hdr
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")}
        k["stall_cycles_per_issue (top)"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:7])
        out["kernels"].append(k)
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
