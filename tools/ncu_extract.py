"""Extracts the judged subset of an ncu report into a small text file (profiles/*.txt) and, optionally, the per-line stall
summary of the source page.  usage: python tools/ncu_extract.py report.ncu-rep out.txt [--stalls]"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r"^(Kernel Name|gpu__time_duration|sm__cycles_active|.*pipe_tensor.*pct_of_peak_sustained_(elapsed|active)|"
                  r"sm__inst_executed_pipe_tensor|dram__bytes_(read|write)\.sum|dram__cycles_active|gpu__dram_throughput|"
                  r"l1tex__m_xbar2l1tex_read_bytes\.sum|lts__t_sector_hit_rate\.pct|lts__t_bytes\.sum|"
                  r"launch__(cluster_size|grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_\w+|cluster_max_active)|"
                  r"smsp__warp_issue_stalled.*_per_warp_active\.pct|smsp__issue_active\.avg\.pct|sm__throughput\.avg\.pct|"
                  r"smsp__inst_executed\.sum|sm__warps_active\.avg\.pct|smsp__cycles_active\.avg|"
                  r"smsp__average_warp.*_per_issue_active|smsp__pcsamp_warps_issue_stalled)")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units, vals = rows[0], rows[1], rows[2]
    lines = []
    for n, u, v in zip(names, units, vals):
        short = n.split(".TriageCompute.")[-1] if ".TriageCompute." in n else n
        if KEEP.match(short) and v != "":
            lines.append(f"{n} [{u}] = {v}")
    return lines


def stalls(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    sidx = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not" not in h]
    tot, agg, ins = 0, {}, []
    for r in rows[2:]:
        if len(r) < 20:
            continue
        s = int(r[ix["# Samples"]] or 0)
        tot += s
        for h, i in sidx:
            agg[h] = agg.get(h, 0) + int(r[i] or 0)
        ins.append((s, r[1].strip()))
    lines = [f"# warp-state samples: {tot}; SASS instructions: {len(ins)}", "# stall reasons (all samples):"]
    for h, v in sorted(agg.items(), key=lambda x: -x[1])[:12]:
        lines.append(f"  {h:28s} {v:8d}  {100.0 * v / max(tot, 1):5.1f} %")
    lines.append(f"# top {top} instructions by samples:")
    for s, t in sorted(ins, key=lambda x: -x[0])[:top]:
        lines.append(f"  {s:6d}  {t[:110]}")
    return lines


if __name__ == "__main__":
    rep, dst = sys.argv[1], sys.argv[2]
    lines = raw(rep)
    if "--stalls" in sys.argv:
        lines += stalls(rep)
    open(dst, "w").write("\n".join(lines) + "\n")
    print(dst, len(lines), "lines")
