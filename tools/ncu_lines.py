"""Per-CUDA-source-line view of an ncu report captured with --import-source on: warp-state samples per source line (the
aggregated rows of `--print-source sass,cuda`), the top lines with their main stall reasons, and 25-line buckets of one file.
usage: python tools/ncu_lines.py report.ncu-rep [file-substring] [top N]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else "dsx_stack.cu"
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    cur, hdr, lines = "", None, []
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r[0].isdigit():
            continue
        ix = {h: i for i, h in enumerate(hdr)}
        sc = hdr.index("Warp Stall Sampling (All Samples)")
        try:
            s = int(r[sc])
        except ValueError:
            continue
        stalls = {h[6:]: int(r[i]) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit() and int(r[i])}
        lines.append((cur, int(r[0]), r[1], s, stalls))
    tot = sum(l[3] for l in lines)
    print("total samples", tot)
    byfile = {}
    for f, n, src, s, st in lines:
        byfile[f] = byfile.get(f, 0) + s
    for f, s in sorted(byfile.items(), key=lambda kv: -kv[1]):
        print(f"  {s:7d} {100.0 * s / max(tot, 1):5.1f}%  {f}")
    sel = [l for l in lines if want in l[0]]
    print(f"--- top {top} lines of {want}")
    for f, n, src, s, st in sorted(sel, key=lambda l: -l[3])[:top]:
        why = " ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{s:7d} {100.0 * s / max(tot, 1):5.1f}%  {n:>5}  {src.strip()[:110]}   [{why}]")
    print("--- samples per 25-line bucket")
    b = {}
    for f, n, src, s, st in sel:
        b[n // 25 * 25] = b.get(n // 25 * 25, 0) + s
    for k in sorted(b):
        if b[k]:
            print(f"lines {k:5d}-{k + 24:5d}: {b[k]:7d} {100.0 * b[k] / max(tot, 1):5.1f}%")


if __name__ == "__main__":
    main()
