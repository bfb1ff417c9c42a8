"""Per-CUDA-source-line instruction / stall-sample totals from
`ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:K > f.csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] != "":  # a source-line summary row
        lines.append(r)
ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples")
def f(x):
    try: return float(x)
    except Exception: return 0.0
ti = sum(f(r[ci]) for r in lines); ts = sum(f(r[cs]) for r in lines)
print("total warp-instr %d  samples %d" % (ti, ts))
for r in sorted(lines, key=lambda r: -f(r[ci]))[:n]:
    print("%5.1f%% instr %5.1f%% samp  L%-4s %s" % (100 * f(r[ci]) / ti, 100 * f(r[cs]) / max(ts, 1), r[0], r[1].strip()[:105]))

if len(sys.argv) > 3:  # region totals: "name:lo-hi,name:lo-hi"
    print("regions:")
    for spec in sys.argv[3].split(","):
        name, rng = spec.split(":")
        lo, hi = [int(x) for x in rng.split("-")]
        sel = [r for r in lines if lo <= int(r[0]) <= hi]
        print("  %-14s %5.1f%% instr %5.1f%% samples" % (name, 100 * sum(f(r[ci]) for r in sel) / ti, 100 * sum(f(r[cs]) for r in sel) / max(ts, 1)))
