"""Summarise an `ncu --page source --csv` dump: stall reasons and the hottest SASS instructions.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:k_raster > src.csv; python tools/ncu_top.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def num(r, h):
    try: return float(r[col[h]])
    except Exception: return 0.0
total = sum(num(r, "# Samples") for r in data)
print("kernel:", rows[0][1], " samples:", int(total), " sass instrs:", len(data),
      " warp-instr executed:", int(sum(num(r, "Instructions Executed") for r in data)))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = sorted(((sum(num(r, h) for r in data), h) for h in stalls), reverse=True)
print("stalls:", ", ".join("%s %.1f%%" % (h[6:], 100 * v / max(total, 1)) for v, h in agg[:8]))
top = sorted(data, key=lambda r: -num(r, "# Samples"))[:n]
for r in top:
    st = sorted(((num(r, h), h[6:]) for h in stalls), reverse=True)[:2]
    print("%5.1f%%  exec %8d  %-70s %s" % (100 * num(r, "# Samples") / max(total, 1), num(r, "Instructions Executed"),
          r[col["Source"]].strip()[:70], " ".join("%s:%d" % (h, v) for v, h in st if v)))
