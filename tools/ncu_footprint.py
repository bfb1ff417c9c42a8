"""Hot-code footprint of a kernel from `ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:K`:
how many distinct SASS instructions carry the execution (instruction cache: ~32 KB = 2048 instructions hold the full
issue rate, tools/ubench/icache.cu), and which source lines own them."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; cur = None; per = collections.Counter(); exe = collections.Counter(); text = {}; ex = []
for r in rows:
    if r and r[0] == "Line No": hdr = r; ci = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0] != "": cur = int(r[0]); text[cur] = r[1].strip(); continue
    try: e = float(r[ci])
    except Exception: e = 0.0
    if e > 0: ex.append(e)
    if e >= 1000: per[cur] += 1; exe[cur] += e
tot = sum(ex); s = sorted(ex, reverse=True)
print("executed warp instructions %.2f M; distinct SASS instructions executed %d (%.0f KB)" % (tot / 1e6, len(ex), len(ex) * 16 / 1024))
for frac in (0.9, 0.99):
    acc = 0
    for i, e in enumerate(s):
        acc += e
        if acc >= frac * tot:
            print("  %.0f%% of the execution comes from %d instructions (%.0f KB)" % (100 * frac, i + 1, (i + 1) * 16 / 1024)); break
for ln, c in per.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print("  %4d hot instrs x %6.0fk  L%-5d %s" % (c, exe[ln] / c / 1e3, ln, text.get(ln, "")[:100]))
