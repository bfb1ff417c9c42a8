"""Writes profiles/ncu_summary.json (what bench.py reports as roofline.traffic) from a `ncu --set full` summary made by
tools/ncu_summarize.py and, optionally, a no-flush per-kernel DRAM capture (tools/gpu_scripts/l2_residency.sh).
The file records the SHA-256 of minirender_b200/csrc/mr_kernels.cu it was measured on: bench.py ignores it otherwise.
usage: python tools/make_ncu_summary.py profiles/r02_final_ncu_full.json [profiles/r02_final_frame_dram_no_flush.csv]"""
import csv, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
full = json.load(open(sys.argv[1]))
out = {"kernels_sha256": hashlib.sha256(open(os.path.join(ROOT, "minirender_b200", "csrc", "mr_kernels.cu"), "rb").read()).hexdigest()}
for k, v in full["kernels"].items():
    out[k] = {"dram_bytes_per_launch": v.get("dram_bytes_per_launch"), "duration_us": v["duration"]["value"],
              "warp_instructions": v.get("warp_instructions", {}).get("value"), "issue_slots_busy_pct": v.get("issue_slots_busy_pct", {}).get("value"),
              "source": "%s (ncu --set full, caches flushed before the kernel)" % os.path.relpath(sys.argv[1], ROOT)}
if len(sys.argv) > 2:
    acc = {}
    for r in csv.reader(open(sys.argv[2])):
        if len(r) > 14 and r[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            name = "k_geom" if "k_geom" in r[4] else "k_raster" if "k_raster" in r[4] else None
            if name:
                acc.setdefault(name, {}).setdefault(r[0], 0.0)
                acc[name][r[0]] += float(r[-1].replace(",", ""))
    for name, d in acc.items():
        vals = sorted(d.values())
        out.setdefault(name, {})["dram_bytes_per_launch_no_flush"] = vals[len(vals) // 2]
    if acc:
        out["frame"] = {"dram_bytes_per_frame_no_flush": sum(out[n]["dram_bytes_per_launch_no_flush"] for n in acc),
                        "source": "%s (ncu --cache-control none --replay-mode application: consecutive frames, L2 left as the pipeline leaves it)" % os.path.relpath(sys.argv[2], ROOT)}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
