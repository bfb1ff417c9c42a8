"""Condenses an .ncu-rep (`ncu --set full`) into a small JSON per kernel for profiles/.
usage: python tools/ncu_summarize.py gpurun_out/prof.ncu-rep profiles/r01_vN_full.json"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = {
    "gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "smsp__inst_executed.sum": "warp_instructions", "sm__inst_executed_pipe_fp64.sum": "fp64_warp_instructions",
    "smsp__inst_executed_pipe_fma.sum": "fma_pipe_warp_instructions",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct_of_peak",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fp32_pipe_cycles_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_warp_instruction",
    "sm__cycles_elapsed.max": "elapsed_cycles", "sm__cycles_active.avg": "sm_active_cycles",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_rate_pct",
}
res = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("::")[-1].split("(")[0].split("<")[0]
    d = {}
    for i, h in enumerate(hdr):
        if h in want:
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            d[want[h]] = {"value": v, "unit": units[i]}
    def val(k, mult=None):
        return d[k]["value"] * (mult or 1) if k in d else None
    um = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    if "dram_read" in d and "dram_write" in d:
        d["dram_bytes_per_launch"] = d["dram_read"]["value"] * um.get(d["dram_read"]["unit"], 1) + d["dram_write"]["value"] * um.get(d["dram_write"]["unit"], 1)
    res.setdefault(name, []).append(d)
summary = {k: v[-1] for k, v in res.items()}
json.dump({"source": rep, "kernels": summary}, open(out, "w"), indent=1)
for k, v in summary.items():
    print(k, {a: (round(b["value"], 2) if isinstance(b, dict) else b) for a, b in v.items()})
