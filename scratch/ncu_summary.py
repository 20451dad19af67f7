"""profiles/<tag>_ncu_metrics.json + .md from `ncu --set full` reports: the metrics the roofline discussion uses."""
import csv, json, subprocess, sys
KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct_of_peak",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_smem_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_bytes",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
}
def unit_scale(name, unit, val):
    u = unit.lower()
    if name == "gpu__time_duration.sum":
        return val * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
    if "bytes" in name:
        return val * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return val
out = {}
for tag, rep in (a.split("=") for a in sys.argv[2:]):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units, v = rows[0], rows[1], rows[2]
    d = {"kernel": v[h.index("Kernel Name")]}
    for i, k in enumerate(h):
        if k in KEYS:
            try:
                d[KEYS[k]] = unit_scale(k, units[i], float(v[i].replace(",", "")))
            except ValueError:
                pass
    if "dram_read_bytes" in d:
        d["dram_bytes"] = d["dram_read_bytes"] + d.get("dram_write_bytes", 0.0)
        d["dram_gbs"] = d["dram_bytes"] / (d["duration_us"] * 1e-6) / 1e9
    out[tag] = d
json.dump(out, open(sys.argv[1] + ".json", "w"), indent=1)
with open(sys.argv[1] + ".md", "w") as f:
    f.write("# ncu --set full captures (cold cache, clocks not locked; `--clock-control none`)\n\n")
    f.write("| capture | kernel | us | DRAM MB (r+w) | DRAM GB/s | DRAM % | L2 % | L1/smem % | tensor pipe % | issue % | regs | grid x block | dyn smem |\n|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|\n")
    for tag, d in out.items():
        f.write(f"| {tag} | `{d['kernel'][:60]}` | {d.get('duration_us',0):.1f} | {d.get('dram_bytes',0)/1e6:.1f} | {d.get('dram_gbs',0):.0f} | "
                f"{d.get('dram_pct_of_peak',0):.1f} | {d.get('l2_pct_of_peak',0):.1f} | {d.get('l1_smem_pct_of_peak',0):.1f} | "
                f"{d.get('tensor_pipe_pct_active',0):.1f} | {d.get('issue_active_pct',0):.1f} | {d.get('registers_per_thread',0):.0f} | "
                f"{d.get('grid',0):.0f} x {d.get('block',0):.0f} | {d.get('dyn_smem_bytes',0):.0f} |\n")
print(open(sys.argv[1] + ".md").read())
