"""Regenerate profiles/<tag>_* from the raw artefacts a gpurun call left in gpurun_out/ (launch list CSV, ncu reports, bench JSON).
usage: python scratch/make_profiles.py r01g"""
import csv, json, os, re, shutil, subprocess, sys
tag = sys.argv[1]
G, P = "gpurun_out", "profiles"
MINE = ("pn::", "umma::", "fa::", "pairmm::", "skinny")

def launches():
    path = f"{G}/launches_{tag}.csv"
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg, tot = {}, 0.0
    for r in rows:
        t = float(r["Metric Value"].replace(",", "")) / 1e3
        name = re.sub(r"\(.*", "", r["Kernel Name"])[:100]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
    mine = sum(t for k, (n, t) in agg.items() if any(m in k for m in MINE))
    shutil.copy(path, f"{P}/{tag}_launches_one_forward.csv")
    out = [f"# {tag}: ncu launch list of ONE detector forward (bs=2, 800x1333, fp32)", "",
           "Command (B200 via gpurun): `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv "
           f"--log-file gpurun_out/launches_{tag}.csv python bench.py --profile` (warm-up, then exactly one eager forward between "
           f"cudaProfilerStart/Stop).  Raw CSV: `{tag}_launches_one_forward.csv`.  ncu times are cold-cache and serialised: compare "
           "SHARES, not absolutes (the graph-replayed step is what `bench.py` reports).", "",
           f"{len(rows)} launches, {tot/1e3:.2f} ms summed; hand-written kernels (`pn::*`, `pn::umma::*`, `pn::fa::*`, `pn::pairmm::*`): "
           f"{mine/1e3:.2f} ms = {100*mine/tot:.1f} %; the rest is PyTorch/cuDNN upstream plumbing (ResNet-50 convs, FPN 3x3 conv, "
           "max-pool, residual adds).", "", "| total us | share | launches | avg us | kernel |", "|---:|---:|---:|---:|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| {t:.1f} | {100*t/tot:.1f} % | {n} | {t/n:.1f} | `{k.strip()}` |")
    open(f"{P}/{tag}_launches_summary.md", "w").write("\n".join(out) + "\n")
    print("launches:", len(rows), f"{tot/1e3:.2f} ms, mine {100*mine/tot:.1f}%")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]

def ncu(rep):
    path = f"{G}/{rep}_{tag}.ncu-rep"
    if not os.path.exists(path):
        return None
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"Kernel Name": r[hdr.index("Kernel Name")]}
        for w in WANT:
            if w in hdr:
                d[w] = (r[hdr.index(w)], units[hdr.index(w)])
        res.append(d)
    return res

if __name__ == "__main__":
    launches()
    out = {}
    for rep in ("umma_gemm", "fa_umma", "pair_umma", "topk"):
        r = ncu(rep)
        if r:
            out[rep] = r
    json.dump(out, open(f"{P}/{tag}_ncu_metrics.json", "w"), indent=1)
    for f, dst in ((f"{G}/bench_b.json", f"{P}/{tag}_bench.json"), (f"{G}/bench_ref_b.json", f"{P}/{tag}_bench_reference_arm.json")):
        if os.path.exists(f):
            shutil.copy(f, dst)
    print(json.dumps({k: [{m: v for m, v in d.items()} for d in v] for k, v in out.items()}, indent=0)[:3000])
