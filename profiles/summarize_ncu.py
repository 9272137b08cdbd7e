"""Turns gpurun_out/<tag>_launches.csv (+ <tag>_prof.ncu-rep) into the committed summaries:
profiles/<tag>_launches_summary.txt, profiles/<tag>_ncu_raw_summary.csv, profiles/ncu_traffic.json.

    python profiles/summarize_ncu.py r1
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
src = os.path.join(ROOT, "gpurun_out")
out = os.path.join(ROOT, "profiles")

# ---- launch list -------------------------------------------------------------------------
lines = [l for l in open(os.path.join(src, f"{tag}_launches.csv")) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    a = agg.setdefault(r["Kernel Name"], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
unit = rows[0]["Metric Unit"]
with open(os.path.join(out, f"{tag}_launches_summary.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {len(rows)} launches of "
            f"`python bench.py --steps N --warmup 3 --no-cpu-baseline` (cold-cache, serialised: compare SHARES)\n")
    f.write(f"# total {tot/1e6:.3f} ms ({unit})\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{v[1]/1e6:10.3f} ms {v[0]:5d}x {100*v[1]/tot:6.2f}%  avg {v[1]/v[0]/1e3:9.1f} us  {k}\n")
print(open(os.path.join(out, f"{tag}_launches_summary.txt")).read())

# ---- full-set capture ---------------------------------------------------------------------
rep = os.path.join(src, f"{tag}_prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_active.avg",
            "lts__t_bytes.sum", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
            "smsp__cycles_elapsed.avg.per_second"]
    idx = [i for i, h in enumerate(hdr) if h in want]
    with open(os.path.join(out, f"{tag}_ncu_raw_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for row in r[2:]:
            w.writerow([row[i] for i in idx])
    print(open(os.path.join(out, f"{tag}_ncu_raw_summary.csv")).read())
    # per-launch DRAM traffic keyed the way bench.py names the kernels
    names = {"clip_s_kernel<0>": "clip_s_kernel<FWD> (logits + exp-sums)",
             "clip_s_kernel<1>": "clip_s_kernel<DZ> (logits recompute + dL/dZ panel)",
             "clip_s_kernel<8>": "clip_s_kernel<FWD_E> (logits + exp-sums + kept exponentials)",
             "dz_from_exp_kernel": "dz_from_exp_kernel (in-place rescale of the kept panel, HBM-bound)",
             "gemm_kernel<0, 1": "gemm_kernel<K,MN> (dA = Wz . B)",
             "gemm_kernel<1, 1": "gemm_kernel<MN,MN> (dB = Wz^T . A)"}
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tpath = os.path.join(out, "ncu_traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    fresh = set()
    for row in r[2:]:
        for short, long in names.items():
            if short in row[ki] and long not in fresh:
                fresh.add(long)
                traffic[long] = float(row[ri]) * scale[units[ri]] + float(row[wi]) * scale[units[wi]]
    with open(os.path.join(out, "ncu_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print(traffic)
