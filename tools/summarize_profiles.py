"""Turns the ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py <round-tag> <launches.csv> <full.ncu-rep>
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches_csv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list
rows = [r for r in csv.reader(l for l in open(launches_csv) if not l.startswith("=="))]
ix = {h: i for i, h in enumerate(rows[0])}
seq, per = [], collections.OrderedDict()
for r in rows[1:]:
    name = r[ix["Kernel Name"]].split("::")[-1].split("(")[0]
    ns = float(r[ix["Metric Value"]])
    seq.append((name, ns))
    per.setdefault(name, []).append(ns)
with open(os.path.join(out, f"{tag}_launches_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:scan_tc|merge_|prep_|mask_kernel|seed_' "
            "-c 60 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-check\n"
            "(per-launch times are cold-cache and serialised under the profiler: compare SHARES, not absolutes)\n\n")
    f.write(f"{'kernel':26s} {'launches':>8s} {'total ms':>10s} {'mean ms':>10s}\n")
    for n, v in per.items():
        f.write(f"{n:26s} {len(v):8d} {sum(v) / 1e6:10.3f} {sum(v) / len(v) / 1e6:10.3f}\n")
    step_names = ("prep_queries_kernel", "seed_thresholds_kernel", "mask_kernel", "scan_tc_kernel", "scan_fp32_kernel", "merge_emit_kernel")
    last_merge = max(i for i, (n, _) in enumerate(seq) if n.startswith("merge_emit_kernel"))
    first = max(i for i, (n, _) in enumerate(seq[:last_merge]) if n.startswith("prep_queries_kernel"))
    step = [x for x in seq[first:last_merge + 1] if x[0].startswith(step_names)]
    tot = sum(ns for _, ns in step)
    f.write("\none bb_search step (last prep_queries .. merge_emit):\n")
    for n, ns in step:
        f.write(f"  {n:24s} {ns / 1e6:9.3f} ms  {100 * ns / tot:5.1f}% of step\n")
print(open(os.path.join(out, f"{tag}_launches_summary.txt")).read())

# ---- full capture of the scan kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, vals = rr[0], rr[1], rr[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
metrics = {}
for h, u, v in zip(hdr, units, vals):
    if h in want:
        metrics[h] = (v, u)
to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
dram = sum(float(metrics[k][0]) * to_bytes.get(metrics[k][1], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sr = list(csv.reader(src.splitlines()))
sh = sr[1]
six = {h: i for i, h in enumerate(sh)}
data = sr[2:]
tot = sum(int(r[six["# Samples"]] or 0) for r in data)
top = sorted(data, key=lambda r: -int(r[six["# Samples"]] or 0))[:25]
with open(os.path.join(out, f"{tag}_scan_tc_ncu_full.txt"), "w") as f:
    f.write(f"ncu --set full --clock-control none --import-source on -k regex:scan_tc -c 1 ... ({os.path.basename(rep)})\n\n")
    for k in want:
        if k in metrics:
            f.write(f"{k:90s} {metrics[k][0]} {metrics[k][1]}\n")
    f.write(f"\ndram bytes per launch (read+write): {dram:.4e}\n")
    f.write(f"\ntop SASS instructions by warp-stall samples (total {tot}):\n")
    for r in top:
        st = {k.replace("stall_", ""): int(r[six[k]] or 0) for k in sh if k.startswith("stall_") and "Not Issued" not in k}
        st = {k: v for k, v in st.items() if v > 0.05 * max(1, int(r[six["# Samples"]]))}
        f.write(f"  {int(r[six['# Samples']]):8d}  exec={r[six['Instructions Executed']]:>12s}  {r[six['Source']].strip()[:70]:70s} {st}\n")
json.dump({"kernel": "scan_tc_kernel<2,0,0>", "per_n_gpus": {"1": dram}, "source": f"profiles/{tag}_scan_tc_ncu_full.txt",
           "note": "one ncu --set full capture per GPU count; bench.py prints null for counts that were not captured"},
          open(os.path.join(out, "scan_dram_bytes.json"), "w"))
print(open(os.path.join(out, f"{tag}_scan_tc_ncu_full.txt")).read()[:3000])
