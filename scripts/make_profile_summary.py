"""Turn gpurun_out/{launches.csv, prof.ncu-rep} into the tracked round summary under profiles/."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
g = os.path.join(ROOT, "gpurun_out")

# ---- launch list (gpu__time_duration per launch; cold-cache, serialised: compare shares) ----
launch_file = sys.argv[3] if len(sys.argv) > 3 else "launches.csv"
rows = [r for r in csv.reader(open(os.path.join(g, launch_file))) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = [(r[ki], float(r[vi].replace(",", "")), r[ui]) for r in rows[1:]]
with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
    f.write("kernel,gpu__time_duration.sum,unit\n")
    for k, v, u in launches:
        f.write(f"\"{k}\",{v},{u}\n")
ours = [(k, v) for k, v, _ in launches if "wg::" in k or "pack_" in k]
step = {}
for k, v in ours:
    base = k.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("(")[0].replace("void ", "").replace("wg::", "").strip()
    step.setdefault(base.split("<")[0], []).append(v)
share = {k: sum(v) / len(v) for k, v in step.items()}
tot = sum(share.values())

# ---- full-set metrics of the three hot kernels ----
rep = sys.argv[2] if len(sys.argv) > 2 else "prof.ncu-rep"
raw = subprocess.run(["ncu", "-i", os.path.join(g, rep), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h = rr[0]
def col(name):
    return h.index(name) if name in h else None
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_dynamic", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.max"]
kernels = []
for r in rr[2:]:
    d = {"kernel": r[col("Kernel Name")]}
    for k in keep:
        c = col(k)
        if c is not None:
            try:
                d[k] = float(r[c].replace(",", ""))
            except ValueError:
                d[k] = r[c]
            d[k + ".unit"] = rr[1][c]
    kernels.append(d)
summary = {"tag": tag, "command": "python scripts/profile_step.py 2 (B=4096, S=34, T=168: the bench workload, fp32 path) "
                                     "+ `... 2 4096 tensor` for the tensor-path kernels",
           "launch_share_ns": share, "launch_share_frac": {k: v / tot for k, v in share.items()}, "kernels": kernels}
json.dump(summary, open(os.path.join(out_dir, f"{tag}_ncu_summary.json"), "w"), indent=1)
with open(os.path.join(out_dir, f"{tag}_ncu_summary.md"), "w") as f:
    f.write(f"# ncu summary {tag} — bench workload (S=34, T=168, B=4096), one forward\n\n")
    f.write("Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`), mean per kernel of the step:\n\n")
    f.write("| kernel | ns per launch | share of step |\n|---|---:|---:|\n")
    for k, v in share.items():
        f.write(f"| {k} | {v:,.0f} | {v / tot:.1%} |\n")
    f.write("\n`ncu --set full` of one launch each:\n\n| kernel | ms | DRAM read GB | DRAM write GB | FMA pipe active % | tensor pipe active % | L2->SM GB | issue active % | regs | grid x block |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
    for d in kernels:
        def unit_gb(key):
            v, u = d.get(key, 0.0), d.get(key + ".unit", "")
            return v * {"Gbyte": 1, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}.get(u, 1)
        t = d.get("gpu__time_duration.sum", 0.0)
        tu = d.get("gpu__time_duration.sum.unit", "")
        t_ms = t * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(tu, 1)
        f.write(f"| {d['kernel'][:60]} | {t_ms:.3f} | {unit_gb('dram__bytes_read.sum'):.3f} | {unit_gb('dram__bytes_write.sum'):.3f} | "
                f"{d.get('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} | "
                f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} | {unit_gb('l1tex__m_xbar2l1tex_read_bytes.sum'):.2f} | "
                f"{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.1f} | "
                f"{d.get('launch__registers_per_thread', 0):.0f} | {d.get('launch__grid_size', 0):.0f} x {d.get('launch__block_size', 0):.0f} |\n")
print(open(os.path.join(out_dir, f"{tag}_ncu_summary.md")).read())
