"""Markdown / JSON table of the kernels in an .ncu-rep (one row per profiled launch).

    python scripts/ncu_table.py <report.ncu-rep> <out_prefix> "<title>" [kernel-regex]
"""
import csv, io, json, re, subprocess, sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
rx = re.compile(sys.argv[4]) if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, units = rr[0], rr[1]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
scale_b = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}
scale_t = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}
kernels = []
for r in rr[2:]:
    name = r[h.index("Kernel Name")]
    if rx and not rx.search(name):
        continue
    d = {"kernel": name}
    for k in keep:
        if k in h:
            c = h.index(k)
            try:
                d[k] = float(r[c].replace(",", ""))
            except ValueError:
                d[k] = r[c]
            d[k + ".unit"] = units[c]
    kernels.append(d)
json.dump({"title": title, "report": rep, "kernels": kernels}, open(out + ".json", "w"), indent=1)
with open(out + ".md", "w") as f:
    f.write(f"# {title}\n\n`ncu --set full --clock-control none`, one launch per row:\n\n")
    f.write("| kernel | ms | DRAM read GB | DRAM write GB | FMA pipe active % | issue active % | warps active % | regs | grid x block | smem KB | shared bank-conflict share |\n")
    f.write("|---|---:|---:|---:|---:|---:|---:|---:|---|---:|---:|\n")
    for d in kernels:
        g = lambda k: d.get(k, 0.0)
        t = g("gpu__time_duration.sum") * scale_t.get(d.get("gpu__time_duration.sum.unit", "ms"), 1.0)
        rd = g("dram__bytes_read.sum") * scale_b.get(d.get("dram__bytes_read.sum.unit", "byte"), 1e-9)
        wr = g("dram__bytes_write.sum") * scale_b.get(d.get("dram__bytes_write.sum.unit", "byte"), 1e-9)
        wf = g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        bc = g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / wf if wf else 0.0
        su = d.get("launch__shared_mem_per_block_dynamic.unit", "byte")
        smem = g("launch__shared_mem_per_block_dynamic") * (1.0 if "Kbyte" in su else 1024.0 if "Mbyte" in su else 1 / 1024)
        f.write(f"| {d['kernel'][:70]} | {t:.3f} | {rd:.3f} | {wr:.3f} | "
                f"{g('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
                f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {g('launch__registers_per_thread'):.0f} | "
                f"{g('launch__grid_size'):.0f} x {g('launch__block_size'):.0f} | {smem:.0f} | {bc:.0%} |\n")
print(open(out + ".md").read())
