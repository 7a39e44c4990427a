"""Top stalled SASS instructions of one kernel from an .ncu-rep source page."""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[col[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print({k: round(v / tot, 3) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]})
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top]:
    dom = sorted(stalls, key=lambda s: -f(r, s))[:2]
    print(f"{f(r,'# Samples')/tot*100:5.1f}%  {r[col['Source']][:70]:70s} {dom[0]}={f(r,dom[0]):.0f} {dom[1]}={f(r,dom[1]):.0f} wf={r[col['L1 Wavefronts Shared']]}/{r[col['L1 Wavefronts Shared Ideal']]}")
