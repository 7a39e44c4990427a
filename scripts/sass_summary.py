"""profiles/<tag>_sass_summary.txt: per kernel of the built library, the SASS mnemonics that show what it runs on
(FFMA2 packed FP32, UTC*MMA tcgen05, LDTM/STTM tensor memory, UBLKCP / UTMA* bulk copies, HMMA legacy tensor path).

    python scripts/sass_summary.py r02
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "windgnn_b200", "lib", "libwindgnn_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout  # noqa: F841 (kept for arch line)
WATCH = ["FFMA2", "FFMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS",
         "HMMA", "MUFU", "SYNCS", "BAR"]
cur, counts, order = None, {}, []
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_total"] += 1
        if op in WATCH:
            counts[cur][op] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
out = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
arch = re.search(r"arch = (sm_\w+)", sass)
with open(out, "w") as f:
    f.write(f"# SASS mnemonic counts per kernel of windgnn_b200/lib/libwindgnn_b200.so ({arch.group(1) if arch else '?'})\n")
    f.write("# cuobjdump -sass | count per function; UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk,\n")
    f.write("# FFMA2 = packed FP32 FMA (fma.rn.f32x2), HMMA = legacy mma.sync path (none expected)\n")
    tot = collections.Counter()
    for raw, nice in zip(order, demangle):
        c = counts[raw]
        tot.update({k: v for k, v in c.items() if k != "_total"})
        short = re.sub(r"\(.*", "", nice).replace("wg::", "").replace("(anonymous namespace)::", "")
        cols = " ".join(f"{k}={c[k]}" for k in WATCH if c[k])
        f.write(f"{short:70s} instrs={c['_total']:6d}  {cols}\n")
    f.write("\nTOTAL " + " ".join(f"{k}={tot[k]}" for k in WATCH) + "\n")
print(open(out).read()[-1200:])
