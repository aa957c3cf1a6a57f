"""Per-CUDA-source-line executed instructions and stall samples of one kernel in an .ncu-rep
(ncu --import-source on, compiled with -lineinfo).   python tools/ncu_lines.py rep.ncu-rep [top N] [launches to skip]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
skip = ["--launch-skip", sys.argv[3], "--launch-count", "1"] if len(sys.argv) > 3 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + skip, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, fname, agg = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None or not r[0].isdigit():
        continue
    extra = len(r) - len(hdr)  # commas inside the source text split it into extra fields
    g = lambda name: r[hdr.index(name) + extra]
    try:
        inst, smp = int(g("Instructions Executed")), int(g("# Samples"))
    except ValueError:
        continue
    stalls = {k[6:]: int(g(k)) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and g(k).isdigit() and int(g(k)) > 0}
    agg.append((inst, smp, fname, int(r[0]), ",".join(r[1:2 + extra]).strip()[:100], stalls))
tot, ts = sum(a[0] for a in agg), sum(a[1] for a in agg)
print(f"total warp instructions {tot}, samples {ts}")
print("-- by executed instructions")
for a in sorted(agg, key=lambda a: -a[0])[:top]:
    print(f"{100 * a[0] / tot:5.1f}% inst {100 * a[1] / max(ts, 1):5.1f}% smp  {a[2]}:{a[3]:<4d} {a[4]}")
print("-- by stall samples")
for a in sorted(agg, key=lambda a: -a[1])[:top // 2]:
    st = ", ".join(f"{k}={v}" for k, v in sorted(a[5].items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * a[1] / max(ts, 1):5.1f}% smp {100 * a[0] / tot:5.1f}% inst  {a[2]}:{a[3]:<4d} {a[4][:70]}   [{st}]")
