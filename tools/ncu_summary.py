"""Print the key metrics of an .ncu-rep (raw page) and the dynamic SASS opcode mix (source page)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
sel = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if sel and sel.replace("(int)", "") not in name.replace("(int)", ""):
        continue
    print("==", name[:90])
    for k in KEYS:
        if k in hdr:
            print(f"  {k:88s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    st = sorted(((float(r[hdr.index(k)]), k) for k in STALL), reverse=True)
    print("  stalls/issue:", ", ".join(f"{k.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, k in st[:9]))
if sel:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    mix, smp, tot, cur = collections.Counter(), collections.Counter(), 0, False
    h = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = sel.replace("(int)", "") in r[1].replace("(int)", "")
            continue
        if r and r[0] == "Address":
            h = r
            continue
        if not cur or h is None or len(r) < len(h):
            continue
        ia, isrc, ismp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
        if not r[ia].isdigit():
            continue
        t = r[isrc].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        mix[op] += int(r[ia]); smp[op] += int(r[ismp]); tot += int(r[ia])
    print(f"  dynamic instructions {tot}, static {sum(1 for _ in mix)} opcodes")
    ts = sum(smp.values())
    for op, c in mix.most_common(14):
        print(f"    {op:8s} {100 * c / tot:5.1f}% of instrs   {100 * smp[op] / max(ts, 1):5.1f}% of samples")
