"""One-screen table of a bench.py JSON line (headline + configs + reference arms).   python tools/bench_table.py file.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])


def row(k, e):
    if not e or "error" in e:
        print(f"{k:16s} ERROR {e and e['error']}")
        return
    r = e.get("roofline") or {}
    tc = e.get("torch_cuda") or {}
    st, tp = tc.get("strict", {}), tc.get("train_pipe", {})
    print(f"{k:16s} {e['value']:10.4g} pts/s {e['ms_per_step']:9.3f} ms/step  frac {r.get('frac') or 0:5.3f} ({r.get('bound')})  kernel {r.get('kernel_ms') or 0:8.3f} ms | "
          f"ref-CUDA strict {st.get('value') or 0:10.4g} (pipe {tp.get('value') or 0:9.4g}) at {tc.get('points_per_step')} pts, {tc.get('max_memory_gb') or 0:5.1f} GB"
          f"  => x{e.get('vs_torch_cuda') or 0:6.2f}")


print(f"N={d['n_gpus']} steps={d['steps']} clocks={d['clocks']}")
row(d["config"]["workload"].split(" ")[0][:16], d)
print(f"  e2e {d['e2e']['value']:.4g} (h2d {d['e2e']['h2d_bytes_per_step']} B/step)  device-sampled e2e {(d.get('e2e_device_sampling') or {}).get('value')}"
      f"  host {d.get('host_ms_per_step'):.3f} ms/step  launches {d['gpu_launches']}  cpu_baseline {d.get('cpu_baseline') and d['cpu_baseline']['value']}")
for k, e in d.get("configs", {}).items():
    row(k, e)
