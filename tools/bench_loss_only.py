import sys, torch, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for config, prec in (("burgers", "fp32"), ("ldc", "tf32"), ("ldc", "fp32x3")):
    dev = torch.device("cuda:0")
    tr = bench.Trainer(bench.workload_configs()[config], dev, precision=prec)
    rng = np.random.default_rng(0)
    hb = tr.sample_host(rng)
    tr.bind_set({k: torch.as_tensor(v, device=dev) for k, v in hb.items()})
    eng = tr.engine
    eng.domains = tr.sets[0]
    def t(fn, reps=20):
        for _ in range(3): fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    print(config, prec, 'full %.3f ms  loss-only %.3f ms' % (t(lambda: eng.launch()), t(lambda: eng.launch(loss_only=True))))
