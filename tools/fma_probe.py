import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idrlnet_b200 import native as nv
lib = nv.lib()
scratch = torch.zeros(16, device='cuda')
for name in ('pinn_fp32_fma_probe', 'pinn_fp32_fma2_probe'):
    fn = getattr(lib, name)
    st = torch.cuda.current_stream().cuda_stream
    for blocks in (148 * 4, 148 * 8):
        fn(2000, blocks, scratch.data_ptr(), st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        flops = fn(20000, blocks, scratch.data_ptr(), st)
        b.record()
        torch.cuda.synchronize()
        print(name, blocks, '%.1f TFLOP/s' % (flops / (a.elapsed_time(b) * 1e-3) / 1e12))
