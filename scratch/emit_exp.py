import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from faqcs_b200 import synth
from faqcs_b200.api import Engine, Options
dev = torch.device('cuda', 0)
w = synth.c2(250000)
reps = 8
d1 = torch.from_numpy(w.r1).to(dev).repeat(reps); d2 = torch.from_numpy(w.r2).to(dev).repeat(reps)
for name, opt in (("default", Options()), ("q0_all_plain", Options(quality=0, low_complexity_cutoff_ratio=1.0, max_num_poly_N=200)),
                  ("qc_only", Options(qc_only=True))):
    with Engine(opt) as e:
        e.autodetect(w.r1, w.r2)
        for _ in range(3):
            r = e.process_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel())
        t = e.last_timing()
        print(name, {k: round(v, 3) for k, v in t.items()}, 'out', sum(r.stream_bytes) / 1e6, 'MB valid', r.n_valid)
# plain D2D copy of the same volume for reference
a = torch.empty(d1.numel() + d2.numel(), dtype=torch.uint8, device=dev); b = torch.empty_like(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): b.copy_(a)
e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
print('torch copy of', a.numel() / 1e6, 'MB:', e0.elapsed_time(e1), 'ms')
