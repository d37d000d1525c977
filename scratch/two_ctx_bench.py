"""Several contexts on ONE device, each driven by its own host thread on its own stream, batches resident in HBM:
does the device overlap one context's issue-bound k_trim with another's memory-bound k_frame / k_emit?"""
import sys, time, threading, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from faqcs_b200 import synth
from faqcs_b200.api import Engine, Options

PAIRS = 2_000_000
w = synth.c2(PAIRS)
d1 = torch.from_numpy(np.asarray(w.r1)).cuda(); d2 = torch.from_numpy(np.asarray(w.r2)).cuda()
ITERS = 30
for n_eng in (1, 2, 3):
    engines = [Engine(Options(discard_output=True)) for _ in range(n_eng)]
    for e in engines:
        e.autodetect(w.r1, w.r2)
        for _ in range(3):
            e.process_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel(), 0, True)
    torch.cuda.synchronize()
    bar = threading.Barrier(n_eng + 1)
    def work(e):
        bar.wait()
        for _ in range(ITERS):
            e.process_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel(), 0, True)
    th = [threading.Thread(target=work, args=(e,)) for e in engines]
    for t in th: t.start()
    bar.wait(); t0 = time.perf_counter()
    for t in th: t.join()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(os.environ.get("FAQCS_B200_LIB", "base").split("/")[-1], "ctas/sm", os.environ.get("FAQCS_B200_TRIM_CTAS_PER_SM", "-"), "engines", n_eng,
          "ms/batch", round(dt / (ITERS * n_eng) * 1e3, 3), "G reads/s", round(2 * PAIRS * ITERS * n_eng / dt / 1e9, 3), flush=True)
    for e in engines: e.close()
