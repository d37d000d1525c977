# PCIe probe: H2D, D2H and simultaneous bidirectional copy rates from pinned host memory (sizes of one C2 bench step).
import torch, time
n_in, n_out = 1356_000_000, 1250_000_000
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
for _ in range(2): run(True, True, 1)
a = run(True, False); b = run(False, True); c = run(True, True)
print(f"H2D alone {n_in/a/1e9:.1f} GB/s ({a*1e3:.1f} ms)  D2H alone {n_out/b/1e9:.1f} GB/s ({b*1e3:.1f} ms)  both {c*1e3:.1f} ms -> H2D {n_in/c/1e9:.1f} + D2H {n_out/c/1e9:.1f} GB/s")
