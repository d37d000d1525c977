# Per-N PCIe probe: every rank moves one C2 batch worth of bytes between PINNED host memory and its GPU,
# all ranks at once.  torchrun --nproc-per-node N scratch/pcie_probe_n.py  ->  one JSON line from rank 0.
#   h2d_only : the end-to-end leg in pieces mode (input up, ~10 % of the output size back)
#   duplex   : the byte-mode leg of round 1 (input up, output down at the same time)
import json, os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_in, n_out = 1_356_000_000, 1_318_000_000
h_in, h_out = torch.empty(n_in, dtype=torch.uint8).pin_memory(), torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n_in, dtype=torch.uint8, device="cuda"), torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=6):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t) / reps], device="cuda")
    if world > 1: dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())
for _ in range(2): run(True, True, 1)
a, b, c = run(True, False), run(False, True), run(True, True)
if rank == 0:
    print(json.dumps({"n_gpus": world, "cpus": os.cpu_count(),
                      "h2d_only": {"ms": a * 1e3, "GBps_per_gpu": n_in / a / 1e9, "GBps_total": world * n_in / a / 1e9},
                      "d2h_only": {"ms": b * 1e3, "GBps_per_gpu": n_out / b / 1e9, "GBps_total": world * n_out / b / 1e9},
                      "duplex": {"ms": c * 1e3, "GBps_total": world * (n_in + n_out) / c / 1e9},
                      "reads_per_s_bound_pieces_mode": world * 4_000_000 / a, "reads_per_s_bound_byte_mode": world * 4_000_000 / c}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
