import os, sys, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
print("rank", rank, {k: v for k, v in os.environ.items() if "NCCL" in k}, file=sys.stderr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = torch.ones(4, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
print("done", rank, file=sys.stderr)
dist.destroy_process_group()
