"""Pinned-memory copy bandwidth per GPU, alone and with all ranks at once (run under torchrun)."""
import os, sys, time, subprocess
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    print(subprocess.run("lscpu | grep -i -E 'numa|model name|^cpu\\(s\\)'", shell=True, capture_output=True, text=True).stdout)
    print("affinity", sorted(os.sched_getaffinity(0)))
N = 1 << 30
h_in = torch.empty(N, dtype=torch.uint8).pin_memory(); h_out = torch.empty(2 * N, dtype=torch.uint8).pin_memory()
d_in = torch.empty(N, dtype=torch.uint8, device="cuda"); d_out = torch.empty(2 * N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(mode):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return 4 * N / dt / 1e9, 8 * N / dt / 1e9
for mode in ("h2d", "d2h", "both"):
    a, b = run(mode)
    msg = f"rank {rank} {mode}: " + (f"H2D {a:.1f} GB/s " if mode != "d2h" else "") + (f"D2H {b:.1f} GB/s" if mode != "h2d" else "")
    print(msg, flush=True)
    if world > 1: dist.barrier()
if world > 1: dist.destroy_process_group()
