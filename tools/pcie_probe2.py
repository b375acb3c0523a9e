"""Chunked pinned copies on two streams without any dependency between them (what the copy pipeline would do if nothing else
mattered): 32 chunks of 50 MB host->device and 2 x 50 MB device->host per pass.  PROBE_MODE = both | h2d | d2h;
START_AT = wall-clock second at which every process of a contention test starts its timed passes."""
import os, time, torch
MODE = os.environ.get("PROBE_MODE", "both")
N = 256 * 1080 * 1920 * 3
C = N // 32
h_in = torch.empty(N, dtype=torch.uint8).pin_memory(); h_o1 = torch.empty(N, dtype=torch.uint8).pin_memory(); h_o2 = torch.empty(N, dtype=torch.uint8).pin_memory()
d_in = torch.empty(N, dtype=torch.uint8, device="cuda"); d_o = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def one_pass():
    for c in range(32):
        sl = slice(c * C, (c + 1) * C)
        if MODE in ("both", "h2d"):
            with torch.cuda.stream(s1): d_in[sl].copy_(h_in[sl], non_blocking=True)
        if MODE in ("both", "d2h"):
            with torch.cuda.stream(s2):
                h_o1[sl].copy_(d_o[sl], non_blocking=True); h_o2[sl].copy_(d_o[sl], non_blocking=True)
one_pass(); torch.cuda.synchronize()
t_start = float(os.environ.get("START_AT", "0"))
while time.time() < t_start: time.sleep(0.001)
t0 = time.perf_counter()
for _ in range(8): one_pass()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
nbytes = {"both": 3, "h2d": 1, "d2h": 2}[MODE] * N
print(f"GPU {os.environ.get('CUDA_VISIBLE_DEVICES')} {MODE}: {8 * 256 / dt:.0f} frame-equivalents/s ({8 * nbytes / dt / 1e9:.1f} GB/s)", flush=True)
