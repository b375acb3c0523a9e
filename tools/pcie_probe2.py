"""Chunked pinned copies on two streams without any dependency between them (what the copy pipeline would do if nothing else
mattered): 32 chunks of 50 MB host->device and 2 x 50 MB device->host per pass."""
import os, time, torch
N = 256 * 1080 * 1920 * 3
C = N // 32
h_in = torch.empty(N, dtype=torch.uint8).pin_memory(); h_o1 = torch.empty(N, dtype=torch.uint8).pin_memory(); h_o2 = torch.empty(N, dtype=torch.uint8).pin_memory()
d_in = torch.empty(N, dtype=torch.uint8, device="cuda"); d_o = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def one_pass():
    for c in range(32):
        sl = slice(c * C, (c + 1) * C)
        with torch.cuda.stream(s1): d_in[sl].copy_(h_in[sl], non_blocking=True)
        with torch.cuda.stream(s2):
            h_o1[sl].copy_(d_o[sl], non_blocking=True); h_o2[sl].copy_(d_o[sl], non_blocking=True)
one_pass(); torch.cuda.synchronize()
t_start = float(os.environ.get("START_AT", "0"))
while time.time() < t_start: time.sleep(0.001)
t0 = time.perf_counter()
for _ in range(8): one_pass()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"GPU {os.environ.get('CUDA_VISIBLE_DEVICES')} raw chunked copies: {8 * 256 / dt:.0f} frame-equivalents/s ({8 * 3 * N / dt / 1e9:.1f} GB/s both ways)", flush=True)
