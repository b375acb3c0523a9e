"""End-to-end (host buffers) throughput of dvc_process_host on this process's GPU; start several at once to see contention."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamic_video_compression_surveillance_b200 import pipeline as P
import bench
h, w, ne = 1080, 1920, 256
pipe = P.FramePipeline(w, h, "window", max_batch=128, **bench.LOOP)
pipe.begin_stream(np.zeros((h, w), np.uint8))
hin = P.pinned_empty((ne, h, w, 3)); hov = P.pinned_empty((ne, h, w, 3)); hcp = P.pinned_empty((ne, h, w, 3))
hin.random_(0, 255)
for _ in range(2):
    pipe.process_host(hin, hov, hcp)
t_start = float(os.environ.get("START_AT", "0"))
while time.time() < t_start:
    time.sleep(0.001)
t0 = time.perf_counter()
for _ in range(8):
    pipe.process_host(hin, hov, hcp)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"GPU {os.environ.get('CUDA_VISIBLE_DEVICES')} chunk {os.environ.get('DVC_HOST_CHUNK', 'default')}: {8 * ne / dt:.0f} fps", flush=True)
