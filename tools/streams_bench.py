"""BASELINE config 4 on one GPU: S independent 1080p streams (own handle, own state), round-robin batches of B frames each,
frames resident in HBM.  Prints aggregate frames/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamic_video_compression_surveillance_b200 import pipeline as P
from dynamic_video_compression_surveillance_b200.synth import make_clip
import bench
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
R = int(sys.argv[3]) if len(sys.argv) > 3 else 6
h, w = 1080, 1920
dev = torch.device("cuda")
clips = []
for s in range(min(S, 4)):                      # four distinct clips are enough to keep the content realistic
    clips.append(bench.device_clip(make_clip("1080p", R * B + 1, seed=s), R * B + 1, dev))
pipes = []
for s in range(S):
    p = P.FramePipeline(w, h, "window", max_batch=B, **bench.LOOP)
    p.begin_stream(P.bgr2gray(clips[s % 4][:1])[0].cpu().numpy())
    p.set_overlap(True)
    pipes.append(p)
ov = torch.empty((S, B, h, w, 3), dtype=torch.uint8, device=dev); cp = torch.empty_like(ov)
def one_round(r):
    for s, p in enumerate(pipes):
        p.process_device(clips[s % 4][1 + r * B:1 + (r + 1) * B], ov[s], cp[s])
for p in pipes: p.flush()
one_round(0); [p.flush() for p in pipes]; torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(1, R):
    one_round(r)
for p in pipes: p.flush()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"{S} streams x {B}-frame batches: {(R - 1) * S * B / (ms * 1e-3):.0f} frames/s aggregate ({ms / ((R - 1) * S):.3f} ms per stream batch)")
