"""Time (and let ncu capture) the 8x8 degrade kernels on resident 1080p frames: fd flavour with block_size 8 and the mco flavour."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamic_video_compression_surveillance_b200 import pipeline as P
n, h, w = 96, 1080, 1920
g = torch.Generator(device="cuda").manual_seed(0)
frames = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
masks = torch.zeros((n, h, w), dtype=torch.uint8, device="cuda")
masks[:, 300:500, 600:900] = 255
for fl in ("fd", "mco"):
    for _ in range(2):
        P.degrade_blend(frames, masks, 8, 100, fl, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        P.degrade_blend(frames, masks, 8, 100, fl, False)
    e1.record(); torch.cuda.synchronize()
    print(fl, f"{1e3 * e0.elapsed_time(e1) / (5 * n):.2f} us per 1080p frame (whole stage-level call)", flush=True)
