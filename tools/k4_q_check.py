"""Exact comparison of the block_size-4 degrade kernel with the oracle on many random blocks for awkward quantisation levels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamic_video_compression_surveillance_b200 import pipeline as P
from oracle import stage_ops as so
r = np.random.default_rng(5)
t, h, w = 4, 240, 640
frames = r.integers(0, 256, (t, h, w, 3), dtype=np.uint8)
acc = np.zeros((t, h, w), np.uint8)
print("cv2 dct4 closed form:", so.cv2_dct4_matches_closed_form(), "selection:", {k: v for k, v in os.environ.items() if k.startswith("DVC_")})
for q in (0.3, 7.3, 0.05, 33.3, 100.0, 3.0, 0.011, 1234.5):
    comp, _ = P.degrade_blend(torch.from_numpy(frames).cuda(), torch.from_numpy(acc).cuda(), 4, q, "fd", False)
    comp = comp.cpu().numpy()
    bad_blocks = 0
    for i in range(t):
        ref = so.degrade_fd(frames[i], acc[i], 4, q)
        d = (comp[i] != ref).any(axis=2)
        bad_blocks += int(d.reshape(h // 4, 4, w // 4, 4).any(axis=(1, 3)).sum())
    print(f"q={q}: {bad_blocks} differing blocks of {t * (h // 4) * (w // 4)}")
