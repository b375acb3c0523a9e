"""Fuzz the block_size-4 degrade kernel over random geometries: prints one checksum line per case.  Run it twice with different
kernel selections (e.g. default and DVC_K4_PERSIST=0, or DVC_K4_PACKED=0) and diff the outputs."""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamic_video_compression_surveillance_b200 import pipeline as P

r = np.random.default_rng(2024)
for case in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    h = 4 * int(r.integers(1, 60)); w = 8 * int(r.integers(1, 400 if case % 5 else 700)); t = int(r.integers(1, 10))
    frames = r.integers(0, 256, (t, h, w, 3), dtype=np.uint8)
    dens = float(r.choice([0.0, 0.001, 0.02, 0.5]))
    acc = (r.random((t, h, w)) < dens).astype(np.uint8) * r.integers(1, 256, (t, h, w), dtype=np.uint8)
    q = float(r.choice([100.0, 100.0, 7.3, 0.3, 250.0]))
    want_ov = bool(case % 3)
    cnt = torch.zeros(5, dtype=torch.int64, device="cuda")
    comp, ov = P.degrade_blend(torch.from_numpy(frames).cuda(), torch.from_numpy(acc).cuda(), 4, q, "fd", want_ov, counters=cnt)
    torch.cuda.synchronize()
    hc = hashlib.md5(comp.cpu().numpy().tobytes()).hexdigest()[:12]
    ho = hashlib.md5(ov.cpu().numpy().tobytes()).hexdigest()[:12] if ov is not None else "-"
    print(case, (t, h, w), q, dens, hc, ho, cnt.cpu().tolist())
