import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamic_video_compression_surveillance_b200 import pipeline as P
from oracle import stage_ops as so
r = np.random.default_rng(5)
t, h, w = 4, 240, 640
frames = r.integers(0, 256, (t, h, w, 3), dtype=np.uint8)
acc = np.zeros((t, h, w), np.uint8)
q = float(sys.argv[1]) if len(sys.argv) > 1 else 0.011
comp = P.degrade_blend(torch.from_numpy(frames).cuda(), torch.from_numpy(acc).cuda(), 4, q, "fd", False)[0].cpu().numpy()
out = []
for i in range(t):
    ref = so.degrade_fd(frames[i], acc[i], 4, q)
    ycc = so.bgr2ycrcb(frames[i])
    d = (comp[i] != ref).any(axis=2).reshape(h // 4, 4, w // 4, 4).any(axis=(1, 3))
    for by, bx in zip(*np.nonzero(d)):
        blk = ycc[by * 4:by * 4 + 4, bx * 4:bx * 4 + 4, 0]
        out.append(dict(y=blk.tolist(), got=comp[i][by * 4:by * 4 + 4, bx * 4:bx * 4 + 4, 0].tolist(),
                        ref=ref[by * 4:by * 4 + 4, bx * 4:bx * 4 + 4, 0].tolist(), bx=int(bx), by=int(by)))
import json
json.dump(dict(q=q, cases=out[:12]), open("gpurun_out/k4_q_dump.json", "w"))
print(len(out), "bad blocks; first:", out[0] if out else None)
