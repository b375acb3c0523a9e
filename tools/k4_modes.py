"""K4 in isolation (stage-level call on resident frames and pre-packed planes is not exposed, so this times the 'degrade'
group of the loop with CUDA events) under the env switches given on the command line.  One line per call."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamic_video_compression_surveillance_b200 import pipeline as P
from dynamic_video_compression_surveillance_b200.synth import make_clip
import bench

T = int(os.environ.get("PROBE_T", "128")); NB = int(os.environ.get("PROBE_NB", "6")); h, w = 1080, 1920
dev = torch.device("cuda")
clip = make_clip("1080p", NB * T + 1, seed=0)
fr = bench.device_clip(clip, NB * T + 1, dev)
cp = torch.empty((NB * T, h, w, 3), dtype=torch.uint8, device=dev); ov = torch.empty_like(cp)
pipe = P.FramePipeline(w, h, "window", max_batch=T, **{k: v for k, v in bench.LOOP.items()})
pipe.begin_stream(P.bgr2gray(fr[:1])[0].cpu().numpy())
res = []
for outs in os.environ.get("PROBE_OUTS", "both").split(","):
    def run():
        for i in range(NB):
            pipe.process_device(fr[1 + i * T:1 + (i + 1) * T], ov[i * T:(i + 1) * T] if outs != "compressed" else None,
                                cp[i * T:(i + 1) * T] if outs != "overlay" else None)
    run(); run(); torch.cuda.synchronize()
    pipe.profile(True); pipe.profile_read()
    for _ in range(6):
        run()
    pr = pipe.profile_read(); pipe.profile(False)
    ms, n = pr["degrade"]
    res.append(f"{outs}: {ms / n / T * 1e3:6.2f} us/frame")
q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.mem,clocks.sm,temperature.gpu,temperature.memory,power.draw", "--format=csv,noheader"],
                   capture_output=True, text=True).stdout.strip()
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith(("DVC_", "PROBE_")))
print(f"{tag:50s} | " + " | ".join(res) + " | " + q, flush=True)
