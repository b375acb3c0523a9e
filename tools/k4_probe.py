"""Box probe: raw copy / fill / read bandwidth with torch next to the loop's per-kernel times (1080p, batch 128)."""
import subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamic_video_compression_surveillance_b200 import pipeline as P
from dynamic_video_compression_surveillance_b200.synth import make_clip

print(subprocess.run(["nvidia-smi", "--query-gpu=serial,temperature.gpu,clocks.mem,clocks.sm", "--format=csv,noheader"],
                     capture_output=True, text=True).stdout.strip())
dev = torch.device("cuda")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


N = 1 << 31
x = torch.empty(N, dtype=torch.uint8, device=dev)
y = torch.empty(N, dtype=torch.uint8, device=dev)
print("copy  GB/s %.0f" % (2 * N / timeit(lambda: y.copy_(x)) / 1e6))
print("fill  GB/s %.0f" % (N / timeit(lambda: y.zero_()) / 1e6))
del x, y
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
T, h, w = 128, 1080, 1920
clip = make_clip("1080p", 3 * T + 1, seed=0)
fr = bench.device_clip(clip, 3 * T + 1, dev)
cp = torch.empty((3 * T, h, w, 3), dtype=torch.uint8, device=dev); ov = torch.empty_like(cp)
pipe = P.FramePipeline(w, h, "window", max_batch=T, **{k: v for k, v in bench.LOOP.items()})
pipe.begin_stream(P.bgr2gray(fr[:1])[0].cpu().numpy())
for outs in ("both", "compressed", "overlay"):
    def run():
        for i in range(3):
            pipe.process_device(fr[1 + i * T:1 + (i + 1) * T], ov[i * T:(i + 1) * T] if outs != "compressed" else None,
                                cp[i * T:(i + 1) * T] if outs != "overlay" else None)
    run(); torch.cuda.synchronize()
    pipe.profile(True); pipe.profile_read()
    for _ in range(5):
        run()
    pr = pipe.profile_read(); pipe.profile(False)
    ms, n = pr["degrade"]
    nb = T * h * w * (3.125 + 3 * (2 if outs == "both" else 1))
    print(f"K4 {outs:10s} {ms / n * 1e3:8.1f} us per {T} frames  {nb / (ms / n) / 1e6:6.0f} GB/s   front {pr['front'][0] / pr['front'][1] * 1e3:.1f} us")
