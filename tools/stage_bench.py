"""Per-stage timings of the stage-level C-ABI entry points at 1080p (CUDA events, 64 frames per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamic_video_compression_surveillance_b200 import pipeline as P
import bench
from dynamic_video_compression_surveillance_b200.synth import make_clip

dev = torch.device("cuda")
T, h, w = 64, 1080, 1920
clip = make_clip("1080p", T + 1, seed=0)
fr = bench.device_clip(clip, T + 1, dev)[1:].contiguous()
mask = torch.zeros((T, h, w), dtype=torch.uint8, device=dev)
for t in range(T):
    for (x, y, ww, hh) in clip.rect_positions(t + 1):
        mask[t, y:y + hh, x:x + ww] = 255


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3 / T       # us per frame


print("degrade fd  bs=4 : %7.2f us/frame" % timeit(lambda: P.degrade_blend(fr, mask, 4, 100, "fd", True)))
print("degrade fd  bs=8 : %7.2f us/frame" % timeit(lambda: P.degrade_blend(fr, mask, 8, 100, "fd", True)))
print("degrade mco bs=8 : %7.2f us/frame" % timeit(lambda: P.degrade_blend(fr, mask, 8, 100, "mco", False)))
print("bgr2gray         : %7.2f us/frame" % timeit(lambda: P.bgr2gray(fr)))
print("morph close e2   : %7.2f us/frame" % timeit(lambda: P.morph(mask, "close", 2, "ellipse")))
print("morph dilate r15 : %7.2f us/frame" % timeit(lambda: P.morph(mask, "dilate", 15, "rect")))
print("contour filter   : %7.2f us/frame" % timeit(lambda: P.contour_filter(mask, 500)))
print("temporal ring K30: %7.2f us/frame" % timeit(lambda: P.temporal_ring(mask, 30, 0.2)))
