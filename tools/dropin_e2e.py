"""End-to-end timing of the drop-in entry point on a real mp4v file (decode -> GPU loop -> two encodes)."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic_video_compression_surveillance_b200", "dropin")); sys.path.insert(1, ROOT)
import cv2, numpy as np
from dynamic_video_compression_surveillance_b200 import host_loop
from dynamic_video_compression_surveillance_b200.synth import make_clip
import frame_differencing as fd

h, w, n = 1080, 1920, int(sys.argv[1]) if len(sys.argv) > 1 else 240
clip = make_clip("1080p", n, seed=1)
yy, xx = np.mgrid[0:h, 0:w]
clip.background = np.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) * 255 // (w + h))], -1).astype(np.uint8)
tmp = tempfile.mkdtemp()
src = os.path.join(tmp, "cam.mp4")
wr = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*"mp4v"), 30, (w, h))
buf = np.empty((h, w, 3), np.uint8)
for t in range(n):
    wr.write(clip.render_into(t, buf))
wr.release()
cap = cv2.VideoCapture(src); t0 = time.time(); k = 0
while cap.read()[0]:
    k += 1
print(f"decode only: {k / (time.time() - t0):.1f} fps")
orig = host_loop.run_fd_stream
for threaded in (False, True, True):
    host_loop.run_fd_stream = lambda *a, **kw: orig(*a, threaded=threaded, **kw)
    t0 = time.time()
    fd.filter_and_dilate_movements(src, os.path.join(tmp, f"out_{threaded}"), max_batch=16)
    dt = time.time() - t0
    print(f"threaded={threaded}: {n - 1} frames in {dt:.2f} s = {(n - 1) / dt:.1f} fps")
