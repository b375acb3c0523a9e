"""Random loop configurations against the oracle loops (masks, overlay and compressed frames exact).  One line per case."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dynamic_video_compression_surveillance_b200 import pipeline as P
from dynamic_video_compression_surveillance_b200.synth import make_clip
from oracle import loops, stage_ops as so

r = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
exact = so.cv2_dct4_matches_closed_form()
bad = 0
for case in range(int(sys.argv[1]) if len(sys.argv) > 1 else 16):
    h = 4 * int(r.integers(8, 40)); w = 8 * int(r.integers(6, 40)); n = int(r.integers(6, 40))
    clip = make_clip((h, w), n, seed=int(r.integers(0, 1000)), temporal_noise=bool(r.integers(0, 2)))
    frames = clip.frames()
    mb = int(r.choice([1, 3, 8, 16]))
    ov = np.empty((n - 1, h, w, 3), np.uint8); cp = np.empty_like(ov); mk = np.empty((n - 1, h, w), np.uint8)
    if case % 2 == 0:
        kw = dict(window_size=int(r.choice([1, 2, 5, 9, 30, 31])), alpha_fraction=float(r.choice([0.0, 0.2, 0.5, 0.9, 1.0])),
                  morph_kernel=int(r.choice([0, 2, 3, 5])), morph_shape=str(r.choice(["ellipse", "rect"])),
                  kernel_size=int(r.choice([0, 1, 2, 7, 10, 15])), motion_threshold=float(r.choice([0.5, 3.0, 40.0, 200.0])),
                  quantization_level=float(r.choice([100, 30, 7.5])))
        ref = loops.window_loop(list(frames), **kw)
        pipe = P.FramePipeline(w, h, "window", max_batch=mb, **kw)
        pipe.begin_stream(so.bgr2gray(frames[0]))
        mask_ref = np.stack(ref["mask"])
    else:
        kw = dict(min_area=float(r.choice([0, 20, 500, 2000])), kernel_size=int(r.choice([1, 3, 7, 8, 12])),
                  release_factor=float(r.choice([0.5, 0.3, 0.05, 0.9])), motion_threshold=float(r.choice([0.5, 2.0, 25.0])),
                  quantization_level=float(r.choice([100, 30, 7.5])))
        ref = loops.fd_loop(list(frames), **kw)
        pipe = P.FramePipeline(w, h, "fd", max_batch=mb, **kw)
        pipe.begin_stream(loops.first_frame_gray_fd(frames[0]))
        mask_ref = np.stack(ref["acc"])
    for i in range(0, n - 1, mb):                       # several process_host calls: state carried across calls
        j = min(n - 1, i + mb)
        pipe.process_host(np.ascontiguousarray(frames[1 + i:1 + j]), ov[i:j], cp[i:j], mk[i:j])
    pipe.close()
    ok_m = np.array_equal(mk, mask_ref)
    ok_o = np.array_equal(ov, np.stack(ref["overlay"]))
    d = np.abs(cp.astype(int) - np.stack(ref["compressed"]).astype(int))
    ok_c = d.max() == 0 if exact else np.mean(d > 1) < 1e-3
    bad += not (ok_m and ok_o and ok_c)
    print(case, (n, h, w), mb, kw, "mask", ok_m, "overlay", ok_o, "compressed", ok_c, int(d.max()), flush=True)
print("FAILED" if bad else "all cases equal", bad)
