"""A reduced pass over every kernel of the library for compute-sanitizer (memcheck / synccheck / racecheck): both loops on small
frames (plain, clipped sizes, a stream group), the stage ops, the large-halo morphology variant, 8x8 / mco / odd block sizes.
Results are checked against the oracle where that is cheap, so a sanitizer-clean run is also a correct one.

    compute-sanitizer --tool memcheck python tools/sanitizer_run.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamic_video_compression_surveillance_b200 import pipeline as P  # noqa: E402
from dynamic_video_compression_surveillance_b200.synth import make_clip  # noqa: E402
from oracle import loops, stage_ops as so  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def loop(mode, h, w, n, S=1, **kw):
    clips = [make_clip((h, w), n, seed=60 + s).frames() for s in range(S)]
    seeds = np.stack([(so.bgr2gray(c[0]) if mode == "window" else loops.first_frame_gray_fd(c[0])) for c in clips])
    body = np.ascontiguousarray(np.stack([c[1:] for c in clips]))
    pipe = P.FramePipeline(w, h, mode, max_batch=5, n_streams=S, **kw)
    pipe.begin_stream(seeds if S > 1 else seeds[0])
    shp = body.shape if S > 1 else body.shape[1:]
    ov = np.empty(shp, np.uint8); cp = np.empty(shp, np.uint8); mk = np.empty(shp[:-1], np.uint8)
    pipe.process_host(body if S > 1 else body[0], ov, cp, mk)
    dbody = dev(body if S > 1 else body[0])
    T = 4
    sl = (slice(None), slice(0, T)) if S > 1 else (slice(0, T),)
    fin = dbody[sl].contiguous()
    pipe.set_overlap(True)
    pipe.process_device(fin, torch.empty_like(fin), torch.empty_like(fin), torch.empty(fin.shape[:-1], dtype=torch.uint8, device="cuda"))
    pipe.flush(); torch.cuda.synchronize()
    blob = pipe.get_state(); pipe.set_state(blob)
    pipe.close()
    for s in range(S):
        ref = (loops.window_loop if mode == "window" else loops.fd_loop)(list(clips[s]), **kw)
        got_m = mk[s] if S > 1 else mk
        assert np.array_equal(got_m, np.stack(ref["mask" if mode == "window" else "acc"])), (mode, h, w, s)
        assert np.array_equal(cp[s] if S > 1 else cp, np.stack(ref["compressed"])), (mode, h, w, s)
    print("loop ok", mode, h, w, "streams", S, kw, flush=True)


def main():
    torch.cuda.set_device(0)
    loop("window", 96, 128, 12, window_size=5, alpha_fraction=0.2, morph_kernel=2, kernel_size=7)
    loop("window", 62, 110, 9, window_size=3, alpha_fraction=0.4, morph_kernel=3, kernel_size=5)          # unaligned width, clipped blocks
    loop("window", 64, 96, 9, S=3, window_size=12, alpha_fraction=0.3, morph_kernel=2, kernel_size=3)       # unfused vote path, group
    loop("fd", 96, 128, 12, min_area=40)
    loop("fd", 62, 110, 9, min_area=30, block_size=8)
    loop("fd", 64, 96, 9, S=2, min_area=30, kernel_size=5)
    r = np.random.default_rng(0)
    # stage ops
    m = (r.random((2, 150, 3840)) < 0.02).astype(np.uint8) * 255
    for op, k, shape in (("close", 15, "rect"), ("open", 15, "rect"), ("dilate", 29, "rect"), ("erode", 10, "rect"), ("close", 5, "ellipse")):
        got = P.morph(dev(m), op, k, shape).cpu().numpy()
        ker = so.structuring_rect(k) if shape == "rect" else so.structuring_ellipse(k)
        fn = {"close": so.morph_close, "open": so.morph_open, "dilate": so.dilate, "erode": so.erode}[op]
        assert np.array_equal(got[0], fn(m[0], ker)), (op, k, shape)
    print("morphology ok", flush=True)
    f = r.integers(0, 256, (2, 70, 100, 3), dtype=np.uint8)
    a = (r.random((2, 70, 100)) < 0.004).astype(np.uint8) * 200
    for bs, fl in ((4, "fd"), (8, "fd"), (8, "mco"), (3, "fd"), (6, "fd")):
        comp, ov = P.degrade_blend(dev(f), dev(a), bs, 100, fl, True)
        ref = so.degrade_fd(f[0], a[0], bs, 100) if fl == "fd" else so.degrade_mco(f[0], a[0])
        assert np.array_equal(comp.cpu().numpy()[0], ref), (bs, fl)
    print("degrade ok", flush=True)
    x = (r.standard_normal((500, 8, 8)) * 50).astype(np.float32)
    P.dct_blocks(dev(x)); P.dct_blocks(dev(x[:, :6, :3].copy()), inverse=True)
    g = r.integers(0, 256, (2, 70, 100), dtype=np.uint8)
    import cv2
    assert np.array_equal(P.gaussian_blur(dev(g), 25, 30.0).cpu().numpy()[0], cv2.GaussianBlur(g[0], (25, 25), 30))
    blobs = np.zeros((2, 120, 160), np.uint8); blobs[:, 20:70, 30:90] = 255; blobs[:, 30:50, 40:60] = 0; blobs[1, 100:104, 5:9] = 255
    assert np.array_equal(P.contour_filter(dev(blobs), 50).cpu().numpy()[0], so.contour_filter_cv2(blobs[0], 50))
    assert np.array_equal(P.mask_rectangles(dev(blobs)).cpu().numpy()[1], so.mask_rectangles_cv2(blobs[1]))
    img = r.integers(0, 256, (2, 90, 120, 3), dtype=np.uint8)
    assert np.array_equal(P.resize_linear(dev(img), (60, 45)).cpu().numpy()[0], cv2.resize(img[0], (60, 45)))
    P.temporal_ring(dev(m[:, :40, :200].copy()), 40, 0.3)
    print("stage ops ok", flush=True)
    torch.cuda.synchronize()
    print("sanitizer_run: all checks passed")


if __name__ == "__main__":
    main()
