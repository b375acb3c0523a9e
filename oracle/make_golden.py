"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Usage::

    python -m oracle.make_golden            # rewrites every fixture

Each fixture stores the clip recipe (synthetic generator arguments, so inputs
are regenerated rather than stored), the reference's intermediate masks
(bit-packed), the EMA state planes, a SHA-256 of every output frame, and the
last two output frames in full.  The run is repeated and must reproduce itself
(cv2 4.13.0 showed a first-call nondeterminism in GaussianBlur).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dynamic_video_compression_surveillance_b200.synth import make_clip  # noqa: E402
from oracle import cv2_proxy  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

FD_CASES = {
    # name: (height, width, n_frames, seed, temporal_noise, kwargs of filter_and_dilate_movements)
    "fd_default_96x128": (96, 128, 26, 11, False, {}),
    "fd_main_cfg_64x96": (64, 96, 20, 12, False, dict(block_size=8, kernel_size=10, release_factor=0.3)),
    "fd_minarea50_noise_72x112": (72, 112, 18, 13, True, dict(min_area=50, motion_threshold=6.0, kernel_size=3)),
    # frame size that is not a multiple of the block size: the reference slices the edge blocks (frame_differencing.py:117-121)
    "fd_clipped_126x218": (126, 218, 16, 14, False, dict(min_area=50)),
}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pack(masks):
    return np.stack([np.packbits(m != 0, axis=-1) for m in masks])


def make_fd(name, spec):
    h, w, n, seed, noise, kw = spec
    frames = list(make_clip((h, w), n, seed=seed, temporal_noise=noise).frames())
    a = cv2_proxy.run_reference_fd(frames, **kw)
    b = cv2_proxy.run_reference_fd(frames, **kw)
    for k in ("raw", "filtered", "dilated", "acc", "overlay", "compressed"):
        assert len(a[k]) == n - 1, (k, len(a[k]))
        assert all(np.array_equal(x, y) for x, y in zip(a[k], b[k])), f"reference not reproducible: {k}"
    import cv2
    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"),
        recipe=np.array([h, w, n, seed, int(noise)]),
        kwargs=np.array(repr(kw)),
        cv2_version=np.array(cv2.__version__), numpy_version=np.array(np.__version__),
        raw=pack(a["raw"]), filtered=pack(a["filtered"]), dilated=pack(a["dilated"]),
        acc=np.stack(a["acc"]),
        overlay_sha=np.array([sha(x) for x in a["overlay"]]),
        compressed_sha=np.array([sha(x) for x in a["compressed"]]),
        overlay_tail=np.stack(a["overlay"][-2:]), compressed_tail=np.stack(a["compressed"][-2:]),
        execution_times=np.array(a["execution_times"]),
    )
    print(name, "ok", "static px in last acc:", int((a["acc"][-1] == 0).sum()))


def make_mco(name="mco_compress_64x96"):
    h, w, n, seed = 64, 96, 6, 21
    frames = list(make_clip((h, w), n, seed=seed).frames())
    rng = np.random.default_rng(seed)
    masks = []
    for t in range(n):
        m = np.zeros((h, w), np.uint8)
        x0, y0 = int(rng.integers(0, w - 30)), int(rng.integers(0, h - 20))
        m[y0:y0 + 20, x0:x0 + 30] = 255
        m[int(rng.integers(0, h)), int(rng.integers(0, w))] = 3      # non-binary value as from the lossy codec
        masks.append(m)
    out = cv2_proxy.run_reference_mco_compress(frames, masks)
    assert len(out) == n
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), recipe=np.array([h, w, n, seed]),
                        masks=np.stack(masks), out=np.stack(out))
    print(name, "ok")


def make_window(name="window_vote_48x80"):
    h, w, n, seed = 48, 80, 40, 31
    rng = np.random.default_rng(seed)
    raws = [(rng.random((h, w)) < 0.25).astype(np.uint8) * 255 for _ in range(n)]
    res = {}
    for alpha, K, mk in ((0.2, 30, 2), (0.2, 5, 2), (0.5, 4, 3), (0.34, 7, 5)):
        sm = cv2_proxy.run_reference_window_vote(raws, alpha, K, mk)
        res[f"a{alpha}_K{K}_m{mk}"] = pack(sm)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), recipe=np.array([h, w, n, seed]), raws=pack(raws), **res)
    print(name, "ok")


OF_CASES = {
    # name: (height, width, n_frames, seed, kwargs of temporal_smoothing_flow)
    "of_default_240x352": (240, 352, 44, 41, {}),                                                    # window 30, ellipse 2 (the call site's values)
    "of_k5_m3_80x112": (80, 112, 24, 42, dict(window_size=5, alpha_fraction=0.5, morph_kernel=3, flow_threshold=0.3)),
}


def make_of(name, spec):
    """temporal_smoothing_flow (motion_compression_opt.py:29-109) run unmodified: raw flow masks in, and the voted,
    morphed and rectangle masks the reference produced from them."""
    h, w, n, seed, kw = spec
    clip = make_clip((h, w), n, seed=seed)
    yy, xx = np.mgrid[0:h, 0:w]          # smooth background: Farneback needs texture gradients, not white noise
    clip.background = np.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) * 255 // (w + h))], -1).astype(np.uint8)
    frames = list(clip.frames())
    a = cv2_proxy.run_reference_temporal_smoothing_flow(frames, **kw)
    b = cv2_proxy.run_reference_temporal_smoothing_flow(frames, **kw)
    for k in ("raw", "voted", "morphed", "rect"):
        assert len(a[k]) == n - 1, (k, len(a[k]))
        assert all(np.array_equal(x, y) for x, y in zip(a[k], b[k])), f"reference not reproducible: {k}"
    assert all(np.array_equal(x, y) for x, y in zip(a["overlay"], frames[1:]))      # overlay.mp4 is the unmodified frame (:99)
    import cv2
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), recipe=np.array([h, w, n, seed]), kwargs=np.array(repr(kw)),
                        cv2_version=np.array(cv2.__version__), raw=pack(a["raw"]), voted=pack(a["voted"]),
                        morphed=pack(a["morphed"]), rect=pack(a["rect"]))
    print(name, "ok", "motion px per frame:", [int((m != 0).sum()) for m in a["rect"][-3:]])


def make_config1(name="fd_config1_480x640x300"):
    """BASELINE configs[0]: frame_differencing.py, all defaults, on the synthetic 640x480 300-frame clip; the unmodified
    reference takes ~3 minutes, so only hashes (every mask, overlay and compressed frame) and the reference's own timing
    file are stored."""
    import time
    h, w, n, seed = 480, 640, 300, 0
    frames = list(make_clip((h, w), n, seed=seed).frames())
    t0 = time.time()
    a = cv2_proxy.run_reference_fd(frames)
    dt = time.time() - t0
    import cv2
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), recipe=np.array([h, w, n, seed, 0]), kwargs=np.array(repr({})),
                        cv2_version=np.array(cv2.__version__), numpy_version=np.array(np.__version__),
                        acc_sha=np.array([sha(x) for x in a["acc"]]), raw_sha=np.array([sha(x) for x in a["raw"]]),
                        filtered_sha=np.array([sha(x) for x in a["filtered"]]),
                        overlay_sha=np.array([sha(x) for x in a["overlay"]]),
                        compressed_sha=np.array([sha(x) for x in a["compressed"]]),
                        static_px_last=np.array(int((a["acc"][-1] == 0).sum())),
                        execution_times=np.array(a["execution_times"]),
                        cpu_seconds=np.array(dt), cpu_threads=np.array(cv2.getNumThreads()), cpu_count=np.array(os.cpu_count()))
    print(name, "ok", f"{dt:.1f} s for {n - 1} frames = {dt / (n - 1):.3f} s/frame;", a["execution_times"].replace("\n", " | "))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    only = sys.argv[1:]                       # optional fixture names: regenerate just those
    for name, spec in FD_CASES.items():
        if not only or name in only:
            make_fd(name, spec)
    if not only or "mco_compress_64x96" in only:
        make_mco()
    if not only or "window_vote_48x80" in only:
        make_window()
    for name, spec in OF_CASES.items():
        if not only or name in only:
            make_of(name, spec)
    if "fd_config1_480x640x300" in only:      # ~3 minutes: only on request
        make_config1()


if __name__ == "__main__":
    main()
