"""Helpers shared by the golden-fixture tests (CPU oracle tests and GPU parity tests)."""
import ast
import hashlib
import os

import numpy as np

from dynamic_video_compression_surveillance_b200.synth import make_clip

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FD_FIXTURES = ["fd_default_96x128", "fd_main_cfg_64x96", "fd_minarea50_noise_72x112"]
# frame size not a multiple of the block size: edge blocks clipped by the reference (frame_differencing.py:117-121)
FD_CLIPPED_FIXTURE = "fd_clipped_126x218"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def unpack(bits, w):
    return (np.unpackbits(bits, axis=-1)[..., :w] * 255).astype(np.uint8)


def load_fd(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    h, w, n, seed, noise = (int(v) for v in z["recipe"])
    kw = ast.literal_eval(str(z["kwargs"]))
    frames = make_clip((h, w), n, seed=seed, temporal_noise=bool(noise)).frames()
    return z, frames, kw, (h, w, n)


OF_FIXTURES = ["of_default_240x352", "of_k5_m3_80x112"]


def load_of(name):
    """Masks tapped from the UNMODIFIED temporal_smoothing_flow (oracle/make_golden.py::make_of): raw flow masks in, voted /
    morphed / rectangle masks out, plus the keyword arguments of the run."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    h, w, n, seed = (int(v) for v in z["recipe"])
    kw = dict(flow_threshold=0.5, alpha_fraction=0.2, window_size=30, morph_kernel=2)       # motion_compression_opt.py:29-30
    kw.update(ast.literal_eval(str(z["kwargs"])))
    return {k: unpack(z[k], w) for k in ("raw", "voted", "morphed", "rect")}, kw, (h, w, n)
