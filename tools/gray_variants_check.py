"""K1 gray-conversion variants of the measure build (DVC_LIB_FLAVOUR=measure DVC_GRAY_IMPL=0|1|2): window-mode masks on
random frames against the oracle (random frames make every gray value matter, threshold 3 keeps the mask non-trivial)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamic_video_compression_surveillance_b200 import pipeline as P, _lib  # noqa: E402
from oracle import loops, stage_ops as so  # noqa: E402

assert _lib.load().dvc_measure_build() == 1, "needs the -DDVC_MEASURE flavour"
r = np.random.default_rng(31)
h, w, n = 64, 160, 12
frames = r.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
frames[1::2] = np.clip(frames[0:-1:2].astype(int) + r.integers(-4, 5, frames[1::2].shape), 0, 255).astype(np.uint8)
cfg = dict(window_size=3, alpha_fraction=0.5, morph_kernel=0, kernel_size=0, motion_threshold=3.0)
ref = loops.window_loop(list(frames), degrade=False, **cfg)
pipe = P.FramePipeline(w, h, "window", max_batch=16, **cfg)
pipe.begin_stream(so.bgr2gray(frames[0]))
mk = torch.empty((n - 1, h, w), dtype=torch.uint8, device="cuda")
pipe.process_device(torch.from_numpy(frames[1:]).cuda(), None, None, mk)
torch.cuda.synchronize()
assert np.array_equal(mk.cpu().numpy(), np.stack(ref["mask"]))
print("masks equal for DVC_GRAY_IMPL =", os.environ.get("DVC_GRAY_IMPL"))
