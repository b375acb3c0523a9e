"""Contour filter on masks the bench does not have: dense salt-and-pepper noise (more row runs than the shared-memory node arrays
hold: the one-launch kernel falls back to its global arrays) and many small blobs.  Prints microseconds per 1080p frame.

    DVC_LIB_FLAVOUR=measure DVC_CCL_SWEEP=0|1 python tools/ccl_noise_bench.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamic_video_compression_surveillance_b200 import pipeline as P  # noqa: E402


def main():
    r = np.random.default_rng(3)
    h, w, n = 1080, 1920, 32
    cases = {}
    for d in (0.0005, 0.005, 0.05, 0.5):
        cases["noise %g" % d] = (r.random((n, h, w)) < d).astype(np.uint8) * 255
    blobs = np.zeros((n, h, w), np.uint8)
    for k in range(n):
        for _ in range(300):
            y, x = int(r.integers(0, h - 40)), int(r.integers(0, w - 40))
            blobs[k, y:y + int(r.integers(3, 40)), x:x + int(r.integers(3, 40))] = 255
    cases["300 blobs"] = blobs
    for name, m in cases.items():
        d = torch.from_numpy(m).cuda()
        P.contour_filter(d, 500)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            P.contour_filter(d, 500)
        b.record()
        torch.cuda.synchronize()
        print("%-14s %8.1f us per frame (incl. packing / unpacking the uint8 masks)" % (name, a.elapsed_time(b) * 1000 / (3 * n)))


if __name__ == "__main__":
    main()
