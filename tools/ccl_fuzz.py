"""Random shapes / masks / thresholds through dvc_contour_filter_u8 (the one-launch kernel) against cv2's findContours /
contourArea / drawContours.  Complements the fixed cases of tests/test_gpu_parity.py::test_contour_filter*.

    python tools/ccl_fuzz.py [n_cases] [seed]
"""
import os
import sys

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamic_video_compression_surveillance_b200 import pipeline as P  # noqa: E402
from oracle import stage_ops as so  # noqa: E402


def random_mask(r, h, w):
    kind = int(r.integers(0, 6))
    m = np.zeros((h, w), np.uint8)
    if kind == 0:
        m = (r.random((h, w)) < r.random() ** 2).astype(np.uint8) * 255
    elif kind == 1:
        for _ in range(int(r.integers(1, 40))):
            cx, cy = int(r.integers(0, w)), int(r.integers(0, h))
            cv2.ellipse(m, (cx, cy), (int(r.integers(1, max(2, w // 3))), int(r.integers(1, max(2, h // 3)))), float(r.integers(0, 180)),
                        0, 360, 255, int(r.choice([-1, 1, 2, 3])))
    elif kind == 2:
        for _ in range(int(r.integers(1, 30))):
            x, y = int(r.integers(0, w)), int(r.integers(0, h))
            cv2.rectangle(m, (x, y), (x + int(r.integers(1, 80)), y + int(r.integers(1, 60))), 255, int(r.choice([-1, 1, 2, 5])))
    elif kind == 3:
        g = cv2.GaussianBlur(r.integers(0, 256, (h, w)).astype(np.uint8), (5, 5), 0)
        m = (cv2.absdiff(g, cv2.GaussianBlur(r.integers(0, 256, (h, w)).astype(np.uint8), (5, 5), 0)) > int(r.integers(5, 40))).astype(np.uint8) * 255
    elif kind == 4:
        for y in range(1, h, 2):
            m[y, :] = 255
            m[y, int(r.integers(0, w))] = 0
    else:
        m[:] = 255
        for _ in range(int(r.integers(1, 30))):
            x, y = int(r.integers(0, w)), int(r.integers(0, h))
            cv2.circle(m, (x, y), int(r.integers(1, 20)), 0, -1)
    if r.random() < 0.5:
        noise = r.random((h, w)) < 0.01
        m[noise] = 255 - m[noise]
    return m


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    r = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    for case in range(n):
        h = int(r.choice([1, 2, 7, 16, 17, 31, 60, 97, 128, 270, 540]))
        w = int(r.choice([1, 5, 31, 32, 33, 63, 64, 65, 127, 200, 480, 960, 1920, 2049, 2300, 4100]))
        if h * w > 1_200_000:
            h = max(1, 1_200_000 // w)
        k = int(r.integers(1, 5))
        masks = np.stack([random_mask(r, h, w) for _ in range(k)])
        min_area = float(r.choice([0, 0.5, 3, 20, 100, 500, 2500]))
        got = P.contour_filter(torch.from_numpy(masks).cuda(), min_area).cpu().numpy()
        for i in range(k):
            if not np.array_equal(got[i], so.contour_filter_cv2(masks[i], min_area)):
                bad += 1
                print("MISMATCH case", case, "frame", i, "shape", (h, w), "min_area", min_area)
    print("contour filter fuzz:", n, "cases,", bad, "mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
