"""The lane-level model of the one-launch contour filter (tools/ccl_sweep_model.py, the restatement csrc/k_ccl_sweep.cuh was
written from) against the oracle's contour filter, on small masks with small words: run boundaries, neighbour bits and the
union-find links of both phases fall on lane boundaries all the time.  The GPU kernel itself is checked against cv2 in
tests/test_gpu_parity.py::test_contour_filter*."""
import importlib.util
import os

import numpy as np

from oracle import stage_ops as so

_spec = importlib.util.spec_from_file_location(
    "ccl_sweep_model", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "ccl_sweep_model.py"))
model = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(model)


def test_model_matches_oracle_on_random_masks():
    assert model.main(400, seed=5) == 0


def test_model_holes_inside_holes_and_runs_on_word_boundaries():
    h, w = 20, 33
    ring = np.zeros((h, w), np.uint8)
    for k in range(0, 9, 2):
        ring[k:h - k, k:w - k] = 255 if (k // 2) % 2 == 0 else 0
    bars = np.zeros((h, w), np.uint8)
    bars[::3, 8:] = 255
    bars[1::3, :16] = 255
    for img in (ring, bars, 255 - ring):
        for wb in (2, 4, 8):
            for min_area in (0, 3, 40):
                got, _ = model.Model(h, w, wb, 3).contour_filter(img, min_area)
                assert np.array_equal(got, so.contour_filter(img, min_area)), (wb, min_area)
