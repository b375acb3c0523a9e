"""The per-frame loop bodies restated on arrays (CPU port of the hot path).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Same cv2 calls in the same
order as the reference, with file I/O, logging and timing removed:

* ``fd_loop``      frame_differencing.py:67-133 (the drop-in, "fd-exact" mode)
* ``window_loop``  the loop BASELINE.json's north_star names: gray/absdiff/threshold
  (frame_differencing.py:92,96-97, no pre-blur) -> K-frame window vote and
  close/open (motion_compression_opt.py:61-62,84-90) -> dilate, overlay and
  block-DCT degrade (frame_differencing.py:106,110-130)
* ``mco_compress`` motion_compression_opt.py:141-185

``literal_blocks=True`` runs the reference's Python double loop over blocks
verbatim (that is what the reference spends 93-99 % of its time in and what
bench.py times as the CPU baseline); the default uses the vectorised form that
tests/test_oracle_vs_cv2.py shows to be identical.
"""
from __future__ import annotations

from collections import deque

import cv2
import numpy as np

from . import stage_ops as so


def _stable_blur(gray, ksize, sigma):
    """cv2.GaussianBlur, guarded: cv2 4.13.0's multi-threaded fixed-point blur races on tiny images
    (height of the order of the thread count; row 1 comes back with garbage in some calls).  Frames
    under 64 rows are blurred single-threaded, where the result is deterministic and equals the
    closed form in stage_ops.gaussian_blur5."""
    if gray.shape[0] >= 64:
        return cv2.GaussianBlur(gray, ksize, sigma)
    n = cv2.getNumThreads()
    cv2.setNumThreads(1)
    try:
        return cv2.GaussianBlur(gray, ksize, sigma)
    finally:
        cv2.setNumThreads(n)


def first_frame_gray_fd(frame0: np.ndarray) -> np.ndarray:
    """frame_differencing.py:75-77: gray of frame 0 with the heavy (25,25),sigma=30 blur."""
    return _stable_blur(cv2.cvtColor(frame0, cv2.COLOR_BGR2GRAY), (25, 25), 30)


def literal_degrade_fd(frame, acc, block_size, quantization_level):
    """frame_differencing.py:115-130, the Python block loop kept as written."""
    h, w = acc.shape
    frame_ycrcb = cv2.cvtColor(frame, cv2.COLOR_BGR2YCrCb)
    channels = list(cv2.split(frame_ycrcb))
    for y in range(0, h, block_size):
        for x in range(0, w, block_size):
            if acc[y:y + block_size, x:x + block_size].mean() == 0:
                block = channels[0][y:y + block_size, x:x + block_size]
                dct_block = cv2.dct(block.astype(np.float32) - 128)
                quantized_block = np.round(dct_block / quantization_level) * quantization_level
                idct_block = cv2.idct(quantized_block) + 128
                channels[0][y:y + block_size, x:x + block_size] = np.clip(idct_block, 0, 255)
                channels[1][y:y + block_size, x:x + block_size] = 128
                channels[2][y:y + block_size, x:x + block_size] = 128
    return cv2.cvtColor(cv2.merge(channels), cv2.COLOR_YCrCb2BGR)


def colour_round_trip_only(frame):
    """The cv2 calls around the Python block loop without the loop itself (frame_differencing.py:115-116,129-130): the
    "cv2-only stage baseline" of BASELINE.md section 3 (``degrade="colour_only"`` in the loops below)."""
    channels = list(cv2.split(cv2.cvtColor(frame, cv2.COLOR_BGR2YCrCb)))
    return cv2.cvtColor(cv2.merge(channels), cv2.COLOR_YCrCb2BGR)


def fd_loop(frames, block_size=4, motion_threshold=0.5, min_area=500, kernel_size=7, release_factor=0.5,
            quantization_level=100, literal_blocks=False, prev_gray=None, acc=None, degrade=True):
    """frame_differencing.py:67-133 on an in-memory clip.  ``frames[0]`` seeds prev_gray unless
    ``prev_gray``/``acc`` state is passed in (chunk hand-off), in which case every frame is processed.
    Returns dict of per-frame lists: raw, filtered, dilated, acc, overlay, compressed (+ final state)."""
    frames = list(frames)
    if prev_gray is None:
        prev_gray = first_frame_gray_fd(frames[0])
        frames = frames[1:]
    h, w = prev_gray.shape
    kernel = np.ones((kernel_size, kernel_size), np.uint8)
    if acc is None:
        acc = np.zeros((h, w), np.uint8)
    out = dict(raw=[], filtered=[], dilated=[], acc=[], overlay=[], compressed=[])
    for frame in frames:
        gray = _stable_blur(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), (5, 5), 0)
        diff = cv2.absdiff(prev_gray, gray)
        _, raw = cv2.threshold(diff, motion_threshold, 255, cv2.THRESH_BINARY)
        filtered = so.contour_filter_cv2(raw, min_area)
        dilated = cv2.dilate(filtered, kernel, iterations=1)
        acc = cv2.addWeighted(acc, release_factor, dilated, 1 - release_factor, 0)
        overlay = frame.copy()
        overlay[acc > 127] = [0, 0, 255]
        out["raw"].append(raw)
        out["filtered"].append(filtered)
        out["dilated"].append(dilated)
        out["acc"].append(acc)
        out["overlay"].append(overlay)
        if degrade == "colour_only":
            out["compressed"].append(colour_round_trip_only(frame))
        elif degrade:
            if literal_blocks:
                comp = literal_degrade_fd(frame, acc, block_size, quantization_level)
            else:
                comp = so.degrade_fd(frame, acc, block_size, quantization_level)
            out["compressed"].append(comp)
        prev_gray = gray
    out["state_prev_gray"] = prev_gray
    out["state_acc"] = acc
    return out


def morph_kernel_ellipse(k: int) -> np.ndarray:
    return cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))


def window_loop(frames, window_size=5, alpha_fraction=0.2, morph_kernel=2, morph_shape="ellipse", kernel_size=7,
                block_size=4, motion_threshold=0.5, quantization_level=100, literal_blocks=False,
                prev_gray=None, history=None, degrade=True):
    """The north_star loop (see module docstring).  ``frames[0]`` seeds prev_gray unless state is given.
    ``history`` is the list of the last (<= window_size-1... window_size) raw masks for chunk hand-off."""
    frames = list(frames)
    if prev_gray is None:
        prev_gray = cv2.cvtColor(frames[0], cv2.COLOR_BGR2GRAY)
        frames = frames[1:]
    q = deque(history or [], maxlen=window_size)
    mk = (morph_kernel_ellipse(morph_kernel) if morph_shape == "ellipse"
          else np.ones((morph_kernel, morph_kernel), np.uint8)) if morph_kernel > 0 else None
    dk = np.ones((kernel_size, kernel_size), np.uint8) if kernel_size > 0 else None
    out = dict(raw=[], voted=[], mask=[], overlay=[], compressed=[])
    for frame in frames:
        gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
        diff = cv2.absdiff(prev_gray, gray)
        _, raw = cv2.threshold(diff, motion_threshold, 255, cv2.THRESH_BINARY)
        q.append(raw)
        cumulative = np.sum(np.array(q), axis=0)
        voted = (cumulative >= (alpha_fraction * len(q) * 255)).astype(np.uint8) * 255
        m = voted
        if mk is not None:
            m = cv2.morphologyEx(m, cv2.MORPH_CLOSE, mk)
            m = cv2.morphologyEx(m, cv2.MORPH_OPEN, mk)
        if dk is not None:
            m = cv2.dilate(m, dk, iterations=1)
        overlay = frame.copy()
        overlay[m > 127] = [0, 0, 255]
        out["raw"].append(raw)
        out["voted"].append(voted)
        out["mask"].append(m)
        out["overlay"].append(overlay)
        if degrade == "colour_only":
            out["compressed"].append(colour_round_trip_only(frame))
        elif degrade:
            if literal_blocks:
                comp = literal_degrade_fd(frame, m, block_size, quantization_level)
            else:
                comp = so.degrade_fd(frame, m, block_size, quantization_level)
            out["compressed"].append(comp)
        prev_gray = gray
    out["state_prev_gray"] = prev_gray
    out["state_history"] = list(q)
    return out


def literal_degrade_mco(frame_in, frame_mask):
    """motion_compression_opt.py:152-183 kept as written (8x8 blocks, three channels, re-gray)."""
    QTY_aggressive = np.full((8, 8), 100, dtype=np.float32)
    frame_ycrcb = cv2.cvtColor(frame_in, cv2.COLOR_BGR2YCrCb)
    channels = list(cv2.split(frame_ycrcb))
    for i in range(0, frame_mask.shape[0], 8):
        for j in range(0, frame_mask.shape[1], 8):
            block_mask = frame_mask[i:i + 8, j:j + 8]
            if block_mask.size == 0 or block_mask.shape[0] < 8 or block_mask.shape[1] < 8:
                continue
            if block_mask.mean() == 0:
                for c in range(3):
                    block = channels[c][i:i + 8, j:j + 8]
                    if block.shape == (8, 8):
                        dct_block = cv2.dct(block.astype(np.float32) - 128)
                        quantized_block = np.round(dct_block / QTY_aggressive) * QTY_aggressive
                        idct_block = cv2.idct(quantized_block) + 128
                        channels[c][i:i + 8, j:j + 8] = np.clip(idct_block, 0, 255)
    frame_processed = cv2.cvtColor(cv2.merge(channels), cv2.COLOR_YCrCb2BGR)
    for i in range(0, frame_mask.shape[0], 8):
        for j in range(0, frame_mask.shape[1], 8):
            block_mask = frame_mask[i:i + 8, j:j + 8]
            if block_mask.size == 0 or block_mask.shape[0] < 8 or block_mask.shape[1] < 8:
                continue
            if block_mask.mean() == 0:
                roi = frame_processed[i:i + 8, j:j + 8]
                gray_roi = cv2.cvtColor(roi, cv2.COLOR_BGR2GRAY)
                frame_processed[i:i + 8, j:j + 8] = cv2.cvtColor(gray_roi, cv2.COLOR_GRAY2BGR)
    return frame_processed


def mco_compress(frames, masks, literal_blocks=False):
    """motion_compression_opt.py:141-185 on arrays: one output frame per (frame, mask) pair."""
    out = []
    for f, m in zip(frames, masks):
        if m.ndim == 3:
            m = cv2.cvtColor(m, cv2.COLOR_BGR2GRAY)
        out.append(literal_degrade_mco(f, m) if literal_blocks else so.degrade_mco(f, m))
    return out
