"""Array-backed stand-in for the ``cv2`` module seen by the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  ``run_reference_fd`` imports
/root/reference/frame_differencing.py, rebinds that module's ``cv2`` global to a
proxy whose VideoCapture / VideoWriter read and write numpy arrays (the mp4v
codec is lossy and would destroy bit-exactness) and whose
``threshold`` / ``dilate`` / ``addWeighted`` tap the intermediate masks.  Every
other attribute forwards to the real cv2, so all arithmetic is the reference's.
Only usable where /root/reference exists (the build container); its outputs are
committed as fixtures by oracle/make_golden.py.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile

import cv2 as _cv2
import numpy as np

REFERENCE_DIR = os.environ.get("DVC_REFERENCE_DIR", "/root/reference")


class _FakeCapture:
    def __init__(self, store, path):
        self._frames = store.get(path)
        self._i = 0
        self._fps = store.get(("fps", path), 30.0)

    def isOpened(self):
        return self._frames is not None

    def get(self, prop):
        if prop == _cv2.CAP_PROP_FPS:
            return float(self._fps)
        if prop == _cv2.CAP_PROP_FRAME_WIDTH:
            return float(self._frames[0].shape[1])
        if prop == _cv2.CAP_PROP_FRAME_HEIGHT:
            return float(self._frames[0].shape[0])
        if prop == _cv2.CAP_PROP_FRAME_COUNT:
            return float(len(self._frames))
        return 0.0

    def read(self):
        if self._frames is None or self._i >= len(self._frames):
            return False, None
        f = np.array(self._frames[self._i], copy=True)
        self._i += 1
        return True, f

    def release(self):
        pass


class _FakeWriter:
    def __init__(self, store, path, fourcc, fps, size, isColor=True):
        self._out = store.setdefault(path, [])
        del self._out[:]
        store[("fps", path)] = fps

    def write(self, frame):
        self._out.append(np.array(frame, copy=True))

    def release(self):
        pass

    def isOpened(self):
        return True


class Cv2Proxy:
    """Forwards to the real cv2; swaps file I/O for arrays and records taps."""

    def __init__(self, store, taps=None):
        self._store = store
        self._taps = taps if taps is not None else {}

    def __getattr__(self, name):
        return getattr(_cv2, name)

    def VideoCapture(self, path):
        return _FakeCapture(self._store, path)

    def VideoWriter(self, path, fourcc, fps, size, isColor=True):
        return _FakeWriter(self._store, path, fourcc, fps, size, isColor)

    def _tap(self, key, value):
        self._taps.setdefault(key, []).append(np.array(value, copy=True))

    def threshold(self, src, thresh, maxval, typ):
        r = _cv2.threshold(src, thresh, maxval, typ)
        self._tap("threshold", r[1])
        return r

    def dilate(self, src, kernel, iterations=1):
        r = _cv2.dilate(src, kernel, iterations=iterations)
        self._tap("dilate_in", src)
        self._tap("dilate", r)
        return r

    def addWeighted(self, a, alpha, b, beta, gamma):
        r = _cv2.addWeighted(a, alpha, b, beta, gamma)
        self._tap("addWeighted", r)
        return r

    # taps of temporal_smoothing_flow (motion_compression_opt.py:82,89-90)
    def cartToPolar(self, x, y, *a, **k):
        r = _cv2.cartToPolar(x, y, *a, **k)
        self._tap("flow_magnitude", r[0])
        return r

    def morphologyEx(self, src, op, kernel, *a, **k):
        r = _cv2.morphologyEx(src, op, kernel, *a, **k)
        self._tap("morph_in_%d" % op, src)
        self._tap("morph_out_%d" % op, r)
        return r


def _load_reference_module(name):
    path = os.path.join(REFERENCE_DIR, name + ".py")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path}: the reference is only present in the build container")
    spec = importlib.util.spec_from_file_location("_dvc_ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_reference_fd(frames, **kwargs):
    """Run the reference's ``filter_and_dilate_movements`` (frame_differencing.py:21) unmodified on an
    in-memory clip.  Returns dict(overlay=[...], compressed=[...], raw=[...], filtered=[...],
    dilated=[...], acc=[...]) with one entry per processed frame (len(frames) - 1)."""
    import logging
    mod = _load_reference_module("frame_differencing")
    store, taps = {}, {}
    mod.cv2 = Cv2Proxy(store, taps)
    with tempfile.TemporaryDirectory() as tmp:
        store["clip.mp4"] = frames
        root = logging.getLogger()
        before = list(root.handlers)
        try:
            mod.filter_and_dilate_movements("clip.mp4", tmp, **kwargs)
        finally:
            for h in list(root.handlers):
                if h not in before:
                    root.removeHandler(h)
                    h.close()
        out_dir = os.path.join(tmp, "clip")
        res = dict(overlay=store[os.path.join(out_dir, "dilated_motion_mask_video.mp4")],
                   compressed=store[os.path.join(out_dir, "compressed_final_video.mp4")],
                   raw=taps.get("threshold", []), filtered=taps.get("dilate_in", []),
                   dilated=taps.get("dilate", []), acc=taps.get("addWeighted", []))
        with open(os.path.join(out_dir, "execution_times.txt")) as f:
            res["execution_times"] = f.read()
    return res


def run_reference_mco_compress(frames, masks):
    """Run the reference's ``compress_with_motion`` (motion_compression_opt.py:111) unmodified on an
    in-memory clip and per-frame uint8 masks ([H, W], any values).  Returns the output frames."""
    mod = _load_reference_module("motion_compression_opt")
    store = {}
    mod.cv2 = Cv2Proxy(store)
    store["in.mp4"] = frames
    store["mask.mp4"] = masks
    with tempfile.TemporaryDirectory() as tmp:
        mod.compress_with_motion("in.mp4", "mask.mp4", tmp)
        return store[os.path.join(tmp, "compressed.mp4")]


def run_reference_temporal_smoothing_flow(frames, **kwargs):
    """Run the reference's ``temporal_smoothing_flow`` (motion_compression_opt.py:29-109) unmodified on an in-memory clip.
    Farneback flow runs as the reference calls it; the proxy taps what flows between the statements the GPU path
    replaces: the flow magnitude (cartToPolar, :82; ``mag > flow_threshold`` is the raw mask, :83), the voted mask going
    into MORPH_CLOSE (:86 -> :89), the mask leaving MORPH_OPEN (:90), and the rectangle mask handed to the mask writer
    (:93-98).  Returns dict(raw=[...], voted=[...], morphed=[...], rect=[...], overlay=[...], result=(frames, total, avg))."""
    mod = _load_reference_module("motion_compression_opt")
    store, taps = {}, {}
    mod.cv2 = Cv2Proxy(store, taps)
    store["in.mp4"] = frames
    thr = kwargs.get("flow_threshold", 0.5)
    with tempfile.TemporaryDirectory() as tmp:
        result = mod.temporal_smoothing_flow("in.mp4", tmp, **kwargs)
        rect = store[os.path.join(tmp, kwargs.get("mask_save_name", "mask.mp4"))]
        overlay = store[os.path.join(tmp, kwargs.get("save_name", "overlay.mp4"))]
    raw = [(m > thr).astype(np.uint8) * 255 for m in taps.get("flow_magnitude", [])]
    return dict(raw=raw, voted=taps.get("morph_in_%d" % _cv2.MORPH_CLOSE, []), morphed=taps.get("morph_out_%d" % _cv2.MORPH_OPEN, []),
                rect=rect, overlay=overlay, result=result)


def run_reference_window_vote(raw_masks, alpha_fraction, window_size, morph_kernel):
    """The reference's window vote + close/open lines (motion_compression_opt.py:61-62,83-90) are
    inside a function that also runs Farneback flow, so they cannot be called in isolation; this
    replays exactly those statements on given 0/255 masks using the reference's own objects
    (collections.deque, np.sum, cv2.morphologyEx)."""
    from collections import deque
    q = deque(maxlen=window_size)
    kernel = _cv2.getStructuringElement(_cv2.MORPH_ELLIPSE, (morph_kernel, morph_kernel))
    out = []
    for m in raw_masks:
        q.append(m)
        cumulative = np.sum(np.array(q), axis=0)
        sm = (cumulative >= (alpha_fraction * len(q) * 255)).astype(np.uint8) * 255
        sm = _cv2.morphologyEx(sm, _cv2.MORPH_CLOSE, kernel)
        sm = _cv2.morphologyEx(sm, _cv2.MORPH_OPEN, kernel)
        out.append(sm)
    return out
