"""Pins oracle/loops.py to fixtures produced by the UNMODIFIED reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import loops, stage_ops as so
from tests.golden_util import FD_CLIPPED_FIXTURE, FD_FIXTURES, GOLDEN, OF_FIXTURES, load_fd, load_of, sha, unpack


@pytest.mark.parametrize("name", FD_FIXTURES + [FD_CLIPPED_FIXTURE])
def test_fd_loop_matches_reference_fixture(name):
    z, frames, kw, (h, w, n) = load_fd(name)
    got = loops.fd_loop(list(frames), **kw)
    for key in ("raw", "filtered", "dilated"):
        ref = unpack(z[key], w)
        assert len(got[key]) == n - 1
        for t in range(n - 1):
            assert np.array_equal(got[key][t], ref[t]), (key, t)
    assert np.array_equal(np.stack(got["acc"]), z["acc"])
    assert [sha(x) for x in got["overlay"]] == list(z["overlay_sha"])
    assert [sha(x) for x in got["compressed"]] == list(z["compressed_sha"])
    assert np.array_equal(np.stack(got["compressed"][-2:]), z["compressed_tail"])


@pytest.mark.parametrize("name", FD_FIXTURES[:2])
def test_fd_numpy_stage_ops_match_reference_fixture(name):
    """Same fixture through the pure-numpy restatements (no cv2 for the integer ops)."""
    z, frames, kw, (h, w, n) = load_fd(name)
    bs, q = kw.get("block_size", 4), kw.get("quantization_level", 100)
    k, rf = kw.get("kernel_size", 7), kw.get("release_factor", 0.5)
    thr, min_area = kw.get("motion_threshold", 0.5), kw.get("min_area", 500)
    prev = loops.first_frame_gray_fd(frames[0])
    acc = np.zeros((h, w), np.uint8)
    for t in range(1, n):
        gray = so.gaussian_blur5(so.bgr2gray(frames[t]))
        raw = so.threshold_binary(so.absdiff(prev, gray), thr)
        assert np.array_equal(raw, unpack(z["raw"], w)[t - 1])
        filt = so.contour_filter(raw, min_area)
        assert np.array_equal(filt, unpack(z["filtered"], w)[t - 1])
        dil = so.dilate(filt, so.structuring_rect(k))
        assert np.array_equal(dil, unpack(z["dilated"], w)[t - 1])
        acc = so.add_weighted(acc, rf, dil, 1 - rf)
        assert np.array_equal(acc, z["acc"][t - 1])
        assert sha(so.overlay_paint(frames[t], acc)) == str(z["overlay_sha"][t - 1])
        assert sha(so.degrade_fd(frames[t], acc, bs, q)) == str(z["compressed_sha"][t - 1])
        prev = gray


def test_literal_block_loop_equals_vectorised():
    z, frames, kw, _ = load_fd(FD_FIXTURES[0])
    a = loops.fd_loop(list(frames[:14]), literal_blocks=True, **kw)
    b = loops.fd_loop(list(frames[:14]), literal_blocks=False, **kw)
    assert all(np.array_equal(x, y) for x, y in zip(a["compressed"], b["compressed"]))


def test_mco_compress_matches_reference_fixture():
    from dynamic_video_compression_surveillance_b200.synth import make_clip
    z = np.load(os.path.join(GOLDEN, "mco_compress_64x96.npz"))
    h, w, n, seed = (int(v) for v in z["recipe"])
    frames = make_clip((h, w), n, seed=seed).frames()
    for lit in (False, True):
        out = loops.mco_compress(list(frames), list(z["masks"]), literal_blocks=lit)
        assert np.array_equal(np.stack(out), z["out"]), lit


def test_window_vote_matches_reference_fixture():
    z = np.load(os.path.join(GOLDEN, "window_vote_48x80.npz"))
    h, w, n, seed = (int(v) for v in z["recipe"])
    raws = unpack(z["raws"], w)
    for key in z.files:
        if not key.startswith("a"):
            continue
        alpha, K, mk = key[1:].split("_")
        alpha, K, mk = float(alpha), int(K[1:]), int(mk[1:])
        ref = unpack(z[key], w)
        mc = so.window_min_counts(alpha, K)
        kernel = so.structuring_ellipse(mk)
        for t in range(n):
            win = raws[max(0, t - K + 1):t + 1]
            cnt = (win != 0).sum(axis=0)
            voted = np.where(cnt >= mc[len(win) - 1], 255, 0).astype(np.uint8)
            got = so.morph_open(so.morph_close(voted, kernel), kernel)
            assert np.array_equal(got, ref[t]), (key, t)


@pytest.mark.parametrize("name", OF_FIXTURES)
def test_of_mask_chain_matches_unmodified_temporal_smoothing_flow(name):
    """motion_compression_opt.py:83-97 as run by the unmodified function (Farneback included): the numpy restatements of
    the window vote, close / open and contours -> rectangles reproduce every tapped mask."""
    m, kw, (h, w, n) = load_of(name)
    K, alpha = kw["window_size"], kw["alpha_fraction"]
    mc = so.window_min_counts(alpha, K)
    kernel = so.structuring_ellipse(kw["morph_kernel"])
    for t in range(n - 1):
        win = m["raw"][max(0, t - K + 1):t + 1]
        voted = np.where((win != 0).sum(axis=0) >= mc[len(win) - 1], 255, 0).astype(np.uint8)
        assert np.array_equal(voted, m["voted"][t]), ("voted", t)
        assert np.array_equal(so.window_vote(list(win), alpha), m["voted"][t])
        morphed = so.morph_open(so.morph_close(voted, kernel), kernel)
        assert np.array_equal(morphed, m["morphed"][t]), ("morphed", t)
        assert np.array_equal(so.mask_rectangles(morphed), m["rect"][t]), ("rect", t)


def test_config1_fixture_is_present_and_self_consistent():
    """BASELINE configs[0] (640x480, 300 frames, fd defaults) through the unmodified reference: hashes only; the oracle
    loop reproduces the first frames here, the GPU test (tests/test_gpu_parity.py) checks all 299."""
    z = np.load(os.path.join(GOLDEN, "fd_config1_480x640x300.npz"))
    h, w, n, seed, noise = (int(v) for v in z["recipe"])
    assert (h, w, n) == (480, 640, 300) and len(z["compressed_sha"]) == n - 1
    from dynamic_video_compression_surveillance_b200.synth import make_clip
    frames = make_clip((h, w), 4, seed=seed).frames()
    got = loops.fd_loop(list(frames))
    assert [sha(x) for x in got["acc"]] == list(z["acc_sha"][:3])
    assert [sha(x) for x in got["overlay"]] == list(z["overlay_sha"][:3])
    assert [sha(x) for x in got["compressed"]] == list(z["compressed_sha"][:3])
