"""Host-side multi-GPU logic on CPU: stream sharding, frame chunks with a temporal halo (validated with the
oracle: chunked == unchunked), and the statistics all-reduce over gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dynamic_video_compression_surveillance_b200 import sharding
from dynamic_video_compression_surveillance_b200.synth import make_clip
from oracle import loops


def test_shard_streams_partition():
    for n, w in ((64, 8), (64, 1), (7, 4), (3, 8), (0, 2)):
        parts = [sharding.shard_streams(n, w, r) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert sharding.shard_streams(64, 8, 3) == list(range(24, 32))


def test_frame_chunks_cover_and_halo():
    chunks = sharding.frame_chunks(1800, 8, halo=5)
    assert chunks[0].start == 1 and chunks[-1].stop == 1800
    for a, b in zip(chunks, chunks[1:]):
        assert a.stop == b.start
        assert b.warm_start == b.start - 5
    assert chunks[0].warm_start == 1
    assert sharding.ema_exact_halo(0.5) == 14 and sharding.ema_exact_halo(0.3) == 6 and sharding.ema_exact_halo(0.7) is None


def test_window_loop_chunked_with_halo_equals_unchunked():
    """K-window mode: a halo of K frames rebuilds the state exactly."""
    n, K = 41, 5
    frames = make_clip((72, 96), n, seed=4).frames()
    cfg = dict(window_size=K, alpha_fraction=0.2, morph_kernel=2, kernel_size=7, degrade=False)
    whole = loops.window_loop(list(frames), **cfg)["mask"]
    for ch in sharding.frame_chunks(n, 3, halo=K):
        part = loops.window_loop(list(frames[ch.warm_start - 1:ch.stop]), **cfg)["mask"]      # frame warm_start-1 seeds
        keep = part[ch.start - ch.warm_start:]
        for t, m in zip(range(ch.start, ch.stop), keep):
            assert np.array_equal(m, whole[t - 1]), (ch, t)


def test_fd_loop_chunked_with_ema_halo_equals_unchunked():
    """EMA mode, release_factor 0.5: 14 warm-up frames re-converge the uint8 accumulator exactly.  The chunk's seed
    frame goes through the 5x5 blur like every non-first frame of the stream (only the global first frame gets
    the heavy blur, frame_differencing.py:77)."""
    import cv2
    n = 60
    frames = make_clip((72, 96), n, seed=6).frames()
    whole = loops.fd_loop(list(frames), degrade=False)["acc"]
    halo = sharding.ema_exact_halo(0.5)
    for ch in sharding.frame_chunks(n, 2, halo=halo)[1:]:
        seed = loops._stable_blur(cv2.cvtColor(frames[ch.warm_start - 1], cv2.COLOR_BGR2GRAY), (5, 5), 0)
        part = loops.fd_loop(list(frames[ch.warm_start:ch.stop]), prev_gray=seed, acc=np.zeros((72, 96), np.uint8),
                             degrade=False)["acc"]
        keep = part[ch.start - ch.warm_start:]
        for t, m in zip(range(ch.start, ch.stop), keep):
            assert np.array_equal(m, whole[t - 1]), t


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    streams = sharding.shard_streams(5, world, rank)
    frames = sum(10 + s for s in streams)
    local = dict(frames=frames, pixels=frames * 100, motion_pixels=frames * 7, blocks=frames * 25, static_blocks=frames * 20)
    total = sharding.reduce_counters(local)
    q.put((rank, streams, total))
    dist.destroy_process_group()


def test_counter_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_streams = sorted(sum((r[1] for r in res), []))
    assert all_streams == list(range(5))
    frames = sum(10 + s for s in range(5))
    for _, _, total in res:
        assert total == dict(frames=frames, pixels=frames * 100, motion_pixels=frames * 7, blocks=frames * 25,
                             static_blocks=frames * 20)
    assert sharding.motion_percentage(res[0][2]) == pytest.approx(7.0)
    assert sharding.static_block_percentage(res[0][2]) == pytest.approx(80.0)


def test_shard_streams_weighted():
    from dynamic_video_compression_surveillance_b200.sharding import shard_streams_weighted
    sh = shard_streams_weighted(64, [792, 792, 792, 794, 1195, 1203, 1197, 1196])      # the per-GPU end-to-end rates measured on an 8-GPU box
    assert [len(x) for x in sh] == [6, 6, 6, 6, 10, 10, 10, 10]
    assert [i for part in sh for i in part] == list(range(64))
    assert [len(x) for x in shard_streams_weighted(64, [1.0] * 8)] == [8] * 8
    assert [len(x) for x in shard_streams_weighted(5, [1, 0, 1])] == [3, 0, 2]
    assert [len(x) for x in shard_streams_weighted(7, [3, 1])] == [5, 2]
