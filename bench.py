#!/usr/bin/env python
"""Benchmark of the per-frame hot loop (BASELINE.json metric: 1080p frames/s, HBM roofline fraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One process per GPU (torchrun for N > 1).  A step is one pass of the loop over the whole synthetic clip
(config #1 of BASELINE.json: one 1080p stream, 1800 frames, K=5 window vote; every rank runs its own stream, so
per-GPU work is fixed: weak scaling).  `value` is timed with the clip and the output buffers resident in HBM
(11.2 GB in, 22.4 GB out, far larger than L2); `e2e` is timed through the host-buffer C-ABI call
(dvc_process_host) with pinned host frames, H2D and D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "1080p_frames_per_second"
UNIT = "frames/s"
LOOP = dict(window_size=5, alpha_fraction=0.2, morph_kernel=2, morph_shape="ellipse", kernel_size=7, block_size=4,
            motion_threshold=0.5, quantization_level=100)
# bytes per pixel per frame
K4_ALG_BYTES_PER_PX = 3 + 3 + 3 + 1.0 / 8          # read BGR, write compressed + overlay, read the mask bit-plane
SURVEY_LOOP_BYTES_PER_PX = 28                      # SURVEY.md section 8(d) accounting of the north_star loop (uint8 masks)
ACTUAL_LOOP_BYTES_PER_PX = 3 + 1.0 / 8 + K4_ALG_BYTES_PER_PX + 4 * (1.0 / 8)   # this design: K1 + mask planes + K4


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--frames", type=int, default=1800)
    p.add_argument("--resolution", default="1080p")
    p.add_argument("--max-batch", type=int, default=int(os.environ.get("DVC_BENCH_BATCH", "225")))
    p.add_argument("--e2e-frames", type=int, default=256)
    p.add_argument("--mode", default="window", choices=["window", "fd"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-overlap", action="store_true", help="strict per-batch stream order instead of the two-stream pipeline")
    # overrides of the loop parameters for the other BASELINE configs (e.g. config 3: --resolution 4k --frames 900
    # --kernel-size 15 --morph-kernel 15 --morph-shape rect); the default line is config 2
    p.add_argument("--kernel-size", type=int, default=None)
    p.add_argument("--morph-kernel", type=int, default=None)
    p.add_argument("--morph-shape", default=None, choices=["ellipse", "rect"])
    p.add_argument("--window-size", type=int, default=None)
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def mark(self):
        """Start of the timed region: samples before this wall-clock instant are ignored."""
        import datetime
        self.t_mark = datetime.datetime.now()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        self.power_max = None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                import datetime
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                if getattr(self, "t_mark", None) and ts < self.t_mark:
                    continue
                sm.append(float(f[1])); mx.append(float(f[2]))
                self.power_max = max(self.power_max or 0.0, float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sm_mhz_min": min(sm) if sm else None,
                "power_w_max": self.power_max}


# ---------------------------------------------------------------------------------------------------
# synthetic clip
# ---------------------------------------------------------------------------------------------------
def device_clip(clip, n_frames, device):
    """Render the synthetic clip directly into HBM (same bytes as clip.frames(): static background +
    rectangles pasted in order)."""
    import torch
    bg = torch.from_numpy(clip.background).to(device)
    out = torch.empty((n_frames,) + tuple(bg.shape), dtype=torch.uint8, device=device)
    out[:] = bg
    for t in range(n_frames):
        for (x, y, w, h), rect in zip(clip.rect_positions(t), clip.rects):
            out[t, y:y + h, x:x + w] = torch.tensor(rect[5], dtype=torch.uint8, device=device)
    return out


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the same loop (literal Python block loop, as the reference runs it)
# ---------------------------------------------------------------------------------------------------
def cpu_loop_fps(clip, n_frames: int, mode: str, state=None):
    """Time the CPU port on frames 1..n_frames (frame 0 seeds prev_gray); returns (fps, seconds, state)."""
    import cv2
    from oracle import loops
    frames = [clip[t] for t in range(n_frames + 1)] if state is None else [clip[t] for t in state["next"]]
    t0 = time.perf_counter()
    if mode == "window":
        kw = dict(LOOP); kw["literal_blocks"] = True
        if state is None:
            r = loops.window_loop(frames, **kw)
        else:
            r = loops.window_loop(frames, prev_gray=state["prev_gray"], history=state["history"], **kw)
        st = dict(prev_gray=r["state_prev_gray"], history=r["state_history"])
    else:
        kw = dict(block_size=LOOP["block_size"], kernel_size=LOOP["kernel_size"], literal_blocks=True)
        if state is None:
            r = loops.fd_loop(frames, **kw)
        else:
            r = loops.fd_loop(frames, prev_gray=state["prev_gray"], acc=state["acc"], **kw)
        st = dict(prev_gray=r["state_prev_gray"], acc=r["state_acc"])
    dt = time.perf_counter() - t0
    n = len(r["compressed"])
    return n / dt, dt, st, cv2.getNumThreads()


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the loop on this box's host cores.  The reference is
    pure Python + cv2 and cannot travel as a compiled artefact, so this runs the oracle port (same cv2 calls, same
    Python block loop: oracle/loops.py), one frame of the same 1080p workload per step."""
    if rank != 0:
        return
    import cv2
    from dynamic_video_compression_surveillance_b200.synth import make_clip, RESOLUTIONS
    h, w = RESOLUTIONS[args.resolution]
    clip = make_clip(args.resolution, args.frames, seed=0)
    fps0, _, state, threads = cpu_loop_fps(clip, 1, args.mode)              # seeds state (counts as warm-up 0)
    t_next = 2
    for _ in range(max(0, args.warmup - 1)):
        state["next"] = [t_next]; t_next += 1
        _, _, state, _ = cpu_loop_fps(clip, 0, args.mode, state)
    times = []
    for _ in range(args.steps):
        state["next"] = [t_next]; t_next += 1
        _, dt, state, _ = cpu_loop_fps(clip, 0, args.mode, state)
        times.append(dt)
    total = sum(times)
    fps = args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, h, w) | {"sample": "1 frame of the clip per step"},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.steps} consecutive 1080p frames, one per step; oracle/loops.py literal block loop; "
                                       f"cv2 {cv2.__version__} threads={threads}, os.cpu_count()={os.cpu_count()}"},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, h, w):
    name = {"window": f"north_star loop: gray/absdiff/threshold -> K={LOOP['window_size']} window vote -> {LOOP['morph_shape']}-"
                      f"{LOOP['morph_kernel']} close/open -> {LOOP['kernel_size']}x{LOOP['kernel_size']} dilate -> "
                      "overlay + 4x4 block-DCT degrade",
            "fd": "frame_differencing.py loop: gray/blur5/absdiff/threshold -> contour filter -> 7x7 dilate -> EMA -> overlay + "
                  "4x4 block-DCT degrade"}[args.mode]
    cfg_name = "BASELINE configs[1]" if (args.resolution, args.frames) == ("1080p", 1800) else "BASELINE-style config"
    return {"workload": f"{cfg_name}: single {args.resolution} ({w}x{h}) synthetic stream per GPU, {args.frames} frames, "
                        f"{name}", "mode": args.mode, "frames_per_step": args.frames, "max_batch": args.max_batch, "two_stream_overlap": not args.no_overlap,
            "l2_policy": f"inputs (clip {args.frames * h * w * 3 / 1e9:.1f} GB) and outputs ({2 * args.frames * h * w * 3 / 1e9:.1f} GB) per step "
                         "are far larger than the 126 MB L2"}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from dynamic_video_compression_surveillance_b200 import pipeline as P
    from dynamic_video_compression_surveillance_b200.synth import make_clip, RESOLUTIONS

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h, w = RESOLUTIONS[args.resolution]
    n, B = args.frames, args.max_batch
    clip = make_clip(args.resolution, n + 1, seed=rank)            # stream id = rank
    frames = device_clip(clip, n + 1, dev)
    ov = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
    cp = torch.empty_like(ov)
    kw = dict(LOOP)
    if args.mode == "fd":
        for k in ("window_size", "alpha_fraction", "morph_kernel", "morph_shape"):
            kw.pop(k)
    pipe = P.FramePipeline(w, h, args.mode, max_batch=B, device=local_rank, **kw)
    if args.mode == "window":
        seed_gray = P.bgr2gray(frames[:1])[0].cpu().numpy()
    else:
        import cv2
        seed_gray = cv2.GaussianBlur(P.bgr2gray(frames[:1])[0].cpu().numpy(), (25, 25), 30)   # stays on the host (fd:77)
    pipe.begin_stream(seed_gray)
    pipe.set_overlap(not args.no_overlap)
    body = frames[1:]

    def one_pass():
        for i in range(0, n, B):
            pipe.process_device(body[i:i + B], ov[i:i + B], cp[i:i + B])
        pipe.flush()          # the current stream re-joins the library's internal streams

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                      # started early: nvidia-smi's own start-up must not fall into the timed region
    for _ in range(max(3, args.warmup)):
        one_pass()
    barrier()
    launches0 = pipe.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    e0.record()
    for _ in range(args.steps):
        one_pass()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = pipe.launch_count() - launches0
    # second timed region, same K steps, strict stream order (kernels serialised) with CUDA events around every kernel
    # group on its launch stream: per-kernel times for the roofline are not inflated by concurrently running kernels
    pipe.set_overlap(False)
    one_pass()
    barrier()
    pipe.profile(True)
    pipe.profile_read()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        one_pass()
    p1.record()
    barrier()
    ms_serial = p0.elapsed_time(p1)
    prof = pipe.profile_read()
    pipe.profile(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([v for v in pipe.counters().values()], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)              # the only collective: final statistics
    ms = float(t.item())
    total_frames = args.steps * n * world
    value = total_frames / (ms / 1e3)

    # ---- roofline of the dominant kernel (K4) from the live CUDA-event times -------------------------
    peak, peak_src = peaks()
    px = h * w
    k4_ms, k4_launches = prof["degrade"]
    frames_per_launch = n * args.steps / max(1, k4_launches)
    k4_bytes = K4_ALG_BYTES_PER_PX * px * frames_per_launch
    k4_gbs = k4_bytes / (k4_ms / max(1, k4_launches) * 1e-3) / 1e9 if k4_ms > 0 else 0.0
    kernel_ms = {k: v[0] for k, v in prof.items() if v[1]}
    roofline = {"bound": "hbm", "kernel": "k_degrade4s (K4, persistent TMA ring: overlay + YCrCb + 4x4 DCT degrade + statistics)",
                "achieved": k4_gbs, "peak": peak, "unit": "GB/s", "frac": k4_gbs / peak, "peak_source": peak_src,
                "traffic": None, "alg_bytes_per_px": K4_ALG_BYTES_PER_PX, "frames_per_launch": frames_per_launch,
                "avg_launch_ms": k4_ms / max(1, k4_launches), "kernel_share_of_step": k4_ms / ms_serial if ms_serial else None,
                "timed_region": "second pass of the same K steps in strict stream order (no two-stream overlap) with CUDA events "
                                "around each kernel group on its launch stream",
                "serialised_fps_per_gpu": n * args.steps / (ms_serial / 1e3),
                "kernel_ms_in_timed_region": kernel_ms,
                "loop": {"fps_per_gpu": value / world,
                         "survey_accounting": {"bytes_per_px": SURVEY_LOOP_BYTES_PER_PX,
                                               "frac": (value / world) * px * SURVEY_LOOP_BYTES_PER_PX / 1e9 / peak},
                         "this_design": {"bytes_per_px": ACTUAL_LOOP_BYTES_PER_PX,
                                         "frac": (value / world) * px * ACTUAL_LOOP_BYTES_PER_PX / 1e9 / peak}}}
    # K4 writes twice what it reads, and HBM absorbs writes more slowly than a read/write mix: report the kernel's write rate
    # next to a write-only (fill) rate measured here, as context for `frac` (which stays against the copy peak)
    try:
        fill_buf = torch.empty(1 << 31, dtype=torch.uint8, device=dev)
        best = None
        for _ in range(6):
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(); fill_buf.zero_(); f1.record(); torch.cuda.synchronize()
            t_ms = f0.elapsed_time(f1)
            best = t_ms if best is None else min(best, t_ms)
        fill_gbs = fill_buf.numel() / (best * 1e-3) / 1e9
        del fill_buf
        k4_write_gbs = 6.0 * px * frames_per_launch / (k4_ms / max(1, k4_launches) * 1e-3) / 1e9 if k4_ms > 0 else 0.0
        roofline["write_path"] = {"written_bytes_per_px": 6.0, "achieved_write": k4_write_gbs, "fill_peak_measured_here": fill_gbs,
                                  "unit": "GB/s", "frac_of_fill_peak": k4_write_gbs / fill_gbs if fill_gbs else None}
    except Exception:
        pass
    traffic_file = os.path.join(ROOT, "profiles", "k4_dram_bytes_per_launch.json")
    if os.path.exists(traffic_file):
        try:
            tf = json.load(open(traffic_file))
            roofline["traffic"] = tf["dram_bytes_per_frame"] * frames_per_launch
            roofline["traffic_source"] = tf.get("source")
        except Exception:
            pass

    # ---- end to end through the host-buffer call ------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        ne = min(args.e2e_frames, n)
        hin = P.pinned_empty((ne, h, w, 3)); hov = P.pinned_empty((ne, h, w, 3)); hcp = P.pinned_empty((ne, h, w, 3))
        hin.copy_(body[:ne].cpu())
        for _ in range(3):
            pipe.process_host(hin, hov, hcp)
        barrier()
        t0 = time.perf_counter()
        reps = max(1, args.steps)
        for _ in range(reps):
            pipe.process_host(hin, hov, hcp)          # returns when the outputs are in host memory
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": reps * ne * world / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": ne * px * 3,
               "d2h_bytes_per_step": 2 * ne * px * 3, "frames_per_step": ne,
               "note": "dvc_process_host: pinned host frames in, overlay + compressed frames out, chunks of max_batch "
                       "double-buffered on copy/compute streams; wall clock around the blocking call"}
        del hin, hov, hcp

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        nf = 3
        fps, dt, _, threads = cpu_loop_fps(clip, nf, args.mode)
        import cv2
        cpu_baseline = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{nf} frames of the same 1080p clip ({dt:.1f} s); oracle/loops.py with the reference's literal "
                                  f"Python block loop; cv2 {cv2.__version__} threads={threads}, os.cpu_count()={os.cpu_count()}"}
    if rank == 0:
        names = list(pipe.counters().keys())
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": workload_config(args, h, w), "clocks": clocks,
                "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "statistics": dict(zip(names, [int(v) for v in cnt.tolist()]))}
        print(json.dumps(line))
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    for key, val in (("kernel_size", args.kernel_size), ("morph_kernel", args.morph_kernel), ("morph_shape", args.morph_shape),
                     ("window_size", args.window_size)):
        if val is not None:
            LOOP[key] = val
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
