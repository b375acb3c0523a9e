"""Drop-in for the reference module of the same name: ``windows.py`` imports ``process_single_video_fd`` from
here (windows.py:14,154).  Names, signatures, defaults, output files, log lines, ``execution_times.txt`` and the
never-raise error convention follow frame_differencing.py:7-196; the per-frame loop (:85-133) runs on the GPU
(dynamic_video_compression_surveillance_b200.host_loop.run_fd_stream -> C ABI).
"""
import logging
import os
import sys
import time

import cv2

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from dynamic_video_compression_surveillance_b200 import host_loop as _hl  # noqa: E402

_LOG_FORMAT = "%(asctime)s - %(levelname)s - %(message)s"


def setup_logging(output_dir):
    """frame_differencing.py:7-19: make the folder, basicConfig with a processing.log file handler + stream handler."""
    os.makedirs(output_dir, exist_ok=True)
    log_file = os.path.join(output_dir, "processing.log")
    logging.basicConfig(level=logging.INFO, format=_LOG_FORMAT,
                        handlers=[logging.FileHandler(log_file, mode="w"), logging.StreamHandler()])
    logging.info(f"Logging configured. Log file saved in: {log_file}")


def filter_and_dilate_movements(video_path, output_dir, block_size=4, search_area=16, motion_threshold=0.5,
                                min_area=500, kernel_size=7, release_factor=0.5, quantization_level=100,
                                scale_factor=1.0, progress_callback=None, *, max_batch=16, device=None, stats_out=None):
    """frame_differencing.py:21-159.  ``search_area`` is accepted and ignored, exactly as there.  Keyword-only
    extras: ``max_batch`` (frames per GPU batch), ``device``, ``stats_out`` (dict that receives the GPU counters)."""
    t_start = time.time()
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        logging.error("Unable to open the video.")
        return
    out_dir = os.path.join(output_dir, _hl.video_stem(video_path))
    os.makedirs(out_dir, exist_ok=True)
    setup_logging(out_dir)

    fps = int(cap.get(cv2.CAP_PROP_FPS))
    size = (int(int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)) * scale_factor), int(int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)) * scale_factor))
    fourcc = cv2.VideoWriter_fourcc(*_hl.FOURCC)
    sinks = (cv2.VideoWriter(os.path.join(out_dir, _hl.FD_OVERLAY_NAME), fourcc, fps, size),
             cv2.VideoWriter(os.path.join(out_dir, _hl.FD_COMPRESSED_NAME), fourcc, fps, size))
    ok, first = cap.read()
    if not ok:
        logging.error("Unable to read the first frame of the video.")
        cap.release()
        return

    run = _hl.FdRun()
    params = dict(block_size=block_size, motion_threshold=motion_threshold, min_area=min_area, kernel_size=kernel_size,
                  release_factor=release_factor, quantization_level=quantization_level)
    try:
        run = _hl.run_fd_stream(cap, first, sinks, size, params, max_batch, device, progress_callback)
    except Exception as e:                                   # frame_differencing.py:140-141: log, do not raise
        logging.error("Error during processing: " + str(e), exc_info=True)
    finally:
        cap.release()
        for s in sinks:
            s.release()
    if stats_out is not None:
        stats_out.update(run.counters)
    if run.counters:
        _hl.write_gpu_statistics(out_dir, run.counters)

    total = time.time() - t_start
    avg = sum(run.per_frame_s) / len(run.per_frame_s) if run.per_frame_s else 0
    times_path = os.path.join(out_dir, _hl.TIMES_NAME)
    _hl.write_execution_times(times_path, [_hl.StageTiming("Frame Differencing", run.frames, total, avg)], total)
    logging.info(f"Execution statistics saved in: {times_path}")


def process_single_video_fd(video_path, output_dir, block_size=4, search_area=16, motion_threshold=0.5, min_area=500,
                            kernel_size=7, release_factor=0.5, quantization_level=100, scale_factor=1.0,
                            progress_callback=None):
    """frame_differencing.py:161-196: per-video folder, logging, banner lines, then the processing."""
    name = _hl.video_stem(video_path)
    out_dir = os.path.join(output_dir, name)
    os.makedirs(out_dir, exist_ok=True)
    setup_logging(out_dir)
    logging.info(f"=== Start processing (Frame Differencing) for '{name}' ===")
    filter_and_dilate_movements(video_path, output_dir, block_size=block_size, search_area=search_area,
                                motion_threshold=motion_threshold, min_area=min_area, kernel_size=kernel_size,
                                release_factor=release_factor, quantization_level=quantization_level,
                                scale_factor=scale_factor, progress_callback=progress_callback)
    logging.info(f"=== Processing successfully completed for '{name}'. ===")
