"""The files the reference's GUI / performance_analysis.py read back: execution_times.txt must parse with the statements of
performance_analysis.py:38-107 (restated here: non-empty stripped lines, regex ``:\\s*([\\d\\.]+)``, the section headers)."""
import json
import os
import re

from dynamic_video_compression_surveillance_b200 import host_loop as hl

PATTERN = r":\s*([\d\.]+)"


def _parse(path):
    lines = [line.strip() for line in open(path) if line.strip() != ""]
    num = lambda s: re.search(PATTERN, s).group(1)
    total = [l for l in lines if l.startswith("Total video processing time:")]
    if lines[0].startswith("Motion Detection:"):
        ci = next(i for i, l in enumerate(lines) if l.startswith("Compression:"))
        return dict(md_frames=int(num(lines[1])), md_time=float(num(lines[2])), md_avg=float(num(lines[3])),
                    cp_frames=int(num(lines[ci + 1])), cp_time=float(num(lines[ci + 2])), cp_avg=float(num(lines[ci + 3])),
                    total=float(num(total[0])))
    assert lines[0].startswith("Frame Differencing:")
    return dict(md_frames=int(num(lines[1])), md_time=float(num(lines[2])), md_avg=float(num(lines[3])), total=float(num(total[0])))


def test_fd_execution_times_layout(tmp_path):
    p = str(tmp_path / hl.TIMES_NAME)
    hl.write_execution_times(p, [hl.StageTiming("Frame Differencing", 299, 1.9049, 0.00637)], 1.9049)
    text = open(p).read().splitlines()
    # frame_differencing.py:152-157, byte for byte
    assert text == ["Frame Differencing:", "  Frames processed: 299", "  Total time: 1.90 seconds",
                    "  Average time per frame: 0.0064 seconds", "", "Total video processing time: 1.90 seconds"]
    d = _parse(p)
    assert d == dict(md_frames=299, md_time=1.90, md_avg=0.0064, total=1.90)


def test_of_execution_times_layout(tmp_path):
    p = str(tmp_path / hl.TIMES_NAME)
    hl.write_execution_times(p, [hl.StageTiming("Motion Detection", 89, 12.345, 0.1387), hl.StageTiming("Compression", 89, 3.2, 0.03596)], 15.545)
    d = _parse(p)                                       # motion_compression_opt.py:235-244
    assert d == dict(md_frames=89, md_time=12.35, md_avg=0.1387, cp_frames=89, cp_time=3.20, cp_avg=0.0360, total=15.54) or \
        d == dict(md_frames=89, md_time=12.35, md_avg=0.1387, cp_frames=89, cp_time=3.20, cp_avg=0.0360, total=15.55)
    assert open(p).read().startswith("Motion Detection:\n  Frames processed: 89\n")


def test_gpu_statistics_file(tmp_path):
    path = hl.write_gpu_statistics(str(tmp_path), dict(frames=10, pixels=1000, motion_pixels=25, blocks=100, static_blocks=90))
    d = json.load(open(path))
    assert os.path.basename(path) == "gpu_statistics.json"
    assert d["motion_pixel_percent"] == 2.5 and d["static_block_percent"] == 90.0 and d["frames"] == 10


def test_video_stem_matches_reference_naming():
    # frame_differencing.py:32 / motion_compression_opt.py:197: os.path.splitext(os.path.basename(path))[0]
    assert hl.video_stem("/a/b/cam.01.mp4") == "cam.01" and hl.video_stem("clip.avi") == "clip"
