"""Drop-in modules with the reference's module names: put this directory first on ``sys.path`` (or use
``install()``) and ``windows.py``'s ``from frame_differencing import process_single_video_fd`` /
``from motion_compression_opt import process_single_video_of`` (windows.py:13-14) resolve to the GPU path."""
import os
import sys


def install() -> None:
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    root = os.path.dirname(os.path.dirname(here))
    if root not in sys.path:
        sys.path.insert(1, root)
