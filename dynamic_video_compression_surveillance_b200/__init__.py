"""B200-native per-frame hot loop of carlozamu/dynamic-video-compression-surveillance.

Layout:
  csrc/        hand-written sm_100a kernels + the C ABI (include/dvc_b200.h) -> libdvc_b200.so
  _lib.py      ctypes binding of that ABI (fails loudly when the library is missing; no CPU fallback)
  pipeline.py  Python host objects over the ABI: ``FramePipeline`` (per-stream state + batch loop),
               stage-level ops on torch CUDA tensors
  sharding.py  multi-GPU partitioning: camera streams per rank, frame chunks with a temporal halo
  dropin/      ``frame_differencing`` / ``motion_compression_opt`` modules with the reference's names and
               signatures (windows.py:13-14 imports them unchanged)
  synth.py     deterministic synthetic clips shared by product benchmarks, oracle and tests
"""
from .synth import SyntheticClip, make_clip  # noqa: F401

__all__ = ["SyntheticClip", "make_clip"]
