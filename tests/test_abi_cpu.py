"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/dvc_b200.h declares.  No compute call is made (there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dynamic_video_compression_surveillance_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_are_exported_and_bound(lib):
    from dynamic_video_compression_surveillance_b200 import _lib
    header = open(os.path.join(ROOT, "include", "dvc_b200.h")).read()
    declared = set(re.findall(r"\b(dvc_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_defaults(lib):
    from dynamic_video_compression_surveillance_b200._lib import DvcConfig
    assert lib.dvc_abi_version() == 3
    cfg = DvcConfig()
    lib.dvc_default_config(ctypes.byref(cfg))
    # the reference's defaults (frame_differencing.py:21-30; motion_compression_opt.py:29-31)
    assert (cfg.block_size, cfg.kernel_size, cfg.window_size, cfg.morph_kernel) == (4, 7, 30, 2)
    assert abs(cfg.motion_threshold - 0.5) < 1e-9 and cfg.min_area == 500 and cfg.release_factor == 0.5
    assert cfg.quantization_level == 100 and cfg.alpha_fraction == 0.2


def test_no_silent_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dynamic_video_compression_surveillance_b200 import pipeline
    with pytest.raises(RuntimeError):
        pipeline.FramePipeline(64, 64)
    from dynamic_video_compression_surveillance_b200._lib import DvcConfig
    cfg = DvcConfig()
    lib.dvc_default_config(ctypes.byref(cfg))
    cfg.width = cfg.height = 64
    h = ctypes.c_void_p()
    assert lib.dvc_create(ctypes.byref(cfg), ctypes.byref(h)) < 0
    assert b"no CUDA device" in lib.dvc_last_error(None)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dynamic_video_compression_surveillance_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)


def test_packed_fp32_is_not_contracted(lib):
    """ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 despite the rounding modifiers (seen with 12.9), which breaks the
    bit-exact DCT.  The K4 ring kernel is written so that no packed product feeds a packed addition; its SASS must hold exactly
    the packed operations of the algorithm: 2 x 16 FFMA2 in the DCT butterflies + 2 x 16 in the quotient, 56 FMUL2, 160 FADD2."""
    import shutil
    import subprocess
    from dynamic_video_compression_surveillance_b200 import build
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", build.LIB], capture_output=True, text=True).stdout
    counts, fn = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        for op in ("FFMA2", "FMUL2", "FADD2"):
            if fn and "k_degrade4s" in fn and (" " + op + " ") in line:
                counts.setdefault(fn, {}).setdefault(op, 0)
                counts[fn][op] += 1
    assert counts, "k_degrade4s not found in the library"
    for fn, c in counts.items():
        assert (c.get("FFMA2"), c.get("FMUL2"), c.get("FADD2")) == (64, 56, 160), (fn, c)
    # the packed 8x8 path (k_degrade4p.cuh::degrade_plane8_smem, rolled loops: each pass appears once): 24 + 6 FFMA2 in the
    # forward row / column bodies, 16 in the eight quotients, 24 + 6 in the inverse bodies; the four packed products per column
    # pair that feed additions are 8 scalar FMULs
    counts8, fn = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        for op in ("FFMA2", "FADD2", "FMUL"):
            if fn and "k_degrade8" in fn and (" " + op + " ") in line:
                counts8.setdefault(fn, {}).setdefault(op, 0)
                counts8[fn][op] += 1
    assert len(counts8) == 2, "k_degrade8<0> and k_degrade8<1> expected"
    for fn, c in counts8.items():
        assert (c.get("FFMA2"), c.get("FADD2"), c.get("FMUL")) == (76, 72, 8), (fn, c)


def test_product_library_reads_no_environment(lib):
    """Measurement switches (kernel generations, copy-only probes) are compiled out of the product build: it reports so and
    its sources call getenv only under #ifdef DVC_MEASURE (a stray variable must never change what the loop computes)."""
    import glob
    from dynamic_video_compression_surveillance_b200 import _lib
    assert lib.dvc_measure_build() == 0
    assert _lib.LIB_PATH.endswith("libdvc_b200.so")
    uses = []
    for f in glob.glob(os.path.join(os.path.dirname(_lib.LIB_PATH), "csrc", "*")):
        src = open(f).read()
        uses += [(os.path.basename(f), ln.strip()) for ln in src.splitlines() if "getenv(" in ln]
    # the only getenv in the sources is measure_env()'s, inside #ifdef DVC_MEASURE (cudart's own imports do not count)
    assert len(uses) == 1 and uses[0][1].startswith("static int measure_env("), uses
    src = open(os.path.join(os.path.dirname(_lib.LIB_PATH), "csrc", "dvc_b200.cu")).read()
    i = src.index("static int measure_env(")
    assert src.rfind("#ifdef DVC_MEASURE", 0, i) > src.rfind("#endif", 0, i)
