# contour-filter sweep kernel: parity subset, then fd-mode loop timing
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "contour or mask_rectangles" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin_gpu.py -m gpu -x -q -k "fd or config1 or stream_group or sizes_not or random_loop or smoke" 2>&1 | tail -4
timeout 600 python bench.py --mode fd --steps 6 --no-cpu-baseline --no-e2e --no-fd --no-streams > gpurun_out/fd.log 2>gpurun_out/fd.err || tail -c 600 gpurun_out/fd.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/fd.log").read().strip().splitlines()[-1]); print("fd", round(l["value"]), {k:round(v,1) for k,v in l["roofline"]["kernel_ms_in_timed_region"].items()})
PY
