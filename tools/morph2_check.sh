# merged MORPH_ELLIPSE (2, 2) pair: morphology / window-loop / OF-fixture parity, then the window line
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin_gpu.py -m gpu -x -q -k "morph or window or loop or of_ or stream or smoke or config" 2>&1 | tail -3
timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-e2e --no-fd --no-streams > gpurun_out/win.log 2>gpurun_out/win.err || tail -c 600 gpurun_out/win.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/win.log").read().strip().splitlines()[-1]); print("win", round(l["value"]), {k:round(v,1) for k,v in l["roofline"]["kernel_ms_in_timed_region"].items()})
PY
