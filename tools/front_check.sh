# front kernels: parity subset (fd + window), then fd-mode loop timing and the default window line
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin_gpu.py -m gpu -x -q -k "gray or fd or config1 or window or loop or stream or smoke or sizes" 2>&1 | tail -4
timeout 600 python bench.py --mode fd --steps 6 --no-cpu-baseline --no-e2e --no-fd --no-streams > gpurun_out/fd.log 2>gpurun_out/fd.err || tail -c 600 gpurun_out/fd.err
timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-e2e --no-fd --no-streams > gpurun_out/win.log 2>gpurun_out/win.err || tail -c 600 gpurun_out/win.err
python - <<'PY'
import json
for n in ("fd", "win"):
    l=json.loads(open("gpurun_out/%s.log" % n).read().strip().splitlines()[-1]); print(n, round(l["value"]), {k:round(v,1) for k,v in l["roofline"]["kernel_ms_in_timed_region"].items()})
PY
