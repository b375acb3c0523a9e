# full GPU suite, then window-mode and fd-mode loop timings
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-fd --no-streams > gpurun_out/cfg2.log 2>gpurun_out/cfg2.err || tail -c 400 gpurun_out/cfg2.err
python bench.py --mode fd --steps 6 --no-cpu-baseline --no-e2e --no-fd --no-streams > gpurun_out/fd.log 2>gpurun_out/fd.err || tail -c 600 gpurun_out/fd.err
python - <<'PY'
import json
for f in ("cfg2","fd"):
    l=json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1]); print(f, round(l["value"]), "serial", round(l["roofline"]["serialised_fps_per_gpu"]), {k:round(v,1) for k,v in l["roofline"]["kernel_ms_in_timed_region"].items()})
PY
