# morphology parity subset, then config 3 (4K, rect 15) and config 2 loop timings
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "morph or config3 or window_loop or of_mask or stream_group" 2>&1 | tail -3
python bench.py --steps 10 --resolution 4k --frames 900 --kernel-size 15 --morph-kernel 15 --morph-shape rect --no-cpu-baseline --no-e2e --no-fd --no-streams > gpurun_out/cfg3.log 2>gpurun_out/cfg3.err || tail -c 400 gpurun_out/cfg3.err
python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-fd --no-streams > gpurun_out/cfg2.log 2>gpurun_out/cfg2.err || tail -c 400 gpurun_out/cfg2.err
python - <<'PY'
import json
for f in ("cfg3","cfg2"):
    l=json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1]); print(f, round(l["value"]), {k:round(v,1) for k,v in l["roofline"]["kernel_ms_in_timed_region"].items()})
PY
