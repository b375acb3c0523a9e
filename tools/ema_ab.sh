# EMA kernel: 8 vs 16 pixels per thread (measure flavour), then the product library's parity subset
for r in 16 8; do
DVC_LIB_FLAVOUR=measure DVC_EMA_PX=$r python bench.py --mode fd --steps 6 --no-cpu-baseline --no-e2e --no-fd --no-streams > gpurun_out/fd_e$r.log 2>gpurun_out/fd.err || tail -c 600 gpurun_out/fd.err
python - <<PY
import json
l=json.loads(open("gpurun_out/fd_e$r.log").read().strip().splitlines()[-1]); print("ema px $r fd", round(l["value"]), {k:round(v,1) for k,v in l["roofline"]["kernel_ms_in_timed_region"].items()})
PY
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin_gpu.py -m gpu -x -q -k "ema or fd or config1 or stream_group or sizes_not or random_loop or smoke" 2>&1 | tail -4
