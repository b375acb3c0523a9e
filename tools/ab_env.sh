run() { python bench.py --steps 20 --no-cpu-baseline --no-e2e --no-overlap > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['roofline']['kernel_ms_in_timed_region'], d['clocks']['power_w_max'])"; }
mkdir -p gpurun_out
nvidia-smi --query-gpu=serial,temperature.gpu,memory.used --format=csv,noheader
run first
python -m pytest tests/test_gpu_parity.py -q -x -k "contour or morph" 2>&1 | tail -1
nvidia-smi --query-gpu=serial,temperature.gpu,memory.used --format=csv,noheader
run after_pytest_small
python -m pytest tests -m gpu -q -x 2>&1 | tail -1
nvidia-smi --query-gpu=serial,temperature.gpu,memory.used --format=csv,noheader
run after_pytest_full
sleep 20
run after_sleep
