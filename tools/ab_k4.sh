# A/B of K4 variants inside one job (same box): env switches read at first launch of each process
mkdir -p gpurun_out
nvidia-smi --query-gpu=serial,temperature.gpu --format=csv,noheader
run() { python bench.py --steps 20 --no-cpu-baseline --no-e2e --no-overlap > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'degrade', round(d['roofline']['kernel_ms_in_timed_region']['degrade'],1), 'frac', round(d['roofline']['frac'],3))"; }
for v in "$@"; do
  env $v bash -c "$(declare -f run); run '$v'"
done
