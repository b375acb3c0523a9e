# three-stream pipeline (front | mask | degrade) against the two-stream one (measure flavour), then parity of the product library
for p in 0 1; do
for mode in fd window; do
DVC_LIB_FLAVOUR=measure DVC_FRONT_STREAM=$p python bench.py --mode $mode --steps 6 --no-cpu-baseline --no-e2e --no-fd --no-streams > gpurun_out/fs_${mode}_$p.log 2>gpurun_out/fd.err || tail -c 600 gpurun_out/fd.err
python - <<PY
import json
l=json.loads(open("gpurun_out/fs_${mode}_$p.log").read().strip().splitlines()[-1]); print("front stream $p $mode", round(l["value"]), "serialised", round(l["roofline"]["serialised_fps_per_gpu"]))
PY
done
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
