# contour-filter sweep kernel: rows per chunk A/B (measure flavour)
for r in ${ROWS:-4 8 16}; do
DVC_LIB_FLAVOUR=measure DVC_CCL_ROWS=$r python bench.py --mode fd --steps 6 --no-cpu-baseline --no-e2e --no-fd --no-streams > gpurun_out/fd_$r.log 2>gpurun_out/fd.err || tail -c 600 gpurun_out/fd.err
python - <<PY
import json
l=json.loads(open("gpurun_out/fd_$r.log").read().strip().splitlines()[-1]); print("rows $r fd", round(l["value"]), {k:round(v,1) for k,v in l["roofline"]["kernel_ms_in_timed_region"].items()})
PY
done
