# A/B of the fused K1 + vote segment length (measure flavour)
export DVC_LIB_FLAVOUR=measure
for s in 32 48 64 80 113; do
  DVC_FUSE_SEG=$s python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-fd --no-streams > gpurun_out/seg.log 2>/dev/null
  python - <<PY
import json
l=json.loads(open("gpurun_out/seg.log").read().strip().splitlines()[-1]); print("seg $s", round(l["value"]), "serial", round(l["roofline"]["serialised_fps_per_gpu"]), {k:round(v,1) for k,v in l["roofline"]["kernel_ms_in_timed_region"].items()})
PY
done
