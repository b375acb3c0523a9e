# environment switches exist only in the -DDVC_MEASURE flavour of the library: build it (here or before gpurun) and select it
export DVC_LIB_FLAVOUR=measure
[ -f dynamic_video_compression_surveillance_b200/libdvc_b200_measure.so ] || python dynamic_video_compression_surveillance_b200/build.py --measure
# A/B of env switches inside one job (same box): one bench per argument, prints loop fps and per-kernel-group ms
mkdir -p gpurun_out
nvidia-smi --query-gpu=serial,temperature.gpu --format=csv,noheader
run() { python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', round(d['value']), 'serial', round(r['serialised_fps_per_gpu']), {k: round(v,1) for k,v in r['kernel_ms_in_timed_region'].items()}, 'K4 frac', round(r['frac'],3))"; }
for v in "$@"; do
  env $v bash -c "$(declare -f run); run '$v'"
done
