# environment switches exist only in the -DDVC_MEASURE flavour of the library: build it (here or before gpurun) and select it
export DVC_LIB_FLAVOUR=measure
[ -f dynamic_video_compression_surveillance_b200/libdvc_b200_measure.so ] || python dynamic_video_compression_surveillance_b200/build.py --measure
# A/B of K4 variants inside one job (same box): env switches read at first launch of each process
mkdir -p gpurun_out
nvidia-smi --query-gpu=serial,temperature.gpu --format=csv,noheader
run() { python bench.py --steps 20 --no-cpu-baseline --no-e2e --no-overlap > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', round(d['value']), 'serial', round(r['serialised_fps_per_gpu']), 'degrade', round(r['kernel_ms_in_timed_region']['degrade'],1), 'frac', round(r['frac'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['sm_mhz_min'], d['clocks']['reasons'], d['clocks']['power_w_max'])"; }
for v in "$@"; do
  env $v bash -c "$(declare -f run); run '$v'"
done
