# environment switches exist only in the -DDVC_MEASURE flavour of the library: build it (here or before gpurun) and select it
export DVC_LIB_FLAVOUR=measure
[ -f dynamic_video_compression_surveillance_b200/libdvc_b200_measure.so ] || python dynamic_video_compression_surveillance_b200/build.py --measure
# usage: bash tools/e2e_probe.sh "ENV=.. ENV=.." ...   : each argument is one configuration, run alone on GPU 0 and then on 2 GPUs at once
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg CUDA_VISIBLE_DEVICES=0 python tools/e2e_probe.py
  START_AT=$(python -c "import time; print(time.time() + 25)")
  env $cfg START_AT=$START_AT CUDA_VISIBLE_DEVICES=0 python tools/e2e_probe.py & env $cfg START_AT=$START_AT CUDA_VISIBLE_DEVICES=1 python tools/e2e_probe.py & wait
done
