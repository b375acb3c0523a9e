# environment switches exist only in the -DDVC_MEASURE flavour of the library: build it (here or before gpurun) and select it
export DVC_LIB_FLAVOUR=measure
[ -f dynamic_video_compression_surveillance_b200/libdvc_b200_measure.so ] || python dynamic_video_compression_surveillance_b200/build.py --measure
nvidia-smi --query-gpu=serial,temperature.gpu,temperature.memory,clocks.mem,clocks.max.mem --format=csv,noheader
for v in "$@"; do env $v python tools/k4_modes.py 2>&1 | tail -1; done
