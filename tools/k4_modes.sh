nvidia-smi --query-gpu=serial,temperature.gpu,temperature.memory,clocks.mem,clocks.max.mem --format=csv,noheader
for v in "$@"; do env $v python tools/k4_modes.py 2>&1 | tail -1; done
