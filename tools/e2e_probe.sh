for c in 8 16 32; do
  export DVC_HOST_CHUNK=$c
  CUDA_VISIBLE_DEVICES=0 python tools/e2e_probe.py
  export START_AT=$(python -c "import time; print(time.time() + 25)")
  CUDA_VISIBLE_DEVICES=0 python tools/e2e_probe.py & CUDA_VISIBLE_DEVICES=1 python tools/e2e_probe.py & wait
  unset START_AT
done
