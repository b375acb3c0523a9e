# A/B of the 8x8 degrade kernels' occupancy target (measure flavour): CTAs of 128 threads per SM
export DVC_LIB_FLAVOUR=measure
for m in 4 5 6; do echo "DVC_K8_MINB=$m"; DVC_K8_MINB=$m python tools/k8_probe.py; done
