set -e
python tools/k8_probe.py > gpurun_out/k8_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_degrade8" -s 9 -c 1 -f -o gpurun_out/prof_k8mco_r2 python tools/k8_probe.py > gpurun_out/ncu_k8.log 2>&1
cat gpurun_out/k8_plain.log
