# ncu --set full of k_morph_chain in BASELINE config 3 (4K, rect 15 chain) and config 2 (1080p, ellipse-2 + 7x7): one launch each
set -e
C3="python bench.py --steps 1 --warmup 1 --resolution 4k --frames 90 --max-batch 90 --kernel-size 15 --morph-kernel 15 --morph-shape rect --no-cpu-baseline --no-e2e --no-fd --no-streams"
C2="python bench.py --steps 1 --warmup 1 --frames 225 --no-cpu-baseline --no-e2e --no-fd --no-streams"
$C3 > gpurun_out/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_morph_chain -s 3 -c 1 -f -o gpurun_out/prof_morph_c3 $C3 > gpurun_out/ncu_c3.log 2>&1
$C2 > gpurun_out/plain_c2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_morph_chain -s 3 -c 1 -f -o gpurun_out/prof_morph_c2 $C2 > gpurun_out/ncu_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
