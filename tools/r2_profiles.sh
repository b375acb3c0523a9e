# Round-2 evidence for the default bench command: launch list (per-launch device times) and ncu --set full of the window-loop kernels.
set -e
C="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fd --no-streams --no-e2e"
$C > gpurun_out/plain_r2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|dvc" -c 600 --csv --log-file gpurun_out/r2_bench_launch_list.csv $C > gpurun_out/ncu_r2_list.log 2>&1
C2="python bench.py --steps 1 --warmup 3 --frames 450 --no-cpu-baseline --no-fd --no-streams --no-e2e"
$C2 > gpurun_out/plain_r2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_degrade4s|k_gray_diff_vote|k_morph_chain" -s 9 -c 3 -f -o gpurun_out/prof_r2_loop $C2 > gpurun_out/ncu_r2_full.log 2>&1
ls -la gpurun_out/prof_r2_loop.ncu-rep gpurun_out/r2_bench_launch_list.csv
