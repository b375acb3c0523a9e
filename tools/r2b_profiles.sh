# Round-2 (second half) evidence: default bench record, launch list of the same command, ncu --set full of the window-loop kernels
# and of the fd-mode mask kernels (fused front, one-launch contour filter, EMA), fd-mode launch list.
set -e
true
C="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fd --no-streams --no-e2e"
$C > gpurun_out/plain_r2b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|dvc" -c 600 --csv --log-file gpurun_out/${TAG:-r2b}_bench_launch_list.csv $C > gpurun_out/ncu_${TAG:-r2b}_list.log 2>&1
C2="python bench.py --steps 1 --warmup 3 --frames 450 --no-cpu-baseline --no-fd --no-streams --no-e2e"
$C2 > gpurun_out/plain_r2b2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_degrade4s|k_gray_diff_vote|k_morph_chain" -s 9 -c 3 -f -o gpurun_out/prof_${TAG:-r2b}_loop $C2 > gpurun_out/ncu_${TAG:-r2b}_full.log 2>&1
C3="python bench.py --mode fd --steps 1 --warmup 1 --frames 225 --no-cpu-baseline --no-e2e --no-fd --no-streams"
$C3 > gpurun_out/plain_r2b3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_fd_front|k_ccl_sweep|k_ema|k_morph_chain|k_degrade4s" -s 10 -c 5 -f -o gpurun_out/prof_${TAG:-r2b}_fd $C3 > gpurun_out/ncu_${TAG:-r2b}_fd.log 2>&1
C4="python bench.py --mode fd --steps 2 --warmup 3 --no-cpu-baseline --no-fd --no-streams --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|dvc" -c 400 --csv --log-file gpurun_out/${TAG:-r2b}_bench_fd_launch_list.csv $C4 > gpurun_out/ncu_${TAG:-r2b}_fdlist.log 2>&1
ls -la gpurun_out/prof_${TAG:-r2b}_loop.ncu-rep gpurun_out/prof_${TAG:-r2b}_fd.ncu-rep gpurun_out/${TAG:-r2b}_bench_launch_list.csv gpurun_out/${TAG:-r2b}_bench_fd_launch_list.csv
