# ncu --set full of the fd-mode mask kernels (contour filter, fused front, EMA): one launch each from a steady-state batch
set -e
C="python bench.py --mode fd --steps 1 --warmup 1 --frames 225 --no-cpu-baseline --no-e2e --no-fd --no-streams"
$C > gpurun_out/plain_fd.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_ccl|k_fd_front|k_ema" -s 27 -c 9 -f -o gpurun_out/prof_fd_r2b $C > gpurun_out/ncu_fd_r2b.log 2>&1
ls -la gpurun_out/prof_fd_r2b.ncu-rep
