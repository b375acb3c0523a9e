# ncu --set full of the fd-mode front kernel: one launch from a steady-state batch
set -e
C="python bench.py --mode fd --steps 1 --warmup 1 --frames 225 --no-cpu-baseline --no-e2e --no-fd --no-streams"
$C > gpurun_out/plain_fd.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_fd_front" -s 4 -c 1 -f -o gpurun_out/prof_fdfront_r2 $C > gpurun_out/ncu_fd_r2.log 2>&1
ls -la gpurun_out/prof_fdfront_r2.ncu-rep
