# On an N-GPU box (default 8): host topology, raw pinned-copy contention probe (1, 2, 4, N GPUs at once; both directions and each alone
# at N), then the benchmark at N GPUs.  Everything lands in gpurun_out/scale_*.
N=${1:-8}
mkdir -p gpurun_out
{
  echo "== topology"; nvidia-smi topo -m 2>&1 | head -40
  echo "== cpu"; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core"
  echo "== memory"; free -g | head -2; (numactl -H 2>/dev/null || echo "numactl not installed") | head -12
  echo "== pcie link per GPU"; nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max --format=csv
} > gpurun_out/scale_topology.txt 2>&1
probe() {  # $1 = number of GPUs at once, $2 = mode
  START_AT=$(python -c "import time; print(time.time() + 22)")
  for g in $(seq 0 $(($1 - 1))); do START_AT=$START_AT PROBE_MODE=$2 CUDA_VISIBLE_DEVICES=$g python tools/pcie_probe2.py & done
  wait
}
{
  for k in 1 2 4 $N; do [ $k -le $N ] && { echo "== $k GPU(s) at once, both directions"; probe $k both; }; done
  echo "== $N GPUs at once, host->device only"; probe $N h2d
  echo "== $N GPUs at once, device->host only"; probe $N d2h
} > gpurun_out/scale_pcie_probe.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 \
  > gpurun_out/scale_bench_${N}gpu.log 2> gpurun_out/scale_bench_${N}gpu.err
tail -c 300 gpurun_out/scale_bench_${N}gpu.err
cat gpurun_out/scale_pcie_probe.txt
