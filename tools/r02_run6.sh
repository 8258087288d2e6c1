#!/bin/bash
# 2 GPUs: the multi-GPU parity suite (N=2 cases) and the 2-GPU bench line
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/run6
mkdir -p "$out"
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > "$out/multi.log" 2>&1; echo "multi rc=$?" | tee "$out/summary.txt"
tail -3 "$out/multi.log" | tee -a "$out/summary.txt"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > "$out/bench2.json" 2> "$out/bench2.err"; echo "bench2 rc=$?" | tee -a "$out/summary.txt"
