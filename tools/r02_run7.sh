#!/bin/bash
# 8 GPUs: bench at N=8 and N=4 (same box), a trace of the 8-GPU solve, and three N=8 parity cases
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/run7
mkdir -p "$out"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 > "$out/bench8.json" 2> "$out/bench8.err"; echo "bench8 rc=$?" | tee "$out/summary.txt"
timeout 400 $TR --nproc-per-node 8 --master-port 29552 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --trace "$out/trace8.json" > "$out/bench8t.json" 2> "$out/bench8t.err"; echo "bench8 trace rc=$?" | tee -a "$out/summary.txt"
timeout 400 $TR --nproc-per-node 4 --master-port 29553 bench.py --gpus 4 --steps 5 --warmup 3 > "$out/bench4.json" 2> "$out/bench4.err"; echo "bench4 rc=$?" | tee -a "$out/summary.txt"
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "c_oracle_at_257 or (bit_identical_to_push and 8)" > "$out/multi8.log" 2>&1; echo "multi8 rc=$?" | tee -a "$out/summary.txt"
tail -3 "$out/multi8.log" | tee -a "$out/summary.txt"
