#!/bin/bash
# 2 GPUs: pattern.c on y-slabs (parity vs one GPU, config-5 bench line at N=2), minimal config 4 after the base-grid fix
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/run9
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k pattern > "$out/multi_pattern.log" 2>&1; echo "pattern slabs rc=$?" | tee "$out/summary.txt"
tail -4 "$out/multi_pattern.log" | tee -a "$out/summary.txt"
timeout 300 python -m pytest tests/test_gpu_r2_minimal.py tests/test_gpu_minimal.py tests/test_gpu_r2_shim_minimal.py -x -q > "$out/minimal.log" 2>&1; echo "minimal rc=$?" | tee -a "$out/summary.txt"
tail -3 "$out/minimal.log" | tee -a "$out/summary.txt"
timeout 300 python bench.py --config c4 --steps 3 --warmup 2 > "$out/bench_c4.json" 2> "$out/bench_c4.err"; echo "c4 rc=$?" | tee -a "$out/summary.txt"
timeout 300 python bench.py --config c5 --steps 3 --warmup 2 > "$out/bench_c5_1.json" 2> "$out/bench_c5_1.err"; echo "c5 N=1 rc=$?" | tee -a "$out/summary.txt"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --config c5 --gpus 2 --steps 3 --warmup 2 > "$out/bench_c5_2.json" 2> "$out/bench_c5_2.err"; echo "c5 N=2 rc=$?" | tee -a "$out/summary.txt"
