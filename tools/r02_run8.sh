#!/bin/bash
# 1 GPU: the whole gpu suite after the kernel changes (ywrap, random init, recognition), the secondary bench configs
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/run8
mkdir -p "$out"
timeout 1200 python -m pytest tests -x -q -m gpu > "$out/gpu.log" 2>&1; echo "gpu suite rc=$?" | tee "$out/summary.txt"
tail -4 "$out/gpu.log" | tee -a "$out/summary.txt"
for c in c1 c2 c4 c5; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 > "$out/bench_$c.json" 2> "$out/bench_$c.err"; echo "bench $c rc=$?" | tee -a "$out/summary.txt"
done
