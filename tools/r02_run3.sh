#!/bin/bash
# round 2, GPU call 3: one system-scope fence per kernel, shared profiling events, 7-nodes-per-thread march variant
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/run3
mkdir -p "$out"
timeout 900 python -m pytest tests -x -q -m gpu > "$out/gpu.log" 2>&1; echo "gpu suite rc=$?" | tee "$out/summary.txt"
tail -3 "$out/gpu.log" | tee -a "$out/summary.txt"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > "$out/bench1.json" 2> "$out/bench1.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
for A in "--full" "--full --march-alt 0" "8" "8 --march-alt 0" "8 --force-mg 2" "8 --force-mg 2 --march-alt 0" "4" "4 --force-mg 2" "2" "2 --force-mg 2"; do
    timeout 300 python tools/slab_bench.py $A >> "$out/slab.jsonl" 2>> "$out/slab.err"
done
cat "$out/slab.jsonl" | tee -a "$out/summary.txt"
