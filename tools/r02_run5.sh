#!/bin/bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/run5
mkdir -p "$out"
for A in "8" "8 --force-mg 2 --trace" "4 --force-mg 2" "2 --force-mg 2"; do
    timeout 300 python tools/slab_bench.py $A >> "$out/slab.jsonl" 2>> "$out/slab.err"
done
timeout 900 python -m pytest tests -x -q -m gpu > "$out/gpu.log" 2>&1; echo "gpu suite rc=$?" | tee "$out/summary.txt"
tail -3 "$out/gpu.log" | tee -a "$out/summary.txt"
