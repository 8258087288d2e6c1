#!/bin/bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/run4
mkdir -p "$out"
for A in "8 --trace" "8 --force-mg 2 --trace" "8 --force-mg 2 --port-opts 8" "8 --force-mg 2 --port-opts 16" "8 --force-mg 2 --port-opts 32" "8 --force-mg 2 --port-opts 48" "8 --force-mg 2 --port-opts 40" "--full" "--full --march-alt 0"; do
    timeout 300 python tools/slab_bench.py $A >> "$out/slab.jsonl" 2>> "$out/slab.err"
done
