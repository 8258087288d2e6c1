#!/bin/bash
# round 2, GPU call 2: re-run what failed in call 1, the new BASELINE-size oracle parity tests, a first bench with the
# polled host scalars, and the thin-slab kernel measurements (one rank's share of the 8-GPU run on one GPU).
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/run2
mkdir -p "$out"
timeout 900 python -m pytest tests -x -q -m gpu > "$out/gpu.log" 2>&1; echo "gpu suite rc=$?" | tee "$out/summary.txt"
tail -3 "$out/gpu.log" | tee -a "$out/summary.txt"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > "$out/bench1.json" 2> "$out/bench1.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
for N in 8 4; do
  for M in "" "--march 8,256" "--march 7,256" "--march 7,512" "--march 4,512"; do
    timeout 300 python tools/slab_bench.py $N $M >> "$out/slab.jsonl" 2>> "$out/slab.err"
  done
  timeout 300 python tools/slab_bench.py $N --force-mg 2 >> "$out/slab.jsonl" 2>> "$out/slab.err"
  timeout 300 python tools/slab_bench.py $N --force-mg 2 --march 7,256 >> "$out/slab.jsonl" 2>> "$out/slab.err"
done
cat "$out/slab.jsonl" | tee -a "$out/summary.txt"
