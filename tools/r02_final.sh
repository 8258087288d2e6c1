#!/bin/bash
# end-of-round check on one GPU: the whole gpu suite, smoke(), the default bench line, the f3 line
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/final
mkdir -p "$out"
timeout 1200 python -m pytest tests -x -q -m gpu > "$out/gpu.log" 2>&1; echo "gpu suite rc=$?" | tee "$out/summary.txt"
tail -4 "$out/gpu.log" | tee -a "$out/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$out/summary.txt"
tail -2 "$out/smoke.log" | tee -a "$out/summary.txt"
timeout 600 python bench.py > "$out/bench.json" 2> "$out/bench.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
timeout 300 python bench.py --config f3 --steps 2 --warmup 1 > "$out/bench_f3.json" 2> "$out/bench_f3.err"; echo "f3 rc=$?" | tee -a "$out/summary.txt"
