#!/bin/bash
# end-of-round check on one GPU: the whole gpu suite (every failure listed, no -x) and smoke(); `bench` as the first
# argument adds the default bench line and the f3 line
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/final
mkdir -p "$out"
timeout 1200 python -m pytest tests -q -m gpu --durations=8 > "$out/gpu.log" 2>&1; echo "gpu suite rc=$?" | tee "$out/summary.txt"
grep -E "^(FAILED|ERROR)" "$out/gpu.log" | tee -a "$out/summary.txt"
tail -3 "$out/gpu.log" | tee -a "$out/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$out/summary.txt"
tail -2 "$out/smoke.log" | tee -a "$out/summary.txt"
if [ "${1:-}" = bench ]; then
    timeout 600 python bench.py > "$out/bench.json" 2> "$out/bench.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
    timeout 300 python bench.py --config f3 --steps 2 --warmup 1 > "$out/bench_f3.json" 2> "$out/bench_f3.err"; echo "f3 rc=$?" | tee -a "$out/summary.txt"
fi
