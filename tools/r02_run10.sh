#!/bin/bash
# 2 GPUs: pattern slab parity (relaxed step comparison); why was c5 at N=1 6x slower on the 2-GPU box than on the 1-GPU box?
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/run10
mkdir -p "$out"
nproc > "$out/host.txt"; cat /sys/fs/cgroup/cpu.max >> "$out/host.txt" 2>/dev/null; nvidia-smi topo -m >> "$out/host.txt" 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k pattern > "$out/multi_pattern.log" 2>&1; echo "pattern slabs rc=$?" | tee "$out/summary.txt"
tail -4 "$out/multi_pattern.log" | tee -a "$out/summary.txt"
timeout 300 python bench.py --config c5 --steps 2 --warmup 1 > "$out/c5_plain.json" 2> "$out/c5_plain.err"
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --config c5 --steps 2 --warmup 1 > "$out/c5_vis0.json" 2> "$out/c5_vis0.err"
OMP_NUM_THREADS=1 timeout 300 python bench.py --config c5 --steps 2 --warmup 1 > "$out/c5_omp1.json" 2> "$out/c5_omp1.err"
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --config c5 --steps 2 --warmup 1 > "$out/c5_vis1.json" 2> "$out/c5_vis1.err"
for f in c5_plain c5_vis0 c5_omp1 c5_vis1; do python -c "
import json
b=json.loads([l for l in open('$out/$f.json') if l.startswith('{')][-1]); print('$f', b['ms_per_step'], b['ms_per_ts_step'])" | tee -a "$out/summary.txt"; done
