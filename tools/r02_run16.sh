#!/bin/bash
# 8 GPUs: pattern.c on y-slabs (parity at N=8, config-5 bench lines at N=8 and 4), the unchanged fish.c with -p4b_gpus 8/4,
# the headline bench at N=8 (default and with the 129^3 level replicated as well)
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/run16
mkdir -p "$out"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "pattern and 8" > "$out/pattern8.log" 2>&1; echo "pattern N=8 rc=$?" | tee "$out/summary.txt"
tail -3 "$out/pattern8.log" | tee -a "$out/summary.txt"
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "one_process and (8 or 4)" > "$out/fishN.log" 2>&1; echo "fish -p4b_gpus rc=$?" | tee -a "$out/summary.txt"
tail -3 "$out/fishN.log" | tee -a "$out/summary.txt"
timeout 300 $TR --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 --steps 5 --warmup 3 > "$out/bench8.json" 2> "$out/bench8.err"; echo "bench8 rc=$?" | tee -a "$out/summary.txt"
timeout 300 $TR --nproc-per-node 8 --master-port 29572 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --rep-points 2200000 > "$out/bench8_rep129.json" 2> "$out/bench8_rep129.err"; echo "bench8 rep129 rc=$?" | tee -a "$out/summary.txt"
timeout 300 $TR --nproc-per-node 8 --master-port 29573 bench.py --config c5 --gpus 8 --steps 3 --warmup 2 > "$out/bench_c5_8.json" 2> "$out/bench_c5_8.err"; echo "c5 N=8 rc=$?" | tee -a "$out/summary.txt"
timeout 300 $TR --nproc-per-node 4 --master-port 29574 bench.py --config c5 --gpus 4 --steps 3 --warmup 2 > "$out/bench_c5_4.json" 2> "$out/bench_c5_4.err"; echo "c5 N=4 rc=$?" | tee -a "$out/summary.txt"
( cd p4pdes_b200/bin; for n in 1 8; do ( time ./fish -fsh_dim 3 -da_refine 8 -pc_mg_levels 7 -pc_type mg -mg_levels_pc_type jacobi -ksp_rtol 1e-10 -ksp_converged_reason -log_view -p4b_gpus $n ) 2>&1 | grep -v "^\[p4b" | tail -9; done ) > "$out/fish513.log" 2>&1
