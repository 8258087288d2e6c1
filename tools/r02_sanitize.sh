#!/bin/bash
# compute-sanitizer on the exchange protocol (2 GPUs): memcheck / racecheck / synccheck of slab solves that use the fused
# HaloPort exchange -- the unchanged fish.c with -p4b_gpus 2 (one process, two host threads) and the plane-marching kernels
# forced onto a small grid through the torchrun worker.
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/run18
mkdir -p "$out"
F="./p4pdes_b200/bin/fish -fsh_dim 3 -da_refine 5 -pc_type mg -mg_levels_pc_type jacobi -ksp_rtol 1e-10 -ksp_converged_reason -p4b_gpus 2"
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 $F > "$out/fish_$tool.log" 2>&1; echo "fish -p4b_gpus 2 $tool rc=$?" | tee -a "$out/summary.txt"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Linear solve|error \|u" "$out/fish_$tool.log" | tee -a "$out/summary.txt"
done
W="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 tests/mgpu_worker.py --refine 5 --rtol 1e-10 --march-min-plane 1 --rep-points 1 --no-oracle"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --print-limit 20 $W > "$out/march_$tool.log" 2>&1; echo "march kernels on slabs $tool rc=$?" | tee -a "$out/summary.txt"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|MGPU_RESULT" "$out/march_$tool.log" | cut -c1-300 | tee -a "$out/summary.txt"
done
