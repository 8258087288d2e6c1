#!/bin/bash
# round 2 ncu evidence (one GPU): launch list of one 513^3 solve and a --set full capture of the finest-level kernels
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/run17
mkdir -p "$out"
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file "$out/launches.csv" $B > "$out/ncu_list.log" 2>&1; echo "launch list rc=$?" | tee "$out/summary.txt"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"march|xp_update|r_update|prolong_add3d|restrict3d" -s 40 -c 22 -o "$out/finest" -f $B > "$out/ncu_full.log" 2>&1; echo "full capture rc=$?" | tee -a "$out/summary.txt"
ncu -i "$out/finest.ncu-rep" --page raw --csv > "$out/finest_raw.csv" 2>> "$out/ncu_full.log"
python profiles/summarize_ncu.py list "$out/launches.csv" "$out/launch_list.md" >> "$out/summary.txt" 2>&1
python profiles/summarize_ncu.py raw "$out/finest_raw.csv" "$out/ncu_full_finest_kernels.md" "$out/traffic.json" >> "$out/summary.txt" 2>&1
timeout 300 python profiles/microbench.py > "$out/microbench.jsonl" 2> "$out/microbench.err"; echo "microbench rc=$?" | tee -a "$out/summary.txt"
rm -f "$out/finest.ncu-rep"
