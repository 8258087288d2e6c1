#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget was spent, in one go.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_pending.sh'
# Writes gpurun_out/pending/: the pytest log of the never-run device tests (marker gpu_pending), then timings of the
# unchanged C drivers (BASELINE configs 4 and 5) through both residual routes, and of the native hosts against the Python
# hosts.  Tests that pass here are promoted by renaming their marker to `gpu` (tests/conftest.py registers both).
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/pending
mkdir -p "$out"
python - <<'PY' > "$out/build.log" 2>&1
import __graft_entry__ as g
g.build()
PY
# 1. the validated suite first (it must still be green after the round-1 CPU-only changes: FD differencing step, shim)
timeout 900 python -m pytest tests -x -q -m gpu > "$out/gpu.log" 2>&1; echo "gpu suite rc=$?" | tee -a "$out/summary.txt"
# 2. the pending device tests
timeout 1200 python -m pytest tests -q -m gpu_pending -rA > "$out/gpu_pending.log" 2>&1; echo "gpu_pending rc=$?" | tee -a "$out/summary.txt"
tail -8 "$out/gpu_pending.log" | tee -a "$out/summary.txt"
# 3. the unchanged drivers at the BASELINE sizes
MG="-pc_type mg -mg_levels_pc_type jacobi"
C4="-da_grid_x 33 -da_grid_y 33 -snes_grid_sequence 6 -snes_fd_color -snes_converged_reason -log_view $MG"
( time timeout 600 ./p4pdes_b200/bin/minimal $C4 ) > "$out/minimal_c4_recognised.log" 2>&1
( time timeout 900 ./p4pdes_b200/bin/minimal $C4 -p4b_recognise_residual 0 ) > "$out/minimal_c4_hostcallback.log" 2>&1
C5="-da_grid_x 8 -da_grid_y 8 -da_refine 8 -ts_monitor -snes_converged_reason -p4b_mg_rscale 0.25 -log_view $MG"
( time timeout 600 ./p4pdes_b200/bin/pattern $C5 -ts_type beuler -ts_dt 5 -ts_max_time 25 ) > "$out/pattern_c5_beuler.log" 2>&1
( time timeout 600 ./p4pdes_b200/bin/pattern $C5 -ts_max_time 50 ) > "$out/pattern_c5_arkimex.log" 2>&1
grep -h "SNESSolve\|residual\|real" "$out"/minimal_c4_*.log "$out"/pattern_c5_*.log | tee -a "$out/summary.txt"
# 3b. GMRES orthogonalisation A/B on the same box: modified Gram-Schmidt (default) vs classical with batched dots
( time timeout 600 ./p4pdes_b200/bin/minimal $C4 -p4b_gmres_cgs 1 ) > "$out/minimal_c4_cgs.log" 2>&1
( time timeout 600 ./p4pdes_b200/bin/pattern $C5 -ts_type beuler -ts_dt 5 -ts_max_time 25 -p4b_gmres_cgs 1 ) > "$out/pattern_c5_beuler_cgs.log" 2>&1
grep -h "SNESSolve\|real" "$out"/minimal_c4_cgs.log "$out"/pattern_c5_beuler_cgs.log | tee -a "$out/summary.txt"
# 4. native hosts against the Python hosts (same kernels, no interpreter between them)
python - <<'PY' 2>&1 | tee -a "$out/summary.txt"
import time
from p4pdes_b200 import minimal as pm, pattern as pp
from p4pdes_b200.fish import Context
ctx = Context()
for native in (False, True):
    r = pm.minimal_main("-da_grid_x 33 -da_grid_y 33 -snes_grid_sequence 6 -snes_fd_color -pc_type mg", ctx, native=native)
    print("minimal 2049^2 native=%s: %.3f s, Newton %s, error %.3e" % (native, r.seconds, [s.its for s in r.stages], r.errinf))
    r = pp.pattern_main("-da_grid_x 8 -da_grid_y 8 -da_refine 8 -ts_type beuler -ts_dt 5 -ts_max_time 25 -pc_type mg "
                        "-p4b_mg_rscale 0.25", ctx, native=native)
    print("pattern 2048^2 beuler native=%s: %.3f s for %d steps" % (native, r.seconds, len(r.steps)))
PY
