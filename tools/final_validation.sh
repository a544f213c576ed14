#!/bin/bash
# Round-end evidence run (under gpurun, one GPU): smoke, bench line, per-op table, ncu launch list, ncu --set full of the
# hot-path kernels.  Everything lands in gpurun_out/; tools/summarize_ncu.py turns it into the files under profiles/.
set -u
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 240 python bench.py 2> gpurun_out/bench.err | tail -1 > gpurun_out/bench_n1.json; echo "bench rc=$?"
timeout 200 python tools/bench_ops.py > gpurun_out/ops.md 2> gpurun_out/ops.err; echo "ops rc=$?"
timeout 120 python tools/bench_pooler_layouts.py 2>/dev/null | tail -1 > gpurun_out/layouts.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py 12 > gpurun_out/ncu_launch.log 2>&1; echo "launches rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on \
    -k regex:"roi_align_nhwc|nchw_to_nhwc|paste_masks|nms_scan|nms_mask" -s 8 -c 8 -o gpurun_out/step_full -f \
    python tools/profile_step.py 3 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out | head -30
