#!/bin/bash
# Round-end evidence run (under gpurun, one GPU): GPU tests, smoke, bench line (both arms), per-op table, ncu launch list of the
# bench command, ncu --set full of the hot-path kernels.  Everything lands in gpurun_out/; tools/summarize_ncu.py turns the
# captures into the text files committed under profiles/.  Every command carries its own timeout.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py 2> gpurun_out/bench.err | tail -1 > gpurun_out/bench_n1.json; echo "bench rc=$?"
timeout 400 python bench.py --impl reference 2> gpurun_out/bench_ref.err | tail -1 > gpurun_out/bench_ref_n1.json; echo "bench ref rc=$?"
timeout 120 python tools/kernel_times.py > gpurun_out/kernel_times.txt 2>/dev/null; echo "kernel_times rc=$?"
timeout 120 python tools/kernel_times.py --autograd --bf16 > gpurun_out/kernel_times_autograd_bf16.txt 2>/dev/null; echo "kernel_times autograd rc=$?"
timeout 120 python tools/bench_pooler_fwd.py > gpurun_out/pooler_fwd.json 2>/dev/null; echo "pooler fwd rc=$?"
timeout 500 python tools/bench_ops.py --out gpurun_out/ops.md > gpurun_out/ops.log 2> gpurun_out/ops.err; echo "ops rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1; echo "launches rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:"dcn_fwd_tc|dcn_bwd_data_tc|dcn_bwd_weight_cols|roi_align_bwd_nhwc|roi_align_nhwc|nms_rank|nms_mask|nms_scan" -c 24 \
    -o gpurun_out/train_full -f python tools/profile_train.py --iters 1 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out | head -40
