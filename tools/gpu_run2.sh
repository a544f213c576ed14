#!/bin/bash
# GPU run 2 of the round: new kernels (column-shared RoIAlign forward, Fast R-CNN / dense-head candidate kernels, half inputs)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --maxfail=30 > gpurun_out/pytest2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest2.log
rm -f gpurun_out/pooler_ab.jsonl
for m in 2 0 1; do D2B_NHWC_MODE=$m timeout 200 python tools/bench_pooler_fwd.py >> gpurun_out/pooler_ab.jsonl 2>gpurun_out/pooler_ab_$m.err; done
cat gpurun_out/pooler_ab.jsonl
timeout 500 python bench.py 2> gpurun_out/bench2.err | tail -1 > gpurun_out/bench2.json; echo "bench rc=$?"; tail -3 gpurun_out/bench2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench2.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("fp32_transport"), d["stages_ms"], d["roofline_other"]["roi_align_fwd_box_pooler"]["frac"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"roi_align_nhwc" -c 4 -o gpurun_out/pooler_fwd_full -f python tools/bench_pooler_fwd.py > gpurun_out/ncu_pooler.log 2>&1; echo "ncu rc=$?"
