#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --maxfail=30 > gpurun_out/pytest5.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest5.log
timeout 600 python bench.py 2> gpurun_out/bench5.err | tail -1 > gpurun_out/bench5.json; echo "bench rc=$?"; tail -5 gpurun_out/bench5.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench5.json"))
    e = d["e2e"]
    print("value", d["value"], "e2e", e["value"], e["graphed"], e["fp32_transport"]["value"], e["fp32_transport_eager"]["value"])
    print(d["stages_ms"])
except Exception as ex:
    print("bench parse failed", ex)
PY
timeout 200 python tools/kernel_times.py --autograd --bf16 > gpurun_out/kernel_times_autograd_bf16.txt 2>gpurun_out/kt.err; head -12 gpurun_out/kernel_times_autograd_bf16.txt
