#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/pooler_ab.jsonl
for m in 0; do D2B_NHWC_MODE=$m timeout 200 python tools/bench_pooler_fwd.py >> gpurun_out/pooler_ab.jsonl 2>gpurun_out/pooler_ab_$m.err; done
cat gpurun_out/pooler_ab.jsonl
timeout 200 python -m pytest tests -q -m gpu -k "roi_align or pooler or half" > gpurun_out/pytest3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest3.log
timeout 600 python bench.py 2> gpurun_out/bench3.err | tail -1 > gpurun_out/bench3.json; echo "bench rc=$?"; tail -5 gpurun_out/bench3.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench3.json"))
    print("value", d["value"], "e2e", json.dumps(d["e2e"])[:900])
except Exception as e:
    print("bench parse failed", e)
PY
