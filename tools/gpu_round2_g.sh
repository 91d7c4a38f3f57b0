#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/g_pytest.log 2>&1; tail -15 gpurun_out/g_pytest.log
timeout 300 python tools/graph_time.py 32 2048 2048
timeout 900 python bench.py --steps 200 --warmup 40 > gpurun_out/g_bench1.json 2> gpurun_out/g_bench1.err; echo "bench rc=$?"; tail -3 gpurun_out/g_bench1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/g_bench1.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','check')}, indent=None))
print(json.dumps(d['roofline'], indent=None)[:900])
print(json.dumps(d['extra'], indent=None)[:3000])
PY
