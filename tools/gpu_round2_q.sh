#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 200 --warmup 40 > gpurun_out/q_bench8.json 2> gpurun_out/q_bench8.err; echo "bench8 rc=$?"; tail -2 gpurun_out/q_bench8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/q_bench8.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus')}), json.dumps(d['e2e'])[:300])
print(json.dumps(d['extra'], indent=None)[:2200])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 train_bench.py --model upconv --tf32 --steps 30 --max-seconds 200 2>&1 | tail -1 | cut -c1-400
