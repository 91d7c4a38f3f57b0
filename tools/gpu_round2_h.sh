#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests -m gpu -q -x -k "train" > gpurun_out/h_pytest_train.log 2>&1; tail -5 gpurun_out/h_pytest_train.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 200 --warmup 40 > gpurun_out/h_bench2.json 2> gpurun_out/h_bench2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/h_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/h_bench2.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus')}), json.dumps(d['e2e'])[:400])
print(json.dumps(d['extra'], indent=None)[:2500])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 train_bench.py --model upconv --tf32 --steps 30 --max-seconds 200 2>&1 | tail -2
