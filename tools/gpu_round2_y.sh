#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_train.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -3
timeout 300 python tools/mlp_time.py 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/y_launches_encoder.csv python tools/enc_once.py > gpurun_out/y_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/y_launches_encoder.csv "python tools/enc_once.py" | head -12
