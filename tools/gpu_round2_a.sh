#!/bin/bash
# first GPU call of round 2: parity of the new Chamfer sweep, timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -x -k "nn_distance or chamfer or fwd_grad or fused or host_pipeline or wide_index or buffers" > gpurun_out/a_pytest_nn.log 2>&1
echo "pytest nn rc=$?" >> gpurun_out/a_pytest_nn.log
tail -5 gpurun_out/a_pytest_nn.log
for g in randn chair dups; do timeout 300 python tools/graph_time.py 32 2048 2048 --gen=$g; done > gpurun_out/a_time.log 2>&1
timeout 300 python tools/graph_time.py 64 2048 2048 >> gpurun_out/a_time.log 2>&1
timeout 300 python tools/graph_time.py 8 16384 16384 >> gpurun_out/a_time.log 2>&1
timeout 300 python tools/graph_time.py 4 2048 2048 >> gpurun_out/a_time.log 2>&1
cat gpurun_out/a_time.log
timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 -k "not (nn_distance or chamfer or fwd_grad or fused or host_pipeline or wide_index or buffers)" > gpurun_out/a_pytest_rest.log 2>&1
echo "pytest rest rc=$?" >> gpurun_out/a_pytest_rest.log
tail -30 gpurun_out/a_pytest_rest.log
