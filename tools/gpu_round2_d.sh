#!/bin/bash
mkdir -p gpurun_out
./tools/wip/microbench_nn | tee gpurun_out/d_microbench_nn.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "nn_distance or chamfer or fwd_grad or fused or wide_index or buffers or host_pipeline" > gpurun_out/d_pytest_nn.log 2>&1; tail -3 gpurun_out/d_pytest_nn.log
for g in randn chair dups; do
  timeout 300 python tools/nn_once.py $g --lib=pointnet_autoencoder_b200/libpnae.so --time
done
timeout 300 python tools/graph_time.py 32 2048 2048
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -k "emd or sweep_sizes" > gpurun_out/d_pytest_emd.log 2>&1; tail -5 gpurun_out/d_pytest_emd.log
