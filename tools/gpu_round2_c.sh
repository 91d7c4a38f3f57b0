#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "nn_distance or chamfer or fwd_grad or fused or wide_index or buffers" > gpurun_out/c_pytest_nn.log 2>&1; tail -3 gpurun_out/c_pytest_nn.log
for g in randn chair dups; do
  timeout 300 python tools/nn_once.py $g --lib=pointnet_autoencoder_b200/libpnae.so --time
done
timeout 300 python tools/graph_time.py 32 2048 2048
timeout 600 ncu --set full --import-source on --clock-control none -k regex:nn_fwd_kernel -s 2 -c 1 -o gpurun_out/c_fwd_full -f python tools/nn_once.py randn > gpurun_out/c_ncu_full.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:nn_finalize_kernel -s 2 -c 1 -o gpurun_out/c_fin_full -f python tools/nn_once.py randn >> gpurun_out/c_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 900 python tools/emd_truth_table.py chair 32 2048 > gpurun_out/c_truth_chair.txt 2>&1; tail -3 gpurun_out/c_truth_chair.txt
timeout 900 python tools/emd_truth_table.py randn 8 2048 > gpurun_out/c_truth_randn.txt 2>&1; tail -12 gpurun_out/c_truth_randn.txt
