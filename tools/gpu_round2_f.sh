#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "nn_distance or chamfer or fwd_grad or fused or wide_index or buffers" > gpurun_out/f_pytest_nn.log 2>&1; tail -3 gpurun_out/f_pytest_nn.log
for lib in tools/wip/libpnae_r1.so pointnet_autoencoder_b200/libpnae.so tools/wip/variants/libpnae_occ6.so tools/wip/variants/libpnae_fin256.so tools/wip/variants/libpnae_unroll1.so; do
 for g in randn chair dups; do
  timeout 300 python tools/nn_once.py $g --lib=$lib --time
 done
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launch_randn.csv python tools/nn_once.py randn > gpurun_out/f_ncu_randn.log 2>&1
python tools/launch_summary.py gpurun_out/f_launch_randn.csv "nn_once randn"
