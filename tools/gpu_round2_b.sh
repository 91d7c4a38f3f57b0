#!/bin/bash
mkdir -p gpurun_out
for g in randn chair dups; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launch_$g.csv python tools/nn_once.py $g > gpurun_out/b_ncu_$g.log 2>&1
  python tools/launch_summary.py gpurun_out/b_launch_$g.csv "nn_once $g"
done
for g in randn chair dups; do
  timeout 300 python tools/nn_once.py $g --lib=tools/wip/libpnae_r1.so --time
  timeout 300 python tools/nn_once.py $g --lib=pointnet_autoencoder_b200/libpnae.so --time
done
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -k "emd or sweep_sizes or host_pipeline or graph_step" > gpurun_out/b_pytest.log 2>&1
tail -40 gpurun_out/b_pytest.log
