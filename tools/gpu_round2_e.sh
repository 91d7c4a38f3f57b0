#!/bin/bash
mkdir -p gpurun_out
for g in randn dups; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e_launch_$g.csv python tools/nn_once.py $g > gpurun_out/e_ncu_$g.log 2>&1
  python tools/launch_summary.py gpurun_out/e_launch_$g.csv "nn_once $g"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e_launch_r1.csv python tools/nn_once.py randn --lib=tools/wip/libpnae_r1.so > gpurun_out/e_ncu_r1.log 2>&1
python tools/launch_summary.py gpurun_out/e_launch_r1.csv "nn_once randn r1 lib"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:nn_finalize_kernel -s 2 -c 1 -o gpurun_out/e_fin_full -f python tools/nn_once.py randn > gpurun_out/e_ncu_full.log 2>&1
