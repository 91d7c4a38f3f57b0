#!/bin/bash
mkdir -p gpurun_out
PNAE_LIB_OVERRIDE=tools/wip/variants/libpnae_tf32x1.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/o_launch_enc.csv python tools/enc_once.py > gpurun_out/o_ncu_enc.log 2>&1
python tools/launch_summary.py gpurun_out/o_launch_enc.csv "enc_once tf32x1" | grep mlp_ | cut -c1-100
