#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/m_launch_enc.csv python tools/enc_once.py > gpurun_out/m_ncu_enc.log 2>&1
python tools/launch_summary.py gpurun_out/m_launch_enc.csv "enc_once" | cut -c1-150
