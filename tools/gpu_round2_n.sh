#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp_layer_kernel -s 3 -c 1 -o gpurun_out/n_full_mlp -f python tools/enc_once.py > gpurun_out/n_ncu.log 2>&1
ls -la gpurun_out/n_full_mlp.ncu-rep
