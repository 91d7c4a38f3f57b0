#!/bin/bash
# round-2 profiles of the final kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 16 --warmup 8 --no-cpu-baseline --no-emd --no-train --no-refgpu"
timeout 600 $CMD > gpurun_out/j_bench_short.json 2> gpurun_out/j_bench_short.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/j_launches.csv $CMD > gpurun_out/j_ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/j_launches.csv "$CMD" | tee gpurun_out/j_launches_summary.csv
for k in nn_fwd_kernel nn_finalize_kernel nn_bwd_kernel; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 2 -c 1 -o gpurun_out/j_full_$k -f python tools/run_once.py chamfer > gpurun_out/j_ncu_$k.log 2>&1
done
for k in approx_match_kernel match_cost_factors_kernel; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 1 -c 1 -o gpurun_out/j_full_$k -f python tools/run_once.py emd > gpurun_out/j_ncu_$k.log 2>&1
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:encoder_conv_pool_kernel -s 1 -c 1 -o gpurun_out/j_full_encoder -f python tools/run_once.py enc > gpurun_out/j_ncu_enc.log 2>&1
ls -la gpurun_out/j_full_*.ncu-rep
