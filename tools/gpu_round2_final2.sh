#!/bin/bash
# closing run, part 2: bench lines with the encoder figures from graph replays, and the encoder kernels' full captures
mkdir -p gpurun_out
timeout 1300 python bench.py > gpurun_out/f_bench_1gpu.json 2> gpurun_out/f_bench_1gpu.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 24 --warmup 8 > gpurun_out/f_bench_1gpu_driver_args.json 2> gpurun_out/f_bench_driver.err; echo "bench(driver args) rc=$?"
for k in encoder_conv_pool_kernel mlp_first_kernel mlp_apply_bf16_kernel; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 2 -c 1 -o gpurun_out/f_full_$k -f python tools/enc_once.py > gpurun_out/f_ncu_$k.log 2>&1
done
ls -la gpurun_out/f_full_*.ncu-rep
