#!/bin/bash
# round-2 closing run on one B200: GPU tests, smoke, the bench lines, the ncu launch lists of the bench command and of
# the encoder forward, and one full ncu capture of every encoder kernel (the Chamfer / EMD captures in profiles/ are of
# kernels that did not change afterwards)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/f_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1300 python bench.py > gpurun_out/f_bench_1gpu.json 2> gpurun_out/f_bench_1gpu.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 24 --warmup 8 > gpurun_out/f_bench_1gpu_driver_args.json 2> gpurun_out/f_bench_driver.err; echo "bench(driver args) rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_reference_arm.json 2> gpurun_out/f_bench_ref.err; echo "bench(reference) rc=$?"
CMD="python bench.py --steps 16 --warmup 8 --no-cpu-baseline --no-emd --no-train --no-refgpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f_launches.csv $CMD > gpurun_out/f_ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/f_launches.csv "$CMD" > gpurun_out/f_launches_summary.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches_encoder.csv python tools/enc_once.py > gpurun_out/f_ncu_enc_launch.log 2>&1
python tools/launch_summary.py gpurun_out/f_launches_encoder.csv "python tools/enc_once.py" > gpurun_out/f_launches_encoder_summary.csv
for k in encoder_conv_pool_kernel mlp_first_kernel mlp_apply_bf16_kernel conv5_finish_kernel; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 2 -c 1 -o gpurun_out/f_full_$k -f python tools/enc_once.py > gpurun_out/f_ncu_$k.log 2>&1
done
# the third forward's layer-3 (64 -> 64) and layer-4 (64 -> 128) kernels
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp_layer_tc_kernel -s 7 -c 2 -o gpurun_out/f_full_mlp_layer_tc_kernel -f python tools/enc_once.py > gpurun_out/f_ncu_mlp_layer.log 2>&1
ls -la gpurun_out/f_full_*.ncu-rep
