#!/bin/bash
# the bench line on N GPUs of one box, launched the way the driver launches it:  gpurun --gpus N -- 'bash tools/gpu_round2_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 \
    bench.py --gpus $N --steps 200 --warmup 40 > gpurun_out/m_bench_${N}gpu.json 2> gpurun_out/m_bench_${N}gpu.err
echo "bench rc=$?"; tail -2 gpurun_out/m_bench_${N}gpu.err | cut -c1-300; cut -c1-600 gpurun_out/m_bench_${N}gpu.json
