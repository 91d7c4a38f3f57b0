#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_train.py -m gpu -q --maxfail=10 > gpurun_out/l_pytest_enc.log 2>&1; tail -12 gpurun_out/l_pytest_enc.log
python - <<'PY'
import torch, time, sys
sys.path.insert(0,'.')
from pointnet_autoencoder_b200.encoder import PointNetEncoder
def t(fn, it=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): g.replay()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1)/it*1e3
pc = torch.randn(32,2048,3,device='cuda')
for mode in (True, "conv5", False):
    enc = PointNetEncoder(fused=mode).cuda().train()
    with torch.no_grad():
        print("encoder forward, training-mode BN, B=32 N=2048, fused=%r: %.1f us" % (mode, t(lambda: enc(pc))))
PY
