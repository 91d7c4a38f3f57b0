#!/bin/bash
# tuning sweep (runs on the GPU box; rebuilds the library per variant)
for v in "-DPNAE_NN_X=0" "-DPNAE_NN_BWD_NOTRIGGER=1"; do
  PNAE_NVCC_DEFS="$v" python -m pointnet_autoencoder_b200.build > /dev/null 2>&1 || echo "build failed: $v"
  echo "$v: $(PNAE_NVCC_DEFS="$v" python tools/graph_time.py 2>&1 | head -3 | tr '\n' ' ')"
  echo "$v: $(PNAE_NVCC_DEFS="$v" python bench.py --steps 800 --warmup 104 | cut -c60-200)"
done
