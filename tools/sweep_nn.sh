#!/bin/bash
for v in "-DPNAE_NN_FINWAVES=1" "-DPNAE_NN_FINWAVES=2" "-DPNAE_NN_FINWAVES=4" "-DPNAE_NN_FINTHREADS=256 -DPNAE_NN_FINOCC=4"; do
  PNAE_NVCC_DEFS="$v" python -m pointnet_autoencoder_b200.build > /dev/null 2>&1 || echo "build failed: $v"
  echo "$v: $(PNAE_NVCC_DEFS="$v" python tools/graph_time.py 2>&1 | head -1)"
done
