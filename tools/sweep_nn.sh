#!/bin/bash
# tuning sweep (runs on the GPU box; rebuilds the library per variant)
for fl in 2 4 8; do
  PNAE_NVCC_DEFS="-DPNAE_NN_FINLANES=$fl" python -m pointnet_autoencoder_b200.build > /dev/null 2>&1
  echo "finlanes=$fl: $(PNAE_NVCC_DEFS="-DPNAE_NN_FINLANES=$fl" python tools/graph_time.py 2>&1 | head -1)"
done
