#!/bin/bash
for v in "-DPNAE_NN_FIN_NOP=1"; do
  PNAE_NVCC_DEFS="$v" python -m pointnet_autoencoder_b200.build > /dev/null 2>&1 || echo "build failed: $v"
  for b in 32 64 128; do echo "$v B=$b: $(PNAE_NVCC_DEFS="$v" python tools/graph_time.py $b 2048 2048 2>&1 | head -1)"; done
done
