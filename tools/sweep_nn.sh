#!/bin/bash
# tuning sweep for the Chamfer forward kernel (runs on the GPU box; rebuilds the library per variant)
for ctas in 2 3 4; do
  PNAE_NVCC_DEFS="-DPNAE_NN_CTAS=$ctas" python -m pointnet_autoencoder_b200.build > /dev/null 2>&1
  echo "ctas=$ctas: $(PNAE_NVCC_DEFS="-DPNAE_NN_CTAS=$ctas" python tools/graph_time.py 2>&1 | head -1)"
done
