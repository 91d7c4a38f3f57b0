#!/bin/bash
# tuning sweep for the Chamfer forward kernel (runs on the GPU box; rebuilds the library per variant)
for ctas in 3 4 5 6; do
  PNAE_NVCC_DEFS="-DPNAE_NN_CTAS=$ctas" python -m pointnet_autoencoder_b200.build --verbose 2>&1 | grep -A2 "Function properties.*nn_fwd" | grep -E "registers|spill" | tr '\n' ' '
  echo "ctas=$ctas: $(PNAE_NVCC_DEFS="-DPNAE_NN_CTAS=$ctas" python tools/graph_time.py 2>&1 | head -1)"
done
