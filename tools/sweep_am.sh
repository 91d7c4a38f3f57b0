#!/bin/bash
# tuning sweep for approx_match (runs on the GPU box; rebuilds the library per variant)
for cfg in "256 256 2" "256 512 2" "512 128 1" "512 256 1" "256 192 2"; do
  set -- $cfg
  D="-DPNAE_AM_THREADS=$1 -DPNAE_AM_TS=$2 -DPNAE_AM_CTAS=$3"
  PNAE_NVCC_DEFS="$D" python -m pointnet_autoencoder_b200.build > /dev/null 2>&1
  echo "threads=$1 ts=$2 ctas=$3: $(PNAE_NVCC_DEFS="$D" python tools/graph_time.py --emd 2>&1 | grep approx_match)"
done
