#!/bin/bash
# tuning sweep for the prefetch variant of approx_match (apply approx_match_prefetch.patch first; runs on the GPU box);
# needs TS <= THREADS and (2*THREADS) % TS == 0
for cfg in "256 128 2" "256 256 2" "128 128 4" "128 64 4" "256 64 2" "512 128 1"; do
  set -- $cfg
  D="-DPNAE_AM_THREADS=$1 -DPNAE_AM_TS=$2 -DPNAE_AM_CTAS=$3"
  PNAE_NVCC_DEFS="$D" python -m pointnet_autoencoder_b200.build > /dev/null 2>&1 || echo "build failed $cfg"
  echo "threads=$1 ts=$2 ctas=$3: $(PNAE_NVCC_DEFS="$D" python tools/graph_time.py --emd 2>&1 | grep approx_match)"
done
