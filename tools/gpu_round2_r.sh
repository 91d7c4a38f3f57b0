#!/bin/bash
for n in "" occ12 occ16 lanes2 lanes8 w2 t64o16; do
  if [ -z "$n" ]; then echo "--- default"; timeout 300 python tools/graph_time.py 32 2048 2048 | grep "fwd  \|fused"; else
  echo "--- $n"; PNAE_LIB_OVERRIDE=tools/wip/variants/libpnae_$n.so timeout 300 python tools/graph_time.py 32 2048 2048 | grep "fwd  \|fused"; fi
done
