#!/bin/bash
mkdir -p gpurun_out
echo "--- current"; timeout 300 python tools/graph_time.py 32 2048 2048 --gen=chair --emd | tail -3
echo "--- prefetch"; PNAE_LIB_OVERRIDE=tools/wip/variants/libpnae_amprefetch.so timeout 300 python tools/graph_time.py 32 2048 2048 --gen=chair --emd | tail -3
echo "--- prefetch B=4"; PNAE_LIB_OVERRIDE=tools/wip/variants/libpnae_amprefetch.so timeout 300 python tools/graph_time.py 4 2048 2048 --gen=chair --emd | tail -3
PNAE_LIB_OVERRIDE=tools/wip/variants/libpnae_amprefetch.so timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -k "emd or sweep_sizes or approx or match or full_size" > gpurun_out/p_pytest_emd.log 2>&1; tail -4 gpurun_out/p_pytest_emd.log
