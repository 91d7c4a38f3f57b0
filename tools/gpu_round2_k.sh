#!/bin/bash
for lib in pointnet_autoencoder_b200/libpnae.so tools/wip/variants/libpnae_ctas5.so tools/wip/variants/libpnae_ctas3.so tools/wip/variants/libpnae_unroll1.so tools/wip/variants/libpnae_unroll4.so tools/wip/variants/libpnae_ctas5u1.so; do
  timeout 300 python tools/nn_once.py randn --lib=$lib --time
done
