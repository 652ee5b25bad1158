#!/bin/bash
# timing experiments for the conv kernel (results are garbage with MASIC_CONV_DEBUG != 0)
for d in 0 1 2 3; do
  echo "== MASIC_CONV_DEBUG=$d"
  MASIC_CONV_DEBUG=$d python tools/conv_perf.py --iters 5 2>&1 | grep -E "g_a_conv2|g_s_conv3|gmm_l0|g_a_conv1|subpix"
done
