# which kernels does the pipelined (three engines) step actually wait for?  `value` with groups of steps dropped
# (MASIC_ENGINE_SKIP: timing experiment, results are wrong).   bash tools/skip_ladder.sh [group ...]
run() { MASIC_ENGINE_SKIP="$1" timeout 300 python bench.py --steps 30 --warmup 6 --blocks none 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('%-60s value %.1f pairs/s  %.3f ms   e2e %.1f' % ('skip: $1', d['value'], d['ms_per_step'], d['e2e']['value']))
"; }
if [ $# -gt 0 ]; then for g in "$@"; do run "$g"; done; exit 0; fi
run ""
for g in "x1.pack_nhwc" "R.warp(x1)" "R.warp(x1_hat)" "mask_" "mask2weights" "pre_conv" "after_conv" "gmm_likelihood" \
         "latent_prep" "entropy_bottleneck" "deconv4" "h_a." "h_s." "context" "conv4" "gmm.l0" "gmm.l1" "gmm.l2,gmm.weights" \
         "conv1+gdn" "conv2+gdn" "conv3+gdn" "deconv1" "deconv2" "deconv3"; do run "$g"; done
run ""
