for d in 0 1 2 3 4 16 19; do echo "== MASIC_CONV_DEBUG=$d"; MASIC_CONV_DEBUG=$d timeout 120 python tools/conv_perf.py --only mb_gl 2>&1 | grep -v "^sum"; done
