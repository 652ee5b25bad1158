set -x
python tools/coresident_probe.py 3 120 > gpurun_out/cores.txt 2>&1
for r in 1024 4096 12288 24576 49152; do MASIC_CONV_SMEM_RESERVE=$r python tools/coresident_probe.py 3 120 >> gpurun_out/cores.txt 2>&1; done
tail -8 gpurun_out/cores.txt
