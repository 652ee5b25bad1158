for n in 20 120; do
python tools/coresident_probe.py 3 $n >> gpurun_out/cores2.txt 2>&1
MASIC_WARP_BLOCK=128 python tools/coresident_probe.py 3 $n >> gpurun_out/cores2.txt 2>&1
MASIC_WARP_BLOCK=128 MASIC_CONV_SMEM_RESERVE=2048 python tools/coresident_probe.py 3 $n >> gpurun_out/cores2.txt 2>&1
MASIC_WARP_BLOCK=64 MASIC_CONV_SMEM_RESERVE=2048 python tools/coresident_probe.py 3 $n >> gpurun_out/cores2.txt 2>&1
done
cat gpurun_out/cores2.txt
