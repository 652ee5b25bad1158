#!/bin/bash
# ncu evidence of a round (run under gpurun; every ncu command is preceded by the same command without ncu):
#   gpurun_out/r2_launches.csv   launch list of the bench command          -> profiles/r2_launches.md
#   gpurun_out/r2_traffic.csv    DRAM bytes + time of every launch of one eager step -> profiles/r2_conv_traffic.json
#   gpurun_out/r2_full.ncu-rep   ncu --set full of one launch per kernel family      -> profiles/r2_ncu_full.md
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --blocks none > gpurun_out/plain_bench.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --blocks none > gpurun_out/ncu_bench.log 2>&1
echo "launches rc=$?"
python tools/ncu_target.py ALL > gpurun_out/plain_all.log 2>&1 &&
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/r2_traffic.csv python tools/ncu_target.py ALL > gpurun_out/ncu_all.log 2>&1
echo "traffic rc=$?"
python tools/ncu_target.py > gpurun_out/plain_sel.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_full \
    python tools/ncu_target.py > gpurun_out/ncu_sel.log 2>&1
echo "full rc=$?"
