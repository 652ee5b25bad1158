timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_bench.py --steps 1 --warmup 2 > gpurun_out/train_ncu.log 2>&1
tail -2 gpurun_out/train_ncu.log
wc -l gpurun_out/train_launches.csv
