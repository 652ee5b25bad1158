python -m pytest tests/test_train_ops_gpu.py tests/test_trainer_gpu.py -x -q -m gpu 2>&1 | tail -4
python tools/train_bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-330
