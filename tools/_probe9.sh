python -m pytest tests/test_trainer_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/train_bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c90-200
MASIC_TRAIN_LANES=0 python tools/train_bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c90-200
