python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python tools/step_profile.py > gpurun_out/step_profile_new.txt 2>&1
grep -E "warp|mask|total" gpurun_out/step_profile_new.txt
python tools/coresident_probe.py 3 20
