python -m pytest tests/test_image_gpu.py -x -q -m gpu 2>&1 | tail -5
python tools/step_profile.py 2>&1 | grep -E "warp|mask|total" > gpurun_out/sp_warp.txt
MASIC_WARP_FAST=0 python tools/step_profile.py 2>&1 | grep -E "warp|mask|total" > gpurun_out/sp_warp0.txt
paste gpurun_out/sp_warp0.txt gpurun_out/sp_warp.txt | cut -c1-200
