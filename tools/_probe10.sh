timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/train_bucket_check.py > gpurun_out/bucket_check.log 2>&1
echo rc=$?; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bucket_check.log | grep -i "world\|error\|Traceback" -A3 | head -30
