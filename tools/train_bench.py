"""Training-step benchmark (BASELINE.json configs[4]): 512x896 patches, batch 2 per GPU, data-parallel with one NCCL
all-reduce of the flat gradient buffer per step.

    python tools/train_bench.py [--steps 10] [--warmup 3] [--profile]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py

Prints one JSON line on rank 0: pairs/s (whole job), ms per step (device-timed, max over ranks), useful TFLOP/s
(3 x the forward conv FLOPs per pair: forward + data gradient + weight gradient)."""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.hsic import HSIC  # noqa: E402

H, W, B = 512, 896, 2
FWD_FLOP_PER_PAIR = 281.96e9            # SURVEY §8d (forward, 512x896)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(0)
    net = HSIC().to(dev).train()
    tr = net.trainer(B, H, W, dev, lmbda=0.01)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
    aux = torch.optim.Adam(net.aux_parameters(), lr=1e-3, fused=True)
    g = torch.Generator().manual_seed(100 + rank)
    x1 = torch.rand(4, B, 3, H, W, generator=g).to(dev)
    x2 = torch.rand(4, B, 3, H, W, generator=g).to(dev)
    Hm = torch.eye(3).repeat(B, 1, 1)
    Hm[:, 0, 2] = 12.0
    Hm = Hm.to(dev)
    losses = []
    for i in range(max(1, a.warmup)):
        losses.append(tr.train_step(x1[i % 4], x2[i % 4], Hm, opt, aux)["loss"])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        losses.append(tr.train_step(x1[i % 4], x2[i % 4], Hm, opt, aux)["loss"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    if rank == 0:
        line = {"metric": "training stereo pairs/s at 512x896 (forward+backward+allreduce+2x Adam), batch 2 per GPU",
                "value": world * B / (ms / 1e3), "unit": "pairs/s", "n_gpus": world, "ms_per_step": ms, "steps": a.steps,
                "useful_tflops_per_gpu": 3 * FWD_FLOP_PER_PAIR * B / (ms / 1e3) / 1e12, "dtype": "bf16", "data": "synthetic",
                "scaling": "weak", "loss_first": losses[0], "loss_last": losses[-1],
                "config": {"workload": "MASIC codec training step, 512x896 patches, batch 2 per GPU (BASELINE.json configs[4])",
                           "parallelism": f"dp{world}: one NCCL all-reduce of the flat fp32 gradient buffer "
                                          f"({tr.flat_grad.numel()} floats) per step"}}
        print(json.dumps(line), flush=True)
        if a.profile:
            for name, t in sorted(tr.profile(), key=lambda kv: -kv[1]):
                print(f"{t:9.3f} ms  {name}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
