"""Data-parallel training step, bucketed (two CUDA graphs, the early-final half of the gradients all-reduced under the rest
of the backward pass) against the plain form (one graph, one all-reduce): after a few steps every rank must hold exactly
the same parameters (a bucket reduced before its gradients were final would break that), and the loss trajectory must
agree with the plain form's to rounding noise (the warp backward accumulates with atomics, so two runs are not
bit-identical); the step time is printed for both.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_bucket_check.py"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.hsic import HSIC  # noqa: E402

H, W, B = 512, 896, 2
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
g = torch.Generator().manual_seed(100 + rank)                 # different data on every rank
x1 = torch.rand(4, B, 3, H, W, generator=g).to(dev)
x2 = torch.rand(4, B, 3, H, W, generator=g).to(dev)
Hm = torch.eye(3).repeat(B, 1, 1)
Hm[:, 0, 2] = 12.0
Hm = Hm.to(dev)
ng = torch.Generator().manual_seed(7 + rank)


def run(buckets: bool, steps: int = 4, timed: int = 10):
    os.environ["MASIC_TRAIN_BUCKETS"] = "1" if buckets else "0"
    torch.manual_seed(0)
    net = HSIC().to(dev).train()
    tr = net.trainer(B, H, W, dev, lmbda=0.01)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
    aux = torch.optim.Adam(net.aux_parameters(), lr=1e-3, fused=True)
    torch.manual_seed(1234 + rank)                            # the in-step uniform noise: same sequence in both runs
    losses = [tr.train_step(x1[i % 4], x2[i % 4], Hm, opt, aux)["loss"] for i in range(steps)]
    state = [p.detach().clone() for p in net.parameters()]
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(timed):
        tr.train_step(x1[i % 4], x2[i % 4], Hm, opt, aux)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / timed], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return state, losses, float(t)


sa, la, ta = run(True)
sb, lb, tb = run(False)
same = all(torch.equal(a, b) for a, b in zip(sa, sb))
worst = max(float((a - b).abs().max()) for a, b in zip(sa, sb))
# every rank must also hold the same parameters as rank 0
chk = torch.stack([p.double().sum() for p in sa]).sum().reshape(1)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
agree = all(float(c) == float(allc[0]) for c in allc)
if rank == 0:
    print(f"world {world}: bucketed {ta:.3f} ms/step, plain {tb:.3f} ms/step; ranks hold identical parameters: {agree}; "
          f"bucketed vs plain: bit-identical {same}, max |diff| {worst:.3e}; losses {la} vs {lb}", flush=True)
dist.destroy_process_group()
# Adam's first steps are sign-sensitive for near-zero gradients, so rounding noise in a gradient can move a parameter by
# lr: the forms are compared on the loss trajectory, the ranks on exact parameter equality
close = all(abs(a - b) <= 1e-5 * abs(b) for a, b in zip(la, lb))
sys.exit(0 if agree and close else 1)
