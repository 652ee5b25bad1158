"""Per-step CUDA-event durations when the engine's three lanes run CONCURRENTLY (as in the captured graph),
next to the serial (eager, one stream) durations: shows which launches are stretched by co-running kernels.
    python tools/lane_profile.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.hsic import HSIC  # noqa: E402

h, w = 1216, 2176
torch.manual_seed(0)
net = HSIC().eval().cuda()
eng = net.engine_for(1, h, w, torch.device("cuda:0"))
serial = dict(eng.profile_steps(iters=3))
main = torch.cuda.current_stream()
side = [torch.cuda.Stream(), torch.cuda.Stream()]
lanes = [main] + side


def run(record):
    for sd in side:
        sd.wait_stream(main)
    events = {}
    for kind, a, lane in eng.sched:
        if kind == "run":
            with torch.cuda.stream(lanes[lane]):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(lanes[lane])
                eng.steps[a][1]()
                e1.record(lanes[lane])
                record.append((a, lane, e0, e1))
        elif kind == "record":
            ev = torch.cuda.Event(); ev.record(lanes[lane]); events[a] = ev
        else:
            lanes[lane].wait_event(events[a])
    for sd in side:
        main.wait_stream(sd)


for _ in range(2):
    run([])
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
rec = []
t0.record()
run(rec)
t1.record()
torch.cuda.synchronize()
print(f"concurrent eager wall: {t0.elapsed_time(t1):.3f} ms; serial sum {sum(serial.values()):.3f} ms")
tot = {0: 0.0, 1: 0.0, 2: 0.0}
for a, lane, e0, e1 in rec:
    n = eng.steps[a][0]
    d = e0.elapsed_time(e1)
    tot[lane] += d
    start = t0.elapsed_time(e0)
    flag = " <<<" if d > 1.3 * serial[n] + 0.01 else ""
    print(f"lane{lane} start {start:7.3f} dur {d:7.4f} serial {serial[n]:7.4f}  {n}{flag}")
print("per-lane busy:", tot)
