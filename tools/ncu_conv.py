"""ncu target: selected conv_perf layers, one profiled launch each (cudaProfilerStart/Stop).
    python tools/ncu_conv.py g_a_conv2 mb_n256 ..."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import conv_perf

names = sys.argv[1:] or ["g_a_conv2"]
allL = dict(conv_perf.LAYERS); allL.update(conv_perf.MICRO)
plans = [(n, conv_perf.build(n, **allL[n])) for n in names]
for _, p in plans:
    p.launch(); p.launch()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _, p in plans:
    p.launch()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", names)
