"""HSIC.compress / decompress round trip at BASELINE configs[3] (1216x2176): timings of both symbol orders,
bpp of the files vs the ideal code length, exact reproduction of forward()'s latents and reconstructions.
    python tools/codec_roundtrip.py [H W] [--raster]"""
import sys
import tempfile
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.hsic import HSIC  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
H, W = (int(args[0]), int(args[1])) if len(args) >= 2 else (1216, 2176)
orders = ["wavefront_streams", "wavefront"] + (["raster"] if "--raster" in sys.argv else [])
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = HSIC().eval()
with torch.no_grad():
    net.encoder1.g_a_conv4.weight.mul_(8.0)          # non-degenerate latents (random init gives all-zero symbols)
    net.encoder2.g_a_conv4.weight.mul_(8.0)
net = net.to(dev)
net.update(force=True)
g = torch.Generator().manual_seed(5)
x1, x2 = torch.rand(1, 3, H, W, generator=g).to(dev), torch.rand(1, 3, H, W, generator=g).to(dev)
Hm = torch.tensor([[[1.0, 0.01, 20.0], [0.0, 1.0, 3.0], [1e-6, 0.0, 1.0]]], device=dev)
with torch.no_grad(), tempfile.TemporaryDirectory() as tmp:
    fwd = net(x1, x2, Hm)
    for order in orders:
        best_c = best_d = 1e9
        for rep in range(4):                         # first repetition builds plans and buffers; host-side coding times
            t0 = time.time()                         # jitter with the box (threads, tmpfs): best of the last three
            enc = net.compress(x1, x2, Hm, "p", tmp, y_order=order)
            torch.cuda.synchronize()
            t1 = time.time()
            dec = net.decompress(x1, x2, Hm, "p", tmp, device=dev)
            torch.cuda.synchronize()
            t2 = time.time()
            if rep:
                best_c, best_d = min(best_c, t1 - t0), min(best_d, t2 - t1)
        t0, t1, t2 = 0.0, best_c, best_c + best_d
        ok = all(torch.equal(dec[k], fwd[k]) for k in ("y1_hat", "x1_hat", "x2_hat")) and torch.equal(dec["y2_hat"], enc["y2_hat"])
        print(f"{order:9s} {H}x{W}: compress {1e3 * (t1 - t0):8.1f} ms (y coding {1e3 * enc['enctime']:.1f}), "
              f"decompress {1e3 * (t2 - t1):8.1f} ms (y decoding {1e3 * dec['dectime']:.1f}); {enc['n_symbols']} symbols, "
              f"bpp_real {enc['bpp_real']:.5f}, y bytes {enc['y_bytes']} vs ideal {enc['y_bits_ideal'] / 8:.0f}; exact={ok}")
