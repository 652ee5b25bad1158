"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in a container that has /root/reference (it cannot travel to the GPU box):

    make -C oracle ref && python tests/golden/make_golden.py

The reference is imported where it lies (oracle/refimport.py); its two C++ extensions are
the binaries `make -C oracle ref` compiled from its own sources.  Every fixture records
the seeds and the reference call that produced it.  The fixtures pin oracle/ (see
tests/test_oracle_pinned.py), and through the oracle the CUDA path.
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
from oracle import refimport  # noqa: E402

torch.set_num_threads(8)
torch.backends.mkldnn.enabled = True


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()


def entropy_inputs(shape, seed, scale=4.0):
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(*shape, generator=g) * scale
    flat = v.view(-1)
    n = flat.numel()
    # exact .5 ties, integers, and far-out values (bypass / table overflow)
    flat[: n // 16] = torch.round(flat[: n // 16]) + 0.5
    flat[n // 16: n // 8] = torch.round(flat[n // 16: n // 8])
    flat[-8:] = torch.tensor([60.0, -75.5, 1000.25, -1000.75, 0.5, -0.5, 1.5, -2.5])
    return v


def main():
    MASIC = refimport.import_masic()
    import compressai
    from compressai._CXX import pmf_to_quantized_cdf as ref_pmf2cdf
    from compressai.entropy_models import (EntropyBottleneck, GaussianConditional,
                                           GaussianMixtureConditional_gf)
    from compressai.layers import GDN
    import kornia

    meta = {"reference": "ywz978020607/MASIC @ /root/reference (unmodified)",
            "torch": torch.__version__, "numpy": np.__version__,
            "note": "kornia is oracle/shims/kornia (restated 0.5.0): warp parity is unpinned by the reference"}

    # ---- (1) pmf_to_quantized_cdf known-answer tests (ops.cpp:40-109)
    rng = np.random.default_rng(1234)
    kats = []
    cases = [[0.1, 0.2, 0.7], [1.0], [0.5, 0.5], [1e-9, 1.0, 1e-9], [0.0, 1.0, 0.0, 0.0], [1e-7] * 40 + [1.0]]
    for t in range(60):
        n = int(rng.integers(2, 48))
        p = rng.random(n).astype(np.float32) ** int(rng.integers(1, 9))
        if t % 3 == 0:
            p[int(rng.integers(0, n))] = 0.0
        if t % 4 == 0:
            p = p / p.sum()
        if t % 7 == 0:
            p = (p * 1e-5).astype(np.float32)
        if p.sum() * 65536 < 1:
            continue
        cases.append([float(v) for v in p])
    for p in cases:
        try:
            kats.append({"pmf": [float(np.float32(v)) for v in p], "cdf": list(ref_pmf2cdf([float(v) for v in p], 16))})
        except Exception as e:  # noqa: BLE001
            kats.append({"pmf": [float(np.float32(v)) for v in p], "error": type(e).__name__})
    (HERE / "pmf_cdf_kat.json").write_text(json.dumps({"precision": 16, "cases": kats}))

    # ---- (2) EntropyBottleneck (entropy_models.py:242-430)
    torch.manual_seed(11)
    eb = EntropyBottleneck(16).eval()
    with torch.no_grad():
        eb.quantiles.data[:, 0, 0] = -torch.rand(16) * 12 - 2.3
        eb.quantiles.data[:, 0, 1] = torch.randn(16) * 0.7
        eb.quantiles.data[:, 0, 2] = torch.rand(16) * 14 + 1.7
        for f in eb._factors:
            f.data = torch.randn_like(f) * 0.5
        for m in eb._matrices:
            m.data = m.data + torch.randn_like(m) * 0.3
    z = entropy_inputs((1, 16, 6, 7), seed=3, scale=4.0)
    with torch.no_grad():
        z_hat, z_lik = eb(z)
        eb.update(force=True)
        sym = eb._quantize(z, "symbols", eb._medians().detach().view(1, -1, 1, 1))
        strings = eb.compress(z)
        z_dec = eb.decompress(strings, z.shape[-2:])
        aux = eb.loss()
    sd = {k: v.clone() for k, v in eb.state_dict().items()}
    np.savez_compressed(HERE / "eb.npz", z=z.numpy(), z_hat=z_hat.numpy(), lik=z_lik.numpy(), symbols=sym.numpy(),
                        z_dec=z_dec.numpy(), string0=np.frombuffer(strings[0], dtype=np.uint8),
                        aux_loss=np.float32(aux.item()),
                        **{"sd/" + k: v.numpy() for k, v in sd.items()})
    # default-init tables (row 0 starts [0,1218,2497,...], length 23, offset -10)
    torch.manual_seed(0)
    eb0 = EntropyBottleneck(128).eval()
    eb0.update()
    np.savez_compressed(HERE / "eb_init_seed0.npz", offset=eb0._offset.numpy(), cdf=eb0._quantized_cdf.numpy(),
                        length=eb0._cdf_length.numpy(), biases0=eb0._biases[0].detach().numpy())

    # ---- (3) GaussianConditional (entropy_models.py:433-562), 64-level table (google.py:195-201)
    from compressai.models.google import get_scale_table
    table = get_scale_table()
    gc = GaussianConditional(list(map(float, table))).eval()
    gc.update()
    g = torch.Generator().manual_seed(5)
    scales = torch.exp(torch.rand(1, 12, 9, 10, generator=g) * 9 - 3.2)      # 0.04 .. 330
    scales.view(-1)[:6] = torch.tensor([0.0, 0.05, 0.11, 0.1100001, 256.0, 300.0])
    scales.view(-1)[6:6 + 64] = torch.tensor([float(v) for v in table])      # exactly on table entries
    means = torch.randn(1, 12, 9, 10, generator=g)
    y = entropy_inputs((1, 12, 9, 10), seed=6, scale=3.0)
    with torch.no_grad():
        idx = gc.build_indexes(scales)
        y_hat, y_lik = gc(y, scales, means)
        gsym = gc._quantize(y, "symbols", means)
        gstr = gc.compress(y, idx, means)
        y_dec = gc.decompress(gstr, idx, means)
    rows = [0, 1, 17, 40, 63]
    np.savez_compressed(HERE / "gc.npz", scale_table=np.asarray(table, dtype=np.float32), scales=scales.numpy(),
                        means=means.numpy(), y=y.numpy(), indexes=idx.numpy(), y_hat=y_hat.numpy(),
                        lik=y_lik.numpy(), symbols=gsym.numpy(), y_dec=y_dec.numpy(),
                        string0=np.frombuffer(gstr[0], dtype=np.uint8),
                        offset=gc._offset.numpy(), length=gc._cdf_length.numpy(),
                        cdf_rows=np.asarray(rows), cdf_sel=gc._quantized_cdf[rows].numpy(),
                        cdf_sha256=np.frombuffer(bytes.fromhex(sha(gc._quantized_cdf)), dtype=np.uint8),
                        cdf_shape=np.asarray(gc._quantized_cdf.shape))

    # ---- (4) GaussianMixtureConditional_gf (entropy_models.py:713-858): what HSIC.gaussian1/2 are
    K, M = 5, 8
    gm = GaussianMixtureConditional_gf(K=K).eval()
    g = torch.Generator().manual_seed(7)
    yg = entropy_inputs((2, M, 10, 12), seed=8, scale=3.0)
    sig = torch.exp(torch.rand(2, M * K, 10, 12, generator=g) * 9 - 3.5)
    sig.view(-1)[:5] = torch.tensor([0.0, 0.05, 0.11, 300.0, 1e-9])
    mu = torch.randn(2, M * K, 10, 12, generator=g) * 2
    wl = torch.randn(2, M * K, 10, 12, generator=g)
    w = torch.softmax(wl.view(2, K, M, 10, 12), dim=1).reshape(2, M * K, 10, 12)
    with torch.no_grad():
        yq, yl = gm(yg, sig, mu, w)
    np.savez_compressed(HERE / "gmm.npz", K=K, y=yg.numpy(), sigma=sig.numpy(), mu=mu.numpy(), w_logits=wl.numpy(),
                        w=w.numpy(), y_hat=yq.numpy(), lik=yl.numpy())

    # ---- (5) GDN / IGDN (layers/gdn.py:41-92)
    torch.manual_seed(21)
    out = {}
    for inv in (False, True):
        layer = GDN(12, inverse=inv).eval()
        with torch.no_grad():
            layer.beta.data = layer.beta.data + torch.rand(12) * 0.5
            layer.gamma.data = torch.sqrt(torch.rand(12, 12) * 0.05 + 0.1 * torch.eye(12))
            layer.gamma.data[0, 1] = 1e-7            # below the 2^-18 bound
            layer.beta.data[2] = 1e-5                # below the beta bound
            x = torch.randn(2, 12, 9, 11) * 2
            yv = layer(x)
        tag = "igdn" if inv else "gdn"
        out.update({f"{tag}/x": x.numpy(), f"{tag}/y": yv.numpy(), f"{tag}/beta": layer.beta.detach().numpy(),
                    f"{tag}/gamma": layer.gamma.detach().numpy()})
    np.savez_compressed(HERE / "gdn.npz", **out)

    # ---- (6) HSIC: state_dict layout + seeded forward (MASIC.py:652-851)
    torch.manual_seed(0)
    net = MASIC.HSIC(N=128, M=192, K=5).eval()
    sd = net.state_dict()
    layout = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()]
    n_main = sum(p.numel() for p in net.parameters())
    n_aux = sum(p.numel() for p in net.aux_parameters())
    (HERE / "hsic_layout.json").write_text(json.dumps(
        {"entries": layout, "main_params": n_main, "aux_params": n_aux,
         "param_sha256": {k: sha(sd[k]) for k in ("encoder1.g_a_conv1.weight", "encoder1.g_a_conv2.bias",
                                                  "decoder2.after_conv.weight", "entropy_bottleneck2._biases.3",
                                                  "_h_s2_same_resolution.gmm_weights.4.weight",
                                                  "mask2weights_unit.maskconv.6.bias",
                                                  "context_prediction2.weight")}}, indent=0))

    sys.path.insert(0, str(ROOT))
    from oracle.hsic import synthetic_homography

    def run_case(tag, h, w, scale_y):
        torch.manual_seed(0)
        net = MASIC.HSIC(N=128, M=192, K=5).eval()
        if scale_y != 1.0:
            with torch.no_grad():
                net.encoder1.g_a_conv4.weight.mul_(scale_y)
                net.encoder2.g_a_conv4.weight.mul_(scale_y)
        g = torch.Generator().manual_seed(100)
        x1 = torch.rand(1, 3, h, w, generator=g)
        x2 = torch.rand(1, 3, h, w, generator=g)
        Hm = synthetic_homography(1, seed=1)
        with torch.no_grad():
            o = net(x1, x2, Hm)
        npx = h * w
        import math
        bpp = {k: float(torch.log(v).sum() / (-math.log(2) * npx)) for k, v in o["likelihoods"].items()}
        mse1 = float(torch.mean((o["x1_hat"] - x1) ** 2))
        mse2 = float(torch.mean((o["x2_hat"] - x2) ** 2))
        cy, cx = h // 2 - 24, w // 2 - 24
        np.savez_compressed(
            HERE / f"hsic_forward_{tag}.npz", h=h, w=w, scale_y=scale_y, x_seed=100, h_seed=1, H=Hm.numpy(),
            y1_hat=o["y1_hat"].numpy(), z1_hat=o["z1_hat"].numpy(),
            lik_y1=o["likelihoods"]["y1"].numpy(), lik_y2=o["likelihoods"]["y2"].numpy(),
            lik_z1=o["likelihoods"]["z1"].numpy(), lik_z2=o["likelihoods"]["z2"].numpy(),
            x1_hat_crop=o["x1_hat"][:, :, cy:cy + 48, cx:cx + 48].numpy(),
            x2_hat_crop=o["x2_hat"][:, :, cy:cy + 48, cx:cx + 48].numpy(),
            x1_hat_mean=float(o["x1_hat"].mean()), x2_hat_mean=float(o["x2_hat"].mean()),
            mask_R_sum=float(o["x1_mask_R"].sum()), mask_L_sum=float(o["x1_mask_L"].sum()),
            mask_R_row=o["x1_mask_R"][0, 0, h // 2].numpy(), mask_L_row=o["x1_mask_L"][0, 0, h // 2].numpy(),
            bpp_y1=bpp["y1"], bpp_y2=bpp["y2"], bpp_z1=bpp["z1"], bpp_z2=bpp["z2"], mse1=mse1, mse2=mse2,
            y1_nonzero=int((o["y1_hat"] != 0).sum()))
        print(tag, "bpp", sum(bpp.values()), bpp, "mse", mse1, mse2, "y1 nonzero", int((o["y1_hat"] != 0).sum()))

    run_case("init_128x192", 128, 192, 1.0)
    run_case("scaled8_128x192", 128, 192, 8.0)
    run_case("scaled50_128x192", 128, 192, 50.0)

    # ---- (7) warp / mask / mask2weights (kornia restatement; unpinned by the reference)
    g = torch.Generator().manual_seed(9)
    img = torch.rand(2, 3, 40, 56, generator=g)
    Hm = synthetic_homography(2, seed=3)
    Hm[:, 0, 2] *= 0.3
    wimg = kornia.warp_perspective(img, Hm, (40, 56))
    m_r, m_l = MASIC.mask(img, Hm)
    np.savez_compressed(HERE / "warp.npz", img=img.numpy(), H=Hm.numpy(), warped=wimg.numpy(), mask_R=m_r.numpy(),
                        mask_L=m_l.numpy())

    (HERE / "META.json").write_text(json.dumps(meta, indent=1))
    print("fixtures written to", HERE)


if __name__ == "__main__":
    main()
