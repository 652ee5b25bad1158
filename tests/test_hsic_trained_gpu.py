"""Parity in the regime the codec is used in: at random init the reconstructions sit at ~5 dB,
where |dPSNR| <= 0.01 dB says nothing.  Here a model is first TRAINED with the repo's own CUDA training step
(HSICTrainer.train_step) on smooth synthetic stereo pairs until its PSNR passes 25 dB, then the eval forward of the
CUDA engine is compared with the oracle (CPU fp32 restatement pinned to the reference) on the same trained weights:
|dPSNR| <= 0.01 dB, |dbpp| <= 0.1 % — the north-star's tolerances.  Margins go to gpurun_out/r2_parity.jsonl."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_trained_regime_psnr_and_bpp_parity():
    from masic_b200.hsic import HSIC
    from tools.train_regime import compare_with_oracle, train_to_psnr
    from parity_common import BPP_RTOL, PSNR_ATOL, record as _record
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = HSIC().to(dev)
    steps, psnr_train = train_to_psnr(net, dev, target_db=27.5, max_steps=1200, size=(256, 256), lr=3e-4, lmbda=0.05,
                                      log=print)
    assert psnr_train >= 25.0, f"training reached only {psnr_train:.2f} dB in {steps} steps"
    # the training step's atomics make every run a little different: top up (smaller steps) until the EVAL-mode
    # reconstructions of the comparison pair are in the regime too
    for extra in range(4):
        r = compare_with_oracle(net, dev, 256, 384)
        if min(r["psnr1_oracle"], r["psnr2_oracle"]) >= 25.0:
            break
        more, psnr_train = train_to_psnr(net, dev, target_db=99.0, max_steps=150, size=(256, 256), lr=1e-4, lmbda=0.05,
                                         seed=10 + extra, log=print)
        steps += more
    for (h, w) in ((256, 384), (512, 512)):
        r = compare_with_oracle(net, dev, h, w)
        r.update(train_steps=steps, train_psnr_db=psnr_train, tol=dict(dbpp_rel=BPP_RTOL, dpsnr_db=PSNR_ATOL))
        _record(f"trained_regime_{h}x{w}", **r)
        print(r)
        assert min(r["psnr1_oracle"], r["psnr2_oracle"]) >= 24.0, r     # the comparison is not vacuous
        assert r["dpsnr1_db"] <= PSNR_ATOL and r["dpsnr2_db"] <= PSNR_ATOL, r
        assert r["dbpp_rel"] <= BPP_RTOL, r
