"""Shared by the GPU parity tests: the north-star's tolerances and the recorder of measured margins."""
import json
from pathlib import Path

BPP_RTOL = 1e-3          # |d bpp| <= 0.1 %
PSNR_ATOL = 0.01         # |d PSNR| <= 0.01 dB


def record(name, **kw):
    """Measured margins -> gpurun_out/r2_parity.jsonl (summarised into profiles/r2_parity.json after the GPU run)."""
    out = Path(__file__).resolve().parents[1] / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        with open(out / "r2_parity.jsonl", "a") as f:
            f.write(json.dumps({"case": name, **kw}) + "\n")
    except OSError:
        pass
