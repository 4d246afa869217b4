#!/usr/bin/env python
"""Stand-alone timing of the fused tail kernel (head + iSTFT + sub-band synthesis) on synthetic logits.

    python tools/tail_bench.py [--config ljs_mb_istft_vits] [--batch 64] [--frames 862] [--reps 20]

Prints ms per launch and the achieved algorithmic GB/s (5632 B per latent frame).  Used under ncu for the
kernel's profile (`ncu --set full -k regex:tail ...`).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mb_istft_vits_b200 import Engine, get_config, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="ljs_mb_istft_vits")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=862)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    cfg = get_config(a.config)
    eng = Engine(cfg, synth.make_state_dict(cfg, seed=1234), precision="bf16")
    L = a.frames
    for u in cfg["upsample_rates"]:
        L *= u
    nch = 18 * (cfg["subbands"] if cfg["variant"] != "istft" else 1)
    logits = torch.randn((a.batch, L + 1, nch), device="cuda") * 0.5
    for _ in range(3):
        eng.tail(logits, a.frames, want_mb=False, want_spec=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        eng.tail(logits, a.frames, want_mb=False, want_spec=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    nbytes = a.batch * a.frames * 5632.0
    print("tail %s B=%d T=%d: %.4f ms  %.0f GB/s algorithmic" % (a.config, a.batch, a.frames, ms, nbytes / ms / 1e6))


if __name__ == "__main__":
    main()
