#!/usr/bin/env python
"""Time the tail kernels alone at BASELINE size (B = 64, T = 862): the fused conv_post + tail kernel (16-bit paths) on a
random operand tensor and the stand-alone tail kernel on random fp32 logits.  CUDA events, 20 launches after warm-up.
    python tools/tail_bench.py [--config ljs_mb_istft_vits] [--precision bf16]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mb_istft_vits_b200 import Engine, get_config, synth  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="ljs_mb_istft_vits")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=862)
    a = ap.parse_args()
    cfg = get_config(a.config)
    eng = Engine(cfg, synth.make_state_dict(cfg, seed=1234), precision=a.precision)
    B, T = a.batch, a.frames
    L = 16 * T
    logits = torch.randn((B, L + 1, 72), device="cuda") * 0.5
    us = timed(lambda: eng.tail(logits, T, want_mb=False, want_spec=False))
    print(f"stand-alone tail on fp32 logits : {us:7.1f} us  ({B * T * 5632 / us / 1e3:.0f} GB/s algorithmic)")
    C = cfg["upsample_initial_channel"] // 4
    act = (torch.randn((B, L + 1, C), device="cuda") * 0.5).to(torch.bfloat16 if a.precision == "bf16" else torch.float16)
    us = timed(lambda: eng.tail_fused(act, T))
    fl = 2.0 * 72 * 7 * C * B * (L + 1)
    print(f"fused conv_post + tail          : {us:7.1f} us  ({fl / us / 1e6:.0f} TFLOP/s, {B * T * (32 * C + 1024) / us / 1e3:.0f} GB/s algorithmic)")


if __name__ == "__main__":
    main()
