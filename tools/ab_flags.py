#!/usr/bin/env python
"""Interleaved A/B of two mbv_config.flags settings on ONE box in ONE process: two engines, each step one CUDA-graph replay
(as bench.py times it), blocks of `--steps` replays alternating A, B, A, B ... so that both arms see the same clock /
power state.  Prints the per-block ms per step and the medians.

    python tools/ab_flags.py --a 0 --b 512 [--blocks 6] [--steps 20] [--precision bf16] [--out file]
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
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--a", type=int, default=0)
    ap.add_argument("--b", type=int, required=True)
    ap.add_argument("--blocks", type=int, default=6)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    cfg = get_config(a.config)
    sd = synth.make_state_dict(cfg, seed=1234)
    z_p, mask, _ = synth.make_latents(cfg, a.batch, a.frames, seed=1234)
    z_p, mask = z_p.cuda(), mask.cuda()
    arms = []
    for flags in (a.a, a.b):
        eng = Engine(cfg, sd, precision=a.precision, flags=flags)
        for _ in range(3):
            eng.flow_decode(z_p, mask, want_z=False)
        torch.cuda.synchronize()
        arms.append((flags, eng, eng.capture_flow_decode(z_p, mask)[0], []))
    for _, _, gr, _ in arms:
        for _ in range(3):
            gr.replay()
    torch.cuda.synchronize()
    for _ in range(a.blocks):
        for _, _, gr, ms in arms:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1) / a.steps)
    lines = []
    for flags, _, _, ms in arms:
        s = sorted(ms)
        lines.append("flags %4d  median %.3f ms/step  min %.3f  blocks: %s" % (flags, s[len(s) // 2], s[0], " ".join("%.3f" % m for m in ms)))
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
