#!/usr/bin/env python
"""Per-launch device times of one hot-path step (eager launches, CUDA events around every kernel).

    python tools/launch_times.py [--config ljs_mb_istft_vits] [--batch 64] [--frames 862] [--reps 5] [--out file]

Prints, per launch of the step, the layer description, the median time over the repetitions, the layer's
algorithmic TFLOP/s and its share of the step.  In-step numbers (warm L2, sustained clocks), unlike an ncu
launch list whose launches are serialised with cold caches.
"""
import argparse
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mb_istft_vits_b200 import Engine, get_config, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="ljs_mb_istft_vits")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=862)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--out", default=None)
    ap.add_argument("--flags", type=int, default=0, help="extra mbv_config flags (16 = fused pairs, 32 = cluster pairs)")
    ap.add_argument("--ab-flags", type=int, default=None,
                    help="A/B: a second engine with these flags, run alternately with the first in the same process; prints both columns")
    a = ap.parse_args()
    cfg = get_config(a.config)
    sd = synth.make_state_dict(cfg, seed=1234)
    eng = Engine(cfg, sd, precision=a.precision, flags=a.flags)
    z_p, mask, _ = synth.make_latents(cfg, a.batch, a.frames, seed=1234)
    z_p, mask = z_p.cuda(), mask.cuda()
    g = None
    if cfg.get("gin_channels", 0):
        g = torch.randn((a.batch, cfg["gin_channels"], 1), generator=torch.Generator().manual_seed(7)).cuda()
    engb = Engine(cfg, sd, precision=a.precision, flags=a.ab_flags) if a.ab_flags is not None else None
    for _ in range(3):
        eng.flow_decode(z_p, mask, g, want_z=False)
        if engb is not None:
            engb.flow_decode(z_p, mask, g, want_z=False)
    torch.cuda.synchronize()
    runs, runs_b = [], []
    for e, r in ((eng, runs), (engb, runs_b)):
        if e is not None:
            e.set_profiling(True)
            e.profile_read()
    for _ in range(a.reps):
        for e, r in ((eng, runs), (engb, runs_b)):
            if e is None:
                continue
            e.flow_decode(z_p, mask, g, want_z=False)
            torch.cuda.synchronize()
            r.append(e.profile_read_launches())
    eng.set_profiling(False)
    n = len(runs[0])
    if engb is not None:  # interleaved A/B: same box, same clock state
        nb = len(runs_b[0])
        med = lambda rr, i: sorted(x[i][1] for x in rr)[len(rr) // 2] * 1e3
        lines = ["A: flags %d   B: flags %d   (median of %d alternating repetitions, us)" % (a.flags, a.ab_flags, a.reps)]
        ta = tb = 0.0
        for i in range(max(n, nb)):
            ua = med(runs, i) if i < n else float("nan")
            ub = med(runs_b, i) if i < nb else float("nan")
            ta += ua if i < n else 0.0
            tb += ub if i < nb else 0.0
            d = runs[0][i][0] if i < n else runs_b[0][i][0]
            lines.append("%3d  %-52s A %8.1f  B %8.1f  B-A %+7.1f" % (i, d, ua, ub, ub - ua))
        lines.append("sum of launches: A %.3f ms (%d launches)   B %.3f ms (%d launches)" % (ta / 1e3, n, tb / 1e3, nb))
        txt = "\n".join(lines)
        print(txt)
        if a.out:
            open(a.out, "w").write(txt + "\n")
        return
    lines = []
    total = 0.0
    B = a.batch
    for i in range(n):
        ts = sorted(r[i][1] for r in runs)
        med = ts[len(ts) // 2]
        total += med
        d = runs[0][i][0]
        tf = ""
        m = re.match(r"conv m(\d+) Ci(\d+) N(\d+) k(\d+) d(\d+) ph(\d+) L(\d+) nt(\d+)", d)
        if m:
            mode, ci, nn, k, dil, ph, L, nt = map(int, m.groups())
            flops = 2.0 * ci * nn * k * ph * L * B  # padded rows included: what the tensor pipe executes
            tf = "%7.0f TF/s(padded)" % (flops / (med * 1e-3) / 1e12)
        lines.append("%3d  %-52s %8.1f us  %s" % (i, d, med * 1e3, tf))
    lines.append("sum of launches: %.3f ms over %d launches" % (total, n))
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
