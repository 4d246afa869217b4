#!/usr/bin/env python
"""Where does the gap between the device-resident step and the host-to-host (`e2e`) step come from?

    python tools/e2e_probe.py [--steps 20] [--out file]

Times the `HostStream` serving loop of bench.py's e2e leg (pinned latents in, pinned waveform out, two slots, one captured
graph per slot) with its copies switched on and off, each variant from an idle GPU like the bench legs:
    resident   one graph replayed back to back (bench.py's `value`)
    slots      the two slot graphs alternating, event waits as in HostStream.submit, NO copies
    in         + the host->device copies of every step
    out        + the device->host copies of every step (no uploads)
    full       HostStream.submit itself (bench.py's `e2e`)
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mb_istft_vits_b200 import Engine, HostStream, get_config, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="ljs_mb_istft_vits")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=862)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cfg = get_config(args.config)
    sd = synth.make_state_dict(cfg, seed=1234)
    eng = Engine(cfg, sd, precision="bf16", device=0)
    B, T = args.batch, args.frames
    zp_h, m_h, _ = synth.make_latents(cfg, B, T, seed=1234)
    zp_pin = [zp_h.clone().pin_memory() for _ in range(2)]
    m_pin = [m_h.clone().pin_memory() for _ in range(2)]
    wav_pin = [torch.empty((B, 1, 256 * T), dtype=torch.float32).pin_memory() for _ in range(2)]
    z_p, mask = zp_h.to(dev), m_h.to(dev)
    for _ in range(3):
        eng.flow_decode(z_p, mask, want_z=False, want_mb=False, want_spec=False)
    graph, _ = eng.capture_flow_decode(z_p, mask)
    hs = HostStream(eng, depth=2, fused=True)
    for i in range(2):   # builds both slots (and their graphs)
        hs.submit(zp_pin[i], m_pin[i], wav_pin[i])
    hs.drain()

    def loop(n, do_in, do_out):
        for i in range(n):
            s = hs.slots[i % 2]
            with torch.cuda.stream(hs.s_in):
                hs.s_in.wait_event(s["ev_cmp"])
                if do_in:
                    s["zp"].copy_(zp_pin[i % 2], non_blocking=True)
                    s["m"].copy_(m_pin[i % 2], non_blocking=True)
                s["ev_in"].record(hs.s_in)
            with torch.cuda.stream(hs.s_cmp):
                hs.s_cmp.wait_event(s["ev_in"])
                hs.s_cmp.wait_event(s["ev_out"])
                s["graph"].replay()
                s["ev_cmp"].record(hs.s_cmp)
            with torch.cuda.stream(hs.s_out):
                hs.s_out.wait_event(s["ev_cmp"])
                if do_out:
                    wav_pin[i % 2].copy_(s["wav"], non_blocking=True)
                s["ev_out"].record(hs.s_out)

    def full(n):
        for i in range(n):
            hs.submit(zp_pin[i % 2], m_pin[i % 2], wav_pin[i % 2])

    def resident(n):
        with torch.cuda.stream(hs.s_cmp):
            for _ in range(n):
                graph.replay()

    variants = [("resident", resident), ("slots", lambda n: loop(n, False, False)), ("in", lambda n: loop(n, True, False)),
                ("out", lambda n: loop(n, False, True)), ("full", full)]
    lines = []
    for rep in range(args.reps):
        for name, fn in variants:
            hs.drain()
            torch.cuda.synchronize()
            time.sleep(2.0)
            fn(3)
            hs.drain()
            torch.cuda.synchronize()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            cur = torch.cuda.current_stream()
            x0.record(cur)
            for st in (hs.s_in, hs.s_cmp, hs.s_out):
                st.wait_event(x0)
            fn(args.steps)
            for st in (hs.s_in, hs.s_cmp, hs.s_out):
                cur.wait_stream(st)
            x1.record(cur)
            torch.cuda.synchronize()
            lines.append(f"rep {rep}  {name:9s} {x0.elapsed_time(x1) / args.steps:8.3f} ms per step")
            print(lines[-1], flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write(f"e2e_probe: {args.config} B={B} T={T}, {args.steps} steps per variant, each from an idle GPU (2 s) + 3 warm-up steps\n")
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
