#!/usr/bin/env python
"""First-light probe for the GPU box: runs each precision / tcgen05 staging mode in its own subprocess (a device
trap poisons the CUDA context) and prints one line per mode.  Results go to gpurun_out/probe.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

MODES = [
    ("fp32", 0), ("bf16_simt", 4), ("bf16_tc", 0), ("tf32_simt", 4), ("tf32_tc", 0),
]


def child(mode, flags, case):
    import torch
    import mbistft_oracle as orc
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import load_case
    from mb_istft_vits_b200 import Engine
    cfg, sd, t, meta = load_case(case)
    prec = mode.split("_")[0]
    eng = Engine(cfg, sd, precision=prec, flags=flags)
    g = t.get("g")
    g = g.cuda() if g is not None else None
    z = eng.flow_reverse(t["z_p"].cuda(), t["mask"].cuda(), g)
    wav, o_mb, spec, phase = eng.decode((t["z"] * t["mask"]).cuda(), g)
    torch.cuda.synchronize()
    res = dict(mode=mode, case=case,
               z_err=float((z.cpu() - t["z"]).abs().max()),
               wav_err=orc.max_abs_over_peak(wav.cpu(), t["o"]), wav_snr=orc.snr_db(wav.cpu(), t["o"]),
               spec_err=orc.max_abs_over_peak(spec.cpu(), t["spec"]),
               phase_err=float((phase.cpu() - t["phase"]).abs().max()))
    if o_mb is not None:
        res["omb_err"] = orc.max_abs_over_peak(o_mb.cpu(), t["o_mb"])
    print("RESULT " + json.dumps(res))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(sys.argv[2], int(sys.argv[3]), sys.argv[4])
        return
    cases = sys.argv[1:] or ["mini_mb", "mb"]
    out = []
    for case in cases:
        for mode, flags in MODES:
            try:
                p = subprocess.run([sys.executable, __file__, "child", mode, str(flags), case], capture_output=True,
                                   text=True, timeout=180)
                lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                if lines:
                    r = json.loads(lines[-1][7:])
                else:
                    r = dict(mode=mode, case=case, error=(p.stderr[-600:] + p.stdout[-300:]))
            except subprocess.TimeoutExpired:
                r = dict(mode=mode, case=case, error="timeout")
            out.append(r)
            print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
