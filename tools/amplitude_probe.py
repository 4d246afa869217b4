#!/usr/bin/env python
"""How far does the bf16 path with an fp16 residual stream hold the 40 dB bar as activations grow?

Sweeps the weight-gain scale g_scale of the synthetic weights (activations grow roughly like g_scale^depth) on the
multi-band config and prints, per scale: the peak of the fp32 residual stream as seen by the oracle, and the waveform
SNR of bf16 + fp16 stream / bf16 + fp32 stream / fp16 single stream against the CPU oracle.
    python tools/amplitude_probe.py [--out file]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import mbistft_oracle as orc  # noqa: E402
from mb_istft_vits_b200 import Engine, get_config, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    cfg = get_config("ljs_mb_istft_vits")
    lines = ["g_scale   wav peak   logits max   bf16+fp16 stream   bf16+fp32 stream   fp16 single stream   (waveform SNR dB vs fp32 oracle)"]
    for gs in (1.0, 1.7, 2.0, 2.5, 3.0, 3.5, 4.0):
        sd = synth.make_state_dict(cfg, seed=1234, g_scale=gs)
        z_p, mask, _ = synth.make_latents(cfg, 1, 48, seed=9)
        z_ref, (o_ref, _, _, _) = orc.flow_decode(sd, cfg, z_p, mask)
        logits = orc.decoder_logits(sd, cfg, z_ref * mask)
        row = [gs, float(o_ref.abs().max()), float(logits.abs().max())]
        for prec, res in (("bf16", "fp16"), ("bf16", "fp32"), ("fp16", "fp16")):
            eng = Engine(cfg, sd, precision=prec, residual=res)
            wav = eng.flow_decode(z_p.cuda(), mask.cuda())[1].cpu()
            row.append(orc.snr_db(wav, o_ref))
            eng.close()
        lines.append("%6.2f %10.3g %12.3g %18.1f %18.1f %20.1f" % tuple(row))
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
